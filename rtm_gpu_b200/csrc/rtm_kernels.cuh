// rtm_kernels.cuh -- sm_100a device code of the RTM time loop.
//
// One launch per time step.  A launch covers a batch of shots (1-D grid) and, per
// shot, two kinds of CTA:
//   * interior tiles (fast path): TMA-staged halo tile of the current wavefield in shared
//     memory, each thread walks 4 x-adjacent cells (one float4) down NR rows; x neighbours = the
//     row segment as float4 shared loads, z neighbours = one float4 shared load per row offset
//     (a register column over z was measured slower at this occupancy, profiles/README.md),
//     coalesced float4 global loads/stores for the other streams;
//   * ring tiles: the N2-wide hybrid absorbing ring.  They evaluate the two-way update on
//     the ring plus a one-cell halo, apply the one-way solution and the blend, and move
//     the boundary strips (save in the forward pass, restore in the backward pass).
// All CTAs read only time slots k-1/k-2 and write disjoint cells of slot k, so there is no
// ordering requirement inside a launch and buffers rotate by pointer.
//
// FP32 contract.  Results must equal the reference's CUDA build bit for bit, so every
// operation that nvcc could contract or reorder is written with an explicit intrinsic, in
// the order the reference's sm_100a SASS has (DESIGN.md "FP contract").
// Reference kernels restated here (kernel.cu): Add/Add_Con :46-114, Hybrid1/2/3 :116-208,
// Equal :18-45, BKEqual :222-245, BKAdd_EFF(_Con) :246-320, BKAdd(_Con) :339-418,
// BKHybrid1/2 :421-487, Rel_Compen/Rel_NonCompen :489-517; Deliver/Deliver_EFF :210-221,
// :323-337 disappear (pointer rotation).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtmk {

constexpr int kThreads = 256;  // 8 warps per CTA
constexpr int kTX      = 128;  // interior tile width  (32 lanes x float4)
// rows per thread (tile height = 8 warps x rows); tuned per kernel on the B200 (profiles/)
#ifndef RTM_NR_F
#define RTM_NR_F 4
#endif
#ifndef RTM_NR_B
#define RTM_NR_B 2
#endif
constexpr int kWarps   = kThreads / 32;
#ifndef RTM_RING_TX
#define RTM_RING_TX 126
#endif
constexpr int kRingTX  = RTM_RING_TX;  // ring tile extent along the band: grown by one cell on both sides = 32 float4 groups
constexpr int kMaxR    = 16;
constexpr int kSliceBins = 768;  // velocity bins whose coefficient rows a tile can stage in shared memory
__host__ __device__ constexpr int slice_bytes(int RP) { return RP >= 8 ? 24576 : kSliceBins * (RP + 5) * 4; }

enum SumKind { SUM_FLOAT = 0, SUM_DOUBLE = 1 };

// Division by a run-time constant as multiply-high + shift (0 <= n < 2^31): the CTA prologues
// divide block indices by tile counts, which costs ~10 % of the forward kernel's instructions as
// real integer divisions.
struct FastDiv {
    unsigned mul, shr;
    int      d;
};
__host__ inline FastDiv make_fastdiv(int d)
{
    FastDiv f{0u, 0u, d < 1 ? 1 : d};
    if (f.d > 1) {
        int cl = 0;
        while ((1 << cl) < f.d) ++cl;
        const int p = 31 + cl;
        f.mul = (unsigned)(((1ull << p) + (unsigned long long)f.d - 1) / (unsigned long long)f.d);
        f.shr = (unsigned)(p - 32);
    }
    return f;
}
__host__ __device__ __forceinline__ int fast_div(int n, const FastDiv& f)
{
#ifdef __CUDA_ARCH__
    return f.d == 1 ? n : (int)(__umulhi((unsigned)n, f.mul) >> f.shr);
#else   // host (tests): the same multiply-high in 64-bit arithmetic
    return f.d == 1 ? n : (int)((unsigned)(((unsigned long long)(unsigned)n * f.mul) >> 32) >> f.shr);
#endif
}

struct Geo {
    int   NZ, NX, N2, mod_NZ, mod_NX;
    int   pitch, padL;       // internal row pitch (floats) and left pad: cell (z,x) at z*pitch+padL+x
    long long shot_stride;   // floats between shots of one field buffer
    int   nfdmax;            // strip width; Taylor radius
    int   mmax;              // longest operator actually present in the model (<= nfdmax): stencil halo
    int   NT;
    int   iLSTE, iCompen;
    float tao2, h2, taoh, taoh2, hzx2_1, vmin, dv;
    float tao, h;            // for the corner coefficient r = v*tao/h
    double A;                // 1.0 + (double)hzx2_1
    int   s_l, s_r, s_z, ds, n;
    // tiling
    int   ntx, ntz_f, ntz_b; // interior tiles (forward / backward kernels use different tile heights)
    int   nband, nside;      // ring tiles per band (top/bottom) and per side (left/right)
    FastDiv fd_ntx, fd_nring;  // dividers by ntx and by the ring tile count 2*nband+2*nside
    // operator
    const float* c;          // LS: packed table; TE: unused
    const int*   Index;
    float  cTE[kMaxR + 1];   // Taylor coefficients
    double cc0TE;            // (1+hzx2_1)*c[0] in double
    float  cc0f;             // same as float when exactly representable (cc0_exact), else unused
    int    cc0_exact;        // TE: cc0TE is a float; LS: 1+hzx2_1 is a power of two (A*c0 exact in float)
    const float* avel;       // [NZ][pitch] a = ((v*v)*tao2)*h2, precomputed with the kernel's own rounding
    // adaptive operator, interior fast path: per-cell velocity bin (2 B/cell side array, replaces
    // the per-step IEEE division + double add + two dependent Index loads) and per-tile bin
    // ranges so that a tile can stage its slice of Index/c in shared memory
    const unsigned short* bins;   // [NZ][pitch]
    const int2* tile_bins_f;      // [ntz_f*ntx] (min bin, max bin) of each forward interior tile
    const int2* tile_bins_b;      // [ntz_b*ntx]
    const int2* tile_bins_b2;     // [ntz_b*ntx] same tiles grown by the operator radius (two-step kernel)
    const float* v;          // [NZ][pitch], shared by all shots
    // adaptive operator no longer than 4 everywhere (streaming kernels): every bin's coefficients zero-padded to
    // 8 floats (one 32-byte sector per cell, read through L1) and its length, in global memory
    const float* ls_rows;    // [nvel][8]
    const int*   ls_len;     // [nvel]
    int          ls_nbins;
    float  w[65];            // blend weights l/N2
};

struct Strips {              // boundary strips, all shots and time slots (64-bit offsets)
    float *up, *dw;          // [S][NT][nfdmax][mod_NX]   rows z = N2-nfdmax+j / NZ-N2+j
    float *lf, *rt;          // [S][NT][mod_NZ][nfdmax]   cols x = N2-nfdmax+j / NX-N2+j
};

// ------------------------------------------------------------------ PTX helpers (TMA)
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// The waiting thread may be suspended by the hardware until the phase completes or the time hint (ns) runs out; with a
// generous hint a waiting warp polls a handful of times instead of spinning (round 2: the spinning warps of the streaming
// kernel issued 30 % of its instructions and cost power under the 1 kW cap).
#ifndef RTM_MBAR_HINT_NS
#define RTM_MBAR_HINT_NS 20000
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"((uint32_t)RTM_MBAR_HINT_NS)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int x, int z, int s)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"((uint64_t)map),
                 "r"(x), "r"(z), "r"(s)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int x, int z, int s)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(x), "r"(z), "r"(s)
        : "memory");
}

// 4-byte asynchronous global -> shared copies (ring tiles: every thread has all its halo
// loads in flight at once instead of one exposed HBM latency per element)
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ exact arithmetic
// final sum of the two-way update, a = ((v*v)*tao2)*h2
__device__ __forceinline__ float finish_float(float a, float w1, float p1, float p0)
{
    return __fmaf_rn(a, w1, __fsub_rn(__fadd_rn(p1, p1), p0));  // Add, BKAdd, BKAdd_Con
}
__device__ __forceinline__ float finish_double(float a, float w1, float p1, float p0)
{
    // Add_Con, BKAdd_EFF, BKAdd_EFF_Con: 2.0*P1 - P0 + (float)(a*w1) in double
    // (2.0*P1 is exact in double, so fma(2, P1, -P0) rounds once exactly like (P1+P1)-P0)
    return __double2float_rn(
        __dadd_rn(__fma_rn((double)p1, 2.0, -(double)p0), (double)__fmul_rn(a, w1)));
}
__device__ __forceinline__ float vel_factor(const Geo& G, float v)
{
    return __fmul_rn(__fmul_rn(__fmul_rn(v, v), G.tao2), G.h2);
}
// first term of the stencil sum, w1 = (float)(((1.0+hzx2_1)*c0)*P1) evaluated in double by the
// reference.  When (1+hzx2_1)*c0 is exactly a float the double product is exact, so a single
// float multiply rounds identically (no conversions needed).
__device__ __forceinline__ float w1_first_te(const Geo& G, float p1)
{
    return G.cc0_exact ? __fmul_rn(G.cc0f, p1) : __double2float_rn(__dmul_rn(G.cc0TE, (double)p1));
}
__device__ __forceinline__ float w1_first_ls(const Geo& G, float c0, float p1)
{
    return G.cc0_exact ? __fmul_rn(__fmul_rn(G.cc0f, c0), p1)   // cc0f = 1+hzx2_1 (a power of two)
                       : __double2float_rn(__dmul_rn(__dmul_rn(G.A, (double)c0), (double)p1));
}
// The same for the four cells of a float4 group, with the (uniform) test outside the cell loop: inside it the compiler
// if-converts the double path, and its two conversions, two double multiplies and the conversion back are then issued
// predicated-off for every cell (round 2, ncu source counters of the streaming kernel: 7 % of its instructions, 15 % of
// its stall samples).  Used by the streaming kernels only: in the tile kernels the real branch cost more than it saved
// (spills in the row loops: forward step at radius 8 100.7 -> 122.2 us, profiles/README.md).
__device__ __forceinline__ void w1_first_ls4(const Geo& G, const float (&c0)[4], const float (&p1)[4], float (&w1)[4])
{
    if (G.cc0_exact) {
#pragma unroll
        for (int q = 0; q < 4; ++q) w1[q] = __fmul_rn(__fmul_rn(G.cc0f, c0[q]), p1[q]);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) w1[q] = __double2float_rn(__dmul_rn(__dmul_rn(G.A, (double)c0[q]), (double)p1[q]));
    }
}

// data cell test of BKAdd (:349-353): row s_z, s_l <= x <= s_r, (x-s_l)%ds==0
__device__ __forceinline__ int data_index(const Geo& G, int z, int x)
{
    if (z != G.s_z || x < G.s_l || x > G.s_r) return -1;
    const int d = x - G.s_l;
    return (d % G.ds == 0) ? d / G.ds : -1;
}

}  // namespace rtmk

// =====================================================================================
// Interior fast path
// =====================================================================================
namespace rtmk {

// Shared-memory tile of the current field for one interior tile: rows [z0-RP, z0+TZ+RP),
// columns [x0-RP, x0+kTX+RP), dense, pitch kTX+2RP floats, written by one TMA box copy.
template <int RP, int NR> struct Tile {
    static constexpr int TZ    = kWarps * NR;
    static constexpr int SP    = kTX + 2 * RP;
    static constexpr int ROWS  = TZ + 2 * RP;
    static constexpr int BYTES = SP * ROWS * 4;
};

__device__ __forceinline__ void unpack(const float4& a, float (&o)[4])
{
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
}

// Operator table of the adaptive path for one interior tile.  Normal case ("rows"): the
// coefficient rows of the tile's bins [bmin,bmax] are staged in shared memory, zero-padded to a
// fixed stride of RP+4 floats (16-byte aligned, so a cell fetches its coefficients with float4
// loads and no Index lookups), next to one length per bin.  Fallback (slice larger than the
// shared-memory budget, e.g. a salt flank crossing the tile): the packed global tables.
struct LsTable {
    const float* rows;   // shared: [(bmax-bmin+1)][RP+4]
    const int*   len;    // shared: [(bmax-bmin+1)]
    const int*   ip;     // fallback: global Index
    const float* cp;     // fallback: global c
    int bmin, bmax;
    bool staged;
};

// Called by all threads of the CTA; the caller synchronises afterwards.
template <int RP>
__device__ __forceinline__ LsTable ls_stage_slice(const Geo& G, int2 tb, float* smem_words)
{
    constexpr int STRIDE = RP + 4;
    LsTable T;
    T.bmin = tb.x; T.bmax = tb.y;
    T.ip = G.Index; T.cp = G.c;
    const int nb = tb.y - tb.x + 1;
    T.rows = smem_words;
    T.len  = nullptr;
    T.staged = false;
    if (RP >= 8) {
        // long operators: padded rows would mostly hold zeros; stage the packed slice instead
        // (Index[bmin..bmax+1] and c[Index[bmin]..Index[bmax+1])), addressed like the global tables
        const int nI = nb + 1;
        const int cBeg = __ldg(G.Index + tb.x), nC = __ldg(G.Index + tb.y + 1) - cBeg;
        if (nI + nC <= slice_bytes(RP) / 4) {
            int*   sI = reinterpret_cast<int*>(smem_words);
            float* sC = smem_words + nI;
            for (int i = threadIdx.x; i < nI; i += kThreads) sI[i] = __ldg(G.Index + tb.x + i);
            for (int i = threadIdx.x; i < nC; i += kThreads) sC[i] = __ldg(G.c + cBeg + i);
            T.ip = sI - tb.x;
            T.cp = sC - cBeg;
        }
        return T;
    }
    T.staged = nb <= kSliceBins;
    T.len  = reinterpret_cast<const int*>(smem_words + nb * STRIDE);
    if (T.staged) {
        int* sl = reinterpret_cast<int*>(smem_words + nb * STRIDE);
        for (int r = threadIdx.x; r < nb; r += kThreads) {  // one bin (row) per thread
            const int top = __ldg(G.Index + tb.x + r), M = __ldg(G.Index + tb.x + r + 1) - top - 1;
            float* row = smem_words + r * STRIDE;
#pragma unroll
            for (int g = 0; g < STRIDE / 4; ++g) {
                float4 v4;
                v4.x = (4 * g + 0 <= M) ? __ldg(G.c + top + 4 * g + 0) : 0.0f;
                v4.y = (4 * g + 1 <= M) ? __ldg(G.c + top + 4 * g + 1) : 0.0f;
                v4.z = (4 * g + 2 <= M) ? __ldg(G.c + top + 4 * g + 2) : 0.0f;
                v4.w = (4 * g + 3 <= M) ? __ldg(G.c + top + 4 * g + 3) : 0.0f;
                *reinterpret_cast<float4*>(row + 4 * g) = v4;
            }
            sl[r] = M;
        }
    }
    return T;
}

// Stencil sums w1 of four x-adjacent cells of one row.  `sc` points at the first of the four
// cells inside the shared tile.  x neighbours: the row segment [x-RP, x+3+RP] as float4 shared
// loads; z neighbours: one float4 shared load per row offset (all four cells share it), which
// keeps the register footprint small enough for 3-4 resident CTAs per SM.
// M <= RP is the uniform length of the Taylor operator; the adaptive operator brings a length
// and table offset per cell (from the cell's velocity bin).
// SPT: pitch of the shared tile (floats); 0 = run-time pitch `spr` (ring tiles).
// RTM_LS_UNIFORM=1: adaptive operators of the radius classes 8, 12, 16 (packed-table path): the term loop runs to the
// longest operator of the WARP.  (The kernels of the radius-4 class keep per-lane bounds: with the uniform bound and selects
// instead of branches there, the forward step of the mixed-length C5 model got slower, 89.1 -> 93.5 us.)
#ifndef RTM_LS_UNIFORM
#define RTM_LS_UNIFORM 1
#endif
template <int RP, bool LS, int SPT = kTX + 2 * RP>
__device__ __forceinline__ void stencil_row(const Geo& G, const float* sc, int M, const LsTable& T,
                                            uint2 bins4, float (&w1)[4], float (&p1)[4], int spr = 0)
{
    const int SP = SPT ? SPT : spr;
    float xr[4 + 2 * RP];  // columns x-RP .. x+3+RP of this row
#pragma unroll
    for (int g = 0; g < (4 + 2 * RP) / 4; ++g) {
        const float4 t4 = *reinterpret_cast<const float4*>(sc - RP + 4 * g);
        xr[4 * g + 0] = t4.x; xr[4 * g + 1] = t4.y; xr[4 * g + 2] = t4.z; xr[4 * g + 3] = t4.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) p1[q] = xr[RP + q];

    if (LS) {
        const int b4[4] = {(int)(bins4.x & 0xffffu), (int)(bins4.x >> 16), (int)(bins4.y & 0xffffu), (int)(bins4.y >> 16)};
        int Mc[4], Mx = 0;
        if (T.staged) {
            constexpr int STRIDE = RP + 4;
            const float* row[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = min(max(b4[q], T.bmin), T.bmax) - T.bmin;  // (cells of a partial group past the interior)
                row[q] = T.rows + r * STRIDE;
                Mc[q]  = T.len[r];
                Mx     = max(Mx, Mc[q]);
            }
#pragma unroll
            for (int g = 0; g <= RP / 4; ++g) {  // coefficients 4g .. 4g+3 of the four cells
                if (4 * g > Mx) break;
                float cg[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) unpack(*reinterpret_cast<const float4*>(row[q] + 4 * g), cg[q]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int l = 4 * g + j;
                    if (l > RP) break;
                    if (l == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) w1[q] = w1_first_ls(G, cg[q][0], p1[q]);
                    } else if (l <= Mx) {
                        float zm[4], zp[4];
                        unpack(*reinterpret_cast<const float4*>(sc - l * SP), zm);
                        unpack(*reinterpret_cast<const float4*>(sc + l * SP), zp);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (l <= Mc[q]) {
                                const float s = __fadd_rn(zm[q], zp[q]);
                                const float t = __fmaf_rn(s, G.hzx2_1, xr[RP + q - l]);
                                const float u = __fadd_rn(t, xr[RP + q + l]);
                                w1[q]         = __fmaf_rn(cg[q][j], u, w1[q]);
                            }
                        }
                    }
                }
            }
        } else {
            const float* cq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int b   = min(max(b4[q], T.bmin), T.bmax);
                const int top = T.ip[b];
                Mc[q] = T.ip[b + 1] - top - 1;
                cq[q] = T.cp + top;
                Mx    = max(Mx, Mc[q]);
                w1[q] = w1_first_ls(G, cq[q][0], p1[q]);
            }
            // the term loop runs to the longest operator of the (converged part of the) warp: a uniform bound needs no
            // divergence handling around every term (measured, profiles/r2_c18_*: forced radius 12 +7.7 %, 8 +3.5 %, 6 +2.8 %)
            if (RTM_LS_UNIFORM && RP >= 8) Mx = __reduce_max_sync(__activemask(), Mx);
#pragma unroll
            for (int l = 1; l <= RP; ++l) {
                if (l <= Mx) {
                    float zm[4], zp[4];
                    unpack(*reinterpret_cast<const float4*>(sc - l * SP), zm);
                    unpack(*reinterpret_cast<const float4*>(sc + l * SP), zp);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (l <= Mc[q]) {   // (the coefficient is only there for l <= Mc: a real test)
                            const float s = __fadd_rn(zm[q], zp[q]);
                            const float t = __fmaf_rn(s, G.hzx2_1, xr[RP + q - l]);
                            const float u = __fadd_rn(t, xr[RP + q + l]);
                            w1[q]         = __fmaf_rn(cq[q][l], u, w1[q]);
                        }
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) w1[q] = w1_first_te(G, p1[q]);
        auto term = [&](int l) {
            float zm[4], zp[4];
            unpack(*reinterpret_cast<const float4*>(sc - l * SP), zm);
            unpack(*reinterpret_cast<const float4*>(sc + l * SP), zp);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float s = __fadd_rn(zm[q], zp[q]);
                const float t = __fmaf_rn(s, G.hzx2_1, xr[RP + q - l]);
                const float u = __fadd_rn(t, xr[RP + q + l]);
                w1[q]         = __fmaf_rn(G.cTE[l], u, w1[q]);
            }
        };
        if (M == RP) {  // common case (radius 4, 8, 12, 16): straight-line code
#pragma unroll
            for (int l = 1; l <= RP; ++l) term(l);
        } else {
#pragma unroll
            for (int l = 1; l <= RP; ++l)
                if (l <= M) term(l);
        }
    }
}

// The adaptive operator of the streaming kernels (every length <= 4): a cell's coefficients are one zero-padded row of 8
// floats in a global table (Geo::ls_rows), its length one int (Geo::ls_len).  The table offsets and lengths of the four
// cells of a float4 group are looked up once (ls_cells) and shared by every field evaluated at those cells; the
// arithmetic per cell is stencil_row<RP, true>'s, term by term (terms beyond a cell's own length are skipped, not
// multiplied by the zero padding).
struct LsCells { int off[4]; int Mc[4]; int Mx; };   // Mx: longest operator among the cells of the WARP (uniform)
// (called by all 32 lanes of a warp)
__device__ __forceinline__ LsCells ls_cells(const Geo& G, uint2 bins4)
{
    LsCells L;
    const int b4[4] = {(int)(bins4.x & 0xffffu), (int)(bins4.x >> 16), (int)(bins4.y & 0xffffu), (int)(bins4.y >> 16)};
    L.Mx = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int b = min(b4[q], G.ls_nbins - 1);
        L.off[q] = b * 8;
        L.Mc[q]  = __ldg(G.ls_len + b);
        L.Mx     = max(L.Mx, L.Mc[q]);
    }
    L.Mx = __reduce_max_sync(0xffffffffu, L.Mx);   // uniform loop bounds: no divergence handling around the terms
    return L;
}
template <int SPT>
__device__ __forceinline__ void stencil_row_ls4(const Geo& G, const float* sc, const LsCells& L, float (&w1)[4], float (&p1)[4])
{
    constexpr int RP = 4, SP = SPT;
    float xr[4 + 2 * RP];
#pragma unroll
    for (int g = 0; g < (4 + 2 * RP) / 4; ++g) {
        const float4 t4 = *reinterpret_cast<const float4*>(sc - RP + 4 * g);
        xr[4 * g + 0] = t4.x; xr[4 * g + 1] = t4.y; xr[4 * g + 2] = t4.z; xr[4 * g + 3] = t4.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) p1[q] = xr[RP + q];
    float cg[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) unpack(__ldg(reinterpret_cast<const float4*>(G.ls_rows + L.off[q])), cg[q]);
    {
        const float c0[4] = {cg[0][0], cg[1][0], cg[2][0], cg[3][0]};
        w1_first_ls4(G, c0, p1, w1);
    }
    auto term = [&](int l, const float (&cl)[4]) {
        float zm[4], zp[4];
        unpack(*reinterpret_cast<const float4*>(sc - l * SP), zm);
        unpack(*reinterpret_cast<const float4*>(sc + l * SP), zp);
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // (a cell shorter than l keeps its sum: selected, not branched around)
            const float s = __fadd_rn(zm[q], zp[q]);
            const float t = __fmaf_rn(s, G.hzx2_1, xr[RP + q - l]);
            const float u = __fadd_rn(t, xr[RP + q + l]);
            const float w = __fmaf_rn(cl[q], u, w1[q]);
            w1[q]         = l <= L.Mc[q] ? w : w1[q];
        }
    };
#pragma unroll
    for (int l = 1; l <= 3; ++l) {
        if (l <= L.Mx) {
            const float cl[4] = {cg[0][l], cg[1][l], cg[2][l], cg[3][l]};
            term(l, cl);
        }
    }
    if (L.Mx >= 4) {
        float c4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) c4[q] = __ldg(G.ls_rows + L.off[q] + 4);
        term(4, c4);
    }
}

// =====================================================================================
// Ring tiles (hybrid absorbing boundary)
// =====================================================================================
// One ring tile.  Output rectangle [za,zb) x [xa,xb) lies in the ring; the two-way update is
// evaluated on that rectangle grown by one cell (clipped to the array), because the one-way
// formulas need the UNBLENDED two-way values of neighbours (Hybrid1 reads DFW2 before
// Hybrid2 blends it).  Returns the blended value through `emit(z, x, value)`.
struct RingRect { int za, zb, xa, xb; };

__device__ __forceinline__ RingRect ring_rect(const Geo& G, int tile)
{
    RingRect r;
    const int N2 = G.N2;
    if (tile < 2 * G.nband) {  // top / bottom band: all columns
        const bool top = tile < G.nband;
        const int  i   = top ? tile : tile - G.nband;
        r.za = top ? 0 : G.NZ - N2;
        r.zb = r.za + N2;
        r.xa = i * kRingTX;
        r.xb = min(r.xa + kRingTX, G.NX);
    } else {  // left / right side: interior rows
        tile -= 2 * G.nband;
        const bool left = tile < G.nside;
        const int  i    = left ? tile : tile - G.nside;
        r.xa = left ? 0 : G.NX - N2;
        r.xb = r.xa + N2;
        r.za = N2 + i * kRingTX;
        r.zb = min(r.za + kRingTX, G.NZ - N2);
    }
    return r;
}

// shared memory needed by a ring tile (floats): current field with the stencil halo (rows: R,
// columns: RP), previous field, velocity factor and two-way result on the compute rectangle padded to float4 groups,
// and three values per output cell staged for the one-way phase (velocity at the cell, velocity
// of the edge formula's tangential term, one caller-defined value)
__host__ __device__ inline int ring_smem_floats(int N2, int R, int RP)
{
    const int wb = (kRingTX + 2 + 3) / 4 * 4, ws = (N2 + 2 + 3) / 4 * 4;  // band / side compute widths
    const int band = (N2 + 2 + 2 * R) * (wb + 2 * RP) + 3 * (N2 + 2) * wb;
    const int side = (kRingTX + 2 + 2 * R) * (ws + 2 * RP) + 3 * (kRingTX + 2) * ws;
    return (band > side ? band : side) + 3 * N2 * kRingTX;
}

// Visit the cells of an h x w rectangle with the CTA's 256 threads without integer division:
// wide rows (band tiles) go row-per-warp, narrow rows (side tiles) 16 or 32 columns per row slot.
template <class F> __device__ __forceinline__ void for_cells(int h, int w, F f)
{
    if (w > 32) {  // (row, 32-column chunk) pairs dealt round-robin to the warps
        const int nch = (w + 31) >> 5;
        const int dq = kWarps / nch, dr = kWarps - dq * nch;   // task index advances by kWarps = dq rows + dr chunks
        const int wp = threadIdx.x >> 5;
        int r = wp / nch, ch = wp - r * nch;
        while (r < h) {
            const int c = (ch << 5) + (threadIdx.x & 31);
            if (c < w) f(r, c);
            r += dq; ch += dr;
            if (ch >= nch) { ch -= nch; ++r; }
        }
    } else if (w > 16) {
        const int c = threadIdx.x & 31;
        if (c < w)
            for (int r = threadIdx.x >> 5; r < h; r += kThreads / 32) f(r, c);
    } else {
        const int c = threadIdx.x & 15;
        if (c < w)
            for (int r = threadIdx.x >> 4; r < h; r += kThreads / 16) f(r, c);
    }
}

// `stage(z, x)` names one more global float per output cell (or nullptr) that is fetched together
// with the halo and handed to `emit(z, x, value, staged)` (backward pass: the boundary-strip value
// that BKEqual restores at that cell).
template <int RP, bool LS, class Stage, class Emit>
__device__ __forceinline__ void ring_tile(const Geo& G, int tile, const float* __restrict__ P1,
                                          const float* __restrict__ P0, int sum_kind,
                                          bool inject, int r_u, int r_x, float wavelet,
                                          const float* __restrict__ seis_row, float* smem, Stage stage, Emit emit)
{
    const RingRect o  = ring_rect(G, tile);
    const int      NZ = G.NZ, NX = G.NX, N2 = G.N2, R = G.mmax, pitch = G.pitch;
    const float*   V  = G.v + G.padL;  // cell (z,x) of the model at V[z*pitch+x]
    // compute rectangle = output grown by 1, clipped; padded to whole float4 groups in x
    const int cza = max(o.za - 1, 0), czb = min(o.zb + 1, NZ);
    const int cxa = max(o.xa - 1, 0), cxb = min(o.xb + 1, NX);
    const int ch = czb - cza, cw = cxb - cxa;
    const int NG = (cw + 3) >> 2, CW = 4 * NG;
    const int SP = CW + 2 * RP;                // pitch of the current-field tile
    float* s1 = smem;                          // (ch+2R) x SP: row 0 = z cza-R, column 0 = x cxa-RP
    float* s0 = s1 + (ch + 2 * R) * SP;        // ch x CW   previous field
    float* s2 = s0 + ch * CW;                  // ch x CW   unblended two-way result
    const int oh = o.zb - o.za, ow = o.xb - o.xa;
    float* sAv = s2 + ch * CW;                 // ch x CW   a = ((v*v)*tao2)*h2 of the two-way update
    float* sVb = sAv + ch * CW;                // oh x ow   velocity at the cell
    float* sVq = sVb + oh * ow;                // oh x ow   velocity of the edge formula's taoh2 term (Q2)
    float* sAx = sVq + oh * ow;                // oh x ow   caller's staged value

    // all global reads of the tile as asynchronous copies: every thread has its loads in flight
    // together, one wait for the lot
    for_cells(ch + 2 * R, SP, [&](int r, int cidx) {
        int gz = cza - R + r, gx = cxa - RP + cidx;
        if (gz < 0) gz = -gz;                      // mirror about the array edge (:65-68)
        if (gz >= NZ) gz = 2 * NZ - 2 - gz;
        if (gx < 0) gx = -gx;
        if (gx >= NX) gx = 2 * NX - 2 - gx;
        gx = min(max(gx, 0), NX - 1);              // (padding columns past the operator's reach)
        cp_async4(s1 + r * SP + cidx, P1 + (size_t)gz * pitch + gx);
    });
    for_cells(ch, CW, [&](int r, int cidx) {
        const size_t cell = (size_t)(cza + r) * pitch + min(cxa + cidx, NX - 1);
        cp_async4(s0 + r * CW + cidx, P0 + cell);
        cp_async4(sAv + r * CW + cidx, G.avel + G.padL + cell);
    });
    for_cells(oh, ow, [&](int oz, int ox) {
        const int z = o.za + oz, x = o.xa + ox;
        const int dz = min(z, NZ - 1 - z), dx = min(x, NX - 1 - x);
        const int a  = min(dz, dx);
        cp_async4(sVb + oz * ow + ox, V + (size_t)z * pitch + x);
        int fz = a, fx = x;              // top :124 / bottom :132 (and the corner cells, which do not use it)
        if (dz >= dx) {                  // left :128 / right :136: flat index (N2-l)*NX + row
            fx = z;                      // (a*NX + z) / NX and % NX without the division
            while (fx >= NX) { fx -= NX; ++fz; }
        }
        cp_async4(sVq + oz * ow + ox, V + (size_t)fz * pitch + fx);
        const float* g = stage(z, x);
        if (g) cp_async4(sAx + oz * ow + ox, g);
    });
    cp_async_wait_all();
    __syncthreads();

    // two-way update of the compute rectangle: one float4 group (4 cells) per thread and pass,
    // the interior tiles' row stencil on the shared tile (global operator tables for the adaptive one)
    {
        LsTable T{};
        T.ip = G.Index; T.cp = G.c; T.bmin = 0; T.bmax = 0xffff; T.staged = false;
        const unsigned short* BN = G.bins + G.padL;
        int sh = 0;
        while ((1 << sh) < NG) ++sh;               // groups per row rounded up to a power of two
        for (int it = threadIdx.x; it < (ch << sh); it += kThreads) {
            const int lz = it >> sh, g = it & ((1 << sh) - 1);
            if (g >= NG) continue;
            const int z = cza + lz, x = cxa + 4 * g;
            const size_t row = (size_t)z * pitch;
            int xq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) xq[q] = min(x + q, NX - 1);
            uint2 b4 = make_uint2(0u, 0u);
            if (LS) {
                b4.x = (unsigned)__ldg(BN + row + xq[0]) | ((unsigned)__ldg(BN + row + xq[1]) << 16);
                b4.y = (unsigned)__ldg(BN + row + xq[2]) | ((unsigned)__ldg(BN + row + xq[3]) << 16);
            }
            float w1[4], p1[4], p0[4], val[4];
            stencil_row<RP, LS, 0>(G, s1 + (lz + R) * SP + 4 * g + RP, G.nfdmax, T, b4, w1, p1, SP);
            unpack(*reinterpret_cast<const float4*>(s0 + lz * CW + 4 * g), p0);
            float avq[4];
            unpack(*reinterpret_cast<const float4*>(sAv + lz * CW + 4 * g), avq);  // ((v*v)*tao2)*h2, the kernels' own rounding
#pragma unroll
            for (int q = 0; q < 4; ++q)
                val[q] = sum_kind == SUM_FLOAT ? finish_float(avq[q], w1[q], p1[q], p0[q]) : finish_double(avq[q], w1[q], p1[q], p0[q]);
            if (seis_row && z == G.s_z) {  // replacement (BKAdd :349-353)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = data_index(G, z, x + q);
                    if (j >= 0) {
                        const float d = seis_row[j];
                        if (d != 0.0f) val[q] = d;
                    }
                }
            }
            if (inject && z == r_u && r_x >= x && r_x < x + 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (x + q == r_x) val[q] = __fadd_rn(val[q], wavelet);
            }
            *reinterpret_cast<float4*>(s2 + lz * CW + 4 * g) = make_float4(val[0], val[1], val[2], val[3]);
        }
    }
    __syncthreads();

    for_cells(oh, ow, [&](int oz, int ox) {
        const int z = o.za + oz, x = o.xa + ox;
        const int lz = z - cza, lx = x - cxa;
        const float vb = sVb[oz * ow + ox];
        const int dz = min(z, NZ - 1 - z), dx = min(x, NX - 1 - x);
        const int a  = min(dz, dx);
        const int sz = (z < NZ - 1 - z) ? 1 : -1, sx = (x < NX - 1 - x) ? 1 : -1;
#define S2(zz, xx) s2[(lz + (zz)) * CW + lx + (xx)]
#define S0(zz, xx) s0[(lz + (zz)) * CW + lx + (xx)]
#define S1(zz, xx) s1[(lz + R + (zz)) * SP + lx + RP + (xx)]
        float Pb;
#ifdef RTM_TIMING_SKIP_ONEWAY  // timing-only builds (wrong results): ring without the one-way solution
        emit(z, x, S2(0, 0), sAx[oz * ow + ox]);
        return;
#endif
        if (abs(dz - dx) <= 1) {
            // corner cells (Hybrid1 :138-155): r1 = sqrt((v*tao/h)^2/2) at the cell itself
            // (GPU_velocity_real.cpp:104-117; tao/h enters as v*tao/h in float)
            const float r  = __fdiv_rn(__fmul_rn(vb, G.tao), G.h);
            const float r2 = __double2float_rn(__dmul_rn(__dmul_rn((double)r, (double)r), 0.5));
            const float r1 = __fsqrt_rn(r2);
            const float rcp = __frcp_rn(__fmaf_rn(2.0f, r1, 1.0f));
            const float nb  = __fadd_rn(S2(0, sx), S2(sz, 0));
            Pb = __fmul_rn(rcp, __fmaf_rn(r1, nb, S1(0, 0)));
        } else {
            int   iz, ix, tz, tx;
            if (dz < dx) {  // top :124 / bottom :132
                iz = sz; ix = 0; tz = 0; tx = 1;
            } else {        // left :128 / right :136
                iz = 0; ix = sx; tz = 1; tx = 0;
            }
            const float vq  = sVq[oz * ow + ox];  // Q2: velocity at the reference's (mis-indexed) cell
            const float tv  = __fmul_rn(G.taoh, vb);
            const float rcp = __frcp_rn(__fadd_rn(tv, 1.0f));
            const float p2i = S2(iz, ix), p0i = S0(iz, ix), p0b = S0(0, 0);
            const float p1b = S1(0, 0), p1i = S1(iz, ix);
            const float A1  = __fadd_rn(__fsub_rn(p2i, p0i), p0b);
            float B = __fadd_rn(__fmul_rn(-2.0f, p1b), p0b);
            B       = __fadd_rn(B, p2i);
            B       = __fsub_rn(B, __fmul_rn(2.0f, p1i));
            B       = __fadd_rn(B, p0i);
            float D = __fsub_rn(S2(iz + tz, ix + tx), __fmul_rn(2.0f, p2i));
            D       = __fadd_rn(D, S2(iz - tz, ix - tx));
            D       = __fadd_rn(D, S0(tz, tx));
            D       = __fsub_rn(D, __fmul_rn(2.0f, p0b));
            D       = __fadd_rn(D, S0(-tz, -tx));
            const float c2 = __fmul_rn(__fmul_rn(G.taoh2, vq), vq);
            Pb = __fmul_rn(rcp, __fmaf_rn(c2, D, __fmaf_rn(tv, A1, -B)));
        }
        // blend (Hybrid2 :160-183): fma(1-w, P2, w*Pb)
        const float w   = G.w[N2 - a];
        const float val = __fmaf_rn(__fsub_rn(1.0f, w), S2(0, 0), __fmul_rn(w, Pb));
#undef S2
#undef S0
#undef S1
        emit(z, x, val, sAx[oz * ow + ox]);
    });
}

__device__ __forceinline__ void store4(float* dst, const float (&o)[4], int x, int xend)
{
    if (x + 3 < xend) {
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (x + q < xend) dst[q] = o[q];
    }
}

// Position of a block in a launch that carries ring CTAs.  The nrc ring CTAs are dealt evenly among
// the interior CTAs -- block b is a ring CTA when b is one of the first nrc multiples of `period` --
// so that an SM works on a ring tile (long, instruction-bound) next to interior tiles (HBM-bound)
// for most of the launch instead of the ring tiles forming its first waves.  period = 1: ring first.
struct BlockRole { bool is_ring; int index; };   // index among the ring CTAs / among the interior CTAs
__host__ __device__ __forceinline__ BlockRole block_role(int b, int nrc, int period, const FastDiv& fd_period)
{
    const int  q = fast_div(b, fd_period);
    const bool ring = q < nrc && b == q * period;           // (nrc == 0: never)
    return BlockRole{ring, ring ? q : b - (q + 1 < nrc ? q + 1 : nrc)};
}
// Ring CTAs of a launch of `total` blocks: every period-th block over the first eighths/8 of the grid
// (measured on C2: 8/8 306 700, 7/8 305 600, 5/8 299 200, ring first 290 000 Mcell-updates/s).
// Needs (ring_ctas - 1) * period < total.
__host__ inline int ring_period_for(bool interleave, int ring_ctas, int total, int eighths = 8)
{
    if (!interleave || ring_ctas == 0) return 1;
    eighths = eighths < 1 ? 1 : (eighths > 8 ? 8 : eighths);
    const int p = (int)((long long)total * eighths / 8 / ring_ctas);
    return p < 1 ? 1 : p;
}

// ------------------------------------------------------------------------------------
// Forward step: slot k from slots k-1 (P1, via TMA / raw pointer) and k-2 (P0).
// grid = (ring tiles + interior tiles, shots)
// ------------------------------------------------------------------------------------
struct FwdArgs {
    const float* P1;   // slot k-1, all shots
    const float* P0;   // slot k-2
    float*       P2;   // slot k
    const int2*  src;  // per shot (r_u, r_x)
    float        wavelet;
    int          k;    // time slot being produced
    int          nshots;
    int          tma_s0;  // offset of shot 0 along the tensor map's 3rd dimension (store-all: slot*S)
    const int*   tiles;   // interior tiles of this launch (null: all tiles 0..ntiles-1)
    int          ntiles;  // interior tiles per shot in this launch
    FastDiv      fd_ntiles;
    int          lookahead;  // CTAs ahead whose TMA box is prefetched into L2 (0: off; all-tiles launches only)
    int          lookahead_p0, tma_s0_p0;  // ... and its slot k-2 box (tma_s0_p0: like tma_s0, for the P0 map)
    int          do_ring; // this launch also carries the ring tiles
    int          ring_period;  // every ring_period-th block (from block 0) is a ring CTA; 1 = all ring CTAs first
    FastDiv      fd_period;
    Strips       st;   // may hold nulls when strips are not wanted (pure modelling)
    float*       gather;  // [S][NT][n] time-major, or null
};

#ifndef RTM_FWD_MINB
#define RTM_FWD_MINB 4
#endif
#ifndef RTM_BWD_MINB
#define RTM_BWD_MINB 3
#endif
#ifndef RTM_FWD_MINB_LS_BIG
#define RTM_FWD_MINB_LS_BIG 3
#endif
#ifndef RTM_FWD_MINB_R12
#define RTM_FWD_MINB_R12 3
#endif
#ifndef RTM_BWD_MINB_R12
#define RTM_BWD_MINB_R12 2
#endif
#ifndef RTM_FWD_MINB_LS
#define RTM_FWD_MINB_LS 3
#endif
#ifndef RTM_BWD_MINB_LS
#define RTM_BWD_MINB_LS 2
#endif
template <int RP, bool LS, int NR>
__global__ void __launch_bounds__(kThreads, (RP <= 4 ? (LS ? RTM_FWD_MINB_LS : RTM_FWD_MINB) : (LS ? RTM_FWD_MINB_LS_BIG : (RP <= 8 ? 3 : RTM_FWD_MINB_R12))))
fwd_step_kernel(const __grid_constant__ CUtensorMap tmP1, const __grid_constant__ CUtensorMap tmP0,
                const __grid_constant__ Geo G, const FwdArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // 1-D grid: the interior tiles shot by shot, with the ring tiles of all shots dealt evenly among
    // them over the grid (block_role(); they are the longest-running CTAs).
    // A time step may be split into several launches, one per operator-length class of the
    // interior tiles (adaptive operator): each carries its own tile list; one of them the ring.
    const int  nring = a.do_ring ? 2 * G.nband + 2 * G.nside : 0, nint = a.ntiles;
    const BlockRole role = block_role(blockIdx.x, a.nshots * nring, a.ring_period, a.fd_period);
    const bool is_ring = role.is_ring;
    const int  bi   = role.index;
    const int  shot = is_ring ? fast_div(bi, G.fd_nring) : fast_div(bi, a.fd_ntiles);
    const int  bt   = bi - shot * (is_ring ? nring : nint);  // ring tile / position in the tile list
    const long long so = (long long)shot * G.shot_stride + G.padL;  // (z=0,x=0) of this shot
    const int2 src = a.src[shot];
    const int  sum_kind = G.iLSTE == 0 ? SUM_FLOAT : SUM_DOUBLE;  // Add vs Add_Con

    if (is_ring) {
#ifdef RTM_TIMING_SKIP_RING   // timing-only builds (wrong results): what the ring tiles cost
        return;
#endif
        float* P2 = a.P2 + so;
        const int N2 = G.N2, nf = G.nfdmax, NZ = G.NZ, NX = G.NX, k = a.k;
        const Strips st = a.st;
        float* gather = a.gather;
        ring_tile<RP, LS>(G, bt, a.P1 + so, a.P0 + so, sum_kind, true, src.x, src.y, a.wavelet,
                      nullptr, reinterpret_cast<float*>(smem_raw),
                      [](int, int) -> const float* { return nullptr; },
                      [&](int z, int x, float val, float) {
            P2[(size_t)z * G.pitch + x] = val;
            if (st.up) {  // Hybrid3 :184-208 (slot k)
                const size_t sx = ((size_t)shot * G.NT + k) * nf * G.mod_NX;
                const size_t sz = ((size_t)shot * G.NT + k) * nf * G.mod_NZ;
                if (x >= N2 && x < NX - N2) {
                    if (z >= N2 - nf && z < N2) st.up[sx + (size_t)(z - (N2 - nf)) * G.mod_NX + x - N2] = val;
                    else if (z >= NZ - N2 && z < NZ - N2 + nf) st.dw[sx + (size_t)(z - (NZ - N2)) * G.mod_NX + x - N2] = val;
                } else if (z >= N2 && z < NZ - N2) {
                    if (x >= N2 - nf && x < N2) st.lf[sz + (size_t)(z - N2) * nf + x - (N2 - nf)] = val;
                    else if (x >= NX - N2 && x < NX - N2 + nf) st.rt[sz + (size_t)(z - N2) * nf + x - (NX - N2)] = val;
                }
            }
            if (gather) {
                const int j = data_index(G, z, x);
                if (j >= 0) gather[((size_t)shot * G.NT + k) * G.n + j] = val;
            }
        });
        return;
    }

    // ---- interior tile
    const int t   = a.tiles ? a.tiles[bt] : bt;
    const int tz  = fast_div(t, G.fd_ntx), tx = t - tz * G.ntx;
    const int z0  = G.N2 + tz * (kWarps * NR), x0 = G.N2 + tx * kTX;  // first interior cell of the tile
    float*    sP  = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + Tile<RP, NR>::BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, Tile<RP, NR>::BYTES);
        tma_load_3d(sP, &tmP1, bar, G.padL + x0 - RP, z0 - RP, a.tma_s0 + shot);
    }
    // L2 look-ahead: the boxes of the CTA `lookahead` blocks further on (its TMA copy then hits L2)
    int la_t = -1, la_s = 0;
    if (tid == 0 && a.lookahead > 0 && bi + a.lookahead < a.nshots * nint) {
        la_s = fast_div(bi + a.lookahead, a.fd_ntiles);
        const int i2 = bi + a.lookahead - la_s * nint;
        la_t = a.tiles ? __ldg(a.tiles + i2) : i2;
    }
    const int zend = G.NZ - G.N2, xend = G.NX - G.N2;
    LsTable T{};
    if (LS) {  // this tile's slice of the operator table -> shared memory (behind the halo tile)
        T = ls_stage_slice<RP>(G, G.tile_bins_f[t], reinterpret_cast<float*>(smem_raw + Tile<RP, NR>::BYTES + 16));
        __syncthreads();
    }
    const int lz0 = warp * NR, lx0 = lane * 4;
    const int z = z0 + lz0, x = x0 + lx0;
    if (x >= xend || z >= zend) return;  // (whole 4-cell groups; rows are warp-uniform)
    const int nrow = min(NR, zend - z);

    // coalesced float4 loads of the other streams while the TMA copy is in flight.
    // The velocity only enters through a = ((v*v)*tao2)*h2 (read precomputed) and, for the
    // adaptive operator, through the cell's bin (2-byte side array).
    const size_t cell0 = (size_t)z * G.pitch + x;  // row pointers advance by `pitch` per row
    const float* p0p = a.P0 + so + cell0;
    float*       p2p = a.P2 + so + cell0;
    const float* vp  = G.avel + G.padL + cell0;
    const unsigned short* bp = G.bins + G.padL + cell0;
    const float* sp  = sP + (lz0 + RP) * Tile<RP, NR>::SP + lx0 + RP;
    const bool full  = x + 3 < xend;
    const int  src_r = (src.y >= x && src.y < x + 4) ? src.x - z : -1;  // row of this thread holding the source
    const int  gat_r = a.gather ? G.s_z - z : -1;
    float4 p0n = *reinterpret_cast<const float4*>(p0p);
    float4 vn  = __ldg(reinterpret_cast<const float4*>(vp));
    uint2  bn  = make_uint2(0u, 0u);
    if (LS) bn = __ldg(reinterpret_cast<const uint2*>(bp));
    mbar_wait(bar, 0);
    if (la_t >= 0) {
        const int tz2 = fast_div(la_t, G.fd_ntx), tx2 = la_t - tz2 * G.ntx;
        tma_prefetch_3d(&tmP1, G.padL + G.N2 + tx2 * kTX - RP, G.N2 + tz2 * (kWarps * NR) - RP, a.tma_s0 + la_s);
        if (a.lookahead_p0)
            tma_prefetch_3d(&tmP0, G.padL + G.N2 + tx2 * kTX - RP, G.N2 + tz2 * (kWarps * NR) - RP, a.tma_s0_p0 + la_s);
    }

#pragma unroll
    for (int r = 0; r < NR; ++r) {
        if (r >= nrow) break;
        float vq[4], pq[4];
        unpack(vn, vq);
        unpack(p0n, pq);
        const uint2 bc = bn;
        if (r + 1 < nrow) {  // next row's loads fly during this row's arithmetic
            p0n = *reinterpret_cast<const float4*>(p0p + G.pitch);
            vn  = __ldg(reinterpret_cast<const float4*>(vp + G.pitch));
            if (LS) bn = __ldg(reinterpret_cast<const uint2*>(bp + G.pitch));
        }
        float w1[4], p1[4], o[4];
        stencil_row<RP, LS>(G, sp, G.nfdmax, T, bc, w1, p1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            o[q] = LS ? finish_float(vq[q], w1[q], p1[q], pq[q])     // Add
                      : finish_double(vq[q], w1[q], p1[q], pq[q]);   // Add_Con
        }
        if (r == src_r) {  // :74-77
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (x + q == src.y) o[q] = __fadd_rn(o[q], a.wavelet);
        }
        if (full) {
            *reinterpret_cast<float4*>(p2p) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (x + q < xend) p2p[q] = o[q];
        }
        if (r == gat_r) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = (x + q < xend) ? data_index(G, z + r, x + q) : -1;
                if (j >= 0) a.gather[((size_t)shot * G.NT + a.k) * G.n + j] = o[q];
            }
        }
        p0p += G.pitch; p2p += G.pitch; vp += G.pitch; bp += G.pitch; sp += Tile<RP, NR>::SP;
    }
}

// ------------------------------------------------------------------------------------
// Backward step k: source slot k reconstructed in the interior (in place over slot k+2),
// receiver slot k on the full grid with data replacement and ABC, imaging accumulators.
// ------------------------------------------------------------------------------------
struct BwdArgs {
    const float* Sk;   // store-all mode: the stored forward field of slot k (S1/S02/strips unused)
    const float* S1;   // source slot k+1 (ring = strips of slot k+1)
    const float* S02;  // source slot k+2
    float*       S2;   // out: source slot k (interior), ring := strips of slot k.  May alias S02 (in place)
    const float* R1;   // receiver, current
    const float* R0;   // receiver, previous
    float*       R2;   // receiver, new
    const int2*  src;
    float        wavelet;
    int          k;
    int          nshots;
    const int*   tiles;   // interior tiles of this launch (null: all tiles 0..ntiles-1)
    int          ntiles;
    FastDiv      fd_ntiles;
    int          lookahead;     // CTAs ahead whose TMA boxes are prefetched into L2 (0: off)
    int          lookahead_p0;  // bit 0: also its receiver slot k+2 box, bit 1: also its source slot k+2 box
    int          do_ring;
    int          ring_period;
    FastDiv      fd_period;
    Strips       st;
    const float* seis;  // [S][NT][n] time-major; row k+1 is imposed
    float *sumS, *sumR, *rel1, *rel2;  // accumulators, field layout
};

// STORE: the source field of slot k is read from the stored forward wavefield instead of being
// reconstructed (RTM_FLAG_STORE_ALL; not a reference mode).
template <int RP, bool LS, int NR, bool STORE>
__global__ void __launch_bounds__(kThreads, (RP <= 4 ? (LS ? RTM_BWD_MINB_LS : RTM_BWD_MINB) : ((RP <= 8 && !LS) ? 3 : RTM_BWD_MINB_R12)))
bwd_step_kernel(const __grid_constant__ CUtensorMap tmS1, const __grid_constant__ CUtensorMap tmR1,
                const __grid_constant__ CUtensorMap tmS0, const __grid_constant__ CUtensorMap tmR0,
                const __grid_constant__ Geo G, const BwdArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int  nring = a.do_ring ? 2 * G.nband + 2 * G.nside : 0, nint = a.ntiles;
    const BlockRole role = block_role(blockIdx.x, a.nshots * nring, a.ring_period, a.fd_period);
    const bool is_ring = role.is_ring;
    const int  bi   = role.index;
    const int  shot = is_ring ? fast_div(bi, G.fd_nring) : fast_div(bi, a.fd_ntiles);
    const int  bt   = bi - shot * (is_ring ? nring : nint);
    const long long so = (long long)shot * G.shot_stride + G.padL;
    const int2 src = a.src[shot];
    const float* seis_row = a.seis + ((size_t)shot * G.NT + (a.k + 1)) * G.n;

    if (is_ring) {
#ifdef RTM_TIMING_SKIP_RING
        return;
#endif
        float* R2 = a.R2 + so;
        float* SX = a.S2 + so;
        const int N2 = G.N2, nf = G.nfdmax, NZ = G.NZ, NX = G.NX, k = a.k;
        const Strips st = a.st;
        // BKEqual :222-245, one step early: the ring of the buffer that becomes the "current"
        // source field at step k-1 receives the strips of slot k.  The strip value of a cell is
        // fetched with the tile's halo (`stage`) and written back by `emit`.
        auto strip_src = [&](int z, int x) -> const float* {
            if (STORE) return nullptr;
            const size_t sx = ((size_t)shot * G.NT + k) * nf * G.mod_NX;
            const size_t sz = ((size_t)shot * G.NT + k) * nf * G.mod_NZ;
            if (x >= N2 && x < NX - N2) {
                if (z >= N2 - nf && z < N2) return st.up + sx + (size_t)(z - (N2 - nf)) * G.mod_NX + x - N2;
                if (z >= NZ - N2 && z < NZ - N2 + nf) return st.dw + sx + (size_t)(z - (NZ - N2)) * G.mod_NX + x - N2;
            } else if (z >= N2 && z < NZ - N2) {
                if (x >= N2 - nf && x < N2) return st.lf + sz + (size_t)(z - N2) * nf + x - (N2 - nf);
                if (x >= NX - N2 && x < NX - N2 + nf) return st.rt + sz + (size_t)(z - N2) * nf + x - (NX - N2);
            }
            return nullptr;
        };
        auto in_strips = [&](int z, int x) -> bool {
            if (STORE) return false;
            if (x >= N2 && x < NX - N2) return (z >= N2 - nf && z < N2) || (z >= NZ - N2 && z < NZ - N2 + nf);
            if (z >= N2 && z < NZ - N2) return (x >= N2 - nf && x < N2) || (x >= NX - N2 && x < NX - N2 + nf);
            return false;
        };
        ring_tile<RP, LS>(G, bt, a.R1 + so, a.R0 + so, SUM_FLOAT, false, 0, 0, 0.0f, seis_row,
                      reinterpret_cast<float*>(smem_raw), strip_src,
                      [&](int z, int x, float val, float strip) {
            R2[(size_t)z * G.pitch + x] = val;
            if (in_strips(z, x)) SX[(size_t)z * G.pitch + x] = strip;
        });
        return;
    }

    const int t   = a.tiles ? a.tiles[bt] : bt;
    const int tz  = fast_div(t, G.fd_ntx), tx = t - tz * G.ntx;
    const int z0  = G.N2 + tz * (kWarps * NR), x0 = G.N2 + tx * kTX;
    constexpr int NTILE = STORE ? 1 : 2;  // halo tiles in shared memory
    float*    sS  = reinterpret_cast<float*>(smem_raw);
    float*    sR  = reinterpret_cast<float*>(smem_raw + (NTILE - 1) * Tile<RP, NR>::BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + NTILE * Tile<RP, NR>::BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, NTILE * Tile<RP, NR>::BYTES);
        if (!STORE) tma_load_3d(sS, &tmS1, bar, G.padL + x0 - RP, z0 - RP, shot);
        tma_load_3d(sR, &tmR1, bar, G.padL + x0 - RP, z0 - RP, shot);
    }
    int la_t = -1, la_s = 0;  // L2 look-ahead, as in the forward kernel
    if (tid == 0 && a.lookahead > 0 && bi + a.lookahead < a.nshots * nint) {
        la_s = fast_div(bi + a.lookahead, a.fd_ntiles);
        const int i2 = bi + a.lookahead - la_s * nint;
        la_t = a.tiles ? __ldg(a.tiles + i2) : i2;
    }
    const int zend = G.NZ - G.N2, xend = G.NX - G.N2;
    LsTable T{};
    if (LS) {
        T = ls_stage_slice<RP>(G, G.tile_bins_b[t], reinterpret_cast<float*>(smem_raw + NTILE * Tile<RP, NR>::BYTES + 16));
        __syncthreads();
    }
    const int lz0 = warp * NR, lx0 = lane * 4;
    const int z = z0 + lz0, x = x0 + lx0;
    if (x >= xend || z >= zend) return;
    const int nrow = min(NR, zend - z);
    const bool compen = G.iCompen == 1;

    size_t o = so + (size_t)z * G.pitch + x;  // advances by `pitch` per row
    const float* vp = G.avel + G.padL + (size_t)z * G.pitch + x;
    const unsigned short* bp = G.bins + G.padL + (size_t)z * G.pitch + x;
    const float* spS = sS + (lz0 + RP) * Tile<RP, NR>::SP + lx0 + RP;
    const float* spR = sR + (lz0 + RP) * Tile<RP, NR>::SP + lx0 + RP;
    const bool full  = x + 3 < xend;
    const int  src_r = (src.y >= x && src.y < x + 4) ? src.x - z : -1;
    const int  dat_r = G.s_z - z;  // row of this thread on the data line (if in 0..NR-1)
    float4 vn  = __ldg(reinterpret_cast<const float4*>(vp));
    uint2  bn  = make_uint2(0u, 0u);
    if (LS) bn = __ldg(reinterpret_cast<const uint2*>(bp));
    float4 s0n = *reinterpret_cast<const float4*>((STORE ? a.Sk : a.S02) + o);  // STORE: slot k itself
    float4 r0n = *reinterpret_cast<const float4*>(a.R0 + o);
    mbar_wait(bar, 0);
    if (la_t >= 0) {
        const int tz2 = fast_div(la_t, G.fd_ntx), tx2 = la_t - tz2 * G.ntx;
        const int xx = G.padL + G.N2 + tx2 * kTX - RP, zz = G.N2 + tz2 * (kWarps * NR) - RP;
        if (!STORE) tma_prefetch_3d(&tmS1, xx, zz, la_s);
        tma_prefetch_3d(&tmR1, xx, zz, la_s);
        if (a.lookahead_p0 & 1) tma_prefetch_3d(&tmR0, xx, zz, la_s);
        if (!STORE && (a.lookahead_p0 & 2)) tma_prefetch_3d(&tmS0, xx, zz, la_s);
    }

    auto put = [&](float* base, const float (&val)[4]) {
        if (full) {
            *reinterpret_cast<float4*>(base + o) = make_float4(val[0], val[1], val[2], val[3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (x + q < xend) base[o + q] = val[q];
        }
    };

#pragma unroll
    for (int r = 0; r < NR; ++r) {
        if (r >= nrow) break;
        float av[4], s0[4], r0[4];
        unpack(vn, av);
        unpack(s0n, s0);
        unpack(r0n, r0);
        // accumulators of this row and the next row's streams fly during the arithmetic
        const float4 a1 = *reinterpret_cast<const float4*>(a.rel1 + o);
        const float4 a2 = *reinterpret_cast<const float4*>(a.rel2 + o);
        float4 aS = make_float4(0.f, 0.f, 0.f, 0.f), aR = aS;
        if (compen) {
            aS = *reinterpret_cast<const float4*>(a.sumS + o);
            aR = *reinterpret_cast<const float4*>(a.sumR + o);
        }
        const uint2 bc = bn;
        if (r + 1 < nrow) {
            vn  = __ldg(reinterpret_cast<const float4*>(vp + G.pitch));
            if (LS) bn = __ldg(reinterpret_cast<const uint2*>(bp + G.pitch));
            s0n = *reinterpret_cast<const float4*>((STORE ? a.Sk : a.S02) + o + G.pitch);
            r0n = *reinterpret_cast<const float4*>(a.R0 + o + G.pitch);
        }
        float w1[4], p1[4], S2[4], R2[4];
        // source field: BKAdd_EFF / BKAdd_EFF_Con, double final sum, + wavelet at the source
        if (STORE) {
#pragma unroll
            for (int q = 0; q < 4; ++q) S2[q] = s0[q];
        } else {
            stencil_row<RP, LS>(G, spS, G.nfdmax, T, bc, w1, p1);
#pragma unroll
            for (int q = 0; q < 4; ++q) S2[q] = finish_double(av[q], w1[q], p1[q], s0[q]);
            if (r == src_r) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (x + q == src.y) S2[q] = __fadd_rn(S2[q], a.wavelet);
            }
        }
        // receiver field: BKAdd / BKAdd_Con, float final sum, data replacement
        stencil_row<RP, LS>(G, spR, G.nfdmax, T, bc, w1, p1);
#pragma unroll
        for (int q = 0; q < 4; ++q) R2[q] = finish_float(av[q], w1[q], p1[q], r0[q]);
        if (r == dat_r) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = (x + q < xend) ? data_index(G, z + r, x + q) : -1;
                if (j >= 0) {
                    const float d = seis_row[j];
                    if (d != 0.0f) R2[q] = d;
                }
            }
        }
        // imaging (Rel_Compen :503-517 / Rel_NonCompen :489-501), S = source, R = receiver
        float r1v[4], r2v[4], sSv[4], sRv[4];
        unpack(a1, r1v);
        unpack(a2, r2v);
        unpack(aS, sSv);
        unpack(aR, sRv);
        if (compen) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                sSv[q] = __fadd_rn(S2[q], sSv[q]);
                sRv[q] = __fadd_rn(R2[q], sRv[q]);
                r1v[q] = __fmaf_rn(sRv[q], sSv[q], r1v[q]);
                r2v[q] = __fmaf_rn(S2[q], S2[q], r2v[q]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                r1v[q] = __fmaf_rn(R2[q], S2[q], r1v[q]);
                r2v[q] = __fmaf_rn(S2[q], S2[q], r2v[q]);
            }
        }
        if (!STORE) put(a.S2, S2);
        put(a.R2, R2);
        put(a.rel1, r1v);
        put(a.rel2, r2v);
        if (compen) {
            put(a.sumS, sSv);
            put(a.sumR, sRv);
        }
        o += G.pitch; vp += G.pitch; bp += G.pitch; spS += Tile<RP, NR>::SP; spR += Tile<RP, NR>::SP;
    }
}


// ------------------------------------------------------------------------------------
// Two backward steps (k and k-1) of one INNER interior tile in a single pass.
//
// Slots k and k-1 depend on slots k+1/k+2 only through a neighbourhood of 2 operator radii, so a
// tile whose neighbourhood holds no absorbing-ring cell can advance two steps while its fields
// and imaging accumulators cross HBM once: per grid cell and PAIR of steps 68 B (S,R of slots
// k+1,k+2 in, S,R of slots k,k-1 out, v, four accumulators read-modify-write) instead of 2 x 60 B.
//   phase A: slot k of both fields on the tile grown by RP cells -> shared memory ("mid" tiles);
//            the current fields (slot k+1) arrive as TMA boxes with a 2*RP halo;
//   phase B: slot k-1 on the tile itself from the mid tiles (P0 = centre of the TMA boxes), both
//            imaging updates in time order, all stores.
// Every cell value is produced by exactly the operations of the single-step kernel, so results
// are bit-identical to stepping twice.  Ring tiles and the frame of interior tiles next to the
// ring run the single-step kernel for slot k (concurrently) and for slot k-1 (afterwards).
// Buffers: slots k+2,k+1 are only read, slots k,k-1 go to two other buffers (4 per field).
// ------------------------------------------------------------------------------------
struct Acc4Maps { CUtensorMap m[4]; };  // rel1, rel2, sumS, sumR with a box of one two-step tile

struct Bwd2Args {
    const float* S0;   // source slot k+2 (slot k+1 comes through the tensor map)
    float*       Sk;   // out: source slot k
    float*       Skm;  // out: source slot k-1
    const float* R0;   // receiver slot k+2
    float*       Rk;
    float*       Rkm;
    const int2*  src;
    float        wavelet_k, wavelet_km;  // source term of steps k and k-1
    int          k;
    int          nshots;
    const int*   tiles;   // inner tiles of this launch
    int          ntiles;
    int          rect_t0, rect_nx;  // rect_nx > 0: the tiles form a rectangle rect_nx wide starting at tile rect_t0,
                                    // row stride rect_dz tiles (no list lookup before the TMA copies are issued)
    int          rect_dz;
    FastDiv      fd_ntiles, fd_rect;
    int          lookahead;  // CTAs ahead whose TMA boxes are prefetched into L2 (0: off)
    int          lookahead_more;  // bit 0: also its slot k+2 boxes, bit 1: also its accumulator tiles
    const float* seis;    // [S][NT][n]; row k+1 is imposed in step k, row k in step k-1
    float *sumS, *sumR, *rel1, *rel2;
};

// shape of the two-step kernel, tuned on the B200 (profiles/README.md)
#ifndef RTM_NR_B2
#define RTM_NR_B2 2      // rows per thread in phase B: tile height 8 * RTM_NR_B2, a multiple of the single-step tile's
#endif
template <int RP> struct Tile2 {
    static constexpr int NR   = RTM_NR_B2;                 // rows per thread in phase B
    static constexpr int TZ   = kWarps * NR;               // one or two tiles of the single-step tiling
    static_assert(TZ % (kWarps * RTM_NR_B) == 0, "two-step tiles are stacks of single-step tiles");
    static constexpr int NRA  = (TZ + 2 * RP) / kWarps;    // grown rows per thread in phase A
    static constexpr int SPA  = kTX + 4 * RP, ROWSA = TZ + 4 * RP;  // slot k+1 boxes (halo 2*RP)
    static constexpr int SPB  = kTX + 2 * RP, ROWSB = TZ + 2 * RP;  // slot k mid tiles (halo RP)
    static constexpr int CUR_BYTES = SPA * ROWSA * 4, MID_BYTES = SPB * ROWSB * 4;
    static constexpr int BYTES = 2 * CUR_BYTES + 2 * MID_BYTES;
};

#ifndef RTM_BWD2_MINB
#define RTM_BWD2_MINB (RTM_NR_B2 <= 2 ? 3 : 2)
#endif
template <int RP, bool LS>
__global__ void __launch_bounds__(kThreads, (RP <= 4 ? RTM_BWD2_MINB : (RP <= 8 && RTM_NR_B2 <= 2 ? 2 : 1)))
bwd2_step_kernel(const __grid_constant__ CUtensorMap tmS1, const __grid_constant__ CUtensorMap tmR1,
                 const __grid_constant__ CUtensorMap tmS0, const __grid_constant__ CUtensorMap tmR0,
                 const __grid_constant__ Acc4Maps tmAcc, const __grid_constant__ Geo G, const Bwd2Args a)
{
    using T2 = Tile2<RP>;
    constexpr int NR = T2::NR, SPA = T2::SPA, SPB = T2::SPB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int  shot = fast_div(blockIdx.x, a.fd_ntiles);
    const int  ti   = blockIdx.x - shot * a.ntiles;
    const int  tr   = fast_div(ti, a.fd_rect);
    const int  t    = a.rect_nx > 0 ? a.rect_t0 + tr * a.rect_dz + (ti - tr * a.rect_nx) : a.tiles[ti];
    const long long so = (long long)shot * G.shot_stride + G.padL;
    const int2 src = a.src[shot];
    const int  tz  = fast_div(t, G.fd_ntx), tx = t - tz * G.ntx;
    const int  z0  = G.N2 + tz * (kWarps * RTM_NR_B), x0 = G.N2 + tx * kTX;  // t counts single-step tiles
    float*    sS1 = reinterpret_cast<float*>(smem_raw);
    float*    sR1 = reinterpret_cast<float*>(smem_raw + T2::CUR_BYTES);
    float*    mS  = reinterpret_cast<float*>(smem_raw + 2 * T2::CUR_BYTES);
    float*    mR  = reinterpret_cast<float*>(smem_raw + 2 * T2::CUR_BYTES + T2::MID_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + T2::BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, 2 * T2::CUR_BYTES);
        tma_load_3d(sS1, &tmS1, bar, G.padL + x0 - 2 * RP, z0 - 2 * RP, shot);
        tma_load_3d(sR1, &tmR1, bar, G.padL + x0 - 2 * RP, z0 - 2 * RP, shot);
    }
    // L2 look-ahead: the boxes of a CTA that starts a little later (its TMA copies then hit L2);
    // a listed tile index is fetched now and used after the wait for this CTA's own boxes
    int la_t = -1, la_s = 0;
    if (tid == 0 && a.lookahead > 0 && (int)blockIdx.x + a.lookahead < (int)gridDim.x) {
        const int b2 = blockIdx.x + a.lookahead;
        la_s = fast_div(b2, a.fd_ntiles);
        const int i2 = b2 - la_s * a.ntiles, r2 = fast_div(i2, a.fd_rect);
        la_t = a.rect_nx > 0 ? a.rect_t0 + r2 * a.rect_dz + (i2 - r2 * a.rect_nx) : __ldg(a.tiles + i2);
    }
    auto look_ahead = [&]() {
        if (la_t < 0) return;
        const int tz2 = fast_div(la_t, G.fd_ntx), tx2 = la_t - tz2 * G.ntx;
        const int zz = G.N2 + tz2 * (kWarps * RTM_NR_B) - 2 * RP, xx = G.padL + G.N2 + tx2 * kTX - 2 * RP;
        tma_prefetch_3d(&tmS1, xx, zz, la_s);
        tma_prefetch_3d(&tmR1, xx, zz, la_s);
        if (a.lookahead_more & 1) {  // slot k+2 (P0 of phase A)
            tma_prefetch_3d(&tmS0, xx, zz, la_s);
            tma_prefetch_3d(&tmR0, xx, zz, la_s);
        }
        if (a.lookahead_more & 2) {  // accumulators (box = the tile itself)
            const int na = G.iCompen == 1 ? 4 : 2;
            for (int i = 0; i < na; ++i) tma_prefetch_3d(&tmAcc.m[i], xx + 2 * RP, zz + 2 * RP, la_s);
        }
    };
    LsTable T{};
    if (LS) {
        T = ls_stage_slice<RP>(G, G.tile_bins_b2[t], reinterpret_cast<float*>(smem_raw + T2::BYTES + 16));
        __syncthreads();
    }
    const float* AV = G.avel + G.padL;
    const unsigned short* BN = G.bins + G.padL;
    const float* S0 = a.S0 + so;
    const float* R0 = a.R0 + so;

    // ---- phase A: slot k on rows [z0-RP, z0+TZ+RP), columns [x0-RP, x0+kTX+RP)
    {
        constexpr int NRA = T2::ROWSB / kWarps;  // grown rows per warp
        constexpr int GH  = RP / 4;              // halo float4 groups per side
        const float* seisA = a.seis + ((size_t)shot * G.NT + (a.k + 1)) * G.n;
        // one float4 group (4 x-adjacent cells) of grown row rg, both fields
        auto item = [&](int rg, int g, const float4& s0v, const float4& r0v, const float4& avv, const uint2& b2) {
            const int z = z0 - RP + rg, x = x0 - RP + 4 * g;
            float av[4], p0[4], w1[4], p1[4], o[4];
            unpack(avv, av);
            stencil_row<RP, LS, SPA>(G, sS1 + (rg + RP) * SPA + 4 * g + RP, G.nfdmax, T, b2, w1, p1);
            unpack(s0v, p0);
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = finish_double(av[q], w1[q], p1[q], p0[q]);
            if (z == src.x && src.y >= x && src.y < x + 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (x + q == src.y) o[q] = __fadd_rn(o[q], a.wavelet_k);
            }
            *reinterpret_cast<float4*>(mS + rg * SPB + 4 * g) = make_float4(o[0], o[1], o[2], o[3]);
            stencil_row<RP, LS, SPA>(G, sR1 + (rg + RP) * SPA + 4 * g + RP, G.nfdmax, T, b2, w1, p1);
            unpack(r0v, p0);
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = finish_float(av[q], w1[q], p1[q], p0[q]);
            if (z == G.s_z) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = data_index(G, z, x + q);
                    if (j >= 0) {
                        const float d = seisA[j];
                        if (d != 0.0f) o[q] = d;
                    }
                }
            }
            *reinterpret_cast<float4*>(mR + rg * SPB + 4 * g) = make_float4(o[0], o[1], o[2], o[3]);
        };
        // main part: lane -> group GH+lane (the tile's own columns), warp -> NRA consecutive rows
        const int rg0 = warp * NRA;
        const size_t cell0 = (size_t)(z0 - RP + rg0) * G.pitch + x0 + 4 * lane;
        {
            size_t cell = cell0;
            float4 s0n = __ldg(reinterpret_cast<const float4*>(S0 + cell));
            float4 r0n = __ldg(reinterpret_cast<const float4*>(R0 + cell));
            float4 avn = __ldg(reinterpret_cast<const float4*>(AV + cell));
            uint2  bn  = make_uint2(0u, 0u);
            if (LS) bn = __ldg(reinterpret_cast<const uint2*>(BN + cell));
            mbar_wait(bar, 0);
            look_ahead();
#pragma unroll
            for (int r = 0; r < NRA; ++r) {
                const float4 s0c = s0n, r0c = r0n, avc = avn;
                const uint2  bc  = bn;
                if (r + 1 < NRA) {
                    cell += G.pitch;
                    s0n = __ldg(reinterpret_cast<const float4*>(S0 + cell));
                    r0n = __ldg(reinterpret_cast<const float4*>(R0 + cell));
                    avn = __ldg(reinterpret_cast<const float4*>(AV + cell));
                    if (LS) bn = __ldg(reinterpret_cast<const uint2*>(BN + cell));
                }
                item(rg0 + r, GH + lane, s0c, r0c, avc, bc);
            }
        }
        // the 2*GH halo groups of every grown row, flattened over the CTA
        for (int i = tid; i < T2::ROWSB * 2 * GH; i += kThreads) {
            const int rg = i / (2 * GH), j = i % (2 * GH);
            const int g  = j < GH ? j : SPB / 4 - 2 * GH + j;
            const size_t ce = (size_t)(z0 - RP + rg) * G.pitch + x0 - RP + 4 * g;
            const float4 s0c = __ldg(reinterpret_cast<const float4*>(S0 + ce));
            const float4 r0c = __ldg(reinterpret_cast<const float4*>(R0 + ce));
            const float4 avc = __ldg(reinterpret_cast<const float4*>(AV + ce));
            uint2 bc = make_uint2(0u, 0u);
            if (LS) bc = __ldg(reinterpret_cast<const uint2*>(BN + ce));
            item(rg, g, s0c, r0c, avc, bc);
        }
    }
    // first-row loads of phase B fly across the barrier (phase A's registers are dead here)
    const size_t cellB = (size_t)(z0 + warp * NR) * G.pitch + x0 + lane * 4;
    const bool   compenB = G.iCompen == 1;
    float4 avB = __ldg(reinterpret_cast<const float4*>(AV + cellB));
    float4 a1B = *reinterpret_cast<const float4*>(a.rel1 + so + cellB);
    float4 a2B = *reinterpret_cast<const float4*>(a.rel2 + so + cellB);
    float4 aSB = make_float4(0.f, 0.f, 0.f, 0.f), aRB = aSB;
    if (compenB) {
        aSB = *reinterpret_cast<const float4*>(a.sumS + so + cellB);
        aRB = *reinterpret_cast<const float4*>(a.sumR + so + cellB);
    }
    __syncthreads();

    // ---- phase B: slot k-1 on the tile (always a full tile), imaging updates of both steps
    {
        const int lz0 = warp * NR, lx0 = lane * 4;
        const int z = z0 + lz0, x = x0 + lx0;
        const bool compen = G.iCompen == 1;
        const float* seisB = a.seis + ((size_t)shot * G.NT + a.k) * G.n;
        size_t cell = (size_t)z * G.pitch + x;
        size_t o    = so + cell;
        float4 avn = avB;
        uint2  bn  = make_uint2(0u, 0u);
        if (LS) bn = __ldg(reinterpret_cast<const uint2*>(BN + cell));
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            float4 a1 = a1B, a2 = a2B, aS = aSB, aR = aRB;
            if (r > 0) {
                a1 = *reinterpret_cast<const float4*>(a.rel1 + o);
                a2 = *reinterpret_cast<const float4*>(a.rel2 + o);
                if (compen) {
                    aS = *reinterpret_cast<const float4*>(a.sumS + o);
                    aR = *reinterpret_cast<const float4*>(a.sumR + o);
                }
            }
            float av[4];
            unpack(avn, av);
            const uint2 bc = bn;
            if (r + 1 < NR) {
                avn = __ldg(reinterpret_cast<const float4*>(AV + cell + G.pitch));
                if (LS) bn = __ldg(reinterpret_cast<const uint2*>(BN + cell + G.pitch));
            }
            float w1[4], sk[4], rk[4], p0[4], skm[4], rkm[4];
            // source field, slot k-1
            stencil_row<RP, LS, SPB>(G, mS + (lz0 + r + RP) * SPB + lx0 + RP, G.nfdmax, T, bc, w1, sk);
            unpack(*reinterpret_cast<const float4*>(sS1 + (lz0 + r + 2 * RP) * SPA + lx0 + 2 * RP), p0);
#pragma unroll
            for (int q = 0; q < 4; ++q) skm[q] = finish_double(av[q], w1[q], sk[q], p0[q]);
            if (z + r == src.x && src.y >= x && src.y < x + 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (x + q == src.y) skm[q] = __fadd_rn(skm[q], a.wavelet_km);
            }
            // receiver field, slot k-1
            stencil_row<RP, LS, SPB>(G, mR + (lz0 + r + RP) * SPB + lx0 + RP, G.nfdmax, T, bc, w1, rk);
            unpack(*reinterpret_cast<const float4*>(sR1 + (lz0 + r + 2 * RP) * SPA + lx0 + 2 * RP), p0);
#pragma unroll
            for (int q = 0; q < 4; ++q) rkm[q] = finish_float(av[q], w1[q], rk[q], p0[q]);
            if (z + r == G.s_z) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = data_index(G, z + r, x + q);
                    if (j >= 0) {
                        const float d = seisB[j];
                        if (d != 0.0f) rkm[q] = d;
                    }
                }
            }
            // imaging, step k then step k-1 (Rel_Compen :503-517 / Rel_NonCompen :489-501)
            float r1v[4], r2v[4], sSv[4], sRv[4];
            unpack(a1, r1v);
            unpack(a2, r2v);
            unpack(aS, sSv);
            unpack(aR, sRv);
            if (compen) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    sSv[q] = __fadd_rn(sk[q], sSv[q]);
                    sRv[q] = __fadd_rn(rk[q], sRv[q]);
                    r1v[q] = __fmaf_rn(sRv[q], sSv[q], r1v[q]);
                    r2v[q] = __fmaf_rn(sk[q], sk[q], r2v[q]);
                    sSv[q] = __fadd_rn(skm[q], sSv[q]);
                    sRv[q] = __fadd_rn(rkm[q], sRv[q]);
                    r1v[q] = __fmaf_rn(sRv[q], sSv[q], r1v[q]);
                    r2v[q] = __fmaf_rn(skm[q], skm[q], r2v[q]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    r1v[q] = __fmaf_rn(rk[q], sk[q], r1v[q]);
                    r2v[q] = __fmaf_rn(sk[q], sk[q], r2v[q]);
                    r1v[q] = __fmaf_rn(rkm[q], skm[q], r1v[q]);
                    r2v[q] = __fmaf_rn(skm[q], skm[q], r2v[q]);
                }
            }
            auto put = [&](float* base, const float (&val)[4]) {
                *reinterpret_cast<float4*>(base + o) = make_float4(val[0], val[1], val[2], val[3]);
            };
            put(a.Sk, sk);
            put(a.Skm, skm);
            put(a.Rk, rk);
            put(a.Rkm, rkm);
            put(a.rel1, r1v);
            put(a.rel2, r2v);
            if (compen) {
                put(a.sumS, sSv);
                put(a.sumR, sRv);
            }
            o += G.pitch; cell += G.pitch;
        }
    }
}

}  // namespace rtmk
