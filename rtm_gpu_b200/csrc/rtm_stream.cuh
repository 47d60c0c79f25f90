// rtm_stream.cuh -- z-streaming, warp-specialised two-step kernels (sm_100a).
//
// The pair-stepping idea of bwd2_step_kernel (two time steps of an INNER region per pass over HBM)
// in the form the round-1 profiles asked for: instead of one 128 x 16 tile per CTA -- whose halo rows
// make phase A evaluate 24 rows for 16 and whose loads are exposed at every CTA start -- a CTA owns a
// SEGMENT, a 128-wide column of 8n rows, and marches down it in blocks of 8 rows:
//
//   warp 9 (producer)   one lane issues, one block ahead of the arithmetic, the TMA copies of the next
//                       stage: 8 new rows of the current fields (slot k+1, halo 2*RP in x), 8 rows of
//                       the previous fields (slot k+2, halo RP; they land IN the mid ring, where
//                       phase A overwrites them in place) and the 8 x 128 accumulator blocks;
//   warp 8 (halo)       evaluates the 2 x RP halo columns of phase A (16 float4 groups per block);
//   warps 0..7          phase A: warp w evaluates row w of mid block i (slot k) from the ring of
//                       current-field rows;  named barrier;  phase B: row w of out block i (slot k-1)
//                       from the mid ring, both imaging updates, all stores.
//
// Shared memory holds RINGS of rows, 3 slots of 8 rows plus a copy of slot 0 behind slot 2, so the
// +-RP row window of any row is contiguous: every stencil load keeps a compile-time offset from one
// base register, as in the tile kernels.  No consumer warp issues a global load for a wavefield or an
// accumulator (only the L2-resident velocity factor), nothing waits for HBM inside the arithmetic, and
// in z nothing is evaluated twice: phase A runs 8.5 rows per 8 output rows instead of 12.
// Every cell value is produced by the operation sequence of the single-step kernels (stencil_row,
// finish_float / finish_double, same injection / replacement / imaging order), so results stay
// bit-identical to single stepping (tests/test_gpu_fuse2.py, test_gpu_stream.py).
//
// The same kernel with BWD = false advances the FORWARD pass two steps per pass (one field, no
// accumulators): slot k from k-1/k-2, then slot k+1 from k/k-1, source term and gather recording in
// both phases.  Reference kernels restated: Add_Con :82-114 (forward); BKAdd_EFF_Con :287-320,
// BKAdd_Con :381-418, Rel_Compen / Rel_NonCompen :489-517 (backward; BKAdd_EFF :246-283 and BKAdd :339-380
// with the adaptive operator).  Operators up to radius 4: the fixed-length (Taylor) one, and -- LS = true, backward
// only -- the adaptive one when no velocity bin of the model needs more (per-cell coefficient rows from the padded
// global table, stencil_row_ls4); longer operators keep the tile kernels of rtm_kernels.cuh.
#pragma once
#include "rtm_kernels.cuh"

#include <vector>

namespace rtmk {

template <int RP> struct Strm {
    static_assert(RP == 4, "ring geometry below assumes RP == 4 and 8-row blocks");
    static constexpr int BR = 8;                       // rows per block = consumer warps
    static constexpr int NSLOT = 3;                    // ring slots (live: 2 blocks, incoming: 1)
    static constexpr int RING = NSLOT * BR + BR;       // ring rows: slots 0..2 + a copy of slot 0
    static constexpr int W1 = kTX + 4 * RP;            // pitch of the current-field ring (halo 2*RP)
    static constexpr int WM = kTX + 2 * RP;            // pitch of the mid ring (halo RP)
    static constexpr int CUR_BLK = BR * W1 * 4, MID_BLK = BR * WM * 4, ACC_BLK = BR * kTX * 4;
    static constexpr int CUR_RING = RING * W1 * 4, MID_RING = RING * WM * 4;
    static constexpr int kThreadsS = 32 * (BR + 2);    // 8 consumer warps + the halo warp + the producer warp
    static_assert(CUR_BLK % 128 == 0 && MID_BLK % 128 == 0 && CUR_RING % 128 == 0 && MID_RING % 128 == 0, "TMA destinations");
    __host__ __device__ static constexpr int bytes(bool bwd)
    {
        return (bwd ? 2 : 1) * (CUR_RING + MID_RING) + (bwd ? 2 * 4 * ACC_BLK : 0) + 64;
    }
};

// current fields (box W1 x 8), previous fields (box WM x 8), accumulators rel1, rel2, sumS, sumR (box 128 x 8)
struct StrmMaps { CUtensorMap cur[2], prev[2], acc[4]; };

struct StrmArgs {
    float *Ak[2];      // out: slot of phase A (backward: k, forward: k)      [0] source / forward field, [1] receiver
    float *Bk[2];      // out: slot of phase B (backward: k-1, forward: k+1)
    const int2* src;
    float  wavelet_a, wavelet_b;   // source term of the two steps
    int    k;                      // slot produced by phase A
    int    nshots;
    const int4* segs;              // (x0, z0, blocks, edges) : out rows [z0, z0 + 8*blocks), columns [x0, x0+128); edges: bit 0 left,
                                   // 1 right, 2 top, 3 bottom = the segment touches the thin frame there and stores its slot-k halo
    int    nseg;
    int    xend, zend;             // the streamed region ends here (exclusive): the last column / block may be partial
    FastDiv fd_nseg;
    const float* seis;             // backward: [S][NT][n], row k+1 imposed in phase A, row k in phase B
    float* gather;                 // forward: [S][NT][n] or null
    float *sumS, *sumR, *rel1, *rel2;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// RTM_MBAR_TRAP=1 (debug builds): bounded spin, a protocol error becomes a trap (CUDA error) instead of a hung GPU
#ifndef RTM_MBAR_TRAP
#define RTM_MBAR_TRAP 0
#endif
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity)
{
    if (!RTM_MBAR_TRAP) { mbar_wait(bar, parity); return; }
    const uint32_t addr = smem_u32(bar);
    for (int tries = 0;; ++tries) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (tries > (1 << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_consumers(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// float4 store of the cells x..x+3 that lie in [xbeg, xend)
__device__ __forceinline__ void store4c(float* dst, const float (&o)[4], int x, int xbeg, int xend)
{
    if (x >= xbeg && x + 3 < xend) {
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (x + q >= xbeg && x + q < xend) dst[q] = o[q];
    }
}

// RTM_STRM_BARSYNC=1 (checking builds): the A -> B barrier of stream2_kernel as `bar.sync 1, 288` instead of the split
// mbarrier arrive / wait.  compute-sanitizer's racecheck does not count an mbarrier phase completed by plain arrivals as
// an ordering between shared-memory accesses of different warps (it reports every phase-A write of the mid ring against
// the phase-B reads of the other warps); with the named barrier in the same place the same run is clean
// (profiles/README.md, "racecheck").
#ifndef RTM_STRM_BARSYNC
#define RTM_STRM_BARSYNC 0
#endif
#ifndef RTM_STRM_MINB_B
#define RTM_STRM_MINB_B 2
#endif
#ifndef RTM_STRM_MINB_F
#define RTM_STRM_MINB_F 3
#endif

#ifndef RTM_STRM_MAXREG_B
#define RTM_STRM_MAXREG_B 88   // 2 CTAs x 320 threads x 88 registers leave 9216 registers = one thin_frame CTA next to them
#endif
#ifndef RTM_STRM_MAXREG_LS
#define RTM_STRM_MAXREG_LS 96  // adaptive operator: per-cell coefficient rows and lengths on top (2 x 320 x 96 = 61440 registers)
#endif
template <int RP, bool BWD, bool LS = false>
__global__ void __launch_bounds__(Strm<RP>::kThreadsS, BWD ? RTM_STRM_MINB_B : RTM_STRM_MINB_F) __maxnreg__(BWD ? (LS ? RTM_STRM_MAXREG_LS : RTM_STRM_MAXREG_B) : 64)
stream2_kernel(const __grid_constant__ StrmMaps tm, const __grid_constant__ Geo G, const StrmArgs a)
{
    using T = Strm<RP>;
    constexpr int NF = BWD ? 2 : 1, BR = T::BR, W1 = T::W1, WM = T::WM;
    constexpr int CURF = T::RING * W1, MIDF = T::RING * WM;   // floats per field ring
    constexpr int COPY = T::NSLOT * BR;                       // row offset of the copy of slot 0
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float*    cur  = reinterpret_cast<float*>(smem_raw);                                        // [NF][RING][W1]
    float*    mid  = reinterpret_cast<float*>(smem_raw + NF * T::CUR_RING);                      // [NF][RING][WM]
    float*    accb = reinterpret_cast<float*>(smem_raw + NF * (T::CUR_RING + T::MID_RING));      // [2][4][BR][kTX]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NF * (T::CUR_RING + T::MID_RING) + (BWD ? 2 * 4 * T::ACC_BLK : 0));
    uint64_t* done = full + 3;
    uint64_t* midr = full + 5;   // "mid block i is complete" (all 9 warps arrive, all wait): the A -> B barrier, split in arrive / wait

    const int  shot = fast_div(blockIdx.x, a.fd_nseg);
    const int4 sg   = __ldg(a.segs + (blockIdx.x - shot * a.nseg));
    const int  x0 = sg.x, z0 = sg.y, n = sg.z, edges = sg.w;
    const int  tid = threadIdx.x, lane = tid & 31;
    const int  warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
    const long long so = (long long)shot * G.shot_stride + G.padL;
    const int2 src = a.src[shot];
    const bool compen = BWD && G.iCompen == 1;
    const int  nacc = compen ? 4 : 2;

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(full + i, 1);
        for (int i = 0; i < 2; ++i) mbar_init(done + i, BR);
        for (int i = 0; i < 2; ++i) mbar_init(midr + i, BR + 1);
    }
    __syncthreads();

    // stage s = current-field rows [z0-8+8s, +8) (+ the copy behind slot 2 when it lands in slot 0),
    // previous-field rows of mid block s-1 = [z0-RP+8(s-1), +8), accumulators of out block s-1 = rows [z0+8(s-2), +8)
    auto issue = [&](int s) {
        const int slot = s % 3;
        uint64_t* bar = full + slot;
        const bool has_prev = s >= 1, has_acc = BWD && s >= 2;
        uint32_t bytes = NF * T::CUR_BLK * (slot == 0 ? 2 : 1);
        if (has_prev) bytes += NF * T::MID_BLK;
        if (has_acc) bytes += nacc * T::ACC_BLK;
        mbar_expect_tx(bar, bytes);
        const int xc = G.padL + x0 - 2 * RP, zc = z0 - BR + BR * s;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            tma_load_3d(cur + f * CURF + slot * BR * W1, &tm.cur[f], bar, xc, zc, shot);
            if (slot == 0) tma_load_3d(cur + f * CURF + 3 * BR * W1, &tm.cur[f], bar, xc, zc, shot);
            if (has_prev) tma_load_3d(mid + f * MIDF + slot * BR * WM, &tm.prev[f], bar, xc + RP, z0 - RP + BR * (s - 1), shot);
        }
        if (has_acc)
            for (int i = 0; i < nacc; ++i)
                tma_load_3d(accb + ((s & 1) * 4 + i) * BR * kTX, &tm.acc[i], bar, xc + 2 * RP, z0 + BR * (s - 2), shot);
    };

    const float* AV = G.avel + G.padL;
    // adaptive operator (every length <= RP): a cell's coefficient row comes from the padded global table, addressed by its
    // velocity bin -- stencil_row's "staged" form with the rows in global memory (L1 hits: neighbouring cells share bins)
    const LsTable T0{};
    const unsigned short* BN = G.bins + G.padL;
    const uint2 nobins = make_uint2(0u, 0u);

    // Row of block `blk` handled in phase A: r (0..7); first mid column colm (float4 group); own: the group lies in the
    // segment's own columns (its values are stored, the halo groups' are not)
    int rA = warp, colm = RP + 4 * lane;
    bool workA = warp < BR;
    if (warp == BR) {            // halo warp: lanes 0..15 take the 2 halo groups of the 8 rows
        rA = lane >> 1; colm = (lane & 1) ? kTX + RP : 0; workA = lane < 2 * BR;
    }
    if (warp == BR + 1) {
        // ---- producer warp: one lane keeps the TMA copies one block ahead of the arithmetic.  It must not share a warp
        // with phase-A work: it waits for EVERY consumer warp to finish block i-1 before it can refill their slots, and
        // work queued behind that wait would make all consumers wait for it at the A -> B barrier.
        if (lane == 0) {
            issue(0); issue(1); issue(2);
            for (int i = 1; i + 2 <= n + 1; ++i) {
                mbar_wait_b(done + ((i - 1) & 1), ((i - 1) >> 1) & 1);   // every consumer warp is past B(i-1)
                issue(i + 2);
            }
        }
        return;
    }
    const int xA = x0 - RP + colm;              // global column of the first cell of the phase-A group
    const int xB = x0 + 4 * lane;               // phase B: the segment's own columns

    const int zlast = min(z0 + BR * n, a.zend); // rows of this segment that are stored: [z0, zlast)
    // Phase A also evaluates slot k on the RP cells around the segment.  Where those are thin-frame cells (the segment lies on
    // the edge of the streamed region) their values are the thin frame's slot k: stored here, so that the thin frame needs a
    // kernel of its own only for the second slot of the pair (which waits for the ring's one-way solution of slot k).
    const int zloA = (edges & 4) ? z0 - RP : z0, zhiA = (edges & 8) ? a.zend + RP : zlast;
    const int xloA = (edges & 1) ? x0 - RP : x0, xhiA = (edges & 2) ? a.xend + RP : min(x0 + kTX, a.xend);
    int slot_i = 0, slot_s = 1;                 // ring slots of stages i and i+1
    // rows of iteration 0 and their running addresses (every iteration moves BR rows down)
    const size_t rowstep = (size_t)BR * G.pitch;
    int zA = z0 - RP + rA, zO = z0 - BR + warp;
    const float* pavA = AV + (size_t)zA * G.pitch + xA;
    const float* pavB = AV + (ptrdiff_t)zO * G.pitch + xB;
    size_t soA = (size_t)so + (size_t)zA * G.pitch + xA, soB = (size_t)so + (size_t)((ptrdiff_t)zO * G.pitch) + xB;
    // velocity factor a = ((v*v)*tao2)*h2 of the two rows (L2-resident, shared by all shots), loaded one iteration ahead
    float4 avA4 = make_float4(0.f, 0.f, 0.f, 0.f), avB4 = avA4;
    uint2  bnA2 = nobins, bnB2 = nobins;        // ... and the velocity bins of the same cells (adaptive operator)
    if (workA && zA < G.NZ) {
        avA4 = __ldg(reinterpret_cast<const float4*>(pavA));
        if (LS) bnA2 = __ldg(reinterpret_cast<const uint2*>(BN + (pavA - AV)));
    }
    for (int i = 0; i <= n; ++i, zA += BR, zO += BR, pavA += rowstep, pavB += rowstep, soA += rowstep, soB += rowstep) {
        const int s = i + 1;
        const bool doA = workA;                                // (mid block n is needed whole: out block n reads RP rows past its end)
        const bool doB = warp < BR && i >= 1 && zO < zlast && xB < a.xend;   // (rows / groups past the region's end feed nobody)
        const float4 avAc = avA4, avBc = avB4;                 // this iteration's; the next iteration's go in flight now
        const uint2  bnAc = bnA2, bnBc = bnB2;
        if (i < n) {
            if (workA && zA + BR < G.NZ) {
                avA4 = __ldg(reinterpret_cast<const float4*>(pavA + rowstep));
                if (LS) bnA2 = __ldg(reinterpret_cast<const uint2*>(BN + (pavA + rowstep - AV)));
            }
            if (warp < BR && zO + BR < zlast && xB < a.xend) {
                avB4 = __ldg(reinterpret_cast<const float4*>(pavB + rowstep));
                if (LS) bnB2 = __ldg(reinterpret_cast<const uint2*>(BN + (pavB + rowstep - AV)));
            }
        }

        if (i == 0) mbar_wait_b(full + 0, 0);
        mbar_wait_b(full + slot_s, (s / 3) & 1);

        // ---- phase A: mid block i (slot k / forward: k), rows [z0-RP+8i, +8)
        float oA[NF][4];
        LsCells LC{};
        if (LS) LC = ls_cells(G, bnAc);   // (all lanes: warp-uniform operator length inside)
        if (doA) {
            // centre row in the current-field ring: rows 0..RP-1 of the mid block lie in stage i, the others in stage i+1
            int rc = rA < RP ? slot_i * BR + rA + RP : slot_s * BR + rA - RP;
            if (rc < RP) rc += COPY;
            const int rm = slot_s * BR + rA;              // mid block i lives in the slot of stage i+1
            float av[4];
            unpack(avAc, av);
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const float* pc = cur + f * CURF + rc * W1 + colm + RP;
                float*       pm = mid + f * MIDF + rm * WM + colm;
                float w1[4], p1[4], p0[4];
                float (&o)[4] = oA[f];
                if constexpr (LS) stencil_row_ls4<W1>(G, pc, LC, w1, p1);
                else stencil_row<RP, false, W1>(G, pc, G.nfdmax, T0, nobins, w1, p1);
                unpack(*reinterpret_cast<const float4*>(pm), p0);
                if (f == 0) {   // source field / forward field: double final sum (Add_Con, BKAdd_EFF_Con), + wavelet
#pragma unroll
                    for (int q = 0; q < 4; ++q) o[q] = finish_double(av[q], w1[q], p1[q], p0[q]);
                    if (zA == src.x && src.y >= xA && src.y < xA + 4) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (xA + q == src.y) o[q] = __fadd_rn(o[q], a.wavelet_a);
                    }
                } else {        // receiver field: float final sum (BKAdd_Con), data replacement
#pragma unroll
                    for (int q = 0; q < 4; ++q) o[q] = finish_float(av[q], w1[q], p1[q], p0[q]);
                    if (zA == G.s_z) {
                        const float* seisA = a.seis + ((size_t)shot * G.NT + (a.k + 1)) * G.n;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int j = data_index(G, zA, xA + q);
                            if (j >= 0) {
                                const float d = seisA[j];
                                if (d != 0.0f) o[q] = d;
                            }
                        }
                    }
                }
                const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(pm) = o4;
                if (slot_s == 0) *reinterpret_cast<float4*>(pm + 3 * BR * WM) = o4;   // the copy behind slot 2
            }
            fence_async_smem();   // these slots are refilled by TMA later (generic -> async proxy order)
        }
        // A -> B barrier, split: arrive now, store this warp's slot-k values, then wait for the other warps' rows
        __syncwarp();
#if !RTM_STRM_BARSYNC
        if (lane == 0) mbar_arrive(midr + (i & 1));
#endif
        if (doA && zA >= zloA && zA < zhiA && xA + 3 >= xloA && xA < xhiA) {
#pragma unroll
            for (int f = 0; f < NF; ++f) store4c(a.Ak[f] + soA, oA[f], xA, xloA, xhiA);
            if (!BWD && a.gather && zA == G.s_z) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = (xA + q >= xloA && xA + q < xhiA) ? data_index(G, zA, xA + q) : -1;
                    if (j >= 0) a.gather[((size_t)shot * G.NT + a.k) * G.n + j] = oA[0][q];
                }
            }
        }
#if RTM_STRM_BARSYNC
        bar_consumers(32 * (BR + 1));   // checking build: the same A -> B barrier as a named hardware barrier
#else
        mbar_wait_b(midr + (i & 1), (i >> 1) & 1);
#endif

        // ---- phase B: out block i (slot k-1 / forward: k+1), rows [z0+8(i-1), +8)
        if (LS) LC = ls_cells(G, bnBc);
        if (doB) {
            int rb = warp < RP ? slot_i * BR + warp + RP : slot_s * BR + warp - RP;   // centre row in the mid ring
            if (rb < RP) rb += COPY;
            const int rp = slot_i * BR + warp;                                         // same row, current-field ring (P0)
            float av[4];
            unpack(avBc, av);
            float ok[NF][4], okm[NF][4];
            const size_t o = soB;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                float w1[4], p0[4];
                if constexpr (LS) stencil_row_ls4<WM>(G, mid + f * MIDF + rb * WM + RP + 4 * lane, LC, w1, ok[f]);
                else stencil_row<RP, false, WM>(G, mid + f * MIDF + rb * WM + RP + 4 * lane, G.nfdmax, T0, nobins, w1, ok[f]);
                unpack(*reinterpret_cast<const float4*>(cur + f * CURF + rp * W1 + 2 * RP + 4 * lane), p0);
                if (f == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) okm[f][q] = finish_double(av[q], w1[q], ok[f][q], p0[q]);
                    if (zO == src.x && src.y >= xB && src.y < xB + 4) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (xB + q == src.y) okm[f][q] = __fadd_rn(okm[f][q], a.wavelet_b);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) okm[f][q] = finish_float(av[q], w1[q], ok[f][q], p0[q]);
                    if (zO == G.s_z) {
                        const float* seisB = a.seis + ((size_t)shot * G.NT + a.k) * G.n;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int j = data_index(G, zO, xB + q);
                            if (j >= 0) {
                                const float d = seisB[j];
                                if (d != 0.0f) okm[f][q] = d;
                            }
                        }
                    }
                }
                store4c(a.Bk[f] + o, okm[f], xB, 0, a.xend);
            }
            if constexpr (!BWD) {
                if (a.gather && zO == G.s_z) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int j = xB + q < a.xend ? data_index(G, zO, xB + q) : -1;
                        if (j >= 0) a.gather[((size_t)shot * G.NT + a.k + 1) * G.n + j] = okm[0][q];
                    }
                }
            } else {
                // imaging, step k then step k-1 (Rel_Compen :503-517 / Rel_NonCompen :489-501); S = [0], R = [1]
                constexpr int F1 = NF - 1;   // (NF == 2 here; keeps the forward instantiation well-formed)
                const float* ab = accb + ((s & 1) * 4) * BR * kTX + warp * kTX + 4 * lane;
                float r1v[4], r2v[4], sSv[4], sRv[4];
                unpack(*reinterpret_cast<const float4*>(ab), r1v);
                unpack(*reinterpret_cast<const float4*>(ab + BR * kTX), r2v);
                if (compen) {
                    unpack(*reinterpret_cast<const float4*>(ab + 2 * BR * kTX), sSv);
                    unpack(*reinterpret_cast<const float4*>(ab + 3 * BR * kTX), sRv);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        sSv[q] = __fadd_rn(ok[0][q], sSv[q]);
                        sRv[q] = __fadd_rn(ok[F1][q], sRv[q]);
                        r1v[q] = __fmaf_rn(sRv[q], sSv[q], r1v[q]);
                        r2v[q] = __fmaf_rn(ok[0][q], ok[0][q], r2v[q]);
                        sSv[q] = __fadd_rn(okm[0][q], sSv[q]);
                        sRv[q] = __fadd_rn(okm[F1][q], sRv[q]);
                        r1v[q] = __fmaf_rn(sRv[q], sSv[q], r1v[q]);
                        r2v[q] = __fmaf_rn(okm[0][q], okm[0][q], r2v[q]);
                    }
                    store4c(a.sumS + o, sSv, xB, 0, a.xend);
                    store4c(a.sumR + o, sRv, xB, 0, a.xend);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        r1v[q] = __fmaf_rn(ok[F1][q], ok[0][q], r1v[q]);
                        r2v[q] = __fmaf_rn(ok[0][q], ok[0][q], r2v[q]);
                        r1v[q] = __fmaf_rn(okm[F1][q], okm[0][q], r1v[q]);
                        r2v[q] = __fmaf_rn(okm[0][q], okm[0][q], r2v[q]);
                    }
                }
                store4c(a.rel1 + o, r1v, xB, 0, a.xend);
                store4c(a.rel2 + o, r2v, xB, 0, a.xend);
            }
        }
        // this warp is done with iteration i: its slots may be refilled by TMA
        __syncwarp();
        if (warp < BR && lane == 0) mbar_arrive(done + (i & 1));
        slot_i = slot_s;
        slot_s = slot_s == 2 ? 0 : slot_s + 1;
    }
}

// ------------------------------------------------------------------------------------
// The forward pass, one slot per pass, in the same streaming form -- `stream1_fwd_kernel`.
//
// The forward pass gains nothing from two slots per pass (one field leaves too little arithmetic per block to pay
// for the second phase and its barrier, and the pair pipeline adds ring and thin-frame launches), but the tile
// kernel's per-CTA TMA wait and instruction-bound row loop make it clock-sensitive: under the 1 kW power cap of a
// full-length run it falls from 142 to 154-158 us per step.  Here a CTA marches down a column segment of the WHOLE
// interior (no thin frame: single steps need none): the producer warp keeps 8 rows of the current field (box 136 x 8,
// ring of 3 slots + copy as above) and 8 x 128 of the previous field one block ahead; consumer warp w evaluates row w
// of every block and stores it.  No warp depends on another warp's results, so the only synchronisation is with the
// TMA engine.  Add_Con :82-114 (double final sum, additive source, gather recording).
// Measured (profiles/r2_c9_*): bit-exact, but 134 us per 32-shot step against 124 us for the tile kernel's interior
// launch -- 70 % of the issue slots busy: one row of one field per warp and block does not amortise the per-block
// waits and address work the way the tile kernel's 4-row loop does.  Kept behind RTM_STREAM1_FWD=1.
// ------------------------------------------------------------------------------------
template <int RP> struct Strm1 {
    static constexpr int BR = 8, NSLOT = 3, RING = NSLOT * BR + BR;
    static constexpr int W1 = kTX + 2 * RP;
    static constexpr int CUR_BLK = BR * W1 * 4, P0_BLK = BR * kTX * 4, CUR_RING = RING * W1 * 4;
    static constexpr int kThreadsS = 32 * (BR + 1);
    static_assert(CUR_BLK % 128 == 0 && CUR_RING % 128 == 0, "TMA destinations");
    __host__ __device__ static constexpr int bytes() { return CUR_RING + NSLOT * P0_BLK + 64; }
};
struct Strm1Maps { CUtensorMap cur, prev; };   // boxes (128+2RP) x 8 and 128 x 8

#ifndef RTM_STRM1_MINB
#define RTM_STRM1_MINB 4
#endif
template <int RP>
__global__ void __launch_bounds__(Strm1<RP>::kThreadsS, RTM_STRM1_MINB)
stream1_fwd_kernel(const __grid_constant__ Strm1Maps tm, const __grid_constant__ Geo G, const StrmArgs a)
{
    using T = Strm1<RP>;
    constexpr int BR = T::BR, W1 = T::W1, COPY = T::NSLOT * BR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float*    cur  = reinterpret_cast<float*>(smem_raw);                   // [RING][W1]
    float*    p0s  = reinterpret_cast<float*>(smem_raw + T::CUR_RING);     // [NSLOT][BR][kTX]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + T::CUR_RING + T::NSLOT * T::P0_BLK);
    uint64_t* done = full + 3;
    const int  shot = fast_div(blockIdx.x, a.fd_nseg);
    const int4 sg   = __ldg(a.segs + (blockIdx.x - shot * a.nseg));
    const int  x0 = sg.x, z0 = sg.y, n = sg.z;
    const int  tid = threadIdx.x, lane = tid & 31;
    const int  warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const long long so = (long long)shot * G.shot_stride + G.padL;
    const int2 src = a.src[shot];
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(full + i, 1);
        for (int i = 0; i < 2; ++i) mbar_init(done + i, BR);
    }
    __syncthreads();
    // stage s: current-field rows [z0-RP+8s, +8) (+ the copy when it lands in slot 0), previous-field rows of out block s-1
    auto issue = [&](int s) {
        const int slot = s % 3;
        uint64_t* bar = full + slot;
        mbar_expect_tx(bar, T::CUR_BLK * (slot == 0 ? 2 : 1) + (s >= 1 ? T::P0_BLK : 0));
        const int xc = G.padL + x0 - RP, zc = z0 - RP + BR * s;
        tma_load_3d(cur + slot * BR * W1, &tm.cur, bar, xc, zc, shot);
        if (slot == 0) tma_load_3d(cur + 3 * BR * W1, &tm.cur, bar, xc, zc, shot);
        if (s >= 1) tma_load_3d(p0s + slot * BR * kTX, &tm.prev, bar, xc + RP, z0 + BR * (s - 1), shot);
    };
    if (warp == BR) {   // producer warp
        if (lane == 0) {
            issue(0); issue(1);
            if (n >= 2) issue(2);
            for (int i = 1; i + 2 <= n; ++i) {
                mbar_wait_b(done + ((i - 1) & 1), ((i - 1) >> 1) & 1);
                issue(i + 2);
            }
        }
        return;
    }
    const LsTable T0{};
    const int x = x0 + 4 * lane;
    const int zlast = min(z0 + BR * n, a.zend);
    const size_t rowstep = (size_t)BR * G.pitch;
    int z = z0 + warp;
    const float* pav = G.avel + G.padL + (size_t)z * G.pitch + x;
    size_t soz = (size_t)so + (size_t)z * G.pitch + x;
    const bool colok = x < a.xend;
    float4 av4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (colok && z < zlast) av4 = __ldg(reinterpret_cast<const float4*>(pav));
    int slot_i = 0, slot_s = 1;
    for (int i = 0; i < n; ++i, z += BR, pav += rowstep, soz += rowstep) {
        const int s = i + 1;
        const float4 avc = av4;
        if (i + 1 < n && colok && z + BR < zlast) av4 = __ldg(reinterpret_cast<const float4*>(pav + rowstep));
        if (i == 0) mbar_wait_b(full + 0, 0);
        mbar_wait_b(full + slot_s, (s / 3) & 1);
        if (colok && z < zlast) {
            int rc = warp < RP ? slot_i * BR + warp + RP : slot_s * BR + warp - RP;
            if (rc < RP) rc += COPY;
            float av[4], w1[4], p1[4], p0[4], o[4];
            unpack(avc, av);
            stencil_row<RP, false, W1>(G, cur + rc * W1 + RP + 4 * lane, G.nfdmax, T0, make_uint2(0u, 0u), w1, p1);
            unpack(*reinterpret_cast<const float4*>(p0s + (slot_s * BR + warp) * kTX + 4 * lane), p0);
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = finish_double(av[q], w1[q], p1[q], p0[q]);
            if (z == src.x && src.y >= x && src.y < x + 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (x + q == src.y) o[q] = __fadd_rn(o[q], a.wavelet_a);
            }
            store4c(a.Ak[0] + soz, o, x, 0, a.xend);
            if (a.gather && z == G.s_z) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = x + q < a.xend ? data_index(G, z, x + q) : -1;
                    if (j >= 0) a.gather[((size_t)shot * G.NT + a.k) * G.n + j] = o[q];
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(done + (i & 1));
        slot_i = slot_s;
        slot_s = slot_s == 2 ? 0 : slot_s + 1;
    }
}

// ------------------------------------------------------------------------------------
// The thin frame: the RP interior cells next to the absorbing ring.  They cannot advance two slots per
// pass (slot k-1 there needs the ring's one-way solution of slot k), so they -- and only they, not the
// 128 x 16 tiles around them as in the tile form -- are stepped singly, between the two-step passes of
// the streamed region.  One CTA of 128 threads per strip tile:
//   kind 0 (top / bottom strip)   4 rows x 128 columns: warp = row, lane = float4 group
//   kind 1 (left / right strip)  64 rows x 2 groups   : thread = (row, group); the strip's RP columns need
//                                 not start on a float4 boundary, so two groups cover them
// The current field(s) arrive as one TMA box with the stencil halo (two box shapes), everything else is
// the single-step arithmetic of fwd_step_kernel / bwd_step_kernel on one float4 group.
// ------------------------------------------------------------------------------------
struct ThinTile { int x0, z0, xbeg, xend, zend, kind, pad0, pad1; };   // cells [z0,zend) x ([x0,..) clipped to [xbeg,xend))
struct ThinMaps { CUtensorMap row[2], col[2]; };                      // boxes (128+2RP) x (4+2RP) and (8+2RP) x (64+2RP)

struct ThinArgs {
    const float* P0[2];    // previous slot (forward: k-2, backward: k+2) of the forward/source field [0] and the receiver field [1]
    float*       P2[2];    // out (slot k)
    const int2*  src;
    float        wavelet;
    int          k, nshots;
    const ThinTile* tiles;
    int          ntiles;
    FastDiv      fd_ntiles;
    const float* seis;     // backward: row k+1 is imposed
    float*       gather;   // forward
    float *sumS, *sumR, *rel1, *rel2;
    int    twice;          // backward: also apply the imaging update of the CURRENT slot (k+1) first -- its field values were
                           // stored by the streaming kernel's halo, its imaging update was left to this kernel
};

template <int RP> struct Thin {
    static constexpr int kThreadsT = 128;
    static constexpr int ROW_W = kTX + 2 * RP, ROW_H = 4 + 2 * RP;     // kind 0 box
    static constexpr int COL_W = 8 + 2 * RP, COL_H = 64 + 2 * RP;      // kind 1 box
    static constexpr int FLOATS = ROW_W * ROW_H > COL_W * COL_H ? ROW_W * ROW_H : COL_W * COL_H;
    __host__ __device__ static constexpr int bytes(bool bwd) { return (bwd ? 2 : 1) * ((FLOATS * 4 + 127) / 128 * 128) + 16; }
};

template <int RP, bool BWD, bool LS = false>
__global__ void __launch_bounds__(Thin<RP>::kThreadsT)
thin_frame_kernel(const __grid_constant__ ThinMaps tm, const __grid_constant__ Geo G, const ThinArgs a)
{
    using T = Thin<RP>;
    constexpr int NF = BWD ? 2 : 1;
    constexpr int FSTRIDE = (T::FLOATS * 4 + 127) / 128 * 128 / 4;     // floats between the two field tiles
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float*    tile = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar  = reinterpret_cast<uint64_t*>(smem_raw + NF * FSTRIDE * 4);
    const int shot = fast_div(blockIdx.x, a.fd_ntiles);
    const ThinTile tl = a.tiles[blockIdx.x - shot * a.ntiles];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool colk = tl.kind == 1;
    const int  SP = colk ? T::COL_W : T::ROW_W;

    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, NF * (colk ? T::COL_W * T::COL_H : T::ROW_W * T::ROW_H) * 4);
#pragma unroll
        for (int f = 0; f < NF; ++f)
            tma_load_3d(tile + f * FSTRIDE, colk ? &tm.col[f] : &tm.row[f], bar, G.padL + tl.x0 - RP, tl.z0 - RP, shot);
    }
    const int lr = colk ? tid >> 1 : warp, lg = colk ? tid & 1 : lane;   // row and float4 group inside the tile
    const int z = tl.z0 + lr, x = tl.x0 + 4 * lg;
    const bool work = z < tl.zend && x + 3 >= tl.xbeg && x < tl.xend;
    const long long so = (long long)shot * G.shot_stride + G.padL;
    const int2 src = a.src[shot];
    const size_t cell = (size_t)z * G.pitch + x;
    float4 av4 = make_float4(0.f, 0.f, 0.f, 0.f), p04[2] = {av4, av4}, a1 = av4, a2 = av4, aS = av4, aR = av4;
    const bool compen = BWD && G.iCompen == 1;
    uint2 bn2 = make_uint2(0u, 0u);
    if (work) {   // (whole float4 groups: the row pitch leaves room past the last column)
        av4 = __ldg(reinterpret_cast<const float4*>(G.avel + G.padL + cell));
        if (LS) bn2 = __ldg(reinterpret_cast<const uint2*>(G.bins + G.padL + cell));
#pragma unroll
        for (int f = 0; f < NF; ++f) p04[f] = *reinterpret_cast<const float4*>(a.P0[f] + so + cell);
        if (BWD) {
            a1 = *reinterpret_cast<const float4*>(a.rel1 + so + cell);
            a2 = *reinterpret_cast<const float4*>(a.rel2 + so + cell);
            if (compen) {
                aS = *reinterpret_cast<const float4*>(a.sumS + so + cell);
                aR = *reinterpret_cast<const float4*>(a.sumR + so + cell);
            }
        }
    }
    mbar_wait_b(bar, 0);
    if (!work) return;
    LsTable T0{};
    if (LS) { T0.rows = G.ls_rows; T0.len = G.ls_len; T0.bmin = 0; T0.bmax = G.ls_nbins - 1; T0.staged = true; }
    float av[4], o[NF][4], c1[NF][4];   // c1: the current fields at the cell
    unpack(av4, av);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        float w1[4], p0[4];
        float (&p1)[4] = c1[f];
        stencil_row<RP, LS, 0>(G, tile + f * FSTRIDE + (lr + RP) * SP + 4 * lg + RP, G.nfdmax, T0, bn2, w1, p1, SP);
        unpack(p04[f], p0);
        if (f == 0) {   // forward field (Add_Con) / reconstructed source field (BKAdd_EFF_Con): double final sum, source term
#pragma unroll
            for (int q = 0; q < 4; ++q) o[f][q] = finish_double(av[q], w1[q], p1[q], p0[q]);
            if (z == src.x && src.y >= x && src.y < x + 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (x + q == src.y) o[f][q] = __fadd_rn(o[f][q], a.wavelet);
            }
        } else {        // receiver field (BKAdd_Con): float final sum, data replacement
#pragma unroll
            for (int q = 0; q < 4; ++q) o[f][q] = finish_float(av[q], w1[q], p1[q], p0[q]);
            if (z == G.s_z) {
                const float* seis_row = a.seis + ((size_t)shot * G.NT + (a.k + 1)) * G.n;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = data_index(G, z, x + q);
                    if (j >= 0) {
                        const float d = seis_row[j];
                        if (d != 0.0f) o[f][q] = d;
                    }
                }
            }
        }
        store4c(a.P2[f] + so + cell, o[f], x, tl.xbeg, tl.xend);
    }
    if constexpr (!BWD) {
        if (a.gather && z == G.s_z) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = (x + q >= tl.xbeg && x + q < tl.xend) ? data_index(G, z, x + q) : -1;
                if (j >= 0) a.gather[((size_t)shot * G.NT + a.k) * G.n + j] = o[0][q];
            }
        }
    } else {
        constexpr int F1 = NF - 1;
        float r1v[4], r2v[4], sSv[4], sRv[4];
        unpack(a1, r1v);
        unpack(a2, r2v);
        unpack(aS, sSv);
        unpack(aR, sRv);
        if (a.twice) {  // slot k+1 first (time order), with the current fields at the cell
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (compen) {
                    sSv[q] = __fadd_rn(c1[0][q], sSv[q]);
                    sRv[q] = __fadd_rn(c1[F1][q], sRv[q]);
                    r1v[q] = __fmaf_rn(sRv[q], sSv[q], r1v[q]);
                } else {
                    r1v[q] = __fmaf_rn(c1[F1][q], c1[0][q], r1v[q]);
                }
                r2v[q] = __fmaf_rn(c1[0][q], c1[0][q], r2v[q]);
            }
        }
        if (compen) {   // Rel_Compen :503-517
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                sSv[q] = __fadd_rn(o[0][q], sSv[q]);
                sRv[q] = __fadd_rn(o[F1][q], sRv[q]);
                r1v[q] = __fmaf_rn(sRv[q], sSv[q], r1v[q]);
                r2v[q] = __fmaf_rn(o[0][q], o[0][q], r2v[q]);
            }
            store4c(a.sumS + so + cell, sSv, x, tl.xbeg, tl.xend);
            store4c(a.sumR + so + cell, sRv, x, tl.xbeg, tl.xend);
        } else {        // Rel_NonCompen :489-501
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                r1v[q] = __fmaf_rn(o[F1][q], o[0][q], r1v[q]);
                r2v[q] = __fmaf_rn(o[0][q], o[0][q], r2v[q]);
            }
        }
        store4c(a.rel1 + so + cell, r1v, x, tl.xbeg, tl.xend);
        store4c(a.rel2 + so + cell, r2v, x, tl.xbeg, tl.xend);
    }
}

// ------------------------------------------------------------------------------------
// Regions of the streaming form (host side; checked on the CPU by tests/test_launch_geometry.py), for the forward and
// the backward pass alike:
//   ring        the N2 outermost cells                    ring_kernel / ring tiles of the single-step kernels
//   thin frame  the RP interior cells next to the ring    thin_frame_kernel, stepped singly
//   streamed    the rest [C0, xe) x [R0, ze)              stream2_kernel, two slots per pass: 128-wide columns (the last may
//               be partial) of 8-row blocks (the last may be partial), cut into segments of at most seg_blocks blocks;
//               "ib" = the columns / pieces next to the thin frame (the inner-inner segments' halo never reaches a cell
//               stepped singly), "ii" = everything inside them
// ------------------------------------------------------------------------------------
struct StreamRegions {
    bool ok = false;                 // the grid is large enough for the streaming form
    std::vector<int4> ii, ib;        // segments (x0, z0, blocks, edges), see StrmArgs::segs
    std::vector<ThinTile> thin;
    double stream_cells = 0;         // cells per shot advanced two slots per pass
    int C0 = 0, R0 = 0, xe = 0, ze = 0;
};
__host__ inline StreamRegions make_stream_regions(const Geo& G, int RPc, int seg_blocks)
{
    using T = Strm<4>;
    StreamRegions r;
    const int C0 = G.N2 + RPc, R0 = G.N2 + RPc, xe = G.NX - G.N2 - RPc, ze = G.NZ - G.N2 - RPc;
    r.C0 = C0; r.R0 = R0; r.xe = xe; r.ze = ze;
    const int ncol = (xe - C0 + kTX - 1) / kTX, nblk = (ze - R0 + T::BR - 1) / T::BR;
    const int nright = (xe - C0 - (ncol - 1) * kTX >= 2 * RPc) ? 1 : 2;   // right "ib" columns: at least 2*RP cells
    if (!(ncol >= 1 + nright && nblk >= 2)) return r;
    // a column's blocks in pieces of (nearly) equal length <= seg_blocks.  "ib": the whole first / last column(s)
    // and, of every other column, its first and last piece -- long pieces, so the ib launch streams as
    // efficiently as the ii launch; their outer 8 rows / 2*RP columns are what the ii segments' halo may reach
    for (int col = 0; col < ncol; ++col) {
        const bool edge = col == 0 || col >= ncol - nright;
        const int pieces = (nblk + seg_blocks - 1) / seg_blocks;
        std::vector<int2> pc;   // (first block, blocks)
        for (int p = 0, b = 0; p < pieces; ++p) {
            const int len = nblk / pieces + (p < nblk % pieces ? 1 : 0);
            pc.push_back(make_int2(b, len));
            b += len;
        }
        int nlast = 1;          // trailing pieces that go to ib: at least 8 valid rows
        while (nlast < (int)pc.size() && ze - (R0 + pc[pc.size() - nlast].x * T::BR) < T::BR) ++nlast;
        for (int p = 0; p < (int)pc.size(); ++p) {
            const bool border = edge || p == 0 || p >= (int)pc.size() - nlast;
            const int edges = (col == 0 ? 1 : 0) | (col == ncol - 1 ? 2 : 0) | (p == 0 ? 4 : 0) | (p == (int)pc.size() - 1 ? 8 : 0);
            (border ? r.ib : r.ii).push_back(make_int4(C0 + col * kTX, R0 + pc[p].x * T::BR, pc[p].y, edges));
        }
    }
    r.stream_cells = (double)(xe - C0) * (ze - R0);
    // thin-frame strips
    for (int x0 = G.N2; x0 < G.NX - G.N2; x0 += kTX) {   // top and bottom: RP rows, all interior columns
        r.thin.push_back(ThinTile{x0, G.N2, G.N2, G.NX - G.N2, G.N2 + RPc, 0, 0, 0});
        r.thin.push_back(ThinTile{x0, ze, G.N2, G.NX - G.N2, ze + RPc, 0, 0, 0});
    }
    const int xr = xe - ((G.padL + xe) % 4);              // float4-aligned start of the right strip's groups
    for (int z0 = R0; z0 < ze; z0 += 64) {                // left and right: RP columns, the rows in between
        r.thin.push_back(ThinTile{G.N2, z0, G.N2, C0, z0 + 64 < ze ? z0 + 64 : ze, 1, 0, 0});
        r.thin.push_back(ThinTile{xr, z0, xe, G.NX - G.N2, z0 + 64 < ze ? z0 + 64 : ze, 1, 0, 0});
    }
    r.ok = true;
    return r;
}

}  // namespace rtmk
