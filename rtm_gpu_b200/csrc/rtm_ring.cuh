// rtm_ring.cuh -- the hybrid absorbing ring as a kernel of its own (both operators).
//
// ring_tile<> of rtm_kernels.cuh spends most of its instructions outside the arithmetic (round-1
// attribution, profiles/r1_final_ring_attribution.txt: 40 % in the 4-byte cp.async fills with their
// mirror / clipping address work, 29 % in the one-way phase, where every cell re-derives its class, its
// one-way coefficients -- a division, a square root and a reciprocal at the corners -- and its
// boundary-strip slot every time step).  Here, per ring tile:
//   * the current field with its stencil halo, the previous field and the velocity factor arrive as THREE
//     TMA BOXES (out-of-array parts zero-filled, then mirrored in shared memory: the reference's mirror
//     rule :65-68 only ever applies inside the ring);
//   * everything about a ring cell that does not change from step to step is computed ONCE PER MODEL by
//     ring_coef_kernel with the very intrinsic sequence ring_tile<> uses -- the reciprocal, the tv / r1
//     factor, the c2 factor (with the reference's mis-indexed velocity, Q2), the blend weight, the cell
//     class and its boundary-strip offset -- and read back as one float4 + one int per cell;
//   * tile geometry is uniform (output rectangle grown by one cell, never clipped: values computed outside
//     the array are never consumed), so all loops have compile-time-friendly shapes.
// The arithmetic per cell is ring_tile<>'s, operation by operation (Hybrid1 :116-158 edge and corner
// formulas, Hybrid2 :160-183 blend, Hybrid3 / Equal strip save, BKEqual strip restore one step early),
// so results are bit-identical (tests/test_gpu_parity.py, test_gpu_shapes.py run both forms).
#pragma once
#include "rtm_kernels.cuh"
#include "rtm_stream.cuh"

#include <algorithm>

namespace rtmk {

// Tiling of the ring for ring_kernel.  A TMA box must start on a 16-byte boundary in global memory and the float4
// groups of the row stencil on one in shared memory, so the COMPUTE rectangle of every tile starts at a column x with
// (padL + x) % 4 == 0: band tiles are 120 columns wide, grown by 4 columns on both sides (128 = 32 groups), the first one
// starting left of the array; side tiles take 16 (N2 + 2, rounded) columns from an aligned column at or left of their
// first needed one.  Rows carry no such constraint: grown by one row.
constexpr int kRing2TX = 120;
struct RingGeo {            // per context
    int RP;                 // rounded operator radius (x halo of the box)
    int R;                  // operator radius (z halo)
    int chB, cwB, spB;      // band tiles:  compute rows N2+2, compute width 128, box pitch 128+2RP
    int chS, cwS, spS;      // side tiles:  compute rows 128,  compute width (N2+5 rounded up to 4), box pitch cwS+2RP
    int xshift;             // band tile i covers columns [120 i - xshift, 120 (i+1) - xshift)
    int cxl, cxr;           // first compute column of the left / right side tiles
    int nband, nside, ntiles;
    FastDiv fd_ntiles;
    int cells;              // cells per tile in the coefficient arrays (N2 * 126, rounded up to 4)
    int n1, nc;             // floats reserved for the halo box / for each compute-rectangle array (both tile kinds, 128-byte multiples)
    __host__ __device__ int smem_bytes() const { return (n1 + 3 * nc) * 4 + 16; }
};
__host__ inline RingGeo make_ring_geo(const Geo& G, int R, int RP)
{
    RingGeo g;
    const int N2 = G.N2;
    g.RP = RP; g.R = R;
    g.chB = N2 + 2; g.cwB = kRing2TX + 8; g.spB = g.cwB + 2 * RP;
    g.chS = kRingTX + 2; g.cwS = (N2 + 5 + 3) / 4 * 4; g.spS = g.cwS + 2 * RP;
    g.xshift = G.padL % 4;
    g.cxl = -(G.padL % 4 ? G.padL % 4 : 4);
    g.cxr = (G.NX - N2 - 1) - (G.padL + G.NX - N2 - 1) % 4;
    g.nband = (G.NX + g.xshift + kRing2TX - 1) / kRing2TX;
    g.nside = (G.mod_NZ + kRingTX - 1) / kRingTX;
    g.ntiles = 2 * g.nband + 2 * g.nside;
    g.fd_ntiles = make_fastdiv(g.ntiles);
    g.cells = (N2 * kRingTX + 3) / 4 * 4;
    auto up32 = [](int n) { return (n + 31) / 32 * 32; };
    g.n1 = up32(std::max((g.chB + 2 * R) * g.spB, (g.chS + 2 * R) * g.spS));
    g.nc = up32(std::max(g.chB * g.cwB, g.chS * g.cwS));
    return g;
}
// output rectangle (clipped to the array) and compute-rectangle origin of a tile
__host__ __device__ inline RingRect ring2_rect(const Geo& G, const RingGeo& g, int tile, int* cz0, int* cx0)
{
    RingRect r;
    const int N2 = G.N2;
    if (tile < 2 * g.nband) {
        const bool top = tile < g.nband;
        const int  i   = top ? tile : tile - g.nband;
        r.za = top ? 0 : G.NZ - N2;
        r.zb = r.za + N2;
        const int xa = i * kRing2TX - g.xshift;
        r.xa = xa < 0 ? 0 : xa;
        r.xb = xa + kRing2TX < G.NX ? xa + kRing2TX : G.NX;
        *cz0 = r.za - 1; *cx0 = xa - 4;
    } else {
        tile -= 2 * g.nband;
        const bool left = tile < g.nside;
        const int  i    = left ? tile : tile - g.nside;
        r.xa = left ? 0 : G.NX - N2;
        r.xb = r.xa + N2;
        r.za = N2 + i * kRingTX;
        r.zb = r.za + kRingTX < G.NZ - N2 ? r.za + kRingTX : G.NZ - N2;
        *cz0 = r.za - 1; *cx0 = left ? g.cxl : g.cxr;
    }
    return r;
}
struct RingMaps { CUtensorMap p1b, p1s, p0b, p0s, avb, avs; };   // current (with halo) / previous field, velocity factor; band / side boxes

// meta word of a ring cell: bits 0-2 class (0 top, 1 bottom, 2 left, 3 right edge; 4..7 corner with sz<0 -> +1, sx<0 -> +2),
// bits 3-5 strip array (0: not a strip cell; 1 up, 2 dw, 3 lf, 4 rt), bits 6.. offset inside the array's time slot
struct RingCoef { const float4* coef; const int* meta; };   // [tiles][cells]: (rcp, tv | r1, c2, w)

// One thread per ring cell (tile-major), once per model.
__global__ void ring_coef_kernel(Geo G, RingGeo rg, float4* coef, int* meta)
{
    const int tile = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x, cells = rg.cells;
    int cz0, cx0;
    const RingRect o = ring2_rect(G, rg, tile, &cz0, &cx0);
    const int oh = o.zb - o.za, ow = o.xb - o.xa;
    if (c >= cells) return;
    float4 cf = make_float4(0.f, 0.f, 0.f, 0.f);
    int m = 0;
    if (c < oh * ow) {
        const int oz = c / ow, ox = c - oz * ow, z = o.za + oz, x = o.xa + ox;
        const int NZ = G.NZ, NX = G.NX, N2 = G.N2, nf = G.nfdmax;
        const float* V = G.v + G.padL;
        const float vb = V[(size_t)z * G.pitch + x];
        const int dz = min(z, NZ - 1 - z), dx = min(x, NX - 1 - x), a = min(dz, dx);
        const int sz = (z < NZ - 1 - z) ? 1 : -1, sx = (x < NX - 1 - x) ? 1 : -1;
        if (abs(dz - dx) <= 1) {   // corner (Hybrid1 :138-155): r1 = sqrt((v*tao/h)^2/2) at the cell itself
            const float r  = __fdiv_rn(__fmul_rn(vb, G.tao), G.h);
            const float r2 = __double2float_rn(__dmul_rn(__dmul_rn((double)r, (double)r), 0.5));
            const float r1 = __fsqrt_rn(r2);
            cf.x = __frcp_rn(__fmaf_rn(2.0f, r1, 1.0f));
            cf.y = r1;
            m = 4 + (sz < 0 ? 1 : 0) + (sx < 0 ? 2 : 0);
        } else {
            int fz = a, fx = x;              // top :124 / bottom :132: velocity of row a, this column
            if (dz >= dx) {                  // left :128 / right :136: flat index (N2-l)*NX + row (Q2)
                fx = z;
                while (fx >= NX) { fx -= NX; ++fz; }
            }
            const float vq = V[(size_t)fz * G.pitch + fx];
            const float tv = __fmul_rn(G.taoh, vb);
            cf.x = __frcp_rn(__fadd_rn(tv, 1.0f));
            cf.y = tv;
            cf.z = __fmul_rn(__fmul_rn(G.taoh2, vq), vq);
            m = dz < dx ? (sz > 0 ? 0 : 1) : (sx > 0 ? 2 : 3);
        }
        cf.w = G.w[N2 - a];
        // boundary strips (:23-43, :189-206): the nf cells just outside the interior
        int arr = 0, off = 0;
        if (x >= N2 && x < NX - N2) {
            if (z >= N2 - nf && z < N2) { arr = 1; off = (z - (N2 - nf)) * G.mod_NX + x - N2; }
            else if (z >= NZ - N2 && z < NZ - N2 + nf) { arr = 2; off = (z - (NZ - N2)) * G.mod_NX + x - N2; }
        } else if (z >= N2 && z < NZ - N2) {
            if (x >= N2 - nf && x < N2) { arr = 3; off = (z - N2) * nf + x - (N2 - nf); }
            else if (x >= NX - N2 && x < NX - N2 + nf) { arr = 4; off = (z - N2) * nf + x - (NX - N2); }
        }
        m |= (arr << 3) | (off << 6);
    }
    coef[(size_t)tile * cells + c] = cf;
    meta[(size_t)tile * cells + c] = m;
}

struct RingArgs {
    const float* P1;    // current field (only for nothing: kept for symmetry; the data comes through the tensor maps)
    float*       P2;    // out: forward field / receiver field, slot k
    float*       SX;    // backward: buffer whose ring receives the strips of slot k (BKEqual one step early), or null
    const int2*  src;
    float        wavelet;
    int          inject;   // forward: add the source term
    int          k, nshots;
    Strips       st;       // forward: save (may hold nulls); backward: restore source
    const float* seis;     // backward: [S][NT][n], row k+1 imposed (null in the forward pass)
    float*       gather;   // forward: [S][NT][n] or null
    int          sum_double;   // forward, Taylor operator: Add_Con's double final sum; backward receiver field: float
    RingCoef     rc;
    RingGeo      rg;
};

#ifndef RTM_RING_MINB
#define RTM_RING_MINB 4
#endif

template <int RP, bool BWD, bool LS>
__global__ void __launch_bounds__(kThreads, RTM_RING_MINB)
ring_kernel(const __grid_constant__ RingMaps tm, const __grid_constant__ Geo G, const RingArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int shot = fast_div(blockIdx.x, a.rg.fd_ntiles), tile = blockIdx.x - shot * a.rg.ntiles;
    const bool band = tile < 2 * a.rg.nband;
    int cz0, cx0;                                      // compute rectangle origin (may lie outside the array: never consumed there)
    const RingRect o = ring2_rect(G, a.rg, tile, &cz0, &cx0);
    const int R = a.rg.R, NZ = G.NZ, NX = G.NX;
    const int ch = band ? a.rg.chB : a.rg.chS, CW = band ? a.rg.cwB : a.rg.cwS, SP = band ? a.rg.spB : a.rg.spS;
    const int NG = CW >> 2;
    float* s1  = reinterpret_cast<float*>(smem_raw);   // (ch+2R) x SP: row 0 = z cz0-R, column 0 = x cx0-RP
    float* s0  = s1 + a.rg.n1;                         // ch x CW previous field
    float* sAv = s0 + a.rg.nc;                         // ch x CW velocity factor
    float* s2  = sAv + a.rg.nc;                        // ch x CW unblended two-way result
    uint64_t* bar = reinterpret_cast<uint64_t*>(s2 + a.rg.nc);
    const int tid = threadIdx.x;
    const long long so = (long long)shot * G.shot_stride + G.padL;

    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, ((ch + 2 * R) * SP + 2 * ch * CW) * 4);
        tma_load_3d(s1, band ? &tm.p1b : &tm.p1s, bar, G.padL + cx0 - RP, cz0 - R, shot);
        tma_load_3d(s0, band ? &tm.p0b : &tm.p0s, bar, G.padL + cx0, cz0, shot);
        tma_load_3d(sAv, band ? &tm.avb : &tm.avs, bar, G.padL + cx0, cz0, 0);
    }
    // this thread's output cells (tile-major cell index c = oz*ow + ox): coefficients, meta word and -- backward --
    // the boundary-strip value to restore, all in flight during the TMA copies
    const int oh = o.zb - o.za, ow = o.xb - o.xa, ncell = oh * ow;
    const float4* cfp = a.rc.coef + (size_t)tile * a.rg.cells;
    const int*    mtp = a.rc.meta + (size_t)tile * a.rg.cells;
    // (the per-cell coefficients and meta words are read straight from global memory in the one-way loop, one cell ahead:
    //  they are per-model constants, L2-resident and coalesced; staging them cost 25 KB of shared memory per CTA, i.e. one
    //  resident ring CTA instead of three next to a streaming CTA)
    float4 cfn = make_float4(0.f, 0.f, 0.f, 0.f);
    int    mn  = 0;
    if (tid < ncell) { cfn = __ldg(cfp + tid); mn = __ldg(mtp + tid); }
    const int2 src = a.src[shot];
    const size_t slot = (size_t)shot * G.NT + a.k;
    const size_t sxs = slot * G.nfdmax * G.mod_NX, szs = slot * G.nfdmax * G.mod_NZ;   // this slot inside up/dw and lf/rt
    mbar_wait(bar, 0);
    // mirror about the array edge (:65-68) where the box reaches outside the array (TMA wrote zeros there)
    {
        const int rows = ch + 2 * R;
        const int ztop = cz0 - R, xleft = cx0 - RP;            // array coordinates of box row 0 / column 0
        if (ztop < 0 || ztop + rows > NZ) {                    // rows outside: whole box width
            const int nout = ztop < 0 ? -ztop : ztop + rows - NZ;
            for (int r = tid >> 5; r < nout; r += kWarps) {    // a warp per outside row
                const int zr = ztop < 0 ? r : rows - nout + r; // box row outside the array
                const int z  = ztop + zr, zm = z < 0 ? -z : 2 * NZ - 2 - z;
                // (a short last tile: the box reaches far past the array, rows whose mirror image lies outside the box feed nobody)
                if (zm - ztop >= 0 && zm - ztop < rows)
                    for (int cidx = tid & 31; cidx < SP; cidx += 32) s1[zr * SP + cidx] = s1[(zm - ztop) * SP + cidx];
            }
            __syncthreads();
        }
        if (xleft < 0 || xleft + SP > NX) {                    // columns outside: all rows (mirrored rows included)
            const int nl = xleft < 0 ? -xleft : 0, nr = xleft + SP > NX ? xleft + SP - NX : 0;
            for (int j = tid & 31; j < nl + nr; j += 32) {     // a lane per outside column, the warps share the rows
                const int cidx = j < nl ? j : SP - nr + (j - nl);
                const int x = xleft + cidx, xm = x < 0 ? -x : 2 * NX - 2 - x;
                if (xm - xleft >= 0 && xm - xleft < SP)
                    for (int r = tid >> 5; r < rows; r += kWarps) s1[r * SP + cidx] = s1[r * SP + xm - xleft];
            }
        }
        __syncthreads();
    }

    // two-way update of the compute rectangle, one float4 group per thread and pass (the interior tiles' row stencil)
    {
        LsTable T0{};   // adaptive operator: the packed global tables (as ring_tile<>), addressed by the cell's velocity bin
        T0.ip = G.Index; T0.cp = G.c; T0.bmin = 0; T0.bmax = 0xffff; T0.staged = false;
        const unsigned short* BN = G.bins + G.padL;
        const float* seis_row = BWD && a.seis ? a.seis + ((size_t)shot * G.NT + (a.k + 1)) * G.n : nullptr;
        int sh = 0;
        while ((1 << sh) < NG) ++sh;
        for (int it = tid; it < (ch << sh); it += kThreads) {
            const int lz = it >> sh, g = it & ((1 << sh) - 1);
            if (g >= NG) continue;
            const int z = cz0 + lz, x = cx0 + 4 * g;
            float w1[4], p1[4], p0[4], av[4], val[4];
            uint2 b4 = make_uint2(0u, 0u);
            if (LS) {   // (cells of the compute rectangle outside the array: any valid bin, their values feed nobody)
                const size_t row = (size_t)min(max(z, 0), NZ - 1) * G.pitch;
                int xq[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) xq[q] = min(max(x + q, 0), NX - 1);
                b4.x = (unsigned)__ldg(BN + row + xq[0]) | ((unsigned)__ldg(BN + row + xq[1]) << 16);
                b4.y = (unsigned)__ldg(BN + row + xq[2]) | ((unsigned)__ldg(BN + row + xq[3]) << 16);
            }
            stencil_row<RP, LS, 0>(G, s1 + (lz + R) * SP + 4 * g + RP, G.nfdmax, T0, b4, w1, p1, SP);
            unpack(*reinterpret_cast<const float4*>(s0 + lz * CW + 4 * g), p0);
            unpack(*reinterpret_cast<const float4*>(sAv + lz * CW + 4 * g), av);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                val[q] = a.sum_double ? finish_double(av[q], w1[q], p1[q], p0[q]) : finish_float(av[q], w1[q], p1[q], p0[q]);
            if (seis_row && z == G.s_z) {   // replacement (BKAdd :349-353)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = data_index(G, z, x + q);
                    if (j >= 0) {
                        const float d = seis_row[j];
                        if (d != 0.0f) val[q] = d;
                    }
                }
            }
            if (a.inject && z == src.x && src.y >= x && src.y < x + 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (x + q == src.y) val[q] = __fadd_rn(val[q], a.wavelet);
            }
            *reinterpret_cast<float4*>(s2 + lz * CW + 4 * g) = make_float4(val[0], val[1], val[2], val[3]);
        }
    }
    __syncthreads();

    // one-way solution, blend, stores.  Thread t takes cells t, t + 256, ...: (oz, ox) advance without divisions
    float* P2 = a.P2 + so;
    float* const sb1 = a.st.up ? a.st.up + sxs : nullptr;
    float* const sb2 = a.st.up ? a.st.dw + sxs : nullptr;
    float* const sb3 = a.st.up ? a.st.lf + szs : nullptr;
    float* const sb4 = a.st.up ? a.st.rt + szs : nullptr;
    const int dq = kThreads / ow, dr = kThreads - dq * ow;
    int oz = tid / ow, ox = tid - oz * ow;
    for (int c = tid; c < ncell; c += kThreads, oz += dq, ox += dr) {
        if (ox >= ow) { ox -= ow; ++oz; }
        const float4 cf = cfn;
        const int    m  = mn;
        if (c + kThreads < ncell) { cfn = __ldg(cfp + c + kThreads); mn = __ldg(mtp + c + kThreads); }
        const int lz = o.za + oz - cz0, lx = o.xa + ox - cx0;  // position in the compute rectangle
        const int kind = m & 7;
        const float* q2 = s2 + lz * CW + lx;
        const float* q0 = s0 + lz * CW + lx;
        const float* q1 = s1 + (lz + R) * SP + lx + RP;
        float Pb;
        if (kind >= 4) {        // corner: Pb = rcp * fma(r1, P2a + P2b, P1)
            const int sz = (kind & 1) ? -1 : 1, sx = (kind & 2) ? -1 : 1;
            const float nb = __fadd_rn(q2[sx], q2[sz * CW]);
            Pb = __fmul_rn(cf.x, __fmaf_rn(cf.y, nb, q1[0]));
        } else {                // edge: inner neighbour i = towards the interior, t = along the edge
            const int iz = kind == 0 ? 1 : (kind == 1 ? -1 : 0), ix = kind == 2 ? 1 : (kind == 3 ? -1 : 0);
            const int tz = kind >= 2 ? 1 : 0, tx = kind >= 2 ? 0 : 1;
            const int oi2 = iz * CW + ix, ot2 = tz * CW + tx, oi1 = iz * SP + ix;
            const float p2i = q2[oi2], p0i = q0[oi2], p0b = q0[0];
            const float p1b = q1[0], p1i = q1[oi1];
            const float A1  = __fadd_rn(__fsub_rn(p2i, p0i), p0b);
            float B = __fadd_rn(__fmul_rn(-2.0f, p1b), p0b);
            B       = __fadd_rn(B, p2i);
            B       = __fsub_rn(B, __fmul_rn(2.0f, p1i));
            B       = __fadd_rn(B, p0i);
            float D = __fsub_rn(q2[oi2 + ot2], __fmul_rn(2.0f, p2i));
            D       = __fadd_rn(D, q2[oi2 - ot2]);
            D       = __fadd_rn(D, q0[ot2]);
            D       = __fsub_rn(D, __fmul_rn(2.0f, p0b));
            D       = __fadd_rn(D, q0[-ot2]);
            Pb = __fmul_rn(cf.x, __fmaf_rn(cf.z, D, __fmaf_rn(cf.y, A1, -B)));
        }
        const float val = __fmaf_rn(__fsub_rn(1.0f, cf.w), q2[0], __fmul_rn(cf.w, Pb));   // Hybrid2 :160-183
        const int z = o.za + oz, x = o.xa + ox;
        P2[(size_t)z * G.pitch + x] = val;
        const int arr = (m >> 3) & 7;
        if (arr && sb1) {
            float* sp = (arr == 1 ? sb1 : arr == 2 ? sb2 : arr == 3 ? sb3 : sb4) + (m >> 6);
            if (BWD) { if (a.SX) a.SX[so + (size_t)z * G.pitch + x] = *sp; }   // BKEqual :222-245 (slot k, one step early)
            else *sp = val;                                                     // Hybrid3 :184-208
        }
        if (!BWD && a.gather) {
            const int j = data_index(G, z, x);
            if (j >= 0) a.gather[slot * G.n + j] = val;
        }
    }
}

}  // namespace rtmk
