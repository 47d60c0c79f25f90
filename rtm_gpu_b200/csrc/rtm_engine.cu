// rtm_engine.cu -- context, memory layout, time loops and the C ABI (include/rtm_b200.h).
//
// HBM layout (per context = per GPU), S = max_batch shots advanced together:
//   field buffers  8 x [S][NZ][pitch] f32   pitch = multiple of 32 floats, the interior's first
//                                           column sits on a 128-byte boundary (padL)
//                  forward pass: 3 rotate (slots k-2,k-1,k).  backward pass: four buffers per
//                  field (source / receiver): slots k+2,k+1 are read, slots k,k-1 written by one
//                  PAIR of steps (two-step kernel on the inner tiles, single-step kernel on the
//                  ring and the frame of tiles next to it), then the roles swap.
//   accumulators   4 x [S][NZ][pitch]       sumS, sumR, rel1, rel2 (only the interior is used)
//   velocity       [NZ][pitch]              shared by all shots
//   strips         up/dw [S][NT][nfdmax][mod_NX], lf/rt [S][NT][mod_NZ][nfdmax]   (64-bit sizes)
//   traces         [S][NT][n] time-major (observed data or recorded gather)
//   images         up/down [S][mod_NX][mod_NZ]; stack 2 x [mod_NX][mod_NZ]
// There is no CPU fallback: every entry point fails with RTM_ERR_NO_DEVICE/RTM_ERR_CUDA when
// no B200-class device is usable.
#include "../../include/rtm_b200.h"
#include "host/rtm_host.h"
#include "rtm_kernels.cuh"
#include "rtm_stream.cuh"
#include "rtm_ring.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace rtmk;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
int rtm_fail(int code, const char* fmt, ...);
extern "C" const char* rtm_last_error(void) { return g_err.c_str(); }
extern "C" const char* rtm_version(void) { return "rtm_b200 0.1 (sm_100a)"; }

int rtm_fail(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return rtm_fail(RTM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                            __FILE__, __LINE__);                                              \
    } while (0)

// ------------------------------------------------------------------------------------ aux kernels
namespace {

// a = ((v*v)*tao2)*h2 for every cell, with the kernels' own rounding (Taylor path reads it
// instead of recomputing it every step)
__global__ void avel_kernel(const float* v, float* a, Geo G, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = vel_factor(G, v[i]);
}

// velocity bin of every cell, (int)((v-vmin)/dv+0.5) exactly as the reference's kernels compute it
// (kernel.cu:55-56), stored once per model as a 2-byte side array
__global__ void bins_kernel(const float* v, unsigned short* bins, Geo G, size_t n, int nvel)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % G.pitch) - G.padL;
    int b = 0;
    if (x >= 0 && x < G.NX) {
        const float q = __fdiv_rn(__fsub_rn(v[i], G.vmin), G.dv);
        b = __double2int_rz(__dadd_rn((double)q, 0.5));
        b = min(max(b, 0), nvel - 1);
    }
    bins[i] = (unsigned short)b;
}

// (min bin, max bin) over the interior cells of every interior tile of height tile_rows
// (grow > 0: the tile grown by `grow` cells on every side, clipped to the interior)
__global__ void tile_bins_kernel(const unsigned short* bins, Geo G, int tile_rows, int ntx, int grow, int2* out)
{
    __shared__ int smin, smax;
    if (threadIdx.x == 0) { smin = 0x7fffffff; smax = 0; }
    __syncthreads();
    const int t = blockIdx.x, z0 = G.N2 + (t / ntx) * tile_rows - grow, x0 = G.N2 + (t % ntx) * kTX - grow;
    const int w = kTX + 2 * grow;
    int lo = 0x7fffffff, hi = 0;
    for (int i = threadIdx.x; i < (tile_rows + 2 * grow) * w; i += blockDim.x) {
        const int z = z0 + i / w, x = x0 + i % w;
        if (z >= G.N2 && x >= G.N2 && z < G.NZ - G.N2 && x < G.NX - G.N2) {
            const int b = bins[(size_t)z * G.pitch + G.padL + x];
            lo = min(lo, b); hi = max(hi, b);
        }
    }
    atomicMin(&smin, lo);
    atomicMax(&smax, hi);
    __syncthreads();
    if (threadIdx.x == 0) out[t] = make_int2(min(smin, smax), smax);
}

__global__ void init_source_kernel(float* F1, Geo G, const int2* src, float val)
{
    const int s = blockIdx.x;
    F1[(long long)s * G.shot_stride + G.padL + (size_t)src[s].x * G.pitch + src[s].y] = val;  // :803
}

// Equal (kernel.cu:18-45): strips of an existing field as time slot k; also samples the
// gather for that slot.  One thread per strip cell; grid.y = shot.
__global__ void strips_from_field_kernel(const float* F, Geo G, Strips st, int k, float* gather)
{
    const int shot = blockIdx.y;
    const float* P = F + (long long)shot * G.shot_stride + G.padL;
    const int nf = G.nfdmax, N2 = G.N2;
    const size_t nx = (size_t)nf * G.mod_NX, nz = (size_t)nf * G.mod_NZ;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (st.up && i < nx) {
        const int j = i / G.mod_NX, x = i % G.mod_NX;
        const size_t o = ((size_t)shot * G.NT + k) * nx + i;
        st.up[o] = P[(size_t)(N2 - nf + j) * G.pitch + N2 + x];
        st.dw[o] = P[(size_t)(G.NZ - N2 + j) * G.pitch + N2 + x];
    }
    if (st.up && i < nz) {
        const int z = i / nf, j = i % nf;
        const size_t o = ((size_t)shot * G.NT + k) * nz + i;
        st.lf[o] = P[(size_t)(N2 + z) * G.pitch + N2 - nf + j];
        st.rt[o] = P[(size_t)(N2 + z) * G.pitch + G.NX - N2 + j];
    }
    if (gather && i < (size_t)G.n)
        gather[((size_t)shot * G.NT + k) * G.n + i] = P[(size_t)G.s_z * G.pitch + G.s_l + i * G.ds];
}

// Accumulator start values (kernel.cu:859-876; host arithmetic there: no fused ops).
// BW0 = slot NT-1, BW1 = slot NT-2; FW0/FW1 are the forward INITIAL conditions.
__global__ void acc_init_kernel(Geo G, const float* BW0f, const float* BW1f, const int2* src,
                                float fw1src, float* sumS, float* sumR, float* rel1, float* rel2)
{
    const int shot = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= G.NX) return;
    const size_t o = (long long)shot * G.shot_stride + G.padL + (size_t)z * G.pitch + x;
    const float FW0 = 0.0f, FW1 = (z == src[shot].x && x == src[shot].y) ? fw1src : 0.0f;
    const float BW0 = BW0f[o], BW1 = BW1f[o];
    if (G.iCompen == 1) {
        const float s = __fadd_rn(BW0, BW1), r = __fadd_rn(FW0, FW1);
        sumS[o] = s;
        sumR[o] = r;
        rel1[o] = __fadd_rn(__fmul_rn(r, s), __fmul_rn(FW0, BW0));
    } else {
        sumS[o] = 0.0f;
        sumR[o] = 0.0f;
        rel1[o] = __fadd_rn(__fmul_rn(FW1, BW1), __fmul_rn(FW0, BW0));
    }
    rel2[o] = __fadd_rn(__fmul_rn(BW1, BW1), __fmul_rn(BW0, BW0));
}

// Per-shot image filter (kernel.cu:935-949): up = -Lap5(rel1) * v^2 / vmax^2, Laplacian
// mirrored about the interior edge, evaluated in double as the reference's expression is;
// output x-outer/z-inner.  Also the max |rel2| of the interior (:963-970).
__global__ void image_up_kernel(Geo G, const float* rel1, const float* rel2, float vmax2,
                                float* up, int* maxbits)
{
    const int shot = blockIdx.z;
    const int zi = blockIdx.x * blockDim.x + threadIdx.x, xi = blockIdx.y;  // interior coords
    if (zi >= G.mod_NZ) return;
    const int N2 = G.N2, i = zi + N2, j = xi + N2;
    const float* M1 = rel1 + (long long)shot * G.shot_stride + G.padL;
    int i1 = i - 1, i2 = i + 1, j1 = j - 1, j2 = j + 1;
    if (i1 < N2) i1 = 2 * N2 - i1;
    if (j1 < N2) j1 = 2 * N2 - j1;
    if (i2 >= G.NZ - N2) i2 = 2 * (G.NZ - N2 - 1) - i2;
    if (j2 >= G.NX - N2) j2 = 2 * (G.NX - N2 - 1) - j2;
    float lap = __fadd_rn(M1[(size_t)i * G.pitch + j2], M1[(size_t)i * G.pitch + j1]);
    lap       = __fadd_rn(lap, M1[(size_t)i2 * G.pitch + j]);
    lap       = __fadd_rn(lap, M1[(size_t)i1 * G.pitch + j]);
    lap       = __fsub_rn(lap, __fmul_rn(4.0f, M1[(size_t)i * G.pitch + j]));
    const double vv = (double)G.v[G.padL + (size_t)i * G.pitch + j];
    const double d  = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(-1.0, (double)lap), vv), vv), (double)vmax2);
    up[((size_t)shot * G.mod_NX + xi) * G.mod_NZ + zi] = __double2float_rn(d);
    const float m = fabsf(rel2[(long long)shot * G.shot_stride + G.padL + (size_t)i * G.pitch + j]);
    atomicMax(maxbits + shot, __float_as_int(m));  // non-negative floats order like ints
}

__global__ void image_down_kernel(Geo G, const float* rel2, const int* maxbits, float whitecoe,
                                  float* down, float* stable_out)
{
    const int shot = blockIdx.z;
    const int zi = blockIdx.x * blockDim.x + threadIdx.x, xi = blockIdx.y;
    if (zi >= G.mod_NZ) return;
    const float stable = __fmul_rn(__int_as_float(maxbits[shot]), whitecoe);  // :971
    if (zi == 0 && xi == 0) stable_out[shot] = stable;
    const float r = rel2[(long long)shot * G.shot_stride + G.padL + (size_t)(zi + G.N2) * G.pitch + xi + G.N2];
    down[((size_t)shot * G.mod_NX + xi) * G.mod_NZ + zi] = __fadd_rn(r, stable);
}

// Stack (kernel.cu:1026-1039): float accumulation in shot order.
__global__ void stack_add_kernel(size_t ncell, int nshots, const float* up, const float* down,
                                 float* stack /* up then down */)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    float a = stack[i], b = stack[ncell + i];
    for (int s = 0; s < nshots; ++s) {
        a = __fadd_rn(a, up[(size_t)s * ncell + i]);
        b = __fadd_rn(b, down[(size_t)s * ncell + i]);
    }
    stack[i]         = a;
    stack[ncell + i] = b;
}

// [S][n][NT] (trace-major, the reference's file layout :831-837) <-> [S][NT][n] (time-major)
__global__ void transpose_traces_kernel(const float* in, float* out, int rows, int cols)
{
    __shared__ float t[32][33];
    const size_t base = (size_t)blockIdx.z * rows * cols;
    int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 32 + threadIdx.y;
    for (int i = 0; i < 32; i += 8)
        if (c < cols && r + i < rows) t[threadIdx.y + i][threadIdx.x] = in[base + (size_t)(r + i) * cols + c];
    __syncthreads();
    c = blockIdx.y * 32 + threadIdx.x;
    r = blockIdx.x * 32 + threadIdx.y;
    for (int i = 0; i < 32; i += 8)
        if (c < rows && r + i < cols) out[base + (size_t)(r + i) * rows + c] = t[threadIdx.x][threadIdx.y + i];
}

// resample() (Resample.cpp:193-225) for a whole gather on the device, fused with the transpose
// into the engine's time-major layout: raw [S][n][NT1] at dxin -> out [S][NT][n] at dxout.
// Same table, same float/double evaluation order as csrc/host/resample.cpp (bit-identical).
__global__ void resample_traces_kernel(const float* raw, float* out, int n, int NT1, int NT, float dxout,
                                       float xouts, float xoutb, const float* table)
{
    __shared__ float t[32][33];
    const size_t ibase = (size_t)blockIdx.z * n * NT1, obase = (size_t)blockIdx.z * n * NT;
    const int k = blockIdx.x * 32 + threadIdx.x;
    for (int i = 0; i < 32; i += 8) {
        const int j = blockIdx.y * 32 + threadIdx.y + i;
        float val = 0.0f;
        if (k < NT && j < n) {
            const float* yin  = raw + ibase + (size_t)j * NT1;
            const float xout  = __fmul_rn((float)k, dxout);
            const float xoutn = __fadd_rn(xoutb, __fmul_rn(xout, xouts));
            const int   ix    = (int)xoutn;
            int         ky    = -11 + ix;
            const float frac  = __fsub_rn(xoutn, (float)ix);
            const int   kt    = frac >= 0.0f ? (int)((double)__fmul_rn(frac, 512.0f) + 0.5)
                                             : (int)(((double)frac + 1.0) * 512.0 - 0.5);
            const float* w = table + kt * 8;
#pragma unroll
            for (int q = 0; q < 8; ++q, ++ky) {
                const float y = (ky < 0 || ky >= NT1) ? 0.0f : yin[ky];
                const float p = __fmul_rn(y, w[q]);
                val = (q == 0) ? p : __fadd_rn(val, p);
            }
        }
        t[threadIdx.y + i][threadIdx.x] = val;
    }
    __syncthreads();
    const int j = blockIdx.y * 32 + threadIdx.x;
    for (int i = 0; i < 32; i += 8) {
        const int kk = blockIdx.x * 32 + threadIdx.y + i;
        if (kk < NT && j < n) out[obase + (size_t)kk * n + j] = t[threadIdx.x][threadIdx.y + i];
    }
}

}  // namespace

// ------------------------------------------------------------------------------------ context
struct rtm_ctx {
    int        device = 0;
    rtm_params p{};
    Geo        G{};
    int        S = 1, RP = 4;
    bool       have_model = false, have_op = false;
    float      vmax = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // the tile classes of one time step run concurrently
    cudaEvent_t  fork_ev = nullptr, join_ev[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr;
    cudaEvent_t  ev_ii[2] = {nullptr, nullptr}, ev_ib[2] = {nullptr, nullptr};  // pair stepping, alternating
    // device memory
    static constexpr int kFields = 8;
    float* field[kFields] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* acc[4]   = {nullptr, nullptr, nullptr, nullptr};
    // Interior tiles grouped by the operator length they need (rounded up to 4): every class runs
    // the kernel template, halo box and register budget of its own radius.  The Taylor operator
    // has one class holding all tiles (tile lists null).
    struct TileClass {
        int  RP = 4;
        int  n_f = 0, n_b = 0;              // tiles per shot in the forward / backward tiling
        int *d_tiles_f = nullptr, *d_tiles_b = nullptr;
        // pair stepping: inner tiles advanced two steps per pass (classified by the operator of the
        // tile grown by one radius), frame tiles (next to the ring, or too long an operator) stepped singly
        // (inner tiles: "ii" = all eight neighbours are inner tiles too, "ib" = next to a frame tile)
        int  n_b2 = 0, n_ii = 0, n_ib = 0, n_bf = 0;
        int  ii_rect[3] = {0, 0, 0};        // first tile, width, row stride when the ii tiles form a rectangle
        int *d_tiles_ii = nullptr, *d_tiles_ib = nullptr, *d_tiles_bf = nullptr;
        CUtensorMap tmap_f[kFields], tmap_b[kFields], tmap_b2[kFields], tmap_store;
        size_t smem_f = 0, smem_b = 0, smem_b2 = 0;  // dynamic shared memory already granted to the kernels
        // z-streaming two-step kernel (rtm_stream.cuh): streamed region as column segments, thin frame as strips
        bool stream_mode = false;
        int4 *d_segs_ii = nullptr, *d_segs_ib = nullptr;
        int  n_segs_ii = 0, n_segs_ib = 0;
        long ii_blocks = 0;
        double stream_cells = 0;                                  // cells per shot advanced two slots per pass
        ThinTile* d_thin = nullptr;
        int  n_thin = 0;
        CUtensorMap tmap_s_cur[kFields], tmap_s_prev[kFields];   // boxes (128+4RP) x 8 and (128+2RP) x 8
        CUtensorMap tmap_t_row[kFields], tmap_t_col[kFields];    // thin-frame boxes
        CUtensorMap tmap_s_own[kFields];                         // 128 x 8 (previous field of the single-step streaming forward kernel)
        int4* d_segs_f1 = nullptr;                               // ... and its segments: the whole interior
        int   n_segs_f1 = 0;
        bool smem_s2 = false, smem_s2f = false, smem_t = false;
    };
    // Pair stepping of the backward pass (two-step kernel on the inner tiles).  Measured on the
    // B200 (profiles/README.md) it pays for the Taylor operator up to radius 4 once a launch holds a
    // few waves of inner tiles; for longer operators and the adaptive operator it is opt-in:
    //   RTM_FUSE2=0 off, =1 forced (any operator, any size), unset: automatic;
    //   RTM_FUSE2_MAXRP=4|8 longest (rounded) radius stepped in pairs.
    bool   fuse2 = true;
    bool   fuse2_forced = false;
    int    fuse2_maxrp = 4;
    static constexpr int kFuse2MinCtas = 1500;  // automatic mode: inner-inner tiles x shots per launch
    // L2 look-ahead (profiles/README.md): one thread of every interior CTA prefetches the TMA boxes
    // of the CTA `distance` blocks ahead (cp.async.bulk.prefetch.tensor), so that CTA's copies hit L2
    int    lookahead_f = 148, lookahead_b = 148, lookahead_b2 = 148, lookahead_more = 1, lookahead_p0 = 3;
    // ring CTAs dealt evenly among the interior CTAs of a launch (RTM_RING_INTERLEAVE=0: all ring CTAs first)
    bool   ring_interleave = true;
    int    ring_spread = 8;                 // ... over the first ring_spread/8 of the grid (RTM_RING_SPREAD)
    Acc4Maps tmap_acc;
    // z-streaming form of the two-step kernel (Taylor operator, radius <= 4): RTM_STREAM2=0 keeps the tile form;
    // RTM_SEG_TILES = longest segment in 16-row tiles
    bool   stream2 = true;
    int    seg_tiles = 12;                  // segments of <= 24 blocks (2301 x 751: 4 pieces per column; measured 8 -> 12: backward -0.9 %, profiles/r2_c27/28_*)
    bool   lookahead_f_auto = true;         // forward look-ahead distance by batch size unless RTM_LOOKAHEAD_F is set
    bool   stream1_fwd = false;             // single-step forward pass by stream1_fwd_kernel (RTM_STREAM1_FWD=1).  Measured slower than the
                                            // tile kernel (156 vs 143 us per step: 70 % issue-bound, one row of one field per warp and block
                                            // does not amortise the block overhead), so off by default
    int    fuse2_fwd = 0;                   // forward pass in pairs: RTM_FUSE2_FWD=1 on (measured slower than single steps: the forward
                                            // step already runs at 74 % of the HBM peak and pays the pipeline's extra launches), default off
    CUtensorMap tmap_s_acc[4];              // rel1, rel2, sumS, sumR with a box of 128 x 8
    // the absorbing ring as a kernel of its own (rtm_ring.cuh; Taylor operator): RTM_RING2=0 keeps ring_tile<> everywhere,
    // RTM_RING2_FWD=0 keeps it in the single-step forward launches
    bool   ring2 = true, ring2_fwd = true, ring2_bwd = true, ring_ready = false;
    RingGeo rgeo{};
    cudaStream_t ring_stream = nullptr;     // the ring kernel runs next to the interior launch of a single step
    cudaEvent_t  ring_fork = nullptr, ring_join = nullptr;
    float4* d_ring_coef = nullptr;
    int*    d_ring_meta = nullptr;
    CUtensorMap tmap_r_p1b[kFields], tmap_r_p1s[kFields], tmap_r_p0b[kFields], tmap_r_p0s[kFields], tmap_r_avb, tmap_r_avs;
    bool   dry = false;                     // launch helpers only set kernel attributes
    long   nlaunch = 0;                     // kernels launched (graph replays included)
    std::map<long long, long> graph_launches;
    std::vector<TileClass> classes;
    std::vector<int> h_M;                   // operator length per velocity bin (adaptive operator)
    // store-all mode (RTM_FLAG_STORE_ALL): every forward time slot stays in HBM, [NT][S][NZ][pitch]
    float* store = nullptr;
    bool   store_mode = false;
    float* d_v = nullptr;
    float* d_avel = nullptr;
    unsigned short* d_bins = nullptr;
    int2 *d_tile_bins_f = nullptr, *d_tile_bins_b = nullptr, *d_tile_bins_b2 = nullptr;
    int    nvel = 0;
    float* d_c = nullptr;
    int*   d_Index = nullptr;
    float* d_ls_rows = nullptr;             // adaptive operator no longer than 4: padded coefficient rows [nvel][8] + lengths (streaming kernels)
    int*   d_ls_len = nullptr;
    bool   ring2_bwd_all = false;           // RTM_RING2_BWD=2
    bool   ring_par = true;                 // stream mode: ring launches of the pair loop on the ring stream, next to the ib / thin launches
    bool   fuse2_on = false;                // pairs of steps for this model + operator (prepare_classes)
    bool   ring_frame_only = false;         // adaptive operator: ring_kernel only for the frame steps of the pair loop
    Strips st{nullptr, nullptr, nullptr, nullptr};
    float* d_traces = nullptr;   // [S][NT][n]
    float* d_stage  = nullptr;   // [S][n][NT] transpose staging
    float* d_raw    = nullptr;   // [S][n][NT1] raw-rate traces (rtm_migrate_raw)
    size_t raw_floats = 0;
    float* d_sinc   = nullptr;   // 513 x 8 interpolation table
    int2*  d_src = nullptr;
    float *d_up = nullptr, *d_down = nullptr, *d_stack = nullptr, *d_stable = nullptr;
    int*   d_maxbits = nullptr;
    int    stack_shots = 0;
    size_t field_floats = 0;
    float  last_forward_ms = 0;
    // The launches of a whole time loop depend only on (loop kind, shots in the batch): they are
    // captured once into a CUDA graph and replayed for every later batch (one graph launch per
    // loop instead of NT-2 kernel launches).  Invalidated when the model/operator changes.
    std::map<long long, cudaGraphExec_t> graphs;
    bool   use_graphs = true;
    rtm_stats stats{};
};

static int encode_tmap(rtm_ctx* c, CUtensorMap* m, float* base, int RP, int tile_rows, long long nslab = 0, int box_w = 0)  // RP = 0: box = the tile; box_w > 0: box_w x tile_rows
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess)
            return rtm_fail(RTM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
        fn = (EncodeFn)p;
    }
    const Geo& G = c->G;
    cuuint64_t dims[3]    = {(cuuint64_t)G.pitch, (cuuint64_t)G.NZ, (cuuint64_t)(nslab ? nslab : c->S)};
    cuuint64_t strides[2] = {(cuuint64_t)G.pitch * 4, (cuuint64_t)G.shot_stride * 4};
    cuuint32_t box[3]     = {(cuuint32_t)(kTX + 2 * RP), (cuuint32_t)(tile_rows + 2 * RP), 1};
    if (box_w > 0) { box[0] = (cuuint32_t)box_w; box[1] = (cuuint32_t)tile_rows; }
    cuuint32_t estr[3]    = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return rtm_fail(RTM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return RTM_OK;
}

extern "C" int rtm_ctx_device(rtm_ctx* c) { return c ? c->device : -1; }
extern "C" int rtm_store_all_active(rtm_ctx* c) { return (c && c->store_mode) ? 1 : 0; }

extern "C" int rtm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" void rtm_destroy(rtm_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.second);
    for (auto& k : c->classes) { cudaFree(k.d_tiles_f); cudaFree(k.d_tiles_b); cudaFree(k.d_tiles_ii); cudaFree(k.d_tiles_ib); cudaFree(k.d_tiles_bf); cudaFree(k.d_segs_ii); cudaFree(k.d_segs_ib); cudaFree(k.d_thin); cudaFree(k.d_segs_f1); }
    cudaFree(c->store);
    cudaFree(c->d_ring_coef); cudaFree(c->d_ring_meta);
    for (auto& f : c->field) cudaFree(f);
    for (auto& f : c->acc) cudaFree(f);
    cudaFree(c->d_v); cudaFree(c->d_avel); cudaFree(c->d_bins); cudaFree(c->d_tile_bins_f); cudaFree(c->d_tile_bins_b); cudaFree(c->d_tile_bins_b2); cudaFree(c->d_c); cudaFree(c->d_Index); cudaFree(c->d_ls_rows); cudaFree(c->d_ls_len);
    cudaFree(c->st.up); cudaFree(c->st.dw); cudaFree(c->st.lf); cudaFree(c->st.rt);
    cudaFree(c->d_traces); cudaFree(c->d_stage); cudaFree(c->d_raw); cudaFree(c->d_sinc); cudaFree(c->d_src);
    cudaFree(c->d_up); cudaFree(c->d_down); cudaFree(c->d_stack); cudaFree(c->d_stable);
    cudaFree(c->d_maxbits);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evA) cudaEventDestroy(c->evA);
    if (c->evB) cudaEventDestroy(c->evB);
    for (auto& a : c->aux) if (a) cudaStreamDestroy(a);
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    for (auto& e : c->join_ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_ii) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_ib) if (e) cudaEventDestroy(e);
    if (c->ring_stream) cudaStreamDestroy(c->ring_stream);
    if (c->ring_fork) cudaEventDestroy(c->ring_fork);
    if (c->ring_join) cudaEventDestroy(c->ring_join);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// Row layout of a field buffer: the first interior column sits on a 128-byte boundary (padL), rows
// are a multiple of 32 floats long.
static void field_layout(const rtm_params& p, int* padL, int* pitch)
{
    const int RP = (p.nfdmax + 3) / 4 * 4, NX = p.mod_NX + 2 * p.N2;
    *padL = (32 - p.N2 % 32) % 32;
    if (*padL + p.N2 < RP) *padL += 32;  // the TMA box may start left of the first interior column
    *pitch = (*padL + NX + 4 + 31) / 32 * 32;
}

// Device memory a context needs, by the formulas rtm_create / the first migrate / rtm_migrate_raw allocate
// with: `fixed` (model-sized arrays, stack) + max_batch x `per_shot` (fields, accumulators, boundary strips,
// traces, images).  NT1 = samples per raw trace (rtm_migrate_raw stages them when NT1 != NT; 0: not used).
extern "C" int rtm_memory_estimate(const rtm_params* p, int NT1, size_t* fixed, size_t* per_shot)
{
    if (!p || p->mod_NZ < 1 || p->mod_NX < 1 || p->N2 < 1 || p->nfdmax < 1 || p->NT < 1 || p->n < 1)
        return rtm_fail(RTM_ERR_ARG, "rtm_memory_estimate: bad parameters");
    int padL = 0, pitch = 0;
    field_layout(*p, &padL, &pitch);
    const size_t NZ = (size_t)p->mod_NZ + 2 * p->N2, stride = NZ * (size_t)pitch, ncell = (size_t)p->mod_NX * p->mod_NZ;
    const size_t ntiles = ((size_t)p->mod_NX / kTX + 1) * ((size_t)p->mod_NZ / (kWarps * RTM_NR_B) + 1);
    if (fixed)
        *fixed = 2 * (stride + 64) * 4 + (stride + 64) * 2          // velocity, velocity factor, bins
                 + 2 * ncell * 4                                     // stack
                 + (rtm_ctx::kFields + 4) * 64 * 4 + 6 * ntiles * 8  // slack of the field buffers, tile lists
                 + (256ull << 20);                                   // CUDA context, graphs, allocator granularity
    if (per_shot)
        *per_shot = (rtm_ctx::kFields + 4) * stride * 4              // wavefields + imaging accumulators
                    + 2 * (size_t)p->NT * p->nfdmax * ((size_t)p->mod_NX + p->mod_NZ) * 4   // boundary strips (4 arrays)
                    + 2 * (size_t)p->NT * p->n * 4                   // traces time-major + staging
                    + ((NT1 > 0 && NT1 != p->NT) ? (size_t)NT1 * p->n * 4 : 0)
                    + 2 * ncell * 4 + 64;                            // per-shot images
    return RTM_OK;
}
// Page-locked host buffers for the traces / images of rtm_migrate(_raw): copies from pinned memory run at full PCIe
// rate and asynchronously to the host.
extern "C" int rtm_host_alloc_pinned(void** p, size_t bytes)
{
    if (!p) return rtm_fail(RTM_ERR_ARG, "rtm_host_alloc_pinned: null argument");
    *p = nullptr;
    if (cudaHostAlloc(p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return rtm_fail(RTM_ERR_CUDA, "rtm_host_alloc_pinned: cannot page-lock %zu bytes", bytes);
    }
    return RTM_OK;
}
extern "C" void rtm_host_free_pinned(void* p)
{
    if (p) cudaFreeHost(p);
}
extern "C" int rtm_device_free_bytes(int device, size_t* free_bytes)
{
    if (!free_bytes) return rtm_fail(RTM_ERR_ARG, "rtm_device_free_bytes: null argument");
    size_t total = 0;
    CK(cudaSetDevice(device));
    CK(cudaMemGetInfo(free_bytes, &total));
    return RTM_OK;
}

extern "C" int rtm_create(int device, const rtm_params* p, rtm_ctx** out)
{
    if (!p || !out) return rtm_fail(RTM_ERR_ARG, "rtm_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return rtm_fail(RTM_ERR_NO_DEVICE, "rtm_create: no CUDA device (this engine has no CPU path)");
    if (device < 0 || device >= ndev) return rtm_fail(RTM_ERR_ARG, "rtm_create: device %d of %d", device, ndev);
    if (p->mod_NZ < 4 || p->mod_NX < 4 || p->N2 < 1 || p->N2 > 64 || p->NT < 3)
        return rtm_fail(RTM_ERR_ARG, "rtm_create: bad grid (mod_NZ=%d mod_NX=%d N2=%d NT=%d)", p->mod_NZ, p->mod_NX, p->N2, p->NT);
    if (p->nfdmax < 1 || p->nfdmax > kMaxR || p->nfdmax > p->N2)
        return rtm_fail(RTM_ERR_ARG, "rtm_create: need 1 <= nfdmax (%d) <= min(N2=%d, %d): the boundary strips lie inside the ring (kernel.cu:23-43)", p->nfdmax, p->N2, kMaxR);
    if (p->n < 1 || p->ds < 1 || p->s_z < 0 || p->s_z >= p->mod_NZ + 2 * p->N2 || p->s_l < 0 ||
        p->s_l + (p->n - 1) * p->ds >= p->mod_NX + 2 * p->N2)
        return rtm_fail(RTM_ERR_ARG, "rtm_create: data positions outside the padded grid");
    if (p->mod_NZ <= 2 * p->N2 + 4 || p->mod_NX <= 2 * p->N2 + 4)
        return rtm_fail(RTM_ERR_ARG, "rtm_create: model %d x %d must be larger than 2*N2+4 = %d in both directions (the ring tiles assume an interior between the two bands)",
                        p->mod_NX, p->mod_NZ, 2 * p->N2 + 4);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return rtm_fail(RTM_ERR_NO_DEVICE, "rtm_create: device %d is sm_%d%d; this engine is built for sm_100a only", device, prop.major, prop.minor);
    {   // a ring tile stages its halo, previous field and one-way inputs in shared memory: wide rings with long operators do not fit
        const int RPmax = (p->nfdmax + 3) / 4 * 4;
        const size_t ring_bytes = (size_t)ring_smem_floats(p->N2, p->nfdmax, RPmax) * 4;
        if (ring_bytes > (size_t)prop.sharedMemPerBlockOptin)
            return rtm_fail(RTM_ERR_ARG, "rtm_create: absorbing ring N2=%d with operator length %d needs %zu bytes of shared memory per ring tile, "
                            "the device grants %zu: use a narrower ring or a shorter operator", p->N2, p->nfdmax, ring_bytes, (size_t)prop.sharedMemPerBlockOptin);
    }

    rtm_ctx* c = new rtm_ctx;
    c->device = device;
    c->p      = *p;
    c->S      = std::max(1, p->max_batch);
    c->RP     = (p->nfdmax + 3) / 4 * 4;
    Geo& G    = c->G;
    G.mod_NZ = p->mod_NZ; G.mod_NX = p->mod_NX; G.N2 = p->N2;
    G.NZ = p->mod_NZ + 2 * p->N2; G.NX = p->mod_NX + 2 * p->N2;
    field_layout(*p, &G.padL, &G.pitch);
    G.shot_stride = (long long)G.NZ * G.pitch;
    G.nfdmax = p->nfdmax; G.mmax = p->nfdmax; G.NT = p->NT; G.iLSTE = p->iLSTE; G.iCompen = p->iCompen;
    G.tao = p->tao; G.h = p->h;
    // derived scalars exactly as kernel.cu:614-626
    G.taoh  = p->tao / p->h;
    G.tao2  = (float)((double)p->tao * (double)p->tao);
    G.h2    = (float)(1 / ((double)p->h * (double)p->h));
    G.taoh2 = G.tao2 * G.h2 / 2;
    const float hzx = p->hz / p->h;
    G.hzx2_1 = 1 / (hzx * hzx);
    G.A      = 1.0 + (double)G.hzx2_1;
    G.s_l = p->s_l; G.s_z = p->s_z; G.n = p->n; G.ds = p->ds; G.s_r = (p->n - 1) * p->ds + p->s_l;
    G.ntx = (G.mod_NX + kTX - 1) / kTX;
    G.ntz_f = (G.mod_NZ + kWarps * RTM_NR_F - 1) / (kWarps * RTM_NR_F);
    G.ntz_b = (G.mod_NZ + kWarps * RTM_NR_B - 1) / (kWarps * RTM_NR_B);
    G.nband = (G.NX + kRingTX - 1) / kRingTX; G.nside = (G.mod_NZ + kRingTX - 1) / kRingTX;
    G.fd_ntx = make_fastdiv(G.ntx); G.fd_nring = make_fastdiv(2 * G.nband + 2 * G.nside);
    if (const char* e = std::getenv("RTM_NO_GRAPH")) c->use_graphs = std::atoi(e) == 0;
    if (const char* e = std::getenv("RTM_FUSE2")) { c->fuse2 = std::atoi(e) != 0; c->fuse2_forced = c->fuse2; }
    // (adaptive operator: pairs only in the streaming form, i.e. when no bin's operator is longer than 4 -- prepare_classes)
    if (const char* e = std::getenv("RTM_RING_PAR")) c->ring_par = std::atoi(e) != 0;
    if (const char* e = std::getenv("RTM_FUSE2_MAXRP")) c->fuse2_maxrp = std::atoi(e);
    if (const char* e = std::getenv("RTM_STREAM2")) c->stream2 = std::atoi(e) != 0;
    if (const char* e = std::getenv("RTM_RING2")) c->ring2 = std::atoi(e) != 0;
    if (const char* e = std::getenv("RTM_RING2_FWD")) c->ring2_fwd = std::atoi(e) != 0;
    if (const char* e = std::getenv("RTM_RING2_BWD")) { c->ring2_bwd = std::atoi(e) != 0; c->ring2_bwd_all = std::atoi(e) == 2; }
    if (const char* e = std::getenv("RTM_FUSE2_FWD")) c->fuse2_fwd = std::atoi(e) != 0 ? 1 : 0;
    if (const char* e = std::getenv("RTM_STREAM1_FWD")) c->stream1_fwd = std::atoi(e) != 0;
    if (const char* e = std::getenv("RTM_SEG_TILES")) c->seg_tiles = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RTM_RING_INTERLEAVE")) c->ring_interleave = std::atoi(e) != 0;
    if (const char* e = std::getenv("RTM_RING_SPREAD")) c->ring_spread = std::atoi(e);
    if (const char* e = std::getenv("RTM_LOOKAHEAD_F")) { c->lookahead_f = std::atoi(e); c->lookahead_f_auto = false; }
    if (const char* e = std::getenv("RTM_LOOKAHEAD_B2")) c->lookahead_b2 = std::atoi(e);
    if (const char* e = std::getenv("RTM_LOOKAHEAD_MORE")) c->lookahead_more = std::atoi(e);
    if (const char* e = std::getenv("RTM_LOOKAHEAD_B")) c->lookahead_b = std::atoi(e);
    if (const char* e = std::getenv("RTM_LOOKAHEAD_P0")) c->lookahead_p0 = std::atoi(e);
    if (p->flags & RTM_FLAG_STORE_ALL) c->fuse2 = false;
    for (int i = 0; i <= p->N2; ++i) G.w[i] = (float)((1.0 * i) / (1.0 * p->N2));  // :688-691

    auto fail = [&](int rc) { rtm_destroy(c); return rc; };
#define CKC(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(rtm_fail(RTM_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                 \
                                 cudaGetErrorString(e_), __FILE__, __LINE__));                 \
    } while (0)
    CKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {   // aux[2] carries the ring/frame chain of the pair stepping: small launches, served first
        int lo = 0, hi = 0;
        CKC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int i = 0; i < 3; ++i) CKC(cudaStreamCreateWithPriority(&c->aux[i], cudaStreamNonBlocking, i == 2 ? hi : lo));
    }
    {
        int lo = 0, hi = 0;
        CKC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const char* e = std::getenv("RTM_RING_PRIO");
        CKC(cudaStreamCreateWithPriority(&c->ring_stream, cudaStreamNonBlocking, (e && std::atoi(e) != 0) ? hi : lo));
    }
    CKC(cudaEventCreateWithFlags(&c->ring_fork, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&c->ring_join, cudaEventDisableTiming));
    for (auto& e : c->ev_ii) CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : c->ev_ib) CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
    for (auto& e : c->join_ev) CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CKC(cudaEventCreate(&c->ev0));
    CKC(cudaEventCreate(&c->ev1));
    CKC(cudaEventCreate(&c->evA));
    CKC(cudaEventCreate(&c->evB));
    c->field_floats = (size_t)c->S * G.shot_stride + 64;  // slack for float4 loads past the last row
    const size_t ncell = (size_t)G.mod_NX * G.mod_NZ;
    size_t need = (rtm_ctx::kFields + 4) * c->field_floats * 4 + (size_t)G.shot_stride * 4 +
                  2 * (size_t)c->S * G.NT * G.n * 4 + 2 * c->S * ncell * 4;
    size_t free_b = 0, total_b = 0;
    CKC(cudaMemGetInfo(&free_b, &total_b));
    if (need > free_b)
        return fail(rtm_fail(RTM_ERR_ARG, "rtm_create: batch of %d shots needs %.1f GB before strips, %.1f GB free", c->S, need / 1e9, free_b / 1e9));
    for (auto& f : c->field) { CKC(cudaMalloc(&f, c->field_floats * 4)); CKC(cudaMemset(f, 0, c->field_floats * 4)); }
    for (auto& f : c->acc) { CKC(cudaMalloc(&f, c->field_floats * 4)); CKC(cudaMemset(f, 0, c->field_floats * 4)); }
    CKC(cudaMalloc(&c->d_v, ((size_t)G.shot_stride + 64) * 4));
    CKC(cudaMemset(c->d_v, 0, ((size_t)G.shot_stride + 64) * 4));
    CKC(cudaMalloc(&c->d_avel, ((size_t)G.shot_stride + 64) * 4));
    CKC(cudaMemset(c->d_avel, 0, ((size_t)G.shot_stride + 64) * 4));
    CKC(cudaMalloc(&c->d_traces, (size_t)c->S * G.NT * G.n * 4));
    CKC(cudaMalloc(&c->d_stage, (size_t)c->S * G.NT * G.n * 4));
    CKC(cudaMalloc(&c->d_src, sizeof(int2) * c->S));
    CKC(cudaMalloc(&c->d_up, c->S * ncell * 4));
    CKC(cudaMalloc(&c->d_down, c->S * ncell * 4));
    CKC(cudaMalloc(&c->d_stack, 2 * ncell * 4));
    CKC(cudaMemset(c->d_stack, 0, 2 * ncell * 4));
    CKC(cudaMalloc(&c->d_stable, sizeof(float) * c->S));
    CKC(cudaMalloc(&c->d_maxbits, sizeof(int) * c->S));
    G.v = c->d_v;
    G.avel = c->d_avel;
    CKC(cudaMalloc(&c->d_bins, ((size_t)G.shot_stride + 64) * sizeof(unsigned short)));
    CKC(cudaMemset(c->d_bins, 0, ((size_t)G.shot_stride + 64) * sizeof(unsigned short)));
    CKC(cudaMalloc(&c->d_tile_bins_f, sizeof(int2) * G.ntx * G.ntz_f));
    CKC(cudaMalloc(&c->d_tile_bins_b, sizeof(int2) * G.ntx * G.ntz_b));
    CKC(cudaMalloc(&c->d_tile_bins_b2, sizeof(int2) * G.ntx * G.ntz_b));
    G.bins = c->d_bins; G.tile_bins_f = c->d_tile_bins_f; G.tile_bins_b = c->d_tile_bins_b; G.tile_bins_b2 = c->d_tile_bins_b2;
    if (p->flags & RTM_FLAG_STORE_ALL) {
        // keep the whole forward wavefield when it fits (with 4 GB of head-room for the strips-free
        // rest); otherwise fall back to boundary saving + reverse-time reconstruction
        const size_t slot_floats = (size_t)c->S * G.shot_stride;
        const size_t bytes = ((size_t)G.NT * slot_floats + 64) * 4;
        CKC(cudaMemGetInfo(&free_b, &total_b));
        if (bytes + (4ull << 30) < free_b && (size_t)G.NT * c->S < (1ull << 31)) {
            CKC(cudaMalloc(&c->store, bytes));
            CKC(cudaMemset(c->store, 0, bytes));
            c->store_mode = true;
        }
    }
    // The allocations above were cleared with legacy-stream memsets, which are asynchronous to
    // the host and NOT ordered against the context's non-blocking stream: drain them here.
    CKC(cudaDeviceSynchronize());
#undef CKC
    *out = c;
    return RTM_OK;
}

static void drop_graphs(rtm_ctx* c);

// Side arrays of the adaptive operator (need both the model and the operator table).
static int prepare_classes(rtm_ctx* c);

static int prepare_ls(rtm_ctx* c)
{
    if (c->G.iLSTE != 0) return c->have_op ? prepare_classes(c) : RTM_OK;
    if (!c->have_model || !c->have_op) return RTM_OK;
    const Geo& G = c->G;
    if (c->nvel > 65535)
        return rtm_fail(RTM_ERR_ARG, "adaptive operator with %d velocity bins: the per-cell bin array is 16-bit, use a larger dv", c->nvel);
    const size_t n = (size_t)G.shot_stride;
    bins_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_v, c->d_bins, G, n, c->nvel);
    tile_bins_kernel<<<G.ntx * G.ntz_f, 256, 0, c->stream>>>(c->d_bins, G, kWarps * RTM_NR_F, G.ntx, 0, c->d_tile_bins_f);
    tile_bins_kernel<<<G.ntx * G.ntz_b, 256, 0, c->stream>>>(c->d_bins, G, kWarps * RTM_NR_B, G.ntx, 0, c->d_tile_bins_b);
    tile_bins_kernel<<<G.ntx * G.ntz_b, 256, 0, c->stream>>>(c->d_bins, G, kWarps * RTM_NR_B, G.ntx, c->RP, c->d_tile_bins_b2);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    return prepare_classes(c);
}

// Group the interior tiles by the radius class they need and encode the TMA boxes per class.
static int prepare_classes(rtm_ctx* c)
{
    const Geo& G = c->G;
    for (auto& k : c->classes) { cudaFree(k.d_tiles_f); cudaFree(k.d_tiles_b); cudaFree(k.d_tiles_ii); cudaFree(k.d_tiles_ib); cudaFree(k.d_tiles_bf); cudaFree(k.d_segs_ii); cudaFree(k.d_segs_ib); cudaFree(k.d_thin); cudaFree(k.d_segs_f1); }
    c->classes.clear();
    const int nf = G.ntx * G.ntz_f, nb = G.ntx * G.ntz_b;
    struct Lists { std::vector<int> fwd, bwd, ii, ib, frame; };
    std::map<int, Lists> lists;  // RP -> tile lists
    const bool ls = G.iLSTE == 0;
    // pairs of steps: the fixed-length operator; the adaptive operator only in the streaming form (its tile form measured
    // slower than single steps, profiles/README.md), i.e. when every bin's operator fits radius 4 -- or when forced
    const bool ls_stream = ls && c->stream2 && c->RP == 4 && G.ls_rows != nullptr;
    const bool fuse2 = c->fuse2 && (!ls || c->fuse2_forced || ls_stream);
    // an inner tile: full, and grown by one (rounded) radius it still lies in the interior
    const int TZb = kWarps * RTM_NR_B;
    auto inner = [&](int t) {
        const int z0 = G.N2 + (t / G.ntx) * TZb, x0 = G.N2 + (t % G.ntx) * kTX;
        return z0 - c->RP >= G.N2 && x0 - c->RP >= G.N2 && z0 + TZb + c->RP <= G.NZ - G.N2 &&
               x0 + kTX + c->RP <= G.NX - G.N2;
    };
    std::vector<int2> tb(std::max(nf, nb)), tb2(nb);
    if (ls) CK(cudaMemcpy(tb2.data(), c->d_tile_bins_b2, sizeof(int2) * nb, cudaMemcpyDeviceToHost));
    auto radius_class = [&](int2 r) {
        int m = 1;
        for (int b = std::max(r.x, 0); b <= r.y && b < (int)c->h_M.size(); ++b) m = std::max(m, c->h_M[b]);
        return (m + 3) / 4 * 4;
    };
    // Tiles advanced by the two-step kernel: vertical pairs (32 rows) of inner tiles whose grown
    // operator is short enough.  covered[t]: tile t belongs to such a pair; head[t]: it is the upper one.
    std::vector<int> rp2(nb, 0);
    std::vector<char> ok(nb, 0), covered(nb, 0), head(nb, 0);
    for (int t = 0; t < nb; ++t) {
        rp2[t] = ls ? radius_class(tb2[t]) : c->RP;
        ok[t]  = fuse2 && inner(t) && rp2[t] <= c->fuse2_maxrp && rp2[t] <= 8;
    }
    constexpr int gsz = Tile2<4>::TZ / (kWarps * RTM_NR_B);  // single-step tiles stacked in one two-step tile
    for (int tx = 0; tx < G.ntx; ++tx)
        for (int tz = 0; tz + gsz <= G.ntz_b;) {
            const int t = tz * G.ntx + tx;
            bool all = true;
            for (int i = 0; i < gsz; ++i) all = all && ok[t + i * G.ntx];
            if (!all) { ++tz; continue; }
            head[t] = 1;
            for (int i = 0; i < gsz; ++i) {
                const int u = t + i * G.ntx;
                covered[u] = 1;
                rp2[t] = std::max(rp2[t], rp2[u]);
                if (ls) tb2[t] = make_int2(std::min(tb2[t].x, tb2[u].x), std::max(tb2[t].y, tb2[u].y));
            }
            tz += gsz;
        }
    if (ls) CK(cudaMemcpy(c->d_tile_bins_b2, tb2.data(), sizeof(int2) * nb, cudaMemcpyHostToDevice));
    for (int pass = 0; pass < 2; ++pass) {
        const int n = pass ? nb : nf;
        if (ls) CK(cudaMemcpy(tb.data(), pass ? c->d_tile_bins_b : c->d_tile_bins_f, sizeof(int2) * n, cudaMemcpyDeviceToHost));
        for (int t = 0; t < n; ++t) {
            const int rp = ls ? radius_class(tb[t]) : c->RP;
            if (!pass) { lists[rp].fwd.push_back(t); continue; }
            lists[rp].bwd.push_back(t);
            if (!covered[t]) { lists[rp].frame.push_back(t); continue; }
            if (!head[t]) continue;
            bool all = true;  // every tile around the pair is advanced by the two-step kernel as well
            for (int dz = -1; dz <= gsz; ++dz)
                for (int dx = -1; dx <= 1; ++dx) all = all && covered[t + dz * G.ntx + dx];
            (all ? lists[rp2[t]].ii : lists[rp2[t]].ib).push_back(t);
        }
    }
    auto upload = [&](const std::vector<int>& v, int** d) -> int {
        if (v.empty()) return RTM_OK;
        CK(cudaMalloc(d, sizeof(int) * v.size()));
        CK(cudaMemcpy(*d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice));
        return RTM_OK;
    };
    for (auto it = lists.rbegin(); it != lists.rend(); ++it) {  // longest operators first
        rtm_ctx::TileClass k;
        k.RP = it->first;
        const Lists& L = it->second;
        k.n_f = (int)L.fwd.size(); k.n_b = (int)L.bwd.size(); k.n_bf = (int)L.frame.size();
        k.n_ii = (int)L.ii.size(); k.n_ib = (int)L.ib.size(); k.n_b2 = k.n_ii + k.n_ib;
        // (a class that holds every tile needs no list: the kernel then uses the tile number itself)
        if (k.n_f < nf) if (int rc = upload(L.fwd, &k.d_tiles_f)) return rc;
        if (k.n_b < nb) if (int rc = upload(L.bwd, &k.d_tiles_b)) return rc;
        if (int rc = upload(L.ii, &k.d_tiles_ii)) return rc;
        if (k.n_ii > 1) {  // a rectangle?  (rows of equal width, equal row stride)
            const std::vector<int>& v = L.ii;
            int w = 1;
            while (w < k.n_ii && v[w] == v[0] + w) ++w;
            const int dz = w < k.n_ii ? v[w] - v[0] : G.ntx;
            bool rect = k.n_ii % w == 0;
            for (int i = 0; rect && i < k.n_ii; ++i) rect = v[i] == v[0] + (i / w) * dz + i % w;
            if (rect) { k.ii_rect[0] = v[0]; k.ii_rect[1] = w; k.ii_rect[2] = dz; }
        }
        if (int rc = upload(L.ib, &k.d_tiles_ib)) return rc;
        if (int rc = upload(L.frame, &k.d_tiles_bf)) return rc;
        if (c->stream2 && fuse2 && (!ls || ls_stream) && k.RP == 4 && lists.size() == 1) {
            // z-streaming form (rtm_stream.cuh).  Regions, for the forward and the backward pass alike:
            //   ring        the N2 outermost cells                    ring tiles of the single-step kernels
            //   thin frame  the RP interior cells next to the ring    thin_frame_kernel, stepped singly
            //   streamed    the rest [C0, xe) x [R0, ze)              stream2_kernel, two slots per pass: 128-wide columns
            //               (the last may be partial) of 8-row blocks (the last may be partial), cut into segments;
            //               "ib" = the columns / blocks next to the thin frame (the inner-inner segments' halo
            //               never reaches a cell stepped singly), "ii" = everything inside them
            using T = Strm<4>;
            const int seg_blocks = std::max(2, c->seg_tiles * TZb / T::BR);
            const StreamRegions sr = make_stream_regions(G, k.RP, seg_blocks);   // rtm_stream.cuh (tests/test_launch_geometry.py)
            if (sr.ok) {
                auto up4 = [&](const std::vector<int4>& v, int4** d, int* n) -> int {
                    *n = (int)v.size();
                    if (v.empty()) return RTM_OK;
                    CK(cudaMalloc(d, sizeof(int4) * v.size()));
                    CK(cudaMemcpy(*d, v.data(), sizeof(int4) * v.size(), cudaMemcpyHostToDevice));
                    return RTM_OK;
                };
                if (int rc = up4(sr.ii, &k.d_segs_ii, &k.n_segs_ii)) return rc;
                if (int rc = up4(sr.ib, &k.d_segs_ib, &k.n_segs_ib)) return rc;
                k.ii_blocks = 0;
                for (auto& sgm : sr.ii) k.ii_blocks += sgm.z;
                k.stream_cells = sr.stream_cells;
                k.n_thin = (int)sr.thin.size();
                CK(cudaMalloc(&k.d_thin, sizeof(ThinTile) * sr.thin.size()));
                CK(cudaMemcpy(k.d_thin, sr.thin.data(), sizeof(ThinTile) * sr.thin.size(), cudaMemcpyHostToDevice));
                for (int i = 0; i < rtm_ctx::kFields; ++i) {
                    int rc = encode_tmap(c, &k.tmap_s_cur[i], c->field[i], 0, T::BR, 0, T::W1);
                    if (!rc) rc = encode_tmap(c, &k.tmap_s_prev[i], c->field[i], 0, T::BR, 0, T::WM);
                    if (!rc) rc = encode_tmap(c, &k.tmap_t_row[i], c->field[i], 0, Thin<4>::ROW_H, 0, Thin<4>::ROW_W);
                    if (!rc) rc = encode_tmap(c, &k.tmap_t_col[i], c->field[i], 0, Thin<4>::COL_H, 0, Thin<4>::COL_W);
                    if (rc) return rc;
                }
                {   // single-step streaming forward kernel: every interior column, blocks of 8 rows from the first interior row
                    const int ncol1 = (G.mod_NX + kTX - 1) / kTX, nblk1 = (G.mod_NZ + T::BR - 1) / T::BR;
                    std::vector<int4> f1;
                    for (int col = 0; col < ncol1; ++col) {
                        const int pieces = (nblk1 + seg_blocks - 1) / seg_blocks;
                        for (int p = 0, b = 0; p < pieces; ++p) {
                            const int len = nblk1 / pieces + (p < nblk1 % pieces ? 1 : 0);
                            f1.push_back(make_int4(G.N2 + col * kTX, G.N2 + b * T::BR, len, 0));
                            b += len;
                        }
                    }
                    if (int rc = up4(f1, &k.d_segs_f1, &k.n_segs_f1)) return rc;
                    for (int i = 0; i < rtm_ctx::kFields; ++i)
                        if (int rc = encode_tmap(c, &k.tmap_s_own[i], c->field[i], 0, T::BR, 0, kTX)) return rc;
                }
                k.stream_mode = true;
            }
        }
        for (int i = 0; i < rtm_ctx::kFields; ++i) {
            int rc = encode_tmap(c, &k.tmap_f[i], c->field[i], k.RP, kWarps * RTM_NR_F);
            if (!rc) rc = encode_tmap(c, &k.tmap_b[i], c->field[i], k.RP, kWarps * RTM_NR_B);
            if (!rc && k.n_b2) rc = encode_tmap(c, &k.tmap_b2[i], c->field[i], 2 * k.RP, Tile2<4>::TZ);
            if (rc) return rc;
        }
        if (c->store_mode)
            if (int rc = encode_tmap(c, &k.tmap_store, c->store, k.RP, kWarps * RTM_NR_F, (long long)G.NT * c->S)) return rc;
        // accumulators: rel1, rel2, sumS, sumR (acc[] holds sumS, sumR, rel1, rel2)
        const int order[4] = {2, 3, 0, 1};
        for (int i = 0; i < 4; ++i) {
            if (int rc = encode_tmap(c, &c->tmap_acc.m[i], c->acc[order[i]], 0, Tile2<4>::TZ)) return rc;
            if (int rc = encode_tmap(c, &c->tmap_s_acc[i], c->acc[order[i]], 0, Strm<4>::BR, 0, kTX)) return rc;
        }
        c->classes.push_back(k);
    }
    // ring kernel (rtm_ring.cuh): per-cell one-way coefficients of this model, tensor maps of the tile boxes
    c->ring_ready = false;
    const bool streaming = c->classes.size() == 1 && c->classes[0].stream_mode;
    c->fuse2_on = fuse2 && (!ls || c->fuse2_forced || streaming);
    c->ring_frame_only = false;
    // (measured with the adaptive operator, profiles/r2_c11_*: the ring on a side stream gains nothing in the forward pass and
    //  loses 3-12 % in the backward pass, where its two-way phase gathers per-cell coefficients from the global tables next to
    //  interior tiles that do the same; RTM_RING2_LS=1 enables it there anyway)
    const bool ring_ls = std::getenv("RTM_RING2_LS") && std::atoi(std::getenv("RTM_RING2_LS")) != 0;
    if (c->ring2 && (!ls || ring_ls || (streaming && c->fuse2_on)) && c->have_model && c->have_op) {
        c->ring_frame_only = ls && !ring_ls;
        c->rgeo = make_ring_geo(G, G.mmax, c->RP);
        const int nring = c->rgeo.ntiles;
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, c->device));
        if ((size_t)c->rgeo.smem_bytes() <= (size_t)prop.sharedMemPerBlockOptin && c->rgeo.chS + 2 * c->rgeo.R <= 256 && c->rgeo.spB <= 256) {
            cudaFree(c->d_ring_coef); cudaFree(c->d_ring_meta);
            c->d_ring_coef = nullptr; c->d_ring_meta = nullptr;
            CK(cudaMalloc(&c->d_ring_coef, sizeof(float4) * (size_t)nring * c->rgeo.cells));
            CK(cudaMalloc(&c->d_ring_meta, sizeof(int) * (size_t)nring * c->rgeo.cells));
            dim3 grid((c->rgeo.cells + 255) / 256, nring);
            ring_coef_kernel<<<grid, 256, 0, c->stream>>>(G, c->rgeo, c->d_ring_coef, c->d_ring_meta);
            CK(cudaGetLastError());
            const RingGeo& r = c->rgeo;
            for (int i = 0; i < rtm_ctx::kFields; ++i) {
                int rc = encode_tmap(c, &c->tmap_r_p1b[i], c->field[i], 0, r.chB + 2 * r.R, 0, r.spB);
                if (!rc) rc = encode_tmap(c, &c->tmap_r_p1s[i], c->field[i], 0, r.chS + 2 * r.R, 0, r.spS);
                if (!rc) rc = encode_tmap(c, &c->tmap_r_p0b[i], c->field[i], 0, r.chB, 0, r.cwB);
                if (!rc) rc = encode_tmap(c, &c->tmap_r_p0s[i], c->field[i], 0, r.chS, 0, r.cwS);
                if (rc) return rc;
            }
            if (int rc = encode_tmap(c, &c->tmap_r_avb, c->d_avel, 0, r.chB, 1, r.cwB)) return rc;
            if (int rc = encode_tmap(c, &c->tmap_r_avs, c->d_avel, 0, r.chS, 1, r.cwS)) return rc;
            c->ring_ready = true;
        }
    }
    CK(cudaDeviceSynchronize());
    return RTM_OK;
}

extern "C" int rtm_set_model(rtm_ctx* c, const float* v, float vmin, float vmax, float dv)
{
    if (!c || !v) return rtm_fail(RTM_ERR_ARG, "rtm_set_model: null argument");
    if (!(dv > 0)) return rtm_fail(RTM_ERR_ARG, "rtm_set_model: dv must be positive");
    CK(cudaSetDevice(c->device));
    Geo& G = c->G;
    // everything on the context's own (non-blocking) stream: a plain cudaMemcpy from pageable
    // memory may return before the DMA lands and is not ordered against that stream
    CK(cudaMemcpy2DAsync(c->d_v + G.padL, (size_t)G.pitch * 4, v, (size_t)G.NX * 4, (size_t)G.NX * 4, G.NZ,
                         cudaMemcpyHostToDevice, c->stream));
    G.vmin = vmin; G.dv = dv; c->vmax = vmax;
    {
        const size_t n = (size_t)G.shot_stride;
        avel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_v, c->d_avel, G, n);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c->stream));
    }
    c->have_model = true;
    drop_graphs(c);
    return prepare_ls(c);
}

extern "C" int rtm_set_operator(rtm_ctx* c, const int* Index, int nvel, const float* coef, int NC)
{
    if (!c || !coef) return rtm_fail(RTM_ERR_ARG, "rtm_set_operator: null argument");
    CK(cudaSetDevice(c->device));
    Geo& G = c->G;
    if (G.iLSTE == 0) {
        if (!Index || nvel < 1 || NC < 1) return rtm_fail(RTM_ERR_ARG, "rtm_set_operator: adaptive operator needs Index[nvel+1] and c[NC]");
        if (Index[nvel] != NC) return rtm_fail(RTM_ERR_ARG, "rtm_set_operator: Index[nvel]=%d != NC=%d", Index[nvel], NC);
        for (int i = 0; i < nvel; ++i)
            if (Index[i + 1] - Index[i] - 1 > G.nfdmax)
                return rtm_fail(RTM_ERR_ARG, "rtm_set_operator: bin %d has length %d > nfdmax %d", i, Index[i + 1] - Index[i] - 1, G.nfdmax);
        cudaFree(c->d_c); cudaFree(c->d_Index);
        c->d_c = nullptr; c->d_Index = nullptr;
        // one spare entry: the lookup reads Index[bin+1] and c[Index[bin]] for any cell
        CK(cudaMalloc(&c->d_c, sizeof(float) * (NC + 1)));
        CK(cudaMalloc(&c->d_Index, sizeof(int) * (nvel + 2)));
        CK(cudaMemsetAsync(c->d_c, 0, sizeof(float) * (NC + 1), c->stream));
        CK(cudaMemcpyAsync(c->d_c, coef, sizeof(float) * NC, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_Index, Index, sizeof(int) * (nvel + 1), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_Index + nvel + 1, Index + nvel, sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        G.c = c->d_c; G.Index = c->d_Index;
        cudaFree(c->d_ls_rows); cudaFree(c->d_ls_len);
        c->d_ls_rows = nullptr; c->d_ls_len = nullptr;
        G.ls_rows = nullptr; G.ls_len = nullptr; G.ls_nbins = 0;
        int longest = 1;
        for (int i = 0; i < nvel; ++i) longest = std::max(longest, Index[i + 1] - Index[i] - 1);
        if (longest <= 4) {   // rows of the streaming kernels: c[Index[b] .. Index[b]+M] zero-padded to 8 floats, M per bin
            std::vector<float> rows((size_t)nvel * 8, 0.0f);
            std::vector<int> len(nvel, -1);
            for (int b = 0; b < nvel; ++b) {
                len[b] = Index[b + 1] - Index[b] - 1;
                for (int l = 0; l <= len[b]; ++l) rows[(size_t)b * 8 + l] = coef[Index[b] + l];
            }
            CK(cudaMalloc(&c->d_ls_rows, sizeof(float) * rows.size()));
            CK(cudaMalloc(&c->d_ls_len, sizeof(int) * len.size()));
            CK(cudaMemcpy(c->d_ls_rows, rows.data(), sizeof(float) * rows.size(), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(c->d_ls_len, len.data(), sizeof(int) * len.size(), cudaMemcpyHostToDevice));
            G.ls_rows = c->d_ls_rows; G.ls_len = c->d_ls_len; G.ls_nbins = nvel;
        }
        // (1+hzx2_1) a power of two (e.g. hz == h): A*c0 is exact in float for every bin
        int ex = 0;
        G.cc0_exact = (std::frexp(G.A, &ex) == 0.5 && (double)(float)G.A == G.A) ? 1 : 0;
        G.cc0f = (float)G.A;
    } else {
        if (NC != G.nfdmax + 1) return rtm_fail(RTM_ERR_ARG, "rtm_set_operator: Taylor operator needs nfdmax+1=%d coefficients, got %d", G.nfdmax + 1, NC);
        for (int l = 0; l <= kMaxR; ++l) G.cTE[l] = l < NC ? coef[l] : 0.0f;
        G.cc0TE = G.A * (double)coef[0];
        G.cc0f = (float)G.cc0TE;
        G.cc0_exact = ((double)G.cc0f == G.cc0TE) ? 1 : 0;
    }
    // The stencil halo (shared-memory tile, TMA box) is sized by the longest operator that is
    // actually present, which can be far below nfdmax (e.g. fast models need length 2-3 only).
    G.mmax = G.nfdmax;
    if (G.iLSTE == 0) {
        G.mmax = 1;
        for (int i = 0; i < nvel; ++i) G.mmax = std::max(G.mmax, Index[i + 1] - Index[i] - 1);
    }
    c->RP = (G.mmax + 3) / 4 * 4;
    c->h_M.clear();
    if (G.iLSTE == 0) {
        c->h_M.resize(nvel);
        for (int i = 0; i < nvel; ++i) c->h_M[i] = Index[i + 1] - Index[i] - 1;
    }
    c->have_op = true;
    c->nvel = nvel;
    drop_graphs(c);
    return prepare_ls(c);
}

// ------------------------------------------------------------------------------------ launches
static int ring_period(const rtm_ctx* c, int ring_ctas, int total)  // see block_role()
{
    return ring_period_for(c->ring_interleave, ring_ctas, total, c->ring_spread);
}
// One launch = the interior tiles of one class (+ the ring tiles when do_ring).
// buf / p0buf: field buffers of slots k-1 / k-2 (buf < 0: store-all slab); frame: only the ring tiles
// (stream mode: the thin frame has its own kernel, the rest is streamed)
template <int RP, bool LS> static int launch_fwd(rtm_ctx* c, rtm_ctx::TileClass& k, cudaStream_t st, int ns, int buf, int p0buf, bool frame, FwdArgs a)
{
    const Geo& G = c->G;
    const int nring = a.do_ring ? 2 * G.nband + 2 * G.nside : 0;
    size_t smem = (size_t)Tile<RP, RTM_NR_F>::BYTES + 16 + (LS ? (size_t)slice_bytes(RP) : 0);
    if (a.do_ring) smem = std::max(smem, (size_t)ring_smem_floats(G.N2, G.mmax, RP) * 4);
    if (smem > k.smem_f) {  // per device, once
        CK(cudaFuncSetAttribute(fwd_step_kernel<RP, LS, RTM_NR_F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k.smem_f = smem;
    }
    a.tiles = frame ? nullptr : k.d_tiles_f; a.ntiles = frame ? 0 : k.n_f; a.fd_ntiles = make_fastdiv(a.ntiles);   // frame (stream mode): ring tiles only
    // L2 look-ahead distance: 148 CTAs tuned in round 1 at 8 shots per launch; with 64 shots per launch 296-444 is the flat
    // optimum (forward step 273 -> 252 us), with one shot per launch on the large grids 148 stays ahead by ~1 % (profiles/r2_c28_*)
    a.lookahead = c->lookahead_f_auto ? (ns >= 48 ? 296 : 148) : c->lookahead_f;
    dim3 grid((unsigned)((nring + a.ntiles) * ns));
    if (grid.x == 0 || c->dry) return RTM_OK;
    a.ring_period = ring_period(c, nring * ns, (int)grid.x); a.fd_period = make_fastdiv(a.ring_period);
    ++c->nlaunch;
    a.lookahead_p0 = 1;
    a.tma_s0_p0 = buf < 0 ? a.tma_s0 - c->S : 0;
    fwd_step_kernel<RP, LS, RTM_NR_F><<<grid, kThreads, smem, st>>>(buf < 0 ? k.tmap_store : k.tmap_f[buf],
                                                                 buf < 0 ? k.tmap_store : k.tmap_f[p0buf], G, a);
    return RTM_OK;
}
// frame: only the tiles that the two-step kernel does not cover
template <int RP, bool LS, bool STORE> static int launch_bwd(rtm_ctx* c, rtm_ctx::TileClass& k, cudaStream_t st, int ns, int s1, int r1, int s0, int r0, BwdArgs a, bool frame)
{
    const Geo& G = c->G;
    const int nring = a.do_ring ? 2 * G.nband + 2 * G.nside : 0;
    size_t smem = (size_t)(STORE ? 1 : 2) * Tile<RP, RTM_NR_B>::BYTES + 16 + (LS ? (size_t)slice_bytes(RP) : 0);
    if (a.do_ring) smem = std::max(smem, (size_t)ring_smem_floats(G.N2, G.mmax, RP) * 4);
    if (smem > k.smem_b) {
        CK(cudaFuncSetAttribute(bwd_step_kernel<RP, LS, RTM_NR_B, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k.smem_b = smem;
    }
    a.tiles = frame ? k.d_tiles_bf : k.d_tiles_b; a.ntiles = frame ? k.n_bf : k.n_b;
    if (frame && k.stream_mode) { a.tiles = nullptr; a.ntiles = 0; }   // ring tiles only: thin_frame_kernel + stream2_kernel do the interior
    a.fd_ntiles = make_fastdiv(a.ntiles);
    dim3 grid((unsigned)((nring + a.ntiles) * ns));
    if (grid.x == 0 || c->dry) return RTM_OK;
    a.ring_period = ring_period(c, nring * ns, (int)grid.x); a.fd_period = make_fastdiv(a.ring_period);
    ++c->nlaunch;
    a.lookahead = c->lookahead_b; a.lookahead_p0 = c->lookahead_p0;
    bwd_step_kernel<RP, LS, RTM_NR_B, STORE><<<grid, kThreads, smem, st>>>(k.tmap_b[STORE ? r1 : s1], k.tmap_b[r1], k.tmap_b[STORE ? r0 : s0],
                                                                          k.tmap_b[r0], G, a);
    return RTM_OK;
}
template <int RP, bool LS> static int launch_bwd2(rtm_ctx* c, rtm_ctx::TileClass& k, cudaStream_t st, int ns, int s1, int r1, int s0, int r0, Bwd2Args a, bool border)
{
    if (k.n_b2 == 0) return RTM_OK;
    const size_t smem = (size_t)Tile2<RP>::BYTES + 16 + (LS ? (size_t)slice_bytes(RP) : 0);
    if (smem > k.smem_b2) {
        CK(cudaFuncSetAttribute(bwd2_step_kernel<RP, LS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k.smem_b2 = smem;
    }
    a.tiles = border ? k.d_tiles_ib : k.d_tiles_ii; a.ntiles = border ? k.n_ib : k.n_ii;
    a.rect_t0 = k.ii_rect[0]; a.rect_nx = border ? 0 : k.ii_rect[1]; a.rect_dz = k.ii_rect[2];
    a.fd_ntiles = make_fastdiv(a.ntiles); a.fd_rect = make_fastdiv(a.rect_nx); a.lookahead = c->lookahead_b2;
    if (c->dry || a.ntiles == 0) return RTM_OK;
    ++c->nlaunch;
    a.lookahead_more = c->lookahead_more;
    bwd2_step_kernel<RP, LS><<<(unsigned)(a.ntiles * ns), kThreads, smem, st>>>(k.tmap_b2[s1], k.tmap_b2[r1], k.tmap_b2[s0], k.tmap_b2[r0],
                                                                              c->tmap_acc, c->G, a);
    return RTM_OK;
}
// The same pair of steps by the z-streaming kernel: one CTA per column segment and shot.
static int launch_stream_bwd(rtm_ctx* c, rtm_ctx::TileClass& k, cudaStream_t st, int ns, int s1, int r1, int s0, int r0, const Bwd2Args& a2, bool border)
{
    using T = Strm<4>;
    const bool ls = c->G.iLSTE == 0;
    if (!k.smem_s2) {
        CK(cudaFuncSetAttribute(stream2_kernel<4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::bytes(true)));
        CK(cudaFuncSetAttribute(stream2_kernel<4, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaFuncSetAttribute(stream2_kernel<4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::bytes(true)));
        CK(cudaFuncSetAttribute(stream2_kernel<4, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        k.smem_s2 = true;
    }
    StrmArgs a{};
    a.Ak[0] = a2.Sk; a.Ak[1] = a2.Rk; a.Bk[0] = a2.Skm; a.Bk[1] = a2.Rkm;
    a.src = a2.src; a.wavelet_a = a2.wavelet_k; a.wavelet_b = a2.wavelet_km; a.k = a2.k; a.nshots = ns;
    a.segs = border ? k.d_segs_ib : k.d_segs_ii; a.nseg = border ? k.n_segs_ib : k.n_segs_ii;
    a.fd_nseg = make_fastdiv(a.nseg);
    a.seis = a2.seis; a.gather = nullptr;
    a.sumS = a2.sumS; a.sumR = a2.sumR; a.rel1 = a2.rel1; a.rel2 = a2.rel2;
    a.xend = c->G.NX - c->G.N2 - k.RP; a.zend = c->G.NZ - c->G.N2 - k.RP;
    if (c->dry || a.nseg == 0) return RTM_OK;
    StrmMaps tm;
    tm.cur[0] = k.tmap_s_cur[s1]; tm.cur[1] = k.tmap_s_cur[r1];
    tm.prev[0] = k.tmap_s_prev[s0]; tm.prev[1] = k.tmap_s_prev[r0];
    for (int i = 0; i < 4; ++i) tm.acc[i] = c->tmap_s_acc[i];
    ++c->nlaunch;
    if (ls) stream2_kernel<4, true, true><<<(unsigned)(a.nseg * ns), T::kThreadsS, T::bytes(true), st>>>(tm, c->G, a);
    else    stream2_kernel<4, true, false><<<(unsigned)(a.nseg * ns), T::kThreadsS, T::bytes(true), st>>>(tm, c->G, a);
    return RTM_OK;
}
// The thin frame (stream mode), one slot: backward (b1 = current source / receiver buffers, P0/P2 in `a`) or forward.
template <bool BWD> static int launch_thin(rtm_ctx* c, rtm_ctx::TileClass& k, cudaStream_t st, int ns, int cur0, int cur1, ThinArgs a)
{
    using T = Thin<4>;
    if (!k.smem_t) {
        CK(cudaFuncSetAttribute(thin_frame_kernel<4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::bytes(true)));
        CK(cudaFuncSetAttribute(thin_frame_kernel<4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::bytes(true)));
        CK(cudaFuncSetAttribute(thin_frame_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::bytes(false)));
        k.smem_t = true;
    }
    a.nshots = ns; a.tiles = k.d_thin; a.ntiles = k.n_thin; a.fd_ntiles = make_fastdiv(k.n_thin);
    if (c->dry || k.n_thin == 0) return RTM_OK;
    ThinMaps tm;
    tm.row[0] = k.tmap_t_row[cur0]; tm.col[0] = k.tmap_t_col[cur0];
    tm.row[1] = k.tmap_t_row[cur1]; tm.col[1] = k.tmap_t_col[cur1];
    ++c->nlaunch;
    if (BWD && c->G.iLSTE == 0) thin_frame_kernel<4, BWD, true><<<(unsigned)(k.n_thin * ns), T::kThreadsT, T::bytes(BWD), st>>>(tm, c->G, a);
    else                        thin_frame_kernel<4, BWD, false><<<(unsigned)(k.n_thin * ns), T::kThreadsT, T::bytes(BWD), st>>>(tm, c->G, a);
    return RTM_OK;
}
static int dispatch_bwd2_class(rtm_ctx* c, rtm_ctx::TileClass& k, cudaStream_t st, int ns, int s1, int r1, int s0, int r0, const Bwd2Args& a, bool border)
{
    const bool ls = c->G.iLSTE == 0;
    if (k.n_b2 == 0 && !k.stream_mode) return RTM_OK;
    if (k.stream_mode) return launch_stream_bwd(c, k, st, ns, s1, r1, s0, r0, a, border);
    switch (k.RP) {
    case 4: return ls ? launch_bwd2<4, true>(c, k, st, ns, s1, r1, s0, r0, a, border) : launch_bwd2<4, false>(c, k, st, ns, s1, r1, s0, r0, a, border);
    case 8: return ls ? launch_bwd2<8, true>(c, k, st, ns, s1, r1, s0, r0, a, border) : launch_bwd2<8, false>(c, k, st, ns, s1, r1, s0, r0, a, border);
    }
    return rtm_fail(RTM_ERR_ARG, "two-step kernel: unsupported operator radius %d", k.RP);
}
// The classes of one time step touch disjoint cells: they are forked onto side streams so that
// small classes share the GPU with the large ones, and joined before the next step.
template <class Launch> static int fork_join(rtm_ctx* c, Launch launch, cudaStream_t serial = nullptr)
{
    const int n = (int)c->classes.size();
    if (serial) {  // all classes one after the other on the given stream
        for (int i = 0; i < n; ++i)
            if (int rc = launch(c->classes[i], serial, i == 0)) return rc;
        return RTM_OK;
    }
    if (n > 1) {
        CK(cudaEventRecord(c->fork_ev, c->stream));
        for (int i = 1; i < n && i <= 3; ++i) CK(cudaStreamWaitEvent(c->aux[i - 1], c->fork_ev, 0));
    }
    for (int i = 0; i < n; ++i) {
        cudaStream_t st = (i >= 1 && i <= 3) ? c->aux[i - 1] : c->stream;
        if (int rc = launch(c->classes[i], st, i == 0)) return rc;
    }
    for (int i = 1; i < n && i <= 3; ++i) {
        CK(cudaEventRecord(c->join_ev[i - 1], c->aux[i - 1]));
        CK(cudaStreamWaitEvent(c->stream, c->join_ev[i - 1], 0));
    }
    return RTM_OK;
}
// buf: index of the field buffer holding slot k-1, or -1 for the store-all slab
// The ring tiles of one slot by ring_kernel: cur / prev = field buffers of the current / previous slot of the field that
// carries the absorbing boundary (forward field; backward: receiver field).
template <int RP, bool BWD, bool LS> static int launch_ring_t(rtm_ctx* c, cudaStream_t st, unsigned grid, int smem, const RingMaps& tm, const RingArgs& a)
{
    static thread_local int granted[64] = {0};   // per device: dynamic shared memory already granted to this instantiation
    if (granted[c->device & 63] < smem) {
        CK(cudaFuncSetAttribute(ring_kernel<RP, BWD, LS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        granted[c->device & 63] = smem;
    }
    if (c->dry) return RTM_OK;
    ++c->nlaunch;
    ring_kernel<RP, BWD, LS><<<grid, kThreads, smem, st>>>(tm, c->G, a);
    return RTM_OK;
}
template <bool BWD> static int launch_ring(rtm_ctx* c, cudaStream_t st, int ns, int cur, int prev, RingArgs a)
{
    const int smem = c->rgeo.smem_bytes();
    a.nshots = ns; a.rc = RingCoef{c->d_ring_coef, c->d_ring_meta}; a.rg = c->rgeo;
    RingMaps tm;
    tm.p1b = c->tmap_r_p1b[cur]; tm.p1s = c->tmap_r_p1s[cur]; tm.p0b = c->tmap_r_p0b[prev]; tm.p0s = c->tmap_r_p0s[prev];
    tm.avb = c->tmap_r_avb; tm.avs = c->tmap_r_avs;
    const unsigned grid = (unsigned)(c->rgeo.ntiles * ns);
    const bool ls = c->G.iLSTE == 0;
    switch (c->RP) {
    case 4:  return ls ? launch_ring_t<4, BWD, true>(c, st, grid, smem, tm, a) : launch_ring_t<4, BWD, false>(c, st, grid, smem, tm, a);
    case 8:  return ls ? launch_ring_t<8, BWD, true>(c, st, grid, smem, tm, a) : launch_ring_t<8, BWD, false>(c, st, grid, smem, tm, a);
    case 12: return ls ? launch_ring_t<12, BWD, true>(c, st, grid, smem, tm, a) : launch_ring_t<12, BWD, false>(c, st, grid, smem, tm, a);
    case 16: return ls ? launch_ring_t<16, BWD, true>(c, st, grid, smem, tm, a) : launch_ring_t<16, BWD, false>(c, st, grid, smem, tm, a);
    }
    return rtm_fail(RTM_ERR_ARG, "unsupported operator radius %d", c->RP);
}
static RingArgs ring_args_fwd(rtm_ctx* c, const FwdArgs& f)
{
    RingArgs r{};
    r.P2 = f.P2; r.SX = nullptr; r.src = f.src; r.wavelet = f.wavelet; r.inject = 1; r.k = f.k; r.st = f.st;
    r.seis = nullptr; r.gather = f.gather; r.sum_double = c->G.iLSTE == 0 ? 0 : 1;
    return r;
}
static int dispatch_fwd(rtm_ctx* c, int ns, int buf, int p0buf, FwdArgs a, bool frame = false, cudaStream_t serial = nullptr)
{
    if (c->ring_ready && buf >= 0 && (frame || (c->ring2_fwd && !c->ring_frame_only))) {
        // the ring by its own kernel: alone (frame step of the pair loop), or next to the interior launch on a side stream
        if (frame) return launch_ring<false>(c, serial ? serial : c->stream, ns, buf, p0buf, ring_args_fwd(c, a));
        CK(cudaEventRecord(c->ring_fork, c->stream));
        CK(cudaStreamWaitEvent(c->ring_stream, c->ring_fork, 0));
        if (int rc = launch_ring<false>(c, c->ring_stream, ns, buf, p0buf, ring_args_fwd(c, a))) return rc;
        CK(cudaEventRecord(c->ring_join, c->ring_stream));
        if (c->stream1_fwd && c->classes.size() == 1 && c->classes[0].stream_mode && c->classes[0].n_segs_f1 > 0) {
            // the interior by the single-step streaming kernel
            rtm_ctx::TileClass& k = c->classes[0];
            StrmArgs sa{};
            sa.Ak[0] = a.P2; sa.src = a.src; sa.wavelet_a = a.wavelet; sa.k = a.k; sa.nshots = ns;
            sa.segs = k.d_segs_f1; sa.nseg = k.n_segs_f1; sa.fd_nseg = make_fastdiv(k.n_segs_f1);
            sa.gather = a.gather; sa.xend = c->G.NX - c->G.N2; sa.zend = c->G.NZ - c->G.N2;
            if (!c->dry) {
                Strm1Maps tm1;
                tm1.cur = k.tmap_s_prev[buf]; tm1.prev = k.tmap_s_own[p0buf];
                ++c->nlaunch;
                stream1_fwd_kernel<4><<<(unsigned)(k.n_segs_f1 * ns), Strm1<4>::kThreadsS, Strm1<4>::bytes(), c->stream>>>(tm1, c->G, sa);
            }
            CK(cudaStreamWaitEvent(c->stream, c->ring_join, 0));
            return RTM_OK;
        }
        const bool ls0 = c->G.iLSTE == 0;
        int rc = fork_join(c, [&](rtm_ctx::TileClass& k, cudaStream_t st, bool) -> int {
            a.do_ring = 0;
            switch (k.RP) {
            case 4:  return ls0 ? launch_fwd<4, true>(c, k, st, ns, buf, p0buf, false, a) : launch_fwd<4, false>(c, k, st, ns, buf, p0buf, false, a);
            case 8:  return ls0 ? launch_fwd<8, true>(c, k, st, ns, buf, p0buf, false, a) : launch_fwd<8, false>(c, k, st, ns, buf, p0buf, false, a);
            case 12: return ls0 ? launch_fwd<12, true>(c, k, st, ns, buf, p0buf, false, a) : launch_fwd<12, false>(c, k, st, ns, buf, p0buf, false, a);
            case 16: return ls0 ? launch_fwd<16, true>(c, k, st, ns, buf, p0buf, false, a) : launch_fwd<16, false>(c, k, st, ns, buf, p0buf, false, a);
            }
            return rtm_fail(RTM_ERR_ARG, "unsupported operator radius %d", k.RP);
        });
        if (rc) return rc;
        CK(cudaStreamWaitEvent(c->stream, c->ring_join, 0));
        return RTM_OK;
    }
    const bool ls = c->G.iLSTE == 0;
    return fork_join(c, [&](rtm_ctx::TileClass& k, cudaStream_t st, bool first) -> int {
        a.do_ring = first ? 1 : 0;
        switch (k.RP) {
        case 4:  return ls ? launch_fwd<4, true>(c, k, st, ns, buf, p0buf, frame, a) : launch_fwd<4, false>(c, k, st, ns, buf, p0buf, frame, a);
        case 8:  return ls ? launch_fwd<8, true>(c, k, st, ns, buf, p0buf, frame, a) : launch_fwd<8, false>(c, k, st, ns, buf, p0buf, frame, a);
        case 12: return ls ? launch_fwd<12, true>(c, k, st, ns, buf, p0buf, frame, a) : launch_fwd<12, false>(c, k, st, ns, buf, p0buf, frame, a);
        case 16: return ls ? launch_fwd<16, true>(c, k, st, ns, buf, p0buf, frame, a) : launch_fwd<16, false>(c, k, st, ns, buf, p0buf, frame, a);
        }
        return rtm_fail(RTM_ERR_ARG, "unsupported operator radius %d", k.RP);
    }, serial);
}
// Two forward steps (slots k, k+1) of the inner segments by the z-streaming kernel.
static int launch_stream_fwd(rtm_ctx* c, rtm_ctx::TileClass& k, cudaStream_t st, int ns, int b1, int b0, int bk, int bk1, const FwdArgs& f,
                             float wavelet_k1, bool border)
{
    using T = Strm<4>;
    if (!k.smem_s2f) {
        CK(cudaFuncSetAttribute(stream2_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::bytes(false)));
        CK(cudaFuncSetAttribute(stream2_kernel<4, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        k.smem_s2f = true;
    }
    StrmArgs a{};
    a.Ak[0] = c->field[bk]; a.Bk[0] = c->field[bk1];
    a.src = f.src; a.wavelet_a = f.wavelet; a.wavelet_b = wavelet_k1; a.k = f.k; a.nshots = ns;
    a.segs = border ? k.d_segs_ib : k.d_segs_ii; a.nseg = border ? k.n_segs_ib : k.n_segs_ii;
    a.fd_nseg = make_fastdiv(a.nseg);
    a.gather = f.gather;
    a.xend = c->G.NX - c->G.N2 - k.RP; a.zend = c->G.NZ - c->G.N2 - k.RP;
    if (c->dry || a.nseg == 0) return RTM_OK;
    StrmMaps tm;
    tm.cur[0] = k.tmap_s_cur[b1]; tm.prev[0] = k.tmap_s_prev[b0];
    tm.cur[1] = tm.cur[0]; tm.prev[1] = tm.prev[0];
    for (int i = 0; i < 4; ++i) tm.acc[i] = c->tmap_s_acc[i];
    ++c->nlaunch;
    stream2_kernel<4, false><<<(unsigned)(a.nseg * ns), T::kThreadsS, T::bytes(false), st>>>(tm, c->G, a);
    return RTM_OK;
}
template <bool STORE> static int dispatch_bwd_t(rtm_ctx* c, int ns, int s1, int r1, int s0, int r0, BwdArgs a, bool frame, cudaStream_t serial, bool no_ring = false)
{
    const bool ls = c->G.iLSTE == 0;
    return fork_join(c, [&](rtm_ctx::TileClass& k, cudaStream_t st, bool first) -> int {
        a.do_ring = (first && !no_ring) ? 1 : 0;
        switch (k.RP) {
        case 4:  return ls ? launch_bwd<4, true, STORE>(c, k, st, ns, s1, r1, s0, r0, a, frame) : launch_bwd<4, false, STORE>(c, k, st, ns, s1, r1, s0, r0, a, frame);
        case 8:  return ls ? launch_bwd<8, true, STORE>(c, k, st, ns, s1, r1, s0, r0, a, frame) : launch_bwd<8, false, STORE>(c, k, st, ns, s1, r1, s0, r0, a, frame);
        case 12: return ls ? launch_bwd<12, true, STORE>(c, k, st, ns, s1, r1, s0, r0, a, frame) : launch_bwd<12, false, STORE>(c, k, st, ns, s1, r1, s0, r0, a, frame);
        case 16: return ls ? launch_bwd<16, true, STORE>(c, k, st, ns, s1, r1, s0, r0, a, frame) : launch_bwd<16, false, STORE>(c, k, st, ns, s1, r1, s0, r0, a, frame);
        }
        return rtm_fail(RTM_ERR_ARG, "unsupported operator radius %d", k.RP);
    }, serial);
}
static int dispatch_bwd(rtm_ctx* c, int ns, int s1, int r1, int s0, int r0, const BwdArgs& a, bool frame = false, cudaStream_t serial = nullptr)
{
    if (frame && c->ring_ready && !c->store_mode && c->classes.size() == 1 && c->classes[0].stream_mode) {
        // stream mode: a frame step is the ring alone (receiver field + strip restore), by ring_kernel
        RingArgs r{};
        r.P2 = a.R2; r.SX = a.S2; r.src = a.src; r.inject = 0; r.k = a.k; r.st = a.st; r.seis = a.seis; r.sum_double = 0;
        return launch_ring<true>(c, serial ? serial : c->stream, ns, r1, r0, r);
    }
    // (measured on 4096^2, one shot per launch, profiles/r2_c23_*: radius 12 +2.4 %, radius 8 -11 % -- there the ring launch ends up
    //  behind the interior launch instead of next to it; RTM_RING2_BWD=2 forces it for every radius)
    if (!frame && !serial && c->ring_ready && c->ring2_bwd && (c->RP != 8 || c->ring2_bwd_all) && !c->ring_frame_only && !c->store_mode) {
        // a single step of all tiles: the ring by its own kernel on a side stream, the interior tiles without ring CTAs
        RingArgs r{};
        r.P2 = a.R2; r.SX = a.S2; r.src = a.src; r.inject = 0; r.k = a.k; r.st = a.st; r.seis = a.seis; r.sum_double = 0;
        CK(cudaEventRecord(c->ring_fork, c->stream));
        CK(cudaStreamWaitEvent(c->ring_stream, c->ring_fork, 0));
        if (int rc = launch_ring<true>(c, c->ring_stream, ns, r1, r0, r)) return rc;
        CK(cudaEventRecord(c->ring_join, c->ring_stream));
        if (int rc = dispatch_bwd_t<false>(c, ns, s1, r1, s0, r0, a, false, nullptr, true)) return rc;
        CK(cudaStreamWaitEvent(c->stream, c->ring_join, 0));
        return RTM_OK;
    }
    return c->store_mode ? dispatch_bwd_t<true>(c, ns, s1, r1, s0, r0, a, frame, serial) : dispatch_bwd_t<false>(c, ns, s1, r1, s0, r0, a, frame, serial);
}

static void drop_graphs(rtm_ctx* c)
{
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.second);
    c->graphs.clear();
    c->graph_launches.clear();
}

// Replay (or capture, the first time) the launches issued by `body` as one CUDA graph.
template <class Body> static int run_as_graph(rtm_ctx* c, long long key, Body body)
{
    auto it = c->graphs.find(key);
    if (it == c->graphs.end()) {
        cudaGraph_t graph = nullptr;
        const long n0 = c->nlaunch;
        CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = body();
        const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        c->graph_launches[key] = c->nlaunch - n0;
        c->nlaunch = n0;
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return rtm_fail(RTM_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(e));
        cudaGraphExec_t exec = nullptr;
        const cudaError_t e2 = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) return rtm_fail(RTM_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
        it = c->graphs.emplace(key, exec).first;
    }
    CK(cudaGraphLaunch(it->second, c->stream));
    c->nlaunch += c->graph_launches[key];
    return RTM_OK;
}

static int ensure_strips(rtm_ctx* c)
{
    if (c->st.up) return RTM_OK;
    const Geo& G = c->G;
    const size_t nx = (size_t)c->S * G.NT * G.nfdmax * G.mod_NX * 4;
    const size_t nz = (size_t)c->S * G.NT * G.nfdmax * G.mod_NZ * 4;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (2 * (nx + nz) > free_b)
        return rtm_fail(RTM_ERR_ARG, "boundary strips for %d shots x %d steps need %.1f GB, %.1f GB free: lower max_batch", c->S, G.NT, 2 * (nx + nz) / 1e9, free_b / 1e9);
    CK(cudaMalloc(&c->st.up, nx)); CK(cudaMalloc(&c->st.dw, nx));
    CK(cudaMalloc(&c->st.lf, nz)); CK(cudaMalloc(&c->st.rt, nz));
    return RTM_OK;
}

// Forward loop for `ns` shots (sources already in d_src).  On return field[*last1] holds slot
// NT-1 and field[*last0] slot NT-2.
static int run_forward(rtm_ctx* c, int ns, const int* r_u, const int* r_x, bool strips, float* gather,
                       int nsnap, const int* snap_k, float* snaps_host, bool use_store, float** last1,
                       float** last0)
{
    const Geo& G = c->G;
    std::vector<int2> src(ns);
    for (int s = 0; s < ns; ++s) {
        if (r_u[s] < 0 || r_u[s] >= G.NZ || r_x[s] < 0 || r_x[s] >= G.NX)
            return rtm_fail(RTM_ERR_ARG, "source %d at (%d,%d) outside the %dx%d grid", s, r_u[s], r_x[s], G.NZ, G.NX);
        src[s] = make_int2(r_u[s], r_x[s]);
    }
    CK(cudaMemcpyAsync(c->d_src, src.data(), sizeof(int2) * ns, cudaMemcpyHostToDevice, c->stream));
    // time slot k lives in one of three rotating buffers, or -- store-all mode -- in its own slab
    const size_t slab = (size_t)c->S * G.shot_stride;
    // forward pass in pairs (inner segments two slots per pass, ring + frame tiles singly): 4 rotating buffers
    bool pairs = false;
    if (!use_store && nsnap == 0 && c->fuse2_fwd != 0 && G.iLSTE != 0 && c->classes.size() == 1 && c->classes[0].stream_mode)
        pairs = c->fuse2_fwd == 1 || c->fuse2_forced || (long)(c->classes[0].stream_cells / 4096.0) * ns >= rtm_ctx::kFuse2MinCtas;
    const int NB = pairs ? 4 : 3;
    auto slot = [&](int k) -> float* { return use_store ? c->store + (size_t)k * slab : c->field[k % NB]; };
    if (use_store) {
        CK(cudaMemsetAsync(slot(0), 0, 2 * slab * 4, c->stream));
    } else {
        for (int b = 0; b < NB; ++b) CK(cudaMemsetAsync(c->field[b], 0, c->field_floats * 4, c->stream));
    }
    const float fw1 = (float)(rtm::ricker(0.0f, c->p.f0) / 2.0);  // :803
    init_source_kernel<<<ns, 1, 0, c->stream>>>(slot(1), G, c->d_src, fw1);
    Strips st = (strips && !use_store) ? c->st : Strips{nullptr, nullptr, nullptr, nullptr};
    {
        size_t m = std::max({(size_t)G.nfdmax * G.mod_NX, (size_t)G.nfdmax * G.mod_NZ, (size_t)G.n});
        dim3 grid((unsigned)((m + 255) / 256), ns);
        if (st.up || gather) {
            strips_from_field_kernel<<<grid, 256, 0, c->stream>>>(slot(0), G, st, 0, gather);
            strips_from_field_kernel<<<grid, 256, 0, c->stream>>>(slot(1), G, st, 1, gather);
        }
    }
    auto snapshot = [&](int k) -> int {
        for (int i = 0; i < nsnap; ++i) {
            if (snap_k[i] != k) continue;
            CK(cudaStreamSynchronize(c->stream));
            for (int s = 0; s < ns; ++s)
                CK(cudaMemcpy2DAsync(snaps_host + ((size_t)s * nsnap + i) * G.NZ * G.NX, (size_t)G.NX * 4,
                                     slot(k) + (size_t)s * G.shot_stride + G.padL, (size_t)G.pitch * 4,
                                     (size_t)G.NX * 4, G.NZ, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        return RTM_OK;
    };
    if (nsnap) { if (int rc = snapshot(0)) return rc; if (int rc = snapshot(1)) return rc; }
    int NT2;
    rtm_derived(c->p.h, c->p.hz, c->p.tao, c->p.tao, c->p.f0, 2, nullptr, &NT2, nullptr, nullptr, nullptr, nullptr, nullptr);
    auto wavelet = [&](int k) { return (k < NT2) ? rtm::ricker((k - 1) * c->p.tao, c->p.f0) : 0.0f; };  // :812-813
    auto args = [&](int k) {
        FwdArgs a{};
        a.P1 = slot(k - 1); a.P0 = slot(k - 2); a.P2 = slot(k);
        a.src = c->d_src;
        a.wavelet = wavelet(k);
        a.k = k; a.nshots = ns; a.st = st; a.gather = gather;
        a.tma_s0 = use_store ? (k - 1) * c->S : 0;
        return a;
    };
    auto step = [&](int k) -> int {
        return dispatch_fwd(c, ns, use_store ? -1 : (k - 1) % NB, (k - 2 + NB) % NB, args(k));
    };
    // slots kfirst, kfirst+1, ... NT-1 in pairs (k, k+1); same two-stream pipeline as the backward pass:
    //   A (main):  inner-inner segments, two-step kernel
    //   B (aux 2): ring + frame tiles slot k -> inner segments next to the frame -> ring + frame tiles slot k+1
    auto pair_loop = [&](int kfirst) -> int {
        cudaStream_t A = c->stream, B = c->aux[2];
        rtm_ctx::TileClass& kc = c->classes[0];
        CK(cudaEventRecord(c->fork_ev, A));
        CK(cudaStreamWaitEvent(B, c->fork_ev, 0));
        int j = 0;
        for (int k = kfirst; k + 1 < G.NT; k += 2, ++j) {
            const int b1 = (k - 1) % NB, b0 = (k - 2) % NB, bk = k % NB, bk1 = (k + 1) % NB;
            const FwdArgs a0 = args(k), a1 = args(k + 1);
            if (j > 0) CK(cudaStreamWaitEvent(A, c->ev_ib[(j - 1) & 1], 0));
            if (int rc = launch_stream_fwd(c, kc, A, ns, b1, b0, bk, bk1, a0, a1.wavelet, false)) return rc;
            CK(cudaEventRecord(c->ev_ii[j & 1], A));
            auto thin = [&](const FwdArgs& f, int cur, int prev, int out) -> int {
                ThinArgs t{};
                t.P0[0] = c->field[prev]; t.P2[0] = c->field[out];
                t.src = f.src; t.wavelet = f.wavelet; t.k = f.k; t.gather = f.gather;
                return launch_thin<false>(c, kc, B, ns, cur, cur, t);
            };
            if (int rc = dispatch_fwd(c, ns, b1, b0, a0, true, B)) return rc;   // (slot k of the thin frame: the ib segments' halo)
            if (j > 0) CK(cudaStreamWaitEvent(B, c->ev_ii[(j - 1) & 1], 0));
            if (int rc = launch_stream_fwd(c, kc, B, ns, b1, b0, bk, bk1, a0, a1.wavelet, true)) return rc;
            CK(cudaEventRecord(c->ev_ib[j & 1], B));
            if (int rc = dispatch_fwd(c, ns, bk, b1, a1, true, B)) return rc;
            if (int rc = thin(a1, bk, b1, bk1)) return rc;
        }
        CK(cudaEventRecord(c->join_ev[2], B));
        CK(cudaStreamWaitEvent(A, c->join_ev[2], 0));
        return RTM_OK;
    };
    auto steps_from = [&](int k) -> int {   // slots k .. NT-1
        if (pairs) {
            if ((G.NT - k) % 2) { if (int rc = step(k)) return rc; ++k; }
            if (k + 1 < G.NT) return pair_loop(k);
            return RTM_OK;
        }
        for (; k < G.NT; ++k) if (int rc = step(k)) return rc;
        return RTM_OK;
    };
    const long nl0 = c->nlaunch;
    CK(cudaEventRecord(c->ev0, c->stream));
    if (c->use_graphs && nsnap == 0 && G.NT > 3) {
        if (int rc = step(2)) return rc;  // (also sets the kernel's shared-memory attribute before capture)
        if (pairs) {   // kernel attributes of the pair path cannot be set during capture: a dry pass
            c->dry = true;
            const int rc = pair_loop(G.NT - 2);
            c->dry = false;
            if (rc) return rc;
        }
        const long long key = 1 + 2 * (st.up ? 1 : 0) + 4 * (gather ? 1 : 0) + 8 * (use_store ? 1 : 0) + 16LL * ns + (pairs ? (1LL << 40) : 0);
        if (int rc = run_as_graph(c, key, [&]() -> int { return steps_from(3); })) return rc;
    } else {
        if (nsnap == 0) {
            if (int rc = steps_from(2)) return rc;
        } else {
            for (int k = 2; k < G.NT; ++k) {
                if (int rc = step(k)) return rc;
                if (int rc = snapshot(k)) return rc;
            }
        }
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    const double cu = (double)(G.NT - 2) * G.NZ * G.NX * ns;
    c->stats.cell_updates += cu;
    c->stats.algorithmic_bytes += cu * 16.0;
    {   // bytes of the schedule that ran (include/rtm_b200.h, rtm_stats)
        const double steps = G.NT - 2, ring = (double)G.NZ * G.NX - (double)G.mod_NZ * G.mod_NX, inner = (double)G.mod_NZ * G.mod_NX;
        double pair_cells = 0;
        if (pairs) {
            const rtm_ctx::TileClass& kc = c->classes[0];
            pair_cells = kc.stream_cells;
        }
        const double pair_steps = pairs ? 2.0 * ((G.NT - 3) / 2) : 0.0;           // slots 3.. in pairs (slot 2 and an odd rest singly)
        const double strips = st.up ? 2.0 * G.nfdmax * ((double)G.mod_NX + G.mod_NZ) * 4 : 0.0;
        c->stats.executed_bytes_forward += ns * (pair_steps * pair_cells * 8.0 + (steps * inner - pair_steps * pair_cells) * 12.0 +
                                                 steps * (ring * 12.0 + strips)) + steps * (double)G.NZ * G.NX * 4.0;
        c->stats.pair_cell_steps_forward += ns * pair_steps * pair_cells;
    }
    c->stats.forward_seconds += ms * 1e-3;
    c->stats.kernel_launches += (c->nlaunch - nl0) + 3;
    c->last_forward_ms = ms;
    *last1 = slot(G.NT - 1); *last0 = slot(G.NT - 2);
    return RTM_OK;
}

extern "C" int rtm_forward(rtm_ctx* c, int nshots, const int* r_u, const int* r_x, float* gathers,
                           int nsnap, const int* snap_k, float* snaps)
{
    if (!c || !r_u || !r_x || nshots < 1) return rtm_fail(RTM_ERR_ARG, "rtm_forward: bad argument");
    if (!c->have_model || !c->have_op) return rtm_fail(RTM_ERR_STATE, "rtm_forward: set the model and the operator first");
    if (nsnap && (!snap_k || !snaps)) return rtm_fail(RTM_ERR_ARG, "rtm_forward: snapshots requested without buffers");
    CK(cudaSetDevice(c->device));
    const Geo& G = c->G;
    for (int first = 0; first < nshots; first += c->S) {
        const int ns = std::min(c->S, nshots - first);
        float *l1, *l0;
        if (int rc = run_forward(c, ns, r_u + first, r_x + first, false, gathers ? c->d_traces : nullptr, nsnap,
                                 snap_k, snaps ? snaps + (size_t)first * nsnap * G.NZ * G.NX : nullptr, false, &l1, &l0))
            return rc;
        if (gathers) {
            dim3 grid((G.n + 31) / 32, (G.NT + 31) / 32, ns);  // [NT][n] -> [n][NT]
            transpose_traces_kernel<<<grid, dim3(32, 8), 0, c->stream>>>(c->d_traces, c->d_stage, G.NT, G.n);
            CK(cudaMemcpyAsync(gathers + (size_t)first * G.n * G.NT, c->d_stage, (size_t)ns * G.n * G.NT * 4,
                               cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        c->stats.shots += ns;
        c->stats.device_seconds += c->last_forward_ms * 1e-3;
    }
    return RTM_OK;
}

// One batch of the shot loop body (kernel.cu:798-990) with the traces already in d_traces.
static int migrate_batch(rtm_ctx* c, int ns, const int* r_u, const int* r_x, float* up, float* down, float* stable)
{
    const Geo& G = c->G;
    const bool store = c->store_mode;
    if (!store) if (int rc = ensure_strips(c)) return rc;
    float *l1, *l0;
    CK(cudaEventRecord(c->evA, c->stream));
    if (int rc = run_forward(c, ns, r_u, r_x, true, nullptr, 0, nullptr, nullptr, store, &l1, &l0)) return rc;
    // Backward pass.  Source field: Sa = slot k+2 (starts as slot NT-1), Sb = slot k+1 (NT-2), Sc/Sd
    // receive slots k / k-1; receiver field Ra..Rd likewise, starting from zero (store-all: only
    // the receiver buffers are used).
    int Sa = -1, Sb = -1, Sc = -1, Sd = -1, Ra = -1, Rb = -1, Rc = -1, Rd = -1;
    {
        int rest[rtm_ctx::kFields], n = 0;
        for (int b = 0; b < rtm_ctx::kFields; ++b) {
            if (c->field[b] == l1) Sa = b;
            else if (c->field[b] == l0) Sb = b;
            else rest[n++] = b;
        }
        if (store) { Ra = rest[0]; Rb = rest[1]; Rc = rest[2]; Rd = rest[3]; }
        else { Sc = rest[0]; Sd = rest[1]; Ra = rest[2]; Rb = rest[3]; Rc = rest[4]; Rd = rest[5]; }
    }
    CK(cudaMemsetAsync(c->field[Ra], 0, c->field_floats * 4, c->stream));
    CK(cudaMemsetAsync(c->field[Rb], 0, c->field_floats * 4, c->stream));
    CK(cudaMemsetAsync(c->field[Rc], 0, c->field_floats * 4, c->stream));
    const float fw1 = (float)(rtm::ricker(0.0f, c->p.f0) / 2.0);
    {
        dim3 grid((G.NX + 127) / 128, G.NZ, ns);
        acc_init_kernel<<<grid, 128, 0, c->stream>>>(G, l1, l0, c->d_src, fw1, c->acc[0], c->acc[1], c->acc[2], c->acc[3]);
    }
    int NT2;
    rtm_derived(c->p.h, c->p.hz, c->p.tao, c->p.tao, c->p.f0, 2, nullptr, &NT2, nullptr, nullptr, nullptr, nullptr, nullptr);
    const size_t slab = (size_t)c->S * G.shot_stride;
    auto wavelet = [&](int k) { return (k < NT2) ? rtm::ricker((k + 1) * c->p.tao, c->p.f0) : 0.0f; };  // :889-890
    // arguments of one single-step launch for slot k: source prev/out buffers, receiver prev/out
    auto args1 = [&](int k, int s0, int s2, int r0, int r1, int r2) {
        BwdArgs a{};
        a.Sk = store ? c->store + (size_t)k * slab : nullptr;
        a.S1 = nullptr; a.S02 = store ? nullptr : c->field[s0]; a.S2 = store ? nullptr : c->field[s2];
        a.R1 = c->field[r1]; a.R0 = c->field[r0]; a.R2 = c->field[r2];
        a.src = c->d_src;
        a.wavelet = wavelet(k);
        a.k = k; a.nshots = ns; a.st = c->st; a.seis = c->d_traces;
        a.sumS = c->acc[0]; a.sumR = c->acc[1]; a.rel1 = c->acc[2]; a.rel2 = c->acc[3];
        return a;
    };
    // one step, all tiles; the source slot k replaces slot k+2 in place
    auto bstep = [&](int k) -> int {
        if (int rc = dispatch_bwd(c, ns, store ? Rb : Sb, Rb, store ? Ra : Sa, Ra, args1(k, Sa, Sa, Ra, Rb, Rc))) return rc;
        std::swap(Sa, Sb);
        const int t = Ra; Ra = Rb; Rb = Rc; Rc = t;
        return RTM_OK;
    };
    // Slots kfirst, kfirst-1, ..., 0 in pairs (k, k-1), software-pipelined over two streams:
    //   A (main):  inner-inner tiles, two-step kernel                      L2ii(j)
    //   B (aux 2): ring + frame tiles slot k (single-step kernel)          L1(j)
    //              inner tiles next to the frame, two-step kernel          L2ib(j)
    //              ring + frame tiles slot k-1                             L3(j)
    // L2ii(j) needs L2ib(j-1) (its halo reaches into those tiles, and it overwrites what they read);
    // L2ib(j) needs L2ii(j-1) for the same two reasons; everything else is ordered by stream B.
    auto pair_loop = [&](int kfirst) -> int {
        cudaStream_t A = c->stream, B = c->aux[2];
        CK(cudaEventRecord(c->fork_ev, A));
        CK(cudaStreamWaitEvent(B, c->fork_ev, 0));
        int j = 0;
        for (int k = kfirst; k >= 1; k -= 2, ++j) {
            Bwd2Args a2{};
            a2.S0 = c->field[Sa]; a2.Sk = c->field[Sc]; a2.Skm = c->field[Sd];
            a2.R0 = c->field[Ra]; a2.Rk = c->field[Rc]; a2.Rkm = c->field[Rd];
            a2.src = c->d_src; a2.wavelet_k = wavelet(k); a2.wavelet_km = wavelet(k - 1);
            a2.k = k; a2.nshots = ns; a2.seis = c->d_traces;
            a2.sumS = c->acc[0]; a2.sumR = c->acc[1]; a2.rel1 = c->acc[2]; a2.rel2 = c->acc[3];
            if (j > 0) CK(cudaStreamWaitEvent(A, c->ev_ib[(j - 1) & 1], 0));
            for (auto& kc : c->classes)
                if (int rc = dispatch_bwd2_class(c, kc, A, ns, Sb, Rb, Sa, Ra, a2, false)) return rc;
            CK(cudaEventRecord(c->ev_ii[j & 1], A));
            // the thin frame (stream mode): slot k's values come out of the streamed segments' halo (ib launch), slot k-1 -- which
            // needs the ring of slot k -- from thin_frame_kernel, which also applies the imaging update of slot k there
            auto thin = [&](int kk, int s1, int r1, int s0, int r0, int s2, int r2) -> int {
                for (auto& kc : c->classes) {
                    if (!kc.stream_mode) continue;
                    ThinArgs t{};
                    t.P0[0] = c->field[s0]; t.P0[1] = c->field[r0]; t.P2[0] = c->field[s2]; t.P2[1] = c->field[r2];
                    t.src = c->d_src; t.wavelet = wavelet(kk); t.k = kk; t.seis = c->d_traces;
                    t.sumS = c->acc[0]; t.sumR = c->acc[1]; t.rel1 = c->acc[2]; t.rel2 = c->acc[3];
                    t.twice = 1;
                    if (int rc = launch_thin<true>(c, kc, B, ns, s1, r1, t)) return rc;
                }
                return RTM_OK;
            };
            // Stream mode: the ring of slot k depends on the previous pair only and touches no cell the ib launch writes, and
            // the ring of slot k-1 and the thin frame of slot k-1 write disjoint cells from the same inputs (ring of slot k,
            // ib launch): each ring launch runs on the ring stream next to its neighbour on B (RTM_RING_PAR=0: in B's order).
            const bool par = c->ring_par && c->ring_ready && c->classes.size() == 1 && c->classes[0].stream_mode;
            cudaStream_t RS = par ? c->ring_stream : B;
            if (par) { CK(cudaEventRecord(c->ring_fork, B)); CK(cudaStreamWaitEvent(RS, c->ring_fork, 0)); }
            if (int rc = dispatch_bwd(c, ns, Sb, Rb, Sa, Ra, args1(k, Sa, Sc, Ra, Rb, Rc), true, RS)) return rc;
            if (par) CK(cudaEventRecord(c->ring_join, RS));
            if (j > 0) CK(cudaStreamWaitEvent(B, c->ev_ii[(j - 1) & 1], 0));
            for (auto& kc : c->classes)
                if (int rc = dispatch_bwd2_class(c, kc, B, ns, Sb, Rb, Sa, Ra, a2, true)) return rc;
            CK(cudaEventRecord(c->ev_ib[j & 1], B));
            if (par) {
                CK(cudaStreamWaitEvent(B, c->ring_join, 0));
                CK(cudaEventRecord(c->ring_fork, B));
                CK(cudaStreamWaitEvent(RS, c->ring_fork, 0));
            }
            if (int rc = dispatch_bwd(c, ns, Sc, Rc, Sb, Rb, args1(k - 1, Sb, Sd, Rb, Rc, Rd), true, RS)) return rc;
            if (par) CK(cudaEventRecord(c->ring_join, RS));
            if (int rc = thin(k - 1, Sc, Rc, Sb, Rb, Sd, Rd)) return rc;
            if (par) CK(cudaStreamWaitEvent(B, c->ring_join, 0));
            std::swap(Sa, Sc); std::swap(Sb, Sd);
            std::swap(Ra, Rc); std::swap(Rb, Rd);
        }
        CK(cudaEventRecord(c->join_ev[2], B));
        CK(cudaStreamWaitEvent(A, c->join_ev[2], 0));
        return RTM_OK;
    };
    bool pairs = c->fuse2_on && !store;
    if (pairs) {
        long nii = 0, nb2 = 0;
        for (auto& kc : c->classes) { nii += kc.stream_mode ? (long)(kc.stream_cells / 4096.0) /* streamed cells in tile pairs of 128 x 32: independent of the segment length */ : kc.n_ii; nb2 += kc.stream_mode ? kc.n_segs_ii + kc.n_segs_ib : kc.n_b2; }
        pairs = nb2 > 0 && (c->fuse2_forced || nii * ns >= rtm_ctx::kFuse2MinCtas);
    }
    auto loop = [&](int kfirst) -> int {  // slots kfirst .. 0
        int k = kfirst;
        if (pairs) {
            if ((k + 1) % 2) { if (int rc = bstep(k)) return rc; --k; }
            if (k >= 1) if (int rc = pair_loop(k)) return rc;
        } else {
            for (; k >= 0; --k) if (int rc = bstep(k)) return rc;
        }
        return RTM_OK;
    };
    const long nl0 = c->nlaunch;
    CK(cudaEventRecord(c->ev0, c->stream));
    if (c->use_graphs && G.NT > 3) {
        {   // kernel attributes cannot be set during capture: a dry pass over both kinds of step
            const int sv[8] = {Sa, Sb, Sc, Sd, Ra, Rb, Rc, Rd};
            c->dry = true;
            int rc = bstep(G.NT - 3);
            if (!rc && pairs) rc = pair_loop(1);
            c->dry = false;
            Sa = sv[0]; Sb = sv[1]; Sc = sv[2]; Sd = sv[3]; Ra = sv[4]; Rb = sv[5]; Rc = sv[6]; Rd = sv[7];
            if (rc) return rc;
        }
        if (int rc = run_as_graph(c, 2 + 4 * (pairs ? 1 : 0) + 8 * (store ? 1 : 0) + 16LL * ns, [&]() -> int { return loop(G.NT - 3); }))
            return rc;
    } else {
        if (int rc = loop(G.NT - 3)) return rc;
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    // per-shot image post-processing and stacking
    const size_t ncell = (size_t)G.mod_NX * G.mod_NZ;
    CK(cudaMemsetAsync(c->d_maxbits, 0, sizeof(int) * ns, c->stream));
    {
        dim3 grid((G.mod_NZ + 127) / 128, G.mod_NX, ns);
        image_up_kernel<<<grid, 128, 0, c->stream>>>(G, c->acc[2], c->acc[3], c->vmax * c->vmax, c->d_up, c->d_maxbits);
        image_down_kernel<<<grid, 128, 0, c->stream>>>(G, c->acc[3], c->d_maxbits, c->p.whitecoe, c->d_down, c->d_stable);
        stack_add_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, c->stream>>>(ncell, ns, c->d_up, c->d_down, c->d_stack);
    }
    CK(cudaEventRecord(c->evB, c->stream));
    CK(cudaGetLastError());
    if (up) CK(cudaMemcpyAsync(up, c->d_up, ns * ncell * 4, cudaMemcpyDeviceToHost, c->stream));
    if (down) CK(cudaMemcpyAsync(down, c->d_down, ns * ncell * 4, cudaMemcpyDeviceToHost, c->stream));
    if (stable) CK(cudaMemcpyAsync(stable, c->d_stable, ns * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    const double steps = (double)(G.NT - 2) * ns;
    c->stats.cell_updates += steps * ((double)G.NZ * G.NX + (store ? 0.0 : (double)ncell));
    c->stats.algorithmic_bytes += steps * (double)G.NZ * G.NX * ((G.iCompen == 1 ? 60.0 : 44.0) - (store ? 8.0 : 0.0));
    c->stats.backward_seconds += ms * 1e-3;
    {   // bytes of the schedule that ran (include/rtm_b200.h, rtm_stats)
        const double nsteps = G.NT - 2, ring = (double)G.NZ * G.NX - (double)ncell, acc = G.iCompen == 1 ? 32.0 : 16.0;
        double pair_cells = 0;
        if (pairs)
            for (auto& kc : c->classes) pair_cells += kc.stream_mode ? kc.stream_cells : (double)kc.n_b2 * kTX * Tile2<4>::TZ;
        const double pair_steps = pairs ? 2.0 * ((G.NT - 2) / 2) : 0.0;
        const double single = store ? 12.0 + 4.0 + acc : 24.0 + acc;   // store-all: receiver field + stored source slot
        const double strips = store ? 0.0 : 2.0 * 2.0 * G.nfdmax * ((double)G.mod_NX + G.mod_NZ) * 4;   // read + written into the ring
        c->stats.executed_bytes_backward += ns * (pair_steps * pair_cells * (16.0 + acc / 2) + (nsteps * (double)ncell - pair_steps * pair_cells) * single +
                                                  nsteps * (ring * 12.0 + strips)) + nsteps * (double)G.NZ * G.NX * 4.0;
        c->stats.pair_cell_steps_backward += ns * pair_steps * pair_cells;
    }
    CK(cudaEventElapsedTime(&ms, c->evA, c->evB));
    c->stats.device_seconds += ms * 1e-3;  // whole batch: init, both loops, image post, stack
    c->stats.kernel_launches += (c->nlaunch - nl0) + 4;
    c->stats.shots += ns;
    c->stack_shots += ns;
    return RTM_OK;
}

static int upload_traces(rtm_ctx* c, int ns, const float* seis)
{
    const Geo& G = c->G;
    CK(cudaMemcpyAsync(c->d_stage, seis, (size_t)ns * G.n * G.NT * 4, cudaMemcpyHostToDevice, c->stream));
    dim3 grid((G.NT + 31) / 32, (G.n + 31) / 32, ns);  // [n][NT] -> [NT][n]
    transpose_traces_kernel<<<grid, dim3(32, 8), 0, c->stream>>>(c->d_stage, c->d_traces, G.n, G.NT);
    CK(cudaGetLastError());
    return RTM_OK;
}

// Raw-rate traces -> d_traces: resampled on the device when NT1 != NT (kernel.cu:839-845), plain
// transpose otherwise (the reference copies in that case whatever the two rates are).
static int upload_raw_traces(rtm_ctx* c, int ns, const float* raw, int NT1, float tao1)
{
    const Geo& G = c->G;
    if (NT1 == G.NT) return upload_traces(c, ns, raw);
    const size_t need = (size_t)c->S * G.n * NT1;
    if (need > c->raw_floats) {
        cudaFree(c->d_raw);
        c->d_raw = nullptr;
        CK(cudaMalloc(&c->d_raw, need * 4));
        c->raw_floats = need;
    }
    if (!c->d_sinc) {
        int ns_, nt_;
        const float* tb = rtm::sinc_table(&ns_, &nt_);
        CK(cudaMalloc(&c->d_sinc, sizeof(float) * ns_ * nt_));
        CK(cudaMemcpyAsync(c->d_sinc, tb, sizeof(float) * ns_ * nt_, cudaMemcpyHostToDevice, c->stream));
    }
    CK(cudaMemcpyAsync(c->d_raw, raw, (size_t)ns * G.n * NT1 * 4, cudaMemcpyHostToDevice, c->stream));
    const float xouts = (float)(1.0 / tao1), xoutb = (float)(8.0 - 0.0f * xouts);
    dim3 grid((G.NT + 31) / 32, (G.n + 31) / 32, ns);
    resample_traces_kernel<<<grid, dim3(32, 8), 0, c->stream>>>(c->d_raw, c->d_traces, G.n, NT1, G.NT, c->p.tao, xouts,
                                                                 xoutb, c->d_sinc);
    CK(cudaGetLastError());
    return RTM_OK;
}

extern "C" int rtm_migrate_raw(rtm_ctx* c, int nshots, const int* r_u, const int* r_x, const float* seis_raw,
                               int NT1, float tao1, float* up, float* down, float* stable)
{
    if (!c || !r_u || !r_x || !seis_raw || nshots < 1 || NT1 < 1 || !(tao1 > 0))
        return rtm_fail(RTM_ERR_ARG, "rtm_migrate_raw: bad argument");
    if (!c->have_model || !c->have_op) return rtm_fail(RTM_ERR_STATE, "rtm_migrate_raw: set the model and the operator first");
    CK(cudaSetDevice(c->device));
    const Geo& G = c->G;
    const size_t ncell = (size_t)G.mod_NX * G.mod_NZ;
    for (int first = 0; first < nshots; first += c->S) {
        const int ns = std::min(c->S, nshots - first);
        if (int rc = upload_raw_traces(c, ns, seis_raw + (size_t)first * G.n * NT1, NT1, tao1)) return rc;
        if (int rc = migrate_batch(c, ns, r_u + first, r_x + first, up ? up + first * ncell : nullptr,
                                   down ? down + first * ncell : nullptr, stable ? stable + first : nullptr))
            return rc;
    }
    return RTM_OK;
}

// Device resampling on its own (parity tests): traces host [ntr][NT1] -> host [ntr][NT_out].
extern "C" int rtm_resample_device(rtm_ctx* c, int ntr, const float* in, int NT1, float tao1, float* out)
{
    if (!c || !in || !out || ntr < 1) return rtm_fail(RTM_ERR_ARG, "rtm_resample_device: bad argument");
    CK(cudaSetDevice(c->device));
    const Geo& G = c->G;
    if (ntr > c->S * G.n) return rtm_fail(RTM_ERR_ARG, "rtm_resample_device: at most max_batch*n = %d traces", c->S * G.n);
    // reuse the gather path: treat the traces as one shot of `ntr` traces
    Geo saved = c->G;
    const int S_saved = c->S;
    c->G.n = ntr; c->S = 1;
    int rc = (NT1 == G.NT) ? upload_traces(c, 1, in) : upload_raw_traces(c, 1, in, NT1, tao1);
    c->G = saved; c->S = S_saved;
    if (rc) return rc;
    dim3 grid((ntr + 31) / 32, (G.NT + 31) / 32, 1);  // [NT][ntr] -> [ntr][NT]
    transpose_traces_kernel<<<grid, dim3(32, 8), 0, c->stream>>>(c->d_traces, c->d_stage, G.NT, ntr);
    CK(cudaMemcpyAsync(out, c->d_stage, (size_t)ntr * G.NT * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return RTM_OK;
}

extern "C" int rtm_migrate(rtm_ctx* c, int nshots, const int* r_u, const int* r_x, const float* seis,
                           float* up, float* down, float* stable)
{
    if (!c || !r_u || !r_x || !seis || nshots < 1) return rtm_fail(RTM_ERR_ARG, "rtm_migrate: bad argument");
    if (!c->have_model || !c->have_op) return rtm_fail(RTM_ERR_STATE, "rtm_migrate: set the model and the operator first");
    CK(cudaSetDevice(c->device));
    const Geo& G = c->G;
    const size_t ncell = (size_t)G.mod_NX * G.mod_NZ;
    for (int first = 0; first < nshots; first += c->S) {
        const int ns = std::min(c->S, nshots - first);
        if (int rc = upload_traces(c, ns, seis + (size_t)first * G.n * G.NT)) return rc;
        if (int rc = migrate_batch(c, ns, r_u + first, r_x + first, up ? up + first * ncell : nullptr,
                                   down ? down + first * ncell : nullptr, stable ? stable + first : nullptr))
            return rc;
    }
    return RTM_OK;
}

extern "C" int rtm_upload_gathers(rtm_ctx* c, int nshots, const float* seis)
{
    if (!c || !seis || nshots < 1 || nshots > c->S) return rtm_fail(RTM_ERR_ARG, "rtm_upload_gathers: 1 <= nshots <= max_batch");
    CK(cudaSetDevice(c->device));
    if (int rc = upload_traces(c, nshots, seis)) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return RTM_OK;
}

extern "C" int rtm_migrate_resident(rtm_ctx* c, int nshots, const int* r_u, const int* r_x)
{
    if (!c || !r_u || !r_x || nshots < 1 || nshots > c->S) return rtm_fail(RTM_ERR_ARG, "rtm_migrate_resident: 1 <= nshots <= max_batch");
    if (!c->have_model || !c->have_op) return rtm_fail(RTM_ERR_STATE, "rtm_migrate_resident: set the model and the operator first");
    CK(cudaSetDevice(c->device));
    return migrate_batch(c, nshots, r_u, r_x, nullptr, nullptr, nullptr);
}

extern "C" int rtm_stack_reset(rtm_ctx* c)
{
    if (!c) return rtm_fail(RTM_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->d_stack, 0, 2 * (size_t)c->G.mod_NX * c->G.mod_NZ * 4, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->stack_shots = 0;
    return RTM_OK;
}
extern "C" int rtm_stack_get(rtm_ctx* c, float* up_sum, float* down_sum, int* nshots)
{
    if (!c) return rtm_fail(RTM_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    const size_t ncell = (size_t)c->G.mod_NX * c->G.mod_NZ;
    if (up_sum) CK(cudaMemcpyAsync(up_sum, c->d_stack, ncell * 4, cudaMemcpyDeviceToHost, c->stream));
    if (down_sum) CK(cudaMemcpyAsync(down_sum, c->d_stack + ncell, ncell * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (nshots) *nshots = c->stack_shots;
    return RTM_OK;
}
extern "C" int rtm_stack_device(rtm_ctx* c, void** dev_ptr, size_t* nfloats, int* nshots)
{
    if (!c || !dev_ptr) return rtm_fail(RTM_ERR_ARG, "null argument");
    *dev_ptr = c->d_stack;
    if (nfloats) *nfloats = 2 * (size_t)c->G.mod_NX * c->G.mod_NZ;
    if (nshots) *nshots = c->stack_shots;
    return RTM_OK;
}
extern "C" int rtm_stack_finalize(const float* up_sum, const float* down_sum, int nrec, int iNorm,
                                  size_t ncell, float* image, float* illum)
{
    if (!up_sum || !down_sum || !image || nrec < 1) return rtm_fail(RTM_ERR_ARG, "rtm_stack_finalize: bad argument");
    for (size_t i = 0; i < ncell; ++i) {  // kernel.cu:1042-1059
        float a = up_sum[i] / nrec, b = down_sum[i] / nrec;
        if (iNorm == 1) a = a / b;
        image[i] = a;
        if (illum) illum[i] = b;
    }
    return RTM_OK;
}

extern "C" int rtm_get_stats(rtm_ctx* c, rtm_stats* out)
{
    if (!c || !out) return rtm_fail(RTM_ERR_ARG, "null argument");
    *out = c->stats;
    return RTM_OK;
}
extern "C" int rtm_reset_stats(rtm_ctx* c)
{
    if (!c) return rtm_fail(RTM_ERR_ARG, "null context");
    c->stats = rtm_stats{};
    return RTM_OK;
}

// NCCL reduce of the per-GPU stacks for contexts living in one process: rtm_nccl.cpp.
// Fallback used when NCCL cannot be loaded or fails: peer copies over NVLink into the first
// context's GPU and a device-side add, in context order.
namespace {
__global__ void add_into_kernel(float* dst, const float* src, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __fadd_rn(dst[i], src[i]);
}
}  // namespace

// out: nfl floats on the first context's GPU; the contexts' own stacks are only read.
int rtm_stack_reduce_p2p(rtm_ctx** ctxs, int nctx, float* out)
{
    rtm_ctx* root = ctxs[0];
    const size_t n = 2 * (size_t)root->G.mod_NX * root->G.mod_NZ;
    CK(cudaSetDevice(root->device));
    float* tmp = nullptr;
    CK(cudaMalloc(&tmp, n * 4));
    auto fail = [&](int rc) { cudaFree(tmp); return rc; };
    if (cudaMemcpyAsync(out, root->d_stack, n * 4, cudaMemcpyDeviceToDevice, root->stream) != cudaSuccess)
        return fail(rtm_fail(RTM_ERR_CUDA, "rtm_stack_reduce: device copy failed"));
    for (int i = 1; i < nctx; ++i) {
        if (2 * (size_t)ctxs[i]->G.mod_NX * ctxs[i]->G.mod_NZ != n) return fail(rtm_fail(RTM_ERR_ARG, "rtm_stack_reduce: contexts differ in image size"));
        if (cudaMemcpyPeerAsync(tmp, root->device, ctxs[i]->d_stack, ctxs[i]->device, n * 4, root->stream) != cudaSuccess)
            return fail(rtm_fail(RTM_ERR_CUDA, "rtm_stack_reduce: peer copy from device %d failed", ctxs[i]->device));
        add_into_kernel<<<(unsigned)((n + 255) / 256), 256, 0, root->stream>>>(out, tmp, n);
    }
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(root->stream) != cudaSuccess)
        return fail(rtm_fail(RTM_ERR_CUDA, "rtm_stack_reduce: peer-copy reduce failed"));
    cudaFree(tmp);
    return RTM_OK;
}
