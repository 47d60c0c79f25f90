/* rtm_b200.h -- C ABI of the B200-native 2D acoustic RTM engine.
 *
 * The reference (caixh90/RTM_GPU) has no library boundary: `int main()` in kernel.cu
 * launches 17 CUDA kernels on raw device pointers.  This header is
 * the seam that replaces those launch sites and the host code around them; every entry
 * point names the reference lines it stands in for.  Plain C types only, no CUDA or
 * torch types; every function returns 0 on success or a negative rtm_status, and
 * rtm_last_error() returns the message of the calling thread's last failure.
 *
 * Threading: one context per GPU, each context driven by one host thread at a time
 * (the reference's serial shot loop, kernel.cu:791, becomes one thread per GPU, each
 * migrating batches of shots).  Host buffers are owned by the caller, device buffers by
 * the context.
 *
 * Index conventions are the reference's after kernel.cu:607-612: padded grid
 * NZ = mod_NZ + 2*N2 rows (z) by NX = mod_NX + 2*N2 columns (x), 0-based, x fastest.
 */
#ifndef RTM_B200_H
#define RTM_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RTM_OK             = 0,
    RTM_ERR_ARG        = -1, /* bad argument / inconsistent configuration */
    RTM_ERR_CUDA       = -2, /* CUDA runtime or driver error (message has the call) */
    RTM_ERR_STATE      = -3, /* call order (e.g. migrate before set_model) */
    RTM_ERR_IO         = -4,
    RTM_ERR_NCCL       = -5,
    RTM_ERR_NO_DEVICE  = -6  /* no CUDA device: the engine has no CPU fallback */
} rtm_status;

const char *rtm_last_error(void);
const char *rtm_version(void);

/* ------------------------------------------------------------------ engine context */

typedef struct rtm_ctx rtm_ctx;

typedef struct {
    /* grid and operator (2D_Real_RVSP_RTM.txt / Parameter.txt, kernel.cu:544-604) */
    int   mod_NZ, mod_NX;   /* model size without the absorbing ring                 */
    int   N2;               /* hybrid ABC width                                      */
    int   nfdmax;           /* longest operator (strip width); nfdmax <= N2          */
    int   NT;               /* time slots per shot, kernel.cu:613                    */
    int   iLSTE;            /* 0 adaptive least-squares operator, 1 fixed Taylor     */
    int   iCompen;          /* 1 Rel_Compen imaging, 0 Rel_NonCompen                 */
    float h, hz, tao, f0;   /* dx, dz, dt, Ricker peak frequency                     */
    float whitecoe;         /* illumination whitening, kernel.cu:971                 */
    /* data (surface) positions, padded 0-based (kernel.cu:607-610) */
    int   s_l, s_z, n, ds;
    /* engine */
    int   max_batch;        /* shots advanced together per kernel launch (>=1)       */
    int   flags;            /* RTM_FLAG_*                                            */
} rtm_params;

#define RTM_FLAG_NONE 0
/* Keep the whole forward wavefield in HBM ([NT][batch][NZ][pitch] floats) when it fits, and
 * image with it instead of reconstructing the source field backwards in time; falls back to
 * boundary saving + reconstruction (the reference's scheme, the default) when it does not fit.
 * NOT a reference mode: images differ from the reference's at the level of its reconstruction
 * error (~1e-4 relative).  rtm_store_all_active() tells which was chosen. */
#define RTM_FLAG_STORE_ALL 1

/* Replaces cudaSetDevice + the 22 cudaMalloc calls (kernel.cu:527, 758-779).
 * Constraints the reference does not have (RTM_ERR_ARG with a message otherwise):
 *   1 <= nfdmax <= min(N2, 16)       the boundary strips lie inside the ring (kernel.cu:23-43)
 *   1 <= N2 <= 64                    and the shared-memory staging of one ring tile must fit the
 *                                    device's opt-in limit (227 KB): about N2 <= 55 with operators of
 *                                    length 16, N2 <= 60 with length 4
 *   mod_NZ, mod_NX > 2*N2 + 4        the ring tiles assume an interior between the two bands
 *   adaptive operator: at most 65535 velocity bins (16-bit per-cell bin array; use a larger dv) */
int  rtm_create(int device, const rtm_params *params, rtm_ctx **out);
void rtm_destroy(rtm_ctx *ctx); /* kernel.cu:1241-1256 */

/* Padded velocity [NZ][NX] (host), snapped vmin/vmax and bin width dv
 * (velocity(), GPU_velocity_real.cpp:6; kernel.cu:704-721, 781). */
int rtm_set_model(rtm_ctx *ctx, const float *v_padded, float vmin, float vmax, float dv);

/* Packed operator table (funMandC / order, kernel.cu:744-753, 785-786).
 * iLSTE==0: Index[nvel+1], c[NC] with c[Index[b]..Index[b+1]) the coefficients of bin b.
 * iLSTE==1: Index may be NULL, c[nfdmax+1]. */
int rtm_set_operator(rtm_ctx *ctx, const int *Index, int nvel, const float *c, int NC);

/* Forward modelling of `nshots` virtual sources at (r_u[i], r_x[i])
 * (kernel.cu:798-821: Equal, Add|Add_Con, Hybrid1-3, Deliver).
 *   gathers  host [nshots][n][NT] or NULL: gather[j][k] = slot_k[s_z][s_l + j*ds]
 *   snap_k   nsnap requested time slots; snaps host [nshots][nsnap][NZ][NX] or NULL */
int rtm_forward(rtm_ctx *ctx, int nshots, const int *r_u, const int *r_x, float *gathers,
                int nsnap, const int *snap_k, float *snaps);

/* Full migration of `nshots` virtual sources (the body of the shot loop,
 * kernel.cu:798-990): forward modelling with boundary-strip saving, reverse-time source
 * reconstruction, receiver back-propagation with data replacement, imaging, per-shot
 * Laplacian filter and whitening.  Shots are processed max_batch at a time.
 *   seis   host [nshots][n][NT] observed traces at the modelling sample rate
 *   up     host [nshots][mod_NX][mod_NZ] (RVSP_RTM_up_<m>.dat layout) or NULL
 *   down   host [nshots][mod_NX][mod_NZ] (RVSP_RTM_down_<m>.dat layout) or NULL
 *   stable host [nshots] whitening constants (kernel.cu:971-972) or NULL
 * Every shot is also added, in shot order, to the context's device-resident stack. */
int rtm_migrate(rtm_ctx *ctx, int nshots, const int *r_u, const int *r_x, const float *seis,
                float *up, float *down, float *stable);

/* Same from traces at their recording rate: seis_raw host [nshots][n][NT1] sampled at tao1.  When
 * NT1 != NT they are resampled to the modelling rate ON THE DEVICE (resample(), Resample.cpp:193-225,
 * fused with the transpose into the engine's time-major layout), bit-identically to the host
 * routine; when NT1 == NT they are used as they are (kernel.cu:839-856). */
int rtm_migrate_raw(rtm_ctx *ctx, int nshots, const int *r_u, const int *r_x, const float *seis_raw,
                    int NT1, float tao1, float *up, float *down, float *stable);
/* The device resampler on its own: in host [ntr][NT1] at tao1 -> out host [ntr][NT] at tao. */
int rtm_resample_device(rtm_ctx *ctx, int ntr, const float *in, int NT1, float tao1, float *out);

/* Same, with the observed traces already resident on the device in the engine's
 * time-major layout (see rtm_upload_gathers); used to time the hot path alone. */
int rtm_upload_gathers(rtm_ctx *ctx, int nshots, const float *seis);
int rtm_migrate_resident(rtm_ctx *ctx, int nshots, const int *r_u, const int *r_x);

/* Stack (kernel.cu:992-1059).  The context keeps sum_m up_m and sum_m down_m on the
 * device.  rtm_stack_get copies them to the host ([mod_NX][mod_NZ]); rtm_stack_device
 * exposes the device buffers (2 x mod_NX*mod_NZ floats, contiguous: up then down) so a
 * multi-process launcher can reduce them with its own communicator; rtm_stack_reduce
 * does the single NCCL reduce for contexts living in one process (one per GPU). */
int rtm_stack_reset(rtm_ctx *ctx);
int rtm_stack_get(rtm_ctx *ctx, float *up_sum, float *down_sum, int *nshots);
int rtm_stack_device(rtm_ctx *ctx, void **dev_ptr, size_t *nfloats, int *nshots);
int rtm_stack_reduce(rtm_ctx **ctxs, int nctx, float *up_sum, float *down_sum, int *nshots);
/* rtm_stack_reduce sums into a scratch buffer on the first context's GPU: the contexts' own stacks
 * are left as they are, so it can be called again (e.g. after more shots).  If libnccl cannot be
 * loaded, or ncclCommInitAll / ncclReduce fail, it falls back to NVLink peer copies + a device add.
 * Backend of the last call: "nccl", "p2p" or "none". */
const char *rtm_stack_reduce_backend(void);
/* Creates the in-process NCCL communicators for `devices` ahead of time (ncclCommInitAll takes seconds; they are cached
 * and reused by rtm_stack_reduce over the same devices).  Optional; a failure only means the reduce will try again. */
int rtm_stack_reduce_prepare(const int *devices, int n);
/* sum/nrec, optional up/down normalisation (kernel.cu:1042-1059); in/out [mod_NX][mod_NZ] */
int rtm_stack_finalize(const float *up_sum, const float *down_sum, int nrec, int iNorm,
                       size_t ncell, float *image, float *illum);

/* Counters since rtm_create / last reset. */
typedef struct {
    double cell_updates;     /* wavefield samples advanced one step                    */
    double device_seconds;   /* CUDA-event time of the time loops                      */
    double forward_seconds, backward_seconds;
    double algorithmic_bytes;/* SURVEY.md 8(d) byte model for the same work            */
    long   kernel_launches;
    long   shots;
    /* byte model of the schedule that really ran (two-step passes move fewer bytes than 8(d)):
     * cells advanced two slots per pass 8 B (forward) / 32 B (backward, 16 B without compensation)
     * per cell-step, cells stepped singly 12 B / 56 B (40 B), absorbing-ring cells 12 B + their
     * boundary-strip traffic, the velocity factor once per launch (it is shared by the batch) */
    double executed_bytes_forward, executed_bytes_backward;
    double pair_cell_steps_forward, pair_cell_steps_backward;   /* cell-steps advanced by two-step passes */
} rtm_stats;
int rtm_get_stats(rtm_ctx *ctx, rtm_stats *out);
int rtm_reset_stats(rtm_ctx *ctx);
int rtm_device_count(void);
/* Device memory a context will allocate, by the engine's own formulas: `fixed` bytes + max_batch x
 * `per_shot` bytes (wavefields, accumulators, boundary strips of a migration, traces, images).
 * NT1 = samples per raw trace for rtm_migrate_raw (0: unused).  Uses mod_NZ, mod_NX, N2, nfdmax, NT, n
 * of *p.  rtm_device_free_bytes reports what cudaMemGetInfo sees on `device`.  The drop-in driver
 * sizes its shot batches with the two (the reference has no counterpart: one shot at a time). */
int rtm_memory_estimate(const rtm_params *p, int NT1, size_t *fixed, size_t *per_shot);
int rtm_device_free_bytes(int device, size_t *free_bytes);
/* Page-locked host memory for the buffers handed to rtm_migrate / rtm_migrate_raw (optional: pageable
 * buffers work too, with slower copies).  The reference reads traces into plain malloc memory (:827-838). */
int  rtm_host_alloc_pinned(void **p, size_t bytes);
void rtm_host_free_pinned(void *p);
int rtm_store_all_active(rtm_ctx *ctx); /* 1 if RTM_FLAG_STORE_ALL was requested and fits */

/* ------------------------------------------------------------------ host-side pieces
 * (pure CPU; the reference's main() does these before/after the device loop) */

/* Ricker wavelet sample f(t1,f0), kernel.cu:1261-1266. */
float rtm_ricker(float t1, float f0);
/* r_u = abs(N/hz)+N2-1, kernel.cu:794-795. */
int   rtm_source_row(float depth_m, float hz, int N2);
/* Derived scalars, kernel.cu:613-626; any output may be NULL. */
void  rtm_derived(float h, float hz, float tao, float tao1, float f0, int NT1, int *NT, int *NT2,
                  float *taoh, float *tao2, float *h2, float *taoh2, float *hzx2_1);
/* velocity(): raw [mod_NX][mod_NZ] -> padded [NZ][NX], GPU_velocity_real.cpp:6-100. */
void  rtm_pad_velocity(const float *vraw, int mod_NZ, int mod_NX, int N2, int ifv, float *v);
/* vmin/vmax snapping and bin usage, kernel.cu:704-738.  need: int[need_cap] or NULL.
 * Returns nvel. */
int   rtm_velocity_bins(const float *v, long ncell, float dv, float *vmin, float *vmax, int *need,
                        int need_cap);
/* order(2*M,c), LSMOrCon_rec_2D.cpp:526-551; c[M+1]. */
void  rtm_taylor_operator(int M, float *c);
/* funMandC, LSMOrCon_rec_2D.cpp:22-73.  M[nvel], Index[nvel+1]; returns NC and writes at
 * most c_cap coefficients to c (call with c==NULL to size).  verbose!=0 prints the
 * reference's search log to stdout. */
int   rtm_ls_operator(int nthita, int nfdmax, int nfdmin, int nvel, float tao, float h, float df,
                      float eps, float fmax, float vmin, float dv, float hzx, const int *need,
                      int *M, int *Index, float *c, int c_cap, int verbose);
/* CAL2DFDCOE_LSM, LSMOrCon_rec_2D.cpp:236-287; c[M+1] doubles. */
void  rtm_ls_coefficients(double *c, double r, double bmax, int M, double hzx);

/* resample(), Resample.cpp:193-225: one trace from nxin samples at dxin to nxout at dxout. */
void  rtm_resample(int nxin, float dxin, const float *yin, int nxout, float dxout, float *yout);

/* SEG-Y (big-endian; sample formats 1 IBM, 2 int32, 3 int16, 5 IEEE).
 * rtm_segy_decode / rtm_segy_encode: segy2trace / trace2segy, segy.cpp:653-695.
 * rtm_segy_info / rtm_segy_read: a whole file as [ntr][ns] floats -- a SEG-Y velocity model is read
 * this way (trace = x position, sample = depth), giving the raw [mod_NX][mod_NZ] layout.
 * rtm_segy_write_image: WriteSGY, SGYWrite.cpp:3-55 (headers from a template file). */
void rtm_segy_decode(const unsigned char *buf, float *out, int ns, int format);
void rtm_segy_encode(unsigned char *buf, const float *in, int ns, int format);
int  rtm_segy_info(const char *path, int *ns, int *ntr, int *format, float *dt);
int  rtm_segy_read(const char *path, float *out, int ns, int ntr);
int  rtm_segy_write_image(const char *template_path, const char *out_path, const float *data, int ntr,
                          int ns, int dt_value, const float *SX, const float *SY, float RX, float RY,
                          const float *DSR);

/* Post-stack chain (kernel.cu:1110-1179; host-only): D2T / T2D (DisToTimeAndTimeToDis1D.cpp:115, :31)
 * and phase_correction (phase_correction_ricker_decon.cpp:153).  V, D are [Nx][Nz] (x outer); the
 * converters return the samples per output trace and write at most cap floats (call with NULL to size). */
int  rtm_depth_to_time(const float *V, const float *D, int Nx, int Nz, float dz, float dt, float *T, int cap);
int  rtm_time_to_depth(const float *V, const float *D, int Nx, int Nt, int Nz_V, float dtime, float ddepth,
                       float *Z, int cap);
void rtm_phase_rotate(const float *din, float *dout, int ntr, int nt, float angle_deg);

/* Drop-in driver: everything main() does up to the stacked image (kernel.cu:525-1108),
 * reading the reference's input files and writing its output files.
 *   run_file  path of 2D_Real_RVSP_RTM.txt
 *   ngpu      number of GPUs (one host thread each); <=0 means all visible
 *   batch     shots per launch per GPU; <=0 picks a default from the grid size */
int rtm_run_driver(const char *run_file, int ngpu, int batch, int verbose);

#ifdef __cplusplus
}
#endif
#endif
