/* oracle/rtm_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's hot path (caixh90/RTM_GPU, kernel.cu):
 * the per-shot time loop (forward modelling with hybrid ABC and boundary-strip
 * saving, reverse-time source reconstruction, receiver back-propagation with
 * data replacement, imaging condition) plus the per-shot image post-processing
 * and the stack.  Every function cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library; the product (rtm_gpu_b200/) must never call it.
 *
 * Pinning (how this oracle is itself checked):
 *   - `contract == 0` reproduces C semantics with no FMA contraction and is
 *     checked BIT-FOR-BIT against the reference's own code executed on the host
 *     (oracle/_ref/ref_cpu, built from /root/reference by oracle/Makefile with
 *     -ffp-contract=off); its outputs are committed under tests/golden/.
 *   - `contract == 1` reproduces the FMA contraction pattern nvcc 12.9 applies
 *     to the unmodified reference for sm_100a (read off `cuobjdump -sass` of
 *     oracle/_ref/ref_cuda; SURVEY.md 3.5 and DESIGN.md "FP contract") and is
 *     checked on the B200 against oracle/_ref/ref_cuda (tests -m gpu).
 *   The reference ships no tests or golden vectors of its own.
 */
#ifndef RTM_ORACLE_H
#define RTM_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    /* grid (padded, 0-based)                      kernel.cu:607-628 */
    int   mod_NZ, mod_NX, N2, nfdmax;
    int   NT;            /* number of time slots                                     */
    int   iLSTE;         /* 0 = adaptive least-squares operator, 1 = fixed Taylor    */
    int   iCompen;       /* 1 = Rel_Compen imaging, 0 = Rel_NonCompen                */
    float h, hz, tao, f0;
    float vmin, dv;      /* snapped vmin and bin width   kernel.cu:715-721           */
    float vmax;          /* snapped vmax (image filter scaling) kernel.cu:718-720    */
    float whitecoe;
    /* acquisition, padded 0-based                 kernel.cu:607-610 */
    int   s_l, s_z, n, ds;
    int   contract;      /* 0: plain C rounding; 1: nvcc sm_100a FMA pattern         */
} oracle_params;

/* Ricker wavelet, kernel.cu:1261-1266 */
float oracle_ricker(float t1, float f0);

/* Derived scalars, kernel.cu:613-626.  Any pointer may be NULL. */
void oracle_derived(float h, float hz, float tao, float tao1, float f0, int NT1,
                    int *NT, int *NT2, float *taoh, float *tao2, float *h2, float *taoh2,
                    float *hzx2_1);

/* Velocity padding, GPU_velocity_real.cpp:6-118.
 * vraw: [mod_NX][mod_NZ] (x outer, z inner);  v, r1: [NZ][NX] */
void oracle_pad_velocity(const float *vraw, int mod_NZ, int mod_NX, int N2, int ifv,
                         float tao, float h, float *v, float *r1);

/* vmin/vmax snapping and bin usage, kernel.cu:704-738.  need: int[nvel] (may be NULL
 * on the first call to size it). Returns nvel. */
int oracle_velocity_bins(const float *v, long ncell, float dv, float *vmin, float *vmax,
                         int *need, int need_cap);

/* Taylor coefficients, LSMOrCon_rec_2D.cpp:526-551 (order(2*M, c)); c has M+1 entries */
void oracle_taylor(int M, float *c);

/* Forward modelling of one virtual source (kernel.cu:798-825).
 *  v [NZ][NX]; c/Index: packed operator table (Index unused when iLSTE==1)
 *  gather    : [n][NT] or NULL   (SURVEY 3.2 definition)
 *  last0,last1: slot NT-2 / slot NT-1 full grids [NZ][NX] or NULL
 *  strips    : opaque, from oracle_strips_alloc(), or NULL
 *  snap_k/snap_out: nsnap requested slots -> snap_out[i] ([NZ][NX] each) */
typedef struct oracle_strips oracle_strips;
oracle_strips *oracle_strips_alloc(const oracle_params *p);
void oracle_strips_free(oracle_strips *s);

void oracle_forward(const oracle_params *p, const float *v, const float *c, const int *Index,
                    int r_u, int r_x, float *gather, float *last0, float *last1,
                    oracle_strips *strips, int nsnap, const int *snap_k, float **snap_out);

/* Full migration of one virtual source (kernel.cu:798-990).
 *  seis: [n][NT] observed traces (already at the modelling sample rate)
 *  up, down: [mod_NX][mod_NZ] (x outer, z inner) = RVSP_RTM_up_/down_ file layout
 *  rel1, rel2: raw accumulators, interior [mod_NZ][mod_NX], may be NULL
 *  stable: whitening constant printed by the reference (kernel.cu:971-972) */
void oracle_migrate_shot(const oracle_params *p, const float *v, const float *c,
                         const int *Index, int r_u, int r_x, const float *seis,
                         float *up, float *down, float *rel1, float *rel2, float *stable);

/* Same with store_all != 0: the engine's non-reference RTM_FLAG_STORE_ALL mode (imaging uses the
 * stored forward field instead of its reconstruction). */
void oracle_migrate_shot_ex(const oracle_params *p, const float *v, const float *c,
                            const int *Index, int r_u, int r_x, const float *seis, float *up,
                            float *down, float *rel1, float *rel2, float *stable, int store_all);

/* Stack over shots, kernel.cu:992-1059.  ups/downs: nshot images [mod_NX][mod_NZ];
 * out: [mod_NX][mod_NZ];  returns stacked up (optionally / stacked down if iNorm). */
void oracle_stack(const float *const *ups, const float *const *downs, int nshot,
                  int mod_NZ, int mod_NX, int iNorm, float *out_up, float *out_down);

#ifdef __cplusplus
}
#endif
#endif
