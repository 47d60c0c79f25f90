/* oracle/rtm_oracle.c -- TEST INFRASTRUCTURE ONLY; see rtm_oracle.h.
 *
 * Build with -ffp-contract=off (oracle/Makefile) so that every `a*b+c` below is
 * two roundings unless it goes through mad() with contract==1.
 *
 * FP contract of the reference (two variants, both restated here):
 *   contract==0  the C expressions of kernel.cu evaluated with one rounding per
 *                operator (float unless an operand is a double literal);
 *   contract==1  the same, except where nvcc 12.9 -arch=sm_100a fuses a multiply
 *                into the following add (FFMA).  The fused sites were read from
 *                `cuobjdump -sass oracle/_ref/ref_cuda`:
 *                  stencil   t = fma(s, hzx2_1, Px-), w1 = fma(c_l, u, w1)   (all 6 kernels)
 *                  float sum P2 = fma(a, w1, (P1+P1)-P0)      (Add, BKAdd, BKAdd_Con)
 *                  double sum: no fusion                      (Add_Con, BKAdd_EFF, BKAdd_EFF_Con)
 *                  Hybrid1   edge: rcp*fma(c2, D, fma(tv, A1, -B)); corner: rcp*fma(r1, Pa+Pb, P1)
 *                  Hybrid2   fma(1-w, P2, w*Pb)
 *                  Rel_*     fma(.,.,rel1), fma(S,S,rel2)
 */
#include "rtm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_PI 3.1415926535898 /* kernel.cu:16 (a double literal) */

static inline float mad(int contract, float a, float b, float c)
{
    if (contract) return fmaf(a, b, c);
    float t = a * b;
    return t + c;
}

/* ------------------------------------------------------------------ scalars */

float oracle_ricker(float t1, float f0) /* kernel.cu:1261-1266 */
{
    float  t00 = 1 / f0;
    double a   = ORACLE_PI * f0 * (t1 - t00); /* (pi*f0) in double times float difference */
    double a2  = a * a;                       /* pow(a,2) */
    double y   = (1 - 2 * a2) * exp(-a2);
    return (float)y;
}

void oracle_derived(float h, float hz, float tao, float tao1, float f0, int NT1, int *NT,
                    int *NT2, float *taoh, float *tao2, float *h2, float *taoh2, float *hzx2_1)
{
    /* kernel.cu:613-626 */
    float t2  = (float)((double)tao * (double)tao);       /* tao2=pow(tao,2)   */
    float hh2 = (float)(1 / ((double)h * (double)h));     /* h2=1/pow(h,2)     */
    float hzx = hz / h;
    if (NT) *NT = (int)((NT1 - 1) * tao1 / tao + 1.5);
    if (NT2) *NT2 = (int)(2.0 / (f0 * tao)) + 1;
    if (taoh) *taoh = tao / h;
    if (tao2) *tao2 = t2;
    if (h2) *h2 = hh2;
    if (taoh2) *taoh2 = t2 * hh2 / 2;
    if (hzx2_1) *hzx2_1 = 1 / (hzx * hzx);
}

void oracle_taylor(int M, float *c) /* LSMOrCon_rec_2D.cpp:526-551, order(2*M,c) */
{
    float *x = (float *)malloc(sizeof(float) * (M + 1));
    c[0]     = 0.0f;
    for (int i = 0; i <= M; i++) x[i] = 1.0f;
    for (int i = 1; i <= M; i++) {
        for (int j = 1; j <= M; j++) {
            if (j != i) {
                double j2 = (double)j * j, i2 = (double)i * i;
                x[i]      = (float)(x[i] * fabs(j2 / (j2 - i2)));
            }
        }
        double sgn = ((i + 1) % 2 == 0) ? 1.0 : -1.0; /* pow(-1,i+1) */
        c[i]       = (float)(sgn / ((double)i * i) * x[i]);
        c[0]       = c[0] - 2 * c[i];
    }
    free(x);
}

/* ----------------------------------------------------------------- velocity */

void oracle_pad_velocity(const float *vraw, int mod_NZ, int mod_NX, int N2, int ifv, float tao,
                         float h, float *v, float *r1)
{
    /* GPU_velocity_real.cpp:11-100: edge-replicate padding == clamp to the interior */
    int NZ = mod_NZ + 2 * N2, NX = mod_NX + 2 * N2;
    for (int i = 0; i < NZ; i++) {
        int zi = i - N2;
        if (zi < 0) zi = 0;
        if (zi > mod_NZ - 1) zi = mod_NZ - 1;
        for (int j = 0; j < NX; j++) {
            int xj = j - N2;
            if (xj < 0) xj = 0;
            if (xj > mod_NX - 1) xj = mod_NX - 1;
            v[(size_t)i * NX + j] = vraw[(size_t)xj * mod_NZ + zi];
        }
    }
    if (ifv == 1) { /* :84-100 mirror in x */
        for (int i = 0; i < NZ; i++)
            for (int j = 0; j < NX / 2; j++) {
                float a                        = v[(size_t)i * NX + j];
                v[(size_t)i * NX + j]          = v[(size_t)i * NX + NX - 1 - j];
                v[(size_t)i * NX + NX - 1 - j] = a;
            }
    }
    if (r1) { /* :104-117 */
        for (size_t i = 0; i < (size_t)NZ * NX; i++) {
            float r  = v[i] * tao / h;
            float r2 = (float)(((double)r * (double)r) / 2);
            r1[i]    = sqrtf(r2);
        }
    }
}

int oracle_velocity_bins(const float *v, long ncell, float dv, float *vmin_out, float *vmax_out,
                         int *need, int need_cap)
{
    /* kernel.cu:704-738 */
    float vmin = v[0], vmax = v[0], vel;
    for (long i = 0; i < ncell; i++) {
        if (v[i] < vmin) vmin = v[i];
        if (v[i] > vmax) vmax = v[i];
    }
    vel = ((int)(vmin / dv)) * dv;
    if (vel > vmin) vmin = vel - dv; else vmin = vel;
    vel = ((int)(vmax / dv)) * dv;
    if (vel < vmax) vmax = vel + dv; else vmax = vel;
    int nvel = (int)((vmax - vmin) / dv + 1.5);
    if (need) {
        for (int i = 0; i < nvel && i < need_cap; i++) need[i] = 0;
        for (long i = 0; i < ncell; i++) {
            int k = (int)((v[i] - vmin) / dv + 0.5);
            if (k >= 0 && k < need_cap) need[k] = 1;
        }
    }
    *vmin_out = vmin;
    *vmax_out = vmax;
    return nvel;
}

/* --------------------------------------------------------------- time loop */

typedef struct {
    int          NZ, NX, N2, mod_NZ, mod_NX, nfdmax, iLSTE, contract;
    float        tao2, h2, taoh, taoh2, hzx2_1, vmin, dv;
    const float *v, *c;
    const int   *Index;
    float        w[64 + 1]; /* blend weights, kernel.cu:688-691 */
    float       *r1;        /* corner coefficient (only ring diagonals are used) */
} octx;

struct oracle_strips {
    int    NT, mod_NZ, mod_NX, nfdmax;
    float *lf, *rt, *up, *dw;
};

oracle_strips *oracle_strips_alloc(const oracle_params *p)
{
    oracle_strips *s = (oracle_strips *)calloc(1, sizeof *s);
    s->NT = p->NT; s->mod_NZ = p->mod_NZ; s->mod_NX = p->mod_NX; s->nfdmax = p->nfdmax;
    size_t nz = (size_t)p->NT * p->mod_NZ * p->nfdmax, nx = (size_t)p->NT * p->mod_NX * p->nfdmax;
    s->lf = (float *)calloc(nz, sizeof(float));
    s->rt = (float *)calloc(nz, sizeof(float));
    s->up = (float *)calloc(nx, sizeof(float));
    s->dw = (float *)calloc(nx, sizeof(float));
    return s;
}
void oracle_strips_free(oracle_strips *s)
{
    if (!s) return;
    free(s->lf); free(s->rt); free(s->up); free(s->dw); free(s);
}

static void octx_init(octx *g, const oracle_params *p, const float *v, const float *c,
                      const int *Index)
{
    memset(g, 0, sizeof *g);
    g->N2 = p->N2; g->mod_NZ = p->mod_NZ; g->mod_NX = p->mod_NX;
    g->NZ = p->mod_NZ + 2 * p->N2; g->NX = p->mod_NX + 2 * p->N2;
    g->nfdmax = p->nfdmax; g->iLSTE = p->iLSTE; g->contract = p->contract;
    g->vmin = p->vmin; g->dv = p->dv; g->v = v; g->c = c; g->Index = Index;
    oracle_derived(p->h, p->hz, p->tao, p->tao, p->f0, 2, 0, 0, &g->taoh, &g->tao2, &g->h2,
                   &g->taoh2, &g->hzx2_1);
    for (int i = 0; i <= p->N2 && i <= 64; i++) g->w[i] = (float)((1.0 * i) / (1.0 * p->N2));
    size_t n = (size_t)g->NZ * g->NX;
    g->r1    = (float *)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; i++) { /* GPU_velocity_real.cpp:104-117 */
        float r  = v[i] * p->tao / p->h;
        float r2 = (float)(((double)r * (double)r) / 2);
        g->r1[i] = sqrtf(r2);
    }
}
static void octx_done(octx *g) { free(g->r1); }

enum { SUM_FLOAT = 0, SUM_DOUBLE = 1 };

/* Two-way update of one cell.  Add :46-80, Add_Con :82-114, BKAdd_EFF :246-283,
 * BKAdd_EFF_Con :286-320, BKAdd :339-380, BKAdd_Con :381-418.
 * sum_kind: which of the kernels' final sums (float `2*` or double `2.0*`). */
static float two_way(const octx *g, const float *P1, const float *P0, int z, int x, int sum_kind)
{
    const int    NZ = g->NZ, NX = g->NX, ct = g->contract;
    const size_t o  = (size_t)z * NX + x;
    const float  vv = g->v[o];
    const float *c;
    int          M;
    if (g->iLSTE == 0) {
        int top = g->Index[(int)((vv - g->vmin) / g->dv + 0.5)];
        int end = g->Index[(int)((vv - g->vmin) / g->dv + 1.5)];
        c = g->c + top;
        M = end - top - 1;
    } else {
        c = g->c;
        M = g->nfdmax;
    }
    float w1 = (float)((1.0 + g->hzx2_1) * c[0] * P1[o]); /* double: ((1+hzx2_1)*c0)*P1 */
    for (int l = 1; l <= M; l++) {
        int z1 = z - l, z2 = z + l, x1 = x - l, x2 = x + l;
        if (z1 < 0) z1 = -z1;
        if (z2 >= NZ) z2 = 2 * NZ - 2 - z2;
        if (x1 < 0) x1 = -x1;
        if (x2 >= NX) x2 = 2 * NX - 2 - x2;
        float s = P1[(size_t)z1 * NX + x] + P1[(size_t)z2 * NX + x];
        float t = mad(ct, s, g->hzx2_1, P1[(size_t)z * NX + x1]);
        float u = t + P1[(size_t)z * NX + x2];
        w1      = mad(ct, c[l], u, w1);
    }
    float a = vv * vv * g->tao2 * g->h2;
    if (sum_kind == SUM_FLOAT) {
        float base = 2 * P1[o] - P0[o];
        return mad(ct, a, w1, base);
    }
    float aw = a * w1;
    return (float)(2.0 * P1[o] - P0[o] + aw);
}

/* Hybrid one-way solution Pb of ring cell (z,x) (Hybrid1 :116-158 == BKHybrid1 :421-463)
 * followed by the blend (Hybrid2 :160-183 == BKHybrid2 :464-487).
 * P2 holds the UNBLENDED two-way solution everywhere (the reference blends in a later
 * launch); the blended value is returned. */
static float hybrid_cell(const octx *g, const float *P0, const float *P1, const float *P2,
                         int z, int x)
{
    const int NZ = g->NZ, NX = g->NX, N2 = g->N2, ct = g->contract;
    int dz = z < NZ - 1 - z ? z : NZ - 1 - z;
    int dx = x < NX - 1 - x ? x : NX - 1 - x;
    int a  = dz < dx ? dz : dx;      /* distance of this ring layer from the array edge */
    int l  = N2 - a;                 /* layer number 1..N2, weight w[l]                   */
    int sz = (z < NZ - 1 - z) ? 1 : -1; /* step towards the interior                       */
    int sx = (x < NX - 1 - x) ? 1 : -1;
#define AT(P, zz, xx) (P)[(size_t)(zz) * NX + (xx)]
    float Pb;
    int   dd = dz > dx ? dz - dx : dx - dz;
    if (dd <= 1) {
        /* corner cells (three per corner per layer), :138-155 */
        float r1  = AT(g->r1, z, x);
        float rcp = 1 / (2 * r1 + 1);
        float nb  = AT(P2, z, x + sx) + AT(P2, z + sz, x);
        Pb        = rcp * mad(ct, r1, nb, AT(P1, z, x));
    } else {
        int   iz, ix, tz, tx; /* inner neighbour and tangential unit step */
        float vq;             /* velocity used in the taoh2 term (mis-indexed in the reference) */
        if (dz < dx) {        /* top (:124) or bottom (:132) edge: tangent along x */
            iz = z + sz; ix = x; tz = 0; tx = 1;
            vq = AT(g->v, a, x);
        } else {              /* left (:128) or right (:136) edge: tangent along z */
            iz = z; ix = x + sx; tz = 1; tx = 0;
            vq = g->v[(size_t)a * NX + z]; /* flat index (N2-l)*NX + row */
        }
        float vb  = AT(g->v, z, x);
        float tv  = g->taoh * vb;
        float rcp = 1 / (tv + 1);
        float A1  = AT(P2, iz, ix) - AT(P0, iz, ix) + AT(P0, z, x);
        float B   = -2 * AT(P1, z, x) + AT(P0, z, x) + AT(P2, iz, ix) - 2 * AT(P1, iz, ix) +
                  AT(P0, iz, ix);
        float D = AT(P2, iz + tz, ix + tx) - 2 * AT(P2, iz, ix) + AT(P2, iz - tz, ix - tx) +
                  AT(P0, z + tz, x + tx) - 2 * AT(P0, z, x) + AT(P0, z - tz, x - tx);
        float c2 = g->taoh2 * vq * vq;
        Pb       = rcp * mad(ct, c2, D, mad(ct, tv, A1, -B));
    }
#undef AT
    float w  = g->w[l];
    float wb = w * Pb;
    return mad(ct, 1 - w, P2[(size_t)z * NX + x], wb);
}

/* Hybrid1 + Hybrid2 on the whole ring, in place on P2 (uses a scratch copy of the ring). */
static void hybrid_abc(const octx *g, const float *P0, const float *P1, float *P2, float *scratch)
{
    const int NZ = g->NZ, NX = g->NX, N2 = g->N2;
    for (int z = 0; z < NZ; z++)
        for (int x = 0; x < NX; x++) {
            int ring = z < N2 || z >= NZ - N2 || x < N2 || x >= NX - N2;
            if (ring) scratch[(size_t)z * NX + x] = hybrid_cell(g, P0, P1, P2, z, x);
        }
    for (int z = 0; z < NZ; z++)
        for (int x = 0; x < NX; x++) {
            int ring = z < N2 || z >= NZ - N2 || x < N2 || x >= NX - N2;
            if (ring) P2[(size_t)z * NX + x] = scratch[(size_t)z * NX + x];
        }
}

/* Equal :18-45 / Hybrid3 :184-208: save the strips (width nfdmax, just outside the
 * interior) of field P as time slot k. */
static void strips_save(const octx *g, oracle_strips *s, const float *P, int k)
{
    const int NX = g->NX, NZ = g->NZ, N2 = g->N2, nf = g->nfdmax, mz = g->mod_NZ, mx = g->mod_NX;
    size_t    oz = (size_t)k * mz * nf, ox = (size_t)k * mx * nf;
    for (int x = 0; x < mz; x++)
        for (int y = 0; y < nf; y++) {
            s->lf[oz + (size_t)x * nf + y] = P[(size_t)(N2 + x) * NX + N2 - y - 1];
            s->rt[oz + (size_t)x * nf + y] = P[(size_t)(N2 + x) * NX + NX - N2 + y];
        }
    for (int x = 0; x < nf; x++)
        for (int y = 0; y < mx; y++) {
            s->up[ox + (size_t)x * mx + y] = P[(size_t)(N2 - x - 1) * NX + y + N2];
            s->dw[ox + (size_t)x * mx + y] = P[(size_t)(NZ - N2 + x) * NX + y + N2];
        }
}
/* BKEqual :222-245 */
static void strips_restore(const octx *g, const oracle_strips *s, float *P, int k)
{
    const int NX = g->NX, NZ = g->NZ, N2 = g->N2, nf = g->nfdmax, mz = g->mod_NZ, mx = g->mod_NX;
    size_t    oz = (size_t)k * mz * nf, ox = (size_t)k * mx * nf;
    for (int x = 0; x < mz; x++)
        for (int y = 0; y < nf; y++) {
            P[(size_t)(N2 + x) * NX + N2 - y - 1]  = s->lf[oz + (size_t)x * nf + y];
            P[(size_t)(N2 + x) * NX + NX - N2 + y] = s->rt[oz + (size_t)x * nf + y];
        }
    for (int x = 0; x < nf; x++)
        for (int y = 0; y < mx; y++) {
            P[(size_t)(N2 - x - 1) * NX + y + N2]  = s->up[ox + (size_t)x * mx + y];
            P[(size_t)(NZ - N2 + x) * NX + y + N2] = s->dw[ox + (size_t)x * mx + y];
        }
}

static void forward_impl(const oracle_params *p, const octx *g, int r_u, int r_x, float *gather,
                         float *F0, float *F1, oracle_strips *strips, int nsnap,
                         const int *snap_k, float **snap_out)
{
    /* kernel.cu:798-821.  On return F1 = slot NT-1, F0 = slot NT-2. */
    const int    NZ = g->NZ, NX = g->NX, NT = p->NT;
    const size_t n  = (size_t)NZ * NX;
    int          NT2;
    oracle_derived(p->h, p->hz, p->tao, p->tao, p->f0, 2, 0, &NT2, 0, 0, 0, 0, 0);
    float *F2 = (float *)calloc(n, sizeof(float)), *scr = (float *)calloc(n, sizeof(float));
    memset(F0, 0, n * sizeof(float));
    memset(F1, 0, n * sizeof(float));
    F1[(size_t)r_u * NX + r_x] = (float)(oracle_ricker(0.0f, p->f0) / 2.0); /* :803 */
    if (strips) { strips_save(g, strips, F0, 0); strips_save(g, strips, F1, 1); } /* Equal */
    for (int s = 0; s < 2; s++) {
        const float *P = s ? F1 : F0;
        if (gather)
            for (int j = 0; j < p->n; j++)
                gather[(size_t)j * NT + s] = P[(size_t)p->s_z * NX + p->s_l + j * p->ds];
        for (int i = 0; i < nsnap; i++)
            if (snap_k[i] == s) memcpy(snap_out[i], P, n * sizeof(float));
    }
    for (int k = 2; k < NT; k++) {
        float wavelet = (k < NT2) ? oracle_ricker((k - 1) * p->tao, p->f0) : 0.0f; /* :812 */
        int   kind    = g->iLSTE == 0 ? SUM_FLOAT : SUM_DOUBLE; /* Add vs Add_Con */
        for (int z = 0; z < NZ; z++)
            for (int x = 0; x < NX; x++) F2[(size_t)z * NX + x] = two_way(g, F1, F0, z, x, kind);
        F2[(size_t)r_u * NX + r_x] += wavelet; /* :74-77 */
        hybrid_abc(g, F0, F1, F2, scr);        /* Hybrid1, Hybrid2 */
        if (strips) strips_save(g, strips, F2, k); /* Hybrid3 */
        if (gather)
            for (int j = 0; j < p->n; j++)
                gather[(size_t)j * NT + k] = F2[(size_t)p->s_z * NX + p->s_l + j * p->ds];
        for (int i = 0; i < nsnap; i++)
            if (snap_k[i] == k) memcpy(snap_out[i], F2, n * sizeof(float));
        memcpy(F0, F1, n * sizeof(float)); /* Deliver :210-221 */
        memcpy(F1, F2, n * sizeof(float));
    }
    free(F2);
    free(scr);
}

void oracle_forward(const oracle_params *p, const float *v, const float *c, const int *Index,
                    int r_u, int r_x, float *gather, float *last0, float *last1,
                    oracle_strips *strips, int nsnap, const int *snap_k, float **snap_out)
{
    octx g;
    octx_init(&g, p, v, c, Index);
    size_t n  = (size_t)g.NZ * g.NX;
    float *F0 = (float *)malloc(n * sizeof(float)), *F1 = (float *)malloc(n * sizeof(float));
    forward_impl(p, &g, r_u, r_x, gather, F0, F1, strips, nsnap, snap_k, snap_out);
    if (last0) memcpy(last0, F0, n * sizeof(float));
    if (last1) memcpy(last1, F1, n * sizeof(float));
    free(F0);
    free(F1);
    octx_done(&g);
}

void oracle_migrate_shot(const oracle_params *p, const float *v, const float *c,
                         const int *Index, int r_u, int r_x, const float *seis, float *up,
                         float *down, float *rel1_out, float *rel2_out, float *stable_out)
{
    oracle_migrate_shot_ex(p, v, c, Index, r_u, r_x, seis, up, down, rel1_out, rel2_out, stable_out, 0);
}

/* store_all != 0: NOT a reference mode.  The source field used by the imaging condition is the
 * stored forward field of slot k instead of its reverse-time reconstruction (the engine's
 * RTM_FLAG_STORE_ALL); everything else is unchanged. */
void oracle_migrate_shot_ex(const oracle_params *p, const float *v, const float *c,
                            const int *Index, int r_u, int r_x, const float *seis, float *up,
                            float *down, float *rel1_out, float *rel2_out, float *stable_out,
                            int store_all)
{
    octx g;
    octx_init(&g, p, v, c, Index);
    const int    NZ = g.NZ, NX = g.NX, N2 = g.N2, NT = p->NT, ct = g.contract;
    const int    mz = g.mod_NZ, mx = g.mod_NX;
    const size_t n  = (size_t)NZ * NX;
    int          NT2;
    oracle_derived(p->h, p->hz, p->tao, p->tao, p->f0, 2, 0, &NT2, 0, 0, 0, 0, 0);

    oracle_strips *st = oracle_strips_alloc(p);
    float *A = (float *)malloc(n * sizeof(float)), *B = (float *)malloc(n * sizeof(float));
    float **stored = 0;
    int    *stored_k = 0;
    if (store_all) {
        stored   = (float **)malloc(sizeof(float *) * NT);
        stored_k = (int *)malloc(sizeof(int) * NT);
        for (int k = 0; k < NT; k++) { stored[k] = (float *)malloc(n * sizeof(float)); stored_k[k] = k; }
    }
    forward_impl(p, &g, r_u, r_x, 0, A, B, st, store_all ? NT : 0, stored_k, stored); /* B = slot NT-1, A = slot NT-2 */

    /* hand-off :822-825: BW0 = slot NT-1, BW1 = slot NT-2 (current) */
    float *S0 = B, *S1 = A, *S2 = (float *)calloc(n, sizeof(float));
    float *R0 = (float *)calloc(n, sizeof(float)), *R1 = (float *)calloc(n, sizeof(float));
    float *R2 = (float *)calloc(n, sizeof(float)), *scr = (float *)calloc(n, sizeof(float));
    float *sumS = (float *)calloc(n, sizeof(float)), *sumR = (float *)calloc(n, sizeof(float));
    float *rel1 = (float *)calloc(n, sizeof(float)), *rel2 = (float *)calloc(n, sizeof(float));

    /* accumulator start values :859-876 (host arithmetic: never fused).  FW0/FW1 are the
     * forward INITIAL-condition arrays: zero, and f(0)/2 at the source cell. */
    float fw1src = (float)(oracle_ricker(0.0f, p->f0) / 2.0);
    for (size_t i = 0; i < n; i++) {
        float FW0 = 0.0f, FW1 = (i == (size_t)r_u * NX + r_x) ? fw1src : 0.0f;
        float BW0 = S0[i], BW1 = S1[i];
        if (p->iCompen == 1) {
            sumS[i]  = BW0 + BW1;
            sumR[i]  = FW0 + FW1;
            float t1 = sumR[i] * sumS[i], t2 = FW0 * BW0;
            rel1[i]  = t1 + t2;
        } else {
            float t1 = FW1 * BW1, t2 = FW0 * BW0;
            rel1[i]  = t1 + t2;
        }
        float q1 = BW1 * BW1, q2 = BW0 * BW0;
        rel2[i]  = q1 + q2;
    }

    for (int k = NT - 3; k >= 0; k--) { /* :887-931 */
        float wavelet = (k < NT2) ? oracle_ricker((k + 1) * p->tao, p->f0) : 0.0f;
        strips_restore(&g, st, S1, k + 1); /* BKEqual: slot k+1 */
        for (int z = N2; z < NZ - N2; z++) /* BKAdd_EFF / _Con: interior, double sum */
            for (int x = N2; x < NX - N2; x++) {
                float val = two_way(&g, S1, S0, z, x, SUM_DOUBLE);
                if (z == r_u && x == r_x) val += wavelet;
                S2[(size_t)z * NX + x] = store_all ? stored[k][(size_t)z * NX + x] : val;
            }
        for (int z = N2; z < NZ - N2; z++) /* Deliver_EFF: interior only */
            for (int x = N2; x < NX - N2; x++) {
                size_t o = (size_t)z * NX + x;
                S0[o] = S1[o];
                S1[o] = S2[o];
            }
        for (int z = 0; z < NZ; z++) /* BKAdd / BKAdd_Con: full grid, float sum */
            for (int x = 0; x < NX; x++) {
                size_t o    = (size_t)z * NX + x;
                int    hit  = 0;
                if (z == p->s_z && x >= p->s_l && x <= p->s_l + (p->n - 1) * p->ds &&
                    (x - p->s_l) % p->ds == 0) {
                    float d = seis[(size_t)((x - p->s_l) / p->ds) * NT + (k + 1)];
                    if (d != 0) { R2[o] = d; hit = 1; } /* replacement, :349-353 */
                }
                if (!hit) R2[o] = two_way(&g, R1, R0, z, x, SUM_FLOAT);
            }
        hybrid_abc(&g, R0, R1, R2, scr); /* BKHybrid1/2 */
        memcpy(R0, R1, n * sizeof(float)); /* Deliver */
        memcpy(R1, R2, n * sizeof(float));
        for (int z = N2; z < NZ - N2; z++) /* Rel_* :489-517 (interior is all that is used) */
            for (int x = N2; x < NX - N2; x++) {
                size_t o = (size_t)z * NX + x;
                float  S = S2[o], R = R2[o];
                if (p->iCompen == 1) {
                    sumS[o] = sumS[o] + S;
                    sumR[o] = sumR[o] + R;
                    rel1[o] = mad(ct, sumR[o], sumS[o], rel1[o]);
                } else {
                    rel1[o] = mad(ct, R, S, rel1[o]);
                }
                rel2[o] = mad(ct, S, S, rel2[o]);
            }
    }

    /* per-shot image post-processing on the host, :935-990 */
    float vmax2 = p->vmax * p->vmax;
    for (int j = N2; j < NX - N2; j++)
        for (int i = N2; i < NZ - N2; i++) {
            int i1 = i - 1, i2 = i + 1, j1 = j - 1, j2 = j + 1;
            if (i1 < N2) i1 = 2 * N2 - i1;
            if (j1 < N2) j1 = 2 * N2 - j1;
            if (i2 >= NZ - N2) i2 = 2 * (NZ - N2 - 1) - i2;
            if (j2 >= NX - N2) j2 = 2 * (NX - N2 - 1) - j2;
            float lap = rel1[(size_t)i * NX + j2] + rel1[(size_t)i * NX + j1] +
                        rel1[(size_t)i2 * NX + j] + rel1[(size_t)i1 * NX + j] -
                        4 * rel1[(size_t)i * NX + j];
            float vv = v[(size_t)i * NX + j];
            if (up) up[(size_t)(j - N2) * mz + (i - N2)] = (float)(-1.0 * lap * vv * vv / vmax2);
        }
    float MIGmax = 0.0f;
    for (int j = N2; j < NX - N2; j++)
        for (int i = N2; i < NZ - N2; i++)
            if (fabsf(rel2[(size_t)i * NX + j]) > MIGmax) MIGmax = fabsf(rel2[(size_t)i * NX + j]);
    float stable = MIGmax * p->whitecoe;
    if (stable_out) *stable_out = stable;
    for (int j = N2; j < NX - N2; j++)
        for (int i = N2; i < NZ - N2; i++) {
            if (down) down[(size_t)(j - N2) * mz + (i - N2)] = rel2[(size_t)i * NX + j] + stable;
            if (rel1_out) rel1_out[(size_t)(i - N2) * mx + (j - N2)] = rel1[(size_t)i * NX + j];
            if (rel2_out) rel2_out[(size_t)(i - N2) * mx + (j - N2)] = rel2[(size_t)i * NX + j];
        }

    if (stored) { for (int k = 0; k < NT; k++) free(stored[k]); free(stored); free(stored_k); }
    free(A); free(B); free(S2); free(R0); free(R1); free(R2); free(scr);
    free(sumS); free(sumR); free(rel1); free(rel2);
    oracle_strips_free(st);
    octx_done(&g);
}

void oracle_stack(const float *const *ups, const float *const *downs, int nshot, int mod_NZ,
                  int mod_NX, int iNorm, float *out_up, float *out_down)
{
    /* kernel.cu:992-1059: float accumulation in shot order, /nrec, optional division */
    size_t n = (size_t)mod_NZ * mod_NX;
    for (size_t i = 0; i < n; i++) {
        float a = 0.0f, b = 0.0f;
        for (int m = 0; m < nshot; m++) { a += ups[m][i]; b += downs[m][i]; }
        a = a / nshot;
        b = b / nshot;
        if (iNorm == 1) a = a / b;
        out_up[i] = a;
        if (out_down) out_down[i] = b;
    }
}
