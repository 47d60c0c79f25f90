/* oracle/ref_host_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" entry points around the reference's host-only translation units so
 * tests can call the *real* reference functions (built into
 * oracle/_ref/libref_host.so by oracle/Makefile; the sources are compiled where
 * they lie under /root/reference and are never copied into this repository).
 *
 *   funMandC / order      /root/reference/LSMOrCon_rec_2D.cpp:22, :526
 *   velocity              /root/reference/GPU_velocity_real.cpp:6
 *   resample              /root/reference/Resample.cpp:193
 *   D2T / T2D             /root/reference/DisToTimeAndTimeToDis1D.cpp:115, :31
 *   phase_correction      /root/reference/phase_correction_ricker_decon.cpp:153
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
/* same include order as /root/reference/kernel.cu:7-13 (a later file's `pi` macro
 * would otherwise clash with an earlier file's identifiers) */
#include "phase_correction_ricker_decon.cpp"
#include "LSMOrCon_rec_2D.cpp"
#include "GPU_velocity_real.cpp"
#include "DisToTimeAndTimeToDis1D.cpp"
#include "segy.h"
#include "Resample.cpp"
#include "SGYWrite.cpp"

extern "C" {

/* Returns NC; fills M[nvel], Index[nvel+1]; *c_out is malloc'ed (free with ref_free). */
int ref_funMandC(int nthita, int nfdmax, int nfdmin, int nvel, double tao, double h, double df,
                 double eps, double fmax, double vmin, double vmax, double dv,
                 int *fdcoeneed, int *M, int *Index, float **c_out, double hzx)
{
    return funMandC(nthita, nfdmax, nfdmin, nvel, tao, h, df, eps, fmax, vmin, vmax, dv,
                    fdcoeneed, M, Index, 0, c_out, hzx);
}
void ref_free(void *p) { free(p); }

void ref_order(int N, float *c) { order(N, c); }

void ref_cal2dfdcoe_lsm(double *c, double r, double bb, int M, double hzx)
{
    CAL2DFDCOE_LSM(c, r, bb, M, hzx);
}

void ref_callenfd2d_ls(int nfre, int nfdmin, int nfdmax, double vel, double tao, double h,
                       double df, int nthita, double eps, int *lenfd, double hzx)
{
    double *hk = (double *)malloc(sizeof(double) * nfre);
    double *de = (double *)malloc(sizeof(double) * nfre);
    callenfd2d_ls(hk, de, nfre, nfdmin, nfdmax, vel, tao, h, df, nthita, eps, lenfd, 0, hzx);
    free(hk); free(de);
}

void ref_velocity(const char *path, float *v, float *v_2, float *r, float *r_1, float *r_2,
                  int NZ, int NX, int N2, float tao, float h, int choice)
{
    velocity((char *)path, v, v_2, r, r_1, r_2, NZ, NX, NZ * NX, N2, tao, h, choice);
}

void ref_resample(int nxin, float dxin, float *yin, int nxout, float dxout, float *yout)
{
    resample(nxin, dxin, yin, nxout, dxout, yout);
}

void ref_segy2trace(const char *buf, float *trace, int ns, int format) { segy2trace(buf, trace, ns, format); }
void ref_trace2segy(char *buf, const float *trace, int ns, int format) { trace2segy(buf, trace, ns, format); }
void ref_segy2head(const char *buf, int *words, int nk) { segy2head(buf, words, nk); }
void ref_head2segy(char *buf, const int *words, int nk) { head2segy(buf, words, nk); }
/* WriteSGY reads "SGY_Model.sgy" from the current directory (SGYWrite.cpp:14) */
void ref_WriteSGY(float *Data, int NX, int NT, int tao3, float *SX, float *SY, float RX, float RY, float *DSR, char *name)
{
    WriteSGY(Data, NX, NT, tao3, SX, SY, RX, RY, DSR, name);
}

int ref_D2T(const char *file, float *V, float *D, int Nx, int Nz, int st, int en, float dx, float dz, float dt)
{
    return D2T(file, V, D, Nx, Nz, st, en, dx, dz, dt);
}
int ref_T2D(const char *file, float *V, float *D, int Nx, int Nz, int Nz_V, int st, int en, float dx, float dz, float dt)
{
    return T2D(file, V, D, Nx, Nz, Nz_V, st, en, dx, dz, dt);
}
void ref_phase_correction(float *din, float *dout, int ntr, int nt, float angle) { phase_correction(din, dout, ntr, nt, angle); }

} /* extern "C" */
