// Post-stack chain of the reference's driver surface (kernel.cu:1110-1179): depth -> two-way time
// (D2T), constant phase rotation through a radix-2 FFT (phase_correction), time -> depth (T2D).
// Reference: DisToTimeAndTimeToDis1D.cpp:8-180, phase_correction_ricker_decon.cpp:153-223, :325-387.
// Host-only cosmetics after the stack; re-implemented so the driver can produce the reference's
// product files byte for byte (tests/test_host.py::test_poststack_*).  Single precision where the
// reference is single, double inside the FFT, same evaluation order.
#include "rtm_host.h"

#include <cmath>
#include <vector>

namespace rtm {
namespace {

// Piecewise-linear resampling of (x[i], y[i]) onto the regular grid j*d, j < nout
// (interpolate, DisToTimeAndTimeToDis1D.cpp:8-28).  out[] must be zero beyond index 0 on entry.
void regrid_linear(float* out, const float* x, const float* y, int nout, int nin, float d)
{
    int k1 = 0;
    out[0] = y[0];
    for (int i = 1; i < nin; ++i) {
        const int k2 = (int)(x[i] / d);
        for (int j = k1 + 1; j <= k2; ++j) {
            if (j == nout) continue;
            out[j] = y[i - 1] + (j * d - x[i - 1]) / (x[i] - x[i - 1]) * (y[i] - y[i - 1]);
        }
        k1 = k2;
    }
}

// Transliterated for bit-identity from kbfft (phase_correction_ricker_decon.cpp:325-387): the
// reference's twiddle recurrence (its literal 6.283185306, the 0.000001 test) and butterfly order
// fix every rounding of the rotated traces, so the routine keeps them statement by statement.
// In-place-style radix-2 FFT on n = 2^k points: (pr,pi) in, (fr,fi) out; pr/pi are used as the
// twiddle table afterwards.  inverse != 0: conjugate twiddles and 1/n scaling.  polar != 0:
// pr/pi are finally overwritten with amplitude/(n/2) and phase (kbfft :325-387; the caller of the
// phase rotation relies on that for the DC and Nyquist bins).
void fft_radix2(double* pr, double* pi, int n, int k, double* fr, double* fi, int inverse, int polar)
{
    for (int it = 0; it < n; ++it) {  // bit reversal
        int m = it, is = 0;
        for (int i = 0; i < k; ++i) { const int j = m / 2; is = 2 * is + (m - 2 * j); m = j; }
        fr[it] = pr[is];
        fi[it] = pi[is];
    }
    pr[0] = 1.0; pi[0] = 0.0;
    double p = 6.283185306 / (1.0 * n), q, s;
    pr[1] = std::cos(p); pi[1] = -std::sin(p);
    if (inverse) pi[1] = -pi[1];
    for (int i = 2; i < n; ++i) {  // w^i = w^(i-1) * w with three multiplications
        p = pr[i - 1] * pr[1]; q = pi[i - 1] * pi[1];
        s = (pr[i - 1] + pi[i - 1]) * (pr[1] + pi[1]);
        pr[i] = p - q; pi[i] = s - p - q;
    }
    for (int it = 0; it <= n - 2; it += 2) {  // first stage
        const double vr = fr[it], vi = fi[it];
        fr[it] = vr + fr[it + 1]; fi[it] = vi + fi[it + 1];
        fr[it + 1] = vr - fr[it + 1]; fi[it + 1] = vi - fi[it + 1];
    }
    int m = n / 2, nv = 2;
    for (int l0 = k - 2; l0 >= 0; --l0) {
        m /= 2; nv *= 2;
        for (int it = 0; it <= (m - 1) * nv; it += nv)
            for (int j = 0; j <= nv / 2 - 1; ++j) {
                const int a = it + j, b = it + j + nv / 2;
                p = pr[m * j] * fr[b];
                q = pi[m * j] * fi[b];
                s = pr[m * j] + pi[m * j];
                s = s * (fr[b] + fi[b]);
                const double oddr = p - q, oddi = s - p - q;
                fr[b] = fr[a] - oddr; fi[b] = fi[a] - oddi;
                fr[a] = fr[a] + oddr; fi[a] = fi[a] + oddi;
            }
    }
    if (inverse)
        for (int i = 0; i < n; ++i) { fr[i] = fr[i] / (1.0 * n); fi[i] = fi[i] / (1.0 * n); }
    if (polar)
        for (int i = 0; i < n; ++i) {
            pr[i] = std::sqrt(fr[i] * fr[i] + fi[i] * fi[i]);
            pr[i] = pr[i] / (n / 2);
            if (std::fabs(fr[i]) < 0.000001 * std::fabs(fi[i])) pi[i] = (fi[i] * fr[i]) > 0 ? 90.0 : -90.0;
            else pi[i] = std::atan(fi[i] / fr[i]);
        }
}

}  // namespace

int depth_to_time(const float* V, const float* D, int Nx, int Nz, float dz, float dt, std::vector<float>& T)
{
    // D2T :115-180 over all Nx columns: t(z) = sum 2 dz / v, then linear regridding to dt
    std::vector<float> t0((size_t)Nx * Nz, 0.0f);
    float tmax = 0;
    for (int i = 0; i < Nx; ++i) {
        float t = 0;
        for (int j = 1; j < Nz; ++j) {
            t += 2 * dz / V[(size_t)i * Nz + j];
            t0[(size_t)i * Nz + j] = t;
        }
        if (tmax < t) tmax = t;
    }
    const int Nt = (int)(tmax / dt);
    T.assign((size_t)Nx * (Nt > 0 ? Nt : 0), 0.0f);
    for (int i = 0; i < Nx && Nt > 0; ++i)
        regrid_linear(&T[(size_t)i * Nt], &t0[(size_t)i * Nz], &D[(size_t)i * Nz], Nt, Nz, dt);
    return Nt;
}

int time_to_depth(const float* V, const float* D, int Nx, int Nt_in, int Nz_V, float dtime, float ddepth,
                  std::vector<float>& Z)
{
    // T2D :31-113 (called with "dz = time step, dt = depth step"): z(t) = sum dt * v(z) / 2 with the
    // velocity looked up at the running depth, then linear regridding to the depth step
    std::vector<float> z0((size_t)Nx * Nt_in, 0.0f);
    float zmax = 0;
    for (int i = 0; i < Nx; ++i) {
        float z = 0;
        for (int j = 1; j < Nt_in; ++j) {
            int jz = (int)((int)z / ddepth);  // `(int)temp/dt`: the cast binds to temp
            if (jz >= Nz_V) jz = Nz_V - 1;
            z += dtime * V[(size_t)i * Nz_V + jz] / 2;
            z0[(size_t)i * Nt_in + j] = z;
        }
        if (zmax < z) zmax = z;
    }
    const int Nz = (int)(zmax / ddepth);
    Z.assign((size_t)Nx * (Nz > 0 ? Nz : 0), 0.0f);
    for (int i = 0; i < Nx && Nz > 0; ++i)
        regrid_linear(&Z[(size_t)i * Nz], &z0[(size_t)i * Nt_in], &D[(size_t)i * Nt_in], Nz, Nt_in, ddepth);
    return Nz;
}

void phase_rotate(const float* din, float* dout, int ntr, int nt, float angle)
{
    // phase_correction :153-223: zero-pad to 2^k, forward FFT, rotate bins 1..n/2-1 by `angle`
    // degrees, Hermitian mirror, inverse FFT.  Bins 0 and n/2 are NOT spectrum values at that
    // point: the forward call leaves amplitude/phase there (kept, it is what the reference does).
    const double theta = angle * 3.1415926535898 / 180;
    int n = 2, k = 1;
    while (n < nt) { n *= 2; ++k; }
    const double ct = std::cos(theta), st = std::sin(theta);
    const int nh = n / 2;
    std::vector<double> pr(n), pi(n), fr(n), fi(n);
    for (int itr = 0; itr < ntr; ++itr) {
        for (int i = 0; i < nt; ++i) { pr[i] = din[(size_t)itr * nt + i]; pi[i] = 0.0; }
        for (int i = nt; i < n; ++i) { pr[i] = 0.0; pi[i] = 0.0; }
        fft_radix2(pr.data(), pi.data(), n, k, fr.data(), fi.data(), 0, 1);
        for (int i = 1; i < nh; ++i) {
            pr[i] = fr[i] * ct - fi[i] * st;
            pi[i] = fi[i] * ct + fr[i] * st;
        }
        for (int i = n - 1; i > nh; --i) { pr[i] = pr[n - i]; pi[i] = -pi[n - i]; }
        fft_radix2(pr.data(), pi.data(), n, k, fr.data(), fi.data(), 1, 0);
        for (int i = 0; i < nt; ++i) dout[(size_t)itr * nt + i] = (float)fr[i];
    }
}

}  // namespace rtm
