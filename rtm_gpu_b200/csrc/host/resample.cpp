// Trace resampling tao1 -> tao (reference: Resample.cpp:69-225, used when NT != NT1,
// kernel.cu:839-845).  The algorithm is the classic 8-point tabulated-sinc interpolation:
// 513 fractional shifts, each an 8-tap least-squares approximation of the band-limited sinc
// (band edge 0.066 + 0.265 ln 8 of Nyquist) obtained from a symmetric Toeplitz solve
// (Levinson recursion).  toeplitz_solve (= stoepd, :130-163), the table build (= mksinc) and
// resample_trace (= intt8r) are TRANSLITERATIONS for bit-identity: the float/double evaluation
// order of the classic routines is kept statement by statement so that resampled traces are
// bit-identical to the reference's (tests/test_host.py::test_resample_matches_reference); the
// device version (resample_traces_kernel, rtm_engine.cu) is new.
#include "rtm_host.h"

#include <cmath>
#include <mutex>

namespace rtm {
namespace {

constexpr int    kTaps   = 8;
constexpr int    kShifts = 513;
constexpr double kPi     = 3.1415926535898;

double sinc_pi(double x) { return x == 0.0 ? 1.0 : std::sin(kPi * x) / (kPi * x); }

// Solve the symmetric Toeplitz system R f = g (R from its first row r) by Levinson
// recursion; a[] is the prediction-error filter workspace.  Transliterated for bit-identity from
// stoepd (Resample.cpp:130-163).
void toeplitz_solve(int n, const double* r, const double* g, double* f, double* a)
{
    if (r[0] == 0.0) return;
    a[0] = 1.0;
    double v = r[0];
    f[0] = g[0] / r[0];
    for (int j = 1; j < n; ++j) {
        a[j] = 0.0;
        f[j] = 0.0;
        double e = 0.0;
        for (int i = 0; i < j; ++i) e += a[i] * r[j - i];
        double c = e / v;
        v -= c * e;
        for (int i = 0; i <= j / 2; ++i) {
            const double bot = a[j - i] - c * a[i];
            a[i] -= c * a[j - i];
            a[j - i] = bot;
        }
        double w = 0.0;
        for (int i = 0; i < j; ++i) w += f[i] * r[j - i];
        c = (w - g[j]) / v;
        for (int i = 0; i <= j; ++i) f[i] -= c * a[j - i];
    }
}

struct SincTable {
    float t[kShifts][kTaps];
    SincTable()
    {
        double fmax = 0.066 + 0.265 * std::log((double)kTaps);  // mksinc :175-192
        fmax = (fmax < 1.0) ? fmax : 1.0;
        for (int s = 1; s < kShifts - 1; ++s) {
            const float d = (float)s / (float)(kShifts - 1);
            double r[20], g[20], f[20], work[20];
            for (int j = 0; j < kTaps; ++j) {
                r[j] = sinc_pi(fmax * j);
                g[j] = sinc_pi(fmax * (kTaps / 2 - j - 1 + d));
            }
            toeplitz_solve(kTaps, r, g, f, work);
            for (int j = 0; j < kTaps; ++j) t[s][j] = (float)f[j];
        }
        for (int j = 0; j < kTaps; ++j) t[0][j] = t[kShifts - 1][j] = 0.0f;
        t[0][kTaps / 2 - 1]       = 1.0f;
        t[kShifts - 1][kTaps / 2] = 1.0f;
    }
};

const SincTable& table()
{
    static SincTable tb;  // thread-safe initialisation
    return tb;
}

}  // namespace

const float* sinc_table(int* nshifts, int* ntaps)
{
    if (nshifts) *nshifts = kShifts;
    if (ntaps) *ntaps = kTaps;
    return &table().t[0][0];
}

void resample_trace(int nxin, float dxin, const float* yin, int nxout, float dxout, float* yout)
{
    // intt8r :69-129 with fxin = 0 and zero extrapolation on both sides (resample :193-225)
    const SincTable& tb = table();
    const float fxin = 0.0f, yinl = 0.0f, yinr = 0.0f;
    const int   ioutb = -3 - 8;
    const float xouts = (float)(1.0 / dxin);
    const float xoutb = (float)(8.0 - fxin * xouts);
    const float ntm1  = (float)(kShifts - 1);
    for (int i = 0; i < nxout; ++i) {
        const float xout  = i * dxout;
        const float xoutn = xoutb + xout * xouts;
        const int   ix    = (int)xoutn;
        int         k     = ioutb + ix;
        const float frac  = xoutn - (float)ix;
        const int   kt    = (int)(frac >= 0.0 ? frac * ntm1 + 0.5 : (frac + 1.0) * ntm1 - 0.5);
        const float* w    = tb.t[kt];
        float sum = 0.0f;
        for (int j = 0; j < kTaps; ++j, ++k) {
            const float y = (k < 0) ? yinl : (k >= nxin ? yinr : yin[k]);
            const float p = y * w[j];
            sum = (j == 0) ? p : sum + p;
        }
        yout[i] = sum;
    }
}

}  // namespace rtm
