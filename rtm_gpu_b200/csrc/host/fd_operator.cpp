// Finite-difference operator preparation (host, FP64): the adaptive "optimal"
// time-space-domain least-squares coefficients with a per-velocity operator
// length, and the fixed Taylor coefficients.
//
// Replaces the reference's LSMOrCon_rec_2D.cpp (funMandC :22-73, CAL2DFDCOE_LSM
// :236-287, fgaus/fgausf :139-233, Gauss :74-137, callenfd2d_ls :293-346,
// calfdlen_ls :347-524, order :526-551).  The search (bisection over velocity bins,
// threaded), the quadrature (ls_coefficients: the fgaus state machine as plain loops over cached basis values) and the
// table packing are written from the algorithm; solve_scaled_pivot is a
// TRANSLITERATION of the reference's Gauss routine (same pivoting decisions, same
// operation order -- any other elimination order changes the last bits of the
// coefficients).  The FP64 evaluation order of every sum is kept so that the float32
// tables that reach the device are bit-identical to the reference's (checked against
// goldens produced by the reference's own code, tests/test_host.py::test_ls_operator_golden).
#include "rtm_host.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <thread>
#include <vector>

namespace rtm {
namespace {

// Independent work items on all host cores.  Every item is computed exactly as in the serial
// code, so results do not depend on the number of threads.
template <class F> void parallel_for(int n, F f)
{
    const int nt = std::max(1, std::min<int>(n, (int)std::thread::hardware_concurrency()));
    if (nt == 1) { for (int i = 0; i < n; ++i) f(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&] { for (int i = next++; i < n; i = next++) f(i); });
    for (auto& th : pool) th.join();
}

constexpr double kPi = 3.1415926535898;  // the reference's literal (LSMOrCon_rec_2D.cpp:4)

// 5-point Gauss-Legendre rule on [-1,1] (10-digit table, as in the reference :143-146)
constexpr double kNode[5]   = {-0.9061798459, -0.5384693101, 0.0, 0.5384693101, 0.9061798459};
constexpr double kWeight[5] = {0.2369268851, 0.4786286705, 0.5688888889, 0.4786286705, 0.2369268851};
constexpr int    kPanels    = 4;  // js = {4,4}, :239

// One factor of the least-squares basis: the spatial dispersion term of the i-th
// coefficient divided by the temporal term, at angle theta and wavenumber*h = beta.
// (fgausf :216-233; `inv_r2` = pow(r,-2) from libm exactly as there.)
inline double basis(int i, double theta_c, double theta_s, double beta, double inv_r2,
                    double one_minus_car, double hzx, double hzx2)
{
    return ((1 + 1 / hzx2) - std::cos(i * beta * theta_c) - std::cos(hzx * i * beta * theta_s) / hzx2) /
           (inv_r2 * one_minus_car);
}

// Dense solve A x = b: rows scaled by their (signed) largest-magnitude entry, then
// Gaussian elimination with partial pivoting, then back substitution.  Transliterated for
// bit-identity from Gauss (LSMOrCon_rec_2D.cpp:74-137): statement order and pivot tests are the
// reference's, only the containers differ.
bool solve_scaled_pivot(std::vector<std::vector<double>>& A, std::vector<double>& b,
                        std::vector<double>& x)
{
    const int n = (int)b.size();
    for (int i = 0; i < n; ++i) {
        double big = A[i][0];
        for (int j = 0; j < n; ++j)
            if (std::fabs(A[i][j]) > std::fabs(big)) big = A[i][j];
        if (std::fabs(big) < 1e-10) return false;
        for (int j = 0; j < n; ++j) A[i][j] = A[i][j] / big;
        b[i] = b[i] / big;
    }
    for (int i = 0; i < n - 1; ++i) {
        int piv = i;
        for (int j = i; j < n; ++j)
            if (std::fabs(A[j][i]) > std::fabs(A[piv][i])) piv = j;
        if (piv != i) {
            std::swap(b[i], b[piv]);
            for (int j = i; j < n; ++j) std::swap(A[i][j], A[piv][j]);
        }
        for (int p = i + 1; p < n; ++p) {
            const double m = A[p][i] / A[i][i];
            b[p] = b[p] - m * b[i];
            for (int j = i; j < n; ++j) A[p][j] = A[p][j] - m * A[i][j];
        }
    }
    x[n - 1] = b[n - 1] / A[n - 1][n - 1];
    for (int i = n - 2; i >= 0; --i) {
        double m = 0.0;
        for (int j = i + 1; j < n; ++j) m = m + A[i][j] * x[j];
        x[i] = (b[i] - m) / A[i][i];
    }
    return true;
}

}  // namespace

// One least-squares system (CAL2DFDCOE_LSM :236-287).  Every matrix entry and right-hand side is a composite 2-D
// Gauss-Legendre integral over theta in [0, 2 pi] (outer) and beta in [0, bmax] (inner), 4 panels x 5 nodes per axis, of
//   f_i * f_j  (normal matrix entry)   or   f_i  (right-hand side).
// The accumulation order -- inner sum first, panel centres advanced by repeated addition of the panel width, inner
// half-width folded in when the inner integral is added to the outer sum, outer half-width applied last -- is that of
// fgaus :139-202.  The reference evaluates the basis anew for every (i, j) at the same 400 nodes: M (M + 3) / 2 integrals
// x 400 nodes x up to 2 basis values.  The values depend on (node, i) only, so they are computed once (M x 400) and the
// sums formed from them in the same order: same operands, same operation order, same bits
// (tests/test_host.py::test_ls_operator_golden), ~15x less libm work at M = 10.
void ls_coefficients(double* c, double r, double bmax, int M, double hzx)
{
    constexpr int NA = kPanels * 5, NB = kPanels * 5;   // nodes per axis
    const double hzx2   = hzx * hzx;
    const double inv_r2 = std::pow(r, -2);
    const double ha = 0.5 * (2 * kPi - 0.0) / kPanels;
    const double hb = 0.5 * (bmax - 0.0) / kPanels;
    std::vector<double> F((size_t)NA * NB * M);          // [node a][node b][i-1]
    {
        double ca = ha + 0.0;
        for (int pa = 0, na = 0; pa < kPanels; ++pa) {
            for (int ka = 0; ka < 5; ++ka, ++na) {
                const double theta = ha * kNode[ka] + ca;
                const double tc = std::cos(theta), ts = std::sin(theta);
                double cb = hb + 0.0;
                for (int pb = 0, nb = 0; pb < kPanels; ++pb) {
                    for (int kb = 0; kb < 5; ++kb, ++nb) {
                        const double beta = hb * kNode[kb] + cb;
                        const double omc  = 1 - std::cos(r * beta);
                        double* f = &F[((size_t)na * NB + nb) * M];
                        for (int i = 1; i <= M; ++i) f[i - 1] = basis(i, tc, ts, beta, inv_r2, omc, hzx, hzx2);
                    }
                    cb = cb + hb * 2.0;
                }
            }
            ca = ca + ha * 2.0;
        }
    }
    auto integral = [&](int i, int j, bool pair) {
        double outer = 0.0;
        for (int na = 0; na < NA; ++na) {
            double inner = 0.0;
            for (int nb = 0; nb < NB; ++nb) {
                const double* fv = &F[((size_t)na * NB + nb) * M];
                double f = fv[i - 1];
                if (pair) f = f * fv[j - 1];
                inner = f * kWeight[nb % 5] + inner;
            }
            outer = inner * hb * kWeight[na % 5] + outer;
        }
        return outer * ha;
    };
    std::vector<std::vector<double>> A(M, std::vector<double>(M));
    std::vector<double> rhs(M), x(M);
    for (int i = 0; i < M; ++i) {
        for (int j = i; j < M; ++j) A[i][j] = integral(i + 1, j + 1, true);
        rhs[i] = integral(i + 1, 0, false);
    }
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < i; ++j) A[i][j] = A[j][i];
    solve_scaled_pivot(A, rhs, x);
    double s = 0.0;
    for (int i = 1; i <= M; ++i) {
        c[i] = x[i - 1];
        s += c[i];
    }
    c[0] = -2.0 * s;
}

int operator_length(const OperatorSearch& q, double vel, int start_len, std::FILE* log)
{
    // callenfd2d_ls :293-346: smallest length in [start_len, nfdmax] whose relative
    // phase-velocity error stays <= eps for all frequencies k*df (k < nfre) and all
    // propagation angles in [0, pi/4]; nfdmax (with a message) if none does.
    const int    nfre = q.nfre;
    const double hzx2 = q.hzx * q.hzx;
    std::vector<double> hk(nfre), c(q.nfdmax + 1, 0.0);
    double b = 2.0 * kPi * q.df * q.h / vel;
    for (int i = 0; i < nfre; ++i) hk[i] = b * i;
    const double r  = vel * q.tao / q.h;
    b               = 2 * kPi * q.df * (nfre - 1) * q.tao / r;
    const double ra = q.h / vel;
    int len = 0, j;
    for (j = start_len; j <= q.nfdmax; ++j) {
        ls_coefficients(c.data(), r, b, j, q.hzx);
        // the serial loop stops at the first frequency whose error exceeds eps at some angle
        // n < nthita; whether such a frequency exists does not depend on the order, so all
        // frequencies are tested concurrently
        std::atomic<bool> failed(false);
        parallel_for(nfre - 1, [&](int kk) {
            if (failed.load(std::memory_order_relaxed)) return;
            const int k = kk + 1;
            const double rat = 2 / (r * hk[k]);
            int n;
            for (n = 0; n <= q.nthita; ++n) {
                const double nth = n * kPi / (4 * q.nthita);
                double mid = 0;
                for (int l = 1; l <= j; ++l) {
                    const double sz = std::sin(l * q.hzx * hk[k] * std::sin(nth) / 2);
                    const double sx = std::sin(l * hk[k] * std::cos(nth) / 2);
                    mid = mid + c[l] * (sz * sz / hzx2 + sx * sx);
                }
                double err = rat * std::asin(std::sqrt(r * r * mid));
                err        = std::fabs(ra * (1.0 / err - 1.0));
                if (err > q.eps) break;
            }
            if (n < q.nthita) failed = true;  // (a failure at exactly n == nthita is not caught, as in the reference)
        });
        len = j;
        if (!failed) break;
    }
    if (j == q.nfdmax + 1 && log) std::fprintf(log, "M=%d is not enough", j);
    return len;
}

bool operator_lengths(const OperatorSearch& q, int nvel, double vmin, double dv, int* len,
                      std::FILE* log)
{
    // calfdlen_ls :347-524.  The length is assumed non-increasing in velocity; the table
    // is filled from the high-velocity end backwards in strides `inc`, bisecting each
    // stride whose end points disagree.  The probe sequence (and the start length handed
    // to each probe) is the reference's, so the table is identical even where the
    // monotonicity assumption fails.
    auto vel_of = [&](int i) { return vmin + (i - 1) * dv; };  // 1-based bin
    auto fill   = [&](int a, int bnd, int l) { for (int i = a; i <= bnd; ++i) len[i - 1] = l; };
    int probes = 0, effort = 0;
    const int ibeg = 1, iend = nvel;

    int lbeg = operator_length(q, vel_of(ibeg), q.nfdmin, log);
    ++probes; effort += lbeg - 2 + 1;
    if (lbeg == 0) return false;
    len[ibeg - 1] = lbeg;
    if (log) std::fprintf(log, "%d, %d\n", ibeg, len[ibeg - 1]);

    int lend = operator_length(q, vel_of(iend), q.nfdmin, log);
    ++probes; effort += lend - 2 + 1;
    if (lend == 0) return false;
    len[iend - 1] = lend;

    auto finish = [&]() {
        if (log) {
            std::fprintf(log, "effectiveness =%f\n ", ((float)probes) / nvel);
            std::fprintf(log, "total search times for fd length : %d\n", effort);
        }
        return true;
    };
    if (lbeg == lend) { fill(ibeg + 1, iend, lbeg); return finish(); }

    int inc = (iend - ibeg) / (2 * (lbeg - lend));
    if (inc < 1) inc = 1;

    int  i1, l1, i2 = iend, l2 = lend, i3 = 0, l3 = 0;
    bool pending = false;  // a coarse stride (i3,l3) is waiting while its upper part is bisected
    auto probe = [&](int i, int start) {
        int l = operator_length(q, vel_of(i), start, log);
        ++probes; effort += l - start + 1;
        return l;
    };
    for (;;) {  // step one stride down from i2
        i1 = i2 - inc;
        if (i1 < ibeg) {
            i1 = ibeg; l1 = lbeg;
            if (i1 == i2) return finish();
        } else {
            l1 = probe(i1, l2);
            if (l1 == 0) return false;
        }
        for (;;) {  // resolve [i1, i2)
            if (l1 == l2 || i1 == i2 - 1) {
                if (l1 == l2) fill(i1, i2 - 1, l1); else len[i1 - 1] = l1;
                i2 = i1; l2 = l1;
                if (pending) {
                    pending = false;
                    if (l3 == l2) { fill(i3, i2 - 1, l2); i2 = i3; l2 = l3; break; }
                    i1 = i3; l1 = l3;
                    continue;
                }
                if (l2 == lbeg) { fill(ibeg, i2 - 1, l2); return finish(); }
                break;
            }
            const int i4 = (i1 + i2) / 2;
            if (!pending) { pending = true; i3 = i1; l3 = l1; }
            const int l4 = probe(i4, l2);
            if (l4 == 0) return false;
            if (l4 == l2) { fill(i4, i2 - 1, l2); i2 = i4; l2 = l4; }
            else          { i1 = i4; l1 = l4; }
        }
    }
}

int build_ls_operator(const OperatorSearch& q, int nvel, double vmin, double dv, const int* need,
                      std::vector<int>& M, std::vector<int>& Index, std::vector<float>& c,
                      std::FILE* log)
{
    // funMandC :22-73
    M.assign(nvel, 0);
    Index.assign(nvel + 1, 0);
    if (!operator_lengths(q, nvel, vmin, dv, M.data(), log) && log)
        std::fprintf(log, "ERROR in calculate M");
    for (int i = 0; i < nvel; ++i)
        if (need[i] == 0) M[i] = -1;
    for (int i = 1; i <= nvel; ++i) Index[i] = Index[i - 1] + M[i - 1] + 1;
    const int NC = Index[nvel];
    c.assign(NC, 0.0f);
    const double rat = q.tao / q.h;
    const double ra  = 2.0 * kPi * q.fmax * q.tao;
    parallel_for(nvel, [&](int i) {  // one least-squares system per used velocity bin
        if (need[i] != 1) return;
        std::vector<double> cd(q.nfdmax + 1);
        const double r = (vmin + i) * rat;  // (the reference assumes dv == 1 here, :58)
        const double b = ra / r;
        ls_coefficients(cd.data(), r, b, M[i], q.hzx);
        for (int l = 0; l <= M[i]; ++l) c[l + Index[i]] = (float)cd[l];
    });
    return NC;
}

void taylor_operator(int M, float* c)
{
    // order(2*M, c) :526-551: central second-derivative Taylor weights; the running
    // products are rounded to float at every step exactly as there.
    std::vector<float> x(M + 1, 1.0f);
    c[0] = 0.0f;
    for (int i = 1; i <= M; ++i) {
        for (int j = 1; j <= M; ++j) {
            if (j == i) continue;
            const double j2 = (double)j * j, i2 = (double)i * i;
            x[i] = (float)(x[i] * std::fabs(j2 / (j2 - i2)));
        }
        const double sign = ((i + 1) % 2 == 0) ? 1.0 : -1.0;
        c[i] = (float)(sign / ((double)i * i) * x[i]);
        c[0] = c[0] - 2 * c[i];
    }
}

}  // namespace rtm
