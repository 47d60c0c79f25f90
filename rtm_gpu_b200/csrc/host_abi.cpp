// C-ABI wrappers of the host-side pieces (include/rtm_b200.h, "host-side pieces").
#include "../../include/rtm_b200.h"
#include "host/rtm_host.h"

#include <cstring>
#include <vector>

extern "C" float rtm_ricker(float t1, float f0) { return rtm::ricker(t1, f0); }

extern "C" int rtm_source_row(float depth_m, float hz, int N2) { return rtm::source_row(depth_m, hz, N2); }

extern "C" void rtm_derived(float h, float hz, float tao, float tao1, float f0, int NT1, int* NT, int* NT2,
                            float* taoh, float* tao2, float* h2, float* taoh2, float* hzx2_1)
{
    rtm::RunConfig c;
    c.h = h; c.hz = hz; c.tao = tao; c.tao1 = tao1; c.f0 = f0; c.NT1 = NT1;
    const rtm::Geometry g = rtm::derive_geometry(c);
    if (NT) *NT = g.NT;
    if (NT2) *NT2 = g.NT2;
    if (taoh) *taoh = g.taoh;
    if (tao2) *tao2 = g.tao2;
    if (h2) *h2 = g.h2;
    if (taoh2) *taoh2 = g.taoh2;
    if (hzx2_1) *hzx2_1 = g.hzx2_1;
}

extern "C" void rtm_pad_velocity(const float* vraw, int mod_NZ, int mod_NX, int N2, int ifv, float* v)
{
    rtm::pad_velocity(vraw, mod_NZ, mod_NX, N2, ifv, v);
}

extern "C" int rtm_velocity_bins(const float* v, long ncell, float dv, float* vmin, float* vmax, int* need,
                                 int need_cap)
{
    const rtm::VelocityBins b = rtm::velocity_bins(v, ncell, dv);
    if (vmin) *vmin = b.vmin;
    if (vmax) *vmax = b.vmax;
    if (need)
        for (int i = 0; i < b.nvel && i < need_cap; ++i) need[i] = b.need[i];
    return b.nvel;
}

extern "C" void rtm_taylor_operator(int M, float* c) { rtm::taylor_operator(M, c); }

extern "C" void rtm_ls_coefficients(double* c, double r, double bmax, int M, double hzx)
{
    rtm::ls_coefficients(c, r, bmax, M, hzx);
}

extern "C" int rtm_ls_operator(int nthita, int nfdmax, int nfdmin, int nvel, float tao, float h, float df,
                               float eps, float fmax, float vmin, float dv, float hzx, const int* need,
                               int* M, int* Index, float* c, int c_cap, int verbose)
{
    // the reference hands its float parameters to funMandC's double arguments (kernel.cu:746)
    rtm::OperatorSearch q;
    q.nthita = nthita; q.nfdmax = nfdmax; q.nfdmin = nfdmin;
    q.tao = tao; q.h = h; q.df = df; q.eps = eps; q.fmax = fmax; q.hzx = hzx;
    q.nfre = (int)(q.fmax / q.df) + 1;  // LSMOrCon_rec_2D.cpp:29
    std::vector<int> Mv, Iv;
    std::vector<float> cv;
    const int NC = rtm::build_ls_operator(q, nvel, vmin, dv, need, Mv, Iv, cv, verbose ? stdout : nullptr);
    if (M) std::memcpy(M, Mv.data(), sizeof(int) * nvel);
    if (Index) std::memcpy(Index, Iv.data(), sizeof(int) * (nvel + 1));
    if (c)
        for (int i = 0; i < NC && i < c_cap; ++i) c[i] = cv[i];
    return NC;
}

extern "C" void rtm_resample(int nxin, float dxin, const float* yin, int nxout, float dxout, float* yout)
{
    rtm::resample_trace(nxin, dxin, yin, nxout, dxout, yout);
}

// ---- SEG-Y
int rtm_fail(int code, const char* fmt, ...);

extern "C" void rtm_segy_decode(const unsigned char* buf, float* out, int ns, int format) { rtm::segy_decode_samples(buf, out, ns, format); }
extern "C" void rtm_segy_encode(unsigned char* buf, const float* in, int ns, int format) { rtm::segy_encode_samples(buf, in, ns, format); }

extern "C" int rtm_segy_info(const char* path, int* ns, int* ntr, int* format, float* dt)
{
    std::string err;
    int a, b, c;
    float d;
    if (!rtm::segy_read_info(path, a, b, c, d, err)) return rtm_fail(RTM_ERR_IO, "%s", err.c_str());
    if (ns) *ns = a;
    if (ntr) *ntr = b;
    if (format) *format = c;
    if (dt) *dt = d;
    return RTM_OK;
}

extern "C" int rtm_segy_read(const char* path, float* out, int ns, int ntr)
{
    std::string err;
    if (!rtm::segy_read_traces(path, out, ns, ntr, err)) return rtm_fail(RTM_ERR_IO, "%s", err.c_str());
    return RTM_OK;
}

extern "C" int rtm_segy_write_image(const char* template_path, const char* out_path, const float* data, int ntr,
                                    int ns, int dt_value, const float* SX, const float* SY, float RX, float RY,
                                    const float* DSR)
{
    std::string err;
    if (!rtm::segy_write_image(template_path, out_path, data, ntr, ns, dt_value, SX, SY, RX, RY, DSR, err))
        return rtm_fail(RTM_ERR_IO, "%s", err.c_str());
    return RTM_OK;
}

// ---- post-stack chain
extern "C" int rtm_depth_to_time(const float* V, const float* D, int Nx, int Nz, float dz, float dt, float* T, int cap)
{
    std::vector<float> out;
    const int nt = rtm::depth_to_time(V, D, Nx, Nz, dz, dt, out);
    if (T) for (size_t i = 0; i < out.size() && i < (size_t)cap; ++i) T[i] = out[i];
    return nt;
}
extern "C" int rtm_time_to_depth(const float* V, const float* D, int Nx, int Nt, int Nz_V, float dtime, float ddepth,
                                 float* Z, int cap)
{
    std::vector<float> out;
    const int nz = rtm::time_to_depth(V, D, Nx, Nt, Nz_V, dtime, ddepth, out);
    if (Z) for (size_t i = 0; i < out.size() && i < (size_t)cap; ++i) Z[i] = out[i];
    return nz;
}
extern "C" void rtm_phase_rotate(const float* din, float* dout, int ntr, int nt, float angle_deg)
{
    rtm::phase_rotate(din, dout, ntr, nt, angle_deg);
}
