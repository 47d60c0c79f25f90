// Velocity model preparation on the host (reference: GPU_velocity_real.cpp:6-118 and
// kernel.cu:704-738).
#include "rtm_host.h"

#include <algorithm>
#include <cstdio>

namespace rtm {

void pad_velocity(const float* vraw, int mod_NZ, int mod_NX, int N2, int ifv, float* v)
{
    // The reference replicates the outermost interior sample into the N2-wide ring side
    // by side and corner by corner; that equals clamping the source index.  The raw file
    // is x-outer / z-inner (GPU_velocity_real.cpp:11-19); ifv==1 mirrors the padded model
    // in x (:84-100).
    const int NZ = mod_NZ + 2 * N2, NX = mod_NX + 2 * N2;
    for (int z = 0; z < NZ; ++z) {
        const int zs = std::clamp(z - N2, 0, mod_NZ - 1);
        float* row   = v + (size_t)z * NX;
        for (int x = 0; x < NX; ++x) {
            const int xd = (ifv == 1) ? NX - 1 - x : x;
            const int xs = std::clamp(xd - N2, 0, mod_NX - 1);
            row[x] = vraw[(size_t)xs * mod_NZ + zs];
        }
    }
}

bool read_velocity(const char* path, int mod_NZ, int mod_NX, std::vector<float>& vraw, std::string& err)
{
    std::FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open velocity file ") + path; return false; }
    vraw.resize((size_t)mod_NZ * mod_NX);
    size_t got = std::fread(vraw.data(), sizeof(float), vraw.size(), f);
    std::fclose(f);
    if (got != vraw.size()) { err = std::string("short velocity file ") + path; return false; }
    return true;
}

VelocityBins velocity_bins(const float* v, long ncell, float dv)
{
    VelocityBins b;
    float vmin = v[0], vmax = v[0];
    for (long i = 0; i < ncell; ++i) {
        vmin = std::min(vmin, v[i]);
        vmax = std::max(vmax, v[i]);
    }
    // snap outward to the dv grid (kernel.cu:715-720); float arithmetic as there
    float vel = ((int)(vmin / dv)) * dv;
    vmin = (vel > vmin) ? vel - dv : vel;
    vel = ((int)(vmax / dv)) * dv;
    vmax = (vel < vmax) ? vel + dv : vel;
    b.vmin = vmin;
    b.vmax = vmax;
    b.nvel = (int)((vmax - vmin) / dv + 1.5);
    b.need.assign(b.nvel, 0);
    for (long i = 0; i < ncell; ++i) {
        const int k = (int)((v[i] - vmin) / dv + 0.5);
        if (k >= 0 && k < b.nvel) b.need[k] = 1;
    }
    return b;
}

}  // namespace rtm
