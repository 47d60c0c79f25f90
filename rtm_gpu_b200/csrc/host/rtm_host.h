// Host-side (C++) pieces of the RTM engine: everything the reference does on the CPU
// before and after the device time loop that the hot path depends on.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace rtm {

// ---------------------------------------------------------------- run configuration
// The reference's driver surface: 2D_Real_RVSP_RTM.txt (28 values, each preceded by a
// free-text label line; kernel.cu:542-600), Parameter.txt (12 values; :602-604) and the
// receiver-depth list (:693-700).  Field names follow the reference.
struct RunConfig {
    int   nfdmax = 0, nfdmin = 0, N2 = 0;
    float f0 = 0, fmax = 0, df = 0;
    int   nthita = 0;
    float eps = 0, dv = 0;
    int   iLSTE = 0, ifv = 0;
    float whitecoe = 0, hz = 0, tao = 0;
    int   iNorm = 0, iCompen = 0, Nsmooth = 0;
    float wthite_phase = 0, angle = 0;
    int   NX_BG = 0, NX_ED = 0, NZ_BG = 0, NZ_ED = 0;
    std::string OutNameseis, OutNameVp, OutNameDPR, OutPara, Result;
    // Parameter.txt
    float h = 0, tao1 = 0;
    int   mod_NZ = 0, mod_NX = 0, NT1 = 0, s_l = 0, s_z = 0, n = 0, ds = 0, r_x = 0, nrec = 0, dr = 0;
    std::vector<float> INRE;  // receiver depths (metres)
};

// Geometry and scalars derived exactly as kernel.cu:607-628 (float arithmetic where the
// reference uses float).  Indices are padded and 0-based.
struct Geometry {
    int   NZ = 0, NX = 0, NT = 0, NT2 = 0;
    int   s_l = 0, s_r = 0, s_z = 0, r_x = 0;
    float taoh = 0, tao2 = 0, h2 = 0, taoh2 = 0, hzx = 0, hzx2_1 = 0;
};

bool  parse_run_file(const char* path, RunConfig& cfg, std::string& err);
bool  parse_parameter_file(const char* path, RunConfig& cfg, std::string& err);
bool  parse_depth_file(const char* path, RunConfig& cfg, std::string& err);
Geometry derive_geometry(const RunConfig& cfg);
int   source_row(float depth_m, float hz, int N2);          // r_u, kernel.cu:794-795
float ricker(float t1, float f0);                            // f(), kernel.cu:1261-1266
void  echo_config(const RunConfig& cfg, const Geometry& g, std::FILE* out);  // :629-662

// ---------------------------------------------------------------- model
// raw [mod_NX][mod_NZ] -> padded [NZ][NX] with edge replication and optional x flip
// (GPU_velocity_real.cpp:6-100)
void pad_velocity(const float* vraw, int mod_NZ, int mod_NX, int N2, int ifv, float* v);
bool read_velocity(const char* path, int mod_NZ, int mod_NX, std::vector<float>& vraw, std::string& err);

struct VelocityBins {
    float vmin = 0, vmax = 0;  // snapped outward to multiples of dv
    int   nvel = 0;
    std::vector<int> need;     // bin used by at least one cell
};
VelocityBins velocity_bins(const float* v, long ncell, float dv);  // kernel.cu:704-738

// ---------------------------------------------------------------- FD operator
struct OperatorSearch {
    int    nthita = 0, nfdmax = 0, nfdmin = 0, nfre = 0;
    double tao = 0, h = 0, df = 0, eps = 0, fmax = 0, hzx = 0;
};
void ls_coefficients(double* c, double r, double bmax, int M, double hzx);
int  operator_length(const OperatorSearch& q, double vel, int start_len, std::FILE* log);
bool operator_lengths(const OperatorSearch& q, int nvel, double vmin, double dv, int* len, std::FILE* log);
int  build_ls_operator(const OperatorSearch& q, int nvel, double vmin, double dv, const int* need,
                       std::vector<int>& M, std::vector<int>& Index, std::vector<float>& c,
                       std::FILE* log);
void taylor_operator(int M, float* c);

// ---------------------------------------------------------------- data preparation
// resample(), Resample.cpp:193-225: 8-point tabulated-sinc interpolation of one trace
void resample_trace(int nxin, float dxin, const float* yin, int nxout, float dxout, float* yout);
// the 513 x 8 interpolation table (for the device version of the same interpolation)
const float* sinc_table(int* nshifts, int* ntaps);

// ---------------------------------------------------------------- SEG-Y (segy.cpp, SGYWrite.cpp)
float    ibm_to_float(uint32_t word);
uint32_t float_to_ibm(float y);
void segy_decode_samples(const unsigned char* buf, float* out, int ns, int format);   // segy2trace :653
void segy_encode_samples(unsigned char* buf, const float* in, int ns, int format);    // trace2segy :675
void segy_unpack_header(const unsigned char* buf, int* words91);                      // segy2head :697
void segy_pack_header(unsigned char* buf, const int* words91);                        // head2segy :746
bool segy_read_info(const char* path, int& ns, int& ntr, int& format, float& dt, std::string& err);
bool segy_read_traces(const char* path, float* out, int ns, int ntr, std::string& err);
bool segy_write_image(const char* template_path, const char* out_path, const float* data, int ntr, int ns,
                      int dt_value, const float* SX, const float* SY, float RX, float RY, const float* DSR,
                      std::string& err);                                              // WriteSGY

// ---------------------------------------------------------------- post-stack chain (kernel.cu:1110-1179)
// V, D: [Nx][Nz] (x outer).  Return the number of output samples per trace; T/Z: [Nx][n].
int  depth_to_time(const float* V, const float* D, int Nx, int Nz, float dz, float dt, std::vector<float>& T);   // D2T
int  time_to_depth(const float* V, const float* D, int Nx, int Nt_in, int Nz_V, float dtime, float ddepth,
                   std::vector<float>& Z);                                                                      // T2D
void phase_rotate(const float* din, float* dout, int ntr, int nt, float angle_deg);                             // phase_correction

}  // namespace rtm
