// The reference's driver surface: parameter files, derived geometry, Ricker wavelet
// and the stdout echo (kernel.cu:542-662, 693-700, 794-795, 1261-1266).
#include "rtm_host.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rtm {
namespace {

// The run file alternates a free-text label line and a value line (kernel.cu:544-599).
// Lines may end in CRLF; surrounding blanks are ignored like fscanf's whitespace skip.
std::string trimmed(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

struct ValueReader {
    std::vector<std::string> values;
    size_t next = 0;
    bool   ok   = true;
    std::string get()
    {
        if (next >= values.size()) { ok = false; return "0"; }
        return values[next++];
    }
    int   i() { return (int)std::strtol(get().c_str(), nullptr, 10); }
    float f() { return std::strtof(get().c_str(), nullptr); }
    std::string s()
    {
        // %s stops at the first blank
        std::string v = get();
        size_t sp = v.find_first_of(" \t");
        return sp == std::string::npos ? v : v.substr(0, sp);
    }
};

}  // namespace

bool parse_run_file(const char* path, RunConfig& c, std::string& err)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = std::string("cannot open run file ") + path; return false; }
    ValueReader r;
    std::string line;
    std::vector<std::string> lines;
    while (std::getline(in, line)) {
        line = trimmed(line);
        if (!line.empty()) lines.push_back(line);  // fscanf("\n") swallows blank lines
    }
    for (size_t k = 1; k < lines.size(); k += 2) r.values.push_back(lines[k]);
    c.nfdmax = r.i(); c.nfdmin = r.i(); c.N2 = r.i();
    c.f0 = r.f(); c.fmax = r.f(); c.df = r.f();
    c.nthita = r.i(); c.eps = r.f(); c.dv = r.f();
    c.iLSTE = r.i(); c.ifv = r.i();
    c.whitecoe = r.f(); c.hz = r.f(); c.tao = r.f();
    c.iNorm = r.i(); c.iCompen = r.i(); c.Nsmooth = r.i();
    c.wthite_phase = r.f(); c.angle = r.f();
    c.NX_BG = r.i(); c.NX_ED = r.i(); c.NZ_BG = r.i(); c.NZ_ED = r.i();
    c.OutNameseis = r.s(); c.OutNameVp = r.s(); c.OutNameDPR = r.s(); c.OutPara = r.s();
    c.Result = r.s();
    if (!r.ok) { err = std::string("run file ") + path + " has fewer than 28 values"; return false; }
    return true;
}

bool parse_parameter_file(const char* path, RunConfig& c, std::string& err)
{
    std::ifstream in(path);
    if (!in) { err = std::string("cannot open parameter file ") + path; return false; }
    // h tao1 mod_NZ mod_NX NT1 s_l s_z n ds r_x nrec dr   (kernel.cu:603)
    if (!(in >> c.h >> c.tao1 >> c.mod_NZ >> c.mod_NX >> c.NT1 >> c.s_l >> c.s_z >> c.n >> c.ds >>
          c.r_x >> c.nrec >> c.dr)) {
        err = std::string("parameter file ") + path + " needs 12 values";
        return false;
    }
    // every later size derives from these: reject what the reference would silently turn into a crash
    auto bad = [&](const char* name, double v) { err = std::string("parameter file ") + path + ": " + name + " = " + std::to_string(v) + " is not positive"; return false; };
    if (!(c.h > 0)) return bad("h", c.h);
    if (!(c.tao1 > 0)) return bad("tao1", c.tao1);
    if (c.mod_NZ < 1) return bad("mod_NZ", c.mod_NZ);
    if (c.mod_NX < 1) return bad("mod_NX", c.mod_NX);
    if (c.NT1 < 1) return bad("NT1", c.NT1);
    if (c.n < 1) return bad("n", c.n);
    if (c.ds < 1) return bad("ds", c.ds);
    if (c.nrec < 1) return bad("nrec", c.nrec);
    return true;
}

bool parse_depth_file(const char* path, RunConfig& c, std::string& err)
{
    std::ifstream in(path);
    if (!in) { err = std::string("cannot open receiver depth file ") + path; return false; }
    if (c.nrec < 1 || c.nrec > (1 << 24)) { err = "nrec = " + std::to_string(c.nrec) + " (Parameter.txt) must be a positive count"; return false; }
    c.INRE.assign(c.nrec, 0.0f);
    for (int i = 0; i < c.nrec; ++i)
        if (!(in >> c.INRE[i])) { err = "receiver depth file has fewer than nrec values"; return false; }
    return true;
}

Geometry derive_geometry(const RunConfig& c)
{
    Geometry g;
    g.s_l = c.s_l + c.N2 - 1;
    g.r_x = c.r_x + c.N2 - 1;
    g.s_z = c.s_z + c.N2 - 1;
    g.s_r = (c.n - 1) * c.ds + g.s_l;
    g.NZ  = c.mod_NZ + 2 * c.N2;
    g.NX  = c.mod_NX + 2 * c.N2;
    g.NT  = (int)((c.NT1 - 1) * c.tao1 / c.tao + 1.5);
    g.taoh   = c.tao / c.h;
    g.tao2   = (float)((double)c.tao * (double)c.tao);    // pow(tao,2)
    g.h2     = (float)(1 / ((double)c.h * (double)c.h));  // 1/pow(h,2)
    g.taoh2  = g.tao2 * g.h2 / 2;
    g.NT2    = (int)(2.0 / (c.f0 * c.tao)) + 1;
    g.hzx    = c.hz / c.h;
    g.hzx2_1 = 1 / (g.hzx * g.hzx);
    return g;
}

int source_row(float depth_m, float hz, int N2)
{
    const int N = (int)depth_m;  // kernel.cu:794
    return (int)(std::fabs(N / hz) + N2 - 1);
}

float ricker(float t1, float f0)
{
    const float  t00 = 1 / f0;
    const double a   = 3.1415926535898 * f0 * (t1 - t00);
    const double a2  = a * a;
    return (float)((1 - 2 * a2) * std::exp(-a2));
}

void echo_config(const RunConfig& c, const Geometry& g, std::FILE* out)
{
    const int Nmax = g.NX < g.NZ ? g.NZ : g.NX;
    std::fprintf(out, "ifv=%d\n", c.ifv);
    std::fprintf(out, "Nmax=%d\n", Nmax);
    std::fprintf(out, "hzx=%f,hzx2_1=%f\n", g.hzx, g.hzx2_1);
    std::fprintf(out, "The maximum length of operator\nnfdmax=%d\n", c.nfdmax);
    std::fprintf(out, "The minimum length of operator\nnfdmin=%d\n", c.nfdmin);
    std::fprintf(out, "The hyrid absorbing boundary width\nN2=%d\n", c.N2);
    std::fprintf(out, "space interval\nh=%f\n", c.h);
    std::fprintf(out, "time interval\ntao=%f\n", c.tao);
    std::fprintf(out, "z grid dimension\nmod_NZ=%d\n", c.mod_NZ);
    std::fprintf(out, "x grid dimension\nmod_NX=%d\n", c.mod_NX);
    std::fprintf(out, "actual grid number in z\nNZ=%d\n", g.NZ);
    std::fprintf(out, "actual grid number in x\nNX=%d\n", g.NX);
    std::fprintf(out, "Number of time number\nNT=%d\n", g.NT);
    std::fprintf(out, "Source X\ns_x=%d\n", g.s_l);
    std::fprintf(out, "Source Z\ns_z=%d\n", g.s_z);
    std::fprintf(out, "The number of sources\nn=%d\n", c.n);
    std::fprintf(out, "The interval of sources\nds=%d\n", c.ds);
    std::fprintf(out, "Dominant Frequency\nf0=%f\n", c.f0);
    std::fprintf(out, "Maximum Frequency\nfmax=%f\n", c.fmax);
    std::fprintf(out, "Interval of Frequency\ndf=%f\n", c.df);
    std::fprintf(out, "Azimuth of the plane wave divide into the number\nnthita=%d\n", c.nthita);
    std::fprintf(out, "dispersion value\neps=%f\n", c.eps);
    std::fprintf(out, "velocity interval\ndv=%f\n", c.dv);
    std::fprintf(out, "Interval of receiver\ndr=%d\n", c.dr);
    std::fprintf(out, "The number of receivers\nnrec=%d\n", c.nrec);
    std::fprintf(out, "Receiver Z\nr_x=%d\n", g.r_x);
    std::fprintf(out, "LSM-0,TEM-1\niLSTE=%d\n", c.iLSTE);
    std::fprintf(out, "hz=%f\n", c.hz);
    std::fprintf(out, "iNorm=%d\n", c.iNorm);
    std::fprintf(out, "iCompen=%d\n", c.iCompen);
    std::fprintf(out, "Nsmooth=%d\n", c.Nsmooth);
    std::fprintf(out, "wthite_phase=%f\n", c.wthite_phase);
    std::fprintf(out, "whitecoe=%f\n", c.whitecoe);
    std::fprintf(out, "angle=%f\n", c.angle);
}

}  // namespace rtm
