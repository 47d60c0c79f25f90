// SEG-Y input/output for the driver surface (reference: segy.cpp:362-760 sample/header codec,
// SGYWrite.cpp:3-55 image writer).  Big-endian on disk; sample formats 1 (IBM float),
// 2 (int32), 3 (int16), 5 (IEEE float).  Written from the SEG-Y rev1 layout; the IBM <-> IEEE
// conversions keep the reference's truncating behaviour so files are byte-identical
// (tests/test_host.py::test_segy_*).
#include "rtm_host.h"

#include <cstdint>
#include <cstdio>
#include <cstring>

namespace rtm {
namespace {

constexpr int kTextBytes = 3200, kBinBytes = 400, kTraceHdrBytes = 240, kNKeys = 91;
constexpr int kBinDt = 16, kBinNs = 20, kBinFormat = 24;  // offsets inside the binary header
// byte width of the 91 standard trace-header words, in file order (sums to 240)
const char kKeyWidth[kNKeys + 1] =
    "4444444222244444444224444222222222222222222222222222222222222222222222244444224222224242244";

inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int      be16(const unsigned char* p) { return (int16_t)(((unsigned)p[0] << 8) | p[1]); }
inline void put32(unsigned char* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
inline void put16(unsigned char* p, int v) { const uint16_t u = (uint16_t)(int16_t)v; p[0] = u >> 8; p[1] = u & 0xff; }

}  // namespace

// ibm_to_float / float_to_ibm: transliterated for bit-identity from ibm_to_float / float_to_ibm of
// segy.cpp:552-650 (same shift-and-normalise steps, hence the same results for every bit pattern,
// including the saturation and underflow cases); the header codec and file IO around them are new.
float ibm_to_float(uint32_t x)
{
    // base-16 exponent (excess 64), 24-bit fraction -> IEEE single; the fraction is normalised by
    // shifting left (no rounding is ever needed), overflow saturates to the largest float
    if ((x & 0x7fffffffu) == 0) return 0.0f;
    uint32_t s = x & 0x80000000u, f = x & 0x00ffffffu;
    int e = (int)((x & 0x7f000000u) >> 24) - 64;
    e = (e >= 0) ? (e << 2) : -((-e) << 2);
    e -= 1;
    if (f != 0)
        while ((f & 0x00800000u) == 0) { f <<= 1; e -= 1; }
    f &= 0x007fffffu;
    e += 127;
    if (e >= 255) s |= 0x7f7fffffu;
    else if (e > 0) s |= ((uint32_t)e << 23) | f;
    float y;
    std::memcpy(&y, &s, 4);
    return y;
}

uint32_t float_to_ibm(float y)
{
    // IEEE single -> IBM: bits that do not fit the base-16 fraction are truncated (not rounded),
    // underflow gives a signed zero, overflow saturates
    uint32_t x;
    std::memcpy(&x, &y, 4);
    if ((x & 0x7fffffffu) == 0) return x;
    uint32_t s = x & 0x80000000u, f = ((x & 0x007fffffu) << 1) | 0x01000000u;
    int e = (int)((x & 0x7f800000u) >> 23) - 127;
    if (e >= 0) { f <<= (e & 3); e >>= 2; }
    else        { f >>= ((-e) & 3); e = -((-e) >> 2); }
    if (f & 0x0f000000u) { f >>= 4; e += 1; }
    e += 64;
    if (e > 127) s |= 0x7fffffffu;
    else if (e >= 0) s |= ((uint32_t)e << 24) | f;
    return s;
}

void segy_decode_samples(const unsigned char* buf, float* out, int ns, int format)
{
    const int nb = (format == 3) ? 2 : 4;
    for (int i = 0; i < ns; ++i, buf += nb) {
        switch (format) {
        case 1: out[i] = ibm_to_float(be32(buf)); break;
        case 2: out[i] = (float)(int32_t)be32(buf); break;
        case 3: out[i] = (float)be16(buf); break;
        case 5: { uint32_t u = be32(buf); std::memcpy(&out[i], &u, 4); break; }
        default: out[i] = 0.0f; break;
        }
    }
}

void segy_encode_samples(unsigned char* buf, const float* in, int ns, int format)
{
    const int nb = (format == 3) ? 2 : 4;
    for (int i = 0; i < ns; ++i, buf += nb) {
        switch (format) {
        case 1: put32(buf, float_to_ibm(in[i])); break;
        case 2: put32(buf, (uint32_t)(int32_t)in[i]); break;
        case 3: put16(buf, (int)in[i]); break;
        case 5: { uint32_t u; std::memcpy(&u, &in[i], 4); put32(buf, u); break; }
        default: break;
        }
    }
}

void segy_unpack_header(const unsigned char* buf, int* words)
{
    for (int i = 0; i < kNKeys; ++i) {
        if (kKeyWidth[i] == '2') { words[i] = be16(buf); buf += 2; }
        else                     { words[i] = (int32_t)be32(buf); buf += 4; }
    }
}

void segy_pack_header(unsigned char* buf, const int* words)
{
    for (int i = 0; i < kNKeys; ++i) {
        if (kKeyWidth[i] == '2') { put16(buf, words[i]); buf += 2; }
        else                     { put32(buf, (uint32_t)words[i]); buf += 4; }
    }
}

bool segy_read_info(const char* path, int& ns, int& ntr, int& format, float& dt, std::string& err)
{
    std::FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open SEG-Y file ") + path; return false; }
    unsigned char bh[kBinBytes];
    bool ok = std::fseek(f, kTextBytes, SEEK_SET) == 0 && std::fread(bh, 1, kBinBytes, f) == (size_t)kBinBytes;
    long size = 0;
    if (ok) { std::fseek(f, 0, SEEK_END); size = std::ftell(f); }
    std::fclose(f);
    if (!ok) { err = std::string("short SEG-Y file ") + path; return false; }
    format = be16(bh + kBinFormat);
    ns = be16(bh + kBinNs) & 0xffff;
    dt = (float)(be16(bh + kBinDt) / 1000000.);  // segydt(), segy.cpp:524-528
    if (format != 1 && format != 2 && format != 3 && format != 5) { err = "unsupported SEG-Y sample format"; return false; }
    const long trace_bytes = kTraceHdrBytes + (long)ns * (format == 3 ? 2 : 4);
    ntr = (int)((size - kTextBytes - kBinBytes) / trace_bytes);
    if (ns <= 0 || ntr <= 0) { err = "empty SEG-Y file"; return false; }
    return true;
}

bool segy_read_traces(const char* path, float* out, int ns, int ntr, std::string& err)
{
    int ns2, ntr2, format;
    float dt;
    if (!segy_read_info(path, ns2, ntr2, format, dt, err)) return false;
    if (ns2 != ns || ntr2 < ntr) { err = "SEG-Y file does not have the expected ns/ntr"; return false; }
    std::FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open SEG-Y file ") + path; return false; }
    const int nb = ns * (format == 3 ? 2 : 4);
    std::vector<unsigned char> buf(nb);
    std::fseek(f, kTextBytes + kBinBytes, SEEK_SET);
    for (int i = 0; i < ntr; ++i) {
        if (std::fseek(f, kTraceHdrBytes, SEEK_CUR) != 0 || std::fread(buf.data(), 1, nb, f) != (size_t)nb) {
            std::fclose(f);
            err = "short SEG-Y trace";
            return false;
        }
        segy_decode_samples(buf.data(), out + (size_t)i * ns, ns, format);
    }
    std::fclose(f);
    return true;
}

bool segy_write_image(const char* template_path, const char* out_path, const float* data, int ntr, int ns,
                      int dt_value, const float* SX, const float* SY, float RX, float RY, const float* DSR,
                      std::string& err)
{
    // WriteSGY, SGYWrite.cpp:3-55: text + binary header and the first trace header come from the
    // template; ns and dt are overwritten; per trace the words offset(11), sx, sy, gx, gy (21..24)
    // are set; samples are encoded in the template's format
    std::FILE* t = std::fopen(template_path, "rb");
    if (!t) { err = std::string("cannot open SEG-Y template ") + template_path; return false; }
    std::vector<unsigned char> head(kTextBytes + kBinBytes), th(kTraceHdrBytes);
    const bool ok = std::fread(head.data(), 1, head.size(), t) == head.size() &&
                    std::fread(th.data(), 1, th.size(), t) == th.size();
    std::fclose(t);
    if (!ok) { err = "SEG-Y template is too short"; return false; }
    unsigned char* bh = head.data() + kTextBytes;
    const int format = be16(bh + kBinFormat);
    put16(bh + kBinNs, ns);
    const float dtf = (float)dt_value;                      // set_segydt(), segy.cpp:530-539
    const float scale = (dtf < 1.0) ? 1000000.f : 1000.f;
    put16(bh + kBinDt, (int)(scale * dtf));
    std::FILE* o = std::fopen(out_path, "wb");
    if (!o) { err = std::string("cannot write ") + out_path; return false; }
    std::fwrite(head.data(), 1, head.size(), o);
    const int nb = ns * (format == 3 ? 2 : 4);
    std::vector<unsigned char> buf(nb);
    int words[kNKeys];
    for (int i = 0; i < ntr; ++i) {
        segy_unpack_header(th.data(), words);
        words[11] = (int)DSR[i];
        words[21] = (int)SX[i];
        words[22] = (int)SY[i];
        words[23] = (int)RX;
        words[24] = (int)RY;
        segy_pack_header(th.data(), words);
        std::fwrite(th.data(), 1, th.size(), o);
        segy_encode_samples(buf.data(), data + (size_t)i * ns, ns, format);
        std::fwrite(buf.data(), 1, nb, o);
    }
    std::fclose(o);
    return true;
}

}  // namespace rtm
