"""Host-side check of the launch geometry helpers of rtm_kernels.cuh / rtm_ring.cuh / rtm_stream.cuh (compiled for the
host with nvcc, no GPU needed): the division by multiply-high, the mapping of block indices to ring / interior CTAs,
the tiling of the ring kernel and the regions of the streaming form (segments, thin frame)."""
import shutil
import subprocess

import pytest

from refcase import ROOT

SRC = r"""
#include "rtm_kernels.cuh"
#include <cstdio>
#include <vector>
using namespace rtmk;
int main()
{
    // fast_div == integer division for every divisor / numerator the launches can produce
    const int ds[] = {1, 2, 3, 5, 7, 18, 47, 50, 126, 846, 3456, 24649, 65535, 1 << 20};
    for (int d : ds) {
        const FastDiv f = make_fastdiv(d);
        for (long long n = 0; n < (1ll << 31); n += (n < 100000 ? 1 : 104729))
            if (fast_div((int)n, f) != (int)(n / d)) { std::printf("fast_div(%lld, %d)\n", n, d); return 1; }
        if (fast_div(2147483647, f) != 2147483647 / d) { std::printf("fast_div(max, %d)\n", d); return 1; }
    }
    // block_role: a bijection onto ring CTAs 0..nrc-1 and interior CTAs 0..nint-1, ring CTAs at multiples of the
    // period, the last one inside the grid; both with and without interleaving
    const int nrcs[] = {0, 1, 7, 50, 400, 1600, 3184}, nints[] = {0, 1, 3, 126, 1008, 3456, 13824, 197192};
    for (int il = 0; il <= 8; ++il)   // 0: ring CTAs first; 1..8: dealt over the first il/8 of the grid
        for (int nrc : nrcs)
            for (int nint : nints) {
                const int total = nrc + nint;
                if (total == 0) continue;
                const int period = ring_period_for(il != 0, nrc, total, il);
                if (period < 1 || (nrc > 0 && (long long)(nrc - 1) * period >= total)) { std::printf("period %d %d %d\n", nrc, nint, period); return 2; }
                if (!il && period != 1) return 3;
                const FastDiv fd = make_fastdiv(period);
                std::vector<char> ring(nrc, 0), inner(nint, 0);
                int last_inner = -1;
                for (int b = 0; b < total; ++b) {
                    const BlockRole r = block_role(b, nrc, period, fd);
                    if (r.index < 0 || r.index >= (r.is_ring ? nrc : nint)) { std::printf("range %d %d %d %d\n", nrc, nint, b, r.index); return 4; }
                    char& seen = r.is_ring ? ring[r.index] : inner[r.index];
                    if (seen) { std::printf("twice %d %d %d\n", nrc, nint, b); return 5; }
                    seen = 1;
                    if (r.is_ring && b != r.index * period) return 6;
                    if (!r.is_ring) { if (r.index != last_inner + 1) return 7; last_inner = r.index; }  // tile order kept
                }
            }
    std::printf("ok\n");
    return 0;
}
"""


def test_fast_div_and_block_roles(tmp_path):
    nvcc = shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not on PATH")
    src = tmp_path / "geom.cu"
    src.write_text(SRC)
    exe = tmp_path / "geom"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-I",
           str(ROOT / "rtm_gpu_b200" / "csrc"), "-o", str(exe), str(src), "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    cc = subprocess.run(cmd, capture_output=True, text=True)
    if cc.returncode != 0:  # (nvcc shares temporary names under /tmp with concurrent builds: one retry)
        cc = subprocess.run(cmd, capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr[-500:])


RING_SRC = r"""
#include "rtm_ring.cuh"
#include <cstdio>
#include <vector>
using namespace rtmk;
int main()
{
    // ring_kernel's tiling: every ring cell lies in exactly one tile's output rectangle; every compute rectangle starts on a
    // 16-byte boundary in global memory (TMA box) and holds the output rectangle grown by one cell; a tile's cells fit the
    // per-tile stride of the coefficient arrays; the boxes fit the shared-memory layout
    const int grids[][3] = {{2301, 751, 10}, {677, 210, 10}, {4096, 4096, 12}, {120, 100, 10}, {60, 400, 10}, {333, 150, 16},
                            {333, 150, 3}, {700, 280, 10}, {20000, 64, 10}, {90, 70, 12}, {40, 150, 10}, {4096, 512, 12}};
    for (auto& g3 : grids)
        for (int R : {1, 4, 5, 8, 12}) {
            Geo G{};
            G.mod_NX = g3[0]; G.mod_NZ = g3[1]; G.N2 = g3[2];
            if (R > G.N2) continue;
            const int RP = (R + 3) / 4 * 4;
            G.NX = G.mod_NX + 2 * G.N2; G.NZ = G.mod_NZ + 2 * G.N2;
            G.padL = (32 - G.N2 % 32) % 32;
            if (G.padL + G.N2 < RP) G.padL += 32;
            G.pitch = (G.padL + G.NX + 4 + 31) / 32 * 32;
            const RingGeo rg = make_ring_geo(G, R, RP);
            std::vector<int> hit((size_t)G.NZ * G.NX, 0);
            for (int t = 0; t < rg.ntiles; ++t) {
                int cz0, cx0;
                const RingRect o = ring2_rect(G, rg, t, &cz0, &cx0);
                const bool band = t < 2 * rg.nband;
                const int ch = band ? rg.chB : rg.chS, cw = band ? rg.cwB : rg.cwS, sp = band ? rg.spB : rg.spS;
                if ((G.padL + cx0) % 4 != 0) { std::printf("align %d %d %d tile %d\n", g3[0], g3[1], g3[2], t); return 1; }
                if (o.za >= o.zb || o.xa >= o.xb) continue;   // (an empty trailing tile)
                if (o.za - 1 < cz0 || o.zb + 1 > cz0 + ch || o.xa - 1 < cx0 || o.xb + 1 > cx0 + cw) { std::printf("grow %d %d %d tile %d\n", g3[0], g3[1], g3[2], t); return 2; }
                if ((o.zb - o.za) * (o.xb - o.xa) > rg.cells) { std::printf("cells\n"); return 3; }
                if ((ch + 2 * R) * sp > rg.n1 || ch * cw > rg.nc || sp != cw + 2 * RP) { std::printf("smem\n"); return 4; }
                for (int z = o.za; z < o.zb; ++z)
                    for (int x = o.xa; x < o.xb; ++x) ++hit[(size_t)z * G.NX + x];
            }
            for (int z = 0; z < G.NZ; ++z)
                for (int x = 0; x < G.NX; ++x) {
                    const bool ring = z < G.N2 || z >= G.NZ - G.N2 || x < G.N2 || x >= G.NX - G.N2;
                    if (hit[(size_t)z * G.NX + x] != (ring ? 1 : 0)) { std::printf("cover %d %d %d at %d %d: %d\n", g3[0], g3[1], g3[2], z, x, hit[(size_t)z * G.NX + x]); return 5; }
                }
        }
    std::printf("ok\n");
    return 0;
}
"""


def test_ring_kernel_tiling(tmp_path):
    nvcc = shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not on PATH")
    src = tmp_path / "ringgeo.cu"
    src.write_text(RING_SRC)
    exe = tmp_path / "ringgeo"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-I",
           str(ROOT / "rtm_gpu_b200" / "csrc"), "-o", str(exe), str(src), "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    cc = subprocess.run(cmd, capture_output=True, text=True)
    if cc.returncode != 0:
        cc = subprocess.run(cmd, capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr


STREAM_SRC = r"""
#include "rtm_stream.cuh"
#include <cstdio>
#include <vector>
using namespace rtmk;
// Every interior cell belongs to exactly one of: a streamed segment (ii or ib), a thin-frame tile.  Inner-inner
// segments keep their distance from everything stepped singly; TMA box starts are 16-byte aligned.
static int check(int mod_NX, int mod_NZ, int N2, int seg_blocks)
{
    Geo G{};
    G.N2 = N2; G.mod_NX = mod_NX; G.mod_NZ = mod_NZ; G.NX = mod_NX + 2 * N2; G.NZ = mod_NZ + 2 * N2;
    const int RP = 4, BR = Strm<4>::BR;
    G.padL = (32 - N2 % 32) % 32;   // field_layout() of rtm_engine.cu: the first interior column sits on a 128-byte boundary
    if (G.padL + N2 < RP) G.padL += 32;
    const StreamRegions r = make_stream_regions(G, RP, seg_blocks);
    const int C0 = N2 + RP, R0 = N2 + RP, xe = G.NX - N2 - RP, ze = G.NZ - N2 - RP;
    if (!r.ok) return (xe - C0 > 2 * kTX && ze - R0 >= 2 * BR) ? 1 : 0;   // must be available on any grid of 3+ columns
    if (r.C0 != C0 || r.R0 != R0 || r.xe != xe || r.ze != ze) return 2;
    std::vector<int> own((size_t)G.NZ * G.NX, 0);
    auto mark = [&](int z, int x, int who) { if (z < 0 || z >= G.NZ || x < 0 || x >= G.NX) return false; int& o = own[(size_t)z * G.NX + x]; if (o) return false; o = who; return true; };
    double cells = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (const int4& s : (pass ? r.ib : r.ii)) {
            if (s.z < 1 || (s.x - C0) % kTX || (s.y - R0) % BR) return 3;
            if ((G.padL + s.x - 2 * RP) % 4) return 4;                       // TMA box start (current fields, halo 2 RP)
            if (s.z > seg_blocks && seg_blocks >= 2) return 5;
            const int zend = s.y + BR * s.z < ze ? s.y + BR * s.z : ze, xend = s.x + kTX < xe ? s.x + kTX : xe;
            if (zend <= s.y || xend <= s.x) return 6;                        // no empty segment
            for (int z = s.y; z < zend; ++z)
                for (int x = s.x; x < xend; ++x) { if (!mark(z, x, pass ? 2 : 1)) return 7; ++cells; }
            // edge flags: the segment touches the thin frame on that side
            const int e = (s.x == C0 ? 1 : 0) | (s.x + kTX >= xe ? 2 : 0) | (s.y == R0 ? 4 : 0) | (s.y + BR * s.z >= ze ? 8 : 0);
            if (e != s.w) return 8;
            if (!pass) {   // inner-inner: 8 rows / 2 RP columns inside the streamed region on every side
                if (s.y - BR < R0 || zend + BR > ze || s.x - 2 * RP < C0 || xend + 2 * RP > xe) return 9;
            }
        }
    if (cells != r.stream_cells) return 10;
    for (const ThinTile& t : r.thin) {
        if ((G.padL + t.x0) % 4) return 11;                                  // float4 groups of the tile
        const int rows = t.kind == 0 ? 4 : 64, cols = t.kind == 0 ? kTX : 8;
        if (t.zend - t.z0 > rows || t.zend <= t.z0) return 12;
        for (int z = t.z0; z < t.zend; ++z)
            for (int x = t.x0; x < t.x0 + cols; ++x)
                if (x >= t.xbeg && x < t.xend && !mark(z, x, 3)) return 13;
    }
    for (int z = 0; z < G.NZ; ++z)
        for (int x = 0; x < G.NX; ++x) {
            const bool interior = z >= N2 && z < G.NZ - N2 && x >= N2 && x < G.NX - N2;
            const bool streamed = z >= R0 && z < ze && x >= C0 && x < xe;
            const int o = own[(size_t)z * G.NX + x];
            if (!interior && o) return 14;
            if (interior && !o) return 15;
            if (interior && (streamed ? o == 3 : o != 3)) return 16;
        }
    return 0;
}
int main()
{
    const int grids[][3] = {{2301, 751, 10}, {700, 280, 10}, {700, 200, 10}, {4096, 4096, 12}, {20000, 320, 10}, {677, 210, 10},
                            {401, 97, 5}, {394, 100, 16}, {385, 64, 10}, {300, 50, 10}, {1024, 33, 10}};
    const int segs[] = {2, 4, 6, 10, 16, 24, 32, 1000};
    for (auto& g : grids)
        for (int sb : segs) {
            const int rc = check(g[0], g[1], g[2], sb);
            if (rc) { std::printf("grid %d x %d N2 %d seg_blocks %d: check %d\n", g[0], g[1], g[2], sb, rc); return 1; }
        }
    std::printf("ok\n");
    return 0;
}
"""


def test_stream_regions(tmp_path):
    """make_stream_regions (rtm_stream.cuh, used by prepare_classes): segments + thin-frame tiles partition the interior."""
    nvcc = shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not on PATH")
    src = tmp_path / "streamgeo.cu"
    src.write_text(STREAM_SRC)
    exe = tmp_path / "streamgeo"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-I",
           str(ROOT / "rtm_gpu_b200" / "csrc"), "-I", str(ROOT / "include"), "-o", str(exe), str(src), "-lcudart_static", "-ldl",
           "-lpthread", "-lrt"]
    cc = subprocess.run(cmd, capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr
