"""Host-side check of the launch geometry helpers of rtm_kernels.cuh (compiled for the host with nvcc,
no GPU needed): the division by multiply-high and the mapping of block indices to ring / interior CTAs."""
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
