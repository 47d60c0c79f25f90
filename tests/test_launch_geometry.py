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
