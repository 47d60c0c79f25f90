#!/usr/bin/env python
"""Attribute the instructions and stall samples of one kernel to source-code regions.

ncu's source page lists SASS instructions with counters but no source lines in CSV form; nvdisasm -g
lists the same instructions with line info.  Both are in program order, so they are joined by position.

  cuobjdump -xelf rtm_engine.sm_100a.cubin rtm_gpu_b200/librtm_b200.so   # (in a scratch directory)
  nvdisasm -g rtm_engine.sm_100a.cubin > cur.sass
  ncu -i prof.ncu-rep --page source --csv > prof_src.csv
  python tools/ncu_attribution.py cur.sass prof_src.csv 'fwd_step_kernelILi4ELb0' \
         '{"ring: one-way":[585,656],"interior tiles":[750,851]}' [top_lines]

The JSON maps region names to line ranges of rtm_gpu_b200/csrc/rtm_kernels.cuh (the library must be the
build that was profiled).  Output: samples / warp instructions per region (profiles/r1_final_ring_attribution.txt).
"""
import collections
import csv
import json
import re
import sys
from pathlib import Path

SRC = Path(__file__).resolve().parents[1] / "rtm_gpu_b200" / "csrc" / "rtm_kernels.cuh"


def main():
    sass, csvf, kern, buckets = sys.argv[1], sys.argv[2], sys.argv[3], json.loads(sys.argv[4])
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    lines = open(sass).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
    cur, ins = None, []
    for l in lines[start + 1:]:
        if (l.startswith(".text.") or l.startswith(".section")) and ins:
            break
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S.*;", l):
            ins.append(cur)
    rows = list(csv.reader(open(csvf)))
    hdr, data = rows[1], rows[2:]
    if len(ins) != len(data):
        sys.exit(f"instruction counts differ ({len(ins)} in the SASS, {len(data)} in the profile): not the same build")
    i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    agg = collections.defaultdict(lambda: [0, 0])
    for c, r in zip(ins, data):
        a = agg[c if c else ("?", 0)]
        a[0] += int(r[i_s])
        a[1] += int(r[i_i])
    tot = [sum(a[i] for a in agg.values()) for i in range(2)]
    print(f"total samples {tot[0]} warp-instr {tot[1]}")

    def bucket(f, ln):
        if f != SRC.name:
            return "(headers) " + f
        for name, (lo, hi) in buckets.items():
            if lo <= ln <= hi:
                return name
        return "other"
    bag = collections.defaultdict(lambda: [0, 0])
    for (f, ln), a in agg.items():
        b = bag[bucket(f, ln)]
        b[0] += a[0]
        b[1] += a[1]
    for name, a in sorted(bag.items(), key=lambda kv: -kv[1][0]):
        print(f"{name:<44s} samples {a[0]:6d} ({100 * a[0] / tot[0]:4.1f}%)  warp-instr {a[1]:9d} ({100 * a[1] / tot[1]:4.1f}%)")
    src = SRC.read_text().splitlines()
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f, ln, a, src[ln - 1].strip()[:90] if f == SRC.name else "")


if __name__ == "__main__":
    main()
