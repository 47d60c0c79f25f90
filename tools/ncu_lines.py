#!/usr/bin/env python
"""Per-source-line and per-opcode attribution of one kernel from an ncu source page.

  cuobjdump -xelf all rtm_gpu_b200/librtm_b200.so ; nvdisasm -g rtm_engine.sm_100a.cubin > cur.sass   (the profiled build)
  ncu -i prof.ncu-rep --page source --csv | gzip > prof_source.csv.gz
  python tools/ncu_lines.py cur.sass prof_source.csv.gz 'stream2_kernelILi4ELb1ELb1' [top]

nvdisasm -g and the source page list the same SASS in program order; they are joined by position
(like tools/ncu_attribution.py, which buckets by line ranges of one file)."""
import collections
import csv
import gzip
import io
import re
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parents[1] / "rtm_gpu_b200" / "csrc"
sass, page, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
lines = open(sass).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
cur, seq = None, []
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+", l):
        seq.append(cur)
op = gzip.open if page.endswith(".gz") else open
rows = list(csv.reader(io.TextIOWrapper(op(page, "rb"))))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
assert len(seq) == len(data), (len(seq), len(data), "the library is not the profiled build")
by_line, by_op = collections.defaultdict(lambda: [0, 0]), collections.defaultdict(lambda: [0, 0])
for loc, r in zip(seq, data):
    ni, ns = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    for d, k in ((by_line, loc), (by_op, m.group(2) if m else "?")):
        d[k][0] += ni
        d[k][1] += ns
ti, ts = sum(a[0] for a in by_op.values()), sum(a[1] for a in by_op.values())
thr = sum(int(r[ix["Thread Instructions Executed"]]) for r in data)
pon = sum(int(r[ix["Predicated-On Thread Instructions Executed"]]) for r in data)
print(f"# {rows[0][1]}\n# {len(data)} SASS instructions, {ti} warp instructions executed, {ts} stall samples, "
      f"{100 * pon / max(thr, 1):.1f} % of thread instructions predicated on")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print("# stall samples: " + ", ".join(f"{h[6:]} {100 * v / ts:.1f} %" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
print("\nopcode       warp inst   share  samples")
for k, (ni, ns) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"{k:10s} {ni:11d} {100 * ni / ti:6.1f}% {100 * ns / ts:7.1f}%")
print("\nsource line                     inst   samples")
src = {}
for (f, ln), (ni, ns) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:top]:
    if f not in src:
        src[f] = (CSRC / f).read_text().splitlines() if (CSRC / f).exists() else []
    text = src[f][ln - 1].strip()[:100] if ln <= len(src[f]) else ""
    print(f"{f}:{ln:<5d} {100 * ni / ti:6.1f}% {100 * ns / ts:7.1f}%   {text}")
