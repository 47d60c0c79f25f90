#!/usr/bin/env python
"""Per-kernel SASS evidence from the in-tree library (no GPU needed): TMA loads / prefetches / stores,
mbarrier operations, cp.async, shared and global loads/stores, FP64 work.
    python tools/sass_summary.py [rtm_gpu_b200/librtm_b200.so] > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "rtm_gpu_b200/librtm_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
counts, name = collections.defaultdict(collections.Counter), None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and name:
        counts[name][m.group(1)] += 1
        counts[name]["instr"] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
cols = ["instr", "UTMALDG", "UTMAPF", "UTMASTG", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG", "F2F", "FP64"]
print(f"# {so}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a).  UTMALDG = cp.async.bulk.tensor load (TMA),")
print("# UTMAPF = cp.async.bulk.prefetch.tensor, UTMASTG = TMA store, SYNCS = mbarrier ops, LDGSTS = cp.async, FP64 = DFMA+DADD+DMUL")
print(f"{'kernel':<44}" + "".join(f"{c:>8}" for c in cols))
rows = []
for mangled, nm in zip(counts, names):
    short = re.sub(r"\(.*", "", nm.replace("(anonymous namespace)::", "")).replace("void ", "").replace("rtmk::", "")
    c = counts[mangled]
    c["FP64"] = c["DFMA"] + c["DADD"] + c["DMUL"]
    rows.append((short, [c[k] for k in cols]))
for short, vals in sorted(rows):
    print(f"{short:<44}" + "".join(f"{v:>8}" for v in vals))
