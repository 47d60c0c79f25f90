#!/usr/bin/env python
"""A whole job through the drop-in executable rtm_gpu_b200/rtm_b200 (one host thread per GPU, reader / writer threads,
one in-process NCCL reduce), files in -> files out, for BASELINE.json configs[1] (C2: 2301 x 751, 8th order, NT 7501,
64 shots) or configs[2] (C3: RVSP 677 x 210, adaptive 2..10, NT 3501, 240 shots).  Writes the reference's input files
into a scratch directory, runs the executable with --timing and prints one JSON line (shots/hour of the shot loop from
the executable's own timing line, and of the whole program by wall clock).
  python tools/run_driver_job.py --config c2|c3 [--gpus N] [--shots M] [--batch B] [--scratch DIR]"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from refcase import Case, write_inputs  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2", choices=["c2", "c3"])
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--shots", type=int, default=0)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--scratch", default=None)
a = ap.parse_args()

w = bench.make_workload(a.config)
shots = a.shots or w.total_shots
if a.config == "c3":
    depths = [200.0 + 20.0 * (i % 190) + (i // 190) for i in range(shots)]   # distinct file names, rows inside the model
    r_x = 11
else:
    depths = [float(8 + i) for i in range(shots)]   # 8, 9, 10 m ...: distinct file names, source rows 2, 2, 2, 2, 3, ... (int(depth)/hz)
    r_x = 1150
case = Case(name=a.config, nfdmax=w.nfdmax, nfdmin=w.nfdmin, N2=w.N2, f0=w.f0, fmax=w.fmax, df=1.0, nthita=w.nthita, eps=w.eps,
            dv=w.dv, iLSTE=w.iLSTE, ifv=0, whitecoe=w.whitecoe, hz=w.hz, tao=w.tao, iNorm=w.iNorm, iCompen=w.iCompen, angle=90.0,
            NX_BG=0, NX_ED=w.mod_NX, NZ_BG=0, NZ_ED=w.mod_NZ, h=w.h, tao1=w.tao1, mod_NZ=w.mod_NZ, mod_NX=w.mod_NX, NT1=w.NT1,
            s_l=w.s_l, s_z=w.s_z, n=w.n, ds=w.ds, r_x=r_x, nrec=shots, dr=1, depths=depths)
vel = w.velocity()
one = np.empty((1, w.n, w.NT), np.float32)
w.traces(one, 0)
wd = Path(tempfile.mkdtemp(prefix=f"rtm_{a.config}_", dir=a.scratch))
try:
    t_in = time.perf_counter()
    out = write_inputs(case, wd, vel, {})
    for j, d in enumerate(depths):   # (one trace file per shot; scaled copies keep the generation cheap)
        (one[0] * np.float32(1.0 + 0.001 * j)).tofile(wd / "in" / ("NEW_L10-1932-X_%d.dat" % int(d)))
    t_in = time.perf_counter() - t_in
    cmd = [str(ROOT / "rtm_gpu_b200" / "rtm_b200"), "--quiet", "--timing"]
    if a.gpus:
        cmd += ["--gpus", str(a.gpus)]
    if a.batch:
        cmd += ["--batch", str(a.batch)]
    env = dict(os.environ)
    try:
        import nvidia.nccl
        env["RTM_NCCL_LIB"] = str(Path(nvidia.nccl.__path__[0]) / "lib" / "libnccl.so.2")
    except Exception:
        pass
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=str(wd), capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    ok = p.returncode == 0 and (out / "RVSP_Migration_Real_new2.dat").exists()
    img = np.fromfile(out / "RVSP_Migration_Real_new2.dat", np.float32) if ok else np.zeros(1)
    timing = [l for l in p.stdout.splitlines() if l.startswith("rtm_b200 timing")]
    m = re.search(r"batch (\d+): ([0-9.]+) s = ([0-9.]+) shots/hour", " ".join(timing))
    cu = shots * w.cell_updates_per_shot()
    print(json.dumps({"config": w.name + ", drop-in executable rtm_b200 (files in -> files out)", "shots": shots, "gpus": a.gpus or "all",
                      "ok": ok, "input_files_written_s": t_in,
                      "shot_loop_seconds": float(m.group(2)) if m else None, "batch": int(m.group(1)) if m else None,
                      "shots_per_hour_shot_loop": float(m.group(3)) if m else None,
                      "Mcell_updates_per_s_shot_loop": cu / float(m.group(2)) / 1e6 if m else None,
                      "wall_seconds_whole_program": dt, "shots_per_hour_whole_program": shots / dt * 3600,
                      "image_finite": bool(np.isfinite(img).all()), "image_l2": float(np.linalg.norm(img.astype(np.float64))),
                      "timing": timing, "stderr": p.stderr[-300:]}))
finally:
    shutil.rmtree(wd, ignore_errors=True)
