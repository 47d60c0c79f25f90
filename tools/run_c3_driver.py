#!/usr/bin/env python
"""BASELINE.json configs[2]: RVSP geometry (677 x 210, h = hz = 20 m, adaptive operator 2..10, hybrid
ABC width 10, NT = 3501 at 1 ms), 240 virtual sources in a well, sharded over the GPUs of the box
by the drop-in driver executable (one host thread per GPU, one NCCL reduce).  Writes the reference's
input files into a scratch directory, runs rtm_gpu_b200/rtm_b200 and reports shots/hour.
  python tools/run_c3_driver.py [--gpus N] [--shots 240]"""
import argparse
import json
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

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--shots", type=int, default=240)
ap.add_argument("--batch", type=int, default=0)
a = ap.parse_args()

case = Case(name="c3", nfdmax=10, nfdmin=2, N2=10, f0=15.0, fmax=31.0, df=1.0, nthita=1000, eps=1e-5, dv=1.0,
            iLSTE=0, ifv=0, whitecoe=1e-4, hz=20.0, tao=0.001, iNorm=1, iCompen=1, angle=90.0,
            NX_BG=0, NX_ED=676, NZ_BG=0, NZ_ED=210, h=20.0, tao1=0.001, mod_NZ=210, mod_NX=677, NT1=3501,
            s_l=21, s_z=3, n=130, ds=5, r_x=11, nrec=a.shots, dr=1,
            depths=[200.0 + 20.0 * (i % 190) + (i // 190) for i in range(a.shots)])
x = np.arange(case.mod_NX, dtype=np.float64)[:, None] * case.h
z = np.arange(case.mod_NZ, dtype=np.float64)[None, :] * case.hz
v = 1500.0 + 0.55 * z + 0.02 * x + 300.0 * (z > 1500.0 + 0.05 * x) + 500.0 * (z > 3000.0 - 0.03 * x)
vel = np.rint(np.clip(v, 1500.0, 4500.0)).astype(np.float32)
k = np.arange(case.NT1, dtype=np.float32)[None, :]
i = np.arange(case.n, dtype=np.float32)[:, None]
base = (np.sin(0.03 * k + 0.05 * i) * np.exp(-((k - 1400.0 - 3.0 * i) / 600.0) ** 2)).astype(np.float32)
wd = Path(tempfile.mkdtemp(prefix="rtm_c3_"))
try:
    data = {d: base * np.float32(1.0 + 0.001 * j) for j, d in enumerate(case.depths)}
    out = write_inputs(case, wd, vel, data)
    cmd = [str(ROOT / "rtm_gpu_b200" / "rtm_b200"), "--quiet", "--timing"]
    if a.gpus:
        cmd += ["--gpus", str(a.gpus)]
    if a.batch:
        cmd += ["--batch", str(a.batch)]
    env = None
    try:
        import os
        import nvidia.nccl
        env = dict(os.environ, RTM_NCCL_LIB=str(Path(nvidia.nccl.__path__[0]) / "lib" / "libnccl.so.2"))
    except Exception:
        pass
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=str(wd), capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    ok = p.returncode == 0 and (out / "RVSP_Migration_Real_new2.dat").exists()
    img = np.fromfile(out / "RVSP_Migration_Real_new2.dat", np.float32) if ok else np.zeros(1)
    NT = case.NT
    cu = a.shots * (NT - 2) * (2.0 * case.NZ * case.NX + case.mod_NZ * case.mod_NX)
    print(json.dumps({"config": "C3 RVSP 677x210 adaptive 2..10, driver executable", "shots": a.shots, "gpus": a.gpus or "all",
                      "ok": ok, "wall_seconds_whole_program": dt, "shots_per_hour_whole_program": a.shots / dt * 3600,
                      "Mcell_updates_per_s_whole_program": cu / dt / 1e6, "image_finite": bool(np.isfinite(img).all()),
                      "image_l2": float(np.linalg.norm(img.astype(np.float64))), "timing": [l for l in p.stdout.splitlines() if l.startswith("rtm_b200 timing")], "stderr": p.stderr[-300:]}))
finally:
    shutil.rmtree(wd, ignore_errors=True)
