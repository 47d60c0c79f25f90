#!/usr/bin/env python
"""Placement of the ring CTAs inside a launch: RTM_RING_INTERLEAVE / RTM_RING_SPREAD variants in one
process (the knobs are read by rtm_create).  C2 grid, NT=301, 32 shots per launch, 1 warm-up + 2 timed steps.
  python tools/ring_spread_sweep.py [setting ...]     setting = il0 | s<eighths>, default: il0 s8 s7 s6 s4"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import rtm_gpu_b200 as R  # noqa: E402

B = 32
w = bench.Workload(301)
v = R.pad_velocity(w.velocity(), w.N2, 0)
vmin, vmax, nvel, _ = R.velocity_bins(v, w.dv)
coef = R.taylor_operator(w.nfdmax)
seis = np.empty((B, w.n, w.NT), np.float32)
w.traces(seis, 0)
r_u, r_x = w.sources(0, B)
for setting in sys.argv[1:] or ["il0", "s8", "s7", "s6", "s4"]:
    os.environ["RTM_RING_INTERLEAVE"] = "0" if setting == "il0" else "1"
    os.environ["RTM_RING_SPREAD"] = setting[1:] if setting[0] == "s" else "8"
    with R.Engine(0, mod_NZ=w.mod_NZ, mod_NX=w.mod_NX, N2=w.N2, nfdmax=w.nfdmax, NT=w.NT, iLSTE=w.iLSTE,
                  iCompen=w.iCompen, h=w.h, hz=w.hz, tao=w.tao, f0=w.f0, whitecoe=w.whitecoe,
                  s_l=w.s_l + w.N2 - 1, s_z=w.s_z + w.N2 - 1, n=w.n, ds=w.ds, max_batch=B) as eng:
        eng.set_model(v, vmin, vmax, w.dv)
        eng.set_operator(coef)
        eng.upload_gathers(seis)
        eng.migrate_resident(r_u, r_x)
        eng.reset_stats()
        for _ in range(2):
            eng.migrate_resident(r_u, r_x)
        st = eng.stats()
    steps = 2 * (w.NT - 2)
    print(f"{setting:4s} value {st['cell_updates'] / st['device_seconds'] / 1e6:8.0f}  fwd {1e6 * st['forward_seconds'] / steps:6.1f} us"
          f"  bwd {1e6 * st['backward_seconds'] / steps:6.1f} us", flush=True)
