#!/usr/bin/env python
"""Throughput of the time loop on the other BASELINE.json configurations (C3, C4, C5; the
headline C2 is bench.py).  Prints one JSON line per configuration; results are copied into
profiles/.  Inputs are resident on the device (the `value` convention of bench.py)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import rtm_gpu_b200 as R  # noqa: E402

PEAK = 6549.8
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def model(mod_NX, mod_NZ, h, hz):
    x = np.arange(mod_NX, dtype=np.float64)[:, None] * h
    z = np.arange(mod_NZ, dtype=np.float64)[None, :] * hz
    zmax, xmax = mod_NZ * hz, mod_NX * h
    v = 1500.0 + 2400.0 * z / zmax + 200.0 * x / xmax
    for f, dip, dvel in ((0.25, 0.04, 250.0), (0.5, -0.06, 350.0), (0.75, 0.03, 400.0)):
        v = v + dvel * (z > f * zmax + dip * x)
    lens = ((x - 0.55 * xmax) / (0.12 * xmax)) ** 2 + ((z - 0.6 * zmax) / (0.12 * zmax)) ** 2 < 1.0
    return np.rint(np.clip(np.where(lens, 4300.0, v), 1500.0, 4500.0)).astype(np.float32)


def run(name, mod_NX, mod_NZ, N2, nfdmax, nfdmin, iLSTE, NT, h, tao, f0, fmax, batch, n, ds, reps=2, eps=1e-5, flags=0):
    vel = model(mod_NX, mod_NZ, h, h)
    v = R.pad_velocity(vel, N2, 0)
    vmin, vmax, nvel, need = R.velocity_bins(v, 1.0)
    t0 = time.time()
    if iLSTE == 0:
        _, M, Index, c = R.ls_operator(200, nfdmax, nfdmin, nvel, tao, h, 1.0, eps, fmax, vmin, 1.0, 1.0, need)
        mhist = np.bincount(M[M >= 0], minlength=nfdmax + 1).tolist()
    else:
        Index, c, mhist = None, R.taylor_operator(nfdmax), None
    t_op = time.time() - t0
    NZ, NX = mod_NZ + 2 * N2, mod_NX + 2 * N2
    eng = R.Engine(0, mod_NZ=mod_NZ, mod_NX=mod_NX, N2=N2, nfdmax=nfdmax, NT=NT, iLSTE=iLSTE, iCompen=1, h=h, hz=h,
                   tao=tao, f0=f0, whitecoe=1e-4, s_l=N2, s_z=N2 + 2, n=n, ds=ds, max_batch=batch, flags=flags)
    store = eng.store_all_active()
    eng.set_model(v, vmin, vmax, 1.0)
    eng.set_operator(c, Index)
    k = np.arange(NT, dtype=np.float32)[None, None, :]
    i = np.arange(n, dtype=np.float32)[None, :, None]
    seis = (np.sin(0.02 * k + 0.003 * i) * np.exp(-((k - 0.4 * NT) / (0.2 * NT)) ** 2)).astype(np.float32)
    seis = np.repeat(seis, batch, axis=0)
    eng.upload_gathers(seis)
    r_u = np.full(batch, N2 + 2, np.int32)
    r_x = (N2 + np.linspace(0.1 * mod_NX, 0.9 * mod_NX, batch)).astype(np.int32)
    eng.migrate_resident(r_u, r_x)  # warm-up
    eng.reset_stats()
    for _ in range(reps):
        eng.migrate_resident(r_u, r_x)
    st = eng.stats()
    eng.close()
    steps = (NT - 2) * reps
    fwd_us = 1e6 * st["forward_seconds"] / steps
    bwd_us = 1e6 * st["backward_seconds"] / steps
    cells = NZ * NX * batch
    bwd_bytes = 52.0 if store else 60.0
    out = {"config": name, "store_all": store, "grid": [mod_NX, mod_NZ], "N2": N2, "operator": "taylor" if iLSTE else "adaptive",
           "nfdmax": nfdmax, "nfdmin": nfdmin, "length_histogram": mhist, "NT": NT, "batch": batch,
           "Mcell_updates_per_s": st["cell_updates"] / st["device_seconds"] / 1e6,
           "fwd_us": fwd_us, "fwd_GBps": 16.0 * cells / fwd_us / 1e3, "fwd_frac": 16.0 * cells / fwd_us / 1e3 / PEAK,
           "bwd_us": bwd_us, "bwd_GBps": bwd_bytes * cells / bwd_us / 1e3, "bwd_frac": bwd_bytes * cells / bwd_us / 1e3 / PEAK,
           "shots_per_hour": reps * batch / st["device_seconds"] * 3600, "operator_seconds": t_op, "peak_GBps": PEAK}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c3", "c5", "c4"])
    a = ap.parse_args()
    if "c3" in a.which:  # RVSP shape, adaptive 2..10, 240 shots sharded: 30 per GPU per batch
        run("C3 RVSP 677x210 adaptive", 677, 210, 10, 10, 2, 0, 3501, 20.0, 1e-3, 15.0, 31.0, 30, 130, 5)
    if "c5" in a.which:  # 4096^2, radius sweep through both operator paths
        for R_ in (4, 8, 12):
            run(f"C5 4096^2 taylor R={R_}", 4096, 4096, 12, R_, 2, 1, 120, 10.0, 5e-4, 15.0, 31.0, 1, 4096, 1)
        run("C5 4096^2 adaptive 2..12", 4096, 4096, 12, 12, 2, 0, 120, 20.0, 1e-3, 15.0, 34.0, 1, 4096, 1)
        for R_ in (4, 8, 12):
            run(f"C5 4096^2 adaptive forced R={R_}", 4096, 4096, 12, R_, R_, 0, 120, 20.0, 1e-3, 15.0, 34.0, 1, 4096, 1)
    if "store" in a.which:  # C2 grid, 2 shots per launch: reconstruction vs store-all (NT 3000: 2 x 3000 x 7.3 MB = 44 GB)
        for fl in (0, 1):
            run("C2 2301x751 taylor R=4, 2 shots, NT 3000" + (" STORE_ALL" if fl else ""), 2301, 751, 10, 4, 2, 1, 3000, 4.0, 4e-4, 20.0, 50.0, 2, 2301, 1, reps=1, flags=fl)
    if "c5a" in a.which:  # adaptive-path tuning subset
        run("C5 4096^2 adaptive 2..12", 4096, 4096, 12, 12, 2, 0, 120, 20.0, 1e-3, 15.0, 34.0, 1, 4096, 1)
        run("C5 4096^2 adaptive forced R=4", 4096, 4096, 12, 4, 4, 0, 120, 20.0, 1e-3, 15.0, 34.0, 1, 4096, 1)
        run("C5 2048x1024 adaptive 2..4 (M mostly 2-3)", 2048, 1024, 10, 4, 2, 0, 200, 10.0, 1e-3, 15.0, 31.0, 4, 2048, 1)
    if "c4" in a.which:  # 20000 x 5000, boundary-save reconstruction; 1000 of the 10000 steps
        run("C4 20000x5000 adaptive 2..10 (NT 1000 of 10000)", 20000, 5000, 10, 10, 2, 0, 1000, 10.0, 1e-3, 15.0, 31.0, 1, 4000, 5, reps=1)
