#!/usr/bin/env python
"""bench.py -- the RTM hot path on B200, measured (see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): Marmousi-shaped synthetic model 2301 x 751, dx = dz = 4 m,
8th-order fixed (Taylor) operator, hybrid ABC width 10, compensated imaging, dt = 0.4 ms,
NT = 7501 (3 s), 2301 data traces at the surface, 64 virtual sources.  A "step" is the full
migration (forward modelling with boundary-strip saving, reverse-time reconstruction +
receiver back-propagation + imaging, per-shot image filter and stacking) of one batch of
SHOTS_PER_STEP = 64 shots (one step = the configuration's 64 shots in one batch; default K=2 steps).

  value  Mcell-updates/s, whole job, inputs resident in HBM (CUDA events inside the library,
         on the stream the kernels run on; max over ranks)
  e2e    same metric through the host-buffer C-ABI call rtm_migrate(): pinned host traces in,
         per-shot images out, H2D/D2H copies inside the timed region (wall clock, synchronised)
  roofline      the backward time step (source reconstruction + receiver step + ABC + imaging):
                algorithmic bytes (60 B per grid cell per step) / average time per step
                from CUDA events / measured HBM peak.  The backward pass advances two steps per pass
                on the inner tiles (bwd2_step_kernel: 68 B per cell and PAIR of steps), so the
                algorithmic figure can exceed the peak; `dram` gives the same with the bytes that
                really crossed HBM (ncu, profiles/traffic.json)
  cpu_baseline  the reference's own kernels run on the host (oracle/_ref/ref_cpu_fast), one
                process per shot on all host cores, on a bounded sample of the same workload

`--impl reference` times only that CPU arm.  N>1: one process per GPU (torchrun), shots
sharded, weak scaling (every GPU migrates SHOTS_PER_STEP shots per step), one NCCL reduce of
the stacked images at the end of the last step.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SHOTS_PER_STEP = 64   # shots per launch (round 1: 8 -> 267, 16 -> 279, 32 -> 286; round 2: 32 -> 329.6, 48 -> 334.3, 64 -> 338.7 Gcell-updates/s at NT=301)
TOTAL_SHOTS = 64


class Workload:
    """BASELINE.json configs[1] -- the default -- and, through make_workload(), the
    other configurations BASELINE.json names (C3, C4, C5)."""
    name = "marmousi_2301x751_o8_64shots"
    key = "c2"
    mod_NX, mod_NZ = 2301, 751
    N2, nfdmax, nfdmin = 10, 4, 2
    h = hz = 4.0
    tao = tao1 = 4.0e-4
    NT1 = 7501
    f0 = 20.0
    fmax = 50.0
    nthita = 200
    eps = 1.0e-5
    iLSTE, iCompen, iNorm = 1, 1, 1
    dv = 1.0
    whitecoe = 1.0e-4
    s_l, s_z, n, ds = 1, 3, 2301, 1      # 1-based, as in Parameter.txt
    src_depth_m = 8.0                    # virtual sources at depth index 2
    total_shots = TOTAL_SHOTS
    shots_per_step = SHOTS_PER_STEP
    default_steps = 2
    vertical_sources = False             # True: RVSP geometry, sources down a well at column r_x

    def __init__(self, scale_nt: int | None = None):
        if scale_nt:
            self.NT1 = scale_nt
        self.NT = int(np.float64(np.float32(np.float32(self.NT1 - 1) * np.float32(self.tao1)) / np.float32(self.tao)) + 1.5)
        self.NZ, self.NX = self.mod_NZ + 2 * self.N2, self.mod_NX + 2 * self.N2

    def velocity(self) -> np.ndarray:
        """[mod_NX][mod_NZ] float32, integer-valued: gradient + 3 dipping reflectors + lens."""
        x = np.arange(self.mod_NX, dtype=np.float64)[:, None] * self.h
        z = np.arange(self.mod_NZ, dtype=np.float64)[None, :] * self.hz
        if self.key == "c2":
            v = 1500.0 + 0.6 * z + 0.02 * x
            for z0, dip, dvel in ((700.0, 0.05, 250.0), (1500.0, -0.08, 400.0), (2300.0, 0.03, 600.0)):
                v = v + dvel * (z > z0 + dip * x)
            lens = ((x - 5200.0) / 900.0) ** 2 + ((z - 1800.0) / 300.0) ** 2 < 1.0
            v = np.where(lens, 4300.0, v)
        else:   # the same structure scaled to the model's extent (tools/perf_configs.py)
            zmax, xmax = self.mod_NZ * self.hz, self.mod_NX * self.h
            v = 1500.0 + 2400.0 * z / zmax + 200.0 * x / xmax
            for f, dip, dvel in ((0.25, 0.04, 250.0), (0.5, -0.06, 350.0), (0.75, 0.03, 400.0)):
                v = v + dvel * (z > f * zmax + dip * x)
            lens = ((x - 0.55 * xmax) / (0.12 * xmax)) ** 2 + ((z - 0.6 * zmax) / (0.12 * zmax)) ** 2 < 1.0
            v = np.where(lens, 4300.0, v)
        return np.rint(np.clip(v, 1500.0, 4500.0)).astype(np.float32)

    def sources(self, first: int, count: int):
        """(r_u, r_x) of virtual sources first..first+count-1, padded 0-based."""
        idx = (np.arange(first, first + count) % self.total_shots)
        if self.vertical_sources:   # well receivers every cell from 200 m, at column r_x = 11
            depth = 200.0 + self.hz * idx
            r_u = (np.abs(depth.astype(np.int32) / np.float32(self.hz)).astype(np.int32) + self.N2 - 1).astype(np.int32)
            r_u = np.minimum(r_u, self.NZ - self.N2 - 2)
            return r_u, np.full(count, 11 + self.N2 - 1, np.int32)
        xs = np.linspace(40, self.mod_NX - 41, self.total_shots).astype(np.int32)
        r_x = xs[idx] + self.N2 - 1
        r_u = np.full(count, int(abs(int(self.src_depth_m) / self.hz) + self.N2 - 1), np.int32)
        return r_u, r_x.astype(np.int32)

    def traces(self, out: np.ndarray, first: int):
        """Synthetic observed data [count][n][NT] (analytic, non-zero everywhere)."""
        count = out.shape[0]
        k = np.arange(self.NT, dtype=np.float32)[None, :]
        i = np.arange(self.n, dtype=np.float32)[:, None]
        for s in range(count):
            ph = np.float32(0.37 * (first + s))
            out[s] = np.sin(0.02 * k + 0.003 * i + ph) * np.exp(-((k - 0.4 * self.NT - 0.5 * i) / (0.2 * self.NT)) ** 2)

    def cell_updates_per_shot(self) -> float:
        return (self.NT - 2) * (2.0 * self.NZ * self.NX + self.mod_NZ * self.mod_NX)

    def operator(self, R, v):
        """(vmin, vmax, Index, c, length histogram) through the product's host code."""
        vmin, vmax, nvel, need = R.velocity_bins(v, self.dv)
        if self.iLSTE == 0:
            hzx = float(np.float32(self.hz) / np.float32(self.h))
            _, M, Index, c = R.ls_operator(self.nthita, self.nfdmax, self.nfdmin, nvel, self.tao, self.h, 1.0, self.eps,
                                           self.fmax, vmin, self.dv, hzx, need)
            return vmin, vmax, Index, c, np.bincount(M[M >= 0], minlength=self.nfdmax + 1).tolist()
        return vmin, vmax, None, R.taylor_operator(self.nfdmax), None


def make_workload(spec: str, scale_nt: int | None = None) -> Workload:
    """c2 (default) | c3 | c4 | c5[:R][:taylor]   -- BASELINE.json configs[1..4]."""
    key, *opt = spec.lower().split(":")
    if key == "c2":
        return Workload(scale_nt)

    class W(Workload):
        pass
    W.key = key
    if key == "c3":     # RVSP shape per 2D_Real_RVSP_RTM.txt: 677 x 210, adaptive 2..10, 240 well receivers, 30 per launch
        W.name = "rvsp_677x210_adaptive2-10_240shots"
        W.mod_NX, W.mod_NZ, W.N2, W.nfdmax, W.nfdmin = 677, 210, 10, 10, 2
        W.h = W.hz = 20.0
        W.tao = W.tao1 = 1.0e-3
        W.NT1, W.f0, W.fmax, W.iLSTE = 3501, 15.0, 31.0, 0
        W.s_l, W.s_z, W.n, W.ds = 21, 3, 130, 5
        W.total_shots, W.shots_per_step, W.default_steps, W.vertical_sources = 240, 30, 8, True
    elif key == "c4":   # 20000 x 5000, 10 000 steps, adaptive 2..10, boundary-strip reconstruction, 8 shots (one per GPU at 8)
        W.name = "large_20000x5000_adaptive2-10_nt10000_8shots"
        W.mod_NX, W.mod_NZ, W.N2, W.nfdmax, W.nfdmin = 20000, 5000, 10, 10, 2
        W.h = W.hz = 10.0
        W.tao = W.tao1 = 1.0e-3
        W.NT1, W.f0, W.fmax, W.iLSTE = 10000, 15.0, 31.0, 0
        W.s_l, W.s_z, W.n, W.ds = 1, 3, 4000, 5
        W.total_shots, W.shots_per_step, W.default_steps = 8, 1, 1
        W.src_depth_m = 20.0
    elif key == "c5":   # 4096^2, N2 = 12, radius sweep: c5:R = adaptive path forced to radius R, c5:R:taylor, c5 = adaptive 2..12
        R_ = int(opt[0]) if opt and opt[0].isdigit() else 0
        taylor = "taylor" in opt
        W.mod_NX, W.mod_NZ, W.N2 = 4096, 4096, 12
        W.nfdmax, W.nfdmin = (R_, R_) if R_ else (12, 2)
        if taylor:
            W.h = W.hz = 10.0
            W.tao = W.tao1 = 5.0e-4
            W.iLSTE, W.nfdmin = 1, 2
        else:
            W.h = W.hz = 20.0
            W.tao = W.tao1 = 1.0e-3
            W.iLSTE = 0
        W.name = "sweep_4096x4096_" + (f"taylor_r{R_}" if taylor else (f"adaptive_forced_r{R_}" if R_ else "adaptive2-12"))
        W.NT1, W.f0, W.fmax = 500, 15.0, 34.0
        W.s_l, W.s_z, W.n, W.ds = 1, 3, 4096, 1
        W.total_shots, W.shots_per_step, W.default_steps = 4, 1, 1
        W.src_depth_m = 40.0
    else:
        raise SystemExit(f"bench.py: unknown --config {spec!r} (c2, c3, c4, c5[:R][:taylor])")
    return W(scale_nt)


# ------------------------------------------------------------------ helpers
def measured_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ reference arm (CPU)
def _ref_case(w: Workload, nt: int, r_x: int, depths):
    from refcase import Case
    return Case(name="bench", nfdmax=w.nfdmax, nfdmin=w.nfdmin, N2=w.N2, f0=w.f0, fmax=w.fmax, dv=w.dv, nthita=w.nthita,
                eps=w.eps, iLSTE=w.iLSTE, ifv=0, whitecoe=w.whitecoe, hz=w.hz, tao=w.tao, iNorm=w.iNorm, iCompen=w.iCompen,
                NX_BG=0, NX_ED=w.mod_NX, NZ_BG=0, NZ_ED=w.mod_NZ, h=w.h, tao1=w.tao1, mod_NZ=w.mod_NZ,
                mod_NX=w.mod_NX, NT1=nt, s_l=w.s_l, s_z=w.s_z, n=w.n, ds=w.ds, r_x=r_x, nrec=len(depths), dr=1,
                depths=list(depths))


def _run_ref_processes(exe, cases, vel, datas, keep_images=False):
    """One process of `exe` per case, all at once; returns (seconds, [(ups, downs)] or None)."""
    from refcase import read_shot_images, write_inputs
    base = Path(tempfile.mkdtemp(prefix="rtm_ref_"))
    try:
        dirs = []
        for p, (c, d) in enumerate(zip(cases, datas)):
            write_inputs(c, base / f"p{p}", vel, d)
            dirs.append(base / f"p{p}")
        t0 = time.perf_counter()
        procs = [subprocess.Popen([str(exe)], cwd=str(d), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d in dirs]
        for pr in procs:
            pr.wait()
        dt = time.perf_counter() - t0
        if not all((d / "out" / "RVSP_RTM_up_1.dat").exists() for d in dirs):
            return None, None
        images = [read_shot_images(c, d / "out") for c, d in zip(cases, dirs)] if keep_images else None
        return dt, images
    finally:
        shutil.rmtree(base, ignore_errors=True)


def _sample_data(w: Workload, nt: int):
    k = np.arange(nt, dtype=np.float32)[None, :]
    i = np.arange(w.n, dtype=np.float32)[:, None]
    return (np.sin(0.02 * k + 0.003 * i) * np.exp(-((k - 0.4 * nt) / (0.2 * nt)) ** 2)).astype(np.float32)


def run_reference_cpu(w: Workload, nproc: int, nt_sample: int):
    """The reference's own kernels on the host (oracle/_ref/ref_cpu_fast, built from /root/reference by
    oracle/Makefile): one process per shot, `nproc` processes at once, each migrating one shot of the
    workload's grid with NT = nt_sample time slots.  The reference's main() has a fixed cost per process
    (file IO, operator search, per-shot host loops, post-stack stage): it is timed with an empty time loop
    (NT = 3) and only the difference is charged to the loop."""
    from refcase import REF_DIR
    exe = REF_DIR / "ref_cpu_fast"
    if not exe.exists():
        return None
    vel = w.velocity()
    data = _sample_data(w, nt_sample)

    def timed(nt):
        cases = [_ref_case(w, nt, 40 + (p * 37) % max(1, w.mod_NX - 80), [w.src_depth_m]) for p in range(nproc)]
        d_nt = np.ascontiguousarray(data[:, :nt])
        return _run_ref_processes(exe, cases, vel, [{w.src_depth_m: d_nt}] * nproc)[0]
    dt_fixed = timed(3)
    dt_full = timed(nt_sample)
    if dt_fixed is None or dt_full is None:
        return None
    dt = max(dt_full - dt_fixed, 1e-3)
    cu = nproc * (nt_sample - 2) * (2.0 * w.NZ * w.NX + w.mod_NZ * w.mod_NX)
    return {"value": cu / dt / 1e6, "unit": "Mcell-updates/s", "cores": nproc, "kind": "reference", "seconds": dt,
            "sample": f"{nproc} concurrent processes x 1 shot each, full {w.mod_NX}x{w.mod_NZ} grid, NT={nt_sample} of {w.NT} "
                      f"time slots; the reference's own kernels and main() run on the host through oracle/shim "
                      f"(-O3 AVX2/FMA); {dt_full:.1f} s minus {dt_fixed:.1f} s fixed cost measured with an empty time loop"}


def run_reference_cuda(w: Workload, nt: int, shot=None):
    """The reference's own CUDA build (oracle/_ref/ref_cuda: unmodified kernel.cu, nvcc sm_100a) on this GPU.
    Seconds per shot = (run of 3 shots - run of 1 shot) / 2: start-up, model input, operator search and the
    post-stack stage cancel out.  shot = (r_x 1-based, depth in m, traces [n][nt]): the 1-shot run migrates
    exactly that shot and its images are returned for the parity check."""
    from refcase import REF_DIR
    exe = REF_DIR / "ref_cuda"
    if not exe.exists():
        return None, None
    vel = w.velocity()
    if shot is None:
        shot = (40, w.src_depth_m, _sample_data(w, nt))
    r_x, depth, data = shot
    data = np.ascontiguousarray(data, np.float32)
    dt1, img = _run_ref_processes(exe, [_ref_case(w, nt, r_x, [depth])], vel, [{depth: data}], keep_images=True)
    depths3 = [depth, depth + w.hz, depth + 2 * w.hz]
    dt3, _ = _run_ref_processes(exe, [_ref_case(w, nt, r_x, depths3)], vel, [{d: data for d in depths3}])
    if dt1 is None or dt3 is None:
        return None, None
    dt = max(dt3 - dt1, 1e-3) / 2.0
    cu = (nt - 2) * (2.0 * w.NZ * w.NX + w.mod_NZ * w.mod_NX)
    rec = {"value": cu / dt / 1e6, "unit": "Mcell-updates/s", "kind": "reference CUDA build (unmodified kernel.cu, nvcc sm_100a)",
           "seconds_per_shot": dt,
           "sample": f"full {w.mod_NX}x{w.mod_NZ} grid, NT={nt} of {w.NT} time slots; seconds per shot = (run of 3 shots {dt3:.2f} s - "
                     f"run of 1 shot {dt1:.2f} s) / 2, whole main() of the unmodified reference"}
    return rec, (img[0][0][0], img[0][1][0])


def config_dict(w: Workload, args, B, world, extra=None):
    d = {"workload": w.name if not args.nt else w.name + f"_NT{w.NT}_debug",
         "grid": [w.mod_NX, w.mod_NZ], "padded_grid": [w.NX, w.NZ], "order": 2 * w.nfdmax,
         "operator": "taylor" if w.iLSTE == 1 else f"adaptive {w.nfdmin}..{w.nfdmax}",
         "NT": w.NT, "shots_per_step_per_gpu": B, "shots_timed": args.steps * B * world,
         "parallelism": f"shots sharded over {world} GPU(s), one NCCL reduce of the stack",
         "l2": "inputs larger than L2 (working set per step %.0f MB)" % (9 * B * w.NZ * w.NX * 4 / 1e6)}
    d.update(extra or {})
    return d


def reference_arm(args, w: Workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ncores = len(os.sched_getaffinity(0))
    nt = args.ref_nt
    vals = []
    for _ in range(args.warmup):
        run_reference_cpu(w, ncores, max(8, nt // 8))
    for _ in range(args.steps):
        r = run_reference_cpu(w, ncores, nt)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_cpu_fast missing (make -C oracle ref needs /root/reference)"}))
            return 0
        vals.append(r)
    secs = sum(r["seconds"] for r in vals)
    cu = sum(r["value"] * r["seconds"] for r in vals)
    v = cu / secs
    B = args.shots_per_step or w.shots_per_step
    line = {"impl": "reference", "metric": "Mcell-updates/s", "value": v, "unit": "Mcell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(w, args, B, max(1, args.gpus), {"sample": "each step is a BOUNDED SAMPLE of this workload on the host cores: " + vals[-1]["sample"]}),
            "cpu_baseline": {"value": v, "unit": "Mcell-updates/s", "cores": vals[-1]["cores"], "kind": vals[-1]["kind"],
                             "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ in-process multi-GPU arm
def inproc_arm(args, w: Workload):
    """north_star's multi-GPU design measured directly: ONE process, one host thread and one context per GPU
    (ctypes releases the GIL inside the C ABI), shots sharded, a single in-process ncclReduce of the stacks
    (rtm_stack_reduce: ncclCommInitAll + ncclReduce).  Same metric and timing rules as the torchrun arm:
    device time per context from CUDA events, max over GPUs."""
    import torch
    import rtm_gpu_b200 as R
    N = args.gpus
    if torch.cuda.device_count() < N:
        raise SystemExit(f"bench.py --inproc: {N} GPUs requested, {torch.cuda.device_count()} visible")
    B = args.shots_per_step or w.shots_per_step
    v = R.pad_velocity(w.velocity(), w.N2, 0)
    vmin, vmax, Index, coef, _ = w.operator(R, v)
    seis = np.empty((B, w.n, w.NT), np.float32)
    w.traces(seis, 0)
    engines, results = [None] * N, [None] * N

    def setup(g):
        e = R.Engine(g, mod_NZ=w.mod_NZ, mod_NX=w.mod_NX, N2=w.N2, nfdmax=w.nfdmax, NT=w.NT, iLSTE=w.iLSTE,
                     iCompen=w.iCompen, h=w.h, hz=w.hz, tao=w.tao, f0=w.f0, whitecoe=w.whitecoe,
                     s_l=w.s_l + w.N2 - 1, s_z=w.s_z + w.N2 - 1, n=w.n, ds=w.ds, max_batch=B)
        e.set_model(v, vmin, vmax, w.dv)
        e.set_operator(coef, Index)
        e.upload_gathers(seis)
        engines[g] = e

    def work(g, nsteps, first_step):
        e = engines[g]
        for i in range(nsteps):
            r_u, r_x = w.sources(((first_step + i) * N + g) * B, B)
            e.migrate_resident(r_u, r_x)
        results[g] = e.stats()

    def par(fn, *a):
        th = [threading.Thread(target=fn, args=(g, *a)) for g in range(N)]
        [t.start() for t in th]
        [t.join() for t in th]
    par(setup)
    devs = np.arange(N, dtype=np.int32)
    R.lib().rtm_stack_reduce_prepare(R._i(devs), N)   # communicators created ahead, as the executable does next to its shot loop
    par(work, args.warmup, 0)
    R.stack_reduce(engines)                           # warm-up reduce
    for e in engines:
        e.reset_stats()
        e.stack_reset()
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    par(work, args.steps, args.warmup)
    t_shots = time.perf_counter() - t0
    up, down, nshots, backend = R.stack_reduce(engines)
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_s = max(r["device_seconds"] for r in results) + (wall - t_shots)   # the reduce follows the slowest GPU
    cu = sum(r["cell_updates"] for r in results)
    launches = sum(r["kernel_launches"] for r in results)
    assert nshots == args.steps * B * N
    line = {"metric": "Mcell-updates/s", "value": cu / dev_s / 1e6, "unit": "Mcell-updates/s", "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "b200-inproc",
            "config": config_dict(w, args, B, N, {"parallelism": f"ONE process, one host thread + one context per GPU ({N}), "
                                                  f"one in-process reduce of the stacks (backend: {backend})"}),
            "per_gpu_value": cu / dev_s / 1e6 / N, "shots_per_hour": args.steps * B * N / dev_s * 3600.0,
            "wall_ms_per_step": 1e3 * wall / args.steps, "reduce_ms": 1e3 * (wall - t_shots), "reduce_backend": backend,
            "stack_checksum": float(np.abs(up).sum()), "clocks": clocks, "gpu_launches": int(launches)}
    print(json.dumps(line))
    for e in engines:
        e.close()
    return 0


# ------------------------------------------------------------------ our arm (GPU)
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 2 for c2; per-config otherwise)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", help="c2 (default, BASELINE configs[1]) | c3 | c4 | c5[:R][:taylor]")
    ap.add_argument("--shots-per-step", type=int, default=0)
    ap.add_argument("--nt", type=int, default=0, help="override NT1 (debug only; the result is then not the named workload)")
    ap.add_argument("--ref-nt", type=int, default=100, help="time slots of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference CUDA build (second baseline + parity check)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--inproc", action="store_true", help="one process, one host thread per GPU, in-process ncclReduce (no torchrun)")
    args = ap.parse_args()
    w = make_workload(args.config, args.nt or None)
    if not args.steps:
        args.steps = w.default_steps
    if args.impl == "reference":
        return reference_arm(args, w)
    if args.inproc:
        return inproc_arm(args, w)

    import torch
    import torch.distributed as dist
    import rtm_gpu_b200 as R

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.shots_per_step or w.shots_per_step

    # model + operator through the product's host code
    v = R.pad_velocity(w.velocity(), w.N2, 0)
    t_op = time.perf_counter()
    vmin, vmax, Index, coef, mhist = w.operator(R, v)
    t_op = time.perf_counter() - t_op
    eng = R.Engine(local, mod_NZ=w.mod_NZ, mod_NX=w.mod_NX, N2=w.N2, nfdmax=w.nfdmax, NT=w.NT, iLSTE=w.iLSTE,
                   iCompen=w.iCompen, h=w.h, hz=w.hz, tao=w.tao, f0=w.f0, whitecoe=w.whitecoe,
                   s_l=w.s_l + w.N2 - 1, s_z=w.s_z + w.N2 - 1, n=w.n, ds=w.ds, max_batch=B)
    eng.set_model(v, vmin, vmax, w.dv)
    eng.set_operator(coef, Index)

    # pinned host buffers (inputs of the e2e call)
    seis_t = torch.empty((B, w.n, w.NT), dtype=torch.float32, pin_memory=True)
    seis = seis_t.numpy()
    w.traces(seis, rank * B)
    up_t = torch.empty((B, w.mod_NX, w.mod_NZ), dtype=torch.float32, pin_memory=True)
    down_t = torch.empty_like(up_t).pin_memory()
    stable = np.zeros(B, np.float32)
    L = R.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_stack():
        if world == 1:
            return
        ptr, nfl, _ = eng.stack_device()

        class _Buf:
            __cuda_array_interface__ = {"shape": (nfl,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        t = torch.as_tensor(_Buf(), device=torch.device("cuda", local))
        dist.reduce(t, dst=0)

    step_shot = [0]
    last_sources = [None]

    def step_resident(last=False):
        r_u, r_x = w.sources((step_shot[0] * world + rank) * B, B)
        step_shot[0] += 1
        eng.migrate_resident(r_u, r_x)
        if last:
            reduce_stack()

    def step_e2e(last=False):
        r_u, r_x = w.sources((step_shot[0] * world + rank) * B, B)
        last_sources[0] = ((step_shot[0] * world + rank) * B, r_u.copy(), r_x.copy())
        step_shot[0] += 1
        R._check(L.rtm_migrate(eng._h, B, R._i(r_u), R._i(r_x), seis_t.numpy().ctypes.data_as(R._fp),
                               up_t.numpy().ctypes.data_as(R._fp), down_t.numpy().ctypes.data_as(R._fp),
                               stable.ctypes.data_as(R._fp)))
        if last:
            reduce_stack()

    # ---- kernel-resident measurement
    eng.upload_gathers(seis)
    for _ in range(args.warmup):
        step_resident()
    eng.reset_stats()
    eng.stack_reset()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_resident(last=(i == args.steps - 1))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    st = eng.stats()
    dev_s = st["device_seconds"]
    tt = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_s, wall = float(tt[0]), float(tt[1])
    cu_total = st["cell_updates"] * world
    value = cu_total / dev_s / 1e6

    # roofline of the dominant kernel(s) (fused backward time step), measured live by CUDA events
    peak, peak_src = measured_peak()
    nsteps = (w.NT - 2) * args.steps                    # time steps per pass in the timed region
    bwd_bytes_per_launch = (60.0 if w.iCompen == 1 else 44.0) * w.NZ * w.NX * B
    bwd_ms = 1e3 * st["backward_seconds"] / nsteps
    achieved = bwd_bytes_per_launch / (bwd_ms * 1e-3) / 1e9
    fwd_ms = 1e3 * st["forward_seconds"] / nsteps
    fwd_bytes = 16.0 * w.NZ * w.NX * B
    pairs_b = st["pair_cell_steps_backward"] > 0
    pairs_f = st["pair_cell_steps_forward"] > 0
    ex_b = st["executed_bytes_backward"] / nsteps       # byte model of the schedule that ran (rtm_stats)
    ex_f = st["executed_bytes_forward"] / nsteps
    traffic = dram = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and w.key == "c2" and not args.nt:
        try:   # ncu dram__bytes of one backward / forward time step, recorded at tj["shots_per_launch"] shots per launch
            tj = json.loads(tf.read_text())
            # which schedule ran: the streaming form steps ~99 % of the interior in pairs, the tile form ~85 %
            pair_frac = st["pair_cell_steps_backward"] / (nsteps * B * float(w.mod_NZ * w.mod_NX))
            key = tj.get("schedule_keys", {}).get("stream" if pair_frac > 0.93 else ("pairs" if pairs_b else "single"))
            rec = tj.get(key) if key else None
            if rec:
                scale = B / float(rec.get("shots_per_launch", 8))
                traffic = rec["bwd_dram_bytes_per_step"] * scale
                dram = {"bytes_per_step": traffic, "achieved": traffic / (bwd_ms * 1e-3) / 1e9, "frac": traffic / (bwd_ms * 1e-3) / 1e9 / peak,
                        "forward_bytes_per_step": rec["fwd_dram_bytes_per_step"] * scale,
                        "forward_frac": rec["fwd_dram_bytes_per_step"] * scale / (fwd_ms * 1e-3) / 1e9 / peak,
                        "source": f"profiles/traffic.json[{key}] (ncu dram__bytes_read+write, {rec.get('shots_per_launch', 8)} shots per launch), scaled to {B}"}
        except Exception:
            traffic = dram = None
    roofline = {"bound": "hbm",
                "kernel": ("backward time step = stream2_kernel<BWD> (streamed segments, two steps per pass, TMA-fed rings) + ring_kernel + "
                           "thin_frame_kernel (absorbing ring and the cells next to it, stepped singly); tile form: bwd2_step_kernel + bwd_step_kernel"
                           if pairs_b else "bwd_step_kernel (source reconstruction + receiver step + ABC + imaging)"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bwd_bytes_per_launch,
                "avg_launch_ms": bwd_ms,
                "note": ("per time step of the batch; SURVEY 8(d) single-step byte model.  Two-step passes move fewer bytes, so `frac` "
                         "may exceed 1: `executed` repeats it with the byte model of the schedule that ran (rtm_stats), `dram` with the "
                         "bytes that crossed HBM (ncu)"),
                "executed": {"bytes_per_step": ex_b, "achieved": ex_b / (bwd_ms * 1e-3) / 1e9, "frac": ex_b / (bwd_ms * 1e-3) / 1e9 / peak,
                             "pair_fraction_of_cell_steps": st["pair_cell_steps_backward"] / (nsteps * B * float(w.mod_NZ * w.mod_NX))},
                "dram": dram,
                "forward_step": {"achieved": fwd_bytes / (fwd_ms * 1e-3) / 1e9, "avg_launch_ms": fwd_ms,
                                 "frac": fwd_bytes / (fwd_ms * 1e-3) / 1e9 / peak,
                                 "kernel": "stream2_kernel<FWD> + fwd_step_kernel (ring + frame tiles)" if pairs_f else "fwd_step_kernel",
                                 "executed": {"bytes_per_step": ex_f, "achieved": ex_f / (fwd_ms * 1e-3) / 1e9,
                                              "frac": ex_f / (fwd_ms * 1e-3) / 1e9 / peak}}}
    launches = st["kernel_launches"]

    # ---- end to end through the host-buffer ABI
    e2e = None
    if not args.no_e2e:
        step_shot[0] = 0
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_e2e(last=(i == args.steps - 1))
        barrier()
        e2e_s = time.perf_counter() - t0
        tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt[0])
        e2e = {"value": args.steps * B * world * w.cell_updates_per_shot() / e2e_s / 1e6, "unit": "Mcell-updates/s",
               "h2d_bytes_per_step": int(seis.nbytes), "d2h_bytes_per_step": int(2 * up_t.numel() * 4 + stable.nbytes),
               "ms_per_step": 1e3 * e2e_s / args.steps, "shots_per_hour": args.steps * B * world / e2e_s * 3600.0}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = run_reference_cpu(w, len(os.sched_getaffinity(0)), args.ref_nt)
            if cpu is None:
                cpu = {"value": None, "unit": "Mcell-updates/s", "cores": 0, "kind": "reference",
                       "sample": "unavailable: oracle/_ref/ref_cpu_fast missing"}
        line = {"metric": "Mcell-updates/s", "value": value, "unit": "Mcell-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": config_dict(w, args, B, world, {"operator_length_histogram": mhist, "operator_seconds": t_op} if mhist else None),
                "per_gpu_value": value / world,
                "shots_per_hour": args.steps * B * world / dev_s * 3600.0,
                "wall_ms_per_step": 1e3 * wall / args.steps,
                "clocks": clocks, "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches)}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_cpu_baseline and not args.no_ref_cuda:
            # second reported baseline: the reference's own CUDA build on this GPU, after our run; its 1-shot run
            # migrates shot 0 of the last timed e2e step, whose images are compared with ours bit for bit
            est_s = 4 * w.cell_updates_per_shot() / 10.0e9     # ~10 Gcell-updates/s, 4 shots
            try:
                if est_s > 240:
                    line["ref_cuda_baseline"] = {"value": None, "unavailable": f"skipped: 4 reference shots at this size need ~{est_s / 60:.0f} min "
                                                 "(parity at this shape: tests/test_gpu_named_configs.py)"}
                    line["parity_checked"] = None
                else:
                    shot = None
                    if e2e is not None and last_sources[0] is not None and not w.vertical_sources:
                        first, r_u, r_x = last_sources[0]
                        shot = (int(r_x[0]) - w.N2 + 1, w.src_depth_m, seis[0])
                    elif e2e is not None and last_sources[0] is not None:
                        first, r_u, r_x = last_sources[0]
                        shot = (11, 200.0 + w.hz * (first % w.total_shots), seis[0])
                    ours = (up_t.numpy()[0].copy(), down_t.numpy()[0].copy()) if shot is not None else None
                    eng.close()
                    rc, img = run_reference_cuda(w, w.NT, shot)
                    if rc is not None:
                        line["ref_cuda_baseline"] = rc
                    if ours is not None and img is not None:
                        same = bool(np.array_equal(ours[0], img[0]) and np.array_equal(ours[1], img[1]))
                        den = float(np.linalg.norm(img[0].astype(np.float64)))
                        line["parity_checked"] = same
                        line["parity"] = {"against": "oracle/_ref/ref_cuda (reference CUDA build) on this GPU, after the timed region",
                                          "what": "up/down images of shot 0 of the last timed e2e step, all %d time slots" % w.NT,
                                          "bit_exact": same,
                                          "rel_l2_up": float(np.linalg.norm((ours[0].astype(np.float64) - img[0]))) / den if den > 0 else None}
            except Exception as ex:  # noqa: BLE001
                line["ref_cuda_baseline"] = {"value": None, "unavailable": str(ex)[:200]}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
