#!/usr/bin/env python
"""bench.py -- the RTM hot path on B200, measured (see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): Marmousi-shaped synthetic model 2301 x 751, dx = dz = 4 m,
8th-order fixed (Taylor) operator, hybrid ABC width 10, compensated imaging, dt = 0.4 ms,
NT = 7501 (3 s), 2301 data traces at the surface, 64 virtual sources.  A "step" is the full
migration (forward modelling with boundary-strip saving, reverse-time reconstruction +
receiver back-propagation + imaging, per-shot image filter and stacking) of one batch of
SHOTS_PER_STEP = 32 shots (two steps cover the 64 shots; the default K=4 steps migrate them twice).

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

SHOTS_PER_STEP = 32   # shots per launch: 8 -> 267, 16 -> 279, 32 -> 286 Gcell-updates/s at NT=301 (profiles/README.md)
TOTAL_SHOTS = 64


class Workload:
    """BASELINE.json configs[1]."""
    name = "marmousi_2301x751_o8_64shots"
    mod_NX, mod_NZ = 2301, 751
    N2, nfdmax, nfdmin = 10, 4, 2
    h = hz = 4.0
    tao = tao1 = 4.0e-4
    NT1 = 7501
    f0 = 20.0
    iLSTE, iCompen, iNorm = 1, 1, 1
    dv = 1.0
    whitecoe = 1.0e-4
    s_l, s_z, n, ds = 1, 3, 2301, 1      # 1-based, as in Parameter.txt
    src_depth_m = 8.0                    # virtual sources at depth index 2

    def __init__(self, scale_nt: int | None = None):
        if scale_nt:
            self.NT1 = scale_nt
        self.NT = int(np.float64(np.float32(np.float32(self.NT1 - 1) * np.float32(self.tao1)) / np.float32(self.tao)) + 1.5)
        self.NZ, self.NX = self.mod_NZ + 2 * self.N2, self.mod_NX + 2 * self.N2

    def velocity(self) -> np.ndarray:
        """[mod_NX][mod_NZ] float32, integer-valued: gradient + 3 dipping reflectors + lens."""
        x = np.arange(self.mod_NX, dtype=np.float64)[:, None] * self.h
        z = np.arange(self.mod_NZ, dtype=np.float64)[None, :] * self.hz
        v = 1500.0 + 0.6 * z + 0.02 * x
        for z0, dip, dvel in ((700.0, 0.05, 250.0), (1500.0, -0.08, 400.0), (2300.0, 0.03, 600.0)):
            v = v + dvel * (z > z0 + dip * x)
        lens = ((x - 5200.0) / 900.0) ** 2 + ((z - 1800.0) / 300.0) ** 2 < 1.0
        v = np.where(lens, 4300.0, v)
        return np.rint(np.clip(v, 1500.0, 4500.0)).astype(np.float32)

    def sources(self, first: int, count: int):
        """(r_u, r_x) of virtual sources first..first+count-1 of the 64, padded 0-based."""
        xs = np.linspace(40, self.mod_NX - 41, TOTAL_SHOTS).astype(np.int32)
        idx = (np.arange(first, first + count) % TOTAL_SHOTS)
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
def run_reference_cpu(w: Workload, nproc: int, nt_sample: int, binary: str = "ref_cpu_fast"):
    """The reference's own kernels on the host (oracle/_ref/ref_cpu_fast, built from
    /root/reference by oracle/Makefile): one process per shot, `nproc` processes at once,
    each migrating one shot of the workload's grid with NT = nt_sample time slots.
    With binary="ref_cuda" the same procedure times the reference's own CUDA build (one process,
    one GPU: it is single-GPU and serial over shots)."""
    import dataclasses
    from refcase import REF_DIR, Case, write_inputs
    exe = REF_DIR / binary
    kind = "reference"
    if not exe.exists():
        return None
    case = Case(name="bench", nfdmax=w.nfdmax, nfdmin=w.nfdmin, N2=w.N2, f0=w.f0, fmax=2.5 * w.f0, dv=w.dv,
                iLSTE=w.iLSTE, ifv=0, whitecoe=w.whitecoe, hz=w.hz, tao=w.tao, iNorm=w.iNorm, iCompen=w.iCompen,
                NX_BG=0, NX_ED=w.mod_NX, NZ_BG=0, NZ_ED=w.mod_NZ, h=w.h, tao1=w.tao1, mod_NZ=w.mod_NZ,
                mod_NX=w.mod_NX, NT1=nt_sample, s_l=w.s_l, s_z=w.s_z, n=w.n, ds=w.ds, r_x=1150, nrec=1, dr=1,
                depths=[w.src_depth_m])
    vel = w.velocity()
    k = np.arange(nt_sample, dtype=np.float32)[None, :]
    i = np.arange(w.n, dtype=np.float32)[:, None]
    data = (np.sin(0.02 * k + 0.003 * i) * np.exp(-((k - 0.4 * nt_sample) / (0.2 * nt_sample)) ** 2)).astype(np.float32)
    def timed_run(nt, nshots=1):
        depths = [w.src_depth_m + i for i in range(nshots)]
        c_nt = dataclasses.replace(case, NT1=nt, nrec=nshots, depths=depths)
        d_nt = np.ascontiguousarray(data[:, :nt])
        base = Path(tempfile.mkdtemp(prefix="rtm_refcpu_"))
        try:
            dirs = []
            for p in range(nproc):
                c = dataclasses.replace(c_nt, r_x=40 + (p * 37) % (w.mod_NX - 80))
                write_inputs(c, base / f"p{p}", vel, {d: d_nt for d in depths})
                dirs.append(base / f"p{p}")
            t0 = time.perf_counter()
            procs = [subprocess.Popen([str(exe)], cwd=str(d), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                     for d in dirs]
            for p in procs:
                p.wait()
            dt = time.perf_counter() - t0
            ok = all((d / "out" / "RVSP_RTM_up_1.dat").exists() for d in dirs)
        finally:
            shutil.rmtree(base, ignore_errors=True)
        return dt if ok else None

    # the reference's main() has a fixed cost per process (file IO, per-shot host loops, post-stack
    # stage); time it with an empty time loop (NT=3) and charge only the difference to the loop
    if binary == "ref_cuda":
        # per-shot cost of the reference's CUDA build: (3 shots) - (1 shot), whole time axis; start-up,
        # model input and the post-stack stage cancel out
        dt_fixed = timed_run(nt_sample, 1)
        dt_full = timed_run(nt_sample, 3)
        if dt_fixed is None or dt_full is None:
            return None
        dt = max(dt_full - dt_fixed, 1e-3) / 2.0
    else:
        dt_fixed = timed_run(3)
        dt_full = timed_run(nt_sample)
        if dt_fixed is None or dt_full is None:
            return None
        dt = max(dt_full - dt_fixed, 1e-3)
    cu = nproc * (nt_sample - 2) * (2.0 * w.NZ * w.NX + w.mod_NZ * w.mod_NX)
    if binary == "ref_cuda":
        return {"value": cu / dt / 1e6, "unit": "Mcell-updates/s", "kind": "reference CUDA build (unmodified kernel.cu, nvcc sm_100a)",
                "seconds_per_shot": dt, "sample": f"full {w.mod_NX}x{w.mod_NZ} grid, NT={nt_sample} of {w.NT} time slots; seconds per shot = "
                                         f"(run of 3 shots {dt_full:.2f} s - run of 1 shot {dt_fixed:.2f} s) / 2, whole main() of the unmodified reference"}
    return {"value": cu / dt / 1e6, "unit": "Mcell-updates/s", "cores": nproc, "kind": kind, "seconds": dt,
            "sample": f"{nproc} concurrent processes x 1 shot each, full {w.mod_NX}x{w.mod_NZ} grid, NT={nt_sample} of {w.NT} "
                      f"time slots; the reference's own kernels and main() run on the host through oracle/shim "
                      f"(-O3 AVX2/FMA); {dt_full:.1f} s minus {dt_fixed:.1f} s fixed cost measured with an empty time loop"}


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
    line = {"impl": "reference", "metric": "Mcell-updates/s", "value": v, "unit": "Mcell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w.name, "grid": [w.mod_NX, w.mod_NZ], "order": 2 * w.nfdmax, "NT": w.NT,
                       "step": vals[-1]["sample"]},
            "cpu_baseline": {"value": v, "unit": "Mcell-updates/s", "cores": vals[-1]["cores"], "kind": vals[-1]["kind"],
                             "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ our arm (GPU)
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shots-per-step", type=int, default=SHOTS_PER_STEP)
    ap.add_argument("--nt", type=int, default=0, help="override NT1 (debug only; the result is then not the named workload)")
    ap.add_argument("--ref-nt", type=int, default=100, help="time slots of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    w = Workload(args.nt or None)
    if args.impl == "reference":
        return reference_arm(args, w)

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
    B = args.shots_per_step

    # model + operator through the product's host code
    v = R.pad_velocity(w.velocity(), w.N2, 0)
    vmin, vmax, nvel, _ = R.velocity_bins(v, w.dv)
    coef = R.taylor_operator(w.nfdmax)
    eng = R.Engine(local, mod_NZ=w.mod_NZ, mod_NX=w.mod_NX, N2=w.N2, nfdmax=w.nfdmax, NT=w.NT, iLSTE=w.iLSTE,
                   iCompen=w.iCompen, h=w.h, hz=w.hz, tao=w.tao, f0=w.f0, whitecoe=w.whitecoe,
                   s_l=w.s_l + w.N2 - 1, s_z=w.s_z + w.N2 - 1, n=w.n, ds=w.ds, max_batch=B)
    eng.set_model(v, vmin, vmax, w.dv)
    eng.set_operator(coef)

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

    def step_resident(last=False):
        r_u, r_x = w.sources((step_shot[0] * world + rank) * B, B)
        step_shot[0] += 1
        eng.migrate_resident(r_u, r_x)
        if last:
            reduce_stack()

    def step_e2e(last=False):
        r_u, r_x = w.sources((step_shot[0] * world + rank) * B, B)
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

    # roofline of the dominant kernel (fused backward step), measured live by CUDA events
    peak, peak_src = measured_peak()
    bwd_launches = (w.NT - 2) * args.steps
    bwd_bytes_per_launch = 60.0 * w.NZ * w.NX * B
    bwd_ms = 1e3 * st["backward_seconds"] / bwd_launches
    achieved = bwd_bytes_per_launch / (bwd_ms * 1e-3) / 1e9
    fwd_ms = 1e3 * st["forward_seconds"] / bwd_launches
    traffic = None
    pairs = os.environ.get("RTM_FUSE2", "1") != "0"
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())  # recorded at tj["shots_per_launch"] shots per launch: scale to this run's batch
            traffic = tj.get("bwd_pair_dram_bytes_per_step" if pairs else "bwd_step_kernel_dram_bytes_per_launch")
            if traffic:
                traffic = traffic * B / float(tj.get("shots_per_launch", 8))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm",
                "kernel": ("backward time step = bwd2_step_kernel (inner tiles, two steps per pass) + bwd_step_kernel (ring + frame tiles)"
                           if pairs else "bwd_step_kernel (source reconstruction + receiver step + ABC + imaging)"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bwd_bytes_per_launch,
                "avg_launch_ms": bwd_ms,
                "note": ("per time step; pair stepping moves fewer bytes than the single-step algorithmic 60 B/cell, "
                         "so `frac` may exceed 1; `dram` = bytes that crossed HBM per step (ncu) over the same time"),
                "dram": (None if not traffic else {"bytes_per_step": traffic, "achieved": traffic / (bwd_ms * 1e-3) / 1e9,
                                                   "frac": traffic / (bwd_ms * 1e-3) / 1e9 / peak}),
                "forward_step": {"achieved": 16.0 * w.NZ * w.NX * B / (fwd_ms * 1e-3) / 1e9, "avg_launch_ms": fwd_ms,
                                 "frac": 16.0 * w.NZ * w.NX * B / (fwd_ms * 1e-3) / 1e9 / peak}}
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
                "config": {"workload": w.name if not args.nt else w.name + f"_NT{w.NT}_debug",
                           "grid": [w.mod_NX, w.mod_NZ], "padded_grid": [w.NX, w.NZ], "order": 2 * w.nfdmax,
                           "NT": w.NT, "shots_per_step_per_gpu": B, "shots_timed": args.steps * B * world,
                           "parallelism": f"shots sharded over {world} GPU(s), one NCCL reduce of the stack",
                           "l2": "inputs larger than L2 (working set per step %.0f MB)" % (9 * B * w.NZ * w.NX * 4 / 1e6)},
                "per_gpu_value": value / world,
                "shots_per_hour": args.steps * B * world / dev_s * 3600.0,
                "wall_ms_per_step": 1e3 * wall / args.steps,
                "clocks": clocks, "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches)}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_cpu_baseline:
            try:  # second reported baseline: the reference's own CUDA build on this GPU (after our run)
                eng.close()
                rc = run_reference_cpu(w, 1, w.NT, binary="ref_cuda")  # one whole shot of the workload
                if rc is not None:
                    line["ref_cuda_baseline"] = rc
            except Exception as ex:  # noqa: BLE001
                line["ref_cuda_baseline"] = {"value": None, "unavailable": str(ex)[:200]}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
