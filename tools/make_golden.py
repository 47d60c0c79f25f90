#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (oracle/_ref, built from
/root/reference by `make -C oracle ref`) on small synthetic cases.

Only runs in the build container (needs oracle/_ref); the committed .npz files are what
travels.  Inputs are regenerated deterministically from tests/refcase.py, so a fixture
holds only reference OUTPUTS:
  up_<m>, down_<m>   per-shot images (RVSP_RTM_up_/down_<m>.dat)         kernel.cu:951-990
  final              stacked/normalised window (RVSP_Migration_Real_new2.dat)  :1061-1084
  stable             whitening constants printed per shot                       :971-972
  gather_<m>         forward field sampled at the data positions per step (shim tap after
                     Hybrid3, kernel.cu:819; the reference writes no gathers itself)
  snap_last1_<m>, snap_last0_<m>, rel1_<m>, rel2_<m>
                     the four device->host copies per shot (:822-823, :932-933)
  M, Index, c        operator table from the reference's funMandC / order
                     (LSMOrCon_rec_2D.cpp:22, :526), via oracle/_ref/libref_host.so
Usage: python tools/make_golden.py [case ...]
"""
import re
import shutil
import sys
import tempfile
from dataclasses import asdict
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
import oraclelib as O  # noqa: E402
from refcase import (Case, data_tiny, read_final_image, read_shot_images, run_reference,  # noqa: E402
                     velocity_tiny, write_inputs)
from golden_cases import CPU_GOLDEN_CASES, GOLDEN_CASES  # noqa: E402


def build(case: Case, outpath: Path):
    wd = Path(tempfile.mkdtemp(prefix="rtm_golden_", dir=str(ROOT / "gpurun_out")))
    try:
        vel = velocity_tiny(case)
        data = {d: data_tiny(case, d) for d in case.depths}
        out = write_inputs(case, wd, vel, data)
        stdout = run_reference(wd, "ref_cpu", env={"RTM_SHIM_DUMP_D2H": str(wd / "d2h.bin"),
                                                   "RTM_SHIM_GATHER": str(wd / "gather_")})
        ups, downs = read_shot_images(case, out)
        res = {"final": read_final_image(case, out)}
        stables = [float(x) for x in re.findall(r"^(\d+\.\d{16})\s*$", stdout, re.M)]
        assert len(stables) == case.nrec, stdout[-500:]
        res["stable"] = np.array(stables, np.float64)
        d2h = np.fromfile(wd / "d2h.bin", np.float32).reshape(case.nrec, 4, case.NZ, case.NX)
        for m in range(case.nrec):
            res[f"up_{m}"] = ups[m]
            res[f"down_{m}"] = downs[m]
            res[f"gather_{m}"] = np.fromfile(wd / f"gather_{m+1}.bin", np.float32).reshape(case.n, case.NT)
        # snapshots / raw accumulators: shot 0 only (size); interior of the accumulators only
        N2 = case.N2
        res["snap_last1_0"] = d2h[0, 0]
        res["snap_last0_0"] = d2h[0, 1]
        res["rel1_0"] = d2h[0, 2][N2:-N2, N2:-N2].copy()
        res["rel2_0"] = d2h[0, 3][N2:-N2, N2:-N2].copy()
        # operator table from the reference's own host code
        v, _ = O.pad_velocity(vel, case.N2, case.ifv, case.tao, case.h)
        vmin, vmax, nvel, need = O.velocity_bins(v, case.dv)
        m = re.search(r"vmin=([\d.]+)\s+vmax=([\d.]+)\s+nvel=(\d+)", stdout)
        assert (float(m.group(1)), float(m.group(2)), int(m.group(3))) == (vmin, vmax, nvel)
        if case.iLSTE == 0:
            NC, M, Index, c = O.ref_funMandC(case.nthita, case.nfdmax, case.nfdmin, nvel, case.tao,
                                             case.h, case.df, case.eps, case.fmax, vmin, vmax,
                                             case.dv, need, np.float32(case.hz) / np.float32(case.h))
            res["M"], res["Index"], res["c"] = M, Index, c
        else:
            c = np.zeros(case.nfdmax + 1, np.float32)
            O.refhost().ref_order(2 * case.nfdmax, c.ctypes.data_as(O.fp))
            res["c"] = c
        res["vrange"] = np.array([vmin, vmax, nvel], np.float64)
        np.savez_compressed(outpath, **res)
        print(f"{case.name}: wrote {outpath} ({outpath.stat().st_size/1024:.0f} KiB)")
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def build_resample(outpath: Path):
    """resample() (Resample.cpp:193-225) on a damped sine, four rate changes."""
    k = np.arange(400, dtype=np.float64)
    yin = (np.sin(0.07 * k) * np.exp(-((k - 180) / 90.0) ** 2)).astype(np.float32)
    cfg = [(0.002, 799, 0.001), (0.001, 160, 0.0025), (0.001, 400, 0.001), (0.004, 1201, 0.00133)]
    res = {"yin": yin, "dxin": np.array([c[0] for c in cfg], np.float32),
           "nxout": np.array([c[1] for c in cfg]), "dxout": np.array([c[2] for c in cfg], np.float32)}
    for i, (dxin, nxout, dxout) in enumerate(cfg):
        out = np.zeros(nxout, np.float32)
        O.refhost().ref_resample(len(yin), dxin, yin.ctypes.data_as(O.fp), nxout, dxout,
                                 out.ctypes.data_as(O.fp))
        res[f"yout_{i}"] = out
    np.savez_compressed(outpath, **res)
    print(f"resample: wrote {outpath}")


def segy_test_values():
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.standard_normal(200) * 10.0 ** rng.integers(-30, 30, 200), [0.0, -0.0, 1.0, -1.0, 0.1, 3.4e38,
                        -3.4e38, 1e-40, 7.2e75 % 1e38, 16.0, 15.999999, 1.0 / 16, 255.5, -4096.25]])
    return x.astype(np.float32)


def build_segy(outpath: Path):
    """trace2segy / segy2trace (segy.cpp:653-695) for the four sample formats."""
    import ctypes as C
    x = segy_test_values()
    res = {"x": x}
    for fmt in (1, 2, 3, 5):
        xin = x if fmt in (1, 5) else np.clip(x, -3e4, 3e4).astype(np.float32)
        nb = 2 if fmt == 3 else 4
        buf = C.create_string_buffer(len(xin) * nb)
        O.refhost().ref_trace2segy(buf, xin.ctypes.data_as(O.fp), len(xin), fmt)
        back = np.zeros(len(xin), np.float32)
        O.refhost().ref_segy2trace(buf.raw, back.ctypes.data_as(O.fp), len(xin), fmt)
        res[f"in_{fmt}"] = xin
        res[f"bytes_{fmt}"] = np.frombuffer(buf.raw, np.uint8).copy()
        res[f"back_{fmt}"] = back
    np.savez_compressed(outpath, **res)
    print(f"segy: wrote {outpath}")


if __name__ == "__main__":
    if sys.argv[1:] == ["segy"]:
        build_segy(ROOT / "tests" / "golden" / "segy.npz")
        sys.exit(0)
    if sys.argv[1:] == ["resample"]:
        build_resample(ROOT / "tests" / "golden" / "resample.npz")
        sys.exit(0)
    cases = {**GOLDEN_CASES, **CPU_GOLDEN_CASES}
    want = sys.argv[1:] or list(cases)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    for name in want:
        build(cases[name], ROOT / "tests" / "golden" / f"{name}.npz")
    if not sys.argv[1:]:
        build_resample(ROOT / "tests" / "golden" / "resample.npz")
