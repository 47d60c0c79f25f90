"""Test infrastructure: synthetic RVSP cases in the reference's own file formats.

Writes the inputs the reference's main() reads (kernel.cu:542-604, 693-700, 827-838;
GPU_velocity_real.cpp:11-19), can run the reference binaries built by oracle/Makefile
(oracle/_ref/ref_cpu, ref_cuda) in a scratch directory, and reads back the per-shot
and stacked images.  Used by tests/, tools/make_golden.py and bench.py's reference arm.
Nothing here is product code.
"""
from __future__ import annotations

import os
import struct
import subprocess
from dataclasses import dataclass, field, asdict
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF_DIR = ROOT / "oracle" / "_ref"

# label lines are free text (the reference skips them with %[^\n]); these are ours
_LABELS = [
    "max operator length", "min operator length", "hybrid ABC width", "dominant frequency",
    "max frequency for optimal operator", "frequency interval", "number of azimuths",
    "dispersion tolerance", "velocity interval", "LSM-0 TEM-1", "flip velocity in x",
    "white coefficient", "dz", "dt", "normalise", "compensate", "nsmooth (unused)",
    "white coefficient phase (unused)", "phase angle", "trace begin", "trace end",
    "depth begin", "depth end", "seismic data dir", "velocity file", "receiver depth file",
    "parameter file", "result dir",
]


@dataclass
class Case:
    """One run configuration = 2D_Real_RVSP_RTM.txt + Parameter.txt + receiver depths."""
    name: str = "tiny"
    nfdmax: int = 10
    nfdmin: int = 2
    N2: int = 10
    f0: float = 15.0
    fmax: float = 31.0
    df: float = 1.0
    nthita: int = 1000
    eps: float = 1.0e-5
    dv: float = 1.0
    iLSTE: int = 0
    ifv: int = 0
    whitecoe: float = 1.0e-4
    hz: float = 20.0
    tao: float = 0.001
    iNorm: int = 1
    iCompen: int = 1
    Nsmooth: int = 0
    wthite_phase: float = 1.0e-3
    angle: float = 90.0
    NX_BG: int = 0
    NX_ED: int = 120
    NZ_BG: int = 0
    NZ_ED: int = 100
    # Parameter.txt (1-based grid indices)
    h: float = 20.0
    tao1: float = 0.001
    mod_NZ: int = 100
    mod_NX: int = 120
    NT1: int = 400
    s_l: int = 21
    s_z: int = 3
    n: int = 20
    ds: int = 5
    r_x: int = 11
    nrec: int = 2
    dr: int = 1
    depths: list = field(default_factory=lambda: [300.0, 500.0])

    # ---- derived (kernel.cu:607-628), float32 arithmetic where the reference uses float
    @property
    def NZ(self): return self.mod_NZ + 2 * self.N2
    @property
    def NX(self): return self.mod_NX + 2 * self.N2
    @property
    def NT(self):
        f = np.float32
        return int(np.float64(f(f(self.NT1 - 1) * f(self.tao1)) / f(self.tao)) + 1.5)
    @property
    def r_u(self):
        f = np.float32
        return [int(f(abs(f(int(d)) / f(self.hz))) + f(self.N2 - 1)) for d in self.depths]
    @property
    def s_l0(self): return self.s_l + self.N2 - 1
    @property
    def s_z0(self): return self.s_z + self.N2 - 1
    @property
    def r_x0(self): return self.r_x + self.N2 - 1


def velocity_tiny(c: Case) -> np.ndarray:
    """[mod_NX][mod_NZ] float32; layered + lateral gradient."""
    z = np.arange(c.mod_NZ)[None, :]
    x = np.arange(c.mod_NX)[:, None]
    v = np.where(z < 40, 1500.0, np.where(z < 70, 2500.0, 3500.0)) + 2.0 * x
    return v.astype(np.float32)


def data_tiny(c: Case, depth: float) -> np.ndarray:
    """[n][NT1] float32 observed traces."""
    k = np.arange(c.NT1)[None, :].astype(np.float64)
    i = np.arange(c.n)[:, None].astype(np.float64)
    d = np.sin(0.05 * k + 0.3 * i + 0.001 * int(depth)) * np.exp(-((k - 200.0) / 80.0) ** 2)
    return d.astype(np.float32)


def write_sgy_template(path: Path, ns: int = 16, fmt: int = 1):
    """A minimal SEG-Y file (3200 + 400 + one 240-byte trace header + samples) standing in
    for the reference's SGY_Model.sgy, which WriteSGY only uses as a header template
    (SGYWrite.cpp:14-37)."""
    ebc = b" " * 3200
    bh = bytearray(400)
    struct.pack_into(">h", bh, 16, 1000)   # dt (us)       bytes 3217-3218
    struct.pack_into(">h", bh, 20, ns)     # ns            bytes 3221-3222
    struct.pack_into(">h", bh, 24, fmt)    # format code   bytes 3225-3226
    th = bytearray(240)
    struct.pack_into(">h", th, 114, ns)
    struct.pack_into(">h", th, 116, 1000)
    path.write_bytes(ebc + bytes(bh) + bytes(th) + b"\0" * (4 * ns))


def write_inputs(c: Case, workdir: Path, vel: np.ndarray, data: dict, crlf: bool = True):
    """Lay out a run directory exactly as the reference expects; returns the result dir."""
    workdir = Path(workdir)
    (workdir / "in").mkdir(parents=True, exist_ok=True)
    (workdir / "out").mkdir(parents=True, exist_ok=True)
    ind, outd = str(workdir / "in") + "/", str(workdir / "out") + "/"
    vals = [c.nfdmax, c.nfdmin, c.N2, c.f0, c.fmax, c.df, c.nthita, c.eps, c.dv, c.iLSTE, c.ifv,
            c.whitecoe, c.hz, c.tao, c.iNorm, c.iCompen, c.Nsmooth, c.wthite_phase, c.angle,
            c.NX_BG, c.NX_ED, c.NZ_BG, c.NZ_ED, ind, ind + "vel.dat", ind + "Depth_Of_Receiver.txt",
            ind + "Parameter.txt", outd]
    eol = "\r\n" if crlf else "\n"
    txt = ""
    for lab, v in zip(_LABELS, vals):
        if isinstance(v, float):
            v = repr(float(np.float32(v))) if abs(v) >= 1e-3 else "%.9e" % v
        txt += lab + eol + str(v) + eol
    (workdir / "2D_Real_RVSP_RTM.txt").write_text(txt, newline="")
    par = [c.h, c.tao1, c.mod_NZ, c.mod_NX, c.NT1, c.s_l, c.s_z, c.n, c.ds, c.r_x, c.nrec, c.dr]
    (workdir / "in" / "Parameter.txt").write_text(
        " \n".join(("%.9g" % p) if isinstance(p, float) else str(p) for p in par) + "\n")
    (workdir / "in" / "Depth_Of_Receiver.txt").write_text(
        " ".join("%g" % d for d in c.depths) + " \n")
    assert vel.shape == (c.mod_NX, c.mod_NZ) and vel.dtype == np.float32
    vel.tofile(workdir / "in" / "vel.dat")
    for depth, d in data.items():
        assert d.shape == (c.n, c.NT1) and d.dtype == np.float32
        d.tofile(workdir / "in" / ("NEW_L10-1932-X_%d.dat" % int(depth)))
    write_sgy_template(workdir / "SGY_Model.sgy")
    return workdir / "out"


def run_reference(workdir: Path, which: str = "ref_cpu", env: dict | None = None,
                  timeout: float = 3600) -> str:
    """Run a reference binary (cwd = workdir, where 2D_Real_RVSP_RTM.txt lives)."""
    exe = REF_DIR / which
    if not exe.exists():
        raise FileNotFoundError(f"{exe} not built (make -C oracle ref)")
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([str(exe)], cwd=str(workdir), env=e, capture_output=True, text=True,
                       timeout=timeout, errors="replace")
    if p.returncode != 0:
        raise RuntimeError(f"{which} failed rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    return p.stdout


def read_shot_images(c: Case, outdir: Path):
    ups, downs = [], []
    for m in range(c.nrec):
        ups.append(np.fromfile(outdir / f"RVSP_RTM_up_{m+1}.dat", np.float32).reshape(c.mod_NX, c.mod_NZ))
        downs.append(np.fromfile(outdir / f"RVSP_RTM_down_{m+1}.dat", np.float32).reshape(c.mod_NX, c.mod_NZ))
    return ups, downs


def read_final_image(c: Case, outdir: Path):
    a = np.fromfile(outdir / "RVSP_Migration_Real_new2.dat", np.float32)
    return a.reshape(c.NX_ED - c.NX_BG, c.NZ_ED - c.NZ_BG)


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.linalg.norm((a - b).ravel()); n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d
