"""A small run of the z-streaming two-step kernels (backward + forward pairs) for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_stream.py"""
import os
import sys
from pathlib import Path

os.environ.setdefault("RTM_FUSE2", "1")
os.environ.setdefault("RTM_STREAM2", "1")
os.environ.setdefault("RTM_FUSE2_FWD", "1")
os.environ.setdefault("RTM_SEG_TILES", "3")
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import rtm_gpu_b200 as R  # noqa: E402
from refcase import Case  # noqa: E402
from test_gpu_shapes import layered, traces  # noqa: E402

LS = os.environ.get("RTM_SAN_LS", "0") == "1"   # the adaptive operator (lengths 2..4) through the same kernels
if LS:
    case = Case(name="stream_ls", nfdmax=4, nfdmin=2, N2=10, f0=15.0, fmax=31.0, iLSTE=0, hz=20.0, h=20.0, tao=1e-3, tao1=1e-3,
                mod_NZ=200, mod_NX=700, NT1=12, s_l=5, s_z=40, n=230, ds=3, r_x=1, nrec=2, NX_ED=700, NZ_ED=200, nthita=100, dv=1.0)
else:
    case = Case(name="stream", nfdmax=4, nfdmin=2, N2=10, f0=20.0, iLSTE=1, hz=5.0, h=5.0, tao=5e-4, tao1=5e-4,
                mod_NZ=200, mod_NX=700, NT1=12, s_l=5, s_z=40, n=230, ds=3, r_x=1, nrec=2, NX_ED=700, NZ_ED=200)
v = R.pad_velocity(layered(case), case.N2, 0)
vmin, vmax, nvel, need = R.velocity_bins(v, case.dv)
Index, coef = None, R.taylor_operator(4)
if LS:
    _, M, Index, coef = R.ls_operator(case.nthita, case.nfdmax, case.nfdmin, nvel, case.tao, case.h, case.df, case.eps, case.fmax,
                                      vmin, case.dv, 1.0, need)
seis = traces(case, 2)
with R.engine_for_case(case, max_batch=2) as e:
    e.set_model(v, vmin, vmax, case.dv)
    e.set_operator(coef, Index)
    u, d, s = e.migrate([60, 90], [300, 420], seis)
    g, _ = e.forward([60, 90], [300, 420])
    print("launches", e.stats()["kernel_launches"], "paired cell-steps", e.stats()["pair_cell_steps_backward"], float(np.abs(u).max()), float(np.abs(g).max()))
print("done")
