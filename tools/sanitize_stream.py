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

case = Case(name="stream", nfdmax=4, nfdmin=2, N2=10, f0=20.0, iLSTE=1, hz=5.0, h=5.0, tao=5e-4, tao1=5e-4,
            mod_NZ=200, mod_NX=700, NT1=12, s_l=5, s_z=40, n=230, ds=3, r_x=1, nrec=2, NX_ED=700, NZ_ED=200)
v = R.pad_velocity(layered(case), case.N2, 0)
vmin, vmax, _, _ = R.velocity_bins(v, case.dv)
seis = traces(case, 2)
with R.engine_for_case(case, max_batch=2) as e:
    e.set_model(v, vmin, vmax, case.dv)
    e.set_operator(R.taylor_operator(4))
    u, d, s = e.migrate([60, 90], [300, 420], seis)
    g, _ = e.forward([60, 90], [300, 420])
    print("launches", e.stats()["kernel_launches"], float(np.abs(u).max()), float(np.abs(g).max()))
print("done")
