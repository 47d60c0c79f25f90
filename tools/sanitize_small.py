"""Tiny runs of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import dataclasses
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import rtm_gpu_b200 as R  # noqa: E402
from golden_cases import GOLDEN_CASES  # noqa: E402
from refcase import data_tiny  # noqa: E402
from test_gpu_parity import make_engine, prepare  # noqa: E402

for name, flags in (("tiny_te_compen", 0), ("tiny_ls_compen", 0), ("small_aniso_flip", 0), ("tiny_ls_noncompen", R.STORE_ALL)):
    case = dataclasses.replace(GOLDEN_CASES[name], NT1=14)
    v, vmin, vmax, Index, c = prepare(case)
    seis = np.stack([data_tiny(case, d)[:, :14] for d in case.depths])
    with make_engine(case, v, vmin, vmax, Index, c, max_batch=2, flags=flags) as e:
        u, d, s = e.migrate(case.r_u, [case.r_x0] * case.nrec, seis)
        g, so = e.forward(case.r_u, [case.r_x0] * case.nrec, snaps=(2, 13))
        raw = np.stack([data_tiny(dataclasses.replace(case, NT1=9), dd) for dd in case.depths])
        e.migrate_raw(case.r_u, [case.r_x0] * case.nrec, raw, case.tao * 1.7)
    print(name, flags, float(np.abs(u).max()))
print("done")
