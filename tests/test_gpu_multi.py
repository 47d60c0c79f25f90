"""Multi-GPU (needs >= 2 devices; skipped otherwise): shots sharded over GPUs, one reduce of the
stacked images.  Only the summation order differs from the single-GPU run."""
import dataclasses
import os
import shutil
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

import rtm_gpu_b200 as R
from golden_cases import GOLDEN_CASES
from refcase import ROOT, data_tiny, read_final_image, read_shot_images, rel_l2, velocity_tiny, write_inputs
from test_gpu_parity import make_engine, prepare

pytestmark = pytest.mark.gpu


def _ndev():
    return R.lib().rtm_device_count()


def _nccl_path():
    try:
        import nvidia.nccl
        p = Path(nvidia.nccl.__path__[0]) / "lib" / "libnccl.so.2"
        return str(p) if p.exists() else None
    except Exception:
        return None


@pytest.mark.parametrize("backend", ["nccl", "p2p"])
def test_in_process_stack_reduce(backend, monkeypatch):
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    import torch  # noqa: F401  (loads the bundled libnccl.so.2 into the process)
    monkeypatch.setenv("RTM_REDUCE", backend)
    case = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], NT1=100)
    v, vmin, vmax, Index, c = prepare(case)
    r_u, r_x = [24, 34, 40, 55], [20, 31, 64, 100]
    seis = np.stack([data_tiny(case, 100 * i)[:, :100] for i in range(4)])
    with make_engine(case, v, vmin, vmax, Index, c, max_batch=4) as e:
        e.migrate(r_u, r_x, seis)
        su1, sd1, n1 = e.stack_get()
    e0 = make_engine(case, v, vmin, vmax, Index, c, max_batch=2, device=0)
    e1 = make_engine(case, v, vmin, vmax, Index, c, max_batch=2, device=1)
    e0.migrate(r_u[:2], r_x[:2], seis[:2])
    e1.migrate(r_u[2:], r_x[2:], seis[2:])
    su, sd, n, used = R.stack_reduce([e0, e1])
    e0.close(); e1.close()
    assert n == 4 and n1 == 4 and used == backend
    assert rel_l2(su, su1) < 1e-6 and rel_l2(sd, sd1) < 1e-6


def test_driver_two_gpus_matches_one_gpu():
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    case = dataclasses.replace(GOLDEN_CASES["tiny_ls_compen"], nrec=3, depths=[300.0, 500.0, 700.0], NT1=200)
    outs = {}
    for ngpu in (1, 2):
        wd = Path(tempfile.mkdtemp(prefix="rtm_drv_"))
        try:
            data = {d: data_tiny(case, d) for d in case.depths}
            out = write_inputs(case, wd, velocity_tiny(case), data)
            env = dict(os.environ)
            if _nccl_path():
                env["RTM_NCCL_LIB"] = _nccl_path()
            p = subprocess.run([str(ROOT / "rtm_gpu_b200" / "rtm_b200"), "--gpus", str(ngpu), "--quiet"], cwd=str(wd),
                               capture_output=True, text=True, env=env, timeout=600)
            assert p.returncode == 0, p.stderr[-2000:]
            outs[ngpu] = (read_shot_images(case, out), read_final_image(case, out))
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    (u1, d1), f1 = outs[1]
    (u2, d2), f2 = outs[2]
    for m in range(case.nrec):
        assert np.array_equal(u1[m], u2[m]) and np.array_equal(d1[m], d2[m])  # per-shot images: same bits
    assert rel_l2(f2, f1) < 1e-6  # stack: summation order only
