"""The engine against the REFERENCE'S OWN CUDA BUILD on the same B200.

oracle/_ref/ref_cuda is the unmodified kernel.cu compiled by nvcc for sm_100a with the
reference's default flags (oracle/Makefile); it travels to the GPU box prebuilt.  It is run
as a black box on the reference's file surface; its per-shot images and stacked image are
the ground truth of north_star ("images matching the reference within 1e-5 relative L2").
Bar: rel-L2 <= 1e-5, and in fact bit-exact (the engine keeps the reference build's FP32
operation order and FMA contraction)."""
import shutil
import tempfile
from pathlib import Path

import numpy as np
import pytest

import rtm_gpu_b200 as R
from golden_cases import GOLDEN_CASES
from refcase import (REF_DIR, data_tiny, read_final_image, read_shot_images, rel_l2, run_reference,
                     velocity_tiny, write_inputs)
from test_gpu_parity import make_engine, prepare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_images_match_reference_cuda_build(name):
    if not (REF_DIR / "ref_cuda").exists():
        pytest.skip("oracle/_ref/ref_cuda was not built (make -C oracle ref)")
    case = GOLDEN_CASES[name]
    wd = Path(tempfile.mkdtemp(prefix="rtm_refcuda_"))
    try:
        data = {d: data_tiny(case, d) for d in case.depths}
        out = write_inputs(case, wd, velocity_tiny(case), data)
        run_reference(wd, "ref_cuda")
        ups, downs = read_shot_images(case, out)
        final = read_final_image(case, out)
    finally:
        shutil.rmtree(wd, ignore_errors=True)

    v, vmin, vmax, Index, c = prepare(case)
    seis = np.stack([data[d] for d in case.depths])
    with make_engine(case, v, vmin, vmax, Index, c, max_batch=2) as e:
        up, down, stable = e.migrate(case.r_u, [case.r_x0] * case.nrec, seis)
        su, sd, _ = e.stack_get()
    for m in range(case.nrec):
        eu, ed = rel_l2(up[m], ups[m]), rel_l2(down[m], downs[m])
        assert eu <= 1e-5 and ed <= 1e-5, (m, eu, ed)
        assert np.array_equal(up[m], ups[m]) and np.array_equal(down[m], downs[m]), (m, eu, ed)
    img, _ = R.stack_finalize(su, sd, case.nrec, case.iNorm)
    if case.ifv == 0:
        win = img[case.NX_BG:case.NX_ED, case.NZ_BG:case.NZ_ED]
        assert rel_l2(win, final) <= 1e-5
        assert np.array_equal(win, final, equal_nan=True)
