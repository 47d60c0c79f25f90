"""Parity of the CUDA engine (through the C ABI) on the B200.

Bars (FP32):
  * bit-exact against the oracle in `contract=1` mode, which restates the reference's CUDA
    build (nvcc sm_100a FMA pattern);
  * rel-L2 <= 2e-4 against the golden vectors, which come from the reference run on the host
    WITHOUT FMA contraction -- the reference's own FMA-on/off discrepancy is 1e-5..1e-3
    (SURVEY 4.3), so this second check only guards against gross errors;
  * bit-exact (hence <= 1e-5) against the reference's own CUDA build run on the same GPU
    (oracle/_ref/ref_cuda), see test_gpu_vs_ref_cuda.py.
"""
import numpy as np
import pytest

import oraclelib as O
import rtm_gpu_b200 as R
from golden_cases import GOLDEN_CASES
from refcase import ROOT, data_tiny, rel_l2, velocity_tiny

pytestmark = pytest.mark.gpu


def prepare(case):
    """Model + operator through the product's own host code."""
    v = R.pad_velocity(velocity_tiny(case), case.N2, case.ifv)
    vmin, vmax, nvel, need = R.velocity_bins(v, case.dv)
    if case.iLSTE == 0:
        hzx = float(np.float32(case.hz) / np.float32(case.h))
        _, _, Index, c = R.ls_operator(case.nthita, case.nfdmax, case.nfdmin, nvel, case.tao, case.h,
                                       case.df, case.eps, case.fmax, vmin, case.dv, hzx, need)
    else:
        Index, c = None, R.taylor_operator(case.nfdmax)
    return v, vmin, vmax, Index, c


def make_engine(case, v, vmin, vmax, Index, c, **kw):
    e = R.engine_for_case(case, **kw)
    e.set_model(v, vmin, vmax, case.dv)
    e.set_operator(c, Index)
    return e


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_migrate_bit_exact_vs_oracle(name):
    case = GOLDEN_CASES[name]
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    v, vmin, vmax, Index, c = prepare(case)
    seis = np.stack([data_tiny(case, d) for d in case.depths])
    with make_engine(case, v, vmin, vmax, Index, c, max_batch=2) as e:
        up, down, stable = e.migrate(case.r_u, [case.r_x0] * case.nrec, seis)
        su, sd, ns = e.stack_get()
    assert ns == case.nrec
    p = O.make_params(case, vmin, vmax, contract=1)
    ups, downs = [], []
    for m in range(case.nrec):
        ou, od, _, _, ostable = O.migrate_shot(p, v, c, Index, case.r_u[m], case.r_x0, seis[m])
        assert np.array_equal(up[m], ou), f"up shot {m}: rel-L2 {rel_l2(up[m], ou):.3e}"
        assert np.array_equal(down[m], od), f"down shot {m}: rel-L2 {rel_l2(down[m], od):.3e}"
        assert stable[m] == np.float32(ostable)
        # reference on the host without FMA contraction: FP-noise-level agreement only
        # (the uncompensated image is a small difference of large terms: the reference's own
        #  FMA-on/off discrepancy reaches percent level there)
        assert rel_l2(up[m], g[f"up_{m}"]) < (2e-3 if case.iCompen == 1 else 1e-1)
        assert rel_l2(down[m], g[f"down_{m}"]) < 2e-4
        ups.append(ou)
        downs.append(od)
    img, ill = R.stack_finalize(su, sd, case.nrec, case.iNorm)
    oimg, oill = O.stack(ups, downs, case.iNorm)
    assert np.array_equal(img, oimg, equal_nan=True) and np.array_equal(ill, oill)


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_forward_gathers_and_snapshots(name):
    case = GOLDEN_CASES[name]
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    v, vmin, vmax, Index, c = prepare(case)
    NT = case.NT
    snaps = (0, 1, 2, 3, 10, NT - 2, NT - 1)
    with make_engine(case, v, vmin, vmax, Index, c, max_batch=2) as e:
        gather, so = e.forward(case.r_u, [case.r_x0] * case.nrec, snaps=snaps)
    p = O.make_params(case, vmin, vmax, contract=1)
    for m in range(case.nrec):
        og, l0, l1, osn = O.forward(p, v, c, Index, case.r_u[m], case.r_x0, snaps=snaps)
        for i, k in enumerate(snaps):
            assert np.array_equal(so[m, i], osn[i]), f"slot {k} shot {m}: {rel_l2(so[m, i], osn[i]):.3e}"
        assert np.array_equal(gather[m], og)
        assert rel_l2(gather[m][:, 2:], g[f"gather_{m}"][:, 2:]) < 2e-4
    assert rel_l2(so[0, -1], g["snap_last1_0"]) < 2e-4


def test_batching_does_not_change_results():
    """Shots are independent: one by one, or four at a time in one launch, same bits."""
    import dataclasses
    case = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], NT1=150)
    v, vmin, vmax, Index, c = prepare(case)
    r_u = [24, 34, 40, 55]
    r_x = [20, 31, 64, 100]
    seis = np.stack([data_tiny(case, 100 * i)[:, :150] for i in range(4)])
    with make_engine(case, v, vmin, vmax, Index, c, max_batch=1) as e1:
        u1, d1, s1 = e1.migrate(r_u, r_x, seis)
    with make_engine(case, v, vmin, vmax, Index, c, max_batch=4) as e4:
        u4, d4, s4 = e4.migrate(r_u, r_x, seis)
    p = O.make_params(case, vmin, vmax, contract=1)
    for m in range(4):
        ou, od, *_ = O.migrate_shot(p, v, c, Index, r_u[m], r_x[m], seis[m])
        assert np.array_equal(u1[m], ou), f"batch=1 shot {m} differs from the oracle ({rel_l2(u1[m], ou):.2e})"
        assert np.array_equal(u4[m], ou), f"batch=4 shot {m} differs from the oracle ({rel_l2(u4[m], ou):.2e})"
    assert np.array_equal(u1, u4) and np.array_equal(d1, d4) and np.array_equal(s1, s4)


def test_argument_errors():
    case = GOLDEN_CASES["tiny_te_compen"]
    v, vmin, vmax, Index, c = prepare(case)
    e = R.engine_for_case(case)
    with pytest.raises(R.RtmError, match="set the model"):
        e.forward([24], [20])
    e.set_model(v, vmin, vmax, case.dv)
    with pytest.raises(R.RtmError, match="Taylor operator needs"):
        e.set_operator(c[:-1])
    e.set_operator(c)
    with pytest.raises(R.RtmError, match="outside"):
        e.forward([10_000], [20])
    e.close()
    import dataclasses
    with pytest.raises(R.RtmError, match="nfdmax"):
        R.engine_for_case(dataclasses.replace(case, nfdmax=12, N2=10))


def test_fresh_contexts_are_deterministic():
    """Regression: set-up copies/memsets must be ordered against the context's non-blocking
    stream (a plain cudaMemcpy/cudaMemset is not).  Re-creating contexts back to back used to
    let the first kernels race with the still-pending model upload."""
    import dataclasses
    case = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], NT1=60)
    v, vmin, vmax, Index, c = prepare(case)
    r_u, r_x = [24, 34, 40], [20, 31, 64]
    seis = np.stack([data_tiny(case, 100 * i)[:, :60] for i in range(3)])
    first = None
    for it in range(12):
        with make_engine(case, v, vmin, vmax, Index, c, max_batch=1 + it % 3) as e:
            u, d, s = e.migrate(r_u, r_x, seis)
        if first is None:
            first = (u, d, s)
        assert np.array_equal(u, first[0]) and np.array_equal(d, first[1]) and np.array_equal(s, first[2]), it


@pytest.mark.parametrize("name", ["tiny_te_compen", "tiny_ls_noncompen", "small_aniso_flip"])
def test_store_all_mode(name):
    """RTM_FLAG_STORE_ALL (keep the whole forward wavefield in HBM): bit-exact against the oracle's
    store-all variant, and within the reference's own reconstruction error of the default mode."""
    case = GOLDEN_CASES[name]
    v, vmin, vmax, Index, c = prepare(case)
    seis = np.stack([data_tiny(case, d) for d in case.depths])
    with make_engine(case, v, vmin, vmax, Index, c, max_batch=2, flags=R.STORE_ALL) as e:
        assert e.store_all_active()
        up, down, stable = e.migrate(case.r_u, [case.r_x0] * case.nrec, seis)
    with make_engine(case, v, vmin, vmax, Index, c, max_batch=2) as e:
        assert not e.store_all_active()
        up0, down0, _ = e.migrate(case.r_u, [case.r_x0] * case.nrec, seis)
    p = O.make_params(case, vmin, vmax, contract=1)
    for m in range(case.nrec):
        ou, od, _, _, ost = O.migrate_shot(p, v, c, Index, case.r_u[m], case.r_x0, seis[m], store_all=1)
        assert np.array_equal(up[m], ou) and np.array_equal(down[m], od) and stable[m] == np.float32(ost)
        assert rel_l2(down[m], down0[m]) < 1e-3
        assert rel_l2(up[m], up0[m]) < (2e-2 if case.iCompen == 1 else 0.5)


def test_store_all_falls_back_when_it_does_not_fit():
    import dataclasses
    big = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], mod_NX=2301, mod_NZ=751, NT1=7501, n=100, ds=5,
                              NX_ED=2301, NZ_ED=751)
    e = R.engine_for_case(big, max_batch=8, flags=R.STORE_ALL)  # 8 x 7501 x 7.3 MB = 440 GB
    assert not e.store_all_active()
    e.close()


def test_device_resampler_matches_host_routine():
    """resample() on the device (fused with the transpose) == the host routine, bit for bit; the
    host routine itself is pinned to the reference's function in tests/test_host.py."""
    import dataclasses
    case = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], tao=0.001, NT1=239)  # engine NT = 239
    rng = np.random.default_rng(11)
    with R.engine_for_case(case, max_batch=2) as e:
        for NT1, tao1 in [(120, 0.002), (477, 0.0005), (300, 0.0013), (96, 0.0025)]:
            tr = rng.standard_normal((37, NT1)).astype(np.float32)
            got = e.resample_device(tr, tao1)
            want = np.stack([R.resample(t, tao1, 239, 0.001) for t in tr])
            assert np.array_equal(got, want), (NT1, tao1)


def test_migrate_raw_equals_host_resampling_then_migrate():
    import dataclasses
    case = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], tao1=0.002, NT1=120)  # NT = 239
    v, vmin, vmax, Index, c = prepare(case)
    raw = np.stack([data_tiny(case, d) for d in case.depths])  # [2][n][120] at 2 ms
    with make_engine(case, v, vmin, vmax, Index, c, max_batch=2) as e:
        NT = e.params.NT
        u1, d1, s1 = e.migrate_raw(case.r_u, [case.r_x0] * 2, raw, case.tao1)
        host = np.stack([[R.resample(t, case.tao1, NT, case.tao) for t in shot] for shot in raw])
        u2, d2, s2 = e.migrate(case.r_u, [case.r_x0] * 2, host)
    assert np.array_equal(u1, u2) and np.array_equal(d1, d2) and np.array_equal(s1, s2)
