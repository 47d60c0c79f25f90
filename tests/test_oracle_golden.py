"""The oracle (oracle/rtm_oracle.c, contract=0) against the reference's own outputs
(tests/golden/*.npz, produced by tools/make_golden.py from oracle/_ref/ref_cpu).
Bar: bit-exact (the oracle restates the same FP32 operation order)."""
import numpy as np
import pytest

import oraclelib as O
from golden_cases import CPU_GOLDEN_CASES, GOLDEN_CASES

ALL_CASES = {**GOLDEN_CASES, **CPU_GOLDEN_CASES}
from refcase import ROOT, data_tiny, velocity_tiny


def load(name):
    return np.load(ROOT / "tests" / "golden" / f"{name}.npz")


def setup(case, g):
    vel = velocity_tiny(case)
    v, r1 = O.pad_velocity(vel, case.N2, case.ifv, case.tao, case.h)
    vmin, vmax, nvel, need = O.velocity_bins(v, case.dv)
    assert (vmin, vmax, nvel) == tuple(g["vrange"])
    c = g["c"]
    Index = g["Index"] if case.iLSTE == 0 else None
    return v, vmin, vmax, c, Index


@pytest.mark.parametrize("name", list(ALL_CASES))
def test_migrate_bit_exact(name):
    case, g = ALL_CASES[name], load(name)
    v, vmin, vmax, c, Index = setup(case, g)
    p = O.make_params(case, vmin, vmax, contract=0)
    ups, downs = [], []
    for m, (depth, ru) in enumerate(zip(case.depths, case.r_u)):
        up, down, rel1, rel2, stable = O.migrate_shot(p, v, c, Index, ru, case.r_x0,
                                                      data_tiny(case, depth))
        assert np.array_equal(up, g[f"up_{m}"]), f"up_{m}"
        assert np.array_equal(down, g[f"down_{m}"]), f"down_{m}"
        assert "%.16f" % stable == "%.16f" % g["stable"][m]
        if m == 0:
            assert np.array_equal(rel1, g["rel1_0"])
            assert np.array_equal(rel2, g["rel2_0"])
        ups.append(up)
        downs.append(down)
    img, _ = O.stack(ups, downs, case.iNorm)
    # window (kernel.cu:1061-1108); ifv==1 reverses the trace order and shifts by one
    if case.ifv == 1:
        win = img[case.NX_ED:case.NX_BG:-1, case.NZ_BG:case.NZ_ED] if case.NX_ED < case.mod_NX else None
    else:
        win = img[case.NX_BG:case.NX_ED, case.NZ_BG:case.NZ_ED]
    if win is not None:
        assert np.array_equal(win, g["final"], equal_nan=True)


@pytest.mark.parametrize("name", list(ALL_CASES))
def test_forward_gather_and_snapshots(name):
    case, g = ALL_CASES[name], load(name)
    v, vmin, vmax, c, Index = setup(case, g)
    p = O.make_params(case, vmin, vmax, contract=0)
    for m, ru in enumerate(case.r_u):
        gather, last0, last1, _ = O.forward(p, v, c, Index, ru, case.r_x0)
        # the tap records slots 2..NT-1; slots 0,1 are the initial conditions
        assert np.array_equal(gather[:, 2:], g[f"gather_{m}"][:, 2:])
        if m == 0:
            assert np.array_equal(last1, g["snap_last1_0"])
            assert np.array_equal(last0, g["snap_last0_0"])


def test_taylor_coefficients():
    g = load("tiny_te_compen")
    assert np.array_equal(O.taylor(4), g["c"])
    # SURVEY 4.3 known answers for order(8, c)
    np.testing.assert_allclose(O.taylor(4), [-2.84722209, 1.60000002, -0.200000003,
                                             0.0253968257, -0.0017857143], rtol=1e-7)


def test_contract_mode_differs_at_fp_noise_level():
    """contract=1 (nvcc's FMA pattern) must differ from contract=0 only at the FP32 noise
    floor the reference itself shows between FMA on/off builds (~1e-5)."""
    case, g = GOLDEN_CASES["tiny_te_compen"], load("tiny_te_compen")
    v, vmin, vmax, c, Index = setup(case, g)
    p1 = O.make_params(case, vmin, vmax, contract=1)
    _, _, last1, _ = O.forward(p1, v, c, Index, case.r_u[0], case.r_x0, want_gather=False)
    ref = g["snap_last1_0"]
    err = np.linalg.norm((last1 - ref).astype(np.float64)) / np.linalg.norm(ref.astype(np.float64))
    assert 0 < err < 1e-4
