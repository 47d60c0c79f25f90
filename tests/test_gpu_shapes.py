"""Parity at the benchmark shapes and across operator radii (all through the C ABI, all
bit-exact against the oracle in reference-CUDA-build mode).  Time axes are short so that the
CPU oracle finishes in seconds; the spatial extents are the real ones, which is what exercises
the tile decomposition (partial tiles, ring tile counts, every RP template)."""
import dataclasses

import numpy as np
import pytest

import oraclelib as O
import rtm_gpu_b200 as R
from refcase import Case, rel_l2

pytestmark = pytest.mark.gpu


def layered(case, vtop=1500.0, grad=0.9, lens=True):
    z = np.arange(case.mod_NZ, dtype=np.float64)[None, :]
    x = np.arange(case.mod_NX, dtype=np.float64)[:, None]
    v = vtop + grad * z * case.hz / 4.0 + 0.05 * x
    v = v + 300.0 * (z > 0.45 * case.mod_NZ + 0.03 * x)
    if lens:
        v = np.where(((x - 0.6 * case.mod_NX) / (0.1 * case.mod_NX)) ** 2 + ((z - 0.6 * case.mod_NZ) / (0.15 * case.mod_NZ)) ** 2 < 1, 4200.0, v)
    return np.rint(np.clip(v, 1500.0, 4500.0)).astype(np.float32)


def traces(case, nshots):
    k = np.arange(case.NT, dtype=np.float32)[None, None, :]
    i = np.arange(case.n, dtype=np.float32)[None, :, None]
    s = np.arange(nshots, dtype=np.float32)[:, None, None]
    d = np.sin(0.3 * k + 0.01 * i + s) * np.exp(-((k - 0.5 * case.NT) / (0.3 * case.NT)) ** 2)
    d[:, ::7, :] = 0.0  # dead traces: zero samples are NOT imposed (kernel.cu:349-353)
    return d.astype(np.float32)


def run_case(case, r_u, r_x, snaps):
    vel = layered(case)
    v = R.pad_velocity(vel, case.N2, case.ifv)
    vmin, vmax, nvel, need = R.velocity_bins(v, case.dv)
    if case.iLSTE == 0:
        hzx = float(np.float32(case.hz) / np.float32(case.h))
        _, M, Index, c = R.ls_operator(case.nthita, case.nfdmax, case.nfdmin, nvel, case.tao, case.h, case.df,
                                       case.eps, case.fmax, vmin, case.dv, hzx, need)
    else:
        Index, c = None, R.taylor_operator(case.nfdmax)
    seis = traces(case, len(r_u))
    with R.engine_for_case(case, max_batch=len(r_u)) as e:
        e.set_model(v, vmin, vmax, case.dv)
        e.set_operator(c, Index)
        gather, so = e.forward(r_u, r_x, snaps=snaps)
        up, down, stable = e.migrate(r_u, r_x, seis)
    p = O.make_params(case, vmin, vmax, contract=1)
    for m in range(len(r_u)):
        og, l0, l1, osn = O.forward(p, v, c, Index, r_u[m], r_x[m], snaps=snaps)
        for i, k in enumerate(snaps):
            assert np.array_equal(so[m, i], osn[i]), f"shot {m} slot {k}: rel-L2 {rel_l2(so[m, i], osn[i]):.2e}"
        assert np.isfinite(osn[-1]).all(), "test configuration is unstable"
        assert np.array_equal(gather[m], og)
        ou, od, _, _, ost = O.migrate_shot(p, v, c, Index, r_u[m], r_x[m], seis[m])
        assert np.array_equal(up[m], ou), f"shot {m} up: rel-L2 {rel_l2(up[m], ou):.2e}"
        assert np.array_equal(down[m], od) and stable[m] == np.float32(ost)


def test_marmousi_width_taylor8():
    """BASELINE configs[1] grid (2301 x 751, dx 4 m, 8th order), 14 time slots, 2 shots."""
    case = Case(name="c2", nfdmax=4, nfdmin=2, N2=10, f0=20.0, iLSTE=1, hz=4.0, h=4.0, tao=4e-4, tao1=4e-4,
                mod_NZ=751, mod_NX=2301, NT1=14, s_l=1, s_z=3, n=2301, ds=1, r_x=1, nrec=2, NX_ED=2301, NZ_ED=751)
    run_case(case, [11, 11], [60, 2200], snaps=(2, 7, 13))


def test_rvsp_shape_adaptive():
    """configs[2] shape: 677 x 210, h = hz = 20, adaptive operator 2..10, ring width 10."""
    case = Case(name="c3", nfdmax=10, nfdmin=2, N2=10, f0=15.0, fmax=31.0, iLSTE=0, hz=20.0, h=20.0, tao=1e-3,
                tao1=1e-3, mod_NZ=210, mod_NX=677, NT1=40, s_l=21, s_z=3, n=130, ds=5, r_x=11, nrec=3,
                NX_ED=676, NZ_ED=210, nthita=200)
    run_case(case, [19, 40, 120], [20, 20, 20], snaps=(2, 20, 39))


@pytest.mark.parametrize("radius,N2", [(12, 12), (8, 10), (14, 16), (5, 9), (1, 3)])
def test_taylor_radii(radius, N2):
    """configs[4] style radius sweep through the fixed-length operator: covers the RP = 4, 8, 12,
    16 kernel templates (and radii that are not a multiple of four)."""
    case = Case(name="radius", nfdmax=radius, nfdmin=2, N2=N2, f0=15.0, iLSTE=1, hz=10.0, h=10.0, tao=4e-4,
                tao1=4e-4, mod_NZ=150, mod_NX=333, NT1=30, s_l=3, s_z=2, n=60, ds=5, r_x=100, nrec=2,
                NX_ED=333, NZ_ED=150)
    run_case(case, [N2 + 20, N2 + 70], [N2 + 50, N2 + 250], snaps=(2, 15, 29))


@pytest.mark.parametrize("nfdmax,N2,fmax", [(12, 12, 34.0), (7, 9, 28.0), (16, 16, 38.0)])
def test_adaptive_lengths(nfdmax, N2, fmax):
    """Adaptive operator with per-cell lengths spread over 2..nfdmax (h = 20 m, 1 ms): every cell of
    a 4-cell group may have its own length; RP = 8, 12, 16 templates of the adaptive path."""
    case = Case(name="adaptive", nfdmax=nfdmax, nfdmin=2, N2=N2, f0=15.0, fmax=fmax, iLSTE=0, hz=20.0, h=20.0,
                tao=1e-3, tao1=1e-3, mod_NZ=150, mod_NX=333, NT1=30, s_l=3, s_z=2, n=60, ds=5, r_x=100, nrec=2,
                NX_ED=333, NZ_ED=150, nthita=100, dv=1.0)
    run_case(case, [N2 + 20, N2 + 70], [N2 + 50, N2 + 250], snaps=(2, 15, 29))


def test_source_inside_the_ring_and_non_compensated():
    case = Case(name="ringsrc", nfdmax=4, nfdmin=2, N2=10, iLSTE=1, iCompen=0, hz=10.0, h=10.0, tao=5e-4,
                tao1=5e-4, mod_NZ=90, mod_NX=200, NT1=40, s_l=3, s_z=1, n=40, ds=5, r_x=3, nrec=2,
                NX_ED=200, NZ_ED=90)
    run_case(case, [9, 4], [12, 5], snaps=(2, 3, 39))  # (N2-1, x) and a cell deep in the ring


def test_tall_narrow_grid():
    """NZ > NX and a grid narrower than one ring tile: each band is a single clipped tile holding both
    corner squares, the sides take several tiles, and the left/right edge formula's velocity index
    (kernel.cu:128/:136, a flat (N2-l)*NX + row) wraps over several rows of the model."""
    case = Case(name="tall", nfdmax=4, nfdmin=2, N2=10, iLSTE=1, iCompen=1, hz=10.0, h=10.0, tao=5e-4,
                tao1=5e-4, mod_NZ=400, mod_NX=60, NT1=40, s_l=1, s_z=2, n=12, ds=5, r_x=30, nrec=2,
                NX_ED=60, NZ_ED=400)
    run_case(case, [14, 405], [66, 12], snaps=(2, 20, 39))  # 4-5 cells from the top/right and bottom/left ring


def test_tall_narrow_grid_adaptive():
    """The same shape with the adaptive operator (ring tiles read the global operator tables)."""
    case = Case(name="tall_ls", nfdmax=8, nfdmin=2, N2=10, f0=15.0, fmax=31.0, iLSTE=0, hz=20.0, h=20.0, tao=1e-3,
                tao1=1e-3, mod_NZ=300, mod_NX=70, NT1=30, s_l=3, s_z=2, n=12, ds=5, r_x=30, nrec=2,
                NX_ED=70, NZ_ED=300, nthita=100, dv=1.0)
    run_case(case, [13, 306], [76, 13], snaps=(2, 15, 29))
