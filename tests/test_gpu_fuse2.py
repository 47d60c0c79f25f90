"""Pair stepping of the backward pass (two time steps of the inner tiles per kernel pass,
`bwd2_step_kernel`; ring and frame tiles by the single-step kernel): bit-exact against the oracle
on grids that have inner tiles, with the data line, the source and odd/even step counts placed
so that every branch of the two-step kernel is taken, and bit-identical to single stepping
(RTM_FUSE2=0) at the benchmark grid width."""
import numpy as np
import pytest

import rtm_gpu_b200 as R
from refcase import Case
from test_gpu_shapes import layered, run_case, traces

pytestmark = pytest.mark.gpu


@pytest.fixture
def fuse_env(monkeypatch):
    def set_(maxrp=4, on=1):
        monkeypatch.setenv("RTM_FUSE2", str(on))
        monkeypatch.setenv("RTM_FUSE2_MAXRP", str(maxrp))
    return set_


@pytest.mark.parametrize("nt1,s_z,icompen", [(31, 52, 1), (30, 3, 1), (24, 70, 0)])
def test_pairs_taylor8_vs_oracle(fuse_env, nt1, s_z, icompen):
    """8th-order Taylor operator, 700 x 200 (6 x 13 tiles, 4 x 11 of them inner); data line inside
    the inner tiles (s_z = 52, 70) or in the frame (3); odd and even numbers of steps; sources in an
    inner tile and in a frame tile."""
    fuse_env(4)
    case = Case(name="pairs", nfdmax=4, nfdmin=2, N2=10, f0=20.0, iLSTE=1, iCompen=icompen, hz=5.0, h=5.0, tao=5e-4,
                tao1=5e-4, mod_NZ=200, mod_NX=700, NT1=nt1, s_l=5, s_z=s_z, n=230, ds=3, r_x=1, nrec=2,
                NX_ED=700, NZ_ED=200)
    run_case(case, [10 + 75, 10 + 3], [10 + 300, 10 + 420], snaps=(2,))


@pytest.mark.parametrize("graph", [1, 0])
def test_pairs_sources_and_data_line_on_tile_edges(fuse_env, monkeypatch, graph):
    """Sources one cell inside / outside tile edges (they lie in the grown region of the neighbouring
    tiles, which inject them redundantly), data line on the last row of a tile row; replayed graph
    and eager launches (RTM_NO_GRAPH)."""
    fuse_env(4)
    monkeypatch.setenv("RTM_NO_GRAPH", "0" if graph else "1")
    case = Case(name="edges", nfdmax=4, nfdmin=2, N2=10, f0=20.0, iLSTE=1, hz=5.0, h=5.0, tao=5e-4,
                tao1=5e-4, mod_NZ=200, mod_NX=700, NT1=26, s_l=5, s_z=10 + 16 * 4 - 10, n=230, ds=3, r_x=1, nrec=3,
                NX_ED=700, NZ_ED=200)
    #            tile rows start at z = 10 + 16 i, tile columns at x = 10 + 128 j
    run_case(case, [10 + 16 * 3, 10 + 16 * 5 - 1, 10 + 16 * 2 + 15], [10 + 128 * 2 - 1, 10 + 128 * 3, 10 + 128 * 2 + 127],
             snaps=(2,))


def test_pairs_taylor16_vs_oracle(fuse_env):
    """Radius 8 (RP = 8 template of the two-step kernel, 2 CTAs per SM)."""
    fuse_env(8)
    case = Case(name="pairs8", nfdmax=8, nfdmin=2, N2=10, f0=15.0, iLSTE=1, hz=10.0, h=10.0, tao=4e-4,
                tao1=4e-4, mod_NZ=180, mod_NX=650, NT1=27, s_l=3, s_z=60, n=200, ds=3, r_x=100, nrec=2,
                NX_ED=650, NZ_ED=180)
    run_case(case, [10 + 80, 10 + 20], [10 + 300, 10 + 500], snaps=(2,))


@pytest.mark.parametrize("nfdmax,fmax,maxrp", [(4, 22.0, 4), (7, 28.0, 8)])
def test_pairs_adaptive_vs_oracle(fuse_env, nfdmax, fmax, maxrp):
    """Adaptive operator: coefficient rows staged for the tile grown by one radius."""
    fuse_env(maxrp)
    case = Case(name="pairs_ls", nfdmax=nfdmax, nfdmin=2, N2=10, f0=15.0, fmax=fmax, iLSTE=0, hz=20.0, h=20.0,
                tao=1e-3, tao1=1e-3, mod_NZ=180, mod_NX=650, NT1=28, s_l=3, s_z=60, n=200, ds=3, r_x=100, nrec=2,
                NX_ED=650, NZ_ED=180, nthita=100, dv=1.0)
    run_case(case, [10 + 80, 10 + 20], [10 + 300, 10 + 500], snaps=(2,))


def _migrate(case, vel, seis, r_u, r_x, monkeypatch, fuse, maxrp=4):
    monkeypatch.setenv("RTM_FUSE2", "1" if fuse else "0")
    monkeypatch.setenv("RTM_FUSE2_MAXRP", str(maxrp))
    v = R.pad_velocity(vel, case.N2, case.ifv)
    vmin, vmax, nvel, need = R.velocity_bins(v, case.dv)
    if case.iLSTE == 0:
        hzx = float(np.float32(case.hz) / np.float32(case.h))
        _, M, Index, c = R.ls_operator(case.nthita, case.nfdmax, case.nfdmin, nvel, case.tao, case.h, case.df,
                                       case.eps, case.fmax, vmin, case.dv, hzx, need)
    else:
        Index, c = None, R.taylor_operator(case.nfdmax)
    with R.engine_for_case(case, max_batch=len(r_u)) as e:
        e.set_model(v, vmin, vmax, case.dv)
        e.set_operator(c, Index)
        out = e.migrate(r_u, r_x, seis)
        launches = e.stats()["pair_cell_steps_backward"]
    return out, launches


def test_pairs_equal_single_steps_marmousi_width(monkeypatch):
    """2301 x 751, 8th order, 301 time slots, 2 shots: pair stepping and single stepping give the
    same images bit for bit; the pair path really ran (rtm_stats counts the cell-steps advanced
    two slots per pass)."""
    case = Case(name="c2", nfdmax=4, nfdmin=2, N2=10, f0=20.0, iLSTE=1, hz=4.0, h=4.0, tao=4e-4, tao1=4e-4,
                mod_NZ=751, mod_NX=2301, NT1=301, s_l=1, s_z=3, n=2301, ds=1, r_x=1, nrec=2, NX_ED=2301, NZ_ED=751)
    vel = layered(case)
    seis = traces(case, 2)
    r_u, r_x = [12, 300], [700, 1500]
    (u1, d1, s1), n1 = _migrate(case, vel, seis, r_u, r_x, monkeypatch, fuse=False)
    (u2, d2, s2), n2 = _migrate(case, vel, seis, r_u, r_x, monkeypatch, fuse=True)
    assert np.abs(u1).max() > 0 and np.isfinite(u1).all()
    assert np.array_equal(u1, u2) and np.array_equal(d1, d2) and np.array_equal(s1, s2)
    assert n1 == 0 and n2 > 0   # cell-steps advanced two slots per pass


def test_pairs_equal_single_steps_adaptive(monkeypatch):
    """Adaptive operator 2..8 on 1100 x 400 with a data line inside the inner tiles, 120 slots."""
    case = Case(name="ls", nfdmax=8, nfdmin=2, N2=10, f0=15.0, fmax=30.0, iLSTE=0, hz=20.0, h=20.0, tao=1e-3,
                tao1=1e-3, mod_NZ=400, mod_NX=1100, NT1=120, s_l=3, s_z=100, n=360, ds=3, r_x=100, nrec=2,
                NX_ED=1100, NZ_ED=400, nthita=100, dv=1.0)
    vel = layered(case)
    seis = traces(case, 2)
    r_u, r_x = [150, 30], [500, 900]
    (u1, d1, s1), n1 = _migrate(case, vel, seis, r_u, r_x, monkeypatch, fuse=False)
    (u2, d2, s2), n2 = _migrate(case, vel, seis, r_u, r_x, monkeypatch, fuse=True, maxrp=8)
    assert np.abs(u1).max() > 0 and np.isfinite(u1).all()
    assert np.array_equal(u1, u2) and np.array_equal(d1, d2) and np.array_equal(s1, s2)
    assert n1 == 0 and n2 > 0
