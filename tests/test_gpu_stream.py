"""The z-streaming two-step kernels (rtm_gpu_b200/csrc/rtm_stream.cuh: warp-specialised, TMA-fed row
rings, segments of a tile column) against the oracle and against the other forms of the same step:

  * bit-exact against the oracle for every segment length (RTM_SEG_TILES = 1, 2, 3, 5: 2..10 blocks per
    segment, i.e. every residue of the 3-slot ring and its copy rows), with sources and the data line
    on block / segment / tile-column borders, odd and even step counts, compensated and plain imaging;
  * identical bits from the streaming form, the tile form (RTM_STREAM2=0) and single stepping (RTM_FUSE2=0)
    at the benchmark grid width.
"""
import numpy as np
import pytest

import rtm_gpu_b200 as R
from refcase import Case
from test_gpu_shapes import layered, run_case, traces

pytestmark = pytest.mark.gpu


def env(monkeypatch, fuse=1, stream=1, seg=8, fwd=None):
    monkeypatch.setenv("RTM_FUSE2", str(fuse))
    monkeypatch.setenv("RTM_STREAM2", str(stream))
    monkeypatch.setenv("RTM_SEG_TILES", str(seg))
    if fwd is not None:
        monkeypatch.setenv("RTM_FUSE2_FWD", str(fwd))


@pytest.mark.parametrize("seg,nt1,icompen", [(1, 27, 1), (2, 30, 1), (3, 31, 0), (5, 26, 1), (8, 29, 1)])
def test_stream_vs_oracle_segment_lengths(monkeypatch, seg, nt1, icompen):
    """8th-order Taylor operator on 700 x 280 (6 x 18 tiles of 128 x 16, 4 x 16 of them inner): a tile column is
    cut into segments of `seg` tiles; sources on the first / last row of a segment and next to a column border;
    data line inside the streamed region."""
    env(monkeypatch, seg=seg)
    case = Case(name="stream", nfdmax=4, nfdmin=2, N2=10, f0=20.0, iLSTE=1, iCompen=icompen, hz=5.0, h=5.0, tao=5e-4,
                tao1=5e-4, mod_NZ=280, mod_NX=700, NT1=nt1, s_l=5, s_z=10 + 16 * 3 - 1 - 9, n=230, ds=3, r_x=1, nrec=3,
                NX_ED=700, NZ_ED=280)
    #      rows: tile rows start at z = 10 + 16 i; columns: tile columns at x = 10 + 128 j
    run_case(case, [10 + 16 * seg, 10 + 16 * 2 * seg + 15, 10 + 16 * 5 + 7], [10 + 128 * 2 - 1, 10 + 128 * 3, 10 + 300],
             snaps=(2,))


@pytest.mark.parametrize("graph", [1, 0])
def test_stream_data_line_on_block_borders(monkeypatch, graph):
    """Data line on the last row of an 8-row block and sources in the halo columns of the neighbouring segment;
    graph replay and eager launches."""
    env(monkeypatch, seg=2)
    monkeypatch.setenv("RTM_NO_GRAPH", "0" if graph else "1")
    case = Case(name="borders", nfdmax=4, nfdmin=2, N2=10, f0=20.0, iLSTE=1, hz=5.0, h=5.0, tao=5e-4,
                tao1=5e-4, mod_NZ=200, mod_NX=700, NT1=26, s_l=5, s_z=10 + 16 * 2 + 7 - 9, n=230, ds=3, r_x=1, nrec=3,
                NX_ED=700, NZ_ED=200)
    run_case(case, [10 + 16 * 3 + 8, 10 + 16 * 4 + 7, 10 + 16 * 2], [10 + 128 * 2 - 3, 10 + 128 * 3 + 2, 10 + 128 * 2 + 127],
             snaps=(2,))


def _migrate(case, vel, seis, r_u, r_x):
    v = R.pad_velocity(vel, case.N2, case.ifv)
    vmin, vmax, _, _ = R.velocity_bins(v, case.dv)
    with R.engine_for_case(case, max_batch=len(r_u)) as e:
        e.set_model(v, vmin, vmax, case.dv)
        e.set_operator(R.taylor_operator(case.nfdmax))
        out = e.migrate(r_u, r_x, seis)
        g, _ = e.forward(r_u, r_x)
        paired = e.stats()["pair_cell_steps_backward"]
    return out, g, paired


def test_stream_equals_tile_form_equals_single_steps(monkeypatch):
    """2301 x 751, 8th order, 241 time slots, 2 shots: three forms of the time loop, one set of bits."""
    case = Case(name="c2", nfdmax=4, nfdmin=2, N2=10, f0=20.0, iLSTE=1, hz=4.0, h=4.0, tao=4e-4, tao1=4e-4,
                mod_NZ=751, mod_NX=2301, NT1=241, s_l=1, s_z=3, n=2301, ds=1, r_x=1, nrec=2, NX_ED=2301, NZ_ED=751)
    vel = layered(case)
    seis = traces(case, 2)
    r_u, r_x = [12, 300], [700, 1500]
    env(monkeypatch, fuse=0, fwd=0)
    (u1, d1, s1), g1, n1 = _migrate(case, vel, seis, r_u, r_x)
    env(monkeypatch, fuse=1, stream=0, fwd=0)
    (u2, d2, s2), g2, n2 = _migrate(case, vel, seis, r_u, r_x)
    env(monkeypatch, fuse=1, stream=1, seg=8, fwd=1)
    (u3, d3, s3), g3, n3 = _migrate(case, vel, seis, r_u, r_x)
    env(monkeypatch, fuse=1, stream=1, seg=3, fwd=1)
    (u4, d4, s4), g4, n4 = _migrate(case, vel, seis, r_u, r_x)
    assert np.abs(u1).max() > 0 and np.isfinite(u1).all()
    assert np.array_equal(u1, u2) and np.array_equal(d1, d2) and np.array_equal(s1, s2)
    assert np.array_equal(u1, u3) and np.array_equal(d1, d3) and np.array_equal(s1, s3)
    assert np.array_equal(u1, u4) and np.array_equal(d1, d4) and np.array_equal(s1, s4)
    assert np.array_equal(g1, g2) and np.array_equal(g1, g3) and np.array_equal(g1, g4)
    assert n1 == 0 and n2 > 0 and n3 > n2 and n4 == n3   # cell-steps advanced two slots per pass: none / tile form / streamed region


# ---------------------------------------------------------------------------------------------------------------
# The adaptive operator in the streaming form: when no velocity bin needs an operator longer than 4 the pair of
# steps runs through stream2_kernel<4, true, LS> / thin_frame_kernel<.., LS> / ring_kernel<.., LS> (per-cell
# coefficient rows from the padded global table).
# ---------------------------------------------------------------------------------------------------------------
def _adaptive_case(name, nt1, mod_NX=700, mod_NZ=280, icompen=1, nrec=3):
    return Case(name=name, nfdmax=4, nfdmin=2, N2=10, f0=15.0, fmax=31.0, iLSTE=0, iCompen=icompen, hz=20.0, h=20.0, tao=1e-3,
                tao1=1e-3, mod_NZ=mod_NZ, mod_NX=mod_NX, NT1=nt1, s_l=5, s_z=10 + 16 * 3 - 1 - 9, n=230, ds=3, r_x=1, nrec=nrec,
                NX_ED=mod_NX, NZ_ED=mod_NZ, nthita=100, dv=1.0)


@pytest.mark.parametrize("seg,nt1,icompen", [(1, 27, 1), (3, 30, 0), (8, 29, 1)])
def test_stream_adaptive_vs_oracle(monkeypatch, seg, nt1, icompen):
    """Adaptive operator 2..4 on 700 x 280 (lengths 2, 3 and 4 inside one float4 group where the layers meet)."""
    env(monkeypatch, seg=seg)
    case = _adaptive_case("stream_ls", nt1, icompen=icompen)
    run_case(case, [10 + 16 * seg, 10 + 16 * 2 * seg + 15, 10 + 16 * 5 + 7], [10 + 128 * 2 - 1, 10 + 128 * 3, 10 + 300],
             snaps=(2,))


def test_stream_adaptive_runs_in_pairs_and_equals_single_steps(monkeypatch):
    """2301 x 751, adaptive 2..4, 121 slots: the default policy (nothing forced) picks the streaming pairs for a launch of
    4 shots; same bits as single stepping."""
    case = _adaptive_case("c2_ls", 121, mod_NX=2301, mod_NZ=751, nrec=4)
    vel = layered(case)
    v = R.pad_velocity(vel, case.N2, 0)
    vmin, vmax, nvel, need = R.velocity_bins(v, case.dv)
    hzx = float(np.float32(case.hz) / np.float32(case.h))
    _, M, Index, c = R.ls_operator(case.nthita, case.nfdmax, case.nfdmin, nvel, case.tao, case.h, case.df, case.eps, case.fmax,
                                   vmin, case.dv, hzx, need)
    assert M.max() <= 4 and len(set(M[M > 0].tolist())) >= 2, "the model should need more than one operator length"
    seis = traces(case, 4)
    r_u, r_x = [12, 300, 500, 740], [700, 1500, 30, 2290]

    def migrate():
        with R.engine_for_case(case, max_batch=4) as e:
            e.set_model(v, vmin, vmax, case.dv)
            e.set_operator(c, Index)
            out = e.migrate(r_u, r_x, seis)
            return out, e.stats()["pair_cell_steps_backward"]
    monkeypatch.delenv("RTM_FUSE2", raising=False)
    (u2, d2, s2), n2 = migrate()
    monkeypatch.setenv("RTM_FUSE2", "0")
    (u1, d1, s1), n1 = migrate()
    assert n1 == 0 and n2 > 0
    assert np.abs(u1).max() > 0 and np.isfinite(u1).all()
    assert np.array_equal(u1, u2) and np.array_equal(d1, d2) and np.array_equal(s1, s2)
