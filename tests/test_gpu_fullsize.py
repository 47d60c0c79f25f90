"""BASELINE.json configs[1] at its FULL size (2301 x 751, 8th order, NT = 7501) through
size-independent properties -- the CPU oracle would need hours there:
  * determinism and batching invariance: a shot migrated alone or inside a batch, twice: same bits;
  * the imaging condition fed by the reverse-time RECONSTRUCTED source field (boundary strips,
    the reference's scheme) agrees with the one fed by the STORED forward field (55 GB resident,
    RTM_FLAG_STORE_ALL) to the accuracy of the reconstruction itself;
  * translation invariance: in a laterally invariant model a source moved by one tile width
    produces the same wavefield moved by one tile width, bit for bit, until the absorbing ring
    is reached (exercises every tile alignment of the real grid).
"""
import numpy as np
import pytest

import rtm_gpu_b200 as R
from refcase import rel_l2

pytestmark = pytest.mark.gpu

MOD_NX, MOD_NZ, N2, NT = 2301, 751, 10, 7501


def engine(max_batch, flags=0, NT_=NT, n=2301):
    return R.Engine(0, mod_NZ=MOD_NZ, mod_NX=MOD_NX, N2=N2, nfdmax=4, NT=NT_, iLSTE=1, iCompen=1, h=4.0, hz=4.0,
                    tao=4.0e-4, f0=20.0, whitecoe=1.0e-4, s_l=N2, s_z=N2 + 2, n=n, ds=1, max_batch=max_batch,
                    flags=flags)


def model(lateral=True):
    x = np.arange(MOD_NX, dtype=np.float64)[:, None] * 4.0
    z = np.arange(MOD_NZ, dtype=np.float64)[None, :] * 4.0
    v = 1500.0 + 0.6 * z + (0.02 * x if lateral else 0.0 * x)
    v = v + 300.0 * (z > 1200.0 + (0.05 * x if lateral else 0.0))
    return np.rint(np.clip(v, 1500.0, 4500.0)).astype(np.float32)


def traces(nshots, n=2301):
    k = np.arange(NT, dtype=np.float32)[None, :]
    i = np.arange(n, dtype=np.float32)[:, None]
    one = (np.sin(0.02 * k + 0.003 * i) * np.exp(-((k - 3000.0 - 0.5 * i) / 1500.0) ** 2)).astype(np.float32)
    return np.stack([one * (1.0 + 0.1 * s) for s in range(nshots)])


def setup(e, lateral=True):
    v = R.pad_velocity(model(lateral), N2, 0)
    vmin, vmax, _, _ = R.velocity_bins(v, 1.0)
    e.set_model(v, vmin, vmax, 1.0)
    e.set_operator(R.taylor_operator(4))


def test_full_size_batching_determinism_and_store_all():
    seis = traces(3)
    r_u, r_x = [N2 + 1] * 3, [N2 + 300, N2 + 1150, N2 + 2000]
    with engine(3) as e:
        setup(e)
        u3, d3, s3 = e.migrate(r_u, r_x, seis)
        u3b, d3b, _ = e.migrate(r_u, r_x, seis)
    assert np.array_equal(u3, u3b) and np.array_equal(d3, d3b)          # deterministic
    assert np.isfinite(u3).all() and np.abs(u3).max() > 0
    with engine(1) as e:
        setup(e)
        u1, d1, s1 = e.migrate(r_u[1:2], r_x[1:2], seis[1:2])
    assert np.array_equal(u1[0], u3[1]) and np.array_equal(d1[0], d3[1]) and s1[0] == s3[1]  # batching
    with engine(1, flags=R.STORE_ALL) as e:                               # 7501 x 7.4 MB = 55 GB in HBM
        assert e.store_all_active()
        setup(e)
        us, ds_, _ = e.migrate(r_u[1:2], r_x[1:2], seis[1:2])
    assert rel_l2(ds_[0], d1[0]) < 1e-3      # illumination: sum of S^2
    assert rel_l2(us[0], u1[0]) < 5e-2       # filtered, compensated cross-correlation


def test_full_width_translation_invariance():
    nt = 1200                                 # the wavefront stays clear of the lateral ring
    with engine(2, NT_=nt, n=64) as e:
        setup(e, lateral=False)
        xs = [N2 + 1000, N2 + 1000 + 128 + 37]  # one tile width plus an odd offset
        _, snaps = e.forward([N2 + 200, N2 + 200], xs, want_gather=False, snaps=(400, 800, nt - 1))
    sh = xs[1] - xs[0]
    for i in range(3):
        a, b = snaps[0, i], snaps[1, i]
        w = slice(N2 + 300, N2 + 1700)        # window around the first source, well inside the grid
        assert np.abs(a[:, w]).max() > 0
        assert np.array_equal(a[N2:-N2, w], b[N2:-N2, w.start + sh:w.stop + sh])
