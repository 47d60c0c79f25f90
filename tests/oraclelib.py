"""ctypes bindings to the parity checkers (TEST INFRASTRUCTURE, never product code):
oracle/librtm_oracle.so (our C restatement) and oracle/_ref/libref_host.so (the
reference's own host functions, only present where oracle/Makefile could build them)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
ORACLE_SO = ROOT / "oracle" / "librtm_oracle.so"
REFHOST_SO = ROOT / "oracle" / "_ref" / "libref_host.so"

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)


class OracleParams(C.Structure):
    _fields_ = [("mod_NZ", C.c_int), ("mod_NX", C.c_int), ("N2", C.c_int), ("nfdmax", C.c_int),
                ("NT", C.c_int), ("iLSTE", C.c_int), ("iCompen", C.c_int),
                ("h", C.c_float), ("hz", C.c_float), ("tao", C.c_float), ("f0", C.c_float),
                ("vmin", C.c_float), ("dv", C.c_float), ("vmax", C.c_float),
                ("whitecoe", C.c_float),
                ("s_l", C.c_int), ("s_z", C.c_int), ("n", C.c_int), ("ds", C.c_int),
                ("contract", C.c_int)]


def _f(a):
    return None if a is None else a.ctypes.data_as(fp)


def _i(a):
    return None if a is None else a.ctypes.data_as(ip)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        if not ORACLE_SO.exists():
            subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "oracle"])
        L = C.CDLL(str(ORACLE_SO))
        L.oracle_ricker.restype = C.c_float
        L.oracle_ricker.argtypes = [C.c_float, C.c_float]
        L.oracle_velocity_bins.restype = C.c_int
        L.oracle_velocity_bins.argtypes = [fp, C.c_long, C.c_float, fp, fp, ip, C.c_int]
        L.oracle_pad_velocity.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                          C.c_float, fp, fp]
        L.oracle_taylor.argtypes = [C.c_int, fp]
        L.oracle_derived.argtypes = [C.c_float] * 5 + [C.c_int, ip, ip, fp, fp, fp, fp, fp]
        L.oracle_strips_alloc.restype = C.c_void_p
        L.oracle_strips_alloc.argtypes = [C.POINTER(OracleParams)]
        L.oracle_strips_free.argtypes = [C.c_void_p]
        L.oracle_forward.argtypes = [C.POINTER(OracleParams), fp, fp, ip, C.c_int, C.c_int, fp, fp,
                                     fp, C.c_void_p, C.c_int, ip, C.POINTER(fp)]
        L.oracle_migrate_shot.argtypes = [C.POINTER(OracleParams), fp, fp, ip, C.c_int, C.c_int,
                                          fp, fp, fp, fp, fp, fp]
        L.oracle_migrate_shot_ex.argtypes = [C.POINTER(OracleParams), fp, fp, ip, C.c_int, C.c_int,
                                             fp, fp, fp, fp, fp, fp, C.c_int]
        L.oracle_stack.argtypes = [C.POINTER(fp), C.POINTER(fp), C.c_int, C.c_int, C.c_int,
                                   C.c_int, fp, fp]
        _oracle = L
    return _oracle


def ricker(t, f0):
    return float(oracle().oracle_ricker(np.float32(t), np.float32(f0)))


def derived(h, hz, tao, tao1, f0, NT1):
    NT, NT2 = C.c_int(), C.c_int()
    fl = [C.c_float() for _ in range(5)]
    oracle().oracle_derived(h, hz, tao, tao1, f0, NT1, C.byref(NT), C.byref(NT2),
                            *[C.byref(x) for x in fl])
    return dict(NT=NT.value, NT2=NT2.value, taoh=fl[0].value, tao2=fl[1].value, h2=fl[2].value,
                taoh2=fl[3].value, hzx2_1=fl[4].value)


def pad_velocity(vraw, N2, ifv, tao, h):
    mod_NX, mod_NZ = vraw.shape
    v = np.empty((mod_NZ + 2 * N2, mod_NX + 2 * N2), np.float32)
    r1 = np.empty_like(v)
    oracle().oracle_pad_velocity(_f(np.ascontiguousarray(vraw)), mod_NZ, mod_NX, N2, ifv, tao, h,
                                 _f(v), _f(r1))
    return v, r1


def velocity_bins(v, dv):
    vmin, vmax = C.c_float(), C.c_float()
    nvel = oracle().oracle_velocity_bins(_f(v), v.size, dv, C.byref(vmin), C.byref(vmax), None, 0)
    need = np.zeros(nvel, np.int32)
    oracle().oracle_velocity_bins(_f(v), v.size, dv, C.byref(vmin), C.byref(vmax), _i(need), nvel)
    return vmin.value, vmax.value, nvel, need


def taylor(M):
    c = np.zeros(M + 1, np.float32)
    oracle().oracle_taylor(M, _f(c))
    return c


def make_params(case, vmin, vmax, NT=None, contract=0):
    p = OracleParams()
    p.mod_NZ, p.mod_NX, p.N2, p.nfdmax = case.mod_NZ, case.mod_NX, case.N2, case.nfdmax
    p.NT = case.NT if NT is None else NT
    p.iLSTE, p.iCompen = case.iLSTE, case.iCompen
    p.h, p.hz, p.tao, p.f0 = case.h, case.hz, case.tao, case.f0
    p.vmin, p.dv, p.vmax, p.whitecoe = vmin, case.dv, vmax, case.whitecoe
    p.s_l, p.s_z, p.n, p.ds = case.s_l0, case.s_z0, case.n, case.ds
    p.contract = contract
    return p


def forward(p, v, c, Index, r_u, r_x, want_gather=True, snaps=()):
    NZ, NX = v.shape
    gather = np.zeros((p.n, p.NT), np.float32) if want_gather else None
    last0 = np.zeros((NZ, NX), np.float32)
    last1 = np.zeros((NZ, NX), np.float32)
    sk = np.asarray(list(snaps), np.int32)
    so = [np.zeros((NZ, NX), np.float32) for _ in snaps]
    arr = (fp * max(1, len(so)))(*[_f(a) for a in so])
    oracle().oracle_forward(C.byref(p), _f(v), _f(c), _i(Index), r_u, r_x, _f(gather), _f(last0),
                            _f(last1), None, len(so), _i(sk) if len(so) else None, arr)
    return gather, last0, last1, so


def migrate_shot(p, v, c, Index, r_u, r_x, seis, store_all=0):
    up = np.zeros((p.mod_NX, p.mod_NZ), np.float32)
    down = np.zeros_like(up)
    rel1 = np.zeros((p.mod_NZ, p.mod_NX), np.float32)
    rel2 = np.zeros_like(rel1)
    stable = C.c_float()
    seis = np.ascontiguousarray(seis, np.float32)
    assert seis.shape == (p.n, p.NT)
    oracle().oracle_migrate_shot_ex(C.byref(p), _f(v), _f(c), _i(Index), r_u, r_x, _f(seis), _f(up),
                                    _f(down), _f(rel1), _f(rel2), C.byref(stable), store_all)
    return up, down, rel1, rel2, stable.value


def stack(ups, downs, iNorm):
    mod_NX, mod_NZ = ups[0].shape
    ua = (fp * len(ups))(*[_f(a) for a in ups])
    da = (fp * len(downs))(*[_f(a) for a in downs])
    out = np.zeros((mod_NX, mod_NZ), np.float32)
    outd = np.zeros_like(out)
    oracle().oracle_stack(ua, da, len(ups), mod_NZ, mod_NX, iNorm, _f(out), _f(outd))
    return out, outd


# ------------------------------------------------------------------ real reference host code
_refhost = None


def refhost():
    """The reference's own host functions (None when oracle/_ref was not built)."""
    global _refhost
    if _refhost is None and REFHOST_SO.exists():
        L = C.CDLL(str(REFHOST_SO))
        L.ref_funMandC.restype = C.c_int
        L.ref_funMandC.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_double] * 8 + \
                                  [ip, ip, ip, C.POINTER(fp), C.c_double]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_order.argtypes = [C.c_int, fp]
        L.ref_cal2dfdcoe_lsm.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int,
                                         C.c_double]
        L.ref_callenfd2d_ls.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_double] * 4 + \
                                       [C.c_int, C.c_double, ip, C.c_double]
        L.ref_velocity.argtypes = [C.c_char_p, fp, fp, fp, fp, fp, C.c_int, C.c_int, C.c_int,
                                   C.c_float, C.c_float, C.c_int]
        L.ref_resample.argtypes = [C.c_int, C.c_float, fp, C.c_int, C.c_float, fp]
        L.ref_segy2trace.argtypes = [C.c_char_p, fp, C.c_int, C.c_int]
        L.ref_trace2segy.argtypes = [C.c_char_p, fp, C.c_int, C.c_int]
        L.ref_segy2head.argtypes = [C.c_char_p, ip, C.c_int]
        L.ref_head2segy.argtypes = [C.c_char_p, ip, C.c_int]
        L.ref_WriteSGY.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp, fp, C.c_float, C.c_float, fp, C.c_char_p]
        L.ref_D2T.argtypes = [C.c_char_p, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        L.ref_T2D.argtypes = [C.c_char_p, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        L.ref_phase_correction.argtypes = [fp, fp, C.c_int, C.c_int, C.c_float]
        _refhost = L
    return _refhost


def ref_funMandC(nthita, nfdmax, nfdmin, nvel, tao, h, df, eps, fmax, vmin, vmax, dv, need, hzx):
    L = refhost()
    M = np.zeros(nvel, np.int32)
    Index = np.zeros(nvel + 1, np.int32)
    need = np.ascontiguousarray(need, np.int32)
    cp = fp()
    NC = L.ref_funMandC(nthita, nfdmax, nfdmin, nvel, float(np.float32(tao)), float(np.float32(h)),
                        float(np.float32(df)), float(np.float32(eps)), float(np.float32(fmax)),
                        float(np.float32(vmin)), float(np.float32(vmax)), float(np.float32(dv)),
                        _i(need), _i(M), _i(Index), C.byref(cp), float(np.float32(hzx)))
    c = np.ctypeslib.as_array(cp, shape=(NC,)).copy()
    L.ref_free(cp)
    return NC, M, Index, c
