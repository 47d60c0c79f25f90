"""rtm_gpu_b200 -- Python face of the B200-native RTM engine.

The product is the C-ABI shared library `librtm_b200.so` (include/rtm_b200.h), built in-tree
from rtm_gpu_b200/csrc by `python -m rtm_gpu_b200.build`.  This module only loads it with
ctypes and mirrors the C entry points for the parity tests and the benchmark; there is no
Python or CPU implementation of the time loop behind it: without the library or without a
GPU every engine call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
import os

STORE_ALL = 1  # RTM_FLAG_STORE_ALL
LIB_PATH = Path(os.environ.get("RTM_LIB_PATH", PKG / "librtm_b200.so"))  # override: kernel-variant experiments

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


class RtmError(RuntimeError):
    pass


class Params(C.Structure):
    """rtm_params (include/rtm_b200.h)."""
    _fields_ = [("mod_NZ", C.c_int), ("mod_NX", C.c_int), ("N2", C.c_int), ("nfdmax", C.c_int),
                ("NT", C.c_int), ("iLSTE", C.c_int), ("iCompen", C.c_int),
                ("h", C.c_float), ("hz", C.c_float), ("tao", C.c_float), ("f0", C.c_float),
                ("whitecoe", C.c_float),
                ("s_l", C.c_int), ("s_z", C.c_int), ("n", C.c_int), ("ds", C.c_int),
                ("max_batch", C.c_int), ("flags", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("cell_updates", C.c_double), ("device_seconds", C.c_double),
                ("forward_seconds", C.c_double), ("backward_seconds", C.c_double),
                ("algorithmic_bytes", C.c_double), ("kernel_launches", C.c_long), ("shots", C.c_long),
                ("executed_bytes_forward", C.c_double), ("executed_bytes_backward", C.c_double),
                ("pair_cell_steps_forward", C.c_double), ("pair_cell_steps_backward", C.c_double)]


# every symbol include/rtm_b200.h declares (checked by tests/test_host.py::test_abi_exports_every_declared_symbol)
ABI_SYMBOLS = [
    "rtm_last_error", "rtm_version", "rtm_create", "rtm_destroy", "rtm_set_model", "rtm_set_operator",
    "rtm_forward", "rtm_migrate", "rtm_migrate_raw", "rtm_resample_device", "rtm_upload_gathers", "rtm_migrate_resident", "rtm_stack_reset",
    "rtm_stack_get", "rtm_stack_device", "rtm_stack_reduce", "rtm_stack_reduce_backend", "rtm_stack_reduce_prepare", "rtm_stack_finalize", "rtm_get_stats",
    "rtm_reset_stats", "rtm_device_count", "rtm_memory_estimate", "rtm_device_free_bytes", "rtm_host_alloc_pinned", "rtm_host_free_pinned", "rtm_store_all_active", "rtm_ricker", "rtm_source_row", "rtm_derived",
    "rtm_pad_velocity", "rtm_velocity_bins", "rtm_taylor_operator", "rtm_ls_operator",
    "rtm_ls_coefficients", "rtm_resample", "rtm_segy_decode", "rtm_segy_encode", "rtm_segy_info",
    "rtm_segy_read", "rtm_segy_write_image", "rtm_depth_to_time", "rtm_time_to_depth", "rtm_phase_rotate",
    "rtm_run_driver",
]

_lib = None


def lib():
    """Load librtm_b200.so (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RtmError(f"{LIB_PATH} is missing: run `python -m rtm_gpu_b200.build` "
                       "(the engine has no Python/CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    L.rtm_last_error.restype = C.c_char_p
    L.rtm_version.restype = C.c_char_p
    L.rtm_create.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p)]
    L.rtm_destroy.argtypes = [C.c_void_p]
    L.rtm_destroy.restype = None
    L.rtm_set_model.argtypes = [C.c_void_p, _fp, C.c_float, C.c_float, C.c_float]
    L.rtm_set_operator.argtypes = [C.c_void_p, _ip, C.c_int, _fp, C.c_int]
    L.rtm_forward.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _fp, C.c_int, _ip, _fp]
    L.rtm_migrate.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _fp, _fp, _fp, _fp]
    L.rtm_migrate_raw.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _fp, C.c_int, C.c_float, _fp, _fp, _fp]
    L.rtm_resample_device.argtypes = [C.c_void_p, C.c_int, _fp, C.c_int, C.c_float, _fp]
    L.rtm_upload_gathers.argtypes = [C.c_void_p, C.c_int, _fp]
    L.rtm_migrate_resident.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
    L.rtm_stack_reset.argtypes = [C.c_void_p]
    L.rtm_stack_get.argtypes = [C.c_void_p, _fp, _fp, _ip]
    L.rtm_stack_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), _ip]
    L.rtm_stack_reduce.argtypes = [C.POINTER(C.c_void_p), C.c_int, _fp, _fp, _ip]
    L.rtm_stack_reduce_backend.restype = C.c_char_p
    L.rtm_stack_reduce_prepare.argtypes = [_ip, C.c_int]
    L.rtm_stack_finalize.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_size_t, _fp, _fp]
    L.rtm_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.rtm_reset_stats.argtypes = [C.c_void_p]
    L.rtm_store_all_active.argtypes = [C.c_void_p]
    L.rtm_memory_estimate.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.rtm_device_free_bytes.argtypes = [C.c_int, C.POINTER(C.c_size_t)]
    L.rtm_ricker.restype = C.c_float
    L.rtm_ricker.argtypes = [C.c_float, C.c_float]
    L.rtm_source_row.argtypes = [C.c_float, C.c_float, C.c_int]
    L.rtm_derived.restype = None
    L.rtm_derived.argtypes = [C.c_float] * 5 + [C.c_int, _ip, _ip, _fp, _fp, _fp, _fp, _fp]
    L.rtm_pad_velocity.restype = None
    L.rtm_pad_velocity.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]
    L.rtm_velocity_bins.argtypes = [_fp, C.c_long, C.c_float, _fp, _fp, _ip, C.c_int]
    L.rtm_taylor_operator.restype = None
    L.rtm_taylor_operator.argtypes = [C.c_int, _fp]
    L.rtm_ls_operator.argtypes = [C.c_int] * 4 + [C.c_float] * 8 + [_ip, _ip, _ip, _fp, C.c_int, C.c_int]
    L.rtm_ls_coefficients.restype = None
    L.rtm_ls_coefficients.argtypes = [_dp, C.c_double, C.c_double, C.c_int, C.c_double]
    L.rtm_resample.restype = None
    L.rtm_resample.argtypes = [C.c_int, C.c_float, _fp, C.c_int, C.c_float, _fp]
    _bp = C.POINTER(C.c_ubyte)
    L.rtm_segy_decode.restype = None
    L.rtm_segy_decode.argtypes = [_bp, _fp, C.c_int, C.c_int]
    L.rtm_segy_encode.restype = None
    L.rtm_segy_encode.argtypes = [_bp, _fp, C.c_int, C.c_int]
    L.rtm_segy_info.argtypes = [C.c_char_p, _ip, _ip, _ip, _fp]
    L.rtm_segy_read.argtypes = [C.c_char_p, _fp, C.c_int, C.c_int]
    L.rtm_segy_write_image.argtypes = [C.c_char_p, C.c_char_p, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp,
                                       C.c_float, C.c_float, _fp]
    L.rtm_depth_to_time.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float, _fp, C.c_int]
    L.rtm_time_to_depth.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _fp, C.c_int]
    L.rtm_phase_rotate.restype = None
    L.rtm_phase_rotate.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_float]
    L.rtm_run_driver.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    _lib = L
    return L


def _f(a):
    return None if a is None else a.ctypes.data_as(_fp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _check(rc):
    if rc != 0:
        raise RtmError(f"rtm error {rc}: {lib().rtm_last_error().decode(errors='replace')}")


# ------------------------------------------------------------------ host-side pieces
def ricker(t1, f0):
    return float(lib().rtm_ricker(np.float32(t1), np.float32(f0)))


def source_row(depth_m, hz, N2):
    return int(lib().rtm_source_row(np.float32(depth_m), np.float32(hz), N2))


def derived(h, hz, tao, tao1, f0, NT1):
    NT, NT2 = C.c_int(), C.c_int()
    fl = [C.c_float() for _ in range(5)]
    lib().rtm_derived(h, hz, tao, tao1, f0, NT1, C.byref(NT), C.byref(NT2), *[C.byref(x) for x in fl])
    return dict(NT=NT.value, NT2=NT2.value, taoh=fl[0].value, tao2=fl[1].value, h2=fl[2].value,
                taoh2=fl[3].value, hzx2_1=fl[4].value)


def pad_velocity(vraw, N2, ifv=0):
    vraw = np.ascontiguousarray(vraw, np.float32)
    mod_NX, mod_NZ = vraw.shape
    v = np.empty((mod_NZ + 2 * N2, mod_NX + 2 * N2), np.float32)
    lib().rtm_pad_velocity(_f(vraw), mod_NZ, mod_NX, N2, ifv, _f(v))
    return v


def velocity_bins(v, dv):
    v = np.ascontiguousarray(v, np.float32)
    vmin, vmax = C.c_float(), C.c_float()
    nvel = lib().rtm_velocity_bins(_f(v), v.size, dv, C.byref(vmin), C.byref(vmax), None, 0)
    need = np.zeros(nvel, np.int32)
    lib().rtm_velocity_bins(_f(v), v.size, dv, C.byref(vmin), C.byref(vmax), _i(need), nvel)
    return vmin.value, vmax.value, nvel, need


def taylor_operator(M):
    c = np.zeros(M + 1, np.float32)
    lib().rtm_taylor_operator(M, _f(c))
    return c


def ls_operator(nthita, nfdmax, nfdmin, nvel, tao, h, df, eps, fmax, vmin, dv, hzx, need, verbose=False):
    need = np.ascontiguousarray(need, np.int32)
    M = np.zeros(nvel, np.int32)
    Index = np.zeros(nvel + 1, np.int32)
    cap = nvel * (nfdmax + 1)
    c = np.zeros(cap, np.float32)
    NC = lib().rtm_ls_operator(nthita, nfdmax, nfdmin, nvel, tao, h, df, eps, fmax, vmin, dv, hzx,
                               _i(need), _i(M), _i(Index), _f(c), cap, int(verbose))
    return NC, M, Index, c[:NC].copy()


def ls_coefficients(r, bmax, M, hzx):
    c = np.zeros(M + 1, np.float64)
    lib().rtm_ls_coefficients(c.ctypes.data_as(_dp), r, bmax, M, hzx)
    return c


def resample(yin, dxin, nxout, dxout):
    yin = np.ascontiguousarray(yin, np.float32)
    out = np.zeros(nxout, np.float32)
    lib().rtm_resample(len(yin), dxin, _f(yin), nxout, dxout, _f(out))
    return out


def segy_decode(buf: bytes, ns, fmt):
    out = np.zeros(ns, np.float32)
    b = (C.c_ubyte * len(buf)).from_buffer_copy(buf)
    lib().rtm_segy_decode(b, _f(out), ns, fmt)
    return out


def segy_encode(x, fmt):
    x = np.ascontiguousarray(x, np.float32)
    b = (C.c_ubyte * (len(x) * (2 if fmt == 3 else 4)))()
    lib().rtm_segy_encode(b, _f(x), len(x), fmt)
    return bytes(b)


def segy_read(path):
    ns, ntr, fmt = C.c_int(), C.c_int(), C.c_int()
    dt = C.c_float()
    _check(lib().rtm_segy_info(str(path).encode(), C.byref(ns), C.byref(ntr), C.byref(fmt), C.byref(dt)))
    out = np.zeros((ntr.value, ns.value), np.float32)
    _check(lib().rtm_segy_read(str(path).encode(), _f(out), ns.value, ntr.value))
    return out, fmt.value, dt.value


def segy_write_image(template, out_path, data, dt_value, SX, SY, RX, RY, DSR):
    data = np.ascontiguousarray(data, np.float32)
    ntr, ns = data.shape
    SX, SY, DSR = (np.ascontiguousarray(a, np.float32) for a in (SX, SY, DSR))
    _check(lib().rtm_segy_write_image(str(template).encode(), str(out_path).encode(), _f(data), ntr, ns,
                                      int(dt_value), _f(SX), _f(SY), RX, RY, _f(DSR)))


def depth_to_time(V, D, dz, dt):
    V = np.ascontiguousarray(V, np.float32)
    D = np.ascontiguousarray(D, np.float32)
    Nx, Nz = D.shape
    nt = lib().rtm_depth_to_time(_f(V), _f(D), Nx, Nz, dz, dt, None, 0)
    T = np.zeros((Nx, max(nt, 0)), np.float32)
    lib().rtm_depth_to_time(_f(V), _f(D), Nx, Nz, dz, dt, _f(T), T.size)
    return T


def time_to_depth(V, D, Nz_V, dtime, ddepth):
    V = np.ascontiguousarray(V, np.float32)
    D = np.ascontiguousarray(D, np.float32)
    Nx, Nt = D.shape
    nz = lib().rtm_time_to_depth(_f(V), _f(D), Nx, Nt, Nz_V, dtime, ddepth, None, 0)
    Z = np.zeros((Nx, max(nz, 0)), np.float32)
    lib().rtm_time_to_depth(_f(V), _f(D), Nx, Nt, Nz_V, dtime, ddepth, _f(Z), Z.size)
    return Z


def phase_rotate(d, angle_deg):
    d = np.ascontiguousarray(d, np.float32)
    out = np.zeros_like(d)
    lib().rtm_phase_rotate(_f(d), _f(out), d.shape[0], d.shape[1], angle_deg)
    return out


# ------------------------------------------------------------------ engine
class Engine:
    """One context per GPU (rtm_create ... rtm_destroy)."""

    def __init__(self, device=0, *, mod_NZ, mod_NX, N2, nfdmax, NT, iLSTE, iCompen, h, hz, tao, f0,
                 whitecoe, s_l, s_z, n, ds, max_batch=1, flags=0):
        self.params = Params(mod_NZ, mod_NX, N2, nfdmax, NT, iLSTE, iCompen, h, hz, tao, f0, whitecoe,
                             s_l, s_z, n, ds, max_batch, flags)
        self._h = C.c_void_p()
        _check(lib().rtm_create(device, C.byref(self.params), C.byref(self._h)))
        self.NZ, self.NX = mod_NZ + 2 * N2, mod_NX + 2 * N2

    def close(self):
        if getattr(self, "_h", None):
            lib().rtm_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_model(self, v_padded, vmin, vmax, dv):
        v = np.ascontiguousarray(v_padded, np.float32)
        assert v.shape == (self.NZ, self.NX)
        _check(lib().rtm_set_model(self._h, _f(v), vmin, vmax, dv))

    def set_operator(self, c, Index=None):
        c = np.ascontiguousarray(c, np.float32)
        if Index is not None:
            Index = np.ascontiguousarray(Index, np.int32)
            _check(lib().rtm_set_operator(self._h, _i(Index), len(Index) - 1, _f(c), len(c)))
        else:
            _check(lib().rtm_set_operator(self._h, None, 0, _f(c), len(c)))

    def forward(self, r_u, r_x, want_gather=True, snaps=()):
        p = self.params
        r_u = np.ascontiguousarray(r_u, np.int32)
        r_x = np.ascontiguousarray(r_x, np.int32)
        ns = len(r_u)
        g = np.zeros((ns, p.n, p.NT), np.float32) if want_gather else None
        sk = np.ascontiguousarray(list(snaps), np.int32)
        so = np.zeros((ns, len(sk), self.NZ, self.NX), np.float32) if len(sk) else None
        _check(lib().rtm_forward(self._h, ns, _i(r_u), _i(r_x), _f(g), len(sk),
                                 _i(sk) if len(sk) else None, _f(so)))
        return g, so

    def migrate(self, r_u, r_x, seis, want_images=True):
        p = self.params
        r_u = np.ascontiguousarray(r_u, np.int32)
        r_x = np.ascontiguousarray(r_x, np.int32)
        ns = len(r_u)
        seis = np.ascontiguousarray(seis, np.float32)
        assert seis.shape == (ns, p.n, p.NT), (seis.shape, (ns, p.n, p.NT))
        up = np.zeros((ns, p.mod_NX, p.mod_NZ), np.float32) if want_images else None
        down = np.zeros_like(up) if want_images else None
        stable = np.zeros(ns, np.float32)
        _check(lib().rtm_migrate(self._h, ns, _i(r_u), _i(r_x), _f(seis), _f(up), _f(down), _f(stable)))
        return up, down, stable

    def migrate_raw(self, r_u, r_x, seis_raw, tao1):
        p = self.params
        r_u = np.ascontiguousarray(r_u, np.int32)
        r_x = np.ascontiguousarray(r_x, np.int32)
        ns = len(r_u)
        seis_raw = np.ascontiguousarray(seis_raw, np.float32)
        assert seis_raw.shape[:2] == (ns, p.n)
        up = np.zeros((ns, p.mod_NX, p.mod_NZ), np.float32)
        down = np.zeros_like(up)
        stable = np.zeros(ns, np.float32)
        _check(lib().rtm_migrate_raw(self._h, ns, _i(r_u), _i(r_x), _f(seis_raw), seis_raw.shape[2], tao1,
                                     _f(up), _f(down), _f(stable)))
        return up, down, stable

    def resample_device(self, traces, tao1):
        traces = np.ascontiguousarray(traces, np.float32)
        out = np.zeros((traces.shape[0], self.params.NT), np.float32)
        _check(lib().rtm_resample_device(self._h, traces.shape[0], _f(traces), traces.shape[1], tao1, _f(out)))
        return out

    def upload_gathers(self, seis):
        seis = np.ascontiguousarray(seis, np.float32)
        _check(lib().rtm_upload_gathers(self._h, seis.shape[0], _f(seis)))

    def migrate_resident(self, r_u, r_x):
        r_u = np.ascontiguousarray(r_u, np.int32)
        r_x = np.ascontiguousarray(r_x, np.int32)
        _check(lib().rtm_migrate_resident(self._h, len(r_u), _i(r_u), _i(r_x)))

    def stack_reset(self):
        _check(lib().rtm_stack_reset(self._h))

    def stack_get(self):
        p = self.params
        up = np.zeros((p.mod_NX, p.mod_NZ), np.float32)
        down = np.zeros_like(up)
        ns = C.c_int()
        _check(lib().rtm_stack_get(self._h, _f(up), _f(down), C.byref(ns)))
        return up, down, ns.value

    def stack_device(self):
        ptr, nfl, ns = C.c_void_p(), C.c_size_t(), C.c_int()
        _check(lib().rtm_stack_device(self._h, C.byref(ptr), C.byref(nfl), C.byref(ns)))
        return ptr.value, nfl.value, ns.value

    def stats(self):
        s = Stats()
        _check(lib().rtm_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def store_all_active(self):
        return bool(lib().rtm_store_all_active(self._h))

    def reset_stats(self):
        _check(lib().rtm_reset_stats(self._h))


def stack_reduce(engines):
    """One reduce of the stacks of engines living in this process (one per GPU)."""
    p = engines[0].params
    up = np.zeros((p.mod_NX, p.mod_NZ), np.float32)
    down = np.zeros_like(up)
    ns = C.c_int()
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    _check(lib().rtm_stack_reduce(arr, len(engines), _f(up), _f(down), C.byref(ns)))
    return up, down, ns.value, lib().rtm_stack_reduce_backend().decode()


def stack_finalize(up_sum, down_sum, nrec, iNorm):
    up_sum = np.ascontiguousarray(up_sum, np.float32)
    down_sum = np.ascontiguousarray(down_sum, np.float32)
    img = np.zeros_like(up_sum)
    ill = np.zeros_like(up_sum)
    _check(lib().rtm_stack_finalize(_f(up_sum), _f(down_sum), nrec, iNorm, up_sum.size, _f(img), _f(ill)))
    return img, ill


def engine_for_case(case, NT=None, max_batch=1, device=0, flags=0):
    """Engine configured from a tests/refcase.Case-like object (reference parameter names)."""
    d = derived(case.h, case.hz, case.tao, case.tao1, case.f0, case.NT1)
    return Engine(device, mod_NZ=case.mod_NZ, mod_NX=case.mod_NX, N2=case.N2, nfdmax=case.nfdmax,
                  NT=d["NT"] if NT is None else NT, iLSTE=case.iLSTE, iCompen=case.iCompen, h=case.h,
                  hz=case.hz, tao=case.tao, f0=case.f0, whitecoe=case.whitecoe,
                  s_l=case.s_l + case.N2 - 1, s_z=case.s_z + case.N2 - 1, n=case.n, ds=case.ds,
                  max_batch=max_batch, flags=flags)
