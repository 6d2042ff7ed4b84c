"""ctypes binding over libstabgpu.so (include/stabgpu.h).

Only plain pointers and sizes cross the boundary; numpy arrays are passed as host pointers.
There is no CPU fallback: a missing library raises at import of the first symbol, a missing CUDA
device makes every compute entry point raise `StabGpuError`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

LIB_PATH = os.environ.get("STABGPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libstabgpu.so")
NDOF = 5


class StabGpuError(RuntimeError):
    pass


class Params(C.Structure):
    """struct stabgpu_params: the scalar run state of module stuff (stuff.f90:11-59)."""

    _fields_ = [
        ("ny", C.c_int), ("mattyp", C.c_int), ("wallt", C.c_int), ("top", C.c_int), ("curve", C.c_int),
        ("ider", C.c_int), ("ievec", C.c_int), ("wall", C.c_int),
        ("Ma", C.c_double), ("Re", C.c_double), ("Pr", C.c_double),
        ("gamma", C.c_double), ("gamma1", C.c_double), ("cp", C.c_double),
        ("Te", C.c_double), ("rmue", C.c_double), ("rlme", C.c_double), ("cone", C.c_double),
        ("datmat", C.c_double * 3),
        ("yi", C.c_double), ("ymax", C.c_double), ("x", C.c_double),
    ]

    @classmethod
    def default(cls) -> "Params":
        p = cls()
        lib().stabgpu_params_default(C.byref(p))
        return p

    def copy(self) -> "Params":
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_pp = C.POINTER(Params)


def lib():
    """Load libstabgpu.so (once) and declare the prototypes of include/stabgpu.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StabGpuError(
            f"{LIB_PATH} not found: build it with `make -C stab_b200/csrc` or __graft_entry__.build() "
            "(there is no CPU fallback for the stab hot path)")
    L = C.CDLL(LIB_PATH)

    def proto(name, res, args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args

    i, d, vp, cp = C.c_int, C.c_double, C.c_void_p, C.c_char_p
    proto("stabgpu_init", i, [i])
    proto("stabgpu_init_multi", i, [i, _ip])
    proto("stabgpu_device_count", i, [])
    proto("stabgpu_set_host_staging", i, [i, i])
    proto("stabgpu_host_register", i, [vp, C.c_size_t])
    proto("stabgpu_host_unregister", i, [vp])
    proto("stabgpu_finalize", i, [])
    proto("stabgpu_last_error", cp, [])
    proto("stabgpu_device_info", i, [cp, i, _ip, _dp])
    proto("stabgpu_set_tuning", i, [i, i, i, i])
    proto("stabgpu_set_qr_deflation", i, [i, i])
    proto("stabgpu_set_hess_mode", i, [i])
    proto("stabgpu_set_evec_mode", i, [i])
    proto("stabgpu_set_lu_mode", i, [i])
    proto("stabgpu_params_default", None, [_pp])
    proto("stabgpu_edge_properties", i, [_pp, d])
    proto("stabgpu_sgengrid", i, [i, d, d, vp, vp, vp, vp])
    proto("stabgpu_chebyd", i, [i, vp])
    proto("stabgpu_spline", i, [i, vp, vp, vp])
    proto("stabgpu_speval", i, [i, vp, vp, vp, d, _dp])
    proto("stabgpu_getmean_table", i, [i, vp, i, vp, vp])
    proto("stabgpu_read_profile", i, [cp, _ip, vp, i])
    proto("stabgpu_mean_gradients", i, [i, i, vp, vp, vp, vp, vp, vp, vp, vp])
    proto("stabgpu_circh", i, [_dp, i, vp, vp])
    proto("stabgpu_temporal_batch", i, [_pp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, vp])
    proto("stabgpu_spatial_batch", i, [_pp, vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, vp])
    proto("stabgpu_zgeev_batch", i, [i, i, vp, i, vp, vp, vp])
    proto("stabgpu_temporal_polish", i, [_pp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, d, vp, vp, _dp, _ip])
    proto("stabgpu_polish_batch", i, [i, _pp, vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, i, d, vp, vp, vp, vp])
    proto("stabgpu_temporal_assemble", i, [_pp, vp, vp, vp, vp, vp, vp, vp, vp, vp])
    proto("stabgpu_spatial_assemble", i, [_pp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp])
    proto("stabgpu_debug_stages", i, [i, vp, vp, vp, _ip, _ip, vp, vp])
    proto("stabgpu_debug_spatial_reduce", i, [_pp, vp, vp, vp, vp, vp, vp, vp, vp, vp, _ip])
    proto("stabgpu_plan_create", i, [C.POINTER(vp), i, _pp, vp, vp, vp, vp, vp, vp, i, i])
    proto("stabgpu_plan_upload", i, [vp, i, vp, vp, vp, vp])
    proto("stabgpu_plan_execute", i, [vp])
    proto("stabgpu_plan_enqueue", i, [vp])
    proto("stabgpu_plan_wait", i, [vp])
    proto("stabgpu_plan_download", i, [vp, vp, vp, vp])
    proto("stabgpu_plan_stage_times", i, [vp, C.POINTER(C.c_float)])
    proto("stabgpu_plan_launch_count", C.c_longlong, [vp])
    proto("stabgpu_plan_stream", vp, [vp])
    proto("stabgpu_plan_ilohi", i, [vp, vp])
    proto("stabgpu_plan_profile_hessenberg", i, [vp, i, C.POINTER(C.c_float)])
    proto("stabgpu_plan_profile_eigvec", i, [vp, C.POINTER(C.c_float)])
    proto("stabgpu_plan_capacity", i, [vp])
    proto("stabgpu_plan_eig_dev", vp, [vp])
    proto("stabgpu_plan_destroy", i, [vp])
    proto("stabgpu_mtemporal_points", i, [d, d, d, d, d, d, vp, vp, i])
    proto("stabgpu_mspatial_points", i, [d, d, d, d, d, d, vp, vp, i])
    proto("stabgpu_shard_range", None, [i, i, i, _ip, _ip])
    proto("stabgpu_write_eig_file", i, [cp, _pp, i, i, vp, vp, vp, d, vp, vp, vp, vp, vp, vp])
    _lib = L
    return L


def _check(rc: int, what: str):
    if rc != 0:
        msg = lib().stabgpu_last_error()
        raise StabGpuError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def _f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def _c128(a) -> np.ndarray:
    return np.ascontiguousarray(np.atleast_1d(a), dtype=np.complex128)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _colmajor(a: np.ndarray) -> np.ndarray:
    """(rows, cols) numpy array -> flat column-major buffer."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)


# ---- host-side pieces -----------------------------------------------------------------------------
def edge_properties(p: Params, T0: float = 0.0) -> Params:
    _check(lib().stabgpu_edge_properties(C.byref(p), float(T0)), "stabgpu_edge_properties")
    return p


def sgengrid(ny: int, yi: float, ymax: float):
    y, eta, deta, d2eta = (np.empty(ny) for _ in range(4))
    _check(lib().stabgpu_sgengrid(ny, yi, ymax, _ptr(y), _ptr(eta), _ptr(deta), _ptr(d2eta)), "stabgpu_sgengrid")
    return y, eta, deta, d2eta


def chebyd(N: int) -> np.ndarray:
    buf = np.empty((N + 1, N + 1))
    _check(lib().stabgpu_chebyd(N, _ptr(buf)), "stabgpu_chebyd")
    return buf.T.copy()      # buffer is column-major


def read_profile(path: str, max_rows: int = 100000) -> np.ndarray:
    tab = np.empty((max_rows, 6))
    n = C.c_int(0)
    _check(lib().stabgpu_read_profile(path.encode(), C.byref(n), _ptr(tab), max_rows), "stabgpu_read_profile")
    return tab[: n.value].copy()


def getmean(table: np.ndarray, y: np.ndarray) -> np.ndarray:
    """getmean.f90:27-111 -> vm (ny, 5)."""
    table = _f64(table)
    y = _f64(y)
    ny = y.size
    buf = np.empty((5, ny))
    _check(lib().stabgpu_getmean_table(table.shape[0], _ptr(table), ny, _ptr(y), _ptr(buf)), "stabgpu_getmean_table")
    return buf.T.copy()


def mean_gradients(vm: np.ndarray, deta, d2eta, wallt: int = 0):
    """temporal.f90:135-179 -> D1, D2 (ny, ny), Dt2 wall row, g2vm, g22vm (ny, 5)."""
    ny = vm.shape[0]
    vmc, deta, d2eta = _colmajor(vm), _f64(deta), _f64(d2eta)
    D1, D2 = np.empty((ny, ny)), np.empty((ny, ny))
    Dt2w, g2, g22 = np.empty(ny), np.empty((5, ny)), np.empty((5, ny))
    _check(lib().stabgpu_mean_gradients(ny, wallt, _ptr(vmc), _ptr(deta), _ptr(d2eta), _ptr(D1), _ptr(D2), _ptr(Dt2w),
                                        _ptr(g2), _ptr(g22)), "stabgpu_mean_gradients")
    return D1.T.copy(), D2.T.copy(), Dt2w, g2.T.copy(), g22.T.copy()


def circh(radius: float, y: np.ndarray):
    """circh.f90:35-188 -> (x_out, h5 (ny, 5))."""
    y = _f64(y)
    x = C.c_double(radius)
    buf = np.empty((5, y.size))
    _check(lib().stabgpu_circh(C.byref(x), y.size, _ptr(y), _ptr(buf)), "stabgpu_circh")
    return x.value, buf.T.copy()


def mtemporal_points(amin, amax, ainc, bmin, bmax, binc):
    n = lib().stabgpu_mtemporal_points(amin, amax, ainc, bmin, bmax, binc, None, None, 0)
    if n < 0:
        raise StabGpuError("mtemporal: zero / non-finite increment or more than 10^7 sweep points")
    a, b = np.empty(n), np.empty(n)
    lib().stabgpu_mtemporal_points(amin, amax, ainc, bmin, bmax, binc, _ptr(a), _ptr(b), n)
    return a, b


def mspatial_points(omin, omax, oinc, bmin, bmax, binc):
    n = lib().stabgpu_mspatial_points(omin, omax, oinc, bmin, bmax, binc, None, None, 0)
    if n < 0:
        raise StabGpuError("mspatial: non-finite range or more than 10^7 sweep points")
    a, b = np.empty(n), np.empty(n)
    lib().stabgpu_mspatial_points(omin, omax, oinc, bmin, bmax, binc, _ptr(a), _ptr(b), n)
    return a, b


def shard_range(npts: int, rank: int, world: int):
    lo, hi = C.c_int(0), C.c_int(0)
    lib().stabgpu_shard_range(npts, rank, world, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def write_eig_file(path: str, p: Params, itype: int, ind: int, omega: complex, alpha: complex, beta: complex, x: float,
                   y, eta, deta, d2eta, eig: np.ndarray, evec: Optional[np.ndarray]):
    """temporal.f90:883-890 / spatial.f90:1120-1126; evec is (n, n) with eigenvectors in columns."""
    om, al, be = _c128(omega), _c128(alpha), _c128(beta)
    eig = _c128(eig)
    ev = None if evec is None else np.ascontiguousarray(np.asarray(evec, dtype=np.complex128).T)
    _check(lib().stabgpu_write_eig_file(path.encode(), C.byref(p), itype, ind, _ptr(om), _ptr(al), _ptr(be), float(x),
                                        _ptr(_f64(y)), _ptr(_f64(eta)), _ptr(_f64(deta)), _ptr(_f64(d2eta)), _ptr(eig),
                                        _ptr(ev)), "stabgpu_write_eig_file")


# ---- device ---------------------------------------------------------------------------------------
def init(device: int = -1):
    _check(lib().stabgpu_init(device), "stabgpu_init")


def init_multi(max_devices: int = 0) -> int:
    """stabgpu_init_multi: the batch calls shard over up to `max_devices` GPUs (0 = all).  Returns the device count."""
    n = C.c_int(0)
    _check(lib().stabgpu_init_multi(int(max_devices), C.byref(n)), "stabgpu_init_multi")
    return n.value


def device_count() -> int:
    return int(lib().stabgpu_device_count())


def set_qr_deflation(window: int = -1, nibble: int = -1):
    """stabgpu_set_qr_deflation: deflation window of the QR stage (0 = classic deflation only) and NIBBLE in per cent."""
    _check(lib().stabgpu_set_qr_deflation(int(window), int(nibble)), "stabgpu_set_qr_deflation")


def set_host_staging(pin_mode: int = -1, copy_threads: int = 0):
    _check(lib().stabgpu_set_host_staging(int(pin_mode), int(copy_threads)), "stabgpu_set_host_staging")


def host_register(a: np.ndarray):
    _check(lib().stabgpu_host_register(_ptr(a), a.nbytes), "stabgpu_host_register")


def host_unregister(a: np.ndarray):
    _check(lib().stabgpu_host_unregister(_ptr(a)), "stabgpu_host_unregister")


def finalize():
    lib().stabgpu_finalize()


def device_info():
    name = C.create_string_buffer(256)
    sm, mem = C.c_int(0), C.c_double(0.0)
    _check(lib().stabgpu_device_info(name, 256, C.byref(sm), C.byref(mem)), "stabgpu_device_info")
    return name.value.decode(), sm.value, mem.value


def set_tuning(qr_window=0, qr_shifts=0, qr_threads=0, hess_threads=0):
    _check(lib().stabgpu_set_tuning(qr_window, qr_shifts, qr_threads, hess_threads), "stabgpu_set_tuning")


def set_hess_mode(mode: int):
    lib().stabgpu_set_hess_mode(int(mode))


def set_evec_mode(mode: int):
    lib().stabgpu_set_evec_mode(int(mode))


def set_lu_mode(mode: int):
    lib().stabgpu_set_lu_mode(int(mode))


def _grid_args(p: Params, vm, g2vm, g22vm, deta, d2eta):
    ny = p.ny
    vmc = _colmajor(_f64(vm, (ny, 5)))
    g2c = None if g2vm is None else _colmajor(_f64(g2vm, (ny, 5)))
    g22c = None if g22vm is None else _colmajor(_f64(g22vm, (ny, 5)))
    return vmc, g2c, g22c, _f64(deta, (ny,)), _f64(d2eta, (ny,))


def _evec_out(buf: np.ndarray) -> np.ndarray:
    """(npts, col, row) column-major buffers -> (npts, row, col) views with eigenvectors in columns."""
    return buf.transpose(0, 2, 1)


def temporal_batch(p: Params, vm, deta, d2eta, alpha: Sequence[complex], beta: Sequence[complex], g2vm=None, g22vm=None,
                   Re_pt=None, Ma_pt=None, want_vectors: bool = False, out=None):
    """stabgpu_temporal_batch -> (omg (npts, n), evec (npts, n, n) or None, info (npts,)).
    `out` = (omg, evec_buffer, info) preallocated host buffers (e.g. pinned); the evec buffer is in
    the library's layout (npts, column, row)."""
    vmc, g2c, g22c, de, d2e = _grid_args(p, vm, g2vm, g22vm, deta, d2eta)
    al, be = _c128(alpha), _c128(beta)
    npts, n = al.size, NDOF * p.ny
    assert be.size == npts
    re = None if Re_pt is None else _f64(Re_pt, (npts,))
    ma = None if Ma_pt is None else _f64(Ma_pt, (npts,))
    if out is not None:
        omg, ev, info = out
    else:
        omg = np.empty((npts, n), dtype=np.complex128)
        ev = np.empty((npts, n, n), dtype=np.complex128) if want_vectors else None
        info = np.zeros(npts, dtype=np.int32)
    _check(lib().stabgpu_temporal_batch(C.byref(p), _ptr(vmc), _ptr(g2c), _ptr(g22c), _ptr(de), _ptr(d2e), npts, _ptr(al),
                                        _ptr(be), _ptr(re), _ptr(ma), 1 if want_vectors else 0, _ptr(omg), _ptr(ev),
                                        _ptr(info)), "stabgpu_temporal_batch")
    return omg, (None if ev is None else _evec_out(ev)), info


def spatial_batch(p: Params, vm, deta, d2eta, omega: Sequence[complex], beta: Sequence[complex], h5=None, g2vm=None,
                  g22vm=None, Re_pt=None, Ma_pt=None, want_vectors: bool = False, out=None):
    """stabgpu_spatial_batch -> (alp (npts, 2n), evec (npts, 2n, 2n) or None, info)."""
    vmc, g2c, g22c, de, d2e = _grid_args(p, vm, g2vm, g22vm, deta, d2eta)
    h5c = None if h5 is None else _colmajor(_f64(h5, (p.ny, 5)))
    om, be = _c128(omega), _c128(beta)
    npts, N = om.size, 2 * NDOF * p.ny
    assert be.size == npts
    re = None if Re_pt is None else _f64(Re_pt, (npts,))
    ma = None if Ma_pt is None else _f64(Ma_pt, (npts,))
    if out is not None:
        alp, ev, info = out
    else:
        alp = np.empty((npts, N), dtype=np.complex128)
        ev = np.empty((npts, N, N), dtype=np.complex128) if want_vectors else None
        info = np.zeros(npts, dtype=np.int32)
    _check(lib().stabgpu_spatial_batch(C.byref(p), _ptr(vmc), _ptr(g2c), _ptr(g22c), _ptr(de), _ptr(d2e), _ptr(h5c), npts,
                                       _ptr(om), _ptr(be), _ptr(re), _ptr(ma), 1 if want_vectors else 0, _ptr(alp), _ptr(ev),
                                       _ptr(info)), "stabgpu_spatial_batch")
    return alp, (None if ev is None else _evec_out(ev)), info


def zgeev_batch(A: np.ndarray, want_vectors: bool = False):
    """A: (batch, n, n) complex.  Returns (w (batch, n), V (batch, n, n) or None, info)."""
    A = np.asarray(A, dtype=np.complex128)
    if A.ndim == 2:
        A = A[None]
    batch, n, _ = A.shape
    Ac = np.ascontiguousarray(A.transpose(0, 2, 1))
    w = np.empty((batch, n), dtype=np.complex128)
    V = np.empty((batch, n, n), dtype=np.complex128) if want_vectors else None
    info = np.zeros(batch, dtype=np.int32)
    _check(lib().stabgpu_zgeev_batch(n, batch, _ptr(Ac), 1 if want_vectors else 0, _ptr(w), _ptr(V), _ptr(info)),
           "stabgpu_zgeev_batch")
    return w, (None if V is None else _evec_out(V)), info


def temporal_assemble(p: Params, vm, deta, d2eta, alpha: complex, beta: complex, g2vm=None, g22vm=None):
    vmc, g2c, g22c, de, d2e = _grid_args(p, vm, g2vm, g22vm, deta, d2eta)
    n = NDOF * p.ny
    A0, B0 = np.empty((n, n), dtype=np.complex128), np.empty((n, n), dtype=np.complex128)
    al, be = _c128(alpha), _c128(beta)
    _check(lib().stabgpu_temporal_assemble(C.byref(p), _ptr(vmc), _ptr(g2c), _ptr(g22c), _ptr(de), _ptr(d2e), _ptr(al),
                                           _ptr(be), _ptr(A0), _ptr(B0)), "stabgpu_temporal_assemble")
    return A0.T, B0.T


def spatial_assemble(p: Params, vm, deta, d2eta, omega: complex, beta: complex, h5=None, g2vm=None, g22vm=None):
    vmc, g2c, g22c, de, d2e = _grid_args(p, vm, g2vm, g22vm, deta, d2eta)
    h5c = None if h5 is None else _colmajor(_f64(h5, (p.ny, 5)))
    n = NDOF * p.ny
    Cs = [np.empty((n, n), dtype=np.complex128) for _ in range(3)]
    om, be = _c128(omega), _c128(beta)
    _check(lib().stabgpu_spatial_assemble(C.byref(p), _ptr(vmc), _ptr(g2c), _ptr(g22c), _ptr(de), _ptr(d2e), _ptr(h5c),
                                          _ptr(om), _ptr(be), _ptr(Cs[0]), _ptr(Cs[1]), _ptr(Cs[2])),
           "stabgpu_spatial_assemble")
    return tuple(c.T for c in Cs)


def debug_spatial_reduce(p: Params, vm, deta, d2eta, omega: complex, beta: complex, h5=None, g2vm=None, g22vm=None):
    """[M1 | M2] = C0^-1 [-C1 | -C2] (n x 2n) as the LU stage leaves it (spatial.f90:978-1008), and ZGETRF's info."""
    vmc, g2c, g22c, de, d2e = _grid_args(p, vm, g2vm, g22vm, deta, d2eta)
    h5c = None if h5 is None else _colmajor(_f64(h5, (p.ny, 5)))
    n = NDOF * p.ny
    M = np.empty((2 * n, n), dtype=np.complex128)
    om, be = _c128(omega), _c128(beta)
    info = C.c_int(0)
    _check(lib().stabgpu_debug_spatial_reduce(C.byref(p), _ptr(vmc), _ptr(g2c), _ptr(g22c), _ptr(de), _ptr(d2e), _ptr(h5c),
                                              _ptr(om), _ptr(be), _ptr(M), C.byref(info)), "stabgpu_debug_spatial_reduce")
    return M.T, int(info.value)


def temporal_polish(p: Params, vm, deta, d2eta, alpha: complex, beta: complex, sigma: complex, x0=None, max_iters: int = 8,
                    tol: float = 1e-13, g2vm=None, g22vm=None):
    """stabgpu_temporal_polish: shift-invert inverse iteration near `sigma` on (A0, B0).
    Returns (lambda, x (n,), resid, iters)."""
    vmc, g2c, g22c, de, d2e = _grid_args(p, vm, g2vm, g22vm, deta, d2eta)
    n = NDOF * p.ny
    al, be, sg = _c128(alpha), _c128(beta), _c128(sigma)
    x0c = None if x0 is None else _c128(x0)
    lam = np.zeros(1, dtype=np.complex128)
    x = np.empty(n, dtype=np.complex128)
    resid, iters = C.c_double(0.0), C.c_int(0)
    _check(lib().stabgpu_temporal_polish(C.byref(p), _ptr(vmc), _ptr(g2c), _ptr(g22c), _ptr(de), _ptr(d2e), _ptr(al), _ptr(be),
                                         _ptr(sg), _ptr(x0c), max_iters, tol, _ptr(lam), _ptr(x), C.byref(resid), C.byref(iters)),
           "stabgpu_temporal_polish")
    return complex(lam[0]), x, resid.value, iters.value


def polish_batch(kind: int, p: Params, vm, deta, d2eta, s1, s2, sigma, x0=None, h5=None, g2vm=None, g22vm=None, Re_pt=None,
                 Ma_pt=None, max_iters: int = 12, tol: float = 1e-13, want_vectors: bool = True):
    """stabgpu_polish_batch: one mode per point near sigma[p].  kind 1: A0 x = omega B0 x (s1 = alpha); kind 2:
    (C0 + alpha C1 + alpha^2 C2) x = 0 (s1 = omega).  Returns (lambda (npts,), x (npts, n) or None, resid, iters)."""
    vmc, g2c, g22c, de, d2e = _grid_args(p, vm, g2vm, g22vm, deta, d2eta)
    h5c = None if h5 is None else _colmajor(_f64(h5, (p.ny, 5)))
    a1, a2, sg = _c128(s1), _c128(s2), _c128(sigma)
    npts, n = a1.size, NDOF * p.ny
    assert a2.size == npts and sg.size == npts
    re = None if Re_pt is None else _f64(Re_pt, (npts,))
    ma = None if Ma_pt is None else _f64(Ma_pt, (npts,))
    x0c = None if x0 is None else np.ascontiguousarray(np.asarray(x0, dtype=np.complex128).reshape(npts, n))
    lam = np.zeros(npts, dtype=np.complex128)
    x = np.empty((npts, n), dtype=np.complex128) if want_vectors else None
    resid = np.zeros(npts)
    iters = np.zeros(npts, dtype=np.int32)
    _check(lib().stabgpu_polish_batch(kind, C.byref(p), _ptr(vmc), _ptr(g2c), _ptr(g22c), _ptr(de), _ptr(d2e), _ptr(h5c), npts,
                                      _ptr(a1), _ptr(a2), _ptr(re), _ptr(ma), _ptr(sg), _ptr(x0c), max_iters, tol, _ptr(lam),
                                      _ptr(x), _ptr(resid), _ptr(iters)), "stabgpu_polish_batch")
    return lam, x, resid, iters


def debug_stages(A: np.ndarray):
    """Balanced matrix, scale, ilo, ihi, Hessenberg(+reflectors), tau of one matrix (stage parity tests)."""
    A = np.asarray(A, dtype=np.complex128)
    n = A.shape[0]
    Ac = np.ascontiguousarray(A.T)
    bal, hess = np.empty((n, n), dtype=np.complex128), np.empty((n, n), dtype=np.complex128)
    scale, tau = np.empty(n), np.empty(n, dtype=np.complex128)
    ilo, ihi = C.c_int(0), C.c_int(0)
    _check(lib().stabgpu_debug_stages(n, _ptr(Ac), _ptr(bal), _ptr(scale), C.byref(ilo), C.byref(ihi), _ptr(hess), _ptr(tau)),
           "stabgpu_debug_stages")
    return bal.T, scale, ilo.value, ihi.value, hess.T, tau


class Plan:
    """Device-resident plan (stabgpu_plan_*): upload sweep values, execute kernels, download results."""

    def __init__(self, kind: int, p: Params, vm, deta, d2eta, max_pts: int, want_vectors: bool = False, h5=None,
                 g2vm=None, g22vm=None):
        vmc, g2c, g22c, de, d2e = _grid_args(p, vm, g2vm, g22vm, deta, d2eta)
        h5c = None if h5 is None else _colmajor(_f64(h5, (p.ny, 5)))
        self._h = C.c_void_p()
        self.kind, self.p = kind, p.copy()
        self.N = (1 if kind == 1 else 2) * NDOF * p.ny
        self.want_vectors = bool(want_vectors)
        _check(lib().stabgpu_plan_create(C.byref(self._h), kind, C.byref(p), _ptr(vmc), _ptr(g2c), _ptr(g22c), _ptr(de),
                                         _ptr(d2e), _ptr(h5c), max_pts, 1 if want_vectors else 0), "stabgpu_plan_create")
        self.capacity = lib().stabgpu_plan_capacity(self._h)
        self.npts = 0

    def upload(self, s1, s2, Re_pt=None, Ma_pt=None):
        s1, s2 = _c128(s1), _c128(s2)
        self.npts = s1.size
        re = None if Re_pt is None else _f64(Re_pt, (self.npts,))
        ma = None if Ma_pt is None else _f64(Ma_pt, (self.npts,))
        _check(lib().stabgpu_plan_upload(self._h, self.npts, _ptr(s1), _ptr(s2), _ptr(re), _ptr(ma)), "stabgpu_plan_upload")

    def execute(self):
        _check(lib().stabgpu_plan_execute(self._h), "stabgpu_plan_execute")

    def enqueue(self):
        """Launch one pass on the plan's stream without waiting (pair with wait())."""
        _check(lib().stabgpu_plan_enqueue(self._h), "stabgpu_plan_enqueue")

    def wait(self):
        _check(lib().stabgpu_plan_wait(self._h), "stabgpu_plan_wait")

    def download(self, eig: Optional[np.ndarray] = None, evec: Optional[np.ndarray] = None, info: Optional[np.ndarray] = None):
        """Buffers may be preallocated (e.g. pinned); evec buffer layout is (npts, col, row)."""
        if eig is None:
            eig = np.empty((self.npts, self.N), dtype=np.complex128)
        if evec is None and self.want_vectors:
            evec = np.empty((self.npts, self.N, self.N), dtype=np.complex128)
        if info is None:
            info = np.zeros(self.npts, dtype=np.int32)
        _check(lib().stabgpu_plan_download(self._h, _ptr(eig), _ptr(evec), _ptr(info)), "stabgpu_plan_download")
        return eig, (None if evec is None else _evec_out(evec)), info

    def info(self) -> np.ndarray:
        """Per-point LAPACK-style status of the last pass (no eigenvalue / eigenvector transfer)."""
        info = np.zeros(self.npts, dtype=np.int32)
        _check(lib().stabgpu_plan_download(self._h, None, None, _ptr(info)), "stabgpu_plan_download")
        return info

    def stage_times(self) -> dict:
        ms = (C.c_float * 8)()
        lib().stabgpu_plan_stage_times(self._h, ms)
        names = ("assemble", "lu", "balance", "hessenberg", "prep", "qr", "sort", "evec")
        return dict(zip(names, (float(v) for v in ms)))

    def ilohi(self) -> np.ndarray:
        out = np.zeros((self.npts, 2), dtype=np.int32)
        _check(lib().stabgpu_plan_ilohi(self._h, _ptr(out)), "stabgpu_plan_ilohi")
        return out

    def profile_hessenberg(self, enable: bool) -> dict:
        ms = (C.c_float * 4)()
        lib().stabgpu_plan_profile_hessenberg(self._h, 1 if enable else 0, ms)
        return dict(zip(("panel_step", "gemv", "gemm", "other"), (float(v) for v in ms)))

    def profile_eigvec(self) -> dict:
        """Per-kernel-class times of the eigenvector stage from the last profiled execute."""
        ms = (C.c_float * 3)()
        lib().stabgpu_plan_profile_eigvec(self._h, ms)
        return dict(zip(("invit", "bt_gemm", "finalize"), (float(v) for v in ms)))

    def launch_count(self) -> int:
        return int(lib().stabgpu_plan_launch_count(self._h))

    def stream(self) -> int:
        return int(lib().stabgpu_plan_stream(self._h) or 0)

    def eig_dev(self) -> int:
        """Device address of the sorted eigenvalues (N x npts complex128, column-major)."""
        return int(lib().stabgpu_plan_eig_dev(self._h) or 0)

    def destroy(self):
        if self._h:
            lib().stabgpu_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
