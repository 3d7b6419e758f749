"""ctypes binding of the CUDA C ABI ``include/nfftcu.h`` (``nfft_b200/lib/libnfftcu.so``).

This is the device-pointer face of the engine used by the benchmark, the multi-GPU driver and the
kernel-level parity tests; the reference-facing face is :mod:`nfft_b200.plan` (``nfft_*`` plan
API over the same library).  No CPU fallback: a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

_LIB = None
_LIBPATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libnfftcu.so")

DOUBLE, FLOAT = 0, 1
OPT_TIMING, OPT_PSI_TABLE, OPT_B_KERNEL, OPT_NODE_ORDER, OPT_B_FLUSH, OPT_FFT_PRUNE, OPT_FFT_KERNEL, OPT_WINDOW_IMAGES = 1, 2, 3, 4, 5, 6, 7, 8

# every symbol include/nfftcu.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = (
    "nfftcu_last_error", "nfftcu_device_count", "nfftcu_create", "nfftcu_destroy",
    "nfftcu_get_c_phi_inv", "nfftcu_get_window_params", "nfftcu_set_nodes", "nfftcu_set_nodes_dev",
    "nfftcu_nodes_version", "nfftcu_get_index_x", "nfftcu_trafo", "nfftcu_adjoint",
    "nfftcu_trafo_direct", "nfftcu_adjoint_direct", "nfftcu_trafo_dev", "nfftcu_adjoint_dev",
    "nfftcu_trafo_direct_dev", "nfftcu_adjoint_direct_dev", "nfftcu_stage_D", "nfftcu_stage_F",
    "nfftcu_stage_B", "nfftcu_stage_BT", "nfftcu_stage_DT", "nfftcu_grid_ptr", "nfftcu_set_option",
    "nfftcu_set_stream", "nfftcu_get_stream", "nfftcu_sync", "nfftcu_stage_times", "nfftcu_b_kernel_time", "nfftcu_trafo_refresh", "nfftcu_adjoint_refresh",
    "nfftcu_launch_count", "nfftcu_malloc_device", "nfftcu_free_device", "nfftcu_malloc_pinned",
    "nfftcu_free_pinned", "nfftcu_memcpy_h2d", "nfftcu_memcpy_d2h",
    "nfftcu_solver_create", "nfftcu_solver_destroy", "nfftcu_solver_upload", "nfftcu_solver_download",
    "nfftcu_solver_vector", "nfftcu_solver_before_loop", "nfftcu_solver_step",
)


class NfftCuError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIBPATH):
            raise RuntimeError(f"{_LIBPATH} not built (run __graft_entry__.build()); no CPU fallback")
        L = C.CDLL(_LIBPATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        vp, i64, ci = C.c_void_p, C.c_int64, C.c_int
        L.nfftcu_last_error.restype = C.c_char_p
        L.nfftcu_device_count.restype = ci
        L.nfftcu_create.argtypes = [C.POINTER(vp), ci, ci, C.POINTER(i64), C.POINTER(i64), i64, i64,
                                    C.c_uint, ci]
        L.nfftcu_destroy.argtypes = [vp]
        L.nfftcu_get_c_phi_inv.argtypes = [vp, ci, vp]
        L.nfftcu_get_window_params.argtypes = [vp, vp, vp]
        for name in ("nfftcu_set_nodes", "nfftcu_set_nodes_dev", "nfftcu_get_index_x",
                     "nfftcu_stage_D", "nfftcu_stage_B", "nfftcu_stage_BT", "nfftcu_stage_DT",
                     "nfftcu_set_stream"):
            getattr(L, name).argtypes = [vp, vp]
        for name in ("nfftcu_trafo", "nfftcu_adjoint", "nfftcu_trafo_direct", "nfftcu_adjoint_direct",
                     "nfftcu_trafo_dev", "nfftcu_adjoint_dev", "nfftcu_trafo_direct_dev",
                     "nfftcu_adjoint_direct_dev"):
            getattr(L, name).argtypes = [vp, vp, vp]
        L.nfftcu_nodes_version.argtypes = [vp]
        L.nfftcu_nodes_version.restype = i64
        L.nfftcu_stage_F.argtypes = [vp, ci]
        L.nfftcu_grid_ptr.argtypes = [vp]
        L.nfftcu_grid_ptr.restype = vp
        L.nfftcu_set_option.argtypes = [vp, ci, i64]
        L.nfftcu_get_stream.argtypes = [vp]
        L.nfftcu_get_stream.restype = vp
        L.nfftcu_sync.argtypes = [vp]
        L.nfftcu_stage_times.argtypes = [vp, C.POINTER(C.c_float)]
        L.nfftcu_b_kernel_time.argtypes = [vp, C.POINTER(C.c_float)]
        L.nfftcu_launch_count.argtypes = [vp]
        L.nfftcu_launch_count.restype = i64
        L.nfftcu_malloc_device.argtypes = [C.POINTER(vp), C.c_size_t, ci]
        L.nfftcu_free_device.argtypes = [vp]
        L.nfftcu_malloc_pinned.argtypes = [C.POINTER(vp), C.c_size_t]
        L.nfftcu_free_pinned.argtypes = [vp]
        L.nfftcu_memcpy_h2d.argtypes = [vp, vp, C.c_size_t]
        L.nfftcu_memcpy_d2h.argtypes = [vp, vp, C.c_size_t]
        _LIB = L
    return _LIB


def _ck(status: int):
    if status != 0:
        raise NfftCuError(f"nfftcu error {status}: {lib().nfftcu_last_error().decode()}")


def _ptr(a) -> C.c_void_p:
    """numpy array, torch tensor (data_ptr) or raw int address -> void*"""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


class Engine:
    """One device plan context (``nfftcu_ctx``)."""

    def __init__(self, N: Sequence[int], n: Sequence[int], m: int, M: int, *,
                 precision: str = "double", flags: int = 0, device: int = 0):
        self.L = lib()
        self.precision = precision
        self.prec = DOUBLE if precision == "double" else FLOAT
        self.real = np.float64 if precision == "double" else np.float32
        self.cplx = np.complex128 if precision == "double" else np.complex64
        self.d = len(N)
        self.N, self.n, self.m, self.M = [int(v) for v in N], [int(v) for v in n], int(m), int(M)
        self.N_total = int(np.prod(self.N))
        self.n_total = int(np.prod(self.n))
        self.ctx = C.c_void_p(0)
        Na = (C.c_int64 * self.d)(*self.N)
        na = (C.c_int64 * self.d)(*self.n)
        _ck(self.L.nfftcu_create(C.byref(self.ctx), self.prec, self.d, Na, na, self.m, self.M,
                                 flags, device))

    def close(self):
        if self.ctx:
            _ck(self.L.nfftcu_destroy(self.ctx))
            self.ctx = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, opt: int, value: int):
        _ck(self.L.nfftcu_set_option(self.ctx, opt, value))

    def set_stream(self, cuda_stream: int):
        _ck(self.L.nfftcu_set_stream(self.ctx, C.c_void_p(cuda_stream)))

    def sync(self):
        _ck(self.L.nfftcu_sync(self.ctx))

    @property
    def launches(self) -> int:
        return int(self.L.nfftcu_launch_count(self.ctx))

    @property
    def nodes_version(self) -> int:
        return int(self.L.nfftcu_nodes_version(self.ctx))

    def stage_times(self):
        ms = (C.c_float * 3)()
        _ck(self.L.nfftcu_stage_times(self.ctx, ms))
        return [float(v) for v in ms]

    def b_kernel_time(self) -> float:
        ms = C.c_float(0)
        _ck(self.L.nfftcu_b_kernel_time(self.ctx, C.byref(ms)))
        return float(ms.value)

    # ---- nodes ----
    def set_nodes(self, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=self.real)
        assert x.size == self.M * self.d
        _ck(self.L.nfftcu_set_nodes(self.ctx, _ptr(x)))

    def set_nodes_dev(self, x_dev):
        _ck(self.L.nfftcu_set_nodes_dev(self.ctx, _ptr(x_dev)))

    def index_x(self) -> np.ndarray:
        out = np.empty((max(self.M, 1), 2), dtype=np.int64)
        _ck(self.L.nfftcu_get_index_x(self.ctx, _ptr(out)))
        return out[: self.M]

    def c_phi_inv(self, t: int) -> np.ndarray:
        out = np.empty(self.N[t], dtype=self.real)
        _ck(self.L.nfftcu_get_c_phi_inv(self.ctx, t, _ptr(out)))
        return out

    # ---- host-buffer transforms ----
    def _host(self, fn, src, n_out):
        src = np.ascontiguousarray(src, dtype=self.cplx)
        out = np.empty(max(n_out, 1), dtype=self.cplx)
        _ck(fn(self.ctx, _ptr(src), _ptr(out)))
        return out[:n_out]

    def trafo(self, f_hat): return self._host(self.L.nfftcu_trafo, f_hat, self.M)
    def adjoint(self, f): return self._host(self.L.nfftcu_adjoint, f, self.N_total)
    def trafo_direct(self, f_hat): return self._host(self.L.nfftcu_trafo_direct, f_hat, self.M)
    def adjoint_direct(self, f): return self._host(self.L.nfftcu_adjoint_direct, f, self.N_total)

    # ---- device-pointer transforms (async on the plan's stream) ----
    def trafo_dev(self, f_hat_dev, f_dev): _ck(self.L.nfftcu_trafo_dev(self.ctx, _ptr(f_hat_dev), _ptr(f_dev)))
    def adjoint_dev(self, f_dev, f_hat_dev): _ck(self.L.nfftcu_adjoint_dev(self.ctx, _ptr(f_dev), _ptr(f_hat_dev)))

    # ---- single stages ----
    def stage_D(self, f_hat_dev): _ck(self.L.nfftcu_stage_D(self.ctx, _ptr(f_hat_dev)))
    def stage_F(self, sign: int): _ck(self.L.nfftcu_stage_F(self.ctx, sign))
    def stage_B(self, f_dev): _ck(self.L.nfftcu_stage_B(self.ctx, _ptr(f_dev)))
    def stage_BT(self, f_dev): _ck(self.L.nfftcu_stage_BT(self.ctx, _ptr(f_dev)))
    def stage_DT(self, f_hat_dev): _ck(self.L.nfftcu_stage_DT(self.ctx, _ptr(f_hat_dev)))

    def grid_ptr(self) -> int:
        return int(self.L.nfftcu_grid_ptr(self.ctx) or 0)

    def grid_to_host(self) -> np.ndarray:
        out = np.empty(self.n_total, dtype=self.cplx)
        self.sync()
        _ck(self.L.nfftcu_memcpy_d2h(_ptr(out), C.c_void_p(self.grid_ptr()), out.nbytes))
        return out

    def grid_from_host(self, g: np.ndarray):
        g = np.ascontiguousarray(g, dtype=self.cplx).ravel()
        assert g.size == self.n_total
        self.sync()
        _ck(self.L.nfftcu_memcpy_h2d(C.c_void_p(self.grid_ptr()), _ptr(g), g.nbytes))


class DeviceBuffer:
    """Raw device allocation through the C ABI (for callers without torch)."""

    def __init__(self, nbytes: int, device: int = 0):
        self.ptr = C.c_void_p(0)
        self.nbytes = nbytes
        _ck(lib().nfftcu_malloc_device(C.byref(self.ptr), nbytes, device))

    def data_ptr(self) -> int:
        return int(self.ptr.value or 0)

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        _ck(lib().nfftcu_memcpy_h2d(self.ptr, _ptr(a), a.nbytes))
        return self

    def download(self, dtype, count) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        _ck(lib().nfftcu_memcpy_d2h(_ptr(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().nfftcu_free_device(self.ptr)
            self.ptr = C.c_void_p(0)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
