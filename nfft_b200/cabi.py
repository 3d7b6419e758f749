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
PEER_HANDLE_BYTES = 192
OPT_TIMING, OPT_PSI_TABLE, OPT_B_KERNEL, OPT_NODE_ORDER, OPT_B_FLUSH, OPT_FFT_PRUNE, OPT_FFT_KERNEL, OPT_WINDOW_IMAGES = 1, 2, 3, 4, 5, 6, 7, 8
OPT_SLAB_FFT = 9
OPT_TC5 = 10
FLAG_GAUSSIAN = 1 << 30

# every symbol include/nfftcu.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = (
    "nfftcu_last_error", "nfftcu_device_count", "nfftcu_create", "nfftcu_destroy",
    "nfftcu_get_c_phi_inv", "nfftcu_get_window_params", "nfftcu_get_window_scale", "nfftcu_set_nodes", "nfftcu_set_nodes_dev",
    "nfftcu_nodes_version", "nfftcu_get_index_x", "nfftcu_trafo", "nfftcu_adjoint",
    "nfftcu_trafo_direct", "nfftcu_adjoint_direct", "nfftcu_trafo_begin", "nfftcu_adjoint_begin", "nfftcu_end", "nfftcu_trafo_dev", "nfftcu_adjoint_dev",
    "nfftcu_trafo_direct_dev", "nfftcu_adjoint_direct_dev", "nfftcu_stage_D", "nfftcu_stage_F",
    "nfftcu_stage_B", "nfftcu_stage_BT", "nfftcu_stage_DT", "nfftcu_grid_ptr", "nfftcu_set_option",
    "nfftcu_set_stream", "nfftcu_get_stream", "nfftcu_sync", "nfftcu_stage_times", "nfftcu_b_kernel_time", "nfftcu_trafo_refresh", "nfftcu_adjoint_refresh",
    "nfftcu_launch_count", "nfftcu_malloc_device", "nfftcu_free_device", "nfftcu_malloc_pinned",
    "nfftcu_free_pinned", "nfftcu_memcpy_h2d", "nfftcu_memcpy_d2h",
    "nfftcu_solver_create", "nfftcu_solver_create_batch", "nfftcu_solver_destroy", "nfftcu_solver_upload", "nfftcu_solver_download",
    "nfftcu_solver_vector", "nfftcu_solver_before_loop", "nfftcu_solver_step",
    "nfftcu_tc5_debug", "nfftcu_measure_peaks", "nfftcu_host_alloc", "nfftcu_host_free", "nfftcu_pool_trim", "nfftcu_fingerprint",
    "nfftcu_trafo_batch", "nfftcu_adjoint_batch", "nfftcu_trafo_batch_dev", "nfftcu_adjoint_batch_dev",
    "nfftcu_mri_inh_2d1d", "nfftcu_mri_inh_3d", "nfftcu_adjoint_mul_trafo",
    "nfftcu_get_sorted_slab", "nfftcu_peer_export", "nfftcu_peer_attach", "nfftcu_peer_detach",
    "nfftcu_adjoint_dev_peer", "nfftcu_peer_error", "nfftcu_peer_reduce_only",
    "nfftcu_group_create", "nfftcu_group_destroy", "nfftcu_group_set_nodes", "nfftcu_group_nodes_version",
    "nfftcu_group_get_index_x", "nfftcu_group_trafo", "nfftcu_group_adjoint", "nfftcu_group_trafo_refresh",
    "nfftcu_group_adjoint_refresh", "nfftcu_group_direct", "nfftcu_group_size", "nfftcu_group_ctx",
    "nfftcu_group_times",
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
        L.nfftcu_get_window_scale.argtypes = [vp, C.POINTER(C.c_double)]
        for name in ("nfftcu_set_nodes", "nfftcu_set_nodes_dev", "nfftcu_get_index_x",
                     "nfftcu_stage_D", "nfftcu_stage_B", "nfftcu_stage_BT", "nfftcu_stage_DT",
                     "nfftcu_set_stream"):
            getattr(L, name).argtypes = [vp, vp]
        for name in ("nfftcu_trafo", "nfftcu_adjoint", "nfftcu_trafo_direct", "nfftcu_adjoint_direct", "nfftcu_trafo_begin", "nfftcu_adjoint_begin", "nfftcu_end",
                     "nfftcu_trafo_dev", "nfftcu_adjoint_dev", "nfftcu_trafo_direct_dev",
                     "nfftcu_adjoint_direct_dev"):
            getattr(L, name).argtypes = [vp, vp, vp]
        for name in ("nfftcu_trafo_begin", "nfftcu_adjoint_begin"):
            getattr(L, name).argtypes = [vp, vp, vp]
        L.nfftcu_end.argtypes = [vp]
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
        L.nfftcu_measure_peaks.argtypes = [ci, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.nfftcu_host_alloc.argtypes = [C.c_size_t]
        L.nfftcu_host_alloc.restype = vp
        L.nfftcu_host_free.argtypes = [vp]
        L.nfftcu_host_free.restype = None
        L.nfftcu_pool_trim.restype = None
        L.nfftcu_fingerprint.argtypes = [vp, C.c_size_t]
        L.nfftcu_fingerprint.restype = C.c_uint64
        for name in ("nfftcu_trafo_batch", "nfftcu_adjoint_batch", "nfftcu_trafo_batch_dev", "nfftcu_adjoint_batch_dev"):
            getattr(L, name).argtypes = [vp, ci, vp, vp]
        L.nfftcu_adjoint_mul_trafo.argtypes = [vp, vp, vp, vp, vp]
        L.nfftcu_get_sorted_slab.argtypes = [vp, i64, i64, vp, vp]
        L.nfftcu_peer_export.argtypes = [vp, vp]
        L.nfftcu_peer_attach.argtypes = [vp, ci, ci, vp]
        L.nfftcu_peer_detach.argtypes = [vp]
        L.nfftcu_adjoint_dev_peer.argtypes = [vp, vp, vp]
        L.nfftcu_peer_error.argtypes = [vp]
        L.nfftcu_peer_reduce_only.argtypes = [vp, vp]
        L.nfftcu_group_create.argtypes = [C.POINTER(vp), ci, ci, C.POINTER(i64), C.POINTER(i64), i64, i64, C.c_uint,
                                          C.POINTER(ci), ci]
        L.nfftcu_group_destroy.argtypes = [vp]
        L.nfftcu_group_set_nodes.argtypes = [vp, vp]
        L.nfftcu_group_nodes_version.argtypes = [vp]
        L.nfftcu_group_nodes_version.restype = i64
        L.nfftcu_group_get_index_x.argtypes = [vp, vp]
        for name in ("nfftcu_group_trafo", "nfftcu_group_adjoint"):
            getattr(L, name).argtypes = [vp, vp, vp]
        for name in ("nfftcu_group_trafo_refresh", "nfftcu_group_adjoint_refresh"):
            getattr(L, name).argtypes = [vp, vp, vp, vp, C.POINTER(ci)]
        L.nfftcu_group_direct.argtypes = [vp, ci, vp, vp]
        L.nfftcu_group_size.argtypes = [vp]
        L.nfftcu_group_ctx.argtypes = [vp, ci]
        L.nfftcu_group_ctx.restype = vp
        L.nfftcu_group_times.argtypes = [vp, C.POINTER(C.c_float)]
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


def measure_peaks(device: int = 0):
    """(fp64 tensor TFLOP/s, mma.sync TF32 TFLOP/s, device copy GB/s) measured now on `device` (peaks.cu)."""
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    _ck(lib().nfftcu_measure_peaks(device, C.byref(a), C.byref(b), C.byref(c)))
    return float(a.value), float(b.value), float(c.value)


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
        """Run on the given cudaStream_t.  torch reports its default stream as handle 0, which in the C ABI means
        "own stream"; the legacy default stream is therefore passed as cudaStreamLegacy (0x1).  ``None`` restores
        the engine's own (non-blocking) stream."""
        if cuda_stream is None:
            _ck(self.L.nfftcu_set_stream(self.ctx, C.c_void_p(0)))
        else:
            _ck(self.L.nfftcu_set_stream(self.ctx, C.c_void_p(int(cuda_stream) or 1)))

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

    def window_scale(self) -> float:
        """S = prod_t s_t: the power-of-two factor the device window carries (1.0 for fp64 plans); single-stage calls
        see the grid in these units (include/nfftcu.h, nfftcu_get_window_scale)"""
        s = (C.c_double * self.d)()
        _ck(self.L.nfftcu_get_window_scale(self.ctx, s))
        return float(np.prod([float(v) for v in s]))

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

    # ---- K right-hand sides on one node set ([K][N_total] / [K][M]) ----
    def trafo_batch(self, f_hat) -> np.ndarray:
        f_hat = np.ascontiguousarray(f_hat, dtype=self.cplx).reshape(-1, self.N_total)
        out = np.empty((f_hat.shape[0], max(self.M, 1)), dtype=self.cplx)
        _ck(self.L.nfftcu_trafo_batch(self.ctx, f_hat.shape[0], _ptr(f_hat), _ptr(out)))
        return out[:, : self.M]

    def adjoint_batch(self, f) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=self.cplx).reshape(-1, self.M)
        out = np.empty((f.shape[0], self.N_total), dtype=self.cplx)
        _ck(self.L.nfftcu_adjoint_batch(self.ctx, f.shape[0], _ptr(f), _ptr(out)))
        return out

    def trafo_batch_dev(self, K, f_hat_dev, f_dev): _ck(self.L.nfftcu_trafo_batch_dev(self.ctx, K, _ptr(f_hat_dev), _ptr(f_dev)))
    def adjoint_batch_dev(self, K, f_dev, f_hat_dev): _ck(self.L.nfftcu_adjoint_batch_dev(self.ctx, K, _ptr(f_dev), _ptr(f_hat_dev)))

    def adjoint_mul_trafo(self, dst: "Engine", f_src, b=None) -> np.ndarray:
        """f_dst = A_dst (b .* A_self^H f_src), f_hat never leaves the device (fastsum far field)"""
        f_src = np.ascontiguousarray(f_src, dtype=self.cplx)
        bb = None if b is None else np.ascontiguousarray(b, dtype=self.cplx)
        out = np.empty(max(dst.M, 1), dtype=self.cplx)
        _ck(self.L.nfftcu_adjoint_mul_trafo(self.ctx, dst.ctx, _ptr(f_src), _ptr(bb), _ptr(out)))
        return out[: dst.M]

    # ---- device-pointer transforms (async on the plan's stream) ----
    def trafo_dev(self, f_hat_dev, f_dev): _ck(self.L.nfftcu_trafo_dev(self.ctx, _ptr(f_hat_dev), _ptr(f_dev)))
    def adjoint_dev(self, f_dev, f_hat_dev): _ck(self.L.nfftcu_adjoint_dev(self.ctx, _ptr(f_dev), _ptr(f_hat_dev)))

    # ---- multi-GPU building blocks (include/nfftcu.h "multi-GPU, node-sharded") ----
    def sorted_slab(self, begin: int, end: int, x_out_dev, perm_out_dev):
        """reference-sorted nodes [begin, end) and their original indices -> device buffers (async on the stream)"""
        _ck(self.L.nfftcu_get_sorted_slab(self.ctx, int(begin), int(end), _ptr(x_out_dev), _ptr(perm_out_dev)))

    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        _ck(self.L.nfftcu_peer_export(self.ctx, buf))
        return buf.raw

    def peer_attach(self, rank: int, world: int, all_handles: bytes):
        assert len(all_handles) == world * PEER_HANDLE_BYTES
        _ck(self.L.nfftcu_peer_attach(self.ctx, rank, world, C.c_char_p(all_handles)))

    def adjoint_dev_peer(self, f_dev, f_hat_dev):
        """adjoint with the cross-GPU reduction fused into D^T over peer memory (peer.cu); all ranks call it"""
        _ck(self.L.nfftcu_adjoint_dev_peer(self.ctx, _ptr(f_dev), _ptr(f_hat_dev)))

    def peer_reduce_only(self, f_hat_dev):
        _ck(self.L.nfftcu_peer_reduce_only(self.ctx, _ptr(f_hat_dev)))

    def peer_error(self) -> int:
        return int(self.L.nfftcu_peer_error(self.ctx))

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


SOLVER_Y, SOLVER_W, SOLVER_W_HAT, SOLVER_F_HAT_ITER, SOLVER_R_ITER, SOLVER_Z_HAT_ITER, SOLVER_P_HAT_ITER, SOLVER_V_ITER = range(8)
LANDWEBER, STEEPEST_DESCENT, CGNR, CGNE, NORMS_FOR_LANDWEBER, PRECOMPUTE_WEIGHT, PRECOMPUTE_DAMP = (1 << i for i in range(7))


class BatchSolver:
    """Device-resident inverse NFFT for K right-hand sides on one plan (``nfftcu_solver_create_batch``)."""

    def __init__(self, engine: "Engine", flags: int, K: int):
        self.L, self.eng, self.K, self.flags = lib(), engine, int(K), flags
        self.s = C.c_void_p(0)
        self.L.nfftcu_solver_create_batch.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_uint, C.c_int]
        for name in ("nfftcu_solver_upload", "nfftcu_solver_download"):
            getattr(self.L, name).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for name in ("nfftcu_solver_before_loop", "nfftcu_solver_step"):
            getattr(self.L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        self.L.nfftcu_solver_destroy.argtypes = [C.c_void_p]
        _ck(self.L.nfftcu_solver_create_batch(C.byref(self.s), engine.ctx, flags, self.K))
        self.scal = np.zeros((self.K, 8))

    def upload(self, which: int, a: np.ndarray):
        real = which in (SOLVER_W, SOLVER_W_HAT)
        a = np.ascontiguousarray(a, dtype=self.eng.real if real else self.eng.cplx)
        _ck(self.L.nfftcu_solver_upload(self.s, which, _ptr(a)))

    def download(self, which: int) -> np.ndarray:
        n = self.eng.M if which in (SOLVER_Y, SOLVER_R_ITER, SOLVER_V_ITER) else self.eng.N_total
        out = np.empty((self.K, n), dtype=self.eng.cplx)
        _ck(self.L.nfftcu_solver_download(self.s, which, _ptr(out)))
        return out

    def before_loop(self):
        _ck(self.L.nfftcu_solver_before_loop(self.s, None, None, self.scal.ctypes.data_as(C.POINTER(C.c_double))))
        return self.scal.copy()

    def step(self):
        _ck(self.L.nfftcu_solver_step(self.s, None, None, self.scal.ctypes.data_as(C.POINTER(C.c_double))))
        return self.scal.copy()

    def close(self):
        if self.s:
            _ck(self.L.nfftcu_solver_destroy(self.s))
            self.s = C.c_void_p(0)


class Group:
    """ONE process driving several GPUs (``nfftcu_group_*``): node-sharded trafo / adjoint on HOST arrays."""

    def __init__(self, N: Sequence[int], n: Sequence[int], m: int, M: int, devices: Sequence[int], *,
                 precision: str = "double", flags: int = 0):
        self.L = lib()
        self.precision = precision
        self.real = np.float64 if precision == "double" else np.float32
        self.cplx = np.complex128 if precision == "double" else np.complex64
        self.d, self.M, self.N_total = len(N), int(M), int(np.prod(N))
        self.g = C.c_void_p(0)
        Na = (C.c_int64 * self.d)(*[int(v) for v in N])
        na = (C.c_int64 * self.d)(*[int(v) for v in n])
        dv = (C.c_int * len(devices))(*[int(v) for v in devices])
        _ck(self.L.nfftcu_group_create(C.byref(self.g), DOUBLE if precision == "double" else FLOAT, self.d, Na, na,
                                       int(m), self.M, flags, dv, len(devices)))

    def close(self):
        if self.g:
            _ck(self.L.nfftcu_group_destroy(self.g))
            self.g = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_nodes(self, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=self.real)
        assert x.size == self.M * self.d
        _ck(self.L.nfftcu_group_set_nodes(self.g, _ptr(x)))

    def index_x(self) -> np.ndarray:
        out = np.empty((max(self.M, 1), 2), dtype=np.int64)
        _ck(self.L.nfftcu_group_get_index_x(self.g, _ptr(out)))
        return out[: self.M]

    def trafo(self, f_hat, out=None):
        f_hat = np.ascontiguousarray(f_hat, dtype=self.cplx)
        out = np.empty(max(self.M, 1), dtype=self.cplx) if out is None else out
        _ck(self.L.nfftcu_group_trafo(self.g, _ptr(f_hat), _ptr(out)))
        return out[: self.M]

    def adjoint(self, f, out=None):
        f = np.ascontiguousarray(f, dtype=self.cplx)
        out = np.empty(self.N_total, dtype=self.cplx) if out is None else out
        _ck(self.L.nfftcu_group_adjoint(self.g, _ptr(f), _ptr(out)))
        return out

    def times(self):
        ms = (C.c_float * 3)()
        _ck(self.L.nfftcu_group_times(self.g, ms))
        return [float(v) for v in ms]


def host_alloc(nbytes: int) -> int:
    """page-locked host memory from the library's nfft_malloc backing store (address)"""
    p = lib().nfftcu_host_alloc(nbytes)
    if not p:
        raise MemoryError(nbytes)
    return int(p)


def host_free(addr: int):
    lib().nfftcu_host_free(C.c_void_p(addr))


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
