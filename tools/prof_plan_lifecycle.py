#!/usr/bin/env python
"""Times the plan life cycle through libnfft3_b200.so (init_guru, node upload + precompute, transforms, finalize)
for the cfg2 / cfg5 shape; used to find host-side overheads of plan-per-coil workloads."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from nfft_b200 import plan_abi as abi
from nfft_b200.plan import Plan
from bench_configs import spiral

flags = abi.PRE_PHI_HUT | abi.PRE_PSI | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT
x = spiral(512 * 512, 512)
rng = np.random.default_rng(0)
for rep in range(4):
    t = [time.perf_counter()]
    p = Plan.init_guru(2, [512, 512], 512 * 512, [1024, 1024], 6, flags)
    t.append(time.perf_counter())
    p.x[:] = x
    p.precompute_one_psi()
    t.append(time.perf_counter())
    p.f_hat[:] = rng.random(512 * 512)
    p.trafo(); p.adjoint()
    t.append(time.perf_counter())
    for _ in range(20):
        p.trafo(); p.adjoint()
    t.append(time.perf_counter())
    p.finalize()
    t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("rep %d: init %.1f ms, nodes+precompute %.1f ms, first pair %.1f ms, 20 pairs %.1f ms (%.2f ms/pair), finalize %.1f ms"
          % (rep, d[0], d[1], d[2], d[3], d[3] / 20, d[4]), flush=True)

# the device-resident solver on the same shape: per-call times of solver_init / before_loop / 20 steps / finalize
import ctypes as C
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsolver_dev_b200.so"), mode=os.RTLD_LOCAL)
fn = L.solver_driver_run
fn.restype = C.c_int
M, NN = 512 * 512, 512 * 512
y = np.ascontiguousarray(rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5))
w = np.ones(M); w_hat = np.ones(NN)
f_hat = np.zeros(NN, dtype=np.complex128); dots = np.zeros(64)
ia = lambda a: (C.c_int * len(a))(*a)
p_ = lambda a: a.ctypes.data_as(C.c_void_p)
for iters in (1, 1, 20, 20, 40):
    t0 = time.perf_counter()
    fn(C.c_int(2), ia([512, 512]), C.c_int(M), ia([1024, 1024]), C.c_int(6), C.c_uint(flags), C.c_uint((1 << 2) | (1 << 6)),
       p_(x), p_(y), p_(w), p_(w_hat), C.c_int(iters), p_(f_hat), p_(dots), C.c_double(0.0), None)
    print("device solver driver, %2d iterations: %.1f ms" % (iters, (time.perf_counter() - t0) * 1e3), flush=True)
