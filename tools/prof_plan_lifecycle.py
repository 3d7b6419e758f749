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
