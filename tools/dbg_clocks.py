#!/usr/bin/env python
"""Debug only: runs trafo at cfg3 with a libnfftcu.so built with -DNFFTCU_DBG_CLOCKS and prints the clock64 section
totals of MMA warp 0 / producer warp 0 of every CTA of interp_mma_kernel."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nfft_b200 import cabi
M = 10_000_000
rng = np.random.default_rng(0)
x = rng.random((M, 3)) - 0.5
eng = cabi.Engine([128] * 3, [256] * 3, 6, M)
eng.set_nodes(x)
dev = torch.device("cuda", 0)
fh = torch.randn(128 ** 3, dtype=torch.complex128, device=dev)
f = torch.empty(M, dtype=torch.complex128, device=dev)
L = cabi.lib()
out = (C.c_ulonglong * 16)()
for rep in range(3):
    eng.trafo_dev(fh, f)
    torch.cuda.synchronize()
    L.nfftcu_debug_clocks(out)
    v = list(out)
    nb = max(v[5], 1)
    print("rep %d: batches %d | per batch (MMA warp 0): wait-full %.0f, mma+weights %.0f, refill issue %.0f, tail %.0f, loop total %.0f clks"
          % (rep, v[5], v[0] / nb, v[1] / nb, v[2] / nb, v[3] / nb, v[4] / nb))
eng.close()
