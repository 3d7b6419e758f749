"""debug: where do the NaNs of the fp32 adjoint at M=1e7 come from?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from nfft_b200 import cabi
import common
N, n, m = [128] * 3, [256] * 3, 6
rng = np.random.Generator(np.random.Philox(20260103))
Mmax = 10_000_000
x = (rng.random((Mmax, 3)) - 0.5).astype(np.float32)
x = np.minimum(x, np.nextafter(np.float32(0.5), np.float32(0)))
fh = rng.random((128 ** 3, 2)).astype(np.float32)
f = rng.random((Mmax, 2)).astype(np.float32)
for M in (2_000_000, 5_000_000, 8_000_000, 10_000_000):
    for opts in ({}, {cabi.OPT_WINDOW_IMAGES: 1}, {cabi.OPT_B_FLUSH: 1}, {cabi.OPT_B_KERNEL: 3}):
        eng = cabi.Engine(N, n, m, M, precision="float")
        for k, v in opts.items():
            eng.set_option(k, v)
        eng.set_nodes(x[:M])
        fc = f[:M].copy().view(np.complex64).ravel()
        out = eng.adjoint(fc)
        bad = ~np.isfinite(out.view(np.float32))
        fbuf = cabi.DeviceBuffer(fc.nbytes).upload(fc)
        eng.stage_BT(fbuf)
        g = eng.grid_to_host()
        gbad = ~np.isfinite(g.view(np.float32))
        msg = ""
        if gbad.any():
            idx = np.flatnonzero(gbad.reshape(-1, 2).any(axis=1))
            i0, i1, i2 = np.unravel_index(idx, (256, 256, 256))
            msg = f" first bad cells {list(zip(i0[:5], i1[:5], i2[:5]))} ... i0 range [{i0.min()},{i0.max()}] i1 [{i1.min()},{i1.max()}] i2 [{i2.min()},{i2.max()}] max|g| {np.nanmax(np.abs(g.view(np.float32)[~gbad])):.3e}"
        print(f"M={M} opts={opts}: f_hat nonfinite {bad.sum()} of {bad.size}; grid nonfinite {gbad.sum()}{msg}", flush=True)
        fbuf.free()
        eng.close()
# the fp32 reference on the same data: finite?
if common.have_ref():
    from nfft_b200.plan import Plan
    from nfft_b200 import plan_abi as abi
    flags = (abi.PRE_PHI_HUT | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT | abi.NFFT_SORT_NODES | abi.NFFT_OMP_BLOCKWISE_ADJOINT)
    p = Plan.init_guru(3, N, Mmax, n, m, flags, api=common.ref_api("float"))
    p.x[:] = x
    p.f.view(np.float32)[:] = f.ravel()
    p.adjoint()
    r = p.f_hat.copy()
    print("reference fp32 adjoint nonfinite:", (~np.isfinite(r.view(np.float32))).sum(), "max", np.abs(r).max())
    p.finalize()
