"""Quick GPU check of the tcgen05 fp32 kernels (tc5.cu) against the mma.sync kernels and the oracle.
Usage: timeout 300 python tools/tc5_check.py [--big]"""
import os, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from nfft_b200 import cabi
import common

def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

CASES = [
    dict(N=[16, 16, 16], n=[32, 32, 32], m=6, M=2000),
    dict(N=[32, 32, 32], n=[64, 64, 64], m=6, M=40000),
    dict(N=[24, 20, 18], n=[48, 40, 36], m=4, M=6000),
    dict(N=[16, 16, 32], n=[32, 32, 70], m=2, M=5000),
    dict(N=[64, 64, 64], n=[128, 128, 128], m=6, M=300000),
]
if "--big" in sys.argv:
    CASES = [dict(N=[128, 128, 128], n=[256, 256, 256], m=6, M=10_000_000)]
rng = np.random.default_rng(5)
o = common.oracle("float")
for spec in CASES:
    M, NN = spec["M"], int(np.prod(spec["N"]))
    x = (rng.random((M, 3)) - 0.5).astype(np.float32)
    x = np.minimum(x, np.nextafter(np.float32(0.5), np.float32(0)))
    fh = (rng.random(NN) - 0.5 + 1j * (rng.random(NN) - 0.5)).astype(np.complex64)
    f = (rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)).astype(np.complex64)
    res = {}
    for tc5 in (1, 3):
        eng = cabi.Engine(spec["N"], spec["n"], spec["m"], M, precision="float")
        eng.set_option(cabi.OPT_TC5, tc5)
        eng.set_option(cabi.OPT_TIMING, 1)
        eng.set_nodes(x)
        t0 = time.time()
        got_f = eng.trafo(fh)
        tb = eng.b_kernel_time()
        if tc5 == 3 and os.environ.get("NFFT_B200_TC5_DBG") and int(os.environ["NFFT_B200_TC5_DBG"]) & 8:
            import ctypes
            buf = (ctypes.c_ulonglong * 32)()
            cabi.lib().nfftcu_tc5_debug(buf)
            d = list(buf)
            nbm = max(d[20], 1)   # own batches of MMA warp 0
            nb = nbm
            print("  per batch (cycles): MMA warp 0 (per own batch): wait op_full %.0f, acc_empty %.0f, a_ready %.0f, issue %.0f | epilogue warp 0 (per own batch): "
                  "wait acc_full %.0f, ld %.0f, math+butterfly %.0f, cross-warp %.0f, loop total/batch %.0f | refill PER SLIDE: hazard wait %.0f, cvt+st issue %.0f, "
                  "prefetch+wait::st %.0f, slides %.2f/batch | feeder wait op_empty %.0f" % (
                      d[16] / nbm, d[17] / nbm, d[18] / nbm, d[19] / nbm, d[0] / max(d[3], 1), d[1] / max(d[3], 1), d[2] / max(d[3], 1),
                      d[4] / max(d[3], 1), d[5] / nb, d[8] / max(d[11], 1), d[9] / max(d[11], 1), d[10] / max(d[11], 1), d[11] / nb, d[24] / nb))
        got_fh = eng.adjoint(f)
        tbt = eng.b_kernel_time()
        res[tc5] = (got_f, got_fh, tb, tbt)
        eng.close()
    line = "N=%s m=%d M=%d: trafo tc5 vs legacy %.2e, adjoint %.2e; B kernel %.3f ms (legacy %.3f), BT %.3f (legacy %.3f)" % (
        spec["N"], spec["m"], M, rel(res[3][0], res[1][0]), rel(res[3][1], res[1][1]), res[3][2], res[1][2], res[3][3], res[1][3])
    if M <= 400000:
        want_f = o.trafo(spec["N"], spec["n"], spec["m"], x, fh)
        want_fh = o.adjoint(spec["N"], spec["n"], spec["m"], x, f, True)
        line += "; vs oracle: trafo %.2e (legacy %.2e) adjoint %.2e (legacy %.2e)" % (
            rel(res[3][0], want_f), rel(res[1][0], want_f), rel(res[3][1], want_fh), rel(res[1][1], want_fh))
    print(line, flush=True)
    bad = np.flatnonzero(~np.isfinite(res[3][0]))
    if bad.size:
        print("  non-finite entries:", bad.size, bad[:10])
