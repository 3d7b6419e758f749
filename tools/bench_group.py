#!/usr/bin/env python
"""tools/bench_group.py -- the multi-GPU path a C caller sees: nfft_trafo / nfft_adjoint of libnfft3_b200.so with
NFFT_B200_DEVICES=0,..,P-1 (one process, P GPUs, nfftcu_group_*: slab-partitioned nodes, all-to-all permutation
over peer memory, fused D^T + reduce-scatter, every GPU moving its share of f over its own host link).
Strong scaling of one node set: prints one JSON line per device count.

    python tools/bench_group.py --config cfg4 --devices 1,2,4,8 [--check]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3", choices=sorted(bench.CFGS))
    ap.add_argument("--devices", default="1,2")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--precision", default="double")
    ap.add_argument("--check", action="store_true", help="rel l2 against the reference (oracle/_ref) once")
    ap.add_argument("--nodes", type=int, default=0)
    args = ap.parse_args()
    from nfft_b200 import cabi, plan_abi as abi
    from nfft_b200.plan import Plan
    cfg = dict(bench.CFGS[args.config])
    if args.nodes:
        cfg["M"] = args.nodes
    prec = args.precision
    x, fh, f = bench.synth(cfg, prec, 0)
    ref = None
    if args.check:
        rp = bench.ReferencePlan(cfg, prec, x, fh, f)
        t_ref = sum(rp.pair())
        ref = (rp.f_out.copy(), rp.fh_out.copy(), t_ref)
        rp.close()
    flags = (abi.PRE_PHI_HUT | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT
             | abi.NFFT_SORT_NODES | abi.NFFT_OMP_BLOCKWISE_ADJOINT)
    base = None
    for P in [int(v) for v in args.devices.split(",")]:
        if P > 1:
            os.environ["NFFT_B200_DEVICES"] = ",".join(str(i) for i in range(P))
        else:
            os.environ.pop("NFFT_B200_DEVICES", None)
        # two plans as in bench.py's e2e leg: trafo reads p.f_hat -> p.f, adjoint reads q.f -> q.f_hat
        p = Plan.init_guru(3, cfg["N"], cfg["M"], cfg["n"], cfg["m"], flags, precision=prec)
        p.x[:] = x
        p.f_hat.view(p.api.real)[:] = fh.ravel()
        t0 = time.perf_counter()
        p.trafo()                         # first call: node upload, global sort, slab distribution, window images
        t_first = time.perf_counter() - t0
        p.f.view(p.api.real)[:] = f.ravel()
        p.adjoint()
        tt, ta = [], []
        for _ in range(args.steps):
            p.f_hat.view(p.api.real)[:] = fh.ravel()
            t0 = time.perf_counter()
            p.trafo()
            t1 = time.perf_counter()
            f_out = p.f.copy()
            p.f.view(p.api.real)[:] = f.ravel()
            t2 = time.perf_counter()
            p.adjoint()
            t3 = time.perf_counter()
            tt.append(t1 - t0)
            ta.append(t3 - t2)
        fh_out = p.f_hat.copy()
        phases = None
        if P > 1 and p.c.my_fftw_plan2:      # nfftcu_group_times of the last transform (the adjoint)
            import ctypes as C
            ms = (C.c_float * 3)()
            cabi.lib().nfftcu_group_times(C.c_void_p(p.c.my_fftw_plan2), ms)
            phases = dict(adjoint_h2d_ms=float(ms[0]), adjoint_compute_and_exchange_ms=float(ms[1]), adjoint_d2h_ms=float(ms[2]))
        line = dict(config=bench.workload_config(cfg, P, "strong")["workload"], n_gpus=P, precision=prec,
                    path="nfft_trafo + nfft_adjoint of libnfft3_b200.so on plan-API (page-locked) host buffers, "
                         "NFFT_B200_DEVICES=%s" % os.environ.get("NFFT_B200_DEVICES", "(unset: one device)"),
                    ms_trafo=float(np.median(tt)) * 1e3, ms_adjoint=float(np.median(ta)) * 1e3,
                    ms_pair=float(np.median(tt) + np.median(ta)) * 1e3,
                    points_per_s=cfg["M"] / float(np.median(tt) + np.median(ta)),
                    first_call_s=t_first)
        if base is None:
            base = line["ms_pair"]
        line["speedup_vs_first_listed"] = base / line["ms_pair"]
        if phases:
            line["phases_slowest_device"] = phases
        if ref is not None:
            line["rel_l2"] = dict(trafo=bench.rel_l2(f_out.view(p.api.real), ref[0].view(p.api.real)),
                                  adjoint=bench.rel_l2(fh_out.view(p.api.real), ref[1].view(p.api.real)))
            line["reference_cpu_pair_s"] = ref[2]
        p.finalize()
        cabi.lib().nfftcu_pool_trim()
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
