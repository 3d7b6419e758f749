#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on a B200: 3-D NFFT trafo+adjoint points/s
(cfg3: N=128^3, n=256^3, M=1e7 uniform random nodes, m=6, fp64; rel l2 err against the reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision float]
                    [--config cfg3|cfg4] [--scaling weak|strong]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step = one nfft_trafo + one nfft_adjoint over the rank's nodes (+ the f_hat reduce when N > 1, node-sharded as in
SURVEY 8e).  `value` = nodes all ranks processed / max-over-ranks device time, inputs resident in HBM.
`e2e` = the same step through the reference-facing plan API (nfft_trafo / nfft_adjoint of libnfft3_b200.so) on HOST
buffers the plan API itself allocated (MALLOC_X | MALLOC_F_HAT | MALLOC_F), copies included.
`rel_l2` = ||ours - ref|| / ||ref|| of the TIMED outputs against the reference's own nfft_trafo / nfft_adjoint
(oracle/_ref, unmodified kernel/nfft/nfft.c) run on the same inputs at full size.
`roofline` = the dominant kernel (B^T spreading or B interpolation launch) against the measured HBM peak with the
SURVEY 8d byte model, the FP64-pipe view beside it; `roofline_pipeline` = the same for the whole pair.
`--impl reference` times the reference's own CPU implementation (oracle/_ref, OpenMP, all host threads; FFT stage is
the shim, not FFTW) on the same workload at full size, K measured steps.

Scaling modes: weak (default; every rank holds M nodes of its own drawn over the whole cube) and strong (--scaling
strong; the config's M nodes are sorted once and cut into equal-count slabs, one per rank).
"""
from __future__ import annotations

import os
import sys

_WORLD = int(os.environ.get("WORLD_SIZE", "1"))
_CORES = os.cpu_count() or 1
# torchrun exports OMP_NUM_THREADS=1; the reference (CPU checker / baseline) legs use the host cores, split over ranks
os.environ["OMP_NUM_THREADS"] = str(max(1, _CORES // _WORLD))

import argparse  # noqa: E402
import ctypes as C  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFGS = {
    "cfg3": dict(name="cfg3", d=3, N=[128, 128, 128], n=[256, 256, 256], m=6, M=10_000_000, seed=20260103),
    "cfg4": dict(name="cfg4", d=3, N=[256, 256, 256], n=[512, 512, 512], m=6, M=100_000_000, seed=20260104),
}
METRIC = "3D NFFT trafo+adjoint points/s (N=128^3, M=1e7, fp64, m=6); rel l2 err"
TOL = {"double": 1e-12, "float": 1e-5}


def metric_name(cfg, prec):
    if cfg["name"] == "cfg3" and prec == "double":
        return METRIC
    return "3D NFFT trafo+adjoint points/s (N=%d^3, M=%.0e, %s, m=%d); rel l2 err" % (
        cfg["N"][0], cfg["M"], "fp64" if prec == "double" else "fp32", cfg["m"])


def workload_config(cfg, world, scaling):
    """The `config` object: identical for our arm and the reference arm."""
    per = cfg["M"] if scaling == "weak" else cfg["M"] // world
    return {"workload": "%s: 3-D N=%d^3 n=%d^3 m=%d sigma=2, uniform random nodes, M=%d %s, reference flags "
                        "PRE_PHI_HUT|NFFT_SORT_NODES|NFFT_OMP_BLOCKWISE_ADJOINT (no PRE_PSI flag), "
                        "one nfft_trafo + one nfft_adjoint per step"
                        % (cfg["name"], cfg["N"][0], cfg["n"][0], cfg["m"], cfg["M"],
                           "per GPU" if scaling == "weak" else "in total"),
            "nodes_per_gpu": per, "n_gpus": world, "scaling": scaling,
            "l2": "inputs larger than L2 (grid %d MB, x %d MB, f %d MB per GPU vs 126 MB L2)"
                  % (int(np.prod(cfg["n"])) * 16 // 10 ** 6, per * 24 // 10 ** 6, per * 16 // 10 ** 6)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def byte_model(cfg, csize, M):
    """SURVEY 8d model v1.  C = complex bytes, R = C/2."""
    Cb, Rb, d = csize, csize // 2, cfg["d"]
    NN, nn = int(np.prod(cfg["N"])), int(np.prod(cfg["n"]))
    a_trafo = Cb * NN + Cb * nn + 2 * d * Cb * nn + Cb * nn + M * (Rb * d + Cb)
    a_adj = M * (Rb * d + Cb) + Cb * nn + 2 * d * Cb * nn + Cb * NN + Cb * NN
    a_spread = M * (Rb * d + Cb) + Cb * nn       # read x, f; flush the spread grid once
    a_interp = Cb * nn + M * (Rb * d + Cb)       # read the grid once; read x, write f
    return dict(trafo=a_trafo, adjoint=a_adj, pair=a_trafo + a_adj, spread=a_spread, interp=a_interp)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({nm for r in self.rows if len(r) >= 9 for nm, v in zip(names, r[5:9]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth(cfg, precision, rank, M=None):
    rng = np.random.Generator(np.random.Philox(cfg["seed"] + 1000 * rank))
    real = np.float64 if precision == "double" else np.float32
    M = cfg["M"] if M is None else M
    d, NN = cfg["d"], int(np.prod(cfg["N"]))
    x = (rng.random((M, d)) - 0.5).astype(real)
    if precision == "float":
        x = np.minimum(x, np.nextafter(np.float32(0.5), np.float32(0)))
    fh = rng.random((NN, 2)).astype(real)
    f = rng.random((M, 2)).astype(real)
    if precision == "float":
        # fp32: zero-mean data.  With U[0,1) samples the REFERENCE's nfftf_adjoint overflows at this size: the k = 0
        # bin of the oversampled spectrum is sum_j f_j * phi_hat(0)^3 = 5e6 * (1.5e11)^3 > FLT_MAX before the
        # deconvolution brings it back (checked: oracle/_ref returns inf there), so there would be nothing to compare
        # with.  (This engine folds phi_hat(0) out of the fp32 window as an exact power of two and stays finite.)
        fh -= real(0.5)
        f -= real(0.5)
    return x, fh, f


def rel_l2(a, ref):
    """nfft_error_l_2_complex (kernel/util/error.c:163-166): ||a - ref||_2 / ||ref||_2."""
    a = np.asarray(a, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    den = float(np.linalg.norm(ref))
    return float(np.linalg.norm(a - ref) / (den if den else 1.0))


# ------------------------------------------------------------------------------------------------------
def reference_flags():
    from nfft_b200 import plan_abi as abi
    return (abi.PRE_PHI_HUT | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT
            | abi.NFFT_SORT_NODES | abi.NFFT_OMP_BLOCKWISE_ADJOINT)


def reference_api(prec):
    """The reference's own CPU path (unmodified kernel/nfft/nfft.c, oracle/_ref fast build).  Checker / baseline."""
    from nfft_b200.plan import Api
    so = os.path.join(ROOT, "oracle", "_ref", "libnfft3_ref_fast.so" if prec == "double" else "libnfft3f_ref_fast.so")
    return Api(C.CDLL(so, mode=os.RTLD_LOCAL | os.RTLD_NOW | getattr(os, 'RTLD_DEEPBIND', 0)), prec)


class ReferencePlan:
    """One reference plan at the workload's size; pair() runs nfft_trafo + nfft_adjoint and returns the times."""

    def __init__(self, cfg, prec, x, fh, f):
        from nfft_b200.plan import Plan
        self.fh, self.f = fh, f
        self.p = Plan.init_guru(cfg["d"], cfg["N"], x.shape[0], cfg["n"], cfg["m"], reference_flags(),
                                api=reference_api(prec))
        self.p.x[:] = x

    def pair(self):
        p = self.p
        p.f_hat.view(p.api.real)[:] = self.fh.ravel()
        t0 = time.perf_counter()
        p.trafo()
        t1 = time.perf_counter()
        self.f_out = p.f.copy()
        p.f.view(p.api.real)[:] = self.f.ravel()
        t2 = time.perf_counter()
        p.adjoint()
        t3 = time.perf_counter()
        self.fh_out = p.f_hat.copy()
        return (t1 - t0), (t3 - t2)

    def close(self):
        self.p.finalize()


def run_reference(args, rank, world):
    if rank != 0:
        return
    prec, cfg = args.precision, CFGS[args.config]
    cores = int(os.environ["OMP_NUM_THREADS"]) * world   # rank 0 alone runs: all host cores
    os.environ["OMP_NUM_THREADS"] = str(cores)
    # weak scaling at N>1: the job is N*M nodes; a step here is a bounded sample of it -- one GPU's share, M nodes
    # (CPU points/s does not depend on how many of the N shares are run); strong scaling: the whole node set
    M = cfg["M"]
    x, fh, f = synth(cfg, prec, 0, M)
    rp = ReferencePlan(cfg, prec, x, fh, f)
    ts = []
    t_start = time.perf_counter()
    bounded = None
    for i in range(args.warmup + args.steps):
        a, b = rp.pair()
        ts.append(a + b)
        # safety net for slow hosts: never run for more than ~10 minutes; say so when it triggers
        if time.perf_counter() - t_start > 600 and i + 1 < args.warmup + args.steps:
            bounded = i + 1
            break
    rp.close()
    timed = ts[args.warmup:] if len(ts) > args.warmup else ts[-1:]
    t_pair = float(np.mean(timed))
    value = M / t_pair
    sample = ("the full workload (M=%d nodes on the N=%d^3 grid), %d measured pairs after %d warm-up pairs, "
              "mean %.3f s per pair (min %.3f, max %.3f); FFT stage = oracle/refbuild shim, not FFTW"
              % (M, cfg["N"][0], len(timed), min(args.warmup, len(ts) - len(timed)), t_pair, min(timed), max(timed)))
    if bounded:
        sample += "; stopped after %d pairs (10-minute cap)" % bounded
    line = {"impl": "reference", "metric": metric_name(cfg, prec), "value": value, "unit": "points/s",
            "n_gpus": args.gpus, "steps": len(timed), "warmup": args.warmup, "ms_per_step": t_pair * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if prec == "double" else "f32", "data": "synthetic", "gpu_launches": 0,
            "config": workload_config(cfg, args.gpus, args.scaling),
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "rank 0 alone runs the reference on all host cores; at N>1 it processes one GPU's share of the "
                    "weak-scaled workload (the reference is a single-node CPU code)" if args.gpus > 1 else
                    "the reference's own nfft_trafo + nfft_adjoint on all host cores"}
    print(json.dumps(line), flush=True)


def cufft_comparison(torch, dev, cfg, prec, stage):
    """cuFFT (through torch.fft.fftn) on the same oversampled grid, timed ONLY as a comparison point for the
    hand-written F stage; the product does not link cuFFT.  cuFFT transforms the full zero-padded grid out of place,
    the F stage runs band-pruned passes in place."""
    try:
        g = torch.randn(cfg["n"], dtype=torch.complex128 if prec == "double" else torch.complex64, device=dev)
        for _ in range(3):
            torch.fft.fftn(g)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            torch.fft.fftn(g)
        e1.record()
        torch.cuda.synchronize()
        return {"ours_F_ms": {"trafo": float(stage[0][1]), "adjoint": float(stage[1][1])},
                "cufft_fftn_ms": e0.elapsed_time(e1) / 10,
                "note": "cuFFT = torch.fft.fftn, full n^3 c2c out of place, comparison only; ours = band-pruned in-place passes"}
    except Exception as exc:   # comparison only: never fail the bench on it
        return {"unavailable": str(exc)[:200]}


# ------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from nfft_b200 import cabi, plan_abi as abi
    from nfft_b200.dist import ShardedPlan, slab_partition
    from nfft_b200.plan import Plan

    prec, cfg, scaling = args.precision, dict(CFGS[args.config]), args.scaling
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.nodes:
        cfg["M"] = args.nodes
    csize = 16 if prec == "double" else 8
    if scaling == "weak":
        x_h, fh_h, f_h = synth(cfg, prec, rank)
        if world > 1:   # f_hat is replicated: every rank uses rank 0's coefficients
            fh_h = synth(cfg, prec, 0, 1)[1]
        M_local = cfg["M"]
    else:
        # strong scaling: ONE node set of cfg["M"] nodes, sorted once by the reference key, equal-count slabs
        xg, fh_h, fg = synth(cfg, prec, 0)
        sel = slab_partition(cfg["N"], cfg["n"], cfg["m"], xg, rank, world, precision=prec, device=local_rank)
        x_h, f_h = np.ascontiguousarray(xg[sel]), np.ascontiguousarray(fg[sel])
        del xg, fg, sel
        M_local = x_h.shape[0]
    M_all = M_local
    if world > 1:
        t = torch.tensor([M_local], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        M_all = int(t.item())
    x_d = torch.from_numpy(x_h).to(dev)
    fh_d = torch.from_numpy(fh_h).to(dev)
    f_d = torch.from_numpy(f_h).to(dev)
    fh_out = torch.empty_like(fh_d)
    f_out = torch.empty_like(f_d)

    sp = ShardedPlan(cfg["N"], cfg["n"], cfg["m"], M_local, precision=prec, device=local_rank,
                     reduce=args.reduce)
    eng = sp.engine
    if args.psi_table:
        eng.set_option(cabi.OPT_PSI_TABLE, 1)
    if args.b_kernel:
        eng.set_option(cabi.OPT_B_KERNEL, args.b_kernel)
    if args.b_flush:
        eng.set_option(cabi.OPT_B_FLUSH, args.b_flush)
    t0 = time.perf_counter()
    sp.set_nodes_dev(x_d)
    torch.cuda.synchronize()
    t_nodes = time.perf_counter() - t0

    stage = np.zeros((2, 3))
    kern = np.zeros(2)    # main B / B^T kernel launch durations (CUDA events on the plan's stream)

    def step(record):
        sp.trafo(fh_d, f_out)
        if record:
            stage[0] += eng.stage_times()
            kern[0] += eng.b_kernel_time()
        sp.adjoint(f_d, fh_out)
        if record:
            stage[1] += eng.stage_times()
            kern[1] += eng.b_kernel_time()

    def timed_region(record):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step(record)
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # region 1 -- THE number: K steps, no per-stage timers, no host synchronisation inside
    l0 = eng.launches
    ms_step = timed_region(False)
    launches = eng.launches - l0
    # region 2 -- the same K steps with CUDA-event stage / kernel timers (one host sync per transform)
    eng.set_option(cabi.OPT_TIMING, 1)
    ms_step_timers = timed_region(True)
    eng.set_option(cabi.OPT_TIMING, 0)
    clocks = sampler.stop() if rank == 0 else None
    value = M_all / (ms_step * 1e-3)
    stage /= args.steps
    kern /= args.steps
    f_timed = f_out.cpu().numpy()
    fh_timed = fh_out.cpu().numpy()
    collective_ms = sp.collective_ms(fh_out, reps=5) if world > 1 else 0.0
    reduce_mode = sp.reduce

    # ---- rel l2 err of the timed outputs against the reference's own transforms (same inputs, full size) ----
    rel = None
    cpu_base = None
    if not args.no_check:
        rp = ReferencePlan(cfg, prec, x_h, fh_h, f_h)
        t_tr, t_ad = rp.pair()
        e_t = rel_l2(f_timed, rp.f_out.view(np.float64 if prec == "double" else np.float32))
        fh_ref = torch.from_numpy(rp.fh_out.view(np.float64 if prec == "double" else np.float32).astype(np.float64))
        if world > 1:   # the reduced f_hat is the sum of the per-rank adjoints (linearity); sum the references too
            g = fh_ref.to(dev)
            dist.all_reduce(g)
            fh_ref = g.cpu()
            et = torch.tensor([e_t], device=dev, dtype=torch.float64)
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
            e_t = float(et.item())
        e_a = rel_l2(fh_timed, fh_ref.numpy())
        rp.close()
        rel = {"trafo": e_t, "adjoint": e_a, "tol": TOL[prec], "ok": bool(e_t <= TOL[prec] and e_a <= TOL[prec]),
               "against": "the reference's nfft_trafo / nfft_adjoint (oracle/_ref/libnfft3%s_ref_fast.so = unmodified "
                          "kernel/nfft/nfft.c) on the same inputs at full size; the TIMED outputs are compared; "
                          "error = ||a-ref||_2/||ref||_2 (kernel/util/error.c:163-166); at N>1 trafo = max over ranks "
                          "of the rank's node slice, adjoint = reduced f_hat vs the sum of the per-rank references"
                          % ("" if prec == "double" else "f")}
        if world == 1:
            cpu_base = {"value": M_local / (t_tr + t_ad), "unit": "points/s", "cores": int(os.environ["OMP_NUM_THREADS"]),
                        "kind": "reference",
                        "sample": "the full workload once (M=%d, no warm-up pair): nfft_trafo %.3f s + nfft_adjoint "
                                  "%.3f s, measured in this process; FFT stage = shim, not FFTW; "
                                  "`--impl reference` times K pairs" % (M_local, t_tr, t_ad)}

    # ---- e2e: the public API on HOST buffers, H2D / D2H inside the timed region ----
    # N = 1: the reference-facing plan API (nfft_trafo / nfft_adjoint of libnfft3_b200.so).  N > 1: the process-per-GPU
    # API a multi-GPU application calls (ShardedPlan.trafo_host / adjoint_host: H2D, transform, on-device reduction
    # of f_hat, D2H); the one-process C path (NFFT_B200_DEVICES) is timed by tools/bench_group.py.
    flags = abi.PRE_PHI_HUT | abi.NFFT_SORT_NODES | abi.NFFT_OMP_BLOCKWISE_ADJOINT | abi.FFTW_INIT
    os.environ["NFFT_B200_DEVICE"] = str(local_rank)
    e2e_steps = max(1, min(args.steps, 5))

    def time_loop(fn):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        torch.cuda.synchronize()
        te = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([te], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def e2e_sharded():
        keep = [torch.from_numpy(a).pin_memory() for a in (fh_h, f_h)]
        fh_res, f_res = torch.empty_like(keep[0]).pin_memory(), torch.empty_like(keep[1]).pin_memory()

        def one():
            sp.trafo_host(keep[0], f_res, local_rank)
            sp.adjoint_host(keep[1], fh_res, local_rank)

        def one_overlapped():
            sp.pair_host(keep[0], f_res, keep[1], fh_res, local_rank)
        t_sync = time_loop(one)
        overlapped.append(time_loop(one_overlapped))
        return t_sync

    def e2e_run(lib_buffers):
        """lib_buffers: the plan API's own MALLOC_X/F_HAT/F buffers (what an unmodified C caller uses);
        otherwise caller-owned page-locked buffers (torch pinned)."""
        fl = flags | ((abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F) if lib_buffers else 0)
        extra = []
        p = Plan.init_guru(3, cfg["N"], M_local, cfg["n"], cfg["m"], fl, precision=prec)
        creal = p.api.creal
        if lib_buffers:
            # Two plans share nothing but the node values: like a solver step, trafo reads p.f_hat and writes p.f,
            # adjoint reads q.f and writes q.f_hat; the inputs are filled once (the transforms do not modify them)
            p.x[:] = x_h
            p.f_hat.view(p.api.real)[:] = fh_h.ravel()
            q = Plan.init_guru(3, cfg["N"], M_local, cfg["n"], cfg["m"], fl, precision=prec)
            q.x[:] = x_h
            q.f.view(q.api.real)[:] = f_h.ravel()
            extra.append(q)

            def one():
                p.trafo()
                q.adjoint()

            def one_overlapped():   # split-phase extension: the two plans' copies overlap each other's kernels
                p.trafo_begin()
                q.adjoint_begin()
                p.wait()
                q.wait()
        else:
            def ptr(tn):
                return C.cast(C.c_void_p(tn.data_ptr()), C.POINTER(creal))
            keep = [torch.from_numpy(a).pin_memory() for a in (x_h, fh_h, f_h)]
            fh_res, f_res = torch.empty_like(keep[1]).pin_memory(), torch.empty_like(keep[2]).pin_memory()
            keep += [fh_res, f_res]
            p.c.x = ptr(keep[0])

            def one():
                p.c.f_hat, p.c.f = ptr(keep[1]), ptr(f_res)
                p.trafo()
                p.c.f, p.c.f_hat = ptr(keep[2]), ptr(fh_res)
                p.adjoint()
        te = time_loop(one)
        if lib_buffers:
            overlapped.append(time_loop(one_overlapped))
        p.finalize()
        for q in extra:
            q.finalize()
        return te

    host_link = None
    overlapped = []
    if world > 1:
        te_lib = te_pin = e2e_sharded()
        # what the host links deliver when all ranks copy at once (the e2e step moves h2d + d2h bytes per rank through
        # them): 160 MB up and 160 MB down per rank, concurrently on all ranks, both directions in flight together
        up = torch.empty(20_000_000, dtype=torch.float64).pin_memory()
        dn = torch.empty(20_000_000, dtype=torch.float64).pin_memory()
        du, dd = torch.empty_like(up, device=dev), torch.empty_like(dn, device=dev)
        s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def both():
            with torch.cuda.stream(s_up):
                du.copy_(up, non_blocking=True)
            with torch.cuda.stream(s_dn):
                dn.copy_(dd, non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()
        t_link = time_loop(both)
        host_link = {"bytes_per_rank": 2 * up.numel() * 8, "seconds": t_link,
                     "per_gpu_gbs": 2 * up.numel() * 8 / t_link / 1e9,
                     "aggregate_gbs": world * 2 * up.numel() * 8 / t_link / 1e9,
                     "note": "all ranks copying 160 MB up + 160 MB down at the same time (page-locked memory); the e2e step "
                             "of every rank needs h2d + d2h bytes through these links on top of the device time"}
        del up, dn, du, dd, x_d, f_d, f_out
        sp.close()
    else:
        del sp, x_d, f_d, f_out
        eng.close()
        te_lib = e2e_run(True)
        te_pin = e2e_run(False)
    h2d = fh_h.nbytes + f_h.nbytes           # x is resident: changes are detected by a host-side fingerprint
    d2h = f_h.nbytes + fh_h.nbytes

    if rank == 0:
        peak, peak_src = peaks()
        bm = byte_model(cfg, csize, M_local)
        fp64_now, tf32_now, copy_now = cabi.measure_peaks(local_rank)
        t_spread = kern[1] if kern[1] > 0 else stage[1][2]
        t_interp = kern[0] if kern[0] > 0 else stage[0][2]
        spread_dom = t_spread >= t_interp
        t_dom = max(t_spread, t_interp)
        a_dom = bm["spread"] if spread_dom else bm["interp"]
        hbm_ach = a_dom / (t_dom * 1e-3) / 1e9 if t_dom > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("%s:%d" % (prec, M_local), {}).get("spread" if spread_dom else "interp")
        dmma = args.b_kernel in (0, 3) and cfg["m"] <= 6 and (prec == "double" or args.b_kernel == 3)
        tf32 = prec == "float" and args.b_kernel == 0 and cfg["m"] <= 6   # fp32 plans: 3xTF32 mma.sync kernels
        taps = (2 * cfg["m"] + 2) ** cfg["d"]
        flops = 4.0 * taps * M_local      # per tap: complex value x real weight = 2 FMA = 4 flops
        kname = (("spread_mma_kernel (B^T)" if spread_dom else "interp_mma_kernel (B)") if dmma else
                 ("tc5_spread_kernel (B^T)" if spread_dom else "tc5_interp_kernel (B)") if tf32 else
                 ("spread (B^T)" if spread_dom else "interp (B)"))
        roof = {"bound": "hbm", "kernel": kname, "achieved": hbm_ach, "peak": peak, "unit": "GB/s",
                "frac": (hbm_ach / peak) if hbm_ach else None, "traffic": traffic,
                "algorithmic_bytes_per_launch": a_dom, "launch_ms": t_dom, "peak_source": peak_src,
                "copy_gbs_measured_now": copy_now}
        if dmma or tf32:
            tf = flops / (t_dom * 1e-3) / 1e12
            floor_pair_ms = 2 * flops / (fp64_now * 1e12) * 1e3
            if dmma:
                roof["fp64_pipe"] = {
                    "achieved": tf, "peak": fp64_now, "unit": "TFLOP/s", "frac": tf / fp64_now,
                    "peak_source": "mma.sync m8n8k4 f64 rate measured in this process (nfftcu_measure_peaks, peaks.cu)",
                    "algorithmic_flops_per_launch": flops,
                    "note": "the kernel is bound by the FP64 pipe, not by HBM: useful flops only, the zero padding of the "
                            "16^3 window box (67 %% lane efficiency) is not counted.  FP64 floor for B + B^T at this "
                            "rate: %.2f ms per pair => HBM fraction of the pipeline <= %.3f however the kernels are "
                            "written; the 50 %%-of-HBM target of north_star is out of reach for fp64 on this design"
                            % (floor_pair_ms, bm["pair"] / (floor_pair_ms * 1e-3) / 1e9 / peak)}
            else:
                roof["tf32_pipe"] = {
                    "achieved": 3 * tf, "peak": tf32_now, "unit": "TFLOP/s", "frac": 3 * tf / tf32_now,
                    "peak_source": "mma.sync m16n8k8 TF32 rate measured in this process (nfftcu_measure_peaks); the kernels "
                                   "themselves issue tcgen05.mma kind::tf32 (M=128, N=16/32, K=8: 17 / 23 cycles each with A in "
                                   "tensor memory, profiles/r2l_microbench_tcgen05.txt)",
                    "note": "3xTF32 split: three tensor flops per useful flop counted (useful %.2f TFLOP/s); the tensor pipe is "
                            "~20 %% active (profiles/r2r_full_tc5_fp32.md): the kernels are bound by the hand-offs between their "
                            "warp roles, see DESIGN 4.1d" % tf}
        pipe_ach = bm["pair"] / (ms_step * 1e-3) / 1e9
        line = {
            "metric": metric_name(cfg, prec), "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64" if prec == "double" else "f32", "data": "synthetic",
            "config": workload_config(cfg, world, scaling),
            "rel_l2": rel,
            "implementation": {
                "window": (("plan-time operand images (5.2 KB per 16-node batch for B, 6.2 KB for B^T, %.2f GB together, built in "
                            "set_nodes like precompute_psi, streamed by TMA every launch)" % (M_local / 15.1 * (5248 + 6272) / 1e9))
                           if tf32 else
                           ("plan-time window images (3 KB per 8-node batch, %.2f GB, built in set_nodes like "
                            "precompute_psi, streamed by TMA every launch)" % (eng_images_gb(M_local)))
                           if dmma else ("psi table" if args.psi_table else "psi on the fly")),
                "multi_gpu": ("node-sharded x%d: trafo replicates f_hat and the grid; adjoint reduces f_hat (%s, "
                              "%.3f ms per D^T + reduction measured alone)" % (world, sp_reduce_name(reduce_mode), collective_ms)
                              if world > 1 else "single GPU"),
                "nodes_setup_s": t_nodes,
                **({"fp32_kernels": "B and B^T: tcgen05.mma kind::tf32 (3xTF32 split), grid window / accumulators in tensor memory, "
                                    "operand images by TMA (tc5.cu)"} if tf32 else {})},
            "stage_ms": {"trafo": {"D": stage[0][0], "F": stage[0][1], "B": stage[0][2]},
                         "adjoint": {"DT": stage[1][0], "F": stage[1][1], "BT": stage[1][2]},
                         "ms_per_step_with_timers": ms_step_timers},
            "kernel_ms": {"B": float(kern[0]), "BT": float(kern[1])},
            "roofline": roof,
            "roofline_pipeline": {"bound": "hbm", "achieved": pipe_ach, "peak": peak, "unit": "GB/s",
                                  "frac": pipe_ach / peak, "algorithmic_bytes_per_step": bm["pair"]},
            "peaks_measured_now": {"fp64_tensor_tflops": fp64_now, "tf32_mma_sync_tflops": tf32_now,
                                   "device_copy_gbs": copy_now},
            "e2e": {"value": M_all / te_lib, "unit": "points/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": te_lib * 1e3,
                    "buffers": ("allocated by the plan API itself (MALLOC_X|MALLOC_F_HAT|MALLOC_F through nfft_malloc, "
                                "page-locked); nfft_trafo reads f_hat and writes f, nfft_adjoint reads f and writes "
                                "f_hat, all on the host; x stays resident (host-side fingerprint, re-upload only when "
                                "it changes)" if world == 1 else
                                "process-per-GPU API (ShardedPlan.trafo_host / adjoint_host) on page-locked host "
                                "tensors: rank 0 uploads f_hat and broadcasts it over NVLink, every rank uploads its f, "
                                "transform, on-device reduction of f_hat, every rank downloads its f, rank 0 the "
                                "reduced f_hat; the byte counts are rank 0's"),
                    "caller_pinned_ms_per_step": te_pin * 1e3, "host_link": host_link,
                    "overlapped": ({"ms_per_step": overlapped[0] * 1e3, "value": M_all / overlapped[0],
                                    "how": ("nfft_b200_trafo_begin(p); nfft_b200_adjoint_begin(q); nfft_b200_wait(p); "
                                            "nfft_b200_wait(q) -- the split-phase extension of the plan API (not part of "
                                            "the reference API): same buffers, same bytes, the copies of one plan overlap "
                                            "the kernels of the other" if world == 1 else
                                            "ShardedPlan.pair_host: the adjoint's samples go up while the trafo computes, "
                                            "the trafo's result comes down while the adjoint computes; same bytes")}
                                   if overlapped else None)},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1:
            line["fft_comparison"] = cufft_comparison(torch, dev, cfg, prec, stage)
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def eng_images_gb(M):
    return M / 7.57 * 3072 / 1e9   # ~7.57 nodes per batch at cfg3 / cfg4 density


def sp_reduce_name(mode):
    return {"nccl": "ncclAllReduce behind D^T", "peer": "fused D^T + reduce kernel over NVLink peer memory"}.get(mode, mode)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="double", choices=["double", "float"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CFGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--reduce", default="auto", choices=["auto", "nccl", "peer"],
                    help="adjoint f_hat reduction at N>1: NCCL all-reduce or the fused D^T+reduce peer-memory kernel")
    ap.add_argument("--nodes", type=int, default=0, help="override M (debug)")
    ap.add_argument("--psi-table", action="store_true", help="per-node window table (PRE_PSI analogue)")
    ap.add_argument("--no-check", action="store_true", help="skip the rel-l2 check against the reference (debug)")
    ap.add_argument("--b-kernel", type=int, default=0, help="debug: NFFTCU_OPT_B_KERNEL (1 generic, 2 pencil, 3 DMMA)")
    ap.add_argument("--b-flush", type=int, default=0, help="debug: NFFTCU_OPT_B_FLUSH (1: RED.ADD instead of TMA reductions)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
