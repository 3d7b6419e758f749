#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on a B200: 3-D NFFT trafo+adjoint points/s
(N=128^3, n=256^3, M=1e7 uniform random nodes per GPU, m=6, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision float]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step = one nfft_trafo + one nfft_adjoint over the rank's M nodes (+ the f_hat all-reduce when
N > 1, node-sharded as in SURVEY 8e).  `value` = nodes all ranks processed / max-over-ranks device
time, inputs resident in HBM.  `e2e` = the same step through the reference-facing plan API
(nfft_trafo / nfft_adjoint of libnfft3_b200.so) on pinned HOST buffers, copies included.
`roofline` = the dominant kernel (the B^T spreading launch) against the measured HBM peak, with
the SURVEY 8d byte model; `roofline_pipeline` = the same for the whole trafo+adjoint pair.
`--impl reference` times the reference's own CPU implementation (oracle/_ref, OpenMP, all host
threads; FFT stage is the shim, not FFTW) on bounded samples of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(d=3, N=[128, 128, 128], n=[256, 256, 256], m=6, M=10_000_000, seed=20260103)
METRIC = "3D NFFT trafo+adjoint points/s (N=128^3, M=1e7, fp64, m=6)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# FP64 tensor-core (DMMA) rate of this pool's B200s, measured with tools/microbench4.cu
# (profiles/r01w_microbench_dmma.txt: 36.9 TFLOP/s = 64 FMA/clk/SM at 1.965 GHz, same pipe as DFMA).
# MEASURED_PEAKS.json carries only HBM GB/s and bf16 TF/s; a key "fp64_tflops" there would override this.
FP64_TENSOR_TFLOPS = 36.9
# legacy tensor path used by the fp32 plans (tools/microbench5.cu, profiles/r03i_microbench_tf32.txt)
TF32_MMA_SYNC_TFLOPS = 276.7


def fp64_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        v = json.load(open(p)).get("fp64_tflops")
        if v:
            return float(v), "measured (MEASURED_PEAKS.json fp64_tflops)"
    return FP64_TENSOR_TFLOPS, ("FP64 DMMA rate measured with tools/microbench4.cu on this pool's B200 "
                                "(profiles/r01w_microbench_dmma.txt); MEASURED_PEAKS.json has no FP64 entry")


def byte_model(cfg, csize):
    """SURVEY 8d model v1.  C = complex bytes, R = C/2."""
    Cb, Rb, d = csize, csize // 2, cfg["d"]
    NN, nn, M = int(np.prod(cfg["N"])), int(np.prod(cfg["n"])), cfg["M"]
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


def synth(cfg, precision, rank):
    rng = np.random.Generator(np.random.Philox(cfg["seed"] + 1000 * rank))
    real = np.float64 if precision == "double" else np.float32
    M, d, NN = cfg["M"], cfg["d"], int(np.prod(cfg["N"]))
    x = (rng.random((M, d)) - 0.5).astype(real)
    if precision == "float":
        x = np.minimum(x, np.nextafter(np.float32(0.5), np.float32(0)))
    fh = rng.random((NN, 2)).astype(real)
    f = rng.random((M, 2)).astype(real)
    return x, fh, f


# ------------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """The reference's own CPU path (unmodified kernel/nfft/nfft.c, oracle/_ref fast build)."""
    if rank != 0:
        return
    from nfft_b200 import plan_abi as abi
    from nfft_b200.plan import Api, Plan
    prec = args.precision
    so = os.path.join(ROOT, "oracle", "_ref", "libnfft3_ref_fast.so" if prec == "double" else "libnfft3f_ref_fast.so")
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    api = Api(C.CDLL(so, mode=os.RTLD_LOCAL | os.RTLD_NOW | getattr(os, 'RTLD_DEEPBIND', 0)), prec)
    flags = (abi.PRE_PHI_HUT | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT
             | abi.NFFT_SORT_NODES | abi.NFFT_OMP_BLOCKWISE_ADJOINT)

    def pair_seconds(M, reps):
        cfg = dict(CFG, M=M)
        x, fh, f = synth(cfg, prec, 0)
        p = Plan.init_guru(3, cfg["N"], M, cfg["n"], cfg["m"], flags, api=api)
        p.x[:] = x
        p.f_hat.view(p.api.real)[:] = fh.ravel()
        ts = []
        for _ in range(reps):
            p.f_hat.view(p.api.real)[:] = fh.ravel()
            t0 = time.perf_counter()
            p.trafo()
            p.f.view(p.api.real)[:] = f.ravel()
            t1 = time.perf_counter()
            p.adjoint()
            t2 = time.perf_counter()
            ts.append((t1 - t0) + (t2 - t1))
        p.finalize()
        return ts

    # bounded sample: two node counts on the full N=128^3 grid, per-step time fitted t = a + b*M
    probe = pair_seconds(100_000, 1)[0]
    per_node = max(probe - 0.0, 1e-9) / 100_000
    budget = 8.0   # seconds per timed pair
    M2 = int(min(CFG["M"], max(200_000, budget / per_node)))
    M1 = M2 // 2
    t1s = pair_seconds(M1, args.warmup + args.steps)[args.warmup:]
    t2s = pair_seconds(M2, args.warmup + args.steps)[args.warmup:]
    t1, t2 = float(np.median(t1s)), float(np.median(t2s))
    b = max((t2 - t1) / (M2 - M1), 1e-12)
    a = max(t2 - b * M2, 0.0)
    t_full = a + b * CFG["M"]
    value = CFG["M"] / t_full
    sample = (f"N=128^3 grid, M={M1} and M={M2} nodes, median of {args.steps} pairs each: {t1:.3f}s / {t2:.3f}s; "
              f"fit t=a+b*M (a={a:.3f}s grid work incl. shim FFT, not FFTW; b={b*1e9:.1f}ns/node) "
              f"extrapolated to M=1e7: {t_full:.2f}s per pair")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if prec == "double" else "f32",
            "data": "synthetic", "gpu_launches": 0,
            "config": {"workload": "cfg3: 3-D N=128^3 n=256^3 M=1e7 m=6 sigma=2, reference flags "
                                   "PRE_PHI_HUT|NFFT_SORT_NODES|NFFT_OMP_BLOCKWISE_ADJOINT, psi on the fly"},
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "reference",
                             "sample": sample},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(precision):
    """Bounded CPU sample inside the default run: reference build when present, else the oracle port."""
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--precision", precision], capture_output=True, text=True, timeout=900)
    for ln in r.stdout.splitlines():
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    return {"value": None, "unit": "points/s", "cores": os.cpu_count(), "kind": "reference",
            "sample": "reference leg failed: " + r.stderr[-200:]}


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
    from nfft_b200.dist import ShardedPlan
    from nfft_b200.plan import Plan

    prec = args.precision
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(CFG)
    if args.nodes:
        cfg["M"] = args.nodes
    rdt = torch.float64 if prec == "double" else torch.float32
    csize = 16 if prec == "double" else 8
    x_h, fh_h, f_h = synth(cfg, prec, rank)
    if world > 1:   # f_hat is replicated: every rank uses rank 0's coefficients
        fh_h = synth(cfg, prec, 0)[1]
    x_d = torch.from_numpy(x_h).to(dev)
    fh_d = torch.from_numpy(fh_h).to(dev)
    f_d = torch.from_numpy(f_h).to(dev)
    fh_out = torch.empty_like(fh_d)
    f_out = torch.empty_like(f_d)

    sp = ShardedPlan(cfg["N"], cfg["n"], cfg["m"], cfg["M"], precision=prec, device=local_rank)
    eng = sp.engine
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_option(cabi.OPT_TIMING, 1)
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

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(True)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = eng.launches - l0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * cfg["M"] / (ms_step * 1e-3)
    stage /= args.steps
    kern /= args.steps

    # ---- e2e: reference-facing plan API on pinned host buffers (H2D/D2H inside the timed region) ----
    eng.set_option(cabi.OPT_TIMING, 0)
    flags = abi.PRE_PHI_HUT | abi.NFFT_SORT_NODES | abi.NFFT_OMP_BLOCKWISE_ADJOINT | abi.FFTW_INIT
    os.environ["NFFT_B200_DEVICE"] = str(local_rank)
    xp = torch.from_numpy(x_h).pin_memory()
    fhp = torch.from_numpy(fh_h).pin_memory()
    fp = torch.from_numpy(f_h).pin_memory()
    fh_res = torch.empty_like(fhp).pin_memory()
    f_res = torch.empty_like(fp).pin_memory()
    del sp, x_d, f_d, f_out
    eng.close()
    p = Plan.init_guru(3, cfg["N"], cfg["M"], cfg["n"], cfg["m"], flags, precision=prec)
    creal = p.api.creal

    def ptr(tn):
        return C.cast(C.c_void_p(tn.data_ptr()), C.POINTER(creal))

    p.c.x = ptr(xp)

    def e2e_step():
        p.c.f_hat, p.c.f = ptr(fhp), ptr(f_res)
        p.trafo()
        p.c.f, p.c.f_hat = ptr(fp), ptr(fh_res)
        p.adjoint()
        if world > 1:
            g = fh_res.to(dev, non_blocking=True)
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            fh_res.copy_(g)
            torch.cuda.synchronize()

    e2e_step()
    e2e_steps = max(1, min(args.steps, 5))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    te = (time.perf_counter() - t0) / e2e_steps
    p.finalize()
    tt = torch.tensor([te], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    te = float(tt.item())
    xbytes = x_h.nbytes
    h2d = 2 * xbytes + fh_h.nbytes + f_h.nbytes     # x is re-sent per transform (no psi flag, nfft.c:4889)
    d2h = f_h.nbytes + fh_h.nbytes

    if rank == 0:
        peak, peak_src = peaks()
        bm = byte_model(cfg, csize)
        # dominant kernel: the B or B^T launch itself (stage time minus memset / gather when the kernel timer ran)
        t_spread = kern[1] if kern[1] > 0 else stage[1][2]
        t_interp = kern[0] if kern[0] > 0 else stage[0][2]
        spread_dom = t_spread >= t_interp
        t_dom = max(t_spread, t_interp)
        a_dom = bm["spread"] if spread_dom else bm["interp"]
        hbm_ach = a_dom / (t_dom * 1e-3) / 1e9 if t_dom > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("%s:%d" % (prec, cfg["M"]), {}).get("spread" if spread_dom else "interp")
        dmma = args.b_kernel in (0, 3) and cfg["m"] <= 6 and (prec == "double" or args.b_kernel == 3)
        tf32 = prec == "float" and args.b_kernel == 0 and cfg["m"] <= 6   # fp32 plans: 3xTF32 mma.sync kernels
        taps = (2 * cfg["m"] + 2) ** cfg["d"]
        flops = 4.0 * taps * cfg["M"]      # per tap: complex value x real weight = 2 FMA = 4 flops
        if dmma:
            pk, pk_src = fp64_peak()
            tf = flops / (t_dom * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": ("spread_mma_kernel (B^T)" if spread_dom else "interp_mma_kernel (B)"),
                    "achieved": tf, "peak": pk, "unit": "TFLOP/s", "frac": tf / pk, "traffic": traffic,
                    "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": a_dom,
                    "launch_ms": t_dom, "peak_source": pk_src,
                    "note": "FP64 tensor cores (DMMA m8n8k4): useful flops only, zero padding of the 16^3 window "
                            "(67 % lane efficiency) not counted; traffic includes the plan-time window images "
                            "(3 KB per 8-node batch, 4.06 GB, a PRE_PSI-style table the TMA unit streams once per launch) "
                            "on top of ~1.3 GB for grid + nodes",
                    "hbm": {"achieved": hbm_ach, "peak": peak, "unit": "GB/s", "frac": hbm_ach / peak,
                            "peak_source": peak_src}}
        elif tf32:
            tf = flops / (t_dom * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": ("spread_tf32_kernel (B^T)" if spread_dom else "interp_tf32_kernel (B)"),
                    "achieved": 3 * tf, "peak": TF32_MMA_SYNC_TFLOPS, "unit": "TFLOP/s", "frac": 3 * tf / TF32_MMA_SYNC_TFLOPS,
                    "traffic": traffic, "algorithmic_flops_per_launch": 3 * flops, "algorithmic_bytes_per_launch": a_dom,
                    "launch_ms": t_dom,
                    "peak_source": "legacy mma.sync m16n8k8 TF32 rate measured with tools/microbench5.cu on this pool's "
                                   "B200 (profiles/r03i_microbench_tf32.txt); the register-resident window rules out tcgen05",
                    "note": "3xTF32 split: three tensor flops per useful flop are counted (useful: %.2f TFLOP/s); zero "
                            "padding of the 16^3 window not counted; the kernels are issue-bound, see DESIGN.md 4.1c" % tf,
                    "hbm": {"achieved": hbm_ach, "peak": peak, "unit": "GB/s", "frac": hbm_ach / peak,
                            "peak_source": peak_src}}
        else:
            roof = {"bound": "hbm", "kernel": "spread (B^T)" if spread_dom else "interp (B)", "achieved": hbm_ach,
                    "peak": peak, "unit": "GB/s", "frac": (hbm_ach / peak) if hbm_ach else None, "traffic": traffic,
                    "algorithmic_bytes_per_launch": a_dom, "launch_ms": t_dom, "peak_source": peak_src}
        pipe_ach = bm["pair"] / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if prec == "double" else "f32", "data": "synthetic",
            "config": {"workload": "cfg3: 3-D N=128^3 n=256^3 m=6 sigma=2, M=%d uniform random nodes per GPU, "
                                   "psi %s, node-sharded x%d (trafo: replicated grid; adjoint: all-reduce of f_hat)"
                                   % (cfg["M"], "table" if args.psi_table else "on the fly", world),
                       "l2": "inputs larger than L2 (grid 268 MB, x 240 MB, f 160 MB vs 126 MB L2)",
                       "nodes_setup_s": t_nodes},
            "stage_ms": {"trafo": {"D": stage[0][0], "F": stage[0][1], "B": stage[0][2]},
                         "adjoint": {"DT": stage[1][0], "F": stage[1][1], "BT": stage[1][2]}},
            "kernel_ms": {"B": float(kern[0]), "BT": float(kern[1])},
            "roofline": roof,
            "roofline_pipeline": {"bound": "hbm", "achieved": pipe_ach, "peak": peak, "unit": "GB/s",
                                  "frac": pipe_ach / peak, "algorithmic_bytes_per_step": bm["pair"]},
            "e2e": {"value": world * cfg["M"] / te, "unit": "points/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": te * 1e3},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1:
            line["fft_comparison"] = cufft_comparison(torch, dev, cfg, prec, stage)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_leg(prec)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="double", choices=["double", "float"])
    ap.add_argument("--nodes", type=int, default=0, help="override M per GPU (debug)")
    ap.add_argument("--psi-table", action="store_true", help="per-node window table (PRE_PSI analogue)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--b-kernel", type=int, default=0, help="debug: NFFTCU_OPT_B_KERNEL (1 generic, 2 pencil, 3 DMMA)")
    ap.add_argument("--b-flush", type=int, default=0, help="debug: NFFTCU_OPT_B_FLUSH (1: RED.ADD instead of TMA reductions)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
