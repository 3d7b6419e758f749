#!/usr/bin/env python
"""tools/bench_configs.py -- the BASELINE.json configs that are NOT the bench line (cfg1, cfg2, cfg4 on one
GPU, cfg5), timed for information: device-resident pair time (CUDA events), the same through the plan API on
host buffers, and for cfg5 the reference's own solver.c (CGNR, 20 iterations, 32 coils) running on the engine
(oracle/_ref/libsolver_b200.so) next to the same driver on the reference's CPU NFFT (libsolver_ref.so, timed
on a bounded number of coils).  Prints one JSON line per config.

    python tools/bench_configs.py [--configs cfg1,cfg2,cfg4,cfg5] [--coils 32] [--cpu-coils 1]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def spiral(M, N):
    A, w = 0.5, N / 64 * 50
    t = np.sqrt(np.arange(M) / M)
    x = np.stack([A * t * np.cos(2 * np.pi * w * t), A * t * np.sin(2 * np.pi * w * t)], 1)
    return np.ascontiguousarray(np.clip(x, -0.5, np.nextafter(0.5, 0.0)))


def pair_times(N, n, m, x, precision="double", steps=20, warmup=3):
    import torch
    from nfft_b200 import cabi
    M = x.shape[0]
    NN = int(np.prod(N))
    rng = np.random.default_rng(1)
    cplx = np.complex128 if precision == "double" else np.complex64
    fh = (rng.random(NN) + 1j * rng.random(NN)).astype(cplx)
    f = (rng.random(M) + 1j * rng.random(M)).astype(cplx)
    dev = torch.device("cuda", 0)
    eng = cabi.Engine(N, n, m, M, precision=precision)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    t0 = time.perf_counter()
    eng.set_nodes(x)
    torch.cuda.synchronize()
    t_nodes = time.perf_counter() - t0
    fh_d, f_d = torch.from_numpy(fh).to(dev), torch.from_numpy(f).to(dev)
    f_o, fh_o = torch.empty_like(f_d), torch.empty_like(fh_d)
    eng.set_option(cabi.OPT_TIMING, 1)
    stage = np.zeros((2, 3))
    for _ in range(warmup):
        eng.trafo_dev(fh_d, f_o)
        eng.adjoint_dev(f_d, fh_o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.trafo_dev(fh_d, f_o)
        stage[0] += eng.stage_times()
        eng.adjoint_dev(f_d, fh_o)
        stage[1] += eng.stage_times()
    e1.record()
    torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1) / steps
    eng.set_option(cabi.OPT_TIMING, 0)
    for _ in range(2):
        eng.trafo(fh)
        eng.adjoint(f)
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.trafo(fh)
        eng.adjoint(f)
    ms_host = (time.perf_counter() - t0) / steps * 1e3
    eng.close()
    stage /= steps
    return dict(ms_pair_device=ms_dev, ms_pair_host_api=ms_host, points_per_s=M / (ms_dev * 1e-3),
                nodes_setup_ms=t_nodes * 1e3,
                stage_ms=dict(D=stage[0][0], F_trafo=stage[0][1], B=stage[0][2], BT=stage[1][2], F_adj=stage[1][1],
                              DT=stage[1][0]))


def cfg5(coils, cpu_coils, iters=20):
    """reconstruct_data_2d.c:38-120 per coil: CGNR | PRECOMPUTE_DAMP, disc mask, 20 fixed iterations."""
    from nfft_b200 import plan_abi as abi
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    N, n, m, M = [512, 512], [1024, 1024], 6, 512 * 512
    x = spiral(M, 512)
    NN = 512 * 512
    k = np.stack(np.meshgrid(*[np.arange(-v // 2, v // 2) / v for v in N], indexing="ij"), -1)
    w_hat = np.ascontiguousarray((np.sqrt((k ** 2).sum(-1)) <= 0.5).astype(np.float64).ravel())
    w = np.ones(M)
    CGNR, PRE_D = 1 << 2, 1 << 6
    flags = abi.PRE_PHI_HUT | abi.PRE_PSI | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT
    rng = np.random.default_rng(5)
    ys = [np.ascontiguousarray(rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)) for _ in range(max(coils, cpu_coils))]
    ia = lambda a: (C.c_int * len(a))(*a)          # noqa: E731
    p = lambda a: a.ctypes.data_as(C.c_void_p)     # noqa: E731
    out = {}
    res = {}
    for name, so, nc in (("b200_device_solver", "libsolver_dev_b200.so", coils), ("b200", "libsolver_b200.so", coils),
                         ("ref_cpu", "libsolver_ref.so", cpu_coils)):
        path = os.path.join(ref_dir, so)
        if not os.path.exists(path) or nc <= 0:
            continue
        L = C.CDLL(path, mode=os.RTLD_LOCAL)
        fn = L.solver_driver_run
        fn.restype = C.c_int
        f_hat = np.zeros(NN, dtype=np.complex128)
        dots = np.zeros(iters)
        if name != "ref_cpu":   # warm-up (context creation, module load)
            fn(C.c_int(2), ia(N), C.c_int(M), ia(n), C.c_int(m), C.c_uint(flags), C.c_uint(CGNR | PRE_D),
               p(x), p(ys[0]), p(w), p(w_hat), C.c_int(1), p(f_hat), p(dots), C.c_double(0.0), None)
        t0 = time.perf_counter()
        for c in range(nc):
            rc = fn(C.c_int(2), ia(N), C.c_int(M), ia(n), C.c_int(m), C.c_uint(flags), C.c_uint(CGNR | PRE_D),
                    p(x), p(ys[c]), p(w), p(w_hat), C.c_int(iters), p(f_hat), p(dots), C.c_double(0.0), None)
            assert rc == 0
            if c == 0:
                res[name] = f_hat.copy()
        dt = time.perf_counter() - t0
        out[name] = dict(coils=nc, seconds=dt, seconds_per_coil=dt / nc)
    for name in ("b200", "b200_device_solver"):
        if name in res and "ref_cpu" in res:
            out[name]["rel_l2_coil0_vs_reference"] = float(np.linalg.norm(res[name] - res["ref_cpu"]) / np.linalg.norm(res["ref_cpu"]))
        if name in out:
            out[name]["transforms_per_s"] = coils * iters * 2 / out[name]["seconds"]
        if name in out and "ref_cpu" in out:
            out[name]["speedup_per_coil"] = out["ref_cpu"]["seconds_per_coil"] / out[name]["seconds_per_coil"]
    out["cpu_threads"] = os.cpu_count()
    return out


def cfg5_multi_gpu(coils, gpus, iters=20, batched=False):
    """cfg5 distributed plan-per-GPU (SURVEY 8e "Batched / multi-coil": replicas only, no collective): coil c runs on
    GPU c mod gpus, one worker process per GPU (env NFFT_B200_DEVICE), each driving the device-resident solver
    through the unmodified plan-per-coil call sequence of reconstruct_data_2d.c.  Wall clock of the slowest worker,
    after every worker has warmed up (context creation, module load) and all have passed a start barrier (a file)."""
    import subprocess
    import tempfile
    go = tempfile.NamedTemporaryFile(delete=False).name
    os.unlink(go)
    procs = []
    for r in range(gpus):
        mine = [c for c in range(coils) if c % gpus == r]
        env = dict(os.environ, NFFT_B200_DEVICE=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__),
                                       "--cfg5-batch-worker" if batched else "--cfg5-worker",
                                       ",".join(map(str, mine)), "--go-file", go, "--iters", str(iters)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    # workers print READY after warm-up; release them together
    for p_ in procs:
        line = p_.stdout.readline()
        assert line.startswith("READY"), (line, p_.stderr.read()[-2000:])
    t0 = time.perf_counter()
    open(go, "w").close()
    outs = [p_.communicate(timeout=1200) for p_ in procs]
    wall = time.perf_counter() - t0
    os.unlink(go)
    per = []
    for p_, (so, se) in zip(procs, outs):
        assert p_.returncode == 0, se[-2000:]
        per.append(json.loads([ln for ln in so.splitlines() if ln.startswith("{")][-1]))
    # the workers start together (go file) and time their own coil loops; the wall clock of this parent also contains
    # process teardown (CUDA context destruction), so the job time is the slowest worker's
    slowest = max(w["seconds"] for w in per)
    return dict(gpus=gpus, coils=coils, iters=iters, seconds=slowest, seconds_per_coil=slowest / coils,
                coils_per_s=coils / slowest, transforms_per_s=coils * iters * 2 / slowest, parent_wall_seconds=wall,
                workers=per)


def cfg5_batch_worker(coil_ids, go_file, iters):
    """The same reconstruction with the coils of this GPU as ONE batched solve: one plan, one node set, K = len(coil_ids)
    right-hand sides in lock-step (nfftcu_solver_create_batch on nfftcu_*_batch_dev).  Timed: plan creation, node set-up,
    uploads, before_loop + `iters` steps, download of the iterates."""
    from nfft_b200 import cabi
    N, n, m, M = [512, 512], [1024, 1024], 6, 512 * 512
    x = spiral(M, 512)
    NN = 512 * 512
    k = np.stack(np.meshgrid(*[np.arange(-v // 2, v // 2) / v for v in N], indexing="ij"), -1)
    w_hat = np.ascontiguousarray((np.sqrt((k ** 2).sum(-1)) <= 0.5).astype(np.float64).ravel())
    rng = np.random.default_rng(5)
    K = len(coil_ids)
    ys = np.ascontiguousarray(np.stack([rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5) for _ in coil_ids]))
    dev = int(os.environ.get("NFFT_B200_DEVICE", "0"))

    def solve():
        eng = cabi.Engine(N, n, m, M, device=dev)
        eng.set_option(cabi.OPT_PSI_TABLE, 1)
        eng.set_nodes(x)
        s = cabi.BatchSolver(eng, cabi.CGNR | cabi.PRECOMPUTE_DAMP, K)
        s.upload(cabi.SOLVER_Y, ys)
        s.upload(cabi.SOLVER_W_HAT, w_hat)
        s.upload(cabi.SOLVER_F_HAT_ITER, np.zeros((K, NN), dtype=np.complex128))
        s.before_loop()
        for _ in range(iters):
            sc = s.step()
        out = s.download(cabi.SOLVER_F_HAT_ITER)
        s.close()
        eng.close()
        return out, sc
    solve()            # warm-up: context creation, module load
    print("READY", flush=True)
    while not os.path.exists(go_file):
        time.sleep(0.0005)
    t0 = time.perf_counter()
    out, sc = solve()
    dt = time.perf_counter() - t0
    print(json.dumps(dict(device=str(dev), coils=K, seconds=dt, last_dot_r=float(sc[-1, 2]))), flush=True)


def cfg5_worker(coil_ids, go_file, iters):
    from nfft_b200 import plan_abi as abi
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsolver_dev_b200.so"), mode=os.RTLD_LOCAL)
    fn = L.solver_driver_run
    fn.restype = C.c_int
    N, n, m, M = [512, 512], [1024, 1024], 6, 512 * 512
    x = spiral(M, 512)
    NN = 512 * 512
    k = np.stack(np.meshgrid(*[np.arange(-v // 2, v // 2) / v for v in N], indexing="ij"), -1)
    w_hat = np.ascontiguousarray((np.sqrt((k ** 2).sum(-1)) <= 0.5).astype(np.float64).ravel())
    w = np.ones(M)
    CGNR, PRE_D = 1 << 2, 1 << 6
    flags = abi.PRE_PHI_HUT | abi.PRE_PSI | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT
    ia = lambda a: (C.c_int * len(a))(*a)          # noqa: E731
    p = lambda a: a.ctypes.data_as(C.c_void_p)     # noqa: E731
    rng = np.random.default_rng(5)
    ys = {c: np.ascontiguousarray(rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)) for c in coil_ids}
    f_hat, dots = np.zeros(NN, dtype=np.complex128), np.zeros(iters)
    y0 = next(iter(ys.values())) if ys else np.zeros(M, dtype=np.complex128)
    fn(C.c_int(2), ia(N), C.c_int(M), ia(n), C.c_int(m), C.c_uint(flags), C.c_uint(CGNR | PRE_D), p(x), p(y0), p(w),
       p(w_hat), C.c_int(1), p(f_hat), p(dots), C.c_double(0.0), None)
    print("READY", flush=True)
    while not os.path.exists(go_file):
        time.sleep(0.0005)
    t0 = time.perf_counter()
    for c in coil_ids:
        rc = fn(C.c_int(2), ia(N), C.c_int(M), ia(n), C.c_int(m), C.c_uint(flags), C.c_uint(CGNR | PRE_D), p(x),
                p(ys[c]), p(w), p(w_hat), C.c_int(iters), p(f_hat), p(dots), C.c_double(0.0), None)
        assert rc == 0
    dt = time.perf_counter() - t0
    print(json.dumps(dict(device=os.environ.get("NFFT_B200_DEVICE", "0"), coils=len(coil_ids), seconds=dt,
                          last_dot_r=float(dots[-1]))), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1, help="cfg5: distribute the coils plan-per-GPU over this many GPUs")
    ap.add_argument("--cfg5-worker", default=None)
    ap.add_argument("--cfg5-batch-worker", default=None)
    ap.add_argument("--go-file", default=None)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--configs", default="cfg1,cfg2,cfg4,cfg5")
    ap.add_argument("--coils", type=int, default=32)
    ap.add_argument("--cpu-coils", type=int, default=1)
    ap.add_argument("--cfg4-nodes", type=int, default=100_000_000)
    args = ap.parse_args()
    if args.cfg5_batch_worker is not None:
        cfg5_batch_worker([int(v) for v in args.cfg5_batch_worker.split(",") if v], args.go_file, args.iters)
        return
    if args.cfg5_worker is not None:
        cfg5_worker([int(v) for v in args.cfg5_worker.split(",") if v], args.go_file, args.iters)
        return
    want = args.configs.split(",")
    rng = np.random.default_rng(20260101)
    if "cfg1" in want:
        x = rng.random((10000, 1)) - 0.5
        print(json.dumps(dict(config="cfg1: 1-D N=1024 n=2048 M=10000 m=6 fp64", **pair_times([1024], [2048], 6, x))), flush=True)
    if "cfg2" in want:
        x = spiral(512 * 512, 512)
        for prec in ("double", "float"):
            xx = x if prec == "double" else np.minimum(x.astype(np.float32), np.nextafter(np.float32(0.5), np.float32(0)))
            print(json.dumps(dict(config="cfg2: 2-D N=512^2 n=1024^2 M=512^2 spiral m=6 " + prec,
                                  **pair_times([512, 512], [1024, 1024], 6, xx, prec))), flush=True)
    if "cfg4" in want:
        M = args.cfg4_nodes
        x = rng.random((M, 3)) - 0.5
        print(json.dumps(dict(config="cfg4 on ONE GPU: 3-D N=256^3 n=512^3 M=%d m=6 fp64" % M,
                              **pair_times([256] * 3, [512] * 3, 6, x, steps=3, warmup=1))), flush=True)
    if "cfg3r" in want:
        # 3-D radial trajectory (applications/mri/mri3d style): M = 10^7 nodes on lines through the origin, density ~ 1/r^2
        M = 10_000_000
        lines = 40_000
        v = rng.normal(size=(lines, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        r = (np.arange(M // lines) / (M // lines) - 0.5)
        x = np.ascontiguousarray((v[:, None, :] * r[None, :, None]).reshape(-1, 3))
        x = np.clip(x, -0.5, np.nextafter(0.5, 0.0))
        print(json.dumps(dict(config="cfg3 grid with a 3-D RADIAL trajectory (clustered at the centre): N=128^3 n=256^3 M=%d m=6 fp64" % M,
                              **pair_times([128] * 3, [256] * 3, 6, x, steps=5, warmup=2))), flush=True)
    if "cfg5" in want:
        print(json.dumps(dict(config="cfg5: 2-D CGNR 20 iterations x %d coils, 512^2 spiral, reference solver.c on the engine"
                                     % args.coils, **cfg5(args.coils, args.cpu_coils))), flush=True)
    if "cfg5batch" in want:
        print(json.dumps(dict(config="cfg5 batched: 2-D CGNR 20 iterations x %d coils over %d GPUs, 512^2 spiral, the coils of "
                                     "a GPU as ONE batched solve (nfftcu_solver_create_batch), one worker process per GPU"
                                     % (args.coils, args.gpus), **cfg5_multi_gpu(args.coils, args.gpus, batched=True))), flush=True)
    if "cfg5mg" in want:
        print(json.dumps(dict(config="cfg5 plan-per-GPU: 2-D CGNR 20 iterations x %d coils over %d GPUs, 512^2 spiral, "
                                     "device-resident solver_*_complex, one worker process per GPU, no collective"
                                     % (args.coils, args.gpus), **cfg5_multi_gpu(args.coils, args.gpus))), flush=True)


if __name__ == "__main__":
    main()
