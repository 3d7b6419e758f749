"""Multi-GPU parity (run with -m gpu).  Tests that need two devices skip on a one-GPU box; the one-device cases
exercise the same code paths (slab partition, all-to-all permutation kernels, fused D^T + reduce with its flag
barriers, the process-per-GPU driver with the REAL engine) with world size 1 or two ranks sharing cuda:0.

Checker: the CPU oracle / the reference build (oracle/_ref) on the complete node set.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import common
from common import make_case, oracle, rel_l2
from nfft_b200 import cabi, plan_abi as abi
from nfft_b200.plan import Plan

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        return cabi.lib().nfftcu_device_count()
    except Exception:
        return 0


SPEC3 = dict(d=3, N=[32, 32, 32], n=[64, 64, 64], m=6, M=60001, seed=31)
SPEC2 = dict(d=2, N=[64, 48], n=[128, 96], m=5, M=20011, seed=32)


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("spec", [SPEC3, SPEC2], ids=["3d", "2d"])
@pytest.mark.parametrize("ndev", [1, 2, 4])
def test_group_vs_oracle(spec, precision, ndev):
    """nfftcu_group_* (one process, ndev devices, host pointers): trafo, adjoint and index_x against the oracle."""
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    x, fh, f = make_case(spec, precision)
    o = oracle(precision)
    g = cabi.Group(spec["N"], spec["n"], spec["m"], spec["M"], list(range(ndev)), precision=precision,
                   flags=abi.NFFT_SORT_NODES)
    g.set_nodes(x)
    tol = 1e-12 if precision == "double" else 1e-5
    assert rel_l2(g.trafo(fh), o.trafo(spec["N"], spec["n"], spec["m"], x, fh)) <= tol
    assert rel_l2(g.adjoint(f), o.adjoint(spec["N"], spec["n"], spec["m"], x, f, True)) <= tol
    assert np.array_equal(g.index_x(), o.sort_nodes(spec["n"], spec["m"], x))
    # a second node set on the same group (re-sort, re-distribution), then the first one again
    x2 = np.ascontiguousarray(x[::-1])
    g.set_nodes(x2)
    assert rel_l2(g.trafo(fh), o.trafo(spec["N"], spec["n"], spec["m"], x2, fh)) <= tol
    g.set_nodes(x)
    assert rel_l2(g.adjoint(f), o.adjoint(spec["N"], spec["n"], spec["m"], x, f, True)) <= tol
    g.close()


@pytest.mark.skipif(not common.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("precision", ["double", "float"])
def test_plan_api_on_two_devices_vs_reference(precision, monkeypatch):
    """NFFT_B200_DEVICES=0,1: the unmodified plan API call sequence, node-sharded over two GPUs by
    libnfft3_b200.so, against the reference on the same inputs (incl. a silent node change, index_x)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("NFFT_B200_DEVICES", "0,1")
    flags = (abi.PRE_PHI_HUT | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT
             | abi.NFFT_SORT_NODES | abi.NFFT_OMP_BLOCKWISE_ADJOINT)
    spec = dict(d=3, N=[64, 64, 64], n=[128, 128, 128], m=6, M=400000, seed=33)
    x, fh, f = make_case(spec, precision)
    fh, f = fh - fh.dtype.type(0.5 + 0.5j), f - f.dtype.type(0.5 + 0.5j)   # zero mean: the fp32 reference overflows otherwise
    tol = 1e-12 if precision == "double" else 1e-5
    outs = []
    for kw in (dict(api=common.ref_api(precision)), dict(precision=precision)):
        p = Plan.init_guru(3, spec["N"], spec["M"], spec["n"], 6, flags, **kw)
        p.x[:] = x
        p.f_hat[:] = fh
        p.trafo()
        r = [p.f.copy()]
        p.f[:] = f
        p.adjoint()
        r += [p.f_hat.copy(), p.index_x.copy()]
        p.x[:] = x[::-1]          # silent node change (no psi flag): the next transform must notice
        p.f_hat[:] = fh
        p.trafo()
        r += [p.f.copy(), p.index_x.copy()]
        p.finalize()
        outs.append(r)
    ref, got = outs
    assert rel_l2(got[0], ref[0]) <= tol and rel_l2(got[1], ref[1]) <= tol
    assert np.array_equal(got[2], ref[2])
    assert rel_l2(got[3], ref[3]) <= tol
    assert np.array_equal(got[4], ref[4])


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import common
from nfft_b200.dist import ShardedPlan, slab_partition

rank, world, backend, reduce, ndev, precision = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5]), sys.argv[6]
local = rank % ndev
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
kw = dict(device_id=dev) if backend == "nccl" else dict()
dist.init_process_group(backend, init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world, **kw)
N, n, m, M = [32, 32, 32], [64, 64, 64], 6, 80001
spec = dict(d=3, N=N, n=n, m=m, M=M, seed=41)
x, fh, f = common.make_case(spec, precision)
sel = slab_partition(N, n, m, x, rank, world, precision=precision, device=local)
real = np.float64 if precision == "double" else np.float32
sp = ShardedPlan(N, n, m, len(sel), precision=precision, device=local, reduce=reduce)
# deliberately NO manual set_stream: ShardedPlan must order the engine against torch's stream itself
x_d = torch.from_numpy(np.ascontiguousarray(x[sel])).to(dev)
fh_d = torch.from_numpy(fh.view(real).reshape(-1, 2).copy()).to(dev)
f_d = torch.from_numpy(np.ascontiguousarray(f[sel]).view(real).reshape(-1, 2).copy()).to(dev)
sp.set_nodes_dev(x_d)
f_loc = torch.empty_like(f_d); fh_out = torch.empty_like(fh_d)
errs = []
for it in range(3):        # repeated calls: flag-barrier epochs, buffer reuse
    f_loc.zero_(); fh_out.zero_()
    sp.trafo(fh_d, f_loc)
    sp.adjoint(f_d, fh_out)
    torch.cuda.synchronize()
    o = common.oracle(precision)
    full_f = o.trafo(N, n, m, x, fh)
    full_fh = o.adjoint(N, n, m, x, f, True)
    cplx = np.complex128 if precision == "double" else np.complex64
    e1 = common.rel_l2(f_loc.cpu().numpy().view(cplx).ravel(), full_f[sel])
    e2 = common.rel_l2(fh_out.cpu().numpy().view(cplx).ravel(), full_fh)
    errs.append((e1, e2))
perr = sp.engine.peer_error() if sp.reduce == "peer" else 0
print("RESULT", rank, sp.reduce, max(e[0] for e in errs), max(e[1] for e in errs), perr, flush=True)
dist.barrier()
sp.close()
dist.destroy_process_group()
'''


def _run_workers(tmp_path, world, backend, reduce, ndev, precision):
    port = 29700 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=common.ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), backend, reduce, str(ndev), precision],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600) for p in procs]
    res = []
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-3000:]
        line = [ln for ln in so.splitlines() if ln.startswith("RESULT")][0].split()
        res.append((line[2], float(line[3]), float(line[4]), int(line[5])))
    return res


@pytest.mark.parametrize("precision", ["double", "float"])
def test_sharded_plan_real_engine_two_ranks_one_gpu(tmp_path, precision):
    """Two ranks sharing cuda:0 over gloo, slab-partitioned nodes, the REAL engine, no manual stream binding:
    the slab of f and the reduced f_hat against the oracle on the complete node set."""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    tol = 1e-12 if precision == "double" else 1e-5
    for mode, e1, e2, perr in _run_workers(tmp_path, 2, "gloo", "nccl", 1, precision):
        assert e1 <= tol and e2 <= tol, (e1, e2)


@pytest.mark.parametrize("reduce", ["nccl", "peer"])
@pytest.mark.parametrize("precision", ["double", "float"])
def test_sharded_plan_nccl_two_gpus(tmp_path, precision, reduce):
    """Two ranks on two GPUs over NCCL: ncclAllReduce of f_hat and the fused D^T + peer-memory reduce."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    tol = 1e-12 if precision == "double" else 1e-5
    for mode, e1, e2, perr in _run_workers(tmp_path, 2, "nccl", reduce, 2, precision):
        assert mode == reduce
        assert perr == 0
        assert e1 <= tol and e2 <= tol, (e1, e2)


@pytest.mark.parametrize("precision", ["double", "float"])
def test_fused_reduce_kernel_world1(precision):
    """The fused D^T + reduce kernel and its flag barriers with a single rank (no peer): same result as D^T."""
    import torch
    from nfft_b200.dist import ShardedPlan
    spec = dict(d=3, N=[32, 32, 32], n=[64, 64, 64], m=6, M=50000, seed=43)
    x, fh, f = make_case(spec, precision)
    real = np.float64 if precision == "double" else np.float32
    cplx = np.complex128 if precision == "double" else np.complex64
    dev = torch.device("cuda", 0)
    sp = ShardedPlan(spec["N"], spec["n"], 6, spec["M"], precision=precision, device=0, reduce="peer")
    assert sp.reduce == "peer"
    sp.set_nodes_dev(torch.from_numpy(x).to(dev))
    f_d = torch.from_numpy(f.view(real).reshape(-1, 2).copy()).to(dev)
    out = torch.empty(32 ** 3, 2, dtype=f_d.dtype, device=dev)
    for _ in range(3):
        sp.adjoint(f_d, out)
    torch.cuda.synchronize()
    want = oracle(precision).adjoint(spec["N"], spec["n"], 6, x, f, True)
    assert rel_l2(out.cpu().numpy().view(cplx).ravel(), want) <= (1e-12 if precision == "double" else 1e-5)
    assert sp.engine.peer_error() == 0
    sp.close()
