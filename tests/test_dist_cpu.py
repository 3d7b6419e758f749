"""world_size-2 gloo test of the node-sharded driver (nfft_b200/dist.py) on CPU.

The per-rank compute object is a stand-in built on the CPU oracle (test infrastructure), so what
is exercised here is the host logic: shard ranges, replicated f_hat for trafo, and the single
all-reduce of the partial f_hat for adjoint."""
import os
import subprocess
import sys

import numpy as np
import pytest

import common
from nfft_b200.dist import shard_range

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import common
from nfft_b200.dist import ShardedPlan, shard_range

class OracleEngine:
    """stand-in for nfft_b200.cabi.Engine with the same device-pointer style interface"""
    def __init__(self, N, n, m): self.N, self.n, self.m, self.o = N, n, m, common.oracle("double")
    def set_nodes_dev(self, x): self.x = x.numpy()
    def trafo_dev(self, fh, f): f.copy_(torch.from_numpy(self.o.trafo(self.N, self.n, self.m, self.x, fh.numpy().view(np.complex128).ravel()).view(np.float64).reshape(-1, 2)))
    def adjoint_dev(self, f, fh): fh.copy_(torch.from_numpy(self.o.adjoint(self.N, self.n, self.m, self.x, f.numpy().view(np.complex128).ravel()).view(np.float64).reshape(-1, 2)))
    def close(self): pass

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
N, n, m, M = [12, 16], [32, 32], 4, 1001
rng = np.random.default_rng(5)
x = rng.random((M, 2)) - 0.5
fh = rng.random((12 * 16, 2)); f = rng.random((M, 2))
b, e = shard_range(M, rank, 2)
sp = ShardedPlan(N, n, m, e - b, engine_factory=lambda: OracleEngine(N, n, m))
sp.set_nodes_dev(torch.from_numpy(x[b:e].copy()))
f_loc = torch.empty(e - b, 2, dtype=torch.float64)
sp.trafo(torch.from_numpy(fh.copy()), f_loc)
fh_out = torch.empty(12 * 16, 2, dtype=torch.float64)
sp.adjoint(torch.from_numpy(f[b:e].copy()), fh_out)
o = common.oracle("double")
full_f = o.trafo(N, n, m, x, fh.view(np.complex128).ravel())
full_fh = o.adjoint(N, n, m, x, f.view(np.complex128).ravel())
e1 = common.rel_l2(f_loc.numpy().view(np.complex128).ravel(), full_f[b:e])
e2 = common.rel_l2(fh_out.numpy().view(np.complex128).ravel(), full_fh)
print("RESULT", rank, e1, e2, flush=True)
dist.destroy_process_group()
'''


def test_shard_range_partitions_everything():
    for M in (0, 1, 7, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            r = [shard_range(M, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == M
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_trafo_adjoint_gloo_world2(tmp_path):
    port = 29500 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=common.ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
        tag, rank, e1, e2 = [ln for ln in so.splitlines() if ln.startswith("RESULT")][0].split()
        assert float(e1) <= 1e-14      # trafo slice: bit-for-bit the same gather
        assert float(e2) <= 1e-14      # adjoint: sum of two partial f_hat == full f_hat up to reassociation
