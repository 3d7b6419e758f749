"""Driver for one ncu pass over the kernels added in round 2 at meaningful sizes (tools/gpu_prof_new.sh)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nfft_b200 import cabi
from nfft_b200.dist import ShardedPlan

rng = np.random.default_rng(1)
dev = torch.device("cuda", 0)

# (1) Bluestein: 2-D grid with prime lengths 1009 x 1009 (P = 2048 per line)
N, n, M = [500, 500], [1009, 1009], 200_000
eng = cabi.Engine(N, n, 6, M)
eng.set_nodes(rng.random((M, 2)) - 0.5)
fh = (rng.random(250000) + 1j * rng.random(250000))
for _ in range(2):
    f = eng.trafo(fh)
    eng.adjoint(f)
eng.close()

# (2) batched 2-D transforms, K = 8 right-hand sides at the cfg2 / cfg5 shape
N, n, M, K = [512, 512], [1024, 1024], 512 * 512, 8
t = np.sqrt(np.arange(M) / M)
x = np.clip(np.stack([0.5 * t * np.cos(2 * np.pi * 400 * t), 0.5 * t * np.sin(2 * np.pi * 400 * t)], 1), -0.5, 0.4999999)
eng = cabi.Engine(N, n, 6, M)
eng.set_nodes(x)
fhb = (rng.random((K, 512 * 512)) + 1j * rng.random((K, 512 * 512)))
for _ in range(2):
    fb = eng.trafo_batch(fhb)
    eng.adjoint_batch(fb)
eng.close()

# (3) fused D^T + reduce (one rank) and (4) the one-process group's permutation kernels at the cfg3 shape
N, n, M = [128] * 3, [256] * 3, 10_000_000
x = rng.random((M, 3)) - 0.5
sp = ShardedPlan(N, n, 6, M, device=0, reduce="peer")
sp.set_nodes_dev(torch.from_numpy(x).to(dev))
f_d = torch.rand(M, 2, dtype=torch.float64, device=dev)
out = torch.empty(128 ** 3, 2, dtype=torch.float64, device=dev)
for _ in range(2):
    sp.adjoint(f_d, out)
torch.cuda.synchronize()
sp.close()
del f_d, out
g = cabi.Group(N, n, 6, M, [0])
g.set_nodes(x)
fh3 = rng.random(128 ** 3) + 1j * rng.random(128 ** 3)
for _ in range(2):
    f3 = g.trafo(fh3)
    g.adjoint(f3)
g.close()
print("done")
