"""Node-sharded multi-GPU NFFT (one process per GPU, torch.distributed over NCCL).

The reference is single-process; sharding follows SURVEY 8e:

* ``trafo``   every rank holds the full f_hat, runs D and F redundantly on its own grid and
              interpolates only its node shard -> its slice of f.  No collective.
* ``adjoint`` every rank spreads its node shard into its own grid, runs F and D^T, and the
              partial f_hat (2*N_total reals) are summed with ONE all-reduce.  D^T and F are
              linear, so reducing f_hat instead of the oversampled grid moves sigma^d (8x in
              3-D) fewer bytes.  The reduce is enqueued on the stream the D^T kernel ran on.

``engine_factory`` builds the per-rank compute object (default: the CUDA engine,
:class:`nfft_b200.cabi.Engine`).  The CPU tests (gloo, world_size 2) inject a stand-in so that
the sharding arithmetic and the collective are exercised without a GPU.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple


def shard_range(M_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of rank's nodes: sizes differ by at most one."""
    base, rem = divmod(int(M_total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class ShardedPlan:
    def __init__(self, N: Sequence[int], n: Sequence[int], m: int, M_local: int, *,
                 precision: str = "double", device: int = 0, group=None,
                 engine_factory: Optional[Callable] = None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if engine_factory is None:
            from .cabi import Engine
            engine_factory = lambda: Engine(N, n, m, M_local, precision=precision, device=device)  # noqa: E731
        self.engine = engine_factory()
        self.M_local = int(M_local)

    def set_nodes_dev(self, x_local):
        self.engine.set_nodes_dev(x_local)

    def trafo(self, f_hat, f_local):
        """f_local := B_local F D f_hat   (device tensors; f_hat replicated on every rank)"""
        self.engine.trafo_dev(f_hat, f_local)

    def adjoint(self, f_local, f_hat):
        """f_hat := sum over ranks of D^T F^H B_local^T f_local   (result replicated)"""
        self.engine.adjoint_dev(f_local, f_hat)
        if self.world > 1:
            self.dist.all_reduce(f_hat, op=self.dist.ReduceOp.SUM, group=self.group)

    def close(self):
        self.engine.close()
