"""Node-sharded multi-GPU NFFT, one process per GPU (torch.distributed for the plumbing).

The reference is single-process; sharding follows SURVEY 8e:

* ``trafo``   every rank holds the full f_hat, runs D and F redundantly on its own grid and
              interpolates only its node shard -> its slice of f.  No collective.
* ``adjoint`` every rank spreads its node shard into its own grid and runs F; the partial results
              are summed across ranks.  D^T and F are linear, so the sum is taken over f_hat
              (2*N_total reals, sigma^d = 8x fewer bytes than the grid in 3-D), in one of two ways:
              ``reduce="peer"``  the reduction is PART OF the D^T kernel: rank r reads the band corners of
                                 all ranks' grids through NVLink peer pointers (CUDA IPC), scales by c once
                                 and stores slice r of f_hat into every rank's result buffer
                                 (``nfftcu_adjoint_dev_peer``, csrc/peer.cu);
              ``reduce="nccl"``  local D^T, then ``ncclAllReduce`` of the partial f_hat.
              ``reduce="auto"``  peer when the peers can be attached, else nccl.

Strong scaling (``slab_partition``): the node set is sorted once by the reference key and cut into
equal-count slabs of the sorted order, so every rank works on a compact slab of the grid at the full
node density.  The single-process flavour of the same design (one host thread, P devices, host pointers)
is ``nfftcu_group_*`` (csrc/shard.cu, ``nfft_b200.cabi.Group``), which ``libnfft3_b200.so`` uses when
``NFFT_B200_DEVICES`` lists several devices.

``engine_factory`` builds the per-rank compute object (default: the CUDA engine,
:class:`nfft_b200.cabi.Engine`).  The CPU tests (gloo, world_size 2) inject a stand-in so that
the sharding arithmetic and the collective are exercised without a GPU.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np


def shard_range(M_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of rank's nodes: sizes differ by at most one."""
    base, rem = divmod(int(M_total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def slab_partition(N: Sequence[int], n: Sequence[int], m: int, x_global: np.ndarray, rank: int, world: int, *,
                   precision: str = "double", device: int = 0, order=None) -> np.ndarray:
    """Original indices of the nodes of rank's slab: positions ``shard_range(M, rank, world)`` of the node list
    sorted by the reference key (nfft.c:75-109, stable).  The sort runs on ``device`` (sort.cu, bit-exact against
    the reference permutation); ``order`` may supply a precomputed permutation instead (CPU tests)."""
    M = int(x_global.shape[0])
    b, e = shard_range(M, rank, world)
    if order is not None:
        return np.asarray(order[b:e], dtype=np.int64)
    from . import cabi
    from .plan_abi import NFFT_SORT_NODES
    eng = cabi.Engine(N, n, m, M, precision=precision, flags=NFFT_SORT_NODES, device=device)
    try:
        eng.set_option(cabi.OPT_B_KERNEL, 1)     # reference order only: no tile binning, no window images
        eng.set_nodes(x_global)
        buf = cabi.DeviceBuffer(4 * max(e - b, 1), device)
        eng.sorted_slab(b, e, None, buf)
        eng.sync()
        sel = buf.download(np.uint32, e - b).astype(np.int64)
        buf.free()
    finally:
        eng.close()
    return sel


class ShardedPlan:
    def __init__(self, N: Sequence[int], n: Sequence[int], m: int, M_local: int, *,
                 precision: str = "double", device: int = 0, group=None,
                 engine_factory: Optional[Callable] = None, reduce: str = "auto"):
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
        self._bound_stream = None
        self.reduce = "nccl"
        if reduce not in ("auto", "nccl", "peer"):
            raise ValueError("reduce must be auto, nccl or peer")
        if reduce in ("auto", "peer") and self.world > 1 and hasattr(self.engine, "peer_export"):
            self.reduce = "peer" if self._attach_peers(device, strict=(reduce == "peer")) else "nccl"
        elif reduce == "peer" and hasattr(self.engine, "peer_export"):
            self._attach_peers(device, strict=True)      # world 1: the fused kernel with one rank
            self.reduce = "peer"

    def _attach_peers(self, device: int, strict: bool) -> bool:
        """exchange the CUDA IPC handles of grid / exchange buffer / flags and map the peers (csrc/peer.cu)"""
        import torch
        from .cabi import PEER_HANDLE_BYTES
        ok, err = 1, None
        try:
            mine = self.engine.peer_export()
        except Exception as exc:      # noqa: BLE001 -- any rank failing disables the peer path on all ranks
            ok, err, mine = 0, exc, bytes(PEER_HANDLE_BYTES)
        if self.world > 1:
            dev = torch.device("cuda", device)
            t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
            allh = torch.empty(self.world * PEER_HANDLE_BYTES, dtype=torch.uint8, device=dev)
            self.dist.all_gather_into_tensor(allh, t, group=self.group)
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
            ok = int(flag.item())
            blob = bytes(allh.cpu().numpy().tobytes())
        else:
            blob = mine
        if ok:
            try:
                self.engine.peer_attach(self.rank, self.world, blob)
            except Exception as exc:  # noqa: BLE001
                ok, err = 0, exc
            if self.world > 1:
                flag = torch.tensor([ok], device=torch.device("cuda", device), dtype=torch.int32)
                self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
                ok = int(flag.item())
        if not ok and strict:
            raise RuntimeError(f"peer-memory reduce unavailable: {err}")
        return bool(ok)

    def _bind_stream(self, like):
        """Run the engine on torch's current stream of the tensor's device (nfftcu_set_stream): the engine's own
        stream is cudaStreamNonBlocking, i.e. unordered against torch's streams and the NCCL call that follows."""
        if not getattr(like, "is_cuda", False) or not hasattr(self.engine, "set_stream"):
            return
        import torch
        s = torch.cuda.current_stream(like.device).cuda_stream
        if s != self._bound_stream:
            self.engine.set_stream(s)
            self._bound_stream = s

    def set_nodes_dev(self, x_local):
        self._bind_stream(x_local)
        self.engine.set_nodes_dev(x_local)

    def trafo(self, f_hat, f_local):
        """f_local := B_local F D f_hat   (device tensors; f_hat replicated on every rank)"""
        self._bind_stream(f_hat)
        self.engine.trafo_dev(f_hat, f_local)

    def adjoint(self, f_local, f_hat):
        """f_hat := sum over ranks of D^T F^H B_local^T f_local   (result replicated)"""
        self._bind_stream(f_hat)
        if self.reduce == "peer":
            self.engine.adjoint_dev_peer(f_local, f_hat)
            return
        self.engine.adjoint_dev(f_local, f_hat)
        if self.world > 1:
            self.dist.all_reduce(f_hat, op=self.dist.ReduceOp.SUM, group=self.group)

    # ---- host-buffer face (what a process-per-GPU application calls): page-locked torch tensors in and out ----
    def _staging(self, like_fhat, like_f, device):
        import torch
        if getattr(self, "_st", None) is None:
            dev = torch.device("cuda", device)
            self._st = (torch.empty(like_fhat.shape, dtype=like_fhat.dtype, device=dev),
                        torch.empty(like_f.shape, dtype=like_f.dtype, device=dev))
        return self._st

    def trafo_host(self, f_hat_host, f_local_host, device: int = 0, root: int = 0):
        """f_local_host := B_local F D f_hat.  f_hat is the same on every rank, so it crosses a host link ONCE: rank
        ``root`` uploads its ``f_hat_host`` and broadcasts it over NVLink (``root=None``: every rank uploads its own
        copy); then the transform and the D2H of the rank's slice of f, in stream order; returns after completion."""
        import torch
        fh_d, f_d = self._staging(f_hat_host, f_local_host, device)
        if root is None or self.world == 1:
            fh_d.copy_(f_hat_host, non_blocking=True)
        else:
            if self.rank == root:
                fh_d.copy_(f_hat_host, non_blocking=True)
            self.dist.broadcast(fh_d, src=root, group=self.group)
        self.trafo(fh_d, f_d)
        f_local_host.copy_(f_d, non_blocking=True)
        torch.cuda.current_stream(fh_d.device).synchronize()

    def adjoint_host(self, f_local_host, f_hat_host, device: int = 0, root: int = 0):
        """f_hat_host (on rank ``root``; ``root=None``: on every rank) := sum over ranks of the adjoints.  The reduction
        runs on the device (fused into D^T or NCCL); every rank moves its samples up once, and only the root moves the
        reduced f_hat down."""
        import torch
        fh_d, f_d = self._staging(f_hat_host, f_local_host, device)
        f_d.copy_(f_local_host, non_blocking=True)
        self.adjoint(f_d, fh_d)
        if root is None or self.rank == root or self.world == 1:
            f_hat_host.copy_(fh_d, non_blocking=True)
        torch.cuda.current_stream(fh_d.device).synchronize()

    def pair_host(self, f_hat_host, f_out_host, f_in_host, f_hat_out_host, device: int = 0, root: int = 0):
        """One trafo (f_hat_host -> f_out_host) and one adjoint (f_in_host -> f_hat_out_host) on independent data with
        the copies overlapped with the kernels: the adjoint's samples go up on a side stream while the trafo computes,
        the trafo's result comes down while the adjoint computes.  Same bytes as trafo_host + adjoint_host."""
        import torch
        fh_d, f_d = self._staging(f_hat_host, f_out_host, device)
        if getattr(self, "_st2", None) is None:
            self._st2 = (torch.empty_like(f_d), torch.empty_like(fh_d), torch.cuda.Stream(fh_d.device),
                         torch.cuda.Stream(fh_d.device))
        f2_d, fh2_d, s_up, s_dn = self._st2
        main = torch.cuda.current_stream(fh_d.device)
        s_up.wait_stream(main)                       # earlier work on the staging buffers
        with torch.cuda.stream(s_up):
            f2_d.copy_(f_in_host, non_blocking=True)     # overlaps the trafo's kernels
        if self.world == 1 or root is None:
            fh_d.copy_(f_hat_host, non_blocking=True)
        else:
            if self.rank == root:
                fh_d.copy_(f_hat_host, non_blocking=True)
            self.dist.broadcast(fh_d, src=root, group=self.group)
        self.trafo(fh_d, f_d)
        s_dn.wait_stream(main)                       # trafo done: its result comes down on its own stream ...
        with torch.cuda.stream(s_dn):
            f_out_host.copy_(f_d, non_blocking=True)     # ... while the adjoint computes
        main.wait_stream(s_up)                       # the adjoint needs the uploaded samples only
        self.adjoint(f2_d, fh2_d)
        if root is None or self.rank == root or self.world == 1:
            f_hat_out_host.copy_(fh2_d, non_blocking=True)
        main.synchronize()
        s_dn.synchronize()

    def collective_ms(self, f_hat, reps: int = 5) -> float:
        """Device time of D^T + cross-rank reduction alone (on whatever the grids hold), max over ranks."""
        import torch
        if not getattr(f_hat, "is_cuda", False):
            return 0.0
        self._bind_stream(f_hat)
        scratch = torch.empty_like(f_hat)

        def once():
            if self.reduce == "peer":
                self.engine.peer_reduce_only(scratch)
            else:
                self.engine.stage_DT(scratch)
                if self.world > 1:
                    self.dist.all_reduce(scratch, op=self.dist.ReduceOp.SUM, group=self.group)
        once()
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            once()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=f_hat.device, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def close(self):
        self.engine.close()
