"""Instance sharding across ranks (one process per GPU).

The path has no exchange step: instance ``b`` is evaluated by rank ``b mod G`` on its own
engine.  A collective is used only to *assemble* results for a caller that wants the whole
batch on every rank (``torch.distributed.all_gather``: NCCL over NVLink on GPUs, gloo in
the CPU tests).
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np

__all__ = ["shard_indices", "ShardedBatch", "device_view"]


def shard_indices(n: int, rank: int, world: int) -> np.ndarray:
    """Instances owned by ``rank``: ``b = rank (mod world)``."""
    if not 0 <= rank < world:
        raise ValueError("rank must be in [0, world)")
    return np.arange(rank, n, world, dtype=np.int64)


class _DevicePointer:
    """Minimal ``__cuda_array_interface__`` carrier: lets torch wrap an engine buffer without a copy."""

    def __init__(self, address: int, shape: tuple, strides=None):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(address), False), "version": 3,
                                         "strides": strides}


def device_view(engine, mode: int):
    """The engine's result buffer of ``mode`` as a ``[B][n]`` float64 CUDA tensor (no copy; valid until
    the mode is evaluated again or the engine is closed)."""
    import torch

    address, count, stride = engine.out_device_pointer(mode)
    strides = None if stride == count else (8 * stride, 8)
    return torch.as_tensor(_DevicePointer(address, (engine.B, count), strides), device=torch.device("cuda", engine.device))


class ShardedBatch:
    """Evaluate a global batch of ``n`` instances, each rank computing its shard.

    ``make_evaluator(fixed_local, device)`` builds the local evaluator; by default a
    :class:`~pockit_b200.batched.BatchedSystem` on GPU ``LOCAL_RANK`` (tests inject a CPU
    stand-in).  Every callback takes the *global* arrays and returns the rank's local block;
    :meth:`gather` assembles local blocks into global order on all ranks.
    """

    def __init__(self, system, fixed_all: np.ndarray, rank: Optional[int] = None, world: Optional[int] = None,
                 make_evaluator: Optional[Callable] = None):
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
        self.n = len(fixed_all)
        self.idx = shard_indices(self.n, self.rank, self.world)
        fixed_local = np.ascontiguousarray(np.asarray(fixed_all, dtype=np.float64)[self.idx])
        if make_evaluator is None:
            from .batched import BatchedSystem

            device = int(os.environ.get("LOCAL_RANK", "0"))
            make_evaluator = lambda f, d=device: BatchedSystem(system, fixed=f, device=d)  # noqa: E731
        self.local = make_evaluator(fixed_local) if len(self.idx) else None

    def _take(self, a):
        return np.ascontiguousarray(np.asarray(a, dtype=np.float64)[self.idx])

    def objective(self, X):
        return self.local.objective(self._take(X))

    def gradient(self, X):
        return self.local.gradient(self._take(X))

    def constraints(self, X):
        return self.local.constraints(self._take(X))

    def jacobian(self, X):
        return self.local.jacobian(self._take(X))

    def hessian(self, X, fct_c, fct_o):
        sig = np.broadcast_to(np.asarray(fct_o, dtype=np.float64), (self.n,))
        return self.local.hessian(self._take(X), self._take(fct_c), self._take(sig))

    def gather(self, local_block: np.ndarray) -> np.ndarray:
        """All ranks receive the global ``[n][...]`` array assembled from the shards."""
        local_block = np.asarray(local_block, dtype=np.float64)
        local_block = local_block.reshape(len(self.idx), -1)
        if self.world == 1:
            return local_block
        import torch
        import torch.distributed as dist

        width = local_block.shape[1]
        per = (self.n + self.world - 1) // self.world  # shards differ by at most one instance: pad
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        send = torch.zeros((per, width), dtype=torch.float64, device=dev)
        send[: len(self.idx)] = torch.from_numpy(local_block).to(dev)
        recv = [torch.empty_like(send) for _ in range(self.world)]
        dist.all_gather(recv, send)
        out = np.empty((self.n, width))
        for r, t in enumerate(recv):
            ids = shard_indices(self.n, r, self.world)
            out[ids] = t[: len(ids)].cpu().numpy()
        return out

    def gather_device(self, mode: int):
        """NCCL all-gather of the local engine's latest ``mode`` values, device to device over NVLink:
        every rank receives the global ``[n][width]`` CUDA tensor in instance order.  The local shard is
        read straight from the engine's output buffer (``pk_out_device_pointer``): no host round trip."""
        import torch
        import torch.distributed as dist

        local = device_view(self.local.engine, mode)
        if self.world == 1:
            return local
        width = local.shape[1]
        per = (self.n + self.world - 1) // self.world
        if len(self.idx) == per:
            send = local
        else:  # the last ranks own one instance less: pad
            send = torch.zeros((per, width), dtype=torch.float64, device=local.device)
            send[: len(self.idx)] = local
        recv = torch.empty((self.world, per, width), dtype=torch.float64, device=local.device)
        dist.all_gather_into_tensor(recv.view(-1), send.contiguous().view(-1))
        # rank r holds instances r, r + world, ...: global instance b sits at recv[b % world, b // world]
        return recv.transpose(0, 1).reshape(per * self.world, width)[: self.n]
