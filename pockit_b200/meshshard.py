"""One fine mesh evaluated by several GPUs ("one very fine mesh").

The callback values of a 40 k - 100 k node transcription are 200+ MB per evaluation set and
the consumer (Ipopt / SciPy) lives on the host, so the path is bound by the device-to-host
copy, not by the kernels.  Sharding therefore splits what is expensive -- the slot-streaming
expansion and the copy -- and replicates what is cheap:

* every rank (one process per GPU) receives the whole ``x`` (and multipliers; 2.9 + 1.9 MB for
  robot_arm LGR 2000x20) and runs the per-node programs, the quadrature sums and the system
  program for all nodes: a few microseconds, and it makes the integrals bit-identical on every
  rank without an all-reduce -- the data path has **no collective**;
* rank ``g`` runs the block expansion for one contiguous range of the (list, interval) tiles in
  output order -- a few whole lists plus at most two partial ones -- and its share of the long
  table / constant runs (``plan.ModePlan._shard``), and copies exactly those slot runs (a handful of
  long ones) over *its own* PCIe link into a host buffer shared by all ranks
  (a ``/dev/shm`` mapping, page-locked in every process), at the offsets of the reference pattern;
* rank 0 is the caller: it owns the small callbacks (objective, gradient, constraints), publishes
  ``x`` in the shared mapping, evaluates its own share and waits for the others' completion flags.

Values and patterns are exactly those of the unsharded engine (the same kernels write the same
slots); only who writes which slot changes.  ``torch.distributed`` is used for the rendezvous
(sharing the mapping's name) -- NCCL on the GPU box, gloo in the CPU tests.

Usage (all ranks)::

    ms = MeshShardedSystem(system)          # collective
    if ms.rank == 0:
        ... ms.jacobian(x) / ms.hessian(x, lam, sigma) / ms.evaluate(x, lam, sigma) ...
        ms.close()                          # releases the workers
    else:
        ms.serve()                          # returns when rank 0 closes
"""
from __future__ import annotations

import mmap
import os
import time
from typing import Callable, Optional

import numpy as np

from . import plan as P

__all__ = ["MeshShardedSystem"]

# int64 control words: [0] sequence (x and the command are published), [1] command, [2] sequence for which the
# multipliers are published too, [8+g] done sequence of rank g, [24+g] error flag,
# [40+g] / [56+g] time.monotonic_ns() at which rank g saw the point / finished its share (diagnostics)
_CTRL = 80
_CMD_JAC, _CMD_HESS, _CMD_EXIT = 1, 2, 4
_PAGE = 4096


def _round(n: int, to: int = _PAGE) -> int:
    return (n + to - 1) // to * to


class _SharedBuffers:
    """The mapping all ranks see: control words, inputs, and the two large value arrays."""

    def __init__(self, path: str, create: bool, L: int, m: int, nnz_j: int, nnz_h: int):
        self.path = path
        off = _round(8 * _CTRL)
        self.off = {}
        for name, n in (("x", L), ("lam", max(m, 1)), ("sig", 1), ("jac", max(nnz_j, 1)), ("hess", max(nnz_h, 1))):
            self.off[name] = (off, n)
            off = _round(off + 8 * n)
        self.size = off
        flags = os.O_RDWR | (os.O_CREAT | os.O_EXCL if create else 0)
        fd = os.open(path, flags, 0o600)
        try:
            if create:
                os.ftruncate(fd, self.size)
            self.mm = mmap.mmap(fd, self.size, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        finally:
            os.close(fd)
        self.ctrl = np.frombuffer(self.mm, dtype=np.int64, count=_CTRL, offset=0)
        self.arr = {k: np.frombuffer(self.mm, dtype=np.float64, count=n, offset=o) for k, (o, n) in self.off.items()}
        self.address = np.frombuffer(self.mm, dtype=np.uint8).ctypes.data

    def unlink(self):
        try:
            os.unlink(self.path)
        except FileNotFoundError:
            pass


class MeshShardedSystem:
    def __init__(self, system, rank: Optional[int] = None, world: Optional[int] = None, device: Optional[int] = None,
                 make_engine: Optional[Callable] = None, page_lock: bool = True, shm_dir: str = "/dev/shm",
                 _startup_delay: float = 0.0):
        self.system = system
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
        if self.world > 16:
            raise ValueError("at most 16 ranks per sharded mesh")
        lo = self.lowering = system.lowering
        self.L, self.m = lo.r_s, lo.m
        self.nnz_jac, self.nnz_hess = lo.nnz_jac, lo.nnz_hess_o + lo.nnz_hess_c
        shard = (self.rank, self.world)
        self.link_rates = None
        if make_engine is None:
            from .engine import Engine  # raises without the CUDA library / a device

            if device is None:
                device = int(os.environ.get("LOCAL_RANK", str(self.rank)))
            make_engine = lambda lowering, sh: Engine(lowering, fastmath=system._fastmath, device=device, shard=sh)  # noqa: E731
            if self.world > 1 and os.environ.get("POCKIT_B200_MESH_WEIGHTS", "1") != "0":
                # The path is bound by the device-to-host copies, and the ranks' links are not equally fast on
                # every box (8 GPUs behind two uplinks: one half of the ranks finished 0.7 ms after the other):
                # measure what every rank gets while ALL ranks copy, and size the shares accordingly.  Rank 0
                # also carries the small callbacks and starts a little later: a slightly smaller share.
                self.link_rates = self._measure_link_rates(device)
                w = np.array(self.link_rates, dtype=np.float64)
                w[0] *= 0.9
                shard = (self.rank, self.world, w / w.sum())
        self.engine = make_engine(lo, shard)
        # rendezvous: rank 0 creates the mapping, everybody attaches, then the name is removed
        path = [f"{shm_dir}/pockit_b200_mesh_{os.getpid()}_{time.time_ns() & 0xFFFFFF:x}" if self.rank == 0 else None]
        if self.rank == 0:
            self.buf = _SharedBuffers(path[0], True, self.L, self.m, self.nnz_jac, self.nnz_hess)
        if self.world > 1:
            import torch.distributed as dist

            if not dist.is_initialized():
                raise RuntimeError("MeshShardedSystem with world > 1 needs torch.distributed to be initialised")
            dist.broadcast_object_list(path, src=0)
            if self.rank != 0:
                self.buf = _SharedBuffers(path[0], False, self.L, self.m, self.nnz_jac, self.nnz_hess)
            dist.barrier()
        if self.rank == 0:
            self.buf.unlink()
        self._locked = False
        if page_lock and hasattr(self.engine, "lib"):
            self.engine._check(self.engine.lib.pk_host_register(self.buf.address, self.buf.size))
            self._locked = True
        if _startup_delay:  # test hook: a rank that finishes its start-up late (slow page-locking)
            time.sleep(_startup_delay)
        # the mapping starts zeroed; do NOT read the live counter here: rank 0 may already have published
        # its first point while this rank was still page-locking the mapping (seen at 8 ranks: a worker
        # that read 1 waited for the counter to change forever)
        self._seq = 0
        self._helper = None
        if self.rank == 0:
            from concurrent.futures import ThreadPoolExecutor

            self._helper = ThreadPoolExecutor(max_workers=1, thread_name_prefix="pockit_b200_publish")
        self.last_timeline = {}
        self.pinned_outputs = False
        self._closed = False

    def _measure_link_rates(self, device: int, mbytes: int = 32, reps: int = 6) -> list:
        """GB/s of device-to-host copies into page-locked memory per rank, all ranks copying at the same time
        (collective: every rank calls it)."""
        import torch
        import torch.distributed as dist

        n = (mbytes << 20) // 8
        dev = torch.zeros(n, dtype=torch.float64, device=torch.device("cuda", device))
        host = torch.empty(n, dtype=torch.float64, pin_memory=True)
        host.copy_(dev)
        torch.cuda.synchronize(device)
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.device(device):
            a.record()
            for _ in range(reps):
                host.copy_(dev, non_blocking=True)
            b.record()
            torch.cuda.synchronize(device)
        rate = reps * 8 * n / (a.elapsed_time(b) * 1e-3) / 1e9
        rates = [None] * self.world
        dist.all_gather_object(rates, float(rate))
        return [float(r) for r in rates]

    # ------------------------------------------------------------------ structures (unchanged)
    def jacobianstructure(self):
        return self.system.jacobianstructure()

    def hessianstructure(self):
        return self.system.hessianstructure()

    # ------------------------------------------------------------------ worker side
    @staticmethod
    def _wait(cond, what: str, timeout: float = 120.0):
        t0 = time.perf_counter()
        spins = 0
        while not cond():
            spins += 1
            if spins > 20000:  # a few milliseconds of polling (the gap between two sets of a solver), then stop
                # hogging the core: with one process per GPU the pollers otherwise compete with rank 0 for host cores
                time.sleep(0.00002 if spins < 200000 else 0.0005)
                if time.perf_counter() - t0 > timeout:
                    raise TimeoutError(f"mesh shard: timed out waiting for {what}")

    def _enqueue(self, modes, outs, x_sent: bool, with_multipliers: bool):
        """Start ``modes`` on the engine without waiting (``pk_eval_set_async``): ``x`` travels with the first
        call of a point, later calls run at the resident copy."""
        a = self.buf.arr
        lam = a["lam"][: self.m] if with_multipliers else None
        return self.engine.evaluate(None if x_sent else a["x"], lam, a["sig"] if with_multipliers else None,
                                    modes=list(modes), outs=list(outs), wait=False)

    def _run_share(self, cmd: int, extra_modes=(), lam_ready=None):
        """This rank's share of one evaluation point, in two stages: everything that needs only ``x``
        (the small callbacks on rank 0, the Jacobian share) starts as soon as ``x`` is published -- its
        device-to-host copy is already running when the multipliers arrive (``lam_ready()`` returns, or
        publishes them on rank 0) and the Hessian share follows.  One synchronisation at the end."""
        a = self.buf.arr
        res, x_sent = {}, False
        pending_publish = None
        if self.rank == 0 and lam_ready is not None:
            # the caller publishes the multipliers on a helper thread WHILE it enqueues its own first stage (the
            # C call releases the GIL): the workers are waiting for them, and its own GPU should not wait either
            pending_publish = self._helper.submit(lam_ready)
            lam_ready = pending_publish.result
        first = list(extra_modes) + ([P.JAC] if cmd & _CMD_JAC else [])
        if first:
            outs = [None] * len(extra_modes) + ([a["jac"][: self.nnz_jac]] if cmd & _CMD_JAC else [])
            t0 = time.perf_counter()
            res.update(self._enqueue(first, outs, x_sent, False))
            self._enqueue_ms = (1e3 * (time.perf_counter() - t0), getattr(self.engine, "last_call_ms", None))
            x_sent = True
        if cmd & _CMD_HESS:
            if lam_ready is not None:
                lam_ready()
            res.update(self._enqueue([P.HESS], [a["hess"][: self.nnz_hess]], x_sent, True))
        if res:
            self.engine.sync()
            if P.OBJ in res and np.ndim(res[P.OBJ]) and np.size(res[P.OBJ]) == 1:
                res[P.OBJ] = np.float64(np.asarray(res[P.OBJ]).reshape(-1)[0])  # read only now: the call above was asynchronous
        return res

    def serve(self, idle_timeout: float = 600.0):
        """Ranks > 0: evaluate this rank's share whenever rank 0 publishes a new point.  Gives up when
        nothing arrives for ``idle_timeout`` seconds (a caller that died must not leave workers -- and
        the GPUs they hold -- waiting forever)."""
        if self.rank == 0:
            raise RuntimeError("rank 0 is the caller, not a worker")
        ctrl = self.buf.ctrl
        while True:
            self._wait(lambda: int(ctrl[0]) != self._seq, "the next evaluation point", timeout=idle_timeout)
            self._seq = int(ctrl[0])
            ctrl[40 + self.rank] = time.monotonic_ns()
            cmd = int(ctrl[1])
            if cmd & _CMD_EXIT:
                break
            try:
                self._run_share(cmd, lam_ready=lambda: self._wait(lambda: int(ctrl[2]) == self._seq, "the multipliers", timeout=60.0))
            except Exception:
                ctrl[24 + self.rank] = 1
                ctrl[8 + self.rank] = self._seq
                raise
            ctrl[56 + self.rank] = time.monotonic_ns()
            ctrl[8 + self.rank] = self._seq
        self._release()

    # ------------------------------------------------------------------ caller side (rank 0)
    def _evaluate(self, cmd: int, x, fct_c=None, fct_o=None, extra_modes=()):
        if self.rank != 0:
            raise RuntimeError("only rank 0 calls the callbacks; the other ranks serve()")
        a, ctrl = self.buf.arr, self.buf.ctrl
        x = np.asarray(x, dtype=np.float64)
        if x.size != self.L:
            raise ValueError(f"x must have {self.L} entries")
        lam = None
        if cmd & _CMD_HESS:
            lam = np.asarray(fct_c, dtype=np.float64)
            if lam.size != self.m:
                raise ValueError(f"fct_c must have {self.m} entries")
        t = {"start": time.monotonic_ns()}
        a["x"][:] = x.reshape(-1)
        ctrl[1] = cmd
        self._seq += 1
        ctrl[0] = self._seq  # publish x (x86: stores are not reordered with earlier stores)
        t["x_published"] = time.monotonic_ns()

        def publish_multipliers():  # while this runs, every rank is already uploading x / expanding its Jacobian share
            t["first_stage_enqueued"] = time.monotonic_ns()
            a["lam"][: self.m] = lam.reshape(-1)
            a["sig"][0] = float(fct_o)
            ctrl[2] = self._seq
            t["multipliers_published"] = time.monotonic_ns()

        res = self._run_share(cmd, extra_modes, lam_ready=publish_multipliers if cmd & _CMD_HESS else None)
        t["own_share_done"] = time.monotonic_ns()
        for g in range(1, self.world):
            self._wait(lambda g=g: int(ctrl[8 + g]) == self._seq, f"rank {g}")
            if ctrl[24 + g]:
                raise RuntimeError(f"mesh shard: rank {g} failed")
        t["all_done"] = time.monotonic_ns()
        # where the time of the last point went, in ms since its start (host clock, shared by all ranks)
        self.last_timeline = {k: (v - t["start"]) / 1e6 for k, v in t.items() if k != "start"}
        self.last_timeline["first_enqueue_ms (python, C call)"] = getattr(self, "_enqueue_ms", None)
        self.last_timeline["workers_saw_point"] = [(int(ctrl[40 + g]) - t["start"]) / 1e6 for g in range(1, self.world)]
        self.last_timeline["workers_done"] = [(int(ctrl[56 + g]) - t["start"]) / 1e6 for g in range(1, self.world)]
        return res

    def _view(self, name: str, n: int):
        v = self.buf.arr[name][:n]
        return v if self.pinned_outputs else v.copy()

    def objective(self, x):
        return self.engine.objective(x)

    def gradient(self, x):
        return self.engine.gradient(x)

    def constraints(self, x):
        return self.engine.constraints(x)

    def jacobian(self, x):
        self._evaluate(_CMD_JAC, x)
        return self._view("jac", self.nnz_jac)

    def hessian(self, x, fct_c, fct_o):
        self._evaluate(_CMD_HESS, x, fct_c, fct_o)
        return self._view("hess", self.nnz_hess)

    def evaluate(self, x, fct_c=None, fct_o=1.0):
        """The whole set at one ``x`` (see ``System.evaluate``), Jacobian / Hessian values sharded."""
        cmd = _CMD_JAC | (_CMD_HESS if fct_c is not None else 0)
        res = self._evaluate(cmd, x, fct_c, fct_o, extra_modes=(P.OBJ, P.GRAD, P.CONS))
        out = {"objective": res[P.OBJ], "gradient": res[P.GRAD], "constraints": res[P.CONS],
               "jacobian": self._view("jac", self.nnz_jac)}
        if fct_c is not None:
            out["hessian"] = self._view("hess", self.nnz_hess)
        return out

    # ------------------------------------------------------------------
    def _release(self):
        if self._closed:
            return
        self._closed = True
        if self._helper is not None:
            self._helper.shutdown(wait=True)
        if self._locked:
            try:
                self.engine.lib.pk_host_unregister(self.buf.address)
            except Exception:
                pass
        if hasattr(self.engine, "close"):
            self.engine.close()

    def close(self):
        """Rank 0: release the workers and the engine."""
        if self.rank == 0 and not self._closed:
            self.buf.ctrl[1] = _CMD_EXIT
            self._seq += 1
            self.buf.ctrl[0] = self._seq
        self._release()
