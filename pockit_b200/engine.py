"""ctypes binding of the C-ABI engine (``include/pockit_b200.h``) and the
host-side glue that feeds it a :class:`~pockit_b200.plan.DevicePlan`.

There is no CPU fallback: if ``libpockit_b200.so`` or a CUDA device is missing
the constructor raises, and so does every callback of the owning System.
"""
from __future__ import annotations

import ctypes as C
import os
import time
import weakref
from pathlib import Path
from typing import Optional

import numpy as np

from . import plan as P

__all__ = ["Engine", "load_library", "library_path", "PinnedArray", "PinnedPool", "cubin_cache_stats"]

N_STAGES = 6
N_MODES = 6  # five callbacks + the fused set pipeline (plan.SET)
N_CALLBACKS = 5


class _Job(C.Structure):
    _fields_ = [("type", C.c_int32), ("flags", C.c_int32), ("i", C.c_int64 * 16), ("f", C.c_double * 2)]


class _Dims(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("batch", C.c_int32), ("L", C.c_int64), ("m", C.c_int64),
        ("nnz_jac", C.c_int64), ("nnz_hess", C.c_int64), ("n_scalar", C.c_int64), ("n_table", C.c_int64),
        ("n_fixed", C.c_int64),
    ]


class _NodeProgram(C.Structure):
    _fields_ = [("kernel", C.c_char_p), ("n_nodes", C.c_int64), ("tm_offset", C.c_int64), ("wm_offset", C.c_int64)]


class _ModeDesc(C.Structure):
    _fields_ = [
        ("cuda_source", C.c_char_p), ("nvrtc_options", C.POINTER(C.c_char_p)), ("n_nvrtc_options", C.c_int32),
        ("n_node_programs", C.c_int32), ("node_programs", C.POINTER(_NodeProgram)), ("system_kernel", C.c_char_p),
        ("table_symbol", C.c_char_p), ("table", C.POINTER(C.c_int64)), ("n_table_entries", C.c_int64),
        ("n_scalar", C.c_int64), ("n_out", C.c_int64), ("jobs", C.c_void_p * N_STAGES), ("n_jobs", C.c_int64 * N_STAGES),
        ("grad_offset", C.c_int64), ("grad_count", C.c_int64),
        ("sub_offset", C.c_int64 * N_CALLBACKS), ("sub_count", C.c_int64 * N_CALLBACKS),
    ]


class _AugPhase(C.Structure):
    _fields_ = [
        ("prep_kernel", C.c_char_p), ("node_kernel", C.c_char_p), ("x_offset", C.c_int64),
        ("L", C.c_int64), ("L_xu", C.c_int64), ("L_x_all", C.c_int64),
        ("n_x", C.c_int64), ("n_u", C.c_int64), ("Lm_aug", C.c_int64), ("rows", C.c_int64),
        ("tm_aug", C.c_void_p),
        ("V_ptr", C.c_void_p), ("V_idx", C.c_void_p), ("V_val", C.c_void_p),
        ("T_ptr", C.c_void_p), ("T_idx", C.c_void_p), ("T_val", C.c_void_p),
        ("I_ptr", C.c_void_p), ("I_idx", C.c_void_p), ("I_val", C.c_void_p),
    ]


assert C.sizeof(_Job) == P.JOB_DTYPE.itemsize

_LIB = None


def library_path() -> Path:
    return Path(os.environ.get("POCKIT_B200_LIB", Path(__file__).resolve().parent / "libpockit_b200.so"))


def load_library():
    """Load ``libpockit_b200.so`` (built by ``__graft_entry__.build()``)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not path.exists():
        raise RuntimeError(f"{path} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(str(path))
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_void_p
    sig = {
        "pk_abi_version": ([], C.c_int),
        "pk_last_error": ([], C.c_char_p),
        "pk_device_count": ([C.POINTER(C.c_int)], C.c_int),
        "pk_engine_create": ([C.POINTER(_Dims), C.c_int, C.POINTER(vp)], C.c_int),
        "pk_engine_destroy": ([vp], C.c_int),
        "pk_engine_set_pools": ([vp, vp, C.c_int64, vp, C.c_int64], C.c_int),
        "pk_engine_set_fixed": ([vp, vp], C.c_int),
        "pk_engine_load_mode": ([vp, C.c_int, C.POINTER(_ModeDesc)], C.c_int),
        "pk_engine_get_cubin": ([vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)], C.c_int),
        "pk_eval_objective": ([vp, vp, vp], C.c_int),
        "pk_eval_gradient": ([vp, vp, vp], C.c_int),
        "pk_eval_constraints": ([vp, vp, vp], C.c_int),
        "pk_eval_jacobian": ([vp, vp, vp], C.c_int),
        "pk_eval_hessian": ([vp, vp, vp, vp, vp], C.c_int),
        "pk_eval_set": ([vp, vp, vp, vp, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)], C.c_int),
        "pk_eval_set_async": ([vp, vp, vp, vp, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)], C.c_int),
        "pk_engine_set_output_runs": ([vp, C.c_int, vp, C.c_int64], C.c_int),
        "pk_engine_set_compaction": ([vp, C.c_int, C.c_int64, vp, vp], C.c_int),
        "pk_out_size": ([vp, C.c_int, C.POINTER(C.c_int64)], C.c_int),
        "pk_engine_load_error_estimate": ([vp, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(_AugPhase), C.c_int], C.c_int),
        "pk_eval_error_data": ([vp, vp, vp, vp], C.c_int),
        "pk_upload_x": ([vp, vp], C.c_int),
        "pk_upload_multipliers": ([vp, vp, vp], C.c_int),
        "pk_run": ([vp, C.c_int], C.c_int),
        "pk_run_set": ([vp, C.POINTER(C.c_int), C.c_int], C.c_int),
        "pk_sync": ([vp], C.c_int),
        "pk_download": ([vp, C.c_int, vp], C.c_int),
        "pk_download_range": ([vp, C.c_int, C.c_int64, C.c_int64, vp], C.c_int),
        "pk_out_device_pointer": ([vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64)], C.c_int),
        "pk_time": ([vp, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)], C.c_int),
        "pk_time_stage": ([vp, C.c_int, C.c_uint, C.c_int, C.c_int, C.POINTER(C.c_float)], C.c_int),
        "pk_time_stage_alternating": ([vp, C.POINTER(C.c_int), C.c_int, C.c_uint, C.c_int, C.POINTER(C.c_float)], C.c_int),
        "pk_time_steps": ([vp, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)], C.c_int),
        "pk_kernel_launches": ([vp, C.POINTER(C.c_int64)], C.c_int),
        "pk_expand_variant": ([vp, C.c_int, C.POINTER(C.c_int)], C.c_int),
        "pk_x_uploads": ([vp, C.POINTER(C.c_int64)], C.c_int),
        "pk_cubin_cache_stats": ([C.POINTER(C.c_int64), C.POINTER(C.c_int64)], C.c_int),
        "pk_timeline": ([vp, C.POINTER(C.c_int), C.c_int, vp, C.c_int, C.POINTER(C.c_int)], C.c_int),
        "pk_flush_l2": ([vp], C.c_int),
        "pk_alloc_host": ([C.c_size_t], vp),
        "pk_free_host": ([vp], None),
        "pk_host_register": ([vp, C.c_size_t], C.c_int),
        "pk_host_unregister": ([vp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, res
    if lib.pk_abi_version() != 2:
        raise RuntimeError("libpockit_b200.so: ABI version mismatch")
    _LIB = lib
    return lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class PinnedArray:
    """float64 NumPy view over page-locked host memory (``pk_alloc_host``)."""

    def __init__(self, n: int):
        self._lib = load_library()
        self._p = self._lib.pk_alloc_host(max(8, 8 * int(n)))
        if not self._p:
            raise MemoryError("pk_alloc_host failed")
        self.array = np.ctypeslib.as_array((C.c_double * int(n)).from_address(self._p)) if n else np.zeros(0)

    def __del__(self):
        if getattr(self, "_p", None):
            self._lib.pk_free_host(self._p)
            self._p = None


class PinnedPool:
    """Page-locked result buffers handed out as *leases*.

    The reference returns a fresh array per call (ownership): the caller may keep it for as
    long as it likes.  Device-to-host copies run at full PCIe rate only into page-locked memory, and
    page-locking per call is far too slow, so results are written into buffers of this pool and the
    caller receives a NumPy array that *owns a lease* on its buffer: the buffer goes back to the pool
    when the last array (or view of it) referring to it is garbage-collected, never earlier.  A solver
    that drops the previous Jacobian before asking for the next one keeps hitting the same buffer; one
    that holds on to results simply gets further buffers (beyond ``max_bytes`` outstanding: ordinary
    pageable arrays).  Nothing is ever overwritten behind the caller's back."""

    def __init__(self, lib, max_bytes: int = 4 << 30, keep_free: int = 2):
        self._lib, self.max_bytes, self.keep_free = lib, int(max_bytes), int(keep_free)
        self._free: dict = {}  # doubles -> [address]
        self.bytes = 0         # page-locked bytes owned (leased + free)
        self.allocations = 0
        self._closed = False

    def take(self, n: int) -> np.ndarray:
        n = int(n)
        stack = self._free.get(n)
        if stack:
            addr = stack.pop()
        else:
            if self.bytes + 8 * n > self.max_bytes:
                return np.empty(n, dtype=np.float64)
            addr = self._lib.pk_alloc_host(8 * n)
            if not addr:
                return np.empty(n, dtype=np.float64)
            self.bytes += 8 * n
            self.allocations += 1
        owner = (C.c_double * n).from_address(addr)
        # every array derived from the one returned here keeps `owner` alive (NumPy's base chain ends
        # at the buffer provider), so this fires exactly when the caller has let go of the result
        weakref.finalize(owner, self._give_back, n, addr).atexit = False
        return np.frombuffer(owner, dtype=np.float64)

    def _give_back(self, n: int, addr: int):
        stack = self._free.setdefault(n, [])
        if self._closed or len(stack) >= self.keep_free:
            self._lib.pk_free_host(addr)
            self.bytes -= 8 * n
        else:
            stack.append(addr)

    def close(self):
        """Release the free buffers; leased ones are freed when their arrays die."""
        self._closed = True
        for n, stack in self._free.items():
            for addr in stack:
                self._lib.pk_free_host(addr)
                self.bytes -= 8 * n
        self._free = {}


def cubin_cache_stats() -> tuple:
    """(hits, misses) of the library's process-wide NVRTC result cache."""
    lib = load_library()
    h, m = C.c_int64(), C.c_int64()
    lib.pk_cubin_cache_stats(C.byref(h), C.byref(m))
    return h.value, m.value


class Engine:
    """One CUDA engine for one lowered system (``batch`` independent instances)."""

    def __init__(self, lowering, batch: int = 1, fastmath: bool = False, device: Optional[int] = None,
                 fixed: Optional[np.ndarray] = None, shard: Optional[tuple] = None):
        self.lib = load_library()
        self.lowering = lowering
        self.B = int(batch)
        self.shard = shard
        # The set pipeline (plan.SET): callbacks evaluated as ONE pipeline -- one per-node program with a
        # shared CSE, one reduction, one system program -- when a whole set is run (pk_run_set).
        #   POCKIT_B200_SET=small (default)  the three latency-bound callbacks (objective, gradient,
        #       constraints): one chain of small kernels instead of three beside the two expansions;
        #   POCKIT_B200_SET=1  all five (measured on B200 in round 1: SLOWER than separate pipelines,
        #       robot_arm 76 vs 67 us per set -- its single chain queues behind the block expansions);
        #   POCKIT_B200_SET=0  none.
        # Needs a whole plan (no mesh shard / fused walk).
        choice = os.environ.get("POCKIT_B200_SET", "small")
        subs = {"0": (), "1": P.SET_ORDER}.get(choice, P.SMALL_SET)
        self.has_set = shard is None and os.environ.get("POCKIT_B200_FUSED", "0") != "1" and bool(subs)
        self.set_subs = tuple(subs) if self.has_set else ()
        self.plan = P.DevicePlan(lowering, self.B, fastmath, shard=shard,
                                 fused=shard is None and os.environ.get("POCKIT_B200_FUSED", "0") == "1",
                                 node_groups=int(os.environ.get("POCKIT_B200_NODE_GROUPS", "1")),
                                 set_subs=subs or P.SET_ORDER)
        self.fin = {}
        self._mode_ids = list(range(N_CALLBACKS)) + ([P.SET] if self.has_set else [])
        for m in self._mode_ids:
            self.plan.mode(m)
        for m in self._mode_ids:
            self.fin[m] = self.plan.finalize(m)
        lo = lowering
        self.n_out = {m: self.fin[m]["n_out"] for m in self._mode_ids}  # slots on the device
        self.n_host = dict(self.n_out)  # values the host receives (smaller with a de-duplicated pattern)
        self.compacted = set()  # modes whose duplicates are summed on the device
        dims = _Dims(
            2, self.B, lo.r_s, lo.m, lo.nnz_jac, lo.nnz_hess_o + lo.nnz_hess_c,
            max(1, max(f["n_scalar"] for f in self.fin.values())),
            max(1, max(f["n_table"] for f in self.fin.values())), self.plan.n_fixed,
        )
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("POCKIT_B200_USE_LOCAL_RANK") else 0
        self.device = device
        h = C.c_void_p()
        self._h = None
        self._check(self.lib.pk_engine_create(C.byref(dims), device, C.byref(h)))
        self._h = h
        self._dpool, self._ipool = self.plan.pools.arrays()
        self._check(self.lib.pk_engine_set_pools(h, _ptr(self._dpool), self._dpool.size, _ptr(self._ipool), self._ipool.size))
        self.set_fixed(fixed)
        self._loaded = set()
        self._pinned = {}
        self.reuse_outputs = False
        # results of 64 KB and more come from the lease pool (see PinnedPool); POCKIT_B200_PINNED_POOL=0
        # returns plain pageable NumPy arrays instead (slow device-to-host copies)
        self.pool = PinnedPool(self.lib) if os.environ.get("POCKIT_B200_PINNED_POOL", "1") != "0" else None
        self._opts = ["--fmad=true"] if fastmath else []
        self._one = np.ones(self.B)
        self._zero_lam = np.zeros(self.B * max(1, lo.m))

    # ------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError("pockit_b200 engine: " + self.lib.pk_last_error().decode(errors="replace"))

    def close(self):
        if self._h is not None:
            self.lib.pk_engine_destroy(self._h)
            self._h = None
            if getattr(self, "pool", None) is not None:
                self.pool.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_fixed(self, fixed: Optional[np.ndarray]):
        """FIXED boundary values per instance, ``[batch][n_fixed]`` in the order
        per phase ``[x(t0) (n_x), x(tf) (n_x), t0, tf]`` (entries of non-FIXED slots are ignored)."""
        if fixed is None:
            fixed = np.tile(self.plan.fixed_default, (self.B, 1))
        fixed = np.ascontiguousarray(np.asarray(fixed, dtype=np.float64).reshape(self.B, self.plan.n_fixed))
        self._fixed = fixed
        if fixed.size:
            self._check(self.lib.pk_engine_set_fixed(self._h, _ptr(fixed)))

    def load(self, mode: int):
        if mode in self._loaded:
            return
        f = self.fin[mode]
        lo = self.lowering
        progs = (_NodeProgram * max(1, len(f["kernels"])))()
        for pi, k in enumerate(f["kernels"]):
            # threads per instance: nodes x expression groups of the phase's program
            progs[pi] = _NodeProgram(k.encode(), lo.phases[pi].L_m * f["node_threads"][pi], self.plan.tm_off[pi], self.plan.wm_off[pi])
        opts = (C.c_char_p * max(1, len(self._opts)))(*[o.encode() for o in self._opts])
        table = np.ascontiguousarray(f["table"], dtype=np.int64)
        d = _ModeDesc()
        d.cuda_source = f["source"].encode()
        d.nvrtc_options = opts
        d.n_nvrtc_options = len(self._opts)
        d.n_node_programs = len(f["kernels"])
        d.node_programs = progs
        d.system_kernel = f["sys_kernel"].encode() if f["sys_kernel"] else None
        d.table_symbol = f["table_symbol"].encode()
        d.table = table.ctypes.data_as(C.POINTER(C.c_int64))
        d.n_table_entries = len(table)
        d.n_scalar = f["n_scalar"]
        d.n_out = f["n_out"]
        d.grad_offset, d.grad_count = (int(v) for v in f["grad_range"])
        if mode == P.SET:
            for k in range(N_CALLBACKS):
                d.sub_offset[k], d.sub_count[k] = (int(v) for v in f["sub_range"].get(k, (0, 0)))
        keep = []
        for s in range(N_STAGES):
            arr = np.ascontiguousarray(f["jobs"][s])
            keep.append(arr)
            d.jobs[s] = arr.ctypes.data if len(arr) else None
            d.n_jobs[s] = len(arr)
        self._check(self.lib.pk_engine_load_mode(self._h, mode, C.byref(d)))
        self._loaded.add(mode)
        runs = f.get("runs")
        if runs is not None:  # mesh shard: copy back only what this rank computes
            if len(runs) == 0:
                raise RuntimeError(f"mode {P.MODES[mode]} is evaluated by rank 0 of the sharded mesh, not by rank {self.shard[0]}")
            if not (len(runs) == 1 and runs[0][0] == 0 and runs[0][1] == f["n_out"]):
                self._check(self.lib.pk_engine_set_output_runs(self._h, mode, _ptr(np.ascontiguousarray(runs)), len(runs)))

    def _load_for_set(self, modes):
        """Load what a set evaluation of ``modes`` needs: the modes themselves and, when all five
        callbacks are asked for on a whole (unsharded, uncompacted) plan, the fused set pipeline,
        which the engine then runs instead of five separate pipelines."""
        for m in modes:
            self.load(m)
        if self.has_set and set(self.set_subs) <= set(modes) and not (self.compacted & set(self.set_subs)):
            self.load(P.SET)

    # ------------------------------------------------------------------ host-to-host callbacks
    def _x(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.size != self.B * self.lowering.r_s:
            raise ValueError(f"x must have {self.B * self.lowering.r_s} entries")
        return x

    def _out(self, mode: int, out: Optional[np.ndarray], small_pinned: bool = False) -> np.ndarray:
        """Destination of a mode's values.  ``small_pinned``: page-locked even for tiny results -- a copy into
        pageable memory blocks the host until the mode has finished, which an asynchronous call must not."""
        n = self.B * self.n_host[mode]
        if out is None:
            if self.reuse_outputs:
                # page-locked, engine-owned result buffers: D2H at full PCIe rate, no per-call
                # allocation; the returned array is valid until the next call of the same callback
                if mode not in self._pinned:
                    self._pinned[mode] = PinnedArray(n)
                return self._pinned[mode].array
            if self.pool is not None and (n >= 8192 or small_pinned):
                return self.pool.take(n)
            return np.empty(n, dtype=np.float64)
        if out.size != n or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array of the callback's size")
        return out

    def _shape(self, mode, out):
        return out if self.B == 1 else out.reshape(self.B, self.n_host[mode])

    # ------------------------------------------------------------------ output shaping
    def set_compaction(self, mode: int, seg_ptr: Optional[np.ndarray], perm: Optional[np.ndarray]):
        """De-duplicated pattern for ``mode``: unique entry ``u`` is the sum of the slots
        ``perm[seg_ptr[u]:seg_ptr[u+1]]``; ``None`` restores the reference pattern."""
        self.load(mode)
        if seg_ptr is None:
            self._check(self.lib.pk_engine_set_compaction(self._h, mode, 0, None, None))
            self.n_host[mode] = self.n_out[mode]
            self.compacted.discard(mode)
        else:
            seg_ptr = np.ascontiguousarray(seg_ptr, dtype=np.int64)
            perm = np.ascontiguousarray(perm, dtype=np.int64)
            if len(perm) != self.n_out[mode]:
                raise ValueError("perm must list every slot of the mode once")
            self._check(self.lib.pk_engine_set_compaction(self._h, mode, len(seg_ptr) - 1, _ptr(seg_ptr), _ptr(perm)))
            self.n_host[mode] = len(seg_ptr) - 1
            self.compacted.add(mode)
        self._pinned.pop(mode, None)

    def set_output_runs(self, mode: int, runs: Optional[np.ndarray]):
        """Mesh shard: ``runs`` = ``[(offset, count), ...]`` of the output this engine computes and
        copies back (``None``: everything)."""
        self.load(mode)
        if runs is None or len(runs) == 0:
            self._check(self.lib.pk_engine_set_output_runs(self._h, mode, None, 0))
        else:
            runs = np.ascontiguousarray(runs, dtype=np.int64).reshape(-1, 2)
            self._check(self.lib.pk_engine_set_output_runs(self._h, mode, _ptr(runs), len(runs)))

    def evaluate(self, x, fct_c=None, fct_o=None, modes=None, outs=None, wait: bool = True):
        """Several callbacks at one ``x`` in a single engine call (``pk_eval_set``): ``x`` and the
        multipliers are uploaded once, the modes run concurrently and each result is copied back as
        soon as it is ready.  Returns ``{mode: array}``; the Hessian is included when ``fct_c`` is given.
        ``wait=False`` (``pk_eval_set_async``) returns once everything is enqueued: the arrays are complete
        after :meth:`sync`, and page-locked inputs must stay untouched until then."""
        if modes is None:
            modes = [P.OBJ, P.GRAD, P.CONS, P.JAC] + ([P.HESS] if fct_c is not None else [])
        modes = list(modes)
        self._load_for_set(modes)
        x = None if x is None else self._x(x)  # None: evaluate at the point already resident on the device
        lam = sig = None
        if P.HESS in modes:
            if fct_c is None:
                raise ValueError("the Hessian needs the multipliers fct_c (and fct_o)")
            lam = np.ascontiguousarray(fct_c, dtype=np.float64)
            if lam.size != self.B * self.lowering.m:
                raise ValueError(f"fct_c must have {self.B * self.lowering.m} entries")
            sig = np.ascontiguousarray(np.broadcast_to(np.asarray(1.0 if fct_o is None else fct_o, dtype=np.float64), (self.B,)))
        bufs = [self._out(m, None if outs is None else outs[k], small_pinned=not wait) for k, m in enumerate(modes)]
        marr = (C.c_int * len(modes))(*modes)
        parr = (C.c_void_p * len(modes))(*[b.ctypes.data for b in bufs])
        call = self.lib.pk_eval_set if wait else self.lib.pk_eval_set_async
        t0 = time.perf_counter()
        self._check(call(self._h, None if x is None else _ptr(x), None if lam is None else _ptr(lam),
                         None if sig is None else _ptr(sig), marr, len(modes), parr))
        self.last_call_ms = 1e3 * (time.perf_counter() - t0)  # host time inside the C call (diagnostics)
        if not wait:
            self._keep = (x, lam, sig)  # the staging copies read these until sync()
        res = {}
        for m, b in zip(modes, bufs):
            if m == P.OBJ and self.B == 1 and wait:
                res[m] = np.float64(b[0])
            else:  # asynchronous call: the objective stays a one-element array, complete after sync()
                res[m] = b if m == P.OBJ else self._shape(m, b)
        return res

    def objective(self, x):
        self.load(P.OBJ)
        x, out = self._x(x), self._out(P.OBJ, None)
        self._check(self.lib.pk_eval_objective(self._h, _ptr(x), _ptr(out)))
        return np.float64(out[0]) if self.B == 1 else out

    def gradient(self, x, out=None):
        self.load(P.GRAD)
        x, out = self._x(x), self._out(P.GRAD, out)
        self._check(self.lib.pk_eval_gradient(self._h, _ptr(x), _ptr(out)))
        return self._shape(P.GRAD, out)

    def constraints(self, x, out=None):
        self.load(P.CONS)
        x, out = self._x(x), self._out(P.CONS, out)
        self._check(self.lib.pk_eval_constraints(self._h, _ptr(x), _ptr(out)))
        return self._shape(P.CONS, out)

    def jacobian(self, x, out=None):
        self.load(P.JAC)
        x, out = self._x(x), self._out(P.JAC, out)
        self._check(self.lib.pk_eval_jacobian(self._h, _ptr(x), _ptr(out)))
        return self._shape(P.JAC, out)

    def hessian(self, x, fct_c, fct_o, out=None):
        self.load(P.HESS)
        x, out = self._x(x), self._out(P.HESS, out)
        lam = np.ascontiguousarray(fct_c, dtype=np.float64)
        if lam.size != self.B * self.lowering.m:
            raise ValueError(f"fct_c must have {self.B * self.lowering.m} entries")
        sig = np.ascontiguousarray(np.broadcast_to(np.asarray(fct_o, dtype=np.float64), (self.B,)))
        self._check(self.lib.pk_eval_hessian(self._h, _ptr(x), _ptr(lam), _ptr(sig), _ptr(out)))
        return self._shape(P.HESS, out)

    def _hessian_part(self, x, fct_c, fct_o, offset: int, count: int):
        """Head / tail of the Hessian values: one evaluation of the mode, and only the requested slots
        cross PCIe (``pk_download_range``) -- SciPy's adapter asks for ``hessian_o`` and ``hessian_c``
        separately at every iterate (``optimizer/scipy.py:64-72``)."""
        self.load(P.HESS)
        if P.HESS in self.compacted or self.fin[P.HESS].get("runs") is not None:
            if x is None:
                raise ValueError("shaped outputs (de-duplicated pattern / mesh shard) need x")
            full = self.hessian(x, fct_c, fct_o)  # shaped outputs: both parts live on one merged pattern / shard
            return np.array(full) if P.HESS in self.compacted else np.array(full[..., offset : offset + count])
        lam = np.ascontiguousarray(fct_c, dtype=np.float64)
        if lam.size != self.B * self.lowering.m:
            raise ValueError(f"fct_c must have {self.B * self.lowering.m} entries")
        sig = np.ascontiguousarray(np.broadcast_to(np.asarray(fct_o, dtype=np.float64), (self.B,)))
        if x is not None:  # None: the point already resident on the device (x-keyed cache)
            x = self._x(x)
            self._check(self.lib.pk_upload_x(self._h, _ptr(x)))
        elif self.x_uploads == 0:
            raise ValueError("x = None reuses the resident point, but none was uploaded yet")
        self._check(self.lib.pk_upload_multipliers(self._h, _ptr(lam), _ptr(sig)))
        self._check(self.lib.pk_run(self._h, P.HESS))
        n = self.B * count
        out = self.pool.take(n) if self.pool is not None and n >= 8192 else np.empty(n, dtype=np.float64)
        self._check(self.lib.pk_download_range(self._h, P.HESS, offset, count, _ptr(out)))
        return out if self.B == 1 else out.reshape(self.B, count)

    def hessian_o(self, x):
        return self._hessian_part(x, self._zero_lam[: self.B * self.lowering.m], self._one, 0, self.lowering.nnz_hess_o)

    def hessian_c(self, x, fct_c):
        return self._hessian_part(x, fct_c, np.zeros(self.B), self.lowering.nnz_hess_o, self.lowering.nnz_hess_c)

    def out_device_pointer(self, mode: int) -> tuple:
        """(device address, values per instance, doubles between instances) of the mode's latest result --
        for device-side consumers such as the NCCL gather of an instance-sharded batch
        (``sharding.ShardedBatch``).  The stride exceeds the count when the values are a slice of the
        set pipeline's combined output."""
        self.load(mode)
        p, n, stride = C.c_void_p(), C.c_int64(), C.c_int64()
        self._check(self.lib.pk_out_device_pointer(self._h, mode, C.byref(p), C.byref(n), C.byref(stride)))
        return p.value, n.value, stride.value

    # ------------------------------------------------------------------ continuous error estimate
    def error_estimation_data(self, x):
        """Per phase ``(T_x_aug, I_f_aug)``, each ``[n_x][rows]``: the data
        ``PhaseBase._error_estimation_data_continuous`` (``phasebase.py:1355-1366``) hands to the
        mesh-refinement checks, computed on the device (programs and operators: ``plan.error_estimate``)."""
        if self.B != 1:
            raise ValueError("error-estimate data is available for engines with batch == 1")
        if getattr(self, "_aug", None) is None:
            ee = self.plan.error_estimate()
            keep, arr = [], (_AugPhase * len(ee["phases"]))()
            for k, ph in enumerate(ee["phases"]):
                a = arr[k]
                a.prep_kernel, a.node_kernel = ph["prep"].encode(), ph["node"].encode()
                for f in ("x_offset", "L", "L_xu", "L_x_all", "n_x", "n_u", "Lm_aug", "rows"):
                    setattr(a, f, ph[f])
                tm = np.ascontiguousarray(ph["tm_aug"], dtype=np.float64)
                keep.append(tm)
                a.tm_aug = tm.ctypes.data
                for name in ("V", "T", "I"):
                    m = ph[name].tocsr()
                    m.sort_indices()
                    ptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
                    idx = np.ascontiguousarray(m.indices, dtype=np.int64)
                    val = np.ascontiguousarray(m.data, dtype=np.float64)
                    keep += [ptr, idx, val]
                    setattr(a, name + "_ptr", ptr.ctypes.data)
                    setattr(a, name + "_idx", idx.ctypes.data)
                    setattr(a, name + "_val", val.ctypes.data)
            opts = (C.c_char_p * max(1, len(self._opts)))(*[o.encode() for o in self._opts])
            self._check(self.lib.pk_engine_load_error_estimate(self._h, ee["source"].encode(), opts, len(self._opts), arr, len(ee["phases"])))
            self._aug = [(ph["n_x"], ph["rows"]) for ph in ee["phases"]]
        x = self._x(x)
        n = sum(nx * rows for nx, rows in self._aug)
        tx, jf = np.empty(max(n, 1)), np.empty(max(n, 1))
        self._check(self.lib.pk_eval_error_data(self._h, _ptr(x), _ptr(tx), _ptr(jf)))
        out, off = [], 0
        for nx, rows in self._aug:
            out.append((tx[off : off + nx * rows].reshape(nx, rows).copy(), jf[off : off + nx * rows].reshape(nx, rows).copy()))
            off += nx * rows
        return out

    # ------------------------------------------------------------------ device-resident path
    def upload(self, x, fct_c=None, fct_o=None):
        x = self._x(x)
        self._check(self.lib.pk_upload_x(self._h, _ptr(x)))
        if fct_c is not None:
            lam = np.ascontiguousarray(fct_c, dtype=np.float64)
            sig = np.ascontiguousarray(np.broadcast_to(np.asarray(1.0 if fct_o is None else fct_o, dtype=np.float64), (self.B,)))
            self._check(self.lib.pk_upload_multipliers(self._h, _ptr(lam), _ptr(sig)))
        self.sync()

    def run(self, mode: int):
        self.load(mode)
        self._check(self.lib.pk_run(self._h, mode))

    def run_set(self, modes):
        """Enqueue several callbacks at the uploaded x as one graph (modes overlap)."""
        self._load_for_set(modes)
        arr = (C.c_int * len(modes))(*modes)
        self._check(self.lib.pk_run_set(self._h, arr, len(modes)))

    def sync(self):
        self._check(self.lib.pk_sync(self._h))

    def download(self, mode: int, out=None):
        out = self._out(mode, out)
        self._check(self.lib.pk_download(self._h, mode, _ptr(out)))
        return self._shape(mode, out)

    def time(self, mode: int, iters: int = 10, stages: bool = False):
        """CUDA-event time (ms) of ``iters`` back-to-back runs; optionally per stage."""
        self.load(mode)
        total = C.c_float()
        st = (C.c_float * (N_STAGES + 2))()
        self._check(self.lib.pk_time(self._h, mode, iters, C.byref(total), st if stages else None))
        return (total.value, list(st)) if stages else total.value

    def time_stage(self, mode: int, stage: int, iters: int = 20, flush_l2: bool = True):
        """CUDA-event times (ms) of ``iters`` single launches of one stage of ``mode`` (``plan.ST_*``; 6 =
        per-node programs), each preceded by an untimed L2 flush."""
        self.load(mode)
        out = (C.c_float * iters)()
        self._check(self.lib.pk_time_stage(self._h, mode, 1 << stage, iters, int(flush_l2), out))
        return list(out)

    def time_stage_alternating(self, modes, stage: int, rounds: int = 20) -> float:
        """Total CUDA-event time (ms) of ``rounds`` x (one launch of ``stage`` for every mode in turn), back to back."""
        for m in modes:
            self.load(m)
        arr = (C.c_int * len(modes))(*modes)
        total = C.c_float()
        self._check(self.lib.pk_time_stage_alternating(self._h, arr, len(modes), 1 << stage, rounds, C.byref(total)))
        return total.value

    def time_steps(self, modes, steps: int, flush_l2: bool = True):
        """Per-step CUDA-event times (ms) of running ``modes`` back to back, inputs resident in HBM."""
        self._load_for_set(modes)
        arr = (C.c_int * len(modes))(*modes)
        out = (C.c_float * steps)()
        self._check(self.lib.pk_time_steps(self._h, arr, len(modes), steps, int(flush_l2), out))
        return list(out)

    def timeline(self, modes):
        """Device-side timeline of one evaluation set: list of (mode, tag, edge, microseconds)."""
        for m in modes:
            self.load(m)
        arr = (C.c_int * len(modes))(*modes)
        rows = np.zeros((256, 4))
        n = C.c_int()
        self._check(self.lib.pk_timeline(self._h, arr, len(modes), _ptr(rows), 256, C.byref(n)))
        return [(int(r[0]), int(r[1]), int(r[2]), float(r[3])) for r in rows[: n.value]]

    def expand_kernel(self, mode: int) -> str:
        """Name of the block-expansion kernel ``mode`` launches ('' if it has none)."""
        self.load(mode)
        v = C.c_int()
        self._check(self.lib.pk_expand_variant(self._h, mode, C.byref(v)))
        return {0: "", 1: "pk_expand_blocks", 2: "pk_expand_cols", 3: "pk_expand_bulk", 4: "pk_expand_batch", 5: "pk_expand_slots"}[v.value]

    def flush_l2(self):
        self._check(self.lib.pk_flush_l2(self._h))

    @property
    def x_uploads(self) -> int:
        n = C.c_int64()
        self._check(self.lib.pk_x_uploads(self._h, C.byref(n)))
        return n.value

    @property
    def launches(self) -> int:
        n = C.c_int64()
        self._check(self.lib.pk_kernel_launches(self._h, C.byref(n)))
        return n.value

    def cubin(self, mode: int) -> bytes:
        self.load(mode)
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.pk_engine_get_cubin(self._h, mode, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)
