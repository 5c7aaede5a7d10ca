#!/usr/bin/env python
"""Benchmark of the NLP-callback hot path.

A *step* is one callback evaluation set -- objective + gradient + constraints +
Jacobian values + Hessian-of-Lagrangian values at the same (x, lambda, sigma) --
of BASELINE.json's configs[1]: robot_arm on the LGR transcription, 2000
intervals x 20 points (40 000 collocation nodes, L=360 008, m=240 000,
nnz_J=12 479 754, nnz_H=12 799 740).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1
* ``value``  device-resident throughput: inputs already in HBM, CUDA events on the
  engine's stream around each step, L2 flushed (untimed) between steps.
* ``e2e``    the same metric through the public API with HOST buffers: every step copies
  x, lambda, sigma to the device and all five results back.  ``e2e.value`` is the default
  contract (System.evaluate -> pk_eval_set; results are arrays the caller owns, leased from
  the engine's page-locked pool); ``e2e.pageable_fresh_arrays`` / ``e2e.reused_buffers`` /
  ``e2e.five_callbacks`` are the other ownership / call styles.
* ``roofline`` the dominant kernel (block expansion of the Hessian) launch by launch with the
  L2 flushed, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
* ``all_configs`` the other BASELINE.json configurations (C1, C3, C4, C5) in short form.
* ``cpu_baseline`` the reference's own Numba path on one host core (oracle/_ref, staged by
  ``__graft_entry__.build()`` in the build container), else the oracle port.

N > 1 (torchrun, one rank per GPU) -- the north-star splits, total work fixed (strong scaling):
* ``value`` / ``e2e``: ONE robot_arm mesh sharded over the N GPUs (pockit_b200.meshshard): every
  rank expands and copies back its share of the Jacobian / Hessian values over its own PCIe link;
  no data-path collective.
* ``c5_batched``: BASELINE configs[4], 8192 planar_quadrotor instances sharded by instance, with
  the NCCL all-gather of the results timed separately.

``--impl reference``: the reference's CPU implementation of the same set on all usable host
cores (one single-threaded process per core, the most the reference can use).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
REF_STAGE = ROOT / "oracle" / "_ref"

METRIC = "NLP callback eval-sets/s (objective+gradient+constraints+Jacobian+Hessian) at 40k nodes"
UNIT = "eval-sets/s"
WORKLOADS = {
    "robot_arm": (dict(builder="robot_arm", scheme="radau", mesh=2000, num_point=20),
                  "robot_arm LGR 2000x20 (40000 nodes; BASELINE.json configs[1])"),
    "lqr": (dict(builder="lqr", scheme="lobatto", mesh=10, num_point=10), "LQR LGL 10x10 (BASELINE.json configs[0])"),
    "humanoid": (dict(builder="humanoid", scheme="lobatto", mesh=11112, num_point=10), "humanoid LGL 11112x10 (configs[2])"),
    "humanoid_small": (dict(builder="humanoid", scheme="lobatto", mesh=1000, num_point=10), "humanoid LGL 1000x10 (configs[2])"),
    "rocket": (dict(builder="rocket", scheme="lobatto", mesh=5556, num_point=10), "two-stage rocket LGL 2x5556x10 (configs[3])"),
    "quadrotor": (dict(builder="quadrotor", scheme="lobatto", mesh=14, num_point=6), "quadrotor LGL 14x6, one instance (configs[4])"),
}
C5_INSTANCES = 8192


def build_system(workload: str = "robot_arm", seed_shift: int = 0, package: str = "pockit_b200"):
    """The workload's System on either implementation (``package`` = 'pockit_b200' or 'pockit')."""
    import importlib

    from pockit_b200 import problems

    wl = WORKLOADS[workload][0]
    mod = importlib.import_module(f"{package}.{wl['scheme']}")
    S = problems.BUILDERS[wl["builder"]](mod, mesh=wl["mesh"], num_point=wl["num_point"])
    x, lam, sigma = problems.evaluation_point(S, seed=1 + seed_shift)
    return S, x, lam, sigma


def config_of(lo, workload: str, world: int) -> dict:
    """``config`` of the JSON line -- identical in both arms (workload, sizes, timing conditions)."""
    L, m, nj, nh = int(lo.r_s), int(lo.m), int(lo.nnz_jac), int(lo.nnz_hess_o + lo.nnz_hess_c)
    par = ("one instance on one GPU" if world == 1 else
           f"ONE mesh sharded over {world} GPUs by (list, interval) tiles, no data-path collective")
    return {
        "workload": WORKLOADS[workload][1], "L": L, "m": m, "nnz_jac": nj, "nnz_hess": nh,
        "set_algorithmic_MB": 8 * (6 * L + 2 * m + nj + nh) / 1e6,
        "l2": "GPU arm: flushed between steps (256 MiB fill, untimed); outputs 207 MB > 126 MB L2",
        "parallelism": par,
    }


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled through NVML (same counters nvidia-smi prints)
    every ~2 ms from a thread while the timed regions run."""

    def __init__(self, index: int):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.h = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def __enter__(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        return self

    def _loop(self):
        nv = self.nv
        while not self.stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, reasons))
            except Exception:
                pass
            time.sleep(0.002)

    def __exit__(self, *exc):
        self.stop.set()
        if self.h is not None:
            self.thread.join(timeout=1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        seen = sorted(k for k, bit in names.items() if any(r & bit for _, r in self.rows))
        return {"sm_mhz": statistics.median(sm for sm, _ in self.rows), "sm_max_mhz": self.max_sm,
                "reasons": seen, "samples": len(self.rows)}


# --------------------------------------------------------------------------- CPU arm
def reference_available() -> bool:
    """The reference staged under oracle/_ref (``__graft_entry__.stage_reference``) and Numba importable."""
    if not (REF_STAGE / "pockit" / "__init__.py").exists():
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


def _cpu_worker(kind, workload, seed_shift, warmup, steps, budget_s, barrier, q):
    """One single-threaded CPU process: build the model, warm up, wait for the others, then time
    ``steps`` evaluation sets twice.  ``kind`` 'reference' = the reference's own code (oracle/_ref),
    'port' = oracle/pockit_oracle.py."""
    try:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        os.environ.setdefault("NUMBA_NUM_THREADS", "1")
        if kind == "reference":
            sys.path.insert(0, str(REF_STAGE))
            S, x, lam, sigma = build_system(workload, seed_shift, package="pockit")

            def one():  # the reference substitutes boundary values into the caller's x: hand it copies
                S.objective(x.copy()); S.gradient(x.copy()); S.constraints(x.copy()); S.jacobian(x.copy())
                S.hessian(x.copy(), lam, sigma)
        else:
            from oracle.pockit_oracle import OracleSystem

            S, x, lam, sigma = build_system(workload, seed_shift)
            O = OracleSystem(S)

            def one():
                O.objective(x); O.gradient(x); O.constraints(x); O.jacobian(x); O.hessian(x, lam, sigma)

        one()  # JIT / lambdify / page faults
        t0 = time.perf_counter()
        for _ in range(max(1, warmup)):
            one()
        per = (time.perf_counter() - t0) / max(1, warmup)
        if barrier is not None:
            barrier.wait()
        # bounded sample: as many of the requested steps as fit the time budget (at least 2)
        n = int(max(2, min(steps, budget_s / max(per, 1e-9))))
        passes = []
        for _ in range(2):
            t0 = time.perf_counter()
            for _ in range(n):
                one()
            passes.append(time.perf_counter() - t0)
            if barrier is not None:
                barrier.wait()
        q.put(("ok", n, passes))
    except Exception as exc:  # noqa: BLE001
        if barrier is not None:
            barrier.abort()
        q.put(("error", repr(exc), []))


def cpu_eval_sets_per_s(kind: str, workload: str, workers: int, warmup: int, steps: int, budget_s: float):
    """``workers`` independent single-threaded processes evaluating concurrently (the reference has no
    threading of its own: this is the most it can use of the host).  Returns
    ``(sets/s of the better pass, sets/s of both passes, steps executed per worker, slowest ms per set)``."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    barrier = ctx.Barrier(workers) if workers > 1 else None
    procs = [ctx.Process(target=_cpu_worker, args=(kind, workload, i, warmup, steps, budget_s, barrier, q)) for i in range(workers)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    bad = [r for r in res if r[0] != "ok"]
    if bad:
        raise RuntimeError(f"CPU worker failed: {bad[0][1]}")
    n = min(r[1] for r in res)
    rates = [sum(r[1] / r[2][k] for r in res) for k in range(2)]
    worst = max(max(r[2]) / r[1] for r in res)
    return max(rates), rates, n, 1000.0 * worst


def usable_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args, rank, world):
    if rank != 0:
        return
    kind = "reference" if reference_available() and not args.port else "port"
    S, _, _, _ = build_system(args.workload)
    cores = args.cores or usable_cores()
    try:  # ~3 GB per worker at 40 k nodes (patterns + value arrays + Numba)
        avail = int(next(l for l in open("/proc/meminfo") if l.startswith("MemAvailable")).split()[1]) // (1 << 20)
        cores = max(1, min(cores, avail // 4))
    except Exception:
        pass
    value, passes, n, worst_ms = cpu_eval_sets_per_s(kind, args.workload, cores, min(args.warmup, 2), args.steps, budget_s=45.0)
    what = ("the reference's own package (oracle/_ref/pockit, Numba) through its public System callbacks" if kind == "reference"
            else "oracle/pockit_oracle.py (NumPy restatement; the reference was not staged / Numba is missing)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": n, "steps_requested": args.steps, "warmup": min(args.warmup, 2), "ms_per_step": worst_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_of(S.lowering, args.workload, world),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": cores, "kind": kind, "passes": passes,
            "sample": f"{n} eval-set(s) per worker and pass, 2 passes (value = the better one), after JIT + warm-up, on {cores} "
                      f"concurrent single-threaded worker processes: {what}",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm helpers
def load_peaks():
    peaks, source = {}, "fallback (B200_PROFILING.md: 6.65 TB/s)"
    for pk, label in ((ROOT / "MEASURED_PEAKS.json", "MEASURED_PEAKS.json"),
                      (ROOT / "profiles" / "r01_measured_peaks.json",
                       "profiles/r01_measured_peaks.json (copy of the driver's round-1 MEASURED_PEAKS.json)")):
        if pk.exists():
            peaks, source = json.loads(pk.read_text()), label
            break
    return float(peaks.get("hbm_gbs", 6650.0)), source


def _expand_bytes(eng, P, mode):
    """Algorithmic bytes of one launch of ``mode``'s block expansion: slots written + list rows read
    (+ multipliers for the Hessian), DESIGN.md section 4."""
    jobs = eng.fin[mode]["jobs"][P.ST_EXPAND]
    if not len(jobs):
        return 0
    slots = int(sum(int(j["i"][1]) * int(j["i"][11]) * int(j["i"][4]) for j in jobs))
    rows_read = int(sum(int(j["i"][1]) * int(j["i"][11]) for j in jobs))
    lam_read = int(sum((int(j["i"][11]) // int(j["i"][3])) * int(j["i"][4]) for j in jobs)) if mode == P.HESS else 0
    return 8 * eng.B * (slots + rows_read + lam_read)


def expansion_roofline(eng, lo, P, peak, iters=20):
    """Dominant kernel = the block expansion.  Two live measurements with CUDA events on the engine stream:
    * sustained: the Jacobian and the Hessian expansion launched alternately, back to back, as they follow
      each other inside a set; their outputs together exceed the L2 (robot_arm 207 MB > 126 MB), so every
      byte drains to HBM -- this is ``achieved`` / ``frac``;
    * cold: single Hessian launches, each behind an untimed L2 flush (dirty flush lines to evict, list
      rows and multipliers from HBM) -- ``cold_launch_ms`` / ``cold_frac``."""
    for m in (P.JAC, P.HESS):
        eng.run(m)
    eng.sync()
    b_j, b_h = _expand_bytes(eng, P, P.JAC), _expand_bytes(eng, P, P.HESS)
    if not b_h:
        return None
    eng.time_stage_alternating([P.JAC, P.HESS], P.ST_EXPAND, rounds=3)
    pair_ms = eng.time_stage_alternating([P.JAC, P.HESS], P.ST_EXPAND, rounds=iters) / iters
    n_launch = 2 if b_j else 1
    k_ms = pair_ms / n_launch
    alg = (b_j + b_h) / n_launch
    ach = alg / (k_ms * 1e-3) / 1e9
    cold = statistics.median(eng.time_stage(P.HESS, P.ST_EXPAND, iters=iters, flush_l2=True))
    return {"kernel": f"{eng.expand_kernel(P.HESS)} (Jacobian and Hessian launches alternating)", "achieved": ach, "peak": peak,
            "frac": ach / peak, "algorithmic_bytes_per_launch": alg, "launch_ms": k_ms, "launches_timed": n_launch * iters,
            "output_MB_per_pair": (b_j + b_h) / 1e6, "cold_launch_ms": cold, "cold_frac": b_h / (cold * 1e-3) / 1e9 / peak,
            "hessian_bytes_per_launch": b_h}


def platform_d2h_GBps(torch, nbytes_per_rank: int, maxed, barrier, reps: int = 5) -> float:
    """What the box gives: every rank copies ``nbytes_per_rank`` from its GPU into its own page-locked host
    buffer at the same time (plain cudaMemcpyAsync, CUDA events, max over ranks).  Returns GB/s PER RANK;
    the end-to-end path is bound by world x this figure (the device-to-host link(s) and host memory)."""
    n = max(1, nbytes_per_rank // 8)
    dev = torch.empty(n, dtype=torch.float64, device="cuda").fill_(1.0)
    host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    host.copy_(dev)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        host.copy_(dev, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    t = maxed(a.elapsed_time(b) / 1e3) / reps
    return 8 * n / t / 1e9


def short_config(name, S, x, lam, sigma, P, peak, steps, batch=1, fixed=None):
    """One of the other BASELINE configurations in short form: device-resident set time (flushed), e2e
    through the host API, roofline fractions."""
    from pockit_b200.engine import Engine

    lo = S.lowering
    eng = Engine(lo, batch=batch, fastmath=S._fastmath, fixed=fixed)
    modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    for m in modes:
        eng.load(m)
    eng.upload(x, lam, sigma)
    eng.time_steps(modes, 3, flush_l2=True)
    ms = eng.time_steps(modes, steps, flush_l2=True)
    dev = statistics.mean(ms)
    for _ in range(2):
        eng.evaluate(x, lam, sigma)
    t0 = time.perf_counter()
    n_e2e = max(3, min(steps, 10))
    for _ in range(n_e2e):
        eng.evaluate(x, lam, sigma)
    e2e = (time.perf_counter() - t0) / n_e2e
    L, m, nj, nh = lo.r_s, lo.m, lo.nnz_jac, lo.nnz_hess_o + lo.nnz_hess_c
    set_bytes = 8 * batch * (6 * L + 2 * m + nj + nh)
    rec = {
        "workload": name, "instances": batch, "nodes": int(sum(p.L_m for p in lo.phases)), "L": int(L), "m": int(m),
        "nnz_jac": int(nj), "nnz_hess": int(nh), "device_ms_per_set": dev, "device_eval_sets_per_s": batch / (dev * 1e-3),
        "set_algorithmic_MB": set_bytes / 1e6, "set_roofline_frac": set_bytes / (dev * 1e-3) / 1e9 / peak,
        "e2e_ms_per_set": 1000.0 * e2e, "e2e_eval_sets_per_s": batch / e2e, "expand_kernel": eng.expand_kernel(P.HESS),
    }
    rf = expansion_roofline(eng, lo, P, peak, iters=10)
    if rf is not None:
        rec["expansion_roofline_frac"] = rf["frac"]
        rec["expansion_launch_ms"] = rf["launch_ms"]
        rec["expansion_cold_frac"] = rf["cold_frac"]
    eng.close()
    return rec


def port_sets_per_s(workload: str, sets: int) -> float:
    """Oracle port on one host core (the CPU column of ``all_configs``)."""
    from oracle.pockit_oracle import OracleSystem

    S, x, lam, sigma = build_system(workload)
    O = OracleSystem(S)

    def one():
        O.objective(x); O.gradient(x); O.constraints(x); O.jacobian(x); O.hessian(x, lam, sigma)

    one()
    t0 = time.perf_counter()
    for _ in range(sets):
        one()
    return sets / (time.perf_counter() - t0)


def all_configs(P, peak, steps):
    """C1, C3, C4 of BASELINE.json next to the benchmarked C2 (short form, N = 1 only; C5 is the
    ``c5_batched`` record)."""
    out = {}
    for key, wl, cpu_sets in (("C1", "lqr", 200), ("C3", "humanoid", 2), ("C4", "rocket", 3)):
        S, x, lam, sigma = build_system(wl)
        rec = short_config(WORKLOADS[wl][1], S, x, lam, sigma, P, peak, steps)
        rec["cpu_port_1core_eval_sets_per_s"] = port_sets_per_s(wl, cpu_sets)
        rec["e2e_speedup_vs_cpu_port_1core"] = rec["e2e_eval_sets_per_s"] / rec["cpu_port_1core_eval_sets_per_s"]
        out[key] = rec
    return out


def c5_sharded(P, peak, steps, rank, world, local, dist, torch):
    """BASELINE configs[4] strong-scaled: 8192 instances dealt out by instance (b mod N), every rank on
    its own engine; compute needs no collective, the NCCL all-gather that assembles the batch for a
    caller who wants it on one device is timed separately (CUDA events, max over ranks)."""
    import pockit_b200.lobatto as lob
    from pockit_b200 import problems
    from pockit_b200.batched import fixed_index, fixed_table
    from pockit_b200.sharding import ShardedBatch

    S = problems.quadrotor(lob)
    n = C5_INSTANCES
    rng = np.random.default_rng(0)
    fixed = fixed_table(S, n)
    fixed[:, fixed_index(S, 0, "x0", 0)] += rng.uniform(-0.2, 0.2, n)
    fixed[:, fixed_index(S, 0, "x0", 1)] += rng.uniform(-0.2, 0.2, n)
    x0, lam0, sigma = problems.evaluation_point(S)
    X = x0[None, :] + 1e-2 * rng.normal(size=(n, len(x0)))
    LAM = lam0[None, :] + 0.1 * rng.normal(size=(n, len(lam0)))
    sb = ShardedBatch(S, fixed, rank=rank, world=world)
    eng = sb.local.engine
    modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    Xl, LAMl, sigl = sb._take(X), sb._take(LAM), np.full(len(sb.idx), sigma)
    eng.upload(Xl, LAMl, sigl)

    def maxed(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    eng.time_steps(modes, 3, flush_l2=True)
    barrier()
    ms = eng.time_steps(modes, steps, flush_l2=True)
    compute = maxed(sum(ms) / 1e3) / steps
    # NCCL all-gather of all five results (device to device; the engine buffers are read in place)
    gather = None
    ok = True
    if dist is not None:
        eng.run_set(modes)
        eng.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(3 + steps):
            if it == 3:
                barrier()
                a.record()
            got = [sb.gather_device(m) for m in modes]
        b.record()
        torch.cuda.synchronize()
        gather = maxed(a.elapsed_time(b) / 1e3) / steps
        jac_local = eng.download(P.JAC).reshape(len(sb.idx), -1)
        ok = bool(torch.equal(got[3][sb.idx.tolist()].cpu(), torch.from_numpy(np.ascontiguousarray(jac_local))))
    # end to end per rank: host buffers in, host buffers out, own PCIe link
    for _ in range(2):
        eng.evaluate(Xl, LAMl, sigl)
    barrier()
    n_e2e = max(3, min(steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        eng.evaluate(Xl, LAMl, sigl)
    e2e = maxed(time.perf_counter() - t0) / n_e2e
    lo = S.lowering
    per = 8 * (6 * lo.r_s + 2 * lo.m + lo.nnz_jac + lo.nnz_hess_o + lo.nnz_hess_c)
    rec = {
        "workload": f"planar_quadrotor LGL 14x6, {n} instances sharded by instance (b mod N) over {world} GPU(s); BASELINE.json configs[4]",
        "scaling": "strong", "n_gpus": world, "steps": steps,
        "device_ms_per_batch_set": 1e3 * compute, "device_instance_sets_per_s": n / compute,
        "device_algorithmic_GBps_total": per * n / compute / 1e9, "set_roofline_frac_per_gpu": per * n / world / compute / 1e9 / peak,
        "e2e_ms_per_batch_set": 1e3 * e2e, "e2e_instance_sets_per_s": n / e2e,
        "nccl_all_gather_ms": None if gather is None else 1e3 * gather,
        "gathered_MB_per_rank": 8 * n * (1 + lo.r_s + lo.m + lo.nnz_jac + lo.nnz_hess_o + lo.nnz_hess_c) / 1e6,
        "device_plus_gather_instance_sets_per_s": None if gather is None else n / (compute + gather),
        "gather_matches_local_shard": ok,
        "note": "evaluation has no collective; the all-gather (all five value arrays of all instances to every rank, NVLink) "
                "is only for a caller that wants the assembled batch on one device and is bound by the receiving link",
    }
    sb.local.close()
    return rec


# --------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="robot_arm", choices=sorted(WORKLOADS),
                    help="CPU arm only (--impl reference): time another BASELINE.json configuration")
    ap.add_argument("--cores", type=int, default=0, help="CPU arm only: worker processes (default: all usable host cores)")
    ap.add_argument("--port", action="store_true", help="CPU arm only: time the oracle port even if the reference is staged")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-compact", action="store_true", help="skip the extra de-duplicated-pattern measurement")
    ap.add_argument("--no-all-configs", action="store_true", help="skip the short-form C1 / C3 / C4 / C5 records")
    ap.add_argument("--no-c5", action="store_true", help="N > 1: skip the instance-sharded C5 record")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload != "robot_arm":
        raise SystemExit("bench.py: --workload selects the CPU arm's configuration only; the GPU arm measures BASELINE.json configs[1]")

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as graft

    graft.build()
    from pockit_b200 import plan as P
    from pockit_b200.engine import Engine

    S, x, lam, sigma = build_system("robot_arm")
    lo = S.lowering
    peak, peak_source = load_peaks()
    modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    L, m, nj, nh = lo.r_s, lo.m, lo.nnz_jac, lo.nnz_hess_o + lo.nnz_hess_c
    h2d = 8 * (L + m + 1)
    d2h = 8 * (1 + L + m + nj + nh)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxed(v):
        if dist is None:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -------------------------------------------------------
    # N = 1: the whole mesh on one engine.  N > 1: rank g plans its share of the mesh (the per-node
    # programs are replicated, the slot-streaming work is split); all ranks step concurrently.
    eng = Engine(lo, device=local, shard=None if world == 1 else (rank, world))
    my_modes = modes if (world == 1 or rank == 0) else [P.JAC, P.HESS]
    for md in my_modes:
        eng.load(md)
    eng.upload(x, lam, sigma)
    eng.time_steps(my_modes, max(3, args.warmup), flush_l2=True)
    barrier()
    launches0 = eng.launches
    clk = ClockSampler(local)
    clk.__enter__()
    ms = eng.time_steps(my_modes, args.steps, flush_l2=True)
    eng.sync()
    launches = eng.launches - launches0
    barrier()
    t_max = maxed(sum(ms) / 1000.0)
    value = args.steps / t_max
    launches_all = launches
    if dist is not None:
        t = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches_all = int(t.item())

    line = None
    if world == 1:
        line = bench_single(args, S, eng, x, lam, sigma, P, peak, peak_source, clk, value, t_max, launches_all, (L, m, nj, nh, h2d, d2h))
    else:
        line = bench_sharded(args, S, eng, x, lam, sigma, P, peak, peak_source, clk, value, t_max, launches_all, (L, m, nj, nh, h2d, d2h),
                             rank, world, local, dist, torch, barrier)
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def bench_single(args, S, eng, x, lam, sigma, P, peak, peak_source, clk, value, t_max, launches, dims):
    import torch

    L, m, nj, nh, h2d, d2h = dims
    lo = S.lowering
    modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    S._engine = eng  # the public callbacks below run on this engine

    def timed_host(fn):
        for _ in range(max(3, args.warmup)):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def one_set_call():
        S.evaluate(x, lam, sigma)

    def one_set():
        S.objective(x); S.gradient(x); S.constraints(x); S.jacobian(x); S.hessian(x, lam, sigma)

    # default contract: every call returns arrays the caller owns (leases on the page-locked pool)
    S.pinned_outputs = False
    e2e_t = timed_host(one_set_call)
    e2e_five_t = timed_host(one_set)
    clk.__exit__()
    e2e_each = {}
    for cname, fn in (("objective", lambda: S.objective(x)), ("gradient", lambda: S.gradient(x)),
                      ("constraints", lambda: S.constraints(x)), ("jacobian", lambda: S.jacobian(x)),
                      ("hessian", lambda: S.hessian(x, lam, sigma))):
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        e2e_each[cname] = 1000.0 * (time.perf_counter() - t0) / 5
    # a caller that keeps the previous results while asking for the next (two leases alive at a time)
    keep = [S.evaluate(x, lam, sigma)]
    for _ in range(3):  # the pool grows to three buffers per callback here (page-locking is slow, once)
        keep = [keep[-1], S.evaluate(x, lam, sigma)]
    t0 = time.perf_counter()
    n_hold = max(3, min(args.steps, 10))
    for _ in range(n_hold):
        keep = [keep[-1], S.evaluate(x, lam, sigma)]
    e2e_hold_t = (time.perf_counter() - t0) / n_hold
    del keep
    # opt-in: views of engine-owned buffers, overwritten by the next call of the same callback
    S.pinned_outputs = True
    e2e_reuse_t = timed_host(one_set_call)
    S.pinned_outputs = False
    # plain pageable NumPy arrays (no page-locked pool)
    pool, eng.pool = eng.pool, None
    n_page = max(3, min(args.steps, 10))
    one_set_call()
    t0 = time.perf_counter()
    for _ in range(n_page):
        one_set_call()
    e2e_page_t = (time.perf_counter() - t0) / n_page
    eng.pool = pool

    link = platform_d2h_GBps(torch, d2h, lambda v: v, torch.cuda.synchronize)
    compact = None
    if not args.no_compact:  # opt-in de-duplicated patterns (outside the reference's pattern contract)
        S.compact_patterns = True
        cj, ch = len(S.jacobianstructure()[0]), len(S.hessianstructure()[0])
        for _ in range(3):
            one_set_call()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one_set_call()
        tc = (time.perf_counter() - t0) / args.steps
        dev_ms = eng.time_steps(modes, min(args.steps, 20), flush_l2=True)
        compact = {"note": "System.compact_patterns=True: duplicates summed on the device, unique (row, col) pairs only",
                   "nnz_jac": cj, "nnz_hess": ch, "e2e_eval_sets_per_s": 1.0 / tc, "e2e_ms_per_step": 1000.0 * tc,
                   "d2h_bytes_per_step": 8 * (1 + L + m + cj + ch),
                   "device_resident_ms_per_step": sum(dev_ms) / len(dev_ms)}
        S.compact_patterns = False

    rf = expansion_roofline(eng, lo, P, peak)
    traffic, traffic_source = None, None
    tr = ROOT / "profiles" / "r02_expand_traffic.json"
    if tr.exists():  # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture
        rec = json.loads(tr.read_text())
        name = rf["kernel"].split(" ")[0]
        hit = [r for r in rec["launches"] if r["kernel"].startswith(name)]
        if hit:  # per launch, averaged over the Jacobian and the Hessian launch like `achieved`
            traffic = sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in hit) / len(hit)
            traffic_source = f"profiles/r02_expand_traffic.json (ncu --set full, commit {rec.get('commit', '?')})"
    per_mode = {}
    for mname, mm in zip(("objective", "gradient", "constraints", "jacobian", "hessian"), modes):
        per_mode[mname] = statistics.median(eng.time_steps([mm], 10, flush_l2=True))
    set_bytes = 8 * (6 * L + 2 * m + nj + nh)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_of(lo, "robot_arm", 1),
        "set_roofline": {"algorithmic_MB": set_bytes / 1e6, "achieved_GBps": set_bytes / (t_max / args.steps) / 1e9,
                         "frac": set_bytes / (t_max / args.steps) / 1e9 / peak, "ms_per_callback_flushed": per_mode},
        "roofline": {
            "bound": "hbm", "kernel": rf["kernel"], "achieved": rf["achieved"], "peak": peak, "unit": "GB/s",
            "frac": rf["frac"], "traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_source,
            "algorithmic_bytes_per_launch": rf["algorithmic_bytes_per_launch"], "launch_ms": rf["launch_ms"],
            "how": f"{rf['launches_timed']} launches back to back, Jacobian and Hessian expansion alternating as inside a set "
                   f"({rf['output_MB_per_pair']:.0f} MB of output per pair > 126 MB L2: sustained streaming), CUDA events on the engine stream",
            "cold_launch_ms": rf["cold_launch_ms"], "cold_frac": rf["cold_frac"],
            "cold_how": "median of single Hessian-expansion launches, each behind an untimed L2 flush (256 MiB fill)",
        },
        "e2e": {"value": args.steps / e2e_t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1000.0 * e2e_t / args.steps,
                "platform_d2h_GBps": link, "frac_of_platform_d2h": (d2h / (e2e_t / args.steps) / 1e9) / link,
                "api": "System.evaluate(x, lam, sigma) -> pk_eval_set: one upload, per-mode streams, copies overlapped",
                "outputs": "arrays owned by the caller (leases on the engine's page-locked pool; nothing is overwritten while referenced)",
                "caller_keeps_previous_results": {"value": 1.0 / e2e_hold_t, "ms_per_step": 1000.0 * e2e_hold_t},
                "pageable_fresh_arrays": {"value": 1.0 / e2e_page_t, "ms_per_step": 1000.0 * e2e_page_t,
                                          "outputs": "np.empty per call (POCKIT_B200_PINNED_POOL=0)"},
                "reused_buffers": {"value": args.steps / e2e_reuse_t, "ms_per_step": 1000.0 * e2e_reuse_t / args.steps,
                                   "outputs": "opt-in System.pinned_outputs=True: views of engine buffers, valid until the next call"},
                "five_callbacks": {"value": args.steps / e2e_five_t, "ms_per_step": 1000.0 * e2e_five_t / args.steps,
                                   "h2d_bytes_per_step": 8 * (5 * L + m + 1), "ms_per_callback": e2e_each,
                                   "api": "System.objective/gradient/constraints/jacobian/hessian one after the other"}},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if compact is not None:
        line["compact_patterns"] = compact
    if not args.no_c5:
        c5 = c5_sharded(P, peak, min(args.steps, 20), 0, 1, 0, None, torch)
        c5["cpu_port_1core_eval_sets_per_s"] = port_sets_per_s("quadrotor", 200)
        c5["e2e_speedup_vs_cpu_port_1core"] = c5["e2e_instance_sets_per_s"] / c5["cpu_port_1core_eval_sets_per_s"]
        line["c5_batched"] = c5
    if not args.no_all_configs:
        line["all_configs"] = all_configs(P, peak, min(args.steps, 20))
        if "c5_batched" in line:
            line["all_configs"]["C5"] = "see c5_batched"
        c2 = {"workload": WORKLOADS["robot_arm"][1], "instances": 1, "device_ms_per_set": line["ms_per_step"],
              "device_eval_sets_per_s": value, "set_roofline_frac": line["set_roofline"]["frac"],
              "e2e_ms_per_set": line["e2e"]["ms_per_step"], "e2e_eval_sets_per_s": line["e2e"]["value"],
              "expansion_roofline_frac": rf["frac"], "expansion_cold_frac": rf["cold_frac"]}
        line["all_configs"]["C2"] = c2
    if not args.no_cpu_baseline:
        kind = "reference" if reference_available() else "port"
        v, passes, n, _ = cpu_eval_sets_per_s(kind, "robot_arm", 1, 1, 3, budget_s=20.0)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": 1, "kind": kind, "passes": passes,
            "sample": f"{n} eval-sets of the same workload per pass (2 passes, the better one) after JIT + warm-up, one host core, "
                      + ("the reference's own package (oracle/_ref/pockit, Numba)" if kind == "reference" else "oracle/pockit_oracle.py"),
        }
        if "all_configs" in line:
            line["all_configs"]["C2"]["cpu_reference_1core_eval_sets_per_s" if kind == "reference" else "cpu_port_1core_eval_sets_per_s"] = v
            line["all_configs"]["C2"]["e2e_speedup_vs_cpu_1core"] = line["e2e"]["value"] / v
    return line


def bench_sharded(args, S, eng, x, lam, sigma, P, peak, peak_source, clk, value, t_max, launches, dims,
                  rank, world, local, dist, torch, barrier):
    from pockit_b200.engine import Engine
    from pockit_b200.meshshard import MeshShardedSystem

    L, m, nj, nh, h2d, d2h = dims
    lo = S.lowering
    eng.close()

    def maxed(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    link = platform_d2h_GBps(torch, d2h // world, maxed, barrier)  # all ranks copy their share concurrently
    # ---- end to end: rank 0 is the caller (host buffers in / out), the other ranks serve their shares
    ms = MeshShardedSystem(S, rank=rank, world=world, device=local)
    ms.pinned_outputs = True  # results are views of the shared page-locked mapping all ranks copy into
    line = None
    if rank != 0:
        ms.serve()
        clk.__exit__()
    else:
        want = None
        try:
            ref = Engine(lo, device=local)  # unsharded engine on the same GPU: expected values
            r0 = ref.evaluate(x, lam, sigma)
            want = {"objective": r0[P.OBJ], "gradient": np.array(r0[P.GRAD]), "constraints": np.array(r0[P.CONS]),
                    "jacobian": np.array(r0[P.JAC]), "hessian": np.array(r0[P.HESS])}
            del r0
            rf = expansion_roofline(ref, lo, P, peak)  # the dominant kernel, same launches as at N = 1 (whole mesh, rank 0's GPU)
            ref.close()
            r = ms.evaluate(x, lam, sigma)
            mismatch = [k for k in want if not np.array_equal(np.asarray(r[k]), np.asarray(want[k]))]
            exact = not mismatch
            for _ in range(max(3, args.warmup)):
                ms.evaluate(x, lam, sigma)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ms.evaluate(x, lam, sigma)
            e2e_t = time.perf_counter() - t0
            timeline = dict(ms.last_timeline)
            ms_rates = ms.link_rates
            r = ms.evaluate(x, lam, sigma)  # and once more after the timed loop
            mismatch += [k + " (after the timed loop)" for k in want if not np.array_equal(np.asarray(r[k]), np.asarray(want[k]))]
            exact = not mismatch
            each = {}
            for name, fn in (("jacobian", lambda: ms.jacobian(x)), ("hessian", lambda: ms.hessian(x, lam, sigma))):
                t0 = time.perf_counter()
                for _ in range(5):
                    fn()
                each[name] = 1000.0 * (time.perf_counter() - t0) / 5
        finally:
            ms.close()
        clk.__exit__()
        set_bytes = 8 * (6 * L + 2 * m + nj + nh)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * t_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_of(lo, "robot_arm", world),
            "set_roofline": {"algorithmic_MB": set_bytes / 1e6, "achieved_GBps_total": set_bytes / (t_max / args.steps) / 1e9,
                             "frac_of_all_gpus": set_bytes / (t_max / args.steps) / 1e9 / (peak * world)},
            "roofline": {
                "bound": "hbm", "kernel": rf["kernel"], "achieved": rf["achieved"], "peak": peak, "unit": "GB/s", "frac": rf["frac"],
                "traffic": None, "peak_source": peak_source, "algorithmic_bytes_per_launch": rf["algorithmic_bytes_per_launch"],
                "launch_ms": rf["launch_ms"], "cold_launch_ms": rf["cold_launch_ms"], "cold_frac": rf["cold_frac"],
                "how": "measured on rank 0's GPU with the WHOLE mesh on one engine (the kernel every rank runs on its share): "
                       "Jacobian and Hessian launches alternating back to back, CUDA events on the engine stream",
            },
            "e2e": {"value": args.steps / e2e_t, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1000.0 * e2e_t / args.steps, "ms_per_callback": each,
                    "platform_d2h_GBps_aggregate": link * world, "platform_d2h_GBps_per_rank": link,
                    "frac_of_platform_d2h": (d2h / (e2e_t / args.steps) / 1e9) / (link * world),
                    "api": "MeshShardedSystem.evaluate(x, lam, sigma) on rank 0: x published in a shared page-locked mapping, every rank "
                           "copies its share of the Jacobian / Hessian values back over its own PCIe link",
                    "bit_identical_to_unsharded": bool(exact), "mismatching_outputs": mismatch, "collective_in_data_path": "none",
                    "last_set_timeline_ms": timeline,
                    "share_weights_from_link_rates_GBps": ms_rates},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
    barrier()
    if not args.no_c5:
        rec = c5_sharded(P, peak, min(args.steps, 20), rank, world, local, dist, torch)
        if line is not None:
            line["c5_batched"] = rec
    return line


if __name__ == "__main__":
    main()
