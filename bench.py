#!/usr/bin/env python
"""Benchmark of the NLP-callback hot path.

A *step* is one callback evaluation set -- objective + gradient + constraints +
Jacobian values + Hessian-of-Lagrangian values at the same (x, lambda, sigma) --
of BASELINE.json's configs[1]: robot_arm on the LGR transcription, 2000
intervals x 20 points (40 000 collocation nodes, L=360 008, m=240 000,
nnz_J=12 479 754, nnz_H=12 799 740).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* ``value``  device-resident throughput: inputs already in HBM, CUDA events on the
  engine's stream around each step, L2 flushed (untimed) between steps.
* ``e2e``    the same metric through the public API with HOST buffers: every step copies
  x, lambda, sigma to the device and all five results back.  ``e2e.value`` uses the
  single-call set evaluation (System.evaluate -> pk_eval_set: one upload, copies
  overlapped with compute); ``e2e.five_callbacks`` is the same set through the five
  reference-style callbacks called one after the other (x uploaded five times).
* ``roofline`` the dominant kernel (the block expansion of the Hessian) against the
  measured HBM copy bandwidth in MEASURED_PEAKS.json.
* ``cpu_baseline`` / ``--impl reference``: the CPU oracle port (oracle/pockit_oracle.py,
  a restatement of the reference's NumPy algorithm; the reference itself cannot
  travel to the GPU box) timed on the host cores.

N > 1 (torchrun): every rank evaluates its own independent OCP instance of the
same shape (a parameter sweep sharded by instance, no data-path collective);
``value`` = total eval-sets/s over all ranks, time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "NLP callback eval-sets/s (objective+gradient+constraints+Jacobian+Hessian) at 40k nodes"
UNIT = "eval-sets/s"
WORKLOAD = dict(builder="robot_arm", scheme="radau", mesh=2000, num_point=20)
_NAME = "robot_arm LGR 2000x20 (40000 nodes; BASELINE.json configs[1])"
# the other BASELINE.json configurations: only the CPU arm can be pointed at them
# (`bench.py --impl reference --workload humanoid`), so that tools/measure_configs.py gets its CPU
# column from this file's cpu leg instead of touching oracle/ itself
OTHER_WORKLOADS = {
    "lqr": (dict(builder="lqr", scheme="lobatto", mesh=10, num_point=10), "LQR LGL 10x10 (BASELINE.json configs[0])"),
    "humanoid": (dict(builder="humanoid", scheme="lobatto", mesh=11112, num_point=10), "humanoid LGL 11112x10 (configs[2])"),
    "humanoid_small": (dict(builder="humanoid", scheme="lobatto", mesh=1000, num_point=10), "humanoid LGL 1000x10 (configs[2])"),
    "rocket": (dict(builder="rocket", scheme="lobatto", mesh=5556, num_point=10), "two-stage rocket LGL 2x5556x10 (configs[3])"),
    "quadrotor": (dict(builder="quadrotor", scheme="lobatto", mesh=14, num_point=6), "quadrotor LGL 14x6, one instance (configs[4])"),
}


def workload_name():
    return _NAME


def build_system(seed_shift: int = 0):
    import importlib

    from pockit_b200 import problems

    mod = importlib.import_module(f"pockit_b200.{WORKLOAD['scheme']}")
    S = problems.BUILDERS[WORKLOAD["builder"]](mod, mesh=WORKLOAD["mesh"], num_point=WORKLOAD["num_point"])
    x, lam, sigma = problems.evaluation_point(S, seed=1 + seed_shift)
    return S, x, lam, sigma


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
def _cpu_worker(args):
    seed_shift, sets = args
    S, x, lam, sigma = build_system(seed_shift)
    from oracle.pockit_oracle import OracleSystem

    O = OracleSystem(S)

    def one():
        O.objective(x); O.gradient(x); O.constraints(x); O.jacobian(x); O.hessian(x, lam, sigma)

    one()  # warm-up (lambdify caches, page faults)
    t0 = time.perf_counter()
    for _ in range(sets):
        one()
    return (time.perf_counter() - t0) / sets


def cpu_eval_sets_per_s(workers: int, sets: int):
    """Oracle port on `workers` host processes (the path itself is single-threaded)."""
    if workers == 1:
        per = [_cpu_worker((0, sets))]
    else:
        import multiprocessing as mp

        with mp.get_context("spawn").Pool(workers) as pool:
            per = pool.map(_cpu_worker, [(i, sets) for i in range(workers)])
    return sum(1.0 / t for t in per), max(per)


def run_reference(args, rank, world):
    global _NAME
    if rank != 0:
        return
    if args.workload != "robot_arm":
        wl, _NAME = OTHER_WORKLOADS[args.workload]
        WORKLOAD.clear()
        WORKLOAD.update(wl)
    cores = args.cores or max(1, min(os.cpu_count() or 1, 8))
    sets = max(1, min(args.steps, 5))  # bounded sample: ~1 s per set and worker
    value, worst = cpu_eval_sets_per_s(cores, sets)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * worst, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name()},
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sets} eval-set(s) per worker after 1 warm-up on {cores} independent worker processes "
                      f"(oracle/pockit_oracle.py, NumPy restatement of the reference; one process is single-threaded)",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="robot_arm", choices=["robot_arm"] + sorted(OTHER_WORKLOADS),
                    help="CPU arm only (--impl reference): time another BASELINE.json configuration")
    ap.add_argument("--cores", type=int, default=0, help="CPU arm only: worker processes (default: min(8, host cores))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-compact", action="store_true", help="skip the extra de-duplicated-pattern measurement")
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

    S, x, lam, sigma = build_system(seed_shift=rank)
    lo = S.lowering
    eng = Engine(lo, device=local)
    S._engine = eng  # the public callbacks below run on this engine
    S.pinned_outputs = True
    eng.reuse_outputs = True
    modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    for m in modes:
        eng.load(m)
    eng.upload(x, lam, sigma)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------
    eng.time_steps(modes, max(3, args.warmup), flush_l2=True)
    barrier()
    launches0 = eng.launches
    clk = ClockSampler(local)
    clk.__enter__()
    ms = eng.time_steps(modes, args.steps, flush_l2=True)
    eng.sync()
    launches = eng.launches - launches0
    barrier()
    t_local = sum(ms) / 1000.0
    if dist is not None:
        t = torch.tensor([t_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_max = float(t.item())
    else:
        t_max = t_local
    value = world * args.steps / t_max

    # ---- end to end through the public API (host buffers in, host buffers out) -------------
    def one_set():
        S.objective(x); S.gradient(x); S.constraints(x); S.jacobian(x); S.hessian(x, lam, sigma)

    def one_set_call():
        S.evaluate(x, lam, sigma)

    def timed_host(fn):
        for _ in range(max(3, args.warmup)):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        t_loc = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([t_loc], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return t_loc

    e2e_five_t = timed_host(one_set)
    e2e_t = timed_host(one_set_call)
    clk.__exit__()
    e2e_each = {}
    for cname, fn in (("objective", lambda: S.objective(x)), ("gradient", lambda: S.gradient(x)),
                      ("constraints", lambda: S.constraints(x)), ("jacobian", lambda: S.jacobian(x)),
                      ("hessian", lambda: S.hessian(x, lam, sigma))):
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        e2e_each[cname] = 1000.0 * (time.perf_counter() - t0) / 5
    L, m, nj, nh = lo.r_s, lo.m, lo.nnz_jac, lo.nnz_hess_o + lo.nnz_hess_c
    h2d = 8 * (L + m + 1)
    d2h = 8 * (1 + L + m + nj + nh)
    # opt-in de-duplicated patterns (outside the reference's pattern contract; reported beside it)
    compact = None
    if rank == 0 and not args.no_compact:
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

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------
    peaks, peak_source = {}, "fallback (B200_PROFILING.md: 6.65 TB/s)"
    for pk, label in ((ROOT / "MEASURED_PEAKS.json", "MEASURED_PEAKS.json"),
                      (ROOT / "profiles" / "r01_measured_peaks.json",
                       "profiles/r01_measured_peaks.json (copy of the driver's round-1 MEASURED_PEAKS.json)")):
        if pk.exists():
            peaks, peak_source = json.loads(pk.read_text()), label
            break
    peak = float(peaks.get("hbm_gbs", 6650.0))
    iters = 20
    _, stages = eng.time(P.HESS, iters=iters, stages=True)
    fin = eng.fin[P.HESS]
    if len(fin["jobs"][P.ST_EXPAND]):  # table/W path (irregular blocks): the stand-alone expansion kernel dominates
        kernel = f"{eng.expand_kernel(P.HESS)} (Hessian mode)"
        k_ms = stages[P.ST_EXPAND] / iters
        jobs = fin["jobs"][P.ST_EXPAND]
        slots = int(sum(int(j["i"][1]) * int(j["i"][11]) * int(j["i"][4]) for j in jobs))
        rows_read = int(sum(int(j["i"][1]) for j in jobs))
        alg_bytes = 8 * (slots + rows_read * lo.phases[0].L_m + lo.phases[0].col.n_rows * lo.phases[0].n_x)
    else:  # fused path: the generated per-node program evaluates, chains and writes the slots itself
        kernel = "pk_node_hessian_p0 (NVRTC per-node program with fused block expansion)"
        k_ms = stages[6] / iters
        small = int(sum(int(j["i"][1]) for j in fin["jobs"][P.ST_GENERIC]))
        slots = lo.nnz_hess_o + lo.nnz_hess_c - small
        rows_written = len(eng.plan.mode(P.HESS).rows)
        # slots written + x and multipliers read once + node-table rows written
        alg_bytes = 8 * (slots + lo.r_s + lo.m + rows_written * lo.phases[0].L_m)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    tr = ROOT / "profiles" / "r01_expand_traffic_v8.json"
    if tr.exists():  # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture
        name = kernel.split(" ")[0]
        recs = [r for r in json.loads(tr.read_text())["launches"] if r["kernel"].startswith(name)]
        hess = [r for r in recs if "<1>" in r["kernel"]]  # template argument LAM = true: the Hessian launch
        if hess or recs:
            rec = (hess or recs)[-1]
            traffic = rec["dram_read_bytes"] + rec["dram_write_bytes"]
    set_bytes = 8 * (6 * L + 2 * m + nj + nh)
    per_mode = {}
    for mname, mm in zip(("objective", "gradient", "constraints", "jacobian", "hessian"), modes):
        per_mode[mname] = eng.time(mm, iters=iters) / iters
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_name(), "L": int(L), "m": int(m), "nnz_jac": int(nj), "nnz_hess": int(nh),
            "l2": "flushed between steps (256 MiB fill, untimed)", "parallelism": f"{world} independent instance(s), one per GPU",
            "set_algorithmic_MB": set_bytes / 1e6, "set_GBps": set_bytes / (t_max / args.steps) / 1e9,
            "ms_per_callback": per_mode,
        },
        "roofline": {
            "bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_source,
            "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": k_ms,
        },
        "e2e": {"value": world * args.steps / e2e_t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1000.0 * e2e_t / args.steps,
                "api": "System.evaluate(x, lam, sigma) -> pk_eval_set: one upload, per-mode streams, copies overlapped",
                "five_callbacks": {"value": world * args.steps / e2e_five_t, "ms_per_step": 1000.0 * e2e_five_t / args.steps,
                                   "h2d_bytes_per_step": 8 * (5 * L + m + 1), "ms_per_callback": e2e_each,
                                   "api": "System.objective/gradient/constraints/jacobian/hessian one after the other"},
                "outputs": "page-locked engine buffers (System.pinned_outputs=True)"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if compact is not None:
        line["config"]["compact_patterns"] = compact
    if not args.no_cpu_baseline and world == 1:
        v, worst = cpu_eval_sets_per_s(1, 3)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "3 eval-sets of the same workload after 1 warm-up, oracle/pockit_oracle.py on one host core",
        }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
