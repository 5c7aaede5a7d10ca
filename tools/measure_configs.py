#!/usr/bin/env python
"""Measure the five BASELINE.json configurations on one GPU (plus, through bench.py's CPU arm, the
reference / the CPU port on a bounded sample) and print one JSON line per configuration.  Not the bench contract --
bench.py is -- this fills the table in DESIGN.md / README.md.

    python tools/measure_configs.py [--skip-cpu] [name ...]
"""
import argparse
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CONFIGS = {
    "C1_lqr_lgl_10x10": ("lqr", "lobatto", dict(mesh=10, num_point=10), 1),
    "C2_robot_arm_lgr_2000x20": ("robot_arm", "radau", dict(mesh=2000, num_point=20), 1),
    "C3_humanoid_lgl_1000x10": ("humanoid", "lobatto", dict(mesh=1000, num_point=10), 1),
    "C3_humanoid_lgl_11112x10": ("humanoid", "lobatto", dict(mesh=11112, num_point=10), 1),
    "C4_rocket_lgl_2x5556x10": ("rocket", "lobatto", dict(mesh=5556, num_point=10), 1),
    "C5_quadrotor_lgl_14x6_B8192": ("quadrotor", "lobatto", dict(mesh=14, num_point=6), 8192),
}


CPU_WORKLOAD = {
    "C1_lqr_lgl_10x10": "lqr", "C2_robot_arm_lgr_2000x20": "robot_arm", "C3_humanoid_lgl_1000x10": "humanoid_small",
    "C3_humanoid_lgl_11112x10": "humanoid", "C4_rocket_lgl_2x5556x10": "rocket", "C5_quadrotor_lgl_14x6_B8192": "quadrotor",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--detail", action="store_true")
    args = ap.parse_args()
    import __graft_entry__ as graft

    graft.build()
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.batched import fixed_index, fixed_table
    from pockit_b200.engine import Engine

    modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    for name in args.names or list(CONFIGS):
        builder, scheme, kw, B = CONFIGS[name]
        mod = importlib.import_module(f"pockit_b200.{scheme}")
        t0 = time.perf_counter()
        S = problems.BUILDERS[builder](mod, **kw)
        lo = S.lowering
        t_plan = time.perf_counter() - t0
        x, lam, sigma = problems.evaluation_point(S)
        fixed = None
        if B > 1:
            rng = np.random.default_rng(0)
            fixed = fixed_table(S, B)
            fixed[:, fixed_index(S, 0, "x0", 0)] += rng.uniform(-0.2, 0.2, B)
            fixed[:, fixed_index(S, 0, "x0", 1)] += rng.uniform(-0.2, 0.2, B)
            X = x[None, :] + 1e-2 * rng.normal(size=(B, len(x)))
            LAM = lam[None, :] + 0.1 * rng.normal(size=(B, len(lam)))
        else:
            X, LAM = x, lam
        t0 = time.perf_counter()
        eng = Engine(lo, batch=B, fastmath=S._fastmath, fixed=fixed)
        for m in modes:
            eng.load(m)
        t_jit = time.perf_counter() - t0
        eng.reuse_outputs = True
        eng.upload(X, LAM, np.full(B, sigma))
        eng.time_steps(modes, 5, True)
        ms = eng.time_steps(modes, args.steps, True)
        dev = B * len(ms) / (sum(ms) / 1e3)
        # end to end with host buffers
        def one_set():
            eng.objective(X); eng.gradient(X); eng.constraints(X); eng.jacobian(X); eng.hessian(X, LAM, np.full(B, sigma))
        for _ in range(3):
            one_set()
        reps = max(3, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(reps):
            one_set()
        e2e = B * reps / (time.perf_counter() - t0)
        # ... and as one engine call per set (pk_eval_set: one upload, copies overlapped)
        sig_all = np.full(B, sigma)
        for _ in range(3):
            eng.evaluate(X, LAM, sig_all, modes=modes)
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.evaluate(X, LAM, sig_all, modes=modes)
        e2e_set = B * reps / (time.perf_counter() - t0)
        L, m, nj, nh = lo.r_s, lo.m, lo.nnz_jac, lo.nnz_hess_o + lo.nnz_hess_c
        set_bytes = 8 * (6 * L + 2 * m + nj + nh) * B
        line = dict(
            config=name, batch=B, nodes=int(sum(p.L_m for p in lo.phases)), L=int(L), m=int(m), nnz_jac=int(nj), nnz_hess=int(nh),
            plan_s=round(t_plan, 2), jit_s=round(t_jit, 2), device_sets_per_s=dev, device_ms_per_set=1e3 * B / dev,
            device_GBps=set_bytes / (sum(ms) / len(ms) / 1e3) / 1e9, e2e_sets_per_s=e2e_set, e2e_five_callbacks_sets_per_s=e2e,
        )
        if args.detail:
            names = ["reduce", "defect", "generic", "expand", "grad_range", "grad_scalar", "node", "system"]
            line["ms_per_callback"] = {}
            for mname, mm in zip(P.MODES, range(5)):
                tot, st = eng.time(mm, iters=10, stages=True)
                line["ms_per_callback"][mname] = dict(total=tot / 10, **{k: v / 10 for k, v in zip(names, st) if v > 0})
        if not args.skip_cpu and name in CPU_WORKLOAD:
            # the CPU column comes from bench.py's cpu leg (the only non-test code that runs oracle/)
            import subprocess

            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", CPU_WORKLOAD[name],
                                "--cores", "1", "--steps", "3", "--warmup", "1"], capture_output=True, text=True)
            if r.returncode == 0 and r.stdout.strip():
                cpu = json.loads(r.stdout.strip().splitlines()[-1])
                line["cpu_sets_per_s_1core"] = cpu["value"]
                line["cpu_kind"] = cpu["cpu_baseline"]["kind"]  # "reference" (oracle/_ref staged) or "port"
        print(json.dumps(line), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
