#!/usr/bin/env python
"""Per-mode, per-stage device times (pk_time: CUDA events around back-to-back launches of one stage
at a time, so each figure includes the launch / drain gap of a stand-alone kernel).
    python tools/stage_times.py [robot_arm|humanoid|rocket|quadrotor ...]"""
import sys, importlib, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import __graft_entry__ as graft
graft.build()
from expand_ab import CONFIGS
from pockit_b200 import plan as P, problems
from pockit_b200.engine import Engine
names = ["reduce","defect","generic","expand","grad","-","node","sys"]
for name in sys.argv[1:] or ["humanoid"]:
    builder, scheme, kw, B = CONFIGS[name]
    S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
    x, lam, sigma = problems.evaluation_point(S)
    if B > 1:  # batched configuration: B instances around the evaluation point
        import numpy as np
        rng = np.random.default_rng(0)
        x = x[None, :] + 1e-2 * rng.normal(size=(B, len(x)))
        lam = np.tile(lam, (B, 1))
        sigma = np.full(B, sigma)
    eng = Engine(S.lowering, batch=B, fastmath=S._fastmath); eng.upload(x, lam, sigma)
    for m in eng._mode_ids:
        eng.time(m, iters=5)
        tot, st = eng.time(m, iters=20, stages=True)
        f = eng.fin[m]
        print(name, P.MODES[m], "mode us %.1f" % (1000*tot/20), {k: round(1000*v/20, 1) for k, v in zip(names, st) if v > 0.0005},
              "jobs", {s: len(f["jobs"][s]) for s in range(6) if len(f["jobs"][s])}, "rows", len(eng.plan.mode(m).rows), flush=True)
