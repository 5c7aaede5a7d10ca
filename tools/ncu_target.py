#!/usr/bin/env python
"""Small target for ncu captures: a few evaluation sets of one configuration (L2 flushed between)."""
import importlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import __graft_entry__ as graft

graft.build()
from expand_ab import CONFIGS
from pockit_b200 import plan as P, problems
from pockit_b200.engine import Engine

name = sys.argv[1] if len(sys.argv) > 1 else "robot_arm"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
builder, scheme, kw, B = CONFIGS[name]
S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
x, lam, sigma = problems.evaluation_point(S)
if B > 1:
    import numpy as np
    rng = np.random.default_rng(0)
    x = x[None, :] + 1e-2 * rng.normal(size=(B, len(x)))
    lam = np.tile(lam, (B, 1))
    sigma = np.full(B, sigma)
eng = Engine(S.lowering, batch=B, fastmath=S._fastmath)
eng.upload(x, lam, sigma)
eng.time_steps([P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS], steps, flush_l2=True)
eng.sync()
print("done", eng.launches)
