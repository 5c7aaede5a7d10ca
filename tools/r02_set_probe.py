#!/usr/bin/env python
"""Which callbacks overlap inside an evaluation set?  Whole-set device time (graph replay, L2 flushed)
for subsets of the five modes, under the scheduling switches given as VAR=a,b:
    python tools/r02_set_probe.py [config] [VAR=a,b ...]"""
import importlib
import json
import os
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def main():
    import __graft_entry__ as graft

    graft.build()
    from expand_ab import CONFIGS
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.engine import Engine

    args = sys.argv[1:]
    name = args.pop(0) if args and "=" not in args[0] else "robot_arm"
    sweeps = [(a.split("=")[0], a.split("=")[1].split(",")) for a in args] or [("POCKIT_B200_GRAPH_PRIORITY", ["1"])]
    builder, scheme, kw, B = CONFIGS[name]
    S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
    x, lam, sigma = problems.evaluation_point(S)
    subsets = [
        ("jac+hess", [P.JAC, P.HESS]), ("jac+hess+obj", [P.JAC, P.HESS, P.OBJ]), ("jac+hess+cons", [P.JAC, P.HESS, P.CONS]),
        ("jac+hess+grad", [P.JAC, P.HESS, P.GRAD]), ("all5", [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]),
        ("obj", [P.OBJ]), ("grad", [P.GRAD]), ("cons", [P.CONS]), ("small3", [P.OBJ, P.GRAD, P.CONS]), ("jac", [P.JAC]), ("hess", [P.HESS]),
    ]
    for var, values in sweeps:
        for val in values:
            os.environ[var] = val
            eng = Engine(S.lowering, fastmath=S._fastmath)
            eng.upload(x, lam, sigma)
            rec = {"config": name, var: val}
            for label, modes in subsets:
                eng.time_steps(modes, 5, flush_l2=True)
                ms = eng.time_steps(modes, 40, flush_l2=True)
                rec[label] = round(1000 * statistics.median(ms), 2)
            print(json.dumps(rec), flush=True)
            eng.close()
        os.environ.pop(var, None)


if __name__ == "__main__":
    main()
