#!/usr/bin/env python
"""A/B of set-level scheduling switches (environment variables read when a set is captured):
whole-set device time, L2 flushed, for subsets of modes.   python tools/set_ab.py [config] VAR=a,b ..."""
import importlib
import json
import os
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
    sweeps = [(a.split("=")[0], a.split("=")[1].split(",")) for a in args] or [("POCKIT_B200_CHAIN", ["0", "1"])]
    builder, scheme, kw, B = CONFIGS[name]
    S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
    x, lam, sigma = problems.evaluation_point(S)
    all5 = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    for var, values in sweeps:
        for val in values:
            os.environ[var] = val
            eng = Engine(S.lowering, fastmath=S._fastmath)
            eng.upload(x, lam, sigma)
            rec = {"config": name, var: val}
            for label, modes in (("all5", all5), ("jac+hess", [P.JAC, P.HESS]), ("jac", [P.JAC]), ("hess", [P.HESS]), ("small3", [P.OBJ, P.GRAD, P.CONS])):
                eng.time_steps(modes, 5, flush_l2=True)
                ms = sorted(eng.time_steps(modes, 60, flush_l2=True))
                rec[label + "_us"] = round(1000 * sum(ms) / len(ms), 2)
                rec[label + "_us_median"] = round(1000 * ms[len(ms) // 2], 2)
            print(json.dumps(rec), flush=True)
            eng.close()
        os.environ.pop(var, None)


if __name__ == "__main__":
    main()
