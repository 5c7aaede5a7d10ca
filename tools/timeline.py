#!/usr/bin/env python
"""Device-side timeline of one evaluation set (pk_timeline): when every kernel of every mode
starts / ends relative to the release of the set.   python tools/timeline.py [robot_arm|humanoid|rocket]
Caveat (measured): every timing event costs ~3 us of front-end time, and the marks of concurrent
streams get serialised, so the ORDER of the marks is meaningful, the distances between them are not."""
import importlib
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

    name = sys.argv[1] if len(sys.argv) > 1 else "robot_arm"
    builder, scheme, kw, B = CONFIGS[name]
    S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
    x, lam, sigma = problems.evaluation_point(S)
    eng = Engine(S.lowering, fastmath=S._fastmath)
    modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    eng.upload(x, lam, sigma)
    eng.time_steps(modes, 3)
    tags = {0: "reduce", 1: "defect", 2: "generic", 3: "expand", 4: "grad", 5: "grad", 6: "node", 7: "sys", 8: "compact"}
    for rep in range(3):
        rows = eng.timeline(modes)
    print(f"# {name}: timeline of the last of 3 sets (us since release)")
    for mode, tag, edge, us in sorted(rows, key=lambda r: r[3]):
        who = "set" if mode < 0 else P.MODES[mode]
        print(f"{us:8.2f}  {who:12s} {tags.get(tag, 'end'):8s} {'start' if edge == 0 else 'done'}")
    g = eng.time_steps(modes, 20, flush_l2=True)
    print(f"# graph replay, flushed: {1000 * sum(g) / len(g):.2f} us per set")


if __name__ == "__main__":
    main()
