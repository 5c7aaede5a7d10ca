#!/usr/bin/env python
"""A/B of the two block-expansion kernels (POCKIT_B200_EXPAND=columns|params|bulk) on one GPU:
stage times of the Jacobian / Hessian expansion, whole-set time, and bit-equality of the outputs.

    python tools/expand_ab.py [config ...]     configs: robot_arm humanoid rocket quadrotor
"""
import importlib
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CONFIGS = {
    "robot_arm": ("robot_arm", "radau", dict(mesh=2000, num_point=20), 1),
    "humanoid": ("humanoid", "lobatto", dict(mesh=11112, num_point=10), 1),
    "rocket": ("rocket", "lobatto", dict(mesh=5556, num_point=10), 1),
    "quadrotor": ("quadrotor", "lobatto", dict(mesh=14, num_point=6), 8192),
}


def main():
    import __graft_entry__ as graft

    graft.build()
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.engine import Engine

    modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    for name in sys.argv[1:] or ["robot_arm"]:
        builder, scheme, kw, B = CONFIGS[name]
        S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
        x, lam, sigma = problems.evaluation_point(S)
        if B > 1:
            rng = np.random.default_rng(0)
            x = x[None, :] + 1e-2 * rng.normal(size=(B, len(x)))
            lam = np.tile(lam, (B, 1))
        outs = {}
        for variant in ("columns", "params", "bulk"):
            os.environ["POCKIT_B200_EXPAND"] = variant
            eng = Engine(S.lowering, batch=B, fastmath=S._fastmath)
            try:
                for m in modes:
                    eng.load(m)
            except RuntimeError as exc:  # this configuration does not fit the variant
                print(json.dumps({"config": name, "variant": variant, "skipped": str(exc).splitlines()[0][:160]}), flush=True)
                eng.close()
                continue
            eng.upload(x, lam, sigma)
            rec = {"config": name, "variant": variant}
            for m, tag in ((P.JAC, "jac"), (P.HESS, "hess")):
                eng.time(m, iters=5)
                tot, st = eng.time(m, iters=20, stages=True)
                rec[f"{tag}_mode_us"] = 1000 * tot / 20
                rec[f"{tag}_expand_us"] = 1000 * st[P.ST_EXPAND] / 20
                rec[f"{tag}_node_us"] = 1000 * st[6] / 20
                rec[f"{tag}_generic_us"] = 1000 * st[P.ST_GENERIC] / 20
            eng.time_steps(modes, 5, flush_l2=True)
            ms = eng.time_steps(modes, 50, flush_l2=True)
            rec["set_us_flushed"] = 1000 * sum(ms) / len(ms)
            ms = eng.time_steps(modes, 50, flush_l2=False)
            rec["set_us_warm"] = 1000 * sum(ms) / len(ms)
            eng.run(P.JAC); eng.run(P.HESS); eng.sync()
            outs[variant] = (eng.download(P.JAC).copy(), eng.download(P.HESS).copy())
            print(json.dumps(rec), flush=True)
            eng.close()
        same = all(np.array_equal(a, b) for v in outs if v != "columns" for a, b in zip(outs["columns"], outs[v]))
        print(json.dumps({"config": name, "all_variants_bit_identical": bool(same)}), flush=True)


if __name__ == "__main__":
    main()
