"""Generate golden vectors from the REAL reference (``/root/reference``).

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py [case ...]

For every case the reference model is built with the shared model text in
``pockit_b200.problems``, evaluated at a seeded synthetic point, and its five
callback outputs plus both COO patterns are stored in ``tests/golden/<case>.npz``.
"""
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference")

CASES = {
    # name: (builder, scheme, kwargs)
    "general_lgl": ("general", "lobatto", {}),
    "general_lgr": ("general", "radau", {}),
    "general_linear_lgl": ("general", "lobatto", {"linear_objective": True}),
    "general_linear_lgr": ("general", "radau", {"linear_objective": True}),
    "general_uniform_lgr": ("general", "radau", {"mesh": 5, "num_point": 4}),
    "lqr_lgl_10x10": ("lqr", "lobatto", {"mesh": 10, "num_point": 10}),
    "lqr_lgr_3x4": ("lqr", "radau", {"mesh": 3, "num_point": 4}),
    "robot_arm_lgr_6x20": ("robot_arm", "radau", {"mesh": 6, "num_point": 20}),
    "robot_arm_lgl_5x4": ("robot_arm", "lobatto", {"mesh": 5, "num_point": 4}),
    "humanoid_lgl_4x5": ("humanoid", "lobatto", {"mesh": 4, "num_point": 5}),
    "rocket_lgl_4x5": ("rocket", "lobatto", {"mesh": 4, "num_point": 5}),
    "rocket_lgr_3x3": ("rocket", "radau", {"mesh": 3, "num_point": 3}),
    "quadrotor_lgl_14x6": ("quadrotor", "lobatto", {"mesh": 14, "num_point": 6}),
    "static_only_lgl": ("static_only", "lobatto", {}),
    "no_control_lgl_3x3": ("no_control", "lobatto", {}),
    "no_control_lgr_3x3": ("no_control", "radau", {}),
    "tiny_lgl_1x3": ("tiny", "lobatto", {"mesh": 1, "num_point": 3}),
    "tiny_lgl_2x2": ("tiny", "lobatto", {"mesh": 2, "num_point": 2}),
    "tiny_lgr_1x2": ("tiny", "radau", {"mesh": 1, "num_point": 2}),
    "quadrotor_lgr_5x3": ("quadrotor", "radau", {"mesh": [0, 0.1, 0.3, 0.6, 0.8, 1.0], "num_point": [3, 4, 3, 5, 2]}),
}


def main(names):
    import importlib

    from pockit_b200 import problems

    for name in names:
        builder, scheme, kw = CASES[name]
        mod = importlib.import_module(f"pockit.{scheme}")
        t0 = time.time()
        S = problems.BUILDERS[builder](mod, **kw)
        x, lam, sigma = problems.evaluation_point(S)
        jr, jc = S.jacobianstructure()
        hr, hc = S.hessianstructure()
        out = dict(
            x=x, lam=lam, sigma=np.float64(sigma),
            objective=np.float64(S.objective(x.copy())),
            gradient=S.gradient(x.copy()),
            constraints=S.constraints(x.copy()),
            jacobian=S.jacobian(x.copy()),
            hessian=S.hessian(x.copy(), lam, sigma),
            hessian_o=S.hessian_o(x.copy()),
            jac_row=np.asarray(jr, np.int64), jac_col=np.asarray(jc, np.int64),
            hess_row=np.asarray(hr, np.int64), hess_col=np.asarray(hc, np.int64),
            v_lb=S.v_lb, v_ub=S.v_ub, c_lb=S.c_lb, c_ub=S.c_ub,
        )
        assert all(np.all(np.isfinite(np.asarray(v))) for k, v in out.items() if k not in ("v_lb", "v_ub", "c_lb", "c_ub")), name
        np.savez_compressed(HERE / f"{name}.npz", **out)
        print(f"{name}: L={len(x)} m={len(lam)} nnzJ={len(jr)} nnzH={len(hr)}  ({time.time() - t0:.1f}s)")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
