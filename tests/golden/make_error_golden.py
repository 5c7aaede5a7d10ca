"""Golden vectors of the continuous error-estimate data from the REAL reference
(``PhaseBase._error_estimation_data_continuous``, ``phasebase.py:1355-1366``).

Run in the build container only:   python tests/golden/make_error_golden.py

For every case: the reference model, the seeded evaluation point of ``problems.evaluation_point``,
and per phase ``T_x_aug`` / ``I_f_aug``; stored in ``tests/golden/error_<case>.npz``.
"""
import importlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference")

CASES = {
    "robot_arm_lgr_6x20": ("robot_arm", "radau", {"mesh": 6, "num_point": 20}),
    "robot_arm_lgl_5x4": ("robot_arm", "lobatto", {"mesh": 5, "num_point": 4}),
    "rocket_lgl_4x5": ("rocket", "lobatto", {"mesh": 4, "num_point": 5}),
    "rocket_lgr_3x3": ("rocket", "radau", {"mesh": 3, "num_point": 3}),
    "quadrotor_lgr_5x3": ("quadrotor", "radau", {"mesh": [0, 0.1, 0.3, 0.6, 0.8, 1.0], "num_point": [3, 4, 3, 5, 2]}),
    "general_lgl": ("general", "lobatto", {}),
    "general_lgr": ("general", "radau", {}),
    "humanoid_lgl_4x5": ("humanoid", "lobatto", {"mesh": 4, "num_point": 5}),
    "tiny_lgl_1x3": ("tiny", "lobatto", {"mesh": 1, "num_point": 3}),
    "tiny_lgr_1x2": ("tiny", "radau", {"mesh": 1, "num_point": 2}),
    # hp-style meshes: mixed orders and widths
    "rocket_lgl_hp": ("rocket", "lobatto", {"mesh": [0, 0.05, 0.2, 0.45, 0.5, 1.0], "num_point": [3, 6, 4, 2, 9]}),
    "robot_arm_lgr_hp": ("robot_arm", "radau", {"mesh": [0, 0.3, 0.35, 0.7, 1.0], "num_point": [5, 2, 8, 3]}),
}


def main(names):
    from pockit_b200 import problems

    for name in names:
        builder, scheme, kw = CASES[name]
        S = problems.BUILDERS[builder](importlib.import_module(f"pockit.{scheme}"), **kw)
        x, _, _ = problems.evaluation_point(S)
        out = dict(x=x)
        s = x[S.l_s : S.r_s]
        for i, p in enumerate(S.p):
            xp = x[S.l_p[i] : S.r_p[i]].copy()  # the reference substitutes boundary values in place
            T, I = p._error_estimation_data_continuous(xp, s)
            out[f"T_{i}"], out[f"I_{i}"] = np.asarray(T), np.asarray(I)
        np.savez_compressed(HERE / f"error_{name}.npz", **out)
        print(name, [out[f"T_{i}"].shape for i in range(len(S.p))])


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
