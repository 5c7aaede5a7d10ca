#!/usr/bin/env python
"""End-to-end time of the continuous error-estimate data on the device (System.error_estimation_data,
host x in, host arrays out) for BASELINE configs[1] and the 100 k-node meshes.  One JSON line each."""
import importlib
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def main():
    import __graft_entry__ as graft

    graft.build()
    from expand_ab import CONFIGS
    from pockit_b200 import problems

    for name in sys.argv[1:] or ["robot_arm", "humanoid", "rocket"]:
        builder, scheme, kw, _ = CONFIGS[name]
        S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
        x, _, _ = problems.evaluation_point(S)
        t0 = time.perf_counter()
        S.error_estimation_data(x)  # plans the augmented operators, NVRTC, uploads
        t_first = time.perf_counter() - t0
        for _ in range(3):
            S.error_estimation_data(x)
        t0 = time.perf_counter()
        for _ in range(20):
            out = S.error_estimation_data(x)
        dt = (time.perf_counter() - t0) / 20
        print(json.dumps({"config": name, "first_call_s": round(t_first, 2), "ms_per_call": round(1000 * dt, 3),
                          "values": int(sum(2 * T.size for T, _ in out))}), flush=True)


if __name__ == "__main__":
    main()
