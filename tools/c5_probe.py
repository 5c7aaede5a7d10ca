#!/usr/bin/env python
"""BASELINE configs[4] on one GPU: device-resident set time of the 8192-instance quadrotor batch
(bench.py's c5_sharded at N = 1), for A/B runs of the batch kernels.   python tools/c5_probe.py [VAR=a,b]"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch

    import __graft_entry__ as graft

    graft.build()
    import bench
    from pockit_b200 import plan as P

    peak, _ = bench.load_peaks()
    sweeps = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[1:]] or [("POCKIT_B200_EXPAND", [""])]
    for var, values in sweeps:
        for val in values:
            if val:
                os.environ[var] = val
            else:
                os.environ.pop(var, None)
            rec = bench.c5_sharded(P, peak, 20, 0, 1, 0, None, torch)
            print(json.dumps({var: val, **{k: rec[k] for k in ("device_ms_per_batch_set", "set_roofline_frac_per_gpu", "e2e_ms_per_batch_set")}}), flush=True)
        os.environ.pop(var, None)


if __name__ == "__main__":
    main()
