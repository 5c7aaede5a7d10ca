"""Shared helpers for the parity tests."""
import importlib
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
RTOL, ATOL = 1e-12, 1e-14  # the north-star tolerance: values within 1e-12 relative / 1e-14 absolute


def golden_cases():
    import sys

    sys.path.insert(0, str(GOLDEN))
    from make_golden import CASES

    return {k: v for k, v in CASES.items() if (GOLDEN / f"{k}.npz").exists()}


def build(case, package="pockit_b200"):
    from pockit_b200 import problems

    builder, scheme, kw = golden_cases()[case]
    mod = importlib.import_module(f"{package}.{scheme}")
    return problems.BUILDERS[builder](mod, **kw)


def load(case):
    return np.load(GOLDEN / f"{case}.npz")


def assert_close(got, want, what="", rtol=RTOL, atol=ATOL):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    err = np.abs(got - want)
    tol = atol + rtol * np.abs(want)
    bad = err > tol
    if bad.any():
        k = int(np.argmax(err - tol))
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.size} outside 1e-12 rel / 1e-14 abs; "
            f"worst at {k}: got {got.flat[k]!r} want {want.flat[k]!r} (|diff| {err.flat[k]:.3e})"
        )
