"""GPU: continuous error-estimate data (phasebase.py:1339-1366) computed by the engine
against golden vectors of the real reference and, at a larger size, the CPU oracle."""
import sys

import numpy as np
import pytest

from helpers import GOLDEN, assert_close

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(GOLDEN))
from make_error_golden import CASES  # noqa: E402


def build(case):
    import importlib

    from pockit_b200 import problems

    builder, scheme, kw = CASES[case]
    return problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)


@pytest.mark.parametrize("case", sorted(CASES))
def test_error_data_matches_reference_golden(case):
    S = build(case)
    g = np.load(GOLDEN / f"error_{case}.npz")
    x = g["x"].copy()
    got = S.error_estimation_data(x)
    assert np.array_equal(x, g["x"])  # the engine never writes boundary values into the caller's x
    assert len(got) == len(S.p)
    for i, (T, I) in enumerate(got):
        assert_close(T, g[f"T_{i}"], f"T_x_aug[{i}]")
        assert_close(I, g[f"I_{i}"], f"I_f_aug[{i}]")
    # callbacks still work on the same engine afterwards, and the data is reproducible
    S.jacobian(x)
    again = S.error_estimation_data(x)
    assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(got, again))


@pytest.mark.parametrize("scheme,mesh,n", [("radau", 300, 12), ("lobatto", 400, 7)])
def test_error_data_matches_oracle_on_a_fine_mesh(scheme, mesh, n):
    import importlib

    from oracle.pockit_oracle import OracleSystem
    from pockit_b200 import problems

    S = problems.robot_arm(importlib.import_module(f"pockit_b200.{scheme}"), mesh=mesh, num_point=n)
    x, _, _ = problems.evaluation_point(S, seed=3)
    got, want = S.error_estimation_data(x), OracleSystem(S).error_estimation_data(x)
    for (T, I), (Tw, Iw) in zip(got, want):
        assert_close(T, Tw, "T_x_aug")
        assert_close(I, Iw, "I_f_aug")
    ok = S.check_continuous_intervals(x)
    assert len(ok) == 1 and ok[0].shape == (mesh,)
    # a random point is not a solution: the check must fail somewhere; with huge tolerances it passes
    assert not ok[0].all() and not S.check_continuous(x)
    assert S.check_continuous(x, 1e9, 1e9)
