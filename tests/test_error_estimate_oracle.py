"""CPU tier: the oracle's restatement of the continuous error-estimate data
(``phasebase.py:1339-1366``) against golden vectors of the real reference, and the augmented-mesh
operators it is built from."""
import sys

import numpy as np
import pytest

from helpers import GOLDEN, assert_close

sys.path.insert(0, str(GOLDEN))
from make_error_golden import CASES  # noqa: E402


def build(case):
    import importlib

    from pockit_b200 import problems

    builder, scheme, kw = CASES[case]
    return problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_error_data_matches_reference(case):
    from oracle.pockit_oracle import OracleSystem

    S = build(case)
    g = np.load(GOLDEN / f"error_{case}.npz")
    x_in = g["x"].copy()
    got = OracleSystem(S).error_estimation_data(g["x"])
    assert np.array_equal(g["x"], x_in)
    assert len(got) == len(S.p)
    for i, (T, I) in enumerate(got):
        assert_close(T, g[f"T_{i}"], f"T_x_aug[{i}]")
        assert_close(I, g[f"I_{i}"], f"I_f_aug[{i}]")


def test_augmented_operators_shapes():
    from pockit_b200.discretization import AugmentedCollocation, Collocation

    for scheme in ("lgl", "lgr"):
        col = Collocation(scheme, np.array([0.0, 0.2, 0.5, 1.0]), np.array([3, 5, 4]), 2, 1)
        A = AugmentedCollocation(col)
        assert A.V.shape == (3 * A.L_m, col.L_xu) and A.T.shape == (2 * A.rows, 2 * col.L_x) and A.I.shape == (A.rows, A.L_m)
        # interpolating a constant gives the constant; its translation is zero
        ones = np.ones(col.L_xu)
        np.testing.assert_allclose(A.V.dot(ones), 1.0, atol=1e-9)
        np.testing.assert_allclose(A.T.dot(np.ones(2 * col.L_x)), 0.0, atol=1e-9)


@pytest.mark.parametrize("case", sorted(CASES))
def test_generated_error_programs_match_reference(case):
    """The planner's prep / node programs (compiled as C++ on the host) + the CSR operators reproduce
    the reference's error-estimate data; the same programs run through NVRTC on the GPU."""
    sys.path.insert(0, str(GOLDEN.parent))
    from hostemu import emulate_error_data

    S = build(case)
    g = np.load(GOLDEN / f"error_{case}.npz")
    got = emulate_error_data(S, g["x"])
    for i, (T, I) in enumerate(got):
        assert_close(T, g[f"T_{i}"], f"T_x_aug[{i}]")
        assert_close(I, g[f"I_{i}"], f"I_f_aug[{i}]")


@pytest.mark.parametrize("scheme", ["lobatto", "radau"])
def test_check_continuous_known_answers_on_the_host_emulated_programs(scheme):
    """The reference's known-answer test of the continuous check
    (tests/test_labatto/test_check_lobatto.py:22-36): exact polynomial trajectories pass, perturbed
    ones fail -- here with the generated programs emulated on the host feeding the interval test."""
    import importlib

    sys.path.insert(0, str(GOLDEN.parent))
    from hostemu import emulate_error_data
    from pockit_b200 import problems
    from pockit_b200.optimizer._common import pack_guess
    from pockit_b200.system import continuous_error_intervals

    S, cases = problems.check_system(importlib.import_module(f"pockit_b200.{scheme}"))
    for value, expected in cases:
        x, _, _ = pack_guess(S, value, None)
        (T, I), = emulate_error_data(S, x)
        ok = continuous_error_intervals(S.p[0], T, I, 1e-8, 1e-8, 1e-4)
        assert bool(ok.all()) is expected
