"""GPU tests written after the round's GPU budget was spent: they exercise code whose pieces were
verified separately (host-emulated programs on the CPU tier, the engine paths by the earlier GPU
tests) but could not be run on a B200 themselves this round.  The file sorts last on purpose, so a
surprise here cannot hide the rest of the suite behind ``pytest -x``."""
import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scheme", ["lobatto", "radau"])
def test_check_continuous_known_answers(scheme):
    """The reference's own known-answer test (tests/test_labatto/test_check_lobatto.py:22-36 and the
    radau twin): polynomial trajectories that satisfy x' = u exactly pass, perturbed ones fail."""
    import importlib

    from pockit_b200 import problems

    mod = importlib.import_module(f"pockit_b200.{scheme}")
    S, cases = problems.check_system(mod)
    for value, expected in cases:
        assert S.check_continuous(value) is expected
    with pytest.raises(NotImplementedError):
        S.check_discontinuous(cases[0][0])
    with pytest.raises(ValueError, match="len\\(value\\)"):
        S.check_continuous([cases[0][0][0]])


def test_remeshing_hits_the_cubin_cache():
    """set_discretization + System.update() builds a new engine whose generated programs are the same
    text: the library's process-wide NVRTC cache serves them, and the values follow the new mesh."""
    import pockit_b200.radau as rad
    from oracle.pockit_oracle import OracleSystem
    from pockit_b200 import problems
    from pockit_b200.engine import cubin_cache_stats

    S = problems.robot_arm(rad, mesh=10, num_point=6)
    x, lam, sigma = problems.evaluation_point(S, seed=2)
    S.jacobian(x), S.hessian(x, lam, sigma)
    h0, m0 = cubin_cache_stats()
    S.p[0].set_discretization(14, 9)
    S.update()
    x, lam, sigma = problems.evaluation_point(S, seed=2)
    O = OracleSystem(S)
    assert_close(S.jacobian(x), O.jacobian(x), "jacobian after re-meshing")
    assert_close(S.hessian(x, lam, sigma), O.hessian(x, lam, sigma), "hessian after re-meshing")
    h1, m1 = cubin_cache_stats()
    assert h1 - h0 >= 2 and m1 == m0
