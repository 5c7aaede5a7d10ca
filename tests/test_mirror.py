"""CPU tier (build container only: needs the reference importable): ``pockit_b200.mirror`` turns a
LIVE reference system into its engine-backed twin with identical layout, bounds and COO patterns,
and the twin's plan (host-emulated) reproduces the reference's callbacks at the same point --
parity against the running reference, not only against stored vectors."""
import importlib
import sys

import numpy as np
import pytest

from helpers import assert_close
from hostemu import HostEmu
from pockit_b200 import plan as P
from pockit_b200 import problems

REF = "/root/reference"


def _reference(scheme):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    try:
        return importlib.import_module(f"pockit.{scheme}")
    except Exception:  # not in this environment (e.g. the GPU box)
        pytest.skip("the reference package is not importable here")


import os

# the reference JIT-compiles every function with Numba (tens of seconds per model): two models by
# default, the rest with POCKIT_B200_SLOW_TESTS=1
CASES = [
    ("general", "lobatto", {}),
    ("lqr", "radau", {"mesh": 3, "num_point": 4}),
]
if os.environ.get("POCKIT_B200_SLOW_TESTS") == "1":
    CASES += [
        ("rocket", "lobatto", {"mesh": 3, "num_point": 4}),
        ("general", "radau", {}),
        ("robot_arm", "radau", {"mesh": 4, "num_point": 5}),
        ("quadrotor", "lobatto", {"mesh": 4, "num_point": 4}),
    ]


@pytest.mark.parametrize("builder,scheme,kw", CASES)
def test_twin_of_a_live_reference_system(builder, scheme, kw):
    from pockit_b200.mirror import from_reference

    ref = problems.BUILDERS[builder](_reference(scheme), **kw)
    twin = from_reference(ref)
    assert type(twin).__module__ == f"pockit_b200.{scheme}"
    assert twin.L == ref.L and twin.n_s == ref.n_s and twin.n_p == ref.n_p
    for name in ("v_lb", "v_ub", "c_lb", "c_ub"):
        assert np.array_equal(getattr(twin, name), getattr(ref, name)), name
    for a, b in zip(twin.jacobianstructure() + twin.hessianstructure(), ref.jacobianstructure() + ref.hessianstructure()):
        assert np.array_equal(a, b)
    x, lam, sigma = problems.evaluation_point(twin, seed=4)
    E = HostEmu(twin)
    assert_close(E.run(P.OBJ, x)[0], ref.objective(x.copy()), "objective")
    assert_close(E.run(P.GRAD, x), ref.gradient(x.copy()), "gradient")
    assert_close(E.run(P.CONS, x), ref.constraints(x.copy()), "constraints")
    assert_close(E.run(P.JAC, x), ref.jacobian(x.copy()), "jacobian")
    assert_close(E.run(P.HESS, x, lam, sigma), ref.hessian(x.copy(), lam, sigma), "hessian")


def test_mirror_rejects_an_unfinished_system():
    from pockit_b200.mirror import from_reference

    ref_mod = _reference("lobatto")
    S = ref_mod.System(0)
    with pytest.raises(ValueError, match="not fully configured"):
        from_reference(S)
