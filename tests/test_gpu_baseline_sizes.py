"""GPU parity at the BASELINE.json sizes: all five callbacks of every expansion-kernel variant
against the CPU oracle (values within 1e-12 relative / 1e-14 absolute, patterns by construction
those the oracle was pinned on).

    C2 robot_arm LGR 2000x20      40 000 nodes   (the bench configuration)
    C3 humanoid  LGL 11112x10    100 009 nodes
    C4 rocket    LGL 2x5556x10   100 010 nodes   (two phases, FUNC-linked boundaries)
    C5 quadrotor LGL 14x6, B = 8192 instances differing in the FIXED initial state

The block expansion (``phasebase.py:1120-1124, 1280-1285``) has five kernels (the two for batches,
pk_expand_slots and pk_expand_batch, are exercised by the batched test); every test asserts
that the engine really launches the one it claims to test (``Engine.expand_kernel``):
``columns`` -> pk_expand_blocks (persistent), ``params`` -> pk_expand_cols (the default where a
block is a whole number of 32-byte sectors: C2, the kernel the bench and the roofline figure run
on), ``bulk`` -> pk_expand_bulk (TMA bulk stores; the default for the 9 x 10 blocks of C3 / C4).
"""
import importlib

import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu

CONFIGS = {
    "robot_arm": ("robot_arm", "radau", dict(mesh=2000, num_point=20)),
    "humanoid": ("humanoid", "lobatto", dict(mesh=11112, num_point=10)),
    "rocket": ("rocket", "lobatto", dict(mesh=5556, num_point=10)),
}
KERNEL = {"columns": "pk_expand_blocks", "params": "pk_expand_cols", "bulk": "pk_expand_bulk"}
_cache = {}


def _reference(name):
    """System, evaluation point and the oracle's five outputs (computed once per configuration)."""
    if name not in _cache:
        from oracle.pockit_oracle import OracleSystem
        from pockit_b200 import problems

        builder, scheme, kw = CONFIGS[name]
        S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
        x, lam, sigma = problems.evaluation_point(S, seed=11)
        sigma = 0.8
        O = OracleSystem(S)
        want = {
            "objective": O.objective(x), "gradient": O.gradient(x), "constraints": O.constraints(x),
            "jacobian": O.jacobian(x), "hessian": O.hessian(x, lam, sigma),
        }
        _cache[name] = (S, x, lam, sigma, want)
    return _cache[name]


@pytest.mark.parametrize("variant", ["params", "columns", "bulk", "default"])
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_full_size_callbacks_match_oracle(name, variant, monkeypatch):
    from pockit_b200 import plan as P
    from pockit_b200.engine import Engine

    S, x, lam, sigma, want = _reference(name)
    if variant == "default":
        monkeypatch.delenv("POCKIT_B200_EXPAND", raising=False)
    else:
        monkeypatch.setenv("POCKIT_B200_EXPAND", variant)
    eng = Engine(S.lowering, fastmath=S._fastmath)
    try:
        # default at these sizes: the parameter-driven column walk for sector-aligned blocks (robot_arm: 20 x 20),
        # its TMA bulk-store variant where a block is not a whole number of sectors (LGL n = 10: 9 x 10)
        expect = KERNEL.get(variant, "pk_expand_cols" if name == "robot_arm" else "pk_expand_bulk")
        assert eng.expand_kernel(P.JAC) == expect and eng.expand_kernel(P.HESS) == expect
        x_in = x.copy()
        assert_close(eng.objective(x), want["objective"], "objective")
        assert_close(eng.gradient(x), want["gradient"], "gradient")
        assert_close(eng.constraints(x), want["constraints"], "constraints")
        assert_close(eng.jacobian(x), want["jacobian"], "jacobian")
        assert_close(eng.hessian(x, lam, sigma), want["hessian"], "hessian")
        assert np.array_equal(x, x_in)
        # the whole set in one call (pk_eval_set: the bench's e2e path) gives the same values
        r = eng.evaluate(x, lam, sigma)
        for mode, key in ((P.OBJ, "objective"), (P.GRAD, "gradient"), (P.CONS, "constraints"), (P.JAC, "jacobian"), (P.HESS, "hessian")):
            assert_close(r[mode], want[key], key + " (set call)")
    finally:
        eng.close()


def test_full_size_device_resident_set_matches_oracle():
    """The path bench.py's ``value`` times: upload once, pk_run_set (one CUDA graph, five streams),
    results read back afterwards."""
    from pockit_b200 import plan as P
    from pockit_b200.engine import Engine

    S, x, lam, sigma, want = _reference("robot_arm")
    eng = Engine(S.lowering)
    try:
        modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
        eng.upload(x, lam, sigma)
        eng.time_steps(modes, 3, flush_l2=True)  # graph replays with the L2 flush in between, like the bench
        eng.sync()
        assert eng.expand_kernel(P.HESS) == "pk_expand_cols"
        for mode, key in zip(modes, ("objective", "gradient", "constraints", "jacobian", "hessian")):
            assert_close(np.atleast_1d(eng.download(mode)), np.atleast_1d(want[key]), key)
    finally:
        eng.close()


@pytest.mark.parametrize("fastmath", [False, True])
def test_quadrotor_8192_instances_match_per_instance_oracle(fastmath, monkeypatch):
    """BASELINE configs[4] at full size: B = 8192 instances, start + U(-0.2, 0.2) with rng seed 0;
    32 sampled instances (first, last, 30 random) against an oracle built for each of them.
    ``fastmath=True`` is what the example passes to Numba (``examples/planar_quadrotor.py:59``); here it
    maps to --fmad=true, so those values may differ from the strict ones in the last bits: the
    strict model is held to 1e-12 / 1e-14, the fastmath one to 1e-11 / 1e-13."""
    import pockit_b200.lobatto as lob
    from oracle.pockit_oracle import OracleSystem
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.batched import BatchedSystem, fixed_index, fixed_table

    B = 8192
    rng = np.random.default_rng(0)
    S = problems.quadrotor(lob, fastmath=fastmath)
    fixed = fixed_table(S, B)
    d0, d1 = rng.uniform(-0.2, 0.2, B), rng.uniform(-0.2, 0.2, B)
    base0, base1 = fixed[0, fixed_index(S, 0, "x0", 0)], fixed[0, fixed_index(S, 0, "x0", 1)]
    fixed[:, fixed_index(S, 0, "x0", 0)] += d0
    fixed[:, fixed_index(S, 0, "x0", 1)] += d1
    x0, lam0, _ = problems.evaluation_point(S)
    X = x0[None, :] + 1e-2 * rng.normal(size=(B, len(x0)))
    LAM = lam0[None, :] + 0.1 * rng.normal(size=(B, len(lam0)))
    sig = rng.uniform(0.5, 1.5, B)
    bs = BatchedSystem(S, fixed)
    try:
        # 84 (interval, column) pairs and 420 slots per list and instance.  Jacobian (one job of five lists): the
        # slot-order batch kernel (a thread owns one (instance, interval, row, column) slot of a job and writes it
        # for every list of the job); Hessian (jobs of one or two lists): the column mapping
        assert bs.engine.expand_kernel(P.JAC) == "pk_expand_slots" and bs.engine.expand_kernel(P.HESS) == "pk_expand_batch"
        obj, grad, cons = bs.objective(X), bs.gradient(X), bs.constraints(X)
        jac, hess = bs.jacobian(X), bs.hessian(X, LAM, sig)
        r = bs.engine.evaluate(X, LAM, sig)
        assert np.array_equal(r[P.JAC], jac) and np.array_equal(r[P.HESS], hess) and np.array_equal(r[P.GRAD], grad)
    finally:
        bs.close()
    # the other routes: the persistent kernel (one block per instance) writes the same bits; the
    # table-driven route through pk_generic_jobs (POCKIT_B200_BATCH_TABLES=1: no block kernel at all) agrees
    # to the last few ulp ((unit * width) / 2 is formed on the host there)
    monkeypatch.setenv("POCKIT_B200_EXPAND", "columns")
    bs = BatchedSystem(S, fixed)
    try:
        assert bs.engine.expand_kernel(P.JAC) == "pk_expand_blocks" and bs.engine.expand_kernel(P.HESS) == "pk_expand_blocks"
        assert np.array_equal(bs.jacobian(X), jac) and np.array_equal(bs.hessian(X, LAM, sig), hess)
    finally:
        bs.close()
    # the slot-order kernel with fewer lists per thread than the default (all of a job's)
    monkeypatch.setenv("POCKIT_B200_EXPAND", "slots")
    for per_thread in ("1", "2"):
        monkeypatch.setenv("POCKIT_B200_SLOT_LISTS", per_thread)
        bs = BatchedSystem(S, fixed)
        try:
            assert bs.engine.expand_kernel(P.JAC) == "pk_expand_slots" and bs.engine.expand_kernel(P.HESS) == "pk_expand_slots"
            assert np.array_equal(bs.jacobian(X), jac) and np.array_equal(bs.hessian(X, LAM, sig), hess)
        finally:
            bs.close()
    monkeypatch.delenv("POCKIT_B200_SLOT_LISTS")
    # the column mapping (a thread per (instance, interval, column), block column in registers): 16 unrolled rows
    # with two lists per thread, the exact row count with five
    monkeypatch.setenv("POCKIT_B200_EXPAND", "batch")
    for rows_mode, per_thread in (("", "2"), ("exact", "5")):
        monkeypatch.setenv("POCKIT_B200_BATCH_ROWS", rows_mode)
        monkeypatch.setenv("POCKIT_B200_BATCH_LISTS", per_thread)
        bs = BatchedSystem(S, fixed)
        try:
            assert bs.engine.expand_kernel(P.JAC) == "pk_expand_batch" and bs.engine.expand_kernel(P.HESS) == "pk_expand_batch"
            assert np.array_equal(bs.jacobian(X), jac) and np.array_equal(bs.hessian(X, LAM, sig), hess)
        finally:
            bs.close()
    monkeypatch.delenv("POCKIT_B200_BATCH_ROWS")
    monkeypatch.delenv("POCKIT_B200_BATCH_LISTS")
    monkeypatch.delenv("POCKIT_B200_EXPAND")
    monkeypatch.setenv("POCKIT_B200_BATCH_TABLES", "1")
    bs = BatchedSystem(S, fixed)
    try:
        assert bs.engine.expand_kernel(P.JAC) == "" and bs.engine.expand_kernel(P.HESS) == ""
        np.testing.assert_allclose(bs.jacobian(X), jac, rtol=1e-14, atol=1e-16)
        np.testing.assert_allclose(bs.hessian(X, LAM, sig), hess, rtol=1e-14, atol=1e-16)
    finally:
        bs.close()
    assert jac.shape == (B, 4321) and hess.shape == (B, 2058) and np.all(np.isfinite(hess))
    rtol, atol = (1e-11, 1e-13) if fastmath else (1e-12, 1e-14)
    sample = [0, B - 1] + sorted(int(b) for b in rng.choice(np.arange(1, B - 1), 30, replace=False))
    for b in sample:
        O = OracleSystem(problems.quadrotor(lob, start=(base0 + d0[b], base1 + d1[b]), fastmath=False))
        assert_close(obj[b], O.objective(X[b]), f"objective[{b}]", rtol, atol)
        assert_close(grad[b], O.gradient(X[b]), f"gradient[{b}]", rtol, atol)
        assert_close(cons[b], O.constraints(X[b]), f"constraints[{b}]", rtol, atol)
        assert_close(jac[b], O.jacobian(X[b]), f"jacobian[{b}]", rtol, atol)
        assert_close(hess[b], O.hessian(X[b], LAM[b], sig[b]), f"hessian[{b}]", rtol, atol)
