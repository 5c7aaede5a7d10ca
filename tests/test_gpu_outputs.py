"""GPU tests of the two output-side extensions next to the parity contract:
the single-call evaluation set (``pk_eval_set``) and the opt-in
de-duplicated patterns (``pk_engine_set_compaction``)."""
import importlib

import numpy as np
import pytest

from helpers import assert_close, build, load

pytestmark = pytest.mark.gpu


def _dense(rows, cols, vals, shape):
    M = np.zeros(shape)
    np.add.at(M, (rows, cols), vals)
    return M


@pytest.mark.parametrize("pinned", [False, True])
def test_evaluate_equals_the_five_callbacks(pinned, monkeypatch):
    """One engine call for the whole set returns bit for bit what the five reference-style
    callbacks return one after the other (per-callback pipelines, POCKIT_B200_SET=0; the fused
    set pipeline is compared in test_gpu_parity.py, to the parity tolerance)."""
    import pockit_b200.lobatto as lob
    from pockit_b200 import problems

    monkeypatch.setenv("POCKIT_B200_SET", "0")
    S = problems.rocket(lob, mesh=30, num_point=8)
    x, lam, sigma = problems.evaluation_point(S, seed=11)
    want = dict(objective=S.objective(x), gradient=S.gradient(x), constraints=S.constraints(x),
                jacobian=S.jacobian(x), hessian=S.hessian(x, lam, sigma))
    S.pinned_outputs = pinned
    x_in = x.copy()
    got = S.evaluate(x, lam, sigma)
    assert np.array_equal(x, x_in)
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), k
    first = S.evaluate(x)  # without multipliers: no Hessian
    assert set(first) == {"objective", "gradient", "constraints", "jacobian"}
    assert np.array_equal(first["jacobian"], want["jacobian"])
    # a second point through the same engine (staging buffers are reused)
    x2 = x + 1e-3
    got2 = S.evaluate(x2, lam, 0.3)
    S.pinned_outputs = False
    assert np.array_equal(got2["hessian"], S.hessian(x2, lam, 0.3))
    assert got2["objective"] == S.objective(x2)


@pytest.mark.parametrize("case", ["general_lgl", "general_lgr", "rocket_lgl_4x5", "robot_arm_lgr_6x20", "quadrotor_lgl_14x6"])
def test_compact_patterns_sum_the_reference_duplicates(case):
    """De-duplicated patterns: unique sorted (row, col) pairs whose values are the sums of the
    reference's duplicate entries (golden vectors of the real reference, summed on the host)."""
    S = build(case)
    g = load(case)
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    L, m = S.L, len(S.c_lb)
    jr, jc = S.jacobianstructure()
    hr, hc = S.hessianstructure()
    S.compact_patterns = True
    cjr, cjc = S.jacobianstructure()
    chr_, chc = S.hessianstructure()
    for r, c, n in ((cjr, cjc, L), (chr_, chc, L)):
        key = r.astype(np.int64) * n + c
        assert np.all(np.diff(key) > 0)  # unique, sorted by row then column
    assert len(cjr) <= len(jr) and len(chr_) <= len(hr)
    jac, hess = S.jacobian(x.copy()), S.hessian(x.copy(), lam, sigma)
    assert jac.shape == cjr.shape and hess.shape == chr_.shape
    # the sum of duplicates may cancel: tolerance relative to the sum of magnitudes
    for got, r, c, ref, r0, c0, shape in ((jac, cjr, cjc, g["jacobian"], jr, jc, (m, L)),
                                          (hess, chr_, chc, g["hessian"], hr, hc, (L, L))):
        want = _dense(r0, c0, ref, shape)
        scale = _dense(r0, c0, np.abs(ref), shape)
        G = _dense(r, c, got, shape)
        assert np.all(np.abs(G - want) <= 1e-14 + 4e-12 * scale)
        assert np.count_nonzero(_dense(r, c, np.ones(len(r)), shape) > 1) == 0
    # SciPy-style split Hessians live on the merged pattern
    ho, hcn = S.hessian_o(x.copy()), S.hessian_c(x.copy(), lam)
    assert ho.shape == hess.shape and hcn.shape == hess.shape
    np.testing.assert_allclose(sigma * ho + hcn, hess, rtol=1e-11, atol=1e-13)
    got = S.evaluate(x.copy(), lam, sigma)
    assert np.array_equal(got["jacobian"], jac) and np.array_equal(got["hessian"], hess)
    # and back to the reference contract
    S.compact_patterns = False
    assert_close(S.jacobian(x.copy()), g["jacobian"], "jacobian")
    assert_close(S.hessian(x.copy(), lam, sigma), g["hessian"], "hessian")


def test_compact_patterns_at_full_size():
    """BASELINE configs[1]: 12.8 M Hessian slots collapse to 0.56 M unique entries; the compacted
    values equal the segmented sums of the uncompacted ones."""
    import pockit_b200.radau as rad
    from pockit_b200 import problems

    S = problems.robot_arm(rad, 2000, 20)
    x, lam, sigma = problems.evaluation_point(S)
    jac, hess = S.jacobian(x), S.hessian(x, lam, sigma)
    S.compact_patterns = True
    cj, ch = S.lowering.compaction("jac"), S.lowering.compaction("hess")
    assert len(ch["row"]) == 559989 and len(cj["row"]) == 7919754
    for full, c, got in ((jac, cj, S.jacobian(x)), (hess, ch, S.hessian(x, lam, sigma))):
        want = np.add.reduceat(full[c["perm"]], c["ptr"][:-1])
        scale = np.add.reduceat(np.abs(full)[c["perm"]], c["ptr"][:-1])
        assert got.shape == want.shape
        assert np.all(np.abs(got - want) <= 1e-14 + 1e-12 * scale)


def test_batched_compaction():
    import pockit_b200.lobatto as lob
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.batched import BatchedSystem

    S = problems.quadrotor(lob, fastmath=False)
    B = 5
    rng = np.random.default_rng(4)
    x0, lam0, _ = problems.evaluation_point(S)
    X = x0[None, :] + 1e-2 * rng.normal(size=(B, len(x0)))
    bs = BatchedSystem(S, batch=B)
    full = bs.jacobian(X).copy()
    c = S.lowering.compaction("jac")
    bs.engine.set_compaction(P.JAC, c["ptr"], c["perm"])
    got = bs.jacobian(X)
    assert got.shape == (B, len(c["row"]))
    for b in range(B):
        want = np.add.reduceat(full[b][c["perm"]], c["ptr"][:-1])
        np.testing.assert_allclose(got[b], want, rtol=1e-12, atol=1e-13)
    bs.close()


def test_staged_asynchronous_evaluation_matches_the_blocking_call():
    """pk_eval_set_async: the inputs arrive in stages (x first, the multipliers later -- a mesh-shard
    worker), every stage is enqueued without waiting and one pk_sync completes all results."""
    import pockit_b200.radau as rad
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.engine import Engine

    S = problems.robot_arm(rad, mesh=120, num_point=12)
    x, lam, sigma = problems.evaluation_point(S, seed=6)
    eng = Engine(S.lowering)
    try:
        want = eng.evaluate(x, lam, sigma)
        want = {m: np.array(v, copy=True) for m, v in want.items()}
        for trial in range(3):
            xs = x + 1e-3 * trial
            ref = {m: np.array(v, copy=True) for m, v in eng.evaluate(xs, lam, sigma).items()}
            first = eng.evaluate(xs, modes=[P.OBJ, P.GRAD, P.CONS, P.JAC], wait=False)
            second = eng.evaluate(None, lam, sigma, modes=[P.HESS], wait=False)  # at the resident x
            assert np.ndim(first[P.OBJ]) == 1  # not read before the sync: still the one-element buffer
            eng.sync()
            got = {**first, **second}
            for m in ref:
                assert np.array_equal(np.atleast_1d(got[m]), np.atleast_1d(ref[m])), P.MODES[m]
        # a blocking call right behind an unfinished asynchronous one is ordered correctly too
        eng.evaluate(x, modes=[P.JAC], wait=False)
        again = eng.evaluate(x, lam, sigma)
        for m in want:
            assert np.array_equal(np.atleast_1d(again[m]), np.atleast_1d(want[m])), P.MODES[m]
    finally:
        eng.close()
