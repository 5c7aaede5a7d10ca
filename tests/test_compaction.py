"""CPU tier: the de-duplication tables of the opt-in compact patterns are a
lossless regrouping of the reference's COO patterns."""
import numpy as np
import pytest

from helpers import build, golden_cases, load

CASES = sorted(golden_cases())


@pytest.mark.parametrize("case", CASES)
def test_compaction_tables_regroup_the_reference_pattern(case):
    S = build(case)
    g = load(case)
    lo = S.lowering
    L, m = S.L, len(S.c_lb)
    for kind, (rows, cols), vals, shape in (
        ("jac", S.jacobianstructure(), g["jacobian"], (m, L)),
        ("hess", S.hessianstructure(), g["hessian"], (L, L)),
    ):
        c = lo.compaction(kind)
        n = len(rows)
        assert np.array_equal(np.sort(c["perm"]), np.arange(n))  # every slot exactly once
        assert c["ptr"][0] == 0 and c["ptr"][-1] == n and np.all(np.diff(c["ptr"]) > 0 if n else True)
        key = c["row"] * L + c["col"]
        assert np.all(np.diff(key) > 0)
        # every slot of a segment carries the segment's (row, col); slots inside a segment increase
        seg = np.repeat(np.arange(len(c["row"])), np.diff(c["ptr"]))
        assert np.array_equal(rows[c["perm"]], c["row"][seg]) and np.array_equal(cols[c["perm"]], c["col"][seg])
        inner = np.diff(c["perm"])
        starts = c["ptr"][1:-1] - 1
        assert np.all(np.delete(inner, starts) > 0)
        if n == 0:
            continue
        want = np.zeros(shape)
        np.add.at(want, (rows, cols), vals)
        got = np.zeros(shape)
        got[c["row"], c["col"]] = np.add.reduceat(vals[c["perm"]], c["ptr"][:-1])
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)


def test_compact_structures_switch_with_the_flag():
    S = build("robot_arm_lgr_6x20")
    jr, jc = S.jacobianstructure()
    hr, hc = S.hessianstructure()
    S.compact_patterns = True
    cjr, cjc = S.jacobianstructure()
    chr_, chc = S.hessianstructure()
    assert len(cjr) < len(jr) and len(chr_) < len(hr)
    assert set(zip(cjr.tolist(), cjc.tolist())) == set(zip(jr.tolist(), jc.tolist()))
    assert set(zip(chr_.tolist(), chc.tolist())) == set(zip(hr.tolist(), hc.tolist()))
    for a, b in zip(S.hessianstructure_o(), S.hessianstructure()):
        assert np.array_equal(a, b)
    S.compact_patterns = False
    assert np.array_equal(S.jacobianstructure()[0], jr) and np.array_equal(S.hessianstructure()[1], hc)
