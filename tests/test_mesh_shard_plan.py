"""CPU tier: mesh sharding of ONE fine mesh over several ranks.  Every rank's plan
keeps a share of the slot-streaming jobs; emulated on the host, the shares must partition the
output and reassemble, bit for bit, to the unsharded evaluation (and hence to the reference)."""
import numpy as np
import pytest

from helpers import assert_close, build, load
from hostemu import HostEmu
from pockit_b200 import plan as P


def _assemble(S, G, mode, x, lam=None, sigma=None):
    full = None
    cover = None
    for g in range(G):
        E = HostEmu(S, shard=(g, G))
        out = E.run(mode, x, lam, sigma)
        runs = E.fin[mode]["runs"]
        if full is None:
            full = np.full(len(out), np.nan)
            cover = np.zeros(len(out), dtype=np.int64)
        for off, cnt in runs:
            full[off : off + cnt] = out[off : off + cnt]
            cover[off : off + cnt] += 1
    return full, cover


@pytest.mark.parametrize("G,split_min", [(2, 4), (3, 4), (8, 1024)])  # small runs split too / production threshold
@pytest.mark.parametrize("case", ["robot_arm_lgr_6x20", "rocket_lgl_4x5", "quadrotor_lgl_14x6", "general_lgl", "general_lgr", "tiny_lgl_2x2"])
def test_shards_partition_and_reassemble(case, G, split_min, monkeypatch):
    monkeypatch.setattr(P, "SPLIT_MIN", split_min)
    S, g = build(case), load(case)
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    whole = HostEmu(S)
    for mode, args, name in ((P.JAC, (x,), "jacobian"), (P.HESS, (x, lam, sigma), "hessian"),
                             (P.CONS, (x,), "constraints"), (P.GRAD, (x,), "gradient"), (P.OBJ, (x,), "objective")):
        full, cover = _assemble(S, G, mode, *args)
        assert np.all(cover == 1), f"{name}: every slot must be owned by exactly one rank"
        assert np.array_equal(full, whole.run(mode, *args)), name
        assert_close(full if mode != P.OBJ else full[0], g[name], name)


def test_fine_mesh_shares_are_balanced():
    import pockit_b200.radau as rad
    from pockit_b200 import problems

    S = problems.robot_arm(rad, mesh=400, num_point=20)
    lo = S.lowering
    G = 8
    for mode, n_out in ((P.JAC, lo.nnz_jac), (P.HESS, lo.nnz_hess_o + lo.nnz_hess_c)):
        owned = []
        for g in range(G):
            dp = P.DevicePlan(lo, shard=(g, G))
            for m in range(5):
                dp.mode(m)
            runs = dp.finalize(mode)["runs"]
            owned.append(int(runs[:, 1].sum()))
        assert sum(owned) == n_out
        assert max(owned) <= 1.1 * n_out / G  # near-equal shares: the copy time is what shards


def test_bad_shard_arguments():
    S = build("tiny_lgl_2x2")
    with pytest.raises(ValueError):
        P.DevicePlan(S.lowering, shard=(2, 2))
    with pytest.raises(ValueError):
        P.DevicePlan(S.lowering, shard=(0, 2), fused=True)


@pytest.mark.parametrize("case", ["robot_arm_lgr_6x20", "rocket_lgl_4x5", "general_lgr"])
def test_weighted_shares_partition_and_follow_the_weights(case, monkeypatch):
    """shard = (rank, world, weights): shares proportional to the ranks' measured device-to-host rates
    (meshshard) must still partition the output and reassemble bit for bit."""
    monkeypatch.setattr(P, "SPLIT_MIN", 4)
    S, g = build(case), load(case)
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    whole = HostEmu(S)
    w = [0.5, 1.3, 0.9, 2.0]
    for mode, args in ((P.JAC, (x,)), (P.HESS, (x, lam, sigma)), (P.CONS, (x,))):
        want = whole.run(mode, *args)
        full, cover, owned = np.full(len(want), np.nan), np.zeros(len(want), dtype=np.int64), []
        for r in range(4):
            E = HostEmu(S, shard=(r, 4, w))
            out = E.run(mode, *args)
            runs = E.fin[mode]["runs"]
            owned.append(int(sum(c for _, c in runs)))
            for off, cnt in runs:
                full[off : off + cnt] = out[off : off + cnt]
                cover[off : off + cnt] += 1
        assert np.all(cover == 1) and np.array_equal(full, want)
        if mode != P.CONS and case == "robot_arm_lgr_6x20":  # the splittable part dominates: shares follow the weights
            assert owned[3] > owned[1] > owned[2] > owned[0]
    with pytest.raises(ValueError):
        P.DevicePlan(S.lowering, shard=(0, 4, [1.0, 0.0, 1.0, 1.0]))
    with pytest.raises(ValueError):
        P.DevicePlan(S.lowering, shard=(0, 4, [1.0, 1.0]))
