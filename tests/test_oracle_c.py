"""Pins the C form of the oracle (oracle/beergame_c.c) to the reference's golden vectors and to the
numpy oracle.  Bit-exact (integer quantities)."""
import numpy as np
import pytest

from conftest import CLASSIC_TOTALS, load_golden_classic, load_golden_rollouts
from helpers import oracle_rollout
from oracle import beergame_c as OC
from oracle import beergame_np as O

GOLD = [c for c in load_golden_rollouts() if c["kind"] == "int"]


@pytest.mark.parametrize("case", GOLD, ids=[c["id"] for c in GOLD])
def test_c_oracle_matches_golden(case):
    spec = O.scenario_spec(case["scenario"], case["agents"], case["global_observable"], 100, case["rule"])
    out = OC.rollout(spec, case["init"][None], case["demand"][None], case["actions"][:, None, :])
    assert np.array_equal(out["obs0"][0], case["obs0"])
    assert np.array_equal(out["obs"][:, 0], case["obs"])
    assert np.array_equal(out["rew"][:, 0], case["rew"])


def test_c_oracle_classic_totals():
    for ep, total in zip(load_golden_classic(), CLASSIC_TOTALS):
        spec = O.scenario_spec("classic", tuple(int(a) for a in ep["agents"]), False, 35)
        out = OC.rollout(spec, ep["init"][None], ep["demand"][None, :38], ep["actions"][:, None, :])
        assert np.array_equal(out["obs"][:, 0], ep["obs"]) and out["rew"].sum() == total


@pytest.mark.parametrize("threads", [1, 3])
def test_c_oracle_vs_numpy_batch(threads):
    rs = np.random.RandomState(4)
    n, T = 777, 100
    init = np.concatenate([rs.randint(0, 25, (n, 4)), rs.randint(0, 9, (n, 15))], axis=1)
    demand = rs.randint(0, 9, (n, T + 3))
    acts = rs.randint(-1, 19, (T, n, 4))
    for glob, rule in ((True, "a"), (False, "d+a")):
        ref = oracle_rollout("uniform", (0, 1, 2, 3), glob, rule, T, init, demand, acts)
        spec = O.scenario_spec("uniform", (0, 1, 2, 3), glob, T, rule)
        out = OC.rollout(spec, init, demand, acts, n_threads=threads)
        assert np.array_equal(out["obs"], ref["obs"]) and np.array_equal(out["rew"], ref["rew"])
    ref = oracle_rollout("uniform", (0, 1, 2, 3), True, "a", T, init, demand, None, policy=([19, 20, 20, 14], 0, 16, "a"))
    out = OC.rollout(O.scenario_spec("uniform", (0, 1, 2, 3), True, T), init, demand, policy=([19, 20, 20, 14], 0, 16, "a"),
                     n_threads=threads)
    assert np.array_equal(out["obs"], ref["obs"]) and np.array_equal(out["ep_return"], ref["rew"].sum(0))
