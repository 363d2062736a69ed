import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden_rollouts():
    z = np.load(os.path.join(GOLDEN, "rollouts.npz"))
    cases = []
    for m in z["meta"]:
        cid, scn, ag, glob, rule, kind, seed = str(m).split("|")
        p = f"c{int(cid):03d}_"
        cases.append(dict(
            id=f"{scn}-ag{ag.replace(',', '')}-g{glob}-{rule}-{kind}", scenario=scn,
            agents=tuple(int(x) for x in ag.split(",")), global_observable=bool(int(glob)), rule=rule, kind=kind,
            seed=int(seed), init=z[p + "init"], demand=z[p + "demand"], actions=z[p + "actions"],
            obs0=z[p + "obs0"], obs=z[p + "obs"], rew=z[p + "rew"], done=z[p + "done"],
            holding=z[p + "holding"], backorder=z[p + "backorder"]))
    return cases


def load_golden_classic():
    z = np.load(os.path.join(GOLDEN, "classic.npz"))
    out = []
    c = 0
    while f"k{c:02d}_init" in z:
        p = f"k{c:02d}_"
        out.append({k: z[p + k] for k in ("init", "demand", "actions", "obs0", "obs", "rew", "agents")})
        c += 1
    return out


# tests/test_supplynetwork.py:25-69 of the reference, in the order make_golden.py replays them
CLASSIC_TOTALS = [-1656.0, -798.0, -1260.0, -224.0, -644.0, -658.0, -170.0, -784.0, -644.0,
                  -158.0, -864.0, -888.0, -1656.0]


@pytest.fixture(scope="session")
def golden_rollouts():
    return load_golden_rollouts()


@pytest.fixture(scope="session")
def golden_classic():
    return load_golden_classic()
