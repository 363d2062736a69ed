"""Golden episodes of make_beer_game(demand_type='deterministic_random') (envs.py:519-520) recorded from the live
reference: the demand list drawn at construction, observations / rewards of the 34 periods the reference survives,
and the IndexError of the 35th.  Run in the build container (needs /root/reference):
    python tests/golden/make_deterministic_random.py   ->  tests/golden/deterministic_random.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402

E = H.import_reference()["envs"]
out = {}
for cid, (seed, agents, max_period) in enumerate([(3, ["retailer"], 35), (11, ["retailer", "wholesaler"], 35), (5, ["distributor"], 20)]):
    env = E.make_beer_game(agent_managed_facilities=list(agents), demand_type="deterministic_random", max_period=max_period, seed=seed)
    acts = np.random.RandomState(100 + seed).randint(0, 17, (max_period, len(agents)))
    demand = np.asarray(env.sn.nodes["retailer"].demands._demands if hasattr(env.sn.nodes["retailer"], "demands") else [])
    obs0 = env.reset()
    obs, rew, err_at, err = [], [], -1, ""
    for t in range(max_period):
        try:
            o, r, d, _ = env.step([int(a) for a in acts[t]])
        except IndexError as exc:
            err_at, err = t, str(exc)
            break
        obs.append(o), rew.append(r)
        assert not d
    # second episode replays the same list: reset works again after the crash?  (the reference's reset re-creates the
    # generator, so it does)
    obs0_again = env.reset()
    p = f"c{cid}_"
    out.update({p + "seed": seed, p + "agents": np.array([H.NODE_NAMES.index(a) for a in agents]), p + "max_period": max_period,
                p + "actions": acts, p + "obs0": obs0, p + "obs": np.array(obs), p + "rew": np.array(rew), p + "err_at": err_at,
                p + "err": err, p + "obs0_again": obs0_again})
    print(cid, "crash at step", err_at, repr(err), "return so far", float(np.sum(rew)))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "deterministic_random.npz"), **out)
