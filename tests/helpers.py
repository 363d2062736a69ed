"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import numpy as np

from oracle import beergame_np as O

SO_OFF = (4, 6, 8, 10)
SH_OFF = (11, 13, 15, 17)
DPA = {"uniform": (-8, 0, 16), "normal": (-10, 0, 20), "classic": (0, 0, 16)}


def split_init(init):
    """[n,19] -> (inv0 [n,4], so list, sh list) in the oracle's format."""
    init = np.asarray(init)
    inv0 = init[:, :4]
    so = [init[:, SO_OFF[k]:SO_OFF[k] + O.INFO_LT[k]] for k in range(4)]
    sh = [init[:, SH_OFF[k]:SH_OFF[k] + 2] for k in range(4)]
    return inv0, so, sh


def oracle_rollout(scenario, agents, global_observable, rule, T, init, demand, actions, dtype=np.int64,
                   policy=None):
    """Run the numpy oracle.  actions [T,n,na] or None with policy=(targets, lb, ub, rule)."""
    spec = O.scenario_spec(scenario, agents, global_observable, T, rule)
    n = init.shape[0]
    orc = O.BeerGameOracle(spec, n, dtype)
    inv0, so, sh = split_init(init)
    obs0 = orc.reset(inv0, so, sh, demand)
    steps = actions.shape[0] if actions is not None else T
    obs = np.zeros((steps, n, obs0.shape[1]), np.float32)
    rew = np.zeros((steps, n))
    hold = np.zeros((steps, n, 4))
    back = np.zeros((steps, n, 4))
    acts = np.zeros((steps, n, len(agents)))
    cur = obs0
    for t in range(steps):
        if actions is not None:
            a = actions[t]
        else:
            tg, lb, ub, prule = policy
            # the policy sees the agent-managed facilities' own 7 numbers (local view of each agent)
            full = np.stack([np.stack(orc._node_state(k, False), axis=1) for k in agents], axis=1).reshape(n, -1)
            a = O.base_stock_action(full, tg, lb, ub, prule)
            if np.issubdtype(np.dtype(dtype), np.integer):
                a = a.astype(np.int64)
        acts[t] = a
        cur, r, done = orc.step(a)
        obs[t], rew[t], hold[t], back[t] = cur, r, orc.holding, orc.backorder
    return dict(obs0=obs0, obs=obs, rew=rew, holding=hold, backorder=back, actions=acts, done=done)
