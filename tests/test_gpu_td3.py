"""TD3 + HGE on the device VecEnv (config 5 wiring): runs, stores original transitions with the terminal
observation at episode ends, decays the HGE rate like HgeRateCallback, reports the time split."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_td3_hge_short_run_and_buffer_contents():
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
    from multi_echelon_drl_b200.utils.heuristics import BaseStockPolicy
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    n = 256
    env = VecNormalize(R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="float32", seed=0, use_graph=True),
                       clip_obs=100.0, clip_reward=1000.0)
    cfg = TD3Config(net_arch=(64,), train_freq=8, gradient_steps=4, buffer_size=n * 120, learning_starts=n * 4,
                    hge_rate_at_start=0.5, hge_decay_periods=1)
    model = HgeTD3(env, BaseStockPolicy([19, 20, 20, 14]), cfg)
    model.learn(n * 104)
    assert model.num_timesteps == n * 104 and model.n_updates > 0
    buf = model.buffer
    assert len(buf) == n * 104
    # transition 99 of every env is the episode end: done flag set, next_obs = terminal observation
    end = slice(99 * n, 100 * n)
    assert (buf.done[end] == 1).all() and (buf.done[: 99 * n] == 0).all()
    # rewards are the ORIGINAL (un-normalised) costs: non-positive multiples of 0.5 for integer-free float actions? just sign
    assert (buf.rew[: len(buf)] <= 0).all()
    # stored actions are in the scaled [-1, 1] space (SB3 convention)
    assert buf.act[: len(buf)].abs().max() <= 1.0
    # observations in the buffer are un-normalised (inventory levels etc., not z-scores)
    assert buf.obs[: len(buf)].abs().max() > 5
    # HGE rate decays linearly to 0 over 100 * decay_periods timesteps (utils/callbacks.py:19-21)
    assert model.hge_rate == 0.0
    split = model.time_split()
    assert split["env_step_s"] > 0 and split["update_s"] > 0 and 0 < split["env_step_frac"] < 1
    assert len(model.episode_returns) == 1 and np.isfinite(model.episode_returns[0])
    for p in model.actor.parameters():
        assert torch.isfinite(p).all()
    # past the uniform warm-up the collection ran as three graph replays per vector step (action, env step, store):
    # the device-side write position / length agree with the host's, and consecutive transitions chain
    assert model._collect is not None and model._collect.store is not None and model._collect.heur is not None
    assert int(buf.pos_t.item()) == buf.pos == n * 104 and buf.size_t.item() == len(buf)
    for t in (0, 3, 4, 5, 50, 98, 100, 102):
        assert torch.equal(buf.next_obs[t * n:(t + 1) * n], buf.obs[(t + 1) * n:(t + 2) * n]), t
    assert not torch.equal(buf.next_obs[99 * n:100 * n], buf.obs[100 * n:101 * n])      # terminal observation vs reset
    # the stored action is the scaled form of the order the env received: latest order of the retailer in next_obs
    # (observation column 3 + lead-time slot is env-specific; check through the reward sign and range only)
    assert (buf.act[4 * n: 104 * n].abs() <= 1).all() and buf.act[4 * n: 104 * n].std() > 0.1
    # what the env's action buffer received last = unscale_action(last stored scaled action), inside the action space
    last = buf.act[103 * n: 104 * n]
    assert torch.allclose(env.venv.action_buffer, model.unscale_action(last), atol=1e-5)
    assert env.venv.action_buffer.min() >= 0 and env.venv.action_buffer.max() <= 16 and env.venv.action_buffer.std() > 1


def test_td3_learns_to_beat_random_on_basic_retailer():
    """Sanity: a few thousand updates on the decentralized Retailer env (config 1 shape) lower the cost
    clearly below the uniform-random warm-up policy."""
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    n = 512
    env = VecNormalize(R.make("BeerGameNormalRetailer-v0", n_envs=n, dtype="float32", seed=0), clip_obs=100.0,
                       clip_reward=1000.0)
    cfg = TD3Config(net_arch=(128,), batch_size=512, learning_rate=3e-4, tau=0.02, train_freq=10, gradient_steps=40,
                    action_noise_std=0.2, gamma=0.95, hge_rate_at_start=0.0, buffer_size=n * 100 * 8,
                    learning_starts=n * 100)
    model = HgeTD3(env, None, cfg)
    model.learn(n * 100 * 14)
    # returns are negative costs (higher is better).  Over seeds 0..2 the random warm-up episode costs about
    # 17,000 and training reaches 2,500-3,000 (profiles/td3_margin_probe.py); TD3 can be unstable late in a run,
    # so the check is on the best of the last six episodes with a wide margin.
    first, best_late = model.episode_returns[0], max(model.episode_returns[-6:])
    assert best_late > 0.6 * first, (first, model.episode_returns)


def test_td3_hge_centralized_learns_to_the_heuristics_level():
    """BASELINE config 5 (centralized complex scenario, 4,096 envs, TD3 with heuristic-guided exploration whose rate
    decays linearly like HgeRateCallback, utils/callbacks.py:5-23; collection = three graph replays per vector step,
    step + VecNormalize as one kernel): 80 episodes with batch 2,048, learning rate 3e-4 and 4 updates per vector step
    bring the deterministic policy from the uniform-random level (about -16,900 per episode) past the base-stock
    heuristic that guided it (-1,337).  Measured over seeds 0..7 and three builds (each a different realisation of
    the same distributions): 22 of 24 runs end between -795 and -1,228, two are still at -2,147 / -3,579 when the
    budget ends (profiles/r2_td3_learning.json, which also records what does NOT hold up: learning rate 1e-3
    diverges on one seed in four, 60 episodes leave three seeds in eight near -3,000 .. -5,000).  Two seeds here: both
    must be far from random, the better one at the heuristic's level."""
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
    from multi_echelon_drl_b200.utils.heuristics import BaseStockPolicy
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    n, episodes = 4096, 80
    total = n * 100 * episodes
    evals = []
    for seed in (0, 1):
        env = VecNormalize(R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="float32", seed=0, use_graph=True),
                           clip_obs=100.0, clip_reward=1000.0)
        cfg = TD3Config(batch_size=2048, net_arch=(768,), learning_rate=3e-4, tau=0.01, train_freq=32, gradient_steps=128,
                        action_noise_std=0.3, gamma=0.95, hge_rate_at_start=0.5, hge_decay_periods=total // 200,
                        learning_starts=n * 100, seed=seed)
        model = HgeTD3(env, BaseStockPolicy([19, 20, 20, 14]), cfg)
        model.learn(total)
        assert env.venv.sim.device_flags() == 0
        assert model._collect is not None and model._collect.store is not None      # the graphed collection path ran
        assert model.hge_rate == 0.0                                  # decayed to zero half-way through the run
        first, late = model.episode_returns[0], float(np.mean(model.episode_returns[-5:]))
        assert first < -12000 and late > -8000, model.episode_returns  # training episodes (with exploration noise)
        eval_env = VecNormalize(R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="float32", seed=123),
                                clip_obs=100.0, clip_reward=1000.0)
        mean, std, cnt = model.evaluate(eval_env, n_episodes=1)
        assert cnt == n and mean > -6000, (seed, mean, std)           # uniform random: -16,858
        evals.append(mean)
        model.close()
    assert max(evals) > -2000, evals                                  # heuristic: -1,337


def test_evaluate_and_checkpoint_roundtrip(tmp_path):
    """EvalCallback-style deterministic evaluation on a frozen, statistics-synchronised VecNormalize, and the
    best-model + env-statistics checkpoint (SaveEnvStatsCallback)."""
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    mk = lambda seed: VecNormalize(R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=128, dtype="float32", seed=seed),
                                   clip_obs=100.0, clip_reward=1000.0)
    cfg = TD3Config(net_arch=(32,), train_freq=8, gradient_steps=2, buffer_size=128 * 40, learning_starts=128 * 8,
                    hge_rate_at_start=0.0)
    model = HgeTD3(mk(0), None, cfg)
    model.learn(128 * 24)
    eval_env = mk(99)
    mean, std, n = model.evaluate(eval_env, n_episodes=2)
    assert n == 256 and np.isfinite(mean) and mean < 0 and std > 0
    assert torch.equal(eval_env._obs_rms, model.vec_normalize._obs_rms) and eval_env.training is False
    mean2, _, _ = model.evaluate(mk(99), n_episodes=2)
    assert mean2 == mean                                    # deterministic policy + same env seed
    path = str(tmp_path / "best_model.pt")
    model.save(path)
    other = HgeTD3(mk(0), None, cfg).load(path)
    assert other.num_timesteps == model.num_timesteps
    for a, b in zip(other.actor.parameters(), model.actor.parameters()):
        assert torch.equal(a, b)
    assert torch.equal(other.vec_normalize._obs_rms, model.vec_normalize._obs_rms)
    assert other.evaluate(mk(99), n_episodes=2)[0] == mean


def test_graphed_update_equals_eager_in_effect_and_is_faster():
    """The TD3 update replayed as a CUDA graph (policy_delay updates per replay, sampling inside the graph from a
    device-side buffer length, frozen VecNormalize statistics) trains like the eager loop on the same buffer, keeps
    the update counter exact for odd step counts, and takes less GPU time."""
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    import torch.nn.functional as F
    n = 256
    models = {}
    for graph in (False, True):
        env = VecNormalize(R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, dtype="float32", seed=3),
                           clip_obs=100.0, clip_reward=1000.0)
        cfg = TD3Config(net_arch=(256,), train_freq=100, gradient_steps=0, buffer_size=n * 100, learning_starts=n * 200,
                        hge_rate_at_start=0.0, cuda_graph=graph, seed=5)
        m = HgeTD3(env, None, cfg)
        m.learn(n * 100)                                  # one episode of uniform-random actions fills the buffer
        assert m.n_updates == 0 and len(m.buffer) == n * 100 and m.buffer.size_t.item() == n * 100
        models[graph] = m
    a, b = models[False], models[True]
    assert torch.equal(a.buffer.obs, b.buffer.obs) and torch.equal(a.buffer.rew, b.buffer.rew)
    for pa, pb in zip(a.critics.parameters(), b.critics.parameters()):
        assert torch.equal(pa, pb)

    def critic_loss(m):
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        obs, act, nobs, rew, done = m.buffer.sample(4096, g)
        obs, rew = m._normalize(obs, rew); nobs, _ = m._normalize(nobs, rew)
        with torch.no_grad():
            x = torch.cat([nobs, m.actor_target(nobs)], 1)
            tgt = rew + (1 - done) * m.cfg.gamma * torch.min(m.critics_target[0](x), m.critics_target[1](x)).squeeze(1)
            return float(F.mse_loss(m.critics[0](torch.cat([obs, act], 1)).squeeze(1), tgt))

    l0 = critic_loss(a)
    times = {}
    for graph, m in models.items():
        m.train(7)                                        # graph: 3 eager warm-up cycles, then capture is pending
        assert m.n_updates == 7
        m.train(3)                                        # from an odd count: one eager step to align, capture, one replay
        assert m.n_updates == 10                          # (the capture takes half a second of CPU time: outside the timed region)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        m.train(401)                                      # 200 replays and one eager step for the odd remainder
        e1.record(); torch.cuda.synchronize()
        times[graph] = e0.elapsed_time(e1)
        assert m.n_updates == 411
        for p in list(m.actor.parameters()) + list(m.critics.parameters()) + list(m.actor_target.parameters()):
            assert torch.isfinite(p).all()
    assert b._graph is not None and a._graph is None
    la, lb = critic_loss(a), critic_loss(b)
    assert la < 0.5 * l0 and lb < 0.5 * l0 and 0.33 < la / lb < 3.0, (l0, la, lb)
    # targets moved by Polyak averaging in both
    assert times[True] < 0.8 * times[False], times          # measured: 0.5 vs 1.66 ms per update
    print("update ms per step: eager %.3f graph %.3f" % (times[False] / 401, times[True] / 401))


def test_td3_hge_runs_on_an_eight_node_tree():
    """The whole stack composes on a network the reference's scripts never built: 8-node distribution tree (obs 56,
    8 action columns) -> VecNormalize -> TD3 with heuristic-guided exploration (wide base-stock kernel)."""
    from multi_echelon_drl_b200.spec import tree_spec
    from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
    from multi_echelon_drl_b200.utils.heuristics import BaseStockPolicy
    from multi_echelon_drl_b200.vec_env import SupplyNetVecEnv
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    names = ["hub", "r1", "r2", "r3", "r4", "r5", "factory", "annex"]
    targets = [60, 19, 14, 10, 8, 6, 40, 12]
    spec = tree_spec([dict(name=nm, target=t, demand=d) for nm, t, d in zip(
        names, targets, [None, (0, 8), (0, 6), (0, 4), (0, 3), (0, 2), None, (0, 5)])],
        [("ext", "factory", 1, 2), ("factory", "annex", 1, 3), ("factory", "hub", 2, 2), ("hub", "r3", 2, 1), ("hub", "r1", 2, 2),
         ("hub", "r5", 3, 1), ("hub", "r2", 1, 1), ("hub", "r4", 1, 2)], names, True)
    n = 128
    env = VecNormalize(SupplyNetVecEnv(spec, n, seed=1, dtype="float32"), clip_obs=100.0, clip_reward=1000.0)
    assert env.observation_space.shape == (56,) and env.action_space.shape == (8,)
    cfg = TD3Config(net_arch=(64,), train_freq=10, gradient_steps=4, buffer_size=n * 250, learning_starts=n * 10,
                    hge_rate_at_start=0.9, hge_decay_periods=300)          # decays over 100 x 300 timesteps = 234 vector steps
    model = HgeTD3(env, BaseStockPolicy(targets), cfg)
    model.learn(n * 210)
    assert model.n_updates > 0 and len(model.episode_returns) == 2 and all(np.isfinite(model.episode_returns))
    assert model.timers.heuristic > 0                       # the heuristic branch was taken (p = 0.81 at the start)
    assert model.buffer.obs.shape[1] == 56 and model.buffer.act.shape[1] == 8 and model.buffer.act.abs().max() <= 1
    for p in model.actor.parameters():
        assert torch.isfinite(p).all()


def test_device_replay_buffer_kernels_match_the_indexing_they_replace():
    """sn_replay_add / sn_replay_sample (csrc/replay_kernels.cu) against the PyTorch indexing of ReplayBuffer.add /
    .sample: ring wrap-around, `carry` aliasing the source, the device-side counters, a captured graph replayed, and the
    sampler's uniformity and determinism."""
    from multi_echelon_drl_b200.td3 import ReplayBuffer
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(0)
    n, od, ad, cap = 96, 28, 4, 1000                     # 1000 is not a multiple of 96: the ring wraps mid-step
    a, b = ReplayBuffer(cap, od, ad, dev), ReplayBuffer(cap, od, ad, dev)
    cur = torch.randn((n, od), generator=g, device=dev)
    cur_b = cur.clone()
    for step in range(25):
        nxt = torch.randn((n, od), generator=g, device=dev)
        act = torch.rand((n, ad), generator=g, device=dev) * 2 - 1
        rew = -torch.rand(n, generator=g, device=dev) * 30
        done = (torch.rand(n, generator=g, device=dev) < 0.1).float() if step % 3 == 0 else None
        a.add(cur, nxt, act, rew, done if done is not None else torch.zeros(n, device=dev))
        cur = nxt.clone()
        b.add_on_device(cur_b, nxt, act, rew, done, carry=cur_b)          # carry aliases the source: orig <- new_orig
        b.advance_host(n)
        assert torch.equal(cur_b, nxt)
        for x, y in ((a.obs, b.obs), (a.next_obs, b.next_obs), (a.act, b.act), (a.rew, b.rew), (a.done, b.done)):
            assert torch.equal(x, y), step
        assert int(b.pos_t.item()) == a.pos == b.pos and b.size_t.item() == len(a) == len(b)
    assert a.full and len(a) == cap
    # one captured store replayed: same as eager calls
    c = ReplayBuffer(cap, od, ad, dev)
    src, nxt, act, rew = (torch.zeros((n, od), device=dev), torch.zeros((n, od), device=dev), torch.zeros((n, ad), device=dev),
                          torch.zeros(n, device=dev))
    c.add_on_device(src, nxt, act, rew)                                   # (warm-up outside the capture)
    c.pos_t.zero_(); c.size_t.zero_()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        c.add_on_device(src, nxt, act, rew, None, carry=src)
    ref = ReplayBuffer(cap, od, ad, dev)
    first = torch.randn((n, od), generator=g, device=dev)
    src.copy_(first)
    for step in range(12):
        nx = torch.randn((n, od), generator=g, device=dev)
        nxt.copy_(nx); act.uniform_(-1, 1, generator=g); rew.uniform_(-5, 0, generator=g)
        ref.add(src.clone(), nx, act, rew, torch.zeros(n, device=dev))
        graph.replay()
        assert torch.equal(src, nx)
    assert torch.equal(ref.obs, c.obs) and torch.equal(ref.next_obs, c.next_obs) and torch.equal(ref.rew, c.rew)
    assert int(c.pos_t.item()) == ref.pos and c.size_t.item() == len(ref)
    # sampling: rows come from the filled part only, (obs, act, next_obs, rew, done) of ONE stored row each, uniform,
    # reproducible from (seed, stream position) and different from call to call
    d = ReplayBuffer(5000, od, ad, dev)
    rows = 1234
    tag = torch.arange(rows, device=dev, dtype=torch.float32)
    d.add(tag[:, None].expand(rows, od).contiguous(), (tag[:, None] + 0.5).expand(rows, od).contiguous(),
          (tag[:, None] * 0 + 0.25).expand(rows, ad).contiguous(), -tag, torch.zeros(rows, device=dev))
    d.sample_seed = 7
    o, ac, no, r, dn = d.sample_on_device(200000)
    idx = o[:, 0]
    assert idx.min() >= 0 and idx.max() <= rows - 1 and torch.equal(o, idx[:, None].expand_as(o))
    assert torch.equal(no, o + 0.5) and torch.equal(r, -idx) and (ac == 0.25).all() and (dn == 0).all()
    counts = torch.bincount(idx.long(), minlength=rows).double()
    assert counts.min() > 0 and abs(counts.mean().item() - 200000 / rows) < 1e-9
    chi2 = float(((counts - counts.mean()) ** 2 / counts.mean()).sum())
    assert 0.8 * rows < chi2 < 1.25 * rows, chi2                    # chi-square with 1,233 degrees of freedom: 1,233 +- 50
    assert int(d.draws_t.item()) == 200000
    o2 = d.sample_on_device(1000)[0]
    d.draws_t.fill_(200000)
    o3 = d.sample_on_device(1000)[0]
    d.draws_t.zero_()
    o4 = d.sample_on_device(1000)[0]
    assert torch.equal(o2, o3) and torch.equal(o4, o[:1000]) and not torch.equal(o2, o4)
