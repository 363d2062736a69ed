"""Device VecNormalize vs a numpy restatement of SB3's VecNormalize/RunningMeanStd (Appendix D;
parity with SB3 itself is unpinned: SB3 is not installed).  Tolerance 1e-6 relative on normalised
values (float32 outputs), 1e-10 on the float64 running statistics."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


class RMS:
    def __init__(self, shape):
        self.mean, self.var, self.count = np.zeros(shape), np.ones(shape), 1e-4

    def update(self, x):
        bm, bv, bc = x.mean(0), x.var(0), x.shape[0]
        d = bm - self.mean
        tot = self.count + bc
        self.mean = self.mean + d * bc / tot
        self.var = (self.var * self.count + bv * bc + d * d * self.count * bc / tot) / tot
        self.count = tot


def test_vecnormalize_matches_numpy_restatement(tmp_path):
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    n = 2048
    env = VecNormalize(R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, seed=4, dtype="float32"),
                       clip_obs=100.0, clip_reward=1000.0)
    obs_rms, ret_rms, ret = RMS((28,)), RMS(()), np.zeros(n)
    gamma, eps = 0.99, 1e-8
    nobs = env.reset()
    raw = env.get_original_obs().cpu().numpy().astype(np.float64)
    obs_rms.update(raw)
    np.testing.assert_allclose(nobs.cpu().numpy(), np.clip((raw - obs_rms.mean) / np.sqrt(obs_rms.var + eps), -100, 100),
                               rtol=1e-5, atol=1e-5)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for t in range(205):                                    # crosses two episode ends (auto-reset)
        a = torch.rand((n, 4), device="cuda", generator=g) * 16
        nobs, nrew, dones, infos = env.step(a)
        raw = env.get_original_obs().cpu().numpy().astype(np.float64)
        r = env.get_original_reward().cpu().numpy().astype(np.float64)
        obs_rms.update(raw)
        ret = ret * gamma + r
        ret_rms.update(ret)
        np.testing.assert_allclose(nobs.cpu().numpy(), np.clip((raw - obs_rms.mean) / np.sqrt(obs_rms.var + eps), -100, 100),
                                   rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(nrew.cpu().numpy(), np.clip(r / np.sqrt(ret_rms.var + eps), -1000, 1000), rtol=1e-5, atol=1e-6)
        if dones.all():
            # SB3's VecNormalize.step_wait normalises infos[i]["terminal_observation"] with the statistics as they
            # are AFTER this step's update (the update itself saw the reset observation, not the terminal one)
            tr = infos.terminal_observation.cpu().numpy()
            raw_term = env.venv._term_obs.cpu().numpy().astype(np.float64)
            np.testing.assert_allclose(tr, np.clip((raw_term - obs_rms.mean) / np.sqrt(obs_rms.var + eps), -100, 100),
                                       rtol=1e-5, atol=1e-5)
            assert np.abs(raw_term - raw).max() > 0          # it really is another observation than the returned one
            ret[:] = 0
        np.testing.assert_allclose(env.returns.cpu().numpy(), ret, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(env.obs_rms["mean"].cpu().numpy(), obs_rms.mean, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(env.obs_rms["var"].cpu().numpy(), obs_rms.var, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(env.ret_rms["var"].item(), ret_rms.var, rtol=1e-8)
    assert abs(env.obs_rms["count"].item() - obs_rms.count) < 1e-6
    # frozen statistics (evaluation env): no update, same normalisation; save / load round trip
    env.training = False
    before = env._obs_rms.clone()
    env.step(torch.zeros((n, 4), device="cuda"))
    assert torch.equal(before, env._obs_rms)
    path = str(tmp_path / "vecnormalize.pkl")
    env.save(path)
    env2 = VecNormalize.load(path, R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, seed=5, dtype="float32"))
    assert torch.equal(env2._obs_rms, env._obs_rms) and env2.clip_obs == 100.0
    x = torch.rand((n, 28), device="cuda") * 30
    assert torch.equal(env.normalize_obs(x), env2.normalize_obs(x))


def test_vecnormalize_small_obs_dim_and_reward_only():
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    n = 333
    env = VecNormalize(R.make("BeerGameNormalRetailer-v0", n_envs=n, seed=1, dtype="float32"), norm_obs=False,
                       clip_reward=1000.0, gamma=0.95)
    o = env.reset()
    assert o.shape == (n, 7)
    ret_rms, ret = RMS(()), np.zeros(n)
    for t in range(20):
        o, nrew, _, _ = env.step(torch.full((n, 1), 9.0, device="cuda"))
        assert torch.equal(o, env.get_original_obs())                 # observations untouched
        r = env.get_original_reward().cpu().numpy().astype(np.float64)
        ret = ret * 0.95 + r
        ret_rms.update(ret)
        np.testing.assert_allclose(nrew.cpu().numpy(), np.clip(r / np.sqrt(ret_rms.var + 1e-8), -1000, 1000), rtol=1e-5, atol=1e-6)
    env7 = VecNormalize(R.make("BeerGameNormalRetailer-v0", n_envs=n, seed=1, dtype="float32"), clip_obs=100.0)
    o = env7.reset()
    raw = env7.get_original_obs().cpu().numpy().astype(np.float64)
    rms = RMS((7,)); rms.update(raw)
    np.testing.assert_allclose(o.cpu().numpy(), np.clip((raw - rms.mean) / np.sqrt(rms.var + 1e-8), -100, 100), rtol=1e-5, atol=1e-5)


def test_vecnormalize_on_56_dim_observations_of_an_eight_node_tree():
    """Widest observation the library produces (7 numbers x 8 facilities): running statistics and normalisation
    agree with the numpy restatement."""
    from multi_echelon_drl_b200.spec import tree_spec
    from multi_echelon_drl_b200.vec_env import SupplyNetVecEnv
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    names = ["hub", "r1", "r2", "r3", "r4", "r5", "factory", "annex"]
    spec = tree_spec([dict(name=nm, target=t, demand=d) for nm, t, d in zip(
        names, [60, 19, 14, 10, 8, 6, 40, 12], [None, (0, 8), (0, 6), (0, 4), (0, 3), (0, 2), None, (0, 5)])],
        [("ext", "factory", 1, 2), ("factory", "annex", 1, 3), ("factory", "hub", 2, 2), ("hub", "r3", 2, 1), ("hub", "r1", 2, 2),
         ("hub", "r5", 3, 1), ("hub", "r2", 1, 1), ("hub", "r4", 1, 2)], names, True)
    n = 1000
    env = VecNormalize(SupplyNetVecEnv(spec, n, seed=2, dtype="float32"), clip_obs=100.0, clip_reward=1000.0)
    obs_rms, ret_rms, ret = RMS((56,)), RMS(()), np.zeros(n)
    nobs = env.reset()
    raw = env.get_original_obs().cpu().numpy().astype(np.float64)
    obs_rms.update(raw)
    np.testing.assert_allclose(nobs.cpu().numpy(), np.clip((raw - obs_rms.mean) / np.sqrt(obs_rms.var + 1e-8), -100, 100),
                               rtol=1e-5, atol=1e-5)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for t in range(40):
        nobs, nrew, dones, infos = env.step(torch.rand((n, 8), device="cuda", generator=g) * 16)
        raw = env.get_original_obs().cpu().numpy().astype(np.float64)
        r = env.get_original_reward().cpu().numpy().astype(np.float64)
        obs_rms.update(raw)
        ret = ret * 0.99 + r
        ret_rms.update(ret)
        np.testing.assert_allclose(nobs.cpu().numpy(), np.clip((raw - obs_rms.mean) / np.sqrt(obs_rms.var + 1e-8), -100, 100),
                                   rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(nrew.cpu().numpy(), np.clip(r / np.sqrt(ret_rms.var + 1e-8), -1000, 1000), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(env.obs_rms["mean"].cpu().numpy(), obs_rms.mean, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(env.obs_rms["var"].cpu().numpy(), obs_rms.var, rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("use_graph", [False, True])
def test_running_statistics_equal_numpy_on_the_concatenated_data(use_graph):
    """Pins the running statistics to something that is not our own restatement: after any number of updates
    RunningMeanStd's (mean, var, count) must be the moments of ALL samples seen so far, taken by numpy over their
    concatenation, combined with the initial pseudo-sample (mean 0, var 1, weight 1e-4) - Chan et al.'s parallel
    variance is exact, so however the batches were merged (eager kernels, fused into the step kernel, one-kernel
    mode) the result is the same closed form."""
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.vec_normalize import VecNormalize
    n = 1536
    env = VecNormalize(R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=n, seed=11, dtype="float32", use_graph=use_graph),
                       clip_obs=100.0, clip_reward=1000.0, gamma=0.95)
    seen_obs, seen_ret, ret = [], [], np.zeros(n)
    env.reset()
    seen_obs.append(env.get_original_obs().cpu().numpy().astype(np.float64))
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    for t in range(130):                                       # crosses one auto-reset
        a = torch.rand((n, 4), device="cuda", generator=g) * 16
        _, _, dones, _ = env.step(a)
        seen_obs.append(env.get_original_obs().cpu().numpy().astype(np.float64))
        ret = ret * 0.95 + env.get_original_reward().cpu().numpy().astype(np.float64)
        seen_ret.append(ret.copy())
        if dones.all():
            ret[:] = 0

    def closed_form(samples):
        x = np.concatenate(samples, axis=0)
        w0, cnt = 1e-4, x.shape[0]
        mean = x.sum(0) / (cnt + w0)                           # prior mean 0
        m2 = ((x - mean) ** 2).sum(0) + w0 * (1.0 + mean ** 2)  # prior: variance 1 about mean 0, moved to the new mean
        return mean, m2 / (cnt + w0), cnt + w0

    mean, var, cnt = closed_form(seen_obs)
    np.testing.assert_allclose(env.obs_rms["mean"].cpu().numpy(), mean, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(env.obs_rms["var"].cpu().numpy(), var, rtol=1e-8, atol=1e-9)
    assert abs(env.obs_rms["count"].item() - cnt) < 1e-6
    rmean, rvar, rcnt = closed_form([r[:, None] for r in seen_ret])
    np.testing.assert_allclose(env.ret_rms["mean"].item(), rmean[0], rtol=1e-9)
    np.testing.assert_allclose(env.ret_rms["var"].item(), rvar[0], rtol=1e-8)
    assert abs(env.ret_rms["count"].item() - rcnt) < 1e-6
