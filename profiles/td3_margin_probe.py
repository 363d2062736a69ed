import sys; sys.path.insert(0,'/root/repo')
import torch
from multi_echelon_drl_b200 import register_envs as R
from multi_echelon_drl_b200.td3 import HgeTD3, TD3Config
from multi_echelon_drl_b200.vec_normalize import VecNormalize
for seed in (0, 1, 2):
    n = 512
    env = VecNormalize(R.make("BeerGameNormalRetailer-v0", n_envs=n, dtype="float32", seed=seed), clip_obs=100.0, clip_reward=1000.0)
    cfg = TD3Config(net_arch=(128,), batch_size=512, learning_rate=3e-4, tau=0.02, train_freq=10, gradient_steps=40,
                    action_noise_std=0.2, gamma=0.95, hge_rate_at_start=0.0, buffer_size=n * 100 * 8, learning_starts=n * 100, seed=seed)
    m = HgeTD3(env, None, cfg); m.learn(n * 100 * 14)
    print(seed, [round(x) for x in m.episode_returns])
