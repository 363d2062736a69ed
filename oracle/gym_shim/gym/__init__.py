"""Stand-in for the absent `gym==0.21` package (TEST INFRASTRUCTURE ONLY).

The reference's envs.py does `import gym; from gym.utils import seeding` and uses
gym.Env, gym.spaces.{Box,Discrete,MultiDiscrete} and seeding.np_random
(/root/reference/envs.py:11-12,117,269-277).  gym is not installed in this image and
cannot be fetched (no network), so the golden-vector generator imports this shim
instead.  It carries no simulation logic.
"""
from . import spaces  # noqa: F401
from .utils import seeding  # noqa: F401


class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def seed(self, seed=None):
        return [seed]

    def close(self):
        pass


class Space:
    pass
