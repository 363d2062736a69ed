"""Minimal action/observation space objects (gym 0.21 duck types).  If gymnasium or gym is
importable their classes are used instead so that SB3-style consumers get the real thing."""
import numpy as np

try:                                      # pragma: no cover - not installed in this image
    from gymnasium.spaces import Box, Discrete, MultiDiscrete   # type: ignore
except Exception:                         # noqa: BLE001
    try:                                  # pragma: no cover
        from gym.spaces import Box, Discrete, MultiDiscrete      # type: ignore
    except Exception:                     # noqa: BLE001

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.shape = tuple(shape) if shape is not None else np.shape(low)
                self.low = np.full(self.shape, low, dtype=dtype)
                self.high = np.full(self.shape, high, dtype=dtype)
                self.dtype = np.dtype(dtype)

            def sample(self):
                return np.random.uniform(self.low, self.high).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        class Discrete:
            def __init__(self, n):
                self.n = int(n)
                self.shape = ()
                self.dtype = np.dtype(np.int64)

            def sample(self):
                return int(np.random.randint(self.n))

            def contains(self, x):
                return 0 <= int(x) < self.n

            def __repr__(self):
                return f"Discrete({self.n})"

        class MultiDiscrete:
            def __init__(self, nvec):
                self.nvec = np.asarray(nvec, dtype=np.int64)
                self.shape = self.nvec.shape
                self.dtype = np.dtype(np.int64)

            def sample(self):
                return (np.random.random_sample(self.nvec.shape) * self.nvec).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

            def __repr__(self):
                return f"MultiDiscrete({self.nvec.tolist()})"
