import numpy as np


def np_random(seed=None):
    rng = np.random.RandomState(seed if seed is not None else None)
    return rng, seed
