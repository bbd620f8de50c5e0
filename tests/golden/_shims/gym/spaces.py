import numpy as np


class Space:
    def __init__(self, *a, **k):
        self.shape = k.get("shape", None)


class Box(Space):
    def __init__(self, low=-np.inf, high=np.inf, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self):
        return np.random.uniform(-1.0, 1.0, size=self.shape).astype(self.dtype)


class Discrete(Space):
    pass


class MultiDiscrete(Space):
    pass


class MultiBinary(Space):
    pass


class Tuple(Space):
    pass


class Dict(Space):
    pass
