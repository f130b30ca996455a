"""Minimal `gym.spaces` stand-ins.

The reference declares `spaces.Box(-1e16, 1e16, shape=(5,1))` and `spaces.Discrete(3)`
(/root/reference/basilisk_env/envs/leoPowerAttitudeEnvironment.py:43-53).  Neither `gym` nor
`gymnasium` is installed in the build image, so the package carries these two classes; when a real
`gym`/`gymnasium` is importable its classes are used instead (see `basilisk_env_b200.registration`)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else np.asarray(low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else np.asarray(high, dtype=self.dtype)
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and np.allclose(self.low, other.low) and np.allclose(self.high, other.high)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return int(self._rng.randint(self.n))

    def contains(self, x):
        try:
            xi = int(x)
        except (TypeError, ValueError):
            return False
        return xi == x and 0 <= xi < self.n

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n


def gym_env_base():
    """`gym.Env` when a gym is importable (the reference declares `class leoPowerAttEnv(gym.Env)`,
    /root/reference/basilisk_env/envs/leoPowerAttitudeEnvironment.py:14, so wrappers that isinstance-check accept the
    env), else `object`: gym / gymnasium are not installed in the build image."""
    import importlib
    for modname in ("gym", "gymnasium"):
        try:
            return importlib.import_module(modname).Env
        except Exception:       # not installed (or a broken install): the env classes stand alone
            continue
    return object
