"""Minimal space containers with what the reference reads off a venv (util/make_agent.py:22-47,165-208;
envs/runners/adversarial_runner.py:63-64,73; models/multigrid_models.py:40-43,149-157): `.shape`, `.high`, `.low`, `.n`,
`.dtype`, `sample()` / `seed()` / `contains()` with gym 0.15.7's semantics (a space owns its own RandomState, unseeded until
`seed()` is called: the random adversary of DR + PLR draws its level-building actions with `action_space.sample()`).  The class
of a discrete space must be NAMED `Discrete` (util/__init__.py:135-139 checks `__class__.__name__`)."""
import numpy as np


class _Space(object):
    def __init__(self):
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def __contains__(self, x):
        return self.contains(x)


class Box(_Space):
    def __init__(self, low, high, shape, dtype='float32'):
        super().__init__()
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def sample(self):
        """gym 0.15.7 Box.sample for bounded spaces: uniform in [low, high], high + 1 exclusive for integer dtypes."""
        high = self.high if self.dtype.kind == 'f' else self.high.astype('int64') + 1
        s = self.np_random.uniform(low=self.low, high=high, size=self.shape)
        return (np.floor(s) if self.dtype.kind != 'f' else s).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def __repr__(self):
        return 'Box%s' % (self.shape,)


class Discrete(_Space):
    def __init__(self, n):
        super().__init__()
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return self.np_random.randint(self.n)

    def contains(self, x):
        try:
            v = int(x)
        except (TypeError, ValueError):
            return False
        return v == x and 0 <= v < self.n

    def __repr__(self):
        return 'Discrete(%d)' % self.n
