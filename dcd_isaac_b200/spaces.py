"""Minimal space containers with the attributes the reference reads off a venv
(util/make_agent.py:22-47,165-208; envs/runners/adversarial_runner.py:63-64,73; models/multigrid_models.py:40-43):
`.shape`, `.high`, `.low`, `.n`, `__getitem__`, `.items()`.  The class of a discrete space must be NAMED
`Discrete` (util/__init__.py:135-139 checks `__class__.__name__`)."""
import numpy as np


class Box(object):
    def __init__(self, low, high, shape, dtype='float32'):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def __repr__(self):
        return 'Box%s' % (self.shape,)


class Discrete(object):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def __repr__(self):
        return 'Discrete(%d)' % self.n
