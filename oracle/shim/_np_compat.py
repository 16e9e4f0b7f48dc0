"""numpy alias shim for the oracle harness (TEST INFRASTRUCTURE ONLY).

The reference uses aliases removed in numpy>=1.24 (np.float, np.int, np.bool, np.NINF):
level_replay/level_sampler.py:79-83,94,138; envs/multigrid/multigrid.py:227,288,1004,1073;
envs/multigrid/adversarial.py:409; envs/runners/adversarial_runner.py:418-419,494,530.
"""
import numpy as np

for _n, _v in (("float", float), ("int", int), ("bool", bool), ("NINF", -np.inf)):
    if not hasattr(np, _n):
        setattr(np, _n, _v)

import _stubs  # noqa: E402  permissive stand-ins for the Box2D / plotting / display packages (never used on the MultiGrid path)

_stubs.install()
