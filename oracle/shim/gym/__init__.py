"""Stand-in for the subset of gym==0.15.7 the reference hot path touches.

TEST INFRASTRUCTURE ONLY (oracle). gym is an un-vendored third-party dependency of the
reference (requirements.txt:2) and is not installed in this image; this package restates
the published behaviour of the pieces the reference calls: the spaces containers,
Env/Wrapper attribute forwarding (how step_adversary/mutate_level/encoding reach the env
through envs/wrappers/time_limit.py) and utils.seeding.np_random (sha512 hash_seed ->
MT19937 init_by_array limbs).  Nothing under dcd_isaac_b200/ imports it.
"""
import _np_compat  # noqa: F401  numpy aliases the reference still uses

from . import error, logger, spaces, utils, envs  # noqa: F401
from .core import Env, Wrapper  # noqa: F401
from . import core  # noqa: F401
