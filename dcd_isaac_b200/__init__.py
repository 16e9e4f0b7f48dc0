"""dcd_isaac_b200 -- B200-native MultiGrid adversarial env + PLR hot path (drop-in for linjiw/dcd-isaac's
vectorised env / LevelSampler / LevelStore boundary).  CUDA only: see _lib.py."""
from . import _lib  # noqa: F401
from .registry import ENV_SPECS, env_spec  # noqa: F401


def __getattr__(name):
    if name in ('CudaAdversarialVecEnv', 'CudaMazeVecEnv', 'create_parallel_env'):
        from . import vec_env
        return getattr(vec_env, name)
    raise AttributeError(name)
