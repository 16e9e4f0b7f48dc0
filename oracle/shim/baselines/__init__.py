"""Stand-in for the handful of openai/baselines names envs/wrappers/* import. TEST INFRASTRUCTURE ONLY."""
from . import logger  # noqa: F401
