"""Empty stand-in: envs/multigrid/window.py:13-18 sys.exit()s when matplotlib is missing. TEST INFRASTRUCTURE ONLY."""
from . import pyplot  # noqa: F401
