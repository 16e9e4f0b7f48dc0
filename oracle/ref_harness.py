"""Run the reference's OWN Python modules in-process (TEST INFRASTRUCTURE ONLY).

Puts `oracle/shim` (stand-ins for the un-installed third-party gym / gym_minigrid /
baselines) and `/root/reference` on sys.path so that envs/multigrid/*.py,
envs/wrappers/*.py, level_replay/*.py and algos/storage.py import UNMODIFIED.  Only
`oracle/gen_golden.py` and the `-m "not gpu"` reference-diff tests use it; it needs
/root/reference and therefore never runs on the GPU box.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, 'shim')
REFERENCE = os.environ.get('DCD_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE, 'envs', 'multigrid'))


def activate():
    """Make `import envs.multigrid.adversarial` etc. resolve to the reference."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE)
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    for p in (REFERENCE, SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import _np_compat  # noqa: F401
    return REFERENCE


def make_env(env_name, **kwargs):
    """gym_make through the reference's registry (wraps in its TimeLimit)."""
    activate()
    import envs.multigrid.adversarial  # noqa: F401  registers the MultiGrid-* ids
    from envs.registration import make as gym_make
    return gym_make(env_name, **kwargs)
