"""Run the reference's OWN Python modules in-process (TEST INFRASTRUCTURE ONLY).

Puts `oracle/shim` (stand-ins for the un-installed third-party gym / gym_minigrid /
baselines) and `/root/reference` on sys.path so that envs/multigrid/*.py,
envs/wrappers/*.py, level_replay/*.py and algos/storage.py import UNMODIFIED.  Only
`oracle/gen_golden*.py`, the `-m "not gpu"` reference-diff tests, the GPU tests that run the reference's own runner /
train.py against the drop-in, and bench.py's Python CPU-baseline leg use it.  The tree is /root/reference in the
authoring container and the unmodified staged copy baseline/_ref/reference (oracle/stage_reference.py; git-ignored,
travels with the gpurun snapshot) on the GPU box.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, 'shim')
STAGED = os.path.join(os.path.dirname(HERE), 'baseline', '_ref', 'reference')
REFERENCE = os.environ.get('DCD_REFERENCE', '/root/reference')
if not os.path.isdir(os.path.join(REFERENCE, 'envs', 'multigrid')) and os.path.isdir(os.path.join(STAGED, 'envs', 'multigrid')):
    REFERENCE = STAGED


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
