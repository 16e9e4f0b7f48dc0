"""Run the reference's own scripts (train.py, eval.py) on the B200 path WITHOUT editing them.

    python -m dcd_isaac_b200.dropin /path/to/reference/train.py --env_name MultiGrid-... [the reference's flags]
    python -m dcd_isaac_b200.dropin /path/to/reference/eval.py --benchmark maze ...

`install()` applies the import swaps of INTEGRATION.md section 2 to the already-importable reference packages, then the
script runs under runpy as `__main__`:

  util.create_parallel_env            -> dcd_isaac_b200.vec_env.create_parallel_env      (util/__init__.py:184-220; train.py:72)
  level_replay.LevelSampler/LevelStore -> dcd_isaac_b200.level_sampler / level_store     (adversarial_runner.py:14,101-124)
  eval.py's per-test-env vector env    -> dcd_isaac_b200.eval_envs.make_eval_venv         (eval.py:176-242)
  algos.storage.RolloutStorage         -> dcd_isaac_b200.storage.RolloutStorage (optional, `storage=True`)

The evaluator builds its envs from closures (`Evaluator.make_env` -> `gym_make(env_name)`, then
`ParallelAdversarialVecEnv(make_fn, is_eval=True)` + `VecMonitor` + `VecPreprocessImageWrapper`, eval.py:176-242), so the
swap sits at the three names those closures resolve at call time: `envs.registration.make` hands back a request object
for the MultiGrid evaluation ids, the `ParallelAdversarialVecEnv` name builds the CUDA vector env from it, and the two
wrapper names pass a CUDA vector env through untouched (it already folds monitor + preprocessing into the step kernel).
Everything that is not a MultiGrid env keeps the reference's own code path.
"""
import os
import runpy
import sys

_installed = {}


class EvalEnvRequest(object):
    """What `gym_make(env_name)` returns for a MultiGrid evaluation env under the drop-in: the vector env is built
    later, from N of these, by the ParallelAdversarialVecEnv stand-in."""

    def __init__(self, env_name, kwargs):
        self.env_name, self.kwargs = env_name, dict(kwargs)
        self.full_obs = False


def _is_cuda_venv(venv):
    from .mst_maze import CudaWideMSTMazeVecEnv
    from .vec_env import CudaAdversarialVecEnv
    return isinstance(venv, (CudaAdversarialVecEnv, CudaWideMSTMazeVecEnv, _OnDevice))


class _OnDevice(object):
    """The `device=` argument of VecPreprocessImageWrapper (obs_wrappers.py:77-86,99-100): eval.py's __main__ evaluates on the
    CPU (eval.py:385), so the CUDA vector env's observation / reward tensors are moved to where the agent lives."""

    def __init__(self, venv, device):
        import torch
        self.venv, self._device = venv, torch.device(device)

    def __getattr__(self, name):
        return getattr(self.venv, name)

    def _move(self, obs):
        return {k: v.to(self._device) for k, v in obs.items()}

    def reset(self):
        return self._move(self.venv.reset())

    def step(self, action):
        obs, rew, done, infos = self.venv.step(action)
        return self._move(obs), rew.to(self._device), done, infos


def install(device='cuda:0', storage=False):
    """Patch the reference modules (must be importable: reference root + its third-party deps on sys.path)."""
    if _installed:
        return _installed
    import envs.registration as registration
    import envs.wrappers as wrappers
    import level_replay
    import util

    from . import eval_envs
    from .level_sampler import LevelSampler
    from .level_store import LevelStore
    from .vec_env import create_parallel_env as cuda_create_parallel_env

    ref = dict(create_parallel_env=util.create_parallel_env, make=registration.make,
               ParallelAdversarialVecEnv=wrappers.ParallelAdversarialVecEnv, VecMonitor=wrappers.VecMonitor,
               VecPreprocessImageWrapper=wrappers.VecPreprocessImageWrapper,
               MultiGridFullyObsWrapper=wrappers.MultiGridFullyObsWrapper)

    def create_parallel_env(args, adversary=True):
        if str(args.env_name).startswith('MultiGrid'):
            return cuda_create_parallel_env(args, adversary=adversary, device=device)
        return ref['create_parallel_env'](args, adversary=adversary)

    def make(env_id, **kwargs):
        if eval_envs.is_eval_env(env_id):
            return EvalEnvRequest(env_id, kwargs)
        return ref['make'](env_id, **kwargs)

    def full_obs_wrapper(env, *a, **k):
        if isinstance(env, EvalEnvRequest):
            env.full_obs = True
            return env
        return ref['MultiGridFullyObsWrapper'](env, *a, **k)

    def parallel_venv(env_fns, adversary=True, is_eval=False):
        first = env_fns[0]()
        if isinstance(first, EvalEnvRequest):
            return eval_envs.make_eval_venv(first.env_name, len(env_fns), device=device, full_obs=first.full_obs)
        if hasattr(first, 'close'):
            first.close()
        return ref['ParallelAdversarialVecEnv'](env_fns, adversary=adversary, is_eval=is_eval)

    def vec_monitor(venv, *a, **k):
        return venv if _is_cuda_venv(venv) else ref['VecMonitor'](venv, *a, **k)

    def vec_preprocess(venv, *a, **k):
        if not _is_cuda_venv(venv):
            return ref['VecPreprocessImageWrapper'](venv, *a, **k)
        import torch
        want = k.get('device')
        if want is not None and torch.device(want).type != 'cuda':
            return _OnDevice(venv, want)
        return venv

    util.create_parallel_env = create_parallel_env
    registration.make = make
    for mod in (wrappers,):
        mod.ParallelAdversarialVecEnv = parallel_venv
        mod.VecMonitor = vec_monitor
        mod.VecPreprocessImageWrapper = vec_preprocess
        mod.MultiGridFullyObsWrapper = full_obs_wrapper
    level_replay.LevelSampler = LevelSampler
    level_replay.LevelStore = LevelStore
    # modules that bound the names before install() ran (e.g. a previously imported runner / eval module)
    for name in ('envs.runners.adversarial_runner', 'eval', 'train', '__main__'):
        m = sys.modules.get(name)
        if m is None:
            continue
        if getattr(m, 'LevelSampler', None) is not None and name != '__main__':
            m.LevelSampler, m.LevelStore = LevelSampler, LevelStore
        if name == 'eval':
            m.gym_make, m.ParallelAdversarialVecEnv, m.VecMonitor = make, parallel_venv, vec_monitor
            m.VecPreprocessImageWrapper, m.MultiGridFullyObsWrapper = vec_preprocess, full_obs_wrapper
            m.create_parallel_env = create_parallel_env
    if storage:
        import algos
        import algos.storage as ref_storage
        from .storage import RolloutStorage
        ref['RolloutStorage'] = ref_storage.RolloutStorage
        ref_storage.RolloutStorage = RolloutStorage
        algos.RolloutStorage = RolloutStorage
        ma = sys.modules.get('util.make_agent')
        if ma is not None and hasattr(ma, 'RolloutStorage'):
            ma.RolloutStorage = RolloutStorage
    _installed.update(ref)
    _installed['device'] = device
    return _installed


def uninstall():
    """Restore the reference's own names (tests)."""
    if not _installed:
        return
    import envs.registration as registration
    import envs.wrappers as wrappers
    import level_replay
    import util
    from level_replay.level_sampler import LevelSampler
    from level_replay.level_store import LevelStore
    util.create_parallel_env = _installed['create_parallel_env']
    registration.make = _installed['make']
    for k in ('ParallelAdversarialVecEnv', 'VecMonitor', 'VecPreprocessImageWrapper', 'MultiGridFullyObsWrapper'):
        setattr(wrappers, k, _installed[k])
    level_replay.LevelSampler, level_replay.LevelStore = LevelSampler, LevelStore
    m = sys.modules.get('envs.runners.adversarial_runner')
    if m is not None:
        m.LevelSampler, m.LevelStore = LevelSampler, LevelStore
    if 'RolloutStorage' in _installed:
        import algos
        import algos.storage as ref_storage
        ref_storage.RolloutStorage = _installed['RolloutStorage']
        algos.RolloutStorage = _installed['RolloutStorage']
        ma = sys.modules.get('util.make_agent')
        if ma is not None:
            ma.RolloutStorage = _installed['RolloutStorage']
    _installed.clear()


def run_script(script, argv, device='cuda:0', storage=False):
    """runpy the reference script as __main__ with the swaps installed."""
    script = os.path.abspath(script)
    root = os.path.dirname(script)
    if root not in sys.path:
        sys.path.insert(0, root)
    install(device=device, storage=storage)
    old_argv = sys.argv
    sys.argv = [script] + list(argv)
    try:
        return runpy.run_path(script, run_name='__main__')
    finally:
        sys.argv = old_argv


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    dev = os.environ.get('MGPLR_DEVICE', 'cuda:0')
    run_script(sys.argv[1], sys.argv[2:], device=dev, storage=bool(int(os.environ.get('MGPLR_DROPIN_STORAGE', '0'))))


if __name__ == '__main__':
    main()
