"""ctypes front-end of oracle/c/mg_oracle.c (TEST INFRASTRUCTURE ONLY -- see that file's header).

`OracleEnv` is one reference-semantics MultiGrid adversarial env; `OracleVecEnv` folds the
reference wrapper chain (TimeLimit, worker auto-reset, VecMonitor, VecPreprocessImageWrapper)
for N envs the way envs/wrappers/parallel_wrappers.py + util/__init__.py:184-220 compose them.
"""
import ctypes as C
import hashlib
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# env-name -> constructor arguments, restated from envs/multigrid/adversarial.py:584-786
ENV_SPECS = {
    'MultiGrid-Adversarial-v0': dict(n_clutter=50, size=15, goal_last=0, see_through=1, max_steps=250, tl=250),
    'MultiGrid-GoalLastAdversarial-v0': dict(n_clutter=50, size=15, goal_last=1, see_through=1, max_steps=250, tl=250),
    'MultiGrid-GoalLastOpaqueWallsAdversarial-v0': dict(n_clutter=50, size=15, goal_last=1, see_through=0, max_steps=250, tl=250),
    'MultiGrid-GoalLastFewerBlocksAdversarial-v0': dict(n_clutter=25, size=15, goal_last=1, see_through=1, max_steps=250, tl=250),
    'MultiGrid-GoalLastFewerBlocksAdversarial-EditWN-v0': dict(n_clutter=25, size=15, goal_last=1, see_through=1, max_steps=250, tl=250, editor=2),
    'MultiGrid-GoalLastFewerBlocksAdversarial-EditWNG-v0': dict(n_clutter=25, size=15, goal_last=1, see_through=1, max_steps=250, tl=250, editor=3),
    'MultiGrid-GoalLastVariableBlocksAdversarialEnv-v0': dict(n_clutter=60, size=15, goal_last=1, see_through=1, max_steps=250, tl=250, resample=1),
    'MultiGrid-GoalLastVariableBlocksAdversarialEnv-Edit-v0': dict(n_clutter=60, size=15, goal_last=1, see_through=1, max_steps=250, tl=250, resample=1, editor=3),
    'MultiGrid-GoalLastEmptyAdversarialEnv-Edit-v0': dict(n_clutter=0, size=15, goal_last=1, see_through=1, max_steps=250, tl=250, editor=3),
    'MultiGrid-GoalLastFewerBlocksOpaqueWallsAdversarial-v0': dict(n_clutter=25, size=15, goal_last=1, see_through=0, max_steps=250, tl=250),
    'MultiGrid-MiniGoalLastAdversarial-v0': dict(n_clutter=7, size=6, goal_last=1, see_through=1, max_steps=50, tl=50),
    'MultiGrid-GoalLastAdversarialEnv30-v0': dict(n_clutter=30, size=15, goal_last=1, see_through=1, max_steps=250, tl=50),
    'MultiGrid-GoalLastAdversarialEnv60-v0': dict(n_clutter=60, size=15, goal_last=1, see_through=1, max_steps=250, tl=50),
}


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        'W', 'max_steps', 'max_episode_steps', 'see_through', 'n_clutter', 'resample_n_clutter',
        'choose_goal_last', 'fixed_environment', 'n_editor_actions')]


def build(force=False):
    so = os.path.join(HERE, 'libmg_oracle.so')
    src = os.path.join(HERE, 'c', 'mg_oracle.c')
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(['make', '-C', HERE, 'libmg_oracle.so'], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.mgo_sizeof_env.restype = C.c_int
        L.mgo_rollout_batch.restype = C.c_long
        L.mgo_rng_next.restype = C.c_uint32
        _LIB = L
    return _LIB


def seed_limbs(seed):
    """gym 0.15.7 seeding: _int_list_from_bigint(hash_seed(seed)) (oracle/shim/gym/utils/seeding.py)."""
    h = hashlib.sha512(str(int(seed) % 2 ** 64).encode('utf8')).digest()[:8]
    lo, hi = struct.unpack('<2I', h)
    if hi:
        return [lo, hi]
    return [lo] if lo else [0]


def make_cfg(W=15, max_steps=250, max_episode_steps=250, see_through=True, n_clutter=50,
             resample_n_clutter=False, choose_goal_last=True, fixed_environment=False, n_editor_actions=4):
    return Cfg(W, max_steps, max_episode_steps, int(see_through), n_clutter, int(resample_n_clutter),
               int(choose_goal_last), int(fixed_environment), n_editor_actions)


def cfg_from_name(env_name, fixed_environment=False):
    s = ENV_SPECS[env_name]
    return make_cfg(W=s['size'], max_steps=s['max_steps'], max_episode_steps=s['tl'],
                    see_through=s['see_through'], n_clutter=s['n_clutter'],
                    resample_n_clutter=s.get('resample', 0), choose_goal_last=s['goal_last'],
                    fixed_environment=fixed_environment, n_editor_actions=s.get('editor', 4))


class OracleBatch(object):
    """N oracle envs in one contiguous C array."""

    STATE_FIELDS = ('has_agent', 'ax', 'ay', 'adir', 'gx', 'gy', 'sx', 'sy', 'sdir', 'step_count',
                    'elapsed', 'done_flag', 'adv_step', 'adv_max', 'n_clutter_sampled',
                    'n_clutter_placed', 'dist', 'passable', 'spl', 'ep_len', 'rng_words', 'error')

    def __init__(self, cfg, n):
        self.L = lib()
        self.cfg = cfg
        self.n = n
        self.W = cfg.W
        self.sz = self.L.mgo_sizeof_env()
        self.buf = C.create_string_buffer(self.sz * n)
        self.base = C.addressof(self.buf)
        for i in range(n):
            self.L.mgo_init(self._p(i), C.byref(cfg))

    def _p(self, i):
        return C.c_void_p(self.base + i * self.sz)

    def seed(self, i, seed):
        limbs = seed_limbs(seed)
        arr = (C.c_uint32 * 2)(*(limbs + [0])[:2])
        self.L.mgo_seed(self._p(i), arr, len(limbs))

    def reset(self, i):
        self.L.mgo_reset(self._p(i))

    def reset_agent(self, i):
        return self.L.mgo_reset_agent(self._p(i))

    def reset_random(self, i, n_walls=-1):
        return self.L.mgo_reset_random(self._p(i), n_walls)

    def step_adversary(self, i, loc):
        err = C.c_int(0)
        done = self.L.mgo_step_adversary(self._p(i), int(loc), C.byref(err))
        if err.value:
            raise ValueError('Position passed to step_adversary is outside the grid.')
        return bool(done)

    def reset_to_encoding(self, i, enc):
        enc = np.ascontiguousarray(enc, dtype=np.uint8)
        return self.L.mgo_reset_to_encoding(self._p(i), enc.ctypes.data_as(C.c_void_p))

    def reset_to_actions(self, i, locs):
        a = np.ascontiguousarray(locs, dtype=np.int32)
        return self.L.mgo_reset_to_actions(self._p(i), a.ctypes.data_as(C.c_void_p), len(a))

    def mutate(self, i, locs, ops, goal_choice=0, agent_choice=0):
        l = np.ascontiguousarray(locs, dtype=np.int32)
        o = np.ascontiguousarray(ops, dtype=np.int32)
        need = (C.c_int * 2)()
        nfree = (C.c_int * 2)()
        rc = self.L.mgo_mutate(self._p(i), l.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), len(l),
                               int(goal_choice), int(agent_choice), need, nfree)
        return rc, list(need), list(nfree)

    def encode(self, i):
        out = np.zeros((self.W, self.W, 3), np.uint8)
        self.L.mgo_encode(self._p(i), out.ctypes.data_as(C.c_void_p))
        return out

    def gen_obs(self, i):
        out = np.zeros((5, 5, 3), np.uint8)
        self.L.mgo_gen_obs(self._p(i), out.ctypes.data_as(C.c_void_p))
        return out

    def step(self, i, action):
        r = C.c_double(0)
        done = self.L.mgo_step(self._p(i), int(action), C.byref(r))
        return r.value, bool(done)

    def step_env(self, i, action, reset_random=False, n_walls_resample=-1):
        """-> dict(flags, obs u8[5,5,3], dir, rew f32, trunc_obs, trunc_dir, ep_r, ep_l)"""
        obs = np.zeros((5, 5, 3), np.uint8)
        tr = np.zeros((5, 5, 3), np.uint8)
        d, td, l = C.c_int(0), C.c_int(0), C.c_int(0)
        r, er = C.c_float(0), C.c_float(0)
        f = self.L.mgo_step_env(self._p(i), int(action), int(reset_random), int(n_walls_resample),
                                obs.ctypes.data_as(C.c_void_p), C.byref(d), C.byref(r),
                                tr.ctypes.data_as(C.c_void_p), C.byref(td), C.byref(er), C.byref(l))
        return dict(flags=f, obs=obs, dir=d.value, rew=np.float32(r.value), trunc_obs=tr, trunc_dir=td.value,
                    ep_r=np.float32(er.value), ep_l=l.value)

    def state(self, i):
        out = (C.c_int * 22)()
        self.L.mgo_get_state(self._p(i), out)
        return dict(zip(self.STATE_FIELDS, list(out)))

    def cells(self, i):
        out = np.zeros((self.W, self.W), np.uint8)
        self.L.mgo_get_cells(self._p(i), out.ctypes.data_as(C.c_void_p))
        return out  # [y][x]

    def rng_next(self, i):
        return int(self.L.mgo_rng_next(self._p(i)))

    def rng_randint(self, i, lo, hi):
        return int(self.L.mgo_rng_randint(self._p(i), lo, hi))

    def rollout(self, actions, reset_random=False, want_obs=True, n_threads=1):
        """actions uint8 [T, N] -> (obs f32 [T,N,3,5,5] or None, rew f32 [T,N], flags u8 [T,N], episodes)"""
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        T, N = actions.shape
        assert N == self.n
        obs = np.empty((T, N, 3, 5, 5), np.float32) if want_obs else None
        rew = np.empty((T, N), np.float32)
        flags = np.empty((T, N), np.uint8)
        ep = self.L.mgo_rollout_batch(C.c_void_p(self.base), N, T, actions.ctypes.data_as(C.c_void_p),
                                      int(reset_random),
                                      obs.ctypes.data_as(C.c_void_p) if want_obs else None,
                                      rew.ctypes.data_as(C.c_void_p), flags.ctypes.data_as(C.c_void_p), int(n_threads))
        return obs, rew, flags, int(ep)


def preprocess(u8):
    """VecPreprocessImageWrapper: /10.0 in double, channels first, float32 (obs_wrappers.py:88-115)."""
    a = np.asarray(u8)
    a = a / 10.0
    if a.ndim == 4:
        a = a.transpose(0, 3, 1, 2)
    else:
        a = a.transpose(2, 0, 1)
    return a.astype(np.float32)
