"""CudaAdversarialVecEnv -- the reference's vectorised adversarial-env API on top of libmgplr.so.

It stands in for the whole stack `ParallelAdversarialVecEnv -> VecMonitor -> VecNormalize(ob=False) ->
VecPreprocessImageWrapper` that util.create_parallel_env builds (util/__init__.py:184-220): same method
names, argument meaning, return conventions and error behaviour as envs/wrappers/parallel_wrappers.py:232-460
and envs/wrappers/obs_wrappers.py:157-230, but the N environments live in HBM and every call is one or two
kernel launches.  Observations come back as float32 CUDA tensors already scaled by 1/10 and channels-first;
`done` is a host numpy bool array and `infos` a list of dicts, exactly what
envs/runners/adversarial_runner.py:497-588 consumes.

No CPU fallback: constructing this class without the CUDA library or a CUDA device raises.
"""
import ctypes as C
import hashlib
import struct
import time
import types

import numpy as np
import torch

from . import _lib
from ._lib import EnvConfig, StepOut, check, ptr
from .registry import EDITOR_ACTION_SPACES, env_spec
from .spaces import Box, Discrete

F_DONE, F_TRUNC_KEY, F_TRUNC_VAL, F_GOAL, F_ERROR = 1, 2, 4, 8, 16

_NO_INFO = types.MappingProxyType({})   # read-only stand-in for the `{}` info of an env to which nothing happened


class _ObsRow(object):
    """`info['truncated_obs']`: the observation dict of ONE env, `{k: batch[k][i]}`, sliced out of the step's batched
    truncated-observation tensors only when a key is read (a synchronized time-limit step would otherwise spend ~10 us per env
    on tensor indexing whether or not anybody looks).  Read-only mapping with the dict methods the runner and
    RolloutStorage.insert_truncated_obs use (adversarial_runner.py:546-549, algos/storage.py:188-193)."""
    __slots__ = ('_batch', '_i')

    def __init__(self, batch, i):
        self._batch, self._i = batch, i

    def __getitem__(self, k):
        return self._batch[k][self._i]

    def __iter__(self):
        return iter(self._batch)

    def __len__(self):
        return len(self._batch)

    def __contains__(self, k):
        return k in self._batch

    def keys(self):
        return self._batch.keys()

    def values(self):
        return [self._batch[k][self._i] for k in self._batch]

    def items(self):
        return [(k, self._batch[k][self._i]) for k in self._batch]

    def get(self, k, default=None):
        return self._batch[k][self._i] if k in self._batch else default


class LazyInfos(list):
    """The `infos` list of one vector step.  Almost every env's info is `{}` on almost every step, so the untouched
    entries all hold ONE read-only empty mapping and a real dict is only created for an env that finished / was truncated,
    or the first time the caller indexes an entry to write into it (`infos[i]['cliffhanger'] = True`,
    adversarial_runner.py:526-528).  Nothing mutable is shared: writing through the read-only mapping raises."""
    __slots__ = ()

    def __getitem__(self, i):
        v = list.__getitem__(self, i)
        if v is _NO_INFO:
            v = {}
            list.__setitem__(self, i, v)
        return v


def seed_limbs(seed):
    """gym 0.15.7 `seeding.np_random(seed)`: limbs of hash_seed(seed) fed to RandomState.seed([...])
    (reached from MultiGridEnv.seed, envs/multigrid/multigrid.py:465-468)."""
    h = hashlib.sha512(str(int(seed) % 2 ** 64).encode('utf8')).digest()[:8]
    lo, hi = struct.unpack('<2I', h)
    if hi:
        return (lo, hi, 2)
    return (lo, 0, 1)


class CudaAdversarialVecEnv(object):
    def __init__(self, env_name, num_envs, device='cuda:0', seed=None, fixed_environment=None, spec=None, full_obs=False,
                 host_rng_seed=None, **overrides):
        if spec is None:
            spec = env_spec(env_name, fixed_environment=fixed_environment, **overrides)
        self.env_name = env_name
        self.spec = spec
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.MgplrError('CudaAdversarialVecEnv needs a CUDA device (no CPU fallback)')
        if not torch.cuda.is_available():
            raise _lib.MgplrError('no CUDA device visible: the MultiGrid path has no CPU fallback')
        self.L = _lib.load()
        self.W = int(spec['size'])
        self.editor_actions = list(EDITOR_ACTION_SPACES[spec['editor_actions']])
        self.cfg = EnvConfig(self.W, 5, spec['max_steps'], spec['max_episode_steps'], int(spec['see_through_walls']),
                             spec['n_clutter'], int(spec['resample_n_clutter']), int(spec['choose_goal_last']),
                             int(spec['fixed_environment']), len(self.editor_actions))
        self.n_clutter = spec['n_clutter']
        self.resample_n_clutter = bool(spec['resample_n_clutter'])
        self.random_z_dim = 50
        self.adversary_max_steps = self.n_clutter + 2
        self.adversary_action_dim = (self.W - 2) ** 2
        h = C.c_void_p()
        torch.cuda.set_device(self.device)
        check(self.L.mgplr_venv_create(C.byref(self.cfg), self.num_envs, self.device.index or 0, C.byref(h)),
              'mgplr_venv_create')
        self.h = h
        self.closed = False
        self.tstart = time.time()
        # The env-side draws the reference makes from numpy's GLOBAL stream (random_z, adversarial.py:449-450; mutate_level,
        # :317-397; _resample_n_clutter, :151-156; the corridor mazes' goal, maze.py:153-156) happen inside its env
        # SUBPROCESSES, each with its own unseeded stream -- they never touch the trainer process's np.random, which the
        # level sampler's decisions and draws consume (level_sampler.py:611,616,674).  The drop-in keeps that separation: one
        # venv-owned RandomState, unseeded unless `host_rng_seed` / `host_rng.seed()` asks for repeatability.
        self.host_rng = np.random.RandomState(host_rng_seed)
        self.seed_values = [seed] * self.num_envs
        # spaces (multigrid.py:407-440, adversarial.py:126-145, obs_wrappers.py:118-155 transposes 'image')
        N = self.num_envs
        self.observation_space = {'image': Box(0, 255, (3, 5, 5), 'uint8'), 'direction': Box(0, 3, (1,), 'uint8')}
        self.action_space = Discrete(7)
        self.adversary_observation_space = {
            'image': Box(0, 255, (3, self.W, self.W), 'uint8'),
            'time_step': Box(0, self.adversary_max_steps, (1,), 'uint8'),
            'random_z': Box(0, 1.0, (self.random_z_dim,), 'float32')}
        self.adversary_action_space = Discrete(self.adversary_action_dim)
        # MultiGridFullyObsWrapper (util/__init__.py:175-178 with --use_global_critic / --use_global_policy): every agent
        # observation also carries 'full_obs' = the whole grid's encoding, channels first, unscaled
        self.full_obs = bool(full_obs)
        if self.full_obs:
            self.observation_space['full_obs'] = Box(0, 255, (3, self.W, self.W), 'uint8')
        self.processed_action_dim = 1
        # persistent small device buffers
        dev = self.device
        self._flags = torch.zeros(N, dtype=torch.uint8, device=dev)
        self._ep_r = torch.zeros(N, dtype=torch.float32, device=dev)
        self._ep_l = torch.zeros(N, dtype=torch.int32, device=dev)
        self._done_adv = torch.zeros(N, dtype=torch.uint8, device=dev)
        self._errors = torch.zeros(N, dtype=torch.int32, device=dev)
        # pinned host staging for the host-driven step (actions in; flags + done records out)
        self._h_action = torch.zeros(N, dtype=torch.int64).pin_memory()
        self._h_action_u8 = torch.zeros(N, dtype=torch.uint8).pin_memory()
        self._h_flags = torch.zeros(N, dtype=torch.uint8).pin_memory()
        self._h_done = torch.zeros(N * 16, dtype=torch.uint8).pin_memory()  # mgplr_done_record [N]
        self._h_ndone = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._h_done_np = self._h_done.numpy()
        self._h_flags_np = self._h_flags.numpy()
        self._h_ndone_np = self._h_ndone.numpy()
        self._done_dtype = np.dtype(_lib.DONE_DTYPE)
        self._step_out = None   # StepOut with the persistent destinations of step_env (built on first use)
        self._tr = None
        if seed is not None:
            self.set_seed([seed] * N)

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _assert_not_closed(self):
        assert not self.closed, 'Trying to operate on a CudaAdversarialVecEnv after calling close()'

    def _new_obs(self, n=None, u8=False):
        n = self.num_envs if n is None else n
        obs = {'image': torch.empty(n, 3, 5, 5, dtype=torch.float32, device=self.device),
               'direction': torch.empty(n, 1, dtype=torch.float32, device=self.device)}
        return obs

    def _add_full_obs(self, obs):
        """MultiGridFullyObsWrapper.agent_observation (multigrid_wrappers.py:30-44) for all envs: one launch."""
        if self.full_obs:
            full = torch.empty(self.num_envs, 3, self.W, self.W, dtype=torch.float32, device=self.device)
            check(self.L.mgplr_full_obs(self.h, ptr(full), self._stream()), 'mgplr_full_obs')
            obs['full_obs'] = full
        return obs

    def _out(self, obs=None, **kw):
        o = StepOut()
        if obs is not None:
            o.image = ptr(obs['image'])
            o.direction = ptr(obs['direction'])
        for k, v in kw.items():
            setattr(o, k, ptr(v))
        return o

    def _raise_errors(self):
        """Surface device-side error flags with the reference's exception types."""
        check(self.L.mgplr_get_errors(self.h, ptr(self._errors), 1, self._stream()), 'mgplr_get_errors')
        e = self._errors.cpu().numpy()
        if not e.any():
            return
        if (e & 4).any():
            raise ValueError('Position passed to step_adversary is outside the grid.')
        if (e & 2).any():
            raise ValueError('Trying to place agent at empty start position.')
        if (e & 1).any():
            raise RuntimeError('Rejection sampling failed in place_obj')  # gym.error.RetriesExceededError analogue

    def _adv_obs(self, image, time_step):
        # random_z: np.random.uniform(size=(50,)).astype(float32) per env (adversarial.py:449-450).  In the reference every
        # env subprocess draws it from its own, unseeded global numpy stream, so only the distribution is defined: small
        # batches draw from the venv's host_rng (seedable), large ones on the device (the host draw of N x 50 doubles was
        # 66 of the 82 ms of a 4 096-env PAIRED cycle).
        if self.num_envs <= 256:
            z = torch.from_numpy(self.host_rng.uniform(size=(self.num_envs, self.random_z_dim)).astype(np.float32)).to(self.device)
        else:
            z = torch.rand(self.num_envs, self.random_z_dim, dtype=torch.float32, device=self.device)
        return {'image': image, 'time_step': time_step, 'random_z': z}

    # ------------------------------------------------------------------ seeding
    def set_seed(self, seeds):
        """venv.set_seed(seeds) (parallel_wrappers.py:415-416): per-env MultiGridEnv.seed."""
        self._assert_not_closed()
        seeds = list(seeds)
        assert len(seeds) == self.num_envs
        limbs = np.zeros((self.num_envs, 2), np.uint32)
        cnt = np.zeros(self.num_envs, np.int32)
        for i, s in enumerate(seeds):
            limbs[i, 0], limbs[i, 1], cnt[i] = seed_limbs(s)
        check(self.L.mgplr_seed(self.h, ptr(limbs), ptr(cnt), None, self.num_envs, self._stream()), 'mgplr_seed')
        self.seed_values = seeds
        return [[s] for s in seeds]

    def seed(self, seed, index):
        """venv.seed(seed, index) (parallel_wrappers.py:268-270)."""
        self._assert_not_closed()
        lo, hi, c = seed_limbs(seed)
        limbs = np.array([[lo, hi]], np.uint32)
        cnt = np.array([c], np.int32)
        idx = np.array([index], np.int32)
        check(self.L.mgplr_seed(self.h, ptr(limbs), ptr(cnt), ptr(idx), 1, self._stream()), 'mgplr_seed')
        self.seed_values[index] = seed
        return [seed]

    def get_seed(self):
        return list(self.seed_values)

    # ------------------------------------------------------------------ adversary phase
    def reset(self):
        """venv.reset(): AdversarialEnv.reset on every env (adversarial.py:194-229)."""
        self._assert_not_closed()
        N, W = self.num_envs, self.W
        image = torch.empty(N, 3, W, W, dtype=torch.float32, device=self.device)
        ts = torch.empty(N, 1, dtype=torch.float32, device=self.device)
        check(self.L.mgplr_reset(self.h, ptr(image), ptr(ts), self._stream()), 'mgplr_reset')
        return self._adv_obs(image, ts)

    def step_adversary(self, action):
        """venv.step_adversary(action) (parallel_wrappers.py:288-297, obs_wrappers.py:183-190)."""
        self._assert_not_closed()
        N, W = self.num_envs, self.W
        a = torch.as_tensor(action)
        if a.device.type != 'cuda':
            a_host = a.reshape(-1).to(torch.int64)
            if a_host.numel() and int(a_host.max()) >= self.adversary_action_dim:
                raise ValueError('Position passed to step_adversary is outside the grid.')
            a = a_host.to(self.device)
        a = a.reshape(-1).to(torch.int64).contiguous()
        assert a.numel() == N
        image = torch.empty(N, 3, W, W, dtype=torch.float32, device=self.device)
        ts = torch.empty(N, 1, dtype=torch.float32, device=self.device)
        check(self.L.mgplr_step_adversary(self.h, ptr(a), ptr(image), ptr(ts), ptr(self._done_adv), self._stream()),
              'mgplr_step_adversary')
        done = self._done_adv.cpu().numpy().astype(bool)
        rew = torch.zeros(N, 1, dtype=torch.float32, device=self.device)
        return self._adv_obs(image, ts), rew, done, LazyInfos([_NO_INFO] * N)

    # ------------------------------------------------------------------ level resets
    def reset_agent(self):
        self._assert_not_closed()
        obs = self._new_obs()
        o = self._out(obs)
        check(self.L.mgplr_reset_agent(self.h, C.byref(o), self._stream()), 'mgplr_reset_agent')
        self._raise_errors()
        return self._add_full_obs(obs)

    def reset_random(self):
        self._assert_not_closed()
        obs = self._new_obs()
        o = self._out(obs)
        nw = None
        if self.resample_n_clutter:  # _resample_n_clutter: np.random.randint(0, n_clutter) per env (adversarial.py:151-156)
            nw = self._draw_n_walls()
        check(self.L.mgplr_reset_random(self.h, ptr(nw), C.byref(o), self._stream()), 'mgplr_reset_random')
        self._raise_errors()
        return self._add_full_obs(obs)

    def _draw_n_walls(self):
        """One _resample_n_clutter draw per env: np.random.randint(0, n_clutter) (adversarial.py:151-156)."""
        return torch.from_numpy(self.host_rng.randint(0, self.n_clutter, size=self.num_envs).astype(np.int32)).to(self.device)

    def _levels_to_device(self, levels):
        """list of levels -> ('bytes', u8 [n,W,W,3]) or ('str', i32 [n,len])."""
        if isinstance(levels[0], str):
            locs = [[int(a) for a in l.split()] for l in levels]
            ln = len(locs[0])
            assert all(len(l) == ln for l in locs), 'action-string levels must have equal length'
            return 'str', torch.tensor(locs, dtype=torch.int32).reshape(len(levels), ln).to(self.device), ln
        arr = np.stack([np.asarray(l, dtype=np.uint8).reshape(self.W, self.W, 3) for l in levels])
        return 'bytes', torch.from_numpy(np.ascontiguousarray(arr)).to(self.device), 0

    def _reset_to_levels(self, levels, index):
        kind, data, ln = self._levels_to_device(levels)
        n = len(levels)
        obs = self._new_obs()
        o = self._out(obs)
        idx = None if index is None else torch.tensor(index, dtype=torch.int32, device=self.device)
        if kind == 'bytes':
            check(self.L.mgplr_reset_to_encoding(self.h, ptr(data), ptr(idx), n, C.byref(o), self._stream()),
                  'mgplr_reset_to_encoding')
        else:
            check(self.L.mgplr_reset_to_actions(self.h, ptr(data), ln, ptr(idx), n, C.byref(o), self._stream()),
                  'mgplr_reset_to_actions')
        self._raise_errors()
        return self._add_full_obs(obs)

    def reset_to_level(self, level, index):
        """venv.reset_to_level(level, index) -> obs of that env, leading dim 1 (parallel_wrappers.py:334-340)."""
        self._assert_not_closed()
        obs = self._reset_to_levels([level], [int(index)])
        return {k: v[index:index + 1].clone() for k, v in obs.items()}

    def reset_to_level_batch(self, level):
        self._assert_not_closed()
        assert len(level) == self.num_envs
        return self._reset_to_levels(list(level), None)

    def mutate_level(self, num_edits, edits=None):
        """venv.mutate_level(num_edits) (parallel_wrappers.py:352-359 -> adversarial.py:317-397).

        The draws the reference's env processes make from their global np.random are made here on the host from
        `self.host_rng`, per env in env order and in the reference's order within an env; `edits` = (locs[N][k], ops[N][k], n_edits[N], choice[N][2]) replays recorded draws.
        Returns the agent observation (not used by the runner)."""
        self._assert_not_closed()
        N = self.num_envs
        num_tiles = (self.W - 2) ** 2
        n_ops = len(self.editor_actions)
        if edits is None and N > 256:
            # In the reference every env subprocess makes these draws from its OWN unseeded global numpy stream, so only the
            # per-env semantics are defined (DESIGN.md deviation 2): large batches draw all envs' numbers in two calls and
            # keep the per-env rule -- distinct locations in set-iteration order, one op per kept location.
            raw = self.host_rng.randint(0, num_tiles, (N, num_edits)).tolist()
            ops_raw = self.host_rng.randint(0, n_ops, (N, num_edits))
            mx = max(1, num_edits)
            locs = np.zeros((N, mx), np.int32)
            n_ed = np.zeros(N, np.int32)
            for i, row in enumerate(raw):
                u = list(set(row))
                n_ed[i] = len(u)
                locs[i, :len(u)] = u
            ops = np.ascontiguousarray(ops_raw[:, :mx], dtype=np.int32)
            choice = None
        else:
            if edits is None:
                locs_l, ops_l = [], []
                for _ in range(N):
                    edit_locs = list(set(self.host_rng.randint(0, num_tiles, num_edits)))
                    action_idx = self.host_rng.randint(0, n_ops, len(edit_locs))
                    locs_l.append(edit_locs)
                    ops_l.append(action_idx)
                choice = None
            else:
                locs_l, ops_l, n_l, choice = edits
                locs_l = [list(l[:n]) for l, n in zip(locs_l, n_l)]
                ops_l = [list(o[:n]) for o, n in zip(ops_l, n_l)]
            mx = max(1, max(len(l) for l in locs_l))
            locs = np.zeros((N, mx), np.int32)
            ops = np.zeros((N, mx), np.int32)
            n_ed = np.zeros(N, np.int32)
            for i in range(N):
                n_ed[i] = len(locs_l[i])
                locs[i, :n_ed[i]] = locs_l[i]
                ops[i, :n_ed[i]] = ops_l[i]
        d_locs, d_ops, d_n = (torch.from_numpy(a).to(self.device) for a in (locs, ops, n_ed))
        need = torch.zeros(N, 2, dtype=torch.uint8, device=self.device)
        nfree = torch.zeros(N, 2, dtype=torch.int32, device=self.device)
        check(self.L.mgplr_mutate_edits(self.h, ptr(d_locs), ptr(d_ops), ptr(d_n), mx, ptr(need), ptr(nfree), self._stream()),
              'mgplr_mutate_edits')
        if choice is None:
            need_h, nfree_h = need.cpu().numpy(), nfree.cpu().numpy()
            choice = np.zeros((N, 2), np.int32)
            for i, k in zip(*np.nonzero(need_h)):  # np.random.choice(free_idx) (adversarial.py:308-315), env order, goal first then agent
                if nfree_h[i, k] <= 0:
                    raise ValueError("'a' cannot be empty unless no samples are taken")
                choice[i, k] = self.host_rng.choice(int(nfree_h[i, k]))
        d_choice = torch.from_numpy(np.ascontiguousarray(choice, dtype=np.int32)).to(self.device)
        obs = self._new_obs()
        o = self._out(obs)
        check(self.L.mgplr_mutate_finalize(self.h, ptr(d_choice), C.byref(o), self._stream()), 'mgplr_mutate_finalize')
        self._raise_errors()
        self.last_mutation = (need, nfree)
        return obs

    # ------------------------------------------------------------------ student phase
    def step_env(self, action, reset_random=False):
        """venv.step_env(action, reset_random) -> (obs, reward f32 [N,1], done np.bool_[N], infos[N])
        (vec_env.py:113-118; folded wrappers: time_limit.py:24-33, vec_monitor.py:60-85, obs_wrappers.py:168-181)."""
        self._assert_not_closed()
        N = self.num_envs
        # one allocation for the step's outputs (image | direction | reward are contiguous slices of it); the truncated-
        # observation buffers and the StepOut with the persistent pointers live across calls
        o_dir = (N * 75 + 63) & ~63        # slice starts kept 256-byte aligned (the kernel's bulk / vector stores)
        o_rew = o_dir + ((N + 63) & ~63)
        flat = torch.empty(o_rew + N, dtype=torch.float32, device=self.device)
        obs = {'image': flat[:N * 75].view(N, 3, 5, 5), 'direction': flat[o_dir:o_dir + N].view(N, 1)}
        rew = flat[o_rew:].view(N, 1)
        if self._step_out is None:
            self._tr = self._new_obs()
            if self.full_obs:
                self._tr['full_obs'] = torch.empty(N, 3, self.W, self.W, dtype=torch.float32, device=self.device)
            self._step_out = self._out(None, flags=self._flags, ep_return=self._ep_r, ep_length=self._ep_l,
                                       trunc_image=self._tr['image'], trunc_direction=self._tr['direction'])
            if self.full_obs:
                self._step_out.trunc_full_obs = ptr(self._tr['full_obs'])
        tr, o = self._tr, self._step_out
        base = flat.data_ptr()
        o.image, o.direction, o.reward = base, base + o_dir * 4, base + o_rew * 4
        a = torch.as_tensor(action)
        # DR auto-reset of a resample_n_clutter env (MultiGrid-GoalLastVariableBlocksAdversarialEnv-v0, the shipped
        # mg_60b_uni_dr config): every env that finishes this step draws its own wall count (adversarial.py:151-156,574);
        # one draw per env per step is made up front and only the finished envs consume theirs
        n_walls = self._draw_n_walls() if (reset_random and self.resample_n_clutter) else None
        if a.device.type == 'cuda' or n_walls is not None:
            a = a.reshape(-1).to(torch.int64).to(self.device).contiguous()
            check(self.L.mgplr_step_env(self.h, ptr(a), int(bool(reset_random)), ptr(n_walls), 0, C.byref(o), self._stream()),
                  'mgplr_step_env')
            flags = self._flags.cpu().numpy()
            ep_r = ep_l = None
            if (flags & F_DONE).any():      # episode statistics cross only on the steps that end an episode
                ep_r = self._ep_r.cpu().numpy()
                ep_l = self._ep_l.cpu().numpy()
        else:
            if N <= 65536:   # narrow to one byte per action on the way into pinned memory (8x less PCIe traffic)
                self._h_action_u8.copy_(a.reshape(-1))
                check(self.L.mgplr_step_env_host_u8(self.h, self._h_action_u8.data_ptr(), int(bool(reset_random)), 0, C.byref(o),
                                                    self._h_flags.data_ptr(), self._h_done.data_ptr(), N, self._h_ndone.data_ptr(),
                                                    self._stream()), 'mgplr_step_env_host_u8')
            else:            # (a host-side narrowing pass over millions of int64 would cost more than the transfer)
                self._h_action.copy_(a.reshape(-1))
                check(self.L.mgplr_step_env_host(self.h, self._h_action.data_ptr(), int(bool(reset_random)), 0, C.byref(o),
                                                 self._h_flags.data_ptr(), self._h_done.data_ptr(), N, self._h_ndone.data_ptr(),
                                                 self._stream()), 'mgplr_step_env_host')
            flags = self._h_flags_np.copy()
            nd = int(self._h_ndone_np[0])
            ep_r = ep_l = None
            if nd:   # the finished episodes arrive as compact records: scatter them by env
                rec = self._h_done_np[:nd * 16].view(self._done_dtype)
                ep_r, ep_l = dict(zip(rec['env'].tolist(), rec['ep_return'])), dict(zip(rec['env'].tolist(), rec['ep_length'] & 0xffffff))   # (top byte: the env's flags)
        done = (flags & F_DONE) != 0
        infos = LazyInfos([_NO_INFO] * N)
        if ep_r is not None or flags.any():
            t_now = round(time.time() - self.tstart, 6)
            for i in np.nonzero(flags & (F_DONE | F_TRUNC_KEY))[0].tolist():
                info = infos[i]
                f = flags[i]
                if f & F_TRUNC_KEY:
                    info['truncated'] = bool(f & F_TRUNC_VAL)
                    info['truncated_obs'] = _ObsRow(tr, i)
                    self._step_out = None   # the infos now own these buffers: take fresh ones next step
                if f & F_DONE:
                    info['episode'] = {'r': ep_r[i], 'l': ep_l[i], 't': t_now}
            if (flags & F_ERROR).any():   # an auto-reset failed on the device (the kernel flags it: no extra launch otherwise)
                self._raise_errors()
        return self._add_full_obs(obs), rew, done, infos

    def step_env_device(self, action, out, reset_random=False, last_step=0, n_walls=None):
        """Device-resident step: `action` i64 [N] CUDA tensor, `out` a StepOut of raw pointers into rollout storage.
        No host synchronisation; the caller reads flags/rewards from its own tensors."""
        check(self.L.mgplr_step_env(self.h, ptr(action), int(bool(reset_random)), ptr(n_walls), int(last_step),
                                    C.byref(out), self._stream()), 'mgplr_step_env')

    def rollout_device(self, actions_u8, out, reset_random=False, last_step=0):
        """T transitions in one launch from a recorded u8 [T, N] action stream (mgplr_rollout)."""
        T = int(actions_u8.shape[0])
        check(self.L.mgplr_rollout_ex(self.h, ptr(actions_u8), T, int(bool(reset_random)), int(last_step), C.byref(out),
                                      self._stream()), 'mgplr_rollout_ex')

    def step(self, action):
        """venv.step(action) on the ADVERSARIAL env (parallel_wrappers.py:20-25,243-266).  The reference's worker answers
        a finished env with `env.reset()`, i.e. the EMPTY adversary grid and the ADVERSARY observation dict (image [W,W,3],
        time_step, random_z), which `_flatten_obs` (parallel_wrappers.py:208-216) then cannot stack with the other envs'
        agent observations: it fails with KeyError('direction').  Until an episode ends, step() is the same transition as
        step_env(); the step that ends one raises the reference's error.  Evaluation uses the non-adversarial test envs
        (`CudaMazeVecEnv.step`, eval_envs.make_eval_venv), whose auto-reset is well defined."""
        obs, rew, done, infos = self.step_env(action, reset_random=False)
        if done.any():
            raise KeyError('direction')
        return obs, rew, done, infos

    # ------------------------------------------------------------------ getters
    def get_encodings(self, index=None):
        """list of np.uint8 [W,W,3] (parallel_wrappers.py:422-423 -> adversarial.py:162-164)."""
        enc = self.get_encodings_device().cpu().numpy()
        if index is None or len(index) == 0:
            return [enc[i] for i in range(self.num_envs)]
        return [enc[i] for i in index]

    def get_encodings_device(self):
        enc = torch.empty(self.num_envs, self.W, self.W, 3, dtype=torch.uint8, device=self.device)
        check(self.L.mgplr_get_encodings(self.h, ptr(enc), self._stream()), 'mgplr_get_encodings')
        return enc

    def _metrics(self):
        m = torch.empty(self.num_envs, 4, dtype=torch.int32, device=self.device)
        check(self.L.mgplr_get_metrics(self.h, ptr(m), self._stream()), 'mgplr_get_metrics')
        return m.cpu().numpy()

    def get_num_blocks(self):
        return [int(v) for v in self._metrics()[:, 0]]

    def get_distance_to_goal(self):
        return [int(v) for v in self._metrics()[:, 1]]

    def get_passable(self):
        return [(-1 if v < 0 else bool(v)) for v in self._metrics()[:, 2]]

    def get_shortest_path_length(self):
        return [int(v) for v in self._metrics()[:, 3]]

    def get_agent_state(self):
        """int32 [N,8]: x, y, dir, step_count, elapsed, adversary_step_count, adversary_max_steps, rng words used."""
        s = torch.empty(self.num_envs, 8, dtype=torch.int32, device=self.device)
        check(self.L.mgplr_get_agent_state(self.h, ptr(s), self._stream()), 'mgplr_get_agent_state')
        return s.cpu().numpy()

    def peek_rng(self, index, count=4):
        w = np.zeros(count, np.uint32)
        check(self.L.mgplr_peek_rng(self.h, int(index), ptr(w), count), 'mgplr_peek_rng')
        return w

    def get_max_episode_steps(self):
        return self.spec['max_episode_steps']

    def max_episode_steps(self):
        """parallel_wrappers.py:200-203 (the TimeLimit's _max_episode_steps of env 0)."""
        return self.get_max_episode_steps()

    # baselines' two-phase VecEnv calls (parallel_wrappers.py:232-311): the work is a kernel launch, so *_async runs the call
    # and *_wait hands the result over
    def step_env_async(self, action):
        self._pending_result = self.step_env(action, reset_random=False)

    def step_env_reset_random_async(self, action):
        self._pending_result = self.step_env(action, reset_random=True)

    def step_adversary_async(self, action):
        self._pending_result = self.step_adversary(action)

    def step_async(self, action):
        self._pending_result = self.step(action)

    def step_wait(self):
        res, self._pending_result = self._pending_result, None
        return res

    def level_seed(self, index):
        """parallel_wrappers.py:272-285: the `level_seed` attribute of env `index`; MultiGrid envs have none."""
        raise AttributeError("MultiGrid environments have no attribute 'level_seed'")

    def get_level(self):
        """parallel_wrappers.py:418-420 (`level` attribute): not defined by the MultiGrid envs either."""
        raise AttributeError("MultiGrid environments have no attribute 'level'")

    def get_complexity_info(self):
        raise NotImplementedError('get_complexity_info is the Box2D envs\' (BipedalWalker / CarRacing), out of scope')

    def render_to_screen(self):
        raise NotImplementedError('render_to_screen opens a matplotlib window in env process 0 (parallel_wrappers.py:195-198, '
                                  '--render); use get_images() for RGB frames')

    def get_observation_space(self):
        return self.observation_space

    def get_adversary_observation_space(self):
        return self.adversary_observation_space

    def get_adversary_action_space(self):
        return self.adversary_action_space

    def remote_attr(self, name, data=None, flatten=False, index=None):
        table = {'encoding': self.get_encodings, 'n_clutter_placed': self.get_num_blocks, 'passable': self.get_passable,
                 'shortest_path_length': self.get_shortest_path_length, 'distance_to_goal': self.get_distance_to_goal,
                 'seed_value': self.get_seed}
        if name not in table:
            raise NotImplementedError(name)
        res = table[name]()
        if index is not None and len(index) > 0:
            res = [res[i] for i in index]
        return res if flatten else [[r] for r in res]

    def get_images(self, index=None):
        """venv.get_images() (parallel_wrappers.py:187-193): list of np.uint8 [W*32, W*32, 3] level screenshots, one per env
        (or per env in `index`) -- MultiGridEnv.render(mode='level') with the agent's view highlighted (multigrid.py:1105-1140).
        train.py:204-232 saves the first `screenshot_batch_size` of them.  Large batches are rendered in chunks so that the
        staging buffer stays below ~1 GB."""
        from . import tiles
        self._assert_not_closed()
        if getattr(self, '_tile_table', None) is None:
            self._tile_table = torch.from_numpy(tiles.tile_table()).to(self.device)
        idx = list(range(self.num_envs)) if index is None else [int(i) for i in index]
        side = self.W * tiles.TILE
        per = side * side * 3
        chunk = max(1, (1 << 30) // per)
        out = []
        for lo in range(0, len(idx), chunk):
            sub = idx[lo:lo + chunk]
            d_idx = torch.tensor(sub, dtype=torch.int32, device=self.device)
            img = torch.empty(len(sub), side, side, 3, dtype=torch.uint8, device=self.device)
            check(self.L.mgplr_render_images(self.h, ptr(self._tile_table), ptr(d_idx), len(sub), ptr(img), self._stream()),
                  'mgplr_render_images')
            out.extend(img.cpu().numpy())
        return out

    def state_bytes(self):
        return int(self.L.mgplr_venv_state_bytes(self.h))

    def close(self):
        if not self.closed and getattr(self, 'h', None):
            self.L.mgplr_venv_destroy(self.h)
            self.h = None
        self.closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaMazeVecEnv(CudaAdversarialVecEnv):
    """Zero-shot evaluation mazes (envs/multigrid/maze.py: MazeEnv and subclasses, registered without a TimeLimit) as a
    vectorised env with the evaluator's API (eval.py:206-329): `reset()` -> agent obs, `step(action)` with the worker's
    auto-`reset()` on done (parallel_wrappers.py:20-25), VecMonitor episode infos, preprocessed float32 observations.
    Same step / render kernels as the adversarial env; the level is loaded once with mgplr_load_levels."""

    def __init__(self, env_name, num_envs, device='cuda:0', full_obs=False, host_rng_seed=None):
        from .mazes import MAZES
        if env_name not in MAZES:
            raise KeyError('No registered env with id: %s' % env_name)
        m = MAZES[env_name]
        spec = dict(n_clutter=0, size=m['size'], choose_goal_last=True, see_through_walls=True, max_steps=m['max_steps'],
                    max_episode_steps=32767, resample_n_clutter=False, editor_actions='walls_none_agent_goal',
                    fixed_environment=False)
        super().__init__(env_name, num_envs, device=device, spec=spec, full_obs=full_obs, host_rng_seed=host_rng_seed)  # eval.py:200-202
        self.maze = m
        W = m['size']
        if 'goal' in m:
            goals = [tuple(m['goal'])] * num_envs
        else:  # corridor mazes: row then col drawn per env at construction (maze.py:153-156,180-183)
            goals = []
            for _ in range(num_envs):
                row = self.host_rng.choice(m['goal_rows'])
                col = self.host_rng.choice(m['goal_cols'])
                goals.append((int(col), int(row)))
        distinct = sorted(set(goals))
        enc = np.zeros((len(distinct), W, W, 3), np.uint8)
        for k, g in enumerate(distinct):
            enc[k, :, :, 0] = 1
            for y in range(W):
                for x in range(W):
                    if (m['rows'][y] >> x) & 1:
                        enc[k, x, y] = (2, 5, 0)
            enc[k, g[0], g[1]] = (8, 1, 0)
            enc[k, m['start'][0], m['start'][1]] = (10, 0, 0)
        self.goal_positions = goals
        self._levels = torch.from_numpy(enc).to(self.device)
        self._level_index = torch.tensor([distinct.index(g) for g in goals], dtype=torch.int32, device=self.device)
        self.set_seed([52] * num_envs)  # MultiGridEnv's default seed (multigrid.py:351)
        self.reset()

    def reset(self):
        """MultiGridEnv.reset (multigrid.py:470-502) on every env: regenerate the fixed grid, agent at the start facing 0."""
        self._assert_not_closed()
        obs = self._new_obs()
        o = self._out(obs)
        check(self.L.mgplr_load_levels(self.h, ptr(self._levels), int(self._levels.shape[0]), ptr(self._level_index), 0,
                                       C.byref(o), self._stream()), 'mgplr_load_levels')
        self._raise_errors()
        return self._add_full_obs(obs)

    def step(self, action):
        """venv.step(action): worker.step auto-reset()s on done (parallel_wrappers.py:20-25); for a fixed maze that is the
        start state again, i.e. the same transition as step_env."""
        return self.step_env(action, reset_random=False)

    def get_max_episode_steps(self):
        return None


def create_parallel_env(args, adversary=True, device='cuda:0'):
    """Drop-in for util.create_parallel_env (util/__init__.py:184-220): returns (venv, ued_venv) with
    ued_venv is venv for MultiGrid, seeded [0..N-1] (or [args.seed]*N for singleton_env)."""
    if not args.env_name.startswith('MultiGrid'):
        raise NotImplementedError('only the MultiGrid adversarial environments are built (SURVEY.md 8)')
    if getattr(args, 'normalize_returns', False):
        # VecNormalize(ret=True) (util/__init__.py:192, vec_normalize.py:37-45) rescales rewards by a running return std;
        # no shipped MultiGrid config sets it and the step kernel writes raw rewards, so refuse instead of ignoring it
        raise NotImplementedError('--normalize_returns (VecNormalize ret=True) is not part of the B200 MultiGrid path')
    singleton = bool(getattr(args, 'singleton_env', False))
    venv = CudaAdversarialVecEnv(args.env_name, args.num_processes, device=device,
                                 fixed_environment=True if singleton else None,
                                 full_obs=bool(getattr(args, 'use_global_critic', False) or getattr(args, 'use_global_policy', False)))
    if singleton:
        seeds = [args.seed] * args.num_processes
    else:
        seeds = [i for i in range(args.num_processes)]
    venv.set_seed(seeds)
    return venv, venv
