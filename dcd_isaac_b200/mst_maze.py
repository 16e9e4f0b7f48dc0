"""Kruskal perfect mazes (envs/multigrid/mst_maze.py:17-163: MultiGrid-PerfectMaze{Small,Medium}-v0) as a vectorised
evaluation env.  The maze GENERATOR runs on the host -- it is a union-find over a shuffled edge list, a few hundred Python
operations per reset of an evaluation env (the evaluator runs 2 processes per env, arguments.py:433-436) -- with each
env's own `numpy.random.RandomState`, which is the very object the reference draws from (gym's `np_random`), so the
`choice` / `shuffle` / `randint` streams are the reference's by construction; stepping and rendering are the same CUDA
kernels as every other MultiGrid env, the regenerated levels are uploaded with mgplr_load_levels_at.

PerfectMazeLarge / XL (51 and 101 cells wide) exceed the 32-column bit-plane of the main path: CudaWideMSTMazeVecEnv runs
them on the wide evaluation kernels (csrc/mgplr_wide.cu: rows of 128 columns, the same view code), same host generator.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import check, ptr
from .vec_env import CudaAdversarialVecEnv, F_DONE, F_GOAL, seed_limbs

MST_MAZES = {'MultiGrid-PerfectMazeSmall-v0': 11, 'MultiGrid-PerfectMazeMedium-v0': 21,
             'MultiGrid-PerfectMazeLarge-v0': 51, 'MultiGrid-PerfectMazeXL-v0': 101}


class _UnionFind(object):
    """weighted quick union with path compression (util/unionfind.py); only `connected` / `union` are needed, and the
    maze only depends on the partition, not on which root represents it."""

    def __init__(self):
        self.parent, self.size = {}, {}

    def add(self, x):
        self.parent[x] = x
        self.size[x] = 1

    def find(self, x):
        root = x
        while self.parent[root] != root:
            root = self.parent[root]
        while self.parent[x] != root:
            self.parent[x], x = root, self.parent[x]
        return root

    def connected(self, a, b):
        return self.find(a) == self.find(b)

    def union(self, a, b):
        ra, rb = self.find(a), self.find(b)
        if ra == rb:
            return
        if self.size[ra] < self.size[rb]:
            ra, rb = rb, ra
        self.parent[rb] = ra
        self.size[ra] += self.size[rb]


def grid_edges(h, w):
    """`list(networkx.grid_graph([h, w]).edges)` -- the list the reference shuffles (mst_maze.py:64,76-77)."""
    import networkx
    g = networkx.grid_graph([h, w])
    return list(g.nodes), list(g.edges)


class MSTMazeHost(object):
    """One env's generator state: its RandomState and the current level (for the goal-respawn draws)."""

    def __init__(self, size):
        self.size = size
        self.rs = np.random.RandomState()
        h = (size - 2) // 2 + 1
        self.nodes, self.edges = grid_edges(h, h)
        self.seed(52)            # MultiGridEnv.__init__(seed=52) re-seeds whatever MSTMazeEnv.__init__ was given (multigrid.py:351,459)
        self.first = self.gen()  # ... and calls reset() (multigrid.py:463)

    def seed(self, seed):
        lo, hi, n = seed_limbs(seed)
        self.rs.seed([lo, hi][:n])

    def gen(self):
        """MSTMazeEnv._gen_grid (mst_maze.py:97-115): returns the level's encoding u8 [W][W][3] indexed [x][y]."""
        size = self.size
        corners = [(1, 1), (size - 2, 1), (1, size - 2), (size - 2, size - 2)]
        a_idx, g_idx = self.rs.choice(range(4), size=(2,), replace=False)          # _sample_start_and_goal_pos (:40-53)
        self.start, self.goal = corners[a_idx], corners[g_idx]
        bit = np.ones((size - 2, size - 2), np.uint8)
        ds = _UnionFind()
        for v in self.nodes:
            bit[v[0] * 2][v[1] * 2] = 0
            ds.add(v)
        edges = list(self.edges)
        self.rs.shuffle(edges)                                                         # (:76-77)
        for u, v in edges:
            if not ds.connected(u, v):
                y1, x1, y2, x2 = u[0] * 2, u[1] * 2, v[0] * 2, v[1] * 2
                bit[y1 + (y2 - y1) // 2][x1 + (x2 - x1) // 2] = 0
                ds.union(u, v)
        enc = np.zeros((size, size, 3), np.uint8)
        enc[:, :, 0] = 1
        wall = np.zeros((size, size), bool)   # [x][y]
        wall[0, :] = wall[-1, :] = wall[:, 0] = wall[:, -1] = True
        wall[1:-1, 1:-1] = bit.T != 0         # bit_map[y, x] -> Wall at (x+1, y+1)  (:110-115)
        enc[wall] = (2, 5, 0)
        enc[self.goal[0], self.goal[1]] = (8, 1, 0)
        enc[self.start[0], self.start[1]] = (10, 0, 0)
        self.wall = wall
        return enc

    def respawn(self):
        """agent_is_done's place_one_agent (multigrid.py:821-838,565-632): rejection sampling over the whole grid with the
        agent off the grid; only the draws matter (the worker resets the env right after)."""
        W = self.size
        while True:
            x, y = self.rs.randint(0, W), self.rs.randint(0, W)
            if not self.wall[x, y] and (x, y) != tuple(self.goal):
                return


class CudaMSTMazeVecEnv(CudaAdversarialVecEnv):
    """venv.reset() / venv.step(action) of the Kruskal perfect mazes with the evaluator's API (eval.py:206-329)."""

    def __init__(self, env_name, num_envs, device='cuda:0', full_obs=False):
        if env_name not in MST_MAZES:
            raise KeyError('No registered env with id: %s' % env_name)
        size = MST_MAZES[env_name]
        if size > 32:
            raise ValueError('%s is %d cells wide: use CudaWideMSTMazeVecEnv (eval_envs.make_eval_venv picks it)' % (env_name, size))
        spec = dict(n_clutter=0, size=size, choose_goal_last=True, see_through_walls=True, max_steps=2 * size * size,
                    max_episode_steps=32767, resample_n_clutter=False, editor_actions='walls_none_agent_goal',
                    fixed_environment=False)
        super().__init__(env_name, num_envs, device=device, spec=spec, full_obs=full_obs)
        self.start_dir = 0     # place_agent_at_pos(rand_dir=True) forces direction 0 (multigrid.py:668-672)
        self.hosts = [MSTMazeHost(size) for _ in range(num_envs)]
        self._upload(list(range(num_envs)), [h.first for h in self.hosts])  # the maze built by the constructor's reset()

    # -- seeding goes to the host generators (the device RNG is not used by these envs)
    def set_seed(self, seeds):
        seeds = list(seeds)
        assert len(seeds) == self.num_envs
        for h, s in zip(self.hosts, seeds):
            h.seed(s)
        self.seed_values = seeds
        return [[s] for s in seeds]

    def seed(self, seed, index):
        self.hosts[index].seed(seed)
        self.seed_values[index] = seed
        return [seed]

    def _upload(self, envs, encs=None, obs=None):
        """Regenerate (unless given) and upload the levels of `envs`; observations of those envs go to `obs`."""
        if encs is None:
            encs = [self.hosts[i].gen() for i in envs]
        enc = torch.from_numpy(np.ascontiguousarray(np.stack(encs))).to(self.device)
        idx = torch.tensor(envs, dtype=torch.int32, device=self.device)
        o = self._out(obs) if obs is not None else self._out()
        check(self.L.mgplr_load_levels_at(self.h, ptr(enc), ptr(idx), len(envs), getattr(self, 'start_dir', 0), C.byref(o),
                                          self._stream()), 'mgplr_load_levels_at')
        self._raise_errors()

    def reset(self):
        """MultiGridEnv.reset on every env: a NEW maze each time (mst_maze.py:97-99)."""
        self._assert_not_closed()
        obs = self._new_obs()
        self._upload(list(range(self.num_envs)), obs=obs)
        return self._add_full_obs(obs)

    def step(self, action):
        """venv.step(action): on done the worker calls env.reset() (parallel_wrappers.py:20-25), i.e. a new maze; a goal
        first makes the respawn draws of agent_is_done on the old one."""
        self.full_obs, keep = False, self.full_obs
        try:
            obs, rew, done, infos = self.step_env(action, reset_random=False)
        finally:
            self.full_obs = keep
        flags = self._h_flags.numpy() if torch.as_tensor(action).device.type != 'cuda' else self._flags.cpu().numpy()
        fin = np.nonzero(flags & F_DONE)[0]
        if len(fin):
            for i in fin:
                if flags[i] & F_GOAL:
                    self.hosts[i].respawn()
            self._upload([int(i) for i in fin], obs=obs)
        return self._add_full_obs(obs), rew, done, infos

    def get_max_episode_steps(self):
        return None


class CudaWideMSTMazeVecEnv(object):
    """PerfectMazeLarge / PerfectMazeXL (51 / 101 cells wide, mst_maze.py:128-136) with the evaluator's API (eval.py:206-329:
    `reset()`, `step(action)` with the worker's auto-reset, VecMonitor episode infos, preprocessed float32 observations) on the
    wide evaluation kernels (mgplr_wide_*).  Same host generator as the narrow mazes."""

    def __init__(self, env_name, num_envs, device='cuda:0', full_obs=False):
        import time
        from . import _lib
        from .spaces import Box, Discrete
        if env_name not in MST_MAZES:
            raise KeyError('No registered env with id: %s' % env_name)
        if full_obs:
            raise NotImplementedError('full_obs (use_global_policy) is not built for the wide mazes')
        self.env_name, self.num_envs, self.device = env_name, int(num_envs), torch.device(device)
        if self.device.type != 'cuda' or not torch.cuda.is_available():
            raise _lib.MgplrError('CudaWideMSTMazeVecEnv needs a CUDA device (no CPU fallback)')
        self.W = MST_MAZES[env_name]
        self.max_steps = 2 * self.W * self.W
        self.L = _lib.load()
        h = C.c_void_p()
        torch.cuda.set_device(self.device)
        check(self.L.mgplr_wide_create(self.W, self.max_steps, self.num_envs, self.device.index or 0, C.byref(h)), 'mgplr_wide_create')
        self.h, self.closed, self.tstart = h, False, time.time()
        self.observation_space = {'image': Box(0, 255, (3, 5, 5), 'uint8'), 'direction': Box(0, 3, (1,), 'uint8')}
        self.action_space = Discrete(7)
        N = self.num_envs
        self._flags = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self._ep_r = torch.zeros(N, dtype=torch.float32, device=self.device)
        self._ep_l = torch.zeros(N, dtype=torch.int32, device=self.device)
        self.seed_values = [None] * N
        self.hosts = [MSTMazeHost(self.W) for _ in range(N)]
        self._upload(list(range(N)), [h_.first for h_ in self.hosts])

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _new_obs(self):
        N = self.num_envs
        return {'image': torch.empty(N, 3, 5, 5, dtype=torch.float32, device=self.device),
                'direction': torch.empty(N, 1, dtype=torch.float32, device=self.device)}

    def _out(self, obs, **kw):
        from ._lib import StepOut
        o = StepOut()
        if obs is not None:
            o.image, o.direction = ptr(obs['image']), ptr(obs['direction'])
        for k, v in kw.items():
            setattr(o, k, ptr(v))
        return o

    def set_seed(self, seeds):
        seeds = list(seeds)
        assert len(seeds) == self.num_envs
        for h, s in zip(self.hosts, seeds):
            h.seed(s)
        self.seed_values = seeds
        return [[s] for s in seeds]

    def seed(self, seed, index):
        self.hosts[index].seed(seed)
        self.seed_values[index] = seed
        return [seed]

    def get_seed(self):
        return list(self.seed_values)

    def _upload(self, envs, encs=None, obs=None):
        if encs is None:
            encs = [self.hosts[i].gen() for i in envs]
        enc = torch.from_numpy(np.ascontiguousarray(np.stack(encs))).to(self.device)
        idx = torch.tensor(envs, dtype=torch.int32, device=self.device)
        o = self._out(obs)
        check(self.L.mgplr_wide_load_levels(self.h, ptr(enc), ptr(idx), len(envs), 0, C.byref(o), self._stream()),
              'mgplr_wide_load_levels')

    def reset(self):
        """MultiGridEnv.reset on every env: a NEW maze each time (mst_maze.py:97-99)."""
        obs = self._new_obs()
        self._upload(list(range(self.num_envs)), obs=obs)
        return obs

    def step(self, action):
        """venv.step(action): transition, VecMonitor episode info; a finished env gets its next maze (the worker's auto-reset,
        parallel_wrappers.py:20-25) after the respawn draws agent_is_done makes on the old one when the goal was reached."""
        import time
        from .vec_env import LazyInfos, _NO_INFO
        N = self.num_envs
        a = torch.as_tensor(action).reshape(-1).to(torch.int64).to(self.device).contiguous()
        obs = self._new_obs()
        rew = torch.empty(N, 1, dtype=torch.float32, device=self.device)
        o = self._out(obs, reward=rew, flags=self._flags, ep_return=self._ep_r, ep_length=self._ep_l)
        check(self.L.mgplr_wide_step(self.h, ptr(a), C.byref(o), self._stream()), 'mgplr_wide_step')
        flags = self._flags.cpu().numpy()
        done = (flags & F_DONE) != 0
        infos = LazyInfos([_NO_INFO] * N)
        fin = np.nonzero(done)[0].tolist()
        if fin:
            ep_r, ep_l = self._ep_r.cpu().numpy(), self._ep_l.cpu().numpy()
            t_now = round(time.time() - self.tstart, 6)
            for i in fin:
                infos[i]['episode'] = {'r': ep_r[i], 'l': ep_l[i], 't': t_now}
                if flags[i] & F_GOAL:
                    self.hosts[i].respawn()
            self._upload(fin, obs=obs)
        return obs, rew, done, infos

    def get_encodings(self):
        enc = torch.empty(self.num_envs, self.W, self.W, 3, dtype=torch.uint8, device=self.device)
        check(self.L.mgplr_wide_get_encodings(self.h, ptr(enc), self._stream()), 'mgplr_wide_get_encodings')
        return [e for e in enc.cpu().numpy()]

    def get_max_episode_steps(self):
        return None

    def close(self):
        if not self.closed and getattr(self, 'h', None):
            self.L.mgplr_wide_destroy(self.h)
            self.h = None
        self.closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
