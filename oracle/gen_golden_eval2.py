"""Golden traces of the remaining zero-shot evaluation envs, produced by EXECUTING the reference's classes over the
third-party shim with the worker's auto-reset rule (parallel_wrappers.py:20-25: on done, ob = env.reset()):
PerfectMazeLarge / PerfectMazeXL (envs/multigrid/mst_maze.py:128-136), MiniGrid-SimpleCrossingS9N1/S9N3/S11N5-v0
(envs/multigrid/crossing.py) and MiniGrid-FourRooms-v0 (envs/multigrid/fourrooms.py).  TEST INFRASTRUCTURE ONLY.

  python oracle/gen_golden_eval2.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def _agent(env):
    p, d = env.agent_pos, env.agent_dir
    if hasattr(p[0], '__len__'):   # MultiGridEnv keeps per-agent arrays
        return int(p[0][0]), int(p[0][1]), int(d[0])
    return int(p[0]), int(p[1]), int(d)


def _goal(env):
    if hasattr(env, 'goal_pos'):
        return int(env.goal_pos[0]), int(env.goal_pos[1])
    for x in range(env.width):
        for y in range(env.height):
            c = env.grid.get(x, y)
            if c is not None and c.type == 'goal':
                return x, y


def bfs_next(env):
    """next cell on a shortest path agent -> goal (so that goal episodes occur in the fixture)"""
    W = env.width
    free = lambda x, y: env.grid.get(x, y) is None or env.grid.get(x, y).type in ('goal', 'agent')  # noqa: E731
    ax, ay, _ = _agent(env)
    dst = _goal(env)
    prev = {dst: None}
    q = [dst]
    while q:
        c = q.pop(0)
        if c == (ax, ay):
            break
        for dx, dy in ((1, 0), (0, 1), (-1, 0), (0, -1)):
            n = (c[0] + dx, c[1] + dy)
            if 0 <= n[0] < W and 0 <= n[1] < W and n not in prev and free(*n):
                prev[n] = c
                q.append(n)
    return prev.get((ax, ay))


def _encode(env):
    """grid.encode() with the agent drawn at its cell for the MiniGrid envs (their agent is not a grid object)."""
    import numpy as np
    enc = np.array(env.grid.encode(), np.uint8)
    ax, ay, d = _agent(env)
    enc[ax, ay] = (10, 0, d)
    return enc


def main():
    import numpy as np
    rh.activate()
    import envs.multigrid.crossing  # noqa: F401  registers the ids
    import envs.multigrid.fourrooms  # noqa: F401
    import envs.multigrid.mst_maze  # noqa: F401
    from envs.registration import make as gym_make
    rec = {}
    cases = (('Large', 'MultiGrid-PerfectMazeLarge-v0', 3200, 400), ('XL', 'MultiGrid-PerfectMazeXL-v0', 2200, 700),
             ('CrossS9N1', 'MiniGrid-SimpleCrossingS9N1-v0', 900, 120), ('CrossS9N3', 'MiniGrid-SimpleCrossingS9N3-v0', 900, 120),
             ('CrossS11N5', 'MiniGrid-SimpleCrossingS11N5-v0', 1100, 150), ('FourRooms', 'MiniGrid-FourRooms-v0', 1500, 120))
    for tag, env_id, T, phase in cases:
        env = gym_make(env_id)
        encs = [_encode(env)]            # the level built by the constructor
        env.seed(7)
        o = env.reset()
        encs.append(_encode(env))
        rs = np.random.RandomState(13)
        acts = rs.randint(0, 7, size=T)
        obs, dirs, rews, dones = [np.array(o['image'], np.uint8)], [int(o['direction'][0])], [], []
        for t in range(T):
            a = int(acts[t])
            follow = 0.9 if (t // phase) % 3 != 2 else 0.05  # goal-seeking phases, then a wandering (step-budget) phase
            if rs.rand() < follow:
                nxt = bfs_next(env)
                if nxt is not None:
                    ax, ay, d = _agent(env)
                    want = {(1, 0): 0, (0, 1): 1, (-1, 0): 2, (0, -1): 3}[(nxt[0] - ax, nxt[1] - ay)]
                    a = 2 if want == d else (1 if (want - d) % 4 == 1 else 0)
            acts[t] = a
            o, r, d, info = env.step(a)
            if d:
                o = env.reset()
                encs.append(_encode(env))
            obs.append(np.array(o['image'], np.uint8)); dirs.append(int(o['direction'][0])); rews.append(np.float32(r)); dones.append(bool(d))
        rec.update({tag + '_actions': acts.astype(np.uint8), tag + '_obs': np.stack(obs), tag + '_dirs': np.array(dirs, np.int8),
                    tag + '_rewards': np.array(rews, np.float32), tag + '_dones': np.array(dones), tag + '_encodings': np.stack(encs),
                    tag + '_max_steps': env.max_steps})
        print(env_id, 'episodes', int(np.sum(dones)), 'goals', int((np.array(rews) > 0).sum()), 'levels', len(encs))
    np.savez_compressed(os.path.join(GOLDEN, 'eval_envs2.npz'), **rec)


if __name__ == '__main__':
    main()
