"""Golden traces of the Kruskal perfect mazes (envs/multigrid/mst_maze.py), produced by EXECUTING the reference's
MSTMazeEnv classes over the third-party shim with the worker's auto-reset rule (parallel_wrappers.py:20-25: on done,
ob = env.reset(), i.e. a new maze).  TEST INFRASTRUCTURE ONLY.

  python oracle/gen_golden_mst.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def bfs_next(env):
    """next cell on a shortest path agent -> goal (so that goal episodes occur in the fixture)"""
    W = env.width
    free = lambda x, y: env.grid.get(x, y) is None or env.grid.get(x, y).type in ('goal', 'agent')  # noqa: E731
    src = (int(env.agent_pos[0][0]), int(env.agent_pos[0][1]))
    dst = (int(env.goal_pos[0]), int(env.goal_pos[1]))
    prev = {dst: None}
    q = [dst]
    while q:
        c = q.pop(0)
        if c == src:
            break
        for dx, dy in ((1, 0), (0, 1), (-1, 0), (0, -1)):
            n = (c[0] + dx, c[1] + dy)
            if 0 <= n[0] < W and 0 <= n[1] < W and n not in prev and free(*n):
                prev[n] = c
                q.append(n)
    return prev.get(src)


def main():
    import numpy as np
    rh.activate()
    import envs.multigrid.mst_maze  # noqa: F401  registers the ids
    from envs.registration import make as gym_make
    rec = {}
    for tag, env_id, T in (('Small', 'MultiGrid-PerfectMazeSmall-v0', 1400), ('Medium', 'MultiGrid-PerfectMazeMedium-v0', 2600)):
        env = gym_make(env_id)
        encs = [env.grid.encode()]            # the maze built by the constructor (seed 52)
        env.seed(7)
        o = env.reset()
        encs.append(env.grid.encode())
        rs = np.random.RandomState(13)
        acts = rs.randint(0, 7, size=T)
        obs, dirs, rews, dones = [np.array(o['image'], np.uint8)], [int(o['direction'][0])], [], []
        for t in range(T):
            a = int(acts[t])
            follow = 0.85 if (t // 300) % 2 == 0 else 0.05  # alternate goal-seeking and wandering (step-budget) phases
            if rs.rand() < follow:
                nxt = bfs_next(env)
                if nxt is not None:
                    ax, ay = int(env.agent_pos[0][0]), int(env.agent_pos[0][1])
                    want = {(1, 0): 0, (0, 1): 1, (-1, 0): 2, (0, -1): 3}[(nxt[0] - ax, nxt[1] - ay)]
                    d = int(env.agent_dir[0])
                    a = 2 if want == d else (1 if (want - d) % 4 == 1 else 0)
            acts[t] = a
            o, r, d, info = env.step(a)
            if d:
                o = env.reset()
                encs.append(env.grid.encode())
            obs.append(np.array(o['image'], np.uint8)); dirs.append(int(o['direction'][0])); rews.append(np.float32(r)); dones.append(bool(d))
        rec.update({tag + '_actions': acts.astype(np.uint8), tag + '_obs': np.stack(obs), tag + '_dirs': np.array(dirs, np.int8),
                    tag + '_rewards': np.array(rews, np.float32), tag + '_dones': np.array(dones), tag + '_encodings': np.stack(encs),
                    tag + '_max_steps': env.max_steps})
        print(env_id, 'episodes', int(np.sum(dones)), 'goals', int((np.array(rews) > 0).sum()), 'mazes', len(encs))
    np.savez_compressed(os.path.join(GOLDEN, 'eval_mst_mazes.npz'), **rec)


if __name__ == '__main__':
    main()
