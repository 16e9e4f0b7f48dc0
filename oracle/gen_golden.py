"""Generate tests/golden/*.npz by EXECUTING the reference's own Python modules (TEST INFRASTRUCTURE ONLY).

Run here (needs /root/reference):   python oracle/gen_golden.py [--only env|venv|fullobs|adv|mutate|images|plr|runner]

The reference ships no tests or golden vectors, so these fixtures are the
pinning: every array below is an output of the unmodified reference files
(envs/multigrid/{multigrid,adversarial}.py, envs/wrappers/*.py, util/__init__.py,
level_replay/*.py, algos/storage.py) running over oracle/shim (stand-ins for the un-installed
third-party gym / gym-minigrid / baselines only).  The fixtures are small and committed; the GPU
box never sees /root/reference.
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

# (tag, env id or None, kwargs for a direct AdversarialEnv construction, TimeLimit steps)
ENV_CASES = [
    ('gl15', 'MultiGrid-GoalLastAdversarial-v0', None, 250),
    ('gl15_opaque', 'MultiGrid-GoalLastOpaqueWallsAdversarial-v0', None, 250),
    ('fb15_opaque', 'MultiGrid-GoalLastFewerBlocksOpaqueWallsAdversarial-v0', None, 250),
    ('gl15_tl50', 'MultiGrid-GoalLastAdversarialEnv30-v0', None, 50),
    ('sz25', None, dict(size=25, n_clutter=50, choose_goal_last=True, see_through_walls=True, max_steps=250), 250),
    ('sz25_opaque', None, dict(size=25, n_clutter=50, choose_goal_last=True, see_through_walls=False, max_steps=250), 250),
    ('mini6', 'MultiGrid-MiniGoalLastAdversarial-v0', None, 50),
    # singleton_env: fixed_environment re-seeds on every reset_random (adversarial.py:542-543, util/__init__.py:137-139)
    ('gl15_fixed', 'MultiGrid-GoalLastFewerBlocksOpaqueWallsAdversarial-v0', {'fixed_environment': True}, 250),
]


def _make(env_id, kwargs, tl, seed):
    import importlib
    adv = importlib.import_module("envs.multigrid.adversarial")
    from envs.wrappers import TimeLimit
    if env_id is not None:
        return rh.make_env(env_id, seed=seed, **(kwargs or {}))
    return TimeLimit(adv.AdversarialEnv(seed=seed, **kwargs), max_episode_steps=tl)


def _biased_actions(rs, n):
    """Forward-biased action stream (uniform actions rarely reach goals)."""
    import numpy as np
    a = rs.randint(0, 7, size=n)
    f = rs.rand(n) < 0.45
    a[f] = 2
    return a.astype(np.uint8)


def gen_env_traces():
    """In-process single-env traces through the reference TimeLimit + the worker's step_env rule
    (envs/wrappers/parallel_wrappers.py:27-37, restated in the 6 lines below because it is a closure)."""
    import numpy as np
    for tag, env_id, kwargs, tl in ENV_CASES:
        for mode in ('agent', 'random'):
            out = {}
            n_env, T = (6, 700) if tag != 'mini6' else (6, 300)
            rs = np.random.RandomState(1234)
            encs, starts, obs_l, dir_l, rew_l, flag_l, tobs_l, tdir_l, pos_l, tail_l, met_l = ([] for _ in range(11))
            for i in range(n_env):
                env = _make(env_id, kwargs, tl, seed=i)
                env.seed(i)
                o = env.reset_random()
                encs.append(env.encoding.copy())
                met_l.append([env.n_clutter_placed, env.distance_to_goal, int(env.passable), env.shortest_path_length])
                starts.append(np.array(o['image'], dtype=np.uint8))
                acts = _biased_actions(rs, T)
                ob_i, di_i, re_i, fl_i, to_i, td_i, po_i = [], [], [], [], [], [], []
                for t in range(T):
                    o, r, d, info = env.step(int(acts[t]))
                    flags = 0
                    tob = np.zeros((5, 5, 3), np.uint8)
                    tdir = 0
                    if r != 0:
                        flags |= 8
                    if 'truncated' in info:
                        flags |= 2 | (4 if info['truncated'] else 0)
                        tob = np.array(info['truncated_obs']['image'], dtype=np.uint8)
                        tdir = int(info['truncated_obs']['direction'][0])
                    if d:
                        flags |= 1
                        if mode == 'random':
                            env.reset_random()
                            o = env.reset_agent()
                        else:
                            o = env.reset_agent()
                    ob_i.append(np.array(o['image'], dtype=np.uint8))
                    di_i.append(int(o['direction'][0]))
                    re_i.append(np.float32(r))
                    fl_i.append(flags)
                    to_i.append(tob)
                    td_i.append(tdir)
                    po_i.append([int(env.agent_pos[0][0]), int(env.agent_pos[0][1]), int(env.agent_dir[0]), env.step_count])
                obs_l.append(ob_i); dir_l.append(di_i); rew_l.append(re_i); flag_l.append(fl_i)
                tobs_l.append(to_i); tdir_l.append(td_i); pos_l.append(po_i)
                tail_l.append(env.np_random.randint(0, 2 ** 32, 4, dtype=np.uint64).astype(np.uint32))
                out.setdefault('actions', []).append(acts)
            W = encs[0].shape[0]
            np.savez_compressed(
                os.path.join(GOLDEN, 'env_trace_%s_%s.npz' % (tag, mode)),
                W=W, time_limit=tl, max_steps=env.max_steps, see_through=int(env.see_through_walls),
                n_clutter=env.n_clutter, seeds=np.arange(n_env), fixed=int(bool(env.fixed_environment)),
                encodings=np.stack(encs), metrics=np.array(met_l), first_obs=np.stack(starts),
                actions=np.stack(out['actions']), obs=np.array(obs_l, dtype=np.uint8), dirs=np.array(dir_l, dtype=np.int8),
                rewards=np.array(rew_l, dtype=np.float32), flags=np.array(flag_l, dtype=np.uint8),
                trunc_obs=np.array(tobs_l, dtype=np.uint8), trunc_dir=np.array(tdir_l, dtype=np.int8),
                pos=np.array(pos_l, dtype=np.int16), rng_tail=np.stack(tail_l))
            print('env_trace', tag, mode, 'episodes', int((np.array(flag_l) & 1).sum()),
                  'goals', int(((np.array(flag_l) & 8) > 0).sum()))


def gen_venv():
    """The REAL vectorised path: util.create_parallel_env -> spawn subprocess workers -> VecMonitor ->
    VecNormalize -> VecPreprocessImageWrapper (util/__init__.py:184-220), driven through step_env."""
    import numpy as np
    import torch
    from types import SimpleNamespace
    import util
    for tag, env_name in (('gl15', 'MultiGrid-GoalLastAdversarial-v0'),
                          ('fb15_opaque', 'MultiGrid-GoalLastFewerBlocksOpaqueWallsAdversarial-v0'),
                          ('gl15_tl50', 'MultiGrid-GoalLastAdversarialEnv30-v0')):
        N, T = 4, 320
        args = SimpleNamespace(env_name=env_name, seed=1, singleton_env=False, use_global_critic=False,
                               use_global_policy=False, num_processes=N, normalize_returns=False)
        venv, _ = util.create_parallel_env(args)
        for mode in (False, True):
            venv.set_seed(list(range(N)))
            venv.reset_random()
            enc = np.stack(venv.get_encodings())
            o = venv.reset_agent()
            first = {k: v.numpy().copy() for k, v in o.items()}
            rs = np.random.RandomState(99)
            acts = np.stack([_biased_actions(rs, N) for _ in range(T)])  # [T, N]
            img, dr, rew, dones, tkey, tval, er, el, timg, tdir = ([] for _ in range(10))
            for t in range(T):
                o, r, d, infos = venv.step_env(torch.from_numpy(acts[t].astype(np.int64)).view(N, 1), reset_random=mode)
                img.append(o['image'].numpy().copy()); dr.append(o['direction'].numpy().copy())
                rew.append(r.numpy().copy()); dones.append(np.array(d, dtype=bool))
                tkey.append([('truncated' in i) for i in infos])
                tval.append([bool(i.get('truncated', False)) for i in infos])
                er.append([np.float32(i['episode']['r']) if 'episode' in i else np.float32(0) for i in infos])
                el.append([int(i['episode']['l']) if 'episode' in i else 0 for i in infos])
                timg.append([i['truncated_obs']['image'].numpy() if 'truncated_obs' in i else np.zeros((3, 5, 5), np.float32) for i in infos])
                tdir.append([float(i['truncated_obs']['direction'][0]) if 'truncated_obs' in i else 0.0 for i in infos])
            np.savez_compressed(
                os.path.join(GOLDEN, 'venv_%s_%s.npz' % (tag, 'random' if mode else 'agent')),
                env_name=env_name, encodings=enc, first_image=first['image'], first_direction=first['direction'],
                actions=acts, image=np.stack(img), direction=np.stack(dr), reward=np.stack(rew), done=np.stack(dones),
                trunc_key=np.array(tkey), trunc_val=np.array(tval), ep_r=np.array(er, dtype=np.float32),
                ep_l=np.array(el, dtype=np.int32), trunc_image=np.array(timg, dtype=np.float32),
                trunc_direction=np.array(tdir, dtype=np.float32),
                num_blocks=np.array(venv.get_num_blocks()), passable=np.array(venv.get_passable(), dtype=np.int8),
                spl=np.array(venv.get_shortest_path_length()), max_episode_steps=venv.get_max_episode_steps(),
                adv_steps=venv.adversary_observation_space['time_step'].high[0])
            print('venv', tag, mode, 'episodes', int(np.stack(dones).sum()))
        venv.close()


def gen_fullobs():
    """MultiGridFullyObsWrapper (use_global_critic / use_global_policy, util/__init__.py:175-178): the real vectorised
    path with the extra 'full_obs' observation, through reset_random / reset_agent / step_env / reset_to_level."""
    import numpy as np
    import torch
    from types import SimpleNamespace
    import util
    for tag, env_name in (('gl15', 'MultiGrid-GoalLastAdversarial-v0'), ('tl50', 'MultiGrid-GoalLastAdversarialEnv30-v0')):
        N, T = 4, 140
        args = SimpleNamespace(env_name=env_name, seed=1, singleton_env=False, use_global_critic=True,
                               use_global_policy=False, num_processes=N, normalize_returns=False)
        venv, _ = util.create_parallel_env(args)
        out = dict(env_name=env_name, space_shape=np.array(venv.observation_space['full_obs'].shape))
        for mode in (False, True):
            m = 'random' if mode else 'agent'
            venv.set_seed(list(range(N)))
            o = venv.reset_random()
            out['rr_full_' + m] = o['full_obs'].numpy().copy()
            o = venv.reset_agent()
            out['first_full_' + m] = o['full_obs'].numpy().copy()
            out['first_image_' + m] = o['image'].numpy().copy()
            rs = np.random.RandomState(7)
            acts = np.stack([_biased_actions(rs, N) for _ in range(T)])
            full, img, dones, tfull, tkey = [], [], [], [], []
            for t in range(T):
                o, r, d, infos = venv.step_env(torch.from_numpy(acts[t].astype(np.int64)).view(N, 1), reset_random=mode)
                full.append(o['full_obs'].numpy().copy()); img.append(o['image'].numpy().copy()); dones.append(np.array(d, dtype=bool))
                tkey.append([('truncated_obs' in i) for i in infos])
                tfull.append([i['truncated_obs']['full_obs'].numpy() if ('truncated_obs' in i and 'full_obs' in i['truncated_obs'])
                              else np.zeros(out['space_shape'], np.float32) for i in infos])
            out['actions_' + m] = acts
            out['full_' + m] = np.stack(full); out['image_' + m] = np.stack(img); out['done_' + m] = np.stack(dones)
            out['trunc_key_' + m] = np.array(tkey); out['trunc_full_' + m] = np.array(tfull, dtype=np.float32)
        enc = venv.get_encodings()
        o = venv.reset_to_level_batch(enc)
        out['level_full'] = o['full_obs'].numpy().copy()
        out['level_enc'] = np.stack(venv.get_encodings())
        np.savez_compressed(os.path.join(GOLDEN, 'fullobs_%s.npz' % tag), **out)
        print('fullobs', tag, out['space_shape'], 'episodes', int(out['done_agent'].sum()), int(out['done_random'].sum()),
              'trunc', int(out['trunc_key_agent'].sum()))
        venv.close()


ADV_CASES = [
    ('gl50', 'MultiGrid-GoalLastAdversarial-v0'),
    ('fb25_opaque', 'MultiGrid-GoalLastFewerBlocksOpaqueWallsAdversarial-v0'),
    ('var60', 'MultiGrid-GoalLastVariableBlocksAdversarialEnv-v0'),
    ('goalfirst50', 'MultiGrid-Adversarial-v0'),
    ('empty0', 'MultiGrid-GoalLastEmptyAdversarialEnv-Edit-v0'),
    ('mini6', 'MultiGrid-MiniGoalLastAdversarial-v0'),
]


def gen_adversary():
    """reset() + step_adversary builds, then reset_to_level round trips (string and byte forms)."""
    import numpy as np
    for tag, env_id in ADV_CASES:
        n_env = 8
        rs = np.random.RandomState(7)
        rec = {k: [] for k in ('locs', 'images', 'time_steps', 'dones', 'metrics', 'encoding', 'agent_obs', 'agent_dir',
                               'start_dir', 'replay_str_enc', 'replay_str_dir', 'replay_byte_enc', 'replay_byte_dir',
                               'replay_str_obs', 'replay_byte_obs', 'rng_tail', 'replay_metrics')}
        for i in range(n_env):
            env = rh.make_env(env_id, seed=i)
            env.seed(100 + i)
            o = env.reset()
            S = env.adversary_observation_space['time_step'].high[0]
            A = env.adversary_action_space.n
            locs = rs.randint(0, A, size=S)
            if i % 2 == 1 and S >= 4:  # force collisions: goal on a wall, agent on the goal
                gi = S - 2 if env.choose_goal_last else 0
                ai = S - 1 if env.choose_goal_last else 1
                if env.choose_goal_last and not env.resample_n_clutter:
                    locs[gi] = locs[0]
                locs[ai] = locs[gi]
            if env.resample_n_clutter:
                locs[0] = [100, 0, 168, 30, 85, 2, 140, 60][i]
            imgs, ts, dn = [np.array(o['image'])], [int(o['time_step'][0])], []
            for a in locs:
                o, r, d, info = env.step_adversary(int(a))
                imgs.append(np.array(o['image'])); ts.append(int(o['time_step'][0])); dn.append(bool(d))
            rec['locs'].append(locs); rec['images'].append(np.stack(imgs)); rec['time_steps'].append(ts); rec['dones'].append(dn)
            rec['metrics'].append([env.n_clutter_placed, env.distance_to_goal, int(env.passable), env.shortest_path_length,
                                   env.adversary_max_steps])
            rec['start_dir'].append(env.agent_start_dir)
            ao = env.reset_agent()
            rec['encoding'].append(env.encoding.copy())
            rec['agent_obs'].append(np.array(ao['image'], dtype=np.uint8)); rec['agent_dir'].append(int(ao['direction'][0]))
            # round trips: the SAME env object (keeps consuming its RNG for the fresh start dir)
            level_str = ' '.join(str(int(a)) for a in locs)
            ao = env.reset_to_level(level_str)
            rec['replay_str_enc'].append(env.encoding.copy()); rec['replay_str_dir'].append(env.agent_start_dir)
            rec['replay_str_obs'].append(np.array(ao['image'], dtype=np.uint8))
            enc_bytes = rec['encoding'][-1].tobytes()
            level = np.frombuffer(enc_bytes, dtype=np.uint8).reshape(rec['encoding'][-1].shape)
            ao = env.reset_to_level(level)
            rec['replay_byte_enc'].append(env.encoding.copy()); rec['replay_byte_dir'].append(env.agent_start_dir)
            rec['replay_byte_obs'].append(np.array(ao['image'], dtype=np.uint8))
            rec['replay_metrics'].append([env.n_clutter_placed, env.distance_to_goal, int(env.passable), env.shortest_path_length])
            rec['rng_tail'].append(env.np_random.randint(0, 2 ** 32, 4, dtype=np.uint64).astype(np.uint32))
        np.savez_compressed(os.path.join(GOLDEN, 'adversary_%s.npz' % tag), env_name=env_id, seeds=100 + np.arange(n_env),
                            **{k: np.array(v) for k, v in rec.items()})
        print('adversary', tag, 'metrics', rec['metrics'][:3])


def gen_mutate():
    """mutate_level with the global-np.random draws logged (adversarial.py:317-397)."""
    import numpy as np
    for tag, env_id in (('wng60', 'MultiGrid-GoalLastVariableBlocksAdversarialEnv-Edit-v0'),
                        ('wnag25', 'MultiGrid-GoalLastFewerBlocksAdversarial-v0'),
                        ('empty', 'MultiGrid-GoalLastEmptyAdversarialEnv-Edit-v0'),
                        ('wn25', 'MultiGrid-GoalLastFewerBlocksAdversarial-EditWN-v0')):
        rec = {k: [] for k in ('base_enc', 'locs', 'ops', 'n_edits', 'goal_choice', 'agent_choice', 'need', 'out_enc',
                               'metrics', 'obs', 'dirs', 'np_seed', 'num_edits')}
        np.random.seed(2024)
        n_cases = 48
        for i in range(n_cases):
            env = rh.make_env(env_id, seed=i)
            env.seed(i)
            env.reset_random()
            base = env.encoding.copy()
            num_edits = [1, 5, 5, 12, 30][i % 5]
            log = {'randint': [], 'choice': []}
            o_randint, o_choice = np.random.randint, np.random.choice

            def randint(*a, **k):
                v = o_randint(*a, **k); log['randint'].append(np.array(v).copy()); return v

            def choice(a, *aa, **k):
                v = o_choice(a, *aa, **k); log['choice'].append((np.array(a).copy(), int(v))); return v
            np.random.randint, np.random.choice = randint, choice
            try:
                chain = 1 + (i % 3)
                for c in range(chain):  # chained mutations of the same level
                    log['randint'].clear(); log['choice'].clear()
                    if c > 0:
                        base = env.encoding.copy()
                    last_seed = 5000 + i * 10 + c
                    np.random.seed(last_seed)
                    o = env.mutate_level(num_edits=num_edits)
            finally:
                np.random.randint, np.random.choice = o_randint, o_choice
            raw_locs, ops = log['randint'][0], log['randint'][1]
            order = list(set(raw_locs))  # the reference's own iteration order (adversarial.py:327)
            I = env.width - 2
            gc, ac, need = 0, 0, [0, 0]
            ch = list(log['choice'])
            # which fallbacks fired: goal first, then agent (adversarial.py:371-388)
            enc_now = env.encoding
            if len(ch) == 2:
                need = [1, 1]
                gc = int(np.nonzero(ch[0][0] == ch[0][1])[0][0]); ac = int(np.nonzero(ch[1][0] == ch[1][1])[0][0])
            elif len(ch) == 1:
                gx = ch[0][1] % I + 1; gy = ch[0][1] // I + 1
                idx = int(np.nonzero(ch[0][0] == ch[0][1])[0][0])
                if enc_now[gx, gy, 0] == 8:
                    need = [1, 0]; gc = idx
                else:
                    need = [0, 1]; ac = idx
            pad = 32
            rec['base_enc'].append(base)
            rec['locs'].append(np.pad(np.array(order, dtype=np.int32), (0, pad - len(order))))
            rec['ops'].append(np.pad(np.array(ops, dtype=np.int32), (0, pad - len(order))))
            rec['n_edits'].append(len(order)); rec['goal_choice'].append(gc); rec['agent_choice'].append(ac); rec['need'].append(need)
            rec['out_enc'].append(env.encoding.copy())
            rec['metrics'].append([env.n_clutter_placed, env.distance_to_goal, int(env.passable), env.shortest_path_length])
            rec['obs'].append(np.array(o['image'], dtype=np.uint8)); rec['dirs'].append(int(o['direction'][0]))
            rec['np_seed'].append(last_seed); rec['num_edits'].append(num_edits)
        np.savez_compressed(os.path.join(GOLDEN, 'mutate_%s.npz' % tag), env_name=env_id,
                            n_editor_actions=len(env.editor_actions), **{k: np.array(v) for k, v in rec.items()})
        print('mutate', tag, 'fallbacks', np.array(rec['need']).sum(0))


def gen_images():
    """venv.get_images() (parallel_wrappers.py:187-193 -> render(mode='level')) of the REAL vectorised env: RGB level
    screenshots with the agent's view highlighted, after reset_random / reset_agent and after a few steps (so that agents
    face every direction), see-through and occluded; plus the adversary phase (empty grid, no agent -> no highlight)."""
    import numpy as np
    import torch
    from types import SimpleNamespace
    import util
    out = {}
    for tag, env_name in (('gl15', 'MultiGrid-GoalLastAdversarial-v0'),
                          ('fb15_opaque', 'MultiGrid-GoalLastFewerBlocksOpaqueWallsAdversarial-v0')):
        N = 4
        args = SimpleNamespace(env_name=env_name, seed=1, singleton_env=False, use_global_critic=False,
                               use_global_policy=False, num_processes=N, normalize_returns=False)
        venv, _ = util.create_parallel_env(args)
        venv.set_seed(list(range(N)))
        venv.reset_random()
        venv.reset_agent()
        out[tag + '_env_name'] = env_name
        out[tag + '_enc0'] = np.stack(venv.get_encodings())
        out[tag + '_img0'] = np.stack(venv.get_images())
        rs = np.random.RandomState(5)
        acts = np.stack([_biased_actions(rs, N) for _ in range(9)])
        for t in range(9):
            venv.step_env(torch.from_numpy(acts[t].astype(np.int64)).view(N, 1), reset_random=False)
        out[tag + '_actions'] = acts
        out[tag + '_img1'] = np.stack(venv.get_images())
        venv.reset()
        venv.step_adversary(torch.tensor([[3], [20], [77], [100]]))
        out[tag + '_img_adv'] = np.stack(venv.get_images())
        venv.close()
        print('images', tag, out[tag + '_img0'].shape)
    np.savez_compressed(os.path.join(GOLDEN, 'level_images.npz'), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    a = ap.parse_args()
    rh.activate()
    os.makedirs(GOLDEN, exist_ok=True)
    todo = {'env': gen_env_traces, 'venv': gen_venv, 'fullobs': gen_fullobs, 'adv': gen_adversary, 'mutate': gen_mutate,
            'images': gen_images}
    try:
        from gen_golden_plr import gen_plr
        todo['plr'] = gen_plr
        from gen_golden_runner import gen_runner
        todo['runner'] = gen_runner
    except ImportError:
        pass
    for k, fn in todo.items():
        if a.only in (None, k):
            fn()


if __name__ == '__main__':
    main()
