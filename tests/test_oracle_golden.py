"""The C oracle (oracle/c/mg_oracle.c) against fixtures produced by EXECUTING the reference
(oracle/gen_golden.py).  CPU only.  This is what pins the oracle."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden
from oracle import mg_oracle as mo


def _cfg_from_trace(g, **kw):
    return mo.make_cfg(W=int(g['W']), max_steps=int(g['max_steps']), max_episode_steps=int(g['time_limit']),
                       see_through=bool(g['see_through']), n_clutter=int(g['n_clutter']), **kw)


TRACES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, 'env_trace_*.npz')))


@pytest.mark.parametrize('name', TRACES)
def test_env_trace(name):
    g = golden(name)
    reset_random = name.endswith('_random.npz')
    n_env, T = g['actions'].shape
    b = mo.OracleBatch(_cfg_from_trace(g, fixed_environment=bool(g['fixed'])), n_env)
    for i in range(n_env):
        b.seed(i, int(g['seeds'][i]))
        assert b.reset_random(i) == 0
        assert np.array_equal(b.encode(i), g['encodings'][i])
        s = b.state(i)
        assert [s['n_clutter_placed'], s['dist'], s['passable'], s['spl']] == list(g['metrics'][i])
        assert np.array_equal(b.gen_obs(i), g['first_obs'][i])
        for t in range(T):
            r = b.step_env(i, g['actions'][i, t], reset_random=reset_random)
            assert r['flags'] == g['flags'][i, t], (i, t)
            assert np.array_equal(r['obs'], g['obs'][i, t]), (i, t)
            assert r['dir'] == g['dirs'][i, t]
            assert r['rew'] == g['rewards'][i, t]
            if r['flags'] & 2:
                assert np.array_equal(r['trunc_obs'], g['trunc_obs'][i, t]), (i, t)
                assert r['trunc_dir'] == g['trunc_dir'][i, t]
            s = b.state(i)
            assert [s['ax'], s['ay'], s['adir'], s['step_count']] == list(g['pos'][i, t]), (i, t)
        assert [b.rng_next(i) for _ in range(4)] == list(g['rng_tail'][i])


VENVS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, 'venv_*.npz')))


@pytest.mark.parametrize('name', VENVS)
def test_venv_trace(name):
    """Fold of the real subprocess vector env + VecMonitor + VecPreprocessImageWrapper."""
    g = golden(name)
    reset_random = name.endswith('_random.npz')
    cfg = mo.cfg_from_name(str(g['env_name']))
    T, N = g['actions'].shape
    b = mo.OracleBatch(cfg, N)
    for i in range(N):
        b.seed(i, i)
        b.reset_random(i)
        assert np.array_equal(b.encode(i), g['encodings'][i])
        b.reset_agent(i)
        assert np.array_equal(mo.preprocess(b.gen_obs(i)), g['first_image'][i])
        assert float(b.state(i)['adir']) == g['first_direction'][i, 0]
    for t in range(T):
        for i in range(N):
            r = b.step_env(i, g['actions'][t, i], reset_random=reset_random)
            assert np.array_equal(mo.preprocess(r['obs']), g['image'][t, i]), (t, i)
            assert float(r['dir']) == g['direction'][t, i, 0]
            assert r['rew'] == g['reward'][t, i, 0]
            assert bool(r['flags'] & 1) == bool(g['done'][t, i])
            assert bool(r['flags'] & 2) == bool(g['trunc_key'][t, i])
            assert bool(r['flags'] & 4) == bool(g['trunc_val'][t, i])
            if r['flags'] & 1:
                assert r['ep_r'] == g['ep_r'][t, i] and r['ep_l'] == g['ep_l'][t, i]
            if r['flags'] & 2:
                assert np.array_equal(mo.preprocess(r['trunc_obs']), g['trunc_image'][t, i])
                assert float(r['trunc_dir']) == g['trunc_direction'][t, i]


ADVS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, 'adversary_*.npz')))


@pytest.mark.parametrize('name', ADVS)
def test_adversary_build(name):
    g = golden(name)
    cfg = mo.cfg_from_name(str(g['env_name']))
    n_env, S = g['locs'].shape
    b = mo.OracleBatch(cfg, n_env)
    for i in range(n_env):
        b.seed(i, int(g['seeds'][i]))
        b.reset(i)
        assert np.array_equal(b.encode(i), g['images'][i, 0])
        for k in range(S):
            done = b.step_adversary(i, g['locs'][i, k])
            assert np.array_equal(b.encode(i), g['images'][i, k + 1]), (i, k)
            assert b.state(i)['adv_step'] == g['time_steps'][i][k + 1]
            assert done == bool(g['dones'][i, k])
        s = b.state(i)
        assert [s['n_clutter_placed'], s['dist'], s['passable'], s['spl'], s['adv_max']] == list(g['metrics'][i])
        assert s['sdir'] == g['start_dir'][i]
        assert b.reset_agent(i) == 0
        assert np.array_equal(b.encode(i), g['encoding'][i])
        assert np.array_equal(b.gen_obs(i), g['agent_obs'][i]) and b.state(i)['adir'] == g['agent_dir'][i]
        # reset_to_level round trips on the same env (fresh start dir each time)
        assert b.reset_to_actions(i, g['locs'][i]) == 0
        assert np.array_equal(b.encode(i), g['replay_str_enc'][i]) and b.state(i)['sdir'] == g['replay_str_dir'][i]
        assert np.array_equal(b.gen_obs(i), g['replay_str_obs'][i])
        assert b.reset_to_encoding(i, g['encoding'][i]) == 0
        assert np.array_equal(b.encode(i), g['replay_byte_enc'][i]) and b.state(i)['sdir'] == g['replay_byte_dir'][i]
        assert np.array_equal(b.gen_obs(i), g['replay_byte_obs'][i])
        s = b.state(i)
        assert [s['n_clutter_placed'], s['dist'], s['passable'], s['spl']] == list(g['replay_metrics'][i])
        assert [b.rng_next(i) for _ in range(4)] == list(g['rng_tail'][i])


def test_step_adversary_out_of_range():
    b = mo.OracleBatch(mo.make_cfg(), 1)
    b.seed(0, 0)
    b.reset(0)
    with pytest.raises(ValueError):
        b.step_adversary(0, 169)


MUTS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, 'mutate_*.npz')))


@pytest.mark.parametrize('name', MUTS)
def test_mutate_level(name):
    g = golden(name)
    base = mo.cfg_from_name(str(g['env_name']))
    base.n_editor_actions = int(g['n_editor_actions'])
    n = len(g['n_edits'])
    b = mo.OracleBatch(base, n)
    for i in range(n):
        b.seed(i, 5)
        assert b.reset_to_encoding(i, g['base_enc'][i]) == 0
        k = int(g['n_edits'][i])
        rc, need, _ = b.mutate(i, g['locs'][i, :k], g['ops'][i, :k], g['goal_choice'][i], g['agent_choice'][i])
        assert rc == 0
        assert need == list(g['need'][i]), i
        got, want = b.encode(i), g['out_enc'][i].copy()
        # the start direction is env-RNG state, not part of the mutation: compare modulo the dir byte
        sx, sy = b.state(i)['sx'], b.state(i)['sy']
        want[sx, sy, 2] = got[sx, sy, 2]
        assert np.array_equal(got, want), i
        s = b.state(i)
        assert [s['n_clutter_placed'], s['dist'], s['passable'], s['spl']] == list(g['metrics'][i]), i
