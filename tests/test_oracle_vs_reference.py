"""LIVE diff of the C oracle against the reference's own modules (executed in-process over oracle/shim) on random
levels that the committed fixtures do not cover: dense / impassable mazes (BFS 'unreachable' branch), occlusion on
cluttered views, both grid sizes.  Runs only where /root/reference exists (this container), never on the GPU box."""
import numpy as np
import pytest

from oracle import mg_oracle as mo
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason='reference tree not present')


def _random_encoding(rs, W, density):
    enc = np.zeros((W, W, 3), np.uint8)
    enc[:, :, 0] = 1
    enc[0, :, :] = enc[-1, :, :] = enc[:, 0, :] = enc[:, -1, :] = (2, 5, 0)
    interior = [(x, y) for x in range(1, W - 1) for y in range(1, W - 1)]
    rs.shuffle(interior)
    n_w = int(density * len(interior))
    for (x, y) in interior[:n_w]:
        enc[x, y] = (2, 5, 0)
    gx, gy = interior[n_w]
    ax, ay = interior[n_w + 1]
    enc[gx, gy] = (8, 1, 0)
    enc[ax, ay] = (10, 0, rs.randint(0, 4))
    return enc


@pytest.mark.parametrize('W,see', [(15, False), (15, True), (25, False)])
def test_random_dense_levels(W, see):
    rh.activate()
    import importlib
    adv = importlib.import_module('envs.multigrid.adversarial')
    from envs.wrappers import TimeLimit
    rs = np.random.RandomState(W + see)
    cfg = mo.make_cfg(W=W, see_through=see, n_clutter=50)
    n_cases = 40
    b = mo.OracleBatch(cfg, n_cases)
    n_impassable = 0
    for i in range(n_cases):
        enc = _random_encoding(rs, W, density=[0.1, 0.35, 0.5, 0.65][i % 4])
        env = TimeLimit(adv.AdversarialEnv(size=W, n_clutter=50, choose_goal_last=True, see_through_walls=see, seed=i,
                                           max_steps=250), max_episode_steps=250)
        env.seed(i)
        b.seed(i, i)
        o = env.reset_to_level(enc)
        assert b.reset_to_encoding(i, enc) == 0
        s = b.state(i)
        assert [s['n_clutter_placed'], s['dist'], s['passable'], s['spl']] == \
            [env.n_clutter_placed, env.distance_to_goal, int(env.passable), env.shortest_path_length], i
        n_impassable += not env.passable
        assert np.array_equal(b.encode(i), env.encoding)
        assert np.array_equal(b.gen_obs(i), np.array(o['image'], np.uint8))
        acts = rs.randint(0, 7, size=60)
        acts[rs.rand(60) < 0.5] = 2
        for a in acts:
            o, r, d, info = env.step(int(a))
            if d:
                o = env.reset_agent()
            res = b.step_env(i, int(a))
            assert np.array_equal(res['obs'], np.array(o['image'], np.uint8)), i
            assert res['rew'] == np.float32(r) and bool(res['flags'] & 1) == bool(d)
    assert n_impassable > 3  # the unreachable branch was exercised


def test_runner_fixture_reproducible():
    """The runner-level fixtures are outputs of the reference's own AdversarialRunner.run() over its spawn-subprocess
    vector env (oracle/gen_golden_runner.py): re-executing one case here must reproduce the committed file exactly."""
    import gzip
    import os
    import pickle
    from conftest import GOLDEN
    rh.activate()
    import oracle.gen_golden_runner as gr
    import util
    from algos.storage import RolloutStorage
    from envs.runners.adversarial_runner import AdversarialRunner
    name = 'runner_dr_mini_nohtl'
    with gzip.open(os.path.join(GOLDEN, name + '.pkl.gz'), 'rb') as f:
        g = pickle.load(f)
    args = gr.parse_args(g['argv'])
    venv, ued_venv = util.create_parallel_env(args)
    try:
        runner, agent, _ = gr.build_runner(args, venv, ued_venv, RolloutStorage, AdversarialRunner, None)
        runs = gr.run_case(name, runner, agent, g['n_runs'], g['np_seed'])
    finally:
        venv.close()
    for got, want in zip(runs, g['runs']):
        for k, v in want['storage'].items():
            assert np.array_equal(got['storage'][k], v), k
        assert np.array_equal(got['encodings'], want['encodings'])
        assert got['total_episodes'] == want['total_episodes']


def test_tile_table_matches_reference_render_tile():
    """dcd_isaac_b200.tiles (the 14-tile table behind venv.get_images) == the reference's Grid.render_tile
    (multigrid.py:159-214) over the restated gym-minigrid rendering helpers, pixel for pixel."""
    rh.activate()
    from envs.multigrid import multigrid as mg
    from gym_minigrid import minigrid
    from dcd_isaac_b200 import tiles
    table = tiles.tile_table()
    objs = [None, minigrid.Wall(), minigrid.Goal()] + [mg.Agent(0, d) for d in range(4)]
    for code, obj in enumerate(objs):
        for hl in (False, True):
            mg.Grid.tile_cache = {}
            ref = mg.Grid.render_tile(obj, highlight=[np.bool_(hl)], tile_size=32, cell_type=(obj.type if obj else None))
            img = np.zeros((32, 32, 3), np.uint8)
            img[:] = ref
            assert np.array_equal(img, table[2 * code + int(hl)]), (code, hl)
