"""The reference's own train.py and eval.py, UNMODIFIED (from /root/reference or the staged copy baseline/_ref/reference),
run on the drop-in: `python -m dcd_isaac_b200.dropin <reference>/train.py <the reference's flags>` installs the import swaps
of INTEGRATION.md and runpy's the script.  train.py runs with DEFAULT screenshot / test settings (a screenshot at update 0
through venv.get_images(), the in-training Evaluator on the three default test mazes) for two PPO updates of robust PLR and
writes logs.csv, the screenshot and a checkpoint; eval.py then loads that checkpoint and evaluates it on the zero-shot maze
benchmark's env list (eval.py:340-349 -- passed explicitly, the reference's list is missing a comma) for one episode each."""
import csv
import glob
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

MAZE_BENCHMARK = ['MultiGrid-SixteenRooms-v0', 'MultiGrid-SixteenRoomsFewerDoors-v0', 'MultiGrid-Labyrinth-v0', 'MultiGrid-Labyrinth2-v0',
                  'MultiGrid-Maze-v0', 'MultiGrid-Maze2-v0', 'MultiGrid-LargeCorridor-v0', 'MultiGrid-PerfectMazeMedium-v0',
                  'MultiGrid-PerfectMazeLarge-v0', 'MultiGrid-PerfectMazeXL-v0']


def _env_and_ref():
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip('no reference tree (neither /root/reference nor baseline/_ref/reference)')
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([rh.SHIM, rh.REFERENCE, ROOT, env.get('PYTHONPATH', '')])
    env['OMP_NUM_THREADS'] = '1'
    # eval.py:435 calls torch.load(path, map_location='cpu') on a checkpoint that pickles the runner's LevelSampler / LevelStore
    # objects (adversarial_runner.py:214-215); torch >= 2.6 defaults to weights_only=True, which the reference predates
    env['TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD'] = '1'
    return env, rh.REFERENCE


def test_reference_train_and_eval_scripts_run_on_the_dropin(tmp_path):
    env, ref = _env_and_ref()
    log_dir = str(tmp_path / 'logs')
    common = [sys.executable, '-m', 'dcd_isaac_b200.dropin']
    train = common + [os.path.join(ref, 'train.py'), '--xpid', 'dropin_test', '--log_dir', log_dir,
                      '--env_name', 'MultiGrid-GoalLastFewerBlocksAdversarial-v0', '--ued_algo', 'domain_randomization',
                      '--use_plr', 'true', '--level_replay_strategy', 'positive_value_loss', '--level_replay_score_transform', 'rank',
                      '--level_replay_temperature', '0.3', '--level_replay_rho', '0.5', '--level_replay_prob', '0.5',
                      '--level_replay_seed_buffer_size', '16', '--staleness_coef', '0.3', '--no_exploratory_grad_updates', 'true',
                      '--handle_timelimits', 'true', '--num_processes', '8', '--num_steps', '64', '--num_env_steps', '1536',
                      '--ppo_epoch', '1', '--test_num_episodes', '2', '--checkpoint_interval', '1', '--recurrent_hidden_size', '32',
                      '--verbose']
    out = subprocess.run(train, capture_output=True, text=True, timeout=900, env=env, cwd=str(tmp_path))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    xp = os.path.join(log_dir, 'dropin_test')
    rows = list(csv.DictReader(open(os.path.join(xp, 'logs.csv'))))
    rows = [r for r in rows if r.get('steps') not in (None, '', 'steps')]
    assert len(rows) == 3, rows                                   # three updates logged
    assert int(float(rows[-1]['steps'])) == 3 * 8 * 64
    assert 'solved_rate:MultiGrid-SixteenRooms-v0' in rows[0] and rows[0]['solved_rate:MultiGrid-SixteenRooms-v0'] != ''
    assert float(rows[-1]['sps']) > 0
    shots = glob.glob(os.path.join(xp, 'screenshots', 'update0*.png'))
    assert len(shots) == 1 and os.path.getsize(shots[0]) > 1000    # venv.get_images() -> save_images (train.py:204-232)
    assert os.path.exists(os.path.join(xp, 'model.tar'))           # runner state incl. the pickled LevelSampler / LevelStore
    ck = torch.load(os.path.join(xp, 'model.tar'), map_location='cpu', weights_only=False)
    st = ck['runner_state_dict']
    assert type(st['level_samplers']['agent']).__module__ == 'dcd_isaac_b200.level_sampler'
    assert type(st['level_store']).__module__ == 'dcd_isaac_b200.level_store' and st['num_updates'] == 3
    # ---- eval.py on the checkpoint: the zero-shot maze benchmark, one episode per env
    res_dir = str(tmp_path / 'results')
    ev = common + [os.path.join(ref, 'eval.py'), '--base_path', log_dir, '--xpid', 'dropin_test', '--model_tar', 'model',
                   '--env_names', ','.join(MAZE_BENCHMARK), '--num_episodes', '1', '--num_processes', '1', '--result_path', res_dir]
    out = subprocess.run(ev, capture_output=True, text=True, timeout=1500, env=env, cwd=str(tmp_path))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    res = glob.glob(os.path.join(res_dir, '*.csv'))
    assert len(res) == 1
    table = {r[0]: r[1:] for r in csv.reader(open(res[0]))}
    for name in MAZE_BENCHMARK:   # (solved_rate rows only come with --accumulator mean, eval.py:331-338)
        assert 'test_returns:' + name in table and len(table['test_returns:' + name]) == 1, (name, list(table)[:6])
