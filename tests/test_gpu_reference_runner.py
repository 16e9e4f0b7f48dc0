"""Runner-level parity: the reference's OWN AdversarialRunner (envs/runners/adversarial_runner.py:442-896, run
unmodified from /root/reference or the staged copy baseline/_ref/reference) driving the DROP-IN objects --
CudaAdversarialVecEnv, LevelSampler, LevelStore and (parametrised) the drop-in or the reference RolloutStorage -- must
reproduce the fixtures that the same runner produced over the reference's subprocess envs / sampler / store
(oracle/gen_golden_runner.py).  Then the same fixtures pin the DEVICE path that bench.py times: the masks, bad_masks,
cliffhanger_masks, rewards and observations written by the step kernel itself (step_env_device(last_step=...) and
mgplr_rollout_ex) are compared with what the reference runner put into its storage.

Bit-exact: observations, rewards, masks, bad_masks, cliffhanger_masks, level_seeds, actions, level encodings, seeds,
store contents, replay decisions.  Returns / truncated value predictions: 1e-6 relative (same op order, fp32); sampler
scores 1e-5 relative (the reference averages in torch fp32, the score kernel in fp32 with a different summation tree)."""
import glob
import gzip
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

CASES = sorted(os.path.basename(p)[:-len('.pkl.gz')] for p in glob.glob(os.path.join(GOLDEN, 'runner_*.pkl.gz')))
EXACT = ('obs_image', 'obs_direction', 'rewards', 'masks', 'bad_masks', 'cliffhanger_masks', 'level_seeds', 'actions', 'value_preds')
CLOSE = ('returns', 'truncated_value_preds')


def _load(name):
    with gzip.open(os.path.join(GOLDEN, name + '.pkl.gz'), 'rb') as f:
        return pickle.load(f)


def _reference_modules():
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip('no reference tree (neither /root/reference nor baseline/_ref/reference: run oracle/stage_reference.py)')
    rh.activate()
    import oracle.gen_golden_runner as gr
    return gr


def _compare_storage(got, want, tag):
    for k in EXACT:
        if k in want:
            assert np.array_equal(got[k], want[k]), '%s: %s differs at %s' % (tag, k, np.argwhere(got[k] != want[k])[:4].tolist())
    for k in CLOSE:
        if k in want:
            np.testing.assert_allclose(got[k], want[k], rtol=1e-6, atol=1e-7, err_msg='%s: %s' % (tag, k))
    # truncated observations: rows the runner inserted (episode ended by the TimeLimit, adversarial_runner.py:546-549)
    if 'truncated_obs_image' in want:
        ins = (want['bad_masks'][1:] == 0) & (want['masks'][1:] == 0) & (want['cliffhanger_masks'][1:] == 1)
        ins = np.concatenate([np.zeros_like(ins[:1]), ins])[..., 0]
        assert np.array_equal(got['truncated_obs_image'][ins], want['truncated_obs_image'][ins]), tag
        assert np.array_equal(got['truncated_obs_direction'][ins], want['truncated_obs_direction'][ins]), tag


@pytest.mark.parametrize('storage_kind', ['dropin', 'reference'])
@pytest.mark.parametrize('name', CASES)
def test_reference_runner_over_dropin(name, storage_kind):
    gr = _reference_modules()
    from dcd_isaac_b200 import dropin
    g = _load(name)
    dropin.install(device='cuda:0')
    try:
        import util
        from envs.runners.adversarial_runner import AdversarialRunner
        import envs.runners.adversarial_runner as ar
        from dcd_isaac_b200.level_sampler import LevelSampler
        from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv
        assert ar.LevelSampler is LevelSampler  # the runner builds the drop-in sampler / store
        if storage_kind == 'dropin':
            from dcd_isaac_b200.storage import RolloutStorage
        else:
            from algos.storage import RolloutStorage
        args = gr.parse_args(g['argv'])
        venv, ued_venv = util.create_parallel_env(args)
        assert isinstance(venv, CudaAdversarialVecEnv) and ued_venv is venv
        plr_args = util.make_plr_args(args, venv.observation_space, venv.action_space) if args.use_plr else None
        runner, agent, _ = gr.build_runner(args, venv, ued_venv, RolloutStorage, AdversarialRunner, plr_args, device='cuda:0')
        runs = gr.run_case(name, runner, agent, g['n_runs'], g['np_seed'])
        venv.close()
    finally:
        dropin.uninstall()
    for r, (got, want) in enumerate(zip(runs, g['runs'])):
        tag = '%s run %d' % (name, r)
        assert got['level_replay'] == want['level_replay'], tag
        _compare_storage(got['storage'], want['storage'], tag)
        for k in ('current_level_seeds', 'total_episodes', 'total_seeds', 'student_grad_updates', 'n_snapshots'):
            assert got[k] == want[k], (tag, k, got[k], want[k])
        assert np.array_equal(got['encodings'], want['encodings']), tag
        assert abs(got['mean_agent_return'] - want['mean_agent_return']) <= 1e-6 * max(1.0, abs(want['mean_agent_return']))
        for k, v in want['stats'].items():
            assert (got['stats'][k] is None) == (v is None), (tag, k)
            if v is not None:
                assert abs(got['stats'][k] - v) <= 1e-9 * max(1.0, abs(v)), (tag, k)
        if 'sampler' in want:
            gs, ws = got['sampler'], want['sampler']
            for k in ('seeds', 'unseen_seed_weights', 'seed_staleness'):
                assert np.array_equal(gs[k], ws[k]), (tag, k, gs[k], ws[k])
            np.testing.assert_allclose(gs['seed_scores'], ws['seed_scores'], rtol=1e-5, atol=1e-7, err_msg=tag)
            for k in ('working_seed_set', 'staging_seed_set', 'running_sample_count', 'working_seed_buffer_size'):
                assert gs[k] == ws[k], (tag, k)
            assert got['store_seeds'] == want['store_seeds'], tag
            assert got['store_levels'] == want['store_levels'], tag


DR_CASES = [c for c in CASES if c.startswith('runner_dr_')]


@pytest.mark.parametrize('path', ['step_env_device', 'rollout_ex'])
@pytest.mark.parametrize('name', DR_CASES)
def test_kernel_written_masks_vs_reference_runner(name, path):
    """The benchmarked device path: the step kernel writes masks / bad_masks / cliffhanger_masks (write_step_scalars,
    last-step rule included) straight into rollout storage.  Replaying a DR fixture's action stream through it must give
    the tensors the REFERENCE RUNNER wrote into its storage (no host loop, no infos)."""
    from dcd_isaac_b200.storage import DeviceRolloutStorage
    from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv
    g = _load(name)
    argv = g['argv']
    opt = {argv[i][2:]: argv[i + 1] for i in range(0, len(argv), 2)}
    N, T = int(opt['num_processes']), int(opt['num_steps'])
    htl = opt['handle_timelimits'] == 'true'
    v = CudaAdversarialVecEnv(opt['env_name'], N)
    v.set_seed(list(range(N)))   # util.create_parallel_env
    st = DeviceRolloutStorage(T, N, device='cuda:0')
    last = 1 | (2 if htl else 0)
    for r, want_run in enumerate(g['runs']):
        want = want_run['storage']
        v.reset_random()                      # agent_rollout(is_env=True): DR without PLR (adversarial_runner.py:467-470)
        obs = v.reset_agent()                 # :484-487
        st.obs['image'][0].copy_(obs['image'])
        st.obs['direction'][0].copy_(obs['direction'])
        acts = torch.from_numpy(want['actions'][:, :, 0]).cuda()
        if path == 'step_env_device':
            for t in range(T):
                v.step_env_device(acts[t].contiguous(), st.step_out(t), reset_random=True, last_step=last if t == T - 1 else 0)
        else:
            o = st.step_out(0)
            v.rollout_device(acts.to(torch.uint8).contiguous(), o, reset_random=True, last_step=last)
        torch.cuda.synchronize()
        tag = '%s run %d (%s)' % (name, r, path)
        got = {'obs_image': st.obs['image'], 'obs_direction': st.obs['direction'], 'rewards': st.rewards, 'masks': st.masks,
               'bad_masks': st.bad_masks, 'cliffhanger_masks': st.cliffhanger_masks}
        for k, t_ in got.items():
            a, b = t_.cpu().numpy(), want[k]
            if k in ('masks', 'bad_masks', 'cliffhanger_masks'):
                a, b = a[1:], b[1:]   # row 0 is the previous rollout's last row (after_update), not written by a step
            assert np.array_equal(a, b), '%s: %s differs at %s' % (tag, k, np.argwhere(a != b)[:4].tolist())
        if htl:
            ins = (want['bad_masks'][1:] == 0) & (want['masks'][1:] == 0) & (want['cliffhanger_masks'][1:] == 1)
            ins = np.concatenate([np.zeros_like(ins[:1]), ins])[..., 0]
            assert ins.any()
            assert np.array_equal(st.truncated_obs['image'].cpu().numpy()[ins], want['truncated_obs_image'][ins]), tag
            assert np.array_equal(st.truncated_obs['direction'].cpu().numpy()[ins], want['truncated_obs_direction'][ins]), tag
            # cliffhangers: the runner hands the env's current observation to info['truncated_obs'] (:528); the kernel
            # writes the same observation into the truncated-obs row
            cl = (want['cliffhanger_masks'][T] == 0)[..., 0]
            assert cl.any()
            assert np.array_equal(st.truncated_obs['image'][T].cpu().numpy()[cl], want['obs_image'][T][cl]), tag
        assert np.array_equal(np.stack(v.get_encodings()), want_run['encodings']), tag
    v.close()
