"""Integration: the pieces composed the way envs/runners/adversarial_runner.py composes them for robust PLR with a
random level generator (DR + PLR without use_reset_random_dr: levels built by a uniform-random adversary through
step_adversary and stored as action strings, adversarial_runner.py:455-482,402-412; replay via
sample_replay_level + reset_to_level_batch; per-done re-sampling inside the rollout, :551-558; scoring and
reconciliation, :616-622,797-800).  The student is a random policy: the NN is out of scope."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def test_robust_plr_cycles_compose():
    from dcd_isaac_b200.level_sampler import LevelSampler
    from dcd_isaac_b200.level_store import LevelStore
    from dcd_isaac_b200.storage import DeviceRolloutStorage
    from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv
    N, T, NB = 32, 64, 48
    np.random.seed(0)
    torch.manual_seed(0)
    venv = CudaAdversarialVecEnv('MultiGrid-GoalLastFewerBlocksAdversarial-v0', N)
    venv.set_seed(list(range(N)))
    S = int(venv.adversary_observation_space['time_step'].high[0])
    A = venv.adversary_action_space.n
    sampler = LevelSampler([], venv.observation_space, venv.action_space, num_actors=N, strategy='positive_value_loss',
                           replay_schedule='proportionate', score_transform='rank', temperature=0.1, rho=0.5,
                           replay_prob=0.5, staleness_coef=0.3, sample_full_distribution=True, seed_buffer_size=NB,
                           seed_buffer_priority='replay_support', gamma=0.995)
    store = LevelStore()
    storage = DeviceRolloutStorage(T, N)
    n_replay, n_new, episodes = 0, 0, 0
    for cycle in range(14):
        replay = sampler.sample_replay_decision()
        if replay:
            n_replay += 1
            seeds = [sampler.sample_replay_level() for _ in range(N)]
            venv.reset_to_level_batch([store.get_level(s) for s in seeds])
        else:
            n_new += 1
            obs = venv.reset()
            assert obs['image'].shape == (N, 3, 15, 15)
            traj = []
            for k in range(S):  # uniform-random adversary (models/multigrid_models.py:149-157 random mode)
                a = torch.randint(0, A, (N, 1))
                traj.append(a)
                obs, r, d, infos = venv.step_adversary(a)
            assert d.all()
            levels = [' '.join(str(int(traj[k][i, 0])) for k in range(S)) for i in range(N)]
            seeds = store.insert(levels)
            sampler.observe_external_unseen_sample(seeds, solvable=venv.get_passable())
        cur = list(seeds)
        obs = venv.reset_agent()
        storage.obs['image'][0].copy_(obs['image'])
        for t in range(T):
            action = torch.randint(0, 7, (N, 1))
            action[torch.rand(N, 1) < 0.5] = 2
            obs, reward, done, infos = venv.step_env(action, reset_random=False)
            if t == T - 1:
                done = np.ones_like(done)
            storage.level_seeds[t].copy_(torch.tensor(cur, dtype=torch.int32).view(-1, 1))
            for i, info in enumerate(infos):
                if 'episode' in info:
                    episodes += 1
                    if replay:  # per-done replay re-sampling (adversarial_runner.py:551-558)
                        cur[i] = sampler.sample_replay_level()
                        obs_i = venv.reset_to_level(store.get_level(cur[i]), i)
                        for k in obs:
                            obs[k][i] = obs_i[k].squeeze(0)
            storage.obs['image'][t + 1].copy_(obs['image'])
            storage.rewards[t].copy_(reward)
            storage.masks[t + 1].copy_(torch.from_numpy(1.0 - done.astype(np.float32)).view(-1, 1))
            storage.value_preds[t].uniform_(0, 1)
        storage.compute_returns(torch.zeros(N, 1, device='cuda'), True, 0.995, 0.95)
        sampler.update_with_rollouts(storage)
        sampler.after_update()
        store.reconcile_seeds(set(int(x) for x in sampler.seeds if x >= 0))
        assert set(int(x) for x in sampler.seeds if x >= 0) <= set(store.seed2level)
        assert np.isfinite(sampler.seed_scores).all()
        if (sampler.unseen_seed_weights < 1).any():
            w = sampler.sample_weights()
            assert abs(w.sum() - 1) < 1e-9 and (w[sampler.unseen_seed_weights > 0] == 0).all()
    assert n_new >= 2 and n_replay >= 1 and episodes > 0
    assert sampler.working_seed_buffer_size == NB  # the buffer filled (NB < levels generated) and admission kept working
    assert set(int(x) for x in sampler.seeds if x >= 0) <= set(store.seed2level)
    venv.close()
