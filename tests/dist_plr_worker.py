#!/usr/bin/env python
"""Worker of the multi-GPU PLR equivalence test (tests/test_gpu_multi.py), also runnable by hand:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/dist_plr_worker.py --envs 64 --T 48 --cycles 5

Robust-PLR cycles on byte-encoded reset_random levels (the runner's use_reset_random_dr path, adversarial_runner.py:462-475,
497-635) with the env batch SHARDED over the ranks: each rank steps its own slice through the host API, the level encodings
cross ranks in ShardedPLR.insert_current_levels, the episode records in update_with_rollouts, the per-step done flags in
resample_finished.  Afterwards rank 0 repeats the same cycles UNSHARDED (all envs on its GPU) and every rank's replica --
sampler arrays, staging / working sets, level store contents -- and every rank's slice of the rollout tensors must equal the
single-process result bit for bit."""
import argparse
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def run_cycles(a, rank, world, group_ok):
    from dcd_isaac_b200.distributed import ShardedPLR
    from dcd_isaac_b200.level_sampler import LevelSampler
    from dcd_isaac_b200.level_store import LevelStore
    from dcd_isaac_b200.storage import DeviceRolloutStorage
    from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv
    N, T = a.envs, a.T
    n_local = N // world
    lo = rank * n_local
    dev = torch.device('cuda', torch.cuda.current_device())
    venv = CudaAdversarialVecEnv(a.env_name, n_local, device=dev)
    venv.set_seed(list(range(lo, lo + n_local)))   # util.create_parallel_env seeds env i with i
    sampler = LevelSampler([], venv.observation_space, venv.action_space, num_actors=N, strategy=a.strategy,
                           replay_schedule='fixed', score_transform='rank', temperature=0.3, rho=0.4, replay_prob=0.6,
                           staleness_coef=0.3, sample_full_distribution=True, seed_buffer_size=a.buffer,
                           seed_buffer_priority='replay_support', gamma=0.995, device=dev)
    example = venv.get_encodings(index=[0])[0]
    store = LevelStore(data_info={'numpy': True, 'dtype': example.dtype, 'shape': example.shape})
    plr = ShardedPLR(sampler, store, rank, world, n_local, device=dev if world > 1 else None)
    storage = DeviceRolloutStorage(T, n_local, device=dev)
    np.random.seed(a.seed)              # the trainer process's stream: replay decisions and draws, identical on every rank
    stream = np.random.RandomState(a.seed + 1)   # scripted policy: actions / values for ALL envs, every rank takes its slice
    log = []
    for cycle in range(a.cycles):
        replay = sampler.sample_replay_decision()
        if replay:
            seeds, levels = plr.sample_replay_levels()
            venv.reset_to_level_batch(levels)
        else:
            venv.reset_random()
            seeds = plr.insert_current_levels(venv.get_encodings_device(), solvable_local=venv.get_passable())
        obs = venv.reset_agent()
        storage.obs['image'][0].copy_(obs['image'])
        storage.obs['direction'][0].copy_(obs['direction'])
        for t in range(T):
            act = stream.randint(0, 7, size=N)
            act[stream.rand(N) < 0.5] = 2
            val = stream.rand(N).astype(np.float32)
            obs, reward, done, infos = venv.step_env(torch.from_numpy(act[lo:lo + n_local].astype(np.int64)).view(-1, 1))
            last = t == T - 1
            cliff = np.zeros(n_local, bool)
            if last:
                cliff = ~done
                done = np.ones_like(done)
            storage.level_seeds[t].copy_(torch.tensor(plr.current_level_seeds[lo:lo + n_local], dtype=torch.int32).view(-1, 1))
            if replay:   # (adversarial_runner.py:551-558; the forced dones of the last step are not episode ends)
                ended = np.array(['episode' in info for info in infos])
                for i, (s, level) in plr.resample_finished(ended).items():
                    obs_i = venv.reset_to_level(level, i)
                    for k in obs:
                        obs[k][i] = obs_i[k].squeeze(0)
            bad = np.array([('truncated' in info) for info in infos]) | cliff
            storage.obs['image'][t + 1].copy_(obs['image'])
            storage.obs['direction'][t + 1].copy_(obs['direction'])
            storage.rewards[t].copy_(reward)
            storage.masks[t + 1].copy_(torch.from_numpy(1.0 - done.astype(np.float32)).view(-1, 1))
            storage.bad_masks[t + 1].copy_(torch.from_numpy(1.0 - bad.astype(np.float32)).view(-1, 1))
            storage.cliffhanger_masks[t + 1].copy_(torch.from_numpy(1.0 - cliff.astype(np.float32)).view(-1, 1))
            storage.value_preds[t].copy_(torch.from_numpy(val[lo:lo + n_local]).view(-1, 1))
        nxt = stream.rand(N).astype(np.float32)
        storage.compute_returns(torch.from_numpy(nxt[lo:lo + n_local]).view(-1, 1).to(dev), True, 0.995, 0.95)
        plr.update_with_rollouts(storage)
        sampler.after_update()
        plr.reconcile()
        log.append(dict(replay=bool(replay), seeds=list(plr.current_level_seeds),
                        obs=storage.obs['image'].cpu().numpy().copy(), rewards=storage.rewards.cpu().numpy().copy(),
                        masks=storage.masks.cpu().numpy().copy(), level_seeds=storage.level_seeds.cpu().numpy().copy(),
                        returns=storage.returns.cpu().numpy().copy(), enc=np.stack(venv.get_encodings())))
    venv.close()
    state = dict(seeds=sampler.seeds.copy(), scores=sampler.seed_scores.copy(), stale=sampler.seed_staleness.copy(),
                 unseen=sampler.unseen_seed_weights.copy(), staging=sorted(sampler.staging_seed_set),
                 working=sorted(sampler.working_seed_set), count=sampler.running_sample_count,
                 store={int(k): bytes(v) for k, v in store.seed2level.items()},
                 parents={int(k): list(v) for k, v in store.seed2parent.items()})
    return state, log


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=64)
    ap.add_argument('--T', type=int, default=48)
    ap.add_argument('--cycles', type=int, default=5)
    ap.add_argument('--buffer', type=int, default=96)
    ap.add_argument('--seed', type=int, default=3)
    ap.add_argument('--env_name', default='MultiGrid-MiniGoalLastAdversarial-v0')
    ap.add_argument('--strategy', default='positive_value_loss')
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    state, log = run_cycles(a, rank, world, True)
    ok = True
    msgs = []
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (state, [{k: v for k, v in c.items()} for c in log]))
        if rank == 0:
            ref_state, ref_log = run_cycles(a, 0, 1, False)   # the same cycles, all envs on one GPU
            n_local = a.envs // world
            for r, (st, lg) in enumerate(gathered):
                for k in ('seeds', 'scores', 'stale', 'unseen'):
                    if not np.array_equal(st[k], ref_state[k]):
                        ok = False
                        msgs.append('rank %d: sampler.%s differs' % (r, k))
                for k in ('staging', 'working', 'count', 'store', 'parents'):
                    if st[k] != ref_state[k]:
                        ok = False
                        msgs.append('rank %d: %s differs' % (r, k))
                sl = slice(r * n_local, (r + 1) * n_local)
                for c, (got, want) in enumerate(zip(lg, ref_log)):
                    if got['replay'] != want['replay'] or got['seeds'] != want['seeds']:
                        ok = False
                        msgs.append('rank %d cycle %d: replay decision / level seeds differ' % (r, c))
                    for k in ('obs', 'rewards', 'masks', 'level_seeds', 'returns'):
                        if not np.array_equal(got[k], want[k][:, sl]):
                            ok = False
                            msgs.append('rank %d cycle %d: %s differs' % (r, c, k))
                    if not np.array_equal(got['enc'], want['enc'][sl]):
                        ok = False
                        msgs.append('rank %d cycle %d: level encodings differ' % (r, c))
            n_replay = sum(int(c['replay']) for c in ref_log)
            print('SHARDED_PLR %s world=%d envs=%d cycles=%d replay_cycles=%d working=%d store=%d %s' % (
                'MATCH' if ok else 'MISMATCH', world, a.envs, a.cycles, n_replay, len(ref_state['working']),
                len(ref_state['store']), '; '.join(msgs[:6])), flush=True)
        flag = torch.tensor([1 if ok else 0], device='cuda')
        dist.broadcast(flag, 0)
        ok = bool(flag.item())
        dist.barrier()
        dist.destroy_process_group()
    else:
        print('SINGLE ok working=%d store=%d' % (len(state['working']), len(state['store'])), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
