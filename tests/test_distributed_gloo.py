"""world_size-2 gloo test (CPU) of the N>1 host logic: env sharding, the episode-record all-gather and the
replicated, canonically ordered sampler update give the single-process result on every rank."""
import gzip
import os
import pickle
import socket

import numpy as np
import pytest

from conftest import GOLDEN

torch = pytest.importorskip('torch')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_sampler(A):
    from dcd_isaac_b200.level_sampler import LevelSampler
    return LevelSampler([], None, None, num_actors=A, strategy='positive_value_loss', replay_schedule='proportionate',
                        score_transform='rank', temperature=0.3, rho=0.5, replay_prob=0.8, staleness_coef=0.3,
                        sample_full_distribution=True, seed_buffer_size=64)


def _records(A, T, seed):
    """Synthetic episode records in canonical order (what mgplr_plr_episode_scores emits)."""
    from dcd_isaac_b200._lib import EPISODE_DTYPE
    rs = np.random.RandomState(seed)
    rows = []
    for a in range(A):
        t = 0
        while t < T:
            t2 = min(T, t + rs.randint(1, 20))
            rows.append((a, t, t2, 1 + rs.randint(0, 40), rs.rand(), rs.rand() + 1, rs.rand(), rs.rand(), rs.rand(),
                         int(t2 == T and rs.rand() < 0.5)))
            t = t2
    return np.array(rows, dtype=np.dtype(EPISODE_DTYPE))


def _worker(rank, world, port, A, T, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from dcd_isaac_b200.distributed import all_gather_episode_records, env_shard
    s = _make_sampler(A)
    s.observe_external_unseen_sample(list(range(1, 41)))
    full = _records(A, T, 7)
    lo, hi = env_shard(A, rank, world)
    local = full[(full['actor'] >= lo) & (full['actor'] < hi)].copy()
    local['actor'] -= lo                       # ranks see local actor indices
    rec = all_gather_episode_records(local, lo)
    s._apply_episode_records(rec)
    with open(os.path.join(out_dir, 'rank%d.pkl' % rank), 'wb') as f:
        pickle.dump((rec, s.seeds, s.seed_scores, s.unseen_seed_weights, s.seed_staleness, sorted(s.staging_seed_set)), f)
    dist.destroy_process_group()


def test_sharded_update_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    A, T, world = 8, 64, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, A, T, str(tmp_path)), nprocs=world, join=True)
    s = _make_sampler(A)
    s.observe_external_unseen_sample(list(range(1, 41)))
    full = _records(A, T, 7)
    s._apply_episode_records(full)
    for r in range(world):
        rec, seeds, scores, unseen, stale, staging = pickle.load(open(os.path.join(str(tmp_path), 'rank%d.pkl' % r), 'rb'))
        assert np.array_equal(rec, full)
        assert np.array_equal(seeds, s.seeds) and np.array_equal(scores, s.seed_scores)
        assert np.array_equal(unseen, s.unseen_seed_weights) and np.array_equal(stale, s.seed_staleness)
        assert staging == sorted(s.staging_seed_set)


def test_env_shard():
    from dcd_isaac_b200.distributed import env_shard
    assert [env_shard(1048576, r, 8) for r in (0, 7)] == [(0, 131072), (917504, 1048576)]
    with pytest.raises(AssertionError):
        env_shard(10, 0, 4)
