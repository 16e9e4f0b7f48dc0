"""world_size-2 gloo test (CPU) of the N>1 host logic: env sharding, the episode-record all-gather (canonical order), the
level-encoding all-gather with rank-consistent seeds, replicated replay draws and the done-flag gather of the per-episode
re-sampling give the single-process result on every rank.  (The record walk itself is a CUDA kernel: the NCCL twin of this
test, tests/test_gpu_multi.py, compares whole PLR cycles.)"""
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


class _StubSampler(object):
    """Stands in for the CUDA-backed LevelSampler in this CPU test: deterministic replay draws from the global np.random
    stream (what every rank shares), observation log."""

    def __init__(self):
        self.observed, self.seeds = [], np.arange(1, 200)

    def observe_external_unseen_sample(self, seeds, solvable=None):
        self.observed.append((list(seeds), None if solvable is None else list(solvable)))

    def sample_replay_levels(self, n):
        return [int(np.random.randint(1, 7)) for _ in range(n)]


def _levels(N, W=6):
    """N byte-encoded levels with a few duplicates (dedupe must give the same seeds on every rank)."""
    rs = np.random.RandomState(3)
    out = rs.randint(0, 3, size=(N, W, W, 3)).astype(np.uint8)
    out[5] = out[1]
    out[N - 1] = out[2]
    return out


def _worker(rank, world, port, A, T, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from dcd_isaac_b200.distributed import ShardedPLR, all_gather_episode_records, env_shard
    from dcd_isaac_b200.level_store import LevelStore
    full = _records(A, T, 7)
    lo, hi = env_shard(A, rank, world)
    local = full[(full['actor'] >= lo) & (full['actor'] < hi)].copy()
    local['actor'] -= lo                       # ranks see local actor indices
    rec = all_gather_episode_records(local, lo)
    # level buffer: every rank contributes its envs' levels, inserts the full env-ordered list
    levels = _levels(A)
    store = LevelStore(data_info={'numpy': True, 'dtype': np.uint8, 'shape': levels.shape[1:]})
    plr = ShardedPLR(_StubSampler(), store, rank, world, hi - lo)
    mine = plr.insert_current_levels([l for l in levels[lo:hi]], solvable_local=[bool(i % 2) for i in range(lo, hi)])
    strs = plr.gather_levels(['%d %d' % (i, i * i) for i in range(lo, hi)])   # action-string levels
    np.random.seed(5)
    rep_seeds, rep_levels = plr.sample_replay_levels()
    done_local = [(i % 3 == 0) for i in range(lo, hi)]
    res = plr.resample_finished(done_local)
    with open(os.path.join(out_dir, 'rank%d.pkl' % rank), 'wb') as f:
        pickle.dump(dict(rec=rec, mine=mine, all_seeds=list(plr.current_level_seeds), store=dict(store.seed2level), strs=strs,
                         observed=plr.sampler.observed, rep_seeds=rep_seeds, rep_levels=[bytes(l) for l in rep_levels],
                         res={k: (v[0], bytes(v[1])) for k, v in res.items()}), f)
    dist.destroy_process_group()


def test_sharded_gathers_match_single_process(tmp_path):
    import torch.multiprocessing as mp
    from dcd_isaac_b200.level_store import LevelStore
    A, T, world = 8, 64, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, A, T, str(tmp_path)), nprocs=world, join=True)
    full = _records(A, T, 7)
    levels = _levels(A)
    store = LevelStore(data_info={'numpy': True, 'dtype': np.uint8, 'shape': levels.shape[1:]})
    seeds = store.insert([l.tobytes() for l in levels])       # the single-process insertion (adversarial_runner.py:402-406)
    assert len(set(seeds)) == A - 2                            # two duplicates collapsed
    np.random.seed(5)
    rep = [int(np.random.randint(1, 7)) for _ in range(A)]
    done_all = [(i % 3 == 0) for i in range(A)]
    redraw = [int(np.random.randint(1, 7)) for _ in range(sum(done_all))]
    per = A // world
    for r in range(world):
        d = pickle.load(open(os.path.join(str(tmp_path), 'rank%d.pkl' % r), 'rb'))
        assert np.array_equal(d['rec'], full)                 # canonical (actor-major) order restored
        assert d['all_seeds'][:A] != [] and d['mine'] == seeds[r * per:(r + 1) * per]
        assert d['store'] == dict(store.seed2level)            # same seeds -> same levels on every rank
        assert d['strs'] == ['%d %d' % (i, i * i) for i in range(A)]
        assert d['observed'] == [(seeds, [bool(i % 2) for i in range(A)])]
        assert d['rep_seeds'] == rep[r * per:(r + 1) * per]
        assert d['rep_levels'] == [store.seed2level[s] for s in rep[r * per:(r + 1) * per]]
        want, k = {}, 0
        for i in range(A):
            if done_all[i]:
                if r * per <= i < (r + 1) * per:
                    want[i - r * per] = (redraw[k], store.seed2level[redraw[k]])
                k += 1
        assert d['res'] == want                                # draws made in ascending GLOBAL env order


def test_env_shard():
    from dcd_isaac_b200.distributed import env_shard
    assert [env_shard(1048576, r, 8) for r in (0, 7)] == [(0, 131072), (917504, 1048576)]
    with pytest.raises(AssertionError):
        env_shard(10, 0, 4)
