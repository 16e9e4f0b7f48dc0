"""Host bookkeeping of the LevelSampler drop-in (CPU only; no kernel is involved): the vectorised application of
episode records is bit-identical to the sequential walk the reference performs (level_sampler.py:497-549), including
staging admissions with eviction, alpha < 1, max_score_coef > 0 and the grounded (MaxMC) strategy."""
import copy

import numpy as np
import pytest

from dcd_isaac_b200._lib import EPISODE_DTYPE
from dcd_isaac_b200.level_sampler import LevelSampler


def _records(rs, A, T, seed_pool, p_cliff=0.1):
    rows = []
    for a in range(A):
        t = 0
        while t < T:
            t2 = min(T, t + int(rs.randint(1, 25)))
            rows.append((a, t, t2, int(rs.choice(seed_pool)), rs.rand(), rs.rand() + 1, rs.rand() * (rs.rand() < 0.5), rs.randn(),
                         rs.randn() - 1, int(t2 == T and rs.rand() < p_cliff * 5)))
            t = t2
    return np.array(rows, dtype=np.dtype(EPISODE_DTYPE))


@pytest.mark.parametrize('strategy,alpha,coef', [('positive_value_loss', 1.0, 0.0), ('positive_value_loss', 0.7, 0.3),
                                                 ('grounded_signed_value_loss', 1.0, 0.0), ('value_l1', 0.5, 0.0)])
def test_vectorised_apply_equals_sequential(strategy, alpha, coef):
    rs = np.random.RandomState(0)
    A, T, NB = 96, 64, 48
    mk = lambda: LevelSampler([], None, None, num_actors=A, strategy=strategy, max_score_coef=coef, alpha=alpha,
                              score_transform='rank', temperature=0.3, rho=0.5, staleness_coef=0.3,
                              sample_full_distribution=True, seed_buffer_size=NB, seed_buffer_priority='score')
    seq, vec = mk(), mk()
    next_seed = 1
    for cyc in range(8):
        new = list(range(next_seed, next_seed + 30))
        next_seed += 30
        pool = new if cyc % 2 == 0 else [int(x) for x in seq.seeds if x >= 0] + new[:5]
        for s in (seq, vec):
            s.observe_external_unseen_sample(new if cyc % 2 == 0 else new[:5])
        rec = _records(rs, A, T, pool)
        seq._apply_episode_records(rec, vectorize=False)
        vec._apply_episode_records(rec, vectorize=True)
        for s in (seq, vec):
            s.after_update()
        assert np.array_equal(seq.seeds, vec.seeds), cyc
        assert np.array_equal(seq.seed_scores, vec.seed_scores), cyc
        assert np.array_equal(seq.unseen_seed_weights, vec.unseen_seed_weights)
        assert np.array_equal(seq.seed_staleness, vec.seed_staleness)
        assert seq.staging_seed_set == vec.staging_seed_set and seq.working_seed_set == vec.working_seed_set
        assert seq.seed2index == vec.seed2index
        if seq.grounded_values is not None:
            assert np.array_equal(seq.grounded_values, vec.grounded_values)
    assert seq.working_seed_buffer_size == NB  # the buffer filled and evicted
    assert vec._partials is None               # the dense [actors, slots] arrays were never needed


def test_large_actor_count_is_cheap():
    """131 072 actors x 4000 slots: no dense partial arrays, a replay-rollout update is a vectorised pass."""
    A, NB = 131072, 4000
    s = LevelSampler([], None, None, num_actors=A, strategy='positive_value_loss', sample_full_distribution=True,
                     seed_buffer_size=NB)
    s.seeds[:] = np.arange(1, NB + 1)
    s.seed2index = {int(k): i for i, k in enumerate(s.seeds)}
    s.working_seed_set = set(s.seed2index)
    s.working_seed_buffer_size = NB
    s.unseen_seed_weights[:] = 0
    rs = np.random.RandomState(1)
    n = 300000
    rec = np.zeros(n, dtype=np.dtype(EPISODE_DTYPE))
    rec['actor'] = np.sort(rs.randint(0, A, n))
    rec['t_end'] = rs.randint(1, 250, n)
    rec['seed'] = rs.randint(1, NB + 1, n)
    rec['mean_score'] = rs.rand(n)
    rec['max_score'] = rs.rand(n)
    import time
    t = time.time()
    s._apply_episode_records(rec)
    s.after_update()
    assert time.time() - t < 5.0
    last = {}
    for i in range(n):
        last[int(rec['seed'][i])] = i
    k = int(rec['seed'][n - 1])
    i = last[k]
    want = 0.0 + (float(rec['mean_score'][i]) - 0.0) * float(rec['t_end'][i]) / float(rec['t_end'][i])
    assert s.seed_scores[k - 1] == want and s._partials is None


def test_pickle_round_trip():
    import pickle
    s = LevelSampler([], None, None, num_actors=4, strategy='positive_value_loss', sample_full_distribution=True,
                     seed_buffer_size=8)
    s.observe_external_unseen_sample([1, 2, 3])
    s2 = pickle.loads(pickle.dumps(s))
    assert s2.staging_seed_set == {1, 2, 3} and s2.running_sample_count == 3
