"""GPU parity of the PLR path: GAE, episode scores, sample weights / replay draws and whole LevelSampler +
LevelStore sessions, against fixtures produced by EXECUTING the reference and against the numpy oracle.
Tolerances: GAE bit-exact; scores / weights 1e-5 relative (north_star); sampled indices, seeds, buffer
membership and staleness exact."""
import glob
import gzip
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN, golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

RTOL = 1e-5


def test_gae_golden_bit_exact():
    from dcd_isaac_b200.storage import gae_returns
    g = golden('plr_gae.npz')
    for tag in 'abc':
        r = torch.from_numpy(g['rewards_' + tag]).cuda().unsqueeze(-1).contiguous()
        v = torch.from_numpy(g['values_' + tag]).cuda().unsqueeze(-1).contiguous()
        m = torch.from_numpy(g['masks_' + tag]).cuda().unsqueeze(-1).contiguous()
        out = torch.zeros_like(v)
        gae_returns(r, v, m, out, float(g['gamma']), float(g['gae_lambda']))
        assert np.array_equal(out.cpu().numpy()[:-1, :, 0], g['returns_' + tag][:-1]), tag


def test_gae_vs_oracle_large():
    from dcd_isaac_b200.storage import gae_returns
    from oracle import plr_oracle as po
    rs = np.random.RandomState(0)
    T, N = 256, 4096
    r = (rs.rand(T, N) < 0.02).astype(np.float32) * rs.rand(T, N).astype(np.float32)
    v = rs.randn(T + 1, N).astype(np.float32)
    m = (rs.rand(T + 1, N) > 0.03).astype(np.float32)
    out = torch.zeros(T + 1, N, 1, device='cuda')
    gae_returns(torch.from_numpy(r).cuda().unsqueeze(-1).contiguous(), torch.from_numpy(v).cuda().unsqueeze(-1).contiguous(),
                torch.from_numpy(m).cuda().unsqueeze(-1).contiguous(), out, 0.995, 0.95)
    assert np.array_equal(out.cpu().numpy()[:-1, :, 0], po.gae(r, v, m, 0.995, 0.95))


class _Rollouts(object):
    use_popart = False


def _rollouts_from(rec):
    ro = _Rollouts()
    for k in ('rewards', 'value_preds', 'masks', 'cliffhanger_masks', 'returns'):
        setattr(ro, k, torch.from_numpy(rec[k]).unsqueeze(-1).contiguous())
    ro.level_seeds = torch.from_numpy(rec['level_seeds']).unsqueeze(-1).contiguous()
    if 'action_log_dist' in rec:
        ro.action_log_dist = torch.from_numpy(rec['action_log_dist']).contiguous()
    return ro


@pytest.mark.parametrize('strategy', ['positive_value_loss', 'signed_value_loss', 'value_l1', 'least_confidence', 'min_margin', 'one_step_td_error'])
def test_episode_scores_vs_oracle(strategy):
    from dcd_isaac_b200.level_sampler import LevelSampler
    from oracle import plr_oracle as po
    rs = np.random.RandomState(4)
    T, N = 256, 512
    rec = dict(rewards=rs.rand(T, N).astype(np.float32) * (rs.rand(T, N) < 0.05),
               value_preds=rs.randn(T + 1, N).astype(np.float32), returns=rs.randn(T + 1, N).astype(np.float32),
               level_seeds=rs.randint(1, 100, size=(T, N)).astype(np.int32))
    masks = (rs.rand(T + 1, N) > 0.04).astype(np.float32)
    masks[-1] = 0
    masks[0, ::7] = 0  # a done at t == 0 must be skipped without moving start_t (level_sampler.py:504-505)
    rec['masks'] = masks
    rec['cliffhanger_masks'] = (rs.rand(T + 1, N) > 0.1).astype(np.float32)
    rec['action_log_dist'] = (rs.randn(T, N, 7) * 2).astype(np.float32)
    s = LevelSampler([], None, None, num_actors=N, strategy=strategy, sample_full_distribution=True, seed_buffer_size=64, gamma=0.995)
    got = s.episode_records(_rollouts_from(rec))
    want = po.episode_scores(rec['masks'], rec['cliffhanger_masks'], rec['returns'], rec['value_preds'], rec['rewards'],
                             rec['level_seeds'], strategy, logits=rec['action_log_dist'], gamma=0.995)
    assert len(got) == len(want)
    for k in ('actor', 't_start', 't_end', 'seed', 'cliffhanger'):
        assert np.array_equal(got[k], np.array([w[k] for w in want])), k
    assert np.allclose(got['mean_score'], [w['mean'] for w in want], rtol=RTOL, atol=1e-7)
    assert np.allclose(got['max_score'], [w['max'] for w in want], rtol=RTOL, atol=1e-6)
    assert np.allclose(got['reward_sum'], [w['reward_sum'] for w in want], rtol=RTOL, atol=1e-7)
    assert np.allclose(got['value_min'], [w['value_min'] for w in want], rtol=0, atol=0)


@pytest.mark.parametrize('knob', ['0', '1'])
def test_episode_scores_both_kernels(knob):
    """The score pass has two kernels (one thread per actor; four time segments per actor, chosen by batch size; the knob
    MGPLR_SCORE_SPLIT is read once per process): run the oracle comparison and a recorded sampler session with each forced."""
    import subprocess
    import sys
    if os.environ.get('MGPLR_SCORE_SPLIT_CHILD'):
        pytest.skip('child run')
    env = dict(os.environ, MGPLR_SCORE_SPLIT=knob, MGPLR_SCORE_SPLIT_CHILD='1')
    out = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-x', '-q', '-m', 'gpu', '-k',
                          'episode_scores_vs_oracle or partial_tails'], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]


def test_episode_scores_partial_tails():
    """Not-done tails (masks[T] != 0) and episodes spanning the time segments: the records of the kernel equal a plain
    sequential walk over each actor (t_start / t_end / seed exact, sums to 1e-5)."""
    from dcd_isaac_b200.level_sampler import LevelSampler
    rs = np.random.RandomState(11)
    for T, N, p_done in ((256, 300, 0.01), (64, 130, 0.2), (17, 70, 0.3), (16, 33, 0.0)):
        rec = dict(rewards=rs.rand(T, N).astype(np.float32), value_preds=rs.randn(T + 1, N).astype(np.float32),
                   returns=rs.randn(T + 1, N).astype(np.float32), level_seeds=rs.randint(1, 100, size=(T, N)).astype(np.int32))
        masks = (rs.rand(T + 1, N) > p_done).astype(np.float32)
        masks[-1, ::3] = 0
        rec['masks'] = masks
        rec['cliffhanger_masks'] = np.ones_like(masks)
        s = LevelSampler([], None, None, num_actors=N, strategy='positive_value_loss', sample_full_distribution=True, seed_buffer_size=64)
        got = s.episode_records(_rollouts_from(rec))
        if T == 17:   # the grow-only record buffer started smaller than this rollout's record count and was regrown
            assert len(got) > 4 * N + 64
        want = []
        for e in range(N):
            start = 0
            for t in range(T):
                if masks[t + 1, e] == 0 or t == T - 1:
                    adv = np.maximum(rec['returns'][start:t + 1, e] - rec['value_preds'][start:t + 1, e], 0)
                    want.append((e, start, t + 1, rec['level_seeds'][start, e], adv.astype(np.float64).mean(), adv.max(),
                                 np.cumsum(rec['rewards'][start:t + 1, e], dtype=np.float32)[-1], rec['value_preds'][start:t + 1, e].min(),
                                 0 if masks[t + 1, e] == 0 else 2))
                    start = t + 1
        assert len(got) == len(want), (T, N)
        w = list(zip(*want))
        for k, col in zip(('actor', 't_start', 't_end', 'seed'), w[:4]):
            assert np.array_equal(got[k], np.array(col)), (k, T, N)
        assert np.allclose(got['mean_score'], w[4], rtol=RTOL, atol=1e-7) and np.allclose(got['max_score'], w[5], rtol=0, atol=0)
        assert np.allclose(got['reward_sum'], w[6], rtol=2e-5, atol=1e-6) and np.array_equal(got['value_min'], np.array(w[7], np.float32))
        assert np.array_equal(got['cliffhanger'], np.array(w[8]))


def _rec_array(recs):
    from dcd_isaac_b200._lib import EPISODE_DTYPE
    a = np.zeros(len(recs), dtype=np.dtype(EPISODE_DTYPE))
    for i, r in enumerate(recs):
        a[i] = (r['actor'], r['t_start'], r['t_end'], r['seed'], r['mean'], r['max'], r['reward_sum'], r['value_sum'], r['value_min'],
                r['cliffhanger'])
    return a


@pytest.mark.parametrize('strategy,alpha,coef,priority,transform', [
    ('positive_value_loss', 1.0, 0.0, 'replay_support', 'rank'), ('positive_value_loss', 0.7, 0.3, 'replay_support', 'rank'),
    ('grounded_signed_value_loss', 1.0, 0.0, 'replay_support', 'rank'), ('value_l1', 0.5, 0.0, 'score', 'rank'),
    ('value_l1', 1.0, 0.0, 'replay_support', 'power'), ('uniform', 1.0, 0.0, 'replay_support', 'rank'),
    ('positive_value_loss', 1.0, 0.0, 'replay_support', 'softmax'), ('positive_value_loss', 0.7, 0.0, 'replay_support', 'match_rank'),
    ('value_l1', 1.0, 0.0, 'replay_support', 'match'), ('positive_value_loss', 1.0, 0.0, 'replay_support', 'eps_greedy')])
def test_device_record_walk_vs_oracle(strategy, alpha, coef, priority, transform):
    """mgplr_plr_apply_records (the single-CTA walk over the episode records: EWA updates, staging -> working admission with
    eviction by argmin of sample_weights, incremental ranks) against the sequential oracle that is pinned by the recorded
    reference sessions (oracle.plr_oracle.BufferOracle): identical buffer after every cycle, bit for bit, incl. seed2index with
    its stale entries, alpha < 1, max_score_coef > 0, the MaxMC grounded values, ties (scores rounded to 2 digits) and long
    runs of score changes between admissions (rank refresh by full sort)."""
    from dcd_isaac_b200.level_sampler import LevelSampler
    from oracle import plr_oracle as po
    rs = np.random.RandomState(0)
    A, T, NB = 96, 64, 48
    s = LevelSampler([], None, None, num_actors=A, strategy=strategy, max_score_coef=coef, alpha=alpha, score_transform=transform,
                     temperature=0.3, rho=0.5, staleness_coef=0.3, sample_full_distribution=True, seed_buffer_size=NB,
                     seed_buffer_priority=priority)
    o = po.BufferOracle(NB, strategy=strategy, alpha=alpha, max_score_coef=coef, priority=priority, score_transform=transform,
                        temperature=0.3, staleness_coef=0.3)
    next_seed = 1
    for cyc in range(10):
        new = list(range(next_seed, next_seed + 30))
        next_seed += 30
        if cyc % 2 == 0:
            pool, obs = new, new
        else:
            pool, obs = [int(x) for x in o.seeds if x >= 0] + new[:5], new[:5]
        if cyc == 7:   # evicted seeds come back: stale seed2index entries are exercised
            pool = pool + [1, 2, 3, 31, 32]
        s.observe_external_unseen_sample(obs)
        o.observe(obs)
        recs = []
        for a in range(A):
            t = 0
            while t < T:
                t2 = min(T, t + int(rs.randint(1, 25)))
                recs.append(dict(actor=a, t_start=t, t_end=t2, seed=int(rs.choice(pool)), mean=np.float32(round(rs.rand(), 2)),
                                 max=np.float32(rs.rand() + 1), reward_sum=np.float32(rs.rand() * (rs.rand() < 0.5)),
                                 value_sum=np.float32(rs.randn()), value_min=np.float32(rs.randn() - 1),
                                 cliffhanger=int(t2 == T and rs.rand() < 0.5)))
                t = t2
        o.apply(recs)
        s._apply_episode_records(_rec_array(recs))
        s.after_update()
        assert np.array_equal(s.seeds, o.seeds), cyc
        assert np.array_equal(s.seed_scores, o.scores), cyc
        assert np.array_equal(s.unseen_seed_weights, o.unseen) and np.array_equal(s.seed_staleness, o.stale), cyc
        assert s.staging_seed_set == set(o.stamp) and s.working_seed_set == set(int(x) for x in o.seeds if x >= 0), cyc
        assert s.seed2index == o.index_of and s.working_seed_buffer_size == o.filled, cyc
        if o.grounded is not None:
            assert np.array_equal(s.grounded_values, o.grounded), cyc
    assert o.filled == NB


def test_large_actor_count_update_is_cheap():
    """131 072 actors, a full 4 000-slot buffer, 300 000 finished episodes of working seeds: one upload + one kernel; the
    last record of a slot wins (alpha = 1)."""
    import time
    from dcd_isaac_b200._lib import EPISODE_DTYPE
    from dcd_isaac_b200.level_sampler import LevelSampler
    A, NB, n = 131072, 4000, 300000
    s = LevelSampler([], None, None, num_actors=A, strategy='positive_value_loss', sample_full_distribution=True,
                     seed_buffer_size=NB)
    s.seeds[:] = np.arange(1, NB + 1)
    s.seed2index = {int(k): i for i, k in enumerate(s.seeds)}
    s.working_seed_set = set(s.seed2index)
    s.working_seed_buffer_size = NB
    s.unseen_seed_weights[:] = 0
    rs = np.random.RandomState(1)
    rec = np.zeros(n, dtype=np.dtype(EPISODE_DTYPE))
    rec['actor'] = np.sort(rs.randint(0, A, n))
    rec['t_end'] = rs.randint(1, 250, n)
    rec['seed'] = rs.randint(1, NB + 1, n)
    rec['mean_score'] = rs.rand(n)
    rec['max_score'] = rs.rand(n)
    s._apply_episode_records(rec[:10])   # (creates the device context)
    t = time.time()
    s._apply_episode_records(rec)
    dt = time.time() - t
    assert dt < 2.0, dt
    last = {}
    for i in range(n):
        last[int(rec['seed'][i])] = i
    for k in (int(rec['seed'][n - 1]), 17, 2000):
        i = last[k]
        want = 0.0 + (float(rec['mean_score'][i]) - 0.0) * float(rec['t_end'][i]) / float(rec['t_end'][i])
        assert s.seed_scores[k - 1] == want


def test_not_done_tails_are_carried_and_flushed():
    """A rollout that does not end in `done` (never produced by the reference runner): the tail's score is carried per
    (actor, seed) with its step count, merged step-weighted into that actor's next finished episode on the seed
    (level_sampler.py:199-204), or -- after_update -- scored as it stands (:580-599)."""
    from dcd_isaac_b200._lib import EPISODE_DTYPE
    from dcd_isaac_b200.level_sampler import LevelSampler
    s = LevelSampler([11, 12], None, None, num_actors=2, strategy='value_l1', score_transform='rank', temperature=0.3)
    rec = np.zeros(3, dtype=np.dtype(EPISODE_DTYPE))
    rec[0] = (0, 0, 10, 11, 0.5, 0.9, 0, 0, 0, 0)    # actor 0 finishes an episode on seed 11
    rec[1] = (0, 10, 16, 12, 0.25, 0.4, 0, 0, 0, 2)  # ... and leaves a 6-step tail on seed 12
    rec[2] = (1, 0, 16, 11, 0.75, 0.8, 0, 0, 0, 2)   # actor 1: one 16-step tail on seed 11
    s._apply_episode_records(rec)
    assert s.seed_scores[0] == 0.5 and s.unseen_seed_weights.tolist() == [0.0, 1.0]
    assert s._tails == {(0, 12): (0.25, np.float32(0.4), 6), (1, 11): (0.75, np.float32(0.8), 16)}
    nxt = np.zeros(2, dtype=np.dtype(EPISODE_DTYPE))
    nxt[0] = (0, 0, 4, 12, 1.0, 1.0, 0, 0, 0, 0)     # the 4 remaining steps of actor 0's episode on seed 12
    nxt[1] = (1, 0, 8, 12, 0.125, 0.5, 0, 0, 0, 0)   # actor 1 plays seed 12: its tail on seed 11 stays
    s._apply_episode_records(nxt)
    assert s.seed_scores[1] == 0.125                 # last write wins (alpha = 1): actor 1's episode
    assert (0, 12) not in s._tails and (1, 11) in s._tails
    s2 = LevelSampler([11, 12], None, None, num_actors=2, strategy='value_l1', score_transform='rank', temperature=0.3)
    s2._apply_episode_records(rec)
    s2._apply_episode_records(nxt[:1])
    assert s2.seed_scores[1] == 0.25 + (1.0 - 0.25) * 4 / 10.0   # merged over 6 + 4 steps
    s2.after_update()                                # flushes actor 1's tail on seed 11 as it stands
    assert s2.seed_scores[0] == 0.75 and s2._tails == {}


def test_sample_weights_and_replay_golden():
    from dcd_isaac_b200.level_sampler import LevelSampler
    g = golden('plr_weights.npz')
    for tag in ('n4000', 'n4000_t01', 'n37_nostale', 'n100_ties'):
        temp, sc, st = g['params_' + tag]
        n = len(g['scores_' + tag])
        s = LevelSampler([], None, None, num_actors=4, strategy='positive_value_loss', score_transform='rank',
                         temperature=float(temp), staleness_coef=float(sc), staleness_transform='power',
                         staleness_temperature=float(st), sample_full_distribution=True, seed_buffer_size=n)
        s.seed_scores[:] = g['scores_' + tag]
        s.unseen_seed_weights[:] = g['unseen_' + tag]
        s.seed_staleness[:] = g['stale_' + tag]
        s.seeds[:] = np.arange(1, n + 1)
        s.working_seed_buffer_size = n
        w = s.sample_weights()
        ref = g['weights_' + tag]
        if 'ties' in tag:  # tie order is unspecified in the reference (numpy quicksort): structure only
            assert abs(w.sum() - 1) < 1e-12 and np.array_equal(w == 0, ref == 0)
            from oracle import plr_oracle as po
            want = po.sample_weights(g['scores_' + tag], g['stale_' + tag], g['unseen_' + tag], temperature=float(temp),
                                     staleness_coef=float(sc), staleness_temperature=float(st))
            assert np.allclose(w, want, rtol=1e-12)  # our documented tie rule == the oracle's stable argsort
            continue
        assert np.allclose(w, ref, rtol=1e-9, atol=0)
        np.random.seed(123)
        picks = [s.sample_replay_level() for _ in range(20)]
        picks += s.sample_replay_levels(20)
        assert np.array_equal(np.array(picks) - 1, g['picks_' + tag])
        assert np.array_equal(s.seed_staleness, g['stale_after_' + tag])


@pytest.mark.parametrize('transform', ['softmax', 'match', 'match_rank', 'eps_greedy'])
def test_extra_score_transforms_golden(transform):
    """The transforms beyond constant / rank / power against the reference's own sample_weights() and replay draws
    (tests/golden/plr_transforms.npz), through the device kernels."""
    from dcd_isaac_b200.level_sampler import LevelSampler
    g = golden('plr_transforms.npz')
    for tag in 'abc':
        k = transform + '_' + tag
        temp, sc, eps = g['params_' + k]
        n = len(g['scores_' + k])
        s = LevelSampler([], None, None, num_actors=4, strategy='positive_value_loss', score_transform=transform,
                         temperature=float(temp), eps=float(eps), staleness_coef=float(sc), staleness_transform='power',
                         staleness_temperature=1.0, sample_full_distribution=True, seed_buffer_size=n)
        s.seed_scores[:] = g['scores_' + k]
        s.unseen_seed_weights[:] = g['unseen_' + k]
        s.seed_staleness[:] = g['stale_' + k]
        s.seeds[:] = np.arange(1, n + 1)
        s.working_seed_buffer_size = n
        assert np.allclose(s.sample_weights(), g['weights_' + k], rtol=1e-9, atol=1e-300), k
        np.random.seed(321)
        picks = [s.sample_replay_level() for _ in range(10)] + s.sample_replay_levels(20)
        assert np.array_equal(np.array(picks) - 1, g['picks_' + k]), k
        assert np.array_equal(s.seed_staleness, g['stale_after_' + k]), k


@pytest.mark.parametrize('n,coef,n_unseen,n_draws', [(4000, 0.3, 800, 600), (37, 0.3, 5, 300), (4096, 0.1, 0, 4096), (500, 0.0, 100, 200),
                                                    (64, 0.5, 62, 50), (64, 0.5, 63, 20), (300, 0.3, 30, 100)])
def test_fast_sequential_draws_equal_the_general_path(monkeypatch, n, coef, n_unseen, n_draws):
    """mgplr_plr_sample_replay's Fenwick-tree path (closed-form staleness between the draws of one call) against
    the general block-wide path: same picks, same staleness afterwards -- long calls, repeated picks of the same slot (few seen
    slots), unseen slots, no staleness mix, a single seen slot and all-zero staleness (preconditions fail -> general path)."""
    from dcd_isaac_b200.level_sampler import LevelSampler
    rs = np.random.RandomState(n + n_draws)
    scores = rs.rand(n)
    unseen = np.zeros(n)
    unseen[rs.permutation(n)[:n_unseen]] = 1.0
    stale = np.floor(rs.rand(n) * 40) if n != 300 else np.zeros(n)
    res = []
    for knob in ('0', '2'):
        monkeypatch.setenv('MGPLR_REPLAY_FAST', knob)
        s = LevelSampler([], None, None, num_actors=4, strategy='positive_value_loss', score_transform='rank', temperature=0.3,
                         staleness_coef=coef, staleness_transform='power', staleness_temperature=1.0,
                         sample_full_distribution=True, seed_buffer_size=n)
        s.seed_scores[:] = scores
        s.unseen_seed_weights[:] = unseen
        s.seed_staleness[:] = stale
        s.seeds[:] = np.arange(1, n + 1)
        s.working_seed_buffer_size = n
        np.random.seed(9)
        res.append((np.array(s.sample_replay_levels(n_draws)), s.seed_staleness.copy()))
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])
    assert len(np.unique(res[0][0])) < n_draws or n >= 4000   # (small buffers: slots are picked repeatedly)


def _assert_weights_match(w, ref, s, cyc):
    """Weights equal the reference's up to the order INSIDE groups of exactly equal scores: the reference ranks with
    numpy's default unstable argsort (level_sampler.py:766), so which member of a tie group gets which rank is
    unspecified there; this build breaks ties by index.  The staleness part is deterministic, so the
    rank part is recovered and compared as a sorted multiset per tie group."""
    seen = s.unseen_seed_weights < 1
    assert np.array_equal(w == 0, ref == 0) and abs(w.sum() - 1) < 1e-9
    c = s.staleness_coef
    if c > 0:
        st = np.clip(s.seed_staleness, 0, None) ** (1. / s.staleness_temperature) * seen
        st = st / st.sum() if st.sum() > 0 else seen / len(seen)
    else:
        st = np.zeros_like(w)
    rp, rp_ref = (w - c * st) / (1 - c), (ref - c * st) / (1 - c)
    for v in np.unique(s.seed_scores[seen]):
        idx = seen & (s.seed_scores == v)
        assert np.allclose(np.sort(rp[idx]), np.sort(rp_ref[idx]), rtol=RTOL, atol=1e-9), (cyc, v)


SESSIONS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, 'plr_session_*.pkl.gz')))


@pytest.mark.parametrize('name', SESSIONS)
def test_sampler_session_golden(name):
    """Replays a whole PLR session (decisions, level insertion with dedupe, replay draws incl. the per-done
    re-sampling inside the rollout, scoring, admission / eviction, reconciliation) recorded from the reference."""
    from dcd_isaac_b200.level_sampler import LevelSampler
    from dcd_isaac_b200.level_store import LevelStore
    with gzip.open(os.path.join(GOLDEN, name), 'rb') as f:
        sess = pickle.load(f)
    tag, strategy, buf, temp, sc, rp, rho, A, T = sess['case']
    np.random.seed(77)
    rs = np.random.RandomState(5)
    s = LevelSampler([], None, None, num_actors=A, strategy=strategy, replay_schedule=sess['schedule'],
                     score_transform=sess['transform'], temperature=temp, eps=0.05, rho=rho, replay_prob=rp, alpha=1.0,
                     staleness_coef=sc, staleness_transform='power', staleness_temperature=1.0,
                     sample_full_distribution=True, seed_buffer_size=buf, seed_buffer_priority='replay_support', gamma=0.995)
    store = LevelStore(data_info={'numpy': True, 'dtype': np.uint8, 'shape': (15, 15, 3)})
    next_level = 0
    for cyc, rec in enumerate(sess['log']):
        replay = s.sample_replay_decision()
        assert replay == rec['replay'], cyc
        assert np.random.get_state()[2] == rec['rng_after_decision']
        if replay:
            seeds = [s.sample_replay_level() for _ in range(A)]
            assert np.array_equal(seeds, rec['sampled']), cyc
            assert np.array_equal(s.seed_staleness, rec['staleness_after_sample'])
            levels = store.get_levels_device(seeds)
            assert levels.shape == (A, 15, 15, 3)
            assert np.array_equal(levels[0].cpu().numpy(), store.get_level(seeds[0]))
        else:
            levels = []
            for _ in range(A):  # same level generator as oracle/gen_golden_plr.py
                enc = np.zeros((15, 15, 3), np.uint8)
                enc[:, :, 0] = 1
                enc[0, 0, 0] = 2
                ident = next_level if rs.rand() > 0.15 or next_level == 0 else rs.randint(0, next_level)
                enc[1 + ident % 13, 1 + (ident // 13) % 13, :] = (8, 1, 0)
                enc[1 + (ident // 169) % 13, 13, :] = (2, 5, 0)
                next_level += 1
                levels.append(enc.tobytes())
            seeds = store.insert(levels)
            assert np.array_equal(seeds, rec['inserted']), cyc
            s.observe_external_unseen_sample(seeds, solvable=[bool(x % 3) for x in seeds])
        # consume the generator's RNG exactly as gen_golden_plr._storage does
        done = rs.rand(T, A) < (1.0 / 10)
        done[-1] = True
        rs.rand(T, A); rs.randint(1, 250, size=(T, A)); rs.rand(A); rs.randn(T + 1, A, 1); rs.randn(T, A, 7)
        if replay:
            res = []
            for t in range(T):
                for i in range(A):
                    if done[t, i]:
                        res.append(s.sample_replay_level())
            assert np.array_equal(res, rec['resampled']), cyc
        rs.randn(A, 1)
        s.update_with_rollouts(_rollouts_from(rec))
        s.after_update()
        store.reconcile_seeds(set(int(x) for x in s.seeds if x >= 0))
        assert np.array_equal(s.seeds, rec['seeds']), cyc
        assert np.array_equal(s.unseen_seed_weights, rec['unseen'])
        assert np.allclose(s.seed_scores, rec['seed_scores'], rtol=RTOL, atol=1e-7), cyc
        assert np.array_equal(s.seed_staleness, rec['seed_staleness'])
        assert s.working_seed_buffer_size == rec['working_size']
        assert sorted(s.staging_seed_set) == list(rec['staging'])
        assert sorted(store.seed2level) == list(rec['store_seeds'])
        assert s.running_sample_count == rec['running_sample_count']
        if (s.unseen_seed_weights < 1).any():
            _assert_weights_match(s.sample_weights(), rec['weights'], s, cyc)
            assert np.isclose(float(s.solvable_mass), rec['solvable_mass'], rtol=1e-3)
        if 'grounded_values' in rec:
            assert np.allclose(s.grounded_values, rec['grounded_values'], rtol=RTOL)
        assert np.random.get_state()[2] == rec['rng_pos'], cyc
    # the objects are checkpointed whole (adversarial_runner.py:214-215)
    s2, store2 = pickle.loads(pickle.dumps((s, store)))
    assert np.array_equal(s2.seed_scores, s.seed_scores) and sorted(store2.seed2level) == sorted(store.seed2level)
    if (s2.unseen_seed_weights < 1).any():
        assert np.allclose(s2.sample_weights(), s.sample_weights())
