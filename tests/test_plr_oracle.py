"""The numpy PLR oracle against fixtures produced by executing the reference (CPU only)."""
import numpy as np

from conftest import golden
from oracle import plr_oracle as po


def test_gae_bit_exact():
    g = golden('plr_gae.npz')
    for tag in 'abc':
        got = po.gae(g['rewards_' + tag], g['values_' + tag], g['masks_' + tag], float(g['gamma']), float(g['gae_lambda']))
        assert np.array_equal(got, g['returns_' + tag][:-1]), tag


def test_sample_weights_and_draws():
    g = golden('plr_weights.npz')
    for tag in ('n4000', 'n4000_t01', 'n100_ties', 'n37_nostale'):
        temp, sc, st = g['params_' + tag]
        kw = dict(score_transform='rank', temperature=temp, staleness_coef=sc, staleness_temperature=st)
        w = po.sample_weights(g['scores_' + tag], g['stale_' + tag], g['unseen_' + tag], **kw)
        ref = g['weights_' + tag]
        if 'ties' in tag:
            # The reference ranks with numpy's default (unstable) argsort, so WHICH member of a tie group gets which
            # rank -- and, through the seen-mask, the normaliser -- is unspecified there.  Only structure is comparable.
            assert abs(np.sum(w) - 1.0) < 1e-12 and abs(np.sum(ref) - 1.0) < 1e-12
            assert np.array_equal(w == 0, ref == 0) and (w >= 0).all()
            continue
        assert np.allclose(w, ref, rtol=1e-12, atol=0)
        idx, stale = po.sample_replay(g['scores_' + tag], g['stale_' + tag], g['unseen_' + tag], g['u_' + tag], **kw)
        assert np.array_equal(idx, g['picks_' + tag])
        assert np.array_equal(stale, g['stale_after_' + tag])


def test_extra_score_transforms():
    """softmax / match / match_rank / eps_greedy (level_sampler.py:752-785) against the reference's sample_weights and 30 replay
    draws (oracle/gen_golden_plr.py::gen_transforms); the uniforms are re-drawn from the same np.random seed."""
    g = golden('plr_transforms.npz')
    for transform in ('softmax', 'match', 'match_rank', 'eps_greedy'):
        for tag in 'abc':
            k = transform + '_' + tag
            temp, sc, eps = g['params_' + k]
            kw = dict(score_transform=transform, temperature=temp, staleness_coef=sc, staleness_temperature=1.0, sampler_eps=eps)
            w = po.sample_weights(g['scores_' + k], g['stale_' + k], g['unseen_' + k], **kw)
            assert np.allclose(w, g['weights_' + k], rtol=1e-12, atol=0), k
            np.random.seed(321)
            u = np.array([np.random.random_sample() for _ in range(30)])
            idx, stale = po.sample_replay(g['scores_' + k], g['stale_' + k], g['unseen_' + k], u, **kw)
            assert np.array_equal(idx, g['picks_' + k]), k
            assert np.array_equal(stale, g['stale_after_' + k]), k


def test_closed_form_sequential_draws():
    """The algebra of the Fenwick-tree draw path (staleness in closed form between the draws of one call):
    same picks and final staleness as recomputing sample_weights() before every draw -- on the recorded reference
    fixtures and on random cases with repeated picks, unseen slots and no staleness mix."""
    g = golden('plr_weights.npz')
    for tag in ('n4000', 'n4000_t01', 'n37_nostale'):
        temp, sc, st = g['params_' + tag]
        kw = dict(score_transform='rank', temperature=temp, staleness_coef=sc, staleness_temperature=st)
        idx, stale = po.sample_replay_closed_form(g['scores_' + tag], g['stale_' + tag], g['unseen_' + tag], g['u_' + tag], **kw)
        assert np.array_equal(idx, g['picks_' + tag]) and np.array_equal(stale, g['stale_after_' + tag]), tag
    rs = np.random.RandomState(0)
    for n, coef, n_unseen, n_draws, tr in [(4000, 0.3, 800, 400, 'rank'), (37, 0.3, 5, 300, 'rank'), (500, 0.0, 100, 200, 'rank'),
                                           (64, 0.5, 60, 80, 'power'), (300, 0.1, 0, 300, 'softmax')]:
        scores, unseen = rs.rand(n), np.zeros(n)
        unseen[rs.permutation(n)[:n_unseen]] = 1.0
        stale, u = np.floor(rs.rand(n) * 40), rs.rand(n_draws)
        kw = dict(score_transform=tr, temperature=0.3, staleness_coef=coef, staleness_temperature=1.0)
        i1, s1 = po.sample_replay(scores, stale, unseen, u, **kw)
        i2, s2 = po.sample_replay_closed_form(scores, stale, unseen, u, **kw)
        assert np.array_equal(i1, i2) and np.array_equal(s1, s2), (n, coef, tr)
        assert n >= 4000 or len(np.unique(i1)) < n_draws   # small buffers: slots are picked repeatedly


def test_storage_returns_and_value_loss():
    """discounted returns bit-exact, batched value loss to 1e-6 against the executed reference RolloutStorage."""
    g = golden('plr_storage.npz')
    for tag in 'ab':
        ret = po.discounted_returns(g['rewards_' + tag], g['masks_' + tag], g['values_' + tag][-1], 0.995)
        assert np.array_equal(ret, g['disc_returns_' + tag]), tag
        k = 0
        while 'bvl_%s_%d' % (tag, k) in g.files:
            signed, pos, power, clipped = [int(x) for x in g['bvl_params_%s_%d' % (tag, k)]]
            got = po.batched_value_loss(g['gae_returns_' + tag], g['values_' + tag], bool(signed), bool(pos), power, bool(clipped))
            assert np.allclose(got, g['bvl_%s_%d' % (tag, k)], rtol=1e-5, atol=1e-7), (tag, k)
            k += 1
        assert k == 12


def test_logit_and_td_score_functions():
    """least_confidence / min_margin / one_step_td_error of the oracle against the reference's own score functions
    (called on single episodes of length 1 .. 250 by oracle/gen_golden_plr.py::gen_score_functions)."""
    g = golden('plr_score_functions.npz')
    for k in range(int(g['n'])):
        logits, rewards, values = g['logits_%d' % k], g['rewards_%d' % k], g['values_%d' % k]
        L = len(rewards)
        masks = np.ones((L + 1, 1), np.float32)
        masks[L] = 0
        for tag, strategy in (('lc', 'least_confidence'), ('mm', 'min_margin'), ('td', 'one_step_td_error')):
            rec = po.episode_scores(masks, np.ones_like(masks), None, np.append(values, 0).reshape(L + 1, 1), rewards.reshape(L, 1),
                                    np.ones((L, 1), np.int32), strategy, logits=logits.reshape(L, 1, 7), gamma=float(g['gamma']))
            assert len(rec) == 1 and rec[0]['t_end'] == L
            want = g['%s_%d' % (tag, k)]
            assert np.allclose([rec[0]['mean'], rec[0]['max']], want, rtol=1e-5, atol=1e-6), (strategy, L, rec[0], want)


def test_buffer_oracle_reproduces_reference_sessions():
    """oracle.plr_oracle.BufferOracle (the sequential record walk the CUDA bookkeeping kernel is checked against) replayed
    over the recorded sessions of the reference LevelSampler: same buffer contents, scores, staleness, staging set after
    every cycle -- including staging -> working admissions with eviction by replay support."""
    import glob
    import gzip
    import os
    import pickle
    from conftest import GOLDEN
    n_adm = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, 'plr_session_*.pkl.gz'))):
        with gzip.open(path, 'rb') as f:
            sess = pickle.load(f)
        tag, strategy, buf, temp, sc, rp, rho, A, T = sess['case']
        o = po.BufferOracle(buf, strategy=strategy, score_transform=sess['transform'], temperature=temp, staleness_coef=sc)
        for cyc, rec in enumerate(sess['log']):
            def drawn(seeds):
                if sc > 0:
                    for s in seeds:
                        o.stale = o.stale + 1
                        o.stale[o.index_of[int(s)]] = 0
            if rec['replay']:
                drawn(rec['sampled'])
                drawn(rec['resampled'])
            else:
                o.observe(rec['inserted'])
            sq = lambda x: None if x is None else np.asarray(x)[..., 0] if np.asarray(x).ndim == 3 else np.asarray(x)
            logits = rec.get('action_log_dist')
            recs = po.episode_scores(sq(rec['masks']), sq(rec['cliffhanger_masks']), sq(rec['returns']), sq(rec['value_preds']),
                                     sq(rec['rewards']), sq(rec['level_seeds']), strategy,
                                     logits=None if logits is None else np.asarray(logits), gamma=0.995)
            before = o.filled
            o.apply(recs)
            n_adm += int(o.filled > before)
            assert np.array_equal(o.seeds, rec['seeds']), (tag, cyc)
            assert np.array_equal(o.unseen, rec['unseen']), (tag, cyc)
            assert np.allclose(o.scores, rec['seed_scores'], rtol=1e-5, atol=1e-7), (tag, cyc)
            assert np.array_equal(o.stale, rec['seed_staleness']), (tag, cyc)
            assert o.filled == rec['working_size'] and sorted(o.stamp) == list(rec['staging']), (tag, cyc)
            assert o.count == rec['running_sample_count']
            if 'grounded_values' in rec:
                assert np.allclose(o.grounded, rec['grounded_values'], rtol=1e-5)
    assert n_adm > 5
