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
