"""GPU parity of the RolloutStorage drop-in (dcd_isaac_b200/storage.py) against fixtures produced by executing the
reference's algos/storage.py (oracle/gen_golden_plr.py::gen_storage) and against the numpy oracle at full size."""
import gzip
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN, golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_discounted_returns_value_loss_and_traj_golden():
    from dcd_isaac_b200 import storage as S
    g = golden('plr_storage.npz')
    for tag in 'ab':
        rewards, values, masks = g['rewards_' + tag], g['values_' + tag], g['masks_' + tag]
        T, N = rewards.shape
        ret = torch.zeros(T + 1, N, 1, device='cuda')
        ret[T, :, 0] = _cuda(values[-1])
        S.discounted_returns(_cuda(rewards).unsqueeze(-1), _cuda(masks).unsqueeze(-1), ret, 0.995)
        assert np.array_equal(ret.cpu().numpy()[:, :, 0], g['disc_returns_' + tag]), tag  # bit-exact
        k = 0
        while 'bvl_%s_%d' % (tag, k) in g.files:
            signed, pos, power, clipped = [int(x) for x in g['bvl_params_%s_%d' % (tag, k)]]
            got = S.batched_value_loss(_cuda(g['gae_returns_' + tag]).unsqueeze(-1), _cuda(values).unsqueeze(-1), bool(signed), bool(pos),
                                       power, bool(clipped))
            assert got.shape == (N, 1)
            assert np.allclose(got.cpu().numpy()[:, 0], g['bvl_%s_%d' % (tag, k)], rtol=1e-5, atol=1e-7), (tag, k)  # 1e-5 relative
            k += 1
        assert k == 12


class _Space(object):
    def __init__(self, shape):
        self.shape = shape


class Discrete(object):
    def __init__(self, n):
        self.n = n


class _Stub(object):
    def __init__(self, w):
        self.w = torch.from_numpy(w).cuda()

    def get_value(self, obs, rnn_hxs, masks):
        return (obs['image'].reshape(-1, 75) @ self.w).unsqueeze(-1) + 0.1 * obs['direction'].reshape(-1, 1)


def test_storage_session_golden():
    """The reference RolloutStorage driven through copy_obs_to_index / insert / insert_truncated_obs / compute_returns /
    after_update (LSTM state, proper time limits with a stub critic): every buffer of the drop-in must match."""
    from dcd_isaac_b200.storage import RolloutStorage
    with gzip.open(os.path.join(GOLDEN, 'plr_storage_session.pkl.gz'), 'rb') as f:
        g = pickle.load(f)
    ins = g['inserts']
    T, N = len(ins), ins[0]['action'].shape[0]
    st = RolloutStorage(model=_Stub(g['stub_w']), num_steps=T, num_processes=N,
                        observation_space={'image': _Space((3, 5, 5)), 'direction': _Space((1,))}, action_space=Discrete(7),
                        recurrent_hidden_state_size=4, recurrent_arch='lstm', use_proper_time_limits=True, device='cuda')
    st.copy_obs_to_index({k: _cuda(v) for k, v in g['first'].items()}, 0)
    for r in ins:
        for i, tr in r['trunc'].items():
            st.insert_truncated_obs(tr, index=i)  # numpy observations, as the runner passes them
        st.insert({k: _cuda(v) for k, v in r['obs'].items()}, tuple(_cuda(h) for h in r['hx']), _cuda(r['action']), _cuda(r['logp']),
                  _cuda(r['logd']), _cuda(r['val']), _cuda(r['rew']), _cuda(r['masks']), _cuda(r['bad']), level_seeds=_cuda(r['seeds']),
                  cliffhanger_masks=_cuda(r['cliff']))
    assert st.step == g['step']
    st.compute_returns(_cuda(g['next_value']), True, 0.995, 0.95)
    for k in ('recurrent_hidden_states', 'actions', 'action_log_probs', 'action_log_dist', 'value_preds', 'rewards', 'masks',
              'bad_masks', 'level_seeds'):
        assert np.array_equal(getattr(st, k).cpu().numpy(), g[k]), k
    for k in ('image', 'direction'):
        assert np.array_equal(st.obs[k].cpu().numpy(), g['obs'][k]), k
        assert np.array_equal(st.truncated_obs[k].cpu().numpy(), g['truncated_obs'][k]), k
    # the stub critic is a GPU matmul here and a CPU matmul in the fixture: 1e-5 relative
    assert np.allclose(st.truncated_value_preds.cpu().numpy(), g['truncated_value_preds'], rtol=1e-5, atol=1e-6)
    assert np.allclose(st.returns.cpu().numpy(), g['returns'], rtol=1e-5, atol=1e-5)
    st.after_update()
    au = g['after_update']
    for k in ('image', 'direction'):
        assert np.array_equal(st.obs[k][0].cpu().numpy(), au['obs0'][k])
    assert np.array_equal(st.masks[0].cpu().numpy(), au['masks0']) and np.array_equal(st.bad_masks[0].cpu().numpy(), au['bad0'])
    assert np.array_equal(st.recurrent_hidden_states[0].cpu().numpy(), au['rnn0'])


def test_action_traj_strings_golden():
    from dcd_isaac_b200.storage import RolloutStorage
    g = golden('plr_storage.npz')
    for tag in 'ab':
        acts = g['actions_' + tag]
        T, N = acts.shape
        st = RolloutStorage(None, T, N, {'image': _Space((3, 5, 5)), 'direction': _Space((1,))}, Discrete(7), 1, device='cuda')
        st.actions.copy_(_cuda(acts).unsqueeze(-1))
        assert st.get_action_traj(as_string=True) == [str(x) for x in g['traj_' + tag]]
        assert torch.equal(st.get_action_traj(), st.actions.squeeze(-1))


def test_returns_and_value_loss_vs_oracle_full_size():
    """131 072 actors x T=256 against the numpy oracle: discounted returns bit-exact, value loss 1e-5 relative."""
    from dcd_isaac_b200 import storage as S
    from oracle import plr_oracle as po
    rs = np.random.RandomState(3)
    T, N = 256, 131072
    rewards = (rs.rand(T, N) < 0.01).astype(np.float32) * rs.rand(T, N).astype(np.float32)
    masks = np.ones((T + 1, N), np.float32)
    masks[1:][rs.rand(T, N) < 0.02] = 0
    values = rs.randn(T + 1, N).astype(np.float32)
    want = po.discounted_returns(rewards, masks, values[-1], 0.995)
    ret = torch.zeros(T + 1, N, 1, device='cuda')
    ret[T, :, 0] = _cuda(values[-1])
    S.discounted_returns(_cuda(rewards).unsqueeze(-1), _cuda(masks).unsqueeze(-1), ret, 0.995)
    assert np.array_equal(ret.cpu().numpy()[:, :, 0], want)
    for kw in (dict(), dict(signed=True, power=2, clipped=False), dict(positive_only=True)):
        got = S.batched_value_loss(ret, _cuda(values).unsqueeze(-1), **kw).cpu().numpy()[:, 0]
        assert np.allclose(got, po.batched_value_loss(want, values, **kw), rtol=1e-5, atol=1e-7)


def test_feed_forward_and_recurrent_generators_cover_the_rollout():
    from dcd_isaac_b200.storage import RolloutStorage
    T, N = 8, 6
    st = RolloutStorage(None, T, N, {'image': _Space((3, 5, 5)), 'direction': _Space((1,))}, Discrete(7), 4, recurrent_arch='lstm',
                        device='cuda')
    st.rewards.copy_(torch.arange(T * N, device='cuda', dtype=torch.float32).view(T, N, 1))
    st.returns[:-1].copy_(st.rewards)
    adv = st.rewards.clone()
    seen = []
    for obs, rnn, act, val, ret, msk, logp, adv_t in st.feed_forward_generator(adv, num_mini_batch=4):
        assert obs['image'].shape[1:] == (3, 5, 5) and len(rnn) == 2 and torch.equal(ret, adv_t)
        seen.append(ret.flatten())
    assert sorted(torch.cat(seen).tolist()) == list(range(T * N))
    seen = []
    for obs, rnn, act, val, ret, msk, logp, adv_t in st.recurrent_generator(adv, num_mini_batch=3):
        assert obs['image'].shape == (T * 2, 3, 5, 5) and rnn[0].shape == (2, 4) and torch.equal(ret, adv_t)
        seen.append(ret.flatten())
    assert sorted(torch.cat(seen).tolist()) == list(range(T * N))
