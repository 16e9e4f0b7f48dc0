"""PLR-side golden fixtures, produced by EXECUTING the reference's level_replay/level_sampler.py,
level_replay/level_store.py and algos/storage.py unmodified (TEST INFRASTRUCTURE ONLY).

  python oracle/gen_golden.py --only plr
"""
import gzip
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def _storage(T, N, rs, mean_len=12, handle_timelimits=True, tie_free=False):
    """A reference RolloutStorage filled like adversarial_runner.agent_rollout fills it."""
    import numpy as np
    import torch
    from gym import spaces
    from algos.storage import RolloutStorage
    obs_space = {'image': spaces.Box(0, 255, (3, 5, 5), 'uint8'), 'direction': spaces.Box(0, 3, (1,), 'uint8')}
    st = RolloutStorage(model=None, num_steps=T, num_processes=N, observation_space=obs_space,
                        action_space=spaces.Discrete(7), recurrent_hidden_state_size=1,
                        use_proper_time_limits=False)
    done = rs.rand(T, N) < (1.0 / mean_len)
    done[-1] = True  # the runner forces done on the last step (adversarial_runner.py:530)
    goal = done & (rs.rand(T, N) < 0.5)
    rewards = np.where(goal, 1.0 - 0.9 * rs.randint(1, 250, size=(T, N)) / 250.0, 0.0).astype(np.float32)
    st.rewards.copy_(torch.from_numpy(rewards).unsqueeze(-1))
    st.masks[1:].copy_(torch.from_numpy(1.0 - done.astype(np.float32)).unsqueeze(-1))
    cliff = np.zeros((T, N), bool)
    if handle_timelimits:
        cliff[-1] = rs.rand(N) < 0.6  # not-done envs at the rollout end become cliffhangers (:521-528)
    st.cliffhanger_masks[1:].copy_(torch.from_numpy(1.0 - cliff.astype(np.float32)).unsqueeze(-1))
    noise = rs.randn(T + 1, N, 1).astype(np.float32)
    # tie_free: every advantage is > 0, so no positive_value_loss score is exactly 0.  A seen level whose score ties
    # with the (score 0) unseen slots gets an arbitrary rank in the reference (numpy's unstable argsort), which no
    # re-implementation can reproduce; ties are covered separately (plr_weights.npz 'n100_ties').
    st.value_preds.copy_(torch.from_numpy(noise * 0.001 - 1.0 if tie_free else noise * 0.3 + 0.4))
    st.action_log_dist.copy_(torch.from_numpy(rs.randn(T, N, 7).astype(np.float32)))
    return st, done


def gen_gae():
    import numpy as np
    import torch
    rs = np.random.RandomState(11)
    out = {}
    for tag, T, N in (('a', 256, 32), ('b', 64, 7), ('c', 5, 1)):
        st, _ = _storage(T, N, rs)
        nv = torch.from_numpy(rs.randn(N, 1).astype(np.float32))
        st.compute_returns(nv, True, 0.995, 0.95)
        out['rewards_' + tag] = st.rewards.numpy()[:, :, 0].copy()
        out['values_' + tag] = st.value_preds.numpy()[:, :, 0].copy()
        out['masks_' + tag] = st.masks.numpy()[:, :, 0].copy()
        out['returns_' + tag] = st.returns.numpy()[:, :, 0].copy()
    np.savez_compressed(os.path.join(GOLDEN, 'plr_gae.npz'), gamma=0.995, gae_lambda=0.95, **out)
    print('gae fixtures', sorted(out)[:3])


PLR_CASES = [
    # tag, strategy, buffer, temperature, staleness_coef, replay_prob, rho, num_actors, T
    ('pvl_small', 'positive_value_loss', 24, 0.3, 0.3, 0.8, 0.5, 8, 64),
    ('maxmc_small', 'grounded_signed_value_loss', 24, 0.1, 0.3, 0.5, 0.5, 8, 64),
    ('pvl_4000', 'positive_value_loss', 4000, 0.3, 0.3, 0.5, 0.01, 32, 256),
    ('l1_nostale', 'value_l1', 16, 1.0, 0.0, 0.95, 0.25, 4, 48),
    ('signed_power', 'signed_value_loss', 16, 0.5, 0.1, 0.9, 0.25, 4, 48),
    # policy-logit / TD strategies (the rollouts' action_log_dist is recorded for these)
    ('lc_small', 'least_confidence', 16, 0.3, 0.3, 0.8, 0.25, 4, 48),
    ('mm_small', 'min_margin', 16, 0.3, 0.3, 0.8, 0.25, 4, 48),
    ('td_small', 'one_step_td_error', 16, 0.3, 0.3, 0.8, 0.25, 4, 48),
]


def gen_sampler(only=None):
    """A replayed PLR session: new levels observed, rollouts scored, buffer admission/eviction, replay
    decisions and draws -- all from the reference LevelSampler/LevelStore with a seeded global np.random."""
    import numpy as np
    import torch
    from level_replay import LevelSampler, LevelStore
    from gym import spaces
    for tag, strategy, buf, temp, sc, rp, rho, A, T in PLR_CASES:
        if only is not None and tag not in only:
            continue
        np.random.seed(77)
        rs = np.random.RandomState(5)
        transform = 'power' if tag == 'signed_power' else 'rank'
        sampler = LevelSampler([], {'image': spaces.Box(0, 255, (3, 5, 5), 'uint8')}, spaces.Discrete(7), num_actors=A,
                               strategy=strategy, replay_schedule=('fixed' if tag == 'pvl_4000' else 'proportionate'), score_transform=transform,
                               temperature=temp, eps=0.05, rho=rho, replay_prob=rp, alpha=1.0, staleness_coef=sc,
                               staleness_transform='power', staleness_temperature=1.0, sample_full_distribution=True,
                               seed_buffer_size=buf, seed_buffer_priority='replay_support', use_dense_rewards=False,
                               gamma=0.995)
        store = LevelStore(data_info={'numpy': True, 'dtype': np.uint8, 'shape': (15, 15, 3)})
        log = []
        n_cycles = 14 if buf <= 24 else 8
        next_level = 0
        for cyc in range(n_cycles):
            rec = {}
            replay = bool(sampler.sample_replay_decision())
            rec['replay'] = replay
            rec['rng_after_decision'] = np.random.get_state()[2]
            if replay:
                seeds = [sampler.sample_replay_level() for _ in range(A)]
                rec['sampled'] = np.array(seeds, dtype=np.int64)
                rec['staleness_after_sample'] = sampler.seed_staleness.copy()
            else:
                levels = []
                for _ in range(A):
                    enc = np.zeros((15, 15, 3), np.uint8)
                    enc[:, :, 0] = 1
                    enc[0, 0, 0] = 2
                    # a few duplicates exercise LevelStore's content dedupe
                    ident = next_level if rs.rand() > 0.15 or next_level == 0 else rs.randint(0, next_level)
                    enc[1 + ident % 13, 1 + (ident // 13) % 13, :] = (8, 1, 0)
                    enc[1 + (ident // 169) % 13, 13, :] = (2, 5, 0)
                    next_level += 1
                    levels.append(enc.tobytes())
                seeds = store.insert(levels)
                rec['inserted'] = np.array(seeds, dtype=np.int64)
                solv = [bool(s % 3) for s in seeds]
                sampler.observe_external_unseen_sample(seeds, solvable=solv)
            st, done = _storage(T, A, rs, mean_len=10, tie_free=tag.startswith('pvl'))
            # level_seeds: current seed per actor, re-sampled on done when replaying (adversarial_runner.py:551-588)
            cur = list(seeds)
            ls = np.zeros((T, A), np.int32)
            resampled = []
            for t in range(T):
                ls[t] = cur
                if replay:
                    for i in range(A):
                        if done[t, i]:
                            cur[i] = sampler.sample_replay_level()
                            resampled.append(cur[i])
            st.level_seeds.copy_(torch.from_numpy(ls).unsqueeze(-1))
            rec['resampled'] = np.array(resampled, dtype=np.int64)
            nv = torch.from_numpy(rs.randn(A, 1).astype(np.float32))
            st.compute_returns(nv, True, 0.995, 0.95)
            for k in ('rewards', 'value_preds', 'masks', 'cliffhanger_masks', 'returns'):
                rec[k] = getattr(st, k).numpy()[:, :, 0].copy()
            rec['level_seeds'] = ls
            if strategy in ('least_confidence', 'min_margin'):
                rec['action_log_dist'] = st.action_log_dist.numpy().copy()
            if replay or True:  # robust PLR also scores the non-replay (exploratory) rollouts
                sampler.update_with_rollouts(st)
                sampler.after_update()
            store.reconcile_seeds(set(int(x) for x in sampler.seeds if x >= 0))
            rec['seeds'] = sampler.seeds.copy()
            rec['seed_scores'] = sampler.seed_scores.copy()
            rec['seed_staleness'] = sampler.seed_staleness.copy()
            rec['unseen'] = sampler.unseen_seed_weights.copy()
            rec['weights'] = sampler.sample_weights().copy() if (sampler.unseen_seed_weights < 1).any() else np.zeros(buf)
            rec['working_size'] = sampler.working_seed_buffer_size
            rec['staging'] = np.array(sorted(sampler.staging_seed_set), dtype=np.int64)
            rec['store_seeds'] = np.array(sorted(store.seed2level), dtype=np.int64)
            rec['solvable_mass'] = float(sampler.solvable_mass) if (sampler.unseen_seed_weights < 1).any() else -1.0
            rec['running_sample_count'] = sampler.running_sample_count
            if sampler.grounded_values is not None:
                rec['grounded_values'] = sampler.grounded_values.copy()
            rec['rng_pos'] = np.random.get_state()[2]
            log.append(rec)
        with gzip.open(os.path.join(GOLDEN, 'plr_session_%s.pkl.gz' % tag), 'wb') as f:
            pickle.dump({'case': (tag, strategy, buf, temp, sc, rp, rho, A, T), 'transform': transform,
                         'schedule': 'fixed' if tag == 'pvl_4000' else 'proportionate', 'log': log}, f,
                        protocol=4)
        print('plr session', tag, 'replays', sum(r['replay'] for r in log), 'working', log[-1]['working_size'],
              'max score', float(np.max(log[-1]['seed_scores'])))


def gen_weights():
    """sample_weights / sampling with recorded uniforms on tie-free and tied score vectors."""
    import numpy as np
    from level_replay import LevelSampler
    from gym import spaces
    rs = np.random.RandomState(3)
    out = {}
    for tag, n, temp, sc, ties in (('n4000', 4000, 0.3, 0.3, False), ('n4000_t01', 4000, 0.1, 0.3, False),
                                   ('n100_ties', 100, 0.3, 0.3, True), ('n37_nostale', 37, 1.0, 0.0, False)):
        s = LevelSampler([], {'image': spaces.Box(0, 255, (3, 5, 5), 'uint8')}, spaces.Discrete(7), num_actors=4,
                         strategy='positive_value_loss', score_transform='rank', temperature=temp, rho=0.5,
                         replay_prob=0.5, staleness_coef=sc, staleness_transform='power', staleness_temperature=1.0,
                         sample_full_distribution=True, seed_buffer_size=n)
        scores = rs.rand(n)
        if ties:
            scores = np.round(scores * 5) / 5
        unseen = (rs.rand(n) < 0.2).astype(np.float64)
        stale = np.floor(rs.rand(n) * 50)
        s.seed_scores[:] = scores
        s.unseen_seed_weights[:] = unseen
        s.seed_staleness[:] = stale
        s.seeds[:] = np.arange(1, n + 1)
        s.working_seed_buffer_size = n
        out['scores_' + tag] = scores
        out['unseen_' + tag] = unseen
        out['stale_' + tag] = stale
        out['weights_' + tag] = s.sample_weights()
        out['params_' + tag] = np.array([temp, sc, 1.0])
        np.random.seed(123)
        st = np.random.get_state()
        picks = [s.sample_replay_level() for _ in range(40)]
        np.random.set_state(st)
        out['u_' + tag] = np.array([np.random.random_sample() for _ in range(40)])
        out['picks_' + tag] = np.array(picks, dtype=np.int64) - 1  # seed -> index
        out['stale_after_' + tag] = s.seed_staleness.copy()
    np.savez_compressed(os.path.join(GOLDEN, 'plr_weights.npz'), **out)
    print('weights fixtures ok')


def gen_transforms():
    """sample_weights + 30 sequential replay draws under the score transforms beyond the shipped configs' constant / rank /
    power: softmax, match, match_rank, eps_greedy (level_sampler.py:752-785), with and without the staleness mix."""
    import numpy as np
    from level_replay import LevelSampler
    from gym import spaces
    rs = np.random.RandomState(11)
    out = {}
    for transform in ('softmax', 'match', 'match_rank', 'eps_greedy'):
        for tag, n, temp, sc in (('a', 600, 0.3, 0.3), ('b', 37, 1.0, 0.0), ('c', 4000, 0.5, 0.1)):
            s = LevelSampler([], {'image': spaces.Box(0, 255, (3, 5, 5), 'uint8')}, spaces.Discrete(7), num_actors=4,
                             strategy='positive_value_loss', score_transform=transform, temperature=temp, eps=0.07, rho=0.5,
                             replay_prob=0.5, staleness_coef=sc, staleness_transform='power', staleness_temperature=1.0,
                             sample_full_distribution=True, seed_buffer_size=n)
            scores = rs.rand(n)
            unseen = (rs.rand(n) < 0.2).astype(np.float64)
            stale = np.floor(rs.rand(n) * 50)
            s.seed_scores[:] = scores
            s.unseen_seed_weights[:] = unseen
            s.seed_staleness[:] = stale
            s.seeds[:] = np.arange(1, n + 1)
            s.working_seed_buffer_size = n
            k = transform + '_' + tag
            out['scores_' + k], out['unseen_' + k], out['stale_' + k] = scores, unseen, stale
            out['weights_' + k] = s.sample_weights()
            out['params_' + k] = np.array([temp, sc, 0.07])
            np.random.seed(321)
            picks = [s.sample_replay_level() for _ in range(30)]
            out['picks_' + k] = np.array(picks, dtype=np.int64) - 1
            out['stale_after_' + k] = s.seed_staleness.copy()
    np.savez_compressed(os.path.join(GOLDEN, 'plr_transforms.npz'), **out)
    print('transform fixtures ok')


class _StubPopart(object):
    def denormalize(self, x):
        return x


class _StubModel(object):
    """A deterministic stand-in critic for the truncated-value path (the NN itself is out of scope): the value of an
    observation is a fixed linear functional of its image plus its direction."""

    def __init__(self):
        import torch
        g = torch.Generator().manual_seed(5)
        self.w = torch.randn(75, generator=g) * 0.05

    def get_value(self, obs, rnn_hxs, masks):
        img = obs['image'].reshape(-1, 75)
        return (img @ self.w).unsqueeze(-1) + 0.1 * obs['direction'].reshape(-1, 1)


def gen_storage():
    """algos/storage.py beyond GAE, executed: discounted returns, get_batched_value_loss variants,
    get_action_traj(as_string), insert / after_update bookkeeping and the truncated-value path
    (use_proper_time_limits) with a stub critic."""
    import numpy as np
    import torch
    from gym import spaces
    from algos.storage import RolloutStorage
    rs = np.random.RandomState(23)
    out = {}
    for tag, T, N in (('a', 256, 32), ('b', 33, 5)):
        st, _ = _storage(T, N, rs)
        nv = torch.from_numpy(rs.randn(N, 1).astype(np.float32))
        acts = rs.randint(0, 7, size=(T, N, 1))
        st.actions.copy_(torch.from_numpy(acts))
        st.compute_returns(nv, False, 0.995, 0.95)
        out['rewards_' + tag] = st.rewards.numpy()[:, :, 0].copy()
        out['values_' + tag] = st.value_preds.numpy()[:, :, 0].copy()
        out['masks_' + tag] = st.masks.numpy()[:, :, 0].copy()
        out['disc_returns_' + tag] = st.returns.numpy()[:, :, 0].copy()
        st.compute_returns(nv, True, 0.995, 0.95)
        out['gae_returns_' + tag] = st.returns.numpy()[:, :, 0].copy()
        k = 0
        for signed, pos in ((False, False), (True, False), (False, True)):
            for power in (1, 2):
                for clipped in (True, False):
                    out['bvl_%s_%d' % (tag, k)] = st.get_batched_value_loss(signed=signed, positive_only=pos, power=power,
                                                                            clipped=clipped, batched=True).numpy()[:, 0].copy()
                    out['bvl_params_%s_%d' % (tag, k)] = np.array([signed, pos, power, clipped], np.int32)
                    k += 1
        out['bvl_scalar_' + tag] = np.float64(st.get_batched_value_loss(signed=False, positive_only=True, batched=False))
        out['actions_' + tag] = acts[:, :, 0].astype(np.int64)
        out['traj_' + tag] = np.array(st.get_action_traj(as_string=True))
    # insert / truncated obs / after_update, then the truncated-value GAE
    T, N = 24, 6
    obs_space = {'image': spaces.Box(0, 255, (3, 5, 5), 'uint8'), 'direction': spaces.Box(0, 3, (1,), 'uint8')}
    st = RolloutStorage(model=_StubModel(), num_steps=T, num_processes=N, observation_space=obs_space,
                        action_space=spaces.Discrete(7), recurrent_hidden_state_size=4, recurrent_arch='lstm',
                        use_proper_time_limits=True)
    st.model.popart = _StubPopart()
    first = {'image': torch.from_numpy(rs.rand(N, 3, 5, 5).astype(np.float32)), 'direction': torch.from_numpy(rs.randint(0, 4, (N, 1)).astype(np.float32))}
    st.copy_obs_to_index(first, 0)
    ins = []
    for t in range(T):
        obs = {'image': torch.from_numpy(rs.rand(N, 3, 5, 5).astype(np.float32)),
               'direction': torch.from_numpy(rs.randint(0, 4, (N, 1)).astype(np.float32))}
        hx = (torch.from_numpy(rs.randn(N, 4).astype(np.float32)), torch.from_numpy(rs.randn(N, 4).astype(np.float32)))
        action = torch.from_numpy(rs.randint(0, 7, (N, 1)))
        logp = torch.from_numpy(rs.randn(N, 1).astype(np.float32))
        logd = torch.from_numpy(rs.randn(N, 7).astype(np.float32))
        val = torch.from_numpy(rs.randn(N, 1).astype(np.float32))
        rew = torch.from_numpy((rs.rand(N, 1) < 0.1).astype(np.float32))
        done = rs.rand(N) < 0.15
        trunc = done & (rs.rand(N) < 0.5)
        if t == T - 1:
            done[:] = True
        masks = torch.from_numpy(1.0 - done.astype(np.float32)).unsqueeze(-1)
        bad = torch.from_numpy(1.0 - trunc.astype(np.float32)).unsqueeze(-1)
        cliff = torch.ones(N, 1)
        seeds = torch.from_numpy(rs.randint(1, 100, (N, 1)).astype(np.int32))
        tr_obs = {}
        for i in np.nonzero(trunc)[0]:
            tr = {'image': rs.rand(3, 5, 5).astype(np.float32), 'direction': rs.randint(0, 4, (1,)).astype(np.float32)}
            st.insert_truncated_obs({k: torch.from_numpy(v) for k, v in tr.items()}, index=int(i))
            tr_obs[int(i)] = tr
        st.insert(obs, hx, action, logp, logd, val, rew, masks, bad, level_seeds=seeds, cliffhanger_masks=cliff)
        ins.append(dict(obs={k: v.numpy() for k, v in obs.items()}, hx=[h.numpy() for h in hx], action=action.numpy(),
                        logp=logp.numpy(), logd=logd.numpy(), val=val.numpy(), rew=rew.numpy(), masks=masks.numpy(),
                        bad=bad.numpy(), cliff=cliff.numpy(), seeds=seeds.numpy(), trunc=tr_obs))
    nv = torch.from_numpy(rs.randn(N, 1).astype(np.float32))
    st.compute_returns(nv, True, 0.995, 0.95)
    final = dict(first={k: v.numpy() for k, v in first.items()}, inserts=ins, next_value=nv.numpy(),
                 obs={k: v.numpy().copy() for k, v in st.obs.items()},
                 truncated_obs={k: v.numpy().copy() for k, v in st.truncated_obs.items()},
                 recurrent_hidden_states=st.recurrent_hidden_states.numpy().copy(), actions=st.actions.numpy().copy(),
                 action_log_probs=st.action_log_probs.numpy().copy(), action_log_dist=st.action_log_dist.numpy().copy(),
                 value_preds=st.value_preds.numpy().copy(), rewards=st.rewards.numpy().copy(), masks=st.masks.numpy().copy(),
                 bad_masks=st.bad_masks.numpy().copy(), level_seeds=st.level_seeds.numpy().copy(),
                 truncated_value_preds=st.truncated_value_preds.numpy().copy(), returns=st.returns.numpy().copy(),
                 stub_w=st.model.w.numpy().copy(), step=st.step)
    st.after_update()
    final['after_update'] = dict(obs0={k: v[0].numpy().copy() for k, v in st.obs.items()}, masks0=st.masks[0].numpy().copy(),
                                 bad0=st.bad_masks[0].numpy().copy(), rnn0=st.recurrent_hidden_states[0].numpy().copy())
    np.savez_compressed(os.path.join(GOLDEN, 'plr_storage.npz'), **out)
    with gzip.open(os.path.join(GOLDEN, 'plr_storage_session.pkl.gz'), 'wb') as f:
        pickle.dump(final, f, protocol=4)
    print('storage fixtures', len(out), 'arrays + 1 session')


def gen_score_functions():
    """The reference's own per-episode score functions (level_sampler.py:288-306,425-437) on random episodes, called the
    way _update_with_rollouts calls them (log_softmax of the stored logits): known answers for the numpy oracle."""
    import numpy as np
    import torch
    from level_replay import LevelSampler
    from gym import spaces
    rs = np.random.RandomState(41)
    s = LevelSampler([], {'image': spaces.Box(0, 255, (3, 5, 5), 'uint8')}, spaces.Discrete(7), num_actors=1,
                     strategy='least_confidence', gamma=0.995, seed_buffer_size=4)
    out = {'gamma': 0.995}
    lens = [1, 2, 3, 7, 19, 64, 250]
    for k, L in enumerate(lens):
        logits = (rs.randn(L, 7) * 2).astype(np.float32)
        rewards = (rs.rand(L) < 0.2).astype(np.float32) * rs.rand(L).astype(np.float32)
        values = rs.randn(L).astype(np.float32)
        lp = torch.log_softmax(torch.from_numpy(logits), -1)
        kw = dict(episode_logits=lp, rewards=torch.from_numpy(rewards).unsqueeze(-1), value_preds=torch.from_numpy(values).unsqueeze(-1))
        out['logits_%d' % k], out['rewards_%d' % k], out['values_%d' % k] = logits, rewards, values
        out['lc_%d' % k] = np.array(s._average_least_confidence(**kw), np.float64)
        out['mm_%d' % k] = np.array(s._average_min_margin(**kw), np.float64)
        out['td_%d' % k] = np.array(s._one_step_td_error(**kw), np.float64)
    out['n'] = len(lens)
    np.savez_compressed(os.path.join(GOLDEN, 'plr_score_functions.npz'), **out)
    print('score function fixtures', len(lens), 'episodes')


def gen_plr():
    rh.activate()
    gen_gae()
    gen_score_functions()
    gen_storage()
    gen_weights()
    gen_transforms()
    gen_sampler()


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'transforms':   # only the (round-2) transform fixtures
        rh.activate()
        gen_transforms()
    else:
        gen_plr()
