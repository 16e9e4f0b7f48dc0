"""LevelSampler -- drop-in for level_replay/level_sampler.py (PLR / robust PLR / ACCEL level curriculum).

Same constructor arguments (util.make_plr_args, util/__init__.py:230-252), public methods and attributes as the
reference class (level_replay/level_sampler.py:19-800), picklable like it (the runner checkpoints the object,
envs/runners/adversarial_runner.py:214-215).  The arithmetic runs on the B200:

  * update_with_rollouts: per-actor episode segmentation + score reductions (mgplr_plr_episode_scores_ex) over the rollout
    tensors where they already live, then ONE single-CTA kernel (mgplr_plr_apply_records) walks the episode records in the
    reference's actor-major / time-minor order and applies the order-dependent buffer rules -- EWA score update of working
    seeds, staging -> working admission with eviction of the least-supported slot (argmin of sample_weights, ranks kept
    incrementally) -- on the buffer arrays in HBM.  Nothing per-record happens in Python.
  * sample_weights / sample_replay_level(s): rank / power transforms, staleness mixing, the inverse-CDF draws and the
    staleness updates (mgplr_plr_sample_weights / mgplr_plr_sample_replay, fp64, one CTA, n draws per launch).

The host keeps what the runner and the checkpoints look at: the numpy arrays `seeds`, `seed_scores`, `seed_staleness`,
`unseen_seed_weights` (refreshed from HBM after every update) and the dict / set views `seed2index`, `staging_seed_set`,
`working_seed_set`, `seed2timestamp_buffer`, which the kernel's compact change list (admissions in order, seeds that left
the staging set) brings up to date.  Random decisions consume the GLOBAL np.random stream exactly where the reference does
(level_sampler.py:611,616,674): one random_sample() per replay decision (only when the fill test passes) and one per draw.

Deviations: ties in the rank transform are broken by index (the reference inherits numpy's unspecified
quicksort order); every transform but `max` and the strategies in _KERNEL_STRATEGY are built; a rollout
whose last step is not `done` (never produced by the reference runner) is scored on [start, T) instead of the reference's
off-by-one slices, and such tails are carried per (actor, seed) instead of per (actor, buffer slot).
"""
import numpy as np

INT32_MAX = 2147483647
MAX_BUFFER = 8192     # one CTA sorts / scans the buffer in shared memory (mgplr_plr.cu kMaxBuf)

_TRANSFORMS = {'constant': 0, 'rank': 1, 'power': 2, 'softmax': 3, 'match': 4, 'match_rank': 5, 'eps_greedy': 6}
_KERNEL_STRATEGY = {  # -> (MGPLR_SCORE_* code)
    'positive_value_loss': 0, 'signed_value_loss': 1, 'gae': 1, 'value_l1': 2,
    'grounded_signed_value_loss': 3, 'uniform': 3,
    'least_confidence': 4, 'min_margin': 5, 'one_step_td_error': 6,
}


class LevelSampler(object):
    def __init__(self, seeds, obs_space, action_space, num_actors=1, strategy='random', max_score_coef=0.0,
                 replay_schedule='fixed', score_transform='power', temperature=1.0, eps=0.05, rho=1.0,
                 replay_prob=0.95, alpha=1.0, staleness_coef=0, staleness_transform='power',
                 staleness_temperature=1.0, sample_full_distribution=False, seed_buffer_size=0,
                 seed_buffer_priority='replay_support', use_dense_rewards=False, tscl_window_size=0, gamma=0.999,
                 device=None):
        self.obs_space, self.action_space, self.num_actors = obs_space, action_space, num_actors
        self.strategy, self.max_score_coef, self.alpha = strategy, max_score_coef, alpha
        self.replay_schedule, self.rho, self.replay_prob = replay_schedule, rho, replay_prob
        self.score_transform, self.temperature, self.eps = score_transform, temperature, eps
        self.staleness_coef, self.staleness_transform = staleness_coef, staleness_transform
        self.staleness_temperature = staleness_temperature
        self.gamma, self.use_dense_rewards, self.device = gamma, use_dense_rewards, device
        if strategy.startswith('tscl') or strategy in ('policy_entropy', 'alt_advantage_abs', 'grounded_positive_value_loss'):
            # policy_entropy cannot run in the reference either ([L,7] * [L] broadcast, level_sampler.py:281); alt_returns is not
            # a RolloutStorage buffer; grounded_positive needs a second pass over the rollout with the walk-time grounded value
            raise NotImplementedError('score strategy %r is not part of the B200 build' % strategy)
        if use_dense_rewards and strategy.startswith('grounded'):
            raise NotImplementedError('grounded scores with dense rewards (CarRacing) are out of scope')
        self.seed_buffer_size = len(seeds) if seeds else seed_buffer_size
        n = self.seed_buffer_size
        if n > MAX_BUFFER:
            raise ValueError('seed_buffer_size / len(seeds) = %d exceeds %d: the weight, draw and bookkeeping kernels hold the '
                             'whole buffer in one CTA (DESIGN.md)' % (n, MAX_BUFFER))
        if seeds:
            self.seeds = np.array(seeds, dtype=np.int64)
            self.seed2index = {s: i for i, s in enumerate(seeds)}
        else:
            self.seeds = np.full(n, -1, dtype=np.int64)
            self.seed2index = {}
        self.seed_scores = np.zeros(n, dtype=float)
        self.seed_staleness = np.zeros(n, dtype=float)
        self.unseen_seed_weights = np.ones(n, dtype=float)
        self.grounded_values = np.full(n, -np.inf, dtype=float) if strategy.startswith('grounded') else None
        self.running_sample_count = 0
        self.next_seed_index = 0
        self.track_solvable = False
        self.sample_full_distribution = sample_full_distribution
        self.seed_buffer_priority = seed_buffer_priority
        self.working_seed_buffer_size = 0
        self.staging_seed_set, self.working_seed_set = set(), set()
        self.seed2timestamp_buffer = {}
        # score / max / steps of the not-done tail of a rollout, per (actor, seed): merged into that actor's next finished
        # episode on the seed, or flushed by after_update (the reference's partial_seed_* arrays, level_sampler.py:80-98)
        self._tails = {}
        self._dev = None  # lazily created device mirrors (not pickled)

    # ------------------------------------------------------------------ pickling (adversarial_runner.py:214-215)
    def __getstate__(self):
        st = dict(self.__dict__)
        st['_dev'] = None
        if st.get('device') is not None:
            st['device'] = str(st['device'])
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._dev = None

    # ------------------------------------------------------------------ device plumbing
    def _device_ctx(self):
        import torch
        from . import _lib
        if self._dev is None:
            if not torch.cuda.is_available():
                raise _lib.MgplrError('LevelSampler needs a CUDA device: the PLR kernels have no CPU fallback')
            dev = torch.device(self.device if self.device is not None else 'cuda')
            n = max(1, self.seed_buffer_size)
            f64 = dict(dtype=torch.float64, device=dev)
            self._dev = {
                'torch': torch, 'lib': _lib.load(), 'dev': dev,
                # one block of the five buffer arrays (scores | staleness | unseen | grounded | seeds) -> one copy each way
                'state': torch.zeros(5, n, **f64), 'state_host': torch.zeros(5, n, dtype=torch.float64).pin_memory(),
                'weights': torch.zeros(n, **f64), 'w_score': torch.zeros(n, **f64),
                'scratch': torch.zeros(4, n, **f64), 'scratch_i': torch.zeros(n, dtype=torch.int32, device=dev),
                'counters': torch.zeros(4, dtype=torch.int32, device=dev),
                'ep': None, 'n_ep': torch.zeros(1, dtype=torch.int32, device=dev),
                'w_key': None,   # what the cached score weights were computed from
            }
        return self._dev

    def _stream(self, d):
        return d['torch'].cuda.current_stream(d['dev']).cuda_stream

    def _push_state(self, d):
        """host arrays -> the HBM block (one pinned staging copy)."""
        h = d['state_host'].numpy()
        h[0], h[1], h[2] = self.seed_scores, self.seed_staleness, self.unseen_seed_weights
        h[3] = self.grounded_values if self.grounded_values is not None else 0.0
        h[4] = self.seeds.view(np.float64) if self.seeds.dtype == np.int64 else self.seeds.astype(np.int64).view(np.float64)
        d['state'].copy_(d['state_host'], non_blocking=True)

    def _pull_state(self, d):
        d['state_host'].copy_(d['state'])
        h = d['state_host'].numpy()
        self.seed_scores, self.seed_staleness, self.unseen_seed_weights = h[0].copy(), h[1].copy(), h[2].copy()
        if self.grounded_values is not None:
            self.grounded_values = h[3].copy()
        self.seeds = h[4].view(np.int64).copy()

    def _upload(self):
        """Host arrays -> device mirrors for the weight / draw kernels.  The score half of the weights (the sort) is cached
        while everything it depends on is unchanged; staleness is uploaded every time."""
        from . import _lib
        d = self._device_ctx()
        self._push_state(d)
        st, temp, eps, _, _, _ = self._weight_args()
        key = d['w_key']
        if key is None or key[2:] != (st, temp, eps) or not (np.array_equal(key[0], self.seed_scores) and
                                                             np.array_equal(key[1], self.unseen_seed_weights)):
            _lib.check(d['lib'].mgplr_plr_score_weights(_lib.ptr(d['state'][0]), _lib.ptr(d['state'][2]), self.seed_buffer_size, st,
                                                        temp, eps, _lib.ptr(d['w_score']), self._stream(d)), 'mgplr_plr_score_weights')
            d['w_key'] = (self.seed_scores.copy(), self.unseen_seed_weights.copy(), st, temp, eps)
        return d

    def _transform_code(self, name):
        if name not in _TRANSFORMS:
            raise NotImplementedError('transform %r is not part of the B200 build (%s are)' % (name, ', '.join(sorted(_TRANSFORMS))))
        return _TRANSFORMS[name]

    def _weight_args(self):
        eps = 0.0 if self.staleness_coef > 0 else 1e-3  # level_sampler.py:771
        if self.score_transform == 'eps_greedy':
            eps = float(self.eps)                        # level_sampler.py:762-764
        return (self._transform_code(self.score_transform), float(self.temperature), eps, float(self.staleness_coef),
                self._transform_code(self.staleness_transform), float(self.staleness_temperature))

    # ------------------------------------------------------------------ small reference API
    def seed_range(self):
        if not self.sample_full_distribution:
            return (int(min(self.seeds)), int(max(self.seeds)))
        return (0, INT32_MAX)

    def _init_solvable_tracking(self):
        self.track_solvable = True
        self.staging_seed2solvable = {}
        self.seed_solvable = np.ones(self.seed_buffer_size, dtype=bool)

    @property
    def _proportion_filled(self):
        if self.sample_full_distribution:
            return self.working_seed_buffer_size / self.seed_buffer_size
        return float((self.unseen_seed_weights <= 0).sum()) / len(self.seeds)

    @property
    def requires_value_buffers(self):
        return self.strategy in ['gae', 'value_l1', 'signed_value_loss', 'positive_value_loss',
                                 'grounded_signed_value_loss', 'grounded_positive_value_loss', 'one_step_td_error',
                                 'alt_advantage_abs', 'tscl_window']

    @property
    def _has_working_seed_buffer(self):
        return (not self.sample_full_distribution) or self.seed_buffer_size > 0

    @property
    def is_warm(self):
        return self._proportion_filled >= self.rho

    @property
    def solvable_mass(self):
        if self.track_solvable:
            return np.sum(self.sample_weights()[self.seed_solvable])
        return 1.

    @property
    def max_score(self):
        return max(self.seed_scores)

    # ------------------------------------------------------------------ rollouts -> episode records (device)
    def _score_rollouts(self, rollouts):
        """Launch the episode-score kernels; the records stay in HBM.  Returns (records tensor [cap, 10] i32, device count
        tensor, capacity, level_seeds tensor)."""
        from . import _lib
        d = self._device_ctx()
        t, L, dev = d['torch'], d['lib'], d['dev']

        def cu(x, dtype):
            x = x.detach()
            if x.device != dev:
                x = x.to(dev)
            x = x.to(dtype)
            if x.dim() == 3:
                x = x[:, :, 0]
            return x.contiguous()
        rewards = cu(rollouts.rewards, t.float32)
        T, N = int(rewards.shape[0]), int(rewards.shape[1])
        value_src = rollouts.denorm_value_preds if getattr(rollouts, 'use_popart', False) else rollouts.value_preds
        values, masks = cu(value_src, t.float32), cu(rollouts.masks, t.float32)
        cliff, returns = cu(rollouts.cliffhanger_masks, t.float32), cu(rollouts.returns, t.float32)
        seeds = cu(rollouts.level_seeds, t.int32)
        code = _KERNEL_STRATEGY[self.strategy]
        logits, n_act = None, 0
        if code in (4, 5):  # policy-logit strategies read RolloutStorage.action_log_dist [T,N,A] (level_sampler.py:512-513)
            logits = rollouts.action_log_dist.detach().to(dev).to(t.float32).contiguous()
            n_act = int(logits.shape[-1])
        # grow-only record buffer: an actor finishes at most T episodes, but a rollout holds a few per actor; start from
        # 4 per actor and let the (exact) count returned by the kernel trigger a regrowth + rerun
        cap = max(4 * N + 64, 0 if d['ep'] is None else int(d['ep'].shape[0]))
        while True:
            if d['ep'] is None or d['ep'].shape[0] < cap:
                d['ep'] = t.empty(cap, 10, dtype=t.int32, device=dev)
            _lib.check(L.mgplr_plr_episode_scores_ex(_lib.ptr(masks), _lib.ptr(cliff), _lib.ptr(returns), _lib.ptr(values),
                                                     _lib.ptr(rewards), _lib.ptr(seeds), _lib.ptr(logits), n_act, float(self.gamma),
                                                     T, N, code, _lib.ptr(d['ep']), cap, _lib.ptr(d['n_ep']), self._stream(d)),
                       'mgplr_plr_episode_scores')
            if cap >= N * (T + 1):
                break
            n = int(d['n_ep'].item())
            if n <= cap:
                break
            cap = min(N * (T + 1), max(n, 2 * cap))
        return d['ep'], d['n_ep'], cap, seeds

    def episode_records(self, rollouts):
        """The episode records of a RolloutStorage-like object as a numpy record array (dcd_isaac_b200._lib.EPISODE_DTYPE) in
        actor-major / time-minor order (what the sharded update all-gathers, distributed.py)."""
        from . import _lib
        ep, n_ep, _, _ = self._score_rollouts(rollouts)
        n = int(n_ep.item())
        return ep[:n].cpu().numpy().view(np.dtype(_lib.EPISODE_DTYPE)).reshape(-1)

    def update_with_rollouts(self, rollouts):
        if self.strategy in ['random', 'off']:
            return
        if self.strategy not in _KERNEL_STRATEGY:
            raise ValueError(f'Unsupported strategy, {self.strategy}')
        if not self._has_working_seed_buffer:
            return
        ep, n_ep, cap, seeds = self._score_rollouts(rollouts)
        t = self._dev['torch']
        table = t.unique(seeds).cpu().numpy().astype(np.int64)   # the rollout's distinct level seeds (sorted)
        self._apply_records_device(ep, n_ep, cap, table)

    def _apply_episode_records(self, rec):
        """Apply a numpy record array (e.g. the all-gathered records of every rank, distributed.update_sampler_sharded)."""
        rec = np.ascontiguousarray(rec)
        if len(rec) == 0:
            return
        d = self._device_ctx()
        t = d['torch']
        ep = t.from_numpy(rec.view(np.int32).reshape(-1, 10)).to(d['dev'])
        self._apply_records_device(ep, None, len(rec), np.unique(rec['seed'].astype(np.int64)), host_rec=rec)

    def _apply_records_device(self, ep, n_ep_dev, n_or_cap, table, host_rec=None, pre=None):
        """Upload the buffer arrays and the seed table, run the bookkeeping kernel, refresh the host views."""
        from . import _lib
        d = self._device_ctx()
        t, dev = d['torch'], d['dev']
        n_u = len(table)
        if n_u == 0:
            return
        if self._tails and pre is None:   # a stored tail may have to be merged into one of these records: host pre-pass
            if host_rec is None:
                n = int(n_ep_dev.item())
                host_rec = ep[:n].cpu().numpy().view(np.dtype(_lib.EPISODE_DTYPE)).reshape(-1)
                n_ep_dev, n_or_cap = None, n
            pre = self._merge_tails(host_rec)
        idx = np.fromiter((self.seed2index.get(int(s), -1) for s in table), dtype=np.int32, count=n_u)
        staging = self.staging_seed_set if self.sample_full_distribution else ()
        stamp = np.fromiter((float(self.seed2timestamp_buffer[int(s)]) if int(s) in staging else -1.0 for s in table),
                            dtype=np.float64, count=n_u)
        self._push_state(d)
        d_table = t.from_numpy(table).to(dev)
        d_idx, d_stamp = t.from_numpy(idx).to(dev), t.from_numpy(stamp).to(dev)
        d_status = t.zeros(n_u, dtype=t.int32, device=dev)
        d_log = t.empty(n_u, 2, dtype=t.int32, device=dev)
        d_uid = t.empty(max(1, n_or_cap), dtype=t.int32, device=dev)
        d_pre = None if pre is None else t.from_numpy(np.ascontiguousarray(pre, dtype=np.float64)).to(dev)
        d['counters'].copy_(t.tensor([0, self.working_seed_buffer_size, 0, 0], dtype=t.int32), non_blocking=True)
        st, temp, eps, coef, stt, stemp = self._weight_args()
        kind = 1 if self.strategy == 'uniform' else (2 if self.grounded_values is not None else 0)
        S = d['state']
        _lib.check(d['lib'].mgplr_plr_apply_records(
            _lib.ptr(ep), _lib.ptr(n_ep_dev), 0 if n_ep_dev is not None else int(n_or_cap), int(n_or_cap), _lib.ptr(d_pre),
            _lib.ptr(d_table), n_u, _lib.ptr(d_idx), _lib.ptr(d_stamp), _lib.ptr(d_status), _lib.ptr(d_log), _lib.ptr(d['counters']),
            _lib.ptr(d_uid), _lib.ptr(S[0]), _lib.ptr(S[1]), _lib.ptr(S[2]), _lib.ptr(S[3]) if kind == 2 else None, _lib.ptr(S[4]),
            self.seed_buffer_size, float(self.running_sample_count), float(self.alpha), float(self.max_score_coef), kind,
            0 if self.seed_buffer_priority == 'replay_support' else 1, st, temp, eps, coef, stt, stemp,
            _lib.ptr(d['scratch']), _lib.ptr(d['scratch_i']), self._stream(d)), 'mgplr_plr_apply_records')
        self._pull_state(d)   # (synchronises)
        counters = d['counters'].cpu().numpy()
        n_adm, self.working_seed_buffer_size, n_tail = int(counters[0]), int(counters[1]), int(counters[2])
        # the change list -> dict / set views
        if n_adm:
            log = d_log[:n_adm].cpu().numpy()
            for u, slot in log.tolist():
                seed = int(table[u])
                self.seed2index[seed] = slot     # (entries of evicted seeds are kept, like the reference's: level_sampler.py:250)
                if self.track_solvable:
                    self.seed_solvable[slot] = self.staging_seed2solvable.get(seed, True)
        if self.sample_full_distribution and len(staging):
            status = d_status.cpu().numpy()
            for u in np.nonzero(status)[0].tolist():
                seed = int(table[u])
                self.staging_seed_set.discard(seed)
                self.seed2timestamp_buffer.pop(seed, None)
                if self.track_solvable:
                    self.staging_seed2solvable.pop(seed, None)
                for key in [k for k in self._tails if k[1] == seed]:
                    del self._tails[key]
            self.working_seed_set = set(self.seeds[self.seeds >= 0].tolist())
        if n_tail:
            if host_rec is None:
                n = int(counters[3])
                host_rec = ep[:n].cpu().numpy().view(np.dtype(_lib.EPISODE_DTYPE)).reshape(-1)
            self._store_tails(host_rec, pre)

    # ------------------------------------------------------------------ not-done tails (never produced by the reference runner)
    def _merge_tails(self, rec):
        """[n, 4] doubles (partial score, partial max, partial steps, flush flag) for the kernel: the stored tail of
        (actor, seed) is merged into that actor's first finished, non-cliffhanger episode on the seed
        (level_sampler.py:199-204,237-241), or extended by another tail."""
        pre = np.zeros((len(rec), 4), np.float64)
        pre[:, 1] = -np.inf
        used = set()
        for r in np.nonzero(rec['cliffhanger'] != 1)[0].tolist():
            key = (int(rec['actor'][r]), int(rec['seed'][r]))
            if key in self._tails and key not in used:
                pre[r, :3] = self._tails[key]
                used.add(key)
                if rec['cliffhanger'][r] == 0:
                    del self._tails[key]
        return pre

    def _store_tails(self, rec, pre):
        for r in np.nonzero(rec['cliffhanger'] == 2)[0].tolist():
            actor, seed, n = int(rec['actor'][r]), int(rec['seed'][r]), int(rec['t_end'][r] - rec['t_start'][r])
            if not (seed in self.staging_seed_set or seed in self.seed2index):
                continue
            ps, pm, pn = (0.0, -np.inf, 0.0) if pre is None else pre[r, :3]
            score, mx = float(rec['mean_score'][r]), float(rec['max_score'][r])
            if self.strategy == 'uniform':
                score = mx = 1.0
            elif self.grounded_values is not None:
                score = mx = 0.0      # (level_sampler.py:383-384: a grounded score needs a finished episode)
            self._tails[(actor, seed)] = (ps + (score - ps) * n / float(pn + n), max(pm, mx), pn + n)

    def after_update(self):
        """level_sampler.py:580-599: what is left of unfinished episodes is scored as it stands (their logits are stale
        after the policy update)."""
        if not self._has_working_seed_buffer or not self._tails:
            return
        from . import _lib
        tails, self._tails = self._tails, {}
        keys = [k for k, v in tails.items() if (v[0] != 0 if k[1] not in self.staging_seed_set else v[0] > 0)]
        # the reference flushes working seeds first (actor-major, slot order), then the staging ones actor by actor
        keys.sort(key=lambda k: (k[1] in self.staging_seed_set, k[0], self.seed2index.get(k[1], -1)))
        if not keys:
            return
        rec = np.zeros(len(keys), dtype=np.dtype(_lib.EPISODE_DTYPE))
        pre = np.zeros((len(keys), 4), np.float64)
        for i, k in enumerate(keys):
            rec['actor'][i], rec['seed'][i] = k
            pre[i] = (tails[k][0], tails[k][1], tails[k][2], 1.0)
        d = self._device_ctx()
        ep = d['torch'].from_numpy(rec.view(np.int32).reshape(-1, 10)).to(d['dev'])
        self._apply_records_device(ep, None, len(rec), np.unique(rec['seed'].astype(np.int64)), host_rec=rec, pre=pre)

    # ------------------------------------------------------------------ decisions and sampling: level_sampler.py:601-724
    def _update_staleness(self, selected_idx):
        if self.staleness_coef > 0:
            self.seed_staleness = self.seed_staleness + 1
            self.seed_staleness[selected_idx] = 0

    def sample_replay_decision(self):
        """Replay or explore (level_sampler.py:606-639).  The uniform is only drawn when the buffer is warm (`and`)."""
        fill = self._proportion_filled
        if self.sample_full_distribution:
            if self.seed_buffer_size <= 0 or fill < self.rho:
                return False
            bar = self.replay_prob if self.replay_schedule == 'fixed' else min(fill, self.replay_prob)
            return bool(np.random.rand() < bar)
        if fill < self.rho:
            return False
        if self.replay_schedule == 'fixed':
            return bool(np.random.rand() < self.replay_prob or not fill < 1.0)
        return bool(np.random.rand() < fill)

    def observe_external_unseen_sample(self, seeds, solvable=None):
        """New levels enter the staging set with a timestamp; known ones only reset their staleness (level_sampler.py:645-659)."""
        for i, seed in enumerate(seeds):
            self.running_sample_count += 1
            if seed in self.staging_seed_set or seed in self.working_seed_set:
                slot = self.seed2index.get(seed, None)
                if slot is not None:
                    self._update_staleness(slot)
                continue
            self.staging_seed_set.add(seed)
            self.seed2timestamp_buffer[seed] = self.running_sample_count
            if solvable is not None:
                if not self.track_solvable:
                    self._init_solvable_tracking()
                self.staging_seed2solvable[seed] = solvable[i]

    def sample_weights(self):
        """level_sampler.py:726-750 on the device (fp64); returns a host numpy array."""
        if not (self.unseen_seed_weights < 1).any():
            raise FloatingPointError('invalid value encountered in divide')  # np.seterr(all='raise') in the reference
        from . import _lib
        d = self._upload()
        st, temp, eps, coef, stt, stemp = self._weight_args()
        S = d['state']
        _lib.check(d['lib'].mgplr_plr_sample_weights(_lib.ptr(S[0]), _lib.ptr(S[1]), _lib.ptr(S[2]), self.seed_buffer_size, st, temp,
                                                     eps, coef, stt, stemp, _lib.ptr(d['w_score']), _lib.ptr(d['weights']),
                                                     self._stream(d)), 'mgplr_plr_sample_weights')
        return d['weights'].cpu().numpy().copy()

    def sample_replay_levels(self, n, update_staleness=True):
        """n sequential sample_replay_level() draws in ONE kernel launch; consumes n random_sample() values."""
        from . import _lib
        d = self._upload()
        t = d['torch']
        u = np.array([np.random.random_sample() for _ in range(n)], dtype=np.float64)
        du = t.from_numpy(u).to(d['dev'])
        out = t.zeros(n, dtype=t.int32, device=d['dev'])
        st, temp, eps, coef, stt, stemp = self._weight_args()
        S = d['state']
        stale = S[1] if update_staleness else S[1].clone()   # the kernel advances staleness between draws in place
        _lib.check(d['lib'].mgplr_plr_sample_replay(_lib.ptr(S[0]), _lib.ptr(stale), _lib.ptr(S[2]), self.seed_buffer_size, st, temp,
                                                    eps, coef, stt, stemp, _lib.ptr(d['w_score']), _lib.ptr(du), n, _lib.ptr(out),
                                                    self._stream(d)), 'mgplr_plr_sample_replay')
        idx = out.cpu().numpy()
        if self.staleness_coef > 0 and update_staleness:
            self.seed_staleness = S[1].cpu().numpy().copy()
        return [int(self.seeds[i]) for i in idx]

    def sample_replay_level(self, update_staleness=True):
        return self._sample_replay_level(update_staleness=update_staleness)

    def _sample_replay_level(self, update_staleness=True):
        return self.sample_replay_levels(1, update_staleness=update_staleness)[0]

    def _sample_unseen_level(self):
        if self.sample_full_distribution:   # a fresh id, stamped with the current sample count (level_sampler.py:692-700)
            while True:
                seed = int(np.random.randint(1, INT32_MAX))
                if seed not in self.staging_seed_set and seed not in self.working_seed_set:
                    break
            self.staging_seed_set.add(seed)
            self.seed2timestamp_buffer[seed] = self.running_sample_count
            return seed
        p = self.unseen_seed_weights / self.unseen_seed_weights.sum()
        slot = np.random.choice(range(len(self.seeds)), 1, p=p)[0]
        self._update_staleness(slot)
        return int(self.seeds[slot])

    def sample(self, strategy=None):
        if strategy == 'full_distribution':
            raise ValueError('One-off sampling via full_distribution strategy is not supported.')
        self.running_sample_count += 1
        strategy = strategy or self.strategy
        if not self.sample_full_distribution and strategy == 'random':
            return int(self.seeds[np.random.choice(range(len(self.seeds)))])
        if not self.sample_full_distribution and strategy == 'sequential':
            slot, self.next_seed_index = self.next_seed_index, (self.next_seed_index + 1) % len(self.seeds)
            return int(self.seeds[slot])
        return self._sample_replay_level() if self.sample_replay_decision() else self._sample_unseen_level()
