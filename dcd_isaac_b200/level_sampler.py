"""LevelSampler -- drop-in for level_replay/level_sampler.py (PLR / robust PLR / ACCEL level curriculum).

Same constructor arguments (util.make_plr_args, util/__init__.py:230-252), public methods and attributes as the
reference class (level_replay/level_sampler.py:19-800), picklable like it (the runner checkpoints the object,
envs/runners/adversarial_runner.py:214-215).  What moves to the B200:

  * update_with_rollouts: the per-actor episode segmentation and the score reductions (mean / max of the
    clamped advantages, reward and value sums for MaxMC) run in mgplr_plr_episode_scores over the rollout
    tensors where they already live; the host only walks the compact episode records, in the reference's
    actor-major / time-minor order, to apply the (order-dependent) buffer admission rules.
  * sample_weights / sample_replay_level(s): rank / power transforms, staleness mixing, the inverse-CDF draw and
    the staleness update run in mgplr_plr_sample_weights / mgplr_plr_sample_replay (fp64, one CTA).

Host state is numpy (scores, staleness, seeds ... exactly the reference's arrays) so pickling, `seeds`,
`seed_scores`, `staging_seed_set`, `working_seed_set` behave as before.  Random decisions consume the GLOBAL
np.random stream exactly where the reference does (level_sampler.py:611,616,674): one random_sample() per
replay decision (only when the fill test passes) and one per replay draw.

Deviations (documented in DESIGN.md): ties in the rank transform are broken by index (the reference inherits
numpy's unspecified quicksort order); only the constant / rank / power transforms and the value-based score
strategies are implemented; a rollout whose last step is not `done` (never produced by the reference runner)
is scored on [start, T) instead of the reference's off-by-one slices.
"""
from collections import defaultdict

import numpy as np

INT32_MAX = 2147483647

_TRANSFORMS = {'constant': 0, 'rank': 1, 'power': 2}
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
        self.obs_space = obs_space
        self.action_space = action_space
        self.num_actors = num_actors
        self.strategy = strategy
        self.max_score_coef = max_score_coef
        self.replay_schedule = replay_schedule
        self.score_transform = score_transform
        self.temperature = temperature
        self.eps = eps
        self.rho = rho
        self.replay_prob = replay_prob
        self.alpha = alpha
        self.staleness_coef = staleness_coef
        self.staleness_transform = staleness_transform
        self.staleness_temperature = staleness_temperature
        self.gamma = gamma
        self.use_dense_rewards = use_dense_rewards
        self.device = device
        if strategy.startswith('tscl') or strategy in ('policy_entropy', 'alt_advantage_abs', 'grounded_positive_value_loss'):
            # policy_entropy cannot run in the reference either ([L,7] * [L] broadcast, level_sampler.py:281); alt_returns is not
            # a RolloutStorage buffer; grounded_positive needs a second pass with the host-side grounded value
            raise NotImplementedError('score strategy %r is not part of the B200 build' % strategy)
        if use_dense_rewards and strategy.startswith('grounded'):
            raise NotImplementedError('grounded scores with dense rewards (CarRacing) are out of scope')

        self.seed_buffer_size = seed_buffer_size if not seeds else len(seeds)
        N = self.seed_buffer_size
        self._init_seed_index(seeds)
        self.unseen_seed_weights = np.array([1.] * N)
        self.seed_scores = np.array([0.] * N, dtype=float)
        # The reference's dense [num_actors, N] partial-score arrays (level_sampler.py:80-82) are only ever non-zero for a
        # rollout that does not end in `done`, which the runner never produces (adversarial_runner.py:530).  They are
        # allocated on first use so that 10^5 actors x 4000 slots does not cost gigabytes of host memory.
        self._partials = None
        self._partials_dirty = False
        self.seed_staleness = np.array([0.] * N, dtype=float)
        self.running_sample_count = 0
        self.next_seed_index = 0
        self.track_solvable = False
        self.grounded_values = None
        if self.strategy.startswith('grounded'):
            self.grounded_values = np.array([-np.inf] * N, dtype=float)
        self.sample_full_distribution = sample_full_distribution
        if self.sample_full_distribution:
            self.seed2actor = defaultdict(set)
            self.working_seed_buffer_size = 0
            self.seed_buffer_priority = seed_buffer_priority
            self.staging_seed_set = set()
            self.working_seed_set = set()
            self.seed2timestamp_buffer = {}
            self.partial_seed_scores_buffer = [{} for _ in range(num_actors)]
            self.partial_seed_max_scores_buffer = [{} for _ in range(num_actors)]
            self.partial_seed_steps_buffer = [{} for _ in range(num_actors)]
        self._dev = None  # lazily created device mirrors (not pickled)

    # ------------------------------------------------------------------ lazily allocated partial-score arrays
    def _alloc_partials(self):
        if self._partials is None:
            A, N = self.num_actors, self.seed_buffer_size
            self._partials = (np.zeros((A, N), dtype=float), np.ones((A, N), dtype=float) * float('-inf'),
                              np.zeros((A, N), dtype=np.int32))
        return self._partials

    @property
    def partial_seed_scores(self):
        return self._alloc_partials()[0]

    @property
    def partial_seed_max_scores(self):
        return self._alloc_partials()[1]

    @property
    def partial_seed_steps(self):
        return self._alloc_partials()[2]

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
            N = max(1, self.seed_buffer_size)
            self._dev = {
                'torch': torch, 'lib': _lib.load(), 'dev': dev,
                'scores': torch.zeros(N, dtype=torch.float64, device=dev),
                'stale': torch.zeros(N, dtype=torch.float64, device=dev),
                'unseen': torch.zeros(N, dtype=torch.float64, device=dev),
                'weights': torch.zeros(N, dtype=torch.float64, device=dev),
                'w_score': torch.zeros(N, dtype=torch.float64, device=dev),
                'w_key': None,   # (scores, unseen) the cached score weights were computed from
            }
        return self._dev

    def _upload(self):
        """Host arrays -> device mirrors.  Scores / unseen flags only change in update paths, so their upload and the score
        half of the weights (the sort) are skipped while they are unchanged; staleness is uploaded every time."""
        from . import _lib
        d = self._device_ctx()
        t = d['torch']
        d['stale'].copy_(t.from_numpy(np.ascontiguousarray(self.seed_staleness, dtype=np.float64)))
        key = d['w_key']
        if key is None or not (np.array_equal(key[0], self.seed_scores) and np.array_equal(key[1], self.unseen_seed_weights)):
            d['scores'].copy_(t.from_numpy(np.ascontiguousarray(self.seed_scores, dtype=np.float64)))
            d['unseen'].copy_(t.from_numpy(np.ascontiguousarray(self.unseen_seed_weights, dtype=np.float64)))
            st, temp, eps, _, _, _ = self._weight_args()
            _lib.check(d['lib'].mgplr_plr_score_weights(_lib.ptr(d['scores']), _lib.ptr(d['unseen']), self.seed_buffer_size, st,
                                                        temp, eps, _lib.ptr(d['w_score']),
                                                        t.cuda.current_stream(d['dev']).cuda_stream), 'mgplr_plr_score_weights')
            d['w_key'] = (self.seed_scores.copy(), self.unseen_seed_weights.copy())
        return d

    def _transform_code(self, name):
        if name not in _TRANSFORMS:
            raise NotImplementedError('transform %r is not part of the B200 build (constant, rank, power are)' % name)
        return _TRANSFORMS[name]

    def _weight_args(self):
        eps = 0.0 if self.staleness_coef > 0 else 1e-3  # level_sampler.py:771
        return (self._transform_code(self.score_transform), float(self.temperature), eps, float(self.staleness_coef),
                self._transform_code(self.staleness_transform), float(self.staleness_temperature))

    # ------------------------------------------------------------------ reference API
    def seed_range(self):
        if not self.sample_full_distribution:
            return (int(min(self.seeds)), int(max(self.seeds)))
        return (0, INT32_MAX)

    def _init_seed_index(self, seeds):
        if seeds:
            self.seeds = np.array(seeds, dtype=np.int64)
            self.seed2index = {seed: i for i, seed in enumerate(seeds)}
        else:
            self.seeds = np.zeros(self.seed_buffer_size, dtype=np.int64) - 1
            self.seed2index = {}

    def _init_solvable_tracking(self):
        self.track_solvable = True
        self.staging_seed2solvable = {}
        self.seed_solvable = np.ones(self.seed_buffer_size, dtype=bool)

    @property
    def _proportion_filled(self):
        if self.sample_full_distribution:
            return self.working_seed_buffer_size / self.seed_buffer_size
        num_unseen = (self.unseen_seed_weights > 0).sum()
        return (len(self.seeds) - num_unseen) / len(self.seeds)

    @property
    def requires_value_buffers(self):
        return self.strategy in ['gae', 'value_l1', 'signed_value_loss', 'positive_value_loss',
                                 'grounded_signed_value_loss', 'grounded_positive_value_loss', 'one_step_td_error',
                                 'alt_advantage_abs', 'tscl_window']

    @property
    def _has_working_seed_buffer(self):
        return not self.sample_full_distribution or (self.sample_full_distribution and self.seed_buffer_size > 0)

    # ---- score bookkeeping: level_sampler.py:185-273, unchanged semantics ----
    def update_seed_score(self, actor_index, seed, score, max_score, num_steps):
        if self.sample_full_distribution and seed in self.staging_seed_set:
            return self._partial_update_seed_score_buffer(actor_index, seed, score, num_steps, done=True)
        return self._partial_update_seed_score(actor_index, seed, score, max_score, num_steps, done=True)

    def _partial_update_seed_score(self, actor_index, seed, score, max_score, num_steps, done=False):
        seed_idx = self.seed2index.get(seed, -1)
        if seed_idx < 0:
            return 0, None
        if self._partials is None and done:
            partial_score, partial_max_score, partial_num_steps = np.float64(0.), float('-inf'), np.int32(0)
        else:
            partial_score = self.partial_seed_scores[actor_index][seed_idx]
            partial_max_score = self.partial_seed_max_scores[actor_index][seed_idx]
            partial_num_steps = self.partial_seed_steps[actor_index][seed_idx]
        running_num_steps = partial_num_steps + num_steps
        merged_score = partial_score + (score - partial_score) * num_steps / float(running_num_steps)
        merged_max_score = max(partial_max_score, max_score)
        if done:
            if self._partials is not None:
                self.partial_seed_scores[actor_index][seed_idx] = 0.
                self.partial_seed_max_scores[actor_index][seed_idx] = float('-inf')
                self.partial_seed_steps[actor_index][seed_idx] = 0
            self.unseen_seed_weights[seed_idx] = 0.
            old_score = self.seed_scores[seed_idx]
            total_score = self.max_score_coef * merged_max_score + (1 - self.max_score_coef) * merged_score
            self.seed_scores[seed_idx] = (1 - self.alpha) * old_score + self.alpha * total_score
        else:
            self.partial_seed_scores[actor_index][seed_idx] = merged_score
            self.partial_seed_max_scores[actor_index][seed_idx] = merged_max_score
            self.partial_seed_steps[actor_index][seed_idx] = running_num_steps
            self._partials_dirty = True
        return merged_score, seed_idx

    @property
    def _next_buffer_index(self):
        if self._proportion_filled < 1.0:
            return self.working_seed_buffer_size
        if self.seed_buffer_priority == 'replay_support':
            return self.sample_weights().argmin()
        return self.seed_scores.argmin()

    def _partial_update_seed_score_buffer(self, actor_index, seed, score, num_steps, done=False):
        seed_idx = -1
        self.seed2actor[seed].add(actor_index)
        partial_score = self.partial_seed_scores_buffer[actor_index].get(seed, 0)
        partial_num_steps = self.partial_seed_steps_buffer[actor_index].get(seed, 0)
        running_num_steps = partial_num_steps + num_steps
        merged_score = partial_score + (score - partial_score) * num_steps / float(running_num_steps)
        if done:
            seed_idx = self._next_buffer_index
            if self.seed_scores[seed_idx] <= merged_score or self.unseen_seed_weights[seed_idx] > 0:
                self.unseen_seed_weights[seed_idx] = 0.
                self.working_seed_set.discard(self.seeds[seed_idx])
                self.working_seed_set.add(seed)
                self.seeds[seed_idx] = seed
                self.seed2index[seed] = seed_idx
                self.seed_scores[seed_idx] = merged_score
                if self._partials is not None:
                    self.partial_seed_scores[:, seed_idx] = 0.
                    self.partial_seed_steps[:, seed_idx] = 0
                self.seed_staleness[seed_idx] = self.running_sample_count - self.seed2timestamp_buffer[seed]
                self.working_seed_buffer_size = min(self.working_seed_buffer_size + 1, self.seed_buffer_size)
                if self.track_solvable:
                    self.seed_solvable[seed_idx] = self.staging_seed2solvable.get(seed, True)
            else:
                seed_idx = None
            for a in self.seed2actor[seed]:
                self.partial_seed_scores_buffer[a].pop(seed, None)
                self.partial_seed_steps_buffer[a].pop(seed, None)
            del self.seed2timestamp_buffer[seed]
            del self.seed2actor[seed]
            self.staging_seed_set.remove(seed)
            if self.track_solvable:
                del self.staging_seed2solvable[seed]
        else:
            self.partial_seed_scores_buffer[actor_index][seed] = merged_score
            self.partial_seed_steps_buffer[actor_index][seed] = running_num_steps
        return merged_score, seed_idx

    # ---- rollouts -> scores: level_sampler.py:149-183, 486-578 ----
    def episode_records(self, rollouts):
        """Run the episode-score kernel over a RolloutStorage-like object; returns a numpy record array
        (dcd_isaac_b200._lib.EPISODE_DTYPE) in actor-major / time-minor order."""
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
        T, N = rewards.shape
        value_src = rollouts.denorm_value_preds if getattr(rollouts, 'use_popart', False) else rollouts.value_preds
        values = cu(value_src, t.float32)
        masks = cu(rollouts.masks, t.float32)
        cliff = cu(rollouts.cliffhanger_masks, t.float32)
        returns = cu(rollouts.returns, t.float32)
        seeds = cu(rollouts.level_seeds, t.int32)
        max_eps = int(N) * (int(T) + 1)
        code = _KERNEL_STRATEGY[self.strategy]
        logits, n_act = None, 0
        if code in (4, 5):  # policy-logit strategies read RolloutStorage.action_log_dist [T,N,A] (level_sampler.py:512-513)
            logits = rollouts.action_log_dist.detach().to(dev).to(t.float32).contiguous()
            n_act = int(logits.shape[-1])
        ep = t.zeros(max_eps, 10, dtype=t.int32, device=dev)
        n_ep = t.zeros(1, dtype=t.int32, device=dev)
        stream = t.cuda.current_stream(dev).cuda_stream
        _lib.check(L.mgplr_plr_episode_scores_ex(_lib.ptr(masks), _lib.ptr(cliff), _lib.ptr(returns), _lib.ptr(values),
                                                 _lib.ptr(rewards), _lib.ptr(seeds), _lib.ptr(logits), n_act, float(self.gamma),
                                                 int(T), int(N), code, _lib.ptr(ep), max_eps, _lib.ptr(n_ep), stream),
                   'mgplr_plr_episode_scores')
        n = int(n_ep.item())
        rec = ep[:n].cpu().numpy().view(np.dtype(_lib.EPISODE_DTYPE)).reshape(-1)
        return rec

    def update_with_rollouts(self, rollouts):
        if self.strategy in ['random', 'off']:
            return
        if self.strategy not in _KERNEL_STRATEGY:
            raise ValueError(f'Unsupported strategy, {self.strategy}')
        if not self._has_working_seed_buffer:
            return
        rec = self.episode_records(rollouts)
        self._apply_episode_records(rec)

    def _apply_episode_records(self, rec, vectorize=None):
        """Apply episode records in the reference's order.  Records of seeds in the staging set take the order-dependent
        admission path one by one; the runs between them are independent per buffer slot and are applied with numpy
        (occurrence by occurrence for a slot that appears several times), which is what makes 10^5 actors practical.
        Bit-identical to the sequential walk (tests/test_level_sampler_host.py)."""
        if vectorize is None:
            vectorize = len(rec) >= 256
        if (not vectorize) or self._partials_dirty or len(rec) == 0:
            for r in rec:
                self._apply_one_record(r)
            return
        if (rec['cliffhanger'] == 2).any():  # a not-done tail: partial bookkeeping, sequential
            for r in rec:
                self._apply_one_record(r)
            return
        n = len(rec)
        seeds = rec['seed']
        lo = 0
        if self.sample_full_distribution and self.staging_seed_set:
            cand = np.nonzero(np.isin(seeds, np.fromiter(self.staging_seed_set, dtype=np.int64)) & (rec['cliffhanger'] != 1))[0]
        else:
            cand = ()
        for pos in cand:
            if int(seeds[pos]) not in self.staging_seed_set:
                continue  # already admitted / rejected by an earlier record of this rollout
            self._apply_batch(rec[lo:pos])
            self._apply_one_record(rec[pos])
            lo = pos + 1
        self._apply_batch(rec[lo:n])

    def _apply_batch(self, rec):
        """Vectorised application of records none of whose seeds is in the staging set (all `done`)."""
        rec = rec[rec['cliffhanger'] == 0]
        if len(rec) == 0:
            return
        if len(self.seed2index) == 0:
            return
        keys = np.fromiter(self.seed2index.keys(), dtype=np.int64, count=len(self.seed2index))
        vals = np.fromiter(self.seed2index.values(), dtype=np.int64, count=len(self.seed2index))
        order = np.argsort(keys)
        keys, vals = keys[order], vals[order]
        s = rec['seed'].astype(np.int64)
        pos = np.searchsorted(keys, s)
        pos[pos >= len(keys)] = len(keys) - 1
        known = keys[pos] == s
        rec = rec[known]
        if len(rec) == 0:
            return
        idx = vals[pos[known]]
        # occurrence number of every record within its slot, in record order
        o = np.argsort(idx, kind='stable')
        sorted_idx = idx[o]
        first = np.r_[True, sorted_idx[1:] != sorted_idx[:-1]]
        start = np.maximum.accumulate(np.where(first, np.arange(len(o)), 0))
        occ = np.empty(len(o), dtype=np.int64)
        occ[o] = np.arange(len(o)) - start
        nsteps = (rec['t_end'] - rec['t_start']).astype(np.float64)
        mean_s, max_s = rec['mean_score'].astype(np.float64), rec['max_score'].astype(np.float64)
        grounded = self.grounded_values is not None
        if self.strategy == 'uniform':
            mean_s, max_s = np.ones_like(mean_s), np.ones_like(max_s)
        for k in range(int(occ.max()) + 1):
            m = occ == k
            i_k, n_k = idx[m], nsteps[m]
            if grounded:
                gv = np.maximum(self.grounded_values[i_k], rec['reward_sum'][m].astype(np.float64))
                score = ((0 + n_k) / n_k) * (gv - rec['value_sum'][m].astype(np.float64) / n_k)
                mx = gv - rec['value_min'][m].astype(np.float64)
            else:
                score, mx = mean_s[m], max_s[m]
            merged = 0.0 + (score - 0.0) * n_k / n_k
            total = self.max_score_coef * mx + (1 - self.max_score_coef) * merged
            self.unseen_seed_weights[i_k] = 0.
            self.seed_scores[i_k] = (1 - self.alpha) * self.seed_scores[i_k] + self.alpha * total
            if grounded:
                self.grounded_values[i_k] = gv

    def _apply_one_record(self, r):
        grounded = self.grounded_values is not None
        if True:
            actor, seed_t, n = int(r['actor']), int(r['seed']), int(r['t_end'] - r['t_start'])
            cl = int(r['cliffhanger'])
            if cl == 1:  # cliffhanger episodes are skipped (level_sampler.py:527-528)
                return
            done = cl != 2
            score, max_score, grounded_value = float(r['mean_score']), float(r['max_score']), None
            if self.strategy == 'uniform':
                score, max_score = 1.0, 1.0
            elif grounded:
                # _average_grounded_signed_value_loss (level_sampler.py:351-386) from the per-episode sums
                seed_idx = self.seed2index.get(seed_t, None)
                if done:
                    gv_ = float(r['reward_sum'])
                    grounded_value = max(self.grounded_values[seed_idx], gv_) if seed_idx is not None else gv_
                if self.sample_full_distribution and seed_t in self.partial_seed_steps_buffer[actor]:
                    partial_steps = self.partial_seed_steps_buffer[actor][seed_t]
                elif seed_idx is not None and self._partials is not None:
                    partial_steps = self.partial_seed_steps[actor][seed_idx]
                else:
                    partial_steps = 0
                if done and grounded_value is not None:
                    score = ((partial_steps + n) / n) * (grounded_value - float(r['value_sum']) / n)
                    max_score = grounded_value - float(r['value_min'])
                else:
                    score, max_score = 0, 0
            if done:
                _, seed_idx = self.update_seed_score(actor, seed_t, score, max_score, n)
                if seed_idx is not None and grounded and grounded_value is not None:
                    self.grounded_values[seed_idx] = grounded_value
            elif self.sample_full_distribution and seed_t in self.staging_seed_set:
                self._partial_update_seed_score_buffer(actor, seed_t, score, n)
            else:
                self._partial_update_seed_score(actor, seed_t, score, max_score, n)

    def after_update(self):
        """level_sampler.py:580-599: flush non-zero partial scores as finished episodes with score 0."""
        if not self._has_working_seed_buffer:
            return
        if self._partials is not None:
            for actor_index, seed_idx in zip(*np.nonzero(self.partial_seed_scores)):
                if self.partial_seed_scores[actor_index][seed_idx] != 0:
                    self.update_seed_score(actor_index, self.seeds[seed_idx], 0, float('-inf'), 0)
            self.partial_seed_scores.fill(0)
            self.partial_seed_steps.fill(0)
        self._partials_dirty = False
        if self.sample_full_distribution:
            for actor_index in range(self.num_actors):
                for seed in list(self.partial_seed_scores_buffer[actor_index].keys()):
                    if self.partial_seed_scores_buffer[actor_index][seed] > 0:
                        self.update_seed_score(actor_index, seed, 0, float('-inf'), 0)

    def _update_staleness(self, selected_idx):
        if self.staleness_coef > 0:
            self.seed_staleness = self.seed_staleness + 1
            self.seed_staleness[selected_idx] = 0

    # ---- decisions and sampling: level_sampler.py:606-724 ----
    def sample_replay_decision(self):
        if self.sample_full_distribution:
            proportion_filled = self._proportion_filled
            if self.seed_buffer_size > 0:
                if self.replay_schedule == 'fixed':
                    return bool(proportion_filled >= self.rho and np.random.rand() < self.replay_prob)
                return bool(proportion_filled >= self.rho and
                            np.random.rand() < min(proportion_filled, self.replay_prob))
            return False
        elif self.replay_schedule == 'fixed':
            proportion_seen = self._proportion_filled
            if proportion_seen >= self.rho:
                if np.random.rand() < self.replay_prob or not proportion_seen < 1.0:
                    return True
            return False
        else:
            proportion_seen = self._proportion_filled
            return bool(proportion_seen >= self.rho and np.random.rand() < proportion_seen)

    @property
    def is_warm(self):
        return self._proportion_filled >= self.rho

    def observe_external_unseen_sample(self, seeds, solvable=None):
        for i, seed in enumerate(seeds):
            self.running_sample_count += 1
            if not (seed in self.staging_seed_set or seed in self.working_seed_set):
                self.seed2timestamp_buffer[seed] = self.running_sample_count
                self.staging_seed_set.add(seed)
                if solvable is not None:
                    if not self.track_solvable:
                        self._init_solvable_tracking()
                    self.staging_seed2solvable[seed] = solvable[i]
            else:
                seed_idx = self.seed2index.get(seed, None)
                if seed_idx is not None:
                    self._update_staleness(seed_idx)

    def sample_weights(self):
        """level_sampler.py:726-750 on the device (fp64); returns a host numpy array."""
        if not (self.unseen_seed_weights < 1).any():
            raise FloatingPointError('invalid value encountered in divide')  # np.seterr(all='raise') in the reference
        from . import _lib
        d = self._upload()
        t = d['torch']
        st, temp, eps, coef, stt, stemp = self._weight_args()
        _lib.check(d['lib'].mgplr_plr_sample_weights(_lib.ptr(d['scores']), _lib.ptr(d['stale']), _lib.ptr(d['unseen']),
                                                     self.seed_buffer_size, st, temp, eps, coef, stt, stemp,
                                                     _lib.ptr(d['w_score']), _lib.ptr(d['weights']),
                                                     t.cuda.current_stream(d['dev']).cuda_stream),
                   'mgplr_plr_sample_weights')
        return d['weights'].cpu().numpy().copy()

    def sample_replay_levels(self, n, update_staleness=True):
        """n sequential sample_replay_level() draws in ONE kernel launch; consumes n random_sample() values."""
        from . import _lib
        if not update_staleness:
            raise NotImplementedError('update_staleness=False')
        d = self._upload()
        t = d['torch']
        u = np.array([np.random.random_sample() for _ in range(n)], dtype=np.float64)
        du = t.from_numpy(u).to(d['dev'])
        out = t.zeros(n, dtype=t.int32, device=d['dev'])
        st, temp, eps, coef, stt, stemp = self._weight_args()
        _lib.check(d['lib'].mgplr_plr_sample_replay(_lib.ptr(d['scores']), _lib.ptr(d['stale']), _lib.ptr(d['unseen']),
                                                    self.seed_buffer_size, st, temp, eps, coef, stt, stemp, _lib.ptr(d['w_score']),
                                                    _lib.ptr(du), n, _lib.ptr(out), t.cuda.current_stream(d['dev']).cuda_stream),
                   'mgplr_plr_sample_replay')
        idx = out.cpu().numpy()
        if self.staleness_coef > 0:
            self.seed_staleness = d['stale'].cpu().numpy().copy()
        return [int(self.seeds[i]) for i in idx]

    def sample_replay_level(self, update_staleness=True):
        return self._sample_replay_level(update_staleness=update_staleness)

    def _sample_replay_level(self, update_staleness=True):
        return self.sample_replay_levels(1, update_staleness=update_staleness)[0]

    def _sample_unseen_level(self):
        if self.sample_full_distribution:
            seed = int(np.random.randint(1, INT32_MAX))
            while seed in self.staging_seed_set or seed in self.working_seed_set:
                seed = int(np.random.randint(1, INT32_MAX))
            self.seed2timestamp_buffer[seed] = self.running_sample_count
            self.staging_seed_set.add(seed)
        else:
            sample_weights = self.unseen_seed_weights / self.unseen_seed_weights.sum()
            seed_idx = np.random.choice(range(len(self.seeds)), 1, p=sample_weights)[0]
            seed = self.seeds[seed_idx]
            self._update_staleness(seed_idx)
        return int(seed)

    def sample(self, strategy=None):
        if strategy == 'full_distribution':
            raise ValueError('One-off sampling via full_distribution strategy is not supported.')
        self.running_sample_count += 1
        if not strategy:
            strategy = self.strategy
        if not self.sample_full_distribution:
            if strategy == 'random':
                seed_idx = np.random.choice(range((len(self.seeds))))
                return int(self.seeds[seed_idx])
            if strategy == 'sequential':
                seed_idx = self.next_seed_index
                self.next_seed_index = (self.next_seed_index + 1) % len(self.seeds)
                return int(self.seeds[seed_idx])
        if self.sample_replay_decision():
            return self._sample_replay_level()
        return self._sample_unseen_level()

    @property
    def solvable_mass(self):
        if self.track_solvable:
            return np.sum(self.sample_weights()[self.seed_solvable])
        return 1.

    @property
    def max_score(self):
        return max(self.seed_scores)
