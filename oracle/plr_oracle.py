"""numpy restatement of the rollout math of the PLR path (TEST INFRASTRUCTURE ONLY -- never imported by
dcd_isaac_b200).  Pinned by tests/test_plr_oracle.py against fixtures produced by executing the reference
(oracle/gen_golden_plr.py -> tests/golden/plr_*.npz / .pkl.gz).

  gae             RolloutStorage.compute_gae_returns        algos/storage.py:233-256
  discounted_returns / batched_value_loss   algos/storage.py:258-279,290-327
  episode_scores  LevelSampler._update_with_rollouts + score functions   level_replay/level_sampler.py:486-549,307-349
  sample_weights  LevelSampler.sample_weights / _score_transform          level_sampler.py:726-785
  sample_replay   _sample_replay_level + _update_staleness                level_sampler.py:664-680,601-604
"""
import numpy as np


def gae(rewards, values, masks, gamma, gae_lambda):
    """rewards [T,N], values/masks [T+1,N] float32 -> returns [T,N]; float32 op-by-op like the torch loop."""
    f = np.float32
    r, v, m = rewards.astype(f), values.astype(f), masks.astype(f)
    T = r.shape[0]
    g32, gl32 = f(gamma), f(gamma * gae_lambda)
    out = np.zeros_like(r)
    acc = np.zeros(r.shape[1], dtype=f)
    for t in reversed(range(T)):
        delta = (r[t] + (g32 * v[t + 1]) * m[t + 1]) - v[t]
        acc = delta + (gl32 * m[t + 1]) * acc
        out[t] = acc + v[t]
    return out


def discounted_returns(rewards, masks, last_value, gamma):
    """RolloutStorage.compute_discounted_returns (algos/storage.py:258-279): rewards [T,N], masks [T+1,N], last_value [N]
    -> returns [T+1,N], float32 op by op."""
    f = np.float32
    T = rewards.shape[0]
    out = np.zeros((T + 1, rewards.shape[1]), f)
    out[T] = last_value
    g = f(gamma)
    for t in reversed(range(T)):
        out[t] = (out[t + 1] * g) * masks[t + 1].astype(f) + rewards[t].astype(f)
    return out


def batched_value_loss(returns, values, signed=False, positive_only=False, power=1, clipped=True):
    """RolloutStorage.get_batched_value_loss(batched=True) (algos/storage.py:290-327): returns/values [T+1,N] -> [N]."""
    f = np.float32
    td = returns[:-1].astype(f) - values[:-1].astype(f)
    if signed:
        pass
    elif positive_only:
        td = np.maximum(td, f(0))
    else:
        td = np.abs(td)
    p = td.copy()
    for _ in range(1, power):
        p = p * td
    m = p.astype(np.float64).mean(0).astype(f)
    if clipped:
        m = np.clip(m, f(-1), f(1))
    return m


def _logit_scores(logits, strategy):
    """per-step scores of the policy-logit strategies (level_sampler.py:288-306) from raw logits [L, A]."""
    x = logits.astype(np.float32)
    lp = x - x.max(-1, keepdims=True)
    lp = lp - np.log(np.exp(lp).sum(-1, keepdims=True))  # log_softmax (level_sampler.py:513)
    p = np.sort(np.exp(lp), axis=-1)
    if strategy == 'least_confidence':
        return (np.float32(1) - p[:, -1]).astype(np.float32)
    return (p[:, -1] - p[:, -2]).astype(np.float32)


def episode_scores(masks, cliff, returns, values, rewards, seeds, strategy, logits=None, gamma=0.999):
    """Episode records in actor-major / time-minor order: dicts with actor, t_start, t_end, seed, mean, max,
    reward_sum, value_sum, value_min, cliffhanger."""
    T, N = rewards.shape
    recs = []
    for a in range(N):
        start = 0
        for t in range(1, T + 1):
            if masks[t, a] > 0:
                continue
            sl = slice(start, t)
            adv = returns[sl, a].astype(np.float64) - values[sl, a].astype(np.float64) if returns is not None else None
            adv32 = (returns[sl, a] - values[sl, a]).astype(np.float32) if returns is not None else None
            mean = mx = None
            if strategy == 'positive_value_loss':
                sc = np.maximum(adv32, 0)
            elif strategy == 'value_l1':
                sc = np.abs(adv32)
            elif strategy == 'least_confidence':
                sc = _logit_scores(logits[sl, a], strategy)
            elif strategy == 'min_margin':  # mean: 1 - mean(margin), max: 1 - min(margin)  (level_sampler.py:298-306)
                m = _logit_scores(logits[sl, a], strategy)
                mean, mx = 1.0 - float(np.mean(m.astype(np.float64))), float(np.float32(1) - np.min(m))
                sc = m
            elif strategy == 'one_step_td_error':  # level_sampler.py:425-437
                r32, v32 = rewards[sl, a].astype(np.float32), values[sl, a].astype(np.float32)
                if len(r32) > 1:
                    sc = np.abs((r32[:-1] + np.float32(gamma) * v32[1:]) - v32[:-1])
                else:
                    sc = r32[:1] - v32[:1]
            else:
                sc = adv32
            if mean is None:
                mean, mx = float(np.mean(sc.astype(np.float64))), float(np.max(sc))
            recs.append(dict(actor=a, t_start=start, t_end=t, seed=int(seeds[start, a]),
                             mean=mean, max=mx,
                             reward_sum=float(np.sum(rewards[sl, a].astype(np.float64))),
                             value_sum=float(np.sum(values[sl, a].astype(np.float64))), value_min=float(np.min(values[sl, a])),
                             cliffhanger=int(not (cliff[t, a] > 0))))
            start = t
    return recs


def _transform(name, temperature, vals, eps):
    if name == 'constant':
        return np.ones_like(vals)
    if name == 'rank':
        order = np.flip(np.argsort(vals, kind='stable'))  # ties: higher index first (documented tie rule)
        ranks = np.empty_like(order)
        ranks[order] = np.arange(len(order)) + 1
        return 1 / ranks ** (1. / temperature)
    if name == 'power':
        return (np.array(vals).clip(0) + eps) ** (1. / temperature)
    if name == 'softmax':
        return np.exp(np.array(vals) / temperature)
    if name == 'match':
        return ((1 - np.array(vals)) * np.array(vals)) ** (1. / temperature)
    if name == 'match_rank':
        return _transform('rank', temperature, (1 - np.array(vals)) * np.array(vals), eps)
    if name == 'eps_greedy':          # (eps is the sampler's eps here: level_sampler.py:762-764)
        w = np.zeros_like(vals)
        w[np.argmax(vals)] = 1. - eps
        return w + eps / len(vals)
    raise NotImplementedError(name)


def sample_weights(scores, staleness, unseen, score_transform='rank', temperature=0.3, staleness_coef=0.3,
                   staleness_transform='power', staleness_temperature=1.0, sampler_eps=0.05):
    eps = 0 if staleness_coef > 0 else 1e-3
    if score_transform == 'eps_greedy':
        eps = sampler_eps
    w = _transform(score_transform, temperature, scores, eps) * (1 - unseen)
    z = np.sum(w)
    if z > 0:
        w = w / z
    else:
        w = np.ones_like(w) / len(w) * (1 - unseen)
        w = w / np.sum(w)
    if staleness_coef > 0:
        s = _transform(staleness_transform, staleness_temperature, staleness, 0) * (1 - unseen)
        z = np.sum(s)
        s = s / z if z > 0 else 1. / len(s) * (1 - unseen)
        w = (1 - staleness_coef) * w + staleness_coef * s
    return w


def sample_replay(scores, staleness, unseen, u, **kw):
    """Sequential draws with recorded uniforms; returns (indices, final staleness)."""
    staleness = staleness.copy()
    idx = []
    coef = kw.get('staleness_coef', 0.3)
    for uu in u:
        w = sample_weights(scores, staleness, unseen, **kw)
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        i = int(np.searchsorted(cdf, uu, side='right'))
        i = min(i, len(w) - 1)
        idx.append(i)
        if coef > 0:
            staleness = staleness + 1
            staleness[i] = 0
    return np.array(idx), staleness


def sample_replay_closed_form(scores, staleness, unseen, u, **kw):
    """The same sequential draws WITHOUT recomputing the weights: the algebra behind the Fenwick-tree path of k_sample_replay
    (DESIGN.md 4.4), restated with plain prefix sums so that the CPU suite checks it against sample_replay above.
    Valid for staleness transform `power` with temperature 1 and S_0 > 0, >= 2 seen slots (or no staleness mix)."""
    coef = kw.get('staleness_coef', 0.3)
    assert kw.get('staleness_transform', 'power') == 'power' and kw.get('staleness_temperature', 1.0) == 1.0
    seen = 1.0 - unseen
    # normalised, masked score weights: constant during the call (the power transform's eps depends on the REAL coef)
    eps = 0 if coef > 0 else 1e-3
    if kw.get('score_transform', 'rank') == 'eps_greedy':
        eps = kw.get('sampler_eps', 0.05)
    a = _transform(kw.get('score_transform', 'rank'), kw.get('temperature', 0.3), scores, eps) * seen
    a = a / a.sum() if a.sum() > 0 else seen / max(1.0, seen.sum())
    PA, PS, PN = np.cumsum(a), np.cumsum(staleness * seen), np.cumsum(seen)
    S0, n_seen, n = PS[-1], PN[-1], len(scores)
    assert coef == 0 or (S0 > 0 and n_seen >= 2)
    corr = np.zeros(n)            # s0_j + k_j + 1 for picked slots
    lastk = np.full(n, -1)
    idx = []
    for t, uu in enumerate(u):
        if coef > 0:
            St = S0 + t * n_seen - corr.sum()
            cdf = (1 - coef) * PA + coef / St * (PS + t * PN - np.cumsum(corr))
        else:
            cdf = PA.copy()
        cdf = cdf / cdf[-1]
        i = min(int(np.searchsorted(cdf, uu, side='right')), n - 1)
        idx.append(i)
        if coef > 0:
            if seen[i] > 0:
                corr[i] = staleness[i] + t + 1
            lastk[i] = t
    T = len(u)
    final = staleness.copy()
    if coef > 0:
        final = np.where(lastk >= 0, T - 1 - lastk, staleness + T).astype(staleness.dtype)
    return np.array(idx), final


class BufferOracle(object):
    """Sequential restatement of the sampler's buffer bookkeeping for the full-distribution (robust PLR / ACCEL) mode:
    observe_external_unseen_sample (level_sampler.py:645-659), the per-record score update with staging -> working
    admission (update_seed_score / _partial_update_seed_score / _partial_update_seed_score_buffer / _next_buffer_index,
    level_sampler.py:185-273) driven by episode records in the reference's actor-major / time-minor order
    (_update_with_rollouts, :486-549), for finished episodes (the only kind the runner produces, adversarial_runner.py:530).
    Test infrastructure: the CUDA kernel k_apply_records is compared with this walk, and this walk is pinned by the recorded
    reference sessions (tests/test_plr_oracle.py)."""

    def __init__(self, n_buf, strategy='positive_value_loss', alpha=1.0, max_score_coef=0.0, priority='replay_support',
                 score_transform='rank', temperature=0.3, staleness_coef=0.3, staleness_transform='power',
                 staleness_temperature=1.0):
        self.n = n_buf
        self.strategy, self.alpha, self.coef, self.priority = strategy, alpha, max_score_coef, priority
        self.wkw = dict(score_transform=score_transform, temperature=temperature, staleness_coef=staleness_coef,
                        staleness_transform=staleness_transform, staleness_temperature=staleness_temperature)
        self.seeds = np.zeros(n_buf, np.int64) - 1
        self.scores = np.zeros(n_buf)
        self.stale = np.zeros(n_buf)
        self.unseen = np.ones(n_buf)
        self.grounded = np.full(n_buf, -np.inf) if strategy.startswith('grounded') else None
        self.index_of = {}       # seed2index: entries are never deleted (an evicted seed keeps its stale slot, :250)
        self.stamp = {}          # seed2timestamp_buffer: the staging set
        self.count = 0           # running_sample_count
        self.filled = 0          # working_seed_buffer_size

    def observe(self, seeds):
        for s in seeds:
            s = int(s)
            self.count += 1
            if s in self.stamp or s in set(self.seeds[self.seeds >= 0].tolist()):
                i = self.index_of.get(s)
                if i is not None and self.wkw['staleness_coef'] > 0:   # _update_staleness (:601-604)
                    self.stale = self.stale + 1
                    self.stale[i] = 0
            else:
                self.stamp[s] = self.count

    def _free_slot(self):
        if self.filled < self.n:
            return self.filled
        if self.priority == 'replay_support':
            return int(np.argmin(sample_weights(self.scores, self.stale, self.unseen, **self.wkw)))
        return int(np.argmin(self.scores))

    def apply(self, recs):
        """recs: iterable of dicts with actor, t_start, t_end, seed, mean, max, reward_sum, value_sum, value_min, cliffhanger."""
        for r in recs:
            if r['cliffhanger']:
                continue
            seed, n = int(r['seed']), int(r['t_end'] - r['t_start'])
            idx = self.index_of.get(seed)
            score, mx, gv = float(r['mean']), float(r['max']), None
            if self.strategy == 'uniform':
                score = mx = 1.0
            elif self.grounded is not None:      # _average_grounded_signed_value_loss (:351-386) from the episode sums
                gv = float(r['reward_sum'])
                if idx is not None:
                    gv = max(self.grounded[idx], gv)
                score = (n / n) * (gv - float(r['value_sum']) / n)
                mx = gv - float(r['value_min'])
            merged = 0.0 + (score - 0.0) * n / float(n)
            if seed in self.stamp:               # staging seed: admission (:228-273)
                slot = self._free_slot()
                if self.scores[slot] <= merged or self.unseen[slot] > 0:
                    self.unseen[slot] = 0.0
                    self.seeds[slot] = seed
                    self.index_of[seed] = slot
                    self.scores[slot] = merged
                    self.stale[slot] = self.count - self.stamp[seed]
                    self.filled = min(self.filled + 1, self.n)
                    if gv is not None:
                        self.grounded[slot] = gv
                del self.stamp[seed]
            elif idx is not None:                # working (or stale-index) seed: EWA update (:193-216)
                self.unseen[idx] = 0.0
                total = self.coef * max(float('-inf'), mx) + (1 - self.coef) * merged
                self.scores[idx] = (1 - self.alpha) * self.scores[idx] + self.alpha * total
                if gv is not None:
                    self.grounded[idx] = gv
