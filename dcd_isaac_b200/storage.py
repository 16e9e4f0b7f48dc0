"""Rollout-storage pieces of the hot path (algos/storage.py): buffer layouts the step kernel writes into and
RolloutStorage.compute_gae_returns on the device.

`RolloutStorage` is the drop-in for the reference's class (same constructor, attributes and methods, tensors
resident in HBM; returns / value-loss reductions run as kernels).  `compute_gae_returns(storage, next_value,
gamma, gae_lambda)` can also be called on the reference's own RolloutStorage object (it only touches .rewards/.value_preds/.masks/.returns and, like the reference,
.truncated_value_preds / .denorm_value_preds when those are in use); `DeviceRolloutStorage` is a minimal
storage with the same tensor names and shapes (algos/storage.py:62-112) for the fused rollout path."""
import torch

from . import _lib


def gae_returns(rewards, value_preds, masks, returns, gamma, gae_lambda):
    """returns[t] for t < T from rewards [T,N,1], value_preds / masks / returns [T+1,N,1] (float32 CUDA,
    contiguous).  Bit-identical to the torch loop at algos/storage.py:251-256."""
    for x in (rewards, value_preds, masks, returns):
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise _lib.MgplrError('gae_returns needs contiguous float32 CUDA tensors (no CPU fallback)')
    T, N = rewards.shape[0], rewards.shape[1]
    L = _lib.load()
    _lib.check(L.mgplr_gae(_lib.ptr(rewards), _lib.ptr(value_preds), _lib.ptr(masks), _lib.ptr(returns), T, N,
                           float(gamma), float(gae_lambda), torch.cuda.current_stream(rewards.device).cuda_stream),
               'mgplr_gae')
    return returns


def compute_gae_returns(storage, next_value, gamma, gae_lambda):
    """RolloutStorage.compute_gae_returns (algos/storage.py:233-256) with the reverse scan on the device."""
    storage.value_preds[-1] = next_value
    value_preds = storage.value_preds
    if getattr(storage, 'use_proper_time_limits', False):
        storage._compute_truncated_value_preds()
        value_preds = storage.truncated_value_preds
    if getattr(storage, 'use_popart', False):
        storage.denorm_value_preds = storage.model.popart.denormalize(value_preds)
        value_preds = storage.denorm_value_preds
    return gae_returns(storage.rewards, value_preds.contiguous(), storage.masks, storage.returns, gamma, gae_lambda)


class DeviceRolloutStorage(object):
    """The RolloutStorage tensors of the student rollout (algos/storage.py:62-112), resident in HBM.  The step
    kernel writes obs[t+1], rewards[t], masks[t+1], bad_masks[t+1], cliffhanger_masks[t+1] in place."""

    def __init__(self, num_steps, num_processes, device='cuda', num_actions=7):
        T, N = num_steps, num_processes
        dev = torch.device(device)
        self.num_steps, self.num_processes, self.device = T, N, dev
        self.obs = {'image': torch.zeros(T + 1, N, 3, 5, 5, device=dev), 'direction': torch.zeros(T + 1, N, 1, device=dev)}
        self.truncated_obs = {'image': torch.zeros(T + 1, N, 3, 5, 5, device=dev),
                              'direction': torch.zeros(T + 1, N, 1, device=dev)}
        self.rewards = torch.zeros(T, N, 1, device=dev)
        self.value_preds = torch.zeros(T + 1, N, 1, device=dev)
        self.returns = torch.zeros(T + 1, N, 1, device=dev)
        self.action_log_dist = torch.zeros(T, N, num_actions, device=dev)
        self.actions = torch.zeros(T, N, 1, dtype=torch.long, device=dev)
        self.masks = torch.ones(T + 1, N, 1, device=dev)
        self.bad_masks = torch.ones(T + 1, N, 1, device=dev)
        self.cliffhanger_masks = torch.ones(T + 1, N, 1, device=dev)
        self.level_seeds = torch.zeros(T, N, 1, dtype=torch.int, device=dev)
        self.flags = torch.zeros(T, N, dtype=torch.uint8, device=dev)
        self.use_popart = False
        self.use_proper_time_limits = False
        self.step = 0

    def step_out(self, t):
        """StepOut whose destinations are row t of this storage (the pointers the step kernel writes through)."""
        o = _lib.StepOut()
        o.image, o.direction = _lib.ptr(self.obs['image'][t + 1]), _lib.ptr(self.obs['direction'][t + 1])
        o.reward, o.flags = _lib.ptr(self.rewards[t]), _lib.ptr(self.flags[t])
        o.trunc_image = _lib.ptr(self.truncated_obs['image'][t + 1])
        o.trunc_direction = _lib.ptr(self.truncated_obs['direction'][t + 1])
        o.masks, o.bad_masks = _lib.ptr(self.masks[t + 1]), _lib.ptr(self.bad_masks[t + 1])
        o.cliffhanger_masks = _lib.ptr(self.cliffhanger_masks[t + 1])
        return o

    def compute_returns(self, next_value, use_gae, gamma, gae_lambda):
        if not use_gae:
            raise NotImplementedError('only GAE returns are part of the hot path')
        return compute_gae_returns(self, next_value, gamma, gae_lambda)

    def after_update(self):
        for k in self.obs:
            self.obs[k][0].copy_(self.obs[k][-1])
        self.masks[0].copy_(self.masks[-1])
        self.bad_masks[0].copy_(self.bad_masks[-1])
        self.cliffhanger_masks[0].copy_(self.cliffhanger_masks[-1])


def discounted_returns(rewards, masks, returns, gamma):
    """returns[t] = returns[t+1]*gamma*masks[t+1] + rewards[t] for t < T (algos/storage.py:276-279); returns[T] is the
    bootstrap value already stored by the caller."""
    for x in (rewards, masks, returns):
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise _lib.MgplrError('discounted_returns needs contiguous float32 CUDA tensors (no CPU fallback)')
    T, N = rewards.shape[0], rewards.shape[1]
    _lib.check(_lib.load().mgplr_discounted_returns(_lib.ptr(rewards), _lib.ptr(masks), _lib.ptr(returns), T, N, float(gamma),
                                                    torch.cuda.current_stream(rewards.device).cuda_stream),
               'mgplr_discounted_returns')
    return returns


def batched_value_loss(returns, value_preds, signed=False, positive_only=False, power=1, clipped=True):
    """[N,1] mean episodic value loss per actor (algos/storage.py:290-327) from returns / value_preds [T+1,N,1]."""
    for x in (returns, value_preds):
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise _lib.MgplrError('batched_value_loss needs contiguous float32 CUDA tensors (no CPU fallback)')
    T, N = returns.shape[0] - 1, returns.shape[1]
    out = torch.empty(N, 1, dtype=torch.float32, device=returns.device)
    mode = 1 if signed else (2 if positive_only else 0)
    _lib.check(_lib.load().mgplr_batched_value_loss(_lib.ptr(returns), _lib.ptr(value_preds), T, N, mode, int(power),
                                                    int(bool(clipped)), _lib.ptr(out),
                                                    torch.cuda.current_stream(returns.device).cuda_stream),
               'mgplr_batched_value_loss')
    return out


def _flatten_helper(T, N, _tensor):
    if isinstance(_tensor, dict):
        return {k: _tensor[k].view(T * N, *_tensor[k].size()[2:]) for k in _tensor.keys()}
    return _tensor.view(T * N, *_tensor.size()[2:])


class RolloutStorage(object):
    """Drop-in for algos.storage.RolloutStorage (algos/storage.py:41-560) with every buffer in HBM.

    Same constructor arguments, attribute names, shapes and dtypes (`:62-112`), so `make_agent` / `ACAgent` /
    `PPO.update` / `LevelSampler.update_with_rollouts` use it unchanged.  `insert` keeps the reference's copy semantics;
    the fused path instead hands `step_out(t)` to the step kernel, which writes obs / rewards / masks in place.
    Reductions over the rollout (GAE, discounted returns, batched value loss) are kernels; the truncated-value path
    evaluates the critic ONCE on all truncated observations instead of once per process (`:208-231`)."""

    def __init__(self, model, num_steps, num_processes, observation_space, action_space, recurrent_hidden_state_size,
                 recurrent_arch='rnn', use_proper_time_limits=False, use_popart=False, device='cuda'):
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise _lib.MgplrError('RolloutStorage lives in HBM: device must be a CUDA device (no CPU fallback)')
        self.device = dev
        self.model = model
        self.num_processes = num_processes
        self.recurrent_arch = recurrent_arch
        self.recurrent_hidden_state_size = recurrent_hidden_state_size
        self.is_lstm = recurrent_arch == 'lstm'
        rnn_size = 2 * recurrent_hidden_state_size if self.is_lstm else recurrent_hidden_state_size
        self.use_proper_time_limits = use_proper_time_limits
        self.use_popart = use_popart
        T, N = num_steps, num_processes
        z = lambda *shape, **kw: torch.zeros(*shape, device=dev, **kw)  # noqa: E731
        self.truncated_obs = None
        if isinstance(observation_space, dict):
            self.is_dict_obs = True
            self.obs = {k: z(T + 1, N, *observation_space[k].shape) for k in observation_space}
            if use_proper_time_limits:
                self.truncated_obs = {k: z(T + 1, N, *observation_space[k].shape) for k in observation_space}
        else:
            self.is_dict_obs = False
            self.obs = z(T + 1, N, *observation_space.shape)
            if use_proper_time_limits:
                self.truncated_obs = torch.zeros_like(self.obs)
        self.recurrent_hidden_states = z(T + 1, N, rnn_size)
        self.rewards = z(T, N, 1)
        self.value_preds = z(T + 1, N, 1)
        self.returns = z(T + 1, N, 1)
        self.action_log_probs = z(T, N, 1)
        if action_space.__class__.__name__ == 'Discrete':
            action_shape = 1
            self.action_log_dist = z(T, N, action_space.n)
        else:
            action_shape = action_space.shape[0]
            self.action_log_dist = z(T, N, 1)
        self.actions = z(T, N, action_shape)
        if action_space.__class__.__name__ == 'Discrete':
            self.actions = self.actions.long()
        self.masks = torch.ones(T + 1, N, 1, device=dev)
        self.bad_masks = torch.ones(T + 1, N, 1, device=dev)
        self.cliffhanger_masks = torch.ones(T + 1, N, 1, device=dev)
        self.truncated_value_preds = torch.zeros_like(self.value_preds) if use_proper_time_limits else None
        self.denorm_value_preds = None
        self.level_seeds = z(T, N, 1, dtype=torch.int)
        self.flags = z(T, N, dtype=torch.uint8)  # step-kernel flags (not in the reference; used by the fused path)
        self.num_steps = num_steps
        self.step = 0

    def to(self, device):
        if torch.device(device).type != 'cuda':
            raise _lib.MgplrError('RolloutStorage lives in HBM (no CPU fallback)')
        return self

    # ---- reference bookkeeping (algos/storage.py:142-206)
    def get_obs(self, idx):
        if self.is_dict_obs:
            return {k: self.obs[k][idx] for k in self.obs.keys()}
        return self.obs[idx]

    def _obs_buffers(self):
        return list(self.obs.values()) if self.is_dict_obs else [self.obs]

    def copy_obs_to_index(self, obs, index):
        src = [obs[k] for k in self.obs] if self.is_dict_obs else [obs]
        for buf, o in zip(self._obs_buffers(), src):
            buf[index].copy_(o)

    def insert(self, obs, recurrent_hidden_states, actions, action_log_probs, action_log_dist, value_preds, rewards, masks,
               bad_masks, level_seeds=None, cliffhanger_masks=None):
        if len(rewards.shape) == 3:
            rewards = rewards.squeeze(2)
        t = self.step
        self.copy_obs_to_index(obs, t + 1)
        if self.is_lstm:
            H = self.recurrent_hidden_state_size
            self.recurrent_hidden_states[t + 1, :, :H].copy_(recurrent_hidden_states[0])
            self.recurrent_hidden_states[t + 1, :, H:].copy_(recurrent_hidden_states[1])
        else:
            self.recurrent_hidden_states[t + 1].copy_(recurrent_hidden_states)
        self.actions[t].copy_(actions)
        self.action_log_probs[t].copy_(action_log_probs)
        self.action_log_dist[t].copy_(action_log_dist)
        self.value_preds[t].copy_(value_preds)
        self.rewards[t].copy_(rewards)
        self.masks[t + 1].copy_(masks)
        self.bad_masks[t + 1].copy_(bad_masks)
        if cliffhanger_masks is not None:
            self.cliffhanger_masks[t + 1].copy_(cliffhanger_masks)
        if level_seeds is not None:
            self.level_seeds[t].copy_(level_seeds)
        self.step = (self.step + 1) % self.num_steps

    def step_out(self, t):
        """StepOut whose destinations are row t of this storage: the step kernel then does the obs / reward / mask part
        of insert() itself (the caller still stores actions, values and log-probs and advances `step`)."""
        o = _lib.StepOut()
        o.image, o.direction = _lib.ptr(self.obs['image'][t + 1]), _lib.ptr(self.obs['direction'][t + 1])
        o.reward, o.flags = _lib.ptr(self.rewards[t]), _lib.ptr(self.flags[t])
        if self.truncated_obs is not None:
            o.trunc_image = _lib.ptr(self.truncated_obs['image'][t + 1])
            o.trunc_direction = _lib.ptr(self.truncated_obs['direction'][t + 1])
        o.masks, o.bad_masks = _lib.ptr(self.masks[t + 1]), _lib.ptr(self.bad_masks[t + 1])
        o.cliffhanger_masks = _lib.ptr(self.cliffhanger_masks[t + 1])
        return o

    def insert_truncated_obs(self, obs, index):
        def as_t(a):
            return a.to(self.device).float() if torch.is_tensor(a) else torch.as_tensor(a, dtype=torch.float32, device=self.device)
        if self.is_dict_obs:
            [self.truncated_obs[k][self.step + 1][index].copy_(as_t(obs[k])) for k in self.truncated_obs.keys()]
        else:
            self.truncated_obs[self.step + 1][index].copy_(as_t(obs))

    def after_update(self):
        """The last slot of every [T+1] buffer becomes slot 0 of the next rollout (algos/storage.py:195-204)."""
        for buf in self._obs_buffers() + [self.recurrent_hidden_states, self.masks, self.bad_masks, self.cliffhanger_masks]:
            buf[0].copy_(buf[-1])

    def replace_final_return(self, returns):
        self.rewards[-1] = returns

    # ---- returns (algos/storage.py:208-288)
    def _compute_truncated_value_preds(self):
        """value_preds with the critic's value of the truncated observation wherever bad_masks == 0 (`:208-231`).
        One batched critic call over all (step, process) pairs, in the reference's process-major order.
        Reference quirk kept: `steps = (...).nonzero().squeeze()` is 0-dimensional for a process with exactly ONE
        truncated step, and `len(steps.shape) == 0` then skips that process (`:213-215`) -- only processes with two or
        more truncated steps get truncated values."""
        self.truncated_value_preds.copy_(self.value_preds)
        with torch.no_grad():
            bad = self.bad_masks[:, :, 0].t() == 0  # [process, step]
            bad = bad & (bad.sum(1, keepdim=True) >= 2)
            idx = bad.nonzero()  # rows (process, step), process-major
            if idx.shape[0]:
                proc, steps = idx[:, 0], idx[:, 1]
                if self.is_dict_obs:
                    obs = {k: self.truncated_obs[k][steps, proc] for k in self.truncated_obs.keys()}
                else:
                    obs = self.truncated_obs[steps, proc]
                rnn_hxs = self.recurrent_hidden_states[steps, proc]
                if self.is_lstm:
                    rnn_hxs = self._split_batched_lstm_recurrent_hidden_states(rnn_hxs)
                masks = torch.ones((idx.shape[0], 1), device=self.device)
                self.truncated_value_preds[steps, proc] = self.model.get_value(obs, rnn_hxs, masks)
        return self.truncated_value_preds

    def _bootstrap_values(self, next_value):
        """-> (values the returns bootstrap from, their PopArt-denormalised form or None): the truncated-value buffer under
        use_proper_time_limits, else value_preds (algos/storage.py:238-249,260-270)."""
        self.value_preds[-1] = next_value
        v = self._compute_truncated_value_preds() if self.use_proper_time_limits else self.value_preds
        denorm = None
        if self.use_popart:
            denorm = self.denorm_value_preds = self.model.popart.denormalize(v)
        return v, denorm

    def _value_preds_for_returns(self, next_value):
        v, denorm = self._bootstrap_values(next_value)
        return (v if denorm is None else denorm).contiguous()

    def compute_gae_returns(self, returns_buffer, next_value, gamma, gae_lambda):
        gae_returns(self.rewards, self._value_preds_for_returns(next_value), self.masks, returns_buffer, gamma, gae_lambda)

    def compute_discounted_returns(self, returns_buffer, next_value, gamma):
        v, _ = self._bootstrap_values(next_value)
        self.returns[-1] = v[-1]  # (the reference bootstraps from the un-denormalised buffer, `:272`)
        discounted_returns(self.rewards, self.masks, returns_buffer, gamma)

    def compute_returns(self, next_value, use_gae, gamma, gae_lambda):
        if use_gae:
            self.compute_gae_returns(self.returns, next_value, gamma, gae_lambda)
        else:
            self.compute_discounted_returns(self.returns, next_value, gamma)

    def get_batched_value_loss(self, signed=False, positive_only=False, power=1, clipped=True, batched=True):
        value_preds = self.denorm_value_preds if self.use_popart else self.value_preds
        batch_td = batched_value_loss(self.returns, value_preds.contiguous(), signed, positive_only, power, clipped)
        return batch_td if batched else batch_td.mean().item()

    def get_action_traj(self, as_string=False):
        if as_string:  # the action-string level format (`:371-378`)
            a = self.actions[:, :, 0].t().cpu().numpy()
            return [' '.join(str(int(x)) for x in row) for row in a]
        return self.actions.squeeze(-1)

    def get_batched_action_complexity(self):
        raise NotImplementedError('Lempel-Ziv action complexity is logging only (third-party lempel_ziv_complexity)')

    get_action_complexity = get_batched_action_complexity

    # ---- minibatch generators (algos/storage.py:392-560): index plumbing on device tensors
    def _split_batched_lstm_recurrent_hidden_states(self, hxs):
        H = self.recurrent_hidden_state_size
        return (hxs[:, :H], hxs[:, H:])

    def get_recurrent_hidden_state(self, step):
        if self.is_lstm:
            return self._split_batched_lstm_recurrent_hidden_states(self.recurrent_hidden_states[step, :].squeeze(0))
        return self.recurrent_hidden_states[step]

    def feed_forward_generator(self, advantages, num_mini_batch=None, mini_batch_size=None):
        from torch.utils.data.sampler import BatchSampler, SubsetRandomSampler
        num_steps, num_processes = self.rewards.size()[0:2]
        batch_size = num_processes * num_steps
        if mini_batch_size is None:
            assert batch_size >= num_mini_batch
            mini_batch_size = batch_size // num_mini_batch
        sampler = BatchSampler(SubsetRandomSampler(range(batch_size)), mini_batch_size, drop_last=False)
        for indices in sampler:
            indices = torch.as_tensor(indices, device=self.device)
            if self.is_dict_obs:
                obs_batch = {k: self.obs[k][:-1].view(-1, *self.obs[k].size()[2:])[indices] for k in self.obs.keys()}
            else:
                obs_batch = self.obs[:-1].view(-1, *self.obs.size()[2:])[indices]
            rnn = self.recurrent_hidden_states[:-1].view(-1, self.recurrent_hidden_states.size(-1))[indices]
            actions_batch = self.actions.view(-1, self.actions.size(-1))[indices]
            value_preds_batch = self.value_preds[:-1].view(-1, 1)[indices]
            return_batch = self.returns[:-1].view(-1, 1)[indices]
            masks_batch = self.masks[:-1].view(-1, 1)[indices]
            old_action_log_probs_batch = self.action_log_probs.view(-1, 1)[indices]
            adv_targ = None if advantages is None else advantages.view(-1, 1)[indices]
            if self.is_lstm:
                rnn = self._split_batched_lstm_recurrent_hidden_states(rnn)
            yield obs_batch, rnn, actions_batch, value_preds_batch, return_batch, masks_batch, old_action_log_probs_batch, adv_targ

    def recurrent_generator(self, advantages, num_mini_batch):
        num_processes = self.rewards.size(1)
        assert num_processes >= num_mini_batch
        num_envs_per_batch = num_processes // num_mini_batch
        perm = torch.randperm(num_processes)
        T = self.num_steps
        for start_ind in range(0, num_processes, num_envs_per_batch):
            ind = perm[start_ind:start_ind + num_envs_per_batch].to(self.device)
            N = ind.numel()
            if self.is_dict_obs:
                obs_batch = {k: _flatten_helper(T, N, self.obs[k][:-1, ind]) for k in self.obs.keys()}
            else:
                obs_batch = _flatten_helper(T, N, self.obs[:-1, ind])
            rnn = self.recurrent_hidden_states[0, ind].view(N, -1)
            actions_batch = _flatten_helper(T, N, self.actions[:, ind])
            value_preds_batch = _flatten_helper(T, N, self.value_preds[:-1, ind])
            return_batch = _flatten_helper(T, N, self.returns[:-1, ind])
            masks_batch = _flatten_helper(T, N, self.masks[:-1, ind])
            old_action_log_probs_batch = _flatten_helper(T, N, self.action_log_probs[:, ind])
            adv_targ = None if advantages is None else _flatten_helper(T, N, advantages[:, ind])
            if self.is_lstm:
                rnn = self._split_batched_lstm_recurrent_hidden_states(rnn)
            yield obs_batch, rnn, actions_batch, value_preds_batch, return_batch, masks_batch, old_action_log_probs_batch, adv_targ
