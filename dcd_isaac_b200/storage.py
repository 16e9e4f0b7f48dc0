"""Rollout-storage pieces of the hot path (algos/storage.py): buffer layouts the step kernel writes into and
RolloutStorage.compute_gae_returns on the device.

`compute_gae_returns(storage, next_value, gamma, gae_lambda)` can be called on the reference's own
RolloutStorage object (it only touches .rewards/.value_preds/.masks/.returns and, like the reference,
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
