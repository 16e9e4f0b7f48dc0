"""A scripted stand-in for algos/agent.py:ACAgent (TEST INFRASTRUCTURE ONLY).

Runner-level fixtures (oracle/gen_golden_runner.py) EXECUTE the reference's envs/runners/adversarial_runner.py; the
only thing that cannot be reproduced bit-for-bit on another machine is a neural network's sampling, so the policy is
replaced by this object: actions, values and log-distributions come from a seeded numpy stream, one draw block per
`act()` call, and the critic used by the truncated-value path (algos/storage.py:208-231) is a fixed linear function
of the observation.  The same class drives the reference runner over the drop-in objects in the GPU tests, so both
sides consume identical streams.  It exposes exactly what the runner touches: `storage`, `act`, `get_value`,
`process_action`, `insert`, `update`, `train`, `eval`, `random`, `algo.actor_critic`.
"""
import numpy as np
import torch

SNAP_KEYS = ('rewards', 'masks', 'bad_masks', 'cliffhanger_masks', 'level_seeds', 'actions', 'value_preds', 'returns',
             'truncated_value_preds')


class _Critic(object):
    """`storage.model`: value = 0.25 * mean(image) + 0.125 * direction (deterministic in the observation)."""

    def get_value(self, obs, rnn_hxs, masks):
        img = obs['image'].float()
        v = img.reshape(img.shape[0], -1).mean(dim=1, keepdim=True) * 0.25
        if 'direction' in obs:
            v = v + obs['direction'].float().reshape(img.shape[0], -1)[:, :1] * 0.125
        return v

    def state_dict(self):
        return {}


class _Algo(object):
    def __init__(self):
        self.actor_critic = _Critic()


class ScriptedAgent(object):
    def __init__(self, storage, num_actions, num_processes, seed, forward_bias=0.0):
        self.storage = storage
        self.storage.model = _Critic()
        self.algo = _Algo()
        self.n_act = int(num_actions)
        self.N = int(num_processes)
        self.rs = np.random.RandomState(seed)
        self.forward_bias = forward_bias
        self.snapshots = []
        self.is_recurrent = False

    @property
    def device(self):
        return self.storage.rewards.device

    def act(self, obs, rnn_hxs, masks):
        N, A = self.N, self.n_act
        a = self.rs.randint(0, A, size=N)
        if self.forward_bias > 0:
            a[self.rs.rand(N) < self.forward_bias] = 2  # MiniGridEnv.Actions.forward
        value = self.rs.rand(N).astype(np.float32)
        logits = self.rs.randn(N, A).astype(np.float32)
        dev = self.device
        log_dist = torch.log_softmax(torch.from_numpy(logits), dim=-1).to(dev)
        return (torch.from_numpy(value).view(N, 1).to(dev), torch.from_numpy(a.astype(np.int64)).view(N, 1).to(dev), log_dist,
                torch.zeros(N, 1, device=dev))

    def get_value(self, obs, rnn_hxs, masks):
        return torch.from_numpy(self.rs.rand(self.N).astype(np.float32)).view(self.N, 1).to(self.device)

    def process_action(self, action):
        return action

    def insert(self, *args, **kwargs):
        return self.storage.insert(*args, **kwargs)

    def snapshot(self):
        st = self.storage
        snap = {'obs_' + k: v.detach().cpu().numpy().copy() for k, v in st.obs.items()}
        for k in SNAP_KEYS:
            v = getattr(st, k, None)
            if v is not None:
                snap[k] = v.detach().cpu().numpy().copy()
        if getattr(st, 'truncated_obs', None) is not None:
            for k, v in st.truncated_obs.items():
                snap['truncated_obs_' + k] = v.detach().cpu().numpy().copy()
        return snap

    def update(self, discard_grad=False):
        self.snapshots.append(self.snapshot())
        self.storage.after_update()
        return 0.0, 0.0, 0.0, {}

    def train(self):
        pass

    def eval(self):
        pass

    def random(self):
        pass

    def to(self, device):
        self.storage.to(device)
        return self
