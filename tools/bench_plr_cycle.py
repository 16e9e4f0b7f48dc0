#!/usr/bin/env python
"""PLR-side wall times at the reference's own scale (config 1/2: 32 actors, T=256, buffer 4000), the quantities
BASELINE.md section 3 lists for the reference: sample_replay_level x32, update_with_rollouts, after_update,
sample_weights.  Host wall clock (these are host-driven calls), best of 5.  One JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from dcd_isaac_b200.level_sampler import LevelSampler
from dcd_isaac_b200.storage import DeviceRolloutStorage


def best(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    b = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        b = min(b, time.perf_counter() - t)
    return b * 1e3


def main():
    A, T, NB = 32, 256, 4000
    np.random.seed(0)
    s = LevelSampler([], None, None, num_actors=A, strategy='positive_value_loss', replay_schedule='fixed', score_transform='rank',
                     temperature=0.3, rho=0.5, replay_prob=0.5, staleness_coef=0.3, sample_full_distribution=True,
                     seed_buffer_size=NB)
    # a full buffer with random scores (SURVEY.md 8d synthetic PLR state)
    s.seeds[:] = np.arange(1, NB + 1)
    s.seed2index = {int(k): i for i, k in enumerate(s.seeds)}
    s.working_seed_set = set(int(k) for k in s.seeds)
    s.working_seed_buffer_size = NB
    s.seed_scores[:] = np.random.rand(NB)
    s.unseen_seed_weights[:] = 0
    s.seed_staleness[:] = np.floor(np.random.rand(NB) * 50)
    st = DeviceRolloutStorage(T, A)
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    st.rewards.copy_((torch.rand(T, A, 1, device='cuda', generator=g) < 0.01).float())
    st.value_preds.copy_(torch.rand(T + 1, A, 1, device='cuda', generator=g))
    m = (torch.rand(T + 1, A, 1, device='cuda', generator=g) > 0.02).float(); m[-1] = 0
    st.masks.copy_(m)
    st.level_seeds.copy_(torch.randint(1, NB + 1, (T, A, 1), device='cuda', generator=g, dtype=torch.int32))
    st.compute_returns(torch.zeros(A, 1, device='cuda'), True, 0.995, 0.95)
    out = {
        'sample_replay_level_x32_ms': best(lambda: [s.sample_replay_level() for _ in range(32)]),
        'sample_replay_levels_32_one_launch_ms': best(lambda: s.sample_replay_levels(32)),
        'update_with_rollouts_ms': best(lambda: s.update_with_rollouts(st)),
        'after_update_ms': best(lambda: s.after_update()),
        'sample_weights_ms': best(lambda: s.sample_weights()),
        'compute_returns_gae_ms': best(lambda: st.compute_returns(torch.zeros(A, 1, device='cuda'), True, 0.995, 0.95)),
        'reference_ms (BASELINE.md section 3, 8-core CPU)': {'sample_replay_level_x32': 20.7, 'update_with_rollouts': 17.0,
                                                              'after_update': 44.0},
        'config': '32 actors, T=256, buffer 4000, rank T=0.3, staleness 0.3',
    }
    print(json.dumps(out))


if __name__ == '__main__':
    main()
