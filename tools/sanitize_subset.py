#!/usr/bin/env python
"""Every kernel of libmgplr.so once or a few times at a small batch, for compute-sanitizer:

  compute-sanitizer --tool memcheck  python tools/sanitize_subset.py
  compute-sanitizer --tool racecheck python tools/sanitize_subset.py --racecheck
  compute-sanitizer --tool synccheck python tools/sanitize_subset.py

2 048 envs (64 warp tiles: several tiles per warp at a 1-SM-sized grid is not needed, the persistent grid adapts), the DR
speculation forced on (MGPLR_RR_SPEC=2) so that candidate records, job lists, regeneration jobs (both the warp-per-job
and the lane-parallel storm variant) and commits all run, a time limit of 24 steps so that truncations, storms and
cliffhangers happen within a few dozen steps.  Results are checked against nothing here (the parity tests do that); the
point is the sanitizer's report.  racecheck only sees shared memory: the intended global-memory race of the DR jobs
 is outside its scope and is covered by the bit-exactness tests.
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('MGPLR_RR_SPEC', '2')

import numpy as np  # noqa: E402
import torch  # noqa: E402

from dcd_isaac_b200 import _lib  # noqa: E402
from dcd_isaac_b200._lib import ptr, check  # noqa: E402
from dcd_isaac_b200.storage import DeviceRolloutStorage, RolloutStorage  # noqa: E402
from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv, CudaMazeVecEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=2048)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--racecheck', action='store_true', help='smaller run: racecheck is ~100x slower than memcheck')
    a = ap.parse_args()
    N, T = (512, 40) if a.racecheck else (a.envs, a.steps)
    L = _lib.load()
    rs = np.random.RandomState(0)
    for W, see in ((15, True), (25, False)):
        v = CudaAdversarialVecEnv('MultiGrid-GoalLastFewerBlocksAdversarial-v0', N, size=W, see_through_walls=see, max_episode_steps=24,
                                  full_obs=True)
        v.set_seed(list(range(N)))
        v.seed(7, 3)
        # adversary build: reset + step_adversary (fused small-batch path) then the separate kernels via a large-batch twin below
        v.reset()
        S = v.adversary_max_steps
        for s in range(S):
            v.step_adversary(torch.from_numpy(rs.randint(0, v.adversary_action_dim, size=(N, 1))).cuda())
        v.reset_agent()
        enc = v.get_encodings_device()
        v.get_num_blocks()
        v.reset_random()
        v.reset_to_level_batch([e for e in enc.cpu().numpy()])
        v.reset_to_level(enc[0].cpu().numpy(), 5)
        v.mutate_level(5)
        st = DeviceRolloutStorage(T, N)
        acts = torch.from_numpy(rs.randint(0, 7, size=(T, N))).cuda()
        acts[torch.rand(T, N, device='cuda') < 0.5] = 2
        for rr in (False, True):
            v.reset_agent()
            for t in range(T):
                v.step_env_device(acts[t].contiguous(), st.step_out(t), reset_random=rr, last_step=3 if t == T - 1 else 0)
            v.rollout_device(acts.to(torch.uint8).contiguous(), st.step_out(0), reset_random=rr, last_step=3)
            for t in range(6):   # host-driven step: pinned zero-copy actions (u8 and i64), pageable fallback
                v.step_env(acts[t].cpu().view(N, 1), reset_random=rr)
            check(L.mgplr_step_env_u8(v.h, ptr(acts[0].to(torch.uint8).contiguous()), int(rr), None, 0, C.byref(st.step_out(0)),
                                      torch.cuda.current_stream().cuda_stream))
            v.get_num_blocks(); v.get_agent_state(); v.peek_rng(3)
        # PLR math on the rollout just written
        st.value_preds.copy_(torch.rand_like(st.value_preds))
        st.level_seeds.copy_(torch.randint(1, 50, st.level_seeds.shape, device='cuda', dtype=torch.int32))
        st.compute_returns(torch.rand(N, 1, device='cuda'), True, 0.995, 0.95)
        from dcd_isaac_b200.level_sampler import LevelSampler
        for strat in ('positive_value_loss', 'grounded_signed_value_loss', 'min_margin', 'one_step_td_error'):
            smp = LevelSampler([], None, None, num_actors=N, strategy=strat, score_transform='rank', temperature=0.3,
                               staleness_coef=0.3, sample_full_distribution=True, seed_buffer_size=64, rho=0.1, replay_prob=0.5)
            smp.observe_external_unseen_sample(list(range(1, 50)))
            smp.update_with_rollouts(st)
            smp.after_update()
            if smp.working_seed_buffer_size:
                smp.sample_weights()
                smp.sample_replay_levels(8)
        from dcd_isaac_b200 import storage as S_
        S_.batched_value_loss(st.returns, st.value_preds)
        S_.discounted_returns(st.rewards, st.masks, st.returns, 0.99)
        v.get_images(index=[0, 1])
        v.close()
    # large-batch code paths of the adversary kernels (separate state / image kernels) and the storm regeneration
    v = CudaAdversarialVecEnv('MultiGrid-GoalLastAdversarial-v0', 20000 if not a.racecheck else 16416, max_episode_steps=8)
    v.set_seed(list(range(v.num_envs)))
    v.reset()
    for s in range(3):
        v.step_adversary(torch.from_numpy(rs.randint(0, 169, size=(v.num_envs, 1))).cuda())
    v.reset_random()
    v.reset_agent()
    st = DeviceRolloutStorage(20, v.num_envs)
    for t in range(20):
        v.step_env_device(torch.randint(0, 7, (v.num_envs,), device='cuda'), st.step_out(t), reset_random=True)
    v.close()
    m = CudaMazeVecEnv('MultiGrid-Labyrinth-v0', 64)
    m.reset()
    for t in range(30):
        m.step(torch.randint(0, 7, (64, 1)))
    m.close()
    torch.cuda.synchronize()
    print('sanitize_subset done')


if __name__ == '__main__':
    main()
