#!/usr/bin/env python
"""The drop-in API at the reference's own scale: `venv.step_env(action)` called from Python exactly like
envs/runners/adversarial_runner.py:512-517 does (CPU int64 action tensor in; obs dict of CUDA tensors, reward, numpy done,
list of info dicts out), for the 32 processes of the shipped configs and a few larger batches; plus the adversary build
(`reset` + S `step_adversary`) and a DR reset.  One JSON object per line; BASELINE.md section 3 has the reference's numbers
(4.7-21 ms per vector step of 32 envs, 14-17 ms per 52-step build per env).

  python tools/bench_dropin.py [--steps 2000]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=2000)
    a = ap.parse_args()
    for N in (32, 256, 4096):
        v = CudaAdversarialVecEnv('MultiGrid-GoalLastFewerBlocksAdversarial-v0', N)
        v.set_seed(list(range(N)))
        v.reset_random()
        v.reset_agent()
        rs = np.random.RandomState(0)
        acts = torch.from_numpy(rs.randint(0, 7, size=(a.steps, N, 1)).astype(np.int64))
        for mode in ('cpu action tensor', 'cuda action tensor'):
            A = acts if mode.startswith('cpu') else acts.cuda()
            for t in range(50):
                v.step_env(A[t])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dones = 0
            for t in range(a.steps):
                obs, rew, done, infos = v.step_env(A[t])
                dones += int(done.sum())
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(json.dumps({'op': 'venv.step_env from Python (%s)' % mode, 'envs': N, 'us_per_call': dt / a.steps * 1e6,
                              'env_steps_per_s': N * a.steps / dt, 'episodes': dones}), flush=True)
        # adversary build: reset + S step_adversary with random locations (adversarial_runner.py:481,515)
        S = v.adversary_max_steps
        locs = torch.from_numpy(rs.randint(0, v.adversary_action_dim, size=(S, N, 1)).astype(np.int64))
        for _ in range(2):
            v.reset()
            for s in range(S):
                v.step_adversary(locs[s])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v.reset()
        for s in range(S):
            v.step_adversary(locs[s])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({'op': 'adversary build from Python (reset + %d step_adversary)' % S, 'envs': N, 'ms_per_build': dt * 1e3,
                          'levels_per_s': N / dt}), flush=True)
        t0 = time.perf_counter()
        for _ in range(20):
            v.reset_random()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        print(json.dumps({'op': 'venv.reset_random from Python', 'envs': N, 'us_per_call': dt * 1e6}), flush=True)
        v.close()


if __name__ == '__main__':
    main()
