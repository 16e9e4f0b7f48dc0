#!/usr/bin/env python
"""Episode-score pass (count + scans + score kernel) timed by CUDA events for a few actor counts; run once per
MGPLR_SCORE_SPLIT setting (0 = one thread per actor, 1 = time-split, unset = by size) to compare the two kernels.

  MGPLR_SCORE_SPLIT=1 python tools/bench_scores.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from dcd_isaac_b200 import _lib
from dcd_isaac_b200._lib import check, ptr


def main():
    L = _lib.load()
    T = 256
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    for N in (32, 4096, 32768, 131072, 262144, 524288):
        g = torch.Generator(device='cuda').manual_seed(1)
        r = torch.rand(T, N, device='cuda', generator=g)
        val = torch.rand(T + 1, N, device='cuda', generator=g)
        ret = torch.rand(T + 1, N, device='cuda', generator=g)
        m = (torch.rand(T + 1, N, device='cuda', generator=g) > 0.02).float()
        m[-1] = 0
        cl = torch.ones_like(m)
        seeds = torch.randint(1, 4000, (T, N), device='cuda', dtype=torch.int32)
        ep = torch.zeros(N * (T + 1), 10, dtype=torch.int32, device='cuda')
        nep = torch.zeros(1, dtype=torch.int32, device='cuda')
        fn = lambda: check(L.mgplr_plr_episode_scores(ptr(m), ptr(cl), ptr(ret), ptr(val), ptr(r), ptr(seeds), T, N, 0, ptr(ep),  # noqa: E731
                                                      N * (T + 1), ptr(nep), st()))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print(json.dumps({'op': 'episode scores (positive_value_loss, T=256)', 'actors': N, 'us': best * 1e3,
                          'split_knob': os.environ.get('MGPLR_SCORE_SPLIT', 'auto'), 'episodes': int(nep.item()),
                          'checksum': int(ep[:int(nep.item())].to(torch.int64).sum().item()),
                          'gbs_at_20B': 20.0 * N * T / (best * 1e-3) / 1e9}))
        del r, val, ret, m, cl, seeds, ep


if __name__ == '__main__':
    main()
