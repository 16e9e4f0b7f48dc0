#!/usr/bin/env python
"""Secondary measurements: every non-step row of SURVEY.md 8(a) timed on the device (CUDA events, warm-up 3, best of 5)
with its algorithmic bytes (SURVEY.md 8d "other units") against the measured HBM peak.  One JSON object per line.

  python tools/bench_levelops.py [--envs 4096] [--size 15]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from dcd_isaac_b200 import _lib
from dcd_isaac_b200._lib import check, ptr
from dcd_isaac_b200.level_sampler import LevelSampler
from dcd_isaac_b200.storage import gae_returns
from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv


def timeit(fn, reps=int(os.environ.get('LEVELOPS_REPS', 5)), warm=int(os.environ.get('LEVELOPS_WARM', 3))):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=4096)
    ap.add_argument('--size', type=int, default=15)
    a = ap.parse_args()
    N, W = a.envs, a.size
    peak = 6545.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    L = _lib.load()
    v = CudaAdversarialVecEnv('MultiGrid-GoalLastAdversarial-v0', N, size=W)
    v.set_seed(list(range(N)))
    st = lambda: torch.cuda.current_stream().cuda_stream
    out = []

    def rec(name, secs, units, bytes_per_unit, note=''):
        gbs = units * bytes_per_unit / secs / 1e9
        out.append({'op': name, 'envs': N, 'size': W, 'seconds': secs, 'units_per_s': units / secs, 'unit_bytes': bytes_per_unit,
                    'achieved_gbs': gbs, 'frac_of_measured_hbm': gbs / peak, 'note': note})

    # adversary build: reset + S_adv step_adversary with the [N,3,W,W] f32 observation written every step
    S = v.adversary_max_steps
    img = torch.empty(N, 3, W, W, device='cuda')
    ts = torch.empty(N, 1, device='cuda')
    dn = torch.empty(N, dtype=torch.uint8, device='cuda')
    locs = torch.randint(0, (W - 2) ** 2, (S, N), device='cuda')

    def build():
        check(L.mgplr_reset(v.h, ptr(img), ptr(ts), st()))
        for k in range(S):
            check(L.mgplr_step_adversary(v.h, ptr(locs[k]), ptr(img), ptr(ts), ptr(dn), st()))
    t = timeit(build)
    rec('adversary build (reset + %d step_adversary, image written each step)' % S, t, N * (S + 1), 3 * W * W * 4 + 16,
        'PAIRED adversary rollout; per step: 8 B action + state + 3*W*W*4 B fp32 image')
    t = timeit(lambda: check(L.mgplr_reset_random(v.h, None, None, st())))
    rec('reset_random', t, N, 4 * W + 360, 'latency-bound: ~80 serial MT19937 draws + flood fill per env')
    obs = v._new_obs()
    o = v._out(obs)
    t = timeit(lambda: check(L.mgplr_reset_agent(v.h, C.byref(o), st())))
    rec('reset_agent (+obs)', t, N, 4 * W + 16 + 16 + 304)
    enc = v.get_encodings_device()
    t = timeit(lambda: check(L.mgplr_get_encodings(v.h, ptr(enc), st())))
    rec('get_encodings', t, N, 3 * W * W + 4 * W + 16)
    t = timeit(lambda: check(L.mgplr_reset_to_encoding(v.h, ptr(enc), None, N, C.byref(o), st())))
    rec('reset_to_level_batch (bytes)', t, N, 3 * W * W + W * W + 360, 'SURVEY 8d: 3W^2 read + W^2 + 360')
    acts = locs.t().contiguous().to(torch.int32)
    t = timeit(lambda: check(L.mgplr_reset_to_actions(v.h, ptr(acts), S, None, N, C.byref(o), st())))
    rec('reset_to_level_batch (action strings, %d steps)' % S, t, N, 4 * S + 4 * W + 360)
    K = 5
    ml = torch.randint(0, (W - 2) ** 2, (N, K), device='cuda', dtype=torch.int32)
    mo_ = torch.randint(0, 4, (N, K), device='cuda', dtype=torch.int32)
    mn = torch.full((N,), K, device='cuda', dtype=torch.int32)
    need = torch.zeros(N, 2, dtype=torch.uint8, device='cuda')
    nfree = torch.zeros(N, 2, dtype=torch.int32, device='cuda')
    ch = torch.zeros(N, 2, dtype=torch.int32, device='cuda')

    def mutate():
        check(L.mgplr_mutate_edits(v.h, ptr(ml), ptr(mo_), ptr(mn), K, ptr(need), ptr(nfree), st()))
        check(L.mgplr_mutate_finalize(v.h, ptr(ch), C.byref(o), st()))
    t = timeit(mutate)
    rec('mutate_level (5 edits, two kernels)', t, N, 8 * K + 8 * W + 360)
    # rollout math at T=256
    T = 256
    r = torch.rand(T, N, 1, device='cuda')
    val = torch.rand(T + 1, N, 1, device='cuda')
    m = (torch.rand(T + 1, N, 1, device='cuda') > 0.02).float()
    m[-1] = 0
    ret = torch.zeros(T + 1, N, 1, device='cuda')
    t = timeit(lambda: gae_returns(r, val, m, ret, 0.995, 0.95))
    rec('compute_gae_returns (T=256)', t, N * T, 16, 'SURVEY 8d: 16 B per (t, env)')

    class RO(object):
        use_popart = False
    ro = RO()
    ro.rewards, ro.value_preds, ro.masks, ro.cliffhanger_masks, ro.returns = r, val, m, torch.ones_like(m), ret
    ro.level_seeds = torch.randint(1, 4000, (T, N, 1), device='cuda', dtype=torch.int32)
    s = LevelSampler([], None, None, num_actors=N, strategy='positive_value_loss', sample_full_distribution=True,
                     seed_buffer_size=4000, score_transform='rank', temperature=0.3, staleness_coef=0.3)
    ep = torch.zeros(N * (T + 1), 10, dtype=torch.int32, device='cuda')
    nep = torch.zeros(1, dtype=torch.int32, device='cuda')
    mm, cc, rr_, vv, rw, ls = (x[:, :, 0].contiguous() for x in (m, ro.cliffhanger_masks, ret, val, r, ro.level_seeds))
    t = timeit(lambda: check(L.mgplr_plr_episode_scores(ptr(mm), ptr(cc), ptr(rr_), ptr(vv), ptr(rw), ptr(ls), T, N, 0, ptr(ep),
                                                         N * (T + 1), ptr(nep), st())))
    rec('PLR episode scores (positive_value_loss, T=256)', t, N * T, 20, 'SURVEY 8d: 20 B per (t, env); 3 kernels + 2 scan kernels')
    nb = 4000
    sc = torch.rand(nb, dtype=torch.float64, device='cuda')
    stl = torch.floor(torch.rand(nb, dtype=torch.float64, device='cuda') * 50)
    un = (torch.rand(nb, device='cuda') < 0.1).double()
    wt = torch.zeros(nb, dtype=torch.float64, device='cuda')
    t = timeit(lambda: check(L.mgplr_plr_sample_weights(ptr(sc), ptr(stl), ptr(un), nb, 1, 0.3, 0.0, 0.3, 2, 1.0, None, ptr(wt), st())))
    rec('sample_weights (buffer 4000, rank + staleness)', t, 1, nb * 8 * 4, 'one CTA, fp64; latency-bound (bitonic sort of 4096 keys)')
    u = torch.rand(32, dtype=torch.float64, device='cuda')
    oi = torch.zeros(32, dtype=torch.int32, device='cuda')
    t = timeit(lambda: check(L.mgplr_plr_sample_replay(ptr(sc), ptr(stl), ptr(un), nb, 1, 0.3, 0.0, 0.3, 2, 1.0, None, ptr(u), 32, ptr(oi), st())))
    rec('32 sequential sample_replay_level draws (buffer 4000)', t, 32, nb * 8 * 3, 'reference: 20.7 ms on CPU (BASELINE.md)')
    for o_ in out:
        print(json.dumps(o_))
    v.close()


if __name__ == '__main__':
    main()
