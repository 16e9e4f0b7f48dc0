#!/usr/bin/env python
"""One full environment-side cycle of BASELINE.json configs[2] (PAIRED) and configs[3] (ACCEL), device-resident policies
stood in for by random action tensors (the networks are out of scope), timed with CUDA events after 2 warm-up cycles.

  PAIRED (adversarial_runner.py:475-520 order): venv.reset(); S x step_adversary(loc) with the [N,3,W,W] adversary
      observation written each step; protagonist rollout = reset_agent + T step_env + GAE; antagonist rollout = the same.
  ACCEL  (adversarial_runner.py:455-472,553-555,616-622): sample_replay_levels -> reset_to_level_batch (byte levels) ->
      T step_env -> GAE -> update_with_rollouts; mutate_level(5 edits) -> T step_env -> GAE -> update_with_rollouts ->
      level-store insert of the children's encodings.

One JSON object per line.   python tools/bench_configs.py [--envs 4096]
"""
import argparse
import json
import time
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from dcd_isaac_b200.level_sampler import LevelSampler
from dcd_isaac_b200.level_store import LevelStore
from dcd_isaac_b200.storage import DeviceRolloutStorage
from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def student_rollout(v, st, acts, last=3):
    """reset_agent -> obs[0]; T steps into rollout storage; GAE.  Device-only calls (capturable in a CUDA graph)."""
    import ctypes as C
    from dcd_isaac_b200._lib import check
    T = acts.shape[0]
    o0 = v._out({'image': st.obs['image'][0], 'direction': st.obs['direction'][0]})
    check(v.L.mgplr_reset_agent(v.h, C.byref(o0), v._stream()), 'mgplr_reset_agent')
    for t in range(T):
        v.step_env_device(acts[t], st.step_out(t), last_step=(last if t == T - 1 else 0))
    st.compute_returns(st.value_preds[-1].clone(), True, 0.995, 0.95)


def paired(N, T):
    v = CudaAdversarialVecEnv('MultiGrid-GoalLastAdversarial-v0', N)
    v.set_seed(list(range(N)))
    S = v.adversary_max_steps
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    locs = torch.randint(0, v.adversary_action_dim, (S, N, 1), device='cuda', generator=g)
    acts = [torch.randint(0, 7, (T, N), device='cuda', generator=g) for _ in range(2)]
    sts = [DeviceRolloutStorage(T, N) for _ in range(2)]
    for st in sts:
        st.value_preds.copy_(torch.rand(T + 1, N, 1, device='cuda', generator=g))

    graphs = [None, None]

    def cycle(use_graph=False):
        v.reset()
        for s in range(S):
            v.step_adversary(locs[s])
        for k in range(2):
            if use_graph:
                graphs[k].replay()
            else:
                student_rollout(v, sts[k], acts[k])
        v.get_passable(); v.get_shortest_path_length(); v.get_num_blocks()

    ms = timed(cycle)
    out = {'config': 'configs[2] PAIRED, MultiGrid-GoalLastAdversarial-v0, %d envs, adversary %d steps + 2 x T=%d student rollouts' % (N, S, T),
           'ms_per_cycle': ms, 'agent_env_steps_per_s': 2 * T * N / (ms * 1e-3), 'levels_built_per_s': N / (ms * 1e-3)}
    # the student rollouts have a fixed launch sequence (resident action tensors): captured once, replayed per cycle
    for k in range(2):
        graphs[k] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graphs[k]):
            student_rollout(v, sts[k], acts[k])
    ms_g = timed(lambda: cycle(True))
    out.update({'ms_per_cycle_student_rollouts_as_cuda_graphs': ms_g, 'agent_env_steps_per_s_graphs': 2 * T * N / (ms_g * 1e-3)})
    print(json.dumps(out), flush=True)
    del graphs
    v.close()


def accel(N, T, NB=4000):
    v = CudaAdversarialVecEnv('MultiGrid-GoalLastVariableBlocksAdversarialEnv-Edit-v0', N)
    v.set_seed(list(range(N)))
    np.random.seed(0)
    s = LevelSampler([], None, None, num_actors=N, strategy='positive_value_loss', replay_schedule='fixed', score_transform='rank',
                     temperature=0.3, rho=0.5, replay_prob=0.8, staleness_coef=0.3, sample_full_distribution=True,
                     seed_buffer_size=NB)
    store = LevelStore(data_info={'numpy': True, 'dtype': np.uint8, 'shape': (v.W, v.W, 3)})  # adversarial_runner.py:142-150
    # fill the buffer with DR levels (byte encodings) and random scores (SURVEY.md 8d synthetic PLR state)
    while len(store.seed2level) < NB:
        v.reset_random()
        store.insert([e.tobytes() for e in v.get_encodings()])
    seeds_all = np.array(sorted(store.seed2level.keys()))[:NB]
    s.seeds[:] = seeds_all
    s.seed2index = {int(x): i for i, x in enumerate(s.seeds)}
    s.working_seed_set = set(int(x) for x in s.seeds)
    s.working_seed_buffer_size = NB
    s.seed_scores[:] = np.random.rand(NB)
    s.unseen_seed_weights[:] = 0
    s.seed_staleness[:] = np.floor(np.random.rand(NB) * 50)
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    acts = torch.randint(0, 7, (T, N), device='cuda', generator=g)
    acts[torch.rand(T, N, device='cuda', generator=g) < 0.5] = 2
    st = DeviceRolloutStorage(T, N)
    st.value_preds.copy_(torch.rand(T + 1, N, 1, device='cuda', generator=g))

    phases = {}

    def ph(name, fn, profile):
        if not profile:
            return fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        phases[name] = phases.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return out

    def cycle(profile=False):
        # replay: sample N levels, load them, roll out, update the sampler (adversarial_runner.py:462-466,497-622)
        seeds = ph('sample_replay_levels', lambda: s.sample_replay_levels(N), profile)
        levels = ph('store.get_level', lambda: [store.get_level(int(x)) for x in seeds], profile)
        ph('reset_to_level_batch', lambda: v.reset_to_level_batch(levels), profile)
        st.level_seeds.copy_(torch.as_tensor(np.asarray(seeds, dtype=np.int32), device='cuda').view(1, N, 1).expand(T, N, 1))
        ph('student_rollout', lambda: student_rollout(v, st, acts), profile)
        ph('update_with_rollouts', lambda: (s.update_with_rollouts(st), s.after_update()), profile)
        # ACCEL: mutate the replayed levels, insert the children, evaluate them (adversarial_runner.py:455-461,756-794)
        ph('mutate_level', lambda: v.mutate_level(5), profile)
        enc = ph('get_encodings', lambda: [e.tobytes() for e in v.get_encodings()], profile)
        child = ph('store.insert', lambda: store.insert(enc, parent_seeds=[int(x) for x in seeds]), profile)
        ph('observe_unseen', lambda: s.observe_external_unseen_sample(child), profile)
        st.level_seeds.copy_(torch.as_tensor(np.asarray(child, dtype=np.int32), device='cuda').view(1, N, 1).expand(T, N, 1))
        ph('student_rollout', lambda: student_rollout(v, st, acts), profile)
        ph('update_with_rollouts (admissions)', lambda: (s.update_with_rollouts(st), s.after_update()), profile)
        ph('reconcile', lambda: store.reconcile_seeds(s.working_seed_set), profile)

    ms = timed(cycle, reps=2, warm=1)
    cycle(True)
    print(json.dumps({'config': 'configs[3] ACCEL, MultiGrid-GoalLastVariableBlocksAdversarialEnv-Edit-v0, %d envs, buffer %d, replay + %d-edit mutation, 2 x T=%d rollouts' % (N, NB, 5, T),
                      'ms_per_cycle': ms, 'agent_env_steps_per_s': 2 * T * N / (ms * 1e-3),
                      'phase_ms': {k: round(x, 3) for k, x in phases.items()}}), flush=True)
    v.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=4096)
    ap.add_argument('--T', type=int, default=256)
    a = ap.parse_args()
    paired(a.envs, a.T)
    accel(32, a.T)
    accel(a.envs, a.T)


if __name__ == '__main__':
    main()
