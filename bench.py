#!/usr/bin/env python
"""bench.py -- MultiGrid env-steps/s (with observations) on N B200s, HBM roofline fraction, CPU baseline.

One bench "step" = one pass of the hot path over one batch: a T=256-step PLR rollout of `--envs` MultiGrid
environments per GPU (15x15, 25 blocks: BASELINE.json configs[1]) -- T launches of the step kernel writing
float32 observations, rewards and masks straight into rollout storage -- followed by GAE and the PLR
positive-value-loss episode-score reduction (and, for N>1 GPUs, the NCCL all-gather of the episode records).
Actions/values are synthetic and resident in HBM before the timed region.  `value` = env-steps/s over all
GPUs; `e2e` drives the same rollout through the host-buffer C-ABI call (mgplr_step_env_host every vector step: the
kernel reads the pinned int64 actions over PCIe and writes flags + done records back to pinned host memory).

  python bench.py [--gpus N --steps K --warmup W] [--impl reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_STEP = 360  # algorithmic bytes per env-step, drop-in fp32 layout


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--envs', type=int, default=524288, help='environments per GPU (SURVEY.md 8d sizes: 32 / 4096 / 131072 / 524288 = the 1M-env sweep on 2 GPUs)')
    ap.add_argument('--T', type=int, default=256, help='rollout length (num_steps)')
    ap.add_argument('--size', type=int, default=15)
    ap.add_argument('--blocks', type=int, default=25)
    ap.add_argument('--opaque', type=int, default=0, help='1: see_through_walls=False (occlusion on)')
    ap.add_argument('--reset-random', type=int, default=0, help='1: DR auto-reset (reset_random) instead of PLR reset_agent')
    ap.add_argument('--cpu-envs', type=int, default=16384)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying a CUDA graph')
    return ap.parse_args()


def workload_name(a):
    return 'MultiGrid %dx%d %d-block levels (%s walls), %d envs/GPU, T=%d rollout + GAE + PLR positive_value_loss scores, auto-reset=%s' % (
        a.size, a.size, a.blocks, 'opaque' if a.opaque else 'see-through', a.envs, a.T,
        'reset_random' if a.reset_random else 'reset_agent')


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def traffic_from_profile(N, a):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of k_step_env from the committed `ncu --set full`
    capture (profiles/traffic.json), when it was taken on this workload; else None."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        t = json.load(open(p))['k_step_env']
        if t['envs'] == N and t['size'] == a.size and t['opaque'] == a.opaque and not a.reset_random:
            return t['dram_bytes_per_launch']
    except Exception:
        pass
    return None


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th.start()

    def stop(self):
        self.stop_flag = True
        self.th.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(self.rows)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_rollout_rate(a, n_envs, steps, warmup, threads):
    """The oracle port (oracle/c/mg_oracle.c) on the host cores: same workload, bounded sample."""
    import numpy as np
    from oracle import mg_oracle as mo
    cfg = mo.make_cfg(W=a.size, see_through=not a.opaque, n_clutter=2 * a.blocks)
    b = mo.OracleBatch(cfg, n_envs)
    for i in range(n_envs):
        b.seed(i, i)
        b.reset_random(i)
    rs = np.random.RandomState(1)
    acts = rs.randint(0, 7, size=(a.T, n_envs)).astype(np.uint8)
    acts[rs.rand(a.T, n_envs) < 0.5] = 2
    obs = np.empty((a.T, n_envs, 3, 5, 5), np.float32)
    L = b.L
    times = []
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        L.mgo_rollout_batch(C.c_void_p(b.base), n_envs, a.T, acts.ctypes.data_as(C.c_void_p), int(a.reset_random),
                            obs.ctypes.data_as(C.c_void_p), None, None, threads)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
    tot = sum(times)
    return n_envs * a.T * len(times) / tot, tot / len(times)


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = a.cpu_envs
    rate, per = cpu_rollout_rate(a, n, max(1, a.steps), min(a.warmup, 1), cores)
    sample = '%d envs x T=%d (one step = %.2f s of wall time on %d threads)' % (n, a.T, per, cores)
    line = {
        'impl': 'reference', 'metric': 'MultiGrid env-steps/sec (with obs)', 'value': rate, 'unit': 'env-steps/s',
        'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': per * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': workload_name(a), 'note': 'reference CPU arm = C port of the reference algorithm '
                   '(oracle/c/mg_oracle.c); the reference itself is Python over un-installed third-party deps and cannot travel'},
        'cpu_baseline': {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from dcd_isaac_b200 import _lib
    from dcd_isaac_b200._lib import StepOut, ptr, check
    from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the hot path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    L = _lib.load()
    N, T = a.envs, a.T

    venv = CudaAdversarialVecEnv('MultiGrid-GoalLastFewerBlocksAdversarial-v0', N, device=dev, size=a.size,
                                 n_clutter=2 * a.blocks, see_through_walls=not a.opaque)
    venv.set_seed([rank * N + i for i in range(N)])
    venv.reset_random()  # synthetic levels: reset_random semantics with n_clutter/2 = `blocks` walls

    # rollout storage (algos/storage.py:62-112 layouts)
    obs_img = torch.zeros(T + 1, N, 3, 5, 5, device=dev)
    obs_dir = torch.zeros(T + 1, N, 1, device=dev)
    rewards = torch.zeros(T, N, 1, device=dev)
    masks = torch.ones(T + 1, N, 1, device=dev)
    bad_masks = torch.ones(T + 1, N, 1, device=dev)
    cliff = torch.ones(T + 1, N, 1, device=dev)
    returns = torch.zeros(T + 1, N, 1, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1 + rank)
    values = torch.rand(T + 1, N, 1, device=dev, generator=g)
    level_seeds = torch.randint(1, 4001, (T, N, 1), device=dev, dtype=torch.int32, generator=g)
    actions = torch.randint(0, 7, (T, N), device=dev, generator=g)
    fwd = torch.rand(T, N, device=dev, generator=g) < 0.5  # forward-biased stream so goals are reached
    actions[fwd] = 2
    actions = actions.contiguous()
    flags = torch.zeros(T, N, dtype=torch.uint8, device=dev)
    ep_r = torch.zeros(N, device=dev)
    ep_l = torch.zeros(N, dtype=torch.int32, device=dev)
    max_eps = N * 3  # ~2.2 episodes per env per rollout here (checked below); the kernel drops records beyond the cap
    episodes = torch.zeros(max_eps, 10, dtype=torch.int32, device=dev)
    n_eps = torch.zeros(1, dtype=torch.int32, device=dev)
    gathered = torch.zeros(world * max_eps, 10, dtype=torch.int32, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    outs = []
    for t in range(T):
        o = StepOut()
        o.image, o.direction, o.reward, o.flags = ptr(obs_img[t + 1]), ptr(obs_dir[t + 1]), ptr(rewards[t]), ptr(flags[t])
        o.ep_return, o.ep_length = ptr(ep_r), ptr(ep_l)
        o.masks, o.bad_masks, o.cliffhanger_masks = ptr(masks[t + 1]), ptr(bad_masks[t + 1]), ptr(cliff[t + 1])
        outs.append(o)
    act_ptrs = [ptr(actions[t]) for t in range(T)]
    rr = int(a.reset_random)

    launches = [0]

    def cur_stream():
        return torch.cuda.current_stream(dev).cuda_stream

    def env_steps():
        st = cur_stream()
        for t in range(T):
            last = 3 if t == T - 1 else 0
            check(L.mgplr_step_env(venv.h, act_ptrs[t], rr, None, last, C.byref(outs[t]), st))

    def rollout():
        st = cur_stream()
        # reset_agent -> obs[0] (adversarial_runner.py:484-487)
        check(L.mgplr_reset_agent(venv.h, C.byref(venv._out({'image': obs_img[0], 'direction': obs_dir[0]})), st))
        env_steps()
        check(L.mgplr_gae(ptr(rewards), ptr(values), ptr(masks), ptr(returns), T, N, 0.995, 0.95, st))
        check(L.mgplr_plr_episode_scores(ptr(masks), ptr(cliff), ptr(returns), ptr(values), ptr(rewards), ptr(level_seeds),
                                         T, N, 0, ptr(episodes), max_eps, ptr(n_eps), st))
        if world > 1:
            dist.all_gather_into_tensor(gathered, episodes)
        launches[0] += 1 + T + 1 + 4

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # nvidia-smi sampling runs from the warm-up to the end of the e2e pass (the K timed steps alone last a few
    # milliseconds, shorter than one nvidia-smi query), so every sample is taken under load
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(a.warmup, 3)):
        rollout()
    sync_all()
    use_graph = not a.no_graph
    if use_graph:
        # The rollout's launch sequence is fixed (actions are a device-resident recorded stream), so it is captured
        # once and replayed: launch latency of T+5 kernels is off the critical path (B200 guide: CUDA graphs).
        g_roll, g_steps = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_roll):
            rollout()
        with torch.cuda.graph(g_steps):
            env_steps()
        per_roll = launches[0] // (max(a.warmup, 3) + 1)
        run_roll = g_roll.replay
        run_steps = g_steps.replay
        for _ in range(2):
            run_roll()
        sync_all()
    else:
        per_roll = 1 + T + 1 + 4
        run_roll, run_steps = rollout, env_steps
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.steps):
        run_roll()
    t1.record()
    sync_all()
    ms = t0.elapsed_time(t1)
    # the dominant kernel alone: T step launches, CUDA events on the launching stream
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(a.steps):
        run_steps()
    k1.record()
    sync_all()
    kernel_ms = [k0.elapsed_time(k1) / a.steps]
    # the same T transitions in ONE launch (mgplr_rollout: recorded action stream, env state stays on chip)
    act_u8 = actions.to(torch.uint8).contiguous()
    o_all = StepOut()
    o_all.image, o_all.direction, o_all.reward, o_all.flags = ptr(obs_img[1:]), ptr(obs_dir[1:]), ptr(rewards), ptr(flags)
    o_all.masks, o_all.bad_masks, o_all.cliffhanger_masks = ptr(masks[1:]), ptr(bad_masks[1:]), ptr(cliff[1:])
    check(L.mgplr_rollout(venv.h, ptr(act_u8), T, rr, C.byref(o_all), cur_stream()))
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(a.steps):
        check(L.mgplr_rollout(venv.h, ptr(act_u8), T, rr, C.byref(o_all), cur_stream()))
    f1.record()
    sync_all()
    fused_ms = f0.elapsed_time(f1) / a.steps
    launches[0] = per_roll * a.steps
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    n_launch = launches[0]
    total_steps = N * T * a.steps * world
    value = total_steps / (ms * 1e-3)
    avg_launch_s = (sum(kernel_ms) / len(kernel_ms)) * 1e-3 / T
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_STEP * N / avg_launch_s / 1e9
    n_done = int((flags & 1).sum().item())
    n_goal = int(((flags & 8) > 0).sum().item())
    n_episodes = int(n_eps.item())
    assert n_episodes <= max_eps, 'episode-record buffer too small: %d > %d' % (n_episodes, max_eps)

    # ---- e2e: the host-buffer C-ABI call every vector step (actions from pinned host memory, results to host)
    e2e = None
    if not a.no_e2e:
        h_act = actions.to(torch.int64).cpu().pin_memory()
        h_flg = torch.zeros(N, dtype=torch.uint8).pin_memory()
        h_done = torch.zeros(N * 16, dtype=torch.uint8).pin_memory()
        h_nd = torch.zeros(1, dtype=torch.int32).pin_memory()
        hp = [ptr(h_act[t]) for t in range(T)]
        nd_np = h_nd.numpy()  # (reading the count through numpy: no tensor indexing in the per-step loop)
        p_flg, p_done, p_nd = ptr(h_flg), ptr(h_done), ptr(h_nd)
        step_host = L.mgplr_step_env_host
        out_refs = [C.byref(o) for o in outs]
        done_seen = [0]

        def rollout_host():
            check(L.mgplr_reset_agent(venv.h, C.byref(venv._out({'image': obs_img[0], 'direction': obs_dir[0]})), stream))
            for t in range(T):
                check(step_host(venv.h, hp[t], rr, 3 if t == T - 1 else 0, out_refs[t], p_flg, p_done, N, p_nd, stream))
                done_seen[0] += int(nd_np[0])
            check(L.mgplr_gae(ptr(rewards), ptr(values), ptr(masks), ptr(returns), T, N, 0.995, 0.95, stream))
            check(L.mgplr_plr_episode_scores(ptr(masks), ptr(cliff), ptr(returns), ptr(values), ptr(rewards), ptr(level_seeds),
                                             T, N, 0, ptr(episodes), max_eps, ptr(n_eps), stream))
            if world > 1:
                dist.all_gather_into_tensor(gathered, episodes)
            return int(n_eps.item())  # device->host read of the step's result

        rollout_host()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ksteps = max(1, min(a.steps, 3))
        e0.record()
        for _ in range(ksteps):
            rollout_host()
        e1.record()
        sync_all()
        ems = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ems], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ems = float(tt.item())
        dones_per_rollout = done_seen[0] // (ksteps + 1)
        e2e = {'value': N * T * ksteps * world / (ems * 1e-3), 'unit': 'env-steps/s',
               'h2d_bytes_per_step': T * N * 8, 'd2h_bytes_per_step': T * N + 16 * dones_per_rollout + 4,
               'api': 'mgplr_step_env_host every vector step (T calls per rollout): the kernel reads the pinned int64 actions over PCIe '
                      '(zero-copy), writes flags u8[N] and the done records into pinned host memory, one stream sync per vector step; '
                      'observations / rewards / masks stay in rollout storage'}

    clocks = sampler.stop() if rank == 0 else None
    cpu = None
    if rank == 0 and not a.no_cpu:
        cores = os.cpu_count() or 1
        rate, per = cpu_rollout_rate(a, a.cpu_envs, 8, 1, cores)
        cpu = {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
               'sample': '%d envs x T=%d per timed pass (%.2f s each) on %d threads, oracle/c/mg_oracle.c' % (a.cpu_envs, T, per, cores)}

    if rank == 0:
        line = {
            'metric': 'MultiGrid env-steps/sec (with obs)', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world,
            'steps': a.steps, 'warmup': max(a.warmup, 3), 'ms_per_step': ms / a.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': workload_name(a), 'l2': 'inputs+outputs per rollout (%.1f GB) exceed the 126 MB L2' %
                       (N * T * 410 / 1e9), 'episodes_per_rollout': n_episodes, 'done_steps': n_done, 'goals': n_goal,
                       'state_bytes': venv.state_bytes(), 'launch': 'cuda-graph replay' if use_graph else 'per-kernel launches from Python'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic_from_profile(N, a), 'kernel': 'k_step_env', 'bytes_per_env_step': BYTES_PER_STEP,
                         'avg_launch_us': avg_launch_s * 1e6, 'peak_source': peak_src},
            'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': n_launch, 'clocks': clocks,
            'fused_rollout': {'note': 'extra, not the headline: mgplr_rollout steps the same T transitions in ONE launch from the recorded '
                              'action stream (state stays on chip); this rank only', 'env_steps_per_s': N * T / (fused_ms * 1e-3),
                              'us_per_step': fused_ms * 1e3 / T, 'frac_at_360B': BYTES_PER_STEP * N * T / (fused_ms * 1e-3) / 1e9 / peak},
        }
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    # teardown: CUDA graphs that captured NCCL work must die before the communicator; never hang the driver on exit
    if use_graph:
        del g_roll, g_steps, run_roll, run_steps
    torch.cuda.synchronize(dev)
    venv.close()
    if world > 1:
        try:
            dist.barrier()
        finally:
            os._exit(0)


def main():
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
