#!/usr/bin/env python
"""bench.py -- MultiGrid env-steps/s (with observations) on N B200s, HBM roofline fraction, CPU baselines.

One bench "step" = one pass of the hot path over one batch: a T=256-step PLR rollout of `--envs` MultiGrid
environments per GPU (15x15, 25 blocks: BASELINE.json configs[1]) -- reset_agent, T launches of the step kernel
writing float32 observations, rewards and masks straight into rollout storage, then GAE and the PLR
positive-value-loss episode-score reduction (and, for N>1 GPUs, the NCCL all-gather of the compact episode records).
Actions/values are synthetic and resident in HBM before the timed region.  `value` = env-steps/s over all GPUs.

The rollout is replayed as three CUDA graphs (reset_agent | T step launches | GAE + scores) with CUDA events recorded
between them, so `roofline.avg_launch_us` is the step kernel's launch time INSIDE the timed rollouts
(avg_launch_us * T <= ms_per_step by construction).  `variants` repeats the measurement for the other configurations
north_star names (25x25 / 50 blocks with occlusion, 15x15 with occlusion, DR auto-reset) at the same envs/GPU, each
with its own roofline fraction.  `e2e` drives the same rollout through the host-buffer C-ABI call
(mgplr_step_env_host_u8 every vector step: the kernel reads the pinned uint8 actions over PCIe and appends the done
records back to pinned host memory).  `cpu_baseline` = the C port of the reference algorithm on all host threads;
`cpu_baseline_python` = the reference's OWN vectorised Python path (util.create_parallel_env -> step_env loop from the
staged copy baseline/_ref/reference) on the same host cores.

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
METRIC = 'MultiGrid env-steps/sec (with obs)'


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
    ap.add_argument('--cpu-envs', type=int, default=16384, help='envs per pass of the C-port CPU arm (passes are repeated up to the workload size)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-variants', action='store_true')
    ap.add_argument('--no-python-ref', action='store_true', help='skip the reference-Python CPU leg (subprocess vector env)')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying CUDA graphs')
    return ap.parse_args()


def workload_name(a, size=None, blocks=None, opaque=None, rr=None):
    size = a.size if size is None else size
    blocks = a.blocks if blocks is None else blocks
    opaque = a.opaque if opaque is None else opaque
    rr = a.reset_random if rr is None else rr
    return 'MultiGrid %dx%d %d-block levels (%s walls), %d envs/GPU, T=%d rollout + GAE + PLR positive_value_loss scores, auto-reset=%s' % (
        size, size, blocks, 'opaque' if opaque else 'see-through', a.envs, a.T, 'reset_random' if rr else 'reset_agent')


def config_of(a):
    """Identical in both arms (the driver compares them)."""
    return {'workload': workload_name(a),
            'l2': 'inputs+outputs per rollout (%.1f GB at %d envs) exceed the 126 MB L2' % (a.envs * a.T * 410 / 1e9, a.envs)}


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def traffic_from_profile(N, size, opaque, rr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of k_step_env from the committed `ncu --set full`
    capture (profiles/traffic.json), when it was taken on this workload; else None."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        t = json.load(open(p))['k_step_env']
        if t['envs'] == N and t['size'] == size and t['opaque'] == opaque and not rr:
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


# ---------------------------------------------------------------------------------------------- CPU arms
def cpu_port_rate(a, steps, warmup, threads, budget_s=6.0):
    """The oracle port (oracle/c/mg_oracle.c) on the host cores.  One timed step = `passes` back-to-back T-step rollouts of
    a --cpu-envs batch, passes chosen so that a step covers the workload's env-step count (envs x T) whenever that fits in
    `budget_s` seconds per step, otherwise as many passes as fit (the sample then says what was covered)."""
    import numpy as np
    from oracle import mg_oracle as mo
    n = min(a.cpu_envs, a.envs)
    cfg = mo.make_cfg(W=a.size, see_through=not a.opaque, n_clutter=2 * a.blocks)
    b = mo.OracleBatch(cfg, n)
    for i in range(n):
        b.seed(i, i)
        b.reset_random(i)
    rs = np.random.RandomState(1)
    acts = rs.randint(0, 7, size=(a.T, n)).astype(np.uint8)
    acts[rs.rand(a.T, n) < 0.5] = 2
    obs = np.empty((a.T, n, 3, 5, 5), np.float32)
    L = b.L

    def one_pass():
        t0 = time.perf_counter()
        L.mgo_rollout_batch(C.c_void_p(b.base), n, a.T, acts.ctypes.data_as(C.c_void_p), int(a.reset_random),
                            obs.ctypes.data_as(C.c_void_p), None, None, threads)
        return time.perf_counter() - t0

    one_pass()
    probe = one_pass()
    full = max(1, (a.envs + n - 1) // n)
    passes = max(1, min(full, int(budget_s / max(probe, 1e-6))))
    times = []
    for k in range(warmup + steps):
        dt = sum(one_pass() for _ in range(passes))
        if k >= warmup:
            times.append(dt)
    tot = sum(times)
    covered = passes * n
    sample = '%d passes of %d envs x T=%d per step = %d env-steps (%s the workload\'s %d), %.2f s per step on %d threads, oracle/c/mg_oracle.c' % (
        passes, n, a.T, covered * a.T, 'all of' if covered >= a.envs else '%.1f%% of' % (100.0 * covered / a.envs), a.envs * a.T,
        tot / len(times), threads)
    return covered * a.T * len(times) / tot, tot / len(times), sample


PY_REF_SNIPPET = r'''
import json, os, sys, time
os.environ.setdefault("OMP_NUM_THREADS", "1")
sys.path.insert(0, %(oracle)r)
import ref_harness as rh
rh.activate()
import numpy as np, torch
from types import SimpleNamespace
import util
N, T, env_name = %(n)d, %(T)d, %(env)r
args = SimpleNamespace(env_name=env_name, seed=1, singleton_env=False, use_global_critic=False, use_global_policy=False,
                       num_processes=N, normalize_returns=False)
t0 = time.perf_counter()
venv, _ = util.create_parallel_env(args)
venv.reset_random(); venv.reset_agent()
t_setup = time.perf_counter() - t0
rs = np.random.RandomState(1)
acts = rs.randint(0, 7, size=(T, N, 1)).astype(np.int64)
acts[rs.rand(T, N, 1) < 0.5] = 2
acts = torch.from_numpy(acts)
for t in range(8):
    venv.step_env(acts[t], reset_random=False)
t0 = time.perf_counter()
for t in range(T):
    obs, rew, done, infos = venv.step_env(acts[t], reset_random=False)
dt = time.perf_counter() - t0
venv.close()
print(json.dumps({"n": N, "T": T, "s": dt, "setup_s": t_setup, "rate": N * T / dt}))
'''


def cpu_python_reference(a, timeout_s=170):
    """the reference's vectorised path -- util.create_parallel_env (spawn-subprocess
    ParallelAdversarialVecEnv + VecMonitor + VecNormalize + VecPreprocessImageWrapper) then a T-step venv.step_env loop
    with pre-generated actions -- UNMODIFIED from the staged copy baseline/_ref/reference over oracle/shim (the un-installed
    third-party gym / gym-minigrid / baselines), OMP_NUM_THREADS=1, at num_processes = 32 (configs[0]) and = cpu_count."""
    from oracle import ref_harness as rh
    if not rh.available():
        return {'unavailable': 'no reference tree (baseline/_ref/reference not staged)'}
    if a.size != 15:
        return {'unavailable': 'registered reference envs are 15x15'}
    cores = os.cpu_count() or 1
    env_name = ('MultiGrid-GoalLastFewerBlocksOpaqueWallsAdversarial-v0' if a.opaque else 'MultiGrid-GoalLastFewerBlocksAdversarial-v0')
    runs = []
    t_start = time.time()
    procs = [int(x) for x in os.environ['MGPLR_BENCH_PYREF_PROCS'].split(',')] if os.environ.get('MGPLR_BENCH_PYREF_PROCS') else sorted({32, cores})
    for n in procs:
        left = timeout_s - (time.time() - t_start)
        if left < 20:
            break
        code = PY_REF_SNIPPET % {'oracle': os.path.join(ROOT, 'oracle'), 'n': n, 'T': a.T, 'env': env_name}
        try:
            out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=left,
                                 env=dict(os.environ, OMP_NUM_THREADS='1', CUDA_VISIBLE_DEVICES=''))
            lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
            if out.returncode == 0 and lines:
                runs.append(json.loads(lines[-1]))
            else:
                runs.append({'n': n, 'error': (out.stderr or out.stdout)[-300:]})
        except subprocess.TimeoutExpired:
            runs.append({'n': n, 'error': 'timeout'})
    ok = [r for r in runs if 'rate' in r]
    if not ok:
        return {'unavailable': 'reference vector env did not run: %s' % (runs[-1].get('error') if runs else 'no time')}
    best = max(ok, key=lambda r: r['rate'])
    return {'value': best['rate'], 'unit': 'env-steps/s', 'cores': cores, 'kind': 'reference',
            'sample': 'reference util.create_parallel_env + %d-step venv.step_env loop, num_processes=%d (best of %s), env %s' % (
                a.T, best['n'], [r['n'] for r in ok], env_name),
            'runs': [{'num_processes': r['n'], 'env_steps_per_s': r['rate'], 'ms_per_vector_step': r['s'] / r['T'] * 1e3,
                      'setup_s': r['setup_s']} for r in ok]}


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rate, per, sample = cpu_port_rate(a, max(1, a.steps), min(a.warmup, 1), cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'env-steps/s',
        'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': per * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic', 'config': config_of(a),
        'cpu_baseline': {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                         'what': 'C port of the reference algorithm (oracle/c/mg_oracle.c: env step + observation only, no GAE / PLR -- '
                                 'conservative); the reference itself is Python, see cpu_baseline_python'},
        'e2e': {'value': rate, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    if not a.no_python_ref:
        line['cpu_baseline_python'] = cpu_python_reference(a)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
class Buffers(object):
    """Rollout storage (algos/storage.py:62-112 layouts) + synthetic inputs; shared by the main case and the variants."""

    def __init__(self, torch, dev, N, T, world, rank):
        self.obs_img = torch.zeros(T + 1, N, 3, 5, 5, device=dev)
        self.obs_dir = torch.zeros(T + 1, N, 1, device=dev)
        self.rewards = torch.zeros(T, N, 1, device=dev)
        self.masks = torch.ones(T + 1, N, 1, device=dev)
        self.bad_masks = torch.ones(T + 1, N, 1, device=dev)
        self.cliff = torch.ones(T + 1, N, 1, device=dev)
        self.returns = torch.zeros(T + 1, N, 1, device=dev)
        g = torch.Generator(device=dev)
        g.manual_seed(1 + rank)
        self.values = torch.rand(T + 1, N, 1, device=dev, generator=g)
        self.level_seeds = torch.randint(1, 4001, (T, N, 1), device=dev, dtype=torch.int32, generator=g)
        actions = torch.randint(0, 7, (T, N), device=dev, generator=g)
        fwd = torch.rand(T, N, device=dev, generator=g) < 0.5  # forward-biased stream so goals are reached
        actions[fwd] = 2
        self.actions = actions.contiguous()
        self.actions_u8 = actions.to(torch.uint8).contiguous()
        self.flags = torch.zeros(T, N, dtype=torch.uint8, device=dev)
        self.ep_r = torch.zeros(N, device=dev)
        self.ep_l = torch.zeros(N, dtype=torch.int32, device=dev)
        self.max_eps = N * 3  # ~2.2 episodes per env per rollout here (checked); the kernel drops records beyond the cap
        self.episodes = torch.zeros(self.max_eps, 10, dtype=torch.int32, device=dev)
        self.n_eps = torch.zeros(1, dtype=torch.int32, device=dev)
        # multi-GPU: compact records (actor, seed, mean score, max score, steps|flag: 20 B instead of 40) are what crosses
        # NVLink; two send buffers so that rollout k's gather (side stream) overlaps rollout k+1's launches
        self.compact = [torch.zeros(self.max_eps, 5, dtype=torch.int32, device=dev) for _ in range(2)] if world > 1 else None
        self.gathered = [torch.zeros(world * self.max_eps, 5, dtype=torch.int32, device=dev) for _ in range(2)] if world > 1 else None


def measure_case(torch, dist, L, a, B, dev, rank, world, size, blocks, opaque, rr, steps, warmup, use_graph, want_fused=False):
    """Build a venv for (size, blocks, opaque, rr), run `warmup` + `steps` rollouts into the shared buffers and return the
    timing record; the step kernel is timed INSIDE the rollouts (events between the rollout's three graph segments)."""
    from dcd_isaac_b200._lib import StepOut, ptr, check
    from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv
    N, T = a.envs, a.T
    venv = CudaAdversarialVecEnv('MultiGrid-GoalLastFewerBlocksAdversarial-v0', N, device=dev, size=size,
                                 n_clutter=2 * blocks, see_through_walls=not opaque)
    venv.set_seed([rank * N + i for i in range(N)])
    venv.reset_random()  # synthetic levels: reset_random semantics with n_clutter/2 = `blocks` walls
    outs = []
    for t in range(T):
        o = StepOut()
        o.image, o.direction, o.reward, o.flags = ptr(B.obs_img[t + 1]), ptr(B.obs_dir[t + 1]), ptr(B.rewards[t]), ptr(B.flags[t])
        o.ep_return, o.ep_length = ptr(B.ep_r), ptr(B.ep_l)
        o.masks, o.bad_masks, o.cliffhanger_masks = ptr(B.masks[t + 1]), ptr(B.bad_masks[t + 1]), ptr(B.cliff[t + 1])
        outs.append(o)
    act_ptrs = [ptr(B.actions[t]) for t in range(T)]
    o0 = venv._out({'image': B.obs_img[0], 'direction': B.obs_dir[0]})
    side = torch.cuda.Stream(device=dev) if world > 1 else None

    def cur():
        return torch.cuda.current_stream(dev).cuda_stream

    def seg_pre():   # reset_agent -> obs[0] (adversarial_runner.py:484-487)
        check(L.mgplr_reset_agent(venv.h, C.byref(o0), cur()))

    def seg_steps():
        st = cur()
        for t in range(T):
            check(L.mgplr_step_env(venv.h, act_ptrs[t], int(rr), None, 3 if t == T - 1 else 0, C.byref(outs[t]), st))

    def make_post(k):
        def seg_post():
            st = cur()
            check(L.mgplr_gae(ptr(B.rewards), ptr(B.values), ptr(B.masks), ptr(B.returns), T, N, 0.995, 0.95, st))
            check(L.mgplr_plr_episode_scores(ptr(B.masks), ptr(B.cliff), ptr(B.returns), ptr(B.values), ptr(B.rewards),
                                             ptr(B.level_seeds), T, N, 0, ptr(B.episodes), B.max_eps, ptr(B.n_eps), st))
            if world > 1:   # compact the records for the wire: actor, seed, mean, max, (t_end - t_start) | cliffhanger << 16
                c = B.compact[k]
                c[:, 0].copy_(B.episodes[:, 0]); c[:, 1].copy_(B.episodes[:, 3])
                c[:, 2].copy_(B.episodes[:, 4]); c[:, 3].copy_(B.episodes[:, 5])
                torch.add(B.episodes[:, 2] - B.episodes[:, 1], B.episodes[:, 9], alpha=65536, out=c[:, 4])
        return seg_post

    posts = [make_post(0), make_post(1)]
    for _ in range(max(warmup, 3)):
        seg_pre(); seg_steps(); posts[0]()
    torch.cuda.synchronize(dev)
    if use_graph:
        g_pre, g_steps = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        g_post = [torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()] if world > 1 else [torch.cuda.CUDAGraph()]
        with torch.cuda.graph(g_pre):
            seg_pre()
        with torch.cuda.graph(g_steps):
            seg_steps()
        for k, g in enumerate(g_post):
            with torch.cuda.graph(g):
                posts[k]()
        run_pre, run_steps = g_pre.replay, g_steps.replay
        run_post = [g.replay for g in g_post]
    else:
        run_pre, run_steps, run_post = seg_pre, seg_steps, posts
    gather_done = [None, None]

    def rollout(k, ev=None):
        run_pre()
        if ev is not None:
            ev[0].record()
        run_steps()
        if ev is not None:
            ev[1].record()
        b = k & 1 if world > 1 else 0
        if world > 1 and gather_done[b] is not None:
            torch.cuda.current_stream(dev).wait_event(gather_done[b])   # send buffer b is free again
        run_post[b]()
        if world > 1:   # the gather of rollout k runs on the side stream, under rollout k+1's launches
            ready = torch.cuda.Event()
            ready.record()
            side.wait_event(ready)
            with torch.cuda.stream(side):
                dist.all_gather_into_tensor(B.gathered[b], B.compact[b])
                gather_done[b] = torch.cuda.Event()
                gather_done[b].record()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for k in range(2):
        rollout(k)
    sync_all()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(steps):
        rollout(k, evs[k])
    if world > 1:
        torch.cuda.current_stream(dev).wait_stream(side)   # the last gather ends inside the timed region
    t1.record()
    sync_all()
    ms = t0.elapsed_time(t1)
    steps_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    n_episodes = int(B.n_eps.item())
    assert n_episodes <= B.max_eps, 'episode-record buffer too small: %d > %d' % (n_episodes, B.max_eps)
    peak, peak_src = measured_peak()
    avg_launch_s = steps_ms * 1e-3 / (steps * T)
    achieved = BYTES_PER_STEP * N / avg_launch_s / 1e9
    rec = {
        'workload': workload_name(a, size, blocks, opaque, rr), 'value': N * T * steps * world / (ms * 1e-3), 'unit': 'env-steps/s',
        'ms_per_step': ms / steps,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic_from_profile(N, size, opaque, rr), 'kernel': 'k_step_env', 'bytes_per_env_step': BYTES_PER_STEP,
                     'avg_launch_us': avg_launch_s * 1e6, 'step_segment_ms_per_rollout': steps_ms / steps,
                     'timed': 'CUDA events around the T step launches inside each timed rollout', 'peak_source': peak_src},
        'episodes_per_rollout': n_episodes, 'done_steps': int((B.flags & 1).sum().item()), 'goals': int(((B.flags & 8) > 0).sum().item()),
        'state_bytes': venv.state_bytes(), 'launches_per_rollout': 1 + T + T // 64 + 1 + 4 + (6 if world > 1 else 0),
    }
    extra = {}
    if want_fused:
        # the same T transitions in ONE launch (mgplr_rollout: recorded action stream, env state stays on chip)
        o_all = StepOut()
        o_all.image, o_all.direction, o_all.reward, o_all.flags = ptr(B.obs_img[1:]), ptr(B.obs_dir[1:]), ptr(B.rewards), ptr(B.flags)
        o_all.masks, o_all.bad_masks, o_all.cliffhanger_masks = ptr(B.masks[1:]), ptr(B.bad_masks[1:]), ptr(B.cliff[1:])
        check(L.mgplr_rollout(venv.h, ptr(B.actions_u8), T, int(rr), C.byref(o_all), cur()))
        sync_all()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(steps):
            check(L.mgplr_rollout(venv.h, ptr(B.actions_u8), T, int(rr), C.byref(o_all), cur()))
        f1.record()
        sync_all()
        fused_ms = f0.elapsed_time(f1) / steps
        extra['fused_rollout'] = {'note': 'extra, not the headline: mgplr_rollout steps the same T transitions in ONE launch from the '
                                  'recorded action stream (state stays on chip); this rank only',
                                  'env_steps_per_s': N * T / (fused_ms * 1e-3), 'us_per_step': fused_ms * 1e3 / T,
                                  'frac_at_360B': BYTES_PER_STEP * N * T / (fused_ms * 1e-3) / 1e9 / peak}
    return rec, extra, venv, (outs, o0, sync_all)


def run_ours(a):
    import torch
    import torch.distributed as dist
    from dcd_isaac_b200 import _lib
    from dcd_isaac_b200._lib import ptr, check

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
    use_graph = not a.no_graph
    B = Buffers(torch, dev, N, T, world, rank)

    # nvidia-smi sampling runs from the warm-up to the end of the e2e pass (the K timed steps alone last a few
    # milliseconds, shorter than one nvidia-smi query), so every sample is taken under load
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    main, extra, venv, (outs, o0, sync_all) = measure_case(torch, dist, L, a, B, dev, rank, world, a.size, a.blocks, a.opaque,
                                                           a.reset_random, a.steps, a.warmup, use_graph, want_fused=True)

    # ---- e2e: the host-buffer C-ABI call every vector step (uint8 actions from pinned host memory, results to host)
    e2e = None
    if not a.no_e2e:
        stream = torch.cuda.current_stream(dev).cuda_stream
        h_act = B.actions_u8.cpu().pin_memory()
        h_done = torch.zeros(N * 16, dtype=torch.uint8).pin_memory()
        h_nd = torch.zeros(1, dtype=torch.int32).pin_memory()
        hp = [ptr(h_act[t]) for t in range(T)]
        nd_np = h_nd.numpy()  # (reading the count through numpy: no tensor indexing in the per-step loop)
        # (no flags array: since ABI v4 the done records carry each finished env's flags, and only finished envs have any)
        p_flg, p_done, p_nd = None, ptr(h_done), ptr(h_nd)
        step_host = L.mgplr_step_env_host_u8
        out_refs = [C.byref(o) for o in outs]
        done_seen = [0]
        rr = int(a.reset_random)

        def rollout_host():
            check(L.mgplr_reset_agent(venv.h, C.byref(o0), stream))
            for t in range(T):
                check(step_host(venv.h, hp[t], rr, 3 if t == T - 1 else 0, out_refs[t], p_flg, p_done, N, p_nd, stream))
                done_seen[0] += int(nd_np[0])
            check(L.mgplr_gae(ptr(B.rewards), ptr(B.values), ptr(B.masks), ptr(B.returns), T, N, 0.995, 0.95, stream))
            check(L.mgplr_plr_episode_scores(ptr(B.masks), ptr(B.cliff), ptr(B.returns), ptr(B.values), ptr(B.rewards),
                                             ptr(B.level_seeds), T, N, 0, ptr(B.episodes), B.max_eps, ptr(B.n_eps), stream))
            if world > 1:
                dist.all_gather_into_tensor(B.gathered[0], B.compact[0])
            return int(B.n_eps.item())  # device->host read of the step's result

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
               'h2d_bytes_per_step': T * N, 'd2h_bytes_per_step': 16 * dones_per_rollout + 4 * T + 4,
               'api': 'mgplr_step_env_host_u8 every vector step (T calls per rollout): the kernel reads the pinned uint8 actions over PCIe '
                      '(zero-copy) and appends the done records (env, reward, episode return, length | flags) to pinned host memory, one stream sync per vector step; '
                      'observations / rewards / masks stay in rollout storage'}
    venv.close()

    # ---- the other configurations north_star names, same envs/GPU, each with its own roofline fraction
    variants = []
    if not a.no_variants:
        vsteps = max(2, min(a.steps, 5))
        todo = [(25, 50, 1, 0), (15, 25, 1, 0), (15, 25, 0, 1)]   # (size, blocks, opaque, reset_random)
        for size, blocks, opaque, rr in todo:
            if (size, blocks, opaque, rr) == (a.size, a.blocks, a.opaque, a.reset_random):
                continue
            rec, _, v2, _ = measure_case(torch, dist, L, a, B, dev, rank, world, size, blocks, opaque, rr, vsteps, 3, use_graph)
            v2.close()
            rec['steps'] = vsteps
            variants.append(rec)

    clocks = sampler.stop() if rank == 0 else None
    cpu = cpu_py = None
    if rank == 0 and not a.no_cpu:
        cores = os.cpu_count() or 1
        rate, per, sample = cpu_port_rate(a, 3, 1, cores, budget_s=3.0)
        cpu = {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample}
        if not a.no_python_ref and world == 1:
            cpu_py = cpu_python_reference(a, timeout_s=120)

    if rank == 0:
        cfg = config_of(a)
        line = {
            'metric': METRIC, 'value': main['value'], 'unit': 'env-steps/s', 'n_gpus': world,
            'steps': a.steps, 'warmup': max(a.warmup, 3), 'ms_per_step': main['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic', 'config': cfg,
            'run_info': {'episodes_per_rollout': main['episodes_per_rollout'], 'done_steps': main['done_steps'], 'goals': main['goals'],
                         'state_bytes': main['state_bytes'],
                         'launch': 'cuda-graph replay (3 graphs per rollout, events between them)' if use_graph else 'per-kernel launches from Python',
                         'multi_gpu': None if world == 1 else 'compact 20-byte episode records all-gathered on a side stream under the next rollout'},
            'roofline': main['roofline'], 'cpu_baseline': cpu, 'cpu_baseline_python': cpu_py, 'e2e': e2e,
            'gpu_launches': main['launches_per_rollout'] * a.steps, 'clocks': clocks, 'variants': variants,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    torch.cuda.synchronize(dev)
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
