"""Size-independent properties at BASELINE.json's full sizes (131 072 envs per GPU, 15x15 and 25x25), where the
oracle is too slow to replay everything: round trips, idempotence, determinism, shard-equivalence, and the
algebraic invariants of the episode bookkeeping."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

N_FULL = 131072


def _venv(n, **kw):
    from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv
    return CudaAdversarialVecEnv('MultiGrid-GoalLastAdversarial-v0', n, **kw)


def _rollout(v, T, seed, rr=False):
    from dcd_isaac_b200._lib import StepOut, ptr
    N = v.num_envs
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    acts = torch.randint(0, 7, (T, N), device='cuda', generator=g)
    acts[torch.rand(T, N, device='cuda', generator=g) < 0.5] = 2
    img = torch.zeros(T, N, 3, 5, 5, device='cuda')
    rew = torch.zeros(T, N, 1, device='cuda')
    fl = torch.zeros(T, N, dtype=torch.uint8, device='cuda')
    epl = torch.zeros(N, dtype=torch.int32, device='cuda')
    lens = torch.zeros(N, dtype=torch.int64, device='cuda')
    for t in range(T):
        o = StepOut()
        o.image, o.reward, o.flags, o.ep_length = ptr(img[t]), ptr(rew[t]), ptr(fl[t]), ptr(epl)
        v.step_env_device(acts[t].contiguous(), o, reset_random=rr)
        lens += torch.where((fl[t] & 1) > 0, epl.long(), torch.zeros_like(lens))
    torch.cuda.synchronize()
    return acts, img, rew, fl, lens


@pytest.mark.parametrize('size,see', [(15, True), (25, False)])
def test_encoding_round_trip_full_size(size, see):
    """get_encodings -> reset_to_level_batch -> get_encodings is the identity on walls / goal / agent cell (the agent's
    direction byte is re-drawn by reset(), adversarial.py:271-272), and metrics are reproduced."""
    v = _venv(N_FULL, size=size, see_through_walls=see)
    v.set_seed(list(range(N_FULL)))
    v.reset_random()
    enc = v.get_encodings_device()
    m0 = v._metrics()
    w = _venv(N_FULL, size=size, see_through_walls=see)
    w.set_seed([7] * N_FULL)
    from dcd_isaac_b200._lib import check, ptr
    import ctypes as C
    o = w._out(w._new_obs())
    check(w.L.mgplr_reset_to_encoding(w.h, ptr(enc), None, N_FULL, C.byref(o), w._stream()))
    enc2 = w.get_encodings_device()
    same = (enc == enc2)
    agent = enc[..., 0] == 10
    assert bool(same[..., 0].all()) and bool(same[..., 1].all())
    assert bool((same[..., 2] | agent).all())
    assert np.array_equal(m0, w._metrics())
    # idempotence of reset_agent: a second call changes nothing observable
    a = w.reset_agent()
    b = w.reset_agent()
    assert torch.equal(a['image'], b['image']) and torch.equal(a['direction'], b['direction'])
    v.close(); w.close()


def test_determinism_and_shard_equivalence_full_size():
    """Same seeds -> bit-identical rollouts; and the env-sharded layout of the multi-GPU path (rank r owns envs
    [r*N/2, (r+1)*N/2)) reproduces the single-handle result bit-for-bit."""
    T = 64
    outs = []
    for _ in range(2):
        v = _venv(N_FULL)
        v.set_seed(list(range(N_FULL)))
        v.reset_random()
        outs.append(_rollout(v, T, 5) + (v.get_encodings_device(), torch.from_numpy(v.get_agent_state()).cuda()))
        v.close()
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    acts, img, rew, fl, lens = outs[0][:5]
    half = N_FULL // 2
    for r in range(2):
        from dcd_isaac_b200._lib import StepOut, ptr
        v = _venv(half)
        v.set_seed(list(range(r * half, (r + 1) * half)))
        v.reset_random()
        img_r = torch.zeros(T, half, 3, 5, 5, device='cuda')
        fl_r = torch.zeros(T, half, dtype=torch.uint8, device='cuda')
        for t in range(T):
            o = StepOut()
            o.image, o.flags = ptr(img_r[t]), ptr(fl_r[t])
            v.step_env_device(acts[t, r * half:(r + 1) * half].contiguous(), o)
        torch.cuda.synchronize()
        assert torch.equal(img_r, img[:, r * half:(r + 1) * half])
        assert torch.equal(fl_r, fl[:, r * half:(r + 1) * half])
        v.close()


@pytest.mark.parametrize('rr', [False, True])
def test_episode_invariants_full_size(rr):
    """Checksum-of-checksums style invariants over 131 072 x 300 transitions: every episode is at most
    max_episode_steps long, reported lengths add up to the steps since the last done, a reward is paid exactly on
    goal steps and lies in (0.1, 1], observations only contain the six legal values, the agent cell is empty."""
    T = 300
    v = _venv(N_FULL, see_through_walls=False)
    v.set_seed(list(range(N_FULL)))
    v.reset_random()
    v.reset_agent()
    acts, img, rew, fl, lens = _rollout(v, T, 9, rr=rr)
    done = (fl & 1) > 0
    goal = (fl & 8) > 0
    trunc = (fl & 2) > 0
    assert bool((done | ~goal).all()) and bool((done | ~trunc).all())
    assert bool(((rew[..., 0] > 0) == goal).all())
    r = rew[..., 0][goal]
    assert float(r.min()) > 0.1 - 1e-6 and float(r.max()) <= 1.0
    # every env hits the 250-step limit at least once within 300 steps unless it reached the goal before
    assert bool(done.any(0).all())
    # steps since the last done + sum of finished episode lengths == T
    last_done = torch.where(done, torch.arange(1, T + 1, device='cuda').view(T, 1), torch.zeros(1, dtype=torch.long, device='cuda')).max(0).values
    st = torch.from_numpy(v.get_agent_state()).cuda()
    assert torch.equal(lens + st[:, 4].long(), torch.full_like(lens, T))
    assert torch.equal(st[:, 4].long(), T - last_done)
    legal = torch.zeros_like(img[0], dtype=torch.bool)
    for t in range(T):  # 2.9e9 values: check per step instead of torch.unique over everything
        x = img[t]
        legal = (x == 0.0) | (x == 0.1) | (x == 0.2) | (x == 0.5) | (x == 0.8)
        assert bool(legal.all()), t
    assert bool((img[:, :, 0, 2, 4] == 0.1).all()) and bool((img[:, :, 2] == 0).all())
    v.close()


def test_dr_speculation_equals_in_kernel_rebuild_full_size(monkeypatch):
    """step_env(reset_random=True) at 131 072 envs x 520 steps (two synchronized time-limit storms): the speculative,
    pipelined level generation (default) and the in-kernel rebuild (MGPLR_RR_SPEC=0, the path checked against the oracle
    at 4 096 envs) must agree bit for bit -- observations, rewards, flags, final levels, metrics, agent state and the
    env-RNG stream position (words drawn)."""
    T = 520
    res = []
    for spec in ('1', '0'):
        monkeypatch.setenv('MGPLR_RR_SPEC', spec)
        v = _venv(N_FULL)
        v.set_seed(list(range(N_FULL)))
        v.reset_random()
        acts, img, rew, fl, lens = _rollout(v, T, 21, rr=True)
        res.append((fl, rew, img, v.get_encodings_device(), torch.from_numpy(v._metrics()).cuda(),
                    torch.from_numpy(v.get_agent_state()).cuda()))
        v.close()
    assert int(((res[0][0] & 1) > 0).sum()) > 2 * N_FULL  # every env was rebuilt at least twice
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_dr_speculation_under_cuda_graph_replay():
    """The job lists of the DR speculation alternate by a DEVICE-side launch parity, so a captured sequence of step
    launches -- here an ODD number of them -- can be replayed back to back and still equals launching step by step."""
    from dcd_isaac_b200._lib import StepOut, ptr
    N, K, R = 32768, 7, 40
    g = torch.Generator(device='cuda')
    g.manual_seed(3)
    acts = torch.randint(0, 7, (K * R, N), device='cuda', generator=g)
    acts[torch.rand(K * R, N, device='cuda', generator=g) < 0.6] = 2
    outs = []
    for use_graph in (False, True):
        v = _venv(N)
        v.set_seed(list(range(N)))
        v.reset_random()
        img = torch.zeros(K, N, 3, 5, 5, device='cuda')
        fl = torch.zeros(K, N, dtype=torch.uint8, device='cuda')
        cur = torch.zeros(K, N, dtype=torch.int64, device='cuda')
        hist_img, hist_fl = [], []

        def k_steps():
            for k in range(K):
                o = StepOut()
                o.image, o.flags = ptr(img[k]), ptr(fl[k])
                v.step_env_device(cur[k], o, reset_random=True)

        if use_graph:
            cur.copy_(acts[:K])
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                k_steps()  # warm-up launch outside the capture (first-use attribute calls)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            hist_img.append(img.clone()); hist_fl.append(fl.clone())
            gr = torch.cuda.CUDAGraph()
            cur.copy_(acts[K:2 * K])
            with torch.cuda.graph(gr):
                k_steps()
            gr.replay()
            torch.cuda.synchronize()
            hist_img.append(img.clone()); hist_fl.append(fl.clone())
            for r in range(2, R):
                cur.copy_(acts[r * K:(r + 1) * K])
                gr.replay()
                torch.cuda.synchronize()
                hist_img.append(img.clone()); hist_fl.append(fl.clone())
            del gr
        else:
            for r in range(R):
                cur.copy_(acts[r * K:(r + 1) * K])
                k_steps()
                torch.cuda.synchronize()
                hist_img.append(img.clone()); hist_fl.append(fl.clone())
        outs.append((torch.stack(hist_fl), torch.stack(hist_img), v.get_encodings_device(), torch.from_numpy(v.get_agent_state()).cuda()))
        v.close()
    assert int(((outs[0][0] & 1) > 0).sum()) > N
    for a, b in zip(*outs):
        assert torch.equal(a, b)
