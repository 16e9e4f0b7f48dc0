// mgplr_wide.cu -- MultiGrid mazes wider than the 32-column bit-plane of the main path: the Kruskal perfect mazes
// PerfectMazeLarge (51x51) and PerfectMazeXL (101x101) of the zero-shot benchmark (envs/multigrid/mst_maze.py:128-136,
// eval.py:340-349).  Evaluation-only environments, a handful of processes each (arguments.py:433-436), so this is the plain
// formulation -- one thread per env, rows of four 32-bit words (128 columns), unpacked per-env state -- around the SAME view
// code as the hot kernels: a 32-column window of the wall rows centred on the agent is handed to render_packed /
// emit_packed_f32 (mgplr_env.cuh), so observations are produced by the code the golden traces pin.
//
// Semantics: MultiGridEnv.step / step_one_agent (multigrid.py:943-975,866-941) for one agent, no TimeLimit (the mazes are
// registered without one), VecMonitor episode statistics (vec_monitor.py:60-85), observation preprocessing
// (obs_wrappers.py:104-110).  The maze generator and the goal-respawn draws of agent_is_done run on the host on each env's
// own numpy RandomState (dcd_isaac_b200/mst_maze.py); a finished env gets its next maze through mgplr_wide_load_levels.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <new>

#include "mgplr_env.cuh"

using namespace mgplr;

extern "C" int mgplr_set_error_(int code, const char *msg);
static int wfail(int code, const char *msg) { return mgplr_set_error_(code, msg); }
#define WCK(call)                                                        \
  do {                                                                   \
    cudaError_t _e = (call);                                             \
    if (_e != cudaSuccess) return wfail((int)_e, cudaGetErrorString(_e)); \
  } while (0)

constexpr int kWideWords = 4;   // 128 columns
constexpr int kWideMax = 128;

struct WideState {   // one env
  int32_t ax, ay, adir, gx, gy, sx, sy, sdir;
  int32_t step_count, ep_len, done_flag, pad;
  float ep_ret, pad2, pad3, pad4;
};

struct mgplr_wide {
  int N, W, max_steps, device;
  uint32_t *wall;     // [N][W][4]
  WideState *st;      // [N]
};

// 32-column window [x0, x0 + 32) of row y; everything outside the grid is wall
__device__ __forceinline__ uint32_t wide_window(const uint32_t *rows, int W, int y, int x0) {
  if (y < 0 || y >= W) return 0xffffffffu;
  uint32_t out = 0;
  const uint32_t *r = rows + (size_t)y * kWideWords;
#pragma unroll 4
  for (int k = 0; k < 32; k++) {
    const int x = x0 + k;
    const uint32_t bit = (x < 0 || x >= W) ? 1u : ((r[x >> 5] >> (x & 31)) & 1u);
    out |= bit << k;
  }
  return out;
}

// the agent's observation through the shared view code: a local 32x(2*8+1) frame with the agent at column 12
__device__ void wide_emit(const uint32_t *rows, int W, const WideState &s, float *image, float *direction, int e) {
  constexpr int kCol = 12, kRowPad = 8;
  uint32_t win[2 * kRowPad + 1];
  const int x0 = s.ax - kCol, y0 = s.ay - kRowPad;
  for (int k = 0; k < 2 * kRowPad + 1; k++) win[k] = wide_window(rows, W, y0 + k, x0);
  Env v{};
  v.ax = kCol; v.ay = kRowPad; v.adir = s.adir; v.has_agent = 1;
  const int lgx = s.gx - x0, lgy = s.gy - y0;
  const bool g_in = lgx >= 0 && lgx < 32 && lgy >= 0 && lgy < 2 * kRowPad + 1;
  v.gx = g_in ? lgx : kNone; v.gy = g_in ? lgy : kNone;
  const Rows R{win, 1};
  const PackedView pv = render_packed<true, uint64_t>(R, v, 2 * kRowPad + 1, 32);
  float o[kObsFloats];
  emit_packed_f32<true, true>(pv, o);
  if (image) for (int i = 0; i < kObsFloats; i++) image[(size_t)e * kObsFloats + i] = o[i];
  if (direction) direction[e] = (float)s.adir;
}

__global__ void k_wide_load(mgplr_wide h, const uint8_t *enc, const int32_t *env_index, int n, int start_dir, float *image,
                            float *direction) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int e = env_index ? env_index[k] : k;
  if (e < 0 || e >= h.N) return;
  const int W = h.W;
  uint32_t *rows = h.wall + (size_t)e * W * kWideWords;
  const uint8_t *src = enc + (size_t)k * W * W * 3;
  WideState s;
  memset(&s, 0, sizeof(s));
  s.gx = s.gy = s.sx = s.sy = kNone;
  for (int y = 0; y < W; y++) {
    uint32_t w[kWideWords] = {0, 0, 0, 0};
    for (int x = 0; x < W; x++) {
      const uint8_t t = src[((size_t)x * W + y) * 3];
      if (t == 2) w[x >> 5] |= 1u << (x & 31);
      else if (t == 8) { s.gx = x; s.gy = y; }
      else if (t == 10) { s.sx = x; s.sy = y; if (start_dir < 0) s.sdir = src[((size_t)x * W + y) * 3 + 2] & 3; }
    }
    for (int q = 0; q < kWideWords; q++) rows[(size_t)y * kWideWords + q] = w[q];
  }
  if (start_dir >= 0) s.sdir = start_dir & 3;
  s.ax = s.sx; s.ay = s.sy; s.adir = s.sdir;
  h.st[e] = s;
  wide_emit(rows, W, s, image, direction, e);
}

__global__ void k_wide_step(mgplr_wide h, const int64_t *action, mgplr_step_out o) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= h.N) return;
  const int W = h.W;
  const uint32_t *rows = h.wall + (size_t)e * W * kWideWords;
  WideState s = h.st[e];
  uint32_t flags = 0;
  double rew = 0.0;
  const int a = (int)action[e];
  s.step_count++;
  const int fx = s.ax + ((s.adir == 0) - (s.adir == 2)), fy = s.ay + ((s.adir == 1) - (s.adir == 3));
  if (a == 0) s.adir = (s.adir + 3) & 3;
  else if (a == 1) s.adir = (s.adir + 1) & 3;
  else if (a == 2) {
    if (fx == s.gx && fy == s.gy) {   // agent_is_done (multigrid.py:821-838): the respawn draws are the host generator's
      s.done_flag = 1;
      rew = __dsub_rn(1.0, __dmul_rn(0.9, __ddiv_rn((double)s.step_count, (double)h.max_steps)));
      flags |= MGPLR_F_GOAL;
    } else {
      const bool wall = fx < 0 || fx >= W || fy < 0 || fy >= W || ((rows[(size_t)fy * kWideWords + (fx >> 5)] >> (fx & 31)) & 1u);
      if (!wall) { s.ax = fx; s.ay = fy; }
    }
  }
  const bool done = s.done_flag || s.step_count >= h.max_steps;
  if (flags & MGPLR_F_GOAL) s.ep_ret = (float)__dadd_rn((double)s.ep_ret, rew);
  s.ep_len += 1;
  if (done) {
    flags |= MGPLR_F_DONE;
    if (o.ep_return) o.ep_return[e] = s.ep_ret;
    if (o.ep_length) o.ep_length[e] = s.ep_len;
    s.ep_ret = 0.f; s.ep_len = 0;
    // the worker answers `done` with env.reset() (parallel_wrappers.py:20-25): the host uploads the next maze; until then the
    // env sits at its start (what reset_agent would give)
    s.ax = s.sx; s.ay = s.sy; s.adir = s.sdir; s.step_count = 0; s.done_flag = 0;
  }
  h.st[e] = s;
  wide_emit(rows, W, s, o.image, o.direction, e);
  if (o.reward) o.reward[e] = (float)rew;
  if (o.flags) o.flags[e] = (uint8_t)flags;
}

// AdversarialEnv.encoding analogue (Grid.encode with the agent at its cell): u8 [N][W][W][3]
__global__ void k_wide_encode(mgplr_wide h, uint8_t *enc) {
  const int W = h.W, WW = W * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)h.N * WW) return;
  const int e = (int)(idx / WW), cell = (int)(idx % WW), x = cell / W, y = cell % W;
  const WideState s = h.st[e];
  const uint32_t *rows = h.wall + (size_t)e * W * kWideWords;
  uint8_t t, c, stt = 0;
  if (x == s.ax && y == s.ay) { t = 10; c = 0; stt = (uint8_t)s.adir; }
  else if ((rows[(size_t)y * kWideWords + (x >> 5)] >> (x & 31)) & 1u) { t = 2; c = 5; }
  else if (x == s.gx && y == s.gy) { t = 8; c = 1; }
  else { t = 1; c = 0; }
  uint8_t *p = enc + idx * 3;
  p[0] = t; p[1] = c; p[2] = stt;
}

extern "C" int mgplr_wide_create(int32_t width, int32_t max_steps, int32_t num_envs, int32_t device, mgplr_wide **out) {
  if (!out) return wfail(MGPLR_E_BADARG, "out is NULL");
  *out = nullptr;
  if (width < 5 || width > kWideMax) return wfail(MGPLR_E_UNSUPPORTED, "width must be in [5, 128]");
  if (num_envs < 1 || max_steps < 1) return wfail(MGPLR_E_BADARG, "num_envs and max_steps must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return wfail((int)cudaErrorNoDevice, "mgplr: no CUDA device -- this library has no CPU fallback");
  WCK(cudaSetDevice(device));
  mgplr_wide *h = new (std::nothrow) mgplr_wide();
  if (!h) return wfail(MGPLR_E_BADARG, "out of host memory");
  h->N = num_envs; h->W = width; h->max_steps = max_steps; h->device = device;
  WCK(cudaMalloc((void **)&h->wall, (size_t)num_envs * width * kWideWords * sizeof(uint32_t)));
  WCK(cudaMalloc((void **)&h->st, (size_t)num_envs * sizeof(WideState)));
  WCK(cudaMemset(h->wall, 0xff, (size_t)num_envs * width * kWideWords * sizeof(uint32_t)));
  WCK(cudaMemset(h->st, 0, (size_t)num_envs * sizeof(WideState)));
  *out = h;
  return 0;
}

extern "C" void mgplr_wide_destroy(mgplr_wide *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->wall); cudaFree(h->st);
  delete h;
}

extern "C" int mgplr_wide_load_levels(mgplr_wide *h, const uint8_t *enc, const int32_t *env_index, int32_t n, int32_t start_dir,
                                      const mgplr_step_out *out, void *stream) {
  if (!h || !enc || n < 1 || (!env_index && n > h->N)) return wfail(MGPLR_E_BADARG, "mgplr_wide_load_levels: bad arguments");
  WCK(cudaSetDevice(h->device));
  k_wide_load<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(*h, enc, env_index, n, start_dir, out ? out->image : nullptr,
                                                            out ? out->direction : nullptr);
  WCK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_wide_step(mgplr_wide *h, const int64_t *action, const mgplr_step_out *out, void *stream) {
  if (!h || !action || !out) return wfail(MGPLR_E_BADARG, "mgplr_wide_step: bad arguments");
  WCK(cudaSetDevice(h->device));
  k_wide_step<<<(h->N + 63) / 64, 64, 0, (cudaStream_t)stream>>>(*h, action, *out);
  WCK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_wide_get_encodings(mgplr_wide *h, uint8_t *enc, void *stream) {
  if (!h || !enc) return wfail(MGPLR_E_BADARG, "mgplr_wide_get_encodings: bad arguments");
  WCK(cudaSetDevice(h->device));
  const size_t total = (size_t)h->N * h->W * h->W;
  k_wide_encode<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*h, enc);
  WCK(cudaGetLastError());
  return 0;
}
