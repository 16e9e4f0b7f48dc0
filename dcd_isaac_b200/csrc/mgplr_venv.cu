// mgplr_venv.cu -- kernels and C-ABI entry points for the batched MultiGrid adversarial env (sm_100a).
//
// Hot kernel: k_step_env -- persistent and warp-pipelined: each warp stages 32-env tiles of the wall bit-plane
// in shared memory with TMA bulk loads (struct-of-arrays, bank index = lane), one thread steps one env, and
// each tile's float32 observations leave the SM as ONE 9600-byte bulk asynchronous store straight into the
// rollout-storage tensor while the next tile computes.  See DESIGN.md for the byte accounting and the roofline.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "mgplr_env.cuh"

using namespace mgplr;

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char *msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
static int cuda_fail(cudaError_t e, const char *where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}
#define CK(call)                                              \
  do {                                                        \
    cudaError_t _e = (call);                                  \
    if (_e != cudaSuccess) return cuda_fail(_e, #call);       \
  } while (0)

extern "C" const char *mgplr_last_error(void) { return g_err; }
extern "C" int mgplr_set_error_(int code, const char *msg) { return fail(code, msg); }
extern "C" int mgplr_abi_version(void) { return MGPLR_ABI_VERSION; }

// ------------------------------------------------------------------------------------------ handle
struct mgplr_venv {
  Dev d;
  int device;
  int64_t bytes;
  // scratch for host->device argument staging
  uint32_t *seed_scratch;  // [4][N]
  int64_t *act_dev;        // [N]   (mgplr_step_env_host)
  uint32_t *cnt_dev;       // [2] ping-pong append counters of the host-driven step (each launch zeroes the other one)
  uint8_t *res_pin;        // pinned + device-mapped: [N done records (env = -1: empty slot)][N flags]
  uint8_t *res_pin_dev;    // the same allocation as the device sees it
  uint32_t host_steps;     // host-driven steps issued (selects the ping-pong counter)
  int host_dma;            // host-driven step: chunked copy-engine staging of pinned actions for large batches
  cudaStream_t copy_stream;
  cudaEvent_t copy_done[8];
  int pdl;                 // launch the step kernel with programmatic stream serialization
  int rr_spec;             // DR auto-reset: speculative next-level candidates (MGPLR_RR_SPEC=0 disables)
  int sm_count;
  int steps_since_sweep;   // reset_agent-mode step launches since the last deferred-respawn sweep (kSweepEvery)
};

// dynamic shared memory: `bufs` obs tiles [TILE][75] f32 then wall rows [W][TILE] u32
// (+ the batched-RNG scratch [32][tile] of the in-kernel reset_random)
static size_t step_smem_bytes(int W, int tile, int bufs, bool rr = true) {
  return (size_t)bufs * tile * kObsFloats * 4 + (size_t)W * tile * 4 + (rr ? (size_t)32 * tile * 4 : 0);
}

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA) before a bulk store
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared -> global bulk asynchronous copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}

// ------------------------------------------------------------------------------------------ obs emit (cold paths)
struct OutPtrs {
  float *image, *direction;
  uint8_t *image_u8;
};
__device__ __noinline__ void emit_direct(const Rows &R, const Env &e, const Cfg &c, const OutPtrs &o, int row) {
  if (!o.image && !o.image_u8 && !o.direction) return;
  View v = c.see_through ? render_view<true>(R, e, c.W) : render_view<false>(R, e, c.W);
  if (o.image) emit_obs_f32(v, o.image + (size_t)row * kObsFloats);
  if (o.image_u8) emit_obs_u8(v, o.image_u8 + (size_t)row * kObsFloats);
  if (o.direction) o.direction[row] = (float)e.adir;
}

// ------------------------------------------------------------------------------------------ kernels: level ops
// shared scratch of the batched RNG for 128-thread level kernels: [32][128] words, one column per thread
#define RNG_SCRATCH() __shared__ uint32_t s_rng_[32 * 128]
#define RNG_OF(d, e) Rng((d), (e), s_rng_ + threadIdx.x, 128)

__global__ void k_init(Dev d) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const Rows R = env_rows(d, e);
  gen_grid(R, d.c.W);
  Env s{};
  s.gx = s.gy = s.sx = s.sy = kNone;
  d.hot[e] = pack(s);
  d.adv[e] = (uint32_t)(d.c.n_clutter + 2) << 12;
  d.metrics[e] = make_int4(0, -1, -1, (d.c.W - 2) * (d.c.W - 2) + 1);
  d.mti[e] = 0; d.words[e] = 0; d.err[e] = 0;
  d.limbs[e] = 0; d.limbs[(size_t)d.N + e] = 0; d.limbs[2 * (size_t)d.N + e] = 1;
}

// scratch: [0][k] lo, [1][k] hi, [2][k] count, [3][k] env index
__global__ void k_seed(Dev d, const uint32_t *scratch, int n, int stride, int has_index) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int e = has_index ? (int)scratch[3 * (size_t)stride + k] : k;
  if (e < 0 || e >= d.N) return;
  const uint32_t lo = scratch[k], hi = scratch[(size_t)stride + k], cnt = scratch[2 * (size_t)stride + k];
  d.limbs[e] = lo; d.limbs[(size_t)d.N + e] = hi; d.limbs[2 * (size_t)d.N + e] = cnt;
  mt_seed(d, e, lo, hi, (int)cnt);
  uint4 h = d.hot[e];  // re-seeding replaces the stream: deferred draws of the old stream are moot
  h.y &= 0x00ffffffu;
  d.hot[e] = h;
}

__global__ void k_reset(Dev d) {
  RNG_SCRATCH();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const Rows R = env_rows(d, e);
  Env s = unpack(d.hot[e]);
  uint32_t adv = d.adv[e];
  int4 met;
  Rng rng = RNG_OF(d, e);
  reset_adversary(R, s, adv, met, rng, d.c);
  rng.store();
  d.hot[e] = pack(s); d.adv[e] = adv; d.metrics[e] = met;
}

__global__ void k_step_adversary(Dev d, const int64_t *loc, uint8_t *done) {
  RNG_SCRATCH();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const Rows R = env_rows(d, e);
  Env s = unpack(d.hot[e]);
  uint32_t adv = d.adv[e], err = 0;
  int4 met = d.metrics[e];
  Rng rng = RNG_OF(d, e);
  const long long l = loc[e];
  const bool dn = step_adversary(R, s, adv, met, rng, d.c, (l < 0 || l > 0x7fffffff) ? -1 : (int)l, err);
  rng.store();
  d.hot[e] = pack(s); d.adv[e] = adv; d.metrics[e] = met;
  if (err) d.err[e] |= err;
  if (done) done[e] = dn ? 1 : 0;
}

// Adversary observation image f32 [N][3][W][W] = Grid.encode()/10 with channels first
// (adversarial.py:222-227,532-537; obs_wrappers.py:104-110).  A CTA assembles the images of kAdvGroup consecutive
// envs in shared memory (one thread per cell: type, colour, state planes) and writes them with ONE bulk asynchronous
// store (kAdvGroup * 3*W*W*4 bytes, contiguous in the output and 16-byte aligned for groups of 4 envs); a ragged or
// misaligned last group falls back to plain coalesced stores.
constexpr int kAdvGroup = 4;
constexpr int kFuseAdvMaxEnvs = 16384;
// mode 0: image only; 1: AdversarialEnv.reset first; 2: step_adversary(loc) first -- the state update of the group's
// envs runs on the first n_env threads of the same CTA, so reset()/step_adversary() + observation is ONE launch.
// FUSED is a template parameter so that the image-only instance keeps its small register footprint (occupancy).
// raw != 0: unscaled codes (obs['full_obs'] of MultiGridFullyObsWrapper) instead of /10.
template <bool FUSED>
__global__ void __launch_bounds__(128) k_adv_image(Dev d, float *image, float *time_step, int mode, const int64_t *loc,
                                                   uint8_t *done, int raw = 0) {
  RNG_SCRATCH();
  extern __shared__ __align__(128) float s_img[];
  const int W = d.c.W, WW = W * W, per_env = 3 * WW;
  const int e0 = blockIdx.x * kAdvGroup, n_env = min(kAdvGroup, d.N - e0);
  if (FUSED && mode && (int)threadIdx.x < n_env) {
    const int e = e0 + threadIdx.x;
    const Rows R = env_rows(d, e);
    Env s = unpack(d.hot[e]);
    uint32_t adv = d.adv[e], err = 0;
    int4 met = d.metrics[e];
    Rng rng = RNG_OF(d, e);
    if (mode == 1) reset_adversary(R, s, adv, met, rng, d.c);
    else {
      const long long l = loc[e];
      const bool dn = step_adversary(R, s, adv, met, rng, d.c, (l < 0 || l > 0x7fffffff) ? -1 : (int)l, err);
      if (done) done[e] = dn ? 1 : 0;
    }
    rng.store();
    d.hot[e] = pack(s); d.adv[e] = adv; d.metrics[e] = met;
    if (err) d.err[e] |= err;
  }
  if (FUSED && mode) __syncthreads();
  for (int i = threadIdx.x; i < n_env * WW; i += blockDim.x) {
    const int k = i / WW, cell = i - k * WW, x = cell / W, y = cell - x * W, e = e0 + k;
    const Env s = unpack(d.hot[e]);
    float t, c, st = 0.f;
    if (s.has_agent && x == s.ax && y == s.ay) { t = 1.0f; c = 0.0f; st = s.adir == 0 ? 0.0f : s.adir == 1 ? 0.1f : s.adir == 2 ? 0.2f : 0.3f; }
    else if ((env_rows(d, e).get(y) >> x) & 1u) { t = 0.2f; c = 0.5f; }
    else if (x == s.gx && y == s.gy) { t = 0.8f; c = 0.1f; }
    else { t = 0.1f; c = 0.0f; }
    if (raw) {  // exact small integers (0.2f * 10 is not 2.0f)
      t = (t == 1.0f) ? 10.f : (t == 0.2f) ? 2.f : (t == 0.8f) ? 8.f : 1.f;
      c = (c == 0.5f) ? 5.f : (c == 0.1f) ? 1.f : 0.f;
      st = (s.has_agent && x == s.ax && y == s.ay) ? (float)s.adir : 0.f;
    }
    float *o = s_img + k * per_env + cell;
    o[0] = t; o[WW] = c; o[2 * WW] = st;
  }
  if (time_step && (int)threadIdx.x < n_env) time_step[e0 + threadIdx.x] = (float)(d.adv[e0 + threadIdx.x] & 0xfff);
  float *gdst = image + (size_t)e0 * per_env;
  const uint32_t bytes = (uint32_t)(n_env * per_env * 4);
  if ((bytes & 15u) == 0 && (((uintptr_t)gdst) & 15u) == 0) {
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_store(gdst, s_img, bytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    __syncthreads();
    for (int i = threadIdx.x; i < n_env * per_env; i += blockDim.x) gdst[i] = s_img[i];
  }
}

// The float32 observation of all N envs is written once per rollout (obs[0]): each warp stages its 32 observations in
// shared memory (75 floats per lane, conflict-free) and writes them as 600 consecutive float4 (coalesced); one thread
// writing its own 300 bytes would touch 10 sectors per store instruction.
__global__ void __launch_bounds__(128) k_reset_agent(Dev d, OutPtrs o) {
  RNG_SCRATCH();
  extern __shared__ __align__(16) float s_tile[];  // [4 warps][32][75]
  const int e = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int base = e - lane;
  bool ok = false;
  Env s{};
  if (e < d.N) {
    const Rows R = env_rows(d, e);
    s = unpack(d.hot[e]);
    if (s.pending) {  // replay the deferred respawn draws of the last rollout (level unchanged since)
      Rng rng = RNG_OF(d, e);
      flush_pending(R, s, rng, d.c.W);
      rng.store();
    }
    ok = reset_agent(s);
    if (!ok) d.err[e] |= kErrNoStart;
    else { s.ep_ret = 0.f; s.ep_len = 0; }  // VecMonitor.reset_agent (vec_monitor.py:42-46)
    d.hot[e] = pack(s);
    if (ok && o.image_u8) emit_direct(R, s, d.c, OutPtrs{nullptr, nullptr, o.image_u8}, e);
    if (ok && o.direction) o.direction[e] = (float)s.adir;
  }
  if (!o.image) return;
  float *tile = s_tile + warp * (32 * kObsFloats);
  if (ok) {
    const Rows R = env_rows(d, e);
    const View v = d.c.see_through ? render_view<true>(R, s, d.c.W) : render_view<false>(R, s, d.c.W);
    emit_obs_f32(v, tile + lane * kObsFloats);
  }
  const unsigned okm = __ballot_sync(0xffffffffu, ok);
  float *gdst = o.image + (size_t)base * kObsFloats;
  if (okm == 0xffffffffu && (((uintptr_t)gdst) & 15u) == 0) {
    const float4 *src = reinterpret_cast<const float4 *>(tile);
    float4 *dst = reinterpret_cast<float4 *>(gdst);
    for (int i = lane; i < 32 * kObsFloats / 4; i += 32) dst[i] = src[i];
  } else {  // ragged tile / envs in error keep their previous observation
    for (int k = 0; k < 32; k++)
      if ((okm >> k) & 1u)
        for (int i = lane; i < kObsFloats; i += 32) gdst[k * kObsFloats + i] = tile[k * kObsFloats + i];
  }
}

__global__ void __launch_bounds__(128) k_reset_random(Dev d, const int32_t *n_walls, OutPtrs o) {
  RNG_SCRATCH();
  extern __shared__ uint32_t s_rr[];  // [W][128]: the grid is rebuilt in shared memory and written back once
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const int W = d.c.W;
  const Rows G = env_rows(d, e);
  const Rows R{s_rr + threadIdx.x, 128};
  Env s = unpack(d.hot[e]);
  if (s.pending) for (int r = 0; r < W; r++) R.set(r, G.get(r));  // deferred respawns replay against the old level
  uint32_t adv = d.adv[e], err = 0;
  int4 met;
  Rng rng = RNG_OF(d, e);
  const int nw = (d.c.resample && n_walls) ? n_walls[e] : -1;
  reset_random(R, s, adv, met, rng, d, e, nw, err);
  rng.store();
  for (int r = 0; r < W; r++) G.set(r, R.get(r));
  s.ep_ret = 0.f; s.ep_len = 0;  // VecMonitor.reset_random (vec_monitor.py:48-52)
  d.hot[e] = pack(s); d.adv[e] = adv; d.metrics[e] = met;
  if (err) d.err[e] |= err;
  emit_direct(R, s, d.c, o, e);
}

// reset_to_level, byte form (adversarial.py:271-294, multigrid.py:264-280): reset() draws a FRESH start
// direction; the encoding's agent dir byte is not restored.
__global__ void k_reset_to_encoding(Dev d, const uint8_t *enc, const int32_t *index, int n, OutPtrs o) {
  RNG_SCRATCH();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int e = index ? index[k] : k;
  if (e < 0 || e >= d.N) return;
  const int W = d.c.W;
  const Rows R = env_rows(d, e);
  Env s = unpack(d.hot[e]);
  uint32_t adv = d.adv[e];
  int4 met;
  Rng rng = RNG_OF(d, e);
  reset_adversary(R, s, adv, met, rng, d.c);
  rng.store();
  const uint8_t *src = enc + (size_t)k * W * W * 3;
  for (int y = 0; y < W; y++) {
    uint32_t row = 0;
    for (int x = 0; x < W; x++) {
      const uint8_t t = src[((size_t)x * W + y) * 3];
      if (t == 2) row |= 1u << x;
      else if (t == 8) { s.gx = x; s.gy = y; }
      else if (t == 10) { s.sx = x; s.sy = y; }
    }
    R.set(y, row);
  }
  // set_encoding visits i (x) outer, j (y) inner: with several goals/agents the LAST in that order wins
  // (multigrid.py:267-280); single-goal/agent levels are order independent.
  met = compute_metrics(R, s, W, false);
  if (!reset_agent(s)) d.err[e] |= kErrNoStart;
  d.hot[e] = pack(s); d.adv[e] = adv; d.metrics[e] = met;
  emit_direct(R, s, d.c, o, e);
}

// Fixed-level load (MazeEnv._gen_grid, maze.py:75-94): no env-RNG use, explicit start direction.
__global__ void k_load_levels(Dev d, const uint8_t *enc, int n_levels, const int32_t *level_index, int start_dir, OutPtrs o) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const int W = d.c.W;
  const Rows R = env_rows(d, e);
  Env s = unpack(d.hot[e]);
  int lv = level_index ? level_index[e] : 0;
  if (lv < 0 || lv >= n_levels) lv = 0;
  const uint8_t *src = enc + (size_t)lv * W * W * 3;
  s.gx = s.gy = s.sx = s.sy = kNone;
  for (int y = 0; y < W; y++) {
    uint32_t row = 0;
    for (int x = 0; x < W; x++) {
      const uint8_t t = src[((size_t)x * W + y) * 3];
      if (t == 2) row |= 1u << x;
      else if (t == 8) { s.gx = x; s.gy = y; }
      else if (t == 10) { s.sx = x; s.sy = y; }
    }
    R.set(y, row);
  }
  s.sdir = start_dir & 3;
  s.pending = 0;  // deferred respawn draws belong to the previous level's stream position: dropped with it
  s.ep_ret = 0.f; s.ep_len = 0;
  spec_invalidate(d, e);  // (DR candidates depend on the level through the respawn draws)
  d.metrics[e] = compute_metrics(R, s, W, true);
  if (!reset_agent(s)) d.err[e] |= kErrNoStart;
  d.hot[e] = pack(s);
  emit_direct(R, s, d.c, o, e);
}

// The same for a subset of the envs with one level each: env_index[k] gets level k (environments that regenerate their
// level on every reset, e.g. the Kruskal mazes of envs/multigrid/mst_maze.py, whose generator runs on the host).
__global__ void k_load_levels_at(Dev d, const uint8_t *enc, const int32_t *env_index, int n, int start_dir, OutPtrs o) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int e = env_index[k];
  if (e < 0 || e >= d.N) return;
  const int W = d.c.W;
  const Rows R = env_rows(d, e);
  Env s = unpack(d.hot[e]);
  const uint8_t *src = enc + (size_t)k * W * W * 3;
  s.gx = s.gy = s.sx = s.sy = kNone;
  int enc_dir = 0;
  for (int y = 0; y < W; y++) {
    uint32_t row = 0;
    for (int x = 0; x < W; x++) {
      const uint8_t t = src[((size_t)x * W + y) * 3];
      if (t == 2) row |= 1u << x;
      else if (t == 8) { s.gx = x; s.gy = y; }
      else if (t == 10) { s.sx = x; s.sy = y; enc_dir = src[((size_t)x * W + y) * 3 + 2] & 3; }
    }
    R.set(y, row);
  }
  s.sdir = (start_dir < 0 ? enc_dir : start_dir) & 3;   // start_dir < 0: the direction stored in the encoding's agent cell
  s.pending = 0;
  s.ep_ret = 0.f; s.ep_len = 0;
  spec_invalidate(d, e);
  d.metrics[e] = compute_metrics(R, s, W, true);
  if (!reset_agent(s)) d.err[e] |= kErrNoStart;
  d.hot[e] = pack(s);
  emit_direct(R, s, d.c, o, e);
}

// reset_to_level, action-string form (adversarial.py:274-283).
__global__ void k_reset_to_actions(Dev d, const int32_t *locs, int len, const int32_t *index, int n, OutPtrs o) {
  RNG_SCRATCH();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int e = index ? index[k] : k;
  if (e < 0 || e >= d.N) return;
  const Rows R = env_rows(d, e);
  Env s = unpack(d.hot[e]);
  uint32_t adv = d.adv[e], err = 0;
  int4 met;
  Rng rng = RNG_OF(d, e);
  reset_adversary(R, s, adv, met, rng, d.c);
  if (d.c.resample) adv = (adv & ~(0xfffu << 12)) | ((uint32_t)len << 12);
  for (int i = 0; i < len; i++) {
    if (step_adversary(R, s, adv, met, rng, d.c, locs[(size_t)k * len + i], err))
      if (!reset_agent(s)) err |= kErrNoStart;
  }
  rng.store();
  s.elapsed = 0;
  d.hot[e] = pack(s); d.adv[e] = adv; d.metrics[e] = met;
  if (err) d.err[e] |= err;
  emit_direct(R, s, d.c, o, e);
}

// free-cell test used by mutate_level's fallbacks: free_mask == "grid cell is None" 
__device__ __forceinline__ int count_free(const Rows &R, const Env &s, int W, int pick, int &sel) {
  int cnt = 0;
  sel = -1;
  for (int y = 1; y < W - 1; y++)
    for (int x = 1; x < W - 1; x++)
      if (is_empty(R, s, x, y)) { if (cnt == pick) sel = (y - 1) * (W - 2) + (x - 1); cnt++; }
  return cnt;
}

// mutate_level, edit phase (adversarial.py:317-368).
__global__ void k_mutate_edits(Dev d, const int32_t *locs, const int32_t *ops, const int32_t *n_edits, int max_edits,
                               uint8_t *need, int32_t *n_free) {
  RNG_SCRATCH();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const int W = d.c.W, I = W - 2;
  const Rows R = env_rows(d, e);
  Env s = unpack(d.hot[e]);
  if (s.pending) { Rng rng = RNG_OF(d, e); flush_pending(R, s, rng, W); rng.store(); }
  spec_invalidate(d, e);
  const int k = n_edits[e];
  for (int n = 0; n < k && n < max_edits; n++) {
    const int loc = locs[(size_t)e * max_edits + n], op = ops[(size_t)e * max_edits + n];
    const int x = loc % I + 1, y = loc / I + 1;
    // editor action: 0 '-', 1 '.', then ('a','g') for the 4-action set, 'g' for the 3-action set
    const bool is_a = (d.c.n_editor == 4 && op == 2), is_g = (d.c.n_editor == 4 && op == 3) || (d.c.n_editor == 3 && op == 2);
    // _clean_loc (adversarial.py:296-306)
    R.set(y, R.get(y) & ~(1u << x));
    if (x == s.gx && y == s.gy) { s.gx = s.gy = kNone; }
    else if (s.has_agent && x == s.ax && y == s.ay) { s.sx = s.sy = kNone; s.has_agent = 0; }
    if (op == 0) R.set(y, R.get(y) | (1u << x));
    else if (is_a) { s.has_agent = 1; s.ax = x; s.ay = y; s.adir = 0; s.sx = x; s.sy = y; }
    else if (is_g) { s.gx = x; s.gy = y; }
  }
  int sel;
  const bool need_g = s.gx == kNone, need_a = s.sx == kNone;
  const int nf = (need_g || need_a) ? count_free(R, s, W, -1, sel) : 0;
  if (need) { need[2 * e] = need_g; need[2 * e + 1] = need_a; }
  if (n_free) { n_free[2 * e] = need_g ? nf : 0; n_free[2 * e + 1] = need_a ? (need_g ? nf - 1 : nf) : 0; }
  d.hot[e] = pack(s);
}

// mutate_level, fallback placement + metrics + reset_agent (adversarial.py:370-397).
__global__ void k_mutate_finalize(Dev d, const int32_t *choice, OutPtrs o) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const int W = d.c.W, I = W - 2;
  const Rows R = env_rows(d, e);
  Env s = unpack(d.hot[e]);
  int sel;
  if (s.gx == kNone) {
    count_free(R, s, W, choice ? choice[2 * e] : 0, sel);
    if (sel >= 0) { s.gx = sel % I + 1; s.gy = sel / I + 1; }
  }
  if (s.sx == kNone) {
    count_free(R, s, W, choice ? choice[2 * e + 1] : 0, sel);
    if (sel >= 0) { s.sx = sel % I + 1; s.sy = sel / I + 1; s.has_agent = 1; s.ax = s.sx; s.ay = s.sy; }
  }
  s.step_count = 0;
  spec_invalidate(d, e);
  d.adv[e] = d.adv[e] & ~0xfffu;  // adversary_step_count = 0
  d.metrics[e] = compute_metrics(R, s, W, true);
  if (!reset_agent(s)) d.err[e] |= kErrNoStart;
  d.hot[e] = pack(s);
  emit_direct(R, s, d.c, o, e);
}

// AdversarialEnv.encoding (adversarial.py:162-164): u8 [N][W][W][3] indexed [x][y][c]; one thread per cell.
__global__ void k_encode(Dev d, uint8_t *enc) {
  const int W = d.c.W, WW = W * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)d.N * WW) return;
  const int e = (int)(idx / WW), cell = (int)(idx % WW), x = cell / W, y = cell % W;
  const Env s = unpack(d.hot[e]);
  uint8_t t, c, st = 0;
  if (s.has_agent && x == s.ax && y == s.ay) { t = 10; c = 0; st = (uint8_t)s.adir; }
  else if ((env_rows(d, e).get(y) >> x) & 1u) { t = 2; c = 5; }
  else if (x == s.gx && y == s.gy) { t = 8; c = 1; }
  else { t = 1; c = 0; }
  uint8_t *o = enc + idx * 3;
  o[0] = t; o[1] = c; o[2] = st;
}

// venv.get_images() (parallel_wrappers.py:187-193 -> MultiGridEnv.render(mode='level'), multigrid.py:1105-1140 ->
// Grid.render, :216-261): the level as an RGB mosaic of 32x32 tiles, out u8 [n][W*32][W*32][3].  A maze cell is empty, wall,
// goal or the agent (4 directions), highlighted or not (compute_agent_visibility_mask, multigrid.py:1071-1103: the cells of
// the agent's view that process_vis leaves visible; walls are never highlighted, render_tile :190): 14 tiles, rendered once on
// the host (dcd_isaac_b200/tiles.py) and pasted here.  One CTA per (cell row, image): the 32 pixel rows of a cell row are
// contiguous in the output, written as coalesced 32-bit words.
__global__ void __launch_bounds__(256) k_render_images(Dev d, const uint32_t *tiles /* [14][32][24] words */, const int32_t *index,
                                                       int n, uint8_t *out) {
  __shared__ int s_tile[32];
  const int W = d.c.W, y = blockIdx.x, k = blockIdx.y;
  const int e = index ? index[k] : k;
  if (e < 0 || e >= d.N) return;
  if ((int)threadIdx.x < W) {
    const int x = threadIdx.x;
    const Env s = unpack(d.hot[e]);
    const Rows R = env_rows(d, e);
    int code = 0;
    if (s.has_agent && x == s.ax && y == s.ay) code = 3 + s.adir;
    else if ((R.get(y) >> x) & 1u) code = 1;
    else if (x == s.gx && y == s.gy) code = 2;
    int hl = 0;
    if (s.has_agent && code != 1) {
      const uint32_t vis = d.c.see_through ? render_packed<true, uint64_t>(R, s, W).vis : render_packed<false, uint64_t>(R, s, W).vis;
      const int dd = s.adir, dx = x - s.ax, dy = y - s.ay;
      const bool vertical = dd & 1;
      const int p = vertical ? dy : dx, q = vertical ? dx : dy;
      const int fd = (dd == 0 || dd == 1) ? p : -p, lt = (dd == 0 || dd == 3) ? q : -q;
      const int vy = kV - 1 - fd, vx = lt + kV / 2;
      if ((unsigned)vx < (unsigned)kV && (unsigned)vy < (unsigned)kV) hl = (vis >> (vy * kV + vx)) & 1u;
    }
    s_tile[x] = 2 * code + hl;
  }
  __syncthreads();
  const int row_words = W * 24;                      // one pixel row of the image: W tiles x 96 bytes
  const size_t img_words = (size_t)W * 32 * row_words;
  uint32_t *o = reinterpret_cast<uint32_t *>(out) + (size_t)k * img_words + (size_t)y * 32 * row_words;
  for (int w = threadIdx.x; w < 32 * row_words; w += blockDim.x) {
    const int py = w / row_words, rem = w - py * row_words, x = rem / 24, off = rem - x * 24;
    o[w] = tiles[(s_tile[x] * 32 + py) * 24 + off];
  }
}

// replay deferred respawn draws so that the RNG state seen by the host is the reference's
// (min_pending > 1: the periodic sweep that keeps the 8-bit counter of very easy levels far from saturation)
__global__ void k_flush(Dev d, int min_pending) {
  RNG_SCRATCH();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  Env s = unpack(d.hot[e]);
  if (s.pending < min_pending) return;
  const Rows R = env_rows(d, e);
  Rng rng = RNG_OF(d, e);
  flush_pending(R, s, rng, d.c.W);
  rng.store();
  d.hot[e] = pack(s);
}

__global__ void k_get_metrics(Dev d, int32_t *m) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  int4 v = d.metrics[e];
  if (v.z == kMetricsDirty) {  // level taken from a speculative candidate record: metrics are computed on demand
    v = compute_metrics(env_rows(d, e), unpack(d.hot[e]), d.c.W, true);
    d.metrics[e] = v;
  }
  m[4 * e] = v.x; m[4 * e + 1] = v.y; m[4 * e + 2] = v.z; m[4 * e + 3] = v.w;
}
__global__ void k_get_agent_state(Dev d, int32_t *o) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const Env s = unpack(d.hot[e]);
  const uint32_t adv = d.adv[e];
  int32_t *p = o + 8 * (size_t)e;
  p[0] = s.ax; p[1] = s.ay; p[2] = s.adir; p[3] = s.step_count; p[4] = s.elapsed;
  p[5] = adv & 0xfff; p[6] = (adv >> 12) & 0xfff; p[7] = (int32_t)d.words[e];
}
__global__ void k_get_errors(Dev d, uint32_t *o, int clear) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  o[e] = d.err[e];
  if (clear) d.err[e] = 0;
}

// ------------------------------------------------------------------------------------------ hot kernel: step_env
struct StepArgs {
  const int64_t *action;
  const uint8_t *action_u8;      // narrow action stream (one byte per env) used instead of `action` when not NULL
  const int32_t *n_walls;
  int last_step;  // bit0: last rollout step (adversarial_runner.py:521-530), bit1: use_proper_time_limits
  mgplr_step_out o;
  uint32_t *done_count;          // host-driven step: append-list of finished episodes (NULL otherwise)
  uint32_t *done_count_next;     // the counter of the NEXT host-driven step, zeroed by this launch
  mgplr_done_record *done_list;  // device view of pinned host memory: records cross PCIe as posted writes
  uint8_t *flags_host;           // second flags destination (mapped pinned host memory) or NULL
  int spec;                      // DR auto-reset: speculative next-level candidates enabled (mgplr_env.cuh "SPECULATION")
  int stream_scalars;            // A/B (MGPLR_L2_HINTS bit 2): streaming stores for the per-env scalars, streaming action loads
};

// ---- rare paths, kept out of line with by-value arguments so the common path stays in registers ----

// agent_is_done's respawn (multigrid.py:821-838): `count` random empty cells drawn with the env RNG (the deferred
// ones first, see Env::pending).  Returns the last one as x | y<<8.
__device__ __noinline__ uint32_t rare_respawn(uint32_t *mt, uint32_t *mti, uint32_t *words, uint32_t *spec, int N, int e, uint32_t *rows,
                                              int stride, int W, int gx, int gy, int count) {
  RngSlow rng(mt, mti, words, spec, N, e);
  const uint32_t p = replay_respawns(Rows{rows, stride}, gx, gy, rng, W, count);
  rng.store();
  return p;
}

// info['truncated_obs'] (time_limit.py:29-31) / runner cliffhanger obs (adversarial_runner.py:523-528)
__device__ __noinline__ void rare_emit_trunc(uint32_t *rows, int stride, uint4 hot, int W, int see, float *image, float *direction,
                                             int e, float *full = nullptr) {
  const Env s = unpack(hot);
  const Rows R{rows, stride};
  const View v = see ? render_view<true>(R, s, W) : render_view<false>(R, s, W);
  if (image) emit_obs_f32(v, image + (size_t)e * kObsFloats);
  if (direction) direction[e] = (float)s.adir;
  if (full) {  // MultiGridFullyObsWrapper.agent_observation on the pre-reset state: raw codes, [3][W][W] indexed [c][x][y]
    float *o = full + (size_t)e * 3 * W * W;
    for (int x = 0; x < W; x++)
      for (int y = 0; y < W; y++) {
        float t, c, st = 0.f;
        if (s.has_agent && x == s.ax && y == s.ay) { t = 10.f; c = 0.f; st = (float)s.adir; }
        else if ((R.get(y) >> x) & 1u) { t = 2.f; c = 5.f; }
        else if (x == s.gx && y == s.gy) { t = 8.f; c = 1.f; }
        else { t = 1.f; c = 0.f; }
        o[x * W + y] = t; o[W * W + x * W + y] = c; o[2 * W * W + x * W + y] = st;
      }
  }
}

// worker.step_env's reset_random branch (parallel_wrappers.py:30-33)
__device__ __noinline__ uint4 rare_reset_random(Dev d, uint32_t *rows, int stride, uint4 hot, int e, int n_walls, uint32_t *rng_col,
                                                int rng_stride) {
  Env s = unpack(hot);
  uint32_t adv = d.adv[e], err = 0;
  int4 met;
  Rng rng(d, e, rng_col, rng_stride);
  reset_random(Rows{rows, stride}, s, adv, met, rng, d, e, n_walls, err);
  rng.store();
  d.adv[e] = adv; d.metrics[e] = met;
  if (err) d.err[e] |= err;
  return pack(s);
}

__device__ __noinline__ void rare_emit_u8(uint32_t *rows, int stride, uint4 hot, int W, int see, uint8_t *image_u8, int e);

// One env transition for the thread's env; `rows` = this env's wall rows in shared memory.  Returns flags.
// In reset_agent mode a goal's respawn draw is DEFERRED (Env::pending, mgplr_env.cuh).
template <bool SEE, bool RR, typename EXT>
__device__ __forceinline__ uint32_t step_one(const Dev &d, uint32_t *rows, int stride, Env &s, int e, int a, const StepArgs &A,
                                             float *s_obs, float &rew_out, bool &rows_dirty, uint32_t *rng_col, int rng_stride) {
  const Cfg &c = d.c;
  const Rows R{rows, stride};
  uint32_t flags = 0;
  double rew = 0.0;
  const bool want_trunc = A.o.trunc_image || A.o.trunc_direction || A.o.trunc_full_obs;
  // MultiGridEnv.step / step_one_agent (multigrid.py:943-975,866-941)
  s.step_count++;
  const int fx = s.ax + ((s.adir == 0) - (s.adir == 2)), fy = s.ay + ((s.adir == 1) - (s.adir == 3));
  if (a == 0) s.adir = (s.adir + 3) & 3;
  else if (a == 1) s.adir = (s.adir + 1) & 3;
  else if (a == 2) {
    if (fx == s.gx && fy == s.gy) {
      // agent_is_done: remove the agent, done, respawn (dir forced to 0, multigrid.py:668-672); reward = _reward()
      s.done_flag = 1;
      if (RR || (want_trunc && s.elapsed + 1 >= c.max_episode_steps) || s.pending >= kMaxPending) {
        const uint32_t p = rare_respawn(d.mt, d.mti, d.words, d.spec, d.N, e, rows, stride, c.W, s.gx, s.gy, s.pending + 1);
        s.pending = 0; s.ax = p & 0xff; s.ay = p >> 8; s.adir = 0;
      } else s.pending++;
      rew = __dsub_rn(1.0, __dmul_rn(0.9, __ddiv_rn((double)s.step_count, (double)c.max_steps)));
      flags |= MGPLR_F_GOAL;
    } else if (!is_wall(R, fx, fy)) { s.ax = fx; s.ay = fy; }
  }
  bool done = s.done_flag || s.step_count >= c.max_steps;
  // TimeLimit.step (time_limit.py:24-33)
  s.elapsed++;
  if (s.elapsed >= c.max_episode_steps) {
    flags |= MGPLR_F_TRUNC_KEY | (done ? 0u : MGPLR_F_TRUNC_VAL);
    if (want_trunc) rare_emit_trunc(rows, stride, pack(s), c.W, c.see_through, A.o.trunc_image, A.o.trunc_direction, e, A.o.trunc_full_obs);
    done = true;
  }
  // VecMonitor.step_wait (vec_monitor.py:60-85): eprets(f32) += rews(f64) is evaluated in double
  if (flags & MGPLR_F_GOAL) s.ep_ret = (float)__dadd_rn((double)s.ep_ret, rew);
  s.ep_len += 1;
  if (done) {
    flags |= MGPLR_F_DONE;
    if (A.o.ep_return) A.o.ep_return[e] = s.ep_ret;
    if (A.o.ep_length) A.o.ep_length[e] = s.ep_len;
    s.ep_ret = 0.f; s.ep_len = 0;
    // worker.step_env (parallel_wrappers.py:27-37)
    if (RR) {
      s = unpack(rare_reset_random(d, rows, stride, pack(s), e, (c.resample && A.n_walls) ? A.n_walls[e] : -1, rng_col, rng_stride));
      rows_dirty = true;
    } else if (!reset_agent(s)) d.err[e] |= kErrNoStart;
  } else if ((A.last_step & 3) == 3 && want_trunc) {
    rare_emit_trunc(rows, stride, pack(s), c.W, c.see_through, A.o.trunc_image, A.o.trunc_direction, e, A.o.trunc_full_obs);
  }
  const PackedView v = render_packed<SEE, EXT>(R, s, c.W);
  emit_packed_f32<SEE, false>(v, s_obs);
  if (A.o.image_u8) rare_emit_u8(rows, stride, pack(s), c.W, c.see_through, A.o.image_u8, e);
  if ((flags & MGPLR_F_DONE) && d.err[e]) flags |= MGPLR_F_ERROR;
  rew_out = (float)rew;
  return flags;
}

__device__ __forceinline__ void write_step_scalars(const StepArgs &A, int e, const Env &s, uint32_t flags, float rew,
                                                   float ep_ret = 0.f, int ep_len = 0) {
  const mgplr_step_out &o = A.o;
  if (A.done_count && (flags & MGPLR_F_DONE)) {
    const uint32_t k = atomicAdd(A.done_count, 1u);
    mgplr_done_record r;
    r.env = e; r.reward = rew; r.ep_return = ep_ret; r.ep_length = (int32_t)((uint32_t)ep_len | (flags << 24));
    A.done_list[k] = r;
  }
  const bool done = flags & MGPLR_F_DONE, last = A.last_step & 1, cliff = last && (A.last_step & 2) && !done;
  if (A.stream_scalars) {   // (A/B, MGPLR_L2_HINTS bit 2) written once, read much later: streaming stores
    if (o.direction) __stcs(&o.direction[e], (float)s.adir);
    if (o.reward) __stcs(&o.reward[e], rew);
    if (o.flags) __stcs(reinterpret_cast<unsigned char *>(&o.flags[e]), (unsigned char)flags);
    if (A.flags_host) A.flags_host[e] = (uint8_t)flags;
    if (o.masks) __stcs(&o.masks[e], (done || last) ? 0.f : 1.f);
    if (o.bad_masks) __stcs(&o.bad_masks[e], ((flags & MGPLR_F_TRUNC_KEY) || cliff) ? 0.f : 1.f);
    if (o.cliffhanger_masks) __stcs(&o.cliffhanger_masks[e], cliff ? 0.f : 1.f);
    return;
  }
  if (o.direction) o.direction[e] = (float)s.adir;
  if (o.reward) o.reward[e] = rew;
  if (o.flags) o.flags[e] = (uint8_t)flags;
  if (A.flags_host) A.flags_host[e] = (uint8_t)flags;
  if (o.masks) o.masks[e] = (done || last) ? 0.f : 1.f;
  if (o.bad_masks) o.bad_masks[e] = ((flags & MGPLR_F_TRUNC_KEY) || cliff) ? 0.f : 1.f;
  if (o.cliffhanger_masks) o.cliffhanger_masks[e] = cliff ? 0.f : 1.f;
}

constexpr int kWarpTile = 32;  // envs per warp tile = lanes

// ---- TMA bulk copies + mbarrier (PTX; SASS: UBLKCP / SYNCS) ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk asynchronous copy completing on an mbarrier; bytes % 16 == 0, 16-byte aligned addresses
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// L2 eviction-priority policies (createpolicy) and hinted variants of the copies.  The level (wall bit-plane) and the
// hot records are re-read by every step launch of a rollout and are small (a few MB): they are kept in the 126 MB L2
// (evict_last) while the observation stream, written once and read much later, is marked evict_first so that it does
// not push them out.
__device__ __forceinline__ uint64_t l2_policy(int kind) {  // 0 normal, 1 evict_last, 2 evict_first
  uint64_t p;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_load_hint(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void *gdst, const void *ssrc, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ uint4 ld_hint_u4(const uint4 *p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_hint_u4(uint4 *p, const uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
// Ampere-style asynchronous copies (LDGSTS): 16 bytes per lane, no uniform-datapath instruction on the issuing warp
__device__ __forceinline__ void cp_async16_hint(void *sdst, const void *gsrc, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(sdst)), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory"); }
__device__ __forceinline__ void st_hint_f4(float4 *p, const float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk stores have finished READING shared memory (the global writes complete before the grid does)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Stage the wall rows of the CTA's TILE/32 warp-tiles into s_rows[sub][W][32]: one bulk copy of W*128 bytes per
// sub-tile (one elected thread, completion on `bar`).  The HBM plane is padded to whole tiles, so partial tiles
// take the same path.
template <int TILE>
__device__ __forceinline__ void stage_rows_begin(const Dev &d, uint32_t *s_rows, uint64_t *bar, int base) {
  const int W = d.c.W;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_proxy_async_smem();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)(W * kWarpTile * 4);
    const int n_tiles = (d.N + kWarpTile - 1) / kWarpTile;
    int subs = 0;
    for (int sub = 0; sub < TILE / kWarpTile; sub++) subs += ((base / kWarpTile + sub) < n_tiles);
    mbar_expect_tx(bar, bytes * subs);
    for (int sub = 0; sub < subs; sub++)
      bulk_load(s_rows + sub * W * kWarpTile, d.wall + ((size_t)(base / kWarpTile + sub)) * W * kWarpTile, bytes, bar);
  }
}

// Store the CTA's observation tile: n_envs*75 contiguous floats starting at gdst.  All threads must call.
__device__ __forceinline__ void store_obs_tile(float *gdst, const float *s_obs, int n_envs, bool wait_now) {
  const uint32_t bytes = (uint32_t)n_envs * kObsFloats * 4u;
  if ((bytes & 15u) == 0 && (((uintptr_t)gdst) & 15u) == 0) {
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_store(gdst, s_obs, bytes);
      bulk_commit();
      if (wait_now) bulk_wait_read0();
    }
  } else {
    __syncthreads();
    for (int i = threadIdx.x; i < n_envs * kObsFloats; i += blockDim.x) gdst[i] = s_obs[i];
  }
}

// ---- hot kernel: persistent, warp-pipelined step_env -------------------------------------------------------
// Each WARP owns tiles of 32 consecutive envs (tile = global_warp + k * total_warps; handed out by tickets in the
// speculative DR variant) and runs a software pipeline with no CTA-level synchronisation:
//   * the wall rows of tile k+1 are prefetched by ONE bulk asynchronous copy (TMA, W*128 B) into the other
//     half of a double buffer while tile k computes; completion is signalled on a per-warp mbarrier;
//   * the hot record / action of tile k+1 are prefetched into registers;
//   * the 32x75 float32 observations of tile k leave shared memory as one 9600-byte bulk asynchronous store
//     that drains while tile k+1 steps and renders; the single obs buffer is only re-acquired
//     (cp.async.bulk.wait_group.read) right before tile k+1 emits.
// The DR variant (RR) additionally resets finished envs -- by copying a pre-built candidate level, or by rebuilding in
// the kernel -- and runs the regeneration jobs queued by the previous launch.
// per warp: obs tile | two row buffers | two mbarriers | (DR variant) batched-RNG scratch [32][32] | job queue
constexpr int kPendCap = 96;  // DR variant: regeneration jobs a warp collects before it appends them to the global list
__host__ __device__ inline size_t warp_smem_bytes(int W, bool rr) {
  return ((size_t)kWarpTile * kObsFloats * 4 + 2 * (size_t)W * kWarpTile * 4 + 16 + (rr ? 32 * kWarpTile * 4 + kPendCap * 8 : 0) + 127) &
         ~(size_t)127;
}
// append a warp's collected jobs (one entry per reset env) to the launch's list: ONE atomic per flush, both candidates
__device__ __forceinline__ void flush_pending_jobs(const Dev &d, int rr_par, const uint2 *s_pend, int &pend_n, int lane) {
  __syncwarp();
  if (pend_n) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&d.sched[rr_par], 2u * (uint32_t)pend_n);
    base = __shfl_sync(0xffffffffu, base, 0);
    uint2 *list = d.rr_list + (size_t)rr_par * 2 * d.N + base;
    for (int i = lane; i < pend_n; i += 32) {
      const uint2 jb = s_pend[i];
      list[2 * i] = jb; list[2 * i + 1] = make_uint2(jb.x | 128u, jb.y);
    }
    pend_n = 0;
  }
  __syncwarp();
}

__device__ __noinline__ void rare_emit_u8(uint32_t *rows, int stride, uint4 hot, int W, int see, uint8_t *image_u8, int e) {
  const Env s = unpack(hot);
  const Rows R{rows, stride};
  const View v = see ? render_view<true>(R, s, W) : render_view<false>(R, s, W);
  emit_obs_u8(v, image_u8 + (size_t)e * kObsFloats);
}

// rows of one 32-env tile -> shared memory: ONE bulk asynchronous copy of W*128 bytes completing on `bar`
__device__ __forceinline__ void warp_issue_rows(const Dev &d, uint32_t *s_rows, uint64_t *bar, int tile, int lane, uint64_t pol) {
  if (!d.use_tma) {  // W*128 bytes = 8W 16-byte chunks, lanes in turn (coalesced); one commit group per tile
    const uint4 *src = reinterpret_cast<const uint4 *>(d.wall + (size_t)tile * d.c.W * kWarpTile);
    uint4 *dst = reinterpret_cast<uint4 *>(s_rows);
    for (int i = lane; i < 8 * d.c.W; i += 32) cp_async16_hint(dst + i, src + i, pol);
    cp_async_commit();
    return;
  }
  if (lane == 0) {
    const uint32_t bytes = (uint32_t)(d.c.W * kWarpTile * 4);
    mbar_expect_tx(bar, bytes);
    bulk_load_hint(s_rows, d.wall + (size_t)tile * d.c.W * kWarpTile, bytes, bar, pol);
  }
}

// reset_agent mode: the shared-memory bound is 4 CTAs/SM for W <= 24 (<= 128 registers) and 3 for wider grids; the DR
// variant carries its in-kernel reset_random, the batched-RNG scratch and the regeneration phase (168 registers, 3 CTAs/SM).
template <bool SEE, bool RR, typename EXT>
__global__ void __launch_bounds__(128, RR ? 3 : (sizeof(EXT) == 8 ? 3 : 4)) k_step_env(Dev d, StepArgs A, int tile0, int n_tiles) {  // tiles [tile0, n_tiles)
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = d.c.W, N = d.N;
  const int wpc = blockDim.x >> 5;
  uint8_t *wbase = smem + (size_t)warp * warp_smem_bytes(W, RR);
  float *s_obs = reinterpret_cast<float *>(wbase);
  uint32_t *s_rows = reinterpret_cast<uint32_t *>(wbase + (size_t)kWarpTile * kObsFloats * 4);
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_rows + 2 * W * kWarpTile);
  uint32_t *s_rng = reinterpret_cast<uint32_t *>(bars + 2);  // only present (and used) when RR
  uint2 *s_pend = reinterpret_cast<uint2 *>(s_rng + 32 * kWarpTile);  // (RR) jobs collected by this warp
  int pend_n = 0;
  unsigned long long prof_start = 0;
  int prof_commits = 0;
  if (RR && d.prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_start));
  const int total = gridDim.x * wpc;
  // Static round-robin tile assignment (tile = global_warp + k * total_warps).  A dynamic scheduler on one global
  // atomic counter was measured: ~9 k same-address atomics per launch doubled the launch time (10 -> 20 us).
  int tile = tile0 + blockIdx.x * wpc + warp;
  // Programmatic dependent launch: consecutive step launches of a rollout are serially dependent through the hot
  // records, so the NEXT launch is allowed to become resident as this one's CTAs retire and to run its on-chip
  // prologue; it blocks at griddepcontrol.wait (below) until this grid has completed and its writes are visible.
  asm volatile("griddepcontrol.launch_dependents;");
  if (!RR && tile >= n_tiles) return;  // (the DR variant's idle warps still serve the regeneration phase)
  if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_proxy_async_smem(); }
  const bool use_spec = RR && A.spec;
  const int rr_par = use_spec ? (int)(*(volatile uint32_t *)&d.sched[5] & 1u) : 0;  // stable for the whole launch
  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything above touched only on-chip state
  if (RR && use_spec) {
    // ---- regeneration jobs FIRST: the jobs queued by the previous launch (list of the other parity; its length is
    // final) are the long indivisible items of this launch (~15 us each), so warps take them before any tile; the warps
    // that find the list empty start on the tiles at once, and because the tiles are handed out dynamically below, the
    // job warps simply end up with fewer tiles.  Nothing waits on this launch's own progress.
    const int q = rr_par ^ 1;
    const uint32_t n_jobs = *(volatile uint32_t *)&d.sched[q];
    const uint2 *list = d.rr_list + (size_t)q * 2 * N;
    uint32_t *scr = reinterpret_cast<uint32_t *>(s_obs);  // the observation tile is not in use yet
    if (n_jobs >= 8u * (uint32_t)total) {
      // a storm's worth of jobs: 32 candidates per warp side by side (rr_regen_job_lane), tickets in units of 32
      for (;;) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(&d.sched[2 + q], 32u);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= n_jobs) break;
        if (j + lane < n_jobs) rr_regen_job_lane(d, __ldcg(list + j + lane), scr + lane);
        __syncwarp();
      }
    } else
    for (;;) {
      uint32_t j = 0;
      if (lane == 0) j = atomicAdd(&d.sched[2 + q], 1u);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= n_jobs) break;
      const long long c0 = d.prof ? clock64() : 0;
      rr_regen_job(d, __ldcg(list + j), lane, scr);
      __syncwarp();
      if (d.prof && lane == 0) {
        const unsigned long long dt = (unsigned long long)(clock64() - c0);
        atomicAdd(&d.prof[3], dt); atomicAdd(&d.prof[4], 1ull); atomicMax(&d.prof[5], dt);
      }
    }
    if (d.prof && lane == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      atomicMax(&d.prof[2], t1);               // end of this warp's job phase
    }
    __syncwarp();
  }
  // the state plane of every observation is all zero: written once here, never touched by emit_packed_f32
#pragma unroll
  for (int i = 0; i < kV * kV; i++) s_obs[lane * kObsFloats + 2 * kV * kV + i] = 0.0f;
  __syncwarp();
  const uint64_t pol_keep = l2_policy((d.l2_hints & 16) ? 2 : ((d.l2_hints & 1) && !(d.l2_hints & 32)) ? 1 : 0), pol_stream = l2_policy((d.l2_hints & 1) && !(d.l2_hints & 8) ? 2 : 0);
  const uint64_t pol_hot = (d.l2_hints & 64) ? l2_policy(2) : (d.l2_hints & 128) ? l2_policy(0) : pol_keep;
  if (A.done_count_next && blockIdx.x == 0 && threadIdx.x == 0) *A.done_count_next = 0;
  // Tile assignment: static round-robin (tile = global_warp + k * total_warps) -- a dynamic scheduler on one global
  // counter doubled the reset_agent variant's launch time -- EXCEPT in the speculative DR variant, where warps carry very
  // different loads (jobs, commits): there the next tile is a ticket, requested one tile ahead so that the atomic's
  // round trip hides behind a tile's work.
  const bool dyn = RR && use_spec;
  uint32_t ticket = 0;  // (lane 0) ticket of the tile after `next`
  int next;
  if (dyn) {
    uint32_t t0 = 0, t1 = 0;
    if (lane == 0) { t0 = atomicAdd(&d.sched[6], 1u); t1 = atomicAdd(&d.sched[6], 1u); }
    tile = tile0 + (int)__shfl_sync(0xffffffffu, t0, 0);
    next = tile0 + (int)__shfl_sync(0xffffffffu, t1, 0);
  } else {
    next = tile + total;
  }
  // prologue: first tile's rows and scalars
  if (tile < n_tiles) warp_issue_rows(d, s_rows, &bars[0], tile, lane, pol_keep);
  uint4 nh = make_uint4(0, 0, 0, 0);
  int na = 6;
  uint32_t nsp = 0;  // speculation word of the env (DR variant)
  if (tile < n_tiles && tile * kWarpTile + lane < N) {
    nh = ld_hint_u4(&d.hot[tile * kWarpTile + lane], pol_hot);
    na = A.action_u8 ? (int)A.action_u8[tile * kWarpTile + lane] : (int)(A.stream_scalars ? __ldcs(&A.action[tile * kWarpTile + lane]) : A.action[tile * kWarpTile + lane]);
    if (use_spec) nsp = d.spec[tile * kWarpTile + lane];
  }
  uint32_t phase = 0;  // bit s = parity to wait for on bars[s]
  for (int k = 0; tile < n_tiles; k++) {
    const int st = k & 1, base = tile * kWarpTile, e = base + lane;
    const int n_tile = min(kWarpTile, N - base);
    const bool valid = lane < n_tile;
    const uint4 h = nh;
    const int a = na;
    const uint32_t sp = nsp;
    uint32_t *rows = s_rows + st * W * kWarpTile;
    // prefetch the next tile into the other stage (its last readers, the previous tile, are done: __syncwarp)
    __syncwarp();
    if (dyn && lane == 0) ticket = atomicAdd(&d.sched[6], 1u);  // the tile after `next`: consumed at the end of this iteration
    if (next < n_tiles) {
      warp_issue_rows(d, s_rows + (st ^ 1) * W * kWarpTile, &bars[st ^ 1], next, lane, pol_keep);
      if (next * kWarpTile + lane < N) {
        nh = ld_hint_u4(&d.hot[next * kWarpTile + lane], pol_hot);
        na = A.action_u8 ? (int)A.action_u8[next * kWarpTile + lane] : (int)(A.stream_scalars ? __ldcs(&A.action[next * kWarpTile + lane]) : A.action[next * kWarpTile + lane]);
        if (use_spec) nsp = d.spec[next * kWarpTile + lane];
      }
    }
    if (d.use_tma) {
      mbar_wait(&bars[st], (phase >> st) & 1u);
      phase ^= 1u << st;
    } else {  // this tile's group is the older of at most two pending ones
      if (next < n_tiles) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncwarp();
    }

    const long long pc0 = (RR && d.prof) ? clock64() : 0;
    Env s = unpack(h);
    const Cfg &c = d.c;
    const Rows R{rows + lane, kWarpTile};
    uint32_t flags = 0;
    double rew = 0.0;
    bool dirty = false, need_rr = false;
    int cand_pick = -1;  // DR variant: candidate record to reset from (0 no goal / 1 goal), -1 = none
    int cand_used = 0;   // MT words that record consumed
    uint32_t cand_idx = 0, cand_words_used = 0, new_idx = 0;  // MT cursor / words drawn before and cursor after the reset
    float fin_ret = 0.f;
    int fin_len = 0;
    if (valid) {
      const bool want_trunc = A.o.trunc_image || A.o.trunc_direction || A.o.trunc_full_obs;
      // MultiGridEnv.step / step_one_agent (multigrid.py:943-975,866-941)
      s.step_count++;
      const int fx = s.ax + ((s.adir == 0) - (s.adir == 2)), fy = s.ay + ((s.adir == 1) - (s.adir == 3));
      if (a == 0) s.adir = (s.adir + 3) & 3;
      else if (a == 1) s.adir = (s.adir + 1) & 3;
      else if (a == 2) {
        if (fx == s.gx && fy == s.gy) {
          // agent_is_done (multigrid.py:821-838): done; the respawn draw is deferred unless something observes it
          s.done_flag = 1;
          // DR variant with a valid goal candidate: its record already accounts for the respawn draws (SPECULATION)
          const bool observed = want_trunc && s.elapsed + 1 >= c.max_episode_steps;
          if (use_spec && (sp & spec_valid_bit(spec_epoch(sp), 1)) && !observed && s.pending == 0) cand_pick = 1;
          else if (RR || observed || s.pending >= kMaxPending) {
            const uint32_t p = rare_respawn(d.mt, d.mti, d.words, d.spec, N, e, rows + lane, kWarpTile, W, s.gx, s.gy, s.pending + 1);
            s.pending = 0; s.ax = p & 0xff; s.ay = p >> 8; s.adir = 0;
          } else s.pending++;
          rew = __dsub_rn(1.0, __dmul_rn(0.9, __ddiv_rn((double)s.step_count, (double)c.max_steps)));  // _reward()
          flags |= MGPLR_F_GOAL;
        } else if (!is_wall(R, fx, fy)) { s.ax = fx; s.ay = fy; }
      }
      bool done = s.done_flag || s.step_count >= c.max_steps;
      // TimeLimit.step (time_limit.py:24-33)
      s.elapsed++;
      if (s.elapsed >= c.max_episode_steps) {
        flags |= MGPLR_F_TRUNC_KEY | (done ? 0u : MGPLR_F_TRUNC_VAL);
        if (want_trunc) rare_emit_trunc(rows + lane, kWarpTile, pack(s), W, c.see_through, A.o.trunc_image, A.o.trunc_direction, e, A.o.trunc_full_obs);
        done = true;
      }
      // VecMonitor.step_wait (vec_monitor.py:60-85): eprets(f32) += rews(f64) is evaluated in double
      if (flags & MGPLR_F_GOAL) s.ep_ret = (float)__dadd_rn((double)s.ep_ret, rew);
      s.ep_len += 1;
      if (done) {
        flags |= MGPLR_F_DONE;
        fin_ret = s.ep_ret; fin_len = s.ep_len;
        if (A.o.ep_return) A.o.ep_return[e] = s.ep_ret;
        if (A.o.ep_length) A.o.ep_length[e] = s.ep_len;
        s.ep_ret = 0.f; s.ep_len = 0;
        if (RR) {  // worker.step_env (parallel_wrappers.py:27-37): reset_random
          if (use_spec && !(flags & MGPLR_F_GOAL) && (sp & spec_valid_bit(spec_epoch(sp), 0)) && s.pending == 0) cand_pick = 0;
          if (cand_pick >= 0) {
            // take the pre-built successor level: goal / start (rows and the MT advance are done warp-wide below)
            const uint32_t *rec = cand_record(d, e, spec_epoch(sp), cand_pick) + W;  // (L2 loads: written by another SM)
            const uint32_t gs = __ldcg(rec), cerr = __ldcg(rec + 6);
            cand_used = (int)__ldcg(rec + 5);
            cand_idx = d.mti[e]; cand_words_used = d.words[e];  // (same round trip as the record fields)
            s.gx = gs & 31; s.gy = (gs >> 5) & 31; s.sx = (gs >> 11) & 31; s.sy = (gs >> 16) & 31; s.sdir = (gs >> 22) & 3;
            d.metrics[e] = make_int4(0, 0, kMetricsDirty, 0);  // recomputed on demand by mgplr_get_metrics
            d.adv[e] &= ~0xfffu;  // adversary_step_count = 0 (adversarial.py:546)
            if (cerr) d.err[e] |= cerr;
            reset_agent(s);
            dirty = true;
          } else need_rr = true;  // no candidate: rebuilt below, warp-converged
        } else if (!reset_agent(s)) d.err[e] |= kErrNoStart;
      } else if ((A.last_step & 3) == 3 && want_trunc) {
        rare_emit_trunc(rows + lane, kWarpTile, pack(s), W, c.see_through, A.o.trunc_image, A.o.trunc_direction, e, A.o.trunc_full_obs);
      }
    }
    const long long pc1 = (RR && d.prof) ? clock64() : 0;
    if (RR) {
      // committed candidate records, all lanes together per finished env: one coalesced W-word read of the rows, and
      // the words the record consumed are applied to the env's MT state (so regeneration jobs only read env state)
      for (unsigned rest = __ballot_sync(0xffffffffu, cand_pick >= 0); rest; rest &= rest - 1) {
        const int le = __ffs(rest) - 1, env = base + le;
        const int pick = __shfl_sync(0xffffffffu, cand_pick, le), used_w = __shfl_sync(0xffffffffu, cand_used, le);
        const uint32_t *rec = cand_record(d, env, spec_epoch(__shfl_sync(0xffffffffu, sp, le)), pick);
        if (lane < W) rows[lane * kWarpTile + le] = __ldcg(rec + lane);
        uint32_t idx = __shfl_sync(0xffffffffu, cand_idx, le), used = __shfl_sync(0xffffffffu, cand_words_used, le);
        // the record carries the new MT state words of the span it consumed: coalesced reads, fire-and-forget stores
        {
          uint32_t nw[kSpecState / 32];
#pragma unroll
          for (int i = 0; i < kSpecState / 32; i++) nw[i] = __ldcg(rec + W + 8 + lane + 32 * i);
#pragma unroll
          for (int i = 0; i < kSpecState / 32; i++) {
            const int j = lane + 32 * i;
            if (j < used_w) {
              uint32_t p = idx + j;
              if (p >= 624) p -= 624;
              d.mt[mt_at(env, p)] = nw[i];
            }
          }
          idx += used_w;
          if (idx >= 624) idx -= 624;
          used += used_w;
        }
        if (lane == 0) { d.mti[env] = idx; d.words[env] = used; }
        if (lane == le) new_idx = idx;
      }
      __syncwarp();
      const long long pc2 = d.prof ? clock64() : 0;
      if (d.prof) {
        const unsigned cm = __ballot_sync(0xffffffffu, cand_pick >= 0);
        if (cm && lane == 0) { atomicAdd(&d.prof[16], (unsigned long long)(pc1 - pc0)); atomicAdd(&d.prof[17], (unsigned long long)(pc2 - pc1)); atomicAdd(&d.prof[18], 1ull); }
      }
      // Finished envs without a candidate get a fresh random level the slow way.  Few per tile: the warp rebuilds them
      // one at a time cooperatively; many (a synchronized time-limit storm): every lane rebuilds its own.
      const unsigned m = __ballot_sync(0xffffffffu, need_rr);
      if (d.prof && need_rr) atomicAdd(&d.prof[9], 1ull);
      if (d.prof && cand_pick >= 0) atomicAdd(&d.prof[8], 1ull);
      if (d.prof) prof_commits += __popc(__ballot_sync(0xffffffffu, cand_pick >= 0 || need_rr));
      if (m) {
        dirty = dirty || need_rr;
        if (__popc(m) > 10) {
          if (need_rr)
            s = unpack(rare_reset_random(d, rows + lane, kWarpTile, pack(s), e, (c.resample && A.n_walls) ? A.n_walls[e] : -1,
                                         s_rng + lane, kWarpTile));
        } else {
          const uint4 mine = pack(s);
          for (unsigned rest = m; rest; rest &= rest - 1) {
            const int le = __ffs(rest) - 1;
            uint4 h2;
            h2.x = __shfl_sync(0xffffffffu, mine.x, le); h2.y = __shfl_sync(0xffffffffu, mine.y, le);
            h2.z = __shfl_sync(0xffffffffu, mine.z, le); h2.w = __shfl_sync(0xffffffffu, mine.w, le);
            const int env = base + le;
            const int nw = (c.resample && A.n_walls) ? A.n_walls[env] : -1;
            const uint4 r = coop_reset_random(d, rows + le, kWarpTile, h2, env, nw, lane);
            if (lane == le) s = unpack(r);
          }
        }
      }
    }
    const PackedView v = render_packed<SEE, EXT>(R, s, W);
    // re-acquire the observation buffer: the previous tile's bulk store must have finished reading it
    if (d.use_tma && lane == 0) bulk_wait_read0();
    __syncwarp();
    emit_packed_f32<SEE, false>(v, s_obs + lane * kObsFloats);
    if (A.o.image) {
      float *gdst = A.o.image + (size_t)base * kObsFloats;
      if (n_tile == kWarpTile && (((uintptr_t)gdst) & 15u) == 0 && !d.use_tma) {
        // 600 consecutive 16-byte chunks, lanes in turn: conflict-free LDS.128, fully coalesced STG.128
        __syncwarp();
        const float4 *src = reinterpret_cast<const float4 *>(s_obs);
        float4 *dst = reinterpret_cast<float4 *>(gdst);
#pragma unroll 5
        for (int i = lane; i < kWarpTile * kObsFloats / 4; i += 32) st_hint_f4(dst + i, src[i], pol_stream);
        __syncwarp();
      } else if (n_tile == kWarpTile && (((uintptr_t)gdst) & 15u) == 0) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { bulk_store_hint(gdst, s_obs, kWarpTile * kObsFloats * 4, pol_stream); bulk_commit(); }
      } else {
        __syncwarp();
        for (int i = lane; i < n_tile * kObsFloats; i += 32) gdst[i] = s_obs[i];
        __syncwarp();
      }
    }
    bool queue_me = false;
    uint2 queue_job = make_uint2(0, 0);
    const long long pc3 = (RR && d.prof) ? clock64() : 0;
    if (valid) {
      st_hint_u4(&d.hot[e], pack(s), pol_hot);
      if ((flags & MGPLR_F_DONE) && d.err[e]) flags |= MGPLR_F_ERROR;  // an auto-reset failed: the host raises (mgplr_get_errors)
      write_step_scalars(A, e, s, flags, (float)rew, fin_ret, fin_len);
      if (A.o.image_u8) rare_emit_u8(rows + lane, kWarpTile, pack(s), W, c.see_through, A.o.image_u8, e);
      if (RR && dirty) {
        const Rows G = env_rows(d, e);
        for (int r = 0; r < W; r++) G.set(r, rows[r * kWarpTile + lane]);
        if (use_spec) {  // new level epoch (drops the old candidates); its two candidates are built during the next launch
          const uint32_t ne = (spec_epoch(sp) + 1u) & 127u;
          *(volatile uint32_t *)&d.spec[e] = ne << kSpecEpochShift;
          if (cand_pick < 0) new_idx = *(volatile uint32_t *)&d.mti[e];  // rebuilt the slow way: cursor as stored by the rebuild
          queue_me = true;
          queue_job = make_uint2(((uint32_t)e << 8) | ne, new_idx);
        }
      }
    }
    if (RR && use_spec) {
      const unsigned qm = __ballot_sync(0xffffffffu, queue_me);
      if (d.prof && qm && lane == 0) { atomicAdd(&d.prof[19], (unsigned long long)(pc3 - pc1)); atomicAdd(&d.prof[20], (unsigned long long)(clock64() - pc3)); }
      if (qm) {
        if (pend_n + __popc(qm) > kPendCap) flush_pending_jobs(d, rr_par, s_pend, pend_n, lane);
        if (queue_me) s_pend[pend_n + __popc(qm & ((1u << lane) - 1u))] = queue_job;
        pend_n += __popc(qm);
      }
    }
    tile = next;
    next = dyn ? tile0 + (int)__shfl_sync(0xffffffffu, ticket, 0) : next + total;
  }
  if (RR && use_spec) flush_pending_jobs(d, rr_par, s_pend, pend_n, lane);
  if (lane == 0) bulk_wait_read0();
  unsigned long long prof_t0 = 0;
  if (RR && d.prof && lane == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_t0));
    atomicMax(&d.prof[1], prof_t0);            // end of the last warp's tiles
    atomicMin(&d.prof[0], prof_t0);            // end of the first warp's tiles
    const unsigned long long dur = prof_t0 - prof_start;
    const int cls = prof_commits ? 13 : 10;
    atomicAdd(&d.prof[cls], dur); atomicAdd(&d.prof[cls + 1], 1ull); atomicMax(&d.prof[cls + 2], dur);
  }
  if (RR && use_spec) {
    // last warp out: the list drained at the start of this launch becomes the next launch's append list
    const int q = rr_par ^ 1;
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(&d.sched[4], 1u) == (uint32_t)total - 1u) {  // last warp out: the drained list becomes the next append list
        d.sched[q] = 0; d.sched[2 + q] = 0; d.sched[4] = 0; d.sched[6] = 0;
        d.sched[5] = (uint32_t)q;
      }
    }
  }
}

// T transitions in one launch from a recorded action stream u8 [T][N]; state stays on chip and the observation
// tile is double-buffered so step t's bulk store overlaps step t+1's compute.
template <bool SEE, bool RR, int TILE, typename EXT>
__global__ void __launch_bounds__(TILE) k_rollout(Dev d, const uint8_t *actions, int T, StepArgs A0) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  float *s_obs0 = reinterpret_cast<float *>(smem);
  float *s_obs1 = reinterpret_cast<float *>(smem + (size_t)TILE * kObsFloats * 4);
  uint32_t *s_rows = reinterpret_cast<uint32_t *>(smem + 2 * (size_t)TILE * kObsFloats * 4);
  const int tid = threadIdx.x, base = blockIdx.x * TILE, e = base + tid, W = d.c.W, N = d.N;
  const int n_tile = min(TILE, N - base);
  const bool valid = tid < n_tile;
  stage_rows_begin<TILE>(d, s_rows, &bar, base);
  // the all-zero state plane of both observation buffers is written once (emit_packed_f32 leaves it alone)
#pragma unroll
  for (int i = 0; i < kV * kV; i++) { s_obs0[tid * kObsFloats + 2 * kV * kV + i] = 0.0f; s_obs1[tid * kObsFloats + 2 * kV * kV + i] = 0.0f; }
  Env s{};
  if (valid) s = unpack(d.hot[e]);
  mbar_wait(&bar, 0);
  uint32_t *my_rows = s_rows + (tid >> 5) * W * kWarpTile + (tid & 31);  // [sub][W][32]
  uint32_t *s_rng = s_rows + (TILE / kWarpTile) * W * kWarpTile;           // [32][TILE]
  bool dirty = false;
  for (int t = 0; t < T; t++) {
    StepArgs A = A0;
    const size_t off = (size_t)t * N;
    if (A.o.image) A.o.image += off * kObsFloats;
    if (A.o.direction) A.o.direction += off;
    if (A.o.reward) A.o.reward += off;
    if (A.o.flags) A.o.flags += off;
    if (A.o.ep_return) A.o.ep_return += off;
    if (A.o.ep_length) A.o.ep_length += off;
    if (A.o.trunc_image) A.o.trunc_image += off * kObsFloats;
    if (A.o.trunc_direction) A.o.trunc_direction += off;
    if (A.o.trunc_full_obs) A.o.trunc_full_obs += off * 3 * W * W;
    if (A.o.masks) A.o.masks += off;
    if (A.o.bad_masks) A.o.bad_masks += off;
    if (A.o.cliffhanger_masks) A.o.cliffhanger_masks += off;
    if (A.o.image_u8) A.o.image_u8 += off * kObsFloats;
    A.last_step = (t == T - 1) ? A0.last_step : 0;
    float *s_obs = (t & 1) ? s_obs1 : s_obs0;
    // the store that last used this buffer (step t-2) must have finished reading it: at most one group pending
    if (tid == 0 && t >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    if (valid) {
      float rew;
      const uint32_t flags = step_one<SEE, RR, EXT>(d, my_rows, kWarpTile, s, e, (int)actions[off + e], A, s_obs + tid * kObsFloats,
                                                    rew, dirty, s_rng + tid, TILE);
      write_step_scalars(A, e, s, flags, rew);
    }
    if (A.o.image) store_obs_tile(A.o.image + (size_t)base * kObsFloats, s_obs, n_tile, false);
  }
  if (tid == 0) bulk_wait_read0();
  if (valid) {
    d.hot[e] = pack(s);
    if (RR && dirty) {
      const Rows G = env_rows(d, e);
      for (int r = 0; r < W; r++) G.set(r, my_rows[r * kWarpTile]);
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
static int grid_for(int n, int block) { return (n + block - 1) / block; }

static int check_cfg(const mgplr_env_config *c) {
  if (!c) return fail(MGPLR_E_BADARG, "config is NULL");
  if (c->width < 5 || c->width > 32) return fail(MGPLR_E_UNSUPPORTED, "width must be in [5, 32]");
  if (c->agent_view_size != 5) return fail(MGPLR_E_UNSUPPORTED, "agent_view_size must be 5");
  if (c->max_steps < 1 || c->max_steps > 65535) return fail(MGPLR_E_UNSUPPORTED, "max_steps must be in [1, 65535]");
  if (c->max_episode_steps < 1 || c->max_episode_steps > 32767)
    return fail(MGPLR_E_UNSUPPORTED, "max_episode_steps must be in [1, 32767]");
  if (c->n_clutter < 0 || c->n_clutter > 4000) return fail(MGPLR_E_UNSUPPORTED, "n_clutter must be in [0, 4000]");
  if (c->n_editor_actions < 2 || c->n_editor_actions > 4) return fail(MGPLR_E_BADARG, "n_editor_actions must be 2, 3 or 4");
  return 0;
}

template <typename T>
static cudaError_t dalloc(T **p, size_t count, int64_t &total) {
  total += (int64_t)(count * sizeof(T));
  return cudaMalloc((void **)p, count * sizeof(T));
}

extern "C" int mgplr_venv_create(const mgplr_env_config *cfg, int32_t num_envs, int32_t device, mgplr_venv **out) {
  if (!out) return fail(MGPLR_E_BADARG, "out is NULL");
  *out = nullptr;
  if (int rc = check_cfg(cfg)) return rc;
  if (num_envs < 1) return fail(MGPLR_E_BADARG, "num_envs must be >= 1");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(ce != cudaSuccess ? (int)ce : (int)cudaErrorNoDevice,
                "mgplr: no CUDA device -- this library has no CPU fallback");
  CK(cudaSetDevice(device));
  mgplr_venv *v = new (std::nothrow) mgplr_venv();
  if (!v) return fail(MGPLR_E_BADARG, "out of host memory");
  memset(v, 0, sizeof(*v));
  v->device = device;
  Dev &d = v->d;
  d.N = num_envs;
  // bit 0: evict_last on level rows + hot records, evict_first on the observation bulk stores; bit 2: streaming (.cs) stores for
  // the per-env scalars and streaming action loads; A/B only: bit 3 = no evict_first on the observations, bit 4 = evict_first
  // on rows + hot records, bit 5 = no evict_last on them, bit 6 = evict_first on the hot records; bit 7: hot records evict_normal
  // (default 133 = bits 0, 2, 7)
  d.l2_hints = getenv("MGPLR_L2_HINTS") ? atoi(getenv("MGPLR_L2_HINTS")) : 133;
  if (getenv("MGPLR_L2_PERSIST_MB")) {
    // A/B knob: an L2 set-aside for persisting (evict_last) lines, cudaLimitPersistingL2CacheSize -- device-wide
    int max_persist = 0;
    CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device));
    size_t want = (size_t)atoi(getenv("MGPLR_L2_PERSIST_MB")) << 20;
    if (want > (size_t)max_persist) want = (size_t)max_persist;
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    if (getenv("MGPLR_VERBOSE")) fprintf(stderr, "mgplr: persisting L2 set-aside %zu MB (device max %d MB)\n", want >> 20, max_persist >> 20);
  }
  d.use_tma = getenv("MGPLR_TMA") ? atoi(getenv("MGPLR_TMA")) : 1;
  v->pdl = getenv("MGPLR_PDL") ? atoi(getenv("MGPLR_PDL")) : 0;  // measured: no gain at 131 072 envs, slower at 4 096
  d.c = Cfg{cfg->width, cfg->max_steps, cfg->max_episode_steps, cfg->see_through_walls != 0, cfg->n_clutter,
            cfg->resample_n_clutter != 0, cfg->choose_goal_last != 0, cfg->fixed_environment != 0, cfg->n_editor_actions};
  const size_t N = (size_t)num_envs;
  int64_t total = 0;
  const size_t wall_words = (size_t)cfg->width * 32 * ((N + 31) / 32);  // padded to whole 32-env tiles
  CK(dalloc(&d.wall, wall_words, total));
  CK(cudaMemset(d.wall, 0, wall_words * sizeof(uint32_t)));
  CK(dalloc(&d.hot, N, total));
  CK(dalloc(&d.adv, N, total));
  CK(dalloc(&d.metrics, N, total));
  const size_t mt_words = 624 * 32 * ((N + 31) / 32);  // padded to whole 32-env tiles
  CK(dalloc(&d.mt, mt_words, total));
  CK(dalloc(&d.mti, N, total));
  CK(dalloc(&d.limbs, 3 * N, total));
  CK(dalloc(&d.words, N, total));
  CK(dalloc(&d.err, N, total));
  CK(dalloc(&d.sched, 8, total));
  CK(cudaMemset(d.sched, 0, 8 * sizeof(uint32_t)));
  CK(dalloc(&d.spec, N, total));
  CK(cudaMemset(d.spec, 0, N * sizeof(uint32_t)));
  d.cand = nullptr; d.rr_list = nullptr;  // DR speculation buffers: allocated by the first DR step launch (rr_spec_alloc)
  v->rr_spec = getenv("MGPLR_RR_SPEC") ? atoi(getenv("MGPLR_RR_SPEC")) : 1;  // DR speculation; 0 = in-kernel rebuild only
  v->host_dma = getenv("MGPLR_HOST_DMA") ? atoi(getenv("MGPLR_HOST_DMA")) : 0;
  CK(cudaStreamCreateWithFlags(&v->copy_stream, cudaStreamNonBlocking));
  for (int c = 0; c < 8; c++) CK(cudaEventCreateWithFlags(&v->copy_done[c], cudaEventDisableTiming));
  d.prof = nullptr;
  if (getenv("MGPLR_RR_PROF")) {  // debug: phase timestamps / job cycles of the DR step kernel (mgplr_debug_prof)
    CK(cudaMalloc((void **)&d.prof, 24 * sizeof(unsigned long long)));
    CK(cudaMemset(d.prof, 0, 24 * sizeof(unsigned long long)));
  }
  CK(dalloc(&v->seed_scratch, 4 * N, total));
  CK(dalloc(&v->act_dev, N, total));
  CK(dalloc(&v->cnt_dev, 2, total));
  CK(cudaMemset(v->cnt_dev, 0, 2 * sizeof(uint32_t)));
  CK(cudaHostAlloc((void **)&v->res_pin, 16 * N + N, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void **)&v->res_pin_dev, v->res_pin, 0));
  memset(v->res_pin, 0xff, 16 * N);  // env = -1 in every record slot
  v->bytes = total;
  CK(cudaDeviceGetAttribute(&v->sm_count, cudaDevAttrMultiProcessorCount, device));
  {
    const int W = cfg->width;
#define SET_STEP(SEE, RR, EXT)                                                                          \
  CK(cudaFuncSetAttribute(k_step_env<SEE, RR, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                          (int)(4 * warp_smem_bytes(W, RR))));
#define SET_ROLL(SEE, RR, TL, EXT) \
  CK(cudaFuncSetAttribute(k_rollout<SEE, RR, TL, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bytes(W, TL, 2)));
#define SET_ALL(EXT)                                                                                  \
  SET_STEP(true, false, EXT) SET_STEP(true, true, EXT) SET_STEP(false, false, EXT) SET_STEP(false, true, EXT) \
  SET_ROLL(true, false, 64, EXT) SET_ROLL(true, true, 64, EXT) SET_ROLL(false, false, 64, EXT) SET_ROLL(false, true, 64, EXT)
    if (W <= 24) { SET_ALL(uint32_t) } else { SET_ALL(uint64_t) }
#undef SET_ALL
#undef SET_ROLL
#undef SET_STEP
  }
  {  // kernels whose dynamic shared memory can exceed the 48 KB default at the widest grids
    const int adv_smem = (int)((size_t)kAdvGroup * 3 * cfg->width * cfg->width * sizeof(float));
    CK(cudaFuncSetAttribute(k_adv_image<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, adv_smem));
    CK(cudaFuncSetAttribute(k_adv_image<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, adv_smem));
    CK(cudaFuncSetAttribute(k_reset_random, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg->width * 128 * 4));
    CK(cudaFuncSetAttribute(k_reset_agent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * 32 * kObsFloats * sizeof(float))));
  }
  CK(cudaMemset(d.mt, 0, mt_words * sizeof(uint32_t)));
  k_init<<<grid_for(num_envs, 256), 256>>>(d);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  *out = v;
  return 0;
}

extern "C" void mgplr_venv_destroy(mgplr_venv *v) {
  if (!v) return;
  cudaSetDevice(v->device);
  Dev &d = v->d;
  if (v->copy_stream) cudaStreamDestroy(v->copy_stream);
  for (int c = 0; c < 8; c++) if (v->copy_done[c]) cudaEventDestroy(v->copy_done[c]);
  cudaFree(d.wall); cudaFree(d.hot); cudaFree(d.adv); cudaFree(d.metrics); cudaFree(d.mt); cudaFree(d.mti);
  cudaFree(d.limbs); cudaFree(d.words); cudaFree(d.err); cudaFree(d.sched); cudaFree(d.spec); cudaFree(d.cand); cudaFree(d.rr_list); cudaFree(v->seed_scratch); cudaFree(v->act_dev); cudaFree(v->cnt_dev); cudaFreeHost(v->res_pin);
  delete v;
}

// debug (MGPLR_RR_PROF=1): out[0..5] = first / last warp past its tiles, end of regeneration (globaltimer ns), job cycles sum,
// jobs, max job cycles -- of the launches since the last call; resets the counters
extern "C" int mgplr_debug_prof(mgplr_venv *v, unsigned long long *out) {
  if (!v || !v->d.prof) return fail(MGPLR_E_BADARG, "profiling is off (MGPLR_RR_PROF)");
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, v->d.prof, 24 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  unsigned long long init[24] = {~0ull};
  CK(cudaMemcpy(v->d.prof, init, sizeof(init), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int32_t mgplr_venv_num_envs(const mgplr_venv *v) { return v ? v->d.N : 0; }
extern "C" int64_t mgplr_venv_state_bytes(const mgplr_venv *v) { return v ? v->bytes : 0; }

// make the handle's device current only when it is not already (cudaGetDevice is a thread-local read)
static inline cudaError_t use_device(int device) {
  int cur = -1;
  cudaError_t e = cudaGetDevice(&cur);
  if (e != cudaSuccess) return e;
  return cur == device ? cudaSuccess : cudaSetDevice(device);
}
#define NEED(v)                                                 \
  if (!(v)) return fail(MGPLR_E_BADARG, "venv handle is NULL"); \
  CK(use_device((v)->device));                                  \
  cudaStream_t st = (cudaStream_t)stream

extern "C" int mgplr_seed(mgplr_venv *v, const uint32_t *limbs_host, const int32_t *n_limbs_host, const int32_t *index_host,
                          int32_t n, void *stream) {
  NEED(v);
  if (!limbs_host || !n_limbs_host || n < 1 || n > v->d.N) return fail(MGPLR_E_BADARG, "mgplr_seed: bad arguments");
  const size_t N = (size_t)v->d.N;
  uint32_t *tmp = (uint32_t *)malloc(4 * (size_t)n * sizeof(uint32_t));
  if (!tmp) return fail(MGPLR_E_BADARG, "out of host memory");
  for (int k = 0; k < n; k++) {
    tmp[k] = limbs_host[2 * k]; tmp[(size_t)n + k] = limbs_host[2 * k + 1];
    tmp[2 * (size_t)n + k] = (uint32_t)n_limbs_host[k]; tmp[3 * (size_t)n + k] = index_host ? (uint32_t)index_host[k] : (uint32_t)k;
    if (n_limbs_host[k] < 1 || n_limbs_host[k] > 2) { free(tmp); return fail(MGPLR_E_BADARG, "seed limb count must be 1 or 2"); }
  }
  (void)N;
  cudaError_t e = cudaMemcpyAsync(v->seed_scratch, tmp, 4 * (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  free(tmp);
  if (e != cudaSuccess) return cuda_fail(e, "mgplr_seed copy");
  k_seed<<<grid_for(n, 128), 128, 0, st>>>(v->d, v->seed_scratch, n, n, index_host != nullptr);
  CK(cudaGetLastError());
  return 0;
}

static int launch_adv_image(mgplr_venv *v, float *adv_image, float *time_step, int mode, const int64_t *loc, uint8_t *done,
                            cudaStream_t st, int raw = 0) {
  const size_t smem = (size_t)kAdvGroup * 3 * v->d.c.W * v->d.c.W * sizeof(float);
  const int grid = (v->d.N + kAdvGroup - 1) / kAdvGroup;
  if (mode) k_adv_image<true><<<grid, 128, smem, st>>>(v->d, adv_image, time_step, mode, loc, done);
  else k_adv_image<false><<<grid, 128, smem, st>>>(v->d, adv_image, time_step, 0, nullptr, nullptr, raw);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_reset(mgplr_venv *v, float *adv_image, float *time_step, void *stream) {
  NEED(v);
  // small batches are launch-bound: one fused launch; large batches are bandwidth-bound: a dense state-update kernel
  // (one thread per env) followed by the image kernel beats 4 busy threads per CTA
  if (adv_image && v->d.N <= kFuseAdvMaxEnvs) return launch_adv_image(v, adv_image, time_step, 1, nullptr, nullptr, st);
  k_reset<<<grid_for(v->d.N, 128), 128, 0, st>>>(v->d);
  CK(cudaGetLastError());
  if (adv_image) return launch_adv_image(v, adv_image, time_step, 0, nullptr, nullptr, st);
  return 0;
}

extern "C" int mgplr_step_adversary(mgplr_venv *v, const int64_t *loc, float *adv_image, float *time_step, uint8_t *done,
                                    void *stream) {
  NEED(v);
  if (!loc) return fail(MGPLR_E_BADARG, "loc is NULL");
  if (adv_image && v->d.N <= kFuseAdvMaxEnvs) return launch_adv_image(v, adv_image, time_step, 2, loc, done, st);
  k_step_adversary<<<grid_for(v->d.N, 128), 128, 0, st>>>(v->d, loc, done);
  CK(cudaGetLastError());
  if (adv_image) return launch_adv_image(v, adv_image, time_step, 0, nullptr, nullptr, st);
  return 0;
}

static OutPtrs outptrs(const mgplr_step_out *o) {
  OutPtrs p{nullptr, nullptr, nullptr};
  if (o) { p.image = o->image; p.direction = o->direction; p.image_u8 = o->image_u8; }
  return p;
}

extern "C" int mgplr_reset_agent(mgplr_venv *v, const mgplr_step_out *out, void *stream) {
  NEED(v);
  v->steps_since_sweep = 0;  // k_reset_agent replays every deferred respawn
  k_reset_agent<<<grid_for(v->d.N, 128), 128, 4 * 32 * kObsFloats * sizeof(float), st>>>(v->d, outptrs(out));
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_reset_random(mgplr_venv *v, const int32_t *n_walls, const mgplr_step_out *out, void *stream) {
  NEED(v);
  k_reset_random<<<grid_for(v->d.N, 128), 128, (size_t)v->d.c.W * 128 * 4, st>>>(v->d, n_walls, outptrs(out));
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_reset_to_encoding(mgplr_venv *v, const uint8_t *enc, const int32_t *index, int32_t n,
                                       const mgplr_step_out *out, void *stream) {
  NEED(v);
  if (!enc || n < 1 || n > v->d.N) return fail(MGPLR_E_BADARG, "mgplr_reset_to_encoding: bad arguments");
  k_reset_to_encoding<<<grid_for(n, 128), 128, 0, st>>>(v->d, enc, index, n, outptrs(out));
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_load_levels(mgplr_venv *v, const uint8_t *enc, int32_t n_levels, const int32_t *level_index,
                                 int32_t start_dir, const mgplr_step_out *out, void *stream) {
  NEED(v);
  if (!enc || n_levels < 1) return fail(MGPLR_E_BADARG, "mgplr_load_levels: bad arguments");
  k_flush<<<grid_for(v->d.N, 128), 128, 0, st>>>(v->d, 1);  // the RNG stream keeps its reference position
  k_load_levels<<<grid_for(v->d.N, 128), 128, 0, st>>>(v->d, enc, n_levels, level_index, start_dir, outptrs(out));
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_load_levels_at(mgplr_venv *v, const uint8_t *enc, const int32_t *env_index, int32_t n, int32_t start_dir,
                                    const mgplr_step_out *out, void *stream) {
  NEED(v);
  if (!enc || !env_index || n < 1 || n > v->d.N) return fail(MGPLR_E_BADARG, "mgplr_load_levels_at: bad arguments");
  k_load_levels_at<<<grid_for(n, 128), 128, 0, st>>>(v->d, enc, env_index, n, start_dir, outptrs(out));
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_reset_to_actions(mgplr_venv *v, const int32_t *locs, int32_t len, const int32_t *index, int32_t n,
                                      const mgplr_step_out *out, void *stream) {
  NEED(v);
  if (!locs || len < 0 || len > 4095 || n < 1 || n > v->d.N) return fail(MGPLR_E_BADARG, "mgplr_reset_to_actions: bad arguments");
  k_reset_to_actions<<<grid_for(n, 128), 128, 0, st>>>(v->d, locs, len, index, n, outptrs(out));
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_mutate_edits(mgplr_venv *v, const int32_t *locs, const int32_t *ops, const int32_t *n_edits,
                                  int32_t max_edits, uint8_t *need, int32_t *n_free, void *stream) {
  NEED(v);
  if (!locs || !ops || !n_edits || max_edits < 0) return fail(MGPLR_E_BADARG, "mgplr_mutate_edits: bad arguments");
  k_mutate_edits<<<grid_for(v->d.N, 128), 128, 0, st>>>(v->d, locs, ops, n_edits, max_edits, need, n_free);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_mutate_finalize(mgplr_venv *v, const int32_t *choice, const mgplr_step_out *out, void *stream) {
  NEED(v);
  k_mutate_finalize<<<grid_for(v->d.N, 128), 128, 0, st>>>(v->d, choice, outptrs(out));
  CK(cudaGetLastError());
  return 0;
}

static int launch_step_args(mgplr_venv *v, StepArgs &A, int32_t reset_random, cudaStream_t st, int tile0 = 0, int tile1 = -1);
static int launch_step(mgplr_venv *v, const int64_t *action, int32_t reset_random, const int32_t *n_walls, int32_t last_step,
                       const mgplr_step_out *out, cudaStream_t st) {
  StepArgs A;
  memset(&A, 0, sizeof(A));
  A.stream_scalars = (v->d.l2_hints & 4) != 0;
  A.action = action; A.n_walls = n_walls; A.last_step = last_step;
  if (out) A.o = *out;
  return launch_step_args(v, A, reset_random, st);
}
// candidate records + job lists of the DR speculation (3.5 KB per env at 15x15): only DR users pay for them
static int rr_spec_alloc(mgplr_venv *v) {
  if (v->d.cand) return 0;
  const size_t N = (size_t)v->d.N;
  cudaError_t e = dalloc(&v->d.cand, N * 4 * (size_t)cand_words(v->d.c.W), v->bytes);
  if (e == cudaSuccess) e = dalloc(&v->d.rr_list, 4 * N, v->bytes);  // uint2 entries
  if (e != cudaSuccess) {
    cudaGetLastError();
    v->d.cand = nullptr;
    return fail((int)e, "allocating the DR speculation buffers failed (the first reset_random step must run outside CUDA graph capture)");
  }
  return 0;
}

// Deferred goal respawns are counted in 8 bits per env (Env::pending); a level whose goal sits next to the start can finish an
// episode every step or two, and a saturated counter is flushed INSIDE the step kernel, one serial MT draw at a time (~0.5 ms
// for one warp).  Every kSweepEvery-th reset_agent-mode step launch is therefore preceded by a sweep that replays the draws of
// the (rare) envs that have accumulated kSweepMin or more -- at full speed, with the batched generator -- so that no env
// comes near 255 however long the caller goes without a reset_agent().  8 MB of hot records per sweep at 524 288 envs.
constexpr int kSweepEvery = 64, kSweepMin = 96;

static int launch_step_args(mgplr_venv *v, StepArgs &A, int32_t reset_random, cudaStream_t st, int tile0, int tile1) {
  if (!reset_random && tile0 == 0 && ++v->steps_since_sweep >= kSweepEvery) {
    v->steps_since_sweep = 0;
    k_flush<<<grid_for(v->d.N, 128), 128, 0, st>>>(v->d, kSweepMin);
    CK(cudaGetLastError());
  }
  // persistent grid: as many 4-warp CTAs as fit on the chip (shared-memory bound), capped by the tile count
  const int W = v->d.c.W, wpc = 4;
  const size_t smem = wpc * warp_smem_bytes(W, reset_random != 0);
  const int all_tiles = (v->d.N + kWarpTile - 1) / kWarpTile;
  const int n_tiles = (tile1 < 0 || tile1 > all_tiles) ? all_tiles : tile1;  // exclusive end of this launch's tile range
  const int per_sm = (int)((227 * 1024) / (smem + 1024)) > 0 ? (int)((227 * 1024) / (smem + 1024)) : 1;
  int grid = v->sm_count * per_sm;
  const bool see = v->d.c.see_through, rr = reset_random != 0, narrow = W <= 24;
  // small batches finish less than one env per launch: the in-kernel rebuild is cheaper than a standing job phase
  // ... and a level with many walls does not fit the 224-word look-ahead window (2-3 words per try): no point in queueing jobs
  A.spec = rr && v->rr_spec && !v->d.c.fixed_env && !v->d.c.resample && (v->d.N >= 16384 || v->rr_spec > 1) &&
           v->d.c.n_clutter / 2 <= 56;
  if (A.spec) { if (int rc = rr_spec_alloc(v)) return rc; }
  const int need = (n_tiles - tile0 + wpc - 1) / wpc;
  if (grid > need) grid = need;
  if (grid < 1) return 0;
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3(grid); lc.blockDim = dim3(wpc * 32); lc.dynamicSmemBytes = smem; lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr; lc.numAttrs = (v->pdl && !A.spec) ? 1 : 0;  // (the DR speculation reads a per-launch parity word)
#define LAUNCH(SEE, RR, EXT) CK(cudaLaunchKernelEx(&lc, k_step_env<SEE, RR, EXT>, v->d, A, tile0, n_tiles))
#define BY_MODE(EXT)                                  \
  do {                                                \
    if (see && rr) LAUNCH(true, true, EXT);           \
    else if (see) LAUNCH(true, false, EXT);           \
    else if (rr) LAUNCH(false, true, EXT);            \
    else LAUNCH(false, false, EXT);                   \
  } while (0)
  if (narrow) BY_MODE(uint32_t); else BY_MODE(uint64_t);
#undef BY_MODE
#undef LAUNCH
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_step_env(mgplr_venv *v, const int64_t *action, int32_t reset_random, const int32_t *n_walls,
                              int32_t last_step, const mgplr_step_out *out, void *stream) {
  NEED(v);
  if (!action) return fail(MGPLR_E_BADARG, "action is NULL");
  return launch_step(v, action, reset_random, n_walls, last_step, out, st);
}

extern "C" int mgplr_step_env_u8(mgplr_venv *v, const uint8_t *action, int32_t reset_random, const int32_t *n_walls,
                                 int32_t last_step, const mgplr_step_out *out, void *stream) {
  NEED(v);
  if (!action) return fail(MGPLR_E_BADARG, "action is NULL");
  StepArgs A;
  memset(&A, 0, sizeof(A));
  A.stream_scalars = (v->d.l2_hints & 4) != 0;
  A.action_u8 = action; A.n_walls = n_walls; A.last_step = last_step;
  if (out) A.o = *out;
  return launch_step_args(v, A, reset_random, st);
}

constexpr int kHostChunks = 8;
// Device view of a host pointer when it is pinned (cudaHostAlloc / cudaHostRegister) memory, else NULL.
static void *mapped_view(const void *host) {
  if (!host) return nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// Host-driven transition.  With pinned buffers the whole call is ONE kernel launch and one stream synchronisation:
// the kernel reads the actions straight from the caller's pinned memory (zero-copy over PCIe, prefetched one tile ahead),
// writes the flags straight into the caller's pinned flags buffer and appends the done records to a device-mapped pinned
// list (dense prefix of an env = -1 sentinel-filled array, so no count has to come back).  Pageable buffers take staged
// copies around the same launch.
static int step_env_host_impl(mgplr_venv *v, const void *action_host_any, size_t act_size, int32_t reset_random, int32_t last_step,
                              const mgplr_step_out *out_dev, uint8_t *flags_host, mgplr_done_record *done_host,
                              int32_t done_capacity, int32_t *n_done_host, void *stream) {
  NEED(v);
  if (!action_host_any) return fail(MGPLR_E_BADARG, "action_host is NULL");
  const size_t N = (size_t)v->d.N;
  mgplr_done_record *list_h = (mgplr_done_record *)v->res_pin;
  uint8_t *flags_pin = v->res_pin + 16 * N;
  const bool narrow = act_size == 1;
  const int64_t *action_host = (const int64_t *)action_host_any;  // (the chunked-DMA A/B path below is int64-only)
  const int64_t *act = (const int64_t *)mapped_view(action_host_any);
  // A/B knob (MGPLR_HOST_DMA=1, off): stage the pinned actions with the copy engine in chunks on a side stream and run
  // the step as one sub-launch per chunk behind its copy.  Measured slower than the zero-copy read on both sizes
  // (524 288 envs: 3.2 vs 4.4 G env-steps/s; 131 072: 2.2 vs 2.8): the event hand-offs cost more than the PCIe gain.
  int chunks = 1;
  if (act && v->host_dma && N >= 65536 && !narrow) {
    chunks = (int)(N / 65536);
    if (chunks > kHostChunks) chunks = kHostChunks;
    act = v->act_dev;
  }
  if (!act) {
    CK(cudaMemcpyAsync(v->act_dev, action_host_any, N * act_size, cudaMemcpyHostToDevice, st));
    act = v->act_dev;
  }
  uint8_t *flags_map = flags_host ? (uint8_t *)mapped_view(flags_host) : nullptr;
  const bool stage_flags = flags_host && !flags_map;
  StepArgs A;
  memset(&A, 0, sizeof(A));
  A.stream_scalars = (v->d.l2_hints & 4) != 0;
  if (out_dev) A.o = *out_dev;
  if (narrow) A.action_u8 = (const uint8_t *)act; else A.action = act;
  A.last_step = last_step;
  A.done_count = v->cnt_dev + (v->host_steps & 1u);
  A.done_count_next = v->cnt_dev + ((v->host_steps + 1u) & 1u);
  A.done_list = (mgplr_done_record *)v->res_pin_dev;
  A.flags_host = stage_flags ? v->res_pin_dev + 16 * N : flags_map;
  v->host_steps++;
  if (chunks > 1) {
    const int all_tiles = (int)((N + kWarpTile - 1) / kWarpTile);
    const int per = ((all_tiles + chunks - 1) / chunks + 3) & ~3;  // tiles per chunk, a multiple of the CTA's 4 warps
    for (int c = 0; c < chunks; c++) {
      const int t0 = c * per, t1 = (t0 + per < all_tiles) ? t0 + per : all_tiles;
      if (t0 >= t1) break;
      const size_t e0 = (size_t)t0 * kWarpTile, e1 = ((size_t)t1 * kWarpTile < N) ? (size_t)t1 * kWarpTile : N;
      CK(cudaMemcpyAsync(v->act_dev + e0, action_host + e0, (e1 - e0) * sizeof(int64_t), cudaMemcpyHostToDevice, v->copy_stream));
      CK(cudaEventRecord(v->copy_done[c], v->copy_stream));
      CK(cudaStreamWaitEvent(st, v->copy_done[c], 0));
      if (int rc = launch_step_args(v, A, reset_random, st, t0, t1)) return rc;
    }
  } else if (int rc = launch_step_args(v, A, reset_random, st)) return rc;
  CK(cudaStreamSynchronize(st));
  if (stage_flags) memcpy(flags_host, flags_pin, N);
  // the records are a dense prefix of the sentinel-filled list
  size_t n_done = 0;
  while (n_done < N && list_h[n_done].env >= 0) n_done++;
  if (n_done_host) *n_done_host = (int32_t)n_done;
  if (done_host && done_capacity > 0) {
    const size_t want = n_done < (size_t)done_capacity ? n_done : (size_t)done_capacity;
    memcpy(done_host, list_h, want * sizeof(mgplr_done_record));
  }
  for (size_t k = 0; k < n_done; k++) list_h[k].env = -1;
  return 0;
}

extern "C" int mgplr_step_env_host(mgplr_venv *v, const int64_t *action_host, int32_t reset_random, int32_t last_step,
                                   const mgplr_step_out *out_dev, uint8_t *flags_host, mgplr_done_record *done_host,
                                   int32_t done_capacity, int32_t *n_done_host, void *stream) {
  return step_env_host_impl(v, action_host, sizeof(int64_t), reset_random, last_step, out_dev, flags_host, done_host, done_capacity,
                            n_done_host, stream);
}
extern "C" int mgplr_step_env_host_u8(mgplr_venv *v, const uint8_t *action_host, int32_t reset_random, int32_t last_step,
                                      const mgplr_step_out *out_dev, uint8_t *flags_host, mgplr_done_record *done_host,
                                      int32_t done_capacity, int32_t *n_done_host, void *stream) {
  return step_env_host_impl(v, action_host, 1, reset_random, last_step, out_dev, flags_host, done_host, done_capacity, n_done_host,
                            stream);
}

extern "C" int mgplr_rollout_ex(mgplr_venv *v, const uint8_t *actions, int32_t T, int32_t reset_random, int32_t last_step,
                                const mgplr_step_out *out_t0, void *stream) {
  NEED(v);
  if (!actions || T < 1) return fail(MGPLR_E_BADARG, "mgplr_rollout: bad arguments");
  StepArgs A;
  memset(&A, 0, sizeof(A));
  A.stream_scalars = (v->d.l2_hints & 4) != 0;
  if (out_t0) A.o = *out_t0;
  A.last_step = last_step;  // applied to step T-1 only (k_rollout)
  const int tile = 64;
  const int grid = grid_for(v->d.N, tile);
  const size_t smem = step_smem_bytes(v->d.c.W, tile, 2, reset_random != 0);
  const bool see = v->d.c.see_through, rr = reset_random != 0, narrow = v->d.c.W <= 24;
#define LAUNCH(SEE, RR, EXT) k_rollout<SEE, RR, 64, EXT><<<grid, 64, smem, st>>>(v->d, actions, T, A)
#define BY_MODE(EXT)                                  \
  do {                                                \
    if (see && rr) LAUNCH(true, true, EXT);           \
    else if (see) LAUNCH(true, false, EXT);           \
    else if (rr) LAUNCH(false, true, EXT);            \
    else LAUNCH(false, false, EXT);                   \
  } while (0)
  if (narrow) BY_MODE(uint32_t); else BY_MODE(uint64_t);
#undef BY_MODE
#undef LAUNCH
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_rollout(mgplr_venv *v, const uint8_t *actions, int32_t T, int32_t reset_random,
                             const mgplr_step_out *out_t0, void *stream) {
  return mgplr_rollout_ex(v, actions, T, reset_random, 0, out_t0, stream);
}

extern "C" int mgplr_full_obs(mgplr_venv *v, float *full_obs, void *stream) {
  NEED(v);
  if (!full_obs) return fail(MGPLR_E_BADARG, "full_obs is NULL");
  return launch_adv_image(v, full_obs, nullptr, 0, nullptr, nullptr, st, 1);
}

extern "C" int mgplr_render_images(mgplr_venv *v, const uint8_t *tiles, const int32_t *index, int32_t n, uint8_t *images,
                                   void *stream) {
  NEED(v);
  if (!tiles || !images || n < 1 || (!index && n > v->d.N)) return fail(MGPLR_E_BADARG, "mgplr_render_images: bad arguments");
  if ((((uintptr_t)tiles) & 3u) || (((uintptr_t)images) & 3u)) return fail(MGPLR_E_BADARG, "mgplr_render_images: pointers must be 4-byte aligned");
  k_render_images<<<dim3(v->d.c.W, n), 256, 0, st>>>(v->d, reinterpret_cast<const uint32_t *>(tiles), index, n, images);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mgplr_get_encodings(mgplr_venv *v, uint8_t *enc, void *stream) {
  NEED(v);
  if (!enc) return fail(MGPLR_E_BADARG, "enc is NULL");
  const size_t total = (size_t)v->d.N * v->d.c.W * v->d.c.W;
  k_encode<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(v->d, enc);
  CK(cudaGetLastError());
  return 0;
}
extern "C" int mgplr_get_metrics(mgplr_venv *v, int32_t *metrics, void *stream) {
  NEED(v);
  if (!metrics) return fail(MGPLR_E_BADARG, "metrics is NULL");
  k_get_metrics<<<grid_for(v->d.N, 256), 256, 0, st>>>(v->d, metrics);
  CK(cudaGetLastError());
  return 0;
}
extern "C" int mgplr_get_agent_state(mgplr_venv *v, int32_t *state, void *stream) {
  NEED(v);
  if (!state) return fail(MGPLR_E_BADARG, "state is NULL");
  k_flush<<<grid_for(v->d.N, 128), 128, 0, st>>>(v->d, 1);
  k_get_agent_state<<<grid_for(v->d.N, 256), 256, 0, st>>>(v->d, state);
  CK(cudaGetLastError());
  return 0;
}
extern "C" int mgplr_get_errors(mgplr_venv *v, uint32_t *errors, int32_t clear, void *stream) {
  NEED(v);
  if (!errors) return fail(MGPLR_E_BADARG, "errors is NULL");
  k_get_errors<<<grid_for(v->d.N, 256), 256, 0, st>>>(v->d, errors, clear);
  CK(cudaGetLastError());
  return 0;
}

// host-side replay of the incremental generator on a copy of one env's state (no device mutation)
extern "C" int mgplr_peek_rng(mgplr_venv *v, int32_t index, uint32_t *words_host, int32_t count) {
  if (!v) return fail(MGPLR_E_BADARG, "venv handle is NULL");
  CK(cudaSetDevice(v->device));
  if (index < 0 || index >= v->d.N || !words_host || count < 0) return fail(MGPLR_E_BADARG, "mgplr_peek_rng: bad arguments");
  k_flush<<<grid_for(v->d.N, 128), 128>>>(v->d, 1);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  uint32_t mt[624], idx = 0;
  CK(cudaMemcpy2D(mt, kMtChunk * sizeof(uint32_t), v->d.mt + mt_at(index, 0), 32 * kMtChunk * sizeof(uint32_t),
                  kMtChunk * sizeof(uint32_t), kMtChunks, cudaMemcpyDeviceToHost));  // 39 chunks of 16 words
  CK(cudaMemcpy(&idx, v->d.mti + index, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  for (int k = 0; k < count; k++) {
    const uint32_t i = idx, i1 = (i + 1 == 624) ? 0 : i + 1, im = (i + 397 >= 624) ? i + 397 - 624 : i + 397;
    uint32_t y = (mt[i] & 0x80000000u) | (mt[i1] & 0x7fffffffu);
    y = mt[im] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    mt[i] = y; idx = i1;
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    words_host[k] = y;
  }
  return 0;
}
