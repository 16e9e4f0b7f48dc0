// mgplr_plr.cu -- rollout math of the PLR path on sm_100a: GAE (algos/storage.py:233-256), per-episode
// score reduction (level_replay/level_sampler.py:486-549,307-349), rank/staleness sample weights
// (level_sampler.py:726-785) and sequential replay sampling (level_sampler.py:664-680,601-604).
//
// All rollout tensors are [T or T+1][N] with the actor index fastest (RolloutStorage's [T,N,1]), so one
// thread per actor walking the time axis reads 128 contiguous bytes per warp per step.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mgplr.h"

extern "C" int mgplr_set_error_(int code, const char *msg);  // mgplr_venv.cu (shared last-error slot)
static int pfail(int code, const char *msg) { return mgplr_set_error_(code, msg); }
#define PCK(call)                                                        \
  do {                                                                   \
    cudaError_t _e = (call);                                             \
    if (_e != cudaSuccess) return pfail((int)_e, cudaGetErrorString(_e)); \
  } while (0)

// ------------------------------------------------------------------------------------------ GAE
// delta = r[t] + gamma*v[t+1]*m[t+1] - v[t];  gae = delta + (gamma*lambda)*m[t+1]*gae;  ret[t] = gae + v[t]
// evaluated in float32 with one rounding per operation, in the reference's operand order (no FMA
// contraction), so the result is bit-identical to the torch CPU ops of algos/storage.py:251-256.
__global__ void k_gae(const float *__restrict__ rewards, const float *__restrict__ values, const float *__restrict__ masks,
                      float *__restrict__ returns, int T, int N, float gamma, float gl) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  float gae = 0.0f;
  float v_next = values[(size_t)T * N + e];
  constexpr int kB = 8;  // steps loaded together (3 kB independent loads in flight), consumed in reverse order
  for (int t1 = T; t1 > 0; t1 -= kB) {
    float rb[kB], mb[kB], vb[kB];
#pragma unroll
    for (int u = 0; u < kB; u++) {
      const int t = t1 - 1 - u;
      const bool in = t >= 0;
      rb[u] = in ? rewards[(size_t)t * N + e] : 0.f;
      mb[u] = in ? masks[(size_t)(t + 1) * N + e] : 0.f;
      vb[u] = in ? values[(size_t)t * N + e] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kB; u++) {
      const int t = t1 - 1 - u;
      if (t < 0) break;
      const float r = rb[u], m = mb[u], v = vb[u];
      const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, v_next), m)), v);
      gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, m), gae));
      returns[(size_t)t * N + e] = __fadd_rn(gae, v);
      v_next = v;
    }
  }
}

// gamma / gae_lambda are the Python floats (doubles): `gamma*value_preds` casts gamma to float32, while
// `gamma * gae_lambda * masks` multiplies the two in double FIRST and casts the product (storage.py:252,255).
extern "C" int mgplr_gae(const float *rewards, const float *value_preds, const float *masks, float *returns, int32_t T,
                         int32_t N, double gamma, double gae_lambda, void *stream) {
  if (!rewards || !value_preds || !masks || !returns || T < 1 || N < 1) return pfail(MGPLR_E_BADARG, "mgplr_gae: bad arguments");
  k_gae<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, value_preds, masks, returns, T, N, (float)gamma,
                                                          (float)(gamma * gae_lambda));
  PCK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ discounted returns
// returns[t] = (returns[t+1] * gamma) * masks[t+1] + rewards[t]  (algos/storage.py:276-279), one thread per actor
__global__ void k_discounted_returns(const float *__restrict__ rewards, const float *__restrict__ masks, float *__restrict__ returns,
                                     int T, int N, float gamma) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  float ret = returns[(size_t)T * N + e];
  for (int t = T - 1; t >= 0; t--) {
    ret = __fadd_rn(__fmul_rn(__fmul_rn(ret, gamma), masks[(size_t)(t + 1) * N + e]), rewards[(size_t)t * N + e]);
    returns[(size_t)t * N + e] = ret;
  }
}

extern "C" int mgplr_discounted_returns(const float *rewards, const float *masks, float *returns, int32_t T, int32_t N, double gamma,
                                        void *stream) {
  if (!rewards || !masks || !returns || T < 1 || N < 1) return pfail(MGPLR_E_BADARG, "mgplr_discounted_returns: bad arguments");
  k_discounted_returns<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, masks, returns, T, N, (float)gamma);
  PCK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ batched value loss
// per-actor mean over the rollout of |ret - v| / (ret - v) / max(ret - v, 0), optionally ^power, optionally clamped
// (algos/storage.py:290-327).  The per-step term is float32 like the reference's; the T-term sum is accumulated in
// double and rounded once (torch's float32 column sum differs from it by rounding only: tests use 1e-5 relative).
__global__ void k_batched_value_loss(const float *__restrict__ returns, const float *__restrict__ values, int T, int N, int mode,
                                     int power, int clipped, float *__restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  double acc = 0.0;
  for (int t = 0; t < T; t++) {
    float td = __fsub_rn(returns[(size_t)t * N + e], values[(size_t)t * N + e]);
    if (mode == 0) td = fabsf(td);
    else if (mode == 2) td = fmaxf(td, 0.0f);
    float p = td;
    for (int k = 1; k < power; k++) p = __fmul_rn(p, td);
    acc += (double)p;
  }
  float m = (float)(acc / (double)T);
  if (clipped) m = fminf(fmaxf(m, -1.0f), 1.0f);
  out[e] = m;
}

extern "C" int mgplr_batched_value_loss(const float *returns, const float *value_preds, int32_t T, int32_t N, int32_t mode,
                                        int32_t power, int32_t clipped, float *out, void *stream) {
  if (!returns || !value_preds || !out || T < 1 || N < 1 || mode < 0 || mode > 2 || power < 1 || power > 16)
    return pfail(MGPLR_E_BADARG, "mgplr_batched_value_loss: bad arguments");
  k_batched_value_loss<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(returns, value_preds, T, N, mode, power, clipped, out);
  PCK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ episode scores
// done[t] = !(masks[t] > 0) for t in 0..T.  An episode is [start_t, t) for every done step t >= 1
// (t == 0 is skipped WITHOUT moving start_t, level_sampler.py:504-505).
__global__ void k_count_episodes(const float *__restrict__ masks, int T, int N, int32_t *__restrict__ counts) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  int c = 0;
#pragma unroll 8
  for (int t = 1; t <= T; t++) c += !(masks[(size_t)t * N + e] > 0.f);
  c += (masks[(size_t)T * N + e] > 0.f);  // trailing partial episode (level_sampler.py:551-578)
  counts[e] = c;
}

// Exclusive scan of the per-actor episode counts (canonical actor-major record order), two levels:
// k_scan_blocks scans 1024 counts per CTA and emits the CTA total, k_scan_tops scans the CTA totals in one CTA;
// consumers add block_off[actor / 1024].
__device__ __forceinline__ int32_t block_exclusive_scan_1024(int32_t v, int32_t *warp_sums, int32_t &total) {
  int32_t x = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    int32_t w = warp_sums[threadIdx.x];
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (threadIdx.x >= o) w += y;
    }
    warp_sums[threadIdx.x] = w;
  }
  __syncthreads();
  const int32_t warp_prefix = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
  total = warp_sums[31];
  return warp_prefix + x - v;
}

__global__ void __launch_bounds__(1024) k_scan_blocks(const int32_t *__restrict__ counts, int N, int32_t *__restrict__ offsets,
                                                      int32_t *__restrict__ block_sums) {
  __shared__ int32_t warp_sums[32];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  int32_t total;
  const int32_t ex = block_exclusive_scan_1024((i < N) ? counts[i] : 0, warp_sums, total);
  if (i < N) offsets[i] = ex;
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_tops(int32_t *__restrict__ block_sums, int n_blocks, int32_t *__restrict__ total_out) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    int32_t total;
    const int32_t ex = block_exclusive_scan_1024((i < n_blocks) ? block_sums[i] : 0, warp_sums, total);
    const int32_t c = carry;
    __syncthreads();
    if (i < n_blocks) block_sums[i] = c + ex;
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// per-step score of the policy-logit strategies from one step's logits x[0..A): with lse = log sum exp(x),
// least confidence = 1 - exp(max x - lse); margin = exp(x1 - lse) - exp(x2 - lse) for the two largest logits
__device__ __forceinline__ float logit_score(const float *__restrict__ x, int A, int strategy) {
  float m1 = -INFINITY, m2 = -INFINITY;
  for (int j = 0; j < A; j++) {
    const float v = x[j];
    if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v;
  }
  float sum = 0.f;
  for (int j = 0; j < A; j++) sum += expf(x[j] - m1);
  const float lse = m1 + logf(sum);
  if (strategy == MGPLR_SCORE_LEAST_CONFIDENCE) return 1.0f - expf(m1 - lse);
  return expf(m1 - lse) - expf(m2 - lse);
}

template <int strategy>  // compile-time: the common value-loss instances stay free of the logit / TD code
__global__ void k_episode_scores(const float *__restrict__ masks, const float *__restrict__ cliff, const float *__restrict__ returns,
                                 const float *__restrict__ values, const float *__restrict__ rewards,
                                 const int32_t *__restrict__ seeds, const float *__restrict__ logits, int A, float gamma,
                                 int T, int N,
                                 const int32_t *__restrict__ offsets, const int32_t *__restrict__ block_off,
                                 mgplr_episode *__restrict__ out, int max_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  int k = offsets[e] + block_off[e >> 10];
  int start = 0;
  double sum = 0.0, vsum = 0.0;
  float mx = -INFINITY, rsum = 0.f, vmin = INFINITY;
  float r_prev = 0.f, v_prev = 0.f;  // ONE_STEP_TD: the term of step t-1 needs v[t]
  // The time loop is latency-bound if it issues one step's loads at a time (4 x 128 B in flight per warp): steps are
  // loaded kB at a time (4 kB independent loads in flight) and then consumed in order.
  constexpr int kB = 8;  // (16 was measured slower: 403 vs 235 us at 131 072 actors -- registers, occupancy)
  // (loading batch k+1 while batch k is consumed was measured slower too: 297 us, 96 registers)
  for (int t0 = 0; t0 < T; t0 += kB) {
    float retb[kB], vb[kB], rb[kB], mb[kB];
#pragma unroll
    for (int u = 0; u < kB; u++) {
      const int t = t0 + u;
      const bool in = t < T;
      retb[u] = (in && returns) ? returns[(size_t)t * N + e] : 0.f;
      vb[u] = in ? values[(size_t)t * N + e] : 0.f;
      rb[u] = in ? rewards[(size_t)t * N + e] : 0.f;
      mb[u] = in ? masks[(size_t)(t + 1) * N + e] : 1.f;
    }
#pragma unroll
    for (int u = 0; u < kB; u++) {
      const int t = t0 + u;
      if (t >= T) break;
      // accumulate step t into the running episode [start, ...)
      const float v = vb[u], r = rb[u];
      float a = retb[u] - v;
      if (strategy == MGPLR_SCORE_POSITIVE_VALUE_LOSS) a = fmaxf(a, 0.f);
      else if (strategy == MGPLR_SCORE_VALUE_L1) a = fabsf(a);
      else if (strategy == MGPLR_SCORE_LEAST_CONFIDENCE) a = logit_score(logits + ((size_t)t * N + e) * A, A, strategy);
      else if (strategy == MGPLR_SCORE_MIN_MARGIN) a = -logit_score(logits + ((size_t)t * N + e) * A, A, strategy);  // max of -s = -min s
      if (strategy == MGPLR_SCORE_ONE_STEP_TD) {
        if (t > start) {  // term of step t-1: |r[t-1] + gamma*v[t] - v[t-1]|
          a = fabsf(__fsub_rn(__fadd_rn(r_prev, __fmul_rn(gamma, v)), v_prev));
          sum += (double)a; mx = fmaxf(mx, a);
        }
        r_prev = r; v_prev = v;
      } else {
        sum += (double)a; mx = fmaxf(mx, a);
      }
      rsum += r;  // torch sums the f32 rewards of the slice (level_sampler.py:534)
      vsum += (double)v; vmin = fminf(vmin, v);
      // done at t+1 closes the episode [start, t+1)
      if (!(mb[u] > 0.f)) {
        const int t_end = t + 1;
        if (k < max_out) {
          mgplr_episode ep;
          ep.actor = e; ep.t_start = start; ep.t_end = t_end; ep.seed = seeds ? seeds[(size_t)start * N + e] : -1;
          const int n = t_end - start;
          ep.mean_score = (float)(sum / (double)n); ep.max_score = mx; ep.reward_sum = rsum;
          if (strategy == MGPLR_SCORE_MIN_MARGIN) { ep.mean_score = (float)(1.0 + sum / (double)n); ep.max_score = 1.0f + mx; }
          if (strategy == MGPLR_SCORE_ONE_STEP_TD) {  // n-1 terms; a one-step episode scores r[0] - v[0] (level_sampler.py:431-434)
            if (n > 1) ep.mean_score = (float)(sum / (double)(n - 1));
            else { ep.mean_score = __fsub_rn(r_prev, v_prev); ep.max_score = ep.mean_score; }
          }
          ep.value_sum = (float)vsum; ep.value_min = vmin;
          ep.cliffhanger = cliff ? !(cliff[(size_t)t_end * N + e] > 0.f) : 0;
          out[k] = ep;
        }
        k++;
        start = t_end; sum = 0.0; vsum = 0.0; mx = -INFINITY; rsum = 0.f; vmin = INFINITY;
      }
    }
  }
  if (start < T && k < max_out) {  // not-done tail: a partial record (cliffhanger field = 2)
    mgplr_episode ep;
    ep.actor = e; ep.t_start = start; ep.t_end = T; ep.seed = seeds ? seeds[(size_t)start * N + e] : -1;
    const int n = T - start;
    ep.mean_score = (float)(sum / (double)n); ep.max_score = mx; ep.reward_sum = rsum;
    if (strategy == MGPLR_SCORE_MIN_MARGIN) { ep.mean_score = (float)(1.0 + sum / (double)n); ep.max_score = 1.0f + mx; }
    if (strategy == MGPLR_SCORE_ONE_STEP_TD) {
      if (n > 1) ep.mean_score = (float)(sum / (double)(n - 1));
      else { ep.mean_score = __fsub_rn(r_prev, v_prev); ep.max_score = ep.mean_score; }
    }
    ep.value_sum = (float)vsum; ep.value_min = vmin; ep.cliffhanger = 2;
    out[k] = ep;
  }
}

// ---- time-split variant: kSeg threads per actor.  One thread per actor leaves a B200 under-filled below ~300 000 actors (131 072
// actors = 43 % of the thread slots) and the loop is latency-bound, so the rollout is cut into kSeg time segments: CTA = 32 actors x
// kSeg segments (warp = segment, lane = actor: loads stay coalesced), every thread scores the episodes that END inside its segment.
// The first of them may have begun in an earlier segment: its accumulators are completed after a CTA barrier from the "tail" parts
// (steps after a segment's last done) the earlier segments left in shared memory.  Record order is unchanged because the episode
// counts are taken per (actor, segment) and scanned in that order.
constexpr int kSeg = 4;

struct ScorePart {  // accumulators of a run of consecutive steps of one actor
  double sum, vsum;
  float mx, rsum, vmin;
  int start;        // first step of the run
};

__device__ __forceinline__ void part_reset(ScorePart &p, int start) {
  p.sum = 0.0; p.vsum = 0.0; p.mx = -INFINITY; p.rsum = 0.f; p.vmin = INFINITY; p.start = start;
}

// `a` precedes `b` in time
__device__ __forceinline__ void part_prepend(ScorePart &b, const ScorePart &a) {
  b.sum = a.sum + b.sum; b.vsum = a.vsum + b.vsum; b.mx = fmaxf(a.mx, b.mx); b.rsum = a.rsum + b.rsum; b.vmin = fminf(a.vmin, b.vmin);
  b.start = a.start;
}

template <int strategy>
__device__ __forceinline__ void part_emit(const ScorePart &p, int e, int t_end, float r_last, float v_last, int cliffhanger,
                                          const int32_t *__restrict__ seeds, int N, mgplr_episode *__restrict__ out) {
  mgplr_episode ep;
  ep.actor = e; ep.t_start = p.start; ep.t_end = t_end; ep.seed = seeds ? seeds[(size_t)p.start * N + e] : -1;
  const int n = t_end - p.start;
  ep.mean_score = (float)(p.sum / (double)n); ep.max_score = p.mx; ep.reward_sum = p.rsum;
  if (strategy == MGPLR_SCORE_MIN_MARGIN) { ep.mean_score = (float)(1.0 + p.sum / (double)n); ep.max_score = 1.0f + p.mx; }
  if (strategy == MGPLR_SCORE_ONE_STEP_TD) {
    if (n > 1) ep.mean_score = (float)(p.sum / (double)(n - 1));
    else { ep.mean_score = __fsub_rn(r_last, v_last); ep.max_score = ep.mean_score; }
  }
  ep.value_sum = (float)p.vsum; ep.value_min = p.vmin; ep.cliffhanger = cliffhanger;
  *out = ep;
}

// counts[e * kSeg + s] = episodes of actor e that end inside segment s (+ the not-done tail in the last segment)
__global__ void __launch_bounds__(32 * kSeg) k_count_episodes_split(const float *__restrict__ masks, int T, int N, int L,
                                                                    int32_t *__restrict__ counts) {
  const int e = blockIdx.x * 32 + (threadIdx.x & 31), s = threadIdx.x >> 5;
  if (e >= N) return;
  const int t0 = s * L, t1 = min(T, t0 + L);
  int c = 0;
#pragma unroll 8
  for (int t = t0; t < t1; t++) c += !(masks[(size_t)(t + 1) * N + e] > 0.f);
  if (s == kSeg - 1) c += (masks[(size_t)T * N + e] > 0.f);
  counts[(size_t)e * kSeg + s] = c;
}

template <int strategy>
__global__ void __launch_bounds__(32 * kSeg, 6) k_episode_scores_split(
    const float *__restrict__ masks, const float *__restrict__ cliff, const float *__restrict__ returns,
    const float *__restrict__ values, const float *__restrict__ rewards, const int32_t *__restrict__ seeds,
    const float *__restrict__ logits, int A, float gamma, int T, int N, int L, const int32_t *__restrict__ offsets,
    const int32_t *__restrict__ block_off, mgplr_episode *__restrict__ out, int max_out) {
  __shared__ ScorePart s_tail[kSeg][32], s_head[kSeg][32];  // head: the run up to the segment's first done (kept out of registers)
  __shared__ int s_head_end[kSeg][32];
  __shared__ float s_head_r[kSeg][32], s_head_v[kSeg][32];
  __shared__ uint8_t s_closed[kSeg][32];  // the segment saw a done (its tail does not reach further back)
  const int lane = threadIdx.x & 31, s = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  const bool valid = e < N;
  const int t0 = s * L, t1 = min(T, t0 + L);
  int k = 0;
  if (valid) {
    const size_t ci = (size_t)e * kSeg + s;
    k = offsets[ci] + block_off[ci >> 10];
  }
  ScorePart acc;
  part_reset(acc, t0);
  bool has_head = false;
  float r_prev = 0.f, v_prev = 0.f;
  bool have_prev = false;  // ONE_STEP_TD: step t-1 belongs to the same episode
  if (strategy == MGPLR_SCORE_ONE_STEP_TD && valid && s > 0 && t0 < T && masks[(size_t)t0 * N + e] > 0.f) {
    have_prev = true; r_prev = rewards[(size_t)(t0 - 1) * N + e]; v_prev = values[(size_t)(t0 - 1) * N + e];
  }
  constexpr int kB = 8;
  if (valid)
    for (int tb = t0; tb < t1; tb += kB) {
      float retb[kB], vb[kB], rb[kB], mb[kB];
#pragma unroll
      for (int u = 0; u < kB; u++) {
        const int t = tb + u;
        const bool in = t < t1;
        retb[u] = (in && returns) ? returns[(size_t)t * N + e] : 0.f;
        vb[u] = in ? values[(size_t)t * N + e] : 0.f;
        rb[u] = in ? rewards[(size_t)t * N + e] : 0.f;
        mb[u] = in ? masks[(size_t)(t + 1) * N + e] : 1.f;
      }
#pragma unroll
      for (int u = 0; u < kB; u++) {
        const int t = tb + u;
        if (t >= t1) break;
        const float v = vb[u], r = rb[u];
        float a = retb[u] - v;
        if (strategy == MGPLR_SCORE_POSITIVE_VALUE_LOSS) a = fmaxf(a, 0.f);
        else if (strategy == MGPLR_SCORE_VALUE_L1) a = fabsf(a);
        else if (strategy == MGPLR_SCORE_LEAST_CONFIDENCE) a = logit_score(logits + ((size_t)t * N + e) * A, A, strategy);
        else if (strategy == MGPLR_SCORE_MIN_MARGIN) a = -logit_score(logits + ((size_t)t * N + e) * A, A, strategy);
        if (strategy == MGPLR_SCORE_ONE_STEP_TD) {
          if (have_prev) {
            a = fabsf(__fsub_rn(__fadd_rn(r_prev, __fmul_rn(gamma, v)), v_prev));
            acc.sum += (double)a; acc.mx = fmaxf(acc.mx, a);
          }
          r_prev = r; v_prev = v; have_prev = true;
        } else {
          acc.sum += (double)a; acc.mx = fmaxf(acc.mx, a);
        }
        acc.rsum += r;
        acc.vsum += (double)v; acc.vmin = fminf(acc.vmin, v);
        if (!(mb[u] > 0.f)) {  // done at t+1 closes the running episode
          const int t_end = t + 1;
          if (!has_head) {     // the first one may have begun before this segment: finished after the barrier
            has_head = true;
            s_head[s][lane] = acc; s_head_end[s][lane] = t_end; s_head_r[s][lane] = r_prev; s_head_v[s][lane] = v_prev;
          } else if (k < max_out) {
            part_emit<strategy>(acc, e, t_end, r_prev, v_prev, cliff ? !(cliff[(size_t)t_end * N + e] > 0.f) : 0, seeds, N, out + k);
          }
          k++;
          part_reset(acc, t_end);
          have_prev = false;
        }
      }
    }
  s_tail[s][lane] = acc;
  s_closed[s][lane] = has_head;
  __syncthreads();
  if (!valid) return;
  if (has_head) {
    ScorePart head = s_head[s][lane];
    const int head_end = s_head_end[s][lane];
    const size_t ci = (size_t)e * kSeg + s;
    const int head_k = offsets[ci] + block_off[ci >> 10];  // the segment's first record
    for (int j = s - 1; j >= 0; j--) {
      part_prepend(head, s_tail[j][lane]);
      if (s_closed[j][lane]) break;
    }
    if (head_k < max_out)
      part_emit<strategy>(head, e, head_end, s_head_r[s][lane], s_head_v[s][lane], cliff ? !(cliff[(size_t)head_end * N + e] > 0.f) : 0, seeds, N,
                          out + head_k);
  }
  if (s == kSeg - 1) {  // not-done tail: a partial record (cliffhanger field = 2)
    if (!has_head)
      for (int j = s - 1; j >= 0; j--) {
        part_prepend(acc, s_tail[j][lane]);
        if (s_closed[j][lane]) break;
      }
    if (acc.start < T && k < max_out) {
      // r_last / v_last of a one-step TD tail: the last step of the rollout
      const float r_l = rewards[(size_t)(T - 1) * N + e], v_l = values[(size_t)(T - 1) * N + e];
      part_emit<strategy>(acc, e, T, r_l, v_l, 2, seeds, N, out + k);
    }
  }
}

// scan scratch of the episode-score launcher: one grow-only buffer PER DEVICE (a process may drive several GPUs); growing
// synchronises that device first so that no in-flight launch still uses the old buffer
constexpr int kMaxDevices = 64;
static int32_t *g_scratch[kMaxDevices] = {nullptr};
static size_t g_scratch_n[kMaxDevices] = {0};

extern "C" int mgplr_plr_episode_scores(const float *masks, const float *cliffhanger_masks, const float *returns,
                                        const float *value_preds, const float *rewards, const int32_t *level_seeds, int32_t T,
                                        int32_t N, int32_t strategy, mgplr_episode *episodes, int32_t max_episodes,
                                        int32_t *n_episodes, void *stream) {
  return mgplr_plr_episode_scores_ex(masks, cliffhanger_masks, returns, value_preds, rewards, level_seeds, nullptr, 0, 0.0, T, N,
                                     strategy, episodes, max_episodes, n_episodes, stream);
}

extern "C" int mgplr_plr_episode_scores_ex(const float *masks, const float *cliffhanger_masks, const float *returns,
                                           const float *value_preds, const float *rewards, const int32_t *level_seeds,
                                           const float *action_log_dist, int32_t num_actions, double gamma, int32_t T, int32_t N,
                                           int32_t strategy, mgplr_episode *episodes, int32_t max_episodes,
                                           int32_t *n_episodes, void *stream) {
  if (!masks || !value_preds || !rewards || !episodes || !n_episodes || T < 1 || N < 1 || strategy < 0 || strategy > MGPLR_SCORE_ONE_STEP_TD)
    return pfail(MGPLR_E_BADARG, "mgplr_plr_episode_scores: bad arguments");
  const bool logit = strategy == MGPLR_SCORE_LEAST_CONFIDENCE || strategy == MGPLR_SCORE_MIN_MARGIN;
  if (logit && (!action_log_dist || num_actions < 2)) return pfail(MGPLR_E_BADARG, "action_log_dist required for this strategy");
  if (strategy <= MGPLR_SCORE_VALUE_L1 && !returns) return pfail(MGPLR_E_BADARG, "returns required for this strategy");
  int dev = 0;
  PCK(cudaGetDevice(&dev));
  // one thread per actor fills the machine only for large batches: below that, kSeg time segments per actor (knob MGPLR_SCORE_SPLIT:
  // unset = by size, 0 = never, 1 = always).  Measured (T = 256, whole pass, us): 32 actors 87 -> 39, 4 096 100 -> 44,
  // 32 768 174 -> 98, 131 072 235 -> 236, 262 144 418 -> 439, 524 288 791 -> 824 (80 registers against 64).
  static const int split_knob = [] { const char *s = getenv("MGPLR_SCORE_SPLIT"); return s ? atoi(s) : -1; }();
  const bool split = T >= 4 * kSeg && (split_knob > 0 || (split_knob < 0 && N <= 65536));
  const size_t n_counts = split ? (size_t)N * kSeg : (size_t)N;
  const int n_blocks = (int)((n_counts + 1023) / 1024);
  const size_t need = 2 * n_counts + (size_t)n_blocks;
  if (dev < 0 || dev >= kMaxDevices) return pfail(MGPLR_E_UNSUPPORTED, "device index out of range");
  if (g_scratch_n[dev] < need) {
    if (g_scratch[dev]) { PCK(cudaDeviceSynchronize()); cudaFree(g_scratch[dev]); }
    g_scratch[dev] = nullptr; g_scratch_n[dev] = 0;
    PCK(cudaMalloc((void **)&g_scratch[dev], need * sizeof(int32_t)));
    g_scratch_n[dev] = need;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int32_t *counts = g_scratch[dev], *offsets = counts + n_counts, *block_off = counts + 2 * n_counts;
  const int L = (T + kSeg - 1) / kSeg;
  if (split) k_count_episodes_split<<<(N + 31) / 32, 32 * kSeg, 0, st>>>(masks, T, N, L, counts);
  else k_count_episodes<<<(N + 127) / 128, 128, 0, st>>>(masks, T, N, counts);
  k_scan_blocks<<<n_blocks, 1024, 0, st>>>(counts, (int)n_counts, offsets, block_off);
  k_scan_tops<<<1, 1024, 0, st>>>(block_off, n_blocks, n_episodes);
#define SCORES(S)                                                                                                          \
  if (split)                                                                                                               \
    k_episode_scores_split<S><<<(N + 31) / 32, 32 * kSeg, 0, st>>>(                                                        \
        masks, cliffhanger_masks, (S <= MGPLR_SCORE_VALUE_L1) ? returns : nullptr, value_preds, rewards, level_seeds,      \
        action_log_dist, num_actions, (float)gamma, T, N, L, offsets, block_off, episodes, max_episodes);                  \
  else                                                                                                                     \
    k_episode_scores<S><<<(N + 127) / 128, 128, 0, st>>>(masks, cliffhanger_masks, (S <= MGPLR_SCORE_VALUE_L1) ? returns : nullptr, \
                                                         value_preds, rewards, level_seeds, action_log_dist, num_actions, (float)gamma, \
                                                         T, N, offsets, block_off, episodes, max_episodes)
  switch (strategy) {
    case MGPLR_SCORE_POSITIVE_VALUE_LOSS: SCORES(MGPLR_SCORE_POSITIVE_VALUE_LOSS); break;
    case MGPLR_SCORE_SIGNED_VALUE_LOSS: SCORES(MGPLR_SCORE_SIGNED_VALUE_LOSS); break;
    case MGPLR_SCORE_VALUE_L1: SCORES(MGPLR_SCORE_VALUE_L1); break;
    case MGPLR_SCORE_MAX_MC: SCORES(MGPLR_SCORE_MAX_MC); break;
    case MGPLR_SCORE_LEAST_CONFIDENCE: SCORES(MGPLR_SCORE_LEAST_CONFIDENCE); break;
    case MGPLR_SCORE_MIN_MARGIN: SCORES(MGPLR_SCORE_MIN_MARGIN); break;
    default: SCORES(MGPLR_SCORE_ONE_STEP_TD); break;
  }
#undef SCORES
  PCK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ sample weights
// One CTA.  rank transform: rank 1 = highest score; ties broken by index (higher index first), i.e.
// np.flip(np.argsort(scores, kind='stable')).  weights = 1/rank^(1/T) masked by seen, normalised; staleness:
// clip(s,0)^(1/Ts) masked, normalised; mix (1-c) w + c s.
constexpr int kMaxBuf = 8192;

struct Key {
  double s;
  int32_t i;
};
// "a comes before b" in DESCENDING order with the tie rule above
__device__ __forceinline__ bool before(const Key &a, const Key &b) { return a.s > b.s || (a.s == b.s && a.i > b.i); }

__device__ void block_sort_desc(Key *keys, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const Key a = keys[i], b = keys[ixj];
          if (up ? before(b, a) : before(a, b)) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
}

__device__ double block_sum(double v, double *red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double w = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
    if (threadIdx.x == 0) red[32] = w;
  }
  __syncthreads();
  const double r = red[32];
  __syncthreads();
  return r;
}

// two independent block sums in one pass (each with exactly block_sum's reduction tree, so the bits are block_sum's)
__device__ double block_sum2(double v, double w2, double *red, double &out2) {
  __shared__ double red2[33];
  for (int o = 16; o > 0; o >>= 1) { v += __shfl_down_sync(0xffffffffu, v, o); w2 += __shfl_down_sync(0xffffffffu, w2, o); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = v; red2[threadIdx.x >> 5] = w2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    double a = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0, b = (threadIdx.x < (blockDim.x >> 5)) ? red2[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    if (threadIdx.x == 0) { red[32] = a; red2[32] = b; }
  }
  __syncthreads();
  const double r = red[32];
  out2 = red2[32];
  __syncthreads();
  return r;
}

// _score_transform (level_sampler.py:752-785):
//   1 rank:       w = 1 / rank^(1/T), rank 1 = highest value, ties by index (see before())
//   2 power:      w = (clip(v, 0) + eps)^(1/T)
//   0 constant:   w = 1
//   3 softmax:    w = exp(v / T)
//   4 match:      w = ((1 - v) v)^(1/T)
//   5 match_rank: rank transform of (1 - v) v
//   6 eps_greedy: w = eps / n everywhere, + (1 - eps) at argmax(v) (first index of the maximum: numpy argmax); `eps` is the
//                 sampler's eps here, not the power transform's offset
// ('max' draws its argmax tie-break from np.random inside the transform: not built.)
// out[i] receives the raw transform of vals[i]; `keys` is dynamic shared memory for the sort.
__device__ void transform_vals(int transform, const double *vals, int n, double temperature, double eps, double *out, Key *keys) {
  const double inv_t = 1.0 / temperature;
  if (transform == 1 || transform == 5) {
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      Key k;
      const double v = (i < n) ? vals[i] : 0.0;
      k.s = (i < n) ? (transform == 5 ? __dmul_rn(__dsub_rn(1.0, v), v) : v) : -INFINITY; k.i = (i < n) ? i : -1 - i;
      keys[i] = k;
    }
    __syncthreads();
    block_sort_desc(keys, n2);
    for (int r = threadIdx.x; r < n; r += blockDim.x) out[keys[r].i] = 1.0 / pow((double)(r + 1), inv_t);
  } else if (transform == 2) {
    // pow(x, 1.0) == x exactly; the double-precision pow routine is ~300 instructions per element and was what a replay draw /
    // an admission mostly consisted of with the default staleness temperature 1 (10.7 us per draw at 4 000 slots)
    if (inv_t == 1.0) { for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = fmax(vals[i], 0.0) + eps; }
    else for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = pow(fmax(vals[i], 0.0) + eps, inv_t);
  } else if (transform == 3) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = exp(vals[i] / temperature);
  } else if (transform == 4) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double m = __dmul_rn(__dsub_rn(1.0, vals[i]), vals[i]);
      out[i] = inv_t == 1.0 ? m : pow(m, inv_t);
    }
  } else if (transform == 6) {
    // first index of the maximum: pack (value, -index) comparisons through one shared slot per warp, then thread 0
    __shared__ double s_best_v[32];
    __shared__ int s_best_i[32];
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (vals[i] > bv || (vals[i] == bv && i < bi)) { bv = vals[i]; bi = i; }
    for (int o = 16; o; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, bv, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_best_v[threadIdx.x >> 5] = bv; s_best_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)((blockDim.x + 31) >> 5); w++)
        if (s_best_v[w] > bv || (s_best_v[w] == bv && s_best_i[w] < bi)) { bv = s_best_v[w]; bi = s_best_i[w]; }
      s_best_i[0] = bi;
    }
    __syncthreads();
    const int am = s_best_i[0];
    const double base = eps / (double)n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = (i == am) ? __dadd_rn(__dsub_rn(1.0, eps), base) : base;
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = 1.0;
  }
  __syncthreads();
}

// mask by seen, normalise (uniform over seen when the mass is zero): level_sampler.py:728-736 / 741-746
__device__ void mask_normalise(double *w, const double *unseen, int n, bool uniform_fallback_normalised, double *red) {
  double part = 0.0, seen_part = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = w[i] * (1.0 - unseen[i]);
    w[i] = v; part += v; seen_part += (1.0 - unseen[i]);
  }
  double nseen;
  const double z = block_sum2(part, seen_part, red, nseen);   // (one pass for both sums: same reduction trees, half the barriers)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (z > 0) w[i] = w[i] / z;
    else {
      const double u = (1.0 / (double)n) * (1.0 - unseen[i]);
      w[i] = uniform_fallback_normalised ? u / (nseen / (double)n) : u;
    }
  }
  __syncthreads();
}

struct WeightArgs {
  int score_transform, stale_transform;
  double temperature, eps, coef, stale_temperature;
};

__device__ void score_weights(const double *scores, const double *unseen, int n, const WeightArgs &a, double *w_score, Key *keys,
                              double *red) {
  transform_vals(a.score_transform, scores, n, a.temperature, a.eps, w_score, keys);
  mask_normalise(w_score, unseen, n, true, red);
}

__device__ void mix_staleness(const double *w_score, const double *staleness, const double *unseen, int n, const WeightArgs &a,
                              double *weights, Key *keys, double *red) {
  if (!(a.coef > 0)) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) weights[i] = w_score[i];
    __syncthreads();
    return;
  }
  transform_vals(a.stale_transform, staleness, n, a.stale_temperature, 0.0, weights, keys);
  mask_normalise(weights, unseen, n, false, red);
  for (int i = threadIdx.x; i < n; i += blockDim.x) weights[i] = (1.0 - a.coef) * w_score[i] + a.coef * weights[i];
  __syncthreads();
}

// `cached` != NULL: the normalised score weights were computed before (mgplr_plr_score_weights) and are reused --
// scores do not change between the draws of a rollout, only staleness does, so the sort is paid once per update.
__global__ void __launch_bounds__(1024) k_sample_weights(const double *scores, const double *staleness, const double *unseen, int n,
                                                         WeightArgs a, double *weights, double *w_score, const double *cached) {
  extern __shared__ __align__(16) uint8_t sm[];
  Key *keys = reinterpret_cast<Key *>(sm);
  __shared__ double red[33];
  if (cached) w_score = const_cast<double *>(cached);
  else score_weights(scores, unseen, n, a, w_score, keys, red);
  mix_staleness(w_score, staleness, unseen, n, a, weights, keys, red);
}

__global__ void __launch_bounds__(1024) k_score_weights(const double *scores, const double *unseen, int n, WeightArgs a,
                                                        double *w_score) {
  extern __shared__ __align__(16) uint8_t sm[];
  Key *keys = reinterpret_cast<Key *>(sm);
  __shared__ double red[33];
  score_weights(scores, unseen, n, a, w_score, keys, red);
}

// rank-weight / weight work arrays of the weight and draw kernels: one fixed-size buffer PER DEVICE, allocated on first use
// (calls on one device are expected on one stream at a time, like every other use of a LevelSampler)
static double *g_dscratch_dev[kMaxDevices] = {nullptr};
static double *g_dscratch = nullptr;   // the current device's buffer, set by ensure_dscratch
static int ensure_dscratch(size_t n) {
  int dev = 0;
  PCK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices || n > 2 * (size_t)kMaxBuf) return pfail(MGPLR_E_UNSUPPORTED, "device index out of range");
  if (!g_dscratch_dev[dev]) PCK(cudaMalloc((void **)&g_dscratch_dev[dev], 2 * (size_t)kMaxBuf * sizeof(double)));
  g_dscratch = g_dscratch_dev[dev];
  return 0;
}
static size_t sort_smem(int n) {
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  return (size_t)n2 * sizeof(Key);
}

static int check_transform(int t);
static size_t sort_smem(int n);
static int ensure_dscratch(size_t n);

extern "C" int mgplr_plr_score_weights(const double *scores, const double *unseen, int32_t n, int32_t score_transform,
                                       double temperature, double eps, double *score_weights, void *stream) {
  if (!scores || !unseen || !score_weights || n < 1 || n > kMaxBuf)
    return pfail(MGPLR_E_BADARG, "mgplr_plr_score_weights: bad arguments (n must be in [1, 8192])");
  if (int rc = check_transform(score_transform)) return rc;
  const size_t smem = sort_smem(n);
  PCK(cudaFuncSetAttribute(k_score_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const WeightArgs a{score_transform, 0, temperature, eps, 0.0, 1.0};
  k_score_weights<<<1, 1024, smem, (cudaStream_t)stream>>>(scores, unseen, n, a, score_weights);
  PCK(cudaGetLastError());
  return 0;
}

static int check_transform(int t) {
  return (t >= 0 && t <= 6) ? 0 : pfail(MGPLR_E_UNSUPPORTED, "transform must be 0 constant, 1 rank, 2 power, 3 softmax, 4 match, 5 match_rank or 6 eps_greedy");
}

extern "C" int mgplr_plr_sample_weights(const double *scores, const double *staleness, const double *unseen, int32_t n,
                                        int32_t score_transform, double temperature, double eps, double staleness_coef,
                                        int32_t staleness_transform, double staleness_temperature, const double *score_weights_in,
                                        double *weights, void *stream) {
  if (!scores || !staleness || !unseen || !weights || n < 1 || n > kMaxBuf)
    return pfail(MGPLR_E_BADARG, "mgplr_plr_sample_weights: bad arguments (n must be in [1, 8192])");
  if (int rc = check_transform(score_transform)) return rc;
  if (int rc = check_transform(staleness_transform)) return rc;
  if (int rc = ensure_dscratch(2 * (size_t)kMaxBuf)) return rc;
  const size_t smem = sort_smem(n);
  PCK(cudaFuncSetAttribute(k_sample_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const WeightArgs a{score_transform, staleness_transform, temperature, eps, staleness_coef, staleness_temperature};
  k_sample_weights<<<1, 1024, smem, (cudaStream_t)stream>>>(scores, staleness, unseen, n, a, weights, g_dscratch, score_weights_in);
  PCK(cudaGetLastError());
  return 0;
}

// n_draws sequential draws.  The rank part is fixed across the batch (scores do not change); the staleness
// part is recomputed after every draw ("all +1, chosen -> 0").  Inverse CDF: np.random.choice(p=w) =
// cumsum -> /cdf[-1] -> searchsorted(u, side='right') = #{cdf <= u}.
// Shared-memory staging of the per-draw working set: every draw is a dozen block-wide stages over the buffer arrays, and with
// the arrays in HBM each stage pays a global-memory round trip (10.7 us per draw at 4 000 slots; 4 096 replay draws of an ACCEL
// cycle: 44 ms).  For buffers up to kStageMax slots the staleness / seen-mask / rank-weight / weight arrays live in shared memory
// for the whole launch (same code, same operation order, same bits) and staleness is written back once at the end.
constexpr int kStageMax = 4096;

// ---- fast sequential draws (staged buffers, staleness transform power / temperature 1, or no staleness mix at all) ----
// Between the draws of one call only the staleness changes, and it changes in closed form: before draw t a seen slot that was
// never picked holds s0_i + t, a slot last picked at draw k holds t - k - 1.  With PA / PS / PN the prefix sums of the
// (normalised, masked) score weights, of s0 over seen slots and of the seen mask, and PC the prefix sum of the corrections
// s0_j + k_j + 1 of the picked slots, the cdf the reference builds with cumsum(sample_weights()) is
//     cdf(i) = (1 - c) PA(i) + c / S_t (PS(i) + t PN(i) - PC(i)),      S_t = S_0 + t nSeen - C_t
// so a draw is ONE descent over four Fenwick trees in shared memory (13 levels at 4 096 slots) plus a point update of the
// correction tree, by one thread: ~1 us per draw instead of ~5.5 us of block-wide passes.  The integer-valued parts are exact
// in double; the score part differs from a sequential cumsum in the last bits only, like the block scan of the general path.
// Preconditions (checked in the kernel; the general path runs otherwise): S_0 > 0 and at least two seen slots, which makes
// every S_t > 0 (the reference's all-zero-staleness fallback never triggers).
__device__ void block_inclusive_scan(const double *x, double *P, int n, double *wsum) {
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  double local = 0.0;
  for (int i = lo; i < hi; i++) local += x[i];
  double v = local;
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += y;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double w = (threadIdx.x < (blockDim.x >> 5)) ? wsum[threadIdx.x] : 0.0;
    for (int o = 1; o < 32; o <<= 1) {
      const double y = __shfl_up_sync(0xffffffffu, w, o);
      if (threadIdx.x >= o) w += y;
    }
    wsum[threadIdx.x] = w;
  }
  __syncthreads();
  double run = ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0.0) + v - local;
  for (int i = lo; i < hi; i++) { run += x[i]; P[i] = run; }
  __syncthreads();
}
// x[0..n) -> Fenwick tree in place (f[i], 1-based, stored at x[i-1]); tmp: n doubles of scratch
__device__ void fenwick_build(double *x, double *tmp, int n, double *wsum) {
  block_inclusive_scan(x, tmp, n, wsum);
  for (int i = threadIdx.x + 1; i <= n; i += blockDim.x) {
    const int lb = i & -i;
    x[i - 1] = tmp[i - 1] - (i - lb > 0 ? tmp[i - lb - 1] : 0.0);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) k_sample_replay(const double *scores, double *staleness_g, const double *unseen_g, int n,
                                                        WeightArgs a, const double *u, int n_draws, int32_t *out_index,
                                                        double *w_rank_g, double *weights_g, const double *cached, int staged,
                                                        int fast_allowed) {
  extern __shared__ __align__(16) uint8_t sm[];
  Key *keys = reinterpret_cast<Key *>(sm);
  __shared__ double red[33];
  __shared__ double wsum[32];
  __shared__ int s_pick;
  double *w_rank = w_rank_g, *weights = weights_g, *staleness = staleness_g;
  const double *unseen = unseen_g;
  if (cached) w_rank = const_cast<double *>(cached);
  else score_weights(scores, unseen_g, n, a, w_rank_g, keys, red);
  if (staged) {   // layout: [keys scratch (only used by a rank STALENESS transform) | w_rank | weights | staleness | unseen]
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    double *base = reinterpret_cast<double *>(sm + ((a.stale_transform == 1 || a.stale_transform == 5) ? (size_t)n2 * sizeof(Key) : 0));
    double *s_w = base, *s_wt = base + n, *s_st = base + 2 * (size_t)n, *s_un = base + 3 * (size_t)n;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) { s_w[i] = w_rank[i]; s_st[i] = staleness_g[i]; s_un[i] = unseen_g[i]; }
    __syncthreads();
    w_rank = s_w; weights = s_wt; staleness = s_st; unseen = s_un;
    if (fast_allowed) {
      // layout after the four staged arrays: [fC n doubles][last pick n int32][seen n bytes]
      double *fA = s_w, *fS = s_wt, *s0 = s_st, *fN = s_un, *fC = base + 4 * (size_t)n;
      int32_t *lastk = reinterpret_cast<int32_t *>(fC + n);
      uint8_t *seen = reinterpret_cast<uint8_t *>(lastk + n);
      const bool mix = a.coef > 0;
      double part_s = 0.0, part_n = 0.0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double sn = 1.0 - s_un[i];
        seen[i] = sn > 0.0;
        lastk[i] = -1;
        fS[i] = mix ? s0[i] * sn : 0.0;   // (power transform, temperature 1, eps 0: the weight IS the staleness; staleness >= 0)
        fN[i] = sn;
        part_s += fS[i]; part_n += sn;
      }
      double n_seen;
      const double S0 = block_sum2(part_s, part_n, red, n_seen);
      if (!mix || (S0 > 0.0 && n_seen >= 2.0)) {
        fenwick_build(fA, fC, n, wsum);
        const double TA = fC[n - 1];                 // sum of the score weights (1, or 0 when nothing is seen)
        __syncthreads();
        if (mix) { fenwick_build(fS, fC, n, wsum); fenwick_build(fN, fC, n, wsum); }
        for (int i = threadIdx.x; i < n; i += blockDim.x) fC[i] = 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
          int top = 1;
          while (top * 2 <= n) top *= 2;
          const double cA = 1.0 - a.coef;
          double Ct = 0.0;
          for (int t = 0; t < n_draws; t++) {
            const double St = S0 + (double)t * n_seen - Ct;
            const double cS = mix ? a.coef / St : 0.0, tt = (double)t;
            const double total = mix ? cA * TA + cS * St : TA;
            const double uu = u[t];
            int pos = 0;
            double aA = 0.0, aS = 0.0, aN = 0.0, aC = 0.0;
            for (int k = top; k; k >>= 1) {
              const int np = pos + k;
              if (np > n) continue;
              const double vA = aA + fA[np - 1];
              double cdf, vS = 0.0, vN = 0.0, vC = 0.0;
              if (mix) {
                vS = aS + fS[np - 1]; vN = aN + fN[np - 1]; vC = aC + fC[np - 1];
                cdf = cA * vA + cS * (vS + tt * vN - vC);
              } else cdf = vA;
              if (!(cdf / total > uu)) { pos = np; aA = vA; aS = vS; aN = vN; aC = vC; }
            }
            const int pick = min(pos, n - 1);   // pos leading slots have cdf <= u: the pick is slot pos (searchsorted side='right')
            out_index[t] = pick;
            if (mix) {
              if (seen[pick]) {
                const double corr_new = s0[pick] + tt + 1.0;
                const double corr_old = lastk[pick] >= 0 ? s0[pick] + (double)lastk[pick] + 1.0 : 0.0;
                const double delta = corr_new - corr_old;
                for (int i = pick + 1; i <= n; i += i & -i) fC[i - 1] += delta;
                Ct += delta;
              }
              lastk[pick] = t;
            }
          }
        }
        __syncthreads();
        if (mix)   // _update_staleness applied n_draws times (level_sampler.py:601-604): +1 everywhere, 0 at the pick
          for (int i = threadIdx.x; i < n; i += blockDim.x)
            staleness_g[i] = lastk[i] >= 0 ? (double)(n_draws - 1 - lastk[i]) : s0[i] + (double)n_draws;
        return;
      }
      // preconditions not met: restore the two staged arrays this block overwrote and take the general path
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) s_un[i] = unseen_g[i];
      __syncthreads();
    }
  }
  const double coef = a.coef;
  // contiguous chunk per thread so the scan is a per-thread serial cumsum + a block scan of chunk sums
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  for (int dr = 0; dr < n_draws; dr++) {
    mix_staleness(w_rank, staleness, unseen, n, a, weights, keys, red);
    double local = 0.0;
    for (int i = lo; i < hi; i++) local += weights[i];
    // block exclusive scan of `local`
    double x = local;
    for (int o = 1; o < 32; o <<= 1) {
      const double y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      double w = (threadIdx.x < (blockDim.x >> 5)) ? wsum[threadIdx.x] : 0.0;
      for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      wsum[threadIdx.x] = w;
    }
    if (threadIdx.x == 0) s_pick = n;  // searchsorted returns n if u >= cdf[-1]
    __syncthreads();
    const double total = wsum[(blockDim.x >> 5) - 1];
    double run = ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0.0) + x - local;
    const double uu = u[dr];
    // first index whose normalised cdf exceeds u
    for (int i = lo; i < hi; i++) {
      run += weights[i];
      if (run / total > uu) { atomicMin(&s_pick, i); break; }
    }
    __syncthreads();
    const int pick = min(s_pick, n - 1);
    if (threadIdx.x == 0) out_index[dr] = pick;
    if (coef > 0) {  // _update_staleness (level_sampler.py:601-604)
      for (int i = threadIdx.x; i < n; i += blockDim.x) staleness[i] = (i == pick) ? 0.0 : staleness[i] + 1.0;
    }
    __syncthreads();
  }
  if (staged && coef > 0)
    for (int i = threadIdx.x; i < n; i += blockDim.x) staleness_g[i] = staleness[i];
}

extern "C" int mgplr_plr_sample_replay(const double *scores, double *staleness, const double *unseen, int32_t n,
                                       int32_t score_transform, double temperature, double eps, double staleness_coef,
                                       int32_t staleness_transform, double staleness_temperature, const double *score_weights_in,
                                       const double *u, int32_t n_draws, int32_t *out_index, void *stream) {
  if (!scores || !staleness || !unseen || !u || !out_index || n < 1 || n > kMaxBuf || n_draws < 1)
    return pfail(MGPLR_E_BADARG, "mgplr_plr_sample_replay: bad arguments (n must be in [1, 8192])");
  if (int rc = check_transform(score_transform)) return rc;
  if (int rc = check_transform(staleness_transform)) return rc;
  if (int rc = ensure_dscratch(2 * (size_t)kMaxBuf)) return rc;
  const int staged = n <= kStageMax;
  // MGPLR_REPLAY_FAST: 0 = always the general block-wide path, 1 (default) = Fenwick path for calls of >= 4 draws, 2 = whenever legal
  const int fast_knob = getenv("MGPLR_REPLAY_FAST") ? atoi(getenv("MGPLR_REPLAY_FAST")) : 1;
  const int fast = staged && fast_knob > 0 && (fast_knob > 1 || n_draws >= 4) &&
                   (!(staleness_coef > 0) || (staleness_transform == 2 && staleness_temperature == 1.0));
  size_t smem = sort_smem(n);
  if (staged) smem = ((staleness_transform == 1 || staleness_transform == 5) ? smem : 0) + 4 * (size_t)n * sizeof(double);
  if (fast) smem += (size_t)n * (sizeof(double) + sizeof(int32_t) + 1) + 16;
  if (!score_weights_in && smem < sort_smem(n)) smem = sort_smem(n);
  PCK(cudaFuncSetAttribute(k_sample_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const WeightArgs a{score_transform, staleness_transform, temperature, eps, staleness_coef, staleness_temperature};
  k_sample_replay<<<1, 1024, smem, (cudaStream_t)stream>>>(scores, staleness, unseen, n, a, u, n_draws, out_index, g_dscratch,
                                                          g_dscratch + kMaxBuf, score_weights_in, staged, fast);
  PCK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ sampler bookkeeping
// LevelSampler.update_with_rollouts after the score reductions: the walk over the episode records in canonical (actor-major,
// time-minor) order that applies the EWA score update to working seeds and admits staging seeds into the working buffer
// (level_sampler.py:185-273: update_seed_score, _partial_update_seed_score(_buffer), _next_buffer_index), on the HBM
// resident buffer arrays.  The walk is order dependent (an admission changes which slot the next one may evict), so it is ONE
// CTA: thread 0 interprets the record stream from a shared-memory chunk; when a staging seed meets a FULL buffer the whole CTA
// evaluates argmin(sample_weights()) on the current state (the same device functions as k_sample_weights, so the weights are
// the bits the host API would get) and hands the slot back.  The rank transform does not re-sort per admission: ranks are
// kept incrementally -- a slot whose score changed is "dirty" and one pass fixes every rank against the old and new keys of the
// dirty slots (rank_i = 1 + #{k : key_k before key_i}) -- with a full bitonic sort only when more than kMaxDirty slots changed.
// Seeds are addressed through a small table of the rollout's distinct seeds (sorted; built by the host from its dict / set
// views): cur_idx = seed2index entry or -1 (entries are never deleted, level_sampler.py:250), stamp = staging timestamp or -1.
constexpr int kMaxDirty = 32, kRecChunk = 128;

struct ApplyArgs {
  const mgplr_episode *rec;
  const int32_t *n_rec_dev;   // record count on the device (written by the score kernel) or NULL
  int32_t n_rec, max_rec;
  const double *pre;          // optional [n_rec][4]: partial score, partial max, partial steps, unused (not-done tails merged by the host)
  const int64_t *useeds;      // [n_u] sorted distinct seeds
  int32_t n_u;
  int32_t *uid;               // [max_rec] scratch: index into useeds or -1
  int32_t *cur_idx;           // [n_u] in/out
  const double *stamp;        // [n_u] staging timestamp (seed2timestamp_buffer) or -1
  int32_t *status;            // [n_u] out: 0 untouched, 1 admitted, 2 rejected (left the staging set either way)
  int32_t *adm_log;           // [n_u][2] out: (table index, slot) of the admissions, in order
  int32_t *counters;          // [4]: n admissions, working_seed_buffer_size (in/out), not-done tails seen, records walked
  double *scores, *stale, *unseen, *grounded;
  int64_t *seeds;
  int32_t n_buf;
  double running_count, alpha, max_coef;
  int32_t kind;               // 0: record scores, 1: uniform, 2: grounded (MaxMC)
  int32_t priority;           // 0: replay_support (argmin of sample_weights), 1: lowest score
  WeightArgs wa;
  double *w_score, *weights, *table, *rank_score;
  int32_t *rank;
};

__global__ void k_record_uid(ApplyArgs a) {
  const int n = a.n_rec_dev ? min(*a.n_rec_dev, a.max_rec) : a.n_rec;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t s = a.rec[r].seed;
  int lo = 0, hi = a.n_u;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a.useeds[mid] < s) lo = mid + 1; else hi = mid;
  }
  a.uid[r] = (lo < a.n_u && a.useeds[lo] == s) ? lo : -1;
}

__device__ __forceinline__ bool key_before(double sa, int ia, double sb, int ib) { return sa > sb || (sa == sb && ia > ib); }

__global__ void __launch_bounds__(1024) k_apply_records(ApplyArgs a, int staged) {
  extern __shared__ __align__(16) uint8_t sm[];
  Key *keys = reinterpret_cast<Key *>(sm);
  __shared__ double red[33];
  __shared__ mgplr_episode s_rec[kRecChunk];
  __shared__ int s_uid[kRecChunk];
  __shared__ int s_req, s_next, s_slot, s_chunk_lo, s_have_slot, s_rank_valid, s_ndirty;
  __shared__ int s_dirty[kMaxDirty], s_cnt[kMaxDirty];
  __shared__ double s_old[kMaxDirty], s_new[kMaxDirty];
  __shared__ double s_min[32];
  __shared__ int s_arg[32];
  const int tid = threadIdx.x, N = a.n_buf;
  const int n = a.n_rec_dev ? min(*a.n_rec_dev, a.max_rec) : a.n_rec;
  const bool ranked = a.wa.score_transform == 1 && a.priority == 0;
  if (tid == 0) { s_next = 0; s_chunk_lo = -kRecChunk; s_have_slot = 0; s_rank_valid = 0; s_ndirty = 0; }
  // working set of an admission (every one is a dozen block-wide stages): in shared memory when the buffer fits (kStageMax),
  // laid out [sort keys, later w_score | weights] [staleness] [unseen] [rank-weight table] [ranks]; else in the HBM scratch
  double *w_score = a.w_score, *weights = a.weights, *stale = a.stale, *unseen = a.unseen, *table = a.table;
  int32_t *rank = a.rank;
  if (staged) {
    int n2 = 1;
    while (n2 < N) n2 <<= 1;
    w_score = reinterpret_cast<double *>(sm);
    weights = w_score + N;
    stale = reinterpret_cast<double *>(sm + (size_t)n2 * sizeof(Key));
    unseen = stale + N;
    table = unseen + N;
    rank = reinterpret_cast<int32_t *>(table + N);
    for (int i = tid; i < N; i += blockDim.x) { stale[i] = a.stale[i]; unseen[i] = a.unseen[i]; }
  }
  if (ranked) {
    const double inv_t = 1.0 / a.wa.temperature;
    for (int r = tid; r < N; r += blockDim.x) table[r] = 1.0 / pow((double)(r + 1), inv_t);
  }
  __syncthreads();
  int n_adm = a.counters[0], filled = a.counters[1], n_tail = 0;   // (meaningful in thread 0)
  for (;;) {
    if (tid == 0) {
      int r = s_next, req = 0;
      while (r < n) {
        if (r < s_chunk_lo || r >= s_chunk_lo + kRecChunk) { req = 2; break; }
        const mgplr_episode e = s_rec[r - s_chunk_lo];
        const int u = s_uid[r - s_chunk_lo];
        if (e.cliffhanger == 1 || u < 0) { r++; continue; }                 // cliffhangers are skipped (level_sampler.py:527-528)
        if (e.cliffhanger == 2) { n_tail++; r++; continue; }                // not-done tail: host bookkeeping (never under the runner)
        const int steps = e.t_end - e.t_start;
        const double dn = (double)steps;
        int idx = a.cur_idx[u];
        double score = (double)e.mean_score, mx = (double)e.max_score, gv = 0.0;
        double p_score = 0.0, p_max = -INFINITY, p_steps = 0.0;
        bool flush = false;  // after_update: what is left of an unfinished episode, scored as it stands (level_sampler.py:580-599)
        if (a.pre) {
          p_score = a.pre[4 * (size_t)r]; p_max = a.pre[4 * (size_t)r + 1]; p_steps = a.pre[4 * (size_t)r + 2];
          flush = a.pre[4 * (size_t)r + 3] != 0.0;
        }
        if (flush) { score = 0.0; mx = -INFINITY; }
        else if (a.kind == 1) { score = 1.0; mx = 1.0; }
        else if (a.kind == 2) {  // _average_grounded_signed_value_loss (level_sampler.py:351-386) from the episode sums
          gv = (double)e.reward_sum;
          if (idx >= 0) gv = fmax(a.grounded[idx], gv);
          score = __dmul_rn(__ddiv_rn(__dadd_rn(p_steps, dn), dn), __dsub_rn(gv, __ddiv_rn((double)e.value_sum, dn)));
          mx = __dsub_rn(gv, (double)e.value_min);
        }
        // (explicitly rounded operations: an FMA contraction would round differently from the reference's Python floats)
        const double merged = __dadd_rn(p_score, __ddiv_rn(__dmul_rn(__dsub_rn(score, p_score), dn), __dadd_rn(p_steps, dn)));
        const bool staged = a.stamp[u] >= 0.0 && a.status[u] == 0;
        if (staged) {   // _partial_update_seed_score_buffer(done=True), level_sampler.py:228-273
          int slot;
          if (filled < N) slot = filled;
          else if (s_have_slot) { slot = s_slot; s_have_slot = 0; }
          else { req = 1; break; }
          if (a.scores[slot] <= merged || unseen[slot] > 0.0) {
            unseen[slot] = 0.0;
            a.seeds[slot] = a.useeds[u];
            a.cur_idx[u] = slot;
            a.scores[slot] = merged;
            stale[slot] = __dsub_rn(a.running_count, a.stamp[u]);
            filled = min(filled + 1, N);
            if (a.kind == 2 && !flush) a.grounded[slot] = gv;
            a.status[u] = 1;
            a.adm_log[2 * n_adm] = u; a.adm_log[2 * n_adm + 1] = slot;
            n_adm++;
            idx = slot;
          } else { a.status[u] = 2; idx = -1; }
        } else if (idx >= 0) {   // _partial_update_seed_score(done=True), level_sampler.py:193-216
          unseen[idx] = 0.0;
          const double total = __dadd_rn(__dmul_rn(a.max_coef, fmax(p_max, mx)), __dmul_rn(__dsub_rn(1.0, a.max_coef), merged));
          a.scores[idx] = __dadd_rn(__dmul_rn(__dsub_rn(1.0, a.alpha), a.scores[idx]), __dmul_rn(a.alpha, total));
          if (a.kind == 2 && !flush) a.grounded[idx] = gv;
        }
        if (idx >= 0 && ranked && s_rank_valid) {   // the slot's rank key changed
          int k = 0;
          while (k < s_ndirty && s_dirty[k] != idx) k++;
          if (k == s_ndirty) {
            if (s_ndirty < kMaxDirty) s_dirty[s_ndirty++] = idx;
            else s_rank_valid = 0;                   // too many: full sort at the next admission
          }
        }
        r++;
      }
      s_next = r; s_req = req;
    }
    __syncthreads();
    const int req = s_req;
    if (req == 0) break;
    if (req == 2) {   // next chunk of records into shared memory
      const int lo = s_next;
      for (int i = tid; i < kRecChunk; i += blockDim.x)
        if (lo + i < n) { s_rec[i] = a.rec[lo + i]; s_uid[i] = a.uid[lo + i]; }
      __syncthreads();
      if (tid == 0) s_chunk_lo = lo;
      __syncthreads();
      continue;
    }
    // ---- req == 1: slot = _next_buffer_index on a full buffer (level_sampler.py:218-226), all threads
    if (a.priority == 0) {
      if (ranked) {
        if (!s_rank_valid) {   // full sort (same key order as transform_vals)
          int n2 = 1;
          while (n2 < N) n2 <<= 1;
          for (int i = tid; i < n2; i += blockDim.x) {
            Key k;
            k.s = (i < N) ? a.scores[i] : -INFINITY; k.i = (i < N) ? i : -1 - i;
            keys[i] = k;
          }
          __syncthreads();
          block_sort_desc(keys, n2);
          for (int r = tid; r < N; r += blockDim.x) { rank[keys[r].i] = r + 1; }
          for (int i = tid; i < N; i += blockDim.x) a.rank_score[i] = a.scores[i];
          __syncthreads();
          if (tid == 0) { s_rank_valid = 1; s_ndirty = 0; }
          __syncthreads();
        } else if (s_ndirty) {
          const int D = s_ndirty;
          if (tid < D) { s_old[tid] = a.rank_score[s_dirty[tid]]; s_new[tid] = a.scores[s_dirty[tid]]; s_cnt[tid] = 0; }
          __syncthreads();
          for (int i = tid; i < N; i += blockDim.x) {
            bool dirty = false;
            for (int d = 0; d < D; d++) dirty |= (s_dirty[d] == i);
            const double si = dirty ? a.scores[i] : a.rank_score[i];   // the slot's FINAL key
            if (!dirty) {
              int delta = 0;
              for (int d = 0; d < D; d++)
                delta += (int)key_before(s_new[d], s_dirty[d], si, i) - (int)key_before(s_old[d], s_dirty[d], si, i);
              if (delta) rank[i] += delta;
            }
            for (int d = 0; d < D; d++)   // slots ahead of dirty slot d under the final keys
              if (s_dirty[d] != i && key_before(si, i, s_new[d], s_dirty[d])) atomicAdd(&s_cnt[d], 1);
          }
          __syncthreads();
          if (tid < D) { rank[s_dirty[tid]] = 1 + s_cnt[tid]; a.rank_score[s_dirty[tid]] = s_new[tid]; }
          __syncthreads();
          if (tid == 0) s_ndirty = 0;
          __syncthreads();
        }
        for (int i = tid; i < N; i += blockDim.x) w_score[i] = table[rank[i] - 1];
        __syncthreads();
        mask_normalise(w_score, unseen, N, true, red);
      } else {
        score_weights(a.scores, unseen, N, a.wa, w_score, keys, red);
      }
      mix_staleness(w_score, stale, unseen, N, a.wa, weights, keys, red);
    }
    {   // first index of the minimum (numpy argmin)
      const double *v = a.priority == 0 ? weights : a.scores;
      double best = INFINITY;
      int arg = 0x7fffffff;
      for (int i = tid; i < N; i += blockDim.x) {
        const double x = v[i];
        if (x < best || (x == best && i < arg)) { best = x; arg = i; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oa = __shfl_down_sync(0xffffffffu, arg, o);
        if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
      }
      if ((tid & 31) == 0) { s_min[tid >> 5] = best; s_arg[tid >> 5] = arg; }
      __syncthreads();
      if (tid < 32) {
        best = (tid < (int)(blockDim.x >> 5)) ? s_min[tid] : INFINITY;
        arg = (tid < (int)(blockDim.x >> 5)) ? s_arg[tid] : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_down_sync(0xffffffffu, best, o);
          const int oa = __shfl_down_sync(0xffffffffu, arg, o);
          if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (tid == 0) { s_slot = arg; s_have_slot = 1; }
      }
      __syncthreads();
    }
  }
  if (staged) {
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) { a.stale[i] = stale[i]; a.unseen[i] = unseen[i]; }
  }
  if (tid == 0) { a.counters[0] = n_adm; a.counters[1] = filled; a.counters[2] = n_tail; a.counters[3] = n; }
}

// host launcher: every pointer is a device pointer; scratch = [4 * n_buf] doubles + [n_buf] int32 owned by the caller
extern "C" int mgplr_plr_apply_records(const mgplr_episode *records, const int32_t *n_records_dev, int32_t n_records,
                                       int32_t max_records, const double *pre, const int64_t *table_seeds, int32_t n_table,
                                       int32_t *table_index, const double *table_stamp, int32_t *table_status,
                                       int32_t *admission_log, int32_t *counters, int32_t *record_scratch, double *scores,
                                       double *staleness, double *unseen, double *grounded, int64_t *seeds, int32_t n_buf,
                                       double running_sample_count, double alpha, double max_score_coef, int32_t score_kind,
                                       int32_t priority, int32_t score_transform, double temperature, double eps,
                                       double staleness_coef, int32_t staleness_transform, double staleness_temperature,
                                       double *scratch_f64, int32_t *scratch_i32, void *stream) {
  if (!records || !table_seeds || !table_index || !table_stamp || !table_status || !admission_log || !counters || !record_scratch ||
      !scores || !staleness || !unseen || !seeds || !scratch_f64 || !scratch_i32 || n_buf < 1 || n_buf > kMaxBuf || n_table < 0 ||
      max_records < 0 || (score_kind == 2 && !grounded))
    return pfail(MGPLR_E_BADARG, "mgplr_plr_apply_records: bad arguments (seed_buffer_size must be in [1, 8192])");
  if (int rc = check_transform(score_transform)) return rc;
  if (int rc = check_transform(staleness_transform)) return rc;
  ApplyArgs a;
  a.rec = records; a.n_rec_dev = n_records_dev; a.n_rec = n_records; a.max_rec = max_records; a.pre = pre;
  a.useeds = table_seeds; a.n_u = n_table; a.uid = record_scratch; a.cur_idx = table_index; a.stamp = table_stamp;
  a.status = table_status; a.adm_log = admission_log; a.counters = counters;
  a.scores = scores; a.stale = staleness; a.unseen = unseen; a.grounded = grounded; a.seeds = seeds; a.n_buf = n_buf;
  a.running_count = running_sample_count; a.alpha = alpha; a.max_coef = max_score_coef; a.kind = score_kind; a.priority = priority;
  a.wa = WeightArgs{score_transform, staleness_transform, temperature, eps, staleness_coef, staleness_temperature};
  a.w_score = scratch_f64; a.weights = scratch_f64 + n_buf; a.table = scratch_f64 + 2 * (size_t)n_buf;
  a.rank_score = scratch_f64 + 3 * (size_t)n_buf; a.rank = scratch_i32;
  cudaStream_t st = (cudaStream_t)stream;
  const int upper = n_records_dev ? max_records : n_records;
  if (upper > 0 && n_table > 0) {
    k_record_uid<<<(upper + 255) / 256, 256, 0, st>>>(a);
    PCK(cudaGetLastError());
  }
  const int staged = n_buf <= kStageMax && staleness_transform != 1 && staleness_transform != 5 && score_transform != 5;  // (those sort in the aliased key region)
  const size_t smem = sort_smem(n_buf) + (staged ? (size_t)n_buf * (3 * sizeof(double) + sizeof(int32_t)) : 0);
  PCK(cudaFuncSetAttribute(k_apply_records, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_apply_records<<<1, 1024, smem, st>>>(a, staged);
  PCK(cudaGetLastError());
  return 0;
}
