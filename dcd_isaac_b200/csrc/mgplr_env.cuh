// mgplr_env.cuh -- device-side MultiGrid adversarial-maze semantics for sm_100a.
//
// State layout in HBM: the maze is a bit-packed uint8 grid -- one wall bit per
// cell, row y of env e in the 32-bit word wall[y*N + e] (rows are struct-of-arrays across envs so a
// warp of 32 envs reads 128 contiguous bytes per row) -- plus one 16-byte "hot" record per env with
// the agent / goal / start coordinates and the episode counters.  Cell codes other than wall are
// coordinates, not cells: goal (gx,gy), agent (ax,ay,dir), start (sx,sy,sdir).
//
// Every function cites the reference file:line whose behaviour it reproduces (paths relative to the
// reference root); gym-minigrid 1.0.1 / gym 0.15.7 / numpy legacy RandomState pieces are third-party
// and are implemented from their published algorithms.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mgplr.h"

namespace mgplr {

constexpr int kV = 5;             // agent_view_size of every registered adversarial env (adversarial.py:71)
constexpr int kNone = 63;         // coordinate sentinel for "None"
constexpr int kObsFloats = 3 * kV * kV;
constexpr uint32_t kErrRetries = 1u, kErrNoStart = 2u, kErrBadLoc = 4u;

struct Cfg {
  int W, max_steps, max_episode_steps, see_through, n_clutter, resample, goal_last, fixed_env, n_editor;
};

// Device view of one venv (all pointers are HBM).
struct Dev {
  int N;
  int use_tma;       // 1: TMA bulk copies for the tile rows / observation tile; 0: cp.async + vector stores (A/B knob)
  int l2_hints;      // bit mask of the step kernel's L2 hints (mgplr_venv_create): 5 = eviction priorities + streaming scalar stores
  Cfg c;
  uint32_t *wall;    // [ceil(N/32)][W][32] wall bit-plane rows (bit x of row y), tile-major: see env_rows()
  uint4 *hot;        // [N] x: ax6|ay6|adir2|has1|done1|step16  y: gx5|gy5|hasgoal1|sx5|sy5|hasstart1|sdir2|pending8  z: elapsed16|eplen16  w: ep_ret bits
  uint32_t *adv;     // [N] adversary_step_count12 | adversary_max_steps12 | n_clutter_sampled1
  int4 *metrics;     // [N] n_clutter_placed, distance_to_goal, passable, shortest_path_length
  uint32_t *mt;      // [ceil(N/32)][39][32][16] MT19937 state (numpy RandomState of each env), see mt_at()
  uint32_t *mti;     // [N] index of the next word to generate (incremental twist), 0..623
  uint32_t *limbs;   // [3][N] seed limbs lo, hi, count (re-seed of fixed_environment)
  uint32_t *words;   // [N] MT words consumed since seeding
  uint32_t *err;     // [N] sticky error bits
  // speculative next-level generation of the DR auto-reset (step_env(reset_random=True)), see the "speculation" block
  uint32_t *spec;    // [N] bits 16+2p / 17+2p: candidate 0 (episode ended without a goal) / 1 (ended at the goal) built for a
                     //     level epoch of parity p is valid; bits 20-26: level epoch (bumped by every DR reset)
  uint32_t *cand;    // [N][2 epochs][2][W + 8 + 192] candidate records: wall rows, packed goal/start, -, words consumed, error
                     //     bits, and the NEW MT state words of the consumed span (so that a commit copies, never recomputes)
  uint2 *rr_list;    // [2][2N] regeneration jobs {env << 8 | candidate << 7 | epoch, MT cursor}, one list per launch parity
  unsigned long long *prof;  // debug counters (MGPLR_RR_PROF), else NULL
  uint32_t *sched;   // [8] regeneration phase: [0..1] jobs appended to list p, [2..3] next ticket of list p, [4] warps exited,
                     //     [5] parity of the running / next DR launch (device-side so that graph replays stay consistent)
};
// MT19937 state layout: tile-major AND chunked -- the 624 words of 32 consecutive envs form one 78 KB block laid out as
// [39 chunks][32 envs][16 words]: word i of env e at ((e/32*39 + i/16)*32 + e%32)*16 + i%16.  Both access patterns of this
// code stay sector-efficient: (a) lane-parallel code (32 lanes = 32 consecutive envs) generates 16-33 CONSECUTIVE words
// per refill, i.e. each lane walks a few 64-byte chunks of its own; (b) warp-per-env code (cooperative rebuild, the
// look-ahead window of a regeneration job) reads runs of 32-225 consecutive words of ONE env = a handful of 64-byte
// chunks instead of one 128-byte line per word.  (History: [624][N] put consecutive words of an env N*4 bytes apart;
// [tile][624][32] fixed the page locality but still cost one line per word on the warp-per-env paths.)
constexpr int kMtChunk = 16, kMtChunks = 624 / kMtChunk;  // 39 chunks of 16 words
__host__ __device__ inline size_t mt_at(int e, uint32_t i) {
  return ((((size_t)(e >> 5) * kMtChunks + (i >> 4)) * 32 + (size_t)(e & 31)) << 4) + (i & 15u);
}
constexpr int kMetricsDirty = -2;  // metrics.z of an env whose level came from a candidate record: recomputed by the getter
constexpr uint32_t kSpecValidMask = 15u << 16;
constexpr int kSpecEpochShift = 20;
// validity bit of candidate k for a level epoch ep: a job sets it with one fire-and-forget atomic OR.  Bits are per epoch
// PARITY and every reset stores a fresh word, so a bit set late by a job of the previous level (it runs during the launch
// that may reset the env again) lands on the other parity and is never consulted.
__host__ __device__ inline uint32_t spec_valid_bit(uint32_t ep, int k) { return 1u << (16 + 2 * (ep & 1u) + k); }
constexpr uint32_t kSpecEpochMask = 127u << kSpecEpochShift;
__host__ __device__ inline uint32_t spec_epoch(uint32_t sp) { return (sp & kSpecEpochMask) >> kSpecEpochShift; }
constexpr int kSpecWindow = 224;   // MT words one regeneration job can look ahead (< 227: all computable from the present state)
constexpr int kSpecState = 192;     // new MT state words a candidate record carries (a record that consumed more is not published)
__host__ __device__ inline int cand_words(int W) { return W + 8 + kSpecState; }

// `pending` = goal respawns (multigrid.py:821-838) whose env-RNG draws have not been made yet.  A respawn draw
// depends only on the level (walls + goal; the agent is off the grid while it is drawn) and its result is
// overwritten by reset_agent before anything observes it, so the step kernel only counts it and the draws are
// replayed -- in order, against the unchanged level -- by flush_pending() before the next consumer of the env
// RNG or the next level edit.  This keeps serial MT19937 traffic out of the hot kernel.
constexpr int kMaxPending = 255;
struct Env {
  int ax, ay, adir, has_agent, done_flag, step_count;
  int gx, gy, sx, sy, sdir, pending;
  int elapsed, ep_len;
  float ep_ret;
};

__device__ __forceinline__ Env unpack(const uint4 h) {
  Env e;
  e.ax = h.x & 63; e.ay = (h.x >> 6) & 63; e.adir = (h.x >> 12) & 3; e.has_agent = (h.x >> 14) & 1;
  e.done_flag = (h.x >> 15) & 1; e.step_count = h.x >> 16;
  const bool hg = (h.y >> 10) & 1, hs = (h.y >> 21) & 1;
  e.gx = hg ? (int)(h.y & 31) : kNone; e.gy = hg ? (int)((h.y >> 5) & 31) : kNone;
  e.sx = hs ? (int)((h.y >> 11) & 31) : kNone; e.sy = hs ? (int)((h.y >> 16) & 31) : kNone;
  e.sdir = (h.y >> 22) & 3; e.pending = h.y >> 24;
  e.elapsed = h.z & 0xffff; e.ep_len = h.z >> 16;
  e.ep_ret = __uint_as_float(h.w);
  return e;
}
__device__ __forceinline__ uint4 pack(const Env &e) {
  uint4 h;
  h.x = (uint32_t)e.ax | ((uint32_t)e.ay << 6) | ((uint32_t)e.adir << 12) | ((uint32_t)e.has_agent << 14) |
        ((uint32_t)e.done_flag << 15) | ((uint32_t)e.step_count << 16);
  const uint32_t hg = e.gx != kNone, hs = e.sx != kNone;
  h.y = ((uint32_t)e.gx & 31u) | (((uint32_t)e.gy & 31u) << 5) | (hg << 10) | (((uint32_t)e.sx & 31u) << 11) |
        (((uint32_t)e.sy & 31u) << 16) | (hs << 21) | ((uint32_t)e.sdir << 22) | ((uint32_t)e.pending << 24);
  h.z = ((uint32_t)e.elapsed & 0xffff) | ((uint32_t)e.ep_len << 16);
  h.w = __float_as_uint(e.ep_ret);
  return h;
}

// Wall rows of one env.  In HBM the bit-plane is tile-major: the W rows of 32 consecutive envs form one contiguous
// W*128-byte block (row y of env e at wall[((e/32)*W + y)*32 + e%32]), so a warp's tile is ONE bulk copy and a
// shared-memory image of it is addressed with the same stride (bank index = lane).
struct Rows {
  uint32_t *p;
  int stride;
  __device__ __forceinline__ uint32_t get(int r) const { return p[r * stride]; }
  __device__ __forceinline__ void set(int r, uint32_t v) const { p[r * stride] = v; }
};

__device__ __forceinline__ Rows env_rows(const Dev &d, int e) {
  return Rows{d.wall + ((size_t)(e >> 5) * d.c.W) * 32 + (e & 31), 32};
}

// ---------------------------------------------------------------------------------------------
// numpy legacy RandomState == MT19937, generated incrementally: word i of the next generation is
// x[i+397] ^ twist(x[i], x[i+1]) and only depends on already-updated words when produced in order, so
// producing words in place, in order, is output-identical to the batch twist and never stalls a lane for 624
// dependent iterations.
// Draws are generated SPECULATIVELY IN BATCHES of 16: the words at offsets 0..15 from the cursor are mutually
// independent (word j reads state[j], state[j+1], state[j+397]; none is produced inside a batch shorter than 227), so
// one refill issues 33 independent loads -- one memory round trip instead of one per draw.  Tempered outputs and new
// state words of the batch live in a per-thread column of SHARED memory ([32][threads]: rows 0..15 outputs, 16..31
// state words; bank = thread), so a draw is two conflict-free LDS and one STG committing that word's state.
// Unconsumed words are simply regenerated by the next refill, so the stream is exactly numpy's.  Refills are
// WARP-UNIFORM: when any converged lane runs dry every converged lane refills from its own cursor, which keeps the
// refill code from being replayed once per lane under divergence (it was 75 % of reset_random's instructions).
struct Rng {
  static constexpr int kBatch = 16;
  uint32_t *mt, *mti_p, *words_p, *spec_p;
  int N, e;
  uint32_t idx, used;  // cursor (0..623), words consumed since seeding
  int have, pos;       // words in the batch, next unread word
  bool loaded;
  uint32_t *buf;       // this thread's column of the shared scratch
  int stride;          // threads per row of the scratch
  __device__ __forceinline__ Rng(const Dev &dev, int env, uint32_t *buf_, int stride_)
      : mt(dev.mt), mti_p(dev.mti), words_p(dev.words), spec_p(dev.spec), N(dev.N), e(env), idx(0), used(0), have(0), pos(0),
        loaded(false), buf(buf_), stride(stride_) {}
  // Any consumer of the env RNG other than the DR speculation itself drops the env's candidates: their stream position
  // is about to become stale.
  __device__ __forceinline__ void load() {
    if (!loaded) {
      idx = mti_p[e]; used = words_p[e]; loaded = true;
      const uint32_t sp = spec_p[e];
      if (sp & kSpecValidMask) spec_p[e] = sp & kSpecEpochMask;
    }
  }
  // (kept inline: out of line it shrinks the DR step kernel from 20 k to 12 k instructions, which changed nothing there, but
  // the lane-parallel reset_random kernel got 40 % slower)
  __device__ __forceinline__ void refill() {
    load();
    uint32_t b[kBatch], c[kBatch];
    uint32_t a = mt[mt_at(e, idx)];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      uint32_t i1 = idx + u + 1, im = idx + u + 397;
      if (i1 >= 624) i1 -= 624;
      if (im >= 624) im -= 624;
      if (im >= 624) im -= 624;
      b[u] = mt[mt_at(e, i1)];
      c[u] = mt[mt_at(e, im)];
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      uint32_t y = (a & 0x80000000u) | (b[u] & 0x7fffffffu);
      y = c[u] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      buf[(kBatch + u) * stride] = y;
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      buf[u * stride] = y;
      a = b[u];
    }
    have = kBatch; pos = 0;
  }
  __device__ __forceinline__ void store() {
    if (loaded) { mti_p[e] = idx; words_p[e] = used; }
  }
  // forget everything (after a re-seed replaced the stream)
  __device__ __forceinline__ void reset() { loaded = false; have = 0; pos = 0; }
  __device__ __forceinline__ bool ok() const { return true; }
  __device__ __forceinline__ uint32_t next() {
    if (__any_sync(__activemask(), pos >= have)) refill();
    const uint32_t v = buf[pos * stride];
    mt[mt_at(e, idx)] = buf[(kBatch + pos) * stride];  // commit this word's new state
    idx = (idx + 1 == 624) ? 0 : idx + 1;
    used++;
    pos++;
    return v;
  }
  __device__ __forceinline__ int randint(int lo, int hi) {
    const uint32_t rng = (uint32_t)(hi - lo - 1);
    if (rng == 0) return lo;
    const uint32_t mask = 0xffffffffu >> __clz(rng);
    uint32_t v;
    do { v = next() & mask; } while (v > rng);
    return lo + (int)v;
  }
};

// One word per draw, no scratch: for the rare paths of the hot kernel that cannot spare shared memory.
struct RngSlow {
  uint32_t *mt, *mti_p, *words_p, *spec_p;
  int N, e;
  uint32_t idx, used;
  bool loaded;
  __device__ __forceinline__ RngSlow(uint32_t *mt_, uint32_t *mti_, uint32_t *words_, uint32_t *spec_, int N_, int env)
      : mt(mt_), mti_p(mti_), words_p(words_), spec_p(spec_), N(N_), e(env), idx(0), used(0), loaded(false) {}
  __device__ __forceinline__ void load() {
    if (!loaded) {
      idx = mti_p[e]; used = words_p[e]; loaded = true;
      const uint32_t sp = spec_p[e];
      if (sp & kSpecValidMask) spec_p[e] = sp & kSpecEpochMask;  // see Rng::load
    }
  }
  __device__ __forceinline__ uint32_t next_raw() {
    const uint32_t i = idx, i1 = (i + 1 == 624) ? 0 : i + 1, im = (i + 397 >= 624) ? i + 397 - 624 : i + 397;
    const uint32_t a = mt[mt_at(e, i)], b = mt[mt_at(e, i1)], c = mt[mt_at(e, im)];
    uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    y = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    mt[mt_at(e, i)] = y;
    idx = i1; used++;
    return y;
  }
  __device__ __forceinline__ void store() {
    if (loaded) { mti_p[e] = idx; words_p[e] = used; }
  }
  __device__ __forceinline__ void reset() { loaded = false; }
  __device__ __forceinline__ bool ok() const { return true; }
  __device__ __forceinline__ uint32_t next() {
    load();
    uint32_t y = next_raw();
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  __device__ __forceinline__ int randint(int lo, int hi) {
    const uint32_t rng = (uint32_t)(hi - lo - 1);
    if (rng == 0) return lo;
    const uint32_t mask = 0xffffffffu >> __clz(rng);
    uint32_t v;
    do { v = next() & mask; } while (v > rng);
    return lo + (int)v;
  }
};

// MT19937 init_by_array for one env (RandomState.seed([lo, hi])), SoA state.
static __device__ __noinline__ void mt_seed(const Dev &d, int e, uint32_t k0, uint32_t k1, int klen) {
  uint32_t *mt = d.mt;
  uint32_t prev = 19650218u;
  mt[mt_at(e, 0)] = prev;
#pragma unroll 1
  for (int i = 1; i < 624; i++) { prev = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i; mt[mt_at(e, i)] = prev; }
  int i = 1, j = 0;
  prev = mt[mt_at(e, 0)];
#pragma unroll 1
  for (int k = 624; k; k--) {
    const uint32_t key = (j == 0) ? k0 : k1;
    uint32_t v = (mt[mt_at(e, i)] ^ ((prev ^ (prev >> 30)) * 1664525u)) + key + (uint32_t)j;
    mt[mt_at(e, i)] = v; prev = v;
    i++; j++;
    if (i >= 624) { mt[mt_at(e, 0)] = prev; i = 1; }
    if (j >= klen) j = 0;
  }
#pragma unroll 1
  for (int k = 623; k; k--) {
    uint32_t v = (mt[mt_at(e, i)] ^ ((prev ^ (prev >> 30)) * 1566083941u)) - (uint32_t)i;
    mt[mt_at(e, i)] = v; prev = v;
    i++;
    if (i >= 624) { mt[mt_at(e, 0)] = prev; i = 1; }
  }
  mt[mt_at(e, 0)] = 0x80000000u;
  d.mti[e] = 0;  // numpy's mti = 624 ("regenerate everything") == incremental index 0
  d.words[e] = 0;
  d.spec[e] &= kSpecEpochMask;  // a new stream: the candidates are moot
}

// level edits that do not draw from the env RNG: the candidates' respawn draws depended on the old level
__device__ __forceinline__ void spec_invalidate(const Dev &d, int e) {
  const uint32_t sp = d.spec[e];
  if (sp & kSpecValidMask) d.spec[e] = sp & kSpecEpochMask;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_wall(const Rows &R, int x, int y) { return (R.get(y) >> x) & 1u; }

// grid.get(x, y) is None: no wall, no goal, no agent object (agent object sits at the agent's cell)
__device__ __forceinline__ bool is_empty(const Rows &R, const Env &e, int x, int y) {
  return !is_wall(R, x, y) && !(x == e.gx && y == e.gy) && !(e.has_agent && x == e.ax && y == e.ay);
}

// AdversarialEnv._gen_grid (adversarial.py:166-172): empty grid + wall_rect border.
static __device__ inline void gen_grid(const Rows &R, int W) {
  const uint32_t full = (W >= 32) ? 0xffffffffu : ((1u << W) - 1u);
  const uint32_t mid = 1u | (1u << (W - 1));
  R.set(0, full);
  for (int r = 1; r < W - 1; r++) R.set(r, mid);
  R.set(W - 1, full);
}

// place_obj over the whole grid (multigrid.py:565-632): x=_rand_int(0,W), y=_rand_int(0,H); reject
// non-empty cells; raise after max_tries (max_tries < 0: unbounded).
template <typename RNG>
__device__ __forceinline__ bool place_random(const Rows &R, const Env &e, RNG &rng, int W, int max_tries, int &ox, int &oy) {
  int tries = 0;
  for (;;) {
    if (max_tries >= 0 && tries > max_tries) return false;
    if (!rng.ok()) return false;  // look-ahead window exhausted (speculation only)
    tries++;
    const int x = rng.randint(0, W), y = rng.randint(0, W);
    if (!is_empty(R, e, x, y)) continue;
    ox = x; oy = y;
    return true;
  }
}

// Replay `count` deferred goal respawns (place_one_agent over the whole grid with the agent off the grid).
// Returns the last position as x | y<<8.
template <typename RNG>
__device__ __forceinline__ uint32_t replay_respawns(const Rows &R, int gx, int gy, RNG &rng, int W, int count) {
  Env t{};
  t.gx = gx; t.gy = gy; t.has_agent = 0;
  int px = 0, py = 0;
  for (int i = 0; i < count; i++) place_random(R, t, rng, W, -1, px, py);
  return (uint32_t)px | ((uint32_t)py << 8);
}
template <typename RNG>
__device__ __forceinline__ void flush_pending(const Rows &R, Env &e, RNG &rng, int W) {
  if (e.pending) { replay_respawns(R, e.gx, e.gy, rng, W, e.pending); e.pending = 0; }
}

// Bit-parallel flood fill with the frontier rows held in registers (ROWS is a compile-time bound >= W, all loops
// fully unrolled so every array index is a constant).  Synchronous update: row y of step d+1 is computed from
// rows y-1, y, y+1 of step d (the old row y-1 is carried in `up`).  Returns the hop count to (gx,gy) or -1.
template <int ROWS>
__device__ __forceinline__ int flood_fill(const Rows &R, int W, int sx, int sy, int gx, int gy, uint32_t interior, int max_d) {
  uint32_t reach[ROWS], fr[ROWS];
#pragma unroll
  for (int y = 0; y < ROWS; y++) {
    fr[y] = (y >= 1 && y < W - 1) ? (~R.get(y) & interior) : 0u;
    reach[y] = (y == sy) ? (1u << sx) : 0u;
  }
  for (int d = 1; d <= max_d; d++) {
    uint32_t up = 0, changed = 0, hit = 0;
#pragma unroll
    for (int y = 1; y < ROWS - 1; y++) {
      const uint32_t r = reach[y];
      const uint32_t v = (r | (r << 1) | (r >> 1) | up | reach[y + 1]) & fr[y];
      up = r;
      reach[y] = v;
      changed |= v ^ r;
      hit |= (y == gy) ? ((v >> gx) & 1u) : 0u;
    }
    if (hit) return d;
    if (!changed) return -1;
  }
  return -1;
}

// reset_metrics + compute_metrics (adversarial.py:184-192,407-447): interior wall count, Manhattan
// distance, and reachability / hop count by a bit-parallel flood fill over the interior rows.
static __device__ __noinline__ int4 compute_metrics(const Rows &R, const Env &e, int W, bool do_reset) {
  int4 m;
  const int unreachable = (W - 2) * (W - 2) + 1;
  const uint32_t interior = ((W >= 32) ? 0xffffffffu : ((1u << W) - 1u)) & ~1u & ~(1u << (W - 1));
  int n = 0;
  for (int y = 1; y < W - 1; y++) n += __popc(R.get(y) & interior);
  m.x = n; m.y = -1; m.z = -1; m.w = unreachable;
  (void)do_reset;
  if (e.sx == kNone || e.gx == kNone) return m;
  m.y = abs(e.gx - e.sx) + abs(e.gy - e.sy);
  if (e.sx == e.gx && e.sy == e.gy) { m.z = 1; m.w = 0; return m; }
  const int d = (W <= 16) ? flood_fill<16>(R, W, e.sx, e.sy, e.gx, e.gy, interior, unreachable)
                          : flood_fill<32>(R, W, e.sx, e.sy, e.gx, e.gy, interior, unreachable);
  if (d >= 0) { m.z = 1; m.w = d; } else { m.z = 0; }
  return m;
}

// reset_agent (adversarial.py:238-269) + TimeLimit.reset_agent (time_limit.py:46-48).
__device__ __forceinline__ bool reset_agent(Env &e) {
  e.has_agent = 0; e.adir = e.sdir; e.done_flag = 0;  // reset_agent_status (adversarial.py:231-236)
  if (e.sx == kNone) return false;                     // ValueError at adversarial.py:248-249
  e.has_agent = 1; e.ax = e.sx; e.ay = e.sy;
  e.step_count = 0; e.elapsed = 0;
  return true;
}

// AdversarialEnv.reset (adversarial.py:194-229).
__device__ __forceinline__ void reset_adversary(const Rows &R, Env &e, uint32_t &adv, int4 &met, Rng &rng, const Cfg &c) {
  flush_pending(R, e, rng, c.W);
  e.step_count = 0;
  uint32_t adv_max = (adv >> 12) & 0xfff, sampled = (adv >> 24) & 1;
  if (c.resample) sampled = 0;
  adv = 0u | (adv_max << 12) | (sampled << 24);
  e.sdir = rng.randint(0, 4);
  e.has_agent = 0; e.adir = e.sdir; e.done_flag = 0;
  e.sx = e.sy = kNone; e.gx = e.gy = kNone;
  met = make_int4(0, -1, -1, (c.W - 2) * (c.W - 2) + 1);
  gen_grid(R, c.W);
}

// step_adversary (adversarial.py:452-539), goal_noise == 0.  Returns done.
__device__ __forceinline__ bool step_adversary(const Rows &R, Env &e, uint32_t &adv, int4 &met, Rng &rng, const Cfg &c, int loc,
                                      uint32_t &err) {
  const int W = c.W, I = W - 2, A = I * I;
  if (loc < 0 || loc >= A) { err |= kErrBadLoc; return false; }
  flush_pending(R, e, rng, W);
  {  // the level is edited below, possibly without a draw: DR candidates (their respawn draws) go stale
    const uint32_t sp = rng.spec_p[rng.e];
    if (sp & kSpecValidMask) rng.spec_p[rng.e] = sp & kSpecEpochMask;
  }
  int adv_step = adv & 0xfff, adv_max = (adv >> 12) & 0xfff, sampled = (adv >> 24) & 1;
  if (c.resample && !sampled) {
    adv_max = (int)(((double)loc / (double)A) * (double)c.n_clutter) + 2;
    sampled = 1;
  }
  if (adv_step < adv_max) {
    const int x = loc % I + 1, y = loc / I + 1;
    const bool goal_step = c.goal_last ? (adv_step == adv_max - 2) : (adv_step == 0);
    const bool agent_step = c.goal_last ? (adv_step == adv_max - 1) : (adv_step == 1);
    if (goal_step) {  // remove_wall + put_obj(Goal)
      R.set(y, R.get(y) & ~(1u << x));
      e.gx = x; e.gy = y;
    } else if (agent_step) {
      R.set(y, R.get(y) & ~(1u << x));  // remove_wall
      if (x == e.gx && y == e.gy) {     // goal already here -> place_one_agent(0, rand_dir=False)
        int px = 0, py = 0;
        e.has_agent = 0;
        place_random(R, e, rng, W, -1, px, py);
        e.sx = px; e.sy = py;
      } else { e.sx = x; e.sy = y; }
      e.has_agent = 1; e.ax = e.sx; e.ay = e.sy;
    } else if (is_empty(R, e, x, y)) {
      R.set(y, R.get(y) | (1u << x));
    }
  }
  adv_step++;
  adv = (uint32_t)adv_step | ((uint32_t)adv_max << 12) | ((uint32_t)sampled << 24);
  if (adv_step >= c.n_clutter + 2) { met = compute_metrics(R, e, W, true); return true; }
  return false;
}

// reset_random (adversarial.py:541-581).  n_walls < 0 -> int(n_clutter/2).
__device__ __forceinline__ void reset_random(const Rows &R, Env &e, uint32_t &adv, int4 &met, Rng &rng, const Dev &d, int env,
                                    int n_walls, uint32_t &err) {
  const Cfg &c = d.c;
  const int W = c.W;
  flush_pending(R, e, rng, W);
  if (c.fixed_env) {  // self.seed(self.seed_value) (adversarial.py:542-543)
    mt_seed(d, env, d.limbs[env], d.limbs[(size_t)d.N + env], (int)d.limbs[2 * (size_t)d.N + env]);
    rng.reset();
  }
  e.step_count = 0;
  uint32_t adv_max = (adv >> 12) & 0xfff, sampled = (adv >> 24) & 1;
  e.has_agent = 0; e.adir = e.sdir; e.done_flag = 0;
  e.sx = e.sy = kNone; e.gx = e.gy = kNone;
  gen_grid(R, W);
  int x = 0, y = 0;
  if (!place_random(R, e, rng, W, 100, x, y)) err |= kErrRetries;
  e.gx = x; e.gy = y;
  e.sdir = rng.randint(0, 4);
  place_random(R, e, rng, W, -1, x, y);
  e.sx = x; e.sy = y; e.has_agent = 1; e.ax = x; e.ay = y;
  if (n_walls < 0) n_walls = c.n_clutter / 2;
  else { adv_max = (uint32_t)n_walls + 2; sampled = 1; }  // _resample_n_clutter (adversarial.py:151-156)
  for (int i = 0; i < n_walls; i++) {
    if (!place_random(R, e, rng, W, 100, x, y)) { err |= kErrRetries; break; }
    R.set(y, R.get(y) | (1u << x));
  }
  adv = 0u | (adv_max << 12) | (sampled << 24);
  met = compute_metrics(R, e, W, true);
  if (!reset_agent(e)) err |= kErrNoStart;
}

// ---------------------------------------------------------------------------------------------
// WARP-COOPERATIVE reset_random: all 32 lanes rebuild ONE env's level.  Inside the step kernel a tile usually has
// zero or one finished env; running the lane-serial reset_random there costs the whole ~25 k-instruction generation
// for 1/32 of the lanes.  Here the 32 lanes share the work of one env instead: lane j generates MT19937 word j of a
// 32-word batch (one memory round trip per 32 draws), lane y owns row y of the grid (gen_grid, wall count and the
// flood fill become a handful of shuffles per iteration), and the (inherently serial) rejection sampling runs
// warp-uniformly on broadcast values.  Draw order and count are exactly reset_random()'s.
struct CoopRng {
  uint32_t *mt, *mti_p, *words_p;
  int N, e, lane;
  uint32_t idx, used;
  int have, pos;
  uint32_t w_out, w_nst;
  __device__ __forceinline__ CoopRng(const Dev &dev, int env, int lane_)
      : mt(dev.mt), mti_p(dev.mti), words_p(dev.words), N(dev.N), e(env), lane(lane_), have(0), pos(0), w_out(0), w_nst(0) {
    idx = mti_p[e]; used = words_p[e];
    const uint32_t sp = dev.spec[e];  // see Rng::load
    if (sp & kSpecValidMask) {
      __syncwarp();
      if (lane == 0) dev.spec[e] = sp & kSpecEpochMask;
      __syncwarp();
    }
  }
  __device__ __forceinline__ void commit() {  // write back the state words of the consumed draws
    if (lane < pos) {
      uint32_t i = idx + lane;
      if (i >= 624) i -= 624;
      mt[mt_at(e, i)] = w_nst;
    }
    idx += pos;
    if (idx >= 624) idx -= 624;
    used += pos;
    pos = 0; have = 0;
  }
  __device__ __forceinline__ void refill() {
    commit();
    __syncwarp();
    uint32_t i = idx + lane, i1 = i + 1, im = i + 397;
    if (i >= 624) i -= 624;
    if (i1 >= 624) i1 -= 624;
    if (im >= 624) im -= 624;
    if (im >= 624) im -= 624;
    const uint32_t a = mt[mt_at(e, i)], b = mt[mt_at(e, i1)], c = mt[mt_at(e, im)];
    uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    y = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    w_nst = y;
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    w_out = y;
    have = 32;
  }
  __device__ __forceinline__ uint32_t next() {
    if (pos >= have) refill();
    const uint32_t v = __shfl_sync(0xffffffffu, w_out, pos);
    pos++;
    return v;
  }
  __device__ __forceinline__ int randint(int lo, int hi) {
    const uint32_t rng = (uint32_t)(hi - lo - 1);
    if (rng == 0) return lo;
    const uint32_t mask = 0xffffffffu >> __clz(rng);
    uint32_t v;
    do { v = next() & mask; } while (v > rng);
    return lo + (int)v;
  }
  __device__ __forceinline__ void store() {
    commit();
    if (lane == 0) { mti_p[e] = idx; words_p[e] = used; }
    __syncwarp();
  }
  __device__ __forceinline__ bool ok() const { return true; }
};

// Warp-uniform draws from a precomputed look-ahead window of tempered MT outputs in shared memory (speculation).
struct WinRng {
  const uint32_t *win;
  int pos, n;
  bool overflow;
  __device__ __forceinline__ uint32_t next() {
    if (pos >= n) { overflow = true; return 0u; }
    return win[pos++];
  }
  __device__ __forceinline__ int randint(int lo, int hi) {
    const uint32_t rng = (uint32_t)(hi - lo - 1);
    if (rng == 0) return lo;
    const uint32_t mask = 0xffffffffu >> __clz(rng);
    uint32_t v;
    do { v = next() & mask; } while (v > rng && !overflow);
    return lo + (int)(v > rng ? 0u : v);
  }
  __device__ __forceinline__ bool ok() const { return !overflow; }
};

// The level-building core of reset_random (adversarial.py:546-581), all 32 lanes on warp-uniform values; `R` = W rows in
// shared memory that this warp owns.  Draw order and count are exactly reset_random()'s.
struct LevelOut {
  int gx, gy, sx, sy, sdir;
  int4 met;
  uint32_t err;
};
template <typename RNG>
__device__ __forceinline__ LevelOut coop_level_core(const Rows &R, RNG &rng, int W, int n_walls, int lane) {
  LevelOut o;
  o.err = 0;
  int x = 0, y = 0;
  __syncwarp();
  if (lane < W) {  // gen_grid: lane y writes row y
    const uint32_t full = (W >= 32) ? 0xffffffffu : ((1u << W) - 1u);
    R.set(lane, (lane == 0 || lane == W - 1) ? full : (1u | (1u << (W - 1))));
  }
  __syncwarp();
  if (!coop_place_random(R, kNone, kNone, false, 0, 0, rng, W, 100, x, y)) o.err |= kErrRetries;
  o.gx = x; o.gy = y;
  o.sdir = rng.randint(0, 4);
  coop_place_random(R, o.gx, o.gy, false, 0, 0, rng, W, -1, x, y);
  o.sx = x; o.sy = y;
  for (int i = 0; i < n_walls; i++) {
    if (!coop_place_random(R, o.gx, o.gy, true, o.sx, o.sy, rng, W, 100, x, y)) { o.err |= kErrRetries; break; }
    __syncwarp();
    if (lane == 0) R.set(y, R.get(y) | (1u << x));
    __syncwarp();
  }
  // compute_metrics: lane y owns row y
  const int unreachable = (W - 2) * (W - 2) + 1;
  const uint32_t interior = ((W >= 32) ? 0xffffffffu : ((1u << W) - 1u)) & ~1u & ~(1u << (W - 1));
  const bool inner = lane >= 1 && lane < W - 1;
  const uint32_t row = (lane < W) ? R.get(lane) : 0u;
  int n = inner ? __popc(row & interior) : 0;
  for (int s = 16; s > 0; s >>= 1) n += __shfl_xor_sync(0xffffffffu, n, s);
  o.met = make_int4(n, abs(o.gx - o.sx) + abs(o.gy - o.sy), 0, unreachable);
  const uint32_t fr = inner ? (~row & interior) : 0u;
  uint32_t reach = (lane == o.sy) ? (1u << o.sx) : 0u;
  for (int dd = 1; dd <= unreachable; dd++) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, reach, 1), dn = __shfl_down_sync(0xffffffffu, reach, 1);
    const uint32_t v = (reach | (reach << 1) | (reach >> 1) | (lane > 0 ? up : 0u) | (lane < 31 ? dn : 0u)) & fr;
    const bool changed = __any_sync(0xffffffffu, v != reach);
    reach = v;
    const uint32_t grow = __shfl_sync(0xffffffffu, reach, o.gy);
    if ((grow >> o.gx) & 1u) { o.met.z = 1; o.met.w = dd; break; }
    if (!changed) break;
  }
  return o;
}

// place_obj over the whole grid, warp-uniform (all lanes see the same draws and the same shared-memory rows)
template <typename RNG>
__device__ __forceinline__ bool coop_place_random(const Rows &R, int gx, int gy, bool has_agent, int ax, int ay, RNG &rng, int W,
                                                  int max_tries, int &ox, int &oy) {
  int tries = 0;
  for (;;) {
    if (max_tries >= 0 && tries > max_tries) return false;
    if (!rng.ok()) return false;  // look-ahead window exhausted (speculation only)
    tries++;
    const int x = rng.randint(0, W), y = rng.randint(0, W);
    const bool empty = !((R.get(y) >> x) & 1u) && !(x == gx && y == gy) && !(has_agent && x == ax && y == ay);
    if (!empty) continue;
    ox = x; oy = y;
    return true;
  }
}

// All 32 lanes call this with warp-uniform arguments; `R` is env `env`'s row column in shared memory.  Returns the
// new hot record (uniform) and writes adv / metrics / error flags of the env from lane 0.
static __device__ __noinline__ uint4 coop_reset_random(Dev d, uint32_t *col, int stride, uint4 hot, int env, int n_walls, int lane) {
  const Cfg &c = d.c;
  const int W = c.W;
  const Rows R{col, stride};
  Env e = unpack(hot);
  if (c.fixed_env) {  // self.seed(self.seed_value) (adversarial.py:542-543); deferred respawns of the old stream are moot
    if (lane == 0) mt_seed(d, env, d.limbs[env], d.limbs[(size_t)d.N + env], (int)d.limbs[2 * (size_t)d.N + env]);
    __syncwarp();
    e.pending = 0;
  }
  CoopRng rng(d, env, lane);
  int x = 0, y = 0;
  for (int i = 0; i < e.pending; i++) coop_place_random(R, e.gx, e.gy, false, 0, 0, rng, W, -1, x, y);  // flush_pending
  e.pending = 0;
  e.step_count = 0;
  uint32_t adv = d.adv[env];
  uint32_t adv_max = (adv >> 12) & 0xfff, sampled = (adv >> 24) & 1;
  if (n_walls < 0) n_walls = c.n_clutter / 2;
  else { adv_max = (uint32_t)n_walls + 2; sampled = 1; }
  const LevelOut o = coop_level_core(R, rng, W, n_walls, lane);
  rng.store();
  adv = 0u | (adv_max << 12) | (sampled << 24);
  e.gx = o.gx; e.gy = o.gy; e.sx = o.sx; e.sy = o.sy; e.sdir = o.sdir;
  uint32_t err = o.err;
  if (!reset_agent(e)) err |= kErrNoStart;
  if (lane == 0) {
    d.adv[env] = adv; d.metrics[env] = o.met;
    if (err) d.err[env] |= err;
  }
  return pack(e);
}

// ---------------------------------------------------------------------------------------------
// SPECULATION of the DR auto-reset.  In step_env(reset_random=True) the env RNG is consumed by exactly two things: the
// goal respawn (multigrid.py:821-838: draws that depend only on the current level) and reset_random itself.  So when a
// level starts, its successor is already determined up to ONE bit -- did the episode end at the goal (respawn draws
// first) or not -- and both candidates can be built ahead of time, off the step's critical path.  A regeneration job
// (one warp per env) settles the MT state, computes a look-ahead window of tempered words (all derivable from the
// present state: < 227 words), builds both candidate levels from it in shared memory and stores them as records; the
// step kernel then resets a finished env by COPYING the right record (no RNG, no flood fill) and queues the env for its
// next regeneration, which runs in the tail of the same launch (k_step_env's regeneration phase).
// ---- warp-PARALLEL level construction from the look-ahead window ------------------------------------------------
// reset_random is a chain of rejection-sampled placements; run literally it is ~5 k dependent instructions (20+ us on
// one warp, however the work is split).  From a window of already tempered words it parallelises:
//   * acceptance of a word by randint(0, W) (masked rejection) is a per-word predicate -> 7 ballots;
//   * a placement TRY starting at word i consumes up to the second accepted word: nxt(i), cell(i) are per-position
//     functions -> two byte tables built by all lanes; the k-th try after position p is k applications of nxt:
//     pointer jumping (tables for 1, 2, 4 ... 64 steps);
//   * the goal / start-direction / agent placements are 1-2 tries each (a short uniform walk); the wall tries are
//     enumerated 32 at a time, one per lane: a try places a wall iff its cell is interior, not the goal, not the agent
//     and not the cell of an EARLIER try (__match_any_sync + the rows placed by earlier rounds); the placement stops
//     at the try holding the n_walls-th success (n-th set bit of the success ballot).
// Anything irregular (window exhausted, more than kSpecTries wall tries) just marks the candidate invalid: the step
// kernel then rebuilds that env the slow way.
constexpr int kSpecTries = 96, kSpecLevels = 7, kSpecOver = 255;
struct SpecTables {
  uint32_t *win;    // [kSpecWindow] tempered words
  uint32_t *am;     // [8] acceptance masks (bit j of am[c]: word 32c+j accepted), am[7] = 0
  uint16_t *cell;   // [kSpecWindow] x | y << 5 of the try starting here
  uint8_t *jump;    // [kSpecLevels][kSpecWindow + 32] start of the (2^l)-th next try (kSpecOver: beyond the window)
};
__device__ __forceinline__ int spec_first_acc(const uint32_t *am, int i) {
  if (i >= kSpecWindow) return kSpecOver;
  int c = i >> 5;
  uint32_t m = am[c] & (0xffffffffu << (i & 31));
  while (m == 0) {
    if (++c >= kSpecWindow / 32) return kSpecOver;
    m = am[c];
  }
  return c * 32 + __ffs(m) - 1;
}
__device__ __forceinline__ void spec_build_tables(const SpecTables &T, int W, int lane) {
  const uint32_t rng = (uint32_t)(W - 1), mask = 0xffffffffu >> __clz(rng);
#pragma unroll
  for (int c = 0; c < kSpecWindow / 32; c++) {
    const uint32_t m = __ballot_sync(0xffffffffu, (T.win[32 * c + lane] & mask) <= rng);
    if (lane == 0) T.am[c] = m;
  }
  if (lane == 0) T.am[kSpecWindow / 32] = 0;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < kSpecWindow / 32; c++) {
    const int i = 32 * c + lane;
    const int px = spec_first_acc(T.am, i);
    const int py = (px == kSpecOver) ? kSpecOver : spec_first_acc(T.am, px + 1);
    const bool ok = py != kSpecOver;
    T.jump[i] = ok ? (uint8_t)(py + 1) : (uint8_t)kSpecOver;
    T.cell[i] = ok ? (uint16_t)((T.win[px] & mask) | ((T.win[py] & mask) << 5)) : (uint16_t)0;
  }
  if (lane < 32) T.jump[kSpecWindow + lane] = (uint8_t)kSpecOver;  // positions >= window: absorbing
  __syncwarp();
  constexpr int S = kSpecWindow + 32;
  for (int l = 1; l < kSpecLevels; l++) {
    const uint8_t *src = T.jump + (l - 1) * S;
    uint8_t *dst = T.jump + l * S;
#pragma unroll
    for (int c = 0; c < S / 32; c++) {
      const int i = 32 * c + lane;
      const int m = src[i];
      dst[i] = (m >= kSpecWindow) ? (uint8_t)kSpecOver : src[m];
    }
    __syncwarp();
  }
}

// wall count + Manhattan distance + reachability / hop count with lane y owning row y (compute_metrics, adversarial.py:407-447)
__device__ __forceinline__ int4 coop_metrics(uint32_t row, int gx, int gy, int sx, int sy, int W, int lane) {
  const int unreachable = (W - 2) * (W - 2) + 1;
  const uint32_t interior = ((W >= 32) ? 0xffffffffu : ((1u << W) - 1u)) & ~1u & ~(1u << (W - 1));
  const bool inner = lane >= 1 && lane < W - 1;
  int n = inner ? __popc(row & interior) : 0;
  for (int s = 16; s > 0; s >>= 1) n += __shfl_xor_sync(0xffffffffu, n, s);
  int4 met = make_int4(n, abs(gx - sx) + abs(gy - sy), 0, unreachable);
  const uint32_t fr = inner ? (~row & interior) : 0u;
  uint32_t reach = (lane == sy) ? (1u << sx) : 0u;
  for (int dd = 1; dd <= unreachable; dd++) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, reach, 1), dn = __shfl_down_sync(0xffffffffu, reach, 1);
    const uint32_t v = (reach | (reach << 1) | (reach >> 1) | (lane > 0 ? up : 0u) | (lane < 31 ? dn : 0u)) & fr;
    const bool changed = __any_sync(0xffffffffu, v != reach);
    reach = v;
    const uint32_t grow = __shfl_sync(0xffffffffu, reach, gy);
    if ((grow >> gx) & 1u) { met.z = 1; met.w = dd; break; }
    if (!changed) break;
  }
  return met;
}

// One candidate level.  `respawn`: first replay the goal respawn's draws against the current level `cur` (goal cgx, cgy).
// Returns false when the window / try budget did not suffice.  `lvl` = 32 words of shared memory (the new rows).
__device__ __forceinline__ bool spec_level(const SpecTables &T, const uint32_t *cur, int cgx, int cgy, bool respawn, uint32_t *lvl, int W,
                                           int n_walls, int lane, LevelOut &o, int &consumed) {
  constexpr int S = kSpecWindow + 32;
  int p = 0;
  o.err = 0;
  auto interior = [W](int x, int y) { return x >= 1 && x < W - 1 && y >= 1 && y < W - 1; };
  if (respawn) {
    for (;;) {  // place_obj(max_tries unbounded) on the current level, agent off the grid
      const int q = T.jump[p];
      if (q == kSpecOver) return false;
      const int c = T.cell[p], x = c & 31, y = c >> 5;
      p = q;
      if (!((cur[y] >> x) & 1u) && !(x == cgx && y == cgy)) break;
      if (p >= kSpecWindow) return false;
    }
  }
  if (p >= kSpecWindow) return false;
  for (;;) {  // goal: first interior cell (the fresh grid is empty inside its border)
    const int q = T.jump[p];
    if (q == kSpecOver) return false;
    const int c = T.cell[p];
    p = q;
    o.gx = c & 31; o.gy = c >> 5;
    if (interior(o.gx, o.gy)) break;
    if (p >= kSpecWindow) return false;
  }
  if (p >= kSpecWindow) return false;
  o.sdir = (int)(T.win[p] & 3u);  // randint(0, 4): one word, never rejected
  p++;
  if (p >= kSpecWindow) return false;
  for (;;) {  // agent: interior and not the goal
    const int q = T.jump[p];
    if (q == kSpecOver) return false;
    const int c = T.cell[p];
    p = q;
    o.sx = c & 31; o.sy = c >> 5;
    if (interior(o.sx, o.sy) && !(o.sx == o.gx && o.sy == o.gy)) break;
    if (p >= kSpecWindow) return false;
  }
  // fresh grid: lane y writes row y
  __syncwarp();
  if (lane < W) {
    const uint32_t full = (W >= 32) ? 0xffffffffu : ((1u << W) - 1u);
    lvl[lane] = (lane == 0 || lane == W - 1) ? full : (1u | (1u << (W - 1)));
  }
  __syncwarp();
  consumed = p;
  int placed = 0;
  bool done = n_walls == 0;
  for (int r = 0; r < kSpecTries / 32 && !done; r++) {
    const int k = 32 * r + lane;
    int tp = p;  // start of this lane's try: k applications of nxt
#pragma unroll
    for (int l = 0; l < kSpecLevels; l++)
      if ((k >> l) & 1) tp = T.jump[l * S + tp];
    const int q = (tp < kSpecWindow) ? (int)T.jump[tp] : kSpecOver;
    const bool complete = q != kSpecOver;
    const int c = complete ? (int)T.cell[tp] : 0, x = c & 31, y = c >> 5;
    const bool valid = complete && interior(x, y) && !(x == o.gx && y == o.gy) && !(x == o.sx && y == o.sy);
    const unsigned mm = __match_any_sync(0xffffffffu, valid ? c : (0x10000 + lane));
    const bool success = valid && (lane == __ffs(mm) - 1) && !((lvl[y] >> x) & 1u);
    const unsigned sm = __ballot_sync(0xffffffffu, success);
    const int cnt = __popc(sm), need = n_walls - placed;
    int last = 32;  // lanes <= last place their wall
    if (cnt >= need) {  // the need-th set bit of the success mask
      unsigned t = sm;
      for (int i = 1; i < need; i++) t &= t - 1;
      last = __ffs(t) - 1;
      done = true;
    }
    __syncwarp();
    if (success && lane <= last) atomicOr(&lvl[y], 1u << x);
    if (done) {
      consumed = __shfl_sync(0xffffffffu, q, last);  // the try holding the last wall is complete (it succeeded)
    } else {
      placed += cnt;
      const int comp = __shfl_sync(0xffffffffu, (int)complete, 31);  // the window must cover the whole round
      if (!comp) return false;
    }
    __syncwarp();
  }
  return done;  // metrics are computed lazily (kMetricsDirty): no getter reads them inside a rollout
}

// Record of candidate k of env e built for level epoch `ep` (double-buffered by epoch parity so that a job of the
// previous level can never write the records the current level's commit reads).
__device__ __forceinline__ uint32_t *cand_record(const Dev &d, int e, uint32_t ep, int k) {
  return d.cand + (((size_t)e * 2 + (ep & 1u)) * 2 + k) * cand_words(d.c.W);
}

// One regeneration job = ONE candidate (k = 0: the episode ends without a goal, 1: at the goal) of one env, one warp.
// Jobs queued by launch t run in the tail of launch t+1, next to its tiles, so a job may race with a step warp that
// resets the very env it is reading.  That is harmless by construction: a job only READS env state (the step kernel
// applies a committed record's words to the MT state when it commits), every reset bumps the env's level epoch and
// stores a fresh speculation word, records and validity bits are per epoch parity, and jobs live for one launch --
// whatever a job computed from a torn state lands on the parity that is no longer consulted (spec_valid_bit).
static __device__ __noinline__ void rr_regen_job(Dev d, uint2 jb, int lane, uint32_t *scr /* 1024 words of shared memory */) {
  const Cfg &c = d.c;
  const uint32_t job = jb.x;
  const int e = (int)(job >> 8), k = (int)((job >> 7) & 1u);
  const uint32_t ep = job & 127u;
  const int W = c.W;
  SpecTables T;
  T.win = scr; T.am = scr + kSpecWindow; uint32_t *cur = scr + kSpecWindow + 8, *lvl = cur + 32;
  T.cell = reinterpret_cast<uint16_t *>(lvl + 32);
  T.jump = reinterpret_cast<uint8_t *>(T.cell + kSpecWindow);
  const uint32_t idx = jb.y % 624u;  // the cursor travels with the job: every load of the job is issued in one go
  const uint32_t sp0 = __ldcg(d.spec + e);
  const uint4 hot = __ldcg(d.hot + e);
  const uint32_t cur_row = (k == 1 && lane < W) ? __ldcg(env_rows(d, e).p + lane * 32) : 0xffffffffu;
  // look-ahead window: tempered outputs idx .. idx + kSpecWindow - 1 of the present state (nothing is stored to mt).
  // All 21 loads per lane are issued before the first shared-memory store (which the compiler must order against them).
  uint32_t nst[kSpecState / 32];  // untempered = new state words of positions lane + 32 i
  {
    constexpr int kPer = kSpecWindow / 32;
    uint32_t xs[kPer], bs[kPer], cs[kPer];
#pragma unroll
    for (int i = 0; i < kPer; i++) {
      uint32_t p = idx + lane + 32 * i, p1 = p + 1, pm = p + 397;
      if (p >= 624) p -= 624;
      if (p1 >= 624) p1 -= 624;
      if (pm >= 624) pm -= 624;
      if (pm >= 624) pm -= 624;
      xs[i] = __ldcg(d.mt + mt_at(e, p)); bs[i] = __ldcg(d.mt + mt_at(e, p1)); cs[i] = __ldcg(d.mt + mt_at(e, pm));
    }
#pragma unroll
    for (int i = 0; i < kPer; i++) {
      uint32_t y = (xs[i] & 0x80000000u) | (bs[i] & 0x7fffffffu);
      y = cs[i] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      if (i < kSpecState / 32) nst[i] = y;
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      T.win[lane + 32 * i] = y;
    }
  }
  cur[lane] = cur_row;
  const Env s = unpack(hot);
  __syncwarp();
  if (spec_epoch(sp0) != ep) return;  // the level is gone already
  const long long c1 = d.prof ? clock64() : 0;
  spec_build_tables(T, W, lane);
  const long long c2 = d.prof ? clock64() : 0;
  LevelOut o;
  int consumed = 0;
  const bool ok = spec_level(T, cur, s.gx & 31, s.gy & 31, k == 1, lvl, W, c.n_clutter / 2, lane, o, consumed);
  __syncwarp();
  if (d.prof && lane == 0) {
    atomicAdd(&d.prof[6], (unsigned long long)(c2 - c1)); atomicAdd(&d.prof[7], (unsigned long long)(clock64() - c2));
  }
  if (!ok || consumed > kSpecState) return;
  uint32_t *rec = cand_record(d, e, ep, k);
  if (lane < W) rec[lane] = lvl[lane];
#pragma unroll
  for (int i = 0; i < kSpecState / 32; i++) rec[W + 8 + lane + 32 * i] = nst[i];  // coalesced; the commit copies them into mt
  if (lane == 0) {
    rec[W] = ((uint32_t)o.gx & 31u) | (((uint32_t)o.gy & 31u) << 5) | (1u << 10) | (((uint32_t)o.sx & 31u) << 11) |
             (((uint32_t)o.sy & 31u) << 16) | (1u << 21) | ((uint32_t)o.sdir << 22);
    rec[W + 5] = (uint32_t)consumed;
    rec[W + 6] = o.err;
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    atomicOr(d.spec + e, spec_valid_bit(ep, k));  // result unused: a reduction, no round trip
  }
}

// ---- lane-parallel regeneration for LONG job lists (a synchronized time-limit storm queues two jobs for nearly every
// env) ------------------------------------------------------------------------------------------------------------
// A warp-per-job build costs ~10 us per candidate whatever the list length; 32 candidates built side by side by the 32
// lanes with the plain lane-serial code cost about as much as ONE of them (k_reset_random builds 131 072 levels in 0.23
// ms).  The draws come straight from the present MT state, non-destructively: output j of the look-ahead span is a
// function of present words j, j+1, j+397 for j < 227, so no window has to be stored, and the untempered value IS the new
// state word the record carries.
struct FlyRng {
  const uint32_t *mt;
  uint32_t *rec_state;  // record slots of the new state words (first kSpecState draws)
  int e;
  uint32_t idx;
  int pos;
  bool overflow;
  __device__ __forceinline__ uint32_t next() {
    if (pos >= kSpecWindow) { overflow = true; return 0u; }
    uint32_t p = idx + pos, p1 = p + 1, pm = p + 397;
    if (p >= 624) p -= 624;
    if (p1 >= 624) p1 -= 624;
    if (pm >= 624) pm -= 624;
    if (pm >= 624) pm -= 624;
    const uint32_t x = __ldcg(mt + mt_at(e, p)), b = __ldcg(mt + mt_at(e, p1)), c = __ldcg(mt + mt_at(e, pm));
    uint32_t y = (x & 0x80000000u) | (b & 0x7fffffffu);
    y = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    if (pos < kSpecState) rec_state[pos] = y;
    pos++;
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  __device__ __forceinline__ int randint(int lo, int hi) {
    const uint32_t rng = (uint32_t)(hi - lo - 1);
    if (rng == 0) return lo;
    const uint32_t mask = 0xffffffffu >> __clz(rng);
    uint32_t v;
    do { v = next() & mask; } while (v > rng && !overflow);
    return lo + (int)(v > rng ? 0u : v);
  }
  __device__ __forceinline__ bool ok() const { return !overflow; }
};

// one candidate per LANE; `col` = this lane's column of a [W][32] shared-memory scratch (the level under construction)
static __device__ __noinline__ void rr_regen_job_lane(Dev d, uint2 jb, uint32_t *col) {
  const Cfg &c = d.c;
  const uint32_t job = jb.x;
  const int e = (int)(job >> 8), k = (int)((job >> 7) & 1u);
  const uint32_t ep = job & 127u;
  const int W = c.W;
  if (spec_epoch(__ldcg(d.spec + e)) != ep) return;  // the level is gone already
  const Env s = unpack(__ldcg(d.hot + e));
  uint32_t *rec = cand_record(d, e, ep, k);
  FlyRng rng{d.mt, rec + W + 8, e, jb.y % 624u, 0, false};
  const Rows L{col, 32};
  int x = 0, y = 0;
  uint32_t err = 0;
  if (k == 1) {  // the goal respawn's draws (multigrid.py:821-838) against the current level, agent off the grid
    Env t{};
    t.gx = s.gx; t.gy = s.gy; t.has_agent = 0;
    place_random(env_rows(d, e), t, rng, W, -1, x, y);
  }
  Env n{};
  n.gx = n.gy = n.sx = n.sy = kNone;
  gen_grid(L, W);
  if (!place_random(L, n, rng, W, 100, x, y)) err |= kErrRetries;
  n.gx = x; n.gy = y;
  n.sdir = rng.randint(0, 4);
  place_random(L, n, rng, W, -1, x, y);
  n.sx = x; n.sy = y; n.has_agent = 1; n.ax = x; n.ay = y;
  const int n_walls = c.n_clutter / 2;
  for (int i = 0; i < n_walls; i++) {
    if (!place_random(L, n, rng, W, 100, x, y)) { err |= kErrRetries; break; }
    L.set(y, L.get(y) | (1u << x));
  }
  if (!rng.ok() || rng.pos > kSpecState) return;
  for (int r = 0; r < W; r++) rec[r] = L.get(r);
  rec[W] = ((uint32_t)n.gx & 31u) | (((uint32_t)n.gy & 31u) << 5) | (1u << 10) | (((uint32_t)n.sx & 31u) << 11) |
           (((uint32_t)n.sy & 31u) << 16) | (1u << 21) | ((uint32_t)n.sdir << 22);
  rec[W + 5] = (uint32_t)rng.pos;
  rec[W + 6] = err;
  __threadfence();
  atomicOr(d.spec + e, spec_valid_bit(ep, k));
}

// ---------------------------------------------------------------------------------------------
// Egocentric view (multigrid.py:977-1055 gen_obs_grid/gen_agent_obs, 320-338 slice, 300-318
// rotate_left, 749-782 get_view_exts; gym_minigrid Grid.process_vis / encode).
//
// image[vx][vy] = cell((ax,ay) + f*(V-1-vy) + r*(vx-V/2)), f = DIR_TO_VEC[dir], r = (-f.y, f.x);
// out of bounds = wall.  Rows of the view are built as 5-bit masks (bit vx) from the wall bit-plane.
// Returns wall masks w[vy], visibility masks vis[vy] and the goal's view cell (gvx,gvy) or (-1,-1).
struct View {
  uint32_t w[kV], vis[kV];
  int gvx, gvy;
};

template <bool SEE_THROUGH, typename EXT>
__device__ __forceinline__ View render_view_t(const Rows &R, const Env &e, int W) {
  View v;
  // extended rows: bit (x+PAD) of E = wall at x, everything outside [0,W) is wall.  EXT = uint32_t needs
  // W + 2*PAD <= 32 (W <= 24, the 15x15 mazes); uint64_t covers W <= 32.
  constexpr int PAD = 4;
  const int d = e.adir;
  const bool vertical = d & 1;                  // facing down/up: view rows are world rows
  const int sgn = (d == 0 || d == 1) ? 1 : -1;  // forward sign along its axis
  const EXT ones = ~(EXT)0;
  const EXT border = ((EXT)15) | (ones << (W + PAD));
  EXT E[kV];
#pragma unroll
  for (int k = 0; k < kV; k++) {
    // vertical: world row of view row vy=k is ay + sgn*(4-k); horizontal: world row of view column vx=k is
    // ay + r.y*(k-2) with r = (-f.y, f.x): dir 0 -> r=(0,1), dir 2 -> r=(0,-1)
    const int wy = vertical ? (e.ay + sgn * (kV - 1 - k)) : (e.ay + sgn * (k - kV / 2));
    EXT row = ones;
    if (wy >= 0 && wy < W) row = ((EXT)R.get(wy) << PAD) | border;
    E[k] = row;
  }
#pragma unroll
  for (int vy = 0; vy < kV; vy++) {
    uint32_t m = 0;
    if (vertical) {
      // dir 3 (up): wx = ax + (vx-2); dir 1 (down): r = (-1,0): wx = ax - (vx-2)
      const uint32_t five = (uint32_t)(E[vy] >> (e.ax - 2 + PAD)) & 31u;
      m = (d == 3) ? five : (__brev(five) >> 27);
    } else {
      const int wx = e.ax + sgn * (kV - 1 - vy);
#pragma unroll
      for (int vx = 0; vx < kV; vx++) m |= (uint32_t)((E[vx] >> (wx + PAD)) & 1) << vx;
    }
    v.w[vy] = m;
  }
  // goal in view coordinates: fd = (g-a).f, lt = (g-a).r
  {
    const int dx = e.gx - e.ax, dy = e.gy - e.ay;
    int fd, lt;
    if (d == 0) { fd = dx; lt = dy; } else if (d == 1) { fd = dy; lt = -dx; }
    else if (d == 2) { fd = -dx; lt = -dy; } else { fd = -dy; lt = dx; }
    const int gvy = kV - 1 - fd, gvx = lt + kV / 2;
    const bool in = (e.gx != kNone) && gvx >= 0 && gvx < kV && gvy >= 0 && gvy < kV;
    v.gvx = in ? gvx : -1; v.gvy = in ? gvy : -1;
  }
  if (SEE_THROUGH) {
#pragma unroll
    for (int j = 0; j < kV; j++) v.vis[j] = 31u;
  } else {
    // gym_minigrid Grid.process_vis(agent_pos=(V/2, V-1)), the two sweeps per row as bit closures
    uint32_t m = 1u << (kV / 2);
#pragma unroll
    for (int j = kV - 1; j >= 0; j--) {
      const uint32_t open = ~v.w[j] & 31u;
      uint32_t P = m;
#pragma unroll
      for (int it = 0; it < kV - 1; it++) P |= ((P & open) << 1) & 31u;   // left -> right, i = 0..V-2
      const uint32_t srcL = P & open & 15u;
      uint32_t nxt = srcL | (srcL << 1);
      uint32_t Q = P;
#pragma unroll
      for (int it = 0; it < kV - 1; it++) Q |= (Q & open) >> 1;           // right -> left, i = V-1..1
      const uint32_t srcR = Q & open & 30u;
      nxt |= srcR | (srcR >> 1);
      v.vis[j] = Q;
      m = nxt;
    }
  }
  return v;
}

template <bool SEE_THROUGH>
__device__ __forceinline__ View render_view(const Rows &R, const Env &e, int W) {
  return render_view_t<SEE_THROUGH, uint64_t>(R, e, W);
}

// ---- packed, branch-free variant used by the hot kernels -------------------------------------------------------
// The 5x5 view is ONE 25-bit word Z with bit (vy*5 + vx) = wall at view cell (vx, vy).  All four directions share one
// instruction stream (a warp holds agents facing every way, so a per-direction branch executes every side):
//   * five world rows are fetched in a per-direction order (k-th row = base + step*k) and 5 bits are cut out of each
//     at a per-direction column offset -> X, bit (k*5 + j) = wall at (off + j, base + step*k);
//   * facing right / down the cut is mirrored in both axes: brev(X) >> 7 (a 25-bit reversal);
//   * facing right / left the view rows are world columns: a 5x5 bit-matrix transpose (8 shifted masks).
// dir 3 (up):    cell(vx,vy) = (ax-2+vx,   ay-4+vy)   rows ay-4.. step +1, off ax-2            -> Z = X
// dir 1 (down):  cell(vx,vy) = (ax+2-vx,   ay+4-vy)   rows ay..   step +1, off ax-2, mirrored  -> Z = rev25(X)
// dir 2 (left):  cell(vx,vy) = (ax-4+vy,   ay+2-vx)   rows ay+2.. step -1, off ax-4            -> Z = X^T
// dir 0 (right): cell(vx,vy) = (ax+4-vy,   ay-2+vx)   rows ay+2.. step -1, off ax,   mirrored  -> Z = rev25(X)^T
__device__ __forceinline__ uint32_t transpose5(uint32_t x) {  // bit (r*5+c) -> bit (c*5+r)
  uint32_t y = x & 0x1041041u;
  y |= (x >> 4) & 0x0082082u;
  y |= (x >> 8) & 0x0004104u;
  y |= (x >> 12) & 0x0000208u;
  y |= (x >> 16) & 0x0000010u;
  y |= (x << 4) & 0x0820820u;
  y |= (x << 8) & 0x0410400u;
  y |= (x << 12) & 0x0208000u;
  y |= (x << 16) & 0x0100000u;
  return y;
}

struct PackedView {
  uint32_t wall;  // bit (vy*5+vx): wall, already masked by `vis`
  uint32_t vis;   // bit (vy*5+vx): visible (all ones when see_through_walls)
  int goal;       // vx*5+vy of the goal if it is in view and visible, else -1
};

// (W = number of rows; Wc = number of columns a row word holds, W unless the caller hands in a window of a wider grid)
template <bool SEE_THROUGH, typename EXT>
__device__ __forceinline__ PackedView render_packed(const Rows &R, const Env &e, int W, int Wc = -1) {
  constexpr int PAD = 4;
  if (Wc < 0) Wc = W;
  const int d = e.adir;
  const bool vertical = d & 1, mirrored = d < 2;
  const int base = (d == 3) ? e.ay - 4 : (d == 1) ? e.ay : e.ay + 2;
  const int step = vertical ? 1 : -1;
  const int off = ((d == 0) ? e.ax : (d == 2) ? e.ax - 4 : e.ax - 2) + PAD;
  const EXT ones = ~(EXT)0;
  const EXT border = ((EXT)15) | (ones << (Wc + PAD));
  uint32_t X = 0;
#pragma unroll
  for (int k = 0; k < kV; k++) {
    const int wy = base + step * k;
    EXT row = ones;  // outside the grid = wall
    if (wy >= 0 && wy < W) row = ((EXT)R.get(wy) << PAD) | border;
    X |= ((uint32_t)(row >> off) & 31u) << (5 * k);
  }
  if (mirrored) X = __brev(X) >> 7;
  if (!vertical) X = transpose5(X);
  PackedView v;
  // goal in view coordinates: fd = (g-a).f, lt = (g-a).r
  const int dx = e.gx - e.ax, dy = e.gy - e.ay;
  const int p = vertical ? dy : dx, q = vertical ? dx : dy;  // d=0: fd=dx, lt=dy; 1: dy,-dx; 2: -dx,-dy; 3: -dy,dx
  const int fd = (d == 0 || d == 1) ? p : -p;
  const int lt = (d == 0 || d == 3) ? q : -q;
  const int gvy = kV - 1 - fd, gvx = lt + kV / 2;
  const bool in = (e.gx != kNone) && (unsigned)gvx < (unsigned)kV && (unsigned)gvy < (unsigned)kV;
  if (SEE_THROUGH) {
    v.wall = X; v.vis = 0x1ffffffu;
    v.goal = in ? gvx * kV + gvy : -1;
  } else {
    // gym_minigrid Grid.process_vis(agent_pos=(V/2, V-1)), the two sweeps per row as bit closures
    uint32_t m = 1u << (kV / 2), V = 0;
#pragma unroll
    for (int j = kV - 1; j >= 0; j--) {
      const uint32_t open = ~(X >> (5 * j)) & 31u;
      uint32_t P = m;
#pragma unroll
      for (int it = 0; it < kV - 1; it++) P |= ((P & open) << 1) & 31u;   // left -> right, i = 0..V-2
      const uint32_t srcL = P & open & 15u;
      uint32_t nxt = srcL | (srcL << 1);
      uint32_t Q = P;
#pragma unroll
      for (int it = 0; it < kV - 1; it++) Q |= (Q & open) >> 1;           // right -> left, i = V-1..1
      const uint32_t srcR = Q & open & 30u;
      nxt |= srcR | (srcR >> 1);
      V |= Q << (5 * j);
      m = nxt;
    }
    v.vis = V; v.wall = X & V;
    v.goal = (in && ((V >> (gvy * kV + gvx)) & 1u)) ? gvx * kV + gvy : -1;
  }
  return v;
}

// preprocessed observation [3][5][5] (c, vx, vy) from the packed view: type = unseen 0 | wall 0.2 | goal 0.8 | empty
// 0.1, colour = wall 0.5 | goal 0.1 | 0, state = 0.  One bit test per cell; the goal (at most one cell, never a wall)
// is patched in afterwards; the agent's own cell (V/2, V-1) never holds a wall or the goal and renders as empty.
// ZERO_STATE = false leaves the all-zero state plane alone (the caller zeroed it once and nothing else writes it).
template <bool SEE_THROUGH, bool ZERO_STATE>
__device__ __forceinline__ void emit_packed_f32(const PackedView &v, float *o) {
#pragma unroll
  for (int vx = 0; vx < kV; vx++)
#pragma unroll
    for (int vy = 0; vy < kV; vy++) {
      const uint32_t bit = 1u << (vy * kV + vx);
      const bool bw = v.wall & bit;
      float t = bw ? 0.2f : 0.1f;
      if (!SEE_THROUGH) t = (v.vis & bit) ? t : 0.0f;
      o[vx * kV + vy] = t;
      o[kV * kV + vx * kV + vy] = bw ? 0.5f : 0.0f;
      if (ZERO_STATE) o[2 * kV * kV + vx * kV + vy] = 0.0f;
    }
  if (v.goal >= 0) { o[v.goal] = 0.8f; o[kV * kV + v.goal] = 0.1f; }
}

// cell code of view cell (vx,vy): 0 unseen, 1 empty, 2 wall, 3 goal
__device__ __forceinline__ int view_code(const View &v, int vx, int vy) {
  if (!((v.vis[vy] >> vx) & 1u)) return 0;
  if (vx == kV / 2 && vy == kV - 1) return 1;  // the agent's own cell is blanked (multigrid.py:1009-1013)
  if ((v.w[vy] >> vx) & 1u) return 2;
  if (vx == v.gvx && vy == v.gvy) return 3;
  return 1;
}

// float32(uint8 / 10.0) for the codes that occur (obs_wrappers.py:104-110): type 0,1,2,8 ; colour 0,0,5,1
__device__ __forceinline__ float type_f(int code) { return code == 0 ? 0.0f : code == 1 ? 0.1f : code == 2 ? 0.2f : 0.8f; }
__device__ __forceinline__ float color_f(int code) { return code == 2 ? 0.5f : code == 3 ? 0.1f : 0.0f; }

// write the preprocessed observation [3][5][5] (c, vx, vy) to `o` (shared or global): per cell
// type = unseen 0 | wall 0.2 | goal 0.8 | empty 0.1, colour = wall 0.5 | goal 0.1 | 0, state = 0.  The agent's own
// cell (V/2, V-1) never holds a wall or the goal, so it renders as empty without a special case.
template <bool SEE_THROUGH>
__device__ __forceinline__ void emit_obs_f32_fast(const View &v, float *o) {
#pragma unroll
  for (int vy = 0; vy < kV; vy++) {
    const uint32_t vis = SEE_THROUGH ? 31u : v.vis[vy];
    const uint32_t w = v.w[vy] & vis;
    const uint32_t g = (vy == v.gvy) ? ((1u << v.gvx) & vis & ~w) : 0u;
#pragma unroll
    for (int vx = 0; vx < kV; vx++) {
      const bool bw = (w >> vx) & 1u, bg = (g >> vx) & 1u, bv = (vis >> vx) & 1u;
      float t = bw ? 0.2f : (bg ? 0.8f : 0.1f);
      if (!SEE_THROUGH) t = bv ? t : 0.0f;
      o[vx * kV + vy] = t;
      o[kV * kV + vx * kV + vy] = bw ? 0.5f : (bg ? 0.1f : 0.0f);
      o[2 * kV * kV + vx * kV + vy] = 0.0f;
    }
  }
}
__device__ __forceinline__ void emit_obs_f32(const View &v, float *o) {
#pragma unroll
  for (int vx = 0; vx < kV; vx++)
#pragma unroll
    for (int vy = 0; vy < kV; vy++) {
      const int code = view_code(v, vx, vy);
      o[vx * kV + vy] = type_f(code);
      o[kV * kV + vx * kV + vy] = color_f(code);
      o[2 * kV * kV + vx * kV + vy] = 0.0f;
    }
}
// raw gym_minigrid encoding [vx][vy][3]
__device__ __forceinline__ void emit_obs_u8(const View &v, uint8_t *o) {
#pragma unroll
  for (int vx = 0; vx < kV; vx++)
#pragma unroll
    for (int vy = 0; vy < kV; vy++) {
      const int code = view_code(v, vx, vy);
      uint8_t *p = o + (vx * kV + vy) * 3;
      p[0] = code == 0 ? 0 : code == 1 ? 1 : code == 2 ? 2 : 8;
      p[1] = code == 2 ? 5 : code == 3 ? 1 : 0;
      p[2] = 0;
    }
}

}  // namespace mgplr
