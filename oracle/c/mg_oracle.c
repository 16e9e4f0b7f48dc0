/*
 * mg_oracle.c -- CPU restatement of the reference's MultiGrid adversarial-maze hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle (and the bench.py CPU baseline):
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Nothing under dcd_isaac_b200/ links, imports or calls it; the product path is the
 * CUDA library (include/mgplr.h) and fails loudly when that is missing.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so the
 * oracle is pinned against outputs of the reference's OWN Python modules run in-process
 * (oracle/ref_harness.py + oracle/gen_golden.py -> tests/golden/ fixtures) and, when
 * /root/reference is present, live in tests/test_oracle_vs_reference.py.
 *
 * Representation: one byte per cell (0 empty, 1 wall, 2 goal), agent kept as coordinates --
 * deliberately NOT the bit-plane layout of the CUDA path, so the two do not share bugs.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Third-party pieces (gym-minigrid 1.0.1, gym 0.15.7, numpy legacy RandomState) are restated
 * from their published algorithms; see oracle/shim/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MGO_MAXW 32
#define MGO_V 5
#define CELL_EMPTY 0
#define CELL_WALL 1
#define CELL_GOAL 2

/* ---------------------------------------------------------------- MT19937 (numpy legacy RandomState) */
typedef struct {
  uint32_t mt[624];
  int mti;
} mgo_mt;

static void mt_init_genrand(mgo_mt *r, uint32_t s) {
  r->mt[0] = s;
  for (int i = 1; i < 624; i++) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
  r->mti = 624;
}

/* numpy RandomState.seed(array) -> init_by_array (Matsumoto & Nishimura reference code). */
static void mt_init_by_array(mgo_mt *r, const uint32_t *key, int key_length) {
  mt_init_genrand(r, 19650218u);
  int i = 1, j = 0;
  int k = (624 > key_length ? 624 : key_length);
  for (; k; k--) {
    r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    i++; j++;
    if (i >= 624) { r->mt[0] = r->mt[623]; i = 1; }
    if (j >= key_length) j = 0;
  }
  for (k = 623; k; k--) {
    r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    i++;
    if (i >= 624) { r->mt[0] = r->mt[623]; i = 1; }
  }
  r->mt[0] = 0x80000000u;
  r->mti = 624;
}

static uint32_t mt_next(mgo_mt *r) {
  if (r->mti >= 624) {
    int kk;
    uint32_t y;
    for (kk = 0; kk < 624 - 397; kk++) {
      y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
      r->mt[kk] = r->mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < 623; kk++) {
      y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
      r->mt[kk] = r->mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (r->mt[623] & 0x80000000u) | (r->mt[0] & 0x7fffffffu);
    r->mt[623] = r->mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    r->mti = 0;
  }
  uint32_t y = r->mt[r->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

/* numpy legacy RandomState.randint(lo, hi): masked rejection on 32-bit words; rng==0 draws nothing.
 * This is gym_minigrid MiniGridEnv._rand_int (called at multigrid.py:603-606, adversarial.py:205,567). */
static int mt_randint(mgo_mt *r, int lo, int hi, int *words) {
  uint32_t rng = (uint32_t)(hi - lo - 1);
  if (rng == 0) return lo;
  uint32_t mask = rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  do { v = mt_next(r) & mask; if (words) (*words)++; } while (v > rng);
  return lo + (int)v;
}

/* ---------------------------------------------------------------- env */
typedef struct {
  int W;                  /* grid side (adversarial.py:70 size) */
  int max_steps;          /* env max_steps (adversarial.py:72) */
  int max_episode_steps;  /* TimeLimit (registration max_episode_steps, time_limit.py:20) */
  int see_through;        /* see_through_walls (adversarial.py:76) */
  int n_clutter;          /* adversarial.py:68 */
  int resample_n_clutter; /* adversarial.py:69 */
  int choose_goal_last;   /* adversarial.py:75 */
  int fixed_environment;  /* adversarial.py:79 */
  int n_editor_actions;   /* len(EDITOR_ACTION_SPACES[...]) adversarial.py:40-56: 2 "-.", 3 "-.g", 4 "-.ag" */
} mgo_cfg;

typedef struct {
  mgo_cfg c;
  uint8_t cells[MGO_MAXW * MGO_MAXW]; /* index y*W+x, like gym_minigrid Grid.grid */
  int has_agent;                      /* agent_pos[0] is not None */
  int ax, ay, adir;
  int gx, gy;       /* goal_pos, -1 = None */
  int sx, sy, sdir; /* agent_start_pos (-1 = None), agent_start_dir */
  int step_count, elapsed, done_flag;
  int adv_step, adv_max, n_clutter_sampled;
  int n_clutter_placed, dist, passable, spl;
  float ep_ret; int ep_len; /* VecMonitor eprets / eplens (vec_monitor.py:39-41,62-63) */
  uint32_t limbs[2]; int n_limbs;
  mgo_mt rng;
  int rng_words; /* number of MT words consumed since seeding (test introspection) */
  int error;     /* 1 = RetriesExceededError analogue (multigrid.py:597-599) */
} mgo_env;

int mgo_sizeof_env(void) { return (int)sizeof(mgo_env); }

static int env_randint(mgo_env *e, int lo, int hi) { return mt_randint(&e->rng, lo, hi, &e->rng_words); }

/* MultiGridEnv.seed -> gym seeding.np_random (multigrid.py:465-468); limbs = _int_list_from_bigint(hash_seed(seed)). */
void mgo_seed(mgo_env *e, const uint32_t *limbs, int n) {
  e->n_limbs = n; e->limbs[0] = limbs[0]; e->limbs[1] = n > 1 ? limbs[1] : 0;
  mt_init_by_array(&e->rng, limbs, n);
  e->rng_words = 0;
}

static void reset_metrics(mgo_env *e) { /* adversarial.py:184-188 */
  int W = e->c.W;
  e->dist = -1; e->n_clutter_placed = 0; e->passable = -1; e->spl = (W - 2) * (W - 2) + 1;
}

static void gen_grid(mgo_env *e) { /* adversarial.py:166-172: empty grid + wall_rect */
  int W = e->c.W;
  memset(e->cells, CELL_EMPTY, sizeof(e->cells));
  for (int i = 0; i < W; i++) {
    e->cells[0 * W + i] = CELL_WALL; e->cells[(W - 1) * W + i] = CELL_WALL;
    e->cells[i * W + 0] = CELL_WALL; e->cells[i * W + (W - 1)] = CELL_WALL;
  }
}

static void reset_agent_status(mgo_env *e) { /* adversarial.py:231-236 */
  e->has_agent = 0; e->adir = e->sdir; e->done_flag = 0;
}

/* _count_walls + compute_shortest_path (adversarial.py:190-192,407-447): interior wall count,
 * Manhattan distance, BFS reachability / hop count on the interior 4-grid minus walls. */
static void compute_metrics(mgo_env *e) {
  int W = e->c.W, n = 0;
  for (int y = 1; y < W - 1; y++)
    for (int x = 1; x < W - 1; x++) n += (e->cells[y * W + x] == CELL_WALL);
  e->n_clutter_placed = n;
  if (e->sx < 0 || e->gx < 0) return;
  e->dist = abs(e->gx - e->sx) + abs(e->gy - e->sy);
  int distv[MGO_MAXW * MGO_MAXW];
  int queue[MGO_MAXW * MGO_MAXW];
  for (int i = 0; i < W * W; i++) distv[i] = -1;
  int qh = 0, qt = 0;
  distv[e->sy * W + e->sx] = 0; queue[qt++] = e->sy * W + e->sx;
  const int dx[4] = {1, -1, 0, 0}, dy[4] = {0, 0, 1, -1};
  while (qh < qt) {
    int cur = queue[qh++], cx = cur % W, cy = cur / W;
    for (int k = 0; k < 4; k++) {
      int nx = cx + dx[k], ny = cy + dy[k];
      if (nx < 1 || ny < 1 || nx > W - 2 || ny > W - 2) continue;
      if (e->cells[ny * W + nx] == CELL_WALL) continue;
      if (distv[ny * W + nx] >= 0) continue;
      distv[ny * W + nx] = distv[cur] + 1; queue[qt++] = ny * W + nx;
    }
  }
  int d = distv[e->gy * W + e->gx];
  if (d >= 0) { e->passable = 1; e->spl = d; }
  else { e->passable = 0; e->spl = (W - 2) * (W - 2) + 1; }
}

void mgo_init(mgo_env *e, const mgo_cfg *c) {
  memset(e, 0, sizeof(*e));
  e->c = *c; e->gx = e->gy = e->sx = e->sy = -1; e->adv_max = c->n_clutter + 2;
  gen_grid(e); reset_metrics(e);
}

/* AdversarialEnv.reset (adversarial.py:194-229): one env-RNG word for agent_start_dir. */
void mgo_reset(mgo_env *e) {
  /* no re-seed here: AdversarialEnv.reset overrides MultiGridEnv.reset (multigrid.py:471-472 is not
   * reached); only reset_random re-seeds a fixed_environment (adversarial.py:542-543). */
  e->step_count = 0; e->adv_step = 0;
  if (e->c.resample_n_clutter) e->n_clutter_sampled = 0;
  e->sdir = env_randint(e, 0, 4);
  reset_agent_status(e);
  e->sx = e->sy = -1; e->gx = e->gy = -1;
  reset_metrics(e);
  gen_grid(e);
}

/* place_obj over the whole grid (multigrid.py:565-632): x=_rand_int(0,W), y=_rand_int(0,H); reject
 * non-empty cells and the agent's cell; `num_tries > max_tries` raises.  Returns 0 on success. */
static int place_random(mgo_env *e, int max_tries, int *ox, int *oy) {
  int W = e->c.W, tries = 0;
  for (;;) {
    if (max_tries >= 0 && tries > max_tries) { e->error = 1; return 1; }
    tries++;
    int x = env_randint(e, 0, W), y = env_randint(e, 0, W);
    if (e->cells[y * W + x] != CELL_EMPTY) continue;
    if (e->has_agent && x == e->ax && y == e->ay) continue;
    *ox = x; *oy = y; return 0;
  }
}

/* reset_agent (adversarial.py:238-269) + TimeLimit.reset_agent (time_limit.py:46-48).  Returns 1 for the
 * ValueError('Trying to place agent at empty start position.'). */
int mgo_reset_agent(mgo_env *e) {
  reset_agent_status(e);
  if (e->sx < 0) return 1;
  e->has_agent = 1; e->ax = e->sx; e->ay = e->sy; /* place_agent_at_pos(rand_dir=False) */
  e->step_count = 0; e->elapsed = 0;
  return 0;
}

/* step_adversary (adversarial.py:452-539).  Returns done; *err=1 for loc >= adversary_action_dim. */
int mgo_step_adversary(mgo_env *e, int loc, int *err) {
  int W = e->c.W, A = (W - 2) * (W - 2);
  if (err) *err = 0;
  if (loc >= A) { if (err) *err = 1; return 0; }
  if (e->c.resample_n_clutter && !e->n_clutter_sampled) {
    int nc = (int)(((double)loc / (double)A) * (double)e->c.n_clutter);
    e->adv_max = nc + 2; e->n_clutter_sampled = 1;
  }
  if (e->adv_step < e->adv_max) {
    int x = loc % (W - 2) + 1, y = loc / (W - 2) + 1;
    int goal_step, agent_step;
    if (e->c.choose_goal_last) { goal_step = (e->adv_step == e->adv_max - 2); agent_step = (e->adv_step == e->adv_max - 1); }
    else { goal_step = (e->adv_step == 0); agent_step = (e->adv_step == 1); }
    if (goal_step) { /* goal_noise == 0 only: remove_wall + put_obj(Goal) */
      e->cells[y * W + x] = CELL_GOAL; e->gx = x; e->gy = y;
    } else if (agent_step) {
      if (e->cells[y * W + x] == CELL_WALL) e->cells[y * W + x] = CELL_EMPTY; /* remove_wall */
      if (e->cells[y * W + x] != CELL_EMPTY) { /* goal already here: place_one_agent(0, rand_dir=False) */
        int px, py; e->has_agent = 0;
        place_random(e, -1, &px, &py);
        e->sx = px; e->sy = py;
      } else { e->sx = x; e->sy = y; }
      e->has_agent = 1; e->ax = e->sx; e->ay = e->sy;
    } else { /* wall, only if the cell is empty (and not the agent's, which only matters goal-first) */
      if (e->cells[y * W + x] == CELL_EMPTY && !(e->has_agent && x == e->ax && y == e->ay))
        e->cells[y * W + x] = CELL_WALL;
    }
  }
  e->adv_step++;
  if (e->adv_step >= e->c.n_clutter + 2) { reset_metrics(e); compute_metrics(e); return 1; }
  return 0;
}

/* reset_random (adversarial.py:541-581).  n_walls < 0 -> int(n_clutter/2); else the caller supplies
 * np.random.randint(0, n_clutter) (the GLOBAL-rng draw of _resample_n_clutter, adversarial.py:151-156). */
int mgo_reset_random(mgo_env *e, int n_walls) {
  if (e->c.fixed_environment) mgo_seed(e, e->limbs, e->n_limbs);
  e->step_count = 0; e->adv_step = 0;
  reset_agent_status(e);
  e->sx = e->sy = -1; e->gx = e->gy = -1;
  reset_metrics(e);
  gen_grid(e);
  int x, y;
  if (place_random(e, 100, &x, &y)) return 2;
  e->cells[y * e->c.W + x] = CELL_GOAL; e->gx = x; e->gy = y;
  e->sdir = env_randint(e, 0, 4);
  place_random(e, -1, &x, &y);
  e->sx = x; e->sy = y; e->has_agent = 1; e->ax = x; e->ay = y;
  if (n_walls < 0) n_walls = e->c.n_clutter / 2;
  else { e->adv_max = n_walls + 2; e->n_clutter_sampled = 1; }
  for (int i = 0; i < n_walls; i++) {
    if (place_random(e, 100, &x, &y)) return 2;
    e->cells[y * e->c.W + x] = CELL_WALL;
  }
  compute_metrics(e);
  return mgo_reset_agent(e);
}

/* Grid.encode() of the full grid incl. the agent object (adversarial.py:162-164); out[x][y][3]. */
void mgo_encode(const mgo_env *e, uint8_t *out) {
  int W = e->c.W;
  for (int x = 0; x < W; x++)
    for (int y = 0; y < W; y++) {
      uint8_t *o = out + (x * W + y) * 3;
      uint8_t c = e->cells[y * W + x];
      if (e->has_agent && x == e->ax && y == e->ay) { o[0] = 10; o[1] = 0; o[2] = (uint8_t)e->adir; }
      else if (c == CELL_WALL) { o[0] = 2; o[1] = 5; o[2] = 0; }
      else if (c == CELL_GOAL) { o[0] = 8; o[1] = 1; o[2] = 0; }
      else { o[0] = 1; o[1] = 0; o[2] = 0; }
    }
}

/* reset_to_level, byte form (adversarial.py:271-294 + Grid.set_encoding multigrid.py:264-280):
 * reset() (fresh start dir!), decode enc[x][y][3], compute_metrics, reset_agent. */
int mgo_reset_to_encoding(mgo_env *e, const uint8_t *enc) {
  int W = e->c.W;
  mgo_reset(e);
  for (int i = 0; i < W; i++)
    for (int j = 0; j < W; j++) {
      uint8_t t = enc[(i * W + j) * 3];
      uint8_t c = CELL_EMPTY;
      if (t == 2) c = CELL_WALL;
      else if (t == 8) { c = CELL_GOAL; e->gx = i; e->gy = j; }
      else if (t == 10) { e->sx = i; e->sy = j; }
      e->cells[j * W + i] = c;
    }
  compute_metrics(e);
  e->elapsed = 0;
  return mgo_reset_agent(e);
}

/* reset_to_level, action-string form (adversarial.py:274-283). */
int mgo_reset_to_actions(mgo_env *e, const int *locs, int n) {
  int rc = 0, err = 0;
  mgo_reset(e);
  if (e->c.resample_n_clutter) e->adv_max = n;
  for (int i = 0; i < n; i++) {
    int done = mgo_step_adversary(e, locs[i], &err);
    if (err) return 3;
    if (done) rc = mgo_reset_agent(e);
  }
  e->elapsed = 0;
  return rc;
}

/* mutate_level (adversarial.py:317-397) with the GLOBAL-rng draws made explicit:
 *   locs[k], ops[k]: edit_locs in the reference's iteration order and editor-action indices
 *     (0 '-', 1 '.', then 'a','g' or 'g' depending on the editor set, adversarial.py:40-56);
 *   goal_choice / agent_choice: index into the row-major list of free interior cells that
 *     np.random.choice(free_idx) picked (adversarial.py:308-315), used only if needed.
 * need[0]/need[1] report whether the goal / agent fallback fired, nfree[] the list lengths. */
int mgo_mutate(mgo_env *e, const int *locs, const int *ops, int k, int goal_choice, int agent_choice,
               int *need, int *nfree) {
  int W = e->c.W, I = W - 2;
  uint8_t freem[MGO_MAXW * MGO_MAXW];
  for (int y = 0; y < I; y++)
    for (int x = 0; x < I; x++) freem[y * I + x] = (e->cells[(y + 1) * W + (x + 1)] != CELL_WALL);
  freem[(e->sy - 1) * I + (e->sx - 1)] = 0;
  freem[(e->gy - 1) * I + (e->gx - 1)] = 0;
  char opch[4] = {'-', '.', 'g', 'g'};
  if (e->c.n_editor_actions == 4) { opch[2] = 'a'; opch[3] = 'g'; }
  for (int n = 0; n < k; n++) {
    int loc = locs[n], x = loc % I + 1, y = loc / I + 1;
    char a = opch[ops[n]];
    /* _clean_loc (adversarial.py:296-306): the Agent object sits at the agent's CURRENT cell */
    if (e->cells[y * W + x] == CELL_GOAL) { e->gx = e->gy = -1; }
    else if (e->has_agent && x == e->ax && y == e->ay) { e->sx = e->sy = -1; e->has_agent = 0; }
    e->cells[y * W + x] = CELL_EMPTY;
    if (a == '-') { e->cells[y * W + x] = CELL_WALL; freem[(y - 1) * I + (x - 1)] = 0; }
    else if (a == '.') { freem[(y - 1) * I + (x - 1)] = 1; }
    else if (a == 'a') {
      if (e->sx >= 0) { freem[(e->sy - 1) * I + (e->sx - 1)] = 1; /* grid.set(ax,ay,None): agent object removed */ }
      e->has_agent = 1; e->ax = x; e->ay = y; e->adir = 0; /* place_one_agent -> rand_dir=True -> dir 0 */
      e->sx = x; e->sy = y; freem[(y - 1) * I + (x - 1)] = 0;
    } else { /* 'g' */
      if (e->gx >= 0) { e->cells[e->gy * W + e->gx] = CELL_EMPTY; freem[(e->gy - 1) * I + (e->gx - 1)] = 1; }
      e->cells[y * W + x] = CELL_GOAL; e->gx = x; e->gy = y; freem[(y - 1) * I + (x - 1)] = 0;
    }
  }
  need[0] = need[1] = 0; nfree[0] = nfree[1] = 0;
  for (int pass = 0; pass < 2; pass++) {
    int missing = pass == 0 ? (e->gx < 0) : (e->sx < 0);
    if (!missing) continue;
    need[pass] = 1;
    int cnt = 0, pick = pass == 0 ? goal_choice : agent_choice, sel = -1;
    for (int i = 0; i < I * I; i++) if (freem[i]) { if (cnt == pick) sel = i; cnt++; }
    nfree[pass] = cnt;
    if (sel < 0) return 4;
    freem[sel] = 0;
    int x = sel % I + 1, y = sel / I + 1;
    if (pass == 0) { e->cells[y * W + x] = CELL_GOAL; e->gx = x; e->gy = y; }
    else { e->has_agent = 1; e->ax = x; e->ay = y; e->adir = 0; e->sx = x; e->sy = y; }
  }
  e->step_count = 0; e->adv_step = 0;
  reset_metrics(e); compute_metrics(e);
  return mgo_reset_agent(e);
}

/* gen_obs -> gen_obs_grid -> slice/rotate_left/process_vis/encode (multigrid.py:977-1055,320-338,
 * 300-318,749-782; gym_minigrid Grid.process_vis/encode).  Implemented literally: slice the
 * window, rotate it left dir+1 times, run the visibility sweep, blank the agent cell, encode.
 * out[vx][vy][3] uint8. */
void mgo_gen_obs(const mgo_env *e, uint8_t *out) {
  const int V = MGO_V, W = e->c.W;
  int tx, ty;
  switch (e->adir) { /* get_view_exts */
    case 0: tx = e->ax; ty = e->ay - V / 2; break;
    case 1: tx = e->ax - V / 2; ty = e->ay; break;
    case 2: tx = e->ax - V + 1; ty = e->ay - V / 2; break;
    default: tx = e->ax - V / 2; ty = e->ay - V + 1; break;
  }
  /* codes: 0 empty 1 wall 2 goal 3 agent(self) ; g[j*V+i] like Grid.grid */
  uint8_t g[MGO_V * MGO_V], g2[MGO_V * MGO_V];
  for (int j = 0; j < V; j++)
    for (int i = 0; i < V; i++) {
      int x = tx + i, y = ty + j;
      uint8_t c;
      if (x >= 0 && x < W && y >= 0 && y < W) {
        c = e->cells[y * W + x];
        if (e->has_agent && x == e->ax && y == e->ay) c = 3;
      } else c = CELL_WALL;
      g[j * V + i] = c;
    }
  for (int r = 0; r < e->adir + 1; r++) { /* rotate_left: new(j, V-1-i) = old(i, j) */
    for (int i = 0; i < V; i++)
      for (int j = 0; j < V; j++) g2[(V - 1 - i) * V + j] = g[j * V + i];
    memcpy(g, g2, sizeof(g));
  }
  uint8_t mask[MGO_V][MGO_V]; /* [i][j] */
  if (!e->c.see_through) {
    memset(mask, 0, sizeof(mask));
    mask[V / 2][V - 1] = 1;
    for (int j = V - 1; j >= 0; j--) {
      for (int i = 0; i < V - 1; i++) {
        if (!mask[i][j]) continue;
        if (g[j * V + i] == CELL_WALL) continue;
        mask[i + 1][j] = 1;
        if (j > 0) { mask[i + 1][j - 1] = 1; mask[i][j - 1] = 1; }
      }
      for (int i = V - 1; i >= 1; i--) {
        if (!mask[i][j]) continue;
        if (g[j * V + i] == CELL_WALL) continue;
        mask[i - 1][j] = 1;
        if (j > 0) { mask[i - 1][j - 1] = 1; mask[i][j - 1] = 1; }
      }
    }
  } else memset(mask, 1, sizeof(mask));
  g[(V - 1) * V + V / 2] = CELL_EMPTY; /* agent's own cell -> None (carrying is always None) */
  for (int i = 0; i < V; i++)
    for (int j = 0; j < V; j++) {
      uint8_t *o = out + (i * V + j) * 3;
      uint8_t c = g[j * V + i];
      if (!mask[i][j]) { o[0] = o[1] = o[2] = 0; }
      else if (c == CELL_WALL) { o[0] = 2; o[1] = 5; o[2] = 0; }
      else if (c == CELL_GOAL) { o[0] = 8; o[1] = 1; o[2] = 0; }
      else if (c == 3) { o[0] = 10; o[1] = 0; o[2] = 0; } /* unreachable for a single agent */
      else { o[0] = 1; o[1] = 0; o[2] = 0; }
    }
}

/* MultiGridEnv.step / step_one_agent / agent_is_done (multigrid.py:943-975,866-941,821-838).
 * Returns done; *reward is the Python float (double). */
int mgo_step(mgo_env *e, int action, double *reward) {
  int W = e->c.W;
  static const int DX[4] = {1, 0, -1, 0}, DY[4] = {0, 1, 0, -1};
  e->step_count++;
  *reward = 0.0;
  int fx = e->ax + DX[e->adir], fy = e->ay + DY[e->adir];
  if (action == 0) e->adir = (e->adir + 3) & 3;
  else if (action == 1) e->adir = (e->adir + 1) & 3;
  else if (action == 2) {
    uint8_t fc = e->cells[fy * W + fx];
    if (fc == CELL_GOAL) {
      e->has_agent = 0; e->done_flag = 1;         /* agent_is_done: remove agent, done */
      int px, py; place_random(e, -1, &px, &py);  /* place_one_agent: env RNG, rand_dir=True -> dir 0 */
      e->has_agent = 1; e->ax = px; e->ay = py; e->adir = 0;
      *reward = 1.0 - 0.9 * ((double)e->step_count / (double)e->c.max_steps); /* MiniGridEnv._reward */
    } else if (fc == CELL_EMPTY) { e->ax = fx; e->ay = fy; }
  }
  /* actions 3..6: pickup/drop/toggle/done are no-ops in a maze of walls and a goal */
  return e->done_flag || e->step_count >= e->c.max_steps;
}

/* One vectorised step_env transition for one env, folding the wrapper chain
 * (parallel_wrappers.py:27-37 worker.step_env, time_limit.py:24-33, vec_monitor.py:60-85,
 * obs_wrappers.py:88-115,168-181).  flags: bit0 done, bit1 'truncated' key present,
 * bit2 truncated value, bit3 goal reached.  obs_u8/trunc_u8 are [vx][vy][3]. */
#define MGO_F_DONE 1
#define MGO_F_TRUNC_KEY 2
#define MGO_F_TRUNC_VAL 4
#define MGO_F_GOAL 8
int mgo_step_env(mgo_env *e, int action, int reset_random, int n_walls_resample,
                 uint8_t *obs_u8, int *dir_out, float *rew_out, uint8_t *trunc_u8, int *trunc_dir,
                 float *ep_r, int *ep_l) {
  double rew;
  int flags = 0;
  int done = mgo_step(e, action, &rew);
  if (rew != 0.0) flags |= MGO_F_GOAL;
  e->elapsed++;
  if (e->elapsed >= e->c.max_episode_steps) {
    flags |= MGO_F_TRUNC_KEY;
    if (!done) flags |= MGO_F_TRUNC_VAL;
    if (trunc_u8) { mgo_gen_obs(e, trunc_u8); *trunc_dir = e->adir; }
    done = 1;
  }
  /* VecMonitor: eprets(f32) += rews(f64) -> computed in double, stored as f32 */
  e->ep_ret = (float)((double)e->ep_ret + rew);
  e->ep_len += 1;
  if (done) {
    flags |= MGO_F_DONE;
    *ep_r = e->ep_ret; *ep_l = e->ep_len;
    e->ep_ret = 0.f; e->ep_len = 0;
    if (reset_random) mgo_reset_random(e, e->c.resample_n_clutter ? n_walls_resample : -1);
    else mgo_reset_agent(e);
  }
  mgo_gen_obs(e, obs_u8);
  *dir_out = e->adir;
  *rew_out = (float)rew;
  return flags;
}

/* VecPreprocessImageWrapper._preprocess: image/10.0 (double) -> transpose [c][vx][vy] -> float32
 * (obs_wrappers.py:104-110, util/__init__.py:197-200). */
void mgo_preprocess(const uint8_t *u8, int n_cells, float *out) {
  for (int i = 0; i < n_cells; i++)
    for (int c = 0; c < 3; c++) out[c * n_cells + i] = (float)((double)u8[i * 3 + c] / 10.0);
}

/* ---------------------------------------------------------------- batch drivers (CPU baseline) */
/* T vector steps over N envs with recorded actions[t*N+i]; obs written as float32 [T][N][3][5][5]
 * (obs_out may be NULL -> per-thread scratch; the observation is still rendered and preprocessed).
 * Envs are split over n_threads pthreads (libgomp is absent from this image).  Returns the number
 * of episodes finished. */
#include <pthread.h>
typedef struct {
  mgo_env *envs; int N, T, lo, hi, reset_random; const uint8_t *actions;
  float *obs_out, *rew_out; uint8_t *flags_out; long episodes;
} mgo_job;

static void *rollout_worker(void *arg) {
  mgo_job *j = (mgo_job *)arg;
  long episodes = 0;
  for (int i = j->lo; i < j->hi; i++) {
    uint8_t u8[75], tr[75]; float scratch[75]; int d, td, l = 0; float r, er = 0.f;
    for (int t = 0; t < j->T; t++) {
      int f = mgo_step_env(&j->envs[i], j->actions[(size_t)t * j->N + i], j->reset_random, -1, u8, &d, &r, tr, &td, &er, &l);
      float *o = j->obs_out ? j->obs_out + ((size_t)t * j->N + i) * 75 : scratch;
      mgo_preprocess(u8, 25, o);
      if (j->rew_out) j->rew_out[(size_t)t * j->N + i] = r;
      if (j->flags_out) j->flags_out[(size_t)t * j->N + i] = (uint8_t)f;
      episodes += (f & MGO_F_DONE);
    }
  }
  j->episodes = episodes;
  return NULL;
}

long mgo_rollout_batch(mgo_env *envs, int N, int T, const uint8_t *actions, int reset_random,
                       float *obs_out, float *rew_out, uint8_t *flags_out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256]; mgo_job jobs[256];
  long episodes = 0;
  for (int k = 0; k < n_threads; k++) {
    mgo_job *j = &jobs[k];
    j->envs = envs; j->N = N; j->T = T; j->reset_random = reset_random; j->actions = actions;
    j->obs_out = obs_out; j->rew_out = rew_out; j->flags_out = flags_out; j->episodes = 0;
    j->lo = (int)((long)N * k / n_threads); j->hi = (int)((long)N * (k + 1) / n_threads);
    if (n_threads == 1) rollout_worker(j); else pthread_create(&th[k], NULL, rollout_worker, j);
  }
  for (int k = 0; k < n_threads; k++) { if (n_threads > 1) pthread_join(th[k], NULL); episodes += jobs[k].episodes; }
  return episodes;
}

/* ---------------------------------------------------------------- accessors for ctypes */
void mgo_get_state(const mgo_env *e, int *out) {
  out[0] = e->has_agent; out[1] = e->ax; out[2] = e->ay; out[3] = e->adir;
  out[4] = e->gx; out[5] = e->gy; out[6] = e->sx; out[7] = e->sy; out[8] = e->sdir;
  out[9] = e->step_count; out[10] = e->elapsed; out[11] = e->done_flag;
  out[12] = e->adv_step; out[13] = e->adv_max; out[14] = e->n_clutter_sampled;
  out[15] = e->n_clutter_placed; out[16] = e->dist; out[17] = e->passable; out[18] = e->spl;
  out[19] = e->ep_len; out[20] = e->rng_words; out[21] = e->error;
}
void mgo_get_cells(const mgo_env *e, uint8_t *out) { memcpy(out, e->cells, (size_t)e->c.W * e->c.W); }
uint32_t mgo_rng_next(mgo_env *e) { e->rng_words++; return mt_next(&e->rng); }
int mgo_rng_randint(mgo_env *e, int lo, int hi) { return env_randint(e, lo, hi); }
