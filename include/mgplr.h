/*
 * mgplr.h -- C ABI of the B200-native MultiGrid / PLR hot path (libmgplr.so).
 *
 * Plain pointers and sizes only: no torch types.  Device pointers are raw CUDA device addresses
 * (a PyTorch tensor's data_ptr()), `stream` is a cudaStream_t passed as void* (0 = legacy default
 * stream).  Every entry point returns 0 on success, a positive cudaError_t, or a negative
 * MGPLR_E_* argument error; mgplr_last_error() returns a thread-local message.
 *
 * The reference (linjiw/dcd-isaac) has no FFI layer: its boundary is the duck-typed Python objects
 * consumed by envs/runners/adversarial_runner.py.  Each entry point below therefore names the
 * reference method it replaces (file:line relative to the reference root); the Python mirror of
 * those objects lives in dcd_isaac_b200/{vec_env,level_sampler,level_store}.py and binds these
 * symbols with ctypes (INTEGRATION.md shows the stub).
 *
 * State layout, kernels and rooflines: DESIGN.md.
 */
#ifndef MGPLR_H
#define MGPLR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGPLR_ABI_VERSION 4

#define MGPLR_E_BADARG (-1)
#define MGPLR_E_UNSUPPORTED (-2)

/* step_env / info flag bits (uint8 per env) */
#define MGPLR_F_DONE 1u      /* done returned to the runner (goal, env max_steps or TimeLimit) */
#define MGPLR_F_TRUNC_KEY 2u /* 'truncated' in info (envs/wrappers/time_limit.py:28-31) */
#define MGPLR_F_TRUNC_VAL 4u /* info['truncated'] is True */
#define MGPLR_F_GOAL 8u      /* goal reached this step (reward != 0) */
#define MGPLR_F_ERROR 16u    /* the env finished and has a pending error bit (mgplr_get_errors): the auto-reset failed */

/* AdversarialEnv constructor arguments (envs/multigrid/adversarial.py:67-79) + the registered
 * TimeLimit (envs/registration.py:118-120, envs/wrappers/time_limit.py:15-22). */
typedef struct mgplr_env_config {
  int32_t width;              /* size: grid side W, 5 <= W <= 32 */
  int32_t agent_view_size;    /* must be 5 (every registered adversarial env) */
  int32_t max_steps;          /* env max_steps */
  int32_t max_episode_steps;  /* TimeLimit._max_episode_steps (< 32768) */
  int32_t see_through_walls;  /* 1: no occlusion; 0: gym_minigrid process_vis occlusion */
  int32_t n_clutter;
  int32_t resample_n_clutter;
  int32_t choose_goal_last;
  int32_t fixed_environment;
  int32_t n_editor_actions;   /* 2 '-.', 3 '-.g', 4 '-.ag' (adversarial.py:40-56) */
} mgplr_env_config;

typedef struct mgplr_venv mgplr_venv; /* opaque: N environments resident in HBM */

/* Destinations of one vectorised step (all device pointers; any may be NULL = not wanted).
 * Layouts are those of algos/storage.py:62-112 so the kernel writes straight into
 * RolloutStorage: obs['image'][t+1], obs['direction'][t+1], rewards[t], masks[t+1], ... */
typedef struct mgplr_step_out {
  float *image;             /* f32 [N][3][5][5] = uint8/10.0, channels first (obs_wrappers.py:104-110) */
  float *direction;         /* f32 [N][1] */
  float *reward;            /* f32 [N][1] */
  uint8_t *flags;           /* u8  [N]  MGPLR_F_* */
  float *ep_return;         /* f32 [N]  VecMonitor info['episode']['r'], valid where DONE */
  int32_t *ep_length;       /* i32 [N]  info['episode']['l'], valid where DONE */
  float *trunc_image;       /* f32 [N][3][5][5] info['truncated_obs'], written where TRUNC_KEY */
  float *trunc_direction;   /* f32 [N][1] */
  float *masks;             /* f32 [N][1] 1-done            (adversarial_runner.py:566-567) */
  float *bad_masks;         /* f32 [N][1] 0 iff TRUNC_KEY   (adversarial_runner.py:568-570) */
  float *cliffhanger_masks; /* f32 [N][1] 0 iff cliffhanger (adversarial_runner.py:571-573) */
  uint8_t *image_u8;        /* u8  [N][5][5][3] raw gym_minigrid encoding (packed secondary layout) */
  float *trunc_full_obs;    /* f32 [N][3][W][W] info['truncated_obs']['full_obs'] (MultiGridFullyObsWrapper), where TRUNC_KEY */
} mgplr_step_out;

const char *mgplr_last_error(void);
int mgplr_abi_version(void);

/* util.create_parallel_env (util/__init__.py:184-220): allocate N envs on `device`; each starts as
 * the empty walled grid of AdversarialEnv.__init__ -> reset(). */
int mgplr_venv_create(const mgplr_env_config *cfg, int32_t num_envs, int32_t device, mgplr_venv **out);
void mgplr_venv_destroy(mgplr_venv *v);
int32_t mgplr_venv_num_envs(const mgplr_venv *v);
/* bytes of HBM held by the handle */
int64_t mgplr_venv_state_bytes(const mgplr_venv *v);

/* venv.set_seed / venv.seed(seed, index) (parallel_wrappers.py:268-270,415-416 -> multigrid.py:465-468
 * -> gym seeding.np_random): limbs = _int_list_from_bigint(hash_seed(seed)), HOST array [n][2] (+ count
 * [n]); index = NULL seeds envs 0..n-1.  Runs MT19937 init_by_array on the device. */
int mgplr_seed(mgplr_venv *v, const uint32_t *limbs_host, const int32_t *n_limbs_host, const int32_t *index_host,
               int32_t n, void *stream);

/* venv.reset() -> AdversarialEnv.reset (adversarial.py:194-229).  adv_image f32 [N][3][W][W] (=/10, CHW),
 * time_step f32 [N][1].  random_z is drawn by the host (global np.random, adversarial.py:449-450). */
int mgplr_reset(mgplr_venv *v, float *adv_image, float *time_step, void *stream);

/* venv.step_adversary(action) (parallel_wrappers.py:288-297 -> adversarial.py:452-539).  loc i64 [N].
 * done u8 [N].  Out-of-range locations set the env's error flag (mgplr_get_errors) and do nothing. */
int mgplr_step_adversary(mgplr_venv *v, const int64_t *loc, float *adv_image, float *time_step, uint8_t *done,
                         void *stream);

/* venv.reset_agent() (parallel_wrappers.py:314-321 -> adversarial.py:238-269, time_limit.py:46-48,
 * vec_monitor.py:42-46).  Only `image`, `direction`, `image_u8` of `out` are used. */
int mgplr_reset_agent(mgplr_venv *v, const mgplr_step_out *out, void *stream);

/* venv.reset_random() (parallel_wrappers.py:324-331 -> adversarial.py:541-581).  n_walls i32 [N] or NULL:
 * the global-rng draw of _resample_n_clutter for resample_n_clutter envs (else int(n_clutter/2)). */
int mgplr_reset_random(mgplr_venv *v, const int32_t *n_walls, const mgplr_step_out *out, void *stream);

/* venv.reset_to_level(level, index) / reset_to_level_batch(levels), byte form
 * (parallel_wrappers.py:334-349 -> adversarial.py:271-294, multigrid.py:264-280).
 * enc u8 [n][W][W][3] (device); index i32 [n] (device) or NULL = envs 0..n-1.  Outputs (only image /
 * direction / image_u8) are written at row `index[k]` of the full-N arrays. */
int mgplr_reset_to_encoding(mgplr_venv *v, const uint8_t *enc, const int32_t *index, int32_t n,
                            const mgplr_step_out *out, void *stream);

/* Fixed-level environments (zero-shot evaluation mazes, envs/multigrid/maze.py:23-94 MazeEnv._gen_grid + MultiGridEnv.reset,
 * multigrid.py:470-502): load byte-encoded levels WITHOUT touching the env RNG and with an explicit start direction
 * (place_agent_at_pos(rand_dir=True) forces 0, multigrid.py:668-672).  enc u8 [n_levels][W][W][3] (device); env e gets
 * level level_index[e] (i32 [N] device) or level 0 when level_index is NULL.  Metrics are recomputed, the agent is reset
 * to the start, step / episode counters are cleared. */
int mgplr_load_levels(mgplr_venv *v, const uint8_t *enc, int32_t n_levels, const int32_t *level_index, int32_t start_dir,
                      const mgplr_step_out *out, void *stream);

/* The same for n of the envs with one level each: env env_index[k] (i32 [n], device) gets level k of enc u8 [n][W][W][3];
 * the other envs are untouched.  For environments that regenerate their level on every reset with a host-side generator
 * (the Kruskal perfect mazes, envs/multigrid/mst_maze.py:55-115: worker auto-reset, parallel_wrappers.py:20-25).
 * Outputs are written at rows env_index[k]. */
int mgplr_load_levels_at(mgplr_venv *v, const uint8_t *enc, const int32_t *env_index, int32_t n, int32_t start_dir,
                         const mgplr_step_out *out, void *stream);

/* (start_dir < 0 in mgplr_load_levels_at: every env starts facing the direction stored in its encoding's agent cell -- envs
 * whose generator also draws the start direction, envs/multigrid/fourrooms.py:71-78.) */

/* Same, action-string form: locs i32 [n][len] (device), replayed through step_adversary. */
int mgplr_reset_to_actions(mgplr_venv *v, const int32_t *locs, int32_t len, const int32_t *index, int32_t n,
                           const mgplr_step_out *out, void *stream);

/* venv.mutate_level(num_edits) (parallel_wrappers.py:352-359 -> adversarial.py:317-397) in two phases so the
 * host can make the global-np.random draws exactly where the reference makes them:
 *  edits:    locs i32 [N][max_edits], ops i32 [N][max_edits], n_edits i32 [N] (iteration order of
 *            list(set(randint)), editor-action indices).  need u8 [N][2] / n_free i32 [N][2] report which
 *            fallbacks (goal, agent) are required and the length of the free-cell list.
 *  finalize: choice i32 [N][2] = index into the row-major free list picked by np.random.choice; then metrics
 *            and reset_agent.  */
int mgplr_mutate_edits(mgplr_venv *v, const int32_t *locs, const int32_t *ops, const int32_t *n_edits,
                       int32_t max_edits, uint8_t *need, int32_t *n_free, void *stream);
int mgplr_mutate_finalize(mgplr_venv *v, const int32_t *choice, const mgplr_step_out *out, void *stream);

/* venv.step_env(action, reset_random) (vec_env.py:113-118, parallel_wrappers.py:27-37,299-311) with the
 * wrapper chain folded in (time_limit.py:24-33, vec_monitor.py:60-85, obs_wrappers.py:168-181).
 * action i64 [N] (device).  last_step bit 0: this is the rollout's last step -- adversarial_runner.py:530 forces
 * done, so masks = 0 for every env; bit 1: use_proper_time_limits -- envs that are not done at that step are
 * cliffhangers (adversarial_runner.py:521-528): cliffhanger_masks = 0, bad_masks = 0, and their observation is also
 * written to trunc_image / trunc_direction.  n_walls as in mgplr_reset_random. */
int mgplr_step_env(mgplr_venv *v, const int64_t *action, int32_t reset_random, const int32_t *n_walls,
                   int32_t last_step, const mgplr_step_out *out, void *stream);

/* The same with a narrow action stream: action u8 [N] (device), values 0..6 -- 1 byte per env instead of 8. */
int mgplr_step_env_u8(mgplr_venv *v, const uint8_t *action, int32_t reset_random, const int32_t *n_walls, int32_t last_step,
                      const mgplr_step_out *out, void *stream);

/* One finished episode, as the host needs it to build info['episode'] (vec_monitor.py:66-74). */
typedef struct mgplr_done_record {
  int32_t env;        /* env index */
  float reward;       /* reward of the terminating step */
  float ep_return;    /* info['episode']['r'] */
  int32_t ep_length;  /* info['episode']['l'] in the low 24 bits, the step's MGPLR_F_* flags of this env in the top byte (ABI v4):
                       * the record list alone tells the host everything the flags array does -- every env with a non-zero flag
                       * is DONE -- so flags_host may be NULL and 1 byte per env less crosses PCIe per step */
} mgplr_done_record;
#define MGPLR_DONE_LENGTH(rec_ep_length) ((rec_ep_length) & 0xffffff)
#define MGPLR_DONE_FLAGS(rec_ep_length) (((uint32_t)(rec_ep_length)) >> 24)

/* The same transition driven from HOST buffers (the reference's calling convention: actions arrive as a CPU
 * tensor, adversarial_runner.py:512-517; done / infos go back to the host).  Observations, rewards and masks stay
 * in HBM at the `out_dev` destinations (rollout storage).  With PINNED host buffers (cudaHostAlloc /
 * cudaHostRegister / torch pin_memory) the call is one kernel launch + one stream synchronisation: the kernel reads
 * action i64 [N] straight from the caller's memory over PCIe, writes flags u8 [N] straight into flags_host (optional: NULL skips
 * it -- the done records carry the flags of every env that has any) and appends
 * the (few) done records to a device-mapped pinned list owned by the handle.  Pageable buffers work too (staged
 * copies).  done_host receives min(*n_done_host, done_capacity) records in unspecified order.  Use one stream per
 * handle for these calls. */
int mgplr_step_env_host(mgplr_venv *v, const int64_t *action_host, int32_t reset_random, int32_t last_step,
                        const mgplr_step_out *out_dev, uint8_t *flags_host, mgplr_done_record *done_host,
                        int32_t done_capacity, int32_t *n_done_host, void *stream);

/* mgplr_step_env_host with uint8 actions (MultiGrid has 7 actions): 8x less PCIe traffic per vector step; the caller narrows
 * `action.cpu()` (adversarial_runner.py:512) once.  Everything else as above. */
int mgplr_step_env_host_u8(mgplr_venv *v, const uint8_t *action_host, int32_t reset_random, int32_t last_step,
                           const mgplr_step_out *out_dev, uint8_t *flags_host, mgplr_done_record *done_host,
                           int32_t done_capacity, int32_t *n_done_host, void *stream);

/* T consecutive step_env transitions in ONE launch from a recorded action stream u8 [T][N]: env state
 * stays in shared memory / registers across steps (replayed-seed evaluation, random-policy rollouts).
 * Outputs are the [T]-leading versions of mgplr_step_out fields: image f32 [T][N][3][5][5], direction
 * [T][N][1], reward [T][N][1], flags u8 [T][N]; masks f32 [T][N][1] etc.  Any may be NULL. */
int mgplr_rollout(mgplr_venv *v, const uint8_t *actions, int32_t T, int32_t reset_random,
                  const mgplr_step_out *out_t0, void *stream);
/* Same, with the runner's last-step rule (adversarial_runner.py:521-530; `last_step` as in mgplr_step_env: bit 0 = this is a
 * rollout's last step, bit 1 = use_proper_time_limits) applied to the mask outputs of step T-1. */
int mgplr_rollout_ex(mgplr_venv *v, const uint8_t *actions, int32_t T, int32_t reset_random, int32_t last_step,
                     const mgplr_step_out *out_t0, void *stream);

/* obs['full_obs'] of MultiGridFullyObsWrapper (envs/wrappers/multigrid_wrappers.py:14-51; used with
 * --use_global_critic / --use_global_policy, util/__init__.py:175-178): the whole grid's encoding with the agent cell
 * (10, 0, dir), channels first and NOT scaled (obs_wrappers.py:108-110 only transposes it): f32 [N][3][W][W]. */
int mgplr_full_obs(mgplr_venv *v, float *full_obs, void *stream);

/* venv.get_images() (parallel_wrappers.py:187-193 -> MultiGridEnv.render(mode='level'), multigrid.py:1105-1140,159-261): RGB
 * screenshots of n levels, images u8 [n][W*32][W*32][3] (device).  tiles u8 [14][32][32][3] (device) is the tile table --
 * index 2*code + highlighted, code 0 empty / 1 wall / 2 goal / 3+dir agent (dcd_isaac_b200/tiles.py renders it from the
 * published gym-minigrid tile algorithm).  index i32 [n] (device) selects the envs, NULL = envs 0..n-1. */
int mgplr_render_images(mgplr_venv *v, const uint8_t *tiles, const int32_t *index, int32_t n, uint8_t *images, void *stream);

/* Getters (parallel_wrappers.py:422-448): encodings u8 [N][W][W][3] = AdversarialEnv.encoding;
 * metrics i32 [N][4] = n_clutter_placed, distance_to_goal, passable, shortest_path_length. */
int mgplr_get_encodings(mgplr_venv *v, uint8_t *enc, void *stream);
int mgplr_get_metrics(mgplr_venv *v, int32_t *metrics, void *stream);
/* i32 [N][8]: agent x, y, dir, step_count, elapsed, adversary_step_count, adversary_max_steps, rng words used */
int mgplr_get_agent_state(mgplr_venv *v, int32_t *state, void *stream);
/* u32 [N]: bit0 RetriesExceeded (multigrid.py:597-599), bit1 reset_agent without start position
 * (adversarial.py:248-249), bit2 step_adversary loc out of range (adversarial.py:465-466).  Sticky until read
 * with clear != 0. */
int mgplr_get_errors(mgplr_venv *v, uint32_t *errors, int32_t clear, void *stream);
/* next `count` raw MT19937 words of env `index` WITHOUT consuming them (test introspection) */
int mgplr_peek_rng(mgplr_venv *v, int32_t index, uint32_t *words_host, int32_t count);

/* ---- rollout math (algos/storage.py, level_replay/level_sampler.py); stateless launchers ---- */

/* RolloutStorage.compute_gae_returns (algos/storage.py:233-256).  rewards f32 [T][N], value_preds f32
 * [T+1][N] (already the truncated/denormalised buffer the reference would use, with value_preds[T] =
 * next_value), masks f32 [T+1][N]; returns f32 [T+1][N] gets rows 0..T-1.  Bit-identical op order. */
int mgplr_gae(const float *rewards, const float *value_preds, const float *masks, float *returns, int32_t T,
              int32_t N, double gamma, double gae_lambda, void *stream);

/* RolloutStorage.compute_discounted_returns (algos/storage.py:258-279): returns[T] must hold the bootstrap value
 * (value_preds[-1], or its truncated / denormalised version); rows T-1..0 are filled with
 * returns[t+1] * gamma * masks[t+1] + rewards[t] in float32, the reference's operand order. */
int mgplr_discounted_returns(const float *rewards, const float *masks, float *returns, int32_t T, int32_t N, double gamma,
                             void *stream);

/* RolloutStorage.get_batched_value_loss(batched=True) (algos/storage.py:290-327), used for ACCEL's base-level scores
 * (adversarial_runner.py:608-613,762-770).  returns / value_preds f32 [T+1][N] (rows 0..T-1 are used).
 * mode 0: |returns - value|, 1: signed, 2: positive part; power > 1 raises the per-step term to that power;
 * out f32 [N] = mean over the T steps, clamped to [-1, 1] when `clipped`. */
int mgplr_batched_value_loss(const float *returns, const float *value_preds, int32_t T, int32_t N, int32_t mode,
                             int32_t power, int32_t clipped, float *out, void *stream);

#define MGPLR_SCORE_POSITIVE_VALUE_LOSS 0
#define MGPLR_SCORE_SIGNED_VALUE_LOSS 1
#define MGPLR_SCORE_VALUE_L1 2
#define MGPLR_SCORE_MAX_MC 3 /* per-episode pieces of grounded_* : sum of rewards and value sums */
#define MGPLR_SCORE_LEAST_CONFIDENCE 4 /* 1 - max softmax probability           (level_sampler.py:288-296) */
#define MGPLR_SCORE_MIN_MARGIN 5       /* mean: 1 - mean(p1 - p2), max: 1 - min  (level_sampler.py:298-306) */
#define MGPLR_SCORE_ONE_STEP_TD 6      /* |r[t] + gamma v[t+1] - v[t]|           (level_sampler.py:425-437) */

/* Episode record produced by the scoring kernel, in the reference's actor-major / time-minor order. */
typedef struct mgplr_episode {
  int32_t actor;
  int32_t t_start;
  int32_t t_end;      /* exclusive: the done step */
  int32_t seed;       /* level_seeds[t_start][actor] */
  float mean_score;   /* strategy mean over the episode */
  float max_score;    /* strategy max over the episode */
  float reward_sum;   /* sum of rewards[t_start:t_end] (grounded value, level_sampler.py:534) */
  float value_sum;    /* sum of value_preds[t_start:t_end] */
  float value_min;    /* min of value_preds[t_start:t_end] (max of grounded - v) */
  int32_t cliffhanger; /* 1: cliffhanger_masks[t_end][actor] == 0 -> skipped by the sampler; 2: not-done tail (partial) */
} mgplr_episode;

/* LevelSampler._update_with_rollouts segmentation + score functions (level_sampler.py:486-549,307-349).
 * masks / cliffhanger_masks f32 [T+1][N], returns / value_preds f32 [T+1][N], rewards f32 [T][N],
 * level_seeds i32 [T][N].  Writes up to max_episodes records to `episodes` (device) in canonical order and
 * the count to n_episodes (device i32). */
int mgplr_plr_episode_scores(const float *masks, const float *cliffhanger_masks, const float *returns,
                             const float *value_preds, const float *rewards, const int32_t *level_seeds, int32_t T,
                             int32_t N, int32_t strategy, mgplr_episode *episodes, int32_t max_episodes,
                             int32_t *n_episodes, void *stream);

/* Same with the policy-logit strategies: action_log_dist f32 [T][N][num_actions] (RolloutStorage.action_log_dist; a
 * log_softmax is applied per step as level_sampler.py:513 does) is needed by LEAST_CONFIDENCE / MIN_MARGIN, gamma by
 * ONE_STEP_TD.  action_log_dist may be NULL for the other strategies. */
int mgplr_plr_episode_scores_ex(const float *masks, const float *cliffhanger_masks, const float *returns,
                                const float *value_preds, const float *rewards, const int32_t *level_seeds,
                                const float *action_log_dist, int32_t num_actions, double gamma, int32_t T, int32_t N,
                                int32_t strategy, mgplr_episode *episodes, int32_t max_episodes, int32_t *n_episodes,
                                void *stream);

#define MGPLR_TRANSFORM_CONSTANT 0
#define MGPLR_TRANSFORM_RANK 1
#define MGPLR_TRANSFORM_POWER 2

/* LevelSampler.sample_weights / _score_transform (level_sampler.py:726-785) in fp64 for the constant / rank / power
 * transforms (score and staleness): w = normalise(transform(scores) * seen); s = normalise(transform(staleness) *
 * seen); weights = (1-c) w + c s.  scores / staleness / unseen f64 [n] -> weights f64 [n] (device).  Ties in the
 * rank transform are broken by index (higher index = better rank): np.flip(np.argsort(scores, kind='stable')). */
int mgplr_plr_sample_weights(const double *scores, const double *staleness, const double *unseen, int32_t n,
                             int32_t score_transform, double temperature, double eps, double staleness_coef,
                             int32_t staleness_transform, double staleness_temperature, const double *score_weights,
                             double *weights, void *stream);

/* The score half of sample_weights alone: normalise(transform(scores) * seen) -> score_weights f64 [n].  Scores only
 * change in update_with_rollouts, so callers compute this once per update and pass it as `score_weights` to the two
 * functions around it (NULL there = recompute); only the staleness half is then redone per draw. */
int mgplr_plr_score_weights(const double *scores, const double *unseen, int32_t n, int32_t score_transform,
                            double temperature, double eps, double *score_weights, void *stream);

/* n_draws sequential _sample_replay_level draws (level_sampler.py:664-680 + 601-604) with recorded uniforms
 * u f64 [n_draws] (np.random.choice's single random_sample each): staleness is updated between draws
 * exactly as the reference does.  out_index i32 [n_draws]; staleness f64 [n] updated in place. */
int mgplr_plr_sample_replay(const double *scores, double *staleness, const double *unseen, int32_t n,
                            int32_t score_transform, double temperature, double eps, double staleness_coef,
                            int32_t staleness_transform, double staleness_temperature, const double *score_weights,
                            const double *u, int32_t n_draws, int32_t *out_index, void *stream);

/* The order-dependent half of LevelSampler.update_with_rollouts (level_sampler.py:185-273: update_seed_score,
 * _partial_update_seed_score, _partial_update_seed_score_buffer, _next_buffer_index) on the buffer arrays in HBM: one single-CTA
 * kernel walks `records` (canonical order; n_records_dev = device count written by mgplr_plr_episode_scores, or NULL and
 * n_records) and applies, per finished non-cliffhanger episode, the EWA update (1-alpha) old + alpha (c max + (1-c) mean) of a
 * working seed, or the admission of a staging seed: next free slot, else the slot with the least replay support (argmin of
 * sample_weights on the current state; priority 1: lowest score), replaced iff its score <= the new one, staleness =
 * running_sample_count - staging timestamp.  score_kind 0: record mean / max, 1: uniform (1, 1), 2: MaxMC from reward_sum /
 * value_sum / value_min with the per-slot grounded value (level_sampler.py:351-386,529-547).
 * Seeds are addressed through a table of the rollout's distinct seeds: table_seeds i64 [n_table] sorted, table_index i32 in/out
 * (seed2index entry or -1), table_stamp f64 (staging timestamp or -1 = not in the staging set), table_status i32 out (1 admitted,
 * 2 rejected), admission_log i32 [n_table][2] out (table index, slot) in admission order, counters i32 [4] in/out: admissions,
 * working_seed_buffer_size, not-done tails skipped (left to the host), records walked.  pre f64 [n][4] optional: partial score /
 * max / steps of a stored tail to merge, flush flag.  record_scratch i32 [max_records], scratch_f64 [4 n_buf], scratch_i32 [n_buf]. */
int mgplr_plr_apply_records(const mgplr_episode *records, const int32_t *n_records_dev, int32_t n_records, int32_t max_records,
                            const double *pre, const int64_t *table_seeds, int32_t n_table, int32_t *table_index,
                            const double *table_stamp, int32_t *table_status, int32_t *admission_log, int32_t *counters,
                            int32_t *record_scratch, double *scores, double *staleness, double *unseen, double *grounded,
                            int64_t *seeds, int32_t n_buf, double running_sample_count, double alpha, double max_score_coef,
                            int32_t score_kind, int32_t priority, int32_t score_transform, double temperature, double eps,
                            double staleness_coef, int32_t staleness_transform, double staleness_temperature, double *scratch_f64,
                            int32_t *scratch_i32, void *stream);

/* ---- mazes wider than 32 columns: PerfectMazeLarge (51) / PerfectMazeXL (101), envs/multigrid/mst_maze.py:128-136, part of
 * eval.py's zero-shot maze benchmark (eval.py:340-349).  Evaluation-only: one thread per env, rows of 128 columns, the same
 * view code as the main path; MultiGridEnv.step semantics without a TimeLimit (the mazes are registered without one). ---- */
typedef struct mgplr_wide mgplr_wide;
int mgplr_wide_create(int32_t width, int32_t max_steps, int32_t num_envs, int32_t device, mgplr_wide **out);
void mgplr_wide_destroy(mgplr_wide *h);
/* MultiGridEnv.reset() of n envs with host-generated levels (mst_maze.py:97-115): enc u8 [n][W][W][3] (device), env_index i32
 * [n] (device) or NULL = envs 0..n-1; start_dir as in mgplr_load_levels_at.  Writes out->image / out->direction rows of the
 * loaded envs (full-N arrays). */
int mgplr_wide_load_levels(mgplr_wide *h, const uint8_t *enc, const int32_t *env_index, int32_t n, int32_t start_dir,
                           const mgplr_step_out *out, void *stream);
/* venv.step(action) (multigrid.py:943-975 + vec_monitor.py:60-85 + obs_wrappers.py:104-110): action i64 [N] (device); fills
 * out->image, direction, reward, flags, ep_return, ep_length.  A finished env is put back at its start; the caller then
 * uploads its next maze (the worker's auto-reset, parallel_wrappers.py:20-25). */
int mgplr_wide_step(mgplr_wide *h, const int64_t *action, const mgplr_step_out *out, void *stream);
int mgplr_wide_get_encodings(mgplr_wide *h, uint8_t *enc, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MGPLR_H */
