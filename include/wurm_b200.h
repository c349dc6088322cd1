/* wurm_b200 -- C ABI of the B200-native batched snake environments.
 *
 * This header is the drop-in boundary for the hot path of oscarknagg/wurm: the batched
 * `step` / `reset` / `_observe` of `wurm.envs.SingleSnake` and `wurm.envs.MultiSnake`.  The reference
 * has no FFI of its own (it is pure PyTorch); each entry point below replaces one Python method of
 * the reference, cited as path:line in the reference tree, and is bound from Python with ctypes
 * (wurm_b200/_lib.py; the stub a reference maintainer would add is in INTEGRATION.md).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer borrowed for the duration of the enqueue, in the reference's
 *     own tensor layout.  The library allocates nothing and keeps no state.
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point synchronises.
 *   - Return value: 0 on success, a WURM_E_* code otherwise; wurm_last_error() gives the message
 *     (thread-local).  C++ exceptions never cross this boundary.
 *   - Data-dependent conditions (states outside the supported set, "no available location") are
 *     OR-ed into the device word `status` (WURM_ST_* bits); the caller checks it when it likes.
 *   - Randomness is an input: each draw is either replayed from a caller-supplied tape (the
 *     `*_replay` pointers, how bit-exact parity with the reference is established) or, when the
 *     tape pointer is NULL, derived from Philox4x32-10 keyed by (seed, step, env, stream).
 *     `step` is the caller's call counter; `step_dev` (nullable device pointer) is added to it on
 *     the device, so that a launch captured in a CUDA graph draws fresh numbers at every replay
 *     (the graph bumps the device counter between replays).
 */
#ifndef WURM_B200_H
#define WURM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WURM_ABI_VERSION 6

/* return codes */
#define WURM_OK 0
#define WURM_E_INVALID 1     /* bad argument (size, mode, dtype, NULL pointer, alignment) */
#define WURM_E_UNSUPPORTED 2 /* configuration outside what the kernels implement */
#define WURM_E_CUDA 3        /* CUDA runtime error at launch */

/* bits of the device status word */
#define WURM_ST_MULTI_HEAD 1     /* an env holds more than one head cell (state outside the contract) */
#define WURM_ST_NO_HEAD_PARTIAL 2 /* partial_n observation of an env without a head: zeros written
                                     (the reference raises a view-shape error: single_snake.py:191) */
#define WURM_ST_NO_SPAWN 4       /* no available location to create a snake (multi_snake.py:865,947) */
#define WURM_ST_OVERLAP 8        /* MultiSnake input state with two bodies on one cell at step start */
#define WURM_ST_NOT_COMPACT 16   /* a state handed to wurm_*_compact that the compact records cannot carry exactly */

/* observation modes (single_snake.py:130-195) */
#define WURM_OBS_DEFAULT 0     /* (N,3,S,S) rgb/255 */
#define WURM_OBS_RAW 1         /* (N,3,S,S) copy of the state */
#define WURM_OBS_ONE_CHANNEL 2 /* (N,1,S,S) */
#define WURM_OBS_POSITIONS 3   /* (N,4) head y,x, food y,x */
#define WURM_OBS_PARTIAL 4     /* (N,3*(2n+1)^2) crop around the head */
#define WURM_OBS_NONE -1       /* do not write an observation */

/* Episode statistics accumulated by the step kernels (all-reduced over ranks by the caller; the
 * fields of the reference drivers' log lines, experiments/main.py:264-311).  The buffer is
 * WURM_STATS_SLOTS x WURM_STATS_FIELDS int64 counters; CTAs add to slot (block index % SLOTS) so no
 * single L2 address is hot; the reader sums over slots. */
#define WURM_STATS_SLOTS 32
#define WURM_STATS_FIELDS 5
#define WURM_STAT_ENV_STEPS 0
#define WURM_STAT_EPISODES 1      /* done flags raised */
#define WURM_STAT_REWARD 2        /* food eaten */
#define WURM_STAT_SELF_COLLISIONS 3
#define WURM_STAT_EDGE_COLLISIONS 4

typedef struct WurmSingleCfg {
    int32_t num_envs; /* N */
    int32_t size;     /* S, >= 9 (single_snake.py:346) */
    int32_t obs_mode; /* WURM_OBS_* */
    int32_t obs_n;    /* n of partial_n */
} WurmSingleCfg;

int wurm_abi_version(void);
const char* wurm_last_error(void);

/* Number of float elements of one env's observation for cfg->obs_mode. */
int64_t wurm_single_obs_elems(const WurmSingleCfg* cfg);

/* Replaces SingleSnake.step (wurm/envs/single_snake.py:197-304), including
 * determine_orientations (wurm/utils.py:36-65), the conv2d head move (wurm/_filters.py:7-28),
 * food respawn (_get_food_addition :306-320 + drop_duplicates wurm/utils.py:205-232) and
 * _observe/_get_rgb (:104-195) in ONE launch.
 *   envs     (N,3,S,S) f32, updated in place          actions (N,) int16/int32/int64 (action_bytes
 *   obs      per cfg->obs_mode, may be NULL with       = 2/4/8) or uint8 (= 1), sanitised IN PLACE like :222
 *            WURM_OBS_NONE                            reward (N,) f32, done/self_col/edge_col (N,) u8
 *   packed   (N,) u8 or NULL: the step's per-env results once more as ONE byte (WURM_PACKED_*), for consumers
 *            on the far side of PCIe -- with uint8 actions a host-side policy moves 2 bytes per env-step
 *            instead of the 13 of int64 actions + fp32 reward + done flag.
 *   food_cell_replay (N,) int32 or NULL: cell index y*S+x of the respawned food for envs that eat
 *            this step (-1: none); NULL -> uniform over free interior cells from Philox.
 *   hints    (N,4) int16 or NULL (initialise to -1): scratch owned by the caller in which the kernels leave each
 *            env's (head cell, snake size, food cell, unused) for their next call.  Pure accelerator: every hint
 *            is verified against `envs` before use -- the head cell must hold a head, the food cell food, the
 *            body channel's maximum and its (size, size-1) cells must be what the hints say -- so stale or garbage
 *            hints (the caller edited `envs`) only cost the work they would have saved: a scan of the body channel
 *            and, for even grid sides from 16 up, reading the food and head channels at all.  (Not re-verified:
 *            the absence of a second head or food cell, i.e. states the reference's env_consistency rejects.)  */
#define WURM_PACKED_DONE 1         /* bit 0: done                                                  */
#define WURM_PACKED_SELF 2         /* bit 1: info['self_collision']                                */
#define WURM_PACKED_EDGE 4         /* bit 2: info['edge_collision']                                */
#define WURM_PACKED_REWARD_SHIFT 3 /* bits 3-4: the reward as an integer 0..3 (:271: 0 or 1 on every
                                      state the reference's env_consistency accepts)               */
int wurm_single_step(const WurmSingleCfg* cfg, float* envs, void* actions, int action_bytes,
                     const int32_t* food_cell_replay, uint64_t seed, uint64_t step, const uint64_t* step_dev,
                     float* obs, float* reward,
                     uint8_t* done, uint8_t* self_col, uint8_t* edge_col, int32_t* status,
                     int64_t* stats /* nullable */, int16_t* hints /* nullable */, uint8_t* packed /* nullable */,
                     void* stream);

/* Fused fast path: wurm_single_step followed by wurm_single_reset(done) in ONE launch -- the pair the
 * reference's driver issues every iteration (experiments/main.py:212-227).  Outputs are those of the
 * step (the observation of an env that just ended is its terminal one, which is what main.py feeds
 * its policy next, :227); the state left in `envs` is the one after the reset.  Draws: the step's with
 * call counter `step`, the reset's with `step` + 1, i.e. bit-identical to the two calls in sequence.
 *   spawn_replay (N,4) int32 or NULL: as for wurm_single_reset. */
int wurm_single_step_reset(const WurmSingleCfg* cfg, float* envs, void* actions, int action_bytes,
                           const int32_t* food_cell_replay, const int32_t* spawn_replay, uint64_t seed, uint64_t step,
                           const uint64_t* step_dev, float* obs, float* reward, uint8_t* done, uint8_t* self_col,
                           uint8_t* edge_col, int32_t* status, int64_t* stats /* nullable */, int16_t* hints /* nullable */,
                           uint8_t* packed /* nullable */, void* stream);

/* Replaces the state update of SingleSnake.reset / _create_envs (single_snake.py:322-337, 344-387):
 * envs whose done_mask byte is non-zero are re-created, all others untouched.
 *   spawn_replay (N,4) int32 or NULL: rows (y, x, dir, food_cell), read for done envs only.     */
int wurm_single_reset(const WurmSingleCfg* cfg, float* envs, const uint8_t* done_mask, const int32_t* spawn_replay,
                      uint64_t seed, uint64_t step, const uint64_t* step_dev, int16_t* hints /* nullable */, void* stream);

/* Replaces SingleSnake._observe (single_snake.py:130-195) on the current state. */
int wurm_single_observe(const WurmSingleCfg* cfg, const float* envs, float* obs, int32_t* status, void* stream);

/* ---- SingleSnake with a COMPACT RESIDENT STATE (optional; wurm_b200/csrc/single_compact.cu) ----
 * Instead of the reference's (N,3,S,S) fp32 tensor the env lives in HBM as
 *   cells (N, Cp) uint16, Cp = S*S rounded up to a multiple of 8, 16-byte aligned: bits 0-13 body value, bit 14 head,
 *         bit 15 food;
 *   aux   (N, 4) int16, 8-byte aligned: head cell (-1 none), snake size, neck cell, flags -- derived data the kernels
 *         maintain (wurm_single_compact computes it from scratch);
 * 176 bytes per env at size 9 instead of 972.  Same semantics, same draws, same outputs as the entry points above
 * (bit-identical on every state the records can carry: integral body values < 16384, food / head values 0 or 1; size
 * <= 90); wurm_single_compact / wurm_single_expand convert, the former raising WURM_ST_NOT_COMPACT for anything else. */
int wurm_single_compact_step(const WurmSingleCfg* cfg, uint16_t* cells, int16_t* aux, void* actions, int action_bytes,
                             const int32_t* food_cell_replay, int auto_reset /* fuse reset(done) into the launch */,
                             const int32_t* spawn_replay, uint64_t seed, uint64_t step, const uint64_t* step_dev, float* obs,
                             float* reward, uint8_t* done, uint8_t* self_col, uint8_t* edge_col, int32_t* status,
                             int64_t* stats /* nullable */, uint8_t* packed /* nullable */, void* stream);
int wurm_single_compact_reset(const WurmSingleCfg* cfg, uint16_t* cells, int16_t* aux, const uint8_t* done_mask,
                              const int32_t* spawn_replay, uint64_t seed, uint64_t step, const uint64_t* step_dev, void* stream);
int wurm_single_compact_observe(const WurmSingleCfg* cfg, const uint16_t* cells, const int16_t* aux, float* obs,
                                int32_t* status, void* stream);
/* wurm_single_check (env_consistency, wurm/utils.py:113-178) on records: same report, same WURM_CHK_* bits. */
int wurm_single_compact_check(const WurmSingleCfg* cfg, const uint16_t* cells, const uint8_t* skip /* nullable */, int32_t* report,
                              void* stream);
int wurm_single_compact(const WurmSingleCfg* cfg, const float* envs, uint16_t* cells, int16_t* aux, int32_t* status, void* stream);
int wurm_single_expand(const WurmSingleCfg* cfg, const uint16_t* cells, const int16_t* aux, float* envs, void* stream);

/* Invariant checks (wurm/utils.py:113-178 snake_consistency + env_consistency, and
 * MultiSnake.check_consistency multi_snake.py:733-769) fused into ONE reduction kernel per call.  The
 * reference's drivers run these every step (experiments/main.py:215, experiments/multiagent.py:378)
 * as ~10 full-state reductions with a host sync each.  report: device int32[WURM_CHECK_REPORT]:
 *   [0] OR over the checked envs of the WURM_CHK_* bits (in the order the reference tests them)
 *   [1] number of violating envs (agents for MultiSnake)      [2] smallest violating env index  */
#define WURM_CHECK_REPORT 4
#define WURM_CHK_FOOD_VALUE 1      /* 'An environment has an invalid food pixel'                          utils.py:119-125 */
#define WURM_CHK_HEAD_COUNT 2      /* '...multiple num_heads for a single snake.'                         :127-131 */
#define WURM_CHK_NO_SNAKE 4        /* "environments don't contain a snake."                               :134-136 */
#define WURM_CHK_HEAD_NOT_AT_END 8 /* "...it's head not at the end of the body."                          :139-143 */
#define WURM_CHK_BODY_VALUES 16    /* '...a body with inconsistent values i.e. not range(n)'              :147-153 */
#define WURM_CHK_TOO_SHORT 32      /* 'A snake has size of less than 3.'                                  :156-157 */
#define WURM_CHK_HEAD_ON_FOOD 64   /* 'A food and head pixel is overlapping...'                           :160-164 */
#define WURM_CHK_FOOD_COUNT 128    /* "...doesn't contain exactly one food instance" (SingleSnake only)   :176-178 */
#define WURM_CHK_OVERLAP 256       /* 'An environment contains overlapping snakes'                 multi_snake.py:746-758 */
#define WURM_CHK_DEAD_NOT_ZERO 512 /* 'Dead snake contains non-zero elements.'                     multi_snake.py:766-769 */

/* skip (N) nullable: envs whose byte is non-zero are not checked (the driver checks envs[~done]). */
int wurm_single_check(const WurmSingleCfg* cfg, const float* envs, const uint8_t* skip, int32_t* report, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MultiSnake (wurm/envs/multi_snake.py)
 * ---------------------------------------------------------------------------------------------- */
#define WURM_MULTI_MAX_SNAKES 32
#define WURM_MOBS_NONE -1
#define WURM_MOBS_FULL 0    /* per agent (E,3,S,S): self green, others blue   (multi_snake.py:268-288) */
#define WURM_MOBS_PARTIAL 1 /* per agent (E,3,W,W) crop of the env image      (multi_snake.py:289-332) */

/* Constructor parameters of the reference that the kernels read (multi_snake.py:56-75, 121-129);
 * the Python class keeps them as mutable attributes and rebuilds this struct at every call, because
 * the reference's drivers anneal them between steps (experiments/multiagent.py:337-345). */
typedef struct WurmMultiCfg {
    int32_t num_envs;      /* E */
    int32_t num_snakes;    /* K <= WURM_MULTI_MAX_SNAKES */
    int32_t size;          /* S */
    int32_t obs_mode;      /* WURM_MOBS_* */
    int32_t obs_n;         /* n of partial_n */
    int32_t boost;         /* self.boost */
    int32_t food_on_death; /* food_on_death_prob > 0 */
    float death_threshold; /* float32(1 - food_on_death_prob)          :424 */
    float boost_cost_prob; /* float32(boost_cost_prob)                 :579 */
    int32_t food_mode;     /* 0 'only_one', 1 'random_rate'            :369,380 */
    float food_rate;       /* float32(food_rate)                       :403 */
    float reward_on_death; /*                                          :684 */
    int32_t respawn_any;   /* respawn_mode == 'any'                    :805 */
    int32_t colour_random; /* colour_mode == 'random'                  :800 */
} WurmMultiCfg;

/* The state tensors of the reference (multi_snake.py:100-108,145), all device pointers. */
typedef struct WurmMultiState {
    float* foods;           /* (E,1,S,S)   */
    float* heads;           /* (E*K,1,S,S) */
    float* bodies;          /* (E*K,1,S,S) */
    uint8_t* dones;         /* (E*K)       */
    int64_t* orientations;  /* (E*K)       */
    uint8_t* boost_this_step; /* (E*K)     */
    int16_t* agent_colours; /* (E*K,3)     */
    /* Optional (E*K) scratch owned by the library between calls, NULL to disable: the head cell each snake was left on
     * by the previous call (-1 dead, -2 unknown; initialise to -2).  The heads tensor is 44-48 % of the state bytes and
     * holds one non-zero per snake: a call that finds every live snake's hinted cell still holding a head (and every
     * dead snake still flagged done) skips streaming it.  Hints are verified, never trusted: after the caller edits the
     * tensors they only cost the scan back.  (What is not re-verified is the absence of a SECOND head of a snake: such
     * states are outside the supported set, as for the reference's own check_consistency.) */
    int16_t* head_hints;
    /* COMPACT RESIDENT STATE (optional, NULL = the reference's tensors above are the state).  (E, Cp) uint32 records,
     * Cp = S*S rounded up to a multiple of 4, 16-byte aligned: bits 0-15 body value, bits 16-21 owner snake + 1, bit 29
     * food, all other bits zero; the snakes' head cells live in head_hints (authoritative in this mode: -1 = no head).
     * When set, every entry point below reads and writes the records INSTEAD of foods / heads / bodies (which may be
     * NULL): 16 KB per env at K=16, S=64 instead of 541 KB of fp32 that is 99 % zeros.  wurm_multi_compact /
     * wurm_multi_expand convert between the two forms; results are bit-identical to the dense path on every state
     * the records can carry (integral body values < 65536, food / head values 1, one body per cell, heads on their
     * own bodies), which includes every state the kernels themselves produce from such a state. */
    uint32_t* cells;
    /* With BOTH forms present (tensors and cells): cells_valid != 0 says the records describe the tensors' current content
     * and nothing else was written into the tensors since the library last wrote them -- a step then loads the records,
     * VERIFIES that every cell they name still holds that value in the tensors (and every head cell its head; a mismatch
     * re-loads the env from the tensors), and writes its changes to both forms: the dense state at close to the compact
     * state's cost.  cells_valid == 0: the step loads the tensors as usual and EMITS the records, after which the caller
     * may pass 1.  Ignored (taken as 1) when the tensors are NULL.  What the presence check cannot see -- values ADDED to
     * the tensors since the library last wrote them -- is the caller's word: pass 1 only if nothing else wrote to the tensors.
     * Per env, a head hint of -2 ("unknown") has the same effect as cells_valid == 0 for that env, so a caller replaying a
     * captured launch (CUDA graph) can withdraw the records by filling head_hints with -2.  wurm_multi_reset follows the same
     * flag (records + tensors when valid, tensors only otherwise); wurm_multi_observe / _env_images / _check read the records
     * when valid. */
    int32_t cells_valid;
} WurmMultiState;

/* Replayed random draws of one step, dense per env (NULL struct pointer -> Philox).
 * SURVEY.md Appendix B.2 gives the reference's schedule. */
typedef struct WurmMultiStepDraws {
    int32_t boost_phase_ran; /* the reference ran its boost phase (batch-global condition :503) */
    const float* u_boost;    /* (E,S,S) rand_like :424 via :574, NULL if not drawn */
    const float* u_cost;     /* (E*K)   rand :579,               NULL if not drawn */
    const float* u_reg;      /* (E,S,S) rand_like :424 via :671, NULL if not drawn */
    const int32_t* food_cell;/* (E)     only_one: respawned cell or -1             */
    const float* u_rate;     /* (E,S,S) random_rate: rand :401 scattered to env rows */
} WurmMultiStepDraws;

/* Per-step outputs, (E,K) row-major unless noted: column k is the reference's dict entry of agent k. */
typedef struct WurmMultiStepOut {
    float* rewards;
    uint8_t* snake_collision;
    uint8_t* edge_collision;
    float* food;            /* info food_k */
    float* size;            /* info size_k */
    uint8_t* dones;         /* copy of the updated done flags (the caller's to keep) */
    uint8_t* boost;         /* info boost_k: copy of boost_this_step */
    uint8_t* all_done;      /* (E) dones['__all__'] */
    float* obs;             /* (K,E,3,S,S) or (K,E,3,W,W): obs[k] is agent k's tensor; NULL with WURM_MOBS_NONE */
} WurmMultiStepOut;

/* Replayed draws of one reset (NULL struct pointer -> Philox). */
typedef struct WurmMultiResetDraws {
    const int32_t* create;  /* (E,K+1,2): per snake (seed cell, direction), then (food cell, 0); re-created envs */
    const int32_t* respawn; /* (E,2): (seed cell or -1, direction) for the env's first dead agent */
    const int16_t* colours; /* (E*K,3): new colour of every agent that is still dead */
} WurmMultiResetDraws;

int64_t wurm_multi_obs_elems(const WurmMultiCfg* cfg); /* floats per (agent, env) */

/* Replaces MultiSnake.step (multi_snake.py:462-731) incl. _move_heads, _get_food_overlap,
 * _decay_bodies, _check_collisions, _check_edges, _food_from_death, _add_food, _get_env_images and
 * _observe, in ONE launch.  actions: host array of K device pointers, each (E,) of action_bytes (8/4/2: int64/int32/
 * int16 as in the reference; 1: uint8) integers in [0,8) (agents in dict order). */
int wurm_multi_step(const WurmMultiCfg* cfg, const WurmMultiState* state, const void* const* actions, int action_bytes,
                    const WurmMultiStepDraws* draws, uint64_t seed, uint64_t step, const uint64_t* step_dev,
                    const WurmMultiStepOut* out, int32_t* status, int64_t* stats /* nullable */, void* stream);

/* Fused fast path: wurm_multi_step followed by wurm_multi_reset(dones['__all__']) in ONE launch -- the pair the
 * reference's driver issues every iteration (experiments/multiagent.py:375-377).  Outputs (and the
 * observations) are those of the step; the state left behind is the one after the reset.  The reset's
 * draws use call counter `step` + 1: bit-identical to the two calls in sequence.  reset_draws: replay tape of
 * the reset or NULL. */
int wurm_multi_step_reset(const WurmMultiCfg* cfg, const WurmMultiState* state, const void* const* actions, int action_bytes,
                          const WurmMultiStepDraws* draws, const WurmMultiResetDraws* reset_draws, uint64_t seed,
                          uint64_t step, const uint64_t* step_dev, const WurmMultiStepOut* out, int32_t* status,
                          int64_t* stats /* nullable */, void* stream);

/* Replaces the state update of MultiSnake.reset (multi_snake.py:771-831) incl. _create_envs,
 * _add_snake, _get_snake_addition and get_n_colours.  env_done (E): envs to re-create. */
int wurm_multi_reset(const WurmMultiCfg* cfg, const WurmMultiState* state, const uint8_t* env_done,
                     const WurmMultiResetDraws* draws, uint64_t seed, uint64_t step, const uint64_t* step_dev,
                     int32_t* status, void* stream);

/* fp32 tensors -> compact records (+ head cells); raises WURM_ST_NOT_COMPACT in *status for a state the records cannot
 * carry.  `state` must hold both forms. */
int wurm_multi_compact(const WurmMultiCfg* cfg, const WurmMultiState* state, int32_t* status, void* stream);
/* compact records (+ head cells) -> the reference's fp32 tensors, dense. */
int wurm_multi_expand(const WurmMultiCfg* cfg, const WurmMultiState* state, void* stream);

/* Replaces MultiSnake._observe (multi_snake.py:283-334) on the current state. */
int wurm_multi_observe(const WurmMultiCfg* cfg, const WurmMultiState* state, float* obs, int32_t* status, void* stream);

/* Replaces MultiSnake._get_env_images (multi_snake.py:194-227): img (E,3,S,S) int16. */
int wurm_multi_env_images(const WurmMultiCfg* cfg, const WurmMultiState* state, int16_t* img, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SimpleGridworld (wurm/envs/simple_gridworld.py): the reference's two-channel debug env, selectable
 * from its driver (experiments/main.py:166-168).  envs (N,2,S,S) f32: food, agent.
 * ---------------------------------------------------------------------------------------------- */
typedef struct WurmGridCfg {
    int32_t num_envs; /* N */
    int32_t size;     /* S > 4 (simple_gridworld.py:239) */
    int32_t obs_mode; /* WURM_OBS_DEFAULT (N,3,S,S) | WURM_OBS_RAW (N,2,S,S) | WURM_OBS_POSITIONS (N,4) | WURM_OBS_NONE */
    int32_t start_y;  /* start_location of the agent (reset only) */
    int32_t start_x;
} WurmGridCfg;

/* Replaces SimpleGridworld.step (simple_gridworld.py:135-201) incl. _get_food_addition (:208-220) and
 * _observe (:88-133) in one launch.  Actions are not sanitised (there is no orientation). */
int wurm_grid_step(const WurmGridCfg* cfg, float* envs, const void* actions, int action_bytes,
                   const int32_t* food_cell_replay, uint64_t seed, uint64_t step, const uint64_t* step_dev, float* obs,
                   float* reward, uint8_t* done, int32_t* status, int64_t* stats /* nullable */, void* stream);

/* Replaces the state update of SimpleGridworld.reset / _create_envs (simple_gridworld.py:222-262). */
int wurm_grid_reset(const WurmGridCfg* cfg, float* envs, const uint8_t* done_mask, const int32_t* food_cell_replay,
                    uint64_t seed, uint64_t step, const uint64_t* step_dev, void* stream);

/* Replaces SimpleGridworld._observe (simple_gridworld.py:108-133). */
int wurm_grid_observe(const WurmGridCfg* cfg, const float* envs, float* obs, void* stream);

/* A2C return scan (wurm/rl/a2c.py:49-63): the backward recurrence over a (T,N) trajectory
 * in one launch, fp32 in the reference's operation order.  rewards / values / dones / returns (T,N) row-major, bootstrap
 * (N).  gae_lambda < 0: n-step returns R_t = r_t + gamma R_{t+1} (1 - d_t) (:58-61); otherwise GAE (:50-57).  gamma and
 * gae_lambda are DOUBLES, as in the reference (Python floats): :56 multiplies them in double before the product meets fp32. */
int wurm_a2c_returns(int32_t num_steps, int32_t num_envs, double gamma, double gae_lambda, const float* bootstrap,
                     const float* rewards, const float* values, const uint8_t* dones, float* returns, void* stream);

/* MultiSnake.check_consistency (multi_snake.py:733-769): living snakes against snake_consistency,
 * no two bodies on one cell, dead snakes all-zero.  report as for wurm_single_check (counts agents). */
int wurm_multi_check(const WurmMultiCfg* cfg, const WurmMultiState* state, int32_t* report, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WURM_B200_H */
