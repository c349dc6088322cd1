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
 */
#ifndef WURM_B200_H
#define WURM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WURM_ABI_VERSION 1

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
 *   obs      per cfg->obs_mode, may be NULL with       = 2/4/8), sanitised IN PLACE like :222
 *            WURM_OBS_NONE                            reward (N,) f32, done/self_col/edge_col (N,) u8
 *   food_cell_replay (N,) int32 or NULL: cell index y*S+x of the respawned food for envs that eat
 *            this step (-1: none); NULL -> uniform over free interior cells from Philox.        */
int wurm_single_step(const WurmSingleCfg* cfg, float* envs, void* actions, int action_bytes,
                     const int32_t* food_cell_replay, uint64_t seed, uint64_t step, float* obs, float* reward,
                     uint8_t* done, uint8_t* self_col, uint8_t* edge_col, int32_t* status,
                     int64_t* stats /* nullable */, void* stream);

/* Replaces the state update of SingleSnake.reset / _create_envs (single_snake.py:322-337, 344-387):
 * envs whose done_mask byte is non-zero are re-created, all others untouched.
 *   spawn_replay (N,4) int32 or NULL: rows (y, x, dir, food_cell), read for done envs only.     */
int wurm_single_reset(const WurmSingleCfg* cfg, float* envs, const uint8_t* done_mask, const int32_t* spawn_replay,
                      uint64_t seed, uint64_t step, void* stream);

/* Replaces SingleSnake._observe (single_snake.py:130-195) on the current state. */
int wurm_single_observe(const WurmSingleCfg* cfg, const float* envs, float* obs, int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WURM_B200_H */
