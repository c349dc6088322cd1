// SimpleGridworld (wurm/envs/simple_gridworld.py) for B200 (sm_100a): the reference's two-channel debug
// env -- an agent pixel and one food pixel, the same move / eat / respawn / edge machinery as SingleSnake
// without a body.  An env is 2*S*S floats (392 B at the reference's size 7).  Up to 64 cells: tiles of eight envs moved
// by TMA bulk copies and stepped in shared memory (grid_tile_kernel, round 2), or -- where the env stride would line
// the tile up on one shared-memory bank -- 8 lanes per env with the env held in registers (grid_small_kernel); larger
// grids (grid_env_kernel): 32 lanes per env, strided loads, the agent found with shuffles.  Either way only the two or
// three cells a step changes are stored.
#include <math.h>
#include <stdlib.h>

#include "../../include/wurm_b200.h"
#include "common.cuh"
#include "host_util.h"

namespace wurm {

enum : uint32_t { kStreamGridStepFood = 11, kStreamGridReset = 12 };

struct GridParams {
    float* envs;
    const void* actions;
    const int32_t* food_replay;
    float* obs;
    float* reward;
    uint8_t* done;
    int32_t* status;
    unsigned long long* stats;
    const unsigned long long* step_dev;
    uint64_t seed, step;
    int N, S, C, action_bytes, obs_mode, start_cell;
    uint32_t magic_S;
};

__device__ __forceinline__ uint64_t call_counter(const GridParams& p) { return p.step + (p.step_dev ? *p.step_dev : 0ull); }

// uniform free interior cell (food + agent < EPS, simple_gridworld.py:208-216): rejection sampling by the
// whole warp in lock step, explicit ranking after kRejectionTries misses.  `env` is read through L2.
__device__ __forceinline__ int grid_pick_free(const GridParams& p, const float* env, uint64_t ctr, int e, uint32_t stream) {
    const int S = p.S, C = p.C, I = S - 2;
    auto is_free = [&](int q) { return __ldcg(env + q) + __ldcg(env + C + q) < kEps; };
    for (uint32_t t = 0; t < kRejectionTries; ++t) {
        const int cand = (int)bounded(draw_i(p.seed, ctr, (uint32_t)e, stream, t), (uint32_t)(I * I));
        const int cy = cand / I, q = (1 + cy) * S + 1 + (cand - cy * I);
        if (is_free(q)) return q;
    }
    int nfree = 0;
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x) nfree += is_free(y * S + x);
    if (nfree == 0) return -1;
    int r = (int)bounded(draw_i(p.seed, ctr, (uint32_t)e, stream, kRejectionTries), (uint32_t)nfree);
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x)
            if (is_free(y * S + x) && r-- == 0) return y * S + x;
    return -1;
}

// simple_gridworld.py:88-133 for env e by a group of G lanes (reads through L2: the step's own stores are visible)
template <int G>
__device__ __forceinline__ void grid_observe_env(const GridParams& p, int e, int lane) {
    const unsigned gm = group_mask<G>();
    const int S = p.S, C = p.C;
    const float* env = p.envs + (size_t)e * 2 * C;
    if (p.obs_mode == WURM_OBS_DEFAULT) {                            // black background (:90), agent green, food red
        float* o = p.obs + (size_t)e * 3 * C;
        for (int q = lane; q < C; q += G) {
            const int y = (int)__umulhi((uint32_t)q, p.magic_S), x = q - y * S;
            float r = 0.0f, g = 0.0f;
            if (__ldcg(env + C + q) > kEps) { r = 0.0f; g = 1.0f; }
            if (__ldcg(env + q) > kEps) { r = 1.0f; g = 0.0f; }
            if (y == 0 || x == 0 || y == S - 1 || x == S - 1) r = g = 0.0f;
            o[q] = r; o[C + q] = g; o[2 * C + q] = 0.0f;
        }
    } else if (p.obs_mode == WURM_OBS_RAW) {
        for (int i = lane; i < 2 * C; i += G) p.obs[(size_t)e * 2 * C + i] = __ldcg(env + i);
    } else if (p.obs_mode == WURM_OBS_POSITIONS) {                   // first argmax of agent / food (:119-130)
        int idx[2];
        for (int ch = 0; ch < 2; ++ch) {
            const float* v = env + (ch == 0 ? C : 0);
            float bv = -INFINITY;
            int bq = 0;
            for (int q = lane; q < C; q += G) {
                const float x = __ldcg(v + q);
                if (x > bv) { bv = x; bq = q; }
            }
            const float gv = group_max<G>(bv, gm);
            idx[ch] = -group_max<G>(bv == gv ? -bq : -(1 << 30), gm);
        }
        if (lane == 0) {
            float* o = p.obs + (size_t)e * 4;
            o[0] = (float)(idx[0] / S); o[1] = (float)(idx[0] % S); o[2] = (float)(idx[1] / S); o[3] = (float)(idx[1] % S);
        }
    }
}

// simple_gridworld.py:135-201: G consecutive lanes per env (G = 8 at the reference's size 7: four envs per warp)
template <int G, bool STEP>
__global__ void __launch_bounds__(256) grid_env_kernel(const GridParams p) {
    __shared__ int cnt_s[4];                                          // per-CTA episode statistics
    const unsigned gm = group_mask<G>();
    const int lane = threadIdx.x % G;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool active = e < p.N;
    if (STEP && p.stats) {
        if (threadIdx.x < 4) cnt_s[threadIdx.x] = 0;
        __syncthreads();
    }
    if (STEP && active) {
        const int S = p.S, C = p.C;
        float* food = p.envs + (size_t)e * 2 * C;
        float* head = food + C;
        int hp = -1, hc = 0;
        for (int q = lane; q < C; q += G)
            if (head[q] != 0.0f) { hp = q; ++hc; }
        hp = group_max<G>(hp, gm);
        hc = group_sum<G>(hc, gm);
        long long a;
        a = load_action(p.actions, p.action_bytes, (size_t)e);
        int np = -1, ny = -1, nx = -1;                                // :149-158 the conv2d translates the agent by -OFF[a]
        if (hp >= 0) {
            const int hy = (int)__umulhi((uint32_t)hp, p.magic_S), hx = hp - hy * S;
            ny = hy - off_y((int)a); nx = hx - off_x((int)a);
            if (ny >= 0 && ny < S && nx >= 0 && nx < S) np = ny * S + nx;
        }
        const float ov = np >= 0 ? food[np] : 0.0f;                   // :169 agent-food overlap
        __syncwarp(gm);
        if (lane == 0 && hp >= 0) {
            head[hp] = 0.0f;
            if (np >= 0) {
                head[np] = 1.0f;
                if (ov != 0.0f) food[np] = ov + ov * -1.0f;           // :171
            }
        }
        __syncwarp(gm);
        if (ov != 0.0f) {                                             // :176-181 respawn
            const int cell = p.food_replay ? p.food_replay[e]
                                           : grid_pick_free(p, food, call_counter(p), e, kStreamGridStepFood);
            if (lane == 0 && cell >= 0) food[cell] = __ldcg(food + cell) + 1.0f;
        }
        const bool interior = np >= 0 && ny >= 1 && ny <= S - 2 && nx >= 1 && nx <= S - 2;
        if (lane == 0) {
            p.reward[e] = 0.0f - ov * -1.0f;
            p.done[e] = !interior;                                    // :189-194 edge collision is the only way to end
            if (hc > 1) atomicOr(p.status, WURM_ST_MULTI_HEAD);
            if (p.stats) {
                atomicAdd(&cnt_s[0], 1);
                if (!interior) atomicAdd(&cnt_s[1], 1);
                if (ov != 0.0f) atomicAdd(&cnt_s[2], 1);
            }
        }
        __syncwarp(gm);
    }
    if (active && p.obs_mode >= 0) grid_observe_env<G>(p, e, lane);
    if (STEP && p.stats) {                                            // one striped slot per CTA: no hot address in L2
        __syncthreads();
        unsigned long long* slot = p.stats + (blockIdx.x % WURM_STATS_SLOTS) * WURM_STATS_FIELDS;
        if (threadIdx.x == 0 && cnt_s[0]) atomicAdd(slot + WURM_STAT_ENV_STEPS, (unsigned long long)cnt_s[0]);
        if (threadIdx.x == 1 && cnt_s[1]) {
            atomicAdd(slot + WURM_STAT_EPISODES, (unsigned long long)cnt_s[1]);
            atomicAdd(slot + WURM_STAT_EDGE_COLLISIONS, (unsigned long long)cnt_s[1]);
        }
        if (threadIdx.x == 2 && cnt_s[2]) atomicAdd(slot + WURM_STAT_REWARD, (unsigned long long)cnt_s[2]);
    }
}

// Envs of at most 64 cells (the reference's size 7: 392 B of state, 588 B of observation): 8 lanes per env,
// each lane keeps its <= 8 cells of both channels in REGISTERS -- all loads of the env are issued at once (one
// DRAM round trip), the step edits registers and stores only the two or three cells it changes, and the
// observation is rendered from the registers instead of re-reading the env through L2.
template <bool STEP>
// 128 threads, >= 8 CTAs per SM (<= 64 registers): best of the measured launch shapes (profiles/r01_sweep_grid.txt)
#ifndef WURM_GRID_MINB
#define WURM_GRID_MINB 8
#endif
#ifndef WURM_GRID_THREADS
#define WURM_GRID_THREADS 128
#endif
__global__ void __launch_bounds__(WURM_GRID_THREADS, WURM_GRID_MINB) grid_small_kernel(const GridParams p) {
    constexpr int G = 8, R = 8;
    __shared__ int cnt_s[4];                                          // per-CTA episode statistics
    const unsigned gm = group_mask<G>();
    const int lane = threadIdx.x % G;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool active = e < p.N;
    const int S = p.S, C = p.C;
    if (STEP && p.stats && threadIdx.x < 4) cnt_s[threadIdx.x] = 0;
    float f[R], h[R];
    long long a = 0;
    float* food = p.envs + (size_t)(active ? e : 0) * 2 * C;
    float* head = food + C;
#pragma unroll
    for (int it = 0; it < R; ++it) {
        const int q = lane + G * it;
        const bool in = active && q < C;
        f[it] = in ? food[q] : 0.0f;
        h[it] = in ? head[q] : 0.0f;
    }
    if (STEP && active) {
        a = load_action(p.actions, p.action_bytes, (size_t)e);
    }
    if (STEP && p.stats) __syncthreads();
    if (STEP && active) {
        int hp = -1, hc = 0;
#pragma unroll
        for (int it = 0; it < R; ++it)
            if (h[it] != 0.0f) { hp = lane + G * it; ++hc; }
        hp = group_max<G>(hp, gm);
        hc = group_sum<G>(hc, gm);
        int np = -1, ny = -1, nx = -1;                                // :149-158 the conv2d translates the agent by -OFF[a]
        if (hp >= 0) {
            const int hy = (int)__umulhi((uint32_t)hp, p.magic_S), hx = hp - hy * S;
            ny = hy - off_y((int)a); nx = hx - off_x((int)a);
            if (ny >= 0 && ny < S && nx >= 0 && nx < S) np = ny * S + nx;
        }
        float ov = 0.0f;                                              // :169 agent-food overlap: food[np], owned by one lane
#pragma unroll
        for (int it = 0; it < R; ++it)
            if (lane + G * it == np) ov = f[it];
        ov = group_sum<G>(ov, gm);                                    // one non-zero term at most: exact
#pragma unroll
        for (int it = 0; it < R; ++it) {
            const int q = lane + G * it;
            if (q == hp) { h[it] = 0.0f; head[q] = 0.0f; }
        }
        __syncwarp(gm);                                               // hp == np cannot happen, but keep the stores ordered
#pragma unroll
        for (int it = 0; it < R; ++it) {
            const int q = lane + G * it;
            if (q == np) {
                h[it] = 1.0f; head[q] = 1.0f;
                if (ov != 0.0f) { f[it] = ov + ov * -1.0f; food[q] = f[it]; }   // :171
            }
        }
        if (ov != 0.0f) {                                             // :176-181 respawn (reads the env through L2)
            __syncwarp(gm);
            const int cell = p.food_replay ? p.food_replay[e]
                                           : grid_pick_free(p, food, call_counter(p), e, kStreamGridStepFood);
#pragma unroll
            for (int it = 0; it < R; ++it) {
                const int q = lane + G * it;
                if (q == cell) { f[it] += 1.0f; food[q] = f[it]; }
            }
        }
        const bool interior = np >= 0 && ny >= 1 && ny <= S - 2 && nx >= 1 && nx <= S - 2;
        if (lane == 0) {
            p.reward[e] = 0.0f - ov * -1.0f;
            p.done[e] = !interior;                                    // :189-194 edge collision is the only way to end
            if (hc > 1) atomicOr(p.status, WURM_ST_MULTI_HEAD);
            if (p.stats) {
                atomicAdd(&cnt_s[0], 1);
                if (!interior) atomicAdd(&cnt_s[1], 1);
                if (ov != 0.0f) atomicAdd(&cnt_s[2], 1);
            }
        }
    }
    if (active && p.obs_mode == WURM_OBS_DEFAULT) {                   // :88-133 black background (:90), agent green, food red
        float* o = p.obs + (size_t)e * 3 * C;
#pragma unroll
        for (int it = 0; it < R; ++it) {
            const int q = lane + G * it;
            if (q < C) {
                const int y = (int)__umulhi((uint32_t)q, p.magic_S), x = q - y * S;
                float r = 0.0f, g = 0.0f;
                if (h[it] > kEps) { r = 0.0f; g = 1.0f; }
                if (f[it] > kEps) { r = 1.0f; g = 0.0f; }
                if (y == 0 || x == 0 || y == S - 1 || x == S - 1) r = g = 0.0f;
                o[q] = r; o[C + q] = g; o[2 * C + q] = 0.0f;
            }
        }
    } else if (active && p.obs_mode == WURM_OBS_RAW) {
        float* o = p.obs + (size_t)e * 2 * C;
#pragma unroll
        for (int it = 0; it < R; ++it) {
            const int q = lane + G * it;
            if (q < C) { o[q] = f[it]; o[C + q] = h[it]; }
        }
    } else if (active && p.obs_mode == WURM_OBS_POSITIONS) {          // first argmax of agent / food (:119-130)
        int idx[2];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float bv = -INFINITY;
            int bq = 0;
#pragma unroll
            for (int it = 0; it < R; ++it) {
                const int q = lane + G * it;
                const float x = ch == 0 ? h[it] : f[it];
                if (q < C && x > bv) { bv = x; bq = q; }
            }
            const float gv = group_max<G>(bv, gm);
            idx[ch] = -group_max<G>(bv == gv ? -bq : -(1 << 30), gm);
        }
        if (lane == 0) {
            float* o = p.obs + (size_t)e * 4;
            o[0] = (float)(idx[0] / S); o[1] = (float)(idx[0] % S); o[2] = (float)(idx[1] / S); o[3] = (float)(idx[1] % S);
        }
    }
    if (STEP && p.stats) {                                            // one striped slot per CTA: no hot address in L2
        __syncthreads();
        unsigned long long* slot = p.stats + (blockIdx.x % WURM_STATS_SLOTS) * WURM_STATS_FIELDS;
        if (threadIdx.x == 0 && cnt_s[0]) atomicAdd(slot + WURM_STAT_ENV_STEPS, (unsigned long long)cnt_s[0]);
        if (threadIdx.x == 1 && cnt_s[1]) {
            atomicAdd(slot + WURM_STAT_EPISODES, (unsigned long long)cnt_s[1]);
            atomicAdd(slot + WURM_STAT_EDGE_COLLISIONS, (unsigned long long)cnt_s[1]);
        }
        if (threadIdx.x == 2 && cnt_s[2]) atomicAdd(slot + WURM_STAT_REWARD, (unsigned long long)cnt_s[2]);
    }
}

// ---------------------------------------------------------------------------------------------
// Small grids as TILES (round 2): the SingleSnake tile design applied to the gridworld.  One one-warp CTA = T = 32 / G
// consecutive envs; the tile's (T,2,S,S) floats arrive by ONE bulk copy (TMA), are stepped in shared memory with the two
// or three changed cells mirrored to HBM, and the observation is rendered into a shared staging area that leaves by one
// bulk store -- instead of 8 lanes per env pulling 32-byte pieces of four different envs per load instruction and
// scattering 32-byte pieces of the observation.  Same arithmetic, same draws.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int grid_pick_free_tile(const GridParams& p, const float* env, uint64_t ctr, int e) {
    const int S = p.S, C = p.C, I = S - 2;
    auto is_free = [&](int q) { return env[q] + env[C + q] < kEps; };
    for (uint32_t t = 0; t < kRejectionTries; ++t) {
        const int cand = (int)bounded(draw_i(p.seed, ctr, (uint32_t)e, kStreamGridStepFood, t), (uint32_t)(I * I));
        const int cy = cand / I, q = (1 + cy) * S + 1 + (cand - cy * I);
        if (is_free(q)) return q;
    }
    int nfree = 0;
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x) nfree += is_free(y * S + x);
    if (nfree == 0) return -1;
    int r = (int)bounded(draw_i(p.seed, ctr, (uint32_t)e, kStreamGridStepFood, kRejectionTries), (uint32_t)nfree);
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x)
            if (is_free(y * S + x) && r-- == 0) return y * S + x;
    return -1;
}

template <int G, bool STEP>
__global__ void __launch_bounds__(32) grid_tile_kernel(const GridParams p, int tile_bytes_padded, int stage_env_floats) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int T = 32 / G;
    const int S = p.S, C = p.C, n2 = 2 * C;
    float* tile = reinterpret_cast<float*>(smem);
    float* stage = reinterpret_cast<float*>(smem + tile_bytes_padded);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + tile_bytes_padded + ((T * stage_env_floats * 4 + 15) & ~15));
    int* cnt_s = reinterpret_cast<int*>(bar + 1);
    const unsigned gm = group_mask<G>();
    const int lane = threadIdx.x, t = lane / G, l = lane % G;
    const int env0 = blockIdx.x * T, nvalid = min(T, p.N - env0);
    const size_t goff = (size_t)env0 * n2;
    const uint32_t bytes = (uint32_t)(nvalid * n2) * 4u;
    const bool bulk = (bytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(p.envs + goff) & 15u) == 0);
    if (bulk) {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_expect_tx(bar, bytes);
            bulk_load(tile, p.envs + goff, bytes, bar);
        }
    } else {
        for (int i = lane; i < nvalid * n2; i += 32) tile[i] = p.envs[goff + i];
    }
    if (lane < 4) cnt_s[lane] = 0;
    const bool valid = t < nvalid;
    const int e = env0 + (valid ? t : 0);
    long long a = 0;
    if (STEP && valid) a = load_action(p.actions, p.action_bytes, (size_t)e);
    __syncwarp();
    if (bulk) mbar_wait(bar, 0);
    float* food = tile + (size_t)t * n2;
    float* head = food + C;
    float* gfood = p.envs + (size_t)e * n2;
    float* ghead = gfood + C;
    if (STEP && valid) {
        int hp = -1, hc = 0;
        for (int q = l; q < C; q += G)
            if (head[q] != 0.0f) { hp = q; ++hc; }
        hp = group_max<G>(hp, gm);
        hc = group_sum<G>(hc, gm);
        int np = -1, ny = -1, nx = -1;                                // :149-158 the conv2d translates the agent by -OFF[a]
        if (hp >= 0) {
            const int hy = (int)__umulhi((uint32_t)hp, p.magic_S), hx = hp - hy * S;
            ny = hy - off_y((int)a); nx = hx - off_x((int)a);
            if (ny >= 0 && ny < S && nx >= 0 && nx < S) np = ny * S + nx;
        }
        const float ov = np >= 0 ? food[np] : 0.0f;                   // :169 agent-food overlap
        __syncwarp(gm);
        if (l == 0 && hp >= 0) {
            head[hp] = 0.0f; ghead[hp] = 0.0f;
            if (np >= 0) {
                head[np] = 1.0f; ghead[np] = 1.0f;
                if (ov != 0.0f) { const float left = ov + ov * -1.0f; food[np] = left; gfood[np] = left; }   // :171
            }
        }
        __syncwarp(gm);
        if (ov != 0.0f) {                                             // :176-181 respawn
            const int cell = p.food_replay ? p.food_replay[e] : grid_pick_free_tile(p, food, call_counter(p), e);
            __syncwarp(gm);
            if (l == 0 && cell >= 0) { const float f = food[cell] + 1.0f; food[cell] = f; gfood[cell] = f; }
        }
        const bool interior = np >= 0 && ny >= 1 && ny <= S - 2 && nx >= 1 && nx <= S - 2;
        if (l == 0) {
            p.reward[e] = 0.0f - ov * -1.0f;
            p.done[e] = !interior;                                    // :189-194 edge collision is the only way to end
            if (hc > 1) atomicOr(p.status, WURM_ST_MULTI_HEAD);
            if (p.stats) {
                atomicAdd(&cnt_s[0], 1);
                if (!interior) atomicAdd(&cnt_s[1], 1);
                if (ov != 0.0f) atomicAdd(&cnt_s[2], 1);
            }
        }
        __syncwarp(gm);
    }
    // ---- observation (:88-133), rendered from the tile into the staging area ----
    if (valid && p.obs_mode == WURM_OBS_DEFAULT) {                    // black background (:90), agent green, food red
        float* o = stage + (size_t)t * 3 * C;
        for (int q = l; q < C; q += G) {
            const int y = (int)__umulhi((uint32_t)q, p.magic_S), x = q - y * S;
            float r = 0.0f, g = 0.0f;
            if (head[q] > kEps) { r = 0.0f; g = 1.0f; }
            if (food[q] > kEps) { r = 1.0f; g = 0.0f; }
            if (y == 0 || x == 0 || y == S - 1 || x == S - 1) r = g = 0.0f;
            o[q] = r; o[C + q] = g; o[2 * C + q] = 0.0f;
        }
    } else if (valid && p.obs_mode == WURM_OBS_POSITIONS) {           // first argmax of agent / food (:119-130)
        int idx[2];
        for (int ch = 0; ch < 2; ++ch) {
            const float* v = ch == 0 ? head : food;
            float bv = -INFINITY;
            int bq = 0;
            for (int q = l; q < C; q += G)
                if (v[q] > bv) { bv = v[q]; bq = q; }
            const float gv = group_max<G>(bv, gm);
            idx[ch] = -group_max<G>(bv == gv ? -bq : -(1 << 30), gm);
        }
        if (l == 0) {
            float* o = p.obs + (size_t)e * 4;
            o[0] = (float)(idx[0] / S); o[1] = (float)(idx[0] % S); o[2] = (float)(idx[1] / S); o[3] = (float)(idx[1] % S);
        }
    }
    fence_proxy_async();                    // generic-proxy writes -> visible to the bulk stores
    __syncwarp();
    if (STEP && p.stats && lane < 3 && cnt_s[lane]) {
        unsigned long long* slot = p.stats + (blockIdx.x % WURM_STATS_SLOTS) * WURM_STATS_FIELDS;
        if (lane == 0) atomicAdd(slot + WURM_STAT_ENV_STEPS, (unsigned long long)cnt_s[0]);
        if (lane == 1) {
            atomicAdd(slot + WURM_STAT_EPISODES, (unsigned long long)cnt_s[1]);
            atomicAdd(slot + WURM_STAT_EDGE_COLLISIONS, (unsigned long long)cnt_s[1]);
        }
        if (lane == 2) atomicAdd(slot + WURM_STAT_REWARD, (unsigned long long)cnt_s[2]);
    }
    if (p.obs_mode == WURM_OBS_DEFAULT || p.obs_mode == WURM_OBS_RAW) {
        const int per_env = p.obs_mode == WURM_OBS_DEFAULT ? 3 * C : n2;
        const float* src = p.obs_mode == WURM_OBS_DEFAULT ? stage : tile;   // raw: a copy of the (updated) tile
        float* dst = p.obs + (size_t)env0 * per_env;
        const uint32_t obytes = (uint32_t)(nvalid * per_env) * 4u;
        if ((obytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
            if (lane == 0) { bulk_store(dst, src, obytes); bulk_commit(); bulk_wait_read_all(); }
        } else {
            for (int i = lane; i < nvalid * per_env; i += 32) dst[i] = src[i];
        }
    }
}

// -1: the tile kernel does not apply.  Grids above 64 cells keep the lane-group kernels, and so does size 8: an env of 128
// floats puts the same cell of all eight envs of a tile on one shared-memory bank (measured on B200, 2^20 envs, `default`
// observations, step kernel: size 5 0.190 ms tile / 0.207 lanes, size 6 0.196 / 0.213, size 7 0.204 / 0.266, size 8 0.400 / 0.261).
template <bool STEP>
static int try_launch_grid_tile(const GridParams& p, cudaStream_t stream) {
    if (p.C > 64 || (2 * p.C) % 32 == 0 || getenv("WURM_GRID_NO_TILE")) return -1;
    constexpr int G = 4, T = 32 / G;
    const int tile_bytes_padded = (T * 2 * p.C * 4 + 15) & ~15;
    const int stage_env_floats = p.obs_mode == WURM_OBS_DEFAULT ? 3 * p.C : 0;
    const int smem = tile_bytes_padded + ((T * stage_env_floats * 4 + 15) & ~15) + 8 + 16 + 16;
    auto kern = grid_tile_kernel<G, STEP>;
    static SmemOptIn opt_in;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), &opt_in, smem, true, "cudaFuncSetAttribute(grid_tile_kernel)")) return rc;
    kern<<<(p.N + T - 1) / T, 32, smem, stream>>>(p, tile_bytes_padded, stage_env_floats);
    return check_launch("grid_tile_kernel");
}

template <bool STEP>
static int launch_grid_env(const GridParams& p, cudaStream_t stream) {
    {
        const int rc = try_launch_grid_tile<STEP>(p, stream);
        if (rc >= 0) return rc;
    }
    const long long threads_g8 = (long long)p.N * 8, threads_g32 = (long long)p.N * 32;
    if (p.C <= 64) grid_small_kernel<STEP><<<(unsigned)((threads_g8 + WURM_GRID_THREADS - 1) / WURM_GRID_THREADS), WURM_GRID_THREADS, 0, stream>>>(p);
    else grid_env_kernel<32, STEP><<<(unsigned)((threads_g32 + 255) / 256), 256, 0, stream>>>(p);
    return check_launch("grid_env_kernel");
}

// simple_gridworld.py:222-262: done envs get the agent at the start location and one food
__global__ void __launch_bounds__(256) grid_reset_kernel(const GridParams p, const uint8_t* __restrict__ done_mask) {
    const int lane = threadIdx.x & 31;
    const int e_base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (e_base >= p.N) return;
    const int e_mine = e_base + lane, C = p.C;
    unsigned todo = __ballot_sync(0xffffffffu, e_mine < p.N && done_mask[e_mine] != 0);
    while (todo) {
        const int e = e_base + __ffs(todo) - 1;
        todo &= todo - 1;
        float* env = p.envs + (size_t)e * 2 * C;
        for (int i = lane; i < 2 * C; i += 32) env[i] = (i == C + p.start_cell) ? 1.0f : 0.0f;
        __syncwarp();
        const int cell = p.food_replay ? p.food_replay[e] : grid_pick_free(p, env, call_counter(p), e, kStreamGridReset);
        if (lane == 0 && cell >= 0) env[cell] = 1.0f;
        __syncwarp();
    }
}

static int plan_grid(const WurmGridCfg* cfg, GridParams* p) {
    if (!cfg) return fail(WURM_E_INVALID, "cfg is NULL");
    if (cfg->num_envs <= 0) return fail(WURM_E_INVALID, "num_envs must be positive");
    if (cfg->size <= 4) return fail(WURM_E_INVALID, "size must be > 4 (reference simple_gridworld.py:239)");
    if (cfg->size > 256) return fail(WURM_E_UNSUPPORTED, "size > 256");
    if (cfg->obs_mode != WURM_OBS_NONE && cfg->obs_mode != WURM_OBS_DEFAULT && cfg->obs_mode != WURM_OBS_RAW &&
        cfg->obs_mode != WURM_OBS_POSITIONS)
        return fail(WURM_E_INVALID, "bad obs_mode (default, raw or positions)");
    p->N = cfg->num_envs; p->S = cfg->size; p->C = cfg->size * cfg->size; p->obs_mode = cfg->obs_mode;
    p->start_cell = cfg->start_y * cfg->size + cfg->start_x;
    p->magic_S = (uint32_t)((0x100000000ull + (uint64_t)p->S - 1) / (uint64_t)p->S);
    return WURM_OK;
}

}  // namespace wurm

using namespace wurm;

extern "C" int wurm_grid_step(const WurmGridCfg* cfg, float* envs, const void* actions, int action_bytes,
                              const int32_t* food_cell_replay, uint64_t seed, uint64_t step, const uint64_t* step_dev,
                              float* obs, float* reward, uint8_t* done, int32_t* status, int64_t* stats, void* stream) {
    GridParams p = {};
    if (int rc = plan_grid(cfg, &p)) return rc;
    if (!envs || !actions || !reward || !done || !status) return fail(WURM_E_INVALID, "NULL pointer");
    if (!valid_action_bytes(action_bytes)) return fail(WURM_E_INVALID, "action_bytes must be 1, 2, 4 or 8");
    if (cfg->obs_mode != WURM_OBS_NONE && !obs) return fail(WURM_E_INVALID, "obs is NULL");
    p.envs = envs; p.actions = actions; p.action_bytes = action_bytes; p.food_replay = food_cell_replay;
    p.seed = seed; p.step = step; p.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
    p.obs = obs; p.reward = reward; p.done = done; p.status = status; p.stats = reinterpret_cast<unsigned long long*>(stats);
    return launch_grid_env<true>(p, (cudaStream_t)stream);
}

extern "C" int wurm_grid_observe(const WurmGridCfg* cfg, const float* envs, float* obs, void* stream) {
    GridParams p = {};
    if (int rc = plan_grid(cfg, &p)) return rc;
    if (!envs || !obs) return fail(WURM_E_INVALID, "NULL pointer");
    if (cfg->obs_mode == WURM_OBS_NONE) return WURM_OK;
    p.envs = const_cast<float*>(envs); p.obs = obs;
    return launch_grid_env<false>(p, (cudaStream_t)stream);
}

extern "C" int wurm_grid_reset(const WurmGridCfg* cfg, float* envs, const uint8_t* done_mask, const int32_t* food_cell_replay,
                               uint64_t seed, uint64_t step, const uint64_t* step_dev, void* stream) {
    GridParams p = {};
    if (int rc = plan_grid(cfg, &p)) return rc;
    if (!envs || !done_mask) return fail(WURM_E_INVALID, "NULL pointer");
    if (cfg->start_y < 0 || cfg->start_y >= cfg->size || cfg->start_x < 0 || cfg->start_x >= cfg->size)
        return fail(WURM_E_INVALID, "start location outside the grid");
    p.envs = envs; p.food_replay = food_cell_replay; p.seed = seed; p.step = step;
    p.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
    grid_reset_kernel<<<(p.N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, done_mask);
    return check_launch("grid_reset_kernel");
}
