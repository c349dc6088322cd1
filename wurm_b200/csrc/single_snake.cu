// SingleSnake hot path for B200 (sm_100a): step / reset / observe as hand-written CUDA.
//
// Replaces wurm/envs/single_snake.py:104-387 and wurm/utils.py:36-65 of the reference (a ~214-op
// ATen tensor program with three conv2d calls and >=5 host syncs per step) with ONE launch per call.
//
// Design:
//   * one CTA per tile of T consecutive environments; the tile's (T,3,S,S) fp32 state is ONE
//     contiguous span of global memory, moved to shared memory with a single TMA bulk copy
//     (cp.async.bulk + mbarrier, SASS UBLKCP) -- no per-element load instructions for state;
//   * G consecutive lanes (a power of two, <=32) own one environment while it sits in shared
//     memory; reductions over the grid (snake size, head cell, free-cell ranking for the food
//     respawn) are warp shuffles / ballots restricted to the group's lane mask;
//   * the reference's conv2d filters become index arithmetic on the head cell;
//   * a step changes O(snake length) cells of an env: only those are written back, cell by cell,
//     where the shared copy is updated (the sectors were just loaded through L2, so the partial
//     writes merge there) -- HBM write traffic drops from the full state to a few sectors per env;
//   * partial_n observations are rendered by the env's own lane group into a shared staging area
//     and leave as ONE bulk store per tile (the tile's observations are contiguous too); `raw` is
//     a bulk store of the tile itself; the full-grid modes are coalesced row stores;
//   * no tensor cores: nothing on this path is a dense contraction.
#include <math.h>
#include <stdlib.h>

#include "../../include/wurm_b200.h"
#include "common.cuh"
#include "host_util.h"
#include "single_device.cuh"

namespace wurm {

// (SingleParams, the per-env step and the partial-observation renderer live in single_device.cuh: shared with the
// compact-state kernels of single_compact.cu)

// single_snake.py:130-195 from the shared tile into the caller's observation buffer.
template <int G>
__device__ __forceinline__ void write_obs(const SingleParams& p, const float* tile, int env0, int nvalid) {
    const int S = p.S, C = p.C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (p.obs_mode == WURM_OBS_DEFAULT || p.obs_mode == WURM_OBS_ONE_CHANNEL) {
        for (int t = warp; t < nvalid; t += nwarps) {
            const float* env = tile + (size_t)t * 3 * C;
            for (int q = lane; q < C; q += 32) {
                const int y = div_S(q, p.magic_S), x = q - y * S;
                const bool border = y == 0 || x == 0 || y == S - 1 || x == S - 1;
                const float f = env[q], h = env[C + q], b = env[2 * C + q];
                if (p.obs_mode == WURM_OBS_DEFAULT) {                // :131-138
                    float* o = p.obs + (size_t)(env0 + t) * 3 * C;
                    o[q] = rgb_channel(f, h, b, border, 0);
                    o[C + q] = rgb_channel(f, h, b, border, 1);
                    o[2 * C + q] = rgb_channel(f, h, b, border, 2);
                } else {                                             // :142-151
                    float v = (b > kEps ? 1.0f : 0.0f) * 0.5f;
                    v += h * 0.5f;
                    v += f * 1.5f;
                    p.obs[(size_t)(env0 + t) * C + q] = border ? -1.0f : v;
                }
            }
        }
    } else if (p.obs_mode == WURM_OBS_POSITIONS) {                   // :152-165 first argmax of head / food
        const unsigned gm = group_mask<G>();
        const int t = threadIdx.x / G, l = threadIdx.x % G;
        if (t < nvalid) {
            const float* env = tile + (size_t)t * 3 * C;
            int idx[2];
            for (int ch = 0; ch < 2; ++ch) {
                const float* v = env + (ch == 0 ? C : 0);
                float bv = -INFINITY;
                int bq = 0;
                for (int q = l; q < C; q += G)
                    if (v[q] > bv) { bv = v[q]; bq = q; }
                const float gv = group_max<G>(bv, gm);
                idx[ch] = -group_max<G>(bv == gv ? -bq : -(1 << 30), gm);   // smallest index attaining the max
            }
            if (l == 0) {
                float* o = p.obs + (size_t)(env0 + t) * 4;
                const int hy = div_S(idx[0], p.magic_S), fy = div_S(idx[1], p.magic_S);
                o[0] = (float)hy; o[1] = (float)(idx[0] - hy * S);
                o[2] = (float)fy; o[3] = (float)(idx[1] - fy * S);
            }
        }
    }
    // WURM_OBS_RAW is a second store of the tile and WURM_OBS_PARTIAL a store of the staging area, both
    // issued by the caller.
}

// One CTA = one tile of T envs.  STEP: load -> step -> store -> observe.  !STEP: load -> observe.
template <int G, bool STEP>
__global__ void __launch_bounds__(256) single_tile_kernel(const SingleParams p) {
    const int KC = p.C;                              // cells per channel
    extern __shared__ __align__(128) unsigned char smem[];
    float* tile = reinterpret_cast<float*>(smem);
    float* stage = reinterpret_cast<float*>(smem + p.tile_bytes_padded);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + p.tile_bytes_padded + p.stage_bytes);
    int* cnt_s = reinterpret_cast<int*>(bar + 1);

    const int env0 = blockIdx.x * p.T;
    const int nvalid = min(p.T, p.N - env0);
    const size_t goff = (size_t)env0 * 3 * KC;
    const int nfloats = nvalid * 3 * KC;
    const uint32_t bytes = (uint32_t)nfloats * 4u;
    const bool bulk = p.bulk_ok && (bytes % 16u == 0u);
    const bool partial = p.obs_mode == WURM_OBS_PARTIAL;
    const int E = 3 * p.W * p.W;                    // floats per env of a partial observation

    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            mbar_arrive_expect_tx(bar, bytes);
            bulk_load(tile, p.envs + goff, bytes, bar);
        }
    } else {
        for (int i = threadIdx.x; i < nfloats; i += blockDim.x) tile[i] = p.envs[goff + i];
    }
    if (threadIdx.x < 4) cnt_s[threadIdx.x] = 0;
    // per-env scalars (action, hints) are fetched while the tile is still in flight
    const int t = threadIdx.x / G, l = threadIdx.x % G;
    long long a_in = 0;
    int hint_head = -1, hint_sz = -1;
    if (STEP && t < nvalid) {
        const size_t e = (size_t)(env0 + t);
        a_in = load_action(p.actions, p.action_bytes, e);
        if (p.hints) { hint_head = p.hints[4 * e]; hint_sz = p.hints[4 * e + 1]; }
    }
    __syncthreads();                        // mbarrier init / fallback tile visible
    if (bulk) mbar_wait(bar, 0);

    bool ended = false;
    if (t < nvalid) {
        float* env = tile + (size_t)t * 3 * KC;
        int hp = -1;
        if (STEP) hp = step_env<G>(p, env, env0 + t, l, cnt_s, a_in, hint_head, hint_sz, ended);
        else if (partial) hp = find_head<G>(p, env, l);
        if (partial) render_partial<G>(p, env, hp, stage + (size_t)t * E, l);
    }
    if (STEP && p.auto_reset) {
        // Fused reset (:322-337): envs that ended are re-created in HBM right away, one after another by the WHOLE
        // warp (a lane group doing it alone would stall the warp's other envs for 3C/G iterations).  The shared
        // copy keeps the terminal state -- the observation returned by step is the terminal one, and it is what
        // the reference's driver feeds its policy next (main.py:227) -- and HBM holds exactly that state at this
        // point, so only its non-zero cells are cleared before the five cells of the new env are written.
        const int lane = threadIdx.x & 31;
        int my_tail = 0, my_mid = 0, my_hd = 0, my_cell = -1;
        const bool mine = ended && l == 0;
        __syncwarp();                       // every group's updates of the shared tile -> visible to the whole warp
        if (mine) new_env_layout(p, p.spawn, call_counter(p) + 1, env0 + t, my_tail, my_mid, my_hd, my_cell);
        unsigned todo = __ballot_sync(0xffffffffu, mine);
        const int C = KC, n = 3 * C;
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int tail = __shfl_sync(0xffffffffu, my_tail, src), mid = __shfl_sync(0xffffffffu, my_mid, src);
            const int hd = __shfl_sync(0xffffffffu, my_hd, src), cell = __shfl_sync(0xffffffffu, my_cell, src);
            const int ts = ((threadIdx.x & ~31) + src) / G;
            const float* old = tile + (size_t)ts * n;
            float* g = p.envs + (size_t)(env0 + ts) * n;
            for (int i = lane; i < n; i += 32)
                if (old[i] != 0.0f) g[i] = 0.0f;
            __syncwarp();
            if (lane < 5) {
                const int idx = lane == 0 ? cell : lane == 1 ? C + hd : lane == 2 ? 2 * C + tail : lane == 3 ? 2 * C + mid : 2 * C + hd;
                const float val = lane == 3 ? 2.0f : lane == 4 ? 3.0f : 1.0f;
                if (idx >= 0) g[idx] = val;
            }
            if (lane == 0 && p.hints) { short* h = p.hints + 4 * (size_t)(env0 + ts); h[0] = (short)hd; h[1] = 3; h[2] = (short)cell; }
        }
    }
    fence_proxy_async();                    // generic-proxy writes -> visible to the bulk stores
    __syncthreads();
    if (STEP && p.stats && threadIdx.x < WURM_STATS_FIELDS) {   // one striped slot per CTA: no hot address in L2
        unsigned long long* slot = p.stats + (blockIdx.x % WURM_STATS_SLOTS) * WURM_STATS_FIELDS;
        const int v = threadIdx.x == 0 ? nvalid : cnt_s[threadIdx.x - 1];
        if (v) atomicAdd(slot + threadIdx.x, (unsigned long long)v);
    }
    const bool raw = p.obs_mode == WURM_OBS_RAW;
    if (raw) {                              // the 'raw' observation is a copy of the (updated) tile
        if (bulk) {
            if (threadIdx.x == 0) bulk_store(p.obs + goff, tile, bytes);
        } else {
            for (int i = threadIdx.x; i < nfloats; i += blockDim.x) p.obs[goff + i] = tile[i];
        }
    }
    bool obs_bulk = false;
    if (partial) {                          // the tile's partial observations are one contiguous span too
        float* dst = p.obs + (size_t)env0 * E;
        const uint32_t obytes = (uint32_t)(nvalid * E) * 4u;
        obs_bulk = p.bulk_ok && (obytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
        if (obs_bulk) {
            if (threadIdx.x == 0) bulk_store(dst, stage, obytes);
        } else {
            for (int i = threadIdx.x; i < nvalid * E; i += blockDim.x) dst[i] = stage[i];
        }
    } else if (p.obs_mode >= 0 && !raw) {
        write_obs<G>(p, tile, env0, nvalid);
    }
    if (((bulk && raw) || obs_bulk) && threadIdx.x == 0) {
        bulk_commit();
        bulk_wait_read_all();               // shared memory must outlive the bulk stores' reads
    }
}

// ---------------------------------------------------------------------------------------------
// EXPERIMENT (WURM_SINGLE_RING=stages): persistent one-warp CTAs with a multi-stage TMA ring
// ---------------------------------------------------------------------------------------------
// single_tile_kernel overlaps load, compute and store by oversubscription: ~20 one-warp CTAs per SM, each doing
// load -> wait -> step -> store once.  This variant makes the overlap explicit instead: grid = a few CTAs per SM, each
// walking tiles blockIdx, blockIdx + gridDim, ... with a ring of STAGES tile buffers -- the bulk load of tile i+STAGES is
// issued (cp.async.bulk + mbarrier expect_tx) as soon as tile i has been stepped, so every warp always has STAGES-1
// loads in flight while it computes -- and two observation staging buffers so that the bulk store of tile i drains while
// tile i+1 is rendered.  Partial observations, full tiles and 16-byte-aligned bases only (the C2 shape); selected by the
// environment variable for the measurement recorded in profiles/r02_c2_ring_experiment.txt.
template <int G, int STAGES>
__global__ void __launch_bounds__(32) single_ring_kernel(const SingleParams p, int ntiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int KC = p.C, n3 = 3 * KC;
    const int E = 3 * p.W * p.W;
    float* tiles = reinterpret_cast<float*>(smem);
    float* stages = reinterpret_cast<float*>(smem + (size_t)STAGES * p.tile_bytes_padded);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * p.tile_bytes_padded + 2 * (size_t)p.stage_bytes);
    int* cnt_s = reinterpret_cast<int*>(bars + STAGES);
    const int lane = threadIdx.x, t = lane / G, l = lane % G;
    const uint32_t bytes = (uint32_t)(p.T * n3) * 4u;
    if (lane < 4) cnt_s[lane] = 0;
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
        fence_mbar_init();
        for (int s = 0; s < STAGES; ++s) {
            const int tile = blockIdx.x + s * gridDim.x;
            if (tile < ntiles) {
                mbar_arrive_expect_tx(bars + s, bytes);
                bulk_load(reinterpret_cast<unsigned char*>(tiles) + (size_t)s * p.tile_bytes_padded, p.envs + (size_t)tile * p.T * n3, bytes, bars + s);
            }
        }
    }
    __syncwarp();
    int done_tiles = 0;
    for (int tile = blockIdx.x, it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        float* tbuf = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(tiles) + (size_t)s * p.tile_bytes_padded);
        float* stage = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stages) + (size_t)(it & 1) * p.stage_bytes);
        const int env0 = tile * p.T;
        const size_t e = (size_t)(env0 + t);
        const long long a_in = load_action(p.actions, p.action_bytes, e);       // per-env scalars while the tile is in flight
        int hint_head = -1, hint_sz = -1;
        if (p.hints) { hint_head = p.hints[4 * e]; hint_sz = p.hints[4 * e + 1]; }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // this staging buffer's last store has read it
        __syncwarp();
        mbar_wait(bars + s, parity);
        float* env = tbuf + (size_t)t * n3;
        bool ended = false;
        const int hp = step_env<G>(p, env, env0 + t, l, cnt_s, a_in, hint_head, hint_sz, ended);
        render_partial<G>(p, env, hp, stage + (size_t)t * E, l);
        if (p.auto_reset) {                                         // fused reset, as in single_tile_kernel
            int my_tail = 0, my_mid = 0, my_hd = 0, my_cell = -1;
            const bool mine = ended && l == 0;
            __syncwarp();
            if (mine) new_env_layout(p, p.spawn, call_counter(p) + 1, env0 + t, my_tail, my_mid, my_hd, my_cell);
            unsigned todo = __ballot_sync(0xffffffffu, mine);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int tail = __shfl_sync(0xffffffffu, my_tail, src), mid = __shfl_sync(0xffffffffu, my_mid, src);
                const int hd = __shfl_sync(0xffffffffu, my_hd, src), cell = __shfl_sync(0xffffffffu, my_cell, src);
                const int ts = src / G;
                const float* old = tbuf + (size_t)ts * n3;
                float* g = p.envs + (size_t)(env0 + ts) * n3;
                for (int i = lane; i < n3; i += 32)
                    if (old[i] != 0.0f) g[i] = 0.0f;
                __syncwarp();
                if (lane < 5) {
                    const int idx = lane == 0 ? cell : lane == 1 ? KC + hd : lane == 2 ? 2 * KC + tail : lane == 3 ? 2 * KC + mid : 2 * KC + hd;
                    const float val = lane == 3 ? 2.0f : lane == 4 ? 3.0f : 1.0f;
                    if (idx >= 0) g[idx] = val;
                }
                if (lane == 0 && p.hints) { short* h = p.hints + 4 * (size_t)(env0 + ts); h[0] = (short)hd; h[1] = 3; h[2] = (short)cell; }
            }
        }
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();                                    // the warp's generic reads / writes of tile and stage come first
            bulk_store(p.obs + (size_t)env0 * E, stage, (uint32_t)(p.T * E) * 4u);
            bulk_commit();
            const int next = tile + STAGES * gridDim.x;             // refill this ring slot
            if (next < ntiles) {
                mbar_arrive_expect_tx(bars + s, bytes);
                bulk_load(tbuf, p.envs + (size_t)next * p.T * n3, bytes, bars + s);
            }
        }
        ++done_tiles;
    }
    __syncwarp();
    if (p.stats && lane < WURM_STATS_FIELDS) {
        unsigned long long* slot = p.stats + (blockIdx.x % WURM_STATS_SLOTS) * WURM_STATS_FIELDS;
        const int v = lane == 0 ? done_tiles * p.T : cnt_s[lane - 1];
        if (v) atomicAdd(slot + lane, (unsigned long long)v);
    }
    if (lane == 0) bulk_wait_read_all();
}

// ---------------------------------------------------------------------------------------------
// Body-only tiles: the steady-state path for larger grids
// ---------------------------------------------------------------------------------------------
// In steady state the previous call left every env's head cell, snake size and food cell.  If the head cell still
// holds a head, the food cell still holds food, and the body channel's maximum and its (size, size-1) cells are what
// the hints say, the step needs NOTHING else from the food and head channels: two thirds of the state are not read.
// For even grid sides a channel is a 16-byte-aligned span, so the tile is loaded as one bulk copy PER ENV of its body
// channel only (all on one mbarrier).  Observations are rendered from the body tile plus the known head and food
// cells.  An env whose hints do not verify takes the general step straight on global memory (the tile kernel's code
// with `env` pointing at HBM): slow, rare, and it re-derives the hints.  Measured on B200 this pays for large envs
// (size 36: 15.5 KB of state) and loses for tiny ones (size 9: scattered sectors, sparse stores into sectors that are
// no longer in L2, two DRAM round trips instead of one), so the host uses it from size 16 up.
template <int G>
__device__ __noinline__ int pick_free_cell_body(const float* body, int head_cell, int S, int C, uint32_t magic, uint32_t rnd, int l,
                                                unsigned gm) {
    const unsigned shift = (G == 32) ? 0u : ((threadIdx.x & 31u) & ~(unsigned)(G - 1));
    const unsigned low = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    auto free_bits = [&](int base) {
        const int q = base + l;
        bool f = false;
        if (q < C) {
            const int y = div_S(q, magic), x = q - y * S;
            f = y >= 1 && y <= S - 2 && x >= 1 && x <= S - 2 && (body[q] + (q == head_cell ? 1.0f : 0.0f) < kEps);
        }
        return (__ballot_sync(gm, f) >> shift) & low;
    };
    int nfree = 0;
    for (int base = 0; base < C; base += G) nfree += __popc(free_bits(base));
    if (nfree == 0) return -1;
    int r = (int)bounded(rnd, (uint32_t)nfree);
    for (int base = 0; base < C; base += G) {
        const unsigned bits = free_bits(base);
        const int cnt = __popc(bits);
        if (r < cnt) return base + (int)__fns(bits, 0, r + 1);
        r -= cnt;
    }
    return -1;
}

// The general step for one env straight on global memory, out of line and with the parameters BY VALUE (a reference
// would pin the kernel's parameter struct to every thread's stack).  Returns (new head cell, food cell, ended).
template <int G>
__device__ __noinline__ int3 slow_step_on_global(const SingleParams p, int e, int l, int* cnt_s, long long a_in, int hint_head,
                                                 int hint_sz, float* partial_out) {
    const unsigned gm = group_mask<G>();
    const int C = p.C;
    float* genv = p.envs + (size_t)e * 3 * C;
    bool ended = false;
    const int np = step_env<G>(p, genv, e, l, cnt_s, a_in, hint_head, hint_sz, ended);
    __syncwarp(gm);
    int fc = 0, fq = -1;
    for (int q = l; q < C; q += G)
        if (genv[q] == 1.0f) { ++fc; fq = q; }
    fc = group_sum<G>(fc, gm);
    fq = group_max<G>(fq, gm);
    const int food_now = fc == 1 ? fq : -1;
    if (l == 0 && p.hints) p.hints[4 * (size_t)e + 2] = (short)food_now;
    // the observation of this env is rendered here, before a fused reset can overwrite the terminal state in HBM
    if (p.obs_mode == WURM_OBS_PARTIAL) render_partial<G>(p, genv, np, partial_out, l);
    if (p.obs_mode == WURM_OBS_DEFAULT || p.obs_mode == WURM_OBS_ONE_CHANNEL) {
        const int S = p.S;
        for (int q = l; q < C; q += G) {
            const int y = div_S(q, p.magic_S), x = q - y * S;
            const bool border = y == 0 || x == 0 || y == S - 1 || x == S - 1;
            const float f = genv[q], h = genv[C + q], b = genv[2 * C + q];
            if (p.obs_mode == WURM_OBS_DEFAULT) {
                float* o = p.obs + (size_t)e * 3 * C;
                o[q] = rgb_channel(f, h, b, border, 0);
                o[C + q] = rgb_channel(f, h, b, border, 1);
                o[2 * C + q] = rgb_channel(f, h, b, border, 2);
            } else {
                float v = (b > kEps ? 1.0f : 0.0f) * 0.5f;
                v += h * 0.5f;
                v += f * 1.5f;
                p.obs[(size_t)e * C + q] = border ? -1.0f : v;
            }
        }
    }
    return make_int3(np, food_now, ended ? 1 : 0);
}

// Envs too large for a shared-memory tile (grid sides above 128: 3*S*S floats no longer fit next to anything else):
// one warp per env, the general step straight on global memory -- slow_step_on_global, the code the body-only kernel
// runs for envs whose hints do not verify -- so that the reference's "any size" holds here too.  Not a tuned path.
template <bool STEP>
__global__ void __launch_bounds__(128) single_global_kernel(const SingleParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int C = p.C, S = p.S, E = 3 * p.W * p.W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* stage = reinterpret_cast<float*>(smem) + (size_t)warp * (p.obs_mode == WURM_OBS_PARTIAL ? E : 0);
    int* cnt_s = reinterpret_cast<int*>(smem + p.stage_bytes);
    if (threadIdx.x < 4) cnt_s[threadIdx.x] = 0;
    __syncthreads();
    const int e = blockIdx.x * nwarps + warp;
    if (e < p.N) {
        float* genv = p.envs + (size_t)e * 3 * C;
        int np = -1;
        bool ended = false;
        if (STEP) {
            const long long a_in = load_action(p.actions, p.action_bytes, (size_t)e);
            int hint_head = -1, hint_sz = -1;
            if (p.hints) { hint_head = p.hints[4 * (size_t)e]; hint_sz = p.hints[4 * (size_t)e + 1]; }
            const int3 res = slow_step_on_global<32>(p, e, lane, cnt_s, a_in, hint_head, hint_sz, stage);
            np = res.x; ended = res.z != 0;
        } else {
            if (p.obs_mode == WURM_OBS_PARTIAL) {
                np = find_head<32>(p, genv, lane);
                render_partial<32>(p, genv, np, stage, lane);
            } else if (p.obs_mode == WURM_OBS_DEFAULT || p.obs_mode == WURM_OBS_ONE_CHANNEL) {
                for (int q = lane; q < C; q += 32) {
                    const int y = div_S(q, p.magic_S), x = q - y * S;
                    const bool border = y == 0 || x == 0 || y == S - 1 || x == S - 1;
                    const float f = genv[q], h = genv[C + q], b = genv[2 * C + q];
                    if (p.obs_mode == WURM_OBS_DEFAULT) {
                        float* o = p.obs + (size_t)e * 3 * C;
                        o[q] = rgb_channel(f, h, b, border, 0); o[C + q] = rgb_channel(f, h, b, border, 1); o[2 * C + q] = rgb_channel(f, h, b, border, 2);
                    } else {
                        float v = (b > kEps ? 1.0f : 0.0f) * 0.5f;
                        v += h * 0.5f;
                        v += f * 1.5f;
                        p.obs[(size_t)e * C + q] = border ? -1.0f : v;
                    }
                }
            }
        }
        __syncwarp();
        if (p.obs_mode == WURM_OBS_PARTIAL) {
            for (int r = lane; r < E; r += 32) p.obs[(size_t)e * E + r] = stage[r];
        } else if (p.obs_mode == WURM_OBS_RAW) {                     // :139-141
            for (int i = lane; i < 3 * C; i += 32) p.obs[(size_t)e * 3 * C + i] = genv[i];
        } else if (p.obs_mode == WURM_OBS_POSITIONS) {               // :152-165 first argmax of head / food
            int idx[2];
            for (int ch = 0; ch < 2; ++ch) {
                const float* v = genv + (ch == 0 ? C : 0);
                float bv = -INFINITY;
                int bq = 0;
                for (int q = lane; q < C; q += 32)
                    if (v[q] > bv) { bv = v[q]; bq = q; }
                const float gv = group_max<32>(bv, 0xffffffffu);
                idx[ch] = -group_max<32>(bv == gv ? -bq : -(1 << 30), 0xffffffffu);
            }
            if (lane == 0) {
                float* o = p.obs + (size_t)e * 4;
                const int hy = div_S(idx[0], p.magic_S), fy = div_S(idx[1], p.magic_S);
                o[0] = (float)hy; o[1] = (float)(idx[0] - hy * S); o[2] = (float)fy; o[3] = (float)(idx[1] - fy * S);
            }
        }
        if (STEP && p.auto_reset && ended) {                         // fused reset (:322-337), see single_tile_kernel
            __syncwarp();
            int tail = 0, mid = 0, hd = 0, cell = -1;
            new_env_layout(p, p.spawn, call_counter(p) + 1, e, tail, mid, hd, cell);
            for (int i = lane; i < 3 * C; i += 32)
                if (genv[i] != 0.0f) genv[i] = 0.0f;
            __syncwarp();
            if (lane < 5) {
                const int idx = lane == 0 ? cell : lane == 1 ? C + hd : lane == 2 ? 2 * C + tail : lane == 3 ? 2 * C + mid : 2 * C + hd;
                const float val = lane == 3 ? 2.0f : lane == 4 ? 3.0f : 1.0f;
                if (idx >= 0) genv[idx] = val;
            }
            if (lane == 0 && p.hints) { short* h = p.hints + 4 * (size_t)e; h[0] = (short)hd; h[1] = 3; h[2] = (short)cell; }
        }
    }
    __syncthreads();
    if (STEP && p.stats && threadIdx.x < WURM_STATS_FIELDS) {
        unsigned long long* slot = p.stats + (blockIdx.x % WURM_STATS_SLOTS) * WURM_STATS_FIELDS;
        const int v = threadIdx.x == 0 ? min(nwarps, p.N - (int)blockIdx.x * nwarps) : cnt_s[threadIdx.x - 1];
        if (v) atomicAdd(slot + threadIdx.x, (unsigned long long)v);
    }
}

// 64 threads (two size-36 envs) and >= 12 CTAs per SM measured best on B200 (profiles/r01_sweep_single_body.txt)
#ifndef WURM_BODY_MINB
#define WURM_BODY_MINB 12
#endif
#ifndef WURM_BODY_THREADS
#define WURM_BODY_THREADS 64
#endif
template <int G>
__global__ void __launch_bounds__(WURM_BODY_THREADS, WURM_BODY_MINB) single_body_kernel(const SingleParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int C = p.C, S = p.S;
    float* body_all = reinterpret_cast<float*>(smem);
    float* stage = reinterpret_cast<float*>(smem + p.tile_bytes_padded);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + p.tile_bytes_padded + p.stage_bytes);
    int* cnt_s = reinterpret_cast<int*>(bar + 1);                     // 4 counters, then per env: fast, head cell, food cell
    int* env_s = cnt_s + 4;
    const int tid = threadIdx.x, lane = tid & 31, t = tid / G, l = tid % G;
    const unsigned gm = group_mask<G>();
    const int env0 = blockIdx.x * p.T, nvalid = min(p.T, p.N - env0);
    const bool valid = t < nvalid;
    const int e = env0 + (valid ? t : 0);
    const bool partial = p.obs_mode == WURM_OBS_PARTIAL;
    const int E = 3 * p.W * p.W;
    float* gfood = p.envs + (size_t)e * 3 * C;
    float* ghead = gfood + C;
    float* gbody = gfood + 2 * C;
    float* body = body_all + (size_t)t * C;

    if (tid == 0) {                                                   // one bulk copy per env: its body channel
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, (uint32_t)(nvalid * C) * 4u);
        for (int i = 0; i < nvalid; ++i)
            bulk_load(body_all + (size_t)i * C, p.envs + ((size_t)(env0 + i) * 3 + 2) * C, (uint32_t)C * 4u, bar);
    }
    if (tid < 4) cnt_s[tid] = 0;
    long long a_in = 0;
    int hint_head = -1, hint_sz = -1, hint_food = -1;
    if (valid) {
        a_in = load_action(p.actions, p.action_bytes, (size_t)e);
        const short4 h = *reinterpret_cast<const short4*>(p.hints + 4 * (size_t)e);
        hint_head = h.x; hint_sz = h.y; hint_food = h.z;
    }
    float hv = 0.0f, fv = 0.0f;                                       // the two verification loads
    if (valid && hint_head >= 0 && hint_head < C) hv = ghead[hint_head];
    if (valid && hint_food >= 0 && hint_food < C) fv = gfood[hint_food];
    __syncthreads();
    mbar_wait(bar, 0);

    float m = -INFINITY;
    int c1 = 0, c2 = 0, p1 = -1, p2 = -1;
    const float hint_size = (float)hint_sz, hint_sm1 = hint_size - 1.0f;
    if (valid) {
#pragma unroll 4
        for (int q = l; q < C; q += G) {
            const float v = body[q];
            m = fmaxf(m, v);
            if (v == hint_size) { ++c1; p1 = q; }
            if (v == hint_sm1) { ++c2; p2 = q; }
        }
    }
    const float size = group_max<G>(m, gm);
    c1 = group_sum<G>(c1, gm);
    c2 = group_sum<G>(c2, gm);
    p1 = group_max<G>(p1, gm);
    p2 = group_max<G>(p2, gm);
    const bool fast = valid && hv != 0.0f && fv == 1.0f && hint_sz >= 1 && size == hint_size && c1 == 1 && c2 == 1;

    int np = -1, food_now = hint_food;
    bool ended = false;
    if (valid && !fast) {
        const int3 res = slow_step_on_global<G>(p, e, l, cnt_s, a_in, hint_head, hint_sz, stage + (size_t)t * E);
        np = res.x; food_now = res.y; ended = res.z != 0;
    }
    if (fast) {
        const int hp = hint_head;
        int k = 0;                                                    // orientation (:212), canonical case
        {
            const int d = p1 - p2;
            const int x2 = p2 - div_S(p2, p.magic_S) * S;
            if (d == -S) k = 0;
            else if (d == 1 && x2 != S - 1) k = 1;
            else if (d == S) k = 2;
            else if (d == -1 && x2 != 0) k = 3;
        }
        const long long a = (a_in + ((long long)k == a_in ? 2 : 0)) % 4;     // :221-222
        if (l == 0 && a != a_in) store_action(p.actions, p.action_bytes, (size_t)e, a);
        int ny, nx;
        {
            const int hy = div_S(hp, p.magic_S), hx = hp - hy * S;       // :225-233
            ny = hy - (a >= 0 ? off_y((int)a) : 0);
            nx = hx - (a >= 0 ? off_x((int)a) : 0);
            if (ny >= 0 && ny < S && nx >= 0 && nx < S) np = ny * S + nx;
        }
        const float ov = (np >= 0 && np == hint_food) ? fv : 0.0f;    // :242 the only food is on the hinted cell
        if (ov == 0.0f) {                                            // :246-249 decay unless it ate
#pragma unroll 4
            for (int q = l; q < C; q += G) {
                const float v = body[q], nv = fmaxf(v - 1.0f, 0.0f);
                if (nv != v) { body[q] = nv; gbody[q] = nv; }
            }
        }
        __syncwarp(gm);
        const bool sc = (np >= 0) && (body[np] > kEps);              // :252
        const bool interior = (np >= 0) && ny >= 1 && ny <= S - 2 && nx >= 1 && nx <= S - 2;
        __syncwarp(gm);
        if (l == 0) {
            ghead[hp] = 0.0f;
            if (np >= 0) {
                ghead[np] = 1.0f;
                const float grown = body[np] + (size + ov);          // :258-262
                body[np] = grown; gbody[np] = grown;
                if (ov != 0.0f) gfood[np] = fv + ov * -1.0f;         // :270-272
            }
        }
        __syncwarp(gm);
        if (ov != 0.0f) {                                            // :277-282 respawn
            int cell;
            if (p.food_replay) cell = p.food_replay[e];
            else {
                const int I = S - 2;
                cell = -1;
                for (uint32_t tr = 0; tr < kRejectionTries && cell < 0; ++tr) {
                    const int cand = (int)bounded(draw_i(p.seed, call_counter(p), (uint32_t)e, kStreamSingleStepFood, tr), (uint32_t)(I * I));
                    const int cy = cand / I, q = (1 + cy) * S + 1 + (cand - cy * I);
                    if (body[q] + (q == np ? 1.0f : 0.0f) < kEps) cell = q;    // the eaten food is gone: the food channel is all zero
                }
                if (cell < 0)
                    cell = pick_free_cell_body<G>(body, np, S, C, p.magic_S,
                                                  draw_i(p.seed, call_counter(p), (uint32_t)e, kStreamSingleStepFood, kRejectionTries), l, gm);
            }
            if (l == 0 && cell >= 0) gfood[cell] = (cell == np ? fv + ov * -1.0f : 0.0f) + 1.0f;
            food_now = cell;
        }
        if (l == 0) {
            p.reward[e] = 0.0f - ov * -1.0f;                             // :271
            p.self_col[e] = sc;
            p.edge_col[e] = !interior;                                   // :290-293
            p.done[e] = sc || !interior;
            if (p.packed) p.packed[e] = pack_result(sc || !interior, sc, !interior, ov);
            short4 h;
            h.x = (short)np; h.y = (short)((np >= 0) ? (int)(size + ov) : -1); h.z = (short)food_now; h.w = 0;
            *reinterpret_cast<short4*>(p.hints + 4 * (size_t)e) = h;
            if (p.stats) {
                if (sc || !interior) atomicAdd(cnt_s + 0, 1);
                if (ov != 0.0f) atomicAdd(cnt_s + 1, (int)ov);
                if (sc) atomicAdd(cnt_s + 2, 1);
                if (!interior) atomicAdd(cnt_s + 3, 1);
            }
        }
        ended = sc || !interior;
        __syncwarp(gm);
        if (partial) {                                               // :166-193 the crop around the new head
            float* out = stage + (size_t)t * E;
            const int W = p.W, WW = W * W, n = p.obs_n;
            if (np < 0) {
                for (int q = l; q < 3 * WW; q += G) out[q] = 0.0f;
                if (l == 0) atomicOr(p.status, WURM_ST_NO_HEAD_PARTIAL);
            } else {
                constexpr float kHalf = 127.0f / 255.0f;
                for (int ij = l; ij < WW; ij += G) {
                    const int i = (int)__umulhi((uint32_t)ij, p.magic_W), j = ij - i * W;
                    const int y = ny - n + i, x = nx - n + j;
                    float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
                    if ((unsigned)(y - 1) < (unsigned)(S - 2) && (unsigned)(x - 1) < (unsigned)(S - 2)) {
                        const int q = y * S + x;
                        v0 = v1 = v2 = 1.0f;
                        if (body[q] > kEps) { v0 = 0.0f; v1 = kHalf; v2 = 0.0f; }
                        if (q == np) { v0 = 0.0f; v1 = 1.0f; v2 = 0.0f; }
                        if (q == food_now) { v0 = 1.0f; v1 = 0.0f; v2 = 0.0f; }
                    }
                    out[ij] = v0; out[WW + ij] = v1; out[2 * WW + ij] = v2;
                }
            }
        }
    }
    if (valid && l == 0) { env_s[3 * t] = fast; env_s[3 * t + 1] = np; env_s[3 * t + 2] = food_now; }

    if (p.auto_reset) {                                              // fused reset, see single_tile_kernel
        int my_tail = 0, my_mid = 0, my_hd = 0, my_cell = -1;
        const bool mine = ended && l == 0;
        __syncwarp();
        if (mine) new_env_layout(p, p.spawn, call_counter(p) + 1, e, my_tail, my_mid, my_hd, my_cell);
        unsigned todo = __ballot_sync(0xffffffffu, mine);
        const int n = 3 * C;
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int tail = __shfl_sync(0xffffffffu, my_tail, src), mid = __shfl_sync(0xffffffffu, my_mid, src);
            const int hd = __shfl_sync(0xffffffffu, my_hd, src), cell = __shfl_sync(0xffffffffu, my_cell, src);
            const int was_fast = __shfl_sync(0xffffffffu, (int)fast, src);
            const int old_head = __shfl_sync(0xffffffffu, np, src), old_food = __shfl_sync(0xffffffffu, food_now, src);
            const int ts = ((tid & ~31) + src) / G;
            float* g = p.envs + (size_t)(env0 + ts) * n;
            if (was_fast) {                                          // the terminal state's non-zero cells are known
                const float* ob = body_all + (size_t)ts * C;
                for (int i = lane; i < C; i += 32)
                    if (ob[i] != 0.0f) g[2 * C + i] = 0.0f;
                if (lane == 0 && old_head >= 0) g[C + old_head] = 0.0f;
                if (lane == 1 && old_food >= 0) g[old_food] = 0.0f;
            } else {
                for (int i = lane; i < n; i += 32)
                    if (g[i] != 0.0f) g[i] = 0.0f;
            }
            __syncwarp();
            if (lane < 5) {
                const int idx = lane == 0 ? cell : lane == 1 ? C + hd : lane == 2 ? 2 * C + tail : lane == 3 ? 2 * C + mid : 2 * C + hd;
                const float val = lane == 3 ? 2.0f : lane == 4 ? 3.0f : 1.0f;
                if (idx >= 0) g[idx] = val;
            }
            if (lane == 0) { short* h = p.hints + 4 * (size_t)(env0 + ts); h[0] = (short)hd; h[1] = 3; h[2] = (short)cell; }
        }
    }
    fence_proxy_async();
    __syncthreads();
    if (p.stats && tid < WURM_STATS_FIELDS) {
        unsigned long long* slot = p.stats + (blockIdx.x % WURM_STATS_SLOTS) * WURM_STATS_FIELDS;
        const int v = tid == 0 ? nvalid : cnt_s[tid - 1];
        if (v) atomicAdd(slot + tid, (unsigned long long)v);
    }
    if (partial) {
        float* dst = p.obs + (size_t)env0 * E;
        const uint32_t obytes = (uint32_t)(nvalid * E) * 4u;
        const bool obs_bulk = (obytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
        if (obs_bulk) {
            if (tid == 0) { bulk_store(dst, stage, obytes); bulk_commit(); bulk_wait_read_all(); }
        } else {
            for (int i = tid; i < nvalid * E; i += blockDim.x) dst[i] = stage[i];
        }
    } else if (p.obs_mode == WURM_OBS_DEFAULT || p.obs_mode == WURM_OBS_ONE_CHANNEL) {
        // single_snake.py:131-151 from the body tile and the known head / food cells (general-path envs have
        // rendered theirs already)
        const int warp = tid >> 5, nwarps = blockDim.x >> 5;
        for (int tt = warp; tt < nvalid; tt += nwarps) {
            if (env_s[3 * tt] == 0) continue;
            const int hcell = env_s[3 * tt + 1], fcell = env_s[3 * tt + 2];
            const float* bt = body_all + (size_t)tt * C;
            for (int q = lane; q < C; q += 32) {
                const int y = div_S(q, p.magic_S), x = q - y * S;
                const bool border = y == 0 || x == 0 || y == S - 1 || x == S - 1;
                const float f = q == fcell ? 1.0f : 0.0f, h = q == hcell ? 1.0f : 0.0f, b = bt[q];
                if (p.obs_mode == WURM_OBS_DEFAULT) {
                    float* o = p.obs + (size_t)(env0 + tt) * 3 * C;
                    o[q] = rgb_channel(f, h, b, border, 0);
                    o[C + q] = rgb_channel(f, h, b, border, 1);
                    o[2 * C + q] = rgb_channel(f, h, b, border, 2);
                } else {
                    float v = (b > kEps ? 1.0f : 0.0f) * 0.5f;
                    v += h * 0.5f;
                    v += f * 1.5f;
                    p.obs[(size_t)(env0 + tt) * C + q] = border ? -1.0f : v;
                }
            }
        }
    }
}

// single_snake.py:322-337 + 344-387: each warp inspects 32 done flags and re-creates the flagged
// envs one after another with coalesced stores; untouched envs cost one byte of traffic.
__global__ void __launch_bounds__(256) single_reset_kernel(const SingleParams p, const uint8_t* __restrict__ done_mask,
                                                           const int32_t* __restrict__ spawn) {
    const int lane = threadIdx.x & 31;
    const int e_base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (e_base >= p.N) return;
    const int e_mine = e_base + lane;
    const bool mine = e_mine < p.N && done_mask[e_mine] != 0;
    unsigned todo = __ballot_sync(0xffffffffu, mine);
    if (todo == 0u) return;
    const int C = p.C, n = 3 * C;
    // every lane draws the layout of ITS env (one Philox evaluation per warp instead of one per finished env)
    int my_tail = 0, my_mid = 0, my_hd = 0, my_cell = -1;
    if (mine) new_env_layout(p, spawn, call_counter(p), e_mine, my_tail, my_mid, my_hd, my_cell);
    while (todo) {
        const int src = __ffs(todo) - 1, e = e_base + src;
        todo &= todo - 1;
        const int tail = __shfl_sync(0xffffffffu, my_tail, src), mid = __shfl_sync(0xffffffffu, my_mid, src);
        const int hd = __shfl_sync(0xffffffffu, my_hd, src), cell = __shfl_sync(0xffffffffu, my_cell, src);
        float* env = p.envs + (size_t)e * n;
        // the new env is zeros plus five cells: 128-bit zero stores over the aligned middle, then (ordered by the
        // warp barrier) the five cells of the LENGTH_3_SNAKES stamp, head and food (:372-385)
        int lead = (4 - (int)((reinterpret_cast<uintptr_t>(env) >> 2) & 3)) & 3;
        if (lead > n) lead = n;
        if (lane < lead) env[lane] = 0.0f;
        const int nvec = (n - lead) >> 2;
        float4* vb = reinterpret_cast<float4*>(env + lead);
        for (int j = lane; j < nvec; j += 32) vb[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const int tail0 = lead + 4 * nvec;
        if (lane < n - tail0) env[tail0 + lane] = 0.0f;
        __syncwarp();
        if (lane < 5) {
            const int idx = lane == 0 ? cell : lane == 1 ? C + hd : lane == 2 ? 2 * C + tail : lane == 3 ? 2 * C + mid : 2 * C + hd;
            const float val = lane == 3 ? 2.0f : lane == 4 ? 3.0f : 1.0f;
            if (idx >= 0) env[idx] = val;
        }
        if (lane == 0 && p.hints) { short* h = p.hints + 4 * (size_t)e; h[0] = (short)hd; h[1] = 3; h[2] = (short)cell; }
    }
}

// wurm/utils.py:113-178 (snake_consistency + env_consistency) as one pass: one warp per env streams its
// 3*S*S floats straight from HBM (coalesced), seven running sums are reduced with shuffles, and the
// verdict of all envs is folded into a 3-word report with atomics.
__global__ void __launch_bounds__(256) single_check_kernel(const float* __restrict__ envs, const uint8_t* __restrict__ skip,
                                                           int N, int C, int* report) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= N || (skip && skip[e])) return;
    const float* env = envs + (size_t)e * 3 * C;
    float sum_food = 0.0f, sum_head = 0.0f, sum_body = 0.0f, max_body = -INFINITY, head_body = 0.0f, head_food = 0.0f;
    int bad_food = 0;
    for (int q = lane; q < C; q += 32) {
        const float f = env[q], h = env[C + q], b = env[2 * C + q];
        bad_food += !(f == 0.0f || f == 1.0f);
        sum_food += f; sum_head += h; sum_body += b;
        max_body = fmaxf(max_body, b);
        head_body += h * b; head_food += h * f;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum_food += __shfl_xor_sync(0xffffffffu, sum_food, o); sum_head += __shfl_xor_sync(0xffffffffu, sum_head, o);
        sum_body += __shfl_xor_sync(0xffffffffu, sum_body, o); head_body += __shfl_xor_sync(0xffffffffu, head_body, o);
        head_food += __shfl_xor_sync(0xffffffffu, head_food, o); bad_food += __shfl_xor_sync(0xffffffffu, bad_food, o);
        max_body = fmaxf(max_body, __shfl_xor_sync(0xffffffffu, max_body, o));
    }
    if (lane == 0) {
        int bits = 0;
        if (bad_food) bits |= WURM_CHK_FOOD_VALUE;
        if (sum_head != 1.0f) bits |= WURM_CHK_HEAD_COUNT;
        if (!(sum_body > 0.0f)) bits |= WURM_CHK_NO_SNAKE;
        if (head_body != max_body) bits |= WURM_CHK_HEAD_NOT_AT_END;
        if ((sqrtf(8.0f * sum_body + 1.0f) - 1.0f) / 2.0f != max_body) bits |= WURM_CHK_BODY_VALUES;
        if (sum_body < 6.0f) bits |= WURM_CHK_TOO_SHORT;
        if (head_food != 0.0f) bits |= WURM_CHK_HEAD_ON_FOOD;
        if (sum_food != 1.0f) bits |= WURM_CHK_FOOD_COUNT;
        if (bits) { atomicOr(report, bits); atomicAdd(report + 1, 1); atomicMin(report + 2, e); }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct SingleLaunch {
    int G, T, threads, blocks, smem;
    bool global;          // the env does not fit a shared-memory tile: single_global_kernel
};

template <bool STEP>
static int launch_global(SingleParams p, cudaStream_t stream) {
    if (p.S > 181) p.hints = nullptr;                   // cell indices no longer fit the int16 hints
    const int warps = 4, E = p.obs_mode == WURM_OBS_PARTIAL ? 3 * p.W * p.W : 0;
    p.stage_bytes = (warps * E * 4 + 15) & ~15;
    const int smem = p.stage_bytes + 16;
    auto kern = single_global_kernel<STEP>;
    static SmemOptIn opt_in;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), &opt_in, smem, false, "cudaFuncSetAttribute(single_global_kernel)")) return rc;
    kern<<<(p.N + warps - 1) / warps, 32 * warps, smem, stream>>>(p);
    return check_launch("single_global_kernel");
}

static int plan_single(const WurmSingleCfg* cfg, SingleParams* p, SingleLaunch* L) {
    if (!cfg) return fail(WURM_E_INVALID, "cfg is NULL");
    const int N = cfg->num_envs, S = cfg->size;
    if (N <= 0) return fail(WURM_E_INVALID, "num_envs must be positive");
    if (S < 9) return fail(WURM_E_INVALID, "size must be >= 9 (reference single_snake.py:346)");
    if (S > 255) return fail(WURM_E_UNSUPPORTED, "size > 255: cell indices no longer fit the 16-bit index arithmetic");
    if (cfg->obs_mode < WURM_OBS_NONE || cfg->obs_mode > WURM_OBS_PARTIAL) return fail(WURM_E_INVALID, "bad obs_mode");
    if (cfg->obs_mode == WURM_OBS_PARTIAL && (cfg->obs_n < 0 || cfg->obs_n > 127)) return fail(WURM_E_INVALID, "bad obs_n");
    const int C = S * S;
    const int W = 2 * cfg->obs_n + 1;
    const int env_bytes = 3 * C * 4;
    const int stage_env_bytes = cfg->obs_mode == WURM_OBS_PARTIAL ? 3 * W * W * 4 : 0;
    int G = 1;
    while (G < 32 && G * 32 < C) G <<= 1;               // ~32 cells per lane
    if (const char* v = getenv("WURM_SINGLE_G")) {       // tuning overrides (power of two <= 32)
        const int g = atoi(v);
        if (g >= 1 && g <= 32 && (g & (g - 1)) == 0) G = g;
    }
    const int per_warp = 32 / G;                        // envs per warp
    // partial observations are staged next to the tile: one-warp CTAs (measured best on B200: 32 threads,
    // ~10 KB, ~20 resident per SM overlapping each other's load / compute / store, no CTA barrier stalls; see
    // profiles/r01_sweep_single_c2*.txt); otherwise up to 256 threads and 64 KB tiles
    int max_threads = stage_env_bytes ? 32 : 256, budget = (stage_env_bytes ? 44 : 64) * 1024;
    if (const char* v = getenv("WURM_SINGLE_THREADS")) max_threads = atoi(v);
    if (const char* v = getenv("WURM_SINGLE_SMEM_KB")) budget = atoi(v) * 1024;
    if (max_threads < 32) max_threads = 32;
    if (max_threads > 256) max_threads = 256;
    int T = max_threads / G;
    if (T * (env_bytes + stage_env_bytes) > budget) T = budget / (env_bytes + stage_env_bytes);
    T = (T / per_warp) * per_warp;
    if (T < per_warp) T = per_warp;
    // prefer a tile whose byte count is a multiple of 16 so the bulk path applies
    if (((size_t)T * env_bytes) % 16 != 0 && T >= 4) T &= ~3;
    if (T < per_warp) T = per_warp;
    const int tile_bytes_padded = (T * env_bytes + 15) & ~15;
    const int stage_bytes = (T * stage_env_bytes + 15) & ~15;
    const int smem = tile_bytes_padded + stage_bytes + 8 + 16;
    L->global = smem > 200 * 1024 || S > 128;             // one env no longer fits a tile: step it on global memory
    p->N = N; p->S = S; p->C = C; p->T = T;
    p->obs_mode = cfg->obs_mode; p->obs_n = cfg->obs_n; p->W = W;
    p->magic_S = (uint32_t)((0x100000000ull + (uint64_t)S - 1) / (uint64_t)S);
    p->magic_W = (uint32_t)((0x100000000ull + (uint64_t)W - 1) / (uint64_t)W);
    p->tile_bytes_padded = tile_bytes_padded;
    p->stage_bytes = stage_bytes;
    L->G = G; L->T = T; L->threads = T * G; L->blocks = (N + T - 1) / T; L->smem = smem;
    return WURM_OK;
}

template <int G, bool STEP>
static int launch_tile(const SingleParams& p, const SingleLaunch& L, cudaStream_t stream) {
    auto kern = single_tile_kernel<G, STEP>;
    static SmemOptIn opt_in;                           // per instantiation, per device inside
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), &opt_in, L.smem, true, "cudaFuncSetAttribute(single_tile_kernel)")) return rc;
    kern<<<L.blocks, L.threads, L.smem, stream>>>(p);
    return check_launch("single_tile_kernel");
}

template <bool STEP>
static int dispatch_tile(const SingleParams& p, const SingleLaunch& L, cudaStream_t stream) {
    switch (L.G) {
        case 1: return launch_tile<1, STEP>(p, L, stream);
        case 2: return launch_tile<2, STEP>(p, L, stream);
        case 4: return launch_tile<4, STEP>(p, L, stream);
        case 8: return launch_tile<8, STEP>(p, L, stream);
        case 16: return launch_tile<16, STEP>(p, L, stream);
        default: return launch_tile<32, STEP>(p, L, stream);
    }
}

static bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

template <int G>
static int launch_body_t(const SingleParams& p, int blocks, int threads, int smem, cudaStream_t stream) {
    auto kern = single_body_kernel<G>;
    static SmemOptIn opt_in;                           // per instantiation, per device inside
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), &opt_in, smem, true, "cudaFuncSetAttribute(single_body_kernel)")) return rc;
    kern<<<blocks, threads, smem, stream>>>(p);
    return check_launch("single_body_kernel");
}

// The body-only path (single_body_kernel) applies to steps with hints, an even grid side from 16 up (a channel is then
// a 16-byte-aligned span worth skipping) and an observation that can be rendered from the body tile; -1 if it does not.
static int try_launch_body(SingleParams p, int G, cudaStream_t stream) {
    const int min_size = getenv("WURM_SINGLE_BODY_MIN_SIZE") ? atoi(getenv("WURM_SINGLE_BODY_MIN_SIZE")) : 16;
    if (!p.hints || (p.S & 1) || p.S < min_size || !aligned16(p.envs) || getenv("WURM_SINGLE_NO_BODY_PATH")) return -1;
    if (p.obs_mode != WURM_OBS_PARTIAL && p.obs_mode != WURM_OBS_NONE && p.obs_mode != WURM_OBS_DEFAULT &&
        p.obs_mode != WURM_OBS_ONE_CHANNEL)
        return -1;
    // measured (profiles/r01_body_path_sizes.txt): with a full-grid observation the path wins from size 16 up; with a
    // partial (or no) observation it wins while several envs share a warp (sizes 16-20) and again once the env is large
    // (size 36: 15.5 KB), but loses in between (size 24: one 2.3 KB env per warp and two DRAM round trips)
    if ((p.obs_mode == WURM_OBS_PARTIAL || p.obs_mode == WURM_OBS_NONE) && p.S > 20 && p.S < 32 &&
        !getenv("WURM_SINGLE_BODY_MIN_SIZE"))
        return -1;
    const int per_warp = 32 / G;
    const int chan_bytes = p.C * 4, stage_env_bytes = p.obs_mode == WURM_OBS_PARTIAL ? 3 * p.W * p.W * 4 : 0;
    int max_threads = stage_env_bytes ? 32 : WURM_BODY_THREADS;
    int T = max_threads / G;
    const int budget = 48 * 1024;
    if (T * (chan_bytes + stage_env_bytes) > budget) T = budget / (chan_bytes + stage_env_bytes);
    T = (T / per_warp) * per_warp;
    if (T < per_warp) T = per_warp;
    if (T > 64) T = 64;
    p.T = T;
    p.tile_bytes_padded = (T * chan_bytes + 15) & ~15;
    p.stage_bytes = (T * stage_env_bytes + 15) & ~15;
    const int smem = p.tile_bytes_padded + p.stage_bytes + 8 + 16 + 3 * T * 4 + 16;
    if (smem > 200 * 1024) return -1;
    const int blocks = (p.N + T - 1) / T, threads = T * G;
    switch (G) {
        case 4: return launch_body_t<4>(p, blocks, threads, smem, stream);
        case 8: return launch_body_t<8>(p, blocks, threads, smem, stream);
        case 16: return launch_body_t<16>(p, blocks, threads, smem, stream);
        case 32: return launch_body_t<32>(p, blocks, threads, smem, stream);
        default: return -1;
    }
}

template <int G, int STAGES>
static int launch_ring_t(const SingleParams& p, int ntiles, int ctas, int smem, cudaStream_t stream) {
    auto kern = single_ring_kernel<G, STAGES>;
    static SmemOptIn opt_in;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), &opt_in, smem, true, "cudaFuncSetAttribute(single_ring_kernel)")) return rc;
    kern<<<ctas, 32, smem, stream>>>(p, ntiles);
    return check_launch("single_ring_kernel");
}

// -1 if the experiment is not selected or does not apply to this call
static int try_launch_ring(const SingleParams& p, const SingleLaunch& L, cudaStream_t stream) {
    const char* v = getenv("WURM_SINGLE_RING");
    if (!v || p.obs_mode != WURM_OBS_PARTIAL || !p.bulk_ok || L.threads != 32 || L.G != 4 || (p.N % p.T) != 0) return -1;
    if ((reinterpret_cast<uintptr_t>(p.obs) & 15u) || ((size_t)p.T * 3 * p.W * p.W * 4) % 16) return -1;
    const int stages = atoi(v);
    const int per_sm = getenv("WURM_SINGLE_RING_CTAS") ? atoi(getenv("WURM_SINGLE_RING_CTAS")) : 12;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = p.N / p.T;
    const int ctas = min(ntiles, sms * per_sm);
    const int smem = stages * p.tile_bytes_padded + 2 * p.stage_bytes + 8 * stages + 16 + 16;
    switch (stages) {
        case 2: return launch_ring_t<4, 2>(p, ntiles, ctas, smem, stream);
        case 3: return launch_ring_t<4, 3>(p, ntiles, ctas, smem, stream);
        case 4: return launch_ring_t<4, 4>(p, ntiles, ctas, smem, stream);
        default: return -1;
    }
}

}  // namespace wurm

using namespace wurm;

extern "C" int64_t wurm_single_obs_elems(const WurmSingleCfg* cfg) {
    if (!cfg) return -1;
    const int64_t C = (int64_t)cfg->size * cfg->size, W = 2 * cfg->obs_n + 1;
    switch (cfg->obs_mode) {
        case WURM_OBS_DEFAULT:
        case WURM_OBS_RAW: return 3 * C;
        case WURM_OBS_ONE_CHANNEL: return C;
        case WURM_OBS_POSITIONS: return 4;
        case WURM_OBS_PARTIAL: return 3 * W * W;
        default: return 0;
    }
}

static int single_step_impl(const WurmSingleCfg* cfg, float* envs, void* actions, int action_bytes,
                            const int32_t* food_cell_replay, int auto_reset, const int32_t* spawn_replay, uint64_t seed,
                            uint64_t step, const uint64_t* step_dev, float* obs, float* reward, uint8_t* done,
                            uint8_t* self_col, uint8_t* edge_col, int32_t* status, int64_t* stats, int16_t* hints,
                            uint8_t* packed, void* stream) {
    SingleParams p = {};
    SingleLaunch L;
    if (int rc = plan_single(cfg, &p, &L)) return rc;
    if (!envs || !actions || !reward || !done || !self_col || !edge_col || !status) return fail(WURM_E_INVALID, "NULL pointer");
    if (!valid_action_bytes(action_bytes)) return fail(WURM_E_INVALID, "action_bytes must be 1, 2, 4 or 8");
    if (cfg->obs_mode != WURM_OBS_NONE && !obs) return fail(WURM_E_INVALID, "obs is NULL");
    p.packed = packed;
    p.envs = envs; p.actions = actions; p.action_bytes = action_bytes; p.food_replay = food_cell_replay;
    p.auto_reset = auto_reset; p.spawn = spawn_replay; p.hints = hints;
    p.seed = seed; p.step = step; p.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
    p.obs = obs; p.reward = reward; p.done = done; p.self_col = self_col;
    p.edge_col = edge_col; p.status = status; p.stats = reinterpret_cast<unsigned long long*>(stats);
    p.bulk_ok = aligned16(envs) && ((size_t)p.T * 3 * p.C * 4) % 16 == 0 && (cfg->obs_mode != WURM_OBS_RAW || aligned16(obs));
    if (L.global) return launch_global<true>(p, (cudaStream_t)stream);
    {
        const int rc = try_launch_body(p, L.G, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    {
        const int rc = try_launch_ring(p, L, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    return dispatch_tile<true>(p, L, (cudaStream_t)stream);
}

extern "C" int wurm_single_step(const WurmSingleCfg* cfg, float* envs, void* actions, int action_bytes,
                                const int32_t* food_cell_replay, uint64_t seed, uint64_t step, const uint64_t* step_dev,
                                float* obs, float* reward, uint8_t* done, uint8_t* self_col, uint8_t* edge_col,
                                int32_t* status, int64_t* stats, int16_t* hints, uint8_t* packed, void* stream) {
    return single_step_impl(cfg, envs, actions, action_bytes, food_cell_replay, 0, nullptr, seed, step, step_dev, obs, reward,
                            done, self_col, edge_col, status, stats, hints, packed, stream);
}

extern "C" int wurm_single_step_reset(const WurmSingleCfg* cfg, float* envs, void* actions, int action_bytes,
                                      const int32_t* food_cell_replay, const int32_t* spawn_replay, uint64_t seed,
                                      uint64_t step, const uint64_t* step_dev, float* obs, float* reward, uint8_t* done,
                                      uint8_t* self_col, uint8_t* edge_col, int32_t* status, int64_t* stats, int16_t* hints,
                                      uint8_t* packed, void* stream) {
    return single_step_impl(cfg, envs, actions, action_bytes, food_cell_replay, 1, spawn_replay, seed, step, step_dev, obs,
                            reward, done, self_col, edge_col, status, stats, hints, packed, stream);
}

extern "C" int wurm_single_observe(const WurmSingleCfg* cfg, const float* envs, float* obs, int32_t* status, void* stream) {
    SingleParams p = {};
    SingleLaunch L;
    if (int rc = plan_single(cfg, &p, &L)) return rc;
    if (!envs || !obs || !status) return fail(WURM_E_INVALID, "NULL pointer");
    if (cfg->obs_mode == WURM_OBS_NONE) return WURM_OK;
    p.envs = const_cast<float*>(envs); p.obs = obs; p.status = status;
    p.bulk_ok = aligned16(envs) && ((size_t)p.T * 3 * p.C * 4) % 16 == 0 && (cfg->obs_mode != WURM_OBS_RAW || aligned16(obs));
    if (L.global) return launch_global<false>(p, (cudaStream_t)stream);
    return dispatch_tile<false>(p, L, (cudaStream_t)stream);
}

extern "C" int wurm_single_reset(const WurmSingleCfg* cfg, float* envs, const uint8_t* done_mask, const int32_t* spawn_replay,
                                 uint64_t seed, uint64_t step, const uint64_t* step_dev, int16_t* hints, void* stream) {
    SingleParams p = {};
    SingleLaunch L;
    if (int rc = plan_single(cfg, &p, &L)) return rc;
    if (!envs || !done_mask) return fail(WURM_E_INVALID, "NULL pointer");
    p.hints = hints;
    p.envs = envs; p.seed = seed; p.step = step; p.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
    const int warps_per_block = 8;
    const int blocks = (p.N + 32 * warps_per_block - 1) / (32 * warps_per_block);
    single_reset_kernel<<<blocks, 32 * warps_per_block, 0, (cudaStream_t)stream>>>(p, done_mask, spawn_replay);
    return check_launch("single_reset_kernel");
}

extern "C" int wurm_single_check(const WurmSingleCfg* cfg, const float* envs, const uint8_t* skip, int32_t* report,
                                 void* stream) {
    if (!cfg || !envs || !report) return fail(WURM_E_INVALID, "NULL pointer");
    if (cfg->num_envs <= 0 || cfg->size < 1) return fail(WURM_E_INVALID, "bad num_envs / size");
    const int warps = 8, blocks = (cfg->num_envs + warps - 1) / warps;
    single_check_kernel<<<blocks, 32 * warps, 0, (cudaStream_t)stream>>>(envs, skip, cfg->num_envs, cfg->size * cfg->size, report);
    return check_launch("single_check_kernel");
}
