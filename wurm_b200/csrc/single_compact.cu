// SingleSnake with a COMPACT RESIDENT STATE for B200 (sm_100a).
//
// The reference keeps an env as (3,S,S) fp32 -- 972 bytes at size 9 of which ~5 cells are non-zero -- and the dense
// kernels of single_snake.cu honour that layout because callers may read and write the tensor between calls.  A caller
// that does not (state='compact' in the Python class) can let the env live in HBM in this file's form instead:
//
//   cells (N, Cp) uint16, Cp = S*S rounded up to a multiple of 8 (rows are 16-byte aligned):
//         bits 0-13 body value, bit 14 head, bit 15 food
//   aux   (N, 4)  int16: head cell (-1 none), snake size (largest body value), neck cell (the cell holding size-1),
//         flags (bit 0: CANONICAL -- exactly one head, sitting on the only cell that holds `size`, and exactly one cell
//         holding size-1; what every env is unless its snake has just died)
//
// 176 bytes per env at size 9.  One CTA = one tile of consecutive envs: the tile's records arrive by ONE bulk copy (TMA,
// UBLKCP), are stepped in shared memory and leave by one bulk store; partial observations are rendered from the
// records into a shared staging area and leave by a second bulk store.  No per-cell traffic to HBM at all.
//
// A canonical env takes the FAST step: orientation from (head, neck), the head move, the decay of every body cell as a
// SIMD-in-word decrement of two 14-bit fields per 32-bit word, collisions as look-ups at the new head cell; the aux
// vector of the next call follows in closed form (new neck = old head).  Everything else -- dead snakes stepped again,
// states folded in from caller-edited tensors -- takes the GENERAL step: the records of that one env are expanded into
// an fp32 scratch env in shared memory and the dense path's own step_env (single_device.cuh) runs on it, executed by
// the whole warp, before the result is folded back and its aux vector re-derived by a scan.  So the compact path is
// bit-identical to the dense one by construction on everything the records can carry (integral body values < 16384,
// food and head values 0 / 1), and wurm_single_compact refuses anything else.
#include <stdlib.h>

#include "../../include/wurm_b200.h"
#include "common.cuh"
#include "host_util.h"
#include "single_device.cuh"

namespace wurm {

constexpr uint32_t kBodyMask = 0x3fffu, kHeadBit = 0x4000u, kFoodBit = 0x8000u;
constexpr int kCanonical = 1;

struct CompactParams {
    SingleParams sp;          // outputs, actions, draws, geometry (sp.envs unused; sp.hints unused)
    uint16_t* cells;
    short* aux;
    int Cp;                   // row pitch in cells
    int scratch_floats;       // 3*C: one fp32 scratch env per warp for the general step
};

__device__ __forceinline__ float cell_food(uint32_t v) { return (v & kFoodBit) ? 1.0f : 0.0f; }
__device__ __forceinline__ float cell_head(uint32_t v) { return (v & kHeadBit) ? 1.0f : 0.0f; }
__device__ __forceinline__ float cell_body(uint32_t v) { return (float)(v & kBodyMask); }

// aux vector of one env from its records, by a group of G lanes (wurm/utils.py:36-65's inputs: head cell, the cells
// holding the two largest body values)
template <int G>
__device__ __forceinline__ short4 derive_aux(const uint16_t* row, int C, int l, unsigned gm) {
    int hp = -1, hc = 0, m = 0;
    for (int q = l; q < C; q += G) {
        const uint32_t v = row[q];
        if (v & kHeadBit) { hp = q; ++hc; }
        m = max(m, (int)(v & kBodyMask));
    }
    hp = group_max<G>(hp, gm); hc = group_sum<G>(hc, gm);
    const int size = group_max<G>(m, gm);
    int c1 = 0, c2 = 0, p1 = -1, p2 = -1;
    for (int q = l; q < C; q += G) {
        const int b = (int)(row[q] & kBodyMask);
        if (b == size) { ++c1; p1 = q; }
        if (b == size - 1) { ++c2; p2 = q; }
    }
    c1 = group_sum<G>(c1, gm); c2 = group_sum<G>(c2, gm);
    p1 = group_max<G>(p1, gm); p2 = group_max<G>(p2, gm);
    const bool canonical = hc == 1 && c1 == 1 && c2 == 1 && p1 == hp && size >= 2;
    short4 a;
    a.x = (short)(hc >= 1 ? hp : -1); a.y = (short)size; a.z = (short)(canonical ? p2 : -1); a.w = (short)(canonical ? kCanonical : 0);
    return a;
}

// the five cells of a freshly created env (:372-385) written over a zeroed row; returns its aux vector
__device__ __forceinline__ short4 stamp_new_env(uint16_t* row, int tail, int mid, int hd, int cell) {
    if (cell >= 0) row[cell] = (uint16_t)kFoodBit;
    row[tail] = 1; row[mid] = 2; row[hd] = (uint16_t)(kHeadBit | 3u);
    short4 a;
    a.x = (short)hd; a.y = 3; a.z = (short)mid; a.w = kCanonical;
    return a;
}

// single_snake.py:130-195 for one env, rendered from its records by a group of G lanes.  partial_n goes to `stage`
// (shared), every other mode straight to the caller's buffer.
template <int G>
__device__ __forceinline__ void observe_env(const SingleParams& p, const uint16_t* row, int e, int hp_known, float* stage, int l,
                                            unsigned gm) {
    const int S = p.S, C = p.C;
    if (p.obs_mode == WURM_OBS_PARTIAL) {                            // :166-193
        const int W = p.W, WW = W * W, n = p.obs_n;
        if (hp_known < 0) {                                          // the reference raises here (:191)
            for (int r = l; r < 3 * WW; r += G) stage[r] = 0.0f;
            if (l == 0) atomicOr(p.status, WURM_ST_NO_HEAD_PARTIAL);
            return;
        }
        constexpr float kHalf = 127.0f / 255.0f;
        const int hy = div_S(hp_known, p.magic_S), hx = hp_known - hy * S;
        for (int ij = l; ij < WW; ij += G) {
            const int i = (int)__umulhi((uint32_t)ij, p.magic_W), j = ij - i * W;
            const int y = hy - n + i, x = hx - n + j;
            float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
            if ((unsigned)(y - 1) < (unsigned)(S - 2) && (unsigned)(x - 1) < (unsigned)(S - 2)) {
                const uint32_t v = row[y * S + x];
                v0 = v1 = v2 = 1.0f;
                if (v & kBodyMask) { v0 = 0.0f; v1 = kHalf; v2 = 0.0f; }
                if (v & kHeadBit) { v0 = 0.0f; v1 = 1.0f; v2 = 0.0f; }
                if (v & kFoodBit) { v0 = 1.0f; v1 = 0.0f; v2 = 0.0f; }
            }
            stage[ij] = v0; stage[WW + ij] = v1; stage[2 * WW + ij] = v2;
        }
    } else if (p.obs_mode == WURM_OBS_DEFAULT || p.obs_mode == WURM_OBS_ONE_CHANNEL || p.obs_mode == WURM_OBS_RAW) {
        for (int q = l; q < C; q += G) {
            const uint32_t v = row[q];
            const int y = div_S(q, p.magic_S), x = q - y * S;
            const bool border = y == 0 || x == 0 || y == S - 1 || x == S - 1;
            const float f = cell_food(v), h = cell_head(v), b = cell_body(v);
            if (p.obs_mode == WURM_OBS_DEFAULT) {                    // :131-138
                float* o = p.obs + (size_t)e * 3 * C;
                o[q] = rgb_channel(f, h, b, border, 0); o[C + q] = rgb_channel(f, h, b, border, 1); o[2 * C + q] = rgb_channel(f, h, b, border, 2);
            } else if (p.obs_mode == WURM_OBS_ONE_CHANNEL) {         // :142-151
                float w = (b > kEps ? 1.0f : 0.0f) * 0.5f;
                w += h * 0.5f;
                w += f * 1.5f;
                p.obs[(size_t)e * C + q] = border ? -1.0f : w;
            } else {                                                 // :139-141 raw: a copy of the state
                float* o = p.obs + (size_t)e * 3 * C;
                o[q] = f; o[C + q] = h; o[2 * C + q] = b;
            }
        }
    } else if (p.obs_mode == WURM_OBS_POSITIONS) {                   // :152-165 first argmax of head / food
        int first_head = 1 << 30, first_food = 1 << 30;
        for (int q = l; q < C; q += G) {
            const uint32_t v = row[q];
            if ((v & kHeadBit) && q < first_head) first_head = q;
            if ((v & kFoodBit) && q < first_food) first_food = q;
        }
        first_head = -group_max<G>(-first_head, gm); first_food = -group_max<G>(-first_food, gm);
        if (first_head == (1 << 30)) first_head = 0;                 // argmax of an all-zero channel
        if (first_food == (1 << 30)) first_food = 0;
        if (l == 0) {
            float* o = p.obs + (size_t)e * 4;
            const int hy = div_S(first_head, p.magic_S), fy = div_S(first_food, p.magic_S);
            o[0] = (float)hy; o[1] = (float)(first_head - hy * S); o[2] = (float)fy; o[3] = (float)(first_food - fy * S);
        }
    }
}

// uniform choice among the free interior cells, ranked with ballots (the fallback after kRejectionTries misses)
template <int G>
__device__ __noinline__ int pick_free_record(const uint16_t* row, int S, int C, uint32_t magic, uint32_t rnd, int l, unsigned gm) {
    const unsigned shift = (G == 32) ? 0u : ((threadIdx.x & 31u) & ~(unsigned)(G - 1));
    const unsigned low = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    auto free_bits = [&](int base) {
        const int q = base + l;
        bool f = false;
        if (q < C) {
            const int y = div_S(q, magic), x = q - y * S;
            f = y >= 1 && y <= S - 2 && x >= 1 && x <= S - 2 && row[q] == 0;
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

// The fast step of one canonical env on its records (single_snake.py:197-300), by a group of G lanes.
// Returns the new head cell; `aux` becomes the env's aux vector after the step.
template <int G>
__device__ __forceinline__ int fast_step(const SingleParams& p, uint16_t* row, int Cp, int e, int l, int* cnt_s, long long a_in,
                                         short4& aux, bool& ended) {
    const unsigned gm = group_mask<G>();
    const int S = p.S, C = p.C;
    const int hp = aux.x, size = aux.y, neck = aux.z;
    int k = 0;                                                       // orientation (:212): head = neck + OFF[k]
    {
        const int d = hp - neck;
        const int x2 = neck - div_S(neck, p.magic_S) * S;
        if (d == -S) k = 0;
        else if (d == 1 && x2 != S - 1) k = 1;
        else if (d == S) k = 2;
        else if (d == -1 && x2 != 0) k = 3;
    }
    const long long a = (a_in + ((long long)k == a_in ? 2 : 0)) % 4;     // :221-222
    if (l == 0 && a != a_in) store_action(p.actions, p.action_bytes, (size_t)e, a);
    int np = -1, ny = -1, nx = -1;                                   // :225-233
    {
        const int hy = div_S(hp, p.magic_S), hx = hp - hy * S;
        ny = hy - (a >= 0 ? off_y((int)a) : 0);
        nx = hx - (a >= 0 ? off_x((int)a) : 0);
        if (ny >= 0 && ny < S && nx >= 0 && nx < S) np = ny * S + nx;
    }
    const int ov = (np >= 0 && (row[np] & kFoodBit)) ? 1 : 0;        // :242
    if (!ov) {                                                       // :246-249 every body cell - 1: two cells per word, eight per load
        uint4* r4 = reinterpret_cast<uint4*>(row);
        auto dec = [](uint32_t w) {                                  // per 16-bit half: body field - 1 where it is non-zero
            const uint32_t x = w & 0x3fff3fffu;
            return w - (((x + 0x3fff3fffu) >> 14) & 0x00010001u);
        };
        for (int j = l; j < (Cp >> 3); j += G) {
            uint4 v = r4[j];
            if ((v.x | v.y | v.z | v.w) & 0x3fff3fffu) {
                v.x = dec(v.x); v.y = dec(v.y); v.z = dec(v.z); v.w = dec(v.w);
                r4[j] = v;
            }
        }
    }
    __syncwarp(gm);
    const bool sc = np >= 0 && (row[np] & kBodyMask) != 0;           // :252
    const bool interior = np >= 0 && ny >= 1 && ny <= S - 2 && nx >= 1 && nx <= S - 2;
    __syncwarp(gm);
    if (l == 0) {
        row[hp] = (uint16_t)(row[hp] & ~kHeadBit);
        if (np >= 0) row[np] = (uint16_t)(kHeadBit | ((row[np] & kBodyMask) + (uint32_t)(size + ov)));   // :258-272 (an eaten food is gone)
    }
    __syncwarp(gm);
    if (ov) {                                                        // :277-282 respawn
        int cell;
        if (p.food_replay) cell = p.food_replay[e];
        else {
            const int I = S - 2;
            cell = -1;
            for (uint32_t t = 0; t < kRejectionTries && cell < 0; ++t) {
                const int cand = (int)bounded(draw_i(p.seed, call_counter(p), (uint32_t)e, kStreamSingleStepFood, t), (uint32_t)(I * I));
                const int cy = cand / I, q = (1 + cy) * S + 1 + (cand - cy * I);
                if (row[q] == 0) cell = q;
            }
            if (cell < 0)
                cell = pick_free_record<G>(row, S, C, p.magic_S,
                                           draw_i(p.seed, call_counter(p), (uint32_t)e, kStreamSingleStepFood, kRejectionTries), l, gm);
        }
        if (l == 0 && cell >= 0) row[cell] = (uint16_t)(row[cell] | kFoodBit);
    }
    if (l == 0) {
        p.reward[e] = (float)ov;                                     // :271
        p.self_col[e] = sc;
        p.edge_col[e] = !interior;                                   // :290-293
        p.done[e] = sc || !interior;
        if (p.packed) p.packed[e] = pack_result(sc || !interior, sc, !interior, (float)ov);
        if (p.stats) {
            if (sc || !interior) atomicAdd(cnt_s + 0, 1);
            if (ov) atomicAdd(cnt_s + 1, ov);
            if (sc) atomicAdd(cnt_s + 2, 1);
            if (!interior) atomicAdd(cnt_s + 3, 1);
        }
    }
    // after the step the only cell holding size' = size + ov is the new head cell, and the old head cell is the only one
    // holding size' - 1 -- unless the head ran into a body (their values added up) or left the grid
    const bool canonical = np >= 0 && !sc;
    aux.x = (short)np; aux.y = (short)(size + ov); aux.z = (short)(canonical ? hp : -1); aux.w = (short)(canonical ? kCanonical : 0);
    __syncwarp(gm);
    ended = sc || !interior;
    return np;
}

// One CTA = one tile of T envs, G lanes per env.  STEP: load -> step (-> reset) -> store, observe.  !STEP: observe.
template <int G, bool STEP>
__global__ void __launch_bounds__(256) single_compact_kernel(const CompactParams cp) {
    const SingleParams& p = cp.sp;
    extern __shared__ __align__(128) unsigned char smem[];
    const int Cp = cp.Cp, C = p.C;
    const int row_bytes = Cp * 2;
    uint16_t* tile = reinterpret_cast<uint16_t*>(smem);
    float* stage = reinterpret_cast<float*>(smem + p.tile_bytes_padded);
    float* scratch_all = reinterpret_cast<float*>(smem + p.tile_bytes_padded + p.stage_bytes);
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(scratch_all + (size_t)(STEP ? nwarps : 0) * cp.scratch_floats);
    int* cnt_s = reinterpret_cast<int*>(bar + 1);
    const unsigned gm = group_mask<G>();

    const int env0 = blockIdx.x * p.T, nvalid = min(p.T, p.N - env0);
    const uint32_t bytes = (uint32_t)(nvalid * row_bytes);
    const bool partial = p.obs_mode == WURM_OBS_PARTIAL;
    const int E = 3 * p.W * p.W;
    uint16_t* gtile = cp.cells + (size_t)env0 * Cp;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) __syncwarp();     // (orders the barrier's initialisation before its first use for racecheck's warp-level model)
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        bulk_load(tile, gtile, bytes, bar);
    }
    if (threadIdx.x < 4) cnt_s[threadIdx.x] = 0;
    const int t = threadIdx.x / G, l = threadIdx.x % G;
    const bool valid = t < nvalid;
    const int e = env0 + (valid ? t : 0);
    long long a_in = 0;
    short4 aux = make_short4(-1, 0, -1, 0);
    if (valid) {
        if (STEP) a_in = load_action(p.actions, p.action_bytes, (size_t)e);
        aux = *reinterpret_cast<const short4*>(cp.aux + 4 * (size_t)e);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    uint16_t* row = tile + (size_t)t * Cp;
    float* my_stage = stage + (size_t)t * E;
    bool ended = false, slow = false;
    if (valid) {
        if (STEP) {
            if (aux.w & kCanonical) {
                const int np = fast_step<G>(p, row, Cp, e, l, cnt_s, a_in, aux, ended);
                observe_env<G>(p, row, e, np, my_stage, l, gm);
            } else {
                slow = true;
            }
        } else {
            observe_env<G>(p, row, e, aux.x, my_stage, l, gm);
        }
    }
    if (STEP) {
        // The general step, for the (rare) envs that are not canonical: one at a time, by the whole warp, on an fp32
        // expansion of the env's records in this warp's scratch -- the dense path's own step_env.
        __syncwarp();
        unsigned todo = __ballot_sync(0xffffffffu, slow && l == 0);
        float* scratch = scratch_all + (size_t)warp * cp.scratch_floats;
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int ts = ((threadIdx.x & ~31) + src) / G, es = env0 + ts;
            const long long a_s = __shfl_sync(0xffffffffu, a_in, src);
            uint16_t* srow = tile + (size_t)ts * Cp;
            for (int q = lane; q < C; q += 32) {
                const uint32_t v = srow[q];
                scratch[q] = cell_food(v); scratch[C + q] = cell_head(v); scratch[2 * C + q] = cell_body(v);
            }
            __syncwarp();
            SingleParams q = p;                                      // (by value: this is the slow path)
            q.hints = nullptr; q.envs = nullptr;
            bool s_ended = false;
            const int np = step_env<32, false>(q, scratch, es, lane, cnt_s, a_s, -1, -1, s_ended);
            __syncwarp();
            for (int qq = lane; qq < C; qq += 32) {
                const float f = scratch[qq], h = scratch[C + qq], b = scratch[2 * C + qq];
                uint32_t v = (uint32_t)(int)b & kBodyMask;
                if (h != 0.0f) v |= kHeadBit;
                if (f != 0.0f) v |= kFoodBit;
                srow[qq] = (uint16_t)v;
            }
            __syncwarp();
            const short4 s_aux = derive_aux<32>(srow, C, lane, 0xffffffffu);
            observe_env<32>(p, srow, es, np, stage + (size_t)ts * E, lane, 0xffffffffu);
            if (lane == src) { aux = s_aux; ended = s_ended; }
            __syncwarp();
        }
        ended = __shfl_sync(0xffffffffu, (int)ended, lane - l) != 0;     // the group's leader knows (general step: only it does)
        if (p.auto_reset && valid && ended) {
            // Fused reset (:322-337): the observation above showed the terminal state (what the reference's driver feeds
            // its policy next, main.py:227); now the env's records are replaced by a fresh env's.
            __syncwarp(gm);
            int tail = 0, mid = 0, hd = 0, cell = -1;
            if (l == 0) new_env_layout(p, p.spawn, call_counter(p) + 1, e, tail, mid, hd, cell);
            uint4* r4 = reinterpret_cast<uint4*>(row);
            for (int j = l; j < (Cp >> 3); j += G) r4[j] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp(gm);
            if (l == 0) aux = stamp_new_env(row, tail, mid, hd, cell);
        }
        if (valid && l == 0) *reinterpret_cast<short4*>(cp.aux + 4 * (size_t)e) = aux;
    }
    fence_proxy_async();                    // generic-proxy writes -> visible to the bulk stores
    __syncthreads();
    if (STEP && p.stats && threadIdx.x < WURM_STATS_FIELDS) {
        unsigned long long* slot = p.stats + (blockIdx.x % WURM_STATS_SLOTS) * WURM_STATS_FIELDS;
        const int v = threadIdx.x == 0 ? nvalid : cnt_s[threadIdx.x - 1];
        if (v) atomicAdd(slot + threadIdx.x, (unsigned long long)v);
    }
    bool stored = false;
    if (threadIdx.x == 0) {
        if (STEP) { bulk_store(gtile, tile, bytes); stored = true; }
    }
    if (partial) {
        float* dst = p.obs + (size_t)env0 * E;
        const uint32_t obytes = (uint32_t)(nvalid * E) * 4u;
        if ((obytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
            if (threadIdx.x == 0) { bulk_store(dst, stage, obytes); stored = true; }
        } else {
            for (int i = threadIdx.x; i < nvalid * E; i += blockDim.x) dst[i] = stage[i];
        }
    }
    if (stored) {
        bulk_commit();
        bulk_wait_read_all();               // shared memory must outlive the bulk stores' reads
    }
}

// single_snake.py:322-337 + 344-387 on records: each warp inspects 32 done flags and re-creates the flagged envs.
__global__ void __launch_bounds__(256) single_compact_reset_kernel(const CompactParams cp, const uint8_t* __restrict__ done_mask,
                                                                   const int32_t* __restrict__ spawn) {
    const SingleParams& p = cp.sp;
    const int lane = threadIdx.x & 31;
    const int e_base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (e_base >= p.N) return;
    const int e_mine = e_base + lane;
    const bool mine = e_mine < p.N && done_mask[e_mine] != 0;
    unsigned todo = __ballot_sync(0xffffffffu, mine);
    if (todo == 0u) return;
    int my_tail = 0, my_mid = 0, my_hd = 0, my_cell = -1;
    if (mine) new_env_layout(p, spawn, call_counter(p), e_mine, my_tail, my_mid, my_hd, my_cell);
    while (todo) {
        const int src = __ffs(todo) - 1, e = e_base + src;
        todo &= todo - 1;
        const int tail = __shfl_sync(0xffffffffu, my_tail, src), mid = __shfl_sync(0xffffffffu, my_mid, src);
        const int hd = __shfl_sync(0xffffffffu, my_hd, src), cell = __shfl_sync(0xffffffffu, my_cell, src);
        uint16_t* row = cp.cells + (size_t)e * cp.Cp;
        uint4* r4 = reinterpret_cast<uint4*>(row);
        for (int j = lane; j < (cp.Cp >> 3); j += 32) r4[j] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        if (lane == 0) *reinterpret_cast<short4*>(cp.aux + 4 * (size_t)e) = stamp_new_env(row, tail, mid, hd, cell);
    }
}

// fp32 (N,3,S,S) <-> records, one warp per env.  TO_COMPACT validates: a value the records cannot carry exactly raises
// WURM_ST_NOT_COMPACT.
template <bool TO_COMPACT>
__global__ void __launch_bounds__(256) single_convert_kernel(const CompactParams cp, float* envs) {
    const SingleParams& p = cp.sp;
    const int lane = threadIdx.x & 31, C = p.C;
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= p.N) return;
    float* env = envs + (size_t)e * 3 * C;
    uint16_t* row = cp.cells + (size_t)e * cp.Cp;
    if (TO_COMPACT) {
        bool bad = false;
        for (int q = lane; q < cp.Cp; q += 32) {
            uint32_t v = 0;
            if (q < C) {
                const float f = env[q], h = env[C + q], b = env[2 * C + q];
                const int bi = (int)b;
                bad |= !(f == 0.0f || f == 1.0f) || !(h == 0.0f || h == 1.0f) || (float)bi != b || bi < 0 || bi > (int)kBodyMask;
                v = ((uint32_t)bi & kBodyMask) | (h != 0.0f ? kHeadBit : 0u) | (f != 0.0f ? kFoodBit : 0u);
            }
            row[q] = (uint16_t)v;
        }
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.status, WURM_ST_NOT_COMPACT);
        __syncwarp();
        const short4 a = derive_aux<32>(row, C, lane, 0xffffffffu);
        if (lane == 0) *reinterpret_cast<short4*>(cp.aux + 4 * (size_t)e) = a;
    } else {
        for (int q = lane; q < C; q += 32) {
            const uint32_t v = row[q];
            env[q] = cell_food(v); env[C + q] = cell_head(v); env[2 * C + q] = cell_body(v);
        }
    }
}

// wurm/utils.py:113-178 (snake_consistency + env_consistency) on records: one warp per env, the same seven sums as
// single_check_kernel (single_snake.cu) taken from the decoded cells, the same verdict bits in the same report.
__global__ void __launch_bounds__(256) single_compact_check_kernel(const uint16_t* __restrict__ cells, const uint8_t* __restrict__ skip,
                                                                   int N, int C, int Cp, int* report) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= N || (skip && skip[e])) return;
    const uint16_t* row = cells + (size_t)e * Cp;
    float sum_food = 0.0f, sum_head = 0.0f, sum_body = 0.0f, max_body = -INFINITY, head_body = 0.0f, head_food = 0.0f;
    for (int q = lane; q < C; q += 32) {
        const uint32_t v = row[q];
        const float f = cell_food(v), h = cell_head(v), b = cell_body(v);
        sum_food += f; sum_head += h; sum_body += b;
        max_body = fmaxf(max_body, b);
        head_body += h * b; head_food += h * f;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum_food += __shfl_xor_sync(0xffffffffu, sum_food, o); sum_head += __shfl_xor_sync(0xffffffffu, sum_head, o);
        sum_body += __shfl_xor_sync(0xffffffffu, sum_body, o); head_body += __shfl_xor_sync(0xffffffffu, head_body, o);
        head_food += __shfl_xor_sync(0xffffffffu, head_food, o);
        max_body = fmaxf(max_body, __shfl_xor_sync(0xffffffffu, max_body, o));
    }
    if (lane == 0) {
        int bits = 0;                                                // (a food pixel is 0 or 1 by construction: no WURM_CHK_FOOD_VALUE)
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
struct CompactLaunch {
    int G, T, threads, blocks, smem;
};

static int plan_compact(const WurmSingleCfg* cfg, uint16_t* cells, int16_t* aux, bool step, CompactParams* cp, CompactLaunch* L) {
    if (!cfg) return fail(WURM_E_INVALID, "cfg is NULL");
    const int N = cfg->num_envs, S = cfg->size;
    if (N <= 0) return fail(WURM_E_INVALID, "num_envs must be positive");
    if (S < 9) return fail(WURM_E_INVALID, "size must be >= 9 (reference single_snake.py:346)");
    if (S > 90) return fail(WURM_E_UNSUPPORTED, "compact state: size > 90 (body values no longer fit 14 bits)");
    if (cfg->obs_mode < WURM_OBS_NONE || cfg->obs_mode > WURM_OBS_PARTIAL) return fail(WURM_E_INVALID, "bad obs_mode");
    if (cfg->obs_mode == WURM_OBS_PARTIAL && (cfg->obs_n < 0 || cfg->obs_n > 127)) return fail(WURM_E_INVALID, "bad obs_n");
    if (!cells || !aux) return fail(WURM_E_INVALID, "NULL pointer");
    if (reinterpret_cast<uintptr_t>(cells) & 15u) return fail(WURM_E_INVALID, "cells must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(aux) & 7u) return fail(WURM_E_INVALID, "aux must be 8-byte aligned");
    SingleParams& p = cp->sp;
    const int C = S * S, W = 2 * cfg->obs_n + 1;
    cp->cells = cells; cp->aux = aux; cp->Cp = (C + 7) & ~7; cp->scratch_floats = (3 * C + 3) & ~3;
    p.N = N; p.S = S; p.C = C;
    p.obs_mode = cfg->obs_mode; p.obs_n = cfg->obs_n; p.W = W;
    p.magic_S = (uint32_t)((0x100000000ull + (uint64_t)S - 1) / (uint64_t)S);
    p.magic_W = (uint32_t)((0x100000000ull + (uint64_t)W - 1) / (uint64_t)W);
    // lanes per env: ~64 cells each (measured on B200 at size 9, profiles/r02_sweep_single_compact.txt: 2 lanes 0.137 ms,
    // 4 lanes 0.172, 8 lanes 0.293 -- the serial parts of a step idle the other lanes of the group)
    int G = 1;
    while (G < 32 && G * 64 < C) G <<= 1;
    if (const char* v = getenv("WURM_COMPACT_G")) {
        const int g = atoi(v);
        if (g >= 1 && g <= 32 && (g & (g - 1)) == 0) G = g;
    }
    int threads = (cfg->obs_mode == WURM_OBS_PARTIAL && C <= 1024) ? 32 : 64;    // one-warp CTAs for small staged tiles (same sweep)
    if (const char* v = getenv("WURM_COMPACT_THREADS")) threads = atoi(v);
    if (threads < 32) threads = 32;
    if (threads > 256) threads = 256;
    threads &= ~31;
    const int row_bytes = cp->Cp * 2, stage_env = cfg->obs_mode == WURM_OBS_PARTIAL ? 3 * W * W * 4 : 0;
    const int scratch = step ? cp->scratch_floats * 4 : 0;
    int T = threads / G;
    while (T > 32 / G && T * (row_bytes + stage_env) + (T * G / 32) * scratch > 96 * 1024) T -= 32 / G;
    threads = T * G;
    p.T = T;
    p.tile_bytes_padded = (T * row_bytes + 15) & ~15;
    p.stage_bytes = (T * stage_env + 15) & ~15;
    const int smem = p.tile_bytes_padded + p.stage_bytes + (threads / 32) * scratch + 8 + 16 + 16;
    if (smem > 227 * 1024) return fail(WURM_E_UNSUPPORTED, "tile does not fit shared memory");
    L->G = G; L->T = T; L->threads = threads; L->blocks = (N + T - 1) / T; L->smem = smem;
    return WURM_OK;
}

template <int G, bool STEP>
static int launch_compact(const CompactParams& cp, const CompactLaunch& L, cudaStream_t stream) {
    auto kern = single_compact_kernel<G, STEP>;
    static SmemOptIn opt_in;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), &opt_in, L.smem, true, "cudaFuncSetAttribute(single_compact_kernel)")) return rc;
    kern<<<L.blocks, L.threads, L.smem, stream>>>(cp);
    return check_launch("single_compact_kernel");
}

template <bool STEP>
static int dispatch_compact(const CompactParams& cp, const CompactLaunch& L, cudaStream_t stream) {
    switch (L.G) {
        case 1: return launch_compact<1, STEP>(cp, L, stream);
        case 2: return launch_compact<2, STEP>(cp, L, stream);
        case 4: return launch_compact<4, STEP>(cp, L, stream);
        case 8: return launch_compact<8, STEP>(cp, L, stream);
        case 16: return launch_compact<16, STEP>(cp, L, stream);
        default: return launch_compact<32, STEP>(cp, L, stream);
    }
}

}  // namespace wurm

using namespace wurm;

extern "C" int wurm_single_compact_step(const WurmSingleCfg* cfg, uint16_t* cells, int16_t* aux, void* actions, int action_bytes,
                                        const int32_t* food_cell_replay, int auto_reset, const int32_t* spawn_replay, uint64_t seed,
                                        uint64_t step, const uint64_t* step_dev, float* obs, float* reward, uint8_t* done,
                                        uint8_t* self_col, uint8_t* edge_col, int32_t* status, int64_t* stats, uint8_t* packed,
                                        void* stream) {
    CompactParams cp = {};
    CompactLaunch L;
    if (int rc = plan_compact(cfg, cells, aux, true, &cp, &L)) return rc;
    if (!actions || !reward || !done || !self_col || !edge_col || !status) return fail(WURM_E_INVALID, "NULL pointer");
    if (!valid_action_bytes(action_bytes)) return fail(WURM_E_INVALID, "action_bytes must be 1, 2, 4 or 8");
    if (cfg->obs_mode != WURM_OBS_NONE && !obs) return fail(WURM_E_INVALID, "obs is NULL");
    SingleParams& p = cp.sp;
    p.actions = actions; p.action_bytes = action_bytes; p.food_replay = food_cell_replay;
    p.auto_reset = auto_reset; p.spawn = spawn_replay; p.packed = packed;
    p.seed = seed; p.step = step; p.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
    p.obs = obs; p.reward = reward; p.done = done; p.self_col = self_col; p.edge_col = edge_col;
    p.status = status; p.stats = reinterpret_cast<unsigned long long*>(stats);
    return dispatch_compact<true>(cp, L, (cudaStream_t)stream);
}

extern "C" int wurm_single_compact_observe(const WurmSingleCfg* cfg, const uint16_t* cells, const int16_t* aux, float* obs,
                                           int32_t* status, void* stream) {
    CompactParams cp = {};
    CompactLaunch L;
    if (int rc = plan_compact(cfg, const_cast<uint16_t*>(cells), const_cast<int16_t*>(aux), false, &cp, &L)) return rc;
    if (!obs || !status) return fail(WURM_E_INVALID, "NULL pointer");
    if (cfg->obs_mode == WURM_OBS_NONE) return WURM_OK;
    cp.sp.obs = obs; cp.sp.status = status;
    return dispatch_compact<false>(cp, L, (cudaStream_t)stream);
}

extern "C" int wurm_single_compact_reset(const WurmSingleCfg* cfg, uint16_t* cells, int16_t* aux, const uint8_t* done_mask,
                                         const int32_t* spawn_replay, uint64_t seed, uint64_t step, const uint64_t* step_dev,
                                         void* stream) {
    CompactParams cp = {};
    CompactLaunch L;
    WurmSingleCfg c = *cfg;
    c.obs_mode = WURM_OBS_NONE;
    if (int rc = plan_compact(&c, cells, aux, false, &cp, &L)) return rc;
    if (!done_mask) return fail(WURM_E_INVALID, "NULL pointer");
    cp.sp.seed = seed; cp.sp.step = step; cp.sp.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
    const int warps_per_block = 8;
    const int blocks = (cp.sp.N + 32 * warps_per_block - 1) / (32 * warps_per_block);
    single_compact_reset_kernel<<<blocks, 32 * warps_per_block, 0, (cudaStream_t)stream>>>(cp, done_mask, spawn_replay);
    return check_launch("single_compact_reset_kernel");
}

static int single_convert(const WurmSingleCfg* cfg, float* envs, uint16_t* cells, int16_t* aux, int32_t* status, void* stream,
                          bool to_compact) {
    CompactParams cp = {};
    CompactLaunch L;
    WurmSingleCfg c = *cfg;
    c.obs_mode = WURM_OBS_NONE;
    if (int rc = plan_compact(&c, cells, aux, false, &cp, &L)) return rc;
    if (!envs || (to_compact && !status)) return fail(WURM_E_INVALID, "NULL pointer");
    cp.sp.status = status;
    const int warps = 8, blocks = (cp.sp.N + warps - 1) / warps;
    if (to_compact) single_convert_kernel<true><<<blocks, 32 * warps, 0, (cudaStream_t)stream>>>(cp, envs);
    else single_convert_kernel<false><<<blocks, 32 * warps, 0, (cudaStream_t)stream>>>(cp, envs);
    return check_launch("single_convert_kernel");
}

extern "C" int wurm_single_compact_check(const WurmSingleCfg* cfg, const uint16_t* cells, const uint8_t* skip, int32_t* report,
                                         void* stream) {
    if (!cfg || !cells || !report) return fail(WURM_E_INVALID, "NULL pointer");
    if (cfg->num_envs <= 0 || cfg->size < 1) return fail(WURM_E_INVALID, "bad num_envs / size");
    const int C = cfg->size * cfg->size, warps = 8, blocks = (cfg->num_envs + warps - 1) / warps;
    single_compact_check_kernel<<<blocks, 32 * warps, 0, (cudaStream_t)stream>>>(cells, skip, cfg->num_envs, C, (C + 7) & ~7, report);
    return check_launch("single_compact_check_kernel");
}

extern "C" int wurm_single_compact(const WurmSingleCfg* cfg, const float* envs, uint16_t* cells, int16_t* aux, int32_t* status,
                                   void* stream) {
    return single_convert(cfg, const_cast<float*>(envs), cells, aux, status, stream, true);
}

extern "C" int wurm_single_expand(const WurmSingleCfg* cfg, const uint16_t* cells, const int16_t* aux, float* envs, void* stream) {
    return single_convert(cfg, envs, const_cast<uint16_t*>(cells), const_cast<int16_t*>(aux), nullptr, stream, false);
}
