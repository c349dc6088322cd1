// Device code shared by the SingleSnake kernels (single_snake.cu: the reference's fp32 layout; single_compact.cu: the
// compact resident state): launch parameters, the per-env step on an fp32 (3,S,S) env held in shared or global memory,
// food placement, env creation and the partial-observation renderer.
#pragma once
#include <math.h>

#include "../../include/wurm_b200.h"
#include "common.cuh"

namespace wurm {

struct SingleParams {
    float* envs;
    void* actions;
    const int32_t* food_replay;
    float* obs;
    float* reward;
    uint8_t* done;
    uint8_t* self_col;
    uint8_t* edge_col;
    uint8_t* packed;             // nullable: one result byte per env (pack_result)
    int32_t* status;
    unsigned long long* stats;   // nullable: WURM_STATS_SLOTS x WURM_STATS_FIELDS counters
    uint64_t seed, step;
    const unsigned long long* step_dev;   // nullable: added to `step` on the device (CUDA-graph replays)
    int auto_reset;       // fused step+reset: envs that end this step are re-created by the same launch
    const int32_t* spawn; // (N,4) replayed (y, x, dir, food_cell) of the fused / stand-alone reset, or NULL
    short* hints;         // (N,4) nullable: (head cell, snake size, food cell, -) left by the previous call -- hints only, always verified
    int N, S, C;          // envs, grid side, cells per channel
    int T;                // envs per tile (= per CTA)
    int action_bytes;     // 2 / 4 / 8
    int obs_mode, obs_n, W;
    uint32_t magic_S;     // ceil(2^32 / S): q / S == __umulhi(q, magic_S) for q < 2^16
    uint32_t magic_W;     // same for the partial-observation window width
    int tile_bytes_padded;
    int stage_bytes;      // shared staging area of the tile's partial observations (0 = none)
    int bulk_ok;          // base pointers 16-byte aligned and full-tile byte count a multiple of 16
};

__device__ __forceinline__ uint64_t call_counter(const SingleParams& p) { return p.step + (p.step_dev ? *p.step_dev : 0ull); }

__device__ __forceinline__ int div_S(int q, uint32_t magic) { return (int)__umulhi((uint32_t)q, magic); }

// int16 colour (single_snake.py:99-123) of one cell, already divided by 255.0f with IEEE rounding
// (the only values are 0, 127/255 and 1).
__device__ __forceinline__ float rgb_channel(float food, float head, float body, bool border, int c) {
    constexpr float kHalf = 127.0f / 255.0f;
    float v = 1.0f;
    if (body > kEps) v = (c == 1) ? kHalf : 0.0f;
    if (head > kEps) v = (c == 1) ? 1.0f : 0.0f;
    if (food > kEps) v = (c == 0) ? 1.0f : 0.0f;
    return border ? 0.0f : v;
}

// wurm/utils.py:36-65 evaluated literally (zero-padded cross-correlation with the four filters);
// only taken for states whose two largest body values are not a unique (size, size-1) pair.
template <int G>
__device__ __noinline__ int orientation_general(const float* body, int S, int C, uint32_t magic, float size, int l,
                                                unsigned gm) {
    const float shift = size - 2.0f;
    auto neck = [&](float v) {
        float n = v - shift;
        n = n > 0.0f ? n : 0.0f;
        if (n > 0.0f) n -= 1.5f;
        return n * 2.0f;
    };
    float best = 0.0f;
    int best_k = 0;
    for (int k = 0; k < 4; ++k) {
        float mk = -INFINITY;
        for (int q = l; q < C; q += G) {
            const int y = div_S(q, magic), x = q - y * S;
            const int yy = y + off_y(k), xx = x + off_x(k);
            const float nb = (yy >= 0 && yy < S && xx >= 0 && xx < S) ? neck(body[yy * S + xx]) : 0.0f;
            mk = fmaxf(mk, nb - neck(body[q]));
        }
        mk = group_max<G>(mk, gm);
        if (k == 0 || mk > best) { best = mk; best_k = k; }
    }
    return best_k;
}

// Uniform choice among the free interior cells of one env (single_snake.py:306-320), raster order,
// ranked with ballots over the group's lanes.  Returns -1 if there is no free cell.
template <int G>
__device__ __noinline__ int pick_free_cell(const float* env, int S, int C, uint32_t magic, uint32_t rnd, int l,
                                           unsigned gm) {
    const unsigned shift = (G == 32) ? 0u : ((threadIdx.x & 31u) & ~(unsigned)(G - 1));
    const unsigned low = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    auto free_bits = [&](int base) {
        const int q = base + l;
        bool f = false;
        if (q < C) {
            const int y = div_S(q, magic), x = q - y * S;
            f = y >= 1 && y <= S - 2 && x >= 1 && x <= S - 2 && (env[q] + env[C + q] + env[2 * C + q] < kEps);
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

// single_snake.py:344-387 _create_envs for env e: seed cell (y, x), direction d and food cell of the new
// env, replayed from `spawn` or drawn from Philox with call counter `ctr`.
__device__ __forceinline__ void new_env_layout(const SingleParams& p, const int32_t* spawn, uint64_t ctr, int e, int& tail,
                                               int& mid, int& hd, int& cell) {
    const int S = p.S;
    int y, x, d;
    if (spawn) {
        y = spawn[4 * (size_t)e]; x = spawn[4 * (size_t)e + 1]; d = spawn[4 * (size_t)e + 2]; cell = spawn[4 * (size_t)e + 3];
    } else {
        const uint4 r = draw(p.seed, ctr, (uint32_t)e, kStreamSingleReset);
        y = 4 + (int)bounded(r.x, (uint32_t)(S - 8));                // :358 randint(4, S-4)
        x = 4 + (int)bounded(r.y, (uint32_t)(S - 8));                // :359
        d = (int)(r.z >> 30);                                        // :366 randint(4)
        // :384 one free interior cell: the r-th interior cell in raster order, skipping the three snake
        // cells (all interior because 4 <= y,x < S-4)
        const int I = S - 2;
        int s0 = (y - off_y(d) - 1) * I + (x - off_x(d) - 1), s1 = (y - 1) * I + (x - 1),
            s2 = (y + off_y(d) - 1) * I + (x + off_x(d) - 1);
        if (s0 > s2) { const int tmp = s0; s0 = s2; s2 = tmp; }      // s1 is always the middle one
        int rr = (int)bounded(r.w, (uint32_t)(I * I - 3));
        if (rr >= s0) ++rr;
        if (rr >= s1) ++rr;
        if (rr >= s2) ++rr;
        const int fy = rr / I;
        cell = (fy + 1) * S + (rr - fy * I + 1);
    }
    tail = (y - off_y(d)) * S + (x - off_x(d)); mid = y * S + x; hd = (y + off_y(d)) * S + (x + off_x(d));
}

// value of element i of the (3,S,S) state of a freshly created env (:372-385 LENGTH_3_SNAKES stamp, head, food)
__device__ __forceinline__ float new_env_value(int i, int C, int tail, int mid, int hd, int cell) {
    float v = 0.0f;
    if (i == cell) v = 1.0f;
    if (i == C + hd) v = 1.0f;
    if (i == 2 * C + tail) v = 1.0f;
    if (i == 2 * C + mid) v = 2.0f;
    if (i == 2 * C + hd) v = 3.0f;
    return v;
}

// single_snake.py:197-300 for one environment held in shared memory, executed by a group of G lanes.
// WT (write-through): mirror every update of `env` into the fp32 state in HBM (p.envs); the compact-state kernels run
// the same code on a scratch expansion of their records and pass false.
template <int G, bool WT = true>
__device__ __forceinline__ int step_env(const SingleParams& p, float* env, int e, int l, int* cnt_s, long long a_in,
                                        int hint_head, int hint_sz, bool& ended) {
    const unsigned gm = group_mask<G>();
    const int S = p.S, C = p.C;
    float* food = env;
    float* head = env + C;
    float* body = env + 2 * C;
    // A step changes O(snake length) cells of the env.  Only those are written back to HBM, cell by
    // cell, at the point where the shared copy is updated (the sectors were loaded through L2 by this
    // CTA's bulk copy microseconds ago, so the partial writes merge there); the tile is never stored
    // wholesale.
    float* gfood = p.envs + (size_t)e * 3 * C;
    float* ghead = gfood + C;
    float* gbody = gfood + 2 * C;

    // Snake size (:210), head cell, and the (size, size-1) cell counts the orientation rule needs.
    // The previous call left (head cell, size) HINTS per env.  They are never trusted: the head hint is
    // used only if that cell really holds a head, the size hint only if the scan finds the same maximum --
    // then the counts taken against the hinted size in the SAME pass are the true ones and the second pass
    // over the body, as well as the scan of the head channel, are skipped.  Stale hints (caller edited
    // `envs`, first step, hints disabled) fall back to the full scans.
    int hp = -1, hc = 0;
    float hint_size = -1.0f;
    if (p.hints) {
        hint_size = (float)hint_sz;
        if (hint_head >= 0 && hint_head < C && head[hint_head] != 0.0f) { hp = hint_head; hc = 1; }
    }
    const bool scan_head = hc == 0;
    float m = -INFINITY;
    int c1 = 0, c2 = 0, p1 = -1, p2 = -1;
    const float hint_sm1 = hint_size - 1.0f;
    if (scan_head) {
#pragma unroll 4
        for (int q = l; q < C; q += G) {
            const float v = body[q];
            m = fmaxf(m, v);
            if (v == hint_size) { ++c1; p1 = q; }
            if (v == hint_sm1) { ++c2; p2 = q; }
            if (head[q] != 0.0f) { hp = q; ++hc; }
        }
        hp = group_max<G>(hp, gm);
        hc = group_sum<G>(hc, gm);
    } else {
#pragma unroll 4
        for (int q = l; q < C; q += G) {
            const float v = body[q];
            m = fmaxf(m, v);
            if (v == hint_size) { ++c1; p1 = q; }
            if (v == hint_sm1) { ++c2; p2 = q; }
        }
    }
    const float size = group_max<G>(m, gm);

    // orientation (:212).  Canonical case: exactly one cell == size (head) and one == size-1
    // (neck): the response of filter k peaks at 2 iff head = neck + OFF[k]; otherwise all four
    // filters tie at 1 and argmax returns 0.
    if (size != hint_size) {                                         // stale size hint: count against the true size
        c1 = c2 = 0; p1 = p2 = -1;
        const float sm1 = size - 1.0f;
#pragma unroll 4
        for (int q = l; q < C; q += G) {
            const float v = body[q];
            if (v == size) { ++c1; p1 = q; }
            if (v == sm1) { ++c2; p2 = q; }
        }
    }
    c1 = group_sum<G>(c1, gm);
    c2 = group_sum<G>(c2, gm);
    p1 = group_max<G>(p1, gm);
    p2 = group_max<G>(p2, gm);
    int k = 0;
    if (c1 == 1 && c2 == 1) {
        const int d = p1 - p2;
        const int x2 = p2 - div_S(p2, p.magic_S) * S;
        if (d == -S) k = 0;
        else if (d == 1 && x2 != S - 1) k = 1;
        else if (d == S) k = 2;
        else if (d == -1 && x2 != 0) k = 3;
    } else {
        k = orientation_general<G>(body, S, C, p.magic_S, size, l, gm);
    }

    // action sanitisation, written back into the caller's tensor (:221-222)
    const long long a = (a_in + ((long long)k == a_in ? 2 : 0)) % 4;
    if (l == 0 && a != a_in) store_action(p.actions, p.action_bytes, (size_t)e, a);

    // head move (:225-233): the conv2d with filter a translates the head channel by -OFF[a];
    // a head that leaves the grid vanishes.
    int np = -1, ny = -1, nx = -1;
    if (hp >= 0) {
        const int hy = div_S(hp, p.magic_S), hx = hp - hy * S;
        ny = hy - (a >= 0 ? off_y((int)a) : 0);
        nx = hx - (a >= 0 ? off_x((int)a) : 0);
        if (ny >= 0 && ny < S && nx >= 0 && nx < S) np = ny * S + nx;
    }

    const float ov = (np >= 0) ? food[np] : 0.0f;                   // :242 head-food overlap
    if (ov == 0.0f) {                                                // :246-249 decay unless it ate
#pragma unroll 4
        for (int q = l; q < C; q += G) {
            const float v = body[q], nv = fmaxf(v - 1.0f, 0.0f);
            if (nv != v) { body[q] = nv; if (WT) gbody[q] = nv; }
        }
    }
    __syncwarp(gm);
    const bool sc = (np >= 0) && (body[np] > kEps);                  // :252 self collision
    const bool interior = (np >= 0) && ny >= 1 && ny <= S - 2 && nx >= 1 && nx <= S - 2;
    __syncwarp(gm);
    if (l == 0 && hp >= 0) {
        head[hp] = 0.0f; if (WT) ghead[hp] = 0.0f;
        if (np >= 0) {
            head[np] = 1.0f; if (WT) ghead[np] = 1.0f;
            const float grown = body[np] + (size + ov);              // :258-262 growth
            body[np] = grown; if (WT) gbody[np] = grown;
            if (ov != 0.0f) { const float left = food[np] + ov * -1.0f; food[np] = left; if (WT) gfood[np] = left; }   // :270-272
        }
    }
    __syncwarp(gm);
    if (ov != 0.0f) {                                                // :277-282 respawn
        int cell;
        if (p.food_replay) cell = p.food_replay[e];
        else {
            // uniform over the free interior cells by rejection (every lane of the group draws the
            // same candidates); after kRejectionTries misses the free cells are ranked explicitly
            const int I = S - 2;
            cell = -1;
            for (uint32_t t = 0; t < kRejectionTries && cell < 0; ++t) {
                const int cand = (int)bounded(draw_i(p.seed, call_counter(p), (uint32_t)e, kStreamSingleStepFood, t), (uint32_t)(I * I));
                const int cy = cand / I, q = (1 + cy) * S + 1 + (cand - cy * I);
                if (env[q] + env[C + q] + env[2 * C + q] < kEps) cell = q;
            }
            if (cell < 0)
                cell = pick_free_cell<G>(env, S, C, p.magic_S,
                                         draw_i(p.seed, call_counter(p), (uint32_t)e, kStreamSingleStepFood, kRejectionTries), l, gm);
        }
        if (l == 0 && cell >= 0) { const float f = food[cell] + 1.0f; food[cell] = f; if (WT) gfood[cell] = f; }
        if (l == 0 && p.hints) p.hints[4 * (size_t)e + 2] = (short)cell;
    }
    if (l == 0) {
        p.reward[e] = 0.0f - ov * -1.0f;                             // :271
        p.self_col[e] = sc;
        p.edge_col[e] = !interior;                                   // :290-293 no head in the interior
        p.done[e] = sc || !interior;
        if (p.packed) p.packed[e] = pack_result(sc || !interior, sc, !interior, ov);
        if (hc > 1) atomicOr(p.status, WURM_ST_MULTI_HEAD);
        if (p.hints) {                                               // for the next call: head cell and size now
            p.hints[4 * (size_t)e] = (short)np;
            p.hints[4 * (size_t)e + 1] = (short)((np >= 0) ? (int)(size + ov) : -1);
        }
        if (p.stats) {                                               // episode statistics, per-CTA partials
            if (sc || !interior) atomicAdd(cnt_s + 0, 1);
            if (ov != 0.0f) atomicAdd(cnt_s + 1, (int)ov);
            if (sc) atomicAdd(cnt_s + 2, 1);
            if (!interior) atomicAdd(cnt_s + 3, 1);
        }
    }
    __syncwarp(gm);                                                  // lane 0's cell updates -> the group's render
    ended = sc || !interior;
    return np;
}

template <int G>
__device__ __forceinline__ int find_head(const SingleParams& p, const float* env, int l) {
    const unsigned gm = group_mask<G>();
    int hp = -1, hc = 0;
    for (int q = l; q < p.C; q += G)
        if (env[p.C + q] != 0.0f) { hp = q; ++hc; }
    hp = group_max<G>(hp, gm);
    hc = group_sum<G>(hc, gm);
    if (l == 0 && hc > 1) atomicOr(p.status, WURM_ST_MULTI_HEAD);
    return (hc == 1) ? hp : -1;
}

// single_snake.py:166-193 partial_n: the (3,W,W) crop of rgb/255 around the head, zero outside the grid,
// rendered by the env's own lane group into the tile's shared staging area (one window cell per lane
// and iteration, three channel rows), from where the whole tile's observations leave in one bulk store.
template <int G>
__device__ __forceinline__ void render_partial(const SingleParams& p, const float* env, int hp, float* out, int l) {
    const int S = p.S, C = p.C, W = p.W, WW = W * W, n = p.obs_n;
    if (hp < 0) {                                                    // the reference raises here (:191)
        for (int r = l; r < 3 * WW; r += G) out[r] = 0.0f;
        if (l == 0) atomicOr(p.status, WURM_ST_NO_HEAD_PARTIAL);
        return;
    }
    constexpr float kHalf = 127.0f / 255.0f;
    const int hy = div_S(hp, p.magic_S), hx = hp - hy * S;
    for (int ij = l; ij < WW; ij += G) {
        const int i = (int)__umulhi((uint32_t)ij, p.magic_W), j = ij - i * W;
        const int y = hy - n + i, x = hx - n + j;
        float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
        if ((unsigned)(y - 1) < (unsigned)(S - 2) && (unsigned)(x - 1) < (unsigned)(S - 2)) {   // inside the grid, not on its border
            const int q = y * S + x;
            v0 = v1 = v2 = 1.0f;                                     // empty cell: white
            if (env[2 * C + q] > kEps) { v0 = 0.0f; v1 = kHalf; v2 = 0.0f; }
            if (env[C + q] > kEps) { v0 = 0.0f; v1 = 1.0f; v2 = 0.0f; }
            if (env[q] > kEps) { v0 = 1.0f; v1 = 0.0f; v2 = 0.0f; }
        }
        out[ij] = v0; out[WW + ij] = v1; out[2 * WW + ij] = v2;
    }
}


}  // namespace wurm
