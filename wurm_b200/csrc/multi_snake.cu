// MultiSnake hot path for B200 (sm_100a): step / reset / observe as hand-written CUDA.
//
// Replaces wurm/envs/multi_snake.py:163-1019 of the reference (a ~330-op ATen tensor program per
// step with conv2d head moves, an (A,K,S,S) repeat_interleave for collisions, masked_select crops
// and >=6 host syncs) with ONE launch per call.
//
// Design:
//   * one CTA per environment.  The env's fp32 state -- (1+2K) grids of S*S floats, 22.5 KB at
//     K=4,S=25 but 540 KB at K=16,S=64, more than an SM's shared memory -- is STREAMED once from
//     HBM with coalesced 128-bit loads and folded on the fly into a compact shared-memory form:
//     one 32-bit record per cell (owner snake, body value), one byte per cell for food, one head
//     index per snake.  The state is almost entirely zeros, so the fold is a handful of shared
//     atomics per env.
//   * the whole step (boost phase, regular phase, food on death, boost cost, food respawn) runs on
//     that compact form: the per-snake logic on one lane per snake of warp 0 (collisions are
//     look-ups at the new head cell, head-to-head clashes a K-wide compare), the per-cell work
//     (decay, food from dead bodies, deletion) as CTA-wide passes over the records;
//   * a step changes O(snake length) cells: the records carry "modified since the load" bits and only
//     those cells are stored back into the reference's fp32 tensors (sparse write-back; a state the
//     compact form cannot carry exactly is expanded densely instead), and the observations are rendered
//     from the compact form straight into the policy's per-agent input buffers;
//   * the same records may also live in HBM between calls (WurmMultiState.cells, 4 bytes per cell): as the state itself
//     (COMPACT: the fp32 tensors are NULL) or BESIDE the tensors as their shadow (COMPACT + SHADOW, the default of the
//     Python class: the tensors stay the state, current after every call).  A step then starts with ONE bulk copy (TMA,
//     cp.async.bulk + mbarrier) of the env's records into the shared record array instead of streaming ~99 % zeros; the
//     shadow variant checks every record it loaded against the tensors (an env they contradict is re-loaded from them) and
//     writes its changes to both forms;
//   * the reset (stand-alone kernel, or fused into the step's launch) decides on the same records; the K sequential snake
//     placements of a re-created env run on one warp over maintained candidate counts (warp_pick);
//   * no tensor cores: nothing here is a dense contraction.
//
// Supported states: the reference's own invariant (MultiSnake.check_consistency, :733-769) --
// bodies of different snakes never share a cell when a step starts, at most one head per snake,
// dead snakes are all-zero.  A violating input raises WURM_ST_OVERLAP / WURM_ST_MULTI_HEAD.
// The reference runs its boost phase when ANY agent of the batch boosts (:503); here each env
// runs it when one of ITS agents boosts, which is identical on supported states (for an env
// without boosting agents the phase changes nothing) -- except in replay mode, where the flag
// recorded from the reference is followed literally.
#include <math.h>
#include <stdlib.h>

#include "../../include/wurm_b200.h"
#include "common.cuh"
#include "host_util.h"

namespace wurm {

constexpr int kMaxK = WURM_MULTI_MAX_SNAKES;

struct MultiParams {
    // state (reference layouts)
    float* foods;
    float* heads;
    float* bodies;
    uint8_t* dones;
    long long* orientations;
    uint8_t* boost_this_step;
    short* colours;
    short* head_hints;    // (E*K) nullable: head cell left by the previous call (-1 dead, -2 unknown), verified before use
    uint32_t* cells;      // compact resident state (E, Cp) or NULL: see WurmMultiState.cells; head_hints is then authoritative
    int Cp;               // row pitch of `cells`: C rounded up to a multiple of 4 (rows stay 16-byte aligned)
    int cells_valid;      // the records describe the current state (load from them); 0: load the dense tensors, emit the records
    // step inputs
    const void* actions[kMaxK];
    int action_bytes;
    int replay, boost_phase_ran;
    const float* u_boost;
    const float* u_cost;
    const float* u_reg;
    const int* food_cell;
    const float* u_rate;
    uint64_t seed, step;
    const unsigned long long* step_dev;   // nullable: added to `step` on the device (CUDA-graph replays)
    // rules
    int E, K, S, C;
    int boost, food_on_death, food_mode, respawn_any, colour_random;
    float death_thr, boost_cost_prob, food_rate, reward_on_death;
    // step outputs
    float* rewards;
    uint8_t* snake_col;
    uint8_t* edge_col;
    float* food_cons;
    float* sizes;
    uint8_t* dones_out;
    uint8_t* boost_out;
    uint8_t* all_done;
    float* obs;
    short* img;
    int obs_mode, obs_n, W;
    // reset inputs
    const uint8_t* env_done;
    const int* create;
    const int* respawn;
    const short* colours_replay;
    int* status;
    unsigned long long* stats;
    uint32_t magic_S, magic_C, magic_W;
    int auto_reset;       // fused step+reset: the env's reset runs in the step's launch, with call counter + 1
};

__device__ __forceinline__ uint64_t call_counter(const MultiParams& p) { return p.step + (p.step_dev ? *p.step_dev : 0ull); }

__device__ __forceinline__ int fdiv(int q, uint32_t magic) { return (int)__umulhi((uint32_t)q, magic); }
// i / C for i < K*C (up to 32 * 181^2): with magic = ceil(2^32 / C) the multiply-shift estimate is exact only while
// i * (C * magic - 2^32) < 2^32, which large K and S exceed for the last cells of a snake's grid; the estimate is
// never more than one too high, so one compare corrects it.
__device__ __forceinline__ int fdiv_C(int i, int C, uint32_t magic_C) {
    const int q = (int)__umulhi((uint32_t)i, magic_C);
    return q * C > i ? q - 1 : q;
}

// One 32-bit record per cell holds everything the step needs to know about it:
//   bits  0-15  body value            bits 16-21  owner snake + 1      (together: the "live" body, 0 = none)
//   bits 22-27  owner + 1 when loaded bit 28      body modified since the load
//   bit  29     food                  bit 30      food when loaded
//   bit  31     the cell is on the env's live list
// so that the write-back can tell exactly which cells of which tensors changed.
// The LIVE LIST holds every cell whose record ever became non-zero (a body or food at load time, a new head cell,
// spawned food): ~1 % of the grid.  The per-cell passes of a step (decay, deaths, boost cost, write-back) walk the
// list -- a couple of iterations with every lane busy -- instead of all S*S records.
constexpr uint32_t kLive = 0x003FFFFFu, kDirty = 1u << 28, kFood = 1u << 29, kFood0 = 1u << 30, kListed = 1u << 31;
__device__ __forceinline__ uint32_t make_rec(int owner, int value) { return ((uint32_t)(owner + 1) << 16) | (uint32_t)value; }
__device__ __forceinline__ int rec_owner(uint32_t r) { return (int)((r >> 16) & 63u) - 1; }
__device__ __forceinline__ int rec_owner0(uint32_t r) { return (int)((r >> 22) & 63u) - 1; }
__device__ __forceinline__ int rec_value(uint32_t r) { return (int)(r & 0xffffu); }
__device__ __forceinline__ bool rec_body(uint32_t r) { return (r & kLive) != 0u; }

__device__ __forceinline__ float4 ld_stream(const float4* ptr) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr));
    return v;
}

// Calls f(i, v) for every non-zero base[i], i < n.  128-bit streaming loads, U in flight per thread (4 for the small
// CTAs, whose registers decide how many envs are resident; 8 for the 256-thread CTAs of large grids: C5 1.65 -> 1.58 ms);
// loads past the end are clamped to the last vector instead of predicated (their hits are dropped by the index test).
template <int U = 4, typename F>
__device__ __forceinline__ void scan_nonzero(const float* base, int n, F&& f) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    int lead = (4 - (int)((reinterpret_cast<uintptr_t>(base) >> 2) & 3)) & 3;
    if (lead > n) lead = n;
    const int nvec = (n - lead) >> 2;
    const int tail0 = lead + 4 * nvec;
    if (tid < lead + (n - tail0)) {                                  // the (at most 3 + 3) unaligned edge elements
        const int i = tid < lead ? tid : tail0 + (tid - lead);
        const float v = base[i];
        if (v != 0.0f) f(i, v);
    }
    const float4* vb = reinterpret_cast<const float4*>(base + lead);
    for (int j0 = tid; j0 < nvec; j0 += U * nthr) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld_stream(vb + min(j0 + u * nthr, nvec - 1));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (v[u].x != 0.0f || v[u].y != 0.0f || v[u].z != 0.0f || v[u].w != 0.0f) {
                const int j = j0 + u * nthr, i = lead + 4 * j;
                if (j < nvec) {
                    if (v[u].x != 0.0f) f(i, v[u].x);
                    if (v[u].y != 0.0f) f(i + 1, v[u].y);
                    if (v[u].z != 0.0f) f(i + 2, v[u].z);
                    if (v[u].w != 0.0f) f(i + 3, v[u].w);
                }
            }
        }
    }
}

// base[i] = gen(i) for i < n, 128-bit stores where the address allows.
template <typename G>
__device__ __forceinline__ void store_floats(float* base, int n, G&& gen) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    int lead = (4 - (int)((reinterpret_cast<uintptr_t>(base) >> 2) & 3)) & 3;
    if (lead > n) lead = n;
    if (tid < lead) base[tid] = gen(tid);
    const int nvec = (n - lead) >> 2;
    float4* vb = reinterpret_cast<float4*>(base + lead);
    for (int j = tid; j < nvec; j += nthr) {
        const int i = lead + 4 * j;
        vb[j] = make_float4(gen(i), gen(i + 1), gen(i + 2), gen(i + 3));
    }
    const int tail0 = lead + 4 * nvec;
    if (tid < n - tail0) base[tail0 + tid] = gen(tail0 + tid);
}

// (float)v / 255.0f, correctly rounded, for integer v in [-1024, 70000] (verified exhaustively on the
// host and by the parity tests): quotient estimate + one fused residual correction instead of the
// ~12-instruction IEEE division sequence.
__device__ __forceinline__ float div255(int v) {
    constexpr float kRcp = 1.0f / 255.0f;
    const float a = (float)v, q = a * kRcp;
    return fmaf(fmaf(-q, 255.0f, a), kRcp, q);
}

struct MultiSmem {
    uint32_t* cell;   // C records
    int* hp;          // head cell per snake, -1 none
    int* size;        // max body value per snake
    int* hcnt;        // head cells seen per snake
    int* done;
    int* decay;
    int* cost;
    int* boost;
    int* sum;         // sum of body values per snake (invariant check)
    int* misc;        // [0] food cells, [1] live-list length, [2] run_boost, [3] force full write-back, [6] queued hits,
                      // [8] the live list overflowed its capacity: the per-cell passes visit every cell instead
    int lcap;         // capacity of the live list
    int2* queue;      // non-zero elements found by the load scan: (tensor << 30 | index, value bits), see load_env
    int qcap;
    short* col;       // K*3
    unsigned char* reset;   // ResetScratch of the fused step+reset path (>= 768 bytes)
    unsigned short* list;   // live list (<= C entries); ALIASES the reset scratch too: it is dead once the state is written back
    float* tab;       // 2*32*3: rendered colour of snake o's body (2o) / head (2o+1) cell, see build_colour_table; ALIASES
                      // the reset scratch, which is only used after the observation is written -- shared memory per
                      // CTA decides how many envs are resident per SM, and 768 bytes more cost 9 % at K=4, S=25
};

// Shared memory per CTA decides how many small envs are resident per SM (K=4, S=25: 768 bytes more cost 9 %), so the
// per-snake arrays are sized by K (rounded up to 4), not by the 32-snake maximum.
__host__ __device__ __forceinline__ int snakes_padded(int K) { return (K + 3) & ~3; }
__host__ __device__ __forceinline__ int queue_capacity(int K) { return 12 * K + 32; }
// Bytes of the region that serves, in turn, the live list, the colour table (768) and the fused reset's scratch
// (occupancy bytes + picks).  Dense layout: the list can hold every cell.  Compact layout (multi_env_kernel<.,.,true>):
// the region is only as large as the reset scratch needs -- the list holds what fits (a few times the typical ~1 % of the
// grid) and an env with more live cells than that is walked cell by cell instead (misc[8]); no load queue either.
// Shared memory per CTA is what limits the number of resident envs, and the kernel is latency-bound.
// the fused / stand-alone reset's scratch: occupancy bytes, block_pick's per-warp counts, the pick, K seed cells and directions
__host__ __device__ __forceinline__ int reset_scratch_size(int C) { return ((C + 15) & ~15) + (16 + 4 + 32 + 32) * 4; }
// ... and, elsewhere (shared memory per CTA is what the step kernels' occupancy hangs on: the fused reset puts them into the
// record array, which is dead by then): seed-candidate counts per 16-byte vector of the occupancy bytes and per lane
__host__ __device__ __forceinline__ int reset_counts_size(int C) { return (((((C + 15) & ~15) >> 4) + 15) & ~15) + 32 * 4; }
__host__ __device__ __forceinline__ int scratch_bytes(int C, bool compact) {
    int scratch = reset_scratch_size(C);
    if (scratch < 768) scratch = 768;
    if (!compact && scratch < 2 * C) scratch = 2 * C;
    return scratch;
}

__device__ __forceinline__ MultiSmem carve(unsigned char* smem, int C, int K, bool compact = false) {
    MultiSmem s;
    const int KP = snakes_padded(K);
    s.cell = reinterpret_cast<uint32_t*>(smem);
    s.queue = reinterpret_cast<int2*>(s.cell + ((C + 1) & ~1));        // 8-byte aligned
    s.qcap = compact ? 0 : queue_capacity(K);
    s.lcap = compact ? min(C, scratch_bytes(C, true) / 2) : C;
    s.hp = reinterpret_cast<int*>(s.queue + s.qcap);
    s.size = s.hp + KP; s.hcnt = s.size + KP; s.done = s.hcnt + KP; s.decay = s.done + KP; s.cost = s.decay + KP;
    s.boost = s.cost + KP; s.sum = s.boost + KP; s.misc = s.sum + KP;
    s.col = reinterpret_cast<short*>(s.misc + 12);
    s.reset = reinterpret_cast<unsigned char*>(s.col + 3 * KP);
    s.tab = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s.reset) + 15) & ~(uintptr_t)15);
    s.list = reinterpret_cast<unsigned short*>(s.tab);
    return s;
}

static size_t multi_smem_bytes(int C, int K, bool compact = false) {
    // records + load queue + per-snake arrays + misc + colours + the shared scratch region
    const size_t KP = (size_t)snakes_padded(K);
    return (size_t)((C + 1) & ~1) * 4 + (size_t)(compact ? 0 : queue_capacity(K)) * 8 + 8 * KP * 4 + 12 * 4 + 3 * KP * 2 +
           (size_t)scratch_bytes(C, compact) + 32;
}

// Food transitions keep the env's food-cell count (misc[0]) current, so _add_food needs no counting pass.
__device__ __forceinline__ void list_push(const MultiSmem& s, int q) {
    const int n = atomicAdd(&s.misc[1], 1);
    if (n < s.lcap) s.list[n] = (unsigned short)q;
    else s.misc[8] = 1;
}
// The per-cell passes walk the live list, or every cell of the grid once the list has overflowed.  The choice is taken
// ONCE per pass (a push that overflows while a pass is running must not change what its loop indices mean); cells that
// become live during a pass are for the next one, as with the list.
struct LiveWalk {
    int cnt;
    bool all;
};
__device__ __forceinline__ LiveWalk live_walk(const MultiSmem& s, int C) {
    LiveWalk w;
    w.all = s.misc[8] != 0;
    w.cnt = w.all ? C : min(s.misc[1], s.lcap);
    return w;
}
__device__ __forceinline__ int live_at(const MultiSmem& s, const LiveWalk& w, int n) { return w.all ? n : (int)s.list[n]; }
__device__ __forceinline__ void set_food(const MultiSmem& s, int q) {
    const uint32_t old = atomicOr(&s.cell[q], kFood | kListed);
    if (!(old & kFood)) atomicAdd(&s.misc[0], 1);
    if (!(old & kListed)) list_push(s, q);
}
__device__ __forceinline__ void clear_food(const MultiSmem& s, int q) {
    if (atomicAnd(&s.cell[q], ~kFood) & kFood) atomicSub(&s.misc[0], 1);
}

// Folds one non-zero element of env e's tensors into the compact form (what: 0 food, 1 head, 2 body).  A value the
// compact form cannot carry exactly (food / head != 1, non-integral body) or two bodies on one cell set misc[3]: the
// env is then written back in full (which normalises it) instead of cell by cell.
template <bool CHECK>
__device__ __forceinline__ void fold_nonzero(const MultiSmem& s, int C, uint32_t magic_C, int32_t* status, int what, int i, float v) {
    bool odd;
    if (what == 0) {
        if (!(atomicOr(&s.cell[i], kFood | kFood0 | kListed) & kListed)) list_push(s, i);
        atomicAdd(&s.misc[0], 1);
        odd = v != 1.0f;
        if (CHECK && odd) s.misc[5] = 1;                             // a food pixel that is neither 0 nor 1
    } else if (what == 1) {
        const int k = fdiv_C(i, C, magic_C);
        atomicMax(&s.hp[k], i - k * C);
        atomicAdd(&s.hcnt[k], v == 1.0f ? 1 : 2);                    // a head value other than 1 counts as "not one head"
        odd = v != 1.0f;
    } else {
        const int k = fdiv_C(i, C, magic_C), val = (int)v;
        const uint32_t owner = (uint32_t)(k + 1);
        const uint32_t old = atomicOr(&s.cell[i - k * C], (owner << 22) | (owner << 16) | ((uint32_t)val & 0xffffu) | kListed);
        odd = (float)val != v || val < 1 || val > 65535;
        if (old & kLive) { atomicOr(status, WURM_ST_OVERLAP); odd = true; }
        if (!(old & kListed)) list_push(s, i - k * C);
        atomicMax(&s.size[k], val);
        if (CHECK) atomicAdd(&s.sum[k], val);
    }
    if (odd) s.misc[3] = 1;
}

// (plain arguments: a reference to the kernel's parameter struct would force a copy of it onto the stack)
template <bool CHECK>
__device__ __noinline__ void fold_nonzero_overflow(unsigned char* smem, int C, int K, uint32_t magic_C, int32_t* status, int what,
                                                   int i, float v, bool compact_layout) {
    fold_nonzero<CHECK>(carve(smem, C, K, compact_layout), C, magic_C, status, what, i, v);
}

// Streams env e's tensors from HBM into the compact shared-memory form.  ONE copy of the scan loop runs over the
// three tensors; the ~1 % non-zero elements it meets are only QUEUED (a counter bump and one 8-byte store, whichever
// lane meets them), and folded afterwards with every lane busy -- handling each hit where it is found would run the
// whole fold with one or two lanes active per hit.  Hits beyond the queue's capacity are folded on the spot.
// `use_hints`: thread k < K brings snake k's head hint, the value of the heads tensor at that cell and the snake's done
// flag (loaded by the caller before the scans start, looked at only after the foods / bodies streams so that their
// latency hides behind them).  A live snake's hint verifies if the cell holds a head, a dead snake's if it is still
// flagged done; if every snake's does, the heads tensor -- one non-zero per snake in 44-48 % of the state's bytes --
// is not streamed at all.
template <bool CHECK = false, int U = 4>
__device__ __forceinline__ void load_env(const MultiParams& p, const MultiSmem& s, int e, bool use_hints = false, int hint_h = -1,
                                         float hint_val = 0.0f, bool hint_dead = false) {
    const int C = p.C, K = p.K;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int what = pass == 0 ? 0 : pass == 1 ? 2 : 1;            // foods, bodies, heads
        if (what == 1 && use_hints) {
            if (threadIdx.x < 32) {
                const bool mine_ok = threadIdx.x >= K || ((hint_h >= 0 && hint_h < C) ? (!hint_dead && hint_val == 1.0f)
                                                                                     : (hint_h == -1 && hint_dead));
                const bool all_ok = __all_sync(0xffffffffu, mine_ok);
                if (all_ok && threadIdx.x < K) { s.hp[threadIdx.x] = hint_h; s.hcnt[threadIdx.x] = hint_h >= 0 ? 1 : 0; }
                __syncwarp();
                if (threadIdx.x == 0) s.misc[7] = all_ok;
            }
            __syncthreads();
            if (s.misc[7]) break;
        }
        const float* base = what == 0 ? p.foods + (size_t)e * C : (what == 1 ? p.heads : p.bodies) + (size_t)e * K * C;
        scan_nonzero<U>(base, what == 0 ? C : K * C, [&](int i, float v) {
            const int n = atomicAdd(&s.misc[6], 1);
            if (n < s.qcap) s.queue[n] = make_int2((int)((unsigned)i | ((unsigned)what << 30)), __float_as_int(v));
            else fold_nonzero_overflow<CHECK>(reinterpret_cast<unsigned char*>(s.cell), C, K, p.magic_C, p.status, what, i, v, s.qcap == 0);
        });
    }
    __syncthreads();
    const int nq = min(s.misc[6], s.qcap);
    for (int n = threadIdx.x; n < nq; n += blockDim.x) {
        const int2 h = s.queue[n];
        fold_nonzero<CHECK>(s, C, p.magic_C, p.status, (int)((unsigned)h.x >> 30), h.x & 0x3fffffff, __int_as_float(h.y));
    }
}

// COMPACT RESIDENT STATE.  Between calls the env may live in HBM in the kernel's own form instead of the reference's
// fp32 tensors: one 32-bit record per cell -- bits 0-15 body value, 16-21 owner + 1, bit 29 food, everything else zero
// -- plus each snake's head cell in `head_hints` (authoritative in this mode: -1 = none).  16 KB per env at K=16, S=64
// instead of the 278 KB of fp32 zeros the dense path streams.  Loading is a copy into shared memory that queues the
// ~1 % non-zero cells on the live list as it goes; the per-call bits (owner / food when loaded, listed) are derived here.
__device__ __forceinline__ uint32_t hbm_record(uint32_t rec) { return rec & (kLive | kFood); }

// VERIFY (shadowed dense state, below): the same walk also checks the records against the reference's tensors; returns this
// thread's verdict (to be combined across the CTA by the caller).
template <bool CHECK = false, bool VERIFY = false>
// (VERIFY: the caller has fetched this thread's snake's head cell `pre_h` and the heads tensor's value there `head_val` already)
// `bar`: the records' 128-bit vectors are already on their way into s.cell by ONE bulk copy (TMA) the caller issued at the top
// of the kernel, completion on this mbarrier -- instead of every thread pulling its vectors one dependent load after the other.
__device__ __forceinline__ bool load_env_compact(const MultiParams& p, const MultiSmem& s, int e, int pre_h = -1, float head_val = 1.0f,
                                                 uint64_t* bar = nullptr) {
    const int C = p.C, K = p.K, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    const uint32_t* g = p.cells + (size_t)e * p.Cp;
    bool ok = true;
    int head_cell = -1;
    if (tid < K) {
        const int h = VERIFY ? pre_h : (int)p.head_hints[(size_t)e * K + tid];
        head_cell = (h >= 0 && h < C) ? h : -1;
        s.hp[tid] = head_cell;
        s.hcnt[tid] = head_cell >= 0 ? 1 : 0;
        if (VERIFY && h == -2) ok = false;                           // "unknown": the records of this env are stale
    }
    // Pass 1: copy the records into shared memory as they are and put the cells that hold anything on the live list.
    // ~99 % of the 128-bit vectors are zero (one ballot tells); for the others the list slots are claimed with ONE
    // shared-memory atomic per warp and component (ballot + popc ranking), not one per cell.
    const uint4* g4 = reinterpret_cast<const uint4*>(g);
    uint4* c4 = reinterpret_cast<uint4*>(s.cell);
    const int nvec = C >> 2;
    if (bar) mbar_wait(bar, 0);
    for (int j0 = tid - lane; j0 < nvec; j0 += nthr) {               // warp-uniform trip count
        const int j = j0 + lane;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (j < nvec) {
            if (bar) {
                v = c4[j];
            } else {
                asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(g4 + j));
                c4[j] = v;
            }
        }
        if (__ballot_sync(0xffffffffu, (v.x | v.y | v.z | v.w) != 0u) == 0u) continue;
        const uint32_t comp[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const bool nz = comp[c] != 0u;
            const unsigned m = __ballot_sync(0xffffffffu, nz);
            if (m == 0u) continue;
            int base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(&s.misc[1], __popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (nz) {
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < s.lcap) s.list[slot] = (unsigned short)(4 * j + c);
                else s.misc[8] = 1;
            }
        }
    }
    if (tid < (C & 3)) {
        const int q = (C & ~3) + tid;
        const uint32_t v = g[q];
        s.cell[q] = v;
        if (v != 0u) list_push(s, q);
    }
    __syncthreads();
    // Pass 2, every lane busy: the per-call bits of the listed records (owner / food when loaded, listed), the env's food
    // count and each snake's size.
    const LiveWalk lw = live_walk(s, C);
    for (int n = tid; n < lw.cnt; n += nthr) {
        const int q = live_at(s, lw, n);
        if (s.cell[q] == 0u) continue;                               // (only met when walking the whole grid)
        uint32_t rec = hbm_record(s.cell[q]) | kListed;
        // shadowed dense state: every cell the records name must still hold that value in the tensors (loads issued first,
        // compared after the shared-memory work of this iteration)
        float t_food = 1.0f, t_body = 0.0f;
        if (VERIFY) {
            if (rec & kFood) t_food = p.foods[(size_t)e * C + q];
            if (rec & kLive) t_body = p.bodies[((size_t)e * K + rec_owner(rec)) * C + q];
        }
        rec |= ((rec >> 16) & 63u) << 22;                            // owner when loaded
        if (rec & kFood) { rec |= kFood0; atomicAdd(&s.misc[0], 1); }
        if (rec & kLive) {
            const int k = rec_owner(rec), val = rec_value(rec);
            atomicMax(&s.size[k], val);
            if (CHECK) atomicAdd(&s.sum[k], val);
        }
        s.cell[q] = rec;
        if (VERIFY) ok &= t_food == 1.0f && (!(rec & kLive) || t_body == (float)rec_value(rec));
    }
    if (VERIFY && tid < K)                                           // a live snake's head cell holds its head; a dead snake is flagged done
        ok &= head_cell >= 0 ? (s.done[tid] == 0 && head_val == 1.0f) : (s.done[tid] != 0);
    return ok;
}

// SHADOWED DENSE STATE.  A caller who keeps the reference's fp32 tensors may let the library keep the records beside them
// (WurmMultiState.cells with cells_valid): a step then loads the records, checks that every cell they name still holds that
// food / body value in the tensors and every head cell its head -- PRESENCE is verified, a few scattered 4-byte loads; that
// nothing ELSE appeared in the tensors is the caller's word (the Python class gives it only while torch's version counters say
// the tensors were not written to) -- and writes its changes to both forms.  A mismatch re-loads the env from the tensors.
// (out of line, parameters by value: the cold path must not weigh on the kernel's code or pin its parameter struct)
__device__ __noinline__ void reload_from_tensors(const MultiParams p, unsigned char* smem_raw, int e) {
    const MultiSmem s = carve(smem_raw, p.C, p.K, true);
    const int C = p.C, K = p.K, tid = threadIdx.x, nthr = blockDim.x;
    for (int q = tid; q < C; q += nthr) s.cell[q] = 0u;
    if (tid < K) { s.hp[tid] = -1; s.size[tid] = 0; s.hcnt[tid] = 0; s.sum[tid] = 0; }
    if (tid == 0) { s.misc[0] = 0; s.misc[1] = 0; s.misc[3] = 0; s.misc[6] = 0; s.misc[8] = 0; s.misc[10] = 1; }
    __syncthreads();
    load_env<false, 4>(p, s, e);                                     // no hints: the heads tensor is scanned too
    __syncthreads();
}

// multi_snake.py:197-206: int16 colour of a body (or, is_head, head) cell of snake o
__device__ __forceinline__ void snake_rgb(const MultiSmem& s, int o, bool is_head, int rgb[3]) {
    float inten = 1.0f * 1.0f / 3.0f + (is_head ? 1.0f : 0.0f) * 1.0f / 3.0f;                  // :197
    inten *= 1.0f + 0.5f * (s.boost[o] ? 1.0f : 0.0f);                                      // :198
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = (int)(short)(inten * (float)s.col[3 * o + c]);      // :201-206
}

// multi_snake.py:194-227 _get_env_images: int16 colour of cell q (canonical state: one owner per cell)
__device__ __forceinline__ void env_pixel(const MultiParams& p, const MultiSmem& s, int q, int y, int x, int rgb[3]) {
    const uint32_t rec = s.cell[q];
    rgb[0] = rgb[1] = rgb[2] = 0;
    if (rec_body(rec)) {
        const int o = rec_owner(rec);
        snake_rgb(s, o, s.hp[o] == q, rgb);
    }
    if (rec & kFood) rgb[0] += 255;                                                          // :208-209
    if (rgb[0] == 0 && rgb[1] == 0 && rgb[2] == 0) rgb[0] = rgb[1] = rgb[2] = 255;           // :214-219
    if (y == 0 || x == 0 || y == p.S - 1 || x == p.S - 1) rgb[0] = rgb[1] = rgb[2] = 0;      // :225
}

// The partial observation shows, for almost every window cell, one of a handful of colours: black (border /
// outside), white (empty), red (food), or a snake's body / head colour.  The 2K snake colours are rendered once
// per env -- through the same expressions as env_pixel, so bit-identical -- and the per-cell work is a look-up.
__device__ __forceinline__ void build_colour_table(const MultiParams& p, const MultiSmem& s) {
    for (int t = threadIdx.x; t < 2 * p.K; t += blockDim.x) {
        int rgb[3];
        snake_rgb(s, t >> 1, (t & 1) != 0, rgb);
        if (rgb[0] == 0 && rgb[1] == 0 && rgb[2] == 0) rgb[0] = rgb[1] = rgb[2] = 255;       // :214-219
#pragma unroll
        for (int c = 0; c < 3; ++c) s.tab[3 * t + c] = div255(rgb[c]);
    }
}

// multi_snake.py:283-334 from the compact form into the per-agent buffers obs[k][e].
__device__ __forceinline__ void write_multi_obs(const MultiParams& p, const MultiSmem& s, int e) {
    const int C = p.C, K = p.K, S = p.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (p.obs_mode == WURM_MOBS_PARTIAL) {                           // :289-332
        const int n = p.obs_n, W = p.W, WW = W * W;
        __syncthreads();                                              // the live list (same memory) is dead from here on
        build_colour_table(p, s);
        __syncthreads();
        // colour of window cell (y, x); zero padding :301-302, black border :225
        auto pixel = [&](int y, int x, float& v0, float& v1, float& v2) {
            v0 = v1 = v2 = 0.0f;
            if ((unsigned)(y - 1) < (unsigned)(S - 2) && (unsigned)(x - 1) < (unsigned)(S - 2)) {
                const int q = y * S + x;
                const uint32_t rec = s.cell[q];
                if (!(rec & kFood)) {
                    v0 = v1 = v2 = 1.0f;                              // empty: white (255 / 255)
                    if (rec_body(rec)) {
                        const int ow = rec_owner(rec);
                        const float* t = s.tab + 3 * (2 * ow + (s.hp[ow] == q ? 1 : 0));
                        v0 = t[0]; v1 = t[1]; v2 = t[2];
                    }
                } else if (!rec_body(rec)) {
                    v0 = 1.0f;                                        // food: (255, 0, 0)
                } else {                                              // food under a body: the general expression
                    int rgb[3];
                    env_pixel(p, s, q, y, x, rgb);
                    v0 = div255(rgb[0]); v1 = div255(rgb[1]); v2 = div255(rgb[2]);            // :296
                }
            }
        };
        // (Unrolling the per-lane window cells so that the channel stores get constant offsets was tried and LOST on
        // B200 -- C4 0.340 -> 0.376 ms: the one-warp-per-env shape is bound by instruction fetch, and four inlined copies
        // of the cell colouring cost more than the address arithmetic they save.)
        for (int k = warp; k < K; k += nwarps) {
            float* o = p.obs + ((size_t)k * p.E + e) * 3 * WW;
            const int hp = s.hp[k];
            if (s.done[k] || hp < 0 || s.hcnt[k] > 1) {           // :320-323 zeros for dead agents
                for (int r = lane; r < 3 * WW; r += 32) o[r] = 0.0f;
                continue;
            }
            const int hy = fdiv(hp, p.magic_S), hx = hp - hy * S;
            for (int ij = lane; ij < WW; ij += 32) {              // one window cell per lane, three channel rows
                const int i = fdiv(ij, p.magic_W), j = ij - i * W;
                float v0, v1, v2;
                pixel(hy - n + i, hx - n + j, v0, v1, v2);
                o[ij] = v0; o[WW + ij] = v1; o[2 * WW + ij] = v2;
            }
        }
    } else if (p.obs_mode == WURM_MOBS_FULL) {                       // :268-281 _observe_agent
        for (int k = warp; k < K; k += nwarps) {
            float* o = p.obs + ((size_t)k * p.E + e) * 3 * C;
            for (int q = lane; q < C; q += 32) {
                const int y = fdiv(q, p.magic_S), x = q - y * S;
                const uint32_t rec = s.cell[q];
                int r = 255, g = 255, b = 255;
                if (rec & kFood) { r = 255; g = 0; b = 0; }
                if (rec_body(rec)) {
                    const int ow = rec_owner(rec);
                    const bool is_head = s.hp[ow] == q;
                    if (ow == k) { r = 0; g = is_head ? 192 : 96; b = 0; }
                    else { r = 0; g = 0; b = is_head ? 192 : 96; }
                }
                if (y == 0 || x == 0 || y == S - 1 || x == S - 1) r = g = b = 0;
                o[q] = div255(r);
                o[C + q] = div255(g);
                o[2 * C + q] = div255(b);
            }
        }
    }
    if (p.img) {                                                     // _get_env_images as (E,3,S,S) int16
        for (int q = threadIdx.x; q < C; q += blockDim.x) {
            const int y = fdiv(q, p.magic_S), x = q - y * S;
            int rgb[3];
            env_pixel(p, s, q, y, x, rgb);
            for (int c = 0; c < 3; ++c) p.img[((size_t)e * 3 + c) * C + q] = (short)rgb[c];
        }
    }
}

// Block-wide uniform choice among the cells q < C with pred(q), in raster order: every thread counts its
// contiguous chunk, a shuffle scan ranks the chunks, `rnd` picks a rank and the owning thread finds the
// cell (written to *out).  Returns the number of candidate cells.  `counts`: >= 33 ints of shared scratch.
template <typename P>
__device__ __forceinline__ int block_pick(int C, int* counts, uint32_t rnd, int* out, P&& pred) {
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int chunk = (C + nthr - 1) / nthr, q0 = tid * chunk, q1 = min(C, q0 + chunk);
    int mine = 0;
    for (int q = q0; q < q1; ++q) mine += pred(q) ? 1 : 0;
    int incl = mine;                                                  // inclusive scan within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) counts[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nwarps; ++w) {
        const int c = counts[w];
        if (w < warp) before += c;
        total += c;
    }
    const int first = before + incl - mine;                           // rank of this thread's first candidate
    if (total > 0) {
        const int r = (int)bounded(rnd, (uint32_t)total);
        if (r >= first && r < first + mine) {
            int left = r - first;
            for (int q = q0; q < q1; ++q)
                if (pred(q) && left-- == 0) { *out = q; break; }
        }
    }
    __syncthreads();
    return total;
}

// ---- pieces of the reset shared by the stand-alone kernel and the fused step+reset path ----
struct ResetScratch {
    uint8_t* occ;      // C occupancy bytes
    uint8_t* vcount;   // seed candidates per 16-byte vector of occ (decide_recreate keeps them current while it places snakes)
    int* lcount;       // ... and per lane of warp 0 (32)
    int* counts;       // block_pick scratch (one int per warp)
    int* pick;         // [0] chosen cell
    int* snake_cell;   // K seed cells of a re-created env
    int* snake_dir;    // K directions
};

static size_t reset_scratch_bytes(int C) { return (size_t)reset_scratch_size(C); }

// (`counts_base`: reset_counts_size(C) bytes, 16-byte aligned, for vcount / lcount)
__device__ __forceinline__ ResetScratch carve_reset(unsigned char* base, int C, unsigned char* counts_base) {
    ResetScratch r;
    const int bytes16 = (C + 15) & ~15;
    r.occ = base;
    r.vcount = counts_base;
    r.lcount = reinterpret_cast<int*>(counts_base + (((bytes16 >> 4) + 15) & ~15));
    r.counts = reinterpret_cast<int*>(base + bytes16);
    r.pick = r.counts + 16;
    r.snake_cell = r.pick + 4;
    r.snake_dir = r.snake_cell + 32;
    return r;
}

// :846-858 / :927-941: a snake may be seeded on cell q if q is at least two cells from the wall and nothing
// occupies its 3x3 neighbourhood.  The occupancy bytes carry that test precomputed: bit 0 = the cell is occupied,
// bit 1 = "no seed here" (too close to the wall, or an occupied cell in the 3x3 neighbourhood), so that the
// candidate scans of block_pick read one byte per cell instead of nine.
constexpr uint8_t kOccupied = 1, kNoSeed = 2, kBorder = 4;        // (kBorder: decide_recreate's map only -- no food on the wall)
__device__ __forceinline__ bool spawnable(const uint8_t* occ, int q) { return !(occ[q] & kNoSeed); }
__device__ __forceinline__ bool seed_margin(const MultiParams& p, int q) {
    const int S = p.S, y = fdiv(q, p.magic_S), x = q - y * S;
    return y < 2 || y > S - 3 || x < 2 || x > S - 3;
}
// ORs bits into occupancy byte q; several threads may hit one byte (or its word) at once: a shared-memory atomic on
// the enclosing 32-bit word (the map starts 16-byte aligned)
__device__ __forceinline__ uint8_t occ_or(uint8_t* occ, int q, uint8_t bits) {        // returns the byte as it was
    return (uint8_t)(atomicOr(reinterpret_cast<unsigned*>(occ + (q & ~3)), (unsigned)bits << (8 * (q & 3))) >> (8 * (q & 3)));
}
// marks neighbour nb (0..8, row-major 3x3) of occupied cell q
// (returns the neighbour's cell if this call is what took it out of the seed candidates, -1 otherwise)
__device__ __forceinline__ int block_around(const MultiParams& p, uint8_t* occ, int q, int nb) {
    const int S = p.S, y = fdiv(q, p.magic_S) + nb / 3 - 1, x = q - fdiv(q, p.magic_S) * S + nb % 3 - 1;
    if (y >= 0 && y < S && x >= 0 && x < S && !(occ_or(occ, y * S + x, kNoSeed) & kNoSeed)) return y * S + x;
    return -1;
}
// completes an occupancy map whose bytes hold bit 0 only: wall margin and 3x3 dilation into bit 1.  Whole CTA.
__device__ __forceinline__ void finish_occupancy(const MultiParams& p, uint8_t* occ) {
    for (int q = threadIdx.x; q < p.C; q += blockDim.x) {
        // (the byte is read through the atomic too: neighbours may be ORing into the same word right now)
        const unsigned w = atomicOr(reinterpret_cast<unsigned*>(occ + (q & ~3)), seed_margin(p, q) ? (unsigned)kNoSeed << (8 * (q & 3)) : 0u);
        if ((w >> (8 * (q & 3))) & kOccupied)
            for (int nb = 0; nb < 9; ++nb) block_around(p, occ, q, nb);
    }
    __syncthreads();
}

// cells of a length-3 snake seeded at `cell` facing d (LENGTH_3_SNAKES): tail 1, seed 2, head 3
__device__ __forceinline__ void snake_cells(const MultiParams& p, int cell, int d, int& tl, int& hd) {
    const int S = p.S, y = fdiv(cell, p.magic_S), x = cell - y * S;
    hd = (y + off_y(d)) * S + (x + off_x(d));
    tl = (y - off_y(d)) * S + (x - off_x(d));
}

// block_pick's choice -- the bounded(rnd, total)-th cell, in raster order, whose occupancy byte has none of the bits `mask`
// (any of bits 0-3) -- made by ONE warp without a block barrier.  Lane l owns the contiguous span of 16-byte vectors
// [l * per_lane, (l + 1) * per_lane) of the map (`bytes16`: its length rounded up to 16; the padding bytes carry every bit:
// never candidates).  All 32 lanes must call; all get the cell (or -1).
// The K snakes of a re-created env are placed one after the other, each on the map the previous one left: as block-wide picks
// with ~6 block barriers each that was the longest dependent chain of the whole reset (~60 us at 16 snakes on a 64 x 64 grid,
// and every launch lasts as long as its unluckiest CTA).  COUNTED: the candidates per vector (sc.vcount) and per lane
// (sc.lcount) are already known -- decide_recreate counts once and then only subtracts what each new snake blocks -- so a pick
// is one shuffle scan plus the owner's walk over its <= per_lane byte counts.
__device__ __forceinline__ uint32_t occ_candidates(uint32_t w, uint32_t m4) {      // bit 8b = byte b of w is a candidate
    uint32_t z = w & m4;
    z |= z >> 1; z |= z >> 2;
    return ~z & 0x01010101u;
}
__device__ __forceinline__ int occ_count(const uint4& v, uint32_t m4) {
    return __popc(occ_candidates(v.x, m4)) + __popc(occ_candidates(v.y, m4)) + __popc(occ_candidates(v.z, m4)) +
           __popc(occ_candidates(v.w, m4));
}
template <bool COUNTED>
__device__ __forceinline__ int warp_pick(const ResetScratch& sc, int bytes16, uint32_t mask, uint32_t rnd) {
    const int lane = threadIdx.x & 31;
    const uint4* o4 = reinterpret_cast<const uint4*>(sc.occ);
    const int nvec = bytes16 >> 4, per_lane = (nvec + 31) >> 5;
    const int j0 = lane * per_lane, j1 = min(nvec, j0 + per_lane);
    const uint32_t m4 = mask * 0x01010101u;
    int mine = 0;
    if (COUNTED) mine = sc.lcount[lane];
    else for (int j = j0; j < j1; ++j) mine += occ_count(o4[j], m4);     // (independent loads: they pipeline)
    int incl = mine;                                                  // inclusive scan over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return -1;
    const int r = (int)bounded(rnd, (uint32_t)total);
    int cell = -1;
    if (r >= incl - mine && r < incl) {                               // the owner: which vector, which word, which byte
        int left = r - (incl - mine);
        for (int j = j0; j < j1 && cell < 0; ++j) {
            if (COUNTED) {
                const int n = sc.vcount[j];
                if (left >= n) { left -= n; continue; }
            }
            const uint4 v = o4[j];
            const uint32_t cw[4] = {occ_candidates(v.x, m4), occ_candidates(v.y, m4), occ_candidates(v.z, m4), occ_candidates(v.w, m4)};
#pragma unroll
            for (int wd = 0; wd < 4; ++wd) {
                const int n = __popc(cw[wd]);
                if (cell < 0 && left < n) cell = 16 * j + 4 * wd + (int)(__fns(cw[wd], 0, left + 1) >> 3);
                left -= n;
            }
        }
    }
    const unsigned owner = __ballot_sync(0xffffffffu, cell >= 0);
    return __shfl_sync(0xffffffffu, cell, __ffs(owner) - 1);
}

// _create_envs (:996-1019): K snakes placed one after the other (_add_snake :911-994) and one food, decided on a
// fresh occupancy map; results in sc.snake_cell / sc.snake_dir and the returned food cell.  Whole CTA calls; warp 0 places.
__device__ __forceinline__ int decide_recreate_inline(const MultiParams& p, int e, uint64_t ctr, const ResetScratch& sc) {
    const int C = p.C, K = p.K, S = p.S, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    const int bytes16 = (C + 15) & ~15;
    for (int q = tid; q < bytes16; q += nthr) {
        uint8_t b = 0xff;                                             // padding: never a candidate
        if (q < C) {
            const int y = fdiv(q, p.magic_S), x = q - y * S;
            b = (seed_margin(p, q) ? kNoSeed : 0) | ((y < 1 || y > S - 2 || x < 1 || x > S - 2) ? kBorder : 0);
        }
        sc.occ[q] = b;
    }
    __syncthreads();
    if (tid < 32) {
        const int nvec = bytes16 >> 4, per_lane = (nvec + 31) >> 5;
        if (!p.create) {                                              // seed candidates per vector and per lane, counted once
            const uint4* o4 = reinterpret_cast<const uint4*>(sc.occ);
            int mine = 0;
            for (int j = lane * per_lane; j < min(nvec, (lane + 1) * per_lane); ++j) {
                const int n = occ_count(o4[j], kNoSeed * 0x01010101u);
                sc.vcount[j] = (uint8_t)n;
                mine += n;
            }
            sc.lcount[lane] = mine;
            __syncwarp();
        }
        for (int k = 0; k < K; ++k) {
            int cell, d;
            if (p.create) {
                cell = p.create[((size_t)e * (K + 1) + k) * 2];
                d = p.create[((size_t)e * (K + 1) + k) * 2 + 1];
            } else {
                const uint4 r = draw(p.seed, ctr, (uint32_t)e, kStreamMultiCreateSnake | ((uint32_t)k << 4));
                cell = warp_pick<true>(sc, bytes16, kNoSeed, r.x);
                d = (int)(r.y >> 30);
            }
            if (lane == 0) {
                sc.snake_cell[k] = cell; sc.snake_dir[k] = d;
                if (cell < 0) atomicOr(p.status, WURM_ST_NO_SPAWN);   // the reference raises (:947)
            }
            if (cell >= 0 && lane < 27) {                             // the new snake's three cells and their surroundings
                int tl, hd;
                snake_cells(p, cell, d, tl, hd);
                const int c = lane < 9 ? tl : lane < 18 ? cell : hd;
                if (lane % 9 == 4) occ_or(sc.occ, c, kOccupied);
                const int gone = block_around(p, sc.occ, c, lane % 9);     // a cell this snake takes out of the seed candidates
                if (gone >= 0 && !p.create) {
                    const int j = gone >> 4;
                    atomicSub(reinterpret_cast<unsigned*>(sc.vcount + (j & ~3)), 1u << (8 * (j & 3)));       // (the byte is >= 1: no borrow)
                    atomicSub(&sc.lcount[j / per_lane], 1);
                }
            }
            __syncwarp();
        }
        int fcell;                                                    // :1016 one food on a free interior cell
        if (p.create) fcell = p.create[((size_t)e * (K + 1) + K) * 2];
        else fcell = warp_pick<false>(sc, bytes16, kOccupied | kBorder, draw(p.seed, ctr, (uint32_t)e, kStreamMultiCreateFood).x);
        if (lane == 0) sc.pick[0] = fcell;
    }
    __syncthreads();
    return sc.pick[0];
}

// Out of line, plain arguments (a reference to the kernel's parameter struct would force a copy of it onto the stack): a
// fraction of a per cent of the envs are re-created per step, and the step kernels are sensitive to the size of their code.
__device__ __noinline__ int decide_recreate_cold(int C, int K, int S, uint32_t magic_S, const int32_t* create, uint64_t seed, int32_t* status,
                                                 int e, uint64_t ctr, ResetScratch sc) {
    MultiParams q = {};
    q.C = C; q.K = K; q.S = S; q.magic_S = magic_S; q.create = create; q.seed = seed; q.status = status;
    return decide_recreate_inline(q, e, ctr, sc);
}
__device__ __forceinline__ int decide_recreate(const MultiParams& p, int e, uint64_t ctr, const ResetScratch& sc) {
    return decide_recreate_cold(p.C, p.K, p.S, p.magic_S, p.create, p.seed, p.status, e, ctr, sc);
}

// Stores the cells of the re-created env's snakes and food and revives its agents (:790-798).  Whole CTA; the
// caller has zeroed (or is zeroing, to the same values) whatever else the tensors held.
// (rec_out / dense_out: into the records and / or the reference's tensors -- compile-time constants in the step kernels)
__device__ __forceinline__ void write_recreated(const MultiParams& p, int e, const ResetScratch& sc, int fcell, bool rec_out, bool dense) {
    const int C = p.C, K = p.K, tid = threadIdx.x;
    uint32_t* gc = rec_out ? p.cells + (size_t)e * p.Cp : nullptr;
    if (tid == 0 && fcell >= 0) {
        if (gc) gc[fcell] = kFood;
        if (dense) p.foods[(size_t)e * C + fcell] = 1.0f;
    }
    if (tid < K) {
        const size_t n = (size_t)e * K + tid;
        if (sc.snake_cell[tid] >= 0) {
            int tl, hd;
            snake_cells(p, sc.snake_cell[tid], sc.snake_dir[tid], tl, hd);
            if (gc) { gc[hd] = make_rec(tid, 3); gc[sc.snake_cell[tid]] = make_rec(tid, 2); gc[tl] = make_rec(tid, 1); }
            if (dense) {
                p.heads[n * C + hd] = 1.0f;
                p.bodies[n * C + hd] = 3.0f; p.bodies[n * C + sc.snake_cell[tid]] = 2.0f; p.bodies[n * C + tl] = 1.0f;
            }
            p.orientations[n] = sc.snake_dir[tid];                    // :793
            if (p.head_hints) p.head_hints[n] = (short)hd;
        } else if (p.head_hints) {
            p.head_hints[n] = gc ? -1 : -2;
        }
        p.dones[n] = 0;                                               // :798
    }
}

// :805-829 + _get_snake_addition :838-909: seed cell (or -1) and direction for the env's first dead snake, on the
// occupancy map the caller prepared.  Whole CTA.
__device__ __forceinline__ int decide_respawn(const MultiParams& p, int e, uint64_t ctr, const ResetScratch& sc, int& d) {
    if (threadIdx.x == 0) sc.pick[0] = -1;
    __syncthreads();
    if (p.respawn) {
        if (threadIdx.x == 0) sc.pick[0] = p.respawn[2 * (size_t)e];
        d = p.respawn[2 * (size_t)e + 1];
        __syncthreads();
    } else {
        finish_occupancy(p, sc.occ);
        const uint4 r = draw(p.seed, ctr, (uint32_t)e, kStreamMultiRespawn);
        block_pick(p.C, sc.counts, r.x, sc.pick, [&](int q) { return spawnable(sc.occ, q); });
        d = (int)(r.y >> 30);
    }
    return sc.pick[0];
}

// one thread: the respawned snake's cells, orientation and done flag (:826-829)
__device__ __forceinline__ void write_respawned(const MultiParams& p, int e, int k, int cell, int d, bool rec_out, bool dense_out) {
    const size_t n = (size_t)e * p.K + k;
    if (cell >= 0) {
        int tl, hd;
        snake_cells(p, cell, d, tl, hd);
        if (rec_out) {
            uint32_t* gc = p.cells + (size_t)e * p.Cp;
            gc[hd] = make_rec(k, 3); gc[cell] = make_rec(k, 2); gc[tl] = make_rec(k, 1);
        }
        if (dense_out) {
            p.heads[n * p.C + hd] = 1.0f;
            p.bodies[n * p.C + hd] = 3.0f; p.bodies[n * p.C + cell] = 2.0f; p.bodies[n * p.C + tl] = 1.0f;
        }
    }
    p.orientations[n] = d;                                            // :828 even when the spawn failed
    p.dones[n] = cell < 0;                                            // :829
    if (p.head_hints) {
        int tl, hd = -1;
        if (cell >= 0) snake_cells(p, cell, d, tl, hd);
        p.head_hints[n] = (short)hd;
    }
}

// get_n_colours (:163-169) for one dead snake (:800-803)
__device__ __forceinline__ void recolour(const MultiParams& p, int e, int k, uint64_t ctr) {
    short* col = p.colours + 3 * ((size_t)e * p.K + k);
    if (p.colours_replay) {
        for (int c = 0; c < 3; ++c) col[c] = p.colours_replay[3 * ((size_t)e * p.K + k) + c];
    } else {
        const uint4 r = draw(p.seed, ctr, (uint32_t)e, kStreamMultiColour | ((uint32_t)k << 4));
        const float c0 = unit_float(r.x) / 1.5f, c1 = unit_float(r.y), c2 = unit_float(r.z);
        const float norm = sqrtf(c0 * c0 + c1 * c1 + c2 * c2);
        col[0] = (short)(c0 / norm * 192.0f); col[1] = (short)(c1 / norm * 192.0f); col[2] = (short)(c2 / norm * 192.0f);
    }
}

// The new state of env e goes back as records: the listed records that changed (4 bytes each; `all`: every record, after a
// load from the dense tensors); the head cells go into head_hints with the per-agent outputs.
__device__ __forceinline__ void records_writeback(const MultiParams& p, const MultiSmem& s, int e, bool all) {
    const int C = p.C, tid = threadIdx.x, nthr = blockDim.x;
    uint32_t* gc = p.cells + (size_t)e * p.Cp;
    if (all) {
        for (int q = tid; q < p.Cp; q += nthr) gc[q] = q < C ? hbm_record(s.cell[q]) : 0u;
        return;
    }
    const LiveWalk lw = live_walk(s, C);
    for (int n = tid; n < lw.cnt; n += nthr) {
        const int q = live_at(s, lw, n);
        const uint32_t rec = s.cell[q];
        if ((rec & kDirty) || (((rec >> 29) ^ (rec >> 30)) & 1u)) gc[q] = hbm_record(rec);
    }
}

// ... and / or into the reference's fp32 tensors.  The compact form knows exactly which cells changed: only those are stored
// (a step touches O(snake length) cells of an S*S grid: a few sectors per snake instead of the full state); an input the
// records could not carry exactly (misc[3]) is expanded densely instead, which normalises it.
__device__ __forceinline__ void dense_writeback(const MultiParams& p, const MultiSmem& s, int e, bool valid, int k, int a_hp0, int a_hp) {
    const int C = p.C, K = p.K, tid = threadIdx.x, nthr = blockDim.x;
    if (s.misc[3]) {
        store_floats(p.foods + (size_t)e * C, C, [&](int i) { return (s.cell[i] & kFood) ? 1.0f : 0.0f; });
        store_floats(p.heads + (size_t)e * K * C, K * C, [&](int i) {
            const int kk = fdiv_C(i, C, p.magic_C);
            return s.hp[kk] == i - kk * C ? 1.0f : 0.0f;
        });
        store_floats(p.bodies + (size_t)e * K * C, K * C, [&](int i) {
            const int kk = fdiv_C(i, C, p.magic_C);
            const uint32_t rec = s.cell[i - kk * C];
            return (rec_body(rec) && rec_owner(rec) == kk) ? (float)rec_value(rec) : 0.0f;
        });
        return;
    }
    float* bodies = p.bodies + (size_t)e * K * C;
    const LiveWalk lw = live_walk(s, C);
    for (int n = tid; n < lw.cnt; n += nthr) {
        const int q = live_at(s, lw, n);
        const uint32_t rec = s.cell[q];
        if (rec & kDirty) {
            const int ko = rec_owner0(rec), kn = rec_body(rec) ? rec_owner(rec) : -1;
            if (ko >= 0 && ko != kn) bodies[(size_t)ko * C + q] = 0.0f;
            if (kn >= 0) bodies[(size_t)kn * C + q] = (float)rec_value(rec);
        }
        if (((rec >> 29) ^ (rec >> 30)) & 1u) p.foods[(size_t)e * C + q] = (rec & kFood) ? 1.0f : 0.0f;
    }
    if (valid && a_hp0 != a_hp) {
        float* head = p.heads + ((size_t)e * K + k) * C;
        if (a_hp0 >= 0) head[a_hp0] = 0.0f;
        if (a_hp >= 0) head[a_hp] = 1.0f;
    }
}

// One CTA = one environment.  STEP: load -> step -> store -> observe.  !STEP: load -> observe.
// THREADS grows with the grid (~24 cells per thread at most, measured sweep in profiles/r01_sweep_multi.txt): for
// small grids one warp per env, so that the ~12 barriers of the step cost nothing, no warp idles while the
// lane-per-snake logic runs and 32 envs stay resident per SM (K=4, S=25: 0.466 ms per launch against 0.489 /
// 0.533 ms with 64 / 128 threads; staging the raw env through shared memory with TMA was tried and lost to the
// occupancy it costs); 256 threads for S=64, where streaming 540 KB per env wants the loads of many threads in flight.
// COMPACT: the records are loaded instead of the tensors; SHADOW (with COMPACT): the tensors exist too -- the records are verified
// against them and every change goes to both forms (template parameters: the record-only kernel is bound by instruction issue
// and fetch, and carrying the shadow's code as run-time branches cost it 15 %).
template <bool STEP, int THREADS, bool COMPACT = false, bool SHADOW = false>
// (the compact 128-thread shape serves grids from 56 x 56 up, where shared memory allows ~10 CTAs per SM anyway)
#ifndef WURM_SHADOW_CTAS_128
#define WURM_SHADOW_CTAS_128 9
#endif
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 4 : THREADS == 128 ? (SHADOW ? WURM_SHADOW_CTAS_128 : COMPACT ? 10 : 12) : THREADS == 64 ? 20 : 32)
multi_env_kernel(const MultiParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const MultiSmem s = carve(smem_raw, p.C, p.K, COMPACT);
    const int C = p.C, K = p.K, S = p.S;
    const int e = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;

    // COMPACT: the env's records set off for shared memory first thing, as ONE bulk copy (TMA) straight into s.cell; everything
    // up to the load's record walk -- hints, flags, colours, the first barrier -- runs under its latency.  (Before: each thread
    // pulled its C / (4 * THREADS) vectors one after the other, every ballot waiting for its own load: 21 % of the C5 kernel's
    // stall samples sat on that loop, profiles/r02_ncu_C5_compact.txt.)
    __shared__ __align__(8) uint64_t rec_bar;
    if (COMPACT && tid == 0) {
        mbar_init(&rec_bar, 1);
        fence_mbar_init();
        const uint32_t bytes = (uint32_t)(C >> 2) * 16u;
        mbar_arrive_expect_tx(&rec_bar, bytes);
        bulk_load(s.cell, p.cells + (size_t)e * p.Cp, bytes, &rec_bar);
    }

    // per-snake scalars (action, orientation, boost-cost draw) are fetched by warp 0 before the env is streamed
    // in, so that their latency hides behind the load instead of heading the serial per-snake logic
    // (measured on B200: +2 % at K=16,S=64 with 256 threads; a loss for the small-grid variant, where the values are
    // fetched right before use instead)
    // Head hints (see load_env): this thread's snake's hint and done flag, fetched before anything else so that the
    // dependent look-up into the heads tensor can be issued ahead of the scans
    int hint_h = -1;
    bool hint_dead = false;
    const bool use_hints = (!COMPACT || SHADOW) && p.head_hints != nullptr;     // (SHADOW: always there, and part of the state)
    if (use_hints && tid < K) {
        hint_h = p.head_hints[(size_t)e * K + tid];
        hint_dead = p.dones[(size_t)e * K + tid] != 0;
    }
    // the call counter is read where a draw needs it (rare: something died or was eaten); hoisting the load to the top
    // of the kernel was measured and lost (a register held across the whole step)
#define ctr_now call_counter(p)
    // (An L2 prefetch of the env's foods / bodies lines ahead of the load scan was measured for the small-CTA shapes:
    // C4 0.3397 -> 0.3357 ms, within noise of not being worth the code; profiles/r02_multi_compact_experiments.txt.)
    constexpr bool kPrefetch = THREADS == 256 || (COMPACT && THREADS >= 64);
    long long pre_action = 0, pre_orient = 0;
    float pre_cost = 0.0f;
    if (STEP && kPrefetch && tid < K) {
        const size_t n = (size_t)e * K + tid;
        pre_action = load_action(p.actions[tid], p.action_bytes, (size_t)e);
        pre_orient = p.orientations[n];
        if (p.replay && p.u_cost) pre_cost = p.u_cost[n];
    }
    if (!COMPACT) {                                                   // (the compact load writes every record itself)
        uint4* c4 = reinterpret_cast<uint4*>(s.cell);                 // 16-byte aligned: the start of the dynamic shared memory
        for (int j = tid; j < (C >> 2); j += nthr) c4[j] = make_uint4(0u, 0u, 0u, 0u);
        if (tid < (C & 3)) s.cell[(C & ~3) + tid] = 0u;
    }
    if (tid < K) {
        s.hp[tid] = -1; s.size[tid] = 0; s.hcnt[tid] = 0; s.decay[tid] = 0; s.cost[tid] = 0; s.sum[tid] = 0;
        s.done[tid] = p.dones[(size_t)e * K + tid] != 0;
        s.boost[tid] = !STEP ? (p.boost_this_step[(size_t)e * K + tid] != 0) : 0;
    }
    if (tid < 12) s.misc[tid] = 0;
    for (int t = tid; t < 3 * K; t += nthr) s.col[t] = p.colours[(size_t)e * K * 3 + t];
    // Head hints: the heads tensor's value at this thread's snake's hinted cell (the hint itself was fetched first thing)
    float hint_val = 0.0f;
    if (use_hints && tid < K && hint_h >= 0 && hint_h < C) hint_val = p.heads[((size_t)e * K + tid) * C + hint_h];
    __syncthreads();
    if (COMPACT) {
        if (SHADOW) {                                                 // shadowed dense state: the records must still match the tensors
            const bool ok = load_env_compact<false, true>(p, s, e, hint_h, hint_val, &rec_bar);
            if (__syncthreads_or(!ok)) reload_from_tensors(p, smem_raw, e);     // (both ways end on a barrier)
        } else {
            load_env_compact<false>(p, s, e, -1, 1.0f, &rec_bar);
            __syncthreads();
        }
    } else {
        load_env<false, (THREADS >= 256 ? 8 : 4)>(p, s, e, use_hints, hint_h, hint_val, hint_dead);
        __syncthreads();
    }

    if (STEP) {
        // ---- per-snake registers, live on lane k of warp 0 ----
        const int k = lane;
        const bool valid = (warp == 0) && (k < K);
        int a_hp = -1, a_hp0 = -1, a_size = 0, a_mv = 0;
        bool a_done = true, a_done0 = true, a_boosted = false, a_scol = false, a_ecol = false;
        float a_reward = 0.0f, a_foodc = 0.0f;
        bool run_boost = false;
        if (warp == 0) {
            if (valid) {
                const size_t n = (size_t)e * K + k;
                if (!kPrefetch) {
                    pre_action = load_action(p.actions[k], p.action_bytes, (size_t)e);
                    pre_orient = p.orientations[n];
                }
                const long long a = pre_action;
                a_hp = a_hp0 = s.hp[k]; a_size = s.size[k];
                a_done = a_done0 = s.done[k] != 0;                    // :490
                long long m = a % 4;                                  // :483
                if (pre_orient == m) m = (m + 2) % 4;                 // :336-339
                a_mv = (int)m;
                p.orientations[n] = (m + 2) % 4;                      // :355-357 (dead agents too)
                a_boosted = (a > 3) && (a_size >= 4);                 // :484,497-498
                p.boost_this_step[n] = a_boosted;                     // :499
                s.boost[k] = a_boosted;
                if (s.hcnt[k] > 1) atomicOr(p.status, WURM_ST_MULTI_HEAD);
            }
            const bool any_boost = __ballot_sync(0xffffffffu, valid && a_boosted) != 0u;
            run_boost = p.replay ? (p.boost_phase_ran != 0) : (p.boost && any_boost);     // :503
            if (lane == 0) s.misc[2] = run_boost;
        }
        __syncthreads();
        run_boost = s.misc[2] != 0;

        for (int phase = run_boost ? 0 : 1; phase < 2; ++phase) {
            const bool boost_phase = phase == 0;
            const bool active = valid && (boost_phase ? a_boosted : true);
            bool ov = false;
            if (warp == 0) {
                if (active && a_hp >= 0) {                            // :509 / :613 _move_heads
                    const int y = fdiv(a_hp, p.magic_S), x = a_hp - y * S;
                    const int ny = y - off_y(a_mv), nx = x - off_x(a_mv);
                    a_hp = (ny >= 0 && ny < S && nx >= 0 && nx < S) ? ny * S + nx : -1;
                }
                if (valid) s.hp[k] = a_hp;
                __syncwarp();
                ov = valid && a_hp >= 0 && (s.cell[a_hp] & kFood);    // :514 / :618 overlap of ALL heads
                __syncwarp();
                if (ov) clear_food(s, a_hp);                          // :517 / :622
                if (valid) s.decay[k] = active && !ov;                // :523-526 / :627-628
                if (active && ov) { a_reward += 1.0f; a_foodc += 1.0f; }   // :527-529 / :629-631
            }
            __syncthreads();
            const LiveWalk lw = live_walk(s, C);
            for (int n = tid; n < lw.cnt; n += nthr) {  // _decay_bodies :362-363
                const int q = live_at(s, lw, n);
                const uint32_t rec = s.cell[q];
                if (rec_body(rec) && s.decay[rec_owner(rec)])
                    s.cell[q] = ((rec_value(rec) == 1) ? (rec & ~kLive) : rec - 1u) | kDirty;
            }
            __syncthreads();
            if (warp == 0) {
                bool col = false;
                if (active && a_hp >= 0) {                            // :534-545 / :636-642
                    col = rec_body(s.cell[a_hp]);                     // any body, own included
                    for (int j = 0; j < K; ++j) col |= (j != k) && (s.hp[j] == a_hp);   // another head
                }
                __syncwarp();
                if (active) { a_done |= col; a_scol |= col; }         // :546-547 / :643-644
                if (active && a_hp >= 0) {                            // :553 / :650 growth at the head cell
                    const uint32_t add = (uint32_t)(a_size + (ov ? 1 : 0));
                    const uint32_t cur = s.cell[a_hp];
                    if (!rec_body(cur)) {                             // head-on: first claim wins
                        if (atomicCAS(&s.cell[a_hp], cur, cur | make_rec(k, (int)add) | kDirty | kListed) == cur && !(cur & kListed))
                            list_push(s, a_hp);
                    }
                    else if (rec_owner(cur) == k) s.cell[a_hp] = (cur + add) | kDirty;   // self collision: values add up
                    // a collider's head value on ANOTHER snake's cell is dropped: the collider is
                    // deleted below and that cell counts as covered by the other body either way
                    a_size += ov ? 1 : 0;                             // :555 / :652
                    const int y = fdiv(a_hp, p.magic_S), x = a_hp - y * S;
                    const bool edge = y == 0 || x == 0 || y == S - 1 || x == S - 1;   // :560 / :657
                    a_done |= edge; a_ecol |= edge;
                }
                if (valid) s.done[k] = a_done;
                if (boost_phase) {                                    // :579-592 boost cost
                    bool cost = false;
                    if (valid && a_boosted) {
                        const float u = p.replay ? (kPrefetch ? pre_cost : p.u_cost[(size_t)e * K + k])
                                                 : unit_float(draw_i(p.seed, ctr_now, (uint32_t)e, kStreamMultiBoostCost, (uint32_t)k));
                        cost = u < p.boost_cost_prob;
                    }
                    if (valid) s.cost[k] = cost;
                    if (cost) { a_reward -= 1.0f; a_size -= 1; }      // :590-591
                }
            }
            __syncthreads();
            {   // food from dead bodies (:416-428), boost cost on the bodies (:583-589), deletion (:595 / :676)
                const float* U = p.replay ? (boost_phase ? p.u_boost : p.u_reg) : nullptr;
                const uint32_t stream = boost_phase ? kStreamMultiDeathBoost : kStreamMultiDeathRegular;
                const LiveWalk lw = live_walk(s, C);
                for (int n = tid; n < lw.cnt; n += nthr) {
                    const int q = live_at(s, lw, n);
                    uint32_t rec = s.cell[q];
                    if (!rec_body(rec)) continue;
                    const int o = rec_owner(rec);
                    bool food = false;
                    if (p.food_on_death && s.done[o]) {
                        const int y = fdiv(q, p.magic_S), x = q - y * S;
                        if (!(y == 1 || x == 0 || y == S - 1 || x == S - 1)) {   // sic: row 1 (:418)
                            const float u = p.replay ? (U ? U[(size_t)e * C + q] : 0.0f)
                                                     : unit_float(draw_i(p.seed, ctr_now, (uint32_t)e, stream, (uint32_t)q));
                            food = u > p.death_thr;
                        }
                    }
                    if (boost_phase && s.cost[o]) {
                        if (rec_value(rec) == 1) { food = true; rec = (rec & ~kLive) | kDirty; }   // the tail becomes food
                        else rec = (rec - 1u) | kDirty;
                    }
                    if (s.done[o]) rec = (rec & ~kLive) | kDirty;
                    s.cell[q] = rec;
                    if (food) set_food(s, q);
                }
            }
            if (warp == 0 && valid && a_done) { a_hp = -1; s.hp[k] = -1; }
            __syncthreads();
        }

        // ---- _add_food (:368-410) ----
        {
            const int nfood = s.misc[0];
            auto cell_free = [&](int q) { return (s.cell[q] & (kLive | kFood)) == 0u; };
            if (p.food_mode == 0) {                                   // only_one
                if (nfood == 0 && tid == 0) {
                    int cell = -1;
                    if (p.replay) cell = p.food_cell[e];
                    else {
                        const int I = S - 2;
                        for (uint32_t t = 0; t < kRejectionTries && cell < 0; ++t) {
                            const int cand = (int)bounded(draw_i(p.seed, ctr_now, (uint32_t)e, kStreamMultiFoodOne, t), (uint32_t)(I * I));
                            const int cy = cand / I, q = (1 + cy) * S + 1 + (cand - cy * I);
                            if (cell_free(q)) cell = q;
                        }
                        if (cell < 0) {                               // nearly full board: rank the free cells
                            int nfree = 0;
                            for (int y = 1; y < S - 1; ++y)
                                for (int x = 1; x < S - 1; ++x) nfree += cell_free(y * S + x);
                            if (nfree > 0) {
                                int r = (int)bounded(draw_i(p.seed, ctr_now, (uint32_t)e, kStreamMultiFoodOne, kRejectionTries), (uint32_t)nfree);
                                for (int y = 1; y < S - 1 && cell < 0; ++y)
                                    for (int x = 1; x < S - 1; ++x)
                                        if (cell_free(y * S + x) && r-- == 0) { cell = y * S + x; break; }
                            }
                        }
                    }
                    if (cell >= 0) set_food(s, cell);
                }
            } else if (nfood < 8 * K) {                               // random_rate, max_food :127
                for (int q = tid; q < C; q += nthr) {
                    const int y = fdiv(q, p.magic_S), x = q - y * S;
                    if (y < 1 || y > S - 2 || x < 1 || x > S - 2 || !cell_free(q)) continue;
                    const float u = p.replay ? p.u_rate[(size_t)e * C + q]
                                             : unit_float(draw_i(p.seed, ctr_now, (uint32_t)e, kStreamMultiFoodRate, (uint32_t)q));
                    if (u < p.food_rate) set_food(s, q);
                }
            }
        }

        // ---- per-agent outputs ----
        if (warp == 0) {
            if (valid) {
                const size_t n = (size_t)e * K + k;
                if (a_done && !a_done0) a_reward += p.reward_on_death;   // :683-685
                p.rewards[n] = a_reward;
                p.snake_col[n] = a_scol;
                p.edge_col[n] = a_ecol;
                p.food_cons[n] = a_foodc;
                p.sizes[n] = (float)a_size;
                p.dones[n] = a_done;
                p.dones_out[n] = a_done;
                // (a snake with several head cells is not something a hint or a record can stand for: "unknown")
                if (p.head_hints) p.head_hints[n] = (short)(s.hcnt[k] > 1 ? -2 : a_done ? -1 : a_hp);
                p.boost_out[n] = a_boosted;
            }
            const unsigned alive = __ballot_sync(0xffffffffu, valid && !a_done);
            const unsigned scols = __ballot_sync(0xffffffffu, valid && a_scol);
            const unsigned ecols = __ballot_sync(0xffffffffu, valid && a_ecol);
            const int eaten = group_sum<32>(valid ? (int)a_foodc : 0, 0xffffffffu);
            if (lane == 0) {
                s.misc[4] = (int)alive;                               // for the fused reset below
                p.all_done[e] = alive == 0u;                          // :703
                if (p.stats) {
                    unsigned long long* slot = p.stats + (blockIdx.x % WURM_STATS_SLOTS) * WURM_STATS_FIELDS;
                    atomicAdd(slot + WURM_STAT_ENV_STEPS, 1ull);
                    if (alive == 0u) atomicAdd(slot + WURM_STAT_EPISODES, 1ull);
                    if (eaten) atomicAdd(slot + WURM_STAT_REWARD, (unsigned long long)eaten);
                    if (scols) atomicAdd(slot + WURM_STAT_SELF_COLLISIONS, (unsigned long long)__popc(scols));
                    if (ecols) atomicAdd(slot + WURM_STAT_EDGE_COLLISIONS, (unsigned long long)__popc(ecols));
                }
            }
        }
        __syncthreads();

        // ---- write the new state back: into the records, the reference's tensors, or both (whichever the caller keeps) ----
        const bool resync = !COMPACT || (SHADOW && s.misc[10] != 0);  // the records do not describe the loaded state: emit them all
        if (COMPACT || p.cells != nullptr) records_writeback(p, s, e, resync);
        if (!COMPACT || SHADOW) dense_writeback(p, s, e, valid, k, a_hp0, a_hp);
    }
    write_multi_obs(p, s, e);

    if (STEP && p.auto_reset) {
        // ---- fused reset (multi_snake.py:771-831), bit-identical to wurm_multi_reset called right after this
        // step: same draws (call counter + 1), same result.  The compact form already knows what the stand-alone
        // kernel has to scan the tensors for: which cells are occupied, which snakes are dead.
        __syncthreads();                                              // the observation is done with the records
        // (the candidate counts of a re-creation go where the records are: nothing reads those once the old food is cleared)
        const ResetScratch sc = carve_reset(reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(s.reset) + 15) & ~(uintptr_t)15), C,
                                            reinterpret_cast<unsigned char*>(s.cell));
        const uint64_t ctr = ctr_now + 1;
        const unsigned alive_mask = (unsigned)s.misc[4];              // snakes alive after this step (ballot of warp 0)
        const unsigned dead_mask = ~alive_mask & (K >= 32 ? 0xffffffffu : ((1u << K) - 1u));
        const int first_dead = dead_mask ? __ffs(dead_mask) - 1 : -1;
        const bool all_dead = alive_mask == 0u;
        const bool rec_out = COMPACT || p.cells != nullptr;
        constexpr bool dense_out = !COMPACT || SHADOW;
        if (all_dead) {                                               // :787-798 re-create the env
            // all snakes are dead, so their tensors are already zero (the step deleted them); only food is left
            for (int q = tid; q < C; q += nthr)
                if (s.cell[q] & kFood) {
                    if (rec_out) p.cells[(size_t)e * p.Cp + q] = 0u;
                    if (dense_out) p.foods[(size_t)e * C + q] = 0.0f;
                }
            __syncthreads();                                          // the records are dead from here on
            const int fcell = decide_recreate(p, e, ctr, sc);
            write_recreated(p, e, sc, fcell, rec_out, dense_out);
        } else if (first_dead >= 0) {
            if (p.colour_random && tid < K && s.done[tid]) recolour(p, e, tid, ctr);     // :800-803
            if (p.respawn_any) {                                      // :805-829
                for (int q = tid; q < C; q += nthr) sc.occ[q] = (s.cell[q] & (kLive | kFood)) != 0u;
                __syncthreads();
                int d;
                const int cell = decide_respawn(p, e, ctr, sc, d);
                if (tid == 0) write_respawned(p, e, first_dead, cell, d, rec_out, dense_out);
            }
        }
    }
}

// MultiSnake.check_consistency (multi_snake.py:733-769) on the compact form: one CTA per env streams the
// state once; the per-snake verdicts come from the head index, body maximum / sum and the cell records.
__global__ void __launch_bounds__(256) multi_check_kernel(const MultiParams p, int* report) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const MultiSmem s = carve(smem_raw, p.C, p.K);
    const int C = p.C, K = p.K, e = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    for (int q = tid; q < C; q += nthr) s.cell[q] = 0u;
    if (tid < K) {
        s.hp[tid] = -1; s.size[tid] = 0; s.hcnt[tid] = 0; s.sum[tid] = 0;
        s.done[tid] = p.dones[(size_t)e * K + tid] != 0;
    }
    if (tid < 12) s.misc[tid] = 0;
    __syncthreads();
    MultiParams q = p;
    q.status = &s.misc[4];                      // overlap / multi-head of THIS env, not the env object's status word
    if (p.cells && p.cells_valid) load_env_compact<true>(q, s, e);
    else load_env<true>(q, s, e);
    __syncthreads();
    if (tid < K) {
        const int k = tid;
        int bits = 0;
        const int hp = s.hp[k], size = s.size[k], total = s.sum[k];
        if (s.done[k]) {                        // :766-769 dead snakes hold nothing
            if (hp >= 0 || total > 0) bits |= WURM_CHK_DEAD_NOT_ZERO;
        } else {                                // :741-742 snake_consistency of the living
            if (s.misc[5]) bits |= WURM_CHK_FOOD_VALUE;
            if (s.hcnt[k] != 1) bits |= WURM_CHK_HEAD_COUNT;
            if (total <= 0) bits |= WURM_CHK_NO_SNAKE;
            const uint32_t rec = hp >= 0 ? s.cell[hp] : 0u;
            if (!(hp >= 0 && rec_body(rec) && rec_owner(rec) == k && rec_value(rec) == size)) bits |= WURM_CHK_HEAD_NOT_AT_END;
            if ((sqrtf(8.0f * (float)total + 1.0f) - 1.0f) / 2.0f != (float)size) bits |= WURM_CHK_BODY_VALUES;
            if (total < 6) bits |= WURM_CHK_TOO_SHORT;
            if (hp >= 0 && (rec & kFood)) bits |= WURM_CHK_HEAD_ON_FOOD;
        }
        if (k == 0 && (s.misc[4] & WURM_ST_OVERLAP)) bits |= WURM_CHK_OVERLAP;               // :746-758
        if (bits) { atomicOr(report, bits); atomicAdd(report + 1, 1); atomicMin(report + 2, e); }
    }
}

// Conversions between the reference's fp32 tensors and the compact resident state (see load_env_compact), one CTA per
// env.  TO_COMPACT folds the tensors exactly as a dense step would and stores every record; a state the records cannot
// carry exactly (food / head values other than 1, non-integral or oversized body values, two bodies on one cell, a
// head that does not sit on its own body, two heads of one snake) raises WURM_ST_NOT_COMPACT.  !TO_COMPACT expands the
// records into the three tensors with dense 128-bit stores.
template <bool TO_COMPACT>
__global__ void __launch_bounds__(256) multi_convert_kernel(const MultiParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const MultiSmem s = carve(smem_raw, p.C, p.K);
    const int C = p.C, K = p.K, e = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    if (TO_COMPACT)
        for (int q = tid; q < C; q += nthr) s.cell[q] = 0u;
    if (tid < K) { s.hp[tid] = -1; s.size[tid] = 0; s.hcnt[tid] = 0; s.sum[tid] = 0; s.done[tid] = 0; }
    if (tid < 12) s.misc[tid] = 0;
    __syncthreads();
    if (TO_COMPACT) {
        load_env<false>(p, s, e);                                     // no hints: the heads tensor is scanned
        __syncthreads();
        uint32_t* gc = p.cells + (size_t)e * p.Cp;
        for (int q = tid; q < p.Cp; q += nthr) gc[q] = q < C ? hbm_record(s.cell[q]) : 0u;
        if (tid < K) {
            const int hp = s.hp[tid], hc = s.hcnt[tid];
            bool bad = hc > 1;
            if (hc == 1) {
                const uint32_t rec = s.cell[hp];
                bad = !(rec_body(rec) && rec_owner(rec) == tid);      // a head must sit on its own body to be a record
            }
            p.head_hints[(size_t)e * K + tid] = (short)((hc == 1 && !bad) ? hp : -1);
            if (bad) atomicOr(p.status, WURM_ST_NOT_COMPACT);
        }
        if (tid == 0 && s.misc[3]) atomicOr(p.status, WURM_ST_NOT_COMPACT);
    } else {
        load_env_compact<false>(p, s, e);
        __syncthreads();
        store_floats(p.foods + (size_t)e * C, C, [&](int i) { return (s.cell[i] & kFood) ? 1.0f : 0.0f; });
        store_floats(p.heads + (size_t)e * K * C, K * C, [&](int i) {
            const int kk = fdiv_C(i, C, p.magic_C);
            return s.hp[kk] == i - kk * C ? 1.0f : 0.0f;
        });
        store_floats(p.bodies + (size_t)e * K * C, K * C, [&](int i) {
            const int kk = fdiv_C(i, C, p.magic_C);
            const uint32_t rec = s.cell[i - kk * C];
            return (rec_body(rec) && rec_owner(rec) == kk) ? (float)rec_value(rec) : 0.0f;
        });
    }
}

// ---------------------------------------------------------------------------------------------
// reset (multi_snake.py:771-831)
// ---------------------------------------------------------------------------------------------
// Calls f(q, record) for every non-zero record of one env's row, whole CTA: 128-bit loads, four in flight per thread before
// the first is looked at (one dependent 4-byte load per cell was 55 % of the stand-alone reset's stall samples).
template <typename F>
__device__ __forceinline__ void walk_records(const uint32_t* row, int Cp, F&& f) {
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
    const int nvec = Cp >> 2, tid = threadIdx.x, nthr = blockDim.x;
    for (int j0 = tid; j0 < nvec; j0 += 4 * nthr) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * nthr;
            v[u] = j < nvec ? r4[j] : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if ((v[u].x | v[u].y | v[u].z | v[u].w) == 0u) continue;
            const int q = 4 * (j0 + u * nthr);
            if (v[u].x) f(q, v[u].x);
            if (v[u].y) f(q + 1, v[u].y);
            if (v[u].z) f(q + 2, v[u].z);
            if (v[u].w) f(q + 3, v[u].w);
        }
    }
}

// The CTA-wide part of the stand-alone reset of ONE env (every thread calls it with the same e).
__device__ __forceinline__ void reset_one_env(const MultiParams& p, unsigned char* smem_raw, int e) {
    const int C = p.C, K = p.K, tid = threadIdx.x, nthr = blockDim.x;
    const ResetScratch sc = carve_reset(smem_raw, C, smem_raw + ((reset_scratch_size(C) + 15) & ~15));
    const uint64_t ctr = call_counter(p);

    const bool recreate = p.env_done[e] != 0;
    int first_dead = -1, ndead = 0;
    for (int k = K - 1; k >= 0; --k)
        if (p.dones[(size_t)e * K + k]) { first_dead = k; ++ndead; }
    if (!recreate && (ndead == 0 || !p.respawn_any)) return;          // nothing to do for this env

    if (p.cells && p.cells_valid) {
        // the records are the state, or a valid shadow of the tensors: decide on the records, write to whichever forms exist
        uint32_t* gc = p.cells + (size_t)e * p.Cp;
        const bool dense = p.foods != nullptr;
        if (recreate) {
            const int fcell = decide_recreate(p, e, ctr, sc);
            walk_records(gc, p.Cp, [&](int q, uint32_t rec) {
                gc[q] = 0u;
                if (dense) {
                    if (rec & kFood) p.foods[(size_t)e * C + q] = 0.0f;
                    if (rec_body(rec)) p.bodies[((size_t)e * K + rec_owner(rec)) * C + q] = 0.0f;     // (none on a consistent state)
                }
            });
            if (dense && tid < K) {
                const int h = p.head_hints[(size_t)e * K + tid];
                if (h >= 0 && h < C) p.heads[((size_t)e * K + tid) * C + h] = 0.0f;                   // (none on a consistent state)
            }
            __syncthreads();
            write_recreated(p, e, sc, fcell, p.cells != nullptr, p.foods != nullptr);
            return;
        }
        for (int q = tid; q < C; q += nthr) sc.occ[q] = 0;
        __syncthreads();
        walk_records(gc, p.Cp, [&](int q, uint32_t rec) {
            sc.occ[q] = 1;                                            // (leftovers count as occupied, as in the dense path)
            if (rec_body(rec) && rec_owner(rec) == first_dead) {      // leftovers of the dead snake
                gc[q] = rec & ~kLive;
                if (dense) p.bodies[((size_t)e * K + first_dead) * C + q] = 0.0f;
            }
        });
        __syncthreads();
        int d;
        const int cell = decide_respawn(p, e, ctr, sc, d);
        if (tid == 0) write_respawned(p, e, first_dead, cell, d, p.cells != nullptr, p.foods != nullptr);
        return;
    }
    float* gfood = p.foods + (size_t)e * C;
    float* ghead = p.heads + (size_t)e * K * C;
    float* gbody = p.bodies + (size_t)e * K * C;
    if (recreate) {                                                   // :787-798
        const int fcell = decide_recreate(p, e, ctr, sc);
        // An env is re-created when all its snakes are dead, i.e. (on a consistent state) its head and body grids
        // are already all-zero: instead of storing (1+2K)*S*S values, the old tensors are scanned (cheap 128-bit
        // loads) and only non-zero leftovers are cleared, then the 4K+1 cells of the new env are stored.
        scan_nonzero(gfood, C, [&](int i, float) { gfood[i] = 0.0f; });
        scan_nonzero(ghead, K * C, [&](int i, float) { ghead[i] = 0.0f; });
        scan_nonzero(gbody, K * C, [&](int i, float) { gbody[i] = 0.0f; });
        __syncthreads();
        write_recreated(p, e, sc, fcell, p.cells != nullptr, p.foods != nullptr);
        return;                                                       // every agent alive: no respawn
    }

    // :805-829 respawn the first dead snake of the env where there is room
    for (int q = tid; q < C; q += nthr) sc.occ[q] = 0;
    __syncthreads();
    scan_nonzero(gfood, C, [&](int i, float) { sc.occ[i] = 1; });
    scan_nonzero(ghead, K * C, [&](int i, float) {
        const int kk = fdiv_C(i, C, p.magic_C);
        sc.occ[i - kk * C] = 1;
        if (kk == first_dead) ghead[i] = 0.0f;                        // leftovers of the dead snake (none on a consistent state)
    });
    scan_nonzero(gbody, K * C, [&](int i, float) {
        const int kk = fdiv_C(i, C, p.magic_C);
        sc.occ[i - kk * C] = 1;
        if (kk == first_dead) gbody[i] = 0.0f;
    });
    __syncthreads();
    int d;
    const int cell = decide_respawn(p, e, ctr, sc, d);
    if (tid == 0) write_respawned(p, e, first_dead, cell, d, p.cells != nullptr, p.foods != nullptr);
}

// One CTA looks at 32 consecutive envs: lane i of warp 0 reads env i's flags (coalesced), a ballot gives
// the envs with work to do, and the CTA handles those one after another.  In the common case (nothing
// died) an env costs K+1 bytes of traffic and a share of one warp instruction, not a CTA launch.
// Envs per CTA.  Envs that need CTA-wide work (re-creation, respawn) are handled one after the other, so the kernel lasts as
// long as its unluckiest CTA; 8 or 16 envs per CTA measured a few per cent better than 4 or 32 (scripts/time_reset.py).
#ifndef WURM_RESET_ENVS
#define WURM_RESET_ENVS 8
#endif
constexpr int kResetEnvs = WURM_RESET_ENVS;
__global__ void __launch_bounds__(256, 4) multi_reset_kernel(const MultiParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned recreate_s, dead_s;
    const int e0 = blockIdx.x * kResetEnvs, K = p.K;
    if (threadIdx.x < 32) {
        const int e = e0 + (int)threadIdx.x;
        const unsigned rec = __ballot_sync(0xffffffffu, (int)threadIdx.x < kResetEnvs && e < p.E && p.env_done[e] != 0);
        if (threadIdx.x == 0) { recreate_s = rec; dead_s = 0u; }
    }
    __syncthreads();
    // one (env, snake) pair per thread and iteration -- the 32 * K done flags of the CTA's envs are one contiguous span -- instead
    // of one lane per env walking its K flags (and recolouring its dead snakes) one dependent load after the other
    const unsigned recreate = recreate_s;
    const int pairs = min(kResetEnvs, p.E - e0) * K;
    for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
        const int el = i / K, k = i - el * K;
        if ((recreate >> el) & 1u) continue;                          // a re-created env has every snake alive again
        if (!p.dones[(size_t)e0 * K + i]) continue;
        atomicOr(&dead_s, 1u << el);
        if (p.colour_random) recolour(p, e0 + el, k, call_counter(p));    // :800-803 a new colour for every dead snake
    }
    __syncthreads();
    unsigned todo = recreate | (p.respawn_any ? dead_s : 0u);         // CTA-wide work: re-creation or a respawn
    while (todo) {
        const int e = e0 + __ffs(todo) - 1;
        todo &= todo - 1;
        __syncthreads();                                              // the previous env is done with shared memory
        reset_one_env(p, smem_raw, e);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int plan_multi(const WurmMultiCfg* cfg, const WurmMultiState* st, MultiParams* p) {
    if (!cfg || !st) return fail(WURM_E_INVALID, "cfg or state is NULL");
    if (cfg->num_envs <= 0) return fail(WURM_E_INVALID, "num_envs must be positive");
    if (cfg->num_snakes <= 0 || cfg->num_snakes > kMaxK) return fail(WURM_E_UNSUPPORTED, "num_snakes must be in [1, 32]");
    if (cfg->size < 7) return fail(WURM_E_INVALID, "size must be >= 7 (a length-3 snake two cells from the wall)");
    if (cfg->size > 181) return fail(WURM_E_UNSUPPORTED, "size > 181: body values no longer fit the 16-bit cell record");
    if (cfg->obs_mode < WURM_MOBS_NONE || cfg->obs_mode > WURM_MOBS_PARTIAL) return fail(WURM_E_INVALID, "bad obs_mode");
    if (cfg->obs_mode == WURM_MOBS_PARTIAL && (cfg->obs_n < 0 || cfg->obs_n > 127)) return fail(WURM_E_INVALID, "bad obs_n");
    if (cfg->food_mode != 0 && cfg->food_mode != 1) return fail(WURM_E_INVALID, "bad food_mode");
    if (!st->dones || !st->orientations || !st->boost_this_step || !st->agent_colours) return fail(WURM_E_INVALID, "NULL state pointer");
    if (st->cells) {
        if (!st->head_hints) return fail(WURM_E_INVALID, "compact state needs head_hints (the snakes' head cells)");
        if (reinterpret_cast<uintptr_t>(st->cells) & 15u) return fail(WURM_E_INVALID, "cells must be 16-byte aligned");
    } else if (!st->foods || !st->heads || !st->bodies) {
        return fail(WURM_E_INVALID, "NULL state pointer");
    }
    p->foods = st->foods; p->heads = st->heads; p->bodies = st->bodies; p->dones = st->dones;
    p->orientations = reinterpret_cast<long long*>(st->orientations); p->boost_this_step = st->boost_this_step;
    p->colours = st->agent_colours;
    p->head_hints = (cfg->size * cfg->size <= 32767) ? st->head_hints : nullptr;
    p->cells = st->cells; p->Cp = (cfg->size * cfg->size + 3) & ~3;
    p->cells_valid = st->cells ? (st->cells_valid != 0 || !st->foods) : 0;      // records without tensors are the state by definition
    p->E = cfg->num_envs; p->K = cfg->num_snakes; p->S = cfg->size; p->C = cfg->size * cfg->size;
    p->boost = cfg->boost; p->food_on_death = cfg->food_on_death; p->food_mode = cfg->food_mode;
    p->respawn_any = cfg->respawn_any; p->colour_random = cfg->colour_random;
    p->death_thr = cfg->death_threshold; p->boost_cost_prob = cfg->boost_cost_prob; p->food_rate = cfg->food_rate;
    p->reward_on_death = cfg->reward_on_death;
    p->obs_mode = cfg->obs_mode; p->obs_n = cfg->obs_n; p->W = 2 * cfg->obs_n + 1;
    p->magic_S = (uint32_t)((0x100000000ull + (uint64_t)p->S - 1) / (uint64_t)p->S);
    p->magic_C = (uint32_t)((0x100000000ull + (uint64_t)p->C - 1) / (uint64_t)p->C);
    p->magic_W = (uint32_t)((0x100000000ull + (uint64_t)p->W - 1) / (uint64_t)p->W);
    return WURM_OK;
}

template <bool STEP, int THREADS, bool COMPACT = false, bool SHADOW = false>
static int launch_multi_env_t(const MultiParams& p, cudaStream_t stream) {
    auto kern = multi_env_kernel<STEP, THREADS, COMPACT, SHADOW>;
    size_t smem = multi_smem_bytes(p.C, p.K, COMPACT);
    if (const char* v = getenv("WURM_MULTI_SMEM_PAD")) smem += (size_t)atoi(v);     // occupancy experiments
    if (smem > 227 * 1024) return fail(WURM_E_UNSUPPORTED, "env does not fit shared memory");
    static SmemOptIn opt_in;                           // per instantiation, per device inside
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), &opt_in, (int)smem, false, "cudaFuncSetAttribute(multi_env_kernel)")) return rc;
    kern<<<p.E, THREADS, smem, stream>>>(p);
    return check_launch("multi_env_kernel");
}

template <bool STEP>
static int launch_multi_env(const MultiParams& p, cudaStream_t stream) {
    // ~24 cells per thread at most (profiles/r01_sweep_multi.txt): 32 threads up to S=31, 64 to S=39, 128 to S=55, 256 above
    int threads = 32;
    while (threads < 256 && p.C > 24 * threads) threads <<= 1;
    if (const char* v = getenv("WURM_MULTI_THREADS")) threads = atoi(v);       // tuning override: 32, 64, 128, 256
    if (p.cells && p.cells_valid) {
        // the records are the state (or a valid shadow of it): nothing to stream, so the CTA is sized by the per-cell passes alone
        threads = 32;
        while (threads < 256 && p.C > 48 * threads) threads <<= 1;
        if (const char* v = getenv("WURM_MULTI_COMPACT_THREADS")) threads = atoi(v);
        if (p.foods) {                                                 // shadowed dense state
            switch (threads) {
                case 32: return launch_multi_env_t<STEP, 32, true, true>(p, stream);
                case 64: return launch_multi_env_t<STEP, 64, true, true>(p, stream);
                case 128: return launch_multi_env_t<STEP, 128, true, true>(p, stream);
                default: return launch_multi_env_t<STEP, 256, true, true>(p, stream);
            }
        }
        switch (threads) {
            case 32: return launch_multi_env_t<STEP, 32, true>(p, stream);
            case 64: return launch_multi_env_t<STEP, 64, true>(p, stream);
            case 128: return launch_multi_env_t<STEP, 128, true>(p, stream);
            default: return launch_multi_env_t<STEP, 256, true>(p, stream);
        }
    }
    switch (threads) {
        case 32: return launch_multi_env_t<STEP, 32>(p, stream);
        case 64: return launch_multi_env_t<STEP, 64>(p, stream);
        case 128: return launch_multi_env_t<STEP, 128>(p, stream);
        default: return launch_multi_env_t<STEP, 256>(p, stream);
    }
}

}  // namespace wurm

using namespace wurm;

extern "C" int64_t wurm_multi_obs_elems(const WurmMultiCfg* cfg) {
    if (!cfg) return -1;
    const int64_t C = (int64_t)cfg->size * cfg->size, W = 2 * cfg->obs_n + 1;
    return cfg->obs_mode == WURM_MOBS_FULL ? 3 * C : cfg->obs_mode == WURM_MOBS_PARTIAL ? 3 * W * W : 0;
}

static int multi_step_impl(const WurmMultiCfg* cfg, const WurmMultiState* state, const void* const* actions, int action_bytes,
                           const WurmMultiStepDraws* draws, int auto_reset, const WurmMultiResetDraws* reset_draws, uint64_t seed,
                           uint64_t step, const uint64_t* step_dev, const WurmMultiStepOut* out, int32_t* status, int64_t* stats,
                           void* stream) {
    MultiParams p = {};
    if (int rc = plan_multi(cfg, state, &p)) return rc;
    if (!actions || !out || !status) return fail(WURM_E_INVALID, "NULL pointer");
    if (!valid_action_bytes(action_bytes)) return fail(WURM_E_INVALID, "action_bytes must be 1, 2, 4 or 8");
    for (int k = 0; k < p.K; ++k) {
        if (!actions[k]) return fail(WURM_E_INVALID, "NULL action pointer");
        p.actions[k] = actions[k];
    }
    if (!out->rewards || !out->snake_collision || !out->edge_collision || !out->food || !out->size || !out->all_done ||
        !out->dones || !out->boost)
        return fail(WURM_E_INVALID, "NULL output pointer");
    if (cfg->obs_mode != WURM_MOBS_NONE && !out->obs) return fail(WURM_E_INVALID, "obs is NULL");
    p.action_bytes = action_bytes;
    if (draws) {
        p.replay = 1; p.boost_phase_ran = draws->boost_phase_ran;
        p.u_boost = draws->u_boost; p.u_cost = draws->u_cost; p.u_reg = draws->u_reg;
        p.food_cell = draws->food_cell; p.u_rate = draws->u_rate;
        if (p.boost_phase_ran && !p.u_cost) return fail(WURM_E_INVALID, "replay: boost phase ran but u_cost is NULL");
        if (p.food_on_death && !p.u_reg) return fail(WURM_E_INVALID, "replay: u_reg is NULL");
        if (p.food_on_death && p.boost_phase_ran && !p.u_boost) return fail(WURM_E_INVALID, "replay: u_boost is NULL");
        if (p.food_mode == 0 && !p.food_cell) return fail(WURM_E_INVALID, "replay: food_cell is NULL");
        if (p.food_mode == 1 && !p.u_rate) return fail(WURM_E_INVALID, "replay: u_rate is NULL");
    }
    p.seed = seed; p.step = step; p.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
    p.rewards = out->rewards; p.snake_col = out->snake_collision; p.edge_col = out->edge_collision;
    p.food_cons = out->food; p.sizes = out->size; p.all_done = out->all_done; p.obs = out->obs;
    p.dones_out = out->dones; p.boost_out = out->boost;
    p.status = status; p.stats = reinterpret_cast<unsigned long long*>(stats);
    p.auto_reset = auto_reset;
    if (auto_reset && reset_draws) {
        p.create = reset_draws->create; p.respawn = reset_draws->respawn; p.colours_replay = reset_draws->colours;
        if (!p.create || (p.respawn_any && !p.respawn) || (p.colour_random && !p.colours_replay))
            return fail(WURM_E_INVALID, "replay: NULL reset draw array");
    }
    return launch_multi_env<true>(p, (cudaStream_t)stream);
}

extern "C" int wurm_multi_step(const WurmMultiCfg* cfg, const WurmMultiState* state, const void* const* actions,
                               int action_bytes, const WurmMultiStepDraws* draws, uint64_t seed, uint64_t step,
                               const uint64_t* step_dev, const WurmMultiStepOut* out, int32_t* status, int64_t* stats,
                               void* stream) {
    return multi_step_impl(cfg, state, actions, action_bytes, draws, 0, nullptr, seed, step, step_dev, out, status, stats, stream);
}

extern "C" int wurm_multi_step_reset(const WurmMultiCfg* cfg, const WurmMultiState* state, const void* const* actions,
                                     int action_bytes, const WurmMultiStepDraws* draws, const WurmMultiResetDraws* reset_draws,
                                     uint64_t seed, uint64_t step, const uint64_t* step_dev, const WurmMultiStepOut* out,
                                     int32_t* status, int64_t* stats, void* stream) {
    return multi_step_impl(cfg, state, actions, action_bytes, draws, 1, reset_draws, seed, step, step_dev, out, status, stats,
                           stream);
}

extern "C" int wurm_multi_observe(const WurmMultiCfg* cfg, const WurmMultiState* state, float* obs, int32_t* status,
                                  void* stream) {
    MultiParams p = {};
    if (int rc = plan_multi(cfg, state, &p)) return rc;
    if (!obs || !status) return fail(WURM_E_INVALID, "NULL pointer");
    if (cfg->obs_mode == WURM_MOBS_NONE) return WURM_OK;
    p.obs = obs; p.status = status;
    return launch_multi_env<false>(p, (cudaStream_t)stream);
}

extern "C" int wurm_multi_env_images(const WurmMultiCfg* cfg, const WurmMultiState* state, int16_t* img, int32_t* status,
                                     void* stream) {
    MultiParams p = {};
    if (int rc = plan_multi(cfg, state, &p)) return rc;
    if (!img || !status) return fail(WURM_E_INVALID, "NULL pointer");
    p.obs_mode = WURM_MOBS_NONE; p.img = img; p.status = status;
    return launch_multi_env<false>(p, (cudaStream_t)stream);
}

extern "C" int wurm_multi_reset(const WurmMultiCfg* cfg, const WurmMultiState* state, const uint8_t* env_done,
                                const WurmMultiResetDraws* draws, uint64_t seed, uint64_t step, const uint64_t* step_dev,
                                int32_t* status, void* stream) {
    MultiParams p = {};
    if (int rc = plan_multi(cfg, state, &p)) return rc;
    if (!env_done || !status) return fail(WURM_E_INVALID, "NULL pointer");
    p.env_done = env_done; p.seed = seed; p.step = step; p.status = status;
    p.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
    if (draws) {
        p.replay = 1; p.create = draws->create; p.respawn = draws->respawn; p.colours_replay = draws->colours;
        if (!p.create || (p.respawn_any && !p.respawn) || (p.colour_random && !p.colours_replay))
            return fail(WURM_E_INVALID, "replay: NULL draw array");
    }
    // (64 registers: twice as many of these smaller CTAs are resident, and most of them only read their envs' flags;
    // measured against 128 / 256 threads: 0.040 against 0.045 ms at K=4, S=25 and 0.050 against 0.052 ms at K=16, S=64)
    const int threads = p.C <= 1024 ? 64 : 128;
    const size_t smem = ((reset_scratch_bytes(p.C) + 15) & ~(size_t)15) + reset_counts_size(p.C) + 16;
    multi_reset_kernel<<<(p.E + kResetEnvs - 1) / kResetEnvs, threads, smem, (cudaStream_t)stream>>>(p);
    return check_launch("multi_reset_kernel");
}

static int multi_convert(const WurmMultiCfg* cfg, const WurmMultiState* state, int32_t* status, void* stream, bool to_compact) {
    MultiParams p = {};
    if (int rc = plan_multi(cfg, state, &p)) return rc;
    if (!state->cells || !state->foods || !state->heads || !state->bodies) return fail(WURM_E_INVALID, "conversion needs both forms of the state");
    if (to_compact && !status) return fail(WURM_E_INVALID, "NULL pointer");
    p.status = status;
    const size_t smem = multi_smem_bytes(p.C, p.K);
    if (smem > 227 * 1024) return fail(WURM_E_UNSUPPORTED, "env does not fit shared memory");
    static SmemOptIn opt_in[2];
    const void* kern = to_compact ? reinterpret_cast<const void*>(multi_convert_kernel<true>) : reinterpret_cast<const void*>(multi_convert_kernel<false>);
    if (int rc = ensure_dynamic_smem(kern, &opt_in[to_compact ? 1 : 0], (int)smem, false, "cudaFuncSetAttribute(multi_convert_kernel)")) return rc;
    if (to_compact) multi_convert_kernel<true><<<p.E, 256, smem, (cudaStream_t)stream>>>(p);
    else multi_convert_kernel<false><<<p.E, 256, smem, (cudaStream_t)stream>>>(p);
    return check_launch("multi_convert_kernel");
}

extern "C" int wurm_multi_compact(const WurmMultiCfg* cfg, const WurmMultiState* state, int32_t* status, void* stream) {
    return multi_convert(cfg, state, status, stream, true);
}

extern "C" int wurm_multi_expand(const WurmMultiCfg* cfg, const WurmMultiState* state, void* stream) {
    return multi_convert(cfg, state, nullptr, stream, false);
}

extern "C" int wurm_multi_check(const WurmMultiCfg* cfg, const WurmMultiState* state, int32_t* report, void* stream) {
    MultiParams p = {};
    if (int rc = plan_multi(cfg, state, &p)) return rc;
    if (!report) return fail(WURM_E_INVALID, "NULL pointer");
    const size_t smem = multi_smem_bytes(p.C, p.K);
    static SmemOptIn opt_in;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(multi_check_kernel), &opt_in, (int)smem, false, "cudaFuncSetAttribute(multi_check_kernel)")) return rc;
    multi_check_kernel<<<p.E, p.C <= 1024 ? 128 : 256, smem, (cudaStream_t)stream>>>(p, report);
    return check_launch("multi_check_kernel");
}
