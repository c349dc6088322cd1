// Shared device helpers for the wurm_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wurm {

constexpr float kEps = 1e-6f;  // reference config.py:11

// orientation k <=> head = neck + OFF[k]; action a moves the head by -OFF[a]
// (reference wurm/_filters.py:7-28 read as cross-correlation taps)
__device__ __forceinline__ int off_y(int k) { return (k == 0) ? -1 : (k == 2) ? 1 : 0; }
__device__ __forceinline__ int off_x(int k) { return (k == 1) ? 1 : (k == 3) ? -1 : 0; }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 and the draw streams (must agree with oracle/wurm_oracle.c, which restates them
// independently).  counter = (unit, stream, step_lo, step_hi), key = (seed_lo, seed_hi).
// ---------------------------------------------------------------------------------------------
// The i-th 32-bit draw of stream s for env e at call counter `step` is word (i % 4) of
// philox(counter = (e, s | (i / 4) << 4, step_lo, step_hi), key = (seed_lo, seed_hi)).
enum DrawStream : uint32_t {
    kStreamSingleStepFood = 0, kStreamSingleReset = 1,
    kStreamMultiDeathBoost = 2, kStreamMultiBoostCost = 3, kStreamMultiDeathRegular = 4,
    kStreamMultiFoodOne = 5, kStreamMultiFoodRate = 6, kStreamMultiCreateSnake = 7,
    kStreamMultiCreateFood = 8, kStreamMultiRespawn = 9, kStreamMultiColour = 10
};
constexpr uint32_t kRejectionTries = 32;

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 draw(uint64_t seed, uint64_t step, uint32_t unit, uint32_t stream) {
    return philox4x32_10(unit, stream, (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
}

__device__ __forceinline__ uint32_t draw_i(uint64_t seed, uint64_t step, uint32_t unit, uint32_t stream, uint32_t i) {
    const uint4 r = draw(seed, step, unit, stream | ((i >> 2) << 4));
    const uint32_t w = i & 3u;
    return w == 0 ? r.x : w == 1 ? r.y : w == 2 ? r.z : r.w;
}

// uniform float in [0,1) with 24 random bits (the resolution of torch.rand)
__device__ __forceinline__ float unit_float(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// uniform integer in [0, n) by multiply-shift
__device__ __forceinline__ uint32_t bounded(uint32_t r, uint32_t n) { return __umulhi(r, n); }

// ---------------------------------------------------------------------------------------------
// Action vectors: int64 / int32 / int16 as in the reference (single_snake.py:198-200), plus uint8 (action_bytes 1), the
// width a host-side policy uploads when PCIe bytes matter (8 -> 1 byte per env-step).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long load_action(const void* actions, int action_bytes, size_t e) {
    if (action_bytes == 8) return ((const long long*)actions)[e];
    if (action_bytes == 4) return ((const int*)actions)[e];
    if (action_bytes == 2) return ((const short*)actions)[e];
    return ((const unsigned char*)actions)[e];
}
__device__ __forceinline__ void store_action(void* actions, int action_bytes, size_t e, long long a) {
    if (action_bytes == 8) ((long long*)actions)[e] = a;
    else if (action_bytes == 4) ((int*)actions)[e] = (int)a;
    else if (action_bytes == 2) ((short*)actions)[e] = (short)a;
    else ((unsigned char*)actions)[e] = (unsigned char)a;
}
__host__ __device__ __forceinline__ bool valid_action_bytes(int b) { return b == 1 || b == 2 || b == 4 || b == 8; }

// One byte per env-step carrying everything a host-side consumer needs back (WURM_PACKED_* in the header):
// bit 0 done, bit 1 self collision, bit 2 edge collision, bits 3-4 the reward as a small integer (0..3).
__device__ __forceinline__ unsigned char pack_result(bool done, bool self_col, bool edge_col, float reward) {
    const int r = reward <= 0.0f ? 0 : reward >= 3.0f ? 3 : (int)reward;
    return (unsigned char)((done ? 1 : 0) | (self_col ? 2 : 0) | (edge_col ? 4 : 0) | (r << 3));
}

// ---------------------------------------------------------------------------------------------
// Bulk asynchronous copies (TMA 1-D, SASS UBLKCP) between global memory and a CTA's shared tile,
// completion tracked by an mbarrier (loads) or a bulk async-group (stores).
// Addresses and byte counts must be multiples of 16.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Sub-warp groups: G consecutive lanes (G a power of two <= 32) cooperate on one environment.
// ---------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ unsigned group_mask() {
    if constexpr (G == 32) {
        return 0xffffffffu;
    } else {
        const unsigned lane = threadIdx.x & 31u;
        return ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));
    }
}
template <int G>
__device__ __forceinline__ float group_max(float v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(m, v, o));
    return v;
}
template <int G>
__device__ __forceinline__ int group_max(int v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(m, v, o));
    return v;
}
template <int G>
__device__ __forceinline__ int group_sum(int v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}

}  // namespace wurm
