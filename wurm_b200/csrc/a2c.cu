// A2C return scan (reference wurm/rl/a2c.py:49-63) for B200 (sm_100a): the reference walks the trajectory
// backwards in a Python loop of ~6 tensor ops per time step over (T,N) tensors; here one thread per
// environment carries the recurrence in a register, reading (t, n) elements that are coalesced across the warp.
// fp32 in the reference's operation order, so the result is bit-identical: the reference's Python scalars are
// doubles that meet fp32 tensors one at a time, so `gamma` enters as float(gamma) and `gamma * gae_lambda` (:56) as the
// fp32 rounding of the DOUBLE product -- the ABI therefore takes both as doubles.
#include "../../include/wurm_b200.h"
#include "common.cuh"
#include "host_util.h"

namespace wurm {

__global__ void __launch_bounds__(256) a2c_returns_kernel(int T, int N, float gamma, float gamma_lambda, bool use_gae, const float* __restrict__ bootstrap,
                                                          const float* __restrict__ rewards, const float* __restrict__ values,
                                                          const uint8_t* __restrict__ dones, float* __restrict__ returns) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    if (!use_gae) {                                                   // n-step returns :58-61
        float R = bootstrap[n] * (dones[(size_t)(T - 1) * N + n] ? 0.0f : 1.0f);
        for (int t = T - 1; t >= 0; --t) {
            const float m = dones[(size_t)t * N + n] ? 0.0f : 1.0f;
            R = rewards[(size_t)t * N + n] + gamma * R * m;
            returns[(size_t)t * N + n] = R;
        }
    } else {                                                          // generalised advantage estimation :50-57
        float gae = 0.0f, next = bootstrap[n];
        for (int t = T - 1; t >= 0; --t) {
            const float m = dones[(size_t)t * N + n] ? 0.0f : 1.0f;
            const float v = values[(size_t)t * N + n];
            const float delta = rewards[(size_t)t * N + n] + gamma * next * m - v;
            gae = delta + gamma_lambda * m * gae;                     // :56 `self.gamma * self.gae_lambda` is a Python double product
            returns[(size_t)t * N + n] = gae + v;
            next = v;
        }
    }
}

}  // namespace wurm

using namespace wurm;

extern "C" int wurm_a2c_returns(int32_t num_steps, int32_t num_envs, double gamma, double gae_lambda, const float* bootstrap,
                                const float* rewards, const float* values, const uint8_t* dones, float* returns, void* stream) {
    if (num_steps <= 0 || num_envs <= 0) return fail(WURM_E_INVALID, "num_steps and num_envs must be positive");
    if (!bootstrap || !rewards || !dones || !returns || (gae_lambda >= 0.0 && !values)) return fail(WURM_E_INVALID, "NULL pointer");
    a2c_returns_kernel<<<(num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        num_steps, num_envs, (float)gamma, (float)(gamma * gae_lambda), gae_lambda >= 0.0, bootstrap, rewards, values, dones, returns);
    return check_launch("a2c_returns_kernel");
}
