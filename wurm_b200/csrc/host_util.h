// Host-side error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

namespace wurm {
// Records `msg` as the calling thread's last error and returns `code`.
int fail(int code, const char* msg);
int fail_cuda(cudaError_t err, const char* where);
// cudaGetLastError() after a launch -> WURM_OK / WURM_E_CUDA.
int check_launch(const char* kernel);

// Dynamic shared-memory opt-in of one kernel (cudaFuncAttributeMaxDynamicSharedMemorySize).  The attribute is PER
// DEVICE, so the largest size configured so far is remembered per device ordinal: a process that steps an env on
// cuda:0 and then another on cuda:1 configures both.  One zero-initialised static instance per kernel instantiation;
// concurrent callers at worst set the attribute twice.
struct SmemOptIn {
    int configured[64];
};
// Makes sure `kernel` may be launched with `bytes` of dynamic shared memory on the current device (`always`: also
// below the 48 KB default, for kernels that want the max-shared carve-out preference).  WURM_OK or WURM_E_CUDA.
int ensure_dynamic_smem(const void* kernel, SmemOptIn* cache, int bytes, bool prefer_shared_carveout, const char* name);
}  // namespace wurm
