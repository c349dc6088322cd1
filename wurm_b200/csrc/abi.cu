// ABI version and the thread-local error string of the wurm_b200 C ABI (include/wurm_b200.h).
#include <stdio.h>

#include "../../include/wurm_b200.h"
#include "host_util.h"

namespace wurm {

static thread_local char g_last_error[256] = "";

int fail(int code, const char* msg) {
    snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
    return code;
}

int fail_cuda(cudaError_t err, const char* where) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", where, cudaGetErrorString(err));
    return WURM_E_CUDA;
}

int check_launch(const char* kernel) {
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? WURM_OK : fail_cuda(err, kernel);
}

int ensure_dynamic_smem(const void* kernel, SmemOptIn* cache, int bytes, bool prefer_shared_carveout, const char* name) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return fail_cuda(err, "cudaGetDevice");
    const bool cached = dev >= 0 && dev < 64;
    // configured[] holds bytes + 1 so that the zero-initialised state means "never configured on this device"
    if (cached && bytes + 1 <= cache->configured[dev]) return WURM_OK;
    if (bytes > 48 * 1024 || prefer_shared_carveout) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (err != cudaSuccess) return fail_cuda(err, name);
        if (prefer_shared_carveout)
            cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (cached) cache->configured[dev] = bytes + 1;
    return WURM_OK;
}

}  // namespace wurm

extern "C" int wurm_abi_version(void) { return WURM_ABI_VERSION; }
extern "C" const char* wurm_last_error(void) { return wurm::g_last_error; }
