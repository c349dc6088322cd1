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

}  // namespace wurm

extern "C" int wurm_abi_version(void) { return WURM_ABI_VERSION; }
extern "C" const char* wurm_last_error(void) { return wurm::g_last_error; }
