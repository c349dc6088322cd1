// Host-side error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

namespace wurm {
// Records `msg` as the calling thread's last error and returns `code`.
int fail(int code, const char* msg);
int fail_cuda(cudaError_t err, const char* where);
// cudaGetLastError() after a launch -> WURM_OK / WURM_E_CUDA.
int check_launch(const char* kernel);
}  // namespace wurm
