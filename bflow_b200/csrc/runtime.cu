// Error plumbing and trivial entry points of the C ABI (include/bflow_b200.h).
#include <string.h>
#include <stdio.h>
#include "common.cuh"

namespace bflow {
static thread_local char g_err[512] = "";
void set_error(const char* msg) {
    strncpy(g_err, msg, sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return BFLOW_ERR_CUDA;
    }
    return BFLOW_OK;
}
}  // namespace bflow

extern "C" int bflow_abi_version(void) { return 1; }
extern "C" const char* bflow_last_error(void) { return bflow::g_err; }
extern "C" int bflow_built_for_sm(void) { return 100; }

extern "C" int bflow_zero(void* ptr, unsigned long long bytes, void* stream) {
    BFLOW_REQUIRE(ptr != nullptr || bytes == 0, "bflow_zero: null pointer");
    if (bytes == 0) return BFLOW_OK;
    cudaError_t e = cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        bflow::set_error(cudaGetErrorString(e));
        return BFLOW_ERR_CUDA;
    }
    return BFLOW_OK;
}
