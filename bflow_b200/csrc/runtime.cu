// Error plumbing and trivial entry points of the C ABI (include/bflow_b200.h).
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
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
static unsigned long long* g_tl_buf = nullptr;
static int g_tl_cap = 0, g_tl_next = 0;
static char g_tl_names[4096][24];
unsigned long long* timeline_next_slot(const char* name) {
    if (g_tl_buf == nullptr || g_tl_next >= g_tl_cap || g_tl_next >= 4096) return nullptr;
    strncpy(g_tl_names[g_tl_next], name, 23);
    g_tl_names[g_tl_next][23] = 0;
    return g_tl_buf + 2 * (g_tl_next++);
}
int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    return dev;
}
int num_sms() {
    static int sms[64] = {};
    const int dev = current_device();
    if (dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0) sms[dev] = 148;
    }
    return sms[dev];
}
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("BFLOW_PDL");
        // on by default: measured on B200 inside the captured graph, 3.875 -> 3.762 ms per forward (the update block's launches fill at
        // most 114 of 148 SMs, so the next kernel's CTAs run their prologue on the idle ones and sit in griddepcontrol.wait).  Only the
        // generic tensor-core kernel and the lookup are launched this way: extending it to the slab / stem / head / InstanceNorm kernels
        // was measured worse (3.92 ms: dependents that start early then hold warps and registers on SMs the running kernel needs)
        on = (e != nullptr && strcmp(e, "0") == 0) ? 0 : 1;
    }
    return on == 1;
}
}  // namespace bflow

#ifndef BFLOW_SOURCE_HASH
#define BFLOW_SOURCE_HASH "unknown"
#endif
extern "C" int bflow_abi_version(void) { return BFLOW_ABI_VERSION; }
extern "C" int bflow_sizeof_conv_desc(void) { return (int)sizeof(bflow_conv_desc); }
extern "C" int bflow_sizeof_lookup_desc(void) { return (int)sizeof(bflow_lookup_desc); }
extern "C" int bflow_sizeof_lookup_otf_desc(void) { return (int)sizeof(bflow_lookup_otf_desc); }
extern "C" const char* bflow_source_hash(void) { return BFLOW_SOURCE_HASH; }
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

// development: device buffer of `capacity` {start, end} pairs (caller initialises start = ~0, end = 0); nullptr switches it off.
// Returns the number of slots handed out so far; bflow_timeline_name(i) names the launch that owns slot i.
extern "C" int bflow_timeline(void* buf, int capacity) {
    const int used = bflow::g_tl_next;
    bflow::g_tl_buf = reinterpret_cast<unsigned long long*>(buf);
    bflow::g_tl_cap = capacity;
    bflow::g_tl_next = 0;
    return used;
}
extern "C" int bflow_timeline_used(void) { return bflow::g_tl_next; }
extern "C" const char* bflow_timeline_name(int i) { return (i >= 0 && i < bflow::g_tl_next) ? bflow::g_tl_names[i] : ""; }
