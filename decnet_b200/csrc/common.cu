// decnet_b200/csrc/common.cu -- error reporting, launch accounting, device info.
#include <cstdlib>
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace decnet {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t err, const char *what) {
    if (err == cudaSuccess) return 0;
    set_error("%s failed: %s (%s)", what, cudaGetErrorName(err), cudaGetErrorString(err));
    return DECNET_ERR_CUDA_BASE + static_cast<int>(err);
}

int after_launch(const char *kernel_name) {
    ++g_launches;
    return cuda_status(cudaGetLastError(), kernel_name);
}

bool pdl_enabled() {
    static const bool on = [] { const char *e = std::getenv("DECNET_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

int sm_count_cached() {
    static std::mutex mu;
    static int cache[64];
    static bool have[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lk(mu);
    if (!have[dev]) {
        int n = 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
        cache[dev] = n;
        have[dev] = true;
    }
    return cache[dev];
}

}  // namespace decnet

extern "C" {

int decnet_abi_version(void) { return DECNET_ABI_VERSION; }

const char *decnet_last_error(void) { return decnet::g_err; }

int decnet_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    DECNET_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) { DECNET_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (cc_major) { DECNET_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { DECNET_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return DECNET_OK;
}

int64_t decnet_launch_count(void) { return decnet::g_launches; }
void decnet_reset_launch_count(void) { decnet::g_launches = 0; }

}  // extern "C"
