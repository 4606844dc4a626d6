// Error reporting, launch counter and device queries behind the C ABI (include/ewvit.h).
#include "ewvit_common.cuh"

#include <atomic>
#include <mutex>

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void ewvit_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void ewvit_count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

int ewvit_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int ewvit_check_device() {
    static int ok[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        ewvit_set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
        return EWVIT_ERR_CUDA;
    }
    if (dev >= 0 && dev < 64 && ok[dev]) return EWVIT_OK;
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        ewvit_set_error("device %d is sm_%d%d; libewvit.so is built for sm_100a only (no fallback path)", dev,
                        major, minor);
        return EWVIT_ERR_NO_DEVICE;
    }
    if (dev >= 0 && dev < 64) ok[dev] = 1;
    return EWVIT_OK;
}

extern "C" {
int ewvit_abi_version(void) { return 1; }
const char *ewvit_last_error(void) { return g_err; }
uint64_t ewvit_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
}
