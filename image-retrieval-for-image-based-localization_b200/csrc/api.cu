// Error reporting, launch accounting and cached device attributes for libcir_b200.so.
#include "common.cuh"

namespace cir {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += n; }

const DeviceInfo& device_info() {
    static thread_local DeviceInfo info[16];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) dev = 0;
    DeviceInfo& d = info[dev];
    if (d.device != dev) {
        cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&d.coop, cudaDevAttrCooperativeLaunch, dev);
        d.device = dev;
    }
    return d;
}

}  // namespace cir

extern "C" {

const char* cir_last_error(void) { return cir::g_err; }

int cir_version(void) { return 100; }

int64_t cir_launch_count(int reset) {
    int64_t v = cir::g_launches;
    if (reset) cir::g_launches = 0;
    return v;
}

}  // extern "C"
