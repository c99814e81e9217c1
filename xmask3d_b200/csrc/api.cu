// libxm3d — error reporting, version and device queries of the C ABI (include/xm3d.h).
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace xm3d {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

thread_local cudaEvent_t g_pool_ev[2] = {nullptr, nullptr};

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return XM3D_OK;
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return XM3D_ERR_CUDA;
}

static int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}

int sm_count() {
    static std::atomic<int> cached[64];
    const int dev = current_device();
    if (dev >= 0 && dev < 64) {
        const int c = cached[dev].load(std::memory_order_relaxed);
        if (c) return c;
    }
    int n = 0;
    if (dev < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;          // B200; used only to size workspaces when no device is visible
    }
    if (dev < 64) cached[dev].store(n, std::memory_order_relaxed);
    return n;
}

bool first_use_on_device(std::atomic<uint64_t> *flag) {
    const int dev = current_device();
    if (dev < 0 || dev >= 64) return true;                     // unknown device: redo the (idempotent) setup
    const uint64_t bit = 1ull << dev;
    return (flag->fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

}  // namespace xm3d

extern "C" int xm3d_version(void) { return XM3D_VERSION; }

extern "C" const char *xm3d_last_error(void) { return xm3d::g_err; }

extern "C" int64_t xm3d_launch_count(void) { return xm3d::g_launches.load(std::memory_order_relaxed); }

extern "C" int xm3d_device_info(int32_t *sm, int32_t *cc_major, int32_t *cc_minor) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        cudaGetLastError();
        xm3d::set_error("xm3d_device_info: no CUDA device");
        return XM3D_ERR_CUDA;
    }
    if (sm) *sm = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return XM3D_OK;
}

extern "C" void xm3d_set_pool_events(void *before, void *after) {
    xm3d::g_pool_ev[0] = static_cast<cudaEvent_t>(before);
    xm3d::g_pool_ev[1] = static_cast<cudaEvent_t>(after);
}
