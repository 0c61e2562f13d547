// C-ABI glue: error reporting, device queries.  See include/pats_b200.h.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace pats {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int invalid(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return PATS_E_INVALID;
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return PATS_E_CUDA;
}

int current_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        set_error("cudaGetDevice failed");
        return -1;
    }
    if (dev < 0 || dev >= kMaxDevices) {
        set_error("device ordinal %d beyond the %d devices the library keeps state for", dev, kMaxDevices);
        return -1;
    }
    return dev;
}

int sm_count() {
    static int cached[kMaxDevices];  // 0 = not asked yet
    const int dev = current_device();
    if (dev < 0) return 0;
    int n = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);
    if (n > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    __atomic_store_n(&cached[dev], n, __ATOMIC_RELAXED);
    return n;
}

}  // namespace pats

PATS_API int pats_version(void) { return PATS_B200_VERSION; }
PATS_API const char *pats_last_error(void) { return pats::g_err; }
PATS_API int pats_sm_count(void) { return pats::sm_count(); }
