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

int sm_count() {
    static int cached = -1;
    if (cached >= 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached = n;
    return n;
}

}  // namespace pats

PATS_API int pats_version(void) { return PATS_B200_VERSION; }
PATS_API const char *pats_last_error(void) { return pats::g_err; }
PATS_API int pats_sm_count(void) { return pats::sm_count(); }
