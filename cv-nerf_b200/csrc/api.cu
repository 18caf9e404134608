// Library-level entry points: ABI version, last-error string, device query.
#include <stdarg.h>

#include "common.cuh"

namespace nerf {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

}  // namespace nerf

extern "C" int nerf_b200_abi_version(void) { return NERF_B200_ABI_VERSION; }

extern "C" const char* nerf_b200_last_error(void) { return nerf::g_last_error; }

extern "C" int nerf_b200_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}
