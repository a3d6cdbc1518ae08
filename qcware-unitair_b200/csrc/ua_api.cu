// Library-level plumbing of the C ABI: version, error string, launch counter.
#include <stdarg.h>
#include <string.h>

#include "ua_common.cuh"

namespace ua {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return UA_ERR_CUDA;
    }
    return UA_OK;
}

}  // namespace ua

extern "C" int ua_version(void) { return 100; }
extern "C" const char *ua_last_error(void) { return ua::g_err; }
extern "C" unsigned long long ua_launch_count(void) { return ua::g_launches.load(); }
