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

// ------------------------------------------------------------------ peer memory (CUDA IPC)
#include <cuda.h>

extern "C" int ua_ipc_export(const void *ptr, unsigned char handle_out[64], long long *offset_out) {
    using namespace ua;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!ptr || !handle_out || !offset_out) { set_error("ua_ipc_export: null pointer"); return UA_ERR_INVALID; }
    // the handle names the whole allocation: report where `ptr` sits inside it
    void *base = nullptr;
    size_t size = 0;
    typedef CUresult (*RangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
    static RangeFn range_fn = nullptr;
    if (!range_fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            range_fn = reinterpret_cast<RangeFn>(p);
        else
            cudaGetLastError();
    }
    if (!range_fn || range_fn(reinterpret_cast<CUdeviceptr *>(&base), &size, (CUdeviceptr)(uintptr_t)ptr) != CUDA_SUCCESS) {
        set_error("ua_ipc_export: cannot find the allocation of %p", ptr);
        return UA_ERR_CUDA;
    }
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, base);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("ua_ipc_export: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return UA_ERR_CUDA;
    }
    memcpy(handle_out, &h, 64);
    *offset_out = (long long)((const char *)ptr - (const char *)base);
    return UA_OK;
}

extern "C" int ua_ipc_open(const unsigned char handle[64], long long offset, void **ptr_out) {
    using namespace ua;
    if (!handle || !ptr_out || offset < 0) { set_error("ua_ipc_open: bad argument"); return UA_ERR_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *base = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("ua_ipc_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
        return UA_ERR_CUDA;
    }
    *ptr_out = (char *)base + offset;
    return UA_OK;
}

extern "C" int ua_ipc_close(void *ptr, long long offset) {
    using namespace ua;
    if (!ptr) return UA_OK;
    const cudaError_t e = cudaIpcCloseMemHandle((char *)ptr - offset);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("ua_ipc_close: %s", cudaGetErrorString(e));
        return UA_ERR_CUDA;
    }
    return UA_OK;
}
