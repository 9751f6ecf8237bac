// Library-level C-ABI entry points: version, device probe, last-error string.
#include <cstdio>
#include <cstring>

#include "tr_common.h"

namespace tr {
static thread_local std::string g_last_error;

int fail(int code, const std::string &msg) { g_last_error = msg; return code; }

int fail_cuda(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) in `%s` at %s:%d", (int) e, cudaGetErrorString(e), what, file, line);
    g_last_error = buf;
    cudaGetLastError();  // clear the sticky (non-fatal) error state
    return e == cudaErrorMemoryAllocation ? TR_ERR_OUT_OF_MEMORY : TR_ERR_CUDA;
}
}  // namespace tr

extern "C" {

const char *tr_last_error(void) { return tr::g_last_error.c_str(); }

int tr_abi_version(void) { return TR_ABI_VERSION; }

// The product has no CPU path: without an sm_100 class device every compute entry point fails.
int tr_device_check(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return tr::fail(TR_ERR_NO_DEVICE, "no CUDA device visible: mitransient_b200 has no CPU fallback");
    }
    int dev = 0;
    TR_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    TR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    TR_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) {
        char buf[160];
        snprintf(buf, sizeof(buf), "device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, major, minor);
        return tr::fail(TR_ERR_NO_DEVICE, buf);
    }
    return TR_OK;
}

// ---- multi-GPU film assembly over NVLink peer memory (see the header)
int tr_peer_alloc(uint64_t bytes, void **dev_ptr, unsigned char *handle) {
    if (!dev_ptr || !handle || bytes == 0) return tr::fail(TR_ERR_INVALID_ARGUMENT, "tr_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == TR_PEER_HANDLE_BYTES, "IPC handle size");
    int rc = tr_device_check();
    if (rc != TR_OK) return rc;
    void *p = nullptr;
    TR_CUDA(cudaMalloc(&p, bytes));
    TR_CUDA_OR(cudaMemset(p, 0, bytes), { cudaFree(p); });
    cudaIpcMemHandle_t h;
    TR_CUDA_OR(cudaIpcGetMemHandle(&h, p), { cudaFree(p); });
    memcpy(handle, &h, sizeof(h));
    *dev_ptr = p;
    return TR_OK;
}

int tr_peer_open(const unsigned char *handle, void **dev_ptr) {
    if (!dev_ptr || !handle) return tr::fail(TR_ERR_INVALID_ARGUMENT, "tr_peer_open: bad argument");
    int rc = tr_device_check();
    if (rc != TR_OK) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    TR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return TR_OK;
}

int tr_peer_close(void *dev_ptr) {
    if (!dev_ptr) return TR_OK;
    TR_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return TR_OK;
}

int tr_peer_free(void *dev_ptr) {
    if (!dev_ptr) return TR_OK;
    TR_CUDA(cudaFree(dev_ptr));
    return TR_OK;
}

}  // extern "C"
