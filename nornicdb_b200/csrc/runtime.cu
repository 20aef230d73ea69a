// runtime.cu — error state, device info and grow-only workspaces.
#include <stdarg.h>
#include <string.h>

#include "kernels.cuh"

namespace nk {

static thread_local char g_err[512] = {0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }
void clear_error() { g_err[0] = 0; }

int query_device_info(int device_id, DeviceInfo *out) {
    cudaDeviceProp prop;
    NK_CUDA_OK(cudaGetDeviceProperties(&prop, device_id));
    out->device_id = device_id;
    out->num_sms = prop.multiProcessorCount;
    out->max_smem_optin = prop.sharedMemPerBlockOptin;
    out->cc = prop.major * 10 + prop.minor;
    return 0;
}

int ws_reserve(void **p, size_t *cur, size_t need) {
    if (need <= *cur && *p) return 0;
    if (*p) {
        // Earlier launches on any stream may still be using the old buffer.
        NK_CUDA_OK(cudaDeviceSynchronize());
        NK_CUDA_OK(cudaFree(*p));
        *p = nullptr;
        *cur = 0;
    }
    size_t sz = need + need / 4;  // slack so slowly growing Q / k do not re-allocate every call
    if (sz < 4096) sz = 4096;
    NK_CUDA_OK(cudaMalloc(p, sz));
    *cur = sz;
    return 0;
}

int Workspace::release() {
    void *ptrs[] = {cand, partial, keys, keys2, below, queries, out_idx, out_score, qaux, rownorm, flags, sub_rows, sub_gather, scratch};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    *this = Workspace();
    return 0;
}

}  // namespace nk
