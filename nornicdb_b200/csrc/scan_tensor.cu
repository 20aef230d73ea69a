// scan_tensor.cu — tcgen05 / TMEM / TMA fused distance + top-k scan (large-Q path).  Placeholder until the
// tensor-core kernel lands: reports "unsupported" so AUTO dispatch stays on the CUDA-core scan.
#include "kernels.cuh"

namespace nk {
bool scan_tensor_supported(const DeviceInfo &, const ScanArgs &) { return false; }
int scan_tensor(const DeviceInfo &, const ScanArgs &, Workspace &, uint64_t *, uint64_t *) {
    set_error("tensor path not built");
    return -1;
}
}  // namespace nk
