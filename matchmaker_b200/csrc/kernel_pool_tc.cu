// tcgen05 forward kernel for cosine + RBF kernel pooling (placeholder until the 3xTF32 pipeline lands):
// reports "not handled" so the dispatcher uses the FFMA kernel.
#include "host_util.cuh"

namespace mmb {
struct KpParams;
int kernel_pool_fwd_tc(const KpParams&, const DeviceInfo&, cudaStream_t, bool* handled) {
  *handled = false;
  return MMB200_OK;
}
}  // namespace mmb
