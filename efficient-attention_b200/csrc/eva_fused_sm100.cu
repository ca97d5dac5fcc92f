// Fused sm_100a EVA forward (TMA + tcgen05).  Placeholder until the kernel lands: reports
// "unsupported" so eva_forward takes the generic two-stage path.
#include "common.cuh"
#include "launch.h"

namespace eva {

bool fused_supported(const Geo&, int, const View&, const View&, const View&, const uint8_t*, const EvaAdaptive&,
                     const float*, long long) {
  return false;
}
size_t fused_workspace_bytes(const Geo&) { return 0; }
cudaError_t launch_fused(const Geo&, int, const View&, const View&, const View&, const EvaAdaptive&, const float*,
                         const float*, long long, void*, void*, cudaStream_t, const char** msg) {
  *msg = "fused path not built";
  return cudaErrorNotSupported;
}

}  // namespace eva
