// lbdrn_tc.cu -- tcgen05 / TMA tensor-core decode path (placeholder until the kernel lands).
#include "lbdrn_internal.h"

namespace lbdrn {
bool tc_supported(const Net&) { return false; }
int tc_decode(const Net&, const void*, const float*, uint16_t*, cudaStream_t) {
  return fail(LBDRN_E_UNSUPPORTED, "tensor-core path not built");
}
}  // namespace lbdrn
