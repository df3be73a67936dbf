// lbdrn_tc.cuh -- tcgen05 / TMA tensor-core decode path (placeholder until the kernel lands).
#pragma once
#include <string>

#include "lbdrn_common.cuh"

namespace lbdrn {
inline bool tc_supported(const Net&) { return false; }
inline int tc_decode(const Net&, const void*, const float*, uint16_t*, cudaStream_t, std::string& err, long long&) {
  err = "tensor-core path not built";
  return LBDRN_E_UNSUPPORTED;
}
}  // namespace lbdrn
