// Mask element access shared by the kernels: masks arrive in whatever element type the caller has
// (torch.bool, HF int64 attention_mask, matchmaker float masks); nonzero = real token.
#pragma once

#include <stdint.h>

#include "../../include/matchmaker_b200.h"

namespace mmb {

__device__ __forceinline__ bool mask_at(const void* mask, int mask_dtype, int64_t idx) {
  switch (mask_dtype) {
    case MMB200_MASK_U8:
      return static_cast<const uint8_t*>(mask)[idx] != 0;
    case MMB200_MASK_I32:
      return static_cast<const int32_t*>(mask)[idx] != 0;
    case MMB200_MASK_I64:
      return static_cast<const int64_t*>(mask)[idx] != 0;
    case MMB200_MASK_F32:
      return static_cast<const float*>(mask)[idx] != 0.0f;
    default:
      return true;
  }
}

// Raw mask word (no test: keeps the load's consumer away from the load) + deferred test.
__device__ __forceinline__ uint64_t mask_raw(const void* mask, int mask_dtype, int64_t idx) {
  switch (mask_dtype) {
    case MMB200_MASK_U8:
      return static_cast<const uint8_t*>(mask)[idx];
    case MMB200_MASK_I32:
    case MMB200_MASK_F32:
      return static_cast<const uint32_t*>(mask)[idx];
    case MMB200_MASK_I64:
      return static_cast<const uint64_t*>(mask)[idx];
    default:
      return 1;
  }
}
__device__ __forceinline__ bool mask_test(uint64_t raw, int mask_dtype) {
  return mask_dtype == MMB200_MASK_F32 ? (__uint_as_float(static_cast<uint32_t>(raw)) != 0.0f) : (raw != 0);
}

}  // namespace mmb
