// Parameter block shared by the kernel-pooling kernels (kernel_pool.cu: FFMA forward/backward;
// kernel_pool_ts.cu: tcgen05 forward).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mmb {

struct KpParams {
  const float* q;
  const float* d;
  const void* q_mask;
  const void* d_mask;
  const float* mu;
  const float* sigma;
  const float* alpha;
  const float* weight;
  int64_t B;
  int32_t Lq, Ld, D, K, mask_dtype;
  float log_scale;
  // forward outputs
  float* score;
  float* per_kernel;
  float* per_kernel_query;
  float* cosine;
  // backward
  const float* S;
  const float* grad_score;
  float* grad_q;
  float* grad_d;
  float* ws_weight;  // [B,K]
  float* ws_alpha;   // [B,K]
};

struct DeviceInfo;
// tcgen05 forward with the document operand in tensor memory (kernel_pool_ts.cu); *handled = false when the shape is
// outside the kernel's envelope
int kernel_pool_fwd_ts(const KpParams& P, const DeviceInfo& dev, cudaStream_t stream, bool* handled);

}  // namespace mmb
