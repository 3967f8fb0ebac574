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
  // variants of the same pooling: TK-Sparse gates every document term (cikm20_tk_sparse.py:135),
  // IDCM's ESM clamps at 1e-4 and adds the bias of its Linear(11, 1) (sigir21_idcm.py:185-186)
  const float* gate;     // [B, Ld] multiplier of the activations of document term j (values < 0 count as 0), or nullptr
  float clamp_min;       // floor of alpha_k * S_ik before the log (1e-10 in KNRM / TK / TK-Sparse / Conv-KNRM)
  float bias;            // added to the score
  float* grad_gate;      // [B, Ld] backward output, or nullptr
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
