// Internal interface of the max-sim kernels (shared by maxsim.cu and maxsim_host.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mmb {

struct MaxsimParams {
  const void* q;
  const void* d;
  const void* q_mask;
  const void* d_mask;
  const int32_t* pair_q;
  const int32_t* pair_d;
  const int32_t* pair_dmask;  // row of d_mask used for pair p (default: the document index)
  const int32_t* rows_needed;  // [n_d] or NULL: rows of document di worth fetching (1 + last unmasked row)
  float* out;
  int32_t* argmax;
  int64_t n_q, n_d, n_pairs;
  int64_t pair_base;  // query of pair p (when pair_q == NULL) is (p + pair_base) / docs_per_query
  int32_t docs_per_query, Lq, Ld, dim, mask_dtype;
};

struct DeviceInfo;
// maxsim_qm.cu: "queries on M" tcgen05 kernel (Lq <= 32, dim 64/128); *handled = false if out of envelope.
int maxsim_qm_launch(const MaxsimParams& P, int dtype, const DeviceInfo& dev, cudaStream_t stream, bool* handled);
// rows_needed[di] = 1 + index of the last unmasked row of document di (Ld when d_mask == NULL)
int maxsim_rows_needed_launch(const void* d_mask, int mask_dtype, int32_t* rows_needed, int64_t n_d, int Ld, cudaStream_t stream);

// Validates, picks SIMT or tcgen05 and launches on `stream`.
int maxsim_fwd_device(const MaxsimParams& P, int dtype, int impl, cudaStream_t stream);

}  // namespace mmb
