// Microbenchmark: cycles per tcgen05.mma on sm_100a for the shapes the kernels in this repo use.
// A from shared memory (SS) or tensor memory (TS); kind::tf32 (K = 8) and kind::f16 (K = 16); M = 128.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../matchmaker_b200/csrc -o umma_bench umma_bench.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace mmb;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 SS tf32, 1 TS tf32, 2 SS bf16, 3 TS bf16.  ndst: number of distinct accumulators cycled through.
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int ndst, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t fmt = mode < 2 ? kFmtTF32 : kFmtBF16;
    const uint32_t idesc = make_idesc(fmt, 128, N);
    const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smem));
    const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smem) + 16384);
    const uint32_t a_tmem = tb + 480;  // last 32 columns
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tb + (uint32_t)((i % ndst) * N);
      if (mode == 0) umma_tf32(d, adesc, bdesc, idesc, 1u);
      else if (mode == 1) umma_tf32_ts(d, a_tmem, bdesc, idesc, 1u);
      else if (mode == 2) umma_f16(d, adesc, bdesc, idesc, 1u);
      else umma_f16_ts(d, a_tmem, bdesc, idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const char* names[4] = {"SS tf32 K=8 ", "TS tf32 K=8 ", "SS bf16 K=16", "TS bf16 K=16"};
  const int iters = 4096;
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {64, 128, 256})
      for (int ndst : {1, 2}) {
        if (ndst * N > 448) continue;
        for (int grid : {1, 148}) {
          k<<<grid, 128, 80 * 1024>>>(mode, N, ndst, iters, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
          printf("%s M=128 N=%3d ndst=%d grid=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma\n", names[mode], N, ndst, grid,
                 (double)h[0] / iters, (double)h[1] / iters);
        }
      }
  return 0;
}
