// Microbenchmark: MUFU.EX2 issue rate per SM on sm_100a, alone and mixed with FP32 work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int FP_PER_MUFU>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); acc[i] = 0.f; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float u = x[i];
#pragma unroll
      for (int f = 0; f < FP_PER_MUFU; ++f) u = fmaf(u, 0.999f, -0.0001f);
      acc[i] += ex2f(u);
      x[i] = u;
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int F>
void run(int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<F><<<148, threads>>>(out, cyc, iters);
  k<F><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double mufu = (double)iters * 8 * threads;
  printf("fp_per_mufu=%d threads=%4d: %.2f MUFU/clk/SM, %.2f instr-lanes/clk/SM\n", F, threads, mufu / h[0], mufu * (F + 2) / h[0]);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int t : {128, 256, 512, 1024}) run<0>(t);
  for (int t : {256, 512}) run<1>(t);
  for (int t : {256, 512}) run<3>(t);
  for (int t : {256, 512}) run<4>(t);
  for (int t : {256, 512}) run<8>(t);
  return 0;
}
