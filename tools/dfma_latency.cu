// DFMA throughput as a function of resident warps per SM and independent chains per thread (B200, sm_100a).
// Answers: how much FP64 pipe can 8 warps/SM (255 registers/thread) feed at a given ILP?
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm, int nsm, double* out) {
  int threads = 32 * (warps_per_sm >= 4 ? 4 : warps_per_sm);  // 128-thread blocks
  int blocks_per_sm = warps_per_sm * 32 / threads;
  int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // dynamic smem limits residency to exactly blocks_per_sm blocks per SM
  size_t smem = 200 * 1024 / blocks_per_sm;
  cudaFuncSetAttribute(k<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<ILP><<<nsm * blocks_per_sm, threads, smem>>>(out, 10, 0.999999, 1e-9);
  cudaEventRecord(e0);
  k<ILP><<<nsm * blocks_per_sm, threads, smem>>>(out, iters, 0.999999, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 2.0 * 16 * ILP * (double)iters * nsm * blocks_per_sm * threads;
  printf("warps/SM=%2d ILP=%d  %.2f TFLOP/s\n", warps_per_sm, ILP, fl / (ms * 1e-3) / 1e12);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 64 * 32);
  for (int w : {4, 8, 12, 16, 32, 64}) { run<1>(w, p.multiProcessorCount, out); run<2>(w, p.multiProcessorCount, out); run<4>(w, p.multiProcessorCount, out); run<8>(w, p.multiProcessorCount, out); }
  return 0;
}
