// Throughput of MUFU.TANH / MUFU.EX2 / FFMA2 per SM and clock on this GPU (a roofline denominator for GELU-bound epilogues).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mufu_rate tools/probes/mufu_rate.cu && gpurun_out/mufu_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void rate_kernel(float* out, int iters, long long* cycles) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int warps) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * warps * 32);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  const int iters = 4096;
  rate_kernel<OP><<<sms, warps * 32>>>(out, iters, cyc);
  rate_kernel<OP><<<sms, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[1024];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < sms; ++i) mean += (double)h[i];
  mean /= sms;
  printf("%-10s warps/SM %2d: %.2f lane-ops / clk / SM\n", name, warps, (double)iters * 8 * warps * 32 / mean);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    run<0>("tanh", w);
    run<1>("ex2", w);
    run<2>("ffma", w);
  }
  return 0;
}
