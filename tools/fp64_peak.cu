// fp64_peak.cu — measures the DFMA issue rate of the device (the FP64 roofline denominator that
// MEASURED_PEAKS.json does not carry).  8 independent FMA chains per thread, 1024 threads per CTA,
// 2 CTAs per SM worth of work per SM; reports TFLOP/s (2 flops per FMA) from CUDA-event time.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 2, threads = 1024, iters = 1 << 14;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double flops = 2.0 * 8 * (double)iters * blocks * threads;
  printf("{\"device\": \"%s\", \"sms\": %d, \"fp64_tflops\": %.3f, \"ms\": %.4f, \"dfma_per_clk_per_sm_at_1965MHz\": %.2f}\n", p.name,
         p.multiProcessorCount, flops / (best * 1e-3) / 1e12, best,
         flops / 2 / (best * 1e-3) / p.multiProcessorCount / 1.965e9);
  return 0;
}
