// Measures the fp64 issue ceilings of one B200: mma.sync.m8n8k4.f64 (DMMA) and plain DFMA, registers only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_peak scripts/dmma_peak.cu ; prints one JSON line.
// Used to put the fp64-pipe bound next to the HBM roofline of the dense preconditioner apply.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void __launch_bounds__(512) k_dmma(double *out, int iters) {
  double acc[CHAINS][2];
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c][0] = acc[c][1] = (double)c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) dmma(acc[c][0], acc[c][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(512) k_dfma(double *out, int iters) {
  double acc[CHAINS];
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = (double)c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 2, threads = 512, iters = 20000;
  double *out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms_dmma = 0, ms_dfma = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_dmma<8><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms_dmma, e0, e1);
    cudaEventRecord(e0);
    k_dfma<8><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms_dfma, e0, e1);
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"cuda\"}\n"); return 1; }
  const double warps = (double)blocks * threads / 32;
  const double dmma_flops = warps * iters * 8.0 * (2.0 * 8 * 8 * 4);
  const double dfma_flops = (double)blocks * threads * iters * 8.0 * 2.0;
  // dependent-chain sweep at the persistent kernel's occupancy (1 CTA of 512 threads per SM): DMMAs per SM per us
  float ms_c[4] = {0, 0, 0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_dmma<1><<<sms, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_c[0], e0, e1);
    cudaEventRecord(e0); k_dmma<2><<<sms, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_c[1], e0, e1);
    cudaEventRecord(e0); k_dmma<4><<<sms, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_c[2], e0, e1);
    cudaEventRecord(e0); k_dmma<8><<<sms, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_c[3], e0, e1);
  }
  const double w1 = (double)threads / 32 * iters;   // DMMAs per SM per chain
  printf("{\"one_cta_512\": {\"chains1\": %.1f, \"chains2\": %.1f, \"chains4\": %.1f, \"chains8\": %.1f, \"unit\": \"DMMA/SM/us\"}}\n",
         w1 * 1 / (ms_c[0] * 1e3), w1 * 2 / (ms_c[1] * 1e3), w1 * 4 / (ms_c[2] * 1e3), w1 * 8 / (ms_c[3] * 1e3));
  printf("{\"sms\": %d, \"dmma_m8n8k4_tflops\": %.3f, \"dfma_tflops\": %.3f, \"dmma_per_sm_per_us\": %.2f}\n", sms,
         dmma_flops / (ms_dmma * 1e-3) / 1e12, dfma_flops / (ms_dfma * 1e-3) / 1e12,
         warps * iters * 8.0 / sms / (ms_dmma * 1e3));
  return 0;
}
