// barrier_bench3.cu -- flavours of the fused "barrier + all-reduce in one round trip" phase end (packets carrying a
// value and the phase number), against the counter barrier + partials read-back of round 1.
//   F0: 16-byte packets, ld/st.relaxed.gpu.v2.b64          F1: 16-byte packets, ld.global.cg.v2 / st.global.cg.v2
//   F2: NCCL-LL style 8-byte words {32 value bits, 32-bit phase}, ld/st.relaxed.gpu.b64
//   F3: as F1 with ld.volatile / st.volatile               A : counter (red.release + ld.relaxed poll) + partials
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
constexpr int NRED = 4, THREADS = 512;
constexpr unsigned FULL = 0xffffffffu;
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}
struct __align__(16) Packet { double v; unsigned long long ph; };

template <int F> __device__ __forceinline__ void st_pk(Packet *p, double v, unsigned long long ph) {
  if (F == 0) asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(ph) : "memory");
  else if (F == 1) asm volatile("st.global.cg.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(ph) : "memory");
  else asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(ph) : "memory");
}
template <int F> __device__ __forceinline__ void ld_pk(const Packet *p, double &v, unsigned long long &ph) {
  long long vv;
  if (F == 0) asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(vv), "=l"(ph) : "l"(p) : "memory");
  else if (F == 1) asm volatile("ld.global.cg.v2.b64 {%0, %1}, [%2];" : "=l"(vv), "=l"(ph) : "l"(p) : "memory");
  else asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(vv), "=l"(ph) : "l"(p) : "memory");
  v = __longlong_as_double(vv);
}

template <int F> __global__ void __launch_bounds__(THREADS, 1) k_fused(int phases, Packet *pk, unsigned long long ph0, double *out, double *scratch) {
  __shared__ double sm_warp[(THREADS / 32) * NRED];
  __shared__ double sm_out[8 * NRED];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, G = gridDim.x;
  unsigned long long ph = ph0;
  double carry = 1.0, acc[NRED];
  for (int it = 0; it < phases; ++it) {
    scratch[(size_t)blockIdx.x * THREADS + threadIdx.x] = carry;
#pragma unroll
    for (int q = 0; q < NRED; ++q) acc[q] = carry * (q + 1) / (double)(G * THREADS);
#pragma unroll
    for (int q = 0; q < NRED; ++q) { double v = warp_sum(acc[q]); if (lane == 0) sm_warp[warp * NRED + q] = v; }
    __syncthreads();
    ++ph;
    if (F == 2) {
      unsigned long long *buf = reinterpret_cast<unsigned long long *>(pk) + (size_t)(ph & 1ull) * G * NRED * 2;
      if (threadIdx.x == 0) {
        double s[NRED];
        for (int q = 0; q < NRED; ++q) { s[q] = 0.0; for (int w = 0; w < nwarps; ++w) s[q] += sm_warp[w * NRED + q]; }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        for (int q = 0; q < NRED; ++q) {
          const unsigned long long bits = (unsigned long long)__double_as_longlong(s[q]);
          const unsigned long long w0 = (bits & 0xffffffffull) | ((ph & 0xffffffffull) << 32), w1 = (bits >> 32) | ((ph & 0xffffffffull) << 32);
          asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(buf + ((size_t)blockIdx.x * NRED + q) * 2), "l"(w0) : "memory");
          asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(buf + ((size_t)blockIdx.x * NRED + q) * 2 + 1), "l"(w1) : "memory");
        }
      }
      const int npoll = (G + 31) >> 5;
      if (warp < npoll) {
        const int c = warp * 32 + lane;
        double v[NRED] = {0, 0, 0, 0};
        if (c < G) {
          const unsigned long long *src = buf + (size_t)c * NRED * 2;
          bool ok;
          unsigned long long w[2 * NRED];
          do {
#pragma unroll
            for (int q = 0; q < 2 * NRED; ++q) asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w[q]) : "l"(src + q) : "memory");
            ok = true;
#pragma unroll
            for (int q = 0; q < 2 * NRED; ++q) ok = ok && ((w[q] >> 32) == (ph & 0xffffffffull));
          } while (!ok);
#pragma unroll
          for (int q = 0; q < NRED; ++q) v[q] = __longlong_as_double((long long)((w[2 * q] & 0xffffffffull) | (w[2 * q + 1] << 32)));
        }
#pragma unroll
        for (int q = 0; q < NRED; ++q) { const double t = warp_sum(v[q]); if (lane == 0) sm_out[warp * NRED + q] = t; }
      }
    } else {
      Packet *buf = pk + (size_t)(ph & 1ull) * G * NRED;
      if (threadIdx.x == 0) {
        double s[NRED];
        for (int q = 0; q < NRED; ++q) { s[q] = 0.0; for (int w = 0; w < nwarps; ++w) s[q] += sm_warp[w * NRED + q]; }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        for (int q = 0; q < NRED; ++q) st_pk<F>(buf + (size_t)blockIdx.x * NRED + q, s[q], ph);
      }
      const int npoll = (G + 31) >> 5;
      if (warp < npoll) {
        const int c = warp * 32 + lane;
        double v[NRED] = {0, 0, 0, 0};
        if (c < G) {
          const Packet *src = buf + (size_t)c * NRED;
          bool ok;
          do {
            unsigned long long t[NRED];
#pragma unroll
            for (int q = 0; q < NRED; ++q) ld_pk<F>(src + q, v[q], t[q]);
            ok = true;
#pragma unroll
            for (int q = 0; q < NRED; ++q) ok = ok && (t[q] == ph);
          } while (!ok);
        }
#pragma unroll
        for (int q = 0; q < NRED; ++q) { const double t = warp_sum(v[q]); if (lane == 0) sm_out[warp * NRED + q] = t; }
      }
    }
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (G + 31) / 32; ++w) tot += sm_out[w * NRED];
    carry = tot + __ldcg(scratch + (size_t)((blockIdx.x + 1) % G) * THREADS + threadIdx.x) * 1e-30;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = carry;
}

__global__ void __launch_bounds__(THREADS, 1) k_counter(int phases, double *partials, unsigned *counter, double *out, double *scratch) {
  __shared__ double sm_warp[(THREADS / 32) * NRED];
  __shared__ double sm_out[NRED];
  unsigned epoch = 0;
  int parity = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, G = gridDim.x;
  double acc[NRED], carry = 1.0;
  for (int ph = 0; ph < phases; ++ph) {
    scratch[(size_t)blockIdx.x * THREADS + threadIdx.x] = carry;
#pragma unroll
    for (int q = 0; q < NRED; ++q) acc[q] = carry * (q + 1) / (double)(G * THREADS);
#pragma unroll
    for (int q = 0; q < NRED; ++q) { double v = warp_sum(acc[q]); if (lane == 0) sm_warp[warp * NRED + q] = v; }
    __syncthreads();
    double *slot = partials + (size_t)parity * G * NRED;
    if ((int)threadIdx.x < NRED) { double s = 0.0; for (int w = 0; w < nwarps; ++w) s += sm_warp[w * NRED + threadIdx.x]; slot[(size_t)blockIdx.x * NRED + threadIdx.x] = s; }
    __syncthreads();
    epoch += G;
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned v;
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - epoch) < 0);
    }
    __syncthreads();
    if (warp == 0) {
      double s[NRED] = {0, 0, 0, 0};
      for (int c = lane; c < G; c += 32)
#pragma unroll
        for (int q = 0; q < NRED; ++q) s[q] += __ldcg(slot + (size_t)c * NRED + q);
#pragma unroll
      for (int q = 0; q < NRED; ++q) { const double t = warp_sum(s[q]); if (lane == 0) sm_out[q] = t; }
    }
    __syncthreads();
    carry = sm_out[0] + __ldcg(scratch + (size_t)((blockIdx.x + 1) % G) * THREADS + threadIdx.x) * 1e-30;
    __syncthreads();
    parity ^= 1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = carry;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
template <class L> float timeit(L launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); for (int i = 0; i < 5; ++i) launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / 5;
}
int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  Packet *pk; double *out, *scratch, *partials; unsigned *counter;
  CK(cudaMalloc(&pk, sizeof(Packet) * 2 * 256 * NRED)); CK(cudaMemset(pk, 0, sizeof(Packet) * 2 * 256 * NRED));
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&scratch, sizeof(double) * 256 * THREADS)); CK(cudaMalloc(&partials, sizeof(double) * 2 * 256 * NRED)); CK(cudaMalloc(&counter, 64));
  const int grids[] = {sms, 16};
  for (int G : grids) {
    for (int variant = 0; variant < 5; ++variant) {
      float t[2]; const int counts[2] = {20, 420};
      unsigned long long ph0 = 0;
      for (int rep = 0; rep < 2; ++rep) {
        int phases = counts[rep];
        auto launch = [&]() {
          if (variant == 4) {
            CK(cudaMemsetAsync(counter, 0, 4));
            void *args[] = {&phases, &partials, &counter, &out, &scratch};
            CK(cudaLaunchCooperativeKernel((void *)k_counter, dim3(G), dim3(THREADS), args, 0, 0));
          } else {
            void *args[] = {&phases, &pk, &ph0, &out, &scratch};
            void *fn = variant == 0 ? (void *)k_fused<0> : variant == 1 ? (void *)k_fused<1> : variant == 2 ? (void *)k_fused<2> : (void *)k_fused<3>;
            CK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(THREADS), args, 0, 0));
            ph0 += phases;
          }
        };
        t[rep] = timeit(launch);
      }
      double hv; CK(cudaMemcpy(&hv, out, 8, cudaMemcpyDeviceToHost));
      printf("{\"variant\": \"%s\", \"grid\": %d, \"us_per_phase\": %.3f, \"check\": %.6f}\n",
             variant == 0 ? "F0 relaxed.gpu.v2" : variant == 1 ? "F1 cg.v2" : variant == 2 ? "F2 LL 8-byte" : variant == 3 ? "F3 volatile.v2" : "A counter+partials",
             G, 1e3 * (t[1] - t[0]) / 400.0, hv);
      fflush(stdout);
    }
  }
  return 0;
}
