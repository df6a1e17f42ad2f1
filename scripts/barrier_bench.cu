// barrier_bench.cu -- microbenchmark of the grid-wide "phase end" (barrier + 4-scalar all-reduce) variants of the
// persistent kernel.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench scripts/barrier_bench.cu
//   A : arrival counter (red.release / ld.acquire poll), partials read back after the barrier  (round-1 protocol)
//   B : all-to-all flagged slots: every CTA publishes {value, epoch} pairs with 16-byte stores, every CTA polls all
//       slots -- barrier and all-reduce in ONE L2 round trip, no atomics
//   C : cluster barrier (barrier.cluster) for grids of one cluster (<= 16 CTAs), values through distributed smem
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

constexpr int NRED = 4;
constexpr int THREADS = 512;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---------------- variant A ----------------
__global__ void __launch_bounds__(THREADS, 1) k_A(int phases, double *partials, unsigned *counter, double *out, double *scratch) {
  __shared__ double sm_warp[(THREADS / 32) * NRED];
  __shared__ double sm_out[NRED];
  unsigned epoch = 0;
  int parity = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double acc[NRED];
  double carry = 1.0;
  for (int ph = 0; ph < phases; ++ph) {
    scratch[(size_t)blockIdx.x * THREADS + threadIdx.x] = carry;     // a data write the barrier has to publish
#pragma unroll
    for (int q = 0; q < NRED; ++q) acc[q] = carry * (q + 1) / (double)(gridDim.x * THREADS);
#pragma unroll
    for (int q = 0; q < NRED; ++q) {
      double v = warp_sum(acc[q]);
      if (lane == 0) sm_warp[warp * NRED + q] = v;
    }
    __syncthreads();
    double *slot = partials + (size_t)parity * gridDim.x * NRED;
    if ((int)threadIdx.x < NRED) {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s += sm_warp[w * NRED + threadIdx.x];
      slot[(size_t)blockIdx.x * NRED + threadIdx.x] = s;
    }
    __syncthreads();
    epoch += gridDim.x;
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      while ((int)(ld_acquire_u32(counter) - epoch) < 0) { }
    }
    __syncthreads();
    if (warp == 0) {
      double s[NRED];
#pragma unroll
      for (int q = 0; q < NRED; ++q) s[q] = 0.0;
      for (int c = lane; c < (int)gridDim.x; c += 32)
#pragma unroll
        for (int q = 0; q < NRED; ++q) s[q] += __ldcg(slot + (size_t)c * NRED + q);
#pragma unroll
      for (int q = 0; q < NRED; ++q) {
        const double t = warp_sum(s[q]);
        if (lane == 0) sm_out[q] = t;
      }
    }
    __syncthreads();
    carry = sm_out[0] + __ldcg(scratch + (size_t)((blockIdx.x + 1) % gridDim.x) * THREADS + threadIdx.x) * 1e-30;
    __syncthreads();
    parity ^= 1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = carry;
}

// ---------------- variant B ----------------
struct __align__(16) Flagged { double v; unsigned long long tag; };

__device__ __forceinline__ void st_flagged(Flagged *p, double v, unsigned long long tag) {
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
}
__device__ __forceinline__ void ld_flagged(const Flagged *p, double &v, unsigned long long &tag) {
  long long vv;
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(vv), "=l"(tag) : "l"(p) : "memory");
  v = __longlong_as_double(vv);
}

__global__ void __launch_bounds__(THREADS, 1) k_B(int phases, Flagged *slots, unsigned long long epoch0, double *out, double *scratch) {
  __shared__ double sm_warp[(THREADS / 32) * NRED];
  __shared__ double sm_red[8 * NRED];
  unsigned long long epoch = epoch0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int G = gridDim.x;
  const int npoll = (G + 31) / 32;           // polling warps
  double acc[NRED];
  double carry = 1.0;
  for (int ph = 0; ph < phases; ++ph) {
    scratch[(size_t)blockIdx.x * THREADS + threadIdx.x] = carry;
#pragma unroll
    for (int q = 0; q < NRED; ++q) acc[q] = carry * (q + 1) / (double)(gridDim.x * THREADS);
#pragma unroll
    for (int q = 0; q < NRED; ++q) {
      double v = warp_sum(acc[q]);
      if (lane == 0) sm_warp[warp * NRED + q] = v;
    }
    __syncthreads();
    ++epoch;
    Flagged *base = slots + (size_t)(epoch & 1) * G * NRED;
    if ((int)threadIdx.x < NRED) {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s += sm_warp[w * NRED + threadIdx.x];
      __threadfence();                                   // release: the CTA's data writes (ordered by the bar.sync) first
      st_flagged(base + (size_t)blockIdx.x * NRED + threadIdx.x, s, epoch);
    }
    if (warp < npoll) {
      const int c = warp * 32 + lane;
      double v[NRED];
#pragma unroll
      for (int q = 0; q < NRED; ++q) v[q] = 0.0;
      if (c < G) {
#pragma unroll
        for (int q = 0; q < NRED; ++q) {
          unsigned long long tag;
          do { ld_flagged(base + (size_t)c * NRED + q, v[q], tag); } while (tag != epoch);
        }
      }
      __threadfence();                                   // acquire side
#pragma unroll
      for (int q = 0; q < NRED; ++q) {
        const double t = warp_sum(v[q]);
        if (lane == 0) sm_red[warp * NRED + q] = t;
      }
    }
    __syncthreads();
    double tot[NRED];
#pragma unroll
    for (int q = 0; q < NRED; ++q) {
      double t = 0.0;
      for (int w = 0; w < npoll; ++w) t += sm_red[w * NRED + q];
      tot[q] = t;
    }
    carry = tot[0] + __ldcg(scratch + (size_t)((blockIdx.x + 1) % gridDim.x) * THREADS + threadIdx.x) * 1e-30;
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = carry;
}

// ---------------- variant C: one cluster ----------------
__global__ void __launch_bounds__(THREADS, 1) k_C(int phases, double *out, double *scratch) {
  __shared__ double sm_warp[(THREADS / 32) * NRED];
  __shared__ double sm_pub[2][NRED];
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int G = cluster.num_blocks();
  double acc[NRED];
  double carry = 1.0;
  int parity = 0;
  for (int ph = 0; ph < phases; ++ph) {
    scratch[(size_t)blockIdx.x * THREADS + threadIdx.x] = carry;
#pragma unroll
    for (int q = 0; q < NRED; ++q) acc[q] = carry * (q + 1) / (double)(gridDim.x * THREADS);
#pragma unroll
    for (int q = 0; q < NRED; ++q) {
      double v = warp_sum(acc[q]);
      if (lane == 0) sm_warp[warp * NRED + q] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < NRED) {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s += sm_warp[w * NRED + threadIdx.x];
      sm_pub[parity][threadIdx.x] = s;
    }
    __threadfence();          // global data writes visible at gpu scope before the cluster barrier releases the readers
    cluster.sync();
    double tot[NRED];
#pragma unroll
    for (int q = 0; q < NRED; ++q) tot[q] = 0.0;
    for (int c = 0; c < G; ++c) {
      const double *rp = cluster.map_shared_rank(&sm_pub[parity][0], c);
#pragma unroll
      for (int q = 0; q < NRED; ++q) tot[q] += rp[q];
    }
    carry = tot[0] + __ldcg(scratch + (size_t)((blockIdx.x + 1) % gridDim.x) * THREADS + threadIdx.x) * 1e-30;
    parity ^= 1;
  }
  cluster.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = carry;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <class F> float time_it(F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 10; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / 10;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double *partials, *out, *scratch; unsigned *counter; Flagged *slots;
  CK(cudaMalloc(&partials, sizeof(double) * 2 * 256 * NRED));
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&counter, 64));
  CK(cudaMalloc(&scratch, sizeof(double) * 256 * THREADS));
  CK(cudaMalloc(&slots, sizeof(Flagged) * 2 * 256 * NRED));
  CK(cudaMemset(slots, 0, sizeof(Flagged) * 2 * 256 * NRED));
  const int grids[] = {sms, 64, 32, 16, 8};
  for (int G : grids) {
    for (int variant = 0; variant < 3; ++variant) {
      if (variant == 2 && G > 16) continue;
      float t[2];
      const int counts[2] = {20, 420};
      unsigned long long epoch0 = 0;
      for (int rep = 0; rep < 2; ++rep) {
        int phases = counts[rep];
        auto launch = [&]() {
          if (variant == 0) {
            CK(cudaMemsetAsync(counter, 0, 4));
            void *args[] = {&phases, &partials, &counter, &out, &scratch};
            CK(cudaLaunchCooperativeKernel((void *)k_A, dim3(G), dim3(THREADS), args, 0, 0));
          } else if (variant == 1) {
            void *args[] = {&phases, &slots, &epoch0, &out, &scratch};
            CK(cudaLaunchCooperativeKernel((void *)k_B, dim3(G), dim3(THREADS), args, 0, 0));
            epoch0 += phases;
          } else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G); cfg.blockDim = dim3(THREADS);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            if (G > 8) CK(cudaFuncSetAttribute(k_C, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            CK(cudaLaunchKernelEx(&cfg, k_C, phases, out, scratch));
          }
        };
        t[rep] = time_it(launch);
      }
      double hv; CK(cudaMemcpy(&hv, out, 8, cudaMemcpyDeviceToHost));
      printf("{\"variant\": \"%c\", \"grid\": %d, \"us_per_phase\": %.3f, \"us_launch\": %.2f, \"check\": %.6f}\n", "ABC"[variant], G,
             1e3 * (t[1] - t[0]) / 400.0, 1e3 * t[0] - 20 * 1e3 * (t[1] - t[0]) / 400.0, hv);
    }
  }
  return 0;
}
