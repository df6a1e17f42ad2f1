// barrier_bench2.cu -- where does the time of a grid-wide phase end go?  Instrumented variants (clock64 in CTA 0 /
// thread 0, averaged over the phases).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench2 ...
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int THREADS = 512;
constexpr int NSEG = 8;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// MODE 0: red.release + ld.acquire poll                      (round-1 protocol)
// MODE 1: fence.acq_rel.gpu ; red.relaxed ; ld.relaxed poll ; fence.acq_rel.gpu
// MODE 2: as 1 but per-CTA flag words (st.relaxed to own slot; one warp polls all slots with ld.relaxed) -- no atomics
// MODE 3: as 2 without any fence (NOT a correct barrier for data; lower bound of the flag round trip)
// DATA : 0 = no data traffic, 1 = every thread stores one double before and loads a neighbour CTA's after
template <int MODE, int DATA>
__global__ void __launch_bounds__(THREADS, 1) k_bar(int phases, unsigned *counter, unsigned *flags, unsigned epoch0, double *scratch,
                                                    long long *seg_out, double *out) {
  const int G = gridDim.x;
  unsigned epoch = epoch0;
  unsigned target = epoch0 * G;
  long long seg[NSEG];
  for (int i = 0; i < NSEG; ++i) seg[i] = 0;
  double carry = 1.0;
  const bool timer = (blockIdx.x == 0 && threadIdx.x == 0);
  for (int ph = 0; ph < phases; ++ph) {
    long long t0 = clock64();
    if (DATA) scratch[(size_t)blockIdx.x * THREADS + threadIdx.x] = carry;
    __syncthreads();
    long long t1 = clock64();
    ++epoch;
    target += G;
    long long t2 = t1, t3 = t1, t4 = t1;
    if (MODE == 0) {
      if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        t2 = clock64();
        while ((int)(ld_acquire_u32(counter) - target) < 0) { }
        t3 = clock64();
        t4 = t3;
      }
    } else if (MODE == 1) {
      if (threadIdx.x == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        t2 = clock64();
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        while ((int)(ld_relaxed_u32(counter) - target) < 0) { }
        t3 = clock64();
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        t4 = clock64();
      }
    } else {
      if (threadIdx.x == 0) {
        if (MODE == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        t2 = clock64();
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * 32), "r"(epoch) : "memory");
      }
      if (threadIdx.x < 32 * ((G + 31) / 32)) {
        const int c = threadIdx.x;
        if (c < G) while ((int)(ld_relaxed_u32(flags + c * 32) - epoch) < 0) { }   // monotonic: a fast CTA may already be one phase ahead
        __syncwarp();
        if (threadIdx.x == 0) t3 = clock64();
        if (MODE == 2 && (threadIdx.x & 31) == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        if (threadIdx.x == 0) t4 = clock64();
      }
    }
    __syncthreads();
    long long t5 = clock64();
    if (DATA) carry += __ldcg(scratch + (size_t)((blockIdx.x + 1) % G) * THREADS + threadIdx.x) * 1e-30;
    long long t6 = clock64();
    if (timer) {
      seg[0] += t1 - t0;   // data store + bar.sync
      seg[1] += t2 - t1;   // release side (fence / red.release)
      seg[2] += t3 - t2;   // arrive + poll
      seg[3] += t4 - t3;   // acquire fence
      seg[4] += t5 - t4;   // closing bar.sync
      seg[5] += t6 - t5;   // data load
    }
  }
  if (timer) {
    for (int i = 0; i < NSEG; ++i) seg_out[i] = seg[i];
    out[0] = carry;
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE, int DATA> void run(int G, unsigned *counter, unsigned *flags, double *scratch, long long *seg, double *out) {
  const int phases = 400;
  unsigned epoch0 = 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaMemset(counter, 0, 4));
    CK(cudaMemset(flags, 0, 4 * 32 * 256));
    epoch0 = 0;
    int ph = phases;
    void *args[] = {&ph, &counter, &flags, &epoch0, &scratch, &seg, &out};
    CK(cudaEventRecord(e0));
    CK(cudaLaunchCooperativeKernel((void *)k_bar<MODE, DATA>, dim3(G), dim3(THREADS), args, 0, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  long long h[NSEG];
  CK(cudaMemcpy(h, seg, sizeof(h), cudaMemcpyDeviceToHost));
  printf("{\"mode\": %d, \"data\": %d, \"grid\": %d, \"us_per_phase\": %.3f, \"cycles\": {\"store_sync\": %.0f, \"release\": %.0f, \"arrive_poll\": %.0f, "
         "\"acquire\": %.0f, \"sync\": %.0f, \"load\": %.0f}}\n", MODE, DATA, G, 1e3 * best / phases, h[0] / (double)phases, h[1] / (double)phases,
         h[2] / (double)phases, h[3] / (double)phases, h[4] / (double)phases, h[5] / (double)phases);
  fflush(stdout);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned *counter, *flags; double *scratch, *out; long long *seg;
  CK(cudaMalloc(&counter, 64)); CK(cudaMalloc(&flags, 4 * 32 * 256));
  CK(cudaMalloc(&scratch, sizeof(double) * 256 * THREADS)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&seg, 8 * NSEG));
  const int grids[] = {sms, 16};
  for (int G : grids) {
    run<0, 0>(G, counter, flags, scratch, seg, out);
    run<0, 1>(G, counter, flags, scratch, seg, out);
    run<1, 0>(G, counter, flags, scratch, seg, out);
    run<1, 1>(G, counter, flags, scratch, seg, out);
    run<2, 0>(G, counter, flags, scratch, seg, out);
    run<2, 1>(G, counter, flags, scratch, seg, out);
    run<3, 0>(G, counter, flags, scratch, seg, out);
    run<3, 1>(G, counter, flags, scratch, seg, out);
  }
  return 0;
}
