// dense_inverse.cu -- in-place inverse of a dense SPD matrix in HBM (fp64), blocked Gauss-Jordan.
//
// Used ONCE per setQ to build the parity-mode preconditioner (Q + 0.1 I)^-1 that the reference
// obtains through a CHOLMOD factorisation (ref: src/QuadraticProblem.cpp:31-42).  SPD input, so no
// pivoting is needed.  Per block step kb (block size B):
//   Pinv = inv(A_kk);  Rw = Pinv * A_k,:  ;  C = A_:,k
//   A_ij -= C_i Rw_j (i,j != k);  A_ik = -C_i Pinv;  A_kj = Rw_j;  A_kk = Pinv
// 2 N^3 flops in N/B rank-B updates; each update streams A once (N^2 * 16 B of traffic).
#include <cuda_runtime.h>
#include "dpgo_kernels.cuh"

namespace dpgo {

constexpr int GJB = 32;     // pivot block
constexpr int GJT = 64;     // update tile

// invert the bs x bs pivot block (identity padded to GJB) -- one CTA of GJB x GJB threads
__global__ void k_gj_pivot(const double *__restrict__ A, int N, int k0, int bs, double *__restrict__ piv) {
  __shared__ double P[GJB][GJB + 1];
  const int i = threadIdx.y, j = threadIdx.x;
  double v = (i == j) ? 1.0 : 0.0;
  if (i < bs && j < bs) v = A[(size_t)(k0 + i) + (size_t)N * (k0 + j)];
  P[i][j] = v;
  __syncthreads();
  for (int k = 0; k < GJB; ++k) {
    const double pkk = P[k][k];
    const double pik = P[i][k], pkj = P[k][j];
    __syncthreads();
    const double inv = 1.0 / pkk;
    double nv;
    if (i == k && j == k) nv = inv;
    else if (i == k) nv = pkj * inv;
    else if (j == k) nv = -pik * inv;
    else nv = P[i][j] - pik * pkj * inv;
    P[i][j] = nv;
    __syncthreads();
  }
  piv[i + GJB * j] = P[i][j];
}

// Rw[q, j] = sum_p Pinv[q,p] A[k0+p, j]  (GJB x N, stored q + GJB*j);  C[i, q] = A[i, k0+q] (i + N*q)
__global__ void k_gj_panels(const double *__restrict__ A, int N, int k0, int bs, const double *__restrict__ piv,
                            double *__restrict__ Rw, double *__restrict__ C) {
  __shared__ double P[GJB][GJB + 1];
  const int t = threadIdx.x;                 // 256 threads
  for (int e = t; e < GJB * GJB; e += blockDim.x) P[e % GJB][e / GJB] = piv[e];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + t;
  if (j < N) {
    double a[GJB];
#pragma unroll
    for (int p = 0; p < GJB; ++p) a[p] = (p < bs) ? A[(size_t)(k0 + p) + (size_t)N * j] : 0.0;
#pragma unroll 4
    for (int q = 0; q < GJB; ++q) {
      double s = 0.0;
#pragma unroll
      for (int p = 0; p < GJB; ++p) s = fma(P[q][p], a[p], s);
      Rw[q + (size_t)GJB * j] = s;
    }
    // column panel: A symmetric at every step?  No -- Gauss-Jordan intermediates are not symmetric,
    // so read the true column entries.
    for (int q = 0; q < GJB; ++q) C[(size_t)j + (size_t)N * q] = (q < bs) ? A[(size_t)j + (size_t)N * (k0 + q)] : 0.0;
  }
}

// rank-GJB update of one GJT x GJT tile
__global__ void __launch_bounds__(256) k_gj_update(double *__restrict__ A, int N, int k0, int bs,
                                                   const double *__restrict__ piv, const double *__restrict__ Rw,
                                                   const double *__restrict__ C) {
  __shared__ double sC[GJB][GJT + 1];   // sC[q][i]
  __shared__ double sR[GJB][GJT + 1];   // sR[q][j]
  const int i0 = blockIdx.y * GJT, j0 = blockIdx.x * GJT;
  const int t = threadIdx.x;
  for (int e = t; e < GJB * GJT; e += 256) {
    const int q = e / GJT, ii = e % GJT;
    sC[q][ii] = (i0 + ii < N) ? C[(size_t)(i0 + ii) + (size_t)N * q] : 0.0;
  }
  for (int e = t; e < GJB * GJT; e += 256) {
    const int jj = e / GJB, q = e % GJB;
    sR[q][jj] = (j0 + jj < N) ? Rw[q + (size_t)GJB * (j0 + jj)] : 0.0;
  }
  __syncthreads();
  const int ti = (t % 16) * 4, tj = (t / 16) * 4;
  double acc[4][4];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
#pragma unroll 8
  for (int q = 0; q < GJB; ++q) {
    double c[4], r[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) { c[x] = sC[q][ti + x]; r[x] = sR[q][tj + x]; }
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = fma(c[x], r[y], acc[x][y]);
  }
#pragma unroll
  for (int y = 0; y < 4; ++y) {
    const int j = j0 + tj + y;
    if (j >= N) continue;
    const bool jin = (j >= k0 && j < k0 + bs);
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int i = i0 + ti + x;
      if (i >= N) continue;
      const bool iin = (i >= k0 && i < k0 + bs);
      double *p = A + (size_t)i + (size_t)N * j;
      if (!iin && !jin) {
        *p = *p - acc[x][y];
      } else if (iin && !jin) {
        *p = sR[i - k0][tj + y];
      } else if (!iin && jin) {
        // -C_i * Pinv[:, j-k0]
        double s = 0.0;
        for (int q = 0; q < bs; ++q) s = fma(sC[q][ti + x], piv[q + GJB * (j - k0)], s);
        *p = -s;
      } else {
        *p = piv[(i - k0) + GJB * (j - k0)];
      }
    }
  }
}

cudaError_t dense_spd_inverse(double *A, int N, cudaStream_t stream) {
  double *piv = nullptr, *Rw = nullptr, *C = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&piv, sizeof(double) * GJB * GJB)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&Rw, sizeof(double) * GJB * (size_t)N)) != cudaSuccess) { cudaFree(piv); return e; }
  if ((e = cudaMalloc(&C, sizeof(double) * GJB * (size_t)N)) != cudaSuccess) { cudaFree(piv); cudaFree(Rw); return e; }
  const int tiles = (N + GJT - 1) / GJT;
  for (int k0 = 0; k0 < N; k0 += GJB) {
    const int bs = (N - k0 < GJB) ? (N - k0) : GJB;
    k_gj_pivot<<<1, dim3(GJB, GJB), 0, stream>>>(A, N, k0, bs, piv);
    k_gj_panels<<<(N + 255) / 256, 256, 0, stream>>>(A, N, k0, bs, piv, Rw, C);
    k_gj_update<<<dim3(tiles, tiles), 256, 0, stream>>>(A, N, k0, bs, piv, Rw, C);
  }
  e = cudaStreamSynchronize(stream);
  cudaFree(piv); cudaFree(Rw); cudaFree(C);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace dpgo
