// dpgo_kernels.cu -- sm_100a kernels of the pose-graph hot path.
//
//  k_optimize<R,DH> : ONE persistent cooperative kernel per QuadraticOptimizer::optimize() call
//                     (ref: src/QuadraticOptimizer.cpp:34-149 + ROPTLIB RTRNewton/tCG).  All phases
//                     -- fused [X.Q + G, f, tangent projection, |g|, preconditioner] passes,
//                     Riemannian Hessian-vector products of the truncated-CG loop, vector updates,
//                     QF retraction -- run inside it, separated by grid barriers; scalar
//                     reductions are fixed-order (deterministic) and every CTA replays the same
//                     scalar control flow.  The host sees only the result record.
//  k_spmv<R,DH>     : the Q.X product alone (Out = X Q [+ G]) -- the roofline kernel.
//  small kernels    : Stiefel (polar) projection, public-pose packing, G assembly.
#include "dpgo_device.cuh"
#include "dpgo_kernels.cuh"
#include <cooperative_groups.h>
#include <algorithm>

namespace dpgo {

// ---------------------------------------------------------------------------------------------
// phase-end reduction: block partials -> global partials -> grid barrier -> every CTA sums all
// partials in the same fixed order, so all CTAs hold bit-identical scalars.
// ---------------------------------------------------------------------------------------------
struct BlockCtx {
  double *sm_warp;    // [nwarps * NRED]
  double *sm_out;     // [2 * NRED]  totals of the phase, double-buffered by parity
  unsigned epoch;
  int parity;
};

// The reduction steps around the barrier are kept short (they are pure latency, ~80 times per step): the 16 warp
// partials are combined by a shuffle tree in warp 0 (fixed order), lane 0 stores the CTA's partials and arrives at the
// barrier right behind them (the release covers the store), the whole of warp 0 polls the counter (one broadcast request)
// and goes straight on to fetch all CTAs' partials; the totals are published through a parity-double-buffered shared slot,
// so a phase end has two bar.syncs, not four.  (A fused variant -- 16-byte {value, phase} packets polled all-to-all, barrier
// and all-reduce in one round trip -- was measured and is NOT faster: scripts/barrier_bench3.cu, 3.55 vs 3.40 us; with
// gpu-scope "strong" 16-byte accesses it is 3x slower.)
template <int NUSED = NRED>
__device__ __forceinline__ void phase_end(const KParams &kp, BlockCtx &bc, double (&acc)[NRED]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int q = 0; q < NUSED; ++q) {
    double v = warp_sum(acc[q]);
    if (lane == 0) bc.sm_warp[warp * NRED + q] = v;
  }
  __syncthreads();                                       // also: every write of this phase is ordered before the release below
  bc.epoch += gridDim.x;
  double *slot = kp.partials + (size_t)bc.parity * kp.grid * NRED;
  double *outp = bc.sm_out + bc.parity * NRED;
  if (warp == 0) {
    if (NUSED > 0) {
      double s[NUSED > 0 ? NUSED : 1];
#pragma unroll
      for (int q = 0; q < NUSED; ++q) {
        double v = (lane < nwarps) ? bc.sm_warp[lane * NRED + q] : 0.0;
        s[q] = warp_sum(v);                              // fixed shuffle tree over the warps' partials
      }
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NUSED; ++q) slot[(size_t)blockIdx.x * NRED + q] = s[q];
      }
    }
    if (!kp.cluster) {
      if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(kp.bar_counter) : "memory");
      if (kp.strict_acquire) { while ((int)(ld_acquire_u32(kp.bar_counter) - bc.epoch) < 0) { } }
      else { while ((int)(ld_relaxed_u32(kp.bar_counter) - bc.epoch) < 0) { } }
    }
  }
  if (kp.cluster) {
    // the grid is one thread-block cluster: the hardware cluster barrier (every thread arrives; release at cluster scope
    // publishes the CTA's global writes to the other CTAs of the cluster, which read them from L2) replaces the
    // atomic counter and its polling round trips -- ~0.3 us instead of >= 1.3 us per phase end
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    if (kp.strict_acquire) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    else asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  }
  if (warp == 0) {
    if (NUSED > 0) {
      double t[NUSED > 0 ? NUSED : 1];
#pragma unroll
      for (int q = 0; q < NUSED; ++q) t[q] = 0.0;
      for (int c = lane; c < kp.grid; c += 32) {
#pragma unroll
        for (int q = 0; q < NUSED; ++q) t[q] += __ldcg(slot + (size_t)c * NRED + q);
      }
#pragma unroll
      for (int q = 0; q < NUSED; ++q) {
        const double v = warp_sum(t[q]);
        if (lane == 0) outp[q] = v;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NUSED; ++q) acc[q] = outp[q];
  bc.parity ^= 1;
}

// The CTA's row range and (when it fits) a shared-memory copy of its block-CSR structure, set up once per launch: the
// sparse phases then start their X gathers without first waiting for two dependent global loads (row pointer, indices).
struct CtaRows {
  int r0, r1;
  const int *rowptr;      // indexable by the global row j (shared copy: local block offsets; else the global array)
  const int *bcol;        // indexable by those block offsets
  const double *bval;     // rebased to match
};

template <int R> struct RowIter {
  static constexpr int SG = SubGroup<R>::SG;
  int r0, r1, stride, jb, sgw, a, c;
  __device__ RowIter(const CtaRows &cr) {
    r0 = cr.r0;
    r1 = cr.r1;
    const int lane = threadIdx.x & 31;
    constexpr int SGW = 32 / SG;                 // sub-groups per warp
    sgw = lane / SG;
    stride = (blockDim.x >> 5) * SGW;
    jb = r0 + (threadIdx.x >> 5) * SGW;
    const int l = lane & (SG - 1);
    a = l >> 2;
    c = l & 3;
  }
};

// ---------------------------------------------------------------------------------------------
// Phase E: everything that is needed at a base point in ONE pass over Q
//   EG = X Q + G, f = 0.5 <XQ, X> + <X, G>, S = sym(Y^T EG_Y), RG = P_X(EG), |RG|^2,
//   Z0 = P_X(M^-1 RG) for the pose-local preconditioners, <Z0, RG>.
// (ref: QuadraticProblem::f / EucGrad / RieGrad, src/QuadraticProblem.cpp:50-66,89-101; the reference
//  spends 5 separate X.Q products on these values per optimize() call.)
// acc: [0] <XQ,X>  [1] <X,G>  [2] |RG|^2  [3] <Z0,RG>
// ---------------------------------------------------------------------------------------------
template <int R, int DH>
__device__ void phase_eval(const KParams &kp, const CtaRows &cr, int cb, bool save_xin, int precond, double (&acc)[NRED]) {
  constexpr int TS = R * DH;
  const double *X = kp.v[V_X0 + cb];
  double *EG = kp.v[V_EG0 + cb], *RG = kp.v[V_RG0 + cb], *Z0 = kp.v[V_Z00 + cb], *S = kp.S[cb];
  RowIter<R> it(cr);
  const bool valid = (it.a < R) && (it.c < DH);
  const int e = it.c * R + it.a;
  for (int jb = it.jb; jb < it.r1; jb += it.stride) {
    const int j = jb + it.sgw;
    const bool act = (j < it.r1);
    const int js = act ? j : it.r1 - 1;
    const bool ld = act && valid;
    double xq = gather_tile<R, DH, true, false>(cr.rowptr, cr.bcol, cr.bval, X, nullptr, 0.0, js, it.a, it.c);
    const size_t idx = (size_t)js * TS + e;
    const double x = ld ? __ldcg(X + idx) : 0.0;
    const double g = ld ? __ldcg(kp.G + idx) : 0.0;
    if (!ld) xq = 0.0;
    const double eg = xq + g;
    acc[0] = fma(xq, x, acc[0]);
    acc[1] = fma(x, g, acc[1]);
    double ya[3], sym[3];
    const double rg = tangent_project_elem<R, DH>(x, eg, it.a, it.c, ya, sym);
    if (ld) {
      EG[idx] = eg;
      RG[idx] = rg;
      if (save_xin) kp.v[V_XIN][idx] = x;
      acc[2] = fma(rg, rg, acc[2]);
    }
    if (act && it.a == 0 && it.c < 3) {
      double *s = S + (size_t)js * 9 + it.c * 3;
      s[0] = sym[0]; s[1] = sym[1]; s[2] = sym[2];
    }
    if (precond != DPGO_PRECOND_DENSE_EXACT && precond != DPGO_PRECOND_SPARSE_EXACT) {
      double z0 = rg;
      if (precond == DPGO_PRECOND_BLOCK_JACOBI) {
        double t = jacobi_elem<R, DH>(kp.dinv, js, ld ? rg : 0.0, it.a, it.c);
        double ya2[3], sym2[3];
        z0 = tangent_project_elem<R, DH>(x, t, it.a, it.c, ya2, sym2);
      }
      if (ld) {
        Z0[idx] = z0;
        acc[3] = fma(z0, rg, acc[3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Phase H: Riemannian Hessian-vector product of the tCG direction (ref: QuadraticProblem::
// EucHessianEta, src/QuadraticProblem.cpp:68-73, + ROPTLIB Stiefel::EucHvToHv + projection):
//   delta_new = -zsrc + beta * delta_old      (formed on the fly for neighbour tiles, stored for own)
//   HD = P_X( delta_new Q - [delta_new_Y S]_pose ),   acc[0] = <delta_new, HD>
// If onfly == false the operand is read as is from `zsrc` (single-operation entry point).
// ---------------------------------------------------------------------------------------------
template <int R, int DH>
__device__ void phase_hess(const KParams &kp, const CtaRows &cr, int cb, const double *zsrc, const double *dold, double *dnew,
                           double beta, bool onfly, double (&acc)[NRED]) {
  constexpr int TS = R * DH;
  constexpr int D = DH - 1;
  const double *X = kp.v[V_X0 + cb];
  const double *S = kp.S[cb];
  double *HD = kp.v[V_HD];
  RowIter<R> it(cr);
  const bool valid = (it.a < R) && (it.c < DH);
  const int e = it.c * R + it.a;
  const bool ticking = (kp.phase_ns != nullptr) && blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long th0 = 0, th1 = 0;
  if (ticking) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(th0));
  for (int jb = it.jb; jb < it.r1; jb += it.stride) {
    const int j = jb + it.sgw;
    const bool act = (j < it.r1);
    const int js = act ? j : it.r1 - 1;
    const bool ld = act && valid;
    // the row's own operands first: these loads are independent of the gather and overlap its round trips
    const size_t idx = (size_t)js * TS + e;
    double dl = 0.0, dprev = 0.0;
    if (ld) {
      dl = __ldcg(zsrc + idx);
      if (onfly && beta != 0.0) dprev = __ldcg(dold + idx);
    }
    const double x = ld ? __ldcg(X + idx) : 0.0;
    const bool rot = ld && (it.c < D);
    double s0 = 0, s1 = 0, s2 = 0;
    if (rot) {
      const double *s = S + (size_t)js * 9 + it.c * 3;
      s0 = __ldcg(s); s1 = __ldcg(s + 1); s2 = __ldcg(s + 2);
    }
    double hq;
    if (onfly) hq = gather_tile<R, DH, true, true>(cr.rowptr, cr.bcol, cr.bval, zsrc, dold, beta, js, it.a, it.c);
    else hq = gather_tile<R, DH, true, false>(cr.rowptr, cr.bcol, cr.bval, zsrc, nullptr, 0.0, js, it.a, it.c);
    if (ticking && jb == it.jb) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(th1)); kp.phase_ns[27] += th1 - th0; }
    if (ld && onfly) {
      dl = -dl;
      if (beta != 0.0) dl = fma(beta, dprev, dl);
      dnew[idx] = dl;
    }
    const double dr = rot ? dl : 0.0;
    const double d0 = quad_get(dr, 0), d1 = quad_get(dr, 1), d2 = (D > 2) ? quad_get(dr, 2) : 0.0;
    double w = ld ? hq : 0.0;
    if (rot) w -= (d0 * s0 + d1 * s1 + d2 * s2);
    double ya[3], sym[3];
    const double hd = tangent_project_elem<R, DH>(x, w, it.a, it.c, ya, sym);
    if (ld) {
      HD[idx] = hd;
      acc[0] = fma(dl, hd, acc[0]);
    }
  }
  if (ticking) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(th1)); kp.phase_ns[28] += th1 - th0; }
}

// ---------------------------------------------------------------------------------------------
// Phase U: eta += alpha delta, res += alpha HD, |res|^2 and (pose-local preconditioners)
// z = P_X(M^-1 res), <z,res>.   acc: [0] |res|^2  [1] <z,res>
// ---------------------------------------------------------------------------------------------
template <int R, int DH>
__device__ void phase_update(const KParams &kp, const CtaRows &cr, int cb, const double *dcur, double alpha, bool first, int precond,
                             double (&acc)[NRED]) {
  constexpr int TS = R * DH;
  const double *X = kp.v[V_X0 + cb];
  const double *RG = kp.v[V_RG0 + cb];
  double *ETA = kp.v[V_ETA], *RES = kp.v[V_RES], *Z = kp.v[V_Z];
  const double *HD = kp.v[V_HD];
  RowIter<R> it(cr);
  const bool valid = (it.a < R) && (it.c < DH);
  const int e = it.c * R + it.a;
  for (int jb = it.jb; jb < it.r1; jb += it.stride) {
    const int j = jb + it.sgw;
    const bool act = (j < it.r1);
    const int js = act ? j : it.r1 - 1;
    const bool ld = act && valid;
    const size_t idx = (size_t)js * TS + e;
    double res = 0.0, x = 0.0;
    if (ld) {
      const double dl = __ldcg(dcur + idx), hd = __ldcg(HD + idx);
      const double eta0 = first ? 0.0 : __ldcg(ETA + idx);
      const double res0 = first ? __ldcg(RG + idx) : __ldcg(RES + idx);
      res = fma(alpha, hd, res0);
      ETA[idx] = fma(alpha, dl, eta0);
      RES[idx] = res;
      acc[0] = fma(res, res, acc[0]);
      x = __ldcg(X + idx);
    }
    if (precond != DPGO_PRECOND_DENSE_EXACT && precond != DPGO_PRECOND_SPARSE_EXACT) {
      double z = res;
      if (precond == DPGO_PRECOND_BLOCK_JACOBI) {
        double t = jacobi_elem<R, DH>(kp.dinv, js, res, it.a, it.c);
        double ya[3], sym[3];
        z = tangent_project_elem<R, DH>(x, t, it.a, it.c, ya, sym);
      }
      if (ld) {
        Z[idx] = z;
        acc[1] = fma(z, res, acc[1]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Dense-inverse preconditioner (parity mode; ref: QuadraticProblem::PreConditioner,
// src/QuadraticProblem.cpp:75-87 = CHOLMOD solve with Q+0.1I, then projection):
//   phase_dense : CTA b owns the row slab [k0,k1) of the symmetric Pinv (a contiguous 8*N*(k1-k0) byte
//                 stream, every load coalesced, up to `per` independent loads in flight per thread) and
//                 writes its partial product  part_b = V[:, k0:k1] * Pinv[k0:k1, :]   (r x N)
//   phase_pz    : T = sum_b part_b in fixed CTA order (deterministic), Z = P_X(T), acc[0] = <Z, V>
// ---------------------------------------------------------------------------------------------
template <int R> __device__ void phase_dense(const KParams &kp, const double *V, double *sV) {
  const int N = kp.N;
  const int per = kp.dense_per;
  const int k0 = min(N, (int)blockIdx.x * per), k1 = min(N, k0 + per);
  const int nk = k1 - k0;
  if (nk <= 0) return;
  for (int q = threadIdx.x; q < nk * R; q += blockDim.x) sV[q] = __ldcg(V + (size_t)k0 * R + q);
  __syncthreads();
  double *part = kp.dense_part + (size_t)blockIdx.x * R * N;
  const double *P = kp.pinv + (size_t)k0 * N;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    double acc[R];
#pragma unroll
    for (int a = 0; a < R; ++a) acc[a] = 0.0;
    int kk = 0;
    for (; kk + 16 <= nk; kk += 16) {
      double p[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) p[u] = ld_stream(P + (size_t)(kk + u) * N + c);
#pragma unroll
      for (int u = 0; u < 16; ++u)
#pragma unroll
        for (int a = 0; a < R; ++a) acc[a] = fma(sV[(kk + u) * R + a], p[u], acc[a]);
    }
    for (; kk + 4 <= nk; kk += 4) {
      double p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) p[u] = ld_stream(P + (size_t)(kk + u) * N + c);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int a = 0; a < R; ++a) acc[a] = fma(sV[(kk + u) * R + a], p[u], acc[a]);
    }
    for (; kk < nk; ++kk) {
      const double p = ld_stream(P + (size_t)kk * N + c);
#pragma unroll
      for (int a = 0; a < R; ++a) acc[a] = fma(sV[kk * R + a], p, acc[a]);
    }
#pragma unroll
    for (int a = 0; a < R; ++a) part[(size_t)c * R + a] = acc[a];
  }
}

// ---- TMA-fed variant of phase_dense -----------------------------------------------------------------------
// Warp 15 is the producer: it streams the CTA's row slab of Pinv through a DENSE_NST-deep shared-memory ring with
// 1-D bulk TMA copies (one chunk = one row x DENSE_SEG columns = 15 KB, contiguous in HBM); warps 0..14 consume
// (thread t owns columns t + 480 m, m < 4, of the current column segment).  ~90 KB in flight per SM, no register
// staging, so the stream runs close to the copy roofline instead of being limited by loads in flight per thread.
constexpr int DENSE_NST = 6;
constexpr int DENSE_CONS = OPT_THREADS - 32;          // 480 consumer threads
constexpr int DENSE_SEG = DENSE_CONS * 4;             // 1920 columns per chunk
constexpr int DENSE_RING_DOUBLES = 6 * (8 * ((OPT_THREADS / 32 - 1) * 32 + 4) + 8 * 5);   // max(6 x 1920, the symmetric variant's 6 stages) doubles
static_assert(DENSE_RING_DOUBLES >= DENSE_NST * DENSE_SEG, "ring must hold the full-matrix stages too");

struct DenseRing {
  double *buf;          // DENSE_NST * DENSE_SEG doubles
  uint64_t *full;       // DENSE_NST
  uint64_t *empty;      // DENSE_NST
  unsigned count;       // chunks issued / consumed so far in this kernel (same in every thread)
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait_parity(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DW_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DW_DONE;\n"
      "bra DW_LOOP;\n"
      "DW_DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void mbar_wait_parity_s(uint32_t bar_s, unsigned parity) {   // same, 32-bit shared address
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DWS_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DWS_DONE;\n"
      "bra DWS_LOOP;\n"
      "DWS_DONE:\n"
      "}\n" ::"r"(bar_s),
      "r"(parity)
      : "memory");
}

template <int R> __device__ void phase_dense_tma(const KParams &kp, const double *V, double *sV, DenseRing &ring) {
  const int N = kp.N;
  const int per = kp.dense_per;
  const int k0 = min(N, (int)blockIdx.x * per), k1 = min(N, k0 + per);
  const int nk = k1 - k0;
  const int nseg = (N + DENSE_SEG - 1) / DENSE_SEG;
  const unsigned total = (unsigned)(nk > 0 ? nk * nseg : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (nk > 0) {
    for (int q = threadIdx.x; q < nk * R; q += blockDim.x) sV[q] = __ldcg(V + (size_t)k0 * R + q);
  }
  __syncthreads();
  if (nk > 0) {
    if (warp == (OPT_THREADS / 32 - 1)) {
      // ---------------- producer ----------------
      if (lane == 0) {
        // evict-first: the 8 N^2 byte stream must not push the partial products / work vectors out of L2
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        unsigned c = ring.count;
        for (int sg = 0; sg < nseg; ++sg) {
          const int c0 = sg * DENSE_SEG;
          const unsigned bytes = (unsigned)min(DENSE_SEG, N - c0) * 8u;
          for (int kk = 0; kk < nk; ++kk, ++c) {
            const int st = c % DENSE_NST;
            if (c >= DENSE_NST) mbar_wait_parity(&ring.empty[st], ((c / DENSE_NST) - 1) & 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&ring.full[st])), "r"(bytes)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                             smem_addr(ring.buf + (size_t)st * DENSE_SEG)),
                         "l"(kp.pinv + (size_t)(k0 + kk) * N + c0), "r"(bytes), "r"(smem_addr(&ring.full[st])), "l"(pol)
                         : "memory");
          }
        }
      }
    } else {
      // ---------------- consumers ----------------
      double *part = kp.dense_part + (size_t)blockIdx.x * R * N;
      unsigned c = ring.count;
      for (int sg = 0; sg < nseg; ++sg) {
        const int c0 = sg * DENSE_SEG;
        double acc[4][R];
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int a = 0; a < R; ++a) acc[m][a] = 0.0;
        for (int kk = 0; kk < nk; ++kk, ++c) {
          const int st = c % DENSE_NST;
          mbar_wait_parity(&ring.full[st], (c / DENSE_NST) & 1);
          const double *sp = ring.buf + (size_t)st * DENSE_SEG + threadIdx.x;
          double v[R];
#pragma unroll
          for (int a = 0; a < R; ++a) v[a] = sV[kk * R + a];
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const double pv = sp[m * DENSE_CONS];            // columns past N hold stale data, never stored
#pragma unroll
            for (int a = 0; a < R; ++a) acc[m][a] = fma(v[a], pv, acc[m][a]);
          }
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&ring.empty[st])) : "memory");
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int col = c0 + (int)threadIdx.x + m * DENSE_CONS;
          if (col < N) {
#pragma unroll
            for (int a = 0; a < R; ++a) part[(size_t)col * R + a] = acc[m][a];
          }
        }
      }
    }
  }
  ring.count += total;
}

// ---- symmetric variant: read only the upper triangle of Pinv --------------------------------------------------
// T = V * Pinv with Pinv symmetric.  The matrix is cut into column segments of SYM_SEG = 480 columns and row groups
// of 8 rows; a chunk (J, g) is the 8 x <=480 piece of group g inside segment J, restricted to columns >= 8g (upper
// trapezoid: half the bytes).  Every 8x8 tile right of the diagonal feeds TWO products from shared memory:
//   direct      T[:, ctile..] += V[:, g0..]   * P[g0.., ctile..]      (per column; registers, summed over the groups)
//   transposed  T[:, g0..]    += V[:, ctile..] * P[g0.., ctile..]^T    (per chunk; summed over the 15 consumer warps)
// and the diagonal tile feeds the direct product only.  Both are mma.sync m8n8k4 (DMMA): the 8x8 tile is read from
// shared memory in the two B-fragment layouts, so the k-reduction of the transposed product happens inside the tensor
// op.  The chunks are ordered segment-major and the host cuts that sequence into `grid` contiguous runs of equal cost
// (a chunk costs about the same whatever its width, see ensure_dense), so a CTA works inside one or two segments:
//   * direct partials: one panel slot per (CTA, segment) -- ~grid/nseg+1 slots per column instead of `grid`;
//   * transposed results: chunk (J, g) is owned by exactly one CTA, which writes rows 8g..8g+7 of slot J of dense_t2.
// phase_pz adds, per element, the <= ccount[J] panel slots and the <= nseg transposed slots in fixed order.
// The upper triangle is stored chunk-major (ppack: a chunk's 8 rows x (width rounded up to 8, + 4 pad) doubles are
// contiguous, chunks in processing order), so a CTA's whole run is ONE contiguous byte range of HBM.
// Warp 15 is the producer (bulk TMA: one copy per chunk + the 8 V columns of the group, L2 evict-first for the
// matrix), warps 0..14 consume through full/empty mbarriers; the consumers park their transposed fragments in shared
// memory and meet at a named barrier once per SYM_WIN/2 chunks; the sum over the warps overlaps the next chunks.
constexpr int SYM_NST = 6;
constexpr int SYM_SEG = (OPT_THREADS / 32 - 1) * 32;   // 480 columns per segment: 32 per consumer warp
constexpr int SYM_SROW = SYM_SEG + 4;                   // staged row pitch: 3872 B = 32 mod 128 (conflict-light tile reads)
constexpr int SYM_VOFF = 8 * SYM_SROW;                  // the group's V columns (8 x R doubles) follow the 8 rows
constexpr int SYM_STAGE = SYM_VOFF + 8 * 5;             // doubles per stage (R <= 5)
constexpr int SYM_RING_DOUBLES = SYM_NST * SYM_STAGE;
static_assert(SYM_RING_DOUBLES <= DENSE_RING_DOUBLES && SYM_NST <= DENSE_NST, "ring too small for the symmetric variant");
constexpr int SYM_WIN = 4;                              // chunks per transposed-fragment window
constexpr int SYM_META_DOUBLES = SYM_WIN;               // 2 ints per window slot: segment, first row
static_assert((OPT_THREADS / 32 - 1) * SYM_WIN * 8 * 5 <= DENSE_PER_MAX * 5, "the fragment window lives in the V staging area");

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int R> __device__ void phase_dense_sym(const KParams &kp, const double *V, DenseRing &ring, double *sAcc, int *sMeta) {
  const int N = kp.N;
  const int nseg = (N + SYM_SEG - 1) / SYM_SEG;
  const int lin0 = ld_const(kp.sym_cut + blockIdx.x), lin1 = ld_const(kp.sym_cut + blockIdx.x + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NCW = OPT_THREADS / 32 - 1;             // consumer warps
  int J0 = 0;
  while (J0 + 1 < nseg && ld_const(kp.sym_segptr + J0 + 1) <= lin0) ++J0;
  if (warp == NCW) {
    // ---------------- producer ----------------
    if (lane == 0) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      unsigned c = ring.count;
      int lin = lin0;
      for (int J = J0; J < nseg && lin < lin1; ++J) {
        const int p0 = ld_const(kp.sym_segptr + J), p1 = ld_const(kp.sym_segptr + J + 1);
        const int gb = min(p1, lin1) - p0;
        for (int g = lin - p0; g < gb; ++g) {
          const int g0 = 8 * g;
          const int nrows = min(8, N - g0);
          const unsigned vbytes = (unsigned)(nrows * R) * 8u;
          const long long o0 = __ldg(kp.sym_off + p0 + g), o1 = __ldg(kp.sym_off + p0 + g + 1);
          const unsigned cbytes = (unsigned)(o1 - o0) * 8u;
          const int st = c % SYM_NST;
          if (c >= SYM_NST) mbar_wait_parity(&ring.empty[st], ((c / SYM_NST) - 1) & 1);
          double *stage = ring.buf + (size_t)st * SYM_STAGE;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&ring.full[st])), "r"(cbytes + vbytes)
                       : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_addr(stage + SYM_VOFF)),
                       "l"(V + (size_t)g0 * R), "r"(vbytes), "r"(smem_addr(&ring.full[st]))
                       : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                           smem_addr(stage)),
                       "l"(kp.ppack + o0), "r"(cbytes), "r"(smem_addr(&ring.full[st])), "l"(pol)
                       : "memory");
          ++c;
        }
        lin = p0 + gb;
      }
    }
  } else {
    // ---------------- consumers ----------------
    const int a = lane >> 2, k = lane & 3;                // A fragment: row a, column k ; D fragment: row a, columns 2k, 2k+1
    const int bn = lane >> 2, bk = lane & 3;              // B fragment: column n = lane>>2, row k = lane&3
    const int ac = min(a, R - 1);                         // rows a >= R of the fragments are never stored: any finite operand will do
    // lane-constant offsets (doubles) of the interior path: full-width chunk, every tile right of the diagonal
    const int offD = bk * SYM_SROW + 32 * warp + bn;      // direct      B[kk][n] = P[g0 + kk][ctile + n]
    const int offT = bn * SYM_SROW + 32 * warp + bk;      // transposed  B[c'][n] = P[g0 + n][ctile + c']
    const int offV = SYM_VOFF + k * R + ac;               // V[a, g0 + k]
    double *fragLane = sAcc + (size_t)(warp * SYM_WIN * 8 + 2 * k) * R + ac;
    const uint32_t full_s = smem_addr(ring.full), empty_s = smem_addr(ring.empty);
    int st = (int)(ring.count % SYM_NST);
    unsigned ph = (ring.count / SYM_NST) & 1u;
    int cidx = 0;                                         // chunks done in this phase
    int lin = lin0;
    for (int J = J0; J < nseg && lin < lin1; ++J) {
      const int s0 = J * SYM_SEG, s1 = min(N, s0 + SYM_SEG);
      const int p0 = ld_const(kp.sym_segptr + J), p1 = ld_const(kp.sym_segptr + J + 1);
      const int gb = min(p1, lin1) - p0;
      const int cw0 = s0 + 32 * warp;
      const bool fullseg = (s1 - s0 == SYM_SEG);
      double va[4][2];                                    // transposed-product A operands: V[a, ctile + 4q + k]
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int col = cw0 + 8 * t + 4 * q + k;
          va[t][q] = (a < R && col < N) ? __ldcg(V + (size_t)col * R + a) : 0.0;
        }
      double D1[4][2];
#pragma unroll
      for (int t = 0; t < 4; ++t) D1[t][0] = D1[t][1] = 0.0;
      for (int g = lin - p0; g < gb; ++g) {
        const int g0 = 8 * g;
        double D2_0 = 0.0, D2_1 = 0.0;
        mbar_wait_parity_s(full_s + 8u * (unsigned)st, ph);
        const double *base = ring.buf + (size_t)st * SYM_STAGE;
        if (fullseg && g0 + 8 <= s0) {
          // interior chunk (>90 % of the bytes): no masks, constant pitch
          const double a1_0 = base[offV], a1_1 = base[offV + 4 * R];
          const double *bD = base + offD, *bT = base + offT;
          double E2_0 = 0.0, E2_1 = 0.0;                       // second transposed chain: halves the dependent DMMA depth
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            dmma884(D1[t][0], D1[t][1], a1_0, bD[8 * t]);
            dmma884(D1[t][0], D1[t][1], a1_1, bD[4 * SYM_SROW + 8 * t]);
            if (t & 1) {
              dmma884(E2_0, E2_1, va[t][0], bT[8 * t]);
              dmma884(E2_0, E2_1, va[t][1], bT[8 * t + 4]);
            } else {
              dmma884(D2_0, D2_1, va[t][0], bT[8 * t]);
              dmma884(D2_0, D2_1, va[t][1], bT[8 * t + 4]);
            }
          }
          D2_0 += E2_0;
          D2_1 += E2_1;
        } else {
          // diagonal / ragged chunk: tiles left of the diagonal or past the segment end are skipped (warp-uniform)
          const int col_lo = max(s0, g0);
          const int pitch = ((s1 - col_lo + 7) & ~7) + 4;     // chunk row pitch in ppack (= 32 or 96 bytes mod 128)
          const double a1_0 = (a < R && g0 + k < N) ? base[SYM_VOFF + k * R + a] : 0.0;            // V[a, g0 + k]
          const double a1_1 = (a < R && g0 + 4 + k < N) ? base[SYM_VOFF + (4 + k) * R + a] : 0.0;  // V[a, g0 + 4 + k]
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int ctile = cw0 + 8 * t;
            if (ctile < g0 || ctile >= s1) continue;
            const int x = ctile - col_lo;
            dmma884(D1[t][0], D1[t][1], a1_0, base[bk * pitch + x + bn]);
            dmma884(D1[t][0], D1[t][1], a1_1, base[(4 + bk) * pitch + x + bn]);
            if (ctile > g0) {
              dmma884(D2_0, D2_1, va[t][0], base[bn * pitch + x + bk]);
              dmma884(D2_0, D2_1, va[t][1], base[bn * pitch + x + 4 + bk]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_s + 8u * (unsigned)st) : "memory");
        if (++st == SYM_NST) { st = 0; ph ^= 1u; }
        // transposed result of this chunk: rows g0 + 2k, g0 + 2k + 1 (fragment columns), component a.  The fragments of
        // SYM_WIN consecutive chunks are parked per warp in shared memory (two halves used alternately).
        const int slot = cidx & (SYM_WIN - 1);
        if (a < R) {
          fragLane[slot * 8 * R] = D2_0;
          fragLane[slot * 8 * R + R] = D2_1;
        }
        if (threadIdx.x == 0) { sMeta[2 * slot] = J; sMeta[2 * slot + 1] = g0; }
        ++cidx;
        if ((slot & (SYM_WIN / 2 - 1)) == SYM_WIN / 2 - 1 || p0 + g + 1 == lin1) {
          // one half of the window is complete: the warps meet, then up to 160 threads sum it while the others go on
          // filling the other half.  One barrier per half window: a half is rewritten only after the NEXT barrier, which
          // the summing threads reach after they are done with it.
          __syncwarp();
          asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
          const int h0 = slot & (SYM_WIN / 2), nsl = (slot & (SYM_WIN / 2 - 1)) + 1;
          for (int q = threadIdx.x; q < nsl * 8 * R; q += NCW * 32) {
            const int sl = h0 + q / (8 * R), rem = q % (8 * R);
            const int Jm = sMeta[2 * sl], g0m = sMeta[2 * sl + 1];
            if (g0m + rem / R < N) {
              double sum = 0.0;
#pragma unroll
              for (int w = 0; w < NCW; ++w) sum += sAcc[(size_t)((w * SYM_WIN + sl) * 8) * R + rem];
              kp.dense_t2[(size_t)Jm * R * N + (size_t)g0m * R + rem] = sum;
            }
          }
        }
      }
      lin = p0 + gb;
      // direct partials of this (CTA, segment) run: every column of the segment, zeros included
      if (a < R) {
        double *part = kp.dense_part + (size_t)(blockIdx.x - ld_const(kp.sym_cfirst + J)) * R * N;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int col = cw0 + 8 * t + 2 * k;
          if (col < N) part[(size_t)col * R + a] = D1[t][0];
          if (col + 1 < N) part[(size_t)(col + 1) * R + a] = D1[t][1];
        }
      }
    }
  }
  ring.count += (unsigned)(lin1 - lin0);
  __syncthreads();
}

// Sum of the partial panels + tangent projection.  Step 1 is a flat job over the CTA's contiguous element range
// [r0*TS, r1*TS): thread (e, p) sums slabs p, p+P, ... of element e with 32 independent L2 loads in flight (P = how
// many times the range fits into the CTA), fixed order; step 2 combines the P parts in order and projects per pose.
template <int R, int DH>
__device__ void phase_pz(const KParams &kp, const CtaRows &cr, int cb, const double *V, double *Zout, double *sT, double (&acc)[NRED]) {
  constexpr int TS = R * DH;
  constexpr int PZ_TILE = DENSE_PER_MAX * R;          // elements staged per round (sT = the V staging area of the dense phases)
  const int nseg_sym = (kp.N + SYM_SEG - 1) / SYM_SEG;
  const double *X = kp.v[V_X0 + cb];
  const bool symm = kp.sym_ok != 0;
  const int nslabs_full = (kp.N + kp.dense_per - 1) / kp.dense_per;
  const size_t stride = (size_t)R * kp.N;
  RowIter<R> it(cr);
  const bool valid = (it.a < R) && (it.c < DH);
  const int e = it.c * R + it.a;
  const int rows_per_round = PZ_TILE / TS;
  for (int rb = it.r0; rb < it.r1; rb += rows_per_round) {
    const int re = min(it.r1, rb + rows_per_round);
    const int E = (re - rb) * TS;                                   // flat elements of this round
    const int P = max(1, min(4, (int)blockDim.x / max(E, 1)));      // slab parts per element (E * P <= PZ_TILE)
    const size_t base = (size_t)rb * TS;
    __syncthreads();                                                 // sT free (previous round consumed)
    for (int q = threadIdx.x; q < E * P; q += blockDim.x) {
      const int part = q / E, el = q - part * E;
      const size_t idx = base + el;
      const int col = (int)(idx / R);
      // symmetric variant: ccount[J] direct panels of the column's segment J, then the transposed slots of the
      // segments J(row group) .. nseg-1; one virtual list so that the parts split it evenly
      const int J = col / SYM_SEG;
      const int ndir = symm ? ld_const(kp.sym_ccount + J) : nslabs_full;
      const int jt0 = (col & ~7) / SYM_SEG;
      const int nslabs = symm ? ndir + (nseg_sym - jt0) : nslabs_full;
      const int b0 = (int)((long long)nslabs * part / P), b1 = (int)((long long)nslabs * (part + 1) / P);
      const double *pp = kp.dense_part + idx;
      // slot b >= ndir is segment jt0 + b - ndir of the transposed partials
      const double *pt = symm ? kp.dense_t2 + idx + (ptrdiff_t)(jt0 - ndir) * (ptrdiff_t)stride : pp;
      double t = 0.0;
      for (int b = b0; b < b1; b += 32) {                                     // 32 independent L2 loads in flight, summed in order
        double tt[32];
#pragma unroll
        for (int u = 0; u < 32; ++u)
          tt[u] = (b + u < b1) ? __ldcg(((b + u) < ndir ? pp : pt) + (size_t)(b + u) * stride) : 0.0;
#pragma unroll
        for (int u = 0; u < 32; ++u) t += tt[u];
      }
      sT[q] = t;
    }
    __syncthreads();
    for (int jb = rb + (it.jb - it.r0); jb < re; jb += it.stride) {
      const int j = jb + it.sgw;
      const bool act = (j < re);
      const int js = act ? j : re - 1;
      const bool ld = act && valid;
      const size_t idx = (size_t)js * TS + e;
      const double x = ld ? __ldcg(X + idx) : 0.0;
      double t = 0.0;
      if (ld) {
        const int el = (js - rb) * TS + e;
        for (int part = 0; part < P; ++part) t += sT[part * E + el];
      }
      double ya[3], sym[3];
      const double z = tangent_project_elem<R, DH>(x, t, it.a, it.c, ya, sym);
      if (ld) {
        Zout[idx] = z;
        acc[0] = fma(z, __ldcg(V + idx), acc[0]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Sparse exact preconditioner (ref: QuadraticProblem::PreConditioner, src/QuadraticProblem.cpp:75-87 = CHOLMOD solve
// with Q + 0.1 I, then projection).  One phase of the nested-dissection block solve (nd_precond.h): the CTA walks its
// steps of the host-built plan:
//   gathers   : tiles of the phase's input vector -> shared memory (forward: residual minus the children's
//               contributions; backward: the ancestors' solution), one sub-group per tile
//   jobs      : one warp = one 8-row panel x a column piece; lane (row = lane & 7, cp = lane >> 3) reads columns
//               cp, cp + 4, ... : 256 contiguous bytes per warp load; r accumulators per lane; 2 shuffle steps
//   epilogues : one sub-group per pose: sums the panel's partial slots in fixed order, writes t / contribution / x;
//               solution tiles are projected onto the tangent space at X and dotted with V on the fly.
// acc[0] accumulates <Z, V> over the phases of one application.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int4 ld_int4(const void *p) { return __ldg(reinterpret_cast<const int4 *>(p)); }

template <int R, int DH>
__device__ void phase_nd(const KParams &kp, int ph, const double *V, int cb, double *Zout, double *ys, double *slots,
                         int4 *grec, double (&acc)[NRED]) {
  constexpr int TS = R * DH;
  constexpr int SG = SubGroup<R>::SG;
  constexpr int SGW = 32 / SG;
  constexpr int SLOT = nd::PANEL_ROWS * R;
  constexpr int NC = nd::INLINE_CONTRIB;
  const KNd &N = kp.nd;
  const int dir = N.dir[ph];
  const int4 ca = ld_int4(N.cta_phase + N.cta0[ph] + blockIdx.x),
             cbq = ld_int4(reinterpret_cast<const int4 *>(N.cta_phase + N.cta0[ph] + blockIdx.x) + 1);
  const int s0 = ca.x, s1 = ca.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sgw = lane / SG, l = lane & (SG - 1), a = l >> 2, c = l & 3;
  const bool valid = (a < R) && (c < DH);
  const int e = c * R + a;
  const int nsub = nwarps * SGW;
  const double *X = kp.v[V_X0 + cb];
  const double *src = (dir == 0) ? V : N.TX;
  const bool ticking = (kp.phase_ns != nullptr) && blockIdx.x == 0 && threadIdx.x == 0;
  for (int si = s0; si < s1; ++si) {
    int g0 = ca.z, g1 = ca.w, j0 = cbq.x, j1 = cbq.y, e0 = cbq.z, e1 = cbq.w;      // the first step sits in the CTA record
    if (si != s0) {
      const int4 sa = ld_int4(N.steps + si), sb = ld_int4(reinterpret_cast<const int4 *>(N.steps + si) + 1);
      g0 = sa.x; g1 = sa.y; j0 = sa.z; j1 = sa.w; e0 = sb.x; e1 = sb.y;
    }
    unsigned long long tg0 = 0, tg1 = 0, tg2 = 0;
    if (ticking) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tg0));
    // ---- gathers: the step's records are staged in shared memory with one coalesced sweep (one L2 round trip), then
    //      the loop is flat over (tile, element) with 4 independent elements per thread and round ----
    {
      const int ng = g1 - g0;
      const int4 *gsrc = reinterpret_cast<const int4 *>(N.gathers + g0);
      for (int q = threadIdx.x; q < 2 * ng; q += blockDim.x) grec[q] = __ldg(gsrc + q);
      __syncthreads();
      constexpr int GU = 4;
      const int total = ng * TS, nthr = blockDim.x;
      for (int base = threadIdx.x; base < total; base += GU * nthr) {
        int t[GU], el[GU];
        bool ok[GU];
        double v[GU], cv[GU][NC];
#pragma unroll
        for (int u = 0; u < GU; ++u) {
          const int idx = base + u * nthr;
          ok[u] = idx < total;
          t[u] = ok[u] ? idx / TS : 0;
          el[u] = idx - t[u] * TS;
          const int4 ra = grec[2 * t[u]], rb = grec[2 * t[u] + 1];             // (ytile, src, nc, cext), first contribution tiles
          v[u] = ok[u] ? __ldcg(src + (size_t)ra.y * TS + el[u]) : 0.0;
          const int ci[NC] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
          for (int q = 0; q < NC; ++q) cv[u][q] = (ok[u] && q < ra.z) ? __ldcg(N.C + (size_t)ci[q] * TS + el[u]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < GU; ++u) {
          const int4 ra = grec[2 * t[u]];
#pragma unroll
          for (int q = 0; q < NC; ++q) v[u] -= cv[u][q];
          for (int k = NC; k < ra.z && ok[u]; ++k)
            v[u] -= __ldcg(N.C + (size_t)ld_const(N.csrc + ra.w + k - NC) * TS + el[u]);
          if (ok[u]) ys[(size_t)ra.x * TS + el[u]] = v[u];
        }
      }
    }
    __syncthreads();
    if (ticking) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tg1));
    // ---- jobs: DMMA m8n8k4 -- A = 8 panel rows x 4 columns (one coalesced 256-byte warp load), B = 4 columns x r
    //      right-hand sides from shared memory, D = 8 x 8 accumulators (r columns used): 1 LDG + 1 LDS + 1 DMMA per
    //      4 columns instead of 1 LDG + r LDS + r DFMA per lane, and no shuffle reduction.  Two accumulator chains;
    //      the next job's record is fetched ahead.
    {
      const int r8 = lane >> 2, k4 = lane & 3;
      const bool bval_lane = (r8 < R);
      int ji = j0 + warp;
      int4 ja = make_int4(0, 0, 0, 0), jb = make_int4(0, 0, 0, 0);
      if (ji < j1) { ja = ld_int4(N.jobs + ji); jb = ld_int4(reinterpret_cast<const int4 *>(N.jobs + ji) + 1); }
      while (ji < j1) {
        const int jn = ji + nwarps;
        int4 na = make_int4(0, 0, 0, 0), nb = make_int4(0, 0, 0, 0);
        if (jn < j1) { na = ld_int4(N.jobs + jn); nb = ld_int4(reinterpret_cast<const int4 *>(N.jobs + jn) + 1); }
        const long long mat = ((long long)(unsigned)ja.x) | ((long long)ja.y << 32);
        const int ncols = ja.z, ycol = ja.w, slot = jb.x, accum = jb.y;
        const double *mp = N.blob + mat + (size_t)k4 * nd::PANEL_ROWS + r8;      // A[r8][k4] of the first column group
        const double *yp = ys + (size_t)(ycol + k4) * R + (bval_lane ? r8 : 0);    // B[k4][r8]
        double d0 = 0.0, d1 = 0.0, f0 = 0.0, f1 = 0.0;
        for (int j = 0; j < ncols; j += 32) {                                        // 8 column groups (32 columns) per round
          double am[8], bm[8];
          if (j + 32 <= ncols) {
#pragma unroll
            for (int u = 0; u < 8; ++u) am[u] = ld_stream(mp + (size_t)(j + 4 * u) * nd::PANEL_ROWS);
#pragma unroll
            for (int u = 0; u < 8; ++u) bm[u] = bval_lane ? yp[(size_t)(j + 4 * u) * R] : 0.0;
          } else {                                                                   // last, ragged round: same 8 independent loads, predicated
#pragma unroll
            for (int u = 0; u < 8; ++u) am[u] = (j + 4 * u + k4 < ncols) ? ld_stream(mp + (size_t)(j + 4 * u) * nd::PANEL_ROWS) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) bm[u] = (bval_lane && j + 4 * u + k4 < ncols) ? yp[(size_t)(j + 4 * u) * R] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
            dmma884(d0, d1, am[u], bm[u]);
            dmma884(f0, f1, am[u + 1], bm[u + 1]);
          }
        }
        d0 += f0;
        d1 += f1;
        {
          double *sl = slots + (size_t)slot * SLOT + r8 * R + 2 * k4;              // D[r8][2 k4], D[r8][2 k4 + 1]
          if (2 * k4 < R) sl[0] = accum ? sl[0] + d0 : d0;
          if (2 * k4 + 1 < R) sl[1] = accum ? sl[1] + d1 : d1;
        }
        ji = jn; ja = na; jb = nb;
      }
    }
    __syncthreads();
    if (ticking) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tg2));
    // ---- epilogues (warp-uniform trip count: the projection uses full-warp shuffles); global loads first, the next
    //      item's record is fetched ahead ----
    {
      int eb = e0 + warp * SGW;
      int4 ea = make_int4(-1, 0, 0, 0), ec = make_int4(0, 0, 0, 0), ed = make_int4(0, 0, 0, 0);
      if (eb + sgw < e1) {
        const int4 *rp = reinterpret_cast<const int4 *>(N.epis + eb + sgw);
        ea = __ldg(rp); ec = __ldg(rp + 1); ed = __ldg(rp + 2);
      }
      while (eb < e1) {
        const int en = eb + nsub;
        int4 fa = make_int4(-1, 0, 0, 0), fc = make_int4(0, 0, 0, 0), fd = make_int4(0, 0, 0, 0);
        if (en + sgw < e1) {
          const int4 *rp = reinterpret_cast<const int4 *>(N.epis + en + sgw);
          fa = __ldg(rp); fc = __ldg(rp + 1); fd = __ldg(rp + 2);
        }
        const bool act = (eb + sgw < e1);
        const int kind = ea.x, slot0 = ea.y, nslots = ea.z, half = ea.w, out = ec.x, aux = ec.y, nc = ec.z, cext = ec.w;
        const bool ld = act && valid;
        const bool sol = (kind == nd::EPI_ROOT) || (kind == nd::EPI_B_OWN);
        // independent global loads: t (backward), X and V tiles (solution poses), contributions (boundary rows)
        double tval = 0.0, xq = 0.0, vv = 0.0, cv[NC] = {0.0, 0.0, 0.0, 0.0};
        if (ld) {
          if (kind == nd::EPI_B_OWN) tval = __ldcg(N.TX + (size_t)out * TS + e);
          if (sol) { xq = __ldcg(X + (size_t)aux * TS + e); vv = __ldcg(V + (size_t)aux * TS + e); }
          if (kind == nd::EPI_F_BND) {
            const int ci[NC] = {ed.x, ed.y, ed.z, ed.w};
#pragma unroll
            for (int q = 0; q < NC; ++q) cv[q] = (q < nc) ? __ldcg(N.C + (size_t)ci[q] * TS + e) : 0.0;
          }
        }
        double sum = 0.0;
        if (ld) {
          const double *sl = slots + (size_t)slot0 * SLOT + (half * DH + c) * R + a;
          for (int k = 0; k < nslots; ++k) sum += sl[(size_t)k * SLOT];
        }
        double x = 0.0;
        if (ld) {
          if (kind == nd::EPI_F_OWN) {
            N.TX[(size_t)out * TS + e] = sum;
          } else if (kind == nd::EPI_F_BND) {
#pragma unroll
            for (int q = 0; q < NC; ++q) sum += cv[q];
            for (int k = NC; k < nc; ++k) sum += __ldcg(N.C + (size_t)ld_const(N.csrc + cext + k - NC) * TS + e);
            N.C[(size_t)out * TS + e] = sum;
          } else if (sol) {
            x = (kind == nd::EPI_ROOT) ? sum : tval - sum;
            N.TX[(size_t)out * TS + e] = x;
          }
        }
        if (__any_sync(FULL, sol)) {                      // forward sweeps below the root have nothing to project
          double ya[3], sym[3];
          const double z = tangent_project_elem<R, DH>(sol ? xq : 0.0, x, a, c, ya, sym);
          if (ld && sol) {
            Zout[(size_t)aux * TS + e] = z;
            acc[0] = fma(z, vv, acc[0]);
          }
        }
        eb = en; ea = fa; ec = fc; ed = fd;
      }
    }
    if (ticking) {
      unsigned long long tg3;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tg3));
      kp.phase_ns[24] += tg1 - tg0; kp.phase_ns[25] += tg2 - tg1; kp.phase_ns[26] += tg3 - tg2;
      unsigned long long *pp = kp.phase_ns + 32 + 3 * min(ph, 9);
      pp[0] += tg1 - tg0; pp[1] += tg2 - tg1; pp[2] += tg3 - tg2;
    }
    // the next step's gathers / jobs rewrite ys / slots only after every warp is past its epilogues
    if (si + 1 < s1) __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Phase RT: candidate X' = R_X(eta_final) with eta_final = eta + tau * delta (tau = 0 unless tCG
// stopped on the trust-region boundary / negative curvature), and the model-decrease dots
//   acc[0] = <eta, g>, acc[1] = <eta, H eta> with H eta = (res - g) + tau * HD (tCG recurrences).
// mode 1 (RGD, ref src/QuadraticOptimizer.cpp:124-149): eta = -step * RG.
// mode 2: eta read as is from V_AUX (single-operation entry point).
// ---------------------------------------------------------------------------------------------
template <int R, int DH>
__device__ void phase_retract(const KParams &kp, const CtaRows &cr, int cb, int mode, const double *dcur, double tau, bool eta_zero,
                              double step, double (&acc)[NRED]) {
  constexpr int TS = R * DH;
  const double *X = kp.v[V_X0 + cb];
  const double *RG = kp.v[V_RG0 + cb];
  double *X2 = kp.v[V_X0 + (1 - cb)];
  RowIter<R> it(cr);
  const bool valid = (it.a < R) && (it.c < DH);
  const int e = it.c * R + it.a;
  for (int jb = it.jb; jb < it.r1; jb += it.stride) {
    const int j = jb + it.sgw;
    const bool act = (j < it.r1);
    const int js = act ? j : it.r1 - 1;
    const bool ld = act && valid;
    const size_t idx = (size_t)js * TS + e;
    double w = 0.0;
    if (ld) {
      const double x = __ldcg(X + idx);
      double eta;
      if (mode == 1) {
        eta = -step * __ldcg(RG + idx);
      } else if (mode == 2) {
        eta = __ldcg(kp.v[V_AUX] + idx);
      } else {
        const double g = __ldcg(RG + idx);
        eta = eta_zero ? 0.0 : __ldcg(kp.v[V_ETA] + idx);
        double heta = eta_zero ? 0.0 : (__ldcg(kp.v[V_RES] + idx) - g);
        if (tau != 0.0) {
          eta = fma(tau, __ldcg(dcur + idx), eta);
          heta = fma(tau, __ldcg(kp.v[V_HD] + idx), heta);
        }
        acc[0] = fma(eta, g, acc[0]);
        acc[1] = fma(eta, heta, acc[1]);
      }
      w = x + eta;
    }
    const double q = qf_retract_elem<R, DH>(w, it.a, it.c);
    if (ld) X2[idx] = q;
  }
}

// Final phase: make X0 hold the result and accumulate |X_out - X_in|^2 (ref: relativeChange,
// src/QuadraticOptimizer.cpp:54).
template <int R, int DH> __device__ void phase_final(const KParams &kp, const CtaRows &cr, int cur, double (&acc)[NRED]) {
  constexpr int TS = R * DH;
  RowIter<R> it(cr);
  const bool valid = (it.a < R) && (it.c < DH);
  const int e = it.c * R + it.a;
  for (int jb = it.jb; jb < it.r1; jb += it.stride) {
    const int j = jb + it.sgw;
    if (j < it.r1 && valid) {
      const size_t idx = (size_t)j * TS + e;
      const double x = __ldcg(kp.v[V_X0 + cur] + idx);
      const double d = x - __ldcg(kp.v[V_XIN] + idx);
      acc[0] = fma(d, d, acc[0]);
      if (cur != 0) kp.v[V_X0][idx] = x;
    }
  }
}

__device__ __forceinline__ void zero(double (&acc)[NRED]) {
#pragma unroll
  for (int q = 0; q < NRED; ++q) acc[q] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// The persistent kernel
// ---------------------------------------------------------------------------------------------
template <int R, int DH> __global__ void __launch_bounds__(OPT_THREADS, 1) k_optimize(const KParams kp) {
  extern __shared__ double smem[];
  BlockCtx bc;
  bc.sm_warp = smem;
  bc.sm_out = smem + (OPT_THREADS / 32) * NRED;
  // the CTA's rows and, when they fit, its slice of the block-CSR structure in shared memory (SP_CACHE_INTS ints)
  int *sp_ints = reinterpret_cast<int *>(bc.sm_out + 2 * NRED);
  CtaRows cr;
  cr.r0 = ld_const(kp.cta_rows + blockIdx.x);
  cr.r1 = ld_const(kp.cta_rows + blockIdx.x + 1);
  cr.rowptr = kp.rowptr;
  cr.bcol = kp.bcol;
  cr.bval = kp.bval;
  {
    const int nrows = cr.r1 - cr.r0;
    const int blk0 = (nrows > 0) ? ld_const(kp.rowptr + cr.r0) : 0, blk1 = (nrows > 0) ? ld_const(kp.rowptr + cr.r1) : 0;
    if (nrows > 0 && nrows + 1 + (blk1 - blk0) <= SP_CACHE_INTS) {
      for (int q = threadIdx.x; q <= nrows; q += blockDim.x) sp_ints[q] = ld_const(kp.rowptr + cr.r0 + q) - blk0;
      for (int q = threadIdx.x; q < blk1 - blk0; q += blockDim.x) sp_ints[nrows + 1 + q] = ld_const(kp.bcol + blk0 + q);
      cr.rowptr = sp_ints - cr.r0;              // indexed by the global row
      cr.bcol = sp_ints + nrows + 1;            // indexed by the local block offset
      cr.bval = kp.bval + (size_t)blk0 * 16;
    }
    __syncthreads();
  }
  double *sV = reinterpret_cast<double *>(sp_ints + SP_CACHE_INTS);   // dense-preconditioner staging / sparse-plan areas
  DenseRing ring;
  ring.buf = sV + (size_t)DENSE_PER_MAX * R;    // 16-byte aligned: every preceding block is a multiple of 2 doubles
  ring.full = reinterpret_cast<uint64_t *>(ring.buf + (size_t)DENSE_RING_DOUBLES);
  ring.empty = ring.full + DENSE_NST;
  ring.count = 0;
  int *sMeta = reinterpret_cast<int *>(ring.empty + DENSE_NST);         // symmetric variant: (segment, first row) per window slot
  // bulk-TMA streaming needs 16-byte aligned rows (N even) and only pays off for a real stream
  const bool dense_tma = (kp.prm.precond == DPGO_PRECOND_DENSE_EXACT) && (kp.pinv != nullptr) && ((kp.N & 1) == 0) && (kp.N >= 2048);
  const bool dense_sym = dense_tma && (kp.sym_ok != 0);
  if (dense_tma) {
    if (dense_sym) {       // stale shared memory must be finite: tiles past N are multiplied by zero operands
      for (int q = threadIdx.x; q < DENSE_RING_DOUBLES; q += blockDim.x) ring.buf[q] = 0.0;
    }
    if (threadIdx.x == 0) {
      for (int st = 0; st < DENSE_NST; ++st) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&ring.full[st])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&ring.empty[st])), "r"(OPT_THREADS / 32 - 1));
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  bc.epoch = *kp.bar_epoch;
  bc.parity = 0;
  // diagnostic phase clock: CTA 0 / thread 0 charges the time since the previous tick to a phase kind
  // (0 eval, 1 dense apply, 2 partial sums + projection, 3 Hessian product, 4 tCG update, 5 retraction, 6 final)
  unsigned long long tick_last = 0;
  const bool ticking = (kp.phase_ns != nullptr) && blockIdx.x == 0 && threadIdx.x == 0;
  auto tick = [&](int kind) {
    if (ticking) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (kind >= 0) kp.phase_ns[kind] += t - tick_last;
      tick_last = t;
    }
  };
  tick(-1);
  const dpgo_opt_params_t prm = kp.prm;
  const int precond = prm.precond;
  const bool exact = (precond == DPGO_PRECOND_DENSE_EXACT) || (precond == DPGO_PRECOND_SPARSE_EXACT);
  double acc[NRED];
  // sparse exact preconditioner: shared memory = gathered input tiles + partial-sum slots (aliases the dense ring)
  double *nd_ys = sV;
  double *nd_slots = sV + (size_t)kp.nd.max_ytiles * R * DH;
  int4 *nd_grec = reinterpret_cast<int4 *>((reinterpret_cast<uintptr_t>(nd_slots + (size_t)kp.nd.max_slots * nd::PANEL_ROWS * R) + 15) & ~(uintptr_t)15);
  // Z = P_X( (Q + 0.1 I)^-1 V ), returns <Z, V> in acc[0]; every phase ends with a grid barrier
  auto apply_exact = [&](const double *Vv, int cbx, double *Zout) {
    if (precond == DPGO_PRECOND_SPARSE_EXACT) {
      zero(acc);
      for (int ph = 0; ph < kp.nd.nphases; ++ph) {
        phase_nd<R, DH>(kp, ph, Vv, cbx, Zout, nd_ys, nd_slots, nd_grec, acc);
        if (ph + 1 < kp.nd.nphases) { phase_end<0>(kp, bc, acc); tick(8 + min(ph, 15)); }
      }
      phase_end<1>(kp, bc, acc);
      tick(8 + min(kp.nd.nphases - 1, 15));
    } else {
      if (dense_sym) phase_dense_sym<R>(kp, Vv, ring, sV, sMeta);
      else if (dense_tma) phase_dense_tma<R>(kp, Vv, sV, ring);
      else phase_dense<R>(kp, Vv, sV);
      zero(acc); phase_end<0>(kp, bc, acc);
      tick(1);
      zero(acc); phase_pz<R, DH>(kp, cr, cbx, Vv, Zout, sV, acc);
      phase_end<1>(kp, bc, acc);
      tick(2);
    }
  };
  dpgo_opt_result_t res;
  res.success = 0; res.tcg_status = DPGO_TCG_NOT_RUN; res.tcg_iterations = 0; res.outer_iterations = 0;
  res.rejections = 0; res.spmv_passes = 0; res.precond_applies = 0; res.reserved0 = 0;
  res.f_init = res.gradnorm_init = res.f_opt = res.gradnorm_opt = res.relative_change = res.elapsed_ms = 0.0;
  res.quad_init = res.lin_init = 0.0;

  int cur = 0;   // which X buffer holds the current iterate

  // ---- single-operation entry points -------------------------------------------------------
  if (kp.op == OP_PHASE_BENCH) {          // diagnostic: tr_max_inner empty phases (barrier + 1-scalar reduction)
    for (int i = 0; i < prm.tr_max_inner; ++i) { zero(acc); acc[0] = 1.0; phase_end<1>(kp, bc, acc); }
    if (blockIdx.x == 0 && threadIdx.x == 0) { res.f_init = acc[0]; *kp.result = res; *kp.bar_epoch = bc.epoch; }
    return;
  }
  if (kp.op == OP_PRECON) {
    if (exact) {
      apply_exact(kp.v[V_AUX], 0, kp.v[V_Z]);
    } else {
      // reuse phase_update with res := AUX (first = false, alpha = 0 would need RES); do it directly
      constexpr int TS = R * DH;
      RowIter<R> it(cr);
      const bool valid = (it.a < R) && (it.c < DH);
      const int e = it.c * R + it.a;
      for (int jb = it.jb; jb < it.r1; jb += it.stride) {
        const int j = jb + it.sgw;
        const bool act = (j < it.r1);
        const int js = act ? j : it.r1 - 1;
        const bool ld = act && valid;
        const size_t idx = (size_t)js * TS + e;
        const double x = ld ? __ldcg(kp.v[V_X0] + idx) : 0.0;
        double t = ld ? __ldcg(kp.v[V_AUX] + idx) : 0.0;
        if (precond == DPGO_PRECOND_BLOCK_JACOBI) t = jacobi_elem<R, DH>(kp.dinv, js, t, it.a, it.c);
        double ya[3], sym[3];
        const double z = tangent_project_elem<R, DH>(x, t, it.a, it.c, ya, sym);
        if (ld) kp.v[V_Z][idx] = z;
      }
      zero(acc); phase_end(kp, bc, acc);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *kp.bar_epoch = bc.epoch;
    return;
  }
  if (kp.op == OP_RETRACT) {
    zero(acc);
    phase_retract<R, DH>(kp, cr, 0, 2, nullptr, 0.0, false, 0.0, acc);
    phase_end(kp, bc, acc);
    if (blockIdx.x == 0 && threadIdx.x == 0) *kp.bar_epoch = bc.epoch;
    return;
  }

  // ---- statistics at the input point (ref: src/QuadraticOptimizer.cpp:36-37) -------------------
  zero(acc);
  phase_eval<R, DH>(kp, cr, 0, true, precond, acc);
  phase_end(kp, bc, acc);
  tick(0);
  res.spmv_passes++;
  double f1 = 0.5 * acc[0] + acc[1];
  double gn = sqrt(acc[2]);
  double zr0 = acc[3];
  res.f_init = f1;
  res.gradnorm_init = gn;
  res.quad_init = acc[0];
  res.lin_init = acc[1];
  res.f_opt = f1;
  res.gradnorm_opt = gn;

  if (kp.op == OP_EVAL) {
    res.success = 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) { *kp.result = res; *kp.bar_epoch = bc.epoch; }
    return;
  }
  if (kp.op == OP_RHESS) {
    zero(acc);
    phase_hess<R, DH>(kp, cr, 0, kp.v[V_AUX], nullptr, nullptr, 0.0, false, acc);
    phase_end(kp, bc, acc);
    if (blockIdx.x == 0 && threadIdx.x == 0) { *kp.result = res; *kp.bar_epoch = bc.epoch; }
    return;
  }

  if (prm.algorithm == DPGO_ALG_RGD) {
    // ---- one fixed-step Riemannian gradient-descent step (ref :124-149) ----------------------
    zero(acc);
    phase_retract<R, DH>(kp, cr, 0, 1, nullptr, 0.0, false, prm.rgd_stepsize, acc);
    phase_end(kp, bc, acc);
    zero(acc);
    phase_eval<R, DH>(kp, cr, 1, false, DPGO_PRECOND_NONE, acc);
    phase_end(kp, bc, acc);
    res.spmv_passes++;
    res.f_opt = 0.5 * acc[0] + acc[1];
    res.gradnorm_opt = sqrt(acc[2]);
    res.outer_iterations = 1;
    cur = 1;
  } else if (gn >= prm.tr_tolerance) {           // ref :67-70 early exit otherwise
    // ---- Riemannian trust region ------------------------------------------------------------
    const bool single = (prm.tr_iterations == 1);      // ref :92-110 shrink-until-accepted mode
    double Delta = prm.tr_initial_radius;
    const double Delta_max = single ? Delta : 5.0 * prm.tr_initial_radius;   // ref :80-81,:96-97
    int total_steps = 0;
    int cb = 0;                      // base buffer
    bool z0_valid = !exact;
    int iter = 0;
    while (true) {
      // -- z0 = M^-1 g for the dense preconditioner (pose-local ones were fused into phase E)
      if (!z0_valid) {
        apply_exact(kp.v[V_RG0 + cb], cb, kp.v[V_Z00 + cb]);
        zr0 = acc[0];
        z0_valid = true;
      }
      // -- truncated CG (ROPTLIB SolversTR::tCG_TR; theta = 1, kappa = 0.1, Min_Inner_Iter = 0)
      double z_r = zr0, d_Pd = zr0, e_Pd = 0.0, e_Pe = 0.0;
      const double n0 = gn;
      res.precond_applies++;                       // z0 = M^-1 g
      double beta = 0.0, tau = 0.0;
      const double *zsrc = kp.v[V_Z00 + cb];
      int pd = 0;                                  // delta_old lives in V_D0 + pd
      bool eta_zero = true;
      int status = DPGO_TCG_MAXITER;
      const double *dcur = nullptr;
      for (int j = 0; j < prm.tr_max_inner; ++j) {
        double *dnew = kp.v[V_D0 + (1 - pd)];
        zero(acc);
        phase_hess<R, DH>(kp, cr, cb, zsrc, kp.v[V_D0 + pd], dnew, beta, true, acc);
        phase_end<1>(kp, bc, acc);
        tick(3);
        res.spmv_passes++;
        res.tcg_iterations++;
        pd = 1 - pd;
        dcur = dnew;
        const double d_Hd = acc[0];
        const double alpha = z_r / d_Hd;
        const double e_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
        if (d_Hd <= 0.0 || e_new >= Delta * Delta) {
          tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
          status = (d_Hd <= 0.0) ? DPGO_TCG_NEGCURVTURE : DPGO_TCG_EXCREGION;
          break;
        }
        e_Pe = e_new;
        zero(acc);
        phase_update<R, DH>(kp, cr, cb, dcur, alpha, eta_zero, precond, acc);
        phase_end<2>(kp, bc, acc);
        tick(4);
        eta_zero = false;
        const double nr = sqrt(acc[0]);
        const double n0t = n0;                         // n0^theta, theta = 1
        if (nr <= n0 * fmin(n0t, 0.1)) {
          status = (0.1 < n0t) ? DPGO_TCG_LCON : DPGO_TCG_SCON;
          break;
        }
        double zr_new = acc[1];
        if (exact) {
          apply_exact(kp.v[V_RES], cb, kp.v[V_Z]);
          zr_new = acc[0];
        }
        res.precond_applies++;
        beta = zr_new / z_r;
        z_r = zr_new;
        zsrc = kp.v[V_Z];
        e_Pd = beta * (e_Pd + alpha * d_Pd);
        d_Pd = z_r + beta * beta * d_Pd;
      }
      res.tcg_status = status;
      res.outer_iterations++;
      // -- candidate point, model decrease, actual decrease
      zero(acc);
      phase_retract<R, DH>(kp, cr, cb, 0, dcur, tau, eta_zero, 0.0, acc);
      phase_end<2>(kp, bc, acc);
      tick(5);
      const double denom = -acc[0] - 0.5 * acc[1];
      zero(acc);
      phase_eval<R, DH>(kp, cr, 1 - cb, false, precond, acc);
      phase_end(kp, bc, acc);
      tick(0);
      res.spmv_passes++;
      const double f2 = 0.5 * acc[0] + acc[1];
      const double gn2 = sqrt(acc[2]);
      const double rho = (denom != 0.0) ? (f1 - f2) / denom : -1.0;
      const bool accepted = rho > 0.1;                 // ROPTLIB Acceptence_Rho
      if (single) {
        if (accepted) {
          cb = 1 - cb; res.f_opt = f2; res.gradnorm_opt = gn2;
          break;
        }
        res.rejections++;
        if (total_steps > 10) break;                   // ref :101-103 return the initial guess
        Delta *= 0.25;                                 // ref :104-107
        total_steps++;
      } else {
        // ROPTLIB SolversTR radius update (Shrinked_tau = 0.25, Magnified_tau = 2)
        if (rho < 0.25) Delta *= 0.25;
        else if (rho > 0.75 && (status == DPGO_TCG_NEGCURVTURE || status == DPGO_TCG_EXCREGION))
          Delta = fmin(2.0 * Delta, Delta_max);
        if (accepted) {
          cb = 1 - cb; f1 = f2; gn = gn2; zr0 = acc[3];
          res.f_opt = f2; res.gradnorm_opt = gn2;
          z0_valid = !exact;
        } else {
          res.rejections++;
        }
        ++iter;
        if (gn < prm.tr_tolerance || iter >= prm.tr_iterations) break;
      }
    }
    cur = cb;
  }

  zero(acc);
  phase_final<R, DH>(kp, cr, cur, acc);
  phase_end<1>(kp, bc, acc);
  tick(6);
  res.relative_change = sqrt(acc[0] / (double)kp.n);
  res.success = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) { *kp.result = res; *kp.bar_epoch = bc.epoch; }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone Q.X product: Out = X Q (+ G).  One sub-group per pose tile, plain grid.
// ---------------------------------------------------------------------------------------------
constexpr int SPMV_THREADS = 256;

template <int R, int DH>
__global__ void __launch_bounds__(SPMV_THREADS) k_spmv(int n, const int *__restrict__ rowptr,
                                                       const int *__restrict__ bcol,
                                                       const double *__restrict__ bval,
                                                       const double *__restrict__ X,
                                                       const double *__restrict__ G, double *__restrict__ out) {
  constexpr int SG = SubGroup<R>::SG;
  constexpr int TS = R * DH;
  const int lane = threadIdx.x & 31;
  const int l = lane & (SG - 1), a = l >> 2, c = l & 3;
  const int sg_global = (blockIdx.x * SPMV_THREADS + threadIdx.x) / SG;
  const int j = sg_global;
  const bool act = j < n;
  const int js = act ? j : n - 1;
  double v = gather_tile<R, DH, false, false>(rowptr, bcol, bval, X, nullptr, 0.0, js, a, c);
  if (act && a < R && c < DH) {
    const size_t idx = (size_t)js * TS + c * R + a;
    if (G != nullptr) v += __ldg(G + idx);
    out[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Stiefel (polar-factor) projection per pose, ref: LiftedSEManifold::project,
// src/manifold/LiftedSEManifold.cpp:34-45 + projectToStiefelManifold, src/DPGO_utils.cpp:479-485
// (U V^T of the thin SVD = polar factor).  One thread per pose: one-sided (Hestenes) Jacobi SVD of
// the r x d block -- columns are rotated until mutually orthogonal (Y V = U Sigma), which keeps high
// relative accuracy for ill-conditioned blocks -- then out = U V^T.
// ---------------------------------------------------------------------------------------------
// The input is the linear combination c0 A + c1 B + c2 C (B, C optional): the Nesterov updates of the accelerated RBCD,
// Y = proj((1 - alpha) X + alpha V) and V = proj(V + gamma (X - Y)) (ref src/PGOAgent.cpp:1077-1091), are one launch each.
// (M and out may be the same buffer: a thread reads its whole tile before it writes it -- hence no __restrict__.)
template <int R, int DH> __global__ void k_stiefel_project(int n, const double *M, double *out, double c0, const double *B, double c1,
                                                            const double *Cc, double c2) {
  constexpr int D = DH - 1;
  constexpr int TS = R * DH;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  auto in = [&](int e) {
    double v = c0 * M[(size_t)j * TS + e];
    if (B) v = fma(c1, B[(size_t)j * TS + e], v);
    if (Cc) v = fma(c2, Cc[(size_t)j * TS + e], v);
    return v;
  };
  double y[D][R];
  double V[D][D];
#pragma unroll
  for (int c = 0; c < D; ++c) {
#pragma unroll
    for (int a = 0; a < R; ++a) y[c][a] = in(c * R + a);
#pragma unroll
    for (int q = 0; q < D; ++q) V[c][q] = (c == q) ? 1.0 : 0.0;     // V[c] = column c of V
  }
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int p = 0; p < D; ++p)
#pragma unroll
      for (int q = p + 1; q < D; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
        for (int a = 0; a < R; ++a) {
          alpha = fma(y[p][a], y[p][a], alpha);
          beta = fma(y[q][a], y[q][a], beta);
          gamma = fma(y[p][a], y[q][a], gamma);
        }
        if (fabs(gamma) <= 1e-17 * sqrt(alpha * beta) || gamma == 0.0) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
        for (int a = 0; a < R; ++a) {
          const double yp = y[p][a], yq = y[q][a];
          y[p][a] = cs * yp - sn * yq;
          y[q][a] = sn * yp + cs * yq;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double vp = V[p][k], vq = V[q][k];
          V[p][k] = cs * vp - sn * vq;
          V[q][k] = sn * vp + cs * vq;
        }
      }
    if (!rotated) break;
  }
  // normalise the rotated columns: U = (Y V) Sigma^-1
#pragma unroll
  for (int c = 0; c < D; ++c) {
    double s = 0;
#pragma unroll
    for (int a = 0; a < R; ++a) s = fma(y[c][a], y[c][a], s);
    const double inv = (s > 0.0) ? 1.0 / sqrt(s) : 0.0;
#pragma unroll
    for (int a = 0; a < R; ++a) y[c][a] *= inv;
  }
  // out = U V^T : out[a, c] = sum_k U[a,k] V[c,k]   (V[k][c'] holds entry c' of column k)
#pragma unroll
  for (int c = 0; c < D; ++c)
#pragma unroll
    for (int a = 0; a < R; ++a) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < D; ++k) s = fma(y[k][a], V[k][c], s);
      out[(size_t)j * TS + c * R + a] = s;
    }
#pragma unroll
  for (int a = 0; a < R; ++a) out[(size_t)j * TS + D * R + a] = in(D * R + a);
}

// ---------------------------------------------------------------------------------------------
// Boundary-pose exchange helpers (ref: PGOAgent::getSharedPoseDict src/PGOAgent.cpp:95-105,
// PGOAgent::constructGMatrix :783-859)
// ---------------------------------------------------------------------------------------------
__global__ void k_pack_tiles(int ts, int count, const int *__restrict__ pose, const double *__restrict__ X,
                             double *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * ts) return;
  const int s = t / ts, e = t - s * ts;
  out[t] = X[(size_t)pose[s] * ts + e];
}

// One thread per (pose with shared edges, element): walks that pose's edges in a fixed order
// (deterministic sum).  outgoing: G_p += -(X_j Om) T^T ; incoming: G_p += -(X_i T) Om.
template <int R, int DH>
__global__ void k_build_G(int nposes, const int *__restrict__ pose_ids, const int *__restrict__ pose_ptr,
                          const int *__restrict__ edge_slot, const int *__restrict__ edge_out,
                          const double *__restrict__ edge_T, const double *__restrict__ edge_om,
                          const double *__restrict__ gathered, double *__restrict__ G) {
  constexpr int TS = R * DH;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nposes * TS) return;
  const int pi = t / TS, e = t - pi * TS;
  const int c = e / R, a = e - c * R;          // element (a, c) of the tile
  double acc = 0.0;
  for (int k = pose_ptr[pi]; k < pose_ptr[pi + 1]; ++k) {
    const double *Xn = gathered + (size_t)edge_slot[k] * TS;     // neighbour tile
    const double *T = edge_T + (size_t)k * DH * DH;              // row-major DH x DH
    const double *om = edge_om + (size_t)k * DH;
    double s = 0.0;
    if (edge_out[k]) {        // L[a,c] = -sum_q Xn[a,q] om[q] T[c][q]
#pragma unroll
      for (int q = 0; q < DH; ++q) s = fma(Xn[q * R + a] * om[q], T[c * DH + q], s);
    } else {                  // L[a,c] = -sum_q Xn[a,q] T[q][c] om[c]
#pragma unroll
      for (int q = 0; q < DH; ++q) s = fma(Xn[q * R + a], T[q * DH + c], s);
      s *= om[c];
    }
    acc -= s;
  }
  G[(size_t)pose_ids[pi] * TS + e] = acc;
}

// ---------------------------------------------------------------------------------------------
// Q from edge records on the device (ref: constructConnectionLaplacianSE, src/DPGO_utils.cpp:199-271, and the diagonal terms
// of PGOAgent::constructQMatrix, src/PGOAgent.cpp:720-781): one thread per block entry walks the block's contribution
// list in input order (deterministic).  Per edge i -> j with T = [R t; 0 1], Om = w diag(kappa.., tau):
//   kind 0  Q_ii += T Om T^T     kind 1  Q_jj += Om     kind 2  Q_ij = -T Om     kind 3  Q_ji = -Om T^T     kind 4  static block
// ---------------------------------------------------------------------------------------------
__global__ void k_assemble_Q(int64_t nb, const int *__restrict__ cptr, const int2 *__restrict__ contrib, const double *__restrict__ eT,
                             const double *__restrict__ eom, const double *__restrict__ ew, const double *__restrict__ sblk,
                             double *__restrict__ bval) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nb * 16) return;
  const int64_t b = tid >> 4;
  const int k = (int)((tid >> 2) & 3), c = (int)(tid & 3);
  double v = 0.0;
  for (int q = cptr[b]; q < cptr[b + 1]; ++q) {
    const int2 cc = contrib[q];
    if (cc.y == 4) { v += sblk[(size_t)cc.x * 16 + k * 4 + c]; continue; }
    const double *T = eT + (size_t)cc.x * 16, *om = eom + (size_t)cc.x * 4;
    const double w = ew[cc.x];
    if (cc.y == 0) {
      double s = 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) s = fma(T[k * 4 + u] * (om[u] * w), T[c * 4 + u], s);
      v += s;
    } else if (cc.y == 1) {
      if (k == c) v += om[k] * w;
    } else if (cc.y == 2) {
      v -= T[k * 4 + c] * (om[c] * w);
    } else {
      v -= (om[k] * w) * T[c * 4 + k];
    }
  }
  bval[tid] = v;
}

// Robust re-weighting (ref: PGOAgent::updateLoopClosuresWeights, src/PGOAgent.cpp:1181-1245; computeMeasurementError,
// src/DPGO_utils.cpp:494-500; RobustCost::weight, src/DPGO_robust.cpp:23-66): one thread per private edge evaluates
// r^2 = kappa |Y_i R - Y_j|^2 + tau |p_j - p_i - Y_i t|^2 at the resident iterate and the weight of the chosen loss.
// cost: 0 L2, 1 L1, 2 Huber(param), 3 TLS(param), 4 Geman-McClure, 5 GNC_TLS(mu, param = cbar)
template <int R, int DH>
__global__ void k_edge_weights(int64_t m, const int *__restrict__ p1, const int *__restrict__ p2, const double *__restrict__ eT,
                               const double *__restrict__ eom, const int *__restrict__ fixed, const double *__restrict__ X, int cost,
                               double mu, double param, double *__restrict__ w, double *__restrict__ resid) {
  constexpr int D = DH - 1, TS = R * DH;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const double *T = eT + (size_t)e * 16, *om = eom + (size_t)e * 4;
  const double *X1 = X + (size_t)p1[e] * TS, *X2 = X + (size_t)p2[e] * TS;
  double rot = 0.0, tra = 0.0;
  for (int a = 0; a < R; ++a) {
    for (int c = 0; c < D; ++c) {
      double s = -X2[c * R + a];
      for (int q = 0; q < D; ++q) s = fma(X1[q * R + a], T[q * 4 + c], s);
      rot = fma(s, s, rot);
    }
    double s = X2[D * R + a] - X1[D * R + a];
    for (int q = 0; q < D; ++q) s = fma(-X1[q * R + a], T[q * 4 + D], s);
    tra = fma(s, s, tra);
  }
  const double r2 = om[0] * rot + om[D] * tra;
  if (resid) resid[e] = r2;
  if (fixed && fixed[e]) return;
  const double r = sqrt(r2);
  double wt = 1.0;
  if (cost == 1) wt = 1.0 / r;
  else if (cost == 2) wt = (r < param) ? 1.0 : param / r;
  else if (cost == 3) wt = (r < param) ? 1.0 : 0.0;
  else if (cost == 4) { const double s = 1.0 + r2; wt = 1.0 / (s * s); }
  else if (cost == 5) {
    const double c2 = param * param;
    if (r2 >= c2 * (mu + 1.0) / mu) wt = 0.0;
    else if (r2 <= c2 * mu / (mu + 1.0)) wt = 1.0;
    else wt = sqrt(c2 * mu * (mu + 1.0) / r2) - mu;
  }
  w[e] = wt;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
// Dynamic shared memory of one launch: the reduction scratch, plus what the launch's preconditioner stages (the dense
// ring or the sparse plan's tiles and slots).  Asking for no more than needed leaves the rest of the SM's 228 KB to L1,
// which now keeps the constant data (block-CSR, plan records) across phases.
template <int R, int DH> static size_t optimize_smem_doubles(const KParams &kp, bool max_only) {
  const size_t base = (OPT_THREADS / 32) * NRED + 2 * NRED + SP_CACHE_INTS / 2;
  const size_t dense = (size_t)DENSE_PER_MAX * R + (size_t)DENSE_RING_DOUBLES + 2 * DENSE_NST + (size_t)SYM_META_DOUBLES;
  if (max_only || kp.prm.precond == DPGO_PRECOND_DENSE_EXACT) return base + dense;
  if (kp.prm.precond == DPGO_PRECOND_SPARSE_EXACT)
    return base + (size_t)kp.nd.max_ytiles * R * DH + (size_t)(kp.nd.max_slots + 1) * nd::PANEL_ROWS * R + 4 * (size_t)kp.nd.max_gathers + 8;
  return base + 8;
}

template <int R, int DH> static cudaError_t launch_optimize_t(const KParams &kp_in, cudaStream_t stream) {
  KParams kp = kp_in;
  const size_t smem_max = optimize_smem_doubles<R, DH>(kp, true) * sizeof(double);
  static_assert((size_t)ND_YCAP_TILES * R * DH + (size_t)(ND_SLOT_CAP + 1) * nd::PANEL_ROWS * R + 4 * (size_t)ND_YCAP_TILES + 8 <=
                    (size_t)DENSE_PER_MAX * R + (size_t)DENSE_RING_DOUBLES + 2 * DENSE_NST + (size_t)SYM_META_DOUBLES,
                "the sparse plan's capacities must fit the kernel's maximum shared memory");
  kp.smem_doubles = (int)optimize_smem_doubles<R, DH>(kp, false);
  const size_t smem = (size_t)kp.smem_doubles * sizeof(double);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_optimize<R, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  {
    // shared-memory carve-out hint: what this launch needs (+ the 1 KB the system reserves), so that L1 gets the rest
    static int last_pct[64];
    static bool init = false;
    if (!init) { for (int i = 0; i < 64; ++i) last_pct[i] = -1; init = true; }
    const int pct = std::min(100, (int)((smem + 2048) * 100 / (228 * 1024)) + 1);
    if (dev >= 0 && dev < 64 && last_pct[dev] != pct) {
      cudaFuncSetAttribute(k_optimize<R, DH>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      last_pct[dev] = pct;
    }
  }
  if (kp.cluster) {
    // one cluster = the whole grid: an ordinary (non-cooperative) launch, co-scheduling is guaranteed by the cluster
    static bool np_set[64] = {};
    if (kp.grid > 8 && (dev < 0 || dev >= 64 || !np_set[dev])) {
      cudaError_t e = cudaFuncSetAttribute(k_optimize<R, DH>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
      if (dev >= 0 && dev < 64) np_set[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kp.grid);
    cfg.blockDim = dim3(OPT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)kp.grid;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_optimize<R, DH>, kp);
  }
  void *args[] = {(void *)&kp};
  return cudaLaunchCooperativeKernel((void *)k_optimize<R, DH>, dim3(kp.grid), dim3(OPT_THREADS), args, smem, stream);
}

template <int R, int DH> static int max_cluster_t(int device) {
  const size_t smem = ((OPT_THREADS / 32) * NRED + 2 * NRED + SP_CACHE_INTS / 2 + (size_t)ND_SMEM_NEED_SMALL) * sizeof(double);
  cudaFuncSetAttribute(k_optimize<R, DH>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {16, 8}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(OPT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, k_optimize<R, DH>, &cfg) == cudaSuccess && nclusters >= 1) return cs;
  }
  cudaGetLastError();
  return 0;
}

template <int R, int DH> static int max_grid_t(int device) {
  const size_t smem = ((OPT_THREADS / 32) * NRED + 2 * NRED + SP_CACHE_INTS / 2 + (size_t)DENSE_PER_MAX * R + (size_t)DENSE_RING_DOUBLES + 2 * DENSE_NST + (size_t)SYM_META_DOUBLES) * sizeof(double);
  cudaFuncSetAttribute(k_optimize<R, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_optimize<R, DH>, OPT_THREADS, smem) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  return per_sm > 0 ? sms : 0;      // one CTA per SM
}

#define DPGO_DISPATCH(R_, DH_, ...)                                    \
  do {                                                                   \
    if ((DH_) == 4) {                                                    \
      switch (R_) {                                                      \
        case 3: { constexpr int R = 3, DH = 4; __VA_ARGS__; } break;     \
        case 4: { constexpr int R = 4, DH = 4; __VA_ARGS__; } break;     \
        case 5: { constexpr int R = 5, DH = 4; __VA_ARGS__; } break;     \
        default: break;                                                  \
      }                                                                  \
    } else if ((DH_) == 3) {                                             \
      switch (R_) {                                                      \
        case 2: { constexpr int R = 2, DH = 3; __VA_ARGS__; } break;     \
        case 3: { constexpr int R = 3, DH = 3; __VA_ARGS__; } break;     \
        case 5: { constexpr int R = 5, DH = 3; __VA_ARGS__; } break;     \
        default: break;                                                  \
      }                                                                  \
    }                                                                    \
  } while (0)

cudaError_t launch_optimize(int r, int dh, const KParams &kp, cudaStream_t stream) {
  cudaError_t e = cudaErrorInvalidValue;
  DPGO_DISPATCH(r, dh, e = (launch_optimize_t<R, DH>(kp, stream)));
  return e;
}

int optimize_max_grid(int r, int dh, int device) {
  int g = 0;
  DPGO_DISPATCH(r, dh, g = (max_grid_t<R, DH>(device)));
  return g;
}

int optimize_max_cluster(int r, int dh, int device) {
  int g = 0;
  DPGO_DISPATCH(r, dh, g = (max_cluster_t<R, DH>(device)));
  return g;
}

cudaError_t launch_spmv(int r, int dh, int n, const int *rowptr, const int *bcol, const double *bval,
                        const double *X, const double *G, double *out, cudaStream_t stream) {
  bool ok = false;
  DPGO_DISPATCH(r, dh, {
    constexpr int SG = SubGroup<R>::SG;
    const int per_block = SPMV_THREADS / SG;
    const int blocks = (n + per_block - 1) / per_block;
    k_spmv<R, DH><<<blocks, SPMV_THREADS, 0, stream>>>(n, rowptr, bcol, bval, X, G, out);
    ok = true;
  });
  if (!ok) return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_stiefel_project(int r, int dh, int n, const double *M, double *out, cudaStream_t stream, double c0,
                                   const double *B, double c1, const double *C, double c2) {
  bool ok = false;
  DPGO_DISPATCH(r, dh, {
    k_stiefel_project<R, DH><<<(n + 127) / 128, 128, 0, stream>>>(n, M, out, c0, B, c1, C, c2);
    ok = true;
  });
  if (!ok) return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_pack_tiles(int ts, int count, const int *pose, const double *X, double *out, cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  const int total = count * ts;
  k_pack_tiles<<<(total + 255) / 256, 256, 0, stream>>>(ts, count, pose, X, out);
  return cudaGetLastError();
}


// dense A <- block-CSR Q (+ shift * I): one thread per (block, entry); the output tile of block b is found by
// binary search in rowptr.  A[(dh*i + k) + N*(dh*j + c)] = bval[b][k][c] with i = bcol[b], j = row of b.
__global__ void k_bsr_to_dense(int n, int dh, int64_t nb, const int *__restrict__ rowptr, const int *__restrict__ bcol,
                               const double *__restrict__ bval, double *__restrict__ A, int N) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb * 16) return;
  const int64_t b = t >> 4;
  const int k = (int)((t >> 2) & 3), c = (int)(t & 3);
  if (k >= dh || c >= dh) return;
  int lo = 0, hi = n;                 // largest j with rowptr[j] <= b
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (rowptr[mid] <= b) lo = mid; else hi = mid;
  }
  const int j = lo, i = bcol[b];
  A[(size_t)(dh * i + k) + (size_t)N * (dh * j + c)] = bval[t];
}
// one block per chunk: copy the 8 x width piece of the dense inverse into its padded chunk-major slot (zero fill)
__global__ void k_pack_sym(const double *__restrict__ pinv, int N, int nchunks, const int *__restrict__ segptr, int nseg,
                           const long long *__restrict__ off, double *__restrict__ ppack) {
  for (int lin = blockIdx.x; lin < nchunks; lin += gridDim.x) {
    int J = 0;
    while (J + 1 < nseg && segptr[J + 1] <= lin) ++J;
    const int s0 = J * SYM_SEG, s1 = min(N, s0 + SYM_SEG);
    const int g0 = 8 * (lin - segptr[J]);
    const int col_lo = max(s0, g0), width = s1 - col_lo;
    const int pitch = ((width + 7) & ~7) + 4;
    double *dst = ppack + off[lin];
    for (int q = threadIdx.x; q < 8 * pitch; q += blockDim.x) {
      const int rr = q / pitch, x = q - rr * pitch;
      dst[q] = (g0 + rr < N && x < width) ? pinv[(size_t)(g0 + rr) * N + col_lo + x] : 0.0;
    }
  }
}

__global__ void k_add_diag(double *__restrict__ A, int N, double shift) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < N) A[(size_t)t * N + t] += shift;
}

cudaError_t launch_pack_sym(const double *pinv, int N, int nchunks, const int *segptr, int nseg, const long long *off, double *ppack,
                            cudaStream_t stream) {
  k_pack_sym<<<std::min(nchunks, 148 * 16), 256, 0, stream>>>(pinv, N, nchunks, segptr, nseg, off, ppack);
  return cudaGetLastError();
}

cudaError_t launch_bsr_to_dense(int n, int dh, int64_t nb, const int *rowptr, const int *bcol, const double *bval,
                                double shift, double *A, int N, cudaStream_t stream) {
  if (nb > 0) k_bsr_to_dense<<<(unsigned)((nb * 16 + 255) / 256), 256, 0, stream>>>(n, dh, nb, rowptr, bcol, bval, A, N);
  k_add_diag<<<(N + 255) / 256, 256, 0, stream>>>(A, N, shift);
  return cudaGetLastError();
}

cudaError_t launch_build_G(int r, int dh, int nposes, const int *pose_ids, const int *pose_ptr, const int *edge_slot,
                           const int *edge_out, const double *edge_T, const double *edge_om, const double *gathered,
                           double *G, cudaStream_t stream) {
  if (nposes <= 0) return cudaSuccess;
  bool ok = false;
  DPGO_DISPATCH(r, dh, {
    const int total = nposes * R * DH;
    k_build_G<R, DH><<<(total + 127) / 128, 128, 0, stream>>>(nposes, pose_ids, pose_ptr, edge_slot, edge_out, edge_T,
                                                              edge_om, gathered, G);
    ok = true;
  });
  if (!ok) return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_assemble_Q(int64_t nb, const int *cptr, const int2 *contrib, const double *eT, const double *eom, const double *ew,
                              const double *sblk, double *bval, cudaStream_t stream) {
  if (nb > 0) k_assemble_Q<<<(unsigned)((nb * 16 + 255) / 256), 256, 0, stream>>>(nb, cptr, contrib, eT, eom, ew, sblk, bval);
  return cudaGetLastError();
}

cudaError_t launch_edge_weights(int r, int dh, int64_t m, const int *p1, const int *p2, const double *eT, const double *eom,
                                const int *fixed, const double *X, int cost, double mu, double param, double *w, double *resid,
                                cudaStream_t stream) {
  if (m <= 0) return cudaSuccess;
  bool ok = false;
  DPGO_DISPATCH(r, dh, {
    k_edge_weights<R, DH><<<(unsigned)((m + 127) / 128), 128, 0, stream>>>(m, p1, p2, eT, eom, fixed, X, cost, mu, param, w, resid);
    ok = true;
  });
  if (!ok) return cudaErrorInvalidValue;
  return cudaGetLastError();
}


}  // namespace dpgo
