// dpgo_spmv_tma.cu -- the Q.X product (Out = X Q [+ G]) as a persistent, TMA-fed streaming kernel.
//
// Why: the measurement-block stream (128 B per block, contiguous for a contiguous range of pose tiles)
// is >75 % of the algorithmic bytes; a per-row gather kernel serialises 5-8 dependent DRAM latencies per
// row (rowptr -> indices -> blocks, batch after batch) and tops out near 28 % of HBM peak (profiles/).
// Here the row structure and the DRAM stream are decoupled:
//   * the host cuts the pose tiles into "row groups" of <= BT blocks (consecutive rows);
//   * each CTA walks its groups with an NSTAGE-deep ring of shared-memory stages; one elected thread
//     issues three 1-D bulk TMA copies per group (cp.async.bulk ... mbarrier::complete_tx): the 4x4 blocks,
//     their column indices, and the row-pointer slice -- 100+ KB in flight per SM, L2 evict-first policy
//     so the stream does not push the pose tiles (X) out of L2;
//   * the warps consume a group from shared memory: indices and blocks come from smem (no dependent global
//     round trips), only the X gathers go to L2, issued for a whole batch of 8 blocks at once.
// Lane mapping and reduce-scatter are those of dpgo_device.cuh (lane (a,k) holds P_i[a,k] and row k of
// the block; result element (a,c) ends in lane (a,c)).
#include "dpgo_device.cuh"
#include "dpgo_kernels.cuh"

#include <cstdlib>

namespace dpgo {

namespace {

constexpr int TMA_THREADS = 512;
constexpr int TMA_NSTAGE = 4;
constexpr int SPMV_BATCH = 12;   // blocks whose loads are issued together (predicated); rows <= 12 blocks take one round

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// 1-D bulk copy global -> shared, completion reported on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

}  // namespace

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// group g covers rows [gi[g].x, gi[g+1].x) and blocks [gi[g].y, gi[g+1].y)
//
// Warp roles: warp NCONS is the producer (one lane issues the bulk TMA copies of a group as soon as its
// stage has been released by all consumer warps -- "empty" mbarrier, count NCONS); warps 0..NCONS-1 are
// consumers that wait on the stage's "full" mbarrier (transaction bytes), process their share of the
// group's rows and release the stage.  No CTA-wide barrier in the steady state: fast warps run up to
// NSTAGE-1 groups ahead.
// PF != 0: the last consumer warp becomes a prefetch warp that runs ONE group ahead of the slowest consumer and pulls the
// pose tiles of that group's column indices into L1 (PF = 1: prefetch.global.L1, SASS CCTL.PF1; PF = 2: discarded
// ld.global.ca), so that the consumers' gathers hit L1 instead of paying an L2 round trip per row.
template <int R, int DH, int BT, int NST, int PF, int LEAN>
__global__ void __launch_bounds__(TMA_THREADS, 2)
    k_spmv_tma(int ngroups, const int2 *__restrict__ gi, const int *__restrict__ rowptr, const int *__restrict__ bcol,
               const double *__restrict__ bval, const double *__restrict__ X, const double *__restrict__ G,
               double *__restrict__ out) {
  constexpr int TS = R * DH;
  constexpr int IDX_CAP = BT + 8;        // ints per stage for indices (alignment slack)
  constexpr int RP_CAP = BT + 12;        // ints per stage for the row-pointer slice (a group has <= BT rows)
  constexpr int NCONS = TMA_THREADS / 32 - 1 - (PF ? 1 : 0);
  constexpr int NARR = NCONS + (PF ? 1 : 0);            // arrivals that release a stage
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sq = reinterpret_cast<double *>(smem_raw);                                     // NSTAGE * BT * 16
  int *sidx = reinterpret_cast<int *>(smem_raw + (size_t)NST * BT * 128);        // NSTAGE * IDX_CAP
  int *srp = sidx + NST * IDX_CAP;                                                // NSTAGE * RP_CAP
  uint64_t *full = reinterpret_cast<uint64_t *>(srp + NST * RP_CAP);              // NSTAGE
  uint64_t *empty = full + NST;                                                    // NSTAGE

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NARR); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NARR) {
    // ---------------- producer ----------------
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int it = 0;
      for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
        const int s = it % NST;
        if (it >= NST) mbar_wait(&empty[s], (unsigned)(((it / NST) - 1) & 1));
        const int2 g0 = __ldg(gi + g), g1 = __ldg(gi + g + 1);
        const int b0 = g0.y, b1 = g1.y, r0 = g0.x, r1 = g1.x;
        const unsigned qbytes = (unsigned)(b1 - b0) * 128u;
        const int ib = b0 & ~3;                                   // 16-byte aligned index window
        const unsigned ibytes = (unsigned)(((b1 - ib) + 3) & ~3) * 4u;
        const int rb = r0 & ~3;                                   // rowptr[r0 .. r1] inclusive
        const unsigned rbytes = (unsigned)(((r1 + 1 - rb) + 3) & ~3) * 4u;
        mbar_expect_tx(&full[s], qbytes + ibytes + rbytes);
        if (qbytes) tma_load_1d(sq + (size_t)s * BT * 16, bval + (size_t)b0 * 16, qbytes, &full[s], pol);
        tma_load_1d(sidx + s * IDX_CAP, bcol + ib, ibytes, &full[s], pol);
        tma_load_1d(srp + s * RP_CAP, rowptr + rb, rbytes, &full[s], pol);
      }
    }
    return;
  }

  if (PF != 0 && warp == NCONS) {
    // ---------------- prefetch warp ----------------
    const char *Xb = reinterpret_cast<const char *>(X);
    constexpr int SECT = (TS * 8 + 31) / 32;               // 32-byte sectors per pose tile
    int it = 0;
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
      const int s = it % NST;
      // one group ahead, not more: L1 holds about two groups' tiles per CTA
      if (it >= 2) mbar_wait(&empty[(it - 2) % NST], (unsigned)((((it - 2) / NST)) & 1));
      mbar_wait(&full[s], (unsigned)((it / NST) & 1));
      const int b0 = __ldg(gi + g).y;
      const int nsec = (__ldg(gi + g + 1).y - b0) * SECT;
      const int *idx_s = sidx + s * IDX_CAP + (b0 - (b0 & ~3));
      for (int c = lane; c < nsec; c += 32) {
        const int u = c / SECT, part = c - u * SECT;
        const char *src = Xb + ((size_t)idx_s[u] * (TS * 8) + (size_t)(part * 32));
        if constexpr (PF == 1) {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(src));
        } else {
          unsigned long long sink;
          asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(sink) : "l"(src));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    return;
  }

  // ---------------- consumers ----------------
  const int a = lane >> 2, k = lane & 3;
  const bool valid = (a < R) && (k < DH);
  const int off = k * R + a;
  const int bn = lane >> 2;                            // B fragment: column n = lane>>2, row k = lane&3
  const bool bvalid = bn < 4;
  const int boff = k * 4 + bn;
  int it = 0;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
    const int s = it % NST;
    mbar_wait(&full[s], (unsigned)((it / NST) & 1));
    const int2 g0 = __ldg(gi + g), g1 = __ldg(gi + g + 1);
    const int r0 = g0.x, r1 = g1.x, b0 = g0.y;
    const double *q_s = sq + (size_t)s * BT * 16;
    const int *idx_s = sidx + s * IDX_CAP + (b0 - (b0 & ~3));
    const int *rp_s = srp + s * RP_CAP + (r0 - (r0 & ~3));

    // One warp per pose tile.  The 4x4 block product runs on the fp64 tensor pipe (mma.sync m8n8k4, SASS DMMA):
    // the A fragment IS the lane mapping (lane = 4a + k holds P_i[a,k]), the B fragment is one 8-byte shared
    // load per lane (lanes 0..15 cover the 128-byte block exactly once -> a single smem wavefront), and the
    // accumulator fragment leaves Out_j[a, 2k..2k+1] in lane (a,k), k < 2 -- no shuffles, no 128-bit smem
    // broadcasts.  The starting warp rotates with the group so that short groups load all warps evenly.
    const int wrot = (warp + it * 7) % NCONS;
    const double *Xl = X + off;
    const double *q_l = q_s + boff;
    if constexpr (LEAN != 0) {
      // Predicate-light variant.  Rows >= R of the A fragment and columns >= 4 of the B fragment only reach accumulator
      // entries that are never stored, so those lanes load any in-bounds address instead of a predicated zero; batches
      // of 4 blocks run without per-block predicates, only the row's last 1-3 blocks are predicated.
      const double *Xg = X + ((a < R) ? off : k * R);        // lanes a >= R: row 0 of the tile (result row discarded)
      constexpr bool KPRED = (DH < 4);                        // k >= DH must contribute zeros (it is summed over)
      const bool kok = k < DH;
      for (int j = r0 + wrot; j < r1; j += NCONS) {
        const int lb0 = rp_s[j - r0] - b0, lb1 = rp_s[j - r0 + 1] - b0;
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
        for (int bb = lb0; bb < lb1; bb += SPMV_BATCH) {
          const int nb = min(lb1 - bb, SPMV_BATCH);          // warp-uniform
          const int nfull = nb >> 2, rem = nb & 3;
          const int *ib = idx_s + bb;
          double x[SPMV_BATCH];
#pragma unroll
          for (int t = 0; t < SPMV_BATCH / 4; ++t) {
            if (t < nfull) {
#pragma unroll
              for (int u = 4 * t; u < 4 * t + 4; ++u) {
                if (KPRED) { x[u] = 0.0; if (kok) x[u] = __ldg(Xg + (size_t)ib[u] * TS); }
                else x[u] = __ldg(Xg + (size_t)ib[u] * TS);
              }
            } else if (t == nfull && rem != 0) {
#pragma unroll
              for (int u = 4 * t; u < 4 * t + 4; ++u) {
                x[u] = 0.0;
                if (u - 4 * t < rem && (!KPRED || kok)) x[u] = __ldg(Xg + (size_t)ib[u] * TS);
              }
            }
          }
#pragma unroll
          for (int t = 0; t < SPMV_BATCH / 4; ++t) {
            const double *qb = q_l + (size_t)(bb + 4 * t) * 16;
            if (t < nfull) {
              const double q0 = qb[0], q1 = qb[16], q2 = qb[32], q3 = qb[48];
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(c0), "+d"(c1) : "d"(x[4 * t]), "d"(q0));
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(e0), "+d"(e1) : "d"(x[4 * t + 1]), "d"(q1));
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(c0), "+d"(c1) : "d"(x[4 * t + 2]), "d"(q2));
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(e0), "+d"(e1) : "d"(x[4 * t + 3]), "d"(q3));
            } else if (t == nfull && rem != 0) {
              // x is zero beyond the row's last block, so the B operand there may be anything finite: it is another
              // row's block (or, past the stage's last block, index words reinterpreted); keep it a predicated zero
              double q0 = qb[0], q1 = 0.0, q2 = 0.0;
              if (rem > 1) q1 = qb[16];
              if (rem > 2) q2 = qb[32];
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(c0), "+d"(c1) : "d"(x[4 * t]), "d"(q0));
              if (rem > 1)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(e0), "+d"(e1) : "d"(x[4 * t + 1]), "d"(q1));
              if (rem > 2)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0), "+d"(c1) : "d"(x[4 * t + 2]), "d"(q2));
            }
          }
        }
        c0 += e0;
        c1 += e1;
        if (a < R && k < 2) {
          double *o = out + ((size_t)j * TS + a + (size_t)(2 * k) * R);
          if (G != nullptr) {
            const double *gp = G + ((size_t)j * TS + a + (size_t)(2 * k) * R);
            c0 += __ldg(gp);
            if (2 * k + 1 < DH) c1 += __ldg(gp + R);
          }
          o[0] = c0;
          if (2 * k + 1 < DH) o[R] = c1;
        }
      }
    } else {
      for (int j = r0 + wrot; j < r1; j += NCONS) {
        const int lb0 = rp_s[j - r0] - b0, lb1 = rp_s[j - r0 + 1] - b0;
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;      // two accumulator pairs (shorter DMMA chains)
        for (int b = lb0; b < lb1; b += SPMV_BATCH) {
          const int nrem = lb1 - b;                            // warp-uniform
          double x[SPMV_BATCH];
  #pragma unroll
          for (int u0 = 0; u0 < SPMV_BATCH; u0 += 4) {         // all global gathers of the batch in flight together
            if (u0 < nrem) {
  #pragma unroll
              for (int u = u0; u < u0 + 4; ++u) {
                x[u] = 0.0;
                if (valid && u < nrem) x[u] = __ldg(Xl + (size_t)idx_s[b + u] * TS);
              }
            }
          }
  #pragma unroll
          for (int u0 = 0; u0 < SPMV_BATCH; u0 += 4) {
            if (u0 < nrem) {
              const double *qb = q_l + (size_t)(b + u0) * 16;
              double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
              if (bvalid) {
                q0 = qb[0];
                if (u0 + 1 < nrem) q1 = qb[16];
                if (u0 + 2 < nrem) q2 = qb[32];
                if (u0 + 3 < nrem) q3 = qb[48];
              }
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(c0), "+d"(c1) : "d"(x[u0]), "d"(q0));
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(e0), "+d"(e1) : "d"(x[u0 + 1]), "d"(q1));
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(c0), "+d"(c1) : "d"(x[u0 + 2]), "d"(q2));
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(e0), "+d"(e1) : "d"(x[u0 + 3]), "d"(q3));
            }
          }
        }
        c0 += e0;
        c1 += e1;
        // lane (a, k): c0 = Out_j[a, 2k], c1 = Out_j[a, 2k+1]
        if (a < R && k < 2) {
          const int cA = 2 * k, cB = 2 * k + 1;
          const size_t base = (size_t)j * TS + a;
          if (cA < DH) {
            double v = c0;
            if (G != nullptr) v += __ldg(G + base + (size_t)cA * R);
            out[base + (size_t)cA * R] = v;
          }
          if (cB < DH) {
            double v = c1;
            if (G != nullptr) v += __ldg(G + base + (size_t)cB * R);
            out[base + (size_t)cB * R] = v;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);                 // this warp is done with stage s
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Version 2: the pose tiles a group's blocks multiply are staged in shared memory as well, by dedicated gather warps.
//
// ncu on version 1 (profiles/r02_spmv.md): a consumer warp handles one pose tile at a time and waits a full L2 round
// trip for its 8-byte gathers before the first DMMA (long_scoreboard 3.4, 1.6 us per row and warp): latency-, not
// bandwidth-bound at 0.62 of the HBM peak.  Here the streams are decoupled by mbarriers (one CTA of 1024 threads per SM):
//   warp 30, lane 0 : bulk TMA of a group's column indices + row pointers into a ring of NI small slots (ifull / iempty);
//                     this ring is deeper than the data ring, so the indices are there long before they are needed
//   warp 31, lane 0 : bulk TMA of the group's 4x4 blocks into a ring of NST stages (full / empty)
//   NG gather warps : once a group's indices have landed and its stage is free, asynchronous 16-byte copies (cp.async,
//                     SASS LDGSTS; no registers held) of the pose tiles the blocks refer to, block position by block
//                     position, into the stage's tile buffer; completion is reported on xfull by
//                     cp.async.mbarrier.arrive
//   math warps      : TEAMS teams that take the groups in turn; a team waits for the barriers of its group, multiplies
//                     its rows purely from shared memory (DMMA, as in version 1), stores and releases stage and slot.
// A stage is refilled (blocks by TMA, tiles by the gather warps, concurrently) the moment its team releases it, so the
// DRAM stream and the L2 round trips of the gathers overlap the products of the other stages.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TMA2_THREADS = 1024;

template <int BYTES> __device__ __forceinline__ void cp_async(uint32_t dst, const void *src) {
  if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

template <int R, int DH, int BT, int NST, int NI, int NG, int TEAMS> struct Spmv2Layout {
  static constexpr int TS = R * DH;
  static constexpr int IDX_CAP = BT + 8;
  static constexpr int RP_CAP = BT + 12;
  static constexpr size_t SQ = (size_t)NST * BT * 128;
  static constexpr size_t SX = (size_t)NST * BT * TS * 8;
  static constexpr size_t SI = (size_t)NI * (IDX_CAP + RP_CAP) * 4;
  static constexpr size_t BARS = (size_t)(3 * NST + 2 * NI) * 8;
  static constexpr size_t BYTES = SQ + SX + SI + BARS + 128;
};

template <int R, int DH, int BT, int NST, int NI, int NG, int TEAMS>
__global__ void __launch_bounds__(TMA2_THREADS, 1)
    k_spmv_tma2(int ngroups, const int2 *__restrict__ gi, const int *__restrict__ rowptr, const int *__restrict__ bcol,
                const double *__restrict__ bval, const double *__restrict__ X, const double *__restrict__ G,
                double *__restrict__ out) {
  using L = Spmv2Layout<R, DH, BT, NST, NI, NG, TEAMS>;
  constexpr int TS = R * DH;
  constexpr int CH = ((TS * 8) % 16 == 0) ? 16 : 8;    // bytes per asynchronous copy
  constexpr int NCH = TS * 8 / CH;                      // copies per pose tile
  constexpr int IDX_CAP = L::IDX_CAP, RP_CAP = L::RP_CAP;
  constexpr int NWARP = TMA2_THREADS / 32;
  constexpr int NMATH = NWARP - 2 - NG;                 // warps [0, NMATH) multiply, [NMATH, NMATH+NG) gather, then 2 TMA warps
  constexpr int NMT = NMATH / TEAMS;                    // math warps per team
  static_assert(NMATH % TEAMS == 0, "math warps must split evenly into teams");
  static_assert(NST % TEAMS == 0 && NI % TEAMS == 0, "a stage / slot is always used by the same team");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sq = reinterpret_cast<double *>(smem_raw);                                      // NST * BT * 16
  double *sx = sq + (size_t)NST * BT * 16;                                                // NST * BT * TS
  int *sidx = reinterpret_cast<int *>(sx + (size_t)NST * BT * TS);                        // NI * IDX_CAP
  int *srp = sidx + NI * IDX_CAP;                                                         // NI * RP_CAP
  uint64_t *full = reinterpret_cast<uint64_t *>(srp + NI * RP_CAP);
  uint64_t *xfull = full + NST;
  uint64_t *empty = xfull + NST;
  uint64_t *ifull = empty + NST;
  uint64_t *iempty = ifull + NI;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&xfull[s], NG * 32);
      mbar_init(&empty[s], NMT);
    }
    for (int s = 0; s < NI; ++s) {
      mbar_init(&ifull[s], 1);
      mbar_init(&iempty[s], NG + NMT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NWARP - 1) {
    // ---------------- TMA producer of the 4x4 blocks ----------------
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int it = 0;
      for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
        const int s = it % NST;
        if (it >= NST) mbar_wait(&empty[s], (unsigned)(((it / NST) - 1) & 1));
        const int b0 = __ldg(gi + g).y, b1 = __ldg(gi + g + 1).y;
        const unsigned qbytes = (unsigned)(b1 - b0) * 128u;
        mbar_expect_tx(&full[s], qbytes);
        if (qbytes) tma_load_1d(sq + (size_t)s * BT * 16, bval + (size_t)b0 * 16, qbytes, &full[s], pol);
      }
    }
    return;
  }
  if (warp == NWARP - 2) {
    // ---------------- TMA producer of the indices and row pointers ----------------
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int it = 0;
      for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
        const int si = it % NI;
        if (it >= NI) mbar_wait(&iempty[si], (unsigned)(((it / NI) - 1) & 1));
        const int2 g0 = __ldg(gi + g), g1 = __ldg(gi + g + 1);
        const int b0 = g0.y, b1 = g1.y, r0 = g0.x, r1 = g1.x;
        const int ib = b0 & ~3;                                   // 16-byte aligned index window
        const unsigned ibytes = (unsigned)(((b1 - ib) + 3) & ~3) * 4u;
        const int rb = r0 & ~3;                                   // rowptr[r0 .. r1] inclusive
        const unsigned rbytes = (unsigned)(((r1 + 1 - rb) + 3) & ~3) * 4u;
        mbar_expect_tx(&ifull[si], ibytes + rbytes);
        tma_load_1d(sidx + si * IDX_CAP, bcol + ib, ibytes, &ifull[si], pol);
        tma_load_1d(srp + si * RP_CAP, rowptr + rb, rbytes, &ifull[si], pol);
      }
    }
    return;
  }

  if (warp >= NMATH) {
    // ---------------- gather warps ----------------
    const char *Xb = reinterpret_cast<const char *>(X);
    const int gl = (warp - NMATH) * 32 + lane;
    int it = 0;
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
      const int s = it % NST, si = it % NI;
      const int b0 = __ldg(gi + g).y;
      const int nch = (__ldg(gi + g + 1).y - b0) * NCH;
      const int *idx_s = sidx + si * IDX_CAP + (b0 - (b0 & ~3));
      const uint32_t xs = smem_u32(sx + (size_t)s * BT * TS);
      mbar_wait(&ifull[si], (unsigned)((it / NI) & 1));
      if (it >= NST) mbar_wait(&empty[s], (unsigned)(((it / NST) - 1) & 1));
      for (int c = gl; c < nch; c += NG * 32) {
        const int u = c / NCH, part = c - u * NCH;
        cp_async<CH>(xs + (uint32_t)(c * CH), Xb + ((size_t)idx_s[u] * (TS * 8) + (size_t)(part * CH)));
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&xfull[s])) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&iempty[si]);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }

  // ---------------- math warps ----------------
  const int team = warp % TEAMS, wm = warp / TEAMS;
  const int a = lane >> 2, k = lane & 3;
  const bool valid = (a < R) && (k < DH);
  const int off = k * R + a;
  const int bn = lane >> 2;
  const bool bvalid = bn < 4;
  const int boff = k * 4 + bn;
  const int gstep = (int)gridDim.x * TEAMS;
  int it = team;
  int g = blockIdx.x + team * (int)gridDim.x;
  int2 g0 = make_int2(0, 0), g1 = make_int2(0, 0);
  if (g < ngroups) { g0 = __ldg(gi + g); g1 = __ldg(gi + g + 1); }
  for (; g < ngroups; g += gstep, it += TEAMS) {
    const int s = it % NST, si = it % NI;
    const int r0 = g0.x, r1 = g1.x, b0 = g0.y;
    if (g + gstep < ngroups) { g0 = __ldg(gi + g + gstep); g1 = __ldg(gi + g + gstep + 1); }   // off the critical path
    const double *q_l = sq + (size_t)s * BT * 16 + boff;
    const double *x_l = sx + (size_t)s * BT * TS + off;
    const int *rp_s = srp + si * RP_CAP + (r0 - (r0 & ~3));
    const int wrot = (wm + (it / TEAMS) * 7) % NMT;
    mbar_wait(&ifull[si], (unsigned)((it / NI) & 1));
    mbar_wait(&full[s], (unsigned)((it / NST) & 1));
    mbar_wait(&xfull[s], (unsigned)((it / NST) & 1));
    for (int j = r0 + wrot; j < r1; j += NMT) {
      const int lb0 = rp_s[j - r0] - b0, lb1 = rp_s[j - r0 + 1] - b0;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      for (int b = lb0; b < lb1; b += 4) {
        const int nrem = lb1 - b;                            // warp-uniform
        const double *xb = x_l + (size_t)b * TS;
        const double *qb = q_l + (size_t)b * 16;
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
        if (valid) {
          x0 = xb[0];
          if (nrem > 1) x1 = xb[TS];
          if (nrem > 2) x2 = xb[2 * TS];
          if (nrem > 3) x3 = xb[3 * TS];
        }
        if (bvalid) {
          q0 = qb[0];
          if (nrem > 1) q1 = qb[16];
          if (nrem > 2) q2 = qb[32];
          if (nrem > 3) q3 = qb[48];
        }
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(x0), "d"(q0));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(e0), "+d"(e1) : "d"(x1), "d"(q1));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(x2), "d"(q2));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(e0), "+d"(e1) : "d"(x3), "d"(q3));
      }
      c0 += e0;
      c1 += e1;
      if (a < R && k < 2) {
        const int cA = 2 * k, cB = 2 * k + 1;
        const size_t base = (size_t)j * TS + a;
        if (cA < DH) {
          double v = c0;
          if (G != nullptr) v += __ldg(G + base + (size_t)cA * R);
          out[base + (size_t)cA * R] = v;
        }
        if (cB < DH) {
          double v = c1;
          if (G != nullptr) v += __ldg(G + base + (size_t)cB * R);
          out[base + (size_t)cB * R] = v;
        }
      }
    }
    __syncwarp();
    if (lane == 0) { mbar_arrive(&empty[s]); mbar_arrive(&iempty[si]); }
  }
}

template <int R, int DH, int BT, int NST, int NI, int NG, int TEAMS>
static cudaError_t launch_tma2_t(int ngroups, const int2 *gi, const int *rowptr, const int *bcol, const double *bval,
                                 const double *X, const double *G, double *out, int sms, cudaStream_t stream) {
  using L = Spmv2Layout<R, DH, BT, NST, NI, NG, TEAMS>;
  static_assert(L::BYTES <= 232448, "stage layout exceeds the shared memory of an SM");
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_spmv_tma2<R, DH, BT, NST, NI, NG, TEAMS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)L::BYTES);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  int grid = sms;
  if (grid > ngroups) grid = ngroups;
  if (grid < 1) grid = 1;
  k_spmv_tma2<R, DH, BT, NST, NI, NG, TEAMS><<<grid, TMA2_THREADS, L::BYTES, stream>>>(ngroups, gi, rowptr, bcol, bval, X, G, out);
  return cudaGetLastError();
}

static int spmv_cfg() {
  static const int c = [] { const char *e = std::getenv("DPGO_SPMV_CFG"); return e ? std::atoi(e) : 0; }();
  return c;
}

// blocks (and rows) per row group the host cuts the pose tiles into (dpgo_capi.cu, set_Q)
int spmv_group_blocks() {
  static const bool v1 = [] { const char *e = std::getenv("DPGO_SPMV_V1"); return e && std::atoi(e) != 0; }();
  if (v1) return SPMV_GROUP_BLOCKS;
  switch (spmv_cfg()) {
    case 1: case 3: case 4: return 128;
    case 2: case 6: return 96;
    default: return 192;
  }
}

template <int R, int DH, int NST = TMA_NSTAGE, int PF = 0, int LEAN = 0>
static cudaError_t launch_tma_t(int ngroups, const int2 *gi, const int *rowptr, const int *bcol, const double *bval,
                                const double *X, const double *G, double *out, int sms, cudaStream_t stream) {
  constexpr int BT = SPMV_GROUP_BLOCKS;
  const size_t smem = (size_t)NST * BT * 128 + (size_t)NST * (BT + 8) * 4 + (size_t)NST * (BT + 12) * 4 + 2 * NST * 8 + 128;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_spmv_tma<R, DH, BT, NST, PF, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (PF != 0) {
      // leave the rest of the SM's unified memory to L1: the prefetched pose tiles live there
      e = cudaFuncSetAttribute(k_spmv_tma<R, DH, BT, NST, PF, LEAN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                               (int)((2 * smem * 100 + 228 * 1024 - 1) / (228 * 1024)));
      if (e != cudaSuccess) return e;
    }
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  int grid = 2 * sms;
  if (grid > ngroups) grid = ngroups;
  if (grid < 1) grid = 1;
  k_spmv_tma<R, DH, BT, NST, PF, LEAN><<<grid, TMA_THREADS, smem, stream>>>(ngroups, gi, rowptr, bcol, bval, X, G, out);
  return cudaGetLastError();
}

cudaError_t launch_spmv_tma(int r, int dh, int ngroups, const int2 *gi, const int *rowptr, const int *bcol,
                            const double *bval, const double *X, const double *G, double *out, int sms,
                            cudaStream_t stream) {
  // DPGO_SPMV_V1=1: the register-gather kernel of round 1 (A/B switch)
  static const bool v1 = [] { const char *e = std::getenv("DPGO_SPMV_V1"); return e && std::atoi(e) != 0; }();
  cudaError_t e = cudaErrorInvalidValue;
  if (!v1 && r == 5 && dh == 4 && spmv_cfg() != 0) {
    switch (spmv_cfg()) {
      case 1: return launch_tma2_t<5, 4, 128, 6, 8, 8, 2>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 2: return launch_tma2_t<5, 4, 96, 8, 12, 8, 2>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 3: return launch_tma2_t<5, 4, 128, 6, 8, 6, 2>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 4: return launch_tma2_t<5, 4, 128, 6, 8, 8, 1>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 5: return launch_tma2_t<5, 4, 192, 4, 6, 8, 2>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 6: return launch_tma2_t<5, 4, 96, 8, 12, 6, 2>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 20: return launch_tma_t<5, 4, 4, 0, 1>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 10: return launch_tma_t<5, 4, 3, 0>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 11: return launch_tma_t<5, 4, 3, 1>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 12: return launch_tma_t<5, 4, 3, 2>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 13: return launch_tma_t<5, 4, 4, 1>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      case 14: return launch_tma_t<5, 4, 4, 2>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
      default: break;
    }
  }
#define DPGO_SPMV_CASE(RR, DD)                                                                              \
  if (r == RR && dh == DD)                                                                                  \
    e = v1 ? launch_tma_t<RR, DD>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream)                 \
           : launch_tma2_t<RR, DD, SPMV_GROUP_BLOCKS, 4, 6, 8, 1>(ngroups, gi, rowptr, bcol, bval, X, G, out, sms, stream);
  DPGO_SPMV_CASE(3, 4)
  DPGO_SPMV_CASE(4, 4)
  DPGO_SPMV_CASE(5, 4)
  DPGO_SPMV_CASE(2, 3)
  DPGO_SPMV_CASE(3, 3)
  DPGO_SPMV_CASE(5, 3)
#undef DPGO_SPMV_CASE
  return e;
}

}  // namespace dpgo
