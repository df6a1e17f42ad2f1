// dpgo_kernels.cuh -- kernel-side parameter block shared by dpgo_kernels.cu and dpgo_capi.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dpgo_b200.h"
#include "nd_precond.h"

namespace dpgo {

// work vectors, each r x (d+1)n fp64 in HBM
enum VecId {
  V_X0 = 0, V_X1,        // iterate double buffer (X0 is the externally visible one)
  V_EG0, V_EG1,          // Euclidean gradient at X0 / X1
  V_RG0, V_RG1,          // Riemannian gradient
  V_Z00, V_Z01,          // preconditioned gradient M^-1 g at X0 / X1
  V_ETA, V_RES, V_Z,     // tCG: step, residual, preconditioned residual
  V_D0, V_D1,            // tCG: search direction double buffer
  V_HD,                  // H[delta]
  V_T,                   // scratch (Stiefel projection output)
  V_XIN,                 // copy of the input iterate (relative change)
  V_AUX,                 // operand of the single-operation entry points
  V_COUNT
};

enum OpCode {
  OP_OPTIMIZE = 0,   // full RTR / RGD call
  OP_EVAL = 1,       // f, EG, RG, |RG| at X0
  OP_RHESS = 2,      // HD = Hess f(X0)[AUX]
  OP_PRECON = 3,     // Z  = P_X0( M^-1 AUX )
  OP_RETRACT = 4,    // X1 = R_X0(AUX)
  OP_PHASE_BENCH = 9, // diagnostic: empty phases
};

constexpr int NRED = 4;            // scalars reduced per phase
constexpr int OPT_THREADS = 512;   // persistent kernel block size
constexpr int SPMV_GROUP_BLOCKS = 192;  // blocks per row group of the TMA-fed SpMV (24 KB of Q per smem stage)
constexpr int ND_YCAP_TILES = 600;  // sparse exact preconditioner: pose tiles of a phase's input vector staged in shared memory per step
constexpr int ND_SLOT_CAP = 240;    // ... and partial-sum slots (8 rows x r doubles) per step
constexpr int ND_SMEM_NEED_SMALL = 16000;   // doubles a small agent's sparse plan needs at most (occupancy query of the cluster launch)
constexpr int SP_CACHE_INTS = 2048;  // shared-memory copy of a CTA's block-CSR structure (row pointers + block columns), 8 KB
constexpr int DENSE_PER_MAX = 512;  // max rows of the dense inverse one CTA owns (smem staging of V): N <= 75k at 148 CTAs

// sparse exact preconditioner (nd_precond.h): the static plan and the panel blob in HBM / L2
struct KNd {
  int nphases;
  int max_ytiles, max_slots;       // shared-memory tiles / partial-sum slots the plan needs per step
  int max_gathers;                 // gather records staged per step (<= max_ytiles)
  int dir[nd::MAX_PHASES];         // per phase: 0 forward (input = residual), 1 backward (input = ancestors' solution)
  int cta0[nd::MAX_PHASES];        // per phase: first (phase, CTA) record
  const nd::CtaPhase *cta_phase;
  const nd::Step *steps;
  const nd::Gather *gathers;
  const nd::Job *jobs;
  const nd::Epi *epis;
  const int *csrc;
  const double *blob;
  double *TX;            // n tiles, permuted order: t (forward) then x (backward), unprojected
  double *C;             // contribution tiles of the forward sweep
};

struct KParams {
  int n;                 // poses
  int N;                 // (d+1) n
  int grid;              // CTAs of the persistent kernel
  int op;
  const int *rowptr;     // n+1
  const int *bcol;       // nb
  const double *bval;    // nb*16
  const double *dinv;    // n*16   block-Jacobi inverse blocks, may be null
  const double *pinv;    // N*N    dense inverse of Q+0.1I, may be null
  double *dense_part;    // grid * r * N  per-CTA partial products of the dense preconditioner
  int dense_per;         // rows of pinv per CTA
  int sym_ok;            // symmetric (upper-triangle) variant of the dense preconditioner is planned
  const double *ppack;   // upper triangle of pinv, chunk-major: chunk = 8 rows x (padded width + 4) doubles, contiguous
  const long long *sym_off; // nchunks+1: first double of every chunk in ppack
  const int *sym_cut;    // grid+1: per-CTA range of chunk indices (segment-major order of (segment, 8-row group))
  const int *sym_segptr; // nseg+1: first chunk index of every column segment
  const int *sym_cfirst; // nseg: first CTA that touches the segment
  const int *sym_ccount; // nseg: number of (consecutive) CTAs that touch it = partial panels of its columns
  double *dense_t2;      // nseg x r x N transposed-product partials, slot = column segment
  const int *cta_rows;   // grid+1 balanced row partition
  const double *G;       // linear term r x N
  double *v[V_COUNT];
  double *S[2];          // n*9   sym(Y^T EG_Y) per pose at X0 / X1
  double *partials;      // 2 * grid * NRED
  unsigned *bar_counter;
  unsigned *bar_epoch;
  dpgo_opt_params_t prm;
  dpgo_opt_result_t *result;   // device copy of the result record
  KNd nd;                // sparse exact preconditioner (nd.nphases == 0: not prepared)
  int cluster;           // 1: the whole grid is ONE thread-block cluster (<= 16 CTAs): phase ends use barrier.cluster
  int strict_acquire;    // 1: the grid barrier polls with ld.acquire (L1 invalidated every phase); 0: relaxed poll (default)
  int smem_doubles;      // dynamic shared memory of this launch, in doubles
  unsigned long long *phase_ns; // diagnostic (nullable): per phase kind, ns seen by CTA 0 (dpgo_debug_phase_times)
};

// launchers (dpgo_kernels.cu)
cudaError_t launch_optimize(int r, int dh, const KParams &kp, cudaStream_t stream);
cudaError_t launch_spmv(int r, int dh, int n, const int *rowptr, const int *bcol, const double *bval,
                        const double *X, const double *G, double *out, cudaStream_t stream);
cudaError_t launch_spmv_tma(int r, int dh, int ngroups, const int2 *groups, const int *rowptr, const int *bcol,
                            const double *bval, const double *X, const double *G, double *out, int sms,
                            cudaStream_t stream);
int spmv_group_blocks();                            // blocks per row group of the TMA-fed SpMV in use
int optimize_max_grid(int r, int dh, int device);   // co-resident CTA count for the persistent kernel
int optimize_max_cluster(int r, int dh, int device); // largest single-cluster grid (16, 8 or 0) the kernel can be launched with
cudaError_t launch_stiefel_project(int r, int dh, int n, const double *M, double *out, cudaStream_t stream, double c0 = 1.0,
                                   const double *B = nullptr, double c1 = 0.0, const double *C = nullptr, double c2 = 0.0);
cudaError_t launch_pack_tiles(int ts, int count, const int *pose, const double *X, double *out, cudaStream_t stream);
cudaError_t launch_build_G(int r, int dh, int nposes, const int *pose_ids, const int *pose_ptr, const int *edge_slot,
                           const int *edge_out, const double *edge_T, const double *edge_om, const double *gathered,
                           double *G, cudaStream_t stream);
cudaError_t launch_assemble_Q(int64_t nb, const int *cptr, const int2 *contrib, const double *eT, const double *eom, const double *ew,
                              const double *sblk, double *bval, cudaStream_t stream);
cudaError_t launch_edge_weights(int r, int dh, int64_t m, const int *p1, const int *p2, const double *eT, const double *eom,
                                const int *fixed, const double *X, int cost, double mu, double param, double *w, double *resid,
                                cudaStream_t stream);
cudaError_t launch_pack_sym(const double *pinv, int N, int nchunks, const int *segptr, int nseg, const long long *off, double *ppack,
                            cudaStream_t stream);
cudaError_t launch_bsr_to_dense(int n, int dh, int64_t nb, const int *rowptr, const int *bcol, const double *bval,
                                double shift, double *A, int N, cudaStream_t stream);

// dense SPD inverse in place (dense_inverse.cu); A is N x N, ld = N, symmetric positive definite
cudaError_t dense_spd_inverse(double *A, int N, cudaStream_t stream);

}  // namespace dpgo
