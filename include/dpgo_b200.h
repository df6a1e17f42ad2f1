/*
 * dpgo_b200.h -- C ABI of the B200-native pose-graph-optimisation hot path.
 *
 * This is the drop-in boundary: every entry point replaces one piece of the reference's
 * (tjcunhao/dpo, fork of mit-acl/dpgo) C++ interface for the per-iteration path.  The
 * reference has no C ABI of its own; the C++ mirror under include/DPGO/ (same class names and
 * signatures as the reference) and the Python mirror dpo_b200/ are thin hosts over this file.
 * "ref:" comments cite the reference interface each function stands in for (paths relative
 * to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; all functions return a status code
 *     (DPGO_OK == 0) and never throw or abort.  dpgo_last_error() gives a thread-local message.
 *   - dense arrays are COLUMN-MAJOR r x (d+1)n doubles; pose i is the contiguous r x (d+1)
 *     tile at columns [(d+1)i, (d+1)(i+1)) (layout pinned by ref tests/testEigenMap.cpp:12-36).
 *   - "host" pointers are caller-owned CPU memory (the library copies in/out);
 *     "dev" pointers are CUDA device memory on the problem's device.
 *   - one opaque handle <-> one GPU <-> one CUDA stream; a handle is used by one thread at a
 *     time (ref: a QuadraticProblem is used by one thread at a time, src/PGOAgent.cpp:676-682).
 *   - all arithmetic is fp64.  There is NO CPU fallback: without a usable CUDA device
 *     dpgo_problem_create() fails with DPGO_ERR_NO_DEVICE.
 */
#ifndef DPGO_B200_H
#define DPGO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPGO_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define DPGO_API __attribute__((visibility("default")))
#else
#define DPGO_API
#endif

/* ---- status codes ------------------------------------------------------------------ */
enum {
  DPGO_OK = 0,
  DPGO_ERR_INVALID_ARG = 1, /* shape / pointer / range violation (ref: assert() on shapes,
                               src/QuadraticProblem.cpp:32-33,51-52) */
  DPGO_ERR_NO_DEVICE = 2,   /* no CUDA device / device index out of range */
  DPGO_ERR_CUDA = 3,        /* a CUDA runtime call or kernel failed */
  DPGO_ERR_STATE = 4,       /* call order violation (e.g. optimise before set_Q) */
  DPGO_ERR_UNSUPPORTED = 5, /* d not in {2,3}; r outside the compiled set (d=3: 3..5, d=2: 2,3,5); N too large for
                               the dense preconditioner */
  DPGO_ERR_ALLOC = 6
};

/* ---- enums mirrored from the reference ----------------------------------------------- */
/* ref: include/DPGO/DPGO_types.h:29-35 (ROPTALG) */
enum { DPGO_ALG_RTR = 0, DPGO_ALG_RGD = 1 };

/* Preconditioner used inside truncated CG.
 * ref: src/QuadraticProblem.cpp:31-42,75-87: exact solve with Q + 0.1 I (CHOLMOD) followed by
 * tangent projection.  DENSE_EXACT applies the same operator through a dense inverse resident
 * in HBM (per-iteration trace parity with the reference); BLOCK_JACOBI is the SpMV-only
 * throughput mode (same fixed points, different inner iterates); NONE is projection only.
 * SPARSE_EXACT applies the same operator as DENSE_EXACT through a nested-dissection block factorisation (dense
 * Schur-complement blocks on 2-4 macro levels, L2-resident for sphere2500-sized agents; O(n log n)-ish memory instead
 * of O(n^2)): the default, and what "exact" means everywhere below. */
enum { DPGO_PRECOND_NONE = 0, DPGO_PRECOND_BLOCK_JACOBI = 1, DPGO_PRECOND_DENSE_EXACT = 2, DPGO_PRECOND_SPARSE_EXACT = 3 };

/* ref: ROPTLIB tCGstatusSet as recorded by src/QuadraticOptimizer.cpp:115 */
enum {
  DPGO_TCG_NEGCURVTURE = 0, DPGO_TCG_EXCREGION = 1, DPGO_TCG_LCON = 2, DPGO_TCG_SCON = 3,
  DPGO_TCG_MAXITER = 4, DPGO_TCG_NOT_RUN = -1
};

typedef struct dpgo_problem dpgo_problem_t; /* opaque: QuadraticProblem + manifold + device state */

/* Solver knobs.  ref: include/DPGO/QuadraticOptimizer.h:36-70 setters; defaults
 * src/QuadraticOptimizer.cpp:20-29 (RTR, step 1e-3, 1 iteration, tol 1e-2, radius 10, 50 inner). */
typedef struct dpgo_opt_params {
  int32_t algorithm;          /* DPGO_ALG_* */
  int32_t tr_iterations;      /* setTrustRegionIterations */
  int32_t tr_max_inner;       /* setTrustRegionMaxInnerIterations */
  int32_t precond;            /* DPGO_PRECOND_* (must have been prepared by set_Q) */
  double rgd_stepsize;        /* setGradientDescentStepsize */
  double tr_tolerance;        /* setTrustRegionTolerance */
  double tr_initial_radius;   /* setTrustRegionInitialRadius */
} dpgo_opt_params_t;

/* ref: include/DPGO/DPGO_types.h:40-59 (ROPTResult) + bookkeeping counters. */
typedef struct dpgo_opt_result {
  int32_t success;
  int32_t tcg_status;        /* status of the last tCG solve */
  int32_t tcg_iterations;    /* inner iterations summed over all attempts */
  int32_t outer_iterations;  /* RTR attempts executed (accepted + rejected) */
  int32_t rejections;        /* rejected attempts */
  int32_t spmv_passes;       /* passes over Q executed inside the call */
  int32_t precond_applies;   /* applications of the tCG preconditioner M^-1 */
  int32_t reserved0;
  double f_init, gradnorm_init, f_opt, gradnorm_opt, relative_change, elapsed_ms;
  double quad_init, lin_init; /* <XQ,X> and <X,G> at the input point: f = quad/2 + lin; summed over agents,
                                 (quad + lin)/2 is the centralised cost (shared-edge cross terms count once) */
} dpgo_opt_result_t;

/* ---- library / device ---------------------------------------------------------------- */
DPGO_API int dpgo_abi_version(void);
DPGO_API const char *dpgo_last_error(void);
DPGO_API int dpgo_device_count(int *count);
DPGO_API void dpgo_opt_params_default(dpgo_opt_params_t *p);

/* ---- problem lifetime.  ref: QuadraticProblem ctor/dtor, include/DPGO/QuadraticProblem.h:33-35 */
DPGO_API int dpgo_problem_create(int n, int d, int r, int device, dpgo_problem_t **out);
DPGO_API int dpgo_problem_destroy(dpgo_problem_t *p);
/* Run the handle's work on a caller-provided cudaStream_t (e.g. torch's current stream). NULL
 * restores the handle's own stream. */
DPGO_API int dpgo_problem_set_stream(dpgo_problem_t *p, void *cuda_stream);
DPGO_API int dpgo_problem_sync(dpgo_problem_t *p);
DPGO_API int dpgo_problem_dims(const dpgo_problem_t *p, int *n, int *d, int *r, int64_t *num_blocks);
/* How the persistent step kernel of this handle is launched; takes effect at the next set_Q (it sizes the row partition
 * and the block-solve plan).  0: cooperative grid over all SMs (default:
 * one agent owns the GPU, as in the reference's one-agent-per-machine deployment, examples/MultiRobotExample.cpp:229-334);
 * 1: ONE thread-block cluster of <= 16 CTAs (non-cooperative launch; hardware cluster barriers end the phases), so that the
 * steps of several small agents of one colour class run side by side on one GPU (dpgo_agents_round_async); -1: default. */
DPGO_API int dpgo_problem_set_launch_mode(dpgo_problem_t *p, int mode);
DPGO_API int dpgo_problem_launch_info(const dpgo_problem_t *p, int *grid, int *cluster);

/* ---- cost matrices ------------------------------------------------------------------- */
/* ref: QuadraticProblem::setQ(const SparseMatrix&), src/QuadraticProblem.cpp:31-42.
 * Q is the scalar row-major CSR exactly as Eigen::SparseMatrix<double,RowMajor> exposes it
 * (outerIndexPtr / innerIndexPtr / valuePtr), (d+1)n x (d+1)n, symmetric.  The library converts
 * it to (d+1)x(d+1) block-CSR in HBM and prepares the preconditioners named in precond_mask
 * (bit i = DPGO_PRECOND_i). */
DPGO_API int dpgo_problem_set_Q_csr(dpgo_problem_t *p, int nrows, const int32_t *rowptr, const int32_t *colind,
                           const double *values, unsigned precond_mask);
/* Same from block triplets: nb blocks, block k at (brow[k], bcol[k]) holds the (d+1)x(d+1)
 * sub-matrix Q[(d+1)brow.., (d+1)bcol..] row-major in blocks[k*(d+1)^2 ..]; duplicates are summed
 * (what ref constructConnectionLaplacianSE / PGOAgent::constructQMatrix produce,
 * src/DPGO_utils.cpp:264-271, src/PGOAgent.cpp:720-781). */
DPGO_API int dpgo_problem_set_Q_blocks(dpgo_problem_t *p, int64_t nb, const int32_t *brow, const int32_t *bcol,
                              const double *blocks, unsigned precond_mask);
/* Q assembled ON THE DEVICE from raw edge records (ref constructConnectionLaplacianSE, src/DPGO_utils.cpp:199-271, and the
 * shared-edge diagonal terms of PGOAgent::constructQMatrix, src/PGOAgent.cpp:720-781).  m private edges p1 -> p2 (local pose
 * ids), R: m x d x d row-major, t: m x d, kappa / tau / weight: m (weight NULL = 1), fixed_weight: m flags (NULL = none; a
 * fixed edge keeps its weight in dpgo_problem_robust_reweight: the reference's isKnownInlier, e.g. odometry); plus
 * num_static blocks added at (static_pose, static_pose), (d+1)x(d+1) row-major each (already weighted).  The block pattern
 * is built on the host once, the values by k_assemble_Q; re-assembly after a weight change keeps the pattern. */
DPGO_API int dpgo_problem_set_edges(dpgo_problem_t *p, int64_t m, const int32_t *p1, const int32_t *p2, const double *R,
                                    const double *t, const double *kappa, const double *tau, const double *weight,
                                    const int32_t *fixed_weight, int64_t num_static, const int32_t *static_pose,
                                    const double *static_blocks, unsigned precond_mask);
/* Robust re-weighting at the resident iterate (ref PGOAgent::updateLoopClosuresWeights, src/PGOAgent.cpp:1181-1245;
 * RobustCost::weight, src/DPGO_robust.cpp:23-66): w_e = weight(sqrt(kappa |Y_i R - Y_j|^2 + tau |p_j - p_i - Y_i t|^2)) for
 * every non-fixed edge, then Q is re-assembled on the device and the preconditioners are refreshed.
 * cost: 0 L2, 1 L1, 2 Huber(param), 3 TLS(param), 4 Geman-McClure, 5 GNC_TLS(mu, param = cbar).
 * weights_host / residuals2_host (nullable): the new weights and the squared residuals, m each. */
DPGO_API int dpgo_problem_robust_reweight(dpgo_problem_t *p, int cost, double mu, double param, double *weights_host,
                                          double *residuals2_host);
/* replace the edge weights (m values) and re-assemble Q on the device */
DPGO_API int dpgo_problem_set_edge_weights(dpgo_problem_t *p, const double *weights_host);
/* ref: QuadraticProblem::setG, src/QuadraticProblem.cpp:44-48.  Dense r x (d+1)n column-major,
 * or the reference's sparse form (row-major CSR with r rows).  NULL / nnz == 0 clears G. */
DPGO_API int dpgo_problem_set_G_dense(dpgo_problem_t *p, const double *G_host);
DPGO_API int dpgo_problem_set_G_csr(dpgo_problem_t *p, const int32_t *rowptr, const int32_t *colind,
                           const double *values);

/* ---- evaluation (host in / host out) --------------------------------------------------- */
/* ref: QuadraticProblem::f, src/QuadraticProblem.cpp:50-60 */
DPGO_API int dpgo_problem_f(dpgo_problem_t *p, const double *X_host, double *f_out);
/* ref: QuadraticProblem::EucGrad, :62-66  (Out = X Q + G) */
DPGO_API int dpgo_problem_egrad(dpgo_problem_t *p, const double *X_host, double *out_host);
/* ref: QuadraticProblem::EucHessianEta, :68-73  (Out = V Q) */
DPGO_API int dpgo_problem_ehess(dpgo_problem_t *p, const double *V_host, double *out_host);
/* ref: QuadraticProblem::RieGrad / RieGradNorm, :89-101.  Either output may be NULL. */
DPGO_API int dpgo_problem_rgrad(dpgo_problem_t *p, const double *X_host, double *out_host, double *norm_out);
/* f, Riemannian gradient norm in ONE pass over Q (what optimize() needs before/after) */
DPGO_API int dpgo_problem_f_rgradnorm(dpgo_problem_t *p, const double *X_host, double *f_out, double *norm_out);
/* Riemannian Hessian-vector product at X: ROPTLIB Problem::HessianEta = EucHessianEta +
 * Stiefel::EucHvToHv + projection (call site src/QuadraticOptimizer.cpp:76-119). */
DPGO_API int dpgo_problem_rhess(dpgo_problem_t *p, const double *X_host, const double *V_host, double *out_host);
/* ref: QuadraticProblem::PreConditioner, :75-87 */
DPGO_API int dpgo_problem_precon(dpgo_problem_t *p, int precond, const double *X_host, const double *V_host,
                        double *out_host);

/* ---- manifold (St(d,r) x R^r)^n ------------------------------------------------------- */
/* ROPTLIB ProductManifold::Projection (tangent projection at X), call sites
 * src/QuadraticProblem.cpp:82,95, src/QuadraticOptimizer.cpp:139 */
DPGO_API int dpgo_manifold_tangent_project(dpgo_problem_t *p, const double *X_host, const double *Z_host,
                                  double *out_host);
/* ROPTLIB ProductManifold::Retraction (QF per pose), src/QuadraticOptimizer.cpp:146 */
DPGO_API int dpgo_manifold_retract(dpgo_problem_t *p, const double *X_host, const double *eta_host,
                          double *out_host);
/* ref: LiftedSEManifold::project, src/manifold/LiftedSEManifold.cpp:34-45 (polar factor per pose) */
DPGO_API int dpgo_manifold_project(dpgo_problem_t *p, const double *M_host, double *out_host);

/* ---- optimiser ------------------------------------------------------------------------ */
/* ref: QuadraticOptimizer::optimize(const Matrix&), src/QuadraticOptimizer.cpp:34-59: one
 * RTR (Riemannian trust region, truncated-CG inner solve) or RGD call, the whole loop on the
 * device in one persistent kernel; host sees only X_out and the result record. */
DPGO_API int dpgo_optimize(dpgo_problem_t *p, const dpgo_opt_params_t *params, const double *X_in_host,
                  double *X_out_host, dpgo_opt_result_t *result);

/* ---- device-resident path (iterate lives in HBM between calls) --------------------------- */
DPGO_API int dpgo_problem_upload_X(dpgo_problem_t *p, const double *X_host);
DPGO_API int dpgo_problem_download_X(dpgo_problem_t *p, double *X_host);
/* the same without the closing synchronisation (pinned host buffers; order with dpgo_problem_sync): one round of a
 * multi-agent run = upload_X_async, exchange, optimize_resident_async, download_X_async, sync */
DPGO_API int dpgo_problem_upload_X_async(dpgo_problem_t *p, const double *X_host);
DPGO_API int dpgo_problem_download_X_async(dpgo_problem_t *p, double *X_host);
/* resident iterate <- device buffer (asynchronous device-to-device copy on the handle's stream) */
DPGO_API int dpgo_problem_copy_X_from_device(dpgo_problem_t *p, const double *X_dev);
DPGO_API int dpgo_problem_device_X(dpgo_problem_t *p, double **X_dev);     /* r x (d+1)n, read/write */
DPGO_API int dpgo_problem_device_G(dpgo_problem_t *p, double **G_dev);
/* optimise the resident iterate in place; asynchronous on the handle's stream */
DPGO_API int dpgo_optimize_resident_async(dpgo_problem_t *p, const dpgo_opt_params_t *params);
/* wait for the last async optimise and fetch its result record */
DPGO_API int dpgo_optimize_result(dpgo_problem_t *p, dpgo_opt_result_t *result);
/* the Q.X product kernel alone on device buffers (the roofline kernel): Out = X Q (+ G) */
DPGO_API int dpgo_spmv_device(dpgo_problem_t *p, const double *X_dev, double *out_dev, int add_G);
DPGO_API int64_t dpgo_spmv_algorithmic_bytes(const dpgo_problem_t *p, int add_G);
/* bytes one application of the preconditioner (ref: QuadraticProblem::PreConditioner, src/QuadraticProblem.cpp:75-87)
 * has to move: the operator's unique data (upper triangle of the symmetric dense inverse when the symmetric
 * kernel is planned, i.e. after the first exact-mode optimise; the full matrix otherwise) + input and output vector */
DPGO_API int64_t dpgo_precond_algorithmic_bytes(const dpgo_problem_t *p, int preconditioner);
/* host only (no device needed): the work decomposition the symmetric dense apply uses for an N x N operator on `grid`
 * CTAs -- column segments of 480, 8-row groups, chunks in segment-major order.  segptr[nseg+1] = first chunk of every
 * segment, cut[grid+1] = chunk range of every CTA, cfirst/ccount[nseg] = the consecutive CTAs that touch a segment
 * (= partial-panel slots of its columns), chunk_offset[nchunks+1] = first double of every chunk in the packed upper
 * triangle.  chunk_cost <= 0 selects the built-in cost model.  DPGO_ERR_UNSUPPORTED when no plan exists for this size
 * (odd N, N < 2048, fewer chunks than CTAs); the library then streams the full matrix. */
DPGO_API int dpgo_sym_plan_sizes(int N, int *num_segments, int *num_chunks);
DPGO_API int dpgo_sym_plan(int N, int grid, double chunk_cost, int32_t *segptr, int32_t *cut, int32_t *cfirst,
                           int32_t *ccount, int64_t *chunk_offset);
/* ---- sparse exact preconditioner: diagnostics ------------------------------------------------------------------ */
/* info[16] of the prepared hierarchy (prepares it if needed): 0 macro levels, 1 macro nodes, 2 phases per application,
 * 3 bytes of all blocks, 4 matrix bytes streamed per application, 5 largest own block (scalars), 6 largest boundary
 * (scalars), 7 dissection depth, 8 steps, 9 jobs, 10 epilogues, 11.. reserved */
DPGO_API int dpgo_nd_info(dpgo_problem_t *p, int64_t *info16);
/* HOST ONLY, verification of the planning code on machines without a GPU (never used by a product path): builds the
 * hierarchy, the blocks and the phase plan for the block matrix given as in dpgo_problem_set_Q_blocks and runs a host
 * emulation of the plan exactly as the kernel interprets it:  Z = (Q + shift I)^-1 V  (no projection), V and Z
 * r x (d+1)n column-major.  force_cuts < 0 lets the cost model choose the macro levels.  info16 as in dpgo_nd_info. */
DPGO_API int dpgo_nd_debug_emulate(int n, int d, int r, int64_t nb, const int32_t *brow, const int32_t *bcol,
                                   const double *blocks, double shift, int grid, int force_cuts, int leaf_size,
                                   const double *V_host, double *Z_host, int64_t *info16);
/* diagnostic: cost of one empty phase of the persistent kernel (grid barrier + scalar reduction) and of its launch */
DPGO_API int dpgo_debug_phase_latency(dpgo_problem_t *p, int phases, double *us_per_phase, double *us_launch);
/* diagnostic: phase clock of the persistent kernel.  enable != 0 switches it on (subsequent optimise calls
 * accumulate, per phase kind, the nanoseconds CTA 0 spent up to the closing grid barrier); every call returns the
 * accumulated milliseconds in ms_by_kind[8] (0 eval pass, 1 dense preconditioner apply, 2 partial sums + projection,
 * 3 Hessian product, 4 tCG update, 5 retraction, 6 final, 7 unused) and resets them; enable == 0 switches it off. */
DPGO_API int dpgo_debug_phase_times(dpgo_problem_t *p, int enable, double *ms_by_kind);
/* same with 32 slots: 0..7 as above (1 = the whole exact-preconditioner application when the dense inverse is used),
 * 8 + k = phase k of the sparse exact preconditioner's application (k < 16), 24 / 25 / 26 = gathers / panel jobs /
 * epilogues of those phases as seen by CTA 0, others unused */
DPGO_API int dpgo_debug_phase_times32(dpgo_problem_t *p, int enable, double *ms_by_kind);
/* same with 64 slots: 32 + 3 k + {0, 1, 2} = gathers / panel jobs / epilogues of phase k (k < 10) as seen by CTA 0 */
DPGO_API int dpgo_debug_phase_times64(dpgo_problem_t *p, int enable, double *ms_by_kind);

/* ---- chordal initialisation on the GPU ------------------------------------------------------------------------
 * ref: chordalInitialization, src/DPGO_utils.cpp:273-461 (two sparse least-squares problems, SPQR there; gauge R_0 = I,
 * t_0 = 0) + projectToRotationGroup :463-477.  Both normal systems are 3 x 3-block connection Laplacians, solved by
 * Jacobi-preconditioned conjugate gradients whose product is the TMA-fed block-CSR kernel of the hot path.
 * m edges p1 -> p2 (pose ids), R: m x d x d row-major, t: m x d, kappa / tau: m.  T_host: d x (d+1)n column-major
 * ([R_p t_p] per pose, the layout of the reference's Matrix).  tol: relative residual (<= 0: 1e-11); iterations2[2]
 * (nullable) receives the CG iteration counts of the two solves.  dpgo_chordal_last_error() for the message. */
DPGO_API int dpgo_chordal_initialization(int n, int d, int64_t m, const int32_t *p1, const int32_t *p2, const double *R,
                                         const double *t, const double *kappa, const double *tau, int device, double tol,
                                         int max_iter, double *T_host, int32_t *iterations2);
DPGO_API const char *dpgo_chordal_last_error(void);

/* ---- plain device helpers for hosts that drive several GPUs without linking the CUDA runtime themselves (the C++
 *      multi-GPU runner: exchange buffers + one stream per GPU, NCCL calls on those streams) ---------------------- */
DPGO_API int dpgo_device_set(int device);                                /* cudaSetDevice for the calling thread */
DPGO_API int dpgo_device_malloc(int device, size_t bytes, void **ptr);   /* zero-initialised */
DPGO_API int dpgo_device_free(int device, void *ptr);
DPGO_API int dpgo_stream_create(int device, void **cuda_stream);
DPGO_API int dpgo_stream_destroy(int device, void *cuda_stream);
DPGO_API int dpgo_stream_synchronize(int device, void *cuda_stream);

/* ---- boundary-pose exchange (multi-agent, one agent per GPU) ----------------------------- */
/* ref: PGOAgent::getSharedPoseDict, src/PGOAgent.cpp:95-105: register which local poses are
 * public; pack gathers their r x (d+1) tiles into a contiguous device buffer (the NCCL
 * all-gather send buffer), slot s <- pose public_pose[s]. */
DPGO_API int dpgo_agent_set_public_poses(dpgo_problem_t *p, int num_public, const int32_t *public_pose);
DPGO_API int dpgo_agent_pack_public(dpgo_problem_t *p, double *send_dev);
/* ref: PGOAgent::constructGMatrix, src/PGOAgent.cpp:783-859.  Shared edge e touches local pose
 * local_pose[e]; its neighbour pose is tile nbr_slot[e] of the gathered buffer; outgoing[e]!=0
 * means this agent owns the edge tail (G_p1 += -X_j Om T^T) else the head (G_p2 += -X_i T Om).
 * T is (d+1)x(d+1) row-major per edge, omega the (d+1) diagonal weights (kappa..,tau)*weight. */
DPGO_API int dpgo_agent_set_shared_edges(dpgo_problem_t *p, int num_edges, const int32_t *local_pose,
                                const int32_t *nbr_slot, const int32_t *outgoing, const double *T,
                                const double *omega);
/* rebuild G in HBM from the gathered neighbour tiles (deterministic: edges grouped per pose) */
DPGO_API int dpgo_agent_build_G(dpgo_problem_t *p, const double *gathered_dev, int64_t num_slots);
/* ---- Nesterov-accelerated RBCD on the resident iterate (ref src/PGOAgent.cpp:685-695 iterate, :1040-1091 updateGamma /
 *      updateAlpha / updateY / updateV / restart; the scalars gamma, alpha stay with the host, which follows the reference's
 *      recurrences).  All calls are asynchronous on the handle's stream. */
DPGO_API int dpgo_agent_accel_init(dpgo_problem_t *p);                     /* V = Y = XPrev = X        (ref :60-62, :1040-1052) */
DPGO_API int dpgo_agent_accel_begin(dpgo_problem_t *p, double alpha);       /* XPrev = X; Y = proj((1 - alpha) X + alpha V)  (ref :1077-1083) */
/* optimized == 0: X = Y (ref updateX(false, true), :1095-1098); then V = proj(V + gamma (X - Y))  (ref :1085-1091) */
DPGO_API int dpgo_agent_accel_end(dpgo_problem_t *p, double gamma, int optimized);
DPGO_API int dpgo_agent_accel_restart_begin(dpgo_problem_t *p);             /* X = XPrev  (then the caller takes a plain step) */
DPGO_API int dpgo_agent_accel_restart_end(dpgo_problem_t *p);               /* V = Y = X */
/* public tiles of the auxiliary iterate Y (ref getAuxSharedPoseDict, :107-118) */
DPGO_API int dpgo_agent_pack_public_aux(dpgo_problem_t *p, double *send_dev);
/* X = Y, then optimise X in place (ref updateX(true, true): the step starts from the auxiliary iterate) */
DPGO_API int dpgo_optimize_resident_from_aux_async(dpgo_problem_t *p, const dpgo_opt_params_t *params);
/* One RBCD round of the active agents of one GPU with one call (ref: the body of the round loop,
 * examples/MultiRobotExample.cpp:229-334: updateNeighborPoses -> iterate() -> getSharedPoseDict per selected agent).
 * Every agent works on its own stream between a fork from and a join into main_stream: G rebuild from gathered_dev ->
 * RTR step -> pack of its public tiles into send_dev[i].  main_stream NULL = the stream the first handle is set to.
 * pack_after_join != 0 issues the packs in a second fork/join
 * (needed when neighbouring agents are active in the same round and send_dev aliases gathered_dev). */
DPGO_API int dpgo_agents_round_async(dpgo_problem_t *const *agents, int num_active, const dpgo_opt_params_t *params,
                            const double *gathered_dev, int64_t num_slots, double *const *send_dev, void *main_stream,
                            int pack_after_join);
/* The host boundary of a round with one call per direction (ref: the host matrices PGOAgent::setX / getX move,
 * src/PGOAgent.cpp:66-93): direction 0 = X of every listed agent from (pinned) host memory, then its public tiles packed
 * into send_dev[i] (send_dev may be NULL); direction 1 = X back to host memory.  Asynchronous on `stream` (NULL: the
 * stream the first handle is set to); a repeated call is replayed as a CUDA graph. */
DPGO_API int dpgo_agents_host_io_async(dpgo_problem_t *const *agents, int count, double *const *X_host,
                              double *const *send_dev, int direction, void *stream);
/* per-agent Riemannian gradient norm / cost of the resident iterate (greedy selection input) */
DPGO_API int dpgo_agent_f_rgradnorm_resident(dpgo_problem_t *p, double *f_out, double *norm_out);

#ifdef __cplusplus
}
#endif
#endif /* DPGO_B200_H */
