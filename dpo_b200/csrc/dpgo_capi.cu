// dpgo_capi.cu -- implementation of the C ABI in include/dpgo_b200.h (host side of the library:
// handle management, CSR -> block-CSR conversion, preconditioner setup, H2D/D2H staging, launches).
// No CPU compute fallback exists here: every numeric entry point runs the sm_100a kernels.
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "dpgo_kernels.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

#define DPGO_CUDA(call)                                                                           \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(DPGO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));             \
  } while (0)

#define DPGO_REQUIRE(cond, code, msg) \
  do {                                \
    if (!(cond)) return fail(code, msg); \
  } while (0)

#define DPGO_TRY(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != DPGO_OK) return _s; \
  } while (0)

template <class T> void free_dev(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

}  // namespace

struct dpgo_problem {
  int n = 0, d = 0, r = 0, dh = 0, N = 0, ts = 0;
  int device = 0, sms = 0, grid = 0, max_grid = 0, max_cluster = 0;
  bool cluster = false;          // the persistent kernel runs as ONE thread-block cluster (small agents)
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev_done = nullptr, ev_fork = nullptr;   // fork / join of dpgo_agents_round_async
  uint64_t generation = 0;       // bumped whenever device buffers a captured round refers to may have been replaced
  struct RoundGraph { std::vector<uint64_t> key; cudaGraphExec_t exec = nullptr; int uses = 0; bool failed = false; };
  std::vector<RoundGraph> round_graphs;      // CUDA graphs of dpgo_agents_round_async, kept by the first agent of the round
  int launch_mode = -1;          // -1: by DPGO_CLUSTER_MAX_POSES (default off), 0: full cooperative grid, 1: one thread-block cluster
  // Q in block-CSR
  int64_t nb = 0;
  bool have_Q = false;
  unsigned precond_mask = 0;
  int *d_rowptr = nullptr, *d_bcol = nullptr, *d_cta_rows = nullptr;
  int2 *d_groups = nullptr;      // row groups of the TMA-fed SpMV
  int ngroups = 0;
  double *d_bval = nullptr, *d_dinv = nullptr, *d_pinv = nullptr, *d_dense_part = nullptr, *d_dense_t2 = nullptr;
  int dense_per = 1;
  int sym_ok = 0;                // symmetric (upper-triangle) dense preconditioner planned
  double *d_ppack = nullptr;
  long long *d_sym_off = nullptr;
  int *d_sym_cut = nullptr, *d_sym_segptr = nullptr, *d_sym_cfirst = nullptr, *d_sym_ccount = nullptr;
  // host copy of the block-CSR (lazy preconditioner setup) and the sparse exact preconditioner
  std::vector<int> h_rowptr, h_bcol;
  std::vector<double> h_bval;
  bool nd_ready = false;
  dpgo::KNd nd = {};
  dpgo::nd::Hierarchy *nd_H = nullptr;
  int64_t nd_info[16] = {};
  // edge records for the device-side Q assembly / robust re-weighting (dpgo_problem_set_edges)
  int64_t ne = 0;
  int *d_e_p1 = nullptr, *d_e_p2 = nullptr, *d_e_fixed = nullptr, *d_cptr = nullptr;
  int2 *d_contrib = nullptr;
  double *d_eT = nullptr, *d_eom = nullptr, *d_ew = nullptr, *d_sblk = nullptr, *d_eres = nullptr;
  // vectors
  double *d_G = nullptr;
  double *d_acc[3] = {nullptr, nullptr, nullptr};   // Nesterov acceleration: Y, V, XPrev (allocated by accel_init)
  double *d_vec[dpgo::V_COUNT] = {};
  double *d_S[2] = {nullptr, nullptr};
  double *d_partials = nullptr;
  unsigned *d_bar = nullptr;     // [0] arrival counter, [1] epoch
  unsigned long long *d_phase_ns = nullptr;   // diagnostic phase clock (8 slots), allocated on request
  dpgo_opt_result_t *d_result = nullptr;
  dpgo_opt_result_t *h_result = nullptr;   // pinned
  bool async_pending = false;
  std::chrono::high_resolution_clock::time_point async_t0;
  // exchange
  int num_public = 0;
  int *d_public = nullptr;
  int num_edges = 0, num_shared_poses = 0, max_slot = -1;
  bool G_dirty = true;           // G may hold values that dpgo_agent_build_G does not overwrite
  int *d_pose_ids = nullptr, *d_pose_ptr = nullptr, *d_edge_slot = nullptr, *d_edge_out = nullptr;
  double *d_edge_T = nullptr, *d_edge_om = nullptr;

  size_t vec_bytes() const { return sizeof(double) * (size_t)r * (size_t)N; }
};

namespace {

void fill_kparams(const dpgo_problem *p, dpgo::KParams &kp, int op, const dpgo_opt_params_t &prm) {
  kp.n = p->n;
  kp.N = p->N;
  kp.grid = p->grid;
  kp.op = op;
  kp.rowptr = p->d_rowptr;
  kp.bcol = p->d_bcol;
  kp.bval = p->d_bval;
  kp.dinv = p->d_dinv;
  kp.pinv = p->d_pinv;
  kp.dense_part = p->d_dense_part;
  kp.dense_per = p->dense_per;
  kp.sym_ok = p->sym_ok;
  kp.ppack = p->d_ppack;
  kp.sym_off = p->d_sym_off;
  kp.sym_cut = p->d_sym_cut;
  kp.sym_segptr = p->d_sym_segptr;
  kp.sym_cfirst = p->d_sym_cfirst;
  kp.sym_ccount = p->d_sym_ccount;
  kp.dense_t2 = p->d_dense_t2;
  kp.cta_rows = p->d_cta_rows;
  kp.G = p->d_G;
  for (int i = 0; i < dpgo::V_COUNT; ++i) kp.v[i] = p->d_vec[i];
  kp.S[0] = p->d_S[0];
  kp.S[1] = p->d_S[1];
  kp.partials = p->d_partials;
  kp.bar_counter = p->d_bar;
  kp.bar_epoch = p->d_bar + 1;
  kp.nd = p->nd;
  if (!p->nd_ready) kp.nd.nphases = 0;
  static const int strict = [] { const char *e = std::getenv("DPGO_STRICT_ACQUIRE"); return (e && e[0] == '1') ? 1 : 0; }();
  kp.strict_acquire = strict;
  kp.cluster = p->cluster ? 1 : 0;
  kp.smem_doubles = 0;
  kp.phase_ns = p->d_phase_ns;
  kp.prm = prm;
  kp.result = p->d_result;
}

cudaError_t run_spmv(const dpgo_problem *p, const double *X, const double *G, double *out) {
  static const bool force_gather = [] { const char *e = std::getenv("DPGO_SPMV_KERNEL"); return e && std::string(e) == "gather"; }();
  if (p->ngroups > 0 && !force_gather)
    return dpgo::launch_spmv_tma(p->r, p->dh, p->ngroups, p->d_groups, p->d_rowptr, p->d_bcol, p->d_bval, X, G, out, p->sms,
                                 p->stream);
  return dpgo::launch_spmv(p->r, p->dh, p->n, p->d_rowptr, p->d_bcol, p->d_bval, X, G, out, p->stream);
}

// Work decomposition of the symmetric (upper-triangle) dense apply (phase_dense_sym) -- host only, also exported as
// dpgo_sym_plan so that CPU tests can check it.  Chunks (segment J of 480 columns, 8-row group g with 8g < end of J) in
// segment-major order are cut into `grid` contiguous runs of equal cost; per segment the consecutive CTAs that touch it
// get the partial-panel slots 0..ccount-1; `off` is the chunk-major packed layout (8 rows x (width up to 8, + 4) doubles).
// Cost model of one chunk in column units: streamed width + a fixed part.  Measured on sphere2500 (phase clock, dense
// apply): fixed = 48 -> 99.5 us, 160 -> 93.1, 320 -> 88.2, 720 -> 83.9, 1000 -> 83.7: a chunk costs about the same
// whatever its width (waits, fragment parking, the masked diagonal path), so the runs get nearly equal chunk COUNTS.
constexpr double SYM_CHUNK_COST = 800.0;
constexpr int SYM_PLAN_SEG = 480;      // = dpgo::SYM_SEG of the kernel (15 consumer warps x 32 columns)

struct SymPlan {
  int nseg = 0, nchunks = 0;
  std::vector<int> segptr, cut, cfirst, ccount;
  std::vector<long long> off;
};

bool make_sym_plan(int64_t N, int G, double chunk_cost, SymPlan &pl) {
  if (N % 2 != 0 || N < 2048 || G < 1) return false;       // bulk TMA needs 16-byte aligned rows; small N: plain apply
  const int SEG = SYM_PLAN_SEG, nseg = (int)((N + SEG - 1) / SEG);
  pl.nseg = nseg;
  pl.segptr.assign((size_t)nseg + 1, 0);
  for (int J = 0; J < nseg; ++J) {
    const int s1 = (int)std::min<int64_t>(N, (int64_t)(J + 1) * SEG);
    pl.segptr[(size_t)J + 1] = pl.segptr[(size_t)J] + (s1 + 7) / 8;
  }
  const int nchunks = pl.nchunks = pl.segptr[(size_t)nseg];
  std::vector<double> cum((size_t)nchunks + 1, 0.0);
  pl.off.assign((size_t)nchunks + 1, 0);
  {
    int lin = 0;
    for (int J = 0; J < nseg; ++J) {
      const int s0 = J * SEG, s1 = (int)std::min<int64_t>(N, (int64_t)s0 + SEG);
      for (int g = 0; 8 * g < s1; ++g, ++lin) {
        const int width = s1 - std::max(s0, 8 * g);
        cum[(size_t)lin + 1] = cum[(size_t)lin] + (double)width + chunk_cost;
        pl.off[(size_t)lin + 1] = pl.off[(size_t)lin] + 8LL * (((width + 7) & ~7) + 4);
      }
    }
  }
  pl.cut.assign((size_t)G + 1, 0);
  {
    int lin = 0;
    for (int b = 1; b < G; ++b) {
      const double target = cum[(size_t)nchunks] * b / G;
      while (lin < nchunks && cum[(size_t)lin + 1] <= target) ++lin;
      pl.cut[(size_t)b] = lin;
    }
    pl.cut[(size_t)G] = nchunks;
  }
  pl.cfirst.assign((size_t)nseg, 0);
  pl.ccount.assign((size_t)nseg, 0);
  int maxslots = 0;
  for (int J = 0; J < nseg; ++J) {
    int first = -1, last = -1;
    for (int b = 0; b < G; ++b)
      if (pl.cut[(size_t)b] < pl.segptr[(size_t)J + 1] && pl.cut[(size_t)b + 1] > pl.segptr[(size_t)J] &&
          pl.cut[(size_t)b] < pl.cut[(size_t)b + 1]) {
        if (first < 0) first = b;
        last = b;
      }
    pl.cfirst[(size_t)J] = std::max(first, 0);
    pl.ccount[(size_t)J] = (first < 0) ? 0 : last - first + 1;
    maxslots = std::max(maxslots, pl.ccount[(size_t)J]);
  }
  bool runs_ok = (maxslots <= G);   // the panels live in dense_part (grid x r x N)
  for (int b = 0; b < G; ++b) runs_ok = runs_ok && (pl.cut[(size_t)b] < pl.cut[(size_t)b + 1]);   // slots assume no empty run
  return runs_ok;
}

// The dense inverse (Q + 0.1 I)^-1 is built on first use (one-shot: scatter block-CSR into N x N in HBM, blocked
// Gauss-Jordan in place) -- problems that are only evaluated (e.g. the drivers' centralised problem) never pay for it.
int ensure_dense(dpgo_problem *p) {
  if (p->d_pinv) return DPGO_OK;
  if (!(p->precond_mask & (1u << DPGO_PRECOND_DENSE_EXACT)))
    return fail(DPGO_ERR_STATE, "dense exact preconditioner was not requested in set_Q (precond_mask)");
  const size_t N = (size_t)p->N;
  if (N * N * sizeof(double) > (size_t)48 << 30)
    return fail(DPGO_ERR_UNSUPPORTED, "dense exact preconditioner limited to N^2*8 <= 48 GiB; use block-Jacobi");
  p->dense_per = (int)((N + p->grid - 1) / p->grid);
  if (p->dense_per > dpgo::DENSE_PER_MAX)
    return fail(DPGO_ERR_UNSUPPORTED, "dense exact preconditioner: N too large for the per-CTA slab; use block-Jacobi");
  DPGO_CUDA(cudaMalloc(&p->d_dense_part, sizeof(double) * (size_t)p->grid * p->r * N));
  DPGO_CUDA(cudaMemsetAsync(p->d_dense_part, 0, sizeof(double) * (size_t)p->grid * p->r * N, p->stream));
  DPGO_CUDA(cudaMalloc(&p->d_pinv, N * N * sizeof(double)));
  DPGO_CUDA(cudaMemsetAsync(p->d_pinv, 0, N * N * sizeof(double), p->stream));
  cudaError_t e = dpgo::launch_bsr_to_dense(p->n, p->dh, p->nb, p->d_rowptr, p->d_bcol, p->d_bval, 0.1, p->d_pinv, p->N,
                                            p->stream);
  if (e == cudaSuccess) e = dpgo::dense_spd_inverse(p->d_pinv, p->N, p->stream);
  if (e != cudaSuccess) {
    free_dev(p->d_pinv);
    free_dev(p->d_dense_part);
    return fail(DPGO_ERR_CUDA, std::string("dense preconditioner setup: ") + cudaGetErrorString(e));
  }
  // symmetric (upper-triangle) variant: plan, packed copy, partial buffers (falls back to the full matrix otherwise)
  static const bool no_sym = [] { const char *e2 = std::getenv("DPGO_DENSE_FULL"); return e2 && e2[0] == '1'; }();
  static const double chunk_cost = [] { const char *e4 = std::getenv("DPGO_SYM_CHUNK_COST"); return e4 ? std::atof(e4) : SYM_CHUNK_COST; }();
  p->sym_ok = 0;
  SymPlan plan;
  if (!no_sym && make_sym_plan((int64_t)N, p->grid, chunk_cost, plan)) {
    const int nseg = plan.nseg, nchunks = plan.nchunks;
    DPGO_CUDA(cudaMalloc(&p->d_sym_cut, sizeof(int) * plan.cut.size()));
    DPGO_CUDA(cudaMalloc(&p->d_sym_segptr, sizeof(int) * plan.segptr.size()));
    DPGO_CUDA(cudaMalloc(&p->d_sym_cfirst, sizeof(int) * plan.cfirst.size()));
    DPGO_CUDA(cudaMalloc(&p->d_sym_ccount, sizeof(int) * plan.ccount.size()));
    DPGO_CUDA(cudaMalloc(&p->d_dense_t2, sizeof(double) * (size_t)nseg * p->r * N));
    DPGO_CUDA(cudaMemsetAsync(p->d_dense_t2, 0, sizeof(double) * (size_t)nseg * p->r * N, p->stream));
    DPGO_CUDA(cudaMalloc(&p->d_sym_off, sizeof(long long) * plan.off.size()));
    DPGO_CUDA(cudaMalloc(&p->d_ppack, sizeof(double) * (size_t)plan.off[(size_t)nchunks]));
    DPGO_CUDA(cudaMemcpy(p->d_sym_off, plan.off.data(), sizeof(long long) * plan.off.size(), cudaMemcpyHostToDevice));
    DPGO_CUDA(cudaMemcpy(p->d_sym_cut, plan.cut.data(), sizeof(int) * plan.cut.size(), cudaMemcpyHostToDevice));
    DPGO_CUDA(cudaMemcpy(p->d_sym_segptr, plan.segptr.data(), sizeof(int) * plan.segptr.size(), cudaMemcpyHostToDevice));
    DPGO_CUDA(cudaMemcpy(p->d_sym_cfirst, plan.cfirst.data(), sizeof(int) * plan.cfirst.size(), cudaMemcpyHostToDevice));
    DPGO_CUDA(cudaMemcpy(p->d_sym_ccount, plan.ccount.data(), sizeof(int) * plan.ccount.size(), cudaMemcpyHostToDevice));
    DPGO_CUDA(dpgo::launch_pack_sym(p->d_pinv, (int)N, nchunks, p->d_sym_segptr, nseg, p->d_sym_off, p->d_ppack, p->stream));
    DPGO_CUDA(cudaStreamSynchronize(p->stream));
    p->sym_ok = 1;
  }
  return DPGO_OK;
}

void nd_fill_info(const dpgo::nd::Hierarchy &H, const dpgo::nd::Plan &P, int64_t *info) {
  for (int i = 0; i < 16; ++i) info[i] = 0;
  int smax = 0, bmax = 0;
  for (const auto &m : H.nodes) { smax = std::max(smax, (int)m.own.size() * H.dh); bmax = std::max(bmax, (int)m.bnd.size() * H.dh); }
  info[0] = H.nstages; info[1] = (int64_t)H.nodes.size(); info[2] = (int64_t)P.phases.size(); info[3] = H.blob_doubles * 8;
  info[4] = P.bytes_per_apply; info[5] = smax; info[6] = bmax; info[7] = H.nd_depth; info[8] = (int64_t)P.steps.size();
  info[9] = (int64_t)P.jobs.size(); info[10] = (int64_t)P.epis.size(); info[11] = P.max_ytiles; info[12] = P.max_slots;
}

dpgo::nd::Options nd_options(int grid, int r, bool cluster = false) {
  dpgo::nd::Options opt;
  opt.grid = grid;
  opt.r = r;
  // a phase end is a hardware cluster barrier in cluster mode: deeper dissections pay off earlier (measured on the
  // 16-agent torus3D workload, 312 poses per agent: 6585 -> 8250 rounds/s; sphere2500's 156-pose agents keep their plan)
  if (cluster) opt.t_phase_us = 2.0;
  opt.warps = dpgo::OPT_THREADS / 32;
  opt.ycap_tiles = dpgo::ND_YCAP_TILES;
  opt.slot_cap = dpgo::ND_SLOT_CAP;
  if (const char *e = std::getenv("DPGO_ND_CUTS")) opt.force_ncuts = std::atoi(e);
  if (const char *e = std::getenv("DPGO_ND_LEAF")) opt.leaf_size = std::max(1, std::atoi(e));
  if (const char *e = std::getenv("DPGO_ND_TPHASE_US")) opt.t_phase_us = std::atof(e);
  if (const char *e = std::getenv("DPGO_ND_BW_GBS")) opt.bw_gbs = std::atof(e);
  if (const char *e = std::getenv("DPGO_ND_TTILE_US")) opt.t_tile_us = std::atof(e);
  return opt;
}

template <class T> int upload_array(const std::vector<T> &h, const T *&d, cudaStream_t stream) {
  T *ptr = nullptr;
  const size_t bytes = sizeof(T) * std::max<size_t>(h.size(), 1);
  DPGO_CUDA(cudaMalloc(&ptr, bytes));
  if (!h.empty()) DPGO_CUDA(cudaMemcpyAsync(ptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, stream));
  d = ptr;
  return DPGO_OK;
}

void free_nd(dpgo_problem *p) {
  ++p->generation;
  auto fr = [](const void *q) { if (q) cudaFree(const_cast<void *>(q)); };
  fr(p->nd.cta_phase); fr(p->nd.steps); fr(p->nd.gathers); fr(p->nd.jobs); fr(p->nd.epis); fr(p->nd.csrc);
  fr(p->nd.blob); fr(p->nd.TX); fr(p->nd.C);
  p->nd = dpgo::KNd();
  delete p->nd_H;
  p->nd_H = nullptr;
  p->nd_ready = false;
}

// The nested-dissection block factorisation of Q + 0.1 I is built on first use (host: ordering, symbolic, plan; the
// dense algebra of large blocks on the device), like the dense inverse.
int ensure_nd(dpgo_problem *p) {
  if (p->nd_ready) return DPGO_OK;
  if (!(p->precond_mask & (1u << DPGO_PRECOND_SPARSE_EXACT)))
    return fail(DPGO_ERR_STATE, "sparse exact preconditioner was not requested in set_Q (precond_mask)");
  namespace nd = dpgo::nd;
  free_nd(p);
  nd::Plan plan;
  std::vector<double> blob;
  nd::Hierarchy *H = new nd::Hierarchy();
  try {
    const nd::Options opt = nd_options(p->grid, p->r, p->cluster);
    nd::BsrView Q{p->n, p->dh, p->h_rowptr.data(), p->h_bcol.data(), p->h_bval.data()};
    nd::build_hierarchy(Q, opt, *H);
    nd::build_numeric(Q, opt, *H, blob);
    nd::build_plan(*H, opt, plan);
  } catch (const std::exception &e) {
    delete H;
    return fail(DPGO_ERR_UNSUPPORTED, std::string("sparse exact preconditioner setup: ") + e.what());
  }
  if (plan.max_ytiles > dpgo::ND_YCAP_TILES || plan.max_slots > dpgo::ND_SLOT_CAP) {
    delete H;
    return fail(DPGO_ERR_UNSUPPORTED, "sparse exact preconditioner: plan exceeds the shared-memory capacities");
  }
  p->nd_H = H;
  nd_fill_info(*H, plan, p->nd_info);
  if ((int)plan.phases.size() > dpgo::nd::MAX_PHASES) {
    delete H;
    p->nd_H = nullptr;
    return fail(DPGO_ERR_UNSUPPORTED, "sparse exact preconditioner: too many phases");
  }
  for (size_t k = 0; k < plan.phases.size(); ++k) { p->nd.dir[k] = plan.phases[k].dir; p->nd.cta0[k] = plan.phases[k].cta0; }
  p->nd.max_ytiles = std::max(plan.max_ytiles, 1);
  p->nd.max_slots = std::max(plan.max_slots, 1);
  p->nd.max_gathers = p->nd.max_ytiles;          // a step gathers at most what its shared-memory tiles hold
  DPGO_TRY(upload_array(plan.cta_phase, p->nd.cta_phase, p->stream));
  DPGO_TRY(upload_array(plan.steps, p->nd.steps, p->stream));
  DPGO_TRY(upload_array(plan.gathers, p->nd.gathers, p->stream));
  DPGO_TRY(upload_array(plan.jobs, p->nd.jobs, p->stream));
  DPGO_TRY(upload_array(plan.epis, p->nd.epis, p->stream));
  DPGO_TRY(upload_array(plan.csrc, p->nd.csrc, p->stream));
  DPGO_TRY(upload_array(blob, p->nd.blob, p->stream));
  const size_t tile = sizeof(double) * (size_t)p->ts;
  DPGO_CUDA(cudaMalloc(&p->nd.TX, tile * (size_t)p->n));
  DPGO_CUDA(cudaMalloc(&p->nd.C, tile * (size_t)H->cbuf_tiles));
  DPGO_CUDA(cudaMemsetAsync(p->nd.TX, 0, tile * (size_t)p->n, p->stream));
  DPGO_CUDA(cudaMemsetAsync(p->nd.C, 0, tile * (size_t)H->cbuf_tiles, p->stream));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  p->nd.nphases = (int)plan.phases.size();
  p->nd_ready = true;
  ++p->generation;
  return DPGO_OK;
}

int check_precond(dpgo_problem *p, int precond) {
  if (precond < 0 || precond > 3) return fail(DPGO_ERR_INVALID_ARG, "unknown preconditioner id");
  if (precond == DPGO_PRECOND_SPARSE_EXACT) return ensure_nd(p);
  if (precond == DPGO_PRECOND_BLOCK_JACOBI && !p->d_dinv)
    return fail(DPGO_ERR_STATE, "block-Jacobi preconditioner was not prepared by set_Q (precond_mask)");
  if (precond == DPGO_PRECOND_DENSE_EXACT) return ensure_dense(p);
  return DPGO_OK;
}

int run_op(dpgo_problem *p, int op, const dpgo_opt_params_t &prm) {
  DPGO_REQUIRE(p->have_Q, DPGO_ERR_STATE, "set_Q has not been called");
  dpgo::KParams kp;
  fill_kparams(p, kp, op, prm);
  DPGO_CUDA(dpgo::launch_optimize(p->r, p->dh, kp, p->stream));
  return DPGO_OK;
}

// ---- host-side block assembly ---------------------------------------------------------------
struct BlockTriplet {
  int brow, bcol;      // Q sub-block at rows dh*brow.., cols dh*bcol..
  double v[16];        // padded 4x4, v[k*4+c] = Q[dh*brow+k, dh*bcol+c]
};

// block triplets -> block-CSR (rows = output tiles, duplicates summed in a fixed order)
void assemble_bsr(int n, const std::vector<BlockTriplet> &trip, std::vector<int> &rowptr, std::vector<int> &bcol,
                  std::vector<double> &bval) {
  // sort by (output tile = bcol, neighbour tile = brow); stable so duplicate summation order is fixed
  std::vector<int64_t> order(trip.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
    if (trip[x].bcol != trip[y].bcol) return trip[x].bcol < trip[y].bcol;
    return trip[x].brow < trip[y].brow;
  });
  rowptr.assign((size_t)n + 1, 0);
  bcol.clear();
  bval.clear();
  bcol.reserve(trip.size());
  bval.reserve(trip.size() * 16);
  int last_j = -1, last_i = -1;
  for (int64_t o : order) {
    const BlockTriplet &t = trip[o];
    if (t.bcol == last_j && t.brow == last_i) {
      double *dst = &bval[bval.size() - 16];
      for (int e = 0; e < 16; ++e) dst[e] += t.v[e];
    } else {
      bcol.push_back(t.brow);
      bval.insert(bval.end(), t.v, t.v + 16);
      rowptr[t.bcol + 1]++;
      last_j = t.bcol;
      last_i = t.brow;
    }
  }
  for (int j = 0; j < n; ++j) rowptr[j + 1] += rowptr[j];
}

// block-Jacobi inverse blocks (Q_jj + 0.1 I)^-1, stored [k][c] padded
void jacobi_blocks(int n, int dh, const std::vector<int> &rowptr, const std::vector<int> &bcol, const std::vector<double> &bval,
                   std::vector<double> &dinv) {
  dinv.assign((size_t)n * 16, 0.0);
  for (int j = 0; j < n; ++j) {
    double A[4][8];
    for (int k = 0; k < 4; ++k)
      for (int c = 0; c < 8; ++c) A[k][c] = (c >= 4 && c - 4 == k) ? 1.0 : 0.0;
    for (int k = 0; k < dh; ++k) A[k][k] = 0.1;
    for (int k = dh; k < 4; ++k) A[k][k] = 1.0;
    for (int b = rowptr[j]; b < rowptr[j + 1]; ++b)
      if (bcol[b] == j)
        for (int k = 0; k < dh; ++k)
          for (int c = 0; c < dh; ++c) A[k][c] += bval[(size_t)b * 16 + k * 4 + c];
    for (int k = 0; k < 4; ++k) {          // Gauss-Jordan, SPD so no pivoting
      const double inv = 1.0 / A[k][k];
      for (int c = 0; c < 8; ++c) A[k][c] *= inv;
      for (int i = 0; i < 4; ++i)
        if (i != k) {
          const double f = A[i][k];
          for (int c = 0; c < 8; ++c) A[i][c] -= f * A[k][c];
        }
    }
    for (int k = 0; k < dh; ++k)
      for (int c = 0; c < dh; ++c) dinv[(size_t)j * 16 + k * 4 + c] = A[k][4 + c];
  }
}

int build_from_triplets(dpgo_problem *p, std::vector<BlockTriplet> &trip, unsigned precond_mask) {
  const int n = p->n, dh = p->dh;
  std::vector<int> rowptr, bcol;
  std::vector<double> bval;
  assemble_bsr(n, trip, rowptr, bcol, bval);
  const int64_t nb = (int64_t)bcol.size();

  std::vector<double> dinv;
  if (precond_mask & (1u << DPGO_PRECOND_BLOCK_JACOBI)) jacobi_blocks(n, dh, rowptr, bcol, bval, dinv);

  // persistent-kernel grid and balanced row partition
  const int sg = (p->r > 4) ? 32 : ((p->r > 2) ? 16 : 8);
  const int rows_per_pass = (dpgo::OPT_THREADS / 32) * (32 / sg);
  int grid = p->max_grid;
  const bool dense = (precond_mask & ((1u << DPGO_PRECOND_DENSE_EXACT) | (1u << DPGO_PRECOND_SPARSE_EXACT))) != 0;
  if (!dense) grid = std::max(1, std::min(grid, (n + rows_per_pass - 1) / rows_per_pass));
  // Launch mode 1 (dpgo_problem_set_launch_mode; or DPGO_CLUSTER_MAX_POSES=<n> for handles left at the default): the step
  // kernel runs as ONE thread-block cluster (<= 16 CTAs) whose phase ends are hardware cluster barriers instead of the
  // atomic-counter grid barrier.  For one agent alone this is SLOWER than the full grid (scripts/phase_times.py --agents 8 /
  // 16, sphere2500: 0.259 vs 0.144 ms per step at 312 poses, 0.176 vs 0.138 ms at 156 poses: 10-16 SMs stream the
  // preconditioner blocks 2.5x slower than 148 SMs); its point is that a cluster launch is not cooperative, so the agents of
  // a colour class run side by side on one GPU and the round captures into a CUDA graph (dpgo_agents_round_async).
  static const int cluster_max_poses = [] { const char *e = std::getenv("DPGO_CLUSTER_MAX_POSES"); return e ? std::atoi(e) : 0; }();
  p->cluster = false;
  const bool want_cluster = p->launch_mode == 1 || (p->launch_mode < 0 && n <= cluster_max_poses);
  if (p->max_cluster >= 8 && want_cluster && !(precond_mask & (1u << DPGO_PRECOND_DENSE_EXACT))) {
    // (always taking 16 CTAs was measured: 156-pose agents 9674 -> 8823 rounds/s; one CTA per 16 poses is enough)
    grid = std::max(1, std::min(p->max_cluster, (n + rows_per_pass - 1) / rows_per_pass));
    p->cluster = true;
  }
  std::vector<int> cta_rows(grid + 1, 0);
  {
    // cost model: blocks + constant epilogue weight per row
    const double wrow = 4.0;
    double total = 0.0;
    for (int j = 0; j < n; ++j) total += (rowptr[j + 1] - rowptr[j]) + wrow;
    double accw = 0.0;
    int ci = 1;
    for (int j = 0; j < n; ++j) {
      accw += (rowptr[j + 1] - rowptr[j]) + wrow;
      while (ci < grid && accw >= total * ci / grid) cta_rows[ci++] = j + 1;
    }
    while (ci <= grid) cta_rows[ci++] = n;
  }

  // upload
  cudaSetDevice(p->device);
  free_dev(p->d_rowptr); free_dev(p->d_bcol); free_dev(p->d_bval); free_dev(p->d_dinv); free_dev(p->d_pinv);
  free_dev(p->d_cta_rows); free_dev(p->d_partials); free_dev(p->d_dense_part); free_dev(p->d_groups);
  free_dev(p->d_dense_t2); free_dev(p->d_ppack); free_dev(p->d_sym_off); free_dev(p->d_sym_cut); free_dev(p->d_sym_segptr); free_dev(p->d_sym_cfirst); free_dev(p->d_sym_ccount);
  p->sym_ok = 0;
  p->have_Q = false;
  p->ngroups = 0;
  free_nd(p);
  p->h_rowptr = rowptr;
  p->h_bcol = bcol;
  p->h_bval = bval;
  // +8 ints of slack: the bulk-TMA windows are rounded out to 16 bytes
  DPGO_CUDA(cudaMalloc(&p->d_rowptr, sizeof(int) * (n + 1 + 8)));
  DPGO_CUDA(cudaMalloc(&p->d_bcol, sizeof(int) * (std::max<int64_t>(nb, 1) + 8)));
  DPGO_CUDA(cudaMemsetAsync(p->d_rowptr, 0, sizeof(int) * (n + 1 + 8), p->stream));
  DPGO_CUDA(cudaMemsetAsync(p->d_bcol, 0, sizeof(int) * (std::max<int64_t>(nb, 1) + 8), p->stream));
  DPGO_CUDA(cudaMalloc(&p->d_bval, sizeof(double) * 16 * std::max<int64_t>(nb, 1)));
  DPGO_CUDA(cudaMalloc(&p->d_cta_rows, sizeof(int) * (grid + 1)));
  DPGO_CUDA(cudaMalloc(&p->d_partials, sizeof(double) * 2 * grid * dpgo::NRED));
  DPGO_CUDA(cudaMemcpyAsync(p->d_rowptr, rowptr.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, p->stream));
  if (nb) {
    DPGO_CUDA(cudaMemcpyAsync(p->d_bcol, bcol.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, p->stream));
    DPGO_CUDA(cudaMemcpyAsync(p->d_bval, bval.data(), sizeof(double) * 16 * nb, cudaMemcpyHostToDevice, p->stream));
  }
  DPGO_CUDA(cudaMemcpyAsync(p->d_cta_rows, cta_rows.data(), sizeof(int) * (grid + 1), cudaMemcpyHostToDevice, p->stream));
  DPGO_CUDA(cudaMemsetAsync(p->d_partials, 0, sizeof(double) * 2 * grid * dpgo::NRED, p->stream));
  if (!dinv.empty()) {
    DPGO_CUDA(cudaMalloc(&p->d_dinv, sizeof(double) * dinv.size()));
    DPGO_CUDA(cudaMemcpyAsync(p->d_dinv, dinv.data(), sizeof(double) * dinv.size(), cudaMemcpyHostToDevice, p->stream));
  }
  // row groups for the TMA-fed SpMV: consecutive rows, <= SPMV_GROUP_BLOCKS blocks and rows each
  {
    const int BT = dpgo::spmv_group_blocks();
    std::vector<int2> groups;
    bool ok = true;
    int rr = 0;
    while (rr < n) {
      const int start = rr;
      int blocks = 0;
      while (rr < n && (rr - start) < BT && blocks + (rowptr[rr + 1] - rowptr[rr]) <= BT) {
        blocks += rowptr[rr + 1] - rowptr[rr];
        ++rr;
      }
      if (rr == start) { ok = false; break; }        // a single row exceeds the stage: fall back to the gather kernel
      groups.push_back(make_int2(start, rowptr[start]));
    }
    if (ok && nb > 0) {
      groups.push_back(make_int2(n, (int)nb));
      DPGO_CUDA(cudaMalloc(&p->d_groups, sizeof(int2) * groups.size()));
      DPGO_CUDA(cudaMemcpyAsync(p->d_groups, groups.data(), sizeof(int2) * groups.size(), cudaMemcpyHostToDevice, p->stream));
      p->ngroups = (int)groups.size() - 1;
    }
  }
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  p->nb = nb;
  p->grid = grid;
  p->precond_mask = precond_mask;

  p->have_Q = true;
  return DPGO_OK;
}

int upload_vec(dpgo_problem *p, int id, const double *host) {
  DPGO_CUDA(cudaMemcpyAsync(p->d_vec[id], host, p->vec_bytes(), cudaMemcpyHostToDevice, p->stream));
  return DPGO_OK;
}
int download_vec(dpgo_problem *p, int id, double *host) {
  DPGO_CUDA(cudaMemcpyAsync(host, p->d_vec[id], p->vec_bytes(), cudaMemcpyDeviceToHost, p->stream));
  return DPGO_OK;
}
int fetch_result(dpgo_problem *p) {
  DPGO_CUDA(cudaMemcpyAsync(p->h_result, p->d_result, sizeof(dpgo_opt_result_t), cudaMemcpyDeviceToHost, p->stream));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

#define DPGO_CHECK_HANDLE(p)                                                         \
  do {                                                                               \
    if (!(p)) return fail(DPGO_ERR_INVALID_ARG, "null problem handle");              \
    cudaError_t _e = cudaSetDevice((p)->device);                                     \
    if (_e != cudaSuccess) return fail(DPGO_ERR_CUDA, cudaGetErrorString(_e));       \
  } while (0)

}  // namespace

extern "C" {

int dpgo_abi_version(void) { return DPGO_B200_ABI_VERSION; }
const char *dpgo_last_error(void) { return g_last_error.c_str(); }

int dpgo_device_count(int *count) {
  DPGO_REQUIRE(count, DPGO_ERR_INVALID_ARG, "null count");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(DPGO_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *count = c;
  return DPGO_OK;
}

void dpgo_opt_params_default(dpgo_opt_params_t *p) {
  if (!p) return;
  p->algorithm = DPGO_ALG_RTR;        // ref: src/QuadraticOptimizer.cpp:22-28
  p->tr_iterations = 1;
  p->tr_max_inner = 50;
  p->precond = DPGO_PRECOND_SPARSE_EXACT;
  p->rgd_stepsize = 1e-3;
  p->tr_tolerance = 1e-2;
  p->tr_initial_radius = 1e1;
}

int dpgo_problem_create(int n, int d, int r, int device, dpgo_problem_t **out) {
  DPGO_REQUIRE(out, DPGO_ERR_INVALID_ARG, "null output handle");
  *out = nullptr;
  DPGO_REQUIRE(n >= 1, DPGO_ERR_INVALID_ARG, "n must be >= 1");
  DPGO_REQUIRE(n <= 100000000, DPGO_ERR_UNSUPPORTED, "n above 1e8 poses: element offsets are 32-bit in the kernels");
  DPGO_REQUIRE(d == 2 || d == 3, DPGO_ERR_UNSUPPORTED, "d must be 2 or 3");
  DPGO_REQUIRE(r >= d, DPGO_ERR_INVALID_ARG, "r must be >= d (ref: assert(r >= d), src/QuadraticProblem.cpp:19)");
  DPGO_REQUIRE((d == 3 && r <= 5) || (d == 2 && (r <= 3 || r == 5)), DPGO_ERR_UNSUPPORTED,
               "unsupported rank (compiled instantiations: d=3: r in 3..5; d=2: r in {2,3,5})");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(DPGO_ERR_NO_DEVICE, "no CUDA device available: the B200 path has no CPU fallback");
  DPGO_REQUIRE(device >= 0 && device < count, DPGO_ERR_NO_DEVICE, "device index out of range");
  DPGO_CUDA(cudaSetDevice(device));
  dpgo_problem *p = new (std::nothrow) dpgo_problem();
  if (!p) return fail(DPGO_ERR_ALLOC, "host allocation failed");
  p->n = n; p->d = d; p->r = r; p->dh = d + 1; p->N = (d + 1) * n; p->ts = r * (d + 1);
  {
    static uint64_t serial = 0;             // handles are created from one thread at a time (as the rest of this API)
    p->generation = (++serial) << 24;       // a recycled address never matches the key of a captured round
  }
  p->device = device;
  auto bail = [&](int code, const std::string &m) { dpgo_problem_destroy(p); return fail(code, m); };
  if (cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
    return bail(DPGO_ERR_CUDA, "cudaDeviceGetAttribute failed");
  if (cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(DPGO_ERR_CUDA, "cudaStreamCreate failed");
  p->stream = p->own_stream;
  p->max_grid = dpgo::optimize_max_grid(r, d + 1, device);
  p->max_cluster = dpgo::optimize_max_cluster(r, d + 1, device);
  if (p->max_grid <= 0) return bail(DPGO_ERR_CUDA, "persistent kernel cannot be made resident on this device");
  const size_t vb = p->vec_bytes();
  for (int i = 0; i < dpgo::V_COUNT; ++i) {
    if (cudaMalloc(&p->d_vec[i], vb) != cudaSuccess) return bail(DPGO_ERR_ALLOC, "device allocation failed (vectors)");
    cudaMemsetAsync(p->d_vec[i], 0, vb, p->stream);
  }
  if (cudaMalloc(&p->d_G, vb) != cudaSuccess) return bail(DPGO_ERR_ALLOC, "device allocation failed (G)");
  cudaMemsetAsync(p->d_G, 0, vb, p->stream);
  for (int i = 0; i < 2; ++i) {
    if (cudaMalloc(&p->d_S[i], sizeof(double) * 9 * (size_t)n) != cudaSuccess) return bail(DPGO_ERR_ALLOC, "device allocation failed (S)");
    cudaMemsetAsync(p->d_S[i], 0, sizeof(double) * 9 * (size_t)n, p->stream);
  }
  if (cudaMalloc(&p->d_bar, 2 * sizeof(unsigned)) != cudaSuccess) return bail(DPGO_ERR_ALLOC, "device allocation failed");
  cudaMemsetAsync(p->d_bar, 0, 2 * sizeof(unsigned), p->stream);
  if (cudaMalloc(&p->d_result, sizeof(dpgo_opt_result_t)) != cudaSuccess) return bail(DPGO_ERR_ALLOC, "device allocation failed");
  if (cudaMallocHost(&p->h_result, sizeof(dpgo_opt_result_t)) != cudaSuccess) return bail(DPGO_ERR_ALLOC, "pinned allocation failed");
  if (cudaStreamSynchronize(p->stream) != cudaSuccess) return bail(DPGO_ERR_CUDA, "device initialisation failed");
  // empty Q (ref: ctor calls setQ(SparseMatrix(N,N)), src/QuadraticProblem.cpp:23)
  std::vector<BlockTriplet> none;
  int s = build_from_triplets(p, none, 1u << DPGO_PRECOND_BLOCK_JACOBI);
  if (s != DPGO_OK) { std::string m = g_last_error; dpgo_problem_destroy(p); return fail(s, m); }
  *out = p;
  return DPGO_OK;
}

int dpgo_problem_destroy(dpgo_problem_t *p) {
  if (!p) return DPGO_OK;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  free_dev(p->d_rowptr); free_dev(p->d_bcol); free_dev(p->d_cta_rows); free_dev(p->d_bval); free_dev(p->d_groups);
  free_dev(p->d_dinv); free_dev(p->d_pinv); free_dev(p->d_dense_part); free_dev(p->d_G);
  free_dev(p->d_dense_t2); free_dev(p->d_ppack); free_dev(p->d_sym_off); free_dev(p->d_sym_cut); free_dev(p->d_sym_segptr); free_dev(p->d_sym_cfirst); free_dev(p->d_sym_ccount);
  for (int i = 0; i < dpgo::V_COUNT; ++i) free_dev(p->d_vec[i]);
  free_dev(p->d_S[0]); free_dev(p->d_S[1]); free_dev(p->d_partials); free_dev(p->d_bar); free_dev(p->d_phase_ns); free_dev(p->d_result);
  free_dev(p->d_acc[0]); free_dev(p->d_acc[1]); free_dev(p->d_acc[2]);
  free_dev(p->d_e_p1); free_dev(p->d_e_p2); free_dev(p->d_e_fixed); free_dev(p->d_cptr); free_dev(p->d_contrib);
  free_dev(p->d_eT); free_dev(p->d_eom); free_dev(p->d_ew); free_dev(p->d_sblk); free_dev(p->d_eres);
  free_dev(p->d_public); free_dev(p->d_pose_ids); free_dev(p->d_pose_ptr); free_dev(p->d_edge_slot);
  free_dev(p->d_edge_out); free_dev(p->d_edge_T); free_dev(p->d_edge_om);
  free_nd(p);
  if (p->h_result) cudaFreeHost(p->h_result);
  for (auto &g : p->round_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  p->round_graphs.clear();
  if (p->ev_done) cudaEventDestroy(p->ev_done);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->own_stream) cudaStreamDestroy(p->own_stream);
  delete p;
  return DPGO_OK;
}

int dpgo_problem_set_stream(dpgo_problem_t *p, void *cuda_stream) {
  DPGO_CHECK_HANDLE(p);
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  p->stream = cuda_stream ? (cudaStream_t)cuda_stream : p->own_stream;
  return DPGO_OK;
}

int dpgo_problem_sync(dpgo_problem_t *p) {
  DPGO_CHECK_HANDLE(p);
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

int dpgo_problem_set_launch_mode(dpgo_problem_t *p, int mode) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(mode >= -1 && mode <= 1, DPGO_ERR_INVALID_ARG, "launch mode must be -1 (default), 0 (grid) or 1 (cluster)");
  DPGO_REQUIRE(mode != 1 || p->max_cluster >= 8, DPGO_ERR_UNSUPPORTED, "this device cannot co-schedule a cluster of >= 8 CTAs of the step kernel");
  p->launch_mode = mode;
  return DPGO_OK;
}

int dpgo_problem_launch_info(const dpgo_problem_t *p, int *grid, int *cluster) {
  DPGO_REQUIRE(p, DPGO_ERR_INVALID_ARG, "null problem handle");
  if (grid) *grid = p->grid;
  if (cluster) *cluster = p->cluster ? 1 : 0;
  return DPGO_OK;
}

int dpgo_problem_dims(const dpgo_problem_t *p, int *n, int *d, int *r, int64_t *num_blocks) {
  DPGO_REQUIRE(p, DPGO_ERR_INVALID_ARG, "null problem handle");
  if (n) *n = p->n;
  if (d) *d = p->d;
  if (r) *r = p->r;
  if (num_blocks) *num_blocks = p->nb;
  return DPGO_OK;
}

int dpgo_problem_set_Q_csr(dpgo_problem_t *p, int nrows, const int32_t *rowptr, const int32_t *colind,
                           const double *values, unsigned precond_mask) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(nrows == p->N, DPGO_ERR_INVALID_ARG, "Q must be (d+1)n x (d+1)n");
  DPGO_REQUIRE(rowptr, DPGO_ERR_INVALID_ARG, "null CSR row pointer");
  DPGO_REQUIRE(rowptr[0] == 0, DPGO_ERR_INVALID_ARG, "CSR row pointer must start at 0");
  for (int i = 0; i < nrows; ++i)
    if (rowptr[i + 1] < rowptr[i]) return fail(DPGO_ERR_INVALID_ARG, "CSR row pointer is not non-decreasing");
  DPGO_REQUIRE(rowptr[nrows] == 0 || (colind && values), DPGO_ERR_INVALID_ARG, "null CSR arrays");
  const int dh = p->dh;
  std::vector<BlockTriplet> trip;
  trip.reserve((size_t)rowptr[nrows] / (dh * dh) + 16);
  for (int ib = 0; ib < p->n; ++ib) {
    const size_t first = trip.size();
    for (int k = 0; k < dh; ++k) {
      const int rr = ib * dh + k;
      for (int q = rowptr[rr]; q < rowptr[rr + 1]; ++q) {
        const int cc = colind[q];
        if (cc < 0 || cc >= p->N) return fail(DPGO_ERR_INVALID_ARG, "column index out of range");
        const int jb = cc / dh, c = cc - jb * dh;
        size_t t = first;
        for (; t < trip.size(); ++t)
          if (trip[t].bcol == jb) break;
        if (t == trip.size()) {
          BlockTriplet bt;
          bt.brow = ib;
          bt.bcol = jb;
          std::memset(bt.v, 0, sizeof(bt.v));
          trip.push_back(bt);
        }
        trip[t].v[k * 4 + c] += values[q];
      }
    }
  }
  return build_from_triplets(p, trip, precond_mask);
}

int dpgo_problem_set_Q_blocks(dpgo_problem_t *p, int64_t nb, const int32_t *brow, const int32_t *bcol,
                              const double *blocks, unsigned precond_mask) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(nb >= 0 && (nb == 0 || (brow && bcol && blocks)), DPGO_ERR_INVALID_ARG, "null block arrays");
  const int dh = p->dh;
  std::vector<BlockTriplet> trip((size_t)nb);
  for (int64_t q = 0; q < nb; ++q) {
    if (brow[q] < 0 || brow[q] >= p->n || bcol[q] < 0 || bcol[q] >= p->n)
      return fail(DPGO_ERR_INVALID_ARG, "block index out of range");
    trip[q].brow = brow[q];
    trip[q].bcol = bcol[q];
    std::memset(trip[q].v, 0, sizeof(trip[q].v));
    for (int k = 0; k < dh; ++k)
      for (int c = 0; c < dh; ++c) trip[q].v[k * 4 + c] = blocks[(size_t)q * dh * dh + k * dh + c];
  }
  return build_from_triplets(p, trip, precond_mask);
}

int dpgo_problem_set_G_dense(dpgo_problem_t *p, const double *G_host) {
  DPGO_CHECK_HANDLE(p);
  p->G_dirty = true;
  if (!G_host) {
    DPGO_CUDA(cudaMemsetAsync(p->d_G, 0, p->vec_bytes(), p->stream));
  } else {
    DPGO_CUDA(cudaMemcpyAsync(p->d_G, G_host, p->vec_bytes(), cudaMemcpyHostToDevice, p->stream));
  }
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

int dpgo_problem_set_G_csr(dpgo_problem_t *p, const int32_t *rowptr, const int32_t *colind, const double *values) {
  DPGO_CHECK_HANDLE(p);
  if (!rowptr) return dpgo_problem_set_G_dense(p, nullptr);
  std::vector<double> G((size_t)p->r * p->N, 0.0);
  for (int a = 0; a < p->r; ++a)
    for (int q = rowptr[a]; q < rowptr[a + 1]; ++q) {
      if (colind[q] < 0 || colind[q] >= p->N) return fail(DPGO_ERR_INVALID_ARG, "G column index out of range");
      G[(size_t)colind[q] * p->r + a] += values[q];
    }
  return dpgo_problem_set_G_dense(p, G.data());
}

// ---- evaluation ---------------------------------------------------------------------------------
static int eval_at(dpgo_problem_t *p, const double *X_host) {
  DPGO_REQUIRE(X_host, DPGO_ERR_INVALID_ARG, "null X");
  DPGO_TRY(upload_vec(p, dpgo::V_X0, X_host));
  dpgo_opt_params_t prm;
  dpgo_opt_params_default(&prm);
  prm.precond = DPGO_PRECOND_NONE;
  DPGO_TRY(run_op(p, dpgo::OP_EVAL, prm));
  return DPGO_OK;
}

int dpgo_problem_f(dpgo_problem_t *p, const double *X_host, double *f_out) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(f_out, DPGO_ERR_INVALID_ARG, "null output");
  DPGO_TRY(eval_at(p, X_host));
  DPGO_TRY(fetch_result(p));
  *f_out = p->h_result->f_init;
  return DPGO_OK;
}

int dpgo_problem_egrad(dpgo_problem_t *p, const double *X_host, double *out_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(out_host, DPGO_ERR_INVALID_ARG, "null output");
  DPGO_TRY(eval_at(p, X_host));
  DPGO_TRY(download_vec(p, dpgo::V_EG0, out_host));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

int dpgo_problem_ehess(dpgo_problem_t *p, const double *V_host, double *out_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(V_host && out_host, DPGO_ERR_INVALID_ARG, "null argument");
  DPGO_REQUIRE(p->have_Q, DPGO_ERR_STATE, "set_Q has not been called");
  DPGO_TRY(upload_vec(p, dpgo::V_AUX, V_host));
  DPGO_CUDA(run_spmv(p, p->d_vec[dpgo::V_AUX], nullptr, p->d_vec[dpgo::V_HD]));
  DPGO_TRY(download_vec(p, dpgo::V_HD, out_host));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

int dpgo_problem_rgrad(dpgo_problem_t *p, const double *X_host, double *out_host, double *norm_out) {
  DPGO_CHECK_HANDLE(p);
  DPGO_TRY(eval_at(p, X_host));
  if (out_host) DPGO_TRY(download_vec(p, dpgo::V_RG0, out_host));
  DPGO_TRY(fetch_result(p));
  if (norm_out) *norm_out = p->h_result->gradnorm_init;
  return DPGO_OK;
}

int dpgo_problem_f_rgradnorm(dpgo_problem_t *p, const double *X_host, double *f_out, double *norm_out) {
  DPGO_CHECK_HANDLE(p);
  DPGO_TRY(eval_at(p, X_host));
  DPGO_TRY(fetch_result(p));
  if (f_out) *f_out = p->h_result->f_init;
  if (norm_out) *norm_out = p->h_result->gradnorm_init;
  return DPGO_OK;
}

int dpgo_problem_rhess(dpgo_problem_t *p, const double *X_host, const double *V_host, double *out_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(X_host && V_host && out_host, DPGO_ERR_INVALID_ARG, "null argument");
  DPGO_TRY(upload_vec(p, dpgo::V_X0, X_host));
  DPGO_TRY(upload_vec(p, dpgo::V_AUX, V_host));
  dpgo_opt_params_t prm;
  dpgo_opt_params_default(&prm);
  prm.precond = DPGO_PRECOND_NONE;
  DPGO_TRY(run_op(p, dpgo::OP_RHESS, prm));
  DPGO_TRY(download_vec(p, dpgo::V_HD, out_host));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

int dpgo_problem_precon(dpgo_problem_t *p, int precond, const double *X_host, const double *V_host, double *out_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(X_host && V_host && out_host, DPGO_ERR_INVALID_ARG, "null argument");
  DPGO_TRY(check_precond(p, precond));
  DPGO_TRY(upload_vec(p, dpgo::V_X0, X_host));
  DPGO_TRY(upload_vec(p, dpgo::V_AUX, V_host));
  dpgo_opt_params_t prm;
  dpgo_opt_params_default(&prm);
  prm.precond = precond;
  DPGO_TRY(run_op(p, dpgo::OP_PRECON, prm));
  DPGO_TRY(download_vec(p, dpgo::V_Z, out_host));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

int dpgo_manifold_tangent_project(dpgo_problem_t *p, const double *X_host, const double *Z_host, double *out_host) {
  return dpgo_problem_precon(p, DPGO_PRECOND_NONE, X_host, Z_host, out_host);
}

int dpgo_manifold_retract(dpgo_problem_t *p, const double *X_host, const double *eta_host, double *out_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(X_host && eta_host && out_host, DPGO_ERR_INVALID_ARG, "null argument");
  DPGO_TRY(upload_vec(p, dpgo::V_X0, X_host));
  DPGO_TRY(upload_vec(p, dpgo::V_AUX, eta_host));
  dpgo_opt_params_t prm;
  dpgo_opt_params_default(&prm);
  prm.precond = DPGO_PRECOND_NONE;
  DPGO_TRY(run_op(p, dpgo::OP_RETRACT, prm));
  DPGO_TRY(download_vec(p, dpgo::V_X1, out_host));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

int dpgo_manifold_project(dpgo_problem_t *p, const double *M_host, double *out_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(M_host && out_host, DPGO_ERR_INVALID_ARG, "null argument");
  DPGO_TRY(upload_vec(p, dpgo::V_AUX, M_host));
  DPGO_CUDA(dpgo::launch_stiefel_project(p->r, p->dh, p->n, p->d_vec[dpgo::V_AUX], p->d_vec[dpgo::V_T], p->stream));
  DPGO_TRY(download_vec(p, dpgo::V_T, out_host));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

// ---- optimiser -----------------------------------------------------------------------------------
static int check_params(dpgo_problem_t *p, const dpgo_opt_params_t *prm) {
  DPGO_REQUIRE(prm, DPGO_ERR_INVALID_ARG, "null params");
  DPGO_REQUIRE(prm->algorithm == DPGO_ALG_RTR || prm->algorithm == DPGO_ALG_RGD, DPGO_ERR_INVALID_ARG, "unknown algorithm");
  DPGO_REQUIRE(prm->tr_iterations >= 1 && prm->tr_max_inner >= 1, DPGO_ERR_INVALID_ARG, "iteration counts must be >= 1");
  DPGO_REQUIRE(prm->tr_initial_radius > 0 && prm->tr_tolerance >= 0, DPGO_ERR_INVALID_ARG, "bad radius / tolerance");
  if (prm->algorithm == DPGO_ALG_RTR) DPGO_TRY(check_precond(p, prm->precond));
  return DPGO_OK;
}

int dpgo_optimize(dpgo_problem_t *p, const dpgo_opt_params_t *params, const double *X_in_host, double *X_out_host,
                  dpgo_opt_result_t *result) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(X_in_host && X_out_host, DPGO_ERR_INVALID_ARG, "null X");
  DPGO_TRY(check_params(p, params));
  const auto t0 = std::chrono::high_resolution_clock::now();
  DPGO_TRY(upload_vec(p, dpgo::V_X0, X_in_host));
  DPGO_TRY(run_op(p, dpgo::OP_OPTIMIZE, *params));
  DPGO_TRY(download_vec(p, dpgo::V_X0, X_out_host));
  DPGO_TRY(fetch_result(p));
  const auto t1 = std::chrono::high_resolution_clock::now();
  p->h_result->elapsed_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  if (result) *result = *p->h_result;
  return DPGO_OK;
}

int dpgo_problem_upload_X(dpgo_problem_t *p, const double *X_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(X_host, DPGO_ERR_INVALID_ARG, "null X");
  DPGO_TRY(upload_vec(p, dpgo::V_X0, X_host));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

int dpgo_problem_download_X(dpgo_problem_t *p, double *X_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(X_host, DPGO_ERR_INVALID_ARG, "null X");
  DPGO_TRY(download_vec(p, dpgo::V_X0, X_host));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  return DPGO_OK;
}

int dpgo_problem_upload_X_async(dpgo_problem_t *p, const double *X_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(X_host, DPGO_ERR_INVALID_ARG, "null X");
  return upload_vec(p, dpgo::V_X0, X_host);
}

int dpgo_problem_download_X_async(dpgo_problem_t *p, double *X_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(X_host, DPGO_ERR_INVALID_ARG, "null X");
  return download_vec(p, dpgo::V_X0, X_host);
}

int dpgo_problem_copy_X_from_device(dpgo_problem_t *p, const double *X_dev) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(X_dev, DPGO_ERR_INVALID_ARG, "null X");
  DPGO_CUDA(cudaMemcpyAsync(p->d_vec[dpgo::V_X0], X_dev, p->vec_bytes(), cudaMemcpyDeviceToDevice, p->stream));
  return DPGO_OK;
}

int dpgo_problem_device_X(dpgo_problem_t *p, double **X_dev) {
  DPGO_REQUIRE(p && X_dev, DPGO_ERR_INVALID_ARG, "null argument");
  *X_dev = p->d_vec[dpgo::V_X0];
  return DPGO_OK;
}

int dpgo_problem_device_G(dpgo_problem_t *p, double **G_dev) {
  DPGO_REQUIRE(p && G_dev, DPGO_ERR_INVALID_ARG, "null argument");
  *G_dev = p->d_G;
  p->G_dirty = true;             // the caller may write G directly
  return DPGO_OK;
}

int dpgo_optimize_resident_async(dpgo_problem_t *p, const dpgo_opt_params_t *params) {
  DPGO_CHECK_HANDLE(p);
  DPGO_TRY(check_params(p, params));
  p->async_t0 = std::chrono::high_resolution_clock::now();
  DPGO_TRY(run_op(p, dpgo::OP_OPTIMIZE, *params));
  p->async_pending = true;
  return DPGO_OK;
}

int dpgo_optimize_result(dpgo_problem_t *p, dpgo_opt_result_t *result) {
  DPGO_CHECK_HANDLE(p);
  DPGO_TRY(fetch_result(p));
  if (p->async_pending) {
    const auto t1 = std::chrono::high_resolution_clock::now();
    p->h_result->elapsed_ms = std::chrono::duration<double, std::milli>(t1 - p->async_t0).count();
    p->async_pending = false;
  }
  if (result) *result = *p->h_result;
  return DPGO_OK;
}

int dpgo_debug_phase_latency(dpgo_problem_t *p, int phases, double *us_per_phase, double *us_launch) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(phases >= 1 && us_per_phase, DPGO_ERR_INVALID_ARG, "bad arguments");
  dpgo_opt_params_t prm;
  dpgo_opt_params_default(&prm);
  prm.precond = DPGO_PRECOND_NONE;
  cudaEvent_t e0, e1;
  DPGO_CUDA(cudaEventCreate(&e0));
  DPGO_CUDA(cudaEventCreate(&e1));
  float ms[2] = {0, 0};
  const int counts[2] = {1, phases + 1};
  for (int rep = 0; rep < 2; ++rep) {
    prm.tr_max_inner = counts[rep];
    for (int w = 0; w < 3; ++w) DPGO_TRY(run_op(p, dpgo::OP_PHASE_BENCH, prm));
    DPGO_CUDA(cudaEventRecord(e0, p->stream));
    for (int w = 0; w < 10; ++w) DPGO_TRY(run_op(p, dpgo::OP_PHASE_BENCH, prm));
    DPGO_CUDA(cudaEventRecord(e1, p->stream));
    DPGO_CUDA(cudaEventSynchronize(e1));
    DPGO_CUDA(cudaEventElapsedTime(&ms[rep], e0, e1));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *us_per_phase = 1e3 * (ms[1] - ms[0]) / 10.0 / phases;
  if (us_launch) *us_launch = 1e3 * ms[0] / 10.0 - *us_per_phase;
  return DPGO_OK;
}

static int phase_times_impl(dpgo_problem_t *p, int enable, double *ms_by_kind, int nout) {
  DPGO_CHECK_HANDLE(p);
  if (enable && !p->d_phase_ns) {
    DPGO_CUDA(cudaMalloc(&p->d_phase_ns, 64 * sizeof(unsigned long long)));
    DPGO_CUDA(cudaMemsetAsync(p->d_phase_ns, 0, 64 * sizeof(unsigned long long), p->stream));
  }
  if (p->d_phase_ns) {
    unsigned long long ns[64];
    DPGO_CUDA(cudaStreamSynchronize(p->stream));
    DPGO_CUDA(cudaMemcpy(ns, p->d_phase_ns, sizeof(ns), cudaMemcpyDeviceToHost));
    if (ms_by_kind)
      for (int i = 0; i < nout; ++i) ms_by_kind[i] = 1e-6 * (double)ns[i];
    DPGO_CUDA(cudaMemset(p->d_phase_ns, 0, sizeof(ns)));
    if (!enable) { cudaFree(p->d_phase_ns); p->d_phase_ns = nullptr; }
  } else if (ms_by_kind) {
    for (int i = 0; i < nout; ++i) ms_by_kind[i] = 0.0;
  }
  return DPGO_OK;
}

int dpgo_debug_phase_times(dpgo_problem_t *p, int enable, double *ms_by_kind) { return phase_times_impl(p, enable, ms_by_kind, 8); }
int dpgo_debug_phase_times32(dpgo_problem_t *p, int enable, double *ms_by_kind) { return phase_times_impl(p, enable, ms_by_kind, 32); }
int dpgo_debug_phase_times64(dpgo_problem_t *p, int enable, double *ms_by_kind) { return phase_times_impl(p, enable, ms_by_kind, 64); }

int dpgo_spmv_device(dpgo_problem_t *p, const double *X_dev, double *out_dev, int add_G) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(X_dev && out_dev, DPGO_ERR_INVALID_ARG, "null argument");
  DPGO_REQUIRE(p->have_Q, DPGO_ERR_STATE, "set_Q has not been called");
  DPGO_CUDA(run_spmv(p, X_dev, add_G ? p->d_G : nullptr, out_dev));
  return DPGO_OK;
}

int64_t dpgo_spmv_algorithmic_bytes(const dpgo_problem_t *p, int add_G) {
  if (!p) return 0;
  // nb*((d+1)^2*8 + 4) + (n+1)*4 + 2*r*(d+1)*n*8 (+ r*(d+1)*n*8 if G is read)
  const int64_t dh = p->dh;
  int64_t b = p->nb * (dh * dh * 8 + 4) + ((int64_t)p->n + 1) * 4 + 2 * (int64_t)p->r * dh * p->n * 8;
  if (add_G) b += (int64_t)p->r * dh * p->n * 8;
  return b;
}

int dpgo_sym_plan_sizes(int N, int *num_segments, int *num_chunks) {
  DPGO_REQUIRE(num_segments && num_chunks, DPGO_ERR_INVALID_ARG, "null argument");
  DPGO_REQUIRE(N >= 1, DPGO_ERR_INVALID_ARG, "bad dimension");
  const int nseg = (N + SYM_PLAN_SEG - 1) / SYM_PLAN_SEG;
  int nch = 0;
  for (int J = 0; J < nseg; ++J) nch += (std::min(N, (J + 1) * SYM_PLAN_SEG) + 7) / 8;
  *num_segments = nseg;
  *num_chunks = nch;
  return DPGO_OK;
}

int dpgo_sym_plan(int N, int grid, double chunk_cost, int32_t *segptr, int32_t *cut, int32_t *cfirst, int32_t *ccount,
                  int64_t *chunk_offset) {
  DPGO_REQUIRE(segptr && cut && cfirst && ccount && chunk_offset, DPGO_ERR_INVALID_ARG, "null argument");
  SymPlan pl;
  if (!make_sym_plan((int64_t)N, grid, chunk_cost > 0 ? chunk_cost : SYM_CHUNK_COST, pl))
    return fail(DPGO_ERR_UNSUPPORTED, "no symmetric plan for this size (odd N, N < 2048, or more CTAs than chunks)");
  std::copy(pl.segptr.begin(), pl.segptr.end(), segptr);
  std::copy(pl.cut.begin(), pl.cut.end(), cut);
  std::copy(pl.cfirst.begin(), pl.cfirst.end(), cfirst);
  std::copy(pl.ccount.begin(), pl.ccount.end(), ccount);
  std::copy(pl.off.begin(), pl.off.end(), chunk_offset);
  return DPGO_OK;
}

int64_t dpgo_precond_algorithmic_bytes(const dpgo_problem_t *p, int preconditioner) {
  if (!p) return 0;
  const int64_t N = (int64_t)p->dh * p->n, vec = (int64_t)p->r * N * 8;
  if (preconditioner == DPGO_PRECOND_BLOCK_JACOBI) return (int64_t)p->n * 16 * 8 + 2 * vec;
  if (preconditioner == DPGO_PRECOND_SPARSE_EXACT) return p->nd_ready ? p->nd_info[4] + 2 * vec : 0;
  if (preconditioner != DPGO_PRECOND_DENSE_EXACT) return 0;
  // the inverse is symmetric: the unique data is its upper triangle in 8-row groups (what phase_dense_sym reads);
  // the full-matrix variants read all of it
  const int64_t mat = p->sym_ok ? (N * N + 8 * N) / 2 * 8 : N * N * 8;
  return mat + 2 * vec;
}

int dpgo_nd_info(dpgo_problem_t *p, int64_t *info16) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(info16, DPGO_ERR_INVALID_ARG, "null argument");
  DPGO_REQUIRE(p->have_Q, DPGO_ERR_STATE, "set_Q has not been called");
  DPGO_TRY(ensure_nd(p));
  std::copy(p->nd_info, p->nd_info + 16, info16);
  return DPGO_OK;
}

int dpgo_nd_debug_emulate(int n, int d, int r, int64_t nb, const int32_t *brow, const int32_t *bcol, const double *blocks,
                          double shift, int grid, int force_cuts, int leaf_size, const double *V_host, double *Z_host,
                          int64_t *info16) {
  DPGO_REQUIRE(n >= 1 && (d == 2 || d == 3) && r >= 1 && r <= 8 && grid >= 1, DPGO_ERR_INVALID_ARG, "bad dimensions");
  DPGO_REQUIRE(nb >= 0 && (nb == 0 || (brow && bcol && blocks)) && V_host && Z_host, DPGO_ERR_INVALID_ARG, "null argument");
  const int dh = d + 1;
  std::vector<BlockTriplet> trip((size_t)nb);
  for (int64_t q = 0; q < nb; ++q) {
    if (brow[q] < 0 || brow[q] >= n || bcol[q] < 0 || bcol[q] >= n) return fail(DPGO_ERR_INVALID_ARG, "block index out of range");
    trip[(size_t)q].brow = brow[q];
    trip[(size_t)q].bcol = bcol[q];
    std::memset(trip[(size_t)q].v, 0, sizeof(trip[(size_t)q].v));
    for (int k = 0; k < dh; ++k)
      for (int c = 0; c < dh; ++c) trip[(size_t)q].v[k * 4 + c] = blocks[(size_t)q * dh * dh + k * dh + c];
  }
  std::vector<int> rowptr, bc;
  std::vector<double> bv;
  assemble_bsr(n, trip, rowptr, bc, bv);
  namespace nd = dpgo::nd;
  try {
    nd::Options opt = nd_options(grid, r);
    opt.force_ncuts = force_cuts;
    opt.shift = shift;
    if (leaf_size > 0) opt.leaf_size = leaf_size;
    nd::BsrView Q{n, dh, rowptr.data(), bc.data(), bv.data()};
    nd::Hierarchy H;
    nd::Plan plan;
    std::vector<double> blob;
    nd::build_hierarchy(Q, opt, H);
    nd::build_numeric(Q, opt, H, blob);
    nd::build_plan(H, opt, plan);
    if (plan.max_ytiles > opt.ycap_tiles || plan.max_slots > opt.slot_cap)
      return fail(DPGO_ERR_UNSUPPORTED, "plan exceeds the shared-memory capacities");
    nd::emulate_apply(H, plan, blob, r, V_host, Z_host);
    if (info16) nd_fill_info(H, plan, info16);
    if (const char *dump = std::getenv("DPGO_ND_DUMP_PLAN")) {            // per (phase, CTA) work statistics, CSV
      if (FILE *fp = std::fopen(dump, "w")) {
        std::fprintf(fp, "phase,dir,stage,cta,steps,gather_tiles,jobs,job_cols,max_warp_cols,epis\n");
        for (size_t ph = 0; ph < plan.phases.size(); ++ph)
          for (int c = 0; c < plan.grid; ++c) {
            const nd::CtaPhase &cp = plan.cta_phase[(size_t)plan.phases[ph].cta0 + c];
            long gt = 0, nj = 0, jc = 0, ne = 0, mw = 0;
            for (int si = cp.s0; si < cp.s1; ++si) {
              const nd::Step &st = plan.steps[(size_t)si];
              gt += st.g1 - st.g0; nj += st.j1 - st.j0; ne += st.e1 - st.e0;
              std::vector<long> w((size_t)opt.warps, 0);
              for (int j = st.j0; j < st.j1; ++j) { jc += plan.jobs[(size_t)j].ncols; w[(size_t)((j - st.j0) % opt.warps)] += plan.jobs[(size_t)j].ncols + 40; }
              mw += *std::max_element(w.begin(), w.end());
            }
            std::fprintf(fp, "%zu,%d,%d,%d,%d,%ld,%ld,%ld,%ld,%ld\n", ph, plan.phases[ph].dir, plan.phases[ph].stage, c, cp.s1 - cp.s0, gt, nj, jc, mw, ne);
          }
        std::fclose(fp);
      }
    }
  } catch (const std::exception &e) {
    return fail(DPGO_ERR_UNSUPPORTED, std::string("sparse exact preconditioner: ") + e.what());
  }
  return DPGO_OK;
}

// ---- Q from edge records on the device, robust re-weighting -----------------------------------------------------
static int reassemble_Q(dpgo_problem *p) {
  DPGO_CUDA(dpgo::launch_assemble_Q(p->nb, p->d_cptr, p->d_contrib, p->d_eT, p->d_eom, p->d_ew, p->d_sblk, p->d_bval, p->stream));
  // the host copy feeds the lazily built exact preconditioners; block-Jacobi blocks are refreshed right away
  DPGO_CUDA(cudaMemcpyAsync(p->h_bval.data(), p->d_bval, sizeof(double) * 16 * (size_t)p->nb, cudaMemcpyDeviceToHost, p->stream));
  DPGO_CUDA(cudaStreamSynchronize(p->stream));
  if (p->d_dinv) {
    std::vector<double> dinv;
    jacobi_blocks(p->n, p->dh, p->h_rowptr, p->h_bcol, p->h_bval, dinv);
    DPGO_CUDA(cudaMemcpyAsync(p->d_dinv, dinv.data(), sizeof(double) * dinv.size(), cudaMemcpyHostToDevice, p->stream));
    DPGO_CUDA(cudaStreamSynchronize(p->stream));
  }
  free_nd(p);                                            // (Q + 0.1 I)^-1 changed: rebuilt on next use
  free_dev(p->d_pinv); free_dev(p->d_dense_part); free_dev(p->d_dense_t2); free_dev(p->d_ppack); free_dev(p->d_sym_off);
  free_dev(p->d_sym_cut); free_dev(p->d_sym_segptr); free_dev(p->d_sym_cfirst); free_dev(p->d_sym_ccount);
  p->sym_ok = 0;
  return DPGO_OK;
}

int dpgo_problem_set_edges(dpgo_problem_t *p, int64_t m, const int32_t *p1, const int32_t *p2, const double *R, const double *t,
                           const double *kappa, const double *tau, const double *weight, const int32_t *fixed_weight,
                           int64_t num_static, const int32_t *static_pose, const double *static_blocks, unsigned precond_mask) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(m >= 0 && (m == 0 || (p1 && p2 && R && t && kappa && tau)), DPGO_ERR_INVALID_ARG, "null edge arrays");
  DPGO_REQUIRE(num_static >= 0 && (num_static == 0 || (static_pose && static_blocks)), DPGO_ERR_INVALID_ARG, "null static blocks");
  const int d = p->d, dh = p->dh, n = p->n;
  for (int64_t e = 0; e < m; ++e)
    if (p1[e] < 0 || p1[e] >= n || p2[e] < 0 || p2[e] >= n) return fail(DPGO_ERR_INVALID_ARG, "edge endpoint out of range");
  for (int64_t q = 0; q < num_static; ++q)
    if (static_pose[q] < 0 || static_pose[q] >= n) return fail(DPGO_ERR_INVALID_ARG, "static block pose out of range");
  // pattern: zero-valued triplets give the block-CSR structure (and the usual launch tables) ...
  std::vector<BlockTriplet> trip;
  trip.reserve((size_t)(4 * m + num_static));
  auto add = [&](int bi, int bj) { BlockTriplet bt; bt.brow = bi; bt.bcol = bj; std::memset(bt.v, 0, sizeof(bt.v)); trip.push_back(bt); };
  for (int64_t e = 0; e < m; ++e) { add(p1[e], p1[e]); add(p2[e], p2[e]); add(p1[e], p2[e]); add(p2[e], p1[e]); }
  for (int64_t q = 0; q < num_static; ++q) add(static_pose[q], static_pose[q]);
  DPGO_TRY(build_from_triplets(p, trip, precond_mask));
  // ... and every block's contribution list in input order: block (bi, bj) = entry with bcol == bi in row bj
  const std::vector<int> &rowptr = p->h_rowptr, &bcol = p->h_bcol;
  auto find_block = [&](int bi, int bj) {
    const int *lo = bcol.data() + rowptr[(size_t)bj], *hi = bcol.data() + rowptr[(size_t)bj + 1];
    return (int)(std::lower_bound(lo, hi, bi) - bcol.data());
  };
  const int64_t nb = p->nb;
  std::vector<int> cnt((size_t)nb + 1, 0);
  std::vector<std::pair<int, int2>> items;             // (block, (index, kind))
  items.reserve((size_t)(4 * m + num_static));
  for (int64_t e = 0; e < m; ++e) {
    items.push_back({find_block(p1[e], p1[e]), make_int2((int)e, 0)});
    items.push_back({find_block(p2[e], p2[e]), make_int2((int)e, 1)});
    items.push_back({find_block(p1[e], p2[e]), make_int2((int)e, 2)});
    items.push_back({find_block(p2[e], p1[e]), make_int2((int)e, 3)});
  }
  for (int64_t q = 0; q < num_static; ++q) items.push_back({find_block(static_pose[q], static_pose[q]), make_int2((int)q, 4)});
  for (auto &it : items) cnt[(size_t)it.first + 1]++;
  for (int64_t b = 0; b < nb; ++b) cnt[(size_t)b + 1] += cnt[(size_t)b];
  std::vector<int2> contrib(items.size());
  {
    std::vector<int> fill(cnt.begin(), cnt.end() - 1);
    for (auto &it : items) contrib[(size_t)fill[(size_t)it.first]++] = it.second;     // input order inside a block
  }
  std::vector<double> eT((size_t)m * 16, 0.0), eom((size_t)m * 4, 0.0), ew((size_t)m, 1.0), sb((size_t)num_static * 16, 0.0);
  std::vector<int> fx((size_t)m, 0), q1((size_t)m), q2((size_t)m);
  for (int64_t e = 0; e < m; ++e) {
    double *T = &eT[(size_t)e * 16];
    for (int a = 0; a < d; ++a) {
      for (int b = 0; b < d; ++b) T[a * 4 + b] = R[(size_t)e * d * d + a * d + b];
      T[a * 4 + d] = t[(size_t)e * d + a];
      eom[(size_t)e * 4 + a] = kappa[e];
    }
    T[d * 4 + d] = 1.0;
    eom[(size_t)e * 4 + d] = tau[e];
    if (weight) ew[(size_t)e] = weight[e];
    if (fixed_weight) fx[(size_t)e] = fixed_weight[e] ? 1 : 0;
    q1[(size_t)e] = p1[e];
    q2[(size_t)e] = p2[e];
  }
  for (int64_t q = 0; q < num_static; ++q)
    for (int a = 0; a < dh; ++a)
      for (int b = 0; b < dh; ++b) sb[(size_t)q * 16 + a * 4 + b] = static_blocks[(size_t)q * dh * dh + a * dh + b];
  free_dev(p->d_e_p1); free_dev(p->d_e_p2); free_dev(p->d_e_fixed); free_dev(p->d_cptr); free_dev(p->d_contrib);
  free_dev(p->d_eT); free_dev(p->d_eom); free_dev(p->d_ew); free_dev(p->d_sblk); free_dev(p->d_eres);
  p->ne = m;
  auto up = [&](auto *&dst, const auto &src) -> cudaError_t {
    using T = typename std::remove_reference<decltype(src[0])>::type;
    cudaError_t e = cudaMalloc(&dst, sizeof(T) * std::max<size_t>(src.size(), 1));
    if (e == cudaSuccess && !src.empty()) e = cudaMemcpy(dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice);
    return e;
  };
  DPGO_CUDA(up(p->d_e_p1, q1)); DPGO_CUDA(up(p->d_e_p2, q2)); DPGO_CUDA(up(p->d_e_fixed, fx)); DPGO_CUDA(up(p->d_cptr, cnt));
  DPGO_CUDA(up(p->d_contrib, contrib)); DPGO_CUDA(up(p->d_eT, eT)); DPGO_CUDA(up(p->d_eom, eom)); DPGO_CUDA(up(p->d_ew, ew));
  DPGO_CUDA(up(p->d_sblk, sb));
  DPGO_CUDA(cudaMalloc(&p->d_eres, sizeof(double) * std::max<int64_t>(m, 1)));
  return reassemble_Q(p);
}

int dpgo_problem_set_edge_weights(dpgo_problem_t *p, const double *weights_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(p->d_cptr, DPGO_ERR_STATE, "dpgo_problem_set_edges has not been called");
  DPGO_REQUIRE(weights_host || p->ne == 0, DPGO_ERR_INVALID_ARG, "null weights");
  if (p->ne) DPGO_CUDA(cudaMemcpyAsync(p->d_ew, weights_host, sizeof(double) * (size_t)p->ne, cudaMemcpyHostToDevice, p->stream));
  return reassemble_Q(p);
}

int dpgo_problem_robust_reweight(dpgo_problem_t *p, int cost, double mu, double param, double *weights_host, double *residuals2_host) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(p->d_cptr, DPGO_ERR_STATE, "dpgo_problem_set_edges has not been called");
  DPGO_REQUIRE(cost >= 0 && cost <= 5, DPGO_ERR_INVALID_ARG, "unknown robust cost");
  DPGO_REQUIRE(cost != 5 || mu > 0, DPGO_ERR_INVALID_ARG, "GNC needs mu > 0");
  DPGO_CUDA(dpgo::launch_edge_weights(p->r, p->dh, p->ne, p->d_e_p1, p->d_e_p2, p->d_eT, p->d_eom, p->d_e_fixed, p->d_vec[dpgo::V_X0], cost,
                                      mu, param, p->d_ew, p->d_eres, p->stream));
  if (weights_host && p->ne)
    DPGO_CUDA(cudaMemcpyAsync(weights_host, p->d_ew, sizeof(double) * (size_t)p->ne, cudaMemcpyDeviceToHost, p->stream));
  if (residuals2_host && p->ne)
    DPGO_CUDA(cudaMemcpyAsync(residuals2_host, p->d_eres, sizeof(double) * (size_t)p->ne, cudaMemcpyDeviceToHost, p->stream));
  return reassemble_Q(p);
}

// ---- plain device helpers ----------------------------------------------------------------------
int dpgo_device_set(int device) {
  DPGO_CUDA(cudaSetDevice(device));
  return DPGO_OK;
}
int dpgo_device_malloc(int device, size_t bytes, void **ptr) {
  DPGO_REQUIRE(ptr, DPGO_ERR_INVALID_ARG, "null argument");
  *ptr = nullptr;
  DPGO_CUDA(cudaSetDevice(device));
  DPGO_CUDA(cudaMalloc(ptr, std::max<size_t>(bytes, 8)));
  DPGO_CUDA(cudaMemset(*ptr, 0, std::max<size_t>(bytes, 8)));
  return DPGO_OK;
}
int dpgo_device_free(int device, void *ptr) {
  DPGO_CUDA(cudaSetDevice(device));
  if (ptr) DPGO_CUDA(cudaFree(ptr));
  return DPGO_OK;
}
int dpgo_stream_create(int device, void **cuda_stream) {
  DPGO_REQUIRE(cuda_stream, DPGO_ERR_INVALID_ARG, "null argument");
  DPGO_CUDA(cudaSetDevice(device));
  cudaStream_t s;
  DPGO_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *cuda_stream = (void *)s;
  return DPGO_OK;
}
int dpgo_stream_destroy(int device, void *cuda_stream) {
  DPGO_CUDA(cudaSetDevice(device));
  if (cuda_stream) DPGO_CUDA(cudaStreamDestroy((cudaStream_t)cuda_stream));
  return DPGO_OK;
}
int dpgo_stream_synchronize(int device, void *cuda_stream) {
  DPGO_CUDA(cudaSetDevice(device));
  DPGO_CUDA(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  return DPGO_OK;
}

// ---- boundary-pose exchange --------------------------------------------------------------------
int dpgo_agent_set_public_poses(dpgo_problem_t *p, int num_public, const int32_t *public_pose) {
  DPGO_CHECK_HANDLE(p);
  ++p->generation;
  DPGO_REQUIRE(num_public >= 0 && (num_public == 0 || public_pose), DPGO_ERR_INVALID_ARG, "bad public pose list");
  for (int s = 0; s < num_public; ++s)
    if (public_pose[s] < 0 || public_pose[s] >= p->n) return fail(DPGO_ERR_INVALID_ARG, "public pose index out of range");
  free_dev(p->d_public);
  p->num_public = num_public;
  if (num_public) {
    DPGO_CUDA(cudaMalloc(&p->d_public, sizeof(int) * num_public));
    DPGO_CUDA(cudaMemcpy(p->d_public, public_pose, sizeof(int) * num_public, cudaMemcpyHostToDevice));
  }
  return DPGO_OK;
}

int dpgo_agent_pack_public(dpgo_problem_t *p, double *send_dev) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(send_dev || p->num_public == 0, DPGO_ERR_INVALID_ARG, "null send buffer");
  DPGO_CUDA(dpgo::launch_pack_tiles(p->ts, p->num_public, p->d_public, p->d_vec[dpgo::V_X0], send_dev, p->stream));
  return DPGO_OK;
}

int dpgo_agent_set_shared_edges(dpgo_problem_t *p, int num_edges, const int32_t *local_pose, const int32_t *nbr_slot,
                                const int32_t *outgoing, const double *T, const double *omega) {
  DPGO_CHECK_HANDLE(p);
  ++p->generation;
  DPGO_REQUIRE(num_edges >= 0 && (num_edges == 0 || (local_pose && nbr_slot && outgoing && T && omega)),
               DPGO_ERR_INVALID_ARG, "bad shared edge arrays");
  const int dh = p->dh;
  for (int e = 0; e < num_edges; ++e)
    if (local_pose[e] < 0 || local_pose[e] >= p->n || nbr_slot[e] < 0)
      return fail(DPGO_ERR_INVALID_ARG, "shared edge index out of range");
  // group edges by local pose, keeping input order inside a pose (= reference accumulation order)
  std::vector<int> order(num_edges);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return local_pose[x] < local_pose[y]; });
  std::vector<int> pose_ids, pose_ptr, slot(num_edges), outg(num_edges);
  std::vector<double> Ts((size_t)num_edges * dh * dh), oms((size_t)num_edges * dh);
  for (int q = 0; q < num_edges; ++q) {
    const int e = order[q];
    if (pose_ids.empty() || pose_ids.back() != local_pose[e]) {
      pose_ids.push_back(local_pose[e]);
      pose_ptr.push_back(q);
    }
    slot[q] = nbr_slot[e];
    outg[q] = outgoing[e] ? 1 : 0;
    std::memcpy(&Ts[(size_t)q * dh * dh], T + (size_t)e * dh * dh, sizeof(double) * dh * dh);
    std::memcpy(&oms[(size_t)q * dh], omega + (size_t)e * dh, sizeof(double) * dh);
  }
  pose_ptr.push_back(num_edges);
  free_dev(p->d_pose_ids); free_dev(p->d_pose_ptr); free_dev(p->d_edge_slot); free_dev(p->d_edge_out);
  free_dev(p->d_edge_T); free_dev(p->d_edge_om);
  p->num_edges = num_edges;
  p->G_dirty = true;
  p->num_shared_poses = (int)pose_ids.size();
  p->max_slot = -1;
  for (int e = 0; e < num_edges; ++e) p->max_slot = std::max(p->max_slot, (int)nbr_slot[e]);
  if (num_edges) {
    DPGO_CUDA(cudaMalloc(&p->d_pose_ids, sizeof(int) * pose_ids.size()));
    DPGO_CUDA(cudaMalloc(&p->d_pose_ptr, sizeof(int) * pose_ptr.size()));
    DPGO_CUDA(cudaMalloc(&p->d_edge_slot, sizeof(int) * num_edges));
    DPGO_CUDA(cudaMalloc(&p->d_edge_out, sizeof(int) * num_edges));
    DPGO_CUDA(cudaMalloc(&p->d_edge_T, sizeof(double) * Ts.size()));
    DPGO_CUDA(cudaMalloc(&p->d_edge_om, sizeof(double) * oms.size()));
    DPGO_CUDA(cudaMemcpy(p->d_pose_ids, pose_ids.data(), sizeof(int) * pose_ids.size(), cudaMemcpyHostToDevice));
    DPGO_CUDA(cudaMemcpy(p->d_pose_ptr, pose_ptr.data(), sizeof(int) * pose_ptr.size(), cudaMemcpyHostToDevice));
    DPGO_CUDA(cudaMemcpy(p->d_edge_slot, slot.data(), sizeof(int) * num_edges, cudaMemcpyHostToDevice));
    DPGO_CUDA(cudaMemcpy(p->d_edge_out, outg.data(), sizeof(int) * num_edges, cudaMemcpyHostToDevice));
    DPGO_CUDA(cudaMemcpy(p->d_edge_T, Ts.data(), sizeof(double) * Ts.size(), cudaMemcpyHostToDevice));
    DPGO_CUDA(cudaMemcpy(p->d_edge_om, oms.data(), sizeof(double) * oms.size(), cudaMemcpyHostToDevice));
  }
  return DPGO_OK;
}

int dpgo_agent_build_G(dpgo_problem_t *p, const double *gathered_dev, int64_t num_slots) {
  DPGO_CHECK_HANDLE(p);
  DPGO_REQUIRE(gathered_dev || p->num_edges == 0, DPGO_ERR_INVALID_ARG, "null gathered buffer");
  DPGO_REQUIRE((int64_t)p->max_slot < num_slots || p->num_edges == 0, DPGO_ERR_INVALID_ARG,
               "a shared edge refers to a slot beyond the gathered buffer (exchange plan / slot table mismatch)");
  // k_build_G assigns every tile of a pose with shared edges; the other tiles of G are zero and stay zero, so G is cleared
  // only when something else may have written it (set_G, a new edge table)
  if (p->G_dirty) {
    DPGO_CUDA(cudaMemsetAsync(p->d_G, 0, p->vec_bytes(), p->stream));
    p->G_dirty = false;
  }
  if (p->num_edges)
    DPGO_CUDA(dpgo::launch_build_G(p->r, p->dh, p->num_shared_poses, p->d_pose_ids, p->d_pose_ptr, p->d_edge_slot,
                                   p->d_edge_out, p->d_edge_T, p->d_edge_om, gathered_dev, p->d_G, p->stream));
  return DPGO_OK;
}

// ---- Nesterov acceleration on the resident iterate ----------------------------------------------------
#define DPGO_ACC_READY(p) DPGO_REQUIRE((p)->d_acc[0], DPGO_ERR_STATE, "dpgo_agent_accel_init has not been called")
int dpgo_agent_accel_init(dpgo_problem_t *p) {
  DPGO_CHECK_HANDLE(p);
  for (int i = 0; i < 3; ++i) {
    if (!p->d_acc[i]) DPGO_CUDA(cudaMalloc(&p->d_acc[i], p->vec_bytes()));
    DPGO_CUDA(cudaMemcpyAsync(p->d_acc[i], p->d_vec[dpgo::V_X0], p->vec_bytes(), cudaMemcpyDeviceToDevice, p->stream));
  }
  return DPGO_OK;
}
int dpgo_agent_accel_begin(dpgo_problem_t *p, double alpha) {
  DPGO_CHECK_HANDLE(p);
  DPGO_ACC_READY(p);
  double *X = p->d_vec[dpgo::V_X0], *Y = p->d_acc[0], *V = p->d_acc[1], *XP = p->d_acc[2];
  DPGO_CUDA(cudaMemcpyAsync(XP, X, p->vec_bytes(), cudaMemcpyDeviceToDevice, p->stream));
  DPGO_CUDA(dpgo::launch_stiefel_project(p->r, p->dh, p->n, X, Y, p->stream, 1.0 - alpha, V, alpha));
  return DPGO_OK;
}
int dpgo_agent_accel_end(dpgo_problem_t *p, double gamma, int optimized) {
  DPGO_CHECK_HANDLE(p);
  DPGO_ACC_READY(p);
  double *X = p->d_vec[dpgo::V_X0], *Y = p->d_acc[0], *V = p->d_acc[1];
  if (!optimized) DPGO_CUDA(cudaMemcpyAsync(X, Y, p->vec_bytes(), cudaMemcpyDeviceToDevice, p->stream));
  DPGO_CUDA(dpgo::launch_stiefel_project(p->r, p->dh, p->n, V, V, p->stream, 1.0, X, gamma, Y, -gamma));
  return DPGO_OK;
}
int dpgo_agent_accel_restart_begin(dpgo_problem_t *p) {
  DPGO_CHECK_HANDLE(p);
  DPGO_ACC_READY(p);
  DPGO_CUDA(cudaMemcpyAsync(p->d_vec[dpgo::V_X0], p->d_acc[2], p->vec_bytes(), cudaMemcpyDeviceToDevice, p->stream));
  return DPGO_OK;
}
int dpgo_agent_accel_restart_end(dpgo_problem_t *p) {
  DPGO_CHECK_HANDLE(p);
  DPGO_ACC_READY(p);
  for (int i = 0; i < 2; ++i)
    DPGO_CUDA(cudaMemcpyAsync(p->d_acc[i], p->d_vec[dpgo::V_X0], p->vec_bytes(), cudaMemcpyDeviceToDevice, p->stream));
  return DPGO_OK;
}
int dpgo_agent_pack_public_aux(dpgo_problem_t *p, double *send_dev) {
  DPGO_CHECK_HANDLE(p);
  DPGO_ACC_READY(p);
  DPGO_REQUIRE(send_dev || p->num_public == 0, DPGO_ERR_INVALID_ARG, "null send buffer");
  DPGO_CUDA(dpgo::launch_pack_tiles(p->ts, p->num_public, p->d_public, p->d_acc[0], send_dev, p->stream));
  return DPGO_OK;
}
int dpgo_optimize_resident_from_aux_async(dpgo_problem_t *p, const dpgo_opt_params_t *params) {
  DPGO_CHECK_HANDLE(p);
  DPGO_ACC_READY(p);
  DPGO_CUDA(cudaMemcpyAsync(p->d_vec[dpgo::V_X0], p->d_acc[0], p->vec_bytes(), cudaMemcpyDeviceToDevice, p->stream));
  return dpgo_optimize_resident_async(p, params);
}

}  // extern "C"

namespace {

struct StreamSwap {                        // the handle's work goes to another stream for the duration of a call
  dpgo_problem *p; cudaStream_t saved;
  StreamSwap(dpgo_problem *q, cudaStream_t to) : p(q), saved(q->stream) { q->stream = to; }
  ~StreamSwap() { p->stream = saved; }
};

// Replay a repeated multi-launch sequence as a CUDA graph.  The graphs live with `lead` (the first agent of the call),
// keyed by everything the captured launches depend on.  First use: eager (warms every lazily created resource);
// second use: captured while it is issued, instantiated and launched; later: one cudaGraphLaunch.
template <class Issue> int replay_or_issue(dpgo_problem *lead, const std::vector<uint64_t> &key, cudaStream_t main, Issue issue) {
  dpgo_problem::RoundGraph *entry = nullptr;
  for (auto &g : lead->round_graphs)
    if (g.key == key) { entry = &g; break; }
  if (!entry) {
    if (lead->round_graphs.size() >= 48) return issue();      // e.g. the greedy schedule on many agents: stay eager
    lead->round_graphs.emplace_back();
    entry = &lead->round_graphs.back();
    entry->key = key;
  }
  if (entry->exec) {
    DPGO_CUDA(cudaGraphLaunch(entry->exec, main));
    return DPGO_OK;
  }
  if (entry->failed || entry->uses++ == 0) return issue();
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(main, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    entry->failed = true;
    return issue();
  }
  const int rc = issue();
  const cudaError_t ce = cudaStreamEndCapture(main, &graph);
  if (rc != DPGO_OK || ce != cudaSuccess || !graph) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    entry->failed = true;
    return issue();
  }
  const cudaError_t ie = cudaGraphInstantiate(&entry->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    cudaGetLastError();
    entry->exec = nullptr;
    entry->failed = true;
    return issue();
  }
  DPGO_CUDA(cudaGraphLaunch(entry->exec, main));
  return DPGO_OK;
}

}  // namespace

extern "C" {

static int issue_round(dpgo_problem_t *const *agents, int num_active, const dpgo_opt_params_t *params,
                       const double *gathered_dev, int64_t num_slots, double *const *send_dev, cudaStream_t main,
                       int pack_after_join) {
  dpgo_problem *lead = agents[0];
  const int passes = pack_after_join ? 2 : 1;
  for (int pass = 0; pass < passes; ++pass) {
    DPGO_CUDA(cudaEventRecord(lead->ev_fork, main));
    for (int i = 0; i < num_active; ++i) {
      dpgo_problem *p = agents[i];
      StreamSwap swap(p, p->own_stream);      // the agent's kernels go to its own stream
      if (p->stream != main) DPGO_CUDA(cudaStreamWaitEvent(p->stream, lead->ev_fork, 0));
      if (pass == 0) {
        DPGO_TRY(dpgo_agent_build_G(p, gathered_dev, num_slots));
        DPGO_TRY(dpgo_optimize_resident_async(p, params));
      }
      if (pass == passes - 1) DPGO_TRY(dpgo_agent_pack_public(p, send_dev[i]));
      if (p->stream != main) {
        DPGO_CUDA(cudaEventRecord(p->ev_done, p->stream));
        DPGO_CUDA(cudaStreamWaitEvent(main, p->ev_done, 0));
      }
    }
  }
  return DPGO_OK;
}

// One RBCD round of the agents of one GPU, issued with one call.  Every active agent works on its OWN stream:
//   main stream --fork--> [ G rebuild from the gathered tiles -> RTR step (persistent kernel) -> pack of its public tiles ] --join--> main
// so agents launched as single thread-block clusters (dpgo_problem_set_launch_mode(p, 1)) share the GPU: up to 8 clusters
// of 16 CTAs run side by side.  Agents of one colour class are never neighbours, so a pack into the (aliased) gathered
// buffer cannot race with another active agent's G rebuild; with pack_after_join != 0 (every agent active on the previous
// round's poses) the packs are issued in a second fork/join instead.
// A round with the same agents, buffers and parameters as an earlier one is replayed as a CUDA graph (the cluster launches
// are ordinary launches, so the fork/join captures): 1 driver call per round instead of ~7 per agent, which is what
// bounds 8 agents x ~100 us of GPU work otherwise.  DPGO_ROUND_GRAPH=0 keeps the eager launches.
int dpgo_agents_round_async(dpgo_problem_t *const *agents, int num_active, const dpgo_opt_params_t *params,
                            const double *gathered_dev, int64_t num_slots, double *const *send_dev, void *main_stream,
                            int pack_after_join) {
  DPGO_REQUIRE(num_active >= 0 && (num_active == 0 || (agents && send_dev)) && params, DPGO_ERR_INVALID_ARG, "bad arguments");
  if (num_active == 0) return DPGO_OK;
  bool graphable = true;
  for (int i = 0; i < num_active; ++i) {
    DPGO_CHECK_HANDLE(agents[i]);
    DPGO_REQUIRE(agents[i]->device == agents[0]->device, DPGO_ERR_INVALID_ARG, "the agents of a round must live on one device");
    DPGO_TRY(check_params(agents[i], params));
    if (!agents[i]->ev_done) DPGO_CUDA(cudaEventCreateWithFlags(&agents[i]->ev_done, cudaEventDisableTiming));
    // a cooperative launch does not capture; a pending G clear or an unbuilt factorisation must run eagerly first
    if (!agents[i]->cluster || agents[i]->G_dirty || agents[i]->d_phase_ns ||
        ((agents[i]->precond_mask & (1u << DPGO_PRECOND_SPARSE_EXACT)) && !agents[i]->nd_ready))
      graphable = false;
  }
  dpgo_problem *lead = agents[0];
  cudaStream_t main = main_stream ? (cudaStream_t)main_stream : lead->stream;   // NULL: the stream the first handle is set to
  DPGO_CUDA(cudaSetDevice(lead->device));
  if (!lead->ev_fork) DPGO_CUDA(cudaEventCreateWithFlags(&lead->ev_fork, cudaEventDisableTiming));
  static const bool use_graph = [] { const char *e = std::getenv("DPGO_ROUND_GRAPH"); return !e || std::atoi(e) != 0; }();
  if (main == cudaStreamLegacy || main == nullptr) graphable = false;   // the legacy default stream cannot be captured
  if (!use_graph || !graphable)
    return issue_round(agents, num_active, params, gathered_dev, num_slots, send_dev, main, pack_after_join);

  std::vector<uint64_t> key;
  key.reserve(3 * (size_t)num_active + 8 + sizeof(*params) / 8 + 1);
  for (int i = 0; i < num_active; ++i) {
    key.push_back((uint64_t)(uintptr_t)agents[i]);
    key.push_back(agents[i]->generation);
    key.push_back((uint64_t)(uintptr_t)send_dev[i]);
  }
  key.push_back((uint64_t)(uintptr_t)gathered_dev);
  key.push_back((uint64_t)num_slots);
  key.push_back((uint64_t)(uintptr_t)main);
  key.push_back((uint64_t)pack_after_join);
  {
    uint64_t w[(sizeof(*params) + 7) / 8] = {};
    std::memcpy(w, params, sizeof(*params));
    key.insert(key.end(), w, w + sizeof(w) / 8);
  }
  auto issue = [&]() { return issue_round(agents, num_active, params, gathered_dev, num_slots, send_dev, main, pack_after_join); };
  return replay_or_issue(lead, key, main, issue);
}

// The host boundary of a round with one call per direction (the end-to-end path of DistributedPGO.step_host):
//   direction 0: X of every listed agent from (pinned) host memory, then its public tiles packed into send_dev[i]
//   direction 1: X of every listed agent back to host memory
// all on `stream`; a repeated call (same agents, buffers, stream) is replayed as a CUDA graph of memcpy / kernel nodes.
int dpgo_agents_host_io_async(dpgo_problem_t *const *agents, int count, double *const *X_host, double *const *send_dev,
                              int direction, void *stream) {
  DPGO_REQUIRE(count >= 0 && (count == 0 || (agents && X_host)) && (direction == 0 || direction == 1), DPGO_ERR_INVALID_ARG,
               "bad arguments");
  if (count == 0) return DPGO_OK;
  for (int i = 0; i < count; ++i) {
    DPGO_CHECK_HANDLE(agents[i]);
    DPGO_REQUIRE(X_host[i], DPGO_ERR_INVALID_ARG, "null host buffer");
    DPGO_REQUIRE(agents[i]->device == agents[0]->device, DPGO_ERR_INVALID_ARG, "the agents of a call must live on one device");
  }
  dpgo_problem *lead = agents[0];
  cudaStream_t main = stream ? (cudaStream_t)stream : lead->stream;
  DPGO_CUDA(cudaSetDevice(lead->device));
  if (!lead->ev_fork) DPGO_CUDA(cudaEventCreateWithFlags(&lead->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < count; ++i)
    if (!agents[i]->ev_done) DPGO_CUDA(cudaEventCreateWithFlags(&agents[i]->ev_done, cudaEventDisableTiming));
  auto issue = [&]() -> int {
    // every agent's copy (+ pack) on its own stream between a fork from and a join into `main`: the copies of different
    // agents overlap each other and the packs
    DPGO_CUDA(cudaEventRecord(lead->ev_fork, main));
    for (int i = 0; i < count; ++i) {
      dpgo_problem *p = agents[i];
      StreamSwap swap(p, p->own_stream);
      if (p->stream != main) DPGO_CUDA(cudaStreamWaitEvent(p->stream, lead->ev_fork, 0));
      if (direction == 0) {
        DPGO_TRY(upload_vec(p, dpgo::V_X0, X_host[i]));
        if (send_dev && send_dev[i]) DPGO_TRY(dpgo_agent_pack_public(p, send_dev[i]));
      } else {
        DPGO_TRY(download_vec(p, dpgo::V_X0, X_host[i]));
      }
      if (p->stream != main) {
        DPGO_CUDA(cudaEventRecord(p->ev_done, p->stream));
        DPGO_CUDA(cudaStreamWaitEvent(main, p->ev_done, 0));
      }
    }
    return DPGO_OK;
  };
  static const bool use_graph = [] { const char *e = std::getenv("DPGO_ROUND_GRAPH"); return !e || std::atoi(e) != 0; }();
  if (!use_graph || main == cudaStreamLegacy || main == nullptr) return issue();
  std::vector<uint64_t> key;
  key.reserve(3 * (size_t)count + 4);
  key.push_back(0x696f0000ull + (uint64_t)direction);       // "io": never equal to a round key (those start with a pointer)
  for (int i = 0; i < count; ++i) {
    key.push_back((uint64_t)(uintptr_t)agents[i]);
    key.push_back(agents[i]->generation);
    key.push_back((uint64_t)(uintptr_t)X_host[i]);
    key.push_back((uint64_t)(uintptr_t)((send_dev && direction == 0) ? send_dev[i] : nullptr));
  }
  key.push_back((uint64_t)(uintptr_t)main);
  return replay_or_issue(lead, key, main, issue);
}

int dpgo_agent_f_rgradnorm_resident(dpgo_problem_t *p, double *f_out, double *norm_out) {
  DPGO_CHECK_HANDLE(p);
  dpgo_opt_params_t prm;
  dpgo_opt_params_default(&prm);
  prm.precond = DPGO_PRECOND_NONE;
  DPGO_TRY(run_op(p, dpgo::OP_EVAL, prm));
  DPGO_TRY(fetch_result(p));
  if (f_out) *f_out = p->h_result->f_init;
  if (norm_out) *norm_out = p->h_result->gradnorm_init;
  return DPGO_OK;
}

}  // extern "C"
