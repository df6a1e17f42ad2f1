// nd_precond.h -- host side of the sparse EXACT preconditioner  Z = (Q + shift I)^-1 V  (r right-hand sides at once).
//
// ref: QuadraticProblem::setQ (src/QuadraticProblem.cpp:31-42: CHOLMOD factorisation of Q + 0.1 I) and
//      QuadraticProblem::PreConditioner (:75-87: solve, then tangent projection).
//
// B200 design: a GPU triangular solve along an elimination tree is a chain of ~2 log2(n) dependent grid-wide steps,
// and a grid-wide step costs >= 1.2 us on B200 whatever the protocol (scripts/barrier_bench2.cu).  So the tree is
// flattened: a nested-dissection tree of the pose graph is cut into a FEW macro levels (usually 2-3); every macro node v
// (a fragment of the dissection tree: its separators / leaf interiors) keeps
//     W_v = S_v^-1        dense inverse of its Schur complement (own x own),
//     F_v = W_v E_v       coupling to the ancestors' variables it touches (own x bnd),
// so that the solve is  2 * levels - 1  phases of small dense panel products:
//     forward  (leaves -> root):  y_v = b_v - sum_children c_child ;  t_v = W_v y_v ;  c_v = F_v^T y_v + pass-through
//     root:                       x_root = W_root y_root
//     backward (root -> leaves):  x_v = t_v - F_v x_bnd(v)
// All blocks together are ~10x the sparse factor but ~15-30x smaller than the dense inverse, and L2-resident for
// sphere2500-sized agents.  Matrices are stored as 8-row panels, column-major inside a panel (a warp reads 256
// contiguous bytes per step); the work of every phase is a host-built static plan (per CTA: steps of gathers into
// shared memory / warp jobs = panel x column piece / epilogues per pose), interpreted by phase_nd in dpgo_kernels.cu.
// Everything is deterministic (fixed summation orders).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace dpgo {
namespace nd {

// ---- device-facing plan records (plain ints, uploaded as they are) ------------------------------------------
// Records are sized in whole 16-byte words and carry what the kernel needs WITHOUT a second dependent load: the first
// step of a CTA sits in its (phase, CTA) record, the first four contribution tiles sit in the gather / epilogue record.
struct Step { int g0, g1, j0, j1, e0, e1, pad0, pad1; };              // ranges into gathers / jobs / epis
struct CtaPhase { int s0, s1; int g0, g1, j0, j1, e0, e1; };          // steps [s0, s1) of one CTA in one phase + step s0 inline
constexpr int INLINE_CONTRIB = 4;
struct Gather { int ytile; int src; int nc; int cext; int ci[INLINE_CONTRIB]; };   // smem tile <- source tile - sum of nc contribution tiles
struct Job { long long mat; int ncols; int ycol; int slot; int accum; int pad0, pad1; };   // one warp: panel piece x y
struct Epi { int kind; int slot0; int nslots; int half; int out; int aux; int nc; int cext; int ci[INLINE_CONTRIB]; };   // one pose (dh rows of a panel)
enum EpiKind { EPI_F_OWN = 0, EPI_F_BND = 1, EPI_B_OWN = 2, EPI_ROOT = 3 };
struct Phase { int dir; int stage; int cta0; int pad; };              // dir 0 forward (source = V, pose ids), 1 backward (source = TX); cta0 = first CtaPhase record
constexpr int MAX_PHASES = 16;

constexpr int PANEL_ROWS = 8;

struct Options {
  int grid = 148;            // CTAs of the persistent kernel
  int r = 5;                 // right-hand sides (rows of the residual)
  int warps = 16;            // warps per CTA
  int leaf_size = 12;        // dissection stops below this many poses
  int max_cuts = 3;          // macro levels <= max_cuts + 1
  int ycap_tiles = 600;      // shared-memory capacity for gathered tiles per step
  int slot_cap = 240;        // shared-memory partial-sum slots (8 x r doubles each) per step
  double t_phase_us = 3.5;   // cost model: fixed part of one phase (grid barrier + the stages' dependent L2 round trips)
  double t_tile_us = 0.012;  // cost model: staging one tile of a phase's largest input vector (per CTA, redundant over CTAs)
  double bw_gbs = 3000.0;    // cost model: effective streaming bandwidth of the panel products
  int force_ncuts = -1;      // >= 0: use exactly this many cuts (tests)
  double shift = 0.1;
};

struct MacroNode {
  int stage = 0;             // 0 = deepest macro level ... nstages-1 = root level
  int parent = -1;
  std::vector<int> own;      // pose ids, elimination order
  std::vector<int> bnd;      // pose ids owned by ancestors that the Schur complement touches
  std::vector<int> children;
  int perm0 = 0;             // first permuted tile index of own
  int cbuf0 = 0;             // first contribution tile of this node
  int64_t gf_off = 0, gb_off = 0;   // panel blobs (doubles): forward [W ; F^T] (rows own+bnd, cols own), backward F (rows own, cols bnd)
};

struct Hierarchy {
  int n = 0, dh = 4;
  std::vector<MacroNode> nodes;
  int nstages = 0;
  std::vector<int> perm;       // permuted tile index -> pose id
  std::vector<int> iperm;      // pose id -> permuted tile index
  std::vector<int> node_of;    // pose id -> macro node
  int cbuf_tiles = 0;
  int64_t blob_doubles = 0;
  std::vector<int> cuts;       // dissection levels where macro levels start (diagnostic)
  int nd_depth = 0;
};

struct Plan {
  std::vector<Phase> phases;       // 2 * nstages - 1
  std::vector<CtaPhase> cta_phase; // phases * grid
  std::vector<Step> steps;
  std::vector<Gather> gathers;
  std::vector<Job> jobs;
  std::vector<Epi> epis;
  std::vector<int> csrc;           // contribution tile ids beyond the INLINE_CONTRIB kept in the records
  int grid = 0, r = 0;
  int max_ytiles = 0, max_slots = 0;
  int64_t bytes_per_apply = 0;     // matrix bytes streamed by one application (all phases)
};

// block-CSR input: rowptr[n+1], bcol[nb], bval[nb*16] with bval[b][k][c] = Q[dh*bcol[b]+k, dh*j+c] for b in row j
struct BsrView { int n; int dh; const int *rowptr; const int *bcol; const double *bval; };

// 1. ordering + macro levels (symbolic).  Throws std::runtime_error on impossible input.
void build_hierarchy(const BsrView &Q, const Options &opt, Hierarchy &H);
// 2. numeric: fills the panel blob (host, OpenMP over columns).  big_node (optional) may take over the dense algebra
//    of one node (device offload); return false to let the host do it.
struct DenseNodeOps {
  // in: Foo (s x s, row-major, SPD), Fob (s x b), Fbb (b x b);  out: W = Foo^-1 (s x s), Fm = W Fob (s x b), Fbb -= Fob^T Fm
  virtual bool factor(int s, int b, double *Foo, double *Fob, double *Fbb) = 0;
  virtual ~DenseNodeOps() {}
};
void build_numeric(const BsrView &Q, const Options &opt, Hierarchy &H, std::vector<double> &blob, DenseNodeOps *big_node = nullptr,
                   int big_threshold = 1 << 30);
// 3. static work plan of every phase
void build_plan(const Hierarchy &H, const Options &opt, Plan &P);
// host emulation of the plan exactly as the kernel interprets it (verification only; never on a product path):
// V, Z are r x (dh n) column-major (pose tiles), Z = (Q + shift I)^-1 V   (no tangent projection)
void emulate_apply(const Hierarchy &H, const Plan &P, const std::vector<double> &blob, int r, const double *V, double *Z);

std::string describe(const Hierarchy &H, const Plan &P);

}  // namespace nd
}  // namespace dpgo
