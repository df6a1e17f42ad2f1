// QuadraticProblem.cpp -- host mirror of the reference's QuadraticProblem over the C ABI
// (include/dpgo_b200.h).  Every numeric method is one call into libdpgo_b200.so (sm_100a kernels); there is
// no CPU implementation behind it.  ref: src/QuadraticProblem.cpp:15-109.
#include <DPGO/QuadraticProblem.h>

#include <cstdlib>
#include <stdexcept>

#include <map>
#include <memory>
#include <tuple>

#include "dpgo_b200.h"

namespace DPGO {

namespace {
void check(int code, const char *what) {
  if (code != DPGO_OK) throw std::runtime_error(std::string(what) + ": " + dpgo_last_error());
}
}  // namespace

int QuadraticProblem::defaultDevice() {
  const char *e = std::getenv("DPGO_DEVICE");
  return e ? std::atoi(e) : 0;
}

QuadraticProblem::QuadraticProblem(size_t nIn, size_t dIn, size_t rIn)
    : n(nIn), d(dIn), r(rIn), mQ((Eigen::Index)((dIn + 1) * nIn), (Eigen::Index)((dIn + 1) * nIn)),
      mG((Eigen::Index)rIn, (Eigen::Index)((dIn + 1) * nIn)), mDevice(defaultDevice()),
      mPrecondMask((1u << DPGO_PRECOND_BLOCK_JACOBI) | (1u << DPGO_PRECOND_SPARSE_EXACT)) {
  assert(r >= d);
}

QuadraticProblem::~QuadraticProblem() {
  if (mHandle) dpgo_problem_destroy(mHandle);
}

void QuadraticProblem::ensureHandle() const {
  if (mHandle) return;
  check(dpgo_problem_create((int)n, (int)d, (int)r, mDevice, &mHandle), "dpgo_problem_create");
  if (mCluster) check(dpgo_problem_set_launch_mode(mHandle, 1), "dpgo_problem_set_launch_mode");
}

void QuadraticProblem::setClusterLaunch(bool on) {
  if (mHandle) throw std::runtime_error("QuadraticProblem::setClusterLaunch must precede the first use");
  mCluster = on;
}

void QuadraticProblem::setDevice(int device) {
  if (mHandle) throw std::runtime_error("QuadraticProblem::setDevice must precede the first use");
  mDevice = device;
}

void QuadraticProblem::setPreconditioners(bool blockJacobi, bool exact, Preconditioner exactKind) {
  const unsigned exactBit = (exactKind == Preconditioner::DenseExact) ? (1u << DPGO_PRECOND_DENSE_EXACT) : (1u << DPGO_PRECOND_SPARSE_EXACT);
  mPrecondMask = (blockJacobi ? (1u << DPGO_PRECOND_BLOCK_JACOBI) : 0u) | (exact ? exactBit : 0u);
  if (mHandle && mQ.nonZeros() > 0) setQ(SparseMatrix(mQ));
}

void QuadraticProblem::setQ(const SparseMatrix &QIn) {
  assert((unsigned)QIn.rows() == (d + 1) * n && (unsigned)QIn.cols() == (d + 1) * n);
  mQ = QIn;
  ensureHandle();
  // uploads Q as block-CSR and prepares the preconditioners (the reference factors Q + 0.1 I with CHOLMOD here)
  check(dpgo_problem_set_Q_csr(mHandle, (int)mQ.rows(), mQ.outerIndexPtr(), mQ.innerIndexPtr(), mQ.valuePtr(), mPrecondMask),
        "dpgo_problem_set_Q_csr");
}

void QuadraticProblem::setG(const SparseMatrix &GIn) {
  assert((unsigned)GIn.rows() == r && (unsigned)GIn.cols() == (d + 1) * n);
  mG = GIn;
  ensureHandle();
  check(dpgo_problem_set_G_csr(mHandle, mG.outerIndexPtr(), mG.innerIndexPtr(), mG.valuePtr()), "dpgo_problem_set_G_csr");
}

void QuadraticProblem::setG(const Matrix &GDense) {
  assert((unsigned)GDense.rows() == r && (unsigned)GDense.cols() == (d + 1) * n);
  ensureHandle();
  check(dpgo_problem_set_G_dense(mHandle, GDense.data()), "dpgo_problem_set_G_dense");
}

double QuadraticProblem::f(const Matrix &Y) const {
  assert((unsigned)Y.rows() == r && (unsigned)Y.cols() == (d + 1) * n);
  ensureHandle();
  double out = 0;
  check(dpgo_problem_f(mHandle, Y.data(), &out), "dpgo_problem_f");
  return out;
}

Matrix QuadraticProblem::EucGrad(const Matrix &Y) const {
  ensureHandle();
  Matrix out((Eigen::Index)r, (Eigen::Index)((d + 1) * n));
  check(dpgo_problem_egrad(mHandle, Y.data(), out.data()), "dpgo_problem_egrad");
  return out;
}

Matrix QuadraticProblem::EucHessianEta(const Matrix &Vin) const {
  ensureHandle();
  Matrix out((Eigen::Index)r, (Eigen::Index)((d + 1) * n));
  check(dpgo_problem_ehess(mHandle, Vin.data(), out.data()), "dpgo_problem_ehess");
  return out;
}

Matrix QuadraticProblem::RieGrad(const Matrix &Y) const {
  ensureHandle();
  Matrix out((Eigen::Index)r, (Eigen::Index)((d + 1) * n));
  check(dpgo_problem_rgrad(mHandle, Y.data(), out.data(), nullptr), "dpgo_problem_rgrad");
  return out;
}

double QuadraticProblem::RieGradNorm(const Matrix &Y) const {
  ensureHandle();
  double nrm = 0;
  check(dpgo_problem_rgrad(mHandle, Y.data(), nullptr, &nrm), "dpgo_problem_rgrad");
  return nrm;
}

Matrix QuadraticProblem::RieHessianEta(const Matrix &Y, const Matrix &Vin) const {
  ensureHandle();
  Matrix out((Eigen::Index)r, (Eigen::Index)((d + 1) * n));
  check(dpgo_problem_rhess(mHandle, Y.data(), Vin.data(), out.data()), "dpgo_problem_rhess");
  return out;
}

Matrix QuadraticProblem::PreConditioner(const Matrix &Y, const Matrix &Vin) const {
  ensureHandle();
  Matrix out((Eigen::Index)r, (Eigen::Index)((d + 1) * n));
  const int which = (mPrecondMask & (1u << DPGO_PRECOND_SPARSE_EXACT)) ? DPGO_PRECOND_SPARSE_EXACT
                    : (mPrecondMask & (1u << DPGO_PRECOND_DENSE_EXACT)) ? DPGO_PRECOND_DENSE_EXACT : DPGO_PRECOND_BLOCK_JACOBI;
  if (dpgo_problem_precon(mHandle, which, Y.data(), Vin.data(), out.data()) != DPGO_OK) {
    printf("Preconditioner failed.\n");      // ref :83-86: fall back to the identity
    return Vin;
  }
  return out;
}

// ---- LiftedSEManifold (ref src/manifold/LiftedSEManifold.cpp:34-45) ------------------------------------
namespace {
// One device scratch problem per (n, d, r, device) and thread, created on first use and kept: the manifold operations
// sit on the Nesterov per-iteration path (PGOAgent::updateY / updateV call project() for every agent and iterate),
// where creating and destroying a handle per call would mean ~16 cudaMalloc / cudaFree with device-wide syncs.
struct ScratchProblem {
  dpgo_problem *h = nullptr;
  ScratchProblem(size_t n, size_t d, size_t r, int device) {
    check(dpgo_problem_create((int)n, (int)d, (int)r, device, &h), "dpgo_problem_create");
  }
  ~ScratchProblem() { dpgo_problem_destroy(h); }
  ScratchProblem(const ScratchProblem &) = delete;
  ScratchProblem &operator=(const ScratchProblem &) = delete;
};
struct ScratchRef { dpgo_problem *h; };
ScratchRef scratch(size_t n, size_t d, size_t r) {
  static thread_local std::map<std::tuple<size_t, size_t, size_t, int>, std::unique_ptr<ScratchProblem>> cache;
  const int device = QuadraticProblem::defaultDevice();
  auto key = std::make_tuple(n, d, r, device);
  auto it = cache.find(key);
  if (it == cache.end()) {
    if (cache.size() >= 8) cache.clear();              // bounded: the agents of one process share few shapes
    it = cache.emplace(key, std::unique_ptr<ScratchProblem>(new ScratchProblem(n, d, r, device))).first;
  }
  return ScratchRef{it->second->h};
}
}  // namespace

Matrix LiftedSEManifold::project(const Matrix &M) const {
  assert(M.rows() == (int)r_ && M.cols() == (int)((d_ + 1) * n_));
  ScratchRef sp = scratch(n_, d_, r_);
  Matrix out(M.rows(), M.cols());
  check(dpgo_manifold_project(sp.h, M.data(), out.data()), "dpgo_manifold_project");
  return out;
}

Matrix LiftedSEManifold::tangentProject(const Matrix &X, const Matrix &Z) const {
  ScratchRef sp = scratch(n_, d_, r_);
  Matrix out(X.rows(), X.cols());
  check(dpgo_manifold_tangent_project(sp.h, X.data(), Z.data(), out.data()), "dpgo_manifold_tangent_project");
  return out;
}

Matrix LiftedSEManifold::retract(const Matrix &X, const Matrix &eta) const {
  ScratchRef sp = scratch(n_, d_, r_);
  Matrix out(X.rows(), X.cols());
  check(dpgo_manifold_retract(sp.h, X.data(), eta.data(), out.data()), "dpgo_manifold_retract");
  return out;
}

}  // namespace DPGO
