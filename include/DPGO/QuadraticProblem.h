// QuadraticProblem.h -- f(X) = 0.5 <Q, X^T X> + <X, G> on the lifted SE manifold, resident on one B200.
// Same public interface as the reference's include/DPGO/QuadraticProblem.h:31-109 (ctor, setQ/setG, f,
// RieGrad, RieGradNorm, getters); the ROPTLIB virtuals become plain methods because the solver lives in
// the CUDA library (include/dpgo_b200.h), not in ROPTLIB.
#ifndef DPGO_B200_QUADRATICPROBLEM_H
#define DPGO_B200_QUADRATICPROBLEM_H

#include <DPGO/DPGO_types.h>
#include <DPGO/manifold/LiftedSEManifold.h>

struct dpgo_problem;   // opaque C-ABI handle

namespace DPGO {

class QuadraticProblem {
 public:
  QuadraticProblem(size_t nIn, size_t dIn, size_t rIn);
  virtual ~QuadraticProblem();
  QuadraticProblem(const QuadraticProblem &) = delete;
  QuadraticProblem &operator=(const QuadraticProblem &) = delete;

  unsigned int num_poses() const { return n; }
  unsigned int dimension() const { return d; }
  unsigned int relaxation_rank() const { return r; }

  SparseMatrix getQ() const { return mQ; }
  SparseMatrix getG() const { return mG; }
  void setQ(const SparseMatrix &QIn);
  void setG(const SparseMatrix &GIn);
  void setG(const Matrix &GDense);           // dense r x (d+1)n form (avoids the sparse detour)

  double f(const Matrix &Y) const;
  Matrix EucGrad(const Matrix &Y) const;               // Y Q + G
  Matrix EucHessianEta(const Matrix &V) const;         // V Q
  Matrix RieGrad(const Matrix &Y) const;
  double RieGradNorm(const Matrix &Y) const;
  Matrix RieHessianEta(const Matrix &Y, const Matrix &V) const;
  Matrix PreConditioner(const Matrix &Y, const Matrix &V) const;   // P_Y((Q + 0.1 I)^-1 V)

  // B200 extensions
  void setDevice(int device);                           // before the first setQ; default: env DPGO_DEVICE or 0
  void setClusterLaunch(bool on);                       // before the first setQ: step kernel as ONE thread-block cluster (dpgo_problem_set_launch_mode)
  void setPreconditioners(bool blockJacobi, bool exact, Preconditioner exactKind = Preconditioner::SparseExact);
  dpgo_problem *handle() const { return mHandle; }
  static int defaultDevice();

 private:
  const size_t n = 0, d = 0, r = 0;
  SparseMatrix mQ, mG;
  int mDevice;
  bool mCluster = false;
  unsigned mPrecondMask;
  mutable dpgo_problem *mHandle = nullptr;
  void ensureHandle() const;
};

}  // namespace DPGO
#endif
