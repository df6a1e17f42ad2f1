// LiftedSEManifold.h -- the product manifold (St(d,r) x R^r)^n.  The reference wraps ROPTLIB objects here
// (include/DPGO/manifold/LiftedSEManifold.h:19-41); on the B200 the manifold operations are kernels of
// libdpgo_b200.so, so this class only carries the dimensions and the host-visible project().
#ifndef DPGO_B200_LIFTEDSEMANIFOLD_H
#define DPGO_B200_LIFTEDSEMANIFOLD_H

#include <DPGO/DPGO_types.h>

namespace DPGO {

class LiftedSEManifold {
 public:
  LiftedSEManifold(int r, int d, int n) : r_(r), d_(d), n_(n) {}
  // per-pose orthogonal projection of the r x d blocks onto the Stiefel manifold (GPU kernel)
  Matrix project(const Matrix &M) const;
  // tangent-space projection at X and QF retraction (GPU kernels)
  Matrix tangentProject(const Matrix &X, const Matrix &Z) const;
  Matrix retract(const Matrix &X, const Matrix &eta) const;

 private:
  size_t r_, d_, n_;
};

}  // namespace DPGO
#endif
