// QuadraticOptimizer.h -- one RTR or RGD call on a QuadraticProblem; same interface as the reference's
// include/DPGO/QuadraticOptimizer.h:20-76.  optimize() is ONE persistent CUDA kernel (dpgo_optimize).
#ifndef DPGO_B200_QUADRATICOPTIMIZER_H
#define DPGO_B200_QUADRATICOPTIMIZER_H

#include <DPGO/DPGO_types.h>
#include <DPGO/QuadraticProblem.h>

namespace DPGO {

class QuadraticOptimizer {
 public:
  // The optimizer borrows the problem (no ownership), as in the reference (:80).
  explicit QuadraticOptimizer(QuadraticProblem *p);
  ~QuadraticOptimizer();

  // One call = one launch of the persistent kernel on the problem's device: uploads Y, runs either a single
  // fixed-step gradient-descent step or `trustRegionIterations` trust-region steps (truncated CG inside), and returns
  // the new iterate.  Statistics of the call (f and |grad| before / after, tCG exit) are kept for getOptResult().
  Matrix optimize(const Matrix &Y);

  void setProblem(QuadraticProblem *p) { problem = p; }
  void setVerbose(bool v) { verbose = v; }
  void setAlgorithm(ROPTALG alg) { algorithm = alg; }                      // RTR or RGD
  void setGradientDescentStepsize(double s) { gradientDescentStepsize = s; }
  // iterations == 1 selects the reference's shrink-the-radius-until-accepted mode (PGOAgent::updateX)
  void setTrustRegionIterations(unsigned iter) { trustRegionIterations = iter; }
  void setTrustRegionTolerance(double tol) { trustRegionTolerance = tol; }  // on the Riemannian gradient norm
  void setTrustRegionInitialRadius(double radius) { trustRegionInitialRadius = radius; }
  void setTrustRegionMaxInnerIterations(int iter) { trustRegionMaxInnerIterations = iter; }
  // B200 extension: which M^-1 the truncated CG uses (default: the reference's exact (Q + 0.1 I)^-1 operator)
  void setPreconditioner(Preconditioner pc) { preconditioner = pc; }
  ROPTResult getOptResult() const { return result; }

 private:
  QuadraticProblem *problem;
  ROPTALG algorithm;
  ROPTResult result;
  Preconditioner preconditioner;
  double gradientDescentStepsize;
  double trustRegionTolerance;
  double trustRegionInitialRadius;
  unsigned trustRegionIterations;
  int trustRegionMaxInnerIterations;
  bool verbose;
};

}  // namespace DPGO
#endif
