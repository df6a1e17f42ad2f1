// QuadraticOptimizer.h -- one RTR or RGD call on a QuadraticProblem; same interface as the reference's
// include/DPGO/QuadraticOptimizer.h:20-76.  optimize() is ONE persistent CUDA kernel (dpgo_optimize).
#ifndef DPGO_B200_QUADRATICOPTIMIZER_H
#define DPGO_B200_QUADRATICOPTIMIZER_H

#include <DPGO/DPGO_types.h>
#include <DPGO/QuadraticProblem.h>

namespace DPGO {

class QuadraticOptimizer {
 public:
  explicit QuadraticOptimizer(QuadraticProblem *p);
  ~QuadraticOptimizer();

  Matrix optimize(const Matrix &Y);

  void setProblem(QuadraticProblem *p) { problem = p; }
  void setVerbose(bool v) { verbose = v; }
  void setAlgorithm(ROPTALG alg) { algorithm = alg; }
  void setGradientDescentStepsize(double s) { gradientDescentStepsize = s; }
  void setTrustRegionIterations(unsigned iter) { trustRegionIterations = iter; }
  void setTrustRegionTolerance(double tol) { trustRegionTolerance = tol; }
  void setTrustRegionInitialRadius(double radius) { trustRegionInitialRadius = radius; }
  void setTrustRegionMaxInnerIterations(int iter) { trustRegionMaxInnerIterations = iter; }
  void setPreconditioner(Preconditioner pc) { preconditioner = pc; }   // B200 extension
  ROPTResult getOptResult() const { return result; }

 private:
  QuadraticProblem *problem;
  ROPTALG algorithm;
  ROPTResult result;
  double gradientDescentStepsize;
  unsigned trustRegionIterations;
  double trustRegionTolerance;
  double trustRegionInitialRadius;
  int trustRegionMaxInnerIterations;
  Preconditioner preconditioner;
  bool verbose;
};

}  // namespace DPGO
#endif
