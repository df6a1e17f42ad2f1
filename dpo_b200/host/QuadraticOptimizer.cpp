// QuadraticOptimizer.cpp -- one RTR / RGD call = one dpgo_optimize() = one persistent CUDA kernel.
// ref: src/QuadraticOptimizer.cpp:20-59 (defaults, statistics, fOpt <= fInit assertion).
#include <DPGO/QuadraticOptimizer.h>

#include <cassert>
#include <stdexcept>

#include "dpgo_b200.h"

namespace DPGO {

QuadraticOptimizer::QuadraticOptimizer(QuadraticProblem *p)
    : problem(p), algorithm(ROPTALG::RTR), preconditioner(Preconditioner::SparseExact), gradientDescentStepsize(1e-3),
      trustRegionTolerance(1e-2), trustRegionInitialRadius(1e1), trustRegionIterations(1),
      trustRegionMaxInnerIterations(50), verbose(false) {
  result.success = false;
}

QuadraticOptimizer::~QuadraticOptimizer() = default;

Matrix QuadraticOptimizer::optimize(const Matrix &Y) {
  dpgo_opt_params_t prm;
  dpgo_opt_params_default(&prm);
  prm.algorithm = (algorithm == ROPTALG::RTR) ? DPGO_ALG_RTR : DPGO_ALG_RGD;
  prm.rgd_stepsize = gradientDescentStepsize;
  prm.tr_iterations = (int)trustRegionIterations;
  prm.tr_tolerance = trustRegionTolerance;
  prm.tr_initial_radius = trustRegionInitialRadius;
  prm.tr_max_inner = trustRegionMaxInnerIterations;
  prm.precond = (int)preconditioner;
  dpgo_opt_result_t res;
  Matrix YOpt(Y.rows(), Y.cols());
  if (!problem->handle()) (void)problem->f(Y);      // materialise the device problem (empty Q) if never used
  if (dpgo_optimize(problem->handle(), &prm, Y.data(), YOpt.data(), &res) != DPGO_OK)
    throw std::runtime_error(std::string("dpgo_optimize: ") + dpgo_last_error());
  result.success = res.success != 0;
  result.fInit = res.f_init;
  result.gradNormInit = res.gradnorm_init;
  result.fOpt = res.f_opt;
  result.gradNormOpt = res.gradnorm_opt;
  result.relativeChange = res.relative_change;
  result.elapsedMs = res.elapsed_ms;
  result.tCGStatus = (res.tcg_status >= 0) ? (ROPTLIB::tCGstatusSet)res.tcg_status : ROPTLIB::TR_MAXITER;
  result.tCGIterations = res.tcg_iterations;
  result.outerIterations = res.outer_iterations;
  result.rejections = res.rejections;
  result.spmvPasses = res.spmv_passes;
  if (verbose)
    printf("[dpgo_b200] f %.10g -> %.10g  |g| %.6g -> %.6g  tCG x%d  passes over Q %d  %.3f ms\n", result.fInit, result.fOpt,
           result.gradNormInit, result.gradNormOpt, result.tCGIterations, result.spmvPasses, result.elapsedMs);
  assert(result.fOpt <= result.fInit + 1e-9 * (1.0 + std::abs(result.fInit)));
  return YOpt;
}

}  // namespace DPGO
