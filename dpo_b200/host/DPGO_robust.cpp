// DPGO_robust.cpp -- robust cost weights and the GNC mu schedule (host scalar math).
// Behaviour of the reference's src/DPGO_robust.cpp:17-103.
#include <DPGO/DPGO_robust.h>
#include <DPGO/DPGO_utils.h>

#include <cmath>
#include <stdexcept>

namespace DPGO {

RobustCost::RobustCost(RobustCostType costType, const RobustCostParameters &params) : mCostType(costType), mParams(params) {
  reset();
}

double RobustCost::weight(double r) {
  if (mCostType == RobustCostType::L2) return 1.0;
  if (mCostType == RobustCostType::L1) return 1.0 / r;
  if (mCostType == RobustCostType::Huber) return r < mParams.HuberThreshold ? 1.0 : mParams.HuberThreshold / r;
  if (mCostType == RobustCostType::TLS) return r < mParams.TLSThreshold ? 1.0 : 0.0;
  if (mCostType == RobustCostType::GM) {
    const double a = 1.0 + r * r;
    return 1.0 / (a * a);
  }
  if (mCostType == RobustCostType::GNC_TLS) {
    // graduated non-convexity surrogate of truncated least squares, eq. (14) of the GNC paper
    const double rSq = r * r, barcSq = mParams.GNCBarc * mParams.GNCBarc;
    if (rSq >= (mu + 1) / mu * barcSq) return 0.0;
    if (rSq <= mu / (mu + 1) * barcSq) return 1.0;
    return std::sqrt(barcSq * mu * (mu + 1) / rSq) - mu;
  }
  throw std::runtime_error("weight function for selected cost function is not implemented !");
}

void RobustCost::reset() {
  if (mCostType == RobustCostType::GNC_TLS) {
    mu = mParams.GNCInitMu;
    mGNCIteration = 0;
  }
}

void RobustCost::update() {
  if (mCostType != RobustCostType::GNC_TLS) return;
  if (++mGNCIteration > mParams.GNCMaxNumIters) {
    printf("GNC: reached maximum iterations.");
    return;
  }
  mu *= mParams.GNCMuStep;
}

double RobustCost::computeErrorThresholdAtQuantile(double quantile, size_t dimension) {
  // residual threshold such that P(chi2_dof <= t^2) = quantile, dof = dimension of SE(d) tangent space
  const size_t dof = (dimension == 2) ? 3 : 6;
  return std::sqrt(chi2inv(quantile, dof));
}

}  // namespace DPGO
