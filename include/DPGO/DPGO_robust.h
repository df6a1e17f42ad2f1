// DPGO_robust.h -- robust-cost weight functions (host scalar math; off (L2) on the benchmark path).
// Interface-compatible with the reference's include/DPGO/DPGO_robust.h:20-124.
#ifndef DPGO_B200_ROBUST_H
#define DPGO_B200_ROBUST_H

#include <iostream>
#include <string>
#include <vector>

namespace DPGO {

enum RobustCostType { L2, L1, TLS, Huber, GM, GNC_TLS };
const std::vector<std::string> RobustCostNames{"L2", "L1", "TLS", "Huber", "GM", "GNC_TLS"};

struct RobustCostParameters {
  unsigned GNCMaxNumIters;   // GNC outer iterations
  double GNCBarc;            // TLS threshold
  double GNCMuStep;          // mu multiplier per update
  double GNCInitMu;          // initial mu
  double TLSThreshold;
  double HuberThreshold;
  RobustCostParameters(unsigned GNCMaxIters = 100, double GNCBarcIn = 5.0, double GNCMuStepIn = 1.4,
                       double GNCInitMuIn = 1e-4, double TLSThresholdIn = 10, double HuberThresholdIn = 3)
      : GNCMaxNumIters(GNCMaxIters), GNCBarc(GNCBarcIn), GNCMuStep(GNCMuStepIn), GNCInitMu(GNCInitMuIn),
        TLSThreshold(TLSThresholdIn), HuberThreshold(HuberThresholdIn) {}
  friend std::ostream &operator<<(std::ostream &os, const RobustCostParameters &p) {
    os << "Robust cost parameters: GNC iters " << p.GNCMaxNumIters << ", barc " << p.GNCBarc << ", mu step "
       << p.GNCMuStep << ", init mu " << p.GNCInitMu << ", TLS " << p.TLSThreshold << ", Huber " << p.HuberThreshold
       << std::endl;
    return os;
  }
};

class RobustCost {
 public:
  RobustCost(RobustCostType costType, const RobustCostParameters &params);
  double weight(double r);   // weight of a residual r
  void reset();
  void update();             // advance the GNC schedule
  static double computeErrorThresholdAtQuantile(double quantile, size_t dimension);

 private:
  RobustCostType mCostType;
  RobustCostParameters mParams;
  double mu = 0;
  unsigned mGNCIteration = 0;
};

}  // namespace DPGO
#endif
