// RelativeSEMeasurement.h -- one relative SE(d) measurement (r1,p1) -> (r2,p2); field-compatible with the
// reference's include/DPGO/RelativeSEMeasurement.h:21-50.
#ifndef DPGO_B200_RELATIVESEMEASUREMENT_H
#define DPGO_B200_RELATIVESEMEASUREMENT_H

#include <DPGO/DPGO_types.h>

namespace DPGO {

struct RelativeSEMeasurement {
  size_t r1 = 0, r2 = 0;   // robots
  size_t p1 = 0, p2 = 0;   // poses
  Matrix R;                // rotation d x d
  Matrix t;                // translation d x 1
  double kappa = 0;        // rotational precision
  double tau = 0;          // translational precision
  bool isKnownInlier = false;
  double weight = 1.0;     // GNC weight in (0,1)

  RelativeSEMeasurement() = default;
  RelativeSEMeasurement(size_t first_robot, size_t second_robot, size_t first_pose, size_t second_pose,
                        const Eigen::MatrixXd &relative_rotation, const Eigen::VectorXd &relative_translation,
                        double rotational_precision, double translational_precision)
      : r1(first_robot), r2(second_robot), p1(first_pose), p2(second_pose), R(relative_rotation),
        t(relative_translation), kappa(rotational_precision), tau(translational_precision) {}

  friend std::ostream &operator<<(std::ostream &os, const RelativeSEMeasurement &m) {
    os << "(" << m.r1 << "," << m.p1 << ") -> (" << m.r2 << "," << m.p2 << ")\nR:\n" << m.R << "\nt:\n" << m.t
       << "\nkappa " << m.kappa << " tau " << m.tau << " inlier " << m.isKnownInlier << " weight " << m.weight
       << std::endl;
    return os;
  }
};

}  // namespace DPGO
#endif
