// RelativeSEMeasurement.h -- one relative SE(d) measurement between pose p1 of robot r1 and pose p2 of robot r2;
// field-compatible with the reference's include/DPGO/RelativeSEMeasurement.h:21-50 (its drivers fill and read the
// public members directly: examples/MultiRobotExample.cpp:120-129, examples/SingleRobotExample.cpp:63-67).
//
// The measurement says  R2 ~= R1 R,  t2 ~= t1 + R1 t  in the frame of the first pose; kappa and tau are the isotropic
// Langevin / Gaussian precisions that enter the connection Laplacian (Omega = diag(kappa I_d, tau), scaled by weight).
#ifndef DPGO_B200_RELATIVESEMEASUREMENT_H
#define DPGO_B200_RELATIVESEMEASUREMENT_H

#include <DPGO/DPGO_types.h>

namespace DPGO {

struct RelativeSEMeasurement {
  // endpoints: (robot, pose index inside that robot's trajectory)
  size_t r1 = 0, r2 = 0;
  size_t p1 = 0, p2 = 0;
  // relative transform, d x d and d x 1
  Matrix R;
  Matrix t;
  // precisions
  double kappa = 0;
  double tau = 0;
  // robust estimation state: a known inlier keeps weight 1; otherwise GNC moves the weight inside [0, 1]
  bool isKnownInlier = false;
  double weight = 1.0;

  RelativeSEMeasurement() = default;

  // argument order of the reference's 8-argument constructor
  RelativeSEMeasurement(size_t first_robot, size_t second_robot, size_t first_pose, size_t second_pose,
                        const Eigen::MatrixXd &relative_rotation, const Eigen::VectorXd &relative_translation,
                        double rotational_precision, double translational_precision)
      : r1(first_robot), r2(second_robot), p1(first_pose), p2(second_pose), R(relative_rotation),
        t(relative_translation), kappa(rotational_precision), tau(translational_precision) {}

  // true when both endpoints belong to the same robot (odometry or a private loop closure)
  bool isPrivate() const { return r1 == r2; }

  friend std::ostream &operator<<(std::ostream &out, const RelativeSEMeasurement &m) {
    out << "edge (" << m.r1 << "," << m.p1 << ") -> (" << m.r2 << "," << m.p2 << ")  kappa=" << m.kappa
        << " tau=" << m.tau << " weight=" << m.weight << (m.isKnownInlier ? " [known inlier]" : "") << "\n";
    out << "R =\n" << m.R << "\nt =\n" << m.t << std::endl;
    return out;
  }
};

}  // namespace DPGO
#endif
