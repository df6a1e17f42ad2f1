// PGOLogger.cpp -- CSV dump / load (off by default; not on the hot path).  Formats follow the reference's
// src/PGOLogger.cpp:18-225: trajectories as "pose_index,qx,qy,qz,qw,tx,ty,tz", measurements as
// "robot_src,pose_src,robot_dst,pose_dst,qx,qy,qz,qw,tx,ty,tz,kappa,tau,is_known_inlier,weight".
#include <DPGO/PGOLogger.h>

#include <cmath>
#include <fstream>
#include <sstream>

namespace DPGO {

namespace {
void rotToQuat(const Matrix &R, double q[4]) {   // q = (x, y, z, w)
  const double tr = R(0, 0) + R(1, 1) + R(2, 2);
  if (tr > 0) {
    const double s = std::sqrt(tr + 1.0) * 2;
    q[3] = 0.25 * s; q[0] = (R(2, 1) - R(1, 2)) / s; q[1] = (R(0, 2) - R(2, 0)) / s; q[2] = (R(1, 0) - R(0, 1)) / s;
  } else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) {
    const double s = std::sqrt(1.0 + R(0, 0) - R(1, 1) - R(2, 2)) * 2;
    q[3] = (R(2, 1) - R(1, 2)) / s; q[0] = 0.25 * s; q[1] = (R(0, 1) + R(1, 0)) / s; q[2] = (R(0, 2) + R(2, 0)) / s;
  } else if (R(1, 1) > R(2, 2)) {
    const double s = std::sqrt(1.0 + R(1, 1) - R(0, 0) - R(2, 2)) * 2;
    q[3] = (R(0, 2) - R(2, 0)) / s; q[0] = (R(0, 1) + R(1, 0)) / s; q[1] = 0.25 * s; q[2] = (R(1, 2) + R(2, 1)) / s;
  } else {
    const double s = std::sqrt(1.0 + R(2, 2) - R(0, 0) - R(1, 1)) * 2;
    q[3] = (R(1, 0) - R(0, 1)) / s; q[0] = (R(0, 2) + R(2, 0)) / s; q[1] = (R(1, 2) + R(2, 1)) / s; q[2] = 0.25 * s;
  }
}
Matrix quatToRot(double x, double y, double z, double w) {
  const double nrm = std::sqrt(x * x + y * y + z * z + w * w);
  x /= nrm; y /= nrm; z /= nrm; w /= nrm;
  Matrix R(3, 3);
  R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - w * z); R(0, 2) = 2 * (x * z + w * y);
  R(1, 0) = 2 * (x * y + w * z); R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - w * x);
  R(2, 0) = 2 * (x * z - w * y); R(2, 1) = 2 * (y * z + w * x); R(2, 2) = 1 - 2 * (x * x + y * y);
  return R;
}
}  // namespace

void PGOLogger::logMeasurements(std::vector<RelativeSEMeasurement> &measurements, const std::string &filename) {
  if (measurements.empty() || measurements[0].R.rows() != 3) return;   // 3-D only, as the reference
  std::ofstream file(logDirectory + filename);
  if (!file.is_open()) return;
  file << "robot_src,pose_src,robot_dst,pose_dst,qx,qy,qz,qw,tx,ty,tz,kappa,tau,is_known_inlier,weight\n";
  file << std::setprecision(17);
  for (const auto &m : measurements) {
    double q[4];
    rotToQuat(m.R, q);
    file << m.r1 << "," << m.p1 << "," << m.r2 << "," << m.p2 << "," << q[0] << "," << q[1] << "," << q[2] << "," << q[3]
         << "," << m.t(0) << "," << m.t(1) << "," << m.t(2) << "," << m.kappa << "," << m.tau << "," << m.isKnownInlier
         << "," << m.weight << "\n";
  }
}

void PGOLogger::logTrajectory(unsigned d, unsigned n, const Matrix &T, const std::string &filename) {
  if (d != 3) return;
  std::ofstream file(logDirectory + filename);
  if (!file.is_open()) return;
  file << "pose_index,qx,qy,qz,qw,tx,ty,tz\n" << std::setprecision(17);
  for (unsigned i = 0; i < n; ++i) {
    double q[4];
    rotToQuat(T.block(0, i * 4, 3, 3), q);
    file << i << "," << q[0] << "," << q[1] << "," << q[2] << "," << q[3] << "," << T(0, i * 4 + 3) << ","
         << T(1, i * 4 + 3) << "," << T(2, i * 4 + 3) << "\n";
  }
}

Matrix PGOLogger::loadTrajectory(const std::string &filename) {
  std::ifstream in(logDirectory + filename);
  std::string line;
  std::vector<std::vector<double>> rows;
  std::getline(in, line);
  while (std::getline(in, line)) {
    std::stringstream ss(line);
    std::string tok;
    std::vector<double> v;
    while (std::getline(ss, tok, ',')) v.push_back(std::stod(tok));
    if (v.size() == 8) rows.push_back(v);
  }
  Matrix T(3, 4 * (Eigen::Index)rows.size());
  for (size_t i = 0; i < rows.size(); ++i) {
    T.block(0, 4 * i, 3, 3) = quatToRot(rows[i][1], rows[i][2], rows[i][3], rows[i][4]);
    for (int c = 0; c < 3; ++c) T(c, 4 * i + 3) = rows[i][5 + c];
  }
  return T;
}

std::vector<RelativeSEMeasurement> PGOLogger::loadMeasurements(const std::string &filename) {
  std::vector<RelativeSEMeasurement> out;
  std::ifstream in(logDirectory + filename);
  std::string line;
  std::getline(in, line);
  while (std::getline(in, line)) {
    std::stringstream ss(line);
    std::string tok;
    std::vector<double> v;
    while (std::getline(ss, tok, ',')) v.push_back(std::stod(tok));
    if (v.size() != 15) continue;
    Matrix t(3, 1);
    t(0) = v[8]; t(1) = v[9]; t(2) = v[10];
    RelativeSEMeasurement m((size_t)v[0], (size_t)v[2], (size_t)v[1], (size_t)v[3], quatToRot(v[4], v[5], v[6], v[7]), t,
                            v[11], v[12]);
    m.isKnownInlier = v[13] != 0;
    m.weight = v[14];
    out.push_back(m);
  }
  return out;
}

}  // namespace DPGO
