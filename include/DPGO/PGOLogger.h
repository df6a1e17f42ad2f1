// PGOLogger.h -- CSV dump / load of trajectories and measurements (off by default, not on the hot path).
// Interface of the reference's include/DPGO/PGOLogger.h:17-63.
#ifndef DPGO_B200_PGOLOGGER_H
#define DPGO_B200_PGOLOGGER_H

#include <DPGO/DPGO_types.h>
#include <DPGO/RelativeSEMeasurement.h>

namespace DPGO {

class PGOLogger {
 public:
  explicit PGOLogger(std::string logDir) : logDirectory(std::move(logDir)) {}
  void logMeasurements(std::vector<RelativeSEMeasurement> &measurements, const std::string &filename);
  void logTrajectory(unsigned d, unsigned n, const Matrix &T, const std::string &filename);
  Matrix loadTrajectory(const std::string &filename);
  std::vector<RelativeSEMeasurement> loadMeasurements(const std::string &filename);

 private:
  std::string logDirectory;
};

}  // namespace DPGO
#endif
