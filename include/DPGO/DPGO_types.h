// DPGO_types.h -- boundary types of the B200 host library, source-compatible with the reference's
// include/DPGO/DPGO_types.h:20-68 (Matrix, SparseMatrix, ROPTALG, ROPTResult, PoseID, PoseDict).
#ifndef DPGO_B200_TYPES_H
#define DPGO_B200_TYPES_H

#include <Eigen/Core>
#include <Eigen/SparseCore>

#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

// The reference records the truncated-CG exit reason with ROPTLIB's enum (DPGO_types.h:58).  ROPTLIB
// is not a dependency here; the enumerators are kept so that client code reading tCGStatus compiles.
namespace ROPTLIB {
enum tCGstatusSet { TR_NEGCURVTURE = 0, TR_EXCREGION, TR_LCON, TR_SCON, TR_MAXITER, TCGSTATUSSETLENGTH };
}

namespace DPGO {

typedef Eigen::VectorXd Vector;
typedef Eigen::MatrixXd Matrix;
typedef Eigen::DiagonalMatrix<double, Eigen::Dynamic> DiagonalMatrix;
typedef Eigen::SparseMatrix<double, Eigen::RowMajor> SparseMatrix;

// local solver run by an agent on its block
enum ROPTALG { RTR, RGD };

// B200 extension: which operator truncated CG is preconditioned with (see include/dpgo_b200.h)
// SparseExact = the reference's operator (Q + 0.1 I)^-1 through a nested-dissection block factorisation (default);
// DenseExact = the same operator through a dense inverse (A/B runs)
enum class Preconditioner { None = 0, BlockJacobi = 1, DenseExact = 2, SparseExact = 3 };

// statistics of one QuadraticOptimizer::optimize() call
struct ROPTResult {
  ROPTResult(bool suc = false, double f0 = 0, double gn0 = 0, double fStar = 0, double gnStar = 0,
             double relchange = 0, double ms = 0)
      : success(suc), fInit(f0), gradNormInit(gn0), fOpt(fStar), gradNormOpt(gnStar), relativeChange(relchange),
        elapsedMs(ms), tCGStatus(ROPTLIB::TR_MAXITER) {}
  bool success;
  double fInit, gradNormInit, fOpt, gradNormOpt, relativeChange, elapsedMs;
  ROPTLIB::tCGstatusSet tCGStatus;
  // B200 extras (not in the reference)
  int tCGIterations = 0, outerIterations = 0, rejections = 0, spmvPasses = 0;
};

typedef std::pair<unsigned, unsigned> PoseID;   // (robot, pose)
typedef std::map<PoseID, Matrix, std::less<>, Eigen::aligned_allocator<std::pair<const PoseID, Matrix>>> PoseDict;

}  // namespace DPGO
#endif
