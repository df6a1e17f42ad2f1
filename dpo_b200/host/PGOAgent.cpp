// PGOAgent.cpp -- host side of one agent of the distributed pose-graph optimiser.
//
// Keeps the reference's behaviour at the interface (src/PGOAgent.cpp, cited per method) while every numeric
// step of the per-iteration path runs on the GPU through QuadraticProblem / QuadraticOptimizer ->
// libdpgo_b200.so.  Host work per selected iteration: assemble G from the cached neighbour poses (a few
// hundred small products) and hand X to one persistent kernel.
#include <DPGO/DPGO_utils.h>
#include <DPGO/PGOAgent.h>
#include <DPGO/QuadraticOptimizer.h>

#include <unistd.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <random>

namespace DPGO {

namespace {
// homogeneous (d+1)x(d+1) form of a measurement and its diagonal weights
Matrix homogeneous(const RelativeSEMeasurement &m, unsigned d) {
  Matrix T = Matrix::Zero(d + 1, d + 1);
  T.block(0, 0, d, d) = m.R;
  T.block(0, d, d, 1) = m.t;
  T(d, d) = 1;
  return T;
}
Matrix weights(const RelativeSEMeasurement &m, unsigned d) {
  Matrix Om = Matrix::Zero(d + 1, d + 1);
  for (unsigned k = 0; k < d; ++k) Om(k, k) = m.weight * m.kappa;
  Om(d, d) = m.weight * m.tau;
  return Om;
}
std::vector<RelativeSEMeasurement> concat(const std::vector<RelativeSEMeasurement> &a, const std::vector<RelativeSEMeasurement> &b) {
  std::vector<RelativeSEMeasurement> out(a);
  out.insert(out.end(), b.begin(), b.end());
  return out;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------
// construction / iterate access (ref :30-123)
// ---------------------------------------------------------------------------------------------------
PGOAgent::PGOAgent(unsigned ID, const PGOAgentParameters &params)
    : mID(ID), d(params.d), r(params.r), n(1), mParams(params), mState(PGOAgentState::WAIT_FOR_DATA),
      mStatus(ID, mState, 0, 0, false, 0), mRobustCost(params.robustCostType, params.robustCostParams),
      mProblemPtr(nullptr), mInstanceNumber(0), mIterationNumber(0), mNumPosesReceived(0), mLogger(params.logDirectory) {
  if (mParams.verbose) std::cout << "Initializing PGO agent...\n" << params << std::endl;
  X = Matrix::Zero(r, d + 1);
  X.block(0, 0, d, d) = Matrix::Identity(d, d);
  if (mID == 0) setLiftingMatrix(fixedStiefelVariable(d, r));   // agent 0 generates the shared lifting matrix
  resetTeamStatus();
}

PGOAgent::~PGOAgent() {
  endOptimizationLoop();
  delete mProblemPtr;
}

void PGOAgent::setX(const Matrix &Xin) {
  lock_guard<mutex> lock(mPosesMutex);
  assert(mState != PGOAgentState::WAIT_FOR_DATA);
  assert(Xin.rows() == relaxation_rank() && Xin.cols() == (dimension() + 1) * num_poses());
  mState = PGOAgentState::INITIALIZED;
  X = Xin;
  if (mParams.acceleration) initializeAcceleration();
  if (mParams.verbose) printf("Robot %u resets trajectory estimates. New trajectory length = %u\n", getID(), num_poses());
}

bool PGOAgent::getX(Matrix &Mout) {
  lock_guard<mutex> lock(mPosesMutex);
  Mout = X;
  return true;
}

bool PGOAgent::getSharedPose(unsigned index, Matrix &Mout) {
  if (mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mPosesMutex);
  if (index >= num_poses()) return false;
  Mout = X.block(0, index * (d + 1), r, d + 1);
  return true;
}

bool PGOAgent::getAuxSharedPose(unsigned index, Matrix &Mout) {
  assert(mParams.acceleration);
  if (mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mPosesMutex);
  if (index >= num_poses()) return false;
  Mout = Y.block(0, index * (d + 1), r, d + 1);
  return true;
}

bool PGOAgent::getSharedPoseDict(PoseDict &map) {
  if (mState != PGOAgentState::INITIALIZED) return false;
  map.clear();
  lock_guard<mutex> lock(mPosesMutex);
  for (const PoseID &pid : localSharedPoseIDs) map[pid] = X.block(0, pid.second * (d + 1), r, d + 1);
  return true;
}

bool PGOAgent::getAuxSharedPoseDict(PoseDict &map) {
  assert(mParams.acceleration);
  if (mState != PGOAgentState::INITIALIZED) return false;
  map.clear();
  lock_guard<mutex> lock(mPosesMutex);
  for (const PoseID &pid : localSharedPoseIDs) map[pid] = Y.block(0, pid.second * (d + 1), r, d + 1);
  return true;
}

void PGOAgent::setLiftingMatrix(const Matrix &M) {
  assert(M.rows() == r && M.cols() == d);
  YLift.emplace(M);
}

bool PGOAgent::getLiftingMatrix(Matrix &M) const {
  assert(mID == 0);
  if (!YLift) return false;
  M = YLift.value();
  return true;
}

void PGOAgent::setGlobalAnchor(const Matrix &M) {
  assert(M.rows() == relaxation_rank() && M.cols() == dimension() + 1);
  globalAnchor.emplace(M);
}

// ---------------------------------------------------------------------------------------------------
// pose graph (ref :126-248)
// ---------------------------------------------------------------------------------------------------
void PGOAgent::setPoseGraph(const std::vector<RelativeSEMeasurement> &inputOdometry,
                            const std::vector<RelativeSEMeasurement> &inputPrivateLoopClosures,
                            const std::vector<RelativeSEMeasurement> &inputSharedLoopClosures, const Matrix &TInit) {
  assert(!isOptimizationRunning());
  assert(mState == PGOAgentState::WAIT_FOR_DATA);
  assert(n == 1);
  if (inputOdometry.empty()) return;
  for (const auto &e : inputOdometry) addOdometry(e);
  for (const auto &e : inputPrivateLoopClosures) addPrivateLoopClosure(e);
  for (const auto &e : inputSharedLoopClosures) addSharedLoopClosure(e);

  const unsigned rows = dimension(), cols = (dimension() + 1) * num_poses();
  bool local_init = true;
  if (TInit.rows() > 0 && TInit.cols() > 0) {
    if ((unsigned)TInit.rows() == rows && (unsigned)TInit.cols() == cols) {
      local_init = false;
    } else {
      printf("Error: provided initial trajectory has wrong dimension! Expect (%u,%u), received (%ld, %ld). "
             "Using local initialization. \n", rows, cols, (long)TInit.rows(), (long)TInit.cols());
    }
  }
  delete mProblemPtr;
  mProblemPtr = new QuadraticProblem(num_poses(), dimension(), relaxation_rank());
  if (mParams.device >= 0) mProblemPtr->setDevice(mParams.device);
  if (mParams.cluster) mProblemPtr->setClusterLaunch(true);
  mProblemPtr->setPreconditioners(true, mParams.preconditioner == Preconditioner::DenseExact || mParams.preconditioner == Preconditioner::SparseExact,
                                  mParams.preconditioner);
  constructQMatrix();               // Q does not depend on the neighbours
  if (!local_init) {
    if (mParams.verbose) printf("Using provided trajectory initialization.\n");
    TLocalInit.emplace(TInit);
  } else {
    if (mParams.verbose) printf("Using internal trajectory initialization.\n");
    localInitialization();
  }
  mState = PGOAgentState::WAIT_FOR_INITIALIZATION;
  // agent 0 (or any agent when cross-robot initialisation is off) defines the global frame
  if (mID == 0 || !mParams.multirobot_initialization) {
    X = YLift.value() * TLocalInit.value();
    XInit.emplace(X);
    mState = PGOAgentState::INITIALIZED;
    if (mParams.acceleration) initializeAcceleration();
    if (mParams.logData) mLogger.logTrajectory(dimension(), num_poses(), TLocalInit.value(), "trajectory_initial.csv");
  }
}

void PGOAgent::addOdometry(const RelativeSEMeasurement &factor) {
  assert(mState != PGOAgentState::INITIALIZED);
  assert(factor.r1 == mID && factor.r2 == mID && factor.p1 + 1 == factor.p2);
  assert(factor.R.rows() == d && factor.R.cols() == d && factor.t.rows() == d && factor.t.cols() == 1);
  n = std::max(n, (unsigned)factor.p2 + 1);
  lock_guard<mutex> lock(mMeasurementsMutex);
  odometry.push_back(factor);
}

void PGOAgent::addPrivateLoopClosure(const RelativeSEMeasurement &factor) {
  assert(mState != PGOAgentState::INITIALIZED);
  assert(factor.r1 == mID && factor.r2 == mID);
  n = std::max(n, (unsigned)std::max(factor.p1 + 1, factor.p2 + 1));
  lock_guard<mutex> lock(mMeasurementsMutex);
  privateLoopClosures.push_back(factor);
}

void PGOAgent::addSharedLoopClosure(const RelativeSEMeasurement &factor) {
  assert(mState != PGOAgentState::INITIALIZED);
  const bool outgoing = (factor.r1 == mID);
  assert(outgoing ? factor.r2 != mID : factor.r2 == mID);
  const size_t mine = outgoing ? factor.p1 : factor.p2;
  const size_t otherRobot = outgoing ? factor.r2 : factor.r1, otherPose = outgoing ? factor.p2 : factor.p1;
  n = std::max(n, (unsigned)mine + 1);
  localSharedPoseIDs.insert(std::make_pair(mID, (unsigned)mine));
  neighborSharedPoseIDs.insert(std::make_pair((unsigned)otherRobot, (unsigned)otherPose));
  neighborRobotIDs.insert((unsigned)otherRobot);
  lock_guard<mutex> lock(mMeasurementsMutex);
  sharedLoopClosures.push_back(factor);
}

// ---------------------------------------------------------------------------------------------------
// cross-robot frame alignment (ref :250-432)
// ---------------------------------------------------------------------------------------------------
Matrix PGOAgent::computeNeighborTransform(const PoseID &nID, const Matrix &var) {
  assert(YLift);
  assert(var.rows() == r && var.cols() == d + 1);
  RelativeSEMeasurement &m = findSharedLoopClosureWithNeighbor(nID);
  // world1: my frame before alignment, world2: the neighbour's (global) frame;
  // frame1: my public pose, frame2: the neighbour's public pose
  Matrix dT = homogeneous(m, d);
  Matrix T_world2_frame2 = Matrix::Identity(d + 1, d + 1);
  T_world2_frame2.block(0, 0, d, d + 1) = YLift.value().transpose() * var;   // round the neighbour pose back to SE(d)
  const Matrix &T = TLocalInit.value();
  Matrix T_frame1_frame2, T_world1_frame1 = Matrix::Identity(d + 1, d + 1);
  if (m.r1 == nID.first) {          // incoming edge
    T_frame1_frame2 = dT.inverse();
    T_world1_frame1.block(0, 0, d, d + 1) = T.block(0, m.p2 * (d + 1), d, d + 1);
  } else {                          // outgoing edge
    T_frame1_frame2 = dT;
    T_world1_frame1.block(0, 0, d, d + 1) = T.block(0, m.p1 * (d + 1), d, d + 1);
  }
  Matrix T_world2_frame1 = T_world2_frame2 * T_frame1_frame2.inverse();
  Matrix T_world2_world1 = T_world2_frame1 * T_world1_frame1.inverse();
  checkRotationMatrix(T_world2_world1.block(0, 0, d, d));
  return T_world2_world1;
}

Matrix PGOAgent::computeRobustNeighborTransformTwoStage(unsigned neighborID, const PoseDict &poseDict) {
  std::vector<Matrix> RVec;
  std::vector<Vector> tVec;
  for (const auto &kv : poseDict)
    if (neighborSharedPoseIDs.count(kv.first)) {
      const Matrix T = computeNeighborTransform(kv.first, kv.second);
      RVec.emplace_back(T.block(0, 0, d, d));
      tVec.emplace_back(Vector(T.block(0, d, d, 1)));
    }
  const int m = (int)RVec.size();
  Matrix ROpt;
  Vector tOpt;
  std::vector<size_t> inliers;
  robustSingleRotationAveraging(ROpt, inliers, RVec, Vector::Ones(m), angular2ChordalSO3(0.5));   // ~30 degrees
  printf("[RobustRelativeTransform] This robot %u, neighbor %u: finds %i inliers out of %i measurements.\n", getID(),
         neighborID, (int)inliers.size(), m);
  if (inliers.empty()) throw std::runtime_error("Robust single rotation averaging returns empty inlier set!");
  std::vector<Vector> tIn;
  for (size_t idx : inliers) tIn.push_back(tVec[idx]);
  singleTranslationAveraging(tOpt, tIn);
  Matrix TOpt = Matrix::Identity(d + 1, d + 1);
  TOpt.block(0, 0, d, d) = ROpt;
  TOpt.block(0, d, d, 1) = tOpt;
  return TOpt;
}

Matrix PGOAgent::computeRobustNeighborTransform(unsigned neighborID, const PoseDict &poseDict) {
  std::vector<Matrix> RVec;
  std::vector<Vector> tVec;
  for (const auto &kv : poseDict)
    if (neighborSharedPoseIDs.count(kv.first)) {
      const Matrix T = computeNeighborTransform(kv.first, kv.second);
      RVec.emplace_back(T.block(0, 0, d, d));
      tVec.emplace_back(Vector(T.block(0, d, d, 1)));
    }
  const int m = (int)RVec.size();
  const Vector kappa = Vector::Constant(m, 1.82);     // rotation stddev ~30 degrees
  const Vector tau = Vector::Constant(m, 0.01);       // translation stddev 10 m
  const double cbar = RobustCost::computeErrorThresholdAtQuantile(0.9, 3);
  Matrix ROpt;
  Vector tOpt;
  std::vector<size_t> inliers;
  robustSinglePoseAveraging(ROpt, tOpt, inliers, RVec, tVec, kappa, tau, cbar);
  printf("[RobustRelativeTransform] This robot %u, neighbor %u: finds %i inliers out of %i measurements.\n", getID(),
         neighborID, (int)inliers.size(), m);
  if (inliers.empty()) throw std::runtime_error("Robust single pose averaging returns empty inlier set!");
  Matrix TOpt = Matrix::Identity(d + 1, d + 1);
  TOpt.block(0, 0, d, d) = ROpt;
  TOpt.block(0, d, d, 1) = tOpt;
  return TOpt;
}

void PGOAgent::initializeInGlobalFrame(unsigned neighborID, const PoseDict &poseDict) {
  assert(YLift);
  bool halted = false;
  if (isOptimizationRunning()) {
    if (mParams.verbose) printf("Robot %u halting optimization thread...\n", getID());
    halted = true;
    endOptimizationLoop();
  }
  lock_guard<mutex> tLock(mPosesMutex);
  lock_guard<mutex> mLock(mMeasurementsMutex);
  lock_guard<mutex> nLock(mNeighborPosesMutex);
  neighborPoseDict.clear();
  neighborAuxPoseDict.clear();
  Matrix T_world2_world1;
  try {
    T_world2_world1 = computeRobustNeighborTransformTwoStage(neighborID, poseDict);
  } catch (const std::runtime_error &e) {
    printf("Robust initialization is not successful! Abort and wait to try again...\n");
    return;
  }
  Matrix T = TLocalInit.value();
  Matrix Tw1 = Matrix::Identity(d + 1, d + 1);
  for (size_t i = 0; i < num_poses(); ++i) {
    Tw1.block(0, 0, d, d + 1) = T.block(0, i * (d + 1), d, d + 1);
    Matrix Tw2 = T_world2_world1 * Tw1;
    T.block(0, i * (d + 1), d, d + 1) = Tw2.block(0, 0, d, d + 1);
  }
  X = YLift.value() * T;
  XInit.emplace(X);
  mState = PGOAgentState::INITIALIZED;
  if (mParams.acceleration) initializeAcceleration();
  if (mParams.logData) mLogger.logTrajectory(dimension(), num_poses(), T, "trajectory_initial.csv");
  if (halted) startOptimizationLoop(mRate);
}

// ---------------------------------------------------------------------------------------------------
// neighbour caches (ref :434-479)
// ---------------------------------------------------------------------------------------------------
void PGOAgent::updateNeighborPoses(unsigned neighborID, const PoseDict &poseDict) {
  assert(neighborID != mID);
  const auto neighborState = getNeighborStatus(neighborID).state;
  if (mState == PGOAgentState::WAIT_FOR_INITIALIZATION && neighborState == PGOAgentState::INITIALIZED)
    initializeInGlobalFrame(neighborID, poseDict);
  for (const auto &kv : poseDict) {
    assert(kv.first.first == neighborID && kv.second.rows() == r && kv.second.cols() == d + 1);
    mNumPosesReceived++;
    if (!neighborSharedPoseIDs.count(kv.first)) continue;
    if (mState == PGOAgentState::INITIALIZED && neighborState == PGOAgentState::INITIALIZED) {
      lock_guard<mutex> lock(mNeighborPosesMutex);
      neighborPoseDict[kv.first] = kv.second;
    }
  }
}

void PGOAgent::updateAuxNeighborPoses(unsigned neighborID, const PoseDict &poseDict) {
  assert(mParams.acceleration && neighborID != mID);
  for (const auto &kv : poseDict) {
    assert(kv.first.first == neighborID && kv.second.rows() == r && kv.second.cols() == d + 1);
    mNumPosesReceived++;
    if (!neighborSharedPoseIDs.count(kv.first)) continue;
    if (mState == PGOAgentState::INITIALIZED && getNeighborStatus(neighborID).state == PGOAgentState::INITIALIZED) {
      lock_guard<mutex> lock(mNeighborPosesMutex);
      neighborAuxPoseDict[kv.first] = kv.second;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// rounding (ref :481-581)
// ---------------------------------------------------------------------------------------------------
bool PGOAgent::getTrajectoryInLocalFrame(Matrix &Trajectory) {
  if (mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mPosesMutex);
  Matrix T = Matrix(X.block(0, 0, r, d)).transpose() * X;     // anchor on my first pose
  Matrix t0 = T.block(0, d, d, 1);
  for (unsigned i = 0; i < n; ++i) {
    T.block(0, i * (d + 1), d, d) = projectToRotationGroup(T.block(0, i * (d + 1), d, d));
    T.block(0, i * (d + 1) + d, d, 1) = T.block(0, i * (d + 1) + d, d, 1) - t0;
  }
  Trajectory = T;
  return true;
}

bool PGOAgent::getTrajectoryInGlobalFrame(Matrix &Trajectory) {
  if (!globalAnchor || mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mPosesMutex);
  const Matrix Ya = globalAnchor.value().block(0, 0, r, d);
  Matrix T = Ya.transpose() * X;
  Matrix t0 = Ya.transpose() * Matrix(globalAnchor.value().block(0, d, r, 1));
  for (unsigned i = 0; i < n; ++i) {
    T.block(0, i * (d + 1), d, d) = projectToRotationGroup(T.block(0, i * (d + 1), d, d));
    T.block(0, i * (d + 1) + d, d, 1) = T.block(0, i * (d + 1) + d, d, 1) - t0;
  }
  Trajectory = T;
  return true;
}

bool PGOAgent::getPoseInGlobalFrame(unsigned poseID, Matrix &T) {
  if (!globalAnchor || mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mPosesMutex);
  if (poseID >= num_poses()) return false;
  const Matrix Ya = globalAnchor.value().block(0, 0, r, d);
  Matrix t0 = Ya.transpose() * Matrix(globalAnchor.value().block(0, d, r, 1));
  Matrix Ti = Ya.transpose() * Matrix(X.block(0, poseID * (d + 1), r, d + 1));
  Ti.block(0, d, d, 1) -= t0;
  T = Ti;
  return true;
}

bool PGOAgent::getNeighborPoseInGlobalFrame(unsigned neighborID, unsigned poseID, Matrix &T) {
  if (!globalAnchor || mState != PGOAgentState::INITIALIZED) return false;
  lock_guard<mutex> lock(mNeighborPosesMutex);
  auto it = neighborPoseDict.find(std::make_pair(neighborID, poseID));
  if (it == neighborPoseDict.end()) return false;
  const Matrix Ya = globalAnchor.value().block(0, 0, r, d);
  Matrix t0 = Ya.transpose() * Matrix(globalAnchor.value().block(0, d, r, 1));
  Matrix Ti = Ya.transpose() * it->second;
  Ti.block(0, d, d, 1) -= t0;
  T = Ti;
  return true;
}

std::vector<unsigned> PGOAgent::getNeighborPublicPoses(const unsigned &neighborID) const {
  assert(neighborRobotIDs.count(neighborID));
  std::vector<unsigned> out;
  for (const PoseID &pid : neighborSharedPoseIDs)
    if (pid.first == neighborID) out.push_back(pid.second);
  return out;
}

std::vector<unsigned> PGOAgent::getNeighbors() const { return std::vector<unsigned>(neighborRobotIDs.begin(), neighborRobotIDs.end()); }

// ---------------------------------------------------------------------------------------------------
// reset (ref :583-640)
// ---------------------------------------------------------------------------------------------------
void PGOAgent::reset() {
  endOptimizationLoop();
  if (mParams.logData) {
    std::vector<RelativeSEMeasurement> all = concat(concat(odometry, privateLoopClosures), sharedLoopClosures);
    mLogger.logMeasurements(all, "measurements.csv");
    Matrix T;
    if (getTrajectoryInGlobalFrame(T)) {
      mLogger.logTrajectory(dimension(), num_poses(), T, "trajectory_optimized.csv");
      std::cout << "Saved optimized trajectory to " << mParams.logDirectory << std::endl;
    }
    writeMatrixToFile(X, mParams.logDirectory + "X.txt");
  }
  mInstanceNumber++;
  mIterationNumber = 0;
  mNumPosesReceived = 0;
  mState = PGOAgentState::WAIT_FOR_DATA;       // the lifting matrix is kept
  mStatus = PGOAgentStatus(getID(), mState, mInstanceNumber, mIterationNumber, false, 0);
  odometry.clear();
  privateLoopClosures.clear();
  sharedLoopClosures.clear();
  neighborPoseDict.clear();
  neighborAuxPoseDict.clear();
  localSharedPoseIDs.clear();
  neighborSharedPoseIDs.clear();
  neighborRobotIDs.clear();
  resetTeamStatus();
  delete mProblemPtr;
  mProblemPtr = nullptr;
  mRobustCost.reset();
  globalAnchor.reset();
  TLocalInit.reset();
  XInit.reset();
  mOptimizationRequested = mPublishPublicPosesRequested = mPublishWeightsRequested = false;
  n = 1;
  X = Matrix::Zero(r, d + 1);
  X.block(0, 0, d, d) = Matrix::Identity(d, d);
}

// ---------------------------------------------------------------------------------------------------
// one RBCD iteration (ref :642-718)
// ---------------------------------------------------------------------------------------------------
void PGOAgent::iterate(bool doOptimization) {
  mIterationNumber++;
  if (mIterationNumber == 50 && mParams.logData) {
    Matrix T;
    if (getTrajectoryInGlobalFrame(T)) mLogger.logTrajectory(dimension(), num_poses(), T, "trajectory_early_stop.csv");
  }
  if (shouldUpdateLoopClosureWeights()) {      // GNC: never for the L2 cost
    updateLoopClosuresWeights();
    mRobustCost.update();
    if (!mParams.robustOptWarmStart) {
      assert(XInit);
      X = XInit.value();
      printf("Warm start is disabled. Robot %u resets trajectory estimates.\n", getID());
    }
    if (mParams.acceleration) initializeAcceleration();
  }
  if (mState != PGOAgentState::INITIALIZED) return;
  XPrev = X;
  std::unique_lock<mutex> tLock(mPosesMutex);
  std::unique_lock<mutex> mLock(mMeasurementsMutex);
  std::unique_lock<mutex> nLock(mNeighborPosesMutex);
  bool success;
  if (mParams.acceleration) {
    updateGamma();
    updateAlpha();
    updateY();
    success = updateX(doOptimization, true);
    updateV();
    if (shouldRestart()) restartNesterovAcceleration(doOptimization);
    mPublishPublicPosesRequested = true;
  } else {
    success = updateX(doOptimization, false);
    if (doOptimization) mPublishPublicPosesRequested = true;
  }
  if (doOptimization) {
    mStatus.agentID = getID();
    mStatus.state = mState;
    mStatus.instanceNumber = instance_number();
    mStatus.iterationNumber = iteration_number();
    mStatus.relativeChange = std::sqrt((X - XPrev).squaredNorm() / num_poses());
    bool ready = success && mStatus.relativeChange <= mParams.relChangeTol;
    if (computeConvergedLoopClosureRatio() < mParams.robustOptMinConvergenceRatio) ready = false;
    mStatus.readyToTerminate = ready;
  }
}

// ---------------------------------------------------------------------------------------------------
// cost matrices (ref :720-859)
// ---------------------------------------------------------------------------------------------------
void PGOAgent::constructQMatrix() {
  // private edges contribute their full Laplacian; a shared edge contributes only the diagonal block of the
  // pose this agent owns: T Om T^T for an outgoing edge (at p1), Om for an incoming one (at p2)
  SparseMatrix Q = constructConnectionLaplacianSE(concat(odometry, privateLoopClosures));
  if ((unsigned)Q.rows() != (d + 1) * n) {       // private edges may not reach the last public pose
    SparseMatrix Qfull((d + 1) * n, (d + 1) * n);
    for (Eigen::Index k = 0; k < Q.outerSize(); ++k)
      for (SparseMatrix::InnerIterator it(Q, k); it; ++it) Qfull.coeffRef(it.row(), it.col()) = it.value();
    Q = Qfull;
  }
  for (const auto &m : sharedLoopClosures) {
    const Matrix T = homogeneous(m, d), Om = weights(m, d);
    const bool outgoing = (m.r1 == mID);
    const size_t idx = outgoing ? m.p1 : m.p2;
    const Matrix W = outgoing ? Matrix(T * Om * T.transpose()) : Om;
    for (size_t col = 0; col < d + 1; ++col)
      for (size_t row = 0; row < d + 1; ++row) Q.coeffRef(idx * (d + 1) + row, idx * (d + 1) + col) += W(row, col);
  }
  assert(mProblemPtr);
  mProblemPtr->setQ(Q);
}

bool PGOAgent::constructGMatrix(const PoseDict &poseDict) {
  // dense r x (d+1)n accumulation (the reference builds a sparse matrix with coeffRef; the values are the same):
  // outgoing edge: G_p1 += -X_j Om T^T ; incoming edge: G_p2 += -X_i T Om
  Matrix G = Matrix::Zero(r, (d + 1) * n);
  for (const auto &m : sharedLoopClosures) {
    const Matrix T = homogeneous(m, d), Om = weights(m, d);
    const bool outgoing = (m.r1 == mID);
    const PoseID nID = outgoing ? std::make_pair((unsigned)m.r2, (unsigned)m.p2) : std::make_pair((unsigned)m.r1, (unsigned)m.p1);
    auto it = poseDict.find(nID);
    if (it == poseDict.end()) {
      if (mParams.verbose) printf("constructGMatrix: robot %u cannot find neighbor pose (%u, %u)\n", getID(), nID.first, nID.second);
      return false;
    }
    const size_t idx = outgoing ? m.p1 : m.p2;
    const Matrix L = outgoing ? Matrix(it->second * Om * T.transpose()) : Matrix(it->second * T * Om);
    G.block(0, idx * (d + 1), r, d + 1) -= L;
  }
  assert(mProblemPtr);
  mProblemPtr->setG(G);
  return true;
}

// ---------------------------------------------------------------------------------------------------
// asynchronous mode (ref :861-925)
// ---------------------------------------------------------------------------------------------------
void PGOAgent::startOptimizationLoop(double freq) {
  assert(!mParams.acceleration);        // asynchronous updates are restricted to non-accelerated mode
  if (isOptimizationRunning()) {
    if (mParams.verbose) printf("startOptimizationLoop: optimization thread already running! \n");
    return;
  }
  mRate = freq;
  mOptimizationThread = new thread(&PGOAgent::runOptimizationLoop, this);
}

void PGOAgent::runOptimizationLoop() {
  if (mParams.verbose) printf("Robot %u optimization thread running at %f Hz.\n", getID(), mRate);
  std::random_device rd;
  std::mt19937 rng(rd());
  std::exponential_distribution<double> waitTime(mRate);     // Poisson clock
  while (true) {
    usleep((useconds_t)(1e6 * waitTime(rng)));
    iterate(true);
    if (mEndLoopRequested) break;
  }
}

void PGOAgent::endOptimizationLoop() {
  if (!isOptimizationRunning()) return;
  mEndLoopRequested = true;
  mOptimizationThread->join();
  delete mOptimizationThread;
  mOptimizationThread = nullptr;
  mEndLoopRequested = false;
  if (mParams.verbose) printf("Robot %u optimization thread exited. \n", getID());
}

bool PGOAgent::isOptimizationRunning() { return mOptimizationThread != nullptr; }

RelativeSEMeasurement &PGOAgent::findSharedLoopClosureWithNeighbor(const PoseID &nID) {
  for (auto &m : sharedLoopClosures)
    if ((m.r1 == nID.first && m.p1 == nID.second) || (m.r2 == nID.first && m.p2 == nID.second)) return m;
  throw std::runtime_error("Cannot find shared loop closure with neighbor.");
}

RelativeSEMeasurement &PGOAgent::findSharedLoopClosure(const PoseID &srcID, const PoseID &dstID) {
  for (auto &m : sharedLoopClosures)
    if (m.r1 == srcID.first && m.p1 == srcID.second && m.r2 == dstID.first && m.p2 == dstID.second) return m;
  throw std::runtime_error("Cannot find specified shared loop closure.");
}

// ---------------------------------------------------------------------------------------------------
// initialisation + single-robot solve (ref :945-990)
// ---------------------------------------------------------------------------------------------------
void PGOAgent::localInitialization() {
  Matrix T0;
  if (mParams.robustCostType == RobustCostType::L2) T0 = chordalInitialization(dimension(), num_poses(), concat(odometry, privateLoopClosures));
  else T0 = odometryInitialization(dimension(), num_poses(), odometry);   // robust mode does not trust loop closures
  assert(T0.rows() == d && T0.cols() == (d + 1) * n);
  TLocalInit.emplace(T0);
}

Matrix PGOAgent::localPoseGraphOptimization() {
  if (!TLocalInit) localInitialization();
  SparseMatrix Q = constructConnectionLaplacianSE(concat(odometry, privateLoopClosures));
  QuadraticProblem problem(n, d, d);        // rank r = d
  problem.setPreconditioners(true, mParams.preconditioner == Preconditioner::DenseExact || mParams.preconditioner == Preconditioner::SparseExact,
                             mParams.preconditioner);
  problem.setQ(Q);
  QuadraticOptimizer optimizer(&problem);
  optimizer.setVerbose(mParams.verbose);
  optimizer.setTrustRegionInitialRadius(10);
  optimizer.setTrustRegionIterations(10);
  optimizer.setTrustRegionTolerance(1e-1);
  optimizer.setTrustRegionMaxInnerIterations(50);
  optimizer.setPreconditioner(mParams.preconditioner);
  Matrix Topt = optimizer.optimize(TLocalInit.value());
  mLastResult = optimizer.getOptResult();
  if (mParams.verbose) printf("Optimization time: %f sec.\n", mLastResult.elapsedMs / 1e3);
  return Topt;
}

// ---------------------------------------------------------------------------------------------------
// termination, Nesterov acceleration (ref :1007-1091)
// ---------------------------------------------------------------------------------------------------
bool PGOAgent::shouldTerminate() {
  if (iteration_number() > mParams.maxNumIters) {
    printf("Reached maximum iterations.\n");
    return true;
  }
  for (size_t robot = 0; robot < mParams.numRobots; ++robot) {
    assert(mTeamStatus[robot].agentID == robot);
    if (mTeamStatus[robot].state != PGOAgentState::INITIALIZED) return false;
  }
  for (size_t robot = 0; robot < mParams.numRobots; ++robot)
    if (!mTeamStatus[robot].readyToTerminate) return false;
  return true;
}

bool PGOAgent::shouldRestart() const { return mParams.acceleration && ((mIterationNumber + 1) % mParams.restartInterval == 0); }

void PGOAgent::restartNesterovAcceleration(bool doOptimization) {
  if (!mParams.acceleration || mState != PGOAgentState::INITIALIZED) return;
  if (mParams.verbose) printf("Robot %u restarts Nesteorv acceleration.\n", getID());
  X = XPrev;
  updateX(doOptimization, false);
  V = X;
  Y = X;
  gamma = 0;
  alpha = 0;
}

void PGOAgent::initializeAcceleration() {
  assert(mParams.acceleration);
  if (mState != PGOAgentState::INITIALIZED) return;
  XPrev = X;
  gamma = 0;
  alpha = 0;
  V = X;
  Y = X;
}

void PGOAgent::updateGamma() {
  assert(mParams.acceleration && mState == PGOAgentState::INITIALIZED);
  const double N = mParams.numRobots;
  gamma = (1 + std::sqrt(1 + 4 * N * N * gamma * gamma)) / (2 * N);
}

void PGOAgent::updateAlpha() {
  assert(mParams.acceleration && mState == PGOAgentState::INITIALIZED);
  alpha = 1 / (gamma * mParams.numRobots);
}

void PGOAgent::updateY() {
  assert(mParams.acceleration && mState == PGOAgentState::INITIALIZED);
  LiftedSEManifold manifold(relaxation_rank(), dimension(), num_poses());
  Y = manifold.project((1 - alpha) * X + alpha * V);          // per-pose Stiefel projection on the GPU
}

void PGOAgent::updateV() {
  assert(mParams.acceleration && mState == PGOAgentState::INITIALIZED);
  LiftedSEManifold manifold(relaxation_rank(), dimension(), num_poses());
  V = manifold.project(V + gamma * (X - Y));
}

// ---------------------------------------------------------------------------------------------------
// the local update (ref :1093-1165)
// ---------------------------------------------------------------------------------------------------
bool PGOAgent::updateX(bool doOptimization, bool acceleration) {
  if (!doOptimization) {
    if (acceleration) X = Y;
    return true;
  }
  if (mParams.verbose) printf("Robot %u optimize at iteration %u... \n", getID(), iteration_number());
  if (acceleration) assert(mParams.acceleration);
  assert(mState == PGOAgentState::INITIALIZED);
  if (mParams.robustCostType != RobustCostType::L2) constructQMatrix();      // weights changed
  const bool hasG = constructGMatrix(acceleration ? neighborAuxPoseDict : neighborPoseDict);
  if (!hasG) {
    if (mParams.verbose) printf("Robot %u could not construct G matrix. Skip update...\n", getID());
    return false;
  }
  QuadraticOptimizer optimizer(mProblemPtr);
  optimizer.setVerbose(mParams.verbose);
  optimizer.setAlgorithm(mParams.algorithm);
  optimizer.setTrustRegionTolerance(1e-2);          // force progress
  optimizer.setTrustRegionIterations(1);
  optimizer.setTrustRegionMaxInnerIterations(10);
  optimizer.setTrustRegionInitialRadius(100);
  optimizer.setPreconditioner(mParams.preconditioner);
  const Matrix &start = acceleration ? Y : X;
  assert(start.rows() == relaxation_rank() && start.cols() == (dimension() + 1) * num_poses());
  X = optimizer.optimize(start);
  mLastResult = optimizer.getOptResult();
  if (mParams.verbose)
    printf("df: %f, gn0: %f, gn1: %f, df/gn0: %f\n", mLastResult.fInit - mLastResult.fOpt, mLastResult.gradNormInit,
           mLastResult.gradNormOpt, (mLastResult.fInit - mLastResult.fOpt) / mLastResult.gradNormInit);
  return true;
}

void PGOAgent::resetTeamStatus() {
  mTeamStatus.clear();
  for (unsigned robot = 0; robot < mParams.numRobots; ++robot) mTeamStatus.emplace_back(robot);
}

// ---------------------------------------------------------------------------------------------------
// GNC weights (ref :1174-1289); host scalar loops, inactive for the L2 cost
// ---------------------------------------------------------------------------------------------------
bool PGOAgent::shouldUpdateLoopClosureWeights() const {
  if (mParams.robustCostType == RobustCostType::L2) return false;
  return ((mIterationNumber + 1) % mParams.robustOptInnerIters == 0);
}

void PGOAgent::updateLoopClosuresWeights() {
  assert(mState == PGOAgentState::INITIALIZED);
  auto poseOf = [&](const Matrix &M, size_t idx, Matrix &Yp, Matrix &pp) {
    Yp = M.block(0, idx * (d + 1), r, d);
    pp = M.block(0, idx * (d + 1) + d, r, 1);
  };
  for (auto &m : privateLoopClosures) {
    if (m.isKnownInlier) continue;
    Matrix Y1, p1, Y2, p2;
    poseOf(X, m.p1, Y1, p1);
    poseOf(X, m.p2, Y2, p2);
    m.weight = mRobustCost.weight(std::sqrt(computeMeasurementError(m, Y1, p1, Y2, p2)));
  }
  // an agent updates the weights of the edges it shares with higher-numbered agents only
  for (auto &m : sharedLoopClosures) {
    if (m.isKnownInlier) continue;
    const bool outgoing = (m.r1 == getID());
    const size_t other = outgoing ? m.r2 : m.r1;
    if (other < getID()) continue;
    auto it = neighborPoseDict.find(outgoing ? std::make_pair((unsigned)m.r2, (unsigned)m.p2) : std::make_pair((unsigned)m.r1, (unsigned)m.p1));
    if (it == neighborPoseDict.end()) {
      printf("Agent %u cannot update edge: (%zu, %zu) -> (%zu, %zu). \n", getID(), m.r1, m.p1, m.r2, m.p2);
      continue;
    }
    Matrix Y1, p1, Y2, p2;
    if (outgoing) {
      poseOf(X, m.p1, Y1, p1);
      Y2 = it->second.block(0, 0, r, d);
      p2 = it->second.block(0, d, r, 1);
    } else {
      poseOf(X, m.p2, Y2, p2);
      Y1 = it->second.block(0, 0, r, d);
      p1 = it->second.block(0, d, r, 1);
    }
    m.weight = mRobustCost.weight(std::sqrt(computeMeasurementError(m, Y1, p1, Y2, p2)));
  }
  mPublishWeightsRequested = true;
}

double PGOAgent::computeConvergedLoopClosureRatio() {
  if (mParams.robustCostType != RobustCostType::GNC_TLS) return 1.0;
  double total = 0, decided = 0;
  auto count = [&](const std::vector<RelativeSEMeasurement> &ms) {
    for (const auto &m : ms) {
      if (m.isKnownInlier) continue;
      if (m.weight == 1 || m.weight == 0) decided += 1;
      total += 1;
    }
  };
  count(privateLoopClosures);
  count(sharedLoopClosures);
  return decided / total;
}

bool PGOAgent::isDuplicateMeasurement(const RelativeSEMeasurement &m, const vector<RelativeSEMeasurement> &measurements) {
  for (const auto &m2 : measurements)
    if (m.r1 == m2.r1 && m.r2 == m2.r2 && m.p1 == m2.p1 && m.p2 == m2.p2) return true;
  return false;
}

}  // namespace DPGO
