// PGOAgent.h -- one robot of the distributed pose-graph optimiser, B200 host side.
//
// Public interface = the reference's include/DPGO/PGOAgent.h:209-490 (same method names, argument
// meaning and return conventions) so that examples/MultiRobotExample.cpp and
// examples/SingleRobotExample.cpp compile and link unchanged.  The bookkeeping (measurement stores,
// neighbour caches, state machine, Nesterov scalars, GNC weights) stays on the host as in the
// reference; every numeric step on the per-iteration path -- cost / gradient / Hessian-vector products,
// truncated-CG trust-region step, retraction, Stiefel projection -- runs in libdpgo_b200.so.
#ifndef DPGO_B200_PGOAGENT_H
#define DPGO_B200_PGOAGENT_H

#include <DPGO/DPGO_robust.h>
#include <DPGO/DPGO_types.h>
#include <DPGO/PGOLogger.h>
#include <DPGO/QuadraticProblem.h>
#include <DPGO/RelativeSEMeasurement.h>
#include <DPGO/manifold/LiftedSEManifold.h>

#include <unistd.h>

#include <mutex>
#include <optional>
#include <set>
#include <stdexcept>
#include <thread>
#include <utility>
#include <vector>

namespace DPGO {

using std::lock_guard;
using std::mutex;
using std::set;
using std::thread;
using std::vector;

// life cycle of an agent; each state only moves to the next one
enum PGOAgentState { WAIT_FOR_DATA, WAIT_FOR_INITIALIZATION, INITIALIZED };

struct PGOAgentParameters {
  unsigned d;                   // problem dimension
  unsigned r;                   // relaxation rank
  unsigned numRobots;
  ROPTALG algorithm;            // local solver
  bool multirobot_initialization;
  bool acceleration;            // Nesterov acceleration
  unsigned restartInterval;
  RobustCostType robustCostType;
  RobustCostParameters robustCostParams;
  bool robustOptWarmStart;
  unsigned robustOptInnerIters;
  double robustOptMinConvergenceRatio;
  unsigned maxNumIters;
  double relChangeTol;
  bool verbose;
  bool logData;
  std::string logDirectory;
  Preconditioner preconditioner = Preconditioner::SparseExact;   // B200 extension
  int device = -1;              // B200 extension: CUDA device of this agent's problem (-1: env DPGO_DEVICE or 0)
  bool cluster = false;         // B200 extension: step kernel as one thread-block cluster, so that several agents share a GPU

  PGOAgentParameters(unsigned dIn, unsigned rIn, unsigned numRobotsIn = 1, ROPTALG algorithmIn = ROPTALG::RTR,
                     bool accel = false, unsigned restartInt = 30, RobustCostType costType = RobustCostType::L2,
                     RobustCostParameters costParams = RobustCostParameters(), bool robust_opt_warm_start = true,
                     unsigned robust_opt_inner_iters = 30, double robust_opt_min_convergence_ratio = 0.8,
                     unsigned maxIters = 500, double changeTol = 5e-3, bool v = false, bool log = false,
                     std::string logDir = "")
      : d(dIn), r(rIn), numRobots(numRobotsIn), algorithm(algorithmIn), multirobot_initialization(true),
        acceleration(accel), restartInterval(restartInt), robustCostType(costType), robustCostParams(costParams),
        robustOptWarmStart(robust_opt_warm_start), robustOptInnerIters(robust_opt_inner_iters),
        robustOptMinConvergenceRatio(robust_opt_min_convergence_ratio), maxNumIters(maxIters), relChangeTol(changeTol),
        verbose(v), logData(log), logDirectory(std::move(logDir)) {}

  friend std::ostream &operator<<(std::ostream &os, const PGOAgentParameters &p) {
    os << "PGOAgent parameters: d=" << p.d << " r=" << p.r << " robots=" << p.numRobots
       << " algorithm=" << p.algorithm << " acceleration=" << p.acceleration << " restart=" << p.restartInterval
       << " cost=" << RobustCostNames[p.robustCostType] << " maxIters=" << p.maxNumIters
       << " relChangeTol=" << p.relChangeTol << " verbose=" << p.verbose << " log=" << p.logData << std::endl;
    return os;
  }
};

// what an agent tells its peers about itself
struct PGOAgentStatus {
  unsigned agentID;
  PGOAgentState state;
  unsigned instanceNumber;
  unsigned iterationNumber;
  bool readyToTerminate;
  double relativeChange;
  explicit PGOAgentStatus(unsigned id, PGOAgentState s = PGOAgentState::WAIT_FOR_DATA, unsigned instance = 0,
                          unsigned iteration = 0, bool ready_to_terminate = false, double relative_change = 0)
      : agentID(id), state(s), instanceNumber(instance), iterationNumber(iteration),
        readyToTerminate(ready_to_terminate), relativeChange(relative_change) {}
};

class PGOAgent {
 public:
  PGOAgent(unsigned ID, const PGOAgentParameters &params);
  ~PGOAgent();

  // pose graph + one RBCD iteration
  void setPoseGraph(const std::vector<RelativeSEMeasurement> &inputOdometry,
                    const std::vector<RelativeSEMeasurement> &inputPrivateLoopClosures,
                    const std::vector<RelativeSEMeasurement> &inputSharedLoopClosures, const Matrix &TInit = Matrix());
  void iterate(bool doOptimization = true);
  virtual void reset();
  void initializeAcceleration();

  inline unsigned getID() const { return mID; }
  inline QuadraticProblem *problem() const { return mProblemPtr; }   // B200 extension: the device-resident problem
  inline unsigned num_poses() const { return n; }
  inline unsigned dimension() const { return d; }
  inline unsigned relaxation_rank() const { return r; }
  inline unsigned instance_number() const { return mInstanceNumber; }
  inline unsigned iteration_number() const { return mIterationNumber; }
  inline PGOAgentStatus getStatus() {
    mStatus.agentID = getID();
    mStatus.state = mState;
    mStatus.instanceNumber = instance_number();
    mStatus.iterationNumber = iteration_number();
    return mStatus;
  }
  inline PGOAgentStatus getNeighborStatus(unsigned neighborID) const { return mTeamStatus[neighborID]; }
  inline void setNeighborStatus(const PGOAgentStatus &status) { mTeamStatus[status.agentID] = status; }

  std::vector<unsigned> getNeighborPublicPoses(const unsigned &neighborID) const;
  std::vector<unsigned> getNeighbors() const;

  // rounded trajectories
  bool getTrajectoryInLocalFrame(Matrix &Trajectory);
  bool getTrajectoryInGlobalFrame(Matrix &Trajectory);
  bool getPoseInGlobalFrame(unsigned poseID, Matrix &T);
  bool getNeighborPoseInGlobalFrame(unsigned neighborID, unsigned poseID, Matrix &T);

  // boundary-pose exchange
  bool getSharedPose(unsigned index, Matrix &Mout);
  bool getAuxSharedPose(unsigned index, Matrix &Mout);
  bool getSharedPoseDict(PoseDict &map);
  bool getAuxSharedPoseDict(PoseDict &map);
  void updateNeighborPoses(unsigned neighborID, const PoseDict &poseDict);
  void updateAuxNeighborPoses(unsigned neighborID, const PoseDict &poseDict);

  void setX(const Matrix &Xin);
  bool getX(Matrix &Mout);

  bool shouldTerminate();
  bool shouldRestart() const;
  void restartNesterovAcceleration(bool doOptimization);

  // asynchronous mode
  void startOptimizationLoop(double freq);
  void endOptimizationLoop();
  bool isOptimizationRunning();

  bool getLiftingMatrix(Matrix &M) const;
  void setLiftingMatrix(const Matrix &M);
  void setGlobalAnchor(const Matrix &M);

  // cross-robot frame alignment
  Matrix computeNeighborTransform(const PoseID &nID, const Matrix &var);
  Matrix computeRobustNeighborTransformTwoStage(unsigned neighborID, const PoseDict &poseDict);
  Matrix computeRobustNeighborTransform(unsigned neighborID, const PoseDict &poseDict);
  void initializeInGlobalFrame(unsigned neighborID, const PoseDict &poseDict);

  Matrix localPoseGraphOptimization();

  // result record of the last local solve (B200 extension, read-only)
  const ROPTResult &lastResult() const { return mLastResult; }

 protected:
  unsigned mID, d, r, n;
  const PGOAgentParameters mParams;
  PGOAgentState mState;
  PGOAgentStatus mStatus;
  RobustCost mRobustCost;
  QuadraticProblem *mProblemPtr;
  double mRate{};
  unsigned mInstanceNumber, mIterationNumber, mNumPosesReceived;
  PGOLogger mLogger;
  std::vector<PGOAgentStatus> mTeamStatus;
  bool mOptimizationRequested = false, mPublishPublicPosesRequested = false, mPublishWeightsRequested = false;
  volatile bool mEndLoopRequested = false;

  Matrix X;                             // iterate before rounding, r x (d+1)n
  std::optional<Matrix> XInit, TLocalInit, YLift, globalAnchor;
  vector<RelativeSEMeasurement> odometry, privateLoopClosures, sharedLoopClosures;
  PoseDict neighborPoseDict;
  set<PoseID> localSharedPoseIDs, neighborSharedPoseIDs;
  set<unsigned> neighborRobotIDs;
  mutex mPosesMutex, mNeighborPosesMutex, mMeasurementsMutex;
  thread *mOptimizationThread = nullptr;
  ROPTResult mLastResult;

  void addOdometry(const RelativeSEMeasurement &factor);
  void addPrivateLoopClosure(const RelativeSEMeasurement &factor);
  void addSharedLoopClosure(const RelativeSEMeasurement &factor);
  void constructQMatrix();                              // quadratic part (private edges + shared diagonal terms)
  bool constructGMatrix(const PoseDict &poseDict);      // linear part from the neighbours' public poses
  void localInitialization();
  void runOptimizationLoop();
  RelativeSEMeasurement &findSharedLoopClosureWithNeighbor(const PoseID &nID);
  RelativeSEMeasurement &findSharedLoopClosure(const PoseID &srcID, const PoseID &dstID);
  bool shouldUpdateLoopClosureWeights() const;
  void updateLoopClosuresWeights();
  double computeConvergedLoopClosureRatio();

 private:
  PoseDict neighborAuxPoseDict;
  double gamma{}, alpha{};
  Matrix Y, V, XPrev;
  void updateGamma();
  void updateAlpha();
  bool updateX(bool doOptimization = false, bool acceleration = false);
  void updateY();
  void updateV();
  void resetTeamStatus();
  static bool isDuplicateMeasurement(const RelativeSEMeasurement &m, const vector<RelativeSEMeasurement> &measurements);
};

}  // namespace DPGO
#endif
