// DeviceRBCD.h -- B200 extension of the C++ host API: the device-resident multi-GPU RBCD runner.
//
// The reference's drivers keep every agent's iterate in host Eigen matrices and exchange public poses through
// std::map "PoseDict"s (examples/MultiRobotExample.cpp:229-334, src/PGOAgent.cpp:95-105,434-458).  This runner keeps
// the iterates in HBM and replaces that "network" by ONE ncclAllGather per round over NVLink:
//
//   per round, on every GPU's stream:   dpgo_agent_pack_public  (public tiles -> send buffer)
//                                       ncclAllGather           (padded public-pose slots of all agents)
//                                       dpgo_agent_build_G      (linear term from the gathered tiles, ref :783-859)
//                                       dpgo_optimize_resident_async for the agents of the round
//
// K agents are spread over N GPUs of one node (K % N == 0, contiguous blocks), one process, one stream and one NCCL
// communicator per GPU (ncclCommInitAll).  Schedules: "greedy" (the reference's: one agent per round, argmax of the block
// gradient norms -- reproduces the shipped traces), "coloured" (all agents of one colour class of the agent graph per
// round: same RBCD semantics, concurrent), "parallel" (all agents on the previous round's poses).
// When a GPU hosts several agents of one colour class, their steps run side by side (one thread-block cluster and one
// stream per agent, the whole round of a GPU replayed as a CUDA graph): DeviceRBCDOptions::concurrent.
#ifndef DPGO_DEVICE_RBCD_H
#define DPGO_DEVICE_RBCD_H

#include <DPGO/DPGO_types.h>
#include <DPGO/PGOAgent.h>
#include <DPGO/RelativeSEMeasurement.h>

#include <memory>
#include <string>
#include <vector>

namespace DPGO {

struct DeviceRBCDOptions {
  unsigned r = 5;
  unsigned gpus = 1;
  std::string schedule = "greedy";
  ROPTALG algorithm = ROPTALG::RTR;
  Preconditioner preconditioner = Preconditioner::SparseExact;
  // pose -> agent (one entry per pose, e.g. from a graph-partition file, ref examples/MultiRobotExample.cpp:76-91);
  // empty: contiguous ranges, the last agent takes the remainder (ref :95-109)
  std::vector<unsigned> owner;
  // the active agents of a round that share a GPU step side by side, each as one thread-block cluster on its own stream
  // (dpgo_agents_round_async; greedy / coloured schedules): -1 = when some GPU hosts >= 2 agents of one colour class
  int concurrent = -1;
};

struct DeviceRBCDStats {
  double cost = 0;        // 2 f of the assembled iterate (centralised cost)
  double gradnorm = 0;    // norm of the centralised Riemannian gradient
  std::vector<unsigned> active;
};

class DeviceRBCD {
 public:
  // graph: the global pose graph (global pose ids); XInit: r x (d+1)n lifted initial iterate
  DeviceRBCD(const std::vector<RelativeSEMeasurement> &graph, size_t n, unsigned numAgents, const Matrix &XInit,
             const DeviceRBCDOptions &options);
  ~DeviceRBCD();
  DeviceRBCD(const DeviceRBCD &) = delete;
  DeviceRBCD &operator=(const DeviceRBCD &) = delete;

  void exchange();                              // pack -> all-gather -> G rebuild, asynchronous on the GPU streams
  DeviceRBCDStats step(bool evaluate = true);   // one round (+ central cost / gradient norm / greedy selection)
  void runRounds(unsigned rounds);              // rounds without evaluation (throughput), asynchronous; call sync()
  void sync();
  Matrix assemble();                            // r x (d+1)n iterate on the host
  bool concurrent() const;
  unsigned numColours() const { return mNumColours; }
  const std::vector<unsigned> &colours() const { return mColour; }
  unsigned round() const { return mRound; }
  size_t allGatherBytesPerGpu() const;

 private:
  struct Impl;
  void roundConcurrent(const std::vector<unsigned> &active);
  std::unique_ptr<Impl> impl;
  unsigned mNumColours = 1, mRound = 0;
  std::vector<unsigned> mColour;
};

}  // namespace DPGO
#endif
