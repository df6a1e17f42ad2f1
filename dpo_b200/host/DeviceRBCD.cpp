// DeviceRBCD.cpp -- device-resident multi-GPU RBCD runner of the C++ host API: iterates stay in HBM, the agents'
// public poses travel by ONE ncclAllGather per round (see include/DPGO/DeviceRBCD.h).
// ref: examples/MultiRobotExample.cpp:63-151 (partition), :229-334 (synchronous driver, greedy selection),
//      src/PGOAgent.cpp:95-105,434-458 (public-pose exchange), :783-859 (G), :1131-1137 (updateX constants).
#include <DPGO/DeviceRBCD.h>
#include <DPGO/QuadraticProblem.h>

#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <stdexcept>

#include "dpgo_b200.h"

namespace DPGO {

namespace {
void check(int code, const char *what) {
  if (code != DPGO_OK) throw std::runtime_error(std::string(what) + ": " + dpgo_last_error());
}
void checkNccl(ncclResult_t r, const char *what) {
  if (r != ncclSuccess) throw std::runtime_error(std::string(what) + ": " + ncclGetErrorString(r));
}
}  // namespace

struct DeviceRBCD::Impl {
  unsigned d = 3, r = 5, dh = 4, ts = 20, K = 1, N = 1, perGpu = 1, pmax = 1;
  size_t n = 0;
  std::string schedule;
  std::vector<std::unique_ptr<PGOAgent>> agents;
  std::vector<size_t> count;
  std::vector<std::vector<size_t>> globalOf;      // agent -> global pose ids in local order
  std::vector<dpgo_problem *> h;
  std::vector<int> gpuOf;
  std::vector<void *> stream;
  std::vector<double *> send, gathered;
  std::vector<ncclComm_t> comm;
  std::vector<std::vector<unsigned>> neighbors;
  dpgo_opt_params_t prm;
  unsigned selected = 0;
  bool concurrent = false;         // active agents of a GPU side by side (cluster launches, own streams)
  bool gatheredCurrent = false;    // concurrent mode: the gathered buffers hold every agent's current public tiles
};

DeviceRBCD::DeviceRBCD(const std::vector<RelativeSEMeasurement> &graph, size_t n, unsigned numAgents, const Matrix &XInit,
                       const DeviceRBCDOptions &opt)
    : impl(new Impl()) {
  Impl &I = *impl;
  if (graph.empty()) throw std::runtime_error("DeviceRBCD: empty pose graph");
  I.d = (unsigned)graph[0].t.size();
  I.r = opt.r;
  I.dh = I.d + 1;
  I.ts = I.r * I.dh;
  I.K = numAgents;
  I.N = std::max(1u, opt.gpus);
  I.n = n;
  I.schedule = opt.schedule;
  if (I.schedule != "greedy" && I.schedule != "coloured" && I.schedule != "parallel")
    throw std::runtime_error("DeviceRBCD: schedule must be greedy, coloured or parallel");
  if (I.K == 0 || n / I.K == 0) throw std::runtime_error("DeviceRBCD: more agents than poses");
  if (I.K % I.N != 0) throw std::runtime_error("DeviceRBCD: the agents must divide evenly over the GPUs");
  int ndev = 0;
  check(dpgo_device_count(&ndev), "dpgo_device_count");
  if ((int)I.N > ndev) throw std::runtime_error("DeviceRBCD: fewer CUDA devices than requested GPUs");
  I.perGpu = I.K / I.N;
  const unsigned K = I.K, d = I.d, dh = I.dh, r = I.r;

  // ---- ownership: partition file or contiguous ranges (ref examples/MultiRobotExample.cpp:76-151) ----
  const size_t per = n / K;
  std::vector<unsigned> owner(n), local(n);
  if (!opt.owner.empty() && opt.owner.size() != n) throw std::runtime_error("DeviceRBCD: owner map must have one entry per pose");
  I.count.assign(K, 0);
  for (size_t g = 0; g < n; ++g) {
    owner[g] = opt.owner.empty() ? (unsigned)std::min<size_t>(g / per, K - 1) : opt.owner[g];
    if (owner[g] >= K) throw std::runtime_error("DeviceRBCD: agent id out of range in the owner map");
    local[g] = (unsigned)I.count[owner[g]]++;
  }
  I.globalOf.assign(K, {});
  for (size_t g = 0; g < n; ++g) I.globalOf[owner[g]].push_back(g);
  std::vector<std::vector<RelativeSEMeasurement>> odo(K), priv(K), shared(K);
  for (const auto &e : graph) {
    const unsigned a1 = owner[e.p1], a2 = owner[e.p2];
    RelativeSEMeasurement m(a1, a2, local[e.p1], local[e.p2], e.R, e.t, e.kappa, e.tau);
    m.weight = e.weight;
    if (a1 != a2) { shared[a1].push_back(m); shared[a2].push_back(m); }
    else if (e.p1 + 1 == e.p2) odo[a1].push_back(m);
    else priv[a1].push_back(m);
  }

  // ---- agent graph and its greedy colouring in agent order (needed before the agents exist: it decides the launch mode) ----
  I.neighbors.assign(K, {});
  for (unsigned a = 0; a < K; ++a) {
    for (const auto &m : shared[a]) I.neighbors[a].push_back((unsigned)(m.r1 == a ? m.r2 : m.r1));
    std::sort(I.neighbors[a].begin(), I.neighbors[a].end());
    I.neighbors[a].erase(std::unique(I.neighbors[a].begin(), I.neighbors[a].end()), I.neighbors[a].end());
  }
  mColour.assign(K, 0);
  {
    std::vector<int> col(K, -1);
    for (unsigned a = 0; a < K; ++a) {
      int c = 0;
      for (bool clash = true; clash; ) {
        clash = false;
        for (unsigned b : I.neighbors[a])
          if (col[b] == c) { clash = true; ++c; break; }
      }
      col[a] = c;
      mColour[a] = (unsigned)c;
      mNumColours = std::max(mNumColours, (unsigned)c + 1);
    }
  }
  {
    unsigned most = 1;
    if (I.schedule == "coloured")
      for (unsigned g = 0; g < I.N; ++g)
        for (unsigned c = 0; c < mNumColours; ++c) {
          unsigned cnt = 0;
          for (unsigned a = g * I.perGpu; a < (g + 1) * I.perGpu; ++a) cnt += (mColour[a] == c);
          most = std::max(most, cnt);
        }
    I.concurrent = opt.concurrent < 0 ? (most >= 2) : (opt.concurrent != 0);
    if (I.concurrent && I.schedule == "parallel")
      throw std::runtime_error("DeviceRBCD: concurrent rounds are implemented for the greedy and coloured schedules");
  }

  // ---- streams, agents (Q on the agent's GPU), resident iterates ----
  I.stream.assign(I.N, nullptr);
  for (unsigned g = 0; g < I.N; ++g) check(dpgo_stream_create((int)g, &I.stream[g]), "dpgo_stream_create");
  I.h.assign(K, nullptr);
  I.gpuOf.assign(K, 0);
  Matrix lift;
  for (unsigned a = 0; a < K; ++a) {
    PGOAgentParameters prm(d, r, K);
    prm.algorithm = opt.algorithm;
    prm.preconditioner = opt.preconditioner;
    prm.device = (int)(a / I.perGpu);
    prm.cluster = I.concurrent;
    I.gpuOf[a] = prm.device;
    I.agents.emplace_back(new PGOAgent(a, prm));
    if (a == 0) I.agents[0]->getLiftingMatrix(lift);
    else I.agents[a]->setLiftingMatrix(lift);
    // a zero trajectory of the right shape skips the agent's own chordal initialisation: X comes from XInit
    I.agents[a]->setPoseGraph(odo[a], priv[a], shared[a], Matrix::Zero(d, dh * I.count[a]));
    Matrix Xa0(r, dh * I.count[a]);
    for (size_t q = 0; q < I.count[a]; ++q) Xa0.block(0, q * dh, r, dh) = Matrix(XInit).block(0, I.globalOf[a][q] * dh, r, dh);
    I.agents[a]->setX(Xa0);
    I.h[a] = I.agents[a]->problem()->handle();
    if (!I.h[a]) throw std::runtime_error("DeviceRBCD: agent without a device problem");
    check(dpgo_problem_set_stream(I.h[a], I.stream[(size_t)I.gpuOf[a]]), "dpgo_problem_set_stream");
    Matrix Xa;
    I.agents[a]->getX(Xa);
    check(dpgo_problem_upload_X(I.h[a], Xa.data()), "dpgo_problem_upload_X");
  }

  // ---- exchange plan: public poses, padded slots, per-agent edge tables ----
  std::vector<std::vector<int32_t>> pub(K);
  for (unsigned a = 0; a < K; ++a) {
    for (const auto &m : shared[a]) pub[a].push_back((int32_t)(m.r1 == a ? m.p1 : m.p2));
    std::sort(pub[a].begin(), pub[a].end());
    pub[a].erase(std::unique(pub[a].begin(), pub[a].end()), pub[a].end());
    I.pmax = std::max<unsigned>(I.pmax, (unsigned)pub[a].size());
  }
  for (unsigned a = 0; a < K; ++a) {
    const size_t m = shared[a].size();
    std::vector<int32_t> loc(m), slot(m), outg(m);
    std::vector<double> T(m * dh * dh, 0.0), om(m * dh, 0.0);
    for (size_t e = 0; e < m; ++e) {
      const RelativeSEMeasurement &s = shared[a][e];
      const bool out = (s.r1 == a);
      const unsigned b = (unsigned)(out ? s.r2 : s.r1);
      const int32_t q = (int32_t)(out ? s.p2 : s.p1);
      loc[e] = (int32_t)(out ? s.p1 : s.p2);
      const auto it = std::lower_bound(pub[b].begin(), pub[b].end(), q);
      slot[e] = (int32_t)(b * I.pmax + (unsigned)(it - pub[b].begin()));
      outg[e] = out ? 1 : 0;
      for (unsigned i = 0; i < d; ++i) {
        for (unsigned j = 0; j < d; ++j) T[e * dh * dh + i * dh + j] = s.R(i, j);
        T[e * dh * dh + i * dh + d] = s.t(i);
        om[e * dh + i] = s.weight * s.kappa;
      }
      T[e * dh * dh + d * dh + d] = 1.0;
      om[e * dh + d] = s.weight * s.tau;
    }
    check(dpgo_agent_set_public_poses(I.h[a], (int)pub[a].size(), pub[a].data()), "dpgo_agent_set_public_poses");
    check(dpgo_agent_set_shared_edges(I.h[a], (int)m, loc.data(), slot.data(), outg.data(), T.data(), om.data()),
          "dpgo_agent_set_shared_edges");
  }
  // ---- exchange buffers and communicators ----
  const size_t slotElems = (size_t)I.pmax * I.ts;
  I.send.assign(I.N, nullptr);
  I.gathered.assign(I.N, nullptr);
  for (unsigned g = 0; g < I.N; ++g) {
    void *p = nullptr;
    check(dpgo_device_malloc((int)g, sizeof(double) * K * slotElems, &p), "dpgo_device_malloc");
    I.gathered[g] = static_cast<double *>(p);
    if (I.N == 1) {
      I.send[g] = I.gathered[g];                       // a single GPU packs straight into the gathered layout
    } else {
      check(dpgo_device_malloc((int)g, sizeof(double) * I.perGpu * slotElems, &p), "dpgo_device_malloc");
      I.send[g] = static_cast<double *>(p);
    }
  }
  if (I.N > 1) {
    std::vector<int> devs(I.N);
    for (unsigned g = 0; g < I.N; ++g) devs[g] = (int)g;
    I.comm.assign(I.N, nullptr);
    checkNccl(ncclCommInitAll(I.comm.data(), (int)I.N, devs.data()), "ncclCommInitAll");
  }
  dpgo_opt_params_default(&I.prm);
  I.prm.algorithm = (opt.algorithm == ROPTALG::RTR) ? DPGO_ALG_RTR : DPGO_ALG_RGD;
  I.prm.precond = (int)opt.preconditioner;
  I.prm.tr_tolerance = 1e-2;          // ref src/PGOAgent.cpp:1134-1137
  I.prm.tr_iterations = 1;
  I.prm.tr_max_inner = 10;
  I.prm.tr_initial_radius = 100;
}

DeviceRBCD::~DeviceRBCD() {
  if (!impl) return;
  Impl &I = *impl;
  try { sync(); } catch (...) {}
  for (ncclComm_t c : I.comm)
    if (c) ncclCommDestroy(c);
  I.agents.clear();                                    // problems first: they use the streams
  for (unsigned g = 0; g < I.N; ++g) {
    if (I.N > 1 && I.send[g]) dpgo_device_free((int)g, I.send[g]);
    if (I.gathered[g]) dpgo_device_free((int)g, I.gathered[g]);
    if (I.stream[g]) dpgo_stream_destroy((int)g, I.stream[g]);
  }
}

size_t DeviceRBCD::allGatherBytesPerGpu() const { return sizeof(double) * impl->perGpu * impl->pmax * impl->ts; }

void DeviceRBCD::exchange() {
  Impl &I = *impl;
  const size_t slotElems = (size_t)I.pmax * I.ts;
  for (unsigned a = 0; a < I.K; ++a) {
    const size_t g = (size_t)I.gpuOf[a];
    double *dst = (I.N == 1) ? I.gathered[g] + a * slotElems : I.send[g] + (a % I.perGpu) * slotElems;
    check(dpgo_agent_pack_public(I.h[a], dst), "dpgo_agent_pack_public");
  }
  if (I.N > 1) {
    checkNccl(ncclGroupStart(), "ncclGroupStart");
    for (unsigned g = 0; g < I.N; ++g) {
      check(dpgo_device_set((int)g), "dpgo_device_set");
      checkNccl(ncclAllGather(I.send[g], I.gathered[g], I.perGpu * slotElems, ncclDouble, I.comm[g], (cudaStream_t)I.stream[g]),
                "ncclAllGather");
    }
    checkNccl(ncclGroupEnd(), "ncclGroupEnd");
  }
  for (unsigned a = 0; a < I.K; ++a)
    check(dpgo_agent_build_G(I.h[a], I.gathered[(size_t)I.gpuOf[a]], (int64_t)I.K * I.pmax), "dpgo_agent_build_G");
}

void DeviceRBCD::sync() {
  Impl &I = *impl;
  for (unsigned g = 0; g < I.N; ++g) check(dpgo_stream_synchronize((int)g, I.stream[g]), "dpgo_stream_synchronize");
}

static std::vector<unsigned> activeSet(const std::string &schedule, unsigned K, unsigned selected, unsigned round,
                                       const std::vector<unsigned> &colour, unsigned ncolours) {
  std::vector<unsigned> act;
  if (schedule == "greedy") act.push_back(selected);
  else
    for (unsigned a = 0; a < K; ++a)
      if (schedule == "parallel" || colour[a] == round % ncolours) act.push_back(a);
  return act;
}

bool DeviceRBCD::concurrent() const { return impl->concurrent; }

// the active agents of every GPU side by side: per GPU ONE call (fork from the GPU's stream, per agent G rebuild -> RTR
// step -> pack on its own stream, join), then the all-gather that publishes the new public tiles
void DeviceRBCD::roundConcurrent(const std::vector<unsigned> &active) {
  Impl &I = *impl;
  const size_t slotElems = (size_t)I.pmax * I.ts;
  for (unsigned g = 0; g < I.N; ++g) {
    std::vector<dpgo_problem *> hs;
    std::vector<double *> dst;
    for (unsigned a : active)
      if ((unsigned)I.gpuOf[a] == g) {
        hs.push_back(I.h[a]);
        dst.push_back((I.N == 1) ? I.gathered[g] + a * slotElems : I.send[g] + (a % I.perGpu) * slotElems);
      }
    if (hs.empty()) continue;
    check(dpgo_agents_round_async(hs.data(), (int)hs.size(), &I.prm, I.gathered[g], (int64_t)I.K * I.pmax, dst.data(), I.stream[g], 0),
          "dpgo_agents_round_async");
  }
  if (I.N > 1) {
    checkNccl(ncclGroupStart(), "ncclGroupStart");
    for (unsigned g = 0; g < I.N; ++g) {
      check(dpgo_device_set((int)g), "dpgo_device_set");
      checkNccl(ncclAllGather(I.send[g], I.gathered[g], I.perGpu * slotElems, ncclDouble, I.comm[g], (cudaStream_t)I.stream[g]),
                "ncclAllGather");
    }
    checkNccl(ncclGroupEnd(), "ncclGroupEnd");
  }
}

void DeviceRBCD::runRounds(unsigned rounds) {
  Impl &I = *impl;
  for (unsigned it = 0; it < rounds; ++it) {
    const std::vector<unsigned> act = activeSet(I.schedule, I.K, I.selected, mRound, mColour, mNumColours);
    if (I.concurrent) {
      if (!I.gatheredCurrent) { exchange(); I.gatheredCurrent = true; }
      roundConcurrent(act);
    } else {
      exchange();
      for (unsigned a : act) check(dpgo_optimize_resident_async(I.h[a], &I.prm), "dpgo_optimize_resident_async");
    }
    ++mRound;
  }
}

DeviceRBCDStats DeviceRBCD::step(bool evaluate) {
  Impl &I = *impl;
  DeviceRBCDStats st;
  st.active = activeSet(I.schedule, I.K, I.selected, mRound, mColour, mNumColours);
  dpgo_opt_result_t res;
  if (I.concurrent) {
    if (!I.gatheredCurrent) { exchange(); I.gatheredCurrent = true; }
    roundConcurrent(st.active);
  } else {
    exchange();
    for (unsigned a : st.active) check(dpgo_optimize_resident_async(I.h[a], &I.prm), "dpgo_optimize_resident_async");
    for (unsigned a : st.active) check(dpgo_optimize_result(I.h[a], &res), "dpgo_optimize_result");
  }
  ++mRound;
  if (!evaluate) return st;
  exchange();                                          // fresh neighbour poses for the central gradient
  double cost = 0, gn2 = 0, best = -1;
  unsigned arg = I.selected;
  for (unsigned a = 0; a < I.K; ++a) {
    double f = 0, nrm = 0;
    check(dpgo_agent_f_rgradnorm_resident(I.h[a], &f, &nrm), "dpgo_agent_f_rgradnorm_resident");
    check(dpgo_optimize_result(I.h[a], &res), "dpgo_optimize_result");
    cost += res.quad_init + res.lin_init;              // sum over agents = 2 f_central (cross terms counted once)
    gn2 += nrm * nrm;
    if (nrm > best) { best = nrm; arg = a; }
  }
  st.cost = cost;
  st.gradnorm = std::sqrt(gn2);
  if (I.schedule == "greedy" && !I.neighbors[I.selected].empty()) I.selected = arg;   // ref :308-325
  return st;
}

Matrix DeviceRBCD::assemble() {
  Impl &I = *impl;
  Matrix X(I.r, I.dh * I.n);
  for (unsigned a = 0; a < I.K; ++a) {
    Matrix Xa(I.r, I.dh * I.count[a]);
    check(dpgo_problem_download_X(I.h[a], Xa.data()), "dpgo_problem_download_X");
    for (size_t q = 0; q < I.count[a]; ++q) X.block(0, I.globalOf[a][q] * I.dh, I.r, I.dh) = Xa.block(0, q * I.dh, I.r, I.dh);
  }
  return X;
}

}  // namespace DPGO
