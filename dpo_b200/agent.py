"""PGOAgent mirror + the multi-GPU runner (one agent per GPU, boundary poses over one all-gather).

`PGOAgent` keeps the reference's public interface for the parts on / next to the hot path
(ref include/DPGO/PGOAgent.h:209-490, src/PGOAgent.cpp): setPoseGraph, setX/getX, getSharedPoseDict,
updateNeighborPoses, iterate, getNeighbors, getTrajectoryInLocalFrame, localPoseGraphOptimization.
Host bookkeeping stays on the host, every numeric step runs in libdpgo_b200.so.

`ExchangePlan` + `DistributedPGO` are the B200-native replacement of the reference's in-process
"network" (examples/MultiRobotExample.cpp:245-256): public poses are packed on the device,
exchanged with ONE all-gather per round (NCCL over NVLink when the tensors are CUDA), and G is
rebuilt on the device from the gathered tiles (ref PGOAgent::constructGMatrix, src/PGOAgent.cpp:783-859).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi as capi
from . import posegraph as pg
from .posegraph import EdgeSet
from .problem import QuadraticOptimizer, QuadraticProblem, ROPTALG

PoseID = Tuple[int, int]


@dataclass
class PGOAgentParameters:
    """ref: include/DPGO/PGOAgent.h:59-136 (fields on the hot path; robust-cost knobs are out of scope)."""
    d: int
    r: int
    numRobots: int = 1
    algorithm: int = ROPTALG.RTR
    acceleration: bool = False
    restartInterval: int = 30
    maxNumIters: int = 500
    relChangeTol: float = 5e-3
    verbose: bool = False
    preconditioner: int = capi.PRECOND_SPARSE_EXACT     # B200 extension (reference operator by default)
    device: int = 0
    cluster: bool = False          # B200 extension: step kernel as one thread-block cluster (several agents per GPU)


class PGOAgentState:
    WAIT_FOR_DATA, WAIT_FOR_INITIALIZATION, INITIALIZED = 0, 1, 2


# ---------------------------------------------------------------------------------------------------
# partitioning (ref examples/MultiRobotExample.cpp:63-151)
# ---------------------------------------------------------------------------------------------------
def contiguous_owner(n: int, k: int) -> np.ndarray:
    """Pose -> agent, contiguous ranges, last agent takes the remainder (ref :95-109)."""
    per = n // k
    if per <= 0:
        raise ValueError("More robots than total number of poses! Decrease the number of robots")
    return np.minimum(np.arange(n) // per, k - 1).astype(np.int64)


def partition_edges(edges: EdgeSet, owner: np.ndarray, k: int):
    """Split a global edge list into per-agent (odometry, private, shared) sets with local pose ids
    (ref :115-151).  Returns (parts, counts, global_index_of[a])."""
    n = owner.shape[0]
    counts = np.bincount(owner, minlength=k).astype(np.int64)
    local = np.empty(n, dtype=np.int64)
    glob = []
    for a in range(k):
        idx = np.flatnonzero(owner == a)
        local[idx] = np.arange(idx.shape[0])
        glob.append(idx)
    a1, a2 = owner[edges.p1], owner[edges.p2]
    re = EdgeSet(edges.d, a1, a2, local[edges.p1], local[edges.p2], edges.R, edges.t, edges.kappa, edges.tau,
                 edges.weight)
    same = a1 == a2
    odo = edges.p1 + 1 == edges.p2                     # ref :134 tests GLOBAL ids
    parts = []
    for a in range(k):
        mine = same & (a1 == a)
        parts.append((re.take(np.flatnonzero(mine & odo)), re.take(np.flatnonzero(mine & ~odo)),
                      re.take(np.flatnonzero(~same & ((a1 == a) | (a2 == a))))))
    return parts, counts, glob


# ---------------------------------------------------------------------------------------------------
# exchange plan: who publishes what, where it lands in the gathered buffer
# ---------------------------------------------------------------------------------------------------
class ExchangePlan:
    """Static tables of the boundary-pose exchange for k agents.

    public[a]   sorted local ids of agent a's public poses (ref localSharedPoseIDs, src/PGOAgent.cpp:236-245)
    pmax        padded slot count per agent (all-gather needs equal counts)
    slot(b, q)  = b * pmax + position of q in public[b]
    edge tables for agent a: local pose, neighbour slot, outgoing flag, T (row-major), omega
    """

    def __init__(self, shared: Sequence[EdgeSet], k: int):
        self.k = k
        self.public: List[np.ndarray] = []
        for a in range(k):
            s = shared[a]
            mine = np.where(s.r1 == a, s.p1, s.p2)
            self.public.append(np.unique(mine).astype(np.int32))
        self.pmax = max(1, max(len(p) for p in self.public))
        self._pos = [{int(q): i for i, q in enumerate(p)} for p in self.public]
        self.tables = []
        for a in range(k):
            s = shared[a]
            out = (s.r1 == a)
            local = np.where(out, s.p1, s.p2).astype(np.int32)
            nbr_agent = np.where(out, s.r2, s.r1)
            nbr_pose = np.where(out, s.p2, s.p1)
            slot = np.array([int(b) * self.pmax + self._pos[int(b)][int(q)] for b, q in zip(nbr_agent, nbr_pose)],
                            dtype=np.int32)
            self.tables.append(dict(local=local, slot=slot, outgoing=out.astype(np.int32),
                                    T=np.ascontiguousarray(s.homogeneous()), omega=np.ascontiguousarray(s.omega()),
                                    neighbors=sorted(set(int(b) for b in nbr_agent))))

    def slot(self, agent: int, local_pose: int) -> int:
        return agent * self.pmax + self._pos[agent][int(local_pose)]

    def colouring(self) -> List[int]:
        """Greedy colouring of the agent graph: agents of one colour share no edge, so they may update
        concurrently with exactly the sequential RBCD semantics (hard part 4)."""
        colour = [-1] * self.k
        for a in range(self.k):
            used = {colour[b] for b in self.tables[a]["neighbors"] if colour[b] >= 0}
            c = 0
            while c in used:
                c += 1
            colour[a] = c
        return colour


# ---------------------------------------------------------------------------------------------------
# PGOAgent
# ---------------------------------------------------------------------------------------------------
class PGOAgent:
    def __init__(self, ID: int, params: PGOAgentParameters):
        self.mID = int(ID)
        self.mParams = params
        self.d, self.r, self.n = params.d, params.r, 1
        self.mState = PGOAgentState.WAIT_FOR_DATA
        self.mIterationNumber = 0
        self.mInstanceNumber = 0
        self.X = np.zeros((self.r, self.d + 1))
        self.X[:self.d, :self.d] = np.eye(self.d)
        self.YLift: Optional[np.ndarray] = pg.fixedStiefelVariable(self.d, self.r) if ID == 0 else None
        self.globalAnchor: Optional[np.ndarray] = None
        self.mProblem: Optional[QuadraticProblem] = None
        self.neighborPoseDict: Dict[PoseID, np.ndarray] = {}
        self.localSharedPoseIDs: List[PoseID] = []
        self.neighborSharedPoseIDs: set = set()
        self.neighborRobotIDs: List[int] = []
        self.relativeChange = 0.0
        self.readyToTerminate = False
        self.lastResult = None
        self.TLocalInit: Optional[np.ndarray] = None
        # Nesterov acceleration (ref src/PGOAgent.cpp:1040-1091)
        self.gamma = 0.0
        self.alpha = 0.0
        self.Y: Optional[np.ndarray] = None
        self.V: Optional[np.ndarray] = None
        self.XPrev: Optional[np.ndarray] = None
        self.neighborAuxPoseDict: Dict[PoseID, np.ndarray] = {}

    # -- getters (ref .h:237-262) --
    def getID(self): return self.mID
    def num_poses(self): return self.n
    def dimension(self): return self.d
    def relaxation_rank(self): return self.r
    def iteration_number(self): return self.mIterationNumber
    def instance_number(self): return self.mInstanceNumber
    def getNeighbors(self): return list(self.neighborRobotIDs)

    def getLiftingMatrix(self):
        assert self.mID == 0
        return self.YLift

    def setLiftingMatrix(self, M):
        M = np.asarray(M, dtype=float)
        assert M.shape == (self.r, self.d)
        self.YLift = M

    def setGlobalAnchor(self, M):
        self.globalAnchor = np.asarray(M, dtype=float)

    # -- pose graph (ref src/PGOAgent.cpp:126-195) --
    def setPoseGraph(self, odometry: EdgeSet, privateLoopClosures: EdgeSet, sharedLoopClosures: EdgeSet,
                     TInit: Optional[np.ndarray] = None, n: Optional[int] = None) -> None:
        assert self.mState == PGOAgentState.WAIT_FOR_DATA and self.n == 1
        if len(odometry) == 0 and n is None:
            return
        self.odometry, self.privateLoopClosures, self.sharedLoopClosures = odometry, privateLoopClosures, sharedLoopClosures
        nn = 1
        for s in (odometry, privateLoopClosures):
            if len(s):
                nn = max(nn, int(max(s.p1.max(), s.p2.max())) + 1)
        sh = sharedLoopClosures
        if len(sh):
            mine = np.where(sh.r1 == self.mID, sh.p1, sh.p2)
            nn = max(nn, int(mine.max()) + 1)
            other_r = np.where(sh.r1 == self.mID, sh.r2, sh.r1)
            other_p = np.where(sh.r1 == self.mID, sh.p2, sh.p1)
            self.localSharedPoseIDs = [(self.mID, int(q)) for q in np.unique(mine)]
            self.neighborSharedPoseIDs = {(int(a), int(q)) for a, q in zip(other_r, other_p)}
            self.neighborRobotIDs = sorted({int(a) for a in other_r})
        self.n = nn if n is None else int(n)
        self.mProblem = QuadraticProblem(self.n, self.d, self.r, device=self.mParams.device,
                                         preconditioners=self._precond_set(), cluster=self.mParams.cluster)
        self.constructQMatrix()
        if TInit is not None and np.shape(TInit) == (self.d, (self.d + 1) * self.n):
            self.TLocalInit = np.array(TInit, dtype=float)
        else:
            self.localInitialization()
        self.mState = PGOAgentState.WAIT_FOR_INITIALIZATION
        if self.mID == 0 and self.YLift is not None:
            self.X = self.YLift @ self.TLocalInit
            self.mState = PGOAgentState.INITIALIZED

    def _precond_set(self):
        s = {capi.PRECOND_BLOCK_JACOBI}
        s.add(self.mParams.preconditioner)
        s.discard(capi.PRECOND_NONE)
        return tuple(sorted(s))

    def localInitialization(self) -> None:
        """ref src/PGOAgent.cpp:945-962 (L2 cost -> chordal initialisation on the private edges)."""
        priv = EdgeSet.join([self.odometry, self.privateLoopClosures])
        self.TLocalInit = pg.chordalInitialization(self.d, self.n, priv)

    def constructQMatrix(self) -> None:
        """Private edges' Laplacian + diagonal terms of the shared edges (ref src/PGOAgent.cpp:720-781)."""
        priv = EdgeSet.join([self.odometry, self.privateLoopClosures])
        sh = self.sharedLoopClosures
        idx, W = None, None
        if len(sh):
            T, om = sh.homogeneous(), sh.omega()
            out = sh.r1 == self.mID
            W = np.zeros_like(T)
            ar = np.arange(self.d + 1)
            W[:, ar, ar] = om                                              # incoming: Omega at p2
            Wout = (T * om[:, None, :]) @ np.transpose(T, (0, 2, 1))       # outgoing: T Omega T^T at p1
            W[out] = Wout[out]
            idx = np.where(out, sh.p1, sh.p2).astype(np.int32)
        # the private edges' Laplacian is assembled on the device from the edge records (k_assemble_Q); the shared edges'
        # diagonal terms enter as static blocks; odometry edges keep their weights under robust re-weighting
        fixed = np.concatenate([np.ones(len(self.odometry), dtype=np.int32), np.zeros(len(self.privateLoopClosures), dtype=np.int32)])
        self.mProblem.setEdges(priv, idx, W, fixed=fixed)

    def constructGMatrix(self, poseDict: Dict[PoseID, np.ndarray]) -> bool:
        """Host form (dictionary of neighbour poses), as the reference does it (src/PGOAgent.cpp:783-859)."""
        sh = self.sharedLoopClosures
        dh = self.d + 1
        G = np.zeros((self.r, dh * self.n))
        T, om = sh.homogeneous(), sh.omega()
        for k in range(len(sh)):
            if sh.r1[k] == self.mID:
                nid = (int(sh.r2[k]), int(sh.p2[k]))
                if nid not in poseDict:
                    return False
                G[:, int(sh.p1[k]) * dh:(int(sh.p1[k]) + 1) * dh] -= (poseDict[nid] * om[k][None, :]) @ T[k].T
            else:
                nid = (int(sh.r1[k]), int(sh.p1[k]))
                if nid not in poseDict:
                    return False
                G[:, int(sh.p2[k]) * dh:(int(sh.p2[k]) + 1) * dh] -= (poseDict[nid] @ T[k]) * om[k][None, :]
        self.mProblem.setG(G)
        return True

    # -- iterate exchange (ref src/PGOAgent.cpp:55-118, 434-458) --
    def setX(self, Xin) -> None:
        Xin = np.asarray(Xin, dtype=float)
        assert self.mState != PGOAgentState.WAIT_FOR_DATA
        assert Xin.shape == (self.r, (self.d + 1) * self.n)
        self.X = Xin.copy()
        self.mState = PGOAgentState.INITIALIZED
        if self.mParams.acceleration:
            self.initializeAcceleration()                  # ref :60-62

    def getX(self) -> np.ndarray:
        return self.X.copy()

    def getSharedPose(self, index: int):
        if self.mState != PGOAgentState.INITIALIZED or index >= self.n:
            return None
        dh = self.d + 1
        return self.X[:, index * dh:(index + 1) * dh].copy()

    def getSharedPoseDict(self) -> Optional[Dict[PoseID, np.ndarray]]:
        if self.mState != PGOAgentState.INITIALIZED:
            return None
        dh = self.d + 1
        return {pid: self.X[:, pid[1] * dh:(pid[1] + 1) * dh].copy() for pid in self.localSharedPoseIDs}

    def updateNeighborPoses(self, neighborID: int, poseDict: Dict[PoseID, np.ndarray]) -> None:
        assert neighborID != self.mID
        for nid, var in poseDict.items():
            assert nid[0] == neighborID and var.shape == (self.r, self.d + 1)
            if nid in self.neighborSharedPoseIDs and self.mState == PGOAgentState.INITIALIZED:
                self.neighborPoseDict[nid] = np.array(var)

    # -- Nesterov acceleration (ref src/PGOAgent.cpp:60-62, 107-118, 460-479, 1040-1091) --
    def initializeAcceleration(self) -> None:
        self.XPrev, self.V, self.Y = self.X.copy(), self.X.copy(), self.X.copy()
        self.gamma = self.alpha = 0.0

    def getAuxSharedPoseDict(self) -> Optional[Dict[PoseID, np.ndarray]]:
        if self.mState != PGOAgentState.INITIALIZED or self.Y is None:
            return None
        dh = self.d + 1
        return {pid: self.Y[:, pid[1] * dh:(pid[1] + 1) * dh].copy() for pid in self.localSharedPoseIDs}

    def updateAuxNeighborPoses(self, neighborID: int, poseDict: Dict[PoseID, np.ndarray]) -> None:
        assert neighborID != self.mID
        for nid, var in poseDict.items():
            if nid in self.neighborSharedPoseIDs and self.mState == PGOAgentState.INITIALIZED:
                self.neighborAuxPoseDict[nid] = np.array(var)

    # -- one RBCD step (ref src/PGOAgent.cpp:642-718, updateX :1093-1165) --
    def iterate(self, doOptimization: bool = True) -> bool:
        self.mIterationNumber += 1
        if self.mState != PGOAgentState.INITIALIZED:
            return True
        if not self.mParams.acceleration:
            return self._updateX(doOptimization, False)
        if self.Y is None:
            self.initializeAcceleration()
        self.XPrev = self.X.copy()
        N = float(self.mParams.numRobots)
        self.gamma = (1 + np.sqrt(1 + 4 * N * N * self.gamma * self.gamma)) / (2 * N)        # ref :1065-1069
        self.alpha = 1.0 / (self.gamma * N)                                                    # ref :1071-1075
        self.Y = self.mProblem.project((1 - self.alpha) * self.X + self.alpha * self.V)       # ref :1077-1083
        ok = self._updateX(doOptimization, True)
        self.V = self.mProblem.project(self.V + self.gamma * (self.X - self.Y))               # ref :1085-1091
        if (self.mIterationNumber + 1) % self.mParams.restartInterval == 0:                    # ref :1033-1052
            self.X = self.XPrev
            self._updateX(doOptimization, False)
            self.V, self.Y = self.X.copy(), self.X.copy()
            self.gamma = self.alpha = 0.0
        return ok

    def _updateX(self, doOptimization: bool, acceleration: bool) -> bool:
        if not doOptimization:
            if acceleration:
                self.X = self.Y.copy()
            return True
        XPrev = self.X
        if not self.constructGMatrix(self.neighborAuxPoseDict if acceleration else self.neighborPoseDict):
            if self.mParams.verbose:
                print(f"Robot {self.mID} could not construct G matrix. Skip update...")
            self.readyToTerminate = False
            return False
        opt = QuadraticOptimizer(self.mProblem)
        opt.setVerbose(self.mParams.verbose)
        opt.setAlgorithm(self.mParams.algorithm)
        opt.setTrustRegionTolerance(1e-2)              # ref :1134-1137
        opt.setTrustRegionIterations(1)
        opt.setTrustRegionMaxInnerIterations(10)
        opt.setTrustRegionInitialRadius(100)
        opt.setPreconditioner(self.mParams.preconditioner)
        self.X = np.array(opt.optimize(self.Y if acceleration else self.X))
        self.lastResult = opt.getOptResult()
        self.relativeChange = float(np.sqrt(np.sum((self.X - XPrev) ** 2) / self.n))
        self.readyToTerminate = self.relativeChange <= self.mParams.relChangeTol
        return True

    def localPoseGraphOptimization(self) -> np.ndarray:
        """ref src/PGOAgent.cpp:964-990: r = d problem on the private edges, RTR 10 outer / 50 inner."""
        if self.TLocalInit is None:
            self.localInitialization()
        priv = EdgeSet.join([self.odometry, self.privateLoopClosures])
        prob = QuadraticProblem(self.n, self.d, self.d, device=self.mParams.device, preconditioners=self._precond_set())
        prob.setQ_blocks(*pg.connection_laplacian_blocks(priv))
        opt = QuadraticOptimizer(prob)
        opt.setVerbose(self.mParams.verbose)
        opt.setTrustRegionInitialRadius(10)
        opt.setTrustRegionIterations(10)
        opt.setTrustRegionTolerance(1e-1)
        opt.setTrustRegionMaxInnerIterations(50)
        opt.setPreconditioner(self.mParams.preconditioner)
        Topt = np.array(opt.optimize(self.TLocalInit))
        self.lastResult = opt.getOptResult()
        prob.close()
        return Topt

    def getTrajectoryInLocalFrame(self) -> Optional[np.ndarray]:
        """ref src/PGOAgent.cpp:481-498."""
        if self.mState != PGOAgentState.INITIALIZED:
            return None
        d, dh = self.d, self.d + 1
        T = self.X[:, :d].T @ self.X
        t0 = T[:, d].copy()
        for i in range(self.n):
            T[:, i * dh:i * dh + d] = pg.projectToRotationGroup(T[:, i * dh:i * dh + d])
            T[:, i * dh + d] -= t0
        return T

    # -- device-resident exchange path -----------------------------------------------------------------
    def attach_exchange(self, plan: ExchangePlan) -> None:
        lib, h = self.mProblem._lib, self.mProblem._h
        pub = np.ascontiguousarray(plan.public[self.mID], dtype=np.int32)
        capi.check(lib.dpgo_agent_set_public_poses(h, len(pub), capi.iptr(pub)))
        tb = plan.tables[self.mID]
        capi.check(lib.dpgo_agent_set_shared_edges(h, len(tb["local"]), capi.iptr(tb["local"]), capi.iptr(tb["slot"]),
                                                   capi.iptr(tb["outgoing"]), capi.dptr(tb["T"]), capi.dptr(tb["omega"])))

    def pack_public(self, send_ptr: int) -> None:
        capi.check(self.mProblem._lib.dpgo_agent_pack_public(self.mProblem._h, C.c_void_p(send_ptr)))

    def build_G(self, gathered_ptr: int, num_slots: int) -> None:
        capi.check(self.mProblem._lib.dpgo_agent_build_G(self.mProblem._h, C.c_void_p(gathered_ptr), num_slots))


# ---------------------------------------------------------------------------------------------------
# multi-agent runner
# ---------------------------------------------------------------------------------------------------
@dataclass
class RoundStats:
    cost: float            # 2 f_central
    gradnorm: float        # |grad_central|
    selected: List[int]


def auto_concurrent(colour: Sequence[int], k: int, world: int, schedule: str, acceleration: bool) -> bool:
    """Default launch mode of a k-agent run over `world` ranks (contiguous blocks of k/world agents per rank): the agents of
    a round step side by side (thread-block clusters, dpgo_agents_round_async) when some rank hosts >= 2 agents of one
    colour class under the coloured schedule.  A pure function of the global plan, so every rank decides alike."""
    if schedule != "coloured" or acceleration:
        return False
    per_rank = k // max(world, 1)
    ncol = max(colour) + 1
    most = max(sum(1 for a in range(q * per_rank, (q + 1) * per_rank) if colour[a] == c)
               for q in range(max(world, 1)) for c in range(ncol))
    return most >= 2


class DistributedPGO:
    """One agent per rank (torch.distributed) or all agents in one process (single-GPU simulation).

    schedule = "greedy"   : the reference's synchronous driver (one agent per round, argmax of the per-agent
                            gradient norm; examples/MultiRobotExample.cpp:229-334) -- parity mode;
             = "coloured" : all agents of one colour class per round (concurrent, same RBCD semantics);
             = "parallel" : every agent every round on the neighbours' previous poses.
    Per round: pack public poses -> ONE all-gather -> device-side G rebuild -> local optimise ->
    3-scalar all-gather for the central cost / gradient norm / selection.
    """

    def __init__(self, edges: EdgeSet, n: int, k: int, r: int = 5, algorithm: int = ROPTALG.RTR,
                 preconditioner: int = capi.PRECOND_SPARSE_EXACT, schedule: str = "greedy",
                 owner: Optional[np.ndarray] = None, X_init: Optional[np.ndarray] = None,
                 rank: Optional[int] = None, world: Optional[int] = None, device: int = 0, dist=None,
                 acceleration: bool = False, restart_interval: int = 30, concurrent: Optional[bool] = None):
        """concurrent: the active agents of a round that share a GPU step side by side (each as one thread-block cluster on
        its own stream, one C call per round: dpgo_agents_round_async) instead of one after the other as full-grid kernels.
        None = automatic: on when some round has >= 2 active agents on a rank (greedy / coloured schedules, no
        acceleration).  The iterates of the two modes agree to rounding, not bitwise (different reduction trees), so
        comparisons across world sizes must pin the mode."""
        import torch
        self.acceleration, self.restart_interval = bool(acceleration), int(restart_interval)
        self.torch = torch
        self.dist = dist
        self.k, self.n, self.r, self.d = k, n, r, edges.d
        self.rank = rank
        self.distributed = dist is not None and world is not None and world > 1
        self.world = world if self.distributed else 1
        if self.distributed:
            assert k % world == 0, "agents must divide evenly over the ranks (contiguous blocks of k/world agents)"
        self.owner = contiguous_owner(n, k) if owner is None else np.asarray(owner, dtype=np.int64)
        parts, counts, glob = partition_edges(edges, self.owner, k)
        self.counts, self.glob = counts, glob
        self.plan = ExchangePlan([p[2] for p in parts], k)
        self.colour = self.plan.colouring()
        self.ncolours = max(self.colour) + 1
        self.schedule = schedule
        if X_init is None:
            X_init = pg.fixedStiefelVariable(self.d, r) @ pg.chordalInitialization(self.d, n, edges)
        per_rank = k // self.world
        self.local_ids = list(range(rank * per_rank, (rank + 1) * per_rank)) if self.distributed else list(range(k))
        if concurrent is None:
            concurrent = auto_concurrent(self.colour, k, self.world, schedule, self.acceleration)
        if concurrent and (self.acceleration or schedule == "parallel"):
            raise ValueError("concurrent rounds are implemented for the greedy and coloured schedules without acceleration")
        self.concurrent = bool(concurrent)
        self.agents: Dict[int, PGOAgent] = {}
        dh = self.d + 1
        self.dev = torch.device("cuda", device)
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        for a in self.local_ids:
            prm = PGOAgentParameters(self.d, r, k, algorithm=algorithm, preconditioner=preconditioner, device=device,
                                     cluster=self.concurrent)
            ag = PGOAgent(a, prm)
            ag.mState = PGOAgentState.WAIT_FOR_DATA
            ag.YLift = None
            ag.setPoseGraph(*parts[a], TInit=np.zeros((self.d, dh * int(counts[a]))), n=int(counts[a]))
            cols = (glob[a][:, None] * dh + np.arange(dh)[None, :]).ravel()
            ag.setX(X_init[:, cols])
            ag.mProblem.set_stream(stream)
            ag.mProblem.upload_X(ag.X)
            ag.attach_exchange(self.plan)
            ag.opt = QuadraticOptimizer(ag.mProblem)
            ag.opt.setAlgorithm(algorithm)
            ag.opt.setTrustRegionTolerance(1e-2)
            ag.opt.setTrustRegionIterations(1)
            ag.opt.setTrustRegionMaxInnerIterations(10)
            ag.opt.setTrustRegionInitialRadius(100)
            ag.opt.setPreconditioner(preconditioner)
            self.agents[a] = ag
        ts = r * dh
        self.slot_elems = self.plan.pmax * ts
        self.gathered = torch.zeros(k * self.slot_elems, dtype=torch.float64, device=self.dev)
        # agents are laid out in agent order in the gathered buffer; a rank owns a contiguous run of them, so its
        # send buffer is one contiguous piece (rank-major all-gather order == agent order)
        if self.distributed:
            self.send_all = torch.zeros(per_rank * self.slot_elems, dtype=torch.float64, device=self.dev)
            base = self.local_ids[0]
            self.send = {a: self.send_all[(a - base) * self.slot_elems:(a - base + 1) * self.slot_elems]
                         for a in self.local_ids}
        else:
            self.send_all = None
            self.send = {a: self.gathered[a * self.slot_elems:(a + 1) * self.slot_elems] for a in self.local_ids}
        if self.acceleration:
            # Nesterov-accelerated RBCD (ref src/PGOAgent.cpp:685-695,1040-1091): auxiliary iterates resident per agent,
            # their public tiles in a second gathered buffer (ref getAuxSharedPoseDict / updateAuxNeighborPoses)
            self.gathered_aux = torch.zeros_like(self.gathered)
            if self.distributed:
                self.send_aux_all = torch.zeros_like(self.send_all)
                base = self.local_ids[0]
                self.send_aux = {a: self.send_aux_all[(a - base) * self.slot_elems:(a - base + 1) * self.slot_elems]
                                 for a in self.local_ids}
            else:
                self.send_aux_all = None
                self.send_aux = {a: self.gathered_aux[a * self.slot_elems:(a + 1) * self.slot_elems] for a in self.local_ids}
            self.acc = {a: dict(gamma=0.0, alpha=0.0, it=0) for a in self.local_ids}
            for a in self.local_ids:
                capi.check(self.agents[a].mProblem._lib.dpgo_agent_accel_init(self.agents[a].mProblem._h))
        self.stats_local = torch.zeros(4 * len(self.local_ids), dtype=torch.float64, device=self.dev)
        self.stats_all = torch.zeros(4 * k, dtype=torch.float64, device=self.dev)
        self.selected = [0]
        self.round = 0
        self._main_stream = stream if stream else 1          # 0 = torch's legacy default stream = cudaStreamLegacy (handle 1)
        self._gathered_current = False       # concurrent mode: `gathered` holds every agent's current public tiles

    # -- the exchange: ONE all-gather of the padded public-pose tiles --------------------------------
    def exchange(self, build: bool = True) -> None:
        for a in self.local_ids:
            self.agents[a].pack_public(self.send[a].data_ptr())
        if self.distributed:
            self.dist.all_gather_into_tensor(self.gathered, self.send_all)
        if build:
            for a in self.local_ids:
                self.agents[a].build_G(self.gathered.data_ptr(), self.k * self.plan.pmax)

    def _round_concurrent(self, active: List[int], publish: bool = True) -> None:
        """G rebuild -> RTR step -> pack for every active local agent, each on its own stream, with one C call; then the
        all-gather that publishes the new public tiles.  Asynchronous (no host synchronisation)."""
        mine = [a for a in self.local_ids if a in active]
        lib = self.agents[self.local_ids[0]].mProblem._lib
        if mine:
            hs = (C.c_void_p * len(mine))(*[self.agents[a].mProblem._h for a in mine])
            sd = (C.c_void_p * len(mine))(*[C.c_void_p(self.send[a].data_ptr()) for a in mine])
            capi.check(lib.dpgo_agents_round_async(hs, len(mine), C.byref(self.agents[mine[0]].opt._p),
                                                   C.c_void_p(self.gathered.data_ptr()), self.k * self.plan.pmax, sd,
                                                   C.c_void_p(self._main_stream), 0))
        if self.distributed and publish:
            self.dist.all_gather_into_tensor(self.gathered, self.send_all)

    def _active(self) -> List[int]:
        if self.schedule == "greedy":
            return list(self.selected)
        if self.schedule == "coloured":
            c = self.round % self.ncolours
            return [a for a in range(self.k) if self.colour[a] == c]
        return list(range(self.k))

    def evaluate(self) -> Tuple[float, float, np.ndarray]:
        """Central cost 2f, |grad|, per-agent gradient norms from per-agent (quad, lin, |g|^2)."""
        vals = np.zeros((self.k, 4))
        for a in self.local_ids:
            res = self.agents[a].opt.problem_stats()
            vals[a] = res
        if self.distributed:
            t = self.torch
            mine = np.ascontiguousarray(vals[self.local_ids[0]:self.local_ids[-1] + 1]).ravel()
            self.stats_local.copy_(t.from_numpy(mine))
            self.dist.all_gather_into_tensor(self.stats_all, self.stats_local)
            vals = self.stats_all.cpu().numpy().reshape(self.k, 4)
        cost = float(np.sum(vals[:, 0] + vals[:, 1]))          # 2 f_central = sum(<XQ,X> + <X,G>)
        gn2 = vals[:, 2]
        return cost, float(np.sqrt(np.sum(gn2))), np.sqrt(gn2)

    def _step_accelerated(self) -> List[int]:
        """One round of the accelerated schedule, every update on the device.  Order as in the reference's driver
        (examples/MultiRobotExample.cpp:236-279): every agent advances gamma / alpha / Y; the idle agents finish their
        iterate(false) (X = Y, V, restart); public tiles of X and of Y are exchanged; the active agents step from Y
        with G built from the neighbours' auxiliary poses."""
        active = self._active()
        N = float(self.k)
        lib = self.agents[self.local_ids[0]].mProblem._lib

        def restart_due(a):
            return (self.acc[a]["it"] + 1) % self.restart_interval == 0

        def finish_restart(a):
            capi.check(lib.dpgo_agent_accel_restart_end(self.agents[a].mProblem._h))
            self.acc[a]["gamma"] = self.acc[a]["alpha"] = 0.0

        for a in self.local_ids:
            st, h = self.acc[a], self.agents[a].mProblem._h
            st["it"] += 1
            st["gamma"] = (1 + np.sqrt(1 + 4 * N * N * st["gamma"] ** 2)) / (2 * N)
            st["alpha"] = 1.0 / (st["gamma"] * N)
            capi.check(lib.dpgo_agent_accel_begin(h, st["alpha"]))
            if a not in active:
                capi.check(lib.dpgo_agent_accel_end(h, st["gamma"], 0))
                if restart_due(a):                         # ref :1040-1052 with doOptimization == false: X = XPrev
                    capi.check(lib.dpgo_agent_accel_restart_begin(h))
                    finish_restart(a)
        # exchange X tiles and Y tiles
        for a in self.local_ids:
            self.agents[a].pack_public(self.send[a].data_ptr())
            capi.check(lib.dpgo_agent_pack_public_aux(self.agents[a].mProblem._h, C.c_void_p(self.send_aux[a].data_ptr())))
        if self.distributed:
            self.dist.all_gather_into_tensor(self.gathered, self.send_all)
            self.dist.all_gather_into_tensor(self.gathered_aux, self.send_aux_all)
        for a in self.local_ids:
            if a not in active:
                continue
            ag, st = self.agents[a], self.acc[a]
            h = ag.mProblem._h
            ag.build_G(self.gathered_aux.data_ptr(), self.k * self.plan.pmax)
            capi.check(lib.dpgo_optimize_resident_from_aux_async(h, C.byref(ag.opt.params())))
            capi.check(lib.dpgo_agent_accel_end(h, st["gamma"], 1))
            if restart_due(a):                             # X = XPrev, one plain step on the neighbours' X, V = Y = X
                capi.check(lib.dpgo_agent_accel_restart_begin(h))
                ag.build_G(self.gathered.data_ptr(), self.k * self.plan.pmax)
                ag.opt.optimize_resident_async()
                finish_restart(a)
        for a in self.local_ids:
            if a in active:
                self.agents[a].lastResult = self.agents[a].opt.fetch_result()
                self.agents[a].mIterationNumber += 1
        return active

    def step(self, evaluate: bool = True) -> Optional[RoundStats]:
        """One round: exchange, active agents optimise, (optionally) exchange again + evaluate + select."""
        if self.acceleration:
            active = self._step_accelerated()
            self.round += 1
            if not evaluate:
                return None
            self.exchange()
            cost, gn, per_agent = self.evaluate()
            if self.schedule == "greedy" and self.plan.tables[self.selected[0]]["neighbors"]:
                self.selected = [int(np.argmax(per_agent))]
            return RoundStats(cost, gn, active)
        active = self._active()
        if self.concurrent:
            if not self._gathered_current:
                self.exchange(build=False)
                self._gathered_current = True
            self._round_concurrent(active)
            for a in self.local_ids:
                if a in active:
                    self.agents[a].mIterationNumber += 1
        else:
            self.exchange()
            for a in self.local_ids:
                if a in active:
                    self.agents[a].opt.optimize_resident_async()
            for a in self.local_ids:
                if a in active:
                    if evaluate:
                        self.agents[a].lastResult = self.agents[a].opt.fetch_result()
                    self.agents[a].mIterationNumber += 1
        self.round += 1
        if not evaluate:
            return None
        self.exchange()                                   # fresh neighbour poses for the central gradient
        cost, gn, per_agent = self.evaluate()
        if self.schedule == "greedy":
            cur = self.selected[0]
            if self.plan.tables[cur]["neighbors"]:
                self.selected = [int(np.argmax(per_agent))]           # ref :308-325
        return RoundStats(cost, gn, active)

    # -- one round with host buffers in and out (the public host-level call of the runner) ---------------------------
    def _host_buffers(self):
        if not hasattr(self, "_hx"):
            torch = self.torch
            self._hx, self._hx_keep = {}, {}
            for a in self.local_ids:
                ag = self.agents[a]
                t = torch.empty(((self.d + 1) * ag.n, self.r), dtype=torch.float64).pin_memory()
                self._hx_keep[a] = t
                self._hx[a] = t.numpy().T                   # (r, N) Fortran-ordered view of the pinned buffer
                self._hx[a][...] = ag.X
        return self._hx

    def step_host(self) -> None:
        """One round with every iterate crossing the host boundary: per local agent X is uploaded from pinned host
        memory, the public poses are exchanged (pack -> one all-gather -> G rebuild, on the device), the active agents
        optimise, and their iterates are read back to the host.  ag.X (host) is the state between rounds."""
        hx = self._host_buffers()
        for a in self.local_ids:
            if hx[a] is not self.agents[a].X:
                hx[a][...] = self.agents[a].X
            if not self.concurrent:
                self.agents[a].mProblem.upload_X_async(hx[a])
        active = self._active()
        if self.concurrent:
            # one call per direction for the host boundary (uploads + packs / downloads, replayed as CUDA graphs), one
            # for the round, one synchronisation
            mine = [a for a in self.local_ids if a in active]
            lib = self.agents[self.local_ids[0]].mProblem._lib
            if not hasattr(self, "_io"):
                self._io = {}
            def arrays(ids, with_send):
                key = (tuple(ids), with_send)
                if key not in self._io:
                    hs = (C.c_void_p * len(ids))(*[self.agents[a].mProblem._h for a in ids])
                    hp = (C.c_void_p * len(ids))(*[C.c_void_p(self._hx_keep[a].data_ptr()) for a in ids])
                    sd = (C.c_void_p * len(ids))(*[C.c_void_p(self.send[a].data_ptr()) for a in ids]) if with_send else None
                    self._io[key] = (hs, hp, sd)
                return self._io[key]
            hs, hp, sd = arrays(self.local_ids, True)
            capi.check(lib.dpgo_agents_host_io_async(hs, len(self.local_ids), hp, sd, 0, C.c_void_p(self._main_stream)))
            if self.distributed:
                self.dist.all_gather_into_tensor(self.gathered, self.send_all)
            self._round_concurrent(active, publish=False)    # the next host round re-publishes every agent's tiles
            self._gathered_current = False
            if mine:
                hs, hp, _ = arrays(mine, False)
                capi.check(lib.dpgo_agents_host_io_async(hs, len(mine), hp, None, 1, C.c_void_p(self._main_stream)))
                self.agents[mine[0]].mProblem.sync()
            for a in mine:
                self.agents[a].X = hx[a]
                self.agents[a].mIterationNumber += 1
            self.round += 1
            return
        else:
            self.exchange()
            for a in self.local_ids:
                if a in active:
                    self.agents[a].opt.optimize_resident_async()
                    self.agents[a].mProblem.download_X_async(hx[a])
        for a in self.local_ids:
            if a in active:
                self.agents[a].mProblem.sync()
                self.agents[a].X = hx[a]
                self.agents[a].mIterationNumber += 1
        self.round += 1

    # -- the same round through the HOST-level interface (reference protocol: host matrices in and out) -------
    def step_host_dict(self) -> None:
        """One round with every iterate crossing the host boundary, as a user of the reference API would drive it:
        getSharedPoseDict -> (all-gather of the host-packed public poses) -> updateNeighborPoses -> iterate(), i.e.
        per active agent H2D of X and G, one persistent kernel, D2H of X.  Used for the end-to-end number."""
        torch = self.torch
        dh, ts = self.d + 1, self.r * (self.d + 1)
        if not hasattr(self, "_host_send"):
            self._host_send = torch.zeros(len(self.local_ids) * self.slot_elems, dtype=torch.float64).pin_memory()
            self._host_gath = torch.zeros(self.k * self.slot_elems, dtype=torch.float64).pin_memory()
        hs = self._host_send.numpy()
        for li, a in enumerate(self.local_ids):
            ag = self.agents[a]
            base = li * self.slot_elems
            for s, q in enumerate(self.plan.public[a]):
                hs[base + s * ts: base + (s + 1) * ts] = ag.X[:, q * dh:(q + 1) * dh].ravel(order="F")
        if self.distributed:
            self.send_all.copy_(self._host_send, non_blocking=True)
            self.dist.all_gather_into_tensor(self.gathered, self.send_all)
            self._host_gath.copy_(self.gathered, non_blocking=False)
            hg = self._host_gath.numpy()
        else:
            hg = hs
        active = self._active()
        for a in self.local_ids:
            if a not in active:
                continue
            ag = self.agents[a]
            for b in self.plan.tables[a]["neighbors"]:
                poses = {}
                for s, q in enumerate(self.plan.public[b]):
                    off = (b * self.plan.pmax + s) * ts
                    poses[(b, int(q))] = hg[off:off + ts].reshape(self.r, dh, order="F")
                ag.updateNeighborPoses(b, poses)
            ag.iterate(True)
        self.round += 1

    def host_bytes_per_step(self):
        """(h2d, d2h) bytes this rank moves per step_host round: every local agent's X in, the active agents' X out."""
        vb = [self.r * (self.d + 1) * int(self.counts[a]) * 8 for a in self.local_ids]
        nact = max(1, len(self.local_ids) // max(self.ncolours, 1)) if self.schedule == "coloured" else len(self.local_ids)
        return int(sum(vb)), int(sum(sorted(vb)[-nact:]))

    def assemble(self) -> np.ndarray:
        """Gather the full iterate on the host (all local agents; distributed: rank-local block only)."""
        dh = self.d + 1
        X = np.zeros((self.r, dh * self.n))
        for a in self.local_ids:
            cols = (self.glob[a][:, None] * dh + np.arange(dh)[None, :]).ravel()
            X[:, cols] = self.agents[a].mProblem.download_X()
        return X
