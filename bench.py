#!/usr/bin/env python
"""bench.py -- RTR iterations/sec on sphere2500 (BASELINE.json metric) + the Q.X SpMV roofline.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --steps K --warmup W  (CPU restatement of the reference path)

A "step" is one QuadraticOptimizer::optimize() call with the constants PGOAgent::updateX uses
(ref src/PGOAgent.cpp:1131-1137: RTR, tol 1e-2, 1 outer iteration, <= 10 tCG iterations, radius 100)
and the reference's preconditioner operator (Q + 0.1 I)^-1.  Steps walk the optimisation trajectory from
the chordal initialisation; the iterate is reset to the initial point every CYCLE steps, before the early-exit
tolerance is reached, so every timed step does full work.

 N = 1 : sphere2500 as ONE agent (r = 5).   value = steps/s with the iterate resident in HBM;
         e2e   = the same through QuadraticOptimizer.optimize() with pinned host buffers (H2D + D2H per step).
 N > 1 : sphere2500 split contiguously into N agents, one per GPU; every round = pack public poses ->
         one NCCL all-gather -> device-side G rebuild -> the agents of the round's colour class take one RTR
         step.  value = agent steps/s summed over ranks ("strong": the graph is fixed, the split grows).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATASET = "sphere2500"
RANK_R = 5
CYCLE = 6            # steps per trajectory before resetting to the initial point (see docstring)
METRIC = "rtr_iters_per_sec_sphere2500"
UNIT = "iter/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            pk = json.load(open(path))
            return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str):
    """DRAM read+write bytes per launch from the committed ncu capture (profiles/traffic.json), else None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places that may execute oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps: int, warmup: int, sample_desc_only: bool = False):
    """Time the CPU restatement of the reference path on the SAME workload (sphere2500, one agent, r=5,
    RTR with updateX constants, exact (Q+0.1I)^-1 preconditioner).  Returns (steps_per_sec, info)."""
    try:
        from oracle import cpu_port                      # C++ restatement (g++ -O3 -march=native), if built
        have_port = cpu_port.available()
    except Exception:
        cpu_port, have_port = None, False
    from oracle import dpgo_oracle as orc
    meas, n = orc.read_g2o(os.path.join(ROOT, "data", DATASET + ".g2o"))
    Q = orc.construct_connection_laplacian(meas, n)
    X0 = orc.fixed_stiefel_variable(meas.d, RANK_R) @ orc.chordal_initialization(meas, n)
    if have_port:
        # single thread = the reference's default (ENABLE_OPENMP OFF; Eigen's dense*sparse product is serial);
        # a threaded variant (products over columns) is tried as well and the faster one is reported
        runner = cpu_port.Runner(Q, n, meas.d, RANK_R, threads=1)
        best = 1
        try:
            t0 = time.perf_counter(); runner.step(X0); runner.step(X0); t1 = time.perf_counter() - t0
            nthr = max(2, min(8, os.cpu_count() or 2))
            alt = cpu_port.Runner(Q, n, meas.d, RANK_R, threads=nthr)
            t0 = time.perf_counter(); alt.step(X0); alt.step(X0); t2 = time.perf_counter() - t0
            if t2 < 0.9 * t1:
                runner, best = alt, nthr
        except Exception:
            pass
        kind, cores = "port", best
        label = ("C++ restatement of the reference path (scalar-CSR X*Q, min-degree sparse LDL^T solves of Q+0.1I, "
                 "ROPTLIB RTR/tCG restated; 10+j products per call), g++ -O3")
    else:
        prob = orc.QuadraticProblem(n, meas.d, RANK_R)
        prob.set_Q(Q)

        class _PyRunner:
            def step(self, X):
                oo = orc.QuadraticOptimizer(prob)
                oo.tr_tolerance, oo.tr_iterations, oo.tr_max_inner, oo.tr_initial_radius = 1e-2, 1, 10, 100.0
                return oo.optimize(X)
        runner = _PyRunner()
        kind, cores, label = "port", 1, "NumPy/SciPy oracle (scipy.sparse SpMM + SuperLU solves), single thread"
    X = X0
    for i in range(warmup):
        X = runner.step(X) if (i + 1) % CYCLE else runner.step(X0)
    X = X0
    t0 = time.perf_counter()
    for i in range(steps):
        if i % CYCLE == 0:
            X = X0
        X = runner.step(X)
    dt = time.perf_counter() - t0
    info = {"value": steps / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{steps} consecutive optimize() steps of the bench workload ({DATASET}, 1 agent, r={RANK_R}, "
                      f"reset every {CYCLE}); {label}", "host_cpus": os.cpu_count()}
    return steps / dt, dt, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    val, dt, info = cpu_reference_steps(steps, min(args.warmup, 2))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "data/sphere2500.g2o (public dataset), chordal init",
            "config": workload_config(1, "host cores only"), "cpu_baseline": info,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(n_agents: int, schedule: str):
    return {"workload": f"{DATASET}.g2o SE(3), 2500 poses / 4949 edges, {n_agents} agent(s), r={RANK_R}, "
                        f"one RTR step per optimize() (tol 1e-2, <=10 tCG, radius 100), preconditioner (Q+0.1I)^-1",
            "agents": n_agents, "schedule": schedule, "cycle_reset": CYCLE,
            "l2_policy": "sphere2500 working set (Q 1.6 MB + vectors) is L2-resident by nature of the named dataset; "
                         "the dense preconditioner (N^2*8 = 800 MB at 1 agent) and the SpMV-roofline inputs exceed L2"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def spmv_roofline(torch, dp, pg, peak_gbs, peak_src, dims=(100, 100, 40), reps=20):
    """Q.X product alone on a synthetic grid far larger than L2 (SURVEY 8d config 5 family)."""
    edges, n, _ = pg.synthetic_grid_graph(*dims, edges_per_pose=4.0, seed=0)
    prob = dp.QuadraticProblem(n, 3, RANK_R, preconditioners=(dp.PRECOND_BLOCK_JACOBI,))
    prob.setQ_blocks(*pg.connection_laplacian_blocks(edges))
    prob.set_stream(torch.cuda.current_stream().cuda_stream)
    X = torch.randn(RANK_R * 4 * n, dtype=torch.float64, device="cuda")
    out = torch.empty_like(X)
    for _ in range(3):
        prob.spmv_device(X.data_ptr(), out.data_ptr(), False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        prob.spmv_device(X.data_ptr(), out.data_ptr(), False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = prob.spmv_algorithmic_bytes(False)
    ach = nbytes / (ms * 1e-3) / 1e9
    res = {"kernel": "k_spmv_tma<5,4,192> (Out = X Q; bulk-TMA producer, DMMA consumers)", "workload": f"synthetic grid {dims[0]}x{dims[1]}x{dims[2]} = {n} poses, "
           f"{len(edges)} edges, r={RANK_R}, nb={prob.num_blocks()} blocks", "bound": "hbm", "achieved": ach,
           "peak": peak_gbs, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak_gbs,
           "algorithmic_bytes_per_launch": nbytes, "us_per_launch": ms * 1e3, "launches_timed": reps,
           "traffic": ncu_traffic("k_spmv_tma")}
    prob.close()
    del X, out
    torch.cuda.empty_cache()
    return res


def run_gpu_arm(args):
    import torch
    import dpo_b200 as dp
    from dpo_b200 import posegraph as pg
    from dpo_b200.agent import DistributedPGO

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    K, W = args.steps, max(args.warmup, 3)
    peak, peak_src = measured_peaks()

    edges, n = pg.read_g2o_file(os.path.join(ROOT, "data", DATASET + ".g2o"))
    d = edges.d
    X0 = pg.fixedStiefelVariable(d, RANK_R) @ pg.chordalInitialization(d, n, edges)
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "data/sphere2500.g2o (public benchmark pose graph shipped with the reference), chordal initialisation"}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        prob = dp.QuadraticProblem(n, d, RANK_R, device=local_rank,
                                   preconditioners=(dp.PRECOND_BLOCK_JACOBI, dp.PRECOND_SPARSE_EXACT, dp.PRECOND_DENSE_EXACT))
        prob.setQ_blocks(*pg.connection_laplacian_blocks(edges))
        prob.set_stream(torch.cuda.current_stream().cuda_stream)
        opt = dp.QuadraticOptimizer(prob)
        opt.setTrustRegionTolerance(1e-2)
        opt.setTrustRegionIterations(1)
        opt.setTrustRegionMaxInnerIterations(10)
        opt.setTrustRegionInitialRadius(100)
        PRE = dp.PRECOND_DENSE_EXACT if args.precond == "dense" else dp.PRECOND_SPARSE_EXACT
        opt.setPreconditioner(PRE)
        X0d = torch.from_numpy(np.asfortranarray(X0).ravel(order="F").copy()).to(dev)

        def resident_steps(count, collect=None):
            for i in range(count):
                if i % CYCLE == 0:
                    prob.copy_X_from_device(X0d.data_ptr())
                opt.optimize_resident_async()
                if collect is not None:
                    collect.append(opt.fetch_result())

        # ---- correctness trail of one cycle (also warms everything) ----
        trail = []
        resident_steps(CYCLE, trail)
        resident_steps(W)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        resident_steps(K)
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        # ---- end to end: public API, pinned host buffers, H2D + D2H inside the timed region ----
        def pinned():                       # (r, N) Fortran-ordered view of a pinned (N, r) tensor
            t = torch.empty(((d + 1) * n, RANK_R), dtype=torch.float64).pin_memory()
            return t, t.numpy().T
        keep0, X0f = pinned()
        X0f[...] = X0
        keepA, bufA = pinned()
        keepB, bufB = pinned()
        bufs = [bufA, bufB]
        host = X0f
        for i in range(W):
            host = opt.optimize(X0f if i % CYCLE == 0 else host, out=bufs[i & 1])
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            host = opt.optimize(X0f if i % CYCLE == 0 else host, out=bufs[i & 1])
        barrier()
        e2e_dt = time.perf_counter() - t0
        clocks = sampler.stop()
        last = opt.getOptResult()
        vec_bytes = RANK_R * (d + 1) * n * 8
        # ---- roofline of the dominant kernel of the step: the persistent k_optimize launch ----
        spmv_b = prob.spmv_algorithmic_bytes(True)
        N = (d + 1) * n
        pre_b = prob.precond_algorithmic_bytes(PRE)      # bytes of the operator's blocks one application streams
        pre_full = N * N * 8 + 2 * vec_bytes                                # the full dense operator
        per_step, per_step_full, flops = [], [], []
        for rs in trail:
            per_step.append(rs.spmv_passes * spmv_b + rs.precond_applies * pre_b)
            per_step_full.append(rs.spmv_passes * spmv_b + rs.precond_applies * pre_full)
            flops.append(rs.precond_applies * 2.0 * RANK_R * N * N)
        alg_bytes = float(np.mean(per_step))
        alg_full = float(np.mean(per_step_full))
        ms_step = ms_total / K
        ach = alg_bytes / (ms_step * 1e-3) / 1e9
        line.update({
            "value": K / (ms_total * 1e-3), "ms_per_step": ms_step,
            "config": workload_config(1, "single agent"),
            "e2e": {"value": K / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": vec_bytes,
                    "d2h_bytes_per_step": vec_bytes + 96, "ms_per_step": 1e3 * e2e_dt / K,
                    "api": "dpo_b200.QuadraticOptimizer.optimize(Y) -> dpgo_optimize (host buffers)"},
            "gpu_launches": K,
            "clocks": clocks,
            "roofline": {"kernel": "k_optimize<5,4> (one persistent launch per step)", "bound": "hbm", "achieved": ach,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                         "algorithmic_bytes_per_launch": alg_bytes, "traffic": ncu_traffic("k_optimize"),
                         "traffic_note": "ncu capture of one launch (a 10-tCG step: 11 preconditioner applications)",
                         "note": "bytes = spmv_passes*(132 nb + 4(n+1) + 96 r n) + precond_applies*(P + 16 r N), averaged "
                                 "over the cycle; P = unique bytes of the symmetric dense (Q+0.1I)^-1 = 4 N (N+8) when the "
                                 "upper-triangle kernel is planned, else 8 N^2; the dense stream dominates",
                         "full_matrix": {"algorithmic_bytes_per_launch": alg_full,
                                         "achieved": alg_full / (ms_step * 1e-3) / 1e9,
                                         "frac": alg_full / (ms_step * 1e-3) / 1e9 / peak,
                                         "note": "same time against the bytes of the full N x N operator (what a "
                                                 "non-symmetric apply would stream)"},
                         "fp64": {"useful_tflops": float(np.mean(flops)) / (ms_step * 1e-3) / 1e12,
                                  "note": "2 r N^2 flops per preconditioner application; the symmetric apply runs on "
                                          "DMMA m8n8k4 with 5 of 8 M rows used (scripts/dmma_peak.cu measures the pipe)"}},
            "trajectory": [{"f": rs.f_opt, "gradnorm": rs.gradnorm_opt, "tcg": rs.tcg_iterations,
                            "status": rs.tcg_status, "spmv_passes": rs.spmv_passes} for rs in trail],
        })
        # ---- extras: Jacobi throughput mode and the RGD step on the same workload ----
        extras = {}
        for name, setter in (("rtr_block_jacobi", lambda: (opt.setAlgorithm(dp.ROPTALG.RTR), opt.setPreconditioner(dp.PRECOND_BLOCK_JACOBI))),
                             ("rgd", lambda: opt.setAlgorithm(dp.ROPTALG.RGD))):
            setter()
            resident_steps(W)
            barrier()
            e0.record()
            resident_steps(K)
            e1.record()
            barrier()
            extras[name + "_iters_per_sec"] = K / (e0.elapsed_time(e1) * 1e-3)
        line["extra"] = extras
        prob.close()
        if not args.no_spmv:
            line["roofline_spmv"] = spmv_roofline(torch, dp, pg, peak, peak_src)
        # ---- CPU restatement of the reference path on the host cores, bounded sample ----
        if not args.no_cpu:
            _, _, info = cpu_reference_steps(2 * CYCLE, 1)
            line["cpu_baseline"] = info
        print(json.dumps(line))
    else:
        # everything of the multi-GPU rounds (pack, NCCL all-gather, G rebuild, optimise) lives on one side stream so
        # that a whole cycle of rounds can be captured into a CUDA graph and replayed (launch-bound inner loop)
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            run = DistributedPGO(edges, n, world, r=RANK_R, schedule="coloured", X_init=X0, rank=rank, world=world,
                                 device=local_rank, dist=dist)
        ag = run.agents[rank]          # world == number of agents here: one agent per GPU
        dh = d + 1
        cols = (run.glob[rank][:, None] * dh + np.arange(dh)[None, :]).ravel()
        X0d = torch.from_numpy(np.asfortranarray(X0[:, cols]).ravel(order="F").copy()).to(dev)
        rounds_per_cycle = CYCLE * run.ncolours

        def run_rounds(count):
            """`count` RBCD rounds; the iterate is reset to the initial point at the start of every cycle."""
            mine = 0
            with torch.cuda.stream(side):
                for i in range(count):
                    c = i % rounds_per_cycle
                    if c == 0:
                        ag.mProblem.copy_X_from_device(X0d.data_ptr())
                    run.exchange()
                    if run.colour[rank] == c % run.ncolours:
                        ag.opt.optimize_resident_async()
                        mine += 1
            return mine

        # (capturing a cycle -- cooperative kernels + NCCL all-gathers -- into one CUDA graph was tried and hangs at
        #  replay on this stack, so the rounds are launched eagerly from the host)
        graph_note = "eager launches"
        run_rounds(max(W, 2 * rounds_per_cycle))
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(side):
            e0.record()
        my_steps = run_rounds(K)                     # exactly K timed rounds
        with torch.cuda.stream(side):
            e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1), float(my_steps)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total = float(tmax[0])
        total_steps = float(tsum[1])
        line["steps"] = K
        with torch.cuda.stream(side):
            st = run.step(evaluate=True)
        # ---- end to end: the same coloured rounds driven through the host-level API (host matrices in and out,
        #      public poses packed on the host, H2D / all-gather / D2H inside the timed region) ----
        ag.mProblem.sync()
        ag.X = np.array(X0[:, cols])
        run.round = 0
        KE = max(4, K // 4)
        with torch.cuda.stream(side):
            for _ in range(2 * run.ncolours):
                run.step_host()
        barrier()
        t0 = time.perf_counter()
        host_steps = 0
        with torch.cuda.stream(side):
            for i in range(KE):
                if i % rounds_per_cycle == 0:
                    ag.X = np.array(X0[:, cols])
                    run.round = 0
                if run.colour[rank] == run.round % run.ncolours:
                    host_steps += 1
                run.step_host()
        barrier()
        e2e_dt = time.perf_counter() - t0
        te = torch.tensor([e2e_dt, float(host_steps)], dtype=torch.float64, device=dev)
        te_max, te_sum = te.clone(), te.clone()
        dist.all_reduce(te_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(te_sum, op=dist.ReduceOp.SUM)
        h2d, d2h = run.host_bytes_per_step()
        if rank == 0:
            clocks = sampler.stop()
            pub_bytes = run.plan.pmax * RANK_R * dh * 8
            line.update({
                "value": total_steps / (ms_total * 1e-3), "ms_per_step": ms_total / K,
                "config": dict(workload_config(world, f"coloured RBCD ({run.ncolours} colours), one agent per GPU"),
                               rounds=K, agent_steps=int(total_steps), launch_mode=graph_note,
                               allgather_bytes_per_rank=pub_bytes, parallelism=f"agents{world}"),
                "rounds_per_sec": K / (ms_total * 1e-3),
                "e2e": {"value": float(te_sum[1]) / float(te_max[0]), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "rounds": KE,
                        "api": "PGOAgent.updateNeighborPoses + PGOAgent.iterate(True) with host matrices; public poses "
                               "packed on the host, H2D -> NCCL all-gather -> D2H every round"},
                "gpu_launches": int(K * 2 + my_steps), "clocks": clocks,
                "roofline": None, "final": {"cost": st.cost, "gradnorm": st.gradnorm},
            })
            print(json.dumps(line))
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-spmv", action="store_true", help="skip the synthetic SpMV roofline leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--precond", default="sparse", choices=["sparse", "dense"],
                    help="exact preconditioner implementation: nested-dissection block solve (default) or the dense inverse (A/B)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
