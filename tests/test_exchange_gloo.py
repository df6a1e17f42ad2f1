"""N>1 host logic on CPU: world-size-2 gloo run of the boundary-pose exchange plan (dpo_b200.agent.ExchangePlan).

The device kernels (k_pack_tiles, k_build_G) cannot run without a GPU, so each rank emulates them with the SAME
index tables in NumPy: pack public tiles -> ONE all_gather (gloo) -> rebuild G from the gathered slots, and compares
with the oracle's dictionary-based constructGMatrix.  This pins the slot / padding / edge tables across processes."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ds, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from dpo_b200 import posegraph as pg
    from dpo_b200.agent import ExchangePlan, contiguous_owner, partition_edges
    from oracle import dpgo_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        edges, n = pg.read_g2o_file(os.path.join(ROOT, "data", ds + ".g2o"))
        d, r, dh = edges.d, 5, edges.d + 1
        owner = contiguous_owner(n, world)
        parts, counts, glob = partition_edges(edges, owner, world)
        plan = ExchangePlan([p[2] for p in parts], world)
        rng = np.random.default_rng(7)                     # same on every rank
        Xfull = rng.standard_normal((r, dh * n))
        cols = (glob[rank][:, None] * dh + np.arange(dh)[None, :]).ravel()
        Xloc = Xfull[:, cols]
        ts = r * dh
        # --- pack (what k_pack_tiles does): slot s <- tile public[rank][s], column-major tiles
        send = np.zeros(plan.pmax * ts)
        for s, q in enumerate(plan.public[rank]):
            send[s * ts:(s + 1) * ts] = Xloc[:, q * dh:(q + 1) * dh].ravel(order="F")
        gathered = torch.zeros(world * plan.pmax * ts, dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, torch.from_numpy(send))
        gathered = gathered.numpy()
        # --- rebuild G (what k_build_G does) from the tables
        tb = plan.tables[rank]
        G = np.zeros((r, dh * int(counts[rank])))
        for e in range(len(tb["local"])):
            Xn = gathered[tb["slot"][e] * ts:(tb["slot"][e] + 1) * ts].reshape(r, dh, order="F")
            T, om = tb["T"][e], tb["omega"][e]
            L = (Xn * om[None, :]) @ T.T if tb["outgoing"][e] else (Xn @ T) * om[None, :]
            p = tb["local"][e]
            G[:, p * dh:(p + 1) * dh] -= L
        # --- oracle: dictionary form with the true neighbour poses
        sh = parts[rank][2]
        m = orc.Measurements(d, sh.r1, sh.r2, sh.p1, sh.p2, sh.R, sh.t, sh.kappa, sh.tau, sh.weight)
        oa = orc.PGOAgent(rank, d, r)
        oa.n = int(counts[rank])
        oa.shared_lc = m
        oa.problem = orc.QuadraticProblem(oa.n, d, r)
        poses = {}
        for b in range(world):
            if b == rank:
                continue
            cb = (glob[b][:, None] * dh + np.arange(dh)[None, :]).ravel()
            Xb = Xfull[:, cb]
            for q in plan.public[b]:
                poses[(b, int(q))] = Xb[:, q * dh:(q + 1) * dh]
        assert oa.construct_G(poses)
        err = float(np.abs(G - oa.problem.G).max())
        ret[rank] = (err, int(plan.pmax), len(tb["local"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ds", ["smallGrid3D", "CSAIL"])
def test_exchange_plan_world2_gloo(ds):
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 300)
    mgr = mp.Manager()
    ret = mgr.dict()
    procs = [mp.get_context("spawn").Process(target=_worker, args=(rk, 2, port, ds, ret)) for rk in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert len(ret) == 2
    for rk in range(2):
        err, pmax, nedges = ret[rk]
        assert err <= 1e-12, (rk, err)
        assert pmax >= 1 and nedges >= 1
    assert ret[0][2] == ret[1][2]          # both agents see the same cut edges


def test_partition_and_colouring():
    sys.path.insert(0, ROOT)
    from dpo_b200 import posegraph as pg
    from dpo_b200.agent import ExchangePlan, contiguous_owner, partition_edges
    edges, n = pg.read_g2o_file(os.path.join(ROOT, "data", "smallGrid3D.g2o"))
    owner = contiguous_owner(n, 5)
    assert owner[0] == 0 and owner[-1] == 4 and np.all(np.diff(owner) >= 0)       # ref ex:95-109 contiguous blocks
    parts, counts, glob = partition_edges(edges, owner, 5)
    assert counts.sum() == n
    n_private = sum(len(p[0]) + len(p[1]) for p in parts)
    n_shared = sum(len(p[2]) for p in parts)
    assert n_private + n_shared // 2 == len(edges)                                 # shared edges appear in both agents
    plan = ExchangePlan([p[2] for p in parts], 5)
    col = plan.colouring()
    for a in range(5):
        for b in plan.tables[a]["neighbors"]:
            assert col[a] != col[b]
