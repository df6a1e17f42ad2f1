"""Chordal initialisation on the GPU (dpgo_chordal_initialization: two Jacobi-preconditioned CG solves over the hot path's
block-CSR product kernel + SO(d) projection) against the oracle's sparse direct solves and the constants the reference
publishes (vis.ipynb:108746,108748: cost 2f and gradient norm at the chordal point, r = d)."""
import os

import numpy as np
import pytest

from oracle import dpgo_oracle as orc

pytestmark = pytest.mark.gpu

CHORDAL = {"sphere2500": (1971.17, 265.247), "smallGrid3D": (1561.38, 237.586), "torus3D": (24669.2, 320.591),
           "parking-garage": (1.41536, 2.3906), "CSAIL": (31.4848, 5.44293)}


@pytest.mark.parametrize("ds", ["tinyGrid3D", "smallGrid3D", "CSAIL", "sphere2500", "torus3D", "parking-garage"])
def test_chordal_gpu_matches_oracle_and_published_constants(ds, data_dir):
    from dpo_b200 import posegraph as pg
    edges, n = pg.read_g2o_file(os.path.join(data_dir, ds + ".g2o"))
    meas, _ = orc.read_g2o(os.path.join(data_dir, ds + ".g2o"))
    T, its = pg.chordalInitializationGPU(edges.d, n, edges, return_iterations=True)
    To = orc.chordal_initialization(meas, n)
    d = edges.d
    # rotations are exactly orthonormal with det +1; pose 0 is the gauge
    Rg = np.transpose(T.reshape(d, d + 1, n, order="F")[:, :d, :], (2, 0, 1))        # (n, d, d): R_p
    assert np.abs(np.einsum("nab,nac->nbc", Rg, Rg) - np.eye(d)[None]).max() <= 1e-13
    assert np.all(np.linalg.det(Rg) > 0.99)
    assert np.abs(T[:, :d] - np.eye(d)).max() <= 1e-14 and np.abs(T[:, d]).max() == 0.0
    scale = max(1.0, np.abs(To).max())
    assert np.abs(T - To).max() <= 1e-7 * scale, (its, np.abs(T - To).max())
    if ds in CHORDAL:
        p = orc.QuadraticProblem(n, d, d)
        p.set_Q(orc.construct_connection_laplacian(meas, n))
        cost, gn = CHORDAL[ds]
        assert abs(2 * p.f(T) - cost) <= 6e-6 * cost
        assert abs(p.rie_grad_norm(T) - gn) <= 6e-6 * gn


def test_chordal_gpu_large_synthetic_grid():
    """Size the host direct solves do not like: 64k poses / 256k edges; property check -- the chordal point of a graph with
    small noise is close to the ground truth and has a small cost."""
    from dpo_b200 import posegraph as pg
    edges, n, Tgt = pg.synthetic_grid_graph(40, 40, 40, edges_per_pose=4.0, seed=2)
    T, its = pg.chordalInitializationGPU(3, n, edges, return_iterations=True)
    Rg = np.transpose(T.reshape(3, 4, n, order="F")[:, :3, :], (2, 0, 1))            # (n, 3, 3)
    Rt = np.transpose(np.asarray(Tgt).reshape(3, 4, n, order="F")[:, :3, :], (2, 0, 1))
    Rrel = np.einsum("ba,nbc->nac", Rt[0], Rt)                                       # ground truth in the gauge of pose 0
    ang = np.arccos(np.clip((np.einsum("nab,nab->n", Rg, Rrel) - 1) / 2, -1, 1))
    # measurement noise is 0.05 rad per edge; the chordal point stays within a few noise levels of the ground truth
    assert np.median(ang) < 0.15 and np.max(ang) < 1.0 and its[0] > 0 and its[1] > 0
