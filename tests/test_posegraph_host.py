"""Host-side data-format code of the product (reader, Laplacian blocks, initial guesses) vs the oracle."""
import os

import numpy as np
import pytest

from dpo_b200 import posegraph as pg
from oracle import dpgo_oracle as orc


@pytest.mark.parametrize("ds", ["tinyGrid3D", "smallGrid3D", "CSAIL", "sphere2500"])
def test_reader_and_laplacian(ds, data_dir):
    path = os.path.join(data_dir, ds + ".g2o")
    e, n = pg.read_g2o_file(path)
    m, n2 = orc.read_g2o(path)
    assert n == n2 and len(e) == len(m) and e.d == m.d
    assert np.array_equal(e.p1, m.p1) and np.array_equal(e.p2, m.p2)
    assert np.abs(e.R - m.R).max() <= 1e-15 and np.abs(e.t - m.t).max() == 0
    assert np.allclose(e.kappa, m.kappa, rtol=1e-14) and np.allclose(e.tau, m.tau, rtol=1e-14)
    Q = pg.constructConnectionLaplacianSE(e, n)
    Qo = orc.construct_connection_laplacian(m, n)
    assert abs(Q - Qo).max() <= 1e-12 * abs(Qo).max()
    assert abs(Q - Q.T).max() <= 1e-12 * abs(Qo).max()


@pytest.mark.parametrize("ds", ["tinyGrid3D", "smallGrid3D", "CSAIL"])
def test_initialisation(ds, data_dir):
    path = os.path.join(data_dir, ds + ".g2o")
    e, n = pg.read_g2o_file(path)
    m, _ = orc.read_g2o(path)
    T = pg.chordalInitialization(e.d, n, e)
    To = orc.chordal_initialization(m, n)
    assert np.abs(T - To).max() <= 1e-8 * max(1.0, np.abs(To).max())
    odo = np.where(e.p1 + 1 == e.p2)[0]
    if len(odo) == n - 1:
        assert np.abs(pg.odometryInitialization(e.d, n, e.take(odo)) -
                      orc.odometry_initialization(m.subset(odo), n)).max() <= 1e-10
    Y = pg.fixedStiefelVariable(e.d, 5)
    assert np.abs(Y.T @ Y - np.eye(e.d)).max() <= 1e-14                      # ref tests/testUtils.cpp:12-19
    assert np.array_equal(Y, pg.fixedStiefelVariable(e.d, 5))                # ref :21-25 (repeatable)


def test_synthetic_grid():
    e, n, Tgt = pg.synthetic_grid_graph(10, 10, 4, edges_per_pose=4.0, seed=0)
    assert n == 400 and len(e) == 1600
    assert len({(int(a), int(b)) for a, b in zip(e.p1, e.p2)}) == 1600
    assert np.all(e.p1 < e.p2)
    # the boustrophedon chain is present
    assert {(k, k + 1) for k in range(n - 1)} <= {(int(a), int(b)) for a, b in zip(e.p1, e.p2)}
    # measurements are close to the ground truth => small cost at ground truth relative to the scale
    p = orc.QuadraticProblem(n, 3, 3)
    m = orc.Measurements(3, e.r1, e.r2, e.p1, e.p2, e.R, e.t, e.kappa, e.tau, e.weight)
    p.set_Q(orc.construct_connection_laplacian(m, n))
    assert 2 * p.f(Tgt) / len(e) < 200 * 3 * 0.05 ** 2 * 4 + 100 * 3 * 0.1 ** 2 * 2


def test_block_partition_of_the_synthetic_lattice():
    """BASELINE config 5 hygiene: 8 agents as 2 x 2 x 2 blocks of the lattice instead of contiguous id ranges (slabs in
    which every pose is public): far fewer public poses, balanced blocks, a valid owner map for the exchange plan."""
    from dpo_b200.agent import ExchangePlan, contiguous_owner, partition_edges
    nx, ny, nz, k = 20, 20, 8, 8
    e, n, _ = pg.synthetic_grid_graph(nx, ny, nz, edges_per_pose=4.0, seed=0)
    c = pg.grid_lattice_coords(nx, ny, nz)
    assert c.shape == (n, 3) and np.abs(np.diff(c, axis=0)).sum(axis=1).max() == 1     # the id path visits lattice neighbours
    owner = pg.grid_block_owner(nx, ny, nz, k)
    assert owner.shape == (n,) and set(owner.tolist()) == set(range(k))
    assert np.bincount(owner).min() == np.bincount(owner).max() == n // k
    pub = {}
    for name, own in (("blocks", owner), ("ranges", contiguous_owner(n, k))):
        parts, counts, glob = partition_edges(e, own, k)
        plan = ExchangePlan([p[2] for p in parts], k)
        pub[name] = sum(len(q) for q in plan.public)
    assert pub["ranges"] == n                       # slabs: every pose is public
    assert pub["blocks"] < 0.7 * pub["ranges"]     # (random closures reach 3 lattice steps across a block face)
    with pytest.raises(ValueError):
        pg.grid_block_owner(2, 2, 2, 16)
