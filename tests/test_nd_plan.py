"""CPU checks of the sparse exact preconditioner's host side (dpo_b200/csrc/nd_precond.cpp): nested dissection, macro
levels, block algebra and the static phase plan, through a host emulation of the plan exactly as the kernel interprets
it (dpgo_nd_debug_emulate: host only, verification only).  Reference operator: (Q + 0.1 I)^-1 by sparse LU
(ref src/QuadraticProblem.cpp:31-42,75-87)."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from dpo_b200 import _capi as capi
from dpo_b200 import posegraph as pg

INFO = ("levels", "nodes", "phases", "block_bytes", "bytes_per_apply", "max_own", "max_bnd", "nd_depth", "steps", "jobs",
        "epilogues", "max_ytiles", "max_slots")


def emulate(n, d, r, brow, bcol, blocks, V, grid=148, cuts=-1, leaf=0, shift=0.1):
    lib = capi.load_library()
    brow = np.ascontiguousarray(brow, dtype=np.int32)
    bcol = np.ascontiguousarray(bcol, dtype=np.int32)
    blocks = np.ascontiguousarray(blocks, dtype=np.float64)
    Vf = np.asfortranarray(V, dtype=np.float64)
    Z = np.asfortranarray(np.full(Vf.shape, np.nan))
    info = (C.c_int64 * 16)()
    capi.check(lib.dpgo_nd_debug_emulate(n, d, r, len(brow), capi.iptr(brow), capi.iptr(bcol), capi.dptr(blocks), shift, grid,
                                         cuts, leaf, capi.dptr(Vf), capi.dptr(Z), info))
    return Z, {k: int(info[i]) for i, k in enumerate(INFO)}


def dense_reference(n, dh, brow, bcol, blocks, V, shift=0.1):
    rows = (np.asarray(brow)[:, None, None] * dh + np.arange(dh)[None, :, None]) + np.zeros((1, 1, dh), dtype=np.int64)
    cols = (np.asarray(bcol)[:, None, None] * dh + np.arange(dh)[None, None, :]) + np.zeros((1, dh, 1), dtype=np.int64)
    Q = sp.csr_matrix((np.asarray(blocks).ravel(), (rows.ravel(), cols.ravel())), shape=(dh * n, dh * n))
    A = (Q + shift * sp.identity(dh * n)).tocsc()
    return spla.splu(A).solve(np.asarray(V).T).T


def relerr(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("ds,r,grid,cuts", [("tinyGrid3D", 3, 148, -1), ("smallGrid3D", 5, 148, -1), ("smallGrid3D", 5, 8, 2),
                                            ("CSAIL", 3, 148, -1), ("CSAIL", 2, 20, 3), ("sphere2500", 5, 148, -1),
                                            ("sphere2500", 5, 148, 1), ("sphere2500", 3, 37, 3), ("parking-garage", 5, 148, -1),
                                            ("torus3D", 5, 148, -1), ("sphere2500", 5, 148, 0),
                                            # grids of the cluster launch mode (<= 16 CTAs interpret the whole plan)
                                            ("smallGrid3D", 5, 2, -1), ("CSAIL", 3, 10, 2), ("parking-garage", 5, 16, 2)])
def test_emulated_plan_matches_sparse_lu(ds, r, grid, cuts, data_dir):
    edges, n = pg.read_g2o_file(os.path.join(data_dir, ds + ".g2o"))
    brow, bcol, blocks = pg.connection_laplacian_blocks(edges)
    dh = edges.d + 1
    V = np.random.default_rng(1).standard_normal((r, dh * n))
    Z, info = emulate(n, edges.d, r, brow, bcol, blocks, V, grid=grid, cuts=cuts)
    assert relerr(Z, dense_reference(n, dh, brow, bcol, blocks, V)) <= 1e-12
    assert info["phases"] == 2 * info["levels"] - 1
    if cuts >= 0 and info["nd_depth"] > cuts:
        assert info["levels"] == cuts + 1
    assert info["max_ytiles"] <= 600 and info["max_slots"] <= 240
    if ds == "sphere2500" and cuts < 0:
        # the point of the design: well below the dense inverse (800 MB) and the O(N^2) cap
        assert info["block_bytes"] < 80e6


def test_ragged_and_degenerate_graphs():
    rng = np.random.default_rng(3)

    def spd_blocks(n, edge_list, dh):
        """random connection-Laplacian-like SPD block matrix: sum over edges of B^T B with B = [M, -I]"""
        brow, bcol, blocks = [], [], []
        for (i, j) in edge_list:
            M = rng.standard_normal((dh, dh))
            brow += [i, j, i, j]
            bcol += [i, j, j, i]
            blocks += [M.T @ M, np.eye(dh), -M.T, -M]
        if not edge_list:
            brow, bcol, blocks = [0], [0], [np.zeros((dh, dh))]
        return np.array(brow), np.array(bcol), np.array(blocks)

    cases = {
        "single pose": (1, []),
        "two poses": (2, [(0, 1)]),
        "chain": (57, [(i, i + 1) for i in range(56)]),
        "two components + isolated pose": (41, [(i, i + 1) for i in range(19)] + [(i, i + 1) for i in range(20, 39)]),
        "dense clique (no separator)": (30, [(i, j) for i in range(30) for j in range(i + 1, 30)]),
        "star": (64, [(0, i) for i in range(1, 64)]),
        "random sparse": (300, [(int(a), int(b)) for a, b in rng.integers(0, 300, size=(700, 2)) if a != b]),
    }
    for name, (n, el) in cases.items():
        for d in (2, 3):
            dh = d + 1
            brow, bcol, blocks = spd_blocks(n, el, dh)
            for r in (d, 5):
                V = rng.standard_normal((r, dh * n))
                for grid, leaf in ((148, 0), (4, 3)):
                    Z, info = emulate(n, d, r, brow, bcol, blocks, V, grid=grid, leaf=leaf)
                    assert relerr(Z, dense_reference(n, dh, brow, bcol, blocks, V)) <= 1e-11, (name, d, r, grid, info)


def test_column_chunked_and_slot_limited_runs():
    """A graph whose single block exceeds the shared-memory tile capacity (column-chunked steps with partial sums
    carried in the slots) and a 1-CTA grid (runs longer than the slot capacity are split)."""
    rng = np.random.default_rng(5)
    n, d, r = 700, 3, 5                  # clique-like: diameter 2 -> one leaf of 700 poses > 600 tiles
    dh = d + 1
    el = [(0, i) for i in range(1, n)] + [(i, i + 1) for i in range(1, n - 1)] + [(1, i) for i in range(3, n, 2)]
    brow, bcol, blocks = [], [], []
    for (i, j) in el:
        M = rng.standard_normal((dh, dh))
        brow += [i, j, i, j]
        bcol += [i, j, j, i]
        blocks += [M.T @ M, np.eye(dh), -M.T, -M]
    V = rng.standard_normal((r, dh * n))
    ref = dense_reference(n, dh, brow, bcol, np.array(blocks), V)
    for grid, cuts in ((148, 0), (1, 0), (3, -1)):
        Z, info = emulate(n, d, r, np.array(brow), np.array(bcol), np.array(blocks), V, grid=grid, cuts=cuts)
        assert relerr(Z, ref) <= 1e-11, (grid, cuts, info)
        if cuts == 0:
            assert info["max_own"] == dh * n and info["max_ytiles"] <= 600
