"""The oracle's C++ restatement (the timed CPU baseline) agrees with the NumPy oracle and performs the
reference's operation count (10 + j products, j + 1 exact solves per RTR call)."""
import os

import numpy as np
import pytest

from oracle import cpu_port
from oracle import dpgo_oracle as orc


@pytest.mark.parametrize("ds,r", [("smallGrid3D", 5), ("CSAIL", 5), ("sphere2500", 3)])
def test_cpu_port_matches_numpy_oracle(ds, r, data_dir):
    assert cpu_port.available()
    meas, n = orc.read_g2o(os.path.join(data_dir, ds + ".g2o"))
    d = meas.d
    Q = orc.construct_connection_laplacian(meas, n)
    X0 = orc.fixed_stiefel_variable(d, r) @ orc.chordal_initialization(meas, n)
    prob = orc.QuadraticProblem(n, d, r)
    prob.set_Q(Q)
    run = cpu_port.Runner(Q, n, d, r)
    Xo, Xc = X0, X0
    for _ in range(3):
        oo = orc.QuadraticOptimizer(prob)
        oo.tr_tolerance, oo.tr_iterations, oo.tr_max_inner, oo.tr_initial_radius = 1e-2, 1, 10, 100.0
        Xo = oo.optimize(Xo)
        Xc = run.step(Xc)
        res = run.result
        assert res.tcg_iterations == oo.result.tcg_iterations and res.tcg_status == oo.result.tcg_status
        assert abs(res.f_opt - oo.result.fOpt) <= 1e-9 * abs(oo.result.fOpt)
        assert np.linalg.norm(Xc - Xo) <= 1e-9 * np.linalg.norm(Xo)
        if res.tcg_iterations:
            assert res.spmv == 10 + res.tcg_iterations
            assert res.solves in (res.tcg_iterations, res.tcg_iterations + 1)
