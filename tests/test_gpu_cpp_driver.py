"""Our own C++ driver (examples/MultiAgentPGO.cpp over libDPGO.so): the C++ PGOAgent host path on more datasets
than the reference's hard-wired example, including Nesterov acceleration (ref src/PGOAgent.cpp:685-695,1040-1091;
Stiefel projection on the GPU) against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import dpgo_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "examples", "MultiAgentPGO")


def run_driver(tmp_path, ds, *flags):
    if not os.path.exists(EXE):
        pytest.skip("build/examples/MultiAgentPGO not built")
    trace = os.path.join(str(tmp_path), "trace.csv")
    res = subprocess.run([EXE, os.path.join(ROOT, "data", ds + ".g2o"), "--trace", trace] + list(flags),
                         capture_output=True, text=True, timeout=900)
    print(res.stdout[-500:], res.stderr[-1500:])
    assert res.returncode == 0
    return np.loadtxt(trace, delimiter=",").reshape(-1, 4)


def test_cpp_driver_reproduces_golden_trace_sphere2500(tmp_path, golden_dir):
    tr = run_driver(tmp_path, "sphere2500", "--robots", "5", "--iters", "80", "--stop", "0")
    gold = np.loadtxt(os.path.join(golden_dir, "NPsphere2500_head400.txt"), delimiter=",")[:80]
    assert np.max(np.abs(tr[:, 2] - gold[:, 0]) / gold[:, 0]) <= 5e-9
    assert np.max(np.abs(tr[:, 3] - gold[:, 1]) / gold[:, 1]) <= 5e-8


def test_cpp_driver_accelerated_matches_oracle(tmp_path, data_dir):
    iters = 100
    tr = run_driver(tmp_path, "smallGrid3D", "--robots", "5", "--iters", str(iters), "--stop", "0", "--accel")
    meas, n = orc.read_g2o(os.path.join(data_dir, "smallGrid3D.g2o"))
    drv = orc.MultiRobotDriver(meas, n, 5, r=5, acceleration=True)
    ot = drv.run(iters)
    assert [int(a) for a in tr[:, 1]] == ot.selected          # same greedy selection sequence
    assert np.max(np.abs(tr[:, 2] - np.array(ot.cost)) / np.array(ot.cost)) <= 1e-8
    assert np.max(np.abs(tr[:, 3] - np.array(ot.gradnorm)) / np.array(ot.gradnorm)) <= 1e-6


def test_cpp_driver_throughput_mode_converges(tmp_path):
    tr = run_driver(tmp_path, "smallGrid3D", "--robots", "4", "--iters", "600", "--stop", "0.1", "--jacobi")
    assert tr[-1, 3] < 0.1
    assert abs(tr[-1, 2] - 1025.398) <= 2e-4 * 1025.398


def _device_count():
    import ctypes
    from dpo_b200 import _capi
    c = ctypes.c_int(0)
    _capi.load_library().dpgo_device_count(ctypes.byref(c))
    return c.value


@pytest.mark.parametrize("gpus", [1, 2])
def test_cpp_resident_runner_reproduces_golden_trace(tmp_path, golden_dir, gpus):
    """DPGO::DeviceRBCD (C++ host, iterates resident in HBM, public poses by ncclAllGather): greedy schedule, 5 agents
    on 1 GPU -- and, where the box has them, coloured 8 agents over 2 GPUs against the oracle's coloured driver."""
    if gpus > _device_count():
        pytest.skip(f"needs {gpus} GPUs")
    if gpus == 1:
        tr = run_driver(tmp_path, "torus3D", "--robots", "5", "--iters", "40", "--stop", "0", "--resident")
        gold = np.loadtxt(os.path.join(golden_dir, "NPtorus3D_head400.txt"), delimiter=",")[:40]
        assert np.max(np.abs(tr[:, 2] - gold[:, 0]) / gold[:, 0]) <= 5e-9
        assert np.max(np.abs(tr[:, 3] - gold[:, 1]) / gold[:, 1]) <= 5e-8
    else:
        tr = run_driver(tmp_path, "torus3D", "--robots", "8", "--iters", "10", "--stop", "0", "--resident", "--gpus", "2",
                        "--schedule", "coloured")
        meas, n = orc.read_g2o(os.path.join(ROOT, "data", "torus3D.g2o"))
        drv = orc.MultiRobotDriver(meas, n, 8, r=5, schedule="coloured")
        ot = drv.run(10)
        assert np.max(np.abs(tr[:, 2] - np.array(ot.cost)) / np.array(ot.cost)) <= 1e-8
        assert np.max(np.abs(tr[:, 3] - np.array(ot.gradnorm)) / np.array(ot.gradnorm)) <= 1e-7


def test_cpp_resident_runner_concurrent_agents_on_one_gpu(tmp_path):
    """8 agents, coloured schedule, ONE GPU: DeviceRBCD steps the 4 agents of a colour class side by side (thread-block
    clusters on their own streams, dpgo_agents_round_async, CUDA-graph replay) -- against the oracle's coloured driver."""
    tr = run_driver(tmp_path, "torus3D", "--robots", "8", "--iters", "10", "--stop", "0", "--resident", "--schedule", "coloured")
    meas, n = orc.read_g2o(os.path.join(ROOT, "data", "torus3D.g2o"))
    drv = orc.MultiRobotDriver(meas, n, 8, r=5, schedule="coloured")
    ot = drv.run(10)
    assert np.max(np.abs(tr[:, 2] - np.array(ot.cost)) / np.array(ot.cost)) <= 1e-8
    assert np.max(np.abs(tr[:, 3] - np.array(ot.gradnorm)) / np.array(ot.gradnorm)) <= 1e-7


def test_cpp_resident_runner_with_partition_file(tmp_path, golden_dir):
    """--partition: ownership from the reference's graph-partition file through the C++ resident runner."""
    tr = run_driver(tmp_path, "CSAIL", "--robots", "5", "--iters", "50", "--stop", "0", "--resident", "--partition",
                    os.path.join(golden_dir, "partition5_strong_CSAIL.txt"))
    gold = np.loadtxt(os.path.join(golden_dir, "strongCSAIL_head400.txt"), delimiter=",")[:50]
    assert np.max(np.abs(tr[:, 2] - gold[:, 0]) / gold[:, 0]) <= 5e-9
    assert np.max(np.abs(tr[:, 3] - gold[:, 1]) / gold[:, 1]) <= 5e-8
