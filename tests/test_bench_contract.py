"""bench.py contract (CPU side): the reference arm runs without a GPU and prints ONE JSON line with the keys the
driver reads; the product arm must refuse to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_json_line():
    out = run_bench("--impl", "reference", "--steps", "3", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        base = json.load(fh)
    assert d["impl"] == "reference" and "unavailable" not in d
    assert d["steps"] == 3 and d["warmup"] >= 1 and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "iter/s" and d["dtype"] == "f64"
    if isinstance(base.get("metric"), str):
        assert "iter" in base["metric"].lower() or "rtr" in d["metric"]
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) <= 1e-6 * 1e3
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_fails_loudly_without_gpu():
    out = run_bench("--steps", "1", "--warmup", "1", "--no-cpu", "--no-spmv")
    assert out.returncode != 0
    assert "{\"metric\"" not in out.stdout                   # no bench line from a CPU fallback
