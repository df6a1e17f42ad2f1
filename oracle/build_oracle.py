"""Builds the oracle's C++ restatement (oracle/cpp/cpu_port.cpp -> oracle/_build/libdpgo_cpu_port.so).
Flags follow the reference's upstream CMakeLists (-O3 -march=native -std=c++17).  Building the checker is not
using it: only tests/, smoke() and bench.py's CPU legs load the result."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp", "cpu_port.cpp")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libdpgo_cpu_port.so")


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(HERE, "cpp", "sparse_ldl_oracle.h")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    # -march=native would tie the binary to THIS host's CPU; the GPU box may differ, so target a portable x86-64-v3
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", "-fopenmp", SRC, "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle C++ build failed:\n" + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
