"""dpo_b200 -- B200-native distributed pose-graph optimisation hot path.

Host-side Python mirror of the reference's QuadraticProblem / QuadraticOptimizer / PGOAgent
interface over the C ABI (include/dpgo_b200.h, libdpgo_b200.so).  No CPU compute fallback.
"""
from ._capi import (ALG_RGD, ALG_RTR, PRECOND_BLOCK_JACOBI, PRECOND_DENSE_EXACT, PRECOND_NONE, PRECOND_SPARSE_EXACT, DpgoError,
                    load_library)
from .problem import QuadraticOptimizer, QuadraticProblem, ROPTALG

__all__ = ["QuadraticProblem", "QuadraticOptimizer", "ROPTALG", "DpgoError", "load_library", "ALG_RTR", "ALG_RGD",
           "PRECOND_NONE", "PRECOND_BLOCK_JACOBI", "PRECOND_DENSE_EXACT", "PRECOND_SPARSE_EXACT"]
