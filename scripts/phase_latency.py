import sys, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np, dpo_b200 as dp
from dpo_b200 import posegraph as pg, _capi
edges,n=pg.read_g2o_file('/root/repo/data/sphere2500.g2o')
for precs,label in (((dp.PRECOND_BLOCK_JACOBI,),'jacobi-grid'),((dp.PRECOND_BLOCK_JACOBI,dp.PRECOND_DENSE_EXACT),'dense-grid(148)')):
    prob=dp.QuadraticProblem(n,3,5,preconditioners=precs)
    prob.setQ_blocks(*pg.connection_laplacian_blocks(edges))
    a=C.c_double(); b=C.c_double()
    _capi.check(prob._lib.dpgo_debug_phase_latency(prob._h, 50, C.byref(a), C.byref(b)))
    print(label, 'us/phase %.2f  launch+first %.2f'%(a.value,b.value))
