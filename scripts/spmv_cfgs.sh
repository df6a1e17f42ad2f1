#!/bin/bash
# A/B of the SpMV kernel configurations (run under gpurun): DPGO_SPMV_CFG selects the instantiation for r=5, d=3
for c in ${@:-0 1 2 3 4 5 6}; do echo "cfg $c"; DPGO_SPMV_CFG=$c timeout 200 python scripts/spmv_sweep.py 2>&1 | tail -1; done
echo v1; DPGO_SPMV_V1=1 timeout 200 python scripts/spmv_sweep.py 2>&1 | tail -1
