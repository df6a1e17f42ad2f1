import sys, json, os
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import dpo_b200 as dp
from dpo_b200 import posegraph as pg
import bench
peak, src = bench.measured_peaks()
out=[]
for dims in ((25,20,20),(50,50,40),(100,100,40),(100,100,100)):
    r=bench.spmv_roofline(torch, dp, pg, peak, src, dims=dims, reps=20)
    out.append((dims[0]*dims[1]*dims[2], round(r["us_per_launch"],1), round(r["frac"],4)))
print(out)
