import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from wfcrl_b200.backend import FlorisBatch
from wfcrl_b200.layouts import layout_xy
lx, ly = layout_xy("Turb32_Row5_")
B, T = 8192, len(lx)
fb = FlorisBatch(lx, ly, B, precision="f64", kernel="fast", max_iter=10**6)
rng = np.random.default_rng(0)
fb.reset(np.clip(8*rng.weibull(8,B),3,28), rng.normal(270,20,B)%360, host_trig=False)
a = torch.rand(B,T,device="cuda")*10-5
for k in range(6): fb.step(a)
torch.cuda.synchronize()
