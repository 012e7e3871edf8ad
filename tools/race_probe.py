import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wfcrl_b200.backend import FlorisBatch
from wfcrl_b200.layouts import layout_xy
name = sys.argv[1]
lx, ly = layout_xy(name); B, T = 6, len(lx)
fb = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=5)
rng = np.random.default_rng(0)
ws = np.array([3.2, 3.6, 4.0, 3.1, 3.3, 3.5]); wd = rng.normal(270, 20, B) % 360
fb.reset(ws, wd)
for k in range(2):
    out = fb.step(torch.as_tensor(rng.uniform(-5, 5, (B, T)).astype(np.float32), device="cuda"))
torch.cuda.synchronize()
print(name, "W", os.environ.get("WFCRL_B200_FIX_WARPS"), "redone", int(fb.get_state("ambiguous").sum()))
