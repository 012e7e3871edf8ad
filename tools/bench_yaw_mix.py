"""Device-resident step time on HornsRev1 as a function of how many turbines are actually yawed.

The step kernel skips the tip-vortex pairs of sources at exactly zero yaw (wf_fast.cu, V sweep).  The bench workload
(uniform random actions on every turbine) never hits that case; trained wake-steering policies mostly do: they hold the
back rows at zero.  Actions here are 0 for the unyawed turbines, so their yaw stays at the reset value of exactly 0.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from wfcrl_b200.backend import FlorisBatch
from wfcrl_b200.layouts import layout_xy

B = 8192
lx, ly = layout_xy("HornsRev1_")
T = len(lx)
rows = []
for frac in (1.0, 0.5, 0.25, 0.1, 0.0):
    fb = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=10 ** 6)
    rng = np.random.default_rng(0)
    fb.reset(np.clip(8 * rng.weibull(8, B), 3, 28), rng.normal(270, 20, B) % 360, host_trig=False)
    g = torch.Generator(device="cuda").manual_seed(1)
    moving = (torch.rand(B, T, device="cuda", generator=g) < frac).float()
    acts = [(torch.rand(B, T, device="cuda", generator=g) * 10 - 5) * moving for _ in range(4)]
    for k in range(5):
        out = fb.step(acts[k % 4])
    torch.cuda.synchronize()
    n = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n):
        out = fb.step(acts[k % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    yawed = float((out["yaw"] != 0).float().mean())
    rows.append({"yawed_fraction": yawed, "ms_per_step": ms, "env_steps_per_s": B / ms * 1e3})
    print(rows[-1], flush=True)
    fb.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/yaw_mix.json", "w"), indent=1)
