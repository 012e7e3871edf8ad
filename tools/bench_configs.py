"""Throughput of the other BASELINE.json configs (parity-test cases, not bench lines): device-resident env steps."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from wfcrl_b200.backend import FlorisBatch
from wfcrl_b200.layouts import layout_xy

CONFIGS = [("Turb6_Row2_", 1, "f64", "basic"), ("Turb6_Row2_", 1, "f32", "fast"), ("Ablaincourt_", 4096, "f32", "fast"),
           ("Turb16_TCRWP_", 16384, "f32", "fast"), ("Turb_TCRWP_", 16384, "f32", "fast"),
           ("Turb32_Row5_", 8192, "f64", "fast"), ("Turb32_Row5_", 8192, "f64", "basic"),
           ("HornsRev1_", 8192, "f32", "fast"), ("HornsRev1_", 65536, "f32", "fast"), ("HornsRev2_", 8192, "f32", "fast")]
rows = []
for name, B, prec, kern in CONFIGS:
    lx, ly = layout_xy(name)
    T = len(lx)
    fb = FlorisBatch(lx, ly, B, precision=prec, kernel=kern, max_iter=10 ** 6)
    rng = np.random.default_rng(0)
    fb.reset(np.clip(8 * rng.weibull(8, B), 3, 28), rng.normal(270, 20, B) % 360, host_trig=False)
    acts = [(torch.rand(B, T, device="cuda") * 10 - 5) for _ in range(4)]
    for k in range(5):
        fb.step(acts[k % 4])
    torch.cuda.synchronize()
    n = 200 if T * B < 200000 else 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for k in range(n):
        fb.step(acts[k % 4])
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / n
    rows.append({"layout": name, "turbines": T, "envs": B, "precision": prec, "kernel": kern, "ms_per_step": ms,
                 "env_steps_per_s": B / ms * 1e3, "wall_us_per_call": wall / n * 1e6})
    print(rows[-1], flush=True)
    fb.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/configs.json", "w"), indent=1)
