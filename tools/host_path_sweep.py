"""wf_step_host wall time per call: pinned buffers mapped into the kernel (zero-copy) vs device staging + copy engines."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from wfcrl_b200.backend import FlorisBatch
from wfcrl_b200.layouts import layout_xy

rows = []
for name in ("Ablaincourt_", "HornsRev1_"):
    lx, ly = layout_xy(name)
    T = len(lx)
    for B in (1, 16, 64, 256, 512, 1024, 2048, 4096, 8192):
        fb = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=10 ** 6)
        rng = np.random.default_rng(0)
        fb.reset(np.clip(8 * rng.weibull(8, B), 3, 28), rng.normal(270, 20, B) % 360, host_trig=False)
        a = torch.empty(B, T).uniform_(-5, 5).pin_memory()
        row = {"layout": name, "envs": B}
        for mode in ("zero_copy", "staged"):
            os.environ["WFCRL_B200_HOST_PATH"] = mode
            for _ in range(5):
                fb.step_host(a)
            n = 200 if B * T < 50000 else 30
            t0 = time.perf_counter()
            for _ in range(n):
                fb.step_host(a)
            row[mode + "_us"] = (time.perf_counter() - t0) / n * 1e6
        rows.append(row)
        print(row, flush=True)
        fb.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/host_path_sweep.json", "w"), indent=1)
