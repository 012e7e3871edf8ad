"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel, both precisions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from wfcrl_b200.backend import FlorisBatch
from wfcrl_b200.layouts import layout_xy

for name, prec, kern in (("Turb6_Row2_", "f64", "basic"), ("Ablaincourt_", "f32", "fast"), ("HornsRev1_", "f32", "fast"),
                         ("HornsRev2_", "f32", "fast"), ("Turb32_Row5_", "f64", "fast"), ("Turb32_Row5_", "f32", "fast")):
    lx, ly = layout_xy(name)
    B, T = 6, len(lx)
    fb = FlorisBatch(lx, ly, B, precision=prec, kernel=kern, max_iter=5)
    rng = np.random.default_rng(0)
    ws = np.clip(8 * rng.weibull(8, B), 3, 28)
    ws[:3] = [3.2, 3.6, 4.0]   # low wind: turbines on the foot of the power curve -> the FP64 re-solve kernel runs (strict FP32)
    wd = rng.normal(270, 20, B) % 360
    wd[1] = 270.0              # x-ties on the row layouts: direct vortex evaluation next to the table
    fb.reset(ws, wd)
    if kern == "fast":
        fb.set_autoreset(True, seed=3, env_id_offset=11, turbulence_intensity_range=(0.05, 0.1))
    n_fix = 0
    for k in range(5):
        out = fb.step(torch.as_tensor(rng.uniform(-5, 5, (B, T)).astype(np.float32), device="cuda"))
        if prec == "f32" and kern == "fast":
            n_fix += int(fb.get_state("ambiguous").sum())
        if kern == "fast" and bool(out["truncated"].any()):
            fb.autoreset_finish()   # in-kernel auto-reset: wind draw in the geometry kernel, table rebuild, warm-up solve
    mask = out["truncated"].clone()
    fb.reset_masked(mask, torch.full((B,), 9.0, dtype=torch.float64, device="cuda"),
                    torch.full((B,), 265.0, dtype=torch.float64, device="cuda"))
    fb.reset_sampled(mask, seed=7, env_id_offset=3, turbulence_intensity_range=(0.04, 0.12))
    yaw = torch.zeros(B, T, device="cuda")
    yaw[:, ::2] = 7.0  # mixed yawed / unyawed sources
    fb.update_command(yaw.double())
    fb.step_host(torch.zeros(B, T).pin_memory())          # zero-copy route (mapped pinned buffers)
    os.environ["WFCRL_B200_HOST_PATH"] = "staged"
    fb.step_host(torch.zeros(B, T).pin_memory())          # staged route (copy engines)
    del os.environ["WFCRL_B200_HOST_PATH"]
    torch.cuda.synchronize()
    print(name, prec, kern, "ok", float(out["reward"].sum()), "env solves redone in FP64 during the env steps:", n_fix)
    fb.close()
