"""Where does the FP32 kernel's error tail come from?  Per-turbine local-wind-speed error vs the FP64 kernel, binned by the
free-stream speed, plus the worst cases.  Output: gpurun_out/err_probe.json / stdout."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from wfcrl_b200.backend import FlorisBatch
from wfcrl_b200.layouts import layout_xy

name = os.environ.get("PROBE_LAYOUT", "HornsRev1_")
amp = float(os.environ.get("PROBE_YAW", "5"))
B = 65536
lx, ly = layout_xy(name)
T = len(lx)
rng = np.random.default_rng(5)
ws = np.clip(8 * rng.weibull(8, B), 3, 28)
wd = np.clip(rng.normal(270, 20, B) % 360, 0, 360)
yaw = rng.uniform(-amp, amp, (B, T)).astype(np.float32).astype(np.float64)
yaw_t = torch.as_tensor(yaw, device="cuda")
res = {}
for prec in ("f64", "f32"):
    fb = FlorisBatch(lx, ly, B, precision=prec, kernel="fast", max_iter=10)
    fb.reset(ws, wd, host_trig=True, warmup_solves=0)
    o = fb.update_command(yaw_t)
    torch.cuda.synchronize()
    res[prec] = {k: o[k].double().cpu().numpy() for k in ("power", "wind_speed", "wind_direction", "load")}
    if prec == "f32":
        flag = fb.get_state("ambiguous").astype(bool)
    order = fb.get_state("order")
    fb.close()
ok = ~flag
r64, r32 = res["f64"], res["f32"]
ews = np.abs(r32["wind_speed"] - r64["wind_speed"]) / r64["wind_speed"]
epw = np.abs(r32["power"] - r64["power"]) / np.maximum(r64["power"], 1.0)
eti = np.abs(r32["load"][..., 0] - r64["load"][..., 0]) / r64["load"][..., 0]
ews[~ok] = 0
epw[~ok] = 0
eti[~ok] = 0
print("flagged", int(flag.sum()), "of", B)
bins = [3, 3.5, 4, 4.5, 5, 6, 7, 8, 9, 10, 12, 14, 28.1]
for lo, hi in zip(bins[:-1], bins[1:]):
    m = (ws >= lo) & (ws < hi)
    if m.sum() == 0:
        continue
    print(f"ws [{lo:4.1f},{hi:4.1f}) n={int(m.sum()):6d}  ws_err max {ews[m].max():.2e} p99.9 {np.percentile(ews[m], 99.9):.2e} med {np.median(ews[m]):.2e}"
          f" | power_err max {epw[m].max():.2e} p99.9 {np.percentile(epw[m], 99.9):.2e} | ti_err max {eti[m].max():.2e}")
# position of each turbine in the sorted order
pos = np.argsort(order, axis=1)
flat = np.argsort(ews.ravel())[::-1][:25]
print("worst local-wind-speed errors (env, turbine, sorted position, ws, wd, yaw, ws_local64, ti64, P64 kW, err_ws, err_P, err_ti):")
for f in flat:
    b, t = divmod(int(f), T)
    print(b, t, int(pos[b, t]), f"{ws[b]:.3f} {wd[b]:.2f} {yaw[b, t]:+.2f} {r64['wind_speed'][b, t]:.4f} {r64['load'][b, t, 0] / 1e7:.4f} "
          f"{r64['power'][b, t] / 1e3:.1f} {ews[b, t]:.2e} {epw[b, t]:.2e} {eti[b, t]:.2e}")
flat = np.argsort(epw.ravel())[::-1][:15]
print("worst power errors:")
for f in flat:
    b, t = divmod(int(f), T)
    print(b, t, int(pos[b, t]), f"{ws[b]:.3f} {wd[b]:.2f} {yaw[b, t]:+.2f} {r64['wind_speed'][b, t]:.4f} {r64['load'][b, t, 0] / 1e7:.4f} "
          f"{r64['power'][b, t] / 1e3:.1f} {ews[b, t]:.2e} {epw[b, t]:.2e} {eti[b, t]:.2e}")
