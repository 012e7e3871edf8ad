import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from tests._util import layout
from tests.test_fp32_strict_gpu import _solve
name, yaw_amp = "HornsRev1_", 40.0
lx, ly = layout(name); T = len(lx); B = 32768
rng = np.random.default_rng(int(yaw_amp) + T)
ws = np.clip(8 * rng.weibull(8, B), 3, 28)
ws[: B // 16] = rng.uniform(3.0, 4.5, B // 16)
wd = np.clip(rng.normal(270, 20, B) % 360, 0, 360)
yaw = rng.uniform(-yaw_amp, yaw_amp, (B, T)).astype(np.float32).astype(np.float64)
yaw_t = torch.as_tensor(yaw, device="cuda")
ref, _, _ = _solve(lx, ly, ws, wd, yaw_t, "f64")
got, flag, it = _solve(lx, ly, ws, wd, yaw_t, "f32")
raw, _, _ = _solve(lx, ly, ws, wd, yaw_t, "f32", strict=False)
err = np.abs(got["power"] - ref["power"]) / np.maximum(ref["power"], 1.0)
bad = np.argwhere(err > 5e-5)
print("flagged", flag.sum(), "bad", len(bad))
for b, t in bad:
    print(b, t, "ws", ws[b], "wd", wd[b], "yaw", yaw[b, t], "P64", ref["power"][b, t], "P32", got["power"][b, t], "err", err[b, t],
          "wsl64", ref["wind_speed"][b, t], "wsl32", got["wind_speed"][b, t], "relws", abs(got["wind_speed"][b,t]-ref["wind_speed"][b,t])/ref["wind_speed"][b,t],
          "ti64", ref["load"][b, t, 0]/1e7, "ti32", got["load"][b, t, 0]/1e7, "flag", flag[b])
    e_env = np.abs(got["wind_speed"][b] - ref["wind_speed"][b]) / ref["wind_speed"][b]
    print("   env max ws err", e_env.max(), "at", e_env.argmax(), " ti err max", (np.abs(got["load"][b,:,0]-ref["load"][b,:,0])/ref["load"][b,:,0]).max())
d = np.abs(got["power"][flag] - ref["power"][flag]) / np.maximum(ref["power"][flag], 1.0)
print("flagged envs: max rel diff to f64", d.max())
