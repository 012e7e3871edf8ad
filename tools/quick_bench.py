"""Device-timed env steps only (no host path, no CPU leg): quick A/B numbers while tuning a kernel.
usage: python tools/quick_bench.py [layout] [envs] [precision] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from wfcrl_b200.backend import FlorisBatch
from wfcrl_b200.layouts import layout_xy

name = sys.argv[1] if len(sys.argv) > 1 else "HornsRev1_"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
prec = sys.argv[3] if len(sys.argv) > 3 else "f32"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
lx, ly = layout_xy(name)
T = len(lx)
rng = np.random.default_rng(0)
ws = np.clip(8 * rng.weibull(8, B), 3, 28)
wd = np.clip(rng.normal(270, 20, B) % 360, 0, 360)
fb = FlorisBatch(lx, ly, B, precision=prec, kernel="fast", max_iter=100000)
fb.reset(ws, wd, host_trig=False)
pool = [(torch.rand(B, T, device="cuda") * 10 - 5).contiguous() for _ in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for k in range(5):
    fb.step(pool[k % 4])
torch.cuda.synchronize()
res = {}
for do_flush in (True, False):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        if do_flush:
            flush.zero_()
        ev[k][0].record()
        fb.step(pool[k % 4])
        ev[k][1].record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    res["flush" if do_flush else "noflush"] = (float(np.mean(ms)), ms[0])
info = fb.device_info()
print(f"{name} B={B} {prec} T={T} regs={info['regs_per_thread']} ctas/sm={info['ctas_per_sm']} smem={info['smem_per_cta']} "
      f"tag={os.environ.get('TAG', '')} | flushed: mean {res['flush'][0]:.4f} ms min {res['flush'][1]:.4f} -> {B / res['flush'][0] / 1e3:.3f} M steps/s"
      f" | back-to-back: mean {res['noflush'][0]:.4f} ms -> {B / res['noflush'][0] / 1e3:.3f} M steps/s", flush=True)
fb.close()
