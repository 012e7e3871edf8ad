"""Does the order of the envs inside the batch matter for the step time (tail wave = last 0.46 wave of CTAs)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wfcrl_b200.backend import FlorisBatch
from wfcrl_b200.layouts import layout_xy

B = 8192
lx, ly = layout_xy("HornsRev1_")
T = len(lx)
rng = np.random.default_rng(0)
ws0 = np.clip(8 * rng.weibull(8, B), 3, 28)
wd0 = rng.normal(270, 20, B) % 360
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def run(order, label):
    fb = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=10 ** 6)
    fb.reset(ws0[order], wd0[order], host_trig=False)
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = [(torch.rand(B, T, device="cuda", generator=g) * 10 - 5) for _ in range(4)]
    for k in range(5):
        fb.step(acts[k % 4])
    ts = []
    for k in range(30):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fb.step(acts[k % 4]); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{label:28s} {np.mean(ts):.4f} ms  ({B / np.mean(ts) / 1e3:.3f} M env-steps/s)", flush=True)
    fb.close()

run(np.arange(B), "random order")
run(np.argsort(wd0), "sorted by wd ascending")
run(np.argsort(-wd0), "sorted by wd descending")
dev = np.abs(((wd0 - 270 + 180) % 360) - 180)
run(np.argsort(dev), "aligned (|wd-270| small) first")
run(np.argsort(-dev), "aligned last")
run(np.argsort(ws0), "slow wind first")
