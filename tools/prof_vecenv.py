import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wfcrl_b200 import environments as envs

name, B = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("HornsRev1_Floris", 8192)
env = envs.make_vec(name, B, precision="f32", max_num_steps=100000)
obs = env.reset(seed=0)
a = torch.randn(B, env.num_turbines, device="cuda")
print(name, B)
def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("backend.step      ms", timeit(lambda: env.backend.step(a)))
print("env.step          ms", timeit(lambda: env.step(a)))
print("env.step + policy ms", timeit(lambda: env.step(-0.05 * env.backend.out["yaw"] + torch.randn_like(a))))
t0 = time.perf_counter()
for _ in range(200): env.step(a)
print("python-side us per env.step call (async)", (time.perf_counter() - t0) / 200 * 1e6)
torch.cuda.synchronize()
