"""Batched rollout: 8192 HornsRev1 environments stepped by one kernel launch per step, with device-side auto-reset and
episode statistics gathered across ranks (run under torchrun for several GPUs; the env batch is sharded by rank)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wfcrl_b200 import environments as envs  # noqa: E402
from wfcrl_b200.dist import shard_range  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total_envs, steps = 8192 * world, 600
lo, hi = shard_range(total_envs, rank, world)
env = envs.make_vec("HornsRev1_Floris", hi - lo, device=local, precision="f32", max_num_steps=500, env_id_offset=lo)
obs = env.reset(seed=0)
policy_gain = -0.05  # toy proportional policy: steer every turbine back towards zero yaw, plus exploration noise
for k in range(10):  # warm-up (lazy CUDA / RNG initialisation)
    obs, *_ = env.step(policy_gain * obs["yaw"] + torch.randn_like(obs["yaw"]))
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(steps):
    action = policy_gain * obs["yaw"] + torch.randn_like(obs["yaw"])
    obs, reward, terminated, truncated, info = env.step(action)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
stats = env.episode_statistics()  # the only collective: all-gather of 4 doubles per rank
if rank == 0:
    print(f"{total_envs * steps / dt / 1e6:.2f} M env-steps/s over {world} GPU(s); finished episodes: {stats}")
env.close()
if world > 1:
    dist.destroy_process_group()
