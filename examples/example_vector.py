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
total_envs = int(os.environ.get("WFCRL_TOTAL_ENVS", 8192 * world))  # fixed total: N-GPU runs reproduce the 1-GPU episodes
steps = int(os.environ.get("WFCRL_STEPS", 600))
lo, hi = shard_range(total_envs, rank, world)
env = envs.make_vec("HornsRev1_Floris", hi - lo, device=local, precision="f32", max_num_steps=500, env_id_offset=lo)
obs = env.reset(seed=0)
policy_gain = -0.05  # toy proportional policy: steer every turbine back towards zero yaw, plus exploration noise
gids = torch.arange(lo, hi, device=obs["yaw"].device, dtype=torch.float64)[:, None]
cols = torch.arange(obs["yaw"].shape[1], device=obs["yaw"].device, dtype=torch.float64)[None, :]


def noise(k):
    """Exploration noise in [-2, 2) as a function of (global env id, turbine, step): identical however the batch is
    sharded over ranks, so an N-GPU run reproduces the single-GPU episodes."""
    return torch.sin(12.9898 * gids + 78.233 * cols + 37.719 * k).mul(43758.5453).frac().float() * 4 - 2


for k in range(-10, 0):  # warm-up (lazy CUDA initialisation)
    obs, *_ = env.step(policy_gain * obs["yaw"] + noise(k))
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(steps):
    obs, reward, terminated, truncated, info = env.step(policy_gain * obs["yaw"] + noise(k))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
stats = env.episode_statistics()  # the only collective: all-gather of 4 doubles per rank
if rank == 0:
    print(f"{total_envs * steps / dt / 1e6:.2f} M env-steps/s over {world} GPU(s); finished episodes: {stats}")
env.close()
if world > 1:
    dist.destroy_process_group()
