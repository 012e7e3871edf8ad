"""N-GPU concurrent pinned host<->device copy probe (plain cudaMemcpyAsync through torch, no kernels): what the box's host
side can move when every GPU of the job copies at once -- the ceiling of bench.py's end-to-end (`e2e`) figure.

    python tools/pcie_probe_multi.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe_multi.py

Per rank, the byte counts of one HornsRev1 x 8192 step of wf_step_host: 2.62 MB host->device, 21.0 MB device->host.
Patterns: (a) one copy per direction, (b) the library's pattern: 6 chunks x (1 + 8) copies on 6 streams.
Rank 0 prints one JSON object (per-rank and aggregate GB/s, and the env-steps/s those rates allow)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

B, T = 8192, 80
H2D = B * T * 4
FIELDS = [B * T * 4 * 4, B * T * 4, B * T * 4, B * T * 4, B * T * 4, B * 4, B * 8, B]  # load, yaw, ws, wd, power, reward, freewind, trunc
D2H = sum(FIELDS)
h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory()
d_in = torch.empty(H2D, dtype=torch.uint8, device=dev)
h_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in FIELDS]
d_out = [torch.empty(n, dtype=torch.uint8, device=dev) for n in FIELDS]
h_packed = torch.empty(D2H, dtype=torch.uint8).pin_memory()
d_packed = torch.empty(D2H, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream(device=dev) for _ in range(6)]


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def one_copy_per_direction():
    with torch.cuda.stream(streams[0]):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(streams[1]):
        h_packed.copy_(d_packed, non_blocking=True)


def library_pattern():
    nch = 6
    for c in range(nch):
        with torch.cuda.stream(streams[c]):
            a, b = H2D * c // nch, H2D * (c + 1) // nch
            d_in[a:b].copy_(h_in[a:b], non_blocking=True)
            for hf, df in zip(h_out, d_out):
                n = hf.numel()
                a, b = n * c // nch, n * (c + 1) // nch
                hf[a:b].copy_(df[a:b], non_blocking=True)


def measure(fn, seconds=2.0):
    for _ in range(5):
        fn()
    barrier()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        n += 20
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt / n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


res = {"n_gpus": world, "h2d_bytes_per_step": H2D, "d2h_bytes_per_step": D2H, "host_cpus": len(os.sched_getaffinity(0))}
for name, fn in (("one_copy_per_direction", one_copy_per_direction), ("library_pattern_6x9_copies", library_pattern)):
    s = measure(fn)
    res[name] = {"seconds_per_step_max_over_ranks": s, "aggregate_GBps": world * (H2D + D2H) / s / 1e9,
                 "per_gpu_GBps": (H2D + D2H) / s / 1e9, "env_steps_per_s_ceiling": world * B / s}
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
