"""Plain device-to-host copy bandwidth into pinned memory, per rank and aggregated (torchrun-aware): the ceiling of bench.py's `e2e` figure,
which brings one 8.29 MB colour image per frame back to the host. Usage: python tools/d2h_bandwidth.py  |  torchrun --nproc-per-node N tools/d2h_bandwidth.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
frames = 128
x = torch.empty((frames, 1080, 1920), dtype=torch.int32, device="cuda")
h = torch.empty((frames, 1080, 1920), dtype=torch.int32).pin_memory()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(2):
    h.copy_(x, non_blocking=True)
barrier()
t0 = time.perf_counter()
reps = 4
for _ in range(reps):
    h.copy_(x, non_blocking=True)
barrier()
elapsed = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
gb = reps * frames * 1080 * 1920 * 4 / 1e9
if rank == 0:
    per_rank = gb / float(elapsed.item())
    print(json.dumps({"n_gpus": world, "d2h_gb_s_per_gpu": per_rank, "d2h_gb_s_total": per_rank * world, "frames_per_s_ceiling_total": per_rank * world * 1e9 / (1080 * 1920 * 4)}))
if world > 1:
    dist.destroy_process_group()
