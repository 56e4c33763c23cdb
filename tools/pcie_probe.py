"""Ceiling of the end-to-end leg: every rank copies 1080p RGBA8 images from its GPU into its own pinned host buffer, all ranks at once.
Run: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py
Rank 0 prints one JSON line with the aggregate device-to-host rate; bench.py's e2e (frames/s x 8.29 MB) is read against it."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    bind = os.environ.get("PROBE_BIND", "1") == "1"
    if bind:
        import bench
        bench.bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    views = 256 // world
    dev = torch.zeros((views, 1080, 1920), dtype=torch.int32, device="cuda")
    host = torch.empty((views, 1080, 1920), dtype=torch.int32).pin_memory()
    for _ in range(2):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    if rank == 0:
        total = reps * 256 * 1080 * 1920 * 4
        print(json.dumps({"n_gpus": world, "bound_to_numa_node": bind, "d2h_gb_per_s_aggregate": total / float(el.item()) / 1e9,
                          "frames_per_s_ceiling": reps * 256 / float(el.item())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
