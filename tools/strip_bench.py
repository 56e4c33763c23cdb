"""Screen-strip mode across GPUs: every rank rasterises rows strip_rows(height, world)[rank] of ONE frame from a
replicated scene (dfpsr_renderer_set_clip_rows). Two ways to assemble the frame are timed:
  gather  a single NCCL all_gather over NVLink puts the whole frame on every rank (the library baseline)
  peer    the presenting rank (0) owns the frame, the others map it (dfpsr_peer_open) and their tile kernels store their strips straight
          into it over NVLink; flags in peer memory order the frames (shard.PeerStripFrame) — no collective, no staging copy
Run: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/strip_bench.py [--scene terrain|tiny]
Rank 0 checks the gathered frame against a full-frame render on its own GPU and prints one JSON line."""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from dfpsr_b200 import abi, lib, scenes, shard


def run(cuda, scene, iters, rank, world):
    """One frame in row strips across the `world` ranks of the initialised process group; every rank must call this. Returns the result
    dict on rank 0 (None elsewhere). Raises if the assembled frame differs from the full-frame render of rank 0's own GPU."""
    if scene == "terrain":
        w, h = 1920, 1080
        sc = scenes.terrain_scene()
        tex = lib.DeviceTexture(sc["texture"], 5)
        model = lib.DeviceModel(sc["points"], sc["polygons"], abi.FILTER_SOLID, tex)
        cam = lib.camera(scenes.orbit_camera(7, w, h))
    else:
        w, h = 3840, 2160
        sc = scenes.tiny_triangle_scene(1000, 999)
        model = lib.DeviceModel(sc["points"], sc["polygons"])
        cam = lib.camera(scenes.top_down_camera(1000, 999, w, h))
    ident = abi.Transform3D.identity()
    sp = lib.stream_ptr()
    color = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    depth = torch.zeros((h, w), dtype=torch.float32, device="cuda")
    r = C.c_void_p()
    lib.check(cuda.dfpsr_renderer_create(C.byref(r)))
    bounds = shard.strip_rows(h, world, align=4)

    def frame(strip):
        lib.check(cuda.dfpsr_renderer_begin_cleared(r, C.byref(lib.image(color)), C.byref(lib.image(depth)), 0, 0.0))
        if strip:
            lib.check(cuda.dfpsr_renderer_set_clip_rows(r, bounds[rank][0], bounds[rank][1]))
        lib.check(cuda.dfpsr_renderer_give_task(r, C.byref(model.desc), C.byref(ident), C.byref(cam), sp))
        lib.check(cuda.dfpsr_renderer_end(r, sp))
        if strip and world > 1:
            lib.check(cuda.dfpsr_renderer_flush(r))  # the all_gather is not queued through the library
            shard.gather_strips(color, bounds)

    psf = shard.PeerStripFrame(shard.CudaPeerTransport(cuda), h, w, rank, world) if world > 1 else None
    counter = [0]

    def peer_frame(_strip):
        counter[0] += 1
        k = counter[0]
        psf.begin_frame(k)
        lib.check(cuda.dfpsr_renderer_begin_cleared(r, C.byref(lib.image_from_ptr(psf.color_ptr, w, h)), C.byref(lib.image(depth)), 0, 0.0))
        lib.check(cuda.dfpsr_renderer_set_clip_rows(r, psf.rows[0], psf.rows[1]))
        lib.check(cuda.dfpsr_renderer_give_task(r, C.byref(model.desc), C.byref(ident), C.byref(cam), sp))
        lib.check(cuda.dfpsr_renderer_end(r, sp))
        psf.end_frame(k)      # presenter: its stream now waits for every strip
        psf.release_frame(k)  # presenter: the frame is consumed at this point of its stream (a display hand-off would sit before this)

    def timed(strip, frame=frame):
        for _ in range(3):
            frame(strip)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            frame(strip)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / iters], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    ms_strip = timed(True)
    lib.check(cuda.dfpsr_renderer_flush(r))
    gathered = color.clone()
    ms_peer, peer_frame_copy, peer_timed_out = None, None, 0
    if psf is not None:
        ms_peer = timed(True, peer_frame)
        lib.check(cuda.dfpsr_renderer_flush(r))
        peer_timed_out = psf.timed_out()
        if rank == 0:
            peer_frame_copy = lib.tensor_from_ptr(psf.color_ptr, (h, w)).clone()
    ms_full = timed(False)
    lib.check(cuda.dfpsr_renderer_flush(r))
    torch.cuda.synchronize()
    same = bool(torch.equal(gathered, color))
    peer_same = bool(torch.equal(peer_frame_copy, color)) if peer_frame_copy is not None else None
    result = None
    if rank == 0:
        result = {"scene": scene, "n_gpus": world, "width": w, "height": h, "strip_frame_ms": ms_strip, "single_gpu_frame_ms": ms_full,
                  "speedup": ms_full / ms_strip, "gathered_equals_full_frame": same, "gather_bytes_per_rank": (bounds[rank][1] - bounds[rank][0]) * w * 4,
                  "peer_strip_frame_ms": ms_peer, "peer_speedup": (ms_full / ms_peer) if ms_peer else None, "peer_equals_full_frame": peer_same,
                  "peer_waits_timed_out": peer_timed_out}
    if psf is not None:
        psf.close()
    lib.check(cuda.dfpsr_renderer_destroy(r))
    assert same, "strip-sharded frame differs from the full-frame render"
    assert peer_same is not False and not peer_timed_out, "peer-store strip frame differs from the full-frame render (or a wait timed out)"
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="tiny", choices=["terrain", "tiny"])
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    cuda = lib.load()
    lib.check(cuda.dfpsr_init(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    result = run(cuda, args.scene, args.iters, rank, world)
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
