"""A/B timing of the tile kernel's modes on the headline workload (256 x 1080p terrain views per launch), one 1080p frame at a time and
the 2 M tiny-triangle frame: per-kernel device time per frame (CUDA events around every launch, dfpsr_profile_*).
usage: python tools/tile_ab.py [views]        DFPSR_TILE_MODE=immediate selects the round-1 kernel (set per run: it is read once)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from dfpsr_b200 import abi, lib, scenes  # noqa: E402

views = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cuda = lib.load()
lib.check(cuda.dfpsr_init(0))
sc = scenes.terrain_scene()
tex = lib.DeviceTexture(sc["texture"], 5)
model = lib.DeviceModel(sc["points"], sc["polygons"], abi.FILTER_SOLID, tex)
W, H = 1920, 1080
color = torch.empty((views, H, W), dtype=torch.int32, device="cuda")
depth = torch.empty((views, H, W), dtype=torch.float32, device="cuda")
cams = (abi.Camera * views)(*[lib.camera(scenes.orbit_camera(v, W, H, frames_per_lap=views)) for v in range(views)])
ci = (abi.Image * views)(*[lib.image(color[v]) for v in range(views)])
di = (abi.Image * views)(*[lib.image(depth[v]) for v in range(views)])
ident = abi.Transform3D.identity()
s = lib.stream_ptr()


def profile(call, repeats, frames):
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(repeats):
        call()
    stop.record()
    torch.cuda.synchronize()
    wall = start.elapsed_time(stop) * 1000.0 / (repeats * frames)
    lib.check(cuda.dfpsr_profile_reset())
    lib.check(cuda.dfpsr_profile_enable(1))
    for _ in range(repeats):
        call()
    torch.cuda.synchronize()
    lib.check(cuda.dfpsr_profile_enable(0))
    prof = lib.profile_snapshot()
    per = {k: 1000.0 * ms / (repeats * frames) for k, (ms, n) in prof.items()}
    return wall, per


for name, precision in (("exact", 0), ("tolerance", 1)):
    lib.check(cuda.dfpsr_set_default_precision(precision))
    wall, per = profile(lambda: lib.check(cuda.dfpsr_model_render_views(C.byref(model.desc), C.byref(ident), ci, di, cams, views, 1, s)), 3, views)
    print(f"[{name}] batch of {views}: {wall:.2f} us/frame ({1e6 / wall:.0f} fps) | " + ", ".join(f"{k} {v:.2f}" for k, v in sorted(per.items(), key=lambda kv: -kv[1])), flush=True)
    wall, per = profile(lambda: lib.check(cuda.dfpsr_model_render_views(C.byref(model.desc), C.byref(ident), ci, di, cams, 1, 1, s)), 50, 1)
    print(f"[{name}] single frame: {wall:.1f} us/frame | " + ", ".join(f"{k} {v:.1f}" for k, v in sorted(per.items(), key=lambda kv: -kv[1])), flush=True)

if "--tiny" in sys.argv:
    del color, depth
    nx, nz = 1000, 999
    ts = scenes.tiny_triangle_scene(nx, nz)
    tmodel = lib.DeviceModel(ts["points"], ts["polygons"])
    TW, TH = 3840, 2160
    tc, td = torch.empty((TH, TW), dtype=torch.int32, device="cuda"), torch.empty((TH, TW), dtype=torch.float32, device="cuda")
    tcam = (abi.Camera * 1)(lib.camera(scenes.top_down_camera(nx, nz, TW, TH)))
    tci, tdi = (abi.Image * 1)(lib.image(tc)), (abi.Image * 1)(lib.image(td))
    for name, precision in (("exact", 0), ("tolerance", 1)):
        lib.check(cuda.dfpsr_set_default_precision(precision))
        wall, per = profile(lambda: lib.check(cuda.dfpsr_model_render_views(C.byref(tmodel.desc), C.byref(ident), tci, tdi, tcam, 1, 1, s)), 10, 1)
        print(f"[{name}] 2M tiny triangles 4K: {wall:.1f} us | " + ", ".join(f"{k} {v:.1f}" for k, v in sorted(per.items(), key=lambda kv: -kv[1])), flush=True)
