"""One 1080p terrain frame at a time (BASELINE config 1): wall clock per frame, per-kernel device time, host time inside the call."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from dfpsr_b200 import abi, lib, scenes  # noqa: E402

cuda = lib.load()
lib.check(cuda.dfpsr_init(0))
sc = scenes.terrain_scene()
tex = lib.DeviceTexture(sc["texture"], 5)
model = lib.DeviceModel(sc["points"], sc["polygons"], abi.FILTER_SOLID, tex)
color = torch.empty((1080, 1920), dtype=torch.int32, device="cuda")
depth = torch.empty((1080, 1920), dtype=torch.float32, device="cuda")
cams = (abi.Camera * 1)(lib.camera(scenes.orbit_camera(7, 1920, 1080)))
ci, di = (abi.Image * 1)(lib.image(color)), (abi.Image * 1)(lib.image(depth))
ident = abi.Transform3D.identity()
s = lib.stream_ptr()
call = lambda: lib.check(cuda.dfpsr_model_render_views(C.byref(model.desc), C.byref(ident), ci, di, cams, 1, 1, s))
for _ in range(20):
    call()
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
host = 0.0
for _ in range(n):
    h0 = time.perf_counter()
    call()
    host += time.perf_counter() - h0
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n
lib.check(cuda.dfpsr_profile_reset())
lib.check(cuda.dfpsr_profile_enable(1))
for _ in range(20):
    call()
torch.cuda.synchronize()
lib.check(cuda.dfpsr_profile_enable(0))
prof = lib.profile_snapshot()
print(f"wall {1e6 * wall:.1f} us per frame ({1 / wall:.0f} fps), host time inside the call {1e6 * host / n:.1f} us")
print("kernels per frame: " + ", ".join(f"{k} {1000 * ms / 20:.1f}us" for k, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])), "| sum", sum(1000 * ms / 20 for ms, c in prof.values()))
