"""Per-kernel device time of BASELINE config 3 (2 M tiny vertex-coloured triangles at 3840x2160). Run on a GPU box."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from dfpsr_b200 import abi, lib, scenes  # noqa: E402

cuda = lib.load()
lib.check(cuda.dfpsr_init(0))
nx, nz = 1000, 999
tiny = scenes.tiny_triangle_scene(nx, nz)
model = lib.DeviceModel(tiny["points"], tiny["polygons"])
color = torch.empty((2160, 3840), dtype=torch.int32, device="cuda")
depth = torch.empty((2160, 3840), dtype=torch.float32, device="cuda")
cams = (abi.Camera * 1)(lib.camera(scenes.top_down_camera(nx, nz, 3840, 2160)))
ci, di = (abi.Image * 1)(lib.image(color)), (abi.Image * 1)(lib.image(depth))
ident = abi.Transform3D.identity()
s = lib.stream_ptr()
call = lambda: lib.check(cuda.dfpsr_model_render_views(C.byref(model.desc), C.byref(ident), ci, di, cams, 1, 1, s))
for _ in range(5):
    call()
torch.cuda.synchronize()
lib.check(cuda.dfpsr_profile_reset())
lib.check(cuda.dfpsr_profile_enable(1))
for _ in range(10):
    call()
torch.cuda.synchronize()
lib.check(cuda.dfpsr_profile_enable(0))
prof = lib.profile_snapshot()
print("kernels per frame: " + ", ".join(f"{k} {1000 * ms / 10:.1f}us" for k, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])), "| sum", round(sum(1000 * ms / 10 for ms, c in prof.values()), 1))
