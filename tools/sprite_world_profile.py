"""Where a spriteWorld_draw frame goes: host planning time, per-kernel device time (CUDA events), wall clock. Run on a GPU box."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import torch  # noqa: E402

import sprite_world_scene as sws  # noqa: E402
from dfpsr_b200 import abi, lib  # noqa: E402

cuda = lib.load()
lib.check(cuda.dfpsr_init(0))
assets = sws.build_assets()
script = sws.sandbox_script(800, 600, lights=16, frames=14)
pw, planner = sws.ProductWorld(cuda, lib.check, assets, shadow_res=256), sws.ProductWorld(cuda, lib.check, assets, shadow_res=256)
target = torch.zeros((600, 800), dtype=torch.int32, device="cuda")
frame = 0
for action in script:
    if action[0] != "draw":
        pw.apply(action)
        planner.apply(action)
        continue
    ops, count = C.POINTER(abi.SpriteWorldOp)(), C.c_int32()
    t0 = time.perf_counter()
    lib.check(cuda.dfpsr_sprite_world_plan_frame(planner.world, 800, 600, C.byref(ops), C.byref(count)))
    plan_ms = 1000.0 * (time.perf_counter() - t0)
    profiled = frame >= 10
    if profiled:
        lib.check(cuda.dfpsr_profile_reset())
        lib.check(cuda.dfpsr_profile_enable(1))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lib.check(cuda.dfpsr_sprite_world_draw(pw.world, C.byref(lib.image(target)), lib.stream_ptr()))
    torch.cuda.synchronize()
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    line = f"frame {frame}: {count.value} ops, host plan {plan_ms:.3f} ms, draw wall {wall_ms:.3f} ms"
    if profiled:
        lib.check(cuda.dfpsr_profile_enable(0))
        prof = lib.profile_snapshot()
        line += " | " + ", ".join(f"{k} {1000 * ms:.0f}us x{n}" for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]))
    print(line)
    frame += 1
