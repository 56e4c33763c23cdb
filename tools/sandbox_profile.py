"""Per-kernel device time of one Sandbox frame (800x600, 16 shadowed point lights) through the library's profiling hooks."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, time
import sandbox_scene
from dfpsr_b200 import lib
cuda = lib.load(); lib.check(cuda.dfpsr_init(0))
sb = sandbox_scene.build(800, 600, lights=16, seed=5)
gpu = sandbox_scene.CudaSandbox(cuda, sb)
gpu.composite()
for _ in range(3):
    gpu.light_fused()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20):
    gpu.light_fused()
torch.cuda.synchronize()
print("wall ms per frame (light passes)", (time.perf_counter() - t) / 20 * 1e3)
lib.check(cuda.dfpsr_profile_reset()); lib.check(cuda.dfpsr_profile_enable(1))
gpu.composite(); gpu.light_fused()
torch.cuda.synchronize()
lib.check(cuda.dfpsr_profile_enable(0))
tot = 0
for k, (ms, n) in sorted(lib.profile_snapshot().items(), key=lambda kv: -kv[1][0]):
    print(f"{k:40s} {n:4d} launches {ms*1e3:9.1f} us"); tot += ms
print("sum of kernels us", tot * 1e3)
