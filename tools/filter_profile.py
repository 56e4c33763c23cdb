"""BASELINE config 5 kernels (filter_mapRgbaU8 affine op, bilinear half / double resize) at 8192x8192: device time and GB/s. Run on a GPU box."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench_extras  # noqa: E402
import sandbox_scene  # noqa: E402
from dfpsr_b200 import abi, lib  # noqa: E402

cuda = lib.load()
lib.check(cuda.dfpsr_init(0))
IM, s = lib.image, lib.stream_ptr()
size = 8192
src = torch.randint(0, 2 ** 31 - 1, (size, size), dtype=torch.int32, device="cuda")
mapped, half, up = torch.empty_like(src), torch.empty((size // 2, size // 2), dtype=torch.int32, device="cuda"), torch.empty_like(src)
prm = np.array(sandbox_scene.CHAIN_AFFINE, np.int32)
ms = bench_extras._time(torch, lambda: lib.check(cuda.dfpsr_filter_map(C.byref(IM(mapped)), abi.MAP_AFFINE, prm.ctypes.data, 8, C.byref(IM(src)), 0, 0, s)), iters=20)
print(f"map {ms * 1000:.1f} us, {8 * size * size / ms / 1e6:.0f} GB/s")
ms = bench_extras._time(torch, lambda: lib.check(cuda.dfpsr_filter_resize(C.byref(IM(half)), C.byref(IM(mapped)), abi.SAMPLER_LINEAR, 0, None, s)), iters=20)
print(f"resize half {ms * 1000:.1f} us, {4 * (size * size + size * size // 4) / ms / 1e6:.0f} GB/s")
ms = bench_extras._time(torch, lambda: lib.check(cuda.dfpsr_filter_resize(C.byref(IM(up)), C.byref(IM(half)), abi.SAMPLER_LINEAR, 0, None, s)), iters=20)
print(f"resize double {ms * 1000:.1f} us, {4 * (size * size + size * size // 4) / ms / 1e6:.0f} GB/s")
ms = bench_extras._time(torch, lambda: mapped.copy_(src), iters=20)
print(f"torch copy {ms * 1000:.1f} us, {8 * size * size / ms / 1e6:.0f} GB/s")
