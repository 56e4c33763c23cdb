"""Times the fused bilinear up-scale (dfpsr_filter_resize, width and height grow) of the library selected by DFPSR_LIB:
4096^2 -> 8192^2 (the bench's figure), 1920x1080 -> 3840x2160 and 640x480 -> 1600x1200. usage: python tools/resize_time.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from dfpsr_b200 import abi, lib

cuda = lib.load(); lib.check(cuda.dfpsr_init(0))
s = lib.stream_ptr()
for (sw, sh, tw, th) in ((4096, 4096, 8192, 8192), (1920, 1080, 3840, 2160), (640, 480, 1600, 1200)):
    src = torch.randint(-2**31, 2**31 - 1, (sh, sw), dtype=torch.int32, device="cuda")
    dst = torch.empty((th, tw), dtype=torch.int32, device="cuda")
    call = lambda: lib.check(cuda.dfpsr_filter_resize(C.byref(lib.image(dst)), C.byref(lib.image(src)), abi.SAMPLER_LINEAR, 0, None, s))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        call()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    gbs = 4 * (sw * sh + tw * th) / ms / 1e6
    print(f"{os.environ.get('DFPSR_LIB', 'default')}: {sw}x{sh}->{tw}x{th} {ms * 1000:.1f} us {gbs:.0f} GB/s", flush=True)
