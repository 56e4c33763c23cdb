"""Rebuilds the library with -D tuning macros (on the GPU box) and times the 4096^2 -> 8192^2 bilinear up-scale for each.
usage: python tools/resize_sweep.py "-DRESIZE_UP_MIN_BLOCKS=6" ..."""
import os, subprocess, sys
for flags in sys.argv[1:]:
    env = dict(os.environ, DFPSR_NVCC_EXTRA=flags)
    subprocess.check_call([sys.executable, "-c", "from dfpsr_b200 import build; build.build(force=True)"], env=env)
    code = '''
import sys
sys.path[:0]=["/root/repo","/root/repo/tests"]
from dfpsr_b200 import lib
import bench_extras
cuda=lib.load(); lib.check(cuda.dfpsr_init(0))
out=bench_extras.run(cuda, lib, cpu=False)
print(round(out["filter_chain_8192"]["resize_up_ms"],4))
'''
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    print(flags, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
