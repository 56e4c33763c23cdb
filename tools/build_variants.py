"""Builds tuning variants of the CUDA library HERE (nvcc cross-compiles without a GPU) so that one gpurun call can time them all:
python tools/build_variants.py name1="-DX=1 -DY=2" name2="..."  ->  dfpsr_b200/variants/libdfpsr_b200_<name>.so (git-ignored, travels to the box).
Select one at run time with DFPSR_LIB=dfpsr_b200/variants/libdfpsr_b200_<name>.so."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dfpsr_b200 import build  # noqa: E402

out_dir = os.path.join(ROOT, "dfpsr_b200", "variants")
os.makedirs(out_dir, exist_ok=True)


def one(spec):
    name, flags = spec.split("=", 1)
    target = os.path.join(out_dir, f"libdfpsr_b200_{name}.so")
    subprocess.check_call([build.NVCC] + build.FLAGS + flags.split() + ["-o", target] + build.sources())
    return target


with ThreadPoolExecutor(4) as pool:
    for path in pool.map(one, sys.argv[1:]):
        print(path)
