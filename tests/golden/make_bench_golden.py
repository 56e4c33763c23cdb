"""Golden hashes for bench.py's parity check: the first and last view of every rank's shard of the 256-view terrain batch (BASELINE config 4,
sharded over 1 / 2 / 4 / 8 GPUs), rendered by the compiled, UNMODIFIED reference (oracle/_ref/libdfpsr_ref_scalar.so, needs /root/reference).

Run in the build container:  python tests/golden/make_bench_golden.py   ->  tests/golden/bench_views.json"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import refbind  # noqa: E402
from dfpsr_b200 import scenes  # noqa: E402

VIEWS, W, H = 256, 1920, 1080


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


ref = refbind.Ref("scalar")
sc = scenes.terrain_scene()
out = {"views_per_lap": VIEWS, "width": W, "height": H, "views": {}}
for view in sorted({k * 32 for k in range(8)} | {k * 32 + 31 for k in range(8)}):
    tex = ref.texture(sc["texture"], 5)
    model = ref.model(sc["points"], sc["polygons"], diffuse=tex)
    col, dep = ref.rgba(array=np.zeros((H, W), np.uint32)), ref.f32(array=np.zeros((H, W), np.float32))
    ref.render(model, scenes.orbit_camera(view, W, H, frames_per_lap=VIEWS), col, dep, mode=1)
    out["views"][str(view)] = {"color_sha256": sha(ref.read_rgba(col)), "depth_sha256": sha(ref.read_f32(dep))}
    ref.free_all()
    print(view, out["views"][str(view)]["color_sha256"][:16], flush=True)
json.dump(out, open(os.path.join(HERE, "bench_views.json"), "w"), indent=1, sort_keys=True)
