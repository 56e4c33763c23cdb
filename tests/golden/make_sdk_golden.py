"""Golden fixtures from the SDK's REAL terrain media: SDK/terrain/media/{HeightMap,Cloud,RampIsland}.png go through the
SDK's own scene generators (SDK/terrain/main.cpp:232-373, compiled where they lie into oracle/_ref by oracle/Makefile) and the unmodified reference
renders the resulting model at 1920x1080. Needs /root/reference.

Run in the build container:  python tests/golden/make_sdk_golden.py
  tests/golden/sdk_terrain_scene.npz   the generated scene (points, polygons, level 0 of the colour texture), so that the tests can render
                                       exactly this scene without the reference's media
  tests/golden/sdk_terrain.json        sha256 of the reference's colour and depth buffers for a few frames of the SDK's camera orbit"""
import ctypes as C
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

MEDIA = "/root/reference/Source/SDK/terrain/media"
W, H = 1920, 1080


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


ref = refbind.Ref("scalar")
texture_id = C.c_int(-1)
model = ref.lib.ref_sdk_terrain_scene(MEDIA.encode(), C.byref(texture_id))
points, polygons, parts, filter_, names = ref.dump_model(model)
pixels, info = ref.texture_pixels(texture_id.value)
width, height, levels = 1 << int(info[0]), 1 << int(info[1]), int(info[2]) + 1
level0 = pixels[len(pixels) - width * height:].reshape(height, width)  # the pyramid is stored smallest level first
out = {"width": W, "height": H, "points": int(len(points)), "polygons": int(len(polygons)), "texture": [width, height, levels], "frames": []}
for frame in (0, 17, 43):
    col, dep = ref.rgba(array=np.zeros((H, W), np.uint32)), ref.f32(array=np.zeros((H, W), np.float32))
    ref.render(model, scenes.orbit_camera(frame, W, H), col, dep, mode=1)  # the SDK's orbit: offset (sin t, 1, cos t) * 10 about (32, 0, -32), t = 2 pi frame / 60
    c, d = ref.read_rgba(col), ref.read_f32(dep)
    out["frames"].append({"frame": frame, "color_sha256": sha(c), "depth_sha256": sha(d), "covered": float((d > 0).mean())})
    print(frame, out["frames"][-1]["color_sha256"][:16], out["frames"][-1]["covered"], flush=True)
np.savez_compressed(os.path.join(HERE, "sdk_terrain_scene.npz"), points=points.astype(np.float32), polygons=polygons, texture=level0.astype(np.uint32), levels=np.int32(levels))
json.dump(out, open(os.path.join(HERE, "sdk_terrain.json"), "w"), indent=1, sort_keys=True)
print(out["points"], "points", out["polygons"], "polygons", out["texture"])
