"""Golden fixtures from the Sandbox's REAL sprite media: SDK/sandbox/media/images/{Floor,Pillar,WoodenBarrel}.png + .ini
are loaded by the unmodified reference's own loader (spriteWorld_loadSpriteTypeFromFile), placed in a small lit world and drawn by the
unmodified reference (spriteWorld_draw, two frames). Needs /root/reference.

Run in the build container:  python tests/golden/make_sandbox_media_golden.py
  tests/golden/sandbox_media.npz    the decoded atlases and the numbers of the .ini files exactly as the reference parsed them (its own
                                    decimal parser, not Python's), so that the tests can build the same sprite types without the media
  tests/golden/sandbox_media.json   the script and sha256 of the reference's colour / diffuse / normal / light / height buffers per frame"""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import refbind  # noqa: E402
import sprite_world_scene as sws  # noqa: E402

MEDIA = "/root/reference/Source/SDK/sandbox/media/images"
NAMES = ["Floor", "Pillar", "WoodenBarrel"]
F = np.float32

ref = refbind.Ref("scalar")
to_float = lambda text: F(ref.lib.ref_string_to_double(text.strip().encode()))


def parse_ini(path):
    """Key=value lines (ref: SDK/SpriteEngine/spriteAPI.cpp:56-101); every decimal goes through the reference's string_toDouble."""
    cfg = {}
    for line in open(path, encoding="utf-8-sig"):
        line = line.strip()
        if not line or line.startswith(";") or "=" not in line:
            continue
        key, value = line.split("=", 1)
        cfg[key.strip().lower()] = value
    t = {"center": (int(cfg["centerx"]), int(cfg["centery"])), "frames": int(cfg["framerows"]), "columns": int(cfg["propertycolumns"]),
         "min": [to_float(v) for v in cfg["minbound"].split(",")], "max": [to_float(v) for v in cfg["maxbound"].split(",")], "points": None, "indices": None}
    if "points" in cfg:
        t["points"] = np.array([to_float(v) for v in cfg["points"].split(",")], F).reshape(-1, 3)
        t["indices"] = np.array([int(v) for v in cfg["triangleindices"].split(",")], np.int32)
    return t


sprites, ids = [], []
for name in NAMES:
    t = parse_ini(os.path.join(MEDIA, name + ".ini"))
    assert t["columns"] == 3
    image = ref.lib.ref_image_load(os.path.join(MEDIA, name + ".png").encode())
    assert image >= 0, name
    t["atlas"] = ref.read_rgba(image)
    t["frame_w"], t["frame_h"] = t["atlas"].shape[1] // 3, t["atlas"].shape[0] // t["frames"]
    sprites.append(t)
    ids.append(ref.lib.ref_sprite_type_load(MEDIA.encode(), name.encode()))
    print(name, t["atlas"].shape, t["frames"], "frames", t["center"], t["min"], t["max"], 0 if t["points"] is None else len(t["points"]), "shadow points", flush=True)
assets = {"sprites": sprites, "models": []}

MINI = sws.MINI
rng = np.random.default_rng(41)
script = []
for gx in range(-3, 4):
    for gz in range(-3, 4):
        script.append(("bg_sprite", 0, int(rng.integers(0, 4)), (gx * MINI, 0, gz * MINI), 0))
for k in range(6):
    script.append(("bg_sprite", 1, int(rng.integers(0, 8)), (int(rng.integers(-2, 3)) * MINI + 300, 0, int(rng.integers(-2, 3)) * MINI - 200), 1))
for k in range(9):
    script.append(("bg_sprite", 2, int(rng.integers(0, 8)), (int(rng.integers(-2500, 2500)), 0, int(rng.integers(-2500, 2500))), 1))
lights = [("directed", (1.0, -1.0, 0.0), 0.1, (255, 255, 255)), ("point", (0.9, 1.1, 0.4), 4.0, 1.0, (255, 170, 100), 1), ("point", (-1.6, 0.8, -1.2), 3.5, 0.9, (110, 190, 255), 1)]
script += lights + [("tmp_sprite", 2, 5, (450, 0, 380), 1), ("draw", 400, 300)]
script += [("clear_temporary",), ("move_camera", 37, -21)] + lights + [("tmp_sprite", 2, 6, (-700, 0, 150), 1), ("draw", 400, 300)]

with tempfile.TemporaryDirectory() as folder:
    frames = sws.run_reference(ref, assets, script, folder, sprite_ids=ids)
out = {"names": NAMES, "script": script, "frames": [sws.frame_hashes(f) for f in frames]}
for f in frames:
    print({k: v[:12] if isinstance(v, str) else v for k, v in sws.frame_hashes(f).items()}, "lit pixels", int((f["light"] & 0xFFFFFF != 0).sum()), flush=True)
arrays = {}
for name, t in zip(NAMES, sprites):
    arrays[name + "_atlas"] = t["atlas"].astype(np.uint32)
    arrays[name + "_numbers"] = np.array([t["center"][0], t["center"][1], t["frames"]], np.int32)
    arrays[name + "_bounds"] = np.array(t["min"] + t["max"], F)
    if t["points"] is not None:
        arrays[name + "_points"], arrays[name + "_indices"] = t["points"], t["indices"]
np.savez_compressed(os.path.join(HERE, "sandbox_media.npz"), **arrays)
json.dump(out, open(os.path.join(HERE, "sandbox_media.json"), "w"), indent=1, sort_keys=True)
print("wrote", os.path.getsize(os.path.join(HERE, "sandbox_media.npz")), "bytes of atlases")
