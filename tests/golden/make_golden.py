"""Generates the golden fixtures in tests/golden/ from the compiled, UNMODIFIED reference
(oracle/_ref/libdfpsr_ref_scalar.so — `make -C oracle ref`, needs /root/reference).

Run in the build container:  python tests/golden/make_golden.py
The fixtures are hashes (sha256 of the raw little-endian pixel rows, tightly packed) of the reference's output on
the deterministic synthetic scenes of dfpsr_b200/scenes.py, plus a few small raw vectors. They pin
  * the C oracle (tests/test_oracle_golden.py, CPU) and
  * the CUDA path at BASELINE.json's full sizes (tests/test_gpu_*.py::*_golden, GPU)
to what the reference itself produces, without /root/reference being present at test time.
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refbind  # noqa: E402
from dfpsr_b200 import abi, scenes  # noqa: E402
import sandbox_scene  # noqa: E402
import sprite_world_scene  # noqa: E402
import draw_scene  # noqa: E402
import mono_scene  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def raster(ref):
    out = {}
    sc = scenes.terrain_scene()
    tex = ref.texture(sc["texture"], 5)
    model = ref.model(sc["points"], sc["polygons"], diffuse=tex)
    entries = []
    for frame in (0, 15, 30, 45):
        col, dep = ref.rgba(shape=(1080, 1920)), ref.f32(shape=(1080, 1920))
        ref.render(model, scenes.orbit_camera(frame, 1920, 1080), col, dep, mode=1)
        entries.append({"frame": frame, "color_sha256": sha(ref.read_rgba(col)), "depth_sha256": sha(ref.read_f32(dep)),
                        "covered": float((ref.read_f32(dep) > 0).mean())})
        ref.free_all()
        tex = ref.texture(sc["texture"], 5)
        model = ref.model(sc["points"], sc["polygons"], diffuse=tex)
    out["terrain_1080p"] = entries
    ref.free_all()
    nx, nz = 1000, 999
    sc = scenes.tiny_triangle_scene(nx, nz)
    model = ref.model(sc["points"], sc["polygons"])
    col, dep = ref.rgba(shape=(2160, 3840)), ref.f32(shape=(2160, 3840))
    ref.render(model, scenes.top_down_camera(nx, nz, 3840, 2160), col, dep, mode=1)
    out["tiny_4k"] = {"nx": nx, "nz": nz, "color_sha256": sha(ref.read_rgba(col)), "depth_sha256": sha(ref.read_f32(dep)),
                      "covered": float((ref.read_f32(dep) > 0).mean())}
    ref.free_all()
    return out


def filters(ref):
    out = {}
    size = 8192
    src = ref.lib.ref_image_create_rgba(size, size, abi.PACK_RGBA)
    ref.lib.ref_filter_map(src, abi.MAP_XOR_PATTERN, None, -1, 0, 0)
    out["source_sha256"] = sha(ref.read_rgba(src))
    mapped = ref.lib.ref_image_create_rgba(size, size, abi.PACK_RGBA)
    params = np.array(sandbox_scene.CHAIN_AFFINE, np.int32)
    ref.lib.ref_filter_map(mapped, abi.MAP_AFFINE, refbind.ptr(params), src, 0, 0)
    out["mapped_sha256"] = sha(ref.read_rgba(mapped))
    for name, (w, h) in {"down_4096": (4096, 4096), "odd_5000x3000": (5000, 3000)}.items():
        r = ref.lib.ref_filter_resize(mapped, abi.SAMPLER_LINEAR, w, h)
        out[name + "_sha256"] = sha(ref.read_rgba(r))
    half = ref.lib.ref_filter_resize(mapped, abi.SAMPLER_LINEAR, 4096, 4096)
    up = ref.lib.ref_filter_resize(half, abi.SAMPLER_LINEAR, 8192, 8192)
    out["up_8192_sha256"] = sha(ref.read_rgba(up))
    near = ref.lib.ref_filter_resize(mapped, abi.SAMPLER_NEAREST, 3000, 5000)
    out["nearest_3000x5000_sha256"] = sha(ref.read_rgba(near))
    ref.free_all()
    return {"filter_chain_8192": out}


def sandbox(ref):
    sb = sandbox_scene.build(800, 600, lights=16, seed=5)
    result = sandbox_scene.run_reference(ref, sb)
    out = {"light_sha256": sha(result["light"]), "color_sha256": sha(result["color"]), "cube0_sha256": sha(result["cubes"][0]),
           "lit_pixels": float((result["light"] & 0xFFFFFF != 0).mean())}
    ref.free_all()
    return {"sandbox_800x600_16": out}


def draw(ref):
    """The 2D draw call sequences of tests/draw_scene.py through the reference's drawAPI."""
    cases = []
    for case in draw_scene.CASES:
        color, depth = draw_scene.run_reference(ref, draw_scene.build(*case))
        cases.append({"color_sha256": sha(color), "depth_sha256": sha(depth)})
        ref.free_all()
    mono = []
    for seed in mono_scene.SEEDS:
        mono.append({"seed": seed, "sha256": mono_scene.sha(mono_scene.run_reference(ref, mono_scene.build(seed)))})
        ref.free_all()
    return {"cases": cases, "mono": mono}


def sprite_world(ref):
    """The scripted Sandbox session of tests/sprite_world_scene.py through the reference's spriteWorld_* API, plus renderDenseModel alone."""
    import tempfile
    assets, script = sprite_world_scene.build_assets(), sprite_world_scene.build_script()
    frames = sprite_world_scene.run_reference(ref, assets, script, tempfile.mkdtemp(prefix="dfpsr_golden_"))
    out = {"script_frames": [dict(sprite_world_scene.frame_hashes(f), camera=[int(v) for v in f["camera"]], ground=[int(v) for v in f["ground"]],
                                  covered=float((f["height"] > -1e5).mean())) for f in frames]}
    out["dense"] = [dict(case=case, **sprite_world_scene.frame_hashes_dense(sprite_world_scene.dense_reference(ref, assets, case))) for case in range(len(sprite_world_scene.DENSE_CASES))]
    out["bake"] = []
    for case in range(len(sprite_world_scene.BAKE_CASES)):
        baked = sprite_world_scene.bake_reference(ref, assets, case)
        out["bake"].append({"atlas_sha256": sha(baked["atlas"]), "shape": list(baked["atlas"].shape), "config": baked["config"], "opaque": float((baked["atlas"] >> 24 != 0).mean())})
    ref.free_all()
    return out


if __name__ == "__main__":
    ref = refbind.Ref("scalar")
    which = sys.argv[1:] or ["raster", "filters", "sandbox", "sprite_world", "draw"]
    for name in which:
        data = {"raster": raster, "filters": filters, "sandbox": sandbox, "sprite_world": sprite_world, "draw": draw}[name](ref)
        path = os.path.join(HERE, name + ".json")
        json.dump(data, open(path, "w"), indent=1, sort_keys=True)
        print("wrote", path)
