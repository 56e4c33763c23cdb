"""CPU: pins the C oracle (oracle/dfpsr_oracle.c) to the golden fixtures in tests/golden/, which were produced by the
compiled, UNMODIFIED reference (tests/golden/make_golden.py, scalar flavour) on the deterministic scenes of
dfpsr_b200/scenes.py at BASELINE.json's full sizes. Runs without /root/reference."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import orcbind
import sandbox_scene
from dfpsr_b200 import abi, scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden(name):
    return json.load(open(os.path.join(GOLDEN, name + ".json")))


@pytest.fixture(scope="module")
def terrain(oracle):
    sc = scenes.terrain_scene()
    buf, tex = orcbind.build_texture(sc["texture"], 5)
    model, keep = orcbind.model_of(sc["points"], sc["polygons"], diffuse=tex)
    return model, (buf, keep)


@pytest.mark.parametrize("index", range(4))
def test_terrain_1080p(oracle, terrain, index):
    entry = golden("raster")["terrain_1080p"][index]
    model, _ = terrain
    c, d = np.zeros((1080, 1920), np.uint32), np.zeros((1080, 1920), np.float32)
    ident = abi.Transform3D.identity()
    cam = orcbind.camera(scenes.orbit_camera(entry["frame"], 1920, 1080))
    n = oracle.orc_model_render(C.byref(model), C.byref(ident), C.byref(orcbind.image_of(c)), C.byref(orcbind.image_of(d)), C.byref(cam))
    assert n > 1000
    assert abs(float((d > 0).mean()) - entry["covered"]) < 1e-12
    assert sha(d) == entry["depth_sha256"]
    assert sha(c) == entry["color_sha256"]


def test_tiny_triangles_4k(oracle):
    g = golden("raster")["tiny_4k"]
    sc = scenes.tiny_triangle_scene(g["nx"], g["nz"])
    model, keep = orcbind.model_of(sc["points"], sc["polygons"])
    c, d = np.zeros((2160, 3840), np.uint32), np.zeros((2160, 3840), np.float32)
    ident = abi.Transform3D.identity()
    cam = orcbind.camera(scenes.top_down_camera(g["nx"], g["nz"], 3840, 2160))
    oracle.orc_model_render(C.byref(model), C.byref(ident), C.byref(orcbind.image_of(c)), C.byref(orcbind.image_of(d)), C.byref(cam))
    assert sha(d) == g["depth_sha256"]
    assert sha(c) == g["color_sha256"]


def test_sandbox_800x600_16_lights(oracle):
    g = golden("sandbox")["sandbox_800x600_16"]
    sb = sandbox_scene.build(800, 600, lights=16, seed=5)
    result = sandbox_scene.run_oracle(oracle, sb)
    assert sha(result["cubes"][0]) == g["cube0_sha256"]
    assert sha(result["light"]) == g["light_sha256"]
    assert sha(result["color"]) == g["color_sha256"]


def test_filter_chain_8192(oracle):
    g = golden("filters")["filter_chain_8192"]
    IM = orcbind.image_of
    size = 8192
    src = np.zeros((size, size), np.uint32)
    oracle.orc_filter_map(C.byref(IM(src)), abi.MAP_XOR_PATTERN, None, None, 0, 0)
    assert sha(src) == g["source_sha256"]
    mapped = np.zeros_like(src)
    params = np.array(sandbox_scene.CHAIN_AFFINE, np.int32)
    oracle.orc_filter_map(C.byref(IM(mapped)), abi.MAP_AFFINE, params.ctypes.data, C.byref(IM(src)), 0, 0)
    assert sha(mapped) == g["mapped_sha256"]
    del src

    def resize(source, w, h, sampler):
        out = np.zeros((h, w), np.uint32)
        scratch = np.zeros(w * source.shape[0], np.uint32)
        oracle.orc_filter_resize(C.byref(IM(out)), C.byref(IM(source)), sampler, 0, scratch.ctypes.data)
        return out

    half = resize(mapped, 4096, 4096, abi.SAMPLER_LINEAR)
    assert sha(half) == g["down_4096_sha256"]
    assert sha(resize(mapped, 5000, 3000, abi.SAMPLER_LINEAR)) == g["odd_5000x3000_sha256"]
    assert sha(resize(mapped, 3000, 5000, abi.SAMPLER_NEAREST)) == g["nearest_3000x5000_sha256"]
    assert sha(resize(half, 8192, 8192, abi.SAMPLER_LINEAR)) == g["up_8192_sha256"]
