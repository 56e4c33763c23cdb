"""GPU: the Sandbox sprite engine on the device — renderDenseModel and whole spriteWorld_draw frames
through the C ABI, bit-exact against (a) the host planner replayed with the C oracle and (b) hashes of the reference's buffers."""
import json
import os

import numpy as np
import pytest

import sprite_world_scene as sws
from dfpsr_b200 import lib

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sprite_world.json")
BUFFERS = ("color", "diffuse", "normal", "light", "height")


@pytest.fixture(scope="module")
def assets():
    return sws.build_assets()


@pytest.mark.parametrize("case", range(len(sws.DENSE_CASES)))
def test_dense_model_render_bit_exact(cuda, oracle, assets, case):
    expected = sws.dense_oracle(oracle, cuda, lib.check, assets, case)
    got = sws.dense_cuda(cuda, lib, assets, case)
    assert np.array_equal(got["rect"], expected["rect"])
    for name in ("height", "diffuse", "normal"):
        assert np.array_equal(got[name].view(np.uint32), expected[name].view(np.uint32)), name
    entry = json.load(open(GOLDEN))["dense"][case]
    hashes = sws.frame_hashes_dense(got)
    assert hashes == {k: entry[k] for k in hashes}


def test_sprite_world_session_bit_exact(cuda, oracle, assets):
    script = sws.build_script()
    cuda.dfpsr_reset_launch_count()
    got = sws.run_cuda(cuda, lib, assets, script)
    launches = cuda.dfpsr_launch_count()
    expected = sws.run_plan_oracle(cuda, lib.check, oracle, assets, script)
    golden = json.load(open(GOLDEN))["script_frames"]
    assert len(got) == len(expected) == len(golden)
    for index, (a, b, entry) in enumerate(zip(got, expected, golden)):
        for name in BUFFERS:
            assert np.array_equal(a[name].view(np.uint32), b[name].view(np.uint32)), (index, name)
        assert sws.frame_hashes(a) == {k: entry[k] for k in BUFFERS}, index
        assert [int(v) for v in a["camera"]] == entry["camera"]
    assert launches > 0


def media_assets():
    """The Sandbox's real sprite media as decoded and parsed by the reference (tests/golden/make_sandbox_media_golden.py)."""
    golden_dir = os.path.dirname(GOLDEN)
    data = np.load(os.path.join(golden_dir, "sandbox_media.npz"))
    entry = json.load(open(os.path.join(golden_dir, "sandbox_media.json")))
    sprites = []
    for name in entry["names"]:
        atlas, numbers, bounds = data[name + "_atlas"], data[name + "_numbers"], data[name + "_bounds"]
        frames = int(numbers[2])
        t = {"atlas": atlas, "frame_w": atlas.shape[1] // 3, "frame_h": atlas.shape[0] // frames, "frames": frames, "center": (int(numbers[0]), int(numbers[1])),
             "min": [np.float32(v) for v in bounds[:3]], "max": [np.float32(v) for v in bounds[3:]], "points": None, "indices": None}
        if name + "_points" in data:
            t["points"], t["indices"] = np.ascontiguousarray(data[name + "_points"], np.float32), np.ascontiguousarray(data[name + "_indices"], np.int32)
        sprites.append(t)
    script = [tuple(tuple(v) if isinstance(v, list) else v for v in action) for action in entry["script"]]
    return {"sprites": sprites, "models": []}, script, entry["frames"]


def test_sandbox_real_media_golden(cuda, oracle):
    """SDK/sandbox/media/images/{Floor,Pillar,WoodenBarrel} (real atlases, real .ini numbers, real shadow shapes) in a lit world: every
    buffer of both frames equals what the unmodified reference drew from the files themselves, and the oracle replay agrees."""
    media, script, golden = media_assets()
    got = sws.run_cuda(cuda, lib, media, script)
    expected = sws.run_plan_oracle(cuda, lib.check, oracle, media, script)
    assert len(got) == len(golden) == 2
    for index, (a, b, entry) in enumerate(zip(got, expected, golden)):
        for name in BUFFERS:
            assert np.array_equal(a[name].view(np.uint32), b[name].view(np.uint32)), (index, name)
        assert sws.frame_hashes(a) == {k: entry[k] for k in BUFFERS}, index


def test_sprite_world_draw_host_round_trip(cuda, assets):
    """dfpsr_sprite_world_draw_host returns the same colour image as the device call."""
    import ctypes as C
    import torch
    script = [a for a in sws.build_script() if a[0] != "draw"][:140]
    pw = sws.ProductWorld(cuda, lib.check, assets)
    for action in script:
        pw.apply(action)
    pw.apply(("point", (0.5, 1.0, 0.5), 4.0, 1.0, (255, 220, 200), 1))
    w, h = 333, 201  # odd sizes: row strides are padded on the device
    host = np.zeros((h, w), np.uint32)
    lib.check(cuda.dfpsr_sprite_world_draw_host(pw.world, host.ctypes.data, w * 4, w, h, 0, lib.stream_ptr()))
    device = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    lib.check(cuda.dfpsr_sprite_world_draw(pw.world, C.byref(lib.image(device)), lib.stream_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(host, device.cpu().numpy().view(np.uint32))
    assert (host != 0).mean() > 0.05  # one point light: most of the frame stays black
    pw.close()


@pytest.mark.parametrize("case", range(len(sws.BAKE_CASES)))
def test_sprite_generate_from_model_matches_reference_golden(cuda, assets, case):
    """sprite_generateFromModel on the device: atlas (colour | 8-bit height | normal per camera angle, cropped) and sprite centre equal the
    reference's (tests/golden/sprite_world.json, generated by the compiled reference)."""
    entry = json.load(open(GOLDEN))["bake"][case]
    baked = sws.bake_cuda(cuda, lib, assets, case)
    assert list(baked["atlas"].shape) == entry["shape"] and baked["config"] == entry["config"]
    assert sws.sha(baked["atlas"]) == entry["atlas_sha256"]
    assert entry["opaque"] > 0.1
