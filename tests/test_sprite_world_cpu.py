"""CPU: the Sandbox sprite engine's HOST side against the compiled reference and the golden fixtures.

  * dfpsr_ortho_system_create == OrthoSystem(cameraTilt, pixelsPerTile)           (bit for bit, every derived matrix)
  * dfpsr_dense_model_build == DenseModel_create                                  (triangles, normals, bounds)
  * oracle renderDenseModel == the reference's                                    (pins the checker the GPU test uses)
  * the product's frame planner (octrees, background blocks, dirty rectangles, pass order) replayed with the oracle's pixel loops
    == spriteWorld_draw of the reference, buffer by buffer, over a scripted session
No compute entry point of the product is called here (no GPU): dfpsr_sprite_world_plan_frame is host logic only.
"""
import ctypes as C
import json
import os
import tempfile

import numpy as np
import pytest

import sprite_world_scene as sws
from dfpsr_b200 import abi, lib

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sprite_world.json")
BUFFERS = ("color", "diffuse", "normal", "light", "height")


@pytest.fixture(scope="module")
def product():
    return lib.load()


@pytest.fixture(scope="module")
def assets():
    return sws.build_assets()


@pytest.mark.parametrize("tilt,pixels", [(-0.6, 150), (-0.6, 64), (-1.0, 64), (-0.3, 200), (-2.5, 33), (-0.75, 128)])
def test_ortho_system_matches_reference(product, ref_scalar, tilt, pixels):
    ours, theirs = abi.OrthoSystem(), abi.OrthoSystem()
    lib.check(product.dfpsr_ortho_system_create(C.byref(ours), tilt, pixels))
    ref_scalar.lib.ref_ortho_system(tilt, pixels, C.byref(theirs))
    assert bytes(ours) == bytes(theirs)


def test_ortho_light_view_is_the_view_the_light_tests_use(product):
    import sandbox_scene
    system = abi.OrthoSystem()
    lib.check(product.dfpsr_ortho_system_create(C.byref(system), -0.6, 150))  # SDK/sandbox/media/Ortho.ini
    view = abi.OrthoView()
    lib.check(product.dfpsr_ortho_camera_light_view(C.byref(system.view[0]), C.byref(view)))
    assert bytes(view) == bytes(sandbox_scene.ortho_view())


def test_dense_model_build_matches_reference(product, ref_scalar, assets):
    for m in assets["models"]:
        tris, mn, mx = sws.dense_build(product, lib.check, m["points"], m["polygons"])
        dense = ref_scalar.lib.ref_dense_model_create(ref_scalar.model(m["points"], m["polygons"]))
        count = ref_scalar.lib.ref_dense_model_triangles(dense, None, None, None)
        assert count == len(tris) > 0
        expected, emn, emx = np.zeros(count, abi.DENSE_TRIANGLE_DTYPE), np.zeros(3, np.float32), np.zeros(3, np.float32)
        ref_scalar.lib.ref_dense_model_triangles(dense, expected.ctypes.data, emn.ctypes.data, emx.ctypes.data)
        assert tris.tobytes() == expected.tobytes()
        assert mn.tobytes() == emn.tobytes() and mx.tobytes() == emx.tobytes()


@pytest.mark.parametrize("case", range(len(sws.DENSE_CASES)))
def test_oracle_dense_model_render_matches_reference(product, oracle, ref_scalar, assets, case):
    expected = sws.dense_reference(ref_scalar, assets, case)
    got = sws.dense_oracle(oracle, product, lib.check, assets, case)
    assert np.array_equal(got["rect"], expected["rect"])
    for name in ("height", "diffuse", "normal"):
        assert np.array_equal(got[name].view(np.uint32), expected[name].view(np.uint32)), name
    culled = case == len(sws.DENSE_CASES) - 1
    assert ((got["diffuse"] != 0).mean() > 0.02) != culled


@pytest.mark.parametrize("case", range(len(sws.DENSE_CASES)))
def test_oracle_dense_model_render_matches_golden(product, oracle, assets, case):
    entry = json.load(open(GOLDEN))["dense"][case]
    got = sws.frame_hashes_dense(sws.dense_oracle(oracle, product, lib.check, assets, case))
    assert got == {k: entry[k] for k in got}


def test_planned_frames_replayed_with_the_oracle_match_the_reference(product, oracle, ref_scalar, assets):
    script = sws.build_script()
    expected = sws.run_reference(ref_scalar, assets, script, tempfile.mkdtemp(prefix="dfpsr_sprites_"))
    got = sws.run_plan_oracle(product, lib.check, oracle, assets, script)
    assert len(got) == len(expected) == sum(1 for a in script if a[0] == "draw")
    for index, (a, b) in enumerate(zip(got, expected)):
        assert np.array_equal(a["camera"], b["camera"]) and np.array_equal(a["ground"], b["ground"]), index
        for name in BUFFERS:
            assert np.array_equal(a[name].view(np.uint32), b[name].view(np.uint32)), (index, name)
    assert (expected[0]["light"] & 0xFFFFFF != 0).mean() > 0.5  # the lights reach the scene
    assert got[1]["ops"] < got[0]["ops"] // 2                   # the second frame only restores dirty rectangles


def test_planned_frames_replayed_with_the_oracle_match_golden(product, oracle, assets):
    """The same comparison against hashes of the reference's buffers: runs without /root/reference."""
    golden = json.load(open(GOLDEN))["script_frames"]
    got = sws.run_plan_oracle(product, lib.check, oracle, assets, sws.build_script())
    assert len(got) == len(golden)
    for index, (frame, entry) in enumerate(zip(got, golden)):
        assert sws.frame_hashes(frame) == {k: entry[k] for k in BUFFERS}, index
        assert [int(v) for v in frame["camera"]] == entry["camera"] and [int(v) for v in frame["ground"]] == entry["ground"]


def test_world_errors_mirror_the_reference(product):
    """ref: SDK/SpriteEngine/spriteAPI.cpp:902, :920 — out-of-bound type indices and null handles are errors, not crashes."""
    system = abi.OrthoSystem()
    lib.check(product.dfpsr_ortho_system_create(C.byref(system), -0.6, 64))
    world = C.c_void_p()
    lib.check(product.dfpsr_sprite_world_create(C.byref(world), C.byref(system), 64))
    bad = sws.sprite_instance(10 ** 6, 0, (0, 0, 0), 0)
    assert product.dfpsr_sprite_world_add_background_sprite(world, C.byref(bad)) != 0
    assert b"out of bound" in product.dfpsr_last_error()
    assert product.dfpsr_sprite_world_add_temporary_sprite(None, C.byref(bad)) != 0
    assert b"null" in product.dfpsr_last_error()
    assert product.dfpsr_sprite_world_set_camera_direction_index(world, 8) != 0
    image = abi.Image.null()
    assert product.dfpsr_sprite_world_draw(world, C.byref(image), None) != 0  # no GPU here / no target: fails loudly either way
    lib.check(product.dfpsr_sprite_world_destroy(world))
