"""The C++14 host shim (dfpsr_b200/host/dsr_b200.h): DFPSR's own API names over the C ABI. The C++ test program renders a scene
through renderer_begin/giveTask/end and model_render exactly like SDK/terrain/main.cpp does; its pixels must equal the oracle's."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import orcbind
from dfpsr_b200 import abi, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "dfpsr_b200", "host", "shim_test")


def write_scene(path, w, h, points, polygons, texture, levels, filt, cam_params):
    loc = cam_params.location
    header = struct.pack("<9i12ff", w, h, len(points), len(polygons), texture.shape[1] if texture is not None else 0, texture.shape[0] if texture is not None else 0,
                         levels, filt, cam_params.perspective, *loc.position, *loc.xAxis, *loc.yAxis, *loc.zAxis, cam_params.widthSlope)
    with open(path, "wb") as f:
        f.write(header)
        f.write(np.ascontiguousarray(points, np.float32).tobytes())
        f.write(np.ascontiguousarray(polygons).tobytes())
        if texture is not None:
            f.write(np.ascontiguousarray(texture, np.uint32).tobytes())


def test_shim_program_is_built():
    assert os.path.exists(EXE), "run `make -C dfpsr_b200/host` (done by __graft_entry__.build())"


def test_shim_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sc = scenes.random_soup(4, 1)
    cam = abi.camera_params(True, scenes.look_at_transform((0, 0, -3), (0, 0, 0)), 32, 32)
    write_scene(tmp_path / "scene.bin", 32, 32, sc["points"], sc["polygons"], None, 1, 0, cam)
    out = subprocess.run([EXE, str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert out.returncode != 0
    assert "no CUDA device" in out.stderr or "no CUDA device" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["terrain", "soup_alpha", "soup_ortho"])
def test_shim_frame_matches_oracle(cuda, oracle, tmp_path, kind):
    if kind == "terrain":
        sc = scenes.terrain_scene()
        w, h, points, polygons, texture, levels, filt = 640, 360, sc["points"], sc["polygons"], sc["texture"], 5, abi.FILTER_SOLID
        cam = scenes.orbit_camera(12, w, h)
    else:
        alpha = kind == "soup_alpha"
        soup = scenes.random_soup(200, 4, textured=True, alpha=alpha)
        w, h, points, polygons, texture, levels = 322, 201, soup["points"], soup["polygons"], scenes.checker_texture(64, 2), 4
        filt = abi.FILTER_ALPHA if alpha else abi.FILTER_SOLID
        persp = kind != "soup_ortho"
        cam = abi.camera_params(persp, scenes.look_at_transform((0.4, 0.2, -0.6), (0.1, -0.2, 2.0)), w, h, width_slope=1.0 if persp else 6.0)
    write_scene(tmp_path / "scene.bin", w, h, points, polygons, texture, levels, filt, cam)
    out = subprocess.run([EXE, str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    raw = np.fromfile(tmp_path / "out.bin", np.uint32)
    got_c, got_d = raw[: w * h].reshape(h, w), raw[w * h:].reshape(h, w)
    buf, tex = orcbind.build_texture(texture, levels)
    model, keep = orcbind.model_of(points, polygons, filt, tex)
    ec = np.zeros((h, w), np.uint32)
    ed = np.zeros((h, w), np.float32) if cam.perspective else np.full((h, w), 1e9, np.float32)
    ident = abi.Transform3D.identity()
    n = oracle.orc_model_render(C.byref(model), C.byref(ident), C.byref(orcbind.image_of(ec)), C.byref(orcbind.image_of(ed)), C.byref(orcbind.camera(cam)))
    assert n > 20
    assert np.array_equal(got_d, ed.view(np.uint32)), "depth"
    assert np.array_equal(got_c, ec), "colour"


@pytest.mark.gpu
def test_shim_sprite_world_session(tmp_path, oracle):
    """spriteWorld_* through the C++ shim == the host planner replayed with the oracle (same calls through ctypes)."""
    import sprite_world_scene as sws
    from dfpsr_b200 import lib
    assets = sws.build_assets()
    sprite, model = assets["sprites"][1], assets["models"][0]
    w, h = 300, 220
    atlas = np.ascontiguousarray(sprite["atlas"])
    header = struct.pack("<9i6f", atlas.shape[1], atlas.shape[0], sprite["frames"], sprite["center"][0], sprite["center"][1], len(model["points"]), len(model["polygons"]), w, h,
                         *[float(v) for v in sprite["min"]], *[float(v) for v in sprite["max"]])
    with open(tmp_path / "assets.bin", "wb") as f:
        f.write(header + atlas.tobytes() + np.ascontiguousarray(model["points"], np.float32).tobytes() + np.ascontiguousarray(model["polygons"]).tobytes())
    result = subprocess.run([EXE, "--sprites", str(tmp_path / "assets.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert result.returncode == 0, result.stdout + result.stderr
    raw = np.fromfile(tmp_path / "out.bin", np.uint32)
    frames, tail = raw[:2 * w * h].reshape(2, h, w), raw[2 * w * h:].view(np.int32)
    # the same session through the C ABI's host planner + the oracle
    handle = lib.load()
    local = {"sprites": [dict(sprite, points=None, indices=None)], "models": [dict(model, shadow_points=model["points"], shadow_polygons=model["polygons"])]}
    pw = sws.ProductWorld(handle, lib.check, local)  # OrthoSystem(-0.6, 64), shadow resolution 64: what the C++ program uses
    for x in range(-3, 4):
        for z in range(-3, 4):
            pw.apply(("bg_sprite", 0, (x + z + 16) % 8, (x * 1024, 0, z * 1024), 1))
    pw.apply(("bg_model", 0, (0.5, 0.0, -0.5), ((1, 0, 0), (0, 1, 0), (0, 0, 1))))
    ex = sws.OracleExecutor(oracle, pw)
    for frame in range(2):
        pw.apply(("clear_temporary",))
        pw.apply(("directed", (1.0, -1.0, 0.0), 0.1, (255, 255, 255)))
        pw.apply(("point", (0.5, 1.5, 0.5), 4.0, 1.0, (255, 200, 150), 1))
        pw.apply(("tmp_sprite", 0, frame, (300 + 200 * frame, 256, -100), 1))
        if frame == 1:
            pw.apply(("move_camera", 12, -7))
        ops, count = C.POINTER(abi.SpriteWorldOp)(), C.c_int32()
        lib.check(handle.dfpsr_sprite_world_plan_frame(pw.world, w, h, C.byref(ops), C.byref(count)))
        expected = ex.frame(ops, count.value, w, h, 0)
        assert np.array_equal(frames[frame], expected["color"]), frame
        assert (frames[frame] != 0).mean() > 0.5
    location, _ground, _index = pw.camera_state(w, h)
    assert list(tail[3:6]) == [int(v) for v in location]
    pw.close()


@pytest.mark.gpu
def test_shim_filter_lambdas_and_pixel_programs(tmp_path):
    """Code written for the reference's filter API (host lambdas capturing images, api/filterAPI.h:62-79) compiles unchanged against the
    shim; the same functions as device pixel programs (NVRTC) give the same pixels; image_writePixel, both addPointLight signatures."""
    result = subprocess.run([EXE, "--filters", str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert result.returncode == 0, result.stdout + result.stderr
    raw = np.fromfile(tmp_path / "out.bin", np.uint32)
    brighter = raw[:64 * 64].reshape(64, 64)
    y, x = np.mgrid[0:64, 0:64]
    red, green = np.minimum(np.minimum(x * 4, 255) * 2, 255), np.minimum(np.minimum(y * 4, 255) * 2, 255)
    assert np.array_equal(brighter, (red | (green << 8) | (255 << 24)).astype(np.uint32))


@pytest.mark.gpu
def test_filter_map_program_matches_numpy(cuda):
    """dfpsr_filter_map_program through the C ABI: an arbitrary integer function of (x, y) and two sources in different pack orders."""
    import torch
    from dfpsr_b200 import lib
    rng = np.random.default_rng(3)
    h, w = 37, 101
    a = rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)
    b = rng.integers(0, 2 ** 32, (16, 16), dtype=np.uint32)
    ta, tb, out = lib.to_device(a), lib.to_device(b), lib.to_device(np.zeros((h, w), np.uint32))
    sources = (abi.Image * 2)(lib.image(ta, abi.PACK_RGBA), lib.image(tb, abi.PACK_BGRA))
    body = b"int4 s = read_clamp(0, x + 1, y - 1); int4 t = read_tile(1, x, y); return make_int4(s.x + t.x - 40, (s.y * t.y) >> 7, s.z ^ t.z, 255 - s.w + source_width(1));"
    lib.check(cuda.dfpsr_filter_map_program(C.byref(lib.image(out, abi.PACK_ARGB)), body, sources, 2, 5, -3, lib.stream_ptr()))
    got = out.cpu().numpy().view(np.uint32)
    yy, xx = np.mgrid[0:h, 0:w]
    xx, yy = xx + 5, yy - 3
    sx, sy = np.clip(xx + 1, 0, w - 1), np.clip(yy - 1, 0, h - 1)
    s = a[sy, sx]
    t = b[yy % 16, xx % 16]
    sr, sg, sb, sa = [((s >> k) & 255).astype(np.int64) for k in (0, 8, 16, 24)]
    tr, tg, tb_, ta_ = [((t >> k) & 255).astype(np.int64) for k in (16, 8, 0, 24)]  # BGRA: red in byte 2
    r, g, bl, al = np.clip(sr + tr - 40, 0, 255), np.clip((sg * tg) >> 7, 0, 255), np.clip(sb ^ tb_, 0, 255), np.clip(255 - sa + 16, 0, 255)
    expected = ((al << 0) | (r << 8) | (g << 16) | (bl << 24)).astype(np.uint32)  # ARGB: alpha in byte 0
    assert np.array_equal(got, expected)
    # the compiled program is cached: a second call with the same text launches at once; a broken body reports the compiler's message
    lib.check(cuda.dfpsr_filter_map_program(C.byref(lib.image(out, abi.PACK_ARGB)), body, sources, 2, 5, -3, lib.stream_ptr()))
    assert cuda.dfpsr_filter_map_program(C.byref(lib.image(out)), b"return 1;", None, 0, 0, 0, lib.stream_ptr()) != 0
    assert b"does not compile" in cuda.dfpsr_last_error()
