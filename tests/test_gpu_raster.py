"""Parity of the CUDA triangle pipeline (through the C ABI) against the oracle: bit-exact colour and depth."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import orcbind
from dfpsr_b200 import abi, lib, scenes
from gpuutil import CudaScene, assert_same_u32, bits, dev, host_f32, host_u32

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "raster.json")


def test_project_points(cuda, oracle):
    sc = scenes.terrain_scene()
    cam_params = scenes.orbit_camera(5, 1920, 1080)
    cam = lib.camera(cam_params)
    assert bytes(cam) == bytes(orcbind.camera(cam_params))
    m2w = abi.Transform3D.identity()
    pts = dev(sc["points"].reshape(-1))
    import torch
    out = torch.zeros(len(sc["points"]) * 40, dtype=torch.uint8, device="cuda")
    lib.check(cuda.dfpsr_project_points(pts.data_ptr(), len(sc["points"]), C.byref(m2w), C.byref(cam), out.data_ptr(), lib.stream_ptr()))
    got = out.cpu().numpy().view(abi.PROJECTED_DTYPE)
    expected = np.zeros(len(sc["points"]), abi.PROJECTED_DTYPE)
    oracle.orc_project_points(orcbind.ptr(sc["points"]), len(sc["points"]), C.byref(m2w), C.byref(orcbind.camera(cam_params)), orcbind.ptr(expected))
    for field in ("cs", "is", "flat"):
        assert np.array_equal(got[field].view(np.uint8), expected[field].view(np.uint8)), field


@pytest.mark.parametrize("frame", [0, 7, 23])
def test_terrain_matches_oracle(cuda, oracle, frame):
    sc = scenes.terrain_scene()
    scene = CudaScene(sc["points"], sc["polygons"], diffuse_level0=sc["texture"], diffuse_levels=5)
    w, h = 640, 360
    cam = scenes.orbit_camera(frame, w, h)
    c0, d0 = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
    got_c, got_d = scene.render_cuda(cuda, cam, c0, d0)
    exp_c, exp_d, commands = scene.render_oracle(oracle, cam, c0, d0)
    assert commands > 500
    assert_same_u32(bits(got_d), bits(exp_d), "depth")
    assert_same_u32(got_c, exp_c, "colour")


CASES = [
    dict(textured=False, light=False, vcol=True, alpha=False, pack=0),
    dict(textured=True, light=False, vcol=False, alpha=False, pack=0),
    dict(textured=True, light=False, vcol=True, alpha=False, pack=1),
    dict(textured=True, light=True, vcol=True, alpha=False, pack=2),
    dict(textured=False, light=True, vcol=False, alpha=False, pack=3),
    dict(textured=False, light=True, vcol=True, alpha=False, pack=0),
    dict(textured=True, light=True, vcol=False, alpha=False, pack=0),
    dict(textured=True, light=False, vcol=True, alpha=True, pack=0),
    dict(textured=False, light=False, vcol=True, alpha=True, pack=1),
    dict(textured=True, light=False, vcol=True, alpha=True, pack=0, use_depth=False),
    dict(textured=True, light=False, vcol=True, alpha=False, pack=0, use_depth=False),
    dict(textured=True, light=False, vcol=True, alpha=False, pack=0, use_color=False),
    dict(textured=True, light=False, vcol=True, alpha=False, pack=0, persp=False),
    dict(textured=False, light=False, vcol=True, alpha=True, pack=0, persp=False),
    dict(textured=True, light=False, vcol=True, alpha=False, pack=0, far=float("inf")),
    dict(textured=True, light=False, vcol=True, alpha=False, pack=0, far=5.0),
]


def soup_case(seed, textured, light, vcol, alpha, pack, use_color=True, use_depth=True, persp=True, far=1000.0, n=300, w=320, h=200):
    soup = scenes.random_soup(n, seed, textured=textured or light, vertex_colors=vcol, alpha=alpha)
    filt = abi.FILTER_ALPHA if alpha else abi.FILTER_SOLID
    scene = CudaScene(soup["points"], soup["polygons"], filt,
                      scenes.checker_texture(64, seed + 1) if textured else None, 4,
                      scenes.checker_texture(32, seed + 2) if light else None, 1)
    rng = np.random.default_rng(seed)
    pos = (rng.random(3) * 2 - 1) * 2
    target = (rng.random(3) * 2 - 1) * 3
    cam = abi.camera_params(persp, scenes.look_at_transform(pos, target), w, h, width_slope=(1.0 if persp else 6.0), far=far)
    color = rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32) if use_color else None
    depth = (np.zeros((h, w), np.float32) if persp else np.full((h, w), 1e9, np.float32)) if use_depth else None
    return scene, cam, color, depth


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("seed", [0, 1])
def test_random_soup_matches_oracle(cuda, oracle, case, seed):
    """Triangles around and behind the camera: culling, near/side clipping, every shader variant, alpha filter,
    light maps, all pack orders, colour-only / depth-only targets, orthogonal cameras, no / near far plane."""
    cfg = CASES[case]
    scene, cam, color, depth = soup_case(100 * case + seed, **cfg)
    got_c, got_d = scene.render_cuda(cuda, cam, color, depth, cfg["pack"])
    exp_c, exp_d, commands = scene.render_oracle(oracle, cam, color, depth, cfg["pack"])
    assert commands > 10
    if depth is not None:
        assert_same_u32(bits(got_d), bits(exp_d), "depth")
    if color is not None:
        assert_same_u32(got_c, exp_c, "colour")


def test_submission_order_across_models(cuda, oracle):
    """Two models in one renderer_begin/end frame: the second (alpha-filtered) must blend over the first."""
    a = scenes.random_soup(150, 11, textured=False)
    b = scenes.random_soup(150, 12, textured=False, alpha=True)
    sa = CudaScene(a["points"], a["polygons"], abi.FILTER_SOLID)
    sb = CudaScene(b["points"], b["polygons"], abi.FILTER_ALPHA)
    w, h = 256, 160
    cam_params = abi.camera_params(True, scenes.look_at_transform((0.3, 0.1, -0.5), (0, 0, 2)), w, h)
    cam = lib.camera(cam_params)
    c0 = np.full((h, w), 0x80402010, np.uint32)
    d0 = np.zeros((h, w), np.float32)
    tc, td = dev(c0), dev(d0)
    r = C.c_void_p()
    lib.check(cuda.dfpsr_renderer_create(C.byref(r)))
    ident = abi.Transform3D.identity()
    lib.check(cuda.dfpsr_renderer_begin(r, C.byref(lib.image(tc)), C.byref(lib.image(td))))
    assert cuda.dfpsr_renderer_begin(r, C.byref(lib.image(tc)), C.byref(lib.image(td))) != 0  # begin twice is an error
    assert b"twice" in cuda.dfpsr_last_error()
    lib.check(cuda.dfpsr_renderer_give_task(r, C.byref(sa.model.desc), C.byref(ident), C.byref(cam), lib.stream_ptr()))
    lib.check(cuda.dfpsr_renderer_give_task(r, C.byref(sb.model.desc), C.byref(ident), C.byref(cam), lib.stream_ptr()))
    lib.check(cuda.dfpsr_renderer_end(r, lib.stream_ptr()))
    assert cuda.dfpsr_renderer_end(r, lib.stream_ptr()) != 0  # end without begin is an error
    count = C.c_int64()
    lib.check(cuda.dfpsr_renderer_last_command_count(r, C.byref(count), lib.stream_ptr()))
    lib.check(cuda.dfpsr_renderer_destroy(r))
    ec, ed, n1 = sa.render_oracle(oracle, cam_params, c0, d0)
    ec, ed, n2 = sb.render_oracle(oracle, cam_params, ec, ed)
    assert count.value == n1 + n2
    assert_same_u32(bits(host_f32(td)), bits(ed), "depth")
    assert_same_u32(host_u32(tc), ec, "colour")


def test_begin_cleared_equals_fill_then_render(cuda, oracle):
    sc = scenes.terrain_scene()
    scene = CudaScene(sc["points"], sc["polygons"], diffuse_level0=sc["texture"], diffuse_levels=5)
    w, h = 322, 182  # not a multiple of the tile size
    cam_params = scenes.orbit_camera(3, w, h)
    cam = lib.camera(cam_params)
    rng = np.random.default_rng(1)
    tc, td = dev(rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)), dev(rng.random((h, w)).astype(np.float32))
    r = C.c_void_p()
    lib.check(cuda.dfpsr_renderer_create(C.byref(r)))
    ident = abi.Transform3D.identity()
    lib.check(cuda.dfpsr_renderer_begin_cleared(r, C.byref(lib.image(tc)), C.byref(lib.image(td)), 0, 0.0))
    lib.check(cuda.dfpsr_renderer_give_task(r, C.byref(scene.model.desc), C.byref(ident), C.byref(cam), lib.stream_ptr()))
    lib.check(cuda.dfpsr_renderer_end(r, lib.stream_ptr()))
    lib.check(cuda.dfpsr_renderer_destroy(r))
    ec, ed, _ = scene.render_oracle(oracle, cam_params, np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32))
    assert_same_u32(bits(host_f32(td)), bits(ed), "depth")
    assert_same_u32(host_u32(tc), ec, "colour")


@pytest.mark.parametrize("persp", [True, False])
def test_render_depth_matches_oracle(cuda, oracle, persp):
    soup = scenes.random_soup(200, 9, textured=False)
    scene = CudaScene(soup["points"], soup["polygons"])
    cam_params = abi.camera_params(persp, scenes.look_at_transform((0.5, 0.2, -0.3), (1, 0.5, 2)), 256, 256, width_slope=(1.0 if persp else 5.0))
    init = np.zeros((256, 256), np.float32) if persp else np.full((256, 256), 1e9, np.float32)
    td = dev(init)
    ident = abi.Transform3D.identity()
    lib.check(cuda.dfpsr_model_render_depth(C.byref(scene.model.desc), C.byref(ident), C.byref(lib.image(td)), C.byref(lib.camera(cam_params)), lib.stream_ptr()))
    expected = init.copy()
    oracle.orc_model_render_depth(C.byref(scene.o_model), C.byref(ident), C.byref(orcbind.image_of(expected)), C.byref(orcbind.camera(cam_params)))
    assert (expected != init).mean() > 0.2
    assert_same_u32(bits(host_f32(td)), bits(expected), "depth")


def test_pre_projected_triangles(cuda, oracle):
    """renderer_giveTask_triangle: projected points come from the caller (host)."""
    soup = scenes.random_soup(120, 21, textured=True)
    w, h = 200, 120
    cam_params = abi.camera_params(True, scenes.look_at_transform((0.1, 0.4, -0.2), (0.5, 0, 2)), w, h)
    ocam = orcbind.camera(cam_params)
    ident = abi.Transform3D.identity()
    projected = np.zeros(len(soup["points"]), abi.PROJECTED_DTYPE)
    oracle.orc_project_points(orcbind.ptr(soup["points"]), len(soup["points"]), C.byref(ident), C.byref(ocam), orcbind.ptr(projected))
    tris = np.zeros(len(soup["polygons"]), abi.TRIANGLE_DTYPE)
    idx = soup["polygons"]["pointIndices"][:, :3]
    tris["pos"] = projected[idx]
    tris["colors"] = soup["polygons"]["colors"][:, :3]
    tris["texCoords"] = soup["polygons"]["texCoords"][:, :3]
    tex0 = scenes.checker_texture(64, 3)
    dtex = lib.DeviceTexture(tex0, 3)
    obuf, otex = orcbind.build_texture(tex0, 3)
    c0, d0 = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
    tc, td = dev(c0), dev(d0)
    r = C.c_void_p()
    lib.check(cuda.dfpsr_renderer_create(C.byref(r)))
    lib.check(cuda.dfpsr_renderer_begin(r, C.byref(lib.image(tc)), C.byref(lib.image(td))))
    lib.check(cuda.dfpsr_renderer_give_task_triangles(r, tris.ctypes.data, len(tris), C.byref(dtex.desc), None, abi.FILTER_SOLID, C.byref(lib.camera(cam_params)), lib.stream_ptr()))
    lib.check(cuda.dfpsr_renderer_end(r, lib.stream_ptr()))
    lib.check(cuda.dfpsr_renderer_destroy(r))
    ec, ed = c0.copy(), d0.copy()
    oracle.orc_render_triangles(tris.ctypes.data, len(tris), C.byref(otex), None, abi.FILTER_SOLID, C.byref(orcbind.image_of(ec)), C.byref(orcbind.image_of(ed)), C.byref(ocam))
    assert (ed > 0).mean() > 0.2
    assert_same_u32(bits(host_f32(td)), bits(ed), "depth")
    assert_same_u32(host_u32(tc), ec, "colour")


def test_many_triangles_per_tile(cuda, oracle):
    """More entries in one tile than the shared-memory sort window holds would need >16384; exercise a long list (several chunks)."""
    soup = scenes.random_soup(3000, 31, extent=1.5, tri_size=1.0, textured=False)
    scene = CudaScene(soup["points"], soup["polygons"])
    w, h = 96, 64
    cam = abi.camera_params(True, scenes.look_at_transform((0, 0, -4), (0, 0, 0)), w, h)
    c0, d0 = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
    got_c, got_d = scene.render_cuda(cuda, cam, c0, d0)
    exp_c, exp_d, commands = scene.render_oracle(oracle, cam, c0, d0)
    assert commands > 1000
    assert_same_u32(bits(got_d), bits(exp_d), "depth")
    assert_same_u32(got_c, exp_c, "colour")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_terrain_1080p_golden(cuda):
    """BASELINE config 1 at full size against hashes produced by the compiled reference (tests/golden/make_golden.py)."""
    golden = json.load(open(GOLDEN))["terrain_1080p"]
    sc = scenes.terrain_scene()
    scene = CudaScene(sc["points"], sc["polygons"], diffuse_level0=sc["texture"], diffuse_levels=5)
    for entry in golden:
        cam = scenes.orbit_camera(entry["frame"], 1920, 1080)
        c, d = scene.render_cuda(cuda, cam, np.zeros((1080, 1920), np.uint32), np.zeros((1080, 1920), np.float32))
        assert sha(d) == entry["depth_sha256"], f"frame {entry['frame']} depth"
        assert sha(c) == entry["color_sha256"], f"frame {entry['frame']} colour"


def test_sdk_terrain_real_media_golden(cuda):
    """BASELINE config 1 on the SDK's REAL assets: SDK/terrain/media through the SDK's own generators (tests/golden/make_sdk_golden.py), rendered at
    1920x1080 in exact mode: sha256 of colour and depth equal the unmodified reference's."""
    golden = json.load(open(os.path.join(os.path.dirname(GOLDEN), "sdk_terrain.json")))
    data = np.load(os.path.join(os.path.dirname(GOLDEN), "sdk_terrain_scene.npz"))
    polygons = data["polygons"].view(abi.POLYGON_DTYPE).reshape(-1)
    assert len(data["points"]) == golden["points"] == 4096 and len(polygons) == golden["polygons"]
    scene = CudaScene(data["points"], polygons, diffuse_level0=data["texture"], diffuse_levels=int(data["levels"]))
    for entry in golden["frames"]:
        cam = scenes.orbit_camera(entry["frame"], 1920, 1080)
        c, d = scene.render_cuda(cuda, cam, np.zeros((1080, 1920), np.uint32), np.zeros((1080, 1920), np.float32))
        assert sha(d) == entry["depth_sha256"], f"frame {entry['frame']} depth"
        assert sha(c) == entry["color_sha256"], f"frame {entry['frame']} colour"


def test_tiny_triangles_4k_golden(cuda):
    """BASELINE config 3 (2 M tiny vertex-coloured triangles at 3840x2160) against the reference's hashes."""
    golden = json.load(open(GOLDEN))["tiny_4k"]
    sc = scenes.tiny_triangle_scene(golden["nx"], golden["nz"])
    scene = CudaScene(sc["points"], sc["polygons"])
    cam = scenes.top_down_camera(golden["nx"], golden["nz"], 3840, 2160)
    c, d = scene.render_cuda(cuda, cam, np.zeros((2160, 3840), np.uint32), np.zeros((2160, 3840), np.float32))
    assert sha(d) == golden["depth_sha256"]
    assert sha(c) == golden["color_sha256"]


@pytest.mark.parametrize("size", [(33, 35), (321, 181), (320, 181), (77, 3), (5, 1), (1, 1), (130, 66)])
@pytest.mark.parametrize("case", [0, 1, 7, 8, 12])
def test_odd_target_sizes(cuda, oracle, case, size):
    """Odd widths (rows that are only 4-byte aligned) and odd heights (the reference's repeated-upper-row quirk in the last row pair,
    ref: shader/fillerTemplates.h:286-331), targets smaller than one tile."""
    cfg = CASES[case]
    scene, cam, color, depth = soup_case(500 + case, w=size[0], h=size[1], **cfg)
    got_c, got_d = scene.render_cuda(cuda, cam, color, depth, cfg["pack"])
    exp_c, exp_d, _ = scene.render_oracle(oracle, cam, color, depth, cfg["pack"])
    if depth is not None:
        assert_same_u32(bits(got_d), bits(exp_d), "depth")
    if color is not None:
        assert_same_u32(got_c, exp_c, "colour")


@pytest.mark.parametrize("world", [2, 3, 8])
def test_strip_mode_equals_full_frame(cuda, oracle, world):
    """Screen strips: every 'rank' draws only its rows; the union equals the full frame bit for bit."""
    from dfpsr_b200 import shard
    sc = scenes.terrain_scene()
    scene = CudaScene(sc["points"], sc["polygons"], diffuse_level0=sc["texture"], diffuse_levels=5)
    w, h = 480, 270
    cam_params = scenes.orbit_camera(9, w, h)
    cam = lib.camera(cam_params)
    exp_c, exp_d, _ = scene.render_oracle(oracle, cam_params, np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32))
    tc, td = dev(np.full((h, w), 0xDEADBEEF, np.uint32)), dev(np.full((h, w), -1.0, np.float32))
    ident = abi.Transform3D.identity()
    r = C.c_void_p()
    lib.check(cuda.dfpsr_renderer_create(C.byref(r)))
    bounds = shard.strip_rows(h, world, align=4)
    for y0, y1 in bounds:
        lib.check(cuda.dfpsr_renderer_begin_cleared(r, C.byref(lib.image(tc)), C.byref(lib.image(td)), 0, 0.0))
        lib.check(cuda.dfpsr_renderer_set_clip_rows(r, y0, y1))
        lib.check(cuda.dfpsr_renderer_give_task(r, C.byref(scene.model.desc), C.byref(ident), C.byref(cam), lib.stream_ptr()))
        lib.check(cuda.dfpsr_renderer_end(r, lib.stream_ptr()))
        got_c, got_d = host_u32(tc), host_f32(td)
        assert_same_u32(got_c[y0:y1], exp_c[y0:y1], f"colour rows {y0}:{y1}")
        assert_same_u32(bits(got_d[y0:y1]), bits(exp_d[y0:y1]), f"depth rows {y0}:{y1}")
        if y1 < h:
            assert np.all(got_c[y1:] == 0xDEADBEEF), "rows below the strip were touched"
    assert cuda.dfpsr_renderer_begin(r, C.byref(lib.image(tc)), C.byref(lib.image(td))) == 0
    assert cuda.dfpsr_renderer_set_clip_rows(r, 2, 10) != 0  # not a multiple of the tile height
    lib.check(cuda.dfpsr_renderer_end(r, lib.stream_ptr()))
    lib.check(cuda.dfpsr_renderer_destroy(r))
    assert_same_u32(host_u32(tc), exp_c, "colour")
    assert_same_u32(bits(host_f32(td)), bits(exp_d), "depth")


def test_render_views_batch_equals_single_renders(cuda, oracle):
    """dfpsr_model_render_views: one submission for many views == one dfpsr_model_render per view (BASELINE config 4)."""
    import torch
    sc = scenes.terrain_scene()
    scene = CudaScene(sc["points"], sc["polygons"], diffuse_level0=sc["texture"], diffuse_levels=5)
    w, h, views = 200, 114, 7
    cams = (abi.Camera * views)(*[lib.camera(scenes.orbit_camera(3 * v, w, h)) for v in range(views)])
    color = torch.full((views, h, w), 7, dtype=torch.int32, device="cuda")
    depth = torch.full((views, h, w), 3.0, dtype=torch.float32, device="cuda")
    colors = (abi.Image * views)(*[lib.image(color[v]) for v in range(views)])
    depths = (abi.Image * views)(*[lib.image(depth[v]) for v in range(views)])
    ident = abi.Transform3D.identity()
    lib.check(cuda.dfpsr_model_render_views(C.byref(scene.model.desc), C.byref(ident), colors, depths, cams, views, 1, lib.stream_ptr()))
    for v in range(views):
        ec, ed, _ = scene.render_oracle(oracle, scenes.orbit_camera(3 * v, w, h), np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32))
        assert_same_u32(host_u32(color[v]), ec, f"view {v} colour")
        assert_same_u32(bits(host_f32(depth[v])), bits(ed), f"view {v} depth")


def test_render_depth_batch_matches_oracle(cuda, oracle):
    """dfpsr_model_render_depth_batch: several casters into several cube-face targets in one submission, fused clear."""
    import torch
    res, faces = 64, 6
    soups = [scenes.random_soup(60, 40 + i, extent=2.0, tri_size=1.0, textured=False) for i in range(3)]
    cscenes = [CudaScene(s["points"], s["polygons"]) for s in soups]
    sides = [((1, 0, 0), (0, 1, 0)), ((-1, 0, 0), (0, 1, 0)), ((0, 1, 0), (0, 0, 1)), ((0, -1, 0), (0, 0, 1)), ((0, 0, 1), (0, 1, 0)), ((0, 0, -1), (0, 1, 0))]
    cam_params = [abi.camera_params(True, abi.Transform3D.make((0, 0, 0), np.stack(scenes.make_axis_system(f, u))), res, res) for f, u in sides]
    cube = torch.full((faces * res, res), 9.0, dtype=torch.float32, device="cuda")
    targets = (abi.Image * faces)(*[lib.image(cube[f * res:(f + 1) * res]) for f in range(faces)])
    n = len(cscenes) * faces
    models = (C.c_void_p * n)()
    transforms = (abi.Transform3D * n)()
    cams = (abi.Camera * n)()
    target_of = (C.c_int32 * n)()
    moves = [abi.Transform3D.make((0.3 * i, -0.2 * i, 0.1), ((1, 0, 0), (0, 1, 0), (0, 0, 1))) for i in range(len(cscenes))]
    k = 0
    for i, cs in enumerate(cscenes):
        for f in range(faces):
            models[k] = C.addressof(cs.model.desc)
            transforms[k] = moves[i]
            cams[k] = lib.camera(cam_params[f])
            target_of[k] = f
            k += 1
    lib.check(cuda.dfpsr_model_render_depth_batch(models, transforms, cams, target_of, n, targets, faces, 1, 0.0, lib.stream_ptr()))
    expected = np.zeros((faces * res, res), np.float32)
    for i, cs in enumerate(cscenes):
        for f in range(faces):
            oracle.orc_model_render_depth(C.byref(cs.o_model), C.byref(moves[i]), C.byref(orcbind.image_of(expected[f * res:(f + 1) * res])), C.byref(orcbind.camera(cam_params[f])))
    assert (expected != 0).mean() > 0.1
    assert_same_u32(bits(host_f32(cube)), bits(expected), "cube map")


def test_thousands_of_triangles_on_one_tile(cuda, oracle):
    """More than 4096 entries in one 32x4 tile: the rank-sort path of sort_lists_kernel."""
    soup = scenes.random_soup(6000, 77, extent=0.4, tri_size=0.8, textured=False)
    scene = CudaScene(soup["points"], soup["polygons"])
    w, h = 40, 12
    cam = abi.camera_params(True, scenes.look_at_transform((0, 0, -3), (0, 0, 0)), w, h, width_slope=0.2)
    c0, d0 = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
    got_c, got_d = scene.render_cuda(cuda, cam, c0, d0)
    exp_c, exp_d, commands = scene.render_oracle(oracle, cam, c0, d0)
    assert commands > 4500
    assert_same_u32(bits(got_d), bits(exp_d), "depth")
    assert_same_u32(got_c, exp_c, "colour")


def test_session_render_views_host_pipeline(cuda, oracle):
    """dfpsr_session_render_views_host: host geometry in, host images out, chunks of 16 views double-buffered (40 views = 3 chunks)."""
    import torch
    sc = scenes.terrain_scene()
    texture = lib.DeviceTexture(sc["texture"], 5)
    w, h, views = 160, 90, 40
    pts = np.ascontiguousarray(sc["points"], np.float32)
    poly = np.ascontiguousarray(sc["polygons"])
    tex_host = texture.pixels.cpu()
    hm = abi.HostModel()
    hm.points, hm.pointCount = pts.ctypes.data, len(pts)
    hm.polygons, hm.polygonCount = poly.ctypes.data, len(poly)
    hm.filter = abi.FILTER_SOLID
    hm.diffusePixels, hm.diffuseLayout = tex_host.data_ptr(), texture.desc
    mn, mx = lib.model_bounds(pts)
    hm.minBound[:], hm.maxBound[:] = mn, mx
    session, slot = C.c_void_p(), C.c_int32()
    lib.check(cuda.dfpsr_session_create(C.byref(session)))
    lib.check(cuda.dfpsr_session_upload_model(session, C.byref(hm), C.byref(slot)))
    cams = (abi.Camera * views)(*[lib.camera(scenes.orbit_camera(v, w, h, frames_per_lap=views)) for v in range(views)])
    color = torch.zeros((views, h, w), dtype=torch.int32).pin_memory()
    depth = torch.zeros((views, h, w), dtype=torch.float32).pin_memory()
    cptr = (C.c_void_p * views)(*[color[v].data_ptr() for v in range(views)])
    dptr = (C.c_void_p * views)(*[depth[v].data_ptr() for v in range(views)])
    ident = abi.Transform3D.identity()
    for _ in range(2):  # second call reuses the buffers and events
        lib.check(cuda.dfpsr_session_render_views_host(session, slot.value, C.byref(ident), cams, views, cptr, w * 4, dptr, w * 4, w, h, abi.PACK_RGBA, 1, lib.stream_ptr()))
    lib.check(cuda.dfpsr_session_destroy(session))
    buf, otex = orcbind.build_texture(sc["texture"], 5)
    omodel, keep = orcbind.model_of(pts, poly, diffuse=otex)
    for v in (0, 15, 16, 31, 39):
        ec, ed = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
        oracle.orc_model_render(C.byref(omodel), C.byref(ident), C.byref(orcbind.image_of(ec)), C.byref(orcbind.image_of(ed)), C.byref(orcbind.camera(scenes.orbit_camera(v, w, h, frames_per_lap=views))))
        assert_same_u32(color[v].numpy().view(np.uint32), ec, f"view {v} colour")
        assert_same_u32(bits(depth[v].numpy()), bits(ed), f"view {v} depth")


OCCLUSION_VARIANTS = [dict(), dict(top_rows=True), dict(perspective=False), dict(seed=11, width=333, height=201, top_rows=True)]


@pytest.mark.parametrize("variant", OCCLUSION_VARIANTS)
def test_occlusion_grid_matches_oracle(cuda, oracle, variant):
    """renderer_occludeFromBox / occludeFromExistingTriangles / occludeFromTopRows / isBoxVisible (ref: api/rendererAPI.cpp:181-477):
    same visibility answers as the oracle, same command count, same pixels."""
    import occlusion_scene
    sc = occlusion_scene.build(**variant)
    expected = occlusion_scene.run_oracle(oracle, sc)
    got = occlusion_scene.run_cuda(cuda, sc)
    assert got["visible"] == expected["visible"]
    assert got["commands"] == expected["commands"]
    assert_same_u32(bits(got["depth"]), bits(expected["depth"]), "depth")
    assert_same_u32(got["color"], expected["color"], "colour")


@pytest.mark.parametrize("variant", OCCLUSION_VARIANTS)
def test_debug_wireframe_matches_oracle(cuda, oracle, variant):
    """renderer_end(renderer, debugWireframe = true) (api/rendererAPI.cpp:362-399): white edges of every command the occlusion grid left."""
    import occlusion_scene
    sc = occlusion_scene.build(**variant)
    expected = occlusion_scene.run_oracle(oracle, sc, wireframe=True)
    plain = occlusion_scene.run_oracle(oracle, sc)
    got = occlusion_scene.run_cuda(cuda, sc, wireframe=True)
    assert int((expected["color"] != plain["color"]).sum()) > 200
    assert_same_u32(got["color"], expected["color"], "colour with the overlay")
    assert_same_u32(bits(got["depth"]), bits(expected["depth"]), "depth")
    # the flag covers one frame only
    again = occlusion_scene.run_cuda(cuda, sc)
    assert_same_u32(again["color"], plain["color"], "colour of the next frame")


def test_debug_wireframe_in_row_strips(cuda, oracle):
    """The overlay obeys dfpsr_renderer_set_clip_rows like the frame itself: two strips drawn into one target equal the full frame."""
    import occlusion_scene
    from dfpsr_b200 import lib as L
    sc = occlusion_scene.build()
    expected = occlusion_scene.run_oracle(oracle, sc, wireframe=True)
    cam = L.camera(sc["camera"])
    tc, td = L.to_device(sc["color0"]), L.to_device(sc["depth0"])
    wall = L.DeviceModel(*sc["wall"])
    models = [L.DeviceModel(m["points"], m["polygons"]) for m in sc["models"]]
    height = sc["color0"].shape[0]
    split = (height // 2) & ~3
    r = C.c_void_p()
    s = L.stream_ptr()
    L.check(cuda.dfpsr_renderer_create(C.byref(r)))
    for top, bottom in ((0, split), (split, height)):
        L.check(cuda.dfpsr_renderer_begin(r, C.byref(L.image(tc)), C.byref(L.image(td))))
        L.check(cuda.dfpsr_renderer_set_clip_rows(r, top, bottom))
        L.check(cuda.dfpsr_renderer_give_task(r, C.byref(wall.desc), C.byref(sc["wall_transform"]), C.byref(cam), s))
        for m, dm in zip(sc["models"], models):
            L.check(cuda.dfpsr_renderer_give_task(r, C.byref(dm.desc), C.byref(m["transform"]), C.byref(cam), s))
        L.check(cuda.dfpsr_renderer_set_debug_wireframe(r, 1))
        L.check(cuda.dfpsr_renderer_end(r, s))
    L.check(cuda.dfpsr_renderer_destroy(r))
    # no occluders here (every model is drawn): compare with an oracle frame without occlusion
    plain = occlusion_scene.run_oracle_without_occluders(oracle, sc, wireframe=True)
    assert_same_u32(tc.cpu().numpy().view(np.uint32), plain["color"], "colour of two strips with the overlay")
    assert expected["color"].shape == plain["color"].shape


@pytest.mark.parametrize("variant", OCCLUSION_VARIANTS)
def test_device_broad_phase_equals_host_tests(cuda, oracle, variant):
    """dfpsr_renderer_give_tasks: isBoxSeen and renderer_isBoxVisible per model on the device (against the occluders given before the call)
    draw the same frame, with the same number of commands, as one renderer_giveTask per model with the tests on the host."""
    import occlusion_scene
    sc = occlusion_scene.build(**variant)
    expected = occlusion_scene.run_oracle(oracle, sc)
    got = occlusion_scene.run_cuda_batched(cuda, sc)
    assert got["commands"] == expected["commands"]
    assert_same_u32(bits(got["depth"]), bits(expected["depth"]), "depth")
    assert_same_u32(got["color"], expected["color"], "colour")


def edge_scenes():
    """Degenerate inputs: (name, points, polygons). Shared with the CPU test that pins the oracle to the reference on the same cases."""
    F = np.float32
    poly = lambda rows: np.array([(r, np.zeros((4, 4), F), np.ones((4, 4), F)) for r in rows], abi.POLYGON_DTYPE)
    rng = np.random.default_rng(77)
    out = []
    out.append(("no polygons", np.zeros((3, 3), F), poly([])))
    out.append(("everything behind the camera", np.array([[0, 0, -5], [1, 0, -5], [0, 1, -6]], F), poly([(0, 1, 2, -1)])))
    out.append(("zero-area triangles", np.array([[0, 0, 3], [1, 1, 3], [2, 2, 3], [0.5, 0.5, 3]], F), poly([(0, 1, 2, -1), (0, 0, 0, -1), (3, 3, 1, -1)])))
    out.append(("one triangle covering the whole target and far beyond", np.array([[-900, -900, 2], [900, -900, 2], [0, 1500, 2]], F), poly([(0, 2, 1, -1), (0, 1, 2, -1)])))
    out.append(("huge and non-finite coordinates", np.array([[1e30, 0, 5], [0, 1e30, 5], [-1e30, -1e30, 5], [0, 0, 1e-30], [np.inf, 1, 2], [np.nan, 0, 3], [0, 0, 2], [1, 0, 2], [0, 1, 2]], F),
                poly([(0, 1, 2, -1), (3, 6, 7, -1), (4, 7, 8, -1), (5, 6, 8, -1), (6, 8, 7, -1)])))
    pts = (rng.random((40, 3)) * 2 - 1).astype(F) * np.array([0.02, 0.02, 0.0], F) + np.array([0, 0, 1.5], F)
    out.append(("sub-pixel triangles", pts, poly([(i, i + 1, i + 2, -1) for i in range(0, 36, 3)] + [(i + 2, i + 1, i, -1) for i in range(0, 36, 3)])))
    quad_pts = np.array([[-1, -1, 2], [1, -1, 2], [1, 1, 2.5], [-1, 1, 2.5], [-1, -1, 2], [1, -1, 2], [1, 1, 2.5], [-1, 1, 2.5]], F)
    out.append(("coincident quads: equal depth everywhere, the first one drawn wins", quad_pts, poly([(0, 3, 2, 1), (4, 7, 6, 5)])))
    out[-1][2]["colors"][1, :, :3] = 0.25
    return out


@pytest.mark.parametrize("size", [(1, 1), (2, 2), (3, 5), (31, 4), (33, 9), (640, 3)])
def test_degenerate_inputs_and_tiny_targets(cuda, oracle, size):
    """Empty models, culled and degenerate triangles, non-finite coordinates, coincident surfaces, on targets smaller than one tile."""
    w, h = size
    cam = abi.camera_params(True, scenes.look_at_transform((0, 0, 0), (0, 0, 1)), w, h)
    for name, points, polygons in edge_scenes():
        scene = CudaScene(points, polygons)
        c0, d0 = np.full((h, w), 0x11223344, np.uint32), np.zeros((h, w), np.float32)
        got_c, got_d = scene.render_cuda(cuda, cam, c0, d0)
        exp_c, exp_d, _ = scene.render_oracle(oracle, cam, c0, d0)
        assert_same_u32(bits(got_d), bits(exp_d), f"depth ({name}, {w}x{h})")
        assert_same_u32(got_c, exp_c, f"colour ({name}, {w}x{h})")


def test_imported_models_render_bit_exact(cuda, oracle):
    """The formats feeding the path end to end: a PLY text and a DMF1 text go through dfpsr_import_* (host), the resulting arrays are
    uploaded and rendered by the CUDA pipeline, and the frame equals the oracle's render of the same arrays (vertex colours from the PLY,
    per-vertex colours with an alpha-filtered pass from the DMF1 model)."""
    from importer_cases import CASES
    texts = {name: (kind, text, options) for name, kind, text, options in CASES}
    for name, filter_ in (("ply_basic", abi.FILTER_SOLID), ("ply_basic_flipped", abi.FILTER_SOLID), ("dmf_detail1", abi.FILTER_ALPHA)):
        kind, text, options = texts[name]
        points, polygons, parts, _, _ = lib.import_model(kind, text, **options)
        assert len(polygons) > 0
        scene = CudaScene(points, polygons, filter_)
        w, h = 160, 120
        for eye in ((1.0, 1.5, -6.0), (3.0, -2.0, 5.0)):
            cam = abi.camera_params(True, scenes.look_at_transform(eye, (1.0, 0.8, 0.5)), w, h)
            c0 = np.full((h, w), 0xFF203040, np.uint32)
            d0 = np.zeros((h, w), np.float32)
            got_c, got_d = scene.render_cuda(cuda, cam, c0, d0)
            exp_c, exp_d, commands = scene.render_oracle(oracle, cam, c0, d0)
            assert_same_u32(bits(got_d), bits(exp_d), f"{name} depth")
            assert_same_u32(got_c, exp_c, f"{name} colour")


def test_large_batches_take_the_in_thread_small_triangle_path(cuda, oracle):
    """Frames with up to 148 x 1024 slots send every triangle taller than one tile row to the unit queue; larger batches keep triangles of up
    to 16 rows with their set-up thread (less queue traffic per triangle). Both must give the oracle's pixels: batches of 24 terrain views
    (182 k slots, mip-mapped texture) and of 200 views of an alpha-filtered, partly clipped triangle soup (160 k slots)."""
    import torch
    ident = abi.Transform3D.identity()

    def batch(scene, cam_params, w, h, check):
        views = len(cam_params)
        cams = (abi.Camera * views)(*[lib.camera(p) for p in cam_params])
        color = torch.zeros((views, h, w), dtype=torch.int32, device="cuda")
        depth = torch.zeros((views, h, w), dtype=torch.float32, device="cuda")
        colors = (abi.Image * views)(*[lib.image(color[v]) for v in range(views)])
        depths = (abi.Image * views)(*[lib.image(depth[v]) for v in range(views)])
        lib.check(cuda.dfpsr_model_render_views(C.byref(scene.model.desc), C.byref(ident), colors, depths, cams, views, 1, lib.stream_ptr()))
        for v in check:
            ec, ed, _ = scene.render_oracle(oracle, cam_params[v], np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32))
            assert_same_u32(bits(host_f32(depth[v])), bits(ed), f"view {v} depth")
            assert_same_u32(host_u32(color[v]), ec, f"view {v} colour")

    sc = scenes.terrain_scene()
    terrain = CudaScene(sc["points"], sc["polygons"], diffuse_level0=sc["texture"], diffuse_levels=5)
    assert 24 * 2 * len(sc["polygons"]) > 148 * 1024
    batch(terrain, [scenes.orbit_camera(5 * v, 320, 182) for v in range(24)], 320, 182, check=(0, 7, 23))

    soup = scenes.random_soup(400, 77, extent=2.5, tri_size=0.6, textured=False)
    assert 200 * 2 * len(soup["polygons"]) > 148 * 1024
    alpha = CudaScene(soup["points"], soup["polygons"], abi.FILTER_ALPHA)
    cams = [abi.camera_params(True, scenes.look_at_transform((3.0 * np.cos(0.1 * v), 0.5 + 0.01 * v, 3.0 * np.sin(0.1 * v)), (0, 0, 0)), 96, 64, near=0.5) for v in range(200)]
    batch(alpha, cams, 96, 64, check=(0, 63, 199))


def test_terrain_1080p_view_batch_golden(cuda):
    """BASELINE config 4 at full size: 48 orbit views of the terrain at 1920x1080 in ONE submission (the bench's path: more than 148 x 2048
    units, so one thread per unit, and more than 148 x 1024 slots, so small triangles stay with their set-up thread). The views the
    compiled reference was hashed on (tests/golden/make_golden.py) must come out with exactly those hashes."""
    import torch
    golden = {e["frame"]: e for e in json.load(open(GOLDEN))["terrain_1080p"]}
    sc = scenes.terrain_scene()
    scene = CudaScene(sc["points"], sc["polygons"], diffuse_level0=sc["texture"], diffuse_levels=5)
    views, w, h = 48, 1920, 1080
    cams = (abi.Camera * views)(*[lib.camera(scenes.orbit_camera(v, w, h)) for v in range(views)])
    color = torch.empty((views, h, w), dtype=torch.int32, device="cuda")
    depth = torch.empty((views, h, w), dtype=torch.float32, device="cuda")
    colors = (abi.Image * views)(*[lib.image(color[v]) for v in range(views)])
    depths = (abi.Image * views)(*[lib.image(depth[v]) for v in range(views)])
    ident = abi.Transform3D.identity()
    lib.check(cuda.dfpsr_model_render_views(C.byref(scene.model.desc), C.byref(ident), colors, depths, cams, views, 1, lib.stream_ptr()))
    checked = 0
    for frame, entry in golden.items():
        if frame < views:
            assert sha(host_f32(depth[frame])) == entry["depth_sha256"], f"frame {frame} depth"
            assert sha(host_u32(color[frame])) == entry["color_sha256"], f"frame {frame} colour"
            checked += 1
    assert checked >= 3


def test_parity_does_not_depend_on_what_the_pools_held(cuda):
    """DFPSR_POISON=1 fills the row-interval and checkpoint pools with 0xFF before every emit pass: a record the tile kernel needs and
    nothing wrote then reads as 'nothing here' instead of whatever an earlier frame left at the same address. The environment variable is
    read once per process, so the parity tests that reuse one renderer for many different frames run again in a child process."""
    import subprocess
    import sys
    env = dict(os.environ, DFPSR_POISON="1")
    result = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                             "-k", "terrain_matches_oracle or random_soup_matches_oracle or odd_target_sizes or occlusion_grid_matches_oracle"],
                            env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert result.returncode == 0, result.stdout[-3000:] + result.stderr[-2000:]
    assert " passed" in result.stdout
