"""Parity of the per-pixel CUDA passes (2D draw calls, Sandbox light, filters) against the oracle: bit-exact."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import orcbind
import sandbox_scene
from dfpsr_b200 import abi, lib
from gpuutil import assert_same_u32, bits, dev, host_f32, host_u32

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")
IM, OI = lib.image, orcbind.image_of


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rand_rgba(rng, h, w):
    return rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)


@pytest.mark.parametrize("pack", [0, 1, 2, 3])
def test_image_fill(cuda, oracle, pack):
    t = dev(np.ones((33, 47), np.uint32))
    lib.check(cuda.dfpsr_image_fill_rgba(C.byref(IM(t, pack)), 300, 20, -5, 128, lib.stream_ptr()))
    e = np.zeros((33, 47), np.uint32)
    oracle.orc_image_fill_rgba(C.byref(OI(e, pack)), 300, 20, -5, 128)
    assert_same_u32(host_u32(t), e, "fill rgba")
    f = dev(np.ones((33, 47), np.float32))
    lib.check(cuda.dfpsr_image_fill_f32(C.byref(IM(f)), -2.5, lib.stream_ptr()))
    assert np.all(host_f32(f) == np.float32(-2.5))


@pytest.mark.parametrize("left,top", [(10, 20), (-20, -10), (100, 70), (200, 10), (0, 0)])
def test_draw_copy_and_higher(cuda, oracle, left, top):
    rng = np.random.default_rng(5)
    Ht, Hs = (rng.random((90, 130)) * 5).astype(np.float32), (rng.random((40, 50)) * 6).astype(np.float32)
    Hs[rng.random((40, 50)) < 0.3] = -np.inf
    At, As, Bt, Bs = rand_rgba(rng, 90, 130), rand_rgba(rng, 40, 50), rand_rgba(rng, 90, 130), rand_rgba(rng, 40, 50)
    th, sh, ta, sa, tb, sbb = dev(Ht), dev(Hs), dev(At), dev(As), dev(Bt), dev(Bs)
    lib.check(cuda.dfpsr_draw_higher(C.byref(IM(th)), C.byref(IM(sh)), C.byref(IM(ta, 1)), C.byref(IM(sa)), C.byref(IM(tb)), C.byref(IM(sbb, 2)), left, top, 0.25, lib.stream_ptr()))
    eh, ea, eb = Ht.copy(), At.copy(), Bt.copy()
    oracle.orc_draw_higher(C.byref(OI(eh)), C.byref(OI(Hs)), C.byref(OI(ea, 1)), C.byref(OI(As)), C.byref(OI(eb)), C.byref(OI(Bs, 2)), left, top, 0.25)
    assert_same_u32(bits(host_f32(th)), bits(eh), "higher height")
    assert_same_u32(host_u32(ta), ea, "higher A")
    assert_same_u32(host_u32(tb), eb, "higher B")
    # height only
    th2 = dev(Ht)
    lib.check(cuda.dfpsr_draw_higher(C.byref(IM(th2)), C.byref(IM(sh)), None, None, None, None, left, top, -0.5, lib.stream_ptr()))
    eh2 = Ht.copy()
    oracle.orc_draw_higher(C.byref(OI(eh2)), C.byref(OI(Hs)), None, None, None, None, left, top, -0.5)
    assert_same_u32(bits(host_f32(th2)), bits(eh2), "higher height only")
    # copies
    tc = dev(At)
    lib.check(cuda.dfpsr_draw_copy_rgba(C.byref(IM(tc, 3)), C.byref(IM(sa, 1)), left, top, lib.stream_ptr()))
    ec = At.copy()
    oracle.orc_draw_copy_rgba(C.byref(OI(ec, 3)), C.byref(OI(As, 1)), left, top)
    assert_same_u32(host_u32(tc), ec, "copy rgba")
    tf = dev(Ht)
    lib.check(cuda.dfpsr_draw_copy_f32(C.byref(IM(tf)), C.byref(IM(sh)), left, top, lib.stream_ptr()))
    ef = Ht.copy()
    oracle.orc_draw_copy_f32(C.byref(OI(ef)), C.byref(OI(Hs)), left, top)
    assert_same_u32(bits(host_f32(tf)), bits(ef), "copy f32")


def test_light_passes(cuda, oracle):
    rng = np.random.default_rng(5)
    view = sandbox_scene.ortho_view()
    w, h = 200, 120
    normal, diffuse, light0 = rand_rgba(rng, h, w), rand_rgba(rng, h, w), rand_rgba(rng, h, w)
    height = (rng.random((h, w)) * 3 - 1).astype(np.float32)
    cube = (rng.random((6 * 64, 64)) * 0.5).astype(np.float32)
    d, col = np.array([1, -1, 0.3], np.float32), np.array([255, 200, 90], np.int32)
    for add in (0, 1):
        tl, tn = dev(light0), dev(normal)
        lib.check(cuda.dfpsr_light_directed(C.byref(view), C.byref(IM(tl)), C.byref(IM(tn)), d.ctypes.data, 0.8, col.ctypes.data, add, lib.stream_ptr()))
        e = light0.copy()
        oracle.orc_light_directed(C.byref(view), C.byref(OI(e)), C.byref(OI(normal)), orcbind.ptr(d), 0.8, orcbind.ptr(col), add)
        assert_same_u32(host_u32(tl), e, f"directed add={add}")
    t_diffuse, t_light0, t_normal, t_height = dev(diffuse), dev(light0), dev(normal), dev(height)
    for pack in range(4):
        tc = dev(np.zeros((h, w), np.uint32))
        lib.check(cuda.dfpsr_light_blend(C.byref(IM(tc, pack)), C.byref(IM(t_diffuse)), C.byref(IM(t_light0)), lib.stream_ptr()))
        e = np.zeros((h, w), np.uint32)
        oracle.orc_light_blend(C.byref(OI(e, pack)), C.byref(OI(diffuse)), C.byref(OI(light0)))
        assert_same_u32(host_u32(tc), e, f"blend pack={pack}")
    wc = np.array([w // 2, h // 2], np.int32)
    for pos, rad, use_cube in [((0.2, 0.8, -0.1), 0.9, False), ((0.5, 1.5, 0.3), 2.5, False), ((-0.3, 0.6, 0.2), 0.7, True), ((0.1, 2.0, 0.0), 3.0, True), ((30.0, 1.0, 0.0), 1.0, False)]:
        p, col = np.array(pos, np.float32), np.array([255, 180, 120], np.int32)
        tl = dev(light0)
        tcube = dev(cube) if use_cube else None
        lib.check(cuda.dfpsr_light_point(C.byref(view), wc.ctypes.data, C.byref(IM(tl)), C.byref(IM(t_normal)), C.byref(IM(t_height)), p.ctypes.data, rad, 1.3, col.ctypes.data, C.byref(IM(tcube)), lib.stream_ptr()))
        e = light0.copy()
        oracle.orc_light_point(C.byref(view), orcbind.ptr(wc), C.byref(OI(e)), C.byref(OI(normal)), C.byref(OI(height)), orcbind.ptr(p), rad, 1.3, orcbind.ptr(col), C.byref(OI(cube if use_cube else None)), 4)
        assert_same_u32(host_u32(tl), e, f"point light {pos} r={rad} cube={use_cube}")


RESIZE_SHAPES = [(83, 61), (40, 30), (83, 100), (83, 20), (120, 61), (31, 61), (160, 130), (200, 122), (17, 200), (300, 45), (1, 1)]


@pytest.mark.parametrize("sampler", [0, 1])
def test_filter_resize(cuda, oracle, sampler):
    rng = np.random.default_rng(6)
    src = rand_rgba(rng, 61, 83)
    ts = dev(src)
    import torch
    for nw, nh in RESIZE_SHAPES:
        for sub in (0, 1):
            tt = dev(np.zeros((nh, nw), np.uint32))
            need = cuda.dfpsr_filter_resize_scratch_bytes(83, 61, nw, nh)
            scratch = torch.zeros(max(need // 4, 1), dtype=torch.int32, device="cuda")
            lib.check(cuda.dfpsr_filter_resize(C.byref(IM(tt)), C.byref(IM(ts)), sampler, sub, scratch.data_ptr(), lib.stream_ptr()))
            e, es = np.zeros((nh, nw), np.uint32), np.zeros(nw * 61 + 4, np.uint32)
            oracle.orc_filter_resize(C.byref(OI(e)), C.byref(OI(src)), sampler, sub, orcbind.ptr(es))
            assert_same_u32(host_u32(tt), e, f"resize to {nw}x{nh} sampler={sampler} sub={sub}")


@pytest.mark.parametrize("sampler", [0, 1])
def test_filter_resize_u8(cuda, oracle, sampler):
    """filter_resize(ImageU8): one byte per pixel, odd strides and unaligned rows included."""
    import torch
    rng = np.random.default_rng(16)
    src = rng.integers(0, 256, (47, 70), dtype=np.uint8)
    ts = torch.from_numpy(src).cuda()
    for nw, nh in [(70, 47), (35, 20), (70, 90), (70, 13), (140, 47), (31, 47), (160, 130), (17, 200), (300, 45), (1, 1), (1024, 600)]:
        tt = torch.zeros((nh, nw), dtype=torch.uint8, device="cuda")
        scratch = torch.zeros(nw * 47 + 4, dtype=torch.uint8, device="cuda")
        lib.check(cuda.dfpsr_filter_resize_u8(C.byref(abi.Image(tt.data_ptr(), nw, nh, nw, 0)), C.byref(abi.Image(ts.data_ptr(), 70, 47, 70, 0)), sampler, scratch.data_ptr(), lib.stream_ptr()))
        e, es = np.zeros((nh, nw), np.uint8), np.zeros(nw * 47 + 4, np.uint8)
        oracle.orc_filter_resize_u8(C.byref(abi.Image(e.ctypes.data, nw, nh, nw, 0)), C.byref(abi.Image(src.ctypes.data, 70, 47, 70, 0)), sampler, orcbind.ptr(es))
        got = tt.cpu().numpy()
        assert np.array_equal(got, e), f"u8 resize to {nw}x{nh} sampler={sampler}: {int((got != e).sum())} bytes differ"


def test_filter_resize_u8_with_strided_images(cuda, oracle):
    """Source and target are windows of larger images (row stride > width, unaligned first pixel): only the window is read and written."""
    import torch
    rng = np.random.default_rng(17)
    big_src = rng.integers(0, 256, (60, 101), dtype=np.uint8)
    big_dst = rng.integers(0, 256, (150, 203), dtype=np.uint8)
    sx, sy, sw, sh = 3, 5, 70, 47
    dx, dy, dw, dh = 7, 9, 160, 130
    ts, td = torch.from_numpy(big_src).cuda(), torch.from_numpy(big_dst).cuda()
    scratch = torch.zeros(dw * sh + 4, dtype=torch.uint8, device="cuda")
    lib.check(cuda.dfpsr_filter_resize_u8(C.byref(abi.Image(td.data_ptr() + dy * 203 + dx, dw, dh, 203, 0)), C.byref(abi.Image(ts.data_ptr() + sy * 101 + sx, sw, sh, 101, 0)),
                                          abi.SAMPLER_LINEAR, scratch.data_ptr(), lib.stream_ptr()))
    expected, es = big_dst.copy(), np.zeros(dw * sh + 4, np.uint8)
    oracle.orc_filter_resize_u8(C.byref(abi.Image(expected.ctypes.data + dy * 203 + dx, dw, dh, 203, 0)), C.byref(abi.Image(big_src.ctypes.data + sy * 101 + sx, sw, sh, 101, 0)),
                                abi.SAMPLER_LINEAR, orcbind.ptr(es))
    got = td.cpu().numpy()
    assert np.array_equal(got, expected)
    assert not np.array_equal(got, big_dst)  # the window changed, and (first assert) nothing outside of it did


def test_filter_map_and_magnify(cuda, oracle):
    rng = np.random.default_rng(7)
    src = rand_rgba(rng, 61, 83)
    ts = dev(src)
    tt = dev(np.zeros((50, 70), np.uint32))
    lib.check(cuda.dfpsr_filter_map(C.byref(IM(tt, 1)), abi.MAP_XOR_PATTERN, None, 0, None, 100, -30, lib.stream_ptr()))
    e = np.zeros((50, 70), np.uint32)
    oracle.orc_filter_map(C.byref(OI(e, 1)), abi.MAP_XOR_PATTERN, None, C.byref(OI(None)), 100, -30)
    assert_same_u32(host_u32(tt), e, "map xor")
    prm = np.array(sandbox_scene.CHAIN_AFFINE, np.int32)
    for sx, sy in [(0, 0), (3, -2), (-5, 4)]:
        tt = dev(np.zeros((61, 83), np.uint32))
        lib.check(cuda.dfpsr_filter_map(C.byref(IM(tt)), abi.MAP_AFFINE, prm.ctypes.data, 8, C.byref(IM(ts)), sx, sy, lib.stream_ptr()))
        e = np.zeros((61, 83), np.uint32)
        oracle.orc_filter_map(C.byref(OI(e)), abi.MAP_AFFINE, orcbind.ptr(prm), C.byref(OI(src)), sx, sy)
        assert_same_u32(host_u32(tt), e, f"map affine start=({sx},{sy})")
    assert cuda.dfpsr_filter_map(C.byref(IM(tt)), 99, None, 0, None, 0, 0, lib.stream_ptr()) != 0
    for pw, ph, tw, th, pack in [(2, 2, 166, 122, 0), (3, 3, 200, 100, 0), (4, 2, 100, 200, 0), (5, 5, 500, 400, 1), (8, 8, 300, 300, 0), (1, 1, 50, 50, 2)]:
        init = rand_rgba(rng, th, tw)
        tt = dev(init)
        lib.check(cuda.dfpsr_filter_block_magnify(C.byref(IM(tt, pack)), C.byref(IM(ts)), pw, ph, lib.stream_ptr()))
        e = init.copy()
        oracle.orc_filter_block_magnify(C.byref(OI(e, pack)), C.byref(OI(src)), pw, ph)
        assert_same_u32(host_u32(tt), e, f"magnify {pw}x{ph}")


def test_texture_pyramid_and_from_image(cuda, oracle):
    from dfpsr_b200 import scenes
    import torch
    level0 = scenes.checker_texture(128, 3)
    dt = lib.DeviceTexture(level0, 6)
    buf, ot = orcbind.build_texture(level0, 6)
    assert (dt.desc.startOffset, dt.desc.maxLevelMask, dt.desc.totalPixels, dt.desc.maxMipLevel) == (ot.startOffset, ot.maxLevelMask, ot.totalPixels, ot.maxMipLevel)
    assert np.array_equal(dt.pixels.cpu().numpy().view(np.uint32), buf)
    # texture_create_RgbaU8(image, levels) from a non-power-of-two image: bilinear resize to 128x64 then pyramid
    rng = np.random.default_rng(9)
    img = rand_rgba(rng, 50, 100)
    desc = abi.Texture()
    lib.check(cuda.dfpsr_texture_layout(C.byref(desc), 100, 50, 4))
    px = torch.zeros(desc.totalPixels, dtype=torch.int32, device="cuda")
    desc.data = px.data_ptr()
    t_img = dev(img)
    lib.check(cuda.dfpsr_texture_from_image(C.byref(desc), C.byref(IM(t_img)), lib.stream_ptr()))
    ot = abi.Texture()
    oracle.orc_texture_layout(C.byref(ot), 100, 50, 4)
    ebuf = np.zeros(ot.totalPixels, np.uint32)
    level = ebuf[ot.startOffset:].reshape(64, 128)
    scratch = np.zeros(128 * 50 + 4, np.uint32)
    oracle.orc_filter_resize(C.byref(OI(level)), C.byref(OI(img)), 1, 0, orcbind.ptr(scratch))
    oracle.orc_texture_generate_pyramid(orcbind.ptr(ebuf), C.byref(ot))
    assert np.array_equal(px.cpu().numpy().view(np.uint32), ebuf)


@pytest.mark.parametrize("batched", [True, False])
def test_sandbox_frame_matches_oracle(cuda, oracle, batched):
    """Small Sandbox frame end to end (compositing, directed light, shadowed point lights, blend) vs the oracle."""
    sb = sandbox_scene.build(320, 240, lights=4, seed=8, sprites=25, casters=3)
    expected = sandbox_scene.run_oracle(oracle, sb)
    gpu = sandbox_scene.CudaSandbox(cuda, sb)
    gpu.composite(batched=batched)
    cubes = gpu.light(keep_cubes=True)
    got = gpu.results()
    for i, (a, b) in enumerate(zip(cubes, expected["cubes"])):
        assert_same_u32(bits(a), bits(b), f"cube map of light {i}")
    assert_same_u32(bits(got["height"]), bits(expected["height"]), "height")
    for key in ("diffuse", "normal", "light", "color"):
        assert_same_u32(got[key], expected[key], key)


def test_sandbox_frame_batched_shadows(cuda, oracle):
    """All cube maps of the frame in one dfpsr_model_render_depth_batch submission: same light buffer, same colours."""
    sb = sandbox_scene.build(320, 240, lights=4, seed=8, sprites=25, casters=3)
    expected = sandbox_scene.run_oracle(oracle, sb)
    gpu = sandbox_scene.CudaSandbox(cuda, sb)
    gpu.composite()
    gpu.light_batched()
    got = gpu.results()
    for i, b in enumerate(expected["cubes"]):
        assert_same_u32(bits(host_f32(gpu.cubes[i])), bits(b), f"cube map of light {i}")
    for key in ("light", "color"):
        assert_same_u32(got[key], expected[key], key)


@pytest.mark.parametrize("size", [(320, 240), (413, 131), (1100, 90)])
def test_sandbox_frame_fused_light(cuda, oracle, size):
    """dfpsr_light_frame: directed + point lights + blend in one kernel == the separate passes, bit for bit. Widths above 1024 exercise
    more than four pixels per thread, 20 lights exercise two shared-memory groups."""
    sb = sandbox_scene.build(size[0], size[1], lights=20 if size[0] > 1000 else 5, seed=9, sprites=12, casters=3)
    expected = sandbox_scene.run_oracle(oracle, sb)
    gpu = sandbox_scene.CudaSandbox(cuda, sb)
    gpu.composite()
    gpu.light_fused()
    got = gpu.results()
    assert (expected["light"] != 0).mean() > 0.3
    for key in ("light", "color"):
        assert_same_u32(got[key], expected[key], key)
    # no directed light: the light buffer starts black (ref: SDK/SpriteEngine/spriteAPI.cpp:783-786); no colour target: light only
    view = sandbox_scene.ortho_view()
    tl = dev(np.full((size[1], size[0]), 0x12345678, np.uint32))
    lib.check(cuda.dfpsr_light_frame(C.byref(view), sb["worldCenter"].ctypes.data, None, None, C.byref(IM(tl)), C.byref(IM(gpu.N)), C.byref(IM(gpu.H)), None, 0, None, 0, lib.stream_ptr()))
    assert np.all(host_u32(tl) == 0)


def test_sandbox_800x600_golden(cuda):
    """BASELINE config 2 at full size against hashes produced by the compiled reference."""
    golden = json.load(open(os.path.join(GOLDEN_DIR, "sandbox.json")))["sandbox_800x600_16"]
    sb = sandbox_scene.build(800, 600, lights=16, seed=5)
    gpu = sandbox_scene.CudaSandbox(cuda, sb)
    gpu.composite()
    cubes = gpu.light(keep_cubes=True)
    got = gpu.results()
    assert sha(cubes[0]) == golden["cube0_sha256"]
    assert sha(got["light"]) == golden["light_sha256"]
    assert sha(got["color"]) == golden["color_sha256"]
    gpu.L.zero_()
    gpu.light_batched()
    got = gpu.results()
    assert sha(got["light"]) == golden["light_sha256"]
    gpu.L.zero_()
    gpu.C.zero_()
    gpu.light_fused()
    got = gpu.results()
    assert sha(host_f32(gpu.cubes[0])) == golden["cube0_sha256"]
    assert sha(got["light"]) == golden["light_sha256"]
    assert sha(got["color"]) == golden["color_sha256"]


def test_filter_chain_8192_golden(cuda):
    """BASELINE config 5 at full size (8192x8192 map + bilinear resizes) against the reference's hashes."""
    import torch
    golden = json.load(open(os.path.join(GOLDEN_DIR, "filters.json")))["filter_chain_8192"]
    size = 8192
    s = lib.stream_ptr()
    src = torch.empty((size, size), dtype=torch.int32, device="cuda")
    lib.check(cuda.dfpsr_filter_map(C.byref(IM(src)), abi.MAP_XOR_PATTERN, None, 0, None, 0, 0, s))
    assert sha(host_u32(src)) == golden["source_sha256"]
    mapped = torch.empty_like(src)
    prm = np.array(sandbox_scene.CHAIN_AFFINE, np.int32)
    lib.check(cuda.dfpsr_filter_map(C.byref(IM(mapped)), abi.MAP_AFFINE, prm.ctypes.data, 8, C.byref(IM(src)), 0, 0, s))
    assert sha(host_u32(mapped)) == golden["mapped_sha256"]

    def resize(source, w, h, sampler):
        out = torch.empty((h, w), dtype=torch.int32, device="cuda")
        need = cuda.dfpsr_filter_resize_scratch_bytes(source.shape[1], source.shape[0], w, h)
        scratch = torch.empty(max(need // 4, 1), dtype=torch.int32, device="cuda")
        lib.check(cuda.dfpsr_filter_resize(C.byref(IM(out)), C.byref(IM(source)), sampler, 0, scratch.data_ptr(), s))
        return out

    half = resize(mapped, 4096, 4096, abi.SAMPLER_LINEAR)
    assert sha(host_u32(half)) == golden["down_4096_sha256"]
    assert sha(host_u32(resize(mapped, 5000, 3000, abi.SAMPLER_LINEAR))) == golden["odd_5000x3000_sha256"]
    assert sha(host_u32(resize(half, 8192, 8192, abi.SAMPLER_LINEAR))) == golden["up_8192_sha256"]
    assert sha(host_u32(resize(mapped, 3000, 5000, abi.SAMPLER_NEAREST))) == golden["nearest_3000x5000_sha256"]


@pytest.mark.parametrize("shape", [(84, 62), (256, 128), (10, 6), (1030, 70)])
@pytest.mark.parametrize("packs", [(0, 0), (1, 0), (0, 3)])
def test_filter_resize_exact_half(cuda, oracle, shape, packs):
    """Bilinear resize to exactly half the size takes the streaming kernel (two byte averages): same bits as the general path."""
    rng = np.random.default_rng(16)
    sw, sh = shape
    src = rand_rgba(rng, sh, sw)
    ts = dev(src)
    tt = dev(np.zeros((sh // 2, sw // 2), np.uint32))
    lib.check(cuda.dfpsr_filter_resize(C.byref(IM(tt, packs[1])), C.byref(IM(ts, packs[0])), abi.SAMPLER_LINEAR, 0, None, lib.stream_ptr()))
    e = np.zeros((sh // 2, sw // 2), np.uint32)
    oracle.orc_filter_resize(C.byref(OI(e, packs[1])), C.byref(OI(src, packs[0])), abi.SAMPLER_LINEAR, 0, None)
    assert_same_u32(host_u32(tt), e, f"half resize {shape} packs={packs}")


@pytest.mark.parametrize("packs", [(0, 0), (2, 1)])
def test_filter_map_affine_streaming(cuda, oracle, packs):
    """Affine map whose reads all fall inside the source (the streaming kernel), incl. a window of a larger source and ragged widths."""
    rng = np.random.default_rng(17)
    src = rand_rgba(rng, 70, 131)
    ts = dev(src)
    prm = np.array([3, -2, 1, 0, -100, 400, 7, 128], np.int32)
    for (tw, th, sx, sy) in [(131, 70, 0, 0), (64, 33, 5, 7), (1, 1, 130, 69), (127, 19, 4, 51)]:
        tt = dev(np.zeros((th, tw), np.uint32))
        lib.check(cuda.dfpsr_filter_map(C.byref(IM(tt, packs[1])), abi.MAP_AFFINE, prm.ctypes.data, 8, C.byref(IM(ts, packs[0])), sx, sy, lib.stream_ptr()))
        e = np.zeros((th, tw), np.uint32)
        oracle.orc_filter_map(C.byref(OI(e, packs[1])), abi.MAP_AFFINE, orcbind.ptr(prm), C.byref(OI(src, packs[0])), sx, sy)
        assert_same_u32(host_u32(tt), e, f"affine map {tw}x{th} at ({sx},{sy})")


def test_reference_rsqrt_fast_path_equals_the_literal_expression():
    """The point light's reciprocal square root is (float)(1.0 / sqrt((double)x)) in the reference's scalar build (base/simd.h:4104). The
    kernels evaluate it with a 22-bit seed + two Newton steps in double and fall back to the literal expression near float rounding
    boundaries; dfpsr_selftest_rsqrt compares both on the device. Every float of 44 binades (squared light distances live in a few of
    them) plus zeros, denormals, infinities, NaNs and negative inputs."""
    cuda = lib.load()
    bad = C.c_uint64(1)
    lib.check(cuda.dfpsr_selftest_rsqrt(0x3A000000, 0x50000000 - 0x3A000000, C.byref(bad), lib.stream_ptr()))
    assert bad.value == 0
    for first, count in ((0, 1 << 20), (0x7F000000, 1 << 24), (0x80000000, 1 << 20), (0xBF000000, 1 << 20), (0x00700000, 1 << 21)):
        lib.check(cuda.dfpsr_selftest_rsqrt(first, count, C.byref(bad), lib.stream_ptr()))
        assert bad.value == 0, hex(first)


@pytest.mark.parametrize("packs", [(2, 0), (1, 0), (3, 3), (2, 1)])
def test_filter_resize_pack_orders(cuda, oracle, packs):
    """Bilinear resize with source and target in different pack orders (the channel-by-channel path; when the orders agree the kernels
    interpolate whole pixels byte lane by byte lane) and in an equal non-RGBA order, over every reference code path."""
    import torch
    rng = np.random.default_rng(23)
    src = rand_rgba(rng, 61, 83)
    ts = dev(src)
    for nw, nh in RESIZE_SHAPES:
        tt = dev(np.zeros((nh, nw), np.uint32))
        need = cuda.dfpsr_filter_resize_scratch_bytes(83, 61, nw, nh)
        scratch = torch.zeros(max(need // 4, 1), dtype=torch.int32, device="cuda")
        lib.check(cuda.dfpsr_filter_resize(C.byref(IM(tt, packs[1])), C.byref(IM(ts, packs[0])), abi.SAMPLER_LINEAR, 0, scratch.data_ptr(), lib.stream_ptr()))
        e, es = np.zeros((nh, nw), np.uint32), np.zeros(nw * 61 + 4, np.uint32)
        oracle.orc_filter_resize(C.byref(OI(e, packs[1])), C.byref(OI(src, packs[0])), abi.SAMPLER_LINEAR, 0, orcbind.ptr(es))
        assert_same_u32(host_u32(tt), e, f"resize to {nw}x{nh} packs={packs}")


def test_canvas_show_is_block_magnify_into_host_memory(cuda, oracle):
    """dfpsr_canvas_show (DsrWindow::showCanvas): the device canvas arrives in the window's host canvas magnified by whole pixels, in the
    window's pack order, partial pixels cut and the rest transparent black — exactly filter_blockMagnify of the reference."""
    rng = np.random.default_rng(12)
    src = rng.integers(0, 2 ** 32, (37, 53), dtype=np.uint32)
    ts = lib.to_device(src)
    for scale, (cw, ch), pack in ((3, (160, 100), abi.PACK_BGRA), (2, (106, 74), abi.PACK_RGBA), (1, (53, 37), abi.PACK_RGBA), (1, (53, 37), abi.PACK_ARGB), (4, (90, 200), abi.PACK_ABGR)):
        stride = cw * 4 + 32
        host = np.full((ch, stride // 4), 0xDEADBEEF, np.uint32)
        lib.check(cuda.dfpsr_canvas_show(C.byref(IM(ts)), scale, host.ctypes.data, stride, cw, ch, pack, lib.stream_ptr()))
        expected = np.zeros((ch, cw), np.uint32)
        oracle.orc_filter_block_magnify(C.byref(OI(expected, pack)), C.byref(OI(src)), scale, scale)
        assert np.array_equal(host[:, :cw], expected), (scale, cw, ch, pack)
        assert (host[:, cw:] == 0xDEADBEEF).all()  # the row padding of the window's canvas is not touched
