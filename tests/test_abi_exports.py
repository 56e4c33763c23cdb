"""CPU checks of the drop-in boundary: the shared library loads, exports every entry point include/dfpsr_b200.h declares,
the ctypes mirrors have the header's layout, and the host-side (no GPU) entry points agree with the oracle and with the
reference's own known-answer tests. No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import orcbind
from dfpsr_b200 import abi, lib, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dfpsr_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dfpsr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_a_plain_c_abi():
    text = open(HEADER).read()
    assert 'extern "C"' in text
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in text and "std::" not in text and "at::" not in text
    names = declared_functions()
    assert len(names) >= 50
    for required in ("dfpsr_renderer_begin", "dfpsr_renderer_give_task", "dfpsr_renderer_end", "dfpsr_model_render", "dfpsr_model_render_depth",
                     "dfpsr_draw_higher", "dfpsr_light_directed", "dfpsr_light_point", "dfpsr_light_blend", "dfpsr_filter_resize", "dfpsr_filter_map"):
        assert required in names


def test_library_exports_every_declared_symbol():
    handle = lib.load()  # sets argtypes for every entry of SIGNATURES; AttributeError if one is missing
    missing = [name for name in declared_functions() if not hasattr(handle, name)]
    assert not missing, f"declared in include/dfpsr_b200.h but not exported: {missing}"
    unbound = [name for name in declared_functions() if name not in lib.SIGNATURES]
    assert not unbound, f"declared but not bound in dfpsr_b200/lib.py: {unbound}"
    assert handle.dfpsr_abi_version() == 1


def test_exported_symbols_are_unmangled():
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in declared_functions():
        assert name in exported


def test_library_is_built_for_sm_100a():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_pod_layouts_match_the_header():
    src = r"""
    #include "dfpsr_b200.h"
    #include <stdio.h>
    #include <stddef.h>
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(dfpsr_transform3d), sizeof(dfpsr_camera), sizeof(dfpsr_polygon), sizeof(dfpsr_projected_point),
             sizeof(dfpsr_triangle), sizeof(dfpsr_image), sizeof(dfpsr_texture), sizeof(dfpsr_ortho_view), sizeof(dfpsr_model), sizeof(dfpsr_sprite_draw), sizeof(dfpsr_host_model));
      printf("%zu %zu %zu\n", offsetof(dfpsr_camera, cullPlanes), offsetof(dfpsr_model, diffuse), offsetof(dfpsr_projected_point, flat));
      return 0; }
    """
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "layout.c")
        open(path, "w").write(src)
        exe = os.path.join(tmp, "layout")
        subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), path, "-o", exe])
        lines = subprocess.check_output([exe], text=True).split("\n")
    sizes = [int(v) for v in lines[0].split()]
    mirrors = [abi.Transform3D, abi.Camera, abi.Polygon, abi.ProjectedPoint, None, abi.Image, abi.Texture, abi.OrthoView, abi.Model, abi.SpriteDraw, abi.HostModel]
    for size, mirror in zip(sizes, mirrors):
        if mirror is not None:
            assert C.sizeof(mirror) == size, mirror.__name__
    assert abi.TRIANGLE_DTYPE.itemsize == sizes[4]
    offsets = [int(v) for v in lines[1].split()]
    assert offsets == [abi.Camera.cullPlanes.offset, abi.Model.diffuse.offset, abi.ProjectedPoint.flat.offset]


def test_compute_entry_points_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    handle = lib.load()
    r = C.c_void_p()
    assert handle.dfpsr_renderer_create(C.byref(r)) != 0
    assert b"no CPU fallback" in handle.dfpsr_last_error()
    assert handle.dfpsr_init(0) != 0
    with pytest.raises(lib.DfpsrError):
        lib.check(handle.dfpsr_init(0))
    ptr, ipc = C.c_void_p(), (C.c_uint8 * 64)()
    assert handle.dfpsr_peer_alloc(C.byref(ptr), 1024, ipc) != 0 and ptr.value is None
    assert handle.dfpsr_peer_open(C.byref(ptr), ipc) != 0
    # a flag wait must carry a time limit: a peer that never signals may not hang the device
    assert handle.dfpsr_peer_wait(C.c_void_p(256), 1, 1, 0, C.c_void_p(512), None) != 0 and b"time limit" in handle.dfpsr_last_error()


@pytest.mark.parametrize("perspective", [True, False])
def test_camera_create_matches_oracle(perspective):
    """dfpsr_camera_create_* is host arithmetic (ref: Camera.h:113-156); it must give the oracle's planes bit for bit."""
    rng = np.random.default_rng(3)
    for i in range(20):
        pos, target = (rng.random(3) * 2 - 1) * 5, (rng.random(3) * 2 - 1) * 5
        params = abi.camera_params(perspective, scenes.look_at_transform(pos, target), 64 + 37 * i, 48 + 21 * i,
                                   width_slope=0.3 + rng.random() * 3, near=0.01 + rng.random(), far=(float("inf") if i % 5 == 0 else 10 + 1000 * rng.random()))
        assert bytes(lib.camera(params)) == bytes(orcbind.camera(params))


def test_camera_is_box_seen_matches_oracle():
    handle, oracle = lib.load(), orcbind.load()
    rng = np.random.default_rng(5)
    seen = set()
    for i in range(300):
        params = abi.camera_params(i % 3 != 0, scenes.look_at_transform((rng.random(3) * 2 - 1) * 4, (rng.random(3) * 2 - 1) * 4), 320, 200,
                                   width_slope=1.0 if i % 3 != 0 else 3.0, far=20.0)
        cam = lib.camera(params)
        lo = ((rng.random(3) * 2 - 1) * 6).astype(np.float32)
        hi = (lo + rng.random(3) * 4).astype(np.float32)
        m2w = abi.Transform3D.make((rng.random(3) * 2 - 1), ((1, 0, 0), (0, 1, 0), (0, 0, 1)))
        got = handle.dfpsr_camera_is_box_seen(C.byref(cam), lo.ctypes.data, hi.ctypes.data, C.byref(m2w))
        expected = oracle.orc_camera_is_box_seen(C.byref(cam), lo.ctypes.data, hi.ctypes.data, C.byref(m2w))
        assert (got != 0) == (expected != 0)
        seen.add(got != 0)
    assert seen == {True, False}


def test_texture_layout_known_answers():
    """ref: test/tests/TextureTest.cpp:22-45 — TextureRgbaU8(4, 4): 16x16 with levels 1, 2, 4, 8, 16."""
    handle, oracle = lib.load(), orcbind.load()
    for layout_fn in (handle.dfpsr_texture_layout, oracle.orc_texture_layout):
        t = abi.Texture()
        layout_fn(C.byref(t), 16, 16, 5)
        assert (t.log2width, t.log2height, t.maxMipLevel) == (4, 4, 4)
        assert t.startOffset == 0b01010101 and t.maxLevelMask == 0b11111111
        assert t.totalPixels == 1 + 4 + 16 + 64 + 256
    t = abi.Texture()
    handle.dfpsr_texture_layout(C.byref(t), 16, 16, 5)
    oracle.orc_texture_layer_offset.restype = C.c_uint32
    oracle.orc_texture_layer_offset.argtypes = [C.POINTER(abi.Texture), C.c_uint32]
    oracle.orc_texture_pixel_offset.restype = C.c_uint32
    oracle.orc_texture_pixel_offset.argtypes = [C.POINTER(abi.Texture), C.c_uint32, C.c_uint32, C.c_uint32]
    # texture_getPixelOffsetToLayer (TextureTest.cpp:30-45)
    assert [oracle.orc_texture_layer_offset(C.byref(t), m) for m in range(5)] == [0b01010101, 0b00010101, 0b00000101, 0b00000001, 0]
    # texture_getPixelOffset (TextureTest.cpp:46-120): (x, y, mip) -> offset, coordinates wrap inside the level
    cases = [((7534, 424, 15), 0), ((25, 85, 4), 0), ((0, 0, 3), 1), ((1, 1, 3), 4), ((246753, 837624, 3), 2), ((6, 9, 3), 3), ((13, 79, 3), 4),
             ((0, 0, 2), 5), ((3, 3, 2), 20), ((65536, 2050, 2), 13), ((991366, 5, 2), 11), ((0, 0, 1), 21), ((7, 3, 1), 52), ((1, 4, 1), 54)]
    for (x, y, mip), expected in cases:
        assert oracle.orc_texture_pixel_offset(C.byref(t), x, y, mip) == expected, (x, y, mip)
    # non-square, fewer levels than fit, and sizes that are rounded up to powers of two
    for w, h, levels in [(1024, 1024, 5), (64, 32, 4), (100, 60, 3), (8, 8, 1), (2, 2, 9)]:
        a, b = abi.Texture(), abi.Texture()
        handle.dfpsr_texture_layout(C.byref(a), w, h, levels)
        oracle.orc_texture_layout(C.byref(b), w, h, levels)
        assert bytes(a) == bytes(b)
    a = abi.Texture()
    handle.dfpsr_texture_layout(C.byref(a), 1024, 1024, 5)
    assert a.totalPixels == 1396736 and a.startOffset == 1396736 - 1024 * 1024  # terrain texture


def test_linear_colour_interpolation_known_answer():
    """ref: test/tests/TextureTest.cpp:13-19 texture_interpolate_color_linear with weights 0, 128, 256, 256."""
    oracle = orcbind.load()
    oracle.orc_interpolate_color_linear.restype = C.c_uint32
    oracle.orc_interpolate_color_linear.argtypes = [C.c_uint32] * 3
    pack = lambda r, g, b, a: r | (g << 8) | (b << 16) | (a << 24)
    lanes_a = [pack(255, 255, 0, 0), pack(175, 84, 253, 150), pack(253, 255, 172, 241), pack(95, 210, 100, 61)]
    lanes_b = [pack(0, 255, 255, 0), pack(215, 162, 71, 139), pack(62, 152, 62, 180), pack(127, 93, 200, 124)]
    expected = [pack(255, 255, 0, 0), pack(195, 123, 162, 144), pack(62, 152, 62, 180), pack(127, 93, 200, 124)]
    for a, b, w, e in zip(lanes_a, lanes_b, [0, 128, 256, 256], expected):
        assert oracle.orc_interpolate_color_linear(a, b, w) == e


def test_pyramid_matches_numpy_statement():
    oracle = orcbind.load()
    level0 = scenes.checker_texture(64, 9)
    buf, t = orcbind.build_texture(level0, 7)
    expected, max_mip = scenes.mip_pyramid(level0, 7)
    assert max_mip == t.maxMipLevel == 6
    assert np.array_equal(buf, expected)
