"""CPU: the oracle's 2D draw calls (draw_rectangle, draw_line, draw_alphaFilter, draw_maxAlpha, draw_alphaClip, draw_silhouette) against
the compiled reference (api/drawAPI.cpp) and the golden fixtures, plus the reference's own known answers for them (test/tests/DrawTest.cpp)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import draw_scene
import mono_scene
import orcbind
from dfpsr_b200 import abi

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "draw.json")


@pytest.mark.parametrize("case", range(len(draw_scene.CASES)))
def test_oracle_draw_calls_match_reference(oracle, ref_scalar, case):
    sc = draw_scene.build(*draw_scene.CASES[case])
    expected_color, expected_depth = draw_scene.run_reference(ref_scalar, sc)
    color, depth = draw_scene.run_oracle(oracle, sc)
    assert np.array_equal(color, expected_color)
    assert np.array_equal(depth.view(np.uint32), expected_depth.view(np.uint32))
    assert (color != sc["base"]).mean() > 0.3


@pytest.mark.parametrize("case", range(len(draw_scene.CASES)))
def test_oracle_draw_calls_match_golden(oracle, case):
    sc = draw_scene.build(*draw_scene.CASES[case])
    color, depth = draw_scene.run_oracle(oracle, sc)
    entry = json.load(open(GOLDEN))["cases"][case]
    assert draw_scene.sha(color) == entry["color_sha256"] and draw_scene.sha(depth) == entry["depth_sha256"]


def from_ascii(alphabet, rows):
    """ref: api/imageAPI.cpp:405-500 image_fromAscii — character i of the alphabet means int(i * 255 / (n - 1))."""
    value = {ch: int(i * (255.0 / (len(alphabet) - 1))) for i, ch in enumerate(alphabet)}
    return np.array([[value[ch] for ch in row] for row in rows], np.int32)


BALL = from_ascii(" .x", [" .xx. ", ".xxxx.", "xxxxxx", "xxxxxx", ".xxxx.", " .xx. "]).astype(np.uint8)  # test/tests/DrawTest.cpp:7-15
LONG = " .,-_':;!+~=^?*abcdefghijklmnopqrstuvwxyz()[]{}|&@#0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"


def channel(image, index):
    return ((image >> (8 * index)) & 255).astype(np.int32)


def test_reference_known_answers_for_rgba_rectangle_and_silhouette(oracle):
    """The reference's own assertions for RGBA drawing, test/tests/DrawTest.cpp:179-297: a white rectangle over the right half of an 8x8
    image, an opaque orange ball and a half-transparent blue ball drawn with draw_silhouette, compared with the ascii images and the
    tolerances (0, 1 and 2) that the reference's test uses."""
    IM = orcbind.image_of
    image = np.zeros((8, 8), np.uint32)
    ball = abi.Image(BALL.ctypes.data, 6, 6, 6, 0)
    oracle.orc_draw_rectangle_rgba(C.byref(IM(image)), 4, 0, 4, 8, np.array([255, 255, 255, 255], np.int32).ctypes.data)
    half = from_ascii(" ,.-x", ["    xxxx"] * 8)
    for c in range(3):
        assert np.array_equal(channel(image, c), half)
    oracle.orc_draw_silhouette(C.byref(IM(image)), C.byref(ball), np.array([255, 127, 0, 255], np.int32).ctypes.data, 1, 1)
    red = from_ascii(" ,.-x", ["    xxxx", "  .xxxxx", " .xxxxxx", " xxxxxxx", " xxxxxxx", " .xxxxxx", "  .xxxxx", "    xxxx"])
    green = from_ascii(" ,.-x", ["    xxxx", "  ,..-xx", " ,....-x", " ......x", " ......x", " ,....-x", "  ,..-xx", "    xxxx"])
    blue = from_ascii(" ,.-x", ["    xxxx", "     .xx", "      .x", "       x", "       x", "      .x", "     .xx", "    xxxx"])
    for c, expected in enumerate((red, green, blue)):
        assert np.abs(channel(image, c) - expected).max() <= 1
    oracle.orc_draw_silhouette(C.byref(IM(image)), C.byref(ball), np.array([0, 0, 255, 127], np.int32).ctypes.data, 3, 3)
    red = from_ascii(LONG, ["    ZZZZ", "  [ZZZZZ", " [ZZZZZZ", " ZZZE[[E", " ZZE[[[[", " [Z[[[[[", "  [[[[[[", "    [[[["])
    green = from_ascii(LONG, ["    ZZZZ", "  g[[DZZ", " g[[[[DZ", " [[[rhhE", " [[rhhh[", " g[hhhr[", "  ghhr[[", "    [[[["])
    blue = from_ascii(LONG, ["    ZZZZ", "     [ZZ", "      [Z", "    g[[Z", "   g[[[Z", "   [[[DZ", "   [[DZZ", "   gZZZZ"])
    for c, expected in enumerate((red, green, blue)):
        assert np.abs(channel(image, c) - expected).max() <= 2


def test_clipped_rectangle_and_lines_small_case(oracle):
    """Hand-checked small case: a rectangle clipped at the upper-left corner, a diagonal line and a horizontal line drawn right to left."""
    IM = orcbind.image_of
    image = np.zeros((4, 6), np.float32)
    oracle.orc_image_fill_f32(C.byref(IM(image)), 3.0)
    oracle.orc_draw_rectangle_f32(C.byref(IM(image)), -2, -1, 4, 3, 9.0)  # covers x 0..1, y 0..1 after clipping
    assert image.tolist() == [[9, 9, 3, 3, 3, 3], [9, 9, 3, 3, 3, 3], [3, 3, 3, 3, 3, 3], [3, 3, 3, 3, 3, 3]]
    oracle.orc_draw_line_f32(C.byref(IM(image)), 1, 0, 4, 3, 1.0)
    oracle.orc_draw_line_f32(C.byref(IM(image)), 5, 2, 0, 2, 7.0)
    assert image.tolist() == [[9, 1, 3, 3, 3, 3], [9, 9, 1, 3, 3, 3], [7, 7, 7, 7, 7, 7], [3, 3, 3, 3, 1, 3]]
    colour = np.zeros((2, 3), np.uint32)
    oracle.orc_draw_rectangle_rgba(C.byref(IM(colour)), 1, 0, 5, 1, np.array([300, -4, 16, 255], np.int32).ctypes.data)  # saturated like image_saturateAndPack
    assert colour.tolist() == [[0, 0xFF1000FF, 0xFF1000FF], [0, 0, 0]]


@pytest.mark.parametrize("seed", mono_scene.SEEDS)
def test_oracle_monochrome_and_mixed_format_draws_match_reference(oracle, ref_scalar, seed):
    """draw_rectangle / draw_line on U8 and U16 images, all thirteen draw_copy overloads (with the reference's conversions, including its
    U16 <- F32 overload that stores the float's first byte) and draw_higher on 16-bit heights with 0, 1 and 2 payloads."""
    sc = mono_scene.build(seed)
    expected, got = mono_scene.run_reference(ref_scalar, sc), mono_scene.run_oracle(oracle, sc)
    assert mono_scene.same(got, expected)
    assert not mono_scene.same(got, mono_scene.results_list(sc["targets"], sc["higher_target"]))  # something was drawn


@pytest.mark.parametrize("seed", mono_scene.SEEDS)
def test_oracle_monochrome_and_mixed_format_draws_match_golden(oracle, seed):
    entry = json.load(open(GOLDEN))["mono"][mono_scene.SEEDS.index(seed)]
    assert mono_scene.sha(mono_scene.run_oracle(oracle, mono_scene.build(seed))) == entry["sha256"]
