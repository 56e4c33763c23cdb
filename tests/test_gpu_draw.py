"""GPU: the 2D draw calls on device images through the C ABI, bit-exact against the oracle and the reference's golden hashes."""
import json
import os

import numpy as np
import pytest

import draw_scene
import mono_scene
from dfpsr_b200 import lib

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "draw.json")


@pytest.mark.parametrize("case", range(len(draw_scene.CASES)))
def test_draw_calls_bit_exact(cuda, oracle, case):
    sc = draw_scene.build(*draw_scene.CASES[case])
    expected_color, expected_depth = draw_scene.run_oracle(oracle, sc)
    color, depth = draw_scene.run_cuda(cuda, lib, sc)
    assert np.array_equal(color, expected_color)
    assert np.array_equal(depth.view(np.uint32), expected_depth.view(np.uint32))
    entry = json.load(open(GOLDEN))["cases"][case]
    assert draw_scene.sha(color) == entry["color_sha256"] and draw_scene.sha(depth) == entry["depth_sha256"]


def test_long_lines_cross_a_large_image(cuda, oracle):
    """Every step of the reference's error-accumulating line in closed form, over thousands of steps and far outside of the image."""
    import ctypes as C
    import torch
    import orcbind
    w, h = 1920, 1080
    rng = np.random.default_rng(9)
    host = np.zeros((h, w), np.uint32)
    dev = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    for k in range(200):
        x1, y1, x2, y2 = (int(v) for v in rng.integers(-3000, 5000, 4))
        colour = np.array([k + 1, 255 - k, (k * 7) & 255, 255], np.int32)
        oracle.orc_draw_line_rgba(C.byref(orcbind.image_of(host)), x1, y1, x2, y2, colour.ctypes.data)
        lib.check(cuda.dfpsr_draw_line_rgba(C.byref(lib.image(dev)), x1, y1, x2, y2, colour.ctypes.data, lib.stream_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), host)
    assert (host != 0).sum() > 20000


@pytest.mark.parametrize("seed", mono_scene.SEEDS)
def test_monochrome_and_mixed_format_draws_bit_exact(cuda, oracle, seed):
    sc = mono_scene.build(seed)
    got = mono_scene.run_cuda(cuda, lib, sc)
    assert mono_scene.same(got, mono_scene.run_oracle(oracle, sc))
    entry = json.load(open(GOLDEN))["mono"][mono_scene.SEEDS.index(seed)]
    assert mono_scene.sha(got) == entry["sha256"]
