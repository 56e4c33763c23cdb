"""Tolerance mode of the tile kernel (dfpsr_set_default_precision(DFPSR_PRECISION_TOLERANCE)): direct evaluation of the interpolation
planes and the hardware reciprocal instead of the reference's replayed addition chains (SURVEY.md §7 hard part 3).

Stated tolerance (BASELINE.json north_star): coverage identical; colour within +-1 LSB per 8-bit channel; depth within DEPTH_REL of the
largest depth of the frame (the reference's running sums drift from the plane by the rounding of up to a thousand additions of values
of that magnitude, so an error bound relative to each pixel's own value does not exist where the plane passes through small values).
Pixels may exceed the colour bound only where a hard decision flips on the last ulps — a depth test between (nearly) coplanar
triangles or the per-quad mip selector at its 2/4/8/16-texel thresholds. The reference's own SSE build differs from its scalar
build in exactly the same way; such pixels are counted, printed and bounded by FLIP_FRACTION.
The comparison runs against the oracle (the reference's scalar flavour) and, when oracle/_ref is present, against the
compiled reference's SSE flavour — the build every x86 user of the reference actually runs."""
import ctypes as C

import numpy as np
import pytest

import orcbind
from dfpsr_b200 import abi, lib, scenes
from gpuutil import CudaScene, bits, dev, host_f32, host_u32

pytestmark = pytest.mark.gpu

DEPTH_REL = 2.0 ** -16   # |depth - reference depth| <= DEPTH_REL x the largest reference depth of the frame (about 256 ulp of that value)
FLIP_FRACTION = 2.0e-4   # pixels allowed to differ by more than 1 LSB (flipped depth ties / mip thresholds), as a fraction of the target


def channel_diff(a, b):
    a8, b8 = a.view(np.uint8).reshape(a.shape + (4,)).astype(np.int16), b.view(np.uint8).reshape(b.shape + (4,)).astype(np.int16)
    return np.abs(a8 - b8).max(axis=-1)


def ulp_diff(a, b):
    ia, ib = bits(a).astype(np.int64), bits(b).astype(np.int64)
    return np.abs(ia - ib)


def compare(name, got_c, got_d, exp_c, exp_d, initial_d):
    covered_got, covered_exp = got_d != initial_d, exp_d != initial_d
    coverage_mismatches = int((covered_got != covered_exp).sum())
    cd = channel_diff(got_c, exp_c)
    over = int((cd > 1).sum())
    ud = ulp_diff(got_d, exp_d)
    scale = float(np.abs(exp_d[np.isfinite(exp_d)]).max())
    ad = np.abs(got_d.astype(np.float64) - exp_d.astype(np.float64))
    ad[~np.isfinite(ad)] = 0.0
    depth_over = int((ad > DEPTH_REL * scale).sum())
    print(f"{name}: coverage mismatches {coverage_mismatches}, pixels differing {int((cd > 0).sum())} (max channel diff {int(cd.max())}), "
          f"pixels over 1 LSB {over}, median / max depth ulp {int(np.median(ud[covered_exp])) if covered_exp.any() else 0} / {int(ud.max())}, "
          f"max |depth error| / frame max depth {ad.max() / scale:.2e}, over the bound {depth_over}, of {got_c.size} pixels")
    assert coverage_mismatches == 0, name
    # a flipped depth tie also changes the depth by more than the ulp bound: both kinds of pixel are counted against the same budget
    assert over <= FLIP_FRACTION * got_c.size, name
    assert depth_over <= FLIP_FRACTION * got_c.size, name
    return over


@pytest.fixture()
def tolerance(cuda):
    lib.check(cuda.dfpsr_set_default_precision(1))
    yield cuda
    lib.check(cuda.dfpsr_set_default_precision(0))


@pytest.mark.parametrize("frame", [0, 7, 23, 41])
def test_terrain_1080p_within_tolerance(tolerance, oracle, frame):
    cuda = tolerance
    sc = scenes.terrain_scene()
    scene = CudaScene(sc["points"], sc["polygons"], diffuse_level0=sc["texture"], diffuse_levels=5)
    w, h = 1920, 1080
    cam = scenes.orbit_camera(frame, w, h)
    c0, d0 = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
    got_c, got_d = scene.render_cuda(cuda, cam, c0, d0)
    exp_c, exp_d, commands = scene.render_oracle(oracle, cam, c0, d0)
    assert commands > 500
    compare(f"terrain frame {frame} vs scalar reference", got_c, got_d, exp_c, exp_d, d0)
    import refbind
    if refbind.available("sse"):
        ref = refbind.Ref("sse")
        tex = ref.texture(sc["texture"], 5)
        model = ref.model(sc["points"], sc["polygons"], diffuse=tex)
        col, dep = ref.rgba(array=c0), ref.f32(array=d0)
        ref.render(model, cam, col, dep)
        sse_c, sse_d = ref.read_rgba(col), ref.read_f32(dep)
        compare(f"terrain frame {frame} vs SSE reference", got_c, got_d, sse_c, sse_d, d0)
        ref.free_all()


@pytest.mark.parametrize("case", [0, 1, 2, 3, 4, 5, 6, 10, 11, 12, 14, 15])
def test_solid_variants_within_tolerance(tolerance, oracle, case):
    """Every solid shader variant, pack order, colour-only / depth-only targets, orthogonal camera, clipping."""
    from test_gpu_raster import CASES, soup_case
    cuda = tolerance
    cfg = CASES[case]
    scene, cam, color, depth = soup_case(900 + case, **cfg)
    got_c, got_d = scene.render_cuda(cuda, cam, color, depth, cfg["pack"])
    exp_c, exp_d, commands = scene.render_oracle(oracle, cam, color, depth, cfg["pack"])
    assert commands > 10
    if depth is not None and color is not None:
        compare(f"soup case {case}", got_c, got_d, exp_c, exp_d, depth)
    elif depth is not None:
        assert ((got_d != depth) == (exp_d != depth)).all()
        ad = np.abs(got_d.astype(np.float64) - exp_d.astype(np.float64))
        assert (ad > DEPTH_REL * float(np.abs(exp_d).max())).sum() <= FLIP_FRACTION * got_d.size * 10
    else:
        # no depth buffer: the last command covering a pixel wins, no ties to flip
        cd = channel_diff(got_c, exp_c)
        assert (cd > 1).sum() <= FLIP_FRACTION * got_c.size * 10


def test_alpha_frames_stay_exact_in_tolerance_mode(tolerance, oracle):
    """Frames that hold alpha-filtered commands always take the exact immediate kernel."""
    from test_gpu_raster import CASES, soup_case
    cuda = tolerance
    cfg = CASES[7]
    scene, cam, color, depth = soup_case(77, **cfg)
    got_c, got_d = scene.render_cuda(cuda, cam, color, depth, cfg["pack"])
    exp_c, exp_d, _ = scene.render_oracle(oracle, cam, color, depth, cfg["pack"])
    assert np.array_equal(bits(got_d), bits(exp_d)) and np.array_equal(got_c, exp_c)


def test_odd_height_quirk_in_tolerance_mode(tolerance, oracle):
    """The reference's repeated upper row in the last row pair of an odd-height target is a coverage rule: it must hold in tolerance mode too."""
    from test_gpu_raster import CASES, soup_case
    cuda = tolerance
    for size in [(33, 35), (321, 181), (77, 3)]:
        scene, cam, color, depth = soup_case(501, w=size[0], h=size[1], **CASES[1])
        got_c, got_d = scene.render_cuda(cuda, cam, color, depth, 0)
        exp_c, exp_d, _ = scene.render_oracle(oracle, cam, color, depth, 0)
        assert ((got_d != depth) == (exp_d != depth)).all()
        cd = channel_diff(got_c, exp_c)
        assert (cd > 1).sum() <= max(3, FLIP_FRACTION * got_c.size * 10), size
