"""Scene formats feeding the path: dfpsr_import_ply / dfpsr_import_dmf1 (host code, no GPU) against the reference's
importer_loadModel (SDK/SpriteEngine/importer.cpp) and importFromContent_DMF1 (model/format/dmf1.cpp): every point, index, colour and
texture coordinate bit for bit, on synthetic texts (goldens committed) and on the reference's own media files when they are present."""
import glob
import hashlib
import json
import os
import tempfile

import numpy as np
import pytest

import refbind
from dfpsr_b200 import abi, lib
from importer_cases import CASES

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "importers.json")
MEDIA = "/root/reference/Source/SDK"


def digest(points, polygons, parts, filter_):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(points, np.float32).tobytes())
    h.update(np.ascontiguousarray(polygons).tobytes())
    h.update(json.dumps([list(p) for p in parts]).encode())
    h.update(str(filter_).encode())
    return h.hexdigest()


def product(kind, text, options):
    pts, polys, parts, filter_, bounds = lib.import_model(kind, text, **options)
    return pts, polys, [(name, count) for name, _, _, _, count in parts], filter_, [n for _, d, l, _, _ in parts for n in (d, l) if n], bounds


def reference(ref, kind, text, options):
    if kind == "ply":
        with tempfile.NamedTemporaryFile("w", suffix=".ply", delete=False, newline="") as f:
            f.write(text)
        try:
            return ref.import_ply(f.name, options.get("flip_x", False), options.get("axis"))
        finally:
            os.unlink(f.name)
    return ref.import_dmf1(text, options.get("detail_level", 2))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_import_matches_golden(case):
    name, kind, text, options = case
    golden = json.load(open(GOLDEN))[name]
    pts, polys, parts, filter_, names, bounds = product(kind, text, options)
    assert [len(pts), len(polys)] == golden["counts"] and names == golden["texture_names"]
    assert digest(pts, polys, parts, filter_) == golden["sha256"]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_import_matches_reference(case):
    if not refbind.available("scalar"):
        pytest.skip("oracle/_ref is not built")
    name, kind, text, options = case
    ref = refbind.Ref("scalar")
    rp, rpoly, rparts, rfilter, rnames = reference(ref, kind, text, options)
    pts, polys, parts, filter_, names, bounds = product(kind, text, options)
    assert pts.view(np.uint32).tolist() == rp.view(np.uint32).tolist()
    assert polys.tobytes() == rpoly.tobytes()
    if kind == "dmf1":  # a PLY file has no part names: the reference imports into a part its caller names
        assert parts == rparts and names == rnames
    assert filter_ == rfilter


def test_reference_media_files():
    if not refbind.available("scalar") or not os.path.isdir(MEDIA):
        pytest.skip("needs oracle/_ref and the reference's media (this container only)")
    ref = refbind.Ref("scalar")
    files = sorted(glob.glob(os.path.join(MEDIA, "sandbox/media/models/*.ply")) + glob.glob(os.path.join(MEDIA, "cube/media/*.dmf")))
    assert len(files) >= 6
    for path in files:
        text = open(path, encoding="utf-8-sig", newline="").read()
        if path.endswith(".ply"):
            for flip in (False, True):
                rp, rpoly, _, _, _ = ref.import_ply(path, flip)
                pts, polys, _, _, _, _ = product("ply", text, {"flip_x": flip})
                assert len(pts) > 0 and pts.view(np.uint32).tolist() == rp.view(np.uint32).tolist(), path
                assert polys.tobytes() == rpoly.tobytes(), path
        else:
            for detail in (0, 1, 2):
                rp, rpoly, rparts, rfilter, rnames = ref.import_dmf1(text, detail)
                pts, polys, parts, filter_, names, _ = product("dmf1", text, {"detail_level": detail})
                assert pts.view(np.uint32).tolist() == rp.view(np.uint32).tolist(), path
                assert polys.tobytes() == rpoly.tobytes() and parts == rparts and names == rnames and filter_ == rfilter, path
        ref.free_all()


def test_import_errors():
    with pytest.raises(lib.DfpsrError):
        lib.import_model("ply", "not a ply file\nat all\n")
    with pytest.raises(lib.DfpsrError):
        lib.import_model("ply", "ply\nformat binary_little_endian 1.0\nend_header\n")
    with pytest.raises(lib.DfpsrError):
        lib.import_model("dmf1", "DMF2 <Part>")
    pts, polys, parts, _, _ = lib.import_model("dmf1", "DMF1")
    assert len(pts) == 0 and len(polys) == 0 and parts == []


def test_imported_model_bounds_start_at_the_origin():
    """ref: implementation/render/model/Model.cpp:281-288 — the bounding box contains the origin and every point."""
    pts, _, _, _, (mn, mx) = lib.import_model("ply", CASES[0][2])
    assert mn == [min(0.0, float(v)) for v in pts.min(axis=0)] and mx == [max(0.0, float(v)) for v in pts.max(axis=0)]


# ---- the dsr:: shim on top of the importers (dfpsr_b200/host/model_test.cpp, host only)
import subprocess  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL_TEST = os.path.join(ROOT, "dfpsr_b200", "host", "model_test")


def parse_model_test(stdout):
    points, parts, names = [], [], []
    for line in stdout.splitlines():
        f = line.split()
        if f[0] == "p":
            points.append([float.fromhex(v) for v in f[1:4]])
        elif f[0] == "part":
            parts.append([])
        elif f[0] == "i":
            parts[-1].append({"indices": [int(v) for v in f[1:5]], "corners": []})
        elif f[0] == "v":
            parts[-1][-1]["corners"].append([float.fromhex(v) for v in f[1:9]])
        elif f[0] == "textures":
            names.append(line.split("'")[1::2])
    return np.array(points, np.float32).reshape(-1, 3), parts, names


def run_model_test(*args):
    if not os.path.exists(MODEL_TEST):
        pytest.skip("run `make -C dfpsr_b200/host` (done by __graft_entry__.build())")
    out = subprocess.run([MODEL_TEST, *args], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return parse_model_test(out.stdout)


def corners_of(polys, index):
    return [list(map(float, polys["texCoords"][index][v])) + list(map(float, polys["colors"][index][v])) for v in range(4)]


@pytest.mark.parametrize("case", [c for c in CASES if "axis" not in c[3]], ids=[c[0] for c in CASES if "axis" not in c[3]])
def test_shim_importers_match_the_c_abi(case, tmp_path):
    name, kind, text, options = case
    path = tmp_path / ("model.ply" if kind == "ply" else "model.dmf")
    with open(path, "w", newline="") as f:
        f.write(text)
    extra = str(int(options.get("flip_x", False))) if kind == "ply" else str(options.get("detail_level", 2))
    points, parts, names = run_model_test("ply" if kind == "ply" else "dmf", str(path), extra)
    pts, polys, own_parts, _, _ = lib.import_model(kind, text, **options)
    assert points.view(np.uint32).tolist() == pts.view(np.uint32).tolist()
    assert [len(p) for p in parts] == [count for _, _, _, _, count in own_parts]
    flat = [polygon for part in parts for polygon in part]
    for i, polygon in enumerate(flat):
        assert polygon["indices"] == polys["pointIndices"][i].tolist() and polygon["corners"] == corners_of(polys, i)
    if kind == "dmf1":
        assert names == [[d, l] for _, d, l, _, _ in own_parts]


def test_shim_polygon_defaults_match_the_reference():
    """model_addTriangle / model_addQuad (ref: api/modelAPI.cpp:146-154) give white corners and texture coordinates spanning the texture
    (Model.cpp:74-103), not zeros: a textured quad built without model_setTexCoord shows the whole texture."""
    points, parts, _ = run_model_test("defaults")
    expected = [[0, 0, 0, 0, 1, 1, 1, 1], [1, 0, 1, 0, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1, 1, 1], [0, 1, 0, 1, 1, 1, 1, 1]]
    assert [p["indices"] for p in parts[0]] == [[0, 1, 2, -1], [0, 1, 3, 2]]
    assert all(p["corners"] == expected for p in parts[0])
    if refbind.available("scalar"):  # the compiled reference builds the same two polygons through its own API
        ref = refbind.Ref("scalar")
        ply = "ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\nelement face 2\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n1 1 0\n3 0 1 2\n4 0 1 3 2\n"
        with tempfile.NamedTemporaryFile("w", suffix=".ply", delete=False) as f:
            f.write(ply)
        try:
            _, rpoly, _, _, _ = ref.import_ply(f.name)  # importer_loadModel calls model_addTriangle / model_addQuad and only sets colours
        finally:
            os.unlink(f.name)
        assert [corners_of(rpoly, i) for i in range(2)] == [expected, expected]
