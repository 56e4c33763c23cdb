"""A deterministic sequence of the 2D draw calls of api/drawAPI.h (rectangles, lines, alpha filter, max alpha, alpha clip, silhouette)
applied to one RGBA8 and one F32 image, through the reference, the C oracle or the CUDA library. TEST INFRASTRUCTURE."""
import ctypes as C
import hashlib

import numpy as np

from dfpsr_b200 import abi, scenes

F = np.float32


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build(seed=3, width=211, height=157, pack=abi.PACK_RGBA):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 2 ** 32, (height, width), dtype=np.uint32)
    depth = rng.random((height, width)).astype(F)
    ops = []
    for _ in range(12):
        ops.append(("rect", int(rng.integers(-40, width)), int(rng.integers(-40, height)), int(rng.integers(0, 120)), int(rng.integers(0, 90)), [int(v) for v in rng.integers(-60, 330, 4)]))
    ops.append(("rect", -5, -5, width + 10, height // 3, [7, 7, 7, 7]))  # a uniform byte pattern (the reference's memset path) over full rows
    ops.append(("rect_f32", 10, 20, 80, 50, 0.0))
    ops.append(("rect_f32", -10, 100, 500, 30, 3.25))
    for _ in range(60):
        x1, y1, x2, y2 = (int(v) for v in rng.integers(-80, max(width, height) + 80, 4))
        ops.append(("line", x1, y1, x2, y2, [int(v) for v in rng.integers(0, 256, 4)]))
    ops += [("line", 5, 9, 150, 9, [255, 0, 0, 255]), ("line", 33, -20, 33, 400, [0, 255, 0, 255]), ("line", 7, 7, 7, 7, [0, 0, 255, 255]),
            ("line", 0, 0, 100, 100, [9, 9, 9, 255]), ("line", 100, 0, 0, 100, [200, 9, 9, 255]), ("line", -50, 10, 300, 80, [1, 2, 3, 4]), ("line", 10, -50, 80, 300, [4, 3, 2, 1])]
    ops.append(("line_f32", 0, 150, 210, 3, 9.5))
    sources = []
    for k in range(6):
        sw, sh = int(rng.integers(20, 140)), int(rng.integers(20, 110))
        src = rng.integers(0, 2 ** 32, (sh, sw), dtype=np.uint32)
        alpha = rng.integers(0, 256, (sh, sw)).astype(np.uint32)
        alpha[rng.random((sh, sw)) < 0.3] = 255
        alpha[rng.random((sh, sw)) < 0.3] = 0
        src = (src & 0x00FFFFFF) | (alpha << 24) if k % 2 == 0 else (src & 0xFFFFFF00) | alpha  # alpha byte for RGBA / ARGB-like orders
        sources.append((src, abi.PACK_RGBA if k % 2 == 0 else abi.PACK_ABGR))
    for k, (src, order) in enumerate(sources):
        left, top = int(rng.integers(-30, width - 10)), int(rng.integers(-30, height - 10))
        kind = ("alpha_filter", "max_alpha", "alpha_clip", "max_alpha_offset", "alpha_filter", "alpha_clip")[k]
        parameter = {"max_alpha_offset": int(rng.integers(-90, 90)), "alpha_clip": int(rng.integers(0, 255))}.get(kind, 0)
        ops.append((kind, k, left, top, parameter))
    silhouettes = []
    for k in range(3):
        sw, sh = int(rng.integers(30, 120)), int(rng.integers(30, 100))
        s = rng.integers(0, 256, (sh, sw)).astype(np.uint8)
        s[rng.random((sh, sw)) < 0.25] = 255
        s[rng.random((sh, sw)) < 0.25] = 0
        silhouettes.append(np.ascontiguousarray(s))
        ops.append(("silhouette", k, int(rng.integers(-20, width - 20)), int(rng.integers(-20, height - 20)), [[255, 128, 0, 255], [10, 200, 90, 140], [300, -5, 77, 0]][k]))
    return {"base": base, "depth": depth, "ops": ops, "sources": sources, "silhouettes": silhouettes, "pack": pack}


def _c4(values):
    return np.array(values, np.int32)


def run_reference(ref, sc):
    lib = ref.lib
    img, dep = ref.rgba(sc["base"], pack=sc["pack"]), ref.f32(sc["depth"])
    sources = [ref.rgba(s, pack=order) for s, order in sc["sources"]]
    silhouettes = [lib.ref_image_create_u8(s.shape[1], s.shape[0], s.ctypes.data) for s in sc["silhouettes"]]
    for op in sc["ops"]:
        k = op[0]
        if k == "rect":
            lib.ref_draw_rectangle_rgba(img, op[1], op[2], op[3], op[4], _c4(op[5]).ctypes.data)
        elif k == "rect_f32":
            lib.ref_draw_rectangle_f32(dep, op[1], op[2], op[3], op[4], op[5])
        elif k == "line":
            lib.ref_draw_line_rgba(img, op[1], op[2], op[3], op[4], _c4(op[5]).ctypes.data)
        elif k == "line_f32":
            lib.ref_draw_line_f32(dep, op[1], op[2], op[3], op[4], op[5])
        elif k == "alpha_filter":
            lib.ref_draw_alpha_filter(img, sources[op[1]], op[2], op[3])
        elif k in ("max_alpha", "max_alpha_offset"):
            lib.ref_draw_max_alpha(img, sources[op[1]], op[2], op[3], op[4])
        elif k == "alpha_clip":
            lib.ref_draw_alpha_clip(img, sources[op[1]], op[2], op[3], op[4])
        elif k == "silhouette":
            lib.ref_draw_silhouette(img, silhouettes[op[1]], _c4(op[4]).ctypes.data, op[2], op[3])
    return ref.read_rgba(img), ref.read_f32(dep)


def _run(call, image_of, img, dep, sources, silhouettes, ops, tail=()):
    for op in ops:
        k = op[0]
        if k == "rect":
            call("draw_rectangle_rgba", C.byref(image_of(img)), op[1], op[2], op[3], op[4], _c4(op[5]).ctypes.data, *tail)
        elif k == "rect_f32":
            call("draw_rectangle_f32", C.byref(image_of(dep)), op[1], op[2], op[3], op[4], op[5], *tail)
        elif k == "line":
            call("draw_line_rgba", C.byref(image_of(img)), op[1], op[2], op[3], op[4], _c4(op[5]).ctypes.data, *tail)
        elif k == "line_f32":
            call("draw_line_f32", C.byref(image_of(dep)), op[1], op[2], op[3], op[4], op[5], *tail)
        elif k == "alpha_filter":
            call("draw_alpha_filter", C.byref(image_of(img)), C.byref(sources[op[1]]), op[2], op[3], *tail)
        elif k in ("max_alpha", "max_alpha_offset"):
            call("draw_max_alpha", C.byref(image_of(img)), C.byref(sources[op[1]]), op[2], op[3], op[4], *tail)
        elif k == "alpha_clip":
            call("draw_alpha_clip", C.byref(image_of(img)), C.byref(sources[op[1]]), op[2], op[3], op[4], *tail)
        elif k == "silhouette":
            call("draw_silhouette", C.byref(image_of(img)), C.byref(silhouettes[op[1]]), _c4(op[4]).ctypes.data, op[2], op[3], *tail)


def run_oracle(oracle, sc):
    import orcbind
    img, dep = sc["base"].copy(), sc["depth"].copy()
    sources = [orcbind.image_of(s, order) for s, order in sc["sources"]]
    silhouettes = [abi.Image(s.ctypes.data, s.shape[1], s.shape[0], s.strides[0], 0) for s in sc["silhouettes"]]
    image_of = lambda a: orcbind.image_of(a, sc["pack"]) if a.dtype == np.uint32 else orcbind.image_of(a)
    _run(lambda name, *args: getattr(oracle, "orc_" + name)(*args), image_of, img, dep, sources, silhouettes, sc["ops"])
    return img, dep


def run_cuda(cuda, lib, sc):
    import torch
    img, dep = lib.to_device(sc["base"]), lib.to_device(sc["depth"])
    keep = [lib.to_device(s) for s, _ in sc["sources"]]
    sources = [lib.image(t, order) for t, (_, order) in zip(keep, sc["sources"])]
    sil_keep = [torch.from_numpy(s).cuda() for s in sc["silhouettes"]]
    silhouettes = [abi.Image(t.data_ptr(), t.shape[1], t.shape[0], t.stride(0), 0) for t in sil_keep]
    image_of = lambda t: lib.image(t, sc["pack"]) if t.dtype == torch.int32 else lib.image(t)
    _run(lambda name, *args: lib.check(getattr(cuda, "dfpsr_" + name)(*args)), image_of, img, dep, sources, silhouettes, sc["ops"], tail=(lib.stream_ptr(),))
    torch.cuda.synchronize()
    return img.cpu().numpy().view(np.uint32), dep.cpu().numpy()


CASES = [(3, 211, 157, abi.PACK_RGBA), (4, 640, 360, abi.PACK_BGRA), (5, 97, 333, abi.PACK_ARGB)]
