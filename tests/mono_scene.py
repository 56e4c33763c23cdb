"""8-bit / 16-bit monochrome images, the thirteen draw_copy overloads and draw_higher on 16-bit heights (api/drawAPI.h:68-103, :128-135)
through the reference, the C oracle or the CUDA library. TEST INFRASTRUCTURE."""
import ctypes as C
import hashlib

import numpy as np

from dfpsr_b200 import abi

F = np.float32
DTYPE = {abi.FORMAT_U8: np.uint8, abi.FORMAT_U16: np.uint16, abi.FORMAT_F32: np.float32, abi.FORMAT_RGBA_U8: np.uint32}
# the overloads the reference has: (target format, source format)
COPIES = [(4, 4), (1, 1), (2, 2), (3, 3), (4, 1), (4, 2), (4, 3), (1, 3), (1, 2), (2, 1), (2, 3), (3, 1), (3, 2)]


def sha(arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def random_image(rng, fmt, h, w):
    if fmt == abi.FORMAT_U8:
        return rng.integers(0, 256, (h, w)).astype(np.uint8)
    if fmt == abi.FORMAT_U16:
        a = rng.integers(0, 65536, (h, w)).astype(np.uint16)
        a[rng.random((h, w)) < 0.4] //= 200  # plenty of values below 256
        return a
    if fmt == abi.FORMAT_F32:
        a = (rng.random((h, w)) * 400 - 60).astype(F)
        a[rng.random((h, w)) < 0.1] = np.nan
        a[rng.random((h, w)) < 0.1] = F(254.5)
        a[rng.random((h, w)) < 0.1] = F(0.5)
        return a
    return rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)


def build(seed=8, width=150, height=91):
    rng = np.random.default_rng(seed)
    sc = {"targets": {f: random_image(rng, f, height, width) for f in DTYPE}, "copies": [], "ops": [], "pack": abi.PACK_BGRA}
    for tf, sf in COPIES:
        sh, sw = int(rng.integers(20, 120)), int(rng.integers(20, 170))
        sc["copies"].append((tf, sf, random_image(rng, sf, sh, sw), int(rng.integers(-30, width - 20)), int(rng.integers(-30, height - 20))))
    for fmt in (abi.FORMAT_U8, abi.FORMAT_U16):
        for _ in range(6):
            sc["ops"].append(("rect", fmt, int(rng.integers(-20, width)), int(rng.integers(-20, height)), int(rng.integers(0, 90)), int(rng.integers(0, 60)), int(rng.integers(-500, 70000))))
        for _ in range(25):
            x1, y1, x2, y2 = (int(v) for v in rng.integers(-60, max(width, height) + 60, 4))
            sc["ops"].append(("line", fmt, x1, y1, x2, y2, int(rng.integers(-100, 70000))))
    # draw_higher on 16-bit heights with 0, 1 and 2 RGBA payloads
    sc["higher"] = []
    for payloads in (0, 1, 2):
        sh, sw = int(rng.integers(30, 100)), int(rng.integers(30, 140))
        hs = rng.integers(0, 65536, (sh, sw)).astype(np.uint16)
        hs[rng.random((sh, sw)) < 0.3] = 0
        sc["higher"].append({"payloads": payloads, "height": hs, "a": random_image(rng, 4, sh, sw), "b": random_image(rng, 4, sh, sw),
                             "left": int(rng.integers(-20, width - 30)), "top": int(rng.integers(-20, height - 30)), "offset": int(rng.integers(-30000, 30000))})
    sc["higher_target"] = {"height": (rng.integers(0, 65536, (height, width)) // 2).astype(np.uint16), "a": random_image(rng, 4, height, width), "b": random_image(rng, 4, height, width)}
    return sc


def results_list(targets, higher):
    return [targets[f] for f in sorted(targets)] + [higher["height"], higher["a"], higher["b"]]


def run_reference(ref, sc):
    lib = ref.lib

    def make(fmt, a, pack=abi.PACK_RGBA):
        if fmt == abi.FORMAT_U8:
            return lib.ref_image_create_u8(a.shape[1], a.shape[0], np.ascontiguousarray(a).ctypes.data)
        if fmt == abi.FORMAT_U16:
            return lib.ref_image_create_u16(a.shape[1], a.shape[0], np.ascontiguousarray(a).ctypes.data)
        if fmt == abi.FORMAT_F32:
            return ref.f32(a)
        return ref.rgba(a, pack=pack)

    def read(fmt, iid, shape):
        if fmt in (abi.FORMAT_U8, abi.FORMAT_U16):
            out = np.zeros(shape, DTYPE[fmt])
            lib.ref_image_read_mono(iid, out.ctypes.data)
            return out
        return ref.read_f32(iid) if fmt == abi.FORMAT_F32 else ref.read_rgba(iid)

    targets = {f: make(f, a, sc["pack"]) for f, a in sc["targets"].items()}
    for tf, sf, src, left, top in sc["copies"]:
        lib.ref_draw_copy_formats(targets[tf], make(sf, src), left, top)
    for op in sc["ops"]:
        if op[0] == "rect":
            lib.ref_draw_rectangle_mono(targets[op[1]], op[2], op[3], op[4], op[5], op[6])
        else:
            lib.ref_draw_line_mono(targets[op[1]], op[2], op[3], op[4], op[5], op[6])
    ht = sc["higher_target"]
    H, A, B = make(abi.FORMAT_U16, ht["height"]), ref.rgba(ht["a"]), ref.rgba(ht["b"], pack=abi.PACK_ARGB)
    for h in sc["higher"]:
        hs, sa, sb = make(abi.FORMAT_U16, h["height"]), ref.rgba(h["a"], pack=abi.PACK_ABGR), ref.rgba(h["b"])
        lib.ref_draw_higher_u16(H, hs, A if h["payloads"] >= 1 else -1, sa if h["payloads"] >= 1 else -1, B if h["payloads"] >= 2 else -1, sb if h["payloads"] >= 2 else -1, h["left"], h["top"], h["offset"])
    shape = sc["targets"][abi.FORMAT_U8].shape
    out_targets = {f: read(f, targets[f], shape) for f in targets}
    return results_list(out_targets, {"height": read(abi.FORMAT_U16, H, shape), "a": ref.read_rgba(A), "b": ref.read_rgba(B)})


def _run(call, image_of, sc, targets, higher_target, make_source, tail=()):
    for tf, sf, src, left, top in sc["copies"]:
        keep = make_source(src)
        call("draw_copy_formats", C.byref(image_of(targets[tf], sc["pack"] if tf == 4 else 0)), tf, C.byref(image_of(keep, 0)), sf, left, top, *tail)
    for op in sc["ops"]:
        name = "draw_rectangle_mono" if op[0] == "rect" else "draw_line_mono"
        call(name, C.byref(image_of(targets[op[1]], 0)), op[1], op[2], op[3], op[4], op[5], op[6], *tail)
    H, A, B = higher_target
    for h in sc["higher"]:
        hs, sa, sb = make_source(h["height"]), make_source(h["a"]), make_source(h["b"])
        null = None
        call("draw_higher_u16", C.byref(image_of(H, 0)), C.byref(image_of(hs, 0)),
             C.byref(image_of(A, abi.PACK_RGBA)) if h["payloads"] >= 1 else null, C.byref(image_of(sa, abi.PACK_ABGR)) if h["payloads"] >= 1 else null,
             C.byref(image_of(B, abi.PACK_ARGB)) if h["payloads"] >= 2 else null, C.byref(image_of(sb, abi.PACK_RGBA)) if h["payloads"] >= 2 else null, h["left"], h["top"], h["offset"], *tail)


def run_oracle(oracle, sc):
    def image_of(a, pack):
        return abi.Image(a.ctypes.data, a.shape[1], a.shape[0], a.strides[0], pack)
    targets = {f: a.copy() for f, a in sc["targets"].items()}
    ht = {k: v.copy() for k, v in sc["higher_target"].items()}
    _run(lambda name, *args: getattr(oracle, "orc_" + name)(*args), image_of, sc, targets, (ht["height"], ht["a"], ht["b"]), lambda a: np.ascontiguousarray(a))
    return results_list(targets, ht)


def run_cuda(cuda, lib, sc):
    import torch

    def dev(a):
        if a.dtype == np.uint32:
            return torch.from_numpy(a.view(np.int32).copy()).cuda()
        if a.dtype == np.uint16:
            return torch.from_numpy(a.view(np.int16).copy()).cuda()
        return torch.from_numpy(np.ascontiguousarray(a).copy()).cuda()

    def image_of(t, pack):
        return abi.Image(t.data_ptr(), t.shape[1], t.shape[0], t.stride(0) * t.element_size(), pack)

    def host(t, dtype):
        return t.cpu().numpy().view(dtype)

    targets = {f: dev(a) for f, a in sc["targets"].items()}
    ht = {k: dev(v) for k, v in sc["higher_target"].items()}
    keep = []

    def make_source(a):
        keep.append(dev(a))
        return keep[-1]

    _run(lambda name, *args: lib.check(getattr(cuda, "dfpsr_" + name)(*args)), image_of, sc, targets, (ht["height"], ht["a"], ht["b"]), make_source, tail=(lib.stream_ptr(),))
    torch.cuda.synchronize()
    out = {f: host(t, DTYPE[f]) for f, t in targets.items()}
    return results_list(out, {"height": host(ht["height"], np.uint16), "a": host(ht["a"], np.uint32), "b": host(ht["b"], np.uint32)})


def same(a, b):
    """Bitwise comparison (NaN-safe for the float image)."""
    return all(x.shape == y.shape and np.array_equal(np.ascontiguousarray(x).view(np.uint8), np.ascontiguousarray(y).view(np.uint8)) for x, y in zip(a, b))


SEEDS = [8, 9]
