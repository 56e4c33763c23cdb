"""Synthetic Sandbox frame (BASELINE config 2): deferred sprite compositing with draw_higher, one directed light,
N shadow-casting point lights with 256x256x6 cube maps rendered by model_renderDepth, and blendLight — the pass
order of SpriteWorldImpl::draw (ref: SDK/SpriteEngine/spriteAPI.cpp:754-816), driven through three back ends:
the compiled reference, the C oracle and the CUDA library. TEST INFRASTRUCTURE (also used by bench.py).
"""
import ctypes as C

import numpy as np

from dfpsr_b200 import abi, scenes

F = np.float32
CHAIN_AFFINE = [2, 1, -1, 1, 0, 10, 255, 0]  # BASELINE config 5 map: (2r, g + 10, 255 - b, a)

# OrthoView 0 of OrthoSystem(cameraTilt = -0.6, pixelsPerTile = 150) — SDK/sandbox/media/Ortho.ini — as produced by the
# reference (ref_ortho_view in oracle/ref_wrap.cpp); float32 bit patterns so the test needs no reference at run time.
VIEW0_BITS = [1060439284, 0, 3207922932, 2147483648, 1065353216, 0, 1060439284, 0, 1060439284,
              1004181219, 2147483648, 0, 0, 0, 3159788189, 0, 1065353216, 3218508445,
              1125509145, 2147483648, 0, 2147483648, 3271557120, 1065353216, 0, 3264789548, 0]
Y_PIXELS_PER_TILE = 128


def ortho_view():
    return abi.OrthoView.from_buffer_copy(np.array(VIEW0_BITS, np.uint32).tobytes())


def view_matrices():
    raw = np.array(VIEW0_BITS, np.uint32).view(F).reshape(3, 3, 3)
    return {"normalToWorld": raw[0], "screenDepthToLight": raw[1], "lightToScreenDepth": raw[2]}


def mat_transform(m, p):
    """ref: math/FMatrix3x3.h:52-58 in float32 (inputs only; exactness is not required here)."""
    p = np.asarray(p, F)
    return (m[0] * p[0] + m[1] * p[1] + m[2] * p[2]).astype(F)


def mat_mul(left, right):
    """ref: math/FMatrix3x3.h:75-77 operator*: rows of left transformed by right."""
    return np.stack([mat_transform(right, left[0]), mat_transform(right, left[1]), mat_transform(right, left[2])]).astype(F)


def box_model(size):
    sx, sy, sz = (F(s) * F(0.5) for s in size)
    pts = np.array([[x, y, z] for x in (-sx, sx) for y in (-sy, sy) for z in (-sz, sz)], F)
    # outward-facing quads (either winding is fine for a closed box: back faces are dropped, front faces drawn)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    poly = np.zeros(len(quads) * 2, abi.POLYGON_DTYPE)
    for i, q in enumerate(quads):
        poly[2 * i]["pointIndices"] = q
        poly[2 * i + 1]["pointIndices"] = q[::-1]  # both windings so that every face casts a shadow
    poly["colors"] = 1.0
    return pts, poly


def make_sprite(rng, w, h, base_height):
    yy, xx = np.mgrid[0:h, 0:w]
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    inside = ((xx - cx) / (w / 2.0)) ** 2 + ((yy - cy) / (h / 2.0)) ** 2 <= 1.0
    height = ((h - yy) / float(Y_PIXELS_PER_TILE) + base_height).astype(F)
    height[~inside] = -np.inf
    nx = np.clip((xx - cx) / (w / 2.0), -1, 1)
    normal = scenes.pack_rgba(np.clip(128 + nx * 100, 0, 255), np.full((h, w), 160), np.clip(128 - 90 * (1 - np.abs(nx)), 0, 255), np.full((h, w), 255))
    tint = rng.integers(60, 256, 3)
    shade = np.clip(0.6 + 0.4 * (1 - np.abs(nx)), 0, 1)
    diffuse = scenes.pack_rgba(tint[0] * shade, tint[1] * shade, tint[2] * shade, np.full((h, w), 255))
    return height, diffuse, normal


def build(width=800, height=600, lights=16, seed=5, sprites=40, casters=5):
    rng = np.random.default_rng(seed)
    m = view_matrices()
    sb = {"width": width, "height": height, "worldCenter": np.array([width // 2, height // 2], np.int32)}
    yy, xx = np.mgrid[0:height, 0:width]
    checker = (((xx // 53) + (yy // 31)) % 2)
    sb["floor_diffuse"] = scenes.pack_rgba(120 + 80 * checker, 130 + 60 * checker, 110 + 30 * checker, np.full((height, width), 255))
    sb["floor_normal"] = scenes.pack_rgba(np.full((height, width), 128), np.full((height, width), 255), np.full((height, width), 128), np.full((height, width), 255))
    sb["floor_height"] = np.zeros((height, width), F)
    sb["sprites"] = []
    for _ in range(sprites):
        w, h = int(rng.integers(30, 110)), int(rng.integers(50, 190))
        hgt, dif, nrm = make_sprite(rng, w, h, float(rng.random() * 0.2))
        sb["sprites"].append({"height": hgt, "diffuse": dif, "normal": nrm, "left": int(rng.integers(-40, width - 20)), "top": int(rng.integers(-60, height - 30)),
                              "offset": float(F(rng.random() * 0.5))})
    sb["directed"] = {"direction": np.array([1, -1, 0], F), "intensity": 0.1, "color": np.array([255, 255, 255], np.int32)}
    grid = int(np.ceil(np.sqrt(lights)))
    sb["lights"] = []
    for i in range(lights):
        gx, gy = i % grid, i // grid
        s = np.array([-2.2 + 4.4 * (gx + 0.5) / grid, 0.6 + 0.8 * rng.random(), -3.0 + 6.0 * (gy + 0.5) / grid], F)  # light space
        pos = mat_transform(m["normalToWorld"], s)  # world space (tiles)
        colour = np.array([int(v) for v in rng.integers(90, 256, 3)], np.int32)
        sb["lights"].append({"position": pos.astype(F), "radius": 4.0, "intensity": 1.0, "color": colour})
    sb["casters"] = []
    for _ in range(casters):
        s = np.array([(rng.random() * 2 - 1) * 2.0, 0.3 + rng.random() * 0.5, (rng.random() * 2 - 1) * 2.5], F)
        sb["casters"].append({"position": mat_transform(m["normalToWorld"], s).astype(F), "model": box_model(rng.random(3) * 0.5 + 0.3)})
    # cube face cameras: Transform3D(FVector3D(), ShadowCubeMapSides[s] * normalToWorld) (ref: spriteAPI.cpp:330-337, :377-384)
    sides = [((1, 0, 0), (0, 1, 0)), ((-1, 0, 0), (0, 1, 0)), ((0, 1, 0), (0, 0, 1)), ((0, -1, 0), (0, 0, 1)), ((0, 0, 1), (0, 1, 0)), ((0, 0, -1), (0, 1, 0))]
    sb["faces"] = []
    for forward, up in sides:
        axes = np.stack(scenes.make_axis_system(forward, up)).astype(F)
        rot = mat_mul(axes, m["normalToWorld"])
        sb["faces"].append(abi.camera_params(True, abi.Transform3D.make((0, 0, 0), rot), 256, 256))
    sb["cube_res"] = 256
    return sb


def caster_transform(caster, light):
    return abi.Transform3D.make((caster["position"] - light["position"]).astype(F), ((1, 0, 0), (0, 1, 0), (0, 0, 1)))


# ------------------------------------------------------------------------------------------------ reference back end

def run_reference(ref, sb, timing=None):
    import refbind
    w, h, res = sb["width"], sb["height"], sb["cube_res"]
    view = ortho_view()
    H, D, N = ref.f32(sb["floor_height"]), ref.rgba(sb["floor_diffuse"]), ref.rgba(sb["floor_normal"])
    for sp in sb["sprites"]:
        ids = (ref.f32(sp["height"]), ref.rgba(sp["diffuse"]), ref.rgba(sp["normal"]))
        ref.lib.ref_draw_higher(H, ids[0], D, ids[1], N, ids[2], sp["left"], sp["top"], sp["offset"])
    L, Cc = ref.rgba(shape=(h, w)), ref.rgba(shape=(h, w))
    d = sb["directed"]
    ref.lib.ref_light_directed(C.byref(view), L, N, refbind.ptr(d["direction"]), d["intensity"], refbind.ptr(d["color"]), 0)
    models = [ref.model(*c["model"]) for c in sb["casters"]]
    cube = ref.f32(shape=(res * 6, res))
    faces = [ref.lib.ref_image_sub(cube, 0, s * res, res, res) for s in range(6)]
    cubes = []
    for light in sb["lights"]:
        ref.lib.ref_image_fill_f32(cube, 0.0)
        for c, mid in zip(sb["casters"], models):
            t = caster_transform(c, light)
            for s in range(6):
                ref.render_depth(mid, sb["faces"][s], faces[s], model_to_world=t)
        cubes.append(ref.read_f32(cube))
        ref.lib.ref_light_point(C.byref(view), refbind.ptr(sb["worldCenter"]), L, N, H, refbind.ptr(light["position"]), light["radius"], light["intensity"], refbind.ptr(light["color"]), cube)
    ref.lib.ref_light_blend(Cc, D, L)
    return {"height": ref.read_f32(H), "diffuse": ref.read_rgba(D), "normal": ref.read_rgba(N), "light": ref.read_rgba(L), "color": ref.read_rgba(Cc), "cubes": cubes}


# ------------------------------------------------------------------------------------------------ oracle back end

def run_oracle(lib, sb):
    import orcbind
    IM = orcbind.image_of
    w, h, res = sb["width"], sb["height"], sb["cube_res"]
    view = ortho_view()
    H, D, N = sb["floor_height"].copy(), sb["floor_diffuse"].copy(), sb["floor_normal"].copy()
    for sp in sb["sprites"]:
        lib.orc_draw_higher(C.byref(IM(H)), C.byref(IM(sp["height"])), C.byref(IM(D)), C.byref(IM(sp["diffuse"])), C.byref(IM(N)), C.byref(IM(sp["normal"])), sp["left"], sp["top"], sp["offset"])
    L, Cc = np.zeros((h, w), np.uint32), np.zeros((h, w), np.uint32)
    d = sb["directed"]
    lib.orc_light_directed(C.byref(view), C.byref(IM(L)), C.byref(IM(N)), orcbind.ptr(d["direction"]), d["intensity"], orcbind.ptr(d["color"]), 0)
    models = [orcbind.model_of(*c["model"]) for c in sb["casters"]]
    cams = [orcbind.camera(p) for p in sb["faces"]]
    cube = np.zeros((res * 6, res), F)
    cubes = []
    for light in sb["lights"]:
        cube[:] = 0.0
        for c, (model, _keep) in zip(sb["casters"], models):
            t = caster_transform(c, light)
            for s in range(6):
                lib.orc_model_render_depth(C.byref(model), C.byref(t), C.byref(IM(cube[s * res:(s + 1) * res])), C.byref(cams[s]))
        cubes.append(cube.copy())
        lib.orc_light_point(C.byref(view), orcbind.ptr(sb["worldCenter"]), C.byref(IM(L)), C.byref(IM(N)), C.byref(IM(H)), orcbind.ptr(light["position"]), light["radius"], light["intensity"], orcbind.ptr(light["color"]), C.byref(IM(cube)), 4)
    lib.orc_light_blend(C.byref(IM(Cc)), C.byref(IM(D)), C.byref(IM(L)))
    return {"height": H, "diffuse": D, "normal": N, "light": L, "color": Cc, "cubes": cubes}


# ------------------------------------------------------------------------------------------------ CUDA back end

class CudaSandbox:
    """Device-resident inputs of a Sandbox frame; frame() runs the passes through the C ABI."""

    def __init__(self, cuda, sb):
        import torch
        from dfpsr_b200 import lib
        self.cuda, self.sb, self.lib = cuda, sb, lib
        self.floor = [lib.to_device(sb["floor_height"]), lib.to_device(sb["floor_diffuse"]), lib.to_device(sb["floor_normal"])]
        self.sprites = [(lib.to_device(sp["height"]), lib.to_device(sp["diffuse"]), lib.to_device(sp["normal"])) for sp in sb["sprites"]]
        w, h, res = sb["width"], sb["height"], sb["cube_res"]
        self.H, self.D, self.N = (torch.empty_like(t) for t in self.floor)
        self.L = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        self.C = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        self.cube = torch.zeros((res * 6, res), dtype=torch.float32, device="cuda")
        self.models = [lib.DeviceModel(*c["model"]) for c in sb["casters"]]
        self.cams = [lib.camera(p) for p in sb["faces"]]
        self.view = ortho_view()
        draws = (abi.SpriteDraw * len(self.sprites))()
        for i, (sp, t) in enumerate(zip(sb["sprites"], self.sprites)):
            draws[i] = abi.SpriteDraw(lib.image(t[0]), lib.image(t[1]), lib.image(t[2]), sp["left"], sp["top"], sp["offset"])
        self.draws = draws

    def composite(self, batched=True):
        cuda, lib = self.cuda, self.lib
        for dst, src in zip((self.H, self.D, self.N), self.floor):
            dst.copy_(src)
        if batched:
            lib.check(cuda.dfpsr_draw_higher_batch(C.byref(lib.image(self.H)), C.byref(lib.image(self.D)), C.byref(lib.image(self.N)), self.draws, len(self.sprites), lib.stream_ptr()))
        else:
            for sp, t in zip(self.sb["sprites"], self.sprites):
                lib.check(cuda.dfpsr_draw_higher(C.byref(lib.image(self.H)), C.byref(lib.image(t[0])), C.byref(lib.image(self.D)), C.byref(lib.image(t[1])), C.byref(lib.image(self.N)), C.byref(lib.image(t[2])), sp["left"], sp["top"], sp["offset"], lib.stream_ptr()))

    def light(self, keep_cubes=False):
        cuda, lib, sb = self.cuda, self.lib, self.sb
        res = sb["cube_res"]
        s = lib.stream_ptr()
        d = sb["directed"]
        IM = lib.image
        lib.check(cuda.dfpsr_light_directed(C.byref(self.view), C.byref(IM(self.L)), C.byref(IM(self.N)), d["direction"].ctypes.data, d["intensity"], d["color"].ctypes.data, 0, s))
        cubes = []
        for light in sb["lights"]:
            lib.check(cuda.dfpsr_image_fill_f32(C.byref(IM(self.cube)), 0.0, s))
            for c, model in zip(sb["casters"], self.models):
                t = caster_transform(c, light)
                for f in range(6):
                    lib.check(cuda.dfpsr_model_render_depth(C.byref(model.desc), C.byref(t), C.byref(IM(self.cube[f * res:(f + 1) * res])), C.byref(self.cams[f]), s))
            if keep_cubes:
                cubes.append(self.cube.cpu().numpy())
            lib.check(cuda.dfpsr_light_point(C.byref(self.view), sb["worldCenter"].ctypes.data, C.byref(IM(self.L)), C.byref(IM(self.N)), C.byref(IM(self.H)), light["position"].ctypes.data, light["radius"], light["intensity"], light["color"].ctypes.data, C.byref(IM(self.cube)), s))
        lib.check(cuda.dfpsr_light_blend(C.byref(IM(self.C)), C.byref(IM(self.D)), C.byref(IM(self.L)), s))
        return cubes

    def prepare_batched(self):
        """Host arrays for dfpsr_model_render_depth_batch: every (light, caster, face) is one task, every (light, face) one target."""
        import torch
        lib, sb = self.lib, self.sb
        res, nl, nc = sb["cube_res"], len(sb["lights"]), len(sb["casters"])
        self.cubes = torch.zeros((nl, res * 6, res), dtype=torch.float32, device="cuda")
        targets = (abi.Image * (nl * 6))(*[lib.image(self.cubes[i][f * res:(f + 1) * res]) for i in range(nl) for f in range(6)])
        n = nl * nc * 6
        models, transforms, cams, target_of = (C.c_void_p * n)(), (abi.Transform3D * n)(), (abi.Camera * n)(), (C.c_int32 * n)()
        k = 0
        for i, light in enumerate(sb["lights"]):
            for c, model in zip(sb["casters"], self.models):
                t = caster_transform(c, light)
                for f in range(6):
                    models[k], transforms[k], cams[k], target_of[k] = C.addressof(model.desc), t, self.cams[f], i * 6 + f
                    k += 1
        self.batch = (models, transforms, cams, target_of, n, targets, nl * 6)

    def light_batched(self):
        """The same passes with all shadow maps of the frame rendered by ONE submission (B200-first order: the reference reuses a single
        cube map and therefore interleaves shadow rendering with the light passes, ref: SDK/SpriteEngine/spriteAPI.cpp:789-810)."""
        cuda, lib, sb = self.cuda, self.lib, self.sb
        if not hasattr(self, "batch"):
            self.prepare_batched()
        s = lib.stream_ptr()
        d = sb["directed"]
        IM = lib.image
        lib.check(cuda.dfpsr_light_directed(C.byref(self.view), C.byref(IM(self.L)), C.byref(IM(self.N)), d["direction"].ctypes.data, d["intensity"], d["color"].ctypes.data, 0, s))
        models, transforms, cams, target_of, n, targets, nt = self.batch
        lib.check(cuda.dfpsr_model_render_depth_batch(models, transforms, cams, target_of, n, targets, nt, 1, 0.0, s))
        for i, light in enumerate(sb["lights"]):
            lib.check(cuda.dfpsr_light_point(C.byref(self.view), sb["worldCenter"].ctypes.data, C.byref(IM(self.L)), C.byref(IM(self.N)), C.byref(IM(self.H)), light["position"].ctypes.data, light["radius"], light["intensity"], light["color"].ctypes.data, C.byref(IM(self.cubes[i])), s))
        lib.check(cuda.dfpsr_light_blend(C.byref(IM(self.C)), C.byref(IM(self.D)), C.byref(IM(self.L)), s))

    def light_fused(self):
        """Shadow maps in one submission, then directed + all point lights + blend in ONE kernel (dfpsr_light_frame)."""
        cuda, lib, sb = self.cuda, self.lib, self.sb
        if not hasattr(self, "batch"):
            self.prepare_batched()
        if not hasattr(self, "frame_lights"):
            d = sb["directed"]
            directed = (abi.DirectedLight * 1)()
            directed[0].direction[:] = [float(v) for v in d["direction"]]
            directed[0].intensity = d["intensity"]
            directed[0].colorRgb[:] = [int(v) for v in d["color"]]
            points = (abi.PointLight * len(sb["lights"]))()
            for i, light in enumerate(sb["lights"]):
                points[i].position[:] = [float(v) for v in light["position"]]
                points[i].radius, points[i].intensity = light["radius"], light["intensity"]
                points[i].colorRgb[:] = [int(v) for v in light["color"]]
                points[i].shadowCubeMap = lib.image(self.cubes[i])
            self.frame_lights = (directed, points)
        s = lib.stream_ptr()
        IM = lib.image
        models, transforms, cams, target_of, n, targets, nt = self.batch
        lib.check(cuda.dfpsr_model_render_depth_batch(models, transforms, cams, target_of, n, targets, nt, 1, 0.0, s))
        directed, points = self.frame_lights
        lib.check(cuda.dfpsr_light_frame(C.byref(self.view), sb["worldCenter"].ctypes.data, C.byref(IM(self.C)), C.byref(IM(self.D)), C.byref(IM(self.L)), C.byref(IM(self.N)), C.byref(IM(self.H)),
                                         directed, len(directed), points, len(points), s))

    def results(self):
        u = lambda t: t.cpu().numpy().view(np.uint32)
        return {"height": self.H.cpu().numpy(), "diffuse": u(self.D), "normal": u(self.N), "light": u(self.L), "color": u(self.C)}
