"""A scripted Sandbox session for the sprite world: synthetic sprite types and dense models, a passive
world, temporary sprites / models / lights, camera moves, removals, a camera rotation and a resize — driven through three back ends:

  * the compiled reference     spriteWorld_* of SDK/SpriteEngine/spriteAPI.cpp through oracle/ref_wrap.cpp (this container only)
  * plan + oracle              the product's HOST planner (dfpsr_sprite_world_plan_frame, no GPU) replayed with the C oracle's pixel loops
  * CUDA                       dfpsr_sprite_world_draw on the device

TEST INFRASTRUCTURE. The same script also produces tests/golden/sprite_world.json (tests/golden/make_golden.py).
"""
import ctypes as C
import hashlib

import numpy as np

from dfpsr_b200 import abi, scenes

F = np.float32
TILT, PIXELS_PER_TILE, SHADOW_RES = -0.6, 64, 64
MINI = 1024


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------------------------------------ assets

def _box(size, offset=(0.0, 0.0, 0.0)):
    sx, sy, sz = (F(s) * F(0.5) for s in size)
    pts = np.array([[x, y, z] for x in (-sx, sx) for y in (-sy, sy) for z in (-sz, sz)], F) + np.array(offset, F)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    return pts, quads


def make_sprite_type(rng, frame_w, frame_h, frames, center, min_bound, max_bound, shadow_box=None):
    """Atlas = `frames` rows of [colour | height | normal]. Returns the description shared by the three back ends."""
    atlas = np.zeros((frame_h * frames, frame_w * 3), np.uint32)
    yy, xx = np.mgrid[0:frame_h, 0:frame_w]
    for f in range(frames):
        cx, cy = (frame_w - 1) / 2.0 + (f % 3 - 1) * 1.5, (frame_h - 1) / 2.0
        inside = ((xx - cx) / (frame_w / 2.0 - 1)) ** 2 + ((yy - cy) / (frame_h / 2.0 - 1)) ** 2 <= 1.0
        tint = rng.integers(70, 256, 3)
        shade = np.clip(0.55 + 0.45 * (1 - np.abs((xx - cx) / (frame_w / 2.0))), 0, 1)
        colour = scenes.pack_rgba(tint[0] * shade, tint[1] * shade, tint[2] * shade, np.where(inside, 255, int(rng.integers(0, 128))))
        height = scenes.pack_rgba(np.clip((frame_h - yy) * 255.0 / frame_h + f, 0, 255), np.full_like(xx, 7), np.full_like(xx, 9), np.full_like(xx, 255))
        nx = np.clip((xx - cx) / (frame_w / 2.0), -1, 1)
        normal = scenes.pack_rgba(np.clip(128 + nx * 100, 0, 255), np.full_like(xx, 170), np.clip(128 - 80 * (1 - np.abs(nx)), 0, 255), np.full_like(xx, 255))
        rows = slice(f * frame_h, (f + 1) * frame_h)
        atlas[rows, 0:frame_w], atlas[rows, frame_w:2 * frame_w], atlas[rows, 2 * frame_w:] = colour, height, normal
    t = {"atlas": atlas, "frame_w": frame_w, "frame_h": frame_h, "frames": frames, "center": center, "min": [F(v) for v in min_bound], "max": [F(v) for v in max_bound],
         "points": None, "indices": None}
    if shadow_box is not None:
        pts, quads = _box(*shadow_box)
        idx = []
        for q in quads:
            for tri in ((q[0], q[1], q[2]), (q[0], q[2], q[3]), (q[2], q[1], q[0]), (q[3], q[2], q[0])):  # both windings: every face casts
                idx.extend(tri)
        t["points"], t["indices"] = pts.astype(F), np.array(idx, np.int32)
    return t


def sprite_ini(t):
    """The reference's <name>.ini (ref: spriteAPI.cpp:56-101). Bounds are dyadic so that the decimal text parses back exactly."""
    fmt = lambda v: repr(float(v))
    lines = [f"CenterX={t['center'][0]}", f"CenterY={t['center'][1]}", f"FrameRows={t['frames']}", "PropertyColumns=3",
             "MinBound=" + ",".join(fmt(v) for v in t["min"]), "MaxBound=" + ",".join(fmt(v) for v in t["max"])]
    if t["points"] is not None:
        lines.append("Points=" + ",".join(fmt(v) for v in t["points"].reshape(-1)))
        lines.append("TriangleIndices=" + ",".join(str(int(v)) for v in t["indices"]))
    return "\n".join(lines) + "\n"


def make_model_type(rng, kind):
    """A vertex-coloured model (quads and triangles) plus a box shadow model."""
    if kind == 0:  # a stepped pyramid
        pts, polys = [], []
        for level, (half, y0, y1) in enumerate(((0.6, 0.0, 0.25), (0.4, 0.25, 0.5), (0.2, 0.5, 0.9))):
            base = len(pts)
            for y in (y0, y1):
                for x, z in ((-half, -half), (half, -half), (half, half), (-half, half)):
                    pts.append((x, y, z))
            for a, b in ((0, 1), (1, 2), (2, 3), (3, 0)):
                polys.append((base + a, base + 4 + a, base + 4 + b, base + b))
                polys.append((base + b, base + 4 + b, base + 4 + a, base + a))
            polys.append((base + 4, base + 7, base + 6, base + 5))
            polys.append((base + 4, base + 5, base + 6, base + 7))
        pts = np.array(pts, F)
    else:  # an octahedron of triangles
        pts = np.array([(0, 0.9, 0), (0.5, 0.45, 0), (0, 0.45, 0.5), (-0.5, 0.45, 0), (0, 0.45, -0.5), (0, 0.0, 0)], F)
        polys = []
        for a, b in ((1, 2), (2, 3), (3, 4), (4, 1)):
            polys += [(0, a, b, -1), (0, b, a, -1), (5, b, a, -1), (5, a, b, -1)]
    poly = np.zeros(len(polys), abi.POLYGON_DTYPE)
    for i, p in enumerate(polys):
        poly[i]["pointIndices"] = p if len(p) == 4 else (p[0], p[1], p[2], -1)
    colours = rng.integers(40, 256, (len(pts), 3)) / F(255.0)
    for i in range(len(poly)):
        for v in range(4):
            idx = poly[i]["pointIndices"][v]
            if idx >= 0:
                poly[i]["colors"][v] = (*colours[idx].astype(F), 1.0)
    spts, squads = _box((0.8, 0.9, 0.8), (0.0, 0.45, 0.0))
    spoly = np.zeros(len(squads) * 2, abi.POLYGON_DTYPE)
    for i, q in enumerate(squads):
        spoly[2 * i]["pointIndices"], spoly[2 * i + 1]["pointIndices"] = q, q[::-1]
    spoly["colors"] = 1.0
    return {"points": pts, "polygons": poly, "shadow_points": spts.astype(F), "shadow_polygons": spoly}


def build_assets(seed=21):
    rng = np.random.default_rng(seed)
    sprites = [
        make_sprite_type(rng, 64, 56, 1, (32, 40), (-0.5, 0.0, -0.5), (0.5, 0.25, 0.5)),                                      # floor tile
        make_sprite_type(rng, 40, 96, 4, (20, 84), (-0.25, 0.0, -0.25), (0.25, 1.5, 0.25), ((0.5, 1.5, 0.5), (0.0, 0.75, 0.0))),  # pillar
        make_sprite_type(rng, 30, 34, 8, (15, 28), (-0.25, 0.0, -0.25), (0.25, 0.5, 0.25), ((0.4, 0.5, 0.4), (0.0, 0.25, 0.0))),  # ball
    ]
    models = [make_model_type(rng, 0), make_model_type(rng, 1)]
    return {"sprites": sprites, "models": models}


def rotation_y(angle):
    c, s = F(np.cos(angle)), F(np.sin(angle))
    return ((c, 0, -s), (0, 1, 0), (s, 0, c))


def build_script(seed=22, width=320, height=240):
    """A list of (action, arguments). Sprite / model type numbers are local to the assets."""
    rng = np.random.default_rng(seed)
    s = []
    for gx in range(-5, 6):  # 121 floor tiles + objects: more than 64 leaves in one octree node, so nodes split and branch
        for gz in range(-5, 6):
            s.append(("bg_sprite", 0, 0, (gx * MINI, 0, gz * MINI), 0))
            r = rng.random()
            if r < 0.25:
                s.append(("bg_sprite", 1, int(rng.integers(0, 8)), (gx * MINI + int(rng.integers(-300, 300)), 0, gz * MINI + int(rng.integers(-300, 300))), 1))
            elif r < 0.45:
                s.append(("bg_sprite", 2, int(rng.integers(0, 8)), (gx * MINI + int(rng.integers(-400, 400)), int(rng.integers(0, 300)), gz * MINI + int(rng.integers(-400, 400))), int(rng.integers(0, 2))))
    s.append(("bg_sprite", 2, 3, (700, 100, -300), 1))
    s.append(("bg_sprite", 2, 3, (700, 100, -300), 1))  # an exact duplicate: equal heights, the first one drawn wins
    for k in range(4):
        pos = (F(rng.random() * 6 - 3), F(0.0), F(rng.random() * 6 - 3))
        s.append(("bg_model", k % 2, pos, rotation_y(rng.random() * 6.28)))
    lights = [("directed", (1.0, -1.0, 0.0), 0.1, (255, 255, 255)), ("directed", (-0.5, -1.0, 0.7), 0.05, (255, 200, 150)),
              ("point", (1.5, 1.2, 0.5), 4.0, 1.0, (255, 160, 90), 1), ("point", (-1.5, 0.9, -1.0), 3.0, 0.8, (90, 200, 255), 1), ("point", (0.0, 2.0, 2.0), 5.0, 0.6, (200, 255, 120), 0)]
    temps = lambda shift: [("tmp_sprite", 2, 1, (300 + shift, 200, 150), 1), ("tmp_sprite", 1, 6, (-900 - shift, 0, 700), 1), ("tmp_model", 1, (F(0.8), F(0.3), F(-0.6 + shift / 2048.0)), rotation_y(0.4))]
    s += lights + temps(0) + [("draw", width, height)]
    s += [("clear_temporary",)] + lights + temps(256) + [("draw", width, height)]                       # dirty-rectangle path
    s += [("move_camera", 40, -25), ("draw", width, height)]                                               # everything dirty
    s += [("remove_sprites", (-1500, -10, -1500), (600, 2000, 900)), ("bg_sprite", 1, 2, (100, 0, 200), 1), ("draw", width, height)]
    s += [("remove_models", (-4000, -10, -4000), (0, 2000, 4000)), ("draw", width, height)]
    s += [("camera_direction", 1), ("draw", width, height)]                                                # all blocks recycled
    s += [("clear_temporary",), ("point", (0.5, 1.0, 0.5), 4.0, 1.0, (255, 255, 255), 1), ("draw", 400, 300)]  # resize, no directed light
    s += [("camera_direction", 6), ("camera_location", (3 * MINI, 0, -2 * MINI)), ("draw", 400, 300)]
    s += [("camera_location", (40 * MINI, 0, 40 * MINI)), ("draw", 400, 300), ("camera_location", (0, 0, 0)), ("draw", 400, 300)]  # far away and back: recycle by distance
    return s


def sprite_instance(type_index, direction, location, shadow):
    inst = abi.SpriteInstance()
    inst.typeIndex, inst.direction, inst.shadowCasting, inst.userData = type_index, direction, shadow, 0
    inst.location[:] = [int(v) for v in location]
    return inst


def model_instance(type_index, position, axes):
    inst = abi.ModelInstance()
    inst.typeIndex, inst.userData = type_index, 0
    inst.location = abi.Transform3D.make(position, axes)
    return inst


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f3(values):
    return np.array(values, F)


def _i3(values):
    return np.array(values, np.int32)


# ------------------------------------------------------------------------------------------------ reference back end

def run_reference(ref, assets, script, folder, shadow_res=SHADOW_RES, keep_frames=True, timing=None, sprite_ids=None):
    """Returns one dict of buffers per draw. sprite_ids: sprite types the reference has loaded already (real media), else they are made from the assets."""
    lib = ref.lib
    model_ids = []
    if sprite_ids is None:
        sprite_ids = []
        for k, t in enumerate(assets["sprites"]):
            atlas = ref.rgba(t["atlas"])
            sprite_ids.append(lib.ref_sprite_type_create(atlas, sprite_ini(t).encode(), folder.encode(), f"sprite{k}_{lib.ref_sprite_type_count()}".encode()))
    for m in assets["models"]:
        dense = lib.ref_dense_model_create(ref.model(m["points"], m["polygons"]))
        model_ids.append(lib.ref_model_type_create(dense, ref.model(m["shadow_points"], m["shadow_polygons"])))
    world = lib.ref_world_create(TILT, PIXELS_PER_TILE, shadow_res)
    frames = []
    for action in script:
        kind = action[0]
        if kind == "bg_sprite":
            lib.ref_world_add_background_sprite(world, C.byref(sprite_instance(sprite_ids[action[1]], action[2], action[3], action[4])))
        elif kind == "tmp_sprite":
            lib.ref_world_add_temporary_sprite(world, C.byref(sprite_instance(sprite_ids[action[1]], action[2], action[3], action[4])))
        elif kind == "bg_model":
            lib.ref_world_add_background_model(world, C.byref(model_instance(model_ids[action[1]], action[2], action[3])))
        elif kind == "tmp_model":
            lib.ref_world_add_temporary_model(world, C.byref(model_instance(model_ids[action[1]], action[2], action[3])))
        elif kind == "directed":
            lib.ref_world_directed_light(world, _ptr(_f3(action[1])), action[2], _ptr(_i3(action[3])))
        elif kind == "point":
            lib.ref_world_point_light(world, _ptr(_f3(action[1])), action[2], action[3], _ptr(_i3(action[4])), action[5])
        elif kind == "clear_temporary":
            lib.ref_world_clear_temporary(world)
        elif kind == "move_camera":
            lib.ref_world_move_camera_in_pixels(world, action[1], action[2])
        elif kind == "camera_direction":
            lib.ref_world_set_camera_direction_index(world, action[1])
        elif kind == "camera_location":
            lib.ref_world_set_camera_location(world, _ptr(_i3(action[1])))
        elif kind == "remove_sprites":
            lib.ref_world_remove_background_sprites(world, _ptr(_i3(action[1])), _ptr(_i3(action[2])))
        elif kind == "remove_models":
            lib.ref_world_remove_background_models(world, _ptr(_i3(action[1])), _ptr(_i3(action[2])))
        elif kind == "draw":
            w, h = action[1], action[2]
            colour = ref.rgba(shape=(h, w))
            t0 = lib.ref_time_seconds()
            lib.ref_world_draw(world, colour)
            if timing is not None:
                timing.append(lib.ref_time_seconds() - t0)
            if not keep_frames:
                continue
            d, n, l, hgt = np.zeros((h, w), np.uint32), np.zeros((h, w), np.uint32), np.zeros((h, w), np.uint32), np.zeros((h, w), F)
            lib.ref_world_read_buffers(world, _ptr(d), _ptr(n), _ptr(l), _ptr(hgt))
            location = np.zeros(3, np.int32)
            lib.ref_world_get_camera_location(world, _ptr(location))
            ground = np.zeros(3, np.int32)
            lib.ref_world_find_ground_at_pixel(world, colour, 17, 200, _ptr(ground))
            frames.append({"color": ref.read_rgba(colour), "diffuse": d, "normal": n, "light": l, "height": hgt, "camera": location.copy(), "ground": ground.copy()})
    return frames


# ------------------------------------------------------------------------------------------------ product world (host side), shared by the two product back ends

class ProductWorld:
    """Creates the types and the world through the C ABI (host only) and applies the non-draw actions of the script."""

    def __init__(self, lib_handle, check, assets, shadow_res=SHADOW_RES):
        self.h, self.check = lib_handle, check
        self.assets = assets
        self.sprite_ids, self.model_ids, self.dense = [], [], []
        h = self.h
        for t in assets["sprites"]:
            cfg = abi.SpriteConfig()
            cfg.centerX, cfg.centerY, cfg.frameRows, cfg.propertyColumns = t["center"][0], t["center"][1], t["frames"], 3
            cfg.minBound[:] = [float(v) for v in t["min"]]
            cfg.maxBound[:] = [float(v) for v in t["max"]]
            if t["points"] is not None:
                cfg.points, cfg.pointCount = t["points"].ctypes.data, len(t["points"])
                cfg.triangleIndices, cfg.triangleIndexCount = t["indices"].ctypes.data, len(t["indices"])
            index = C.c_int32()
            atlas = np.ascontiguousarray(t["atlas"])
            check(h.dfpsr_sprite_type_create(_ptr(atlas), atlas.shape[1], atlas.shape[0], atlas.shape[1] * 4, C.byref(cfg), C.byref(index)))
            self.sprite_ids.append(index.value)
        for m in assets["models"]:
            tris, mn, mx = dense_build(h, check, m["points"], m["polygons"])
            shadow = abi.HostModel()
            shadow.points, shadow.pointCount = m["shadow_points"].ctypes.data, len(m["shadow_points"])
            shadow.polygons, shadow.polygonCount = m["shadow_polygons"].ctypes.data, len(m["shadow_polygons"])
            index = C.c_int32()
            check(h.dfpsr_model_type_create(_ptr(tris), len(tris), _ptr(mn), _ptr(mx), C.byref(shadow), C.byref(index)))
            self.model_ids.append(index.value)
            self.dense.append((tris, mn, mx))
        self.ortho = abi.OrthoSystem()
        check(h.dfpsr_ortho_system_create(C.byref(self.ortho), TILT, PIXELS_PER_TILE))
        self.world = C.c_void_p()
        check(h.dfpsr_sprite_world_create(C.byref(self.world), C.byref(self.ortho), shadow_res))
        self.directed, self.points = [], []

    def close(self):
        self.check(self.h.dfpsr_sprite_world_destroy(self.world))

    def apply(self, action):
        h, check, world = self.h, self.check, self.world
        kind = action[0]
        if kind == "bg_sprite":
            check(h.dfpsr_sprite_world_add_background_sprite(world, C.byref(sprite_instance(self.sprite_ids[action[1]], action[2], action[3], action[4]))))
        elif kind == "tmp_sprite":
            check(h.dfpsr_sprite_world_add_temporary_sprite(world, C.byref(sprite_instance(self.sprite_ids[action[1]], action[2], action[3], action[4]))))
        elif kind == "bg_model":
            check(h.dfpsr_sprite_world_add_background_model(world, C.byref(model_instance(self.model_ids[action[1]], action[2], action[3]))))
        elif kind == "tmp_model":
            check(h.dfpsr_sprite_world_add_temporary_model(world, C.byref(model_instance(self.model_ids[action[1]], action[2], action[3]))))
        elif kind == "directed":
            check(h.dfpsr_sprite_world_create_temporary_directed_light(world, _ptr(_f3(action[1])), action[2], _ptr(_i3(action[3]))))
            self.directed.append(action)
        elif kind == "point":
            check(h.dfpsr_sprite_world_create_temporary_point_light(world, _ptr(_f3(action[1])), action[2], action[3], _ptr(_i3(action[4])), action[5]))
            self.points.append(action)
        elif kind == "clear_temporary":
            check(h.dfpsr_sprite_world_clear_temporary(world))
            self.directed, self.points = [], []
        elif kind == "move_camera":
            check(h.dfpsr_sprite_world_move_camera_in_pixels(world, action[1], action[2]))
        elif kind == "camera_direction":
            check(h.dfpsr_sprite_world_set_camera_direction_index(world, action[1]))
        elif kind == "camera_location":
            check(h.dfpsr_sprite_world_set_camera_location(world, _ptr(_i3(action[1]))))
        elif kind == "remove_sprites":
            check(h.dfpsr_sprite_world_remove_background_sprites(world, _ptr(_i3(action[1])), _ptr(_i3(action[2])), None, None))
        elif kind == "remove_models":
            check(h.dfpsr_sprite_world_remove_background_models(world, _ptr(_i3(action[1])), _ptr(_i3(action[2])), None, None))
        else:
            raise ValueError(kind)

    def camera_state(self, width, height):
        location, ground, index = np.zeros(3, np.int32), np.zeros(3, np.int32), C.c_int32()
        self.check(self.h.dfpsr_sprite_world_get_camera_location(self.world, _ptr(location)))
        self.check(self.h.dfpsr_sprite_world_find_ground_at_pixel(self.world, width, height, 17, 200, _ptr(ground)))
        self.check(self.h.dfpsr_sprite_world_get_camera_direction_index(self.world, C.byref(index)))
        return location, ground, index.value


def dense_build(h, check, points, polygons):
    pts, poly = np.ascontiguousarray(points, F), np.ascontiguousarray(polygons)
    count = h.dfpsr_dense_model_triangle_count(_ptr(poly), len(poly))
    tris = np.zeros(count, abi.DENSE_TRIANGLE_DTYPE)
    mn, mx = np.zeros(3, F), np.zeros(3, F)
    check(h.dfpsr_dense_model_build(_ptr(pts), len(pts), _ptr(poly), len(poly), _ptr(tris), _ptr(mn), _ptr(mx)))
    return tris, mn, mx


# ------------------------------------------------------------------------------------------------ plan + oracle back end (CPU only)

CUBE_SIDES = [((1, 0, 0), (0, 1, 0)), ((-1, 0, 0), (0, 1, 0)), ((0, 1, 0), (0, 0, 1)), ((0, -1, 0), (0, 0, 1)), ((0, 0, 1), (0, 1, 0)), ((0, 0, -1), (0, 1, 0))]


def _mat(m):
    return np.array([list(m.xAxis), list(m.yAxis), list(m.zAxis)], F)


def _mat_transform(m, p):
    p = np.asarray(p, F)
    return (F(p[0]) * m[0] + F(p[1]) * m[1]) + F(p[2]) * m[2]


class OracleExecutor:
    """Replays dfpsr_sprite_world_op lists with the C oracle on numpy buffers."""

    def __init__(self, oracle, product_world):
        import orcbind
        self.o, self.ob, self.pw = oracle, orcbind, product_world
        self.blocks = {}
        self.size = None
        self.frames = {}
        assets = product_world.assets
        self.sprite_frames = {}
        for local, t in enumerate(assets["sprites"]):
            for f in range(t["frames"]):
                rows = slice(f * t["frame_h"], (f + 1) * t["frame_h"])
                colour, hcol, normal = (np.ascontiguousarray(t["atlas"][rows, c * t["frame_w"]:(c + 1) * t["frame_w"]]) for c in range(3))
                height = np.zeros(colour.shape, F)
                IM = orcbind.image_of
                oracle.orc_sprite_scale_height(C.byref(IM(hcol)), C.byref(IM(colour)), float(t["min"][1]), float(t["max"][1]), C.byref(IM(height)))
                self.sprite_frames[(product_world.sprite_ids[local], f)] = (height, colour, normal)
        self.sprite_shadow = {}
        for local, t in enumerate(assets["sprites"]):
            if t["points"] is not None:
                poly = np.zeros(len(t["indices"]) // 3, abi.POLYGON_DTYPE)
                poly["pointIndices"][:, :3] = t["indices"].reshape(-1, 3)
                poly["pointIndices"][:, 3] = -1
                poly["colors"] = 1.0
                self.sprite_shadow[product_world.sprite_ids[local]] = orcbind.model_of(t["points"], poly)
        self.model_shadow = {product_world.model_ids[k]: orcbind.model_of(m["shadow_points"], m["shadow_polygons"]) for k, m in enumerate(assets["models"])}
        self.dense = {product_world.model_ids[k]: product_world.dense[k] for k in range(len(assets["models"]))}

    def _targets(self, block):
        return self.blocks[block] if block >= 0 else (self.H, self.D, self.N)

    def frame(self, ops, count, width, height, camera_index):
        o, IM = self.o, self.ob.image_of
        if self.size != (width, height):
            self.size = (width, height)
            self.D, self.N, self.L = (np.zeros((height, width), np.uint32) for _ in range(3))
            self.H = np.zeros((height, width), F)
        colour = np.zeros((height, width), np.uint32)
        view = self.pw.ortho.view[camera_index]
        light_view = view.light_view()
        pw = self.pw
        centre = np.zeros(3, np.int32)  # worldCenter = (w/2, h/2) - cameraPixel: recomputed like the library does
        cube, cams = None, None
        res = SHADOW_RES
        for k in range(count):
            op = ops[k]
            kind = op.op
            if kind == abi.SW_BLOCK_CLEAR:
                self.blocks[op.block] = (np.full((512, 512), -1000000.0, F), np.zeros((512, 512), np.uint32), np.full((512, 512), 0x80808080, np.uint32))
            elif kind in (abi.SW_BLOCK_SPRITE, abi.SW_SPRITE):
                th, td, tn = self._targets(op.block if kind == abi.SW_BLOCK_SPRITE else -1)
                sh, sd, sn = self.sprite_frames[(op.typeIndex, op.frame)]
                o.orc_draw_higher(C.byref(IM(th)), C.byref(IM(sh)), C.byref(IM(td)), C.byref(IM(sd)), C.byref(IM(tn)), C.byref(IM(sn)), op.left, op.top, op.heightOffset)
            elif kind in (abi.SW_BLOCK_MODEL, abi.SW_MODEL):
                th, td, tn = self._targets(op.block if kind == abi.SW_BLOCK_MODEL else -1)
                tris, mn, mx = self.dense[op.typeIndex]
                origin = np.array(list(op.worldOrigin), F)
                o.orc_dense_model_render(_ptr(tris), len(tris), _ptr(mn), _ptr(mx), C.byref(view), C.byref(IM(th)), C.byref(IM(td)), C.byref(IM(tn)), _ptr(origin), C.byref(op.transform), 0, None)
            elif kind == abi.SW_COPY_BLOCK:
                bh, bd, bn = self.blocks[op.block]
                dst = (slice(op.top, op.top + op.height), slice(op.left, op.left + op.width))
                src = (slice(op.sourceTop, op.sourceTop + op.height), slice(op.sourceLeft, op.sourceLeft + op.width))
                self.H[dst], self.D[dst], self.N[dst] = bh[src], bd[src], bn[src]
            elif kind == abi.SW_LIGHT_CLEAR:
                self.L[:] = 0
            elif kind == abi.SW_LIGHT_DIRECTED:
                a = pw.directed[op.light]
                o.orc_light_directed(C.byref(light_view), C.byref(IM(self.L)), C.byref(IM(self.N)), _ptr(_f3(a[1])), a[2], _ptr(_i3(a[3])), 0 if op.flag else 1)
            elif kind == abi.SW_SHADOW_CLEAR:
                cube = np.zeros((res * 6, res), F)
                if cams is None:
                    n2w = _mat(view.normalToWorldSpace)
                    cams = []
                    for forward, up in CUBE_SIDES:
                        side = np.stack(scenes.make_axis_system(forward, up)).astype(F)
                        rot = np.stack([_mat_transform(n2w, side[0]), _mat_transform(n2w, side[1]), _mat_transform(n2w, side[2])]).astype(F)
                        cams.append(self.ob.camera(abi.camera_params(True, abi.Transform3D.make((0, 0, 0), rot), res, res)))
            elif kind in (abi.SW_SHADOW_SPRITE, abi.SW_SHADOW_MODEL):
                model, _keep = (self.sprite_shadow if kind == abi.SW_SHADOW_SPRITE else self.model_shadow)[op.typeIndex]
                for s in range(6):
                    o.orc_model_render_depth(C.byref(model), C.byref(op.transform), C.byref(IM(cube[s * res:(s + 1) * res])), C.byref(cams[s]))
            elif kind == abi.SW_LIGHT_POINT:
                a = pw.points[op.light]
                location, _ground, _index = pw.camera_state(width, height)
                cam_pixel = mini_offset_to_pixel(view, location)
                wc = np.array([width // 2 - cam_pixel[0], height // 2 - cam_pixel[1]], np.int32)
                shadow = C.byref(IM(cube)) if op.flag else None
                o.orc_light_point(C.byref(light_view), _ptr(wc), C.byref(IM(self.L)), C.byref(IM(self.N)), C.byref(IM(self.H)), _ptr(_f3(a[1])), a[2], a[3], _ptr(_i3(a[4])), shadow, 4)
            elif kind == abi.SW_BLEND:
                o.orc_light_blend(C.byref(IM(colour)), C.byref(IM(self.D)), C.byref(IM(self.L)))
            else:
                raise ValueError(kind)
        return {"color": colour, "diffuse": self.D.copy(), "normal": self.N.copy(), "light": self.L.copy(), "height": self.H.copy()}


def _trunc_div(a, b):
    q = abs(int(a)) // b
    return q if a >= 0 else -q


def mini_offset_to_pixel(view, offset):
    """ref: orthoAPI.cpp:34-38 (C++ integer division truncates towards zero)."""
    x = view.pixelOffsetPerTileX[0] * int(offset[0]) + view.pixelOffsetPerTileZ[0] * int(offset[2])
    y = view.pixelOffsetPerTileX[1] * int(offset[0]) + view.pixelOffsetPerTileZ[1] * int(offset[2]) - int(offset[1]) * view.yPixelsPerTile
    return _trunc_div(x, MINI), _trunc_div(y, MINI)


def run_plan_oracle(lib_handle, check, oracle, assets, script):
    pw = ProductWorld(lib_handle, check, assets)
    ex = OracleExecutor(oracle, pw)
    frames = []
    for action in script:
        if action[0] == "draw":
            w, h = action[1], action[2]
            ops, count = C.POINTER(abi.SpriteWorldOp)(), C.c_int32()
            location, ground, index = pw.camera_state(w, h)
            check(lib_handle.dfpsr_sprite_world_plan_frame(pw.world, w, h, C.byref(ops), C.byref(count)))
            result = ex.frame(ops, count.value, w, h, index)
            result["camera"], result["ground"] = location, ground
            result["ops"] = count.value
            frames.append(result)
        else:
            pw.apply(action)
    pw.close()
    return frames


# ------------------------------------------------------------------------------------------------ CUDA back end

def run_cuda(lib_handle, lib, assets, script):
    import torch
    pw = ProductWorld(lib_handle, lib.check, assets)
    frames = []
    for action in script:
        if action[0] == "draw":
            w, h = action[1], action[2]
            colour = torch.zeros((h, w), dtype=torch.int32, device="cuda")
            location, ground, _index = pw.camera_state(w, h)
            lib.check(lib_handle.dfpsr_sprite_world_draw(pw.world, C.byref(lib.image(colour)), lib.stream_ptr()))
            images = [abi.Image() for _ in range(4)]
            lib.check(lib_handle.dfpsr_sprite_world_get_buffers(pw.world, *[C.byref(im) for im in images]))
            torch.cuda.synchronize()
            out = {"color": colour.cpu().numpy().view(np.uint32), "camera": location, "ground": ground}
            for name, im, dtype in zip(("diffuse", "normal", "light", "height"), images, (np.uint32, np.uint32, np.uint32, F)):
                host = np.zeros((im.height, im.stride // 4), dtype)
                lib.check(lib_handle.dfpsr_download(_ptr(host), im.data, host.nbytes, lib.stream_ptr()))
                torch.cuda.synchronize()
                out[name] = host[:, :im.width].copy()
            frames.append(out)
        else:
            pw.apply(action)
    pw.close()
    return frames


def frame_hashes(frame):
    return {k: sha(frame[k]) for k in ("color", "diffuse", "normal", "light", "height")}


# ------------------------------------------------------------------------------------------------ renderDenseModel alone

# (model, view index, high quality, width, height, world origin, position, rotation angle about Y, uniform scale)
DENSE_CASES = [
    (0, 0, 1, 200, 160, (100.0, 110.0), (0.0, 0.0, 0.0), 0.0, 1.0),
    (0, 3, 0, 200, 160, (90.5, 100.25), (0.3, 0.1, -0.2), 0.7, 1.3),
    (1, 5, 1, 160, 200, (80.0, 150.0), (-0.2, 0.5, 0.4), 2.1, 1.7),
    (1, 2, 0, 97, 61, (10.0, 40.0), (0.0, 0.0, 0.0), 4.0, 2.5),      # partly outside of the target
    (0, 7, 1, 64, 64, (400.0, 400.0), (0.0, 0.0, 0.0), 0.0, 1.0),    # culled: nothing drawn, empty dirty rectangle
]


def dense_case(case):
    model, view_index, hq, w, h, origin, position, angle, scale = DENSE_CASES[case]
    axes = [[F(v) * F(scale) for v in axis] for axis in rotation_y(angle)]
    rng = np.random.default_rng(100 + case)
    height = (rng.random((h, w)) * 0.2 - 40.0).astype(F)  # a background below the model with some noise
    return model, view_index, hq, w, h, np.array(origin, F), abi.Transform3D.make(position, axes), height


def dense_reference(ref, assets, case):
    model, view_index, hq, w, h, origin, transform, height = dense_case(case)
    m = assets["models"][model]
    dense = ref.lib.ref_dense_model_create(ref.model(m["points"], m["polygons"]))
    H, D, N = ref.f32(height), ref.rgba(shape=(h, w)), ref.rgba(shape=(h, w))
    rect = np.zeros(4, np.int32)
    ref.lib.ref_dense_model_render(dense, TILT, PIXELS_PER_TILE, view_index, H, D, N, float(origin[0]), float(origin[1]), C.byref(transform), hq, _ptr(rect))
    return {"height": ref.read_f32(H), "diffuse": ref.read_rgba(D), "normal": ref.read_rgba(N), "rect": rect}


def dense_oracle(oracle, lib_handle, check, assets, case):
    import orcbind
    IM = orcbind.image_of
    model, view_index, hq, w, h, origin, transform, height = dense_case(case)
    m = assets["models"][model]
    tris, mn, mx = dense_build(lib_handle, check, m["points"], m["polygons"])
    ortho = abi.OrthoSystem()
    check(lib_handle.dfpsr_ortho_system_create(C.byref(ortho), TILT, PIXELS_PER_TILE))
    H, D, N = height.copy(), np.zeros((h, w), np.uint32), np.zeros((h, w), np.uint32)
    rect = np.zeros(4, np.int32)
    oracle.orc_dense_model_render(_ptr(tris), len(tris), _ptr(mn), _ptr(mx), C.byref(ortho.view[view_index]), C.byref(IM(H)), C.byref(IM(D)), C.byref(IM(N)), _ptr(origin), C.byref(transform), hq, _ptr(rect))
    return {"height": H, "diffuse": D, "normal": N, "rect": rect}


def dense_cuda(lib_handle, lib, assets, case):
    import torch
    model, view_index, hq, w, h, origin, transform, height = dense_case(case)
    m = assets["models"][model]
    tris, mn, mx = dense_build(lib_handle, lib.check, m["points"], m["polygons"])
    ortho = abi.OrthoSystem()
    lib.check(lib_handle.dfpsr_ortho_system_create(C.byref(ortho), TILT, PIXELS_PER_TILE))
    d_tris = lib.to_device(tris.view(np.uint8).reshape(-1))
    H, D, N = lib.to_device(height), torch.zeros((h, w), dtype=torch.int32, device="cuda"), torch.zeros((h, w), dtype=torch.int32, device="cuda")
    rect = np.zeros(4, np.int32)
    lib.check(lib_handle.dfpsr_dense_model_render(d_tris.data_ptr(), len(tris), _ptr(mn), _ptr(mx), C.byref(ortho.view[view_index]), C.byref(lib.image(H)), C.byref(lib.image(D)), C.byref(lib.image(N)),
                                                  _ptr(origin), C.byref(transform), hq, _ptr(rect), lib.stream_ptr()))
    torch.cuda.synchronize()
    return {"height": H.cpu().numpy(), "diffuse": D.cpu().numpy().view(np.uint32), "normal": N.cpu().numpy().view(np.uint32), "rect": rect}


def frame_hashes_dense(result):
    return {"height": sha(result["height"]), "diffuse": sha(result["diffuse"]), "normal": sha(result["normal"]), "rect": [int(v) for v in result["rect"]],
            "touched": float((result["diffuse"] != 0).mean())}


# ------------------------------------------------------------------------------------------------ BASELINE config 2 through the sprite world

def sandbox_script(width=800, height=600, lights=16, frames=6, seed=31):
    """An 800x600 Sandbox session in the spirit of SDK/sandbox/sandbox.cpp:336-353, :366-493: a tiled floor with objects, one directed light,
    `lights` shadow-casting point lights on a grid and two moving temporary sprites; the camera pans every other frame."""
    rng = np.random.default_rng(seed)
    s = []
    for gx in range(-12, 13):
        for gz in range(-12, 13):
            s.append(("bg_sprite", 0, 0, (gx * MINI, 0, gz * MINI), 0))
            r = rng.random()
            if r < 0.2:
                s.append(("bg_sprite", 1, int(rng.integers(0, 8)), (gx * MINI + int(rng.integers(-300, 300)), 0, gz * MINI + int(rng.integers(-300, 300))), 1))
            elif r < 0.35:
                s.append(("bg_sprite", 2, int(rng.integers(0, 8)), (gx * MINI + int(rng.integers(-400, 400)), 0, gz * MINI + int(rng.integers(-400, 400))), 1))
    for k in range(6):
        s.append(("bg_model", k % 2, (F(rng.random() * 10 - 5), F(0.0), F(rng.random() * 10 - 5)), rotation_y(rng.random() * 6.28)))
    grid = int(np.ceil(np.sqrt(lights)))
    for frame in range(frames):
        s.append(("clear_temporary",))
        s.append(("directed", (1.0, -1.0, 0.0), 0.1, (255, 255, 255)))
        for i in range(lights):
            gx, gz = i % grid, i // grid
            colour = (int(90 + 40 * (i % 4)), int(255 - 30 * (i % 5)), int(120 + 25 * (i % 6)))
            s.append(("point", (F(-4.5 + 9.0 * (gx + 0.5) / grid), F(1.0 + 0.1 * (i % 3)), F(-4.5 + 9.0 * (gz + 0.5) / grid)), 4.0, 1.0, colour, 1))
        s.append(("tmp_sprite", 2, frame % 8, (500 + 150 * frame, 200, -300), 1))
        s.append(("tmp_sprite", 1, (frame + 3) % 8, (-1200, 0, 900 - 100 * frame), 1))
        if frame % 2 == 1:
            s.append(("move_camera", 8, -5))
        s.append(("draw", width, height))
    return s


# ------------------------------------------------------------------------------------------------ sprite baking (sprite_generateFromModel)

BAKE_CASES = [(0, 8), (1, 3), (0, 1)]  # (model, camera angles)


def bake_reference(ref, assets, case):
    model, angles = BAKE_CASES[case]
    m = assets["models"][model]
    info = np.zeros(4, np.int32)
    atlas = ref.lib.ref_sprite_generate_from_model(ref.model(m["points"], m["polygons"]), -1, TILT, PIXELS_PER_TILE, angles, _ptr(info))
    assert atlas >= 0
    return {"atlas": ref.read_rgba(atlas), "config": [int(v) for v in info]}


def bake_cuda(lib_handle, lib, assets, case):
    import torch
    model, angles = BAKE_CASES[case]
    m = assets["models"][model]
    tris, mn, mx = dense_build(lib_handle, lib.check, m["points"], m["polygons"])
    ortho = abi.OrthoSystem()
    lib.check(lib_handle.dfpsr_ortho_system_create(C.byref(ortho), TILT, PIXELS_PER_TILE))
    baked = abi.BakedSprite()
    lib.check(lib_handle.dfpsr_sprite_generate_from_model(_ptr(tris), len(tris), _ptr(mn), _ptr(mx), C.byref(ortho), angles, C.byref(baked), lib.stream_ptr()))
    assert baked.atlas.data
    host = np.zeros((baked.atlas.height, baked.atlas.stride // 4), np.uint32)
    lib.check(lib_handle.dfpsr_download(_ptr(host), baked.atlas.data, host.nbytes, lib.stream_ptr()))
    torch.cuda.synchronize()
    lib.check(lib_handle.dfpsr_free(baked.atlas.data))
    assert list(baked.minBound) == [float(v) for v in mn] and list(baked.maxBound) == [float(v) for v in mx]
    return {"atlas": host[:, :baked.atlas.width].copy(), "config": [baked.centerX, baked.centerY, baked.frameRows, baked.propertyColumns]}
