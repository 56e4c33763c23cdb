"""Synthetic scenes of the shapes BASELINE.json names (no reference assets are read).

All generators are deterministic functions of their seed and return plain numpy arrays in the layouts of
include/dfpsr_b200.h, so the same scene can be given to the CUDA path, the C oracle and the compiled
reference. Shapes follow the SDK programs that define the benchmark configs:
  terrain  — ref: SDK/terrain/main.cpp:118-178 (64x64 height map -> 4096 points, one quad per non-sea cell,
             alternating diagonal, uv = cell / (size - 1)), 1024x1024 colour map with 5 mip levels (:370).
  orbit    — ref: SDK/terrain/main.cpp:405-413 camera orbit.
  tiny     — SURVEY.md §8(d) config 3: 1000 x 999 quad grid of vertex-coloured triangles seen from above.
"""
import math

import numpy as np

from . import abi

F = np.float32


def _v(x):
    return np.asarray(x, dtype=F)


def normalize(v):
    v = _v(v)
    l = F(np.sqrt(F(F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2]))))
    if l == 0:
        return _v([0, 0, 1])
    return _v(v / l)


def cross(a, b):
    a, b = _v(a), _v(b)
    return _v([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def make_axis_system(forward, up=(0.0, 1.0, 0.0)):
    """ref: math/FMatrix3x3.h:44-51 makeAxisSystem (inputs to the renderer, so only determinism matters)."""
    z = normalize(forward)
    x = normalize(cross(normalize(up), z))
    y = normalize(cross(z, x))
    return x, y, z


def look_at_transform(position, target):
    position, target = _v(position), _v(target)
    axes = make_axis_system(target - position)
    return abi.Transform3D.make(position, axes)


def orbit_camera(frame, width, height, frames_per_lap=60, map_size=64, distance=10.0, lift=10.0):
    t = 2.0 * math.pi * frame / frames_per_lap
    center = _v([map_size * 0.5, 0.0, map_size * -0.5])
    offset = _v([math.sin(t) * distance, lift, math.cos(t) * distance])
    axes = make_axis_system(-offset)
    location = abi.Transform3D.make(center + offset, axes)
    return abi.camera_params(True, location, width, height)


def _value_noise(size, cells, rng):
    grid = rng.random((cells + 1, cells + 1))
    xs = np.linspace(0, cells, size, endpoint=False)
    x0 = np.floor(xs).astype(int)
    fx = xs - x0
    fx = fx * fx * (3 - 2 * fx)
    a = grid[np.ix_(x0, x0)]
    b = grid[np.ix_(x0, x0 + 1)]
    c = grid[np.ix_(x0 + 1, x0)]
    d = grid[np.ix_(x0 + 1, x0 + 1)]
    top = a * (1 - fx)[None, :] + b * fx[None, :]
    bottom = c * (1 - fx)[None, :] + d * fx[None, :]
    return top * (1 - fx)[:, None] + bottom * fx[:, None]


def height_map(size=64, seed=7):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size]
    r = np.hypot((xx - size / 2 + 0.5) / (size / 2), (yy - size / 2 + 0.5) / (size / 2))
    island = np.clip(1.6 - r * r * 1.0, 0.0, 1.0) * np.clip(1.25 - r * 0.75, 0.0, 1.0)
    noise = 0.6 * _value_noise(size, 4, rng) + 0.3 * _value_noise(size, 8, rng) + 0.1 * _value_noise(size, 16, rng)
    h = np.clip(island * (0.25 + noise) * 1.05 - 0.06, 0.0, 1.0)
    return np.round(h * 255).astype(np.uint8)


def pack_rgba(r, g, b, a):
    return (np.asarray(r, np.uint32) | (np.asarray(g, np.uint32) << 8) | (np.asarray(b, np.uint32) << 16) | (np.asarray(a, np.uint32) << 24)).astype(np.uint32)


def colour_map(heights, density=16, seed=11):
    """A 1024x1024 lit colour map in the spirit of generateDiffuseMap/updateColorMap (main.cpp:271-307)."""
    size = heights.shape[0] * density
    rng = np.random.default_rng(seed)
    h = np.kron(heights.astype(np.float64), np.ones((density, density)))
    # cheap blur so slopes are smooth
    for _ in range(2):
        h = (h + np.roll(h, 1, 0) + np.roll(h, -1, 0) + np.roll(h, 1, 1) + np.roll(h, -1, 1)) / 5.0
    bump = h + (_value_noise(size, 64, rng) - 0.5) * 18.0
    gx = np.roll(bump, -1, 1) - np.roll(bump, 1, 1)
    gy = np.roll(bump, -1, 0) - np.roll(bump, 1, 0)
    light = np.clip(0.75 - 0.04 * gx - 0.06 * gy, 0.0, 1.6) + 0.2
    t = np.clip(bump / 255.0, 0, 1)
    ramp_r = np.interp(t, [0, 0.05, 0.15, 0.5, 0.8, 1.0], [40, 210, 70, 60, 120, 250])
    ramp_g = np.interp(t, [0, 0.05, 0.15, 0.5, 0.8, 1.0], [70, 200, 150, 110, 110, 250])
    ramp_b = np.interp(t, [0, 0.05, 0.15, 0.5, 0.8, 1.0], [160, 140, 50, 40, 100, 250])
    r = np.clip(ramp_r * light, 0, 255).astype(np.uint32)
    g = np.clip(ramp_g * light, 0, 255).astype(np.uint32)
    b = np.clip(ramp_b * light, 0, 255).astype(np.uint32)
    return pack_rgba(r, g, b, np.full_like(r, 255))


def terrain_scene(map_size=64, density=16, seed=7, highest_ground=5.0):
    heights = height_map(map_size, seed)
    per_unit = F(highest_ground) / F(255.0)
    points = np.zeros((map_size * map_size, 3), F)
    polygons = []
    scale = F(1.0) / F(map_size - 1.0)
    for z in range(map_size):
        for x in range(map_size):
            points[x + z * map_size] = (F(x), F(heights[z, x]) * per_unit, F(-z))
            if x > 0 and z > 0:
                px, pz = x - 1, z - 1
                if heights[pz, px] > 0 or heights[pz, x] > 0 or heights[z, px] > 0 or heights[z, x] > 0:
                    ia, ib, ic, id_ = px + pz * map_size, x + pz * map_size, x + z * map_size, px + z * map_size
                    ta = (F(px) * scale, F(pz) * scale, 0, 0)
                    tb = (F(x) * scale, F(pz) * scale, 0, 0)
                    tc = (F(x) * scale, F(z) * scale, 0, 0)
                    td = (F(px) * scale, F(z) * scale, 0, 0)
                    if (x + z) % 2 == 0:
                        polygons.append(((ia, ib, ic, id_), (ta, tb, tc, td)))
                    else:
                        polygons.append(((ib, ic, id_, ia), (tb, tc, td, ta)))
    poly = np.zeros(len(polygons), abi.POLYGON_DTYPE)
    for i, (idx, tex) in enumerate(polygons):
        poly[i]["pointIndices"] = idx
        poly[i]["texCoords"] = tex
    poly["colors"] = 1.0
    texture = colour_map(heights, density, seed + 4)
    return {"points": points, "polygons": poly, "texture": texture, "texture_levels": 5, "filter": abi.FILTER_SOLID}


def tiny_triangle_scene(nx=1000, nz=999, seed=3):
    """(nx+1) x (nz+1) points, nx*nz quads = 2*nx*nz vertex-coloured triangles, no textures."""
    rng = np.random.default_rng(seed)
    px, pz = nx + 1, nz + 1
    xs, zs = np.meshgrid(np.arange(px, dtype=F), np.arange(pz, dtype=F))
    ys = (rng.random((pz, px)) * 0.3).astype(F)
    points = np.stack([xs, ys, -zs], axis=-1).reshape(-1, 3).astype(F)
    vertex_rgb = rng.random((pz * px, 3)).astype(F)
    gx, gz = np.meshgrid(np.arange(nx), np.arange(nz))
    a = (gx + gz * px).reshape(-1)
    b = a + 1
    c = a + 1 + px
    d = a + px
    poly = np.zeros(nx * nz, abi.POLYGON_DTYPE)
    idx = np.stack([a, b, c, d], axis=1).astype(np.int32)
    poly["pointIndices"] = idx
    poly["colors"][:, :, :3] = vertex_rgb[idx]
    poly["colors"][:, :, 3] = 1.0
    return {"points": points, "polygons": poly, "texture": None, "texture_levels": 0, "filter": abi.FILTER_SOLID}


def top_down_camera(nx, nz, width, height):
    """Perspective camera above the centre of the tiny-triangle grid, looking straight down."""
    position = _v([nx * 0.5, nx * 0.5, -nz * 0.5])
    axes = make_axis_system((0.0, -1.0, 0.0), (0.0, 0.0, 1.0))
    return abi.camera_params(True, abi.Transform3D.make(position, axes), width, height)


def checker_texture(size=64, seed=5):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size]
    base = (((xx // 4) + (yy // 4)) % 2) * 120 + 60
    r = np.clip(base + rng.integers(-40, 40, (size, size)), 0, 255)
    g = np.clip(255 - base + rng.integers(-40, 40, (size, size)), 0, 255)
    b = rng.integers(0, 256, (size, size))
    a = np.clip(rng.integers(0, 320, (size, size)), 0, 255)
    return pack_rgba(r, g, b, a)


def random_soup(count, seed, extent=6.0, tri_size=2.5, textured=True, vertex_colors=True, alpha=False):
    """Random triangles scattered around the origin, many of them crossing the near plane and the frustum
    sides of a camera standing inside the cloud: exercises culling, clipping and every shader variant."""
    rng = np.random.default_rng(seed)
    centers = (rng.random((count, 3)) * 2 - 1) * extent
    points = (centers[:, None, :] + (rng.random((count, 3, 3)) * 2 - 1) * tri_size).reshape(-1, 3).astype(F)
    poly = np.zeros(count, abi.POLYGON_DTYPE)
    poly["pointIndices"][:, :3] = np.arange(count * 3, dtype=np.int32).reshape(count, 3)
    poly["pointIndices"][:, 3] = -1
    if textured:
        poly["texCoords"][:, :3, :] = (rng.random((count, 3, 4)) * 3 - 1).astype(F)
    if vertex_colors:
        col = rng.random((count, 3, 4)).astype(F)
        # a share of constant-colour and white triangles to hit the other shader variants
        kind = rng.integers(0, 4, count)
        col[kind == 1] = col[kind == 1][:, :1, :]
        col[kind == 2] = 1.0
        if not alpha:
            col[:, :, 3] = 1.0
        poly["colors"][:, :3, :] = col
    else:
        poly["colors"] = 1.0
    return {"points": points, "polygons": poly}


def mip_pyramid(level0, levels):
    """numpy statement of texture_generatePyramid (ref: api/textureAPI.cpp:44-87): whole buffer, smallest level first."""
    h, w = level0.shape
    log2w, log2h = int(math.log2(w)), int(math.log2(h))
    max_mip = min(log2w, log2h, levels - 1)
    mips = [level0.astype(np.uint32)]
    for _ in range(max_mip):
        src = mips[-1]
        ch = [(src >> s) & 255 for s in (0, 8, 16, 24)]
        out = [((c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2]) // 4) for c in ch]
        mips.append(pack_rgba(*out))
    return np.concatenate([m.reshape(-1) for m in reversed(mips)]).astype(np.uint32), max_mip
