"""A frame that exercises the renderer's occlusion grid (ref: api/rendererAPI.cpp:181-477): an occluder box in front of a crowd of
small models, renderer_isBoxVisible queries, renderer_occludeFromExistingTriangles in the middle of the submission and
renderer_occludeFromTopRows on a pre-filled depth buffer. Driven through the compiled reference, the C oracle and the CUDA library.
TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

from dfpsr_b200 import abi, scenes
from sandbox_scene import box_model

F = np.float32


def build(seed=3, width=640, height=360, models=60, top_rows=False, perspective=True):
    rng = np.random.default_rng(seed)
    sc = {"width": width, "height": height, "top_rows": top_rows}
    sc["camera"] = abi.camera_params(perspective, scenes.look_at_transform((0.0, 1.0, -6.0), (0.0, 0.5, 0.0)), width, height, width_slope=1.0 if perspective else 6.0)
    # a solid wall (the occluder is a slightly smaller box inside it, as a game would declare it)
    wall_pts, wall_poly = box_model((5.0, 3.0, 0.6))
    wall_poly = wall_poly[::2].copy()  # outward faces only
    wall_poly["colors"][:, :, :3] = (0.8, 0.7, 0.6)
    sc["wall"] = (wall_pts, wall_poly)
    sc["wall_transform"] = abi.Transform3D.make((0.0, 1.0, 0.0), ((1, 0, 0), (0, 1, 0), (0, 0, 1)))
    sc["occluder_box"] = (np.array([-2.4, -1.4, -0.25], F), np.array([2.4, 1.4, 0.25], F))
    sc["models"] = []
    for i in range(models):
        pts, poly = box_model(rng.random(3) * 0.5 + 0.15)
        poly = poly[::2].copy()
        poly["colors"][:, :, :3] = rng.random(3)
        behind = i % 3 != 0
        pos = ((rng.random() * 2 - 1) * (2.0 if behind else 6.0), rng.random() * 2.0, (2.0 + rng.random() * 6.0) if behind else (rng.random() * 10 - 4))
        sc["models"].append({"points": pts, "polygons": poly, "transform": abi.Transform3D.make(pos, ((1, 0, 0), (0, 1, 0), (0, 0, 1)))})
    sc["existing_after"] = models // 2
    depth = np.zeros((height, width), F) if perspective else np.full((height, width), 1e9, F)
    if top_rows:  # something already drawn in the upper part of the depth buffer (1 / distance for perspective cameras)
        depth[: height // 3, width // 4: width // 2] = F(1.0 / 3.0) if perspective else F(3.0)
    sc["depth0"] = depth
    sc["color0"] = np.zeros((height, width), np.uint32)
    return sc


def bounds_of(points):
    lo = np.minimum(points.min(axis=0), 0.0).astype(F)
    hi = np.maximum(points.max(axis=0), 0.0).astype(F)
    return lo, hi


def run_reference(ref, sc, wireframe=False):
    import refbind
    L = ref.lib
    cam = abi.Camera.from_buffer_copy(sc["camera"])
    L.ref_camera_fill(C.byref(cam))
    color, depth = ref.rgba(sc["color0"]), ref.f32(sc["depth0"])
    wall = ref.model(*sc["wall"])
    ids = [ref.model(m["points"], m["polygons"]) for m in sc["models"]]
    L.ref_renderer_begin(color, depth)
    L.ref_renderer_give_task(wall, C.byref(sc["wall_transform"]), C.byref(cam))
    lo, hi = sc["occluder_box"]
    L.ref_renderer_occlude_from_box(refbind.ptr(lo), refbind.ptr(hi), C.byref(sc["wall_transform"]), C.byref(cam))
    if sc["top_rows"]:
        L.ref_renderer_occlude_from_top_rows(C.byref(cam))
    visible = []
    for i, (m, mid) in enumerate(zip(sc["models"], ids)):
        blo, bhi = bounds_of(m["points"])
        visible.append(L.ref_renderer_is_box_visible(refbind.ptr(blo), refbind.ptr(bhi), C.byref(m["transform"]), C.byref(cam)))
        L.ref_renderer_give_task(mid, C.byref(m["transform"]), C.byref(cam))
        if i == sc["existing_after"]:
            L.ref_renderer_occlude_from_existing_triangles()
    assert L.ref_renderer_has_occluders() == 1
    if wireframe:
        L.ref_renderer_end_wireframe()
    else:
        L.ref_renderer_end()
    return {"color": ref.read_rgba(color), "depth": ref.read_f32(depth), "visible": visible}


def run_oracle(lib, sc, wireframe=False):
    import orcbind
    IM = orcbind.image_of
    cam = orcbind.camera(sc["camera"])
    color, depth = sc["color0"].copy(), sc["depth0"].copy()
    r = lib.orc_renderer_create()
    wall, keep = orcbind.model_of(*sc["wall"])
    models = [orcbind.model_of(m["points"], m["polygons"]) for m in sc["models"]]
    lib.orc_renderer_begin(r, C.byref(IM(color)), C.byref(IM(depth)))
    lib.orc_renderer_give_task(r, C.byref(wall), C.byref(sc["wall_transform"]), C.byref(cam))
    lo, hi = sc["occluder_box"]
    lib.orc_renderer_occlude_from_box(r, orcbind.ptr(lo), orcbind.ptr(hi), C.byref(sc["wall_transform"]), C.byref(cam))
    if sc["top_rows"]:
        lib.orc_renderer_occlude_from_top_rows(r, C.byref(cam))
    visible = []
    for i, (m, (om, _k)) in enumerate(zip(sc["models"], models)):
        blo, bhi = bounds_of(m["points"])
        visible.append(lib.orc_renderer_is_box_visible(r, orcbind.ptr(blo), orcbind.ptr(bhi), C.byref(m["transform"]), C.byref(cam)))
        lib.orc_renderer_give_task(r, C.byref(om), C.byref(m["transform"]), C.byref(cam))
        if i == sc["existing_after"]:
            lib.orc_renderer_occlude_from_existing_triangles(r)
    assert lib.orc_renderer_has_occluders(r) == 1
    skipped = C.c_int64()
    lib.orc_renderer_set_debug_wireframe(r, 1 if wireframe else 0)
    n = lib.orc_renderer_end(r, C.byref(skipped))
    lib.orc_renderer_destroy(r)
    return {"color": color, "depth": depth, "visible": visible, "commands": n, "occluded": skipped.value}


def run_oracle_without_occluders(lib, sc, wireframe=False):
    """The same models in the same order through the oracle's renderer with no occluder calls (nothing is skipped)."""
    import orcbind
    IM = orcbind.image_of
    cam = orcbind.camera(sc["camera"])
    color, depth = sc["color0"].copy(), sc["depth0"].copy()
    r = lib.orc_renderer_create()
    wall, keep = orcbind.model_of(*sc["wall"])
    models = [orcbind.model_of(m["points"], m["polygons"]) for m in sc["models"]]
    lib.orc_renderer_begin(r, C.byref(IM(color)), C.byref(IM(depth)))
    lib.orc_renderer_give_task(r, C.byref(wall), C.byref(sc["wall_transform"]), C.byref(cam))
    for m, (om, _k) in zip(sc["models"], models):
        lib.orc_renderer_give_task(r, C.byref(om), C.byref(m["transform"]), C.byref(cam))
    lib.orc_renderer_set_debug_wireframe(r, 1 if wireframe else 0)
    n = lib.orc_renderer_end(r, None)
    lib.orc_renderer_destroy(r)
    return {"color": color, "depth": depth, "commands": n}


def run_cuda(cuda, sc, wireframe=False):
    from dfpsr_b200 import lib
    cam = lib.camera(sc["camera"])
    tc, td = lib.to_device(sc["color0"]), lib.to_device(sc["depth0"])
    wall = lib.DeviceModel(*sc["wall"])
    models = [lib.DeviceModel(m["points"], m["polygons"]) for m in sc["models"]]
    r = C.c_void_p()
    s = lib.stream_ptr()
    lib.check(cuda.dfpsr_renderer_create(C.byref(r)))
    lib.check(cuda.dfpsr_renderer_begin(r, C.byref(lib.image(tc)), C.byref(lib.image(td))))
    lib.check(cuda.dfpsr_renderer_give_task(r, C.byref(wall.desc), C.byref(sc["wall_transform"]), C.byref(cam), s))
    lo, hi = sc["occluder_box"]
    lib.check(cuda.dfpsr_renderer_occlude_from_box(r, lo.ctypes.data, hi.ctypes.data, C.byref(sc["wall_transform"]), C.byref(cam)))
    if sc["top_rows"]:
        lib.check(cuda.dfpsr_renderer_occlude_from_top_rows(r, C.byref(cam), s))
    visible = []
    for i, (m, dm) in enumerate(zip(sc["models"], models)):
        blo, bhi = bounds_of(m["points"])
        v = C.c_int32()
        lib.check(cuda.dfpsr_renderer_is_box_visible(r, blo.ctypes.data, bhi.ctypes.data, C.byref(m["transform"]), C.byref(cam), C.byref(v)))
        visible.append(v.value)
        lib.check(cuda.dfpsr_renderer_give_task(r, C.byref(dm.desc), C.byref(m["transform"]), C.byref(cam), s))
        if i == sc["existing_after"]:
            lib.check(cuda.dfpsr_renderer_occlude_from_existing_triangles(r, s))
    assert cuda.dfpsr_renderer_has_occluders(r) == 1
    lib.check(cuda.dfpsr_renderer_set_debug_wireframe(r, 1 if wireframe else 0))
    lib.check(cuda.dfpsr_renderer_end(r, s))
    count = C.c_int64()
    lib.check(cuda.dfpsr_renderer_last_command_count(r, C.byref(count), s))
    lib.check(cuda.dfpsr_renderer_destroy(r))
    return {"color": tc.cpu().numpy().view(np.uint32), "depth": td.cpu().numpy(), "visible": visible, "commands": count.value}


def run_cuda_batched(cuda, sc):
    """The same frame with the models handed over through dfpsr_renderer_give_tasks (whole-model tests on the device): one call for the
    models up to the occludeFromExistingTriangles call, one for the rest — the occluders do not change inside either group."""
    from dfpsr_b200 import abi, lib
    cam = lib.camera(sc["camera"])
    tc, td = lib.to_device(sc["color0"]), lib.to_device(sc["depth0"])
    wall = lib.DeviceModel(*sc["wall"])
    models = [lib.DeviceModel(m["points"], m["polygons"]) for m in sc["models"]]
    r = C.c_void_p()
    s = lib.stream_ptr()
    lib.check(cuda.dfpsr_renderer_create(C.byref(r)))
    lib.check(cuda.dfpsr_renderer_begin(r, C.byref(lib.image(tc)), C.byref(lib.image(td))))
    lib.check(cuda.dfpsr_renderer_give_task(r, C.byref(wall.desc), C.byref(sc["wall_transform"]), C.byref(cam), s))
    lo, hi = sc["occluder_box"]
    lib.check(cuda.dfpsr_renderer_occlude_from_box(r, lo.ctypes.data, hi.ctypes.data, C.byref(sc["wall_transform"]), C.byref(cam)))
    if sc["top_rows"]:
        lib.check(cuda.dfpsr_renderer_occlude_from_top_rows(r, C.byref(cam), s))
    split = sc["existing_after"] + 1 if 0 <= sc["existing_after"] < len(models) else len(models)
    for first, last in ((0, split), (split, len(models))):
        n = last - first
        if n > 0:
            descs = (abi.Model * n)(*[models[i].desc for i in range(first, last)])
            transforms = (abi.Transform3D * n)(*[sc["models"][i]["transform"] for i in range(first, last)])
            lib.check(cuda.dfpsr_renderer_give_tasks(r, descs, transforms, n, C.byref(cam), s))
        if first == 0 and 0 <= sc["existing_after"] < len(models):
            lib.check(cuda.dfpsr_renderer_occlude_from_existing_triangles(r, s))
    lib.check(cuda.dfpsr_renderer_end(r, s))
    count = C.c_int64()
    lib.check(cuda.dfpsr_renderer_last_command_count(r, C.byref(count), s))
    lib.check(cuda.dfpsr_renderer_destroy(r))
    return {"color": tc.cpu().numpy().view(np.uint32), "depth": td.cpu().numpy(), "commands": count.value}
