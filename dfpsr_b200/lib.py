"""ctypes binding of libdfpsr_b200.so (the product C ABI, include/dfpsr_b200.h) + torch helpers.

torch is used for device memory, streams and torch.distributed only; every kernel is in the library.
"""
import ctypes as C
import os

import numpy as np

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
# DFPSR_LIB selects another build of the same library (tuning variants built by tools/build_variants.py); the product build is the default
LIB_PATH = os.environ.get("DFPSR_LIB") or os.path.join(HERE, "libdfpsr_b200.so")
_lib = None

i32, u32, f32, vp, i64, u64, sz = C.c_int32, C.c_uint32, C.c_float, C.c_void_p, C.c_int64, C.c_uint64, C.c_size_t
P = C.POINTER
SIGNATURES = {
    "dfpsr_abi_version": (i32, []),
    "dfpsr_last_error": (C.c_char_p, []),
    "dfpsr_init": (i32, [i32]),
    "dfpsr_device_count": (i32, []),
    "dfpsr_launch_count": (u64, []),
    "dfpsr_reset_launch_count": (None, []),
    "dfpsr_profile_enable": (i32, [i32]),
    "dfpsr_profile_reset": (i32, []),
    "dfpsr_profile_count": (i32, []),
    "dfpsr_profile_read": (i32, [i32, P(C.c_char_p), P(C.c_double), P(i64)]),
    "dfpsr_malloc": (i32, [P(vp), sz]),
    "dfpsr_free": (i32, [vp]),
    "dfpsr_malloc_host": (i32, [P(vp), sz]),
    "dfpsr_free_host": (i32, [vp]),
    "dfpsr_upload": (i32, [vp, vp, sz, vp]),
    "dfpsr_download": (i32, [vp, vp, sz, vp]),
    "dfpsr_upload_2d": (i32, [vp, sz, vp, sz, sz, sz, vp]),
    "dfpsr_download_2d": (i32, [vp, sz, vp, sz, sz, sz, vp]),
    "dfpsr_stream_synchronize": (i32, [vp]),
    "dfpsr_camera_create_perspective": (i32, [P(abi.Camera), P(abi.Transform3D), f32, f32, f32, f32, f32]),
    "dfpsr_camera_create_orthogonal": (i32, [P(abi.Camera), P(abi.Transform3D), f32, f32, f32]),
    "dfpsr_camera_is_box_seen": (i32, [P(abi.Camera), vp, vp, P(abi.Transform3D)]),
    "dfpsr_texture_layout": (i32, [P(abi.Texture), i32, i32, i32]),
    "dfpsr_texture_generate_pyramid": (i32, [P(abi.Texture), vp]),
    "dfpsr_texture_from_image": (i32, [P(abi.Texture), P(abi.Image), vp]),
    "dfpsr_renderer_create": (i32, [P(vp)]),
    "dfpsr_renderer_destroy": (i32, [vp]),
    "dfpsr_renderer_set_async": (i32, [vp, i32]),
    "dfpsr_set_default_async": (i32, [i32]),
    "dfpsr_renderer_flush": (i32, [vp]),
    "dfpsr_flush": (i32, []),
    "dfpsr_renderer_set_precision": (i32, [vp, i32]),
    "dfpsr_set_default_precision": (i32, [i32]),
    "dfpsr_renderer_begin": (i32, [vp, P(abi.Image), P(abi.Image)]),
    "dfpsr_renderer_begin_cleared": (i32, [vp, P(abi.Image), P(abi.Image), u32, f32]),
    "dfpsr_renderer_occlude_from_box": (i32, [vp, vp, vp, P(abi.Transform3D), P(abi.Camera)]),
    "dfpsr_renderer_occlude_from_top_rows": (i32, [vp, P(abi.Camera), vp]),
    "dfpsr_renderer_occlude_from_existing_triangles": (i32, [vp, vp]),
    "dfpsr_renderer_has_occluders": (i32, [vp]),
    "dfpsr_renderer_is_box_visible": (i32, [vp, vp, vp, P(abi.Transform3D), P(abi.Camera), P(i32)]),
    "dfpsr_renderer_set_clip_rows": (i32, [vp, i32, i32]),
    "dfpsr_model_render_depth_batch": (i32, [vp, vp, vp, vp, i32, vp, i32, i32, f32, vp]),
    "dfpsr_renderer_give_task": (i32, [vp, P(abi.Model), P(abi.Transform3D), P(abi.Camera), vp]),
    "dfpsr_renderer_give_tasks": (i32, [vp, vp, vp, i32, P(abi.Camera), vp]),
    "dfpsr_renderer_give_task_triangles": (i32, [vp, vp, i32, P(abi.Texture), P(abi.Texture), i32, P(abi.Camera), vp]),
    "dfpsr_renderer_end": (i32, [vp, vp]),
    "dfpsr_renderer_set_debug_wireframe": (i32, [vp, i32]),
    "dfpsr_renderer_last_command_count": (i32, [vp, P(i64), vp]),
    "dfpsr_model_render": (i32, [P(abi.Model), P(abi.Transform3D), P(abi.Image), P(abi.Image), P(abi.Camera), vp]),
    "dfpsr_model_render_depth": (i32, [P(abi.Model), P(abi.Transform3D), P(abi.Image), P(abi.Camera), vp]),
    "dfpsr_model_render_views": (i32, [P(abi.Model), P(abi.Transform3D), vp, vp, vp, i32, i32, vp]),
    "dfpsr_project_points": (i32, [vp, i32, P(abi.Transform3D), P(abi.Camera), vp, vp]),
    "dfpsr_image_fill_rgba": (i32, [P(abi.Image), i32, i32, i32, i32, vp]),
    "dfpsr_image_fill_f32": (i32, [P(abi.Image), f32, vp]),
    "dfpsr_draw_copy_rgba": (i32, [P(abi.Image), P(abi.Image), i32, i32, vp]),
    "dfpsr_draw_copy_f32": (i32, [P(abi.Image), P(abi.Image), i32, i32, vp]),
    "dfpsr_draw_higher": (i32, [P(abi.Image)] * 6 + [i32, i32, f32, vp]),
    "dfpsr_draw_higher_batch": (i32, [P(abi.Image), P(abi.Image), P(abi.Image), vp, i32, vp]),
    "dfpsr_draw_rectangle_rgba": (i32, [P(abi.Image), i32, i32, i32, i32, vp, vp]),
    "dfpsr_draw_rectangle_f32": (i32, [P(abi.Image), i32, i32, i32, i32, f32, vp]),
    "dfpsr_draw_line_rgba": (i32, [P(abi.Image), i32, i32, i32, i32, vp, vp]),
    "dfpsr_draw_line_f32": (i32, [P(abi.Image), i32, i32, i32, i32, f32, vp]),
    "dfpsr_draw_alpha_filter": (i32, [P(abi.Image), P(abi.Image), i32, i32, vp]),
    "dfpsr_draw_max_alpha": (i32, [P(abi.Image), P(abi.Image), i32, i32, i32, vp]),
    "dfpsr_draw_alpha_clip": (i32, [P(abi.Image), P(abi.Image), i32, i32, i32, vp]),
    "dfpsr_draw_silhouette": (i32, [P(abi.Image), P(abi.Image), vp, i32, i32, vp]),
    "dfpsr_draw_rectangle_mono": (i32, [P(abi.Image), i32, i32, i32, i32, i32, i32, vp]),
    "dfpsr_draw_line_mono": (i32, [P(abi.Image), i32, i32, i32, i32, i32, i32, vp]),
    "dfpsr_draw_copy_formats": (i32, [P(abi.Image), i32, P(abi.Image), i32, i32, i32, vp]),
    "dfpsr_draw_higher_u16": (i32, [P(abi.Image)] * 6 + [i32, i32, i32, vp]),
    "dfpsr_light_directed": (i32, [P(abi.OrthoView), P(abi.Image), P(abi.Image), vp, f32, vp, i32, vp]),
    "dfpsr_light_point": (i32, [P(abi.OrthoView), vp, P(abi.Image), P(abi.Image), P(abi.Image), vp, f32, f32, vp, P(abi.Image), vp]),
    "dfpsr_light_frame": (i32, [P(abi.OrthoView), vp, P(abi.Image), P(abi.Image), P(abi.Image), P(abi.Image), P(abi.Image), vp, i32, vp, i32, vp]),
    "dfpsr_light_blend": (i32, [P(abi.Image), P(abi.Image), P(abi.Image), vp]),
    "dfpsr_filter_resize_scratch_bytes": (sz, [i32, i32, i32, i32]),
    "dfpsr_filter_resize": (i32, [P(abi.Image), P(abi.Image), i32, i32, vp, vp]),
    "dfpsr_filter_resize_u8": (i32, [P(abi.Image), P(abi.Image), i32, vp, vp]),
    "dfpsr_filter_map": (i32, [P(abi.Image), i32, vp, i32, P(abi.Image), i32, i32, vp]),
    "dfpsr_filter_block_magnify": (i32, [P(abi.Image), P(abi.Image), i32, i32, vp]),
    "dfpsr_ortho_system_create": (i32, [P(abi.OrthoSystem), f32, i32]),
    "dfpsr_ortho_camera_light_view": (i32, [P(abi.OrthoCamera), P(abi.OrthoView)]),
    "dfpsr_dense_model_triangle_count": (i32, [vp, i32]),
    "dfpsr_dense_model_build": (i32, [vp, i32, vp, i32, vp, vp, vp]),
    "dfpsr_dense_model_render": (i32, [vp, i32, vp, vp, P(abi.OrthoCamera), P(abi.Image), P(abi.Image), P(abi.Image), vp, P(abi.Transform3D), i32, vp, vp]),
    "dfpsr_sprite_generate_from_model": (i32, [vp, i32, vp, vp, P(abi.OrthoSystem), i32, P(abi.BakedSprite), vp]),
    "dfpsr_sprite_type_create": (i32, [vp, i32, i32, i32, P(abi.SpriteConfig), P(i32)]),
    "dfpsr_sprite_type_count": (i32, []),
    "dfpsr_model_type_create": (i32, [vp, i32, vp, vp, P(abi.HostModel), P(i32)]),
    "dfpsr_model_type_count": (i32, []),
    "dfpsr_sprite_world_create": (i32, [P(vp), P(abi.OrthoSystem), i32]),
    "dfpsr_sprite_world_destroy": (i32, [vp]),
    "dfpsr_sprite_world_add_background_sprite": (i32, [vp, P(abi.SpriteInstance)]),
    "dfpsr_sprite_world_add_background_model": (i32, [vp, P(abi.ModelInstance)]),
    "dfpsr_sprite_world_add_temporary_sprite": (i32, [vp, P(abi.SpriteInstance)]),
    "dfpsr_sprite_world_add_temporary_model": (i32, [vp, P(abi.ModelInstance)]),
    "dfpsr_sprite_world_remove_background_sprites": (i32, [vp, vp, vp, vp, vp]),
    "dfpsr_sprite_world_remove_background_models": (i32, [vp, vp, vp, vp, vp]),
    "dfpsr_sprite_world_create_temporary_point_light": (i32, [vp, vp, f32, f32, vp, i32]),
    "dfpsr_sprite_world_create_temporary_directed_light": (i32, [vp, vp, f32, vp]),
    "dfpsr_sprite_world_clear_temporary": (i32, [vp]),
    "dfpsr_sprite_world_get_camera_location": (i32, [vp, vp]),
    "dfpsr_sprite_world_set_camera_location": (i32, [vp, vp]),
    "dfpsr_sprite_world_move_camera_in_pixels": (i32, [vp, i32, i32]),
    "dfpsr_sprite_world_get_camera_direction_index": (i32, [vp, P(i32)]),
    "dfpsr_sprite_world_set_camera_direction_index": (i32, [vp, i32]),
    "dfpsr_sprite_world_find_ground_at_pixel": (i32, [vp, i32, i32, i32, i32, vp]),
    "dfpsr_sprite_world_plan_frame": (i32, [vp, i32, i32, P(P(abi.SpriteWorldOp)), P(i32)]),
    "dfpsr_sprite_world_draw": (i32, [vp, P(abi.Image), vp]),
    "dfpsr_sprite_world_draw_host": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "dfpsr_sprite_world_get_buffers": (i32, [vp, P(abi.Image), P(abi.Image), P(abi.Image), P(abi.Image)]),
    "dfpsr_session_create": (i32, [P(vp)]),
    "dfpsr_session_destroy": (i32, [vp]),
    "dfpsr_session_upload_model": (i32, [vp, P(abi.HostModel), P(i32)]),
    "dfpsr_session_render_views_host": (i32, [vp, i32, P(abi.Transform3D), vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "dfpsr_session_render_frame_host": (i32, [vp, i32, P(abi.Transform3D), P(abi.Camera), vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "dfpsr_selftest_rsqrt": (i32, [u32, u32, P(u64), vp]),
    "dfpsr_import_ply": (i32, [C.c_char_p, sz, i32, P(abi.Transform3D), P(abi.ImportedModel)]),
    "dfpsr_import_dmf1": (i32, [C.c_char_p, sz, i32, P(abi.ImportedModel)]),
    "dfpsr_import_free": (None, [P(abi.ImportedModel)]),
    "dfpsr_canvas_show": (i32, [P(abi.Image), i32, vp, i32, i32, i32, i32, vp]),
    "dfpsr_filter_map_program": (i32, [P(abi.Image), C.c_char_p, vp, i32, i32, i32, vp]),
    "dfpsr_peer_alloc": (i32, [P(vp), sz, vp]),
    "dfpsr_peer_free": (i32, [vp]),
    "dfpsr_peer_open": (i32, [P(vp), vp]),
    "dfpsr_peer_close": (i32, [vp]),
    "dfpsr_peer_signal": (i32, [P(vp), i32, u32, vp]),
    "dfpsr_peer_wait": (i32, [vp, i32, u32, u32, vp, vp]),
    "dfpsr_peer_reset_status": (i32, [vp, vp]),
}


class DfpsrError(RuntimeError):
    pass


def load():
    """Loads the CUDA library. Raises if it has not been built — there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DfpsrError(f"{LIB_PATH} is missing: build it with `python -m dfpsr_b200.build` (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the library does not export what the header declares
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise DfpsrError(load().dfpsr_last_error().decode("utf-8", "replace"))


def profile_snapshot():
    """{kernel name: (milliseconds, launches)} accumulated since dfpsr_profile_reset."""
    handle = load()
    out = {}
    for i in range(handle.dfpsr_profile_count()):
        name, ms, n = C.c_char_p(), C.c_double(), C.c_int64()
        check(handle.dfpsr_profile_read(i, C.byref(name), C.byref(ms), C.byref(n)))
        out[name.value.decode()] = (ms.value, n.value)
    return out


def imported_arrays(model):
    """(points (n, 3) float32, polygons (POLYGON_DTYPE), parts [(name, diffuse, light, first, count)]) copied out of a dfpsr_imported_model."""
    pts = np.ctypeslib.as_array(model.points, shape=(model.pointCount * 3,)).reshape(-1, 3).copy() if model.pointCount else np.zeros((0, 3), np.float32)
    if model.polygonCount:
        raw = C.string_at(model.polygons, model.polygonCount * abi.POLYGON_DTYPE.itemsize)
        polys = np.frombuffer(raw, dtype=abi.POLYGON_DTYPE).copy()
    else:
        polys = np.zeros(0, abi.POLYGON_DTYPE)
    parts = [(model.parts[i].name.decode(), model.parts[i].diffuseName.decode(), model.parts[i].lightName.decode(), model.parts[i].firstPolygon, model.parts[i].polygonCount) for i in range(model.partCount)]
    return pts, polys, parts


def import_model(kind, text, flip_x=False, axis=None, detail_level=2):
    """dfpsr_import_ply / dfpsr_import_dmf1 on a str or bytes; returns (points, polygons, parts, filter, (min, max))."""
    handle = load()
    data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    out = abi.ImportedModel()
    if kind == "ply":
        check(handle.dfpsr_import_ply(data, len(data), 1 if flip_x else 0, C.byref(axis) if axis is not None else None, C.byref(out)))
    else:
        check(handle.dfpsr_import_dmf1(data, len(data), detail_level, C.byref(out)))
    try:
        pts, polys, parts = imported_arrays(out)
        return pts, polys, parts, out.filter, (list(out.minBound), list(out.maxBound))
    finally:
        handle.dfpsr_import_free(C.byref(out))


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def image(tensor, pack=abi.PACK_RGBA):
    """dfpsr_image over a 2-D CUDA tensor of 4-byte elements (int32/uint32 view for RGBA8, float32 for F32)."""
    if tensor is None:
        return abi.Image.null()
    assert tensor.is_cuda and tensor.dim() == 2 and tensor.element_size() == 4 and tensor.stride(1) == 1
    return abi.Image(tensor.data_ptr(), tensor.shape[1], tensor.shape[0], tensor.stride(0) * 4, pack)


def image_from_ptr(ptr, width, height, stride_bytes=None, pack=abi.PACK_RGBA):
    """dfpsr_image over raw device memory (e.g. a peer-mapped frame of another rank, dfpsr_peer_open)."""
    return abi.Image(ptr, width, height, stride_bytes if stride_bytes is not None else width * 4, pack)


class _RawCudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def tensor_from_ptr(ptr, shape, typestr="<i4", device="cuda"):
    """A torch view (no copy, no ownership) of raw device memory through __cuda_array_interface__."""
    import torch
    return torch.as_tensor(_RawCudaArray(ptr, shape, typestr), device=device)


def to_device(array, device="cuda"):
    import torch
    a = np.ascontiguousarray(array)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).to(device)


def camera(params):
    """Runs dfpsr_camera_create_* on a POD that only holds the constructor arguments."""
    lib = load()
    out = abi.Camera()
    loc = params.location
    if params.perspective:
        check(lib.dfpsr_camera_create_perspective(C.byref(out), C.byref(loc), params.imageWidth, params.imageHeight, params.widthSlope, params.nearClip, params.farClip))
    else:
        check(lib.dfpsr_camera_create_orthogonal(C.byref(out), C.byref(loc), params.imageWidth, params.imageHeight, params.widthSlope))
    return out


def model_bounds(points):
    """ref: implementation/render/model/Model.cpp:281-288 — the box starts at the origin and only grows."""
    if len(points) == 0:
        return [0.0] * 3, [0.0] * 3
    return [float(min(v, 0.0)) for v in points.min(axis=0)], [float(max(v, 0.0)) for v in points.max(axis=0)]


class DeviceTexture:
    def __init__(self, level0, levels, device="cuda"):
        """Uploads level 0 and generates the pyramid on the device (texture_create + texture_generatePyramid)."""
        import torch
        lib = load()
        h, w = level0.shape
        self.desc = abi.Texture()
        check(lib.dfpsr_texture_layout(C.byref(self.desc), w, h, levels))
        self.pixels = torch.zeros(self.desc.totalPixels, dtype=torch.int32, device=device)
        self.pixels[self.desc.startOffset:] = to_device(np.ascontiguousarray(level0, np.uint32).reshape(-1), device)
        self.desc.data = self.pixels.data_ptr()
        check(lib.dfpsr_texture_generate_pyramid(C.byref(self.desc), stream_ptr()))


class DeviceModel:
    def __init__(self, points, polygons, filter_=abi.FILTER_SOLID, diffuse=None, light=None, device="cuda"):
        pts = np.ascontiguousarray(points, np.float32)
        poly = np.ascontiguousarray(polygons)
        assert poly.dtype == abi.POLYGON_DTYPE
        self.points = to_device(pts.reshape(-1), device)
        self.polygons = to_device(poly.view(np.uint8).reshape(-1), device)
        self.diffuse, self.light = diffuse, light
        m = abi.Model()
        m.points = self.points.data_ptr()
        m.pointCount = len(pts)
        m.polygons = self.polygons.data_ptr()
        m.polygonCount = len(poly)
        m.filter = filter_
        if diffuse is not None:
            m.diffuse = diffuse.desc
        if light is not None:
            m.light = light.desc
        mn, mx = model_bounds(pts)
        m.minBound[:] = mn
        m.maxBound[:] = mx
        self.desc = m
