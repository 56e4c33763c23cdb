"""ctypes binding of oracle/_ref/libdfpsr_ref_{sse,scalar}.so (the compiled, unmodified reference).

TEST INFRASTRUCTURE: imported only by tests/ and by bench.py's reference / cpu_baseline legs.
"""
import atexit
import ctypes as C
import os

import numpy as np

from dfpsr_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

_libs = {}


def available(flavour="scalar"):
    return os.path.exists(os.path.join(REF_DIR, f"libdfpsr_ref_{flavour}.so"))


def load(flavour="scalar"):
    if flavour in _libs:
        return _libs[flavour]
    lib = C.CDLL(os.path.join(REF_DIR, f"libdfpsr_ref_{flavour}.so"))
    i, f, p, d = C.c_int, C.c_float, C.c_void_p, C.c_double
    T, Cam, V = C.POINTER(abi.Transform3D), C.POINTER(abi.Camera), C.POINTER(abi.OrthoView)
    sig = {
        "ref_flavour": (i, []), "ref_thread_count": (i, []), "ref_time_seconds": (d, []), "ref_free_all": (None, []), "ref_shutdown": (None, []),
        "ref_image_create_rgba": (i, [i, i, i]), "ref_image_create_f32": (i, [i, i]), "ref_image_sub": (i, [i, i, i, i, i]),
        "ref_image_width": (i, [i]), "ref_image_height": (i, [i]), "ref_image_stride": (i, [i]),
        "ref_image_write": (None, [i, p, i]), "ref_image_read": (None, [i, p, i]),
        "ref_image_fill_rgba": (None, [i, i, i, i, i]), "ref_image_fill_f32": (None, [i, f]),
        "ref_texture_create": (i, [i, i, i, p]), "ref_texture_from_image": (i, [i, i]),
        "ref_texture_info": (None, [i, p]), "ref_texture_read": (None, [i, p]),
        "ref_model_create": (i, [p, i, p, i, i, i, i]), "ref_model_bounds": (None, [i, p, p]),
        "ref_model_render": (None, [i, T, i, i, Cam, i]),
        "ref_models_render_frame": (None, [p, p, i, i, i, Cam]),
        "ref_model_render_depth": (None, [i, T, i, Cam]),
        "ref_renderer_begin": (None, [i, i]), "ref_renderer_give_task": (None, [i, T, Cam]),
        "ref_renderer_occlude_from_box": (None, [p, p, T, Cam]), "ref_renderer_occlude_from_top_rows": (None, [Cam]),
        "ref_renderer_occlude_from_existing_triangles": (None, []), "ref_renderer_is_box_visible": (i, [p, p, T, Cam]),
        "ref_renderer_has_occluders": (i, []), "ref_renderer_end": (None, []), "ref_renderer_end_wireframe": (None, []),
        "ref_terrain_frame": (d, [i, T, i, i, Cam]),
        "ref_project_points": (None, [p, i, T, Cam, p]),
        "ref_camera_fill": (None, [Cam]), "ref_camera_is_box_seen": (i, [Cam, p, p, T]),
        "ref_draw_higher": (None, [i, i, i, i, i, i, i, i, f]), "ref_draw_copy": (None, [i, i, i, i]),
        "ref_ortho_view": (None, [f, i, i, V, p]),
        "ref_light_directed": (None, [V, i, i, p, f, p, i]),
        "ref_light_point": (None, [V, p, i, i, i, p, f, f, p, i]),
        "ref_light_blend": (None, [i, i, i]),
        "ref_sprite_type_load": (i, [C.c_char_p, C.c_char_p]), "ref_image_load": (i, [C.c_char_p]), "ref_string_to_double": (C.c_double, [C.c_char_p]),
        "ref_filter_resize": (i, [i, i, i, i]), "ref_filter_resize_u8": (i, [i, i, i, i]), "ref_filter_map": (None, [i, i, p, i, i, i]),
        "ref_filter_block_magnify": (None, [i, i, i, i]),
        "ref_image_create_u8": (i, [i, i, p]),
        "ref_image_create_u16": (i, [i, i, p]), "ref_image_read_mono": (None, [i, p]),
        "ref_draw_rectangle_mono": (None, [i, i, i, i, i, i]), "ref_draw_line_mono": (None, [i, i, i, i, i, i]),
        "ref_draw_copy_formats": (None, [i, i, i, i]), "ref_draw_higher_u16": (None, [i, i, i, i, i, i, i, i, i]),
        "ref_draw_rectangle_rgba": (None, [i, i, i, i, i, p]), "ref_draw_rectangle_f32": (None, [i, i, i, i, i, f]),
        "ref_draw_line_rgba": (None, [i, i, i, i, i, p]), "ref_draw_line_f32": (None, [i, i, i, i, i, f]),
        "ref_draw_alpha_filter": (None, [i, i, i, i]), "ref_draw_max_alpha": (None, [i, i, i, i, i]), "ref_draw_alpha_clip": (None, [i, i, i, i, i]),
        "ref_draw_silhouette": (None, [i, i, p, i, i]),
        "ref_ortho_system": (None, [f, i, C.POINTER(abi.OrthoSystem)]),
        "ref_dense_model_create": (i, [i]), "ref_dense_model_triangles": (i, [i, p, p, p]),
        "ref_dense_model_render": (None, [i, f, i, i, i, i, i, f, f, T, i, p]),
        "ref_sprite_type_create": (i, [i, C.c_char_p, C.c_char_p, C.c_char_p]), "ref_sprite_type_count": (i, []),
        "ref_model_type_create": (i, [i, i]),
        "ref_sprite_generate_from_model": (i, [i, i, f, i, i, p]),
        "ref_world_create": (i, [f, i, i]),
        "ref_world_add_background_sprite": (None, [i, C.POINTER(abi.SpriteInstance)]), "ref_world_add_background_model": (None, [i, C.POINTER(abi.ModelInstance)]),
        "ref_world_add_temporary_sprite": (None, [i, C.POINTER(abi.SpriteInstance)]), "ref_world_add_temporary_model": (None, [i, C.POINTER(abi.ModelInstance)]),
        "ref_world_remove_background_sprites": (None, [i, p, p]), "ref_world_remove_background_models": (None, [i, p, p]),
        "ref_world_point_light": (None, [i, p, f, f, p, i]), "ref_world_directed_light": (None, [i, p, f, p]),
        "ref_world_clear_temporary": (None, [i]),
        "ref_world_set_camera_location": (None, [i, p]), "ref_world_get_camera_location": (None, [i, p]),
        "ref_world_move_camera_in_pixels": (None, [i, i, i]), "ref_world_set_camera_direction_index": (None, [i, i]),
        "ref_world_find_ground_at_pixel": (None, [i, i, i, i, p]),
        "ref_import_ply": (i, [C.c_char_p, i, T]), "ref_import_dmf1": (i, [C.c_char_p, i]),
        "ref_sdk_terrain_scene": (i, [C.c_char_p, p]),
        "ref_model_dump_counts": (None, [i, p]), "ref_model_dump": (None, [i, p, p, p]), "ref_model_dump_name": (None, [i, i, C.c_char_p, i]),
        "ref_world_draw": (None, [i, i]), "ref_world_read_buffers": (None, [i, p, p, p, p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _libs[flavour] = lib
    atexit.register(lib.ref_shutdown)
    return lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Ref:
    """Convenience layer: numpy in, numpy out."""

    def __init__(self, flavour="scalar"):
        self.lib = load(flavour)
        self.flavour = flavour

    # ---- images
    def rgba(self, array=None, shape=None, pack=abi.PACK_RGBA):
        if array is not None:
            shape = array.shape
        h, w = shape
        iid = self.lib.ref_image_create_rgba(w, h, pack)
        if array is not None:
            a = np.ascontiguousarray(array, np.uint32)
            self.lib.ref_image_write(iid, ptr(a), w * 4)
        return iid

    def f32(self, array=None, shape=None, fill=None):
        if array is not None:
            shape = array.shape
        h, w = shape
        iid = self.lib.ref_image_create_f32(w, h)
        if array is not None:
            a = np.ascontiguousarray(array, np.float32)
            self.lib.ref_image_write(iid, ptr(a), w * 4)
        elif fill is not None:
            self.lib.ref_image_fill_f32(iid, fill)
        return iid

    def read(self, iid, dtype):
        w, h = self.lib.ref_image_width(iid), self.lib.ref_image_height(iid)
        out = np.empty((h, w), dtype)
        self.lib.ref_image_read(iid, ptr(out), w * 4)
        return out

    def read_rgba(self, iid):
        return self.read(iid, np.uint32)

    def read_f32(self, iid):
        return self.read(iid, np.float32)

    # ---- textures / models
    def texture(self, level0, levels):
        a = np.ascontiguousarray(level0, np.uint32)
        return self.lib.ref_texture_create(a.shape[1], a.shape[0], levels, ptr(a))

    def texture_pixels(self, tid):
        info = np.zeros(6, np.uint32)
        self.lib.ref_texture_info(tid, ptr(info))
        out = np.empty(int(info[5]), np.uint32)
        self.lib.ref_texture_read(tid, ptr(out))
        return out, info

    def model(self, points, polygons, filter_=abi.FILTER_SOLID, diffuse=-1, light=-1):
        pts = np.ascontiguousarray(points, np.float32)
        poly = np.ascontiguousarray(polygons)
        return self.lib.ref_model_create(ptr(pts), len(pts), ptr(poly), len(poly), filter_, diffuse, light)

    def render(self, model, camera, color, depth, mode=1, model_to_world=None):
        m2w = model_to_world or abi.Transform3D.identity()
        self.lib.ref_model_render(model, C.byref(m2w), color, depth, C.byref(camera), mode)

    def render_depth(self, model, camera, depth, model_to_world=None):
        m2w = model_to_world or abi.Transform3D.identity()
        self.lib.ref_model_render_depth(model, C.byref(m2w), depth, C.byref(camera))

    def project(self, points, camera, model_to_world=None):
        m2w = model_to_world or abi.Transform3D.identity()
        pts = np.ascontiguousarray(points, np.float32)
        out = np.zeros(len(pts), abi.PROJECTED_DTYPE)
        self.lib.ref_project_points(ptr(pts), len(pts), C.byref(m2w), C.byref(camera), ptr(out))
        return out

    # ---- importers
    def dump_model(self, model):
        """(points, polygons, parts [(name, polygon count)], filter, requested texture names) of a reference model."""
        counts = np.zeros(5, np.int32)
        self.lib.ref_model_dump_counts(model, ptr(counts))
        points = np.zeros((int(counts[0]), 3), np.float32)
        polygons = np.zeros(int(counts[2]), abi.POLYGON_DTYPE)
        per_part = np.zeros(max(int(counts[1]), 1), np.int32)
        self.lib.ref_model_dump(model, ptr(points), ptr(polygons), ptr(per_part))

        def name(index):
            buf = C.create_string_buffer(256)
            self.lib.ref_model_dump_name(model, index, buf, 256)
            return buf.value.decode()
        parts = [(name(k), int(per_part[k])) for k in range(int(counts[1]))]
        return points, polygons, parts, int(counts[3]), [name(-1 - k) for k in range(int(counts[4]))]

    def import_ply(self, path, flip_x=False, axis=None):
        return self.dump_model(self.lib.ref_import_ply(path.encode(), 1 if flip_x else 0, C.byref(axis or abi.Transform3D.identity())))

    def import_dmf1(self, text, detail_level=2):
        return self.dump_model(self.lib.ref_import_dmf1(text.encode("utf-8"), detail_level))

    def free_all(self):
        self.lib.ref_free_all()
