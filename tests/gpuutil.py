"""Helpers shared by the GPU parity tests: run one operation through the product C ABI on device tensors and
through the oracle on numpy arrays built from the same seeded inputs."""
import ctypes as C

import numpy as np

import orcbind
from dfpsr_b200 import abi, lib


def dev(array):
    return lib.to_device(array)


def host_u32(tensor):
    return tensor.cpu().numpy().view(np.uint32)


def host_f32(tensor):
    return tensor.cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_same_u32(got, expected, what):
    if not np.array_equal(got, expected):
        bad = np.argwhere(got != expected)
        y, x = bad[0]
        raise AssertionError(f"{what}: {len(bad)} of {got.size} pixels differ, first at (x={x}, y={y}): got {got[y, x]:#010x}, expected {expected[y, x]:#010x}")


class CudaScene:
    """A model + textures resident on the device, mirrored by host arrays for the oracle."""

    def __init__(self, points, polygons, filter_=abi.FILTER_SOLID, diffuse_level0=None, diffuse_levels=1, light_level0=None, light_levels=1):
        self.points, self.polygons, self.filter = points, polygons, filter_
        self.d_tex = lib.DeviceTexture(diffuse_level0, diffuse_levels) if diffuse_level0 is not None else None
        self.l_tex = lib.DeviceTexture(light_level0, light_levels) if light_level0 is not None else None
        self.model = lib.DeviceModel(points, polygons, filter_, self.d_tex, self.l_tex)
        self.o_diffuse = orcbind.build_texture(diffuse_level0, diffuse_levels) if diffuse_level0 is not None else (None, None)
        self.o_light = orcbind.build_texture(light_level0, light_levels) if light_level0 is not None else (None, None)
        self.o_model, self._keep = orcbind.model_of(points, polygons, filter_, self.o_diffuse[1], self.o_light[1])

    def render_cuda(self, cuda, cam_params, color, depth, pack=abi.PACK_RGBA, model_to_world=None):
        """color/depth: numpy initial contents or None. Returns (color, depth) numpy after dfpsr_model_render."""
        cam = lib.camera(cam_params)
        m2w = model_to_world or abi.Transform3D.identity()
        tc = dev(color) if color is not None else None
        td = dev(depth) if depth is not None else None
        lib.check(cuda.dfpsr_model_render(C.byref(self.model.desc), C.byref(m2w), C.byref(lib.image(tc, pack)), C.byref(lib.image(td)), C.byref(cam), lib.stream_ptr()))
        return (host_u32(tc) if tc is not None else None), (host_f32(td) if td is not None else None)

    def render_oracle(self, oracle, cam_params, color, depth, pack=abi.PACK_RGBA, model_to_world=None):
        cam = orcbind.camera(cam_params)
        m2w = model_to_world or abi.Transform3D.identity()
        c = color.copy() if color is not None else None
        d = depth.copy() if depth is not None else None
        n = oracle.orc_model_render(C.byref(self.o_model), C.byref(m2w), C.byref(orcbind.image_of(c, pack)), C.byref(orcbind.image_of(d)), C.byref(cam))
        return c, d, n
