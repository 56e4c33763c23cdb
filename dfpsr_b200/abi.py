"""ctypes mirrors of the POD types in include/dfpsr_b200.h.

Shared by the product binding (dfpsr_b200/lib.py), the oracle binding and the reference wrapper binding
(tests/ only), so that the same scene description can be handed to all three.
"""
import ctypes as C

import numpy as np

PACK_RGBA, PACK_BGRA, PACK_ARGB, PACK_ABGR = 0, 1, 2, 3
FILTER_SOLID, FILTER_ALPHA = 0, 1
SAMPLER_NEAREST, SAMPLER_LINEAR = 0, 1
MAP_XOR_PATTERN, MAP_AFFINE, MAP_CONSTANT = 0, 1, 2
FORMAT_U8, FORMAT_U16, FORMAT_F32, FORMAT_RGBA_U8 = 1, 2, 3, 4


class Transform3D(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("xAxis", C.c_float * 3), ("yAxis", C.c_float * 3), ("zAxis", C.c_float * 3)]

    @staticmethod
    def identity():
        return Transform3D.make((0, 0, 0), ((1, 0, 0), (0, 1, 0), (0, 0, 1)))

    @staticmethod
    def make(position, axes):
        t = Transform3D()
        t.position[:] = [float(np.float32(v)) for v in position]
        t.xAxis[:] = [float(np.float32(v)) for v in axes[0]]
        t.yAxis[:] = [float(np.float32(v)) for v in axes[1]]
        t.zAxis[:] = [float(np.float32(v)) for v in axes[2]]
        return t


class Matrix3x3(C.Structure):
    _fields_ = [("xAxis", C.c_float * 3), ("yAxis", C.c_float * 3), ("zAxis", C.c_float * 3)]


class Camera(C.Structure):
    _fields_ = [
        ("perspective", C.c_int32),
        ("location", Transform3D),
        ("widthSlope", C.c_float), ("heightSlope", C.c_float), ("invWidthSlope", C.c_float), ("invHeightSlope", C.c_float),
        ("imageWidth", C.c_float), ("imageHeight", C.c_float), ("nearClip", C.c_float), ("farClip", C.c_float),
        ("cullPlaneCount", C.c_int32), ("clipPlaneCount", C.c_int32),
        ("cullPlanes", (C.c_float * 4) * 6),
        ("clipPlanes", (C.c_float * 4) * 6),
    ]


class Polygon(C.Structure):
    _fields_ = [("pointIndices", C.c_int32 * 4), ("texCoords", (C.c_float * 4) * 4), ("colors", (C.c_float * 4) * 4)]


POLYGON_DTYPE = np.dtype([("pointIndices", np.int32, (4,)), ("texCoords", np.float32, (4, 4)), ("colors", np.float32, (4, 4))])
assert POLYGON_DTYPE.itemsize == 144 and C.sizeof(Polygon) == 144


class ProjectedPoint(C.Structure):
    _fields_ = [("cs", C.c_float * 3), ("is_", C.c_float * 2), ("pad_", C.c_int32), ("flat", C.c_int64 * 2)]


PROJECTED_DTYPE = np.dtype([("cs", np.float32, (3,)), ("is", np.float32, (2,)), ("pad", np.int32), ("flat", np.int64, (2,))])
assert PROJECTED_DTYPE.itemsize == 40 and C.sizeof(ProjectedPoint) == 40

TRIANGLE_DTYPE = np.dtype([("pos", PROJECTED_DTYPE, (3,)), ("colors", np.float32, (3, 4)), ("texCoords", np.float32, (3, 4))])
assert TRIANGLE_DTYPE.itemsize == 216


class Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("stride", C.c_int32), ("packOrder", C.c_int32)]

    @staticmethod
    def null():
        return Image(None, 0, 0, 0, 0)


class Texture(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("log2width", C.c_uint32), ("log2height", C.c_uint32), ("maxMipLevel", C.c_uint32),
        ("startOffset", C.c_uint32), ("maxLevelMask", C.c_uint32), ("totalPixels", C.c_uint32),
    ]


class OrthoView(C.Structure):
    _fields_ = [("normalToWorldSpace", Matrix3x3), ("screenDepthToLightSpace", Matrix3x3), ("lightSpaceToScreenDepth", Matrix3x3)]


class Model(C.Structure):
    _fields_ = [
        ("points", C.c_void_p), ("pointCount", C.c_int32),
        ("polygons", C.c_void_p), ("polygonCount", C.c_int32),
        ("filter", C.c_int32),
        ("diffuse", Texture), ("light", Texture),
        ("minBound", C.c_float * 3), ("maxBound", C.c_float * 3),
    ]


class SpriteDraw(C.Structure):
    _fields_ = [("sourceHeight", Image), ("sourceA", Image), ("sourceB", Image), ("left", C.c_int32), ("top", C.c_int32), ("heightOffset", C.c_float)]


class DirectedLight(C.Structure):
    _fields_ = [("direction", C.c_float * 3), ("intensity", C.c_float), ("colorRgb", C.c_int32 * 3)]


class PointLight(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("radius", C.c_float), ("intensity", C.c_float), ("colorRgb", C.c_int32 * 3), ("shadowCubeMap", Image)]


class HostModel(C.Structure):
    _fields_ = [
        ("points", C.c_void_p), ("pointCount", C.c_int32),
        ("polygons", C.c_void_p), ("polygonCount", C.c_int32),
        ("filter", C.c_int32),
        ("diffusePixels", C.c_void_p), ("diffuseLayout", Texture),
        ("lightPixels", C.c_void_p), ("lightLayout", Texture),
        ("minBound", C.c_float * 3), ("maxBound", C.c_float * 3),
    ]


class OrthoCamera(C.Structure):
    _fields_ = [
        ("id", C.c_int32), ("worldDirection", C.c_int32),
        ("normalToWorldSpace", Matrix3x3),
        ("pixelOffsetPerTileX", C.c_int32 * 2), ("pixelOffsetPerTileZ", C.c_int32 * 2), ("yPixelsPerTile", C.c_int32),
        ("screenDepthToWorldSpace", Matrix3x3), ("worldSpaceToScreenDepth", Matrix3x3),
        ("screenDepthToLightSpace", Matrix3x3), ("lightSpaceToScreenDepth", Matrix3x3),
        ("roundedScreenPixelsToWorldTiles", C.c_float * 4),
    ]

    def light_view(self):
        return OrthoView(self.normalToWorldSpace, self.screenDepthToLightSpace, self.lightSpaceToScreenDepth)


class OrthoSystem(C.Structure):
    _fields_ = [("cameraTilt", C.c_float), ("pixelsPerTile", C.c_int32), ("view", OrthoCamera * 8)]


DENSE_TRIANGLE_DTYPE = np.dtype([(name, np.float32, (3,)) for name in ("colorA", "colorB", "colorC", "posA", "posB", "posC", "normalA", "normalB", "normalC")])
assert DENSE_TRIANGLE_DTYPE.itemsize == 108


class SpriteConfig(C.Structure):
    _fields_ = [
        ("centerX", C.c_int32), ("centerY", C.c_int32), ("frameRows", C.c_int32), ("propertyColumns", C.c_int32),
        ("minBound", C.c_float * 3), ("maxBound", C.c_float * 3),
        ("points", C.c_void_p), ("pointCount", C.c_int32),
        ("triangleIndices", C.c_void_p), ("triangleIndexCount", C.c_int32),
    ]


class BakedSprite(C.Structure):
    _fields_ = [("atlas", Image), ("centerX", C.c_int32), ("centerY", C.c_int32), ("frameRows", C.c_int32), ("propertyColumns", C.c_int32),
                ("minBound", C.c_float * 3), ("maxBound", C.c_float * 3)]


class SpriteInstance(C.Structure):
    _fields_ = [("typeIndex", C.c_int32), ("direction", C.c_int32), ("location", C.c_int32 * 3), ("shadowCasting", C.c_int32), ("userData", C.c_uint64)]


class ModelInstance(C.Structure):
    _fields_ = [("typeIndex", C.c_int32), ("location", Transform3D), ("userData", C.c_uint64)]


(SW_BLOCK_CLEAR, SW_BLOCK_SPRITE, SW_BLOCK_MODEL, SW_COPY_BLOCK, SW_SPRITE, SW_MODEL, SW_LIGHT_CLEAR, SW_LIGHT_DIRECTED,
 SW_SHADOW_CLEAR, SW_SHADOW_SPRITE, SW_SHADOW_MODEL, SW_LIGHT_POINT, SW_BLEND) = range(1, 14)


class SpriteWorldOp(C.Structure):
    _fields_ = [
        ("op", C.c_int32), ("block", C.c_int32), ("typeIndex", C.c_int32), ("frame", C.c_int32),
        ("left", C.c_int32), ("top", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("sourceLeft", C.c_int32), ("sourceTop", C.c_int32),
        ("heightOffset", C.c_float), ("worldOrigin", C.c_float * 2),
        ("transform", Transform3D),
        ("light", C.c_int32), ("flag", C.c_int32),
    ]


def camera_params(perspective, location, width, height, width_slope=1.0, near=0.01, far=1000.0):
    """A camera POD holding only the constructor arguments; derived fields are filled by
    dfpsr_camera_create_* (product), orc_camera_create (oracle) or ref_camera_fill (reference)."""
    c = Camera()
    c.perspective = 1 if perspective else 0
    c.location = location
    c.imageWidth = float(width)
    c.imageHeight = float(height)
    c.widthSlope = float(np.float32(width_slope))
    c.nearClip = float(np.float32(near))
    c.farClip = float(np.float32(far))
    return c


class ImportedPart(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("diffuseName", C.c_char * 64), ("lightName", C.c_char * 64), ("firstPolygon", C.c_int32), ("polygonCount", C.c_int32)]


class ImportedModel(C.Structure):
    _fields_ = [("points", C.POINTER(C.c_float)), ("pointCount", C.c_int32), ("polygons", C.c_void_p), ("polygonCount", C.c_int32),
                ("parts", C.POINTER(ImportedPart)), ("partCount", C.c_int32), ("filter", C.c_int32), ("minBound", C.c_float * 3), ("maxBound", C.c_float * 3)]
