"""Synthetic PLY and DMF1 texts for the importer parity tests (own content; the reference's media files are compared too when present)."""
from dfpsr_b200 import abi

PLY_BASIC = """ply
format ascii 1.0
comment made for the importer tests
element vertex 6
property float x
property float y
property float z
property uchar red
property uchar green
property uchar blue
element face 3
property list uchar uint vertex_indices
end_header
0 0 0 255 0 0
1 0 0.5 0 255 0
1 1 -0.25 0 0 255
0 1 1e-1 128 64 32
2.5 0.125 3 10 20 30
-1.5 -0.333333 7.000001 255 255 255
4 0 1 2 3
3 0 2 4
5 0 1 2 3 4
"""

# CRLF line ends, float colours with alpha, properties the importer ignores (nx, s), an element it ignores (edge), exponents and a ~ sign
PLY_RICH = "\r\n".join([
    "PLY", "Format ASCII 1.0", "element vertex 5",
    "property float x", "property float y", "property float z", "property float nx", "property float red", "property float green",
    "property float blue", "property float alpha", "property float s",
    "element edge 2", "property int vertex1", "property int vertex2",
    "element face 2", "property list uchar int vertex_indices",
    "end_header",
    "0.1 0.2 0.3 0 1 0.5 0.25 0.75 9", "1.5e1 2E-2 ~3.25 0 0.1 0.2 0.3 0.4 9", "-7 8.0625 9.000000001 0 1 1 1 1 9",
    "123456.789 0.000001 -0.5 0 0 0 0 0 9", "3,5 4 5 0 0.9 0.8 0.7 0.6 9",
    "0 1", "1 2",
    "3 4 3 2", "4 0 1 2 3", ""])

DMF_BASIC = """DMF1
FilterType(Alpha)
CullingType(AABB)
<Shape> Name(Shape) ShapeType(8) Radius(1)
	<Point> X(0.15625) Y(0.171875) Z(-0.15625)
<Part> Name(Body)
	Shader[0](M_Diffuse_1Tex)
	Shader[1](M_Shadow_Solid)
	Texture[0](Planks)
	Texture[1](Planks_Normal)
	MinDetailLevel(0) MaxDetailLevel(2)
	<Triangle>
		X[0](0) Y[0](0) Z[0](0) U1[0](0.25) V1[0](0.5) CR[0](1) CG[0](0.5) CB[0](0.25) CA[0](1)
		X[1](1) Y[1](0) Z[1](0) U1[1](1) V1[1](0)
		X[2](1) Y[2](1) Z[2](0.000001) U1[2](1) V1[2](1) U2[2](0.125) V2[2](0.875)
	<Triangle>
		X[0](0) Y[0](0) Z[0](0.000004)
		X[1](1) Y[1](1) Z[1](0)
		X[2](0) Y[2](1) Z[2](0) CA[2](0.5)
<Part> Name(Lid)
	Shader[0](M_Diffuse_2Tex) Texture[0](Metal) Texture[1](LightMap)
	MinDetailLevel(1) MaxDetailLevel(1)
	<Triangle> X[0](0) Y[0](2) Z[0](0) X[1](1) Y[1](2) Z[1](0) X[2](0.5) Y[2](3) Z[2](-1.5e0)
<Part> Name(Plain)
	Shader[0](M_Diffuse_0Tex)
	<Triangle> X[0](5) Y[0](5) Z[0](5) X[1](6) Y[1](5) Z[1](5) X[2](5) Y[2](6) Z[2](5)
<Bone> Name(Root) X(1)
"""

AXIS = abi.Transform3D()
AXIS.position[:] = [0.5, -1.0, 2.0]
AXIS.xAxis[:] = [0.0, 0.0, 1.0]
AXIS.yAxis[:] = [0.0, 2.0, 0.0]
AXIS.zAxis[:] = [-1.0, 0.0, 0.0]

CASES = [
    ("ply_basic", "ply", PLY_BASIC, {"flip_x": False}),
    ("ply_basic_flipped", "ply", PLY_BASIC, {"flip_x": True}),
    ("ply_rich_axis", "ply", PLY_RICH, {"flip_x": False, "axis": AXIS}),
    ("ply_rich_flipped_axis", "ply", PLY_RICH, {"flip_x": True, "axis": AXIS}),
    ("dmf_detail2", "dmf1", DMF_BASIC, {"detail_level": 2}),
    ("dmf_detail1", "dmf1", DMF_BASIC, {"detail_level": 1}),
    ("dmf_detail0", "dmf1", DMF_BASIC, {"detail_level": 0}),
]
