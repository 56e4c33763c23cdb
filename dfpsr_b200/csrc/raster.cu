// raster.cu — the triangle pipeline on sm_100a.
//
//   project_kernel      ref: api/modelAPI.cpp:238-242 + implementation/render/Camera.h:157-190
//   setup_kernel<false> ref: implementation/render/renderCore.cpp:172-341 (cull, clip, back-face) — counting pass
//   scan_kernel         prefix sums that replace List<TriangleDrawCommand>::push (renderCore.cpp:445) and
//                       CommandQueue::execute's 12 strips (renderCore.cpp:449-480) with per-tile lists
//   setup_kernel<true>  emits compact draw commands in submission order + their row intervals
//                       (implementation/render/ITriangle2D.cpp:31-176) + interpolation planes (:182-300)
//   big_rows_kernel     row intervals of tall triangles, one warp per triangle
//   raster_kernel       ref: shader/fillerTemplates.h:108-441 + shader/RgbaMultiply.h:37-175 + api/textureAPI.h:253-495
//                       one CTA per 32x32 screen tile, one thread per aligned 2x2 quad, colour and depth of the
//                       tile held in registers for the whole triangle list, triangles applied in submission order.
//
// Exactness: coverage is the reference's int64 row-interval arithmetic; interpolated (1/W, U/W, V/W) replay the
// reference's chain of float additions from each row pair's outer block start, so colour and depth are
// bit-identical to the reference's scalar build (exact 1/x) — not merely within tolerance.
#include "common.cuh"

#include <vector>
#include <new>

namespace dfpsr {

static const int TILE = 32;              // screen tile edge in pixels
static const int CHUNK = 16;             // commands staged in shared memory per round
static const int SMALL_ROWS = 16;        // triangles up to this many rows are scan-converted by their set-up thread
static const int SETUP_THREADS = 256;

struct PPoint { // == dfpsr_projected_point
	float csx, csy, csz, isx, isy;
	int32_t pad;
	long long fx, fy;
};
static_assert(sizeof(PPoint) == 40, "PPoint layout");
static_assert(sizeof(dfpsr_projected_point) == 40, "dfpsr_projected_point layout");

// One draw command = one front-facing triangle after culling and clipping (ref: renderCore.h:52-70 carries 456 bytes).
struct Cmd {
	float start[3], dx[3], dy[3]; // Projection (ref: ITriangle2D.h:63-76)
	int32_t bx0, bx1;             // clipped pixel bound, columns
	int32_t rowStart, rowCount;   // even-aligned rows (ref: ITriangle2D.cpp:70-75)
	uint32_t rowOffset;           // first entry in the row-interval table
	uint32_t flags;               // CMD_* | diffuse index << 8 | light index << 20
	float red[3], green[3], blue[3], alpha[3]; // scaled vertex colours (ref: RgbaMultiply.h:45-60)
	float u1[3], v1[3], u2[3], v2[3];
	uint32_t pad_;
};
static_assert(sizeof(Cmd) == 160, "Cmd layout");

enum : uint32_t {
	CMD_AFFINE = 1u, CMD_ALPHA = 2u, CMD_HAS_DIFFUSE = 4u, CMD_HAS_LIGHT = 8u, CMD_HAS_FADE = 16u, CMD_COLORLESS = 32u
};

struct BigCmd {
	uint32_t cmdIndex, pad_;
	long long fx[3], fy[3];
};

struct TaskParams {
	const float *points;
	const dfpsr_polygon *polygons;
	const dfpsr_triangle *triangles; // alternative source: pre-projected triangles
	PPoint *projected;
	int32_t pointCount, polygonCount, triangleCount;
	int32_t slotBase, slotCount, blockBase;
	dfpsr_transform3d modelToWorld;
	dfpsr_camera camera;
	int32_t filter, diffuseIndex, lightIndex; // texture table indices or -1
	int32_t width, height, depthOnly;
};

struct FrameDev {
	uint32_t *slotCounts;  // per slot: command count | rows << 3
	uint32_t *blockCmds, *blockRows; // per set-up block: totals, then exclusive offsets after scan_kernel
	uint32_t *tileCount, *tileOffset, *tileCursor;
	uint32_t *totals;      // [0] commands, [1] rows, [2] tile entries, [3] max entries in one tile, [4] big commands
	Cmd *cmds;
	int2 *rows;
	uint32_t *tileList;
	BigCmd *big;
	int32_t tilesX, tilesY, blockCount;
};

// ------------------------------------------------------------------------------------------------ projection

// The reference's int64_t(float) is x86 cvttss2si: out-of-range and NaN give INT64_MIN.
__device__ __forceinline__ long long float_to_i64(float v) {
	if (!(fabsf(v) < 9.2233720368547758e18f)) { return (long long)0x8000000000000000ull; }
	return __float2ll_rz(v);
}

// ref: implementation/render/Camera.h:160-187
__device__ __forceinline__ PPoint camera_to_screen(const dfpsr_camera &c, float x, float y, float z) {
	PPoint r;
	r.csx = x; r.csy = y; r.csz = z; r.pad = 0;
	if (c.perspective) {
		float invDepth = z > 0.0f ? 1.0f / z : 0.0f;
		float centerShear = z * 0.5f;
		float preX = (x * c.invWidthSlope + centerShear) * c.imageWidth;
		float preY = (-y * c.invHeightSlope + centerShear) * c.imageHeight;
		r.isx = preX * invDepth;
		r.isy = preY * invDepth;
	} else {
		r.isx = (x * c.invWidthSlope + 0.5f) * c.imageWidth;
		r.isy = (-y * c.invHeightSlope + 0.5f) * c.imageHeight;
	}
	r.fx = float_to_i64(r.isx * 256.0f);
	r.fy = float_to_i64(r.isy * 256.0f);
	return r;
}

// ref: math/Transform3D.h:41-52, math/FMatrix3x3.h:52-70
__device__ __forceinline__ PPoint world_to_screen(const dfpsr_camera &c, const dfpsr_transform3d &m, float px, float py, float pz) {
	float wx = (px * m.xAxis[0] + py * m.yAxis[0] + pz * m.zAxis[0]) + m.position[0];
	float wy = (px * m.xAxis[1] + py * m.yAxis[1] + pz * m.zAxis[1]) + m.position[1];
	float wz = (px * m.xAxis[2] + py * m.yAxis[2] + pz * m.zAxis[2]) + m.position[2];
	const dfpsr_transform3d &l = c.location;
	float dx = wx - l.position[0], dy = wy - l.position[1], dz = wz - l.position[2];
	float cx = dx * l.xAxis[0] + dy * l.xAxis[1] + dz * l.xAxis[2];
	float cy = dx * l.yAxis[0] + dy * l.yAxis[1] + dz * l.yAxis[2];
	float cz = dx * l.zAxis[0] + dy * l.zAxis[1] + dz * l.zAxis[2];
	return camera_to_screen(c, cx, cy, cz);
}

__global__ void __launch_bounds__(256) project_kernel(const float *__restrict__ points, int32_t count, dfpsr_transform3d m2w, dfpsr_camera camera, PPoint *__restrict__ out) {
	for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		out[i] = world_to_screen(camera, m2w, points[3 * i], points[3 * i + 1], points[3 * i + 2]);
	}
}

// ------------------------------------------------------------------------------------------------ set-up helpers

// ref: math/FPlane3D.h:39-45
__device__ __forceinline__ bool plane_outside(const float *pl, const PPoint &p) {
	return !((((pl[0] * p.csx) + (pl[1] * p.csy) + (pl[2] * p.csz)) - pl[3]) <= 0.0f);
}

// ref: implementation/render/renderCore.cpp:172-198. 0 hidden, 1 full, 2 partial.
__device__ int triangle_visibility(const PPoint *p, const dfpsr_camera &c, bool clipFrustum) {
	int planeCount = clipFrustum ? c.clipPlaneCount : c.cullPlaneCount;
	const float (*planes)[4] = clipFrustum ? c.clipPlanes : c.cullPlanes;
	bool any = false;
	for (int s = 0; s < planeCount; s++) {
		bool o0 = plane_outside(planes[s], p[0]), o1 = plane_outside(planes[s], p[1]), o2 = plane_outside(planes[s], p[2]);
		if (o0 && o1 && o2) { return 0; }
		any = any || o0 || o1 || o2;
	}
	return any ? 2 : 1;
}

// ref: implementation/render/ITriangle2D.cpp:55-60
__device__ __forceinline__ bool is_frontfacing(const PPoint *p) {
	return ((p[2].fx - p[0].fx) * (p[1].fy - p[0].fy)) + ((p[2].fy - p[0].fy) * (p[0].fx - p[1].fx)) < 0;
}

struct Bound { int32_t l, t, r, b; bool any; };

// ref: implementation/render/ITriangle2D.cpp:31-43, :62-75 — pixel bound, cut to the target, rows aligned to 2.
__device__ Bound raster_bound(const PPoint *p, int32_t width, int32_t height) {
	int32_t rx0 = (int32_t)((p[0].fx + 128) / 256), ry0 = (int32_t)((p[0].fy + 128) / 256);
	int32_t rx1 = (int32_t)((p[1].fx + 128) / 256), ry1 = (int32_t)((p[1].fy + 128) / 256);
	int32_t rx2 = (int32_t)((p[2].fx + 128) / 256), ry2 = (int32_t)((p[2].fy + 128) / 256);
	int32_t l = min(rx0, min(rx1, rx2)) - 1, t = min(ry0, min(ry1, ry2)) - 1;
	int32_t r = max(rx0, max(rx1, rx2)) + 1, b = max(ry0, max(ry1, ry2)) + 1;
	Bound out;
	out.any = l < width && r > 0 && t < height && b > 0; // IRect::overlaps (math/IRect.h:77)
	if (!out.any) { out.l = out.t = out.r = out.b = 0; return out; }
	out.l = max(l, 0); out.r = min(r, width);
	int32_t top = max(t, 0), bottom = min(b, height);
	out.t = (top / 2) * 2;
	out.b = ((bottom + 1) / 2) * 2;
	return out;
}

// Row intervals of one triangle: the reference's cutConvexEdge (ITriangle2D.cpp:86-150) in closed form per row.
struct EdgeSet {
	long long limit0[3], offsetX[3], offsetY[3], valueOrigin[3];
	int32_t threshold[3];
	int32_t kind[3]; // 0 none, 1 left cut, 2 right cut, 3 horizontal
	int32_t leftBound, rightBound, topBound;
	bool degenerate;
};

__device__ void edges_setup(EdgeSet &e, const long long *fx, const long long *fy, int32_t l, int32_t t, int32_t r) {
	e.leftBound = l; e.rightBound = r; e.topBound = t;
	e.degenerate = (fx[0] == fx[1] && fy[0] == fy[1]) || (fx[1] == fx[2] && fy[1] == fy[2]) || (fx[2] == fx[0] && fy[2] == fy[0]);
	long long originX = 128 + (long long)l * 256, originY = 128 + (long long)t * 256;
#pragma unroll
	for (int i = 0; i < 3; i++) {
		int j = (i + 1) % 3;
		long long sx = fx[i], sy = fy[i], ex = fx[j], ey = fy[j];
		long long threshold = (sx > ex || (sx == ex && sy > ey)) ? -1 : 0;
		long long normalX = ey - sy, normalY = sx - ex;
		e.offsetX[i] = normalX * 256;
		e.offsetY[i] = normalY * 256;
		e.valueOrigin[i] = ((originX - sx) * normalX) + ((originY - sy) * normalY);
		e.threshold[i] = (int32_t)threshold;
		e.limit0[i] = threshold - e.valueOrigin[i] + (e.offsetX[i] * l);
		e.kind[i] = normalX != 0 ? (normalX < 0 ? 1 : 2) : (normalY != 0 ? 3 : 0);
	}
}

__device__ int2 edges_row(const EdgeSet &e, int32_t y) {
	int32_t left = e.leftBound, right = e.rightBound;
	if (e.degenerate) { return make_int2(e.rightBound, e.leftBound); }
	long long dy = (long long)(y - e.topBound);
#pragma unroll
	for (int i = 0; i < 3; i++) {
		if (e.kind[i] == 1) {
			long long limit = e.limit0[i] - e.offsetY[i] * dy;
			int32_t side = min(max(e.leftBound, (int32_t)((limit + 1) / e.offsetX[i] + 1)), e.rightBound);
			left = max(left, side);
		} else if (e.kind[i] == 2) {
			long long limit = e.limit0[i] - e.offsetY[i] * dy;
			int32_t side = min(max(e.leftBound, (int32_t)(limit / e.offsetX[i] + 1)), e.rightBound);
			right = min(right, side);
		} else if (e.kind[i] == 3) {
			long long valueRow = e.valueOrigin[i] + e.offsetY[i] * dy;
			if (valueRow > (long long)e.threshold[i]) { left = e.rightBound; right = e.leftBound; }
		}
	}
	return make_int2(left, right);
}

// ref: implementation/render/ITriangle2D.cpp:182-300
__device__ void get_projection(Cmd &cmd, const PPoint *p, const float *subB, const float *subC, bool perspective) {
	float px[3] = {p[0].isx, p[1].isx, p[2].isx}, py[3] = {p[0].isy, p[1].isy, p[2].isy};
	float offsetX[3], offsetY[3], mult[3], normalX[3], normalY[3], tw[3];
#pragma unroll
	for (int i = 0; i < 3; i++) {
		int j = (i + 1) % 3;
		offsetX[i] = py[j] - py[i];
		offsetY[i] = px[i] - px[j];
	}
#pragma unroll
	for (int i = 0; i < 3; i++) {
		int o = (i + 2) % 3;
		float other = ((px[o] - px[i]) * offsetX[i]) + ((py[o] - py[i]) * offsetY[i]);
		mult[o] = (other == 0.0f) ? 0.0f : 1.0f / other;
	}
#pragma unroll
	for (int i = 0; i < 3; i++) {
		normalX[i] = offsetX[i] * mult[i];
		normalY[i] = offsetY[i] * mult[i];
	}
#pragma unroll
	for (int i = 0; i < 3; i++) {
		int o = (i + 2) % 3;
		tw[o] = px[i] * -normalX[i] + py[i] * -normalY[i];
	}
	float adx[3] = {normalX[1], normalX[2], normalX[0]};
	float ady[3] = {normalY[1], normalY[2], normalY[0]};
	if (!perspective) {
		float W[3] = {p[0].csz, p[1].csz, p[2].csz};
		cmd.start[0] = W[0] * tw[0] + W[1] * tw[1] + W[2] * tw[2];
		cmd.start[1] = tw[0] * subB[0] + tw[1] * subB[1] + tw[2] * subB[2];
		cmd.start[2] = tw[0] * subC[0] + tw[1] * subC[1] + tw[2] * subC[2];
		cmd.dx[0] = W[0] * adx[0] + W[1] * adx[1] + W[2] * adx[2];
		cmd.dx[1] = adx[0] * subB[0] + adx[1] * subB[1] + adx[2] * subB[2];
		cmd.dx[2] = adx[0] * subC[0] + adx[1] * subC[1] + adx[2] * subC[2];
		cmd.dy[0] = W[0] * ady[0] + W[1] * ady[1] + W[2] * ady[2];
		cmd.dy[1] = ady[0] * subB[0] + ady[1] * subB[1] + ady[2] * subB[2];
		cmd.dy[2] = ady[0] * subC[0] + ady[1] * subC[1] + ady[2] * subC[2];
	} else {
		float IW[3] = {1.0f / p[0].csz, 1.0f / p[1].csz, 1.0f / p[2].csz};
		cmd.start[0] = IW[0] * tw[0] + IW[1] * tw[1] + IW[2] * tw[2];
		cmd.start[1] = IW[0] * tw[0] * subB[0] + IW[1] * tw[1] * subB[1] + IW[2] * tw[2] * subB[2];
		cmd.start[2] = IW[0] * tw[0] * subC[0] + IW[1] * tw[1] * subC[1] + IW[2] * tw[2] * subC[2];
		cmd.dx[0] = IW[0] * adx[0] + IW[1] * adx[1] + IW[2] * adx[2];
		cmd.dx[1] = IW[0] * adx[0] * subB[0] + IW[1] * adx[1] * subB[1] + IW[2] * adx[2] * subB[2];
		cmd.dx[2] = IW[0] * adx[0] * subC[0] + IW[1] * adx[1] * subC[1] + IW[2] * adx[2] * subC[2];
		cmd.dy[0] = IW[0] * ady[0] + IW[1] * ady[1] + IW[2] * ady[2];
		cmd.dy[1] = IW[0] * ady[0] * subB[0] + IW[1] * ady[1] * subB[1] + IW[2] * ady[2] * subB[2];
		cmd.dy[2] = IW[0] * ady[0] * subC[0] + IW[1] * ady[1] * subC[1] + IW[2] * ady[2] * subC[2];
	}
}

__device__ __forceinline__ bool almost_zero(float v) { return v > -0.001f && v < 0.001f; } // ref: fillerTemplates.h:37
__device__ __forceinline__ bool almost_one(float v) { return v > 0.999f && v < 1.001f; }
__device__ __forceinline__ bool almost_same3(const float *c) { return almost_zero(c[0] - c[1]) && almost_zero(c[0] - c[2]) && almost_zero(c[1] - c[2]); }
__device__ __forceinline__ bool almost_one3(const float *c) { return almost_one(c[0]) && almost_one(c[1]) && almost_one(c[2]); }

// ---- clipping (ref: implementation/render/renderCore.cpp:33-170)

struct SubVertex { float x, y, z, subB, subC, value; int state; };

__device__ __forceinline__ float inverse_lerp(float a, float b, float value) {
	float c = b - a;
	return c == 0.0f ? 0.5f : (value - a) / c;
}

__device__ SubVertex sub_lerp(const SubVertex &a, const SubVertex &b, float ratio) {
	SubVertex r;
	float inv = 1.0f - ratio;
	r.x = a.x * inv + b.x * ratio; r.y = a.y * inv + b.y * ratio; r.z = a.z * inv + b.z * ratio;
	r.subB = a.subB * inv + b.subB * ratio;
	r.subC = a.subC * inv + b.subC * ratio;
	r.state = 0; r.value = 0.0f;
	return r;
}

__device__ void clip_plane(SubVertex *v, int &count, const float *pl) {
	const int maxPoints = 9;
	if (!(count >= 3 && count < maxPoints)) { return; }
	int outsideCount = 0, lastOutside = 0;
	for (int i = 0; i < count; i++) {
		float distance = ((pl[0] * v[i].x) + (pl[1] * v[i].y) + (pl[2] * v[i].z)) - pl[3];
		v[i].value = distance;
		if (distance > 0.0f) { outsideCount++; lastOutside = i; v[i].state = 1; } else { v[i].state = 0; }
	}
	if (outsideCount == 0) { return; }
	if (outsideCount >= count) { count = 0; return; }
	if (outsideCount == 1) {
		int cur = lastOutside, prev = (lastOutside - 1 + count) % count, next = (lastOutside + 1) % count;
		float r1 = inverse_lerp(v[prev].value, v[cur].value, 0.0f);
		float r2 = inverse_lerp(v[cur].value, v[next].value, 0.0f);
		SubVertex cutStart = sub_lerp(v[prev], v[cur], r1);
		SubVertex cutEnd = sub_lerp(v[cur], v[next], r2);
		v[lastOutside] = cutStart;
		if (count < maxPoints) {
			for (int k = count - 1; k >= next; k--) { v[k + 1] = v[k]; }
			v[next] = cutEnd;
			count++;
		}
	} else {
		for (int cur = 0; cur < count; cur++) {
			int prev = (cur - 1 + count) % count, next = (cur + 1) % count;
			if (v[cur].state == 1) {
				if (v[prev].state == 0) {
					float r = inverse_lerp(v[prev].value, v[cur].value, 0.0f);
					v[cur] = sub_lerp(v[prev], v[cur], r);
					v[cur].state = 2;
				} else if (v[next].state == 0) {
					float r = inverse_lerp(v[cur].value, v[next].value, 0.0f);
					v[cur] = sub_lerp(v[cur], v[next], r);
					v[cur].state = 2;
				}
			}
		}
		if (outsideCount > 2) {
			for (int i = count - 1; i >= 0; i--) {
				if (v[i].state == 1) {
					for (int k = i; k < count - 1; k++) { v[k] = v[k + 1]; }
					count--;
				}
			}
		}
	}
}

// Calls emit(p, subB, subC) once per draw command of one input triangle, in the reference's order.
// ref: implementation/render/renderCore.cpp:261-341 (colour path) and :407-443 (depth-only path).
template <typename Emit>
__device__ void for_each_command(const TaskParams &task, const PPoint *p, const float *alpha, Emit &&emit) {
	const dfpsr_camera &c = task.camera;
	if (triangle_visibility(p, c, false) == 0) { return; }
	if (!task.depthOnly && task.filter == DFPSR_FILTER_ALPHA && almost_zero(alpha[0]) && almost_zero(alpha[1]) && almost_zero(alpha[2])) { return; }
	if (triangle_visibility(p, c, true) == 1) {
		if (is_frontfacing(p)) {
			const float subB[3] = {0.0f, 1.0f, 0.0f}, subC[3] = {0.0f, 0.0f, 1.0f};
			emit(p, subB, subC);
		}
	} else {
		SubVertex v[9];
		v[0] = SubVertex{p[0].csx, p[0].csy, p[0].csz, 0.0f, 0.0f, 0.0f, 0};
		v[1] = SubVertex{p[1].csx, p[1].csy, p[1].csz, 1.0f, 0.0f, 0.0f, 0};
		v[2] = SubVertex{p[2].csx, p[2].csy, p[2].csz, 0.0f, 1.0f, 0.0f, 0};
		int count = 3;
		for (int s = 0; s < c.clipPlaneCount; s++) { clip_plane(v, count, c.clipPlanes[s]); }
		for (int i = 0; i < count - 2; i++) {
			PPoint q[3] = {camera_to_screen(c, v[0].x, v[0].y, v[0].z), camera_to_screen(c, v[1 + i].x, v[1 + i].y, v[1 + i].z), camera_to_screen(c, v[2 + i].x, v[2 + i].y, v[2 + i].z)};
			if (is_frontfacing(q)) {
				float subB[3], subC[3];
				if (task.depthOnly) { // ref: renderCore.cpp:350 getProjection(FVector3D(), FVector3D(), ...)
					subB[0] = subB[1] = subB[2] = 0.0f; subC[0] = subC[1] = subC[2] = 0.0f;
				} else {
					subB[0] = v[0].subB; subB[1] = v[1 + i].subB; subB[2] = v[2 + i].subB;
					subC[0] = v[0].subC; subC[1] = v[1 + i].subC; subC[2] = v[2 + i].subC;
				}
				emit(q, subB, subC);
			}
		}
	}
}

// Loads the three corners of input triangle `slot` of a task. Returns false for an empty slot.
__device__ bool load_triangle(const TaskParams &task, int32_t local, PPoint *p, float colors[3][4], float tex[3][4]) {
	if (task.triangles != nullptr) {
		const dfpsr_triangle &t = task.triangles[local];
#pragma unroll
		for (int k = 0; k < 3; k++) {
			p[k] = *(const PPoint *)&t.pos[k];
#pragma unroll
			for (int ch = 0; ch < 4; ch++) { colors[k][ch] = t.colors[k][ch]; tex[k][ch] = t.texCoords[k][ch]; }
		}
		return true;
	}
	const dfpsr_polygon &poly = task.polygons[local >> 1];
	int second = local & 1;
	if (second && poly.pointIndices[3] == -1) { return false; }
	int corner[3] = {0, 1 + second, 2 + second}; // ref: api/modelAPI.cpp:252-278 fan (0,1,2), (0,2,3)
#pragma unroll
	for (int k = 0; k < 3; k++) {
		p[k] = task.projected[poly.pointIndices[corner[k]]];
#pragma unroll
		for (int ch = 0; ch < 4; ch++) { colors[k][ch] = poly.colors[corner[k]][ch]; tex[k][ch] = poly.texCoords[corner[k]][ch]; }
	}
	return true;
}

// ------------------------------------------------------------------------------------------------ set-up kernel

template <bool EMIT>
__global__ void __launch_bounds__(SETUP_THREADS) setup_kernel(TaskParams task, FrameDev frame) {
	__shared__ uint32_t warpCmds[SETUP_THREADS / 32], warpRows[SETUP_THREADS / 32];
	int32_t local = blockIdx.x * SETUP_THREADS + threadIdx.x;
	bool active = local < task.slotCount;
	int32_t slot = task.slotBase + local;
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	PPoint p[3];
	float colors[3][4], tex[3][4];
	bool loaded = active && load_triangle(task, local, p, colors, tex);
	const int32_t tilesX = frame.tilesX;

	uint32_t cmdBase = 0, rowBase = 0;
	if (EMIT) {
		// exclusive prefix of (commands, rows) inside the block, on top of the block's scanned offset
		uint32_t packed = active ? frame.slotCounts[slot] : 0u;
		uint32_t nCmd = packed & 7u, nRows = packed >> 3;
		uint32_t incCmd = nCmd, incRows = nRows;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t a = __shfl_up_sync(0xffffffffu, incCmd, d), b = __shfl_up_sync(0xffffffffu, incRows, d);
			if (lane >= d) { incCmd += a; incRows += b; }
		}
		if (lane == 31) { warpCmds[warp] = incCmd; warpRows[warp] = incRows; }
		__syncthreads();
		uint32_t preCmd = 0, preRows = 0;
		for (int w = 0; w < warp; w++) { preCmd += warpCmds[w]; preRows += warpRows[w]; }
		cmdBase = frame.blockCmds[task.blockBase + blockIdx.x] + preCmd + incCmd - nCmd;
		rowBase = frame.blockRows[task.blockBase + blockIdx.x] + preRows + incRows - nRows;
	}

	uint32_t countCmd = 0, countRows = 0;
	if (loaded) {
		float alpha[3] = {colors[0][3], colors[1][3], colors[2][3]};
		for_each_command(task, p, alpha, [&](const PPoint *q, const float *subB, const float *subC) {
			Bound bound = raster_bound(q, task.width, task.height);
			int32_t rowCount = bound.any ? bound.b - bound.t : 0;
			int32_t tx0 = bound.l / TILE, tx1 = (bound.r - 1) / TILE, ty0 = bound.t / TILE, ty1 = (min(bound.b, task.height) - 1) / TILE;
			if (!EMIT) {
				if (rowCount > 0) {
					for (int32_t ty = ty0; ty <= ty1; ty++) {
						for (int32_t tx = tx0; tx <= tx1; tx++) { atomicAdd(&frame.tileCount[ty * tilesX + tx], 1u); }
					}
				}
			} else {
				uint32_t index = cmdBase + countCmd;
				Cmd cmd;
				bool perspective = task.camera.perspective != 0;
				get_projection(cmd, q, subB, subC, perspective);
				cmd.bx0 = bound.l; cmd.bx1 = bound.r;
				cmd.rowStart = bound.t; cmd.rowCount = rowCount;
				cmd.rowOffset = rowBase + countRows;
				uint32_t flags = perspective ? 0u : CMD_AFFINE;
				if (task.filter == DFPSR_FILTER_ALPHA) { flags |= CMD_ALPHA; }
				// ref: shader/RgbaMultiply.h:45-60, :110-116
				float scale = 255.0f;
				if (task.diffuseIndex >= 0) { scale *= 1.0f / 255.0f; flags |= CMD_HAS_DIFFUSE | ((uint32_t)task.diffuseIndex << 8); }
				if (task.lightIndex >= 0) { scale *= 1.0f / 255.0f; flags |= CMD_HAS_LIGHT | ((uint32_t)task.lightIndex << 20); }
#pragma unroll
				for (int k = 0; k < 3; k++) {
					cmd.red[k] = colors[k][0] * scale; cmd.green[k] = colors[k][1] * scale;
					cmd.blue[k] = colors[k][2] * scale; cmd.alpha[k] = colors[k][3] * scale;
					cmd.u1[k] = tex[k][0]; cmd.v1[k] = tex[k][1]; cmd.u2[k] = tex[k][2]; cmd.v2[k] = tex[k][3];
				}
				if (!(almost_same3(cmd.red) && almost_same3(cmd.green) && almost_same3(cmd.blue) && almost_same3(cmd.alpha))) { flags |= CMD_HAS_FADE; }
				if (almost_one3(cmd.red) && almost_one3(cmd.green) && almost_one3(cmd.blue) && almost_one3(cmd.alpha)) { flags |= CMD_COLORLESS; }
				cmd.flags = flags;
				cmd.pad_ = 0;
				frame.cmds[index] = cmd;
				if (rowCount > 0) {
					long long fx[3] = {q[0].fx, q[1].fx, q[2].fx}, fy[3] = {q[0].fy, q[1].fy, q[2].fy};
					if (rowCount <= SMALL_ROWS) {
						EdgeSet edges;
						edges_setup(edges, fx, fy, bound.l, bound.t, bound.r);
						for (int32_t r = 0; r < rowCount; r++) { frame.rows[cmd.rowOffset + r] = edges_row(edges, bound.t + r); }
					} else {
						uint32_t b = atomicAdd(&frame.totals[4], 1u);
						BigCmd big;
						big.cmdIndex = index; big.pad_ = 0;
						for (int k = 0; k < 3; k++) { big.fx[k] = fx[k]; big.fy[k] = fy[k]; }
						frame.big[b] = big;
					}
					for (int32_t ty = ty0; ty <= ty1; ty++) {
						for (int32_t tx = tx0; tx <= tx1; tx++) {
							int32_t tile = ty * tilesX + tx;
							uint32_t pos = atomicAdd(&frame.tileCursor[tile], 1u);
							frame.tileList[frame.tileOffset[tile] + pos] = index;
						}
					}
				}
			}
			countCmd++;
			countRows += (uint32_t)rowCount;
		});
	}

	if (!EMIT) {
		if (active) { frame.slotCounts[slot] = countCmd | (countRows << 3); }
		// block totals for scan_kernel
		uint32_t sumCmd = countCmd, sumRows = countRows;
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			sumCmd += __shfl_xor_sync(0xffffffffu, sumCmd, d);
			sumRows += __shfl_xor_sync(0xffffffffu, sumRows, d);
		}
		if (lane == 0) { warpCmds[warp] = sumCmd; warpRows[warp] = sumRows; }
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t a = 0, b = 0;
			for (int w = 0; w < SETUP_THREADS / 32; w++) { a += warpCmds[w]; b += warpRows[w]; }
			frame.blockCmds[task.blockBase + blockIdx.x] = a;
			frame.blockRows[task.blockBase + blockIdx.x] = b;
		}
	}
}

// One CTA: exclusive scans of the per-block command/row totals and of the per-tile entry counts.
__global__ void __launch_bounds__(1024) scan_kernel(FrameDev frame) {
	__shared__ uint32_t warpSum[3][32];
	__shared__ uint32_t carry[3];
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x < 3) { carry[threadIdx.x] = 0; }
	__syncthreads();
	uint32_t maxTile = 0;
	int32_t tileTotal = frame.tilesX * frame.tilesY;
	int32_t longest = max(frame.blockCount, tileTotal);
	for (int32_t base = 0; base < longest; base += 1024) {
		int32_t i = base + threadIdx.x;
		uint32_t v[3];
		v[0] = i < frame.blockCount ? frame.blockCmds[i] : 0u;
		v[1] = i < frame.blockCount ? frame.blockRows[i] : 0u;
		v[2] = i < tileTotal ? frame.tileCount[i] : 0u;
		maxTile = max(maxTile, v[2]);
		uint32_t inc[3] = {v[0], v[1], v[2]};
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
			for (int k = 0; k < 3; k++) {
				uint32_t a = __shfl_up_sync(0xffffffffu, inc[k], d);
				if (lane >= d) { inc[k] += a; }
			}
		}
		if (lane == 31) { for (int k = 0; k < 3; k++) { warpSum[k][warp] = inc[k]; } }
		__syncthreads();
		uint32_t pre[3] = {carry[0], carry[1], carry[2]};
		for (int w = 0; w < warp; w++) { for (int k = 0; k < 3; k++) { pre[k] += warpSum[k][w]; } }
		if (i < frame.blockCount) {
			frame.blockCmds[i] = pre[0] + inc[0] - v[0];
			frame.blockRows[i] = pre[1] + inc[1] - v[1];
		}
		if (i < tileTotal) {
			frame.tileOffset[i] = pre[2] + inc[2] - v[2];
			frame.tileCursor[i] = 0;
		}
		__syncthreads();
		if (threadIdx.x == 1023) { for (int k = 0; k < 3; k++) { carry[k] = pre[k] + inc[k]; } }
		__syncthreads();
	}
	// block-wide max of tile counts
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) { maxTile = max(maxTile, __shfl_xor_sync(0xffffffffu, maxTile, d)); }
	if (lane == 0) { warpSum[0][warp] = maxTile; }
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t m = 0;
		for (int w = 0; w < 32; w++) { m = max(m, warpSum[0][w]); }
		frame.totals[0] = carry[0];
		frame.totals[1] = carry[1];
		frame.totals[2] = carry[2];
		frame.totals[3] = m;
		frame.totals[4] = 0;
		frame.tileOffset[tileTotal] = carry[2];
	}
}

// Row intervals of the triangles taller than SMALL_ROWS: one warp per triangle, lanes stride over rows.
__global__ void __launch_bounds__(256) big_rows_kernel(FrameDev frame) {
	uint32_t count = frame.totals[4];
	int lane = threadIdx.x & 31;
	for (uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5); b < count; b += gridDim.x * 8) {
		const BigCmd &big = frame.big[b];
		const Cmd &cmd = frame.cmds[big.cmdIndex];
		EdgeSet edges;
		long long fx[3] = {big.fx[0], big.fx[1], big.fx[2]}, fy[3] = {big.fy[0], big.fy[1], big.fy[2]};
		edges_setup(edges, fx, fy, cmd.bx0, cmd.rowStart, cmd.bx1);
		for (int32_t r = lane; r < cmd.rowCount; r += 32) { frame.rows[cmd.rowOffset + r] = edges_row(edges, cmd.rowStart + r); }
	}
}

// ------------------------------------------------------------------------------------------------ texture sampling

struct TexDev {
	const uint32_t *data;
	uint32_t log2width, log2height, maxMipLevel, startOffset, maxLevelMask;
};
static const int MAX_TEXTURES = 64;
struct TexTable { TexDev t[MAX_TEXTURES]; };

// ref: api/textureAPI.h:253-263 weightColors on 16-bit lane pairs (sums never exceed 255 * 256, so lanes cannot carry)
__device__ __forceinline__ uint32_t weight_colors(uint32_t colorA, uint32_t weightA, uint32_t colorB, uint32_t weightB) {
	uint32_t low = (colorA & 0x00FF00FFu) * weightA + (colorB & 0x00FF00FFu) * weightB;
	uint32_t high = ((colorA >> 8) & 0x00FF00FFu) * weightA + ((colorB >> 8) & 0x00FF00FFu) * weightB;
	return ((low >> 8) & 0x00FF00FFu) | (high & 0xFF00FF00u);
}

// ref: api/textureAPI.h:342-438 texture_sample_bilinear<SQUARE=false, *, MIP_INSIDE=true, *>
__device__ __forceinline__ uint32_t sample_bilinear(const TexDev &t, float u, float v, uint32_t mip) {
	uint32_t scaleU = (256u << t.log2width) >> mip, scaleV = (256u << t.log2height) >> mip;
	uint32_t subX = __float2uint_rz((u + 256.0f) * (float)scaleU) - 128u;
	uint32_t subY = __float2uint_rz((v + 256.0f) * (float)scaleV) - 128u;
	uint32_t wx = subX & 0xFFu, wy = subY & 0xFFu;
	uint32_t left = subX >> 8, top = subY >> 8;
	uint32_t maskX = ((1u << t.log2width) - 1u) >> mip, maskY = ((1u << t.log2height) - 1u) >> mip;
	uint32_t right = (left + 1u) & maskX, bottom = (top + 1u) & maskY;
	left &= maskX; top &= maskY;
	uint32_t log2Stride = t.log2width - mip;
	const uint32_t *data = t.data + (t.startOffset & (t.maxLevelMask >> (2u * mip))); // ref: api/textureAPI.h:79-85
	uint32_t c00 = __ldg(data + ((top << log2Stride) | left)), c10 = __ldg(data + ((top << log2Stride) | right));
	uint32_t c01 = __ldg(data + ((bottom << log2Stride) | left)), c11 = __ldg(data + ((bottom << log2Stride) | right));
	uint32_t upper = weight_colors(c00, 256u - wx, c10, wx);
	uint32_t lower = weight_colors(c01, 256u - wx, c11, wx);
	return weight_colors(upper, 256u - wy, lower, wy);
}

// ref: api/textureAPI.h:472-495 — one mip level per quad, from lanes 0, 1, 2 (covered or not)
__device__ __forceinline__ uint32_t mip_level(const TexDev &t, const float *u, const float *v) {
	float offsetU = fmaxf(fabsf(u[0] - u[1]), fabsf(u[0] - u[2])) * (float)(1u << t.log2width);
	float offsetV = fmaxf(fabsf(v[0] - v[1]), fabsf(v[0] - v[2])) * (float)(1u << t.log2height);
	float offset = fmaxf(offsetU, offsetV);
	uint32_t result = 0;
	if (offset > 2.0f) { result = 1; }
	if (offset > 4.0f) { result = 2; }
	if (offset > 8.0f) { result = 3; }
	if (offset > 16.0f) { result = 4; }
	return min(result, t.maxMipLevel);
}

// ref: shader/shaderMethods.h:39-44
__device__ __forceinline__ float interpolate3(const float *d, float wa, float wb, float wc) {
	return d[0] * wa + d[1] * wb + d[2] * wc;
}

// Samples one texture for the four lanes of a quad and multiplies (or assigns) into rgba[lane][channel].
template <bool MULTIPLY>
__device__ __forceinline__ void sample_quad(const TexDev &t, bool highestResolution, const float *cu, const float *cv, const float *wa, const float *wb, const float *wc, float rgba[4][4]) {
	float u[4], v[4];
#pragma unroll
	for (int l = 0; l < 4; l++) { u[l] = interpolate3(cu, wa[l], wb[l], wc[l]); v[l] = interpolate3(cv, wa[l], wb[l], wc[l]); }
	uint32_t mip = highestResolution ? 0u : mip_level(t, u, v);
#pragma unroll
	for (int l = 0; l < 4; l++) {
		uint32_t c = sample_bilinear(t, u[l], v[l], mip);
		float r = (float)(c & 255u), g = (float)((c >> 8) & 255u), b = (float)((c >> 16) & 255u), a = (float)(c >> 24);
		if (MULTIPLY) { rgba[l][0] = rgba[l][0] * r; rgba[l][1] = rgba[l][1] * g; rgba[l][2] = rgba[l][2] * b; rgba[l][3] = rgba[l][3] * a; }
		else { rgba[l][0] = r; rgba[l][1] = g; rgba[l][2] = b; rgba[l][3] = a; }
	}
}

// ------------------------------------------------------------------------------------------------ tile kernel

struct RasterParams {
	dfpsr_image color, depth; // data == nullptr when absent
	int32_t width, height;
	int32_t depthOnly;        // model_renderDepth semantics (1x1 aligned, per-pixel chain)
	int32_t clear;            // targets are defined to be (clearColor, clearDepth) before this frame: no loads, every pixel stored
	uint32_t clearColor;
	float clearDepth;
	uint32_t sortCapacity;    // entries of shared memory available for sorting a tile's list (power of two)
};

// Bitonic sort of s[0..capacity) ascending, capacity a power of two, 256 threads.
__device__ void sort_shared(uint32_t *s, uint32_t capacity) {
	for (uint32_t k = 2; k <= capacity; k <<= 1) {
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t i = threadIdx.x; i < capacity; i += blockDim.x) {
				uint32_t l = i ^ j;
				if (l > i) {
					uint32_t a = s[i], b = s[l];
					bool ascending = (i & k) == 0;
					if ((a > b) == ascending) { s[i] = b; s[l] = a; }
				}
			}
			__syncthreads();
		}
	}
}

__global__ void __launch_bounds__(256, 2) raster_kernel(FrameDev frame, RasterParams rp, TexTable textures) {
	extern __shared__ __align__(16) unsigned char smemRaw[];
	Cmd *sCmd = (Cmd *)smemRaw;                                   // CHUNK commands
	int2 *sRows = (int2 *)(smemRaw + sizeof(Cmd) * CHUNK);        // CHUNK x TILE row intervals
	uint32_t *sList = (uint32_t *)(smemRaw + sizeof(Cmd) * CHUNK + sizeof(int2) * CHUNK * TILE);
	__shared__ uint32_t sCount;

	const int32_t tile = blockIdx.x;
	const int32_t tileX = tile % frame.tilesX, tileY = tile / frame.tilesX;
	const uint32_t n = frame.tileCount[tile];
	if (n == 0 && !rp.clear) { return; }
	const uint32_t *list = frame.tileList + frame.tileOffset[tile];

	const int32_t qx = threadIdx.x & 15, qy = threadIdx.x >> 4;
	const int32_t x0 = tileX * TILE + 2 * qx, y1 = tileY * TILE + 2 * qy, y2 = y1 + 1;
	const bool hasColor = rp.color.data != nullptr, hasDepth = rp.depth.data != nullptr;
	const bool in[4] = {x0 < rp.width && y1 < rp.height, x0 + 1 < rp.width && y1 < rp.height, x0 < rp.width && y2 < rp.height, x0 + 1 < rp.width && y2 < rp.height};
	const uint32_t shifts = pack_shifts(rp.color.packOrder);

	uint32_t col[4];
	float dep[4];
#pragma unroll
	for (int l = 0; l < 4; l++) {
		int32_t px = x0 + (l & 1), py = y1 + (l >> 1);
		col[l] = rp.clearColor; dep[l] = rp.clearDepth;
		if (!rp.clear && in[l]) {
			if (hasColor) { col[l] = row_ptr<uint32_t>(rp.color.data, rp.color.stride, py)[px]; }
			if (hasDepth) { dep[l] = row_ptr<float>(rp.depth.data, rp.depth.stride, py)[px]; }
		}
	}
	bool dirty = rp.clear != 0;

	// The tile's list is consumed in ascending command order in windows of at most sortCapacity entries.
	uint32_t processedBelow = 0; // every entry < processedBelow has been applied
	uint32_t remaining = n;
	while (remaining > 0) {
		uint32_t windowEnd = 0xFFFFFFFFu; // exclusive upper key of this window
		uint32_t windowCount = remaining;
		if (remaining > rp.sortCapacity) {
			// Binary search the largest key bound whose window still fits in shared memory.
			uint32_t lo = processedBelow, hi = 0xFFFFFFFFu; // count(lo) fits, count(hi) does not
			// invariant: entries in [processedBelow, lo) <= capacity; [processedBelow, hi) > capacity
			while (hi - lo > 1) {
				uint32_t mid = lo + (hi - lo) / 2;
				if (threadIdx.x == 0) { sCount = 0; }
				__syncthreads();
				uint32_t mine = 0;
				for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { uint32_t e = list[i]; mine += (e >= processedBelow && e < mid) ? 1u : 0u; }
				atomicAdd(&sCount, mine);
				__syncthreads();
				uint32_t c = sCount;
				__syncthreads();
				if (c <= rp.sortCapacity) { lo = mid; } else { hi = mid; }
			}
			windowEnd = lo;
			if (threadIdx.x == 0) { sCount = 0; }
			__syncthreads();
			for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
				uint32_t e = list[i];
				if (e >= processedBelow && e < windowEnd) { sList[atomicAdd(&sCount, 1u)] = e; }
			}
			__syncthreads();
			windowCount = sCount;
			__syncthreads();
		} else {
			if (threadIdx.x == 0) { sCount = 0; }
			__syncthreads();
			for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
				uint32_t e = list[i];
				if (e >= processedBelow) { sList[atomicAdd(&sCount, 1u)] = e; }
			}
			__syncthreads();
			windowCount = sCount;
			__syncthreads();
		}
		// pad to a power of two and sort
		uint32_t sortSize = 1;
		while (sortSize < windowCount) { sortSize <<= 1; }
		for (uint32_t i = windowCount + threadIdx.x; i < sortSize; i += blockDim.x) { sList[i] = 0xFFFFFFFFu; }
		__syncthreads();
		if (windowCount > 1) { sort_shared(sList, sortSize); }

		for (uint32_t chunkStart = 0; chunkStart < windowCount; chunkStart += CHUNK) {
			const uint32_t chunkCount = min((uint32_t)CHUNK, windowCount - chunkStart);
			// stage CHUNK command records and their row intervals for this tile's 32 rows
			for (uint32_t w = threadIdx.x; w < chunkCount * (sizeof(Cmd) / 16); w += blockDim.x) {
				uint32_t c = w / (sizeof(Cmd) / 16), part = w % (sizeof(Cmd) / 16);
				((uint4 *)&sCmd[c])[part] = __ldg(((const uint4 *)&frame.cmds[sList[chunkStart + c]]) + part);
			}
			__syncthreads();
			for (uint32_t w = threadIdx.x; w < chunkCount * TILE; w += blockDim.x) {
				uint32_t c = w / TILE, r = w % TILE;
				int32_t idx = tileY * TILE + (int32_t)r - sCmd[c].rowStart;
				int2 row = make_int2(0, 0);
				if (idx >= 0 && idx < sCmd[c].rowCount) { row = frame.rows[sCmd[c].rowOffset + idx]; }
				sRows[c * TILE + r] = row;
			}
			__syncthreads();

			for (uint32_t c = 0; c < chunkCount; c++) {
				const Cmd &cmd = sCmd[c];
				int2 upperRow = sRows[c * TILE + 2 * qy], lowerRow = sRows[c * TILE + 2 * qy + 1];
				const uint32_t flags = cmd.flags;
				const bool affine = (flags & CMD_AFFINE) != 0;

				if (rp.depthOnly) {
					// ref: implementation/render/renderCore.cpp:343-387 — per row: value at row.left, then += dx per pixel
#pragma unroll
					for (int l = 0; l < 4; l++) {
						int2 row = (l < 2) ? upperRow : lowerRow;
						int32_t px = x0 + (l & 1), py = y1 + (l >> 1);
						if (px >= row.x && px < row.y && py < rp.height) {
							float value = (cmd.start[0] + (cmd.dx[0] * ((float)row.x + 0.5f))) + (cmd.dy[0] * ((float)py + 0.5f));
							for (int32_t s = row.x; s < px; s++) { value += cmd.dx[0]; }
							if (affine ? (value < dep[l]) : (value > dep[l])) { dep[l] = value; dirty = true; }
						}
					}
					continue;
				}

				// ref: shader/fillerTemplates.h:275-300
				int32_t outerStart = min(upperRow.x, lowerRow.x), outerEnd = max(upperRow.y, lowerRow.y);
				int32_t innerStart = max(upperRow.x, lowerRow.x), innerEnd = min(upperRow.y, lowerRow.y);
				int32_t obs = outerStart & ~1, obe = (outerEnd + 1) & ~1, ibs = (innerStart + 1) & ~1, ibe = innerEnd & ~1;
				if (y2 >= rp.height) { lowerRow.y = lowerRow.x; }
				bool hasTop = upperRow.y > upperRow.x, hasBottom = lowerRow.y > lowerRow.x;
				if (!(hasTop || hasBottom)) { continue; }
				if (x0 < obs || x0 >= obe) { continue; }

				// Replay of the reference's running sums (fillerTemplates.h:329-372): start at the outer block start,
				// add 2*dx per quad; inner (unclipped) runs advance all four lanes separately; after an inner run the
				// base jumps by one multiplication.
				float up[3], lo[3], dx2[3];
				{
					float fx = (float)obs + 0.5f, fy = (float)y1 + 0.5f;
#pragma unroll
					for (int k = 0; k < 3; k++) {
						up[k] = (cmd.start[k] + (cmd.dx[k] * fx)) + (cmd.dy[k] * fy);
						lo[k] = up[k] + cmd.dy[k];
						dx2[k] = cmd.dx[k] * 2.0f;
					}
				}
				float lanes[3][4];
				bool clipSides = true;
				const bool noInner = ibe <= ibs;
				if (noInner || x0 < ibs) {
					for (int32_t s = obs; s < x0; s += 2) {
#pragma unroll
						for (int k = 0; k < 3; k++) { up[k] += dx2[k]; lo[k] += dx2[k]; }
					}
#pragma unroll
					for (int k = 0; k < 3; k++) { lanes[k][0] = up[k]; lanes[k][1] = up[k] + cmd.dx[k]; lanes[k][2] = lo[k]; lanes[k][3] = lo[k] + cmd.dx[k]; }
				} else {
					for (int32_t s = obs; s < ibs; s += 2) {
#pragma unroll
						for (int k = 0; k < 3; k++) { up[k] += dx2[k]; lo[k] += dx2[k]; }
					}
					if (x0 < ibe) {
						clipSides = false;
#pragma unroll
						for (int k = 0; k < 3; k++) { lanes[k][0] = up[k]; lanes[k][1] = up[k] + cmd.dx[k]; lanes[k][2] = lo[k]; lanes[k][3] = lo[k] + cmd.dx[k]; }
						for (int32_t s = ibs; s < x0; s += 2) {
#pragma unroll
							for (int k = 0; k < 3; k++) {
#pragma unroll
								for (int l = 0; l < 4; l++) { lanes[k][l] += dx2[k]; }
							}
						}
					} else {
						float quadCount = (float)((ibe - ibs) / 2);
#pragma unroll
						for (int k = 0; k < 3; k++) { up[k] = up[k] + (dx2[k] * quadCount); lo[k] = lo[k] + (dx2[k] * quadCount); }
						for (int32_t s = ibe; s < x0; s += 2) {
#pragma unroll
							for (int k = 0; k < 3; k++) { up[k] += dx2[k]; lo[k] += dx2[k]; }
						}
#pragma unroll
						for (int k = 0; k < 3; k++) { lanes[k][0] = up[k]; lanes[k][1] = up[k] + cmd.dx[k]; lanes[k][2] = lo[k]; lanes[k][3] = lo[k] + cmd.dx[k]; }
					}
				}

				// ref: shader/fillerTemplates.h:196-243 — weights; :93-138 — visibility
				float wa[4], wb[4], wc[4];
				bool vis[4];
				bool anyVisible = false;
#pragma unroll
				for (int l = 0; l < 4; l++) {
					if (affine) { wb[l] = lanes[1][l]; wc[l] = lanes[2][l]; }
					else { float linearDepth = 1.0f / lanes[0][l]; wb[l] = lanes[1][l] * linearDepth; wc[l] = lanes[2][l] * linearDepth; }
					wa[l] = 1.0f - (wb[l] + wc[l]);
					bool visible = true;
					if (clipSides) {
						int2 row = (l < 2) ? upperRow : lowerRow;
						int32_t px = x0 + (l & 1);
						visible = px >= row.x && px < row.y;
					}
					if (visible && hasDepth) { visible = affine ? (lanes[0][l] < dep[l]) : (lanes[0][l] > dep[l]); }
					vis[l] = visible;
					anyVisible = anyVisible || visible;
				}
				if (!anyVisible) { continue; }

				if (hasColor) {
					// ref: shader/RgbaMultiply.h:75-106
					float rgba[4][4];
					const bool hasDiffuse = (flags & CMD_HAS_DIFFUSE) != 0, hasLight = (flags & CMD_HAS_LIGHT) != 0;
					const bool fade = (flags & CMD_HAS_FADE) != 0, colorless = (flags & CMD_COLORLESS) != 0 && !fade;
					if (hasDiffuse && !hasLight && colorless) {
						sample_quad<false>(textures.t[(flags >> 8) & 0xFFFu], false, cmd.u1, cmd.v1, wa, wb, wc, rgba);
					} else if (hasLight && !hasDiffuse && colorless) {
						sample_quad<false>(textures.t[flags >> 20], true, cmd.u2, cmd.v2, wa, wb, wc, rgba);
					} else {
#pragma unroll
						for (int l = 0; l < 4; l++) {
							if (fade) {
								rgba[l][0] = interpolate3(cmd.red, wa[l], wb[l], wc[l]);
								rgba[l][1] = interpolate3(cmd.green, wa[l], wb[l], wc[l]);
								rgba[l][2] = interpolate3(cmd.blue, wa[l], wb[l], wc[l]);
								rgba[l][3] = interpolate3(cmd.alpha, wa[l], wb[l], wc[l]);
							} else {
								rgba[l][0] = cmd.red[0]; rgba[l][1] = cmd.green[0]; rgba[l][2] = cmd.blue[0]; rgba[l][3] = cmd.alpha[0];
							}
						}
						if (hasDiffuse) { sample_quad<true>(textures.t[(flags >> 8) & 0xFFFu], false, cmd.u1, cmd.v1, wa, wb, wc, rgba); }
						if (hasLight) { sample_quad<true>(textures.t[flags >> 20], true, cmd.u2, cmd.v2, wa, wb, wc, rgba); }
					}
					const bool alphaFilter = (flags & CMD_ALPHA) != 0;
#pragma unroll
					for (int l = 0; l < 4; l++) {
						if (alphaFilter) {
							// ref: shader/fillerTemplates.h:155-176; lanes that are not visible read as 0 when clipping sides
							float opacity = rgba[l][3] * (1.0f / 255.0f);
							uint32_t target = (vis[l] || !clipSides) ? col[l] : 0u;
							float inv = 1.0f - opacity;
							float tr = (float)((target >> (shifts & 31u)) & 255u), tg = (float)((target >> ((shifts >> 8) & 31u)) & 255u);
							float tb = (float)((target >> ((shifts >> 16) & 31u)) & 255u), ta = (float)((target >> ((shifts >> 24) & 31u)) & 255u);
							rgba[l][0] = (rgba[l][0] * opacity) + (tr * inv);
							rgba[l][1] = (rgba[l][1] * opacity) + (tg * inv);
							rgba[l][2] = (rgba[l][2] * opacity) + (tb * inv);
							rgba[l][3] = (rgba[l][3] * opacity) + (ta * inv);
						}
						if (vis[l]) {
							col[l] = pack_rgba_ordered(saturated_byte(rgba[l][0]), saturated_byte(rgba[l][1]), saturated_byte(rgba[l][2]), saturated_byte(rgba[l][3]), shifts);
							dirty = true;
						}
					}
					// ref: shader/fillerTemplates.h:387-441 — alpha filtering leaves depth untouched when both buffers exist
					if (hasDepth && !alphaFilter) {
#pragma unroll
						for (int l = 0; l < 4; l++) { if (vis[l]) { dep[l] = lanes[0][l]; } }
					}
				} else if (hasDepth) {
#pragma unroll
					for (int l = 0; l < 4; l++) { if (vis[l]) { dep[l] = lanes[0][l]; dirty = true; } }
				}
			}
			__syncthreads();
		}
		processedBelow = windowEnd;
		remaining -= windowCount;
	}

	if (dirty) {
		// each thread owns 2 adjacent pixels in two rows: 8-byte stores, 128 contiguous bytes per row per half-warp
		if (hasColor) {
			if (in[0] && in[1]) { *(uint2 *)(row_ptr<uint32_t>(rp.color.data, rp.color.stride, y1) + x0) = make_uint2(col[0], col[1]); }
			else if (in[0]) { row_ptr<uint32_t>(rp.color.data, rp.color.stride, y1)[x0] = col[0]; }
			if (in[2] && in[3]) { *(uint2 *)(row_ptr<uint32_t>(rp.color.data, rp.color.stride, y2) + x0) = make_uint2(col[2], col[3]); }
			else if (in[2]) { row_ptr<uint32_t>(rp.color.data, rp.color.stride, y2)[x0] = col[2]; }
		}
		if (hasDepth) {
			if (in[0] && in[1]) { *(float2 *)(row_ptr<float>(rp.depth.data, rp.depth.stride, y1) + x0) = make_float2(dep[0], dep[1]); }
			else if (in[0]) { row_ptr<float>(rp.depth.data, rp.depth.stride, y1)[x0] = dep[0]; }
			if (in[2] && in[3]) { *(float2 *)(row_ptr<float>(rp.depth.data, rp.depth.stride, y2) + x0) = make_float2(dep[2], dep[3]); }
			else if (in[2]) { row_ptr<float>(rp.depth.data, rp.depth.stride, y2)[x0] = dep[2]; }
		}
	}
}

__global__ void __launch_bounds__(256) zero_kernel(uint32_t *data, int32_t count) {
	for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) { data[i] = 0u; }
}

} // namespace dfpsr

// ------------------------------------------------------------------------------------------------ host side

using namespace dfpsr;

struct FrameTask {
	TaskParams params;
};

struct dfpsr_renderer {
	bool receiving = false;
	dfpsr_image color{}, depth{};
	int32_t width = 0, height = 0;
	bool depthOnly = false;
	bool clear = false;
	uint32_t clearColor = 0;
	float clearDepth = 0.0f;
	std::vector<FrameTask> tasks;
	std::vector<DeviceBuffer> projected; // one per task, reused across frames
	std::vector<DeviceBuffer> uploads;   // host triangle batches
	TexTable textures{};
	int textureCount = 0;
	int32_t slotTotal = 0, blockTotal = 0;
	int64_t lastCommands = -1;
	DeviceBuffer slotCounts, blockCmds, blockRows, tileCount, tileOffset, tileCursor, totals, cmds, rows, tileList, big;
	uint32_t *hostTotals = nullptr; // pinned
	FrameDev frame{};
	bool countsZeroed = false;

	~dfpsr_renderer() {
		for (auto &b : projected) { b.release(); }
		for (auto &b : uploads) { b.release(); }
		DeviceBuffer *all[] = {&slotCounts, &blockCmds, &blockRows, &tileCount, &tileOffset, &tileCursor, &totals, &cmds, &rows, &tileList, &big};
		for (auto *b : all) { b->release(); }
		if (hostTotals) { cudaFreeHost(hostTotals); }
	}
};

static bool image_exists(const dfpsr_image *image) { return image != nullptr && image->data != nullptr; }

static int register_texture(dfpsr_renderer *r, const dfpsr_texture *t) {
	if (t == nullptr || t->data == nullptr) { return -1; }
	for (int i = 0; i < r->textureCount; i++) {
		if (r->textures.t[i].data == t->data && r->textures.t[i].log2width == t->log2width && r->textures.t[i].maxMipLevel == t->maxMipLevel) { return i; }
	}
	if (r->textureCount >= MAX_TEXTURES) { return -2; }
	TexDev &d = r->textures.t[r->textureCount];
	d.data = t->data; d.log2width = t->log2width; d.log2height = t->log2height;
	d.maxMipLevel = t->maxMipLevel; d.startOffset = t->startOffset; d.maxLevelMask = t->maxLevelMask;
	return r->textureCount++;
}

static int renderer_begin_internal(dfpsr_renderer *r, const dfpsr_image *color, const dfpsr_image *depth, bool depthOnly, bool clear, uint32_t clearColor, float clearDepth, cudaStream_t stream) {
	// ref: api/rendererAPI.cpp:151-168
	DFPSR_REQUIRE(!r->receiving, "Called renderer_begin on the same renderer twice without ending the previous batch!");
	r->color = image_exists(color) ? *color : dfpsr_image{};
	r->depth = image_exists(depth) ? *depth : dfpsr_image{};
	if (image_exists(color) && image_exists(depth)) {
		DFPSR_REQUIRE(color->width == depth->width && color->height == depth->height, "renderer_begin: colour buffer %dx%d and depth buffer %dx%d differ", color->width, color->height, depth->width, depth->height);
	}
	if (image_exists(color)) { r->width = color->width; r->height = color->height; }
	else if (image_exists(depth)) { r->width = depth->width; r->height = depth->height; }
	else { r->width = 0; r->height = 0; }
	r->receiving = true;
	r->depthOnly = depthOnly;
	r->clear = clear; r->clearColor = clearColor; r->clearDepth = clearDepth;
	r->tasks.clear();
	r->textureCount = 0;
	r->slotTotal = 0; r->blockTotal = 0;
	r->countsZeroed = false;
	r->frame.tilesX = (r->width + TILE - 1) / TILE;
	r->frame.tilesY = (r->height + TILE - 1) / TILE;
	if (!r->hostTotals) { DFPSR_CHECK_CUDA(cudaMallocHost((void **)&r->hostTotals, 8 * sizeof(uint32_t))); }
	(void)stream;
	return 0;
}

static int ensure_tile_counts(dfpsr_renderer *r, cudaStream_t stream) {
	if (r->countsZeroed) { return 0; }
	int32_t tiles = r->frame.tilesX * r->frame.tilesY;
	if (r->tileCount.reserve((size_t)(tiles + 1) * 4)) { return 1; }
	if (r->tileOffset.reserve((size_t)(tiles + 1) * 4)) { return 1; }
	if (r->tileCursor.reserve((size_t)(tiles + 1) * 4)) { return 1; }
	if (r->totals.reserve(8 * 4)) { return 1; }
	DFPSR_CHECK_CUDA(cudaMemsetAsync(r->tileCount.ptr, 0, (size_t)(tiles + 1) * 4, stream));
	r->countsZeroed = true;
	return 0;
}

// Appends a task: uploads nothing, launches projection + counting set-up.
static int add_task(dfpsr_renderer *r, TaskParams &task, cudaStream_t stream) {
	if (r->width <= 0 || r->height <= 0) { return 0; } // ref: renderCore.cpp:307-309 — no target, nothing to draw
	if (ensure_tile_counts(r, stream)) { return 1; }
	size_t index = r->tasks.size();
	task.slotBase = r->slotTotal;
	task.blockBase = r->blockTotal;
	task.width = r->width; task.height = r->height;
	task.depthOnly = r->depthOnly ? 1 : 0;
	int32_t blocks = (task.slotCount + SETUP_THREADS - 1) / SETUP_THREADS;
	if (blocks == 0) { return 0; }
	if (task.triangles == nullptr) {
		if (r->projected.size() <= index) { r->projected.resize(index + 1); }
		if (r->projected[index].reserve((size_t)task.pointCount * sizeof(PPoint) + 16)) { return 1; }
		task.projected = (PPoint *)r->projected[index].ptr;
		int grid = (task.pointCount + 255) / 256;
		if (grid > sm_count() * 8) { grid = sm_count() * 8; }
		if (grid > 0) { DFPSR_LAUNCH(project_kernel, grid, 256, 0, stream, task.points, task.pointCount, task.modelToWorld, task.camera, task.projected); }
	}
	// per-slot and per-block counters grow with the frame; growing must not lose earlier tasks' counts
	size_t slotsNeeded = (size_t)(r->slotTotal + task.slotCount) * 4, blocksNeeded = (size_t)(r->blockTotal + blocks) * 4;
	if (slotsNeeded > r->slotCounts.capacity || blocksNeeded > r->blockCmds.capacity) {
		// preserve contents when growing mid-frame
		DeviceBuffer *bufs[3] = {&r->slotCounts, &r->blockCmds, &r->blockRows};
		size_t needs[3] = {slotsNeeded, blocksNeeded, blocksNeeded};
		for (int i = 0; i < 3; i++) {
			if (needs[i] <= bufs[i]->capacity) { continue; }
			DeviceBuffer fresh;
			if (fresh.reserve(needs[i] * 2)) { return 1; }
			if (bufs[i]->ptr && !r->tasks.empty()) { DFPSR_CHECK_CUDA(cudaMemcpyAsync(fresh.ptr, bufs[i]->ptr, bufs[i]->capacity, cudaMemcpyDeviceToDevice, stream)); DFPSR_CHECK_CUDA(cudaStreamSynchronize(stream)); }
			bufs[i]->release();
			*bufs[i] = fresh;
		}
	}
	r->frame.slotCounts = (uint32_t *)r->slotCounts.ptr;
	r->frame.blockCmds = (uint32_t *)r->blockCmds.ptr;
	r->frame.blockRows = (uint32_t *)r->blockRows.ptr;
	r->frame.tileCount = (uint32_t *)r->tileCount.ptr;
	r->frame.tileOffset = (uint32_t *)r->tileOffset.ptr;
	r->frame.tileCursor = (uint32_t *)r->tileCursor.ptr;
	r->frame.totals = (uint32_t *)r->totals.ptr;
	DFPSR_LAUNCH(setup_kernel<false>, blocks, SETUP_THREADS, 0, stream, task, r->frame);
	r->slotTotal += task.slotCount;
	r->blockTotal += blocks;
	r->tasks.push_back(FrameTask{task});
	return 0;
}

static int renderer_end_internal(dfpsr_renderer *r, cudaStream_t stream) {
	// ref: api/rendererAPI.cpp:352-402
	DFPSR_REQUIRE(r->receiving, "Called renderer_end without renderer_begin!");
	r->receiving = false;
	r->lastCommands = 0;
	if (r->width <= 0 || r->height <= 0) { return 0; }
	int32_t tiles = r->frame.tilesX * r->frame.tilesY;
	uint32_t maxTile = 0;
	if (!r->tasks.empty()) {
		r->frame.blockCount = r->blockTotal;
		DFPSR_LAUNCH(scan_kernel, 1, 1024, 0, stream, r->frame);
		DFPSR_CHECK_CUDA(cudaMemcpyAsync(r->hostTotals, r->totals.ptr, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
		DFPSR_CHECK_CUDA(cudaStreamSynchronize(stream));
		uint32_t commandTotal = r->hostTotals[0], rowTotal = r->hostTotals[1], entryTotal = r->hostTotals[2];
		maxTile = r->hostTotals[3];
		r->lastCommands = commandTotal;
		if (commandTotal > 0) {
			if (r->cmds.reserve((size_t)commandTotal * sizeof(Cmd))) { return 1; }
			if (r->rows.reserve((size_t)rowTotal * sizeof(int2) + 16)) { return 1; }
			if (r->tileList.reserve((size_t)entryTotal * 4 + 16)) { return 1; }
			if (r->big.reserve((size_t)commandTotal * sizeof(BigCmd))) { return 1; }
			r->frame.cmds = (Cmd *)r->cmds.ptr;
			r->frame.rows = (int2 *)r->rows.ptr;
			r->frame.tileList = (uint32_t *)r->tileList.ptr;
			r->frame.big = (BigCmd *)r->big.ptr;
			for (FrameTask &t : r->tasks) {
				int32_t blocks = (t.params.slotCount + SETUP_THREADS - 1) / SETUP_THREADS;
				DFPSR_LAUNCH(setup_kernel<true>, blocks, SETUP_THREADS, 0, stream, t.params, r->frame);
			}
			DFPSR_LAUNCH(big_rows_kernel, sm_count() * 4, 256, 0, stream, r->frame);
		}
	} else if (r->clear) {
		if (ensure_tile_counts(r, stream)) { return 1; }
		r->frame.tileCount = (uint32_t *)r->tileCount.ptr;
		r->frame.tileOffset = (uint32_t *)r->tileOffset.ptr;
	}
	if (r->tasks.empty() && !r->clear) { return 0; }
	RasterParams rp;
	rp.color = r->color; rp.depth = r->depth;
	rp.width = r->width; rp.height = r->height;
	rp.depthOnly = r->depthOnly ? 1 : 0;
	rp.clear = r->clear ? 1 : 0;
	rp.clearColor = r->clearColor; rp.clearDepth = r->clearDepth;
	uint32_t capacity = 64;
	while (capacity < maxTile && capacity < 16384u) { capacity <<= 1; }
	rp.sortCapacity = capacity;
	size_t smem = sizeof(Cmd) * CHUNK + sizeof(int2) * CHUNK * TILE + (size_t)capacity * 4;
	if (smem > 48 * 1024) {
		DFPSR_CHECK_CUDA(cudaFuncSetAttribute(raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	}
	DFPSR_LAUNCH(raster_kernel, tiles, 256, smem, stream, r->frame, rp, r->textures);
	return 0;
}

static int fill_model_task(dfpsr_renderer *r, TaskParams &task, const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera) {
	memset(&task, 0, sizeof(task));
	task.points = model->points;
	task.polygons = model->polygons;
	task.pointCount = model->pointCount;
	task.polygonCount = model->polygonCount;
	task.slotCount = model->polygonCount * 2;
	task.modelToWorld = *modelToWorld;
	task.camera = *camera;
	task.filter = model->filter;
	task.diffuseIndex = r->depthOnly ? -1 : register_texture(r, &model->diffuse);
	task.lightIndex = r->depthOnly ? -1 : register_texture(r, &model->light);
	DFPSR_REQUIRE(task.diffuseIndex != -2 && task.lightIndex != -2, "more than %d distinct textures in one frame", MAX_TEXTURES);
	return 0;
}

extern "C" {

int dfpsr_renderer_create(dfpsr_renderer **out) {
	DFPSR_REQUIRE(out != nullptr, "dfpsr_renderer_create: null output");
	int n = 0;
	DFPSR_REQUIRE(cudaGetDeviceCount(&n) == cudaSuccess && n > 0, "no CUDA device available; dfpsr_b200 has no CPU fallback");
	*out = new (std::nothrow) dfpsr_renderer();
	DFPSR_REQUIRE(*out != nullptr, "out of host memory");
	return 0;
}

int dfpsr_renderer_destroy(dfpsr_renderer *renderer) {
	delete renderer;
	return 0;
}

int dfpsr_renderer_begin(dfpsr_renderer *renderer, const dfpsr_image *color, const dfpsr_image *depth) {
	DFPSR_REQUIRE(renderer != nullptr, "renderer_begin: renderer does not exist");
	return renderer_begin_internal(renderer, color, depth, false, false, 0u, 0.0f, nullptr);
}

int dfpsr_renderer_begin_cleared(dfpsr_renderer *renderer, const dfpsr_image *color, const dfpsr_image *depth, uint32_t packedClearColor, float clearDepth) {
	DFPSR_REQUIRE(renderer != nullptr, "renderer_begin: renderer does not exist");
	return renderer_begin_internal(renderer, color, depth, false, true, packedClearColor, clearDepth, nullptr);
}

int dfpsr_renderer_give_task(dfpsr_renderer *renderer, const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr && model != nullptr && modelToWorld != nullptr && camera != nullptr, "renderer_giveTask: null argument");
	DFPSR_REQUIRE(renderer->receiving, "Cannot call renderer_giveTask before renderer_begin!");
	// ref: api/modelAPI.cpp:228 — whole-model culling against the cull frustum on the host
	if (!dfpsr_camera_is_box_seen(camera, model->minBound, model->maxBound, modelToWorld)) { return 0; }
	if (model->polygonCount <= 0) { return 0; }
	TaskParams task;
	if (fill_model_task(renderer, task, model, modelToWorld, camera)) { return 1; }
	return add_task(renderer, task, as_stream(stream));
}

int dfpsr_renderer_give_task_triangles(dfpsr_renderer *renderer, const dfpsr_triangle *triangles, int32_t count, const dfpsr_texture *diffuse, const dfpsr_texture *light, int32_t filter, const dfpsr_camera *camera, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr && camera != nullptr, "renderer_giveTask_triangle: null argument");
	DFPSR_REQUIRE(renderer->receiving, "Cannot call renderer_giveTask_triangle before renderer_begin!");
	if (count <= 0) { return 0; }
	DFPSR_REQUIRE(triangles != nullptr, "renderer_giveTask_triangle: null triangles");
	size_t index = renderer->tasks.size();
	if (renderer->uploads.size() <= index) { renderer->uploads.resize(index + 1); }
	if (renderer->uploads[index].reserve((size_t)count * sizeof(dfpsr_triangle))) { return 1; }
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(renderer->uploads[index].ptr, triangles, (size_t)count * sizeof(dfpsr_triangle), cudaMemcpyHostToDevice, as_stream(stream)));
	TaskParams task;
	memset(&task, 0, sizeof(task));
	task.triangles = (const dfpsr_triangle *)renderer->uploads[index].ptr;
	task.triangleCount = count;
	task.slotCount = count;
	task.camera = *camera;
	task.filter = filter;
	task.diffuseIndex = register_texture(renderer, diffuse);
	task.lightIndex = register_texture(renderer, light);
	DFPSR_REQUIRE(task.diffuseIndex != -2 && task.lightIndex != -2, "more than %d distinct textures in one frame", MAX_TEXTURES);
	return add_task(renderer, task, as_stream(stream));
}

int dfpsr_renderer_end(dfpsr_renderer *renderer, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr, "renderer_end: renderer does not exist");
	return renderer_end_internal(renderer, as_stream(stream));
}

int dfpsr_renderer_last_command_count(dfpsr_renderer *renderer, int64_t *count, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr && count != nullptr, "renderer_last_command_count: null argument");
	(void)stream;
	*count = renderer->lastCommands;
	return 0;
}

static thread_local dfpsr_renderer *g_immediate = nullptr;

static int immediate_renderer(dfpsr_renderer **out) {
	if (g_immediate == nullptr) {
		if (dfpsr_renderer_create(&g_immediate)) { return 1; }
	}
	DFPSR_REQUIRE(!g_immediate->receiving, "model_render called re-entrantly");
	*out = g_immediate;
	return 0;
}

int dfpsr_model_render(const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_image *color, const dfpsr_image *depth, const dfpsr_camera *camera, void *stream) {
	if (model == nullptr) { return 0; } // ref: api/modelAPI.cpp:198
	dfpsr_renderer *r;
	if (immediate_renderer(&r)) { return 1; }
	if (renderer_begin_internal(r, color, depth, false, false, 0u, 0.0f, as_stream(stream))) { return 1; }
	int status = dfpsr_renderer_give_task(r, model, modelToWorld, camera, stream);
	if (status) { r->receiving = false; return status; }
	return renderer_end_internal(r, as_stream(stream));
}

int dfpsr_model_render_depth(const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_image *depth, const dfpsr_camera *camera, void *stream) {
	if (model == nullptr || !image_exists(depth)) { return 0; } // ref: api/modelAPI.cpp:203, renderCore.cpp:409
	dfpsr_renderer *r;
	if (immediate_renderer(&r)) { return 1; }
	if (renderer_begin_internal(r, nullptr, depth, true, false, 0u, 0.0f, as_stream(stream))) { return 1; }
	int status = dfpsr_renderer_give_task(r, model, modelToWorld, camera, stream);
	if (status) { r->receiving = false; return status; }
	return renderer_end_internal(r, as_stream(stream));
}

int dfpsr_model_render_views(const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_image *colors, const dfpsr_image *depths, const dfpsr_camera *cameras, int32_t count, int32_t clear, void *stream) {
	DFPSR_REQUIRE(model != nullptr && modelToWorld != nullptr && cameras != nullptr, "model_render_views: null argument");
	dfpsr_renderer *r;
	if (immediate_renderer(&r)) { return 1; }
	for (int32_t i = 0; i < count; i++) {
		const dfpsr_image *color = colors ? colors + i : nullptr, *depth = depths ? depths + i : nullptr;
		if (renderer_begin_internal(r, color, depth, false, clear != 0, 0u, 0.0f, as_stream(stream))) { return 1; }
		int status = dfpsr_renderer_give_task(r, model, modelToWorld, cameras + i, stream);
		if (status) { r->receiving = false; return status; }
		if (renderer_end_internal(r, as_stream(stream))) { return 1; }
	}
	return 0;
}

int dfpsr_project_points(const float *points, int32_t count, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, dfpsr_projected_point *outDevice, void *stream) {
	DFPSR_REQUIRE(points != nullptr && modelToWorld != nullptr && camera != nullptr && outDevice != nullptr, "project_points: null argument");
	if (count <= 0) { return 0; }
	int grid = (count + 255) / 256;
	if (grid > sm_count() * 8) { grid = sm_count() * 8; }
	DFPSR_LAUNCH(project_kernel, grid, 256, 0, as_stream(stream), points, count, *modelToWorld, *camera, (PPoint *)outDevice);
	return 0;
}

} // extern "C"
