/*
 * dfpsr_oracle.c — TEST INFRASTRUCTURE: plain-C restatement of the reference's hot-path algorithm.
 * See dfpsr_oracle.h for scope and parity status (pinned against the compiled reference).
 *
 * Style: straight-line scalar code, one function per reference routine, each citing the reference
 * file:line it follows (paths relative to /root/reference/Source/DFPSR unless they start with SDK/).
 * Build with -ffp-contract=off: the reference is built in ISO C++ mode, i.e. without FMA contraction.
 * Floating point follows the reference's SCALAR flavour (exact 1/x), see SURVEY.md §8c.
 */
#include "dfpsr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* 1 / sqrt(x) of the reference's scalar build (base/simd.h:4104): `1.0f / sqrt(value)` inside namespace dsr resolves to
 * ::sqrt(double) (checked with a static_assert against the reference headers), so the quotient is formed in double and
 * rounded to float once. */
#ifndef ORC_RSQRT
#define ORC_RSQRT(x) ((float)(1.0 / sqrt((double)(x))))
#endif

typedef struct { float x, y, z; } v3;

static v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 v3_from(const float *p) { return v3_make(p[0], p[1], p[2]); }
static float v3_dot(v3 a, v3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); } /* math/FVector.h:74 */

/* math/FVector.h:113-120 */
static v3 v3_normalize(v3 v) {
	float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
	if (l == 0.0f) { return v3_make(0.0f, 0.0f, 1.0f); }
	return v3_make(v.x / l, v.y / l, v.z / l);
}

/* math/FMatrix3x3.h:52-58 */
static v3 mat_transform(const float *xa, const float *ya, const float *za, v3 p) {
	return v3_make(
	  p.x * xa[0] + p.y * ya[0] + p.z * za[0],
	  p.x * xa[1] + p.y * ya[1] + p.z * za[1],
	  p.x * xa[2] + p.y * ya[2] + p.z * za[2]);
}
/* math/FMatrix3x3.h:63-69 */
static v3 mat_transform_transposed(const float *xa, const float *ya, const float *za, v3 p) {
	return v3_make(
	  p.x * xa[0] + p.y * xa[1] + p.z * xa[2],
	  p.x * ya[0] + p.y * ya[1] + p.z * ya[2],
	  p.x * za[0] + p.y * za[1] + p.z * za[2]);
}
/* math/Transform3D.h:41-43 */
static v3 transform_point(const dfpsr_transform3d *t, v3 p) {
	v3 r = mat_transform(t->xAxis, t->yAxis, t->zAxis, p);
	return v3_make(r.x + t->position[0], r.y + t->position[1], r.z + t->position[2]);
}
/* math/Transform3D.h:50-52 */
static v3 transform_point_transposed_inverse(const dfpsr_transform3d *t, v3 p) {
	return mat_transform_transposed(t->xAxis, t->yAxis, t->zAxis, v3_make(p.x - t->position[0], p.y - t->position[1], p.z - t->position[2]));
}

/* ------------------------------------------------------------------------------------------- camera */

static void set_plane(float *dst, v3 normal, float offset) { /* math/FPlane3D.h:36 */
	v3 n = v3_normalize(normal);
	dst[0] = n.x; dst[1] = n.y; dst[2] = n.z; dst[3] = offset;
}

/* implementation/render/Camera.h:56-72 */
static int frustum_perspective(float planes[6][4], float nearClip, float farClip, float widthSlope, float heightSlope) {
	set_plane(planes[0], v3_make(-1.0f, 0.0f, -widthSlope), 0.0f);
	set_plane(planes[1], v3_make(1.0f, 0.0f, -widthSlope), 0.0f);
	set_plane(planes[2], v3_make(0.0f, 1.0f, -heightSlope), 0.0f);
	set_plane(planes[3], v3_make(0.0f, -1.0f, -heightSlope), 0.0f);
	set_plane(planes[4], v3_make(0.0f, 0.0f, -1.0f), -nearClip);
	set_plane(planes[5], v3_make(0.0f, 0.0f, 1.0f), farClip);
	return farClip == INFINITY ? 5 : 6;
}
/* implementation/render/Camera.h:47-55 */
static int frustum_orthogonal(float planes[6][4], float halfWidth, float halfHeight) {
	set_plane(planes[0], v3_make(-1.0f, 0.0f, 0.0f), halfWidth);
	set_plane(planes[1], v3_make(1.0f, 0.0f, 0.0f), halfWidth);
	set_plane(planes[2], v3_make(0.0f, 1.0f, 0.0f), halfHeight);
	set_plane(planes[3], v3_make(0.0f, -1.0f, 0.0f), halfHeight);
	return 4;
}

/* implementation/render/Camera.h:113-156 */
void orc_camera_create(dfpsr_camera *c) {
	const float cullRatio = 1.0001f, clipRatio = 2.0f;
	memset(c->cullPlanes, 0, sizeof(c->cullPlanes));
	memset(c->clipPlanes, 0, sizeof(c->clipPlanes));
	if (c->perspective) {
		float widthSlope = c->widthSlope;
		float heightSlope = widthSlope * c->imageHeight / c->imageWidth;
		c->heightSlope = heightSlope;
		c->cullPlaneCount = frustum_perspective(c->cullPlanes, c->nearClip, c->farClip, widthSlope * cullRatio, heightSlope * cullRatio);
		c->clipPlaneCount = frustum_perspective(c->clipPlanes, c->nearClip, c->farClip, widthSlope * clipRatio, heightSlope * clipRatio);
	} else {
		float halfWidth = c->widthSlope;
		float halfHeight = halfWidth * c->imageHeight / c->imageWidth;
		c->heightSlope = halfHeight;
		c->nearClip = -FLT_MAX;
		c->farClip = INFINITY;
		c->cullPlaneCount = frustum_orthogonal(c->cullPlanes, halfWidth * cullRatio, halfHeight * cullRatio);
		c->clipPlaneCount = frustum_orthogonal(c->clipPlanes, halfWidth * clipRatio, halfHeight * clipRatio);
	}
	c->invWidthSlope = 0.5f / c->widthSlope;
	c->invHeightSlope = 0.5f / c->heightSlope;
}

typedef struct { v3 cs; float isx, isy; int64_t fx, fy; } ppoint;

/* implementation/render/Camera.h:160-187 */
static ppoint camera_to_screen(const dfpsr_camera *c, v3 cs) {
	ppoint r;
	r.cs = cs;
	if (c->perspective) {
		float invDepth;
		if (cs.z > 0.0f) { invDepth = 1.0f / cs.z; } else { invDepth = 0.0f; }
		float centerShear = cs.z * 0.5f;
		float preX = (cs.x * c->invWidthSlope + centerShear) * c->imageWidth;
		float preY = (-cs.y * c->invHeightSlope + centerShear) * c->imageHeight;
		r.isx = preX * invDepth;
		r.isy = preY * invDepth;
	} else {
		r.isx = (cs.x * c->invWidthSlope + 0.5f) * c->imageWidth;
		r.isy = (-cs.y * c->invHeightSlope + 0.5f) * c->imageHeight;
	}
	float subX = r.isx * 256.0f, subY = r.isy * 256.0f; /* constants.h:40 unitsPerPixel */
	r.fx = (int64_t)subX;
	r.fy = (int64_t)subY;
	return r;
}

static ppoint world_to_screen(const dfpsr_camera *c, v3 world) { /* Camera.h:157, :188 */
	return camera_to_screen(c, transform_point_transposed_inverse(&c->location, world));
}

static int plane_inside(const float *plane, v3 p) { /* math/FPlane3D.h:39-45 */
	return (v3_dot(v3_from(plane), p) - plane[3]) <= 0.0f;
}

/* implementation/render/Camera.h:73-95 + :202-217 */
int orc_camera_is_box_seen(const dfpsr_camera *c, const float *mn, const float *mx, const dfpsr_transform3d *m2w) {
	v3 corners[8];
	for (int i = 0; i < 8; i++) {
		v3 p = v3_make((i & 1) ? mx[0] : mn[0], (i & 2) ? mx[1] : mn[1], (i & 4) ? mx[2] : mn[2]);
		corners[i] = transform_point_transposed_inverse(&c->location, transform_point(m2w, p));
	}
	int anyOutside = 0;
	for (int s = 0; s < c->cullPlaneCount; s++) {
		int anyInside = 0;
		for (int p = 0; p < 8; p++) {
			if (plane_inside(c->cullPlanes[s], corners[p])) { anyInside = 1; } else { anyOutside = 1; }
		}
		if (!anyInside) { return 0; }
	}
	return anyOutside ? 1 : 2;
}

static void ppoint_export(dfpsr_projected_point *o, const ppoint *p) {
	o->cs[0] = p->cs.x; o->cs[1] = p->cs.y; o->cs[2] = p->cs.z;
	o->is[0] = p->isx; o->is[1] = p->isy; o->pad_ = 0;
	o->flat[0] = p->fx; o->flat[1] = p->fy;
}
static ppoint ppoint_import(const dfpsr_projected_point *p) {
	ppoint r;
	r.cs = v3_from(p->cs); r.isx = p->is[0]; r.isy = p->is[1]; r.fx = p->flat[0]; r.fy = p->flat[1];
	return r;
}

/* api/modelAPI.cpp:238-242 */
void orc_project_points(const float *points, int32_t count, const dfpsr_transform3d *m2w, const dfpsr_camera *camera, dfpsr_projected_point *out) {
	for (int32_t i = 0; i < count; i++) {
		ppoint p = world_to_screen(camera, transform_point(m2w, v3_from(points + 3 * i)));
		ppoint_export(out + i, &p);
	}
}

/* ------------------------------------------------------------------------------------------- textures */

/* api/textureAPI.cpp:30-41 */
static uint32_t find_log2_size(uint32_t size) {
	for (uint32_t l = 0; l < 15; l++) { if ((1u << l) >= size) { return l; } }
	return 15;
}

/* api/textureAPI.cpp:65-78 + implementation/image/Texture.h:63-102 */
void orc_texture_layout(dfpsr_texture *out, int32_t width, int32_t height, int32_t resolutions) {
	uint32_t log2w = find_log2_size((uint32_t)width), log2h = find_log2_size((uint32_t)height);
	uint32_t maxMip = (uint32_t)(resolutions - 1);
	if (maxMip > log2w) { maxMip = log2w; }
	if (maxMip > log2h) { maxMip = log2h; }
	if (maxMip > 15) { maxMip = 15; }
	uint32_t highest = 1u << (log2w + log2h);
	uint64_t pixelCount = 0;
	uint32_t levelCount = highest;
	for (int32_t level = (int32_t)maxMip; level >= 0; level--) { pixelCount |= levelCount; levelCount >>= 2; }
	out->data = NULL;
	out->log2width = log2w; out->log2height = log2h; out->maxMipLevel = maxMip;
	out->startOffset = (uint32_t)pixelCount & ~highest;
	out->maxLevelMask = highest - 1;
	out->totalPixels = (uint32_t)pixelCount;
}

static uint32_t tex_layer_offset(const dfpsr_texture *t, uint32_t mip) { /* api/textureAPI.h:79-85 */
	return t->startOffset & (t->maxLevelMask >> (2 * mip));
}

/* api/textureAPI.cpp:44-87 */
void orc_texture_generate_pyramid(uint32_t *px, const dfpsr_texture *t) {
	for (uint32_t level = 1; level <= t->maxMipLevel; level++) {
		uint32_t tw = 1u << (t->log2width - level), th = 1u << (t->log2height - level);
		const uint32_t *src = px + tex_layer_offset(t, level - 1);
		uint32_t *dst = px + tex_layer_offset(t, level);
		uint32_t sw = tw * 2;
		for (uint32_t y = 0; y < th; y++) {
			for (uint32_t x = 0; x < tw; x++) {
				uint32_t a = src[(2 * y) * sw + 2 * x], b = src[(2 * y) * sw + 2 * x + 1];
				uint32_t c = src[(2 * y + 1) * sw + 2 * x], d = src[(2 * y + 1) * sw + 2 * x + 1];
				uint32_t out = 0;
				for (int s = 0; s < 32; s += 8) {
					out |= ((((a >> s) & 255u) + ((b >> s) & 255u) + ((c >> s) & 255u) + ((d >> s) & 255u)) / 4u) << s;
				}
				dst[y * tw + x] = out;
			}
		}
	}
}

/* api/textureAPI.h:253-263 weightColors, on 16-bit lane pairs */
static uint32_t weight_colors(uint32_t colorA, uint32_t weightA, uint32_t colorB, uint32_t weightB) {
	uint32_t lowA = colorA & 0x00FF00FFu, lowB = colorB & 0x00FF00FFu;
	uint32_t highA = (colorA & 0xFF00FF00u) >> 8, highB = (colorB & 0xFF00FF00u) >> 8;
	/* 16-bit lanes: products <= 255 * 256 so the two lanes never carry into each other */
	uint32_t low = (((lowA & 0xFFFFu) * weightA + (lowB & 0xFFFFu) * weightB) & 0xFFFFu)
	             | (((((lowA >> 16) * weightA + (lowB >> 16) * weightB)) & 0xFFFFu) << 16);
	uint32_t high = (((highA & 0xFFFFu) * weightA + (highB & 0xFFFFu) * weightB) & 0xFFFFu)
	              | (((((highA >> 16) * weightA + (highB >> 16) * weightB)) & 0xFFFFu) << 16);
	return ((low >> 8) & 0x00FF00FFu) | (high & 0xFF00FF00u);
}

/* api/textureAPI.h:342-438 texture_sample_bilinear<SQUARE=false, SINGLE_LAYER, MIP_INSIDE=true, HIGHEST_RESOLUTION> */
static uint32_t tex_sample_bilinear(const dfpsr_texture *t, float u, float v, uint32_t mip) {
	uint32_t scaleU = (256u << t->log2width) >> mip;
	uint32_t scaleV = (256u << t->log2height) >> mip;
	uint32_t subX = (uint32_t)((u + 256.0f) * (float)scaleU) - 128u;
	uint32_t subY = (uint32_t)((v + 256.0f) * (float)scaleV) - 128u;
	uint32_t wx = subX & 0xFF, wy = subY & 0xFF;
	uint32_t left = subX >> 8, top = subY >> 8;
	uint32_t maskX = ((1u << t->log2width) - 1u) >> mip, maskY = ((1u << t->log2height) - 1u) >> mip;
	uint32_t right = (left + 1) & maskX, bottom = (top + 1) & maskY;
	left &= maskX; top &= maskY;
	uint32_t log2Stride = t->log2width - mip;
	const uint32_t *data = t->data + tex_layer_offset(t, mip);
	uint32_t c00 = data[(top << log2Stride) | left], c10 = data[(top << log2Stride) | right];
	uint32_t c01 = data[(bottom << log2Stride) | left], c11 = data[(bottom << log2Stride) | right];
	/* api/textureAPI.h:315-326 texture_interpolate_color_bilinear */
	uint32_t upper = weight_colors(c00, 256u - wx, c10, wx);
	uint32_t lower = weight_colors(c01, 256u - wx, c11, wx);
	return weight_colors(upper, 256u - wy, lower, wy);
}

/* ---- test hooks (known-answer tests of the reference: test/tests/TextureTest.cpp:13-19, :22-372) */
uint32_t orc_texture_layer_offset(const dfpsr_texture *t, uint32_t mip) { return tex_layer_offset(t, mip); }
/* api/textureAPI.h:111-177 texture_getPixelOffset with tiling: coordinates wrap inside the level */
uint32_t orc_texture_pixel_offset(const dfpsr_texture *t, uint32_t x, uint32_t y, uint32_t mip) {
	if (mip > t->maxMipLevel) { mip = t->maxMipLevel; }
	uint32_t maskX = ((1u << t->log2width) - 1u) >> mip, maskY = ((1u << t->log2height) - 1u) >> mip;
	return tex_layer_offset(t, mip) + (((y & maskY) << (t->log2width - mip)) | (x & maskX));
}
/* api/textureAPI.h:265-275 texture_interpolate_color_linear: weight 0..256 of colorB */
uint32_t orc_interpolate_color_linear(uint32_t colorA, uint32_t colorB, uint32_t weight) { return weight_colors(colorA, 256u - weight, colorB, weight); }
uint32_t orc_texture_sample_bilinear(const dfpsr_texture *t, float u, float v, uint32_t mip) { return tex_sample_bilinear(t, u, v, mip); }

/* api/textureAPI.h:472-495: one level per quad from lanes 0, 1, 2 */
static uint32_t tex_mip_level(const dfpsr_texture *t, const float *u, const float *v) {
	float offsetUX = fabsf(u[0] - u[1]), offsetUY = fabsf(u[0] - u[2]);
	float offsetVX = fabsf(v[0] - v[1]), offsetVY = fabsf(v[0] - v[2]);
	float offsetU = (offsetUX > offsetUY ? offsetUX : offsetUY) * (float)(1u << t->log2width);
	float offsetV = (offsetVX > offsetVY ? offsetVX : offsetVY) * (float)(1u << t->log2height);
	float offset = offsetU > offsetV ? offsetU : offsetV;
	uint32_t result = 0;
	if (offset > 2.0f) { result = 1; }
	if (offset > 4.0f) { result = 2; }
	if (offset > 8.0f) { result = 3; }
	if (offset > 16.0f) { result = 4; }
	if (result > t->maxMipLevel) { result = t->maxMipLevel; }
	return result;
}

/* ------------------------------------------------------------------------------------------- triangle set-up */

typedef struct { int32_t l, t, w, h; } irect;
static irect irect_make(int32_t l, int32_t t, int32_t w, int32_t h) { irect r = {l, t, w, h}; return r; }
static int irect_overlaps(irect a, irect b) { /* math/IRect.h:77 */
	return a.l < b.l + b.w && a.l + a.w > b.l && a.t < b.t + b.h && a.t + a.h > b.t;
}
static int32_t clamp_i32(int32_t lo, int32_t v, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }
static int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
static int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static irect irect_cut(irect a, irect b) { /* math/IRect.h:56-66 */
	if (!irect_overlaps(a, b)) { return irect_make(0, 0, 0, 0); }
	int32_t l = imax(a.l, b.l), t = imax(a.t, b.t), r = imin(a.l + a.w, b.l + b.w), bo = imin(a.t + a.h, b.t + b.h);
	return irect_make(l, t, r - l, bo - t);
}

/* implementation/render/ITriangle2D.cpp:31-43 */
static irect triangle_bound(const ppoint *p) {
	int32_t rx[3], ry[3];
	for (int i = 0; i < 3; i++) {
		rx[i] = (int32_t)((p[i].fx + 128) / 256);
		ry[i] = (int32_t)((p[i].fy + 128) / 256);
	}
	int32_t l = imin(rx[0], imin(rx[1], rx[2])) - 1, t = imin(ry[0], imin(ry[1], ry[2])) - 1;
	int32_t r = imax(rx[0], imax(rx[1], rx[2])) + 1, b = imax(ry[0], imax(ry[1], ry[2])) + 1;
	return irect_make(l, t, r - l, b - t);
}

/* implementation/render/ITriangle2D.cpp:55-60 */
static int is_frontfacing(const ppoint *p) {
	return ((p[2].fx - p[0].fx) * (p[1].fy - p[0].fy)) + ((p[2].fy - p[0].fy) * (p[0].fx - p[1].fx)) < 0;
}

typedef struct { int32_t left, right; } row_interval;

/* implementation/render/ITriangle2D.cpp:86-150 cutConvexEdge, one row at a time in closed form
 * (limit is decremented by offsetY once per row there; the int64 arithmetic is exact so this is identical). */
static void cut_convex_edge(int64_t sx, int64_t sy, int64_t ex, int64_t ey, row_interval *rows, irect bound) {
	int32_t leftBound = bound.l, topBound = bound.t, rightBound = bound.l + bound.w, bottomBound = bound.t + bound.h;
	int64_t originX = 128 + (int64_t)bound.l * 256;
	int64_t originY = 128 + (int64_t)bound.t * 256;
	int64_t threshold = (sx > ex || (sx == ex && sy > ey)) ? -1 : 0;
	int64_t normalX = ey - sy, normalY = sx - ex;
	int64_t offsetX = normalX * 256, offsetY = normalY * 256;
	int64_t valueOrigin = ((originX - sx) * normalX) + ((originY - sy) * normalY);
	if (normalX != 0) {
		int64_t limit0 = threshold - valueOrigin + (offsetX * leftBound);
		for (int32_t y = topBound; y < bottomBound; y++) {
			int64_t limit = limit0 - offsetY * (int64_t)(y - topBound);
			if (normalX < 0) {
				int32_t leftSide = imin(imax(leftBound, (int32_t)((limit + 1) / offsetX + 1)), rightBound);
				rows[y - topBound].left = imax(rows[y - topBound].left, leftSide);
			} else {
				int32_t rightSide = imin(imax(leftBound, (int32_t)(limit / offsetX + 1)), rightBound);
				rows[y - topBound].right = imin(rows[y - topBound].right, rightSide);
			}
		}
	} else if (normalY != 0) {
		for (int32_t y = topBound; y < bottomBound; y++) {
			int64_t valueRow = valueOrigin + offsetY * (int64_t)(y - topBound);
			if (valueRow > threshold) {
				rows[y - topBound].left = rightBound;
				rows[y - topBound].right = leftBound;
			}
		}
	}
}

/* implementation/render/ITriangle2D.cpp:152-169 */
static void rasterize_triangle(const ppoint *p, row_interval *rows, irect bound) {
	int degenerate = (p[0].fx == p[1].fx && p[0].fy == p[1].fy) || (p[1].fx == p[2].fx && p[1].fy == p[2].fy) || (p[2].fx == p[0].fx && p[2].fy == p[0].fy);
	for (int32_t r = 0; r < bound.h; r++) {
		rows[r].left = degenerate ? bound.l + bound.w : bound.l;
		rows[r].right = degenerate ? bound.l : bound.l + bound.w;
	}
	if (!degenerate) {
		for (int i = 0; i < 3; i++) {
			int j = (i + 1) % 3;
			cut_convex_edge(p[i].fx, p[i].fy, p[j].fx, p[j].fy, rows, bound);
		}
	}
}

/* test hook: row intervals [left, right) of one triangle given by its 1/256-pixel corners, inside bound (l, t, w, h) */
void orc_rasterize_rows(const int64_t *fx, const int64_t *fy, int32_t l, int32_t t, int32_t w, int32_t h, int32_t *rowsOut) {
	ppoint p[3];
	memset(p, 0, sizeof(p));
	for (int i = 0; i < 3; i++) { p[i].fx = fx[i]; p[i].fy = fy[i]; }
	rasterize_triangle(p, (row_interval*)rowsOut, irect_make(l, t, w, h));
}
int orc_is_frontfacing(const int64_t *fx, const int64_t *fy) {
	ppoint p[3];
	memset(p, 0, sizeof(p));
	for (int i = 0; i < 3; i++) { p[i].fx = fx[i]; p[i].fy = fy[i]; }
	return is_frontfacing(p);
}

typedef struct { int affine; float start[3], dx[3], dy[3]; } projection;

/* implementation/render/ITriangle2D.cpp:182-300 */
static projection get_projection(const ppoint *p, const float *subB, const float *subC, int perspective) {
	float offsetX[3], offsetY[3], mult[3], normalX[3], normalY[3], targetWeight[3];
	for (int i = 0; i < 3; i++) {
		int j = (i + 1) % 3;
		offsetX[i] = p[j].isy - p[i].isy;
		offsetY[i] = p[i].isx - p[j].isx;
	}
	for (int i = 0; i < 3; i++) {
		int o = (i + 2) % 3;
		float other = ((p[o].isx - p[i].isx) * offsetX[i]) + ((p[o].isy - p[i].isy) * offsetY[i]);
		mult[o] = (other == 0.0f) ? 0.0f : 1.0f / other;
	}
	for (int i = 0; i < 3; i++) {
		normalX[i] = offsetX[i] * mult[i];
		normalY[i] = offsetY[i] * mult[i];
	}
	for (int i = 0; i < 3; i++) {
		int o = (i + 2) % 3;
		targetWeight[o] = p[i].isx * -normalX[i] + p[i].isy * -normalY[i];
	}
	float adx[3] = {normalX[1], normalX[2], normalX[0]};
	float ady[3] = {normalY[1], normalY[2], normalY[0]};
	projection r;
	if (!perspective) {
		float W[3] = {p[0].cs.z, p[1].cs.z, p[2].cs.z};
		r.affine = 1;
		r.start[0] = W[0] * targetWeight[0] + W[1] * targetWeight[1] + W[2] * targetWeight[2];
		r.start[1] = targetWeight[0] * subB[0] + targetWeight[1] * subB[1] + targetWeight[2] * subB[2];
		r.start[2] = targetWeight[0] * subC[0] + targetWeight[1] * subC[1] + targetWeight[2] * subC[2];
		r.dx[0] = W[0] * adx[0] + W[1] * adx[1] + W[2] * adx[2];
		r.dx[1] = adx[0] * subB[0] + adx[1] * subB[1] + adx[2] * subB[2];
		r.dx[2] = adx[0] * subC[0] + adx[1] * subC[1] + adx[2] * subC[2];
		r.dy[0] = W[0] * ady[0] + W[1] * ady[1] + W[2] * ady[2];
		r.dy[1] = ady[0] * subB[0] + ady[1] * subB[1] + ady[2] * subB[2];
		r.dy[2] = ady[0] * subC[0] + ady[1] * subC[1] + ady[2] * subC[2];
	} else {
		float IW[3] = {1.0f / p[0].cs.z, 1.0f / p[1].cs.z, 1.0f / p[2].cs.z};
		r.affine = 0;
		r.start[0] = IW[0] * targetWeight[0] + IW[1] * targetWeight[1] + IW[2] * targetWeight[2];
		r.start[1] = IW[0] * targetWeight[0] * subB[0] + IW[1] * targetWeight[1] * subB[1] + IW[2] * targetWeight[2] * subB[2];
		r.start[2] = IW[0] * targetWeight[0] * subC[0] + IW[1] * targetWeight[1] * subC[1] + IW[2] * targetWeight[2] * subC[2];
		r.dx[0] = IW[0] * adx[0] + IW[1] * adx[1] + IW[2] * adx[2];
		r.dx[1] = IW[0] * adx[0] * subB[0] + IW[1] * adx[1] * subB[1] + IW[2] * adx[2] * subB[2];
		r.dx[2] = IW[0] * adx[0] * subC[0] + IW[1] * adx[1] * subC[1] + IW[2] * adx[2] * subC[2];
		r.dy[0] = IW[0] * ady[0] + IW[1] * ady[1] + IW[2] * ady[2];
		r.dy[1] = IW[0] * ady[0] * subB[0] + IW[1] * ady[1] * subB[1] + IW[2] * ady[2] * subB[2];
		r.dy[2] = IW[0] * ady[0] * subC[0] + IW[1] * ady[1] * subC[1] + IW[2] * ady[2] * subC[2];
	}
	return r;
}

/* implementation/render/ITriangle2D.h:82-100: start + dx * (x + 0.5) + dy * (y + 0.5), left to right */
static void projection_at(const projection *p, int32_t x, int32_t y, float *out) {
	float fx = (float)x + 0.5f, fy = (float)y + 0.5f;
	for (int k = 0; k < 3; k++) { out[k] = p->start[k] + (p->dx[k] * fx) + (p->dy[k] * fy); }
}

/* ------------------------------------------------------------------------------------------- shader */

typedef struct {
	int hasDiffuse, hasLight, hasFade, colorless;
	float red[3], green[3], blue[3], alpha[3]; /* already scaled, shader/RgbaMultiply.h:45-60 */
	float u1[3], v1[3], u2[3], v2[3];
	const dfpsr_texture *diffuse, *light;
} shader_data;

static int almost_zero(float v) { return v > -0.001f && v < 0.001f; } /* shader/fillerTemplates.h:37 */
static int almost_one(float v) { return v > 0.999f && v < 1.001f; }
static int almost_zero3(const float *c) { return almost_zero(c[0]) && almost_zero(c[1]) && almost_zero(c[2]); }
static int almost_one3(const float *c) { return almost_one(c[0]) && almost_one(c[1]) && almost_one(c[2]); }
static int almost_same3(const float *c) { return almost_zero(c[0] - c[1]) && almost_zero(c[0] - c[2]) && almost_zero(c[1] - c[2]); }

/* shader/RgbaMultiply.h:37-60, :110-116. colors/texCoords are [corner][channel]. */
static void shader_setup(shader_data *s, const float colors[3][4], const float tex[3][4], const dfpsr_texture *diffuse, const dfpsr_texture *light) {
	s->diffuse = diffuse; s->light = light;
	s->hasDiffuse = diffuse != NULL && diffuse->data != NULL;
	s->hasLight = light != NULL && light->data != NULL;
	float scale = 255.0f;
	if (s->hasDiffuse) { scale *= 1.0f / 255.0f; }
	if (s->hasLight) { scale *= 1.0f / 255.0f; }
	for (int c = 0; c < 3; c++) {
		s->red[c] = colors[c][0] * scale; s->green[c] = colors[c][1] * scale;
		s->blue[c] = colors[c][2] * scale; s->alpha[c] = colors[c][3] * scale;
		s->u1[c] = tex[c][0]; s->v1[c] = tex[c][1]; s->u2[c] = tex[c][2]; s->v2[c] = tex[c][3];
	}
	s->hasFade = !(almost_same3(s->red) && almost_same3(s->green) && almost_same3(s->blue) && almost_same3(s->alpha));
	s->colorless = almost_one3(s->red) && almost_one3(s->green) && almost_one3(s->blue) && almost_one3(s->alpha);
}

/* shader/shaderMethods.h:39-44 */
static float interpolate(const float *d, float wa, float wb, float wc) {
	return d[0] * wa + d[1] * wb + d[2] * wc;
}

static void unpack_rgba(uint32_t c, float *r, float *g, float *b, float *a) { /* shader/shaderTypes.h:39-43 */
	*r = (float)(c & 255u); *g = (float)((c >> 8) & 255u); *b = (float)((c >> 16) & 255u); *a = (float)(c >> 24);
}

static void sample_quad(const dfpsr_texture *t, int highestResolution, const float *cu, const float *cv, const float *wa, const float *wb, const float *wc, float out[4][4]) {
	float u[4], v[4];
	for (int l = 0; l < 4; l++) { u[l] = interpolate(cu, wa[l], wb[l], wc[l]); v[l] = interpolate(cv, wa[l], wb[l], wc[l]); }
	uint32_t mip = highestResolution ? 0u : tex_mip_level(t, u, v); /* shader/shaderMethods.h:63-83 */
	for (int l = 0; l < 4; l++) {
		uint32_t c = tex_sample_bilinear(t, u[l], v[l], mip);
		unpack_rgba(c, &out[l][0], &out[l][1], &out[l][2], &out[l][3]);
	}
}

/* shader/RgbaMultiply.h:75-106 getPixels_2x2; out[lane][r,g,b,a] */
static void shade_quad(const shader_data *s, const float *wa, const float *wb, const float *wc, float out[4][4]) {
	if (s->hasDiffuse && !s->hasLight && s->colorless && !s->hasFade) {
		sample_quad(s->diffuse, 0, s->u1, s->v1, wa, wb, wc, out);
	} else if (s->hasLight && !s->hasDiffuse && s->colorless && !s->hasFade) {
		sample_quad(s->light, 1, s->u2, s->v2, wa, wb, wc, out);
	} else {
		for (int l = 0; l < 4; l++) {
			if (s->hasFade) {
				out[l][0] = interpolate(s->red, wa[l], wb[l], wc[l]);
				out[l][1] = interpolate(s->green, wa[l], wb[l], wc[l]);
				out[l][2] = interpolate(s->blue, wa[l], wb[l], wc[l]);
				out[l][3] = interpolate(s->alpha, wa[l], wb[l], wc[l]);
			} else {
				out[l][0] = s->red[0]; out[l][1] = s->green[0]; out[l][2] = s->blue[0]; out[l][3] = s->alpha[0];
			}
		}
		float sampled[4][4];
		if (s->hasDiffuse) {
			sample_quad(s->diffuse, 0, s->u1, s->v1, wa, wb, wc, sampled);
			for (int l = 0; l < 4; l++) { for (int c = 0; c < 4; c++) { out[l][c] = out[l][c] * sampled[l][c]; } }
		}
		if (s->hasLight) {
			sample_quad(s->light, 1, s->u2, s->v2, wa, wb, wc, sampled);
			for (int l = 0; l < 4; l++) { for (int c = 0; c < 4; c++) { out[l][c] = out[l][c] * sampled[l][c]; } }
		}
	}
}

static const int packIndex[4][4] = { /* implementation/image/PackOrder.h:85-96: byte index of r, g, b, a */
	{0, 1, 2, 3}, {2, 1, 0, 3}, {1, 2, 3, 0}, {3, 2, 1, 0}
};

static uint32_t saturated_byte(float v) { /* PackOrder.h:186-197: clampUpper 255.1 then truncate */
	float c = v < 255.1f ? v : 255.1f;
	return (uint32_t)c;
}

static uint32_t pack_float_color(const float *rgba, int order) { /* PackOrder.h:204-213 */
	const int *ix = packIndex[order];
	return (saturated_byte(rgba[0]) << (8 * ix[0])) | (saturated_byte(rgba[1]) << (8 * ix[1]))
	     | (saturated_byte(rgba[2]) << (8 * ix[2])) | (saturated_byte(rgba[3]) << (8 * ix[3]));
}

static void unpack_ordered(uint32_t c, int order, float *rgba) { /* shader/shaderTypes.h:44-48 */
	const int *ix = packIndex[order];
	for (int k = 0; k < 4; k++) { rgba[k] = (float)((c >> (8 * ix[k])) & 255u); }
}

/* ------------------------------------------------------------------------------------------- fill */

typedef struct {
	const dfpsr_image *color, *depth; /* either may be NULL */
	int colorWrite, depthRead, depthWrite, alphaFilter, affine;
	int32_t maxHeight;
} fill_mode;

static uint32_t *color_px(const dfpsr_image *im, int32_t x, int32_t y) { return (uint32_t*)((uint8_t*)im->data + (size_t)y * im->stride) + x; }
static float *depth_px(const dfpsr_image *im, int32_t x, int32_t y) { return (float*)((uint8_t*)im->data + (size_t)y * im->stride) + x; }

/* shader/fillerTemplates.h:139-187 fillQuadSuper + :190-244 fillRowSuper body for ONE quad.
 * upper/lower: the running (depth, B, C) plane values for lane 0 and lane 2; lanes 1 and 3 are +dx. */
static void fill_quad(const fill_mode *m, const shader_data *s, int clipSides, int32_t x, int32_t y1, const float *lanes /* [3][4] */, row_interval upperRow, row_interval lowerRow, int hasTop, int hasBottom) {
	int32_t y2 = y1 + 1;
	int32_t px[4] = {x, x + 1, x, x + 1};
	/* fillerTemplates.h:302-331: a row pair with an empty row points both row pointers at the non-empty row ("repeat the
	 * lower/upper row to avoid reading outside"). Only the unclipped inner quads of the last row pair of an odd-height
	 * target ever touch the repeated row: there lanes 2 and 3 read and overwrite the pixels of lanes 0 and 1. */
	int32_t py[4] = {hasTop ? y1 : y2, hasTop ? y1 : y2, hasBottom ? y2 : y1, hasBottom ? y2 : y1};
	float depth[4], wa[4], wb[4], wc[4];
	for (int l = 0; l < 4; l++) {
		depth[l] = lanes[0 * 4 + l];
		if (m->affine) {
			wb[l] = lanes[1 * 4 + l];
			wc[l] = lanes[2 * 4 + l];
		} else {
			float linearDepth = 1.0f / lanes[0 * 4 + l]; /* scalar-flavour reciprocal, base/simd.h:4064-4067 */
			wb[l] = lanes[1 * 4 + l] * linearDepth;
			wc[l] = lanes[2 * 4 + l] * linearDepth;
		}
		wa[l] = 1.0f - (wb[l] + wc[l]);
	}
	/* fillerTemplates.h:93-138 */
	int vis[4];
	for (int l = 0; l < 4; l++) {
		int clip = 1;
		if (clipSides) {
			row_interval row = (l < 2) ? upperRow : lowerRow;
			clip = px[l] >= row.left && px[l] < row.right;
		}
		int front = 1;
		if (m->depthRead) {
			if (clipSides && !clip) { front = 0; }
			else {
				float old = *depth_px(m->depth, px[l], py[l]);
				front = m->affine ? (depth[l] < old) : (depth[l] > old);
			}
		}
		vis[l] = clip && front;
	}
	if (!(vis[0] || vis[1] || vis[2] || vis[3])) { return; }
	if (m->colorWrite) {
		float colors[4][4];
		shade_quad(s, wa, wb, wc, colors);
		int order = m->color->packOrder;
		uint32_t targets[4] = {0u, 0u, 0u, 0u};
		if (m->alphaFilter) {
			/* clippedRead: all four reads happen before any write; lanes that are not visible read 0 when clipping sides */
			for (int l = 0; l < 4; l++) { targets[l] = (vis[l] || !clipSides) ? *color_px(m->color, px[l], py[l]) : 0u; }
		}
		for (int l = 0; l < 4; l++) {
			if (m->alphaFilter) {
				float opacity = colors[l][3] * (1.0f / 255.0f);
				uint32_t target = targets[l];
				float dst[4];
				unpack_ordered(target, order, dst);
				float inv = 1.0f - opacity;
				for (int c = 0; c < 4; c++) { colors[l][c] = (colors[l][c] * opacity) + (dst[c] * inv); }
			}
			if (vis[l]) { *color_px(m->color, px[l], py[l]) = pack_float_color(colors[l], order); }
		}
	}
	if (m->depthWrite) {
		for (int l = 0; l < 4; l++) { if (vis[l]) { *depth_px(m->depth, px[l], py[l]) = depth[l]; } }
	}
}

static uint32_t round_up_even(uint32_t x) { return (x + 1u) & ~1u; }
static uint32_t round_down_even(uint32_t x) { return x & ~1u; }

/* shader/fillerTemplates.h:247-385 fillShapeSuper */
static void fill_shape(const fill_mode *m, const shader_data *s, const projection *proj, int32_t startRow, int32_t rowCount, const row_interval *rows) {
	float dx2[3] = {proj->dx[0] * 2.0f, proj->dx[1] * 2.0f, proj->dx[2] * 2.0f};
	for (int32_t y1 = startRow; y1 < startRow + rowCount; y1 += 2) {
		int32_t y2 = y1 + 1;
		row_interval upperRow = rows[y1 - startRow], lowerRow = rows[y2 - startRow];
		int32_t outerStart = imin(upperRow.left, lowerRow.left), outerEnd = imax(upperRow.right, lowerRow.right);
		int32_t innerStart = imax(upperRow.left, lowerRow.left), innerEnd = imin(upperRow.right, lowerRow.right);
		int32_t outerBlockStart = (int32_t)round_down_even((uint32_t)outerStart), outerBlockEnd = (int32_t)round_up_even((uint32_t)outerEnd);
		int32_t innerBlockStart = (int32_t)round_up_even((uint32_t)innerStart), innerBlockEnd = (int32_t)round_down_even((uint32_t)innerEnd);
		if (y2 >= m->maxHeight) { lowerRow.right = lowerRow.left; }
		int hasTop = upperRow.right > upperRow.left, hasBottom = lowerRow.right > lowerRow.left;
		if (!(hasTop || hasBottom)) { continue; }
		float upper[3], lower[3];
		projection_at(proj, outerBlockStart, y1, upper);
		for (int k = 0; k < 3; k++) { lower[k] = upper[k] + proj->dy[k]; }
		float lanes[12];
		if (innerBlockEnd <= innerBlockStart) {
			for (int32_t x = outerBlockStart; x < outerBlockEnd; x += 2) {
				for (int k = 0; k < 3; k++) { lanes[k * 4 + 0] = upper[k]; lanes[k * 4 + 1] = upper[k] + proj->dx[k]; lanes[k * 4 + 2] = lower[k]; lanes[k * 4 + 3] = lower[k] + proj->dx[k]; }
				fill_quad(m, s, 1, x, y1, lanes, upperRow, lowerRow, hasTop, hasBottom);
				for (int k = 0; k < 3; k++) { upper[k] = upper[k] + dx2[k]; lower[k] = lower[k] + dx2[k]; }
			}
		} else {
			for (int32_t x = outerBlockStart; x < innerBlockStart; x += 2) {
				for (int k = 0; k < 3; k++) { lanes[k * 4 + 0] = upper[k]; lanes[k * 4 + 1] = upper[k] + proj->dx[k]; lanes[k * 4 + 2] = lower[k]; lanes[k * 4 + 3] = lower[k] + proj->dx[k]; }
				fill_quad(m, s, 1, x, y1, lanes, upperRow, lowerRow, hasTop, hasBottom);
				for (int k = 0; k < 3; k++) { upper[k] = upper[k] + dx2[k]; lower[k] = lower[k] + dx2[k]; }
			}
			/* full quads: the four lanes advance independently by repeated addition (fillRowSuper) */
			int32_t quadCount = (innerBlockEnd - innerBlockStart) / 2;
			for (int k = 0; k < 3; k++) { lanes[k * 4 + 0] = upper[k]; lanes[k * 4 + 1] = upper[k] + proj->dx[k]; lanes[k * 4 + 2] = lower[k]; lanes[k * 4 + 3] = lower[k] + proj->dx[k]; }
			for (int32_t x = innerBlockStart; x < innerBlockEnd; x += 2) {
				fill_quad(m, s, 0, x, y1, lanes, upperRow, lowerRow, hasTop, hasBottom);
				for (int k = 0; k < 3; k++) { for (int l = 0; l < 4; l++) { lanes[k * 4 + l] = lanes[k * 4 + l] + dx2[k]; } }
			}
			for (int k = 0; k < 3; k++) { upper[k] = upper[k] + (dx2[k] * (float)quadCount); lower[k] = lower[k] + (dx2[k] * (float)quadCount); }
			for (int32_t x = innerBlockEnd; x < outerBlockEnd; x += 2) {
				for (int k = 0; k < 3; k++) { lanes[k * 4 + 0] = upper[k]; lanes[k * 4 + 1] = upper[k] + proj->dx[k]; lanes[k * 4 + 2] = lower[k]; lanes[k * 4 + 3] = lower[k] + proj->dx[k]; }
				fill_quad(m, s, 1, x, y1, lanes, upperRow, lowerRow, hasTop, hasBottom);
				for (int k = 0; k < 3; k++) { upper[k] = upper[k] + dx2[k]; lower[k] = lower[k] + dx2[k]; }
			}
		}
	}
}

struct orc_renderer;
typedef struct {
	const dfpsr_image *color, *depth;
	int32_t width, height;
	const dfpsr_camera *camera;
	int filter;
	shader_data shader;
	int64_t commands;
	struct orc_renderer *queue; /* when set, triangles are queued like CommandQueue::add instead of being drawn at once */
} draw_context;
static void queue_push(struct orc_renderer *r, const draw_context *ctx, const ppoint *p, const float *subB, const float *subC);

/* implementation/render/renderCore.cpp:203-217 executeTriangleDrawing + shader/fillerTemplates.h:387-441 fillShape */
static void execute_triangle(draw_context *ctx, const ppoint *p, const float *subB, const float *subC) {
	ctx->commands++;
	if (ctx->queue != NULL) { queue_push(ctx->queue, ctx, p, subB, subC); return; }
	irect clip = irect_make(0, 0, ctx->width, ctx->height);
	irect whole = triangle_bound(p);
	if (!irect_overlaps(whole, clip)) { return; }
	irect un = irect_cut(whole, clip);
	int32_t alignedTop = (un.t / 2) * 2, alignedBottom = ((un.t + un.h + 1) / 2) * 2; /* ITriangle2D.cpp:70-75 (non-negative) */
	irect bound = irect_make(un.l, alignedTop, un.w, alignedBottom - alignedTop);
	row_interval *rows = (row_interval*)malloc(sizeof(row_interval) * (size_t)bound.h);
	rasterize_triangle(p, rows, bound);
	projection proj = get_projection(p, subB, subC, ctx->camera->perspective);
	int hasColor = ctx->color != NULL && ctx->color->data != NULL, hasDepth = ctx->depth != NULL && ctx->depth->data != NULL;
	fill_mode m;
	m.color = ctx->color; m.depth = ctx->depth; m.affine = proj.affine;
	m.maxHeight = ctx->height;
	m.alphaFilter = 0;
	if (hasDepth && hasColor) {
		if (ctx->filter != DFPSR_FILTER_SOLID) { m.colorWrite = 1; m.depthRead = 1; m.depthWrite = 0; m.alphaFilter = 1; }
		else { m.colorWrite = 1; m.depthRead = 1; m.depthWrite = 1; }
	} else if (hasDepth) {
		m.colorWrite = 0; m.depthRead = 1; m.depthWrite = 1;
	} else {
		m.colorWrite = 1; m.depthRead = 0; m.depthWrite = 0; m.alphaFilter = ctx->filter != DFPSR_FILTER_SOLID;
	}
	fill_shape(&m, &ctx->shader, &proj, bound.t, bound.h, rows);
	free(rows);
}

/* ------------------------------------------------------------------------------------------- cull / clip */

enum { VIS_HIDDEN = 0, VIS_FULL = 1, VIS_PARTIAL = 2 };

/* implementation/render/renderCore.cpp:172-198 */
static int triangle_visibility(const ppoint *p, const dfpsr_camera *c, int clipFrustum) {
	int planeCount = clipFrustum ? c->clipPlaneCount : c->cullPlaneCount;
	const float (*planes)[4] = clipFrustum ? c->clipPlanes : c->cullPlanes;
	int outside[3][6];
	for (int k = 0; k < 3; k++) { for (int s = 0; s < planeCount; s++) { outside[k][s] = !plane_inside(planes[s], p[k].cs); } }
	for (int s = 0; s < planeCount; s++) { if (outside[0][s] && outside[1][s] && outside[2][s]) { return VIS_HIDDEN; } }
	for (int k = 0; k < 3; k++) { for (int s = 0; s < planeCount; s++) { if (outside[k][s]) { return VIS_PARTIAL; } } }
	return VIS_FULL;
}

typedef struct { v3 cs; float subB, subC; int state; float value; } sub_vertex;

static float inverse_lerp(float a, float b, float value) { /* renderCore.cpp:52-59 */
	float c = b - a;
	if (c == 0.0f) { return 0.5f; }
	return (value - a) / c;
}

static sub_vertex sub_lerp(const sub_vertex *a, const sub_vertex *b, float ratio) { /* renderCore.cpp:41-46 */
	sub_vertex r;
	float inv = 1.0f - ratio;
	r.cs = v3_make(a->cs.x * inv + b->cs.x * ratio, a->cs.y * inv + b->cs.y * ratio, a->cs.z * inv + b->cs.z * ratio);
	r.subB = a->subB * inv + b->subB * ratio;
	r.subC = a->subC * inv + b->subC * ratio;
	r.state = 0; r.value = 0.0f;
	return r;
}

#define MAX_POINTS 9
typedef struct { int count; sub_vertex v[MAX_POINTS]; } clipped_triangle;

/* renderCore.cpp:104-169 ClippedTriangle::clip */
static void clip_plane(clipped_triangle *t, const float *plane) {
	enum { USE = 0, DELETE = 1, MODIFIED = 2 };
	if (!(t->count >= 3 && t->count < MAX_POINTS)) { return; }
	int outsideCount = 0, lastOutside = 0;
	for (int v = 0; v < t->count; v++) {
		float distance = v3_dot(v3_from(plane), t->v[v].cs) - plane[3];
		t->v[v].value = distance;
		if (distance > 0.0f) { outsideCount++; lastOutside = v; t->v[v].state = DELETE; } else { t->v[v].state = USE; }
	}
	if (outsideCount == 0) { return; }
	if (outsideCount >= t->count) { t->count = 0; return; }
	if (outsideCount == 1) {
		int cur = lastOutside, prev = (lastOutside - 1 + t->count) % t->count, next = (lastOutside + 1) % t->count;
		float r1 = inverse_lerp(t->v[prev].value, t->v[cur].value, 0.0f);
		float r2 = inverse_lerp(t->v[cur].value, t->v[next].value, 0.0f);
		sub_vertex cutStart = sub_lerp(&t->v[prev], &t->v[cur], r1);
		sub_vertex cutEnd = sub_lerp(&t->v[cur], &t->v[next], r2);
		t->v[lastOutside] = cutStart;
		/* insertVertex(next, cutEnd), renderCore.cpp:87-99 */
		if (t->count < MAX_POINTS) {
			for (int v = t->count - 1; v >= next; v--) { t->v[v + 1] = t->v[v]; }
			t->v[next] = cutEnd;
			t->count++;
		}
	} else {
		for (int cur = 0; cur < t->count; cur++) {
			int prev = (cur - 1 + t->count) % t->count, next = (cur + 1) % t->count;
			if (t->v[cur].state == DELETE) {
				if (t->v[prev].state == USE) {
					float r = inverse_lerp(t->v[prev].value, t->v[cur].value, 0.0f);
					t->v[cur] = sub_lerp(&t->v[prev], &t->v[cur], r);
					t->v[cur].state = MODIFIED;
				} else if (t->v[next].state == USE) {
					float r = inverse_lerp(t->v[cur].value, t->v[next].value, 0.0f);
					t->v[cur] = sub_lerp(&t->v[cur], &t->v[next], r);
					t->v[cur].state = MODIFIED;
				}
			}
		}
		if (outsideCount > 2) {
			for (int v = t->count - 1; v >= 0; v--) {
				if (t->v[v].state == DELETE) {
					for (int k = v; k < t->count - 1; k++) { t->v[k] = t->v[k + 1]; }
					t->count--;
				}
			}
		}
	}
}

static clipped_triangle clip_triangle(const ppoint *p, const dfpsr_camera *c) { /* renderCore.cpp:70-75, :245-250 */
	clipped_triangle t;
	memset(&t, 0, sizeof(t));
	t.v[0].cs = p[0].cs; t.v[0].subB = 0.0f; t.v[0].subC = 0.0f;
	t.v[1].cs = p[1].cs; t.v[1].subB = 1.0f; t.v[1].subC = 0.0f;
	t.v[2].cs = p[2].cs; t.v[2].subB = 0.0f; t.v[2].subC = 1.0f;
	t.count = 3;
	for (int s = 0; s < c->clipPlaneCount; s++) { clip_plane(&t, c->clipPlanes[s]); }
	return t;
}

/* implementation/render/renderCore.cpp:261-341 renderTriangleFromData + renderTriangleWithShader */
static void render_triangle(draw_context *ctx, const ppoint *p, const float colors[3][4], const float tex[3][4], const dfpsr_texture *diffuse, const dfpsr_texture *light) {
	const dfpsr_camera *c = ctx->camera;
	if (triangle_visibility(p, c, 0) == VIS_HIDDEN) { return; }
	float alphas[3] = {colors[0][3], colors[1][3], colors[2][3]};
	if (ctx->filter == DFPSR_FILTER_ALPHA && almost_zero3(alphas)) { return; }
	shader_setup(&ctx->shader, colors, tex, diffuse, light);
	if (triangle_visibility(p, c, 1) == VIS_FULL) {
		if (is_frontfacing(p)) {
			float subB[3] = {0.0f, 1.0f, 0.0f}, subC[3] = {0.0f, 0.0f, 1.0f};
			execute_triangle(ctx, p, subB, subC);
		}
	} else {
		clipped_triangle t = clip_triangle(p, c);
		for (int i = 0; i < t.count - 2; i++) { /* renderCore.cpp:252-257, :220-240 */
			const sub_vertex *a = &t.v[0], *b = &t.v[1 + i], *cc = &t.v[2 + i];
			float subB[3] = {a->subB, b->subB, cc->subB}, subC[3] = {a->subC, b->subC, cc->subC};
			ppoint q[3] = {camera_to_screen(c, a->cs), camera_to_screen(c, b->cs), camera_to_screen(c, cc->cs)};
			if (is_frontfacing(q)) { execute_triangle(ctx, q, subB, subC); }
		}
	}
}

static int context_init(draw_context *ctx, const dfpsr_image *color, const dfpsr_image *depth, const dfpsr_camera *camera, int filter) {
	memset(ctx, 0, sizeof(*ctx));
	ctx->color = color; ctx->depth = depth; ctx->camera = camera; ctx->filter = filter;
	if (color != NULL && color->data != NULL) { ctx->width = color->width; ctx->height = color->height; } /* renderCore.cpp:289-310 */
	else if (depth != NULL && depth->data != NULL) { ctx->width = depth->width; ctx->height = depth->height; }
	else { return 0; }
	return 1;
}

/* api/modelAPI.cpp:214-281 (== implementation/render/model/Model.cpp:135-197) */
static void submit_model(draw_context *ctx, const dfpsr_model *model, const dfpsr_transform3d *m2w, const dfpsr_camera *camera);

int64_t orc_model_render(const dfpsr_model *model, const dfpsr_transform3d *m2w, const dfpsr_image *color, const dfpsr_image *depth, const dfpsr_camera *camera) {
	draw_context ctx;
	if (!context_init(&ctx, color, depth, camera, model->filter)) { return 0; }
	if (!orc_camera_is_box_seen(camera, model->minBound, model->maxBound, m2w)) { return 0; }
	submit_model(&ctx, model, m2w, camera);
	return ctx.commands;
}

static void submit_model(draw_context *ctxp, const dfpsr_model *model, const dfpsr_transform3d *m2w, const dfpsr_camera *camera) {
	ppoint *projected = (ppoint*)malloc(sizeof(ppoint) * (size_t)(model->pointCount > 0 ? model->pointCount : 1));
	for (int32_t i = 0; i < model->pointCount; i++) {
		projected[i] = world_to_screen(camera, transform_point(m2w, v3_from(model->points + 3 * i)));
	}
	for (int32_t i = 0; i < model->polygonCount; i++) {
		const dfpsr_polygon *poly = model->polygons + i;
		int triangles = poly->pointIndices[3] != -1 ? 2 : 1;
		for (int t = 0; t < triangles; t++) {
			int ia = 0, ib = 1 + t, ic = 2 + t;
			ppoint p[3] = {projected[poly->pointIndices[ia]], projected[poly->pointIndices[ib]], projected[poly->pointIndices[ic]]};
			float colors[3][4], tex[3][4];
			memcpy(colors[0], poly->colors[ia], 16); memcpy(colors[1], poly->colors[ib], 16); memcpy(colors[2], poly->colors[ic], 16);
			memcpy(tex[0], poly->texCoords[ia], 16); memcpy(tex[1], poly->texCoords[ib], 16); memcpy(tex[2], poly->texCoords[ic], 16);
			render_triangle(ctxp, p, colors, tex, &model->diffuse, &model->light);
		}
	}
	free(projected);
}

/* api/rendererAPI.cpp:503-519 */
int64_t orc_render_triangles(const dfpsr_triangle *triangles, int32_t count, const dfpsr_texture *diffuse, const dfpsr_texture *light, int32_t filter, const dfpsr_image *color, const dfpsr_image *depth, const dfpsr_camera *camera) {
	draw_context ctx;
	if (!context_init(&ctx, color, depth, camera, filter)) { return 0; }
	for (int32_t i = 0; i < count; i++) {
		ppoint p[3] = {ppoint_import(&triangles[i].pos[0]), ppoint_import(&triangles[i].pos[1]), ppoint_import(&triangles[i].pos[2])};
		render_triangle(&ctx, p, triangles[i].colors, triangles[i].texCoords, diffuse, light);
	}
	return ctx.commands;
}

/* ------------------------------------------------------------------------------------------- renderer with occlusion grid */

/* api/rendererAPI.cpp:141-150 RendererImpl: deferred queue + 16-pixel-cell occlusion grid */
typedef struct {
	ppoint p[3];
	float subB[3], subC[3];
	shader_data shader;
	int filter;
	const dfpsr_camera *camera;
	dfpsr_camera cameraCopy;
	int occluded;
} queued_command;

struct orc_renderer {
	int receiving;
	dfpsr_image color, depth;
	int32_t width, height, gridWidth, gridHeight;
	float *grid; int32_t gridAllocW, gridAllocH; /* depthGrid keeps its old size when large enough (rendererAPI.cpp:186-189) */
	int occluded;
	queued_command *commands; int64_t count, capacity;
	int64_t lastOccluded;
	int wireframe; /* renderer_end(renderer, debugWireframe = true) for the next frame */
};
#define CELL_SIZE 16 /* api/rendererAPI.cpp:36 */

static void queue_push(struct orc_renderer *r, const draw_context *ctx, const ppoint *p, const float *subB, const float *subC) {
	if (r->count == r->capacity) {
		r->capacity = r->capacity ? r->capacity * 2 : 1024;
		r->commands = (queued_command*)realloc(r->commands, sizeof(queued_command) * (size_t)r->capacity);
	}
	queued_command *c = &r->commands[r->count++];
	memcpy(c->p, p, sizeof(c->p)); memcpy(c->subB, subB, 12); memcpy(c->subC, subC, 12);
	c->shader = ctx->shader; c->filter = ctx->filter; c->cameraCopy = *ctx->camera; c->camera = NULL; c->occluded = 0;
}

orc_renderer *orc_renderer_create(void) { return (orc_renderer*)calloc(1, sizeof(orc_renderer)); }
void orc_renderer_destroy(orc_renderer *r) { if (r) { free(r->grid); free(r->commands); free(r); } }

/* api/rendererAPI.cpp:151-168 */
void orc_renderer_begin(orc_renderer *r, const dfpsr_image *color, const dfpsr_image *depth) {
	r->receiving = 1;
	memset(&r->color, 0, sizeof(r->color)); memset(&r->depth, 0, sizeof(r->depth));
	if (color != NULL && color->data != NULL) { r->color = *color; }
	if (depth != NULL && depth->data != NULL) { r->depth = *depth; }
	if (r->color.data != NULL) { r->width = r->color.width; r->height = r->color.height; }
	else if (r->depth.data != NULL) { r->width = r->depth.width; r->height = r->depth.height; }
	r->gridWidth = (r->width + (CELL_SIZE - 1)) / CELL_SIZE;
	r->gridHeight = (r->height + (CELL_SIZE - 1)) / CELL_SIZE;
	r->occluded = 0;
	r->count = 0;
}
int orc_renderer_has_occluders(const orc_renderer *r) { return r->occluded; }

/* image_readPixel_clamp on the grid image (a missing image reads 0) */
static float grid_read(const orc_renderer *r, int32_t x, int32_t y) {
	if (r->grid == NULL) { return 0.0f; }
	x = clamp_i32(0, x, r->gridAllocW - 1); y = clamp_i32(0, y, r->gridAllocH - 1);
	return r->grid[y * r->gridAllocW + x];
}
/* api/rendererAPI.cpp:181-192 */
static void prepare_for_occlusion(orc_renderer *r) {
	if (!r->occluded) {
		if (!(r->grid != NULL && r->gridAllocW >= r->gridWidth && r->gridAllocH >= r->gridHeight)) {
			free(r->grid);
			r->gridAllocW = r->gridWidth; r->gridAllocH = r->gridHeight;
			r->grid = (float*)malloc(sizeof(float) * (size_t)(r->gridAllocW * r->gridAllocH > 0 ? r->gridAllocW * r->gridAllocH : 1));
		}
		for (int32_t i = 0; i < r->gridAllocW * r->gridAllocH; i++) { r->grid[i] = INFINITY; }
	}
	r->occluded = 1;
}
/* api/rendererAPI.cpp:169-180 (IRect(l, t, w, h): right = l + w) */
static irect outer_cell_bound(const orc_renderer *r, irect pixelBound) {
	int32_t minX = pixelBound.l / CELL_SIZE, maxX = (pixelBound.l + pixelBound.w) / CELL_SIZE + 1;
	int32_t minY = pixelBound.t / CELL_SIZE, maxY = (pixelBound.t + pixelBound.h) / CELL_SIZE + 1;
	if (minX < 0) { minX = 0; } if (minY < 0) { minY = 0; }
	if (maxX > r->gridWidth) { maxX = r->gridWidth; } if (maxY > r->gridHeight) { maxY = r->gridHeight; }
	return irect_make(minX, minY, maxX - minX, maxY - minY);
}
/* api/rendererAPI.cpp:99-126 */
static int point_inside_of_hull(const ppoint *hull, int count, int64_t x, int64_t y) {
	for (int c = 0; c < count; c++) {
		int nc = c + 1 == count ? 0 : c + 1;
		int64_t dirX = hull[nc].fy - hull[c].fy, dirY = hull[c].fx - hull[nc].fx;
		if (!((dirX * (x - hull[c].fx)) + (dirY * (y - hull[c].fy)) <= 0)) { return 0; }
	}
	return 1;
}
/* api/rendererAPI.cpp:218-241 occludeFromSortedHull */
static void occlude_from_sorted_hull(orc_renderer *r, const ppoint *hull, int count, irect pixelBound) {
	if (!(pixelBound.w > CELL_SIZE && pixelBound.h > CELL_SIZE)) { return; }
	float distance = 0.0f;
	for (int c = 0; c < count; c++) { if (hull[c].cs.z > distance) { distance = hull[c].cs.z; } }
	irect outer = outer_cell_bound(r, pixelBound);
	for (int32_t cy = outer.t; cy < outer.t + outer.h; cy++) {
		for (int32_t cx = outer.l; cx < outer.l + outer.w; cx++) {
			int64_t l = (int64_t)cx * CELL_SIZE * 256, t = (int64_t)cy * CELL_SIZE * 256, rr = l + CELL_SIZE * 256, b = t + CELL_SIZE * 256; /* IRect * unitsPerPixel */
			if (point_inside_of_hull(hull, count, l, t) && point_inside_of_hull(hull, count, rr, t) && point_inside_of_hull(hull, count, l, b) && point_inside_of_hull(hull, count, rr, b)) {
				if (distance < grid_read(r, cx, cy)) { r->grid[cy * r->gridAllocW + cx] = distance; }
			}
		}
	}
}
/* api/rendererAPI.cpp:89-95 getPixelBoundFromProjection: merge of 1x1 rectangles at flat / 256 (truncating division) */
static irect pixel_bound_from_projection(const ppoint *hull, int count) {
	int32_t l = (int32_t)(hull[0].fx / 256), t = (int32_t)(hull[0].fy / 256), rr = l + 1, b = t + 1;
	for (int p = 1; p < count; p++) {
		int32_t x = (int32_t)(hull[p].fx / 256), y = (int32_t)(hull[p].fy / 256);
		l = imin(l, x); t = imin(t, y); rr = imax(rr, x + 1); b = imax(b, y + 1);
	}
	return irect_make(l, t, rr - l, b - t);
}
static void box_corners(v3 *out, const float *mn, const float *mx) { /* api/rendererAPI.cpp:259-267 */
	for (int i = 0; i < 8; i++) { out[i] = v3_make((i & 4) ? mx[0] : mn[0], (i & 2) ? mx[1] : mn[1], (i & 1) ? mx[2] : mn[2]); }
}
/* api/rendererAPI.cpp:38-74 jarvisConvexHullAlgorithm */
static int counter_clockwise(const ppoint *p, const ppoint *q, const ppoint *rr) {
	return (q->fy - p->fy) * (rr->fx - q->fx) - (q->fx - p->fx) * (rr->fy - q->fy) < 0;
}
static void jarvis(ppoint *out, int *outCount, const ppoint *in, int n) {
	if (n < 3) { *outCount = n; for (int p = 0; p < n; p++) { out[p] = in[p]; } return; }
	int l = 0;
	*outCount = 0;
	for (int i = 1; i < n; i++) { if (in[i].fx < in[l].fx) { l = i; } }
	int p = l;
	do {
		if (*outCount >= n) { return; }
		out[(*outCount)++] = in[p];
		int q = (p + 1) % n;
		for (int i = 0; i < n; i++) { if (counter_clockwise(&in[p], &in[i], &in[q])) { q = i; } }
		p = q;
	} while (p != l);
}
/* api/rendererAPI.cpp:268-299 occludeFromBox (+ :76-88 projectHull) */
void orc_renderer_occlude_from_box(orc_renderer *r, const float *mn, const float *mx, const dfpsr_transform3d *m2w, const dfpsr_camera *camera) {
	prepare_for_occlusion(r);
	v3 local[8]; ppoint projected[8], hull[8];
	box_corners(local, mn, mx);
	for (int p = 0; p < 8; p++) {
		v3 cameraPoint = transform_point_transposed_inverse(&camera->location, transform_point(m2w, local[p]));
		v3 narrow = v3_make(cameraPoint.x * 0.5f, cameraPoint.y * 0.5f, cameraPoint.z * 1.0f);
		for (int s = 0; s < camera->cullPlaneCount; s++) { if (!plane_inside(camera->cullPlanes[s], narrow)) { return; } }
		projected[p] = camera_to_screen(camera, cameraPoint);
	}
	int count = 0;
	jarvis(hull, &count, projected, 8);
	occlude_from_sorted_hull(r, hull, count, pixel_bound_from_projection(hull, count));
}
/* api/rendererAPI.cpp:242-258 occludeFromExistingTriangles */
void orc_renderer_occlude_from_existing_triangles(orc_renderer *r) {
	prepare_for_occlusion(r);
	for (int64_t t = 0; t < r->count; t++) {
		if (r->commands[t].filter == DFPSR_FILTER_SOLID) { occlude_from_sorted_hull(r, r->commands[t].p, 3, triangle_bound(r->commands[t].p)); }
	}
}
/* api/rendererAPI.cpp:403-477 occludeFromTopRows: scans the first pixel row of every cell row of the depth buffer */
void orc_renderer_occlude_from_top_rows(orc_renderer *r, const dfpsr_camera *camera) {
	prepare_for_occlusion(r);
	if (r->depth.data == NULL) { return; }
	for (int32_t y = 0, gy = 0; y < r->height; y += CELL_SIZE, gy++) {
		const float *depthPixel = depth_px(&r->depth, 0, y);
		int32_t x = 0, right = CELL_SIZE - 1;
		for (int32_t gx = 0; gx < r->gridWidth; gx++) {
			float extreme = camera->perspective ? INFINITY : 0.0f;
			if (right >= r->width) { right = r->width; }
			while (x < right) {
				float v = *depthPixel;
				if (camera->perspective) { if (v < extreme) { extreme = v; } } else { if (v > extreme) { extreme = v; } }
				depthPixel += 1; x += 1;
			}
			float maxDistance = camera->perspective ? 1.0f / extreme : extreme;
			if (maxDistance < r->grid[gy * r->gridAllocW + gx]) { r->grid[gy * r->gridAllocW + gx] = maxDistance; }
			right += CELL_SIZE;
		}
	}
}
/* api/rendererAPI.cpp:302-351 isHullOccluded / isBoxOccluded, negated like renderer_isBoxVisible (:538-543) */
int orc_renderer_is_box_visible(const orc_renderer *r, const float *mn, const float *mx, const dfpsr_transform3d *m2w, const dfpsr_camera *camera) {
	v3 local[8], cameraPoints[8]; ppoint projected[8];
	box_corners(local, mn, mx);
	for (int p = 0; p < 8; p++) {
		cameraPoints[p] = transform_point_transposed_inverse(&camera->location, transform_point(m2w, local[p]));
		projected[p] = camera_to_screen(camera, cameraPoints[p]);
	}
	for (int s = 0; s < camera->cullPlaneCount; s++) {
		int allOutside = 1;
		for (int p = 0; p < 8; p++) { if (plane_inside(camera->cullPlanes[s], cameraPoints[p])) { allOutside = 0; break; } }
		if (allOutside) { return 0; }
	}
	irect pixelBound = pixel_bound_from_projection(projected, 8);
	float closest = INFINITY;
	for (int c = 0; c < 8; c++) { if (projected[c].cs.z < closest) { closest = projected[c].cs.z; } }
	irect outer = outer_cell_bound(r, pixelBound);
	for (int32_t cy = outer.t; cy < outer.t + outer.h; cy++) {
		for (int32_t cx = outer.l; cx < outer.l + outer.w; cx++) { if (closest < grid_read(r, cx, cy)) { return 1; } }
	}
	return 0;
}
/* api/modelAPI.cpp:214-281 model_render_threaded */
void orc_renderer_give_task(orc_renderer *r, const dfpsr_model *model, const dfpsr_transform3d *m2w, const dfpsr_camera *camera) {
	draw_context ctx;
	if (!context_init(&ctx, &r->color, &r->depth, camera, model->filter)) { return; }
	if (!orc_camera_is_box_seen(camera, model->minBound, model->maxBound, m2w)) { return; }
	if (r->occluded && !orc_renderer_is_box_visible(r, model->minBound, model->maxBound, m2w, camera)) { return; }
	ctx.queue = r;
	submit_model(&ctx, model, m2w, camera);
}
void orc_renderer_set_debug_wireframe(orc_renderer *r, int enabled) { r->wireframe = enabled; }
/* api/rendererAPI.cpp:193-217 completeOcclusion + :352-402 endFrame. Returns the queue length; *occludedOut = commands skipped. */
int64_t orc_renderer_end(orc_renderer *r, int64_t *occludedOut) {
	r->receiving = 0;
	int64_t skipped = 0;
	if (r->occluded) {
		for (int64_t t = r->count - 1; t >= 0; t--) {
			queued_command *c = &r->commands[t];
			int anyVisible = 0;
			irect outer = outer_cell_bound(r, triangle_bound(c->p));
			float triangleDepth = c->p[0].cs.z;
			if (c->p[1].cs.z < triangleDepth) { triangleDepth = c->p[1].cs.z; }
			if (c->p[2].cs.z < triangleDepth) { triangleDepth = c->p[2].cs.z; }
			for (int32_t cy = outer.t; cy < outer.t + outer.h; cy++) {
				for (int32_t cx = outer.l; cx < outer.l + outer.w; cx++) {
					if ((double)triangleDepth < (double)grid_read(r, cx, cy) + 0.001) { anyVisible = 1; }
				}
			}
			if (!anyVisible) { c->occluded = 1; skipped++; }
		}
	}
	for (int64_t t = 0; t < r->count; t++) {
		queued_command *c = &r->commands[t];
		if (c->occluded) { continue; }
		draw_context ctx;
		if (!context_init(&ctx, &r->color, &r->depth, &c->cameraCopy, c->filter)) { continue; }
		ctx.shader = c->shader;
		execute_triangle(&ctx, c->p, c->subB, c->subC);
	}
	if (r->wireframe && r->color.data != NULL) { /* api/rendererAPI.cpp:362-399: white edges of every command that was not occluded */
		static const int32_t white[4] = {255, 255, 255, 255};
		for (int64_t t = 0; t < r->count; t++) {
			const queued_command *c = &r->commands[t];
			if (c->occluded) { continue; }
			for (int e = 0; e < 3; e++) {
				const ppoint *a = &c->p[e], *b = &c->p[(e + 1) % 3];
				orc_draw_line_rgba(&r->color, (int32_t)(a->fx / 256), (int32_t)(a->fy / 256), (int32_t)(b->fx / 256), (int32_t)(b->fy / 256), white);
			}
		}
	}
	r->wireframe = 0;
	if (occludedOut != NULL) { *occludedOut = skipped; }
	r->lastOccluded = skipped;
	int64_t n = r->count;
	r->count = 0;
	return n;
}

/* ------------------------------------------------------------------------------------------- depth-only path */

/* implementation/render/renderCore.cpp:343-398 */
static void draw_triangle_depth(const dfpsr_image *depth, const dfpsr_camera *c, const ppoint *p) {
	if (!is_frontfacing(p)) { return; }
	irect clip = irect_make(0, 0, depth->width, depth->height);
	irect whole = triangle_bound(p);
	if (!irect_overlaps(whole, clip)) { return; }
	irect un = irect_cut(whole, clip);
	int32_t alignedTop = (un.t / 2) * 2, alignedBottom = ((un.t + un.h + 1) / 2) * 2;
	irect bound = irect_make(un.l, alignedTop, un.w, alignedBottom - alignedTop);
	row_interval *rows = (row_interval*)malloc(sizeof(row_interval) * (size_t)bound.h);
	rasterize_triangle(p, rows, bound);
	float zero[3] = {0.0f, 0.0f, 0.0f};
	projection proj = get_projection(p, zero, zero, c->perspective);
	for (int32_t y = bound.t; y < bound.t + bound.h; y++) {
		if (y >= depth->height) { break; } /* the reference would write past an odd-height image here; targets are even */
		row_interval row = rows[y - bound.t];
		float w[3];
		projection_at(&proj, row.left, y, w);
		float value = w[0], dx = proj.dx[0];
		for (int32_t x = row.left; x < row.right; x++) {
			float *px = depth_px(depth, x, y);
			if (proj.affine) { if (value < *px) { *px = value; } } else { if (value > *px) { *px = value; } }
			value += dx;
		}
	}
	free(rows);
}

/* implementation/render/renderCore.cpp:407-443 + model/Model.cpp:164-211 */
void orc_model_render_depth(const dfpsr_model *model, const dfpsr_transform3d *m2w, const dfpsr_image *depth, const dfpsr_camera *camera) {
	if (depth == NULL || depth->data == NULL) { return; }
	if (!orc_camera_is_box_seen(camera, model->minBound, model->maxBound, m2w)) { return; }
	ppoint *projected = (ppoint*)malloc(sizeof(ppoint) * (size_t)(model->pointCount > 0 ? model->pointCount : 1));
	for (int32_t i = 0; i < model->pointCount; i++) {
		projected[i] = world_to_screen(camera, transform_point(m2w, v3_from(model->points + 3 * i)));
	}
	for (int32_t i = 0; i < model->polygonCount; i++) {
		const dfpsr_polygon *poly = model->polygons + i;
		int triangles = poly->pointIndices[3] != -1 ? 2 : 1;
		for (int t = 0; t < triangles; t++) {
			ppoint p[3] = {projected[poly->pointIndices[0]], projected[poly->pointIndices[1 + t]], projected[poly->pointIndices[2 + t]]};
			if (triangle_visibility(p, camera, 0) == VIS_HIDDEN) { continue; }
			if (triangle_visibility(p, camera, 1) == VIS_FULL) {
				draw_triangle_depth(depth, camera, p);
			} else {
				clipped_triangle ct = clip_triangle(p, camera);
				for (int k = 0; k < ct.count - 2; k++) {
					ppoint q[3] = {camera_to_screen(camera, ct.v[0].cs), camera_to_screen(camera, ct.v[1 + k].cs), camera_to_screen(camera, ct.v[2 + k].cs)};
					draw_triangle_depth(depth, camera, q);
				}
			}
		}
	}
	free(projected);
}


/* ------------------------------------------------------------------------------------------- 2D draw calls */

static uint32_t pack_bytes_ordered(uint32_t r, uint32_t g, uint32_t b, uint32_t a, int order) {
	const int *ix = packIndex[order];
	return (r << (8 * ix[0])) | (g << (8 * ix[1])) | (b << (8 * ix[2])) | (a << (8 * ix[3]));
}

/* api/imageAPI.cpp:167-185 + api/drawAPI.cpp:72-174 (whole-image rectangle) */
void orc_image_fill_rgba(const dfpsr_image *image, int32_t r, int32_t g, int32_t b, int32_t a) {
	if (image == NULL || image->data == NULL) { return; }
	uint32_t packed = pack_bytes_ordered((uint32_t)clamp_i32(0, r, 255), (uint32_t)clamp_i32(0, g, 255), (uint32_t)clamp_i32(0, b, 255), (uint32_t)clamp_i32(0, a, 255), image->packOrder);
	for (int32_t y = 0; y < image->height; y++) { for (int32_t x = 0; x < image->width; x++) { *color_px(image, x, y) = packed; } }
}
void orc_image_fill_f32(const dfpsr_image *image, float value) {
	if (image == NULL || image->data == NULL) { return; }
	for (int32_t y = 0; y < image->height; y++) { for (int32_t x = 0; x < image->width; x++) { *depth_px(image, x, y) = value; } }
}

/* api/drawAPI.cpp:330-385 ImageIntersection: the part of source placed at (left, top) that lands inside target */
typedef struct { int32_t tx, ty, sx, sy, w, h; } intersection;
static int intersect(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, intersection *out) {
	int32_t x0 = imax(left, 0), y0 = imax(top, 0);
	int32_t x1 = imin(left + source->width, target->width), y1 = imin(top + source->height, target->height);
	if (x1 <= x0 || y1 <= y0) { return 0; }
	out->tx = x0; out->ty = y0; out->sx = x0 - left; out->sy = y0 - top; out->w = x1 - x0; out->h = y1 - y0;
	return 1;
}

static uint32_t repack(uint32_t c, int sourceOrder, int targetOrder) { /* drawAPI.cpp:508-513 */
	const int *s = packIndex[sourceOrder];
	return pack_bytes_ordered((c >> (8 * s[0])) & 255u, (c >> (8 * s[1])) & 255u, (c >> (8 * s[2])) & 255u, (c >> (8 * s[3])) & 255u, targetOrder);
}

/* api/drawAPI.cpp:492-515 */
void orc_draw_copy_rgba(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top) {
	intersection is;
	if (target == NULL || source == NULL || target->data == NULL || source->data == NULL || !intersect(target, source, left, top, &is)) { return; }
	for (int32_t y = 0; y < is.h; y++) {
		for (int32_t x = 0; x < is.w; x++) {
			*color_px(target, is.tx + x, is.ty + y) = repack(*color_px(source, is.sx + x, is.sy + y), source->packOrder, target->packOrder);
		}
	}
}
/* api/drawAPI.cpp:534-539 */
void orc_draw_copy_f32(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top) {
	intersection is;
	if (target == NULL || source == NULL || target->data == NULL || source->data == NULL || !intersect(target, source, left, top, &is)) { return; }
	for (int32_t y = 0; y < is.h; y++) {
		memcpy(depth_px(target, is.tx, is.ty + y), depth_px(source, is.sx, is.sy + y), (size_t)is.w * 4);
	}
}

/* api/drawAPI.cpp:834-904 */
void orc_draw_higher(const dfpsr_image *targetHeight, const dfpsr_image *sourceHeight, const dfpsr_image *targetA, const dfpsr_image *sourceA, const dfpsr_image *targetB, const dfpsr_image *sourceB, int32_t left, int32_t top, float offset) {
	intersection is;
	if (targetHeight == NULL || sourceHeight == NULL || targetHeight->data == NULL || sourceHeight->data == NULL) { return; }
	if (!intersect(targetHeight, sourceHeight, left, top, &is)) { return; }
	int hasA = targetA != NULL && targetA->data != NULL && sourceA != NULL && sourceA->data != NULL;
	int hasB = targetB != NULL && targetB->data != NULL && sourceB != NULL && sourceB->data != NULL;
	for (int32_t y = 0; y < is.h; y++) {
		for (int32_t x = 0; x < is.w; x++) {
			float newHeight = *depth_px(sourceHeight, is.sx + x, is.sy + y);
			if (newHeight > -INFINITY) {
				newHeight += offset;
				float *t = depth_px(targetHeight, is.tx + x, is.ty + y);
				if (newHeight > *t) {
					*t = newHeight;
					if (hasA) { *color_px(targetA, is.tx + x, is.ty + y) = repack(*color_px(sourceA, is.sx + x, is.sy + y), sourceA->packOrder, targetA->packOrder); }
					if (hasB) { *color_px(targetB, is.tx + x, is.ty + y) = repack(*color_px(sourceB, is.sx + x, is.sy + y), sourceB->packOrder, targetB->packOrder); }
				}
			}
		}
	}
}

static v3 mat3_transform(const dfpsr_matrix3x3 *m, v3 p) { return mat_transform(m->xAxis, m->yAxis, m->zAxis, p); }
static v3 mat3_transform_transposed(const dfpsr_matrix3x3 *m, v3 p) { return mat_transform_transposed(m->xAxis, m->yAxis, m->zAxis, p); }

/* ------------------------------------------------------------------------------------------- remaining 2D draw calls (api/drawAPI.cpp) */

static uint32_t clamp_byte(int32_t v) { return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
static uint32_t draw_color(const dfpsr_image *image, const int32_t *c) {
	return pack_bytes_ordered(clamp_byte(c[0]), clamp_byte(c[1]), clamp_byte(c[2]), clamp_byte(c[3]), image->packOrder);
}
static void write_pixel_u32(const dfpsr_image *im, int64_t x, int64_t y, uint32_t value) { /* image_writePixel: out-of-bound writes are ignored */
	if (x >= 0 && x < im->width && y >= 0 && y < im->height) { *color_px(im, (int32_t)x, (int32_t)y) = value; }
}
/* api/drawAPI.cpp:72-88, :152-174 */
static void rectangle_u32(const dfpsr_image *im, int32_t left, int32_t top, int32_t width, int32_t height, uint32_t value) {
	if (im == NULL || im->data == NULL) { return; }
	int64_t right = (int64_t)left + width, bottom = (int64_t)top + height;
	int32_t l = left > 0 ? left : 0, t = top > 0 ? top : 0;
	int32_t r = right < im->width ? (int32_t)right : im->width, b = bottom < im->height ? (int32_t)bottom : im->height;
	for (int32_t y = t; y < b; y++) { for (int32_t x = l; x < r; x++) { *color_px(im, x, y) = value; } }
}
void orc_draw_rectangle_rgba(const dfpsr_image *image, int32_t left, int32_t top, int32_t width, int32_t height, const int32_t *colorRgba) {
	if (image == NULL || image->data == NULL) { return; }
	rectangle_u32(image, left, top, width, height, draw_color(image, colorRgba));
}
void orc_draw_rectangle_f32(const dfpsr_image *image, int32_t left, int32_t top, int32_t width, int32_t height, float value) {
	uint32_t bits; memcpy(&bits, &value, 4);
	rectangle_u32(image, left, top, width, height, bits);
}
/* api/drawAPI.cpp:176-283 drawLineSuper, with the reference's running error term */
static void line_u32(const dfpsr_image *im, int32_t x1, int32_t y1, int32_t x2, int32_t y2, uint32_t value) {
	if (im == NULL || im->data == NULL) { return; }
	int32_t width = im->width, height = im->height;
	if ((x1 < 0 && x2 < 0) || (y1 < 0 && y2 < 0) || (x1 >= width && x2 >= width) || (y1 >= height && y2 >= height)) { return; }
	if (y1 == y2) {
		int32_t l = x1 < x2 ? x1 : x2, r = x1 > x2 ? x1 : x2;
		for (int64_t x = l; x <= r; x++) { write_pixel_u32(im, x, y1, value); }
	} else if (x1 == x2) {
		int32_t t = y1 < y2 ? y1 : y2, b = y1 > y2 ? y1 : y2;
		for (int64_t y = t; y <= b; y++) { write_pixel_u32(im, x1, y, value); }
	} else {
		int64_t adx = (int64_t)x2 - x1, ady = (int64_t)y2 - y1;
		if (adx < 0) { adx = -adx; } if (ady < 0) { ady = -ady; }
		if (ady >= adx) {
			if (y2 < y1) { int32_t tx = x1, ty = y1; x1 = x2; y1 = y2; x2 = tx; y2 = ty; }
			int64_t x = x1, y = y1, tilt = adx * 2, maxError = (int64_t)y2 - y1, error = 0, step = x2 > x1 ? 1 : -1;
			while (y <= y2) {
				write_pixel_u32(im, x, y, value);
				error += tilt;
				if (error >= maxError) { x += step; error -= maxError * 2; }
				y++;
			}
		} else {
			if (x2 < x1) { int32_t tx = x1, ty = y1; x1 = x2; y1 = y2; x2 = tx; y2 = ty; }
			int64_t x = x1, y = y1, tilt = ady * 2, maxError = (int64_t)x2 - x1, error = 0, step = y2 > y1 ? 1 : -1;
			while (x <= x2) {
				write_pixel_u32(im, x, y, value);
				error += tilt;
				if (error >= maxError) { y += step; error -= maxError * 2; }
				x++;
			}
		}
	}
}
void orc_draw_line_rgba(const dfpsr_image *image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, const int32_t *colorRgba) {
	if (image == NULL || image->data == NULL) { return; }
	line_u32(image, x1, y1, x2, y2, draw_color(image, colorRgba));
}
void orc_draw_line_f32(const dfpsr_image *image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, float value) {
	uint32_t bits; memcpy(&bits, &value, 4);
	line_u32(image, x1, y1, x2, y2, bits);
}
static uint32_t byte_mul(uint32_t a, uint32_t b) { return (a * b * 65793u + 8388608u) >> 24; } /* api/drawAPI.cpp:46-54 */
static void unpack_color(uint32_t c, int order, uint32_t *v) { const int *s = packIndex[order]; for (int k = 0; k < 4; k++) { v[k] = (c >> (8 * s[k])) & 255u; } }
/* api/drawAPI.cpp:636-715: op 0 alphaFilter, 1 maxAlpha (parameter = sourceAlphaOffset), 2 alphaClip (parameter = threshold) */
static void image_over(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, int op, int32_t parameter) {
	intersection is;
	if (target == NULL || source == NULL || target->data == NULL || source->data == NULL) { return; }
	if (!intersect(target, source, left, top, &is)) { return; }
	for (int32_t y = 0; y < is.h; y++) {
		for (int32_t x = 0; x < is.w; x++) {
			uint32_t s[4], t[4];
			uint32_t *tp = color_px(target, is.tx + x, is.ty + y);
			unpack_color(*color_px(source, is.sx + x, is.sy + y), source->packOrder, s);
			unpack_color(*tp, target->packOrder, t);
			if (op == 0) {
				uint32_t sr = s[3];
				if (sr > 0) {
					if (sr == 255) { *tp = pack_bytes_ordered(s[0], s[1], s[2], 255, target->packOrder); }
					else {
						uint32_t tr = 255 - sr;
						*tp = pack_bytes_ordered((uint8_t)(byte_mul(t[0], tr) + byte_mul(s[0], sr)), (uint8_t)(byte_mul(t[1], tr) + byte_mul(s[1], sr)), (uint8_t)(byte_mul(t[2], tr) + byte_mul(s[2], sr)), (uint8_t)(byte_mul(t[3], tr) + sr), target->packOrder);
					}
				}
			} else if (op == 1) {
				int32_t sa = (int32_t)s[3];
				if (parameter == 0) {
					if (sa > (int32_t)t[3]) { *tp = pack_bytes_ordered(s[0], s[1], s[2], (uint32_t)sa, target->packOrder); }
				} else if (sa > 0) {
					sa += parameter;
					if (sa > (int32_t)t[3]) {
						if (sa < 0) { sa = 0; } if (sa > 255) { sa = 255; }
						*tp = pack_bytes_ordered(s[0], s[1], s[2], (uint32_t)sa, target->packOrder);
					}
				}
			} else {
				if ((int32_t)s[3] > parameter) { *tp = pack_bytes_ordered(s[0], s[1], s[2], 255, target->packOrder); }
			}
		}
	}
}
void orc_draw_alpha_filter(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top) { image_over(target, source, left, top, 0, 0); }
void orc_draw_max_alpha(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, int32_t sourceAlphaOffset) { image_over(target, source, left, top, 1, sourceAlphaOffset); }
void orc_draw_alpha_clip(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, int32_t threshold) { image_over(target, source, left, top, 2, threshold); }
/* api/drawAPI.cpp:717-757; the silhouette is an 8-bit image (1 byte per pixel) */
void orc_draw_silhouette(const dfpsr_image *target, const dfpsr_image *silhouetteU8, const int32_t *colorRgba, int32_t left, int32_t top) {
	intersection is;
	if (target == NULL || silhouetteU8 == NULL || target->data == NULL || silhouetteU8->data == NULL) { return; }
	if (colorRgba[3] <= 0) { return; }
	uint32_t red = clamp_byte(colorRgba[0]), green = clamp_byte(colorRgba[1]), blue = clamp_byte(colorRgba[2]), alpha = clamp_byte(colorRgba[3]);
	int fullAlpha = colorRgba[3] >= 255;
	if (!intersect(target, silhouetteU8, left, top, &is)) { return; }
	for (int32_t y = 0; y < is.h; y++) {
		for (int32_t x = 0; x < is.w; x++) {
			uint32_t sr = *((const uint8_t*)silhouetteU8->data + (size_t)(is.sy + y) * silhouetteU8->stride + (is.sx + x));
			if (!fullAlpha) { sr = byte_mul(sr, alpha); }
			if (sr == 0) { continue; }
			uint32_t *tp = color_px(target, is.tx + x, is.ty + y);
			if (sr == 255) { *tp = pack_bytes_ordered(red, green, blue, 255, target->packOrder); continue; }
			uint32_t t[4], tr = 255 - sr;
			unpack_color(*tp, target->packOrder, t);
			*tp = pack_bytes_ordered((uint8_t)(byte_mul(t[0], tr) + byte_mul(red, sr)), (uint8_t)(byte_mul(t[1], tr) + byte_mul(green, sr)), (uint8_t)(byte_mul(t[2], tr) + byte_mul(blue, sr)), (uint8_t)(byte_mul(t[3], tr) + sr), target->packOrder);
		}
	}
}

/* ---- 8-bit / 16-bit monochrome images and conversions between formats (api/drawAPI.cpp:130-150, :284-297, :476-486, :519-634, :759-832) */

static uint8_t *u8_px(const dfpsr_image *im, int32_t x, int32_t y) { return (uint8_t*)im->data + (size_t)y * im->stride + x; }
static uint16_t *u16_px(const dfpsr_image *im, int32_t x, int32_t y) { return (uint16_t*)((uint8_t*)im->data + (size_t)y * im->stride) + x; }
static void mono_write(const dfpsr_image *im, int32_t format, int64_t x, int64_t y, uint32_t value) {
	if (x < 0 || x >= im->width || y < 0 || y >= im->height) { return; }
	if (format == DFPSR_FORMAT_U8) { *u8_px(im, (int32_t)x, (int32_t)y) = (uint8_t)value; } else { *u16_px(im, (int32_t)x, (int32_t)y) = (uint16_t)value; }
}
static uint32_t mono_color(int32_t format, int32_t color) { int32_t top = format == DFPSR_FORMAT_U8 ? 255 : 65535; return (uint32_t)(color < 0 ? 0 : (color > top ? top : color)); }
void orc_draw_rectangle_mono(const dfpsr_image *image, int32_t format, int32_t left, int32_t top, int32_t width, int32_t height, int32_t color) {
	if (image == NULL || image->data == NULL) { return; }
	int64_t right = (int64_t)left + width, bottom = (int64_t)top + height;
	int32_t l = left > 0 ? left : 0, t = top > 0 ? top : 0;
	int32_t r = right < image->width ? (int32_t)right : image->width, b = bottom < image->height ? (int32_t)bottom : image->height;
	uint32_t value = mono_color(format, color);
	/* api/drawAPI.cpp:136-146: a 16-bit colour whose two bytes are equal takes the memset path, which is handed 0 instead of the byte */
	if (format == DFPSR_FORMAT_U16 && (value & 0xFFu) == (value >> 8)) { value = 0u; }
	for (int32_t y = t; y < b; y++) { for (int32_t x = l; x < r; x++) { mono_write(image, format, x, y, value); } }
}
void orc_draw_line_mono(const dfpsr_image *image, int32_t format, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int32_t color) {
	/* the same walk as line_u32 with a narrower store: drawn into a 32-bit scratch image of the same size, then copied where it was touched */
	if (image == NULL || image->data == NULL) { return; }
	dfpsr_image scratch = *image;
	uint32_t *mask = (uint32_t*)calloc((size_t)image->width * image->height, 4);
	scratch.data = mask; scratch.stride = image->width * 4;
	line_u32(&scratch, x1, y1, x2, y2, 0xFFFFFFFFu);
	for (int32_t y = 0; y < image->height; y++) { for (int32_t x = 0; x < image->width; x++) { if (mask[(size_t)y * image->width + x]) { mono_write(image, format, x, y, mono_color(format, color)); } } }
	free(mask);
}
static int32_t saturate_float(float value) { /* api/drawAPI.cpp:476-486 */
	if (!(value >= 0.5f)) { return 0; }
	if (value > 254.5f) { return 255; }
	return (uint8_t)(value + 0.5f);
}
static size_t format_size(int32_t format) { return format == DFPSR_FORMAT_U8 ? 1 : (format == DFPSR_FORMAT_U16 ? 2 : 4); }
void orc_draw_copy_formats(const dfpsr_image *target, int32_t targetFormat, const dfpsr_image *source, int32_t sourceFormat, int32_t left, int32_t top) {
	intersection is;
	if (target == NULL || source == NULL || target->data == NULL || source->data == NULL) { return; }
	if (targetFormat == DFPSR_FORMAT_RGBA_U8 && sourceFormat == DFPSR_FORMAT_RGBA_U8) { orc_draw_copy_rgba(target, source, left, top); return; }
	if (!intersect(target, source, left, top, &is)) { return; }
	for (int32_t y = 0; y < is.h; y++) {
		for (int32_t x = 0; x < is.w; x++) {
			const uint8_t *sp = (const uint8_t*)source->data + (size_t)(is.sy + y) * source->stride + (size_t)(is.sx + x) * format_size(sourceFormat);
			uint8_t *tp = (uint8_t*)target->data + (size_t)(is.ty + y) * target->stride + (size_t)(is.tx + x) * format_size(targetFormat);
			if (targetFormat == sourceFormat) { memcpy(tp, sp, format_size(targetFormat)); continue; }
			if (targetFormat == DFPSR_FORMAT_RGBA_U8) {
				int32_t luma;
				if (sourceFormat == DFPSR_FORMAT_U8) { luma = *sp; }
				else if (sourceFormat == DFPSR_FORMAT_U16) { luma = *(const uint16_t*)sp; if (luma > 255) { luma = 255; } }
				else { luma = saturate_float(*(const float*)sp); }
				*(uint32_t*)tp = pack_bytes_ordered((uint32_t)luma, (uint32_t)luma, (uint32_t)luma, 255u, target->packOrder);
			} else if (targetFormat == DFPSR_FORMAT_U8) {
				if (sourceFormat == DFPSR_FORMAT_F32) { *tp = (uint8_t)saturate_float(*(const float*)sp); }
				else { int32_t luma = *(const uint16_t*)sp; if (luma > 255) { luma = 255; } *tp = (uint8_t)luma; }
			} else if (targetFormat == DFPSR_FORMAT_U16) {
				/* from F32 the reference stores *sourcePixel, the first byte of the float (:603-614) */
				*(uint16_t*)tp = *sp;
			} else {
				if (sourceFormat == DFPSR_FORMAT_U8) { *(float*)tp = (float)*sp; }
				else { int32_t luma = *(const uint16_t*)sp; if (luma > 255) { luma = 255; } *(float*)tp = (float)luma; }
			}
		}
	}
}
void orc_draw_higher_u16(const dfpsr_image *targetHeight, const dfpsr_image *sourceHeight, const dfpsr_image *targetA, const dfpsr_image *sourceA, const dfpsr_image *targetB, const dfpsr_image *sourceB, int32_t left, int32_t top, int32_t offset) {
	intersection is;
	if (targetHeight == NULL || sourceHeight == NULL || targetHeight->data == NULL || sourceHeight->data == NULL) { return; }
	if (!intersect(targetHeight, sourceHeight, left, top, &is)) { return; }
	int hasA = targetA != NULL && targetA->data != NULL && sourceA != NULL && sourceA->data != NULL;
	int hasB = targetB != NULL && targetB->data != NULL && sourceB != NULL && sourceB->data != NULL;
	for (int32_t y = 0; y < is.h; y++) {
		for (int32_t x = 0; x < is.w; x++) {
			int32_t newHeight = *u16_px(sourceHeight, is.sx + x, is.sy + y);
			if (newHeight > 0) {
				newHeight += offset;
				if (newHeight < 0) { newHeight = 0; }
				if (newHeight > 65535) { newHeight = 65535; }
				uint16_t *t = u16_px(targetHeight, is.tx + x, is.ty + y);
				if (newHeight > 0 && newHeight > *t) {
					*t = (uint16_t)newHeight;
					if (hasA) { *color_px(targetA, is.tx + x, is.ty + y) = repack(*color_px(sourceA, is.sx + x, is.sy + y), sourceA->packOrder, targetA->packOrder); }
					if (hasB) { *color_px(targetB, is.tx + x, is.ty + y) = repack(*color_px(sourceB, is.sx + x, is.sy + y), sourceB->packOrder, targetB->packOrder); }
				}
			}
		}
	}
}

/* ------------------------------------------------------------------------------------------- Sandbox dense models and sprite heights */

/* The reference converts with cvttss2si; (uint32_t)float goes through the 64-bit conversion on x86-64. */
static int32_t trunc_i32(float v) { return (v > -2147483904.0f && v < 2147483648.0f) ? (int32_t)v : INT32_MIN; }
static uint32_t trunc_u32(float v) { return (v > -9223373136366403584.0f && v < 9223372036854775808.0f) ? (uint32_t)(int64_t)v : 0u; }

/* SDK/SpriteEngine/spriteAPI.cpp:1243-1327 renderDenseModel<HIGH_QUALITY>, one triangle after the other like the reference. */
void orc_dense_model_render(const dfpsr_dense_triangle *triangles, int32_t triangleCount, const float *minBound, const float *maxBound, const dfpsr_ortho_camera *view,
                            const dfpsr_image *height, const dfpsr_image *diffuse, const dfpsr_image *normal, const float *worldOrigin, const dfpsr_transform3d *modelToWorld,
                            int32_t highQuality, int32_t *dirtyRect) {
	/* combineModelToScreenTransform (:27-33): modelToWorld * Transform3D((ox, oy, 0), worldSpaceToScreenDepth), math/Transform3D.h:56-58 */
	const dfpsr_matrix3x3 *w2s = &view->worldSpaceToScreenDepth;
	dfpsr_transform3d o2s;
	{
		v3 p = mat3_transform(w2s, v3_from(modelToWorld->position));
		o2s.position[0] = p.x + worldOrigin[0]; o2s.position[1] = p.y + worldOrigin[1]; o2s.position[2] = p.z + 0.0f;
		v3 ax = mat3_transform(w2s, v3_from(modelToWorld->xAxis)), ay = mat3_transform(w2s, v3_from(modelToWorld->yAxis)), az = mat3_transform(w2s, v3_from(modelToWorld->zAxis));
		o2s.xAxis[0] = ax.x; o2s.xAxis[1] = ax.y; o2s.xAxis[2] = ax.z; o2s.yAxis[0] = ay.x; o2s.yAxis[1] = ay.y; o2s.yAxis[2] = ay.z; o2s.zAxis[0] = az.x; o2s.zAxis[1] = az.y; o2s.zAxis[2] = az.z;
	}
	/* boundingBoxToRectangle (:1132-1141) */
	int32_t bl = 0, bt = 0, br = 0, bb = 0;
	for (int c = 0; c < 8; c++) {
		v3 p = transform_point(&o2s, v3_make((c & 1) ? maxBound[0] : minBound[0], (c & 2) ? maxBound[1] : minBound[1], (c & 4) ? maxBound[2] : minBound[2]));
		int32_t x = trunc_i32(p.x), y = trunc_i32(p.y);
		if (c == 0) { bl = x; bt = y; br = x + 1; bb = y + 1; }
		else { if (x < bl) { bl = x; } if (y < bt) { bt = y; } if (x + 1 > br) { br = x + 1; } if (y + 1 > bb) { bb = y + 1; } }
	}
	const int32_t cw = height->width, ch = height->height;
	if (dirtyRect) { dirtyRect[0] = dirtyRect[1] = dirtyRect[2] = dirtyRect[3] = 0; }
	if (!(bl < cw && br > 0 && bt < ch && bb > 0)) { return; }
	if (dirtyRect) { dirtyRect[0] = bl; dirtyRect[1] = bt; dirtyRect[2] = br - bl; dirtyRect[3] = bb - bt; }
	/* modelToNormalSpace = modelToWorld.transform * transpose(normalToWorldSpace) (:1262, math/FMatrix3x3.h:78-80) */
	const dfpsr_matrix3x3 *n2w = &view->normalToWorldSpace;
	dfpsr_matrix3x3 nt, m2n;
	nt.xAxis[0] = n2w->xAxis[0]; nt.xAxis[1] = n2w->yAxis[0]; nt.xAxis[2] = n2w->zAxis[0];
	nt.yAxis[0] = n2w->xAxis[1]; nt.yAxis[1] = n2w->yAxis[1]; nt.yAxis[2] = n2w->zAxis[1];
	nt.zAxis[0] = n2w->xAxis[2]; nt.zAxis[1] = n2w->yAxis[2]; nt.zAxis[2] = n2w->zAxis[2];
	{
		v3 ax = mat3_transform(&nt, v3_from(modelToWorld->xAxis)), ay = mat3_transform(&nt, v3_from(modelToWorld->yAxis)), az = mat3_transform(&nt, v3_from(modelToWorld->zAxis));
		m2n.xAxis[0] = ax.x; m2n.xAxis[1] = ax.y; m2n.xAxis[2] = ax.z; m2n.yAxis[0] = ay.x; m2n.yAxis[1] = ay.y; m2n.yAxis[2] = ay.z; m2n.zAxis[0] = az.x; m2n.zAxis[1] = az.y; m2n.zAxis[2] = az.z;
	}
	for (int32_t i = 0; i < triangleCount; i++) {
		const dfpsr_dense_triangle *tri = triangles + i;
		v3 a = transform_point(&o2s, v3_from(tri->posA)), b = transform_point(&o2s, v3_from(tri->posB)), c = transform_point(&o2s, v3_from(tri->posC));
		/* getBackCulledTriangleBound (:1143-1156) */
		if (((c.x - a.x) * (b.y - a.y)) + ((c.y - a.y) * (a.x - b.x)) >= 0.0f) { continue; }
		float minX = a.x < b.x ? a.x : b.x; minX = minX < c.x ? minX : c.x;
		float minY = a.y < b.y ? a.y : b.y; minY = minY < c.y ? minY : c.y;
		float maxX = a.x > b.x ? a.x : b.x; maxX = maxX > c.x ? maxX : c.x;
		float maxY = a.y > b.y ? a.y : b.y; maxY = maxY > c.y ? maxY : c.y;
		int32_t l = trunc_i32(minX), t = trunc_i32(minY), r = trunc_i32(maxX) + 1, bo = trunc_i32(maxY) + 1;
		if (!(l < cw && r > 0 && t < ch && bo > 0)) { continue; } /* IRect::cut, math/IRect.h:56-66 */
		if (l < 0) { l = 0; } if (t < 0) { t = 0; } if (r > cw) { r = cw; } if (bo > ch) { bo = ch; }
		if (!(r > l && bo > t)) { continue; }
		/* inverse(FMatrix2x2(B - A, C - A)), math/FMatrix2x2.h:69-76 */
		const float xx = b.x - a.x, xy = b.y - a.y, yx = c.x - a.x, yy = c.y - a.y;
		const float inv = 1.0f / (xx * yy - xy * yx);
		const float m0 = yy * inv, m1 = -xy * inv, m2 = -yx * inv, m3 = xx * inv;
		v3 na = mat3_transform(&m2n, v3_from(tri->normalA)), nb = mat3_transform(&m2n, v3_from(tri->normalB)), nc = mat3_transform(&m2n, v3_from(tri->normalC));
		for (int32_t y = t; y < bo; y++) {
			for (int32_t x = l; x < r; x++) {
				const float ox = ((float)x + 0.5f) - a.x, oy = ((float)y + 0.5f) - a.y;
				const float wb = ox * m0 + oy * m2, wc = ox * m1 + oy * m3;
				const float wa = 1.0f - (wb + wc);
				if (wa >= -0.00001f && wb >= -0.00001f && wc >= -0.00001f) {
					const float h = a.z * wa + b.z * wb + c.z * wc;
					float *hp = depth_px(height, x, y);
					if (h > *hp) {
						*hp = h;
						const float cr = tri->colorA[0] * wa + tri->colorB[0] * wb + tri->colorC[0] * wc;
						const float cg = tri->colorA[1] * wa + tri->colorB[1] * wb + tri->colorC[1] * wc;
						const float cb = tri->colorA[2] * wa + tri->colorB[2] * wb + tri->colorC[2] * wc;
						*color_px(diffuse, x, y) = trunc_u32(cr) | (trunc_u32(cg) << 8) | (trunc_u32(cb) << 16) | (255u << 24);
						v3 n = v3_make(na.x * wa + nb.x * wb + nc.x * wc, na.y * wa + nb.y * wb + nc.y * wc, na.z * wa + nb.z * wb + nc.z * wc);
						if (highQuality) { n = v3_normalize(n); }
						*color_px(normal, x, y) = trunc_u32((n.x + 1.0f) * 127.5f) | (trunc_u32((n.y + 1.0f) * 127.5f) << 8) | (trunc_u32((n.z + 1.0f) * 127.5f) << 16) | (255u << 24);
					}
				}
			}
		}
	}
}

/* SDK/SpriteEngine/spriteAPI.cpp:157-174 scaleHeightImage: heights of one sprite frame from the atlas' height column (RGBA order). */
void orc_sprite_scale_height(const dfpsr_image *heightColumn, const dfpsr_image *colorColumn, float minHeight, float maxHeight, const dfpsr_image *out) {
	const float scale = (maxHeight - minHeight) / 255.0f, offset = minHeight;
	for (int32_t y = 0; y < out->height; y++) {
		for (int32_t x = 0; x < out->width; x++) {
			const float value = (float)(*color_px(heightColumn, x, y) & 255u);
			*depth_px(out, x, y) = ((*color_px(colorColumn, x, y) >> 24) > 127u) ? (value * scale) + offset : -INFINITY;
		}
	}
}

/* ------------------------------------------------------------------------------------------- Sandbox light */


static uint8_t sat_add_u8(uint8_t a, uint8_t b) { uint32_t s = (uint32_t)a + (uint32_t)b; return (uint8_t)(s > 255u ? 255u : s); } /* base/simd.h:2670 */

static uint32_t light_pack(float red, float green, float blue) { /* lightAPI.cpp:52-56 */
	return saturated_byte(red) | (saturated_byte(green) << 8) | (saturated_byte(blue) << 16);
}
static uint32_t sat_add_packed(uint32_t a, uint32_t b) {
	uint32_t r = 0;
	for (int s = 0; s < 32; s += 8) { r |= (uint32_t)sat_add_u8((uint8_t)(a >> s), (uint8_t)(b >> s)) << s; }
	return r;
}

/* SDK/SpriteEngine/lightAPI.cpp:23-68 */
void orc_light_directed(const dfpsr_ortho_view *view, const dfpsr_image *light, const dfpsr_image *normal, const float *direction, float intensity, const int32_t *colorRgb, int32_t add) {
	v3 n = v3_normalize(mat3_transform_transposed(&view->normalToWorldSpace, v3_from(direction)));
	v3 rev = v3_make(-n.x * intensity * 2.0f, -n.y * intensity * 2.0f, -n.z * intensity * 2.0f);
	float colorR = fmaxf(0.0f, (float)colorRgb[0] / 255.0f), colorG = fmaxf(0.0f, (float)colorRgb[1] / 255.0f), colorB = fmaxf(0.0f, (float)colorRgb[2] / 255.0f);
	for (int32_t y = 0; y < light->height; y++) {
		for (int32_t x = 0; x < light->width; x++) {
			uint32_t nc = *color_px(normal, x, y);
			float nx = (float)(nc & 255u) - 128.0f, ny = (float)((nc >> 8) & 255u) - 128.0f, nz = (float)((nc >> 16) & 255u) - 128.0f;
			float dot = (nx * rev.x) + (ny * rev.y) + (nz * rev.z);
			float in = dot > 0.0f ? dot : 0.0f;
			uint32_t packed = light_pack(in * colorR, in * colorG, in * colorB);
			uint32_t *t = color_px(light, x, y);
			*t = add ? sat_add_packed(*t, packed) : packed;
		}
	}
}

/* SDK/SpriteEngine/lightAPI.cpp:107-139 */
static float shadow_transparency(const dfpsr_image *cube, float halfWidth, v3 o) {
	int32_t width = cube->width;
	float absX = o.x < 0.0f ? -o.x : o.x, absY = o.y < 0.0f ? -o.y : o.y, absZ = o.z < 0.0f ? -o.z : o.z;
	int xIsLongest = absX > absY && absX > absZ;
	int yIsLongerThanZ = absY > absZ;
	float depth = xIsLongest ? o.x : (yIsLongerThanZ ? o.y : o.z);
	float slopeUp = (yIsLongerThanZ && !xIsLongest) ? o.z : o.y;
	float slopeSide = xIsLongest ? -o.z : (yIsLongerThanZ ? -o.x : o.x);
	int32_t viewOffset = width * (xIsLongest ? 0 : (yIsLongerThanZ ? 2 : 4));
	int negativeSide = depth < 0.0f;
	if (negativeSide) { depth = -depth; slopeSide = -slopeSide; viewOffset = viewOffset + width; }
	float reciDepth = 1.0f / depth;
	float scale = halfWidth * reciDepth;
	int32_t sampleX = (int32_t)(halfWidth + (slopeSide * scale));
	int32_t sampleY = (int32_t)(halfWidth - (slopeUp * scale));
	int32_t maxPixel = width - 1;
	sampleX = clamp_i32(0, sampleX, maxPixel);
	sampleY = clamp_i32(0, sampleY, maxPixel);
	float shadowReciDepth = *depth_px(cube, sampleX, sampleY + viewOffset);
	return reciDepth * 1.02f > shadowReciDepth ? 1.0f : 0.0f;
}

/* SDK/SpriteEngine/lightAPI.cpp:76-105 calculateBound; returns 0 when empty */
static int light_bound(const dfpsr_ortho_view *view, const int32_t *worldCenter, const dfpsr_image *light, v3 lightSpacePosition, float radius, int32_t alignment, irect *out) {
	v3 rotated = mat3_transform(&view->lightSpaceToScreenDepth, lightSpacePosition);
	int32_t cx = (int32_t)rotated.x + worldCenter[0], cy = (int32_t)rotated.y + worldCenter[1];
	int32_t pixelRadius = (int32_t)(radius * view->lightSpaceToScreenDepth.xAxis[0]);
	if (cx < -pixelRadius || cx > light->width + pixelRadius || cy < -pixelRadius || cy > light->height + pixelRadius) { return 0; }
	int32_t size = (int32_t)((float)pixelRadius * 2.0f);
	irect r = irect_cut(irect_make(0, 0, light->width, light->height), irect_make(cx - pixelRadius, cy - pixelRadius, size, size));
	if (r.w > 0 && r.h > 0 && alignment > 1) {
		int32_t left = (r.l / alignment) * alignment; /* non-negative */
		int32_t right = ((r.l + r.w + alignment - 1) / alignment) * alignment;
		r = irect_make(left, r.t, right - left, r.h);
	}
	*out = r;
	return r.w > 0 && r.h > 0;
}

/* SDK/SpriteEngine/lightAPI.cpp:170-273, single job (DISABLE_MULTI_THREADING) */
void orc_light_point(const dfpsr_ortho_view *view, const int32_t *worldCenter, const dfpsr_image *light, const dfpsr_image *normal, const dfpsr_image *height, const float *position, float radius, float intensity, const int32_t *colorRgb, const dfpsr_image *shadowCubeMap, int32_t laneCount) {
	v3 S = mat3_transform_transposed(&view->normalToWorldSpace, v3_from(position));
	irect bound;
	if (!light_bound(view, worldCenter, light, S, radius, laneCount, &bound)) { return; }
	v3 face = v3_from(view->screenDepthToLightSpace.zAxis);
	float colorR = fmaxf(0.0f, (float)colorRgb[0] * intensity), colorG = fmaxf(0.0f, (float)colorRgb[1] * intensity), colorB = fmaxf(0.0f, (float)colorRgb[2] * intensity);
	float reciprocalRadius = 1.0f / radius;
	int shadow = shadowCubeMap != NULL && shadowCubeMap->data != NULL;
	float cubeCenter = shadow ? (float)shadowCubeMap->width * 0.5f : 0.0f;
	v3 t = mat3_transform(&view->screenDepthToLightSpace, v3_make(0.5f - (float)worldCenter[0] + (float)bound.l, 0.5f - (float)worldCenter[1] + (float)bound.t, 0.0f));
	v3 base = v3_make(t.x - S.x, t.y - S.y, t.z - S.z);
	v3 dx = v3_from(view->screenDepthToLightSpace.xAxis), dy = v3_from(view->screenDepthToLightSpace.yAxis);
	/* createGradient (base/simd.h:474): lane l = start + increment * l (l = 2, 3, ... multiply first) */
	float rowX[8], rowY[8], rowZ[8];
	for (int l = 0; l < laneCount; l++) {
		if (l == 0) { rowX[l] = base.x; rowY[l] = base.y; rowZ[l] = base.z; }
		else if (l == 1) { rowX[l] = base.x + dx.x; rowY[l] = base.y + dx.y; rowZ[l] = base.z + dx.z; }
		else { rowX[l] = base.x + dx.x * (float)l; rowY[l] = base.y + dx.y * (float)l; rowZ[l] = base.z + dx.z * (float)l; }
	}
	v3 dxX = v3_make(dx.x * (float)laneCount, dx.y * (float)laneCount, dx.z * (float)laneCount);
	for (int32_t y = bound.t; y < bound.t + bound.h; y++) {
		float px[8], py[8], pz[8];
		for (int l = 0; l < laneCount; l++) { px[l] = rowX[l]; py[l] = rowY[l]; pz[l] = rowZ[l]; }
		for (int32_t x = bound.l; x < bound.l + bound.w; x += laneCount) {
			for (int l = 0; l < laneCount; l++) {
				int32_t xx = x + l;
				if (xx < light->width) { /* lanes past the width land in row padding in the reference */
					float h = *depth_px(height, xx, y);
					v3 o = v3_make(px[l] + (face.x * h), py[l] + (face.y * h), pz[l] + (face.z * h));
					float sq = (o.x * o.x) + (o.y * o.y) + (o.z * o.z);
					float len = sqrtf(sq);
					float lightRatio = len * reciprocalRadius;
					if (1.0f < lightRatio) { lightRatio = 1.0f; } /* min(1.0f, x): base/simd.h:2198-2215 */
					uint32_t nc = *color_px(normal, xx, y);
					float nx = ((float)(nc & 255u) - 128.0f) * (-1.0f / 128.0f), ny = ((float)((nc >> 8) & 255u) - 128.0f) * (-1.0f / 128.0f), nz = ((float)((nc >> 16) & 255u) - 128.0f) * (-1.0f / 128.0f);
					float distanceIntensity = 1.0f - 2.0f * lightRatio + lightRatio * lightRatio;
					float rs = ORC_RSQRT(sq);
					float dot = ((o.x * rs) * nx) + ((o.y * rs) * ny) + ((o.z * rs) * nz);
					float angleIntensity = dot > 0.0f ? dot : 0.0f;
					float in = angleIntensity * distanceIntensity;
					if (shadow) { in = in * shadow_transparency(shadowCubeMap, cubeCenter, o); }
					uint32_t packed = light_pack(in * colorR, in * colorG, in * colorB);
					uint32_t *target = color_px(light, xx, y);
					*target = sat_add_packed(*target, packed);
				}
				px[l] += dxX.x; py[l] += dxX.y; pz[l] += dxX.z;
			}
		}
		for (int l = 0; l < laneCount; l++) { rowX[l] += dy.x; rowY[l] += dy.y; rowZ[l] += dy.z; }
	}
}

/* SDK/SpriteEngine/lightAPI.cpp:287-323 */
void orc_light_blend(const dfpsr_image *color, const dfpsr_image *diffuse, const dfpsr_image *light) {
	const float scale = (float)(1.0 / 128.0f);
	for (int32_t y = 0; y < color->height; y++) {
		for (int32_t x = 0; x < color->width; x++) {
			uint32_t d = *color_px(diffuse, x, y), l = *color_px(light, x, y);
			float red = ((float)(d & 255u) * (float)(l & 255u)) * scale;
			float green = ((float)((d >> 8) & 255u) * (float)((l >> 8) & 255u)) * scale;
			float blue = ((float)((d >> 16) & 255u) * (float)((l >> 16) & 255u)) * scale;
			*color_px(color, x, y) = pack_bytes_ordered(saturated_byte(red), saturated_byte(green), saturated_byte(blue), 0u, color->packOrder);
		}
	}
}

/* ------------------------------------------------------------------------------------------- filters */

typedef struct { int32_t r, g, b, a; } rgba_i32;

static rgba_i32 read_clamp(const dfpsr_image *im, int32_t x, int32_t y) { /* api/imageAPI.h:284-290 */
	uint32_t c = *color_px(im, clamp_i32(0, x, im->width - 1), clamp_i32(0, y, im->height - 1));
	const int *ix = packIndex[im->packOrder];
	rgba_i32 o = {(int32_t)((c >> (8 * ix[0])) & 255u), (int32_t)((c >> (8 * ix[1])) & 255u), (int32_t)((c >> (8 * ix[2])) & 255u), (int32_t)((c >> (8 * ix[3])) & 255u)};
	return o;
}
static uint32_t saturate_and_pack(rgba_i32 c, int order) { /* PackOrder.h:108-110 */
	return pack_bytes_ordered((uint32_t)clamp_i32(0, c.r, 255), (uint32_t)clamp_i32(0, c.g, 255), (uint32_t)clamp_i32(0, c.b, 255), (uint32_t)clamp_i32(0, c.a, 255), order);
}
static rgba_i32 lerp16(rgba_i32 a, rgba_i32 b, uint32_t ratioB) { /* filterAPI.cpp:86-88: (a * (65536 - w) + b * w) >> 16 in u32 */
	uint32_t ra = 65536u - ratioB;
	rgba_i32 o = {(int32_t)(((uint32_t)a.r * ra + (uint32_t)b.r * ratioB) >> 16), (int32_t)(((uint32_t)a.g * ra + (uint32_t)b.g * ratioB) >> 16),
	              (int32_t)(((uint32_t)a.b * ra + (uint32_t)b.b * ratioB) >> 16), (int32_t)(((uint32_t)a.a * ra + (uint32_t)b.a * ratioB) >> 16)};
	return o;
}
static uint32_t mix_colors_uniform(uint32_t a, uint32_t b, uint32_t fineRatio) { /* filterAPI.cpp:49-63 */
	uint32_t ratio = (fineRatio >> 8) & 0xFFFFu, inv = (256u - ratio) & 0xFFFFu;
	return weight_colors(a, inv, b, ratio);
}

/* api/filterAPI.cpp:156-259 resize_optimized + :118-154 resize_reference, scaleRegion = whole target */
static void resize_single(const dfpsr_image *target, const dfpsr_image *source, int bilinear, int simdAligned) {
	int32_t tw = target->width, th = target->height, sw = source->width, sh = source->height;
	int sameWidth = sw == tw, sameHeight = sh == th, samePack = target->packOrder == source->packOrder;
	if (sameWidth && sameHeight) { orc_draw_copy_rgba(target, source, 0, 0); return; }
	int32_t offsetX = (int32_t)(65536u * (uint32_t)sw / (uint32_t)tw), offsetY = (int32_t)(65536u * (uint32_t)sh / (uint32_t)th);
	int32_t startX = offsetX / 2, startY = offsetY / 2;
	if (bilinear) { startX -= 32768; startY -= 32768; }
	if (sameWidth && (samePack || bilinear)) {
		int32_t readY = startY;
		for (int32_t y = 0; y < th; y++) {
			uint32_t sampleY = (uint32_t)(readY < 0 ? 0 : readY);
			uint32_t upperY = sampleY >> 16, lowerY = upperY + 1;
			if (upperY >= (uint32_t)sh) { upperY = (uint32_t)sh - 1; }
			if (lowerY >= (uint32_t)sh) { lowerY = (uint32_t)sh - 1; }
			uint32_t lowerRatio = sampleY & 65535u;
			for (int32_t x = 0; x < tw; x++) {
				if (bilinear) {
					if (simdAligned) {
						*color_px(target, x, y) = mix_colors_uniform(*color_px(source, x, (int32_t)upperY), *color_px(source, x, (int32_t)lowerY), lowerRatio);
					} else {
						*color_px(target, x, y) = saturate_and_pack(lerp16(read_clamp(source, x, (int32_t)upperY), read_clamp(source, x, (int32_t)lowerY), lowerRatio), target->packOrder);
					}
				} else {
					*color_px(target, x, y) = *color_px(source, x, (int32_t)upperY);
				}
			}
			readY += offsetY;
		}
	} else if (sameHeight) {
		for (int32_t y = 0; y < th; y++) {
			int32_t readX = startX;
			for (int32_t x = 0; x < tw; x++) {
				uint32_t sampleX = (uint32_t)(readX < 0 ? 0 : readX);
				uint32_t leftX = sampleX >> 16, rightRatio = sampleX & 65535u;
				rgba_i32 c = bilinear ? lerp16(read_clamp(source, (int32_t)leftX, y), read_clamp(source, (int32_t)leftX + 1, y), rightRatio) : read_clamp(source, (int32_t)leftX, y);
				*color_px(target, x, y) = saturate_and_pack(c, target->packOrder);
				readX += offsetX;
			}
		}
	} else {
		int32_t readY = startY;
		for (int32_t y = 0; y < th; y++) {
			uint32_t sampleY = (uint32_t)(readY < 0 ? 0 : readY);
			uint32_t upperY = sampleY >> 16, lowerRatio = sampleY & 65535u;
			int32_t readX = startX;
			for (int32_t x = 0; x < tw; x++) {
				uint32_t sampleX = (uint32_t)(readX < 0 ? 0 : readX);
				uint32_t leftX = sampleX >> 16, rightRatio = sampleX & 65535u;
				rgba_i32 c;
				if (bilinear) { /* filterAPI.cpp:75-93 samplePixel */
					rgba_i32 upper = lerp16(read_clamp(source, (int32_t)leftX, (int32_t)upperY), read_clamp(source, (int32_t)leftX + 1, (int32_t)upperY), rightRatio);
					rgba_i32 lower = lerp16(read_clamp(source, (int32_t)leftX, (int32_t)upperY + 1), read_clamp(source, (int32_t)leftX + 1, (int32_t)upperY + 1), rightRatio);
					c = lerp16(upper, lower, lowerRatio);
				} else {
					c = read_clamp(source, (int32_t)leftX, (int32_t)upperY);
				}
				*color_px(target, x, y) = saturate_and_pack(c, target->packOrder);
				readX += offsetX;
			}
			readY += offsetY;
		}
	}
}

/* api/filterAPI.cpp:262-314, :852-860 */
void orc_filter_resize(const dfpsr_image *target, const dfpsr_image *source, int32_t sampler, int32_t sourceIsSubImage, uint32_t *scratch) {
	int bilinear = sampler == DFPSR_SAMPLER_LINEAR;
	if (target->width != source->width && target->height > source->height) {
		dfpsr_image temp = {scratch, target->width, source->height, target->width * 4, target->packOrder};
		resize_single(&temp, source, bilinear, !sourceIsSubImage);
		resize_single(target, &temp, bilinear, 1);
	} else {
		resize_single(target, source, bilinear, !sourceIsSubImage);
	}
}

/* api/filterAPI.cpp:95-110 samplePixel(ImageU8) inside :118-154 resize_reference<*, ImageU8, uint8_t>, scaleRegion = whole target */
static uint32_t u8_clamp(const dfpsr_image *im, int32_t x, int32_t y) { /* api/imageAPI.h image_readPixel_clamp(ImageU8) */
	if (x < 0) { x = 0; } if (x >= im->width) { x = im->width - 1; }
	if (y < 0) { y = 0; } if (y >= im->height) { y = im->height - 1; }
	return ((const uint8_t *)im->data)[(size_t)y * (size_t)im->stride + (size_t)x];
}
static void resize_u8_single(const dfpsr_image *target, const dfpsr_image *source, int bilinear) {
	int32_t offsetX = (int32_t)(65536u * (uint32_t)source->width / (uint32_t)target->width), offsetY = (int32_t)(65536u * (uint32_t)source->height / (uint32_t)target->height);
	int32_t startX = offsetX / 2, startY = offsetY / 2;
	if (bilinear) { startX -= 32768; startY -= 32768; }
	int32_t readY = startY;
	for (int32_t y = 0; y < target->height; y++) {
		uint32_t sampleY = (uint32_t)(readY < 0 ? 0 : readY);
		int32_t upperY = (int32_t)(sampleY >> 16);
		uint32_t lowerRatio = sampleY & 65535u;
		int32_t readX = startX;
		for (int32_t x = 0; x < target->width; x++) {
			uint32_t sampleX = (uint32_t)(readX < 0 ? 0 : readX);
			int32_t leftX = (int32_t)(sampleX >> 16);
			uint32_t rightRatio = sampleX & 65535u, value;
			if (bilinear) {
				uint32_t upper = (u8_clamp(source, leftX, upperY) * (65536u - rightRatio) + u8_clamp(source, leftX + 1, upperY) * rightRatio) >> 16;
				uint32_t lower = (u8_clamp(source, leftX, upperY + 1) * (65536u - rightRatio) + u8_clamp(source, leftX + 1, upperY + 1) * rightRatio) >> 16;
				value = (upper * (65536u - lowerRatio) + lower * lowerRatio) >> 16;
			} else {
				value = u8_clamp(source, leftX, upperY);
			}
			((uint8_t *)target->data)[(size_t)y * (size_t)target->stride + (size_t)x] = (uint8_t)value;
			readX += offsetX;
		}
		readY += offsetY;
	}
}

/* api/filterAPI.cpp:282-314, :862-870 filter_resize(ImageU8): two passes when the width changes and the height grows */
void orc_filter_resize_u8(const dfpsr_image *target, const dfpsr_image *source, int32_t sampler, uint8_t *scratch) {
	int bilinear = sampler == DFPSR_SAMPLER_LINEAR;
	if (target->width != source->width && target->height > source->height) {
		dfpsr_image temp = {scratch, target->width, source->height, target->width, 0};
		resize_u8_single(&temp, source, bilinear);
		resize_u8_single(target, &temp, bilinear);
	} else {
		resize_u8_single(target, source, bilinear);
	}
}

/* api/filterAPI.cpp:759-777 with the enumerated ops of dfpsr_b200.h */
void orc_filter_map(const dfpsr_image *target, int32_t op, const int32_t *params, const dfpsr_image *source, int32_t startX, int32_t startY) {
	if (target == NULL || target->data == NULL) { return; }
	for (int32_t ty = 0; ty < target->height; ty++) {
		for (int32_t tx = 0; tx < target->width; tx++) {
			int32_t x = tx + startX, y = ty + startY;
			rgba_i32 c = {0, 0, 0, 0};
			if (op == DFPSR_MAP_XOR_PATTERN) { c.r = x & 255; c.g = y & 255; c.b = (x ^ y) & 255; c.a = 255; }
			else if (op == DFPSR_MAP_AFFINE) {
				rgba_i32 s = read_clamp(source, x, y);
				c.r = s.r * params[0] + params[4]; c.g = s.g * params[1] + params[5]; c.b = s.b * params[2] + params[6]; c.a = s.a * params[3] + params[7];
			} else if (op == DFPSR_MAP_CONSTANT) { c.r = params[0]; c.g = params[1]; c.b = params[2]; c.a = params[3]; }
			*color_px(target, tx, ty) = saturate_and_pack(c, target->packOrder);
		}
	}
}

/* api/filterAPI.cpp:724-757 + :340-365 */
void orc_filter_block_magnify(const dfpsr_image *target, const dfpsr_image *source, int32_t pixelWidth, int32_t pixelHeight) {
	if (target == NULL || source == NULL || target->data == NULL || source->data == NULL) { return; }
	if (pixelWidth < 1) { pixelWidth = 1; }
	if (pixelHeight < 1) { pixelHeight = 1; }
	int32_t clipWidth = imin(target->width, source->width * pixelWidth); clipWidth -= clipWidth % pixelWidth;
	int32_t clipHeight = imin(target->height, source->height * pixelHeight); clipHeight -= clipHeight % pixelHeight;
	for (int32_t y = 0; y < target->height; y++) {
		for (int32_t x = 0; x < target->width; x++) {
			uint32_t c = 0;
			if (x < clipWidth && y < clipHeight) {
				c = repack(*color_px(source, imin(x / pixelWidth, source->width - 1), imin(y / pixelHeight, source->height - 1)), source->packOrder, target->packOrder);
			}
			*color_px(target, x, y) = c;
		}
	}
}
