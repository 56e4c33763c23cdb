// pixel_ops.cu — the bandwidth-bound per-pixel passes of the hot path on sm_100a:
//   image_fill / draw_copy / draw_higher          ref: api/imageAPI.cpp:167-185, api/drawAPI.cpp:72-174, :492-539, :834-904
//   directed light / point light / blend          ref: SDK/SpriteEngine/lightAPI.cpp:23-323
//   filter_resize / filter_map / blockMagnify     ref: api/filterAPI.cpp:49-314, :724-782
//   texture pyramid                               ref: api/textureAPI.cpp:44-110
// Every kernel streams each pixel once with 16-byte accesses where rows are 16-byte aligned (4 pixels per
// thread), otherwise 4-byte accesses; the arithmetic is the reference's, operation for operation
// (-fmad=false), so integer passes are bit-exact and float passes match the reference's scalar build.
#include "common.cuh"

#include <math.h>
#include <vector>

namespace dfpsr {

static const int PX = 4; // pixels per thread along x
static inline dim3 grid_for(int32_t width, int32_t height, dim3 block) {
	return dim3((unsigned)((width + PX * (int)block.x - 1) / (PX * (int)block.x)), (unsigned)((height + (int)block.y - 1) / (int)block.y));
}
static const dim3 BLOCK(64, 4);
// Streaming kernels: every thread owns 4 pixels (16 bytes) in each of ROWS rows and issues all its loads before the first store, so a
// CTA of 256 threads keeps 16 KB per input in flight — what HBM3e needs to approach its peak (4 KB per CTA reached 56 % of it).
static const int ROWS = 4;
static inline dim3 grid_rows(int32_t width, int32_t height, dim3 block) {
	return dim3((unsigned)((width + PX * (int)block.x - 1) / (PX * (int)block.x)), (unsigned)((height + ROWS * (int)block.y - 1) / (ROWS * (int)block.y)));
}

struct Img {
	uint8_t *data;
	int32_t width, height, stride, packOrder;
};
static inline Img img_of(const dfpsr_image *im) {
	Img r;
	if (im == nullptr) { r.data = nullptr; r.width = r.height = r.stride = r.packOrder = 0; return r; }
	r.data = (uint8_t *)im->data; r.width = im->width; r.height = im->height; r.stride = im->stride; r.packOrder = im->packOrder;
	return r;
}
__device__ __forceinline__ uint32_t *px_u32(const Img &im, int32_t x, int32_t y) { return (uint32_t *)(im.data + (size_t)y * (size_t)im.stride) + x; }
__device__ __forceinline__ float *px_f32(const Img &im, int32_t x, int32_t y) { return (float *)(im.data + (size_t)y * (size_t)im.stride) + x; }
__device__ __forceinline__ bool aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

// Loads / stores up to 4 consecutive 32-bit pixels starting at x (x is a multiple of 4 in image space).
__device__ __forceinline__ void load4(const Img &im, int32_t x, int32_t y, int n, uint32_t *v) {
	const uint32_t *p = px_u32(im, x, y);
	if (n == 4 && aligned16(p)) { uint4 q = *(const uint4 *)p; v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
	else { for (int i = 0; i < 4; i++) { v[i] = i < n ? p[i] : 0u; } }
}
__device__ __forceinline__ void store4(const Img &im, int32_t x, int32_t y, int n, const uint32_t *v) {
	uint32_t *p = px_u32(im, x, y);
	if (n == 4 && aligned16(p)) { *(uint4 *)p = make_uint4(v[0], v[1], v[2], v[3]); }
	else { for (int i = 0; i < n; i++) { p[i] = v[i]; } }
}

__device__ __forceinline__ uint32_t repack(uint32_t c, uint32_t sourceShifts, uint32_t targetShifts) {
	uint32_t r = (c >> (sourceShifts & 31u)) & 255u, g = (c >> ((sourceShifts >> 8) & 31u)) & 255u;
	uint32_t b = (c >> ((sourceShifts >> 16) & 31u)) & 255u, a = (c >> ((sourceShifts >> 24) & 31u)) & 255u;
	return pack_rgba_ordered(r, g, b, a, targetShifts);
}

// ------------------------------------------------------------------------------------------------ fill / copy / higher

__global__ void __launch_bounds__(256) fill_kernel(Img target, uint32_t value) {
	int32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * PX, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= target.width || y >= target.height) { return; }
	uint32_t v[4] = {value, value, value, value};
	store4(target, x, y, min(PX, target.width - x), v);
}

struct Intersection { int32_t tx, ty, sx, sy, w, h; };
// ref: api/drawAPI.cpp:330-385 ImageIntersection
static bool intersect(const Img &target, const Img &source, int32_t left, int32_t top, Intersection &out) {
	int32_t x0 = left > 0 ? left : 0, y0 = top > 0 ? top : 0;
	int32_t x1 = left + source.width < target.width ? left + source.width : target.width;
	int32_t y1 = top + source.height < target.height ? top + source.height : target.height;
	if (x1 <= x0 || y1 <= y0) { return false; }
	out.tx = x0; out.ty = y0; out.sx = x0 - left; out.sy = y0 - top; out.w = x1 - x0; out.h = y1 - y0;
	return true;
}

// convert: 0 = raw 32-bit copy, 1 = repack channels between pack orders
__global__ void __launch_bounds__(256) copy_kernel(Img target, Img source, Intersection is, int convert) {
	int32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * PX, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= is.w || y >= is.h) { return; }
	int n = min(PX, is.w - x);
	uint32_t ss = pack_shifts(source.packOrder), ts = pack_shifts(target.packOrder);
	const uint32_t *s = px_u32(source, is.sx + x, is.sy + y);
	uint32_t *t = px_u32(target, is.tx + x, is.ty + y);
	if (n == 4 && aligned16(s) && aligned16(t)) {
		uint4 q = *(const uint4 *)s;
		if (convert) { q.x = repack(q.x, ss, ts); q.y = repack(q.y, ss, ts); q.z = repack(q.z, ss, ts); q.w = repack(q.w, ss, ts); }
		*(uint4 *)t = q;
	} else {
		for (int i = 0; i < n; i++) { t[i] = convert ? repack(s[i], ss, ts) : s[i]; }
	}
}

// ref: api/drawAPI.cpp:834-904
__global__ void __launch_bounds__(256) higher_kernel(Img targetH, Img sourceH, Img targetA, Img sourceA, Img targetB, Img sourceB, Intersection is, float offset) {
	int32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * PX, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= is.w || y >= is.h) { return; }
	int n = min(PX, is.w - x);
	// both height fields come in as 16-byte loads when the rows allow it (large images: the pass is a stream of heights); colours are only
	// touched where the source is higher
	Img sourceBits = sourceH, targetBits = targetH;
	uint32_t sourceWords[4], targetWords[4];
	load4(sourceBits, is.sx + x, is.sy + y, n, sourceWords);
	load4(targetBits, is.tx + x, is.ty + y, n, targetWords);
	bool changed = false;
#pragma unroll
	for (int i = 0; i < 4; i++) {
		if (i >= n) { break; }
		float newHeight = __uint_as_float(sourceWords[i]);
		if (newHeight > -INFINITY) {
			newHeight += offset;
			if (newHeight > __uint_as_float(targetWords[i])) {
				targetWords[i] = __float_as_uint(newHeight);
				changed = true;
				if (targetA.data) { *px_u32(targetA, is.tx + x + i, is.ty + y) = repack(*px_u32(sourceA, is.sx + x + i, is.sy + y), pack_shifts(sourceA.packOrder), pack_shifts(targetA.packOrder)); }
				if (targetB.data) { *px_u32(targetB, is.tx + x + i, is.ty + y) = repack(*px_u32(sourceB, is.sx + x + i, is.sy + y), pack_shifts(sourceB.packOrder), pack_shifts(targetB.packOrder)); }
			}
		}
	}
	if (changed) { store4(targetBits, is.tx + x, is.ty + y, n, targetWords); }
}

struct SpriteDev { Img sourceH, sourceA, sourceB; int32_t left, top; float offset; };

// All sprites of a batch applied to one target tile in array order: each target pixel is read and written once.
// ref: SDK/SpriteEngine/spriteAPI.cpp:316-323 drawSprite → draw_higher, called in a loop (:525-539, :726-733).
__global__ void __launch_bounds__(256) higher_batch_kernel(Img targetH, Img targetA, Img targetB, const SpriteDev *__restrict__ sprites, int32_t count) {
	int32_t x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
	int32_t tileL = blockIdx.x * 32, tileT = blockIdx.y * 8;
	bool inside = x < targetH.width && y < targetH.height;
	float h = 0.0f;
	uint32_t a = 0, b = 0;
	if (inside) {
		h = *px_f32(targetH, x, y);
		if (targetA.data) { a = *px_u32(targetA, x, y); }
		if (targetB.data) { b = *px_u32(targetB, x, y); }
	}
	bool dirty = false;
	uint32_t ta = pack_shifts(targetA.packOrder), tb = pack_shifts(targetB.packOrder);
	for (int32_t s = 0; s < count; s++) {
		const SpriteDev &sp = sprites[s];
		// tile-uniform rejection
		if (sp.left >= tileL + 32 || sp.left + sp.sourceH.width <= tileL || sp.top >= tileT + 8 || sp.top + sp.sourceH.height <= tileT) { continue; }
		int32_t sx = x - sp.left, sy = y - sp.top;
		if (inside && sx >= 0 && sy >= 0 && sx < sp.sourceH.width && sy < sp.sourceH.height) {
			float newHeight = *px_f32(sp.sourceH, sx, sy);
			if (newHeight > -INFINITY) {
				newHeight += sp.offset;
				if (newHeight > h) {
					h = newHeight;
					dirty = true;
					if (targetA.data) { a = repack(*px_u32(sp.sourceA, sx, sy), pack_shifts(sp.sourceA.packOrder), ta); }
					if (targetB.data) { b = repack(*px_u32(sp.sourceB, sx, sy), pack_shifts(sp.sourceB.packOrder), tb); }
				}
			}
		}
	}
	if (dirty) {
		*px_f32(targetH, x, y) = h;
		if (targetA.data) { *px_u32(targetA, x, y) = a; }
		if (targetB.data) { *px_u32(targetB, x, y) = b; }
	}
}

// ------------------------------------------------------------------------------------------------ Sandbox light

// ref: base/simd.h:2670 saturatedAddition on four bytes
__device__ __forceinline__ uint32_t sat_add_bytes(uint32_t a, uint32_t b) { return __vaddus4(a, b); }

// ref: SDK/SpriteEngine/lightAPI.cpp:52-56 — clampUpper 255.1, truncate, pack r | g << 8 | b << 16
__device__ __forceinline__ uint32_t light_pack(float r, float g, float b) {
	return saturated_byte(r) | (saturated_byte(g) << 8) | (saturated_byte(b) << 16);
}

// ref: SDK/SpriteEngine/lightAPI.cpp:23-68
__global__ void __launch_bounds__(256) directed_kernel(Img light, Img normal, float rx, float ry, float rz, float colorR, float colorG, float colorB, int add) {
	const int32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * PX, y0 = blockIdx.y * (blockDim.y * ROWS) + threadIdx.y;
	if (x >= light.width) { return; }
	const int n = min(PX, light.width - x);
	uint32_t nc[ROWS][4], old[ROWS][4];
#pragma unroll
	for (int j = 0; j < ROWS; j++) {
		const int32_t y = y0 + j * blockDim.y;
		if (y < light.height) {
			load4(normal, x, y, n, nc[j]);
			if (add) { load4(light, x, y, n, old[j]); }
		}
	}
#pragma unroll
	for (int j = 0; j < ROWS; j++) {
		const int32_t y = y0 + j * blockDim.y;
		if (y < light.height) {
			uint32_t out[4];
#pragma unroll
			for (int i = 0; i < 4; i++) {
				float nx = (float)(nc[j][i] & 255u) - 128.0f, ny = (float)((nc[j][i] >> 8) & 255u) - 128.0f, nz = (float)((nc[j][i] >> 16) & 255u) - 128.0f;
				float dot = (nx * rx) + (ny * ry) + (nz * rz);
				float in = dot > 0.0f ? dot : 0.0f;
				uint32_t packed = light_pack(in * colorR, in * colorG, in * colorB);
				out[i] = add ? sat_add_bytes(old[j][i], packed) : packed;
			}
			store4(light, x, y, n, out);
		}
	}
}

// ref: SDK/SpriteEngine/lightAPI.cpp:287-323
__global__ void __launch_bounds__(256) blend_kernel(Img color, Img diffuse, Img light) {
	const int32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * PX, y0 = blockIdx.y * (blockDim.y * ROWS) + threadIdx.y;
	if (x >= color.width) { return; }
	const int n = min(PX, color.width - x);
	uint32_t d[ROWS][4], l[ROWS][4];
#pragma unroll
	for (int j = 0; j < ROWS; j++) {
		const int32_t y = y0 + j * blockDim.y;
		if (y < color.height) { load4(diffuse, x, y, n, d[j]); load4(light, x, y, n, l[j]); }
	}
	const float scale = 0.0078125f; // 1 / 128
	const uint32_t shifts = pack_shifts(color.packOrder);
#pragma unroll
	for (int j = 0; j < ROWS; j++) {
		const int32_t y = y0 + j * blockDim.y;
		if (y < color.height) {
			uint32_t out[4];
#pragma unroll
			for (int i = 0; i < 4; i++) {
				float red = ((float)(d[j][i] & 255u) * (float)(l[j][i] & 255u)) * scale;
				float green = ((float)((d[j][i] >> 8) & 255u) * (float)((l[j][i] >> 8) & 255u)) * scale;
				float blue = ((float)((d[j][i] >> 16) & 255u) * (float)((l[j][i] >> 16) & 255u)) * scale;
				out[i] = pack_rgba_ordered(saturated_byte(red), saturated_byte(green), saturated_byte(blue), 0u, shifts);
			}
			store4(color, x, y, n, out);
		}
	}
}

// (float)(1.0 / sqrt((double)x)) — reciprocalSquareRoot of the reference's scalar build (ref: base/simd.h:4104) — without the FP64 square
// root and division (about 70 double-precision instructions per pixel and light; the point light was bound by them). A 22-bit seed and two
// Newton steps in double land within 2^-50 of the true value; the reference's own result lies within 2^-52 of it. Unless the estimate is
// closer than 2^-47 (relative) to a boundary between two floats, both round to the same float; the few values that are (3.6e-7 of all
// inputs) take the literal expression. Checked on the host over 4e8 inputs with perturbed seeds and on the device by
// dfpsr_selftest_rsqrt (tests/test_gpu_pixel_ops.py).
__device__ __forceinline__ float reference_rsqrt(float x) {
	if (!(x >= 1.17549435e-38f && x <= 3.0e38f)) { return (float)(1.0 / sqrt((double)x)); } // zero, denormal, negative, inf, nan
	const double d = (double)x, halfD = 0.5 * d;
	double y = (double)rsqrtf(x);
	y = y * (1.5 - (halfD * (y * y)));
	y = y * (1.5 - (halfD * (y * y)));
	const float f = (float)y;
	const double ulp = (double)__uint_as_float(__float_as_uint(f) + 1u) - (double)f;
	const double below = (__float_as_uint(f) & 0x7FFFFFu) == 0u ? 0.25 * ulp : 0.5 * ulp; // the spacing halves below a power of two
	const double diff = y - (double)f, slack = y * 7.105427357601002e-15; // 2^-47
	if (diff > 0.5 * ulp - slack || -diff > below - slack) { return (float)(1.0 / sqrt((double)x)); }
	return f;
}

__global__ void __launch_bounds__(256) selftest_rsqrt_kernel(uint32_t firstBits, uint32_t count, unsigned long long *mismatches) {
	unsigned long long bad = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		const float x = __uint_as_float(firstBits + i);
		const float fast = reference_rsqrt(x), literal = (float)(1.0 / sqrt((double)x));
		if (__float_as_uint(fast) != __float_as_uint(literal) && !(fast != fast && literal != literal)) { bad++; }
	}
	if (bad) { atomicAdd(mismatches, bad); }
}

struct PointLightParams {
	int32_t left, top, width, height; // lane-aligned rectangle (ref: lightAPI.cpp:76-105)
	int32_t laneCount;
	float baseX, baseY, baseZ;       // light-space offset of the rectangle's first pixel centre at height 0
	float dxX, dxY, dxZ, dyX, dyY, dyZ, faceX, faceY, faceZ;
	float colorR, colorG, colorB, reciprocalRadius;
	int32_t shadow;
	float cubeCenter;
};

// ref: SDK/SpriteEngine/lightAPI.cpp:107-139
__device__ float shadow_transparency(const Img &cube, float halfWidth, float ox, float oy, float oz) {
	int32_t width = cube.width;
	float absX = ox < 0.0f ? -ox : ox, absY = oy < 0.0f ? -oy : oy, absZ = oz < 0.0f ? -oz : oz;
	bool xIsLongest = absX > absY && absX > absZ;
	bool yIsLongerThanZ = absY > absZ;
	float depth = xIsLongest ? ox : (yIsLongerThanZ ? oy : oz);
	float slopeUp = (yIsLongerThanZ && !xIsLongest) ? oz : oy;
	float slopeSide = xIsLongest ? -oz : (yIsLongerThanZ ? -ox : ox);
	int32_t viewOffset = width * (xIsLongest ? 0 : (yIsLongerThanZ ? 2 : 4));
	if (depth < 0.0f) { depth = -depth; slopeSide = -slopeSide; viewOffset += width; }
	float reciDepth = 1.0f / depth;
	float scale = halfWidth * reciDepth;
	int32_t sampleX = __float2int_rz(halfWidth + (slopeSide * scale));
	int32_t sampleY = __float2int_rz(halfWidth - (slopeUp * scale));
	int32_t maxPixel = width - 1;
	sampleX = min(max(sampleX, 0), maxPixel);
	sampleY = min(max(sampleY, 0), maxPixel);
	float shadowReciDepth = *px_f32(cube, sampleX, sampleY + viewOffset);
	return reciDepth * 1.02f > shadowReciDepth ? 1.0f : 0.0f;
}

// One CTA per image row of the light's rectangle. The reference walks the rectangle with running sums
// (lightBaseRowX += dy per row, lightBasePixel += laneCount * dx per vector, lightAPI.cpp:196-268); the first
// 3 * laneCount threads replay those sums for this row into shared memory, then all threads shade pixels.
__global__ void __launch_bounds__(256) point_light_kernel(Img light, Img normal, Img height, Img cube, PointLightParams p) {
	extern __shared__ float sChain[]; // [vector][component * laneCount + lane]
	const int32_t y = p.top + (int32_t)blockIdx.x;
	const int32_t lanes = p.laneCount, vectors = p.width / lanes;
	if ((int32_t)threadIdx.x < 3 * lanes) {
		int32_t comp = threadIdx.x / lanes, l = threadIdx.x % lanes;
		float base = comp == 0 ? p.baseX : (comp == 1 ? p.baseY : p.baseZ);
		float dx = comp == 0 ? p.dxX : (comp == 1 ? p.dxY : p.dxZ), dy = comp == 0 ? p.dyX : (comp == 1 ? p.dyY : p.dyZ);
		// createGradient (ref: base/simd.h:474): lane 0 = start, lane 1 = start + inc, lane l = start + inc * l
		float v = l == 0 ? base : (l == 1 ? base + dx : base + dx * (float)l);
		for (int32_t r = 0; r < (int32_t)blockIdx.x; r++) { v += dy; }
		float step = dx * (float)lanes;
		for (int32_t k = 0; k < vectors; k++) { sChain[k * 3 * lanes + comp * lanes + l] = v; v += step; }
	}
	__syncthreads();
	for (int32_t i = threadIdx.x; i < p.width; i += blockDim.x) {
		int32_t x = p.left + i;
		if (x >= light.width) { continue; } // lanes past the image width are row padding in the reference
		int32_t k = i / lanes, l = i % lanes;
		float h = *px_f32(height, x, y);
		float ox = sChain[k * 3 * lanes + l] + (p.faceX * h);
		float oy = sChain[k * 3 * lanes + lanes + l] + (p.faceY * h);
		float oz = sChain[k * 3 * lanes + 2 * lanes + l] + (p.faceZ * h);
		float sq = (ox * ox) + (oy * oy) + (oz * oz);
		float lightRatio = sqrtf(sq) * p.reciprocalRadius;
		if (lightRatio >= 1.0f) { continue; } // the distance term is (1 - 2) + 1 = 0 exactly: this light adds zero bytes here
		uint32_t nc = *px_u32(normal, x, y);
		float nx = ((float)(nc & 255u) - 128.0f) * (-1.0f / 128.0f), ny = ((float)((nc >> 8) & 255u) - 128.0f) * (-1.0f / 128.0f), nz = ((float)((nc >> 16) & 255u) - 128.0f) * (-1.0f / 128.0f);
		float distanceIntensity = 1.0f - 2.0f * lightRatio + lightRatio * lightRatio;
		// scalar-build reciprocalSquareRoot: the quotient is formed in double (ref: base/simd.h:4104, see oracle/dfpsr_oracle.c)
		float rs = reference_rsqrt(sq);
		float dot = ((ox * rs) * nx) + ((oy * rs) * ny) + ((oz * rs) * nz);
		float in = (dot > 0.0f ? dot : 0.0f) * distanceIntensity;
		if (in == 0.0f) { continue; } // zero times the shadow term (0 or 1) stays zero
		if (p.shadow) { in = in * shadow_transparency(cube, p.cubeCenter, ox, oy, oz); }
		uint32_t *t = px_u32(light, x, y);
		*t = sat_add_bytes(*t, light_pack(in * p.colorR, in * p.colorG, in * p.colorB));
	}
}

// ---- the whole light part of a Sandbox frame in one pass (ref: SDK/SpriteEngine/spriteAPI.cpp:775-814): directed lights (the first
// one sets, the others add), every point light, blend. Per pixel: normal, height and diffuse are read once, light and colour written
// once (20 B instead of 8 + 16 per point light + 12). Saturating byte additions of non-negative contributions commute, so keeping the
// light accumulator in a register gives the reference's bytes.
static const int FRAME_MAX_DIRECTED = 8;
static const int FRAME_LIGHT_GROUP = 16; // at most this many point lights have their row sums staged in shared memory together (12 threads each)
struct FramePointLight { PointLightParams p; Img cube; };
struct FrameLightParams {
	int32_t directedCount, pointCount, chainStride, group;
	float directed[FRAME_MAX_DIRECTED][6]; // rx, ry, rz, colour r, g, b
	const FramePointLight *points;
};

// One CTA per image row. For every group of point lights, 12 threads per light (3 components x 4 lanes) first replay the reference's
// running sums for this row (lightAPI.cpp:196-268: += dy per row, += laneCount * dx per vector), then all threads shade pixels.
__global__ void __launch_bounds__(256) light_frame_kernel(Img color, Img diffuse, Img light, Img normal, Img height, FrameLightParams fp) {
	extern __shared__ float sChains[]; // [light in group][vector][component * 4 + lane], chainStride floats per light
	const int32_t y = (int32_t)blockIdx.x;
	const int32_t perThread = (light.width + (int32_t)blockDim.x - 1) / (int32_t)blockDim.x;
	// pixels x = threadIdx.x + i * blockDim.x, at most 8 per thread are kept in registers per sweep
	for (int32_t sweep = 0; sweep < perThread; sweep += 8) {
		uint32_t acc[8], nc[8];
		float h[8];
#pragma unroll
		for (int i = 0; i < 8; i++) {
			const int32_t x = (int32_t)threadIdx.x + (sweep + i) * (int32_t)blockDim.x;
			acc[i] = 0u; nc[i] = 0u; h[i] = 0.0f;
			if (sweep + i < perThread && x < light.width) {
				nc[i] = *px_u32(normal, x, y);
				if (height.data) { h[i] = *px_f32(height, x, y); }
				const float nx = (float)(nc[i] & 255u) - 128.0f, ny = (float)((nc[i] >> 8) & 255u) - 128.0f, nz = (float)((nc[i] >> 16) & 255u) - 128.0f;
				for (int32_t d = 0; d < fp.directedCount; d++) {
					const float dot = (nx * fp.directed[d][0]) + (ny * fp.directed[d][1]) + (nz * fp.directed[d][2]);
					const float in = dot > 0.0f ? dot : 0.0f;
					const uint32_t packed = light_pack(in * fp.directed[d][3], in * fp.directed[d][4], in * fp.directed[d][5]);
					acc[i] = d == 0 ? packed : sat_add_bytes(acc[i], packed);
				}
			}
		}
		for (int32_t groupStart = 0; groupStart < fp.pointCount; groupStart += fp.group) {
			const int32_t groupCount = min(fp.group, fp.pointCount - groupStart);
			__syncthreads();
			if ((int32_t)threadIdx.x < groupCount * 12) {
				const int32_t g = threadIdx.x / 12, comp = (threadIdx.x % 12) / 4, l = threadIdx.x % 4;
				const PointLightParams &p = fp.points[groupStart + g].p;
				if (y >= p.top && y < p.top + p.height) {
					const float base = comp == 0 ? p.baseX : (comp == 1 ? p.baseY : p.baseZ);
					const float dx = comp == 0 ? p.dxX : (comp == 1 ? p.dxY : p.dxZ), dy = comp == 0 ? p.dyX : (comp == 1 ? p.dyY : p.dyZ);
					float v = l == 0 ? base : (l == 1 ? base + dx : base + dx * (float)l); // createGradient (ref: base/simd.h:474)
					for (int32_t r = p.top; r < y; r++) { v += dy; }
					const float step = dx * 4.0f;
					float *dst = sChains + g * fp.chainStride + comp * 4 + l;
					const int32_t vectors = p.width / 4;
					for (int32_t k = 0; k < vectors; k++) { dst[k * 12] = v; v += step; }
				}
			}
			__syncthreads();
			for (int32_t g = 0; g < groupCount; g++) {
				const FramePointLight &fl = fp.points[groupStart + g];
				const PointLightParams &p = fl.p;
				if (y < p.top || y >= p.top + p.height) { continue; }
				const float *chain = sChains + g * fp.chainStride;
#pragma unroll
				for (int i = 0; i < 8; i++) {
					const int32_t x = (int32_t)threadIdx.x + (sweep + i) * (int32_t)blockDim.x;
					const int32_t rel = x - p.left;
					if (sweep + i < perThread && x < light.width && rel >= 0 && rel < p.width) {
						const int32_t k = rel >> 2, l = rel & 3;
						const float ox = chain[k * 12 + l] + (p.faceX * h[i]);
						const float oy = chain[k * 12 + 4 + l] + (p.faceY * h[i]);
						const float oz = chain[k * 12 + 8 + l] + (p.faceZ * h[i]);
						const float sq = (ox * ox) + (oy * oy) + (oz * oz);
						float lightRatio = sqrtf(sq) * p.reciprocalRadius;
						// At or beyond the radius the ratio clamps to 1 and the distance term is (1 - 2) + 1 = 0 exactly: the pixel receives
						// max(dot, 0) * 0 = 0 from this light (the corners of the light's square, a fifth of it), and adding zero bytes changes nothing.
						if (lightRatio >= 1.0f) { continue; }
						const float nx = ((float)(nc[i] & 255u) - 128.0f) * (-1.0f / 128.0f), ny = ((float)((nc[i] >> 8) & 255u) - 128.0f) * (-1.0f / 128.0f), nz = ((float)((nc[i] >> 16) & 255u) - 128.0f) * (-1.0f / 128.0f);
						const float distanceIntensity = 1.0f - 2.0f * lightRatio + lightRatio * lightRatio;
						const float rs = reference_rsqrt(sq); // scalar-build reciprocalSquareRoot (ref: base/simd.h:4104)
						const float dot = ((ox * rs) * nx) + ((oy * rs) * ny) + ((oz * rs) * nz);
						float in = (dot > 0.0f ? dot : 0.0f) * distanceIntensity;
						if (in == 0.0f) { continue; } // faces away from the light: zero times the shadow term (0 or 1) stays zero
						if (p.shadow) { in = in * shadow_transparency(fl.cube, p.cubeCenter, ox, oy, oz); }
						acc[i] = sat_add_bytes(acc[i], light_pack(in * p.colorR, in * p.colorG, in * p.colorB));
					}
				}
			}
		}
		const uint32_t shifts = pack_shifts(color.packOrder);
#pragma unroll
		for (int i = 0; i < 8; i++) {
			const int32_t x = (int32_t)threadIdx.x + (sweep + i) * (int32_t)blockDim.x;
			if (sweep + i < perThread && x < light.width) {
				*px_u32(light, x, y) = acc[i];
				if (color.data) {
					const uint32_t d = *px_u32(diffuse, x, y);
					const float scale = 0.0078125f;
					const float red = ((float)(d & 255u) * (float)(acc[i] & 255u)) * scale;
					const float green = ((float)((d >> 8) & 255u) * (float)((acc[i] >> 8) & 255u)) * scale;
					const float blue = ((float)((d >> 16) & 255u) * (float)((acc[i] >> 16) & 255u)) * scale;
					*px_u32(color, x, y) = pack_rgba_ordered(saturated_byte(red), saturated_byte(green), saturated_byte(blue), 0u, shifts);
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ textures and filters

// ref: api/textureAPI.cpp:44-63 downsample, one level
__global__ void __launch_bounds__(256) downsample_kernel(const uint32_t *__restrict__ source, uint32_t *__restrict__ target, uint32_t targetWidth, uint32_t targetHeight) {
	uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= targetWidth || y >= targetHeight) { return; }
	uint32_t sw = targetWidth * 2;
	uint2 top = *(const uint2 *)(source + (size_t)(2 * y) * sw + 2 * x), bottom = *(const uint2 *)(source + (size_t)(2 * y + 1) * sw + 2 * x);
	uint32_t out = 0;
#pragma unroll
	for (int s = 0; s < 32; s += 8) {
		out |= ((((top.x >> s) & 255u) + ((top.y >> s) & 255u) + ((bottom.x >> s) & 255u) + ((bottom.y >> s) & 255u)) / 4u) << s;
	}
	target[(size_t)y * targetWidth + x] = out;
}

struct Rgba { int32_t r, g, b, a; };

// ref: api/imageAPI.h:284-290 image_readPixel_clamp → channels in RGBA order
__device__ __forceinline__ Rgba read_clamp(const Img &im, int32_t x, int32_t y) {
	x = min(max(x, 0), im.width - 1); y = min(max(y, 0), im.height - 1);
	uint32_t c = *px_u32(im, x, y), s = pack_shifts(im.packOrder);
	Rgba o = {(int32_t)((c >> (s & 31u)) & 255u), (int32_t)((c >> ((s >> 8) & 31u)) & 255u), (int32_t)((c >> ((s >> 16) & 31u)) & 255u), (int32_t)((c >> ((s >> 24) & 31u)) & 255u)};
	return o;
}
__device__ __forceinline__ uint32_t saturate_and_pack(Rgba c, uint32_t shifts) { // ref: PackOrder.h:108-110
	return pack_rgba_ordered((uint32_t)min(max(c.r, 0), 255), (uint32_t)min(max(c.g, 0), 255), (uint32_t)min(max(c.b, 0), 255), (uint32_t)min(max(c.a, 0), 255), shifts);
}
__device__ __forceinline__ Rgba lerp16(Rgba a, Rgba b, uint32_t ratioB) { // ref: api/filterAPI.cpp:86-88
	uint32_t ra = 65536u - ratioB;
	Rgba o = {(int32_t)(((uint32_t)a.r * ra + (uint32_t)b.r * ratioB) >> 16), (int32_t)(((uint32_t)a.g * ra + (uint32_t)b.g * ratioB) >> 16),
	          (int32_t)(((uint32_t)a.b * ra + (uint32_t)b.b * ratioB) >> 16), (int32_t)(((uint32_t)a.a * ra + (uint32_t)b.a * ratioB) >> 16)};
	return o;
}
// ref: api/filterAPI.cpp:49-63 mixColorsUniform (8-bit weights on packed colours)
// Bytes 1 and 3 reach their 16-bit lanes, and the four result bytes are gathered, with one byte permute each instead of shift + mask
// pairs: the up-scaling kernel was bound by the integer ALU pipe (76 % busy), the multiplies run on the FMA pipe.
__device__ __forceinline__ uint32_t mix_uniform(uint32_t a, uint32_t b, uint32_t fineRatio) {
	uint32_t ratio = fineRatio >> 8, inv = 256u - ratio;
	uint32_t low = (a & 0x00FF00FFu) * inv + (b & 0x00FF00FFu) * ratio;
	uint32_t high = __byte_perm(a, 0u, 0x4341) * inv + __byte_perm(b, 0u, 0x4341) * ratio;
	return __byte_perm(low, high, 0x7351); // ((low >> 8) & 0x00FF00FF) | (high & 0xFF00FF00)
}

// lerp16 + saturate-and-pack on whole pixels when source and target share a pack order: every byte lane is interpolated on its own
// (ref: api/filterAPI.cpp:86-88, the weights add up to 65536, so a lane's sum stays below 2^24 and its result is byte 2 of the sum;
// interpolated bytes never leave 0..255, so the saturation is the identity). 19 instructions per pixel instead of about 50.
// Each lane sum a_k * ratioA + b_k * ratioB is one two-way dot product (dp2a: 16-bit weights against the byte pair (a_k, b_k)), so a pixel
// costs two permutes to pair the bytes, four dot products and three permutes to gather byte 2 of the sums. ratioB == 0 would need the
// weight 65536, which does not fit 16 bits: that case returns a itself ((a_k * 65536) >> 16).
__device__ __forceinline__ uint32_t lerp16_lanes(uint32_t a, uint32_t b, uint32_t ratioB) {
	const uint32_t weights = ((65536u - ratioB) & 0xFFFFu) | (ratioB << 16);
	const uint32_t pair01 = __byte_perm(a, b, 0x5140), pair23 = __byte_perm(a, b, 0x7362); // (a0, b0, a1, b1), (a2, b2, a3, b3)
	const uint32_t s0 = __dp2a_lo(weights, pair01, 0u), s1 = __dp2a_hi(weights, pair01, 0u);
	const uint32_t s2 = __dp2a_lo(weights, pair23, 0u), s3 = __dp2a_hi(weights, pair23, 0u);
	const uint32_t mixed = __byte_perm(__byte_perm(s0, s1, 0x4462), __byte_perm(s2, s3, 0x4462), 0x5410);
	return ratioB == 0u ? a : mixed;
}

enum { RESIZE_VERTICAL_PACKED = 0, RESIZE_VERTICAL = 1, RESIZE_VERTICAL_NEAREST = 2, RESIZE_HORIZONTAL = 3, RESIZE_GENERAL = 4 };

struct ResizeParams { int32_t offsetX, offsetY, startX, startY, bilinear, path; };

// ref: api/filterAPI.cpp:118-259 — one thread per 4 target pixels of a row; the 16.16 read position is start + i * offset.
__global__ void __launch_bounds__(256) resize_kernel(Img target, Img source, ResizeParams rp) {
	int32_t x0 = (blockIdx.x * blockDim.x + threadIdx.x) * PX, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x0 >= target.width || y >= target.height) { return; }
	int n = min(PX, target.width - x0);
	uint32_t shifts = pack_shifts(target.packOrder);
	int32_t readY = rp.startY + y * rp.offsetY;
	uint32_t sampleY = (uint32_t)(readY < 0 ? 0 : readY);
	uint32_t upperY = sampleY >> 16, lowerRatio = sampleY & 65535u;
	uint32_t out[4];
	if (rp.path <= RESIZE_VERTICAL_NEAREST) {
		uint32_t lowerY = upperY + 1;
		if (upperY >= (uint32_t)source.height) { upperY = (uint32_t)source.height - 1; }
		if (lowerY >= (uint32_t)source.height) { lowerY = (uint32_t)source.height - 1; }
		if (rp.path == RESIZE_VERTICAL_PACKED) {
			uint32_t up[4], lo[4];
			load4(source, x0, (int32_t)upperY, n, up);
			load4(source, x0, (int32_t)lowerY, n, lo);
			for (int i = 0; i < 4; i++) { out[i] = mix_uniform(up[i], lo[i], lowerRatio); }
		} else if (rp.path == RESIZE_VERTICAL) {
			for (int i = 0; i < n; i++) { out[i] = saturate_and_pack(lerp16(read_clamp(source, x0 + i, (int32_t)upperY), read_clamp(source, x0 + i, (int32_t)lowerY), lowerRatio), shifts); }
		} else {
			load4(source, x0, (int32_t)upperY, n, out);
		}
	} else if (rp.bilinear && target.packOrder == source.packOrder) {
		// the same integers on whole pixels, byte lane by byte lane (lerp16_lanes): four clamped reads and three interpolations per pixel
		const int32_t lastX = source.width - 1, lastY = source.height - 1;
		const int32_t rowA = rp.path == RESIZE_HORIZONTAL ? min(y, lastY) : min((int32_t)upperY, lastY), rowB = min((int32_t)upperY + 1, lastY);
		const uint32_t *lineA = row_ptr<uint32_t>(source.data, source.stride, rowA), *lineB = row_ptr<uint32_t>(source.data, source.stride, rowB);
		for (int i = 0; i < n; i++) {
			const int32_t readX = rp.startX + (x0 + i) * rp.offsetX;
			const uint32_t sampleX = (uint32_t)(readX < 0 ? 0 : readX);
			const int32_t leftX = min((int32_t)(sampleX >> 16), lastX), rightX = min((int32_t)(sampleX >> 16) + 1, lastX);
			const uint32_t rightRatio = sampleX & 65535u;
			const uint32_t upper = lerp16_lanes(__ldg(lineA + leftX), __ldg(lineA + rightX), rightRatio);
			if (rp.path == RESIZE_HORIZONTAL) { out[i] = upper; }
			else { out[i] = lerp16_lanes(upper, lerp16_lanes(__ldg(lineB + leftX), __ldg(lineB + rightX), rightRatio), lowerRatio); }
		}
	} else {
		for (int i = 0; i < n; i++) {
			int32_t readX = rp.startX + (x0 + i) * rp.offsetX;
			uint32_t sampleX = (uint32_t)(readX < 0 ? 0 : readX);
			int32_t leftX = (int32_t)(sampleX >> 16);
			uint32_t rightRatio = sampleX & 65535u;
			Rgba c;
			if (rp.path == RESIZE_HORIZONTAL) {
				c = rp.bilinear ? lerp16(read_clamp(source, leftX, y), read_clamp(source, leftX + 1, y), rightRatio) : read_clamp(source, leftX, y);
			} else if (rp.bilinear) {
				Rgba upper = lerp16(read_clamp(source, leftX, (int32_t)upperY), read_clamp(source, leftX + 1, (int32_t)upperY), rightRatio);
				Rgba lower = lerp16(read_clamp(source, leftX, (int32_t)upperY + 1), read_clamp(source, leftX + 1, (int32_t)upperY + 1), rightRatio);
				c = lerp16(upper, lower, lowerRatio);
			} else {
				c = read_clamp(source, leftX, (int32_t)upperY);
			}
			out[i] = saturate_and_pack(c, shifts);
		}
	}
	store4(target, x0, y, n, out);
}

// filter_resize, bilinear, when the reference runs two passes (api/filterAPI.cpp:298-314: the width changes and the height grows): pass 1
// stretches every source row horizontally with 16-bit weights into a temporary image in the target's pack order (:118-154, lerp16 +
// saturate and pack), pass 2 mixes two rows of the temporary with 8-bit weights (mixColorsUniform, :49-63, :200-230). This kernel produces
// the same integers in ONE pass: a thread owns 4 columns x FUSED_ROWS target rows, forms each needed row of the temporary in registers
// (a row is reused by about targetHeight / sourceHeight consecutive target rows) and never writes it to memory:
// 4 B x (source + target) of traffic instead of 4 B x (source + 3 x temporary + target).
#ifndef RESIZE_UP_ROWS
#define RESIZE_UP_ROWS 16
#endif
static const int FUSED_ROWS = RESIZE_UP_ROWS;

#ifndef RESIZE_UP_MIN_BLOCKS
#define RESIZE_UP_MIN_BLOCKS 6 // 42 registers, 48 warps per SM: 115 us for 4096^2 -> 8192^2 against 121 us unbounded and 128 us at 8 (tools/resize_sweep.py)
#endif
// The common thread of the fused up-scale: four whole columns, source and target in the same pack order. Everything that depends on the
// column only is prepared once — the two clamped source columns and the 16-bit weight pair of each target column — so that a row of the
// temporary costs eight loads and four lerps and nothing else. A column whose right weight is zero reads its left pixel twice with the
// weights (65535, 1): (a * 65535 + a) >> 16 == a, the value lerp16_lanes returns for that case, without a select in the loop.
__device__ __forceinline__ void resize_up_whole_columns(const Img &target, const Img &source, const ResizeParams &rp, int32_t x0, int32_t yFirst) {
	const int32_t last = source.width - 1;
	int32_t left[4], right[4];
	uint32_t weights[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const int32_t readX = rp.startX + (x0 + i) * rp.offsetX;
		const uint32_t sampleX = (uint32_t)(readX < 0 ? 0 : readX), ratio = sampleX & 65535u;
		left[i] = min((int32_t)(sampleX >> 16), last);
		right[i] = ratio == 0u ? left[i] : min((int32_t)(sampleX >> 16) + 1, last);
		weights[i] = ratio == 0u ? (65535u | (1u << 16)) : ((65536u - ratio) | (ratio << 16));
	}
	auto stretch_row = [&](int32_t row, uint32_t *value) {
		const uint32_t *line = row_ptr<uint32_t>(source.data, source.stride, row);
		uint32_t a[4], b[4];
#pragma unroll
		for (int i = 0; i < 4; i++) { a[i] = __ldg(line + left[i]); b[i] = __ldg(line + right[i]); }
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const uint32_t pair01 = __byte_perm(a[i], b[i], 0x5140), pair23 = __byte_perm(a[i], b[i], 0x7362);
			const uint32_t s0 = __dp2a_lo(weights[i], pair01, 0u), s1 = __dp2a_hi(weights[i], pair01, 0u);
			const uint32_t s2 = __dp2a_lo(weights[i], pair23, 0u), s3 = __dp2a_hi(weights[i], pair23, 0u);
			value[i] = __byte_perm(__byte_perm(s0, s1, 0x4462), __byte_perm(s2, s3, 0x4462), 0x5410);
		}
	};
	const int32_t lastRow = source.height - 1;
	int32_t upperRow = -1, lowerRow = -1;
	uint32_t upperValue[4] = {0u, 0u, 0u, 0u}, lowerValue[4] = {0u, 0u, 0u, 0u};
	const int32_t yEnd = min(yFirst + FUSED_ROWS, target.height);
	for (int32_t y = yFirst; y < yEnd; y++) {
		const int32_t readY = rp.startY + y * rp.offsetY;
		const uint32_t sampleY = (uint32_t)(readY < 0 ? 0 : readY);
		const int32_t upperY = min((int32_t)(sampleY >> 16), lastRow), lowerY = min((int32_t)(sampleY >> 16) + 1, lastRow);
		if (upperY != upperRow) {
			if (upperY == lowerRow) {
#pragma unroll
				for (int i = 0; i < 4; i++) { upperValue[i] = lowerValue[i]; }
			} else {
				stretch_row(upperY, upperValue);
			}
			upperRow = upperY;
		}
		if (lowerY != lowerRow) {
			if (lowerY == upperY) {
#pragma unroll
				for (int i = 0; i < 4; i++) { lowerValue[i] = upperValue[i]; }
			} else {
				stretch_row(lowerY, lowerValue);
			}
			lowerRow = lowerY;
		}
		uint32_t out[4];
#pragma unroll
		for (int i = 0; i < 4; i++) { out[i] = mix_uniform(upperValue[i], lowerValue[i], sampleY & 65535u); }
		store4(target, x0, y, 4, out);
	}
}

template <bool SAME_ORDER>
__global__ void __launch_bounds__(256, RESIZE_UP_MIN_BLOCKS) resize_up_fused_kernel(Img target, Img source, ResizeParams rp) {
	const int32_t x0 = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x) * PX, yFirst = (int32_t)(blockIdx.y * blockDim.y + threadIdx.y) * FUSED_ROWS;
	if (x0 >= target.width || yFirst >= target.height) { return; }
	if (SAME_ORDER && target.width - x0 >= PX) { resize_up_whole_columns(target, source, rp, x0, yFirst); return; }
	const int n = min(PX, target.width - x0);
	const uint32_t shifts = pack_shifts(target.packOrder);
	int32_t leftX[4];
	uint32_t rightRatio[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const int32_t readX = rp.startX + (x0 + i) * rp.offsetX;
		const uint32_t sampleX = (uint32_t)(readX < 0 ? 0 : readX);
		leftX[i] = (int32_t)(sampleX >> 16); rightRatio[i] = sampleX & 65535u;
	}
	// The two rows of the temporary a target row needs, (upper, lower), live in registers; target rows walk downwards and, when up-scaling,
	// advance by at most one source row, so the old lower row usually becomes the new upper row. No indexed register arrays: every move is static.
	int32_t upperRow = -1, lowerRow = -1;
	uint32_t upperValue[4] = {0u, 0u, 0u, 0u}, lowerValue[4] = {0u, 0u, 0u, 0u};
	const int32_t last = source.width - 1;
	auto stretch_row = [&](int32_t row, uint32_t *value) {
		if (SAME_ORDER) {
			const uint32_t *line = row_ptr<uint32_t>(source.data, source.stride, row); // row is already clamped
#pragma unroll
			for (int i = 0; i < 4; i++) {
				if (i < n) { value[i] = lerp16_lanes(__ldg(line + min(leftX[i], last)), __ldg(line + min(leftX[i] + 1, last)), rightRatio[i]); }
			}
		} else {
#pragma unroll
			for (int i = 0; i < 4; i++) {
				if (i < n) { value[i] = saturate_and_pack(lerp16(read_clamp(source, leftX[i], row), read_clamp(source, leftX[i] + 1, row), rightRatio[i]), shifts); }
			}
		}
	};
	const int32_t yEnd = min(yFirst + FUSED_ROWS, target.height);
	for (int32_t y = yFirst; y < yEnd; y++) {
		const int32_t readY = rp.startY + y * rp.offsetY;
		const uint32_t sampleY = (uint32_t)(readY < 0 ? 0 : readY);
		uint32_t upperY = sampleY >> 16, lowerY = upperY + 1;
		const uint32_t lowerRatio = sampleY & 65535u;
		if (upperY >= (uint32_t)source.height) { upperY = (uint32_t)source.height - 1; }
		if (lowerY >= (uint32_t)source.height) { lowerY = (uint32_t)source.height - 1; }
		if ((int32_t)upperY != upperRow) {
			if ((int32_t)upperY == lowerRow) {
#pragma unroll
				for (int i = 0; i < 4; i++) { upperValue[i] = lowerValue[i]; }
			} else {
				stretch_row((int32_t)upperY, upperValue);
			}
			upperRow = (int32_t)upperY;
		}
		if ((int32_t)lowerY != lowerRow) {
			if (lowerY == upperY) {
#pragma unroll
				for (int i = 0; i < 4; i++) { lowerValue[i] = upperValue[i]; }
			} else {
				stretch_row((int32_t)lowerY, lowerValue);
			}
			lowerRow = (int32_t)lowerY;
		}
		uint32_t out[4];
#pragma unroll
		for (int i = 0; i < 4; i++) { out[i] = mix_uniform(upperValue[i], lowerValue[i], lowerRatio); }
		store4(target, x0, y, n, out);
	}
}

// ref: api/filterAPI.cpp:95-154 resize_reference<*, ImageU8, uint8_t> — one byte per pixel, both axes sampled for every target pixel
// (16-bit weights, clamped reads). One thread per four target pixels of a row, stored as one word when the address allows.
__global__ void __launch_bounds__(256) resize_u8_kernel(Img target, Img source, ResizeParams rp) {
	const int32_t x0 = (blockIdx.x * blockDim.x + threadIdx.x) * PX, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x0 >= target.width || y >= target.height) { return; }
	const int n = min(PX, target.width - x0);
	const int32_t lastX = source.width - 1, lastY = source.height - 1;
	const int32_t readY = rp.startY + y * rp.offsetY;
	const uint32_t sampleY = (uint32_t)(readY < 0 ? 0 : readY), lowerRatio = sampleY & 65535u;
	const uint8_t *upperLine = row_ptr<uint8_t>(source.data, source.stride, min((int32_t)(sampleY >> 16), lastY));
	const uint8_t *lowerLine = row_ptr<uint8_t>(source.data, source.stride, min((int32_t)(sampleY >> 16) + 1, lastY));
	uint32_t out[4] = {0u, 0u, 0u, 0u};
	for (int i = 0; i < n; i++) {
		const int32_t readX = rp.startX + (x0 + i) * rp.offsetX;
		const uint32_t sampleX = (uint32_t)(readX < 0 ? 0 : readX), rightRatio = sampleX & 65535u;
		const int32_t leftX = min((int32_t)(sampleX >> 16), lastX), rightX = min((int32_t)(sampleX >> 16) + 1, lastX);
		if (rp.bilinear) {
			const uint32_t upper = ((uint32_t)__ldg(upperLine + leftX) * (65536u - rightRatio) + (uint32_t)__ldg(upperLine + rightX) * rightRatio) >> 16;
			const uint32_t lower = ((uint32_t)__ldg(lowerLine + leftX) * (65536u - rightRatio) + (uint32_t)__ldg(lowerLine + rightX) * rightRatio) >> 16;
			out[i] = (upper * (65536u - lowerRatio) + lower * lowerRatio) >> 16;
		} else {
			out[i] = __ldg(upperLine + leftX);
		}
	}
	uint8_t *p = row_ptr<uint8_t>(target.data, target.stride, y) + x0;
	if (n == 4 && (((uintptr_t)p) & 3u) == 0) { *(uint32_t *)p = out[0] | (out[1] << 8) | (out[2] << 16) | (out[3] << 24); }
	else { for (int i = 0; i < n; i++) { p[i] = (uint8_t)out[i]; } }
}

struct MapParams { int32_t op, startX, startY; int32_t p[8]; };

// ref: api/filterAPI.cpp:759-777 with the enumerated device ops of dfpsr_b200.h
__global__ void __launch_bounds__(256) map_kernel(Img target, Img source, MapParams mp) {
	int32_t x0 = (blockIdx.x * blockDim.x + threadIdx.x) * PX, ty = blockIdx.y * blockDim.y + threadIdx.y;
	if (x0 >= target.width || ty >= target.height) { return; }
	int n = min(PX, target.width - x0);
	uint32_t shifts = pack_shifts(target.packOrder);
	uint32_t out[4];
	int32_t y = ty + mp.startY;
	uint32_t src[4];
	bool direct = false;
	if (mp.op == DFPSR_MAP_AFFINE) {
		// fast path: the four reads are inside the source and contiguous
		int32_t sx = x0 + mp.startX;
		direct = n == 4 && sx >= 0 && sx + 3 < source.width && y >= 0 && y < source.height;
		if (direct) { load4(source, sx, y, 4, src); }
	}
	uint32_t ss = pack_shifts(source.packOrder);
	for (int i = 0; i < n; i++) {
		int32_t x = x0 + i + mp.startX;
		Rgba c = {0, 0, 0, 0};
		if (mp.op == DFPSR_MAP_XOR_PATTERN) { c.r = x & 255; c.g = y & 255; c.b = (x ^ y) & 255; c.a = 255; }
		else if (mp.op == DFPSR_MAP_AFFINE) {
			Rgba s;
			if (direct) { s.r = (int32_t)((src[i] >> (ss & 31u)) & 255u); s.g = (int32_t)((src[i] >> ((ss >> 8) & 31u)) & 255u); s.b = (int32_t)((src[i] >> ((ss >> 16) & 31u)) & 255u); s.a = (int32_t)((src[i] >> ((ss >> 24) & 31u)) & 255u); }
			else { s = read_clamp(source, x, y); }
			c.r = s.r * mp.p[0] + mp.p[4]; c.g = s.g * mp.p[1] + mp.p[5]; c.b = s.b * mp.p[2] + mp.p[6]; c.a = s.a * mp.p[3] + mp.p[7];
		} else { c.r = mp.p[0]; c.g = mp.p[1]; c.b = mp.p[2]; c.a = mp.p[3]; }
		out[i] = saturate_and_pack(c, shifts);
	}
	store4(target, x0, ty, n, out);
}


__device__ __forceinline__ uint32_t affine_pixel(uint32_t c, uint32_t ss, uint32_t ts, const int32_t *p) {
	int32_t r = (int32_t)((c >> (ss & 31u)) & 255u) * p[0] + p[4], g = (int32_t)((c >> ((ss >> 8) & 31u)) & 255u) * p[1] + p[5];
	int32_t b = (int32_t)((c >> ((ss >> 16) & 31u)) & 255u) * p[2] + p[6], a = (int32_t)((c >> ((ss >> 24) & 31u)) & 255u) * p[3] + p[7];
	return pack_rgba_ordered((uint32_t)min(max(r, 0), 255), (uint32_t)min(max(g, 0), 255), (uint32_t)min(max(b, 0), 255), (uint32_t)min(max(a, 0), 255), ts);
}

// The same pixel when source and target are both in RGBA order: bytes are taken and put back with byte permutes and each clamp is one
// relu-min (15 instructions per pixel instead of 27; the kernel issued at 64 % with DRAM at 52 % before).
__device__ __forceinline__ uint32_t affine_pixel_rgba(uint32_t c, const int32_t *p) {
	const int32_t r = __vimin_s32_relu((int32_t)__byte_perm(c, 0, 0x4440) * p[0] + p[4], 255), g = __vimin_s32_relu((int32_t)__byte_perm(c, 0, 0x4441) * p[1] + p[5], 255);
	const int32_t b = __vimin_s32_relu((int32_t)__byte_perm(c, 0, 0x4442) * p[2] + p[6], 255), a = __vimin_s32_relu((int32_t)__byte_perm(c, 0, 0x4443) * p[3] + p[7], 255);
	return __byte_perm(__byte_perm((uint32_t)r, (uint32_t)g, 0x1140), __byte_perm((uint32_t)b, (uint32_t)a, 0x1140), 0x5410);
}

// filter_map, affine op, every read inside the source (ref: api/filterAPI.cpp:759-777 with image_readPixel_clamp never clamping).
template <bool RGBA_ORDER>
__global__ void __launch_bounds__(256) map_affine_stream_kernel(Img target, Img source, MapParams mp) {
	const int32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * PX, y0 = blockIdx.y * (blockDim.y * ROWS) + threadIdx.y;
	if (x >= target.width) { return; }
	const int n = min(PX, target.width - x);
	const uint32_t ss = pack_shifts(source.packOrder), ts = pack_shifts(target.packOrder);
	uint32_t v[ROWS][4];
#pragma unroll
	for (int j = 0; j < ROWS; j++) {
		const int32_t y = y0 + j * blockDim.y;
		if (y < target.height) { load4(source, x + mp.startX, y + mp.startY, n, v[j]); }
	}
#pragma unroll
	for (int j = 0; j < ROWS; j++) {
		const int32_t y = y0 + j * blockDim.y;
		if (y < target.height) {
#pragma unroll
			for (int i = 0; i < 4; i++) { v[j][i] = RGBA_ORDER ? affine_pixel_rgba(v[j][i], mp.p) : affine_pixel(v[j][i], ss, ts, mp.p); }
			store4(target, x, y, n, v[j]);
		}
	}
}

// filter_resize, bilinear, exactly half the width and half the height: the 16.16 read position of target pixel i is 2 i + 0.5 on both
// axes (ref: api/filterAPI.cpp:118-154 with offset = 131072 and start = 65536 - 32768), so every lerp has weight 32768 and
// (a * 32768 + b * 32768) >> 16 == (a + b) >> 1 per channel: two truncating byte averages horizontally, one vertically.
__global__ void __launch_bounds__(256) resize_half_kernel(Img target, Img source) {
	const int32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * PX, y0 = blockIdx.y * (blockDim.y * 2) + threadIdx.y;
	if (x >= target.width) { return; }
	const int n = min(PX, target.width - x);
	const uint32_t ss = pack_shifts(source.packOrder), ts = pack_shifts(target.packOrder);
	uint32_t up[2][8], lo[2][8];
#pragma unroll
	for (int j = 0; j < 2; j++) {
		const int32_t y = y0 + j * blockDim.y;
		if (y < target.height) {
			load4(source, 2 * x, 2 * y, min(4, 2 * n), up[j]); load4(source, 2 * x + 4, 2 * y, max(0, 2 * n - 4), up[j] + 4);
			load4(source, 2 * x, 2 * y + 1, min(4, 2 * n), lo[j]); load4(source, 2 * x + 4, 2 * y + 1, max(0, 2 * n - 4), lo[j] + 4);
		}
	}
#pragma unroll
	for (int j = 0; j < 2; j++) {
		const int32_t y = y0 + j * blockDim.y;
		if (y < target.height) {
			uint32_t out[4];
#pragma unroll
			for (int i = 0; i < 4; i++) {
				uint32_t c = __vhaddu4(__vhaddu4(up[j][2 * i], up[j][2 * i + 1]), __vhaddu4(lo[j][2 * i], lo[j][2 * i + 1]));
				out[i] = ss == ts ? c : repack(c, ss, ts);
			}
			store4(target, x, y, n, out);
		}
	}
}

// ref: api/filterAPI.cpp:724-757 + :340-365
__global__ void __launch_bounds__(256) magnify_kernel(Img target, Img source, int32_t pixelWidth, int32_t pixelHeight, int32_t clipWidth, int32_t clipHeight) {
	int32_t x0 = (blockIdx.x * blockDim.x + threadIdx.x) * PX, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x0 >= target.width || y >= target.height) { return; }
	int n = min(PX, target.width - x0);
	uint32_t ss = pack_shifts(source.packOrder), ts = pack_shifts(target.packOrder);
	uint32_t out[4];
	for (int i = 0; i < n; i++) {
		int32_t x = x0 + i;
		out[i] = 0;
		if (x < clipWidth && y < clipHeight) {
			out[i] = repack(*px_u32(source, min(x / pixelWidth, source.width - 1), min(y / pixelHeight, source.height - 1)), ss, ts);
		}
	}
	store4(target, x0, y, n, out);
}

static bool exists(const dfpsr_image *im) { return im != nullptr && im->data != nullptr; }

struct V3 { float x, y, z; };
static V3 mat_transform(const dfpsr_matrix3x3 &m, V3 p) { // ref: math/FMatrix3x3.h:52-58
	return V3{p.x * m.xAxis[0] + p.y * m.yAxis[0] + p.z * m.zAxis[0], p.x * m.xAxis[1] + p.y * m.yAxis[1] + p.z * m.zAxis[1], p.x * m.xAxis[2] + p.y * m.yAxis[2] + p.z * m.zAxis[2]};
}
static V3 mat_transform_transposed(const dfpsr_matrix3x3 &m, V3 p) { // ref: math/FMatrix3x3.h:63-69
	return V3{p.x * m.xAxis[0] + p.y * m.xAxis[1] + p.z * m.xAxis[2], p.x * m.yAxis[0] + p.y * m.yAxis[1] + p.z * m.yAxis[2], p.x * m.zAxis[0] + p.y * m.zAxis[1] + p.z * m.zAxis[2]};
}
static V3 normalize3(V3 v) { // ref: math/FVector.h:113-120
	float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
	if (l == 0.0f) { return V3{0.0f, 0.0f, 1.0f}; }
	return V3{v.x / l, v.y / l, v.z / l};
}

static int resize_single(const Img &target, const Img &source, bool bilinear, bool simdAligned, cudaStream_t stream);

} // namespace dfpsr

using namespace dfpsr;

extern "C" {

int dfpsr_image_fill_rgba(const dfpsr_image *image, int32_t red, int32_t green, int32_t blue, int32_t alpha, void *stream) {
	if (!exists(image)) { return 0; } // ref: api/drawAPI.cpp:162
	auto clamp255 = [](int32_t v) { return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
	uint32_t shifts = pack_shifts(image->packOrder);
	uint32_t packed = (clamp255(red) << (shifts & 31u)) | (clamp255(green) << ((shifts >> 8) & 31u)) | (clamp255(blue) << ((shifts >> 16) & 31u)) | (clamp255(alpha) << ((shifts >> 24) & 31u));
	Img t = img_of(image);
	DFPSR_LAUNCH(fill_kernel, grid_for(t.width, t.height, BLOCK), BLOCK, 0, as_stream(stream), t, packed);
	return 0;
}

int dfpsr_image_fill_f32(const dfpsr_image *image, float value, void *stream) {
	if (!exists(image)) { return 0; }
	uint32_t bits;
	memcpy(&bits, &value, 4);
	Img t = img_of(image);
	DFPSR_LAUNCH(fill_kernel, grid_for(t.width, t.height, BLOCK), BLOCK, 0, as_stream(stream), t, bits);
	return 0;
}

int dfpsr_draw_copy_rgba(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, void *stream) {
	if (!exists(target) || !exists(source)) { return 0; } // ref: api/drawAPI.cpp:906-911
	Img t = img_of(target), s = img_of(source);
	Intersection is;
	if (!intersect(t, s, left, top, is)) { return 0; }
	DFPSR_LAUNCH(copy_kernel, grid_for(is.w, is.h, BLOCK), BLOCK, 0, as_stream(stream), t, s, is, t.packOrder != s.packOrder ? 1 : 0);
	return 0;
}

int dfpsr_draw_copy_f32(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, void *stream) {
	if (!exists(target) || !exists(source)) { return 0; }
	Img t = img_of(target), s = img_of(source);
	Intersection is;
	if (!intersect(t, s, left, top, is)) { return 0; }
	DFPSR_LAUNCH(copy_kernel, grid_for(is.w, is.h, BLOCK), BLOCK, 0, as_stream(stream), t, s, is, 0);
	return 0;
}

int dfpsr_draw_higher(const dfpsr_image *targetHeight, const dfpsr_image *sourceHeight, const dfpsr_image *targetA, const dfpsr_image *sourceA, const dfpsr_image *targetB, const dfpsr_image *sourceB, int32_t left, int32_t top, float sourceHeightOffset, void *stream) {
	// ref: api/drawAPI.cpp:962-979 — every image given to an overload must exist, otherwise nothing is drawn
	if (!exists(targetHeight) || !exists(sourceHeight)) { return 0; }
	bool wantA = targetA != nullptr || sourceA != nullptr, wantB = targetB != nullptr || sourceB != nullptr;
	if (wantA && (!exists(targetA) || !exists(sourceA))) { return 0; }
	if (wantB && (!exists(targetB) || !exists(sourceB))) { return 0; }
	DFPSR_REQUIRE(!wantB || wantA, "draw_higher: a second payload image needs a first one");
	Img th = img_of(targetHeight), sh = img_of(sourceHeight);
	if (wantA) { DFPSR_REQUIRE(sourceA->width == sh.width && sourceA->height == sh.height, "draw_higher: sourceA and sourceHeight differ in size"); }
	if (wantB) { DFPSR_REQUIRE(sourceB->width == sh.width && sourceB->height == sh.height, "draw_higher: sourceB and sourceHeight differ in size"); }
	Intersection is;
	if (!intersect(th, sh, left, top, is)) { return 0; }
	DFPSR_LAUNCH(higher_kernel, grid_for(is.w, is.h, BLOCK), BLOCK, 0, as_stream(stream), th, sh, img_of(wantA ? targetA : nullptr), img_of(wantA ? sourceA : nullptr), img_of(wantB ? targetB : nullptr), img_of(wantB ? sourceB : nullptr), is, sourceHeightOffset);
	return 0;
}

int dfpsr_draw_higher_batch(const dfpsr_image *targetHeight, const dfpsr_image *targetA, const dfpsr_image *targetB, const dfpsr_sprite_draw *draws, int32_t count, void *stream) {
	if (!exists(targetHeight) || count <= 0) { return 0; }
	DFPSR_REQUIRE(draws != nullptr, "draw_higher_batch: null draws");
	// The sprite records travel through a small ring of page-locked staging buffers, each with its own device copy and an event behind
	// its upload: a slot is reused four calls later, when its copy has long completed, so the call never waits for the stream
	// (round 1 synchronised the stream here, once per regenerated background block and once per frame of a Sandbox scene).
	struct Slot { SpriteDev *host = nullptr; size_t capacity = 0; DeviceBuffer device; cudaEvent_t uploaded = nullptr; bool busy = false; };
	static thread_local Slot ring[4];
	static thread_local unsigned next = 0;
	Slot &slot = ring[next++ & 3u];
	if (slot.busy) { DFPSR_CHECK_CUDA(cudaEventSynchronize(slot.uploaded)); slot.busy = false; }
	if (slot.uploaded == nullptr) { DFPSR_CHECK_CUDA(cudaEventCreateWithFlags(&slot.uploaded, cudaEventDisableTiming)); }
	if ((size_t)count > slot.capacity) {
		if (slot.host) { cudaFreeHost(slot.host); slot.host = nullptr; }
		slot.capacity = (size_t)count * 2;
		DFPSR_CHECK_CUDA(cudaMallocHost((void **)&slot.host, slot.capacity * sizeof(SpriteDev)));
	}
	if (slot.device.reserve((size_t)count * sizeof(SpriteDev))) { return 1; }
	for (int32_t i = 0; i < count; i++) {
		slot.host[i].sourceH = img_of(&draws[i].sourceHeight);
		slot.host[i].sourceA = img_of(&draws[i].sourceA);
		slot.host[i].sourceB = img_of(&draws[i].sourceB);
		slot.host[i].left = draws[i].left; slot.host[i].top = draws[i].top; slot.host[i].offset = draws[i].heightOffset;
	}
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(slot.device.ptr, slot.host, (size_t)count * sizeof(SpriteDev), cudaMemcpyHostToDevice, as_stream(stream)));
	DFPSR_CHECK_CUDA(cudaEventRecord(slot.uploaded, as_stream(stream)));
	slot.busy = true;
	Img th = img_of(targetHeight);
	dim3 grid((unsigned)((th.width + 31) / 32), (unsigned)((th.height + 7) / 8));
	DFPSR_LAUNCH(higher_batch_kernel, grid, 256, 0, as_stream(stream), th, img_of(exists(targetA) ? targetA : nullptr), img_of(exists(targetB) ? targetB : nullptr), (const SpriteDev *)slot.device.ptr, count);
	return 0;
}

int dfpsr_light_directed(const dfpsr_ortho_view *view, const dfpsr_image *light, const dfpsr_image *normal, const float direction[3], float intensity, const int32_t colorRgb[3], int32_t add, void *stream) {
	DFPSR_REQUIRE(view && exists(light) && exists(normal) && direction && colorRgb, "light_directed: null argument");
	DFPSR_REQUIRE(light->width == normal->width && light->height == normal->height, "light_directed: light and normal buffers differ in size");
	// ref: SDK/SpriteEngine/lightAPI.cpp:26-30 (host-side uniforms)
	V3 n = normalize3(mat_transform_transposed(view->normalToWorldSpace, V3{direction[0], direction[1], direction[2]}));
	float rx = -n.x * intensity * 2.0f, ry = -n.y * intensity * 2.0f, rz = -n.z * intensity * 2.0f;
	float colorR = fmaxf(0.0f, (float)colorRgb[0] / 255.0f), colorG = fmaxf(0.0f, (float)colorRgb[1] / 255.0f), colorB = fmaxf(0.0f, (float)colorRgb[2] / 255.0f);
	Img l = img_of(light);
	DFPSR_LAUNCH(directed_kernel, grid_rows(l.width, l.height, BLOCK), BLOCK, 0, as_stream(stream), l, img_of(normal), rx, ry, rz, colorR, colorG, colorB, add);
	return 0;
}

// ref: SDK/SpriteEngine/lightAPI.cpp:76-105 calculateBound + :190-206 uniforms of addPointLightSuper. Returns false when the light's
// rectangle misses the image.
static bool point_light_params(PointLightParams &p, const dfpsr_ortho_view *view, const int32_t worldCenter[2], int32_t imageWidth, int32_t imageHeight, const float position[3], float radius, float intensity, const int32_t colorRgb[3], const dfpsr_image *shadowCubeMap) {
	const int32_t laneCount = 4; // laneCountX_32Bit of the reference's SSE2 and scalar builds
	V3 S = mat_transform_transposed(view->normalToWorldSpace, V3{position[0], position[1], position[2]});
	V3 rotated = mat_transform(view->lightSpaceToScreenDepth, S);
	int32_t cx = (int32_t)rotated.x + worldCenter[0], cy = (int32_t)rotated.y + worldCenter[1];
	int32_t pixelRadius = (int32_t)(radius * view->lightSpaceToScreenDepth.xAxis[0]);
	if (cx < -pixelRadius || cx > imageWidth + pixelRadius || cy < -pixelRadius || cy > imageHeight + pixelRadius) { return false; }
	int32_t size = (int32_t)((float)pixelRadius * 2.0f);
	int32_t l = cx - pixelRadius, t = cy - pixelRadius, r = l + size, b = t + size;
	if (!(l < imageWidth && r > 0 && t < imageHeight && b > 0)) { return false; }
	l = l > 0 ? l : 0; t = t > 0 ? t : 0; r = r < imageWidth ? r : imageWidth; b = b < imageHeight ? b : imageHeight;
	if (r <= l || b <= t) { return false; }
	l = (l / laneCount) * laneCount;
	r = ((r + laneCount - 1) / laneCount) * laneCount;
	p.left = l; p.top = t; p.width = r - l; p.height = b - t; p.laneCount = laneCount;
	V3 origin = mat_transform(view->screenDepthToLightSpace, V3{0.5f - (float)worldCenter[0] + (float)l, 0.5f - (float)worldCenter[1] + (float)t, 0.0f});
	p.baseX = origin.x - S.x; p.baseY = origin.y - S.y; p.baseZ = origin.z - S.z;
	p.dxX = view->screenDepthToLightSpace.xAxis[0]; p.dxY = view->screenDepthToLightSpace.xAxis[1]; p.dxZ = view->screenDepthToLightSpace.xAxis[2];
	p.dyX = view->screenDepthToLightSpace.yAxis[0]; p.dyY = view->screenDepthToLightSpace.yAxis[1]; p.dyZ = view->screenDepthToLightSpace.yAxis[2];
	p.faceX = view->screenDepthToLightSpace.zAxis[0]; p.faceY = view->screenDepthToLightSpace.zAxis[1]; p.faceZ = view->screenDepthToLightSpace.zAxis[2];
	p.colorR = fmaxf(0.0f, (float)colorRgb[0] * intensity); p.colorG = fmaxf(0.0f, (float)colorRgb[1] * intensity); p.colorB = fmaxf(0.0f, (float)colorRgb[2] * intensity);
	p.reciprocalRadius = 1.0f / radius;
	p.shadow = exists(shadowCubeMap) ? 1 : 0;
	p.cubeCenter = p.shadow ? (float)shadowCubeMap->width * 0.5f : 0.0f;
	return true;
}

int dfpsr_light_point(const dfpsr_ortho_view *view, const int32_t worldCenter[2], const dfpsr_image *light, const dfpsr_image *normal, const dfpsr_image *height, const float position[3], float radius, float intensity, const int32_t colorRgb[3], const dfpsr_image *shadowCubeMap, void *stream) {
	DFPSR_REQUIRE(view && worldCenter && exists(light) && exists(normal) && exists(height) && position && colorRgb, "light_point: null argument");
	PointLightParams p;
	if (!point_light_params(p, view, worldCenter, light->width, light->height, position, radius, intensity, colorRgb, shadowCubeMap)) { return 0; }
	size_t smem = (size_t)(p.width / p.laneCount) * 3 * p.laneCount * sizeof(float);
	if (smem > 48 * 1024) { DFPSR_CHECK_CUDA(cudaFuncSetAttribute(point_light_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }
	DFPSR_LAUNCH(point_light_kernel, p.height, 256, smem, as_stream(stream), img_of(light), img_of(normal), img_of(height), img_of(p.shadow ? shadowCubeMap : nullptr), p);
	return 0;
}

// ref: SDK/SpriteEngine/spriteAPI.cpp:775-814 — the light part of SpriteWorldImpl::draw in ONE kernel.
int dfpsr_light_frame(const dfpsr_ortho_view *view, const int32_t worldCenter[2], const dfpsr_image *color, const dfpsr_image *diffuse, const dfpsr_image *light, const dfpsr_image *normal, const dfpsr_image *height, const dfpsr_directed_light *directed, int32_t directedCount, const dfpsr_point_light *points, int32_t pointCount, void *stream) {
	DFPSR_REQUIRE(view && worldCenter && exists(light) && exists(normal), "light_frame: null argument");
	DFPSR_REQUIRE(directedCount >= 0 && pointCount >= 0 && (directedCount == 0 || directed) && (pointCount == 0 || points), "light_frame: bad light arrays");
	DFPSR_REQUIRE(pointCount == 0 || exists(height), "light_frame: point lights need the height buffer");
	DFPSR_REQUIRE(light->width == normal->width && light->height == normal->height, "light_frame: light and normal buffers differ in size");
	DFPSR_REQUIRE(!exists(color) || (exists(diffuse) && color->width == light->width && color->height == light->height && diffuse->width == light->width && diffuse->height == light->height), "light_frame: colour, diffuse and light buffers differ in size");
	DFPSR_REQUIRE(directedCount <= FRAME_MAX_DIRECTED, "light_frame: at most %d directed lights", FRAME_MAX_DIRECTED);
	FrameLightParams fp;
	memset(&fp, 0, sizeof(fp));
	fp.directedCount = directedCount;
	for (int32_t i = 0; i < directedCount; i++) {
		// ref: SDK/SpriteEngine/lightAPI.cpp:26-30 (host-side uniforms)
		V3 n = normalize3(mat_transform_transposed(view->normalToWorldSpace, V3{directed[i].direction[0], directed[i].direction[1], directed[i].direction[2]}));
		fp.directed[i][0] = -n.x * directed[i].intensity * 2.0f; fp.directed[i][1] = -n.y * directed[i].intensity * 2.0f; fp.directed[i][2] = -n.z * directed[i].intensity * 2.0f;
		fp.directed[i][3] = fmaxf(0.0f, (float)directed[i].colorRgb[0] / 255.0f); fp.directed[i][4] = fmaxf(0.0f, (float)directed[i].colorRgb[1] / 255.0f); fp.directed[i][5] = fmaxf(0.0f, (float)directed[i].colorRgb[2] / 255.0f);
	}
	static thread_local DeviceBuffer staging;
	std::vector<FramePointLight> lights;
	int32_t maxWidth = 0;
	for (int32_t i = 0; i < pointCount; i++) {
		FramePointLight fl;
		if (!point_light_params(fl.p, view, worldCenter, light->width, light->height, points[i].position, points[i].radius, points[i].intensity, points[i].colorRgb, &points[i].shadowCubeMap)) { continue; }
		fl.cube = img_of(fl.p.shadow ? &points[i].shadowCubeMap : nullptr);
		if (fl.p.width > maxWidth) { maxWidth = fl.p.width; }
		lights.push_back(fl);
	}
	fp.pointCount = (int32_t)lights.size();
	if (!lights.empty()) {
		if (staging.reserve(lights.size() * sizeof(FramePointLight))) { return 1; }
		DFPSR_CHECK_CUDA(cudaMemcpyAsync(staging.ptr, lights.data(), lights.size() * sizeof(FramePointLight), cudaMemcpyHostToDevice, as_stream(stream)));
		fp.points = (const FramePointLight *)staging.ptr;
	}
	fp.chainStride = maxWidth * 3;
	// lights per shared-memory group: about 40 KB per CTA keeps five CTAs (40 warps) on an SM; the arithmetic (double-precision rsqrt of
	// the reference's scalar build, cube-map gathers) needs the occupancy more than it needs fewer barriers
	size_t perLight = (size_t)fp.chainStride * sizeof(float);
	fp.group = perLight > 0 ? (int32_t)((40 * 1024) / perLight) : FRAME_LIGHT_GROUP;
	if (fp.group < 1) { fp.group = 1; }
	if (fp.group > FRAME_LIGHT_GROUP) { fp.group = FRAME_LIGHT_GROUP; }
	size_t smem = (size_t)fp.group * perLight;
	DFPSR_REQUIRE(smem <= 200 * 1024, "light_frame: light rectangles of %d pixels exceed the shared memory budget", maxWidth);
	if (smem > 48 * 1024) { DFPSR_CHECK_CUDA(cudaFuncSetAttribute(light_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }
	Img l = img_of(light);
	DFPSR_LAUNCH(light_frame_kernel, l.height, 256, smem, as_stream(stream), img_of(exists(color) ? color : nullptr), img_of(exists(color) ? diffuse : nullptr), l, img_of(normal), img_of(exists(height) ? height : nullptr), fp);
	return 0;
}

int dfpsr_light_blend(const dfpsr_image *color, const dfpsr_image *diffuse, const dfpsr_image *light, void *stream) {
	DFPSR_REQUIRE(exists(color) && exists(diffuse) && exists(light), "light_blend: null argument");
	DFPSR_REQUIRE(color->width == diffuse->width && color->height == diffuse->height && color->width == light->width && color->height == light->height, "light_blend: buffers differ in size");
	Img c = img_of(color);
	DFPSR_LAUNCH(blend_kernel, grid_rows(c.width, c.height, BLOCK), BLOCK, 0, as_stream(stream), c, img_of(diffuse), img_of(light));
	return 0;
}

int dfpsr_texture_generate_pyramid(const dfpsr_texture *texture, void *stream) {
	DFPSR_REQUIRE(texture != nullptr && texture->data != nullptr, "texture_generate_pyramid: texture does not exist");
	uint32_t *px = (uint32_t *)texture->data;
	for (uint32_t level = 1; level <= texture->maxMipLevel; level++) {
		uint32_t tw = 1u << (texture->log2width - level), th = 1u << (texture->log2height - level);
		const uint32_t *src = px + (texture->startOffset & (texture->maxLevelMask >> (2 * (level - 1))));
		uint32_t *dst = px + (texture->startOffset & (texture->maxLevelMask >> (2 * level)));
		dim3 block(32, 8), grid((tw + 31) / 32, (th + 7) / 8);
		DFPSR_LAUNCH(downsample_kernel, grid, block, 0, as_stream(stream), src, dst, tw, th);
	}
	return 0;
}

size_t dfpsr_filter_resize_scratch_bytes(int32_t sourceWidth, int32_t sourceHeight, int32_t newWidth, int32_t newHeight) {
	if (newWidth != sourceWidth && newHeight > sourceHeight) { return (size_t)newWidth * (size_t)sourceHeight * 4; }
	return 0;
}

static int resize_u8_single(const Img &target, const Img &source, bool bilinear, cudaStream_t stream) {
	ResizeParams rp;
	rp.offsetX = (int32_t)(65536u * (uint32_t)source.width / (uint32_t)target.width);
	rp.offsetY = (int32_t)(65536u * (uint32_t)source.height / (uint32_t)target.height);
	rp.startX = rp.offsetX / 2 - (bilinear ? 32768 : 0); rp.startY = rp.offsetY / 2 - (bilinear ? 32768 : 0);
	rp.bilinear = bilinear ? 1 : 0; rp.path = RESIZE_GENERAL;
	DFPSR_LAUNCH(resize_u8_kernel, grid_for(target.width, target.height, BLOCK), BLOCK, 0, stream, target, source, rp);
	return 0;
}

int dfpsr_filter_resize_u8(const dfpsr_image *target, const dfpsr_image *source, int32_t sampler, void *scratch, void *stream) {
	DFPSR_REQUIRE(exists(target) && exists(source), "filter_resize_u8: null argument");
	Img t = img_of(target), s = img_of(source);
	const bool bilinear = sampler == DFPSR_SAMPLER_LINEAR;
	// ref: api/filterAPI.cpp:298-314 resizeToTarget: two passes (and their two roundings) when the width changes and the height grows
	if (t.width != s.width && t.height > s.height) {
		DFPSR_REQUIRE(scratch != nullptr, "filter_resize_u8: up-scaling both dimensions needs a scratch buffer of target.width * source.height bytes");
		Img temp;
		temp.data = (uint8_t *)scratch; temp.width = t.width; temp.height = s.height; temp.stride = t.width; temp.packOrder = 0;
		if (resize_u8_single(temp, s, bilinear, as_stream(stream))) { return 1; }
		return resize_u8_single(t, temp, bilinear, as_stream(stream));
	}
	return resize_u8_single(t, s, bilinear, as_stream(stream));
}

int dfpsr_filter_resize(const dfpsr_image *target, const dfpsr_image *source, int32_t sampler, int32_t sourceIsSubImage, void *scratch, void *stream) {
	DFPSR_REQUIRE(exists(target) && exists(source), "filter_resize: null argument");
	Img t = img_of(target), s = img_of(source);
	bool bilinear = sampler == DFPSR_SAMPLER_LINEAR;
	// ref: api/filterAPI.cpp:298-314 resizeToTarget
	if (t.width != s.width && t.height > s.height && bilinear) {
		// both passes of the reference in one kernel (the scratch buffer is not needed)
		ResizeParams rp;
		rp.offsetX = (int32_t)(65536u * (uint32_t)s.width / (uint32_t)t.width);
		rp.offsetY = (int32_t)(65536u * (uint32_t)s.height / (uint32_t)t.height);
		rp.startX = rp.offsetX / 2 - 32768; rp.startY = rp.offsetY / 2 - 32768;
		rp.bilinear = 1; rp.path = RESIZE_GENERAL;
		dim3 grid((unsigned)((t.width + PX * (int)BLOCK.x - 1) / (PX * (int)BLOCK.x)), (unsigned)((t.height + FUSED_ROWS * (int)BLOCK.y - 1) / (FUSED_ROWS * (int)BLOCK.y)));
		if (t.packOrder == s.packOrder) { DFPSR_LAUNCH(resize_up_fused_kernel<true>, grid, BLOCK, 0, as_stream(stream), t, s, rp); }
		else { DFPSR_LAUNCH(resize_up_fused_kernel<false>, grid, BLOCK, 0, as_stream(stream), t, s, rp); }
		return 0;
	}
	if (t.width != s.width && t.height > s.height) {
		DFPSR_REQUIRE(scratch != nullptr, "filter_resize: up-scaling both dimensions needs the scratch buffer (dfpsr_filter_resize_scratch_bytes)");
		Img temp;
		temp.data = (uint8_t *)scratch; temp.width = t.width; temp.height = s.height; temp.stride = t.width * 4; temp.packOrder = t.packOrder;
		if (resize_single(temp, s, bilinear, !sourceIsSubImage, as_stream(stream))) { return 1; }
		return resize_single(t, temp, bilinear, true, as_stream(stream));
	}
	return resize_single(t, s, bilinear, !sourceIsSubImage, as_stream(stream));
}

int dfpsr_texture_from_image(const dfpsr_texture *texture, const dfpsr_image *image, void *stream) {
	DFPSR_REQUIRE(texture != nullptr && texture->data != nullptr && exists(image), "texture_from_image: null argument");
	// ref: api/textureAPI.cpp:89-110: resize into level 0, then the pyramid. The temporary of a two-pass up-scale
	// is placed in the (not yet generated) lower levels when it fits, else the caller must pre-size with the layout.
	dfpsr_image level0;
	level0.data = (void *)(texture->data + texture->startOffset);
	level0.width = 1 << texture->log2width; level0.height = 1 << texture->log2height;
	level0.stride = level0.width * 4; level0.packOrder = DFPSR_PACK_RGBA;
	size_t need = dfpsr_filter_resize_scratch_bytes(image->width, image->height, level0.width, level0.height);
	static thread_local DeviceBuffer scratch;
	if (need > 0 && scratch.reserve(need)) { return 1; }
	if (dfpsr_filter_resize(&level0, image, DFPSR_SAMPLER_LINEAR, 0, scratch.ptr, stream)) { return 1; }
	return dfpsr_texture_generate_pyramid(texture, stream);
}

int dfpsr_filter_map(const dfpsr_image *target, int32_t op, const int32_t *params, int32_t paramCount, const dfpsr_image *source, int32_t startX, int32_t startY, void *stream) {
	if (!exists(target)) { return 0; } // ref: api/filterAPI.cpp:773
	MapParams mp;
	memset(&mp, 0, sizeof(mp));
	mp.op = op; mp.startX = startX; mp.startY = startY;
	int needed = op == DFPSR_MAP_AFFINE ? 8 : (op == DFPSR_MAP_CONSTANT ? 4 : 0);
	DFPSR_REQUIRE(op >= DFPSR_MAP_XOR_PATTERN && op <= DFPSR_MAP_CONSTANT, "filter_map: unknown op %d", op);
	DFPSR_REQUIRE(paramCount >= needed && (needed == 0 || params != nullptr), "filter_map: op %d needs %d parameters", op, needed);
	for (int i = 0; i < needed; i++) { mp.p[i] = params[i]; }
	DFPSR_REQUIRE(op != DFPSR_MAP_AFFINE || exists(source), "filter_map: the affine op needs a source image");
	Img t = img_of(target);
	if (op == DFPSR_MAP_AFFINE && startX >= 0 && startY >= 0 && (int64_t)startX + t.width <= source->width && (int64_t)startY + t.height <= source->height) {
		if (t.packOrder == DFPSR_PACK_RGBA && source->packOrder == DFPSR_PACK_RGBA) { DFPSR_LAUNCH(map_affine_stream_kernel<true>, grid_rows(t.width, t.height, BLOCK), BLOCK, 0, as_stream(stream), t, img_of(source), mp); }
		else { DFPSR_LAUNCH(map_affine_stream_kernel<false>, grid_rows(t.width, t.height, BLOCK), BLOCK, 0, as_stream(stream), t, img_of(source), mp); }
		return 0;
	}
	DFPSR_LAUNCH(map_kernel, grid_for(t.width, t.height, BLOCK), BLOCK, 0, as_stream(stream), t, img_of(exists(source) ? source : nullptr), mp);
	return 0;
}

int dfpsr_filter_block_magnify(const dfpsr_image *target, const dfpsr_image *source, int32_t pixelWidth, int32_t pixelHeight, void *stream) {
	if (!exists(target) || !exists(source)) { return 0; } // ref: api/filterAPI.cpp:872-876
	if (pixelWidth < 1) { pixelWidth = 1; }
	if (pixelHeight < 1) { pixelHeight = 1; }
	Img t = img_of(target), s = img_of(source);
	int32_t clipWidth = t.width < s.width * pixelWidth ? t.width : s.width * pixelWidth; clipWidth -= clipWidth % pixelWidth;
	int32_t clipHeight = t.height < s.height * pixelHeight ? t.height : s.height * pixelHeight; clipHeight -= clipHeight % pixelHeight;
	DFPSR_LAUNCH(magnify_kernel, grid_for(t.width, t.height, BLOCK), BLOCK, 0, as_stream(stream), t, s, pixelWidth, pixelHeight, clipWidth, clipHeight);
	return 0;
}

} // extern "C"

namespace dfpsr {

// ref: api/filterAPI.cpp:156-259 resize_optimized path selection (scaleRegion = whole target)
static int resize_single(const Img &target, const Img &source, bool bilinear, bool simdAligned, cudaStream_t stream) {
	bool sameWidth = source.width == target.width, sameHeight = source.height == target.height, samePack = target.packOrder == source.packOrder;
	if (sameWidth && sameHeight) {
		Intersection is{0, 0, 0, 0, target.width, target.height};
		DFPSR_LAUNCH(copy_kernel, grid_for(is.w, is.h, BLOCK), BLOCK, 0, stream, target, source, is, samePack ? 0 : 1);
		return 0;
	}
	ResizeParams rp;
	rp.offsetX = (int32_t)(65536u * (uint32_t)source.width / (uint32_t)target.width);
	rp.offsetY = (int32_t)(65536u * (uint32_t)source.height / (uint32_t)target.height);
	rp.startX = rp.offsetX / 2; rp.startY = rp.offsetY / 2;
	if (bilinear) { rp.startX -= 32768; rp.startY -= 32768; }
	rp.bilinear = bilinear ? 1 : 0;
	if (sameWidth && (samePack || bilinear)) { rp.path = bilinear ? (simdAligned ? RESIZE_VERTICAL_PACKED : RESIZE_VERTICAL) : RESIZE_VERTICAL_NEAREST; }
	else if (sameHeight) { rp.path = RESIZE_HORIZONTAL; }
	else { rp.path = RESIZE_GENERAL; }
	if (rp.path == RESIZE_GENERAL && bilinear && source.width == 2 * target.width && source.height == 2 * target.height) {
		dim3 grid((unsigned)((target.width + PX * (int)BLOCK.x - 1) / (PX * (int)BLOCK.x)), (unsigned)((target.height + 2 * (int)BLOCK.y - 1) / (2 * (int)BLOCK.y)));
		DFPSR_LAUNCH(resize_half_kernel, grid, BLOCK, 0, stream, target, source);
		return 0;
	}
	DFPSR_LAUNCH(resize_kernel, grid_for(target.width, target.height, BLOCK), BLOCK, 0, stream, target, source, rp);
	return 0;
}

} // namespace dfpsr

extern "C" int dfpsr_selftest_rsqrt(uint32_t firstBits, uint32_t count, uint64_t *mismatchesHost, void *stream) {
	DFPSR_REQUIRE(mismatchesHost != nullptr, "selftest_rsqrt: null output");
	unsigned long long *counter = nullptr;
	DFPSR_CHECK_CUDA(cudaMalloc((void **)&counter, sizeof(unsigned long long)));
	cudaError_t err = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), dfpsr::as_stream(stream));
	if (err == cudaSuccess && count > 0) {
		dfpsr::selftest_rsqrt_kernel<<<dfpsr::sm_count() * 8, 256, 0, dfpsr::as_stream(stream)>>>(firstBits, count, counter);
		dfpsr::g_launches.fetch_add(1, std::memory_order_relaxed);
		err = cudaGetLastError();
	}
	unsigned long long result = 0;
	if (err == cudaSuccess) { err = cudaMemcpyAsync(&result, counter, sizeof(result), cudaMemcpyDeviceToHost, dfpsr::as_stream(stream)); }
	if (err == cudaSuccess) { err = cudaStreamSynchronize(dfpsr::as_stream(stream)); }
	cudaFree(counter);
	if (err != cudaSuccess) { dfpsr::set_error("selftest_rsqrt: %s", cudaGetErrorString(err)); return 1; }
	*mismatchesHost = (uint64_t)result;
	return 0;
}
