// sprite_world.cu — the Sandbox sprite engine behind the C ABI (SURVEY.md §8 rows a21, a22 and the callers of a19-a26):
//   * OrthoSystem / OrthoView construction            ref: SDK/SpriteEngine/orthoAPI.cpp:5-119 (host arithmetic)
//   * DenseModel build + renderDenseModel on sm_100a   ref: SDK/SpriteEngine/spriteAPI.cpp:1176-1327
//   * sprite / model types, octrees, background blocks, dirty rectangles, temporary objects and lights: SpriteWorldImpl
//                                                      ref: SDK/SpriteEngine/spriteAPI.cpp:190-300, :351-421, :452-816, Octree.h, DirtyRectangles.h
// The reference interleaves control flow and pixel loops on one CPU thread pool. Here a frame is PLANNED on the host into a short
// operation list (the reference's decisions, in the reference's order) and then EXECUTED on the device in a handful of launches:
// one batched draw_higher per block / per frame, one copy kernel for every (block, dirty rectangle) pair of the frame, one
// dfpsr_model_render_depth_batch for all shadow cube maps and one dfpsr_light_frame for every light plus blendLight.
#include "sprite_math.cuh"

#include <algorithm>
#include <ctime>
#include <cstdlib>
#include <memory>
#include <new>
#include <vector>

namespace dfpsr {
namespace sw {

static const int32_t MINI_UNITS_PER_TILE = 1024;                       // ref: orthoAPI.h:28
static const float TILES_PER_MINI_UNIT = 1.0f / (float)MINI_UNITS_PER_TILE; // ref: orthoAPI.h:29
static const float BOTTOM_CLIP_PLANE = -1000000.0f;                   // ref: spriteAPI.cpp:18
static const int32_t BLOCK_SIZE = 512, BLOCK_MAX_DISTANCE = BLOCK_SIZE * 2; // ref: spriteAPI.cpp:513-514

static inline int32_t correct_direction(int32_t direction) { return (int32_t)((uint32_t)(direction + 8 * 1024) % 8u); } // ref: orthoAPI.h:19-21
static inline float mini_to_floating_tile(int32_t mini) { return (float)mini * TILES_PER_MINI_UNIT; }                   // ref: orthoAPI.cpp:129-131
static inline int32_t floating_tile_to_mini(float tile) { return (int32_t)round((double)tile * (double)MINI_UNITS_PER_TILE); } // ref: orthoAPI.cpp:141-143
static inline I3 i3(int32_t x, int32_t y, int32_t z) { I3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

// ref: orthoAPI.cpp:34-38 OrthoView::miniTileOffsetToScreenPixel
static I2 mini_offset_to_pixel(const dfpsr_ortho_camera &v, I3 offset) {
	I2 p;
	p.x = wrap_add(wrap_mul(v.pixelOffsetPerTileX[0], offset.x), wrap_mul(v.pixelOffsetPerTileZ[0], offset.z));
	p.y = wrap_add(wrap_mul(v.pixelOffsetPerTileX[1], offset.x), wrap_mul(v.pixelOffsetPerTileZ[1], offset.z));
	p.y = wrap_sub(p.y, wrap_mul(offset.y, v.yPixelsPerTile));
	p.x /= MINI_UNITS_PER_TILE; p.y /= MINI_UNITS_PER_TILE;
	return p;
}

// ------------------------------------------------------------------------------------------------ ortho system

// ref: implementation/render/Camera.h:157-190 for an orthogonal camera at the origin: worldToScreen(p).is
static F2 ortho_world_to_image(const M3 &cameraSystem, float imageSize, float halfWidth, F3 world) {
	const float halfHeight = halfWidth * imageSize / imageSize; // ref: Camera.h:153
	const float invWidthSlope = 0.5f / halfWidth, invHeightSlope = 0.5f / halfHeight;
	const F3 cs = transform_transposed(cameraSystem, sub(world, f3(0.0f, 0.0f, 0.0f))); // Transform3D.h:51-53
	F2 r;
	r.x = (cs.x * invWidthSlope + 0.5f) * imageSize;
	r.y = (-cs.y * invHeightSlope + 0.5f) * imageSize;
	return r;
}

// ref: orthoAPI.cpp:5-32 OrthoView::OrthoView
static void make_view(dfpsr_ortho_camera &out, int32_t id, I2 roundedX, I2 roundedZ, int32_t yPixelsPerTile, const M3 &normalToWorld, int32_t worldDirection) {
	memset(&out, 0, sizeof(out));
	out.id = id; out.worldDirection = worldDirection;
	store(out.normalToWorldSpace, normalToWorld);
	out.pixelOffsetPerTileX[0] = roundedX.x; out.pixelOffsetPerTileX[1] = roundedX.y;
	out.pixelOffsetPerTileZ[0] = roundedZ.x; out.pixelOffsetPerTileZ[1] = roundedZ.y;
	out.yPixelsPerTile = yPixelsPerTile;
	const M3 tileToScreen = m3(f3((float)roundedX.x, (float)roundedX.y, 0.0f), f3(0.0f, (float)(-yPixelsPerTile), 1.0f), f3((float)roundedZ.x, (float)roundedZ.y, 0.0f));
	const M3 screenToTile = inverse(tileToScreen);
	{ // inverse(FMatrix2x2(xAxis, zAxis)), ref: math/FMatrix2x2.h:69-76
		const float ax = (float)roundedX.x, ay = (float)roundedX.y, bx = (float)roundedZ.x, by = (float)roundedZ.y;
		const float s = 1.0f / (ax * by - ay * bx);
		out.roundedScreenPixelsToWorldTiles[0] = by * s; out.roundedScreenPixelsToWorldTiles[1] = -ay * s;
		out.roundedScreenPixelsToWorldTiles[2] = -bx * s; out.roundedScreenPixelsToWorldTiles[3] = ax * s;
	}
	store(out.screenDepthToWorldSpace, screenToTile);
	store(out.worldSpaceToScreenDepth, tileToScreen);
	const M3 toLight = m3(transform_transposed(normalToWorld, screenToTile.x), transform_transposed(normalToWorld, screenToTile.y), transform_transposed(normalToWorld, screenToTile.z));
	store(out.screenDepthToLightSpace, toLight);
	store(out.lightSpaceToScreenDepth, inverse(toLight));
}

// The camera transform for each cube side (ref: spriteAPI.cpp:330-337) and the rotation of each sprite direction (:340-349)
static M3 cube_side(int s) {
	static const float forward[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
	static const float up[6][3] = {{0, 1, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 1}, {0, 1, 0}, {0, 1, 0}};
	return make_axis_system(f3(forward[s]), f3(up[s]));
}
static M3 sprite_direction(int d) {
	static const float forward[8][3] = {{0, 0, 1}, {1, 0, 1}, {1, 0, 0}, {1, 0, -1}, {0, 0, -1}, {-1, 0, -1}, {-1, 0, 0}, {-1, 0, 1}};
	return make_axis_system(f3(forward[d]), f3(0.0f, 1.0f, 0.0f));
}

// ------------------------------------------------------------------------------------------------ dense models (device)

struct DenseSetup { // one front-facing triangle prepared for a 16x16 pixel tile
	float ax, ay;                 // cornerA
	float m0, m1, m2, m3;         // offsetToWeight: xAxis.x, xAxis.y, yAxis.x, yAxis.y
	float za, zb, zc;
	float color[9], normal[9];    // A, B, C
	int32_t l, t, r, b;
};

struct DenseParams {
	T3 objectToScreen;
	M3 modelToNormal;
	int32_t clipWidth, clipHeight;
	int32_t regionLeft, regionTop, regionRight, regionBottom;
	int32_t triangleCount, highQuality;
	dfpsr_image height, diffuse, normal;
};

static const int DENSE_TILE = 16, DENSE_THREADS = DENSE_TILE * DENSE_TILE;

// One CTA per 16x16 pixel tile. Triangles are taken 256 at a time: thread i transforms triangle i of the chunk and, when its bound
// touches the tile, appends it (in order) to a shared-memory list; then every pixel thread walks the list in triangle order, so ties
// in height resolve like the reference's sequential loop (first triangle wins, `height > *heightPixel`).
template <bool HIGH_QUALITY>
__global__ void __launch_bounds__(DENSE_THREADS) dense_model_kernel(DenseParams p, const dfpsr_dense_triangle *__restrict__ triangles) {
	__shared__ DenseSetup sList[DENSE_THREADS];
	__shared__ int32_t sWarpCount[DENSE_THREADS / 32];
	const int32_t tileLeft = p.regionLeft + (int32_t)blockIdx.x * DENSE_TILE, tileTop = p.regionTop + (int32_t)blockIdx.y * DENSE_TILE;
	const int32_t px = tileLeft + (int32_t)(threadIdx.x % DENSE_TILE), py = tileTop + (int32_t)(threadIdx.x / DENSE_TILE);
	const bool inside = px < p.regionRight && py < p.regionBottom;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	float *heightPixel = nullptr;
	float bestHeight = 0.0f;
	uint32_t bestDiffuse = 0u, bestNormal = 0u;
	bool written = false;
	if (inside) {
		heightPixel = row_ptr<float>(p.height.data, p.height.stride, py) + px;
		bestHeight = *heightPixel;
	}
	for (int32_t base = 0; base < p.triangleCount; base += DENSE_THREADS) {
		const int32_t index = base + (int32_t)threadIdx.x;
		DenseSetup s;
		bool keep = false;
		if (index < p.triangleCount) {
			const dfpsr_dense_triangle tri = triangles[index];
			const F3 a = transform_point(p.objectToScreen, f3(tri.posA)), b = transform_point(p.objectToScreen, f3(tri.posB)), c = transform_point(p.objectToScreen, f3(tri.posC));
			// ref: spriteAPI.cpp:1144-1156 getBackCulledTriangleBound
			if (!(((c.x - a.x) * (b.y - a.y)) + ((c.y - a.y) * (a.x - b.x)) >= 0.0f)) {
				const float minX = fminf(fminf(a.x, b.x), c.x), minY = fminf(fminf(a.y, b.y), c.y);
				const float maxX = fmaxf(fmaxf(a.x, b.x), c.x), maxY = fmaxf(fmaxf(a.y, b.y), c.y);
				int32_t l = f2i(minX), t = f2i(minY), r = f2i(maxX) + 1, bo = f2i(maxY) + 1;
				// IRect::cut with the image bound (ref: math/IRect.h:56-66)
				if (l < p.clipWidth && r > 0 && t < p.clipHeight && bo > 0) {
					l = max(l, 0); t = max(t, 0); r = min(r, p.clipWidth); bo = min(bo, p.clipHeight);
					if (r > l && bo > t && l < tileLeft + DENSE_TILE && r > tileLeft && t < tileTop + DENSE_TILE && bo > tileTop) {
						keep = true;
						s.l = l; s.t = t; s.r = r; s.b = bo;
						s.ax = a.x; s.ay = a.y;
						// inverse(FMatrix2x2(cornerB - cornerA, cornerC - cornerA)) (ref: math/FMatrix2x2.h:69-76)
						const float xx = b.x - a.x, xy = b.y - a.y, yx = c.x - a.x, yy = c.y - a.y;
						const float inv = 1.0f / (xx * yy - xy * yx);
						s.m0 = yy * inv; s.m1 = -xy * inv; s.m2 = -yx * inv; s.m3 = xx * inv;
						s.za = a.z; s.zb = b.z; s.zc = c.z;
						const F3 na = transform(p.modelToNormal, f3(tri.normalA)), nb = transform(p.modelToNormal, f3(tri.normalB)), nc = transform(p.modelToNormal, f3(tri.normalC));
						s.normal[0] = na.x; s.normal[1] = na.y; s.normal[2] = na.z; s.normal[3] = nb.x; s.normal[4] = nb.y; s.normal[5] = nb.z; s.normal[6] = nc.x; s.normal[7] = nc.y; s.normal[8] = nc.z;
#pragma unroll
						for (int k = 0; k < 3; k++) { s.color[k] = tri.colorA[k]; s.color[3 + k] = tri.colorB[k]; s.color[6 + k] = tri.colorC[k]; }
					}
				}
			}
		}
		// ordered compaction of the kept triangles
		const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
		if (lane == 0) { sWarpCount[warp] = __popc(ballot); }
		__syncthreads();
		int32_t offset = 0, total = 0;
#pragma unroll
		for (int w = 0; w < DENSE_THREADS / 32; w++) { const int32_t n = sWarpCount[w]; if (w < warp) { offset += n; } total += n; }
		if (keep) { sList[offset + __popc(ballot & ((1u << lane) - 1u))] = s; }
		__syncthreads();
		if (inside) {
			for (int32_t i = 0; i < total; i++) {
				const DenseSetup &q = sList[i];
				if (px < q.l || px >= q.r || py < q.t || py >= q.b) { continue; }
				// ref: spriteAPI.cpp:1296-1313
				const float ox = ((float)px + 0.5f) - q.ax, oy = ((float)py + 0.5f) - q.ay;
				const float wb = ox * q.m0 + oy * q.m2, wc = ox * q.m1 + oy * q.m3;
				const float wa = 1.0f - (wb + wc);
				if (wa >= -0.00001f && wb >= -0.00001f && wc >= -0.00001f) {
					const float h = q.za * wa + q.zb * wb + q.zc * wc;
					if (h > bestHeight) {
						bestHeight = h;
						written = true;
						const float cr = q.color[0] * wa + q.color[3] * wb + q.color[6] * wc;
						const float cg = q.color[1] * wa + q.color[4] * wb + q.color[7] * wc;
						const float cb = q.color[2] * wa + q.color[5] * wb + q.color[8] * wc;
						bestDiffuse = f2u(cr) | (f2u(cg) << 8) | (f2u(cb) << 16) | (255u << 24);
						F3 n = f3(q.normal[0] * wa + q.normal[3] * wb + q.normal[6] * wc, q.normal[1] * wa + q.normal[4] * wb + q.normal[7] * wc, q.normal[2] * wa + q.normal[5] * wb + q.normal[8] * wc);
						if (HIGH_QUALITY) { n = normalize(n); }
						bestNormal = f2u((n.x + 1.0f) * 127.5f) | (f2u((n.y + 1.0f) * 127.5f) << 8) | (f2u((n.z + 1.0f) * 127.5f) << 16) | (255u << 24);
					}
				}
			}
		}
		__syncthreads();
	}
	if (written) {
		*heightPixel = bestHeight;
		row_ptr<uint32_t>(p.diffuse.data, p.diffuse.stride, py)[px] = bestDiffuse;
		row_ptr<uint32_t>(p.normal.data, p.normal.stride, py)[px] = bestNormal;
	}
}

// ---- scaleHeightImage (ref: spriteAPI.cpp:157-174): F32 heights from the atlas' height column, -inf where the colour is transparent
__global__ void __launch_bounds__(256) scale_height_kernel(const uint32_t *__restrict__ atlas, int32_t atlasStridePixels, int32_t colorLeft, int32_t heightLeft, int32_t top,
                                                           int32_t width, int32_t height, float scale, float offset, float *__restrict__ out) {
	const int32_t x = (int32_t)(blockIdx.x * 32u + (threadIdx.x & 31u)), y = (int32_t)(blockIdx.y * 8u + (threadIdx.x >> 5));
	if (x >= width || y >= height) { return; }
	const uint32_t h = atlas[(size_t)(top + y) * atlasStridePixels + heightLeft + x], c = atlas[(size_t)(top + y) * atlasStridePixels + colorLeft + x];
	const float value = (float)(h & 255u);
	out[(size_t)y * width + x] = ((c >> 24) > 127u) ? (value * scale) + offset : -INFINITY;
}

// ---- block clear: diffuse = 0, normal = 0x80808080, height = bottom clip plane (ref: spriteAPI.cpp:521-523, :548)
__global__ void __launch_bounds__(256) block_clear_kernel(uint4 *__restrict__ diffuse, uint4 *__restrict__ normal, float4 *__restrict__ height, int32_t quads) {
	const int32_t i = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= quads) { return; }
	diffuse[i] = make_uint4(0u, 0u, 0u, 0u);
	normal[i] = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
	height[i] = make_float4(BOTTOM_CLIP_PLANE, BOTTOM_CLIP_PLANE, BOTTOM_CLIP_PLANE, BOTTOM_CLIP_PLANE);
}

// ---- all background copies of a frame (ref: spriteAPI.cpp:554-562 BackgroundBlock::draw = 3 x draw_copy, :673-688 per dirty rectangle)
struct CopyDev {
	const uint32_t *diffuse, *normal; const float *height; // block images, BLOCK_SIZE pixels per row
	int32_t left, top, width, height_, sourceLeft, sourceTop;
};
__global__ void __launch_bounds__(256) background_copy_kernel(const CopyDev *__restrict__ copies, dfpsr_image diffuse, dfpsr_image normal, dfpsr_image height) {
	const CopyDev c = copies[blockIdx.z];
	const int32_t x = (int32_t)(blockIdx.x * 32u + (threadIdx.x & 31u)), y = (int32_t)(blockIdx.y * 8u + (threadIdx.x >> 5));
	if (x >= c.width || y >= c.height_) { return; }
	const size_t src = (size_t)(c.sourceTop + y) * BLOCK_SIZE + (size_t)(c.sourceLeft + x);
	row_ptr<uint32_t>(diffuse.data, diffuse.stride, c.top + y)[c.left + x] = c.diffuse[src];
	row_ptr<uint32_t>(normal.data, normal.stride, c.top + y)[c.left + x] = c.normal[src];
	row_ptr<float>(height.data, height.stride, c.top + y)[c.left + x] = c.height[src];
}

// ---- sprite baking (ref: spriteAPI.cpp:1329-1432 sprite_generateFromModel)
// height image of one camera angle: red = (height - minY) * 255 / (maxY - minY) saturated to a byte, alpha = the colour's alpha (:1375-1381)
__global__ void __launch_bounds__(256) bake_height_kernel(const float *__restrict__ depth, const uint32_t *__restrict__ color, uint32_t *__restrict__ heightImage, int32_t pixels, float minY, float heightScale) {
	const int32_t i = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= pixels) { return; }
	int32_t h = f2i((depth[i] - minY) * heightScale);
	h = h < 0 ? 0 : (h > 255 ? 255 : h);
	heightImage[i] = (uint32_t)h | (color[i] & 0xFF000000u);
}
// bound of the pixels with a non-zero alpha over all angles (:1385-1401): crop = {minX, minY, maxX, maxY}
__global__ void __launch_bounds__(256) bake_crop_kernel(const uint32_t *__restrict__ color, int32_t width, int32_t height, int32_t angles, int32_t *crop) {
	const int32_t x = (int32_t)(blockIdx.x * 32u + (threadIdx.x & 31u)), y = (int32_t)(blockIdx.y * 8u + (threadIdx.x >> 5));
	bool any = false;
	if (x < width && y < height) {
		for (int32_t a = 0; a < angles; a++) { any = any || (color[((size_t)a * height + y) * width + x] >> 24) != 0u; }
	}
	if (any) { atomicMin(crop + 0, x); atomicMin(crop + 1, y); atomicMax(crop + 2, x); atomicMax(crop + 3, y); }
}

// ------------------------------------------------------------------------------------------------ types (process global, like the reference)

struct DeviceModel { // shadow model resident on the device (lazily)
	std::vector<float> points;
	std::vector<dfpsr_polygon> polygons;
	float minBound[3] = {0, 0, 0}, maxBound[3] = {0, 0, 0};
	void *dPoints = nullptr, *dPolygons = nullptr;
	dfpsr_model desc;
	bool exists() const { return !polygons.empty() || !points.empty(); }
	void set_bounds() { // ref: implementation/render/model/Model.cpp:281-288 — both bounds start at the origin
		for (int k = 0; k < 3; k++) { minBound[k] = 0.0f; maxBound[k] = 0.0f; }
		for (size_t i = 0; i + 2 < points.size(); i += 3) {
			for (int k = 0; k < 3; k++) { if (points[i + k] < minBound[k]) { minBound[k] = points[i + k]; } if (points[i + k] > maxBound[k]) { maxBound[k] = points[i + k]; } }
		}
	}
	int ensure_device() {
		if (dPoints || !exists()) { return 0; }
		DFPSR_CHECK_CUDA(cudaMalloc(&dPoints, std::max<size_t>(points.size() * sizeof(float), 16)));
		DFPSR_CHECK_CUDA(cudaMalloc(&dPolygons, std::max<size_t>(polygons.size() * sizeof(dfpsr_polygon), 16)));
		DFPSR_CHECK_CUDA(cudaMemcpy(dPoints, points.data(), points.size() * sizeof(float), cudaMemcpyHostToDevice));
		DFPSR_CHECK_CUDA(cudaMemcpy(dPolygons, polygons.data(), polygons.size() * sizeof(dfpsr_polygon), cudaMemcpyHostToDevice));
		memset(&desc, 0, sizeof(desc));
		desc.points = (const float *)dPoints; desc.pointCount = (int32_t)(points.size() / 3);
		desc.polygons = (const dfpsr_polygon *)dPolygons; desc.polygonCount = (int32_t)polygons.size();
		desc.filter = DFPSR_FILTER_SOLID;
		for (int k = 0; k < 3; k++) { desc.minBound[k] = minBound[k]; desc.maxBound[k] = maxBound[k]; }
		return 0;
	}
};

struct SpriteType {
	I3 minBoundMini, maxBoundMini;
	int32_t centerX = 0, centerY = 0, frameWidth = 0, frameHeight = 0, frameCount = 0, propertyColumns = 0;
	int32_t atlasWidth = 0, atlasHeight = 0;
	float heightScale = 0.0f, heightOffset = 0.0f;
	std::vector<uint32_t> atlas; // host copy, tightly packed rows
	DeviceModel shadow;
	// device residency (lazy)
	uint32_t *dAtlas = nullptr;
	float *dHeights = nullptr; // frameCount images of frameWidth x frameHeight
	int ensure_device(cudaStream_t stream);
	int32_t frame_index(int32_t direction) const { // ref: spriteAPI.cpp:227-231
		static const int32_t frameFromDir[8] = {4, 1, 5, 2, 6, 3, 7, 0};
		return frameFromDir[correct_direction(direction)] % frameCount;
	}
};

struct ModelType {
	std::vector<dfpsr_dense_triangle> triangles;
	float minBound[3], maxBound[3];
	DeviceModel shadow;
	dfpsr_dense_triangle *dTriangles = nullptr;
	int ensure_device() {
		if (dTriangles || triangles.empty()) { return shadow.ensure_device(); }
		DFPSR_CHECK_CUDA(cudaMalloc((void **)&dTriangles, triangles.size() * sizeof(dfpsr_dense_triangle)));
		DFPSR_CHECK_CUDA(cudaMemcpy(dTriangles, triangles.data(), triangles.size() * sizeof(dfpsr_dense_triangle), cudaMemcpyHostToDevice));
		return shadow.ensure_device();
	}
};

// deque-like stability is not needed: worlds refer to types by index
static std::vector<std::unique_ptr<SpriteType>> g_spriteTypes;
static std::vector<std::unique_ptr<ModelType>> g_modelTypes;

int SpriteType::ensure_device(cudaStream_t stream) {
	if (dAtlas) { return shadow.ensure_device(); }
	DFPSR_CHECK_CUDA(cudaMalloc((void **)&dAtlas, atlas.size() * sizeof(uint32_t)));
	DFPSR_CHECK_CUDA(cudaMemcpy(dAtlas, atlas.data(), atlas.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	DFPSR_CHECK_CUDA(cudaMalloc((void **)&dHeights, std::max<size_t>((size_t)frameCount * frameWidth * frameHeight * sizeof(float), 16)));
	for (int32_t f = 0; f < frameCount; f++) {
		dim3 grid((unsigned)((frameWidth + 31) / 32), (unsigned)((frameHeight + 7) / 8));
		DFPSR_LAUNCH(scale_height_kernel, grid, 256, 0, stream, dAtlas, atlasWidth, 0, frameWidth, f * frameHeight, frameWidth, frameHeight, heightScale, heightOffset,
		             dHeights + (size_t)f * frameWidth * frameHeight);
	}
	return shadow.ensure_device();
}

static dfpsr_image atlas_image(const SpriteType &t, int32_t column, int32_t frame) {
	dfpsr_image im;
	im.data = t.dAtlas + (size_t)frame * t.frameHeight * t.atlasWidth + (size_t)column * t.frameWidth;
	im.width = t.frameWidth; im.height = t.frameHeight; im.stride = t.atlasWidth * 4; im.packOrder = DFPSR_PACK_RGBA;
	return im;
}
static dfpsr_image height_image(const SpriteType &t, int32_t frame) {
	dfpsr_image im;
	im.data = t.dHeights + (size_t)frame * t.frameWidth * t.frameHeight;
	im.width = t.frameWidth; im.height = t.frameHeight; im.stride = t.frameWidth * 4; im.packOrder = DFPSR_PACK_RGBA;
	return im;
}

// ------------------------------------------------------------------------------------------------ octree (ref: SDK/SpriteEngine/Octree.h)

template <typename T>
struct Leaf { T content; I3 origin, mn, mx; };

template <typename T>
struct Node {
	I3 ownedMin, ownedMax, leafMin, leafMax;
	bool divided = false;
	std::unique_ptr<Node<T>> child[8];
	std::vector<Leaf<T>> leaves;

	bool inside_owned(I3 o) const { return o.x >= ownedMin.x && o.y >= ownedMin.y && o.z >= ownedMin.z && o.x <= ownedMax.x && o.y <= ownedMax.y && o.z <= ownedMax.z; }
	static int branch_index(bool px, bool py, bool pz) { return (px ? 1 : 0) | (py ? 2 : 0) | (pz ? 4 : 0); }
	bool may_branch(const Leaf<T> &leaf) const { // ref: Octree.h:124-131
		return divided && leaf.mx.x - leaf.mn.x <= (ownedMax.x - ownedMin.x) / 4 && leaf.mx.y - leaf.mn.y <= (ownedMax.y - ownedMin.y) / 4 && leaf.mx.z - leaf.mn.z <= (ownedMax.z - ownedMin.z) / 4;
	}
	int insert(const Leaf<T> &leaf, int depth = 0) { // ref: Octree.h:132-186
		leafMin.x = std::min(leafMin.x, leaf.mn.x); leafMin.y = std::min(leafMin.y, leaf.mn.y); leafMin.z = std::min(leafMin.z, leaf.mn.z);
		leafMax.x = std::max(leafMax.x, leaf.mx.x); leafMax.y = std::max(leafMax.y, leaf.mx.y); leafMax.z = std::max(leafMax.z, leaf.mx.z);
		while (!inside_owned(leaf.origin)) {
			if (ownedMin.x < -100000000 || ownedMax.x > 100000000) { set_error("octree: cannot expand to include the origin (%d, %d, %d)", leaf.origin.x, leaf.origin.y, leaf.origin.z); return 1; }
			// the old node becomes the inner child of a parent twice its size; this object keeps its identity as the parent
			const int inner = branch_index(ownedMin.x + ownedMax.x < 0, ownedMin.y + ownedMax.y < 0, ownedMin.z + ownedMax.z < 0);
			std::unique_ptr<Node<T>> old(new Node<T>());
			old->ownedMin = ownedMin; old->ownedMax = ownedMax; old->leafMin = leafMin; old->leafMax = leafMax; old->divided = divided;
			for (int n = 0; n < 8; n++) { old->child[n] = std::move(child[n]); }
			old->leaves.swap(leaves);
			ownedMin = i3(ownedMin.x * 2, ownedMin.y * 2, ownedMin.z * 2); ownedMax = i3(ownedMax.x * 2, ownedMax.y * 2, ownedMax.z * 2);
			divided = true;
			child[inner] = std::move(old);
		}
		for (int n = 0; n < 8; n++) {
			if (child[n] && child[n]->inside_owned(leaf.origin)) { return child[n]->insert(leaf, depth + 1); }
		}
		if (may_branch(leaf)) {
			const I3 middle = i3((ownedMin.x + ownedMax.x) / 2, (ownedMin.y + ownedMax.y) / 2, (ownedMin.z + ownedMax.z) / 2);
			const int index = branch_index(leaf.origin.x >= middle.x, leaf.origin.y >= middle.y, leaf.origin.z >= middle.z);
			const I3 size = i3((ownedMax.x - ownedMin.x) / 2, (ownedMax.y - ownedMin.y) / 2, (ownedMax.z - ownedMin.z) / 2); // splitBound, Octree.h:56-66
			std::unique_ptr<Node<T>> fresh(new Node<T>());
			fresh->ownedMin = i3(ownedMin.x + ((index & 1) ? size.x : 0), ownedMin.y + ((index & 2) ? size.y : 0), ownedMin.z + ((index & 4) ? size.z : 0));
			fresh->ownedMax = i3(fresh->ownedMin.x + size.x, fresh->ownedMin.y + size.y, fresh->ownedMin.z + size.z);
			fresh->leafMin = leaf.mn; fresh->leafMax = leaf.mx;
			fresh->leaves.push_back(leaf);
			child[index] = std::move(fresh);
		} else {
			leaves.push_back(leaf);
			if (leaves.size() > 64 && depth < 64) {
				divided = true;
				std::vector<Leaf<T>> old;
				old.swap(leaves);
				const size_t before = old.size();
				bool moved = false;
				for (size_t l = 0; l < old.size(); l++) {
					// the reference re-inserts through insert(); a node whose leaves all stay put would recurse forever there, so the
					// last re-insertion must not trigger another split when nothing moved
					const bool last = l + 1 == old.size();
					if (last && !moved && leaves.size() + 1 == before) { leaves.push_back(old[l]); break; }
					const size_t had = leaves.size();
					if (insert(old[l], depth + 1)) { return 1; }
					if (leaves.size() == had) { moved = true; }
				}
			}
		}
		return 0;
	}
	template <typename Filter, typename Operation>
	void find(const Filter &filter, const Operation &operation) { // ref: Octree.h:187-203
		if (!filter(leafMin, leafMax)) { return; }
		for (int32_t l = (int32_t)leaves.size() - 1; l >= 0; l--) {
			Leaf<T> &leaf = leaves[(size_t)l];
			if (filter(leaf.mn, leaf.mx) && operation(leaf.content, leaf.origin, leaf.mn, leaf.mx)) { leaves.erase(leaves.begin() + l); }
		}
		for (int n = 0; n < 8; n++) { if (child[n]) { child[n]->find(filter, operation); } }
	}
};

template <typename T>
struct Octree { // ref: Octree.h:206-255
	std::unique_ptr<Node<T>> side[8];
	int32_t initialSize;
	explicit Octree(int32_t initialSize) : initialSize(initialSize) {}
	int insert(const T &content, I3 origin, I3 mn, I3 mx) {
		Leaf<T> leaf; leaf.content = content; leaf.origin = origin; leaf.mn = mn; leaf.mx = mx;
		const int index = Node<T>::branch_index(origin.x >= 0, origin.y >= 0, origin.z >= 0);
		if (!side[index]) {
			const int32_t required = std::max(std::max(std::max(origin.x, -origin.x), std::max(origin.y, -origin.y)), std::max(origin.z, -origin.z));
			int32_t size = initialSize;
			while (size < required) { size *= 2; }
			std::unique_ptr<Node<T>> fresh(new Node<T>());
			fresh->ownedMin = i3(origin.x < 0 ? -size : 0, origin.y < 0 ? -size : 0, origin.z < 0 ? -size : 0);
			fresh->ownedMax = i3(origin.x < 0 ? 0 : size, origin.y < 0 ? 0 : size, origin.z < 0 ? 0 : size);
			fresh->leafMin = mn; fresh->leafMax = mx;
			fresh->leaves.push_back(leaf);
			side[index] = std::move(fresh);
			return 0;
		}
		return side[index]->insert(leaf);
	}
	template <typename Filter, typename Operation>
	void map(const Filter &filter, const Operation &operation) {
		for (int n = 0; n < 8; n++) { if (side[n]) { side[n]->find(filter, operation); } }
	}
	template <typename Operation>
	void map_box(I3 searchMin, I3 searchMax, const Operation &operation) {
		map([searchMin, searchMax](I3 mn, I3 mx) {
			return searchMax.x >= mn.x && searchMin.x <= mx.x && searchMax.y >= mn.y && searchMin.y <= mx.y && searchMax.z >= mn.z && searchMin.z <= mx.z;
		}, operation);
	}
};

// ------------------------------------------------------------------------------------------------ dirty rectangles (ref: DirtyRectangles.h)

struct DirtyRectangles {
	int32_t width = 0, height = 0;
	std::vector<Rect> rects;
	Rect bound() const { return Rect(0, 0, width, height); }
	void all_dirty() { rects.clear(); rects.push_back(bound()); }
	void none_dirty() { rects.clear(); }
	void set_target_resolution(int32_t w, int32_t h) { if (width != w || height != h) { width = w; height = h; all_dirty(); } }
	void make_region_dirty(Rect region) {
		region = Rect::cut(region, bound());
		if (!region.has_area()) { return; }
		for (int32_t i = 0; i < (int32_t)rects.size(); i++) {
			if (Rect::touches(rects[(size_t)i], region)) {
				region = Rect::merge(region, rects[(size_t)i]);
				rects.erase(rects.begin() + i);
				i = -1;
			}
		}
		rects.push_back(region);
	}
};

// ------------------------------------------------------------------------------------------------ the world

struct PointLightRec { float position[3]; float radius, intensity; int32_t color[3]; int32_t shadowCasting; };
struct DirectedLightRec { float direction[3]; float intensity; int32_t color[3]; };

struct Block { // ref: spriteAPI.cpp:504-569 BackgroundBlock
	Rect worldRegion;
	int32_t cameraId = 0;
	int32_t state = 0; // 0 unused, 1 ready, 2 dirty
	uint32_t *dDiffuse = nullptr, *dNormal = nullptr; float *dHeight = nullptr;
};

struct DeviceImage {
	void *ptr = nullptr; int32_t width = 0, height = 0, stride = 0;
	int ensure(int32_t w, int32_t h) {
		if (ptr && w == width && h == height) { return 0; }
		if (ptr) { cudaFree(ptr); ptr = nullptr; }
		stride = ((w * 4 + 15) / 16) * 16;
		DFPSR_CHECK_CUDA(cudaMalloc(&ptr, std::max<size_t>((size_t)stride * h, 16)));
		DFPSR_CHECK_CUDA(cudaMemset(ptr, 0, (size_t)stride * h));
		width = w; height = h;
		return 0;
	}
	dfpsr_image image() const { dfpsr_image im; im.data = ptr; im.width = width; im.height = height; im.stride = stride; im.packOrder = DFPSR_PACK_RGBA; return im; }
};

} // namespace sw
} // namespace dfpsr

using namespace dfpsr;
using namespace dfpsr::sw;

struct dfpsr_sprite_world {
	dfpsr_ortho_system ortho;
	Octree<dfpsr_sprite_instance> passiveSprites;
	Octree<dfpsr_model_instance> passiveModels;
	std::vector<dfpsr_sprite_instance> temporarySprites;
	std::vector<dfpsr_model_instance> temporaryModels;
	std::vector<PointLightRec> pointLights;
	std::vector<DirectedLightRec> directedLights;
	int32_t cameraIndex = 0;
	I3 cameraLocation = {0, 0, 0};
	int32_t width = 0, height = 0; // size of the deferred buffers (0 before the first frame)
	std::vector<Block> blocks;
	DirtyRectangles dirty;
	int32_t shadowResolution;
	std::vector<dfpsr_sprite_world_op> ops;
	// device side
	DeviceImage diffuse, normal, light, heightBuffer;
	// Shadow cube maps, one width x 6 width F32 image per shadow-casting light of the frame, in two contiguous pools:
	//   cubeStatic   the light's PASSIVE casters only, kept from frame to frame (the reference reuses ONE cube map for all lights and
	//                therefore renders every caster of every light in every frame; 1.5 MB per light buys not doing that)
	//   cubeWorking  static + this frame's temporary casters: a copy of the static map with the temporary casters rendered on top
	//                (depth-only rendering keeps the maximum, so the result does not depend on the split)
	// staticSignature[c] says what cubeStatic[c] holds: resolution, the view's normal-to-world matrix and every passive caster
	// (op, type, transform relative to the light), byte for byte; a light whose list differs is rendered again.
	DeviceBuffer cubeStatic, cubeWorking;
	int32_t cubeCapacity = 0;
	std::vector<std::vector<uint8_t>> staticSignature;
	DeviceBuffer copyStaging;
	dfpsr_sprite_world(const dfpsr_ortho_system &o, int32_t shadowResolution)
	: ortho(o), passiveSprites(MINI_UNITS_PER_TILE * 64), passiveModels(MINI_UNITS_PER_TILE * 64), shadowResolution(shadowResolution) {}
	const dfpsr_ortho_camera &view() const { return ortho.view[cameraIndex]; }
};

// ---- host planning -----------------------------------------------------------------------------------------------------------------

static dfpsr_sprite_world_op make_op(int32_t kind) { dfpsr_sprite_world_op op; memset(&op, 0, sizeof(op)); op.op = kind; op.block = -1; return op; }

// ref: spriteAPI.cpp:837-847 + :856-874 get3DBounds
static void get_3d_bounds(const T3 &transform, F3 localMin, F3 localMax, I3 &worldMin, I3 &worldMax) {
	worldMin = i3(f2i(transform.position.x), f2i(transform.position.y), f2i(transform.position.z));
	worldMax = worldMin;
	for (int c = 0; c < 8; c++) {
		const F3 corner = transform_point(transform, f3((c & 1) ? localMax.x : localMin.x, (c & 2) ? localMax.y : localMin.y, (c & 4) ? localMax.z : localMin.z));
		worldMin.x = std::min(worldMin.x, (int32_t)floor((double)corner.x)); worldMin.y = std::min(worldMin.y, (int32_t)floor((double)corner.y)); worldMin.z = std::min(worldMin.z, (int32_t)floor((double)corner.z));
		worldMax.x = std::max(worldMax.x, (int32_t)ceil((double)corner.x)); worldMax.y = std::max(worldMax.y, (int32_t)ceil((double)corner.y)); worldMax.z = std::max(worldMax.z, (int32_t)ceil((double)corner.z));
	}
}

// ref: spriteAPI.cpp:882-898 getScreenBounds
static Rect screen_bounds(const dfpsr_sprite_world *w, I3 worldMin, I3 worldMax) {
	const T3 worldToPixels = t3(f3(0.0f, 0.0f, 0.0f), m3(w->view().worldSpaceToScreenDepth));
	const F3 mn = scale(f3((float)worldMin.x, (float)worldMin.y, (float)worldMin.z), TILES_PER_MINI_UNIT), mx = scale(f3((float)worldMax.x, (float)worldMax.y, (float)worldMax.z), TILES_PER_MINI_UNIT);
	int32_t l = 0, t = 0, r = 0, b = 0;
	for (int c = 0; c < 8; c++) {
		const F3 p = transform_point(worldToPixels, f3((c & 1) ? mx.x : mn.x, (c & 2) ? mx.y : mn.y, (c & 4) ? mx.z : mn.z));
		const int32_t fl = (int32_t)floor((double)p.x), ft = (int32_t)floor((double)p.y), cr = (int32_t)ceil((double)p.x), cb = (int32_t)ceil((double)p.y);
		if (c == 0) { l = fl; t = ft; r = cr; b = cb; }
		else { l = std::min(l, fl); t = std::min(t, ft); r = std::max(r, cr); b = std::max(b, cb); }
	}
	return Rect(l, t, r - l, b - t);
}

// ref: spriteAPI.cpp:738-752 updatePassiveRegion + :626-639 invalidateBlockAt
static void update_passive_region(dfpsr_sprite_world *w, const Rect &region) {
	const int64_t left = round_down(region.l, BLOCK_SIZE), top = round_down(region.t, BLOCK_SIZE);
	const int64_t right = round_down(region.right() - 1, BLOCK_SIZE), bottom = round_down(region.bottom() - 1, BLOCK_SIZE);
	for (int64_t y = top; y <= bottom; y += BLOCK_SIZE) {
		for (int64_t x = left; x <= right; x += BLOCK_SIZE) {
			for (Block &b : w->blocks) { if (b.state == 1 && b.worldRegion.l == (int32_t)x && b.worldRegion.t == (int32_t)y) { b.state = 2; } }
		}
	}
	w->dirty.all_dirty();
}

// ref: spriteAPI.cpp:452-501 orthoCullingTest
static bool ortho_culling_test(const dfpsr_ortho_camera &view, I3 mn, I3 mx, const Rect &seen) {
	I2 c[8];
	for (int k = 0; k < 8; k++) { c[k] = mini_offset_to_pixel(view, i3((k & 1) ? mx.x : mn.x, (k & 2) ? mx.y : mn.y, (k & 4) ? mx.z : mn.z)); }
	bool allLeft = true, allRight = true, allAbove = true, allBelow = true;
	for (int k = 0; k < 8; k++) {
		allLeft = allLeft && c[k].x < seen.l; allRight = allRight && c[k].x > seen.right();
		allAbove = allAbove && c[k].y < seen.t; allBelow = allBelow && c[k].y > seen.bottom();
	}
	return !(allLeft || allRight || allAbove || allBelow);
}

// ref: spriteAPI.cpp:306-323 drawSprite: placement of a sprite frame relative to a target whose pixel (0, 0) is world pixel -worldCenter
static dfpsr_sprite_world_op sprite_op(int32_t kind, const dfpsr_sprite_instance &sprite, const dfpsr_ortho_camera &view, I2 worldCenter) {
	const SpriteType &type = *g_spriteTypes[(size_t)sprite.typeIndex];
	dfpsr_sprite_world_op op = make_op(kind);
	op.typeIndex = sprite.typeIndex;
	op.frame = type.frame_index(view.worldDirection + sprite.direction);
	const I2 pixel = mini_offset_to_pixel(view, i3(sprite.location[0], sprite.location[1], sprite.location[2]));
	op.left = pixel.x + worldCenter.x - type.centerX; op.top = pixel.y + worldCenter.y - type.centerY;
	op.width = type.frameWidth; op.height = type.frameHeight;
	op.heightOffset = (float)sprite.location[1] * TILES_PER_MINI_UNIT;
	return op;
}

// ref: spriteAPI.cpp:1243-1260 — the part of renderDenseModel that runs before any pixel: transform, pessimistic bound, culling
static Rect dense_pessimistic_bound(const float *minBound, const float *maxBound, const T3 &objectToScreen) {
	Rect result;
	for (int c = 0; c < 8; c++) { // transformCorners order (spriteAPI.cpp:838-847)
		const F3 p = transform_point(objectToScreen, f3((c & 1) ? maxBound[0] : minBound[0], (c & 2) ? maxBound[1] : minBound[1], (c & 4) ? maxBound[2] : minBound[2]));
		const Rect one(f2i(p.x), f2i(p.y), 1, 1);
		result = c == 0 ? one : Rect::merge(result, one);
	}
	return result;
}
static T3 dense_object_to_screen(const dfpsr_ortho_camera &view, const float *worldOrigin, const dfpsr_transform3d &modelToWorld) {
	return mul(t3(modelToWorld), t3(f3(worldOrigin[0], worldOrigin[1], 0.0f), m3(view.worldSpaceToScreenDepth))); // spriteAPI.cpp:27-33
}

static dfpsr_sprite_world_op model_op(int32_t kind, const dfpsr_model_instance &model, I2 worldCenter) {
	dfpsr_sprite_world_op op = make_op(kind);
	op.typeIndex = model.typeIndex;
	op.worldOrigin[0] = (float)worldCenter.x; op.worldOrigin[1] = (float)worldCenter.y;
	op.transform = model.location;
	return op;
}

// ref: spriteAPI.cpp:520-539 BackgroundBlock::draw
static void plan_block(dfpsr_sprite_world *w, int32_t slot) {
	Block &block = w->blocks[(size_t)slot];
	const dfpsr_ortho_camera &view = w->view();
	dfpsr_sprite_world_op clear = make_op(DFPSR_SW_BLOCK_CLEAR);
	clear.block = slot; clear.left = block.worldRegion.l; clear.top = block.worldRegion.t; clear.width = BLOCK_SIZE; clear.height = BLOCK_SIZE;
	w->ops.push_back(clear);
	const Rect region = block.worldRegion;
	const I2 worldCenter = {-region.l, -region.t};
	auto filter = [&view, region](I3 mn, I3 mx) { return ortho_culling_test(view, mn, mx, region); };
	w->passiveSprites.map(filter, [&](dfpsr_sprite_instance &sprite, I3, I3, I3) {
		dfpsr_sprite_world_op op = sprite_op(DFPSR_SW_BLOCK_SPRITE, sprite, view, worldCenter);
		op.block = slot;
		w->ops.push_back(op);
		return false;
	});
	w->passiveModels.map(filter, [&](dfpsr_model_instance &model, I3, I3, I3) {
		dfpsr_sprite_world_op op = model_op(DFPSR_SW_BLOCK_MODEL, model, worldCenter);
		op.block = slot;
		w->ops.push_back(op);
		return false;
	});
	block.state = 1;
}

// ref: spriteAPI.cpp:572-625 updateBlockAt
static int32_t update_block_at(dfpsr_sprite_world *w, const Rect &blockRegion, const Rect &seen) {
	int32_t unused = -1;
	const int32_t cameraId = w->view().id;
	for (int32_t b = 0; b < (int32_t)w->blocks.size(); b++) {
		Block &cur = w->blocks[(size_t)b];
		if (cur.state != 0) {
			if (cur.cameraId == cameraId) {
				if (cur.worldRegion.l == blockRegion.l && cur.worldRegion.t == blockRegion.t) {
					if (cur.state == 2) {
						cur.worldRegion = blockRegion; cur.cameraId = cameraId;
						plan_block(w, b);
						return 1;
					}
					return 0;
				}
				if (cur.worldRegion.right() < seen.l - BLOCK_MAX_DISTANCE || cur.worldRegion.l > seen.right() + BLOCK_MAX_DISTANCE
				 || cur.worldRegion.bottom() < seen.t - BLOCK_MAX_DISTANCE || cur.worldRegion.t > seen.bottom() + BLOCK_MAX_DISTANCE) {
					cur.state = 0; cur.worldRegion = Rect(); cur.cameraId = -1;
					unused = b;
				}
			} else {
				cur.state = 0; cur.worldRegion = Rect(); cur.cameraId = -1;
				unused = b;
			}
		} else {
			unused = b;
		}
	}
	if (unused < 0) { w->blocks.push_back(Block()); unused = (int32_t)w->blocks.size() - 1; }
	Block &target = w->blocks[(size_t)unused];
	target.worldRegion = blockRegion; target.cameraId = cameraId;
	plan_block(w, unused);
	return 1;
}

// ref: spriteAPI.cpp:640-660 updateBlocks
static int32_t update_blocks(dfpsr_sprite_world *w, const Rect &seen, int32_t maxUpdates) {
	int32_t updates = 0;
	const int64_t left = round_down(seen.l, BLOCK_SIZE), top = round_down(seen.t, BLOCK_SIZE);
	const int64_t right = round_down(seen.right() - 1, BLOCK_SIZE), bottom = round_down(seen.bottom() - 1, BLOCK_SIZE);
	for (int64_t y = top; y <= bottom; y += BLOCK_SIZE) {
		for (int64_t x = left; x <= right; x += BLOCK_SIZE) {
			updates += update_block_at(w, Rect((int32_t)x, (int32_t)y, BLOCK_SIZE, BLOCK_SIZE), seen);
			if (maxUpdates > -1 && updates >= maxUpdates) { return updates; }
		}
	}
	return updates;
}

static I2 find_world_center(const dfpsr_sprite_world *w, int32_t width, int32_t height) { // ref: spriteAPI.cpp:751-753
	const I2 camera = mini_offset_to_pixel(w->view(), w->cameraLocation);
	I2 r; r.x = width / 2 - camera.x; r.y = height / 2 - camera.y;
	return r;
}

// ref: spriteAPI.cpp:661-734 drawDeferred + :754-816 draw
static void plan_frame(dfpsr_sprite_world *w, int32_t width, int32_t height) {
	w->ops.clear();
	const dfpsr_ortho_camera &view = w->view();
	const I2 worldCenter = find_world_center(w, width, height);
	w->width = width; w->height = height;
	const Rect seen(-worldCenter.x, -worldCenter.y, width, height);
	w->dirty.set_target_resolution(width, height);
	const int32_t forced = update_blocks(w, seen, -1);
	if (forced < 1) { update_blocks(w, seen.expanded(128), 1); }
	for (int32_t b = 0; b < (int32_t)w->blocks.size(); b++) {
		const Block &block = w->blocks[(size_t)b];
		if (block.state == 0) { continue; }
		for (const Rect &screenClip : w->dirty.rects) {
			// draw_copy into the sub-image screenClip of the targets at (block - worldClip), clipped to the sub-image (api/drawAPI.cpp:492-538)
			const Rect placed(block.worldRegion.l - seen.l, block.worldRegion.t - seen.t, BLOCK_SIZE, BLOCK_SIZE);
			const Rect hit = Rect::cut(placed, screenClip);
			if (!hit.has_area()) { continue; }
			dfpsr_sprite_world_op op = make_op(DFPSR_SW_COPY_BLOCK);
			op.block = b; op.left = hit.l; op.top = hit.t; op.width = hit.w; op.height = hit.h;
			op.sourceLeft = hit.l - placed.l; op.sourceTop = hit.t - placed.t;
			w->ops.push_back(op);
		}
	}
	w->dirty.none_dirty();
	for (const dfpsr_sprite_instance &sprite : w->temporarySprites) {
		const dfpsr_sprite_world_op op = sprite_op(DFPSR_SW_SPRITE, sprite, view, worldCenter);
		w->ops.push_back(op);
		w->dirty.make_region_dirty(Rect(op.left, op.top, op.width, op.height));
	}
	for (const dfpsr_model_instance &model : w->temporaryModels) {
		const dfpsr_sprite_world_op op = model_op(DFPSR_SW_MODEL, model, worldCenter);
		w->ops.push_back(op);
		const ModelType &type = *g_modelTypes[(size_t)model.typeIndex];
		const Rect bound = dense_pessimistic_bound(type.minBound, type.maxBound, dense_object_to_screen(view, op.worldOrigin, model.location));
		if (Rect::overlaps(bound, Rect(0, 0, width, height))) { w->dirty.make_region_dirty(bound); } // a culled model returns IRect()
	}
	// lights (ref: spriteAPI.cpp:775-814)
	if (!w->directedLights.empty()) {
		for (int32_t i = 0; i < (int32_t)w->directedLights.size(); i++) {
			dfpsr_sprite_world_op op = make_op(DFPSR_SW_LIGHT_DIRECTED);
			op.light = i; op.flag = i == 0 ? 1 : 0;
			w->ops.push_back(op);
		}
	} else {
		w->ops.push_back(make_op(DFPSR_SW_LIGHT_CLEAR));
	}
	const M3 normalToWorld = m3(view.normalToWorldSpace);
	for (int32_t i = 0; i < (int32_t)w->pointLights.size(); i++) {
		const PointLightRec &light = w->pointLights[(size_t)i];
		if (light.shadowCasting) {
			dfpsr_sprite_world_op clear = make_op(DFPSR_SW_SHADOW_CLEAR);
			clear.light = i;
			w->ops.push_back(clear);
			const F3 position = f3(light.position);
			auto sprite_shadow = [&](const dfpsr_sprite_instance &sprite) { // ref: spriteAPI.cpp:389-403 renderSpriteShadow
				if (!sprite.shadowCasting || !g_spriteTypes[(size_t)sprite.typeIndex]->shadow.exists()) { return; }
				dfpsr_sprite_world_op op = make_op(DFPSR_SW_SHADOW_SPRITE);
				op.light = i; op.typeIndex = sprite.typeIndex;
				const F3 tile = f3(mini_to_floating_tile(sprite.location[0]), mini_to_floating_tile(sprite.location[1]), mini_to_floating_tile(sprite.location[2]));
				op.transform = pod(t3(sub(tile, position), sprite_direction(sprite.direction)));
				w->ops.push_back(op);
			};
			auto model_shadow = [&](const dfpsr_model_instance &model) { // ref: spriteAPI.cpp:375-388 renderModelShadow
				if (!g_modelTypes[(size_t)model.typeIndex]->shadow.exists()) { return; }
				dfpsr_sprite_world_op op = make_op(DFPSR_SW_SHADOW_MODEL);
				op.light = i; op.typeIndex = model.typeIndex;
				T3 t = t3(model.location);
				t.position = sub(t.position, position);
				op.transform = pod(t);
				w->ops.push_back(op);
			};
			// ref: spriteAPI.cpp:404-421 renderPassiveShadows
			const I3 center = i3(floating_tile_to_mini(light.position[0]), floating_tile_to_mini(light.position[1]), floating_tile_to_mini(light.position[2]));
			const int32_t reach = floating_tile_to_mini(light.radius);
			const I3 mn = i3(center.x - reach, center.y - reach, center.z - reach), mx = i3(center.x + reach, center.y + reach, center.z + reach);
			w->passiveSprites.map_box(mn, mx, [&](dfpsr_sprite_instance &sprite, I3, I3, I3) { sprite_shadow(sprite); return false; });
			w->passiveModels.map_box(mn, mx, [&](dfpsr_model_instance &model, I3, I3, I3) { model_shadow(model); return false; });
			// temporary casters carry flag = 1: the executor keeps the cube map of a light's PASSIVE casters from frame to frame
			const size_t firstTemporary = w->ops.size();
			for (const dfpsr_sprite_instance &sprite : w->temporarySprites) { sprite_shadow(sprite); }
			for (const dfpsr_model_instance &model : w->temporaryModels) { model_shadow(model); }
			for (size_t k = firstTemporary; k < w->ops.size(); k++) { w->ops[k].flag = 1; }
		}
		dfpsr_sprite_world_op op = make_op(DFPSR_SW_LIGHT_POINT);
		op.light = i; op.flag = light.shadowCasting ? 1 : 0;
		w->ops.push_back(op);
	}
	(void)normalToWorld;
	w->ops.push_back(make_op(DFPSR_SW_BLEND));
}

// ---- device execution ----------------------------------------------------------------------------------------------------------------

static int ensure_block_storage(Block &b) {
	if (b.dDiffuse) { return 0; }
	const size_t pixels = (size_t)BLOCK_SIZE * BLOCK_SIZE;
	DFPSR_CHECK_CUDA(cudaMalloc((void **)&b.dDiffuse, pixels * 4));
	DFPSR_CHECK_CUDA(cudaMalloc((void **)&b.dNormal, pixels * 4));
	DFPSR_CHECK_CUDA(cudaMalloc((void **)&b.dHeight, pixels * 4));
	return 0;
}
static dfpsr_image block_image(void *data) { dfpsr_image im; im.data = data; im.width = BLOCK_SIZE; im.height = BLOCK_SIZE; im.stride = BLOCK_SIZE * 4; im.packOrder = DFPSR_PACK_RGBA; return im; }

static int dense_render(const dfpsr_dense_triangle *dTriangles, int32_t count, const float *minBound, const float *maxBound, const dfpsr_ortho_camera &view, const dfpsr_image &height,
                        const dfpsr_image &diffuse, const dfpsr_image &normal, const float *worldOrigin, const dfpsr_transform3d &modelToWorld, bool highQuality, int32_t *dirtyRect, cudaStream_t stream) {
	const T3 objectToScreen = dense_object_to_screen(view, worldOrigin, modelToWorld);
	const Rect bound = dense_pessimistic_bound(minBound, maxBound, objectToScreen);
	const Rect clip(0, 0, height.width, height.height);
	if (dirtyRect) { dirtyRect[0] = dirtyRect[1] = dirtyRect[2] = dirtyRect[3] = 0; }
	if (!Rect::overlaps(bound, clip)) { return 0; }
	if (dirtyRect) { dirtyRect[0] = bound.l; dirtyRect[1] = bound.t; dirtyRect[2] = bound.w; dirtyRect[3] = bound.h; }
	if (count <= 0) { return 0; }
	// every triangle lies inside the transformed bounding box up to rounding: two extra pixels on each side cover that
	const Rect region = Rect::cut(bound.expanded(2), clip);
	DenseParams p;
	p.objectToScreen = objectToScreen;
	p.modelToNormal = mul(t3(modelToWorld).m, transpose(m3(view.normalToWorldSpace))); // spriteAPI.cpp:1262
	p.clipWidth = clip.w; p.clipHeight = clip.h;
	p.regionLeft = region.l; p.regionTop = region.t; p.regionRight = region.right(); p.regionBottom = region.bottom();
	p.triangleCount = count; p.highQuality = highQuality ? 1 : 0;
	p.height = height; p.diffuse = diffuse; p.normal = normal;
	dim3 grid((unsigned)((region.w + DENSE_TILE - 1) / DENSE_TILE), (unsigned)((region.h + DENSE_TILE - 1) / DENSE_TILE));
	if (highQuality) { DFPSR_LAUNCH(dense_model_kernel<true>, grid, DENSE_THREADS, 0, stream, p, dTriangles); }
	else { DFPSR_LAUNCH(dense_model_kernel<false>, grid, DENSE_THREADS, 0, stream, p, dTriangles); }
	return 0;
}

static int flush_sprites(std::vector<dfpsr_sprite_draw> &draws, const dfpsr_image &height, const dfpsr_image &diffuse, const dfpsr_image &normal, cudaStream_t stream) {
	if (draws.empty()) { return 0; }
	const int status = dfpsr_draw_higher_batch(&height, &diffuse, &normal, draws.data(), (int32_t)draws.size(), stream);
	draws.clear();
	return status;
}

static int execute_frame(dfpsr_sprite_world *w, const dfpsr_image *colorTarget, cudaStream_t stream) {
	const int32_t width = w->width, height = w->height;
	if (w->diffuse.ensure(width, height) || w->normal.ensure(width, height) || w->light.ensure(width, height) || w->heightBuffer.ensure(width, height)) { return 1; }
	const dfpsr_image fDiffuse = w->diffuse.image(), fNormal = w->normal.image(), fLight = w->light.image(), fHeight = w->heightBuffer.image();
	const dfpsr_ortho_camera &view = w->view();
	int32_t shadowLights = 0;
	for (const dfpsr_sprite_world_op &op : w->ops) { // residency of everything the frame touches
		if (op.op == DFPSR_SW_SHADOW_CLEAR) { shadowLights++; }
		if (op.op == DFPSR_SW_BLOCK_SPRITE || op.op == DFPSR_SW_SPRITE || op.op == DFPSR_SW_SHADOW_SPRITE) { if (g_spriteTypes[(size_t)op.typeIndex]->ensure_device(stream)) { return 1; } }
		if (op.op == DFPSR_SW_BLOCK_MODEL || op.op == DFPSR_SW_MODEL || op.op == DFPSR_SW_SHADOW_MODEL) { if (g_modelTypes[(size_t)op.typeIndex]->ensure_device()) { return 1; } }
		if (op.block >= 0 && ensure_block_storage(w->blocks[(size_t)op.block])) { return 1; }
	}
	std::vector<dfpsr_sprite_draw> draws;
	std::vector<CopyDev> copies;
	int32_t drawBlock = -2; // block the pending sprite batch targets (-1 = frame buffers)
	auto targets_of = [&](int32_t block, dfpsr_image &h, dfpsr_image &d, dfpsr_image &n) {
		if (block < 0) { h = fHeight; d = fDiffuse; n = fNormal; }
		else { const Block &b = w->blocks[(size_t)block]; h = block_image(b.dHeight); d = block_image(b.dDiffuse); n = block_image(b.dNormal); }
	};
	auto flush = [&]() -> int {
		if (draws.empty()) { return 0; }
		dfpsr_image h, d, n;
		targets_of(drawBlock, h, d, n);
		return flush_sprites(draws, h, d, n, stream);
	};
	auto flush_copies = [&]() -> int {
		if (copies.empty()) { return 0; }
		if (w->copyStaging.reserve(copies.size() * sizeof(CopyDev))) { return 1; }
		DFPSR_CHECK_CUDA(cudaMemcpyAsync(w->copyStaging.ptr, copies.data(), copies.size() * sizeof(CopyDev), cudaMemcpyHostToDevice, stream));
		int32_t maxW = 0, maxH = 0;
		for (const CopyDev &c : copies) { maxW = std::max(maxW, c.width); maxH = std::max(maxH, c.height_); }
		for (size_t first = 0; first < copies.size(); first += 65535) {
			const size_t n = std::min<size_t>(65535, copies.size() - first);
			dim3 grid((unsigned)((maxW + 31) / 32), (unsigned)((maxH + 7) / 8), (unsigned)n);
			DFPSR_LAUNCH(background_copy_kernel, grid, 256, 0, stream, (const CopyDev *)w->copyStaging.ptr + first, fDiffuse, fNormal, fHeight);
		}
		// `copies` is pageable host memory: cudaMemcpyAsync has staged it before returning, so it may be reused at once. (Waiting for the
		// stream here made every frame start by draining the previous frame's light pass: no overlap between frames.)
		copies.clear();
		return 0;
	};
	// lights of the frame
	std::vector<dfpsr_directed_light> directed;
	std::vector<dfpsr_point_light> points(w->pointLights.size());
	// One depth-only submission: (model, transform, face camera) tuples and the cube maps they draw into (target = position of the
	// cube in `cubes` x 6 + face).
	struct ShadowBatch {
		std::vector<const dfpsr_model *> models;
		std::vector<dfpsr_transform3d> transforms;
		std::vector<dfpsr_camera> cameras;
		std::vector<int32_t> cubeOfTask, faceOfTask;
		std::vector<int32_t> cubes;
	};
	ShadowBatch passiveBatch, temporary;
	std::vector<int32_t> cubeOfLight(w->pointLights.size(), -1);
	std::vector<char> usesWorking(w->pointLights.size(), 0);
	std::vector<uint8_t> signature;
	std::vector<const dfpsr_sprite_world_op *> passiveOps;
	int32_t cubeCount = 0, temporaryOfLight = 0;
	dfpsr_camera faceCameras[6];
	float faceStretch[6] = {1, 1, 1, 1, 1, 1}; // Frobenius norm of each face camera's axis system: bounds how far worldToCamera can stretch a length
	bool haveFaceCameras = false;
	const int32_t res = w->shadowResolution;
	const size_t cubeBytes = std::max<size_t>((size_t)res * res * 6 * 4, 16);
	auto ensure_cubes = [&](int32_t count) -> int {
		if (count <= w->cubeCapacity) { return 0; }
		const int32_t grown = std::max(count, w->cubeCapacity * 2);
		// growing the pools drops every cached static map (DeviceBuffer::reserve does not keep contents)
		if (w->cubeStatic.reserve(cubeBytes * (size_t)grown) || w->cubeWorking.reserve(cubeBytes * (size_t)grown)) { return 1; }
		w->cubeCapacity = grown;
		w->staticSignature.assign((size_t)grown, std::vector<uint8_t>());
		return 0;
	};
	auto ensure_face_cameras = [&]() -> int { // ref: spriteAPI.cpp:383, :397 — Camera::createPerspective(Transform3D(FVector3D(), ShadowCubeMapSides[s] * normalToWorld), res, res)
		if (haveFaceCameras) { return 0; }
		const M3 normalToWorld = m3(view.normalToWorldSpace);
		for (int s = 0; s < 6; s++) {
			const dfpsr_transform3d location = pod(t3(f3(0.0f, 0.0f, 0.0f), mul(cube_side(s), normalToWorld)));
			if (dfpsr_camera_create_perspective(&faceCameras[s], &location, (float)res, (float)res, 1.0f, 0.01f, 1000.0f)) { return 1; }
			const dfpsr_transform3d &l = faceCameras[s].location;
			float sum = 0.0f;
			for (int k = 0; k < 3; k++) { sum += l.xAxis[k] * l.xAxis[k] + l.yAxis[k] * l.yAxis[k] + l.zAxis[k] * l.zAxis[k]; }
			faceStretch[s] = sqrtf(sum);
		}
		haveFaceCameras = true;
		return 0;
	};
	// Queues the cube faces a caster can touch; returns how many. Conservative pre-filter: the model's bounding sphere against each face's
	// cull planes. A face whose frustum the sphere misses by a margin is also missed by the exact box test (dfpsr_camera_is_box_seen inside
	// the batch, ref: api/modelAPI.cpp:228) and by every triangle, so skipping the submission cannot change a pixel; it only spares the
	// host the exact test for the four or five faces of the cube a caster cannot touch.
	auto submit_caster = [&](const dfpsr_sprite_world_op &op, ShadowBatch &batch, int32_t cube) -> int32_t {
		if (ensure_face_cameras()) { return 0; }
		const DeviceModel &model = op.op == DFPSR_SW_SHADOW_SPRITE ? g_spriteTypes[(size_t)op.typeIndex]->shadow : g_modelTypes[(size_t)op.typeIndex]->shadow;
		const dfpsr_transform3d &m = op.transform;
		float centre[3], radius = 0.0f;
		{
			const float *mn = model.desc.minBound, *mx = model.desc.maxBound;
			const float cx = (mn[0] + mx[0]) * 0.5f, cy = (mn[1] + mx[1]) * 0.5f, cz = (mn[2] + mx[2]) * 0.5f;
			for (int k = 0; k < 3; k++) { centre[k] = (cx * m.xAxis[k] + cy * m.yAxis[k] + cz * m.zAxis[k]) + m.position[k]; }
			const float hx = (mx[0] - mn[0]) * 0.5f, hy = (mx[1] - mn[1]) * 0.5f, hz = (mx[2] - mn[2]) * 0.5f;
			const float xx = m.xAxis[0] * m.xAxis[0] + m.xAxis[1] * m.xAxis[1] + m.xAxis[2] * m.xAxis[2];
			const float yy = m.yAxis[0] * m.yAxis[0] + m.yAxis[1] * m.yAxis[1] + m.yAxis[2] * m.yAxis[2];
			const float zz = m.zAxis[0] * m.zAxis[0] + m.zAxis[1] * m.zAxis[1] + m.zAxis[2] * m.zAxis[2];
			const float xy = fabsf(m.xAxis[0] * m.yAxis[0] + m.xAxis[1] * m.yAxis[1] + m.xAxis[2] * m.yAxis[2]);
			const float xz = fabsf(m.xAxis[0] * m.zAxis[0] + m.xAxis[1] * m.zAxis[1] + m.xAxis[2] * m.zAxis[2]);
			const float yz = fabsf(m.yAxis[0] * m.zAxis[0] + m.yAxis[1] * m.zAxis[1] + m.yAxis[2] * m.zAxis[2]);
			// |hx X + hy Y + hz Z|^2 over the corner signs <= sum of squares + twice the absolute cross terms (exact for orthogonal axes)
			radius = sqrtf(hx * hx * xx + hy * hy * yy + hz * hz * zz + 2.0f * (hx * hy * xy + hx * hz * xz + hy * hz * yz));
		}
		int32_t submitted = 0;
		for (int s = 0; s < 6; s++) {
			const dfpsr_camera &fc = faceCameras[s];
			const dfpsr_transform3d &l = fc.location;
			const float dx = centre[0] - l.position[0], dy = centre[1] - l.position[1], dz = centre[2] - l.position[2];
			const float px = dx * l.xAxis[0] + dy * l.xAxis[1] + dz * l.xAxis[2], py = dx * l.yAxis[0] + dy * l.yAxis[1] + dz * l.yAxis[2], pz = dx * l.zAxis[0] + dy * l.zAxis[1] + dz * l.zAxis[2];
			const float reach = radius * faceStretch[s] * 1.01f + 1e-3f; // farthest a corner can lie from the centre in camera space, with slack for rounding
			bool outside = false;
			for (int q = 0; q < fc.cullPlaneCount && !outside; q++) {
				const float *pl = fc.cullPlanes[q];
				outside = ((pl[0] * px + pl[1] * py + pl[2] * pz) - pl[3]) > reach;
			}
			if (outside) { continue; }
			batch.models.push_back(&model.desc); batch.transforms.push_back(op.transform); batch.cameras.push_back(faceCameras[s]);
			batch.cubeOfTask.push_back(cube); batch.faceOfTask.push_back(s);
			submitted++;
		}
		return submitted;
	};
	bool blend = false;
	if (ensure_cubes(shadowLights)) { return 1; } // before any light is looked at: growing the pools drops every cached map
	for (const dfpsr_sprite_world_op &op : w->ops) {
		switch (op.op) {
		case DFPSR_SW_BLOCK_CLEAR: {
			if (flush()) { return 1; }
			const Block &b = w->blocks[(size_t)op.block];
			const int32_t quads = BLOCK_SIZE * BLOCK_SIZE / 4;
			DFPSR_LAUNCH(block_clear_kernel, (quads + 255) / 256, 256, 0, stream, (uint4 *)b.dDiffuse, (uint4 *)b.dNormal, (float4 *)b.dHeight, quads);
			break;
		}
		case DFPSR_SW_BLOCK_SPRITE: case DFPSR_SW_SPRITE: {
			const int32_t target = op.op == DFPSR_SW_SPRITE ? -1 : op.block;
			if (target == -1 && flush_copies()) { return 1; }
			if (target != drawBlock) { if (flush()) { return 1; } drawBlock = target; }
			const SpriteType &type = *g_spriteTypes[(size_t)op.typeIndex];
			dfpsr_sprite_draw d;
			d.sourceHeight = height_image(type, op.frame); d.sourceA = atlas_image(type, 0, op.frame); d.sourceB = atlas_image(type, 2, op.frame);
			d.left = op.left; d.top = op.top; d.heightOffset = op.heightOffset;
			draws.push_back(d);
			break;
		}
		case DFPSR_SW_BLOCK_MODEL: case DFPSR_SW_MODEL: {
			if (flush()) { return 1; }
			if (op.op == DFPSR_SW_MODEL && flush_copies()) { return 1; }
			dfpsr_image h, d, n;
			targets_of(op.op == DFPSR_SW_MODEL ? -1 : op.block, h, d, n);
			const ModelType &type = *g_modelTypes[(size_t)op.typeIndex];
			if (dense_render(type.dTriangles, (int32_t)type.triangles.size(), type.minBound, type.maxBound, view, h, d, n, op.worldOrigin, op.transform, false, nullptr, stream)) { return 1; }
			break;
		}
		case DFPSR_SW_COPY_BLOCK: {
			if (flush()) { return 1; }
			const Block &b = w->blocks[(size_t)op.block];
			CopyDev c;
			c.diffuse = b.dDiffuse; c.normal = b.dNormal; c.height = b.dHeight;
			c.left = op.left; c.top = op.top; c.width = op.width; c.height_ = op.height; c.sourceLeft = op.sourceLeft; c.sourceTop = op.sourceTop;
			copies.push_back(c);
			break;
		}
		case DFPSR_SW_LIGHT_CLEAR: break; // dfpsr_light_frame starts from black without directed lights
		case DFPSR_SW_LIGHT_DIRECTED: {
			const DirectedLightRec &l = w->directedLights[(size_t)op.light];
			dfpsr_directed_light d;
			memcpy(d.direction, l.direction, sizeof(d.direction)); d.intensity = l.intensity; memcpy(d.colorRgb, l.color, sizeof(d.colorRgb));
			directed.push_back(d);
			break;
		}
		case DFPSR_SW_SHADOW_CLEAR: {
			cubeOfLight[(size_t)op.light] = cubeCount++;
			signature.clear();
			passiveOps.clear();
			signature.insert(signature.end(), (const uint8_t *)&res, (const uint8_t *)&res + sizeof(res));
			signature.insert(signature.end(), (const uint8_t *)&view.normalToWorldSpace, (const uint8_t *)&view.normalToWorldSpace + sizeof(view.normalToWorldSpace));
			temporaryOfLight = 0;
			break;
		}
		case DFPSR_SW_SHADOW_SPRITE: case DFPSR_SW_SHADOW_MODEL: {
			if (op.flag == 0) { // passive caster: part of the light's signature, rendered only when the signature changes
				signature.insert(signature.end(), (const uint8_t *)&op.op, (const uint8_t *)&op.op + sizeof(op.op));
				signature.insert(signature.end(), (const uint8_t *)&op.typeIndex, (const uint8_t *)&op.typeIndex + sizeof(op.typeIndex));
				signature.insert(signature.end(), (const uint8_t *)&op.transform, (const uint8_t *)&op.transform + sizeof(op.transform));
				passiveOps.push_back(&op);
			} else {
				temporaryOfLight += submit_caster(op, temporary, cubeOfLight[(size_t)op.light]);
			}
			break;
		}
		case DFPSR_SW_LIGHT_POINT: {
			const PointLightRec &l = w->pointLights[(size_t)op.light];
			dfpsr_point_light &p = points[(size_t)op.light];
			memcpy(p.position, l.position, sizeof(p.position)); p.radius = l.radius; p.intensity = l.intensity; memcpy(p.colorRgb, l.color, sizeof(p.colorRgb));
			memset(&p.shadowCubeMap, 0, sizeof(p.shadowCubeMap));
			if (op.flag) {
				const int32_t cube = cubeOfLight[(size_t)op.light];
				if (w->staticSignature[(size_t)cube] != signature) { // the passive casters changed (or the slot is new): render the static map again
					for (const dfpsr_sprite_world_op *passive : passiveOps) { submit_caster(*passive, passiveBatch, cube); }
					passiveBatch.cubes.push_back(cube);
					w->staticSignature[(size_t)cube] = signature;
				}
				if (temporaryOfLight > 0) { temporary.cubes.push_back(cube); }
				usesWorking[(size_t)op.light] = temporaryOfLight > 0;
				p.shadowCubeMap.width = res; p.shadowCubeMap.height = res * 6; p.shadowCubeMap.stride = res * 4; p.shadowCubeMap.packOrder = DFPSR_PACK_RGBA;
			}
			break;
		}
		case DFPSR_SW_BLEND: blend = true; break;
		default: break;
		}
	}
	const bool timing = getenv("DFPSR_SW_TIMING") != nullptr; // developer aid: host-side phase times of one frame on stderr
	auto now = []() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; };
	const double t0 = timing ? now() : 0.0;
	if (flush() || flush_copies()) { return 1; }
	const double t1 = timing ? now() : 0.0;
	// ---- shadows: static maps whose passive casters changed are rendered again (cleared, one submission); lights with temporary casters
	// get a working copy of their static map with the temporary casters rendered on top (one copy per run of neighbouring maps, one
	// submission without clearing). A frame whose passive casters did not change and that has no temporary caster renders no shadow at all.
	auto run_batch = [&](ShadowBatch &batch, const DeviceBuffer &pool, int32_t clear) -> int {
		if (batch.cubes.empty()) { return 0; }
		std::vector<int32_t> position((size_t)w->cubeCapacity, -1);
		std::vector<dfpsr_image> faces;
		for (size_t k = 0; k < batch.cubes.size(); k++) {
			position[(size_t)batch.cubes[k]] = (int32_t)k;
			for (int f = 0; f < 6; f++) {
				dfpsr_image face;
				face.data = (uint8_t *)pool.ptr + cubeBytes * (size_t)batch.cubes[k] + (size_t)f * res * res * 4;
				face.width = res; face.height = res; face.stride = res * 4; face.packOrder = DFPSR_PACK_RGBA;
				faces.push_back(face);
			}
		}
		std::vector<int32_t> targets(batch.models.size());
		for (size_t t = 0; t < targets.size(); t++) { targets[t] = position[(size_t)batch.cubeOfTask[t]] * 6 + batch.faceOfTask[t]; }
		batch.models.reserve(1); batch.transforms.reserve(1); batch.cameras.reserve(1); targets.reserve(1); // non-null data() for a batch without casters
		return dfpsr_model_render_depth_batch(batch.models.data(), batch.transforms.data(), batch.cameras.data(), targets.data(), (int32_t)batch.models.size(),
		                                      faces.data(), (int32_t)faces.size(), clear, 0.0f, stream);
	};
	const size_t shadowTasks = passiveBatch.models.size() + temporary.models.size();
	if (run_batch(passiveBatch, w->cubeStatic, 1)) { return 1; }
	if (!temporary.cubes.empty()) {
		if (g_pendingFrames > 0 && verify_pending_frames()) { return 1; } // the copies below are not queued through DFPSR_LAUNCH
		for (size_t k = 0; k < temporary.cubes.size();) {
			size_t run = 1;
			while (k + run < temporary.cubes.size() && temporary.cubes[k + run] == temporary.cubes[k] + (int32_t)run) { run++; }
			const size_t offset = cubeBytes * (size_t)temporary.cubes[k];
			DFPSR_CHECK_CUDA(cudaMemcpyAsync((uint8_t *)w->cubeWorking.ptr + offset, (const uint8_t *)w->cubeStatic.ptr + offset, cubeBytes * run, cudaMemcpyDeviceToDevice, stream));
			k += run;
		}
		if (run_batch(temporary, w->cubeWorking, 0)) { return 1; }
	}
	for (size_t i = 0; i < points.size(); i++) {
		if (cubeOfLight[i] >= 0 && points[i].shadowCubeMap.width > 0) {
			points[i].shadowCubeMap.data = (uint8_t *)(usesWorking[i] ? w->cubeWorking.ptr : w->cubeStatic.ptr) + cubeBytes * (size_t)cubeOfLight[i];
		}
	}
	const int32_t worldCenter[2] = {find_world_center(w, width, height).x, find_world_center(w, width, height).y};
	dfpsr_ortho_view lightView;
	dfpsr_ortho_camera_light_view(&view, &lightView);
	const double t2 = timing ? now() : 0.0;
	const int status = dfpsr_light_frame(&lightView, worldCenter, blend ? colorTarget : nullptr, &fDiffuse, &fLight, &fNormal, &fHeight, directed.data(), (int32_t)directed.size(), points.data(), (int32_t)points.size(), stream);
	if (timing) { fprintf(stderr, "sprite world frame: sprites+copies %.0f us, shadow batches (%zu tasks, %zu static + %zu working maps) %.0f us, light frame launch %.0f us\n", t1 - t0, shadowTasks, passiveBatch.cubes.size(), temporary.cubes.size(), t2 - t1, now() - t2); }
	return status;
}

// ------------------------------------------------------------------------------------------------ C ABI

extern "C" {

int dfpsr_ortho_system_create(dfpsr_ortho_system *out, float cameraTilt, int32_t pixelsPerTile) { // ref: orthoAPI.cpp:82-119 OrthoSystem::update
	DFPSR_REQUIRE(out != nullptr, "ortho_system_create: null output");
	out->cameraTilt = cameraTilt; out->pixelsPerTile = pixelsPerTile;
	const int32_t yPixelsPerTile = (int32_t)((double)(float)pixelsPerTile / sqrt((double)(cameraTilt * cameraTilt + 1.0f)));
	const F3 up = f3(0.0f, 1.0f, 0.0f);
	static const int32_t worldDirections[8] = {7, 1, 3, 5, 0, 2, 4, 6};
	const float diag = 0.707106781f;
	const F3 forward[8] = {f3(diag, cameraTilt, diag), f3(-diag, cameraTilt, diag), f3(-diag, cameraTilt, -diag), f3(diag, cameraTilt, -diag),
	                       f3(0.0f, cameraTilt, 1.0f), f3(-1.0f, cameraTilt, 0.0f), f3(0.0f, cameraTilt, -1.0f), f3(1.0f, cameraTilt, 0.0f)};
	for (int a = 0; a < 8; a++) {
		const M3 cameraSystem = make_axis_system(forward[a], up);
		F3 normalDirection = cameraSystem.z;
		normalDirection.y = 0.0f;
		const M3 normalToWorld = make_axis_system(normalDirection, up);
		const float size = (float)pixelsPerTile, halfTile = (float)pixelsPerTile * 0.5f;
		const F2 xImage = ortho_world_to_image(cameraSystem, size, 0.5f, f3(1.0f, 0.0f, 0.0f)), zImage = ortho_world_to_image(cameraSystem, size, 0.5f, f3(0.0f, 0.0f, 1.0f));
		I2 xAxis, zAxis;
		xAxis.x = f2i(xImage.x - halfTile); xAxis.y = f2i(xImage.y - halfTile);
		zAxis.x = f2i(zImage.x - halfTile); zAxis.y = f2i(zImage.y - halfTile);
		make_view(out->view[a], a, xAxis, zAxis, yPixelsPerTile, normalToWorld, worldDirections[a]);
	}
	return 0;
}

int dfpsr_ortho_camera_light_view(const dfpsr_ortho_camera *camera, dfpsr_ortho_view *out) {
	DFPSR_REQUIRE(camera && out, "ortho_camera_light_view: null argument");
	out->normalToWorldSpace = camera->normalToWorldSpace;
	out->screenDepthToLightSpace = camera->screenDepthToLightSpace;
	out->lightSpaceToScreenDepth = camera->lightSpaceToScreenDepth;
	return 0;
}

int32_t dfpsr_dense_model_triangle_count(const dfpsr_polygon *polygons, int32_t polygonCount) { // ref: spriteAPI.cpp:1176-1186
	int32_t count = 0;
	for (int32_t i = 0; polygons && i < polygonCount; i++) { count += polygons[i].pointIndices[3] >= 0 ? 2 : 1; }
	return count;
}

int dfpsr_dense_model_build(const float *points, int32_t pointCount, const dfpsr_polygon *polygons, int32_t polygonCount, dfpsr_dense_triangle *out, float minBound[3], float maxBound[3]) {
	DFPSR_REQUIRE((points || pointCount == 0) && (polygons || polygonCount == 0) && minBound && maxBound, "dense_model_build: null argument");
	for (int k = 0; k < 3; k++) { minBound[k] = 0.0f; maxBound[k] = 0.0f; } // ref: Model.cpp:281-288
	for (int32_t i = 0; i < pointCount; i++) {
		for (int k = 0; k < 3; k++) { minBound[k] = std::min(minBound[k], points[i * 3 + k]); maxBound[k] = std::max(maxBound[k], points[i * 3 + k]); }
	}
	std::vector<F3> normals((size_t)pointCount, f3(0.0f, 0.0f, 0.0f));
	auto point = [&](int32_t index) { return f3(points + (size_t)index * 3); };
	for (int32_t i = 0; i < polygonCount; i++) { // ref: spriteAPI.cpp:1195-1205, :1158-1174 getAverageNormal
		const dfpsr_polygon &poly = polygons[i];
		const int32_t vertices = poly.pointIndices[3] >= 0 ? 4 : 3;
		for (int32_t v = 0; v < vertices; v++) { DFPSR_REQUIRE(poly.pointIndices[v] >= 0 && poly.pointIndices[v] < pointCount, "dense_model_build: polygon %d refers to point %d of %d", i, poly.pointIndices[v], pointCount); }
		F3 sum = f3(0.0f, 0.0f, 0.0f);
		for (int32_t t = 0; t < vertices - 2; t++) {
			const F3 a = point(poly.pointIndices[0]), b = point(poly.pointIndices[t + 1]), c = point(poly.pointIndices[t + 2]);
			sum = add(sum, normalize(cross(sub(b, a), sub(c, a))));
		}
		const F3 normal = normalize(sum);
		for (int32_t v = 0; v < vertices; v++) { F3 &n = normals[(size_t)poly.pointIndices[v]]; n = add(n, normal); }
	}
	for (F3 &n : normals) { n = normalize(n); }
	int32_t index = 0;
	for (int32_t i = 0; i < polygonCount; i++) { // ref: spriteAPI.cpp:1211-1232
		const dfpsr_polygon &poly = polygons[i];
		const int32_t vertices = poly.pointIndices[3] >= 0 ? 4 : 3;
		for (int32_t vb = 1; vb < vertices - 1; vb++) {
			const int32_t corner[3] = {0, vb, vb + 1};
			dfpsr_dense_triangle &t = out[index++];
			float *colors[3] = {t.colorA, t.colorB, t.colorC}, *positions[3] = {t.posA, t.posB, t.posC}, *outNormals[3] = {t.normalA, t.normalB, t.normalC};
			for (int c = 0; c < 3; c++) {
				const int32_t p = poly.pointIndices[corner[c]];
				for (int k = 0; k < 3; k++) { colors[c][k] = poly.colors[corner[c]][k] * 255.0f; positions[c][k] = points[(size_t)p * 3 + k]; }
				outNormals[c][0] = normals[(size_t)p].x; outNormals[c][1] = normals[(size_t)p].y; outNormals[c][2] = normals[(size_t)p].z;
			}
		}
	}
	return 0;
}

int dfpsr_dense_model_render(const dfpsr_dense_triangle *triangles, int32_t triangleCount, const float minBound[3], const float maxBound[3], const dfpsr_ortho_camera *view, const dfpsr_image *height, const dfpsr_image *diffuse, const dfpsr_image *normal, const float worldOrigin[2], const dfpsr_transform3d *modelToWorld, int32_t highQuality, int32_t dirtyRect[4], void *stream) {
	int n = 0;
	DFPSR_REQUIRE(cudaGetDeviceCount(&n) == cudaSuccess && n > 0, "no CUDA device available; dfpsr_b200 has no CPU fallback");
	DFPSR_REQUIRE(minBound && maxBound && view && height && diffuse && normal && worldOrigin && modelToWorld, "dense_model_render: null argument");
	DFPSR_REQUIRE(height->data && diffuse->data && normal->data, "dense_model_render: all three targets must exist");
	DFPSR_REQUIRE(height->width == diffuse->width && height->height == diffuse->height && height->width == normal->width && height->height == normal->height, "dense_model_render: targets differ in size");
	DFPSR_REQUIRE(triangles || triangleCount == 0, "dense_model_render: null triangles");
	return dense_render(triangles, triangleCount, minBound, maxBound, *view, *height, *diffuse, *normal, worldOrigin, *modelToWorld, highQuality != 0, dirtyRect, as_stream(stream));
}

int dfpsr_sprite_generate_from_model(const dfpsr_dense_triangle *triangles, int32_t triangleCount, const float minBound[3], const float maxBound[3], const dfpsr_ortho_system *ortho, int32_t cameraAngles, dfpsr_baked_sprite *out, void *stream) {
	int devices = 0;
	DFPSR_REQUIRE(cudaGetDeviceCount(&devices) == cudaSuccess && devices > 0, "no CUDA device available; dfpsr_b200 has no CPU fallback");
	DFPSR_REQUIRE((triangles || triangleCount == 0) && minBound && maxBound && ortho && out, "sprite_generate_from_model: null argument");
	DFPSR_REQUIRE(cameraAngles >= 1 && cameraAngles <= 8, "Need at least one camera angle to generate a sprite!");
	memset(out, 0, sizeof(*out));
	if (minBound[0] > maxBound[0]) { return 0; } // nothing visible (:1345-1348)
	cudaStream_t s = as_stream(stream);
	// ref: :1352-1358 — worst-case square image
	const float worstCaseDiameter = (std::max(maxBound[0], -minBound[0]) + std::max(maxBound[1], -minBound[1]) + std::max(maxBound[2], -minBound[2])) * 2;
	const int32_t size = f2i(worstCaseDiameter) * ortho->pixelsPerTile;
	const int32_t maxRes = (int32_t)(size + 1 - signed_modulo(size - 1, 2)) + 4; // roundUp(size, 2) + 4
	DFPSR_REQUIRE(maxRes > 0 && maxRes <= 16384, "sprite_generate_from_model: the model needs a %d pixel wide image", maxRes);
	const int32_t width = maxRes, height = maxRes;
	const size_t pixels = (size_t)width * height;
	DeviceBuffer depth, planes, dTriangles, dCrop;
	struct Release { DeviceBuffer *b[4]; ~Release() { for (DeviceBuffer *x : b) { x->release(); } } } release{{&depth, &planes, &dTriangles, &dCrop}};
	if (depth.reserve(pixels * 4) || planes.reserve(pixels * 4 * 3 * (size_t)cameraAngles) || dTriangles.reserve(std::max<size_t>((size_t)triangleCount * sizeof(dfpsr_dense_triangle), 16)) || dCrop.reserve(16)) { return 1; }
	uint32_t *color = (uint32_t *)planes.ptr, *heights = color + pixels * cameraAngles, *normals = heights + pixels * cameraAngles;
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(dTriangles.ptr, triangles, (size_t)triangleCount * sizeof(dfpsr_dense_triangle), cudaMemcpyHostToDevice, s));
	DFPSR_CHECK_CUDA(cudaMemsetAsync(planes.ptr, 0, pixels * 4 * 3 * (size_t)cameraAngles, s)); // new images are black and transparent
	const float heightScale = 255.0f / (maxBound[1] - minBound[1]);
	const float origin[2] = {(float)width * 0.5f, (float)height * 0.5f};
	dfpsr_transform3d identity;
	memset(&identity, 0, sizeof(identity));
	identity.xAxis[0] = identity.yAxis[1] = identity.zAxis[2] = 1.0f;
	for (int32_t a = 0; a < cameraAngles; a++) {
		dfpsr_image d, c, n;
		d.data = depth.ptr; c.data = color + pixels * a; n.data = normals + pixels * a;
		d.width = c.width = n.width = width; d.height = c.height = n.height = height; d.stride = c.stride = n.stride = width * 4; d.packOrder = c.packOrder = n.packOrder = DFPSR_PACK_RGBA;
		if (dfpsr_image_fill_f32(&d, -1000000000.0f, stream)) { return 1; }
		if (dense_render((const dfpsr_dense_triangle *)dTriangles.ptr, triangleCount, minBound, maxBound, ortho->view[a], d, c, n, origin, identity, true, nullptr, s)) { return 1; }
		DFPSR_LAUNCH(bake_height_kernel, (unsigned)((pixels + 255) / 256), 256, 0, s, (const float *)depth.ptr, (const uint32_t *)c.data, heights + pixels * a, (int32_t)pixels, minBound[1], heightScale);
	}
	const int32_t cropInit[4] = {width, height, 0, 0};
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(dCrop.ptr, cropInit, sizeof(cropInit), cudaMemcpyHostToDevice, s));
	DFPSR_LAUNCH(bake_crop_kernel, dim3((unsigned)((width + 31) / 32), (unsigned)((height + 7) / 8)), 256, 0, s, (const uint32_t *)color, width, height, cameraAngles, (int32_t *)dCrop.ptr);
	int32_t crop[4];
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(crop, dCrop.ptr, sizeof(crop), cudaMemcpyDeviceToHost, s));
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(s));
	if (crop[0] > crop[2]) { return 0; } // nothing drawn (:1402-1405)
	const int32_t croppedWidth = crop[2] + 1 - crop[0], croppedHeight = crop[3] + 1 - crop[1];
	void *atlas = nullptr;
	const size_t atlasStride = (size_t)croppedWidth * 3 * 4;
	DFPSR_CHECK_CUDA(cudaMalloc(&atlas, atlasStride * croppedHeight * cameraAngles));
	for (int32_t a = 0; a < cameraAngles; a++) { // ref: :1420-1425 — [colour | height | normal] per row of the atlas
		const uint32_t *sources[3] = {color + pixels * a, heights + pixels * a, normals + pixels * a};
		for (int column = 0; column < 3; column++) {
			DFPSR_CHECK_CUDA(cudaMemcpy2DAsync((uint8_t *)atlas + (size_t)a * croppedHeight * atlasStride + (size_t)column * croppedWidth * 4, atlasStride,
			                                   sources[column] + (size_t)crop[1] * width + crop[0], (size_t)width * 4, (size_t)croppedWidth * 4, (size_t)croppedHeight, cudaMemcpyDeviceToDevice, s));
		}
	}
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(s)); // the working images are released on return
	out->atlas.data = atlas; out->atlas.width = croppedWidth * 3; out->atlas.height = croppedHeight * cameraAngles; out->atlas.stride = (int32_t)atlasStride; out->atlas.packOrder = DFPSR_PACK_RGBA;
	out->centerX = width / 2 - crop[0]; out->centerY = height / 2 - crop[1]; out->frameRows = cameraAngles; out->propertyColumns = 3;
	for (int k = 0; k < 3; k++) { out->minBound[k] = minBound[k]; out->maxBound[k] = maxBound[k]; }
	return 0;
}

int dfpsr_sprite_type_create(const uint32_t *atlasHost, int32_t width, int32_t height, int32_t strideBytes, const dfpsr_sprite_config *config, int32_t *typeIndex) {
	DFPSR_REQUIRE(atlasHost && config && typeIndex, "sprite_type_create: null argument");
	DFPSR_REQUIRE(config->frameRows > 0 && config->propertyColumns >= 3, "sprite_type_create: the atlas needs at least one frame row and the colour, height and normal columns");
	DFPSR_REQUIRE(width > 0 && height > 0 && strideBytes >= width * 4, "sprite_type_create: bad atlas dimensions");
	DFPSR_REQUIRE(config->triangleIndexCount % 3 == 0, "sprite_type_create: TriangleIndices must hold multiples of three");
	std::unique_ptr<SpriteType> t(new (std::nothrow) SpriteType());
	DFPSR_REQUIRE(t != nullptr, "out of host memory");
	// ref: spriteAPI.cpp:204-215 — bounds in mini units, rounded outwards
	t->minBoundMini = i3((int32_t)floor((double)(config->minBound[0] * (float)MINI_UNITS_PER_TILE)), (int32_t)floor((double)(config->minBound[1] * (float)MINI_UNITS_PER_TILE)), (int32_t)floor((double)(config->minBound[2] * (float)MINI_UNITS_PER_TILE)));
	t->maxBoundMini = i3((int32_t)ceil((double)(config->maxBound[0] * (float)MINI_UNITS_PER_TILE)), (int32_t)ceil((double)(config->maxBound[1] * (float)MINI_UNITS_PER_TILE)), (int32_t)ceil((double)(config->maxBound[2] * (float)MINI_UNITS_PER_TILE)));
	t->centerX = config->centerX; t->centerY = config->centerY;
	t->frameWidth = width / config->propertyColumns; t->frameHeight = height / config->frameRows;
	t->frameCount = config->frameRows; t->propertyColumns = config->propertyColumns;
	t->atlasWidth = width; t->atlasHeight = height;
	t->heightScale = (config->maxBound[1] - config->minBound[1]) / 255.0f; t->heightOffset = config->minBound[1]; // ref: spriteAPI.cpp:158-159
	t->atlas.resize((size_t)width * height);
	for (int32_t y = 0; y < height; y++) { memcpy(&t->atlas[(size_t)y * width], (const uint8_t *)atlasHost + (size_t)y * strideBytes, (size_t)width * 4); }
	if (config->pointCount > 0) { // ref: spriteAPI.cpp:217-226
		DFPSR_REQUIRE(config->points != nullptr && (config->triangleIndices != nullptr || config->triangleIndexCount == 0), "sprite_type_create: null shadow geometry");
		t->shadow.points.assign(config->points, config->points + (size_t)config->pointCount * 3);
		for (int32_t i = 0; i + 2 < config->triangleIndexCount; i += 3) {
			dfpsr_polygon poly;
			memset(&poly, 0, sizeof(poly));
			for (int k = 0; k < 3; k++) {
				DFPSR_REQUIRE(config->triangleIndices[i + k] >= 0 && config->triangleIndices[i + k] < config->pointCount, "sprite_type_create: triangle index out of bound");
				poly.pointIndices[k] = config->triangleIndices[i + k];
			}
			poly.pointIndices[3] = -1;
			for (int v = 0; v < 4; v++) { for (int k = 0; k < 4; k++) { poly.colors[v][k] = 1.0f; } }
			t->shadow.polygons.push_back(poly);
		}
		t->shadow.set_bounds();
	}
	g_spriteTypes.push_back(std::move(t));
	*typeIndex = (int32_t)g_spriteTypes.size() - 1;
	return 0;
}

int32_t dfpsr_sprite_type_count(void) { return (int32_t)g_spriteTypes.size(); }

int dfpsr_model_type_create(const dfpsr_dense_triangle *triangles, int32_t triangleCount, const float minBound[3], const float maxBound[3], const struct dfpsr_host_model *shadowModel, int32_t *typeIndex) {
	DFPSR_REQUIRE((triangles || triangleCount == 0) && minBound && maxBound && typeIndex, "model_type_create: null argument");
	std::unique_ptr<ModelType> t(new (std::nothrow) ModelType());
	DFPSR_REQUIRE(t != nullptr, "out of host memory");
	if (triangleCount > 0) { t->triangles.assign(triangles, triangles + triangleCount); }
	for (int k = 0; k < 3; k++) { t->minBound[k] = minBound[k]; t->maxBound[k] = maxBound[k]; }
	if (shadowModel && shadowModel->pointCount > 0) {
		t->shadow.points.assign(shadowModel->points, shadowModel->points + (size_t)shadowModel->pointCount * 3);
		if (shadowModel->polygonCount > 0) { t->shadow.polygons.assign(shadowModel->polygons, shadowModel->polygons + shadowModel->polygonCount); }
		t->shadow.set_bounds();
	}
	g_modelTypes.push_back(std::move(t));
	*typeIndex = (int32_t)g_modelTypes.size() - 1;
	return 0;
}

int32_t dfpsr_model_type_count(void) { return (int32_t)g_modelTypes.size(); }

int dfpsr_sprite_world_create(dfpsr_sprite_world **out, const dfpsr_ortho_system *ortho, int32_t shadowResolution) {
	DFPSR_REQUIRE(out && ortho, "sprite_world_create: null argument");
	DFPSR_REQUIRE(shadowResolution > 0, "sprite_world_create: the shadow resolution must be positive");
	*out = new (std::nothrow) dfpsr_sprite_world(*ortho, shadowResolution);
	DFPSR_REQUIRE(*out != nullptr, "out of host memory");
	return 0;
}

int dfpsr_sprite_world_destroy(dfpsr_sprite_world *world) {
	if (!world) { return 0; }
	for (Block &b : world->blocks) { if (b.dDiffuse) { cudaFree(b.dDiffuse); cudaFree(b.dNormal); cudaFree(b.dHeight); } }
	world->cubeStatic.release(); world->cubeWorking.release();
	for (DeviceImage *im : {&world->diffuse, &world->normal, &world->light, &world->heightBuffer}) { if (im->ptr) { cudaFree(im->ptr); } }
	world->copyStaging.release();
	delete world;
	return 0;
}

#define WORLD_REQUIRE(world, name) DFPSR_REQUIRE((world) != nullptr, "The world handle was null in " name)

int dfpsr_sprite_world_add_background_sprite(dfpsr_sprite_world *world, const dfpsr_sprite_instance *sprite) {
	WORLD_REQUIRE(world, "spriteWorld_addBackgroundSprite");
	DFPSR_REQUIRE(sprite != nullptr, "sprite_world_add_background_sprite: null sprite");
	DFPSR_REQUIRE(sprite->typeIndex >= 0 && sprite->typeIndex < (int32_t)g_spriteTypes.size(), "Sprite type index %d is out of bound!", sprite->typeIndex);
	DFPSR_REQUIRE(sprite->direction >= 0 && sprite->direction < 8, "sprite direction %d is out of bound", sprite->direction);
	const SpriteType &type = *g_spriteTypes[(size_t)sprite->typeIndex];
	const I3 location = i3(sprite->location[0], sprite->location[1], sprite->location[2]);
	I3 worldMin, worldMax;
	get_3d_bounds(t3(f3((float)location.x, (float)location.y, (float)location.z), sprite_direction(sprite->direction)),
	              f3((float)type.minBoundMini.x, (float)type.minBoundMini.y, (float)type.minBoundMini.z), f3((float)type.maxBoundMini.x, (float)type.maxBoundMini.y, (float)type.maxBoundMini.z), worldMin, worldMax);
	if (world->passiveSprites.insert(*sprite, location, worldMin, worldMax)) { return 1; }
	const dfpsr_sprite_world_op op = sprite_op(DFPSR_SW_SPRITE, *sprite, world->view(), I2{0, 0});
	update_passive_region(world, Rect(op.left, op.top, op.width, op.height));
	return 0;
}

int dfpsr_sprite_world_add_background_model(dfpsr_sprite_world *world, const dfpsr_model_instance *model) {
	WORLD_REQUIRE(world, "spriteWorld_addBackgroundModel");
	DFPSR_REQUIRE(model != nullptr, "sprite_world_add_background_model: null model");
	DFPSR_REQUIRE(model->typeIndex >= 0 && model->typeIndex < (int32_t)g_modelTypes.size(), "Model type index %d is out of bound!", model->typeIndex);
	const ModelType &type = *g_modelTypes[(size_t)model->typeIndex];
	const T3 location = t3(model->location);
	const I3 origin = i3(floating_tile_to_mini(location.position.x), floating_tile_to_mini(location.position.y), floating_tile_to_mini(location.position.z));
	I3 worldMin, worldMax;
	const float unit = (float)MINI_UNITS_PER_TILE;
	get_3d_bounds(t3(scale(location.position, unit), location.m), scale(f3(type.minBound), unit), scale(f3(type.maxBound), unit), worldMin, worldMax);
	const Rect pixels = screen_bounds(world, worldMin, worldMax);
	if (world->passiveModels.insert(*model, origin, worldMin, worldMax)) { return 1; }
	update_passive_region(world, pixels);
	return 0;
}

int dfpsr_sprite_world_add_temporary_sprite(dfpsr_sprite_world *world, const dfpsr_sprite_instance *sprite) {
	WORLD_REQUIRE(world, "spriteWorld_addTemporarySprite");
	DFPSR_REQUIRE(sprite != nullptr, "sprite_world_add_temporary_sprite: null sprite");
	DFPSR_REQUIRE(sprite->typeIndex >= 0 && sprite->typeIndex < (int32_t)g_spriteTypes.size(), "Sprite type index %d is out of bound!", sprite->typeIndex);
	DFPSR_REQUIRE(sprite->direction >= 0 && sprite->direction < 8, "sprite direction %d is out of bound", sprite->direction);
	world->temporarySprites.push_back(*sprite);
	return 0;
}

int dfpsr_sprite_world_add_temporary_model(dfpsr_sprite_world *world, const dfpsr_model_instance *model) {
	WORLD_REQUIRE(world, "spriteWorld_addTemporaryModel");
	DFPSR_REQUIRE(model != nullptr, "sprite_world_add_temporary_model: null model");
	DFPSR_REQUIRE(model->typeIndex >= 0 && model->typeIndex < (int32_t)g_modelTypes.size(), "Model type index %d is out of bound!", model->typeIndex);
	world->temporaryModels.push_back(*model);
	return 0;
}

int dfpsr_sprite_world_remove_background_sprites(dfpsr_sprite_world *world, const int32_t searchMin[3], const int32_t searchMax[3], dfpsr_sprite_selection filter, void *user) {
	WORLD_REQUIRE(world, "spriteWorld_removeBackgroundSprites");
	DFPSR_REQUIRE(searchMin && searchMax, "sprite_world_remove_background_sprites: null bound");
	world->passiveSprites.map_box(i3(searchMin[0], searchMin[1], searchMin[2]), i3(searchMax[0], searchMax[1], searchMax[2]), [&](dfpsr_sprite_instance &sprite, I3 origin, I3 mn, I3 mx) {
		const int32_t o[3] = {origin.x, origin.y, origin.z}, a[3] = {mn.x, mn.y, mn.z}, b[3] = {mx.x, mx.y, mx.z};
		if (filter && !filter(&sprite, o, a, b, user)) { return false; }
		update_passive_region(world, screen_bounds(world, mn, mx));
		return true;
	});
	return 0;
}

int dfpsr_sprite_world_remove_background_models(dfpsr_sprite_world *world, const int32_t searchMin[3], const int32_t searchMax[3], dfpsr_model_selection filter, void *user) {
	WORLD_REQUIRE(world, "spriteWorld_removeBackgroundModels");
	DFPSR_REQUIRE(searchMin && searchMax, "sprite_world_remove_background_models: null bound");
	world->passiveModels.map_box(i3(searchMin[0], searchMin[1], searchMin[2]), i3(searchMax[0], searchMax[1], searchMax[2]), [&](dfpsr_model_instance &model, I3 origin, I3 mn, I3 mx) {
		const int32_t o[3] = {origin.x, origin.y, origin.z}, a[3] = {mn.x, mn.y, mn.z}, b[3] = {mx.x, mx.y, mx.z};
		if (filter && !filter(&model, o, a, b, user)) { return false; }
		update_passive_region(world, screen_bounds(world, mn, mx));
		return true;
	});
	return 0;
}

int dfpsr_sprite_world_create_temporary_point_light(dfpsr_sprite_world *world, const float position[3], float radius, float intensity, const int32_t colorRgb[3], int32_t shadowCasting) {
	WORLD_REQUIRE(world, "spriteWorld_createTemporary_pointLight");
	DFPSR_REQUIRE(position && colorRgb, "sprite_world_create_temporary_point_light: null argument");
	PointLightRec l;
	memcpy(l.position, position, sizeof(l.position)); l.radius = radius; l.intensity = intensity; memcpy(l.color, colorRgb, sizeof(l.color)); l.shadowCasting = shadowCasting ? 1 : 0;
	world->pointLights.push_back(l);
	return 0;
}

int dfpsr_sprite_world_create_temporary_directed_light(dfpsr_sprite_world *world, const float direction[3], float intensity, const int32_t colorRgb[3]) {
	WORLD_REQUIRE(world, "spriteWorld_createTemporary_directedLight");
	DFPSR_REQUIRE(direction && colorRgb, "sprite_world_create_temporary_directed_light: null argument");
	DirectedLightRec l;
	memcpy(l.direction, direction, sizeof(l.direction)); l.intensity = intensity; memcpy(l.color, colorRgb, sizeof(l.color));
	world->directedLights.push_back(l);
	return 0;
}

int dfpsr_sprite_world_clear_temporary(dfpsr_sprite_world *world) {
	WORLD_REQUIRE(world, "spriteWorld_clearTemporary");
	world->temporarySprites.clear(); world->temporaryModels.clear(); world->pointLights.clear(); world->directedLights.clear();
	return 0;
}

int dfpsr_sprite_world_get_camera_location(const dfpsr_sprite_world *world, int32_t location[3]) {
	WORLD_REQUIRE(world, "spriteWorld_getCameraLocation");
	location[0] = world->cameraLocation.x; location[1] = world->cameraLocation.y; location[2] = world->cameraLocation.z;
	return 0;
}

int dfpsr_sprite_world_set_camera_location(dfpsr_sprite_world *world, const int32_t location[3]) { // ref: spriteAPI.cpp:1055-1061
	WORLD_REQUIRE(world, "spriteWorld_setCameraLocation");
	if (world->cameraLocation.x != location[0] || world->cameraLocation.y != location[1] || world->cameraLocation.z != location[2]) {
		world->cameraLocation = i3(location[0], location[1], location[2]);
		world->dirty.all_dirty();
	}
	return 0;
}

// ref: orthoAPI.cpp:44-56 pixelToTileOffset / pixelToMiniOffset
static I3 pixel_to_mini_offset(const dfpsr_ortho_camera &v, int32_t px, int32_t py) {
	const float *m = v.roundedScreenPixelsToWorldTiles;
	const float x = (float)px, y = (float)py;
	const float tx = x * m[0] + y * m[2], tz = x * m[1] + y * m[3];
	return i3(floating_tile_to_mini(tx), 0, floating_tile_to_mini(tz));
}

int dfpsr_sprite_world_move_camera_in_pixels(dfpsr_sprite_world *world, int32_t offsetX, int32_t offsetY) { // ref: spriteAPI.cpp:1068-1074
	WORLD_REQUIRE(world, "spriteWorld_moveCameraInPixels");
	if (offsetX != 0 || offsetY != 0) {
		const I3 o = pixel_to_mini_offset(world->view(), offsetX, offsetY);
		world->cameraLocation = i3(world->cameraLocation.x + o.x, world->cameraLocation.y + o.y, world->cameraLocation.z + o.z);
		world->dirty.all_dirty();
	}
	return 0;
}

int dfpsr_sprite_world_get_camera_direction_index(const dfpsr_sprite_world *world, int32_t *index) {
	WORLD_REQUIRE(world, "spriteWorld_getCameraDirectionIndex");
	*index = world->cameraIndex;
	return 0;
}

int dfpsr_sprite_world_set_camera_direction_index(dfpsr_sprite_world *world, int32_t index) { // ref: spriteAPI.cpp:1103-1109
	WORLD_REQUIRE(world, "spriteWorld_setCameraDirectionIndex");
	DFPSR_REQUIRE(index >= 0 && index < 8, "camera direction index %d is out of bound", index);
	if (index != world->cameraIndex) { world->cameraIndex = index; world->dirty.all_dirty(); }
	return 0;
}

int dfpsr_sprite_world_find_ground_at_pixel(const dfpsr_sprite_world *world, int32_t targetWidth, int32_t targetHeight, int32_t pixelX, int32_t pixelY, int32_t location[3]) { // ref: spriteAPI.cpp:1050-1053
	WORLD_REQUIRE(world, "spriteWorld_findGroundAtPixel");
	const I2 center = find_world_center(world, targetWidth, targetHeight);
	const I3 r = pixel_to_mini_offset(world->view(), pixelX - center.x, pixelY - center.y);
	location[0] = r.x; location[1] = r.y; location[2] = r.z;
	return 0;
}

int dfpsr_sprite_world_plan_frame(dfpsr_sprite_world *world, int32_t width, int32_t height, const dfpsr_sprite_world_op **ops, int32_t *opCount) {
	WORLD_REQUIRE(world, "spriteWorld_draw");
	DFPSR_REQUIRE(width > 0 && height > 0 && ops && opCount, "sprite_world_plan_frame: bad argument");
	plan_frame(world, width, height);
	*ops = world->ops.data();
	*opCount = (int32_t)world->ops.size();
	return 0;
}

int dfpsr_sprite_world_draw(dfpsr_sprite_world *world, const dfpsr_image *colorTarget, void *stream) {
	WORLD_REQUIRE(world, "spriteWorld_draw");
	int n = 0;
	DFPSR_REQUIRE(cudaGetDeviceCount(&n) == cudaSuccess && n > 0, "no CUDA device available; dfpsr_b200 has no CPU fallback");
	DFPSR_REQUIRE(colorTarget && colorTarget->data && colorTarget->width > 0 && colorTarget->height > 0, "sprite_world_draw: the colour target does not exist");
	plan_frame(world, colorTarget->width, colorTarget->height);
	return execute_frame(world, colorTarget, as_stream(stream));
}

int dfpsr_sprite_world_draw_host(dfpsr_sprite_world *world, uint32_t *colorHost, int32_t strideBytes, int32_t width, int32_t height, int32_t packOrder, void *stream) {
	WORLD_REQUIRE(world, "spriteWorld_draw");
	DFPSR_REQUIRE(colorHost && width > 0 && height > 0 && strideBytes >= width * 4, "sprite_world_draw_host: bad colour image");
	static thread_local DeviceBuffer staging;
	const int32_t stride = ((width * 4 + 15) / 16) * 16;
	if (staging.reserve((size_t)stride * height)) { return 1; }
	dfpsr_image target; target.data = staging.ptr; target.width = width; target.height = height; target.stride = stride; target.packOrder = packOrder;
	if (dfpsr_sprite_world_draw(world, &target, stream)) { return 1; }
	DFPSR_CHECK_CUDA(cudaMemcpy2DAsync(colorHost, (size_t)strideBytes, staging.ptr, (size_t)stride, (size_t)width * 4, (size_t)height, cudaMemcpyDeviceToHost, as_stream(stream)));
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
	return 0;
}

int dfpsr_sprite_world_get_buffers(const dfpsr_sprite_world *world, dfpsr_image *diffuse, dfpsr_image *normal, dfpsr_image *light, dfpsr_image *height) {
	WORLD_REQUIRE(world, "spriteWorld_getDiffuseBuffer");
	if (diffuse) { *diffuse = world->diffuse.image(); }
	if (normal) { *normal = world->normal.image(); }
	if (light) { *light = world->light.image(); }
	if (height) { *height = world->heightBuffer.image(); }
	return 0;
}

} // extern "C"
