// raster.cu — the triangle pipeline on sm_100a: warp-owned 32x4 pixel tiles, batched views, one launch sequence per frame.
//
//   project_kernel       ref: api/modelAPI.cpp:238-242 + implementation/render/Camera.h:157-190 — all tasks of a batch in one launch; leaves
//                        the frustum plane codes of every point in PPoint::pad and clears the frame's counters in its spare grid rows
//   setup_kernel<false>  ref: implementation/render/renderCore.cpp:172-341 (cull, clip, back-face) — counting pass: commands and rows per
//                        slot, an upper bound of the entries of every screen tile (bounding box), units and checkpoints of large commands
//   counts_kernel        ordered prefix sums that replace List<TriangleDrawCommand>::push (renderCore.cpp:445: commands keep submission
//                        order), a segment of the entry pool for every tile (replaces CommandQueue::execute's 12 strips, :449-480, with
//                        per-tile lists) and the frame's totals / verdict for the host — one launch
//   setup_kernel<true>   emits compact draw commands + their row intervals (implementation/render/ITriangle2D.cpp:31-176) + interpolation
//                        planes (:182-300) and bins small commands to exactly the tiles their row intervals touch; large commands are
//                        queued as (command, tile row) units
//   big_units_kernel     one thread per unit: row intervals, binning and the interpolation checkpoints at the tile edges a row pair crosses
//   sort_lists_kernel    only when a tile holds more than 64 entries: restores submission order inside long tile lists
//   raster_kernel        ref: shader/fillerTemplates.h:108-441 + shader/RgbaMultiply.h:37-175 + api/textureAPI.h:253-495
//                        one WARP per tile, one lane per aligned 2x2 quad, colour and depth of the tile in registers for the whole list,
//                        no block-wide barrier anywhere. Commands are taken 16 at a time: lane (command, row pair) loads or replays the
//                        reference's running float sums up to the tile's left edge ONCE and leaves a checkpoint in shared memory; every
//                        quad lane then works through the commands that touch ITS quad in submission order, continuing from the
//                        checkpoint with at most 15 additions. Instances: deferred (visibility pass + every pixel shaded once; exact with
//                        prepared shading inputs, exact raw for frames without textures, tolerance), immediate (alpha-filtered frames),
//                        depth only.
//   wireframe_kernel     ref: api/rendererAPI.cpp:362-399 — renderer_end's debug overlay
//   occlude_existing_kernel / top_rows_kernel   ref: api/rendererAPI.cpp:242-258, :403-477 — the device side of the occlusion grid
//
// Exactness: coverage is the reference's int64 row-interval arithmetic; interpolated (1/W, U/W, V/W) replay the reference's chain of float
// additions from each row pair's outer block start, so colour and depth are bit-identical to the reference's scalar build (exact 1/x).
// Asynchronous frames, pools sized from history and the verdict protocol with the host: see run_frame / verify_frame below and DESIGN 4.7.
#include "common.cuh"

#include <algorithm>
#include <cstddef>
#include <vector>
#include <new>

namespace dfpsr {

#ifndef DFPSR_TILE_W
#define DFPSR_TILE_W 32 // pixels per tile row: 32 (a 32 x 4 tile, 16 x 2 quads) or 16 (16 x 8, 8 x 4 quads); one warp owns a tile of 128 pixels either way
#endif
static const int TILE_W = DFPSR_TILE_W, TILE_H = 128 / DFPSR_TILE_W;
static const int QUADS_X = TILE_W / 2, PAIRS_Y = TILE_H / 2; // lane = quad (qx, qy) = (lane % QUADS_X, lane / QUADS_X)
static const int BATCH = 32 / PAIRS_Y;      // commands whose checkpoints are prepared together (BATCH x PAIRS_Y row pairs = 32 lanes)
static_assert((TILE_W == 32 || TILE_W == 16) && QUADS_X * PAIRS_Y == 32, "tile shape");
static const int SMALL_ROWS = 8;           // large frames: triangles up to this many rows and SMALL_WIDTH columns are scan-converted by their set-up thread
static const int SMALL_WIDTH = 128;
static const int SMALL_TILES = 8;          // counting pass: bounding boxes up to this many tiles are counted by the set-up thread
#ifndef SETUP_THREADS_N
#define SETUP_THREADS_N 64
#endif
static const int SETUP_THREADS = SETUP_THREADS_N; // slots (quads / pre-projected triangles) per set-up CTA
#ifndef RASTER_WARPS_N
#define RASTER_WARPS_N 4
#endif
static const int RASTER_WARPS = RASTER_WARPS_N;
#ifndef RASTER_MIN_BLOCKS
#define RASTER_MIN_BLOCKS 8 // 64 registers: measured 50 us per 1080p terrain frame against 63 us at 128 registers (tools/variant_sweep.py)
#endif
#ifndef RASTER_MIN_BLOCKS_DEFERRED
#define RASTER_MIN_BLOCKS_DEFERRED 8
#endif
#ifndef RASTER_MIN_BLOCKS_TOLERANCE
#define RASTER_MIN_BLOCKS_TOLERANCE 8
#endif
static const int LOCAL_SORT = 64;           // tile lists up to this long are sorted by the tile kernel's own warp (two keys per lane)
static const int SORT_THREADS = 256;
static const int SORT_SMEM = 4096;         // entries of one tile list sorted in shared memory by a CTA; longer lists use the rank sort
static const int SORT_WARP = SORT_SMEM / (SORT_THREADS / 32); // lists up to this long are sorted by single warps

struct PPoint { // == dfpsr_projected_point
	float csx, csy, csz, isx, isy;
	int32_t pad;
	long long fx, fy;
};
static_assert(sizeof(PPoint) == 40, "PPoint layout");
static_assert(sizeof(dfpsr_projected_point) == 40, "dfpsr_projected_point layout");

// One draw command = one front-facing triangle after culling and clipping (ref: renderCore.h:52-70 carries 456 bytes).
struct Cmd {
	float start[3], dx[3], dy[3]; // Projection (ref: ITriangle2D.h:63-76)                        bytes   0..35
	uint32_t flags;               // CMD_* | diffuse index << 8 | light index << 20                     36..39
	uint32_t chkOffset;           // first stored checkpoint record, CHK_NONE when the tile kernel computes them  40..43
	uint32_t chkShape;            // first tile column | tile columns << 16 of the checkpoint table           44..47
	int32_t rowStart, rowCount;   // even-aligned rows (ref: ITriangle2D.cpp:70-75)                      48..55
	uint32_t rowOffset;           // first entry in the row-interval table (always even)                 56..59
	uint32_t pad_;
	float red[3], green[3], blue[3], alpha[3]; // scaled vertex colours (ref: RgbaMultiply.h:45-60)      64..111
	float u1[3], v1[3], u2[3], v2[3];          //                                                       112..159
};
static_assert(sizeof(Cmd) == 160, "Cmd layout");
static const uint32_t CHK_NONE = 0xFFFFFFFFu;

// Stored checkpoint of one (command, row pair, tile column): the reference's running sums where the row pair enters the tile.
// Written by the set-up warp that scan-converts a large triangle, read by the tile kernel (same meaning as Rec::mode / Rec::v there).
struct ChkRec {
	int32_t mode;
	float v[18];
	int32_t at; // column the sums stand at (the tile's left edge, or the row pair's outer block start inside the first tile)
};
static_assert(sizeof(ChkRec) == 80, "ChkRec layout: five 16-byte words, copied as one block into the tile kernel's shared-memory record");

enum : uint32_t {
	CMD_AFFINE = 1u, CMD_ALPHA = 2u, CMD_HAS_DIFFUSE = 4u, CMD_HAS_LIGHT = 8u, CMD_HAS_FADE = 16u, CMD_COLORLESS = 32u
};

// One submission (model or triangle batch) for one view. Lives in device memory; every set-up block copies its task to shared memory.
struct TaskParams {
	const float *points;
	const dfpsr_polygon *polygons;
	const dfpsr_triangle *triangles; // alternative source: pre-projected triangles
	PPoint *projected;
	int32_t pointCount, slotCount;
	int32_t slotBase, blockBase, blockCount;
	int32_t view;
	int32_t filter, diffuseIndex, lightIndex; // texture table indices or -1
	int32_t depthOnly;
	// dfpsr_renderer_give_tasks: the whole-model tests of model_render_threaded (api/modelAPI.cpp:228-234) run on the device, once per set-up CTA
	int32_t cullOnDevice;
	int32_t gridOffset;              // first float of this submission's snapshot of the occlusion grid in FrameDev::gridSnapshots, -1: no occluders
	float minBound[3], maxBound[3];
	dfpsr_transform3d modelToWorld;
	dfpsr_camera camera;
};

// One render target pair of the batch (96 bytes, read by the tile kernel with six 16-byte loads).
struct __align__(16) ViewDev {
	dfpsr_image color, depth; // data == nullptr when absent
	int32_t width, height;
	int32_t clipTop, clipBottom; // rows this process draws (strip mode; multiples of TILE_H or the image height)
	int32_t clear;               // targets are defined to be (clearColor, clearDepth) before this frame: no loads, every pixel stored
	uint32_t clearColor;
	float clearDepth;
	uint32_t tileBase;           // index of this view's first tile in the batch
	int32_t tilesX, tilesY;
	uint32_t packShifts, packSelector; // pack_shifts / pack_selector of the colour target's pack order, formed once on the host (44 instructions per tile otherwise)
};
static_assert(sizeof(ViewDev) == 96, "ViewDev layout");

// One texture of the frame's table (ref: implementation/image/Texture.h:42-95); Cmd::flags carries 12-bit indices into the table.
struct TexDev {
	const uint32_t *data;
	uint32_t log2width, log2height, maxMipLevel, startOffset, maxLevelMask;
	uint32_t pad_;
};
static_assert(sizeof(TexDev) == 32, "TexDev layout (two 16-byte loads)");
static const int MAX_TEXTURES = 4096; // 12 index bits per texture in Cmd::flags

struct FrameDev {
	const TaskParams *tasks;
	const ViewDev *views;
	const TexDev *textures;          // the frame's texture table, uploaded behind the task records
	int32_t checkpoints;             // 1: large commands leave interpolation checkpoints (exact mode); 0: tolerance mode, nothing to replay
	// Pools of the second half of the frame (elements). The host sizes them from earlier frames and does not wait for this frame's counts:
	// when checkCaps is set, the kernels behind the counting pass return at once if the frame does not fit (totals[TOTAL_OVERFLOW], set by
	// counts_kernel) and the host draws the frame again with larger pools when it next looks (renderer_end_internal / verify_frame).
	uint32_t capCmds, capRows, capEntries, capChk, capUnits, capSortTmp;
	int32_t checkCaps;
	int32_t sortScheduled;           // 0: the host did not launch sort_lists_kernel (no earlier frame had a tile list long enough to need it)
	uint32_t *hostSlot;              // mapped pinned memory: the frame's totals and verdict for the host (device alias)
	uint32_t serial;
	int32_t taskCount, viewCount, blockCount;
	uint32_t tileTotal;
	uint32_t *slotCounts;            // per slot: command count | rows << 3
	uint32_t *blockCmds, *blockRows; // per set-up block: totals, then exclusive offsets after scan_blocks_kernel
	uint32_t *tileCount, *tileOffset, *tileCursor;
	uint32_t *totals;                // [0] commands, [1] rows, [2] tile entries (upper bound), [3] max entries in one tile (upper bound),
	                                 // [4] checkpoint records (upper bound), [5] checkpoint cursor, [6] rank-sort scratch cursor,
	                                 // [7] (command, tile row) units of large commands, [8] large command cursor, [9] unit cursor,
	                                 // [10] the frame does not fit the pools, [11] finished CTAs of counts_kernel
	Cmd *cmds;
	int2 *rows;
	uint32_t *tileList;
	ChkRec *chk;                     // checkpoint records of large triangles; totals[4] = records needed (upper bound), totals[5] = cursor
	uint32_t *sortTmp;               // rank-sort scratch, as large as the entry pool; totals[6] = cursor
	struct BigItem *bigItems;        // large commands of the frame (totals[8] = cursor) and, per (command, tile row) unit, the index of its
	uint32_t *bigUnits;              // command in bigItems (totals[7] = units needed, counted by the first pass; totals[9] = cursor)
	const int32_t *blockTask;        // task of every set-up block (null: binary search over the tasks' block ranges)
	int32_t smallRows;               // triangles up to this many rows (and SMALL_WIDTH columns) are scan-converted by their set-up thread
	const float *occlusionGrid;      // 16-pixel cells of the farthest depth at which something can still be visible; null = no occluders
	int32_t gridWidth, gridHeight, gridStride;
	const float *gridSnapshots;      // the grid as it was at every dfpsr_renderer_give_tasks call that found occluders (gridStride x gridRows floats each)
	int32_t gridRows;
};

static const int TOTAL_OVERFLOW = 10, TOTAL_TICKET = 11, TOTAL_WORDS = 12;

// True when the frame's second half must not run: its counts exceed the pools it was launched with.
__device__ __forceinline__ bool frame_dropped(const FrameDev &frame) {
	return frame.checkCaps != 0 && frame.totals[TOTAL_OVERFLOW] != 0u;
}
// The verdict itself, from the totals of the finished counting pass (see counts_kernel / setup_kernel<true>).
__device__ __forceinline__ bool frame_fits(const FrameDev &frame) {
	const uint32_t *t = frame.totals;
	const uint32_t commands = t[0], rows = t[1], entries = t[2], maxTile = t[3], chk = t[4], units = t[7];
	return commands <= frame.capCmds && rows <= frame.capRows && entries <= frame.capEntries && chk <= frame.capChk && units <= frame.capUnits
	       && (maxTile <= (uint32_t)SORT_SMEM || entries <= frame.capSortTmp) && (maxTile <= (uint32_t)LOCAL_SORT || frame.sortScheduled != 0);
}
__device__ __forceinline__ void publish_totals(const FrameDev &frame, bool fits) {
	const uint32_t *t = frame.totals;
	volatile uint32_t *hostSlot = frame.hostSlot;
	hostSlot[0] = t[0]; hostSlot[1] = t[1]; hostSlot[2] = t[2]; hostSlot[3] = t[3]; hostSlot[4] = t[4]; hostSlot[7] = t[7];
	hostSlot[TOTAL_OVERFLOW] = fits ? 0u : 1u;
	hostSlot[TOTAL_TICKET] = frame.serial; // no fences: the host reads the slot after an event behind the publishing kernel, when its writes are visible
}

// ------------------------------------------------------------------------------------------------ projection

// The reference's int64_t(float) is x86 cvttss2si: out-of-range and NaN give INT64_MIN.
__host__ __device__ __forceinline__ long long float_to_i64(float v) {
	if (!(fabsf(v) < 9.2233720368547758e18f)) { return (long long)0x8000000000000000ull; }
#ifdef __CUDA_ARCH__
	return __float2ll_rz(v);
#else
	return (long long)v; // host copy for the occlusion grid (compiled with -ffp-contract=off)
#endif
}

// ref: implementation/render/Camera.h:160-187
__host__ __device__ __forceinline__ PPoint camera_to_screen(const dfpsr_camera &c, float x, float y, float z) {
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

// ref: math/FPlane3D.h:39-45
__device__ __forceinline__ bool plane_outside(const float *pl, const PPoint &p) {
	return !((((pl[0] * p.csx) + (pl[1] * p.csy) + (pl[2] * p.csz)) - pl[3]) <= 0.0f);
}

// One bit per frustum plane the point lies outside of: cull planes in bits 0-5, clip planes in bits 16-21. The projection pass leaves the
// codes in PPoint::pad, so that the set-up threads (three corners, two frustums, up to six planes each: a quarter of their instructions
// when every thread evaluated the planes itself) only combine bits.
__device__ __forceinline__ int32_t point_outcodes(const dfpsr_camera &c, const PPoint &p) {
	uint32_t codes = 0u;
	for (int s = 0; s < c.cullPlaneCount; s++) { if (plane_outside(c.cullPlanes[s], p)) { codes |= 1u << s; } }
	for (int s = 0; s < c.clipPlaneCount; s++) { if (plane_outside(c.clipPlanes[s], p)) { codes |= 0x10000u << s; } }
	return (int32_t)codes;
}

// One launch for every task of the batch: blockIdx.y = task. Grid rows behind the tasks clear the frame's tile counters and totals
// (`zeroWords` words at `zero`, 16-byte aligned), which saves the frame a separate memset.
__global__ void __launch_bounds__(256) project_kernel(const TaskParams *__restrict__ tasks, int32_t taskCount, uint4 *__restrict__ zero, uint32_t zeroWords) {
	chain_enter();
	if ((int32_t)blockIdx.y >= taskCount) {
		const uint32_t quads = (zeroWords + 3u) / 4u; // the buffer is padded to a multiple of 16 bytes
		const uint32_t i = (((uint32_t)blockIdx.y - (uint32_t)taskCount) * gridDim.x + blockIdx.x) * 256u + threadIdx.x;
		if (i < quads) { zero[i] = make_uint4(0u, 0u, 0u, 0u); }
		return;
	}
	__shared__ TaskParams task;
	for (uint32_t w = threadIdx.x; w < sizeof(TaskParams) / 4; w += blockDim.x) { ((uint32_t *)&task)[w] = ((const uint32_t *)&tasks[blockIdx.y])[w]; }
	__syncthreads();
	if (task.triangles != nullptr) { return; }
	for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < task.pointCount; i += gridDim.x * blockDim.x) {
		PPoint p = world_to_screen(task.camera, task.modelToWorld, task.points[3 * i], task.points[3 * i + 1], task.points[3 * i + 2]);
		p.pad = point_outcodes(task.camera, p);
		task.projected[i] = p;
	}
}

// dfpsr_project_points: the projection loop alone.
__global__ void __launch_bounds__(256) project_points_kernel(const float *__restrict__ points, int32_t count, dfpsr_transform3d m2w, dfpsr_camera camera, PPoint *__restrict__ out) {
	for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		out[i] = world_to_screen(camera, m2w, points[3 * i], points[3 * i + 1], points[3 * i + 2]);
	}
}

// ------------------------------------------------------------------------------------------------ set-up helpers

// ref: implementation/render/renderCore.cpp:172-198 on the corners' codes. 0 hidden, 1 full, 2 partial.
__device__ __forceinline__ int triangle_visibility_codes(const PPoint *p, bool clipFrustum) {
	const uint32_t shift = clipFrustum ? 16u : 0u;
	const uint32_t o0 = ((uint32_t)p[0].pad >> shift) & 0xFFFFu, o1 = ((uint32_t)p[1].pad >> shift) & 0xFFFFu, o2 = ((uint32_t)p[2].pad >> shift) & 0xFFFFu;
	if ((o0 & o1 & o2) != 0u) { return 0; }
	return (o0 | o1 | o2) != 0u ? 2 : 1;
}

// ref: implementation/render/ITriangle2D.cpp:55-60
__device__ __forceinline__ bool is_frontfacing(const PPoint *p) {
	return ((p[2].fx - p[0].fx) * (p[1].fy - p[0].fy)) + ((p[2].fy - p[0].fy) * (p[0].fx - p[1].fx)) < 0;
}

struct Bound { int32_t l, t, r, b; bool any; int32_t wl, wt, wr, wb; /* whole (unclipped) bound, ref: ITriangle2D.h wholeBound */ };

// ref: implementation/render/ITriangle2D.cpp:31-43, :62-75 — pixel bound, cut to the clip rectangle (0, clipTop, width, clipBottom - clipTop)
// exactly like executeTriangleDrawing's clipBound (renderCore.cpp:203-217), rows aligned to 2.
__device__ Bound raster_bound(const PPoint *p, int32_t width, int32_t clipTop, int32_t clipBottom) {
	int32_t rx0 = (int32_t)((p[0].fx + 128) / 256), ry0 = (int32_t)((p[0].fy + 128) / 256);
	int32_t rx1 = (int32_t)((p[1].fx + 128) / 256), ry1 = (int32_t)((p[1].fy + 128) / 256);
	int32_t rx2 = (int32_t)((p[2].fx + 128) / 256), ry2 = (int32_t)((p[2].fy + 128) / 256);
	int32_t l = min(rx0, min(rx1, rx2)) - 1, t = min(ry0, min(ry1, ry2)) - 1;
	int32_t r = max(rx0, max(rx1, rx2)) + 1, b = max(ry0, max(ry1, ry2)) + 1;
	Bound out;
	out.wl = l; out.wt = t; out.wr = r; out.wb = b;
	out.any = l < width && r > 0 && t < clipBottom && b > clipTop; // IRect::overlaps (math/IRect.h:77)
	if (!out.any) { out.l = out.t = out.r = out.b = 0; return out; }
	out.l = max(l, 0); out.r = min(r, width);
	int32_t top = max(t, clipTop), bottom = min(b, clipBottom);
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

// Truncating int64 division (C++ semantics), b != 0. The emulated 64-bit divide costs over a hundred instructions and was half of the
// set-up kernel for small triangles. When both operands are below 2^52 the double quotient is within one of the true quotient
// (relative error 2^-53, |a / b| < 2^52), so one exact remainder check repairs it; anything larger takes the generic divide.
__device__ __forceinline__ long long div_trunc(long long a, long long b) {
	const unsigned long long magA = (unsigned long long)(a < 0 ? -a : a), magB = (unsigned long long)(b < 0 ? -b : b);
	if (magA < (1ull << 52) && magB < (1ull << 52)) {
		long long q = __double2ll_rz((double)a / (double)b);
		long long r = a - q * b; // exact: |q * b| <= |a| + |b|
		const long long toward = ((a < 0) == (b < 0)) ? 1 : -1; // direction in which the quotient grows in magnitude
		if (r != 0 && ((r < 0) != (a < 0))) { q -= toward; r += toward * b; }        // overshot: the remainder must carry the sign of a
		else if ((unsigned long long)(r < 0 ? -r : r) >= magB) { q += toward; }        // undershot by one
		return q;
	}
	return a / b;
}

__device__ int2 edges_row(const EdgeSet &e, int32_t y) {
	int32_t left = e.leftBound, right = e.rightBound;
	if (e.degenerate) { return make_int2(e.rightBound, e.leftBound); }
	long long dy = (long long)(y - e.topBound);
#pragma unroll
	for (int i = 0; i < 3; i++) {
		if (e.kind[i] == 1) {
			long long limit = e.limit0[i] - e.offsetY[i] * dy;
			int32_t side = min(max(e.leftBound, (int32_t)(div_trunc(limit + 1, e.offsetX[i]) + 1)), e.rightBound);
			left = max(left, side);
		} else if (e.kind[i] == 2) {
			long long limit = e.limit0[i] - e.offsetY[i] * dy;
			int32_t side = min(max(e.leftBound, (int32_t)(div_trunc(limit, e.offsetX[i]) + 1)), e.rightBound);
			right = min(right, side);
		} else if (e.kind[i] == 3) {
			long long valueRow = e.valueOrigin[i] + e.offsetY[i] * dy;
			if (valueRow > (long long)e.threshold[i]) { left = e.rightBound; right = e.leftBound; }
		}
	}
	return make_int2(left, right);
}

// The same row interval for NARROW bounding boxes (at most NARROW_WIDTH columns) whose left end is not the screen's first column, without
// any division: every edge function is linear in x, so the first column it admits (left cuts) or rejects (right cuts) inside [l, r) is
// the number of columns in front of it, and those are counted by evaluating the int64 edge function at the box's few columns. This is the
// exact crossing the division computes — ceil(limit / offsetX) for left cuts, floor(limit / offsetX) + 1 for right cuts, clamped to
// [l, r] — as long as a negative quotient cannot matter: the reference's division truncates toward zero, which differs from the floor
// only for crossings left of pixel 0, and those clamp to l either way when l >= 1 (edge predicate == row intervals).
// A 2 M-triangle frame spent 45 % of its emit pass in the emulated 64-bit divisions of triangles that are three pixels wide.
static const int NARROW_WIDTH = 8;
__device__ __forceinline__ int2 edges_row_narrow(const EdgeSet &e, int32_t y) {
	int32_t left = e.leftBound, right = e.rightBound;
	if (e.degenerate) { return make_int2(e.rightBound, e.leftBound); }
	const long long dy = (long long)(y - e.topBound);
	const int32_t width = e.rightBound - e.leftBound;
#pragma unroll
	for (int i = 0; i < 3; i++) {
		if (e.kind[i] == 3) {
			const long long valueRow = e.valueOrigin[i] + e.offsetY[i] * dy;
			if (valueRow > (long long)e.threshold[i]) { left = e.rightBound; right = e.leftBound; }
		} else if (e.kind[i] != 0) {
			long long value = e.valueOrigin[i] + e.offsetY[i] * dy; // at column l
			const long long threshold = (long long)e.threshold[i];
			int32_t columns = 0; // left cut: columns in front of the first covered one; right cut: covered columns from l on
#pragma unroll
			for (int k = 0; k < NARROW_WIDTH; k++) {
				const bool covered = value <= threshold;
				columns += (k < width && covered == (e.kind[i] == 2)) ? 1 : 0;
				value += e.offsetX[i];
			}
			if (e.kind[i] == 1) { left = max(left, e.leftBound + columns); } else { right = min(right, e.leftBound + columns); }
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
	if (triangle_visibility_codes(p, false) == 0) { return; } // the corners carry their plane codes (project_kernel / load_triangle)
	if (!task.depthOnly && task.filter == DFPSR_FILTER_ALPHA && almost_zero(alpha[0]) && almost_zero(alpha[1]) && almost_zero(alpha[2])) { return; }
	if (triangle_visibility_codes(p, true) == 1) {
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
	// texture coordinates are 64 of a polygon's 144 bytes (and 48 of a command's 160): a task without textures neither reads nor writes them
	const bool textured = task.diffuseIndex >= 0 || task.lightIndex >= 0;
	if (task.triangles != nullptr) {
		const dfpsr_triangle &t = task.triangles[local];
#pragma unroll
		for (int k = 0; k < 3; k++) {
			p[k] = *(const PPoint *)&t.pos[k];
			p[k].pad = point_outcodes(task.camera, p[k]); // pre-projected corners come from the host without codes
#pragma unroll
			for (int ch = 0; ch < 4; ch++) { colors[k][ch] = t.colors[k][ch]; tex[k][ch] = textured ? t.texCoords[k][ch] : 0.0f; }
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
		for (int ch = 0; ch < 4; ch++) { colors[k][ch] = poly.colors[corner[k]][ch]; tex[k][ch] = textured ? poly.texCoords[corner[k]][ch] : 0.0f; }
	}
	return true;
}


// ------------------------------------------------------------------------------------------------ occlusion grid (ref: api/rendererAPI.cpp:36-351, :403-477)

static const int CELL_SIZE = 16; // ref: api/rendererAPI.cpp:36

struct CellBound { int32_t x0, y0, x1, y1; };
// ref: api/rendererAPI.cpp:169-180 getOuterCellBound (C++ division truncates toward zero, also for negative pixel coordinates)
__host__ __device__ __forceinline__ CellBound outer_cell_bound(int32_t left, int32_t top, int32_t right, int32_t bottom, int32_t gridWidth, int32_t gridHeight) {
	CellBound c;
	c.x0 = left / CELL_SIZE; c.x1 = right / CELL_SIZE + 1; c.y0 = top / CELL_SIZE; c.y1 = bottom / CELL_SIZE + 1;
	if (c.x0 < 0) { c.x0 = 0; } if (c.y0 < 0) { c.y0 = 0; }
	if (c.x1 > gridWidth) { c.x1 = gridWidth; } if (c.y1 > gridHeight) { c.y1 = gridHeight; }
	return c;
}
// ref: api/rendererAPI.cpp:99-135 pointInsideOfHull on the four corners of a cell, in 1/256 pixel units
__host__ __device__ __forceinline__ bool point_inside_of_hull(const PPoint *hull, int count, long long x, long long y) {
	for (int c = 0; c < count; c++) {
		int nc = c + 1 == count ? 0 : c + 1;
		long long dirX = hull[nc].fy - hull[c].fy, dirY = hull[c].fx - hull[nc].fx;
		if (!((dirX * (x - hull[c].fx)) + (dirY * (y - hull[c].fy)) <= 0)) { return false; }
	}
	return true;
}
__host__ __device__ __forceinline__ bool cell_inside_of_hull(const PPoint *hull, int count, int32_t cellX, int32_t cellY) {
	long long l = (long long)cellX * CELL_SIZE * 256, t = (long long)cellY * CELL_SIZE * 256, r = l + CELL_SIZE * 256, b = t + CELL_SIZE * 256;
	return point_inside_of_hull(hull, count, l, t) && point_inside_of_hull(hull, count, r, t) && point_inside_of_hull(hull, count, l, b) && point_inside_of_hull(hull, count, r, b);
}

// ref: api/rendererAPI.cpp:193-217 completeOcclusion for one command: nothing of it can be visible in any cell its whole bound touches
__device__ __forceinline__ bool command_occluded(const FrameDev &frame, const Bound &bound, const PPoint *q) {
	if (frame.occlusionGrid == nullptr) { return false; }
	const CellBound cells = outer_cell_bound(bound.wl, bound.wt, bound.wr, bound.wb, frame.gridWidth, frame.gridHeight);
	const float triangleDepth = fminf(q[0].csz, fminf(q[1].csz, q[2].csz));
	for (int32_t cy = cells.y0; cy < cells.y1; cy++) {
		for (int32_t cx = cells.x0; cx < cells.x1; cx++) {
			if ((double)triangleDepth < (double)frame.occlusionGrid[cy * frame.gridStride + cx] + 0.001) { return false; }
		}
	}
	return true;
}

// ref: api/rendererAPI.cpp:403-477 occludeFromTopRows: per cell the extreme of the FIRST pixel row of its cell row (the reference scans
// columns [16 k - 1, 16 k + 15) for cell k > 0 and [0, 15) for cell 0). out[cell] = the distance the host merges into the grid.
__global__ void __launch_bounds__(256) top_rows_kernel(dfpsr_image depth, int32_t width, int32_t gridWidth, int32_t gridHeight, int32_t perspective, float *__restrict__ out) {
	int32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
	if (cell >= gridWidth * gridHeight) { return; }
	int32_t gx = cell % gridWidth, gy = cell / gridWidth;
	int32_t x0 = gx == 0 ? 0 : gx * CELL_SIZE - 1, x1 = gx * CELL_SIZE + CELL_SIZE - 1;
	if (x1 >= width) { x1 = width; }
	if (x0 > x1) { x0 = x1; } // cells after the first capped one see an empty range
	const float *row = row_ptr<float>(depth.data, depth.stride, gy * CELL_SIZE);
	float extreme = perspective ? INFINITY : 0.0f;
	for (int32_t x = x0; x < x1; x++) {
		float v = row[x];
		if (perspective) { if (v < extreme) { extreme = v; } } else { if (v > extreme) { extreme = v; } }
	}
	out[cell] = perspective ? 1.0f / extreme : extreme;
}

// ------------------------------------------------------------------------------------------------ broad phase on the device
// ref: api/modelAPI.cpp:228-234 — a model is skipped when its bounding box lies outside of the cull frustum (Camera::isBoxSeen,
// implementation/render/Camera.h:202-217) or behind the occluders given so far (renderer_isBoxVisible, api/rendererAPI.cpp:302-351, :538-543).
// The same IEEE operations in the same order as the host versions (dfpsr_camera_is_box_seen in host_math.cpp, box_visible below), so a
// frame submitted through dfpsr_renderer_give_tasks makes exactly the decisions a loop of dfpsr_renderer_give_task makes. One thread per CTA.
__device__ bool task_is_culled(const FrameDev &frame, const TaskParams &task) {
	const dfpsr_camera &c = task.camera;
	const dfpsr_transform3d &m = task.modelToWorld, &l = c.location;
	float cs[8][3];
#pragma unroll
	for (int p = 0; p < 8; p++) { // corner p takes the maximum on axis k when bit k of p is set (x: bit 0, y: bit 1, z: bit 2)
		const float px = (p & 1) ? task.maxBound[0] : task.minBound[0], py = (p & 2) ? task.maxBound[1] : task.minBound[1], pz = (p & 4) ? task.maxBound[2] : task.minBound[2];
		const float wx = (px * m.xAxis[0] + py * m.yAxis[0] + pz * m.zAxis[0]) + m.position[0];
		const float wy = (px * m.xAxis[1] + py * m.yAxis[1] + pz * m.zAxis[1]) + m.position[1];
		const float wz = (px * m.xAxis[2] + py * m.yAxis[2] + pz * m.zAxis[2]) + m.position[2];
		const float dx = wx - l.position[0], dy = wy - l.position[1], dz = wz - l.position[2];
		cs[p][0] = dx * l.xAxis[0] + dy * l.xAxis[1] + dz * l.xAxis[2];
		cs[p][1] = dx * l.yAxis[0] + dy * l.yAxis[1] + dz * l.yAxis[2];
		cs[p][2] = dx * l.zAxis[0] + dy * l.zAxis[1] + dz * l.zAxis[2];
	}
	for (int s = 0; s < c.cullPlaneCount; s++) { // isBoxSeen == 0: some plane has every corner outside (a NaN distance counts as outside)
		const float *pl = c.cullPlanes[s];
		bool anyInside = false;
#pragma unroll
		for (int p = 0; p < 8; p++) { anyInside = anyInside || ((((pl[0] * cs[p][0]) + (pl[1] * cs[p][1]) + (pl[2] * cs[p][2])) - pl[3]) <= 0.0f); }
		if (!anyInside) { return true; }
	}
	if (task.gridOffset < 0) { return false; }
	// renderer_isBoxVisible against the snapshot of the grid: visible when the box's closest corner is nearer than some cell it may touch
	const ViewDev &view = frame.views[task.view];
	const float *grid = frame.gridSnapshots + task.gridOffset;
	int32_t left = 0, top = 0, right = 0, bottom = 0;
	float closest = INFINITY;
#pragma unroll
	for (int p = 0; p < 8; p++) {
		const PPoint q = camera_to_screen(c, cs[p][0], cs[p][1], cs[p][2]);
		const int32_t x = (int32_t)(q.fx / 256), y = (int32_t)(q.fy / 256); // ref: api/rendererAPI.cpp:89-95 getPixelBoundFromProjection
		if (p == 0) { left = x; top = y; right = x + 1; bottom = y + 1; }
		else { left = min(left, x); top = min(top, y); right = max(right, x + 1); bottom = max(bottom, y + 1); }
		if (q.csz < closest) { closest = q.csz; }
	}
	const int32_t gridWidth = (view.width + (CELL_SIZE - 1)) / CELL_SIZE, gridHeight = (view.height + (CELL_SIZE - 1)) / CELL_SIZE;
	const CellBound cells = outer_cell_bound(left, top, right, bottom, gridWidth, gridHeight);
	for (int32_t cy = cells.y0; cy < cells.y1; cy++) {
		for (int32_t cx = cells.x0; cx < cells.x1; cx++) {
			// image_readPixel_clamp on the grid as allocated; a grid that does not exist reads 0
			float cell = 0.0f;
			if (frame.gridStride > 0 && frame.gridRows > 0) { cell = grid[(size_t)min(max(cy, 0), frame.gridRows - 1) * frame.gridStride + min(max(cx, 0), frame.gridStride - 1)]; }
			if (closest < cell) { return false; }
		}
	}
	return true;
}

// ------------------------------------------------------------------------------------------------ set-up kernel

// A command whose tile counting (counting pass) or rows + tile entries (emit pass) are produced cooperatively by one warp.
struct BigItem {
	uint32_t cmdIndex, rowOffset, tileBase, chkOffset;
	uint32_t unitStart; // first (command, tile row) unit of this command in FrameDev::bigUnits
	int32_t tilesX, l, t, r, rowCount;
	int32_t tx0, tx1, ty0, ty1;
	long long fx[3], fy[3];
};

// Finds the task a set-up block belongs to (tasks own contiguous block ranges).
__device__ __forceinline__ int32_t task_of_block(const TaskParams *tasks, int32_t taskCount, int32_t block) {
	int32_t lo = 0, hi = taskCount - 1;
	while (lo < hi) {
		int32_t mid = (lo + hi + 1) >> 1;
		if (tasks[mid].blockBase <= block) { lo = mid; } else { hi = mid - 1; }
	}
	return lo;
}

// Appends `index` to the lists of the tiles of one tile row that pixels [minL, maxR) touch. Four tiles at a time: the atomics that hand out
// the list positions return values, and a loop of one atomic + dependent store per tile is a chain of L2 round trips (a row that crosses
// the whole 1080p target touches 60 tiles; that chain was the longest thread of a single frame's set-up).
__device__ __forceinline__ void emit_tile_row(const FrameDev &frame, uint32_t tileBase, int32_t tilesX, int32_t ty, int32_t minL, int32_t maxR, uint32_t index) {
	if (maxR <= minL) { return; }
	const int32_t first = minL / TILE_W, last = (maxR - 1) / TILE_W;
	if (first == last) { // the common case of small triangles: one tile
		const uint32_t tile = tileBase + (uint32_t)(ty * tilesX + first);
		const uint32_t position = atomicAdd(&frame.tileCursor[tile], 1u);
		frame.tileList[frame.tileOffset[tile] + position] = index;
		return;
	}
	for (int32_t tx = first; tx <= last; tx += 4) {
		const int32_t n = min(4, last - tx + 1);
		uint32_t position[4], offset[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			if (k < n) {
				const uint32_t tile = tileBase + (uint32_t)(ty * tilesX + tx + k);
				position[k] = atomicAdd(&frame.tileCursor[tile], 1u);
				offset[k] = frame.tileOffset[tile];
			}
		}
#pragma unroll
		for (int k = 0; k < 4; k++) { if (k < n) { frame.tileList[offset[k] + position[k]] = index; } }
	}
}

__device__ __forceinline__ void chk_store(ChkRec *dst, int32_t mode, const float *v, int count, int32_t at) {
	// word 0 = mode, words 1..18 = v, word 19 = column; assembled from registers (no local copy of the record)
	uint32_t w[20];
	w[0] = (uint32_t)mode; w[19] = (uint32_t)at;
#pragma unroll
	for (int i = 0; i < 18; i++) { w[1 + i] = i < count ? __float_as_uint(v[i]) : 0u; }
	uint4 *d = (uint4 *)dst;
#pragma unroll
	for (int q = 0; q < 5; q++) { d[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]); }
}

// Walks one row pair of a large triangle like the reference's fillShapeSuper does (shader/fillerTemplates.h:329-372: left-edge quads, the
// unclipped inner run whose four lanes advance separately, the closing multiplication, right-edge quads) and leaves a checkpoint of the
// running sums at the row pair's first quad and at every tile column boundary. recs is indexed by tile column - firstColumn.
// Kept out of line on purpose: inlined into big_units_kernel, nvcc 12.9 generated code for the tile-wise inner loop below that lost tile
// entries of some row pairs (the parity tests fail; the same source is bit-identical to the per-step loop on the host and out of line).
__device__ __noinline__ void chk_walk_row_pair(const float *start, const float *dx, const float *dy, int2 upperRow, int2 lowerRow, int32_t y1, ChkRec *recs, int32_t firstColumn) {
	const int32_t outerStart = min(upperRow.x, lowerRow.x), outerEnd = max(upperRow.y, lowerRow.y);
	const int32_t innerStart = max(upperRow.x, lowerRow.x), innerEnd = min(upperRow.y, lowerRow.y);
	const int32_t obs = outerStart & ~1, obe = (outerEnd + 1) & ~1, ibs = (innerStart + 1) & ~1, ibe = innerEnd & ~1;
	if (obe <= obs) { return; }
	float v[18], dx2[3];
	const float fx = (float)obs + 0.5f, fy = (float)y1 + 0.5f;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		v[k] = (start[k] + (dx[k] * fx)) + (dy[k] * fy);
		v[3 + k] = v[k] + dy[k];
		dx2[k] = dx[k] * 2.0f;
	}
	const bool noInner = ibe <= ibs;
	const int32_t leftEnd = noInner ? obe : ibs;
	int32_t x = obs;
	// No record for the tile column the row pair starts in: the tile kernel evaluates the planes at the outer block start itself (exactly
	// the values above), which is cheaper than storing and fetching 80 bytes — 60 % of all records of a terrain frame were of this kind.
	// The walk therefore ends at the last tile edge inside the row pair; a row pair within one tile column is not walked at all.
	const int32_t lastEdge = ((obe - 1) / TILE_W) * TILE_W; // obe > obs >= 0
	if (lastEdge <= obs) { return; }
	while (x < leftEnd) {
#pragma unroll
		for (int k = 0; k < 6; k++) { v[k] += dx2[k % 3]; }
		x += 2;
		if ((x & (TILE_W - 1)) == 0 && x < obe && (noInner || x <= ibs)) { chk_store(recs + (x / TILE_W - firstColumn), 0, v, 6, x); }
		if (x >= lastEdge) { return; }
	}
	if (noInner) { return; }
	// inner run: v[0..11] = lanes[k][l], v[12..17] = the sums after the run
	{
		const float quadCount = (float)((ibe - ibs) / 2);
		float up[3] = {v[0], v[1], v[2]}, lo[3] = {v[3], v[4], v[5]};
#pragma unroll
		for (int k = 0; k < 3; k++) {
			v[k * 4 + 0] = up[k]; v[k * 4 + 1] = up[k] + dx[k]; v[k * 4 + 2] = lo[k]; v[k * 4 + 3] = lo[k] + dx[k];
			v[12 + k] = up[k] + (dx2[k] * quadCount);
			v[15 + k] = lo[k] + (dx2[k] * quadCount);
		}
	}
	// The inner run carries most of a wide row pair (a triangle across a 1080p target: 900 steps of twelve dependent additions, the
	// longest thread of a single frame's unit kernel). It is walked tile by tile: up to the next tile edge, then whole tiles of sixteen
	// steps without any per-step test, with the checkpoint of each edge in between.
	{
		const int32_t nextEdge = min((x + TILE_W) & ~(TILE_W - 1), ibe);
		while (x < nextEdge) {
#pragma unroll
			for (int i = 0; i < 12; i++) { v[i] += dx2[i / 4]; }
			x += 2;
		}
	}
	while (x < ibe) { // x is a tile edge inside the run
		chk_store(recs + (x / TILE_W - firstColumn), 1, v, 18, x);
		if (x >= lastEdge) { return; }
		const int32_t stop = min(x + TILE_W, ibe);
		while (x < stop) {
#pragma unroll
			for (int i = 0; i < 12; i++) { v[i] += dx2[i / 4]; }
			x += 2;
		}
	}
#pragma unroll
	for (int k = 0; k < 6; k++) { v[k] = v[12 + k]; }
	if ((x & (TILE_W - 1)) == 0 && x < obe) { chk_store(recs + (x / TILE_W - firstColumn), 2, v, 6, x); }
	if (x >= lastEdge) { return; }
	while (x < obe) {
#pragma unroll
		for (int k = 0; k < 6; k++) { v[k] += dx2[k % 3]; }
		x += 2;
		if ((x & (TILE_W - 1)) == 0 && x < obe) { chk_store(recs + (x / TILE_W - firstColumn), 2, v, 6, x); }
		if (x >= lastEdge) { return; }
	}
}

#ifndef SETUP_MIN_BLOCKS
#define SETUP_MIN_BLOCKS 12 // 85 registers: measured 269 us against 314 us (8 blocks, 126 registers) for 2 M tiny triangles
#endif
template <bool EMIT>
__global__ void __launch_bounds__(SETUP_THREADS, SETUP_MIN_BLOCKS) setup_kernel(FrameDev frame) {
	__shared__ TaskParams task;
	__shared__ uint32_t warpCmds[SETUP_THREADS / 32], warpRows[SETUP_THREADS / 32];
	__shared__ BigItem sBig[SETUP_THREADS];
	__shared__ uint32_t sBigCount, sChkCount, sUnitCount, sItemBase, sUnitBase, sChkBase;
	__shared__ uint32_t sUnitEnd[SETUP_THREADS]; // emit pass: inclusive prefix of tile rows over the queued commands
	chain_enter();
	if (EMIT && frame.checkCaps != 0) {
		// the counting pass is complete: every CTA derives the verdict itself; the first one tells the host and the kernels behind this one
		const bool fits = frame_fits(frame);
		if (blockIdx.x == 0 && threadIdx.x == 0) {
			frame.totals[TOTAL_OVERFLOW] = fits ? 0u : 1u;
			publish_totals(frame, fits);
		}
		if (!fits) { return; }
	}
	{
		if (threadIdx.x == 0) { sChkCount = 0; sUnitCount = 0; }
		int32_t t = frame.blockTask ? frame.blockTask[blockIdx.x] : task_of_block(frame.tasks, frame.taskCount, (int32_t)blockIdx.x);
		for (uint32_t w = threadIdx.x; w < sizeof(TaskParams) / 4; w += blockDim.x) { ((uint32_t *)&task)[w] = ((const uint32_t *)&frame.tasks[t])[w]; }
		if (threadIdx.x == 0) { sBigCount = 0; }
	}
	__syncthreads();
	if (task.cullOnDevice) { // uniform per CTA: the model's bounding box against the cull frustum and the occluders, like model_render_threaded
		__shared__ int sCulled;
		if (threadIdx.x == 0) { sCulled = task_is_culled(frame, task) ? 1 : 0; }
		__syncthreads();
		if (sCulled) {
			if (!EMIT) {
				const int32_t culledSlot = task.slotBase + ((int32_t)blockIdx.x - task.blockBase) * SETUP_THREADS + (int32_t)threadIdx.x;
				if (culledSlot < task.slotBase + task.slotCount) { frame.slotCounts[culledSlot] = 0u; }
				if (threadIdx.x == 0) { frame.blockCmds[blockIdx.x] = 0u; frame.blockRows[blockIdx.x] = 0u; }
			}
			return;
		}
	}
	const ViewDev &view = frame.views[task.view];
	const int32_t width = view.width, height = view.height, clipTop = view.clipTop, clipBottom = view.clipBottom;
	const int32_t tilesX = view.tilesX;
	const uint32_t tileBase = view.tileBase;

	const int32_t localBlock = (int32_t)blockIdx.x - task.blockBase;
	const int32_t local = localBlock * SETUP_THREADS + threadIdx.x;
	const bool active = local < task.slotCount;
	const int32_t slot = task.slotBase + local;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	PPoint p[3];
	float colors[3][4], tex[3][4];
	const bool loaded = active && load_triangle(task, local, p, colors, tex);

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
		cmdBase = frame.blockCmds[blockIdx.x] + preCmd + incCmd - nCmd;
		rowBase = frame.blockRows[blockIdx.x] + preRows + incRows - nRows;
	}

	uint32_t countCmd = 0, countRows = 0;
	if (loaded) {
		float alpha[3] = {colors[0][3], colors[1][3], colors[2][3]};
		for_each_command(task, p, alpha, [&](const PPoint *q, const float *subB, const float *subC) {
			Bound bound = raster_bound(q, width, clipTop, clipBottom);
			// ref: api/rendererAPI.cpp:193-217 — a command that the occlusion grid hides stays in the queue (it keeps its index) but draws nothing
			const int32_t rowCount = (bound.any && !command_occluded(frame, bound, q)) ? bound.b - bound.t : 0;
			const int32_t tx0 = bound.l / TILE_W, tx1 = (bound.r - 1) / TILE_W, ty0 = bound.t / TILE_H, ty1 = (min(bound.b, height) - 1) / TILE_H;
			const bool small = rowCount <= frame.smallRows && (bound.r - bound.l) <= SMALL_WIDTH;
			if (!EMIT) {
				if (rowCount > 0) {
					if (!small && !task.depthOnly && frame.checkpoints) { atomicAdd(&sChkCount, (uint32_t)((rowCount / 2) * (tx1 - tx0 + 1))); }
					if (!small) { atomicAdd(&sUnitCount, (uint32_t)((bound.t + rowCount - 1) / TILE_H - bound.t / TILE_H + 1)); }
					int32_t tiles = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
					uint32_t queued = 0xFFFFFFFFu;
					if (tiles > SMALL_TILES) {
						queued = atomicAdd(&sBigCount, 1u);
						if (queued < (uint32_t)SETUP_THREADS) {
							BigItem &it = sBig[queued];
							it.tileBase = tileBase; it.tilesX = tilesX; it.tx0 = tx0; it.tx1 = tx1; it.ty0 = ty0; it.ty1 = ty1;
						}
					}
					if (queued >= (uint32_t)SETUP_THREADS) {
						for (int32_t ty = ty0; ty <= ty1; ty++) {
							for (int32_t tx = tx0; tx <= tx1; tx++) { atomicAdd(&frame.tileCount[tileBase + (uint32_t)(ty * tilesX + tx)], 1u); }
						}
					}
				}
			} else {
				const uint32_t index = cmdBase + countCmd;
				Cmd cmd;
				const bool perspective = task.camera.perspective != 0;
				get_projection(cmd, q, subB, subC, perspective);
				cmd.chkOffset = CHK_NONE; cmd.chkShape = 0u;
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
				// large triangles go to a whole warp: it scan-converts, bins and leaves interpolation checkpoints for the tile kernel
				uint32_t queued = 0xFFFFFFFFu;
				if (rowCount > 0 && !small) {
					queued = atomicAdd(&sBigCount, 1u);
					if (queued < (uint32_t)SETUP_THREADS) {
						BigItem &it = sBig[queued];
						it.cmdIndex = index; it.rowOffset = cmd.rowOffset; it.tileBase = tileBase; it.tilesX = tilesX;
						it.l = bound.l; it.t = bound.t; it.r = bound.r; it.rowCount = rowCount;
						it.tx0 = tx0; it.tx1 = tx1; it.ty0 = 0; it.ty1 = height;
						for (int k = 0; k < 3; k++) { it.fx[k] = q[k].fx; it.fy[k] = q[k].fy; }
						it.chkOffset = CHK_NONE;
#ifndef DFPSR_NO_CHK
						if (!task.depthOnly && frame.checkpoints) {
							// position inside the block's share of the checkpoint pool; the block reserves its share with ONE global atomic
							// after the barrier below and patches the command records (one cursor bumped by every large command of a
							// frame is a serial chain of same-address atomics)
							it.chkOffset = atomicAdd(&sChkCount, (uint32_t)((rowCount / 2) * (tx1 - tx0 + 1)));
							cmd.chkOffset = it.chkOffset;
							cmd.chkShape = (uint32_t)tx0 | ((uint32_t)(tx1 - tx0 + 1) << 16);
						}
#endif
					}
				}
				{
					// assembled field by field: copying the record through a uint4 view of its address would put it on the stack
					uint4 *dst = (uint4 *)&frame.cmds[index];
#define FU(x) __float_as_uint(x)
					dst[0] = make_uint4(FU(cmd.start[0]), FU(cmd.start[1]), FU(cmd.start[2]), FU(cmd.dx[0]));
					dst[1] = make_uint4(FU(cmd.dx[1]), FU(cmd.dx[2]), FU(cmd.dy[0]), FU(cmd.dy[1]));
					dst[2] = make_uint4(FU(cmd.dy[2]), cmd.flags, cmd.chkOffset, cmd.chkShape);
					dst[3] = make_uint4((uint32_t)cmd.rowStart, (uint32_t)cmd.rowCount, cmd.rowOffset, 0u);
					dst[4] = make_uint4(FU(cmd.red[0]), FU(cmd.red[1]), FU(cmd.red[2]), FU(cmd.green[0]));
					dst[5] = make_uint4(FU(cmd.green[1]), FU(cmd.green[2]), FU(cmd.blue[0]), FU(cmd.blue[1]));
					dst[6] = make_uint4(FU(cmd.blue[2]), FU(cmd.alpha[0]), FU(cmd.alpha[1]), FU(cmd.alpha[2]));
					// words 7..9 are only read for commands with a texture; word 7 is written anyway: it shares a 32-byte sector with word 6,
					// and a sector that is only half written has to be filled from DRAM when it leaves the L2
					dst[7] = make_uint4(FU(cmd.u1[0]), FU(cmd.u1[1]), FU(cmd.u1[2]), FU(cmd.v1[0]));
					if (task.diffuseIndex >= 0 || task.lightIndex >= 0) {
						dst[8] = make_uint4(FU(cmd.v1[1]), FU(cmd.v1[2]), FU(cmd.u2[0]), FU(cmd.u2[1]));
						dst[9] = make_uint4(FU(cmd.u2[2]), FU(cmd.v2[0]), FU(cmd.v2[1]), FU(cmd.v2[2]));
					}
#undef FU
				}
				if (rowCount > 0) {
					if (queued >= (uint32_t)SETUP_THREADS) {
						// scan conversion by this thread; a tile row (TILE_H rows) is binned when its last row is done
						long long fx[3] = {q[0].fx, q[1].fx, q[2].fx}, fy[3] = {q[0].fy, q[1].fy, q[2].fy};
						EdgeSet edges;
						edges_setup(edges, fx, fy, bound.l, bound.t, bound.r);
						int32_t minL = 0x7FFFFFFF, maxR = -1;
						const bool narrow = bound.l >= 1 && bound.r - bound.l <= NARROW_WIDTH;
						for (int32_t r = 0; r < rowCount; r++) {
							int32_t y = bound.t + r;
							int2 row = narrow ? edges_row_narrow(edges, y) : edges_row(edges, y);
							frame.rows[cmd.rowOffset + r] = row;
							if (row.y > row.x && y < height) { minL = min(minL, row.x); maxR = max(maxR, row.y); }
							if ((y & (TILE_H - 1)) == TILE_H - 1 || r == rowCount - 1) {
								if (y - (y & (TILE_H - 1)) < height) { emit_tile_row(frame, tileBase, tilesX, y / TILE_H, minL, maxR, index); }
								minL = 0x7FFFFFFF; maxR = -1;
							}
						}
					}
				}
			}
			countCmd++;
			countRows += (uint32_t)rowCount;
		});
	}

	// ---- large commands: the count pass walks bounding boxes warp by warp; the emit pass flattens them into (command, tile row) units and
	// gives every thread of the block the same number of units, so a block full of tall triangles is not eight warps working through 32
	// commands each with a third of their lanes busy (that serial tail was 200 us of a single 1080p frame)
	__syncthreads();
	if (!EMIT) {
		const uint32_t bigCount = min(sBigCount, (uint32_t)SETUP_THREADS);
		for (uint32_t b = warp; b < bigCount; b += SETUP_THREADS / 32) {
			const BigItem &it = sBig[b];
			// wide boxes: lanes over the columns of a row; narrow ones: one lane per tile row (no division per tile)
			if (it.tx1 - it.tx0 >= 16) {
				for (int32_t ty = it.ty0; ty <= it.ty1; ty++) {
					for (int32_t tx = it.tx0 + lane; tx <= it.tx1; tx += 32) { atomicAdd(&frame.tileCount[it.tileBase + (uint32_t)(ty * it.tilesX + tx)], 1u); }
				}
			} else {
				for (int32_t ty = it.ty0 + lane; ty <= it.ty1; ty += 32) {
					for (int32_t tx = it.tx0; tx <= it.tx1; tx++) { atomicAdd(&frame.tileCount[it.tileBase + (uint32_t)(ty * it.tilesX + tx)], 1u); }
				}
			}
		}
	} else {
		const uint32_t bigCount = min(sBigCount, (uint32_t)SETUP_THREADS);
		// inclusive prefix of tile rows per queued command (bigCount <= SETUP_THREADS: one command per thread)
		uint32_t mine = 0;
		if (threadIdx.x < bigCount) { const BigItem &it = sBig[threadIdx.x]; mine = (uint32_t)((it.t + it.rowCount - 1) / TILE_H - it.t / TILE_H + 1); }
		uint32_t inclusive = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t a = __shfl_up_sync(0xffffffffu, inclusive, d); if (lane >= d) { inclusive += a; } }
		if (lane == 31) { warpCmds[warp] = inclusive; }
		__syncthreads();
		uint32_t before = 0;
		for (int w = 0; w < warp; w++) { before += warpCmds[w]; }
		sUnitEnd[threadIdx.x] = before + inclusive;
		__syncthreads();
		const uint32_t unitTotal = bigCount > 0 ? sUnitEnd[bigCount - 1] : 0u;
		// the block reserves its share of the frame-wide queues with one atomic each; big_units_kernel then gives every unit its own thread
		if (threadIdx.x == 0 && bigCount > 0) {
			sItemBase = atomicAdd(&frame.totals[8], bigCount); sUnitBase = atomicAdd(&frame.totals[9], unitTotal);
			sChkBase = sChkCount > 0 ? atomicAdd(&frame.totals[5], sChkCount) : 0u;
		}
		__syncthreads();
		if (threadIdx.x < bigCount) {
			BigItem it = sBig[threadIdx.x];
			it.unitStart = sUnitBase + (threadIdx.x > 0 ? sUnitEnd[threadIdx.x - 1] : 0u);
			if (it.chkOffset != CHK_NONE) { it.chkOffset += sChkBase; frame.cmds[it.cmdIndex].chkOffset = it.chkOffset; }
			frame.bigItems[sItemBase + threadIdx.x] = it;
		}
		for (uint32_t u = threadIdx.x; u < unitTotal; u += SETUP_THREADS) {
			uint32_t lo = 0, hi = bigCount - 1; // first command whose inclusive end exceeds u
			while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sUnitEnd[mid] > u) { hi = mid; } else { lo = mid + 1; } }
			frame.bigUnits[sUnitBase + u] = sItemBase + lo;
		}
	}

	if (!EMIT) {
		if (threadIdx.x == 0 && sChkCount > 0) { atomicAdd(&frame.totals[4], sChkCount); }
		if (threadIdx.x == 0 && sUnitCount > 0) { atomicAdd(&frame.totals[7], sUnitCount); }
		if (active) { frame.slotCounts[slot] = countCmd | (countRows << 3); }
		// block totals for scan_blocks_kernel
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
			frame.blockCmds[blockIdx.x] = a;
			frame.blockRows[blockIdx.x] = b;
		}
	}
}

// One thread per (large command, tile row): scan-converts the row pairs of that tile row (ref: ITriangle2D.cpp:86-176), bins the command to
// the tiles they touch and walks the interpolation chains across the tile columns, leaving the checkpoints the tile kernel continues from.
// The units of the whole batch are spread over the whole grid, so one 1080p frame (about 60 k units) already fills the machine.
// SPLIT: two threads per unit (adjacent lanes), one per row pair of the tile row — twice the threads and half the serial chain per thread for
// frames whose units do not fill the machine anyway (one 1080p terrain frame has 15 k units); the bins of the two row pairs meet in a shuffle.
#ifndef BIG_UNITS_THREADS
#define BIG_UNITS_THREADS 64 // measured on the 256-view batch: 4.27 us per frame against 4.69 (256 threads) and 4.45 (128)
#endif
template <bool SPLIT>
__global__ void __launch_bounds__(BIG_UNITS_THREADS) big_units_kernel(FrameDev frame) {
	chain_enter();
	if (frame_dropped(frame)) { return; }
	// The grid is sized by the host from an earlier frame's unit count (it does not wait for this frame's): CTAs stride over the units
	// the counting pass found. Commands that overflowed a set-up block's queue were finished by their own thread, so the cursor
	// (totals[9]) holds the number of units that were really queued.
	const uint32_t unitTotal = min(frame.totals[7], frame.totals[9]);
	const uint32_t threadTotal = SPLIT ? 2u * unitTotal : unitTotal;
	for (uint32_t first = blockIdx.x * blockDim.x; first < threadTotal; first += gridDim.x * blockDim.x) {
		const uint32_t thread = first + threadIdx.x;
		const uint32_t u = SPLIT ? thread >> 1 : thread;
		const int32_t half = SPLIT ? (int32_t)(thread & 1u) : 0;
		const bool valid = u < unitTotal;
		if (!SPLIT && !valid) { continue; }
		int32_t minL = 0x7FFFFFFF, maxR = -1;
		int32_t ty = 0, height = 0;
		const BigItem *item = nullptr;
		if (valid) {
			const BigItem &it = frame.bigItems[frame.bigUnits[u]];
			item = &it;
			ty = it.t / TILE_H + (int32_t)(u - it.unitStart);
			height = it.ty1; // the view's height travels in ty1 (unused by the emit pass)
			const int32_t yBegin = max(it.t, ty * TILE_H), yEnd = min(it.t + it.rowCount, ty * TILE_H + TILE_H);
			const int32_t yFirst = SPLIT ? yBegin + (TILE_H / 2) * half : yBegin, yLast = SPLIT ? min(yEnd, yFirst + TILE_H / 2) : yEnd;
			if (yFirst < yLast) {
				EdgeSet edges;
				long long fx[3] = {it.fx[0], it.fx[1], it.fx[2]}, fy[3] = {it.fy[0], it.fy[1], it.fy[2]};
				edges_setup(edges, fx, fy, it.l, it.t, it.r);
				float start[3], dx[3], dy[3];
				const bool checkpoints = it.chkOffset != CHK_NONE;
				if (checkpoints) {
					const float *planes = (const float *)&frame.cmds[it.cmdIndex];
#pragma unroll
					for (int k = 0; k < 3; k++) { start[k] = planes[k]; dx[k] = planes[3 + k]; dy[k] = planes[6 + k]; }
				}
				const int32_t columns = it.tx1 - it.tx0 + 1;
				for (int32_t y = yFirst; y < yLast; y += 2) { // rows come in even-aligned pairs
					int2 upperRow = edges_row(edges, y), lowerRow = edges_row(edges, y + 1);
					*(int4 *)&frame.rows[it.rowOffset + (uint32_t)(y - it.t)] = make_int4(upperRow.x, upperRow.y, lowerRow.x, lowerRow.y);
					if (upperRow.y > upperRow.x && y < height) { minL = min(minL, upperRow.x); maxR = max(maxR, upperRow.y); }
					if (lowerRow.y > lowerRow.x && y + 1 < height) { minL = min(minL, lowerRow.x); maxR = max(maxR, lowerRow.y); }
					if (checkpoints && y < height) {
						chk_walk_row_pair(start, dx, dy, upperRow, lowerRow, y, frame.chk + it.chkOffset + (size_t)((y - it.t) / 2) * (size_t)columns, it.tx0);
					}
				}
			}
		}
		if (SPLIT) {
			// the loop bound is warp-uniform (first and threadTotal are), so every lane of the warp reaches the shuffles
			minL = min(minL, __shfl_xor_sync(0xffffffffu, minL, 1));
			maxR = max(maxR, __shfl_xor_sync(0xffffffffu, maxR, 1));
			if (!valid || half != 0) { continue; }
		}
		if (ty * TILE_H < height) { emit_tile_row(frame, item->tileBase, item->tilesX, ty, minL, maxR, item->cmdIndex); }
	}
}

// ref: api/rendererAPI.cpp:242-258 occludeFromExistingTriangles: every solid command queued so far is an occluder for the cells that lie
// completely inside it (occludeFromSortedHull, :218-241). The grid keeps the minimum, so the order of the atomics does not matter.
__global__ void __launch_bounds__(SETUP_THREADS) occlude_existing_kernel(FrameDev frame, float *grid) {
	__shared__ TaskParams task;
	{
		int32_t t = frame.blockTask ? frame.blockTask[blockIdx.x] : task_of_block(frame.tasks, frame.taskCount, (int32_t)blockIdx.x);
		for (uint32_t w = threadIdx.x; w < sizeof(TaskParams) / 4; w += blockDim.x) { ((uint32_t *)&task)[w] = ((const uint32_t *)&frame.tasks[t])[w]; }
	}
	__syncthreads();
	if (task.filter != DFPSR_FILTER_SOLID) { return; }
	if (task.cullOnDevice) { // a model the broad phase drops has no commands that could occlude
		__shared__ int sCulled;
		if (threadIdx.x == 0) { sCulled = task_is_culled(frame, task) ? 1 : 0; }
		__syncthreads();
		if (sCulled) { return; }
	}
	const ViewDev &view = frame.views[task.view];
	const int32_t local = ((int32_t)blockIdx.x - task.blockBase) * SETUP_THREADS + threadIdx.x;
	if (local >= task.slotCount) { return; }
	PPoint p[3];
	float colors[3][4], tex[3][4];
	if (!load_triangle(task, local, p, colors, tex)) { return; }
	float alpha[3] = {colors[0][3], colors[1][3], colors[2][3]};
	const int32_t width = view.width, clipTop = view.clipTop, clipBottom = view.clipBottom;
	for_each_command(task, p, alpha, [&](const PPoint *q, const float *, const float *) {
		Bound bound = raster_bound(q, width, clipTop, clipBottom);
		if (!(bound.wr - bound.wl > CELL_SIZE && bound.wb - bound.wt > CELL_SIZE)) { return; }
		float distance = fmaxf(0.0f, fmaxf(q[0].csz, fmaxf(q[1].csz, q[2].csz)));
		const CellBound cells = outer_cell_bound(bound.wl, bound.wt, bound.wr, bound.wb, frame.gridWidth, frame.gridHeight);
		for (int32_t cy = cells.y0; cy < cells.y1; cy++) {
			for (int32_t cx = cells.x0; cx < cells.x1; cx++) {
				if (cell_inside_of_hull(q, 3, cx, cy)) { atomicMin((int *)&grid[cy * frame.gridStride + cx], __float_as_int(distance)); }
			}
		}
	});
}

// The end of the counting pass in ONE launch (three kernels in round 1: a single 1080p frame is bound by launches, not by work):
//   CTA 0       exclusive scans of the per-block command and row totals (submission order is preserved). Every thread owns SCAN_ITEMS
//               consecutive blocks per round, so 2 M tiny triangles (16 k set-up blocks) need two rounds instead of sixteen.
//   other CTAs  every tile takes a segment of the entry pool sized by its counted upper bound. Order between tiles is irrelevant, so
//               no scan: one atomicAdd per warp.
//   last CTA to finish: compares the frame's totals with the pools it was launched with (frame.cap*) and publishes totals and verdict
//               to the host through mapped pinned memory (not through a device-to-host copy: a copy would queue on the copy engine
//               behind megabytes of finished frames travelling to the host and stall the next chunk's set-up for milliseconds).
static const int SCAN_ITEMS = 8;
static const int COUNTS_THREADS = 1024;
__global__ void __launch_bounds__(COUNTS_THREADS) counts_kernel(FrameDev frame) {
	__shared__ uint32_t warpSum[2][32];
	__shared__ uint32_t carry[2];
	__shared__ bool sLast;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	chain_enter();
	if (blockIdx.x == 0) {
		if (threadIdx.x < 2) { carry[threadIdx.x] = 0; }
		__syncthreads();
		for (int32_t base = 0; base < frame.blockCount; base += COUNTS_THREADS * SCAN_ITEMS) {
			const int32_t first = base + (int32_t)threadIdx.x * SCAN_ITEMS;
			uint32_t v[2][SCAN_ITEMS], total[2] = {0u, 0u};
			if (first + SCAN_ITEMS <= frame.blockCount) { // 32 consecutive bytes per thread and array: two 16-byte loads, a warp reads 1 KB contiguously
				const uint4 a0 = *(const uint4 *)(frame.blockCmds + first), a1 = *(const uint4 *)(frame.blockCmds + first + 4);
				const uint4 b0 = *(const uint4 *)(frame.blockRows + first), b1 = *(const uint4 *)(frame.blockRows + first + 4);
				v[0][0] = a0.x; v[0][1] = a0.y; v[0][2] = a0.z; v[0][3] = a0.w; v[0][4] = a1.x; v[0][5] = a1.y; v[0][6] = a1.z; v[0][7] = a1.w;
				v[1][0] = b0.x; v[1][1] = b0.y; v[1][2] = b0.z; v[1][3] = b0.w; v[1][4] = b1.x; v[1][5] = b1.y; v[1][6] = b1.z; v[1][7] = b1.w;
			} else {
#pragma unroll
				for (int e = 0; e < SCAN_ITEMS; e++) {
					const bool valid = first + e < frame.blockCount;
					v[0][e] = valid ? frame.blockCmds[first + e] : 0u; v[1][e] = valid ? frame.blockRows[first + e] : 0u;
				}
			}
#pragma unroll
			for (int e = 0; e < SCAN_ITEMS; e++) { total[0] += v[0][e]; total[1] += v[1][e]; }
			uint32_t inc[2] = {total[0], total[1]};
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
				for (int k = 0; k < 2; k++) {
					const uint32_t a = __shfl_up_sync(0xffffffffu, inc[k], d);
					if (lane >= d) { inc[k] += a; }
				}
			}
			if (lane == 31) { for (int k = 0; k < 2; k++) { warpSum[k][warp] = inc[k]; } }
			__syncthreads();
			if (warp == 0) { // inclusive scan of the 32 warp totals
				uint32_t w[2] = {warpSum[0][lane], warpSum[1][lane]};
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
					for (int k = 0; k < 2; k++) {
						const uint32_t a = __shfl_up_sync(0xffffffffu, w[k], d);
						if (lane >= d) { w[k] += a; }
					}
				}
				warpSum[0][lane] = w[0]; warpSum[1][lane] = w[1];
			}
			__syncthreads();
			uint32_t running[2];
#pragma unroll
			for (int k = 0; k < 2; k++) { running[k] = carry[k] + (warp > 0 ? warpSum[k][warp - 1] : 0u) + inc[k] - total[k]; }
			uint32_t out[2][SCAN_ITEMS];
#pragma unroll
			for (int e = 0; e < SCAN_ITEMS; e++) { out[0][e] = running[0]; out[1][e] = running[1]; running[0] += v[0][e]; running[1] += v[1][e]; }
			if (first + SCAN_ITEMS <= frame.blockCount) {
				*(uint4 *)(frame.blockCmds + first) = make_uint4(out[0][0], out[0][1], out[0][2], out[0][3]); *(uint4 *)(frame.blockCmds + first + 4) = make_uint4(out[0][4], out[0][5], out[0][6], out[0][7]);
				*(uint4 *)(frame.blockRows + first) = make_uint4(out[1][0], out[1][1], out[1][2], out[1][3]); *(uint4 *)(frame.blockRows + first + 4) = make_uint4(out[1][4], out[1][5], out[1][6], out[1][7]);
			} else {
#pragma unroll
				for (int e = 0; e < SCAN_ITEMS; e++) { if (first + e < frame.blockCount) { frame.blockCmds[first + e] = out[0][e]; frame.blockRows[first + e] = out[1][e]; } }
			}
			__syncthreads();
			if (threadIdx.x == 0) { for (int k = 0; k < 2; k++) { carry[k] += warpSum[k][31]; } }
			__syncthreads();
		}
		if (threadIdx.x == 0) {
			frame.totals[0] = carry[0];
			frame.totals[1] = carry[1];
		}
	} else {
		const uint32_t i = (blockIdx.x - 1u) * COUNTS_THREADS + threadIdx.x;
		const uint32_t c = i < frame.tileTotal ? frame.tileCount[i] : 0u;
		uint32_t inc = c, mx = c;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t a = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) { inc += a; }
		}
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) { mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d)); }
		// one atomic per CTA (round 1: one per warp — 2 000 same-address atomics for a single 1080p frame)
		if (lane == 31) { warpSum[0][warp] = inc; }
		if (lane == 0) { warpSum[1][warp] = mx; }
		__syncthreads();
		if (warp == 0) {
			uint32_t total = warpSum[0][lane], most = warpSum[1][lane];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t a = __shfl_up_sync(0xffffffffu, total, d);
				if (lane >= d) { total += a; }
			}
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) { most = max(most, __shfl_xor_sync(0xffffffffu, most, d)); }
			uint32_t base = 0;
			if (lane == 31 && total > 0) { base = atomicAdd(&frame.totals[2], total); }
			base = __shfl_sync(0xffffffffu, base, 31);
			if (lane == 0 && most > 0) { atomicMax(&frame.totals[3], most); }
			warpSum[0][lane] = base + total - warpSum[0][lane]; // exclusive offset of every warp of the CTA
		}
		__syncthreads();
		if (i < frame.tileTotal) {
			frame.tileOffset[i] = warpSum[0][warp] + inc - c;
			frame.tileCursor[i] = 0;
		}
	}
	// ---- a frame whose host side waits for the counts: the last CTA to arrive sees every total and publishes them. An asynchronous frame
	// needs none of this (no fence, no ticket): every CTA of setup_kernel<true> derives the verdict from the finished totals itself and
	// its first CTA publishes it.
	if (frame.checkCaps != 0) { return; }
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) { sLast = atomicAdd(&frame.totals[TOTAL_TICKET], 1u) == gridDim.x - 1u; }
	__syncthreads();
	if (!sLast) { return; }
	__threadfence();
	if (threadIdx.x == 0) { publish_totals(frame, true); }
}

// Restores ascending command order in the lists that hold more than LOCAL_SORT entries (shorter ones are sorted in registers by raster_kernel).
// Every thread looks at one tile; the long ones are queued in shared memory and sorted by the whole CTA one after the other.
__global__ void __launch_bounds__(SORT_THREADS) sort_lists_kernel(FrameDev frame) {
	// launched with every frame whose host side does not wait for the counts: nothing to do unless some tile list is long
	chain_enter();
	if (frame_dropped(frame) || frame.totals[3] <= (uint32_t)LOCAL_SORT) { return; }
	__shared__ uint32_t s[SORT_SMEM];
	__shared__ uint32_t sQueue[SORT_THREADS];
	__shared__ uint32_t sQueued;
	if (threadIdx.x == 0) { sQueued = 0; }
	__syncthreads();
	{
		const uint32_t tile = blockIdx.x * SORT_THREADS + threadIdx.x;
		if (tile < frame.tileTotal && frame.tileCursor[tile] > (uint32_t)LOCAL_SORT) { sQueue[atomicAdd(&sQueued, 1u)] = tile; }
	}
	__syncthreads();
	const uint32_t queued = sQueued;
	// lists of up to SORT_WARP entries: one warp each, all warps of the CTA in parallel
	{
		const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
		uint32_t *ws = s + warp * SORT_WARP;
		for (uint32_t q = warp; q < queued; q += SORT_THREADS / 32) {
			const uint32_t tile = sQueue[q];
			const uint32_t n = frame.tileCursor[tile];
			if (n > (uint32_t)SORT_WARP) { continue; }
			uint32_t *list = frame.tileList + frame.tileOffset[tile];
			uint32_t size = 64;
			while (size < n) { size <<= 1; }
			for (uint32_t i = lane; i < size; i += 32) { ws[i] = i < n ? list[i] : 0xFFFFFFFFu; }
			__syncwarp();
			for (uint32_t k = 2; k <= size; k <<= 1) {
				for (uint32_t j = k >> 1; j > 0; j >>= 1) {
					for (uint32_t i = lane; i < size; i += 32) {
						uint32_t l = i ^ j;
						if (l > i) {
							uint32_t a = ws[i], b = ws[l];
							bool ascending = (i & k) == 0;
							if ((a > b) == ascending) { ws[i] = b; ws[l] = a; }
						}
					}
					__syncwarp();
				}
			}
			for (uint32_t i = lane; i < n; i += 32) { list[i] = ws[i]; }
			__syncwarp();
		}
	}
	__syncthreads();
	for (uint32_t q = 0; q < queued; q++) {
		const uint32_t tile = sQueue[q];
		const uint32_t n = frame.tileCursor[tile];
		if (n <= (uint32_t)SORT_WARP) { continue; }
		uint32_t *list = frame.tileList + frame.tileOffset[tile];
		if (n <= (uint32_t)SORT_SMEM) {
			uint32_t size = 64;
			while (size < n) { size <<= 1; }
			for (uint32_t i = threadIdx.x; i < size; i += SORT_THREADS) { s[i] = i < n ? list[i] : 0xFFFFFFFFu; }
			__syncthreads();
			for (uint32_t k = 2; k <= size; k <<= 1) {
				for (uint32_t j = k >> 1; j > 0; j >>= 1) {
					for (uint32_t i = threadIdx.x; i < size; i += SORT_THREADS) {
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
			for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS) { list[i] = s[i]; }
			__syncthreads();
		} else {
			// rank sort (keys are distinct): O(n^2 / threads), only for pathological pile-ups of thousands of triangles on one tile.
			// Scratch: every such list claims its own n entries of a pool as large as the whole entry pool (totals[6] is the cursor).
			__shared__ uint32_t sSegment;
			if (threadIdx.x == 0) { sSegment = atomicAdd(&frame.totals[6], n); }
			__syncthreads();
			uint32_t *tmp = frame.sortTmp + sSegment;
			for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS) {
				uint32_t key = list[i], rank = 0;
				for (uint32_t j = 0; j < n; j++) { rank += list[j] < key ? 1u : 0u; }
				tmp[rank] = key;
			}
			__syncthreads();
			for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS) { list[i] = tmp[i]; }
			__syncthreads();
		}
	}
}

// ------------------------------------------------------------------------------------------------ texture sampling

// The frame's texture table lives in device memory next to the task records (any number of textures up to the 12 index bits of Cmd::flags).
__device__ __forceinline__ TexDev load_tex(const TexDev *__restrict__ table, uint32_t index) {
	const uint4 a = __ldg((const uint4 *)(table + index)), b = __ldg((const uint4 *)(table + index) + 1);
	TexDev t;
	t.data = (const uint32_t *)(((unsigned long long)a.y << 32) | (unsigned long long)a.x);
	t.log2width = a.z; t.log2height = a.w; t.maxMipLevel = b.x; t.startOffset = b.y; t.maxLevelMask = b.z;
	return t;
}

// ref: api/textureAPI.h:253-263 weightColors on 16-bit lane pairs (sums never exceed 255 * 256, so lanes cannot carry)
// Bytes 1 and 3 are moved into the 16-bit lanes, and the high bytes of the four lane sums are gathered, with one byte permute each
// instead of shift + mask pairs: the integer ALU pipe is the busiest pipe of the tile kernel (ncu: 50 %), the multiplies run on the FMA pipe.
__device__ __forceinline__ uint32_t weight_colors(uint32_t colorA, uint32_t weightA, uint32_t colorB, uint32_t weightB) {
	uint32_t low = (colorA & 0x00FF00FFu) * weightA + (colorB & 0x00FF00FFu) * weightB;
	uint32_t high = __byte_perm(colorA, 0u, 0x4341) * weightA + __byte_perm(colorB, 0u, 0x4341) * weightB;
	return __byte_perm(low, high, 0x7351); // ((low >> 8) & 0x00FF00FF) | (high & 0xFF00FF00)
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
	// (row << log2Stride) | column == row * stride + column: two row pointers, four loads at 32-bit column offsets
	const uint32_t *upperRow = data + (top << log2Stride), *lowerRow = data + (bottom << log2Stride);
	uint32_t c00 = __ldg(upperRow + left), c10 = __ldg(upperRow + right);
	uint32_t c01 = __ldg(lowerRow + left), c11 = __ldg(lowerRow + right);
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

// __byte_perm selector that moves the bytes (red, green, blue, alpha) of a texel to the positions pack_shifts() describes
__host__ __device__ __forceinline__ uint32_t pack_selector(uint32_t shifts) {
	uint32_t selector = 0u;
#pragma unroll
	for (uint32_t channel = 0; channel < 4; channel++) { selector |= channel << (((shifts >> (8u * channel)) & 31u) >> 1); } // byte position p takes nibble p
	return selector;
}

// byte -> float without the conversion pipe: the byte becomes the low mantissa bits of 2^23, then 2^23 is subtracted (exact)
__device__ __forceinline__ void unpack_bytes(uint32_t c, float &r, float &g, float &b, float &a) {
	r = __uint_as_float(__byte_perm(c, 0x4B000000u, 0x7440)) - 8388608.0f; g = __uint_as_float(__byte_perm(c, 0x4B000000u, 0x7441)) - 8388608.0f;
	b = __uint_as_float(__byte_perm(c, 0x4B000000u, 0x7442)) - 8388608.0f; a = __uint_as_float(__byte_perm(c, 0x4B000000u, 0x7443)) - 8388608.0f;
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
		float r, g, b, a;
		unpack_bytes(c, r, g, b, a);
		if (MULTIPLY) { rgba[l][0] = rgba[l][0] * r; rgba[l][1] = rgba[l][1] * g; rgba[l][2] = rgba[l][2] * b; rgba[l][3] = rgba[l][3] * a; }
		else { rgba[l][0] = r; rgba[l][1] = g; rgba[l][2] = b; rgba[l][3] = a; }
	}
}

// The reciprocal of the interpolated 1/W. Exact mode: the IEEE quotient of the reference's scalar build. Tolerance mode: the hardware
// approximation (about one ulp, like the rcpps + Newton step of the reference's SSE build, base/simd.h:4047-4052).
template <bool EXACT>
__device__ __forceinline__ float reciprocal_w(float v) {
	if (EXACT) { return 1.0f / v; }
	return __fdividef(1.0f, v);
}

// ------------------------------------------------------------------------------------------------ tile kernel

// Checkpoint of one (command, row pair) at the left edge of the tile, written by lane (command, row pair) of the warp.
//   mode 0: v[0..5] = running sums (upper, lower) at position `at`, still left of (or at) the inner block start
//   mode 1: v[0..11] = the four lanes' sums inside the inner run at `at`, v[12..17] = sums after the run's closing multiplication
//   mode 2: v[0..5] = running sums at `at`, right of the inner run
//   mode 3: depth-only path: v[0], v[1] = depth at columns atU / atL of the upper / lower row
//   mode 4: tolerance mode: nothing stored, the quad lanes evaluate the planes directly
//   mode -1: nothing of this command in this row pair of this tile
struct Rec {
	int32_t ul, ur, ll, lr; // row intervals of the pair, as stored
	int32_t mode;           // from here on the layout of ChkRec: a stored checkpoint arrives as ONE 80-byte bulk copy (cp.async.bulk)
	float v[18];
	int32_t at;
};
static_assert(sizeof(Rec) == 96 && offsetof(Rec, mode) == 16 && offsetof(Rec, at) == 16 + offsetof(ChkRec, at), "Rec layout");
struct RecLite { // tolerance mode: the row intervals are all the quad lanes need
	int32_t ul, ur, ll, lr;
	int32_t mode, at;
	int32_t pad_[2];
};
static_assert(sizeof(RecLite) == 32, "RecLite layout");
template <bool EXACT> struct RecOf { typedef Rec type; };
template <> struct RecOf<false> { typedef RecLite type; };

// Ascending bitonic sort of one key per lane inside aligned groups of K lanes.
template <int K>
__device__ __forceinline__ uint32_t warp_sort(uint32_t key, int lane) {
#pragma unroll
	for (int k = 2; k <= K; k <<= 1) {
#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1) {
			uint32_t other = __shfl_xor_sync(0xffffffffu, key, j);
			bool ascending = (lane & k) == 0, lower = (lane & j) == 0;
			bool takeMin = ascending == lower;
			key = takeMin ? min(key, other) : max(key, other);
		}
	}
	return key;
}

// Ascending bitonic sort of 64 keys held two per lane: element i lives in lane (i & 31), register (i >> 5).
__device__ __forceinline__ void warp_sort64(uint32_t &k0, uint32_t &k1, int lane) {
#pragma unroll
	for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1) {
			if (j == 32) {
				// partners are the two registers of a lane; k == 64 here, every pair ascends
				const uint32_t lo = min(k0, k1), hi = max(k0, k1);
				k0 = lo; k1 = hi;
			} else {
				const uint32_t o0 = __shfl_xor_sync(0xffffffffu, k0, j), o1 = __shfl_xor_sync(0xffffffffu, k1, j);
				const bool lower = (lane & j) == 0;
				const bool up0 = (lane & k) == 0 || k == 64, up1 = ((lane + 32) & k) == 0; // direction of the block each element belongs to
				k0 = (up0 == lower) ? min(k0, o0) : max(k0, o0);
				k1 = (up1 == lower) ? min(k1, o1) : max(k1, o1);
			}
		}
	}
}

// Replay of the reference's running sums (shader/fillerTemplates.h:329-372) from the tile's checkpoint to the quad at column x0: the
// interpolated (1/W, U/W, V/W) of the quad's four lanes, bit for bit. clipSides tells whether the quad is one of the row pair's edge
// quads (lanes tested against their row's interval) or part of the unclipped inner run.
__device__ __forceinline__ void chain_to_quad(const Rec &rec, int32_t mode, const float *dx, int32_t x0, int32_t ibs, int32_t ibe, float lanes[3][4], bool &clipSides) {
	const float dx2[3] = {dx[0] * 2.0f, dx[1] * 2.0f, dx[2] * 2.0f};
	const int32_t at = rec.at;
	const bool noInner = ibe <= ibs;
	clipSides = true;
	if (mode == 1 && x0 < ibe) {
		clipSides = false;
#pragma unroll
		for (int k = 0; k < 3; k++) {
#pragma unroll
			for (int l = 0; l < 4; l++) { lanes[k][l] = rec.v[k * 4 + l]; }
		}
		for (int32_t s = at; s < x0; s += 2) {
#pragma unroll
			for (int k = 0; k < 3; k++) {
#pragma unroll
				for (int l = 0; l < 4; l++) { lanes[k][l] += dx2[k]; }
			}
		}
	} else {
		float up[3], lo[3];
		int32_t from = at;
		if (mode == 1) {
#pragma unroll
			for (int k = 0; k < 3; k++) { up[k] = rec.v[12 + k]; lo[k] = rec.v[15 + k]; }
			from = ibe;
		} else {
#pragma unroll
			for (int k = 0; k < 3; k++) { up[k] = rec.v[k]; lo[k] = rec.v[3 + k]; }
		}
		bool inner = false;
		if (mode == 0 && !noInner && x0 >= ibs) {
			// the inner run starts inside this tile: finish the left edge, then either enter the run or jump over it
			for (int32_t s = from; s < ibs; s += 2) {
#pragma unroll
				for (int k = 0; k < 3; k++) { up[k] += dx2[k]; lo[k] += dx2[k]; }
			}
			if (x0 < ibe) {
				inner = true;
				clipSides = false;
#pragma unroll
				for (int k = 0; k < 3; k++) { lanes[k][0] = up[k]; lanes[k][1] = up[k] + dx[k]; lanes[k][2] = lo[k]; lanes[k][3] = lo[k] + dx[k]; }
				for (int32_t s = ibs; s < x0; s += 2) {
#pragma unroll
					for (int k = 0; k < 3; k++) {
#pragma unroll
						for (int l = 0; l < 4; l++) { lanes[k][l] += dx2[k]; }
					}
				}
			} else {
				const float quadCount = (float)((ibe - ibs) / 2);
#pragma unroll
				for (int k = 0; k < 3; k++) { up[k] = up[k] + (dx2[k] * quadCount); lo[k] = lo[k] + (dx2[k] * quadCount); }
				from = ibe;
			}
		}
		if (!inner) {
			for (int32_t s = from; s < x0; s += 2) {
#pragma unroll
				for (int k = 0; k < 3; k++) { up[k] += dx2[k]; lo[k] += dx2[k]; }
			}
#pragma unroll
			for (int k = 0; k < 3; k++) { lanes[k][0] = up[k]; lanes[k][1] = up[k] + dx[k]; lanes[k][2] = lo[k]; lanes[k][3] = lo[k] + dx[k]; }
		}
	}
}

// One pixel of the deferred shading pass, for warps whose winners are not all plain textured commands: the colour command `cmd` gives the
// pixel whose interpolated (1/W, U/W, V/W) are (d, su, sv). PREPARED: a textured command left what its variant needs in (su, sv) when it
// took the pixel (see the visibility pass) — the diffuse texture coordinates (u, v) for a plain textured command, the vertex weights (B, C)
// for the other textured variants; commands without a texture always leave the raw sums.
// ref: shader/fillerTemplates.h:196-243 (weights), shader/RgbaMultiply.h:75-106 (variants), implementation/image/PackOrder.h:186-213 (pack)
template <bool EXACT, bool PREPARED>
__device__ __forceinline__ uint32_t shade_pixel(const Cmd *__restrict__ cmd, const TexDev *__restrict__ texTable, float d, float su, float sv, uint32_t mip, uint32_t shifts) {
	const uint32_t flags = __ldg(&cmd->flags);
	const bool hasDiffuse = (flags & CMD_HAS_DIFFUSE) != 0, hasLight = (flags & CMD_HAS_LIGHT) != 0;
	const bool fade = (flags & CMD_HAS_FADE) != 0, colorless = (flags & CMD_COLORLESS) != 0 && !fade;
	const bool plainDiffuse = hasDiffuse && !hasLight && colorless, plainLight = hasLight && !hasDiffuse && colorless;
	float r, g, b, a;
	if (PREPARED && plainDiffuse) {
		unpack_bytes(sample_bilinear(load_tex(texTable, (flags >> 8) & 0xFFFu), su, sv, mip), r, g, b, a);
		return pack_rgba_ordered(saturated_byte(r), saturated_byte(g), saturated_byte(b), saturated_byte(a), shifts);
	}
	float wb, wc;
	if ((flags & CMD_AFFINE) != 0 || (PREPARED && (hasDiffuse || hasLight))) { wb = su; wc = sv; }
	else { const float linearDepth = reciprocal_w<EXACT>(d); wb = su * linearDepth; wc = sv * linearDepth; }
	const float wa = 1.0f - (wb + wc);
	const float4 *words = (const float4 *)cmd; // 16-byte words of the record: 4..6 colours, 7..9 texture coordinates
	if (!(plainDiffuse || plainLight)) {
		const float4 c0 = __ldg(words + 4), c1 = __ldg(words + 5), c2 = __ldg(words + 6);
		if (fade) {
			const float red[3] = {c0.x, c0.y, c0.z}, green[3] = {c0.w, c1.x, c1.y}, blue[3] = {c1.z, c1.w, c2.x}, alpha[3] = {c2.y, c2.z, c2.w};
			r = interpolate3(red, wa, wb, wc); g = interpolate3(green, wa, wb, wc); b = interpolate3(blue, wa, wb, wc); a = interpolate3(alpha, wa, wb, wc);
		} else {
			r = c0.x; g = c0.w; b = c1.z; a = c2.y;
		}
	}
	if (hasDiffuse || hasLight) {
		const float4 t0 = __ldg(words + 7), t1 = __ldg(words + 8), t2 = __ldg(words + 9);
		if (hasDiffuse) {
			const float cu[3] = {t0.x, t0.y, t0.z}, cv[3] = {t0.w, t1.x, t1.y};
			const TexDev t = load_tex(texTable, (flags >> 8) & 0xFFFu);
			const uint32_t c = sample_bilinear(t, interpolate3(cu, wa, wb, wc), interpolate3(cv, wa, wb, wc), mip);
			float tr, tg, tb, ta;
			unpack_bytes(c, tr, tg, tb, ta);
			if (plainDiffuse) { r = tr; g = tg; b = tb; a = ta; } else { r = r * tr; g = g * tg; b = b * tb; a = a * ta; }
		}
		if (hasLight) {
			const float cu[3] = {t1.z, t1.w, t2.x}, cv[3] = {t2.y, t2.z, t2.w};
			const TexDev t = load_tex(texTable, flags >> 20);
			const uint32_t c = sample_bilinear(t, interpolate3(cu, wa, wb, wc), interpolate3(cv, wa, wb, wc), 0u);
			float tr, tg, tb, ta;
			unpack_bytes(c, tr, tg, tb, ta);
			if (plainLight) { r = tr; g = tg; b = tb; a = ta; } else { r = r * tr; g = g * tg; b = b * tb; a = a * ta; }
		}
	}
	return pack_rgba_ordered(saturated_byte(r), saturated_byte(g), saturated_byte(b), saturated_byte(a), shifts);
}

// ---- bulk asynchronous copies (the TMA's linear mode: cp.async.bulk, completion on a shared-memory mbarrier)
#ifndef DFPSR_CHK_BULK
// 1: stored checkpoints travel global -> shared as 80-byte bulk copies (UBLKCP + mbarrier) instead of five 16-byte loads and six
// shared-memory stores per lane. Built, bit-exact, and measured SLOWER (tile kernel 38.5 against 37.3 us per 1080p frame): UBLKCP issues
// from the uniform datapath, so the per-lane copies of a warp are serialised (nine instructions per record), while the vector loads of
// all lanes issue together. Kept for the record (profiles/r2_tma_experiment.md); off by default.
#define DFPSR_CHK_BULK 0
#endif
__device__ __forceinline__ uint32_t smem_address(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarrier_init(uint64_t *bar, uint32_t arrivals) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_address(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbarrier_expect(uint64_t *bar, uint32_t bytes) { // one arrival that announces `bytes` of bulk copies
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_address(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarrier_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred done;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n"
	    "@!done bra WAIT_%=;\n"
	    "}\n" ::"r"(smem_address(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_to_shared(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_address(dst)), "l"(src), "r"(bytes), "r"(smem_address(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

enum : int { TILE_IMMEDIATE = 0, TILE_DEPTH_ONLY = 1, TILE_DEFERRED = 2, TILE_DEFERRED_RAW = 3 }; // RAW: deferred shading without prepared shading inputs (frames without textures)
static const uint32_t NO_WINNER = 0xFFFFFFFFu;
static const uint32_t SHADER_PATTERN = CMD_AFFINE | CMD_HAS_DIFFUSE | CMD_HAS_LIGHT | CMD_HAS_FADE | CMD_COLORLESS | (0xFFFu << 8), PATTERN_TAKEN = 0x80000000u, PATTERN_MIXED = 0xFFFFFFFFu;

// MODE: TILE_IMMEDIATE  commands are shaded quad by quad in submission order (needed when a frame holds alpha-filtered commands)
//       TILE_DEPTH_ONLY the depth pass of model_renderDepth (renderCore.cpp:343-443)
//       TILE_DEFERRED   frames of solid commands: a visibility pass leaves, per pixel, the command that ends up visible (the depth test is
//                       strict, so that is the first command reaching the largest depth — exactly the command whose colour survives the
//                       reference's sequential loop) together with its interpolated (1/W, U/W, V/W); every pixel is then shaded ONCE.
// EXACT: true   the interpolated values replay the reference's chains of float additions (bit-identical to its scalar build)
//        false  tolerance mode: planes evaluated directly per quad, approximate reciprocal; no checkpoints anywhere (deferred mode only)
template <int MODE, bool EXACT>
__global__ void __launch_bounds__(RASTER_WARPS * 32, (MODE == TILE_DEFERRED ? (EXACT ? RASTER_MIN_BLOCKS_DEFERRED : RASTER_MIN_BLOCKS_TOLERANCE) : RASTER_MIN_BLOCKS)) raster_kernel(FrameDev frame) {
	constexpr bool DEPTH_ONLY = MODE == TILE_DEPTH_ONLY, DEFERRED = MODE == TILE_DEFERRED || MODE == TILE_DEFERRED_RAW;
#ifndef DFPSR_PREPARED_TOLERANCE
#define DFPSR_PREPARED_TOLERANCE 0 // tolerance mode: its hardware reciprocal makes the shading pass's division cheap, 25.4 us either way
#endif
	constexpr bool PREPARED = MODE == TILE_DEFERRED && (EXACT || DFPSR_PREPARED_TOLERANCE != 0); // textured commands hand their shading inputs to the shading pass
	static_assert(EXACT || DEFERRED, "tolerance mode exists for the deferred tile kernel only");
	typedef typename RecOf<EXACT>::type RecT;
	__shared__ __align__(16) RecT sRecAll[RASTER_WARPS][32];
	__shared__ uint32_t sKeysAll[RASTER_WARPS][LOCAL_SORT]; // the tile's command list in submission order (lists up to LOCAL_SORT entries)
	// per (row pair, command): which of the 16 quads of the row pair the command may touch; 16-bit masks, commands c and c + 8 share a word
	__shared__ __align__(16) uint16_t sMaskAll[RASTER_WARPS][32];
	__shared__ __align__(8) uint64_t sBarAll[RASTER_WARPS]; // one mbarrier per warp: completion of the batch's bulk copies
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	RecT *sRec = sRecAll[warp];
	uint16_t *sMask = sMaskAll[warp];
	uint64_t *sBar = &sBarAll[warp];
	uint32_t barParity = 0u;
	constexpr bool BULK = DFPSR_CHK_BULK != 0 && EXACT && MODE != TILE_DEPTH_ONLY;
	if (BULK) {
		if (lane == 0) { mbarrier_init(sBar, 1u); }
		__syncwarp();
	}
	chain_enter();
	if (frame_dropped(frame)) { return; } // the frame did not fit its pools: the host draws it again (see FrameDev::checkCaps)

	// grid = (tile columns of the widest view / RASTER_WARPS, tile rows of the tallest view, views): no division, no search
	ViewDev vw;
	{
		const uint4 *src = (const uint4 *)(frame.views + blockIdx.z);
		uint4 *dst = (uint4 *)&vw;
#pragma unroll
		for (int w = 0; w < (int)(sizeof(ViewDev) / 16); w++) { dst[w] = __ldg(src + w); }
	}
	const int32_t tilesX = vw.tilesX;
	const int32_t tileX = (int32_t)(blockIdx.x * RASTER_WARPS) + warp, tileY = (int32_t)blockIdx.y;
	if (tileX >= tilesX || tileY >= vw.tilesY) { return; }
	const uint32_t tile = vw.tileBase + (uint32_t)(tileY * tilesX + tileX);
	if (tileY * TILE_H >= vw.clipBottom || tileY * TILE_H + TILE_H <= vw.clipTop) { return; }
	const uint32_t n = frame.tileCursor[tile];
	const bool clear = vw.clear != 0;
	if (n == 0 && !clear) { return; }
	if (n == 0) {
		// An empty tile of a cleared target (sky; most tiles of a shadow cube map): nothing to load, sort or shade. Lane = four pixels of
		// one row, one 16-byte store per buffer when the row allows it.
		const int32_t x = tileX * TILE_W + 4 * (lane % (TILE_W / 4)), y = tileY * TILE_H + lane / (TILE_W / 4);
		if (y >= vw.height || x >= vw.width) { return; }
		const int32_t count = min(4, vw.width - x);
		if (!DEPTH_ONLY && vw.color.data != nullptr) {
			uint32_t *dst = row_ptr<uint32_t>(vw.color.data, vw.color.stride, y) + x;
			const uint32_t c = vw.clearColor;
			if (count == 4 && (((uintptr_t)dst) & 15u) == 0) { *(uint4 *)dst = make_uint4(c, c, c, c); }
			else { for (int32_t i = 0; i < count; i++) { dst[i] = c; } }
		}
		if (vw.depth.data != nullptr) {
			float *dst = row_ptr<float>(vw.depth.data, vw.depth.stride, y) + x;
			const float d = vw.clearDepth;
			if (count == 4 && (((uintptr_t)dst) & 15u) == 0) { *(float4 *)dst = make_float4(d, d, d, d); }
			else { for (int32_t i = 0; i < count; i++) { dst[i] = d; } }
		}
		return;
	}
	const uint32_t *list = frame.tileList + frame.tileOffset[tile];

	const int32_t width = vw.width, height = vw.height;
	const int32_t tileLeft = tileX * TILE_W;
	const int32_t qx = lane % QUADS_X, qy = lane / QUADS_X;
	const int32_t x0 = tileLeft + 2 * qx, y1 = tileY * TILE_H + 2 * qy, y2 = y1 + 1;
	const dfpsr_image color = vw.color, depth = vw.depth;
	const bool hasColor = !DEPTH_ONLY && color.data != nullptr, hasDepth = depth.data != nullptr;
	const bool in[4] = {x0 < width && y1 < height, x0 + 1 < width && y1 < height, x0 < width && y2 < height, x0 + 1 < width && y2 < height};
	const uint32_t shifts = vw.packShifts;
	// ref: shader/fillerTemplates.h:286-331 — the last row pair of an odd-height target has no lower row and the reference lets its
	// lower-row pointers repeat the upper row: unclipped quads then read and overwrite lanes 0/1 through lanes 2/3.
	const bool aliasLower = y2 >= height;

	uint32_t col[4];
	float dep[4];
#pragma unroll
	for (int l = 0; l < 4; l++) {
		int32_t px = x0 + (l & 1), py = y1 + (l >> 1);
		col[l] = vw.clearColor; dep[l] = vw.clearDepth;
		if (!clear && in[l]) {
			// deferred mode never reads the colour target: pixels nobody wins keep what they hold (they are not stored)
			if (hasColor && !DEFERRED) { col[l] = row_ptr<uint32_t>(color.data, color.stride, py)[px]; }
			if (hasDepth) { dep[l] = row_ptr<float>(depth.data, depth.stride, py)[px]; }
		}
	}
	bool dirty = clear;
	// deferred mode: the command that is visible at each of the lane's pixels so far, its interpolated U/W and V/W (1/W lives in dep[])
	// and the mip level its quad selected (one byte per pixel)
	uint32_t win[4] = {NO_WINNER, NO_WINNER, NO_WINNER, NO_WINNER};
	float su[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
	uint32_t mips = 0u;
	// the shader pattern (variant flags + diffuse texture) shared by every command that took one of the lane's pixels so far: 0 nothing
	// taken, PATTERN_MIXED different patterns. Conservative for the shading pass's fast path (a pattern that was overwritten still counts).
	uint32_t lanePattern = 0u;

	// lists of up to LOCAL_SORT entries are sorted here and kept in shared memory; longer ones were sorted by sort_lists_kernel
	uint32_t *sKeys = sKeysAll[warp];
	const bool localSort = n <= (uint32_t)LOCAL_SORT;
	if (localSort) {
		uint32_t k0 = (uint32_t)lane < n ? __ldg(list + lane) : 0xFFFFFFFFu;
		uint32_t k1 = (uint32_t)lane + 32u < n ? __ldg(list + 32 + lane) : 0xFFFFFFFFu;
		if (n > 1u) {
			// the set-up pass fills lists roughly in command order: most short lists arrive sorted
			const uint32_t previous0 = __shfl_up_sync(0xffffffffu, k0, 1), last0 = __shfl_sync(0xffffffffu, k0, 31);
			uint32_t previous1 = __shfl_up_sync(0xffffffffu, k1, 1);
			if (lane == 0) { previous1 = last0; }
			if (__any_sync(0xffffffffu, (lane > 0 && previous0 > k0) || previous1 > k1)) {
				if (n <= 8u) { k0 = warp_sort<8>(k0, lane); }
				else if (n <= 16u) { k0 = warp_sort<16>(k0, lane); }
				else if (n <= 32u) { k0 = warp_sort<32>(k0, lane); }
				else { warp_sort64(k0, k1, lane); }
			}
		}
		sKeys[lane] = k0; sKeys[lane + 32] = k1;
		__syncwarp();
	}

	for (uint32_t batchStart = 0; batchStart < n; batchStart += BATCH) {
		const uint32_t batchCount = min((uint32_t)BATCH, n - batchStart);
		// ---- lane (c, r): checkpoint of command c for row pair r of this tile
		const uint32_t c = (uint32_t)lane % (uint32_t)BATCH, r = (uint32_t)lane / (uint32_t)BATCH;
		uint32_t key;
		if (localSort) { key = sKeys[(batchStart + c) & (uint32_t)(LOCAL_SORT - 1)]; }
		else { key = c < batchCount ? __ldg(list + batchStart + c) : 0u; }
		{
			// the record is written straight into shared memory (a local copy that is stored through a uint4 view lives on the stack)
			RecT &rec = sRec[r * BATCH + c];
			int32_t recMode = -1;
			rec.at = 0; rec.ul = rec.ur = rec.ll = rec.lr = 0;
			int32_t quadFirst = 0, quadEnd = 0; // quads [quadFirst, quadEnd) of this row pair can be touched
			const void *stored = nullptr;       // the stored checkpoint of this (command, row pair, tile column), fetched by a bulk copy
			if (c < batchCount) {
				const Cmd *cmd = frame.cmds + key;
				const int4 head = __ldg((const int4 *)cmd + 3); // rowStart, rowCount, rowOffset, pad
				const int32_t rowStart = head.x, rowCount = head.y;
				const uint32_t rowOffset = (uint32_t)head.z;
				const int32_t yTop = tileY * TILE_H + 2 * (int32_t)r;
				const int32_t idx = yTop - rowStart;
				if (idx >= 0 && idx < rowCount) {
					const int4 rr = __ldg((const int4 *)(frame.rows + rowOffset + (uint32_t)idx)); // rows idx and idx + 1 (both even-aligned)
					rec.ul = rr.x; rec.ur = rr.y; rec.ll = rr.z; rec.lr = rr.w;
					if constexpr (!EXACT) {
						const int32_t outerStart = min(rr.x, rr.z), outerEnd = max(rr.y, rr.w);
						const int32_t obs = outerStart & ~1, obe = (outerEnd + 1) & ~1;
						const bool hasTop = rr.y > rr.x, hasBottom = (yTop + 1 < height) && rr.w > rr.z;
						if ((hasTop || hasBottom) && obe > tileLeft && obs < tileLeft + TILE_W) {
							recMode = 4;
							quadFirst = (max(obs, tileLeft) - tileLeft) >> 1; quadEnd = (min(obe, tileLeft + TILE_W) - tileLeft) >> 1;
						}
					} else {
					const uint4 third = __ldg((const uint4 *)cmd + 2); // dy[2], flags, chkOffset, chkShape
					float start[3], dx[3], dy[3];
					{
						const float4 a = __ldg((const float4 *)cmd), b = __ldg((const float4 *)cmd + 1);
						start[0] = a.x; start[1] = a.y; start[2] = a.z; dx[0] = a.w; dx[1] = b.x; dx[2] = b.y; dy[0] = b.z; dy[1] = b.w; dy[2] = __uint_as_float(third.x);
					}
					if (DEPTH_ONLY) {
						// ref: implementation/render/renderCore.cpp:343-387 — per row: value at row.left, then += dx per pixel
						const bool hasU = rec.ur > rec.ul && rec.ur > tileLeft && rec.ul < tileLeft + TILE_W;
						const bool hasL = rec.lr > rec.ll && rec.lr > tileLeft && rec.ll < tileLeft + TILE_W && yTop + 1 < height;
						if (hasU || hasL) {
							recMode = 3;
							const int32_t lo = min(hasU ? rec.ul : 0x7FFFFFFF, hasL ? rec.ll : 0x7FFFFFFF), hi = max(hasU ? rec.ur : -1, hasL ? rec.lr : -1);
							quadFirst = (max(lo, tileLeft) - tileLeft) >> 1; quadEnd = (min(hi, tileLeft + TILE_W) - tileLeft + 1) >> 1;
							float vu = (start[0] + (dx[0] * ((float)rec.ul + 0.5f))) + (dy[0] * ((float)yTop + 0.5f));
							float vl = (start[0] + (dx[0] * ((float)rec.ll + 0.5f))) + (dy[0] * ((float)(yTop + 1) + 0.5f));
							if (hasU) { for (int32_t s = rec.ul; s < tileLeft; s++) { vu += dx[0]; } }
							if (hasL) { for (int32_t s = rec.ll; s < tileLeft; s++) { vl += dx[0]; } }
							rec.v[0] = vu; rec.v[1] = vl;
						}
					} else {
						// ref: shader/fillerTemplates.h:275-300
						const int32_t outerStart = min(rec.ul, rec.ll), outerEnd = max(rec.ur, rec.lr);
						const int32_t innerStart = max(rec.ul, rec.ll), innerEnd = min(rec.ur, rec.lr);
						const int32_t obs = outerStart & ~1, obe = (outerEnd + 1) & ~1, ibs = (innerStart + 1) & ~1, ibe = innerEnd & ~1;
						const bool hasTop = rec.ur > rec.ul, hasBottom = (yTop + 1 < height) && rec.lr > rec.ll;
						if ((hasTop || hasBottom) && obe > tileLeft && obs < tileLeft + TILE_W) {
							// Replay of the reference's running sums (fillerTemplates.h:329-372) from the outer block start to the tile.
							float up[3], lo[3], dx2[3];
							const float fx = (float)obs + 0.5f, fy = (float)yTop + 0.5f;
#pragma unroll
							for (int k = 0; k < 3; k++) {
								up[k] = (start[k] + (dx[k] * fx)) + (dy[k] * fy);
								lo[k] = up[k] + dy[k];
								dx2[k] = dx[k] * 2.0f;
							}
							const int32_t at = max(obs, tileLeft);
							const bool noInner = ibe <= ibs;
							rec.at = at;
							quadFirst = (at - tileLeft) >> 1; quadEnd = (min(obe, tileLeft + TILE_W) - tileLeft) >> 1;
							if (third.z != CHK_NONE && obs < tileLeft) { // the row pair entered the tile from the left: its sums were stored at the tile's edge
								// a set-up warp walked this row pair and left the sums at this tile column
								const uint32_t firstColumn = third.w & 0xFFFFu, columns = third.w >> 16;
								const uint4 *src = (const uint4 *)(frame.chk + third.z + (size_t)(idx >> 1) * columns + ((uint32_t)tileX - firstColumn));
								if (BULK) { stored = src; recMode = 0; } else {
								const uint4 w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2), w3 = __ldg(src + 3), w4 = __ldg(src + 4);
								recMode = (int32_t)w0.x;
								rec.v[0] = __uint_as_float(w0.y); rec.v[1] = __uint_as_float(w0.z); rec.v[2] = __uint_as_float(w0.w);
								rec.v[3] = __uint_as_float(w1.x); rec.v[4] = __uint_as_float(w1.y); rec.v[5] = __uint_as_float(w1.z); rec.v[6] = __uint_as_float(w1.w);
								rec.v[7] = __uint_as_float(w2.x); rec.v[8] = __uint_as_float(w2.y); rec.v[9] = __uint_as_float(w2.z); rec.v[10] = __uint_as_float(w2.w);
								rec.v[11] = __uint_as_float(w3.x); rec.v[12] = __uint_as_float(w3.y); rec.v[13] = __uint_as_float(w3.z); rec.v[14] = __uint_as_float(w3.w);
								rec.v[15] = __uint_as_float(w4.x); rec.v[16] = __uint_as_float(w4.y); rec.v[17] = __uint_as_float(w4.z);
								}
							} else if (noInner || at <= ibs) {
								for (int32_t s = obs; s < at; s += 2) {
#pragma unroll
									for (int k = 0; k < 3; k++) { up[k] += dx2[k]; lo[k] += dx2[k]; }
								}
								recMode = 0;
#pragma unroll
								for (int k = 0; k < 3; k++) { rec.v[k] = up[k]; rec.v[3 + k] = lo[k]; }
							} else {
								for (int32_t s = obs; s < ibs; s += 2) {
#pragma unroll
									for (int k = 0; k < 3; k++) { up[k] += dx2[k]; lo[k] += dx2[k]; }
								}
								const float quadCount = (float)((ibe - ibs) / 2);
								if (at < ibe) {
									float lanes[3][4];
#pragma unroll
									for (int k = 0; k < 3; k++) { lanes[k][0] = up[k]; lanes[k][1] = up[k] + dx[k]; lanes[k][2] = lo[k]; lanes[k][3] = lo[k] + dx[k]; }
									for (int32_t s = ibs; s < at; s += 2) {
#pragma unroll
										for (int k = 0; k < 3; k++) {
#pragma unroll
											for (int l = 0; l < 4; l++) { lanes[k][l] += dx2[k]; }
										}
									}
									recMode = 1;
#pragma unroll
									for (int k = 0; k < 3; k++) {
#pragma unroll
										for (int l = 0; l < 4; l++) { rec.v[k * 4 + l] = lanes[k][l]; }
										rec.v[12 + k] = up[k] + (dx2[k] * quadCount);
										rec.v[15 + k] = lo[k] + (dx2[k] * quadCount);
									}
								} else {
#pragma unroll
									for (int k = 0; k < 3; k++) { up[k] = up[k] + (dx2[k] * quadCount); lo[k] = lo[k] + (dx2[k] * quadCount); }
									for (int32_t s = ibe; s < at; s += 2) {
#pragma unroll
										for (int k = 0; k < 3; k++) { up[k] += dx2[k]; lo[k] += dx2[k]; }
									}
									recMode = 2;
#pragma unroll
									for (int k = 0; k < 3; k++) { rec.v[k] = up[k]; rec.v[3 + k] = lo[k]; }
								}
							}
						}
					}
					}
				}
			}
			if constexpr (BULK) {
				// Stored checkpoints arrive as ONE 80-byte bulk copy per lane (mode, eighteen sums, column: the tail of the record), completion
				// counted on the warp's mbarrier; lanes that computed their checkpoint themselves write their record as before.
				const uint32_t fetching = __ballot_sync(0xffffffffu, stored != nullptr);
				if (fetching != 0u) {
					fence_async_proxy(); // the previous batch read these records through the generic proxy
					if (lane == 0) { mbarrier_expect(sBar, (uint32_t)sizeof(ChkRec) * (uint32_t)__popc(fetching)); }
					if (stored != nullptr) { bulk_copy_to_shared(&rec.mode, stored, (uint32_t)sizeof(ChkRec), sBar); }
					else { rec.mode = recMode; }
					mbarrier_wait(sBar, barParity);
					barParity ^= 1u;
				} else { rec.mode = recMode; }
			} else {
				(void)stored;
				rec.mode = recMode;
			}
			// commands c and c + BATCH / 2 share one 32-bit word of the row pair's masks
			sMask[r * BATCH + 2u * (c % (uint32_t)(BATCH / 2)) + c / (uint32_t)(BATCH / 2)] = (uint16_t)((recMode >= 0 && quadEnd > quadFirst) ? ((0xFFFFFFFFu >> (32 - quadEnd)) & ~((1u << quadFirst) - 1u)) : 0u);
		}
		__syncwarp();
		// Every lane collects the commands of the batch that may touch ITS quad and works through them in submission order. Lanes are
		// independent (a pixel only depends on the commands that cover it), so lanes covered by different triangles shade at the same
		// time instead of idling through each other's commands: the warp needs as many rounds as its busiest quad has commands.
		uint32_t cover = 0;
		{
			constexpr int HALF = BATCH / 2; // words per row pair: word k holds the masks of commands k (low half) and k + HALF (high half)
			const uint4 *masks = (const uint4 *)(sMask + qy * BATCH);
			uint32_t words[HALF];
			{
				const uint4 m0 = masks[0];
				words[0] = m0.x; words[1] = m0.y; words[2] = m0.z; words[3] = m0.w;
				if constexpr (HALF == 8) { const uint4 m1 = masks[1]; words[4] = m1.x; words[5] = m1.y; words[6] = m1.z; words[7] = m1.w; }
			}
			uint32_t both = 0u; // bit k: command k, bit 16 + k: command HALF + k
#pragma unroll
			for (int k = 0; k < HALF; k++) { both |= ((words[k] >> qx) & 0x00010001u) << k; }
			cover = (both & ((1u << HALF) - 1u)) | ((both >> (16 - HALF)) & (((1u << HALF) - 1u) << HALF));
		}

		while (__any_sync(0xffffffffu, cover != 0u)) {
			const bool busy = cover != 0u;
			const uint32_t ci = busy ? (uint32_t)__ffs((int)cover) - 1u : 0u;
			cover &= cover - 1u;
			const uint32_t cmdKey = __shfl_sync(0xffffffffu, key, (int)ci);
			if (!busy) { continue; }
			const RecT &rec = sRec[(uint32_t)qy * BATCH + ci];
			const int32_t mode = rec.mode;
			const Cmd &cmd = frame.cmds[cmdKey];
			int2 upperRow = make_int2(rec.ul, rec.ur), lowerRow = make_int2(rec.ll, rec.lr);

			if constexpr (DEPTH_ONLY) {
				// The reference adds dx once per pixel from the row's left end (renderCore.cpp:343-387). Both rows walk to this quad's first
				// column in ONE loop (the checkpoint holds the sums at max(row.left, tileLeft)); the second column is one more addition.
				const bool affine = (__ldg(&cmd.flags) & CMD_AFFINE) != 0;
				const float dx0 = __ldg(&cmd.dx[0]);
				const int32_t startU = max(upperRow.x, tileLeft), startL = max(lowerRow.x, tileLeft);
				const bool touchU = x0 + 1 >= upperRow.x && x0 < upperRow.y, touchL = x0 + 1 >= lowerRow.x && x0 < lowerRow.y && y2 < height;
				const int32_t stepsU = touchU ? x0 - startU : -1, stepsL = touchL ? x0 - startL : -1; // negative: the row starts right of the first column (or misses the quad)
				float vu = rec.v[0], vl = rec.v[1];
				for (int32_t i = 0, steps = max(stepsU, stepsL); i < steps; i++) {
					if (i < stepsU) { vu += dx0; }
					if (i < stepsL) { vl += dx0; }
				}
				const float value[4] = {vu, stepsU >= 0 ? vu + dx0 : vu, vl, stepsL >= 0 ? vl + dx0 : vl};
#pragma unroll
				for (int l = 0; l < 4; l++) {
					const int2 row = (l < 2) ? upperRow : lowerRow;
					const int32_t px = x0 + (l & 1), py = y1 + (l >> 1);
					if (px >= row.x && px < row.y && py < height) {
						if (affine ? (value[l] < dep[l]) : (value[l] > dep[l])) { dep[l] = value[l]; dirty = true; }
					}
				}
				continue;
			} else {
			const int32_t outerStart = min(upperRow.x, lowerRow.x), outerEnd = max(upperRow.y, lowerRow.y);
			const int32_t innerStart = max(upperRow.x, lowerRow.x), innerEnd = min(upperRow.y, lowerRow.y);
			const int32_t obs = outerStart & ~1, obe = (outerEnd + 1) & ~1, ibs = (innerStart + 1) & ~1, ibe = innerEnd & ~1;
			if (y2 >= height) { lowerRow.y = lowerRow.x; }
			if (x0 < obs || x0 >= obe) { continue; }
			// words 0..2 of the record: start[3], dx[3], dy[3], flags
			const float4 planeA = __ldg((const float4 *)&cmd), planeB = __ldg((const float4 *)&cmd + 1);
			float lanes[3][4];
			bool clipSides;
			uint32_t flags;
			if constexpr (EXACT) {
				flags = __ldg(&cmd.flags);
				const float dx[3] = {planeA.w, planeB.x, planeB.y};
				chain_to_quad(rec, mode, dx, x0, ibs, ibe, lanes, clipSides);
			} else {
				// tolerance mode: the planes at the quad's first pixel centre (ITriangle2D.h:82-100), neighbours by one addition each
				const uint4 planeC = __ldg((const uint4 *)&cmd + 2);
				flags = planeC.y;
				const float start[3] = {planeA.x, planeA.y, planeA.z}, dx[3] = {planeA.w, planeB.x, planeB.y}, dy[3] = {planeB.z, planeB.w, __uint_as_float(planeC.x)};
				const float fx = (float)x0 + 0.5f, fy = (float)y1 + 0.5f;
#pragma unroll
				for (int k = 0; k < 3; k++) {
					lanes[k][0] = __fmaf_rn(dy[k], fy, __fmaf_rn(dx[k], fx, start[k]));
					lanes[k][1] = lanes[k][0] + dx[k];
					lanes[k][2] = lanes[k][0] + dy[k];
					lanes[k][3] = lanes[k][2] + dx[k];
				}
				clipSides = !(ibe > ibs && x0 >= ibs && x0 < ibe);
				(void)mode;
			}
			const bool affine = (flags & CMD_AFFINE) != 0;

			// ref: shader/fillerTemplates.h:93-138 — visibility
			bool vis[4];
			bool anyVisible = false;
			const bool repeatUpper = aliasLower && !clipSides; // lanes 2/3 act on the pixels of lanes 0/1
#pragma unroll
			for (int l = 0; l < 4; l++) {
				bool visible = true;
				if (clipSides) {
					int2 row = (l < 2) ? upperRow : lowerRow;
					int32_t px = x0 + (l & 1);
					visible = px >= row.x && px < row.y;
				}
				if (visible && hasDepth) {
					const float old = (l >= 2 && repeatUpper) ? dep[l - 2] : dep[l];
					visible = affine ? (lanes[0][l] < old) : (lanes[0][l] > old);
				}
				vis[l] = visible;
				anyVisible = anyVisible || visible;
			}
			if (!anyVisible) { continue; }

			if constexpr (DEFERRED) {
				// What the shading pass needs of a pixel a TEXTURED command takes is prepared here, where the perspective division of lanes
				// 0, 1, 2 is done anyway for the quad's mip level (it comes from those lanes of THIS command whether they are visible or
				// not, textureAPI.h:472-495): the texture coordinates (u, v) themselves for plain textured commands (diffuse texture,
				// colourless vertices, no light map: RgbaMultiply.h:75-79), the vertex weights (B, C) for the other textured variants. The
				// shading pass then starts at the sampler instead of repeating division, weights and interpolation per pixel (PREPARED:
				// 37.4 -> 34.4 us per 1080p terrain frame). Commands without any texture keep the raw sums (U/W, V/W): their division at every
				// take costs a frame of tiny vertex-coloured triangles (many takes per pixel) more than the shading pass saves.
				uint32_t mip = 0u;
				float pa[4] = {lanes[1][0], lanes[1][1], lanes[1][2], lanes[1][3]}, pb[4] = {lanes[2][0], lanes[2][1], lanes[2][2], lanes[2][3]};
				if (hasColor && (flags & (CMD_HAS_DIFFUSE | CMD_HAS_LIGHT)) != 0 && (PREPARED || (flags & CMD_HAS_DIFFUSE) != 0)) {
					const bool plainCommand = (flags & (CMD_HAS_DIFFUSE | CMD_HAS_LIGHT | CMD_HAS_FADE | CMD_COLORLESS)) == (CMD_HAS_DIFFUSE | CMD_COLORLESS);
					bool mipNeeded = false;
					TexDev t;
					if ((flags & CMD_HAS_DIFFUSE) != 0) { t = load_tex(frame.textures, (flags >> 8) & 0xFFFu); mipNeeded = t.maxMipLevel > 0u; }
					float cu[3] = {0.0f, 0.0f, 0.0f}, cv[3] = {0.0f, 0.0f, 0.0f};
					if ((PREPARED && plainCommand) || mipNeeded) {
						const float4 t0 = __ldg((const float4 *)&cmd + 7), t1 = __ldg((const float4 *)&cmd + 8);
						cu[0] = t0.x; cu[1] = t0.y; cu[2] = t0.z; cv[0] = t0.w; cv[1] = t1.x; cv[2] = t1.y;
					}
					float u[3] = {0.0f, 0.0f, 0.0f}, v[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
					for (int l = 0; l < 4; l++) {
						const bool forMip = mipNeeded && l < 3;
						if ((PREPARED && vis[l]) || forMip) {
							float wb, wc;
							if (affine) { wb = lanes[1][l]; wc = lanes[2][l]; }
							else { const float linearDepth = reciprocal_w<EXACT>(lanes[0][l]); wb = lanes[1][l] * linearDepth; wc = lanes[2][l] * linearDepth; }
							if (PREPARED) { pa[l] = wb; pb[l] = wc; }
							if ((PREPARED && plainCommand) || forMip) {
								const float wa = 1.0f - (wb + wc);
								const float tu = interpolate3(cu, wa, wb, wc), tv = interpolate3(cv, wa, wb, wc);
								if (l < 3) { u[l] = tu; v[l] = tv; }
								if (PREPARED && plainCommand) { pa[l] = tu; pb[l] = tv; }
							}
						}
					}
					if (mipNeeded) { mip = mip_level(t, u, v); }
				}
				// writes in lane order (clippedWrite); with a repeated upper row lanes 2/3 land on lanes 0/1
#define DFPSR_TAKE(T, L, SEL) { dep[T] = lanes[0][L]; su[T] = pa[L]; sv[T] = pb[L]; win[T] = cmdKey; mips = __byte_perm(mips, mip, SEL); }
				if (vis[0]) DFPSR_TAKE(0, 0, 0x3214)
				if (vis[1]) DFPSR_TAKE(1, 1, 0x3240)
				if (vis[2]) { if (repeatUpper) DFPSR_TAKE(0, 2, 0x3214) else DFPSR_TAKE(2, 2, 0x3410) }
				if (vis[3]) { if (repeatUpper) DFPSR_TAKE(1, 3, 0x3240) else DFPSR_TAKE(3, 3, 0x4210) }
#undef DFPSR_TAKE
				{
					const uint32_t pattern = (flags & SHADER_PATTERN) | PATTERN_TAKEN;
					lanePattern = (lanePattern == 0u || lanePattern == pattern) ? pattern : PATTERN_MIXED;
				}
				dirty = true;
			} else {
			// ref: shader/fillerTemplates.h:196-243 — weights
			float wa[4], wb[4], wc[4];
#pragma unroll
			for (int l = 0; l < 4; l++) {
				if (affine) { wb[l] = lanes[1][l]; wc[l] = lanes[2][l]; }
				else { float linearDepth = 1.0f / lanes[0][l]; wb[l] = lanes[1][l] * linearDepth; wc[l] = lanes[2][l] * linearDepth; }
				wa[l] = 1.0f - (wb[l] + wc[l]);
			}
			if (hasColor) {
				// ref: shader/RgbaMultiply.h:75-106
				float rgba[4][4];
				const bool hasDiffuse = (flags & CMD_HAS_DIFFUSE) != 0, hasLight = (flags & CMD_HAS_LIGHT) != 0;
				const bool fade = (flags & CMD_HAS_FADE) != 0, colorless = (flags & CMD_COLORLESS) != 0 && !fade;
				if (hasDiffuse && !hasLight && colorless) {
					sample_quad<false>(load_tex(frame.textures, (flags >> 8) & 0xFFFu), false, cmd.u1, cmd.v1, wa, wb, wc, rgba);
				} else if (hasLight && !hasDiffuse && colorless) {
					sample_quad<false>(load_tex(frame.textures, flags >> 20), true, cmd.u2, cmd.v2, wa, wb, wc, rgba);
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
					if (hasDiffuse) { sample_quad<true>(load_tex(frame.textures, (flags >> 8) & 0xFFFu), false, cmd.u1, cmd.v1, wa, wb, wc, rgba); }
					if (hasLight) { sample_quad<true>(load_tex(frame.textures, flags >> 20), true, cmd.u2, cmd.v2, wa, wb, wc, rgba); }
				}
				const bool alphaFilter = (flags & CMD_ALPHA) != 0;
				uint32_t packed[4];
#pragma unroll
				for (int l = 0; l < 4; l++) {
					if (alphaFilter) {
						// ref: shader/fillerTemplates.h:155-176; all four reads precede the writes; lanes that are not visible read 0 when clipping sides
						float opacity = rgba[l][3] * (1.0f / 255.0f);
						uint32_t target = (vis[l] || !clipSides) ? ((l >= 2 && repeatUpper) ? col[l - 2] : col[l]) : 0u;
						float inv = 1.0f - opacity;
						float tr = (float)((target >> (shifts & 31u)) & 255u), tg = (float)((target >> ((shifts >> 8) & 31u)) & 255u);
						float tb = (float)((target >> ((shifts >> 16) & 31u)) & 255u), ta = (float)((target >> ((shifts >> 24) & 31u)) & 255u);
						rgba[l][0] = (rgba[l][0] * opacity) + (tr * inv);
						rgba[l][1] = (rgba[l][1] * opacity) + (tg * inv);
						rgba[l][2] = (rgba[l][2] * opacity) + (tb * inv);
						rgba[l][3] = (rgba[l][3] * opacity) + (ta * inv);
					}
					packed[l] = pack_rgba_ordered(saturated_byte(rgba[l][0]), saturated_byte(rgba[l][1]), saturated_byte(rgba[l][2]), saturated_byte(rgba[l][3]), shifts);
				}
				// writes in lane order (clippedWrite); with a repeated upper row lanes 2/3 land on lanes 0/1
				if (vis[0]) { col[0] = packed[0]; }
				if (vis[1]) { col[1] = packed[1]; }
				if (vis[2]) { if (repeatUpper) { col[0] = packed[2]; } else { col[2] = packed[2]; } }
				if (vis[3]) { if (repeatUpper) { col[1] = packed[3]; } else { col[3] = packed[3]; } }
				dirty = true;
				// ref: shader/fillerTemplates.h:387-441 — alpha filtering leaves depth untouched when both buffers exist
				if (hasDepth && !alphaFilter) {
					if (vis[0]) { dep[0] = lanes[0][0]; }
					if (vis[1]) { dep[1] = lanes[0][1]; }
					if (vis[2]) { if (repeatUpper) { dep[0] = lanes[0][2]; } else { dep[2] = lanes[0][2]; } }
					if (vis[3]) { if (repeatUpper) { dep[1] = lanes[0][3]; } else { dep[3] = lanes[0][3]; } }
				}
			} else if (hasDepth) {
				if (vis[0]) { dep[0] = lanes[0][0]; }
				if (vis[1]) { dep[1] = lanes[0][1]; }
				if (vis[2]) { if (repeatUpper) { dep[0] = lanes[0][2]; } else { dep[2] = lanes[0][2]; } }
				if (vis[3]) { if (repeatUpper) { dep[1] = lanes[0][3]; } else { dep[3] = lanes[0][3]; } }
				dirty = true;
			}
			}
			}
		}
		__syncwarp();
	}

	if constexpr (DEFERRED) {
		// ---- shading pass: every pixel that some command won is shaded once, by that command
		if (hasColor) {
			const bool valid[4] = {win[0] != NO_WINNER, win[1] != NO_WINNER, win[2] != NO_WINNER, win[3] != NO_WINNER};
			// Fast path, chosen by the whole warp: every command that took a pixel of the warp's lanes is a plain textured command (diffuse
			// texture, no light map, colourless vertices: RgbaMultiply.h:75-79) and each lane saw one texture only (lanePattern, kept by the
			// visibility pass: no flag loads here). The four pixels are then sampled side by side — sixteen texel loads in flight — instead
			// of one shader variant dispatch per pixel.
			const uint32_t common = lanePattern;
			const bool plain = common == 0u || (common != PATTERN_MIXED && (common & (CMD_HAS_DIFFUSE | CMD_HAS_LIGHT | CMD_HAS_FADE | CMD_COLORLESS)) == (CMD_HAS_DIFFUSE | CMD_COLORLESS));
			if (__all_sync(0xffffffffu, plain)) {
				if (common != 0u) {
					const TexDev t = load_tex(frame.textures, (common >> 8) & 0xFFFu);
					if (!PREPARED) {
						// the visibility pass left the raw sums (U/W, V/W): division, vertex weights and texture coordinates per pixel
						float4 t0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), t1 = t0;
						uint32_t current = NO_WINNER;
#pragma unroll
						for (int l = 0; l < 4; l++) {
							if (valid[l]) {
								if (win[l] != current) { // texture coordinates of the command: words 7 and 8 of its record
									current = win[l];
									const float4 *words = (const float4 *)(frame.cmds + current);
									t0 = __ldg(words + 7); t1 = __ldg(words + 8);
								}
								float wb, wc;
								if (common & CMD_AFFINE) { wb = su[l]; wc = sv[l]; }
								else { const float linearDepth = reciprocal_w<EXACT>(dep[l]); wb = su[l] * linearDepth; wc = sv[l] * linearDepth; }
								const float wa = 1.0f - (wb + wc);
								const float cu[3] = {t0.x, t0.y, t0.z}, cv[3] = {t0.w, t1.x, t1.y};
								su[l] = interpolate3(cu, wa, wb, wc); sv[l] = interpolate3(cv, wa, wb, wc);
							}
						}
					}
					uint32_t texel[4];
#pragma unroll
					for (int l = 0; l < 4; l++) { texel[l] = sample_bilinear(t, su[l], sv[l], (mips >> (8 * l)) & 0xFFu); } // (u, v) by now; pixels nobody won sample (0, 0) and are dropped
					// byte -> float -> min(x, 255.1) -> truncation (RgbaMultiply.h:75-79, PackOrder.h:186-213) is the identity on 0..255: the
					// texel's bytes only move to the target's pack order, one byte permute per pixel
					const uint32_t selector = vw.packSelector;
#pragma unroll
					for (int l = 0; l < 4; l++) { if (valid[l]) { col[l] = __byte_perm(texel[l], 0u, selector); } }
				}
			} else {
#pragma unroll
				for (int l = 0; l < 4; l++) {
					if (valid[l]) { col[l] = shade_pixel<EXACT, PREPARED>(frame.cmds + win[l], frame.textures, dep[l], su[l], sv[l], (mips >> (8 * l)) & 0xFFu, shifts); }
				}
			}
		}
	}

	if (dirty) {
		// each lane owns 2 adjacent pixels in two rows: 8-byte stores (when the row is 8-byte aligned), 128 contiguous bytes per row per half-warp
		if (hasColor) {
			// deferred mode without a clear: pixels nobody won were never loaded and are not stored
			const bool keep = DEFERRED && !clear;
			const bool s0 = in[0] && !(keep && win[0] == NO_WINNER), s1 = in[1] && !(keep && win[1] == NO_WINNER);
			const bool s2 = in[2] && !(keep && win[2] == NO_WINNER), s3 = in[3] && !(keep && win[3] == NO_WINNER);
			uint32_t *upper = row_ptr<uint32_t>(color.data, color.stride, y1) + x0, *lower = row_ptr<uint32_t>(color.data, color.stride, y2) + x0;
			if (s0 && s1 && (((uintptr_t)upper) & 7u) == 0) { *(uint2 *)upper = make_uint2(col[0], col[1]); }
			else { if (s0) { upper[0] = col[0]; } if (s1) { upper[1] = col[1]; } }
			if (s2 && s3 && (((uintptr_t)lower) & 7u) == 0) { *(uint2 *)lower = make_uint2(col[2], col[3]); }
			else { if (s2) { lower[0] = col[2]; } if (s3) { lower[1] = col[3]; } }
		}
		if (hasDepth) {
			float *upper = row_ptr<float>(depth.data, depth.stride, y1) + x0, *lower = row_ptr<float>(depth.data, depth.stride, y2) + x0;
			if (in[0] && in[1] && (((uintptr_t)upper) & 7u) == 0) { *(float2 *)upper = make_float2(dep[0], dep[1]); }
			else { if (in[0]) { upper[0] = dep[0]; } if (in[1]) { upper[1] = dep[1]; } }
			if (in[2] && in[3] && (((uintptr_t)lower) & 7u) == 0) { *(float2 *)lower = make_float2(dep[2], dep[3]); }
			else { if (in[2]) { lower[0] = dep[2]; } if (in[3]) { lower[1] = dep[3]; } }
		}
	}
}

// ------------------------------------------------------------------------------------------------ wireframe overlay

// ref: api/rendererAPI.cpp:362-399 — renderer_end(renderer, debugWireframe = true): after the frame is drawn, the three edges of every
// command the occlusion grid did not remove are drawn on top of the colour buffer as white lines (draw_line between the corners'
// positions in whole pixels, flat / unitsPerPixel). One thread per input triangle repeats the set-up pass's decisions and walks its lines;
// every line has the same colour, so the order of the writes does not matter. A debug view: nothing here is tuned.
__device__ void wireframe_line(const ViewDev &view, int32_t x1, int32_t y1, int32_t x2, int32_t y2) {
	LineParams p;
	if (!line_params(view.width, view.height, x1, y1, x2, y2, p)) { return; }
	for (int32_t i = 0; i < p.steps; i++) {
		int32_t x, y;
		if (line_pixel(p, i, view.width, view.height, x, y) && y >= view.clipTop && y < view.clipBottom) { row_ptr<uint32_t>(view.color.data, view.color.stride, y)[x] = 0xFFFFFFFFu; } // white, alpha 255, in every pack order
	}
}

__global__ void __launch_bounds__(SETUP_THREADS) wireframe_kernel(FrameDev frame) {
	__shared__ TaskParams task;
	__shared__ int sCulled;
	{
		int32_t t = frame.blockTask ? frame.blockTask[blockIdx.x] : task_of_block(frame.tasks, frame.taskCount, (int32_t)blockIdx.x);
		for (uint32_t w = threadIdx.x; w < sizeof(TaskParams) / 4; w += blockDim.x) { ((uint32_t *)&task)[w] = ((const uint32_t *)&frame.tasks[t])[w]; }
	}
	__syncthreads();
	if (threadIdx.x == 0) { sCulled = task.cullOnDevice && task_is_culled(frame, task) ? 1 : 0; }
	__syncthreads();
	if (sCulled) { return; }
	const ViewDev &view = frame.views[task.view];
	if (view.color.data == nullptr) { return; }
	const int32_t local = ((int32_t)blockIdx.x - task.blockBase) * SETUP_THREADS + (int32_t)threadIdx.x;
	PPoint p[3];
	float colors[3][4], tex[3][4];
	if (local >= task.slotCount || !load_triangle(task, local, p, colors, tex)) { return; }
	const float alpha[3] = {colors[0][3], colors[1][3], colors[2][3]};
	for_each_command(task, p, alpha, [&](const PPoint *q, const float *, const float *) {
		const Bound bound = raster_bound(q, view.width, view.clipTop, view.clipBottom);
		if (command_occluded(frame, bound, q)) { return; }
		int32_t x[3], y[3];
		for (int k = 0; k < 3; k++) { x[k] = (int32_t)(q[k].fx / 256); y[k] = (int32_t)(q[k].fy / 256); }
		wireframe_line(view, x[0], y[0], x[1], y[1]);
		wireframe_line(view, x[1], y[1], x[2], y[2]);
		wireframe_line(view, x[2], y[2], x[0], y[0]);
	});
}

} // namespace dfpsr

// ------------------------------------------------------------------------------------------------ host side

using namespace dfpsr;

// the four instances of the tile kernel under the names the launch accounting and the profiles use
static constexpr auto tile_kernel_depth = raster_kernel<TILE_DEPTH_ONLY, true>;
static constexpr auto tile_kernel_immediate = raster_kernel<TILE_IMMEDIATE, true>;
static constexpr auto tile_kernel_deferred = raster_kernel<TILE_DEFERRED, true>;
static constexpr auto tile_kernel_tolerance = raster_kernel<TILE_DEFERRED, false>;
static constexpr auto tile_kernel_deferred_raw = raster_kernel<TILE_DEFERRED_RAW, true>; // exact frames without any texture: nothing to prepare, and the leaner kernel is 12 % faster on 2 M tiny triangles

// A frame whose second half was launched without waiting for its counts (see renderer_end_internal): everything that is needed to
// draw it again should the counting pass report that it did not fit the pools.
struct PendingFrame {
	bool active = false;
	uint32_t serial = 0;
	int slot = 0;
	cudaStream_t stream = nullptr;
	std::vector<ViewDev> views;
	std::vector<TaskParams> tasks;
	std::vector<TexDev> textures;
	std::vector<float> grid, gridSnapshots;
	bool depthOnly = false, occluded = false, exact = true;
	int32_t gridWidth = 0, gridHeight = 0, gridAllocW = 0, gridAllocH = 0;
	int32_t slots = 0;
};

struct dfpsr_renderer {
	bool receiving = false;
	bool depthOnly = false;
	std::vector<ViewDev> views;
	std::vector<TaskParams> tasks;
	std::vector<DeviceBuffer> uploads;   // host triangle batches of this frame
	size_t uploadCount = 0;
	std::vector<TexDev> textures;        // the frame's texture table (uploaded behind the task records)
	bool exact = true;                   // false: tolerance mode (dfpsr_renderer_set_precision)
	bool wireframe = false;              // the next renderer_end draws the commands' edges on top of the frame (dfpsr_renderer_set_debug_wireframe)
	bool async = false;                  // true: renderer_end does not wait for the frame's counts (dfpsr_renderer_set_async)
	int64_t lastCommands = -1;
	DeviceBuffer dTasks, dViews, projected, slotCounts, blockCmds, blockRows, tileCount, tileOffset, tileCursor, cmds, rows, tileList, chk, sortTmp, bigItems, bigUnits;
	uint32_t *hostTotals = nullptr, *hostTotalsDevice = nullptr; // mapped pinned memory and its device alias: two slots of 16 words
	cudaEvent_t counted[2] = {nullptr, nullptr};                 // recorded behind counts_kernel of the frame that uses the slot
	uint32_t serial = 0;
	PendingFrame pending;                                        // at most one frame is unverified at any time
	uint32_t history[TOTAL_WORDS] = {};                          // the totals of the last verified frame, the pools are sized from them
	int64_t historySlots = 0, historyTiles = 0;                  // 0: no history yet, the next frame waits for its counts
	uint32_t recentUnits = 0; int64_t recentSlots = 1;           // the last verified frame itself: sizes the grid of big_units_kernel
	cudaStream_t lastStream = nullptr;
	bool usedStream = false;
	void *pinnedTasks = nullptr;                                 // page-locked staging for the task records (a pageable source of a few hundred
	size_t pinnedTasksCapacity = 0;                              // KB makes cudaMemcpyAsync wait for everything queued on the stream before it)
	// occlusion grid (ref: api/rendererAPI.cpp:145, :181-192): lives on the host, where occluder boxes and visibility queries are evaluated
	std::vector<float> grid;
	int32_t gridWidth = 0, gridHeight = 0, gridAllocW = 0, gridAllocH = 0;
	bool occluded = false;
	DeviceBuffer dGrid;
	std::vector<float> gridSnapshots; // the grid at every dfpsr_renderer_give_tasks call of the frame that found occluders (device broad phase)
	DeviceBuffer dGridSnapshots;

	~dfpsr_renderer();
};

static bool image_exists(const dfpsr_image *image) { return image != nullptr && image->data != nullptr; }

static int register_texture(dfpsr_renderer *r, const dfpsr_texture *t) {
	if (t == nullptr || t->data == nullptr) { return -1; }
	// most recently registered first: consecutive submissions usually share their textures
	for (size_t n = r->textures.size(), i = n; i-- > 0;) {
		const TexDev &d = r->textures[i];
		if (d.data == t->data && d.log2width == (uint32_t)t->log2width && d.log2height == (uint32_t)t->log2height && d.maxMipLevel == (uint32_t)t->maxMipLevel
		    && d.startOffset == t->startOffset && d.maxLevelMask == t->maxLevelMask) { return (int)i; }
		if (n - i >= 64) { break; } // a bounded look-back keeps submission O(1); duplicates in the table are harmless
	}
	if (r->textures.size() >= (size_t)MAX_TEXTURES) { return -2; }
	TexDev d;
	memset(&d, 0, sizeof(d));
	d.data = t->data; d.log2width = (uint32_t)t->log2width; d.log2height = (uint32_t)t->log2height;
	d.maxMipLevel = (uint32_t)t->maxMipLevel; d.startOffset = t->startOffset; d.maxLevelMask = t->maxLevelMask;
	r->textures.push_back(d);
	return (int)r->textures.size() - 1;
}

// ref: api/rendererAPI.cpp:151-168 — one view
static int make_view(ViewDev &v, const dfpsr_image *color, const dfpsr_image *depth, bool clear, uint32_t clearColor, float clearDepth) {
	memset(&v, 0, sizeof(v));
	if (image_exists(color)) { v.color = *color; }
	if (image_exists(depth)) { v.depth = *depth; }
	if (image_exists(color) && image_exists(depth)) {
		DFPSR_REQUIRE(color->width == depth->width && color->height == depth->height, "renderer_begin: colour buffer %dx%d and depth buffer %dx%d differ", color->width, color->height, depth->width, depth->height);
	}
	if (image_exists(color)) { v.width = color->width; v.height = color->height; }
	else if (image_exists(depth)) { v.width = depth->width; v.height = depth->height; }
	DFPSR_REQUIRE(v.width >= 0 && v.height >= 0, "renderer_begin: negative image dimensions");
	v.clipTop = 0; v.clipBottom = v.height;
	v.clear = clear ? 1 : 0; v.clearColor = clearColor; v.clearDepth = clearDepth;
	v.tilesX = (v.width + TILE_W - 1) / TILE_W;
	v.tilesY = (v.height + TILE_H - 1) / TILE_H;
	v.packShifts = pack_shifts(v.color.packOrder); v.packSelector = pack_selector(v.packShifts);
	return 0;
}

static int verify_frame(dfpsr_renderer *r);

static int renderer_begin_internal(dfpsr_renderer *r, bool depthOnly) {
	DFPSR_REQUIRE(!r->receiving, "Called renderer_begin on the same renderer twice without ending the previous batch!");
	// the previous frame is verified before anything of this one is queued: should it have to be drawn again, it still comes first
	if (verify_frame(r)) { return 1; }
	r->receiving = true;
	r->depthOnly = depthOnly;
	r->occluded = false;
	r->views.clear();
	r->tasks.clear();
	r->uploadCount = 0;
	r->textures.clear();
	r->gridSnapshots.clear();
	if (!r->hostTotals) {
		DFPSR_CHECK_CUDA(cudaHostAlloc((void **)&r->hostTotals, 2 * 16 * sizeof(uint32_t), cudaHostAllocMapped));
		DFPSR_CHECK_CUDA(cudaHostGetDevicePointer((void **)&r->hostTotalsDevice, r->hostTotals, 0));
		for (int i = 0; i < 2; i++) { DFPSR_CHECK_CUDA(cudaEventCreateWithFlags(&r->counted[i], cudaEventDisableTiming)); }
	}
	return 0;
}

static int add_model_task(dfpsr_renderer *r, int32_t view, const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera) {
	const ViewDev &v = r->views[(size_t)view];
	if (v.width <= 0 || v.height <= 0) { return 0; } // ref: renderCore.cpp:307-309 — no target, nothing to draw
	// ref: api/modelAPI.cpp:228 — whole-model culling against the cull frustum on the host
	if (!dfpsr_camera_is_box_seen(camera, model->minBound, model->maxBound, modelToWorld)) { return 0; }
	if (model->polygonCount <= 0) { return 0; }
	const int32_t diffuseIndex = r->depthOnly ? -1 : register_texture(r, &model->diffuse);
	const int32_t lightIndex = r->depthOnly ? -1 : register_texture(r, &model->light);
	DFPSR_REQUIRE(diffuseIndex != -2 && lightIndex != -2, "more than %d textures in one frame", MAX_TEXTURES);
	r->tasks.emplace_back(); // value-initialised (all zero) and filled in place: a Sandbox frame queues hundreds of 400-byte tasks
	TaskParams &task = r->tasks.back();
	task.points = model->points;
	task.polygons = model->polygons;
	task.pointCount = model->pointCount;
	task.slotCount = model->polygonCount * 2;
	task.view = view;
	task.modelToWorld = *modelToWorld;
	task.camera = *camera;
	task.filter = model->filter;
	task.depthOnly = r->depthOnly ? 1 : 0;
	task.diffuseIndex = diffuseIndex;
	task.lightIndex = lightIndex;
	task.gridOffset = -1; // tested on the host by the caller
	return 0;
}

// Gives every task its slot, block and projected-point ranges (tasks own contiguous ranges in submission order).
static int layout_tasks(dfpsr_renderer *r, int32_t &slotTotal, int32_t &blockTotal) {
	slotTotal = 0; blockTotal = 0;
	size_t pointTotal = 0;
	for (TaskParams &t : r->tasks) {
		t.slotBase = slotTotal; t.blockBase = blockTotal;
		t.blockCount = (t.slotCount + SETUP_THREADS - 1) / SETUP_THREADS;
		slotTotal += t.slotCount; blockTotal += t.blockCount;
		if (t.triangles == nullptr) { pointTotal += (size_t)t.pointCount; }
	}
	if (r->projected.reserve(pointTotal * sizeof(PPoint) + 16)) { return 1; }
	size_t at = 0;
	for (TaskParams &t : r->tasks) {
		if (t.triangles == nullptr) { t.projected = (PPoint *)r->projected.ptr + at; at += (size_t)t.pointCount; }
	}
	return 0;
}

static int launch_projection(dfpsr_renderer *r, const FrameDev &frame, cudaStream_t stream) {
	int32_t maxPoints = 0;
	for (const TaskParams &t : r->tasks) { if (t.triangles == nullptr && t.pointCount > maxPoints) { maxPoints = t.pointCount; } }
	if (maxPoints > 0) {
		dim3 grid((unsigned)((maxPoints + 255) / 256), (unsigned)r->tasks.size());
		if (grid.x > 1024u) { grid.x = 1024u; }
		DFPSR_REQUIRE(r->tasks.size() <= 65535, "more than 65535 tasks in one frame");
		DFPSR_LAUNCH(project_kernel, grid, 256, 0, stream, frame.tasks, (int32_t)r->tasks.size(), (uint4 *)nullptr, 0u);
	}
	return 0;
}

static double host_now_us() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }

// ---- frames in flight whose counts the host has not looked at yet (asynchronous renderers, at most one frame each)
namespace dfpsr {
thread_local int g_pendingFrames = 0;
thread_local bool g_insideFrame = false;
}
static thread_local std::vector<dfpsr_renderer *> g_pendingRenderers;

static void forget_pending(dfpsr_renderer *r) {
	for (size_t i = 0; i < g_pendingRenderers.size(); i++) {
		if (g_pendingRenderers[i] == r) { g_pendingRenderers.erase(g_pendingRenderers.begin() + (long)i); g_pendingFrames--; break; }
	}
}

dfpsr_renderer::~dfpsr_renderer() {
	if (pending.active) { cudaEventSynchronize(counted[pending.slot]); forget_pending(this); }
	for (auto &b : uploads) { b.release(); }
	DeviceBuffer *all[] = {&dTasks, &dViews, &projected, &slotCounts, &blockCmds, &blockRows, &tileCount, &tileOffset, &tileCursor, &cmds, &rows, &tileList, &chk, &sortTmp, &bigItems, &bigUnits, &dGrid, &dGridSnapshots};
	for (auto *b : all) { b->release(); }
	if (hostTotals) { cudaFreeHost(hostTotals); }
	for (cudaEvent_t e : counted) { if (e) { cudaEventDestroy(e); } }
	if (pinnedTasks) { cudaFreeHost(pinnedTasks); }
}

struct InsideFrame { // the launches of a frame must not trigger the verification hook of DFPSR_LAUNCH
	bool previous;
	InsideFrame() : previous(g_insideFrame) { g_insideFrame = true; }
	~InsideFrame() { g_insideFrame = previous; }
};

// Grows the pools of the second half so that `needed` elements fit with `slack` on top. Growing frees the old allocation, which makes
// the device idle first (cudaFree), so nothing in flight can still be using it.
static int grow_pool(DeviceBuffer &buffer, size_t needed, size_t elementSize, double slack, size_t extraBytes) {
	if (needed * elementSize + extraBytes <= buffer.capacity) { return 0; }
	return buffer.reserve((size_t)((double)needed * slack) * elementSize + extraBytes);
}
static uint32_t pool_elements(const DeviceBuffer &buffer, size_t elementSize, size_t extraBytes) {
	const size_t n = buffer.capacity > extraBytes ? (buffer.capacity - extraBytes) / elementSize : 0;
	return (uint32_t)std::min<size_t>(n, 0xFFFFFFFFu);
}

static int run_frame(dfpsr_renderer *r, cudaStream_t stream, bool allowAsync) {
	// ref: api/rendererAPI.cpp:352-402
	InsideFrame guard;
	static const bool timing = getenv("DFPSR_END_TIMING") != nullptr; // developer aid: host-side phase times of one renderer_end on stderr
	const double tStart = timing ? host_now_us() : 0.0;
	double tUploaded = 0.0, tLaunched = 0.0, tSynced = 0.0;
	r->lastCommands = 0;
	// scratch buffers are reused from frame to frame in stream order: a frame on another stream first waits for the previous one
	if (r->usedStream && r->lastStream != stream) { DFPSR_CHECK_CUDA(cudaStreamSynchronize(r->lastStream)); }
	// ---- lay out the batch
	uint32_t tileTotal = 0;
	bool anyClear = false;
	for (ViewDev &v : r->views) {
		v.tileBase = tileTotal;
		tileTotal += (uint32_t)(v.tilesX * v.tilesY);
		anyClear = anyClear || (v.clear != 0 && v.width > 0 && v.height > 0);
	}
	if (tileTotal == 0 || (r->tasks.empty() && !anyClear)) { return 0; }
	r->lastStream = stream; r->usedStream = true;
	int32_t slotTotal = 0, blockTotal = 0;
	if (layout_tasks(r, slotTotal, blockTotal)) { return 1; }
	const size_t taskCount = r->tasks.size(), viewCount = r->views.size();
	const size_t counterWords = ((size_t)tileTotal + TOTAL_WORDS + 3) & ~(size_t)3; // tile counts + totals, cleared with 16-byte stores
	if (r->tileCount.reserve(counterWords * 4) || r->tileOffset.reserve(((size_t)tileTotal + 1) * 4) || r->tileCursor.reserve(((size_t)tileTotal + 1) * 4)) { return 1; }
	if (r->slotCounts.reserve((size_t)slotTotal * 4 + 16) || r->blockCmds.reserve((size_t)blockTotal * 4 + 16) || r->blockRows.reserve((size_t)blockTotal * 4 + 16)) { return 1; }
	// ---- one upload: views, task records, the frame's texture table and the task of every set-up block (so that a block of a frame with
	// hundreds of tasks — the shadow pass of a Sandbox frame — starts with one load instead of a ten-step binary search of dependent
	// loads). The records go through page-locked staging: the copy is then a plain DMA in stream order. The staging buffer is free again
	// when the frame's counting pass has run, which the host knows before it lays out the next frame (verify_frame).
	const size_t viewBytes = viewCount * sizeof(ViewDev);
	const size_t taskBytes = (taskCount * sizeof(TaskParams) + 15) & ~(size_t)15; // the texture table is read with 16-byte loads
	const size_t textureBytes = r->textures.size() * sizeof(TexDev);
	const size_t tableBytes = taskCount > 1 ? (size_t)blockTotal * sizeof(int32_t) : 0;
	const size_t bytes = viewBytes + taskBytes + textureBytes + tableBytes;
	if (r->dTasks.reserve(bytes + 16)) { return 1; }
	if (bytes > r->pinnedTasksCapacity) {
		if (r->pinnedTasks) { cudaFreeHost(r->pinnedTasks); r->pinnedTasks = nullptr; r->pinnedTasksCapacity = 0; }
		const size_t grown = bytes + bytes / 2 + 4096;
		DFPSR_CHECK_CUDA(cudaHostAlloc(&r->pinnedTasks, grown, cudaHostAllocDefault));
		r->pinnedTasksCapacity = grown;
	}
	{
		uint8_t *staging = (uint8_t *)r->pinnedTasks;
		memcpy(staging, r->views.data(), viewBytes);
		if (taskCount > 0) { memcpy(staging + viewBytes, r->tasks.data(), taskCount * sizeof(TaskParams)); }
		if (textureBytes > 0) { memcpy(staging + viewBytes + taskBytes, r->textures.data(), textureBytes); }
		if (tableBytes > 0) {
			int32_t *table = (int32_t *)(staging + viewBytes + taskBytes + textureBytes);
			int32_t index = 0;
			for (const TaskParams &t : r->tasks) { for (int32_t b = 0; b < t.blockCount; b++) { table[t.blockBase + b] = index; } index++; }
		}
		DFPSR_CHECK_CUDA(cudaMemcpyAsync(r->dTasks.ptr, staging, bytes, cudaMemcpyHostToDevice, stream));
	}
	const uint8_t *deviceRecords = (const uint8_t *)r->dTasks.ptr;
	if (timing) { tUploaded = host_now_us(); }

	FrameDev frame;
	memset(&frame, 0, sizeof(frame));
	frame.views = (const ViewDev *)deviceRecords;
	frame.tasks = (const TaskParams *)(deviceRecords + viewBytes);
	frame.textures = (const TexDev *)(deviceRecords + viewBytes + taskBytes);
	frame.blockTask = tableBytes > 0 ? (const int32_t *)(deviceRecords + viewBytes + taskBytes + textureBytes) : nullptr;
	// Which tile kernel draws the frame: alpha-filtered commands blend in submission order and need the immediate kernel; frames of solid
	// commands take the deferred one (visibility first, every pixel shaded once), in exact or in tolerance mode.
	bool anyAlpha = false;
	for (const TaskParams &t : r->tasks) { anyAlpha = anyAlpha || t.filter == DFPSR_FILTER_ALPHA; }
	static const char *tileModeOverride = getenv("DFPSR_TILE_MODE"); // developer aid: "immediate" forces the round-1 kernel for A/B timing
	const bool immediate = anyAlpha || (tileModeOverride && strcmp(tileModeOverride, "immediate") == 0);
	const bool exactFrame = r->depthOnly || immediate || r->exact;
	frame.checkpoints = exactFrame ? 1 : 0;
	frame.taskCount = (int32_t)taskCount; frame.viewCount = (int32_t)viewCount; frame.blockCount = blockTotal;
	frame.tileTotal = tileTotal;
	frame.slotCounts = (uint32_t *)r->slotCounts.ptr;
	frame.blockCmds = (uint32_t *)r->blockCmds.ptr; frame.blockRows = (uint32_t *)r->blockRows.ptr;
	frame.tileCount = (uint32_t *)r->tileCount.ptr; frame.tileOffset = (uint32_t *)r->tileOffset.ptr; frame.tileCursor = (uint32_t *)r->tileCursor.ptr;
	frame.totals = frame.tileCount + tileTotal;
	// A frame that does not fill the machine (one 1080p terrain frame: 7.6 k slots) is bound by its longest thread: only triangles within one
	// tile row stay with their set-up thread, everything taller goes to the unit queue (one thread per row pair). Large frames keep the
	// serial scan conversion of triangles up to SMALL_ROWS rows: measured on the 256-view batch 43.45 us per frame at 4 rows, 43.6 at 8,
	// 44.0 at 16, 44.7 at 32 — but the 2 M tiny triangles of BASELINE config 3 (2-3 pixels tall, up to 6 aligned rows) take 930 us at 4
	// (a third of them queue a unit each) against 620 us at 8. DFPSR_SMALL_ROWS overrides for experiments.
	static const int smallRowsOverride = getenv("DFPSR_SMALL_ROWS") ? atoi(getenv("DFPSR_SMALL_ROWS")) : -1;
	frame.smallRows = slotTotal <= sm_count() * 1024 ? 4 : SMALL_ROWS;
	if (smallRowsOverride >= 0) { frame.smallRows = smallRowsOverride; }

	if (taskCount == 0) {
		// nothing but clears: every tile list is empty
		DFPSR_CHECK_CUDA(cudaMemsetAsync(r->tileCursor.ptr, 0, (size_t)tileTotal * 4, stream));
	} else {
		// ---- counting pass: projection (+ clearing the counters), culling / clipping / bounding boxes, scans and tile segments
		int32_t maxPoints = 0;
		for (const TaskParams &t : r->tasks) { if (t.triangles == nullptr && t.pointCount > maxPoints) { maxPoints = t.pointCount; } }
		{
			unsigned gx = (unsigned)std::max(1, (maxPoints + 255) / 256);
			if (gx > 1024u) { gx = 1024u; }
			const size_t zeroCtas = (counterWords / 4 + 255) / 256;
			const size_t zeroRows = (zeroCtas + gx - 1) / gx;
			DFPSR_REQUIRE(taskCount + zeroRows <= 65535, "more than 65535 tasks in one frame");
			const dim3 grid(gx, (unsigned)(taskCount + zeroRows));
			DFPSR_LAUNCH_CHAINED(project_kernel, grid, 256, 0, stream, frame.tasks, (int32_t)taskCount, (uint4 *)r->tileCount.ptr, (uint32_t)counterWords);
		}
		if (r->occluded) {
			// completeOcclusion (ref: api/rendererAPI.cpp:193-217) happens inside the set-up kernels, against the grid as it is now
			if (r->dGrid.reserve(r->grid.size() * sizeof(float) + 16)) { return 1; }
			DFPSR_CHECK_CUDA(cudaMemcpyAsync(r->dGrid.ptr, r->grid.data(), r->grid.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
			frame.occlusionGrid = (const float *)r->dGrid.ptr;
			frame.gridWidth = r->gridWidth; frame.gridHeight = r->gridHeight; frame.gridStride = r->gridAllocW;
		}
		if (!r->gridSnapshots.empty()) { // the broad phase of dfpsr_renderer_give_tasks tests against the grid as it was at each of those calls
			if (r->dGridSnapshots.reserve(r->gridSnapshots.size() * sizeof(float) + 16)) { return 1; }
			DFPSR_CHECK_CUDA(cudaMemcpyAsync(r->dGridSnapshots.ptr, r->gridSnapshots.data(), r->gridSnapshots.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
			frame.gridSnapshots = (const float *)r->dGridSnapshots.ptr;
			frame.gridStride = r->gridAllocW; frame.gridRows = r->gridAllocH;
		}
		// ---- pools of the second half. Without history (first frame, synchronous renderer) the host waits for the counts and sizes the
		// pools exactly; with history they are grown ahead of the frame (counts of the last verified frame, scaled to this frame's slots
		// and tiles, doubled) and the frame is launched in one go: the kernels check the pools themselves (FrameDev::checkCaps).
		// a frame more than twice the size of what the history was taken from (slots or tiles) is sized from its own counts instead
		const bool async = allowAsync && r->historySlots > 0 && (int64_t)slotTotal <= 2 * r->historySlots && (int64_t)tileTotal <= 2 * std::max<int64_t>(r->historyTiles, 1);
		if (async) {
			const double scale = std::max(1.0, (double)slotTotal / (double)r->historySlots) * std::max(1.0, (double)tileTotal / (double)std::max<int64_t>(r->historyTiles, 1));
			const uint32_t *h = r->history;
			const size_t wantCmds = (size_t)(h[0] * scale) + 1024, wantRows = (size_t)(h[1] * scale) + 4096, wantEntries = (size_t)(h[2] * scale) + 4096;
			const size_t wantChk = (size_t)(h[4] * scale) + 1024, wantUnits = (size_t)(h[7] * scale) + 1024;
			if (grow_pool(r->cmds, wantCmds + wantCmds / 4, sizeof(Cmd), 2.0, 0) || grow_pool(r->rows, wantRows + wantRows / 4, sizeof(int2), 2.0, 16)
			    || grow_pool(r->tileList, wantEntries + wantEntries / 4, 4, 2.0, 16) || grow_pool(r->chk, wantChk + wantChk / 4, sizeof(ChkRec), 2.0, 16)
			    || grow_pool(r->bigUnits, wantUnits + wantUnits / 4, 4, 2.0, 0)) { return 1; }
			if (grow_pool(r->bigItems, pool_elements(r->cmds, sizeof(Cmd), 0), sizeof(BigItem), 1.0, 0)) { return 1; }
			if (h[3] > (uint32_t)SORT_SMEM / 2 && grow_pool(r->sortTmp, pool_elements(r->tileList, 4, 16), 4, 1.0, 16)) { return 1; }
		}
		auto set_pools = [&]() {
			frame.cmds = (Cmd *)r->cmds.ptr; frame.rows = (int2 *)r->rows.ptr; frame.tileList = (uint32_t *)r->tileList.ptr; frame.chk = (ChkRec *)r->chk.ptr;
			frame.bigItems = (BigItem *)r->bigItems.ptr; frame.bigUnits = (uint32_t *)r->bigUnits.ptr; frame.sortTmp = (uint32_t *)r->sortTmp.ptr;
			frame.capCmds = std::min(pool_elements(r->cmds, sizeof(Cmd), 0), pool_elements(r->bigItems, sizeof(BigItem), 0));
			frame.capRows = pool_elements(r->rows, sizeof(int2), 16); frame.capEntries = pool_elements(r->tileList, 4, 16);
			frame.capChk = pool_elements(r->chk, sizeof(ChkRec), 16); frame.capUnits = pool_elements(r->bigUnits, 4, 0);
			frame.capSortTmp = pool_elements(r->sortTmp, 4, 16);
		};
		set_pools();
		frame.checkCaps = async ? 1 : 0;
		frame.sortScheduled = (!async || r->history[3] > (uint32_t)LOCAL_SORT / 2u) ? 1 : 0;
		const int slot = (int)(r->serial & 1u);
		const uint32_t serial = ++r->serial;
		uint32_t *hostSlot = r->hostTotals + 16 * slot;
		frame.hostSlot = r->hostTotalsDevice + 16 * slot; frame.serial = serial;
		DFPSR_LAUNCH_CHAINED(setup_kernel<false>, blockTotal, SETUP_THREADS, 0, stream, frame);
		DFPSR_LAUNCH_CHAINED(counts_kernel, 1 + (tileTotal + COUNTS_THREADS - 1) / COUNTS_THREADS, COUNTS_THREADS, 0, stream, frame);
		if (timing) { tLaunched = host_now_us(); }
		uint32_t unitEstimate, maxTileEstimate;
		bool secondHalf = true;
		if (!async) {
			DFPSR_CHECK_CUDA(cudaStreamSynchronize(stream));
			if (timing) { tSynced = host_now_us(); }
			DFPSR_REQUIRE(hostSlot[TOTAL_TICKET] == serial, "renderer_end: the counting pass did not report (slot holds frame %u, expected %u)", hostSlot[TOTAL_TICKET], serial);
			const uint32_t commandTotal = hostSlot[0], rowTotal = hostSlot[1], entryTotal = hostSlot[2], maxTile = hostSlot[3];
			r->lastCommands = commandTotal;
			for (int i = 0; i < TOTAL_WORDS; i++) { r->history[i] = hostSlot[i]; }
			r->historySlots = std::max(slotTotal, 1); r->historyTiles = tileTotal;
			r->recentUnits = hostSlot[7]; r->recentSlots = std::max(slotTotal, 1);
			secondHalf = commandTotal > 0;
			unitEstimate = hostSlot[7]; maxTileEstimate = maxTile;
			if (secondHalf) {
				// an asynchronous renderer gets head room at once, so that the next frames fit without another wait
				const double slack = allowAsync ? 1.5 : 1.0;
				if (grow_pool(r->cmds, commandTotal, sizeof(Cmd), slack, 0) || grow_pool(r->rows, rowTotal, sizeof(int2), slack, 16)
				    || grow_pool(r->tileList, entryTotal, 4, slack, 16) || grow_pool(r->chk, hostSlot[4], sizeof(ChkRec), slack, 16)) { return 1; }
				if (unitEstimate > 0 && (grow_pool(r->bigItems, pool_elements(r->cmds, sizeof(Cmd), 0), sizeof(BigItem), 1.0, 0) || grow_pool(r->bigUnits, unitEstimate, 4, slack, 0))) { return 1; }
				if (maxTile > (uint32_t)SORT_SMEM && grow_pool(r->sortTmp, pool_elements(r->tileList, 4, 16), 4, 1.0, 16)) { return 1; }
				set_pools();
				// DFPSR_POISON=1 (tests): the checkpoint and row pools are filled with 0xFF before the emit pass, so that a record the tile
				// kernel needs and nothing wrote reads as "nothing here" (mode -1, empty rows) and shows up as missing pixels instead of
				// passing on what an earlier frame left at the same address.
				static const bool poison = getenv("DFPSR_POISON") != nullptr && atoi(getenv("DFPSR_POISON")) != 0;
				if (poison) {
					if (r->chk.ptr && r->chk.capacity > 0) { DFPSR_CHECK_CUDA(cudaMemsetAsync(r->chk.ptr, 0xFF, r->chk.capacity, stream)); }
					if (r->rows.ptr && r->rows.capacity > 0) { DFPSR_CHECK_CUDA(cudaMemsetAsync(r->rows.ptr, 0xFF, r->rows.capacity, stream)); }
				}
			}
		} else {
			PendingFrame &p = r->pending;
			p.active = true; p.serial = serial; p.slot = slot; p.stream = stream;
			p.views = r->views; p.tasks = r->tasks; p.textures = r->textures;
			p.depthOnly = r->depthOnly; p.occluded = r->occluded; p.exact = r->exact;
			p.gridSnapshots = r->gridSnapshots;
			if (r->occluded) { p.grid = r->grid; p.gridWidth = r->gridWidth; p.gridHeight = r->gridHeight; p.gridAllocW = r->gridAllocW; p.gridAllocH = r->gridAllocH; }
			g_pendingRenderers.push_back(r); g_pendingFrames++;
			// the grid of the unit kernel follows the most recent frame (scaled to this frame's slots), not the largest one seen
			unitEstimate = (uint32_t)std::min(4.0e9, (double)r->recentUnits * ((double)slotTotal / (double)r->recentSlots) * 1.25) + 1u;
			// the list sort is launched when an earlier frame came within a factor of two of needing it; a frame that needs it without
			// having it scheduled is dropped by counts_kernel and drawn again like one that outgrew a pool
			maxTileEstimate = r->history[3] > (uint32_t)LOCAL_SORT / 2u ? 0xFFFFFFFFu : 0u;
			p.slots = slotTotal;
		}
		if (secondHalf) {
			DFPSR_LAUNCH_CHAINED(setup_kernel<true>, blockTotal, SETUP_THREADS, 0, stream, frame);
			if (async) { DFPSR_CHECK_CUDA(cudaEventRecord(r->counted[slot], stream)); } // setup_kernel<true> publishes the verdict of an asynchronous frame
			if (unitEstimate > 0 || async) {
				// frames with few units (a single 1080p frame: 15 k) are latency-bound: two threads per unit; large batches keep one.
				// The kernels stride over the units the device counted, so an estimate only sizes the grid.
				const uint32_t most = (uint32_t)sm_count() * 64u;
				if (unitEstimate <= (uint32_t)sm_count() * 2048u) { DFPSR_LAUNCH_CHAINED(big_units_kernel<true>, std::min(most, (2u * unitEstimate + BIG_UNITS_THREADS - 1u) / BIG_UNITS_THREADS), BIG_UNITS_THREADS, 0, stream, frame); }
				else { DFPSR_LAUNCH_CHAINED(big_units_kernel<false>, std::min(most * 4u, (unitEstimate + BIG_UNITS_THREADS - 1u) / BIG_UNITS_THREADS), BIG_UNITS_THREADS, 0, stream, frame); }
			}
			if (maxTileEstimate > (uint32_t)LOCAL_SORT) { DFPSR_LAUNCH_CHAINED(sort_lists_kernel, (tileTotal + SORT_THREADS - 1) / SORT_THREADS, SORT_THREADS, 0, stream, frame); }
		}
	}
	// grid = (tile columns of the widest view / warps per CTA, tile rows of the tallest view, views)
	uint32_t widest = 0, tallest = 0;
	for (const ViewDev &v : r->views) { widest = std::max(widest, (uint32_t)v.tilesX); tallest = std::max(tallest, (uint32_t)v.tilesY); }
	DFPSR_REQUIRE(viewCount <= 65535 && tallest <= 65535u, "more than 65535 views in one batch or a target taller than 262140 rows");
	const dim3 grid((widest + RASTER_WARPS - 1) / RASTER_WARPS, tallest, (unsigned)viewCount);
	if (r->depthOnly) { DFPSR_LAUNCH_CHAINED(tile_kernel_depth, grid, RASTER_WARPS * 32, 0, stream, frame); }
	else if (immediate) { DFPSR_LAUNCH_CHAINED(tile_kernel_immediate, grid, RASTER_WARPS * 32, 0, stream, frame); }
	else if (exactFrame && r->textures.empty()) { DFPSR_LAUNCH_CHAINED(tile_kernel_deferred_raw, grid, RASTER_WARPS * 32, 0, stream, frame); }
	else if (exactFrame) { DFPSR_LAUNCH_CHAINED(tile_kernel_deferred, grid, RASTER_WARPS * 32, 0, stream, frame); }
	else { DFPSR_LAUNCH_CHAINED(tile_kernel_tolerance, grid, RASTER_WARPS * 32, 0, stream, frame); }
	if (r->wireframe && !r->depthOnly && taskCount > 0) { DFPSR_LAUNCH(wireframe_kernel, blockTotal, SETUP_THREADS, 0, stream, frame); }
	if (timing) {
		fprintf(stderr, "renderer_end: %zu tasks, %zu views | layout+upload %.0f us, first launches %.0f us, wait for totals %.0f us, second launches %.0f us\n",
		        taskCount, viewCount, tUploaded - tStart, tLaunched - tUploaded, tSynced > 0.0 ? tSynced - tLaunched : 0.0, host_now_us() - (tSynced > 0.0 ? tSynced : tLaunched));
	}
	return 0;
}

// Looks at the counts of the renderer's frame in flight (waiting for its counting pass if it has not run yet). A frame that did not fit
// its pools drew nothing: it is drawn again here, through the path that waits for the counts, before anything else is queued.
static int verify_frame(dfpsr_renderer *r) {
	PendingFrame &p = r->pending;
	if (!p.active) { return 0; }
	DFPSR_CHECK_CUDA(cudaEventSynchronize(r->counted[p.slot]));
	p.active = false;
	forget_pending(r);
	const uint32_t *hostSlot = r->hostTotals + 16 * p.slot;
	DFPSR_REQUIRE(hostSlot[TOTAL_TICKET] == p.serial, "renderer: the counting pass of frame %u did not report (slot holds frame %u)", p.serial, hostSlot[TOTAL_TICKET]);
	r->lastCommands = hostSlot[0];
	if (hostSlot[TOTAL_OVERFLOW] == 0u) {
		// the pools follow the largest frame seen: a shrinking scene keeps its head room
		for (int i = 0; i < TOTAL_WORDS; i++) { r->history[i] = std::max(r->history[i], hostSlot[i]); }
		r->recentUnits = hostSlot[7]; r->recentSlots = std::max(p.slots, 1);
		return 0;
	}
	// ---- the frame was dropped on the device: nothing of it (or behind it) may still be running while its inputs are laid out again
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(p.stream));
	std::swap(r->views, p.views); std::swap(r->tasks, p.tasks); std::swap(r->textures, p.textures); std::swap(r->gridSnapshots, p.gridSnapshots);
	std::swap(r->depthOnly, p.depthOnly); std::swap(r->occluded, p.occluded); std::swap(r->exact, p.exact);
	if (r->occluded) { std::swap(r->grid, p.grid); std::swap(r->gridWidth, p.gridWidth); std::swap(r->gridHeight, p.gridHeight); std::swap(r->gridAllocW, p.gridAllocW); std::swap(r->gridAllocH, p.gridAllocH); }
	const bool occludedFrame = r->occluded;
	const int status = run_frame(r, p.stream, false);
	if (occludedFrame) { std::swap(r->grid, p.grid); std::swap(r->gridWidth, p.gridWidth); std::swap(r->gridHeight, p.gridHeight); std::swap(r->gridAllocW, p.gridAllocW); std::swap(r->gridAllocH, p.gridAllocH); }
	std::swap(r->views, p.views); std::swap(r->tasks, p.tasks); std::swap(r->textures, p.textures); std::swap(r->gridSnapshots, p.gridSnapshots);
	std::swap(r->depthOnly, p.depthOnly); std::swap(r->occluded, p.occluded); std::swap(r->exact, p.exact);
	r->historySlots = std::max<int64_t>(r->historySlots, 1); // run_frame refreshed the history from the redrawn frame
	static const bool verbose = getenv("DFPSR_END_TIMING") != nullptr;
	if (verbose) { fprintf(stderr, "renderer: frame %u did not fit its pools and was drawn again\n", p.serial); }
	return status;
}

namespace dfpsr {
// DFPSR_LAUNCH and the copy helpers call this before they queue anything while frames are unverified: whatever consumes a frame through
// this library is queued behind the frame's (possible) second drawing.
int verify_pending_frames() {
	if (g_insideFrame) { return 0; }
	while (!g_pendingRenderers.empty()) {
		if (verify_frame(g_pendingRenderers.back())) { return 1; }
	}
	return 0;
}
}

static int renderer_end_internal(dfpsr_renderer *r, cudaStream_t stream) {
	DFPSR_REQUIRE(r->receiving, "Called renderer_end without renderer_begin!");
	r->receiving = false;
	if (verify_frame(r)) { return 1; }
	// a frame with the wireframe overlay waits for its counts: the overlay is a debug view and is not part of the redraw of a dropped frame
	const int status = run_frame(r, stream, r->async && !r->wireframe);
	r->wireframe = false;
	return status;
}

extern "C" {

// DFPSR_PRECISION=tolerance in the environment changes the initial default (A/B runs of the test-suite and the bench)
static bool initial_precision_exact() { const char *e = getenv("DFPSR_PRECISION"); return !(e && strcmp(e, "tolerance") == 0); }
static thread_local bool g_defaultExact = initial_precision_exact();
// DFPSR_ASYNC=1 in the environment: renderers do not wait for their frames' counts (dfpsr_set_default_async)
static bool initial_async() { const char *e = getenv("DFPSR_ASYNC"); return e && atoi(e) != 0; }
static thread_local bool g_defaultAsync = initial_async();
static thread_local dfpsr_renderer *g_immediate = nullptr;

int dfpsr_renderer_create(dfpsr_renderer **out) {
	DFPSR_REQUIRE(out != nullptr, "dfpsr_renderer_create: null output");
	int n = 0;
	DFPSR_REQUIRE(cudaGetDeviceCount(&n) == cudaSuccess && n > 0, "no CUDA device available; dfpsr_b200 has no CPU fallback");
	*out = new (std::nothrow) dfpsr_renderer();
	DFPSR_REQUIRE(*out != nullptr, "out of host memory");
	(*out)->exact = g_defaultExact;
	(*out)->async = g_defaultAsync;
	return 0;
}

int dfpsr_renderer_set_async(dfpsr_renderer *renderer, int32_t enabled) {
	DFPSR_REQUIRE(renderer != nullptr, "renderer_set_async: renderer does not exist");
	if (!enabled && verify_frame(renderer)) { return 1; }
	renderer->async = enabled != 0;
	return 0;
}

int dfpsr_set_default_async(int32_t enabled) {
	g_defaultAsync = enabled != 0;
	if (g_immediate != nullptr) { return dfpsr_renderer_set_async(g_immediate, enabled); }
	return 0;
}

int dfpsr_renderer_flush(dfpsr_renderer *renderer) {
	DFPSR_REQUIRE(renderer != nullptr, "renderer_flush: renderer does not exist");
	return verify_frame(renderer);
}

int dfpsr_flush(void) { return dfpsr::verify_pending_frames(); }

int dfpsr_renderer_set_debug_wireframe(dfpsr_renderer *renderer, int32_t enabled) {
	DFPSR_REQUIRE(renderer != nullptr, "dfpsr_renderer_set_debug_wireframe: null renderer");
	renderer->wireframe = enabled != 0;
	return 0;
}

int dfpsr_renderer_set_precision(dfpsr_renderer *renderer, int32_t precision) {
	DFPSR_REQUIRE(renderer != nullptr, "renderer_set_precision: renderer does not exist");
	DFPSR_REQUIRE(precision == DFPSR_PRECISION_EXACT || precision == DFPSR_PRECISION_TOLERANCE, "renderer_set_precision: unknown precision %d", precision);
	DFPSR_REQUIRE(!renderer->receiving, "renderer_set_precision: call outside of renderer_begin / renderer_end");
	renderer->exact = precision == DFPSR_PRECISION_EXACT;
	return 0;
}

int dfpsr_set_default_precision(int32_t precision) {
	DFPSR_REQUIRE(precision == DFPSR_PRECISION_EXACT || precision == DFPSR_PRECISION_TOLERANCE, "set_default_precision: unknown precision %d", precision);
	g_defaultExact = precision == DFPSR_PRECISION_EXACT;
	if (g_immediate != nullptr) { g_immediate->exact = g_defaultExact; }
	return 0;
}

int dfpsr_renderer_destroy(dfpsr_renderer *renderer) {
	delete renderer;
	return 0;
}

static int begin_one_view(dfpsr_renderer *r, const dfpsr_image *color, const dfpsr_image *depth, bool depthOnly, bool clear, uint32_t clearColor, float clearDepth) {
	ViewDev v;
	if (make_view(v, color, depth, clear, clearColor, clearDepth)) { return 1; }
	if (renderer_begin_internal(r, depthOnly)) { return 1; }
	r->views.push_back(v);
	return 0;
}

int dfpsr_renderer_begin(dfpsr_renderer *renderer, const dfpsr_image *color, const dfpsr_image *depth) {
	DFPSR_REQUIRE(renderer != nullptr, "renderer_begin: renderer does not exist");
	return begin_one_view(renderer, color, depth, false, false, 0u, 0.0f);
}

int dfpsr_renderer_begin_cleared(dfpsr_renderer *renderer, const dfpsr_image *color, const dfpsr_image *depth, uint32_t packedClearColor, float clearDepth) {
	DFPSR_REQUIRE(renderer != nullptr, "renderer_begin: renderer does not exist");
	return begin_one_view(renderer, color, depth, false, true, packedClearColor, clearDepth);
}

static bool box_visible(const dfpsr_renderer *r, const float *mn, const float *mx, const dfpsr_transform3d *m2w, const dfpsr_camera *camera);

int dfpsr_renderer_set_clip_rows(dfpsr_renderer *renderer, int32_t top, int32_t bottom) {
	DFPSR_REQUIRE(renderer != nullptr && renderer->receiving && renderer->views.size() == 1, "renderer_set_clip_rows: call between renderer_begin and renderer_end");
	ViewDev &v = renderer->views[0];
	DFPSR_REQUIRE(top >= 0 && top <= bottom && bottom <= v.height, "renderer_set_clip_rows: rows [%d, %d) outside the %d-row target", top, bottom, v.height);
	DFPSR_REQUIRE(top % TILE_H == 0 && (bottom % TILE_H == 0 || bottom == v.height), "renderer_set_clip_rows: strip boundaries must be multiples of %d rows", TILE_H);
	v.clipTop = top; v.clipBottom = bottom;
	return 0;
}

int dfpsr_renderer_give_task(dfpsr_renderer *renderer, const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr && model != nullptr && modelToWorld != nullptr && camera != nullptr, "renderer_giveTask: null argument");
	DFPSR_REQUIRE(renderer->receiving, "Cannot call renderer_giveTask before renderer_begin!");
	(void)stream;
	// ref: api/modelAPI.cpp:229-234 — whole models hidden by the occluders given so far are skipped
	if (renderer->occluded && dfpsr_camera_is_box_seen(camera, model->minBound, model->maxBound, modelToWorld) && !box_visible(renderer, model->minBound, model->maxBound, modelToWorld, camera)) { return 0; }
	return add_model_task(renderer, 0, model, modelToWorld, camera);
}

// Many models in one call, the whole-model tests on the device (a broad phase for scenes of thousands of models, where
// one host test per model per frame serialises the submission). ref: api/modelAPI.cpp:214-281, api/rendererAPI.cpp:302-351, :538-543.
int dfpsr_renderer_give_tasks(dfpsr_renderer *renderer, const dfpsr_model *models, const dfpsr_transform3d *modelToWorld, int32_t count, const dfpsr_camera *camera, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr && camera != nullptr, "renderer_giveTasks: null argument");
	DFPSR_REQUIRE(renderer->receiving, "Cannot call renderer_giveTasks before renderer_begin!");
	(void)stream;
	if (count <= 0) { return 0; }
	DFPSR_REQUIRE(models != nullptr && modelToWorld != nullptr, "renderer_giveTasks: null models");
	dfpsr_renderer *r = renderer;
	const ViewDev &v = r->views[0];
	if (v.width <= 0 || v.height <= 0) { return 0; }
	int32_t gridOffset = -1;
	if (r->occluded) { // the models of this call are tested against the occluders given SO FAR, like renderer_giveTask does at its call
		gridOffset = (int32_t)r->gridSnapshots.size();
		r->gridSnapshots.insert(r->gridSnapshots.end(), r->grid.begin(), r->grid.end());
	}
	for (int32_t i = 0; i < count; i++) {
		const dfpsr_model &model = models[i];
		if (model.polygonCount <= 0) { continue; }
		const int32_t diffuseIndex = r->depthOnly ? -1 : register_texture(r, &model.diffuse);
		const int32_t lightIndex = r->depthOnly ? -1 : register_texture(r, &model.light);
		DFPSR_REQUIRE(diffuseIndex != -2 && lightIndex != -2, "more than %d textures in one frame", MAX_TEXTURES);
		r->tasks.emplace_back();
		TaskParams &task = r->tasks.back();
		task.points = model.points; task.polygons = model.polygons;
		task.pointCount = model.pointCount; task.slotCount = model.polygonCount * 2;
		task.view = 0;
		task.modelToWorld = modelToWorld[i]; task.camera = *camera;
		task.filter = model.filter; task.depthOnly = r->depthOnly ? 1 : 0;
		task.diffuseIndex = diffuseIndex; task.lightIndex = lightIndex;
		task.cullOnDevice = 1; task.gridOffset = gridOffset;
		memcpy(task.minBound, model.minBound, sizeof(task.minBound)); memcpy(task.maxBound, model.maxBound, sizeof(task.maxBound));
	}
	return 0;
}

int dfpsr_renderer_give_task_triangles(dfpsr_renderer *renderer, const dfpsr_triangle *triangles, int32_t count, const dfpsr_texture *diffuse, const dfpsr_texture *light, int32_t filter, const dfpsr_camera *camera, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr && camera != nullptr, "renderer_giveTask_triangle: null argument");
	DFPSR_REQUIRE(renderer->receiving, "Cannot call renderer_giveTask_triangle before renderer_begin!");
	if (count <= 0) { return 0; }
	DFPSR_REQUIRE(triangles != nullptr, "renderer_giveTask_triangle: null triangles");
	const ViewDev &v = renderer->views[0];
	if (v.width <= 0 || v.height <= 0) { return 0; }
	// vertex data is copied at submission (ref: api/rendererAPI.h:103-107)
	size_t index = renderer->uploadCount++;
	if (renderer->uploads.size() <= index) { renderer->uploads.resize(index + 1); }
	if (renderer->uploads[index].reserve((size_t)count * sizeof(dfpsr_triangle))) { return 1; }
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(renderer->uploads[index].ptr, triangles, (size_t)count * sizeof(dfpsr_triangle), cudaMemcpyHostToDevice, as_stream(stream)));
	TaskParams task;
	memset(&task, 0, sizeof(task));
	task.triangles = (const dfpsr_triangle *)renderer->uploads[index].ptr;
	task.slotCount = count;
	task.view = 0;
	task.camera = *camera;
	task.filter = filter;
	task.depthOnly = renderer->depthOnly ? 1 : 0;
	task.diffuseIndex = register_texture(renderer, diffuse);
	task.lightIndex = register_texture(renderer, light);
	DFPSR_REQUIRE(task.diffuseIndex != -2 && task.lightIndex != -2, "more than %d textures in one frame", MAX_TEXTURES);
	renderer->tasks.push_back(task);
	return 0;
}

// ---- occlusion (ref: api/rendererAPI.cpp:181-351, :403-477, :521-562)

// ref: api/rendererAPI.cpp:181-192 prepareForOcclusion
static void prepare_for_occlusion(dfpsr_renderer *r) {
	const ViewDev &v = r->views[0];
	r->gridWidth = (v.width + (CELL_SIZE - 1)) / CELL_SIZE;
	r->gridHeight = (v.height + (CELL_SIZE - 1)) / CELL_SIZE;
	if (!r->occluded) {
		if (!(!r->grid.empty() && r->gridAllocW >= r->gridWidth && r->gridAllocH >= r->gridHeight)) {
			r->gridAllocW = r->gridWidth; r->gridAllocH = r->gridHeight;
			r->grid.assign((size_t)(r->gridAllocW * r->gridAllocH > 0 ? r->gridAllocW * r->gridAllocH : 1), INFINITY);
		}
		for (float &cell : r->grid) { cell = INFINITY; }
	}
	r->occluded = true;
}
static float grid_read(const dfpsr_renderer *r, int32_t x, int32_t y) { // image_readPixel_clamp; a grid that does not exist reads 0
	if (r->grid.empty() || r->gridAllocW <= 0 || r->gridAllocH <= 0) { return 0.0f; }
	x = x < 0 ? 0 : (x >= r->gridAllocW ? r->gridAllocW - 1 : x); y = y < 0 ? 0 : (y >= r->gridAllocH ? r->gridAllocH - 1 : y);
	return r->grid[(size_t)y * r->gridAllocW + x];
}
static PPoint host_camera_point(const dfpsr_camera *c, const dfpsr_transform3d *m, const float *p, float *cameraSpace) {
	float wx = (p[0] * m->xAxis[0] + p[1] * m->yAxis[0] + p[2] * m->zAxis[0]) + m->position[0];
	float wy = (p[0] * m->xAxis[1] + p[1] * m->yAxis[1] + p[2] * m->zAxis[1]) + m->position[1];
	float wz = (p[0] * m->xAxis[2] + p[1] * m->yAxis[2] + p[2] * m->zAxis[2]) + m->position[2];
	const dfpsr_transform3d &l = c->location;
	float dx = wx - l.position[0], dy = wy - l.position[1], dz = wz - l.position[2];
	cameraSpace[0] = dx * l.xAxis[0] + dy * l.xAxis[1] + dz * l.xAxis[2];
	cameraSpace[1] = dx * l.yAxis[0] + dy * l.yAxis[1] + dz * l.yAxis[2];
	cameraSpace[2] = dx * l.zAxis[0] + dy * l.zAxis[1] + dz * l.zAxis[2];
	return camera_to_screen(*c, cameraSpace[0], cameraSpace[1], cameraSpace[2]);
}
static bool host_plane_inside(const float *pl, float x, float y, float z) { return (((pl[0] * x) + (pl[1] * y) + (pl[2] * z)) - pl[3]) <= 0.0f; }
static void box_corner(float *out, const float *mn, const float *mx, int i) { out[0] = (i & 4) ? mx[0] : mn[0]; out[1] = (i & 2) ? mx[1] : mn[1]; out[2] = (i & 1) ? mx[2] : mn[2]; }
// ref: api/rendererAPI.cpp:89-95 getPixelBoundFromProjection (left, top, right, bottom)
static void pixel_bound_from_projection(const PPoint *hull, int count, int32_t *bound) {
	bound[0] = (int32_t)(hull[0].fx / 256); bound[1] = (int32_t)(hull[0].fy / 256); bound[2] = bound[0] + 1; bound[3] = bound[1] + 1;
	for (int p = 1; p < count; p++) {
		int32_t x = (int32_t)(hull[p].fx / 256), y = (int32_t)(hull[p].fy / 256);
		if (x < bound[0]) { bound[0] = x; } if (y < bound[1]) { bound[1] = y; }
		if (x + 1 > bound[2]) { bound[2] = x + 1; } if (y + 1 > bound[3]) { bound[3] = y + 1; }
	}
}

// ref: api/rendererAPI.cpp:268-299 occludeFromBox (host: a handful of occluders per frame, the grid is a few thousand cells)
int dfpsr_renderer_occlude_from_box(dfpsr_renderer *renderer, const float minBound[3], const float maxBound[3], const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera) {
	DFPSR_REQUIRE(renderer != nullptr && minBound && maxBound && modelToWorld && camera, "renderer_occludeFromBox: null argument");
	DFPSR_REQUIRE(renderer->receiving && renderer->views.size() == 1, "Cannot call renderer_occludeFromBox without first calling renderer_begin!");
	prepare_for_occlusion(renderer);
	PPoint projected[8], hull[8];
	for (int p = 0; p < 8; p++) { // projectHull (:76-88)
		float local[3], cs[3];
		box_corner(local, minBound, maxBound, p);
		projected[p] = host_camera_point(camera, modelToWorld, local, cs);
		for (int s = 0; s < camera->cullPlaneCount; s++) {
			if (!host_plane_inside(camera->cullPlanes[s], cs[0] * 0.5f, cs[1] * 0.5f, cs[2] * 1.0f)) { return 0; }
		}
	}
	// jarvisConvexHullAlgorithm (:38-74)
	int count = 0, l = 0;
	for (int i = 1; i < 8; i++) { if (projected[i].fx < projected[l].fx) { l = i; } }
	int p = l;
	do {
		if (count >= 8) { break; }
		hull[count++] = projected[p];
		int q = (p + 1) % 8;
		for (int i = 0; i < 8; i++) {
			const PPoint &a = projected[p], &b = projected[i], &c = projected[q];
			if ((b.fy - a.fy) * (c.fx - b.fx) - (b.fx - a.fx) * (c.fy - b.fy) < 0) { q = i; }
		}
		p = q;
	} while (p != l);
	// occludeFromSortedHull (:218-241)
	int32_t bound[4];
	pixel_bound_from_projection(hull, count, bound);
	if (!(bound[2] - bound[0] > CELL_SIZE && bound[3] - bound[1] > CELL_SIZE)) { return 0; }
	float distance = 0.0f;
	for (int c = 0; c < count; c++) { if (hull[c].csz > distance) { distance = hull[c].csz; } }
	CellBound cells = outer_cell_bound(bound[0], bound[1], bound[2], bound[3], renderer->gridWidth, renderer->gridHeight);
	for (int32_t cy = cells.y0; cy < cells.y1; cy++) {
		for (int32_t cx = cells.x0; cx < cells.x1; cx++) {
			if (cell_inside_of_hull(hull, count, cx, cy) && distance < grid_read(renderer, cx, cy)) { renderer->grid[(size_t)cy * renderer->gridAllocW + cx] = distance; }
		}
	}
	return 0;
}

int dfpsr_renderer_has_occluders(const dfpsr_renderer *renderer) { return renderer != nullptr && renderer->occluded ? 1 : 0; } // ref: :557-562

// ref: api/rendererAPI.cpp:302-351 isHullOccluded / isBoxOccluded, negated as in renderer_isBoxVisible (:538-543)
static bool box_visible(const dfpsr_renderer *r, const float *mn, const float *mx, const dfpsr_transform3d *m2w, const dfpsr_camera *camera) {
	PPoint projected[8];
	float cs[8][3];
	for (int p = 0; p < 8; p++) {
		float local[3];
		box_corner(local, mn, mx, p);
		projected[p] = host_camera_point(camera, m2w, local, cs[p]);
	}
	for (int s = 0; s < camera->cullPlaneCount; s++) {
		bool allOutside = true;
		for (int p = 0; p < 8; p++) { if (host_plane_inside(camera->cullPlanes[s], cs[p][0], cs[p][1], cs[p][2])) { allOutside = false; break; } }
		if (allOutside) { return false; }
	}
	int32_t bound[4];
	pixel_bound_from_projection(projected, 8, bound);
	float closest = INFINITY;
	for (int c = 0; c < 8; c++) { if (projected[c].csz < closest) { closest = projected[c].csz; } }
	int32_t gridWidth = (r->views[0].width + (CELL_SIZE - 1)) / CELL_SIZE, gridHeight = (r->views[0].height + (CELL_SIZE - 1)) / CELL_SIZE;
	CellBound cells = outer_cell_bound(bound[0], bound[1], bound[2], bound[3], gridWidth, gridHeight);
	for (int32_t cy = cells.y0; cy < cells.y1; cy++) {
		for (int32_t cx = cells.x0; cx < cells.x1; cx++) { if (closest < grid_read(r, cx, cy)) { return true; } }
	}
	return false;
}
int dfpsr_renderer_is_box_visible(const dfpsr_renderer *renderer, const float minBound[3], const float maxBound[3], const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, int32_t *visible) {
	DFPSR_REQUIRE(renderer != nullptr && minBound && maxBound && modelToWorld && camera && visible, "renderer_isBoxVisible: null argument");
	DFPSR_REQUIRE(renderer->receiving && renderer->views.size() == 1, "Cannot call renderer_isBoxVisible without first calling renderer_begin and giving occluder shapes to the pass!");
	*visible = box_visible(renderer, minBound, maxBound, modelToWorld, camera) ? 1 : 0;
	return 0;
}

// ref: api/rendererAPI.cpp:403-477 occludeFromTopRows: the depth buffer is on the device, so one small kernel reduces the scanned row of
// every cell and the host merges the result into its grid (one synchronisation, like the reference's read of the depth buffer).
int dfpsr_renderer_occlude_from_top_rows(dfpsr_renderer *renderer, const dfpsr_camera *camera, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr && camera != nullptr, "renderer_occludeFromTopRows: null argument");
	DFPSR_REQUIRE(renderer->receiving && renderer->views.size() == 1, "Cannot call renderer_occludeFromTopRows without first calling renderer_begin!");
	const ViewDev &v = renderer->views[0];
	DFPSR_REQUIRE(v.depth.data != nullptr, "Cannot call renderer_occludeFromTopRows without having given a depth buffer in renderer_begin!");
	prepare_for_occlusion(renderer);
	const int32_t cells = renderer->gridWidth * renderer->gridHeight;
	if (cells <= 0) { return 0; }
	if (renderer->dGrid.reserve((size_t)cells * sizeof(float) + renderer->grid.size() * sizeof(float) + 16)) { return 1; }
	std::vector<float> rowExtremes((size_t)cells);
	DFPSR_LAUNCH(top_rows_kernel, (cells + 255) / 256, 256, 0, as_stream(stream), v.depth, v.width, renderer->gridWidth, renderer->gridHeight, camera->perspective, (float *)renderer->dGrid.ptr);
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(rowExtremes.data(), renderer->dGrid.ptr, (size_t)cells * sizeof(float), cudaMemcpyDeviceToHost, as_stream(stream)));
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
	for (int32_t gy = 0; gy < renderer->gridHeight; gy++) {
		for (int32_t gx = 0; gx < renderer->gridWidth; gx++) {
			float maxDistance = rowExtremes[(size_t)gy * renderer->gridWidth + gx];
			float &cell = renderer->grid[(size_t)gy * renderer->gridAllocW + gx];
			if (maxDistance < cell) { cell = maxDistance; }
		}
	}
	return 0;
}

// ref: api/rendererAPI.cpp:242-258 occludeFromExistingTriangles: the triangles queued so far are projected, culled and clipped on the
// device and rasterised into the grid as occluders; the grid then returns to the host for the following visibility queries.
int dfpsr_renderer_occlude_from_existing_triangles(dfpsr_renderer *renderer, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr, "renderer_occludeFromExistingTriangles: renderer does not exist");
	DFPSR_REQUIRE(renderer->receiving && renderer->views.size() == 1, "Cannot call renderer_occludeFromExistingTriangles without first calling renderer_begin!");
	prepare_for_occlusion(renderer);
	dfpsr_renderer *r = renderer;
	if (r->tasks.empty()) { return 0; }
	cudaStream_t s = as_stream(stream);
	int32_t slotTotal = 0, blockTotal = 0;
	if (layout_tasks(r, slotTotal, blockTotal)) { return 1; }
	if (r->dTasks.reserve(r->tasks.size() * sizeof(TaskParams) + 16) || r->dViews.reserve(sizeof(ViewDev)) || r->dGrid.reserve(r->grid.size() * sizeof(float) + 16)) { return 1; }
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(r->dTasks.ptr, r->tasks.data(), r->tasks.size() * sizeof(TaskParams), cudaMemcpyHostToDevice, s));
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(r->dViews.ptr, r->views.data(), sizeof(ViewDev), cudaMemcpyHostToDevice, s));
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(r->dGrid.ptr, r->grid.data(), r->grid.size() * sizeof(float), cudaMemcpyHostToDevice, s));
	FrameDev frame;
	memset(&frame, 0, sizeof(frame));
	frame.tasks = (const TaskParams *)r->dTasks.ptr; frame.views = (const ViewDev *)r->dViews.ptr;
	frame.taskCount = (int32_t)r->tasks.size(); frame.viewCount = 1; frame.blockCount = blockTotal;
	frame.gridWidth = r->gridWidth; frame.gridHeight = r->gridHeight; frame.gridStride = r->gridAllocW; frame.gridRows = r->gridAllocH;
	if (!r->gridSnapshots.empty()) { // tasks of dfpsr_renderer_give_tasks are tested against their own snapshot of the grid here as well
		if (r->dGridSnapshots.reserve(r->gridSnapshots.size() * sizeof(float) + 16)) { return 1; }
		DFPSR_CHECK_CUDA(cudaMemcpyAsync(r->dGridSnapshots.ptr, r->gridSnapshots.data(), r->gridSnapshots.size() * sizeof(float), cudaMemcpyHostToDevice, s));
		frame.gridSnapshots = (const float *)r->dGridSnapshots.ptr;
	}
	if (launch_projection(r, frame, s)) { return 1; }
	DFPSR_LAUNCH(occlude_existing_kernel, blockTotal, SETUP_THREADS, 0, s, frame, (float *)r->dGrid.ptr);
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(r->grid.data(), r->dGrid.ptr, r->grid.size() * sizeof(float), cudaMemcpyDeviceToHost, s));
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(s));
	return 0;
}

int dfpsr_renderer_end(dfpsr_renderer *renderer, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr, "renderer_end: renderer does not exist");
	return renderer_end_internal(renderer, as_stream(stream));
}

int dfpsr_renderer_last_command_count(dfpsr_renderer *renderer, int64_t *count, void *stream) {
	DFPSR_REQUIRE(renderer != nullptr && count != nullptr, "renderer_last_command_count: null argument");
	(void)stream;
	if (verify_frame(renderer)) { return 1; }
	*count = renderer->lastCommands;
	return 0;
}

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
	DFPSR_REQUIRE(modelToWorld != nullptr && camera != nullptr, "model_render: null argument");
	dfpsr_renderer *r;
	if (immediate_renderer(&r)) { return 1; }
	if (begin_one_view(r, color, depth, false, false, 0u, 0.0f)) { return 1; }
	int status = add_model_task(r, 0, model, modelToWorld, camera);
	if (status) { r->receiving = false; return status; }
	return renderer_end_internal(r, as_stream(stream));
}

int dfpsr_model_render_depth(const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_image *depth, const dfpsr_camera *camera, void *stream) {
	if (model == nullptr || !image_exists(depth)) { return 0; } // ref: api/modelAPI.cpp:203, renderCore.cpp:409
	DFPSR_REQUIRE(modelToWorld != nullptr && camera != nullptr, "model_renderDepth: null argument");
	dfpsr_renderer *r;
	if (immediate_renderer(&r)) { return 1; }
	if (begin_one_view(r, nullptr, depth, true, false, 0u, 0.0f)) { return 1; }
	int status = add_model_task(r, 0, model, modelToWorld, camera);
	if (status) { r->receiving = false; return status; }
	return renderer_end_internal(r, as_stream(stream));
}

// Shared by dfpsr_model_render_views and dfpsr_model_render_depth_views: every (model, transform, camera, target) tuple becomes one task
// with its own view; the whole batch goes through ONE sequence of launches.
static int render_batch(const dfpsr_model *const *models, const dfpsr_transform3d *transforms, const dfpsr_image *colors, const dfpsr_image *depths, const dfpsr_camera *cameras, const int32_t *targetOfTask, int32_t taskCount, int32_t viewCount, bool depthOnly, bool clear, uint32_t clearColor, float clearDepth, cudaStream_t stream) {
	dfpsr_renderer *r;
	if (immediate_renderer(&r)) { return 1; }
	if (renderer_begin_internal(r, depthOnly)) { return 1; }
	static const bool timing = getenv("DFPSR_END_TIMING") != nullptr;
	const double tStart = timing ? host_now_us() : 0.0;
	for (int32_t v = 0; v < viewCount; v++) {
		ViewDev view;
		if (make_view(view, (colors && !depthOnly) ? colors + v : nullptr, depths ? depths + v : nullptr, clear, clearColor, clearDepth)) { r->receiving = false; return 1; }
		r->views.push_back(view);
	}
	for (int32_t t = 0; t < taskCount; t++) {
		int32_t v = targetOfTask ? targetOfTask[t] : t;
		if (v < 0 || v >= viewCount) { r->receiving = false; set_error("render batch: task %d names target %d of %d", t, v, viewCount); return 1; }
		if (models[t] == nullptr) { continue; }
		if (add_model_task(r, v, models[t], transforms + t, cameras + t)) { r->receiving = false; return 1; }
	}
	if (timing) { fprintf(stderr, "render_batch: %d submissions -> %zu tasks in %.0f us\n", taskCount, r->tasks.size(), host_now_us() - tStart); }
	return renderer_end_internal(r, stream);
}

int dfpsr_model_render_views(const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_image *colors, const dfpsr_image *depths, const dfpsr_camera *cameras, int32_t count, int32_t clear, void *stream) {
	DFPSR_REQUIRE(model != nullptr && modelToWorld != nullptr && cameras != nullptr, "model_render_views: null argument");
	if (count <= 0) { return 0; }
	std::vector<const dfpsr_model *> models((size_t)count, model);
	std::vector<dfpsr_transform3d> transforms((size_t)count, *modelToWorld);
	return render_batch(models.data(), transforms.data(), colors, depths, cameras, nullptr, count, count, false, clear != 0, 0u, 0.0f, as_stream(stream));
}

int dfpsr_model_render_depth_batch(const dfpsr_model *const *models, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *cameras, const int32_t *targetOfTask, int32_t taskCount, const dfpsr_image *depths, int32_t targetCount, int32_t clear, float clearDepth, void *stream) {
	DFPSR_REQUIRE(models != nullptr && modelToWorld != nullptr && cameras != nullptr && depths != nullptr, "model_render_depth_batch: null argument");
	if (targetCount <= 0) { return 0; }
	return render_batch(models, modelToWorld, nullptr, depths, cameras, targetOfTask, taskCount, targetCount, true, clear != 0, 0u, clearDepth, as_stream(stream));
}

int dfpsr_project_points(const float *points, int32_t count, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, dfpsr_projected_point *outDevice, void *stream) {
	DFPSR_REQUIRE(points != nullptr && modelToWorld != nullptr && camera != nullptr && outDevice != nullptr, "project_points: null argument");
	if (count <= 0) { return 0; }
	int grid = (count + 255) / 256;
	if (grid > sm_count() * 8) { grid = sm_count() * 8; }
	DFPSR_LAUNCH(project_points_kernel, grid, 256, 0, as_stream(stream), points, count, *modelToWorld, *camera, (PPoint *)outDevice);
	return 0;
}

} // extern "C"
