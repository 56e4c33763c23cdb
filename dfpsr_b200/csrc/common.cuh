// common.cuh — shared plumbing for the sm_100a kernels: error handling, launch accounting, pixel packing.
// Everything is compiled with -fmad=false: the reference is built without FMA contraction
// (ISO C++ mode, SURVEY.md §7 hard part 1) and bit-exact parity depends on identical rounding.
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dfpsr_b200.h"

namespace dfpsr {

// ---- error reporting (thread-local message, int status across the C ABI)
void set_error(const char *fmt, ...);
extern thread_local char g_error[512];

#define DFPSR_CHECK_CUDA(expr)                                                                         \
	do {                                                                                               \
		cudaError_t err_ = (expr);                                                                     \
		if (err_ != cudaSuccess) {                                                                     \
			dfpsr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err_), __FILE__, __LINE__); \
			return 1;                                                                                  \
		}                                                                                              \
	} while (0)

#define DFPSR_REQUIRE(cond, ...)              \
	do {                                      \
		if (!(cond)) {                        \
			dfpsr::set_error(__VA_ARGS__);    \
			return 1;                         \
		}                                     \
	} while (0)

// ---- launch accounting: every kernel launch of this library goes through DFPSR_LAUNCH
extern std::atomic<unsigned long long> g_launches;
int check_launch(const char *name);

// Optional per-kernel device timing (dfpsr_profile_enable): CUDA events on the launching stream around each launch.
extern bool g_profile;
void profile_begin(const char *name, cudaStream_t stream);
void profile_end(cudaStream_t stream);

// Frames of asynchronous renderers whose counts the host has not looked at yet (raster.cu). Everything this library queues while such
// frames exist first verifies them, so that a frame that has to be drawn again still comes before its consumers.
extern thread_local int g_pendingFrames;
extern thread_local bool g_insideFrame;
int verify_pending_frames();

#define DFPSR_LAUNCH(kernel, grid, block, smem, stream, ...)                      \
	do {                                                                          \
		if (dfpsr::g_pendingFrames > 0 && !dfpsr::g_insideFrame && dfpsr::verify_pending_frames()) { return 1; } \
		if (dfpsr::g_profile) { dfpsr::profile_begin(#kernel, (stream)); }        \
		kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);               \
		if (dfpsr::g_profile) { dfpsr::profile_end((stream)); }                   \
		dfpsr::g_launches.fetch_add(1, std::memory_order_relaxed);                                                      \
		if (dfpsr::check_launch(#kernel)) { return 1; }                           \
	} while (0)

// Kernels of one frame that follow each other on a stream are launched with programmatic stream serialisation: the next kernel's CTAs
// are scheduled while the previous kernel is still running and wait at chain_enter() (griddepcontrol.wait) until it has completed and
// its writes are visible. This takes the launch latency (2-3 us per kernel) off the critical path of small frames, where seven
// kernels of a few microseconds each follow one another. DFPSR_CHAIN=0 in the environment restores plain launches.
extern bool g_chainLaunches;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
	cudaLaunchConfig_t config = {};
	config.gridDim = grid; config.blockDim = block; config.dynamicSmemBytes = smem; config.stream = stream;
	cudaLaunchAttribute attribute[1];
	attribute[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attribute[0].val.programmaticStreamSerializationAllowed = 1;
	config.attrs = attribute; config.numAttrs = 1;
	return cudaLaunchKernelEx(&config, kernel, KArgs(args)...);
}

#define DFPSR_LAUNCH_CHAINED(kernel, grid, block, smem, stream, ...)                                                         \
	do {                                                                                                                     \
		if (dfpsr::g_pendingFrames > 0 && !dfpsr::g_insideFrame && dfpsr::verify_pending_frames()) { return 1; }              \
		if (dfpsr::g_profile || !dfpsr::g_chainLaunches) {                                                                   \
			if (dfpsr::g_profile) { dfpsr::profile_begin(#kernel, (stream)); }                                               \
			kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                                      \
			if (dfpsr::g_profile) { dfpsr::profile_end((stream)); }                                                          \
		} else {                                                                                                             \
			cudaError_t chainError_ = dfpsr::launch_chained(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__); \
			if (chainError_ != cudaSuccess) { dfpsr::set_error("launch of %s failed: %s", #kernel, cudaGetErrorString(chainError_)); return 1; } \
		}                                                                                                                    \
		dfpsr::g_launches.fetch_add(1, std::memory_order_relaxed);                                                                                                 \
		if (dfpsr::check_launch(#kernel)) { return 1; }                                                                      \
	} while (0)

// First statement of every kernel that may be launched chained: lets the NEXT kernel of the stream start being scheduled, then waits
// until the PREVIOUS one has completed (both are no-ops for plain launches). Nothing that another kernel wrote may be read before it.
__device__ __forceinline__ void chain_enter() {
#if defined(__CUDA_ARCH__)
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

inline cudaStream_t as_stream(void *s) { return (cudaStream_t)s; }

// Number of SMs of the current device (148 on B200); grids of persistent kernels are multiples of it.
int sm_count();

// Growable device buffer owned by a renderer / session.
struct DeviceBuffer {
	void *ptr = nullptr;
	size_t capacity = 0;
	int reserve(size_t bytes);
	void release();
};

// ---- device helpers

// Byte index of red, green, blue, alpha for each pack order (ref: implementation/image/PackOrder.h:85-96).
__host__ __device__ inline uint32_t pack_shifts(int order) {
	// four 8-bit fields: shift of red | green << 8 | blue << 16 | alpha << 24
	switch (order) {
		case DFPSR_PACK_BGRA: return 16u | (8u << 8) | (0u << 16) | (24u << 24);
		case DFPSR_PACK_ARGB: return 8u | (16u << 8) | (24u << 16) | (0u << 24);
		case DFPSR_PACK_ABGR: return 24u | (16u << 8) | (8u << 16) | (0u << 24);
		default: return 0u | (8u << 8) | (16u << 16) | (24u << 24);
	}
}

// ref: implementation/image/PackOrder.h:186-197 — clampUpper(x, 255.1) then truncating conversion.
// The reference's scalar conversion wraps negative inputs; inputs on this path are never negative.
__device__ __forceinline__ uint32_t saturated_byte(float v) {
	float c = v < 255.1f ? v : 255.1f;
	return (uint32_t)__float2int_rz(c);
}

__device__ __forceinline__ uint32_t pack_rgba_ordered(uint32_t r, uint32_t g, uint32_t b, uint32_t a, uint32_t shifts) {
	return (r << (shifts & 31u)) | (g << ((shifts >> 8) & 31u)) | (b << ((shifts >> 16) & 31u)) | (a << ((shifts >> 24) & 31u));
}

// ref: api/drawAPI.cpp:176-283 drawLineSuper. Step j along the major axis writes (major0 + j, minor0 + sign * k(j)) where the reference's
// running error (error += tilt; if error >= maxError { minor += sign; error -= 2 * maxError }) has the closed form
// k(j) = floor((j * tilt + maxError) / (2 * maxError)). Shared by the 2D line kernels (draw_ops.cu) and the wireframe overlay (raster.cu).
struct LineParams { int32_t major0, minor0, sign, steps, firstStep; int32_t majorIsY; long long tilt, maxError; };
__host__ __device__ inline bool line_params(int32_t width, int32_t height, int32_t x1, int32_t y1, int32_t x2, int32_t y2, LineParams &p) {
	if ((x1 < 0 && x2 < 0) || (y1 < 0 && y2 < 0) || (x1 >= width && x2 >= width) || (y1 >= height && y2 >= height)) { return false; }
	const long long dx = (long long)x2 - x1, dy = (long long)y2 - y1;
	const long long adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
	long long length;
	if (ady >= adx) { // vertical, or closer to vertical: walk down (ref: :203-236); a horizontal line has ady == 0 == adx only when both are 0
		if (y2 < y1) { int32_t t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
		p.majorIsY = 1; p.major0 = y1; p.minor0 = x1; p.sign = x2 > x1 ? 1 : -1; p.tilt = 2 * adx; p.maxError = ady; length = ady;
	} else { // closer to horizontal: walk right (ref: :237-276)
		if (x2 < x1) { int32_t t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
		p.majorIsY = 0; p.major0 = x1; p.minor0 = y1; p.sign = y2 > y1 ? 1 : -1; p.tilt = 2 * ady; p.maxError = adx; length = adx;
	}
	// only the steps whose major coordinate lies inside the image can write
	const long long limit = p.majorIsY ? height : width;
	const long long first = p.major0 < 0 ? -(long long)p.major0 : 0;
	long long last = limit - 1 - p.major0;
	if (length < last) { last = length; }
	if (last < first) { return false; }
	p.firstStep = (int32_t)first; p.steps = (int32_t)(last - first + 1);
	return true;
}
// Pixel of step i (counted from firstStep) of a line; false when it falls outside the image.
__device__ __forceinline__ bool line_pixel(const LineParams &p, int32_t i, int32_t width, int32_t height, int32_t &px, int32_t &py) {
	const long long j = (long long)p.firstStep + i;
	const long long k = p.maxError > 0 ? (j * p.tilt + p.maxError) / (2 * p.maxError) : 0;
	const long long major = (long long)p.major0 + j, minor = (long long)p.minor0 + p.sign * k;
	const long long x = p.majorIsY ? minor : major, y = p.majorIsY ? major : minor;
	px = (int32_t)x; py = (int32_t)y;
	return x >= 0 && x < width && y >= 0 && y < height;
}

template <typename T>
__device__ __forceinline__ T *row_ptr(void *data, int32_t stride, int32_t y) {
	return (T *)((uint8_t *)data + (size_t)y * (size_t)stride);
}
template <typename T>
__device__ __forceinline__ const T *row_ptr(const void *data, int32_t stride, int32_t y) {
	return (const T *)((const uint8_t *)data + (size_t)y * (size_t)stride);
}

} // namespace dfpsr
