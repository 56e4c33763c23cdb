// host_api.cu — the non-kernel part of the C ABI: errors, device memory helpers, camera construction
// (host arithmetic mirroring Camera.h), texture layout. Kernels live in raster.cu and pixel_ops.cu.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>
#include <float.h>
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <utility>

namespace dfpsr {

thread_local char g_error[512] = "";
std::atomic<unsigned long long> g_launches{0}; // launches of every thread (relaxed: a statistic, nothing is ordered by it)

void set_error(const char *fmt, ...) {
	va_list args;
	va_start(args, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, args);
	va_end(args);
}

int check_launch(const char *name) {
	cudaError_t err = cudaGetLastError();
	if (err != cudaSuccess) {
		set_error("launch of %s failed: %s", name, cudaGetErrorString(err));
		return 1;
	}
	return 0;
}

static bool initial_chain_launches() { const char *e = getenv("DFPSR_CHAIN"); return !(e && atoi(e) == 0); }
bool g_chainLaunches = initial_chain_launches();

// ---- per-kernel profiling
bool g_profile = false;
namespace {
struct ProfileEntry { const char *name; std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events; double ms = 0.0; long long launches = 0; };
// The table is shared by all threads and guarded by g_profileMutex; the launch a thread is in the middle of is its own.
std::mutex g_profileMutex;
std::vector<ProfileEntry> g_profileEntries;
thread_local int g_profileCurrent = -1;
thread_local cudaEvent_t g_profileStart;
void profile_collect() { // with g_profileMutex held
	for (ProfileEntry &e : g_profileEntries) {
		for (auto &pair : e.events) {
			cudaEventSynchronize(pair.second);
			float ms = 0.0f;
			if (cudaEventElapsedTime(&ms, pair.first, pair.second) == cudaSuccess) { e.ms += ms; e.launches++; }
			cudaEventDestroy(pair.first); cudaEventDestroy(pair.second);
		}
		e.events.clear();
	}
}
}
void profile_begin(const char *name, cudaStream_t stream) {
	{
		std::lock_guard<std::mutex> lock(g_profileMutex);
		g_profileCurrent = -1;
		for (size_t i = 0; i < g_profileEntries.size(); i++) { if (strcmp(g_profileEntries[i].name, name) == 0) { g_profileCurrent = (int)i; break; } }
		if (g_profileCurrent < 0) { g_profileEntries.push_back(ProfileEntry{name, {}, 0.0, 0}); g_profileCurrent = (int)g_profileEntries.size() - 1; }
	}
	cudaEventCreate(&g_profileStart);
	cudaEventRecord(g_profileStart, stream);
}
void profile_end(cudaStream_t stream) {
	if (g_profileCurrent < 0) { return; }
	cudaEvent_t stop;
	cudaEventCreate(&stop);
	cudaEventRecord(stop, stream);
	std::lock_guard<std::mutex> lock(g_profileMutex);
	if (g_profileCurrent >= (int)g_profileEntries.size()) { return; } // the table was reset between the two halves of this launch
	ProfileEntry &entry = g_profileEntries[(size_t)g_profileCurrent];
	entry.events.push_back({g_profileStart, stop});
	if (entry.events.size() >= 4096) { profile_collect(); }
}

int sm_count() {
	static int cached = 0;
	if (cached == 0) {
		int device = 0;
		if (cudaGetDevice(&device) != cudaSuccess) { return 148; }
		if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || cached <= 0) { cached = 148; }
	}
	return cached;
}

int DeviceBuffer::reserve(size_t bytes) {
	if (bytes <= capacity) { return 0; }
	size_t grown = capacity + capacity / 2;
	if (grown < bytes) { grown = bytes; }
	grown = (grown + 255) & ~(size_t)255;
	void *fresh = nullptr;
	cudaError_t err = cudaMalloc(&fresh, grown);
	if (err != cudaSuccess) {
		set_error("cudaMalloc of %zu bytes failed: %s", grown, cudaGetErrorString(err));
		return 1;
	}
	if (ptr) { cudaFree(ptr); } // contents are per-frame scratch; nothing to preserve
	ptr = fresh;
	capacity = grown;
	return 0;
}

void DeviceBuffer::release() {
	if (ptr) { cudaFree(ptr); }
	ptr = nullptr;
	capacity = 0;
}

struct V3 { float x, y, z; };

// ref: math/FVector.h:113-120
static V3 normalize3(V3 v) {
	float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
	if (l == 0.0f) { return V3{0.0f, 0.0f, 1.0f}; }
	return V3{v.x / l, v.y / l, v.z / l};
}

static void set_plane(float *dst, V3 normal, float offset) { // ref: math/FPlane3D.h:36
	V3 n = normalize3(normal);
	dst[0] = n.x; dst[1] = n.y; dst[2] = n.z; dst[3] = offset;
}

// ref: implementation/render/Camera.h:56-72
static int frustum_perspective(float planes[6][4], float nearClip, float farClip, float widthSlope, float heightSlope) {
	set_plane(planes[0], V3{-1.0f, 0.0f, -widthSlope}, 0.0f);
	set_plane(planes[1], V3{1.0f, 0.0f, -widthSlope}, 0.0f);
	set_plane(planes[2], V3{0.0f, 1.0f, -heightSlope}, 0.0f);
	set_plane(planes[3], V3{0.0f, -1.0f, -heightSlope}, 0.0f);
	set_plane(planes[4], V3{0.0f, 0.0f, -1.0f}, -nearClip);
	set_plane(planes[5], V3{0.0f, 0.0f, 1.0f}, farClip);
	return farClip == INFINITY ? 5 : 6;
}

// ref: implementation/render/Camera.h:47-55
static int frustum_orthogonal(float planes[6][4], float halfWidth, float halfHeight) {
	set_plane(planes[0], V3{-1.0f, 0.0f, 0.0f}, halfWidth);
	set_plane(planes[1], V3{1.0f, 0.0f, 0.0f}, halfWidth);
	set_plane(planes[2], V3{0.0f, 1.0f, 0.0f}, halfHeight);
	set_plane(planes[3], V3{0.0f, -1.0f, 0.0f}, halfHeight);
	return 4;
}

static const float cullRatio = 1.0001f; // ref: Camera.h:113
static const float clipRatio = 2.0f;    // ref: Camera.h:118

} // namespace dfpsr

using namespace dfpsr;

extern "C" {

int dfpsr_abi_version(void) { return DFPSR_B200_ABI_VERSION; }

const char *dfpsr_last_error(void) { return g_error; }

int dfpsr_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { return 0; }
	return n;
}

int dfpsr_init(int device) {
	int n = 0;
	cudaError_t err = cudaGetDeviceCount(&n);
	DFPSR_REQUIRE(err == cudaSuccess && n > 0, "no CUDA device available (%s); dfpsr_b200 has no CPU fallback", cudaGetErrorString(err));
	DFPSR_REQUIRE(device >= 0 && device < n, "device %d out of range (found %d)", device, n);
	DFPSR_CHECK_CUDA(cudaSetDevice(device));
	DFPSR_CHECK_CUDA(cudaFree(0));
	return 0;
}

int dfpsr_profile_enable(int enabled) {
	std::lock_guard<std::mutex> lock(g_profileMutex);
	if (!enabled) { profile_collect(); }
	g_profile = enabled != 0;
	return 0;
}
int dfpsr_profile_reset(void) {
	std::lock_guard<std::mutex> lock(g_profileMutex);
	profile_collect();
	g_profileEntries.clear();
	return 0;
}
int dfpsr_profile_count(void) {
	std::lock_guard<std::mutex> lock(g_profileMutex);
	profile_collect();
	return (int)g_profileEntries.size();
}
int dfpsr_profile_read(int index, const char **name, double *milliseconds, int64_t *launches) {
	std::lock_guard<std::mutex> lock(g_profileMutex);
	profile_collect();
	DFPSR_REQUIRE(index >= 0 && index < (int)g_profileEntries.size(), "profile_read: index out of range");
	*name = g_profileEntries[index].name;
	*milliseconds = g_profileEntries[index].ms;
	*launches = g_profileEntries[index].launches;
	return 0;
}

uint64_t dfpsr_launch_count(void) { return g_launches; }
void dfpsr_reset_launch_count(void) { g_launches = 0; }

int dfpsr_malloc(void **devicePtr, size_t bytes) {
	DFPSR_CHECK_CUDA(cudaMalloc(devicePtr, bytes));
	return 0;
}
int dfpsr_free(void *devicePtr) {
	DFPSR_CHECK_CUDA(cudaFree(devicePtr));
	return 0;
}
int dfpsr_malloc_host(void **pinnedPtr, size_t bytes) {
	DFPSR_CHECK_CUDA(cudaMallocHost(pinnedPtr, bytes));
	return 0;
}
int dfpsr_free_host(void *pinnedPtr) {
	DFPSR_CHECK_CUDA(cudaFreeHost(pinnedPtr));
	return 0;
}
int dfpsr_upload(void *devicePtr, const void *hostPtr, size_t bytes, void *stream) {
	if (g_pendingFrames > 0 && verify_pending_frames()) { return 1; } // frames in flight are verified before their targets are touched
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(devicePtr, hostPtr, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
	return 0;
}
int dfpsr_download(void *hostPtr, const void *devicePtr, size_t bytes, void *stream) {
	if (g_pendingFrames > 0 && verify_pending_frames()) { return 1; } // frames in flight are verified before their targets are touched
	DFPSR_CHECK_CUDA(cudaMemcpyAsync(hostPtr, devicePtr, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
	return 0;
}
int dfpsr_upload_2d(void *devicePtr, size_t deviceStride, const void *hostPtr, size_t hostStride, size_t rowBytes, size_t rows, void *stream) {
	if (g_pendingFrames > 0 && verify_pending_frames()) { return 1; } // frames in flight are verified before their targets are touched
	DFPSR_CHECK_CUDA(cudaMemcpy2DAsync(devicePtr, deviceStride, hostPtr, hostStride, rowBytes, rows, cudaMemcpyHostToDevice, as_stream(stream)));
	return 0;
}
int dfpsr_download_2d(void *hostPtr, size_t hostStride, const void *devicePtr, size_t deviceStride, size_t rowBytes, size_t rows, void *stream) {
	if (g_pendingFrames > 0 && verify_pending_frames()) { return 1; } // frames in flight are verified before their targets are touched
	DFPSR_CHECK_CUDA(cudaMemcpy2DAsync(hostPtr, hostStride, devicePtr, deviceStride, rowBytes, rows, cudaMemcpyDeviceToHost, as_stream(stream)));
	return 0;
}
int dfpsr_stream_synchronize(void *stream) {
	if (g_pendingFrames > 0 && verify_pending_frames()) { return 1; } // frames in flight are verified before their targets are touched
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
	return 0;
}

// ref: implementation/render/Camera.h:128-150
int dfpsr_camera_create_perspective(dfpsr_camera *out, const dfpsr_transform3d *location, float imageWidth, float imageHeight, float widthSlope, float nearClip, float farClip) {
	DFPSR_REQUIRE(out && location, "dfpsr_camera_create_perspective: null argument");
	memset(out, 0, sizeof(*out));
	float heightSlope = widthSlope * imageHeight / imageWidth;
	out->perspective = 1;
	out->location = *location;
	out->widthSlope = widthSlope; out->heightSlope = heightSlope;
	out->invWidthSlope = 0.5f / widthSlope; out->invHeightSlope = 0.5f / heightSlope;
	out->imageWidth = imageWidth; out->imageHeight = imageHeight;
	out->nearClip = nearClip; out->farClip = farClip;
	out->cullPlaneCount = frustum_perspective(out->cullPlanes, nearClip, farClip, widthSlope * cullRatio, heightSlope * cullRatio);
	out->clipPlaneCount = frustum_perspective(out->clipPlanes, nearClip, farClip, widthSlope * clipRatio, heightSlope * clipRatio);
	return 0;
}

// ref: implementation/render/Camera.h:152-156
int dfpsr_camera_create_orthogonal(dfpsr_camera *out, const dfpsr_transform3d *location, float imageWidth, float imageHeight, float halfWidth) {
	DFPSR_REQUIRE(out && location, "dfpsr_camera_create_orthogonal: null argument");
	memset(out, 0, sizeof(*out));
	float halfHeight = halfWidth * imageHeight / imageWidth;
	out->perspective = 0;
	out->location = *location;
	out->widthSlope = halfWidth; out->heightSlope = halfHeight;
	out->invWidthSlope = 0.5f / halfWidth; out->invHeightSlope = 0.5f / halfHeight;
	out->imageWidth = imageWidth; out->imageHeight = imageHeight;
	out->nearClip = -FLT_MAX; out->farClip = INFINITY;
	out->cullPlaneCount = frustum_orthogonal(out->cullPlanes, halfWidth * cullRatio, halfHeight * cullRatio);
	out->clipPlaneCount = frustum_orthogonal(out->clipPlanes, halfWidth * clipRatio, halfHeight * clipRatio);
	return 0;
}

// ref: implementation/render/Camera.h:73-95, :202-217
// dfpsr_camera_is_box_seen lives in host_math.cpp (four corners per SSE vector, same IEEE operations in the same order).

// ref: api/textureAPI.cpp:30-41, :65-78; implementation/image/Texture.h:63-102
int dfpsr_texture_layout(dfpsr_texture *out, int32_t width, int32_t height, int32_t resolutions) {
	DFPSR_REQUIRE(out != nullptr, "dfpsr_texture_layout: null output");
	DFPSR_REQUIRE(resolutions >= 1, "Tried to create a texture without any resolutions stored, which would be empty!");
	DFPSR_REQUIRE(width >= 1 && height >= 1, "Tried to create a texture of %d x %d pixels, which would be empty!", width, height);
	DFPSR_REQUIRE(width <= 32768 && height <= 32768, "Tried to create a texture of %d x %d pixels, which exceeds the maximum texture dimensions of 32768 x 32768 pixels!", width, height);
	uint32_t log2w = 15, log2h = 15;
	for (uint32_t l = 0; l < 15; l++) { if ((1u << l) >= (uint32_t)width) { log2w = l; break; } }
	for (uint32_t l = 0; l < 15; l++) { if ((1u << l) >= (uint32_t)height) { log2h = l; break; } }
	uint32_t maxMip = (uint32_t)(resolutions - 1);
	if (maxMip > log2w) { maxMip = log2w; }
	if (maxMip > log2h) { maxMip = log2h; }
	if (maxMip > 15) { maxMip = 15; }
	uint64_t highest = (uint64_t)1 << (log2w + log2h);
	uint64_t pixelCount = 0, levelCount = highest;
	for (int32_t level = (int32_t)maxMip; level >= 0; level--) { pixelCount |= levelCount; levelCount >>= 2; }
	DFPSR_REQUIRE(pixelCount < 4294967296ull, "texture of %d x %d pixels cannot be indexed with 32-bit offsets", width, height);
	out->data = nullptr;
	out->log2width = log2w; out->log2height = log2h; out->maxMipLevel = maxMip;
	out->startOffset = (uint32_t)(pixelCount & ~highest);
	out->maxLevelMask = (uint32_t)(highest - 1);
	out->totalPixels = (uint32_t)pixelCount;
	return 0;
}

} // extern "C"
