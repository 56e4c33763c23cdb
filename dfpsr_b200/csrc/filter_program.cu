// filter_program.cu — filter_mapRgbaU8 / filter_generateRgbaU8 with an ARBITRARY per-pixel function on the device
// (ref: api/filterAPI.h:54-79, api/filterAPI.cpp:759-782).
//
// The reference takes a host lambda `ColorRgbaI32 f(int32_t x, int32_t y)` that may capture images and read them with
// image_readPixel_clamp / _border / _tile. A device cannot run a host lambda, so the function travels as TEXT: the body of
//     int4 pixel(int x, int y)        // (red, green, blue, alpha), any int; saturated to 0..255 and packed afterwards
// in CUDA C++, compiled for the current device with NVRTC the first time it is used and cached by its text (one cache per process).
// Inside the body: x and y (target coordinates plus the start offsets), and for every source image i handed to the call
//     read_clamp(i, x, y)   read_border(i, x, y)   read_border(i, x, y, int4 border)   read_tile(i, x, y)      -> int4 (r, g, b, a)
//     source_width(i)   source_height(i)
// with the reference's out-of-bound rules (clamp to the edge / transparent black or the given border / wrap around) and channels in
// red, green, blue, alpha order whatever the image's pack order. The generated kernel is the same streaming kernel as the built-in
// maps: four pixels per thread, one 16-byte store, every buffer touched once — HBM-bound for any reasonably short body.
// The enumerated ops of dfpsr_filter_map (pixel_ops.cu) stay as pre-compiled instances of the same thing.
#include "common.cuh"

#include <dlfcn.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

using namespace dfpsr;

// ---- NVRTC, loaded on first use (the library itself must load on machines without a CUDA toolkit in the loader path)
typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
	void *handle = nullptr;
	int (*createProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
	int (*compileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
	int (*getProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
	int (*getProgramLog)(nvrtcProgram, char *) = nullptr;
	int (*getCUBINSize)(nvrtcProgram, size_t *) = nullptr;
	int (*getCUBIN)(nvrtcProgram, char *) = nullptr;
	int (*destroyProgram)(nvrtcProgram *) = nullptr;
	const char *(*getErrorString)(int) = nullptr;
	bool load() {
		if (handle) { return true; }
		const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
		for (const char *name : names) { handle = dlopen(name, RTLD_NOW | RTLD_LOCAL); if (handle) { break; } }
		if (!handle) { return false; }
#define DFPSR_NVRTC_SYMBOL(field, symbol) field = (decltype(field))dlsym(handle, symbol); if (!field) { return false; }
		DFPSR_NVRTC_SYMBOL(createProgram, "nvrtcCreateProgram") DFPSR_NVRTC_SYMBOL(compileProgram, "nvrtcCompileProgram")
		DFPSR_NVRTC_SYMBOL(getProgramLogSize, "nvrtcGetProgramLogSize") DFPSR_NVRTC_SYMBOL(getProgramLog, "nvrtcGetProgramLog")
		DFPSR_NVRTC_SYMBOL(getCUBINSize, "nvrtcGetCUBINSize") DFPSR_NVRTC_SYMBOL(getCUBIN, "nvrtcGetCUBIN")
		DFPSR_NVRTC_SYMBOL(destroyProgram, "nvrtcDestroyProgram") DFPSR_NVRTC_SYMBOL(getErrorString, "nvrtcGetErrorString")
#undef DFPSR_NVRTC_SYMBOL
		return true;
	}
};

const int MAX_SOURCES = 4;
struct ProgramImage { void *data; int32_t width, height, stride, packOrder; };
struct ProgramParams { ProgramImage target; ProgramImage source[MAX_SOURCES]; int32_t sourceCount, startX, startY, pad_; };

// The kernel around the caller's body. Kept free of headers: NVRTC compiles it without an include path.
const char *PROLOGUE = R"SRC(
struct ProgramImage { void *data; int width, height, stride, packOrder; };
struct ProgramParams { ProgramImage target; ProgramImage source[4]; int sourceCount, startX, startY, pad_; };
__device__ __forceinline__ unsigned int dfpsr_shifts(int order) { // bit position of red | green << 8 | blue << 16 | alpha << 24 (PackOrder.h:85-96)
	return order == 1 ? (16u | (8u << 8) | (0u << 16) | (24u << 24)) : order == 2 ? (8u | (16u << 8) | (24u << 16) | (0u << 24))
	     : order == 3 ? (24u | (16u << 8) | (8u << 16) | (0u << 24)) : (0u | (8u << 8) | (16u << 16) | (24u << 24));
}
__device__ __forceinline__ int4 dfpsr_unpack(unsigned int c, unsigned int s) {
	return make_int4((int)((c >> (s & 31u)) & 255u), (int)((c >> ((s >> 8) & 31u)) & 255u), (int)((c >> ((s >> 16) & 31u)) & 255u), (int)((c >> ((s >> 24) & 31u)) & 255u));
}
__device__ __forceinline__ int4 dfpsr_fetch(const ProgramImage &image, int x, int y) {
	const unsigned int *row = (const unsigned int *)((const unsigned char *)image.data + (size_t)y * (size_t)image.stride);
	return dfpsr_unpack(__ldg(row + x), dfpsr_shifts(image.packOrder));
}
__device__ __forceinline__ int4 dfpsr_read_clamp(const ProgramParams &P, int i, int x, int y) { // image_readPixel_clamp
	const ProgramImage &image = P.source[i];
	if (image.data == 0 || image.width <= 0 || image.height <= 0) { return make_int4(0, 0, 0, 0); }
	x = x < 0 ? 0 : (x >= image.width ? image.width - 1 : x); y = y < 0 ? 0 : (y >= image.height ? image.height - 1 : y);
	return dfpsr_fetch(image, x, y);
}
__device__ __forceinline__ int4 dfpsr_read_border(const ProgramParams &P, int i, int x, int y, int4 border = make_int4(0, 0, 0, 0)) { // image_readPixel_border
	const ProgramImage &image = P.source[i];
	if (image.data == 0 || x < 0 || y < 0 || x >= image.width || y >= image.height) { return border; }
	return dfpsr_fetch(image, x, y);
}
__device__ __forceinline__ int dfpsr_wrap(int v, int size) { int m = v % size; return m < 0 ? m + size : m; }
__device__ __forceinline__ int4 dfpsr_read_tile(const ProgramParams &P, int i, int x, int y) { // image_readPixel_tile
	const ProgramImage &image = P.source[i];
	if (image.data == 0 || image.width <= 0 || image.height <= 0) { return make_int4(0, 0, 0, 0); }
	return dfpsr_fetch(image, dfpsr_wrap(x, image.width), dfpsr_wrap(y, image.height));
}
#define read_clamp(i, ...) dfpsr_read_clamp(P, (i), __VA_ARGS__)
#define read_border(i, ...) dfpsr_read_border(P, (i), __VA_ARGS__)
#define read_tile(i, ...) dfpsr_read_tile(P, (i), __VA_ARGS__)
#define source_width(i) (P.source[(i)].width)
#define source_height(i) (P.source[(i)].height)
__device__ __forceinline__ int4 pixel(const ProgramParams &P, int x, int y) {
)SRC";

const char *EPILOGUE = R"SRC(
}
__device__ __forceinline__ unsigned int dfpsr_saturate_pack(int4 c, unsigned int s) { // image_saturateAndPack (filterAPI.cpp:759-777)
	const unsigned int r = (unsigned int)min(max(c.x, 0), 255), g = (unsigned int)min(max(c.y, 0), 255), b = (unsigned int)min(max(c.z, 0), 255), a = (unsigned int)min(max(c.w, 0), 255);
	return (r << (s & 31u)) | (g << ((s >> 8) & 31u)) | (b << ((s >> 16) & 31u)) | (a << ((s >> 24) & 31u));
}
extern "C" __global__ void __launch_bounds__(256) dfpsr_map_program(ProgramParams P) {
	const int x0 = (int)(blockIdx.x * blockDim.x + threadIdx.x) * 4, y = (int)(blockIdx.y * blockDim.y + threadIdx.y);
	if (x0 >= P.target.width || y >= P.target.height) { return; }
	const unsigned int s = dfpsr_shifts(P.target.packOrder);
	unsigned int *out = (unsigned int *)((unsigned char *)P.target.data + (size_t)y * (size_t)P.target.stride) + x0;
	const int n = min(4, P.target.width - x0);
	unsigned int packed[4];
#pragma unroll
	for (int i = 0; i < 4; i++) { packed[i] = i < n ? dfpsr_saturate_pack(pixel(P, x0 + i + P.startX, y + P.startY), s) : 0u; }
	if (n == 4 && (((size_t)out) & 15u) == 0) { *(uint4 *)out = make_uint4(packed[0], packed[1], packed[2], packed[3]); }
	else { for (int i = 0; i < n; i++) { out[i] = packed[i]; } }
}
)SRC";

struct Compiled { cudaLibrary_t library = nullptr; cudaKernel_t kernel = nullptr; };
std::mutex g_cacheLock;
std::unordered_map<std::string, Compiled> g_cache; // key: device architecture + body text
Nvrtc g_nvrtc;

int compile_body(const char *body, Compiled &out) {
	int device = 0, major = 0, minor = 0;
	DFPSR_CHECK_CUDA(cudaGetDevice(&device));
	DFPSR_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
	DFPSR_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
	char arch[64];
	snprintf(arch, sizeof(arch), "--gpu-architecture=sm_%d%d%s", major, minor, major >= 9 ? "a" : "");
	const std::string key = std::string(arch) + "\n" + body;
	std::lock_guard<std::mutex> guard(g_cacheLock);
	auto found = g_cache.find(key);
	if (found != g_cache.end()) { out = found->second; return 0; }
	DFPSR_REQUIRE(g_nvrtc.load(), "filter_map_program: libnvrtc.so.12 could not be loaded (%s); the run-time compiler of the CUDA toolkit is needed for per-pixel programs", dlerror() ? dlerror() : "missing symbol");
	const std::string source = std::string(PROLOGUE) + "\n" + body + "\n" + EPILOGUE;
	nvrtcProgram program = nullptr;
	int status = g_nvrtc.createProgram(&program, source.c_str(), "dfpsr_map_program.cu", 0, nullptr, nullptr);
	DFPSR_REQUIRE(status == 0, "filter_map_program: nvrtcCreateProgram failed: %s", g_nvrtc.getErrorString(status));
	const char *options[] = {arch, "--fmad=false", "--std=c++17"};
	status = g_nvrtc.compileProgram(program, 3, options);
	if (status != 0) {
		size_t logSize = 0;
		g_nvrtc.getProgramLogSize(program, &logSize);
		std::vector<char> log(logSize + 1, '\0');
		if (logSize > 0) { g_nvrtc.getProgramLog(program, log.data()); }
		g_nvrtc.destroyProgram(&program);
		set_error("filter_map_program: the pixel function does not compile: %.400s", log.data());
		return 1;
	}
	size_t size = 0;
	g_nvrtc.getCUBINSize(program, &size);
	std::vector<char> cubin(size);
	status = g_nvrtc.getCUBIN(program, cubin.data());
	g_nvrtc.destroyProgram(&program);
	DFPSR_REQUIRE(status == 0 && size > 0, "filter_map_program: no device code came out of NVRTC: %s", g_nvrtc.getErrorString(status));
	Compiled compiled;
	DFPSR_CHECK_CUDA(cudaLibraryLoadData(&compiled.library, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
	DFPSR_CHECK_CUDA(cudaLibraryGetKernel(&compiled.kernel, compiled.library, "dfpsr_map_program"));
	g_cache[key] = compiled;
	out = compiled;
	return 0;
}

} // namespace

extern "C" {

int dfpsr_filter_map_program(const dfpsr_image *target, const char *body, const dfpsr_image *sources, int32_t sourceCount, int32_t startX, int32_t startY, void *stream) {
	if (target == nullptr || target->data == nullptr || target->width <= 0 || target->height <= 0) { return 0; } // ref: api/filterAPI.cpp:773
	DFPSR_REQUIRE(body != nullptr, "filter_map_program: null pixel function");
	DFPSR_REQUIRE(sourceCount >= 0 && sourceCount <= MAX_SOURCES && (sourceCount == 0 || sources != nullptr), "filter_map_program: 0..%d source images, got %d", MAX_SOURCES, sourceCount);
	int devices = 0;
	DFPSR_REQUIRE(cudaGetDeviceCount(&devices) == cudaSuccess && devices > 0, "no CUDA device available; dfpsr_b200 has no CPU fallback");
	Compiled compiled;
	if (compile_body(body, compiled)) { return 1; }
	ProgramParams params;
	memset(&params, 0, sizeof(params));
	params.target = ProgramImage{target->data, target->width, target->height, target->stride, target->packOrder};
	for (int32_t i = 0; i < sourceCount; i++) { params.source[i] = ProgramImage{sources[i].data, sources[i].width, sources[i].height, sources[i].stride, sources[i].packOrder}; }
	params.sourceCount = sourceCount; params.startX = startX; params.startY = startY;
	if (g_pendingFrames > 0 && verify_pending_frames()) { return 1; } // the sources may be frames in flight
	const dim3 block(64, 4), grid((unsigned)(((target->width + 3) / 4 + 63) / 64), (unsigned)((target->height + 3) / 4));
	void *arguments[] = {&params};
	if (g_profile) { profile_begin("map_program_kernel", as_stream(stream)); }
	const cudaError_t launched = cudaLaunchKernel((const void *)compiled.kernel, grid, block, arguments, 0, as_stream(stream));
	if (g_profile) { profile_end(as_stream(stream)); }
	DFPSR_REQUIRE(launched == cudaSuccess, "launch of the compiled pixel function failed: %s", cudaGetErrorString(launched));
	g_launches.fetch_add(1, std::memory_order_relaxed);
	return check_launch("map_program_kernel");
}

} // extern "C"
