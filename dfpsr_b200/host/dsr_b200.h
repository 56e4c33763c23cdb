// dsr_b200.h — C++14 host shim: DFPSR's own API names for the rendering hot path, implemented over the C ABI of
// include/dfpsr_b200.h (hand-written sm_100a kernels). Header-only; link with libdfpsr_b200.so. No CPU fallback: without a
// CUDA device every rendering call throws (the reference reports errors through throwError -> std::exception as well,
// ref: DFPSR/api/stringAPI.h:597-629).
//
// Each declaration cites the reference declaration it mirrors (paths relative to /root/reference/Source/DFPSR).
// Scope: what SDK/terrain/main.cpp:383-433, SDK/cube, templates/basic3D and SDK/SpriteEngine/{spriteAPI,lightAPI}.cpp call on
// the path. Images and textures live in device memory; host access (image_readPixel_*, image_writePixel,
// image_download) goes through a lazily synchronised host mirror, so a frame that is only rendered and presented costs one
// device-to-host copy and nothing else.
#pragma once

#include "../../include/dfpsr_b200.h"

#include <cstdint>
#include <cstring>
#include <cstdio>
#include <utility>
#include <cmath>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace dsr {

// ref: api/stringAPI.h:597 throwError
inline void throwError(const std::string &message) { throw std::runtime_error(message); }
inline void b200_check(int status) { if (status != 0) { throwError(dfpsr_last_error()); } }

// ---------------------------------------------------------------- math PODs (ref: math/FVector.h, FMatrix3x3.h, Transform3D.h)
struct FVector2D { float x = 0.0f, y = 0.0f; FVector2D() {} FVector2D(float x, float y) : x(x), y(y) {} };
struct FVector3D { float x = 0.0f, y = 0.0f, z = 0.0f; FVector3D() {} FVector3D(float x, float y, float z) : x(x), y(y), z(z) {} };
struct FVector4D { float x = 0.0f, y = 0.0f, z = 0.0f, w = 0.0f; FVector4D() {} FVector4D(float x, float y, float z, float w) : x(x), y(y), z(z), w(w) {} };
inline FVector3D operator+(const FVector3D &a, const FVector3D &b) { return FVector3D(a.x + b.x, a.y + b.y, a.z + b.z); }
inline FVector3D operator-(const FVector3D &a, const FVector3D &b) { return FVector3D(a.x - b.x, a.y - b.y, a.z - b.z); }
inline FVector3D operator-(const FVector3D &a) { return FVector3D(-a.x, -a.y, -a.z); }
inline FVector3D operator*(const FVector3D &a, float s) { return FVector3D(a.x * s, a.y * s, a.z * s); }
inline float dotProduct(const FVector3D &a, const FVector3D &b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
inline FVector3D crossProduct(const FVector3D &a, const FVector3D &b) { return FVector3D(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline FVector3D normalize(const FVector3D &v) { // ref: math/FVector.h:113-120
	float l = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
	if (l == 0.0f) { return FVector3D(0.0f, 0.0f, 1.0f); }
	return FVector3D(v.x / l, v.y / l, v.z / l);
}
struct FMatrix3x3 { // ref: math/FMatrix3x3.h:33-51
	FVector3D xAxis = FVector3D(1, 0, 0), yAxis = FVector3D(0, 1, 0), zAxis = FVector3D(0, 0, 1);
	FMatrix3x3() {}
	FMatrix3x3(const FVector3D &x, const FVector3D &y, const FVector3D &z) : xAxis(x), yAxis(y), zAxis(z) {}
	static FMatrix3x3 makeAxisSystem(const FVector3D &forward, const FVector3D &up) {
		FMatrix3x3 r;
		r.zAxis = normalize(forward);
		r.xAxis = normalize(crossProduct(normalize(up), r.zAxis));
		r.yAxis = normalize(crossProduct(r.zAxis, r.xAxis));
		return r;
	}
};
struct Transform3D { // ref: math/Transform3D.h:33-36
	FVector3D position;
	FMatrix3x3 transform;
	Transform3D() {}
	Transform3D(const FVector3D &position, const FMatrix3x3 &transform) : position(position), transform(transform) {}
};
inline dfpsr_transform3d b200_pod(const Transform3D &t) {
	dfpsr_transform3d r;
	r.position[0] = t.position.x; r.position[1] = t.position.y; r.position[2] = t.position.z;
	r.xAxis[0] = t.transform.xAxis.x; r.xAxis[1] = t.transform.xAxis.y; r.xAxis[2] = t.transform.xAxis.z;
	r.yAxis[0] = t.transform.yAxis.x; r.yAxis[1] = t.transform.yAxis.y; r.yAxis[2] = t.transform.yAxis.z;
	r.zAxis[0] = t.transform.zAxis.x; r.zAxis[1] = t.transform.zAxis.y; r.zAxis[2] = t.transform.zAxis.z;
	return r;
}

enum class PackOrderIndex { RGBA = 0, BGRA = 1, ARGB = 2, ABGR = 3 }; // ref: implementation/image/PackOrder.h:37-42
enum class Filter { Solid = 0, Alpha = 1 };                           // ref: implementation/render/constants.h:34
enum class Sampler { Nearest = 0, Linear = 1 };                       // ref: api/filterAPI.h:33-36
struct ColorRgbaI32 { // ref: implementation/image/Color.h
	int32_t red = 0, green = 0, blue = 0, alpha = 0;
	ColorRgbaI32() {}
	ColorRgbaI32(int32_t r, int32_t g, int32_t b, int32_t a) : red(r), green(g), blue(b), alpha(a) {}
};

struct IVector2D { int32_t x = 0, y = 0; IVector2D() {} IVector2D(int32_t x, int32_t y) : x(x), y(y) {} }; // ref: math/IVector.h
struct IVector3D { int32_t x = 0, y = 0, z = 0; IVector3D() {} IVector3D(int32_t x, int32_t y, int32_t z) : x(x), y(y), z(z) {} };

// ---------------------------------------------------------------- device context
inline void *&b200_stream() { static thread_local void *stream = nullptr; return stream; } // cudaStream_t used by this thread's calls
inline void b200_init(int device = 0) { b200_check(dfpsr_init(device)); }

// ---------------------------------------------------------------- images (ref: api/imageAPI.h:47-525, implementation/image/Image.h:58-183)
struct B200Buffer { // ref-counted device allocation + lazily synchronised host mirror
	void *device = nullptr;
	size_t bytes = 0;
	std::vector<uint8_t> host;
	bool hostValid = false, deviceDirty = false; // deviceDirty: kernels wrote since the last download
	bool owned = true;
	explicit B200Buffer(size_t bytes) : bytes(bytes) { b200_check(dfpsr_malloc(&device, bytes > 0 ? bytes : 1)); }
	B200Buffer(void *borrowedDevice, size_t bytes) : device(borrowedDevice), bytes(bytes), deviceDirty(true), owned(false) {} // a view of memory owned elsewhere
	~B200Buffer() { if (device && owned) { dfpsr_free(device); } }
	B200Buffer(const B200Buffer &) = delete;
	B200Buffer &operator=(const B200Buffer &) = delete;
	uint8_t *hostData() {
		if (!hostValid || deviceDirty) {
			host.resize(bytes);
			b200_check(dfpsr_download(host.data(), device, bytes, b200_stream()));
			b200_check(dfpsr_stream_synchronize(b200_stream()));
			hostValid = true; deviceDirty = false;
		}
		return host.data();
	}
	void pushHost() { b200_check(dfpsr_upload(device, host.data(), bytes, b200_stream())); b200_check(dfpsr_stream_synchronize(b200_stream())); }
};

template <typename PIXEL>
class B200Image {
public:
	std::shared_ptr<B200Buffer> buffer;
	int32_t width = 0, height = 0, stride = 0, startOffset = 0;
	PackOrderIndex packOrder = PackOrderIndex::RGBA;
	bool subImage = false;
	dfpsr_image pod() const {
		dfpsr_image r;
		r.data = buffer ? (void *)((uint8_t *)buffer->device + startOffset) : nullptr;
		r.width = width; r.height = height; r.stride = stride; r.packOrder = (int32_t)packOrder;
		return r;
	}
	void touchedByDevice() const { if (buffer) { buffer->deviceDirty = true; } }
};
using ImageRgbaU8 = B200Image<uint32_t>;
using ImageF32 = B200Image<float>;

template <typename PIXEL>
inline B200Image<PIXEL> b200_image_create(int32_t width, int32_t height, PackOrderIndex order) {
	if (width <= 0 || height <= 0) { throwError("image_create: non-positive dimensions"); } // ref: api/imageAPI.cpp:58-66
	B200Image<PIXEL> image;
	image.width = width; image.height = height;
	image.stride = ((width * (int32_t)sizeof(PIXEL) + 255) / 256) * 256; // rows on 256-byte boundaries (the reference pads to its heap alignment, imageAPI.cpp:50)
	image.packOrder = order;
	image.buffer = std::make_shared<B200Buffer>((size_t)image.stride * (size_t)height);
	return image;
}
// ref: api/imageAPI.h:63-70. zeroed = true clears like the reference does.
inline ImageRgbaU8 image_create_RgbaU8(int32_t width, int32_t height, bool zeroed = true) {
	ImageRgbaU8 image = b200_image_create<uint32_t>(width, height, PackOrderIndex::RGBA);
	if (zeroed) { dfpsr_image pod = image.pod(); b200_check(dfpsr_image_fill_rgba(&pod, 0, 0, 0, 0, b200_stream())); image.touchedByDevice(); }
	return image;
}
inline ImageRgbaU8 image_create_RgbaU8_native(int32_t width, int32_t height, PackOrderIndex order, bool zeroed = true) {
	ImageRgbaU8 image = b200_image_create<uint32_t>(width, height, order);
	if (zeroed) { dfpsr_image pod = image.pod(); b200_check(dfpsr_image_fill_rgba(&pod, 0, 0, 0, 0, b200_stream())); image.touchedByDevice(); }
	return image;
}
inline ImageF32 image_create_F32(int32_t width, int32_t height, bool zeroed = true) {
	ImageF32 image = b200_image_create<float>(width, height, PackOrderIndex::RGBA);
	if (zeroed) { dfpsr_image pod = image.pod(); b200_check(dfpsr_image_fill_f32(&pod, 0.0f, b200_stream())); image.touchedByDevice(); }
	return image;
}
template <typename P> inline bool image_exists(const B200Image<P> &image) { return (bool)image.buffer; }          // ref: api/imageAPI.h:96
template <typename P> inline int32_t image_getWidth(const B200Image<P> &image) { return image.width; }            // ref: api/imageAPI.h:104-120
template <typename P> inline int32_t image_getHeight(const B200Image<P> &image) { return image.height; }
template <typename P> inline int32_t image_getStride(const B200Image<P> &image) { return image.stride; }
template <typename P> inline bool image_isSubImage(const B200Image<P> &image) { return image.subImage; }
inline PackOrderIndex image_getPackOrderIndex(const ImageRgbaU8 &image) { return image.packOrder; }
// ref: api/imageAPI.h:380-400 image_getSubImage — a view on the same buffer
template <typename P> inline B200Image<P> image_getSubImage(const B200Image<P> &image, int32_t left, int32_t top, int32_t width, int32_t height) {
	if (!image_exists(image)) { return B200Image<P>(); }
	if (left < 0 || top < 0 || width <= 0 || height <= 0 || left + width > image.width || top + height > image.height) { return B200Image<P>(); } // ref: Image.h:186-206
	B200Image<P> r = image;
	r.startOffset = image.startOffset + top * image.stride + left * (int32_t)sizeof(P);
	r.width = width; r.height = height; r.subImage = true;
	return r;
}
// ref: api/imageAPI.cpp:167-185
inline void image_fill(ImageRgbaU8 &image, const ColorRgbaI32 &color) {
	if (!image_exists(image)) { return; }
	dfpsr_image pod = image.pod();
	b200_check(dfpsr_image_fill_rgba(&pod, color.red, color.green, color.blue, color.alpha, b200_stream()));
	image.touchedByDevice();
}
inline void image_fill(ImageF32 &image, float value) {
	if (!image_exists(image)) { return; }
	dfpsr_image pod = image.pod();
	b200_check(dfpsr_image_fill_f32(&pod, value, b200_stream()));
	image.touchedByDevice();
}
// ref: api/imageAPI.h:207-341 — reads go through the host mirror
inline ColorRgbaI32 b200_unpack(uint32_t c, PackOrderIndex order) {
	static const int index[4][4] = {{0, 1, 2, 3}, {2, 1, 0, 3}, {1, 2, 3, 0}, {3, 2, 1, 0}}; // byte of r, g, b, a (ref: PackOrder.h:85-96)
	const int *i = index[(int)order];
	return ColorRgbaI32((int32_t)((c >> (8 * i[0])) & 255u), (int32_t)((c >> (8 * i[1])) & 255u), (int32_t)((c >> (8 * i[2])) & 255u), (int32_t)((c >> (8 * i[3])) & 255u));
}
inline ColorRgbaI32 image_readPixel_clamp(const ImageRgbaU8 &image, int32_t x, int32_t y) {
	if (!image_exists(image)) { return ColorRgbaI32(); }
	x = x < 0 ? 0 : (x >= image.width ? image.width - 1 : x); y = y < 0 ? 0 : (y >= image.height ? image.height - 1 : y);
	uint32_t c;
	std::memcpy(&c, image.buffer->hostData() + image.startOffset + (size_t)y * image.stride + (size_t)x * 4, 4);
	return b200_unpack(c, image.packOrder);
}
// ref: api/imageAPI.h:228-275 image_readPixel_border (transparent black or the given colour outside) and image_readPixel_tile (wrap around)
inline ColorRgbaI32 image_readPixel_border(const ImageRgbaU8 &image, int32_t x, int32_t y, const ColorRgbaI32 &border = ColorRgbaI32()) {
	if (!image_exists(image) || x < 0 || y < 0 || x >= image.width || y >= image.height) { return border; }
	return image_readPixel_clamp(image, x, y);
}
inline ColorRgbaI32 image_readPixel_tile(const ImageRgbaU8 &image, int32_t x, int32_t y) {
	if (!image_exists(image)) { return ColorRgbaI32(); }
	auto wrap = [](int32_t v, int32_t size) { int32_t m = v % size; return m < 0 ? m + size : m; };
	return image_readPixel_clamp(image, wrap(x, image.width), wrap(y, image.height));
}
inline float image_readPixel_clamp(const ImageF32 &image, int32_t x, int32_t y) {
	if (!image_exists(image)) { return 0.0f; }
	x = x < 0 ? 0 : (x >= image.width ? image.width - 1 : x); y = y < 0 ? 0 : (y >= image.height ? image.height - 1 : y);
	float v;
	std::memcpy(&v, image.buffer->hostData() + image.startOffset + (size_t)y * image.stride + (size_t)x * 4, 4);
	return v;
}
// ref: api/imageAPI.h:184-204 image_writePixel — saturated, silently ignored outside of the image; written through to the device
inline uint32_t b200_pack(const ColorRgbaI32 &color, PackOrderIndex order) {
	static const int index[4][4] = {{0, 1, 2, 3}, {2, 1, 0, 3}, {1, 2, 3, 0}, {3, 2, 1, 0}};
	const int *i = index[(int)order];
	auto sat = [](int32_t v) { return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
	return (sat(color.red) << (8 * i[0])) | (sat(color.green) << (8 * i[1])) | (sat(color.blue) << (8 * i[2])) | (sat(color.alpha) << (8 * i[3]));
}
template <typename P> inline void b200_write_pixel(const B200Image<P> &image, int32_t x, int32_t y, const void *value) {
	if (!image_exists(image) || x < 0 || y < 0 || x >= image.width || y >= image.height) { return; }
	const size_t offset = (size_t)image.startOffset + (size_t)y * image.stride + (size_t)x * sizeof(P);
	if (image.buffer->hostValid && !image.buffer->deviceDirty) { std::memcpy(image.buffer->host.data() + offset, value, sizeof(P)); } // keep a valid mirror valid
	b200_check(dfpsr_upload((uint8_t *)image.buffer->device + offset, value, sizeof(P), b200_stream()));
	b200_check(dfpsr_stream_synchronize(b200_stream())); // `value` is the caller's stack
}
inline void image_writePixel(const ImageRgbaU8 &image, int32_t x, int32_t y, const ColorRgbaI32 &color) { const uint32_t packed = b200_pack(color, image.packOrder); b200_write_pixel(image, x, y, &packed); }
inline void image_writePixel(const ImageRgbaU8 &image, int32_t x, int32_t y, uint32_t packedColor) { b200_write_pixel(image, x, y, &packedColor); }
inline void image_writePixel(const ImageF32 &image, int32_t x, int32_t y, float color) { b200_write_pixel(image, x, y, &color); }
// Whole-image transfers for presentation / asset upload (tightly packed rows on the host side).
template <typename P> inline void image_download(const B200Image<P> &image, P *target, int32_t targetStrideBytes) {
	if (!image_exists(image)) { return; }
	b200_check(dfpsr_download_2d(target, (size_t)targetStrideBytes, (uint8_t *)image.buffer->device + image.startOffset, (size_t)image.stride, (size_t)image.width * sizeof(P), (size_t)image.height, b200_stream()));
	b200_check(dfpsr_stream_synchronize(b200_stream()));
}
template <typename P> inline void image_upload(B200Image<P> &image, const P *source, int32_t sourceStrideBytes) {
	if (!image_exists(image)) { return; }
	b200_check(dfpsr_upload_2d((uint8_t *)image.buffer->device + image.startOffset, (size_t)image.stride, source, (size_t)sourceStrideBytes, (size_t)image.width * sizeof(P), (size_t)image.height, b200_stream()));
	b200_check(dfpsr_stream_synchronize(b200_stream()));
	image.buffer->hostValid = false;
}

// ---------------------------------------------------------------- textures (ref: api/textureAPI.h:450-465, implementation/image/Texture.h)
class TextureRgbaU8 {
public:
	std::shared_ptr<B200Buffer> buffer;
	dfpsr_texture layout{};
	dfpsr_texture pod() const { dfpsr_texture r = layout; r.data = buffer ? (const uint32_t *)buffer->device : nullptr; return r; }
};
inline bool texture_exists(const TextureRgbaU8 &texture) { return (bool)texture.buffer; }
inline int32_t texture_getMaxWidth(const TextureRgbaU8 &texture) { return texture_exists(texture) ? 1 << texture.layout.log2width : 0; }
inline int32_t texture_getMaxHeight(const TextureRgbaU8 &texture) { return texture_exists(texture) ? 1 << texture.layout.log2height : 0; }
inline int32_t texture_getSmallestMipLevel(const TextureRgbaU8 &texture) { return (int32_t)texture.layout.maxMipLevel; }
// ref: api/textureAPI.cpp:65-78 texture_create_RgbaU8(width, height, resolutions)
inline TextureRgbaU8 texture_create_RgbaU8(int32_t width, int32_t height, int32_t resolutions) {
	TextureRgbaU8 t;
	b200_check(dfpsr_texture_layout(&t.layout, width, height, resolutions));
	t.buffer = std::make_shared<B200Buffer>((size_t)t.layout.totalPixels * 4);
	return t;
}
// ref: api/textureAPI.cpp:80-87 texture_generatePyramid
inline void texture_generatePyramid(TextureRgbaU8 &texture) {
	if (!texture_exists(texture)) { return; }
	dfpsr_texture pod = texture.pod();
	b200_check(dfpsr_texture_generate_pyramid(&pod, b200_stream()));
	texture.buffer->deviceDirty = true;
}
// ref: api/textureAPI.cpp:89-110 texture_create_RgbaU8(image, resolutions): bilinear resize to powers of two, then the pyramid
inline TextureRgbaU8 texture_create_RgbaU8(const ImageRgbaU8 &image, int32_t resolutions) {
	if (!image_exists(image)) { return TextureRgbaU8(); }
	TextureRgbaU8 t = texture_create_RgbaU8(image.width, image.height, resolutions);
	dfpsr_texture pod = t.pod();
	dfpsr_image source = image.pod();
	b200_check(dfpsr_texture_from_image(&pod, &source, b200_stream()));
	t.buffer->deviceDirty = true;
	return t;
}
// ref: api/textureAPI.h:460-465 texture_getMipLevelImage — a view on one level (level 0 = full resolution)
inline ImageRgbaU8 texture_getMipLevelImage(const TextureRgbaU8 &texture, int32_t mipLevel) {
	if (!texture_exists(texture) || mipLevel < 0 || mipLevel > (int32_t)texture.layout.maxMipLevel) { return ImageRgbaU8(); }
	ImageRgbaU8 image;
	image.buffer = texture.buffer;
	image.width = (1 << texture.layout.log2width) >> mipLevel; image.height = (1 << texture.layout.log2height) >> mipLevel;
	image.stride = image.width * 4;
	image.startOffset = (int32_t)(texture.layout.startOffset & (texture.layout.maxLevelMask >> (2 * mipLevel))) * 4;
	image.subImage = true;
	return image;
}

// ---------------------------------------------------------------- camera (ref: implementation/render/Camera.h:128-217)
class Camera {
public:
	dfpsr_camera pod{};
	static Camera createPerspective(const Transform3D &location, float imageWidth, float imageHeight, float widthSlope = 1.0f, float nearClip = 0.01f, float farClip = 1000.0f) {
		Camera c; dfpsr_transform3d t = b200_pod(location);
		b200_check(dfpsr_camera_create_perspective(&c.pod, &t, imageWidth, imageHeight, widthSlope, nearClip, farClip));
		return c;
	}
	static Camera createOrthogonal(const Transform3D &location, float imageWidth, float imageHeight, float halfWidth) {
		Camera c; dfpsr_transform3d t = b200_pod(location);
		b200_check(dfpsr_camera_create_orthogonal(&c.pod, &t, imageWidth, imageHeight, halfWidth));
		return c;
	}
	// ref: Camera.h:202-217 — 0 hidden, 1 partial, 2 full
	int isBoxSeen(const FVector3D &minimum, const FVector3D &maximum, const Transform3D &modelToWorld) const {
		float mn[3] = {minimum.x, minimum.y, minimum.z}, mx[3] = {maximum.x, maximum.y, maximum.z};
		dfpsr_transform3d t = b200_pod(modelToWorld);
		return dfpsr_camera_is_box_seen(&pod, mn, mx, &t);
	}
};

// ---------------------------------------------------------------- models (ref: api/modelAPI.h:62-301, implementation/render/model/Model.h:54-84)
struct B200Part {
	std::string name;
	TextureRgbaU8 diffuseMap, lightMap;
	std::vector<dfpsr_polygon> polygons;
	std::shared_ptr<B200Buffer> devicePolygons;
	bool dirty = true;
};
struct B200Model {
	Filter filter = Filter::Solid;
	std::vector<float> points; // x, y, z
	std::vector<B200Part> parts;
	FVector3D minBound, maxBound; // ref: Model.cpp:281-288 — starts at the origin and only grows
	std::shared_ptr<B200Buffer> devicePoints;
	bool pointsDirty = true;
};
using Model = std::shared_ptr<B200Model>;

inline Model model_create() { return std::make_shared<B200Model>(); }
inline bool model_exists(const Model &model) { return (bool)model; }
inline Model model_clone(const Model &model) { // ref: api/modelAPI.h:66 — deep copy of geometry, textures shared
	if (!model) { return Model(); }
	Model r = std::make_shared<B200Model>(*model);
	r->devicePoints.reset(); r->pointsDirty = true;
	for (B200Part &p : r->parts) { p.devicePolygons.reset(); p.dirty = true; }
	return r;
}
inline void b200_require(const Model &model, const char *method) { if (!model) { throwError(std::string(method) + ": the model does not exist"); } }
inline void model_setFilter(const Model &model, Filter filter) { b200_require(model, "model_setFilter"); model->filter = filter; }
inline Filter model_getFilter(const Model &model) { b200_require(model, "model_getFilter"); return model->filter; }
inline int32_t model_addEmptyPart(Model &model, const std::string &name) {
	b200_require(model, "model_addEmptyPart");
	model->parts.emplace_back();
	model->parts.back().name = name;
	return (int32_t)model->parts.size() - 1;
}
inline int32_t model_getNumberOfParts(const Model &model) { b200_require(model, "model_getNumberOfParts"); return (int32_t)model->parts.size(); }
inline int32_t model_getNumberOfPoints(const Model &model) { b200_require(model, "model_getNumberOfPoints"); return (int32_t)(model->points.size() / 3); }
inline int32_t model_addPoint(const Model &model, const FVector3D &position) {
	b200_require(model, "model_addPoint");
	model->points.push_back(position.x); model->points.push_back(position.y); model->points.push_back(position.z);
	// ref: Model.cpp:281-288 expandBound
	if (position.x < model->minBound.x) { model->minBound.x = position.x; } if (position.y < model->minBound.y) { model->minBound.y = position.y; } if (position.z < model->minBound.z) { model->minBound.z = position.z; }
	if (position.x > model->maxBound.x) { model->maxBound.x = position.x; } if (position.y > model->maxBound.y) { model->maxBound.y = position.y; } if (position.z > model->maxBound.z) { model->maxBound.z = position.z; }
	model->pointsDirty = true;
	return (int32_t)(model->points.size() / 3) - 1;
}
inline FVector3D model_getPoint(const Model &model, int32_t pointIndex) {
	b200_require(model, "model_getPoint");
	if (pointIndex < 0 || pointIndex >= model_getNumberOfPoints(model)) { return FVector3D(); } // ref: Model.cpp:34-42 prints and returns a default
	return FVector3D(model->points[3 * pointIndex], model->points[3 * pointIndex + 1], model->points[3 * pointIndex + 2]);
}
inline void model_setPoint(Model &model, int32_t pointIndex, const FVector3D &position) {
	b200_require(model, "model_setPoint");
	if (pointIndex < 0 || pointIndex >= model_getNumberOfPoints(model)) { return; }
	model->points[3 * pointIndex] = position.x; model->points[3 * pointIndex + 1] = position.y; model->points[3 * pointIndex + 2] = position.z;
	if (position.x < model->minBound.x) { model->minBound.x = position.x; } if (position.y < model->minBound.y) { model->minBound.y = position.y; } if (position.z < model->minBound.z) { model->minBound.z = position.z; }
	if (position.x > model->maxBound.x) { model->maxBound.x = position.x; } if (position.y > model->maxBound.y) { model->maxBound.y = position.y; } if (position.z > model->maxBound.z) { model->maxBound.z = position.z; }
	model->pointsDirty = true;
}
inline void model_getBoundingBox(const Model &model, FVector3D &minimum, FVector3D &maximum) { b200_require(model, "model_getBoundingBox"); minimum = model->minBound; maximum = model->maxBound; }
inline B200Part *b200_part(const Model &model, int32_t partIndex, const char *method) {
	b200_require(model, method);
	if (partIndex < 0 || partIndex >= (int32_t)model->parts.size()) { return nullptr; }
	return &model->parts[(size_t)partIndex];
}
inline int32_t b200_add_polygon(Model &model, int32_t partIndex, int32_t a, int32_t b, int32_t c, int32_t d, const char *method) {
	B200Part *part = b200_part(model, partIndex, method);
	if (!part) { return -1; }
	dfpsr_polygon polygon;
	std::memset(&polygon, 0, sizeof(polygon));
	polygon.pointIndices[0] = a; polygon.pointIndices[1] = b; polygon.pointIndices[2] = c; polygon.pointIndices[3] = d;
	// ref: Model.cpp:74-103 Polygon(indexA, indexB, indexC[, indexD]): white corners, texture coordinates spanning the whole texture
	static const float cornerTexCoords[4][4] = {{0.0f, 0.0f, 0.0f, 0.0f}, {1.0f, 0.0f, 1.0f, 0.0f}, {1.0f, 1.0f, 1.0f, 1.0f}, {0.0f, 1.0f, 0.0f, 1.0f}};
	for (int v = 0; v < 4; v++) { for (int ch = 0; ch < 4; ch++) { polygon.colors[v][ch] = 1.0f; polygon.texCoords[v][ch] = cornerTexCoords[v][ch]; } }
	part->polygons.push_back(polygon);
	part->dirty = true;
	return (int32_t)part->polygons.size() - 1;
}
inline int32_t model_addTriangle(Model &model, int32_t partIndex, int32_t pointA, int32_t pointB, int32_t pointC) { return b200_add_polygon(model, partIndex, pointA, pointB, pointC, -1, "model_addTriangle"); }
inline int32_t model_addQuad(Model &model, int32_t partIndex, int32_t pointA, int32_t pointB, int32_t pointC, int32_t pointD) { return b200_add_polygon(model, partIndex, pointA, pointB, pointC, pointD, "model_addQuad"); }
inline int32_t model_getNumberOfPolygons(const Model &model, int32_t partIndex) { B200Part *p = b200_part(model, partIndex, "model_getNumberOfPolygons"); return p ? (int32_t)p->polygons.size() : 0; }
inline dfpsr_polygon *b200_polygon(const Model &model, int32_t partIndex, int32_t polygonIndex, int32_t vertexIndex, const char *method) {
	B200Part *part = b200_part(model, partIndex, method);
	if (!part || polygonIndex < 0 || polygonIndex >= (int32_t)part->polygons.size() || vertexIndex < 0 || vertexIndex > 3) { return nullptr; }
	part->dirty = true;
	return &part->polygons[(size_t)polygonIndex];
}
inline void model_setVertexColor(Model &model, int32_t partIndex, int32_t polygonIndex, int32_t vertexIndex, const FVector4D &color) {
	if (dfpsr_polygon *p = b200_polygon(model, partIndex, polygonIndex, vertexIndex, "model_setVertexColor")) { p->colors[vertexIndex][0] = color.x; p->colors[vertexIndex][1] = color.y; p->colors[vertexIndex][2] = color.z; p->colors[vertexIndex][3] = color.w; }
}
inline void model_setTexCoord(Model &model, int32_t partIndex, int32_t polygonIndex, int32_t vertexIndex, const FVector4D &texCoord) {
	if (dfpsr_polygon *p = b200_polygon(model, partIndex, polygonIndex, vertexIndex, "model_setTexCoord")) { p->texCoords[vertexIndex][0] = texCoord.x; p->texCoords[vertexIndex][1] = texCoord.y; p->texCoords[vertexIndex][2] = texCoord.z; p->texCoords[vertexIndex][3] = texCoord.w; }
}
inline void model_setDiffuseMap(Model &model, int32_t partIndex, const TextureRgbaU8 &diffuseMap) { if (B200Part *p = b200_part(model, partIndex, "model_setDiffuseMap")) { p->diffuseMap = diffuseMap; } }
inline void model_setLightMap(Model &model, int32_t partIndex, const TextureRgbaU8 &lightMap) { if (B200Part *p = b200_part(model, partIndex, "model_setLightMap")) { p->lightMap = lightMap; } }
inline TextureRgbaU8 model_getDiffuseMap(const Model &model, int32_t partIndex) { B200Part *p = b200_part(model, partIndex, "model_getDiffuseMap"); return p ? p->diffuseMap : TextureRgbaU8(); }
inline TextureRgbaU8 model_getLightMap(const Model &model, int32_t partIndex) { B200Part *p = b200_part(model, partIndex, "model_getLightMap"); return p ? p->lightMap : TextureRgbaU8(); }

// Uploads whatever changed since the last draw and returns one dfpsr_model per part (the C ABI's model has one part).
// ---------------------------------------------------------------- importers (ref: SDK/SpriteEngine/importer.h, api/modelAPI.h:303-319)
// Textures are resolved by the caller: `textureNames`, when given, receives per part {diffuse name, light name} ("" when the shader has none);
// the reference's ResourcePool decodes image files, which is outside the path.
inline Model b200_model_from_import(dfpsr_imported_model &imported, std::vector<std::pair<std::string, std::string>> *textureNames) {
	Model model = model_create();
	model->filter = imported.filter == DFPSR_FILTER_ALPHA ? Filter::Alpha : Filter::Solid;
	model->points.assign(imported.points, imported.points + 3 * (size_t)imported.pointCount);
	model->minBound = FVector3D(imported.minBound[0], imported.minBound[1], imported.minBound[2]);
	model->maxBound = FVector3D(imported.maxBound[0], imported.maxBound[1], imported.maxBound[2]);
	for (int32_t p = 0; p < imported.partCount; p++) {
		const dfpsr_imported_part &part = imported.parts[p];
		model->parts.emplace_back();
		model->parts.back().name = part.name;
		model->parts.back().polygons.assign(imported.polygons + part.firstPolygon, imported.polygons + part.firstPolygon + part.polygonCount);
		if (textureNames) { textureNames->push_back({part.diffuseName, part.lightName}); }
	}
	dfpsr_import_free(&imported);
	return model;
}
// ref: SDK/SpriteEngine/importer.cpp:52-262 loadPlyModel on file content
inline Model importer_loadModelFromContent(const std::string &plyContent, bool flipX, const Transform3D &axisConversion) {
	dfpsr_imported_model imported;
	const dfpsr_transform3d axis = b200_pod(axisConversion);
	b200_check(dfpsr_import_ply(plyContent.data(), plyContent.size(), flipX ? 1 : 0, &axis, &imported));
	return b200_model_from_import(imported, nullptr);
}
// ref: SDK/SpriteEngine/importer.cpp:280-290 importer_loadModel(filename, flipX, axisConversion); only PLY, like the reference
inline Model importer_loadModel(const std::string &filename, bool flipX, const Transform3D &axisConversion) {
	const size_t dot = filename.find_last_of('.');
	if (dot == std::string::npos) { throwError("The model's filename " + filename + " does not have an extension!"); }
	std::string extension = filename.substr(dot + 1);
	for (char &c : extension) { if (c >= 'a' && c <= 'z') { c = (char)(c - 'a' + 'A'); } }
	if (extension != "PLY") { throwError("The extension " + extension + " in " + filename + " is not yet supported!"); }
	FILE *file = std::fopen(filename.c_str(), "rb");
	if (!file) { throwError("Failed to load " + filename); }
	std::string content;
	char block[65536];
	size_t got;
	while ((got = std::fread(block, 1, sizeof(block), file)) > 0) { content.append(block, got); }
	std::fclose(file);
	return importer_loadModelFromContent(content, flipX, axisConversion);
}
// ref: api/modelAPI.h:319 importFromContent_DMF1(fileContent, pool, detailLevel = 2)
inline Model importFromContent_DMF1(const std::string &fileContent, int32_t detailLevel = 2, std::vector<std::pair<std::string, std::string>> *textureNames = nullptr) {
	dfpsr_imported_model imported;
	b200_check(dfpsr_import_dmf1(fileContent.data(), fileContent.size(), detailLevel, &imported));
	return b200_model_from_import(imported, textureNames);
}

inline std::vector<dfpsr_model> b200_device_models(const Model &model) {
	std::vector<dfpsr_model> result;
	if (!model || model->points.empty()) { return result; }
	// A task only records device pointers (projection and set-up run at renderer_end), so geometry that a renderer still holds on to —
	// use_count() > 1: B200Renderer keeps every buffer of its queued and in-flight tasks — is never overwritten in place: the changed
	// model gets a fresh buffer and the queued task keeps drawing what was submitted, like the reference's copy at renderer_giveTask.
	if (model->pointsDirty || !model->devicePoints || model->devicePoints->bytes < model->points.size() * 4) {
		if (!model->devicePoints || model->devicePoints->bytes < model->points.size() * 4 || model->devicePoints.use_count() > 1) { model->devicePoints = std::make_shared<B200Buffer>(model->points.size() * 4); }
		b200_check(dfpsr_upload(model->devicePoints->device, model->points.data(), model->points.size() * 4, b200_stream()));
		b200_check(dfpsr_stream_synchronize(b200_stream()));
		model->pointsDirty = false;
	}
	for (B200Part &part : model->parts) {
		if (part.polygons.empty()) { continue; }
		size_t bytes = part.polygons.size() * sizeof(dfpsr_polygon);
		if (part.dirty || !part.devicePolygons || part.devicePolygons->bytes < bytes) {
			if (!part.devicePolygons || part.devicePolygons->bytes < bytes || part.devicePolygons.use_count() > 1) { part.devicePolygons = std::make_shared<B200Buffer>(bytes); }
			b200_check(dfpsr_upload(part.devicePolygons->device, part.polygons.data(), bytes, b200_stream()));
			b200_check(dfpsr_stream_synchronize(b200_stream()));
			part.dirty = false;
		}
		dfpsr_model m;
		std::memset(&m, 0, sizeof(m));
		m.points = (const float *)model->devicePoints->device; m.pointCount = (int32_t)(model->points.size() / 3);
		m.polygons = (const dfpsr_polygon *)part.devicePolygons->device; m.polygonCount = (int32_t)part.polygons.size();
		m.filter = (int32_t)model->filter;
		m.diffuse = part.diffuseMap.pod(); m.light = part.lightMap.pod();
		m.minBound[0] = model->minBound.x; m.minBound[1] = model->minBound.y; m.minBound[2] = model->minBound.z;
		m.maxBound[0] = model->maxBound.x; m.maxBound[1] = model->maxBound.y; m.maxBound[2] = model->maxBound.z;
		result.push_back(m);
	}
	return result;
}

// ---------------------------------------------------------------- renderer (ref: api/rendererAPI.h:54-135, api/rendererAPI.cpp:141-168, :352-402)
struct B200Renderer {
	dfpsr_renderer *handle = nullptr;
	ImageRgbaU8 colorBuffer;
	ImageF32 depthBuffer;
	bool receiving = false;
	// Every device buffer the queued tasks point at (points, polygons, textures, targets), and those of the frame before: the library reads
	// them at renderer_end and, for an asynchronous renderer, possibly once more when it verifies the frame at the next renderer_begin.
	// A model or texture that its owner drops in between therefore stays alive for as long as a task can read it.
	std::vector<std::shared_ptr<B200Buffer>> queued, inFlight;
	void keep(const std::shared_ptr<B200Buffer> &buffer) { if (buffer) { queued.push_back(buffer); } }
	B200Renderer() { b200_check(dfpsr_renderer_create(&handle)); b200_check(dfpsr_renderer_set_async(handle, 1)); } // every consumer of a shim image goes through the library, which verifies frames in flight first
	~B200Renderer() { if (handle) { dfpsr_renderer_destroy(handle); } }
	B200Renderer(const B200Renderer &) = delete;
	B200Renderer &operator=(const B200Renderer &) = delete;
};
using Renderer = std::shared_ptr<B200Renderer>;

inline Renderer renderer_create() { return std::make_shared<B200Renderer>(); }
inline bool renderer_exists(const Renderer &renderer) { return (bool)renderer; }
inline void renderer_begin(Renderer &renderer, ImageRgbaU8 &colorBuffer, ImageF32 &depthBuffer) {
	if (!renderer) { throwError("renderer_begin: renderer does not exist"); }
	dfpsr_image color = colorBuffer.pod(), depth = depthBuffer.pod();
	b200_check(dfpsr_renderer_begin(renderer->handle, &color, &depth)); // "twice without ending" is reported by the library (rendererAPI.cpp:152-154)
	renderer->inFlight.clear(); // the previous frame has been verified by the call above; releasing a buffer waits for the device (cudaFree)
	renderer->colorBuffer = colorBuffer; renderer->depthBuffer = depthBuffer; renderer->receiving = true;
	renderer->keep(colorBuffer.buffer); renderer->keep(depthBuffer.buffer);
}
inline ImageRgbaU8 renderer_getColorBuffer(const Renderer &renderer) { return renderer ? renderer->colorBuffer : ImageRgbaU8(); }
inline ImageF32 renderer_getDepthBuffer(const Renderer &renderer) { return renderer ? renderer->depthBuffer : ImageF32(); }
inline bool renderer_takesTriangles(const Renderer &renderer) { return renderer && renderer->receiving; }
// ref: api/modelAPI.cpp:214-281 model_render_threaded == renderer_giveTask
inline void model_render_threaded(const Model &model, const Transform3D &modelToWorldTransform, Renderer &renderer, const Camera &camera) {
	if (!renderer) { throwError("renderer_giveTask: renderer does not exist"); }
	if (!model) { return; }
	dfpsr_transform3d t = b200_pod(modelToWorldTransform);
	for (const dfpsr_model &m : b200_device_models(model)) { b200_check(dfpsr_renderer_give_task(renderer->handle, &m, &t, &camera.pod, b200_stream())); }
	renderer->keep(model->devicePoints);
	for (const B200Part &part : model->parts) { renderer->keep(part.devicePolygons); renderer->keep(part.diffuseMap.buffer); renderer->keep(part.lightMap.buffer); }
}
inline void renderer_giveTask(Renderer &renderer, const Model &model, const Transform3D &modelToWorldTransform, const Camera &camera) { model_render_threaded(model, modelToWorldTransform, renderer, camera); }
// Many models at once (no counterpart in the reference): the per-model tests of model_render_threaded — isBoxSeen and, with occluders, renderer_isBoxVisible —
// run on the device instead of one after the other on the host; the frame is identical to a loop of renderer_giveTask.
inline void renderer_giveTasks(Renderer &renderer, const std::vector<Model> &models, const std::vector<Transform3D> &modelToWorldTransforms, const Camera &camera) {
	if (!renderer) { throwError("renderer_giveTasks: renderer does not exist"); }
	if (models.size() != modelToWorldTransforms.size()) { throwError("renderer_giveTasks: one transform per model"); }
	std::vector<dfpsr_model> parts;
	std::vector<dfpsr_transform3d> transforms;
	for (size_t i = 0; i < models.size(); i++) {
		if (!models[i]) { continue; }
		for (const dfpsr_model &m : b200_device_models(models[i])) { parts.push_back(m); transforms.push_back(b200_pod(modelToWorldTransforms[i])); }
		renderer->keep(models[i]->devicePoints);
		for (const B200Part &part : models[i]->parts) { renderer->keep(part.devicePolygons); renderer->keep(part.diffuseMap.buffer); renderer->keep(part.lightMap.buffer); }
	}
	if (!parts.empty()) { b200_check(dfpsr_renderer_give_tasks(renderer->handle, parts.data(), transforms.data(), (int32_t)parts.size(), &camera.pod, b200_stream())); }
}
// ref: api/rendererAPI.h:108-129 renderer_giveTask_triangle with points the caller projected (ProjectedPoint, 40 bytes)
inline void renderer_giveTask_triangle(Renderer &renderer, const dfpsr_projected_point &posA, const dfpsr_projected_point &posB, const dfpsr_projected_point &posC,
  const FVector4D &colorA, const FVector4D &colorB, const FVector4D &colorC, const FVector4D &texCoordA, const FVector4D &texCoordB, const FVector4D &texCoordC,
  const TextureRgbaU8 &diffuse, const TextureRgbaU8 &light, Filter filter, const Camera &camera) {
	if (!renderer) { throwError("renderer_giveTask_triangle: renderer does not exist"); }
	dfpsr_triangle tri;
	tri.pos[0] = posA; tri.pos[1] = posB; tri.pos[2] = posC;
	const FVector4D *colors[3] = {&colorA, &colorB, &colorC}, *tex[3] = {&texCoordA, &texCoordB, &texCoordC};
	for (int k = 0; k < 3; k++) {
		tri.colors[k][0] = colors[k]->x; tri.colors[k][1] = colors[k]->y; tri.colors[k][2] = colors[k]->z; tri.colors[k][3] = colors[k]->w;
		tri.texCoords[k][0] = tex[k]->x; tri.texCoords[k][1] = tex[k]->y; tri.texCoords[k][2] = tex[k]->z; tri.texCoords[k][3] = tex[k]->w;
	}
	dfpsr_texture d = diffuse.pod(), l = light.pod();
	b200_check(dfpsr_renderer_give_task_triangles(renderer->handle, &tri, 1, &d, &l, (int32_t)filter, &camera.pod, b200_stream()));
	renderer->keep(diffuse.buffer); renderer->keep(light.buffer);
}
// ref: api/rendererAPI.h:73-97, :131 — the occlusion grid
inline void renderer_occludeFromBox(Renderer &renderer, const FVector3D &minimum, const FVector3D &maximum, const Transform3D &modelToWorldTransform, const Camera &camera, bool debugSilhouette = false) {
	(void)debugSilhouette;
	if (!renderer) { throwError("renderer_occludeFromBox: renderer does not exist"); }
	float mn[3] = {minimum.x, minimum.y, minimum.z}, mx[3] = {maximum.x, maximum.y, maximum.z};
	dfpsr_transform3d t = b200_pod(modelToWorldTransform);
	b200_check(dfpsr_renderer_occlude_from_box(renderer->handle, mn, mx, &t, &camera.pod));
}
inline void renderer_occludeFromTopRows(Renderer &renderer, const Camera &camera) {
	if (!renderer) { throwError("renderer_occludeFromTopRows: renderer does not exist"); }
	b200_check(dfpsr_renderer_occlude_from_top_rows(renderer->handle, &camera.pod, b200_stream()));
}
inline void renderer_occludeFromExistingTriangles(Renderer &renderer) {
	if (!renderer) { throwError("renderer_occludeFromExistingTriangles: renderer does not exist"); }
	b200_check(dfpsr_renderer_occlude_from_existing_triangles(renderer->handle, b200_stream()));
}
inline bool renderer_hasOccluders(const Renderer &renderer) { return renderer && dfpsr_renderer_has_occluders(renderer->handle) != 0; }
inline bool renderer_isBoxVisible(const Renderer &renderer, const FVector3D &minimum, const FVector3D &maximum, const Transform3D &modelToWorldTransform, const Camera &camera) {
	if (!renderer) { throwError("renderer_isBoxVisible: renderer does not exist"); }
	float mn[3] = {minimum.x, minimum.y, minimum.z}, mx[3] = {maximum.x, maximum.y, maximum.z};
	dfpsr_transform3d t = b200_pod(modelToWorldTransform);
	int32_t visible = 0;
	b200_check(dfpsr_renderer_is_box_visible(renderer->handle, mn, mx, &t, &camera.pod, &visible));
	return visible != 0;
}
inline void renderer_end(Renderer &renderer, bool debugWireframe = false) {
	if (!renderer) { throwError("renderer_end: renderer does not exist"); }
	if (debugWireframe) { b200_check(dfpsr_renderer_set_debug_wireframe(renderer->handle, 1)); } // ref: rendererAPI.cpp:362-399, drawn on the device after the frame
	b200_check(dfpsr_renderer_end(renderer->handle, b200_stream()));
	renderer->inFlight.swap(renderer->queued); renderer->queued.clear();
	renderer->colorBuffer.touchedByDevice(); renderer->depthBuffer.touchedByDevice();
	renderer->colorBuffer = ImageRgbaU8(); renderer->depthBuffer = ImageF32(); renderer->receiving = false; // ref: rendererAPI.cpp:480-488
}
// ref: api/modelAPI.cpp:197-206
inline void model_render(const Model &model, const Transform3D &modelToWorldTransform, ImageRgbaU8 &colorBuffer, ImageF32 &depthBuffer, const Camera &camera) {
	if (!model) { return; }
	dfpsr_transform3d t = b200_pod(modelToWorldTransform);
	dfpsr_image color = colorBuffer.pod(), depth = depthBuffer.pod();
	for (const dfpsr_model &m : b200_device_models(model)) { b200_check(dfpsr_model_render(&m, &t, &color, &depth, &camera.pod, b200_stream())); }
	colorBuffer.touchedByDevice(); depthBuffer.touchedByDevice();
}
inline void model_renderDepth(const Model &model, const Transform3D &modelToWorldTransform, ImageF32 &depthBuffer, const Camera &camera) {
	if (!model) { return; }
	dfpsr_transform3d t = b200_pod(modelToWorldTransform);
	dfpsr_image depth = depthBuffer.pod();
	for (const dfpsr_model &m : b200_device_models(model)) { b200_check(dfpsr_model_render_depth(&m, &t, &depth, &camera.pod, b200_stream())); }
	depthBuffer.touchedByDevice();
}

// ---------------------------------------------------------------- 2D draw calls on the path (ref: api/drawAPI.h)
inline void draw_copy(ImageRgbaU8 &target, const ImageRgbaU8 &source, int32_t left = 0, int32_t top = 0) {
	dfpsr_image t = target.pod(), s = source.pod();
	b200_check(dfpsr_draw_copy_rgba(&t, &s, left, top, b200_stream())); target.touchedByDevice();
}
inline void draw_copy(ImageF32 &target, const ImageF32 &source, int32_t left = 0, int32_t top = 0) {
	dfpsr_image t = target.pod(), s = source.pod();
	b200_check(dfpsr_draw_copy_f32(&t, &s, left, top, b200_stream())); target.touchedByDevice();
}
// ref: api/drawAPI.cpp:962-979 draw_higher(F32 height, + 0 / 1 / 2 RGBA payloads)
inline void draw_higher(ImageF32 &targetHeight, const ImageF32 &sourceHeight, int32_t left = 0, int32_t top = 0, float sourceHeightOffset = 0.0f) {
	dfpsr_image th = targetHeight.pod(), sh = sourceHeight.pod();
	b200_check(dfpsr_draw_higher(&th, &sh, nullptr, nullptr, nullptr, nullptr, left, top, sourceHeightOffset, b200_stream())); targetHeight.touchedByDevice();
}
inline void draw_higher(ImageF32 &targetHeight, const ImageF32 &sourceHeight, ImageRgbaU8 &targetA, const ImageRgbaU8 &sourceA, int32_t left = 0, int32_t top = 0, float sourceHeightOffset = 0.0f) {
	dfpsr_image th = targetHeight.pod(), sh = sourceHeight.pod(), ta = targetA.pod(), sa = sourceA.pod();
	b200_check(dfpsr_draw_higher(&th, &sh, &ta, &sa, nullptr, nullptr, left, top, sourceHeightOffset, b200_stream())); targetHeight.touchedByDevice(); targetA.touchedByDevice();
}
inline void draw_higher(ImageF32 &targetHeight, const ImageF32 &sourceHeight, ImageRgbaU8 &targetA, const ImageRgbaU8 &sourceA, ImageRgbaU8 &targetB, const ImageRgbaU8 &sourceB, int32_t left = 0, int32_t top = 0, float sourceHeightOffset = 0.0f) {
	dfpsr_image th = targetHeight.pod(), sh = sourceHeight.pod(), ta = targetA.pod(), sa = sourceA.pod(), tb = targetB.pod(), sb = sourceB.pod();
	b200_check(dfpsr_draw_higher(&th, &sh, &ta, &sa, &tb, &sb, left, top, sourceHeightOffset, b200_stream())); targetHeight.touchedByDevice(); targetA.touchedByDevice(); targetB.touchedByDevice();
}

// ---------------------------------------------------------------- remaining 2D draw calls (ref: api/drawAPI.h:68-161)
struct IRect { int32_t l = 0, t = 0, w = 0, h = 0; IRect() {} IRect(int32_t left, int32_t top, int32_t width, int32_t height) : l(left), t(top), w(width), h(height) {}
	int32_t left() const { return l; } int32_t top() const { return t; } int32_t width() const { return w; } int32_t height() const { return h; } int32_t right() const { return l + w; } int32_t bottom() const { return t + h; } };
using ImageU8 = B200Image<uint8_t>;
inline ImageU8 image_create_U8(int32_t width, int32_t height, bool zeroed = true) { // ref: api/imageAPI.h:47
	ImageU8 image = b200_image_create<uint8_t>(width, height, PackOrderIndex::RGBA);
	if (zeroed) { image.buffer->host.assign(image.buffer->bytes, 0); image.buffer->hostValid = true; image.buffer->pushHost(); }
	return image;
}
inline void draw_rectangle(ImageRgbaU8 &image, const IRect &bound, const ColorRgbaI32 &color) {
	dfpsr_image im = image.pod(); const int32_t c[4] = {color.red, color.green, color.blue, color.alpha};
	b200_check(dfpsr_draw_rectangle_rgba(&im, bound.l, bound.t, bound.w, bound.h, c, b200_stream())); image.touchedByDevice();
}
inline void draw_rectangle(ImageF32 &image, const IRect &bound, float color) { dfpsr_image im = image.pod(); b200_check(dfpsr_draw_rectangle_f32(&im, bound.l, bound.t, bound.w, bound.h, color, b200_stream())); image.touchedByDevice(); }
inline void draw_line(ImageRgbaU8 &image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, const ColorRgbaI32 &color) {
	dfpsr_image im = image.pod(); const int32_t c[4] = {color.red, color.green, color.blue, color.alpha};
	b200_check(dfpsr_draw_line_rgba(&im, x1, y1, x2, y2, c, b200_stream())); image.touchedByDevice();
}
inline void draw_line(ImageF32 &image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, float color) { dfpsr_image im = image.pod(); b200_check(dfpsr_draw_line_f32(&im, x1, y1, x2, y2, color, b200_stream())); image.touchedByDevice(); }
inline void draw_alphaFilter(ImageRgbaU8 &target, const ImageRgbaU8 &source, int32_t left = 0, int32_t top = 0) { dfpsr_image t = target.pod(), s = source.pod(); b200_check(dfpsr_draw_alpha_filter(&t, &s, left, top, b200_stream())); target.touchedByDevice(); }
inline void draw_maxAlpha(ImageRgbaU8 &target, const ImageRgbaU8 &source, int32_t left = 0, int32_t top = 0, int32_t sourceAlphaOffset = 0) { dfpsr_image t = target.pod(), s = source.pod(); b200_check(dfpsr_draw_max_alpha(&t, &s, left, top, sourceAlphaOffset, b200_stream())); target.touchedByDevice(); }
inline void draw_alphaClip(ImageRgbaU8 &target, const ImageRgbaU8 &source, int32_t left = 0, int32_t top = 0, int32_t threshold = 127) { dfpsr_image t = target.pod(), s = source.pod(); b200_check(dfpsr_draw_alpha_clip(&t, &s, left, top, threshold, b200_stream())); target.touchedByDevice(); }
inline void draw_silhouette(ImageRgbaU8 &target, const ImageU8 &silhouette, const ColorRgbaI32 &color, int32_t left = 0, int32_t top = 0) {
	dfpsr_image t = target.pod(), s = silhouette.pod(); const int32_t c[4] = {color.red, color.green, color.blue, color.alpha};
	b200_check(dfpsr_draw_silhouette(&t, &s, c, left, top, b200_stream())); target.touchedByDevice();
}

// ---------------------------------------------------------------- Sandbox deferred light (ref: SDK/SpriteEngine/lightAPI.h:27-31, orthoAPI.h:52-77)
struct OrthoView { dfpsr_ortho_view pod{}; };
inline void b200_light_directed(const OrthoView &view, ImageRgbaU8 &lightBuffer, const ImageRgbaU8 &normalBuffer, const FVector3D &lightDirection, float lightIntensity, const ColorRgbaI32 &lightColor, int add) {
	dfpsr_image l = lightBuffer.pod(), n = normalBuffer.pod();
	float d[3] = {lightDirection.x, lightDirection.y, lightDirection.z};
	int32_t c[3] = {lightColor.red, lightColor.green, lightColor.blue};
	b200_check(dfpsr_light_directed(&view.pod, &l, &n, d, lightIntensity, c, add, b200_stream())); lightBuffer.touchedByDevice();
}
inline void setDirectedLight(const OrthoView &camera, ImageRgbaU8 &lightBuffer, const ImageRgbaU8 &normalBuffer, const FVector3D &lightDirection, float lightIntensity, const ColorRgbaI32 &lightColor) { b200_light_directed(camera, lightBuffer, normalBuffer, lightDirection, lightIntensity, lightColor, 0); }
inline void addDirectedLight(const OrthoView &camera, ImageRgbaU8 &lightBuffer, const ImageRgbaU8 &normalBuffer, const FVector3D &lightDirection, float lightIntensity, const ColorRgbaI32 &lightColor) { b200_light_directed(camera, lightBuffer, normalBuffer, lightDirection, lightIntensity, lightColor, 1); }
inline void addPointLight(const OrthoView &camera, int32_t worldCenterX, int32_t worldCenterY, ImageRgbaU8 &lightBuffer, const ImageRgbaU8 &normalBuffer, const ImageF32 &heightBuffer, const FVector3D &lightPosition, float lightRadius, float lightIntensity, const ColorRgbaI32 &lightColor, const ImageF32 &shadowCubeMap = ImageF32()) {
	dfpsr_image l = lightBuffer.pod(), n = normalBuffer.pod(), h = heightBuffer.pod(), cube = shadowCubeMap.pod();
	float p[3] = {lightPosition.x, lightPosition.y, lightPosition.z};
	int32_t c[3] = {lightColor.red, lightColor.green, lightColor.blue}, wc[2] = {worldCenterX, worldCenterY};
	b200_check(dfpsr_light_point(&camera.pod, wc, &l, &n, &h, p, lightRadius, lightIntensity, c, image_exists(shadowCubeMap) ? &cube : nullptr, b200_stream())); lightBuffer.touchedByDevice();
}
// ref: SDK/SpriteEngine/lightAPI.h:29-30 — the reference's two signatures (world centre as IVector2D, with and without a shadow cube map)
inline void addPointLight(const OrthoView &camera, const IVector2D &worldCenter, ImageRgbaU8 &lightBuffer, const ImageRgbaU8 &normalBuffer, const ImageF32 &heightBuffer, const FVector3D &lightPosition, float lightRadius, float lightIntensity, const ColorRgbaI32 &lightColor, const ImageF32 &shadowCubeMap) {
	addPointLight(camera, worldCenter.x, worldCenter.y, lightBuffer, normalBuffer, heightBuffer, lightPosition, lightRadius, lightIntensity, lightColor, shadowCubeMap);
}
inline void addPointLight(const OrthoView &camera, const IVector2D &worldCenter, ImageRgbaU8 &lightBuffer, const ImageRgbaU8 &normalBuffer, const ImageF32 &heightBuffer, const FVector3D &lightPosition, float lightRadius, float lightIntensity, const ColorRgbaI32 &lightColor) {
	addPointLight(camera, worldCenter.x, worldCenter.y, lightBuffer, normalBuffer, heightBuffer, lightPosition, lightRadius, lightIntensity, lightColor, ImageF32());
}
inline void blendLight(ImageRgbaU8 &colorBuffer, const ImageRgbaU8 &diffuseBuffer, const ImageRgbaU8 &lightBuffer) {
	dfpsr_image c = colorBuffer.pod(), d = diffuseBuffer.pod(), l = lightBuffer.pod();
	b200_check(dfpsr_light_blend(&c, &d, &l, b200_stream())); colorBuffer.touchedByDevice();
}

// ---------------------------------------------------------------- Sandbox sprite world (ref: SDK/SpriteEngine/spriteAPI.h:24-145, orthoAPI.h:92-134)
using Direction = int32_t;
static const int32_t ortho_miniUnitsPerTile = 1024; // ref: orthoAPI.h:28
struct OrthoSystem { // ref: orthoAPI.h:92-134 — the eight views are derived exactly like OrthoSystem::update (orthoAPI.cpp:82-119)
	dfpsr_ortho_system pod{};
	OrthoSystem() {}
	OrthoSystem(float cameraTilt, int32_t pixelsPerTile) { b200_check(dfpsr_ortho_system_create(&pod, cameraTilt, pixelsPerTile)); }
	OrthoView lightView(int32_t cameraIndex) const { OrthoView v; b200_check(dfpsr_ortho_camera_light_view(&pod.view[cameraIndex], &v.pod)); return v; }
};
struct SpriteInstance { // ref: spriteAPI.h:24-36
	int32_t typeIndex; Direction direction; IVector3D location; bool shadowCasting; uint64_t userData;
	SpriteInstance(int32_t typeIndex, Direction direction, const IVector3D &location, bool shadowCasting, uint64_t userData = 0)
	: typeIndex(typeIndex), direction(direction), location(location), shadowCasting(shadowCasting), userData(userData) {}
};
struct ModelInstance { // ref: spriteAPI.h:44-52
	int32_t typeIndex; Transform3D location; uint64_t userData;
	ModelInstance(int32_t typeIndex, const Transform3D &location, uint64_t userData = 0) : typeIndex(typeIndex), location(location), userData(userData) {}
};
inline dfpsr_sprite_instance b200_pod(const SpriteInstance &s) {
	dfpsr_sprite_instance r; r.typeIndex = s.typeIndex; r.direction = s.direction; r.location[0] = s.location.x; r.location[1] = s.location.y; r.location[2] = s.location.z;
	r.shadowCasting = s.shadowCasting ? 1 : 0; r.userData = s.userData; return r;
}
inline dfpsr_model_instance b200_pod(const ModelInstance &m) { dfpsr_model_instance r; r.typeIndex = m.typeIndex; r.location = b200_pod(m.location); r.userData = m.userData; return r; }

// Sprite types. The reference loads <name>.png + <name>.ini (spriteAPI.cpp:190-232); image codecs are outside of the hot path, so the
// decoded atlas (RGBA order, host pixels) and the parsed settings are handed over instead.
struct SpriteConfig { // ref: spriteAPI.cpp:47-55
	int32_t centerX = 0, centerY = 0, frameRows = 1, propertyColumns = 3;
	FVector3D minBound, maxBound;
	std::vector<FVector3D> points; std::vector<int32_t> triangleIndices;
};
inline int32_t spriteWorld_createSpriteType(const uint32_t *atlasPixels, int32_t width, int32_t height, int32_t strideBytes, const SpriteConfig &config) {
	dfpsr_sprite_config c{};
	c.centerX = config.centerX; c.centerY = config.centerY; c.frameRows = config.frameRows; c.propertyColumns = config.propertyColumns;
	c.minBound[0] = config.minBound.x; c.minBound[1] = config.minBound.y; c.minBound[2] = config.minBound.z;
	c.maxBound[0] = config.maxBound.x; c.maxBound[1] = config.maxBound.y; c.maxBound[2] = config.maxBound.z;
	std::vector<float> flat;
	for (const FVector3D &p : config.points) { flat.push_back(p.x); flat.push_back(p.y); flat.push_back(p.z); }
	c.points = flat.data(); c.pointCount = (int32_t)config.points.size();
	c.triangleIndices = config.triangleIndices.data(); c.triangleIndexCount = (int32_t)config.triangleIndices.size();
	int32_t index = -1;
	b200_check(dfpsr_sprite_type_create(atlasPixels, width, height, strideBytes, &c, &index));
	return index;
}
inline int32_t spriteWorld_getSpriteTypeCount() { return dfpsr_sprite_type_count(); } // ref: spriteAPI.h:60
inline int32_t spriteWorld_getModelTypeCount() { return dfpsr_model_type_count(); }   // ref: spriteAPI.h:65
// ref: spriteAPI.cpp:262-277 ModelType(visibleModel, shadowModel): DenseModel_create(visible) + the shadow model's geometry (all parts)
inline int32_t spriteWorld_createModelType(const Model &visibleModel, const Model &shadowModel) {
	b200_require(visibleModel, "spriteWorld_createModelType");
	auto flatten = [](const Model &m, std::vector<dfpsr_polygon> &polygons) { for (const B200Part &part : m->parts) { polygons.insert(polygons.end(), part.polygons.begin(), part.polygons.end()); } };
	std::vector<dfpsr_polygon> visible, shadow;
	flatten(visibleModel, visible);
	std::vector<dfpsr_dense_triangle> triangles((size_t)dfpsr_dense_model_triangle_count(visible.data(), (int32_t)visible.size()));
	float mn[3], mx[3];
	b200_check(dfpsr_dense_model_build(visibleModel->points.data(), (int32_t)(visibleModel->points.size() / 3), visible.data(), (int32_t)visible.size(), triangles.data(), mn, mx));
	dfpsr_host_model host{};
	if (shadowModel) {
		flatten(shadowModel, shadow);
		host.points = shadowModel->points.data(); host.pointCount = (int32_t)(shadowModel->points.size() / 3);
		host.polygons = shadow.data(); host.polygonCount = (int32_t)shadow.size();
	}
	int32_t index = -1;
	b200_check(dfpsr_model_type_create(triangles.data(), (int32_t)triangles.size(), mn, mx, shadowModel ? &host : nullptr, &index));
	return index;
}

struct B200SpriteWorld {
	dfpsr_sprite_world *handle = nullptr;
	OrthoSystem ortho;
	B200SpriteWorld(const OrthoSystem &ortho, int32_t shadowResolution) : ortho(ortho) { b200_check(dfpsr_sprite_world_create(&handle, &ortho.pod, shadowResolution)); }
	~B200SpriteWorld() { dfpsr_sprite_world_destroy(handle); }
	B200SpriteWorld(const B200SpriteWorld &) = delete;
	B200SpriteWorld &operator=(const B200SpriteWorld &) = delete;
};
using SpriteWorld = std::shared_ptr<B200SpriteWorld>;
inline dfpsr_sprite_world *b200_world(const SpriteWorld &world, const char *method) { if (!world) { throwError(std::string("The world handle was null in ") + method); } return world->handle; }
inline SpriteWorld spriteWorld_create(OrthoSystem ortho, int32_t shadowResolution) { return std::make_shared<B200SpriteWorld>(ortho, shadowResolution); } // ref: spriteAPI.h:68
inline void spriteWorld_addBackgroundSprite(SpriteWorld &world, const SpriteInstance &sprite) { dfpsr_sprite_instance s = b200_pod(sprite); b200_check(dfpsr_sprite_world_add_background_sprite(b200_world(world, "spriteWorld_addBackgroundSprite"), &s)); }
inline void spriteWorld_addBackgroundModel(SpriteWorld &world, const ModelInstance &instance) { dfpsr_model_instance m = b200_pod(instance); b200_check(dfpsr_sprite_world_add_background_model(b200_world(world, "spriteWorld_addBackgroundModel"), &m)); }
inline void spriteWorld_addTemporarySprite(SpriteWorld &world, const SpriteInstance &sprite) { dfpsr_sprite_instance s = b200_pod(sprite); b200_check(dfpsr_sprite_world_add_temporary_sprite(b200_world(world, "spriteWorld_addTemporarySprite"), &s)); }
inline void spriteWorld_addTemporaryModel(SpriteWorld &world, const ModelInstance &instance) { dfpsr_model_instance m = b200_pod(instance); b200_check(dfpsr_sprite_world_add_temporary_model(b200_world(world, "spriteWorld_addTemporaryModel"), &m)); }
// ref: spriteAPI.h:76-92 — erase everything touching the box (the selection-callback overloads are available on the C ABI)
inline void spriteWorld_removeBackgroundSprites(SpriteWorld &world, const IVector3D &searchMinBound, const IVector3D &searchMaxBound) {
	const int32_t mn[3] = {searchMinBound.x, searchMinBound.y, searchMinBound.z}, mx[3] = {searchMaxBound.x, searchMaxBound.y, searchMaxBound.z};
	b200_check(dfpsr_sprite_world_remove_background_sprites(b200_world(world, "spriteWorld_removeBackgroundSprites"), mn, mx, nullptr, nullptr));
}
inline void spriteWorld_removeBackgroundModels(SpriteWorld &world, const IVector3D &searchMinBound, const IVector3D &searchMaxBound) {
	const int32_t mn[3] = {searchMinBound.x, searchMinBound.y, searchMinBound.z}, mx[3] = {searchMaxBound.x, searchMaxBound.y, searchMaxBound.z};
	b200_check(dfpsr_sprite_world_remove_background_models(b200_world(world, "spriteWorld_removeBackgroundModels"), mn, mx, nullptr, nullptr));
}
inline void spriteWorld_createTemporary_pointLight(SpriteWorld &world, const FVector3D position, float radius, float intensity, const ColorRgbaI32 &color, bool shadowCasting) { // ref: spriteAPI.h:96
	const float p[3] = {position.x, position.y, position.z}; const int32_t c[3] = {color.red, color.green, color.blue};
	b200_check(dfpsr_sprite_world_create_temporary_point_light(b200_world(world, "spriteWorld_createTemporary_pointLight"), p, radius, intensity, c, shadowCasting ? 1 : 0));
}
inline void spriteWorld_createTemporary_directedLight(SpriteWorld &world, const FVector3D direction, float intensity, const ColorRgbaI32 &color) { // ref: spriteAPI.h:97
	const float d[3] = {direction.x, direction.y, direction.z}; const int32_t c[3] = {color.red, color.green, color.blue};
	b200_check(dfpsr_sprite_world_create_temporary_directed_light(b200_world(world, "spriteWorld_createTemporary_directedLight"), d, intensity, c));
}
inline void spriteWorld_clearTemporary(SpriteWorld &world) { b200_check(dfpsr_sprite_world_clear_temporary(b200_world(world, "spriteWorld_clearTemporary"))); } // ref: spriteAPI.h:99
inline void spriteWorld_draw(SpriteWorld &world, ImageRgbaU8 &colorTarget) { // ref: spriteAPI.h:102
	dfpsr_image c = colorTarget.pod();
	b200_check(dfpsr_sprite_world_draw(b200_world(world, "spriteWorld_draw"), &c, b200_stream())); colorTarget.touchedByDevice();
}
inline IVector3D spriteWorld_findGroundAtPixel(SpriteWorld &world, const ImageRgbaU8 &colorBuffer, const IVector2D &pixelLocation) { // ref: spriteAPI.h:109
	int32_t r[3]; b200_check(dfpsr_sprite_world_find_ground_at_pixel(b200_world(world, "spriteWorld_findGroundAtPixel"), colorBuffer.width, colorBuffer.height, pixelLocation.x, pixelLocation.y, r));
	return IVector3D(r[0], r[1], r[2]);
}
inline void spriteWorld_moveCameraInPixels(SpriteWorld &world, const IVector2D &pixelOffset) { b200_check(dfpsr_sprite_world_move_camera_in_pixels(b200_world(world, "spriteWorld_moveCameraInPixels"), pixelOffset.x, pixelOffset.y)); } // ref: spriteAPI.h:113
inline IVector3D spriteWorld_getCameraLocation(const SpriteWorld &world) { int32_t r[3]; b200_check(dfpsr_sprite_world_get_camera_location(b200_world(world, "spriteWorld_getCameraLocation"), r)); return IVector3D(r[0], r[1], r[2]); } // ref: spriteAPI.h:126
inline void spriteWorld_setCameraLocation(SpriteWorld &world, const IVector3D miniTileLocation) { const int32_t p[3] = {miniTileLocation.x, miniTileLocation.y, miniTileLocation.z}; b200_check(dfpsr_sprite_world_set_camera_location(b200_world(world, "spriteWorld_setCameraLocation"), p)); }
inline int32_t spriteWorld_getCameraDirectionIndex(const SpriteWorld &world) { int32_t i = 0; b200_check(dfpsr_sprite_world_get_camera_direction_index(b200_world(world, "spriteWorld_getCameraDirectionIndex"), &i)); return i; } // ref: spriteAPI.h:132
inline void spriteWorld_setCameraDirectionIndex(SpriteWorld &world, int32_t index) { b200_check(dfpsr_sprite_world_set_camera_direction_index(b200_world(world, "spriteWorld_setCameraDirectionIndex"), index)); }
inline OrthoSystem &spriteWorld_getOrthoSystem(SpriteWorld &world) { b200_world(world, "spriteWorld_getOrthoSystem"); return world->ortho; } // ref: spriteAPI.h:139
// ref: spriteAPI.h:120-123 — views over the world's device buffers of the last frame (not owned: valid until the world is resized or destroyed)
template <typename P> inline B200Image<P> b200_borrowed_image(const dfpsr_image &im) {
	B200Image<P> r;
	if (im.data == nullptr) { return r; }
	r.buffer = std::make_shared<B200Buffer>(im.data, (size_t)im.stride * (size_t)im.height);
	r.width = im.width; r.height = im.height; r.stride = im.stride; r.packOrder = (PackOrderIndex)im.packOrder;
	return r;
}
inline ImageRgbaU8 spriteWorld_getDiffuseBuffer(SpriteWorld &world) { dfpsr_image im{}; b200_check(dfpsr_sprite_world_get_buffers(b200_world(world, "spriteWorld_getDiffuseBuffer"), &im, nullptr, nullptr, nullptr)); return b200_borrowed_image<uint32_t>(im); }
inline ImageRgbaU8 spriteWorld_getNormalBuffer(SpriteWorld &world) { dfpsr_image im{}; b200_check(dfpsr_sprite_world_get_buffers(b200_world(world, "spriteWorld_getNormalBuffer"), nullptr, &im, nullptr, nullptr)); return b200_borrowed_image<uint32_t>(im); }
inline ImageRgbaU8 spriteWorld_getLightBuffer(SpriteWorld &world) { dfpsr_image im{}; b200_check(dfpsr_sprite_world_get_buffers(b200_world(world, "spriteWorld_getLightBuffer"), nullptr, nullptr, &im, nullptr)); return b200_borrowed_image<uint32_t>(im); }
inline ImageF32 spriteWorld_getHeightBuffer(SpriteWorld &world) { dfpsr_image im{}; b200_check(dfpsr_sprite_world_get_buffers(b200_world(world, "spriteWorld_getHeightBuffer"), nullptr, nullptr, nullptr, &im)); return b200_borrowed_image<float>(im); }

// ---------------------------------------------------------------- filters (ref: api/filterAPI.h:41-79)
inline ImageRgbaU8 filter_resize(const ImageRgbaU8 &source, Sampler interpolation, int32_t newWidth, int32_t newHeight) {
	if (!image_exists(source)) { return ImageRgbaU8(); } // ref: api/filterAPI.cpp:852-860
	ImageRgbaU8 result = b200_image_create<uint32_t>(newWidth, newHeight, PackOrderIndex::RGBA);
	size_t need = dfpsr_filter_resize_scratch_bytes(source.width, source.height, newWidth, newHeight);
	std::unique_ptr<B200Buffer> scratch(need > 0 ? new B200Buffer(need) : nullptr);
	dfpsr_image t = result.pod(), s = source.pod();
	b200_check(dfpsr_filter_resize(&t, &s, (int32_t)interpolation, source.subImage ? 1 : 0, scratch ? scratch->device : nullptr, b200_stream()));
	if (scratch) { b200_check(dfpsr_stream_synchronize(b200_stream())); }
	result.touchedByDevice();
	return result;
}
inline ImageU8 filter_resize(const ImageU8 &source, Sampler interpolation, int32_t newWidth, int32_t newHeight) { // ref: api/filterAPI.h:42, api/filterAPI.cpp:862-870
	if (!image_exists(source)) { return ImageU8(); }
	ImageU8 result = b200_image_create<uint8_t>(newWidth, newHeight, PackOrderIndex::RGBA);
	const bool twoPasses = newWidth != source.width && newHeight > source.height;
	std::unique_ptr<B200Buffer> scratch(twoPasses ? new B200Buffer((size_t)newWidth * (size_t)source.height) : nullptr);
	dfpsr_image t = result.pod(), s = source.pod();
	b200_check(dfpsr_filter_resize_u8(&t, &s, (int32_t)interpolation, scratch ? scratch->device : nullptr, b200_stream()));
	if (scratch) { b200_check(dfpsr_stream_synchronize(b200_stream())); }
	result.touchedByDevice();
	return result;
}
// ref: api/filterAPI.h:54-79 filter_mapRgbaU8 / filter_generateRgbaU8. Three ways to say what a pixel is:
//   PixelProgram   the function as CUDA C++ text, compiled for the device on first use (dfpsr_filter_map_program): runs at HBM bandwidth.
//                  filter_mapRgbaU8(target, PixelProgram("int4 s = read_clamp(0, x, y); return make_int4(s.x * 2, s.y * 2, s.z * 2, s.w);", {source}));
//   a host callable  `ColorRgbaI32 f(int32_t x, int32_t y)` exactly like the reference's lambda — code written for the reference compiles
//                  unchanged. A device cannot call into the host, so the callable runs on the host over the image's mirror and the result is
//                  uploaded: correct, and as slow as the reference ("when speed is not critical", filterAPI.h:50). Images it captures are read
//                  through image_readPixel_* (one download each).
//   a device op    one of the pre-compiled DFPSR_MAP_* instances.
struct PixelProgram {
	std::string body;
	std::vector<ImageRgbaU8> sources;
	explicit PixelProgram(const std::string &body, const std::vector<ImageRgbaU8> &sources = std::vector<ImageRgbaU8>()) : body(body), sources(sources) {}
};
inline void filter_mapRgbaU8(ImageRgbaU8 &target, const PixelProgram &program, int32_t startX = 0, int32_t startY = 0) {
	if (!image_exists(target)) { return; }
	std::vector<dfpsr_image> sources;
	for (const ImageRgbaU8 &source : program.sources) { sources.push_back(source.pod()); }
	dfpsr_image t = target.pod();
	b200_check(dfpsr_filter_map_program(&t, program.body.c_str(), sources.data(), (int32_t)sources.size(), startX, startY, b200_stream())); target.touchedByDevice();
}
template <typename F, typename = decltype(std::declval<F &>()(0, 0).red)>
inline void filter_mapRgbaU8(ImageRgbaU8 &target, F &&lambda, int32_t startX = 0, int32_t startY = 0) {
	if (!image_exists(target)) { return; }
	std::vector<uint32_t> rows((size_t)target.width * (size_t)target.height);
	for (int32_t y = 0; y < target.height; y++) {
		for (int32_t x = 0; x < target.width; x++) { rows[(size_t)y * target.width + x] = b200_pack(lambda(x + startX, y + startY), target.packOrder); }
	}
	image_upload(target, rows.data(), target.width * 4);
}
inline ImageRgbaU8 filter_generateRgbaU8(int32_t width, int32_t height, const PixelProgram &program, int32_t startX = 0, int32_t startY = 0) {
	ImageRgbaU8 result = b200_image_create<uint32_t>(width, height, PackOrderIndex::RGBA);
	filter_mapRgbaU8(result, program, startX, startY);
	return result;
}
template <typename F, typename = decltype(std::declval<F &>()(0, 0).red)>
inline ImageRgbaU8 filter_generateRgbaU8(int32_t width, int32_t height, F &&lambda, int32_t startX = 0, int32_t startY = 0) {
	ImageRgbaU8 result = b200_image_create<uint32_t>(width, height, PackOrderIndex::RGBA);
	filter_mapRgbaU8(result, lambda, startX, startY);
	return result;
}
inline void filter_mapRgbaU8(ImageRgbaU8 &target, int32_t deviceOp, const int32_t *params, int32_t paramCount, const ImageRgbaU8 &source = ImageRgbaU8(), int32_t startX = 0, int32_t startY = 0) {
	dfpsr_image t = target.pod(), s = source.pod();
	b200_check(dfpsr_filter_map(&t, deviceOp, params, paramCount, image_exists(source) ? &s : nullptr, startX, startY, b200_stream())); target.touchedByDevice();
}
// ref: implementation/gui/DsrWindow.cpp:255-281 DsrWindow::showCanvas — what a window back-end calls with its own (host) canvas memory:
// block magnify by pixelScale into the canvas's pack order on the device, one copy into the canvas rows.
inline void b200_showCanvas(const ImageRgbaU8 &canvas, int32_t pixelScale, void *hostCanvas, int32_t hostStrideBytes, int32_t hostWidth, int32_t hostHeight, PackOrderIndex hostPackOrder) {
	if (!image_exists(canvas)) { return; }
	dfpsr_image c = canvas.pod();
	b200_check(dfpsr_canvas_show(&c, pixelScale, hostCanvas, hostStrideBytes, hostWidth, hostHeight, (int32_t)hostPackOrder, b200_stream()));
}
inline void filter_blockMagnify(ImageRgbaU8 &target, const ImageRgbaU8 &source, int32_t pixelWidth, int32_t pixelHeight) {
	dfpsr_image t = target.pod(), s = source.pod();
	b200_check(dfpsr_filter_block_magnify(&t, &s, pixelWidth, pixelHeight, b200_stream())); target.touchedByDevice();
}

} // namespace dsr
