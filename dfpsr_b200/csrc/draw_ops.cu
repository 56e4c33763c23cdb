// draw_ops.cu — the 2D draw calls that touch the same device images around the hot path, on sm_100a:
//   draw_rectangle (RgbaU8 / F32)                   ref: api/drawAPI.cpp:72-174
//   draw_line (RgbaU8 / F32)                        ref: api/drawAPI.cpp:176-310
//   draw_alphaFilter / draw_maxAlpha / draw_alphaClip / draw_silhouette   ref: api/drawAPI.cpp:636-744, :926-960
// They exist so that GUI overlays, debug wireframes (api/rendererAPI.cpp:374-399) and sprite tools can draw into device images
// without a device -> host -> device round trip. All integer arithmetic, bit-exact; one thread per pixel (per line step).
#include "common.cuh"

#include <algorithm>
#include <utility>

namespace dfpsr {
namespace {

struct Img { uint8_t *data; int32_t width, height, stride, packOrder; };
inline Img img_of(const dfpsr_image *im) {
	Img r;
	if (im == nullptr) { r.data = nullptr; r.width = r.height = r.stride = r.packOrder = 0; return r; }
	r.data = (uint8_t *)im->data; r.width = im->width; r.height = im->height; r.stride = im->stride; r.packOrder = im->packOrder;
	return r;
}
inline bool exists(const dfpsr_image *im) { return im != nullptr && im->data != nullptr; }
__device__ __forceinline__ uint32_t *px_u32(const Img &im, int32_t x, int32_t y) { return (uint32_t *)(im.data + (size_t)y * (size_t)im.stride) + x; }
__device__ __forceinline__ uint8_t *px_u8(const Img &im, int32_t x, int32_t y) { return im.data + (size_t)y * (size_t)im.stride + x; }

struct Intersection { int32_t tx, ty, sx, sy, w, h; };
// ref: api/drawAPI.cpp:330-385 ImageIntersection
bool intersect(const Img &target, const Img &source, int32_t left, int32_t top, Intersection &out) {
	const int32_t x0 = left > 0 ? left : 0, y0 = top > 0 ? top : 0;
	const int64_t r = (int64_t)left + source.width, b = (int64_t)top + source.height;
	const int32_t x1 = r < target.width ? (int32_t)r : target.width, y1 = b < target.height ? (int32_t)b : target.height;
	if (x1 <= x0 || y1 <= y0) { return false; }
	out.tx = x0; out.ty = y0; out.sx = x0 - left; out.sy = y0 - top; out.w = x1 - x0; out.h = y1 - y0;
	return true;
}

// ref: api/drawAPI.cpp:46-54
__device__ __forceinline__ uint32_t byte_mul(uint32_t a, uint32_t b) { return (a * b * 65793u + 8388608u) >> 24; }

__device__ __forceinline__ void unpack(uint32_t c, uint32_t shifts, uint32_t *v) {
	v[0] = (c >> (shifts & 31u)) & 255u; v[1] = (c >> ((shifts >> 8) & 31u)) & 255u; v[2] = (c >> ((shifts >> 16) & 31u)) & 255u; v[3] = (c >> ((shifts >> 24) & 31u)) & 255u;
}

__global__ void __launch_bounds__(256) rectangle_kernel(Img target, int32_t left, int32_t top, int32_t width, int32_t height, uint32_t value) {
	const int32_t x = (int32_t)(blockIdx.x * 32u + (threadIdx.x & 31u)), y = (int32_t)(blockIdx.y * 8u + (threadIdx.x >> 5));
	if (x < width && y < height) { *px_u32(target, left + x, top + y) = value; }
}

// ref: api/drawAPI.cpp:176-283 drawLineSuper in closed form (LineParams / line_pixel in common.cuh): one thread per step of the major axis
__global__ void __launch_bounds__(256) line_kernel(Img target, LineParams p, uint32_t value) {
	const int32_t i = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= p.steps) { return; }
	int32_t x, y;
	if (line_pixel(p, i, target.width, target.height, x, y)) { *px_u32(target, x, y) = value; }
}

enum { OP_ALPHA_FILTER = 0, OP_MAX_ALPHA = 1, OP_MAX_ALPHA_OFFSET = 2, OP_ALPHA_CLIP = 3 };
template <int OP>
__global__ void __launch_bounds__(256) image_over_kernel(Img target, Img source, Intersection is, int32_t parameter) {
	const int32_t x = (int32_t)(blockIdx.x * 32u + (threadIdx.x & 31u)), y = (int32_t)(blockIdx.y * 8u + (threadIdx.x >> 5));
	if (x >= is.w || y >= is.h) { return; }
	const uint32_t ts = pack_shifts(target.packOrder), ss = pack_shifts(source.packOrder);
	uint32_t *tp = px_u32(target, is.tx + x, is.ty + y);
	uint32_t s[4], t[4];
	unpack(*px_u32(source, is.sx + x, is.sy + y), ss, s);
	if (OP == OP_ALPHA_FILTER) { // ref: api/drawAPI.cpp:636-661
		const uint32_t sourceRatio = s[3];
		if (sourceRatio == 0u) { return; }
		if (sourceRatio == 255u) { *tp = pack_rgba_ordered(s[0], s[1], s[2], 255u, ts); return; }
		unpack(*tp, ts, t);
		const uint32_t targetRatio = 255u - sourceRatio;
		// the sums are stored into bytes by the reference: keep the low 8 bits
		*tp = pack_rgba_ordered((byte_mul(t[0], targetRatio) + byte_mul(s[0], sourceRatio)) & 255u, (byte_mul(t[1], targetRatio) + byte_mul(s[1], sourceRatio)) & 255u,
		                        (byte_mul(t[2], targetRatio) + byte_mul(s[2], sourceRatio)) & 255u, (byte_mul(t[3], targetRatio) + sourceRatio) & 255u, ts);
	} else if (OP == OP_MAX_ALPHA) { // ref: api/drawAPI.cpp:663-679
		unpack(*tp, ts, t);
		if ((int32_t)s[3] > (int32_t)t[3]) { *tp = pack_rgba_ordered(s[0], s[1], s[2], s[3], ts); }
	} else if (OP == OP_MAX_ALPHA_OFFSET) { // ref: api/drawAPI.cpp:680-697
		int32_t sourceAlpha = (int32_t)s[3];
		if (sourceAlpha > 0) {
			sourceAlpha += parameter;
			unpack(*tp, ts, t);
			if (sourceAlpha > (int32_t)t[3]) {
				if (sourceAlpha < 0) { sourceAlpha = 0; }
				if (sourceAlpha > 255) { sourceAlpha = 255; }
				*tp = pack_rgba_ordered(s[0], s[1], s[2], (uint32_t)sourceAlpha, ts);
			}
		}
	} else { // OP_ALPHA_CLIP, ref: api/drawAPI.cpp:700-715
		if ((int32_t)s[3] > parameter) { *tp = pack_rgba_ordered(s[0], s[1], s[2], 255u, ts); }
	}
}

// ref: api/drawAPI.cpp:717-757 drawSilhouette_template / imageImpl_drawSilhouette
__global__ void __launch_bounds__(256) silhouette_kernel(Img target, Img source, Intersection is, uint32_t red, uint32_t green, uint32_t blue, uint32_t alpha, int fullAlpha) {
	const int32_t x = (int32_t)(blockIdx.x * 32u + (threadIdx.x & 31u)), y = (int32_t)(blockIdx.y * 8u + (threadIdx.x >> 5));
	if (x >= is.w || y >= is.h) { return; }
	const uint32_t ts = pack_shifts(target.packOrder);
	uint32_t sourceRatio = *px_u8(source, is.sx + x, is.sy + y);
	if (!fullAlpha) { sourceRatio = byte_mul(sourceRatio, alpha); }
	if (sourceRatio == 0u) { return; }
	uint32_t *tp = px_u32(target, is.tx + x, is.ty + y);
	if (sourceRatio == 255u) { *tp = pack_rgba_ordered(red, green, blue, 255u, ts); return; }
	uint32_t t[4];
	unpack(*tp, ts, t);
	const uint32_t targetRatio = 255u - sourceRatio;
	*tp = pack_rgba_ordered((byte_mul(t[0], targetRatio) + byte_mul(red, sourceRatio)) & 255u, (byte_mul(t[1], targetRatio) + byte_mul(green, sourceRatio)) & 255u,
	                        (byte_mul(t[2], targetRatio) + byte_mul(blue, sourceRatio)) & 255u, (byte_mul(t[3], targetRatio) + sourceRatio) & 255u, ts);
}

// ---- 8-bit and 16-bit monochrome images (ref: api/drawAPI.cpp:130-150, :284-297, :519-634, :759-832)
__device__ __forceinline__ uint16_t *px_u16(const Img &im, int32_t x, int32_t y) { return (uint16_t *)(im.data + (size_t)y * (size_t)im.stride) + x; }
__device__ __forceinline__ float *px_f32(const Img &im, int32_t x, int32_t y) { return (float *)(im.data + (size_t)y * (size_t)im.stride) + x; }

__device__ __forceinline__ void store_mono(const Img &im, int32_t format, int32_t x, int32_t y, uint32_t value) {
	if (format == DFPSR_FORMAT_U8) { *px_u8(im, x, y) = (uint8_t)value; } else { *px_u16(im, x, y) = (uint16_t)value; }
}
__global__ void __launch_bounds__(256) rectangle_mono_kernel(Img target, int32_t format, int32_t left, int32_t top, int32_t width, int32_t height, uint32_t value) {
	const int32_t x = (int32_t)(blockIdx.x * 32u + (threadIdx.x & 31u)), y = (int32_t)(blockIdx.y * 8u + (threadIdx.x >> 5));
	if (x < width && y < height) { store_mono(target, format, left + x, top + y, value); }
}
__global__ void __launch_bounds__(256) line_mono_kernel(Img target, int32_t format, LineParams p, uint32_t value) {
	const int32_t i = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= p.steps) { return; }
	int32_t x, y;
	if (line_pixel(p, i, target.width, target.height, x, y)) { store_mono(target, format, x, y, value); }
}

// ref: api/drawAPI.cpp:476-486 saturateFloat
__device__ __forceinline__ uint32_t saturate_float(float value) {
	if (!(value >= 0.5f)) { return 0u; }
	if (value > 254.5f) { return 255u; }
	return (uint32_t)(uint8_t)__float2int_rz(value + 0.5f);
}

// Every overload of draw_copy between different pixel formats (ref: api/drawAPI.cpp:544-634), one thread per pixel.
__global__ void __launch_bounds__(256) copy_convert_kernel(Img target, int32_t targetFormat, Img source, int32_t sourceFormat, Intersection is) {
	const int32_t x = (int32_t)(blockIdx.x * 32u + (threadIdx.x & 31u)), y = (int32_t)(blockIdx.y * 8u + (threadIdx.x >> 5));
	if (x >= is.w || y >= is.h) { return; }
	const int32_t sx = is.sx + x, sy = is.sy + y, tx = is.tx + x, ty = is.ty + y;
	if (targetFormat == DFPSR_FORMAT_RGBA_U8) { // luma replicated, alpha 255
		uint32_t luma;
		if (sourceFormat == DFPSR_FORMAT_U8) { luma = *px_u8(source, sx, sy); }
		else if (sourceFormat == DFPSR_FORMAT_U16) { luma = min((uint32_t)*px_u16(source, sx, sy), 255u); }
		else { luma = saturate_float(*px_f32(source, sx, sy)); }
		*px_u32(target, tx, ty) = pack_rgba_ordered(luma, luma, luma, 255u, pack_shifts(target.packOrder));
	} else if (targetFormat == DFPSR_FORMAT_U8) {
		if (sourceFormat == DFPSR_FORMAT_F32) { *px_u8(target, tx, ty) = (uint8_t)saturate_float(*px_f32(source, sx, sy)); }
		else if (sourceFormat == DFPSR_FORMAT_U16) { *px_u8(target, tx, ty) = (uint8_t)min((uint32_t)*px_u16(source, sx, sy), 255u); }
		else { *px_u8(target, tx, ty) = *px_u8(source, sx, sy); }
	} else if (targetFormat == DFPSR_FORMAT_U16) {
		if (sourceFormat == DFPSR_FORMAT_U8) { *px_u16(target, tx, ty) = *px_u8(source, sx, sy); }
		// the reference stores the first BYTE of the float (api/drawAPI.cpp:603-614 writes *sourcePixel, not the clamped value): reproduced
		else if (sourceFormat == DFPSR_FORMAT_F32) { *px_u16(target, tx, ty) = (uint16_t)(__float_as_uint(*px_f32(source, sx, sy)) & 255u); }
		else { *px_u16(target, tx, ty) = *px_u16(source, sx, sy); }
	} else { // F32 target
		if (sourceFormat == DFPSR_FORMAT_U8) { *px_f32(target, tx, ty) = (float)*px_u8(source, sx, sy); }
		else if (sourceFormat == DFPSR_FORMAT_U16) { *px_f32(target, tx, ty) = (float)min((uint32_t)*px_u16(source, sx, sy), 255u); } // clamped to 255 by the reference (:624-633)
		else { *px_f32(target, tx, ty) = *px_f32(source, sx, sy); }
	}
}

// ref: api/drawAPI.cpp:759-832 draw_higher on 16-bit heights with 0, 1 or 2 RGBA payloads (height 0 = nothing there)
__global__ void __launch_bounds__(256) higher_u16_kernel(Img targetH, Img sourceH, Img targetA, Img sourceA, Img targetB, Img sourceB, Intersection is, int32_t offset) {
	const int32_t x = (int32_t)(blockIdx.x * 32u + (threadIdx.x & 31u)), y = (int32_t)(blockIdx.y * 8u + (threadIdx.x >> 5));
	if (x >= is.w || y >= is.h) { return; }
	int32_t newHeight = *px_u16(sourceH, is.sx + x, is.sy + y);
	if (newHeight <= 0) { return; }
	newHeight += offset;
	if (newHeight < 0) { newHeight = 0; }
	if (newHeight > 65535) { newHeight = 65535; }
	uint16_t *t = px_u16(targetH, is.tx + x, is.ty + y);
	// without payload images the reference also requires newHeight > 0 (:768); with payloads a zero can never exceed the target
	if (newHeight > (int32_t)*t) {
		*t = (uint16_t)newHeight;
		if (targetA.data) {
			uint32_t c[4];
			unpack(*px_u32(sourceA, is.sx + x, is.sy + y), pack_shifts(sourceA.packOrder), c);
			*px_u32(targetA, is.tx + x, is.ty + y) = pack_rgba_ordered(c[0], c[1], c[2], c[3], pack_shifts(targetA.packOrder));
		}
		if (targetB.data) {
			uint32_t c[4];
			unpack(*px_u32(sourceB, is.sx + x, is.sy + y), pack_shifts(sourceB.packOrder), c);
			*px_u32(targetB, is.tx + x, is.ty + y) = pack_rgba_ordered(c[0], c[1], c[2], c[3], pack_shifts(targetB.packOrder));
		}
	}
}

inline dim3 grid2d(int32_t w, int32_t h) { return dim3((unsigned)((w + 31) / 32), (unsigned)((h + 7) / 8)); }
inline uint32_t clamp255(int32_t v) { return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

int rectangle(const dfpsr_image *image, int32_t left, int32_t top, int32_t width, int32_t height, uint32_t value, cudaStream_t stream) {
	if (!exists(image)) { return 0; }
	const Img t = img_of(image);
	// ref: api/drawAPI.cpp:73-77 — the rectangle is clipped to the image
	const int64_t right = (int64_t)left + width, bottom = (int64_t)top + height;
	const int32_t l = left > 0 ? left : 0, tp = top > 0 ? top : 0;
	const int32_t r = right < t.width ? (int32_t)right : t.width, b = bottom < t.height ? (int32_t)bottom : t.height;
	if (r <= l || b <= tp) { return 0; }
	DFPSR_LAUNCH(rectangle_kernel, grid2d(r - l, b - tp), 256, 0, stream, t, l, tp, r - l, b - tp, value);
	return 0;
}

inline bool line_params(const Img &t, int32_t x1, int32_t y1, int32_t x2, int32_t y2, LineParams &p) { return dfpsr::line_params(t.width, t.height, x1, y1, x2, y2, p); }
int line(const dfpsr_image *image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, uint32_t value, cudaStream_t stream) {
	if (!exists(image)) { return 0; }
	const Img t = img_of(image);
	LineParams p;
	if (!line_params(t, x1, y1, x2, y2, p)) { return 0; }
	DFPSR_LAUNCH(line_kernel, (p.steps + 255) / 256, 256, 0, stream, t, p, value);
	return 0;
}
uint32_t saturate_and_pack(const dfpsr_image *image, const int32_t rgba[4]) { // ref: api/imageAPI.cpp image_saturateAndPack
	const uint32_t shifts = pack_shifts(image->packOrder);
	return (clamp255(rgba[0]) << (shifts & 31u)) | (clamp255(rgba[1]) << ((shifts >> 8) & 31u)) | (clamp255(rgba[2]) << ((shifts >> 16) & 31u)) | (clamp255(rgba[3]) << ((shifts >> 24) & 31u));
}

} // namespace
} // namespace dfpsr

using namespace dfpsr;

extern "C" {

int dfpsr_draw_rectangle_rgba(const dfpsr_image *image, int32_t left, int32_t top, int32_t width, int32_t height, const int32_t colorRgba[4], void *stream) {
	if (!exists(image)) { return 0; }
	DFPSR_REQUIRE(colorRgba != nullptr, "draw_rectangle: null colour");
	return rectangle(image, left, top, width, height, saturate_and_pack(image, colorRgba), as_stream(stream));
}
int dfpsr_draw_rectangle_f32(const dfpsr_image *image, int32_t left, int32_t top, int32_t width, int32_t height, float value, void *stream) {
	uint32_t bits;
	memcpy(&bits, &value, 4);
	return rectangle(image, left, top, width, height, bits, as_stream(stream));
}
int dfpsr_draw_line_rgba(const dfpsr_image *image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, const int32_t colorRgba[4], void *stream) {
	if (!exists(image)) { return 0; }
	DFPSR_REQUIRE(colorRgba != nullptr, "draw_line: null colour");
	return line(image, x1, y1, x2, y2, saturate_and_pack(image, colorRgba), as_stream(stream));
}
int dfpsr_draw_line_f32(const dfpsr_image *image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, float value, void *stream) {
	uint32_t bits;
	memcpy(&bits, &value, 4);
	return line(image, x1, y1, x2, y2, bits, as_stream(stream));
}

#define IMAGE_OVER(OP, parameter)                                                                                         \
	if (!exists(target) || !exists(source)) { return 0; }                                                                 \
	const Img t = img_of(target), s = img_of(source);                                                                     \
	Intersection is;                                                                                                      \
	if (!intersect(t, s, left, top, is)) { return 0; }                                                                    \
	DFPSR_LAUNCH(image_over_kernel<OP>, grid2d(is.w, is.h), 256, 0, as_stream(stream), t, s, is, (parameter));            \
	return 0;

int dfpsr_draw_alpha_filter(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, void *stream) { IMAGE_OVER(OP_ALPHA_FILTER, 0) }
int dfpsr_draw_max_alpha(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, int32_t sourceAlphaOffset, void *stream) {
	if (sourceAlphaOffset == 0) { IMAGE_OVER(OP_MAX_ALPHA, 0) }
	IMAGE_OVER(OP_MAX_ALPHA_OFFSET, sourceAlphaOffset)
}
int dfpsr_draw_alpha_clip(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, int32_t threshold, void *stream) { IMAGE_OVER(OP_ALPHA_CLIP, threshold) }

int dfpsr_draw_silhouette(const dfpsr_image *target, const dfpsr_image *silhouetteU8, const int32_t colorRgba[4], int32_t left, int32_t top, void *stream) {
	if (!exists(target) || !exists(silhouetteU8)) { return 0; }
	DFPSR_REQUIRE(colorRgba != nullptr, "draw_silhouette: null colour");
	if (colorRgba[3] <= 0) { return 0; } // ref: api/drawAPI.cpp:748
	const Img t = img_of(target), s = img_of(silhouetteU8);
	Intersection is;
	if (!intersect(t, s, left, top, is)) { return 0; }
	DFPSR_LAUNCH(silhouette_kernel, grid2d(is.w, is.h), 256, 0, as_stream(stream), t, s, is, clamp255(colorRgba[0]), clamp255(colorRgba[1]), clamp255(colorRgba[2]), clamp255(colorRgba[3]), colorRgba[3] >= 255 ? 1 : 0);
	return 0;
}

static bool mono_format(int32_t format) { return format == DFPSR_FORMAT_U8 || format == DFPSR_FORMAT_U16; }
static bool known_format(int32_t format) { return format >= DFPSR_FORMAT_U8 && format <= DFPSR_FORMAT_RGBA_U8; }

int dfpsr_draw_rectangle_mono(const dfpsr_image *image, int32_t format, int32_t left, int32_t top, int32_t width, int32_t height, int32_t color, void *stream) {
	if (!exists(image)) { return 0; }
	DFPSR_REQUIRE(mono_format(format), "draw_rectangle_mono: the format must be DFPSR_FORMAT_U8 or DFPSR_FORMAT_U16");
	const int32_t top_value = format == DFPSR_FORMAT_U8 ? 255 : 65535;
	uint32_t value = (uint32_t)(color < 0 ? 0 : (color > top_value ? top_value : color)); // ref: api/drawAPI.cpp:130-134
	// ref: api/drawAPI.cpp:136-146 — a 16-bit colour whose two bytes are equal takes the reference's memset path, which is handed 0 instead
	// of the byte, so such rectangles come out black (65535 included). Reproduced: drop-in means the same pixels.
	if (format == DFPSR_FORMAT_U16 && (value & 0xFFu) == (value >> 8)) { value = 0u; }
	const Img t = img_of(image);
	const int64_t right = (int64_t)left + width, bottom = (int64_t)top + height;
	const int32_t l = left > 0 ? left : 0, tp = top > 0 ? top : 0;
	const int32_t r = right < t.width ? (int32_t)right : t.width, b = bottom < t.height ? (int32_t)bottom : t.height;
	if (r <= l || b <= tp) { return 0; }
	DFPSR_LAUNCH(rectangle_mono_kernel, grid2d(r - l, b - tp), 256, 0, as_stream(stream), t, format, l, tp, r - l, b - tp, value);
	return 0;
}

int dfpsr_draw_line_mono(const dfpsr_image *image, int32_t format, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int32_t color, void *stream) {
	if (!exists(image)) { return 0; }
	DFPSR_REQUIRE(mono_format(format), "draw_line_mono: the format must be DFPSR_FORMAT_U8 or DFPSR_FORMAT_U16");
	const int32_t top_value = format == DFPSR_FORMAT_U8 ? 255 : 65535;
	const uint32_t value = (uint32_t)(color < 0 ? 0 : (color > top_value ? top_value : color)); // ref: api/drawAPI.cpp:284-297
	const Img t = img_of(image);
	LineParams p;
	if (!line_params(t, x1, y1, x2, y2, p)) { return 0; }
	DFPSR_LAUNCH(line_mono_kernel, (p.steps + 255) / 256, 256, 0, as_stream(stream), t, format, p, value);
	return 0;
}

int dfpsr_draw_copy_formats(const dfpsr_image *target, int32_t targetFormat, const dfpsr_image *source, int32_t sourceFormat, int32_t left, int32_t top, void *stream) {
	if (!exists(target) || !exists(source)) { return 0; }
	DFPSR_REQUIRE(known_format(targetFormat) && known_format(sourceFormat), "draw_copy_formats: unknown pixel format");
	// ref: api/drawAPI.h:91-103 — the thirteen overloads; an RGBA source only goes to an RGBA target
	DFPSR_REQUIRE(sourceFormat != DFPSR_FORMAT_RGBA_U8 || targetFormat == DFPSR_FORMAT_RGBA_U8, "draw_copy_formats: the reference has no draw_copy from RGBA to a monochrome image");
	if (targetFormat == DFPSR_FORMAT_RGBA_U8 && sourceFormat == DFPSR_FORMAT_RGBA_U8) { return dfpsr_draw_copy_rgba(target, source, left, top, stream); }
	if (targetFormat == DFPSR_FORMAT_F32 && sourceFormat == DFPSR_FORMAT_F32) { return dfpsr_draw_copy_f32(target, source, left, top, stream); }
	const Img t = img_of(target), s = img_of(source);
	Intersection is;
	if (!intersect(t, s, left, top, is)) { return 0; }
	DFPSR_LAUNCH(copy_convert_kernel, grid2d(is.w, is.h), 256, 0, as_stream(stream), t, targetFormat, s, sourceFormat, is);
	return 0;
}

int dfpsr_draw_higher_u16(const dfpsr_image *targetHeight, const dfpsr_image *sourceHeight, const dfpsr_image *targetA, const dfpsr_image *sourceA, const dfpsr_image *targetB, const dfpsr_image *sourceB, int32_t left, int32_t top, int32_t sourceHeightOffset, void *stream) {
	// ref: api/drawAPI.cpp:962-979 — every image given to an overload must exist, otherwise nothing is drawn
	if (!exists(targetHeight) || !exists(sourceHeight)) { return 0; }
	const bool wantA = targetA != nullptr || sourceA != nullptr, wantB = targetB != nullptr || sourceB != nullptr;
	if (wantA && (!exists(targetA) || !exists(sourceA))) { return 0; }
	if (wantB && (!exists(targetB) || !exists(sourceB))) { return 0; }
	DFPSR_REQUIRE(!wantB || wantA, "draw_higher_u16: a second payload image needs a first one");
	const Img th = img_of(targetHeight), sh = img_of(sourceHeight);
	if (wantA) { DFPSR_REQUIRE(sourceA->width == sh.width && sourceA->height == sh.height, "draw_higher_u16: sourceA and sourceHeight differ in size"); }
	if (wantB) { DFPSR_REQUIRE(sourceB->width == sh.width && sourceB->height == sh.height, "draw_higher_u16: sourceB and sourceHeight differ in size"); }
	Intersection is;
	if (!intersect(th, sh, left, top, is)) { return 0; }
	DFPSR_LAUNCH(higher_u16_kernel, grid2d(is.w, is.h), 256, 0, as_stream(stream), th, sh, img_of(wantA ? targetA : nullptr), img_of(wantA ? sourceA : nullptr), img_of(wantB ? targetB : nullptr), img_of(wantB ? sourceB : nullptr), is, sourceHeightOffset);
	return 0;
}

} // extern "C"
