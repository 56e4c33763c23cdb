// ref_wrap.cpp — TEST INFRASTRUCTURE, not product code.
//
// A flat C wrapper around the UNMODIFIED reference (Dawoodoz/DFPSR), compiled by oracle/Makefile from the
// sources where they lie under /root/reference/Source into oracle/_ref/libdfpsr_ref_{sse,scalar}.so.
// No reference source is copied: this file only #includes the reference's public headers and calls
// its public API (model_render, renderer_*, draw_*, filter_*, lightAPI). It is used
//   * by tests/ to pin the C restatement (oracle/dfpsr_oracle.c) and the CUDA path against the real thing,
//   * by bench.py --impl reference / cpu_baseline as the timed CPU implementation.
// The product (dfpsr_b200/) never links or loads it.
#include "../include/dfpsr_b200.h"

#include "DFPSR/includeFramework.h"
#include "DFPSR/implementation/render/model/Model.h"
#include "DFPSR/implementation/render/Camera.h"
#include "DFPSR/api/rendererAPI.h"
#include "SDK/SpriteEngine/lightAPI.h"
#include "SDK/SpriteEngine/orthoAPI.h"
// The sprite engine keeps renderDenseModel, SpriteType, ModelType and SpriteWorldImpl local to its translation unit. Including that
// unit here (it is compiled from where it lies and is NOT built a second time, see oracle/Makefile) lets the wrappers below call them.
#include "SDK/SpriteEngine/spriteAPI.cpp"

#include <vector>
#include <cstring>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <new>

using namespace dsr;

// The SDK's terrain example, compiled where it lies: its scene generators (bump / light / diffuse / colour maps from the three media images,
// the grid model) are called by ref_sdk_terrain_scene below exactly in the order of its dsrMain (SDK/terrain/main.cpp:343-373). The window
// loop of the example is never entered; DSR_MAIN_CALLER is defined away so that the example does not bring its own main().
namespace sdk_terrain {
#undef DSR_MAIN_CALLER
#define DSR_MAIN_CALLER(X)
#include "SDK/terrain/main.cpp"
}

// The reference's heap (DFPSR/base/heap.cpp:671-684) writes a new allocation's header through an AllocationHeader
// pointer, which slices off HeapHeader's own fields (destructor, useCount, flags, binIndex): it relies on every
// 16 MiB arena from operator new being fresh zero pages from mmap. Inside a long-lived Python process glibc raises
// its dynamic mmap threshold after large numpy arrays are freed, arenas then come from recycled (dirty) memory and
// heap_free calls a garbage destructor pointer. The replaceable global allocation functions below (a standard C++
// customisation point, resolved inside this shared object first) hand the reference zeroed memory, which restores
// the condition it assumes without touching its sources.
void *operator new(std::size_t size) {
	void *p = calloc(size ? size : 1, 1);
	if (p == nullptr) throw std::bad_alloc();
	return p;
}
void *operator new[](std::size_t size) { return operator new(size); }
void operator delete(void *p) noexcept { free(p); }
void operator delete[](void *p) noexcept { free(p); }
void operator delete(void *p, std::size_t) noexcept { free(p); }
void operator delete[](void *p, std::size_t) noexcept { free(p); }

namespace {

struct AnyImage {
	int kind = 0; // 1 = RgbaU8, 2 = F32
	ImageRgbaU8 rgba;
	AlignedImageRgbaU8 rgbaAligned; // only for whole images
	OrderedImageRgbaU8 rgbaOrdered; // only for whole RGBA-order images
	ImageF32 f32;
	AlignedImageF32 f32Aligned; // only for whole images
	ImageU8 u8;                 // kind 3
	ImageU16 u16;               // kind 4
};

std::vector<AnyImage> g_images;
std::vector<TextureRgbaU8> g_textures;
std::vector<Model> g_models;
bool g_started = false;

void ensureStarted() {
	if (!g_started) {
		heap_startingApplication();
		g_started = true;
	}
}

FVector3D v3(const float *p) { return FVector3D(p[0], p[1], p[2]); }

Transform3D toTransform(const dfpsr_transform3d *t) {
	return Transform3D(v3(t->position), FMatrix3x3(v3(t->xAxis), v3(t->yAxis), v3(t->zAxis)));
}

FMatrix3x3 toMatrix(const dfpsr_matrix3x3 &m) {
	return FMatrix3x3(v3(m.xAxis), v3(m.yAxis), v3(m.zAxis));
}

Camera toCamera(const dfpsr_camera *c) {
	Transform3D location = toTransform(&c->location);
	if (c->perspective) {
		return Camera::createPerspective(location, c->imageWidth, c->imageHeight, c->widthSlope, c->nearClip, c->farClip);
	} else {
		return Camera::createOrthogonal(location, c->imageWidth, c->imageHeight, c->widthSlope);
	}
}

OrthoView toView(const dfpsr_ortho_view *v) {
	OrthoView result;
	result.normalToWorldSpace = toMatrix(v->normalToWorldSpace);
	result.screenDepthToLightSpace = toMatrix(v->screenDepthToLightSpace);
	result.lightSpaceToScreenDepth = toMatrix(v->lightSpaceToScreenDepth);
	return result;
}

ImageRgbaU8 rgbaOrNull(int id) { return id >= 0 ? g_images[id].rgba : ImageRgbaU8(); }
ImageF32 f32OrNull(int id) { return id >= 0 ? g_images[id].f32 : ImageF32(); }
TextureRgbaU8 texOrNull(int id) { return id >= 0 ? g_textures[id] : TextureRgbaU8(); }

void fromMatrix(dfpsr_matrix3x3 &out, const FMatrix3x3 &m) {
	out.xAxis[0] = m.xAxis.x; out.xAxis[1] = m.xAxis.y; out.xAxis[2] = m.xAxis.z;
	out.yAxis[0] = m.yAxis.x; out.yAxis[1] = m.yAxis.y; out.yAxis[2] = m.yAxis.z;
	out.zAxis[0] = m.zAxis.x; out.zAxis[1] = m.zAxis.y; out.zAxis[2] = m.zAxis.z;
}

} // namespace

extern "C" {

// 0 = SSE2 (rcpps + Newton-Raphson reciprocal), 1 = scalar (exact 1/x). See Makefile.
int ref_flavour() {
	#ifdef USE_SSE2
		return 0;
	#else
		return 1;
	#endif
}

int ref_thread_count() { return getThreadCount(); }

double ref_time_seconds() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void ref_free_all() {
	g_models.clear();
	g_textures.clear();
	g_images.clear();
}

// Called at interpreter exit so that the reference's heap does not complain about a missing DSR_MAIN_CALLER.
void ref_shutdown() {
	if (g_started) {
		heap_hardExitCleaning();
		g_started = false;
	}
}

// ---- images

int ref_image_create_rgba(int w, int h, int packOrder) {
	ensureStarted();
	AnyImage img;
	img.kind = 1;
	if (packOrder == DFPSR_PACK_RGBA) {
		img.rgbaOrdered = image_create_RgbaU8(w, h);
		img.rgbaAligned = img.rgbaOrdered;
	} else {
		img.rgbaAligned = image_create_RgbaU8_native(w, h, (PackOrderIndex)packOrder);
	}
	img.rgba = img.rgbaAligned;
	g_images.push_back(img);
	return (int)g_images.size() - 1;
}

int ref_image_create_f32(int w, int h) {
	ensureStarted();
	AnyImage img;
	img.kind = 2;
	img.f32Aligned = image_create_F32(w, h);
	img.f32 = img.f32Aligned;
	g_images.push_back(img);
	return (int)g_images.size() - 1;
}

int ref_image_sub(int id, int x, int y, int w, int h) {
	AnyImage img;
	img.kind = g_images[id].kind;
	if (img.kind == 1) {
		img.rgba = image_getSubImage(g_images[id].rgba, IRect(x, y, w, h));
	} else {
		img.f32 = image_getSubImage(g_images[id].f32, IRect(x, y, w, h));
	}
	g_images.push_back(img);
	return (int)g_images.size() - 1;
}

int ref_image_width(int id) { return g_images[id].kind == 1 ? image_getWidth(g_images[id].rgba) : image_getWidth(g_images[id].f32); }
int ref_image_height(int id) { return g_images[id].kind == 1 ? image_getHeight(g_images[id].rgba) : image_getHeight(g_images[id].f32); }
int ref_image_stride(int id) { return g_images[id].kind == 1 ? image_getStride(g_images[id].rgba) : image_getStride(g_images[id].f32); }

// Tight or strided host rows in/out (4 bytes per pixel for both kinds).
void ref_image_write(int id, const void *src, int srcStride) {
	const AnyImage &img = g_images[id];
	int w = ref_image_width(id), h = ref_image_height(id);
	for (int y = 0; y < h; y++) {
		void *row = img.kind == 1 ? (void*)image_getSafePointer<uint32_t>(img.rgba, y).getUnsafe() : (void*)image_getSafePointer<float>(img.f32, y).getUnsafe();
		memcpy(row, (const uint8_t*)src + (size_t)y * srcStride, (size_t)w * 4);
	}
}

void ref_image_read(int id, void *dst, int dstStride) {
	const AnyImage &img = g_images[id];
	int w = ref_image_width(id), h = ref_image_height(id);
	for (int y = 0; y < h; y++) {
		const void *row = img.kind == 1 ? (const void*)image_getSafePointer<uint32_t>(img.rgba, y).getUnsafe() : (const void*)image_getSafePointer<float>(img.f32, y).getUnsafe();
		memcpy((uint8_t*)dst + (size_t)y * dstStride, row, (size_t)w * 4);
	}
}

void ref_image_fill_rgba(int id, int r, int g, int b, int a) { image_fill(g_images[id].rgba, ColorRgbaI32(r, g, b, a)); }
void ref_image_fill_f32(int id, float v) { image_fill(g_images[id].f32, v); }

// ---- textures

// level0 is tight w*h RGBA-packed u32 with power-of-two w and h; lower levels are generated by the reference.
int ref_texture_create(int w, int h, int levels, const uint32_t *level0) {
	ensureStarted();
	TextureRgbaU8 tex = texture_create_RgbaU8(w, h, levels);
	int tw = texture_getMaxWidth(tex), th = texture_getMaxHeight(tex);
	for (int y = 0; y < th && y < h; y++) {
		SafePointer<uint32_t> target = texture_getSafePointer(tex, 0u, y);
		memcpy(target.getUnsafe(), level0 + (size_t)y * w, (size_t)(w < tw ? w : tw) * 4);
	}
	texture_generatePyramid(tex);
	g_textures.push_back(tex);
	return (int)g_textures.size() - 1;
}

int ref_texture_from_image(int imageId, int levels) {
	ensureStarted();
	g_textures.push_back(texture_create_RgbaU8(g_images[imageId].rgba, levels));
	return (int)g_textures.size() - 1;
}

// out: log2w, log2h, maxMip, startOffset, maxLevelMask, totalPixels
void ref_texture_info(int id, uint32_t *out) {
	const TextureRgbaU8 &t = g_textures[id];
	out[0] = t.impl_log2width; out[1] = t.impl_log2height; out[2] = t.impl_maxMipLevel;
	out[3] = t.impl_startOffset; out[4] = t.impl_maxLevelMask;
	out[5] = t.impl_startOffset + (1u << (t.impl_log2width + t.impl_log2height));
}

void ref_texture_read(int id, uint32_t *out) {
	uint32_t info[6];
	ref_texture_info(id, info);
	SafePointer<uint32_t> data = g_textures[id].impl_buffer.getSafe<uint32_t>("ref_texture_read");
	memcpy(out, data.getUnsafe(), (size_t)info[5] * 4);
}

// ---- models

int ref_model_create(const float *points, int pointCount, const dfpsr_polygon *polygons, int polygonCount, int filter, int diffuseTex, int lightTex) {
	ensureStarted();
	Model model = model_create();
	model_setFilter(model, filter == DFPSR_FILTER_ALPHA ? Filter::Alpha : Filter::Solid);
	for (int p = 0; p < pointCount; p++) {
		model_addPoint(model, v3(points + 3 * p));
	}
	int part = model_addEmptyPart(model, U"part");
	if (diffuseTex >= 0) { model_setDiffuseMap(model, part, g_textures[diffuseTex]); }
	if (lightTex >= 0) { model_setLightMap(model, part, g_textures[lightTex]); }
	List<Polygon> &target = model->partBuffer[part].polygonBuffer;
	for (int i = 0; i < polygonCount; i++) {
		const dfpsr_polygon &src = polygons[i];
		Polygon polygon(src.pointIndices[0], src.pointIndices[1], src.pointIndices[2], src.pointIndices[3]);
		polygon.pointIndices[3] = src.pointIndices[3];
		for (int c = 0; c < 4; c++) {
			polygon.texCoords[c] = FVector4D(src.texCoords[c][0], src.texCoords[c][1], src.texCoords[c][2], src.texCoords[c][3]);
			polygon.colors[c] = FVector4D(src.colors[c][0], src.colors[c][1], src.colors[c][2], src.colors[c][3]);
		}
		target.push(polygon);
	}
	g_models.push_back(model);
	return (int)g_models.size() - 1;
}

void ref_model_bounds(int model, float *minOut, float *maxOut) {
	FVector3D mn, mx;
	model_getBoundingBox(g_models[model], mn, mx);
	minOut[0] = mn.x; minOut[1] = mn.y; minOut[2] = mn.z;
	maxOut[0] = mx.x; maxOut[1] = mx.y; maxOut[2] = mx.z;
}

// mode 0: model_render (single thread, immediate). mode 1: renderer_begin / renderer_giveTask / renderer_end.
void ref_model_render(int model, const dfpsr_transform3d *modelToWorld, int colorId, int depthId, const dfpsr_camera *camera, int mode) {
	ImageRgbaU8 color = rgbaOrNull(colorId);
	ImageF32 depth = f32OrNull(depthId);
	Camera cam = toCamera(camera);
	Transform3D m2w = toTransform(modelToWorld);
	if (mode == 0) {
		model_render(g_models[model], m2w, color, depth, cam);
	} else {
		static Renderer worker = renderer_create();
		renderer_begin(worker, color, depth);
		renderer_giveTask(worker, g_models[model], m2w, cam);
		renderer_end(worker);
	}
}

// Several models in one renderer_begin / renderer_end frame, in order.
void ref_models_render_frame(const int *models, const dfpsr_transform3d *modelToWorld, int count, int colorId, int depthId, const dfpsr_camera *camera) {
	ImageRgbaU8 color = rgbaOrNull(colorId);
	ImageF32 depth = f32OrNull(depthId);
	Camera cam = toCamera(camera);
	static Renderer worker = renderer_create();
	renderer_begin(worker, color, depth);
	for (int i = 0; i < count; i++) {
		renderer_giveTask(worker, g_models[models[i]], toTransform(modelToWorld + i), cam);
	}
	renderer_end(worker);
}

// ---- the renderer API step by step, for the occlusion grid (api/rendererAPI.h:56-135)
static Renderer &stepRenderer() { static Renderer worker = renderer_create(); return worker; }
void ref_renderer_begin(int colorId, int depthId) {
	ImageRgbaU8 color = rgbaOrNull(colorId);
	ImageF32 depth = f32OrNull(depthId);
	renderer_begin(stepRenderer(), color, depth);
}
void ref_renderer_give_task(int model, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera) {
	renderer_giveTask(stepRenderer(), g_models[model], toTransform(modelToWorld), toCamera(camera));
}
void ref_renderer_occlude_from_box(const float *mn, const float *mx, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera) {
	renderer_occludeFromBox(stepRenderer(), v3(mn), v3(mx), toTransform(modelToWorld), toCamera(camera));
}
void ref_renderer_occlude_from_top_rows(const dfpsr_camera *camera) { renderer_occludeFromTopRows(stepRenderer(), toCamera(camera)); }
void ref_renderer_occlude_from_existing_triangles() { renderer_occludeFromExistingTriangles(stepRenderer()); }
int ref_renderer_is_box_visible(const float *mn, const float *mx, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera) {
	return renderer_isBoxVisible(stepRenderer(), v3(mn), v3(mx), toTransform(modelToWorld), toCamera(camera)) ? 1 : 0;
}
int ref_renderer_has_occluders() { return renderer_hasOccluders(stepRenderer()) ? 1 : 0; }
void ref_renderer_end() { renderer_end(stepRenderer()); }
void ref_renderer_end_wireframe() { renderer_end(stepRenderer(), true); }

void ref_model_render_depth(int model, const dfpsr_transform3d *modelToWorld, int depthId, const dfpsr_camera *camera) {
	ImageF32 depth = f32OrNull(depthId);
	model_renderDepth(g_models[model], toTransform(modelToWorld), depth, toCamera(camera));
}

// The whole SDK terrain frame: clear colour+depth, begin, giveTask, end (SDK/terrain/main.cpp:397-421). Returns seconds.
double ref_terrain_frame(int model, const dfpsr_transform3d *modelToWorld, int colorId, int depthId, const dfpsr_camera *camera) {
	double t0 = ref_time_seconds();
	image_fill(g_images[colorId].rgba, ColorRgbaI32(0, 0, 0, 0));
	image_fill(g_images[depthId].f32, 0.0f);
	ref_model_render(model, modelToWorld, colorId, depthId, camera, 1);
	return ref_time_seconds() - t0;
}

void ref_project_points(const float *points, int count, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, dfpsr_projected_point *out) {
	Camera cam = toCamera(camera);
	Transform3D m2w = toTransform(modelToWorld);
	for (int i = 0; i < count; i++) {
		ProjectedPoint p = cam.worldToScreen(m2w.transformPoint(v3(points + 3 * i)));
		out[i].cs[0] = p.cs.x; out[i].cs[1] = p.cs.y; out[i].cs[2] = p.cs.z;
		out[i].is[0] = p.is.x; out[i].is[1] = p.is.y;
		out[i].pad_ = 0;
		out[i].flat[0] = p.flat.x; out[i].flat[1] = p.flat.y;
	}
}

// Fills the derived fields of a camera POD from the reference's own Camera, for checking dfpsr_camera_create_*.
void ref_camera_fill(dfpsr_camera *c) {
	Camera cam = toCamera(c);
	c->widthSlope = cam.widthSlope; c->heightSlope = cam.heightSlope;
	c->invWidthSlope = cam.invWidthSlope; c->invHeightSlope = cam.invHeightSlope;
	c->nearClip = cam.nearClip; c->farClip = cam.farClip;
	c->cullPlaneCount = cam.getFrustumPlaneCount(false);
	c->clipPlaneCount = cam.getFrustumPlaneCount(true);
	for (int clip = 0; clip < 2; clip++) {
		int n = cam.getFrustumPlaneCount(clip != 0);
		for (int s = 0; s < n; s++) {
			FPlane3D p = cam.getFrustumPlane(s, clip != 0);
			float *dst = clip ? c->clipPlanes[s] : c->cullPlanes[s];
			dst[0] = p.normal.x; dst[1] = p.normal.y; dst[2] = p.normal.z; dst[3] = p.offset;
		}
	}
}

int ref_camera_is_box_seen(const dfpsr_camera *c, const float *mn, const float *mx, const dfpsr_transform3d *modelToWorld) {
	return toCamera(c).isBoxSeen(v3(mn), v3(mx), toTransform(modelToWorld));
}

// ---- draw

void ref_draw_higher(int targetH, int sourceH, int targetA, int sourceA, int targetB, int sourceB, int left, int top, float offset) {
	if (targetA < 0) {
		// The height-only overload is declared (drawAPI.h:142) with a signature its definition (drawAPI.cpp:962)
		// does not match, so it cannot be linked; give it throw-away payload images instead.
		ImageRgbaU8 dummyTarget = image_create_RgbaU8(image_getWidth(g_images[targetH].f32), image_getHeight(g_images[targetH].f32));
		ImageRgbaU8 dummySource = image_create_RgbaU8(image_getWidth(g_images[sourceH].f32), image_getHeight(g_images[sourceH].f32));
		draw_higher(g_images[targetH].f32, g_images[sourceH].f32, dummyTarget, dummySource, left, top, offset);
	} else if (targetB < 0) {
		draw_higher(g_images[targetH].f32, g_images[sourceH].f32, g_images[targetA].rgba, g_images[sourceA].rgba, left, top, offset);
	} else {
		draw_higher(g_images[targetH].f32, g_images[sourceH].f32, g_images[targetA].rgba, g_images[sourceA].rgba, g_images[targetB].rgba, g_images[sourceB].rgba, left, top, offset);
	}
}

void ref_draw_copy(int target, int source, int left, int top) {
	if (g_images[target].kind == 1) {
		draw_copy(g_images[target].rgba, g_images[source].rgba, left, top);
	} else {
		draw_copy(g_images[target].f32, g_images[source].f32, left, top);
	}
}

// ---- Sandbox light

// Views of the reference's own orthogonal system (SDK/sandbox/media/Ortho.ini: tilt -0.6, 150 px per tile).
void ref_ortho_view(float cameraTilt, int pixelsPerTile, int viewIndex, dfpsr_ortho_view *out, int32_t *pixelOffsets /* xAxis.xy, zAxis.xy, yPixelsPerTile */) {
	ensureStarted();
	OrthoSystem system(cameraTilt, pixelsPerTile);
	const OrthoView &v = system.view[viewIndex];
	fromMatrix(out->normalToWorldSpace, v.normalToWorldSpace);
	fromMatrix(out->screenDepthToLightSpace, v.screenDepthToLightSpace);
	fromMatrix(out->lightSpaceToScreenDepth, v.lightSpaceToScreenDepth);
	if (pixelOffsets) {
		pixelOffsets[0] = v.pixelOffsetPerTileX.x; pixelOffsets[1] = v.pixelOffsetPerTileX.y;
		pixelOffsets[2] = v.pixelOffsetPerTileZ.x; pixelOffsets[3] = v.pixelOffsetPerTileZ.y;
		pixelOffsets[4] = v.yPixelsPerTile;
	}
}

void ref_light_directed(const dfpsr_ortho_view *view, int light, int normal, const float *direction, float intensity, const int32_t *color, int add) {
	OrthoView v = toView(view);
	OrderedImageRgbaU8 lightImage = g_images[light].rgbaOrdered;
	OrderedImageRgbaU8 normalImage = g_images[normal].rgbaOrdered;
	if (add) {
		addDirectedLight(v, lightImage, normalImage, v3(direction), intensity, ColorRgbI32(color[0], color[1], color[2]));
	} else {
		setDirectedLight(v, lightImage, normalImage, v3(direction), intensity, ColorRgbI32(color[0], color[1], color[2]));
	}
}

void ref_light_point(const dfpsr_ortho_view *view, const int32_t *worldCenter, int light, int normal, int height, const float *position, float radius, float intensity, const int32_t *color, int cubeMap) {
	OrthoView v = toView(view);
	OrderedImageRgbaU8 lightImage = g_images[light].rgbaOrdered;
	OrderedImageRgbaU8 normalImage = g_images[normal].rgbaOrdered;
	AlignedImageF32 heightImage = g_images[height].f32Aligned;
	if (cubeMap >= 0) {
		AlignedImageF32 cube = g_images[cubeMap].f32Aligned;
		addPointLight(v, IVector2D(worldCenter[0], worldCenter[1]), lightImage, normalImage, heightImage, v3(position), radius, intensity, ColorRgbI32(color[0], color[1], color[2]), cube);
	} else {
		addPointLight(v, IVector2D(worldCenter[0], worldCenter[1]), lightImage, normalImage, heightImage, v3(position), radius, intensity, ColorRgbI32(color[0], color[1], color[2]));
	}
}

void ref_light_blend(int color, int diffuse, int light) {
	AlignedImageRgbaU8 colorImage = g_images[color].rgbaAligned;
	OrderedImageRgbaU8 diffuseImage = g_images[diffuse].rgbaOrdered;
	OrderedImageRgbaU8 lightImage = g_images[light].rgbaOrdered;
	blendLight(colorImage, diffuseImage, lightImage);
}

// ---- filters

int ref_filter_resize(int source, int sampler, int newWidth, int newHeight) {
	AnyImage img;
	img.kind = 1;
	img.rgbaOrdered = filter_resize(g_images[source].rgba, sampler == DFPSR_SAMPLER_LINEAR ? Sampler::Linear : Sampler::Nearest, newWidth, newHeight);
	img.rgbaAligned = img.rgbaOrdered;
	img.rgba = img.rgbaOrdered;
	g_images.push_back(img);
	return (int)g_images.size() - 1;
}

int ref_filter_resize_u8(int source, int sampler, int newWidth, int newHeight) {
	AnyImage img;
	img.kind = 3;
	img.u8 = filter_resize(g_images[source].u8, sampler == DFPSR_SAMPLER_LINEAR ? Sampler::Linear : Sampler::Nearest, newWidth, newHeight);
	g_images.push_back(img);
	return (int)g_images.size() - 1;
}

void ref_filter_map(int target, int op, const int32_t *params, int source, int startX, int startY) {
	ImageRgbaU8 targetImage = g_images[target].rgba;
	if (op == DFPSR_MAP_XOR_PATTERN) {
		filter_mapRgbaU8(targetImage, [](int32_t x, int32_t y) -> ColorRgbaI32 {
			return ColorRgbaI32(x & 255, y & 255, (x ^ y) & 255, 255);
		}, startX, startY);
	} else if (op == DFPSR_MAP_AFFINE) {
		ImageRgbaU8 sourceImage = g_images[source].rgba;
		filter_mapRgbaU8(targetImage, [sourceImage, params](int32_t x, int32_t y) -> ColorRgbaI32 {
			ColorRgbaI32 s = image_readPixel_clamp(sourceImage, x, y);
			return ColorRgbaI32(s.red * params[0] + params[4], s.green * params[1] + params[5], s.blue * params[2] + params[6], s.alpha * params[3] + params[7]);
		}, startX, startY);
	} else if (op == DFPSR_MAP_CONSTANT) {
		filter_mapRgbaU8(targetImage, [params](int32_t x, int32_t y) -> ColorRgbaI32 {
			return ColorRgbaI32(params[0], params[1], params[2], params[3]);
		}, startX, startY);
	}
}

void ref_filter_block_magnify(int target, int source, int pixelWidth, int pixelHeight) {
	filter_blockMagnify(g_images[target].rgba, g_images[source].rgba, pixelWidth, pixelHeight);
}

// ---- the remaining draw calls (api/drawAPI.h:68-161)

int ref_image_create_u8(int w, int h, const uint8_t *pixels) { // tight rows in
	ensureStarted();
	AnyImage img;
	img.kind = 3;
	img.u8 = image_create_U8(w, h);
	for (int y = 0; y < h; y++) { memcpy(image_getSafePointer<uint8_t>(img.u8, y).getUnsafe(), pixels + (size_t)y * w, (size_t)w); }
	g_images.push_back(img);
	return (int)g_images.size() - 1;
}
int ref_image_create_u16(int w, int h, const uint16_t *pixels) { // tight rows in
	ensureStarted();
	AnyImage img;
	img.kind = 4;
	img.u16 = image_create_U16(w, h);
	for (int y = 0; y < h; y++) { memcpy(image_getSafePointer<uint16_t>(img.u16, y).getUnsafe(), pixels + (size_t)y * w, (size_t)w * 2); }
	g_images.push_back(img);
	return (int)g_images.size() - 1;
}
void ref_image_read_mono(int id, void *dst) { // tight rows out, 1 or 2 bytes per pixel
	const AnyImage &img = g_images[id];
	if (img.kind == 3) { int w = image_getWidth(img.u8); for (int y = 0; y < image_getHeight(img.u8); y++) { memcpy((uint8_t*)dst + (size_t)y * w, image_getSafePointer<uint8_t>(img.u8, y).getUnsafe(), (size_t)w); } }
	else { int w = image_getWidth(img.u16); for (int y = 0; y < image_getHeight(img.u16); y++) { memcpy((uint16_t*)dst + (size_t)y * w, image_getSafePointer<uint16_t>(img.u16, y).getUnsafe(), (size_t)w * 2); } }
}
void ref_draw_rectangle_mono(int image, int left, int top, int width, int height, int color) {
	if (g_images[image].kind == 3) { draw_rectangle(g_images[image].u8, IRect(left, top, width, height), color); } else { draw_rectangle(g_images[image].u16, IRect(left, top, width, height), color); }
}
void ref_draw_line_mono(int image, int x1, int y1, int x2, int y2, int color) {
	if (g_images[image].kind == 3) { draw_line(g_images[image].u8, x1, y1, x2, y2, color); } else { draw_line(g_images[image].u16, x1, y1, x2, y2, color); }
}
// every draw_copy overload of api/drawAPI.h:91-103, selected by the kinds of the two images (1 RGBA, 2 F32, 3 U8, 4 U16)
void ref_draw_copy_formats(int target, int source, int left, int top) {
	const AnyImage &t = g_images[target], &s = g_images[source];
	switch (t.kind * 10 + s.kind) {
		case 11: draw_copy(t.rgba, s.rgba, left, top); break; case 13: draw_copy(t.rgba, s.u8, left, top); break;
		case 14: draw_copy(t.rgba, s.u16, left, top); break;  case 12: draw_copy(t.rgba, s.f32, left, top); break;
		case 33: draw_copy(t.u8, s.u8, left, top); break;     case 32: draw_copy(t.u8, s.f32, left, top); break;
		case 34: draw_copy(t.u8, s.u16, left, top); break;    case 44: draw_copy(t.u16, s.u16, left, top); break;
		case 43: draw_copy(t.u16, s.u8, left, top); break;    case 42: draw_copy(t.u16, s.f32, left, top); break;
		case 22: draw_copy(t.f32, s.f32, left, top); break;   case 23: draw_copy(t.f32, s.u8, left, top); break;
		case 24: draw_copy(t.f32, s.u16, left, top); break;
	}
}
void ref_draw_higher_u16(int targetH, int sourceH, int targetA, int sourceA, int targetB, int sourceB, int left, int top, int offset) {
	if (targetA < 0) {
		// like the F32 case above: the height-only overload is declared (drawAPI.h:124) with a signature its definition (drawAPI.cpp:946) does
		// not match; throw-away payload images give the same heights (a clamped height of 0 can never exceed an unsigned target)
		ImageRgbaU8 dummyTarget = image_create_RgbaU8(image_getWidth(g_images[targetH].u16), image_getHeight(g_images[targetH].u16));
		ImageRgbaU8 dummySource = image_create_RgbaU8(image_getWidth(g_images[sourceH].u16), image_getHeight(g_images[sourceH].u16));
		draw_higher(g_images[targetH].u16, g_images[sourceH].u16, dummyTarget, dummySource, left, top, offset);
	}
	else if (targetB < 0) { draw_higher(g_images[targetH].u16, g_images[sourceH].u16, g_images[targetA].rgba, g_images[sourceA].rgba, left, top, offset); }
	else { draw_higher(g_images[targetH].u16, g_images[sourceH].u16, g_images[targetA].rgba, g_images[sourceA].rgba, g_images[targetB].rgba, g_images[sourceB].rgba, left, top, offset); }
}
void ref_draw_rectangle_rgba(int image, int left, int top, int width, int height, const int32_t *c) { draw_rectangle(g_images[image].rgba, IRect(left, top, width, height), ColorRgbaI32(c[0], c[1], c[2], c[3])); }
void ref_draw_rectangle_f32(int image, int left, int top, int width, int height, float value) { draw_rectangle(g_images[image].f32, IRect(left, top, width, height), value); }
void ref_draw_line_rgba(int image, int x1, int y1, int x2, int y2, const int32_t *c) { draw_line(g_images[image].rgba, x1, y1, x2, y2, ColorRgbaI32(c[0], c[1], c[2], c[3])); }
void ref_draw_line_f32(int image, int x1, int y1, int x2, int y2, float value) { draw_line(g_images[image].f32, x1, y1, x2, y2, value); }
void ref_draw_alpha_filter(int target, int source, int left, int top) { draw_alphaFilter(g_images[target].rgba, g_images[source].rgba, left, top); }
void ref_draw_max_alpha(int target, int source, int left, int top, int offset) { draw_maxAlpha(g_images[target].rgba, g_images[source].rgba, left, top, offset); }
void ref_draw_alpha_clip(int target, int source, int left, int top, int threshold) { draw_alphaClip(g_images[target].rgba, g_images[source].rgba, left, top, threshold); }
void ref_draw_silhouette(int target, int source, const int32_t *c, int left, int top) { draw_silhouette(g_images[target].rgba, g_images[source].u8, ColorRgbaI32(c[0], c[1], c[2], c[3]), left, top); }

// ---- Sandbox sprite engine (SDK/SpriteEngine/spriteAPI.cpp, orthoAPI.cpp)

static void fillOrthoCamera(dfpsr_ortho_camera &out, const OrthoView &v) {
	memset(&out, 0, sizeof(out));
	out.id = v.id; out.worldDirection = v.worldDirection;
	fromMatrix(out.normalToWorldSpace, v.normalToWorldSpace);
	out.pixelOffsetPerTileX[0] = v.pixelOffsetPerTileX.x; out.pixelOffsetPerTileX[1] = v.pixelOffsetPerTileX.y;
	out.pixelOffsetPerTileZ[0] = v.pixelOffsetPerTileZ.x; out.pixelOffsetPerTileZ[1] = v.pixelOffsetPerTileZ.y;
	out.yPixelsPerTile = v.yPixelsPerTile;
	fromMatrix(out.screenDepthToWorldSpace, v.screenDepthToWorldSpace);
	fromMatrix(out.worldSpaceToScreenDepth, v.worldSpaceToScreenDepth);
	fromMatrix(out.screenDepthToLightSpace, v.screenDepthToLightSpace);
	fromMatrix(out.lightSpaceToScreenDepth, v.lightSpaceToScreenDepth);
	out.roundedScreenPixelsToWorldTiles[0] = v.roundedScreenPixelsToWorldTiles.xAxis.x; out.roundedScreenPixelsToWorldTiles[1] = v.roundedScreenPixelsToWorldTiles.xAxis.y;
	out.roundedScreenPixelsToWorldTiles[2] = v.roundedScreenPixelsToWorldTiles.yAxis.x; out.roundedScreenPixelsToWorldTiles[3] = v.roundedScreenPixelsToWorldTiles.yAxis.y;
}

void ref_ortho_system(float cameraTilt, int pixelsPerTile, dfpsr_ortho_system *out) {
	ensureStarted();
	OrthoSystem system(cameraTilt, pixelsPerTile);
	out->cameraTilt = system.cameraTilt; out->pixelsPerTile = system.pixelsPerTile;
	for (int a = 0; a < 8; a++) { fillOrthoCamera(out->view[a], system.view[a]); }
}

static std::vector<DenseModel> g_dense;
static std::vector<SpriteWorld> g_worlds;

int ref_dense_model_create(int model) {
	g_dense.push_back(DenseModel_create(g_models[model]));
	return (int)g_dense.size() - 1;
}

int ref_dense_model_triangles(int dense, dfpsr_dense_triangle *out, float *minOut, float *maxOut) {
	const DenseModel &m = g_dense[dense];
	static_assert(sizeof(DenseTriangle) == sizeof(dfpsr_dense_triangle), "DenseTriangle layout");
	if (out) { for (int i = 0; i < m->triangles.length(); i++) { memcpy(out + i, &m->triangles[i], sizeof(DenseTriangle)); } }
	if (minOut) { minOut[0] = m->minBound.x; minOut[1] = m->minBound.y; minOut[2] = m->minBound.z; }
	if (maxOut) { maxOut[0] = m->maxBound.x; maxOut[1] = m->maxBound.y; maxOut[2] = m->maxBound.z; }
	return (int)m->triangles.length();
}

void ref_dense_model_render(int dense, float cameraTilt, int pixelsPerTile, int viewIndex, int height, int diffuse, int normal, float originX, float originY, const dfpsr_transform3d *modelToWorld, int highQuality, int32_t *rect) {
	OrthoSystem system(cameraTilt, pixelsPerTile);
	IRect r;
	if (highQuality) { r = renderDenseModel<true>(g_dense[dense], system.view[viewIndex], g_images[height].f32, g_images[diffuse].rgba, g_images[normal].rgba, FVector2D(originX, originY), toTransform(modelToWorld)); }
	else { r = renderDenseModel<false>(g_dense[dense], system.view[viewIndex], g_images[height].f32, g_images[diffuse].rgba, g_images[normal].rgba, FVector2D(originX, originY), toTransform(modelToWorld)); }
	if (rect) { rect[0] = r.left(); rect[1] = r.top(); rect[2] = r.width(); rect[3] = r.height(); }
}

// The reference only loads sprite types from <name>.png + <name>.ini: the atlas image and the configuration text are written to
// `folder` with the reference's own PNG encoder (lossless) and loaded back through spriteWorld_loadSpriteTypeFromFile.
int ref_sprite_type_create(int atlas, const char *iniText, const char *folder, const char *name) {
	ensureStarted();
	String folderPath = string_combine(folder), spriteName = string_combine(name);
	String base = file_combinePaths(folderPath, spriteName);
	image_save(g_images[atlas].rgba, string_combine(base, U".png"));
	string_save(string_combine(base, U".ini"), string_combine(iniText));
	return spriteWorld_loadSpriteTypeFromFile(folderPath, spriteName);
}

int ref_sprite_type_count() { return spriteWorld_getSpriteTypeCount(); }
// Real media: the SDK's own <name>.png + <name>.ini where they lie (SDK/sandbox/media/images), through the reference's own loader.
int ref_sprite_type_load(const char *folder, const char *name) {
	ensureStarted();
	return spriteWorld_loadSpriteTypeFromFile(string_combine(folder), string_combine(name));
}
int ref_image_load(const char *path) { // decoded by the reference's image_load_RgbaU8 (stb_image), RGBA order
	ensureStarted();
	AnyImage img;
	img.kind = 1;
	img.rgba = image_load_RgbaU8(string_combine(path));
	if (!image_exists(img.rgba)) { return -1; }
	g_images.push_back(img);
	return (int)g_images.size() - 1;
}
// The reference's own decimal parser (sprite configuration numbers go through it: spriteAPI.cpp:56-101)
double ref_string_to_double(const char *text) { return string_toDouble(string_combine(text)); }

int ref_model_type_create(int dense, int shadowModel) {
	ensureStarted();
	return (int)modelTypes.pushConstructGetIndex(g_dense[dense], shadowModel >= 0 ? g_models[shadowModel] : Model());
}

// sprite_generateFromModel (SDK/SpriteEngine/spriteAPI.cpp:1329-1432): returns the atlas as a new image id (-1 when nothing was generated) and the
// numbers of the configuration text: out = {centerX, centerY, frameRows, propertyColumns}.
int ref_sprite_generate_from_model(int visibleModel, int shadowModel, float cameraTilt, int pixelsPerTile, int cameraAngles, int32_t *out) {
	ensureStarted();
	ImageRgbaU8 atlas;
	String configText;
	sprite_generateFromModel(atlas, configText, g_models[visibleModel], shadowModel >= 0 ? g_models[shadowModel] : Model(), OrthoSystem(cameraTilt, pixelsPerTile), U"", cameraAngles);
	if (!image_exists(atlas)) { return -1; }
	SpriteConfig config(configText);
	out[0] = config.centerX; out[1] = config.centerY; out[2] = config.frameRows; out[3] = config.propertyColumns;
	AnyImage img;
	img.kind = 1;
	img.rgba = atlas;
	g_images.push_back(img);
	return (int)g_images.size() - 1;
}

int ref_world_create(float cameraTilt, int pixelsPerTile, int shadowResolution) {
	ensureStarted();
	g_worlds.push_back(spriteWorld_create(OrthoSystem(cameraTilt, pixelsPerTile), shadowResolution));
	return (int)g_worlds.size() - 1;
}

static SpriteInstance toSprite(const dfpsr_sprite_instance *s) {
	return SpriteInstance(s->typeIndex, s->direction, IVector3D(s->location[0], s->location[1], s->location[2]), s->shadowCasting != 0, s->userData);
}
static ModelInstance toModelInstance(const dfpsr_model_instance *m) { return ModelInstance(m->typeIndex, toTransform(&m->location), m->userData); }

void ref_world_add_background_sprite(int world, const dfpsr_sprite_instance *s) { spriteWorld_addBackgroundSprite(g_worlds[world], toSprite(s)); }
void ref_world_add_background_model(int world, const dfpsr_model_instance *m) { spriteWorld_addBackgroundModel(g_worlds[world], toModelInstance(m)); }
void ref_world_add_temporary_sprite(int world, const dfpsr_sprite_instance *s) { spriteWorld_addTemporarySprite(g_worlds[world], toSprite(s)); }
void ref_world_add_temporary_model(int world, const dfpsr_model_instance *m) { spriteWorld_addTemporaryModel(g_worlds[world], toModelInstance(m)); }
void ref_world_remove_background_sprites(int world, const int32_t *mn, const int32_t *mx) { spriteWorld_removeBackgroundSprites(g_worlds[world], IVector3D(mn[0], mn[1], mn[2]), IVector3D(mx[0], mx[1], mx[2])); }
void ref_world_remove_background_models(int world, const int32_t *mn, const int32_t *mx) { spriteWorld_removeBackgroundModels(g_worlds[world], IVector3D(mn[0], mn[1], mn[2]), IVector3D(mx[0], mx[1], mx[2])); }
void ref_world_point_light(int world, const float *position, float radius, float intensity, const int32_t *color, int shadowCasting) {
	spriteWorld_createTemporary_pointLight(g_worlds[world], v3(position), radius, intensity, ColorRgbI32(color[0], color[1], color[2]), shadowCasting != 0);
}
void ref_world_directed_light(int world, const float *direction, float intensity, const int32_t *color) {
	spriteWorld_createTemporary_directedLight(g_worlds[world], v3(direction), intensity, ColorRgbI32(color[0], color[1], color[2]));
}
void ref_world_clear_temporary(int world) { spriteWorld_clearTemporary(g_worlds[world]); }
void ref_world_set_camera_location(int world, const int32_t *p) { spriteWorld_setCameraLocation(g_worlds[world], IVector3D(p[0], p[1], p[2])); }
void ref_world_get_camera_location(int world, int32_t *p) { IVector3D l = spriteWorld_getCameraLocation(g_worlds[world]); p[0] = l.x; p[1] = l.y; p[2] = l.z; }
void ref_world_move_camera_in_pixels(int world, int x, int y) { spriteWorld_moveCameraInPixels(g_worlds[world], IVector2D(x, y)); }
void ref_world_set_camera_direction_index(int world, int index) { spriteWorld_setCameraDirectionIndex(g_worlds[world], index); }
void ref_world_find_ground_at_pixel(int world, int color, int x, int y, int32_t *out) {
	IVector3D l = spriteWorld_findGroundAtPixel(g_worlds[world], g_images[color].rgbaAligned, IVector2D(x, y));
	out[0] = l.x; out[1] = l.y; out[2] = l.z;
}
// colour must be a whole image (not a sub-image)
void ref_world_draw(int world, int color) { spriteWorld_draw(g_worlds[world], g_images[color].rgbaAligned); }
// Copies of the four deferred buffers after a draw: tight rows, 4 bytes per pixel.
void ref_world_read_buffers(int world, uint32_t *diffuse, uint32_t *normal, uint32_t *light, float *height) {
	SpriteWorld &w = g_worlds[world];
	int width = image_getWidth(w->diffuseBuffer), h = image_getHeight(w->diffuseBuffer);
	for (int y = 0; y < h; y++) {
		if (diffuse) { memcpy(diffuse + (size_t)y * width, image_getSafePointer<uint32_t>(w->diffuseBuffer, y).getUnsafe(), (size_t)width * 4); }
		if (normal) { memcpy(normal + (size_t)y * width, image_getSafePointer<uint32_t>(w->normalBuffer, y).getUnsafe(), (size_t)width * 4); }
		if (light) { memcpy(light + (size_t)y * width, image_getSafePointer<uint32_t>(w->lightBuffer, y).getUnsafe(), (size_t)width * 4); }
		if (height) { memcpy(height + (size_t)y * width, image_getSafePointer<float>(w->heightBuffer, y).getUnsafe(), (size_t)width * 4); }
	}
}

} // extern "C"

// ---- model importers (SDK/SpriteEngine/importer.cpp, DFPSR/implementation/render/model/format/dmf1.cpp)

// A resource pool that loads nothing and remembers which texture names the importer asked for, in order.
struct NameOnlyPool : public ResourcePool {
	std::vector<std::string> requested;
	static std::string narrow(const ReadableString &text) { std::string r; for (intptr_t i = 0; i < string_length(text); i++) { r.push_back((char)text[i]); } return r; }
	const ImageRgbaU8 fetchImageRgba(const ReadableString &name) override { requested.push_back(narrow(name)); return ImageRgbaU8(); }
	const TextureRgbaU8 fetchTextureRgba(const ReadableString &name, int32_t resolutions) override { (void)resolutions; requested.push_back(narrow(name)); return TextureRgbaU8(); }
};
static std::vector<std::vector<std::string>> g_importNames; // per imported model: texture names in request order

extern "C" {

// importer_loadModel(filename, flipX, axisConversion) on a PLY file; returns a model id (read it back with ref_model_dump*)
int ref_import_ply(const char *filename, int flipX, const dfpsr_transform3d *axisConversion) {
	ensureStarted();
	Model model = importer_loadModel(String(filename), flipX != 0, toTransform(axisConversion));
	g_models.push_back(model);
	g_importNames.resize(g_models.size());
	return (int)g_models.size() - 1;
}

// The scene of SDK/terrain built by the SDK's own code from the media folder's HeightMap.png, Cloud.png and RampIsland.png
// (SDK/terrain/main.cpp:343-373). Returns the model id; *textureId receives the id of its 5-level colour texture.
int ref_sdk_terrain_scene(const char *mediaPath, int *textureId) {
	ensureStarted();
	using namespace sdk_terrain;
	const String folder = String(mediaPath);
	ImageU8 heightMap = image_get_red(image_load_RgbaU8(file_combinePaths(folder, U"HeightMap.png")));
	ImageU8 genericCloudPattern = image_get_red(image_load_RgbaU8(file_combinePaths(folder, U"Cloud.png")));
	ImageRgbaU8 heightRamp = image_load_RgbaU8(file_combinePaths(folder, U"RampIsland.png"));
	const int32_t colorMapWidth = image_getWidth(heightMap) * tileColorDensity, colorMapHeight = image_getHeight(heightMap) * tileColorDensity;
	ImageF32 bumpMap = image_create_F32(colorMapWidth, colorMapHeight);
	generateBumpMap(bumpMap, heightMap, genericCloudPattern);
	ImageF32 lightMap = image_create_F32(colorMapWidth, colorMapHeight);
	FVector3D sunDirection = normalize(FVector3D(0.3f, -1.0f, 1.0f));
	float ambient = 0.2f;
	generateLightMap(lightMap, bumpMap, sunDirection, ambient);
	ImageRgbaU8 diffuseMap = image_create_RgbaU8(colorMapWidth, colorMapHeight);
	generateDiffuseMap(diffuseMap, bumpMap, heightRamp);
	TextureRgbaU8 colorTexture = texture_create_RgbaU8(colorMapWidth, colorMapHeight, 5);
	ImageRgbaU8 colorMap = texture_getMipLevelImage(colorTexture, 0);
	updateColorMap(colorMap, diffuseMap, lightMap);
	texture_generatePyramid(colorTexture);
	Model ground = createGrid(heightMap, colorTexture);
	g_textures.push_back(colorTexture);
	if (textureId) { *textureId = (int)g_textures.size() - 1; }
	g_models.push_back(ground);
	g_importNames.resize(g_models.size());
	return (int)g_models.size() - 1;
}

int ref_import_dmf1(const char *content, int detailLevel) {
	ensureStarted();
	NameOnlyPool pool;
	Model model = importFromContent_DMF1(String(content), pool, detailLevel);
	g_models.push_back(model);
	g_importNames.resize(g_models.size());
	g_importNames.back() = pool.requested;
	return (int)g_models.size() - 1;
}

// counts[0] = points, counts[1] = parts, counts[2] = polygons over all parts, counts[3] = filter, counts[4] = requested texture names
void ref_model_dump_counts(int model, int *counts) {
	const Model &m = g_models[model];
	counts[0] = model_getNumberOfPoints(m); counts[1] = model_getNumberOfParts(m); counts[2] = 0;
	for (int p = 0; p < counts[1]; p++) { counts[2] += model_getNumberOfPolygons(m, p); }
	counts[3] = model_getFilter(m) == Filter::Alpha ? DFPSR_FILTER_ALPHA : DFPSR_FILTER_SOLID;
	counts[4] = (size_t)model < g_importNames.size() ? (int)g_importNames[model].size() : 0;
}

// points: 3 floats each; polygons in part order (dfpsr_polygon = the reference's Polygon layout); polygonsPerPart: one count per part
void ref_model_dump(int model, float *points, dfpsr_polygon *polygons, int *polygonsPerPart) {
	const Model &m = g_models[model];
	for (int i = 0; i < model_getNumberOfPoints(m); i++) { FVector3D p = model_getPoint(m, i); points[3 * i] = p.x; points[3 * i + 1] = p.y; points[3 * i + 2] = p.z; }
	int at = 0;
	for (int p = 0; p < model_getNumberOfParts(m); p++) {
		const List<Polygon> &source = m->partBuffer[p].polygonBuffer;
		polygonsPerPart[p] = (int)source.length();
		for (int i = 0; i < (int)source.length(); i++, at++) {
			dfpsr_polygon &dst = polygons[at];
			for (int c = 0; c < 4; c++) {
				dst.pointIndices[c] = source[i].pointIndices[c];
				dst.texCoords[c][0] = source[i].texCoords[c].x; dst.texCoords[c][1] = source[i].texCoords[c].y; dst.texCoords[c][2] = source[i].texCoords[c].z; dst.texCoords[c][3] = source[i].texCoords[c].w;
				dst.colors[c][0] = source[i].colors[c].x; dst.colors[c][1] = source[i].colors[c].y; dst.colors[c][2] = source[i].colors[c].z; dst.colors[c][3] = source[i].colors[c].w;
			}
		}
	}
}

// name of part `part` (index >= 0) or requested texture name number -1 - part; truncated to size - 1 characters
void ref_model_dump_name(int model, int part, char *out, int size) {
	std::string text;
	if (part >= 0) { text = NameOnlyPool::narrow(model_getPartName(g_models[model], part)); }
	else { text = g_importNames[model][(size_t)(-1 - part)]; }
	snprintf(out, (size_t)size, "%s", text.c_str());
}

} // extern "C"
