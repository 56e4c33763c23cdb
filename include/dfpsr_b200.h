/*
 * dfpsr_b200.h — C ABI of the B200-native rendering hot path.
 *
 * Every entry point replaces one host-side loop of Dawoodoz/DFPSR (the reference), cited as
 * "ref: <file>:<line>" relative to /root/reference/Source. The reference is a C++14 library, so its
 * "FFI" for this path is the C++ API itself (rendererAPI.h, modelAPI.h, drawAPI.h, filterAPI.h,
 * SDK/SpriteEngine/lightAPI.h). The `dsr::` shim under dfpsr_b200/host/ re-exposes those C++ names
 * on top of this ABI; INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ or torch types.
 *   - every function returns 0 on success, non-zero on error; dfpsr_last_error() gives the message
 *     (thread local). There is NO CPU fallback: without a CUDA device every compute call fails.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream). Calls are asynchronous
 *     on that stream unless the name ends in _host (those take HOST buffers, copy in, run the same
 *     kernels, copy out and synchronise the stream before returning).
 *   - images are descriptors over device memory: {data, width, height, stride in bytes, pack order}.
 *     `data == NULL` means "image does not exist" (ref: api/imageAPI.h:96 image_exists).
 */
#ifndef DFPSR_B200_H
#define DFPSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFPSR_B200_ABI_VERSION 1

/* ---------------------------------------------------------------- POD mirrors */

/* ref: implementation/image/PackOrder.h:37-42 (PackOrderIndex) */
enum { DFPSR_PACK_RGBA = 0, DFPSR_PACK_BGRA = 1, DFPSR_PACK_ARGB = 2, DFPSR_PACK_ABGR = 3 };
/* ref: implementation/render/constants.h:34 (Filter) */
enum { DFPSR_FILTER_SOLID = 0, DFPSR_FILTER_ALPHA = 1 };
/* ref: api/filterAPI.h:33-36 (Sampler) */
enum { DFPSR_SAMPLER_NEAREST = 0, DFPSR_SAMPLER_LINEAR = 1 };

/* ref: math/Transform3D.h:33-36 — position then the three matrix axes, 12 floats. */
typedef struct dfpsr_transform3d {
	float position[3];
	float xAxis[3], yAxis[3], zAxis[3];
} dfpsr_transform3d;

/* ref: math/FMatrix3x3.h:33 */
typedef struct dfpsr_matrix3x3 {
	float xAxis[3], yAxis[3], zAxis[3];
} dfpsr_matrix3x3;

/* ref: implementation/render/Camera.h:128-150. Planes are {nx, ny, nz, offset} with normalised normals
 * (ref: math/FPlane3D.h:36-45); filled by dfpsr_camera_create_*(). */
typedef struct dfpsr_camera {
	int32_t perspective;
	dfpsr_transform3d location;
	float widthSlope, heightSlope, invWidthSlope, invHeightSlope;
	float imageWidth, imageHeight, nearClip, farClip;
	int32_t cullPlaneCount, clipPlaneCount;
	float cullPlanes[6][4];
	float clipPlanes[6][4];
} dfpsr_camera;

/* ref: implementation/render/model/Model.h:54-58 (Polygon, 144 bytes, same field order). */
typedef struct dfpsr_polygon {
	int32_t pointIndices[4]; /* pointIndices[3] == -1 for triangles */
	float texCoords[4][4];   /* u1, v1, u2, v2 per corner */
	float colors[4][4];      /* r, g, b, a per corner, 0..1 */
} dfpsr_polygon;

/* ref: implementation/render/ProjectedPoint.h:33-50 (40 bytes). */
typedef struct dfpsr_projected_point {
	float cs[3];
	float is[2];
	int32_t pad_;
	int64_t flat[2];
} dfpsr_projected_point;

/* One pre-projected triangle as given to renderer_giveTask_triangle (ref: api/rendererAPI.h:108-116). */
typedef struct dfpsr_triangle {
	dfpsr_projected_point pos[3];
	float colors[3][4];
	float texCoords[3][4];
} dfpsr_triangle;

/* Image view over DEVICE memory (ref: implementation/image/Image.h:58-183). stride in bytes. */
typedef struct dfpsr_image {
	void *data;
	int32_t width, height;
	int32_t stride;
	int32_t packOrder; /* RGBA8 images only */
} dfpsr_image;

/* Mip pyramid in one u32 buffer, smallest level first (ref: implementation/image/Texture.h:42-95).
 * data == NULL means "no texture". Use dfpsr_texture_layout() to fill the derived fields. */
typedef struct dfpsr_texture {
	const uint32_t *data; /* device pointer, totalPixels u32 */
	uint32_t log2width, log2height, maxMipLevel;
	uint32_t startOffset, maxLevelMask;
	uint32_t totalPixels;
} dfpsr_texture;

/* The three matrices of OrthoView that the light passes read (ref: SDK/SpriteEngine/orthoAPI.h:52-77). */
typedef struct dfpsr_ortho_view {
	dfpsr_matrix3x3 normalToWorldSpace;
	dfpsr_matrix3x3 screenDepthToLightSpace;
	dfpsr_matrix3x3 lightSpaceToScreenDepth;
} dfpsr_ortho_view;

/* A model on the device: points (3 floats each) and polygons; one part (ref: Model.h:66-84). */
typedef struct dfpsr_model {
	const float *points;           /* device, 3 * pointCount floats */
	int32_t pointCount;
	const dfpsr_polygon *polygons; /* device */
	int32_t polygonCount;
	int32_t filter;
	dfpsr_texture diffuse, light;
	float minBound[3], maxBound[3]; /* model space, ref: Model.h:83 */
} dfpsr_model;

/* Pre-compiled per-pixel operations for filter_mapRgbaU8 / filter_generateRgbaU8 (the general form is dfpsr_filter_map_program).
 * params are int32. Results are saturated to 0..255 and packed in the target's pack order
 * (ref: api/filterAPI.cpp:759-777, image_saturateAndPack). */
enum {
	/* (x & 255, y & 255, (x ^ y) & 255, 255) — BASELINE config 5 source pattern. No params. */
	DFPSR_MAP_XOR_PATTERN = 0,
	/* c' = src(x, y).c * mul[c] + add[c], source read with clamp-to-edge
	 * (ref: api/imageAPI.h image_readPixel_clamp). params = mul[4], add[4]. */
	DFPSR_MAP_AFFINE = 1,
	/* constant colour params[0..3] */
	DFPSR_MAP_CONSTANT = 2
};

/* ---------------------------------------------------------------- library */

int dfpsr_abi_version(void);
const char *dfpsr_last_error(void);
/* Selects the CUDA device for the calling thread and creates its context state. */
int dfpsr_init(int device);
int dfpsr_device_count(void);
/* Number of kernels this library has launched since the last reset (bench.py's gpu_launches). */
uint64_t dfpsr_launch_count(void);
void dfpsr_reset_launch_count(void);

/* Per-kernel device timing for bench.py's roofline: while enabled, every launch is bracketed by CUDA events on its
 * stream; dfpsr_profile_read returns the accumulated time and launch count per kernel name. */
int dfpsr_profile_enable(int enabled);
int dfpsr_profile_reset(void);
int dfpsr_profile_count(void);
int dfpsr_profile_read(int index, const char **name, double *milliseconds, int64_t *launches);

/* Device memory helpers for hosts that do not bring their own allocator. */
int dfpsr_malloc(void **devicePtr, size_t bytes);
int dfpsr_free(void *devicePtr);
int dfpsr_malloc_host(void **pinnedPtr, size_t bytes);
int dfpsr_free_host(void *pinnedPtr);
int dfpsr_upload(void *devicePtr, const void *hostPtr, size_t bytes, void *stream);
int dfpsr_download(void *hostPtr, const void *devicePtr, size_t bytes, void *stream);
int dfpsr_upload_2d(void *devicePtr, size_t deviceStride, const void *hostPtr, size_t hostStride, size_t rowBytes, size_t rows, void *stream);
int dfpsr_download_2d(void *hostPtr, size_t hostStride, const void *devicePtr, size_t deviceStride, size_t rowBytes, size_t rows, void *stream);
int dfpsr_stream_synchronize(void *stream);

/* ---------------------------------------------------------------- camera (host side, no GPU needed) */

/* ref: implementation/render/Camera.h:144-150 Camera::createPerspective */
int dfpsr_camera_create_perspective(dfpsr_camera *out, const dfpsr_transform3d *location, float imageWidth, float imageHeight, float widthSlope, float nearClip, float farClip);
/* ref: implementation/render/Camera.h:152-156 Camera::createOrthogonal */
int dfpsr_camera_create_orthogonal(dfpsr_camera *out, const dfpsr_transform3d *location, float imageWidth, float imageHeight, float halfWidth);
/* ref: implementation/render/Camera.h:202-217 Camera::isBoxSeen — 0 hidden, 1 partial, 2 fully inside. */
int dfpsr_camera_is_box_seen(const dfpsr_camera *camera, const float minBound[3], const float maxBound[3], const dfpsr_transform3d *modelToWorld);

/* ---------------------------------------------------------------- textures */

/* Fills log2/startOffset/maxLevelMask/totalPixels for a width x height texture with `resolutions`
 * levels (ref: api/textureAPI.cpp:65-78 + implementation/image/Texture.h:63-91). width and height are
 * rounded up to powers of two. data is left NULL. */
int dfpsr_texture_layout(dfpsr_texture *out, int32_t width, int32_t height, int32_t resolutions);
/* Box-filters every lower level from level 0 on the device (ref: api/textureAPI.cpp:44-87). */
int dfpsr_texture_generate_pyramid(const dfpsr_texture *texture, void *stream);
/* texture_create_RgbaU8(image, resolutions): bilinear resize of `image` into level 0 then pyramid
 * (ref: api/textureAPI.cpp:89-110). `texture->data` must already point at totalPixels u32. */
int dfpsr_texture_from_image(const dfpsr_texture *texture, const dfpsr_image *image, void *stream);

/* ---------------------------------------------------------------- triangle pipeline */

typedef struct dfpsr_renderer dfpsr_renderer;

/* ref: api/rendererAPI.h:56-58 renderer_create / (handle release) */
int dfpsr_renderer_create(dfpsr_renderer **out);
int dfpsr_renderer_destroy(dfpsr_renderer *renderer);
/* Precision of the interpolated 1/W, U/W, V/W of colour frames (no counterpart in the reference; SURVEY.md §7 hard part 3).
 *   DFPSR_PRECISION_EXACT (default): the reference's chains of float additions are replayed (shader/fillerTemplates.h:329-372), colour
 *     and depth are bit-identical to the reference's scalar build.
 *   DFPSR_PRECISION_TOLERANCE: the planes are evaluated directly at every quad (ITriangle2D.h:82-100) and 1/W uses the hardware
 *     reciprocal, like the rcpps + Newton step of the reference's SSE build (base/simd.h:4047-4052): identical coverage, colours within
 *     +-1 LSB per channel, depth within a few ulp (tests/test_gpu_tolerance.py states the bounds). Frames with alpha-filtered commands and
 *     depth-only frames always use the exact path.
 * dfpsr_set_default_precision applies to renderers created afterwards and to the calling thread's model_render* calls. */
enum { DFPSR_PRECISION_EXACT = 0, DFPSR_PRECISION_TOLERANCE = 1 };
int dfpsr_renderer_set_precision(dfpsr_renderer *renderer, int32_t precision);
int dfpsr_set_default_precision(int32_t precision);
/* Asynchronous frames (no counterpart in the reference, whose renderer_end blocks until the pixels are drawn).
 * By default dfpsr_renderer_end waits once, in the middle of the frame, for the counts of its set-up pass (commands, rows, tile list
 * entries) and sizes its device pools exactly. An asynchronous renderer launches the whole frame without waiting: the pools are sized
 * from the frames it has drawn before (its first frame still waits) and the kernels check them on the device. A frame that does not fit
 * draws NOTHING; the library notices when it next looks at the renderer — dfpsr_renderer_begin / _end / _flush / _destroy, and before
 * every kernel launch, upload, download or dfpsr_stream_synchronize this thread makes through this library — grows the pools and draws
 * the frame again before queueing anything else, so that every consumer that goes through this library sees finished pixels in stream
 * order. Work queued on the stream by other means (the caller's own cudaMemcpyAsync, other libraries) must call dfpsr_renderer_flush
 * (or dfpsr_flush) first. The model, texture and (for give_task_triangles) internal copies a frame reads must stay unchanged until then.
 * dfpsr_set_default_async applies to renderers created afterwards and to the calling thread's model_render* calls. */
int dfpsr_renderer_set_async(dfpsr_renderer *renderer, int32_t enabled);
int dfpsr_set_default_async(int32_t enabled);
int dfpsr_renderer_flush(dfpsr_renderer *renderer);
int dfpsr_flush(void);
/* ref: api/rendererAPI.h:66 renderer_begin. Either image may have data == NULL. Calling begin twice
 * without end is an error (ref: api/rendererAPI.cpp:152-154). */
int dfpsr_renderer_begin(dfpsr_renderer *renderer, const dfpsr_image *color, const dfpsr_image *depth);
/* image_fill(color, packedClearColor) + image_fill(depth, clearDepth) + renderer_begin fused (ref: SDK/terrain/main.cpp:397-416):
 * the tile kernel starts every tile from the clear values instead of loading it, so the two clears cost no memory pass. */
int dfpsr_renderer_begin_cleared(dfpsr_renderer *renderer, const dfpsr_image *color, const dfpsr_image *depth, uint32_t packedClearColor, float clearDepth);
/* Strip mode for multi-GPU frames: only target rows [top, bottom) are drawn by this renderer, exactly like one worker's clipBound in
 * CommandQueue::execute (ref: implementation/render/renderCore.cpp:459-478). Call between begin and end. top and bottom must be multiples
 * of 4 (the tile height; bottom may also be the image height). Pixels do not depend on the split. */
int dfpsr_renderer_set_clip_rows(dfpsr_renderer *renderer, int32_t top, int32_t bottom);
/* Occlusion grid (ref: api/rendererAPI.h:73-97, :131; api/rendererAPI.cpp:181-351, :403-477): 16x16-pixel cells holding the farthest
 * depth at which something may still be visible. Occluder boxes and visibility queries are evaluated on the host (the grid is a few
 * thousand floats); triangles and the depth buffer live on the device, so occlude_from_existing_triangles and occlude_from_top_rows
 * run a kernel and synchronise `stream` once. Commands hidden by the grid are skipped at renderer_end exactly like completeOcclusion
 * does (api/rendererAPI.cpp:193-217), and renderer_give_task skips whole models like model_render_threaded (api/modelAPI.cpp:229-234).
 * All of them must be called between renderer_begin and renderer_end with the frame's camera. */
int dfpsr_renderer_occlude_from_box(dfpsr_renderer *renderer, const float minBound[3], const float maxBound[3], const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera);
int dfpsr_renderer_occlude_from_top_rows(dfpsr_renderer *renderer, const dfpsr_camera *camera, void *stream);
int dfpsr_renderer_occlude_from_existing_triangles(dfpsr_renderer *renderer, void *stream);
int dfpsr_renderer_has_occluders(const dfpsr_renderer *renderer);
int dfpsr_renderer_is_box_visible(const dfpsr_renderer *renderer, const float minBound[3], const float maxBound[3], const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, int32_t *visible);
/* ref: api/modelAPI.cpp:214-281 model_render_threaded / renderer_giveTask. Bound culling
 * (Camera::isBoxSeen) is applied on the host exactly like the reference. The task records the model's DEVICE pointers (points,
 * polygons, textures); projection, set-up and drawing all run at dfpsr_renderer_end, so those buffers must stay alive and unchanged
 * until the frame has been drawn (the reference copies the vertex data at submission; dsr_b200.h keeps the buffers alive for its
 * callers). A frame may use up to 4096 textures. */
int dfpsr_renderer_give_task(dfpsr_renderer *renderer, const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, void *stream);
/* Many models in one call with the whole-model tests ON THE DEVICE (a broad phase for scenes of thousands of models):
 * the tests model_render_threaded makes per model on the host — Camera::isBoxSeen (api/modelAPI.cpp:228) and, when the renderer has occluders,
 * renderer_isBoxVisible (:229-234; api/rendererAPI.cpp:302-351) against the occluders given BEFORE this call — run once per set-up CTA of
 * each model with the same arithmetic, so the frame is identical to a loop of dfpsr_renderer_give_task. models / modelToWorld: HOST arrays. */
int dfpsr_renderer_give_tasks(dfpsr_renderer *renderer, const dfpsr_model *models, const dfpsr_transform3d *modelToWorld, int32_t count, const dfpsr_camera *camera, void *stream);
/* ref: api/rendererAPI.h:108-116 renderer_giveTask_triangle, batched: `triangles` is a HOST array. */
int dfpsr_renderer_give_task_triangles(dfpsr_renderer *renderer, const dfpsr_triangle *triangles, int32_t count, const dfpsr_texture *diffuse, const dfpsr_texture *light, int32_t filter, const dfpsr_camera *camera, void *stream);
/* ref: api/rendererAPI.h:131 renderer_end → CommandQueue::execute (implementation/render/renderCore.cpp:449-480):
 * bins the queued triangles to screen tiles and rasterises/shades every tile in submission order. Everything is queued on `stream`;
 * the call waits for the stream once (for the set-up counts) unless the renderer is asynchronous (dfpsr_renderer_set_async). */
int dfpsr_renderer_end(dfpsr_renderer *renderer, void *stream);
/* ref: api/rendererAPI.h:134-135, api/rendererAPI.cpp:362-399 renderer_end(renderer, debugWireframe = true): when enabled, the NEXT
 * dfpsr_renderer_end draws the edges of every command that the occlusion grid did not remove as white lines on top of the finished colour
 * buffer (the reference's draw_line between the corners' whole-pixel positions). The flag resets after that frame; such a frame always
 * waits for its set-up counts. */
int dfpsr_renderer_set_debug_wireframe(dfpsr_renderer *renderer, int32_t enabled);
/* Number of draw commands (post-clipping triangles) the last frame produced; synchronises `stream`. */
int dfpsr_renderer_last_command_count(dfpsr_renderer *renderer, int64_t *count, void *stream);

/* ref: api/modelAPI.cpp:197-201 model_render — begin + give_task + end on a private renderer. */
int dfpsr_model_render(const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_image *color, const dfpsr_image *depth, const dfpsr_camera *camera, void *stream);
/* ref: api/modelAPI.cpp:202-206 model_renderDepth (1x1 aligned depth-only path, renderCore.cpp:343-443). */
int dfpsr_model_render_depth(const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_image *depth, const dfpsr_camera *camera, void *stream);
/* Many independent views of one model in one submission (BASELINE config 4): view i renders with
 * cameras[i] into colors[i]/depths[i]. Equivalent to `count` calls of dfpsr_model_render. When
 * `clear` is non-zero every target is first cleared to colour 0 / depth 0.0f as the SDK terrain loop does
 * (ref: SDK/terrain/main.cpp:397-402). Arrays are HOST arrays of descriptors. */
int dfpsr_model_render_views(const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_image *colors, const dfpsr_image *depths, const dfpsr_camera *cameras, int32_t count, int32_t clear, void *stream);
/* Many model_renderDepth calls in one submission (the Sandbox shadow pass: 6 cube faces x shadow casters x lights, ref:
 * SDK/SpriteEngine/spriteAPI.cpp:389-401, :793-804): task i renders models[i] with modelToWorld[i] and cameras[i] into depths[targetOfTask[i]],
 * tasks that share a target are applied in array order. When `clear` is non-zero every target is first filled with clearDepth
 * (CubeMapF32::clear, spriteAPI.cpp:360-362) inside the same kernels. All arrays are HOST arrays. */
int dfpsr_model_render_depth_batch(const dfpsr_model *const *models, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *cameras, const int32_t *targetOfTask, int32_t taskCount, const dfpsr_image *depths, int32_t targetCount, int32_t clear, float clearDepth, void *stream);
/* ref: api/modelAPI.cpp:238-242 — the projection loop alone (exposed for parity tests): out[i] = worldToScreen(M * p[i]). */
int dfpsr_project_points(const float *points, int32_t count, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, dfpsr_projected_point *outDevice, void *stream);

/* ---------------------------------------------------------------- 2D draw calls on the path */

/* ref: api/imageAPI.cpp:167-185 image_fill → api/drawAPI.cpp:72-174 draw_rectangle. */
int dfpsr_image_fill_rgba(const dfpsr_image *image, int32_t red, int32_t green, int32_t blue, int32_t alpha, void *stream);
int dfpsr_image_fill_f32(const dfpsr_image *image, float value, void *stream);
/* ref: api/drawAPI.cpp:492-538, :906-924 draw_copy (RgbaU8→RgbaU8 incl. pack-order conversion; F32→F32). */
int dfpsr_draw_copy_rgba(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, void *stream);
int dfpsr_draw_copy_f32(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, void *stream);
/* ref: api/drawAPI.cpp:834-904, :962-979 draw_higher on F32 heights with 0, 1 or 2 RGBA8 payload images
 * (targetA/sourceA and targetB/sourceB may be NULL together). */
int dfpsr_draw_higher(const dfpsr_image *targetHeight, const dfpsr_image *sourceHeight, const dfpsr_image *targetA, const dfpsr_image *sourceA, const dfpsr_image *targetB, const dfpsr_image *sourceB, int32_t left, int32_t top, float sourceHeightOffset, void *stream);

/* The remaining draw calls on RGBA8 / F32 device images, so that overlays and debug drawing stay on the device.
 * ref: api/drawAPI.cpp:72-174 draw_rectangle (clipped; colour saturated and packed in the image's pack order), :176-310 draw_line
 * (the reference's error-accumulating line, every step evaluated in closed form), :636-661 draw_alphaFilter, :663-698 draw_maxAlpha,
 * :700-715 draw_alphaClip, :717-757 draw_silhouette (source = 8-bit image: 1 byte per pixel, stride in bytes). */
int dfpsr_draw_rectangle_rgba(const dfpsr_image *image, int32_t left, int32_t top, int32_t width, int32_t height, const int32_t colorRgba[4], void *stream);
int dfpsr_draw_rectangle_f32(const dfpsr_image *image, int32_t left, int32_t top, int32_t width, int32_t height, float value, void *stream);
int dfpsr_draw_line_rgba(const dfpsr_image *image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, const int32_t colorRgba[4], void *stream);
int dfpsr_draw_line_f32(const dfpsr_image *image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, float value, void *stream);
int dfpsr_draw_alpha_filter(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, void *stream);
int dfpsr_draw_max_alpha(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, int32_t sourceAlphaOffset, void *stream);
int dfpsr_draw_alpha_clip(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, int32_t threshold, void *stream);
int dfpsr_draw_silhouette(const dfpsr_image *target, const dfpsr_image *silhouetteU8, const int32_t colorRgba[4], int32_t left, int32_t top, void *stream);

/* 8-bit and 16-bit monochrome images and the conversions between formats (ref: api/drawAPI.h:68-103, :128-135): the descriptor is the
 * same dfpsr_image (stride in bytes, packOrder ignored for monochrome); the pixel format is passed next to it. */
enum { DFPSR_FORMAT_U8 = 1, DFPSR_FORMAT_U16 = 2, DFPSR_FORMAT_F32 = 3, DFPSR_FORMAT_RGBA_U8 = 4 };
int dfpsr_draw_rectangle_mono(const dfpsr_image *image, int32_t format, int32_t left, int32_t top, int32_t width, int32_t height, int32_t color, void *stream);
int dfpsr_draw_line_mono(const dfpsr_image *image, int32_t format, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int32_t color, void *stream);
/* All thirteen draw_copy overloads (ref: api/drawAPI.cpp:497-634), including their conversions as the reference performs them: luma
 * replicated with alpha 255 into RGBA, saturateFloat rounding for F32 -> 8 bits, U16 clamped to 255 on its way to U8 / F32 / RGBA. */
int dfpsr_draw_copy_formats(const dfpsr_image *target, int32_t targetFormat, const dfpsr_image *source, int32_t sourceFormat, int32_t left, int32_t top, void *stream);
/* ref: api/drawAPI.cpp:759-832 draw_higher on 16-bit heights (0 = empty) with 0, 1 or 2 RGBA payload images. */
int dfpsr_draw_higher_u16(const dfpsr_image *targetHeight, const dfpsr_image *sourceHeight, const dfpsr_image *targetA, const dfpsr_image *sourceA, const dfpsr_image *targetB, const dfpsr_image *sourceB, int32_t left, int32_t top, int32_t sourceHeightOffset, void *stream);

/* One sprite placement for the batched compositor: sources live in an atlas on the device. */
typedef struct dfpsr_sprite_draw {
	dfpsr_image sourceHeight, sourceA, sourceB;
	int32_t left, top;
	float heightOffset;
} dfpsr_sprite_draw;
/* `count` draw_higher calls applied in array order (ref: SDK/SpriteEngine/spriteAPI.cpp:316-323, :525-539
 * — BackgroundBlock::draw loops drawSprite over the octree result). HOST array. */
int dfpsr_draw_higher_batch(const dfpsr_image *targetHeight, const dfpsr_image *targetA, const dfpsr_image *targetB, const dfpsr_sprite_draw *draws, int32_t count, void *stream);

/* ---------------------------------------------------------------- Sandbox deferred light */

/* ref: SDK/SpriteEngine/lightAPI.cpp:23-74 setDirectedLight (add = 0) / addDirectedLight (add = 1). */
int dfpsr_light_directed(const dfpsr_ortho_view *view, const dfpsr_image *light, const dfpsr_image *normal, const float direction[3], float intensity, const int32_t colorRgb[3], int32_t add, void *stream);
/* ref: SDK/SpriteEngine/lightAPI.cpp:76-285 addPointLight. shadowCubeMap may be NULL (no shadows);
 * otherwise a width x 6*width F32 image rendered by dfpsr_model_render_depth (spriteAPI.cpp:351-401). */
int dfpsr_light_point(const dfpsr_ortho_view *view, const int32_t worldCenter[2], const dfpsr_image *light, const dfpsr_image *normal, const dfpsr_image *height, const float position[3], float radius, float intensity, const int32_t colorRgb[3], const dfpsr_image *shadowCubeMap, void *stream);
/* ref: SDK/SpriteEngine/lightAPI.cpp:287-323 blendLight. */
int dfpsr_light_blend(const dfpsr_image *color, const dfpsr_image *diffuse, const dfpsr_image *light, void *stream);

/* One Sandbox frame's light passes in a single kernel (ref: SDK/SpriteEngine/spriteAPI.cpp:775-814, the part of SpriteWorldImpl::draw
 * after drawDeferred): the first directed light overwrites the light buffer (no directed light = black, :783-786), the others and
 * every point light are added with saturation, then colour = blendLight(diffuse, light) when `color` is given. Same bytes as the
 * separate calls above; every buffer is read or written once per pixel. Shadow cube maps must already be rendered
 * (dfpsr_model_render_depth_batch). directed / points are HOST arrays. */
typedef struct dfpsr_directed_light { float direction[3]; float intensity; int32_t colorRgb[3]; } dfpsr_directed_light;
typedef struct dfpsr_point_light { float position[3]; float radius, intensity; int32_t colorRgb[3]; dfpsr_image shadowCubeMap; /* data == NULL: no shadows */ } dfpsr_point_light;
int dfpsr_light_frame(const dfpsr_ortho_view *view, const int32_t worldCenter[2], const dfpsr_image *color, const dfpsr_image *diffuse, const dfpsr_image *light, const dfpsr_image *normal, const dfpsr_image *height, const dfpsr_directed_light *directed, int32_t directedCount, const dfpsr_point_light *points, int32_t pointCount, void *stream);

/* ---------------------------------------------------------------- Sandbox sprite engine: views, dense models, sprite world */

/* ref: SDK/SpriteEngine/orthoAPI.h:40-83 OrthoView, all derived members. */
typedef struct dfpsr_ortho_camera {
	int32_t id, worldDirection;
	dfpsr_matrix3x3 normalToWorldSpace;
	int32_t pixelOffsetPerTileX[2], pixelOffsetPerTileZ[2], yPixelsPerTile;
	dfpsr_matrix3x3 screenDepthToWorldSpace, worldSpaceToScreenDepth, screenDepthToLightSpace, lightSpaceToScreenDepth;
	float roundedScreenPixelsToWorldTiles[4]; /* FMatrix2x2: xAxis.x, xAxis.y, yAxis.x, yAxis.y */
} dfpsr_ortho_camera;
/* ref: SDK/SpriteEngine/orthoAPI.h:92-134 OrthoSystem (8 fixed camera angles). */
typedef struct dfpsr_ortho_system {
	float cameraTilt;
	int32_t pixelsPerTile;
	dfpsr_ortho_camera view[8];
} dfpsr_ortho_system;
/* ref: SDK/SpriteEngine/orthoAPI.cpp:5-32, :82-119 OrthoSystem(cameraTilt, pixelsPerTile) -> update(). Host arithmetic, no GPU needed. */
int dfpsr_ortho_system_create(dfpsr_ortho_system *out, float cameraTilt, int32_t pixelsPerTile);
/* The three matrices the light passes read (dfpsr_ortho_view) taken from a full view. */
int dfpsr_ortho_camera_light_view(const dfpsr_ortho_camera *camera, dfpsr_ortho_view *out);

/* ref: SDK/SpriteEngine/spriteAPI.cpp:236-249 DenseTriangle (108 bytes): colours 0..255, object-space positions, smooth normals. */
typedef struct dfpsr_dense_triangle {
	float colorA[3], colorB[3], colorC[3];
	float posA[3], posB[3], posC[3];
	float normalA[3], normalB[3], normalC[3];
} dfpsr_dense_triangle;
/* ref: SDK/SpriteEngine/spriteAPI.cpp:1176-1233 DenseModel_create: smooth per-point normals, polygons fanned into triangles (0, b, b + 1),
 * vertex colours scaled by 255. Host arithmetic (set-up time). `out` holds dfpsr_dense_model_triangle_count() HOST triangles; the bounds
 * are the model's bounding box (ref: implementation/render/model/Model.cpp:281-288, always contains the origin). */
int32_t dfpsr_dense_model_triangle_count(const dfpsr_polygon *polygons, int32_t polygonCount);
int dfpsr_dense_model_build(const float *points, int32_t pointCount, const dfpsr_polygon *polygons, int32_t polygonCount, dfpsr_dense_triangle *out, float minBound[3], float maxBound[3]);
/* ref: SDK/SpriteEngine/spriteAPI.cpp:1243-1327 renderDenseModel<HIGH_QUALITY>: orthogonal vertex-colour triangles with float barycentric
 * weights (tolerance -0.00001), height test `>`, writing height + diffuse + normal (RGBA order). `triangles` is a DEVICE array.
 * dirtyRect receives {left, top, width, height} of the pessimistic bound the reference returns (all zero when culled); may be NULL. */
int dfpsr_dense_model_render(const dfpsr_dense_triangle *triangles, int32_t triangleCount, const float minBound[3], const float maxBound[3], const dfpsr_ortho_camera *view, const dfpsr_image *height, const dfpsr_image *diffuse, const dfpsr_image *normal, const float worldOrigin[2], const dfpsr_transform3d *modelToWorld, int32_t highQuality, int32_t dirtyRect[4], void *stream);

/* ref: SDK/SpriteEngine/spriteAPI.cpp:1329-1432 sprite_generateFromModel: renders a dense model (HOST triangles from dfpsr_dense_model_build) with
 * renderDenseModel<true> from `cameraAngles` of the system's views into a worst-case square image, converts heights to 8 bits, crops all angles
 * uniformly to the drawn pixels and packs [colour | height | normal] x angles into an atlas. The atlas is a DEVICE image allocated by the call
 * (release atlas.data with dfpsr_free); atlas.data stays NULL when nothing is visible. The remaining fields are the SpriteConfig the reference
 * writes to the .ini (the caller appends its shadow model). Synchronises `stream`. */
typedef struct dfpsr_baked_sprite {
	dfpsr_image atlas;
	int32_t centerX, centerY, frameRows, propertyColumns;
	float minBound[3], maxBound[3];
} dfpsr_baked_sprite;
int dfpsr_sprite_generate_from_model(const dfpsr_dense_triangle *triangles, int32_t triangleCount, const float minBound[3], const float maxBound[3], const dfpsr_ortho_system *ortho, int32_t cameraAngles, dfpsr_baked_sprite *out, void *stream);

/* Sprite types are process-global like the reference's (ref: SDK/SpriteEngine/spriteAPI.cpp:279-289). The reference loads
 * <name>.png + <name>.ini; here the decoded atlas (RGBA order, HOST memory) and the parsed configuration are passed in
 * (ref: spriteAPI.cpp:47-133 SpriteConfig, :190-232 SpriteType): the atlas holds frameRows rows of [colour | height | normal ...]
 * columns; heights become (red * (maxBound.y - minBound.y) / 255 + minBound.y) where the colour's alpha > 127, -inf elsewhere
 * (:157-174 scaleHeightImage, evaluated on the device). points / triangleIndices describe the optional shadow model. */
typedef struct dfpsr_sprite_config {
	int32_t centerX, centerY, frameRows, propertyColumns;
	float minBound[3], maxBound[3];
	const float *points; int32_t pointCount;                   /* 3 floats per point */
	const int32_t *triangleIndices; int32_t triangleIndexCount; /* multiples of three */
} dfpsr_sprite_config;
int dfpsr_sprite_type_create(const uint32_t *atlasHost, int32_t width, int32_t height, int32_t strideBytes, const dfpsr_sprite_config *config, int32_t *typeIndex);
int32_t dfpsr_sprite_type_count(void);
/* ref: SDK/SpriteEngine/spriteAPI.cpp:262-277, :291-299 ModelType(visibleModel, shadowModel): dense triangles (HOST, from
 * dfpsr_dense_model_build) + an optional shadow model (HOST geometry; only points and polygons are read). */
struct dfpsr_host_model; /* declared with the host-buffer entry points below */
int dfpsr_model_type_create(const dfpsr_dense_triangle *triangles, int32_t triangleCount, const float minBound[3], const float maxBound[3], const struct dfpsr_host_model *shadowModel, int32_t *typeIndex);
int32_t dfpsr_model_type_count(void);

/* ref: SDK/SpriteEngine/spriteAPI.h:24-53 */
typedef struct dfpsr_sprite_instance {
	int32_t typeIndex, direction;
	int32_t location[3]; /* mini-tile units (1024 per tile) */
	int32_t shadowCasting;
	uint64_t userData;
} dfpsr_sprite_instance;
typedef struct dfpsr_model_instance {
	int32_t typeIndex;
	dfpsr_transform3d location; /* tile units */
	uint64_t userData;
} dfpsr_model_instance;

/* The host side of SpriteWorldImpl (ref: SDK/SpriteEngine/spriteAPI.cpp:572-816): octrees of passive sprites and models, 512x512
 * background blocks, dirty rectangles, temporary sprites / models / lights and the pass order of a frame. A frame is first PLANNED on
 * the host into a list of operations (this is where the reference's control flow lives) and then EXECUTED on the device with the
 * kernels above: block (re)generation and temporary sprites through dfpsr_draw_higher_batch / dfpsr_dense_model_render, all
 * background copies of the frame in one kernel, every shadow cube map of the frame in one dfpsr_model_render_depth_batch and all
 * lights plus blendLight in one dfpsr_light_frame. */
typedef struct dfpsr_sprite_world dfpsr_sprite_world;
enum {
	DFPSR_SW_BLOCK_CLEAR = 1,   /* block: diffuse = 0, normal = (128,128,128,128), height = -1000000 (spriteAPI.cpp:521-523, :545-551) */
	DFPSR_SW_BLOCK_SPRITE = 2,  /* draw_higher of sprite frame (typeIndex, frame) into block at (left, top) with heightOffset */
	DFPSR_SW_BLOCK_MODEL = 3,   /* renderDenseModel<false> of model typeIndex into block with worldOrigin / transform */
	DFPSR_SW_COPY_BLOCK = 4,    /* draw_copy x3: block pixels from (sourceLeft, sourceTop), size (width, height), to frame buffers at (left, top) */
	DFPSR_SW_SPRITE = 5,        /* temporary sprite into the frame buffers */
	DFPSR_SW_MODEL = 6,         /* temporary model into the frame buffers */
	DFPSR_SW_LIGHT_CLEAR = 7,   /* no directed light: light buffer = 0 */
	DFPSR_SW_LIGHT_DIRECTED = 8,/* light = index into the directed lights; flag = 1 overwrite, 0 add */
	DFPSR_SW_SHADOW_CLEAR = 9,  /* cube map of point light `light` = 0 */
	DFPSR_SW_SHADOW_SPRITE = 10,/* 6 x model_renderDepth of sprite type typeIndex's shadow model with `transform` (relative to the light); flag = 1: a temporary caster */
	DFPSR_SW_SHADOW_MODEL = 11, /* the same for model type typeIndex */
	DFPSR_SW_LIGHT_POINT = 12,  /* addPointLight of light `light`; flag = 1 with its shadow cube map */
	DFPSR_SW_BLEND = 13         /* blendLight into the colour target */
};
typedef struct dfpsr_sprite_world_op {
	int32_t op;
	int32_t block;                 /* background block slot */
	int32_t typeIndex, frame;
	int32_t left, top, width, height;
	int32_t sourceLeft, sourceTop;
	float heightOffset;
	float worldOrigin[2];
	dfpsr_transform3d transform;
	int32_t light, flag;
} dfpsr_sprite_world_op;

int dfpsr_sprite_world_create(dfpsr_sprite_world **out, const dfpsr_ortho_system *ortho, int32_t shadowResolution); /* ref: spriteAPI.cpp:818 */
int dfpsr_sprite_world_destroy(dfpsr_sprite_world *world);
int dfpsr_sprite_world_add_background_sprite(dfpsr_sprite_world *world, const dfpsr_sprite_instance *sprite); /* ref: spriteAPI.cpp:900-916 */
int dfpsr_sprite_world_add_background_model(dfpsr_sprite_world *world, const dfpsr_model_instance *model);    /* ref: spriteAPI.cpp:918-937 */
int dfpsr_sprite_world_add_temporary_sprite(dfpsr_sprite_world *world, const dfpsr_sprite_instance *sprite);  /* ref: spriteAPI.cpp:975-980 */
int dfpsr_sprite_world_add_temporary_model(dfpsr_sprite_world *world, const dfpsr_model_instance *model);     /* ref: spriteAPI.cpp:982-986 */
/* ref: spriteAPI.cpp:939-973. filter == NULL erases everything touching the search box; otherwise it returns non-zero to erase. */
typedef int (*dfpsr_sprite_selection)(dfpsr_sprite_instance *sprite, const int32_t origin[3], const int32_t minBound[3], const int32_t maxBound[3], void *user);
typedef int (*dfpsr_model_selection)(dfpsr_model_instance *model, const int32_t origin[3], const int32_t minBound[3], const int32_t maxBound[3], void *user);
int dfpsr_sprite_world_remove_background_sprites(dfpsr_sprite_world *world, const int32_t searchMin[3], const int32_t searchMax[3], dfpsr_sprite_selection filter, void *user);
int dfpsr_sprite_world_remove_background_models(dfpsr_sprite_world *world, const int32_t searchMin[3], const int32_t searchMax[3], dfpsr_model_selection filter, void *user);
int dfpsr_sprite_world_create_temporary_point_light(dfpsr_sprite_world *world, const float position[3], float radius, float intensity, const int32_t colorRgb[3], int32_t shadowCasting); /* ref: spriteAPI.cpp:988-991 */
int dfpsr_sprite_world_create_temporary_directed_light(dfpsr_sprite_world *world, const float direction[3], float intensity, const int32_t colorRgb[3]); /* ref: spriteAPI.cpp:993-996 */
int dfpsr_sprite_world_clear_temporary(dfpsr_sprite_world *world); /* ref: spriteAPI.cpp:998-1004 */
/* Camera (ref: spriteAPI.cpp:1050-1112). */
int dfpsr_sprite_world_get_camera_location(const dfpsr_sprite_world *world, int32_t location[3]);
int dfpsr_sprite_world_set_camera_location(dfpsr_sprite_world *world, const int32_t location[3]);
int dfpsr_sprite_world_move_camera_in_pixels(dfpsr_sprite_world *world, int32_t offsetX, int32_t offsetY);
int dfpsr_sprite_world_get_camera_direction_index(const dfpsr_sprite_world *world, int32_t *index);
int dfpsr_sprite_world_set_camera_direction_index(dfpsr_sprite_world *world, int32_t index);
int dfpsr_sprite_world_find_ground_at_pixel(const dfpsr_sprite_world *world, int32_t targetWidth, int32_t targetHeight, int32_t pixelX, int32_t pixelY, int32_t location[3]);
/* Plans one frame for a width x height colour target exactly like SpriteWorldImpl::draw (ref: spriteAPI.cpp:754-816) and advances the
 * world's state (block cache, dirty rectangles) WITHOUT touching the GPU. *ops stays valid until the next call on this world.
 * Used by dfpsr_sprite_world_draw and by the host-logic parity tests, which replay the operations with the CPU oracle. */
int dfpsr_sprite_world_plan_frame(dfpsr_sprite_world *world, int32_t width, int32_t height, const dfpsr_sprite_world_op **ops, int32_t *opCount);
/* spriteWorld_draw (ref: spriteAPI.cpp:1006-1009): plans the frame and executes it on the device into `colorTarget` (DEVICE image). */
int dfpsr_sprite_world_draw(dfpsr_sprite_world *world, const dfpsr_image *colorTarget, void *stream);
/* The same with a HOST colour image: the blended frame is copied back and the stream synchronised. */
int dfpsr_sprite_world_draw_host(dfpsr_sprite_world *world, uint32_t *colorHost, int32_t strideBytes, int32_t width, int32_t height, int32_t packOrder, void *stream);
/* ref: spriteAPI.cpp:1078-1096 spriteWorld_get{Diffuse,Normal,Light,Height}Buffer — DEVICE images of the last frame (data == NULL before the first draw). */
int dfpsr_sprite_world_get_buffers(const dfpsr_sprite_world *world, dfpsr_image *diffuse, dfpsr_image *normal, dfpsr_image *light, dfpsr_image *height);

/* ---------------------------------------------------------------- filters */

/* ref: api/filterAPI.cpp:852-860 filter_resize(ImageRgbaU8): `target` (newWidth x newHeight, RGBA order,
 * not a sub-image) receives the stretched `source`. sourceIsSubImage mirrors image_isSubImage(source),
 * which selects the reference's code path and therefore its rounding (filterAPI.cpp:262-279).
 * scratch: device buffer of dfpsr_filter_resize_scratch_bytes() bytes, only used for two-pass up-scaling. */
size_t dfpsr_filter_resize_scratch_bytes(int32_t sourceWidth, int32_t sourceHeight, int32_t newWidth, int32_t newHeight);
int dfpsr_filter_resize(const dfpsr_image *target, const dfpsr_image *source, int32_t sampler, int32_t sourceIsSubImage, void *scratch, void *stream);
/* ref: api/filterAPI.h:42, api/filterAPI.cpp:95-154, :282-314 filter_resize(ImageU8): one byte per pixel (stride in bytes, packOrder
 * ignored). scratch: device buffer of target.width * source.height BYTES, only used when the width changes and the height grows (the
 * reference's two passes and their two roundings). */
int dfpsr_filter_resize_u8(const dfpsr_image *target, const dfpsr_image *source, int32_t sampler, void *scratch, void *stream);
/* ref: api/filterAPI.h:54-79, api/filterAPI.cpp:759-782 filter_mapRgbaU8 / filter_generateRgbaU8 with ANY per-pixel function.
 * The reference takes a host lambda `ColorRgbaI32 f(int32_t x, int32_t y)`; here the function travels as text: `body` is the body of
 *     int4 pixel(int x, int y)      // (red, green, blue, alpha) as ints, saturated to 0..255 and packed in the target's order afterwards
 * in CUDA C++, compiled for the current device with NVRTC on first use and cached by its text. x and y are target coordinates plus
 * startX / startY. Up to 4 source images (device memory, any pack order) are read inside the body with read_clamp(i, x, y),
 * read_border(i, x, y[, int4 border]), read_tile(i, x, y) -> int4 (r, g, b, a) — the rules of image_readPixel_clamp / _border / _tile —
 * and source_width(i) / source_height(i). Example, the brighter image of api/filterAPI.h:67-73:
 *     "int4 s = read_clamp(0, x, y); return make_int4(s.x * 2, s.y * 2, s.z * 2, s.w);"
 * Needs libnvrtc.so.12 of the CUDA toolkit at run time. The enumerated ops of dfpsr_filter_map are pre-compiled instances. */
int dfpsr_filter_map_program(const dfpsr_image *target, const char *body, const dfpsr_image *sources, int32_t sourceCount, int32_t startX, int32_t startY, void *stream);
/* ref: api/filterAPI.cpp:759-782 filter_mapRgbaU8 / filter_generateRgbaU8 with one of the pre-compiled device ops above. */
int dfpsr_filter_map(const dfpsr_image *target, int32_t op, const int32_t *params, int32_t paramCount, const dfpsr_image *source, int32_t startX, int32_t startY, void *stream);
/* ref: implementation/gui/DsrWindow.cpp:255-281 DsrWindow::showCanvas + the back-end's canvas (windowManagers/X11Window.cpp:838-858): the
 * device canvas is block-magnified by pixelScale into the window's canvas — HOST memory of hostWidth x hostHeight pixels in the window
 * system's pack order (BGRA on X11 and Win32) — with the rules of filter_blockMagnify (partial pixels cut, transparent black beyond the
 * source) and arrives there with one device-to-host copy; the call returns when the canvas rows are complete. */
int dfpsr_canvas_show(const dfpsr_image *deviceCanvas, int32_t pixelScale, void *hostCanvas, int32_t hostStrideBytes, int32_t hostWidth, int32_t hostHeight, int32_t hostPackOrder, void *stream);
/* ref: api/filterAPI.cpp:724-757, :872-876 filter_blockMagnify. */
int dfpsr_filter_block_magnify(const dfpsr_image *target, const dfpsr_image *source, int32_t pixelWidth, int32_t pixelHeight, void *stream);

/* ---------------------------------------------------------------- host-buffer entry points (end to end) */

/* The same operations taking HOST images, as the reference API does: host→device copies, kernels,
 * device→host copies, stream synchronised on return. Used by the dsr:: shim and by bench.py's e2e leg. */
typedef struct dfpsr_host_model {
	const float *points; int32_t pointCount;
	const dfpsr_polygon *polygons; int32_t polygonCount;
	int32_t filter;
	const uint32_t *diffusePixels; /* whole pyramid, layout from dfpsr_texture_layout; may be NULL */
	dfpsr_texture diffuseLayout;
	const uint32_t *lightPixels;
	dfpsr_texture lightLayout;
	float minBound[3], maxBound[3];
} dfpsr_host_model;

typedef struct dfpsr_session dfpsr_session;
/* A session owns device mirrors of host objects so that repeated frames only move what changed. */
int dfpsr_session_create(dfpsr_session **out);
int dfpsr_session_destroy(dfpsr_session *session);
/* Uploads (or re-uploads) a model's geometry and textures; returns a model slot id >= 0 in *slot. */
int dfpsr_session_upload_model(dfpsr_session *session, const dfpsr_host_model *model, int32_t *slot);
/* Clears (colour 0, depth 0), renders model `slot` and downloads colour (and depth if depthHost != NULL)
 * into HOST buffers — one SDK terrain frame (ref: SDK/terrain/main.cpp:397-421). */
int dfpsr_session_render_frame_host(dfpsr_session *session, int32_t slot, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, uint32_t *colorHost, int32_t colorStride, float *depthHost, int32_t depthStride, int32_t width, int32_t height, int32_t packOrder, int32_t uploadGeometry, void *stream);

/* The same for `count` views of one model (BASELINE config 4 end to end): colorHost / depthHost are HOST arrays of `count` HOST image
 * pointers (depthHost or single entries may be NULL). Rendering of the next chunk of views overlaps the device-to-host copy of the
 * previous one; pinned host images (dfpsr_malloc_host) make the copies asynchronous. Returns after everything has arrived. */
int dfpsr_session_render_views_host(dfpsr_session *session, int32_t slot, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *cameras, int32_t count, uint32_t *const *colorHost, int32_t colorStride, float *const *depthHost, int32_t depthStride, int32_t width, int32_t height, int32_t packOrder, int32_t uploadGeometry, void *stream);

/* ---------------------------------------------------------------- scene formats feeding the path (host side, no GPU needed) */

/* One part of an imported model: polygons [firstPolygon, firstPolygon + polygonCount) and the texture names its shader asks for
 * (DMF1: M_Diffuse_1Tex / M_Diffuse_2Tex -> Texture[0] / Texture[1]; empty when absent, ref: dmf1.cpp:336-347). */
typedef struct dfpsr_imported_part {
	char name[64], diffuseName[64], lightName[64];
	int32_t firstPolygon, polygonCount;
} dfpsr_imported_part;
/* Host arrays in the layout dfpsr_host_model / dfpsr_model take; released with dfpsr_import_free. */
typedef struct dfpsr_imported_model {
	float *points; int32_t pointCount;
	dfpsr_polygon *polygons; int32_t polygonCount;
	dfpsr_imported_part *parts; int32_t partCount;
	int32_t filter;
	float minBound[3], maxBound[3];
} dfpsr_imported_model;
/* ref: SDK/SpriteEngine/importer.cpp:52-290 importer_loadModel for ASCII PLY 1.0 (vertex x/y/z/red/green/blue/alpha as float or uchar,
 * face vertex_indices lists: quads kept, other polygons fanned; flipX mirrors and reverses the winding; axisConversion may be NULL). */
int dfpsr_import_ply(const char *content, size_t length, int32_t flipX, const dfpsr_transform3d *axisConversion, dfpsr_imported_model *out);
/* ref: DFPSR/implementation/render/model/format/dmf1.cpp importFromContent_DMF1(content, pool, detailLevel): parts within the detail
 * level, points merged within 0.00001, per-vertex colours and two texture coordinate sets. Textures are returned by NAME (the
 * reference resolves them through a ResourcePool; decoding image files is outside the path). */
int dfpsr_import_dmf1(const char *content, size_t length, int32_t detailLevel, dfpsr_imported_model *out);
void dfpsr_import_free(dfpsr_imported_model *model);

/* Diagnostic: the point light's reciprocal square root (ref: base/simd.h:4104, (float)(1.0 / sqrt((double)x)) in the scalar build) is evaluated
 * on the device without FP64 sqrt / division plus an exact fallback; this compares it with the literal expression for the `count` floats whose
 * bit patterns start at firstBits and returns the number of differing results (must be 0). */
int dfpsr_selftest_rsqrt(uint32_t firstBits, uint32_t count, uint64_t *mismatchesHost, void *stream);

/* ---------------------------------------------------------------- strip-sharded frames over NVLink peer memory */

/* One frame split into row strips across GPUs (one process per GPU): the reference's workers all write their strip into the same target
 * (ref: implementation/render/renderCore.cpp:449-480); here the presenting rank owns the frame, exports it with dfpsr_peer_alloc, every
 * other rank maps it with dfpsr_peer_open and uses the mapped pointer as the colour target of dfpsr_renderer_begin[_cleared] +
 * dfpsr_renderer_set_clip_rows, so the tile kernel's own stores deliver the strip through NVLink/NVSwitch — no staging copy, no collective.
 * Hand-shake, all stream-ordered and without the host: a rank ends its frame with dfpsr_peer_signal(flag in the presenter's memory, frame
 * number); the presenter's stream runs dfpsr_peer_wait on those flags before it consumes the frame, then signals "consumed" flags in the
 * ranks' own memory, on which they wait before they overwrite the frame with the next one. */
#define DFPSR_PEER_HANDLE_BYTES 64
#define DFPSR_PEER_MAX_RANKS 16
/* cudaMalloc (zero-filled) + IPC export; handle receives DFPSR_PEER_HANDLE_BYTES bytes to send to the other processes. */
int dfpsr_peer_alloc(void **devicePtr, size_t bytes, uint8_t *handle);
int dfpsr_peer_free(void *devicePtr);
/* Maps an exported allocation of another process (same or other GPU of the box) into this one; close before the owner frees it. */
int dfpsr_peer_open(void **devicePtr, const uint8_t *handle);
int dfpsr_peer_close(void *devicePtr);
/* After everything queued on `stream` so far (peer stores included) is visible system-wide, stores `value` to each of the `count` flags. */
int dfpsr_peer_signal(uint32_t *const *flags, int32_t count, uint32_t value, void *stream);
/* Blocks `stream` (not the host) until flags[0..count) have all reached `value` (wrap-safe >=). Gives up after timeoutMs (1..10000)
 * instead of hanging the device: status (two u32 of device memory, zero at first; give every wait site its own) then counts the wait in
 * status[0] and keeps the value it was waiting for in status[1]. A caller that finds status[0] != 0 must treat the frames since its
 * last check as torn and resynchronise with its peers; dfpsr_peer_reset_status clears the block (stream-ordered). */
int dfpsr_peer_wait(const uint32_t *flags, int32_t count, uint32_t value, uint32_t timeoutMs, uint32_t *status, void *stream);
int dfpsr_peer_reset_status(uint32_t *status, void *stream);

#ifdef __cplusplus
}
#endif

#endif
