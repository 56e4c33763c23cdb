/*
 * dfpsr_oracle.h — TEST INFRASTRUCTURE. CPU restatement (plain C) of the reference's algorithm for the
 * rendering hot path, used only as the checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg. The product (dfpsr_b200/) never links or loads it.
 *
 * Parity status: PINNED. Every function here is checked bit-for-bit against the compiled, unmodified
 * reference (oracle/_ref/libdfpsr_ref_scalar.so, built by oracle/Makefile) by tests/test_oracle_vs_ref.py,
 * and against the golden hashes committed in tests/golden/ (generated from the same reference build by
 * tests/golden/make_golden.py). The reference's own unit tests do not cover this path.
 *
 * All images here are HOST memory described with dfpsr_image (data = host pointer).
 */
#ifndef DFPSR_ORACLE_H
#define DFPSR_ORACLE_H

#include "../include/dfpsr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Camera::createPerspective / createOrthogonal from the constructor arguments already stored in *camera
 * (perspective, location, imageWidth, imageHeight, widthSlope [= halfWidth when orthogonal], nearClip, farClip). */
void orc_camera_create(dfpsr_camera *camera);
int orc_camera_is_box_seen(const dfpsr_camera *camera, const float *minBound, const float *maxBound, const dfpsr_transform3d *modelToWorld);
void orc_project_points(const float *points, int32_t count, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, dfpsr_projected_point *out);

void orc_texture_layout(dfpsr_texture *out, int32_t width, int32_t height, int32_t resolutions);
void orc_texture_generate_pyramid(uint32_t *pixels, const dfpsr_texture *layout);
/* test hooks for the reference's own known-answer tests (test/tests/TextureTest.cpp) and for the coverage property test */
uint32_t orc_texture_layer_offset(const dfpsr_texture *t, uint32_t mip);
uint32_t orc_texture_pixel_offset(const dfpsr_texture *t, uint32_t x, uint32_t y, uint32_t mip);
uint32_t orc_interpolate_color_linear(uint32_t colorA, uint32_t colorB, uint32_t weight);
uint32_t orc_texture_sample_bilinear(const dfpsr_texture *t, float u, float v, uint32_t mip);
void orc_rasterize_rows(const int64_t *fx, const int64_t *fy, int32_t l, int32_t t, int32_t w, int32_t h, int32_t *rowsOut);
int orc_is_frontfacing(const int64_t *fx, const int64_t *fy);

/* model_render / renderer_begin+giveTask+end (identical pixels). color and/or depth may have data == NULL.
 * Returns the number of draw commands (triangles after culling/clipping/back-face removal). */
int64_t orc_model_render(const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_image *color, const dfpsr_image *depth, const dfpsr_camera *camera);
/* renderer_giveTask_triangle for pre-projected triangles. */
int64_t orc_render_triangles(const dfpsr_triangle *triangles, int32_t count, const dfpsr_texture *diffuse, const dfpsr_texture *light, int32_t filter, const dfpsr_image *color, const dfpsr_image *depth, const dfpsr_camera *camera);
/* Renderer object with the deferred queue and the 16-pixel-cell occlusion grid (ref: api/rendererAPI.cpp:141-477). */
typedef struct orc_renderer orc_renderer;
orc_renderer *orc_renderer_create(void);
void orc_renderer_destroy(orc_renderer *r);
void orc_renderer_begin(orc_renderer *r, const dfpsr_image *color, const dfpsr_image *depth);
int orc_renderer_has_occluders(const orc_renderer *r);
void orc_renderer_occlude_from_box(orc_renderer *r, const float *minBound, const float *maxBound, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera);
void orc_renderer_occlude_from_existing_triangles(orc_renderer *r);
void orc_renderer_occlude_from_top_rows(orc_renderer *r, const dfpsr_camera *camera);
int orc_renderer_is_box_visible(const orc_renderer *r, const float *minBound, const float *maxBound, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera);
void orc_renderer_give_task(orc_renderer *r, const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera);
/* renderer_end(renderer, debugWireframe = true) (api/rendererAPI.cpp:362-399) for the next orc_renderer_end. */
void orc_renderer_set_debug_wireframe(orc_renderer *r, int enabled);
int64_t orc_renderer_end(orc_renderer *r, int64_t *occludedOut);
/* model_renderDepth */
void orc_model_render_depth(const dfpsr_model *model, const dfpsr_transform3d *modelToWorld, const dfpsr_image *depth, const dfpsr_camera *camera);

void orc_image_fill_rgba(const dfpsr_image *image, int32_t r, int32_t g, int32_t b, int32_t a);
void orc_image_fill_f32(const dfpsr_image *image, float value);
void orc_draw_copy_rgba(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top);
void orc_draw_copy_f32(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top);
void orc_draw_higher(const dfpsr_image *targetHeight, const dfpsr_image *sourceHeight, const dfpsr_image *targetA, const dfpsr_image *sourceA, const dfpsr_image *targetB, const dfpsr_image *sourceB, int32_t left, int32_t top, float offset);

/* The remaining draw calls on RGBA8 / F32 images (api/drawAPI.cpp:72-310, :636-757). */
void orc_draw_rectangle_rgba(const dfpsr_image *image, int32_t left, int32_t top, int32_t width, int32_t height, const int32_t *colorRgba);
void orc_draw_rectangle_f32(const dfpsr_image *image, int32_t left, int32_t top, int32_t width, int32_t height, float value);
void orc_draw_line_rgba(const dfpsr_image *image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, const int32_t *colorRgba);
void orc_draw_line_f32(const dfpsr_image *image, int32_t x1, int32_t y1, int32_t x2, int32_t y2, float value);
void orc_draw_alpha_filter(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top);
void orc_draw_max_alpha(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, int32_t sourceAlphaOffset);
void orc_draw_alpha_clip(const dfpsr_image *target, const dfpsr_image *source, int32_t left, int32_t top, int32_t threshold);
void orc_draw_silhouette(const dfpsr_image *target, const dfpsr_image *silhouetteU8, const int32_t *colorRgba, int32_t left, int32_t top);

/* 8-bit / 16-bit monochrome images, the mixed-format draw_copy overloads and draw_higher on 16-bit heights (api/drawAPI.cpp:130-150, :284-297, :519-634, :759-832). */
void orc_draw_rectangle_mono(const dfpsr_image *image, int32_t format, int32_t left, int32_t top, int32_t width, int32_t height, int32_t color);
void orc_draw_line_mono(const dfpsr_image *image, int32_t format, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int32_t color);
void orc_draw_copy_formats(const dfpsr_image *target, int32_t targetFormat, const dfpsr_image *source, int32_t sourceFormat, int32_t left, int32_t top);
void orc_draw_higher_u16(const dfpsr_image *targetHeight, const dfpsr_image *sourceHeight, const dfpsr_image *targetA, const dfpsr_image *sourceA, const dfpsr_image *targetB, const dfpsr_image *sourceB, int32_t left, int32_t top, int32_t offset);

/* renderDenseModel<HIGH_QUALITY> (SDK/SpriteEngine/spriteAPI.cpp:1243-1327); dirtyRect = {left, top, width, height} or zeros when culled. */
void orc_dense_model_render(const dfpsr_dense_triangle *triangles, int32_t triangleCount, const float *minBound, const float *maxBound, const dfpsr_ortho_camera *view,
                            const dfpsr_image *height, const dfpsr_image *diffuse, const dfpsr_image *normal, const float *worldOrigin, const dfpsr_transform3d *modelToWorld,
                            int32_t highQuality, int32_t *dirtyRect);
/* scaleHeightImage (SDK/SpriteEngine/spriteAPI.cpp:157-174) for one sprite frame. */
void orc_sprite_scale_height(const dfpsr_image *heightColumn, const dfpsr_image *colorColumn, float minHeight, float maxHeight, const dfpsr_image *out);

/* laneCount: the reference's laneCountX_32Bit (4 for the SSE2/scalar builds, 8 for AVX2); it decides the
 * alignment of the point light's rectangle and the grouping of its incremental position adds.
 * rowsPerJob: 0 = one job for the whole rectangle (DISABLE_MULTI_THREADING build). */
void orc_light_directed(const dfpsr_ortho_view *view, const dfpsr_image *light, const dfpsr_image *normal, const float *direction, float intensity, const int32_t *colorRgb, int32_t add);
void orc_light_point(const dfpsr_ortho_view *view, const int32_t *worldCenter, const dfpsr_image *light, const dfpsr_image *normal, const dfpsr_image *height, const float *position, float radius, float intensity, const int32_t *colorRgb, const dfpsr_image *shadowCubeMap, int32_t laneCount);
void orc_light_blend(const dfpsr_image *color, const dfpsr_image *diffuse, const dfpsr_image *light);

/* filter_resize into an existing RGBA-order target; scratch must hold target.width * source.height u32. */
void orc_filter_resize(const dfpsr_image *target, const dfpsr_image *source, int32_t sampler, int32_t sourceIsSubImage, uint32_t *scratch);
/* filter_resize(ImageU8): one byte per pixel, stride in bytes; scratch must hold target.width * source.height bytes. */
void orc_filter_resize_u8(const dfpsr_image *target, const dfpsr_image *source, int32_t sampler, uint8_t *scratch);
void orc_filter_map(const dfpsr_image *target, int32_t op, const int32_t *params, const dfpsr_image *source, int32_t startX, int32_t startY);
void orc_filter_block_magnify(const dfpsr_image *target, const dfpsr_image *source, int32_t pixelWidth, int32_t pixelHeight);

#ifdef __cplusplus
}
#endif
#endif
