// session.cu — host-buffer entry points: the reference API works on host images, so a drop-in call has to
// move the frame's inputs to the device and its result back. A session keeps device mirrors of models and of
// the render targets so that a frame only moves what changed (geometry on request, the finished colour/depth out).
#include "common.cuh"

#include <stdlib.h>
#include <vector>
#include <new>

using namespace dfpsr;

namespace {

struct ModelSlot {
	DeviceBuffer points, polygons, diffuse, light;
	dfpsr_host_model host{};
	dfpsr_model device{};
};

} // namespace

static const int PIPELINE_CHUNK = 16; // views rendered per submission while the previous chunk is copied to the host

struct dfpsr_session {
	std::vector<ModelSlot *> models;
	dfpsr_renderer *renderer = nullptr;
	DeviceBuffer color, depth;
	// double-buffered targets of dfpsr_session_render_views_host
	DeviceBuffer chunkColor[2], chunkDepth[2];
	cudaStream_t copyStream = nullptr;
	cudaEvent_t rendered[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
	~dfpsr_session() {
		for (int b = 0; b < 2; b++) {
			chunkColor[b].release(); chunkDepth[b].release();
			if (rendered[b]) { cudaEventDestroy(rendered[b]); }
			if (copied[b]) { cudaEventDestroy(copied[b]); }
		}
		if (copyStream) { cudaStreamDestroy(copyStream); }
		for (ModelSlot *m : models) {
			m->points.release(); m->polygons.release(); m->diffuse.release(); m->light.release();
			delete m;
		}
		if (renderer) { dfpsr_renderer_destroy(renderer); }
		color.release(); depth.release();
	}
};

static int upload_texture(DeviceBuffer &buffer, dfpsr_texture &device, const uint32_t *pixels, const dfpsr_texture &layout) {
	device = layout;
	device.data = nullptr;
	if (pixels == nullptr) { return 0; }
	if (buffer.reserve((size_t)layout.totalPixels * 4)) { return 1; }
	DFPSR_CHECK_CUDA(cudaMemcpy(buffer.ptr, pixels, (size_t)layout.totalPixels * 4, cudaMemcpyHostToDevice));
	device.data = (const uint32_t *)buffer.ptr;
	return 0;
}

extern "C" {

int dfpsr_session_create(dfpsr_session **out) {
	DFPSR_REQUIRE(out != nullptr, "session_create: null output");
	*out = new (std::nothrow) dfpsr_session();
	DFPSR_REQUIRE(*out != nullptr, "out of host memory");
	if (dfpsr_renderer_create(&(*out)->renderer)) { delete *out; *out = nullptr; return 1; }
	return 0;
}

int dfpsr_session_destroy(dfpsr_session *session) {
	delete session;
	return 0;
}

int dfpsr_session_upload_model(dfpsr_session *session, const dfpsr_host_model *model, int32_t *slot) {
	DFPSR_REQUIRE(session != nullptr && model != nullptr && slot != nullptr, "session_upload_model: null argument");
	DFPSR_REQUIRE(model->pointCount >= 0 && model->polygonCount >= 0, "session_upload_model: negative counts");
	ModelSlot *m = new (std::nothrow) ModelSlot();
	DFPSR_REQUIRE(m != nullptr, "out of host memory");
	m->host = *model;
	if (m->points.reserve((size_t)model->pointCount * 12 + 16) || m->polygons.reserve((size_t)model->polygonCount * sizeof(dfpsr_polygon) + 16)) { delete m; return 1; }
	DFPSR_CHECK_CUDA(cudaMemcpy(m->points.ptr, model->points, (size_t)model->pointCount * 12, cudaMemcpyHostToDevice));
	DFPSR_CHECK_CUDA(cudaMemcpy(m->polygons.ptr, model->polygons, (size_t)model->polygonCount * sizeof(dfpsr_polygon), cudaMemcpyHostToDevice));
	m->device.points = (const float *)m->points.ptr;
	m->device.pointCount = model->pointCount;
	m->device.polygons = (const dfpsr_polygon *)m->polygons.ptr;
	m->device.polygonCount = model->polygonCount;
	m->device.filter = model->filter;
	if (upload_texture(m->diffuse, m->device.diffuse, model->diffusePixels, model->diffuseLayout)) { delete m; return 1; }
	if (upload_texture(m->light, m->device.light, model->lightPixels, model->lightLayout)) { delete m; return 1; }
	for (int k = 0; k < 3; k++) { m->device.minBound[k] = model->minBound[k]; m->device.maxBound[k] = model->maxBound[k]; }
	session->models.push_back(m);
	*slot = (int32_t)session->models.size() - 1;
	return 0;
}

int dfpsr_session_render_frame_host(dfpsr_session *session, int32_t slot, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *camera, uint32_t *colorHost, int32_t colorStride, float *depthHost, int32_t depthStride, int32_t width, int32_t height, int32_t packOrder, int32_t uploadGeometry, void *stream) {
	DFPSR_REQUIRE(session != nullptr && modelToWorld != nullptr && camera != nullptr, "session_render_frame_host: null argument");
	DFPSR_REQUIRE(slot >= 0 && slot < (int32_t)session->models.size(), "session_render_frame_host: model slot %d does not exist", slot);
	DFPSR_REQUIRE(width > 0 && height > 0, "session_render_frame_host: empty target");
	cudaStream_t s = as_stream(stream);
	ModelSlot *m = session->models[slot];
	if (uploadGeometry) {
		DFPSR_CHECK_CUDA(cudaMemcpyAsync(m->points.ptr, m->host.points, (size_t)m->host.pointCount * 12, cudaMemcpyHostToDevice, s));
		DFPSR_CHECK_CUDA(cudaMemcpyAsync(m->polygons.ptr, m->host.polygons, (size_t)m->host.polygonCount * sizeof(dfpsr_polygon), cudaMemcpyHostToDevice, s));
	}
	int32_t pitch = ((width * 4 + 255) / 256) * 256; // rows start on 256-byte boundaries: every 16-byte access is aligned
	if (session->color.reserve((size_t)pitch * height) || session->depth.reserve((size_t)pitch * height)) { return 1; }
	dfpsr_image color{session->color.ptr, width, height, pitch, packOrder};
	dfpsr_image depth{session->depth.ptr, width, height, pitch, 0};
	// image_fill(colour, 0) + image_fill(depth, 0) + renderer_begin (ref: SDK/terrain/main.cpp:397-416), fused into the tile kernel
	if (dfpsr_renderer_begin_cleared(session->renderer, &color, &depth, 0u, 0.0f)) { return 1; }
	if (dfpsr_renderer_give_task(session->renderer, &m->device, modelToWorld, camera, stream)) { return 1; }
	if (dfpsr_renderer_end(session->renderer, stream)) { return 1; }
	if (verify_pending_frames()) { return 1; } // an asynchronous renderer: the copies below must see the frame's final pixels
	if (colorHost != nullptr) { DFPSR_CHECK_CUDA(cudaMemcpy2DAsync(colorHost, (size_t)colorStride, color.data, (size_t)pitch, (size_t)width * 4, (size_t)height, cudaMemcpyDeviceToHost, s)); }
	if (depthHost != nullptr) { DFPSR_CHECK_CUDA(cudaMemcpy2DAsync(depthHost, (size_t)depthStride, depth.data, (size_t)pitch, (size_t)width * 4, (size_t)height, cudaMemcpyDeviceToHost, s)); }
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(s));
	return 0;
}

// Many views of one model with HOST targets (BASELINE config 4 end to end): views are rendered PIPELINE_CHUNK at a time into one of two
// device buffer sets while the previous chunk travels to the host on a second stream, so the PCIe copy of frame i overlaps the
// rasterisation of frame i + 1. Host images should be pinned (dfpsr_malloc_host) for the copies to be asynchronous.
int dfpsr_session_render_views_host(dfpsr_session *session, int32_t slot, const dfpsr_transform3d *modelToWorld, const dfpsr_camera *cameras, int32_t count, uint32_t *const *colorHost, int32_t colorStride, float *const *depthHost, int32_t depthStride, int32_t width, int32_t height, int32_t packOrder, int32_t uploadGeometry, void *stream) {
	DFPSR_REQUIRE(session != nullptr && modelToWorld != nullptr && cameras != nullptr, "session_render_views_host: null argument");
	DFPSR_REQUIRE(slot >= 0 && slot < (int32_t)session->models.size(), "session_render_views_host: model slot %d does not exist", slot);
	DFPSR_REQUIRE(width > 0 && height > 0, "session_render_views_host: empty target");
	if (count <= 0) { return 0; }
	cudaStream_t s = as_stream(stream);
	ModelSlot *m = session->models[slot];
	if (!session->copyStream) {
		DFPSR_CHECK_CUDA(cudaStreamCreateWithFlags(&session->copyStream, cudaStreamNonBlocking));
		for (int b = 0; b < 2; b++) {
			DFPSR_CHECK_CUDA(cudaEventCreateWithFlags(&session->rendered[b], cudaEventDisableTiming));
			DFPSR_CHECK_CUDA(cudaEventCreateWithFlags(&session->copied[b], cudaEventDisableTiming));
		}
	}
	if (uploadGeometry) {
		DFPSR_CHECK_CUDA(cudaMemcpyAsync(m->points.ptr, m->host.points, (size_t)m->host.pointCount * 12, cudaMemcpyHostToDevice, s));
		DFPSR_CHECK_CUDA(cudaMemcpyAsync(m->polygons.ptr, m->host.polygons, (size_t)m->host.polygonCount * sizeof(dfpsr_polygon), cudaMemcpyHostToDevice, s));
	}
	const int32_t pitch = ((width * 4 + 255) / 256) * 256;
	const size_t frameBytes = (size_t)pitch * (size_t)height;
	int32_t chunkLimit = PIPELINE_CHUNK;
	if (const char *env = getenv("DFPSR_PIPELINE_CHUNK")) { int v = atoi(env); if (v > 0 && v <= 1024) { chunkLimit = v; } } // tuning knob
	const int32_t chunk = count < chunkLimit ? count : chunkLimit;
	for (int b = 0; b < 2; b++) {
		if (session->chunkColor[b].reserve(frameBytes * chunk) || session->chunkDepth[b].reserve(frameBytes * chunk)) { return 1; }
	}
	std::vector<dfpsr_image> colors((size_t)chunk), depths((size_t)chunk);
	int32_t chunkIndex = 0;
	for (int32_t first = 0; first < count; first += chunk, chunkIndex++) {
		const int b = chunkIndex & 1;
		const int32_t n = count - first < chunk ? count - first : chunk;
		if (chunkIndex >= 2) { DFPSR_CHECK_CUDA(cudaStreamWaitEvent(s, session->copied[b], 0)); } // the buffer set is free again
		for (int32_t v = 0; v < n; v++) {
			colors[(size_t)v] = dfpsr_image{(uint8_t *)session->chunkColor[b].ptr + frameBytes * v, width, height, pitch, packOrder};
			depths[(size_t)v] = dfpsr_image{(uint8_t *)session->chunkDepth[b].ptr + frameBytes * v, width, height, pitch, 0};
		}
		// on an error the copies of earlier chunks are still in flight towards the caller's buffers: wait for them before returning
		if (dfpsr_model_render_views(&m->device, modelToWorld, colors.data(), depths.data(), cameras + first, n, 1, stream) || verify_pending_frames()) {
			cudaStreamSynchronize(session->copyStream); cudaStreamSynchronize(s);
			return 1;
		}
		DFPSR_CHECK_CUDA(cudaEventRecord(session->rendered[b], s));
		DFPSR_CHECK_CUDA(cudaStreamWaitEvent(session->copyStream, session->rendered[b], 0));
		for (int32_t v = 0; v < n; v++) {
			if (colorHost != nullptr && colorHost[first + v] != nullptr) {
				if (colorStride == pitch && pitch == width * 4) { DFPSR_CHECK_CUDA(cudaMemcpyAsync(colorHost[first + v], colors[(size_t)v].data, frameBytes, cudaMemcpyDeviceToHost, session->copyStream)); }
				else { DFPSR_CHECK_CUDA(cudaMemcpy2DAsync(colorHost[first + v], (size_t)colorStride, colors[(size_t)v].data, (size_t)pitch, (size_t)width * 4, (size_t)height, cudaMemcpyDeviceToHost, session->copyStream)); }
			}
			if (depthHost != nullptr && depthHost[first + v] != nullptr) {
				if (depthStride == pitch && pitch == width * 4) { DFPSR_CHECK_CUDA(cudaMemcpyAsync(depthHost[first + v], depths[(size_t)v].data, frameBytes, cudaMemcpyDeviceToHost, session->copyStream)); }
				else { DFPSR_CHECK_CUDA(cudaMemcpy2DAsync(depthHost[first + v], (size_t)depthStride, depths[(size_t)v].data, (size_t)pitch, (size_t)width * 4, (size_t)height, cudaMemcpyDeviceToHost, session->copyStream)); }
			}
		}
		DFPSR_CHECK_CUDA(cudaEventRecord(session->copied[b], session->copyStream));
	}
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(session->copyStream));
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(s));
	return 0;
}

// ref: implementation/gui/DsrWindow.cpp:255-281 DsrWindow::showCanvas — the low-resolution canvas is magnified by whole pixels into the
// back-end's canvas (its native pack order, exact pixel size: partial pixels at the right / bottom are cut, anything beyond the source is
// transparent black) and handed to the window system. The back-end's canvas is host memory, so the hand-off is: block magnify on the device
// into a staging image of the canvas's size and pack order, then ONE device-to-host copy into the caller's canvas rows.
int dfpsr_canvas_show(const dfpsr_image *deviceCanvas, int32_t pixelScale, void *hostCanvas, int32_t hostStrideBytes, int32_t hostWidth, int32_t hostHeight, int32_t hostPackOrder, void *stream) {
	DFPSR_REQUIRE(deviceCanvas != nullptr && deviceCanvas->data != nullptr && hostCanvas != nullptr, "canvas_show: null argument");
	DFPSR_REQUIRE(pixelScale >= 1 && hostWidth > 0 && hostHeight > 0 && hostStrideBytes >= hostWidth * 4, "canvas_show: pixel scale %d, canvas %d x %d, stride %d", pixelScale, hostWidth, hostHeight, hostStrideBytes);
	cudaStream_t s = as_stream(stream);
	const uint8_t *source = (const uint8_t *)deviceCanvas->data;
	size_t sourcePitch = (size_t)deviceCanvas->stride;
	static thread_local DeviceBuffer staging;
	if (pixelScale > 1 || hostPackOrder != deviceCanvas->packOrder || deviceCanvas->width != hostWidth || deviceCanvas->height != hostHeight) {
		const int32_t pitch = ((hostWidth * 4 + 255) / 256) * 256;
		if (staging.reserve((size_t)pitch * (size_t)hostHeight)) { return 1; }
		const dfpsr_image target{staging.ptr, hostWidth, hostHeight, pitch, hostPackOrder};
		if (dfpsr_filter_block_magnify(&target, deviceCanvas, pixelScale, pixelScale, stream)) { return 1; }
		source = (const uint8_t *)staging.ptr; sourcePitch = (size_t)pitch;
	} else if (verify_pending_frames()) { return 1; } // the canvas may be a frame in flight: the copy below is not queued through DFPSR_LAUNCH
	DFPSR_CHECK_CUDA(cudaMemcpy2DAsync(hostCanvas, (size_t)hostStrideBytes, source, sourcePitch, (size_t)hostWidth * 4, (size_t)hostHeight, cudaMemcpyDeviceToHost, s));
	DFPSR_CHECK_CUDA(cudaStreamSynchronize(s)); // the window system reads the canvas next
	return 0;
}

} // extern "C"
