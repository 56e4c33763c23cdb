// shim_test.cpp — drives the rendering hot path through DFPSR's own C++ API names (dsr_b200.h), the way
// SDK/terrain/main.cpp:383-433 and the reference's test programs (test/tests/*Test.cpp) do. Used by tests/test_gpu_shim.py:
//   shim_test <scene.bin> <out.bin>   renders the scene the Python test wrote, twice (renderer_begin/giveTask/end and model_render),
//                                     checks the API's state machine and writes colour + depth for comparison with the oracle.
#include "dsr_b200.h"

#include <cstdio>
#include <cstdlib>

using namespace dsr;

#define ASSERT(cond) do { if (!(cond)) { std::fprintf(stderr, "ASSERT failed at %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(2); } } while (0)
#define ASSERT_THROWS(stmt, text) do { bool thrown_ = false; try { stmt; } catch (const std::exception &e) { thrown_ = std::string(e.what()).find(text) != std::string::npos; } \
	if (!thrown_) { std::fprintf(stderr, "expected an error containing \"%s\" at %s:%d\n", text, __FILE__, __LINE__); std::exit(2); } } while (0)

struct SceneHeader { int32_t width, height, pointCount, polygonCount, textureWidth, textureHeight, textureLevels, filter, perspective; float location[12]; float widthSlope; };

template <typename T> static std::vector<T> readArray(FILE *f, size_t count) {
	std::vector<T> v(count);
	if (count > 0 && std::fread(v.data(), sizeof(T), count, f) != count) { std::fprintf(stderr, "short scene file\n"); std::exit(2); }
	return v;
}

// shim_test --filters <out.bin>: code written for the reference's filter API (ref: api/filterAPI.h:62-79) compiles unchanged against the
// shim — host lambdas capturing images — and gives the same pixels as the same functions handed to the device as pixel programs.
static int filterSession(const char *outPath) {
	b200_init(0);
	int32_t width = 64;
	int32_t height = 64;
	// the two generators of api/filterAPI.h:67-73, verbatim
	ImageRgbaU8 fadeImage = filter_generateRgbaU8(width, height, [](int32_t x, int32_t y)->ColorRgbaI32 {
		return ColorRgbaI32(x * 4, y * 4, 0, 255);
	});
	ImageRgbaU8 brighterImage = filter_generateRgbaU8(width, height, [fadeImage](int32_t x, int32_t y)->ColorRgbaI32 {
		ColorRgbaI32 source = image_readPixel_clamp(fadeImage, x, y);
		return ColorRgbaI32(source.red * 2, source.green * 2, source.blue * 2, source.alpha);
	});
	// the same two functions as device programs
	ImageRgbaU8 fadeDevice = filter_generateRgbaU8(width, height, PixelProgram("return make_int4(x * 4, y * 4, 0, 255);"));
	ImageRgbaU8 brighterDevice = filter_generateRgbaU8(width, height, PixelProgram("int4 s = read_clamp(0, x, y); return make_int4(s.x * 2, s.y * 2, s.z * 2, s.w);", {fadeDevice}));
	std::vector<uint32_t> a((size_t)width * height), b(a.size()), c(a.size()), d(a.size());
	image_download(fadeImage, a.data(), width * 4); image_download(fadeDevice, b.data(), width * 4);
	image_download(brighterImage, c.data(), width * 4); image_download(brighterDevice, d.data(), width * 4);
	ASSERT(std::memcmp(a.data(), b.data(), a.size() * 4) == 0);
	ASSERT(std::memcmp(c.data(), d.data(), c.size() * 4) == 0);
	ASSERT((c[(size_t)10 * width + 40] & 255u) == 255u && ((c[(size_t)10 * width + 40] >> 8) & 255u) == 80u); // 40 * 4 * 2 saturates, 10 * 4 * 2 = 80
	// in-place map with start offsets, border and tile reads, a target in another pack order; a program that does not compile is an error
	ImageRgbaU8 target = image_create_RgbaU8_native(50, 30, PackOrderIndex::BGRA);
	auto pattern = [fadeImage](int32_t x, int32_t y)->ColorRgbaI32 {
		ColorRgbaI32 t = image_readPixel_tile(fadeImage, x * 3, y - 70), e = image_readPixel_border(fadeImage, x - 5, y, ColorRgbaI32(9, 8, 7, 6));
		return ColorRgbaI32(t.red + e.red / 2, t.green - 300, e.blue + x, e.alpha);
	};
	filter_mapRgbaU8(target, pattern, -7, 11);
	std::vector<uint32_t> hostResult((size_t)50 * 30), deviceResult(hostResult.size());
	image_download(target, hostResult.data(), 50 * 4);
	filter_mapRgbaU8(target, PixelProgram("int4 t = read_tile(0, x * 3, y - 70), e = read_border(0, x - 5, y, make_int4(9, 8, 7, 6)); return make_int4(t.x + e.x / 2, t.y - 300, e.z + x, e.w);", {fadeDevice}), -7, 11);
	image_download(target, deviceResult.data(), 50 * 4);
	ASSERT(std::memcmp(hostResult.data(), deviceResult.data(), hostResult.size() * 4) == 0);
	ASSERT_THROWS(filter_mapRgbaU8(target, PixelProgram("return no_such_thing;")), "does not compile");
	// image_writePixel: saturated, ignored outside, visible to device work and to reads (ref: api/imageAPI.h:184-204)
	image_writePixel(target, 3, 4, ColorRgbaI32(300, -5, 17, 255));
	image_writePixel(target, -1, 4, ColorRgbaI32(1, 2, 3, 4));
	ColorRgbaI32 written = image_readPixel_clamp(target, 3, 4);
	ASSERT(written.red == 255 && written.green == 0 && written.blue == 17 && written.alpha == 255);
	ImageF32 heights = image_create_F32(8, 8);
	image_writePixel(heights, 2, 2, 1.5f);
	ASSERT(image_readPixel_clamp(heights, 2, 2) == 1.5f && image_readPixel_clamp(heights, 3, 2) == 0.0f);
	// both addPointLight signatures of SDK/SpriteEngine/lightAPI.h:29-30 (world centre as IVector2D)
	OrthoSystem ortho(0.6f, 32);
	OrthoView view = ortho.lightView(0);
	ImageRgbaU8 light = image_create_RgbaU8(64, 64), normal = image_create_RgbaU8(64, 64);
	ImageF32 heightBuffer = image_create_F32(64, 64), cube = image_create_F32(16, 96);
	image_fill(normal, ColorRgbaI32(128, 255, 128, 0));
	addPointLight(view, IVector2D(32, 32), light, normal, heightBuffer, FVector3D(0.0f, 2.0f, 0.0f), 4.0f, 1.0f, ColorRgbaI32(255, 200, 100, 0));
	addPointLight(view, IVector2D(32, 32), light, normal, heightBuffer, FVector3D(0.0f, 2.0f, 0.0f), 4.0f, 1.0f, ColorRgbaI32(255, 200, 100, 0), cube);
	// filter_resize(ImageU8) (ref: api/filterAPI.h:42): a horizontal ramp doubled in both directions stays a ramp; nearest repeats pixels
	ImageU8 ramp = image_create_U8(32, 8);
	std::vector<uint8_t> rampRows((size_t)32 * 8);
	for (int32_t y = 0; y < 8; y++) { for (int32_t x = 0; x < 32; x++) { rampRows[(size_t)y * 32 + x] = (uint8_t)(x * 8); } }
	image_upload(ramp, rampRows.data(), 32);
	ImageU8 smooth = filter_resize(ramp, Sampler::Linear, 64, 16), blocky = filter_resize(ramp, Sampler::Nearest, 64, 16);
	ASSERT(image_getWidth(smooth) == 64 && image_getHeight(smooth) == 16);
	std::vector<uint8_t> smoothRows((size_t)64 * 16), blockyRows(smoothRows.size());
	image_download(smooth, smoothRows.data(), 64); image_download(blocky, blockyRows.data(), 64);
	ASSERT(smoothRows[0] == 0 && smoothRows[(size_t)5 * 64 + 21] == 82 && smoothRows[(size_t)15 * 64 + 63] == 248); // (10*8*0.75 + 11*8*0.25) = 82
	ASSERT(blockyRows[(size_t)3 * 64 + 20] == 80 && blockyRows[(size_t)3 * 64 + 21] == 80 && blockyRows[(size_t)3 * 64 + 22] == 88);
	FILE *out = std::fopen(outPath, "wb");
	ASSERT(out != nullptr);
	std::fwrite(c.data(), 4, c.size(), out);
	std::fwrite(hostResult.data(), 4, hostResult.size(), out);
	std::fclose(out);
	std::printf("shim_test filters ok\n");
	return 0;
}

// shim_test --sprites <assets.bin> <out.bin>: a small Sandbox session through spriteWorld_* (ref: SDK/sandbox/sandbox.cpp:336-353, :366-493):
// one sprite type and one model type from the file, a grid of passive sprites, two lights, a temporary sprite, two frames.
struct SpriteHeader { int32_t atlasWidth, atlasHeight, frameRows, centerX, centerY, pointCount, polygonCount, width, height; float minBound[3], maxBound[3]; };
static int spriteSession(const char *inPath, const char *outPath) {
	FILE *f = std::fopen(inPath, "rb");
	ASSERT(f != nullptr);
	SpriteHeader h;
	ASSERT(std::fread(&h, sizeof(h), 1, f) == 1);
	std::vector<uint32_t> atlas = readArray<uint32_t>(f, (size_t)h.atlasWidth * h.atlasHeight);
	std::vector<float> points = readArray<float>(f, (size_t)h.pointCount * 3);
	std::vector<dfpsr_polygon> polygons = readArray<dfpsr_polygon>(f, (size_t)h.polygonCount);
	std::fclose(f);
	b200_init(0);
	SpriteConfig config;
	config.centerX = h.centerX; config.centerY = h.centerY; config.frameRows = h.frameRows; config.propertyColumns = 3;
	config.minBound = FVector3D(h.minBound[0], h.minBound[1], h.minBound[2]); config.maxBound = FVector3D(h.maxBound[0], h.maxBound[1], h.maxBound[2]);
	const int32_t spriteType = spriteWorld_createSpriteType(atlas.data(), h.atlasWidth, h.atlasHeight, h.atlasWidth * 4, config);
	ASSERT(spriteType == spriteWorld_getSpriteTypeCount() - 1);
	Model visible = model_create();
	const int32_t part = model_addEmptyPart(visible, "part");
	for (int32_t i = 0; i < h.pointCount; i++) { model_addPoint(visible, FVector3D(points[3 * i], points[3 * i + 1], points[3 * i + 2])); }
	for (int32_t i = 0; i < h.polygonCount; i++) {
		const dfpsr_polygon &p = polygons[(size_t)i];
		const int32_t index = p.pointIndices[3] < 0 ? model_addTriangle(visible, part, p.pointIndices[0], p.pointIndices[1], p.pointIndices[2]) : model_addQuad(visible, part, p.pointIndices[0], p.pointIndices[1], p.pointIndices[2], p.pointIndices[3]);
		for (int v = 0; v < 4; v++) { model_setVertexColor(visible, part, index, v, FVector4D(p.colors[v][0], p.colors[v][1], p.colors[v][2], p.colors[v][3])); }
	}
	const int32_t modelType = spriteWorld_createModelType(visible, visible);
	SpriteWorld world = spriteWorld_create(OrthoSystem(-0.6f, 64), 64);
	SpriteWorld none;
	ASSERT_THROWS(spriteWorld_clearTemporary(none), "null");
	ASSERT_THROWS(spriteWorld_addBackgroundSprite(world, SpriteInstance(1000000, 0, IVector3D(), false)), "out of bound");
	for (int32_t x = -3; x <= 3; x++) {
		for (int32_t z = -3; z <= 3; z++) { spriteWorld_addBackgroundSprite(world, SpriteInstance(spriteType, (x + z + 16) % 8, IVector3D(x * ortho_miniUnitsPerTile, 0, z * ortho_miniUnitsPerTile), true)); }
	}
	spriteWorld_addBackgroundModel(world, ModelInstance(modelType, Transform3D(FVector3D(0.5f, 0.0f, -0.5f), FMatrix3x3())));
	ImageRgbaU8 color = image_create_RgbaU8(h.width, h.height);
	std::vector<uint32_t> frames((size_t)h.width * h.height * 2);
	for (int frame = 0; frame < 2; frame++) {
		spriteWorld_clearTemporary(world);
		spriteWorld_createTemporary_directedLight(world, FVector3D(1.0f, -1.0f, 0.0f), 0.1f, ColorRgbaI32(255, 255, 255, 255));
		spriteWorld_createTemporary_pointLight(world, FVector3D(0.5f, 1.5f, 0.5f), 4.0f, 1.0f, ColorRgbaI32(255, 200, 150, 255), true);
		spriteWorld_addTemporarySprite(world, SpriteInstance(spriteType, frame, IVector3D(300 + 200 * frame, 256, -100), true));
		if (frame == 1) { spriteWorld_moveCameraInPixels(world, IVector2D(12, -7)); }
		spriteWorld_draw(world, color);
		image_download(color, frames.data() + (size_t)frame * h.width * h.height, h.width * 4);
	}
	ASSERT(image_getWidth(spriteWorld_getHeightBuffer(world)) == h.width && image_exists(spriteWorld_getDiffuseBuffer(world)));
	const IVector3D ground = spriteWorld_findGroundAtPixel(world, color, IVector2D(10, 20)), camera = spriteWorld_getCameraLocation(world);
	FILE *out = std::fopen(outPath, "wb");
	ASSERT(out != nullptr);
	const int32_t tail[8] = {ground.x, ground.y, ground.z, camera.x, camera.y, camera.z, spriteType, modelType};
	std::fwrite(frames.data(), 4, frames.size(), out);
	std::fwrite(tail, 4, 8, out);
	std::fclose(out);
	std::printf("shim_test sprites ok: %dx%d\n", h.width, h.height);
	return 0;
}

int main(int argc, char **argv) {
	if (argc == 3 && std::string(argv[1]) == "--filters") { return filterSession(argv[2]); }
	if (argc == 4 && std::string(argv[1]) == "--sprites") { return spriteSession(argv[2], argv[3]); }
	if (argc < 3) { std::fprintf(stderr, "usage: shim_test scene.bin out.bin\n"); return 2; }
	FILE *f = std::fopen(argv[1], "rb");
	ASSERT(f != nullptr);
	SceneHeader h;
	ASSERT(std::fread(&h, sizeof(h), 1, f) == 1);
	std::vector<float> points = readArray<float>(f, (size_t)h.pointCount * 3);
	std::vector<dfpsr_polygon> polygons = readArray<dfpsr_polygon>(f, (size_t)h.polygonCount);
	std::vector<uint32_t> texels = readArray<uint32_t>(f, (size_t)h.textureWidth * h.textureHeight);
	std::fclose(f);

	b200_init(0);

	// texture_create_RgbaU8 + upload of level 0 + texture_generatePyramid (ref: SDK/terrain/main.cpp:370, api/textureAPI.cpp:65-87)
	TextureRgbaU8 diffuse;
	if (h.textureWidth > 0) {
		diffuse = texture_create_RgbaU8(h.textureWidth, h.textureHeight, h.textureLevels);
		ImageRgbaU8 level0 = texture_getMipLevelImage(diffuse, 0);
		ASSERT(image_getWidth(level0) == h.textureWidth && image_getHeight(level0) == h.textureHeight);
		image_upload(level0, texels.data(), h.textureWidth * 4);
		texture_generatePyramid(diffuse);
		ASSERT(texture_getMaxWidth(diffuse) == h.textureWidth && texture_getSmallestMipLevel(diffuse) <= h.textureLevels - 1);
	}

	// the model, through the model_* calls (ref: api/modelAPI.h:62-258)
	Model model = model_create();
	ASSERT(model_exists(model));
	model_setFilter(model, h.filter ? Filter::Alpha : Filter::Solid);
	int32_t part = model_addEmptyPart(model, "part");
	ASSERT(part == 0 && model_getNumberOfParts(model) == 1);
	for (int32_t i = 0; i < h.pointCount; i++) { ASSERT(model_addPoint(model, FVector3D(points[3 * i], points[3 * i + 1], points[3 * i + 2])) == i); }
	for (int32_t i = 0; i < h.polygonCount; i++) {
		const dfpsr_polygon &p = polygons[(size_t)i];
		int32_t index = p.pointIndices[3] < 0 ? model_addTriangle(model, part, p.pointIndices[0], p.pointIndices[1], p.pointIndices[2])
		                                      : model_addQuad(model, part, p.pointIndices[0], p.pointIndices[1], p.pointIndices[2], p.pointIndices[3]);
		ASSERT(index == i);
		for (int v = 0; v < 4; v++) {
			model_setVertexColor(model, part, index, v, FVector4D(p.colors[v][0], p.colors[v][1], p.colors[v][2], p.colors[v][3]));
			model_setTexCoord(model, part, index, v, FVector4D(p.texCoords[v][0], p.texCoords[v][1], p.texCoords[v][2], p.texCoords[v][3]));
		}
	}
	ASSERT(model_getNumberOfPolygons(model, part) == h.polygonCount && model_getNumberOfPoints(model) == h.pointCount);
	model_setDiffuseMap(model, part, diffuse);
	ASSERT(model_addTriangle(model, 7, 0, 1, 2) == -1); // out-of-range part: reported, nothing added (ref: Model.cpp:34-42)

	Transform3D location(FVector3D(h.location[0], h.location[1], h.location[2]),
	                     FMatrix3x3(FVector3D(h.location[3], h.location[4], h.location[5]), FVector3D(h.location[6], h.location[7], h.location[8]), FVector3D(h.location[9], h.location[10], h.location[11])));
	Camera camera = h.perspective ? Camera::createPerspective(location, (float)h.width, (float)h.height, h.widthSlope) : Camera::createOrthogonal(location, (float)h.width, (float)h.height, h.widthSlope);

	// ---- one frame as SDK/terrain/main.cpp:397-421 does it
	ImageRgbaU8 colorBuffer = image_create_RgbaU8(h.width, h.height);
	ImageF32 depthBuffer = image_create_F32(h.width, h.height);
	Renderer worker = renderer_create();
	ASSERT(renderer_exists(worker) && !renderer_takesTriangles(worker));
	ASSERT_THROWS(renderer_end(worker), "without renderer_begin");
	image_fill(colorBuffer, ColorRgbaI32(0, 0, 0, 0));
	image_fill(depthBuffer, h.perspective ? 0.0f : 1.0e9f);
	renderer_begin(worker, colorBuffer, depthBuffer);
	ASSERT(renderer_takesTriangles(worker));
	ASSERT_THROWS(renderer_begin(worker, colorBuffer, depthBuffer), "twice");
	renderer_giveTask(worker, model, Transform3D(), camera);
	renderer_end(worker);
	ASSERT(!renderer_takesTriangles(worker) && !image_exists(renderer_getColorBuffer(worker)));

	// ---- the same frame through model_render into a second pair of images: identical pixels
	ImageRgbaU8 color2 = image_create_RgbaU8(h.width, h.height);
	ImageF32 depth2 = image_create_F32(h.width, h.height);
	image_fill(depth2, h.perspective ? 0.0f : 1.0e9f);
	model_render(model, Transform3D(), color2, depth2, camera);
	std::vector<uint32_t> c1((size_t)h.width * h.height), c2(c1.size());
	std::vector<float> d1(c1.size()), d2(c1.size());
	image_download(colorBuffer, c1.data(), h.width * 4); image_download(color2, c2.data(), h.width * 4);
	image_download(depthBuffer, d1.data(), h.width * 4); image_download(depth2, d2.data(), h.width * 4);
	ASSERT(std::memcmp(c1.data(), c2.data(), c1.size() * 4) == 0);
	ASSERT(std::memcmp(d1.data(), d2.data(), d1.size() * 4) == 0);

	// ---- pixel reads and sub-images behave like the reference's handles (ref: api/imageAPI.h:207-341, Image.h:186-206)
	int32_t mx = h.width / 2, my = h.height / 2;
	ColorRgbaI32 centre = image_readPixel_clamp(colorBuffer, mx, my);
	uint32_t packed = c1[(size_t)my * h.width + mx];
	ASSERT(centre.red == (int32_t)(packed & 255u) && centre.alpha == (int32_t)(packed >> 24));
	ASSERT(image_readPixel_clamp(depthBuffer, -5, my) == d1[(size_t)my * h.width]);
	ImageRgbaU8 sub = image_getSubImage(colorBuffer, mx / 2, my / 2, mx, my);
	ASSERT(image_isSubImage(sub) && image_getWidth(sub) == mx && image_getStride(sub) == image_getStride(colorBuffer));
	ColorRgbaI32 viaSub = image_readPixel_clamp(sub, mx - mx / 2, my - my / 2);
	ASSERT(viaSub.red == centre.red && viaSub.green == centre.green && viaSub.blue == centre.blue);
	ASSERT(!image_exists(image_getSubImage(colorBuffer, 1, 1, h.width, h.height)));
	// model_renderDepth into a depth-only target equals the depth of a colourless render where nothing is alpha filtered
	ImageF32 depth3 = image_create_F32(h.width, h.height);
	image_fill(depth3, h.perspective ? 0.0f : 1.0e9f);
	ImageRgbaU8 none;
	model_render(model, Transform3D(), none, depth3, camera);
	std::vector<float> d3(c1.size());
	image_download(depth3, d3.data(), h.width * 4);
	if (!h.filter) { ASSERT(std::memcmp(d1.data(), d3.data(), d1.size() * 4) == 0); }

	// renderer_end(renderer, debugWireframe = true) (ref: api/rendererAPI.h:134-135): the same depth, white edges on top of the colours
	ImageRgbaU8 color4 = image_create_RgbaU8(h.width, h.height);
	ImageF32 depth4 = image_create_F32(h.width, h.height);
	image_fill(depth4, h.perspective ? 0.0f : 1.0e9f);
	renderer_begin(worker, color4, depth4);
	renderer_giveTask(worker, model, Transform3D(), camera);
	renderer_end(worker, true);
	std::vector<uint32_t> c4(c1.size());
	std::vector<float> d4(c1.size());
	image_download(color4, c4.data(), h.width * 4); image_download(depth4, d4.data(), h.width * 4);
	ASSERT(std::memcmp(d1.data(), d4.data(), d1.size() * 4) == 0);
	size_t changed = 0, changedToWhite = 0;
	for (size_t i = 0; i < c1.size(); i++) { if (c4[i] != c1[i]) { changed++; if (c4[i] == 0xFFFFFFFFu) { changedToWhite++; } } }
	ASSERT(changed > 0 && changed == changedToWhite);

	FILE *out = std::fopen(argv[2], "wb");
	ASSERT(out != nullptr);
	std::fwrite(c1.data(), 4, c1.size(), out);
	std::fwrite(d1.data(), 4, d1.size(), out);
	std::fclose(out);
	std::printf("shim_test ok: %d polygons, %dx%d\n", h.polygonCount, h.width, h.height);
	return 0;
}
