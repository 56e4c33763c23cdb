// model_test.cpp — host-side check of the dsr:: shim's model construction and importers (no GPU needed): prints the geometry a PLY / DMF1
// text turns into and the defaults of model_addTriangle / model_addQuad, for tests/test_importers.py to compare with the C ABI's own output.
// usage: model_test ply|dmf <file> [flipX | detailLevel]
#include "dsr_b200.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

using namespace dsr;

static void print_model(const Model &model) {
	std::printf("points %d parts %d\n", model_getNumberOfPoints(model), model_getNumberOfParts(model));
	for (int32_t i = 0; i < model_getNumberOfPoints(model); i++) {
		const FVector3D p = model_getPoint(model, i);
		std::printf("p %a %a %a\n", p.x, p.y, p.z);
	}
	for (int32_t part = 0; part < model_getNumberOfParts(model); part++) {
		std::printf("part %d polygons %d\n", part, model_getNumberOfPolygons(model, part));
		for (const dfpsr_polygon &polygon : model->parts[(size_t)part].polygons) {
			std::printf("i %d %d %d %d\n", polygon.pointIndices[0], polygon.pointIndices[1], polygon.pointIndices[2], polygon.pointIndices[3]);
			for (int v = 0; v < 4; v++) {
				std::printf("v %a %a %a %a %a %a %a %a\n", polygon.texCoords[v][0], polygon.texCoords[v][1], polygon.texCoords[v][2], polygon.texCoords[v][3],
				            polygon.colors[v][0], polygon.colors[v][1], polygon.colors[v][2], polygon.colors[v][3]);
			}
		}
	}
	FVector3D mn, mx;
	model_getBoundingBox(model, mn, mx);
	std::printf("bound %a %a %a %a %a %a\n", mn.x, mn.y, mn.z, mx.x, mx.y, mx.z);
}

int main(int argc, char **argv) {
	try {
		if (argc >= 2 && std::string(argv[1]) == "defaults") { // ref: api/modelAPI.cpp:146-154 + Model.cpp:74-103
			Model model = model_create();
			const int32_t part = model_addEmptyPart(model, "part");
			for (int i = 0; i < 4; i++) { model_addPoint(model, FVector3D((float)(i & 1), (float)(i >> 1), -1.5f)); }
			model_addTriangle(model, part, 0, 1, 2);
			model_addQuad(model, part, 0, 1, 3, 2);
			print_model(model);
			return 0;
		}
		if (argc < 3) { std::fprintf(stderr, "usage: model_test ply|dmf <file> [flipX | detailLevel] | defaults\n"); return 2; }
		if (std::string(argv[1]) == "ply") {
			print_model(importer_loadModel(argv[2], argc > 3 && std::atoi(argv[3]) != 0, Transform3D()));
		} else {
			std::ifstream file(argv[2], std::ios::binary);
			std::stringstream content;
			content << file.rdbuf();
			std::vector<std::pair<std::string, std::string>> names;
			Model model = importFromContent_DMF1(content.str(), argc > 3 ? std::atoi(argv[3]) : 2, &names);
			print_model(model);
			for (const auto &n : names) { std::printf("textures '%s' '%s'\n", n.first.c_str(), n.second.c_str()); }
		}
		return 0;
	} catch (const std::exception &error) {
		std::fprintf(stderr, "error: %s\n", error.what());
		return 1;
	}
}
