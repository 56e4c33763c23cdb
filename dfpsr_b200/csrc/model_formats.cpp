// model_formats.cpp — the scene formats that feed the path: ASCII PLY (ref: SDK/SpriteEngine/importer.cpp) and the
// reference's own DMF1 text format (ref: DFPSR/implementation/render/model/format/dmf1.cpp), parsed on the host into the point /
// polygon arrays the C ABI takes (dfpsr_model, dfpsr_host_model). Host-only code: no kernel, no CUDA call. Numbers are read with the
// reference's own digit-by-digit conversion (ref: api/stringAPI.cpp:1563-1620), not strtod, so every float equals the reference's.
#include "../../include/dfpsr_b200.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

namespace dfpsr { void set_error(const char *fmt, ...); }

namespace {

typedef std::string Str;

bool is_white(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f' || c == '\n' || c == '\r'; } // ref: stringAPI.cpp:1686

Str trim(const Str &s) { // ref: stringAPI.cpp:718 string_removeOuterWhiteSpace
	size_t a = 0, b = s.size();
	while (a < b && is_white(s[a])) { a++; }
	while (b > a && is_white(s[b - 1])) { b--; }
	return s.substr(a, b - a);
}

// ref: stringAPI.cpp:1503-1553 string_split — an element per separator plus the rest when it is not empty
std::vector<Str> split(const Str &s, char separator, bool removeWhiteSpace) {
	std::vector<Str> out;
	size_t start = 0;
	for (size_t i = 0; i < s.size(); i++) {
		if (s[i] == separator) {
			Str e = s.substr(start, i - start);
			out.push_back(removeWhiteSpace ? trim(e) : e);
			start = i + 1;
		}
	}
	if (s.size() > start) { Str e = s.substr(start); out.push_back(removeWhiteSpace ? trim(e) : e); }
	return out;
}

bool same_text(const Str &a, const char *b) { // case-insensitive match on ASCII
	size_t n = strlen(b);
	if (a.size() != n) { return false; }
	for (size_t i = 0; i < n; i++) {
		char x = a[i], y = b[i];
		if (x >= 'a' && x <= 'z') { x = (char)(x - 'a' + 'A'); }
		if (y >= 'a' && y <= 'z') { y = (char)(y - 'a' + 'A'); }
		if (x != y) { return false; }
	}
	return true;
}

long long to_integer(const Str &s, size_t from = 0) { // ref: stringAPI.cpp:1563-1584
	long long result = 0;
	bool negated = false;
	for (size_t i = from; i < s.size(); i++) {
		const char c = s[i];
		if (c == '-' || c == '~') { negated = !negated; }
		else if (c >= '0' && c <= '9') { result = (result * 10) + (int)(c - '0'); }
		else if (c == ',' || c == '.') { break; }
	}
	return negated ? -result : result;
}

double to_double(const Str &s) { // ref: stringAPI.cpp:1586-1620
	double result = 0.0;
	bool negated = false, reachedDecimal = false;
	long long digitDivider = 1;
	for (size_t i = 0; i < s.size(); i++) {
		const char c = s[i];
		if (c == '-' || c == '~') { negated = !negated; }
		else if (c >= '0' && c <= '9') {
			if (reachedDecimal) { digitDivider = digitDivider * 10; result = result + ((double)(c - '0') / (double)digitDivider); }
			else { result = (result * 10) + (double)(c - '0'); }
		} else if (c == ',' || c == '.') { reachedDecimal = true; }
		else if (c == 'e' || c == 'E') { result *= pow(10.0, (double)to_integer(s, i + 1)); break; }
	}
	return negated ? -result : result;
}

struct Builder { // the part of ModelImpl the importers use (ref: implementation/render/model/Model.cpp:281-321)
	std::vector<float> points;
	std::vector<dfpsr_polygon> polygons;
	std::vector<dfpsr_imported_part> parts;
	float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
	int32_t filter = DFPSR_FILTER_SOLID;
	int32_t add_point(float x, float y, float z) {
		const float p[3] = {x, y, z};
		for (int k = 0; k < 3; k++) { if (p[k] < mn[k]) { mn[k] = p[k]; } if (p[k] > mx[k]) { mx[k] = p[k]; } }
		points.push_back(x); points.push_back(y); points.push_back(z);
		return (int32_t)(points.size() / 3) - 1;
	}
	int32_t add_point_if_needed(float x, float y, float z, float threshold) { // ref: Model.cpp:289-321 — the closest point inside the threshold
		float best = threshold;
		int32_t bestIndex = -1;
		for (size_t i = 0; i + 2 < points.size(); i += 3) {
			const float dx = x - points[i], dy = y - points[i + 1], dz = z - points[i + 2];
			const float distance = sqrtf(dx * dx + dy * dy + dz * dz);
			if (distance < best) { best = distance; bestIndex = (int32_t)(i / 3); }
		}
		return bestIndex > -1 ? bestIndex : add_point(x, y, z);
	}
	int32_t add_part(const Str &name) {
		dfpsr_imported_part part;
		memset(&part, 0, sizeof(part));
		snprintf(part.name, sizeof(part.name), "%s", name.c_str());
		part.firstPolygon = (int32_t)polygons.size();
		parts.push_back(part);
		return (int32_t)parts.size() - 1;
	}
	// ref: Model.cpp:74-103 Polygon(indexA, indexB, indexC[, indexD]) — model_addTriangle / model_addQuad defaults
	int32_t add_polygon(int32_t a, int32_t b, int32_t c, int32_t d) {
		dfpsr_polygon p;
		memset(&p, 0, sizeof(p));
		p.pointIndices[0] = a; p.pointIndices[1] = b; p.pointIndices[2] = c; p.pointIndices[3] = d;
		const float tex[4][4] = {{0, 0, 0, 0}, {1, 0, 1, 0}, {1, 1, 1, 1}, {0, 1, 0, 1}};
		for (int k = 0; k < 4; k++) { for (int ch = 0; ch < 4; ch++) { p.texCoords[k][ch] = tex[k][ch]; p.colors[k][ch] = 1.0f; } }
		polygons.push_back(p);
		parts.back().polygonCount++;
		return (int32_t)polygons.size() - 1;
	}
};

int finish(Builder &b, dfpsr_imported_model *out) {
	memset(out, 0, sizeof(*out));
	out->pointCount = (int32_t)(b.points.size() / 3);
	out->polygonCount = (int32_t)b.polygons.size();
	out->partCount = (int32_t)b.parts.size();
	out->filter = b.filter;
	for (int k = 0; k < 3; k++) { out->minBound[k] = b.mn[k]; out->maxBound[k] = b.mx[k]; }
	out->points = (float *)malloc(b.points.size() * sizeof(float) + 1);
	out->polygons = (dfpsr_polygon *)malloc(b.polygons.size() * sizeof(dfpsr_polygon) + 1);
	out->parts = (dfpsr_imported_part *)malloc(b.parts.size() * sizeof(dfpsr_imported_part) + 1);
	if (!out->points || !out->polygons || !out->parts) {
		free(out->points); free(out->polygons); free(out->parts);
		memset(out, 0, sizeof(*out));
		dfpsr::set_error("import: out of host memory");
		return 1;
	}
	if (!b.points.empty()) { memcpy(out->points, b.points.data(), b.points.size() * sizeof(float)); }
	if (!b.polygons.empty()) { memcpy(out->polygons, b.polygons.data(), b.polygons.size() * sizeof(dfpsr_polygon)); }
	if (!b.parts.empty()) { memcpy(out->parts, b.parts.data(), b.parts.size() * sizeof(dfpsr_imported_part)); }
	return 0;
}

struct PlyProperty { Str name; bool list; int32_t scale; };
struct PlyElement { Str name; int32_t count; std::vector<PlyProperty> properties; };
struct PlyVertex { float position[3] = {0, 0, 0}; float color[4] = {1, 1, 1, 1}; };
enum PlyInput { PLY_IGNORE, PLY_VERTEX, PLY_FACE };
PlyInput ply_input(const Str &name) { return same_text(name, "VERTEX") ? PLY_VERTEX : (same_text(name, "FACE") ? PLY_FACE : PLY_IGNORE); }

void set_colors(dfpsr_polygon &p, int vertex, const float *color) { for (int ch = 0; ch < 4; ch++) { p.colors[vertex][ch] = color[ch]; } }

// ref: SDK/SpriteEngine/importer.cpp:52-262 loadPlyModel (into a new model with one part, like importer_loadModel(filename, ...) :280-290)
int load_ply(Builder &b, const Str &content, bool flipX, const dfpsr_transform3d &axis) {
	b.add_part("Imported"); // the reference imports into a part the caller created; one part per file here
	const std::vector<Str> lines = split(content, '\n', true);
	std::vector<PlyElement> elements;
	std::vector<PlyVertex> vertices;
	bool readingContent = false;
	int32_t elementIndex = -1, memberIndex = 0;
	PlyInput mode = PLY_IGNORE;
	if (lines.size() < 2) { dfpsr::set_error("loadPlyModel: Failed to identify line-breaks in the PLY file!"); return 1; }
	if (!same_text(trim(lines[0]), "PLY")) { dfpsr::set_error("loadPlyModel: Failed to identify the file as PLY!"); return 1; }
	if (!same_text(trim(lines[1]), "FORMAT ASCII 1.0")) { dfpsr::set_error("loadPlyModel: Only supporting the ascii 1.0 format!"); return 1; }
	for (size_t l = 0; l < lines.size(); l++) {
		const std::vector<Str> tokens = split(lines[l], ' ', false);
		if (tokens.empty() || same_text(tokens[0], "COMMENT")) { continue; }
		if (readingContent) {
			if (mode == PLY_VERTEX || mode == PLY_FACE) {
				if (mode == PLY_VERTEX) { vertices.push_back(PlyVertex()); }
				const PlyElement &element = elements[(size_t)elementIndex];
				size_t tokenIndex = 0;
				for (size_t pi = 0; pi < element.properties.size(); pi++) {
					if (tokenIndex >= tokens.size()) { break; } // "Undeclared properties" warning in the reference
					const PlyProperty &property = element.properties[pi];
					if (property.list) {
						const int32_t listLength = (int32_t)to_integer(tokens[tokenIndex]);
						tokenIndex++;
						if (mode == PLY_FACE && same_text(property.name, "VERTEX_INDICES")) {
							if (tokenIndex + (size_t)(listLength > 0 ? listLength : 0) > tokens.size()) { dfpsr::set_error("loadPlyModel: a face on line %zu lists more indices than it has", l + 1); return 1; }
							std::vector<int32_t> index((size_t)(listLength > 0 ? listLength : 0));
							for (size_t i = 0; i < index.size(); i++) {
								index[i] = (int32_t)to_integer(tokens[tokenIndex + i]);
								if (index[i] < 0 || (size_t)index[i] >= vertices.size()) { dfpsr::set_error("loadPlyModel: vertex index %d on line %zu is out of bound", index[i], l + 1); return 1; }
							}
							if (listLength == 4) {
								const int order[4] = {flipX ? 3 : 0, flipX ? 2 : 1, flipX ? 1 : 2, flipX ? 0 : 3};
								const int32_t polygon = b.add_polygon(index[order[0]], index[order[1]], index[order[2]], index[order[3]]);
								for (int k = 0; k < 4; k++) { set_colors(b.polygons[(size_t)polygon], k, vertices[(size_t)index[order[k]]].color); }
							} else if (listLength >= 2) {
								int32_t indexA = index[0], indexB = index[1];
								for (int32_t i = 2; i < listLength; i++) { // triangle fan
									const int32_t indexC = index[(size_t)i];
									const int32_t tri[3] = {flipX ? indexC : indexA, indexB, flipX ? indexA : indexC};
									const int32_t polygon = b.add_polygon(tri[0], tri[1], tri[2], -1);
									for (int k = 0; k < 3; k++) { set_colors(b.polygons[(size_t)polygon], k, vertices[(size_t)tri[k]].color); }
									indexB = indexC;
								}
							}
						}
						tokenIndex += (size_t)(listLength > 0 ? listLength : 0);
					} else if (mode == PLY_VERTEX) {
						float value = (float)(to_double(tokens[tokenIndex]) / (double)property.scale);
						PlyVertex &v = vertices.back();
						if (same_text(property.name, "X")) { if (flipX) { value = -value; } v.position[0] = value; }
						else if (same_text(property.name, "Y")) { v.position[1] = value; }
						else if (same_text(property.name, "Z")) { v.position[2] = value; }
						else if (same_text(property.name, "RED")) { v.color[0] = value; }
						else if (same_text(property.name, "GREEN")) { v.color[1] = value; }
						else if (same_text(property.name, "BLUE")) { v.color[2] = value; }
						else if (same_text(property.name, "ALPHA")) { v.color[3] = value; }
					}
					tokenIndex++;
				}
				if (mode == PLY_VERTEX) { // ref: math/Transform3D.h:41-43 transformPoint
					const float *p = vertices.back().position;
					b.add_point((p[0] * axis.xAxis[0] + p[1] * axis.yAxis[0] + p[2] * axis.zAxis[0]) + axis.position[0],
					            (p[0] * axis.xAxis[1] + p[1] * axis.yAxis[1] + p[2] * axis.zAxis[1]) + axis.position[1],
					            (p[0] * axis.xAxis[2] + p[1] * axis.yAxis[2] + p[2] * axis.zAxis[2]) + axis.position[2]);
				}
			}
			memberIndex++;
			if (memberIndex >= elements[(size_t)elementIndex].count) {
				elementIndex++;
				memberIndex = 0;
				if ((size_t)elementIndex >= elements.size()) { return 0; } // remaining lines are ignored
				mode = ply_input(elements[(size_t)elementIndex].name);
			}
		} else if (tokens.size() == 1) {
			if (same_text(tokens[0], "END_HEADER")) {
				readingContent = true; elementIndex = 0; memberIndex = 0;
				if (elements.size() < 2) { dfpsr::set_error("loadPlyModel: Need at least two elements to defined faces and vertices in the model!"); return 1; }
				mode = ply_input(elements[0].name);
			}
		} else if (tokens.size() >= 3) {
			if (same_text(tokens[0], "ELEMENT")) {
				elements.push_back(PlyElement{tokens[1], (int32_t)to_integer(tokens[2]), {}});
				elementIndex = (int32_t)elements.size() - 1;
			} else if (same_text(tokens[0], "PROPERTY")) {
				if (elementIndex < 0) { continue; } // "Cannot declare a property without an element!"
				if (tokens.size() == 3) { elements[(size_t)elementIndex].properties.push_back(PlyProperty{tokens[2], false, same_text(tokens[1], "UCHAR") ? 255 : 1}); }
				else if (tokens.size() == 5 && same_text(tokens[1], "LIST")) { elements[(size_t)elementIndex].properties.push_back(PlyProperty{tokens[4], true, same_text(tokens[3], "UCHAR") ? 255 : 1}); }
				else { dfpsr::set_error("loadPlyModel: Unable to parse property!"); return 1; }
			}
		}
	}
	return 0;
}

// ---- DMF1 (ref: implementation/render/model/format/dmf1.cpp)
struct DmfVertex { float position[3] = {0, 0, 0}; float texCoord[4] = {0, 0, 0, 0}; float color[4] = {1, 1, 1, 1}; };
struct DmfTriangle { DmfVertex vertices[3]; };
struct DmfPart { Str textures[16]; Str shaderZero; int32_t minDetailLevel = 0, maxDetailLevel = 2; std::vector<DmfTriangle> triangles; Str name; };
struct DmfModel { int32_t filter = DFPSR_FILTER_SOLID; std::vector<DmfPart> parts; };
enum { SPACE_MAIN, SPACE_PART, SPACE_TRIANGLE, SPACE_BONE, SPACE_SHAPE, SPACE_POINT, SPACE_UNHANDLED };
enum { WAIT_STATEMENT, WAIT_INDEX_OR_PROPERTY, WAIT_PROPERTY };
struct DmfState { DmfModel *model; int state = WAIT_STATEMENT, space = SPACE_MAIN, propertyIndex = 0; Str lastPropertyName; };

int32_t round_index(double value) { return (int32_t)round(value); }

void dmf_set_property(DmfState &st, const Str &name, int32_t index, const Str &content) { // ref: dmf1.cpp:112-204
	const float value = (float)to_double(content);
	if (st.space == SPACE_MAIN) {
		if (same_text(name, "FilterType")) { st.model->filter = same_text(content, "Alpha") ? DFPSR_FILTER_ALPHA : DFPSR_FILTER_SOLID; }
	} else if (st.space == SPACE_PART) {
		if (st.model->parts.empty()) { return; }
		DmfPart &part = st.model->parts.back();
		if (same_text(name, "Name")) { part.name = content; }
		else if (same_text(name, "Texture")) { if (index >= 0 && index < 16) { part.textures[index] = content; } }
		else if (same_text(name, "Shader")) { if (index == 0) { part.shaderZero = content; } }
		else if (same_text(name, "MinDetailLevel")) { part.minDetailLevel = round_index(value); }
		else if (same_text(name, "MaxDetailLevel")) { part.maxDetailLevel = round_index(value); }
	} else if (st.space == SPACE_TRIANGLE) {
		if (st.model->parts.empty() || st.model->parts.back().triangles.empty() || index < 0 || index > 2) { return; }
		DmfVertex &v = st.model->parts.back().triangles.back().vertices[index];
		if (same_text(name, "X")) { v.position[0] = value; } else if (same_text(name, "Y")) { v.position[1] = value; } else if (same_text(name, "Z")) { v.position[2] = value; }
		else if (same_text(name, "CR")) { v.color[0] = value; } else if (same_text(name, "CG")) { v.color[1] = value; }
		else if (same_text(name, "CB")) { v.color[2] = value; } else if (same_text(name, "CA")) { v.color[3] = value; }
		else if (same_text(name, "U1")) { v.texCoord[0] = value; } else if (same_text(name, "V1")) { v.texCoord[1] = value; }
		else if (same_text(name, "U2")) { v.texCoord[2] = value; } else if (same_text(name, "V2")) { v.texCoord[3] = value; }
	}
}

void dmf_change_namespace(DmfState &st, const Str &name) { // ref: dmf1.cpp:206-230
	if (same_text(name, "Part")) { st.model->parts.push_back(DmfPart()); st.space = SPACE_PART; }
	else if (same_text(name, "Triangle")) {
		if ((st.space == SPACE_PART || st.space == SPACE_TRIANGLE) && !st.model->parts.empty()) { st.model->parts.back().triangles.push_back(DmfTriangle()); st.space = SPACE_TRIANGLE; }
	} else if (same_text(name, "Bone")) { st.space = SPACE_BONE; }
	else if (same_text(name, "Shape")) { st.space = SPACE_SHAPE; }
	else if (same_text(name, "Point")) { st.space = SPACE_POINT; }
	else { st.space = SPACE_UNHANDLED; }
}

void dmf_read_token(DmfState &st, const Str &text, long start, long end) { // ref: dmf1.cpp:234-281 (end is inclusive)
	if (end < start) { return; }
	const char first = text[(size_t)start], last = text[(size_t)end];
	if (first == '(' && last == ')') {
		if (st.state == WAIT_PROPERTY || st.state == WAIT_INDEX_OR_PROPERTY) {
			dmf_set_property(st, st.lastPropertyName, st.propertyIndex, text.substr((size_t)start + 1, (size_t)(end - start - 1)));
			st.state = WAIT_STATEMENT;
			st.propertyIndex = 0;
		}
	} else if (first == '[' && last == ']') {
		if (st.state == WAIT_INDEX_OR_PROPERTY) { st.propertyIndex = round_index(to_double(text.substr((size_t)start + 1, (size_t)(end - start - 1)))); }
	} else if (first == '<' && last == '>') {
		if (st.state == WAIT_STATEMENT && end - start <= 258) { dmf_change_namespace(st, text.substr((size_t)start + 1, (size_t)(end - start - 1))); }
	} else if (st.state == WAIT_STATEMENT && end - start <= 258) {
		st.lastPropertyName = text.substr((size_t)start, (size_t)(end - start + 1));
		st.state = WAIT_INDEX_OR_PROPERTY;
	}
}

int load_dmf1(Builder &b, const Str &text, int32_t detailLevel) {
	DmfModel native;
	DmfState st;
	st.model = &native;
	if (text.size() < 4 || text[0] != 'D' || text[1] != 'M' || text[2] != 'F' || text[3] != '1') { dfpsr::set_error("The file does not start with \"DMF1\"!"); return 1; }
	long tokenStart = 4, readIndex = 4; // ref: dmf1.cpp:284-327 loadNative_DMF1
	char firstCharOfToken = '\0';
	for (readIndex = tokenStart; readIndex < (long)text.size(); readIndex++) {
		const char c = text[(size_t)readIndex];
		if (firstCharOfToken == '\0' && (c == '\t' || c == ' ' || c == '\n' || c == '\r')) { dmf_read_token(st, text, tokenStart, readIndex - 1); tokenStart = readIndex + 1; }
		else if (c == '<' || c == '(' || c == '[') { dmf_read_token(st, text, tokenStart, readIndex - 1); tokenStart = readIndex; firstCharOfToken = c; }
		else if ((firstCharOfToken == '<' && c == '>') || (firstCharOfToken == '(' && c == ')') || (firstCharOfToken == '[' && c == ']')) {
			dmf_read_token(st, text, tokenStart, readIndex); tokenStart = readIndex + 1; firstCharOfToken = '\0';
		}
	}
	dmf_read_token(st, text, tokenStart, readIndex - 1);
	// ref: dmf1.cpp:329-367 convertFromDMF1
	// FilterType is parsed into the native model but convertFromDMF1 never hands it to the result (dmf1.cpp:329-367): the imported model
	// keeps the default Filter::Solid, and so does this importer.
	(void)native.filter;
	for (const DmfPart &part : native.parts) {
		if (detailLevel < part.minDetailLevel || detailLevel > part.maxDetailLevel) { continue; }
		const int32_t index = b.add_part(part.name);
		dfpsr_imported_part &target = b.parts[(size_t)index];
		if (same_text(part.shaderZero, "M_Diffuse_1Tex") || same_text(part.shaderZero, "M_Diffuse_2Tex")) { snprintf(target.diffuseName, sizeof(target.diffuseName), "%s", part.textures[0].c_str()); }
		if (same_text(part.shaderZero, "M_Diffuse_2Tex")) { snprintf(target.lightName, sizeof(target.lightName), "%s", part.textures[1].c_str()); }
		for (const DmfTriangle &t : part.triangles) {
			int32_t point[3];
			for (int k = 0; k < 3; k++) { point[k] = b.add_point_if_needed(t.vertices[k].position[0], t.vertices[k].position[1], t.vertices[k].position[2], 0.00001f); }
			const int32_t polygon = b.add_polygon(point[0], point[1], point[2], -1);
			dfpsr_polygon &p = b.polygons[(size_t)polygon]; // ref: Model.cpp:44-57 Polygon(vertA, vertB, vertC): the fourth corner is zeroed
			for (int k = 0; k < 3; k++) { for (int ch = 0; ch < 4; ch++) { p.texCoords[k][ch] = t.vertices[k].texCoord[ch]; p.colors[k][ch] = t.vertices[k].color[ch]; } }
			for (int ch = 0; ch < 4; ch++) { p.texCoords[3][ch] = 0.0f; p.colors[3][ch] = 0.0f; }
		}
	}
	return 0;
}

} // namespace

extern "C" {

int dfpsr_import_ply(const char *content, size_t length, int32_t flipX, const dfpsr_transform3d *axisConversion, dfpsr_imported_model *out) {
	if (!content || !out) { dfpsr::set_error("import_ply: null argument"); return 1; }
	dfpsr_transform3d identity;
	memset(&identity, 0, sizeof(identity));
	identity.xAxis[0] = identity.yAxis[1] = identity.zAxis[2] = 1.0f;
	Builder b;
	if (load_ply(b, Str(content, length), flipX != 0, axisConversion ? *axisConversion : identity)) { memset(out, 0, sizeof(*out)); return 1; }
	return finish(b, out);
}

int dfpsr_import_dmf1(const char *content, size_t length, int32_t detailLevel, dfpsr_imported_model *out) {
	if (!content || !out) { dfpsr::set_error("import_dmf1: null argument"); return 1; }
	Builder b;
	if (load_dmf1(b, Str(content, length), detailLevel)) { memset(out, 0, sizeof(*out)); return 1; }
	return finish(b, out);
}

void dfpsr_import_free(dfpsr_imported_model *model) {
	if (!model) { return; }
	free(model->points); free(model->polygons); free(model->parts);
	memset(model, 0, sizeof(*model));
}

} // extern "C"
