// model_formats.cpp — the two scene formats that feed the path, read on the host into the point / polygon arrays
// of the C ABI (dfpsr_model, dfpsr_host_model). Host-only: no kernel, no CUDA call.
//
//   ASCII PLY 1.0   the files importer_loadModel accepts (ref: SDK/SpriteEngine/importer.cpp:52-290)
//   DMF1            the reference's own text format (ref: DFPSR/implementation/render/model/format/dmf1.cpp)
//
// Own design, same language: both readers work on spans of the caller's buffer (nothing is copied into strings or lists).
//   PLY   the header is compiled ONCE into a schema — per element a short program of slot operations (store coordinate, store colour
//         channel, emit face, skip) — and the body is executed element by element against that program, instead of re-identifying
//         every property name for every token of every line.
//   DMF1  a scanner cuts the text into lexemes of four kinds (word, <section>, [index], (value)); an assembler folds them into
//         `name [index] (value)` assignments, which a static table routes to their field by (section, name).
// What must be identical to the reference is the accepted language and the resulting numbers, so three behaviours are reproduced on
// purpose and pinned by tests/test_importers.py against the compiled reference: (1) decimal text becomes a double through the
// reference's digit-by-digit arithmetic (ref: api/stringAPI.cpp:1563-1620), not strtod — the last bit of many coordinates depends on
// it; (2) a PLY line is cut at every single space (two spaces make an empty field), a list field takes its length + 2 fields, and an
// element takes at least one line; (3) DMF1 points merge with the CLOSEST earlier point within 0.00001 (Model.cpp:289-321) — found
// here through a hash grid instead of a scan over every earlier point.
#include "../../include/dfpsr_b200.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <unordered_map>
#include <vector>

namespace dfpsr { void set_error(const char *fmt, ...); }

namespace {

// ------------------------------------------------------------------------------------------------ spans of the input

struct Span {
	const char *first, *last; // [first, last)
	size_t size() const { return (size_t)(last - first); }
	bool empty() const { return first >= last; }
};

inline bool blank(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f' || c == '\n' || c == '\r'; }
inline char upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 'a' + 'A') : c; }

Span stripped(Span s) {
	while (s.first < s.last && blank(*s.first)) { s.first++; }
	while (s.last > s.first && blank(s.last[-1])) { s.last--; }
	return s;
}

bool is_word(Span s, const char *word) { // ASCII case-insensitive equality
	const size_t n = strlen(word);
	if (s.size() != n) { return false; }
	for (size_t i = 0; i < n; i++) { if (upper(s.first[i]) != upper(word[i])) { return false; } }
	return true;
}

void copy_name(char *target, size_t capacity, Span s) {
	const size_t n = s.first == nullptr ? 0 : (s.size() < capacity - 1 ? s.size() : capacity - 1);
	if (n > 0) { memcpy(target, s.first, n); }
	target[n] = '\0';
}

// Decimal text to a number with the reference's arithmetic (api/stringAPI.cpp:1563-1620): every '-' or '~' flips the sign, digits
// accumulate as result * 10 + digit, fraction digits add digit / 10^k with an integer power, an exponent multiplies by pow(10, integer).
// Any other character is skipped. strtod would round differently in the last place.
long long whole_number(const char *p, const char *end) {
	long long magnitude = 0;
	bool negative = false;
	for (; p < end; p++) {
		if (*p == '-' || *p == '~') { negative = !negative; }
		else if (*p >= '0' && *p <= '9') { magnitude = magnitude * 10 + (*p - '0'); }
		else if (*p == '.' || *p == ',') { break; }
	}
	return negative ? -magnitude : magnitude;
}
long long whole_number(Span s) { return whole_number(s.first, s.last); }

double real_number(Span s) {
	double magnitude = 0.0;
	long long scale = 1;
	bool negative = false, fraction = false;
	for (const char *p = s.first; p < s.last; p++) {
		const char c = *p;
		if (c == '-' || c == '~') { negative = !negative; }
		else if (c >= '0' && c <= '9') {
			if (fraction) { scale *= 10; magnitude = magnitude + (double)(c - '0') / (double)scale; }
			else { magnitude = magnitude * 10 + (double)(c - '0'); }
		} else if (c == '.' || c == ',') { fraction = true; }
		else if (c == 'e' || c == 'E') { magnitude *= pow(10.0, (double)whole_number(p + 1, s.last)); break; }
	}
	return negative ? -magnitude : magnitude;
}

// ------------------------------------------------------------------------------------------------ the model being assembled

struct Assembly {
	std::vector<float> points;
	std::vector<dfpsr_polygon> polygons;
	std::vector<dfpsr_imported_part> parts;
	float low[3] = {0.0f, 0.0f, 0.0f}, high[3] = {0.0f, 0.0f, 0.0f}; // the reference's bound starts at the origin (Model.cpp:281-287)

	int32_t point(float x, float y, float z) {
		const float p[3] = {x, y, z};
		for (int k = 0; k < 3; k++) { low[k] = p[k] < low[k] ? p[k] : low[k]; high[k] = p[k] > high[k] ? p[k] : high[k]; }
		points.insert(points.end(), p, p + 3);
		return (int32_t)(points.size() / 3) - 1;
	}
	void open_part(Span name) {
		dfpsr_imported_part part;
		memset(&part, 0, sizeof(part));
		copy_name(part.name, sizeof(part.name), name);
		part.firstPolygon = (int32_t)polygons.size();
		parts.push_back(part);
	}
	// A new polygon of the open part with the reference's defaults (Model.cpp:74-103): white corners, texture coordinates spanning the texture.
	dfpsr_polygon &polygon(int32_t a, int32_t b, int32_t c, int32_t d) {
		static const float corner[4][2] = {{0.0f, 0.0f}, {1.0f, 0.0f}, {1.0f, 1.0f}, {0.0f, 1.0f}};
		dfpsr_polygon p;
		memset(&p, 0, sizeof(p));
		const int32_t index[4] = {a, b, c, d};
		for (int k = 0; k < 4; k++) {
			p.pointIndices[k] = index[k];
			p.texCoords[k][0] = p.texCoords[k][2] = corner[k][0]; p.texCoords[k][1] = p.texCoords[k][3] = corner[k][1];
			for (int ch = 0; ch < 4; ch++) { p.colors[k][ch] = 1.0f; }
		}
		polygons.push_back(p);
		parts.back().polygonCount++;
		return polygons.back();
	}
	int release(dfpsr_imported_model *out) {
		memset(out, 0, sizeof(*out));
		out->pointCount = (int32_t)(points.size() / 3); out->polygonCount = (int32_t)polygons.size(); out->partCount = (int32_t)parts.size();
		out->filter = DFPSR_FILTER_SOLID;
		memcpy(out->minBound, low, sizeof(low)); memcpy(out->maxBound, high, sizeof(high));
		out->points = (float *)malloc(points.size() * sizeof(float) + 1);
		out->polygons = (dfpsr_polygon *)malloc(polygons.size() * sizeof(dfpsr_polygon) + 1);
		out->parts = (dfpsr_imported_part *)malloc(parts.size() * sizeof(dfpsr_imported_part) + 1);
		if (!out->points || !out->polygons || !out->parts) {
			free(out->points); free(out->polygons); free(out->parts);
			memset(out, 0, sizeof(*out));
			dfpsr::set_error("import: the host is out of memory");
			return 1;
		}
		if (!points.empty()) { memcpy(out->points, points.data(), points.size() * sizeof(float)); }
		if (!polygons.empty()) { memcpy(out->polygons, polygons.data(), polygons.size() * sizeof(dfpsr_polygon)); }
		if (!parts.empty()) { memcpy(out->parts, parts.data(), parts.size() * sizeof(dfpsr_imported_part)); }
		return 0;
	}
};

// ------------------------------------------------------------------------------------------------ PLY

// What one declared property does with its field(s) of a body line.
enum SlotOp : uint8_t { OP_SKIP, OP_SKIP_LIST, OP_X, OP_Y, OP_Z, OP_RED, OP_GREEN, OP_BLUE, OP_ALPHA, OP_FACE };
struct Slot { SlotOp op; double divisor; };
enum ElementRole : uint8_t { ROLE_OTHER, ROLE_VERTEX, ROLE_FACE };
struct ElementPlan { ElementRole role; long long members; std::vector<Slot> slots; };
struct PlyCorner { float position[3]; float color[4]; };

// Lines of the file, outer white space removed; blank lines and comments never reach the caller.
struct PlyLines {
	const char *at, *end;
	bool next(Span &line) {
		while (at < end) {
			const char *stop = (const char *)memchr(at, '\n', (size_t)(end - at));
			Span raw = {at, stop ? stop : end};
			at = stop ? stop + 1 : end;
			line = stripped(raw);
			if (line.empty()) { continue; }
			Span head = {line.first, (const char *)memchr(line.first, ' ', line.size())};
			if (head.last == nullptr) { head.last = line.last; }
			if (is_word(head, "comment")) { continue; }
			return true;
		}
		return false;
	}
};

// Fields of a line: cut at every single space, so consecutive spaces make empty fields (they count, like in the reference).
void cut_fields(Span line, std::vector<Span> &fields) {
	fields.clear();
	const char *start = line.first;
	for (const char *p = line.first; p < line.last; p++) {
		if (*p == ' ') { fields.push_back(Span{start, p}); start = p + 1; }
	}
	if (start < line.last) { fields.push_back(Span{start, line.last}); }
}

SlotOp scalar_op(ElementRole role, Span name) {
	if (role != ROLE_VERTEX) { return OP_SKIP; }
	static const struct { const char *name; SlotOp op; } table[] = {{"x", OP_X}, {"y", OP_Y}, {"z", OP_Z}, {"red", OP_RED}, {"green", OP_GREEN}, {"blue", OP_BLUE}, {"alpha", OP_ALPHA}};
	for (const auto &entry : table) { if (is_word(name, entry.name)) { return entry.op; } }
	return OP_SKIP;
}

int read_ply(Assembly &model, Span text, bool mirror, const dfpsr_transform3d &axis) {
	PlyLines lines = {text.first, text.last};
	std::vector<Span> fields;
	Span line;
	// ---- signature: "ply", then "format ascii 1.0"
	{
		// the signature lines are the first two lines of the file as they stand (a comment in front of them is not allowed)
		const char *firstBreak = (const char *)memchr(text.first, '\n', text.size());
		if (firstBreak == nullptr) { dfpsr::set_error("PLY import: the text has no line breaks, so it cannot hold a header"); return 1; }
		const char *secondBreak = (const char *)memchr(firstBreak + 1, '\n', (size_t)(text.last - firstBreak - 1));
		if (!is_word(stripped(Span{text.first, firstBreak}), "ply")) { dfpsr::set_error("PLY import: the first line must read \"ply\""); return 1; }
		if (!is_word(stripped(Span{firstBreak + 1, secondBreak ? secondBreak : text.last}), "format ascii 1.0")) { dfpsr::set_error("PLY import: only \"format ascii 1.0\" files are supported"); return 1; }
	}
	model.open_part(Span{"Imported", "Imported" + 8}); // the reference imports into a part its caller created: one part per file here
	// ---- header -> schema
	std::vector<ElementPlan> plan;
	bool headerClosed = false;
	while (!headerClosed && lines.next(line)) {
		cut_fields(line, fields);
		if (fields.size() == 1) { headerClosed = is_word(fields[0], "end_header"); continue; }
		if (fields.size() < 3) { continue; }
		if (is_word(fields[0], "element")) {
			ElementPlan element;
			element.role = is_word(fields[1], "vertex") ? ROLE_VERTEX : (is_word(fields[1], "face") ? ROLE_FACE : ROLE_OTHER);
			element.members = whole_number(fields[2]);
			plan.push_back(element);
		} else if (is_word(fields[0], "property") && !plan.empty()) {
			ElementPlan &element = plan.back();
			if (fields.size() == 3) {
				element.slots.push_back(Slot{scalar_op(element.role, fields[2]), is_word(fields[1], "uchar") ? 255.0 : 1.0});
			} else if (fields.size() == 5 && is_word(fields[1], "list")) {
				element.slots.push_back(Slot{element.role == ROLE_FACE && is_word(fields[4], "vertex_indices") ? OP_FACE : OP_SKIP_LIST, 1.0});
			} else {
				dfpsr::set_error("PLY import: a property is declared as \"property <type> <name>\" or \"property list <type> <type> <name>\"");
				return 1;
			}
		}
	}
	if (!headerClosed) { return 0; } // no body: an empty model, like the reference
	if (plan.size() < 2) { dfpsr::set_error("PLY import: the header must declare at least a vertex and a face element"); return 1; }
	// ---- body: element after element, every member line run through the element's slots
	std::vector<PlyCorner> corners;
	std::vector<int32_t> loop;
	for (const ElementPlan &element : plan) {
		const long long members = element.members > 1 ? element.members : 1; // an element takes at least one line
		for (long long m = 0; m < members; m++) {
			if (!lines.next(line)) { return 0; }
			if (element.role == ROLE_OTHER) { continue; }
			cut_fields(line, fields);
			PlyCorner corner = {{0.0f, 0.0f, 0.0f}, {1.0f, 1.0f, 1.0f, 1.0f}};
			size_t at = 0;
			for (const Slot &slot : element.slots) {
				if (at >= fields.size()) { break; }
				if (slot.op == OP_SKIP_LIST || slot.op == OP_FACE) {
					const long long declared = whole_number(fields[at]);
					const size_t length = declared > 0 ? (size_t)declared : 0;
					if (slot.op == OP_FACE) {
						if (at + 1 + length > fields.size()) { dfpsr::set_error("PLY import: a face announces %lld corners but its line holds fewer", declared); return 1; }
						loop.resize(length);
						for (size_t i = 0; i < length; i++) {
							loop[i] = (int32_t)whole_number(fields[at + 1 + i]);
							if (loop[i] < 0 || (size_t)loop[i] >= corners.size()) { dfpsr::set_error("PLY import: a face refers to vertex %d of %zu", loop[i], corners.size()); return 1; }
						}
						auto paint = [&](dfpsr_polygon &p, int k, int32_t vertex) { memcpy(p.colors[k], corners[(size_t)vertex].color, sizeof(p.colors[k])); };
						if (length == 4) { // quads stay quads; mirroring reverses the winding
							const int32_t q[4] = {loop[mirror ? 3 : 0], loop[mirror ? 2 : 1], loop[mirror ? 1 : 2], loop[mirror ? 0 : 3]};
							dfpsr_polygon &p = model.polygon(q[0], q[1], q[2], q[3]);
							for (int k = 0; k < 4; k++) { paint(p, k, q[k]); }
						} else {
							for (size_t i = 2; i < length; i++) { // everything else becomes a fan around its first corner
								const int32_t t[3] = {mirror ? loop[i] : loop[0], loop[i - 1], mirror ? loop[0] : loop[i]};
								dfpsr_polygon &p = model.polygon(t[0], t[1], t[2], -1);
								for (int k = 0; k < 3; k++) { paint(p, k, t[k]); }
							}
						}
					}
					at += length + 2; // the length field, the entries, and one more field (the reference's list reader does the same)
					continue;
				}
				if (slot.op != OP_SKIP) {
					const float value = (float)(real_number(fields[at]) / slot.divisor);
					switch (slot.op) {
						case OP_X: corner.position[0] = mirror ? -value : value; break;
						case OP_Y: corner.position[1] = value; break;
						case OP_Z: corner.position[2] = value; break;
						default: corner.color[slot.op - OP_RED] = value; break;
					}
				}
				at++;
			}
			if (element.role == ROLE_VERTEX) {
				corners.push_back(corner);
				const float *p = corner.position; // ref: math/Transform3D.h:41-43 — x * xAxis + y * yAxis + z * zAxis, then the offset
				model.point((p[0] * axis.xAxis[0] + p[1] * axis.yAxis[0] + p[2] * axis.zAxis[0]) + axis.position[0],
				            (p[0] * axis.xAxis[1] + p[1] * axis.yAxis[1] + p[2] * axis.zAxis[1]) + axis.position[1],
				            (p[0] * axis.xAxis[2] + p[1] * axis.yAxis[2] + p[2] * axis.zAxis[2]) + axis.position[2]);
			}
		}
	}
	return 0;
}

// ------------------------------------------------------------------------------------------------ DMF1

enum LexemeKind : uint8_t { LEX_WORD, LEX_SECTION, LEX_INDEX, LEX_VALUE };
struct Lexeme { LexemeKind kind; Span body; }; // body: the word itself, or what stands between the brackets

// Cuts the text behind the "DMF1" signature into lexemes. White space separates lexemes outside of brackets only; an opening bracket
// always starts a new lexeme, which ends with the closing bracket of the same kind.
struct DmfScanner {
	const char *at, *end;
	const char *start;
	char open = '\0';
	bool classify(const char *first, const char *last, Lexeme &out) const { // [first, last]
		if (last < first) { return false; }
		const char a = *first, z = *last;
		const size_t length = (size_t)(last - first) + 1;
		if (a == '(' && z == ')') { out = Lexeme{LEX_VALUE, Span{first + 1, last}}; return true; }
		if (a == '[' && z == ']') { out = Lexeme{LEX_INDEX, Span{first + 1, last}}; return true; }
		if (length > 259) { return false; } // names and section titles are at most 259 characters, longer ones are dropped
		if (a == '<' && z == '>') { out = Lexeme{LEX_SECTION, Span{first + 1, last}}; return true; }
		out = Lexeme{LEX_WORD, Span{first, last + 1}};
		return true;
	}
	bool next(Lexeme &out) {
		while (at < end) {
			const char c = *at;
			const char *here = at++;
			if (open == '\0' && (c == ' ' || c == '\t' || c == '\n' || c == '\r')) {
				const char *first = start;
				start = here + 1;
				if (classify(first, here - 1, out)) { return true; }
			} else if (c == '<' || c == '(' || c == '[') {
				const char *first = start;
				start = here; open = c;
				if (classify(first, here - 1, out)) { return true; }
			} else if ((open == '<' && c == '>') || (open == '(' && c == ')') || (open == '[' && c == ']')) {
				const char *first = start;
				start = here + 1; open = '\0';
				if (classify(first, here, out)) { return true; }
			}
		}
		if (start < end) { const char *first = start; start = end; return classify(first, end - 1, out); }
		return false;
	}
};

struct DmfCorner { float position[3] = {0, 0, 0}; float texture[4] = {0, 0, 0, 0}; float color[4] = {1, 1, 1, 1}; };
struct DmfFace { DmfCorner corner[3]; };
struct DmfPiece {
	Span name = {nullptr, nullptr}, textures[16] = {}, shader = {nullptr, nullptr};
	int32_t lowestDetail = 0, highestDetail = 2;
	std::vector<DmfFace> faces;
};
enum DmfSection : uint8_t { IN_FILE, IN_PART, IN_TRIANGLE, IN_IGNORED };

// (section, property name) -> what the assignment writes
enum DmfField : uint8_t { F_PART_NAME, F_PART_TEXTURE, F_PART_SHADER, F_PART_MIN, F_PART_MAX, F_POSITION, F_COLOR, F_TEXTURE };
static const struct { DmfSection section; const char *name; DmfField field; int lane; } DMF_FIELDS[] = {
	{IN_PART, "Name", F_PART_NAME, 0}, {IN_PART, "Texture", F_PART_TEXTURE, 0}, {IN_PART, "Shader", F_PART_SHADER, 0},
	{IN_PART, "MinDetailLevel", F_PART_MIN, 0}, {IN_PART, "MaxDetailLevel", F_PART_MAX, 0},
	{IN_TRIANGLE, "X", F_POSITION, 0}, {IN_TRIANGLE, "Y", F_POSITION, 1}, {IN_TRIANGLE, "Z", F_POSITION, 2},
	{IN_TRIANGLE, "CR", F_COLOR, 0}, {IN_TRIANGLE, "CG", F_COLOR, 1}, {IN_TRIANGLE, "CB", F_COLOR, 2}, {IN_TRIANGLE, "CA", F_COLOR, 3},
	{IN_TRIANGLE, "U1", F_TEXTURE, 0}, {IN_TRIANGLE, "V1", F_TEXTURE, 1}, {IN_TRIANGLE, "U2", F_TEXTURE, 2}, {IN_TRIANGLE, "V2", F_TEXTURE, 3},
};

// Points of a DMF1 model merge with the closest earlier point within `reach` (ties: the earliest). A hash grid of cells `reach` wide
// keeps the search local: a point's partners lie in its own cell or one of the 26 around it. Coordinates too large for the grid's
// integer cells fall back to the plain scan.
struct PointMerger {
	Assembly &model;
	const float reach;
	std::unordered_map<uint64_t, std::vector<int32_t>> cells;
	std::vector<int32_t> unhashed;
	PointMerger(Assembly &m, float r) : model(m), reach(r) {}
	static bool cell_of(const float *p, float reach, int64_t cell[3]) {
		for (int k = 0; k < 3; k++) {
			const double scaled = floor((double)p[k] / (double)reach);
			if (!(fabs(scaled) < 1.0e15)) { return false; }
			cell[k] = (int64_t)scaled;
		}
		return true;
	}
	// equal cells give equal keys; different cells sharing a key only add candidates, every one of which is measured
	static uint64_t key_of(int64_t x, int64_t y, int64_t z) { return ((uint64_t)x * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)y * 0xC2B2AE3D27D4EB4Full) ^ ((uint64_t)z * 0x165667B19E3779F9ull); }
	void consider(int32_t candidate, const float *p, float &best, int32_t &bestIndex) const {
		const float *q = &model.points[(size_t)candidate * 3];
		const float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
		const float distance = sqrtf(dx * dx + dy * dy + dz * dz);
		if (distance < best || (distance == best && bestIndex >= 0 && candidate < bestIndex && distance < reach)) { best = distance; bestIndex = candidate; }
	}
	int32_t merge(float x, float y, float z) {
		const float p[3] = {x, y, z};
		float best = reach;
		int32_t bestIndex = -1;
		int64_t cell[3];
		const bool hashed = cell_of(p, reach, cell);
		if (hashed) {
			for (int64_t dx = -1; dx <= 1; dx++) for (int64_t dy = -1; dy <= 1; dy++) for (int64_t dz = -1; dz <= 1; dz++) {
				auto found = cells.find(key_of(cell[0] + dx, cell[1] + dy, cell[2] + dz));
				if (found == cells.end()) { continue; }
				for (int32_t candidate : found->second) { consider(candidate, p, best, bestIndex); }
			}
			for (int32_t candidate : unhashed) { consider(candidate, p, best, bestIndex); }
		} else {
			for (int32_t candidate = 0; candidate < (int32_t)(model.points.size() / 3); candidate++) { consider(candidate, p, best, bestIndex); }
		}
		if (bestIndex >= 0) { return bestIndex; }
		const int32_t fresh = model.point(x, y, z);
		if (hashed) { cells[key_of(cell[0], cell[1], cell[2])].push_back(fresh); } else { unhashed.push_back(fresh); }
		return fresh;
	}
};

int read_dmf1(Assembly &model, Span text, int32_t detail) {
	if (text.size() < 4 || memcmp(text.first, "DMF1", 4) != 0) { dfpsr::set_error("DMF1 import: the text does not begin with the signature \"DMF1\""); return 1; }
	std::vector<DmfPiece> pieces;
	DmfSection section = IN_FILE;
	// assembler state: a word opens an assignment, an index may follow, a value closes it
	bool pending = false;
	Span pendingName = {nullptr, nullptr};
	int32_t pendingIndex = 0;
	DmfScanner scanner = {text.first + 4, text.last, text.first + 4};
	Lexeme lexeme;
	while (scanner.next(lexeme)) {
		switch (lexeme.kind) {
		case LEX_WORD:
			if (!pending) { pending = true; pendingName = lexeme.body; pendingIndex = 0; }
			break;
		case LEX_INDEX:
			if (pending) { pendingIndex = (int32_t)round(real_number(lexeme.body)); }
			break;
		case LEX_SECTION:
			if (pending) { break; } // a section title inside an open assignment is ignored
			if (is_word(lexeme.body, "Part")) { pieces.emplace_back(); section = IN_PART; }
			else if (is_word(lexeme.body, "Triangle")) {
				if ((section == IN_PART || section == IN_TRIANGLE) && !pieces.empty()) { pieces.back().faces.emplace_back(); section = IN_TRIANGLE; }
			} else { section = IN_IGNORED; } // bones, shapes, points and anything unknown
			break;
		case LEX_VALUE: {
			if (!pending) { break; }
			pending = false;
			if (pieces.empty()) { break; } // only the file section comes before the first part; its FilterType never reaches the imported model (dmf1.cpp:329-367)
			DmfPiece &piece = pieces.back();
			for (const auto &entry : DMF_FIELDS) {
				if (entry.section != section || !is_word(pendingName, entry.name)) { continue; }
				const float number = (float)real_number(lexeme.body);
				switch (entry.field) {
				case F_PART_NAME: piece.name = lexeme.body; break;
				case F_PART_TEXTURE: if (pendingIndex >= 0 && pendingIndex < 16) { piece.textures[pendingIndex] = lexeme.body; } break;
				case F_PART_SHADER: if (pendingIndex == 0) { piece.shader = lexeme.body; } break;
				case F_PART_MIN: piece.lowestDetail = (int32_t)round((double)number); break;
				case F_PART_MAX: piece.highestDetail = (int32_t)round((double)number); break;
				default:
					if (piece.faces.empty() || pendingIndex < 0 || pendingIndex > 2) { break; }
					DmfCorner &corner = piece.faces.back().corner[pendingIndex];
					(entry.field == F_POSITION ? corner.position : (entry.field == F_COLOR ? corner.color : corner.texture))[entry.lane] = number;
					break;
				}
				break;
			}
			break;
		}
		}
	}
	// ---- the parts of the requested detail level become parts of the model
	PointMerger merger(model, 0.00001f);
	for (const DmfPiece &piece : pieces) {
		if (detail < piece.lowestDetail || detail > piece.highestDetail) { continue; }
		model.open_part(piece.name);
		dfpsr_imported_part &part = model.parts.back();
		const bool one = is_word(piece.shader, "M_Diffuse_1Tex"), two = is_word(piece.shader, "M_Diffuse_2Tex");
		if (one || two) { copy_name(part.diffuseName, sizeof(part.diffuseName), piece.textures[0]); }
		if (two) { copy_name(part.lightName, sizeof(part.lightName), piece.textures[1]); }
		for (const DmfFace &face : piece.faces) {
			int32_t index[3];
			for (int k = 0; k < 3; k++) { index[k] = merger.merge(face.corner[k].position[0], face.corner[k].position[1], face.corner[k].position[2]); }
			dfpsr_polygon &p = model.polygon(index[0], index[1], index[2], -1);
			for (int k = 0; k < 3; k++) { memcpy(p.texCoords[k], face.corner[k].texture, sizeof(p.texCoords[k])); memcpy(p.colors[k], face.corner[k].color, sizeof(p.colors[k])); }
			memset(p.texCoords[3], 0, sizeof(p.texCoords[3])); memset(p.colors[3], 0, sizeof(p.colors[3])); // a triangle's fourth corner is all zero (Model.cpp:44-57)
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
	Assembly model;
	if (read_ply(model, Span{content, content + length}, flipX != 0, axisConversion ? *axisConversion : identity)) { memset(out, 0, sizeof(*out)); return 1; }
	return model.release(out);
}

int dfpsr_import_dmf1(const char *content, size_t length, int32_t detailLevel, dfpsr_imported_model *out) {
	if (!content || !out) { dfpsr::set_error("import_dmf1: null argument"); return 1; }
	Assembly model;
	if (read_dmf1(model, Span{content, content + length}, detailLevel)) { memset(out, 0, sizeof(*out)); return 1; }
	return model.release(out);
}

void dfpsr_import_free(dfpsr_imported_model *model) {
	if (!model) { return; }
	free(model->points); free(model->polygons); free(model->parts);
	memset(model, 0, sizeof(*model));
}

} // extern "C"
