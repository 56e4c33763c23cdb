// sprite_math.cuh — float and integer vector arithmetic of the Sandbox sprite engine, written so that every expression rounds exactly
// like the reference's (same operand order, no contraction: the library is built with -fmad=false / -ffp-contract=off).
//   ref: DFPSR/math/FVector.h, FMatrix2x2.h, FMatrix3x3.h, Transform3D.h, IVector.h, IRect.h
#pragma once

#include "common.cuh"
#include <math.h>

namespace dfpsr {
namespace sw {

struct F2 { float x, y; };
struct F3 { float x, y, z; };
struct I2 { int32_t x, y; };
struct I3 { int32_t x, y, z; };
struct M3 { F3 x, y, z; };      // xAxis, yAxis, zAxis
struct T3 { F3 position; M3 m; }; // ref: math/Transform3D.h:33-36

// The reference converts float to int32 with cvttss2si: NaN and out-of-range values give INT32_MIN.
__host__ __device__ inline int32_t f2i(float v) {
	if (!(v > -2147483904.0f && v < 2147483648.0f)) { return (int32_t)0x80000000; }
	return (int32_t)v;
}
// (uint32_t)float on x86-64 goes through the 64-bit signed conversion and keeps the low 32 bits.
__host__ __device__ inline uint32_t f2u(float v) {
	if (!(v > -9223373136366403584.0f && v < 9223372036854775808.0f)) { return 0u; }
	return (uint32_t)(long long)v;
}

__host__ __device__ inline F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ inline F3 f3(const float *p) { return f3(p[0], p[1], p[2]); }
__host__ __device__ inline F3 add(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ inline F3 sub(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ inline F3 scale(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ inline F3 cross(F3 a, F3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } // FVector.h:101
__host__ __device__ inline float length3(F3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }                                  // FVector.h:84-95
__host__ __device__ inline F3 normalize(F3 v) {                                                                                       // FVector.h:113-120
	float l = length3(v);
	if (l == 0.0f) { return f3(0.0f, 0.0f, 1.0f); }
	return f3(v.x / l, v.y / l, v.z / l);
}
// ref: math/FMatrix3x3.h:55-61
__host__ __device__ inline F3 transform(const M3 &m, F3 p) {
	return f3(p.x * m.x.x + p.y * m.y.x + p.z * m.z.x, p.x * m.x.y + p.y * m.y.y + p.z * m.z.y, p.x * m.x.z + p.y * m.y.z + p.z * m.z.z);
}
// ref: math/FMatrix3x3.h:66-72
__host__ __device__ inline F3 transform_transposed(const M3 &m, F3 p) {
	return f3(p.x * m.x.x + p.y * m.x.y + p.z * m.x.z, p.x * m.y.x + p.y * m.y.y + p.z * m.y.z, p.x * m.z.x + p.y * m.z.y + p.z * m.z.z);
}
__host__ __device__ inline M3 m3(F3 x, F3 y, F3 z) { M3 r; r.x = x; r.y = y; r.z = z; return r; }
inline M3 m3(const dfpsr_matrix3x3 &m) { return m3(f3(m.xAxis), f3(m.yAxis), f3(m.zAxis)); }
inline void store(dfpsr_matrix3x3 &out, const M3 &m) {
	out.xAxis[0] = m.x.x; out.xAxis[1] = m.x.y; out.xAxis[2] = m.x.z;
	out.yAxis[0] = m.y.x; out.yAxis[1] = m.y.y; out.yAxis[2] = m.y.z;
	out.zAxis[0] = m.z.x; out.zAxis[1] = m.z.y; out.zAxis[2] = m.z.z;
}
inline M3 mul(const M3 &left, const M3 &right) { return m3(transform(right, left.x), transform(right, left.y), transform(right, left.z)); } // FMatrix3x3.h:78-80
inline M3 transpose(const M3 &m) { return m3(f3(m.x.x, m.y.x, m.z.x), f3(m.x.y, m.y.y, m.z.y), f3(m.x.z, m.y.z, m.z.z)); }              // FMatrix3x3.h:109-121
inline float determinant(const M3 &m) {                                                                                                 // FMatrix3x3.h:82-89
	return m.x.x * m.y.y * m.z.z + m.z.x * m.x.y * m.y.z + m.y.x * m.z.y * m.x.z - m.x.x * m.z.y * m.y.z - m.y.x * m.x.y * m.z.z - m.z.x * m.y.y * m.x.z;
}
inline M3 inverse(const M3 &m) {                                                                                                        // FMatrix3x3.h:91-107
	const float invDet = 1.0f / determinant(m);
	M3 r;
	r.x.x = invDet * (m.y.y * m.z.z - m.y.z * m.z.y);
	r.x.y = -invDet * (m.x.y * m.z.z - m.x.z * m.z.y);
	r.x.z = invDet * (m.x.y * m.y.z - m.x.z * m.y.y);
	r.y.x = -invDet * (m.y.x * m.z.z - m.y.z * m.z.x);
	r.y.y = invDet * (m.x.x * m.z.z - m.x.z * m.z.x);
	r.y.z = -invDet * (m.x.x * m.y.z - m.x.z * m.y.x);
	r.z.x = invDet * (m.y.x * m.z.y - m.y.y * m.z.x);
	r.z.y = -invDet * (m.x.x * m.z.y - m.x.y * m.z.x);
	r.z.z = invDet * (m.x.x * m.y.y - m.x.y * m.y.x);
	return r;
}
inline M3 make_axis_system(F3 forward, F3 up) {                                                                                         // FMatrix3x3.h:46-53
	M3 r;
	const F3 forwardNormalized = normalize(forward);
	r.z = forwardNormalized;
	r.x = normalize(cross(normalize(up), forwardNormalized));
	r.y = normalize(cross(forwardNormalized, r.x));
	return r;
}
__host__ __device__ inline F3 transform_point(const T3 &t, F3 p) { return add(transform(t.m, p), t.position); }                          // Transform3D.h:41-43
inline T3 t3(F3 position, const M3 &m) { T3 r; r.position = position; r.m = m; return r; }
inline T3 t3(const dfpsr_transform3d &t) { return t3(f3(t.position), m3(f3(t.xAxis), f3(t.yAxis), f3(t.zAxis))); }
inline dfpsr_transform3d pod(const T3 &t) {
	dfpsr_transform3d r;
	r.position[0] = t.position.x; r.position[1] = t.position.y; r.position[2] = t.position.z;
	r.xAxis[0] = t.m.x.x; r.xAxis[1] = t.m.x.y; r.xAxis[2] = t.m.x.z;
	r.yAxis[0] = t.m.y.x; r.yAxis[1] = t.m.y.y; r.yAxis[2] = t.m.y.z;
	r.zAxis[0] = t.m.z.x; r.zAxis[1] = t.m.z.y; r.zAxis[2] = t.m.z.z;
	return r;
}
inline T3 mul(const T3 &left, const T3 &right) { return t3(transform_point(right, left.position), mul(left.m, right.m)); }               // Transform3D.h:56-58

// ---- integer rectangles (ref: math/IRect.h): left, top, width, height; right and bottom are exclusive
struct Rect {
	int32_t l = 0, t = 0, w = 0, h = 0;
	Rect() {}
	Rect(int32_t l, int32_t t, int32_t w, int32_t h) : l(l), t(t), w(w), h(h) {}
	int32_t right() const { return l + w; }
	int32_t bottom() const { return t + h; }
	bool has_area() const { return w > 0 && h > 0; }
	Rect expanded(int32_t units) const { return Rect(l - units, t - units, w + units * 2, h + units * 2); }
	static bool overlaps(const Rect &a, const Rect &b) { return a.l < b.right() && a.right() > b.l && a.t < b.bottom() && a.bottom() > b.t; }
	static bool touches(const Rect &a, const Rect &b) { return a.l <= b.right() && a.right() >= b.l && a.t <= b.bottom() && a.bottom() >= b.t; }
	static Rect cut(const Rect &a, const Rect &b) {
		if (!overlaps(a, b)) { return Rect(); }
		const int32_t l = a.l > b.l ? a.l : b.l, t = a.t > b.t ? a.t : b.t;
		const int32_t r = a.right() < b.right() ? a.right() : b.right(), bo = a.bottom() < b.bottom() ? a.bottom() : b.bottom();
		return Rect(l, t, r - l, bo - t);
	}
	static Rect merge(const Rect &a, const Rect &b) {
		const int32_t l = a.l < b.l ? a.l : b.l, t = a.t < b.t ? a.t : b.t;
		const int32_t r = a.right() > b.right() ? a.right() : b.right(), bo = a.bottom() > b.bottom() ? a.bottom() : b.bottom();
		return Rect(l, t, r - l, bo - t);
	}
};

// ref: implementation/math/scalar.h:31-51 signedModulo / roundDown
inline int64_t signed_modulo(int64_t a, int64_t b) { return a >= 0 ? a % b : (b - (-a % b)) % b; }
inline int64_t round_down(int64_t size, int64_t alignment) { return size - signed_modulo(size, alignment); }

} // namespace sw
} // namespace dfpsr
