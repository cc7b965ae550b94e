// TEST INFRASTRUCTURE (oracle/ref/Makefile): the subset of GLSL that the reference's math.glsl, material.glsl and
// light_sampling.glsl use, as C++, so that those files can be compiled BY g++ FROM WHERE THEY LIE (after the purely syntactic
// rewrite of glsl_to_cpp.py: out / inout parameters -> references, swizzles -> member calls, constructors whose arguments draw
// random numbers -> braces, which C++ evaluates left to right as GLSL does) and called by tests/test_cpu_ref_pins.py.
// Compiled with -fsingle-precision-constant -ffp-contract=off: GLSL's literals are floats and nothing is fused.
// This header is ours; nothing of the reference is copied here.
#pragma once
#include <cmath>
#include <cstdint>
#undef assert

typedef unsigned int uint;

struct vec2 {
	float x, y;
	vec2() : x(0), y(0) {}
	explicit vec2(float s) : x(s), y(s) {}
	vec2(float x_, float y_) : x(x_), y(y_) {}
	explicit vec2(const struct vec3& v);
};
struct vec3 {
	float x, y, z;
	vec3() : x(0), y(0), z(0) {}
	explicit vec3(float s) : x(s), y(s), z(s) {}
	vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
	vec3(vec2 a, float z_) : x(a.x), y(a.y), z(z_) {}
	vec2 xy() const { return vec2(x, y); }
	vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
	vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
	vec3& operator+=(vec3 b) { x += b.x; y += b.y; z += b.z; return *this; }
};
inline vec2::vec2(const vec3& v) : x(v.x), y(v.y) {}
struct vec4 {
	float x, y, z, w;
	vec4() : x(0), y(0), z(0), w(0) {}
	vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
	vec4(vec2 a, vec2 b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
	vec2 zw() const { return vec2(z, w); }
};
struct uvec2 { uint x, y; };

inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator-(vec2 a, float s) { return vec2(a.x - s, a.y - s); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }

inline float abs(float x) { return std::fabs(x); }
inline float max(float a, float b) { return b > a ? b : a; }   // GLSL: y if x < y
inline float min(float a, float b) { return b < a ? b : a; }
inline float sqrt(float x) { return std::sqrt(x); }
inline float cos(float x) { return std::cos(x); }
inline float sin(float x) { return std::sin(x); }
inline float asin(float x) { return std::asin(x); }
inline float atan(float y, float x) { return std::atan2(y, x); }
inline bool isnan(float x) { return x != x; }
inline bool isinf(float x) { return std::isinf(x); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }   // the GLSL definition, spelled out
inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, vec3 t) { return vec3(mix(a.x, b.x, t.x), mix(a.y, b.y, t.y), mix(a.z, b.z, t.z)); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec3 clamp(vec3 v, vec3 lo, vec3 hi) { return vec3(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z)); }
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(vec2 a) { return sqrt(dot(a, a)); }
inline float length(vec3 a) { return sqrt(dot(a, a)); }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
inline vec3 normalize(vec3 a) { return a / length(a); }
inline vec3 reflect(vec3 i, vec3 n) { return i - n * (2.0f * dot(n, i)); }

// column-major 3x3, as GLSL's mat3(c0, c1, c2)
struct mat3 {
	vec3 c[3];
	mat3(vec3 a, vec3 b, vec3 d) { c[0] = a; c[1] = b; c[2] = d; }
};
inline vec3 operator*(const mat3& m, vec3 v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
inline mat3 inverse(const mat3& m) {   // adjugate / determinant
	const vec3 a = m.c[0], b = m.c[1], c = m.c[2];
	const vec3 r0 = cross(b, c), r1 = cross(c, a), r2 = cross(a, b);
	const float invDet = 1.0f / dot(r2, c);
	return mat3(vec3(r0.x, r1.x, r2.x) * invDet, vec3(r0.y, r1.y, r2.y) * invDet, vec3(r0.z, r1.z, r2.z) * invDet);
}

// layouts.glsl:6-25 (RESTIR_PT_MATERIAL), :74-88 — the shader-side structs, byte-identical to the C ABI's
struct Material { vec3 baseColor; uint type; uint textureIdx; float metallic; float roughness; float ior; };
struct TriangleLight { vec3 v0; float nx; vec3 v1; float ny; vec3 v2; float nz; vec3 radiance; float area; };
struct LightSampleTableElement { float prob; uint failId; };
static_assert(sizeof(Material) == 32 && sizeof(TriangleLight) == 64 && sizeof(LightSampleTableElement) == 8, "std430 layouts");
// the two storage buffers light_sampling.glsl reads (layouts.glsl:176-177)
static const TriangleLight* uTriangleLights = nullptr;
static const LightSampleTableElement* uLightSampleTable = nullptr;
