// TEST INFRASTRUCTURE (oracle/ref/Makefile): the subset of GLSL 4.60 (+ GL_EXT_ray_query) that the reference's shaders use, as
// C++, so that the shader sources can be compiled BY g++ FROM WHERE THEY LIE under /root/reference (after the purely syntactic
// rewrite of glsl_to_cpp.py) and executed on the CPU by tests/test_cpu_ref_pins.py and tests/test_cpu_ref_shaders.py.
// This header is ours; nothing of the reference is copied here.
//
// Two builds of the built-in function library:
//   default                     plain IEEE operations and libm (sin / cos / tan / atan / asin), nothing fused: "a GPU with exact built-ins"
//   -DGLSL_BUILTINS_CONTRACT    the built-ins whose precision and evaluation order GLSL leaves to the implementation (dot, cross,
//                               length, normalize, distance, mix, reflect, matrix * vector, inverse of an orthonormal mat3,
//                               vector / scalar, sin, cos, tan) are evaluated as the numeric contract of DESIGN.md §2 prescribes
//                               (oracle_math.h / rt_math.cuh): then every remaining difference between the reference's text and the
//                               oracle is a difference in the ALGORITHM, and whole frames can be compared bit for bit.
// Everything the shader text itself spells out (a * b + c, a / b, comparisons, integer work) is compiled as written, with
// -ffp-contract=off -fsingle-precision-constant (GLSL literals are floats; nothing is fused that the contract does not fuse).
// GLSL locals are not initialised by the language; here every vector / struct starts at zero (DESIGN.md "defined behaviours" 1).
#pragma once
// (every free function here is `static`: this header is compiled in two modes into one library, and functions with external
// linkage and different bodies would be merged by the linker)
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#undef assert

typedef unsigned int uint;

#ifdef GLSL_BUILTINS_CONTRACT
#include "../oracle_math.h"   // orc::sincos_, orc::fma_ : the shared definition of the contract's transcendental functions
#define GLSL_FMA(a, b, c) __builtin_fmaf(a, b, c)
#endif

// ---- vectors ----------------------------------------------------------------------------------------------------------------
struct vec3; struct vec4; struct uvec2; struct ivec2;
struct vec2 {
	float x, y;
	vec2() : x(0), y(0) {}
	explicit vec2(float s) : x(s), y(s) {}
	vec2(float x_, float y_) : x(x_), y(y_) {}
	explicit vec2(const vec3& v);
	explicit vec2(const uvec2& v);
	explicit vec2(const ivec2& v);
	vec2 xy() const { return *this; }
	float& operator[](int i) { return (&x)[i]; }
};
struct vec3 {
	float x, y, z;
	vec3() : x(0), y(0), z(0) {}
	explicit vec3(float s) : x(s), y(s), z(s) {}
	vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
	vec3(vec2 a, float z_) : x(a.x), y(a.y), z(z_) {}
	vec3(float x_, vec2 a) : x(x_), y(a.x), z(a.y) {}
	explicit vec3(const vec4& v);
	vec2 xy() const { return vec2(x, y); }
	vec2 yz() const { return vec2(y, z); }
	vec2& yz() { return *reinterpret_cast<vec2*>(&y); }   // contiguous: also usable as an `out` argument (light_sampling.glsl:58)
	vec3 xyz() const { return *this; }
	vec3 rgb() const { return *this; }
	float& operator[](int i) { return (&x)[i]; }
};
struct vec4 {
	float x, y, z, w;
	vec4() : x(0), y(0), z(0), w(0) {}
	explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
	vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
	vec4(vec2 a, vec2 b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
	vec4(vec3 a, float w_) : x(a.x), y(a.y), z(a.z), w(w_) {}
	vec4(vec2 a, float z_, float w_) : x(a.x), y(a.y), z(z_), w(w_) {}
	vec2 xy() const { return vec2(x, y); }
	vec2 zw() const { return vec2(z, w); }
	vec3 xyz() const { return vec3(x, y, z); }
	vec3 rgb() const { return vec3(x, y, z); }
	vec3 yzw() const { return vec3(y, z, w); }
	float& operator[](int i) { return (&x)[i]; }
};
struct ivec2 {
	int x, y;
	ivec2() : x(0), y(0) {}
	explicit ivec2(int s) : x(s), y(s) {}
	ivec2(int x_, int y_) : x(x_), y(y_) {}
	explicit ivec2(const vec2& v) : x(int(v.x)), y(int(v.y)) {}
	explicit ivec2(const uvec2& v);
	ivec2 xy() const { return *this; }
};
struct uvec2 {
	uint x, y;
	uvec2() : x(0), y(0) {}
	explicit uvec2(uint s) : x(s), y(s) {}
	uvec2(uint x_, uint y_) : x(x_), y(y_) {}
	explicit uvec2(const vec2& v) : x(uint(v.x)), y(uint(v.y)) {}
	explicit uvec2(const ivec2& v) : x(uint(v.x)), y(uint(v.y)) {}
	uvec2 xy() const { return *this; }
};
struct uvec3 {
	uint x, y, z;
	uvec3() : x(0), y(0), z(0) {}
	uvec3(uint x_, uint y_, uint z_) : x(x_), y(y_), z(z_) {}
	uvec2 xy() const { return uvec2(x, y); }
};
struct uvec4 {
	uint x, y, z, w;
	uvec4() : x(0), y(0), z(0), w(0) {}
	uvec4(uint x_, uint y_, uint z_, uint w_) : x(x_), y(y_), z(z_), w(w_) {}
	uvec2 xy() const { return uvec2(x, y); }
};
inline vec2::vec2(const vec3& v) : x(v.x), y(v.y) {}
inline vec2::vec2(const uvec2& v) : x(float(v.x)), y(float(v.y)) {}
inline vec2::vec2(const ivec2& v) : x(float(v.x)), y(float(v.y)) {}
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}
inline ivec2::ivec2(const uvec2& v) : x(int(v.x)), y(int(v.y)) {}

// component-wise arithmetic exactly as written: vector (op) vector, vector (op) scalar, scalar (op) vector
#define GLSL_OP2(OP) \
	static inline vec2 operator OP(vec2 a, vec2 b) { return vec2(a.x OP b.x, a.y OP b.y); } \
	static inline vec2 operator OP(vec2 a, float s) { return vec2(a.x OP s, a.y OP s); } \
	static inline vec2 operator OP(float s, vec2 a) { return vec2(s OP a.x, s OP a.y); } \
	static inline vec2& operator OP##=(vec2& a, vec2 b) { a = a OP b; return a; } \
	static inline vec2& operator OP##=(vec2& a, float s) { a = a OP s; return a; }
#define GLSL_OP3(OP) \
	static inline vec3 operator OP(vec3 a, vec3 b) { return vec3(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
	static inline vec3 operator OP(vec3 a, float s) { return vec3(a.x OP s, a.y OP s, a.z OP s); } \
	static inline vec3 operator OP(float s, vec3 a) { return vec3(s OP a.x, s OP a.y, s OP a.z); } \
	static inline vec3& operator OP##=(vec3& a, vec3 b) { a = a OP b; return a; } \
	static inline vec3& operator OP##=(vec3& a, float s) { a = a OP s; return a; }
#define GLSL_OP4(OP) \
	static inline vec4 operator OP(vec4 a, vec4 b) { return vec4(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
	static inline vec4 operator OP(vec4 a, float s) { return vec4(a.x OP s, a.y OP s, a.z OP s, a.w OP s); } \
	static inline vec4 operator OP(float s, vec4 a) { return vec4(s OP a.x, s OP a.y, s OP a.z, s OP a.w); }
GLSL_OP2(+) GLSL_OP2(-) GLSL_OP2(*) GLSL_OP2(/)
GLSL_OP3(+) GLSL_OP3(-) GLSL_OP3(*)
GLSL_OP4(+) GLSL_OP4(-) GLSL_OP4(*) GLSL_OP4(/)
static inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline vec3 operator/(float s, vec3 a) { return vec3(s / a.x, s / a.y, s / a.z); }
#ifdef GLSL_BUILTINS_CONTRACT
// vector / scalar: one correctly rounded reciprocal, three multiplies — how a GPU compiles it (oracle_math.h operator/)
static inline vec3 operator/(vec3 a, float s) { const float r = 1.0f / s; return vec3(a.x * r, a.y * r, a.z * r); }
#else
static inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
#endif
static inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
static inline vec3& operator/=(vec3& a, vec3 b) { a = a / b; return a; }
static inline vec2 operator-(vec2 a) { return vec2(-a.x, -a.y); }
static inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
static inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
static inline ivec2 operator-(ivec2 a, ivec2 b) { return ivec2(a.x - b.x, a.y - b.y); }
static inline uvec2 operator+(uvec2 a, uvec2 b) { return uvec2(a.x + b.x, a.y + b.y); }
static inline bool operator==(uvec2 a, uvec2 b) { return a.x == b.x && a.y == b.y; }
static inline bool operator!=(uvec2 a, uvec2 b) { return !(a == b); }
static inline bool operator==(ivec2 a, ivec2 b) { return a.x == b.x && a.y == b.y; }
static inline bool operator!=(ivec2 a, ivec2 b) { return !(a == b); }

// ---- scalar built-ins -------------------------------------------------------------------------------------------------------
static inline float abs(float x) { return std::fabs(x); }
static inline float max(float a, float b) { return b > a ? b : a; }   // GLSL: y if x < y
static inline float min(float a, float b) { return b < a ? b : a; }
static inline float max(float a, int b) { return max(a, float(b)); }
static inline float max(int a, float b) { return max(float(a), b); }
static inline float min(float a, int b) { return min(a, float(b)); }
static inline float min(int a, float b) { return min(float(a), b); }
static inline int max(int a, int b) { return b > a ? b : a; }
static inline int min(int a, int b) { return b < a ? b : a; }
static inline uint max(uint a, uint b) { return b > a ? b : a; }
static inline uint min(uint a, uint b) { return b < a ? b : a; }
static inline float sqrt(float x) { return std::sqrt(x); }
static inline float floor(float x) { return std::floor(x); }
static inline float fract(float x) { return x - std::floor(x); }
static inline float pow(float x, float y) { return std::pow(x, y); }
static inline float exp(float x) { return std::exp(x); }
static inline float log(float x) { return std::log(x); }
static inline float asin(float x) { return std::asin(x); }
static inline float acos(float x) { return std::acos(x); }
static inline float atan(float y, float x) { return std::atan2(y, x); }
static inline float radians(float deg) { return deg * 0.017453292519943295f; }
#ifdef GLSL_BUILTINS_CONTRACT
static inline float cos(float x) { float s, c; orc::sincos_(x, s, c); return c; }
static inline float sin(float x) { float s, c; orc::sincos_(x, s, c); return s; }
static inline float tan(float x) { float s, c; orc::sincos_(x, s, c); return s / c; }
static inline float mix(float a, float b, float t) { return GLSL_FMA(b, t, a * (1.0f - t)); }
#else
static inline float cos(float x) { return std::cos(x); }
static inline float sin(float x) { return std::sin(x); }
static inline float tan(float x) { return std::tan(x); }
static inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }   // the GLSL definition, spelled out
#endif
static inline bool isnan(float x) { return x != x; }
static inline bool isinf(float x) { return std::isinf(x); }
static inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
static inline float clamp(float x, int lo, int hi) { return clamp(x, float(lo), float(hi)); }
static inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
static inline uint floatBitsToUint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
static inline float uintBitsToFloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uvec2 floatBitsToUint(vec2 v) { return uvec2(floatBitsToUint(v.x), floatBitsToUint(v.y)); }
static inline vec2 uintBitsToFloat(uvec2 v) { return vec2(uintBitsToFloat(v.x), uintBitsToFloat(v.y)); }
template <typename T> static inline T nonuniformEXT(T x) { return x; }

// ---- vector built-ins -------------------------------------------------------------------------------------------------------
static inline vec2 abs(vec2 a) { return vec2(abs(a.x), abs(a.y)); }
static inline vec3 abs(vec3 a) { return vec3(abs(a.x), abs(a.y), abs(a.z)); }
static inline vec3 max(vec3 a, vec3 b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
static inline vec3 min(vec3 a, vec3 b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
static inline vec3 max(vec3 a, float b) { return vec3(max(a.x, b), max(a.y, b), max(a.z, b)); }
static inline vec3 min(vec3 a, float b) { return vec3(min(a.x, b), min(a.y, b), min(a.z, b)); }
static inline vec3 mix(vec3 a, vec3 b, float t) { return vec3(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)); }
static inline vec3 mix(vec3 a, vec3 b, vec3 t) { return vec3(mix(a.x, b.x, t.x), mix(a.y, b.y, t.y), mix(a.z, b.z, t.z)); }
static inline vec3 clamp(vec3 v, vec3 lo, vec3 hi) { return vec3(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z)); }
static inline vec3 clamp(vec3 v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
static inline vec3 pow(vec3 a, vec3 b) { return vec3(pow(a.x, b.x), pow(a.y, b.y), pow(a.z, b.z)); }
static inline vec3 sqrt(vec3 a) { return vec3(sqrt(a.x), sqrt(a.y), sqrt(a.z)); }
static inline vec3 floor(vec3 a) { return vec3(floor(a.x), floor(a.y), floor(a.z)); }
static inline vec3 fract(vec3 a) { return vec3(fract(a.x), fract(a.y), fract(a.z)); }
#ifdef GLSL_BUILTINS_CONTRACT
static inline float dot(vec2 a, vec2 b) { return GLSL_FMA(a.y, b.y, a.x * b.x); }
static inline float dot(vec3 a, vec3 b) { return GLSL_FMA(a.z, b.z, GLSL_FMA(a.y, b.y, a.x * b.x)); }
static inline vec3 cross(vec3 a, vec3 b) {
	return vec3(GLSL_FMA(a.y, b.z, -(a.z * b.y)), GLSL_FMA(a.z, b.x, -(a.x * b.z)), GLSL_FMA(a.x, b.y, -(a.y * b.x)));
}
static inline float length(vec2 a) { return sqrt(dot(a, a)); }
static inline float length(vec3 a) { return sqrt(dot(a, a)); }
static inline vec3 normalize(vec3 a) { return a * (1.0f / length(a)); }
static inline vec3 reflect(vec3 i, vec3 n) { const float k = 2.0f * dot(n, i); return i - n * k; }
#else
static inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
static inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
static inline float length(vec2 a) { return sqrt(dot(a, a)); }
static inline float length(vec3 a) { return sqrt(dot(a, a)); }
static inline vec3 normalize(vec3 a) { return a / length(a); }
static inline vec3 reflect(vec3 i, vec3 n) { return i - n * (2.0f * dot(n, i)); }
#endif
static inline float distance(vec3 a, vec3 b) { return length(a - b); }
static inline float distance(vec2 a, vec2 b) { return length(a - b); }

// ---- matrices (column-major, as GLSL) ---------------------------------------------------------------------------------------
struct mat3 {
	vec3 c[3];
	mat3() {}
	mat3(vec3 a, vec3 b, vec3 d) { c[0] = a; c[1] = b; c[2] = d; }
};
struct mat4 {
	vec4 c[4];
	mat4() {}
};
#ifdef GLSL_BUILTINS_CONTRACT
static inline vec3 operator*(const mat3& m, vec3 v) {
	return vec3(GLSL_FMA(m.c[2].x, v.z, GLSL_FMA(m.c[1].x, v.y, m.c[0].x * v.x)),
	            GLSL_FMA(m.c[2].y, v.z, GLSL_FMA(m.c[1].y, v.y, m.c[0].y * v.x)),
	            GLSL_FMA(m.c[2].z, v.z, GLSL_FMA(m.c[1].z, v.y, m.c[0].z * v.x)));
}
// only ever applied to the orthonormal frame of matLocalToWorld: the transpose (GLSL leaves inverse()'s precision undefined)
static inline mat3 inverse(const mat3& m) {
	return mat3(vec3(m.c[0].x, m.c[1].x, m.c[2].x), vec3(m.c[0].y, m.c[1].y, m.c[2].y), vec3(m.c[0].z, m.c[1].z, m.c[2].z));
}
static inline vec4 operator*(const mat4& m, vec4 v) {
	return vec4(GLSL_FMA(m.c[2].x, v.z, GLSL_FMA(m.c[1].x, v.y, GLSL_FMA(m.c[0].x, v.x, m.c[3].x * v.w))),
	            GLSL_FMA(m.c[2].y, v.z, GLSL_FMA(m.c[1].y, v.y, GLSL_FMA(m.c[0].y, v.x, m.c[3].y * v.w))),
	            GLSL_FMA(m.c[2].z, v.z, GLSL_FMA(m.c[1].z, v.y, GLSL_FMA(m.c[0].z, v.x, m.c[3].z * v.w))),
	            GLSL_FMA(m.c[2].w, v.z, GLSL_FMA(m.c[1].w, v.y, GLSL_FMA(m.c[0].w, v.x, m.c[3].w * v.w))));
}
#else
static inline vec3 operator*(const mat3& m, vec3 v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
static inline mat3 inverse(const mat3& m) {   // adjugate / determinant
	const vec3 a = m.c[0], b = m.c[1], c = m.c[2];
	const vec3 r0 = cross(b, c), r1 = cross(c, a), r2 = cross(a, b);
	const float invDet = 1.0f / dot(r2, c);
	return mat3(vec3(r0.x, r1.x, r2.x) * invDet, vec3(r0.y, r1.y, r2.y) * invDet, vec3(r0.z, r1.z, r2.z) * invDet);
}
static inline vec4 operator*(const mat4& m, vec4 v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w; }
#endif

// ---- packing (gbuffer_util.glsl) --------------------------------------------------------------------------------------------
static inline uint packUnorm4x8(vec4 v) {
	auto q = [](float c) { return uint(std::floor(clamp(c, 0.0f, 1.0f) * 255.0f + 0.5f)); };   // round(clamp(c, 0, 1) * 255)
	return q(v.x) | (q(v.y) << 8) | (q(v.z) << 16) | (q(v.w) << 24);
}
static inline vec4 unpackUnorm4x8(uint p) {
	return vec4(float(p & 0xffu) / 255.0f, float((p >> 8) & 0xffu) / 255.0f, float((p >> 16) & 0xffu) / 255.0f, float(p >> 24) / 255.0f);
}
static inline uint packSnorm2x16(vec2 v) {
	auto q = [](float c) { return uint(int(std::nearbyint(clamp(c, -1.0f, 1.0f) * 32767.0f))) & 0xffffu; };
	return q(v.x) | (q(v.y) << 16);
}
static inline vec2 unpackSnorm2x16(uint p) {
	auto u = [](uint h) { return clamp(float(int16_t(h)) / 32767.0f, -1.0f, 1.0f); };
	return vec2(u(p & 0xffffu), u(p >> 16));
}

// a 4-byte bool for structs that live in buffers (std430 / push constants)
struct gbool {
	uint v;
	gbool() : v(0) {}
	gbool(bool b) : v(b ? 1u : 0u) {}
	operator bool() const { return v != 0; }
};

#ifdef GLSL_COMPAT_RESOURCES
// ---- resources: what the Vulkan driver supplies (images, samplers, the acceleration structure) ------------------------------
// Filtering, addressing and ray / triangle intersection are the DRIVER's, not the reference's: they are delegated to call-backs
// (the test passes the oracle's definitions of them — DESIGN.md "defined behaviours" 4-6), so what is compared is the shader text.
struct RefDriver {
	void* user;
	// closest hit: returns 0 on a miss; bary = barycentrics (u, v), instance = custom index, primitive = triangle index
	int (*traceClosest)(void* user, const float* o, float tmin, const float* d, float tmax, float* bary, uint* instance, uint* primitive);
	int (*traceAny)(void* user, const float* o, float tmin, const float* d, float tmax);
	uint (*countCandidates)(void* user, const float* o, const float* d);
	void (*sampleTexture)(void* user, uint texture, float u, float v, float* rgb);
	// texture() on a G-buffer float image (`which`: 0 = this frame's depth-normal, 1 = the previous frame's)
	void (*sampleDepthNormal)(void* user, int which, float u, float v, float* out4);
};
extern thread_local const RefDriver* glslDriver;

// a storage buffer / descriptor array with Vulkan's robust-access behaviour: a read past the end yields zeros, a write is dropped
// (the shaders do read out of range and then discard the value: di_reservoir.glsl loads the surface of a sample BEFORE it asks
// whether the sample is valid, instanceIdx 0xffffffff)
template <typename T> struct glsl_buffer {
	T* p = nullptr;
	size_t n = 0;
	T& operator[](size_t i) const {
		if (i < n) return p[i];
		static thread_local T outOfRange;
		outOfRange = T();
		return outOfRange;
	}
};

struct sampler2D {          // a material texture (index) or a G-buffer float image (texels)
	int kind = 0;           // 0: material texture `index`; 1: depth-normal (`which`); 2: plain float image for texelFetch only
	uint index = 0;
	int which = 0;
	const float* texels = nullptr;
	int channels = 4;
	uint width = 0, height = 0;
};
struct usampler2D { const uint* texels = nullptr; int channels = 2; uint width = 0, height = 0; };
struct image2D { float* texels = nullptr; uint width = 0, height = 0; };
struct accelerationStructureEXT { int unused = 0; };

static inline vec4 texture(const sampler2D& s, vec2 uv) {
	float o[4] = { 0, 0, 0, 1 };
	if (s.kind == 0) glslDriver->sampleTexture(glslDriver->user, s.index, uv.x, uv.y, o);
	else glslDriver->sampleDepthNormal(glslDriver->user, s.which, uv.x, uv.y, o);
	return vec4(o[0], o[1], o[2], o[3]);
}
// out-of-range coordinates: the oracle's definition (clamped to the edge; Vulkan's robust access would return 0), DESIGN.md
static inline size_t glslTexel(int x, int y, uint w, uint h) {
	x = x < 0 ? 0 : (x > int(w) - 1 ? int(w) - 1 : x);
	y = y < 0 ? 0 : (y > int(h) - 1 ? int(h) - 1 : y);
	return size_t(y) * w + size_t(x);
}
static inline vec4 texelFetch(const sampler2D& s, ivec2 p, int) {
	const float* t = s.texels + glslTexel(p.x, p.y, s.width, s.height) * s.channels;
	return vec4(t[0], s.channels > 1 ? t[1] : 0.0f, s.channels > 2 ? t[2] : 0.0f, s.channels > 3 ? t[3] : 1.0f);
}
static inline uvec4 texelFetch(const usampler2D& s, ivec2 p, int) {
	const uint* t = s.texels + glslTexel(p.x, p.y, s.width, s.height) * s.channels;
	return uvec4(t[0], s.channels > 1 ? t[1] : 0u, s.channels > 2 ? t[2] : 0u, s.channels > 3 ? t[3] : 0u);
}
static inline vec4 imageLoad(const image2D& im, ivec2 p) {
	const float* t = im.texels + (size_t(p.y) * im.width + size_t(p.x)) * 4;
	return vec4(t[0], t[1], t[2], t[3]);
}
static inline void imageStore(const image2D& im, ivec2 p, vec4 v) {
	float* t = im.texels + (size_t(p.y) * im.width + size_t(p.x)) * 4;
	t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
}

// ---- GL_EXT_ray_query over opaque geometry: rayQueryProceedEXT never surfaces a candidate, the committed hit is the closest ---
const uint gl_RayFlagsNoneEXT = 0u, gl_RayFlagsOpaqueEXT = 1u, gl_RayFlagsTerminateOnFirstHitEXT = 4u, gl_RayFlagsSkipClosestHitShaderEXT = 8u;
const uint gl_RayQueryCommittedIntersectionNoneEXT = 0u, gl_RayQueryCommittedIntersectionTriangleEXT = 1u;
const uint gl_RayQueryCandidateIntersectionTriangleEXT = 0u;
struct rayQueryEXT {
	vec3 o, d;
	float tmin = 0, tmax = 0;
	uint flags = 0;
	bool done = false, hit = false;
	vec2 bary;
	uint instance = 0, primitive = 0;
	uint candidates = 0;      // non-opaque queries (debugVisualizeAS): candidate triangles still to surface
};
static inline void rayQueryInitializeEXT(rayQueryEXT& q, const accelerationStructureEXT&, uint flags, uint, vec3 o, float tmin, vec3 d, float tmax) {
	q = rayQueryEXT();
	q.o = o; q.d = d; q.tmin = tmin; q.tmax = tmax; q.flags = flags;
}
static inline bool rayQueryProceedEXT(rayQueryEXT& q) {
	if (!q.done) {
		q.done = true;
		if (q.flags & gl_RayFlagsTerminateOnFirstHitEXT) q.hit = glslDriver->traceAny(glslDriver->user, &q.o.x, q.tmin, &q.d.x, q.tmax) != 0;
		else q.hit = glslDriver->traceClosest(glslDriver->user, &q.o.x, q.tmin, &q.d.x, q.tmax, &q.bary.x, &q.instance, &q.primitive) != 0;
		if (!(q.flags & gl_RayFlagsOpaqueEXT)) q.candidates = glslDriver->countCandidates(glslDriver->user, &q.o.x, &q.d.x);
	}
	if (q.candidates > 0) { q.candidates--; return true; }
	return false;
}
static inline uint rayQueryGetIntersectionTypeEXT(const rayQueryEXT& q, bool committed) {
	if (!committed) return gl_RayQueryCandidateIntersectionTriangleEXT;
	return q.hit ? gl_RayQueryCommittedIntersectionTriangleEXT : gl_RayQueryCommittedIntersectionNoneEXT;
}
static inline vec2 rayQueryGetIntersectionBarycentricsEXT(const rayQueryEXT& q, bool) { return q.bary; }
static inline uint rayQueryGetIntersectionInstanceCustomIndexEXT(const rayQueryEXT& q, bool) { return q.instance; }
static inline uint rayQueryGetIntersectionPrimitiveIndexEXT(const rayQueryEXT& q, bool) { return q.primitive; }

extern thread_local uvec3 gl_GlobalInvocationID;
#endif  // GLSL_COMPAT_RESOURCES

#ifdef GLSL_COMPAT_LIBRARY_STRUCTS
// layouts.glsl:6-25 (RESTIR_PT_MATERIAL), :74-88 — the shader-side structs for the library-only build (ref_glsl.cpp), byte-identical
// to the C ABI's; the whole-shader build (ref_shaders.cpp) takes them from layouts.glsl itself
struct Material { vec3 baseColor; uint type; uint textureIdx; float metallic; float roughness; float ior; };
struct TriangleLight { vec3 v0; float nx; vec3 v1; float ny; vec3 v2; float nz; vec3 radiance; float area; };
struct LightSampleTableElement { float prob; uint failId; };
static_assert(sizeof(Material) == 32 && sizeof(TriangleLight) == 64 && sizeof(LightSampleTableElement) == 8, "std430 layouts");
// the two storage buffers light_sampling.glsl reads (layouts.glsl:176-177)
static const TriangleLight* uTriangleLights = nullptr;
static const LightSampleTableElement* uLightSampleTable = nullptr;
#endif
