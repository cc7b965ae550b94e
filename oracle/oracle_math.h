// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference shaders' arithmetic
// (reference src/shader/math.glsl and the GLSL built-ins it relies on).  Never linked into the product.
// PARITY PINNED against the reference's own shaders compiled for the CPU (oracle/ref, tests/test_cpu_ref_shaders.py):
// the reference ships no tests / golden vectors, so its shader text itself, compiled by g++, is the pin.
//
// Numeric contract shared (by specification, not by code) with the CUDA kernels, so that both produce
// bit-identical fp32 results: IEEE single, round-to-nearest, no implicit contraction (-ffp-contract=off here,
// -fmad=false there); fused multiply-add only where `fma` is spelled out below; divisions and square roots
// correctly rounded; comparisons written so NaN behaves the same; sin/cos from the polynomial below instead
// of libm.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

inline float fma_(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
inline float min_(float a, float b) { return a < b ? a : b; }
inline float max_(float a, float b) { return a > b ? a : b; }
inline float clamp_(float x, float lo, float hi) { return min_(max_(x, lo), hi); }
inline float abs_(float x) { return std::fabs(x); }
inline bool isnan_(float x) { return x != x; }

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };
struct uvec2 { uint32_t x, y; };

inline vec3 V3(float s) { return { s, s, s }; }
inline vec3 V3(float x, float y, float z) { return { x, y, z }; }
inline vec3 V3(const float* p) { return { p[0], p[1], p[2] }; }

inline vec3 operator+(vec3 a, vec3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline vec3 operator-(vec3 a, vec3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline vec3 operator-(vec3 a) { return { -a.x, -a.y, -a.z }; }
inline vec3 operator*(vec3 a, vec3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline vec3 operator*(vec3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
// vec3 / float: one correctly rounded reciprocal, three multiplies (the form a GPU compiler gives GLSL's vector-by-scalar division;
// part of the numeric contract shared with rt_math.cuh)
inline vec3 operator/(vec3 a, float s) { const float r = 1.0f / s; return { a.x * r, a.y * r, a.z * r }; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }

inline vec2 operator+(vec2 a, vec2 b) { return { a.x + b.x, a.y + b.y }; }
inline vec2 operator*(vec2 a, float s) { return { a.x * s, a.y * s }; }

inline float dot(vec3 a, vec3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
inline float dot(vec2 a, vec2 b) { return fma_(a.y, b.y, a.x * b.x); }
inline vec3 cross(vec3 a, vec3 b) {
	return { fma_(a.y, b.z, -(a.z * b.y)), fma_(a.z, b.x, -(a.x * b.z)), fma_(a.x, b.y, -(a.y * b.x)) };
}
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline float length(vec2 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a * (1.0f / length(a)); }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
// GLSL mix
inline float mix(float a, float b, float t) { return fma_(b, t, a * (1.0f - t)); }
inline vec3 mix(vec3 a, vec3 b, float t) { return { mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t) }; }
inline vec3 mix(vec3 a, vec3 b, vec3 t) { return { mix(a.x, b.x, t.x), mix(a.y, b.y, t.y), mix(a.z, b.z, t.z) }; }
// GLSL reflect(I, N) = I - 2 dot(N, I) N
inline vec3 reflect(vec3 I, vec3 N) { float k = 2.0f * dot(N, I); return I - N * k; }
// a*w.x + b*w.y + c*w.z (barycentric interpolation), evaluated as the shader text writes it (ray_layouts.glsl:72-75: no built-in is
// involved, so nothing is fused) — with this the reference's own shaders, compiled for the CPU, reproduce every buffer of every
// pass bit for bit (tests/test_cpu_ref_shaders.py)
inline float interp(float a, float b, float c, vec3 w) { return a * w.x + b * w.y + c * w.z; }
inline vec3 interp(vec3 a, vec3 b, vec3 c, vec3 w) {
	return { interp(a.x, b.x, c.x, w), interp(a.y, b.y, c.y, w), interp(a.z, b.z, c.z, w) };
}

// column-major 4x4 (glm): m[4*j + i] = column j, row i
inline vec3 xformPoint(const float* m, vec3 p) {   // vec3(M * vec4(p, 1))
	return {
		fma_(m[8], p.z, fma_(m[4], p.y, fma_(m[0], p.x, m[12]))),
		fma_(m[9], p.z, fma_(m[5], p.y, fma_(m[1], p.x, m[13]))),
		fma_(m[10], p.z, fma_(m[6], p.y, fma_(m[2], p.x, m[14]))) };
}
inline vec4 xformPoint4(const float* m, vec3 p) {  // M * vec4(p, 1)
	vec3 r = xformPoint(m, p);
	return { r.x, r.y, r.z, fma_(m[11], p.z, fma_(m[7], p.y, fma_(m[3], p.x, m[15]))) };
}
inline vec3 xformDir(const float* m, vec3 v) {     // mat3(M) * v
	return {
		fma_(m[8], v.z, fma_(m[4], v.y, m[0] * v.x)),
		fma_(m[9], v.z, fma_(m[5], v.y, m[1] * v.x)),
		fma_(m[10], v.z, fma_(m[6], v.y, m[2] * v.x)) };
}

const float Pi = 3.14159265358979323846f;
const float PiInv = 1.0f / Pi;

// sin / cos: Cody-Waite reduction by pi/2 + degree-7/8 minimax polynomials (Cephes sinf/cosf coefficients).
// Defined for |x| up to a few hundred; used for the concentric-disk mapping and the camera's tan(FOV/2)
// (reference math.glsl:23-39, camera.glsl:33) in place of the GLSL built-ins, whose precision Vulkan leaves
// implementation-defined.
inline void sincos_(float x, float& s, float& c) {
	float k = std::floor(fma_(x, 0.636619772367581343f, 0.5f));
	float r = fma_(k, -1.5707962512969970703125f, x);
	r = fma_(k, -7.54978995489188e-08f, r);
	float r2 = r * r;
	float ps = fma_(fma_(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f);
	float pc = fma_(fma_(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f);
	float sr = fma_(r * r2, ps, r);
	float cr = fma_(r2 * r2, pc, fma_(-0.5f, r2, 1.0f));
	int q = int(k) & 3;
	if (q == 0) { s = sr; c = cr; }
	else if (q == 1) { s = cr; c = -sr; }
	else if (q == 2) { s = -sr; c = -cr; }
	else { s = -cr; c = sr; }
}
inline float tan_(float x) { float s, c; sincos_(x, s, c); return s / c; }

inline float sqr(float x) { return x * x; }
inline float square(float x) { return x * x; }
inline float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }
inline float luminance(vec3 c) { return dot(c, V3(0.299f, 0.587f, 0.114f)); }
inline bool isBlack(vec3 c) { return luminance(c) < 1e-5f; }
inline bool hasNan(vec3 c) { return isnan_(c.x) || isnan_(c.y) || isnan_(c.z); }
inline float satDot(vec3 a, vec3 b) { return max_(dot(a, b), 0.0f); }
inline float absDot(vec3 a, vec3 b) { return abs_(dot(a, b)); }
inline float MISWeight(float f, float g) { return (f * f) / (f * f + g * g); }

// math.glsl:187-192
inline vec3 clampColor(vec3 c) {
	if (hasNan(c)) return V3(0.0f);
	return { clamp_(c.x, 0.0f, 1e4f), clamp_(c.y, 0.0f, 1e4f), clamp_(c.z, 0.0f, 1e4f) };
}

// math.glsl:227-266 — all integer
inline uint32_t hash2(uint32_t seed) {
	seed = (seed ^ 61u) ^ (seed >> 16u);
	seed *= 9u;
	seed = seed ^ (seed >> 4u);
	seed *= 0x27d4eb2du;
	seed = seed ^ (seed >> 15u);
	return seed;
}
inline uint32_t makeSeed(uint32_t rand, uint32_t index) { return hash2(rand) + hash2(index); }
inline uint32_t makeSeed(uint32_t seed, uvec2 index) {
	return makeSeed((seed + index.x) ^ (index.y - 1u), index.y * (index.x - 2u));
}
inline uint32_t urand(uint32_t& rng) { return rng = hash2(rng); }
// float(u) / 4294967295.0: the divisor rounds to 2^32 in fp32, so this is an exact scale; can return 1.0
inline float sample1f(uint32_t& rng) { return float(urand(rng)) * 2.3283064365386963e-10f; }
// GLSL evaluates constructor arguments left to right
inline vec2 sample2f(uint32_t& rng) { vec2 v; v.x = sample1f(rng); v.y = sample1f(rng); return v; }
inline vec3 sample3f(uint32_t& rng) { vec3 v; v.x = sample1f(rng); v.y = sample1f(rng); v.z = sample1f(rng); return v; }
inline vec4 sample4f(uint32_t& rng) { vec4 v; v.x = sample1f(rng); v.y = sample1f(rng); v.z = sample1f(rng); v.w = sample1f(rng); return v; }

// math.glsl:23-39
inline vec2 toConcentricDisk(vec2 v) {
	if (v.x == 0.0f && v.y == 0.0f) return { 0.0f, 0.0f };
	v.x = v.x * 2.0f - 1.0f;
	v.y = v.y * 2.0f - 1.0f;
	float phi, r;
	if (v.x * v.x > v.y * v.y) {
		r = v.x;
		phi = Pi * v.y / v.x * 0.25f;
	}
	else {
		r = v.y;
		phi = Pi * 0.5f - Pi * v.x / v.y * 0.25f;
	}
	float s, c;
	sincos_(phi, s, c);
	return { r * c, r * s };
}

// math.glsl:69-89
inline vec3 getTangent(vec3 n) { return (abs_(n.z) > 0.999f) ? V3(0.0f, 1.0f, 0.0f) : V3(0.0f, 0.0f, 1.0f); }
struct Frame { vec3 t, b, n; };
inline Frame matLocalToWorld(vec3 n) {
	vec3 t = getTangent(n);
	vec3 b = normalize(cross(n, t));
	t = cross(b, n);
	return { t, b, n };
}
inline vec3 frameToWorld(const Frame& f, vec3 v) {   // mat3(t,b,n) * v
	return { fma_(f.n.x, v.z, fma_(f.b.x, v.y, f.t.x * v.x)),
	         fma_(f.n.y, v.z, fma_(f.b.y, v.y, f.t.y * v.x)),
	         fma_(f.n.z, v.z, fma_(f.b.z, v.y, f.t.z * v.x)) };
}
inline vec3 localToWorld(vec3 n, vec3 v) { return normalize(frameToWorld(matLocalToWorld(n), v)); }
inline vec3 sampleCosineWeightedHemisphere(vec3 n, vec2 u) {
	vec2 uv = toConcentricDisk(u);
	float z = std::sqrt(1.0f - dot(uv, uv));
	return localToWorld(n, V3(uv.x, uv.y, z));
}

// math.glsl:117-120
inline vec2 uvToBary(vec2 uv) { float r = std::sqrt(uv.y); return { 1.0f - r, uv.x * r }; }

// fp32 -> fp16 -> fp32, round to nearest even (the RG16F motion-vector target)
inline float roundThroughHalf(float f) {
	uint32_t x; std::memcpy(&x, &f, 4);
	uint32_t sign = x & 0x80000000u;
	uint32_t ax = x & 0x7fffffffu;
	uint32_t out;
	if (ax >= 0x7f800000u) out = ax;                                  // inf / nan
	else if (ax >= 0x477ff000u) out = 0x7f800000u;                    // rounds to >= 65520 -> inf
	else if (ax < 0x33000001u) out = 0;                               // < half of the smallest subnormal (2^-25 ties to 0)
	else if (ax < 0x38800000u) {                                      // subnormal half: quantum 2^-24
		float a; std::memcpy(&a, &ax, 4);
		float q = a * 16777216.0f;                                    // exact
		float r = std::nearbyintf(q);                                 // RNE (default rounding mode)
		float back = r * 5.9604644775390625e-08f;
		std::memcpy(&out, &back, 4);
	}
	else {
		uint32_t lsb = (ax >> 13) & 1u;
		out = (ax + 0xfffu + lsb) & 0xffffe000u;
	}
	out |= sign;
	float r; std::memcpy(&r, &out, 4);
	return r;
}

} // namespace orc
