// Device math for the sm_100a kernels.  Implements the numeric contract of DESIGN.md §numerics: IEEE fp32,
// round-to-nearest, the translation unit is compiled with -fmad=false so nothing is contracted implicitly;
// fused multiply-adds appear only where __fmaf_rn is written.  Divisions / square roots are the correctly
// rounded ones (-prec-div/-prec-sqrt).  sin/cos come from the polynomial below, not from libdevice.
// GLSL counterparts: reference src/shader/math.glsl.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define RT_DEV __device__ __forceinline__

namespace rt {

RT_DEV float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
RT_DEV float min_(float a, float b) { return a < b ? a : b; }
RT_DEV float max_(float a, float b) { return a > b ? a : b; }
RT_DEV float clamp_(float x, float lo, float hi) { return min_(max_(x, lo), hi); }
RT_DEV float abs_(float x) { return fabsf(x); }
RT_DEV bool isnan_(float x) { return x != x; }

RT_DEV float3 f3(float s) { return make_float3(s, s, s); }
RT_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
RT_DEV float3 f3(const float* p) { return make_float3(p[0], p[1], p[2]); }
RT_DEV float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }

RT_DEV float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DEV float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DEV float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
RT_DEV float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_DEV float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
// vec3 / float: one correctly rounded reciprocal and three multiplies — the form a GPU compiler gives GLSL's vector-by-scalar
// division (the oracle divides the same way; three IEEE divisions per vector cost grisBounceKernel 9 %, profiles/r2_13_*)
RT_DEV float3 operator/(float3 a, float s) { const float r = 1.0f / s; return make_float3(a.x * r, a.y * r, a.z * r); }
RT_DEV float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
RT_DEV float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
RT_DEV float3& operator*=(float3& a, float s) { a = a * s; return a; }
RT_DEV float3& operator/=(float3& a, float s) { a = a / s; return a; }

RT_DEV float dot(float3 a, float3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
RT_DEV float dot(float2 a, float2 b) { return fma_(a.y, b.y, a.x * b.x); }
RT_DEV float3 cross(float3 a, float3 b) {
	return make_float3(fma_(a.y, b.z, -(a.z * b.y)), fma_(a.z, b.x, -(a.x * b.z)), fma_(a.x, b.y, -(a.y * b.x)));
}
RT_DEV float length(float3 a) { return sqrtf(dot(a, a)); }
RT_DEV float length(float2 a) { return sqrtf(dot(a, a)); }
RT_DEV float3 normalize(float3 a) { return a * (1.0f / length(a)); }
RT_DEV float distance(float3 a, float3 b) { return length(a - b); }
RT_DEV float mix(float a, float b, float t) { return fma_(b, t, a * (1.0f - t)); }
RT_DEV float3 mix(float3 a, float3 b, float t) { return make_float3(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)); }
RT_DEV float3 mix(float3 a, float3 b, float3 t) { return make_float3(mix(a.x, b.x, t.x), mix(a.y, b.y, t.y), mix(a.z, b.z, t.z)); }
RT_DEV float3 reflect(float3 I, float3 N) { float k = 2.0f * dot(N, I); return I - N * k; }
// barycentric interpolation as the shader text writes it (ray_layouts.glsl:72-75; the unit is compiled with -fmad=false, so the
// three products and two sums stay separate, as in the oracle and in the reference's text compiled for the CPU)
RT_DEV float interp(float a, float b, float c, float3 w) { return a * w.x + b * w.y + c * w.z; }
RT_DEV float3 interp(float3 a, float3 b, float3 c, float3 w) {
	return make_float3(interp(a.x, b.x, c.x, w), interp(a.y, b.y, c.y, w), interp(a.z, b.z, c.z, w));
}

// column-major 4x4: m[4*j + i] = column j, row i
RT_DEV float3 xformPoint(const float* __restrict__ m, float3 p) {
	return make_float3(
		fma_(m[8], p.z, fma_(m[4], p.y, fma_(m[0], p.x, m[12]))),
		fma_(m[9], p.z, fma_(m[5], p.y, fma_(m[1], p.x, m[13]))),
		fma_(m[10], p.z, fma_(m[6], p.y, fma_(m[2], p.x, m[14]))));
}
RT_DEV float4 xformPoint4(const float* __restrict__ m, float3 p) {
	float3 r = xformPoint(m, p);
	return make_float4(r.x, r.y, r.z, fma_(m[11], p.z, fma_(m[7], p.y, fma_(m[3], p.x, m[15]))));
}
RT_DEV float3 xformDir(const float* __restrict__ m, float3 v) {
	return make_float3(
		fma_(m[8], v.z, fma_(m[4], v.y, m[0] * v.x)),
		fma_(m[9], v.z, fma_(m[5], v.y, m[1] * v.x)),
		fma_(m[10], v.z, fma_(m[6], v.y, m[2] * v.x)));
}

#define RT_PI 3.14159265358979323846f
#define RT_PI_INV (1.0f / RT_PI)

// Cody-Waite reduction by pi/2 + Cephes sinf/cosf minimax polynomials; same formula on the CPU side
RT_DEV void sincos_(float x, float& s, float& c) {
	float k = floorf(fma_(x, 0.636619772367581343f, 0.5f));
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
RT_DEV float tan_(float x) { float s, c; sincos_(x, s, c); return s / c; }

RT_DEV float square(float x) { return x * x; }
RT_DEV float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }
RT_DEV float luminance(float3 c) { return dot(c, make_float3(0.299f, 0.587f, 0.114f)); }
RT_DEV bool isBlack(float3 c) { return luminance(c) < 1e-5f; }
RT_DEV bool hasNan(float3 c) { return isnan_(c.x) || isnan_(c.y) || isnan_(c.z); }
RT_DEV float satDot(float3 a, float3 b) { return max_(dot(a, b), 0.0f); }
RT_DEV float absDot(float3 a, float3 b) { return abs_(dot(a, b)); }
RT_DEV float MISWeight(float f, float g) { return (f * f) / (f * f + g * g); }

RT_DEV float3 clampColor(float3 c) {   // math.glsl:187-192
	if (hasNan(c)) return f3(0.0f);
	return make_float3(clamp_(c.x, 0.0f, 1e4f), clamp_(c.y, 0.0f, 1e4f), clamp_(c.z, 0.0f, 1e4f));
}

// math.glsl:227-266 (integer RNG)
RT_DEV uint32_t hash2(uint32_t seed) {
	seed = (seed ^ 61u) ^ (seed >> 16u);
	seed *= 9u;
	seed = seed ^ (seed >> 4u);
	seed *= 0x27d4eb2du;
	seed = seed ^ (seed >> 15u);
	return seed;
}
RT_DEV uint32_t makeSeed(uint32_t seed, uint32_t x, uint32_t y) {
	return hash2((seed + x) ^ (y - 1u)) + hash2(y * (x - 2u));
}
RT_DEV float sample1f(uint32_t& rng) {
	rng = hash2(rng);
	return __uint2float_rn(rng) * 2.3283064365386963e-10f;   // == float(u) / 4294967295.0 in fp32
}
RT_DEV float2 sample2f(uint32_t& rng) { float2 v; v.x = sample1f(rng); v.y = sample1f(rng); return v; }
RT_DEV float3 sample3f(uint32_t& rng) { float3 v; v.x = sample1f(rng); v.y = sample1f(rng); v.z = sample1f(rng); return v; }
RT_DEV float4 sample4f(uint32_t& rng) { float4 v; v.x = sample1f(rng); v.y = sample1f(rng); v.z = sample1f(rng); v.w = sample1f(rng); return v; }

RT_DEV float2 toConcentricDisk(float2 v) {   // math.glsl:23-39
	if (v.x == 0.0f && v.y == 0.0f) return make_float2(0.0f, 0.0f);
	v.x = v.x * 2.0f - 1.0f;
	v.y = v.y * 2.0f - 1.0f;
	float phi, r;
	if (v.x * v.x > v.y * v.y) {
		r = v.x;
		phi = RT_PI * v.y / v.x * 0.25f;
	}
	else {
		r = v.y;
		phi = RT_PI * 0.5f - RT_PI * v.x / v.y * 0.25f;
	}
	float s, c;
	sincos_(phi, s, c);
	return make_float2(r * c, r * s);
}

struct Frame3 { float3 t, b, n; };
RT_DEV Frame3 matLocalToWorld(float3 n) {   // math.glsl:69-78
	float3 t = (abs_(n.z) > 0.999f) ? make_float3(0.0f, 1.0f, 0.0f) : make_float3(0.0f, 0.0f, 1.0f);
	float3 b = normalize(cross(n, t));
	t = cross(b, n);
	Frame3 f; f.t = t; f.b = b; f.n = n;
	return f;
}
RT_DEV float3 frameToWorld(const Frame3& f, float3 v) {
	return make_float3(fma_(f.n.x, v.z, fma_(f.b.x, v.y, f.t.x * v.x)),
	                   fma_(f.n.y, v.z, fma_(f.b.y, v.y, f.t.y * v.x)),
	                   fma_(f.n.z, v.z, fma_(f.b.z, v.y, f.t.z * v.x)));
}
RT_DEV float3 sampleCosineWeightedHemisphere(float3 n, float2 u) {   // math.glsl:80-89
	float2 uv = toConcentricDisk(u);
	float z = sqrtf(1.0f - dot(uv, uv));
	return normalize(frameToWorld(matLocalToWorld(n), make_float3(uv.x, uv.y, z)));
}
RT_DEV float2 uvToBary(float2 uv) { float r = sqrtf(uv.y); return make_float2(1.0f - r, uv.x * r); }   // :117-120

} // namespace rt
