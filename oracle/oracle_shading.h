// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's BSDF library, light sampling,
// camera and surface fetch (reference src/shader/material.glsl, light_sampling.glsl:1-53, camera.glsl:28-42,
// gbuffer_util.glsl, ray_layouts.glsl:50-104, ray_gbuffer_util.glsl).  PARITY PINNED against the reference's own shaders compiled for the CPU (oracle/ref, tests/test_cpu_ref_shaders.py).
#pragma once
#include "oracle_frame.h"

namespace orc {

// material.glsl:6-20
const uint32_t Diffuse = 1u << 0, Glossy = 1u << 1, Specular = 1u << 2, Reflection = 1u << 4, Transmission = 1u << 5;
const uint32_t InvalidBSDFSample = 0x80000000u;
const uint32_t MatLambert = 1, MatMetallicWorkflow = 2, MatMetal = 3, MatDielectric = 4, MatFake = 6;

struct BSDFSample {
	vec3 wi{ 0, 0, 0 };
	float pdf = 0;
	vec3 bsdf{ 0, 0, 0 };
	uint32_t type = 0;
};

struct SurfaceInfo {
	vec3 pos{ 0, 0, 0 };
	vec3 norm{ 0, 0, 0 };
	vec3 albedo{ 0, 0, 0 };
	uint32_t matIndex = 0;   // left unset for lights in the reference (ray_layouts.glsl:50-56); defined as 0
	bool isLight = false;
};

// ---- BSDFs (material.glsl) ------------------------------------------------------------------------------
inline float fresnelSchlick(float cosTheta, float ior) {                       // :31-34
	float f0 = abs_(1.0f - ior) / (1.0f + ior);
	return mix(f0, 1.0f, pow5(1.0f - cosTheta));
}
inline vec3 fresnelSchlick(float cosTheta, vec3 f0) {                          // :36-38
	return mix(f0, V3(1.0f), pow5(1.0f - cosTheta));
}
// :40-57.  `#define MATERIAL_DIELECTRIC_USE_SCHLICK_APPROX true` + `#if MATERIAL_DIELECTRIC_USE_SCHLICK_APPROX`: inside #if the GLSL
// preprocessor gives an identifier that is no macro — `true` — the value 0, so the compiled reference takes the #else branch, the
// exact Fresnel equations, whatever the macro's name promises (found by running the reference's own shaders on the CPU,
// tests/test_cpu_ref_shaders.py; DESIGN.md §2 "defined behaviours" 9)
inline float fresnel(float cosIn, float ior) {
	if (cosIn < 0) {
		ior = 1.0f / ior;
		cosIn = -cosIn;
	}
	float sinIn = std::sqrt(1.0f - cosIn * cosIn);
	float sinTr = sinIn / ior;
	if (sinTr >= 1.0f) return 1.0f;
	float cosTr = std::sqrt(1.0f - sinTr * sinTr);
	return (square((cosIn - ior * cosTr) / (cosIn + ior * cosTr)) + square((ior * cosIn - cosTr) / (ior * cosIn + cosTr))) * 0.5f;
}
inline float schlickG(float cosTheta, float alpha) {                           // :59-62
	float a = alpha * 0.5f;
	return cosTheta / (cosTheta * (1.0f - a) + a);
}
inline float smithG(float cosWo, float cosWi, float alpha) {                   // :64-66
	return schlickG(abs_(cosWo), alpha) * schlickG(abs_(cosWi), alpha);
}
inline float GTR2Distrib(float cosTheta, float alpha) {                        // :68-77
	if (cosTheta < 1e-6f) return 0.0f;
	float aa = alpha * alpha;
	float denom = cosTheta * cosTheta * (aa - 1.0f) + 1.0f;
	denom = denom * denom * Pi;
	return aa / denom;
}
inline float GTR2Pdf(vec3 n, vec3 m, vec3 wo, float alpha) {                   // :79-82
	return GTR2Distrib(dot(n, m), alpha) * schlickG(dot(n, wo), alpha) * absDot(m, wo) / absDot(n, wo);
}
vec3 GTR2Sample(vec3 n, vec3 wo, float alpha, vec2 r);                         // :84-102
inline bool isGTR2Connectible(float roughness) { return roughness > 0.05f; }   // :104-106
inline bool isGTR2Delta(float roughness) { return roughness < 0.01f; }         // :108-110

vec3 evalBSDF(const RptMaterial& mat, vec3 albedo, vec3 n, vec3 wo, vec3 wi);  // :286-299
float evalPdf(const RptMaterial& mat, vec3 n, vec3 wo, vec3 wi);               // :301-314
bool sampleBSDF(const RptMaterial& mat, vec3 albedo, vec3 n, vec3 wo, vec3 r, BSDFSample& s);   // :316-330
bool isBSDFDelta(const RptMaterial& mat);                                      // :332-344
bool isBSDFConnectible(const RptMaterial& mat);                                // :346-358
inline bool isSampleTypeDelta(uint32_t type) { return (type & Specular) == Specular; }

// ---- lights (light_sampling.glsl:6-53) ------------------------------------------------------------------
struct LightSample {
	vec3 radiance, wi;
	float dist, pdf, jacobian;
	vec2 bary;
	uint32_t id;
};
LightSample sampleLight(const Scene& s, vec3 ref, vec4 r);

// ---- camera (camera.glsl:28-42) -------------------------------------------------------------------------
Ray pinholeCameraSampleRay(const RptCamera& cam, vec2 uv);   // jitter r = 0 at every call site

// ---- surface fetch --------------------------------------------------------------------------------------
void loadSurfaceInfo(const Scene& s, const Intersection& isec, SurfaceInfo& info);   // ray_layouts.glsl:50-104
inline Intersection toIsec(const RptIntersection& r) { return { { r.bary[0], r.bary[1] }, r.instanceIdx, r.triangleIdx }; }
inline RptIntersection fromIsec(const Intersection& i) { return { { i.bary.x, i.bary.y }, i.instanceIdx, i.triangleIdx }; }

// G-buffer access.  texture() on depthNormal is bilinear with REPEAT addressing and 8-bit weight precision;
// texelFetch on albedoMatId uses ivec2(uv * film) (clamped: Vulkan robustness returns 0 out of range,
// we clamp to the edge and document it).
vec4 fetchDepthNormalBilinear(const std::vector<vec4>& img, uint32_t W, uint32_t H, vec2 uv);
bool unpackGBuffer(vec4 depthNormal, uvec2 albedoMatId, float& depth, vec3& normal, vec3& albedo, int& matMeshId);

inline uint32_t packAlbedo(vec3 a) {   // packUnorm4x8(vec4(albedo, 1))
	auto q = [](float c) { return uint32_t(std::floor(clamp_(c, 0.0f, 1.0f) * 255.0f + 0.5f)); };
	return q(a.x) | (q(a.y) << 8) | (q(a.z) << 16) | (255u << 24);
}
inline vec3 unpackAlbedo(uint32_t p) {
	return { float(p & 0xffu) / 255.0f, float((p >> 8) & 0xffu) / 255.0f, float((p >> 16) & 0xffu) / 255.0f };
}

} // namespace orc
