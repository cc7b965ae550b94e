// Shim over the reference's own shader library — src/shader/math.glsl (RNG, sampling helpers), material.glsl (the BSDFs) and
// light_sampling.glsl — compiled by g++ from where the files lie (oracle/ref/Makefile: glsl_to_cpp.py rewrites GLSL-only syntax
// into a temporary directory, glsl_compat.h supplies vecN / mat3 / the built-in functions).  TEST INFRASTRUCTURE: pins the oracle's
// restatement of these functions (oracle_math.h, oracle_shading.cpp) against the reference's text, tests/test_cpu_ref_pins.py.
#include "glsl_compat.h"
#ifdef GLSL_BUILTINS_CONTRACT
#define REF_FN(x) refc_##x
#else
#define REF_FN(x) ref_##x
#endif
namespace REF_FN(lib) {
#include "math.inc"             // generated from $(REF)/src/shader/math.glsl
#include "material.inc"         //                              material.glsl
#include "light_sampling.inc"   //                              light_sampling.glsl
}
using namespace REF_FN(lib);

#define REF_API extern "C" __attribute__((visibility("default")))
// compiled twice (oracle/ref/Makefile): ref_* with IEEE built-ins + libm, refc_* with the numeric contract's built-in library

static vec3 v3(const float* p) { return vec3(p[0], p[1], p[2]); }
static void put(float* o, vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

REF_API uint32_t REF_FN(hash2)(uint32_t x) { return hash2(x); }
REF_API uint32_t REF_FN(make_seed)(uint32_t seed, uint32_t x, uint32_t y) { uvec2 i; i.x = x; i.y = y; return makeSeed(seed, i); }
REF_API float REF_FN(sample1f)(uint32_t* rng) { return sample1f(*rng); }
REF_API void REF_FN(sample4f)(uint32_t* rng, float* out4) { vec4 v = sample4f(*rng); out4[0] = v.x; out4[1] = v.y; out4[2] = v.z; out4[3] = v.w; }
REF_API void REF_FN(sample3f)(uint32_t* rng, float* out3) { put(out3, sample3f(*rng)); }
REF_API void REF_FN(concentric_disk)(float u, float v, float* out2) { vec2 d = toConcentricDisk(vec2(u, v)); out2[0] = d.x; out2[1] = d.y; }
REF_API void REF_FN(cosine_hemisphere)(const float* n, float u, float v, float* out3) { put(out3, sampleCosineWeightedHemisphere(v3(n), vec2(u, v))); }
REF_API void REF_FN(uv_to_bary)(float u, float v, float* out2) { vec2 b = uvToBary(vec2(u, v)); out2[0] = b.x; out2[1] = b.y; }
REF_API float REF_FN(luminance)(const float* c) { return luminance(v3(c)); }
REF_API void REF_FN(clamp_color)(const float* c, float* out3) { put(out3, clampColor(v3(c))); }

REF_API void REF_FN(eval_bsdf)(const Material* m, const float* albedo, const float* n, const float* wo, const float* wi, float* out3, float* pdf) {
	put(out3, evalBSDF(*m, v3(albedo), v3(n), v3(wo), v3(wi)));
	*pdf = evalPdf(*m, v3(n), v3(wo), v3(wi));
}
REF_API int REF_FN(sample_bsdf)(const Material* m, const float* albedo, const float* n, const float* wo, const float* r3, float* wi, float* bsdf, float* pdf, uint32_t* type) {
	BSDFSample s;
	s.pdf = 0.0f; s.type = 0;
	const bool ok = sampleBSDF(*m, v3(albedo), v3(n), v3(wo), v3(r3), s);
	put(wi, s.wi); put(bsdf, s.bsdf); *pdf = s.pdf; *type = s.type;
	return ok ? 1 : 0;
}
REF_API int REF_FN(is_bsdf_delta)(const Material* m) { return isBSDFDelta(*m) ? 1 : 0; }
REF_API int REF_FN(is_bsdf_connectible)(const Material* m) { return isBSDFConnectible(*m) ? 1 : 0; }

// light_sampling.glsl:24-37 through the 8-output overload (:50-53); the two "uniform buffers" are the caller's arrays
REF_API void REF_FN(sample_light)(const void* lightTable, const void* lights, const float* ref, const float* r4,
                              float* radiance, float* wi, float* dist, float* pdf, float* jacobian, float* bary, uint32_t* id) {
	uLightSampleTable = static_cast<const LightSampleTableElement*>(lightTable);
	uTriangleLights = static_cast<const TriangleLight*>(lights);
	vec3 w; vec2 b; uint i = 0;
	const vec3 L = sampleLight(v3(ref), w, *dist, *pdf, *jacobian, b, i, vec4(r4[0], r4[1], r4[2], r4[3]));
	put(radiance, L); put(wi, w); bary[0] = b.x; bary[1] = b.y; *id = i;
}
