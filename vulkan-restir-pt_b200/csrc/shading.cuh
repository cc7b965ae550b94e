// Device shading library: BSDFs, light sampling, camera rays, surface fetch, G-buffer access.
// CUDA counterparts of reference src/shader/{material.glsl, light_sampling.glsl:1-53, camera.glsl:28-42,
// ray_layouts.glsl:50-104, gbuffer_util.glsl, ray_gbuffer_util.glsl}.  Arithmetic follows DESIGN.md §numerics.
#pragma once
#include "bvh_traverse.cuh"

namespace rt {

// material.glsl:6-20
constexpr uint32_t BsdfDiffuse = 1u << 0, BsdfGlossy = 1u << 1, BsdfSpecular = 1u << 2, BsdfReflection = 1u << 4, BsdfTransmission = 1u << 5;
constexpr uint32_t InvalidBSDFSample = 0x80000000u;
constexpr uint32_t MatLambert = 1, MatMetallicWorkflow = 2, MatMetal = 3, MatDielectric = 4, MatFake = 6;

struct Mat {   // RptMaterial in registers
	float3 baseColor;
	uint32_t type, textureIdx;
	float metallic, roughness, ior;
};

RT_DEV Mat loadMaterial(const SceneView& s, uint32_t idx) {
	const float4* p = reinterpret_cast<const float4*>(s.materials + idx);
	const float4 a = __ldg(p), b = __ldg(p + 1);
	Mat m;
	m.baseColor = make_float3(a.x, a.y, a.z);
	m.type = __float_as_uint(a.w);
	m.textureIdx = __float_as_uint(b.x);
	m.metallic = b.y; m.roughness = b.z; m.ior = b.w;
	return m;
}

struct BSDFSample {
	float3 wi;
	float pdf;
	float3 bsdf;
	uint32_t type;
};
RT_DEV BSDFSample emptyBSDFSample() { BSDFSample s; s.wi = f3(0.0f); s.pdf = 0.0f; s.bsdf = f3(0.0f); s.type = 0; return s; }

struct Surface {
	float3 pos, norm, albedo;
	uint32_t matIndex;
	bool isLight;
};

struct Ray { float3 ori, dir; };

// ---- BSDF building blocks ------------------------------------------------------------------------------
RT_DEV float fresnelSchlick(float cosTheta, float ior) {
	float f0 = abs_(1.0f - ior) / (1.0f + ior);
	return mix(f0, 1.0f, pow5(1.0f - cosTheta));
}
RT_DEV float3 fresnelSchlick(float cosTheta, float3 f0) { return mix(f0, f3(1.0f), pow5(1.0f - cosTheta)); }
// material.glsl:40-57: the exact Fresnel equations.  `#if MATERIAL_DIELECTRIC_USE_SCHLICK_APPROX` with the macro defined as `true`
// is false in the GLSL preprocessor (an identifier that is no macro counts as 0), so this #else branch is what the reference runs
RT_DEV float fresnel(float cosIn, float ior) {
	if (cosIn < 0) {
		ior = 1.0f / ior;
		cosIn = -cosIn;
	}
	const float sinIn = sqrtf(1.0f - cosIn * cosIn);
	const float sinTr = sinIn / ior;
	if (sinTr >= 1.0f) return 1.0f;
	const float cosTr = sqrtf(1.0f - sinTr * sinTr);
	return (square((cosIn - ior * cosTr) / (cosIn + ior * cosTr)) + square((ior * cosIn - cosTr) / (ior * cosIn + cosTr))) * 0.5f;
}
RT_DEV float schlickG(float cosTheta, float alpha) {
	float a = alpha * 0.5f;
	return cosTheta / (cosTheta * (1.0f - a) + a);
}
RT_DEV float smithG(float cosWo, float cosWi, float alpha) { return schlickG(abs_(cosWo), alpha) * schlickG(abs_(cosWi), alpha); }
RT_DEV float GTR2Distrib(float cosTheta, float alpha) {
	if (cosTheta < 1e-6f) return 0.0f;
	float aa = alpha * alpha;
	float denom = cosTheta * cosTheta * (aa - 1.0f) + 1.0f;
	denom = denom * denom * RT_PI;
	return aa / denom;
}
RT_DEV float GTR2Pdf(float3 n, float3 m, float3 wo, float alpha) {
	return GTR2Distrib(dot(n, m), alpha) * schlickG(dot(n, wo), alpha) * absDot(m, wo) / absDot(n, wo);
}
// material.glsl:84-102 (inverse of the orthonormal frame = its transpose)
RT_DEV float3 GTR2Sample(float3 n, float3 wo, float alpha, float2 r) {
	const Frame3 fr = matLocalToWorld(n);
	const float3 local = make_float3(dot(fr.t, wo), dot(fr.b, wo), dot(fr.n, wo));
	const float3 vh = normalize(local * make_float3(alpha, alpha, 1.0f));
	const float lenSq = vh.x * vh.x + vh.y * vh.y;
	const float3 t = lenSq > 0.0f ? make_float3(-vh.y, vh.x, 0.0f) / sqrtf(lenSq) : make_float3(1.0f, 0.0f, 0.0f);
	const float3 b = cross(vh, t);
	float2 p = toConcentricDisk(r);
	const float sc = 0.5f * (vh.z + 1.0f);
	p.y = (1.0f - sc) * sqrtf(1.0f - p.x * p.x) + sc * p.y;
	float3 wh = t * p.x + b * p.y + vh * sqrtf(max_(0.0f, 1.0f - dot(p, p)));
	wh = make_float3(wh.x * alpha, wh.y * alpha, max_(0.0f, wh.z));
	return normalize(frameToWorld(fr, wh));
}
RT_DEV bool isGTR2Connectible(float roughness) { return roughness > 0.05f; }
RT_DEV bool isGTR2Delta(float roughness) { return roughness < 0.01f; }

RT_DEV bool refract_(float3 n, float3 wi, float ior, float3& wt) {   // material.glsl:112-131
	const float cosIn = dot(n, wi);
	if (cosIn < 0) ior = 1.0f / ior;
	const float sin2In = max_(0.0f, 1.0f - cosIn * cosIn);
	const float sin2Tr = sin2In / (ior * ior);
	if (sin2Tr >= 1.0f) return false;
	float cosTr = sqrtf(1.0f - sin2Tr);
	if (cosIn < 0) cosTr = -cosTr;
	wt = normalize(-wi / ior + n * (cosIn / ior - cosTr));
	return true;
}

RT_DEV float3 metallicWorkflowBSDF(const Mat& mat, float3 albedo, float3 n, float3 wo, float3 wi) {   // :174-191
	const float alpha = square(mat.roughness);
	const float3 wh = normalize(wo + wi);
	const float cosO = dot(n, wo), cosI = dot(n, wi);
	if (cosI * cosO < 1e-7f) return f3(0.0f);
	const float3 f = fresnelSchlick(dot(wh, wo), mix(f3(0.08f), albedo, mat.metallic));
	const float g = smithG(cosO, cosI, alpha);
	const float d = GTR2Distrib(dot(n, wh), alpha);
	return mix(albedo * RT_PI_INV * (1.0f - mat.metallic), f3(g * d / (4.0f * cosI * cosO)), f);
}
RT_DEV float metallicWorkflowPdf(const Mat& mat, float3 n, float3 wo, float3 wi) {   // :193-201
	const float3 wh = normalize(wo + wi);
	return mix(satDot(n, wi) * RT_PI_INV,
	           GTR2Pdf(n, wh, wo, square(mat.roughness)) / (4.0f * absDot(wh, wo)),
	           1.0f / (2.0f - mat.metallic));
}
RT_DEV float3 metalBSDF(const Mat& mat, float3 albedo, float3 n, float3 wo, float3 wi) {   // :228-248
	if (isGTR2Delta(mat.roughness)) return f3(0.0f);
	const float alpha = square(mat.roughness);
	const float3 wh = normalize(wo + wi);
	const float cosO = dot(n, wo), cosI = dot(n, wi);
	if (cosI * cosO < 1e-7f) return f3(0.0f);
	const float f = fresnelSchlick(absDot(wh, wo), mat.ior);
	const float g = smithG(cosO, cosI, alpha);
	const float d = GTR2Distrib(dot(n, wh), alpha);
	return albedo * f * g * d / (4.0f * cosI * cosO);
}
RT_DEV float metalPdf(const Mat& mat, float3 n, float3 wo, float3 wi) {   // :250-256
	if (isGTR2Delta(mat.roughness)) return 0.0f;
	const float3 wh = normalize(wo + wi);
	return GTR2Pdf(n, wh, wo, square(mat.roughness)) / (4.0f * absDot(wh, wo));
}

RT_DEV float3 evalBSDF(const Mat& mat, float3 albedo, float3 n, float3 wo, float3 wi) {   // :286-299
	if (mat.type == MatLambert) return albedo * RT_PI_INV;
	if (mat.type == MatMetallicWorkflow) return metallicWorkflowBSDF(mat, albedo, n, wo, wi);
	if (mat.type == MatMetal) return metalBSDF(mat, albedo, n, wo, wi);
	return f3(0.0f);
}
RT_DEV float evalPdf(const Mat& mat, float3 n, float3 wo, float3 wi) {   // :301-314
	if (mat.type == MatLambert) return absDot(n, wi) * RT_PI_INV;
	if (mat.type == MatMetallicWorkflow) return metallicWorkflowPdf(mat, n, wo, wi);
	if (mat.type == MatMetal) return metalPdf(mat, n, wo, wi);
	return 0.0f;
}

RT_DEV bool sampleBSDF(const Mat& mat, float3 albedo, float3 n, float3 wo, float3 r, BSDFSample& s) {   // :316-330
	if (mat.type == MatLambert) {   // :141-147
		s.wi = sampleCosineWeightedHemisphere(n, make_float2(r.x, r.y));
		s.pdf = absDot(n, s.wi) * RT_PI_INV;
		s.bsdf = albedo * RT_PI_INV;
		s.type = BsdfDiffuse | BsdfReflection;
		return true;
	}
	if (mat.type == MatMetallicWorkflow) {   // :203-226
		const float alpha = square(mat.roughness);
		s.type = BsdfReflection;
		if (r.z > (1.0f / (2.0f - mat.metallic))) {
			s.wi = sampleCosineWeightedHemisphere(n, make_float2(r.x, r.y));
			s.type |= BsdfDiffuse;
		}
		else {
			const float3 wh = GTR2Sample(n, wo, alpha, make_float2(r.x, r.y));
			s.wi = -reflect(wo, wh);
			s.type |= isGTR2Delta(mat.roughness) ? BsdfSpecular : BsdfGlossy;
		}
		if (dot(n, s.wi) < 0.0f) { s.type = InvalidBSDFSample; return false; }
		s.bsdf = metallicWorkflowBSDF(mat, albedo, n, wo, s.wi);
		s.pdf = metallicWorkflowPdf(mat, n, wo, s.wi);
		return true;
	}
	if (mat.type == MatMetal) {   // :258-284
		const float alpha = square(mat.roughness);
		const bool isDelta = isGTR2Delta(mat.roughness);
		if (isDelta) {
			s.wi = -reflect(wo, n);
		}
		else {
			const float3 wh = GTR2Sample(n, wo, alpha, make_float2(r.x, r.y));
			s.wi = -reflect(wo, wh);
		}
		if (dot(n, s.wi) < 0.0f) { s.type = InvalidBSDFSample; return false; }
		s.bsdf = isDelta ? albedo * fresnelSchlick(absDot(n, wo), mat.ior) : metalBSDF(mat, albedo, n, wo, s.wi);
		s.pdf = isDelta ? 1.0f : metalPdf(mat, n, wo, s.wi);
		s.type = BsdfReflection | (isDelta ? BsdfSpecular : BsdfGlossy);
		return true;
	}
	if (mat.type == MatDielectric) {   // :149-172
		float ior = mat.ior;
		const float pdfReflect = fresnel(dot(n, wo), ior);
		s.bsdf = albedo;
		if (r.z < pdfReflect) {
			s.wi = reflect(-wo, n);
			s.type = BsdfSpecular | BsdfReflection;
			s.pdf = 1.0f;
		}
		else {
			if (!refract_(n, wo, ior, s.wi)) { s.type = InvalidBSDFSample; return false; }
			if (dot(n, wo) < 0) ior = 1.0f / ior;
			s.bsdf /= ior * ior;
			s.type = BsdfSpecular | BsdfTransmission;
			s.pdf = 1.0f;
		}
		return true;
	}
	if (mat.type == MatFake) {   // :278-284
		s.wi = -wo; s.bsdf = albedo; s.pdf = 1.0f; s.type = BsdfSpecular | BsdfTransmission;
		return true;
	}
	return false;
}

RT_DEV bool isBSDFDelta(const Mat& mat) {   // :332-344
	if (mat.type == MatLambert) return false;
	if (mat.type == MatMetallicWorkflow) return isGTR2Delta(mat.roughness) && mat.metallic > 0.9f;
	if (mat.type == MatMetal) return isGTR2Delta(mat.roughness);
	return true;
}
RT_DEV bool isBSDFConnectible(const Mat& mat) {   // :346-358
	if (mat.type == MatLambert) return true;
	if (mat.type == MatMetallicWorkflow) return isGTR2Connectible(mat.roughness) || mat.metallic < 0.9f;
	if (mat.type == MatMetal) return isGTR2Connectible(mat.roughness);
	return false;
}
RT_DEV bool isSampleTypeDelta(uint32_t type) { return (type & BsdfSpecular) == BsdfSpecular; }

// ---- light sampling (light_sampling.glsl:6-37) -----------------------------------------------------------
struct LightSample {
	float3 radiance, wi;
	float dist, pdf, jacobian;
	float2 bary;
	uint32_t id;
};

RT_DEV LightSample sampleLight(const SceneView& s, float3 ref, float4 r) {
	LightSample o;
	const RptLightSampleTableElement head = s.lightTable[0];
	const float sumPower = head.prob;
	const uint32_t numLights = head.failId;
	uint32_t id = uint32_t(float(numLights) * r.x);
	if (id > numLights - 1u) id = numLights - 1u;   // r.x == 1.0 (DESIGN.md "defined behaviours")
	const RptLightSampleTableElement e = s.lightTable[id + 1];
	id = (r.y < e.prob) ? id : e.failId - 1u;
	o.id = id;
	const float4* lp = reinterpret_cast<const float4*>(s.lights + id);
	const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2), l3 = __ldg(lp + 3);
	const float3 radiance = f3(l3);
	const float area = l3.w;
	o.bary = uvToBary(make_float2(r.z, r.w));
	const float3 pos = f3(l0) * (1.0f - o.bary.x - o.bary.y) + f3(l1) * o.bary.x + f3(l2) * o.bary.y;
	o.dist = distance(ref, pos);
	const float3 n = make_float3(l0.w, l1.w, l2.w);
	o.wi = (pos - ref) / o.dist;
	o.jacobian = absDot(n, o.wi) / square(o.dist);
	o.pdf = 1.0f / o.jacobian / area;
	o.radiance = (dot(n, o.wi) > 0) ? f3(0.0f) : radiance;
	o.pdf *= luminance(radiance) * area / sumPower;
	return o;
}

// ---- camera (camera.glsl:28-42 with r = 0) -----------------------------------------------------------------
RT_DEV Ray pinholeCameraSampleRay(const RptCamera& cam, float2 uv) {
	const float2 ndc = make_float2(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f);
	const float aspect = float(cam.filmSize[0]) / float(cam.filmSize[1]);
	const float tanFOV = tan_((cam.FOV * 0.5f) * 0.017453292519943295f);
	const float3 pFocusPlane = make_float3(ndc.x * aspect * tanFOV, ndc.y * 1.0f * tanFOV, 1.0f);
	float3 dir = normalize(pFocusPlane);
	dir = normalize(f3(cam.right) * dir.x + f3(cam.up) * dir.y + f3(cam.front) * dir.z);
	Ray r; r.ori = f3(cam.pos); r.dir = dir;
	return r;
}

// ---- textures -----------------------------------------------------------------------------------------------
RT_DEV int wrapRepeat(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

RT_DEV float3 sampleTexture(const SceneView& s, uint32_t texIdx, float u, float v) {
	const TextureView t = s.textures[texIdx];
	const int W = int(t.width), H = int(t.height);
	auto texel = [&](int x, int y) {
		const uchar4 c = __ldg(t.texels + (size_t(y) * W + x));
		return make_float3(__ldg(s.srgbToLinear + c.x), __ldg(s.srgbToLinear + c.y), __ldg(s.srgbToLinear + c.z));
	};
	if (!(abs_(u) < 1e6f) || !(abs_(v) < 1e6f)) { u = 0.f; v = 0.f; }
	if (t.filter == 1) {
		return texel(wrapRepeat(int(floorf(u * float(W))), W), wrapRepeat(int(floorf(v * float(H))), H));
	}
	const float x = u * float(W) - 0.5f, y = v * float(H) - 0.5f;
	const float fx = floorf(x), fy = floorf(y);
	const float ax = floorf((x - fx) * 256.0f + 0.5f) * 0.00390625f;
	const float ay = floorf((y - fy) * 256.0f + 0.5f) * 0.00390625f;
	const int x0 = wrapRepeat(int(fx), W), x1 = wrapRepeat(int(fx) + 1, W);
	const int y0 = wrapRepeat(int(fy), H), y1 = wrapRepeat(int(fy) + 1, H);
	const float3 top = texel(x0, y0) * (1.0f - ax) + texel(x1, y0) * ax;
	const float3 bot = texel(x0, y1) * (1.0f - ax) + texel(x1, y1) * ax;
	return top * (1.0f - ay) + bot * ay;
}

// ---- surface fetch (ray_layouts.glsl:50-104) ----------------------------------------------------------------
RT_DEV void loadSurfaceInfo(const SceneView& s, uint32_t instanceIdx, uint32_t triangleIdx, float2 isecBary, Surface& info) {
	const float3 bary = make_float3(1.0f - isecBary.x - isecBary.y, isecBary.x, isecBary.y);
	if (s.counters != nullptr) atomicAdd(&s.counters[4], 1ull);
	if (instanceIdx == 0) {
		const float4* lp = reinterpret_cast<const float4*>(s.lights + triangleIdx);
		const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2), l3 = __ldg(lp + 3);
		info.pos = f3(l0) * bary.x + f3(l1) * bary.y + f3(l2) * bary.z;
		info.norm = make_float3(l0.w, l1.w, l2.w);
		info.albedo = f3(l3);
		info.matIndex = 0;   // unset in the reference; defined as 0
		info.isLight = true;
		return;
	}
	const RptObjectInstance* inst = s.instances + (instanceIdx - 1);
	const uint32_t indexOffset = inst->indexOffset;
	info.matIndex = uint32_t(__ldg(s.materialIndices + (indexOffset / 3 + triangleIdx)));
	const uint32_t* ip = s.indices + (indexOffset + triangleIdx * 3);
	const float4* v0 = reinterpret_cast<const float4*>(s.vertices + __ldg(ip));
	const float4* v1 = reinterpret_cast<const float4*>(s.vertices + __ldg(ip + 1));
	const float4* v2 = reinterpret_cast<const float4*>(s.vertices + __ldg(ip + 2));
	const float4 p0 = __ldg(v0), n0 = __ldg(v0 + 1), p1 = __ldg(v1), n1 = __ldg(v1 + 1), p2 = __ldg(v2), n2 = __ldg(v2 + 1);
	const float3 pos = interp(f3(p0), f3(p1), f3(p2), bary);
	const float3 norm = interp(f3(n0), f3(n1), f3(n2), bary);
	const float uvx = interp(p0.w, p1.w, p2.w, bary);
	const float uvy = interp(n0.w, n1.w, n2.w, bary);
	info.pos = xformPoint(inst->transform, pos);
	info.norm = normalize(xformPoint(inst->transformInvT, norm));
	const Mat m = loadMaterial(s, info.matIndex);
	info.albedo = (m.textureIdx == InvalidResourceIdx) ? m.baseColor : sampleTexture(s, m.textureIdx, uvx, uvy);
	info.isLight = false;
}
RT_DEV void loadSurfaceInfo(const SceneView& s, const Hit& h, Surface& info) { loadSurfaceInfo(s, h.instanceIdx, h.triangleIdx, make_float2(h.u, h.v), info); }
RT_DEV void loadSurfaceInfo(const SceneView& s, const RptIntersection& i, Surface& info) { loadSurfaceInfo(s, i.instanceIdx, i.triangleIdx, make_float2(i.bary[0], i.bary[1]), info); }

// ---- G-buffer access ------------------------------------------------------------------------------------------
RT_DEV uint32_t packAlbedo(float3 a) {
	const uint32_t r = uint32_t(floorf(clamp_(a.x, 0.0f, 1.0f) * 255.0f + 0.5f));
	const uint32_t g = uint32_t(floorf(clamp_(a.y, 0.0f, 1.0f) * 255.0f + 0.5f));
	const uint32_t b = uint32_t(floorf(clamp_(a.z, 0.0f, 1.0f) * 255.0f + 0.5f));
	return r | (g << 8) | (b << 16) | (255u << 24);
}
RT_DEV float3 unpackAlbedo(uint32_t p) {
	return make_float3(float(p & 0xffu) / 255.0f, float((p >> 8) & 0xffu) / 255.0f, float((p >> 16) & 0xffu) / 255.0f);
}

// storage row of film row y in the depthNormal images.  A strip of a multi-GPU film additionally keeps film rows 0
// and H-1 in two extra rows behind its stored rows, because REPEAT addressing lets a bilinear tap at the top / bottom
// edge of the film wrap around to the other side (reference sampler: zvk/core/Memory.cpp:75-92).
RT_DEV uint32_t depthNormalRow(const FrameView& f, int y) {
	const int lo = int(f.storeBegin), hi = int(f.storeEnd);
	if (y >= lo && y < hi) return uint32_t(y - lo);
	if (y == 0) return uint32_t(hi - lo);
	if (y == int(f.height) - 1) return uint32_t(hi - lo) + 1u;
	return uint32_t((y < lo ? lo : hi - 1) - lo);   // not reachable for lookups within the halo
}

// texture(uDepthNormal*, uv): bilinear, REPEAT, 8-bit weights
RT_DEV float4 fetchDepthNormalBilinear(const FrameView& f, const float4* __restrict__ img, float2 uv) {
	const int W = int(f.width), H = int(f.height);
	const float x = uv.x * float(W) - 0.5f, y = uv.y * float(H) - 0.5f;
	const float fx = floorf(x), fy = floorf(y);
	const float ax = floorf((x - fx) * 256.0f + 0.5f) * 0.00390625f;
	const float ay = floorf((y - fy) * 256.0f + 0.5f) * 0.00390625f;
	const int x0 = wrapRepeat(int(fx), W), x1 = wrapRepeat(int(fx) + 1, W);
	const size_t r0 = size_t(depthNormalRow(f, wrapRepeat(int(fy), H))) * f.width, r1 = size_t(depthNormalRow(f, wrapRepeat(int(fy) + 1, H))) * f.width;
	const float4 a = img[r0 + x0], b = img[r0 + x1], c = img[r1 + x0], d = img[r1 + x1];
	auto lerp2 = [&](float p, float q, float r, float t) {
		const float top = p * (1.0f - ax) + q * ax;
		const float bot = r * (1.0f - ax) + t * ax;
		return top * (1.0f - ay) + bot * ay;
	};
	return make_float4(lerp2(a.x, b.x, c.x, d.x), lerp2(a.y, b.y, c.y, d.y), lerp2(a.z, b.z, c.z, d.z), lerp2(a.w, b.w, c.w, d.w));
}

// what every ray pass derives from the G-buffer at its own pixel centre
struct Primary {
	bool valid;
	float2 uv;
	float depth;
	float3 norm, albedo;
	int matMeshId, matId;
	Ray ray;
	float3 pos;
};

RT_DEV Primary loadPrimary(const FrameView& f, uint32_t x, uint32_t y) {
	Primary p;
	p.uv = make_float2((float(x) + 0.5f) / float(f.width), (float(y) + 0.5f) / float(f.height));
	const size_t i = f.index(x, y);
	const float4 dn = f.depthNormal[i];
	p.depth = dn.x;
	p.valid = !(p.depth == 0.0f);
	if (!p.valid) return p;
	const uint2 am = f.albedoMatId[i];
	p.norm = make_float3(dn.y, dn.z, dn.w);
	p.albedo = unpackAlbedo(am.x);
	p.matMeshId = int(am.y);
	p.matId = p.matMeshId >> 16;
	p.ray = pinholeCameraSampleRay(f.camera, make_float2(p.uv.x, 1.0f - p.uv.y));
	p.pos = p.ray.ori + p.ray.dir * (p.depth - 1e-4f);
	return p;
}

RT_DEV Surface primarySurface(const Primary& p) {
	Surface sf;
	sf.pos = p.pos; sf.norm = p.norm; sf.albedo = p.albedo; sf.matIndex = uint32_t(p.matId); sf.isLight = false;
	return sf;
}

// surface lookup at an arbitrary uv of this frame's or the previous frame's G-buffer
struct Neighbor {
	bool found;
	size_t pixel;        // storage index of ivec2(uv * film)
	float depth;
	float3 norm, albedo, pos;
	int matMeshId;
};

RT_DEV Neighbor lookupSurface(const FrameView& f, bool previousFrame, float2 uv) {
	Neighbor nb;
	nb.found = false;
	if (uv.x < 0 || uv.y < 0 || uv.x > 1.0f || uv.y > 1.0f) return nb;
	int px = int(uv.x * float(f.width)), py = int(uv.y * float(f.height));
	if (px > int(f.width) - 1) px = int(f.width) - 1;
	if (py > int(f.height) - 1) py = int(f.height) - 1;
	// multi-GPU strips: every lookup is GPU-local.  Current-frame reservoirs exist for owned + halo rows.
	// Previous-frame ones exist for the owned rows, plus — when the neighbour is connected and mirrors the boundary rows of its
	// final reservoirs into this GPU's halo rows — the halo rows whose bilinear taps stay inside the stored rows
	// (prevRowBegin / prevRowEnd, set by the host).  A lookup outside fails.
	if (previousFrame ? (py < int(f.prevRowBegin) || py >= int(f.prevRowEnd)) : (py < int(f.storeBegin) || py >= int(f.storeEnd))) return nb;
	const float4 dn = fetchDepthNormalBilinear(f, previousFrame ? f.depthNormalPrev : f.depthNormal, uv);
	nb.depth = dn.x;
	if (nb.depth == 0.0f) return nb;
	nb.pixel = f.index(uint32_t(px), uint32_t(py));
	const uint2 am = previousFrame ? f.albedoMatIdPrev[nb.pixel] : f.albedoMatId[nb.pixel];
	nb.norm = make_float3(dn.y, dn.z, dn.w);
	nb.albedo = unpackAlbedo(am.x);
	nb.matMeshId = int(am.y);
	const Ray ray = pinholeCameraSampleRay(previousFrame ? f.prevCamera : f.camera, make_float2(uv.x, 1.0f - uv.y));
	nb.pos = ray.ori + ray.dir * (nb.depth - 1e-4f);
	nb.found = true;
	return nb;
}

// multi-GPU strips: film rows of this strip that are halo rows of the strip above / below, and where they live in the
// neighbour's storage (peer memory over NVLink; the neighbour's buffers have the same row-major layout from its storeBegin)
RT_DEV bool rowInUpHalo(const FrameView& f, uint32_t y) { return y < f.rowBegin + f.halo; }
RT_DEV bool rowInDownHalo(const FrameView& f, uint32_t y) { return y + f.halo >= f.rowEnd; }
RT_DEV size_t peerUpIndex(const FrameView& f, uint32_t x, uint32_t y) { return size_t(y - f.peerUpStoreBegin) * f.width + x; }
RT_DEV size_t peerDownIndex(const FrameView& f, uint32_t x, uint32_t y) { return size_t(y - f.peerDownStoreBegin) * f.width + x; }

RT_DEV void accumulate(float4* __restrict__ img, const FrameView& f, uint32_t x, uint32_t y, float3 c) {
	const float n = float(f.camera.frameIndex & 0x7fffffffu);
	const size_t i = f.index(x, y);
	const float4 px = img[i];
	float3 acc = make_float3(px.x, px.y, px.z);
	acc = (acc * n + c) / (n + 1.0f);
	img[i] = make_float4(acc.x, acc.y, acc.z, 1.0f);
}

} // namespace rt
