// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_shading.h).
#include "oracle_shading.h"

namespace orc {

// material.glsl:84-102.  The reference calls inverse(mat3) on an orthonormal frame; the transpose is used
// here (identical in exact arithmetic; GLSL leaves inverse()'s precision undefined).
vec3 GTR2Sample(vec3 n, vec3 wo, float alpha, vec2 r) {
	Frame fr = matLocalToWorld(n);
	vec3 local = V3(dot(fr.t, wo), dot(fr.b, wo), dot(fr.n, wo));
	vec3 vh = normalize(local * V3(alpha, alpha, 1.0f));

	float lenSq = vh.x * vh.x + vh.y * vh.y;
	vec3 t = lenSq > 0.0f ? V3(-vh.y, vh.x, 0.0f) / std::sqrt(lenSq) : V3(1.0f, 0.0f, 0.0f);
	vec3 b = cross(vh, t);

	vec2 p = toConcentricDisk(r);
	float s = 0.5f * (vh.z + 1.0f);
	p.y = (1.0f - s) * std::sqrt(1.0f - p.x * p.x) + s * p.y;

	vec3 wh = t * p.x + b * p.y + vh * std::sqrt(max_(0.0f, 1.0f - dot(p, p)));
	wh = V3(wh.x * alpha, wh.y * alpha, max_(0.0f, wh.z));
	return normalize(frameToWorld(fr, wh));
}

// material.glsl:112-131
static bool refract_(vec3 n, vec3 wi, float ior, vec3& wt) {
	float cosIn = dot(n, wi);
	if (cosIn < 0) ior = 1.0f / ior;
	float sin2In = max_(0.0f, 1.0f - cosIn * cosIn);
	float sin2Tr = sin2In / (ior * ior);
	if (sin2Tr >= 1.0f) return false;
	float cosTr = std::sqrt(1.0f - sin2Tr);
	if (cosIn < 0) cosTr = -cosTr;
	wt = normalize(-wi / ior + n * (cosIn / ior - cosTr));
	return true;
}

// :141-147
static bool lambertSample(vec3 albedo, vec3 n, vec2 r, BSDFSample& s) {
	s.wi = sampleCosineWeightedHemisphere(n, r);
	s.pdf = absDot(n, s.wi) * PiInv;
	s.bsdf = albedo * PiInv;
	s.type = Diffuse | Reflection;
	return true;
}

// :149-172
static bool dielectricSample(RptMaterial mat, vec3 albedo, vec3 n, vec3 wo, vec3 r, BSDFSample& s) {
	float pdfReflect = fresnel(dot(n, wo), mat.ior);
	s.bsdf = albedo;
	if (r.z < pdfReflect) {
		s.wi = reflect(-wo, n);
		s.type = Specular | Reflection;
		s.pdf = 1.0f;
	}
	else {
		if (!refract_(n, wo, mat.ior, s.wi)) {
			s.type = InvalidBSDFSample;
			return false;
		}
		if (dot(n, wo) < 0) mat.ior = 1.0f / mat.ior;
		s.bsdf /= mat.ior * mat.ior;
		s.type = Specular | Transmission;
		s.pdf = 1.0f;
	}
	return true;
}

// :174-191
static vec3 metallicWorkflowBSDF(const RptMaterial& mat, vec3 albedo, vec3 n, vec3 wo, vec3 wi) {
	float alpha = square(mat.roughness);
	vec3 wh = normalize(wo + wi);
	float cosO = dot(n, wo);
	float cosI = dot(n, wi);
	if (cosI * cosO < 1e-7f) return V3(0.0f);
	vec3 f = fresnelSchlick(dot(wh, wo), mix(V3(0.08f), albedo, mat.metallic));
	float g = smithG(cosO, cosI, alpha);
	float d = GTR2Distrib(dot(n, wh), alpha);
	return mix(albedo * PiInv * (1.0f - mat.metallic), V3(g * d / (4.0f * cosI * cosO)), f);
}

// :193-201
static float metallicWorkflowPdf(const RptMaterial& mat, vec3 n, vec3 wo, vec3 wi) {
	vec3 wh = normalize(wo + wi);
	return mix(satDot(n, wi) * PiInv,
	           GTR2Pdf(n, wh, wo, square(mat.roughness)) / (4.0f * absDot(wh, wo)),
	           1.0f / (2.0f - mat.metallic));
}

// :203-226
static bool metallicWorkflowSample(const RptMaterial& mat, vec3 albedo, vec3 n, vec3 wo, vec3 r, BSDFSample& s) {
	float alpha = square(mat.roughness);
	s.type = Reflection;
	if (r.z > (1.0f / (2.0f - mat.metallic))) {
		s.wi = sampleCosineWeightedHemisphere(n, { r.x, r.y });
		s.type |= Diffuse;
	}
	else {
		vec3 wh = GTR2Sample(n, wo, alpha, { r.x, r.y });
		s.wi = -reflect(wo, wh);
		s.type |= isGTR2Delta(mat.roughness) ? Specular : Glossy;
	}
	if (dot(n, s.wi) < 0.0f) {
		s.type = InvalidBSDFSample;
		return false;
	}
	s.bsdf = metallicWorkflowBSDF(mat, albedo, n, wo, s.wi);
	s.pdf = metallicWorkflowPdf(mat, n, wo, s.wi);
	return true;
}

// :228-248
static vec3 metalBSDF(const RptMaterial& mat, vec3 albedo, vec3 n, vec3 wo, vec3 wi) {
	if (isGTR2Delta(mat.roughness)) return V3(0.0f);
	float alpha = square(mat.roughness);
	vec3 wh = normalize(wo + wi);
	float cosO = dot(n, wo);
	float cosI = dot(n, wi);
	if (cosI * cosO < 1e-7f) return V3(0.0f);
	float f = fresnelSchlick(absDot(wh, wo), mat.ior);
	float g = smithG(cosO, cosI, alpha);
	float d = GTR2Distrib(dot(n, wh), alpha);
	return albedo * f * g * d / (4.0f * cosI * cosO);
}

// :250-256
static float metalPdf(const RptMaterial& mat, vec3 n, vec3 wo, vec3 wi) {
	if (isGTR2Delta(mat.roughness)) return 0.0f;
	vec3 wh = normalize(wo + wi);
	return GTR2Pdf(n, wh, wo, square(mat.roughness)) / (4.0f * absDot(wh, wo));
}

// :258-284
static bool metalSample(const RptMaterial& mat, vec3 albedo, vec3 n, vec3 wo, vec3 r, BSDFSample& s) {
	float alpha = square(mat.roughness);
	bool isDelta = isGTR2Delta(mat.roughness);
	if (isDelta) {
		s.wi = -reflect(wo, n);
	}
	else {
		vec3 wh = GTR2Sample(n, wo, alpha, { r.x, r.y });
		s.wi = -reflect(wo, wh);
	}
	if (dot(n, s.wi) < 0.0f) {
		s.type = InvalidBSDFSample;
		return false;
	}
	s.bsdf = isDelta ? albedo * fresnelSchlick(absDot(n, wo), mat.ior) : metalBSDF(mat, albedo, n, wo, s.wi);
	s.pdf = isDelta ? 1.0f : metalPdf(mat, n, wo, s.wi);
	s.type = Reflection | (isDelta ? Specular : Glossy);
	return true;
}

vec3 evalBSDF(const RptMaterial& mat, vec3 albedo, vec3 n, vec3 wo, vec3 wi) {
	switch (mat.type) {
	case MatLambert: return albedo * PiInv;
	case MatMetallicWorkflow: return metallicWorkflowBSDF(mat, albedo, n, wo, wi);
	case MatMetal: return metalBSDF(mat, albedo, n, wo, wi);
	}
	return V3(0.0f);
}

float evalPdf(const RptMaterial& mat, vec3 n, vec3 wo, vec3 wi) {
	switch (mat.type) {
	case MatLambert: return absDot(n, wi) * PiInv;
	case MatMetallicWorkflow: return metallicWorkflowPdf(mat, n, wo, wi);
	case MatMetal: return metalPdf(mat, n, wo, wi);
	}
	return 0.0f;
}

bool sampleBSDF(const RptMaterial& mat, vec3 albedo, vec3 n, vec3 wo, vec3 r, BSDFSample& s) {
	switch (mat.type) {
	case MatLambert: return lambertSample(albedo, n, { r.x, r.y }, s);
	case MatMetallicWorkflow: return metallicWorkflowSample(mat, albedo, n, wo, r, s);
	case MatMetal: return metalSample(mat, albedo, n, wo, r, s);
	case MatDielectric: return dielectricSample(mat, albedo, n, wo, r, s);
	case MatFake:   // :278-284
		s.wi = -wo; s.bsdf = albedo; s.pdf = 1.0f; s.type = Specular | Transmission;
		return true;
	}
	return false;
}

bool isBSDFDelta(const RptMaterial& mat) {
	switch (mat.type) {
	case MatLambert: return false;
	case MatMetallicWorkflow: return isGTR2Delta(mat.roughness) && mat.metallic > 0.9f;
	case MatMetal: return isGTR2Delta(mat.roughness);
	}
	return true;
}

bool isBSDFConnectible(const RptMaterial& mat) {
	switch (mat.type) {
	case MatLambert: return true;
	case MatMetallicWorkflow: return isGTR2Connectible(mat.roughness) || mat.metallic < 0.9f;
	case MatMetal: return isGTR2Connectible(mat.roughness);
	}
	return false;
}

// light_sampling.glsl:6-37
LightSample sampleLight(const Scene& s, vec3 ref, vec4 r) {
	LightSample o;
	float sumPower = s.lightTable[0].prob;
	uint32_t numLights = s.lightTable[0].failId;

	uint32_t id = uint32_t(float(numLights) * r.x);
	// sample1f can return exactly 1.0 -> id == N -> table index N+1 (out of bounds in the reference; Vulkan
	// robust access hides it).  Clamped here; listed under "defined behaviours" in DESIGN.md.
	if (id > numLights - 1u) id = numLights - 1u;
	id = (r.y < s.lightTable[id + 1].prob) ? id : s.lightTable[id + 1].failId - 1u;
	o.id = id;

	const RptTriangleLight& light = s.lights[id];
	vec3 radiance = V3(light.radiance);
	o.bary = uvToBary({ r.z, r.w });
	vec3 pos = V3(light.v0) * (1.0f - o.bary.x - o.bary.y) + V3(light.v1) * o.bary.x + V3(light.v2) * o.bary.y;
	o.dist = distance(ref, pos);
	vec3 n = V3(light.nx, light.ny, light.nz);
	o.wi = (pos - ref) / o.dist;
	o.jacobian = absDot(n, o.wi) / square(o.dist);
	o.pdf = 1.0f / o.jacobian / light.area;
	o.radiance = (dot(n, o.wi) > 0) ? V3(0.0f) : radiance;   // single-sided emitters
	o.pdf *= luminance(radiance) * light.area / sumPower;
	return o;
}

// camera.glsl:28-42 with r = 0
Ray pinholeCameraSampleRay(const RptCamera& cam, vec2 uv) {
	vec2 ndc = { uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f };
	float aspect = float(cam.filmSize[0]) / float(cam.filmSize[1]);
	float tanFOV = tan_((cam.FOV * 0.5f) * 0.017453292519943295f);
	vec3 pFocusPlane = V3(ndc.x * aspect * tanFOV, ndc.y * 1.0f * tanFOV, 1.0f);
	vec3 dir = normalize(pFocusPlane);
	dir = normalize(V3(cam.right) * dir.x + V3(cam.up) * dir.y + V3(cam.front) * dir.z);
	return { V3(cam.pos), dir };
}

// ray_layouts.glsl:50-104
void loadSurfaceInfo(const Scene& s, const Intersection& isec, SurfaceInfo& info) {
	vec3 bary = V3(1.0f - isec.bary.x - isec.bary.y, isec.bary.x, isec.bary.y);
	if (isec.instanceIdx == 0) {
		const RptTriangleLight& light = s.lights[isec.triangleIdx];
		info.pos = V3(light.v0) * bary.x + V3(light.v1) * bary.y + V3(light.v2) * bary.z;
		info.norm = V3(light.nx, light.ny, light.nz);
		info.albedo = V3(light.radiance);
		info.matIndex = 0;
		info.isLight = true;
		return;
	}
	const RptObjectInstance& inst = s.instances[isec.instanceIdx - 1];
	info.matIndex = uint32_t(s.materialIndices[inst.indexOffset / 3 + isec.triangleIdx]);
	const RptMeshVertex& v0 = s.vertices[s.indices[inst.indexOffset + isec.triangleIdx * 3 + 0]];
	const RptMeshVertex& v1 = s.vertices[s.indices[inst.indexOffset + isec.triangleIdx * 3 + 1]];
	const RptMeshVertex& v2 = s.vertices[s.indices[inst.indexOffset + isec.triangleIdx * 3 + 2]];
	vec3 pos = interp(V3(v0.pos), V3(v1.pos), V3(v2.pos), bary);
	vec3 norm = interp(V3(v0.norm), V3(v1.norm), V3(v2.norm), bary);
	float uvx = interp(v0.uvx, v1.uvx, v2.uvx, bary);
	float uvy = interp(v0.uvy, v1.uvy, v2.uvy, bary);
	info.pos = xformPoint(inst.transform, pos);
	info.norm = normalize(xformPoint(inst.transformInvT, norm));   // vec3(invT * vec4(norm, 1.0))
	const RptMaterial& m = s.materials[info.matIndex];
	info.albedo = (m.textureIdx == InvalidResourceIdx) ? V3(m.baseColor) : s.sampleTexture(m.textureIdx, uvx, uvy);
	info.isLight = false;
}

vec4 fetchDepthNormalBilinear(const std::vector<vec4>& img, uint32_t W, uint32_t H, vec2 uv) {
	auto wrap = [](int i, int n) { int m = i % n; return m < 0 ? m + n : m; };
	float x = uv.x * float(W) - 0.5f, y = uv.y * float(H) - 0.5f;
	float fx = std::floor(x), fy = std::floor(y);
	float ax = std::floor((x - fx) * 256.0f + 0.5f) * 0.00390625f;
	float ay = std::floor((y - fy) * 256.0f + 0.5f) * 0.00390625f;
	int x0 = wrap(int(fx), int(W)), x1 = wrap(int(fx) + 1, int(W));
	int y0 = wrap(int(fy), int(H)), y1 = wrap(int(fy) + 1, int(H));
	const vec4& a = img[size_t(y0) * W + x0]; const vec4& b = img[size_t(y0) * W + x1];
	const vec4& c = img[size_t(y1) * W + x0]; const vec4& d = img[size_t(y1) * W + x1];
	auto lerp2 = [&](float p, float q, float r, float t) {
		float top = p * (1.0f - ax) + q * ax;
		float bot = r * (1.0f - ax) + t * ax;
		return top * (1.0f - ay) + bot * ay;
	};
	return { lerp2(a.x, b.x, c.x, d.x), lerp2(a.y, b.y, c.y, d.y), lerp2(a.z, b.z, c.z, d.z), lerp2(a.w, b.w, c.w, d.w) };
}

// gbuffer_util.glsl:25-38
bool unpackGBuffer(vec4 dn, uvec2 am, float& depth, vec3& normal, vec3& albedo, int& matMeshId) {
	depth = dn.x;
	if (depth == 0.0f) return false;
	normal = V3(dn.y, dn.z, dn.w);
	albedo = unpackAlbedo(am.x);
	matMeshId = int(am.y);
	return true;
}

} // namespace orc
