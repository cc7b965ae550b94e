// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's per-pixel passes:
//   G-buffer            GBuffer.vert:19-30, GBuffer.frag:21-54 (rasteriser replaced by a pixel-centre primary ray)
//   naive DI / GI       di_naive.glsl:79-166 (+ .comp), gi_naive.glsl:28-148 (+ .comp)
//   ReSTIR DI           di_reservoir.glsl, di_path_gen.glsl, di_temporal.glsl, di_spatial.glsl (+ .comp)
//   ReSTIR GI           gi_reservoir.glsl, gi_resample_temporal.glsl (+ .comp)
//   ReSTIR PT (GRIS)    gris_reservoir.glsl:1-136, gris_path_trace.glsl, gris_retrace.glsl:42-236,
//                       gris_resample_temporal.glsl, gris_resample_spatial.glsl (+ .comp)
//   post-process        post_proc.frag:16-42
// One function per shader entry; locals the GLSL leaves uninitialised are zero here ("defined
// behaviours").  PARITY PINNED against the reference's own shaders compiled for the CPU (oracle/ref, tests/test_cpu_ref_shaders.py):
// every ray pass below writes the same bits as the reference's .comp shader executed on the same inputs.
#include "oracle_shading.h"

namespace orc {

namespace {

const uint32_t CameraClearFlag = 0x80000000u;
const uint32_t CameraFrameIndexMask = 0x7fffffffu;
const int MaxTracingDepth = 15;

// what every ray pass derives from the G-buffer at its own pixel centre
struct Primary {
	bool valid;
	vec2 uv;
	float depth;
	vec3 norm, albedo;
	int matMeshId, matId;
	Ray ray;
	vec3 pos;
};

Primary loadPrimary(const Frame2D& f, uint32_t x, uint32_t y) {
	Primary p{};
	const uint32_t W = f.width, H = f.height;
	p.uv = { (float(x) + 0.5f) / float(W), (float(y) + 0.5f) / float(H) };
	// texture() at a pixel centre returns that texel (weights quantise to 0 / 1)
	vec4 dn = f.depthNormal[f.cur][size_t(y) * W + x];
	uvec2 am = f.albedoMatId[f.cur][size_t(y) * W + x];
	p.valid = unpackGBuffer(dn, am, p.depth, p.norm, p.albedo, p.matMeshId);
	if (!p.valid) return p;
	p.matId = p.matMeshId >> 16;
	p.ray = pinholeCameraSampleRay(f.camera, { p.uv.x, 1.0f - p.uv.y });
	p.pos = p.ray.ori + p.ray.dir * (p.depth - 1e-4f);
	return p;
}

void accumulate(std::vector<vec4>& img, const Frame2D& f, uint32_t x, uint32_t y, vec3 c) {
	float n = float(f.camera.frameIndex & CameraFrameIndexMask);
	vec4& px = img[size_t(y) * f.width + x];
	vec3 acc = V3(px.x, px.y, px.z);
	acc = (acc * n + c) / (n + 1.0f);
	px = { acc.x, acc.y, acc.z, 1.0f };
}

// neighbour / previous-frame surface lookup shared by the temporal and spatial passes
// (di_temporal.glsl:9-32, di_spatial.glsl:9-32, gris_resample_*.glsl:11-34, gi_resample_temporal.glsl:9-32)
struct Neighbor {
	bool found;
	uint32_t pixel;
	float depth;
	vec3 norm, albedo, pos;
	int matMeshId;
};

Neighbor lookupSurface(const Frame2D& f, bool previousFrame, vec2 uv) {
	Neighbor nb{};
	if (uv.x < 0 || uv.y < 0 || uv.x > 1.0f || uv.y > 1.0f) return nb;
	const uint32_t W = f.width, H = f.height;
	int px = int(uv.x * float(W)), py = int(uv.y * float(H));
	if (px > int(W) - 1) px = int(W) - 1;   // uv == 1.0: out of range in the reference, clamped here
	if (py > int(H) - 1) py = int(H) - 1;
	const uint32_t which = previousFrame ? (f.cur ^ 1u) : f.cur;
	vec4 dn = fetchDepthNormalBilinear(f.depthNormal[which], W, H, uv);
	uvec2 am = f.albedoMatId[which][size_t(py) * W + px];
	if (!unpackGBuffer(dn, am, nb.depth, nb.norm, nb.albedo, nb.matMeshId)) return nb;
	Ray ray = pinholeCameraSampleRay(previousFrame ? f.prevCamera : f.camera, { uv.x, 1.0f - uv.y });
	nb.pos = ray.ori + ray.dir * (nb.depth - 1e-4f);
	nb.pixel = uint32_t(py) * W + uint32_t(px);
	nb.found = true;
	return nb;
}

struct StreamRIS {   // di_naive.glsl:54-77
	vec3 Li{ 0, 0, 0 };
	float weight = 0, sumWeight = 0;
	uint32_t sampleCount = 0;
	void add(vec3 L, float w, float r) {
		sumWeight += w;
		if (r * sumWeight < w) { weight = w; Li = L; sampleCount++; }
	}
};

} // namespace

// ---------------------------------------------------------------------------------------------------------
// G-buffer
// ---------------------------------------------------------------------------------------------------------
void passGBuffer(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1) {
	const uint32_t W = f.width, H = f.height;
	const RptCamera& cam = f.camera;
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < W; x++) {
		size_t i = size_t(y) * W + x;
		vec2 uv = { (float(x) + 0.5f) / float(W), (float(y) + 0.5f) / float(H) };
		Ray ray = pinholeCameraSampleRay(cam, { uv.x, 1.0f - uv.y });
		// the rasteriser draws object instances only (GBufferPass.cpp:50-54): the light mesh is invisible
		Intersection isec = s.traceClosestHit(ray.ori, cam.nearZ, ray.dir, MaxRayDistance, true);
		f.primaryIsec[i] = fromIsec(isec);
		if (isec.instanceIdx == InvalidHitIndex) {
			f.depthNormal[f.cur][i] = { 0, 0, 0, 0 };
			f.albedoMatId[f.cur][i] = { 0, 0 };
			f.motion[i] = { 0, 0 };
			continue;
		}
		const uint32_t instIdx = isec.instanceIdx - 1;
		const RptObjectInstance& inst = s.instances[instIdx];
		uint32_t matIndex = uint32_t(s.materialIndices[inst.indexOffset / 3 + isec.triangleIdx]);
		const RptMeshVertex* v[3];
		for (int c = 0; c < 3; c++) v[c] = &s.vertices[s.indices[inst.indexOffset + isec.triangleIdx * 3 + c]];
		vec3 bary = V3(1.0f - isec.bary.x - isec.bary.y, isec.bary.x, isec.bary.y);
		vec3 posL = interp(V3(v[0]->pos), V3(v[1]->pos), V3(v[2]->pos), bary);
		vec3 P = xformPoint(inst.transform, posL);
		// GBuffer.vert:26: per-vertex normalize(mat3(invT) * n), interpolated, re-normalised in the fragment stage
		vec3 n0 = normalize(xformDir(inst.transformInvT, V3(v[0]->norm)));
		vec3 n1 = normalize(xformDir(inst.transformInvT, V3(v[1]->norm)));
		vec3 n2 = normalize(xformDir(inst.transformInvT, V3(v[2]->norm)));
		vec3 N = normalize(interp(n0, n1, n2, bary));
		float uvx = interp(v[0]->uvx, v[1]->uvx, v[2]->uvx, bary);
		float uvy = interp(v[0]->uvy, v[1]->uvy, v[2]->uvy, bary);
		const RptMaterial& mat = s.materials[matIndex];
		vec3 albedo = (mat.textureIdx == InvalidResourceIdx) ? V3(mat.baseColor) : s.sampleTexture(mat.textureIdx, uvx, uvy);

		// per-instance motion (rpt_scene_update_instances .. rpt_scene_end_motion): the point's position under last frame's placement
		vec3 Plast = s.prevInstances.empty() ? P : xformPoint(s.prevInstances[instIdx].transform, posL);
		vec4 last = xformPoint4(cam.lastProjView, Plast);
		vec2 lastCoord = { (last.x / last.w) * 0.5f + 0.5f, (last.y / last.w) * 0.5f + 0.5f };
		vec2 motion = { lastCoord.x - uv.x, lastCoord.y - uv.y };

		f.depthNormal[f.cur][i] = { length(V3(cam.pos) - P), N.x, N.y, N.z };
		f.albedoMatId[f.cur][i] = { packAlbedo(albedo), (matIndex << 16) | instIdx };
		f.motion[i] = { roundThroughHalf(motion.x), roundThroughHalf(motion.y) };
	}
}

// ---------------------------------------------------------------------------------------------------------
// naive direct illumination (di_naive.comp -> directIllumination2)
// ---------------------------------------------------------------------------------------------------------
static vec3 naiveDirect(const Scene& s, const Frame2D& f, uint32_t x, uint32_t y) {
	Primary p = loadPrimary(f, x, y);
	if (!p.valid) return V3(0.0f);
	uint32_t rng = makeSeed(f.camera.seed, uvec2{ x, y });
	vec3 radiance = V3(0.0f);
	vec3 wo = -p.ray.dir;
	const RptMaterial& mat = s.materials[p.matId];
	StreamRIS resv;

	if (!isBSDFDelta(mat)) {
		LightSample ls = sampleLight(s, p.pos, sample4f(rng));
		bool shadowed = s.traceShadow(p.pos, MinRayDistance, ls.wi, ls.dist - 1e-4f);
		if (!shadowed && ls.pdf > 1e-6f) {
			float bsdfPdf = evalPdf(mat, p.norm, -p.ray.dir, ls.wi);
			float weight = MISWeight(ls.pdf, bsdfPdf);
			vec3 contrib = ls.radiance * evalBSDF(mat, p.albedo, p.norm, wo, ls.wi) * satDot(p.norm, ls.wi) / ls.pdf * weight;
			resv.add(contrib, 100, sample1f(rng));
			radiance += contrib;
		}
	}
	BSDFSample bs;
	vec3 r3 = sample3f(rng);
	if (sampleBSDF(mat, p.albedo, p.norm, wo, r3, bs) && bs.pdf > 1e-6f) {
		Intersection isec = s.traceClosestHit(p.pos, MinRayDistance, bs.wi, MaxRayDistance);
		if (isec.instanceIdx != InvalidHitIndex) {
			SurfaceInfo surf;
			loadSurfaceInfo(s, isec, surf);
			float cosTheta = -dot(bs.wi, surf.norm);
			if (surf.isLight && cosTheta > 0) {
				float dist = length(surf.pos - p.pos);
				float sumPower = s.lightTable[0].prob;
				float lightPdf = luminance(surf.albedo) / sumPower * dist * dist / abs_(cosTheta);
				float weight = isSampleTypeDelta(bs.type) ? 1.0f : MISWeight(bs.pdf, lightPdf);
				float cosTerm = isSampleTypeDelta(bs.type) ? 1.0f : satDot(p.norm, bs.wi);
				vec3 contrib = surf.albedo * bs.bsdf * cosTerm / bs.pdf * weight;
				resv.add(contrib, 1, sample1f(rng));
				radiance += contrib;
			}
		}
	}
	if (resv.weight > 0 && resv.sumWeight > 0) radiance = resv.Li * resv.sumWeight / resv.weight;
	else radiance = V3(0.0f);
	return clampColor(radiance);
}

// RT-pipeline mode (di_naive.rgen -> directIllumination, di_naive.glsl:9-52): one light sample, MIS weight computed and
// then forced to 1 (:47), no BSDF sample
static vec3 naiveDirectRT(const Scene& s, const Frame2D& f, uint32_t x, uint32_t y) {
	Primary p = loadPrimary(f, x, y);
	if (!p.valid) return V3(0.0f);
	uint32_t rng = makeSeed(f.camera.seed, uvec2{ x, y });
	vec3 radiance = V3(0.0f);
	vec3 wo = -p.ray.dir;
	const RptMaterial& mat = s.materials[p.matId];
	if (!isBSDFDelta(mat)) {
		LightSample ls = sampleLight(s, p.pos, sample4f(rng));
		bool shadowed = s.traceShadow(p.pos, MinRayDistance, ls.wi, ls.dist - 1e-4f);
		if (!shadowed && ls.pdf > 1e-6f) {
			float weight = 1.0f;
			radiance += ls.radiance * evalBSDF(mat, p.albedo, p.norm, wo, ls.wi) * satDot(p.norm, ls.wi) / ls.pdf * weight;
		}
	}
	return clampColor(radiance);
}

void passDINaiveRT(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1) {
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++)
		accumulate(f.directOutput, f, x, y, naiveDirectRT(s, f, x, y));
}

void passDINaive(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1) {
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++)
		accumulate(f.directOutput, f, x, y, naiveDirect(s, f, x, y));
}

// ---------------------------------------------------------------------------------------------------------
// naive path tracing for bounces >= 1 (gi_naive.comp -> indirectIllumination)
// ---------------------------------------------------------------------------------------------------------
static vec3 naiveIndirect(const Scene& s, const Frame2D& f, uint32_t x, uint32_t y) {
	Primary p = loadPrimary(f, x, y);
	if (!p.valid) return V3(0.0f);
	Ray ray = p.ray;
	uint32_t rng = makeSeed(f.camera.seed, uvec2{ x, y });
	vec3 radiance = V3(0.0f), throughput = V3(1.0f), lastPos = V3(0.0f);
	vec3 wo = -ray.dir;
	SurfaceInfo surf;
	surf.pos = p.pos; surf.norm = p.norm; surf.albedo = p.albedo; surf.isLight = false;
	RptMaterial mat = s.materials[p.matId];
	BSDFSample bs;

	for (int bounce = 0; bounce < MaxTracingDepth; bounce++) {
		if (bounce > 0) {
			Intersection isec = s.traceClosestHit(ray.ori, MinRayDistance, ray.dir, MaxRayDistance);
			if (isec.instanceIdx == InvalidHitIndex) break;
			loadSurfaceInfo(s, isec, surf);
			mat = s.materials[surf.matIndex];
		}
		if (surf.isLight) {
			float cosTheta = -dot(ray.dir, surf.norm);
			if (bounce > 1 && cosTheta > 0) {
				float weight = 1.0f;
				if (!isSampleTypeDelta(bs.type)) {
					float dist = length(surf.pos - lastPos);
					float sumPower = s.lightTable[0].prob;
					float lightPdf = luminance(surf.albedo) / sumPower * dist * dist / abs_(cosTheta);
					weight = MISWeight(bs.pdf, lightPdf);
				}
				radiance += surf.albedo * weight * throughput;
			}
			break;
		}
		if (bounce > 0 && !isBSDFDelta(mat)) {
			LightSample ls = sampleLight(s, surf.pos, sample4f(rng));
			bool shadowed = s.traceShadow(surf.pos, MinRayDistance, ls.wi, ls.dist - MinRayDistance);
			if (!shadowed && ls.pdf > 1e-6f) {
				float weight = 1.0f;   // MIS weight computed then forced to 1 (gi_naive.glsl:116-118)
				radiance += ls.radiance * evalBSDF(mat, surf.albedo, surf.norm, wo, ls.wi) * satDot(surf.norm, ls.wi) / ls.pdf * weight * throughput;
			}
		}
		if (bounce > 4) {
			float pdfTerminate = max_(1.0f - luminance(throughput), 0.0f);
			if (sample1f(rng) < pdfTerminate) break;
			throughput /= (1.0f - pdfTerminate);
		}
		vec3 r3 = sample3f(rng);
		if (!sampleBSDF(mat, surf.albedo, surf.norm, wo, r3, bs) || bs.pdf < 1e-6f) break;
		float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
		throughput *= bs.bsdf * cosTheta / bs.pdf;
		lastPos = surf.pos;
		wo = -bs.wi;
		ray.dir = bs.wi;
		ray.ori = surf.pos + ray.dir * 1e-4f;
	}
	return clampColor(radiance);
}

void passGINaive(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1) {
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++)
		accumulate(f.indirectOutput, f, x, y, naiveIndirect(s, f, x, y));
}

// ---------------------------------------------------------------------------------------------------------
// ReSTIR DI (di_reservoir.glsl)
// ---------------------------------------------------------------------------------------------------------
namespace {

void diReset(RptDIReservoir& r) { r.sampleCount = 0; r.resampleWeight = 0.0f; r.contribWeight = 0.0f; }
bool diValid(const RptDIReservoir& r) { return !isnan_(r.resampleWeight); }
void diResetIfInvalid(RptDIReservoir& r) { if (!diValid(r)) diReset(r); }
bool diSampleValid(const RptDIReservoir& r) { return r.isec.instanceIdx != InvalidHitIndex; }

// the sample part of a reservoir (DIPathSample) is everything before sampleCount
void diCopySample(RptDIReservoir& dst, const RptDIReservoir& src) { std::memcpy(&dst, &src, 48); }

void diAddSample(RptDIReservoir& resv, const RptDIReservoir& sample, float w, float r) {   // :51-59
	resv.resampleWeight += w;
	resv.sampleCount++;
	if (r * resv.resampleWeight < w) { diCopySample(resv, sample); resv.weight = w; }
}
void diMerge(RptDIReservoir& resv, const RptDIReservoir& rhs, float r) {   // :61-69
	resv.resampleWeight += rhs.resampleWeight;
	resv.sampleCount += rhs.sampleCount;
	if (r * resv.resampleWeight < rhs.resampleWeight) { diCopySample(resv, rhs); resv.weight = rhs.weight; }
}
void diCap(RptDIReservoir& resv, uint32_t cap) {   // :71-76
	if (resv.sampleCount > cap) {
		resv.resampleWeight *= float(cap) / float(resv.sampleCount);
		resv.sampleCount = cap;
	}
}

// di_reservoir.glsl:78-188
vec3 diSampleLi(const Scene& s, const RptDISettings& st, const SurfaceInfo& surf, const RptMaterial& mat, vec3 wo,
                uint32_t rng, uint32_t& resvRng, RptDIReservoir& resv) {
	vec3 radiance = V3(0.0f);
	RptDIReservoir ps{};   // only the DIPathSample part is used
	ps.rng = rng;
	vec4 lightRand = sample4f(rng);
	vec3 scatterRand = sample3f(rng);

	if (st.sampleType != 1 && !isBSDFDelta(mat)) {
		LightSample ls = sampleLight(s, surf.pos, lightRand);
		bool shadowed = s.traceShadow(surf.pos, MinRayDistance, ls.wi, ls.dist - MinRayDistance);
		if (!shadowed && ls.pdf > 1e-6f) {
			float bsdfPdf = evalPdf(mat, surf.norm, wo, ls.wi);
			float weight = MISWeight(ls.pdf, bsdfPdf);
			if (st.sampleType == 0) weight = 1.0f;
			vec3 contrib = ls.radiance * evalBSDF(mat, surf.albedo, surf.norm, wo, ls.wi) * satDot(surf.norm, ls.wi) / ls.pdf * weight;
			float sampleWeight = luminance(contrib);
			if (isnan_(sampleWeight) || sampleWeight < 0) sampleWeight = 0;
			ps.isec = { { ls.bary.x, ls.bary.y }, 0, ls.id };
			vec3 Li = ls.radiance * weight;
			ps.Li[0] = Li.x; ps.Li[1] = Li.y; ps.Li[2] = Li.z;
			ps.jacobian = ls.jacobian;
			ps.samplePdf = ls.pdf;
			ps.isLightSample = 1;
			diAddSample(resv, ps, sampleWeight, sample1f(resvRng));
			radiance += contrib;
		}
	}
	BSDFSample bs;
	ps.isLightSample = 0;
	if (st.sampleType != 0 && sampleBSDF(mat, surf.albedo, surf.norm, wo, scatterRand, bs) && bs.pdf > 1e-6f) {
		Intersection isec = s.traceClosestHit(surf.pos, MinRayDistance, bs.wi, MaxRayDistance);
		if (isec.instanceIdx != InvalidHitIndex) {
			SurfaceInfo hit;
			loadSurfaceInfo(s, isec, hit);
			float cosTheta = -dot(bs.wi, hit.norm);
			if (hit.isLight && cosTheta > 0) {
				float dist = length(hit.pos - surf.pos);
				float sumPower = s.lightTable[0].prob;
				float lightPdf = luminance(hit.albedo) / sumPower * dist * dist / abs_(cosTheta);
				float weight = MISWeight(bs.pdf, lightPdf);
				if (st.sampleType == 1 || isSampleTypeDelta(bs.type)) weight = 1.0f;
				float cosTerm = isSampleTypeDelta(bs.type) ? 1.0f : satDot(surf.norm, bs.wi);
				vec3 contrib = hit.albedo * bs.bsdf * cosTerm / bs.pdf * weight;
				ps.isec = fromIsec(isec);
				vec3 Li = hit.albedo * weight;
				ps.Li[0] = Li.x; ps.Li[1] = Li.y; ps.Li[2] = Li.z;
				ps.jacobian = abs_(cosTheta) / square(dist);
				ps.samplePdf = bs.pdf;
				ps.isLightSample = 0;
				diAddSample(resv, ps, luminance(contrib), sample1f(resvRng));
				radiance += contrib;
			}
		}
	}
	diResetIfInvalid(resv);
	if (resv.sampleCount > 0 && diSampleValid(resv) && resv.weight > 0) {
		float k = resv.resampleWeight / resv.weight;
		resv.Li[0] *= k; resv.Li[1] *= k; resv.Li[2] *= k;
		resv.weight = resv.resampleWeight;
	}
	else {
		resv.isec.instanceIdx = InvalidHitIndex;
		resv.weight = 0;
		resv.resampleWeight = 0;
	}
	resv.sampleCount = 1;
	return radiance;
}

// di_reservoir.glsl:190-224
void diRandomReplay(const Scene& s, const RptDISettings& st, RptDIReservoir& dst, const SurfaceInfo& dstSurf,
                    const RptDIReservoir& src, vec3 wo, uint32_t& rng) {
	const RptMaterial& dstMat = s.materials[dstSurf.matIndex];
	RptDIReservoir replay{};
	diReset(replay);
	diSampleLi(s, st, dstSurf, dstMat, wo, src.rng, rng, replay);
	float jacobian = 1;
	if (diSampleValid(replay)) {
		SurfaceInfo rs;
		loadSurfaceInfo(s, toIsec(replay.isec), rs);
		vec3 wi = normalize(rs.pos - dstSurf.pos);
		vec3 Li = V3(replay.Li) * evalBSDF(dstMat, dstSurf.albedo, dstSurf.norm, wo, wi) * satDot(dstSurf.norm, wi) / replay.samplePdf;
		float dstPHat = luminance(Li * jacobian);
		replay.resampleWeight = src.resampleWeight * dstPHat / src.weight;
		replay.sampleCount = src.sampleCount;
	}
	else {
		replay.resampleWeight = 0;
	}
	if (diValid(replay)) diMerge(dst, replay, sample1f(rng));
}

// di_reservoir.glsl:226-286
void diReconnection(const Scene& s, RptDIReservoir& dst, const SurfaceInfo& dstSurf, RptDIReservoir src, vec3 wo, uint32_t& rng) {
	const RptMaterial& dstMat = s.materials[dstSurf.matIndex];
	SurfaceInfo rc;
	bool srcSampleValid = false;
	float dstPHat = 0, dstSamplePdf = 0, dstJacobian = 0;

	// the reference calls loadSurfaceInfo on the source intersection before checking that it is valid; an
	// invalid index reads out of bounds there, so the validity test is hoisted (results are unused otherwise)
	if (diSampleValid(src)) {
		loadSurfaceInfo(s, toIsec(src.isec), rc);
		float dist = distance(rc.pos, dstSurf.pos);
		vec3 wi = normalize(rc.pos - dstSurf.pos);
		float cosTheta = -dot(rc.norm, wi);
		dstJacobian = abs_(cosTheta) / square(dist);
		float jacobian = dstJacobian / src.jacobian;
		if (dist > 1e-4f) {
			if (cosTheta > 0 && !isnan_(jacobian) && src.jacobian > 0) {
				if (s.traceVisibility(dstSurf.pos, rc.pos)) {
					srcSampleValid = true;
					if (!isnan_(src.samplePdf) && src.samplePdf > 1e-6f) {
						vec3 Li = V3(src.Li) * evalBSDF(dstMat, dstSurf.albedo, dstSurf.norm, wo, wi) * satDot(dstSurf.norm, wi) / src.samplePdf;
						dstPHat = luminance(Li * jacobian);
					}
					if (src.isLightSample) {
						float sumPower = s.lightTable[0].prob;
						dstSamplePdf = luminance(rc.albedo) / sumPower / dstJacobian;
					}
					else {
						dstSamplePdf = evalPdf(dstMat, dstSurf.norm, wo, wi);
					}
				}
			}
		}
	}
	if (srcSampleValid) {
		src.jacobian = dstJacobian;
		src.samplePdf = dstSamplePdf;
		if (src.samplePdf < 1e-6f || isnan_(src.samplePdf)) src.samplePdf = 0;
		src.resampleWeight *= dstPHat / src.weight;
		if (isnan_(src.resampleWeight)) src.resampleWeight = 0;
	}
	else {
		src.resampleWeight = 0;
	}
	if (diValid(src)) diMerge(dst, src, sample1f(rng));
}

void diReuseAndMerge(const Scene& s, const RptDISettings& st, RptDIReservoir& dst, const SurfaceInfo& dstSurf,
                     const RptDIReservoir& src, vec3 wo, uint32_t& rng) {   // :288-295
	if (st.shiftType == 0) diReconnection(s, dst, dstSurf, src, wo, rng);
	else if (st.shiftType == 1) diRandomReplay(s, st, dst, dstSurf, src, wo, rng);
}

SurfaceInfo primarySurface(const Primary& p) {
	SurfaceInfo sf;
	sf.pos = p.pos; sf.norm = p.norm; sf.albedo = p.albedo; sf.matIndex = uint32_t(p.matId); sf.isLight = false;
	return sf;
}

// visibility re-check of the selected sample (di_temporal.glsl:72-81, di_spatial.glsl:78-87)
void diRecheckVisibility(const Scene& s, RptDIReservoir& resv, vec3 pos) {
	if (diValid(resv) && diSampleValid(resv)) {
		SurfaceInfo surf;
		loadSurfaceInfo(s, toIsec(resv.isec), surf);
		if (!s.traceVisibility(pos, surf.pos)) resv.resampleWeight = 0;
	}
}

} // namespace

// di_path_gen.glsl:9-35 (the image store of di_path_gen.comp is commented out: only the reservoir is written)
void passDIPathGen(const Scene& s, Frame2D& f, const RptDISettings& st, uint32_t y0, uint32_t y1) {
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++) {
		Primary p = loadPrimary(f, x, y);
		if (!p.valid) continue;
		uint32_t rng = makeSeed(f.camera.seed, uvec2{ x, y });
		uint32_t resvRng = ~rng;
		RptDIReservoir resv{};
		diReset(resv);
		diSampleLi(s, st, primarySurface(p), s.materials[p.matId], -p.ray.dir, rng, resvRng, resv);
		f.di[f.cur][size_t(y) * f.width + x] = resv;
	}
}

// di_temporal.glsl:34-89
void passDITemporal(const Scene& s, Frame2D& f, const RptDISettings& st, uint32_t y0, uint32_t y1) {
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++) {
		Primary p = loadPrimary(f, x, y);
		if (!p.valid) continue;
		size_t idx = size_t(y) * f.width + x;
		vec2 motion = f.motion[idx];
		uint32_t rng = makeSeed(f.camera.seed, uvec2{ x, y }) ^ 1u;
		uint32_t resvRng = ~rng;
		vec3 wo = -p.ray.dir;
		RptDIReservoir resv = f.di[f.cur][idx];

		if (st.temporalReuse) {
			SurfaceInfo dstSurf = primarySurface(p);
			if ((f.camera.frameIndex & CameraClearFlag) == 0) {
				Neighbor nb = lookupSurface(f, true, { p.uv.x + motion.x, p.uv.y + motion.y });
				if (nb.found && !(nb.matMeshId != p.matMeshId || dot(nb.norm, p.norm) < 0.95f || distance(p.pos, nb.pos) > 0.5f)) {
					const RptDIReservoir& prev = f.di[f.cur ^ 1u][nb.pixel];
					if (diValid(prev)) diReuseAndMerge(s, st, resv, dstSurf, prev, wo, resvRng);
				}
			}
			diRecheckVisibility(s, resv, p.pos);
		}
		diCap(resv, 40);
		diResetIfInvalid(resv);
		f.diTemp[idx] = resv;
	}
}

// di_spatial.glsl:34-119 + di_spatial.comp
void passDISpatial(const Scene& s, Frame2D& f, const RptDISettings& st, uint32_t y0, uint32_t y1) {
	const float texelX = 1.0f / float(f.width), texelY = 1.0f / float(f.height);
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++) {
		Primary p = loadPrimary(f, x, y);
		vec3 radiance = V3(0.0f);
		if (p.valid) {
			size_t idx = size_t(y) * f.width + x;
			uint32_t rng = makeSeed(f.camera.seed, uvec2{ x, y }) ^ 2u;
			vec3 wo = -p.ray.dir;
			const RptMaterial& mat = s.materials[p.matId];
			RptDIReservoir resv = f.diTemp[idx];

			if (st.spatialReuse) {
				SurfaceInfo dstSurf = primarySurface(p);
				for (uint32_t i = 0; i < 10; i++) {
					vec2 d = toConcentricDisk(sample2f(rng));
					vec2 nuv = { p.uv.x + d.x * 20.0f * texelX, p.uv.y + d.y * 20.0f * texelY };
					Neighbor nb = lookupSurface(f, false, nuv);
					if (nb.found && !(dot(nb.norm, p.norm) < 0.9f || distance(p.pos, nb.pos) > 0.4f)) {
						const RptDIReservoir& nr = f.diTemp[nb.pixel];
						if (diValid(nr)) diReuseAndMerge(s, st, resv, dstSurf, nr, wo, rng);
					}
				}
				diRecheckVisibility(s, resv, p.pos);
			}
			diCap(resv, 40);
			diResetIfInvalid(resv);
			f.di[f.cur][idx] = resv;

			if (diValid(resv) && diSampleValid(resv)) {
				SurfaceInfo surf;
				loadSurfaceInfo(s, toIsec(resv.isec), surf);
				vec3 wi = normalize(surf.pos - p.pos);
				if (resv.sampleCount > 0) {
					vec3 Li = V3(resv.Li) * evalBSDF(mat, p.albedo, p.norm, wo, wi) * satDot(p.norm, wi) / resv.samplePdf;
					if (!isBlack(Li)) radiance = Li / luminance(Li) * resv.resampleWeight / float(resv.sampleCount);
				}
			}
			radiance = clampColor(radiance);
		}
		accumulate(f.directOutput, f, x, y, radiance);
	}
}

// ---------------------------------------------------------------------------------------------------------
// ReSTIR GI in one kernel (gi_resample_temporal.glsl:34-212 + .comp)
// ---------------------------------------------------------------------------------------------------------
namespace {

void giReset(RptGIReservoir& r) { r.sampleCount = 0; r.resampleWeight = 0.0f; r.contribWeight = 0.0f; }
bool giValid(const RptGIReservoir& r) { return !isnan_(r.resampleWeight) && r.resampleWeight >= 0; }

} // namespace

void passGIReSTIR(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1) {
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++) {
		size_t idx = size_t(y) * f.width + x;
		Primary p = loadPrimary(f, x, y);
		if (!p.valid) {
			giReset(f.gi[f.cur][idx]);
			accumulate(f.indirectOutput, f, x, y, V3(0.0f));
			continue;
		}
		vec2 motion = f.motion[idx];
		Ray ray = p.ray;
		uint32_t rng = makeSeed(f.camera.seed, uvec2{ x, y });
		vec3 throughputAfter = V3(1.0f), lastPos = V3(0.0f);
		vec3 wo = -ray.dir;
		SurfaceInfo surf;
		surf.pos = p.pos; surf.norm = p.norm; surf.albedo = p.albedo; surf.isLight = false;

		RptGIReservoir ps{};   // GIPathSample part
		ps.rcIsec.instanceIdx = InvalidHitIndex;
		vec3 rcLo = V3(0.0f);
		ps.rcPrevCoord = (y << 16) | x;

		vec3 primaryPos = surf.pos, primaryWo = -ray.dir;
		vec3 primaryScatter = V3(0.0f);
		float primaryPdf = 0.0f;
		RptMaterial mat = s.materials[p.matId];
		BSDFSample bs;

		for (int bounce = 0; bounce < MaxTracingDepth; bounce++) {
			if (bounce > 0) {
				Intersection isec = s.traceClosestHit(ray.ori, MinRayDistance, ray.dir, MaxRayDistance);
				if (isec.instanceIdx == InvalidHitIndex) break;
				loadSurfaceInfo(s, isec, surf);
				mat = s.materials[surf.matIndex];
				if (bounce == 1 && !surf.isLight) ps.rcIsec = fromIsec(isec);
			}
			if (surf.isLight) {
				float cosTheta = -dot(ray.dir, surf.norm);
				if (bounce > 1 && cosTheta > 0) {
					float weight = 1.0f;
					if (!isSampleTypeDelta(bs.type)) {
						float dist = length(surf.pos - lastPos);
						float sumPower = s.lightTable[0].prob;
						float lightPdf = luminance(surf.albedo) / sumPower * dist * dist / abs_(cosTheta);
						weight = MISWeight(bs.pdf, lightPdf);
					}
					rcLo += surf.albedo * weight * throughputAfter;
				}
				break;
			}
			if (bounce > 0 && !isBSDFDelta(mat)) {
				LightSample ls = sampleLight(s, surf.pos, sample4f(rng));
				bool shadowed = s.traceShadow(surf.pos, MinRayDistance, ls.wi, ls.dist - MinRayDistance);
				if (!shadowed && ls.pdf > 1e-6f) {
					float bsdfPdf = absDot(surf.norm, ls.wi) * PiInv;
					float weight = MISWeight(ls.pdf, bsdfPdf);
					rcLo += ls.radiance * evalBSDF(mat, surf.albedo, surf.norm, wo, ls.wi) * satDot(surf.norm, ls.wi) / ls.pdf * weight * throughputAfter;
				}
			}
			if (bounce > 4) {
				float pdfTerminate = max_(1.0f - luminance(throughputAfter), 0.0f);
				if (sample1f(rng) < pdfTerminate) break;
				throughputAfter /= (1.0f - pdfTerminate);
			}
			vec3 r3 = sample3f(rng);
			if (!sampleBSDF(mat, surf.albedo, surf.norm, wo, r3, bs) || bs.pdf < 1e-6f) break;
			float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
			vec3 scatterTerms = bs.bsdf * cosTheta / bs.pdf;
			if (bounce == 0) {
				primaryScatter = bs.bsdf * cosTheta;
				primaryPdf = bs.pdf;
			}
			else {
				throughputAfter *= scatterTerms;
			}
			lastPos = surf.pos;
			wo = -bs.wi;
			ray.dir = bs.wi;
			ray.ori = surf.pos + ray.dir * 1e-4f;
		}
		vec3 radiance = rcLo * primaryScatter / primaryPdf;
		ps.rcLo[0] = rcLo.x; ps.rcLo[1] = rcLo.y; ps.rcLo[2] = rcLo.z;

		RptGIReservoir resv{};
		giReset(resv);
		resv.rcIsec.instanceIdx = 0;   // zero-initialised local, as `resv{}` above
		if ((f.camera.frameIndex & CameraClearFlag) == 0) {
			Neighbor nb = lookupSurface(f, true, { p.uv.x + motion.x, p.uv.y + motion.y });
			if (nb.found && !(nb.matMeshId != p.matMeshId || dot(nb.norm, p.norm) < 0.9f || abs_(nb.depth - p.depth) > 5.0f)) {
				resv = f.gi[f.cur ^ 1u][nb.pixel];
			}
		}
		if (ps.rcIsec.instanceIdx != InvalidHitIndex) {
			float sampleWeight = luminance(radiance);
			if (isnan_(sampleWeight) || sampleWeight < 0.0f || primaryPdf < 1e-6f) sampleWeight = 0.0f;
			// GIReservoirAddSample (gi_reservoir.glsl:37-44)
			resv.resampleWeight += sampleWeight;
			resv.sampleCount++;
			if (sample1f(rng) * resv.resampleWeight < sampleWeight) std::memcpy(&resv, &ps, 32);
		}
		if (!giValid(resv)) giReset(resv);
		if (resv.sampleCount > 40) {
			resv.resampleWeight *= float(40) / float(resv.sampleCount);
			resv.sampleCount = 40;
		}
		const RptMaterial& primaryMat = s.materials[p.matId];
		if (giValid(resv) && resv.sampleCount > 0 && !isBSDFDelta(primaryMat)) {
			// a reservoir that never received a sample carries rcIsec = 0-initialised / stale data; the
			// reference dereferences it regardless.  instanceIdx 0xffffffff would read out of bounds, so the
			// shade step is skipped for it (radiance keeps the unresampled estimate, as when Li is black).
			if (resv.rcIsec.instanceIdx != InvalidHitIndex) {
				SurfaceInfo rc;
				loadSurfaceInfo(s, toIsec(resv.rcIsec), rc);
				vec3 primaryWi = normalize(rc.pos - primaryPos);
				float weight = resv.resampleWeight / float(resv.sampleCount);
				vec3 Li = V3(resv.rcLo) * evalBSDF(primaryMat, p.albedo, p.norm, primaryWo, primaryWi) * satDot(p.norm, primaryWi);
				if (!isBlack(Li) && s.traceVisibility(primaryPos, rc.pos)) radiance = Li / luminance(Li) * weight;
			}
		}
		f.gi[f.cur][idx] = resv;
		accumulate(f.indirectOutput, f, x, y, clampColor(radiance));
	}
}

// ---------------------------------------------------------------------------------------------------------
// ReSTIR PT / GRIS
// ---------------------------------------------------------------------------------------------------------
namespace {

const float GRISDistanceThreshold = 0.01f;
const uint32_t ShiftReconnection = 0;
const uint32_t RcLightSampled = 0, RcLightScattered = 1, RcSurface = 2;

uint32_t flagsRcVertexId(uint32_t fl) { return fl & 0xffu; }
uint32_t flagsRcVertexType(uint32_t fl) { return (fl >> 16) & 0xffu; }
void flagsSetRcVertexId(uint32_t& fl, uint32_t id) { fl = (fl & 0xffffff00u) | (id & 0xffu); }
void flagsSetPathLength(uint32_t& fl, uint32_t id) { fl = (fl & 0xffff00ffu) | ((id & 0xffu) << 8); }
void flagsSetRcVertexType(uint32_t& fl, uint32_t t) { fl = (fl & 0xff00ffffu) | ((t & 0xffu) << 16); }

void setV3(float* d, vec3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

void grisSampleReset(RptGRISReservoir& r) {   // gris_reservoir.glsl:61-69
	r.rcIsec.instanceIdx = InvalidHitIndex;
	setV3(r.rcLi, V3(0.0f)); setV3(r.rcWi, V3(0.0f));
	r.rcPrevSamplePdf = 0; r.rcJacobian = 0; r.flags = 0;
	setV3(r.F, V3(0.0f));
}
bool grisSampleValid(const RptGRISReservoir& r) { return r.rcIsec.instanceIdx != InvalidHitIndex; }
void grisReset(RptGRISReservoir& r) { r.rcIsec.instanceIdx = InvalidHitIndex; r.sampleCount = 0; r.resampleWeight = 0; }   // :75-79
bool grisValid(const RptGRISReservoir& r) { return !isnan_(r.resampleWeight) && r.resampleWeight >= 0; }   // :93-95
void grisCopySample(RptGRISReservoir& dst, const RptGRISReservoir& src) { std::memcpy(&dst, &src, 80); }
void grisMerge(RptGRISReservoir& resv, const RptGRISReservoir& rhs, float r) {   // :114-123
	resv.sampleCount += rhs.sampleCount;
	resv.resampleWeight += rhs.resampleWeight;
	if (r * resv.resampleWeight < rhs.resampleWeight) grisCopySample(resv, rhs);
}
void grisCap(RptGRISReservoir& resv, float cap) {   // :131-136
	if (resv.sampleCount > cap) {
		resv.resampleWeight *= cap / resv.sampleCount;
		resv.sampleCount = cap;
	}
}

struct GrisStream {   // gris_path_trace.glsl:10-33
	RptGRISReservoir sample{};
	float weight = 0, sumWeight = 0;
	uint32_t sampleCount = 0;
	void add(const RptGRISReservoir& ps, float w, float r) {
		sumWeight += w;
		if (r * sumWeight < w) { weight = w; grisCopySample(sample, ps); sampleCount++; }
	}
};

uint32_t nextRcVertexSampleState(uint32_t state, bool connectible) {   // :35-43
	if (state == 2) return 2;
	if (!connectible) return 0;
	return state + 1u;
}

struct RcData {   // GRISReconnectionData
	Intersection rcPrevIsec{ { 0, 0 }, InvalidHitIndex, 0 };
	vec3 rcPrevWo{ 0, 0, 0 };
	vec3 rcPrevThroughput{ 0, 0, 0 };
};

// gris_retrace.glsl:42-136
void traceReplayPath(const Scene& s, const RptGRISSettings& st, Intersection isec, SurfaceInfo surf, Ray ray,
                     uint32_t targetFlags, uint32_t rng, RcData& rc) {
	vec3 throughput = V3(1.0f);
	vec3 wo = -ray.dir;
	RptMaterial mat = s.materials[surf.matIndex];
	BSDFSample bs;
	rc = RcData{};
	uint32_t targetId = flagsRcVertexId(targetFlags);
	if (targetId == 1) {
		rc.rcPrevIsec = isec; rc.rcPrevWo = wo; rc.rcPrevThroughput = throughput;
		return;
	}
	for (int bounce = 0; bounce < MaxTracingDepth; bounce++) {
		if (bounce > 0) {
			isec = s.traceClosestHit(ray.ori, MinRayDistance, ray.dir, MaxRayDistance);
			if (isec.instanceIdx == InvalidHitIndex) break;
			loadSurfaceInfo(s, isec, surf);
			mat = s.materials[surf.matIndex];
		}
		bool isThisVertexConnectible = isBSDFConnectible(mat);
		sample1f(rng);
		if (surf.isLight) break;
		if (uint32_t(bounce) == targetId - 1u) {
			if (isThisVertexConnectible) {
				rc.rcPrevIsec = isec; rc.rcPrevWo = wo; rc.rcPrevThroughput = throughput;
			}
			break;
		}
		sample4f(rng);   // keep the random stream aligned with tracePath
		sample1f(rng);
		if (bounce > 4) {
			float pdfTerminate = max_(1.0f - luminance(throughput) * st.rrScale, 0.0f);
			if (sample1f(rng) < pdfTerminate) break;
			throughput /= (1.0f - pdfTerminate);
		}
		vec3 r3 = sample3f(rng);
		if (!sampleBSDF(mat, surf.albedo, surf.norm, wo, r3, bs) || bs.pdf < 1e-6f) break;
		float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
		throughput *= bs.bsdf * cosTheta / bs.pdf;
		wo = -bs.wi;
		ray.dir = bs.wi;
		ray.ori = surf.pos + ray.dir * 1e-4f;
	}
}

// gris_retrace.glsl:138-236
void grisReuseAndMerge(const Scene& s, const RptGRISSettings& st, RptGRISReservoir& dst, const SurfaceInfo& dstPrimarySurf,
                       const Intersection& dstPrimaryIsec, const Ray& primaryRay, RptGRISReservoir src, uint32_t& rng) {
	RcData rcData;
	SurfaceInfo rcPrevSurf, rcSurf;
	RptMaterial rcMat{}, rcPrevMat{};
	vec3 wi = V3(0.0f), Li = V3(0.0f);
	bool srcSampleValid = false;
	float dstJacobian = 0, jacobian = 0, dstPHat = 0, dstSamplePdf = 0;

	if (grisSampleValid(src)) {
		traceReplayPath(s, st, dstPrimaryIsec, dstPrimarySurf, primaryRay, src.flags, src.primaryRng, rcData);
		if (rcData.rcPrevIsec.instanceIdx != InvalidHitIndex) {
			if (rcData.rcPrevIsec.instanceIdx == SpecialHitIndex) rcPrevSurf = dstPrimarySurf;
			else loadSurfaceInfo(s, rcData.rcPrevIsec, rcPrevSurf);
			loadSurfaceInfo(s, toIsec(src.rcIsec), rcSurf);
			rcMat = s.materials[rcSurf.matIndex];
			rcPrevMat = s.materials[rcPrevSurf.matIndex];
			float dist = distance(rcPrevSurf.pos, rcSurf.pos);
			wi = normalize(rcSurf.pos - rcPrevSurf.pos);
			float cosTheta = -dot(rcSurf.norm, wi);
			dstJacobian = abs_(cosTheta) / square(dist);
			jacobian = dstJacobian / src.rcJacobian;
			if (dist > GRISDistanceThreshold && cosTheta > 0 && !isnan_(jacobian) && src.rcJacobian > 0 && isBSDFConnectible(rcPrevMat)) {
				if (s.traceVisibility(rcPrevSurf.pos, rcSurf.pos)) srcSampleValid = true;
			}
		}
	}
	if (srcSampleValid) {
		uint32_t rcType = flagsRcVertexType(src.flags);
		if (!isnan_(src.rcPrevSamplePdf) && src.rcPrevSamplePdf > 1e-6f) {
			Li = V3(src.rcLi);
			vec3 rcWi = V3(src.rcWi);
			if (rcType == RcSurface && length(rcWi) > 0.5f) {
				Li *= evalBSDF(rcMat, rcSurf.albedo, rcSurf.norm, -wi, rcWi) * satDot(rcSurf.norm, rcWi);
			}
			Li *= evalBSDF(rcPrevMat, rcPrevSurf.albedo, rcPrevSurf.norm, rcData.rcPrevWo, wi) * satDot(rcPrevSurf.norm, wi);
			Li *= rcData.rcPrevThroughput;
			Li /= src.rcPrevSamplePdf;
			if (!isBlack(Li) && !hasNan(Li)) dstPHat = luminance(Li * jacobian);
			if (rcType == RcLightSampled) {
				float sumPower = s.lightTable[0].prob;
				dstSamplePdf = luminance(rcSurf.albedo) / sumPower / dstJacobian;
			}
			else {
				dstSamplePdf = evalPdf(rcPrevMat, rcPrevSurf.norm, rcData.rcPrevWo, wi);
			}
		}
		float srcPHat = luminance(V3(src.F));
		src.rcJacobian = dstJacobian;
		src.rcPrevSamplePdf = dstSamplePdf;
		setV3(src.F, Li);
		if (src.rcPrevSamplePdf < 1e-6f || isnan_(src.rcPrevSamplePdf)) src.rcPrevSamplePdf = 0;
		src.resampleWeight *= dstPHat / srcPHat;
	}
	else {
		src.resampleWeight = 0;
	}
	if (grisValid(src)) grisMerge(dst, src, sample1f(rng));
	grisCap(dst, float(st.cap));
}

} // namespace

// gris_path_trace.glsl:45-305 (the .comp's image store is commented out: only the reservoir is written)
void passGRISPathTrace(const Scene& s, Frame2D& f, const RptGRISSettings& st, uint32_t y0, uint32_t y1) {
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++) {
		Primary p = loadPrimary(f, x, y);
		if (!p.valid) continue;
		Ray ray = p.ray;
		uint32_t rng = makeSeed(f.camera.seed, uvec2{ x, y });
		vec3 throughput = V3(1.0f), rcThroughput = V3(0.0f), lastPos = V3(0.0f);
		vec3 wo = -ray.dir;
		bool isLastVertexConnectible = false;
		uint32_t sampleState = 0, lastSampleState = 0;
		SurfaceInfo surf;
		surf.pos = p.pos; surf.norm = p.norm; surf.albedo = p.albedo; surf.isLight = false;
		RptMaterial mat = s.materials[p.matId];
		BSDFSample bs;
		Intersection isec{ { 0, 0 }, 0, 0 };

		RptGRISReservoir ps{};   // GRISPathSample part of a reservoir
		grisSampleReset(ps);
		ps.primaryRng = rng;
		RptGRISReservoir resv{};
		grisReset(resv);
		GrisStream stream;

		for (int bounce = 0; bounce < MaxTracingDepth; bounce++) {
			if (bounce > 0) {
				isec = s.traceClosestHit(ray.ori, MinRayDistance, ray.dir, MaxRayDistance);
				if (isec.instanceIdx == InvalidHitIndex) break;
				loadSurfaceInfo(s, isec, surf);
				mat = s.materials[surf.matIndex];
			}
			flagsSetPathLength(ps.flags, uint32_t(bounce + 1));
			float cosPrevWi = dot(ray.dir, surf.norm);
			float distToPrev = distance(lastPos, surf.pos);
			float geometryJacobian = abs_(cosPrevWi) / square(distToPrev);
			bool isThisVertexConnectible = surf.isLight || isBSDFConnectible(mat);
			lastSampleState = sampleState;
			sampleState = nextRcVertexSampleState(sampleState, isThisVertexConnectible);
			if (st.shiftType == ShiftReconnection && bounce == 1 && !surf.isLight) {
				sampleState = 2;
				lastSampleState = 1;
			}
			float resvRandSample = sample1f(rng);

			if (surf.isLight) {
				if (bounce > 1 && cosPrevWi < 0) {
					float weight = 1.0f;
					float sumPower = s.lightTable[0].prob;
					float lightPdf = luminance(surf.albedo) / sumPower / geometryJacobian;
					if (!isSampleTypeDelta(bs.type)) weight = MISWeight(bs.pdf, lightPdf);
					vec3 weightedLi = surf.albedo * weight;
					if (sampleState == 2 && lastSampleState == 2) {
						setV3(ps.rcLi, V3(ps.rcLi) + weightedLi * rcThroughput);
						setV3(ps.F, V3(ps.F) + weightedLi * throughput);
					}
					else if ((sampleState == 2 && lastSampleState == 1) && isLastVertexConnectible && distToPrev > GRISDistanceThreshold) {
						ps.rcIsec = fromIsec(isec);
						ps.rcRng = rng;
						ps.rcPrevSamplePdf = bs.pdf;
						ps.rcJacobian = geometryJacobian;
						setV3(ps.rcLi, weightedLi);
						setV3(ps.rcWi, V3(0.0f));
						setV3(ps.F, weightedLi * throughput);
						flagsSetRcVertexId(ps.flags, uint32_t(bounce));
						flagsSetRcVertexType(ps.flags, RcLightScattered);
						stream.add(ps, luminance(V3(ps.F)), resvRandSample);
					}
				}
				break;
			}
			bool connectible = isThisVertexConnectible && isLastVertexConnectible && distToPrev > GRISDistanceThreshold;
			if ((sampleState == 2 && lastSampleState == 1) && (connectible || st.shiftType == ShiftReconnection)) {
				ps.rcIsec = fromIsec(isec);
				ps.rcRng = rng;
				ps.rcPrevSamplePdf = bs.pdf;
				ps.rcJacobian = geometryJacobian;
				flagsSetRcVertexId(ps.flags, uint32_t(bounce));
				flagsSetRcVertexType(ps.flags, RcSurface);
				rcThroughput = V3(1.0f);
			}
			vec4 lightRandSample = sample4f(rng);
			resvRandSample = sample1f(rng);

			if (bounce > 0 && !isBSDFDelta(mat)) {
				LightSample ls = sampleLight(s, surf.pos, lightRandSample);
				bool shadowed = s.traceShadow(surf.pos, MinRayDistance, ls.wi, ls.dist - MinRayDistance);
				if (!shadowed && ls.pdf > 1e-6f) {
					float bsdfPdf = absDot(surf.norm, ls.wi) * PiInv;
					float weight = MISWeight(ls.pdf, bsdfPdf);
					vec3 scatterTerm = evalBSDF(mat, surf.albedo, surf.norm, wo, ls.wi) * satDot(surf.norm, ls.wi);
					vec3 weightedLi = ls.radiance / ls.pdf * weight;
					if (sampleState == 2 && lastSampleState == 2) {
						setV3(ps.rcLi, V3(ps.rcLi) + weightedLi * scatterTerm * rcThroughput);
						setV3(ps.F, V3(ps.F) + weightedLi * scatterTerm * throughput);
					}
					else if (sampleState == 2 && lastSampleState == 1) {
						setV3(ps.rcLi, weightedLi);
						setV3(ps.rcWi, ls.wi);
						setV3(ps.F, weightedLi * scatterTerm * throughput);
						stream.add(ps, luminance(V3(ps.F)), resvRandSample);
					}
					else if (sampleState == 1 && isThisVertexConnectible && ls.dist > GRISDistanceThreshold) {
						ps.rcIsec = { { ls.bary.x, ls.bary.y }, 0, ls.id };
						ps.rcRng = rng;
						ps.rcPrevSamplePdf = ls.pdf;
						ps.rcJacobian = ls.jacobian;
						setV3(ps.rcLi, ls.radiance * weight);
						setV3(ps.rcWi, V3(0.0f));
						setV3(ps.F, weightedLi * scatterTerm * throughput);
						flagsSetRcVertexId(ps.flags, uint32_t(bounce + 1));
						flagsSetRcVertexType(ps.flags, RcLightSampled);
						stream.add(ps, luminance(V3(ps.F)), resvRandSample);
					}
				}
			}
			if (bounce > 4) {
				float pdfTerminate = max_(1.0f - luminance(throughput) * st.rrScale, 0.0f);
				if (sample1f(rng) < pdfTerminate) break;
				throughput /= (1.0f - pdfTerminate);
				rcThroughput /= (1.0f - pdfTerminate);
			}
			vec3 r3 = sample3f(rng);
			if (!sampleBSDF(mat, surf.albedo, surf.norm, wo, r3, bs) || bs.pdf < 1e-6f) break;
			float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
			vec3 scatterTerm = bs.bsdf * cosTheta / bs.pdf;
			throughput *= scatterTerm;
			if (sampleState == 2 && lastSampleState == 2) {
				rcThroughput *= scatterTerm;
			}
			else if (sampleState == 2 && lastSampleState == 1) {
				setV3(ps.rcLi, V3(0.0f));
				setV3(ps.rcWi, bs.wi);
				setV3(ps.F, V3(0.0f));
				rcThroughput /= bs.pdf;
			}
			lastPos = surf.pos;
			wo = -bs.wi;
			ray.dir = bs.wi;
			ray.ori = surf.pos + ray.dir * 1e-4f;
			isLastVertexConnectible = isThisVertexConnectible;
		}
		if (sampleState == 2 && lastSampleState == 2) {
			stream.add(ps, luminance(V3(ps.F)), sample1f(rng));
		}
		grisCopySample(resv, stream.sample);
		if (stream.sumWeight > 0 && stream.weight > 0) {
			float k = stream.sumWeight / stream.weight;
			setV3(resv.F, V3(resv.F) * k);
			setV3(resv.rcLi, V3(resv.rcLi) * k);
			resv.resampleWeight = luminance(V3(resv.F));
		}
		else {
			resv.rcIsec.instanceIdx = InvalidHitIndex;
			setV3(resv.F, V3(0.0f));
		}
		resv.sampleCount = 1;
		f.gris[f.cur][size_t(y) * f.width + x] = resv;
	}
}

// gris_resample_temporal.glsl:36-83
void passGRISTemporal(const Scene& s, Frame2D& f, const RptGRISSettings& st, uint32_t y0, uint32_t y1) {
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++) {
		Primary p = loadPrimary(f, x, y);
		if (!p.valid) continue;
		size_t idx = size_t(y) * f.width + x;
		vec2 motion = f.motion[idx];
		uint32_t rng = makeSeed(f.camera.seed, uvec2{ x, y }) ^ 1u;
		uint32_t resvRng = ~rng;
		RptGRISReservoir resv = f.gris[f.cur][idx];
		Intersection dstPrimaryIsec{ p.uv, SpecialHitIndex, 0 };

		if (st.temporalReuse) {
			SurfaceInfo dstPrimarySurf = primarySurface(p);
			if ((f.camera.frameIndex & CameraClearFlag) == 0) {
				Neighbor nb = lookupSurface(f, true, { p.uv.x + motion.x, p.uv.y + motion.y });
				if (nb.found && !(nb.matMeshId != p.matMeshId || dot(nb.norm, p.norm) < 0.95f || distance(p.pos, nb.pos) > 0.5f)) {
					const RptGRISReservoir& prev = f.gris[f.cur ^ 1u][nb.pixel];
					if (grisValid(prev)) grisReuseAndMerge(s, st, resv, dstPrimarySurf, dstPrimaryIsec, p.ray, prev, resvRng);
				}
			}
		}
		if (!grisValid(resv)) grisReset(resv);
		f.grisTemp[idx] = resv;
	}
}

// gris_resample_spatial.glsl:36-134 + .comp
void passGRISSpatial(const Scene& s, Frame2D& f, const RptGRISSettings& st, uint32_t y0, uint32_t y1) {
	const float texelX = 1.0f / float(f.width), texelY = 1.0f / float(f.height);
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++) {
		Primary p = loadPrimary(f, x, y);
		vec3 radiance = V3(0.0f);
		if (p.valid) {
			size_t idx = size_t(y) * f.width + x;
			uint32_t rng = makeSeed(f.camera.seed, uvec2{ x, y }) ^ 2u;
			RptGRISReservoir resv = f.grisTemp[idx];
			Intersection dstPrimaryIsec{ p.uv, SpecialHitIndex, 0 };
			SurfaceInfo dstPrimarySurf = primarySurface(p);

			if (st.spatialReuse) {
				for (uint32_t i = 0; i < 3; i++) {
					vec2 d = toConcentricDisk(sample2f(rng));
					vec2 nuv = { p.uv.x + d.x * 20.0f * texelX, p.uv.y + d.y * 20.0f * texelY };
					Neighbor nb = lookupSurface(f, false, nuv);
					if (nb.found && !(dot(nb.norm, p.norm) < 0.9f || distance(p.pos, nb.pos) > 0.4f)) {
						const RptGRISReservoir& nr = f.grisTemp[nb.pixel];
						if (grisValid(nr)) grisReuseAndMerge(s, st, resv, dstPrimarySurf, dstPrimaryIsec, p.ray, nr, rng);
					}
				}
			}
			if (!grisValid(resv)) grisReset(resv);
			f.gris[f.cur][idx] = resv;

			if (grisValid(resv) && resv.sampleCount > 0 && grisSampleValid(resv)) {
				RcData rcData;
				traceReplayPath(s, st, dstPrimaryIsec, dstPrimarySurf, p.ray, resv.flags, resv.primaryRng, rcData);
				if (rcData.rcPrevIsec.instanceIdx != InvalidHitIndex) {
					SurfaceInfo rcPrevSurf, rcSurf;
					if (rcData.rcPrevIsec.instanceIdx == SpecialHitIndex) rcPrevSurf = dstPrimarySurf;
					else loadSurfaceInfo(s, rcData.rcPrevIsec, rcPrevSurf);
					loadSurfaceInfo(s, toIsec(resv.rcIsec), rcSurf);
					const RptMaterial& rcMat = s.materials[rcSurf.matIndex];
					const RptMaterial& rcPrevMat = s.materials[rcPrevSurf.matIndex];
					vec3 Li = V3(resv.rcLi);
					uint32_t rcType = flagsRcVertexType(resv.flags);
					vec3 wi = normalize(rcSurf.pos - rcPrevSurf.pos);
					vec3 rcWi = V3(resv.rcWi);
					if (rcType == RcSurface && length(rcWi) > 0.5f) {
						Li *= evalBSDF(rcMat, rcSurf.albedo, rcSurf.norm, -wi, rcWi) * satDot(rcSurf.norm, rcWi);
					}
					Li *= evalBSDF(rcPrevMat, rcPrevSurf.albedo, rcPrevSurf.norm, rcData.rcPrevWo, wi) * satDot(rcPrevSurf.norm, wi);
					Li *= rcData.rcPrevThroughput;
					Li /= resv.rcPrevSamplePdf;
					if (!isBlack(Li) && !hasNan(Li)) radiance = Li / luminance(Li) * resv.resampleWeight / resv.sampleCount;
				}
			}
			radiance = clampColor(radiance);
		}
		accumulate(f.indirectOutput, f, x, y, radiance);
	}
}

// as_visualize.comp:14-27 — candidates = triangles the ray intersects within [tmin, tmax] (nothing is ever
// committed in debugVisualizeAS, ray_query.glsl:72-92, so the interval never shrinks)
void passVisualizeAS(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1) {
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++) {
		vec2 uv = { (float(x) + 0.5f) / float(f.width), (float(y) + 0.5f) / float(f.height) };
		Ray ray = pinholeCameraSampleRay(f.camera, { uv.x, 1.0f - uv.y });
		float level = float(s.countCandidates(ray.ori, ray.dir)) / 100.0f;
		f.directOutput[size_t(y) * f.width + x] = { level, level, level, 1.0f };
	}
}

// post_proc.frag:16-42 (quad UVs map 1:1 to pixels)
void passPostProcess(const Frame2D& f, const RptPostSettings& st, uint8_t* rgba8, uint32_t y0, uint32_t y1) {
	auto filmic1 = [](float c) { return (c * (c * 0.22f + 0.03f) + 0.002f) / (c * (c * 0.22f + 0.3f) + 0.06f) - 1.0f / 30.0f; };
	for (uint32_t y = y0; y < y1; y++) for (uint32_t x = 0; x < f.width; x++) {
		size_t i = size_t(y) * f.width + x;
		float c[3] = { 0, 0, 0 };
		if (st.noDirect == 0) { c[0] += f.directOutput[i].x; c[1] += f.directOutput[i].y; c[2] += f.directOutput[i].z; }
		if (st.noIndirect == 0) { c[0] += f.indirectOutput[i].x; c[1] += f.indirectOutput[i].y; c[2] += f.indirectOutput[i].z; }
		for (int k = 0; k < 3; k++) {
			float v = c[k];
			if (st.toneMapping == 1) v = filmic1(v * 1.6f) / filmic1(11.2f);
			else if (st.toneMapping == 2) v = (v * (v * 2.51f + 0.03f)) / (v * (v * 2.43f + 0.59f) + 0.14f);
			if (st.correctGamma != 0) v = std::pow(v, 1.0f / 2.2f);
			// UNORM8 colour attachment: clamp and round to nearest
			v = v != v ? 0.0f : clamp_(v, 0.0f, 1.0f);
			rgba8[i * 4 + k] = uint8_t(std::floor(v * 255.0f + 0.5f));
		}
		rgba8[i * 4 + 3] = 255;
	}
}

} // namespace orc
