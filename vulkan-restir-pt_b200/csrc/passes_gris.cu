// ReSTIR PT / GRIS (BASELINE.json configs 3 and 4 — the north-star path): candidate path generation with
// reconnection-vertex selection, hybrid-shift temporal reuse, hybrid-shift spatial reuse + final shading.
//   reference src/shader/gris_path_trace.glsl:45-305, gris_retrace.glsl:42-236, gris_reservoir.glsl:37-136,
//   gris_resample_temporal.glsl:11-83, gris_resample_spatial.glsl:11-134 (+ the three .comp entry points)
//   host sequence: GRISReSTIR::render (src/GRISReSTIR.cpp:9-53)
#include "passes.h"
#include "shading.cuh"

namespace rt {

namespace {

constexpr float GRISDistanceThreshold = 0.01f;
constexpr uint32_t ShiftReconnection = 0;
constexpr uint32_t RcLightSampled = 0, RcLightScattered = 1, RcSurface = 2;

// GRISReservoir (96 B) as six 16-byte words; q0..q4 are the GRISPathSample:
//   q0 rcIsec | q1 {rcLi, rcRng} | q2 {rcWi, flags} | q3 {pad, pad, rcPrevSamplePdf, rcJacobian} | q4 {F, primaryRng}
//   q5 {sampleCount (float), resampleWeight, contribWeight, pad}
struct GRISResv {
	float4 q0, q1, q2, q3, q4, q5;
	RT_DEV uint32_t rcInstance() const { return __float_as_uint(q0.z); }
	RT_DEV uint32_t flags() const { return __float_as_uint(q2.w); }
	RT_DEV void setFlags(uint32_t v) { q2.w = __uint_as_float(v); }
	RT_DEV float3 rcLi() const { return f3(q1); }
	RT_DEV void setRcLi(float3 v) { q1.x = v.x; q1.y = v.y; q1.z = v.z; }
	RT_DEV float3 rcWi() const { return f3(q2); }
	RT_DEV void setRcWi(float3 v) { q2.x = v.x; q2.y = v.y; q2.z = v.z; }
	RT_DEV float3 F() const { return f3(q4); }
	RT_DEV void setF(float3 v) { q4.x = v.x; q4.y = v.y; q4.z = v.z; }
	RT_DEV float& rcPrevSamplePdf() { return q3.z; }
	RT_DEV float& rcJacobian() { return q3.w; }
	RT_DEV uint32_t primaryRng() const { return __float_as_uint(q4.w); }
	RT_DEV float& sampleCount() { return q5.x; }
	RT_DEV float& resampleWeight() { return q5.y; }
	RT_DEV bool sampleValid() const { return rcInstance() != InvalidHitIndex; }
	RT_DEV bool valid() const { return !isnan_(q5.y) && q5.y >= 0; }                       // gris_reservoir.glsl:93-95
	RT_DEV void reset() { q0.z = __uint_as_float(InvalidHitIndex); q5.x = 0; q5.y = 0; }   // :75-79
	RT_DEV void copySample(const GRISResv& o) { q0 = o.q0; q1 = o.q1; q2 = o.q2; q3 = o.q3; q4 = o.q4; }
};

RT_DEV GRISResv zeroGRIS() {
	GRISResv r;
	r.q0 = r.q1 = r.q2 = r.q3 = r.q4 = r.q5 = make_float4(0.f, 0.f, 0.f, 0.f);
	return r;
}
RT_DEV GRISResv loadGRIS(const RptGRISReservoir* p) {
	const float4* q = reinterpret_cast<const float4*>(p);
	GRISResv r; r.q0 = q[0]; r.q1 = q[1]; r.q2 = q[2]; r.q3 = q[3]; r.q4 = q[4]; r.q5 = q[5];
	return r;
}
RT_DEV void storeGRIS(RptGRISReservoir* p, const GRISResv& r) {
	float4* q = reinterpret_cast<float4*>(p);
	q[0] = r.q0; q[1] = r.q1; q[2] = r.q2; q[3] = r.q3; q[4] = r.q4; q[5] = r.q5;
}

RT_DEV uint32_t flagsRcVertexId(uint32_t fl) { return fl & 0xffu; }
RT_DEV uint32_t flagsRcVertexType(uint32_t fl) { return (fl >> 16) & 0xffu; }
RT_DEV uint32_t withRcVertexId(uint32_t fl, uint32_t id) { return (fl & 0xffffff00u) | (id & 0xffu); }
RT_DEV uint32_t withPathLength(uint32_t fl, uint32_t id) { return (fl & 0xffff00ffu) | ((id & 0xffu) << 8); }
RT_DEV uint32_t withRcVertexType(uint32_t fl, uint32_t t) { return (fl & 0xff00ffffu) | ((t & 0xffu) << 16); }

RT_DEV void grisMerge(GRISResv& resv, GRISResv& rhs, float r) {   // gris_reservoir.glsl:114-123
	resv.sampleCount() += rhs.sampleCount();
	resv.resampleWeight() += rhs.resampleWeight();
	if (r * resv.resampleWeight() < rhs.resampleWeight()) resv.copySample(rhs);
}
RT_DEV void grisCap(GRISResv& resv, float cap) {   // :131-136
	if (resv.sampleCount() > cap) {
		resv.resampleWeight() *= cap / resv.sampleCount();
		resv.sampleCount() = cap;
	}
}

struct GrisStream {   // gris_path_trace.glsl:10-33
	GRISResv sample;
	float weight, sumWeight;
	RT_DEV void add(const GRISResv& ps, float w, float r) {
		sumWeight += w;
		if (r * sumWeight < w) { weight = w; sample.copySample(ps); }
	}
};

RT_DEV uint32_t nextRcVertexSampleState(uint32_t state, bool connectible) {   // :35-43
	if (state == 2) return 2;
	if (!connectible) return 0;
	return state + 1u;
}

struct RcData {   // GRISReconnectionData (layouts.glsl:158-164), only ever a local
	uint32_t prevInstance, prevTriangle;
	float2 prevBary;
	float3 rcPrevWo, rcPrevThroughput;
};

// gris_retrace.glsl:42-136: replay the BSDF chain from the destination's primary hit with the source path's
// random numbers, consuming them in lock-step with tracePath, up to the vertex before the reconnection vertex
RT_DEV void traceReplayPath(const SceneView& s, const RptGRISSettings& st, const Surface& primarySurf, float2 primaryUv, Ray ray,
                            uint32_t targetFlags, uint32_t rng, RcData& rc) {
	float3 throughput = f3(1.0f);
	float3 wo = -ray.dir;
	Surface surf = primarySurf;
	Mat mat = loadMaterial(s, surf.matIndex);
	BSDFSample bs = emptyBSDFSample();
	rc.prevInstance = InvalidHitIndex; rc.prevTriangle = 0; rc.prevBary = make_float2(0.f, 0.f);
	rc.rcPrevWo = f3(0.0f); rc.rcPrevThroughput = f3(0.0f);
	uint32_t curInst = SpecialHitIndex, curTri = 0;
	float2 curBary = primaryUv;
	const uint32_t targetId = flagsRcVertexId(targetFlags);
	if (targetId == 1) {
		rc.prevInstance = curInst; rc.prevTriangle = curTri; rc.prevBary = curBary;
		rc.rcPrevWo = wo; rc.rcPrevThroughput = throughput;
		return;
	}
	for (int bounce = 0; bounce < 15; bounce++) {
		if (bounce > 0) {
			const Hit h = traceClosestHit(s, ray.ori, MinRayDistance, ray.dir, MaxRayDistance);
			if (h.instanceIdx == InvalidHitIndex) break;
			curInst = h.instanceIdx; curTri = h.triangleIdx; curBary = make_float2(h.u, h.v);
			loadSurfaceInfo(s, h, surf);
			mat = loadMaterial(s, surf.matIndex);
		}
		const bool isThisVertexConnectible = isBSDFConnectible(mat);
		sample1f(rng);
		if (surf.isLight) break;
		if (uint32_t(bounce) == targetId - 1u) {
			if (isThisVertexConnectible) {
				rc.prevInstance = curInst; rc.prevTriangle = curTri; rc.prevBary = curBary;
				rc.rcPrevWo = wo; rc.rcPrevThroughput = throughput;
			}
			break;
		}
		sample4f(rng);
		sample1f(rng);
		if (bounce > 4) {
			const float pdfTerminate = max_(1.0f - luminance(throughput) * st.rrScale, 0.0f);
			if (sample1f(rng) < pdfTerminate) break;
			throughput /= (1.0f - pdfTerminate);
		}
		const float3 r3 = sample3f(rng);
		if (!sampleBSDF(mat, surf.albedo, surf.norm, wo, r3, bs) || bs.pdf < 1e-6f) break;
		const float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
		throughput *= bs.bsdf * cosTheta / bs.pdf;
		wo = -bs.wi;
		ray.dir = bs.wi;
		ray.ori = surf.pos + ray.dir * 1e-4f;
	}
}

// shared by the shift (Li of the shifted path) and the final shading of the spatial pass
RT_DEV float3 reconnectionLi(const SceneView& s, const GRISResv& sample, const RcData& rc, const Surface& rcPrevSurf, const Surface& rcSurf,
                             const Mat& rcPrevMat, float3 wi, float rcPrevSamplePdf) {
	float3 Li = sample.rcLi();
	const uint32_t rcType = flagsRcVertexType(sample.flags());
	const float3 rcWi = sample.rcWi();
	if (rcType == RcSurface && length(rcWi) > 0.5f) {
		const Mat rcMat = loadMaterial(s, rcSurf.matIndex);
		Li *= evalBSDF(rcMat, rcSurf.albedo, rcSurf.norm, -wi, rcWi) * satDot(rcSurf.norm, rcWi);
	}
	Li *= evalBSDF(rcPrevMat, rcPrevSurf.albedo, rcPrevSurf.norm, rc.rcPrevWo, wi) * satDot(rcPrevSurf.norm, wi);
	Li *= rc.rcPrevThroughput;
	Li /= rcPrevSamplePdf;
	return Li;
}

// gris_retrace.glsl:138-236
RT_DEV void grisReuseAndMerge(const SceneView& s, const RptGRISSettings& st, GRISResv& dst, const Surface& dstPrimarySurf, float2 dstUv,
                              const Ray& primaryRay, GRISResv src, uint32_t& rng) {
	RcData rc;
	Surface rcPrevSurf, rcSurf;
	Mat rcPrevMat;
	float3 wi = f3(0.0f), Li = f3(0.0f);
	bool srcSampleValid = false;
	float dstJacobian = 0, jacobian = 0, dstPHat = 0, dstSamplePdf = 0;

	if (src.sampleValid()) {
		traceReplayPath(s, st, dstPrimarySurf, dstUv, primaryRay, src.flags(), src.primaryRng(), rc);
		if (rc.prevInstance != InvalidHitIndex) {
			if (rc.prevInstance == SpecialHitIndex) rcPrevSurf = dstPrimarySurf;
			else loadSurfaceInfo(s, rc.prevInstance, rc.prevTriangle, rc.prevBary, rcPrevSurf);
			loadSurfaceInfo(s, src.rcInstance(), __float_as_uint(src.q0.w), make_float2(src.q0.x, src.q0.y), rcSurf);
			rcPrevMat = loadMaterial(s, rcPrevSurf.matIndex);
			const float dist = distance(rcPrevSurf.pos, rcSurf.pos);
			wi = normalize(rcSurf.pos - rcPrevSurf.pos);
			const float cosTheta = -dot(rcSurf.norm, wi);
			dstJacobian = abs_(cosTheta) / square(dist);
			jacobian = dstJacobian / src.rcJacobian();
			if (dist > GRISDistanceThreshold && cosTheta > 0 && !isnan_(jacobian) && src.rcJacobian() > 0 && isBSDFConnectible(rcPrevMat)) {
				if (traceVisibility(s, rcPrevSurf.pos, rcSurf.pos)) srcSampleValid = true;
			}
		}
	}
	if (srcSampleValid) {
		const uint32_t rcType = flagsRcVertexType(src.flags());
		if (!isnan_(src.rcPrevSamplePdf()) && src.rcPrevSamplePdf() > 1e-6f) {
			Li = reconnectionLi(s, src, rc, rcPrevSurf, rcSurf, rcPrevMat, wi, src.rcPrevSamplePdf());
			if (!isBlack(Li) && !hasNan(Li)) dstPHat = luminance(Li * jacobian);
			if (rcType == RcLightSampled) {
				const float sumPower = s.lightTable[0].prob;
				dstSamplePdf = luminance(rcSurf.albedo) / sumPower / dstJacobian;
			}
			else {
				dstSamplePdf = evalPdf(rcPrevMat, rcPrevSurf.norm, rc.rcPrevWo, wi);
			}
		}
		const float srcPHat = luminance(src.F());
		src.rcJacobian() = dstJacobian;
		src.rcPrevSamplePdf() = dstSamplePdf;
		src.setF(Li);
		if (src.rcPrevSamplePdf() < 1e-6f || isnan_(src.rcPrevSamplePdf())) src.rcPrevSamplePdf() = 0;
		src.resampleWeight() *= dstPHat / srcPHat;
	}
	else {
		src.resampleWeight() = 0;
	}
	if (src.valid()) grisMerge(dst, src, sample1f(rng));
	grisCap(dst, float(st.cap));
}

} // namespace

// gris_path_trace.comp -> tracePath
__global__ void __launch_bounds__(PassBlockX* PassBlockY) grisPathTraceKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	if (!p.valid) return;
	Ray ray = p.ray;
	uint32_t rng = makeSeed(f.camera.seed, x, y);
	float3 throughput = f3(1.0f), rcThroughput = f3(0.0f), lastPos = f3(0.0f);
	float3 wo = -ray.dir;
	bool isLastVertexConnectible = false;
	uint32_t sampleState = 0, lastSampleState = 0;
	Surface surf = primarySurface(p);
	Mat mat = loadMaterial(s, uint32_t(p.matId));
	BSDFSample bs = emptyBSDFSample();
	Hit isec;
	isec.u = 0.f; isec.v = 0.f; isec.instanceIdx = 0; isec.triangleIdx = 0;
	const float sumPower = s.lightTable[0].prob;

	GRISResv ps = zeroGRIS();   // GRISPathSampleReset, gris_reservoir.glsl:61-69
	ps.q0.z = __uint_as_float(InvalidHitIndex);
	ps.q4.w = __uint_as_float(rng);   // primaryRng
	GrisStream stream;
	stream.sample = zeroGRIS(); stream.weight = 0.0f; stream.sumWeight = 0.0f;

	for (int bounce = 0; bounce < 15; bounce++) {
		if (bounce > 0) {
			isec = traceClosestHit(s, ray.ori, MinRayDistance, ray.dir, MaxRayDistance);
			if (isec.instanceIdx == InvalidHitIndex) break;
			loadSurfaceInfo(s, isec, surf);
			mat = loadMaterial(s, surf.matIndex);
		}
		ps.setFlags(withPathLength(ps.flags(), uint32_t(bounce + 1)));
		const float cosPrevWi = dot(ray.dir, surf.norm);
		const float distToPrev = distance(lastPos, surf.pos);
		const float geometryJacobian = abs_(cosPrevWi) / square(distToPrev);
		const bool isThisVertexConnectible = surf.isLight || isBSDFConnectible(mat);
		lastSampleState = sampleState;
		sampleState = nextRcVertexSampleState(sampleState, isThisVertexConnectible);
		if (st.shiftType == ShiftReconnection && bounce == 1 && !surf.isLight) {
			sampleState = 2;
			lastSampleState = 1;
		}
		float resvRandSample = sample1f(rng);
		const float4 isecWord = make_float4(isec.u, isec.v, __uint_as_float(isec.instanceIdx), __uint_as_float(isec.triangleIdx));

		if (surf.isLight) {
			if (bounce > 1 && cosPrevWi < 0) {
				float weight = 1.0f;
				const float lightPdf = luminance(surf.albedo) / sumPower / geometryJacobian;
				if (!isSampleTypeDelta(bs.type)) weight = MISWeight(bs.pdf, lightPdf);
				const float3 weightedLi = surf.albedo * weight;
				if (sampleState == 2 && lastSampleState == 2) {
					ps.setRcLi(ps.rcLi() + weightedLi * rcThroughput);
					ps.setF(ps.F() + weightedLi * throughput);
				}
				else if ((sampleState == 2 && lastSampleState == 1) && isLastVertexConnectible && distToPrev > GRISDistanceThreshold) {
					ps.q0 = isecWord;
					ps.q1.w = __uint_as_float(rng);
					ps.rcPrevSamplePdf() = bs.pdf;
					ps.rcJacobian() = geometryJacobian;
					ps.setRcLi(weightedLi);
					ps.setRcWi(f3(0.0f));
					ps.setF(weightedLi * throughput);
					ps.setFlags(withRcVertexType(withRcVertexId(ps.flags(), uint32_t(bounce)), RcLightScattered));
					stream.add(ps, luminance(ps.F()), resvRandSample);
				}
			}
			break;
		}
		const bool connectible = isThisVertexConnectible && isLastVertexConnectible && distToPrev > GRISDistanceThreshold;
		if ((sampleState == 2 && lastSampleState == 1) && (connectible || st.shiftType == ShiftReconnection)) {
			ps.q0 = isecWord;
			ps.q1.w = __uint_as_float(rng);
			ps.rcPrevSamplePdf() = bs.pdf;
			ps.rcJacobian() = geometryJacobian;
			ps.setFlags(withRcVertexType(withRcVertexId(ps.flags(), uint32_t(bounce)), RcSurface));
			rcThroughput = f3(1.0f);
		}
		const float4 lightRandSample = sample4f(rng);
		resvRandSample = sample1f(rng);

		if (bounce > 0 && !isBSDFDelta(mat)) {
			const LightSample ls = sampleLight(s, surf.pos, lightRandSample);
			const bool shadowed = traceShadow(s, surf.pos, MinRayDistance, ls.wi, ls.dist - MinRayDistance);
			if (!shadowed && ls.pdf > 1e-6f) {
				const float bsdfPdf = absDot(surf.norm, ls.wi) * RT_PI_INV;
				const float weight = MISWeight(ls.pdf, bsdfPdf);
				const float3 scatterTerm = evalBSDF(mat, surf.albedo, surf.norm, wo, ls.wi) * satDot(surf.norm, ls.wi);
				const float3 weightedLi = ls.radiance / ls.pdf * weight;
				if (sampleState == 2 && lastSampleState == 2) {
					ps.setRcLi(ps.rcLi() + weightedLi * scatterTerm * rcThroughput);
					ps.setF(ps.F() + weightedLi * scatterTerm * throughput);
				}
				else if (sampleState == 2 && lastSampleState == 1) {
					ps.setRcLi(weightedLi);
					ps.setRcWi(ls.wi);
					ps.setF(weightedLi * scatterTerm * throughput);
					stream.add(ps, luminance(ps.F()), resvRandSample);
				}
				else if (sampleState == 1 && isThisVertexConnectible && ls.dist > GRISDistanceThreshold) {
					ps.q0 = make_float4(ls.bary.x, ls.bary.y, __uint_as_float(0u), __uint_as_float(ls.id));
					ps.q1.w = __uint_as_float(rng);
					ps.rcPrevSamplePdf() = ls.pdf;
					ps.rcJacobian() = ls.jacobian;
					ps.setRcLi(ls.radiance * weight);
					ps.setRcWi(f3(0.0f));
					ps.setF(weightedLi * scatterTerm * throughput);
					ps.setFlags(withRcVertexType(withRcVertexId(ps.flags(), uint32_t(bounce + 1)), RcLightSampled));
					stream.add(ps, luminance(ps.F()), resvRandSample);
				}
			}
		}
		if (bounce > 4) {
			const float pdfTerminate = max_(1.0f - luminance(throughput) * st.rrScale, 0.0f);
			if (sample1f(rng) < pdfTerminate) break;
			throughput /= (1.0f - pdfTerminate);
			rcThroughput /= (1.0f - pdfTerminate);
		}
		const float3 r3 = sample3f(rng);
		if (!sampleBSDF(mat, surf.albedo, surf.norm, wo, r3, bs) || bs.pdf < 1e-6f) break;
		const float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
		const float3 scatterTerm = bs.bsdf * cosTheta / bs.pdf;
		throughput *= scatterTerm;
		if (sampleState == 2 && lastSampleState == 2) {
			rcThroughput *= scatterTerm;
		}
		else if (sampleState == 2 && lastSampleState == 1) {
			ps.setRcLi(f3(0.0f));
			ps.setRcWi(bs.wi);
			ps.setF(f3(0.0f));
			rcThroughput /= bs.pdf;
		}
		lastPos = surf.pos;
		wo = -bs.wi;
		ray.dir = bs.wi;
		ray.ori = surf.pos + ray.dir * 1e-4f;
		isLastVertexConnectible = isThisVertexConnectible;
	}
	if (sampleState == 2 && lastSampleState == 2) {
		stream.add(ps, luminance(ps.F()), sample1f(rng));
	}
	GRISResv resv = zeroGRIS();
	resv.copySample(stream.sample);
	if (stream.sumWeight > 0 && stream.weight > 0) {
		const float k = stream.sumWeight / stream.weight;
		resv.setF(resv.F() * k);
		resv.setRcLi(resv.rcLi() * k);
		resv.resampleWeight() = luminance(resv.F());
	}
	else {
		resv.q0.z = __uint_as_float(InvalidHitIndex);
		resv.setF(f3(0.0f));
	}
	resv.sampleCount() = 1;
	storeGRIS(f.grisThis + f.index(x, y), resv);
}

// gris_resample_temporal.comp -> temporalReuse
__global__ void __launch_bounds__(PassBlockX* PassBlockY) grisTemporalKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	if (!p.valid) return;
	const size_t idx = f.index(x, y);
	const float2 motion = f.motion[idx];
	const uint32_t rng = makeSeed(f.camera.seed, x, y) ^ 1u;
	uint32_t resvRng = ~rng;
	GRISResv resv = loadGRIS(f.grisThis + idx);

	if (st.temporalReuse) {
		if ((f.camera.frameIndex & 0x80000000u) == 0) {
			const Neighbor nb = lookupSurface(f, true, make_float2(p.uv.x + motion.x, p.uv.y + motion.y));
			if (nb.found && !(nb.matMeshId != p.matMeshId || dot(nb.norm, p.norm) < 0.95f || distance(p.pos, nb.pos) > 0.5f)) {
				const GRISResv prev = loadGRIS(f.grisPrev + nb.pixel);
				if (prev.valid()) grisReuseAndMerge(s, st, resv, primarySurface(p), p.uv, p.ray, prev, resvRng);
			}
		}
	}
	if (!resv.valid()) resv.reset();
	storeGRIS(f.grisTemp + idx, resv);
	// multi-GPU strips: boundary rows go straight into the neighbours' halo rows over NVLink peer memory
	if (f.peerGrisUp != nullptr && y < f.rowBegin + f.halo) storeGRIS(f.peerGrisUp + (size_t(y - f.peerUpStoreBegin) * f.width + x), resv);
	if (f.peerGrisDown != nullptr && y + f.halo >= f.rowEnd) storeGRIS(f.peerGrisDown + (size_t(y - f.peerDownStoreBegin) * f.width + x), resv);
}

// gris_resample_spatial.comp -> spatialReuse
__global__ void __launch_bounds__(PassBlockX* PassBlockY) grisSpatialKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	float3 radiance = f3(0.0f);
	if (p.valid) {
		const size_t idx = f.index(x, y);
		uint32_t rng = makeSeed(f.camera.seed, x, y) ^ 2u;
		GRISResv resv = loadGRIS(f.grisTemp + idx);
		const Surface dstPrimarySurf = primarySurface(p);
		const float texelX = 1.0f / float(f.width), texelY = 1.0f / float(f.height);

		if (st.spatialReuse) {
			for (uint32_t i = 0; i < 3; i++) {
				const float2 d = toConcentricDisk(sample2f(rng));
				const float2 nuv = make_float2(p.uv.x + d.x * 20.0f * texelX, p.uv.y + d.y * 20.0f * texelY);
				const Neighbor nb = lookupSurface(f, false, nuv);
				if (nb.found && !(dot(nb.norm, p.norm) < 0.9f || distance(p.pos, nb.pos) > 0.4f)) {
					const GRISResv nr = loadGRIS(f.grisTemp + nb.pixel);
					if (nr.valid()) grisReuseAndMerge(s, st, resv, dstPrimarySurf, p.uv, p.ray, nr, rng);
				}
			}
		}
		if (!resv.valid()) resv.reset();
		storeGRIS(f.grisThis + idx, resv);

		if (resv.valid() && resv.sampleCount() > 0 && resv.sampleValid()) {
			RcData rc;
			traceReplayPath(s, st, dstPrimarySurf, p.uv, p.ray, resv.flags(), resv.primaryRng(), rc);
			if (rc.prevInstance != InvalidHitIndex) {
				Surface rcPrevSurf, rcSurf;
				if (rc.prevInstance == SpecialHitIndex) rcPrevSurf = dstPrimarySurf;
				else loadSurfaceInfo(s, rc.prevInstance, rc.prevTriangle, rc.prevBary, rcPrevSurf);
				loadSurfaceInfo(s, resv.rcInstance(), __float_as_uint(resv.q0.w), make_float2(resv.q0.x, resv.q0.y), rcSurf);
				const Mat rcPrevMat = loadMaterial(s, rcPrevSurf.matIndex);
				const float3 wi = normalize(rcSurf.pos - rcPrevSurf.pos);
				const float3 Li = reconnectionLi(s, resv, rc, rcPrevSurf, rcSurf, rcPrevMat, wi, resv.rcPrevSamplePdf());
				if (!isBlack(Li) && !hasNan(Li)) radiance = Li / luminance(Li) * resv.resampleWeight() / resv.sampleCount();
			}
		}
		radiance = clampColor(radiance);
	}
	accumulate(f.indirectOutput, f, x, y, radiance);
}

void launchGRISPathTrace(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st) {
	grisPathTraceKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s, p);
}
void launchGRISTemporal(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st) {
	grisTemporalKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s, p);
}
void launchGRISSpatial(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st) {
	grisSpatialKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s, p);
}

} // namespace rt
