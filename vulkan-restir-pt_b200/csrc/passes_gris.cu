// ReSTIR PT / GRIS (BASELINE.json configs 3 and 4 — the north-star path): candidate path generation with
// reconnection-vertex selection, hybrid-shift temporal reuse, hybrid-shift spatial reuse + final shading.
//   reference src/shader/gris_path_trace.glsl:45-305, gris_retrace.glsl:42-236, gris_reservoir.glsl:37-136,
//   gris_resample_temporal.glsl:11-83, gris_resample_spatial.glsl:11-134 (+ the three .comp entry points)
//   host sequence: GRISReSTIR::render (src/GRISReSTIR.cpp:9-53)
#include <cstdlib>
#include "passes.h"
#include "shading.cuh"
#include "persist.cuh"

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

// ---- gris_path_trace.comp -> tracePath, as a wavefront ---------------------------------------------------------
//
// The shader runs one invocation per pixel through a loop of up to 15 bounces with two ray queries per bounce
// (gris_path_trace.glsl:85-257).  Here the loop is cut at its ray queries: per bounce
//     [extend]   trace_queue.cu, closest hit of every queued extension ray
//     [vertex]   grisVertexKernel: surface fetch, state machine, emitter hit, reconnection-vertex choice, random
//                draws, light sample -> shadow-ray queue                                     (:90-175)
//     [shadow]   trace_queue.cu, any hit of every queued shadow ray
//     [scatter]  grisScatterKernel: the three NEE candidate kinds, Russian roulette, BSDF sample, throughput
//                update -> extension-ray queue of the next bounce                             (:176-257)
// and the state of a path lives in global memory between the kernels (PathState below, 160 B per pixel).  A path
// that ends (miss, emitter, roulette, failed BSDF sample, bounce limit) writes its reservoir at once (:259-278).
// The winner of the path's streaming RIS (StreamSampler, :10-33) is kept directly in the pixel's output reservoir
// slot: it is only ever overwritten until the path ends.
// Per-pixel arithmetic, its order and the RNG stream are exactly those of the shader's loop.

struct PathState {
	float3 dir;              // direction that arrived at the current vertex (wo = -dir)
	uint32_t rng;
	float3 throughput, rcThroughput, lastPos;
	float bsPdf;
	uint32_t bsType;
	uint32_t sampleState, lastSampleState;
	bool isLastVertexConnectible, isThisVertexConnectible, streamWritten;
	int bounce;
	float streamWeight, streamSumWeight;
	GRISResv ps;             // q0..q4 = the GRISPathSample under construction
};

RT_DEV void storePathState(float4* __restrict__ base, size_t pix, const PathState& st) {
	float4* w = base + pix * PathStateWords;
	const uint32_t flags = uint32_t(st.bounce) | (st.sampleState << 4) | (st.lastSampleState << 6) | (st.isLastVertexConnectible ? 1u << 8 : 0u)
		| (st.isThisVertexConnectible ? 1u << 9 : 0u) | (st.streamWritten ? 1u << 10 : 0u) | (st.bsType << 16);
	w[0] = make_float4(st.dir.x, st.dir.y, st.dir.z, __uint_as_float(st.rng));
	w[1] = make_float4(st.throughput.x, st.throughput.y, st.throughput.z, st.bsPdf);
	w[2] = make_float4(st.rcThroughput.x, st.rcThroughput.y, st.rcThroughput.z, st.streamWeight);
	w[3] = make_float4(st.lastPos.x, st.lastPos.y, st.lastPos.z, st.streamSumWeight);
	w[4] = make_float4(__uint_as_float(flags), 0.f, 0.f, 0.f);
	w[5] = st.ps.q0; w[6] = st.ps.q1; w[7] = st.ps.q2; w[8] = st.ps.q3; w[9] = st.ps.q4;
}
RT_DEV void loadPathState(const float4* __restrict__ base, size_t pix, PathState& st) {
	const float4* w = base + pix * PathStateWords;
	const float4 a = w[0], b = w[1], c = w[2], d = w[3], e = w[4];
	st.dir = f3(a); st.rng = __float_as_uint(a.w);
	st.throughput = f3(b); st.bsPdf = b.w;
	st.rcThroughput = f3(c); st.streamWeight = c.w;
	st.lastPos = f3(d); st.streamSumWeight = d.w;
	const uint32_t flags = __float_as_uint(e.x);
	st.bounce = int(flags & 15u);
	st.sampleState = (flags >> 4) & 3u; st.lastSampleState = (flags >> 6) & 3u;
	st.isLastVertexConnectible = (flags >> 8) & 1u; st.isThisVertexConnectible = (flags >> 9) & 1u; st.streamWritten = (flags >> 10) & 1u;
	st.bsType = flags >> 16;
	st.ps.q0 = w[5]; st.ps.q1 = w[6]; st.ps.q2 = w[7]; st.ps.q3 = w[8]; st.ps.q4 = w[9];
	st.ps.q5 = make_float4(0.f, 0.f, 0.f, 0.f);
}

// StreamSampler::add (gris_path_trace.glsl:21-32); the selected sample goes straight to the pixel's reservoir slot
RT_DEV void streamAdd(PathState& st, RptGRISReservoir* __restrict__ slot, float w, float r) {
	st.streamSumWeight += w;
	if (r * st.streamSumWeight < w) {
		st.streamWeight = w;
		st.streamWritten = true;
		float4* q = reinterpret_cast<float4*>(slot);
		q[0] = st.ps.q0; q[1] = st.ps.q1; q[2] = st.ps.q2; q[3] = st.ps.q3; q[4] = st.ps.q4;
	}
}

// end of tracePath (gris_path_trace.glsl:259-278)
RT_DEV void finishPath(PathState& st, RptGRISReservoir* __restrict__ slot) {
	if (st.sampleState == 2 && st.lastSampleState == 2) {
		streamAdd(st, slot, luminance(st.ps.F()), sample1f(st.rng));
	}
	float4* q = reinterpret_cast<float4*>(slot);
	const bool scaled = st.streamSumWeight > 0 && st.streamWeight > 0;
	if (st.streamWritten) {
		float4 q0 = q[0], q1 = q[1], q4 = q[4];
		float resampleWeight = 0.0f;
		if (scaled) {
			const float k = st.streamSumWeight / st.streamWeight;
			const float3 F = f3(q4) * k, rcLi = f3(q1) * k;
			q4.x = F.x; q4.y = F.y; q4.z = F.z;
			q1.x = rcLi.x; q1.y = rcLi.y; q1.z = rcLi.z;
			resampleWeight = luminance(F);
			q[1] = q1;
		}
		else {
			q0.z = __uint_as_float(InvalidHitIndex);
			q4.x = 0.f; q4.y = 0.f; q4.z = 0.f;
			q[0] = q0;
		}
		q[4] = q4;
		q[5] = make_float4(1.0f, resampleWeight, 0.f, 0.f);
	}
	else {   // nothing was ever selected: the zero sample, marked invalid
		const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
		q[0] = make_float4(0.f, 0.f, __uint_as_float(InvalidHitIndex), 0.f);
		q[1] = z; q[2] = z; q[3] = z; q[4] = z;
		q[5] = make_float4(1.0f, 0.f, 0.f, 0.f);
	}
}

struct VertexOut {
	float4 lightRandSample;
	float resvRandSample;
};

// [vertex] gris_path_trace.glsl:101-174.  isecWord = the hit that led here.  Returns false when the path ended here.
RT_DEV bool vertexStage(PathState& st, const RptGRISSettings& set, const Surface& surf, const Mat& mat, float4 isecWord, float sumPower,
                        RptGRISReservoir* __restrict__ slot, VertexOut& out) {
	GRISResv& ps = st.ps;
	const int bounce = st.bounce;
	ps.setFlags(withPathLength(ps.flags(), uint32_t(bounce + 1)));
	const float cosPrevWi = dot(st.dir, surf.norm);
	const float distToPrev = distance(st.lastPos, surf.pos);
	const float geometryJacobian = abs_(cosPrevWi) / square(distToPrev);
	st.isThisVertexConnectible = surf.isLight || isBSDFConnectible(mat);
	st.lastSampleState = st.sampleState;
	st.sampleState = nextRcVertexSampleState(st.sampleState, st.isThisVertexConnectible);
	if (set.shiftType == ShiftReconnection && bounce == 1 && !surf.isLight) {
		st.sampleState = 2;
		st.lastSampleState = 1;
	}
	out.resvRandSample = sample1f(st.rng);

	if (surf.isLight) {
		if (bounce > 1 && cosPrevWi < 0) {
			float weight = 1.0f;
			const float lightPdf = luminance(surf.albedo) / sumPower / geometryJacobian;
			if (!isSampleTypeDelta(st.bsType)) weight = MISWeight(st.bsPdf, lightPdf);
			const float3 weightedLi = surf.albedo * weight;
			if (st.sampleState == 2 && st.lastSampleState == 2) {
				ps.setRcLi(ps.rcLi() + weightedLi * st.rcThroughput);
				ps.setF(ps.F() + weightedLi * st.throughput);
			}
			else if ((st.sampleState == 2 && st.lastSampleState == 1) && st.isLastVertexConnectible && distToPrev > GRISDistanceThreshold) {
				ps.q0 = isecWord;
				ps.q1.w = __uint_as_float(st.rng);
				ps.rcPrevSamplePdf() = st.bsPdf;
				ps.rcJacobian() = geometryJacobian;
				ps.setRcLi(weightedLi);
				ps.setRcWi(f3(0.0f));
				ps.setF(weightedLi * st.throughput);
				ps.setFlags(withRcVertexType(withRcVertexId(ps.flags(), uint32_t(bounce)), RcLightScattered));
				streamAdd(st, slot, luminance(ps.F()), out.resvRandSample);
			}
		}
		return false;
	}
	const bool connectible = st.isThisVertexConnectible && st.isLastVertexConnectible && distToPrev > GRISDistanceThreshold;
	if ((st.sampleState == 2 && st.lastSampleState == 1) && (connectible || set.shiftType == ShiftReconnection)) {
		ps.q0 = isecWord;
		ps.q1.w = __uint_as_float(st.rng);
		ps.rcPrevSamplePdf() = st.bsPdf;
		ps.rcJacobian() = geometryJacobian;
		ps.setFlags(withRcVertexType(withRcVertexId(ps.flags(), uint32_t(bounce)), RcSurface));
		st.rcThroughput = f3(1.0f);
	}
	out.lightRandSample = sample4f(st.rng);
	out.resvRandSample = sample1f(st.rng);
	return true;
}

// NEE contributions (gris_path_trace.glsl:186-223) of an unoccluded light sample
RT_DEV void neeStage(PathState& st, const Surface& surf, const Mat& mat, const LightSample& ls, float resvRandSample, RptGRISReservoir* __restrict__ slot) {
	GRISResv& ps = st.ps;
	const float3 wo = -st.dir;
	const float bsdfPdf = absDot(surf.norm, ls.wi) * RT_PI_INV;
	const float weight = MISWeight(ls.pdf, bsdfPdf);
	const float3 scatterTerm = evalBSDF(mat, surf.albedo, surf.norm, wo, ls.wi) * satDot(surf.norm, ls.wi);
	const float3 weightedLi = ls.radiance / ls.pdf * weight;
	if (st.sampleState == 2 && st.lastSampleState == 2) {
		ps.setRcLi(ps.rcLi() + weightedLi * scatterTerm * st.rcThroughput);
		ps.setF(ps.F() + weightedLi * scatterTerm * st.throughput);
	}
	else if (st.sampleState == 2 && st.lastSampleState == 1) {
		ps.setRcLi(weightedLi);
		ps.setRcWi(ls.wi);
		ps.setF(weightedLi * scatterTerm * st.throughput);
		streamAdd(st, slot, luminance(ps.F()), resvRandSample);
	}
	else if (st.sampleState == 1 && st.isThisVertexConnectible && ls.dist > GRISDistanceThreshold) {
		ps.q0 = make_float4(ls.bary.x, ls.bary.y, __uint_as_float(0u), __uint_as_float(ls.id));
		ps.q1.w = __uint_as_float(st.rng);
		ps.rcPrevSamplePdf() = ls.pdf;
		ps.rcJacobian() = ls.jacobian;
		ps.setRcLi(ls.radiance * weight);
		ps.setRcWi(f3(0.0f));
		ps.setF(weightedLi * scatterTerm * st.throughput);
		ps.setFlags(withRcVertexType(withRcVertexId(ps.flags(), uint32_t(st.bounce + 1)), RcLightSampled));
		streamAdd(st, slot, luminance(ps.F()), resvRandSample);
	}
}

// [scatter] gris_path_trace.glsl:226-257.  Returns false when the path ended; else rayOri/st.dir hold the next ray.
RT_DEV bool scatterStage(PathState& st, const RptGRISSettings& set, const Surface& surf, const Mat& mat, float3& rayOri) {
	GRISResv& ps = st.ps;
	const float3 wo = -st.dir;
	if (st.bounce > 4) {
		const float pdfTerminate = max_(1.0f - luminance(st.throughput) * set.rrScale, 0.0f);
		if (sample1f(st.rng) < pdfTerminate) return false;
		st.throughput /= (1.0f - pdfTerminate);
		st.rcThroughput /= (1.0f - pdfTerminate);
	}
	const float3 r3 = sample3f(st.rng);
	BSDFSample bs = emptyBSDFSample();
	bs.pdf = st.bsPdf; bs.type = st.bsType;
	const bool ok = sampleBSDF(mat, surf.albedo, surf.norm, wo, r3, bs);
	if (!ok || bs.pdf < 1e-6f) return false;
	st.bsPdf = bs.pdf; st.bsType = bs.type;
	const float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
	const float3 scatterTerm = bs.bsdf * cosTheta / bs.pdf;
	st.throughput *= scatterTerm;
	if (st.sampleState == 2 && st.lastSampleState == 2) {
		st.rcThroughput *= scatterTerm;
	}
	else if (st.sampleState == 2 && st.lastSampleState == 1) {
		ps.setRcLi(f3(0.0f));
		ps.setRcWi(bs.wi);
		ps.setF(f3(0.0f));
		st.rcThroughput /= bs.pdf;
	}
	st.lastPos = surf.pos;
	st.dir = bs.wi;
	rayOri = surf.pos + st.dir * 1e-4f;
	st.isLastVertexConnectible = st.isThisVertexConnectible;
	st.bounce++;
	return st.bounce < 15;
}

constexpr uint32_t SlotEnded = 0xfffffffeu, SlotNoShadow = 0xffffffffu;
constexpr int ShadeBlock = 128;

// appends one entry per requesting lane to a device queue: one atomic per warp
RT_DEV uint32_t queueAppend(uint32_t* __restrict__ count, bool want) {
	const unsigned mask = __ballot_sync(__activemask(), want);
	if (!want) return 0u;
	const uint32_t lane = threadIdx.x & 31u;
	const int leader = __ffs(int(mask)) - 1;
	uint32_t base = 0;
	if (int(lane) == leader) base = atomicAdd(count, uint32_t(__popc(mask)));
	base = __shfl_sync(mask, base, leader);
	return base + uint32_t(__popc(mask & ((1u << lane) - 1u)));
}

RT_DEV void pushExtensionRay(const FrameView& f, int nextBounce, uint32_t pix, float3 ori, float3 dir) {
	uint32_t* cnt = f.wf.counters + 4 * nextBounce;
	const uint32_t slot = queueAppend(cnt, true);
	float4* rq = f.wf.rays[nextBounce & 1] + 2 * size_t(slot);
	rq[0] = make_float4(ori.x, ori.y, ori.z, MinRayDistance);
	rq[1] = make_float4(dir.x, dir.y, dir.z, MaxRayDistance);
	f.wf.pix[nextBounce & 1][slot] = pix;
}

// bounce 0: the primary hit comes from the G-buffer, there is no NEE (gris_path_trace.glsl:176), so vertex and scatter
// stages run back to back.  One thread per pixel in 8x4-tile order.
__global__ void __launch_bounds__(ShadeBlock) grisBeginKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings set) {
	const uint32_t tilesX = (f.width + 7u) / 8u;
	const uint32_t id = blockIdx.x * ShadeBlock + threadIdx.x;
	const uint32_t tile = id >> 5, within = id & 31u;
	const uint32_t x = (tile % tilesX) * 8u + (within & 7u), y = f.rowBegin + (tile / tilesX) * 4u + (within >> 3);
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	if (!p.valid) return;   // background pixels keep their old reservoir (gris_path_trace.glsl:54-56)
	const uint32_t pix = uint32_t(f.index(x, y));
	RptGRISReservoir* slot = f.grisThis + pix;

	PathState st;
	st.dir = p.ray.dir;
	st.rng = makeSeed(f.camera.seed, x, y);
	st.throughput = f3(1.0f); st.rcThroughput = f3(0.0f); st.lastPos = f3(0.0f);
	st.bsPdf = 0.0f; st.bsType = 0;
	st.sampleState = 0; st.lastSampleState = 0;
	st.isLastVertexConnectible = false; st.isThisVertexConnectible = false; st.streamWritten = false;
	st.bounce = 0;
	st.streamWeight = 0.0f; st.streamSumWeight = 0.0f;
	st.ps = zeroGRIS();   // GRISPathSampleReset, gris_reservoir.glsl:61-69
	st.ps.q0.z = __uint_as_float(InvalidHitIndex);
	st.ps.q4.w = __uint_as_float(st.rng);   // primaryRng

	const Surface surf = primarySurface(p);
	const Mat mat = loadMaterial(s, uint32_t(p.matId));
	VertexOut vo;
	float3 rayOri;
	const float4 isecWord = make_float4(0.f, 0.f, __uint_as_float(0u), __uint_as_float(0u));
	bool go = vertexStage(st, set, surf, mat, isecWord, s.lightTable[0].prob, slot, vo);
	if (go) go = scatterStage(st, set, surf, mat, rayOri);
	if (!go) { finishPath(st, slot); return; }
	storePathState(f.wf.state, pix, st);
	pushExtensionRay(f, 1, pix, rayOri, st.dir);
}

// [vertex] of bounce >= 1: one thread per slot of the bounce's extension queue
__global__ void __launch_bounds__(ShadeBlock) grisVertexKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings set, int bounce) {
	const uint32_t n = f.wf.counters[4 * bounce];
	const float sumPower = s.lightTable[0].prob;
	for (uint32_t slotIdx = blockIdx.x * ShadeBlock + threadIdx.x; slotIdx < n; slotIdx += gridDim.x * ShadeBlock) {
		const uint32_t pix = f.wf.pix[bounce & 1][slotIdx];
		const RptIntersection hit = f.wf.hits[slotIdx];
		RptGRISReservoir* slot = f.grisThis + pix;
		float4* vtx = f.wf.vertex + size_t(slotIdx) * VertexWords;
		PathState st;
		loadPathState(f.wf.state, pix, st);
		if (hit.instanceIdx == InvalidHitIndex) {   // the ray left the scene (gris_path_trace.glsl:92-94)
			finishPath(st, slot);
			vtx[2].w = __uint_as_float(SlotEnded);
			continue;
		}
		Surface surf;
		loadSurfaceInfo(s, hit, surf);
		const Mat mat = loadMaterial(s, surf.matIndex);
		VertexOut vo;
		vo.lightRandSample = make_float4(0.f, 0.f, 0.f, 0.f); vo.resvRandSample = 0.f;
		const float4 isecWord = make_float4(hit.bary[0], hit.bary[1], __uint_as_float(hit.instanceIdx), __uint_as_float(hit.triangleIdx));
		if (!vertexStage(st, set, surf, mat, isecWord, sumPower, slot, vo)) {
			finishPath(st, slot);
			vtx[2].w = __uint_as_float(SlotEnded);
			continue;
		}
		uint32_t shadowIdx = SlotNoShadow;
		const bool nee = !isBSDFDelta(mat);
		if (nee) {
			const LightSample ls = sampleLight(s, surf.pos, vo.lightRandSample);
			shadowIdx = queueAppend(f.wf.counters + 4 * bounce + 1, true);
			float4* rq = f.wf.shadowRays + 2 * size_t(shadowIdx);
			rq[0] = make_float4(surf.pos.x, surf.pos.y, surf.pos.z, MinRayDistance);
			rq[1] = make_float4(ls.wi.x, ls.wi.y, ls.wi.z, ls.dist - MinRayDistance);
		}
		storePathState(f.wf.state, pix, st);
		vtx[0] = make_float4(surf.pos.x, surf.pos.y, surf.pos.z, __uint_as_float(surf.matIndex));
		vtx[1] = make_float4(surf.norm.x, surf.norm.y, surf.norm.z, vo.resvRandSample);
		vtx[2] = make_float4(surf.albedo.x, surf.albedo.y, surf.albedo.z, __uint_as_float(shadowIdx));
		vtx[3] = vo.lightRandSample;
	}
}

// [scatter] of bounce >= 1: one thread per slot of the same queue
__global__ void __launch_bounds__(ShadeBlock) grisScatterKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings set, int bounce) {
	const uint32_t n = f.wf.counters[4 * bounce];
	for (uint32_t slotIdx = blockIdx.x * ShadeBlock + threadIdx.x; slotIdx < n; slotIdx += gridDim.x * ShadeBlock) {
		const float4* vtx = f.wf.vertex + size_t(slotIdx) * VertexWords;
		const float4 v2 = vtx[2];
		const uint32_t shadowIdx = __float_as_uint(v2.w);
		if (shadowIdx == SlotEnded) continue;
		const float4 v0 = vtx[0], v1 = vtx[1];
		const uint32_t pix = f.wf.pix[bounce & 1][slotIdx];
		RptGRISReservoir* slot = f.grisThis + pix;
		PathState st;
		loadPathState(f.wf.state, pix, st);
		Surface surf;
		surf.pos = f3(v0); surf.norm = f3(v1); surf.albedo = f3(v2); surf.matIndex = __float_as_uint(v0.w); surf.isLight = false;
		const Mat mat = loadMaterial(s, surf.matIndex);
		if (shadowIdx != SlotNoShadow && f.wf.occluded[shadowIdx] == 0) {
			const LightSample ls = sampleLight(s, surf.pos, vtx[3]);
			if (ls.pdf > 1e-6f) neeStage(st, surf, mat, ls, v1.w, slot);
		}
		float3 rayOri;
		if (!scatterStage(st, set, surf, mat, rayOri)) { finishPath(st, slot); continue; }
		storePathState(f.wf.state, pix, st);
		pushExtensionRay(f, bounce + 1, pix, rayOri, st.dir);
	}
}

// experiment: vertex + in-line shadow ray + scatter in one kernel (one state round trip per bounce)
__global__ void __launch_bounds__(ShadeBlock) grisShadeFusedKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings set, int bounce) {
	const uint32_t n = f.wf.counters[4 * bounce];
	const float sumPower = s.lightTable[0].prob;
	for (uint32_t slotIdx = blockIdx.x * ShadeBlock + threadIdx.x; slotIdx < n; slotIdx += gridDim.x * ShadeBlock) {
		const uint32_t pix = f.wf.pix[bounce & 1][slotIdx];
		const RptIntersection hit = f.wf.hits[slotIdx];
		RptGRISReservoir* slot = f.grisThis + pix;
		PathState st;
		loadPathState(f.wf.state, pix, st);
		if (hit.instanceIdx == InvalidHitIndex) { finishPath(st, slot); continue; }
		Surface surf;
		loadSurfaceInfo(s, hit, surf);
		const Mat mat = loadMaterial(s, surf.matIndex);
		VertexOut vo;
		vo.lightRandSample = make_float4(0.f, 0.f, 0.f, 0.f); vo.resvRandSample = 0.f;
		const float4 isecWord = make_float4(hit.bary[0], hit.bary[1], __uint_as_float(hit.instanceIdx), __uint_as_float(hit.triangleIdx));
		if (!vertexStage(st, set, surf, mat, isecWord, sumPower, slot, vo)) { finishPath(st, slot); continue; }
		if (!isBSDFDelta(mat)) {
			const LightSample ls = sampleLight(s, surf.pos, vo.lightRandSample);
			const bool shadowed = traceShadow(s, surf.pos, MinRayDistance, ls.wi, ls.dist - MinRayDistance);
			if (!shadowed && ls.pdf > 1e-6f) neeStage(st, surf, mat, ls, vo.resvRandSample, slot);
		}
		float3 rayOri;
		if (!scatterStage(st, set, surf, mat, rayOri)) { finishPath(st, slot); continue; }
		storePathState(f.wf.state, pix, st);
		pushExtensionRay(f, bounce + 1, pix, rayOri, st.dir);
	}
}

// gris_resample_temporal.comp -> temporalReuse
__global__ void __launch_bounds__(PassBlockX* PassBlockY, 8) grisTemporalKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
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
	static const int vertexBlocks = persistentBlocks(reinterpret_cast<const void*>(grisVertexKernel), ShadeBlock);
	static const int scatterBlocks = persistentBlocks(reinterpret_cast<const void*>(grisScatterKernel), ShadeBlock);
	const uint32_t rows = f.rowEnd - f.rowBegin;
	const uint32_t slots = ((f.width + 7u) / 8u) * ((rows + 3u) / 4u) * 32u;
	cudaMemsetAsync(f.wf.counters, 0, size_t(WavefrontMaxBounces) * 4 * sizeof(uint32_t), st);
	grisBeginKernel<<<(slots + ShadeBlock - 1) / ShadeBlock, ShadeBlock, 0, st>>>(f, s, p);
	static const int fusedMode = getenv("RPT_FUSED_SHADE") ? atoi(getenv("RPT_FUSED_SHADE")) : 0;
	static const int fusedBlocks = persistentBlocks(reinterpret_cast<const void*>(grisShadeFusedKernel), ShadeBlock);
	for (int bounce = 1; bounce < 15; bounce++) {
		uint32_t* c = f.wf.counters + 4 * bounce;
		launchTraceQueueClosest(s, f.wf.rays[bounce & 1], c + 0, 0, c + 2, f.wf.hits, st);
		if (fusedMode) { grisShadeFusedKernel<<<fusedBlocks, ShadeBlock, 0, st>>>(f, s, p, bounce); continue; }
		grisVertexKernel<<<vertexBlocks, ShadeBlock, 0, st>>>(f, s, p, bounce);
		launchTraceQueueAny(s, f.wf.shadowRays, c + 1, 0, c + 3, f.wf.occluded, st);
		grisScatterKernel<<<scatterBlocks, ShadeBlock, 0, st>>>(f, s, p, bounce);
	}
}
void launchGRISTemporal(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st) {
	grisTemporalKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s, p);
}
void launchGRISSpatial(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st) {
	grisSpatialKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s, p);
}

} // namespace rt
