// ReSTIR PT / GRIS (BASELINE.json configs 3 and 4 — the north-star path): candidate path generation with
// reconnection-vertex selection, hybrid-shift temporal reuse, hybrid-shift spatial reuse + final shading.
//   reference src/shader/gris_path_trace.glsl:45-305, gris_retrace.glsl:42-236, gris_reservoir.glsl:37-136,
//   gris_resample_temporal.glsl:11-83, gris_resample_spatial.glsl:11-134 (+ the three .comp entry points)
//   host sequence: GRISReSTIR::render (src/GRISReSTIR.cpp:9-53)
#include <algorithm>
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

RT_DEV bool grisMerge(GRISResv& resv, GRISResv& rhs, float r) {   // gris_reservoir.glsl:114-123; true: rhs's sample was taken
	resv.sampleCount() += rhs.sampleCount();
	resv.resampleWeight() += rhs.resampleWeight();
	const bool take = r * resv.resampleWeight() < rhs.resampleWeight();
	if (take) resv.copySample(rhs);
	return take;
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
// random numbers, consuming them in lock-step with tracePath, up to the vertex before the reconnection vertex.
// replayVertex is one iteration of that loop from the point where the vertex's surface is known (:62-135); it is shared by the
// in-line form (traceReplayPath: one thread walks the whole prefix) and the wavefront form (rwStepKernel: one kernel per bounce
// around the queue traversal kernel).  Returns ReplayContinue (ray / wo / throughput / rng advanced to the next vertex),
// ReplayFound (rc filled: this is the vertex before the reconnection vertex) or ReplayFailed (rc untouched).
constexpr int ReplayContinue = 0, ReplayFound = 1, ReplayFailed = 2;
RT_DEV int replayVertex(const RptGRISSettings& st, int bounce, uint32_t targetId, const Surface& surf, const Mat& mat, uint32_t curInst, uint32_t curTri,
                        float2 curBary, float3& throughput, float3& wo, uint32_t& rng, Ray& ray, RcData& rc) {
	const bool isThisVertexConnectible = isBSDFConnectible(mat);
	sample1f(rng);
	if (surf.isLight) return ReplayFailed;
	if (uint32_t(bounce) == targetId - 1u) {
		if (!isThisVertexConnectible) return ReplayFailed;
		rc.prevInstance = curInst; rc.prevTriangle = curTri; rc.prevBary = curBary;
		rc.rcPrevWo = wo; rc.rcPrevThroughput = throughput;
		return ReplayFound;
	}
	sample4f(rng);
	sample1f(rng);
	if (bounce > 4) {
		const float pdfTerminate = max_(1.0f - luminance(throughput) * st.rrScale, 0.0f);
		if (sample1f(rng) < pdfTerminate) return ReplayFailed;
		throughput /= (1.0f - pdfTerminate);
	}
	const float3 r3 = sample3f(rng);
	BSDFSample bs = emptyBSDFSample();
	if (!sampleBSDF(mat, surf.albedo, surf.norm, wo, r3, bs) || bs.pdf < 1e-6f) return ReplayFailed;
	const float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
	throughput *= bs.bsdf * cosTheta / bs.pdf;
	wo = -bs.wi;
	ray.dir = bs.wi;
	ray.ori = surf.pos + ray.dir * 1e-4f;
	return ReplayContinue;
}
RT_DEV void resetRcData(RcData& rc) {
	rc.prevInstance = InvalidHitIndex; rc.prevTriangle = 0; rc.prevBary = make_float2(0.f, 0.f);
	rc.rcPrevWo = f3(0.0f); rc.rcPrevThroughput = f3(0.0f);
}
// CanTrace = false: only the rcVertexId == 1 case (no ray needed) is compiled in
template <bool CanTrace = true>
RT_DEV void traceReplayPath(const SceneView& s, const RptGRISSettings& st, const Surface& primarySurf, float2 primaryUv, Ray ray,
                            uint32_t targetFlags, uint32_t rng, RcData& rc) {
	float3 throughput = f3(1.0f);
	float3 wo = -ray.dir;
	Surface surf = primarySurf;
	Mat mat = loadMaterial(s, surf.matIndex);
	resetRcData(rc);
	uint32_t curInst = SpecialHitIndex, curTri = 0;
	float2 curBary = primaryUv;
	const uint32_t targetId = flagsRcVertexId(targetFlags);
	if (targetId == 1) {
		rc.prevInstance = curInst; rc.prevTriangle = curTri; rc.prevBary = curBary;
		rc.rcPrevWo = wo; rc.rcPrevThroughput = throughput;
		return;
	}
	if (!CanTrace) return;
	for (int bounce = 0; bounce < 15; bounce++) {
		if (bounce > 0) {
			const Hit h = traceClosestHit(s, ray.ori, MinRayDistance, ray.dir, MaxRayDistance);
			if (h.instanceIdx == InvalidHitIndex) break;
			curInst = h.instanceIdx; curTri = h.triangleIdx; curBary = make_float2(h.u, h.v);
			loadSurfaceInfo(s, h, surf);
			mat = loadMaterial(s, surf.matIndex);
		}
		if (replayVertex(st, bounce, targetId, surf, mat, curInst, curTri, curBary, throughput, wo, rng, ray, rc) != ReplayContinue) break;
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

// ---- the hybrid shift (GRISReservoirReuseAndMerge, gris_retrace.glsl:138-236), cut at its visibility ray -------------
// shiftPrepare: replay + reconnection geometry + the validity tests up to the ray (:150-184)
// shiftFinish:  shifted contribution, Jacobian, new target function, reweighting (:186-231)
// The per-pixel passes call both around an in-line visibility ray (grisReuseAndMerge); the wavefront reuse passes
// run them in two kernels with the ray traced from a queue in between.
constexpr uint32_t TaskInvalid = 0, TaskRay = 1, TaskSkip = 3;

struct ShiftTask {
	uint32_t status;                 // TaskInvalid: the source sample cannot be shifted here; TaskRay: visibility decides
	Surface rcPrevSurf, rcSurf;      // the two ends of the reconnection segment (status == TaskRay)
	float3 rcPrevWo, rcPrevThroughput;
};

// the part of shiftPrepare after the replay (:156-184): reconnection geometry and validity tests from the replay's result
RT_DEV void shiftPrepareFromRc(const SceneView& s, const Surface& dstPrimarySurf, const GRISResv& src, const RcData& rc, ShiftTask& t) {
	t.status = TaskInvalid;
	if (rc.prevInstance == InvalidHitIndex) return;
	if (rc.prevInstance == SpecialHitIndex) t.rcPrevSurf = dstPrimarySurf;
	else loadSurfaceInfo(s, rc.prevInstance, rc.prevTriangle, rc.prevBary, t.rcPrevSurf);
	loadSurfaceInfo(s, src.rcInstance(), __float_as_uint(src.q0.w), make_float2(src.q0.x, src.q0.y), t.rcSurf);
	t.rcPrevWo = rc.rcPrevWo; t.rcPrevThroughput = rc.rcPrevThroughput;
	const Mat rcPrevMat = loadMaterial(s, t.rcPrevSurf.matIndex);
	const float dist = distance(t.rcPrevSurf.pos, t.rcSurf.pos);
	const float3 wi = normalize(t.rcSurf.pos - t.rcPrevSurf.pos);
	const float cosTheta = -dot(t.rcSurf.norm, wi);
	const float dstJacobian = abs_(cosTheta) / square(dist);
	const float jacobian = dstJacobian / src.q3.w;
	if (dist > GRISDistanceThreshold && cosTheta > 0 && !isnan_(jacobian) && src.q3.w > 0 && isBSDFConnectible(rcPrevMat)) t.status = TaskRay;
}
template <bool CanTrace = true>
RT_DEV void shiftPrepare(const SceneView& s, const RptGRISSettings& st, const Surface& dstPrimarySurf, float2 dstUv, const Ray& primaryRay,
                         const GRISResv& src, ShiftTask& t) {
	t.status = TaskInvalid;
	if (!src.sampleValid()) return;
	RcData rc;
	traceReplayPath<CanTrace>(s, st, dstPrimarySurf, dstUv, primaryRay, src.flags(), src.primaryRng(), rc);
	shiftPrepareFromRc(s, dstPrimarySurf, src, rc, t);
}

// the visibility ray of traceVisibility(rcPrevSurf.pos, rcSurf.pos), ray_query.glsl:27-38
RT_DEV void shiftVisibilityRay(const ShiftTask& t, float3& dir, float& tmax) {
	dir = normalize(t.rcSurf.pos - t.rcPrevSurf.pos);
	tmax = distance(t.rcSurf.pos, t.rcPrevSurf.pos) - MinRayDistance;
}

RT_DEV void shiftFinish(const SceneView& s, GRISResv& src, const ShiftTask& t, bool srcSampleValid) {
	if (srcSampleValid) {
		const Mat rcPrevMat = loadMaterial(s, t.rcPrevSurf.matIndex);
		const float dist = distance(t.rcPrevSurf.pos, t.rcSurf.pos);
		const float3 wi = normalize(t.rcSurf.pos - t.rcPrevSurf.pos);
		const float cosTheta = -dot(t.rcSurf.norm, wi);
		const float dstJacobian = abs_(cosTheta) / square(dist);
		const float jacobian = dstJacobian / src.rcJacobian();
		RcData rc;
		rc.rcPrevWo = t.rcPrevWo; rc.rcPrevThroughput = t.rcPrevThroughput;
		float3 Li = f3(0.0f);
		float dstPHat = 0, dstSamplePdf = 0;
		const uint32_t rcType = flagsRcVertexType(src.flags());
		if (!isnan_(src.rcPrevSamplePdf()) && src.rcPrevSamplePdf() > 1e-6f) {
			Li = reconnectionLi(s, src, rc, t.rcPrevSurf, t.rcSurf, rcPrevMat, wi, src.rcPrevSamplePdf());
			if (!isBlack(Li) && !hasNan(Li)) dstPHat = luminance(Li * jacobian);
			if (rcType == RcLightSampled) {
				const float sumPower = s.lightTable[0].prob;
				dstSamplePdf = luminance(t.rcSurf.albedo) / sumPower / dstJacobian;
			}
			else {
				dstSamplePdf = evalPdf(rcPrevMat, t.rcPrevSurf.norm, rc.rcPrevWo, wi);
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
}

RT_DEV void grisReuseAndMerge(const SceneView& s, const RptGRISSettings& st, GRISResv& dst, const Surface& dstPrimarySurf, float2 dstUv,
                              const Ray& primaryRay, GRISResv src, uint32_t& rng) {
	ShiftTask t;
	shiftPrepare(s, st, dstPrimarySurf, dstUv, primaryRay, src, t);
	const bool srcSampleValid = t.status == TaskRay && traceVisibility(s, t.rcPrevSurf.pos, t.rcSurf.pos);
	shiftFinish(s, src, t, srcSampleValid);
	if (src.valid()) grisMerge(dst, src, sample1f(rng));
	grisCap(dst, float(st.cap));
}

} // namespace

// ---- gris_path_trace.comp -> tracePath, as a wavefront ---------------------------------------------------------
//
// The shader runs one invocation per pixel through a loop of up to 15 bounces with two ray queries per bounce
// (gris_path_trace.glsl:85-257).  Here the loop is cut at its ray queries and every bounce is ONE shading kernel
// between two traversal launches (trace_queue.cu):
//
//     [extend b]   closest hit of every extension ray queued for bounce b
//     [bounce b]   grisBounceKernel: (1) fold in the light sample of vertex b-1 now that its shadow ray is known,
//                  (2) surface fetch + vertex logic of vertex b (:90-175), (3) light sample of vertex b -> shadow
//                  queue, (4) Russian roulette + BSDF sample (:226-257) -> extension queue of bounce b+1
//     [shadow b]   any hit of every queued shadow ray (overlaps nothing else of this path: consumed at bounce b+1)
//
// Deferring the light sample's contribution by one kernel is exact, not approximate: its VALUE (the three candidate
// kinds of :186-223) is computed at vertex b from the state of that moment and only its application — the
// stream-reservoir insertion / the additions into rcLi and F — waits for the visibility bit; nothing that happens in
// between (roulette, BSDF sampling, throughput update) reads what the application writes, and the next reader (the
// vertex logic of b+1, or the end of the path) runs after it, so every floating-point operation sees the same
// operands in the same order as in the shader.  A path that ends while its light sample is pending is queued once
// more without a ray ("zombie") and finishes at the next kernel.
//
// Path state lives in HBM between kernels as structure-of-arrays planes indexed by QUEUE SLOT (ping-pong over
// bounces), so every warp reads and writes full 512-byte runs; only the two rarely touched words of the path sample
// (rcIsec, rcPrevSamplePdf/rcJacobian) are indexed by pixel.  The winner of the path's streaming RIS (StreamSampler,
// :10-33) is kept directly in the pixel's output reservoir slot: it is only ever overwritten until the path ends.

constexpr int ShadeBlock = 128;
// resident blocks per SM the shading / reuse kernels are compiled for (register cap = 65536 / (128 * blocks)); they are
// bound by the latency of dependent gathers, not by issue slots or DRAM (profiles/r1_08_*), so occupancy is worth a few spills
#ifndef RT_BOUNCE_MINBLOCKS
#define RT_BOUNCE_MINBLOCKS 4
#endif
#ifndef RT_REUSE_MINBLOCKS
#define RT_REUSE_MINBLOCKS 4
#endif
constexpr uint32_t NeeNone = 0, NeeAccumulate = 1, NeeRcVertex = 2, NeeLightSampled = 3;

struct PathState {
	float3 dir;              // direction that arrived at the current vertex (wo = -dir)
	uint32_t rng;
	float3 throughput, rcThroughput, lastPos;
	float bsPdf;
	uint32_t bsType;
	uint32_t sampleState, lastSampleState;
	bool isLastVertexConnectible, isThisVertexConnectible, streamWritten, zombie;
	int bounce;
	float streamWeight, streamSumWeight;
	// light sample of the last vertex, waiting for its shadow ray
	uint32_t neeKind, shadowIdx;
	float neeResvRand;
	float4 nee0, nee1, nee2;
	// the GRISPathSample under construction; q0 / q3 are loaded from the per-pixel cold array on demand
	GRISResv ps;
	bool coldValid, coldDirty;
};

struct PathBuffers {
	float4* hot;             // PathStateWords planes of `capacity` float4 each
	float4* cold;            // 2 float4 per pixel
	uint32_t capacity;
};

RT_DEV void storePathState(const PathBuffers& b, uint32_t slot, uint32_t pix, const PathState& st) {
	const uint32_t flags = uint32_t(st.bounce) | (st.sampleState << 4) | (st.lastSampleState << 6) | (st.isLastVertexConnectible ? 1u << 8 : 0u)
		| (st.isThisVertexConnectible ? 1u << 9 : 0u) | (st.streamWritten ? 1u << 10 : 0u) | (st.zombie ? 1u << 11 : 0u) | (st.neeKind << 12) | (st.bsType << 16);
	float4* w = b.hot + slot;
	const size_t n = b.capacity;
	w[0 * n] = make_float4(st.dir.x, st.dir.y, st.dir.z, __uint_as_float(st.rng));
	w[1 * n] = make_float4(st.throughput.x, st.throughput.y, st.throughput.z, st.bsPdf);
	w[2 * n] = make_float4(st.rcThroughput.x, st.rcThroughput.y, st.rcThroughput.z, st.streamWeight);
	w[3 * n] = make_float4(st.lastPos.x, st.lastPos.y, st.lastPos.z, st.streamSumWeight);
	w[4 * n] = make_float4(__uint_as_float(flags), __uint_as_float(st.shadowIdx), st.neeResvRand, 0.f);
	w[5 * n] = st.ps.q1; w[6 * n] = st.ps.q2; w[7 * n] = st.ps.q4;
	if (st.neeKind != NeeNone) { w[8 * n] = st.nee0; w[9 * n] = st.nee1; w[10 * n] = st.nee2; }
	if (st.coldDirty) { b.cold[2 * size_t(pix)] = st.ps.q0; b.cold[2 * size_t(pix) + 1] = st.ps.q3; }
}
RT_DEV void loadPathState(const PathBuffers& b, uint32_t slot, PathState& st) {
	const float4* w = b.hot + slot;
	const size_t n = b.capacity;
	const float4 a = w[0 * n], bb = w[1 * n], c = w[2 * n], d = w[3 * n], e = w[4 * n];
	st.dir = f3(a); st.rng = __float_as_uint(a.w);
	st.throughput = f3(bb); st.bsPdf = bb.w;
	st.rcThroughput = f3(c); st.streamWeight = c.w;
	st.lastPos = f3(d); st.streamSumWeight = d.w;
	const uint32_t flags = __float_as_uint(e.x);
	st.bounce = int(flags & 15u);
	st.sampleState = (flags >> 4) & 3u; st.lastSampleState = (flags >> 6) & 3u;
	st.isLastVertexConnectible = (flags >> 8) & 1u; st.isThisVertexConnectible = (flags >> 9) & 1u;
	st.streamWritten = (flags >> 10) & 1u; st.zombie = (flags >> 11) & 1u;
	st.neeKind = (flags >> 12) & 3u;
	st.bsType = flags >> 16;
	st.shadowIdx = __float_as_uint(e.y); st.neeResvRand = e.z;
	st.ps.q1 = w[5 * n]; st.ps.q2 = w[6 * n]; st.ps.q4 = w[7 * n];
	if (st.neeKind != NeeNone) { st.nee0 = w[8 * n]; st.nee1 = w[9 * n]; st.nee2 = w[10 * n]; }
	st.ps.q5 = make_float4(0.f, 0.f, 0.f, 0.f);
	st.coldValid = false; st.coldDirty = false;
}
RT_DEV void needCold(const PathBuffers& b, uint32_t pix, PathState& st) {
	if (!st.coldValid) {
		st.ps.q0 = b.cold[2 * size_t(pix)]; st.ps.q3 = b.cold[2 * size_t(pix) + 1];
		st.coldValid = true;
	}
}

// StreamSampler::add (gris_path_trace.glsl:21-32); the selected sample goes straight to the pixel's reservoir slot
RT_DEV void streamAdd(PathState& st, const GRISResv& sample, RptGRISReservoir* __restrict__ slot, float w, float r) {
	st.streamSumWeight += w;
	if (r * st.streamSumWeight < w) {
		st.streamWeight = w;
		st.streamWritten = true;
		float4* q = reinterpret_cast<float4*>(slot);
		q[0] = sample.q0; q[1] = sample.q1; q[2] = sample.q2; q[3] = sample.q3; q[4] = sample.q4;
	}
}

// end of tracePath (gris_path_trace.glsl:259-278)
RT_DEV void finishPath(const PathBuffers& b, uint32_t pix, PathState& st, RptGRISReservoir* __restrict__ slot) {
	if (st.sampleState == 2 && st.lastSampleState == 2) {
		needCold(b, pix, st);
		streamAdd(st, st.ps, slot, luminance(st.ps.F()), sample1f(st.rng));
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

RT_DEV void setRcVertex(PathState& st, float4 isecWord, float prevSamplePdf, float jacobian) {
	st.ps.q0 = isecWord;
	st.ps.q1.w = __uint_as_float(st.rng);
	st.ps.q3 = make_float4(0.f, 0.f, prevSamplePdf, jacobian);   // the two pad words are never written: always 0
	st.coldValid = true; st.coldDirty = true;
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
				setRcVertex(st, isecWord, st.bsPdf, geometryJacobian);
				ps.setRcLi(weightedLi);
				ps.setRcWi(f3(0.0f));
				ps.setF(weightedLi * st.throughput);
				ps.setFlags(withRcVertexType(withRcVertexId(ps.flags(), uint32_t(bounce)), RcLightScattered));
				streamAdd(st, ps, slot, luminance(ps.F()), out.resvRandSample);
			}
		}
		return false;
	}
	const bool connectible = st.isThisVertexConnectible && st.isLastVertexConnectible && distToPrev > GRISDistanceThreshold;
	if ((st.sampleState == 2 && st.lastSampleState == 1) && (connectible || set.shiftType == ShiftReconnection)) {
		setRcVertex(st, isecWord, st.bsPdf, geometryJacobian);
		ps.setFlags(withRcVertexType(withRcVertexId(ps.flags(), uint32_t(bounce)), RcSurface));
		st.rcThroughput = f3(1.0f);
	}
	out.lightRandSample = sample4f(st.rng);
	out.resvRandSample = sample1f(st.rng);
	return true;
}

// NEE of the current vertex (gris_path_trace.glsl:176-223): evaluates the light sample and records what it WILL
// contribute if its shadow ray finds nothing; returns false when nothing could be contributed (no ray needed)
RT_DEV bool neePrepare(PathState& st, const Surface& surf, const Mat& mat, const LightSample& ls, float resvRandSample) {
	st.neeKind = NeeNone;
	if (!(ls.pdf > 1e-6f)) return false;
	const float3 wo = -st.dir;
	const float bsdfPdf = absDot(surf.norm, ls.wi) * RT_PI_INV;
	const float weight = MISWeight(ls.pdf, bsdfPdf);
	const float3 scatterTerm = evalBSDF(mat, surf.albedo, surf.norm, wo, ls.wi) * satDot(surf.norm, ls.wi);
	const float3 weightedLi = ls.radiance / ls.pdf * weight;
	st.neeResvRand = resvRandSample;
	if (st.sampleState == 2 && st.lastSampleState == 2) {
		const float3 dRc = weightedLi * scatterTerm * st.rcThroughput, dF = weightedLi * scatterTerm * st.throughput;
		st.neeKind = NeeAccumulate;
		st.nee0 = make_float4(dRc.x, dRc.y, dRc.z, 0.f);
		st.nee1 = make_float4(dF.x, dF.y, dF.z, 0.f);
		st.nee2 = make_float4(0.f, 0.f, 0.f, 0.f);
	}
	else if (st.sampleState == 2 && st.lastSampleState == 1) {
		const float3 F = weightedLi * scatterTerm * st.throughput;
		st.neeKind = NeeRcVertex;
		st.nee0 = make_float4(weightedLi.x, weightedLi.y, weightedLi.z, 0.f);
		st.nee1 = make_float4(ls.wi.x, ls.wi.y, ls.wi.z, 0.f);
		st.nee2 = make_float4(F.x, F.y, F.z, 0.f);
	}
	else if (st.sampleState == 1 && st.isThisVertexConnectible && ls.dist > GRISDistanceThreshold) {
		const float3 rcLi = ls.radiance * weight, F = weightedLi * scatterTerm * st.throughput;
		st.neeKind = NeeLightSampled;
		st.nee0 = make_float4(ls.bary.x, ls.bary.y, __uint_as_float(ls.id), ls.pdf);
		st.nee1 = make_float4(rcLi.x, rcLi.y, rcLi.z, ls.jacobian);
		st.nee2 = make_float4(F.x, F.y, F.z, __uint_as_float(st.rng));
	}
	return st.neeKind != NeeNone;
}

// ... and its application once the shadow ray is known to be unoccluded.  neeBounce = bounce index of that vertex.
RT_DEV void neeApply(const PathBuffers& b, uint32_t pix, PathState& st, int neeBounce, RptGRISReservoir* __restrict__ slot) {
	GRISResv& ps = st.ps;
	if (st.neeKind == NeeAccumulate) {
		ps.setRcLi(ps.rcLi() + f3(st.nee0));
		ps.setF(ps.F() + f3(st.nee1));
	}
	else if (st.neeKind == NeeRcVertex) {
		// the shader sets rcLi / rcWi / F, inserts the sample, and the BSDF-sampling step right after overwrites the
		// three fields again (:246-250); only the inserted copy ever sees them
		needCold(b, pix, st);
		GRISResv tmp = ps;
		tmp.setRcLi(f3(st.nee0)); tmp.setRcWi(f3(st.nee1)); tmp.setF(f3(st.nee2));
		streamAdd(st, tmp, slot, luminance(tmp.F()), st.neeResvRand);
	}
	else if (st.neeKind == NeeLightSampled) {
		ps.q0 = make_float4(st.nee0.x, st.nee0.y, __uint_as_float(0u), st.nee0.z);
		ps.q1.w = st.nee2.w;
		ps.q3 = make_float4(0.f, 0.f, st.nee0.w, st.nee1.w);
		st.coldValid = true; st.coldDirty = true;
		ps.setRcLi(f3(st.nee1));
		ps.setRcWi(f3(0.0f));
		ps.setF(f3(st.nee2));
		ps.setFlags(withRcVertexType(withRcVertexId(ps.flags(), uint32_t(neeBounce + 1)), RcLightSampled));
		streamAdd(st, ps, slot, luminance(ps.F()), st.neeResvRand);
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

RT_DEV PathBuffers pathBuffers(const FrameView& f, int bounce) {
	PathBuffers b;
	b.hot = f.wf.state[bounce & 1]; b.cold = f.wf.cold; b.capacity = f.wf.capacity;
	return b;
}

// light sample + scatter of the current vertex, then hand the path to the next bounce (or end it)
RT_DEV void shadeAndContinue(const FrameView& f, const SceneView& s, const RptGRISSettings& set, PathState& st, const Surface& surf, const Mat& mat,
                             const VertexOut& vo, uint32_t pix, RptGRISReservoir* __restrict__ slot) {
	const int bounce = st.bounce;
	st.neeKind = NeeNone;
	// the light sample's shadow ray; an empty interval (reported unoccluded without touching the BVH) when there is none
	float4 sh0 = make_float4(0.f, 0.f, 0.f, 1.0f), sh1 = make_float4(0.f, 0.f, 1.f, 0.0f);
	if (bounce > 0 && !isBSDFDelta(mat)) {
		const LightSample ls = sampleLight(s, surf.pos, vo.lightRandSample);
		if (neePrepare(st, surf, mat, ls, vo.resvRandSample)) {
			sh0 = make_float4(surf.pos.x, surf.pos.y, surf.pos.z, MinRayDistance);
			sh1 = make_float4(ls.wi.x, ls.wi.y, ls.wi.z, ls.dist - MinRayDistance);
		}
	}
	float3 rayOri = f3(0.0f);
	const bool go = scatterStage(st, set, surf, mat, rayOri);
	if (!go) st.bounce = bounce;   // a path that ends keeps the index of its last vertex (its light sample may still be pending)
	const PathBuffers next = pathBuffers(f, bounce + 1);
	if (!go && st.neeKind == NeeNone) {   // nothing pending: the path ends here
		finishPath(next, pix, st, slot);   // (cold words are per pixel: either PathBuffers view works)
		return;
	}
	st.zombie = !go;
	// one slot of the next bounce's queue holds the path state, its extension ray AND the shadow ray of this vertex's light
	// sample: one atomic per vertex, and the next kernel reads the visibility bit at its own slot index (coalesced)
	const uint32_t nslot = queueAppend(f.wf.counters + 4 * (bounce + 1));
	float4* rq = f.wf.rays[(bounce + 1) & 1] + 2 * size_t(nslot);
	if (go) {
		rq[0] = make_float4(rayOri.x, rayOri.y, rayOri.z, MinRayDistance);
		rq[1] = make_float4(st.dir.x, st.dir.y, st.dir.z, MaxRayDistance);
	}
	else {   // zombie: an empty interval, the traversal kernel reports a miss without touching the BVH
		rq[0] = make_float4(0.f, 0.f, 0.f, 1.0f);
		rq[1] = make_float4(0.f, 0.f, 1.f, 0.0f);
	}
	if (bounce > 0) {   // (the G-buffer vertex draws no light sample: the shadow queue of bounce 0 is never traced)
		float4* sq = f.wf.shadowRays[bounce & 1] + 2 * size_t(nslot);
		sq[0] = sh0; sq[1] = sh1;
	}
	f.wf.pix[(bounce + 1) & 1][nslot] = pix;
	if (bounce + 1 == WavefrontTailStart) { f.wf.tailList[nslot] = pix; f.wf.tailMark[pix] = f.wf.epoch; }
	storePathState(next, nslot, pix, st);
}

// bounce 0: the primary hit comes from the G-buffer and there is no light sample (gris_path_trace.glsl:176).
// One thread per pixel in 8x4-tile order.
__global__ void __launch_bounds__(ShadeBlock) grisBeginKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings set) {
	const uint32_t tilesX = (f.width + 7u) / 8u;
	const uint32_t id = blockIdx.x * ShadeBlock + threadIdx.x;
	const uint32_t tile = id >> 5, within = id & 31u;
	const uint32_t x = (tile % tilesX) * 8u + (within & 7u), y = f.rowBegin + (tile / tilesX) * 4u + (within >> 3);
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	const uint32_t pix = uint32_t(f.index(x, y));
	RptGRISReservoir* slot = f.grisThis + pix;
	if (!p.valid) {
		// background pixels keep their old reservoir (gris_path_trace.glsl:54-56) — in the reference's ping-pong pair that is the
		// reservoir of two frames ago; here the final reservoirs rotate through three buffers, so it is carried over
		const float4* old = reinterpret_cast<const float4*>(f.grisStale + pix);
		float4* q = reinterpret_cast<float4*>(slot);
		for (int i = 0; i < 6; i++) q[i] = old[i];
		return;
	}

	PathState st;
	st.dir = p.ray.dir;
	st.rng = makeSeed(f.camera.seed, x, y);
	st.throughput = f3(1.0f); st.rcThroughput = f3(0.0f); st.lastPos = f3(0.0f);
	st.bsPdf = 0.0f; st.bsType = 0;
	st.sampleState = 0; st.lastSampleState = 0;
	st.isLastVertexConnectible = false; st.isThisVertexConnectible = false; st.streamWritten = false; st.zombie = false;
	st.bounce = 0;
	st.streamWeight = 0.0f; st.streamSumWeight = 0.0f;
	st.neeKind = NeeNone; st.shadowIdx = 0; st.neeResvRand = 0.0f;
	st.nee0 = st.nee1 = st.nee2 = make_float4(0.f, 0.f, 0.f, 0.f);
	st.ps = zeroGRIS();   // GRISPathSampleReset, gris_reservoir.glsl:61-69
	st.ps.q0.z = __uint_as_float(InvalidHitIndex);
	st.ps.q4.w = __uint_as_float(st.rng);   // primaryRng
	st.coldValid = true; st.coldDirty = true;

	const Surface surf = primarySurface(p);
	const Mat mat = loadMaterial(s, uint32_t(p.matId));
	VertexOut vo;
	vo.lightRandSample = make_float4(0.f, 0.f, 0.f, 0.f); vo.resvRandSample = 0.f;
	const float4 isecWord = make_float4(0.f, 0.f, __uint_as_float(0u), __uint_as_float(0u));
	if (!vertexStage(st, set, surf, mat, isecWord, s.lightTable[0].prob, slot, vo)) {
		finishPath(pathBuffers(f, 1), pix, st, slot);
		return;
	}
	shadeAndContinue(f, s, set, st, surf, mat, vo, pix, slot);
}

// (Measured and rejected, profiles/r2_05_*: the surface fetch of every hit in a light kernel of its own at high occupancy, so that
// this kernel reads a 48-byte surface as a coalesced stream: 2.07 -> 2.10 ms per frame — the extra 96 B per vertex through HBM cost
// what the shorter gather chain saved.)
// bounce >= 1: one thread per slot of the bounce's extension queue
__global__ void __launch_bounds__(ShadeBlock, RT_BOUNCE_MINBLOCKS) grisBounceKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings set, int bounce) {
	const uint32_t n = f.wf.counters[4 * bounce];
	const float sumPower = s.lightTable[0].prob;
	const PathBuffers cur = pathBuffers(f, bounce);
	for (uint32_t slotIdx = blockIdx.x * ShadeBlock + threadIdx.x; slotIdx < n; slotIdx += gridDim.x * ShadeBlock) {
		const uint32_t pix = f.wf.pix[bounce & 1][slotIdx];
		RptGRISReservoir* slot = f.grisThis + pix;
		PathState st;
		loadPathState(cur, slotIdx, st);
		// (1) the light sample of the previous vertex
		if (st.neeKind != NeeNone) {
			if (f.wf.occluded[(bounce - 1) & 1][slotIdx] == 0) neeApply(cur, pix, st, st.zombie ? st.bounce : st.bounce - 1, slot);   // index of the vertex that drew the sample
			st.neeKind = NeeNone;
		}
		if (st.zombie) { finishPath(cur, pix, st, slot); continue; }
		// (2) the vertex the extension ray found
		const RptIntersection hit = f.wf.hits[slotIdx];
		if (hit.instanceIdx == InvalidHitIndex) { finishPath(cur, pix, st, slot); continue; }   // left the scene (:92-94)
		Surface surf;
		loadSurfaceInfo(s, hit, surf);
		const Mat mat = loadMaterial(s, surf.matIndex);
		VertexOut vo;
		vo.lightRandSample = make_float4(0.f, 0.f, 0.f, 0.f); vo.resvRandSample = 0.f;
		const float4 isecWord = make_float4(hit.bary[0], hit.bary[1], __uint_as_float(hit.instanceIdx), __uint_as_float(hit.triangleIdx));
		if (!vertexStage(st, set, surf, mat, isecWord, sumPower, slot, vo)) { finishPath(cur, pix, st, slot); continue; }
		// (3) + (4)
		shadeAndContinue(f, s, set, st, surf, mat, vo, pix, slot);
	}
}

// The tail: after bounce WavefrontTailStart - 1 only a few percent of the paths are alive (VeachAjar 1080p: 87 k of 2 M,
// halving with every bounce), and a wavefront of them is nine rounds of three nearly empty launches whose duration is the
// latency of their longest ray — 1.3 ms end to end, during which the temporal pass waits for the tail's pixels
// (profiles/r1_03_wavefront_launches.csv).  This kernel instead runs each surviving path to its end in one thread with
// in-line traversal: the same stage functions in the same order, so the same bits; SIMD efficiency is irrelevant at
// this size, the kernel takes as long as its longest path.
//
// The kernel runs on a second stream next to the temporal pass, which must not be starved: a thread holds ~170
// registers and mostly waits on memory, so the grid is kept to a few small blocks per SM (a sixth to a third of the
// register file) and every thread pulls its next path from the list as soon as one ends.
constexpr int TailBlock = 64;
// register cap of the tail kernel = 65536 / (64 * blocks).  The compiler's own choice is ~128 for the single-level code; with the
// two-level traversal compiled in as a call it would take 168 and crowd the temporal pass it runs next to (profiles/r2_09_*)
#ifndef RT_TAIL_MINBLOCKS
#define RT_TAIL_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(TailBlock, RT_TAIL_MINBLOCKS) grisTailKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings set) {
	const int first = WavefrontTailStart;
	const uint32_t n = f.wf.counters[4 * first];
	uint32_t* head = f.wf.counters + 4 * first + 2;   // (the fetch counter of the wavefront traversal of this bounce: unused here, zero)
	const float sumPower = s.lightTable[0].prob;
	const PathBuffers buf = pathBuffers(f, first);
	for (;;) {
		const uint32_t slotIdx = atomicAdd(head, 1u);
		if (slotIdx >= n) break;
		const uint32_t pix = f.wf.pix[first & 1][slotIdx];
		RptGRISReservoir* slot = f.grisThis + pix;
		PathState st;
		loadPathState(buf, slotIdx, st);
		// the two rays queued by the last wavefront bounce: the light sample's shadow ray and the extension ray
		float4 sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sh1 = sh0;
		if (st.neeKind != NeeNone) {
			const float4* rq = f.wf.shadowRays[(first - 1) & 1] + 2 * size_t(slotIdx);
			sh0 = rq[0]; sh1 = rq[1];
		}
		float3 rayOri = f3(f.wf.rays[first & 1][2 * size_t(slotIdx)]);
		for (;;) {
			// (1) the light sample of the previous vertex
			if (st.neeKind != NeeNone) {
				if (!traceShadow(s, f3(sh0), sh0.w, f3(sh1), sh1.w)) neeApply(buf, pix, st, st.zombie ? st.bounce : st.bounce - 1, slot);
				st.neeKind = NeeNone;
			}
			if (st.zombie) break;
			// (2) the vertex the extension ray finds
			const Hit h = traceClosestHit(s, rayOri, MinRayDistance, st.dir, MaxRayDistance);
			if (h.instanceIdx == InvalidHitIndex) break;
			Surface surf;
			loadSurfaceInfo(s, h, surf);
			const Mat mat = loadMaterial(s, surf.matIndex);
			VertexOut vo;
			vo.lightRandSample = make_float4(0.f, 0.f, 0.f, 0.f); vo.resvRandSample = 0.f;
			const float4 isecWord = make_float4(h.u, h.v, __uint_as_float(h.instanceIdx), __uint_as_float(h.triangleIdx));
			if (!vertexStage(st, set, surf, mat, isecWord, sumPower, slot, vo)) break;
			// (3) light sample, (4) roulette + BSDF sample: shadeAndContinue without the queues
			const int bounce = st.bounce;
			if (!isBSDFDelta(mat)) {
				const LightSample ls = sampleLight(s, surf.pos, vo.lightRandSample);
				if (neePrepare(st, surf, mat, ls, vo.resvRandSample)) {
					sh0 = make_float4(surf.pos.x, surf.pos.y, surf.pos.z, MinRayDistance);
					sh1 = make_float4(ls.wi.x, ls.wi.y, ls.wi.z, ls.dist - MinRayDistance);
				}
			}
			const bool go = scatterStage(st, set, surf, mat, rayOri);
			if (!go) {
				st.bounce = bounce;
				if (st.neeKind == NeeNone) break;
				st.zombie = true;
			}
		}
		finishPath(buf, pix, st, slot);
	}
}

// ---- temporal and spatial reuse (gris_resample_temporal.comp, gris_resample_spatial.comp) ------------------------------
//
// Each pass is a wavefront of two kernels around one launch of the queue traversal kernel:
//     [gen]    per pixel: pick the candidate reservoir(s) (reprojected pixel / 3 disk neighbours), replay the source
//              path's prefix on this pixel, reconnection geometry and validity tests -> ShiftTask + visibility ray
//     [trace]  any-hit traversal of all visibility rays (trace_queue.cu)
//     [merge]  per pixel: shifted contribution + Jacobian + reweighting of every candidate, reservoir merges in the
//              shader's order, and (spatial) the final shading of the selected sample
// The shader's random-number stream interleaves "2 numbers to place neighbour i" with "1 number for merge i-1, drawn
// only if the shifted reservoir is well-formed".  [gen] assumes the merge draw happens whenever a candidate exists —
// which is decided before the shift — and [merge] verifies it: if a shifted reservoir turns out ill-formed (NaN weight:
// only a zero-contribution sample with a valid reconnection vertex can do that) the pixel is put on the redo list and
// recomputed by the sequential per-pixel code.  Likewise the rare final samples whose shading needs replay rays
// (reconnection vertex beyond the first bounce) go to a list handled by a small kernel with in-line traversal, so the
// merge kernel itself never traverses.

constexpr int ReuseBlock = 128;

RT_DEV void storeShiftTask(const ReuseView& ru, uint32_t cand, uint32_t o, const ShiftTask& t, uint32_t srcPixel) {
	const size_t n = size_t(ru.capacity) * 3;
	float4* w = ru.task + size_t(cand) * ru.capacity + o;
	w[2 * n] = make_float4(t.rcPrevSurf.albedo.x, t.rcPrevSurf.albedo.y, t.rcPrevSurf.albedo.z, __uint_as_float(srcPixel | (t.status << 30)));
	if (t.status != TaskRay) return;
	w[0 * n] = make_float4(t.rcPrevSurf.pos.x, t.rcPrevSurf.pos.y, t.rcPrevSurf.pos.z, __uint_as_float(t.rcPrevSurf.matIndex));
	w[1 * n] = make_float4(t.rcPrevSurf.norm.x, t.rcPrevSurf.norm.y, t.rcPrevSurf.norm.z, __uint_as_float(t.rcSurf.matIndex));
	w[3 * n] = make_float4(t.rcSurf.pos.x, t.rcSurf.pos.y, t.rcSurf.pos.z, t.rcPrevWo.x);
	w[4 * n] = make_float4(t.rcSurf.norm.x, t.rcSurf.norm.y, t.rcSurf.norm.z, t.rcPrevWo.y);
	w[5 * n] = make_float4(t.rcSurf.albedo.x, t.rcSurf.albedo.y, t.rcSurf.albedo.z, t.rcPrevWo.z);
	w[6 * n] = make_float4(t.rcPrevThroughput.x, t.rcPrevThroughput.y, t.rcPrevThroughput.z, 0.f);
}
RT_DEV void storeSkipTask(const ReuseView& ru, uint32_t cand, uint32_t o) {
	ru.task[2 * size_t(ru.capacity) * 3 + size_t(cand) * ru.capacity + o] = make_float4(0.f, 0.f, 0.f, __uint_as_float(TaskSkip << 30));
}
RT_DEV uint32_t loadShiftTask(const ReuseView& ru, uint32_t cand, uint32_t o, ShiftTask& t) {   // returns the source pixel
	const size_t n = size_t(ru.capacity) * 3;
	const float4* w = ru.task + size_t(cand) * ru.capacity + o;
	const float4 c = w[2 * n];
	const uint32_t packed = __float_as_uint(c.w);
	t.status = packed >> 30;
	if (t.status == TaskRay) {
		const float4 a = w[0 * n], b = w[1 * n], d = w[3 * n], e = w[4 * n], g = w[5 * n], h = w[6 * n];
		t.rcPrevSurf.pos = f3(a); t.rcPrevSurf.matIndex = __float_as_uint(a.w);
		t.rcPrevSurf.norm = f3(b); t.rcSurf.matIndex = __float_as_uint(b.w);
		t.rcPrevSurf.albedo = f3(c); t.rcPrevSurf.isLight = false;
		t.rcSurf.pos = f3(d); t.rcSurf.norm = f3(e); t.rcSurf.albedo = f3(g); t.rcSurf.isLight = false;
		t.rcPrevWo = make_float3(d.w, e.w, g.w);
		t.rcPrevThroughput = f3(h);
	}
	return packed & 0x3fffffffu;
}
RT_DEV void storeVisibilityRay(const ReuseView& ru, uint32_t cand, uint32_t o, const ShiftTask* t) {
	float4* rq = ru.rays + 2 * (size_t(cand) * ru.capacity + o);
	if (t != nullptr && t->status == TaskRay) {
		float3 dir; float tmax;
		shiftVisibilityRay(*t, dir, tmax);
		rq[0] = make_float4(t->rcPrevSurf.pos.x, t->rcPrevSurf.pos.y, t->rcPrevSurf.pos.z, MinRayDistance);
		rq[1] = make_float4(dir.x, dir.y, dir.z, tmax);
	}
	else {   // empty interval: reported unoccluded without touching the BVH
		rq[0] = make_float4(0.f, 0.f, 0.f, 1.0f);
		rq[1] = make_float4(0.f, 0.f, 1.f, 0.0f);
	}
}

// -------- temporal ------------------------------------------------------------------------------------------------------
RT_DEV bool temporalCandidate(const FrameView& f, const RptGRISSettings& st, const Primary& p, size_t idx, Neighbor& nb) {
	if (!st.temporalReuse || (f.camera.frameIndex & 0x80000000u) != 0) return false;
	const float2 motion = f.motion[idx];
	nb = lookupSurface(f, true, make_float2(p.uv.x + motion.x, p.uv.y + motion.y));
	return nb.found && !(nb.matMeshId != p.matMeshId || dot(nb.norm, p.norm) < 0.95f || distance(p.pos, nb.pos) > 0.5f);
}

RT_DEV void temporalStore(const FrameView& f, uint32_t x, uint32_t y, size_t idx, GRISResv& resv) {
	if (!resv.valid()) resv.reset();
	storeGRIS(f.grisTemp + idx, resv);
	// multi-GPU strips: boundary rows go straight into the neighbours' halo rows over NVLink peer memory
	if (f.peerGrisUp != nullptr && rowInUpHalo(f, y)) storeGRIS(f.peerGrisUp + peerUpIndex(f, x, y), resv);
	if (f.peerGrisDown != nullptr && rowInDownHalo(f, y)) storeGRIS(f.peerGrisDown + peerDownIndex(f, x, y), resv);
}

// the spatial pass's output = next frame's "previous" reservoirs.  Boundary rows are mirrored into the neighbours' halo rows
// of the same ping-pong buffer, so a reprojection that crosses a cut (camera motion) finds its history on this GPU
RT_DEV void spatialStore(const FrameView& f, uint32_t x, uint32_t y, size_t idx, const GRISResv& resv) {
	storeGRIS(f.grisThis + idx, resv);
	if (f.peerGrisThisUp != nullptr && rowInUpHalo(f, y)) storeGRIS(f.peerGrisThisUp + peerUpIndex(f, x, y), resv);
	if (f.peerGrisThisDown != nullptr && rowInDownHalo(f, y)) storeGRIS(f.peerGrisThisDown + peerDownIndex(f, x, y), resv);
}

// the sequential per-pixel form (gris_resample_temporal.glsl:11-83), used for the pixels of the path-tracing tail
RT_DEV void grisTemporalPixel(const FrameView& f, const SceneView& s, const RptGRISSettings& st, uint32_t x, uint32_t y) {
	const Primary p = loadPrimary(f, x, y);
	if (!p.valid) return;
	const size_t idx = f.index(x, y);
	const uint32_t rng = makeSeed(f.camera.seed, x, y) ^ 1u;
	uint32_t resvRng = ~rng;
	GRISResv resv = loadGRIS(f.grisThis + idx);
	Neighbor nb;
	if (temporalCandidate(f, st, p, idx, nb)) {
		const GRISResv prev = loadGRIS(f.grisPrev + nb.pixel);
		if (prev.valid()) grisReuseAndMerge(s, st, resv, primarySurface(p), p.uv, p.ray, prev, resvRng);
	}
	temporalStore(f, x, y, idx, resv);
}

// A history sample that reconnects beyond the first bounce (paths through glass, mirrors) needs its prefix replayed from this
// pixel: a chain of dependent in-line rays, a few long ones per warp.  Such pixels (1 % of VeachAjar's) are taken out of the
// wavefront — gen marks them TaskDeferred, the merge kernel leaves them alone — and go through the sequential per-pixel form
// on the second stream (grisTemporalListKernel) while the other 99 % run through gen / visibility queue / merge without any
// traversal code in their kernels.
constexpr uint32_t TaskDeferred = 2;
__global__ void __launch_bounds__(ReuseBlock, RT_REUSE_MINBLOCKS) grisTemporalGenKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st, int skipTail) {
	const uint32_t o = blockIdx.x * ReuseBlock + threadIdx.x;
	if (o >= f.ru.capacity) return;
	const uint32_t x = o % f.width, y = f.rowBegin + o / f.width;
	const size_t idx = f.index(x, y);
	const ShiftTask* rayTask = nullptr;
	ShiftTask t;
	bool stored = false;
	if (!(skipTail && f.wf.tailMark[idx] == f.wf.epoch)) {
		const Primary p = loadPrimary(f, x, y);
		Neighbor nb;
		if (p.valid && temporalCandidate(f, st, p, idx, nb)) {
			const GRISResv prev = loadGRIS(f.grisPrev + nb.pixel);
			if (prev.valid()) {
				if (prev.sampleValid() && flagsRcVertexId(prev.flags()) != 1u) {
					f.ru.redoList[atomicAdd(f.ru.counters + 4, 1u)] = o;
					f.ru.task[2 * size_t(f.ru.capacity) * 3 + o] = make_float4(0.f, 0.f, 0.f, __uint_as_float(TaskDeferred << 30));
					stored = true;
				}
				else {
					shiftPrepare<false>(s, st, primarySurface(p), p.uv, p.ray, prev, t);
					storeShiftTask(f.ru, 0, o, t, uint32_t(nb.pixel));
					stored = true;
					rayTask = &t;
				}
			}
		}
	}
	if (!stored) storeSkipTask(f.ru, 0, o);
	storeVisibilityRay(f.ru, 0, o, rayTask);
}

__global__ void __launch_bounds__(ReuseBlock, RT_REUSE_MINBLOCKS) grisTemporalMergeKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st, int skipTail) {
	const uint32_t o = blockIdx.x * ReuseBlock + threadIdx.x;
	if (o >= f.ru.capacity) return;
	const uint32_t x = o % f.width, y = f.rowBegin + o / f.width;
	const size_t idx = f.index(x, y);
	if (skipTail && f.wf.tailMark[idx] == f.wf.epoch) return;
	if (f.depthNormal[idx].x == 0.0f) return;   // background: the shader returns before touching the reservoir
	const uint32_t rng = makeSeed(f.camera.seed, x, y) ^ 1u;
	uint32_t resvRng = ~rng;
	GRISResv resv = loadGRIS(f.grisThis + idx);
	ShiftTask t;
	const uint32_t srcPixel = loadShiftTask(f.ru, 0, o, t);
	if (t.status == TaskDeferred) return;   // grisTemporalListKernel owns this pixel
	if (t.status != TaskSkip) {
		GRISResv prev = loadGRIS(f.grisPrev + srcPixel);
		shiftFinish(s, prev, t, t.status == TaskRay && f.ru.occluded[o] == 0);
		if (prev.valid()) grisMerge(resv, prev, sample1f(resvRng));
		grisCap(resv, float(st.cap));
	}
	temporalStore(f, x, y, idx, resv);
}

__global__ void __launch_bounds__(PassBlockX* PassBlockY, 8) grisTemporalTailKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t n = f.wf.counters[4 * WavefrontTailStart];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t pix = f.wf.tailList[i];
		grisTemporalPixel(f, s, st, pix % f.width, f.storeBegin + pix / f.width);
	}
}

__global__ void __launch_bounds__(PassBlockX* PassBlockY, 8) grisTemporalListKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t n = f.ru.counters[4];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t o = f.ru.redoList[i];
		grisTemporalPixel(f, s, st, o % f.width, f.rowBegin + o / f.width);
	}
}

// -------- replay wavefront ---------------------------------------------------------------------------------------------
// A sample that reconnects beyond the first bounce (paths through glass, mirrors) can only be shifted to another pixel after
// its prefix has been replayed FROM that pixel (gris_retrace.glsl:42-136): up to 14 dependent closest-hit rays with a surface
// fetch and a BSDF sample between them.  One thread per replay with in-line traversal (the list kernels below) runs at 6-7 of 32
// lanes and a quarter of the GPU's warps, and where such samples are dense — the rows of VeachAjar's door, one strip of a
// multi-GPU film — those kernels ARE the pass (profiles/r2_23_*: 4.5 of the 5.8 ms of the spatial pass of the strip that holds
// the door).  So the replays are a wavefront of their own, shaped like the path tracer's: rwBegin*Kernel takes the first vertex
// (the pixel's primary hit, no ray), then per bounce one launch of the queue traversal kernel and one rwStepKernel (surface
// fetch + replayVertex, next ray into the next round's queue), and rwFinish*Kernel does what follows the replay in the shader
// (reconnection geometry, validity tests, visibility ray).  replayVertex is the same function the in-line form calls, in the same
// order on the same operands: the same bits.  Reconnection vertices 2..RwRounds+1 take this way (all but a handful per frame); the
// rest, and whatever exceeds the list's capacity, stay with the in-line list kernels.
constexpr int RwRounds = 6;
static_assert(RwRounds * 4 + 8 <= 64, "rwCounters holds [round][4] words");

RT_DEV void rwStoreRc(const ReuseView& ru, uint32_t k, const RcData& rc) {
	ru.rwRc[k] = make_float4(__uint_as_float(rc.prevInstance), __uint_as_float(rc.prevTriangle), rc.prevBary.x, rc.prevBary.y);
	ru.rwRc[size_t(ru.capacity) + k] = make_float4(rc.rcPrevWo.x, rc.rcPrevWo.y, rc.rcPrevWo.z, rc.rcPrevThroughput.x);
	ru.rwRc[2 * size_t(ru.capacity) + k] = make_float4(rc.rcPrevThroughput.y, rc.rcPrevThroughput.z, 0.f, 0.f);
}
RT_DEV RcData rwLoadRc(const ReuseView& ru, uint32_t k) {
	const float4 a = ru.rwRc[k], b = ru.rwRc[size_t(ru.capacity) + k], c = ru.rwRc[2 * size_t(ru.capacity) + k];
	RcData rc;
	rc.prevInstance = __float_as_uint(a.x); rc.prevTriangle = __float_as_uint(a.y); rc.prevBary = make_float2(a.z, a.w);
	rc.rcPrevWo = make_float3(b.x, b.y, b.z); rc.rcPrevThroughput = make_float3(b.w, c.x, c.y);
	return rc;
}
// the replay of list position k goes on to `round`: its ray into that round's queue
RT_DEV void rwEnqueue(const ReuseView& ru, int round, uint32_t k, uint32_t targetId, const Ray& ray, float3 throughput, uint32_t rng) {
	const uint32_t slot = queueAppend(ru.rwCounters + 4 * round);
	float4* rq = ru.rwRays[round & 1] + 2 * size_t(slot);
	rq[0] = make_float4(ray.ori.x, ray.ori.y, ray.ori.z, MinRayDistance);
	rq[1] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, MaxRayDistance);
	float4* sq = ru.rwState[round & 1] + 2 * size_t(slot);
	sq[0] = make_float4(throughput.x, throughput.y, throughput.z, __uint_as_float(rng));
	sq[1] = make_float4(__uint_as_float(k), __uint_as_float(targetId), 0.f, 0.f);
}
// first vertex of a replay (the destination pixel's primary hit: no ray)
RT_DEV void rwBegin(const FrameView& f, const SceneView& s, const RptGRISSettings& st, uint32_t k, uint32_t o, uint32_t targetFlags, uint32_t rng) {
	const Primary p = loadPrimary(f, o % f.width, f.rowBegin + o / f.width);
	const Surface surf = primarySurface(p);
	const Mat mat = loadMaterial(s, surf.matIndex);
	float3 throughput = f3(1.0f), wo = -p.ray.dir;
	Ray ray = p.ray;
	RcData rc;
	resetRcData(rc);
	const uint32_t targetId = flagsRcVertexId(targetFlags);
	if (replayVertex(st, 0, targetId, surf, mat, SpecialHitIndex, 0u, p.uv, throughput, wo, rng, ray, rc) == ReplayContinue) rwEnqueue(f.ru, 1, k, targetId, ray, throughput, rng);
	else rwStoreRc(f.ru, k, rc);   // (ended at its first vertex: invalid)
}
__global__ void __launch_bounds__(ShadeBlock, RT_REUSE_MINBLOCKS) rwStepKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st, int round) {
	const uint32_t n = f.ru.rwCounters[4 * round];
	for (uint32_t slot = blockIdx.x * ShadeBlock + threadIdx.x; slot < n; slot += gridDim.x * ShadeBlock) {
		const float4 a = f.ru.rwState[round & 1][2 * size_t(slot)], b = f.ru.rwState[round & 1][2 * size_t(slot) + 1];
		const uint32_t k = __float_as_uint(b.x), targetId = __float_as_uint(b.y);
		const RptIntersection hit = f.ru.rwHits[slot];
		RcData rc;
		resetRcData(rc);
		if (hit.instanceIdx != InvalidHitIndex) {
			float3 throughput = f3(a);
			uint32_t rng = __float_as_uint(a.w);
			Ray ray;
			ray.ori = f3(0.0f);
			ray.dir = f3(f.ru.rwRays[round & 1][2 * size_t(slot) + 1]);
			float3 wo = -ray.dir;
			Surface surf;
			loadSurfaceInfo(s, hit, surf);
			const Mat mat = loadMaterial(s, surf.matIndex);
			const int res = replayVertex(st, round, targetId, surf, mat, hit.instanceIdx, hit.triangleIdx, make_float2(hit.bary[0], hit.bary[1]), throughput, wo, rng, ray, rc);
			if (res == ReplayContinue && round < RwRounds) { rwEnqueue(f.ru, round + 1, k, targetId, ray, throughput, rng); continue; }
			if (res != ReplayFound) resetRcData(rc);
		}
		rwStoreRc(f.ru, k, rc);
	}
}
// spatial pass: (pixel, neighbour) pairs of the wavefront replay list
__global__ void __launch_bounds__(ShadeBlock, RT_REUSE_MINBLOCKS) rwBeginShiftKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t n = min(f.ru.counters[5], f.ru.capacity);
	for (uint32_t k = blockIdx.x * ShadeBlock + threadIdx.x; k < n; k += gridDim.x * ShadeBlock) {
		const uint32_t e = f.ru.rwList[k], o = e % f.ru.capacity;
		const uint32_t srcPixel = __float_as_uint(f.ru.task[2 * size_t(f.ru.capacity) * 3 + size_t(e)].w) & 0x3fffffffu;
		const float4* q = reinterpret_cast<const float4*>(f.grisTemp + srcPixel);
		rwBegin(f, s, st, k, o, __float_as_uint(q[2].w), __float_as_uint(q[4].w));
	}
}
__global__ void __launch_bounds__(ShadeBlock, RT_REUSE_MINBLOCKS) rwFinishShiftKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t n = min(f.ru.counters[5], f.ru.capacity);
	for (uint32_t k = blockIdx.x * ShadeBlock + threadIdx.x; k < n; k += gridDim.x * ShadeBlock) {
		const uint32_t e = f.ru.rwList[k], i = e / f.ru.capacity, o = e % f.ru.capacity;
		const uint32_t srcPixel = __float_as_uint(f.ru.task[2 * size_t(f.ru.capacity) * 3 + size_t(e)].w) & 0x3fffffffu;
		const GRISResv nr = loadGRIS(f.grisTemp + srcPixel);
		const Primary p = loadPrimary(f, o % f.width, f.rowBegin + o / f.width);
		ShiftTask t;
		t.status = TaskInvalid;
		if (nr.sampleValid()) shiftPrepareFromRc(s, primarySurface(p), nr, rwLoadRc(f.ru, k), t);
		storeShiftTask(f.ru, i, o, t, srcPixel);
		storeVisibilityRay(f.ru, i, o, &t);
	}
}

// -------- spatial -------------------------------------------------------------------------------------------------------
RT_DEV bool spatialCandidate(const FrameView& f, const Primary& p, uint32_t& rng, Neighbor& nb) {   // gris_resample_spatial.glsl:62-84
	const float texelX = 1.0f / float(f.width), texelY = 1.0f / float(f.height);
	const float2 d = toConcentricDisk(sample2f(rng));
	const float2 nuv = make_float2(p.uv.x + d.x * 20.0f * texelX, p.uv.y + d.y * 20.0f * texelY);
	nb = lookupSurface(f, false, nuv);
	return nb.found && !(dot(nb.norm, p.norm) < 0.9f || distance(p.pos, nb.pos) > 0.4f);
}

// final shading of the selected sample (gris_resample_spatial.glsl:88-131); replays the path prefix
template <bool CanTrace = true>
RT_DEV float3 spatialShade(const SceneView& s, const RptGRISSettings& st, const Primary& p, const Surface& dstPrimarySurf, GRISResv& resv) {
	float3 radiance = f3(0.0f);
	if (resv.valid() && resv.sampleCount() > 0 && resv.sampleValid()) {
		RcData rc;
		traceReplayPath<CanTrace>(s, st, dstPrimarySurf, p.uv, p.ray, resv.flags(), resv.primaryRng(), rc);
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
	return clampColor(radiance);
}

// The same shading from the reconnection data a shift left behind (the reference's GRISReconnectionData, layouts.glsl:158-164,
// which its unfinished retrace pass was to store once per sample, gris_retrace.glsl:238-320): when the selected sample is a
// neighbour's, the shift that brought it to this pixel has already replayed its prefix HERE — same primary surface, same flags,
// same random numbers as the replay of spatialShade — and its ShiftTask holds the result: both ends of the reconnection segment,
// rcPrevWo and rcPrevThroughput.  (A task only wins a merge if its status was TaskRay and its ray unoccluded, so the replay found a
// vertex.)  Same functions on the same operands: the same bits as spatialShade, without replay rays or surface fetches.
RT_DEV float3 spatialShadeFromTask(const SceneView& s, const ShiftTask& t, GRISResv& resv) {
	float3 radiance = f3(0.0f);
	if (resv.valid() && resv.sampleCount() > 0 && resv.sampleValid()) {
		RcData rc;
		rc.prevInstance = 0; rc.prevTriangle = 0; rc.prevBary = make_float2(0.f, 0.f);
		rc.rcPrevWo = t.rcPrevWo; rc.rcPrevThroughput = t.rcPrevThroughput;
		const Mat rcPrevMat = loadMaterial(s, t.rcPrevSurf.matIndex);
		const float3 wi = normalize(t.rcSurf.pos - t.rcPrevSurf.pos);
		const float3 Li = reconnectionLi(s, resv, rc, t.rcPrevSurf, t.rcSurf, rcPrevMat, wi, resv.rcPrevSamplePdf());
		if (!isBlack(Li) && !hasNan(Li)) radiance = Li / luminance(Li) * resv.resampleWeight() / resv.sampleCount();
	}
	return clampColor(radiance);
}

// the sequential per-pixel form (gris_resample_spatial.glsl:11-134), used for the redo list
RT_DEV void grisSpatialPixel(const FrameView& f, const SceneView& s, const RptGRISSettings& st, uint32_t x, uint32_t y) {
	const Primary p = loadPrimary(f, x, y);
	float3 radiance = f3(0.0f);
	if (p.valid) {
		const size_t idx = f.index(x, y);
		uint32_t rng = makeSeed(f.camera.seed, x, y) ^ 2u;
		GRISResv resv = loadGRIS(f.grisTemp + idx);
		const Surface dstPrimarySurf = primarySurface(p);
		if (st.spatialReuse) {
			for (uint32_t i = 0; i < 3; i++) {
				Neighbor nb;
				if (spatialCandidate(f, p, rng, nb)) {
					const GRISResv nr = loadGRIS(f.grisTemp + nb.pixel);
					if (nr.valid()) grisReuseAndMerge(s, st, resv, dstPrimarySurf, p.uv, p.ray, nr, rng);
				}
			}
		}
		if (!resv.valid()) resv.reset();
		spatialStore(f, x, y, idx, resv);
		radiance = spatialShade(s, st, p, dstPrimarySurf, resv);
	}
	accumulate(f.indirectOutput, f, x, y, radiance);
}

// [gen] of the spatial pass is itself two kernels.  Placing the three neighbours is a sequential chain per pixel (the
// random numbers of neighbour i+1 depend on whether neighbour i's reservoir was well-formed) but a short one: disk sample,
// bilinear G-buffer lookup, one 16-byte word of the neighbour's reservoir.  The shift of each chosen neighbour — reservoir
// fetch, replay, two surface fetches (instance -> indices -> vertices -> material -> texture) — is independent of the other
// two, so it runs as one thread per (pixel, neighbour): three times the loads in flight for the same work
// (profiles/r1_17_*: the one-kernel form spent 40 % of its stall samples waiting for the neighbour's reservoir).
constexpr uint32_t TaskPending = 2;   // grisSpatialPickKernel -> grisSpatialShiftKernel: neighbour chosen, shift not yet prepared
constexpr uint32_t TaskReplay = 1;    // the same for grisSpatialShiftListKernel (the value of TaskRay, which only exists after the shift)

__global__ void __launch_bounds__(ReuseBlock) grisSpatialPickKernel(const __grid_constant__ FrameView f, const RptGRISSettings st) {
	const uint32_t o = blockIdx.x * ReuseBlock + threadIdx.x;
	if (o >= f.ru.capacity) return;
	const uint32_t x = o % f.width, y = f.rowBegin + o / f.width;
	const Primary p = loadPrimary(f, x, y);
	const bool active = p.valid && st.spatialReuse != 0;
	uint32_t rng = makeSeed(f.camera.seed, x, y) ^ 2u;
	float4* word = f.ru.task + 2 * size_t(f.ru.capacity) * 3 + o;
	for (uint32_t i = 0; i < 3; i++) {
		uint32_t packed = TaskSkip << 30;
		Neighbor nb;
		if (active && spatialCandidate(f, p, rng, nb)) {
			const float4* q = reinterpret_cast<const float4*>(f.grisTemp + nb.pixel);
			const float w = q[5].y;   // GRISResv::valid()
			if (!isnan_(w) && w >= 0) {
				// a source sample that reconnects beyond the first bounce needs replay rays: grisSpatialShiftListKernel
				const bool replay = __float_as_uint(q[0].z) != InvalidHitIndex && flagsRcVertexId(__float_as_uint(q[2].w)) != 1u;
				packed = uint32_t(nb.pixel) | ((replay ? TaskReplay : TaskPending) << 30);
				if (replay) {
					const uint32_t id = flagsRcVertexId(__float_as_uint(q[2].w));
					uint32_t place = 0xffffffffu;
					if (!f.ru.noReplayWavefront && id <= uint32_t(RwRounds) + 1u) place = atomicAdd(f.ru.counters + 5, 1u);
					if (place < f.ru.capacity) f.ru.rwList[place] = i * f.ru.capacity + o;                 // replay wavefront
					else f.ru.shadeList[atomicAdd(f.ru.counters + 3, 1u)] = i * f.ru.capacity + o;       // in-line list kernel
				}
				sample1f(rng);   // the merge's random number, assumed drawn (verified in the merge kernel)
			}
		}
		word[size_t(i) * f.ru.capacity] = make_float4(0.f, 0.f, 0.f, __uint_as_float(packed));
	}
}

template <bool CanTrace>
RT_DEV void spatialShiftOne(const FrameView& f, const SceneView& s, const RptGRISSettings& st, uint32_t i, uint32_t o, uint32_t srcPixel, const GRISResv& nr) {
	const Primary p = loadPrimary(f, o % f.width, f.rowBegin + o / f.width);
	ShiftTask t;
	shiftPrepare<CanTrace>(s, st, primarySurface(p), p.uv, p.ray, nr, t);
	storeShiftTask(f.ru, i, o, t, srcPixel);
	storeVisibilityRay(f.ru, i, o, &t);
}

// one thread per (pixel, neighbour); blockIdx.y = neighbour index.  Source samples that reconnect at the first bounce
// (most) need no replay ray and are shifted here; the others are on the list of grisSpatialShiftListKernel, which runs
// next to this kernel on a second stream — so this kernel carries no traversal code (80 registers instead of 128) and the
// replay rays run in full warps instead of 4 lanes of 32.
__global__ void __launch_bounds__(ReuseBlock, RT_REUSE_MINBLOCKS) grisSpatialShiftKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t o = blockIdx.x * ReuseBlock + threadIdx.x, i = blockIdx.y;
	if (o >= f.ru.capacity) return;
	const uint32_t packed = __float_as_uint(f.ru.task[2 * size_t(f.ru.capacity) * 3 + size_t(i) * f.ru.capacity + o].w);
	if ((packed >> 30) == TaskReplay) return;
	if ((packed >> 30) != TaskPending) {
		storeVisibilityRay(f.ru, i, o, nullptr);
		return;
	}
	const uint32_t srcPixel = packed & 0x3fffffffu;
	spatialShiftOne<false>(f, s, st, i, o, srcPixel, loadGRIS(f.grisTemp + srcPixel));
}

__global__ void __launch_bounds__(ReuseBlock, RT_REUSE_MINBLOCKS) grisSpatialShiftListKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t n = f.ru.counters[3];
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		const uint32_t e = f.ru.shadeList[k], i = e / f.ru.capacity, o = e % f.ru.capacity;
		const uint32_t srcPixel = __float_as_uint(f.ru.task[2 * size_t(f.ru.capacity) * 3 + size_t(e)].w) & 0x3fffffffu;
		spatialShiftOne<true>(f, s, st, i, o, srcPixel, loadGRIS(f.grisTemp + srcPixel));
	}
}

__global__ void __launch_bounds__(ReuseBlock, RT_REUSE_MINBLOCKS) grisSpatialMergeKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t o = blockIdx.x * ReuseBlock + threadIdx.x;
	if (o >= f.ru.capacity) return;
	const uint32_t x = o % f.width, y = f.rowBegin + o / f.width;
	const Primary p = loadPrimary(f, x, y);
	float3 radiance = f3(0.0f);
	if (p.valid) {
		const size_t idx = f.index(x, y);
		uint32_t rng = makeSeed(f.camera.seed, x, y) ^ 2u;
		GRISResv resv = loadGRIS(f.grisTemp + idx);
		int winner = -1;   // the neighbour whose sample the reservoir holds in the end (-1: the pixel's own)
		if (st.spatialReuse) {
			for (uint32_t i = 0; i < 3; i++) {
				sample2f(rng);   // the two numbers that placed neighbour i
				ShiftTask t;
				const uint32_t srcPixel = loadShiftTask(f.ru, i, o, t);
				if (t.status == TaskSkip) continue;
				GRISResv nr = loadGRIS(f.grisTemp + srcPixel);
				shiftFinish(s, nr, t, t.status == TaskRay && f.ru.occluded[size_t(i) * f.ru.capacity + o] == 0);
				if (!nr.valid()) {   // no random number is drawn for an ill-formed reservoir: the assumed sequence is off from here
					f.ru.redoList[atomicAdd(f.ru.counters + 1, 1u)] = o;
					return;
				}
				if (grisMerge(resv, nr, sample1f(rng))) winner = int(i);
				grisCap(resv, float(st.cap));
			}
		}
		if (!resv.valid()) resv.reset();
		spatialStore(f, x, y, idx, resv);
		if (winner >= 0 && !f.ru.noShadeFromTask) {   // a neighbour's sample: its shift has left the reconnection data of this pixel behind
			ShiftTask t;
			loadShiftTask(f.ru, uint32_t(winner), o, t);
			radiance = spatialShadeFromTask(s, t, resv);
		}
		else {
			if (resv.valid() && resv.sampleCount() > 0 && resv.sampleValid() && flagsRcVertexId(resv.flags()) != 1u) {
				f.ru.shadeList[atomicAdd(f.ru.counters + 0, 1u)] = o;   // shading needs replay rays: grisSpatialShadeListKernel
				return;
			}
			radiance = spatialShade<false>(s, st, p, primarySurface(p), resv);   // rcVertexId == 1: the replay is the primary hit itself, no ray
		}
	}
	accumulate(f.indirectOutput, f, x, y, radiance);
}

__global__ void __launch_bounds__(PassBlockX* PassBlockY) grisSpatialShadeListKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t n = f.ru.counters[0];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t o = f.ru.shadeList[i];
		const uint32_t x = o % f.width, y = f.rowBegin + o / f.width;
		const Primary p = loadPrimary(f, x, y);
		GRISResv resv = loadGRIS(f.grisThis + f.index(x, y));
		accumulate(f.indirectOutput, f, x, y, spatialShade(s, st, p, primarySurface(p), resv));
	}
}
__global__ void __launch_bounds__(PassBlockX* PassBlockY) grisSpatialRedoKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptGRISSettings st) {
	const uint32_t n = f.ru.counters[1];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t o = f.ru.redoList[i];
		grisSpatialPixel(f, s, st, o % f.width, f.rowBegin + o / f.width);
	}
}

void launchGRISPathTraceBounces(const FrameView& f, const SceneView& s, const RptGRISSettings& p, int firstBounce, int lastBounce, cudaStream_t st,
                                KernelClock* clock, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join) {
	const bool twoStreams = side != nullptr && fork != nullptr && join != nullptr;
	static const int bounceBlocks = persistentBlocks(reinterpret_cast<const void*>(grisBounceKernel), ShadeBlock);
	if (firstBounce == 0) {
		const uint32_t rows = f.rowEnd - f.rowBegin;
		const uint32_t slots = ((f.width + 7u) / 8u) * ((rows + 3u) / 4u) * 32u;
		cudaMemsetAsync(f.wf.counters, 0, size_t(WavefrontMaxBounces) * 4 * sizeof(uint32_t), st);
		if (clock) clock->tick(RPT_KERNEL_GRIS_BEGIN);
		grisBeginKernel<<<(slots + ShadeBlock - 1) / ShadeBlock, ShadeBlock, 0, st>>>(f, s, p);
		firstBounce = 1;
	}
	// bounce 15 only drains the paths whose last light sample is still pending
	for (int bounce = firstBounce; bounce <= lastBounce; bounce++) {
		uint32_t* c = f.wf.counters + 4 * bounce;
		// the shadow rays of vertex b-1 and the extension rays of bounce b are independent: on two streams the drain of the
		// one kernel (its last, longest rays) overlaps the body of the other
		const bool overlap = twoStreams && bounce > 1 && bounce < 15;
		if (clock && bounce > 1) clock->tick(overlap ? RPT_KERNEL_TRACE_PAIR : RPT_KERNEL_TRACE_ANY);
		if (overlap) { cudaEventRecord(fork, st); cudaStreamWaitEvent(side, fork, 0); }
		if (bounce > 1) launchTraceQueueAny(s, f.wf.shadowRays[(bounce - 1) & 1], c + 0, 0, c - 4 + 3, f.wf.occluded[(bounce - 1) & 1], overlap ? side : st);   // (slot-aligned with this bounce's queue)
		if (overlap) cudaEventRecord(join, side);
		if (clock && bounce < 15 && !overlap) clock->tick(RPT_KERNEL_TRACE_CLOSEST);
		if (bounce < 15) launchTraceQueueClosest(s, f.wf.rays[bounce & 1], c + 0, 0, c + 2, f.wf.hits, st);
		if (overlap) cudaStreamWaitEvent(st, join, 0);
		if (clock) clock->tick(RPT_KERNEL_GRIS_BOUNCE);
		grisBounceKernel<<<bounceBlocks, ShadeBlock, 0, st>>>(f, s, p, bounce);
	}
}
// every path still alive after bounce WavefrontTailStart - 1, to its end
void launchGRISPathTraceTail(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st) {
	static const int blocks = [] {
		int dev = 0, sms = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		const char* e = getenv("RPT_TAIL_BLOCKS_PER_SM");   // experiments; default = measured optimum (profiles/README.md)
		return (sms > 0 ? sms : 148) * (e ? std::max(1, atoi(e)) : 4);
	}();
	grisTailKernel<<<blocks, TailBlock, 0, st>>>(f, s, p);
}
void launchGRISTemporal(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st, int tailMode, KernelClock* clock,
                        cudaStream_t side, cudaEvent_t fork) {
	static const int blocks = persistentBlocks(reinterpret_cast<const void*>(grisTemporalTailKernel), PassBlockX * PassBlockY);
	if (tailMode == 2) {
		if (clock) clock->tick(RPT_KERNEL_REUSE_MERGE);
		grisTemporalTailKernel<<<blocks, PassBlockX * PassBlockY, 0, st>>>(f, s, p);
		return;
	}
	const uint32_t n = f.ru.capacity, dense = (n + ReuseBlock - 1) / ReuseBlock;
	cudaMemsetAsync(f.ru.counters + 2, 0, 3 * sizeof(uint32_t), st);   // ([0], [1]: the previous frame's shade list may still be running)
	if (clock) clock->tick(RPT_KERNEL_REUSE_GEN);
	grisTemporalGenKernel<<<dense, ReuseBlock, 0, st>>>(f, s, p, tailMode);
	// the deferred pixels (replay chains) on the second stream, next to the visibility rays and the merge of all others; the
	// caller joins `side` before anything reads the pass's output
	if (side != nullptr && fork != nullptr) {
		cudaEventRecord(fork, st);
		cudaStreamWaitEvent(side, fork, 0);
		grisTemporalListKernel<<<blocks, PassBlockX * PassBlockY, 0, side>>>(f, s, p);
	}
	if (clock) clock->tick(RPT_KERNEL_TRACE_ANY);
	launchTraceQueueAny(s, f.ru.rays, nullptr, n, f.ru.counters + 2, f.ru.occluded, st);
	if (clock) clock->tick(RPT_KERNEL_REUSE_MERGE);
	grisTemporalMergeKernel<<<dense, ReuseBlock, 0, st>>>(f, s, p, tailMode);
	if (side == nullptr || fork == nullptr) grisTemporalListKernel<<<blocks, PassBlockX * PassBlockY, 0, st>>>(f, s, p);
}
void launchGRISSpatial(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st, KernelClock* clock,
                       cudaStream_t side, cudaEvent_t fork, cudaEvent_t join, cudaStream_t side2, cudaEvent_t join2,
                       cudaStream_t shadeStream, cudaEvent_t shadeFork, cudaEvent_t shadeDone) {
	static const int listBlocks = persistentBlocks(reinterpret_cast<const void*>(grisSpatialRedoKernel), PassBlockX * PassBlockY);
	const bool twoStreams = side != nullptr && fork != nullptr && join != nullptr;
	// next to the dense kernel the list kernel gets a part of every SM (blocks of 128 threads x 128 registers: a quarter of
	// the register file each), alone all of it
	static const int shiftListFull = persistentBlocks(reinterpret_cast<const void*>(grisSpatialShiftListKernel), ReuseBlock);
	static const int shiftListPart = [] {
		int dev = 0, sms = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		const char* e = getenv("RPT_SHIFTLIST_BLOCKS_PER_SM");   // experiments; default = measured optimum (profiles/README.md)
		return (sms > 0 ? sms : 148) * (e ? std::max(1, atoi(e)) : 3);
	}();
	const int shiftListBlocks = twoStreams ? shiftListPart : shiftListFull;
	const uint32_t n = f.ru.capacity, blocks = (n + ReuseBlock - 1) / ReuseBlock;
	cudaMemsetAsync(f.ru.counters, 0, 16 * sizeof(uint32_t), st);
	cudaMemsetAsync(f.ru.rwCounters, 0, 64 * sizeof(uint32_t), st);
	if (clock) clock->tick(RPT_KERNEL_REUSE_GEN);
	grisSpatialPickKernel<<<blocks, ReuseBlock, 0, st>>>(f, p);
	// the replays (wavefront rounds, then the few long ones as in-line chains) next to the dense shift kernel
	if (twoStreams) { cudaEventRecord(fork, st); cudaStreamWaitEvent(side, fork, 0); }
	// (with a third stream the wavefront rounds and the few long in-line chains run next to each other as well)
	const bool threeStreams = twoStreams && side2 != nullptr && join2 != nullptr;
	if (threeStreams) cudaStreamWaitEvent(side2, fork, 0);
	cudaStream_t rs = threeStreams ? side2 : (twoStreams ? side : st);
	if (!f.ru.noReplayWavefront) {
		static const int rwBlocks = persistentBlocks(reinterpret_cast<const void*>(rwStepKernel), ShadeBlock);
		rwBeginShiftKernel<<<rwBlocks, ShadeBlock, 0, rs>>>(f, s, p);
		for (int round = 1; round <= RwRounds; round++) {
			uint32_t* c = f.ru.rwCounters + 4 * round;
			launchTraceQueueClosest(s, f.ru.rwRays[round & 1], c + 0, 0, c + 2, f.ru.rwHits, rs);
			rwStepKernel<<<rwBlocks, ShadeBlock, 0, rs>>>(f, s, p, round);
		}
		rwFinishShiftKernel<<<rwBlocks, ShadeBlock, 0, rs>>>(f, s, p);
	}
	if (threeStreams) cudaEventRecord(join2, side2);
	grisSpatialShiftListKernel<<<shiftListBlocks, ReuseBlock, 0, twoStreams ? side : st>>>(f, s, p);
	if (twoStreams) cudaEventRecord(join, side);
	grisSpatialShiftKernel<<<dim3(blocks, 3), ReuseBlock, 0, st>>>(f, s, p);
	if (twoStreams) cudaStreamWaitEvent(st, join, 0);
	if (threeStreams) cudaStreamWaitEvent(st, join2, 0);
	if (clock) clock->tick(RPT_KERNEL_TRACE_ANY);
	launchTraceQueueAny(s, f.ru.rays, nullptr, 3 * n, f.ru.counters + 2, f.ru.occluded, st);
	if (clock) clock->tick(RPT_KERNEL_REUSE_MERGE);
	grisSpatialMergeKernel<<<blocks, ReuseBlock, 0, st>>>(f, s, p);
	grisSpatialRedoKernel<<<listBlocks, PassBlockX * PassBlockY, 0, st>>>(f, s, p);
	// The shade list — the selected samples whose final shading needs replay rays: a latency chain — gates nothing but the image:
	// the reservoirs are complete after the merge.  With a stream to itself it runs next to whatever follows the pass on `st` (the
	// next frame's temporal pass); the caller orders the post-process and the next spatial pass behind `shadeDone`.
	if (shadeStream != nullptr && shadeFork != nullptr && shadeDone != nullptr) {
		cudaEventRecord(shadeFork, st);
		cudaStreamWaitEvent(shadeStream, shadeFork, 0);
		grisSpatialShadeListKernel<<<listBlocks, PassBlockX * PassBlockY, 0, shadeStream>>>(f, s, p);
		cudaEventRecord(shadeDone, shadeStream);
	}
	else grisSpatialShadeListKernel<<<listBlocks, PassBlockX * PassBlockY, 0, st>>>(f, s, p);
}

} // namespace rt
