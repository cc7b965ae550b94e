// ReSTIR GI (BASELINE.json config 5), as a wavefront.
//   reference src/shader/gi_resample_temporal.glsl:9-212 (+ .comp), gi_reservoir.glsl:8-50
//
// The shader runs one invocation per pixel: a path of up to 15 bounces with two ray queries per bounce
// (gi_resample_temporal.glsl:61-170), then the temporal reservoir update (:172-193) and the final shading of the
// selected sample behind one visibility ray (:195-209).  As one kernel that is 38 ms per 1080p frame on the 51 M-triangle
// field (profiles/r1_04_config5_*): the rays of a warp need very different numbers of BVH steps.  Here the loop is cut
// at its ray queries, exactly like the ReSTIR PT path tracer (passes_gris.cu), and shares its queues:
//
//     giBeginKernel      bounce 0 (the G-buffer vertex: no light sample, no roulette), BSDF sample -> extension queue 1
//     [extend b]         closest hit of every extension ray of bounce b               (trace_queue.cu)
//     giBounceKernel b   (1) add the light sample of vertex b-1 now that its shadow ray is known, (2) surface fetch,
//                        emitter hit, (3) light sample of vertex b -> shadow queue, (4) roulette + BSDF sample ->
//                        extension queue b+1
//     [shadow b]         any hit of every shadow ray of bounce b
//     giResolveKernel    per pixel: temporal reservoir lookup + update, cap; candidate final shading -> visibility queue
//     [visibility]       any hit
//     giShadeKernel      per pixel: pick the shaded colour, accumulate
//
// Deferring a light sample's addition into rcLo by one kernel is exact: the summand is computed at vertex b from the
// state of that moment; nothing between the shader's addition and the next one (at vertex b+1) reads rcLo, so every
// floating-point operation sees the same operands in the same order.  A path that ends while its light sample is
// pending is queued once more without a ray ("zombie") and finishes in the next kernel.
//
// Per-slot path state: 5 planes of the wavefront state buffer.  Per-pixel records (owned-pixel index o): planes of
// the reuse task buffer — R0 {rcLo, rng at the end of the path}, R1 {primaryScatter, primaryPdf}, R2 psIsec,
// T0 {colour if the final sample is not used, flag}, T1 {colour if it is visible}.
#include "passes.h"
#include "shading.cuh"
#include "persist.cuh"

namespace rt {

namespace {

constexpr int GIBlock = 128;

// GIReservoir (48 B): q0 = rcIsec, q1 = {rcLo, rcPrevCoord}, q2 = {sampleCount, resampleWeight, contribWeight, pad}
struct GIResv {
	float4 q0, q1, q2;
	RT_DEV uint32_t sampleCount() const { return __float_as_uint(q2.x); }
	RT_DEV void setSampleCount(uint32_t c) { q2.x = __uint_as_float(c); }
	RT_DEV bool valid() const { return !isnan_(q2.y) && q2.y >= 0; }
	RT_DEV void reset() { setSampleCount(0); q2.y = 0.0f; q2.z = 0.0f; }
};

struct GIPath {
	float3 dir;               // direction that arrived at the current vertex (wo = -dir)
	uint32_t rng;
	float3 throughputAfter, lastPos, rcLo;
	float bsPdf;
	uint32_t bsType;
	int bounce;
	bool neePending, zombie;
	uint32_t shadowIdx;
	float3 nee;               // the pending light sample's summand
};

RT_DEV float4* recordPlane(const FrameView& f, int plane) { return f.ru.task + size_t(plane) * f.ru.capacity; }
RT_DEV uint32_t ownedIndex(const FrameView& f, uint32_t pix) { return pix - (f.rowBegin - f.storeBegin) * f.width; }

RT_DEV void storePath(const FrameView& f, int parity, uint32_t slot, const GIPath& p) {
	float4* w = f.wf.state[parity] + slot;
	const size_t n = f.wf.capacity;
	const uint32_t flags = uint32_t(p.bounce) | (p.neePending ? 1u << 4 : 0u) | (p.zombie ? 1u << 5 : 0u) | (p.bsType << 8);
	w[0 * n] = make_float4(p.dir.x, p.dir.y, p.dir.z, __uint_as_float(p.rng));
	w[1 * n] = make_float4(p.throughputAfter.x, p.throughputAfter.y, p.throughputAfter.z, p.bsPdf);
	w[2 * n] = make_float4(p.lastPos.x, p.lastPos.y, p.lastPos.z, __uint_as_float(flags));
	w[3 * n] = make_float4(p.rcLo.x, p.rcLo.y, p.rcLo.z, __uint_as_float(p.shadowIdx));
	if (p.neePending) w[4 * n] = make_float4(p.nee.x, p.nee.y, p.nee.z, 0.f);
}
RT_DEV void loadPath(const FrameView& f, int parity, uint32_t slot, GIPath& p) {
	const float4* w = f.wf.state[parity] + slot;
	const size_t n = f.wf.capacity;
	const float4 a = w[0 * n], b = w[1 * n], c = w[2 * n], d = w[3 * n];
	p.dir = f3(a); p.rng = __float_as_uint(a.w);
	p.throughputAfter = f3(b); p.bsPdf = b.w;
	p.lastPos = f3(c);
	const uint32_t flags = __float_as_uint(c.w);
	p.bounce = int(flags & 15u); p.neePending = (flags >> 4) & 1u; p.zombie = (flags >> 5) & 1u; p.bsType = flags >> 8;
	p.rcLo = f3(d); p.shadowIdx = __float_as_uint(d.w);
	p.nee = p.neePending ? f3(w[4 * n]) : f3(0.0f);
}

// the path has ended: what the rest of the shader needs from it
RT_DEV void finishPath(const FrameView& f, uint32_t o, const GIPath& p) {
	recordPlane(f, 0)[o] = make_float4(p.rcLo.x, p.rcLo.y, p.rcLo.z, __uint_as_float(p.rng));
}

// roulette + BSDF sample of the current vertex (gi_resample_temporal.glsl:146-169), then hand the path to the next bounce
// sh0 / sh1 = the shadow ray of this vertex's light sample (an empty interval when there is none): it shares the slot of the
// next bounce's queue with the path state and the extension ray
RT_DEV void scatterAndContinue(const FrameView& f, GIPath& p, const Surface& surf, const Mat& mat, uint32_t pix, uint32_t o, float4 sh0, float4 sh1) {
	const int bounce = p.bounce;
	bool go = true;
	float3 rayOri = f3(0.0f);
	if (bounce > 4) {
		const float pdfTerminate = max_(1.0f - luminance(p.throughputAfter), 0.0f);
		if (sample1f(p.rng) < pdfTerminate) go = false;
		else p.throughputAfter /= (1.0f - pdfTerminate);
	}
	if (go) {
		const float3 r3 = sample3f(p.rng);
		BSDFSample bs = emptyBSDFSample();
		bs.pdf = p.bsPdf; bs.type = p.bsType;
		if (!sampleBSDF(mat, surf.albedo, surf.norm, -p.dir, r3, bs) || bs.pdf < 1e-6f) go = false;
		else {
			p.bsPdf = bs.pdf; p.bsType = bs.type;
			const float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
			if (bounce == 0) {
				const float3 primaryScatter = bs.bsdf * cosTheta;
				recordPlane(f, 1)[o] = make_float4(primaryScatter.x, primaryScatter.y, primaryScatter.z, bs.pdf);
			}
			else {
				p.throughputAfter *= bs.bsdf * cosTheta / bs.pdf;
			}
			p.lastPos = surf.pos;
			p.dir = bs.wi;
			rayOri = surf.pos + p.dir * 1e-4f;
			go = bounce + 1 < 15;
		}
	}
	if (!go && !p.neePending) { finishPath(f, o, p); return; }
	p.zombie = !go;
	if (go) p.bounce = bounce + 1;
	const int parity = (bounce + 1) & 1;
	const uint32_t nslot = queueAppend(f.wf.counters + 4 * (bounce + 1));
	float4* rq = f.wf.rays[parity] + 2 * size_t(nslot);
	if (go) {
		rq[0] = make_float4(rayOri.x, rayOri.y, rayOri.z, MinRayDistance);
		rq[1] = make_float4(p.dir.x, p.dir.y, p.dir.z, MaxRayDistance);
	}
	else {   // zombie: an empty interval, the traversal kernel reports a miss without touching the BVH
		rq[0] = make_float4(0.f, 0.f, 0.f, 1.0f);
		rq[1] = make_float4(0.f, 0.f, 1.f, 0.0f);
	}
	if (bounce > 0) {   // (the G-buffer vertex draws no light sample: the shadow queue of bounce 0 is never traced)
		float4* sq = f.wf.shadowRays[bounce & 1] + 2 * size_t(nslot);
		sq[0] = sh0; sq[1] = sh1;
	}
	f.wf.pix[parity][nslot] = pix;
	storePath(f, parity, nslot, p);
}

__global__ void __launch_bounds__(GIBlock) giBeginKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s) {
	const uint32_t tilesX = (f.width + 7u) / 8u;
	const uint32_t id = blockIdx.x * GIBlock + threadIdx.x;
	const uint32_t tile = id >> 5, within = id & 31u;
	const uint32_t x = (tile % tilesX) * 8u + (within & 7u), y = f.rowBegin + (tile / tilesX) * 4u + (within >> 3);
	if (x >= f.width || y >= f.rowEnd) return;
	const uint32_t pix = uint32_t(f.index(x, y));
	const uint32_t o = ownedIndex(f, pix);
	const Primary pr = loadPrimary(f, x, y);
	if (!pr.valid) {
		// GIReservoirReset on the stored reservoir (gi_resample_temporal.glsl:44): sampleCount, weights
		float4* outResv = reinterpret_cast<float4*>(f.giThis + pix);
		float4 q2 = outResv[2];
		q2.x = __uint_as_float(0u); q2.y = 0.0f; q2.z = 0.0f;
		outResv[2] = q2;
		if (f.peerGiThisUp != nullptr && rowInUpHalo(f, y)) reinterpret_cast<float4*>(f.peerGiThisUp + peerUpIndex(f, x, y))[2] = q2;
		if (f.peerGiThisDown != nullptr && rowInDownHalo(f, y)) reinterpret_cast<float4*>(f.peerGiThisDown + peerDownIndex(f, x, y))[2] = q2;
		accumulate(f.indirectOutput, f, x, y, f3(0.0f));
		return;
	}
	GIPath p;
	p.dir = pr.ray.dir;
	p.rng = makeSeed(f.camera.seed, x, y);
	p.throughputAfter = f3(1.0f); p.lastPos = f3(0.0f); p.rcLo = f3(0.0f);
	p.bsPdf = 0.0f; p.bsType = 0; p.bounce = 0;
	p.neePending = false; p.zombie = false; p.shadowIdx = 0; p.nee = f3(0.0f);
	recordPlane(f, 1)[o] = make_float4(0.f, 0.f, 0.f, 0.f);                                   // primaryScatter, primaryPdf
	recordPlane(f, 2)[o] = make_float4(0.f, 0.f, __uint_as_float(InvalidHitIndex), 0.f);     // psIsec
	const Surface surf = primarySurface(pr);
	const Mat mat = loadMaterial(s, uint32_t(pr.matId));
	const float4 none0 = make_float4(0.f, 0.f, 0.f, 1.0f), none1 = make_float4(0.f, 0.f, 1.f, 0.0f);
	scatterAndContinue(f, p, surf, mat, pix, o, none0, none1);
}

__global__ void __launch_bounds__(GIBlock) giBounceKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, int bounce) {
	const uint32_t n = f.wf.counters[4 * bounce];
	const float sumPower = s.lightTable[0].prob;
	const int parity = bounce & 1;
	for (uint32_t slot = blockIdx.x * GIBlock + threadIdx.x; slot < n; slot += gridDim.x * GIBlock) {
		const uint32_t pix = f.wf.pix[parity][slot];
		const uint32_t o = ownedIndex(f, pix);
		GIPath p;
		loadPath(f, parity, slot, p);
		// (1) the light sample of the previous vertex
		if (p.neePending) {
			if (f.wf.occluded[(bounce - 1) & 1][slot] == 0) p.rcLo += p.nee;
			p.neePending = false;
		}
		if (p.zombie) { finishPath(f, o, p); continue; }
		// (2) the vertex the extension ray found (gi_resample_temporal.glsl:63-105)
		const RptIntersection hit = f.wf.hits[slot];
		if (hit.instanceIdx == InvalidHitIndex) { finishPath(f, o, p); continue; }
		Surface surf;
		loadSurfaceInfo(s, hit, surf);
		const Mat mat = loadMaterial(s, surf.matIndex);
		if (bounce == 1 && !surf.isLight) {
			recordPlane(f, 2)[o] = make_float4(hit.bary[0], hit.bary[1], __uint_as_float(hit.instanceIdx), __uint_as_float(hit.triangleIdx));
		}
		if (surf.isLight) {
			const float cosTheta = -dot(p.dir, surf.norm);
			if (bounce > 1 && cosTheta > 0) {
				float weight = 1.0f;
				if (!isSampleTypeDelta(p.bsType)) {
					const float dist = length(surf.pos - p.lastPos);
					const float lightPdf = luminance(surf.albedo) / sumPower * dist * dist / abs_(cosTheta);
					weight = MISWeight(p.bsPdf, lightPdf);
				}
				p.rcLo += surf.albedo * weight * p.throughputAfter;
			}
			finishPath(f, o, p);
			continue;
		}
		// (3) light sample (:107-144); the shader traces its shadow ray even when the sample cannot contribute
		float4 sh0 = make_float4(0.f, 0.f, 0.f, 1.0f), sh1 = make_float4(0.f, 0.f, 1.f, 0.0f);
		if (!isBSDFDelta(mat)) {
			const LightSample ls = sampleLight(s, surf.pos, sample4f(p.rng));
			if (ls.pdf > 1e-6f) {
				const float bsdfPdf = absDot(surf.norm, ls.wi) * RT_PI_INV;
				const float weight = MISWeight(ls.pdf, bsdfPdf);
				p.nee = ls.radiance * evalBSDF(mat, surf.albedo, surf.norm, -p.dir, ls.wi) * satDot(surf.norm, ls.wi) / ls.pdf * weight * p.throughputAfter;
				p.neePending = true;
				sh0 = make_float4(surf.pos.x, surf.pos.y, surf.pos.z, MinRayDistance);
				sh1 = make_float4(ls.wi.x, ls.wi.y, ls.wi.z, ls.dist - MinRayDistance);
			}
		}
		// (4)
		scatterAndContinue(f, p, surf, mat, pix, o, sh0, sh1);
	}
}

// temporal reservoir update (gi_resample_temporal.glsl:172-193) and the candidate final shading (:195-209)
__global__ void __launch_bounds__(GIBlock) giResolveKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s) {
	const uint32_t o = blockIdx.x * GIBlock + threadIdx.x;
	if (o >= f.ru.capacity) return;
	const uint32_t x = o % f.width, y = f.rowBegin + o / f.width;
	float4* rq = f.ru.rays + 2 * size_t(o);
	rq[0] = make_float4(0.f, 0.f, 0.f, 1.0f);   // empty interval unless a visibility ray is needed
	rq[1] = make_float4(0.f, 0.f, 1.f, 0.0f);
	const Primary p = loadPrimary(f, x, y);
	if (!p.valid) return;   // handled by giBeginKernel
	const size_t idx = f.index(x, y);
	const float4 r0 = recordPlane(f, 0)[o], r1 = recordPlane(f, 1)[o], psIsec = recordPlane(f, 2)[o];
	const float3 rcLo = f3(r0), primaryScatter = f3(r1);
	const float primaryPdf = r1.w;
	uint32_t rng = __float_as_uint(r0.w);
	const uint32_t rcPrevCoord = (y << 16) | x;
	const Mat primaryMat = loadMaterial(s, uint32_t(p.matId));
	const float3 primaryPos = p.pos, primaryWo = -p.ray.dir;
	float3 radiance = rcLo * primaryScatter / primaryPdf;

	GIResv resv;
	resv.q0 = resv.q1 = resv.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
	if ((f.camera.frameIndex & 0x80000000u) == 0) {
		const float2 motion = f.motion[idx];
		const Neighbor nb = lookupSurface(f, true, make_float2(p.uv.x + motion.x, p.uv.y + motion.y));
		if (nb.found && !(nb.matMeshId != p.matMeshId || dot(nb.norm, p.norm) < 0.9f || abs_(nb.depth - p.depth) > 5.0f)) {
			const float4* q = reinterpret_cast<const float4*>(f.giPrev + nb.pixel);
			resv.q0 = q[0]; resv.q1 = q[1]; resv.q2 = q[2];
		}
	}
	if (__float_as_uint(psIsec.z) != InvalidHitIndex) {
		float sampleWeight = luminance(radiance);
		if (isnan_(sampleWeight) || sampleWeight < 0.0f || primaryPdf < 1e-6f) sampleWeight = 0.0f;
		resv.q2.y += sampleWeight;   // GIReservoirAddSample, gi_reservoir.glsl:37-44
		resv.setSampleCount(resv.sampleCount() + 1u);
		if (sample1f(rng) * resv.q2.y < sampleWeight) {
			resv.q0 = psIsec;
			resv.q1 = make_float4(rcLo.x, rcLo.y, rcLo.z, __uint_as_float(rcPrevCoord));
		}
	}
	if (!resv.valid()) resv.reset();
	if (resv.sampleCount() > 40u) {
		resv.q2.y *= float(40) / float(resv.sampleCount());
		resv.setSampleCount(40u);
	}
	float3 visibleRadiance = f3(0.0f);
	uint32_t needRay = 0;
	if (resv.valid() && resv.sampleCount() > 0 && !isBSDFDelta(primaryMat)) {
		const uint32_t rcInst = __float_as_uint(resv.q0.z);
		if (rcInst != InvalidHitIndex) {   // see the matching note in the CPU oracle
			Surface rc;
			loadSurfaceInfo(s, rcInst, __float_as_uint(resv.q0.w), make_float2(resv.q0.x, resv.q0.y), rc);
			const float3 primaryWi = normalize(rc.pos - primaryPos);
			const float weight = resv.q2.y / float(resv.sampleCount());
			const float3 Li = f3(resv.q1) * evalBSDF(primaryMat, p.albedo, p.norm, primaryWo, primaryWi) * satDot(p.norm, primaryWi);
			if (!isBlack(Li)) {
				visibleRadiance = Li / luminance(Li) * weight;
				needRay = 1;
				// traceVisibility(primaryPos, rc.pos), ray_query.glsl:27-38
				const float3 dir = normalize(rc.pos - primaryPos);
				rq[0] = make_float4(primaryPos.x, primaryPos.y, primaryPos.z, MinRayDistance);
				rq[1] = make_float4(dir.x, dir.y, dir.z, distance(rc.pos, primaryPos) - MinRayDistance);
			}
		}
	}
	float4* outResv = reinterpret_cast<float4*>(f.giThis + idx);
	outResv[0] = resv.q0; outResv[1] = resv.q1; outResv[2] = resv.q2;
	// multi-GPU strips: boundary rows mirrored into the neighbours' halo rows (next frame's previous-frame lookups across a cut)
	if (f.peerGiThisUp != nullptr && rowInUpHalo(f, y)) {
		float4* q = reinterpret_cast<float4*>(f.peerGiThisUp + peerUpIndex(f, x, y));
		q[0] = resv.q0; q[1] = resv.q1; q[2] = resv.q2;
	}
	if (f.peerGiThisDown != nullptr && rowInDownHalo(f, y)) {
		float4* q = reinterpret_cast<float4*>(f.peerGiThisDown + peerDownIndex(f, x, y));
		q[0] = resv.q0; q[1] = resv.q1; q[2] = resv.q2;
	}
	recordPlane(f, 3)[o] = make_float4(radiance.x, radiance.y, radiance.z, __uint_as_float(needRay));
	if (needRay) recordPlane(f, 4)[o] = make_float4(visibleRadiance.x, visibleRadiance.y, visibleRadiance.z, 0.f);
}

__global__ void __launch_bounds__(GIBlock) giShadeKernel(const __grid_constant__ FrameView f) {
	const uint32_t o = blockIdx.x * GIBlock + threadIdx.x;
	if (o >= f.ru.capacity) return;
	const uint32_t x = o % f.width, y = f.rowBegin + o / f.width;
	if (f.depthNormal[f.index(x, y)].x == 0.0f) return;   // background: accumulated by giBeginKernel
	const float4 t0 = recordPlane(f, 3)[o];
	float3 radiance = f3(t0);
	if (__float_as_uint(t0.w) != 0u && f.ru.occluded[o] == 0) radiance = f3(recordPlane(f, 4)[o]);
	accumulate(f.indirectOutput, f, x, y, clampColor(radiance));
}

} // namespace

void launchGIReSTIR(const FrameView& f, const SceneView& s, cudaStream_t st, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join) {
	const bool twoStreams = side != nullptr && fork != nullptr && join != nullptr;
	static const int bounceBlocks = persistentBlocks(reinterpret_cast<const void*>(giBounceKernel), GIBlock);
	const uint32_t rows = f.rowEnd - f.rowBegin;
	const uint32_t slots = ((f.width + 7u) / 8u) * ((rows + 3u) / 4u) * 32u;
	const uint32_t n = f.ru.capacity, blocks = (n + GIBlock - 1) / GIBlock;
	cudaMemsetAsync(f.wf.counters, 0, size_t(WavefrontMaxBounces) * 4 * sizeof(uint32_t), st);
	cudaMemsetAsync(f.ru.counters, 0, 16 * sizeof(uint32_t), st);
	giBeginKernel<<<(slots + GIBlock - 1) / GIBlock, GIBlock, 0, st>>>(f, s);
	// bounce 15 only drains the paths whose last light sample is still pending
	for (int bounce = 1; bounce <= 15; bounce++) {
		uint32_t* c = f.wf.counters + 4 * bounce;
		// shadow rays of vertex b-1 next to the extension rays of bounce b, as in launchGRISPathTraceBounces
		const bool overlap = twoStreams && bounce > 1 && bounce < 15;
		if (overlap) { cudaEventRecord(fork, st); cudaStreamWaitEvent(side, fork, 0); }
		if (bounce > 1) launchTraceQueueAny(s, f.wf.shadowRays[(bounce - 1) & 1], c + 0, 0, c - 4 + 3, f.wf.occluded[(bounce - 1) & 1], overlap ? side : st);   // (slot-aligned with this bounce's queue)
		if (overlap) cudaEventRecord(join, side);
		if (bounce < 15) launchTraceQueueClosest(s, f.wf.rays[bounce & 1], c + 0, 0, c + 2, f.wf.hits, st);
		if (overlap) cudaStreamWaitEvent(st, join, 0);
		giBounceKernel<<<bounceBlocks, GIBlock, 0, st>>>(f, s, bounce);
	}
	giResolveKernel<<<blocks, GIBlock, 0, st>>>(f, s);
	launchTraceQueueAny(s, f.ru.rays, nullptr, n, f.ru.counters + 2, f.ru.occluded, st);
	giShadeKernel<<<blocks, GIBlock, 0, st>>>(f);
}

} // namespace rt
