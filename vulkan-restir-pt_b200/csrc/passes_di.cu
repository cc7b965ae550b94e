// ReSTIR DI (BASELINE.json config 2): candidate generation, temporal and spatial reuse.
//   reference src/shader/di_reservoir.glsl, di_path_gen.glsl:9-35, di_temporal.glsl:9-89, di_spatial.glsl:9-119
//   host sequence: TestReSTIR::render (src/TestReSTIR.cpp:9-36)
#include "passes.h"
#include "shading.cuh"

namespace rt {

namespace {

// DIReservoir (64 B) as four 16-byte words: q0 = isec, q1 = {Li, pad}, q2 = {jacobian, samplePdf, rng, isLightSample},
// q3 = {sampleCount, resampleWeight, contribWeight, weight}.  q0..q2 are the DIPathSample.
struct DIResv {
	float4 q0, q1, q2, q3;
	RT_DEV uint32_t instanceIdx() const { return __float_as_uint(q0.z); }
	RT_DEV uint32_t triangleIdx() const { return __float_as_uint(q0.w); }
	RT_DEV float3 Li() const { return f3(q1); }
	RT_DEV float jacobian() const { return q2.x; }
	RT_DEV float samplePdf() const { return q2.y; }
	RT_DEV uint32_t rng() const { return __float_as_uint(q2.z); }
	RT_DEV bool isLightSample() const { return __float_as_uint(q2.w) != 0u; }
	RT_DEV uint32_t sampleCount() const { return __float_as_uint(q3.x); }
	RT_DEV void setSampleCount(uint32_t c) { q3.x = __uint_as_float(c); }
	RT_DEV float& resampleWeight() { return q3.y; }
	RT_DEV float resampleWeight() const { return q3.y; }
	RT_DEV float& weight() { return q3.w; }
	RT_DEV float weight() const { return q3.w; }
	RT_DEV bool valid() const { return !isnan_(q3.y); }
	RT_DEV bool sampleValid() const { return instanceIdx() != InvalidHitIndex; }
	RT_DEV void reset() { setSampleCount(0); q3.y = 0.0f; q3.z = 0.0f; }
	RT_DEV void resetIfInvalid() { if (!valid()) reset(); }
	RT_DEV void copySample(const DIResv& o) { q0 = o.q0; q1 = o.q1; q2 = o.q2; }
};

RT_DEV DIResv zeroDI() {
	DIResv r;
	r.q0 = r.q1 = r.q2 = r.q3 = make_float4(0.f, 0.f, 0.f, 0.f);
	return r;
}
RT_DEV DIResv loadDI(const RptDIReservoir* p) {
	const float4* q = reinterpret_cast<const float4*>(p);
	DIResv r; r.q0 = q[0]; r.q1 = q[1]; r.q2 = q[2]; r.q3 = q[3];
	return r;
}
RT_DEV void storeDI(RptDIReservoir* p, const DIResv& r) {
	float4* q = reinterpret_cast<float4*>(p);
	q[0] = r.q0; q[1] = r.q1; q[2] = r.q2; q[3] = r.q3;
}

RT_DEV void diAddSample(DIResv& resv, const DIResv& sample, float w, float r) {   // di_reservoir.glsl:51-59
	resv.resampleWeight() += w;
	resv.setSampleCount(resv.sampleCount() + 1u);
	if (r * resv.resampleWeight() < w) { resv.copySample(sample); resv.weight() = w; }
}
RT_DEV void diMerge(DIResv& resv, const DIResv& rhs, float r) {   // :61-69
	resv.resampleWeight() += rhs.resampleWeight();
	resv.setSampleCount(resv.sampleCount() + rhs.sampleCount());
	if (r * resv.resampleWeight() < rhs.resampleWeight()) { resv.copySample(rhs); resv.weight() = rhs.weight(); }
}
RT_DEV void diCap(DIResv& resv, uint32_t cap) {   // :71-76
	if (resv.sampleCount() > cap) {
		resv.resampleWeight() *= float(cap) / float(resv.sampleCount());
		resv.setSampleCount(cap);
	}
}

// di_reservoir.glsl:78-188
RT_DEV void diSampleLi(const SceneView& s, const RptDISettings& st, const Surface& surf, const Mat& mat, float3 wo,
                       uint32_t rng, uint32_t& resvRng, DIResv& resv) {
	DIResv ps = zeroDI();
	ps.q2.z = __uint_as_float(rng);
	const float4 lightRand = sample4f(rng);
	const float3 scatterRand = sample3f(rng);

	if (st.sampleType != 1 && !isBSDFDelta(mat)) {
		const LightSample ls = sampleLight(s, surf.pos, lightRand);
		const bool shadowed = traceShadow(s, surf.pos, MinRayDistance, ls.wi, ls.dist - MinRayDistance);
		if (!shadowed && ls.pdf > 1e-6f) {
			const float bsdfPdf = evalPdf(mat, surf.norm, wo, ls.wi);
			float weight = MISWeight(ls.pdf, bsdfPdf);
			if (st.sampleType == 0) weight = 1.0f;
			const float3 contrib = ls.radiance * evalBSDF(mat, surf.albedo, surf.norm, wo, ls.wi) * satDot(surf.norm, ls.wi) / ls.pdf * weight;
			float sampleWeight = luminance(contrib);
			if (isnan_(sampleWeight) || sampleWeight < 0) sampleWeight = 0;
			const float3 Li = ls.radiance * weight;
			ps.q0 = make_float4(ls.bary.x, ls.bary.y, __uint_as_float(0u), __uint_as_float(ls.id));
			ps.q1 = make_float4(Li.x, Li.y, Li.z, 0.0f);
			ps.q2.x = ls.jacobian; ps.q2.y = ls.pdf; ps.q2.w = __uint_as_float(1u);
			diAddSample(resv, ps, sampleWeight, sample1f(resvRng));
		}
	}
	BSDFSample bs = emptyBSDFSample();
	ps.q2.w = __uint_as_float(0u);
	if (st.sampleType != 0 && sampleBSDF(mat, surf.albedo, surf.norm, wo, scatterRand, bs) && bs.pdf > 1e-6f) {
		const Hit h = traceClosestHit(s, surf.pos, MinRayDistance, bs.wi, MaxRayDistance);
		if (h.instanceIdx != InvalidHitIndex) {
			Surface hit;
			loadSurfaceInfo(s, h, hit);
			const float cosTheta = -dot(bs.wi, hit.norm);
			if (hit.isLight && cosTheta > 0) {
				const float dist = length(hit.pos - surf.pos);
				const float sumPower = s.lightTable[0].prob;
				const float lightPdf = luminance(hit.albedo) / sumPower * dist * dist / abs_(cosTheta);
				float weight = MISWeight(bs.pdf, lightPdf);
				if (st.sampleType == 1 || isSampleTypeDelta(bs.type)) weight = 1.0f;
				const float cosTerm = isSampleTypeDelta(bs.type) ? 1.0f : satDot(surf.norm, bs.wi);
				const float3 contrib = hit.albedo * bs.bsdf * cosTerm / bs.pdf * weight;
				const float3 Li = hit.albedo * weight;
				ps.q0 = make_float4(h.u, h.v, __uint_as_float(h.instanceIdx), __uint_as_float(h.triangleIdx));
				ps.q1 = make_float4(Li.x, Li.y, Li.z, 0.0f);
				ps.q2.x = abs_(cosTheta) / square(dist); ps.q2.y = bs.pdf; ps.q2.w = __uint_as_float(0u);
				diAddSample(resv, ps, luminance(contrib), sample1f(resvRng));
			}
		}
	}
	resv.resetIfInvalid();
	if (resv.sampleCount() > 0 && resv.sampleValid() && resv.weight() > 0) {
		const float k = resv.resampleWeight() / resv.weight();
		resv.q1.x *= k; resv.q1.y *= k; resv.q1.z *= k;
		resv.weight() = resv.resampleWeight();
	}
	else {
		resv.q0.z = __uint_as_float(InvalidHitIndex);
		resv.weight() = 0;
		resv.resampleWeight() = 0;
	}
	resv.setSampleCount(1);
}

// di_reservoir.glsl:190-224
RT_DEV void diRandomReplay(const SceneView& s, const RptDISettings& st, DIResv& dst, const Surface& dstSurf, const DIResv& src, float3 wo, uint32_t& rng) {
	const Mat dstMat = loadMaterial(s, dstSurf.matIndex);
	DIResv replay = zeroDI();
	diSampleLi(s, st, dstSurf, dstMat, wo, src.rng(), rng, replay);
	const float jacobian = 1;
	if (replay.sampleValid()) {
		Surface rs;
		loadSurfaceInfo(s, replay.instanceIdx(), replay.triangleIdx(), make_float2(replay.q0.x, replay.q0.y), rs);
		const float3 wi = normalize(rs.pos - dstSurf.pos);
		const float3 Li = replay.Li() * evalBSDF(dstMat, dstSurf.albedo, dstSurf.norm, wo, wi) * satDot(dstSurf.norm, wi) / replay.samplePdf();
		const float dstPHat = luminance(Li * jacobian);
		replay.resampleWeight() = src.resampleWeight() * dstPHat / src.weight();
		replay.setSampleCount(src.sampleCount());
	}
	else {
		replay.resampleWeight() = 0;
	}
	if (replay.valid()) diMerge(dst, replay, sample1f(rng));
}

// di_reservoir.glsl:226-286
RT_DEV void diReconnection(const SceneView& s, DIResv& dst, const Surface& dstSurf, DIResv src, float3 wo, uint32_t& rng) {
	const Mat dstMat = loadMaterial(s, dstSurf.matIndex);
	bool srcSampleValid = false;
	float dstPHat = 0, dstSamplePdf = 0, dstJacobian = 0;
	if (src.sampleValid()) {
		Surface rc;
		loadSurfaceInfo(s, src.instanceIdx(), src.triangleIdx(), make_float2(src.q0.x, src.q0.y), rc);
		const float dist = distance(rc.pos, dstSurf.pos);
		const float3 wi = normalize(rc.pos - dstSurf.pos);
		const float cosTheta = -dot(rc.norm, wi);
		dstJacobian = abs_(cosTheta) / square(dist);
		const float jacobian = dstJacobian / src.jacobian();
		if (dist > 1e-4f) {
			if (cosTheta > 0 && !isnan_(jacobian) && src.jacobian() > 0) {
				if (traceVisibility(s, dstSurf.pos, rc.pos)) {
					srcSampleValid = true;
					if (!isnan_(src.samplePdf()) && src.samplePdf() > 1e-6f) {
						const float3 Li = src.Li() * evalBSDF(dstMat, dstSurf.albedo, dstSurf.norm, wo, wi) * satDot(dstSurf.norm, wi) / src.samplePdf();
						dstPHat = luminance(Li * jacobian);
					}
					if (src.isLightSample()) {
						const float sumPower = s.lightTable[0].prob;
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
		src.q2.x = dstJacobian;
		src.q2.y = dstSamplePdf;
		if (src.q2.y < 1e-6f || isnan_(src.q2.y)) src.q2.y = 0;
		src.resampleWeight() *= dstPHat / src.weight();
		if (isnan_(src.resampleWeight())) src.resampleWeight() = 0;
	}
	else {
		src.resampleWeight() = 0;
	}
	if (src.valid()) diMerge(dst, src, sample1f(rng));
}

RT_DEV void diReuseAndMerge(const SceneView& s, const RptDISettings& st, DIResv& dst, const Surface& dstSurf, const DIResv& src, float3 wo, uint32_t& rng) {
	if (st.shiftType == 0) diReconnection(s, dst, dstSurf, src, wo, rng);
	else if (st.shiftType == 1) diRandomReplay(s, st, dst, dstSurf, src, wo, rng);
}

RT_DEV void diRecheckVisibility(const SceneView& s, DIResv& resv, float3 pos) {   // di_temporal.glsl:72-81
	if (resv.valid() && resv.sampleValid()) {
		Surface surf;
		loadSurfaceInfo(s, resv.instanceIdx(), resv.triangleIdx(), make_float2(resv.q0.x, resv.q0.y), surf);
		if (!traceVisibility(s, pos, surf.pos)) resv.resampleWeight() = 0;
	}
}

} // namespace

__global__ void __launch_bounds__(PassBlockX* PassBlockY) diPathGenKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptDISettings st) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	if (!p.valid) return;
	const uint32_t rng = makeSeed(f.camera.seed, x, y);
	uint32_t resvRng = ~rng;
	DIResv resv = zeroDI();
	diSampleLi(s, st, primarySurface(p), loadMaterial(s, uint32_t(p.matId)), -p.ray.dir, rng, resvRng, resv);
	storeDI(f.diThis + f.index(x, y), resv);
}

__global__ void __launch_bounds__(PassBlockX* PassBlockY) diTemporalKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptDISettings st) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	if (!p.valid) return;
	const size_t idx = f.index(x, y);
	const float2 motion = f.motion[idx];
	const uint32_t rng = makeSeed(f.camera.seed, x, y) ^ 1u;
	uint32_t resvRng = ~rng;
	const float3 wo = -p.ray.dir;
	DIResv resv = loadDI(f.diThis + idx);

	if (st.temporalReuse) {
		const Surface dstSurf = primarySurface(p);
		if ((f.camera.frameIndex & 0x80000000u) == 0) {
			const Neighbor nb = lookupSurface(f, true, make_float2(p.uv.x + motion.x, p.uv.y + motion.y));
			if (nb.found && !(nb.matMeshId != p.matMeshId || dot(nb.norm, p.norm) < 0.95f || distance(p.pos, nb.pos) > 0.5f)) {
				const DIResv prev = loadDI(f.diPrev + nb.pixel);
				if (prev.valid()) diReuseAndMerge(s, st, resv, dstSurf, prev, wo, resvRng);
			}
		}
		diRecheckVisibility(s, resv, p.pos);
	}
	diCap(resv, 40);
	resv.resetIfInvalid();
	storeDI(f.diTemp + idx, resv);
	// multi-GPU strips: boundary rows go straight into the neighbours' halo rows over NVLink peer memory
	if (f.peerDiUp != nullptr && rowInUpHalo(f, y)) storeDI(f.peerDiUp + peerUpIndex(f, x, y), resv);
	if (f.peerDiDown != nullptr && rowInDownHalo(f, y)) storeDI(f.peerDiDown + peerDownIndex(f, x, y), resv);
}

__global__ void __launch_bounds__(PassBlockX* PassBlockY) diSpatialKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s, const RptDISettings st) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	float3 radiance = f3(0.0f);
	if (p.valid) {
		const size_t idx = f.index(x, y);
		uint32_t rng = makeSeed(f.camera.seed, x, y) ^ 2u;
		const float3 wo = -p.ray.dir;
		const Mat mat = loadMaterial(s, uint32_t(p.matId));
		DIResv resv = loadDI(f.diTemp + idx);
		const float texelX = 1.0f / float(f.width), texelY = 1.0f / float(f.height);

		if (st.spatialReuse) {
			const Surface dstSurf = primarySurface(p);
			for (uint32_t i = 0; i < 10; i++) {
				const float2 d = toConcentricDisk(sample2f(rng));
				const float2 nuv = make_float2(p.uv.x + d.x * 20.0f * texelX, p.uv.y + d.y * 20.0f * texelY);
				const Neighbor nb = lookupSurface(f, false, nuv);
				if (nb.found && !(dot(nb.norm, p.norm) < 0.9f || distance(p.pos, nb.pos) > 0.4f)) {
					const DIResv nr = loadDI(f.diTemp + nb.pixel);
					if (nr.valid()) diReuseAndMerge(s, st, resv, dstSurf, nr, wo, rng);
				}
			}
			diRecheckVisibility(s, resv, p.pos);
		}
		diCap(resv, 40);
		resv.resetIfInvalid();
		storeDI(f.diThis + idx, resv);
		// multi-GPU strips: the final reservoirs of the boundary rows are mirrored into the neighbours' halo rows (next frame's
		// previous-frame lookups across a cut, as in passes_gris.cu spatialStore)
		if (f.peerDiThisUp != nullptr && rowInUpHalo(f, y)) storeDI(f.peerDiThisUp + peerUpIndex(f, x, y), resv);
		if (f.peerDiThisDown != nullptr && rowInDownHalo(f, y)) storeDI(f.peerDiThisDown + peerDownIndex(f, x, y), resv);

		if (resv.valid() && resv.sampleValid()) {
			Surface surf;
			loadSurfaceInfo(s, resv.instanceIdx(), resv.triangleIdx(), make_float2(resv.q0.x, resv.q0.y), surf);
			const float3 wi = normalize(surf.pos - p.pos);
			if (resv.sampleCount() > 0) {
				const float3 Li = resv.Li() * evalBSDF(mat, p.albedo, p.norm, wo, wi) * satDot(p.norm, wi) / resv.samplePdf();
				if (!isBlack(Li)) radiance = Li / luminance(Li) * resv.resampleWeight() / float(resv.sampleCount());
			}
		}
		radiance = clampColor(radiance);
	}
	accumulate(f.directOutput, f, x, y, radiance);
}

void launchDIPathGen(const FrameView& f, const SceneView& s, const RptDISettings& p, cudaStream_t st) {
	diPathGenKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s, p);
}
void launchDITemporal(const FrameView& f, const SceneView& s, const RptDISettings& p, cudaStream_t st) {
	diTemporalKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s, p);
}
void launchDISpatial(const FrameView& f, const SceneView& s, const RptDISettings& p, cudaStream_t st) {
	diSpatialKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s, p);
}

} // namespace rt
