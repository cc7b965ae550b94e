// ReSTIR GI in one kernel (BASELINE.json config 5).
//   reference src/shader/gi_resample_temporal.glsl:9-212 (+ .comp), gi_reservoir.glsl:8-50
#include "passes.h"
#include "shading.cuh"

namespace rt {

namespace {

// GIReservoir (48 B): q0 = rcIsec, q1 = {rcLo, rcPrevCoord}, q2 = {sampleCount, resampleWeight, contribWeight, pad}
struct GIResv {
	float4 q0, q1, q2;
	RT_DEV uint32_t sampleCount() const { return __float_as_uint(q2.x); }
	RT_DEV void setSampleCount(uint32_t c) { q2.x = __uint_as_float(c); }
	RT_DEV bool valid() const { return !isnan_(q2.y) && q2.y >= 0; }
	RT_DEV void reset() { setSampleCount(0); q2.y = 0.0f; q2.z = 0.0f; }
};

} // namespace

__global__ void __launch_bounds__(PassBlockX* PassBlockY) giReSTIRKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const size_t idx = f.index(x, y);
	float4* outResv = reinterpret_cast<float4*>(f.giThis + idx);
	const Primary p = loadPrimary(f, x, y);
	if (!p.valid) {
		// GIReservoirReset on the stored reservoir (gi_resample_temporal.glsl:44): sampleCount, weights
		float4 q2 = outResv[2];
		q2.x = __uint_as_float(0u); q2.y = 0.0f; q2.z = 0.0f;
		outResv[2] = q2;
		accumulate(f.indirectOutput, f, x, y, f3(0.0f));
		return;
	}
	const float2 motion = f.motion[idx];
	Ray ray = p.ray;
	uint32_t rng = makeSeed(f.camera.seed, x, y);
	float3 throughputAfter = f3(1.0f), lastPos = f3(0.0f);
	float3 wo = -ray.dir;
	Surface surf = primarySurface(p);

	float4 psIsec = make_float4(0.f, 0.f, __uint_as_float(InvalidHitIndex), 0.f);
	float3 rcLo = f3(0.0f);
	const uint32_t rcPrevCoord = (y << 16) | x;

	const float3 primaryPos = surf.pos, primaryWo = -ray.dir;
	float3 primaryScatter = f3(0.0f);
	float primaryPdf = 0.0f;
	const Mat primaryMat = loadMaterial(s, uint32_t(p.matId));
	Mat mat = primaryMat;
	BSDFSample bs = emptyBSDFSample();
	const float sumPower = s.lightTable[0].prob;

	for (int bounce = 0; bounce < 15; bounce++) {
		if (bounce > 0) {
			const Hit h = traceClosestHit(s, ray.ori, MinRayDistance, ray.dir, MaxRayDistance);
			if (h.instanceIdx == InvalidHitIndex) break;
			loadSurfaceInfo(s, h, surf);
			mat = loadMaterial(s, surf.matIndex);
			if (bounce == 1 && !surf.isLight) psIsec = make_float4(h.u, h.v, __uint_as_float(h.instanceIdx), __uint_as_float(h.triangleIdx));
		}
		if (surf.isLight) {
			const float cosTheta = -dot(ray.dir, surf.norm);
			if (bounce > 1 && cosTheta > 0) {
				float weight = 1.0f;
				if (!isSampleTypeDelta(bs.type)) {
					const float dist = length(surf.pos - lastPos);
					const float lightPdf = luminance(surf.albedo) / sumPower * dist * dist / abs_(cosTheta);
					weight = MISWeight(bs.pdf, lightPdf);
				}
				rcLo += surf.albedo * weight * throughputAfter;
			}
			break;
		}
		if (bounce > 0 && !isBSDFDelta(mat)) {
			const LightSample ls = sampleLight(s, surf.pos, sample4f(rng));
			const bool shadowed = traceShadow(s, surf.pos, MinRayDistance, ls.wi, ls.dist - MinRayDistance);
			if (!shadowed && ls.pdf > 1e-6f) {
				const float bsdfPdf = absDot(surf.norm, ls.wi) * RT_PI_INV;
				const float weight = MISWeight(ls.pdf, bsdfPdf);
				rcLo += ls.radiance * evalBSDF(mat, surf.albedo, surf.norm, wo, ls.wi) * satDot(surf.norm, ls.wi) / ls.pdf * weight * throughputAfter;
			}
		}
		if (bounce > 4) {
			const float pdfTerminate = max_(1.0f - luminance(throughputAfter), 0.0f);
			if (sample1f(rng) < pdfTerminate) break;
			throughputAfter /= (1.0f - pdfTerminate);
		}
		const float3 r3 = sample3f(rng);
		if (!sampleBSDF(mat, surf.albedo, surf.norm, wo, r3, bs) || bs.pdf < 1e-6f) break;
		const float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
		const float3 scatterTerms = bs.bsdf * cosTheta / bs.pdf;
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
	float3 radiance = rcLo * primaryScatter / primaryPdf;

	GIResv resv;
	resv.q0 = resv.q1 = resv.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
	if ((f.camera.frameIndex & 0x80000000u) == 0) {
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
	if (resv.valid() && resv.sampleCount() > 0 && !isBSDFDelta(primaryMat)) {
		const uint32_t rcInst = __float_as_uint(resv.q0.z);
		if (rcInst != InvalidHitIndex) {   // see the matching note in the CPU oracle
			Surface rc;
			loadSurfaceInfo(s, rcInst, __float_as_uint(resv.q0.w), make_float2(resv.q0.x, resv.q0.y), rc);
			const float3 primaryWi = normalize(rc.pos - primaryPos);
			const float weight = resv.q2.y / float(resv.sampleCount());
			const float3 Li = f3(resv.q1) * evalBSDF(primaryMat, p.albedo, p.norm, primaryWo, primaryWi) * satDot(p.norm, primaryWi);
			if (!isBlack(Li) && traceVisibility(s, primaryPos, rc.pos)) radiance = Li / luminance(Li) * weight;
		}
	}
	outResv[0] = resv.q0; outResv[1] = resv.q1; outResv[2] = resv.q2;
	accumulate(f.indirectOutput, f, x, y, clampColor(radiance));
}

void launchGIReSTIR(const FrameView& f, const SceneView& s, cudaStream_t st) {
	giReSTIRKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s);
}

} // namespace rt
