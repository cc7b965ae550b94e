// Naive direct illumination and naive path tracing (BASELINE.json config 1).
//   di_naive.comp -> directIllumination2 (reference src/shader/di_naive.glsl:79-166)
//   gi_naive.comp -> indirectIllumination (reference src/shader/gi_naive.glsl:28-148)
#include "passes.h"
#include "shading.cuh"

namespace rt {

namespace {

struct StreamRIS {   // di_naive.glsl:54-77
	float3 Li;
	float weight, sumWeight;
	RT_DEV void add(float3 L, float w, float r) {
		sumWeight += w;
		if (r * sumWeight < w) { weight = w; Li = L; }
	}
};

} // namespace

__global__ void __launch_bounds__(PassBlockX* PassBlockY) diNaiveKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	float3 radiance = f3(0.0f);
	if (p.valid) {
		uint32_t rng = makeSeed(f.camera.seed, x, y);
		const float3 wo = -p.ray.dir;
		const Mat mat = loadMaterial(s, uint32_t(p.matId));
		StreamRIS resv;
		resv.Li = f3(0.0f); resv.weight = 0.0f; resv.sumWeight = 0.0f;

		if (!isBSDFDelta(mat)) {
			const LightSample ls = sampleLight(s, p.pos, sample4f(rng));
			const bool shadowed = traceShadow(s, p.pos, MinRayDistance, ls.wi, ls.dist - 1e-4f);
			if (!shadowed && ls.pdf > 1e-6f) {
				const float bsdfPdf = evalPdf(mat, p.norm, wo, ls.wi);
				const float weight = MISWeight(ls.pdf, bsdfPdf);
				const float3 contrib = ls.radiance * evalBSDF(mat, p.albedo, p.norm, wo, ls.wi) * satDot(p.norm, ls.wi) / ls.pdf * weight;
				resv.add(contrib, 100.0f, sample1f(rng));
			}
		}
		BSDFSample bs = emptyBSDFSample();
		const float3 r3 = sample3f(rng);
		if (sampleBSDF(mat, p.albedo, p.norm, wo, r3, bs) && bs.pdf > 1e-6f) {
			const Hit h = traceClosestHit(s, p.pos, MinRayDistance, bs.wi, MaxRayDistance);
			if (h.instanceIdx != InvalidHitIndex) {
				Surface surf;
				loadSurfaceInfo(s, h, surf);
				const float cosTheta = -dot(bs.wi, surf.norm);
				if (surf.isLight && cosTheta > 0) {
					const float dist = length(surf.pos - p.pos);
					const float sumPower = s.lightTable[0].prob;
					const float lightPdf = luminance(surf.albedo) / sumPower * dist * dist / abs_(cosTheta);
					const float weight = isSampleTypeDelta(bs.type) ? 1.0f : MISWeight(bs.pdf, lightPdf);
					const float cosTerm = isSampleTypeDelta(bs.type) ? 1.0f : satDot(p.norm, bs.wi);
					const float3 contrib = surf.albedo * bs.bsdf * cosTerm / bs.pdf * weight;
					resv.add(contrib, 1.0f, sample1f(rng));
				}
			}
		}
		if (resv.weight > 0 && resv.sumWeight > 0) radiance = resv.Li * resv.sumWeight / resv.weight;
		radiance = clampColor(radiance);
	}
	accumulate(f.directOutput, f, x, y, radiance);
}

// RT-pipeline mode of the reference (di_naive.rgen -> directIllumination, src/shader/di_naive.glsl:9-52): one light
// sample, its MIS weight computed and then forced to 1 (:47), no BSDF sample.  The only pass whose .rgen and .comp entry
// points run different estimators; the other .rgen files call the same functions as their .comp twins.
__global__ void __launch_bounds__(PassBlockX* PassBlockY) diNaiveRTKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	float3 radiance = f3(0.0f);
	if (p.valid) {
		uint32_t rng = makeSeed(f.camera.seed, x, y);
		const float3 wo = -p.ray.dir;
		const Mat mat = loadMaterial(s, uint32_t(p.matId));
		if (!isBSDFDelta(mat)) {
			const LightSample ls = sampleLight(s, p.pos, sample4f(rng));
			const bool shadowed = traceShadow(s, p.pos, MinRayDistance, ls.wi, ls.dist - 1e-4f);
			if (!shadowed && ls.pdf > 1e-6f) {
				const float weight = 1.0f;
				radiance += ls.radiance * evalBSDF(mat, p.albedo, p.norm, wo, ls.wi) * satDot(p.norm, ls.wi) / ls.pdf * weight;
			}
		}
		radiance = clampColor(radiance);
	}
	accumulate(f.directOutput, f, x, y, radiance);
}

__global__ void __launch_bounds__(PassBlockX* PassBlockY) giNaiveKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const Primary p = loadPrimary(f, x, y);
	float3 radiance = f3(0.0f);
	if (p.valid) {
		Ray ray = p.ray;
		uint32_t rng = makeSeed(f.camera.seed, x, y);
		float3 throughput = f3(1.0f), lastPos = f3(0.0f);
		float3 wo = -ray.dir;
		Surface surf = primarySurface(p);
		Mat mat = loadMaterial(s, uint32_t(p.matId));
		BSDFSample bs = emptyBSDFSample();
		const float sumPower = s.lightTable[0].prob;

		for (int bounce = 0; bounce < 15; bounce++) {
			if (bounce > 0) {
				const Hit h = traceClosestHit(s, ray.ori, MinRayDistance, ray.dir, MaxRayDistance);
				if (h.instanceIdx == InvalidHitIndex) break;
				loadSurfaceInfo(s, h, surf);
				mat = loadMaterial(s, surf.matIndex);
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
					radiance += surf.albedo * weight * throughput;
				}
				break;
			}
			if (bounce > 0 && !isBSDFDelta(mat)) {
				const LightSample ls = sampleLight(s, surf.pos, sample4f(rng));
				const bool shadowed = traceShadow(s, surf.pos, MinRayDistance, ls.wi, ls.dist - MinRayDistance);
				if (!shadowed && ls.pdf > 1e-6f) {
					const float weight = 1.0f;   // MIS weight forced to 1 (gi_naive.glsl:116-118)
					radiance += ls.radiance * evalBSDF(mat, surf.albedo, surf.norm, wo, ls.wi) * satDot(surf.norm, ls.wi) / ls.pdf * weight * throughput;
				}
			}
			if (bounce > 4) {
				const float pdfTerminate = max_(1.0f - luminance(throughput), 0.0f);
				if (sample1f(rng) < pdfTerminate) break;
				throughput /= (1.0f - pdfTerminate);
			}
			const float3 r3 = sample3f(rng);
			if (!sampleBSDF(mat, surf.albedo, surf.norm, wo, r3, bs) || bs.pdf < 1e-6f) break;
			const float cosTheta = isSampleTypeDelta(bs.type) ? 1.0f : absDot(surf.norm, bs.wi);
			throughput *= bs.bsdf * cosTheta / bs.pdf;
			lastPos = surf.pos;
			wo = -bs.wi;
			ray.dir = bs.wi;
			ray.ori = surf.pos + ray.dir * 1e-4f;
		}
		radiance = clampColor(radiance);
	}
	accumulate(f.indirectOutput, f, x, y, radiance);
}

void launchDINaiveRT(const FrameView& f, const SceneView& s, cudaStream_t st) {
	diNaiveRTKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s);
}
void launchDINaive(const FrameView& f, const SceneView& s, cudaStream_t st) {
	diNaiveKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s);
}
void launchGINaive(const FrameView& f, const SceneView& s, cudaStream_t st) {
	giNaiveKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s);
}

} // namespace rt
