// G-buffer, acceleration-structure visualisation, post-process and the raw ray-query kernels.
//   G-buffer: reference GBuffer.vert:19-30 + GBuffer.frag:21-54 (rasterised there; here one primary ray per
//             pixel centre against object instances only — the light mesh is never rasterised,
//             src/GBufferPass.cpp:50-54)
//   visualize: as_visualize.comp:14-27;  post-process: post_proc.frag:16-42
#include <cuda_fp16.h>
#include "passes.h"
#include "shading.cuh"

namespace rt {

__global__ void __launch_bounds__(PassBlockX* PassBlockY) gbufferKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	uint32_t y = f.storeBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width) return;
	// two extra rows behind the stored ones: film rows 0 and H-1 of a strip that does not own them (depth/normal only,
	// for REPEAT-wrapped bilinear taps; see depthNormalRow)
	bool wrapRowOnly = false;
	size_t i;
	if (y >= f.storeEnd) {
		const uint32_t extra = y - f.storeEnd;
		if (extra > 1u) return;
		y = extra == 0u ? 0u : f.height - 1u;
		if (y >= f.storeBegin && y < f.storeEnd) return;   // already stored
		wrapRowOnly = true;
		i = size_t(f.storeEnd - f.storeBegin + extra) * f.width + x;
	}
	else {
		i = f.index(x, y);
	}
	const RptCamera& cam = f.camera;
	const float2 uv = make_float2((float(x) + 0.5f) / float(f.width), (float(y) + 0.5f) / float(f.height));
	const Ray ray = pinholeCameraSampleRay(cam, make_float2(uv.x, 1.0f - uv.y));
	const Hit h = traceRay<TraceClosestNoLights>(s, ray.ori, cam.nearZ, ray.dir, MaxRayDistance);
	if (!wrapRowOnly) {
		RptIntersection pi;
		pi.bary[0] = h.u; pi.bary[1] = h.v; pi.instanceIdx = h.instanceIdx; pi.triangleIdx = h.triangleIdx;
		f.primaryIsec[i] = pi;
	}
	if (h.instanceIdx == InvalidHitIndex) {
		f.depthNormal[i] = make_float4(0.f, 0.f, 0.f, 0.f);
		if (wrapRowOnly) return;
		f.albedoMatId[i] = make_uint2(0u, 0u);
		f.motion[i] = make_float2(0.f, 0.f);
		return;
	}
	const uint32_t instIdx = h.instanceIdx - 1;
	const RptObjectInstance* inst = s.instances + instIdx;
	const uint32_t indexOffset = inst->indexOffset;
	const uint32_t matIndex = uint32_t(__ldg(s.materialIndices + (indexOffset / 3 + h.triangleIdx)));
	const uint32_t* ip = s.indices + (indexOffset + h.triangleIdx * 3);
	const float4* v0 = reinterpret_cast<const float4*>(s.vertices + __ldg(ip));
	const float4* v1 = reinterpret_cast<const float4*>(s.vertices + __ldg(ip + 1));
	const float4* v2 = reinterpret_cast<const float4*>(s.vertices + __ldg(ip + 2));
	const float4 p0 = __ldg(v0), q0 = __ldg(v0 + 1), p1 = __ldg(v1), q1 = __ldg(v1 + 1), p2 = __ldg(v2), q2 = __ldg(v2 + 1);
	const float3 bary = make_float3(1.0f - h.u - h.v, h.u, h.v);
	const float3 P = xformPoint(inst->transform, interp(f3(p0), f3(p1), f3(p2), bary));
	const float3 n0 = normalize(xformDir(inst->transformInvT, f3(q0)));
	const float3 n1 = normalize(xformDir(inst->transformInvT, f3(q1)));
	const float3 n2 = normalize(xformDir(inst->transformInvT, f3(q2)));
	const float3 N = normalize(interp(n0, n1, n2, bary));
	if (wrapRowOnly) {
		f.depthNormal[i] = make_float4(length(f3(cam.pos) - P), N.x, N.y, N.z);
		return;
	}
	const float uvx = interp(p0.w, p1.w, p2.w, bary), uvy = interp(q0.w, q1.w, q2.w, bary);
	const Mat mat = loadMaterial(s, matIndex);
	const float3 albedo = (mat.textureIdx == InvalidResourceIdx) ? mat.baseColor : sampleTexture(s, mat.textureIdx, uvx, uvy);

	// GBuffer.frag:40-44 reprojects the CURRENT world position (the reference's scenes are static).  While an instance update is
	// in motion the surface point is followed back through its instance's previous placement, so the motion vector carries
	// the object's movement as well as the camera's.
	const float3 Plast = s.prevInstances == nullptr ? P : xformPoint(s.prevInstances[instIdx].transform, interp(f3(p0), f3(p1), f3(p2), bary));
	const float4 last = xformPoint4(cam.lastProjView, Plast);
	const float2 lastCoord = make_float2((last.x / last.w) * 0.5f + 0.5f, (last.y / last.w) * 0.5f + 0.5f);
	const float2 motion = make_float2(lastCoord.x - uv.x, lastCoord.y - uv.y);

	f.depthNormal[i] = make_float4(length(f3(cam.pos) - P), N.x, N.y, N.z);
	f.albedoMatId[i] = make_uint2(packAlbedo(albedo), (matIndex << 16) | instIdx);
	// RG16F render target: round to nearest even through fp16
	f.motion[i] = make_float2(__half2float(__float2half_rn(motion.x)), __half2float(__float2half_rn(motion.y)));
}

__global__ void __launch_bounds__(PassBlockX* PassBlockY) visualizeASKernel(const __grid_constant__ FrameView f, const __grid_constant__ SceneView s) {
	const uint32_t x = blockIdx.x * PassBlockX + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * PassBlockY + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const float2 uv = make_float2((float(x) + 0.5f) / float(f.width), (float(y) + 0.5f) / float(f.height));
	const Ray ray = pinholeCameraSampleRay(f.camera, make_float2(uv.x, 1.0f - uv.y));
	uint32_t count = 0;
	traceRay<TraceCount>(s, ray.ori, MinRayDistance, ray.dir, MaxRayDistance, &count);
	const float level = float(count) / 100.0f;
	f.directOutput[f.index(x, y)] = make_float4(level, level, level, 1.0f);
}

RT_DEV float filmic1(float c) { return (c * (c * 0.22f + 0.03f) + 0.002f) / (c * (c * 0.22f + 0.3f) + 0.06f) - 1.0f / 30.0f; }

// writes the owned rows only, densely: out[(y - rowBegin) * width + x]
__global__ void __launch_bounds__(256) postProcessKernel(const __grid_constant__ FrameView f, RptPostSettings p, uchar4* __restrict__ out) {
	const uint32_t x = blockIdx.x * 32 + threadIdx.x;
	const uint32_t y = f.rowBegin + blockIdx.y * 8 + threadIdx.y;
	if (x >= f.width || y >= f.rowEnd) return;
	const size_t i = f.index(x, y);
	float c[3] = { 0.f, 0.f, 0.f };
	if (p.noDirect == 0) { const float4 d = f.directOutput[i]; c[0] += d.x; c[1] += d.y; c[2] += d.z; }
	if (p.noIndirect == 0) { const float4 d = f.indirectOutput[i]; c[0] += d.x; c[1] += d.y; c[2] += d.z; }
	uint32_t q[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		float v = c[k];
		if (p.toneMapping == 1) v = filmic1(v * 1.6f) / filmic1(11.2f);
		else if (p.toneMapping == 2) v = (v * (v * 2.51f + 0.03f)) / (v * (v * 2.43f + 0.59f) + 0.14f);
		if (p.correctGamma != 0) v = powf(v, 1.0f / 2.2f);
		v = v != v ? 0.0f : clamp_(v, 0.0f, 1.0f);
		q[k] = uint32_t(floorf(v * 255.0f + 0.5f));
	}
	const uchar4 px = make_uchar4(uint8_t(q[0]), uint8_t(q[1]), uint8_t(q[2]), 255);
	out[size_t(y - f.rowBegin) * f.width + x] = px;
	// multi-GPU strips with a connected gather: the row also goes straight into the full-film image on the root strip's GPU
	// (NVLink peer store, coalesced 128 B per warp) — the final gather of SURVEY.md §8(e) fused into the pass that makes the pixels
	if (f.gatherImage != nullptr) f.gatherImage[size_t(y) * f.width + x] = px;
}

// raw ray queries for the parity tests / traversal microbenchmarks: rays[2i] = {o, tmin}, rays[2i+1] = {d, tmax}
__global__ void __launch_bounds__(128) traceRaysKernel(const __grid_constant__ SceneView s, const float4* __restrict__ rays, uint32_t n,
                                                        RptIntersection* __restrict__ out, uint8_t* __restrict__ occluded) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4 a = rays[2 * size_t(i)], b = rays[2 * size_t(i) + 1];
	if (out) {
		const Hit h = traceRay<TraceClosest>(s, f3(a), a.w, f3(b), b.w);
		RptIntersection r;
		r.bary[0] = h.u; r.bary[1] = h.v; r.instanceIdx = h.instanceIdx; r.triangleIdx = h.triangleIdx;
		out[i] = r;
	}
	if (occluded) {
		occluded[i] = traceShadow(s, f3(a), a.w, f3(b), b.w) ? 1 : 0;
	}
}

void launchGBuffer(const FrameView& f, const SceneView& s, cudaStream_t st) {
	const uint32_t extraRows = f.striped ? 2u : 0u;
	gbufferKernel<<<passGrid(f.width, f.storeEnd - f.storeBegin + extraRows), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s);
}
void launchVisualizeAS(const FrameView& f, const SceneView& s, cudaStream_t st) {
	visualizeASKernel<<<passGrid(f.width, f.rowEnd - f.rowBegin), dim3(PassBlockX, PassBlockY), 0, st>>>(f, s);
}
void launchPostProcess(const FrameView& f, const RptPostSettings& p, uchar4* rgba8, cudaStream_t st) {
	dim3 grid((f.width + 31) / 32, (f.rowEnd - f.rowBegin + 7) / 8);
	postProcessKernel<<<grid, dim3(32, 8), 0, st>>>(f, p, rgba8);
}
void launchTraceRays(const SceneView& s, const float4* rays, uint32_t n, RptIntersection* out, uint8_t* occluded, cudaStream_t st) {
	if (n == 0) return;
	traceRaysKernel<<<(n + 127) / 128, 128, 0, st>>>(s, rays, n, out, occluded);
}

} // namespace rt
