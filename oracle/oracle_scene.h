// ORACLE — TEST INFRASTRUCTURE ONLY.  Scene storage, ray/triangle intersection and texture fetch on the CPU:
// stands in for what the reference gets from the Vulkan driver (acceleration structures + rayQueryEXT,
// reference src/shader/ray_query.glsl:6-70; samplers, zvk/core/Memory.cpp:75-92).  PARITY UNPINNED.
#pragma once
#include <atomic>
#include <cstdint>
#include <vector>
#include "oracle_math.h"
#include "../include/restirpt.h"

namespace orc {

const uint32_t InvalidHitIndex = 0xffffffffu;
const uint32_t SpecialHitIndex = 0xfffffffeu;
const uint32_t InvalidResourceIdx = 0xffffffffu;
const float MinRayDistance = 1e-4f;
const float MaxRayDistance = 1e7f;

struct Intersection {
	vec2 bary;
	uint32_t instanceIdx;
	uint32_t triangleIdx;
};

// world-space triangle in flattened order: light triangles (instance 0) first, then object instance 1, 2, ...
struct WorldTri {
	vec3 v0, e1, e2;
	uint32_t instanceIdx, triangleIdx;
};

struct Texture {
	uint32_t width, height, filter;
	std::vector<uint8_t> rgba8;
};

struct Counters {
	std::atomic<uint64_t> closestRays{ 0 }, shadowRays{ 0 };
};

struct Scene {
	std::vector<RptMeshVertex> vertices;
	std::vector<uint32_t> indices;
	std::vector<RptMaterial> materials;
	std::vector<int32_t> materialIndices;
	std::vector<RptObjectInstance> instances;
	std::vector<RptTriangleLight> lights;
	std::vector<RptLightSampleTableElement> lightTable;
	std::vector<Texture> textures;
	float srgbToLinear[256];

	std::vector<WorldTri> tris;
	uint32_t firstObjectTri = 0;   // tris[0 .. firstObjectTri) are light triangles

	// BVH2 over tris (median-of-centroid splits with SAH binning); only an accelerator for the brute-force
	// definition below — both give the same answer by construction (conservative boxes, same tie rule)
	struct Node { float lo[3], hi[3]; uint32_t left, count; };   // count > 0: leaf over order[left .. left+count)
	std::vector<Node> nodes;
	std::vector<uint32_t> order;

	mutable Counters counters;
	bool bruteForce = false;

	void build(const RptSceneDesc& d);

	// nearest hit with tmin < t < tmax; ties on t resolved towards the lower flattened triangle index.
	// skipLights: ignore instance 0 (the rasterised G-buffer never draws the light mesh)
	Intersection traceClosestHit(vec3 o, float tmin, vec3 d, float tmax, bool skipLights = false) const;
	bool traceShadow(vec3 o, float tmin, vec3 d, float tmax) const;
	bool traceVisibility(vec3 from, vec3 to) const {   // ray_query.glsl:27-38
		return !traceShadow(from, MinRayDistance, normalize(to - from), distance(to, from) - MinRayDistance);
	}
	// number of candidate triangle tests along a primary ray (as_visualize.comp / debugVisualizeAS)
	uint32_t countCandidates(vec3 o, vec3 d) const;

	vec3 sampleTexture(uint32_t texIdx, float u, float v) const;

private:
	void buildNode(uint32_t nodeIdx, uint32_t begin, uint32_t end, std::vector<vec3>& cen, int depth);
};

// Möller–Trumbore on (v0, e1, e2) with the operation order fixed by the numeric contract (oracle_math.h).
// The barycentric tests carry a small tolerance so that rounding cannot open cracks along shared edges
// (the hardware intersector the reference runs on is watertight; plain Möller–Trumbore is not).
const float BaryEps = 1e-4f;
inline bool intersectTri(const WorldTri& t, vec3 o, vec3 d, float tmin, float tmax, float& outT, float& outU, float& outV) {
	vec3 p = cross(d, t.e2);
	float det = dot(t.e1, p);
	float inv = 1.0f / det;
	vec3 s = o - t.v0;
	float u = dot(s, p) * inv;
	vec3 q = cross(s, t.e1);
	float v = dot(d, q) * inv;
	float tt = dot(t.e2, q) * inv;
	if (u >= -BaryEps && v >= -BaryEps && (u + v) <= 1.0f + BaryEps && tt > tmin && tt < tmax) {
		outT = tt; outU = u; outV = v;
		return true;
	}
	return false;
}

} // namespace orc
