// ORACLE — TEST INFRASTRUCTURE ONLY.  Scene storage, ray/triangle intersection and texture fetch on the CPU:
// stands in for what the reference gets from the Vulkan driver (acceleration structures + rayQueryEXT,
// reference src/shader/ray_query.glsl:6-70; samplers, zvk/core/Memory.cpp:75-92).  These are the DRIVER's parts: there is nothing of the reference to pin them against.
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
	std::vector<RptObjectInstance> prevInstances;   // non-empty while an instance update is "in motion" (rpt_scene_end_motion)
	std::vector<RptTriangleLight> lights;
	std::vector<RptLightSampleTableElement> lightTable;
	std::vector<Texture> textures;
	float srgbToLinear[256];

	// Single-level scenes: every instance flattened to world space (the reference never shares geometry between instances,
	// src/Resource.cpp:183-184), one set.  Two-level scenes (RPT_SCENE_TWO_LEVEL — the reference's own BLAS / TLAS arrangement,
	// src/Scene.cpp:448-547): the light triangles in world space (record 0) and one set per unique mesh in OBJECT space; a ray
	// is taken into the object space of every instance (direction not renormalised, so t is shared) and the triangle test runs
	// there.  Hit definition either way: the triangle test below, minimum t, ties to the lower flattened index.
	//
	// A set = triangles + a BVH2 (median-of-centroid splits with SAH binning), which is only an accelerator for the
	// brute-force definition — both give the same answer by construction (conservative boxes, same tie rule).
	struct Node { float lo[3], hi[3]; uint32_t left, count; };   // count > 0: leaf over order[left .. left+count)
	struct TriSet {
		std::vector<WorldTri> tris;
		std::vector<Node> nodes;
		std::vector<uint32_t> order;
		void buildTree();
		void buildNode(uint32_t nodeIdx, uint32_t begin, uint32_t end, std::vector<vec3>& cen, int depth);
		// calls fn(index in tris) for every triangle whose box the ray interval [tmin, tfar()] may enter (all of them when brute)
		template <typename TFar, typename Fn>
		void candidates(vec3 o, vec3 d, float tmin, TFar tfar, bool brute, Fn fn) const;
	};
	struct InstanceRecord {      // record 0 = the light triangles, k + 1 = object instance k
		float r[3][4];           // world -> object, rows
		uint32_t set;            // index into sets
		uint32_t customIndex;    // Intersection.instanceIdx
		uint32_t flatBase;       // flattened index of the instance's triangle 0
	};
	std::vector<TriSet> sets;                // single-level: one
	std::vector<InstanceRecord> records;     // two-level only
	bool twoLevel = false;
	uint32_t numFlatTris = 0;

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
	// fn(triangle, flattened index, custom instance index, object-space o, d) for every candidate of the ray
	template <typename TFar, typename Fn>
	void forCandidates(vec3 o, vec3 d, float tmin, TFar tfar, Fn fn) const;
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
