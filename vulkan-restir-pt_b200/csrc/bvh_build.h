// Host-callable interface of the GPU BVH builder (bvh_build.cu).
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "rt_types.cuh"

namespace rt {

struct BuildInputs {                     // all pointers are device pointers
	const RptMeshVertex* vertices;
	const uint32_t* indices;
	const RptObjectInstance* instances;
	const RptTriangleLight* lights;
	const uint32_t* triOffsets;          // numInstances + 1 entries; [0] == numLights (flattened index of each instance's first triangle)
	uint32_t numInstances;
	uint32_t numLights;
	uint32_t numTris;                    // lights + all instance triangles
};

struct BuildOutputs {
	WideNode* nodes = nullptr;           // cudaMalloc'd, owned by the caller
	TriRecord* tris = nullptr;           // leaf-ordered triangles
	uint32_t numNodes = 0;
	uint32_t numTris = 0;
	uint32_t depth = 0;                  // levels of the wide tree
	float3 rootLo{}, rootHi{};           // bounds of everything (padded as the leaves are)
	float buildMs = 0.0f;
};

// a wide tree deeper than this cannot be traversed with the fixed stack (bvh_traverse.cuh: one deferred group per level)
constexpr int MaxWideTreeDepth = 46;

cudaError_t buildBvh(const BuildInputs& in, cudaStream_t stream, BuildOutputs* out);

// ---- two-level structure: one BLAS per unique mesh in OBJECT space + a TLAS over the instances --------------------------
// (the reference's own arrangement: a BLAS per model, a light BLAS with custom index 0, a TLAS of instances with 3x4
// transforms — src/Scene.cpp:448-547, zvk/core/AccelerationStructure.cpp:46-136)
struct MeshRange { uint32_t indexOffset, indexCount; };

struct BlasInfo {            // device + host
	float3 lo; uint32_t rootNode;
	float3 hi; uint32_t numTris;
};

struct TwoLevelInputs {
	BuildInputs base;                        // device views, as for buildBvh
	std::vector<MeshRange> meshes;           // unique (indexOffset, indexCount) ranges of the object instances
	std::vector<uint32_t> meshOfInstance;    // numInstances entries
};

struct TwoLevelState {                       // everything rpt_scene_update_instances needs to rebuild the TLAS alone
	WideNode* blasNodes = nullptr;           // all BLASes, concatenated (child / triangle bases are absolute)
	TriRecord* blasTris = nullptr;           //   t1.w = triangle within the mesh, t2.w = the same (tie order within the mesh)
	WideNode* tlasNodes = nullptr;
	TriRecord* tlasLeaves = nullptr;         // t0.w = index into records[], in TLAS leaf order
	InstanceRecord* records = nullptr;       // numInstances + 1: [0] = the light triangles (world space, identity)
	BlasInfo* blasInfo = nullptr;            // device, meshes + 1 entries (the last one is the light BLAS)
	uint32_t* meshOfRecord = nullptr;        // device, numInstances + 1 entries
	uint32_t numBlasNodes = 0, numBlasTris = 0, numTlasNodes = 0, numRecords = 0, maxBlasDepth = 0, tlasDepth = 0;
	float blasMs = 0.0f, tlasMs = 0.0f;
	void release();
};

cudaError_t buildTwoLevel(const TwoLevelInputs& in, cudaStream_t stream, TwoLevelState* out);
// instance records + TLAS from the CURRENT contents of in.instances (device); frees and replaces out->tlasNodes / tlasLeaves
cudaError_t rebuildTlas(const BuildInputs& in, cudaStream_t stream, TwoLevelState* out);

} // namespace rt
