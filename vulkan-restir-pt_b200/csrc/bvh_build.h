// Host-callable interface of the GPU BVH builder (bvh_build.cu).
#pragma once
#include <algorithm>
#include <cstring>
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
	float buildMs = 0.0f;
};

cudaError_t buildBvh(const BuildInputs& in, cudaStream_t stream, BuildOutputs* out);

} // namespace rt
