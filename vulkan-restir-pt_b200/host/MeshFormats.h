// PLY / STL readers and the post-processing the reference asks assimp for (see MeshFormats.cpp)
#pragma once
#include <string>
#include <vector>
#include "rmath.h"
#include "../../include/restirpt.h"

namespace rpt {

struct RawMesh {
	std::vector<vec3> pos, nrm;                 // nrm / uv: empty or one per position
	std::vector<vec2> uv;
	std::vector<std::vector<uint32_t>> faces;   // polygons (>= 3 corners, indices into pos)
};

void readPLY(const std::string& path, RawMesh& out);   // throw std::runtime_error with the reason
void readSTL(const std::string& path, RawMesh& out);
// Triangulate + FlipUVs + GenSmoothNormals (objects) / GenNormals (lights) for meshes without normals
void triangulateRawMesh(const RawMesh& m, bool isLight, std::vector<RptMeshVertex>& verts, std::vector<uint32_t>& indices);

}  // namespace rpt
