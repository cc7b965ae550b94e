// Host scene model.  Mirrors the reference's Scene / Resource / ModelInstance / Material / DiscreteSampler1D
// (src/Scene.h:40-66, src/Resource.h:16-64, src/Model.h:44-103, src/Material.h, src/util/AliasTable.h) with the
// device-visible arrays kept in the exact layouts of include/restirpt.h, ready to hand to rpt_scene_create
// (the DeviceScene replacement).  assimp / pugixml / stb are replaced by a small OBJ reader (Triangulate +
// FlipUVs + per-corner vertices, the flags of src/Resource.cpp:107-118), XmlLite.h and a PPM reader.
#pragma once
#include <string>
#include <vector>
#include "Camera.h"
#include "../../include/restirpt.h"

namespace rpt {

constexpr uint32_t InvalidResourceIdx = 0xffffffffu;

enum MaterialType : uint32_t {   // reference src/Material.h:15-17
	Principled = 0, Lambertian, MetalWorkflow, Metal, Dielectric, ThinDielectric, Fake, Light
};

RptMaterial defaultMaterial();   // reference src/Material.h:19-26 member initialisers

// reference src/util/AliasTable.h:26-71 (DiscreteSampler1D<float>::build)
std::vector<RptLightSampleTableElement> buildAliasTable(std::vector<float> distribution);

struct HostImage {
	std::vector<uint8_t> rgba8;
	uint32_t width = 0, height = 0;
	uint32_t filter = 0;          // 0 linear, 1 nearest
	std::string path;
};

struct MeshInstance {             // reference src/Model.h:35-41
	uint32_t indexOffset = 0, indexCount = 0, vertexOffset = 0, vertexCount = 0;
	int materialIdx = int(InvalidResourceIdx);
};

struct ModelInstance {            // reference src/Model.h:44-103 (data members only)
	uint32_t meshOffset = 0, numMeshes = 0, numIndices = 0, numVertices = 0, refId = 0;
	bool flipNormal = false;
	vec3 pos = vec3(0.0f), scale = vec3(1.0f), rotation = vec3(0.0f);
	std::string name, path;
	mat4 modelMatrix() const;     // reference src/Model.cpp:11-21
};

class Scene {
public:
	// reference Scene::load (src/Scene.cpp:104-123)
	void load(const std::string& xmlPath);
	void clear();

	// programmatic construction (procedural scenes, tests): same bookkeeping as the XML path
	// returns model index into models[isLight]
	uint32_t addModelFromOBJ(const std::string& objPath, bool isLight);
	uint32_t addModelFromFile(const std::string& modelPath, bool isLight);   // by extension: .obj, .ply, .stl
	uint32_t addModelFromTriangles(const std::vector<RptMeshVertex>& verts, const std::vector<uint32_t>& localIndices,
	                               bool isLight, vec3 defaultDiffuse = vec3(0.6f));
	// finalises instance `modelIdx` (transform already set on the ModelInstance): appends an ObjectInstance or
	// the world-space TriangleLights of power `power` (reference Scene::loadModels, src/Scene.cpp:192-300)
	void commitInstance(uint32_t modelIdx, bool isLight, vec3 power);
	// another placement of object model `modelIdx` that SHARES its geometry and per-triangle materials (what a TLAS instance is);
	// returns the new model index (still one ObjectInstance per object model, in model order, once committed)
	uint32_t addModelInstanceOf(uint32_t modelIdx);
	// dynamic scenes: a new placement for object model `modelIdx` (its ObjectInstance is rewritten in place; the device
	// scene follows with Renderer::updateInstances / rpt_scene_update_instances)
	void setObjectTransform(uint32_t modelIdx, vec3 pos, vec3 scale, vec3 rotationDeg);
	void setModelMaterial(uint32_t modelIdx, RptMaterial mat, bool overrideColor, vec3 baseColor, uint32_t textureIdx);
	uint32_t addTexture(HostImage img);                       // returns texture index
	bool loadTextureFile(const std::string& path, uint32_t filter, uint32_t* outIdx);
	void buildLightDataStructure();                           // reference src/Scene.cpp:303-322

	RptSceneDesc desc() const;                                // view for rpt_scene_create
	uint32_t numTriangles() const { return uint32_t(indices[0].size() / 3 + triangleLights.size()); }

public:
	Camera camera;
	// Resource (index 0 = Object, 1 = Light, like Resource::MeshType)
	std::vector<RptMeshVertex> vertices[2];
	std::vector<uint32_t> indices[2];
	std::vector<MeshInstance> meshInstances[2];
	std::vector<ModelInstance> models[2];
	std::vector<RptMaterial> materials;
	std::vector<int32_t> materialIndices;
	std::vector<HostImage> textures;
	// Scene
	std::vector<RptObjectInstance> objectInstances;
	std::vector<RptTriangleLight> triangleLights;
	std::vector<RptLightSampleTableElement> lightSampleTable;
	std::string path;
	// acceleration-structure arrangement asked of rpt_scene_create (RptSceneFlags): BLAS per unique mesh + TLAS instead of the
	// flattened world-space structure
	bool twoLevel = false;

	Scene();

private:
	mutable std::vector<RptTextureDesc> mTexDescs;
};

// ---- procedural scenes (BASELINE.json configs 1 and 5, and the synthetic stand-in for VeachAjar) ----------
// Cornell box per SURVEY.md §8(d) config 1: 5 quad walls, 2 boxes (Lambert + metalWorkflow), 1 quad light.
void makeCornellBox(Scene& scene);
// "Ajar-like" room: closed room lit through a door gap, glossy/metal/dielectric teapot-ish blobs of
// `detail`-controlled tessellation (≈ trisTarget triangles).  Used when the VeachAjar asset is absent.
void makeAjarLikeRoom(Scene& scene, uint32_t trisTarget, uint32_t seed);
// Displaced icosphere mesh instanced on a jittered grid inside a lit box (config 5 stress scene).
// shareGeometry: the instances reference ONE copy of the mesh (and of its material) instead of each carrying their own — the
// "instanced ... through TLAS" variant of config 5, meant for two-level scenes.
void makeInstancedField(Scene& scene, uint32_t meshSubdiv, uint32_t gridN, uint32_t seed, bool shareGeometry = false);

// PNG writer (stored deflate blocks; no zlib needed) and PPM reader
bool writePNG(const std::string& path, const uint8_t* rgba8, uint32_t w, uint32_t h);
bool readPPM(const std::string& path, HostImage& out);
// PNG / JPEG / binary PPM by content (Image.cpp): what the reference's stb_image call yields, 8-bit RGBA
bool readImage(const std::string& path, HostImage& out, std::string* error = nullptr);

} // namespace rpt
