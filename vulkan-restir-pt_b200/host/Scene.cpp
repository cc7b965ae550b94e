#include "Scene.h"
#include "MeshFormats.h"
#include "XmlLite.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace rpt {

static inline float luminance(vec3 c) { return dot(c, vec3(0.299f, 0.587f, 0.114f)); }

RptMaterial defaultMaterial() {
	RptMaterial m;
	m.baseColor[0] = m.baseColor[1] = m.baseColor[2] = 1.0f;
	m.type = Lambertian;
	m.textureIdx = InvalidResourceIdx;
	m.metallic = 0.0f;
	m.roughness = 1.0f;
	m.ior = 1.5f;
	return m;
}

// Alias-method table, 1-based ids, entry 0 = {sum, N}.  Follows the construction order of the reference
// (two explicit stacks, larger-than-one entries donate to smaller ones; src/util/AliasTable.h:26-71), so
// the resulting table — and therefore every light pick — is the same for the same power vector.
std::vector<RptLightSampleTableElement> buildAliasTable(std::vector<float> w) {
	const uint32_t n = uint32_t(w.size());
	std::vector<RptLightSampleTableElement> table(n + 1);
	float total = 0.f;
	for (float v : w) total += v;
	const float norm = float(n) / total;
	for (float& v : w) v *= norm;

	std::vector<RptLightSampleTableElement> over(n * 2 + 1), under(n * 2 + 1);
	int nOver = 0, nUnder = 0;
	for (uint32_t i = 0; i < n; i++) {
		RptLightSampleTableElement e{ w[i], i + 1 };
		if (w[i] > 1.0f) over[nOver++] = e; else under[nUnder++] = e;
	}
	while (nOver && nUnder) {
		RptLightSampleTableElement big = over[--nOver];
		RptLightSampleTableElement small = under[--nUnder];
		table[small.failId] = { small.prob, big.failId };
		big.prob -= (1.0f - small.prob);
		if (big.prob > 1.0f) over[nOver++] = big; else under[nUnder++] = big;
	}
	for (int i = nOver - 1; i >= 0; i--) table[over[i].failId] = over[i];
	for (int i = nUnder - 1; i >= 0; i--) table[under[i].failId] = under[i];
	table[0] = { total, n };
	return table;
}

// T * Rz(rot.x) * Rx(rot.y + 90) * Ry(rot.z) * S(x, z, y): OBJ assets are Y-up, the world is Z-up
mat4 ModelInstance::modelMatrix() const {
	mat4 m(1.0f);
	m = translate(m, pos);
	m = rotate(m, radians(rotation.x), vec3(0.0f, 0.0f, 1.0f));
	m = rotate(m, radians(rotation.y + 90.f), vec3(1.0f, 0.0f, 0.0f));
	m = rotate(m, radians(rotation.z), vec3(0.0f, 1.0f, 0.0f));
	m = rpt::scale(m, vec3(scale.x, scale.z, scale.y));
	return m;
}

Scene::Scene() {
	// materials[0] is a magenta placeholder (reference Resource::Resource, src/Resource.cpp:36-40)
	RptMaterial empty = defaultMaterial();
	empty.baseColor[0] = 1.f; empty.baseColor[1] = 0.f; empty.baseColor[2] = 1.f;
	materials.push_back(empty);
}

void Scene::clear() {
	*this = Scene();
}

// ---------------------------------------------------------------------------------------------------------
// OBJ import with the semantics the reference gets from assimp (src/Resource.cpp:100-181):
//   one vertex per face corner, fan triangulation, V flipped (aiProcess_FlipUVs), missing normals generated
//   (flat), aiProcess_FixInfacingNormals' bounding-box heuristic, one default material (diffuse 0.6) per file.
// ---------------------------------------------------------------------------------------------------------
namespace {

struct ObjCorner { int v, vt, vn; };

const char* skipSpace(const char* p) { while (*p == ' ' || *p == '\t') p++; return p; }

bool parseCorner(const char*& p, ObjCorner& c, int nv, int nvt, int nvn) {
	p = skipSpace(p);
	if (*p == '\0' || *p == '\n' || *p == '\r') return false;
	char* e;
	long v = std::strtol(p, &e, 10);
	if (e == p) return false;
	long vt = 0, vn = 0;
	p = e;
	if (*p == '/') {
		p++;
		if (*p != '/') { vt = std::strtol(p, &e, 10); p = e; }
		if (*p == '/') { p++; vn = std::strtol(p, &e, 10); p = e; }
	}
	c.v = int(v > 0 ? v - 1 : nv + v);
	c.vt = vt == 0 ? -1 : int(vt > 0 ? vt - 1 : nvt + vt);
	c.vn = vn == 0 ? -1 : int(vn > 0 ? vn - 1 : nvn + vn);
	return true;
}

// aiProcess_FixInfacingNormals (bounding box of positions vs positions+normals); flips normals and winding
void fixInfacingNormals(std::vector<RptMeshVertex>& verts, std::vector<uint32_t>& idx) {
	if (verts.empty()) return;
	vec3 lo1(1e10f), hi1(-1e10f), lo0(1e10f), hi0(-1e10f);
	for (auto& v : verts) {
		for (int k = 0; k < 3; k++) {
			lo1[k] = std::min(lo1[k], v.pos[k]); hi1[k] = std::max(hi1[k], v.pos[k]);
			float q = v.pos[k] + v.norm[k];
			lo0[k] = std::min(lo0[k], q); hi0[k] = std::max(hi0[k], q);
		}
	}
	vec3 d0 = hi0 - lo0, d1 = hi1 - lo1;
	for (int k = 0; k < 3; k++) if ((d0[k] > 0.f) != (d1[k] > 0.f)) return;
	if (d1.x < 0.05f * std::sqrt(d1.y * d1.z)) return;
	if (d1.y < 0.05f * std::sqrt(d1.z * d1.x)) return;
	if (d1.z < 0.05f * std::sqrt(d1.y * d1.x)) return;
	if (std::fabs(d0.x * d0.y * d0.z) < std::fabs(d1.x * d1.y * d1.z)) {
		for (auto& v : verts) for (int k = 0; k < 3; k++) v.norm[k] *= -1.0f;
		for (size_t t = 0; t + 2 < idx.size(); t += 3) std::swap(idx[t], idx[t + 2]);
	}
}

} // namespace

uint32_t Scene::addModelFromTriangles(const std::vector<RptMeshVertex>& verts, const std::vector<uint32_t>& localIdx,
                                      bool isLight, vec3 defaultDiffuse) {
	const int L = isLight ? 1 : 0;
	ModelInstance model;
	model.meshOffset = uint32_t(meshInstances[L].size());
	model.refId = uint32_t(models[L].size());

	MeshInstance mesh;
	mesh.vertexOffset = uint32_t(vertices[L].size());
	mesh.vertexCount = uint32_t(verts.size());
	mesh.indexOffset = uint32_t(indices[L].size());
	mesh.indexCount = uint32_t(localIdx.size());
	mesh.materialIdx = int(materials.size());   // material 0 of this model, offset by the pool size
	vertices[L].insert(vertices[L].end(), verts.begin(), verts.end());
	// (no exact-size reserve here: it would defeat the vector's geometric growth and make adding n models O(n^2))
	const size_t firstIndex = indices[L].size();
	indices[L].insert(indices[L].end(), localIdx.begin(), localIdx.end());
	for (size_t k = firstIndex; k < indices[L].size(); k++) indices[L][k] += mesh.vertexOffset;
	if (!isLight) {
		materialIndices.insert(materialIndices.end(), localIdx.size() / 3, mesh.materialIdx);
	}
	meshInstances[L].push_back(mesh);
	model.numMeshes = 1;
	model.numIndices = mesh.indexCount;
	model.numVertices = mesh.vertexCount;

	if (!isLight) {
		RptMaterial m = defaultMaterial();
		m.baseColor[0] = defaultDiffuse.x; m.baseColor[1] = defaultDiffuse.y; m.baseColor[2] = defaultDiffuse.z;
		materials.push_back(m);
	}
	models[L].push_back(model);
	return uint32_t(models[L].size() - 1);
}

// Any model file (reference Resource::createNewModelInstance hands the path to assimp): OBJ by the reader below, PLY and STL by
// MeshFormats.cpp; one mesh, assimp's default material (diffuse 0.6)
uint32_t Scene::addModelFromFile(const std::string& modelPath, bool isLight) {
	std::string ext;
	const size_t dot = modelPath.find_last_of('.');
	if (dot != std::string::npos) ext = modelPath.substr(dot + 1);
	for (char& c : ext) c = char(std::tolower(static_cast<unsigned char>(c)));
	if (ext == "obj") return addModelFromOBJ(modelPath, isLight);
	RawMesh raw;
	if (ext == "ply") readPLY(modelPath, raw);
	else if (ext == "stl") readSTL(modelPath, raw);
	else throw std::runtime_error("Scene: " + modelPath + ": unsupported model format (OBJ, PLY and STL are read)");
	std::vector<RptMeshVertex> verts;
	std::vector<uint32_t> idx;
	triangulateRawMesh(raw, isLight, verts, idx);
	fixInfacingNormals(verts, idx);
	const uint32_t id = addModelFromTriangles(verts, idx, isLight, vec3(0.6f));
	models[isLight ? 1 : 0][id].path = modelPath;
	return id;
}

// One material of an OBJ file's material list (assimp ObjFile::Material: diffuse defaults to 0.6)
namespace {
struct ObjMaterial {
	std::string name;
	vec3 kd = vec3(0.6f);
	std::string mapKd;
};
struct ObjMesh {
	int material = -1;   // index into the file's material list; -1 = none set (-> material 0, the default material)
	std::vector<RptMeshVertex> verts;
	std::vector<uint32_t> idx;
};
struct ObjObject {
	std::string name;
	std::vector<size_t> meshes;
};

std::string restOfLine(const char* q) {
	q = skipSpace(q);
	const char* e = q;
	while (*e && *e != '\n' && *e != '\r') e++;
	while (e > q && (e[-1] == ' ' || e[-1] == '\t')) e--;
	return std::string(q, e);
}
std::string firstWord(const char* q) {
	q = skipSpace(q);
	const char* e = q;
	while (*e && *e != '\n' && *e != '\r' && *e != ' ' && *e != '\t') e++;
	return std::string(q, e);
}
std::string parentDirOf(const std::string& p) {
	size_t k = p.find_last_of("/\\");
	return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

// .mtl: newmtl / Kd / map_Kd (assimp ObjFileMtlImporter; everything else is irrelevant to Resource::createNewModelInstance,
// which reads AI_MATKEY_COLOR_DIFFUSE and the first diffuse texture only, reference src/Resource.cpp:150-176)
void parseMtl(const std::string& mtlPath, std::vector<ObjMaterial>& lib) {
	std::ifstream f(mtlPath, std::ios::binary);
	if (!f) return;   // assimp logs an error and goes on: usemtl then creates named materials with default values
	std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	int cur = -1;
	const char* p = text.c_str();
	const char* end = p + text.size();
	while (p < end) {
		const char* line = p;
		const char* nl = static_cast<const char*>(std::memchr(p, '\n', size_t(end - p)));
		p = nl ? nl + 1 : end;
		line = skipSpace(line);
		if (std::strncmp(line, "newmtl", 6) == 0 && (line[6] == ' ' || line[6] == '\t')) {
			const std::string name = restOfLine(line + 6);
			cur = -1;
			for (size_t i = 0; i < lib.size(); i++) if (lib[i].name == name) cur = int(i);
			if (cur < 0) { ObjMaterial m; m.name = name; lib.push_back(m); cur = int(lib.size() - 1); }
		}
		else if (cur >= 0 && line[0] == 'K' && line[1] == 'd' && (line[2] == ' ' || line[2] == '\t')) {
			char* e;
			lib[cur].kd.x = std::strtof(line + 3, &e); lib[cur].kd.y = std::strtof(e, &e); lib[cur].kd.z = std::strtof(e, &e);
		}
		else if (cur >= 0 && std::strncmp(line, "map_Kd", 6) == 0 && (line[6] == ' ' || line[6] == '\t')) {
			// options (-s, -o, -bm ...) precede the file name: the name is the last token
			std::string rest = restOfLine(line + 6);
			size_t k = rest.find_last_of(" \t");
			lib[cur].mapKd = k == std::string::npos ? rest : rest.substr(k + 1);
		}
	}
}
} // namespace

// OBJ import with assimp's structure (ext/assimp/code/AssetLib/Obj/ObjFileParser.cpp, ObjFileImporter.cpp):
//   * `o` / `g` start a new object, every object starts a mesh, `usemtl` starts another mesh when the current one already
//     has faces of a different material; one aiMesh per non-empty mesh, one vertex per face corner;
//   * the material list starts with "DefaultMaterial" (diffuse 0.6), followed by the .mtl materials in file order and by
//     materials that are used but not defined; every one of them becomes a Material of the pool
//     (reference src/Resource.cpp:150-176), meshInstance.materialIdx = pool offset + index (:228-233);
//   * Resource::createNewModelInstance walks the node tree with a stack (src/Resource.cpp:129-147), i.e. it visits the
//     objects in REVERSE file order; the meshes of one object stay in order;
//   * post-process steps per mesh: fan triangulation, V flipped, flat normals where missing, FixInfacingNormals.
uint32_t Scene::addModelFromOBJ(const std::string& objPath, bool isLight) {
	std::ifstream f(objPath, std::ios::binary);
	if (!f) throw std::runtime_error("OBJ: cannot open " + objPath);
	std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	const std::string dir = parentDirOf(objPath);

	std::vector<vec3> P, N;
	std::vector<vec2> T;
	std::vector<ObjCorner> face;
	std::vector<ObjMaterial> lib(1);
	lib[0].name = "DefaultMaterial";
	std::vector<ObjMesh> meshes;
	std::vector<ObjObject> objects;
	int curObject = -1, curMesh = -1, curMaterial = -1;
	std::string activeGroup;

	auto createMesh = [&]() {
		meshes.emplace_back();
		curMesh = int(meshes.size() - 1);
		if (curObject >= 0) objects[size_t(curObject)].meshes.push_back(size_t(curMesh));
	};
	auto createObject = [&](const std::string& name) {
		objects.emplace_back();
		objects.back().name = name;
		curObject = int(objects.size() - 1);
		createMesh();
		if (curMaterial >= 0) meshes[size_t(curMesh)].material = curMaterial;
	};

	const char* p = text.c_str();
	const char* end = p + text.size();
	while (p < end) {
		const char* line = p;
		const char* nl = static_cast<const char*>(std::memchr(p, '\n', size_t(end - p)));
		p = nl ? nl + 1 : end;
		line = skipSpace(line);
		if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
			char* e; vec3 v;
			v.x = std::strtof(line + 2, &e); v.y = std::strtof(e, &e); v.z = std::strtof(e, &e);
			P.push_back(v);
		}
		else if (line[0] == 'v' && line[1] == 'n') {
			char* e; vec3 v;
			v.x = std::strtof(line + 3, &e); v.y = std::strtof(e, &e); v.z = std::strtof(e, &e);
			N.push_back(v);
		}
		else if (line[0] == 'v' && line[1] == 't') {
			char* e; vec2 v;
			v.x = std::strtof(line + 3, &e); v.y = std::strtof(e, &e);
			T.push_back(v);
		}
		else if (line[0] == 'o' && (line[1] == ' ' || line[1] == '\t')) {
			const std::string name = firstWord(line + 2);
			if (!name.empty()) {
				curObject = -1;
				for (size_t i = 0; i < objects.size(); i++) if (objects[i].name == name) { curObject = int(i); break; }
				if (curObject < 0) createObject(name);
			}
		}
		else if (line[0] == 'g' && (line[1] == ' ' || line[1] == '\t' || line[1] == '\r' || line[1] == '\n' || line[1] == 0)) {
			const std::string name = restOfLine(line + 1);
			if (activeGroup != name) { createObject(name); activeGroup = name; }
		}
		else if (std::strncmp(line, "usemtl", 6) == 0 && (line[6] == ' ' || line[6] == '\t')) {
			const std::string name = restOfLine(line + 6);
			if (name.empty() || (curMaterial >= 0 && lib[size_t(curMaterial)].name == name)) continue;
			int found = -1;
			for (size_t i = 0; i < lib.size(); i++) if (lib[i].name == name) found = int(i);
			if (found < 0) { ObjMaterial m; m.name = name; lib.push_back(m); found = int(lib.size() - 1); }
			curMaterial = found;
			const bool needsNewMesh = curMesh < 0 ||
				(meshes[size_t(curMesh)].material != -1 && meshes[size_t(curMesh)].material != found && !meshes[size_t(curMesh)].idx.empty());
			if (needsNewMesh) createMesh();
			meshes[size_t(curMesh)].material = found;
		}
		else if (std::strncmp(line, "mtllib", 6) == 0 && (line[6] == ' ' || line[6] == '\t')) {
			parseMtl(dir + "/" + restOfLine(line + 6), lib);
		}
		else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
			face.clear();
			const char* q = line + 2;
			ObjCorner c;
			while (parseCorner(q, c, int(P.size()), int(T.size()), int(N.size()))) face.push_back(c);
			if (face.size() < 3) continue;
			for (const ObjCorner& k : face)   // (assimp: "OBJ: vertex index out of range")
				if (k.v < 0 || k.v >= int(P.size())) throw std::runtime_error("OBJ: " + objPath + ": a face references a vertex that does not exist");
			if (curObject < 0) createObject("defaultobject");
			if (curMesh < 0) createMesh();
			ObjMesh& mesh = meshes[size_t(curMesh)];
			// flat normal for corners without one (aiProcess_GenNormals; all shipped assets carry vn)
			vec3 fn(0.f);
			{
				vec3 a = P[face[0].v], b = P[face[1].v], c2 = P[face[2].v];
				vec3 n = cross(b - a, c2 - a);
				float l = length(n);
				fn = l > 0.f ? n / l : vec3(0.f, 0.f, 1.f);
			}
			uint32_t base = uint32_t(mesh.verts.size());
			for (auto& k : face) {
				RptMeshVertex mv;
				vec3 pos = P[k.v];
				vec3 nrm = fn;
				if (k.vn >= 0 && k.vn < int(N.size())) nrm = N[k.vn];
				vec2 uv;
				if (k.vt >= 0 && k.vt < int(T.size())) { uv = T[k.vt]; uv.y = 1.0f - uv.y; }   // aiProcess_FlipUVs
				mv.pos[0] = pos.x; mv.pos[1] = pos.y; mv.pos[2] = pos.z; mv.uvx = uv.x;
				mv.norm[0] = nrm.x; mv.norm[1] = nrm.y; mv.norm[2] = nrm.z; mv.uvy = uv.y;
				mesh.verts.push_back(mv);
			}
			for (uint32_t k = 1; k + 1 < face.size(); k++) {   // fan from corner 0 (convex polygons)
				mesh.idx.push_back(base); mesh.idx.push_back(base + k); mesh.idx.push_back(base + k + 1);
			}
		}
	}

	// ---- Resource::createNewModelInstance (src/Resource.cpp:100-181) ----
	const int L = isLight ? 1 : 0;
	ModelInstance model;
	model.meshOffset = uint32_t(meshInstances[L].size());
	model.refId = uint32_t(models[L].size());
	model.path = objPath;
	const uint32_t materialOffset = uint32_t(materials.size());
	for (size_t o = objects.size(); o-- > 0;) {          // the reference's node stack pops the last child first
		for (size_t m : objects[o].meshes) {
			ObjMesh& src = meshes[m];
			if (src.idx.empty()) continue;
			fixInfacingNormals(src.verts, src.idx);
			MeshInstance mesh;
			mesh.vertexOffset = uint32_t(vertices[L].size());
			mesh.vertexCount = uint32_t(src.verts.size());
			mesh.indexOffset = uint32_t(indices[L].size());
			mesh.indexCount = uint32_t(src.idx.size());
			mesh.materialIdx = int(materialOffset) + (src.material >= 0 ? src.material : 0);
			vertices[L].insert(vertices[L].end(), src.verts.begin(), src.verts.end());
			const size_t firstIndex = indices[L].size();
			indices[L].insert(indices[L].end(), src.idx.begin(), src.idx.end());
			for (size_t k = firstIndex; k < indices[L].size(); k++) indices[L][k] += mesh.vertexOffset;
			if (!isLight) materialIndices.insert(materialIndices.end(), src.idx.size() / 3, mesh.materialIdx);
			meshInstances[L].push_back(mesh);
			model.numMeshes++;
			model.numIndices += mesh.indexCount;
			model.numVertices += mesh.vertexCount;
		}
	}
	if (model.numMeshes == 0) throw std::runtime_error("OBJ: no faces in " + objPath);
	if (!isLight) {
		for (const ObjMaterial& om : lib) {
			RptMaterial m = defaultMaterial();
			m.baseColor[0] = om.kd.x; m.baseColor[1] = om.kd.y; m.baseColor[2] = om.kd.z;
			m.textureIdx = InvalidResourceIdx;
			if (!om.mapKd.empty()) {
				uint32_t loaded;
				if (loadTextureFile(dir + "/" + om.mapKd, 0u, &loaded)) m.textureIdx = loaded;
			}
			materials.push_back(m);
		}
	}
	models[L].push_back(model);
	return uint32_t(models[L].size() - 1);
}

void Scene::setModelMaterial(uint32_t modelIdx, RptMaterial mat, bool overrideColor, vec3 baseColor, uint32_t textureIdx) {
	const ModelInstance& model = models[0][modelIdx];
	for (uint32_t i = 0; i < model.numMeshes; i++) {
		uint32_t materialIdx = uint32_t(meshInstances[0][i + model.meshOffset].materialIdx);
		RptMaterial m = mat;
		m.textureIdx = (textureIdx != InvalidResourceIdx) ? textureIdx : materials[materialIdx].textureIdx;
		if (overrideColor) { m.baseColor[0] = baseColor.x; m.baseColor[1] = baseColor.y; m.baseColor[2] = baseColor.z; }
		else std::memcpy(m.baseColor, materials[materialIdx].baseColor, 12);
		materials[materialIdx] = m;
	}
}

uint32_t Scene::addTexture(HostImage img) {
	textures.push_back(std::move(img));
	return uint32_t(textures.size() - 1);
}

bool Scene::loadTextureFile(const std::string& p, uint32_t filter, uint32_t* outIdx) {
	for (uint32_t i = 0; i < textures.size(); i++) {
		if (textures[i].path == p) { *outIdx = i; return true; }
	}
	HostImage img;
	// A binary PPM side-car "<file>.ppm" (tools/prepare_assets.py writes them with PIL's decoder) wins when present, so that
	// measured workloads keep their texels; otherwise the file itself is decoded (Image.cpp: PNG, JPEG, PPM).
	if (!readPPM(p + ".ppm", img) && !readImage(p, img)) return false;
	img.filter = filter;
	img.path = p;
	*outIdx = addTexture(std::move(img));
	return true;
}

void Scene::commitInstance(uint32_t modelIdx, bool isLight, vec3 power) {
	const int L = isLight ? 1 : 0;
	const ModelInstance& model = models[L][modelIdx];
	mat4 transform = model.modelMatrix();

	if (isLight) {
		if (!(length(power) > 0)) return;
		const MeshInstance& first = meshInstances[1][model.meshOffset];
		uint32_t triCount = model.numIndices / 3;
		size_t base = triangleLights.size();
		float sumArea = 0.f;
		for (uint32_t i = 0; i < triCount; i++) {
			RptTriangleLight tri{};
			vec3 v[3];
			for (int k = 0; k < 3; k++) {
				const RptMeshVertex& mv = vertices[1][indices[1][first.indexOffset + i * 3 + k]];
				vec4 w = transform * vec4(vec3(mv.pos[0], mv.pos[1], mv.pos[2]), 1.f);
				v[k] = vec3(w.x, w.y, w.z);
			}
			vec3 n = cross(v[1] - v[0], v[2] - v[0]);
			tri.area = .5f * length(n);
			n = normalize(n);
			if (model.flipNormal) n = -n;
			std::memcpy(tri.v0, &v[0], 12); std::memcpy(tri.v1, &v[1], 12); std::memcpy(tri.v2, &v[2], 12);
			tri.nx = n.x; tri.ny = n.y; tri.nz = n.z;
			sumArea += tri.area;
			triangleLights.push_back(tri);
		}
		for (size_t i = base; i < triangleLights.size(); i++) {
			triangleLights[i].radiance[0] = power.x / sumArea;
			triangleLights[i].radiance[1] = power.y / sumArea;
			triangleLights[i].radiance[2] = power.z / sumArea;
		}
	}
	else {
		RptObjectInstance inst{};
		mat4 inv = inverse(transform);
		mat4 invT = transpose(inv);
		std::memcpy(inst.transform, &transform, 64);
		std::memcpy(inst.transformInv, &inv, 64);
		std::memcpy(inst.transformInvT, &invT, 64);
		inst.radiance[0] = power.x; inst.radiance[1] = power.y; inst.radiance[2] = power.z;
		inst.indexOffset = meshInstances[0][model.meshOffset].indexOffset;
		inst.indexCount = model.numIndices;
		objectInstances.push_back(inst);
	}
}

uint32_t Scene::addModelInstanceOf(uint32_t modelIdx) {
	if (modelIdx >= models[0].size()) throw std::runtime_error("Scene: no such object model");
	ModelInstance copy = models[0][modelIdx];   // same meshOffset / numIndices: the geometry is referenced, not duplicated
	models[0].push_back(copy);
	return uint32_t(models[0].size() - 1);
}

void Scene::setObjectTransform(uint32_t modelIdx, vec3 pos, vec3 scale, vec3 rotationDeg) {
	if (modelIdx >= models[0].size() || modelIdx >= objectInstances.size()) throw std::runtime_error("Scene: no such object model");
	ModelInstance& model = models[0][modelIdx];
	model.pos = pos; model.scale = scale; model.rotation = rotationDeg;
	const mat4 transform = model.modelMatrix();
	const mat4 inv = inverse(transform);
	const mat4 invT = transpose(inv);
	RptObjectInstance& inst = objectInstances[modelIdx];   // one instance per object model, in model order (commitInstance)
	std::memcpy(inst.transform, &transform, 64);
	std::memcpy(inst.transformInv, &inv, 64);
	std::memcpy(inst.transformInvT, &invT, 64);
}

void Scene::buildLightDataStructure() {
	std::vector<float> power(triangleLights.size());
	for (size_t i = 0; i < triangleLights.size(); i++) {
		const RptTriangleLight& t = triangleLights[i];
		power[i] = luminance(vec3(t.radiance[0], t.radiance[1], t.radiance[2]) * t.area);
	}
	lightSampleTable = buildAliasTable(power);
}

RptSceneDesc Scene::desc() const {
	RptSceneDesc d{};
	d.vertices = vertices[0].data();            d.numVertices = uint32_t(vertices[0].size());
	d.indices = indices[0].data();              d.numIndices = uint32_t(indices[0].size());
	d.materials = materials.data();             d.numMaterials = uint32_t(materials.size());
	d.materialIndices = materialIndices.data(); d.numMaterialIndices = uint32_t(materialIndices.size());
	d.instances = objectInstances.data();       d.numInstances = uint32_t(objectInstances.size());
	d.triangleLights = triangleLights.data();   d.numTriangleLights = uint32_t(triangleLights.size());
	d.lightSampleTable = lightSampleTable.data();
	mTexDescs.clear();
	for (auto& t : textures) mTexDescs.push_back({ t.rgba8.data(), t.width, t.height, t.filter });
	d.textures = mTexDescs.data();              d.numTextures = uint32_t(mTexDescs.size());
	d.flags = twoLevel ? uint32_t(RPT_SCENE_TWO_LEVEL) : 0u;
	return d;
}

// ---------------------------------------------------------------------------------------------------------
// XML scene (reference src/Scene.cpp:28-190, src/Material.cpp:5-69)
// ---------------------------------------------------------------------------------------------------------
namespace {

vec3 parseVec3(const std::string& s, vec3 init = vec3(0.f)) {
	std::stringstream ss(s);
	vec3 v = init;
	ss >> v.x >> v.y >> v.z;
	return v;
}

void loadFloat(const XmlNode& node, const char* childName, float& value) {
	const XmlNode* c = node.child(childName);
	if (!c) return;
	std::stringstream ss(c->attr("value"));
	ss >> value;
}

// returns false for type "default" (keep the imported material)
bool loadMaterialNoBaseColor(const XmlNode& node, RptMaterial& m) {
	m = defaultMaterial();
	std::string type = node.attr("type");
	if (type == "default") return false;
	else if (type == "metalWorkflow") {
		loadFloat(node, "metallic", m.metallic);
		loadFloat(node, "roughness", m.roughness);
		m.type = MetalWorkflow;
	}
	else if (type == "metal") {
		loadFloat(node, "roughness", m.roughness);
		loadFloat(node, "ior", m.ior);
		m.type = Metal;
	}
	else if (type == "dielectric") {
		loadFloat(node, "ior", m.ior);
		loadFloat(node, "roughness", m.roughness);
		m.type = Dielectric;
	}
	else if (type == "thinDielectric") {
		loadFloat(node, "ior", m.ior);
		m.type = ThinDielectric;
	}
	else if (type == "lambertian") m.type = Lambertian;
	else if (type == "fake") m.type = Fake;
	// anything else (e.g. "diffuse" in ajar.xml) keeps the default-constructed Lambertian
	return true;
}

std::string parentDir(const std::string& p) {
	size_t s = p.find_last_of("/\\");
	return s == std::string::npos ? std::string(".") : p.substr(0, s);
}

} // namespace

void Scene::load(const std::string& xmlPath) {
	path = xmlPath;
	std::ifstream f(xmlPath, std::ios::binary);
	if (!f) throw std::runtime_error("Scene: cannot open " + xmlPath);
	std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	XmlParser parser(text);
	auto root = parser.parseDocument();
	if (root->name != "scene") throw std::runtime_error("Scene: failed to load");
	const std::string dir = parentDir(xmlPath);

	if (const XmlNode* integ = root->child("integrator")) {
		if (const XmlNode* size = integ->child("size")) {
			camera.setFilmSize(uint32_t(std::atoi(size->attr("width").c_str())), uint32_t(std::atoi(size->attr("height").c_str())));
		}
	}
	if (const XmlNode* cam = root->child("camera")) {
		if (const XmlNode* n = cam->child("position")) camera.setPos(parseVec3(n->attr("value")));
		if (const XmlNode* n = cam->child("angle")) camera.setAngle(parseVec3(n->attr("value")));
		else if (const XmlNode* l = cam->child("lookAt")) camera.lookAt(parseVec3(l->attr("value")));
		if (const XmlNode* n = cam->child("fov")) camera.setFOV(float(std::atof(n->attr("value").c_str())));
		if (const XmlNode* n = cam->child("lensRadius")) camera.setLensRadius(float(std::atof(n->attr("value").c_str())));
		if (const XmlNode* n = cam->child("focalDistance")) camera.setFocalDist(float(std::atof(n->attr("value").c_str())));
	}

	const XmlNode* modelsNode = root->child("modelInstances");
	if (!modelsNode) throw std::runtime_error("Scene: no <modelInstances>");
	for (auto& inst : modelsNode->children) {
		vec3 power(0.f);
		bool isLight = false;
		if (inst->attr("type") == "light") {
			if (const XmlNode* r = inst->child("radiance")) power = parseVec3(r->attr("value"));
			isLight = true;
		}
		uint32_t id = addModelFromFile(dir + "/" + inst->attr("path"), isLight);
		ModelInstance& model = models[isLight ? 1 : 0][id];
		model.name = inst->attr("name");
		if (const XmlNode* t = inst->child("transform")) {
			model.pos = parseVec3(t->attr("translate"));
			model.scale = parseVec3(t->attr("scale"));
			model.rotation = parseVec3(t->attr("rotate"));
		}
		if (inst->attr("flip") == "true") model.flipNormal = true;

		const XmlNode* matNode = inst->child("material");
		RptMaterial mat;
		if (!isLight && matNode && loadMaterialNoBaseColor(*matNode, mat)) {
			uint32_t textureIdx = InvalidResourceIdx;
			vec3 baseColor(1.0f);
			bool overrideColor = false;
			if (const XmlNode* bc = matNode->child("baseColor")) {
				if (bc->hasAttr("value")) baseColor = parseVec3(bc->attr("value"), vec3(1.0f));
				if (bc->hasAttr("image")) {
					uint32_t filter = (bc->attr("filter") == "nearest") ? 1u : 0u;
					uint32_t loaded;
					if (loadTextureFile(dir + "/" + bc->attr("image"), filter, &loaded)) textureIdx = loaded;
				}
				overrideColor = true;
			}
			setModelMaterial(id, mat, overrideColor, baseColor, textureIdx);
		}
		commitInstance(id, isLight, power);
	}
	buildLightDataStructure();
}

// ---------------------------------------------------------------------------------------------------------
// Procedural scenes
// ---------------------------------------------------------------------------------------------------------
namespace {

struct MeshBuilder {
	std::vector<RptMeshVertex> v;
	std::vector<uint32_t> i;

	// world (Z-up) -> model (Y-up) so that ModelInstance::modelMatrix() (Rx(+90deg)) maps it back
	static vec3 toModel(vec3 w) { return vec3(w.x, w.z, -w.y); }

	uint32_t vert(vec3 pw, vec3 nw, float u, float t) {
		vec3 p = toModel(pw), n = toModel(nw);
		RptMeshVertex mv{ { p.x, p.y, p.z }, u, { n.x, n.y, n.z }, t };
		v.push_back(mv);
		return uint32_t(v.size() - 1);
	}
	// quad a,b,c,d (counter-clockwise seen from the side n points to), per-corner vertices like assimp
	void quad(vec3 a, vec3 b, vec3 c, vec3 d, vec3 n, float uvScale = 1.0f) {
		uint32_t i0 = vert(a, n, 0, 0), i1 = vert(b, n, uvScale, 0), i2 = vert(c, n, uvScale, uvScale);
		i.insert(i.end(), { i0, i1, i2 });
		uint32_t j0 = vert(a, n, 0, 0), j2 = vert(c, n, uvScale, uvScale), j3 = vert(d, n, 0, uvScale);
		i.insert(i.end(), { j0, j2, j3 });
	}
	// axis-aligned box [lo,hi] rotated about Z by `deg` around its centre, outward normals
	void box(vec3 lo, vec3 hi, float deg) {
		vec3 c = (lo + hi) * 0.5f, h = (hi - lo) * 0.5f;
		float cs = std::cos(radians(deg)), sn = std::sin(radians(deg));
		auto R = [&](vec3 q) { return vec3(q.x * cs - q.y * sn, q.x * sn + q.y * cs, q.z); };
		auto P = [&](float sx, float sy, float sz) { return c + R(vec3(sx * h.x, sy * h.y, sz * h.z)); };
		quad(P(-1, -1, 1), P(1, -1, 1), P(1, 1, 1), P(-1, 1, 1), R(vec3(0, 0, 1)));
		quad(P(-1, 1, -1), P(1, 1, -1), P(1, -1, -1), P(-1, -1, -1), R(vec3(0, 0, -1)));
		quad(P(-1, -1, -1), P(1, -1, -1), P(1, -1, 1), P(-1, -1, 1), R(vec3(0, -1, 0)));
		quad(P(1, 1, -1), P(-1, 1, -1), P(-1, 1, 1), P(1, 1, 1), R(vec3(0, 1, 0)));
		quad(P(1, -1, -1), P(1, 1, -1), P(1, 1, 1), P(1, -1, 1), R(vec3(1, 0, 0)));
		quad(P(-1, 1, -1), P(-1, -1, -1), P(-1, -1, 1), P(-1, 1, 1), R(vec3(-1, 0, 0)));
	}
};

uint32_t pcgHash(uint32_t& state) {
	state = state * 747796405u + 2891336453u;
	uint32_t w = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
	return (w >> 22u) ^ w;
}
float pcgFloat(uint32_t& state) { return float(pcgHash(state) >> 8) * (1.0f / 16777216.0f); }

RptMaterial lambert() { return defaultMaterial(); }
RptMaterial metalWorkflow(float metallic, float roughness) {
	RptMaterial m = defaultMaterial();
	m.type = MetalWorkflow; m.metallic = metallic; m.roughness = roughness;
	return m;
}

uint32_t addObject(Scene& s, const MeshBuilder& mb, RptMaterial mat, vec3 color, uint32_t tex = InvalidResourceIdx,
                   vec3 pos = vec3(0.f), vec3 scl = vec3(1.f), vec3 rot = vec3(0.f)) {
	uint32_t id = s.addModelFromTriangles(mb.v, mb.i, false);
	s.models[0][id].pos = pos; s.models[0][id].scale = scl; s.models[0][id].rotation = rot;
	s.setModelMaterial(id, mat, true, color, tex);
	s.commitInstance(id, false, vec3(0.f));
	return id;
}

void addLight(Scene& s, const MeshBuilder& mb, vec3 power) {
	uint32_t id = s.addModelFromTriangles(mb.v, mb.i, true);
	s.commitInstance(id, true, power);
}

// smooth-shaded sphere-ish blob: lat-long grid displaced by a few sine lobes (analytic, seeded)
void blob(MeshBuilder& mb, vec3 centre, float radius, uint32_t nu, uint32_t nv, float bump, uint32_t seed) {
	uint32_t st = seed * 2654435761u + 12345u;
	float ph[6];
	for (float& q : ph) q = pcgFloat(st) * 6.2831853f;
	auto pos = [&](float u, float v) {
		float th = u * 6.2831853f, phi = v * 3.14159265f;
		vec3 d(std::sin(phi) * std::cos(th), std::sin(phi) * std::sin(th), std::cos(phi));
		float r = radius * (1.0f + bump * (std::sin(5 * th + ph[0]) * std::sin(4 * phi + ph[1]) * std::sin(phi)
			+ 0.5f * std::sin(11 * th + ph[2]) * std::sin(9 * phi + ph[3]) * std::sin(phi)));
		return centre + d * r;
	};
	auto nrm = [&](float u, float v) {
		const float e = 1e-3f;
		float v0 = std::min(std::max(v, e), 1.0f - e);
		vec3 du = pos(u + e, v0) - pos(u - e, v0), dv = pos(u, v0 + e) - pos(u, v0 - e);
		vec3 n = cross(du, dv);
		float l = length(n);
		vec3 out = pos(u, v0) - centre;
		if (!(l > 0.f)) return normalize(out);
		n = n / l;
		return dot(n, out) < 0 ? -n : n;
	};
	for (uint32_t j = 0; j < nv; j++) {
		for (uint32_t i = 0; i < nu; i++) {
			float u0 = float(i) / nu, u1 = float(i + 1) / nu, v0 = float(j) / nv, v1 = float(j + 1) / nv;
			vec3 a = pos(u0, v0), b = pos(u0, v1), c = pos(u1, v1), d = pos(u1, v0);
			uint32_t ia = mb.vert(a, nrm(u0, v0), u0, v0), ib = mb.vert(b, nrm(u0, v1), u0, v1), ic = mb.vert(c, nrm(u1, v1), u1, v1);
			if (j + 1 < nv) mb.i.insert(mb.i.end(), { ia, ib, ic });
			uint32_t ja = mb.vert(a, nrm(u0, v0), u0, v0), jc = mb.vert(c, nrm(u1, v1), u1, v1), jd = mb.vert(d, nrm(u1, v0), u1, v0);
			if (j > 0) mb.i.insert(mb.i.end(), { ja, jc, jd });
		}
	}
}

HostImage checkerTexture(uint32_t n, uint32_t cells, uint8_t a, uint8_t b) {
	HostImage img;
	img.width = img.height = n;
	img.rgba8.resize(size_t(n) * n * 4);
	for (uint32_t y = 0; y < n; y++) for (uint32_t x = 0; x < n; x++) {
		uint8_t c = (((x * cells / n) + (y * cells / n)) & 1) ? a : b;
		uint8_t* px = &img.rgba8[(size_t(y) * n + x) * 4];
		px[0] = c; px[1] = c; px[2] = uint8_t(c * 9 / 10); px[3] = 255;
	}
	img.path = "<checker>";
	return img;
}

} // namespace

void makeCornellBox(Scene& scene) {
	scene.clear();
	const vec3 white(0.73f), red(0.65f, 0.05f, 0.05f), green(0.12f, 0.45f, 0.15f);
	// room x in [-1,1], y in [-1,1], z in [0,2]; open towards -y (camera side)
	{ MeshBuilder m; m.quad(vec3(-1, -1, 0), vec3(1, -1, 0), vec3(1, 1, 0), vec3(-1, 1, 0), vec3(0, 0, 1)); addObject(scene, m, lambert(), white); }   // floor
	{ MeshBuilder m; m.quad(vec3(-1, 1, 2), vec3(1, 1, 2), vec3(1, -1, 2), vec3(-1, -1, 2), vec3(0, 0, -1)); addObject(scene, m, lambert(), white); }  // ceiling
	{ MeshBuilder m; m.quad(vec3(-1, 1, 0), vec3(1, 1, 0), vec3(1, 1, 2), vec3(-1, 1, 2), vec3(0, -1, 0)); addObject(scene, m, lambert(), white); }    // back
	{ MeshBuilder m; m.quad(vec3(-1, -1, 0), vec3(-1, 1, 0), vec3(-1, 1, 2), vec3(-1, -1, 2), vec3(1, 0, 0)); addObject(scene, m, lambert(), red); }   // left
	{ MeshBuilder m; m.quad(vec3(1, 1, 0), vec3(1, -1, 0), vec3(1, -1, 2), vec3(1, 1, 2), vec3(-1, 0, 0)); addObject(scene, m, lambert(), green); }    // right
	{ MeshBuilder m; m.box(vec3(-0.70f, 0.05f, 0.0f), vec3(-0.10f, 0.65f, 1.2f), 18.0f); addObject(scene, m, lambert(), white); }                      // tall box
	{ MeshBuilder m; m.box(vec3(0.10f, -0.55f, 0.0f), vec3(0.70f, 0.05f, 0.6f), -17.0f); addObject(scene, m, metalWorkflow(1.0f, 0.3f), vec3(0.93f, 0.92f, 0.92f)); } // short box
	{ MeshBuilder m; m.quad(vec3(-0.25f, -0.25f, 1.998f), vec3(-0.25f, 0.25f, 1.998f), vec3(0.25f, 0.25f, 1.998f), vec3(0.25f, -0.25f, 1.998f), vec3(0, 0, -1));
	  addLight(scene, m, vec3(5.0f)); }
	scene.buildLightDataStructure();
	scene.camera = Camera(vec3(0.0f, -3.4f, 1.0f), vec3(0.0f, 0.0f, 0.0f));
	scene.camera.setFOV(45.0f);
	scene.camera.setFilmSize(640, 360);
}

void makeAjarLikeRoom(Scene& scene, uint32_t trisTarget, uint32_t seed) {
	scene.clear();
	uint32_t checker = scene.addTexture(checkerTexture(256, 16, 200, 60));
	const vec3 wall(0.6f);
	// main room: x in [-3,3], y in [-4,4], z in [0,3]; the wall at x=-3 has a door gap y in [-0.6,0.6], z<2.2
	{ MeshBuilder m; m.quad(vec3(-3, -4, 0), vec3(3, -4, 0), vec3(3, 4, 0), vec3(-3, 4, 0), vec3(0, 0, 1), 6.0f); addObject(scene, m, metalWorkflow(0.0f, 0.5f), vec3(1.f), checker); }
	{ MeshBuilder m;
	  m.quad(vec3(-3, 4, 3), vec3(3, 4, 3), vec3(3, -4, 3), vec3(-3, -4, 3), vec3(0, 0, -1));
	  m.quad(vec3(-3, 4, 0), vec3(3, 4, 0), vec3(3, 4, 3), vec3(-3, 4, 3), vec3(0, -1, 0));
	  m.quad(vec3(3, -4, 0), vec3(-3, -4, 0), vec3(-3, -4, 3), vec3(3, -4, 3), vec3(0, 1, 0));
	  m.quad(vec3(3, 4, 0), vec3(3, -4, 0), vec3(3, -4, 3), vec3(3, 4, 3), vec3(-1, 0, 0));
	  // door wall in three pieces
	  m.quad(vec3(-3, -4, 0), vec3(-3, -0.6f, 0), vec3(-3, -0.6f, 3), vec3(-3, -4, 3), vec3(1, 0, 0));
	  m.quad(vec3(-3, 0.6f, 0), vec3(-3, 4, 0), vec3(-3, 4, 3), vec3(-3, 0.6f, 3), vec3(1, 0, 0));
	  m.quad(vec3(-3, -0.6f, 2.2f), vec3(-3, 0.6f, 2.2f), vec3(-3, 0.6f, 3), vec3(-3, -0.6f, 3), vec3(1, 0, 0));
	  // ante-room behind the door holding the light: x in [-6,-3]
	  m.quad(vec3(-6, -2, 0), vec3(-3, -2, 0), vec3(-3, 2, 0), vec3(-6, 2, 0), vec3(0, 0, 1));
	  m.quad(vec3(-6, 2, 3), vec3(-3, 2, 3), vec3(-3, -2, 3), vec3(-6, -2, 3), vec3(0, 0, -1));
	  m.quad(vec3(-6, 2, 0), vec3(-3, 2, 0), vec3(-3, 2, 3), vec3(-6, 2, 3), vec3(0, -1, 0));
	  m.quad(vec3(-3, -2, 0), vec3(-6, -2, 0), vec3(-6, -2, 3), vec3(-3, -2, 3), vec3(0, 1, 0));
	  m.quad(vec3(-6, -2, 0), vec3(-6, 2, 0), vec3(-6, 2, 3), vec3(-6, -2, 3), vec3(1, 0, 0));
	  addObject(scene, m, lambert(), wall); }
	// the door itself, ajar
	{ MeshBuilder m; m.box(vec3(-0.035f, -0.575f, 0.0f), vec3(0.035f, 0.575f, 2.2f), 0.0f);
	  addObject(scene, m, metalWorkflow(0.0f, 0.05f), vec3(0.45f, 0.25f, 0.12f), InvalidResourceIdx, vec3(-2.68f, -0.28f, 0.0f), vec3(1.f), vec3(-35.0f, 0.f, 0.f)); }
	// table
	{ MeshBuilder m; m.box(vec3(-0.2f, -1.2f, 0.85f), vec3(1.6f, 1.2f, 0.92f), 0.0f);
	  m.box(vec3(-0.1f, -1.1f, 0.0f), vec3(0.0f, -1.0f, 0.85f), 0.f); m.box(vec3(1.4f, -1.1f, 0.0f), vec3(1.5f, -1.0f, 0.85f), 0.f);
	  m.box(vec3(-0.1f, 1.0f, 0.0f), vec3(0.0f, 1.1f, 0.85f), 0.f); m.box(vec3(1.4f, 1.0f, 0.0f), vec3(1.5f, 1.1f, 0.85f), 0.f);
	  addObject(scene, m, metalWorkflow(0.0f, 0.1f), vec3(0.5f, 0.35f, 0.2f)); }
	// three pots on the table; tessellation carries the triangle budget
	uint32_t per = std::max(trisTarget / 3u, 64u);
	uint32_t nv = std::max(4u, uint32_t(std::sqrt(double(per) / 4.0)));
	uint32_t nu = 2 * nv;
	RptMaterial glass = defaultMaterial(); glass.type = Dielectric; glass.ior = 1.5f;
	{ MeshBuilder m; blob(m, vec3(0.7f, -0.7f, 1.22f), 0.3f, nu, nv, 0.06f, seed + 1); addObject(scene, m, metalWorkflow(1.0f, 0.17f), vec3(0.93f, 0.92f, 0.92f)); }
	{ MeshBuilder m; blob(m, vec3(0.7f, 0.0f, 1.22f), 0.3f, nu, nv, 0.06f, seed + 2); addObject(scene, m, metalWorkflow(0.0f, 0.3f), vec3(0.8f)); }
	{ MeshBuilder m; blob(m, vec3(0.7f, 0.7f, 1.22f), 0.3f, nu, nv, 0.06f, seed + 3); addObject(scene, m, glass, vec3(1.0f)); }
	// light in the ante-room, facing the door (+x)
	{ MeshBuilder m; m.quad(vec3(-5.9f, -0.7f, 0.4f), vec3(-5.9f, 0.7f, 0.4f), vec3(-5.9f, 0.7f, 2.6f), vec3(-5.9f, -0.7f, 2.6f), vec3(1, 0, 0));
	  addLight(scene, m, vec3(1000.0f)); }
	scene.buildLightDataStructure();
	scene.camera = Camera(vec3(2.6f, -3.2f, 1.6f));
	scene.camera.lookAt(vec3(-0.5f, 0.3f, 1.0f));
	scene.camera.setFOV(45.0f);
	scene.camera.setFilmSize(1280, 720);
}

void makeInstancedField(Scene& scene, uint32_t meshSubdiv, uint32_t gridN, uint32_t seed, bool shareGeometry) {
	scene.clear();
	const float cell = 1.0f;
	const float half = 0.5f * cell * gridN + 1.0f;
	{ MeshBuilder m;
	  m.quad(vec3(-half, -half, 0), vec3(half, -half, 0), vec3(half, half, 0), vec3(-half, half, 0), vec3(0, 0, 1));
	  m.quad(vec3(-half, half, 6), vec3(half, half, 6), vec3(half, -half, 6), vec3(-half, -half, 6), vec3(0, 0, -1));
	  m.quad(vec3(-half, half, 0), vec3(half, half, 0), vec3(half, half, 6), vec3(-half, half, 6), vec3(0, -1, 0));
	  m.quad(vec3(half, -half, 0), vec3(-half, -half, 0), vec3(-half, -half, 6), vec3(half, -half, 6), vec3(0, 1, 0));
	  m.quad(vec3(half, half, 0), vec3(half, -half, 0), vec3(half, -half, 6), vec3(half, half, 6), vec3(-1, 0, 0));
	  m.quad(vec3(-half, -half, 0), vec3(-half, half, 0), vec3(-half, half, 6), vec3(-half, -half, 6), vec3(1, 0, 0));
	  addObject(scene, m, lambert(), vec3(0.7f)); }
	// one displaced blob mesh, re-added per instance (the reference never shares geometry between
	// instances either: Resource::getModelInstanceByPath always returns nullptr, src/Resource.cpp:183-184)
	uint32_t nv = 4u << meshSubdiv, nu = 2 * nv;
	MeshBuilder proto;
	blob(proto, vec3(0.f), 0.38f, nu, nv, 0.08f, seed);
	uint32_t st = seed * 9781u + 7u;
	uint32_t protoModel = InvalidResourceIdx;
	for (uint32_t gy = 0; gy < gridN; gy++) for (uint32_t gx = 0; gx < gridN; gx++) {
		vec3 p((gx + 0.5f) * cell - 0.5f * cell * gridN + (pcgFloat(st) - 0.5f) * 0.2f,
		       (gy + 0.5f) * cell - 0.5f * cell * gridN + (pcgFloat(st) - 0.5f) * 0.2f,
		       0.45f + pcgFloat(st) * 0.6f);
		float s = 0.8f + 0.4f * pcgFloat(st);
		float k = pcgFloat(st);
		RptMaterial mat = k < 0.5f ? lambert() : metalWorkflow(k < 0.8f ? 0.0f : 1.0f, 0.15f + 0.4f * pcgFloat(st));
		vec3 col(0.35f + 0.6f * pcgFloat(st), 0.35f + 0.6f * pcgFloat(st), 0.35f + 0.6f * pcgFloat(st));
		const vec3 rot(360.f * pcgFloat(st), 0.f, 0.f);
		if (!shareGeometry || protoModel == InvalidResourceIdx) {
			const uint32_t id = addObject(scene, proto, mat, col, InvalidResourceIdx, p, vec3(s), rot);
			if (protoModel == InvalidResourceIdx) protoModel = id;
		}
		else {   // the same triangles (and the first blob's material) under another transform
			const uint32_t id = scene.addModelInstanceOf(protoModel);
			scene.models[0][id].pos = p; scene.models[0][id].scale = vec3(s); scene.models[0][id].rotation = rot;
			scene.commitInstance(id, false, vec3(0.f));
		}
	}
	{ MeshBuilder m; float q = half * 0.5f;
	  m.quad(vec3(-q, -q, 5.99f), vec3(-q, q, 5.99f), vec3(q, q, 5.99f), vec3(q, -q, 5.99f), vec3(0, 0, -1));
	  addLight(scene, m, vec3(60.0f * q * q)); }
	scene.buildLightDataStructure();
	scene.camera = Camera(vec3(-half + 0.5f, -half + 0.5f, 3.0f));
	scene.camera.lookAt(vec3(0.f, 0.f, 0.5f));
	scene.camera.setFOV(45.0f);
	scene.camera.setFilmSize(1920, 1080);
}

// ---------------------------------------------------------------------------------------------------------
// image I/O
// ---------------------------------------------------------------------------------------------------------
bool readPPM(const std::string& p, HostImage& out) {
	FILE* f = std::fopen(p.c_str(), "rb");
	if (!f) return false;
	char magic[3] = { 0 };
	int w = 0, h = 0, maxv = 0;
	bool ok = std::fscanf(f, "%2s", magic) == 1 && std::strcmp(magic, "P6") == 0;
	auto readInt = [&](int& v) {
		int c;
		for (;;) {
			c = std::fgetc(f);
			if (c == '#') { while (c != '\n' && c != EOF) c = std::fgetc(f); }
			else if (c != ' ' && c != '\t' && c != '\n' && c != '\r') break;
		}
		if (c == EOF) return false;
		std::ungetc(c, f);
		return std::fscanf(f, "%d", &v) == 1;
	};
	ok = ok && readInt(w) && readInt(h) && readInt(maxv) && maxv == 255 && w > 0 && h > 0;
	if (ok) {   // the header must not promise more than the file holds (a damaged size would otherwise allocate gigabytes)
		std::fgetc(f);
		const long here = std::ftell(f);
		std::fseek(f, 0, SEEK_END);
		const long size = std::ftell(f);
		std::fseek(f, here, SEEK_SET);
		ok = here >= 0 && size >= here && uint64_t(w) * uint64_t(h) * 3u <= uint64_t(size - here);
	}
	if (ok) {
		std::vector<uint8_t> rgb(size_t(w) * h * 3);
		ok = std::fread(rgb.data(), 1, rgb.size(), f) == rgb.size();
		if (ok) {
			out.width = uint32_t(w); out.height = uint32_t(h);
			out.rgba8.resize(size_t(w) * h * 4);
			for (size_t i = 0; i < size_t(w) * h; i++) {
				out.rgba8[i * 4 + 0] = rgb[i * 3 + 0]; out.rgba8[i * 4 + 1] = rgb[i * 3 + 1];
				out.rgba8[i * 4 + 2] = rgb[i * 3 + 2]; out.rgba8[i * 4 + 3] = 255;
			}
		}
	}
	std::fclose(f);
	return ok;
}

bool writePNG(const std::string& p, const uint8_t* rgba8, uint32_t w, uint32_t h) {
	static uint32_t crcTable[256];
	static bool init = false;
	if (!init) {
		for (uint32_t n = 0; n < 256; n++) {
			uint32_t c = n;
			for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
			crcTable[n] = c;
		}
		init = true;
	}
	auto crc = [&](const uint8_t* d, size_t n, uint32_t c) { for (size_t i = 0; i < n; i++) c = crcTable[(c ^ d[i]) & 0xff] ^ (c >> 8); return c; };
	auto be32 = [](std::vector<uint8_t>& o, uint32_t v) { o.push_back(uint8_t(v >> 24)); o.push_back(uint8_t(v >> 16)); o.push_back(uint8_t(v >> 8)); o.push_back(uint8_t(v)); };
	auto chunk = [&](std::vector<uint8_t>& o, const char* tag, const std::vector<uint8_t>& data) {
		be32(o, uint32_t(data.size()));
		size_t s = o.size();
		o.insert(o.end(), tag, tag + 4);
		o.insert(o.end(), data.begin(), data.end());
		be32(o, crc(&o[s], o.size() - s, 0xffffffffu) ^ 0xffffffffu);
	};
	std::vector<uint8_t> raw;
	raw.reserve(size_t(h) * (w * 4 + 1));
	for (uint32_t y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), rgba8 + size_t(y) * w * 4, rgba8 + size_t(y + 1) * w * 4); }
	std::vector<uint8_t> z = { 0x78, 0x01 };
	uint32_t a = 1, b = 0;
	for (size_t off = 0; off < raw.size(); off += 65535) {
		size_t n = std::min<size_t>(65535, raw.size() - off);
		z.push_back(off + n >= raw.size() ? 1 : 0);
		z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8)); z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
		z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
	}
	for (uint8_t c : raw) { a = (a + c) % 65521; b = (b + a) % 65521; }
	be32(z, (b << 16) | a);
	std::vector<uint8_t> out = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
	std::vector<uint8_t> ihdr;
	be32(ihdr, w); be32(ihdr, h);
	ihdr.insert(ihdr.end(), { 8, 6, 0, 0, 0 });
	chunk(out, "IHDR", ihdr);
	chunk(out, "IDAT", z);
	chunk(out, "IEND", {});
	FILE* f = std::fopen(p.c_str(), "wb");
	if (!f) return false;
	bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
	std::fclose(f);
	return ok;
}

} // namespace rpt
