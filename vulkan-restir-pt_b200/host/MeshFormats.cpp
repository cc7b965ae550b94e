// Mesh files other than OBJ.  The reference hands every model path to assimp (src/Resource.cpp:100-181), so any format assimp
// reads is a valid `path` of a <modelInstances> entry; its shipped scenes only use OBJ (Scene.cpp has that reader).  This file
// adds the two formats scanned and simulated geometry usually comes in — Stanford PLY (ascii and binary_little_endian /
// binary_big_endian: vertex x y z [nx ny nz] [s t | u v | texture_u texture_v], face lists of any integer type) and STL (binary
// and ascii) — with the post-processing the reference asks assimp for: Triangulate (convex fan), FlipUVs, GenSmoothNormals for
// objects / GenNormals for lights when the file has no normals, FixInfacingNormals (Scene.cpp).  The material is assimp's
// default one (diffuse 0.6), which the XML's <material> then overrides as for any model.
// Not pinned against assimp itself (it is not built here: 140 MB of sources behind CMake); checked against the
// OBJ reader on the same geometry (tests/test_cpu_host_abi.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include "MeshFormats.h"

namespace rpt {

namespace {

std::string readFile(const std::string& path, const char* what) {
	std::ifstream f(path, std::ios::binary);
	if (!f) throw std::runtime_error(std::string(what) + ": cannot open " + path);
	return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

std::string lower(std::string s) {
	for (char& c : s) c = char(std::tolower(static_cast<unsigned char>(c)));
	return s;
}

// ---- PLY ---------------------------------------------------------------------------------------------------------
enum PlyType { I8, U8, I16, U16, I32, U32, F32, F64, BadType };
PlyType plyType(const std::string& t) {
	static const std::map<std::string, PlyType> names = {
		{ "char", I8 }, { "int8", I8 }, { "uchar", U8 }, { "uint8", U8 }, { "short", I16 }, { "int16", I16 }, { "ushort", U16 }, { "uint16", U16 },
		{ "int", I32 }, { "int32", I32 }, { "uint", U32 }, { "uint32", U32 }, { "float", F32 }, { "float32", F32 }, { "double", F64 }, { "float64", F64 } };
	auto it = names.find(t);
	return it == names.end() ? BadType : it->second;
}
size_t plySize(PlyType t) { static const size_t s[] = { 1, 1, 2, 2, 4, 4, 4, 8, 0 }; return s[t]; }

struct PlyProperty { std::string name; PlyType type = BadType; bool isList = false; PlyType countType = BadType; };
struct PlyElement { std::string name; size_t count = 0; std::vector<PlyProperty> props; };

struct PlyReader {
	const std::string& data; size_t pos; int format;   // 0 ascii, 1 little endian, 2 big endian
	const std::string& path;
	[[noreturn]] void fail(const std::string& why) const { throw std::runtime_error("PLY: " + path + ": " + why); }
	double number(PlyType t) {
		if (format == 0) {
			while (pos < data.size() && std::isspace(static_cast<unsigned char>(data[pos]))) pos++;
			if (pos >= data.size()) fail("unexpected end of file");
			char* end = nullptr;
			const double v = std::strtod(data.c_str() + pos, &end);
			if (end == data.c_str() + pos) fail("not a number");
			pos = size_t(end - data.c_str());
			return v;
		}
		const size_t n = plySize(t);
		if (pos + n > data.size()) fail("unexpected end of file");
		unsigned char b[8];
		std::memcpy(b, data.data() + pos, n);
		pos += n;
		if (format == 2) std::reverse(b, b + n);
		switch (t) {
		case I8: { int8_t v; std::memcpy(&v, b, 1); return v; }
		case U8: return b[0];
		case I16: { int16_t v; std::memcpy(&v, b, 2); return v; }
		case U16: { uint16_t v; std::memcpy(&v, b, 2); return v; }
		case I32: { int32_t v; std::memcpy(&v, b, 4); return v; }
		case U32: { uint32_t v; std::memcpy(&v, b, 4); return v; }
		case F32: { float v; std::memcpy(&v, b, 4); return v; }
		case F64: { double v; std::memcpy(&v, b, 8); return v; }
		default: fail("bad property type");
		}
	}
};

}  // namespace

void readPLY(const std::string& path, RawMesh& out) {
	const std::string data = readFile(path, "PLY");
	auto bad = [&](const std::string& why) -> std::runtime_error { return std::runtime_error("PLY: " + path + ": " + why); };
	// header: lines up to "end_header"
	size_t pos = 0;
	auto nextLine = [&]() {
		const size_t e = data.find('\n', pos);
		if (e == std::string::npos) throw bad("no end_header");
		std::string line = data.substr(pos, e - pos);
		pos = e + 1;
		if (!line.empty() && line.back() == '\r') line.pop_back();
		return line;
	};
	if (nextLine() != "ply") throw bad("not a PLY file");
	int format = -1;
	std::vector<PlyElement> elements;
	for (;;) {
		std::istringstream ls(nextLine());
		std::string key;
		ls >> key;
		if (key == "end_header") break;
		if (key == "format") {
			std::string f;
			ls >> f;
			format = f == "ascii" ? 0 : f == "binary_little_endian" ? 1 : f == "binary_big_endian" ? 2 : -1;
			if (format < 0) throw bad("unknown format " + f);
		}
		else if (key == "element") {
			PlyElement e;
			long long n = -1;
			ls >> e.name >> n;
			if (e.name.empty() || n < 0 || n > 400000000ll) throw bad("bad element line");
			e.count = size_t(n);
			elements.push_back(e);
		}
		else if (key == "property") {
			if (elements.empty()) throw bad("property before any element");
			PlyProperty p;
			std::string t;
			ls >> t;
			if (t == "list") {
				std::string ct, it;
				ls >> ct >> it >> p.name;
				p.isList = true; p.countType = plyType(ct); p.type = plyType(it);
				if (p.countType == BadType || p.countType == F32 || p.countType == F64) throw bad("bad list count type");
			}
			else {
				p.type = plyType(t);
				ls >> p.name;
			}
			if (p.type == BadType || p.name.empty()) throw bad("bad property line");
			elements.back().props.push_back(p);
		}
		// comment / obj_info: ignored
	}
	if (format < 0) throw bad("no format line");
	PlyReader r{ data, pos, format, path };
	out = RawMesh{};
	for (const PlyElement& e : elements) {
		if (e.name == "vertex") {
			int ix = -1, iy = -1, iz = -1, inx = -1, iny = -1, inz = -1, iu = -1, iv = -1;
			for (size_t k = 0; k < e.props.size(); k++) {
				const std::string n = lower(e.props[k].name);
				if (e.props[k].isList) continue;
				if (n == "x") ix = int(k); else if (n == "y") iy = int(k); else if (n == "z") iz = int(k);
				else if (n == "nx") inx = int(k); else if (n == "ny") iny = int(k); else if (n == "nz") inz = int(k);
				else if (n == "s" || n == "u" || n == "texture_u") iu = int(k);
				else if (n == "t" || n == "v" || n == "texture_v") iv = int(k);
			}
			if (ix < 0 || iy < 0 || iz < 0) throw bad("vertex element without x y z");
			const bool hasN = inx >= 0 && iny >= 0 && inz >= 0, hasUV = iu >= 0 && iv >= 0;
			out.pos.reserve(std::min(e.count, data.size()));   // (a vertex takes at least a byte: a corrupt count cannot reserve more than the file)
			std::vector<double> row(e.props.size());
			for (size_t i = 0; i < e.count; i++) {
				for (size_t k = 0; k < e.props.size(); k++) {
					if (e.props[k].isList) {
						const double n = r.number(e.props[k].countType);
						if (n < 0 || n > 1e6) throw bad("bad list length");
						for (size_t j = 0; j < size_t(n); j++) r.number(e.props[k].type);
						row[k] = 0;
					}
					else row[k] = r.number(e.props[k].type);
				}
				out.pos.push_back(vec3(float(row[size_t(ix)]), float(row[size_t(iy)]), float(row[size_t(iz)])));
				if (hasN) out.nrm.push_back(vec3(float(row[size_t(inx)]), float(row[size_t(iny)]), float(row[size_t(inz)])));
				if (hasUV) out.uv.push_back(vec2{ float(row[size_t(iu)]), float(row[size_t(iv)]) });
			}
		}
		else {
			const bool isFace = e.name == "face";
			for (size_t i = 0; i < e.count; i++) {
				for (const PlyProperty& p : e.props) {
					if (!p.isList) { r.number(p.type); continue; }
					const double n = r.number(p.countType);
					if (n < 0 || n > 1e6) throw bad("bad list length");
					const std::string name = lower(p.name);
					const bool indices = isFace && (name == "vertex_indices" || name == "vertex_index");
					std::vector<uint32_t> face;
					for (size_t j = 0; j < size_t(n); j++) {
						const double v = r.number(p.type);
						if (indices) {
							if (v < 0 || v >= double(out.pos.size())) throw bad("a face references a vertex that does not exist");
							face.push_back(uint32_t(v));
						}
					}
					if (indices && face.size() >= 3) out.faces.push_back(std::move(face));
				}
			}
		}
	}
	if (out.pos.empty() || out.faces.empty()) throw bad("no faces");
}

// ---- STL ---------------------------------------------------------------------------------------------------------
void readSTL(const std::string& path, RawMesh& out) {
	const std::string data = readFile(path, "STL");
	auto bad = [&](const std::string& why) -> std::runtime_error { return std::runtime_error("STL: " + path + ": " + why); };
	out = RawMesh{};
	// binary: 80-byte header, uint32 count, 50 bytes per facet — recognised by its exact size (an ascii file may start with anything
	// after "solid", and binary files that start with "solid" exist)
	bool binary = false;
	if (data.size() >= 84) {
		uint32_t n;
		std::memcpy(&n, data.data() + 80, 4);
		binary = data.size() == 84 + size_t(n) * 50;
	}
	auto addFacet = [&](vec3 n, const vec3 v[3]) {
		const uint32_t base = uint32_t(out.pos.size());
		for (int k = 0; k < 3; k++) { out.pos.push_back(v[k]); out.nrm.push_back(n); }
		out.faces.push_back({ base, base + 1, base + 2 });
	};
	if (binary) {
		uint32_t n;
		std::memcpy(&n, data.data() + 80, 4);
		for (uint32_t i = 0; i < n; i++) {
			float f[12];
			std::memcpy(f, data.data() + 84 + size_t(i) * 50, 48);
			const vec3 v[3] = { vec3(f[3], f[4], f[5]), vec3(f[6], f[7], f[8]), vec3(f[9], f[10], f[11]) };
			addFacet(vec3(f[0], f[1], f[2]), v);
		}
	}
	else {
		std::istringstream in(data);
		std::string tok;
		if (!(in >> tok) || lower(tok) != "solid") throw bad("neither a binary nor an ascii STL file");
		vec3 n(0.f), v[3];
		int nv = 0;
		while (in >> tok) {
			tok = lower(tok);
			if (tok == "facet") {
				std::string w;
				in >> w >> n.x >> n.y >> n.z;
				if (lower(w) != "normal" || !in) throw bad("bad facet line");
				nv = 0;
			}
			else if (tok == "vertex") {
				if (nv >= 3) throw bad("a facet with more than three vertices");
				in >> v[nv].x >> v[nv].y >> v[nv].z;
				if (!in) throw bad("bad vertex line");
				nv++;
			}
			else if (tok == "endfacet") {
				if (nv != 3) throw bad("a facet with fewer than three vertices");
				addFacet(n, v);
			}
		}
	}
	if (out.faces.empty()) throw bad("no facets");
	// a facet normal of (0, 0, 0) means "compute it" (many exporters write zeros): drop all normals then, they are generated
	for (const vec3& n : out.nrm) {
		if (n.x == 0.f && n.y == 0.f && n.z == 0.f) { out.nrm.clear(); break; }
	}
}

// ---- post-processing ---------------------------------------------------------------------------------------------
void triangulateRawMesh(const RawMesh& m, bool isLight, std::vector<RptMeshVertex>& verts, std::vector<uint32_t>& idx) {
	const bool hasN = m.nrm.size() == m.pos.size(), hasUV = m.uv.size() == m.pos.size();
	auto faceNormal = [&](const std::vector<uint32_t>& f) {
		const vec3 n = cross(m.pos[f[1]] - m.pos[f[0]], m.pos[f[2]] - m.pos[f[0]]);
		const float l = length(n);
		return l > 0.f ? n * (1.0f / l) : vec3(0.f, 0.f, 0.f);
	};
	auto put = [&](uint32_t v, vec3 n) {
		RptMeshVertex mv;
		mv.pos[0] = m.pos[v].x; mv.pos[1] = m.pos[v].y; mv.pos[2] = m.pos[v].z;
		mv.norm[0] = n.x; mv.norm[1] = n.y; mv.norm[2] = n.z;
		mv.uvx = hasUV ? m.uv[v].x : 0.f;
		mv.uvy = hasUV ? 1.0f - m.uv[v].y : 0.f;   // aiProcess_FlipUVs
		verts.push_back(mv);
	};
	verts.clear(); idx.clear();
	if (hasN || !isLight) {
		// shared vertices.  Without normals in the file: aiProcess_GenSmoothNormals — the normalised sum of the unit normals of
		// every face that touches a vertex AT THAT POSITION (assimp joins by position, not by index)
		std::vector<vec3> smooth;
		if (!hasN) {
			std::map<std::tuple<float, float, float>, vec3> sum;
			for (const auto& f : m.faces) {
				const vec3 n = faceNormal(f);
				for (uint32_t v : f) {
					vec3& s = sum[std::make_tuple(m.pos[v].x, m.pos[v].y, m.pos[v].z)];
					s = s + n;
				}
			}
			smooth.resize(m.pos.size());
			for (size_t v = 0; v < m.pos.size(); v++) {
				auto it = sum.find(std::make_tuple(m.pos[v].x, m.pos[v].y, m.pos[v].z));
				vec3 s = it == sum.end() ? vec3(0.f) : it->second;
				const float l = length(s);
				smooth[v] = l > 0.f ? s * (1.0f / l) : vec3(0.f, 0.f, 1.f);
			}
		}
		for (uint32_t v = 0; v < m.pos.size(); v++) put(v, hasN ? m.nrm[v] : smooth[v]);
		for (const auto& f : m.faces)
			for (size_t k = 1; k + 1 < f.size(); k++) { idx.push_back(f[0]); idx.push_back(f[k]); idx.push_back(f[k + 1]); }
	}
	else {
		// a light without normals: aiProcess_GenNormals — flat shading, one vertex per corner
		for (const auto& f : m.faces) {
			vec3 n = faceNormal(f);
			if (n.x == 0.f && n.y == 0.f && n.z == 0.f) n = vec3(0.f, 0.f, 1.f);
			const uint32_t base = uint32_t(verts.size());
			for (uint32_t v : f) put(v, n);
			for (uint32_t k = 1; k + 1 < f.size(); k++) { idx.push_back(base); idx.push_back(base + k); idx.push_back(base + k + 1); }
		}
	}
}

}  // namespace rpt
