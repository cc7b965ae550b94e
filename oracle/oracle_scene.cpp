// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_scene.h).
#include <cstdio>
#include <cstdlib>
#include "oracle_scene.h"
#include <algorithm>
#include <cmath>

namespace orc {

void Scene::build(const RptSceneDesc& d) {
	vertices.assign(d.vertices, d.vertices + d.numVertices);
	indices.assign(d.indices, d.indices + d.numIndices);
	materials.assign(d.materials, d.materials + d.numMaterials);
	materialIndices.assign(d.materialIndices, d.materialIndices + d.numMaterialIndices);
	instances.assign(d.instances, d.instances + d.numInstances);
	lights.assign(d.triangleLights, d.triangleLights + d.numTriangleLights);
	lightTable.assign(d.lightSampleTable, d.lightSampleTable + d.numTriangleLights + 1);
	textures.clear();
	for (uint32_t i = 0; i < d.numTextures; i++) {
		Texture t;
		t.width = d.textures[i].width; t.height = d.textures[i].height; t.filter = d.textures[i].filter;
		t.rgba8.assign(d.textures[i].rgba8, d.textures[i].rgba8 + size_t(t.width) * t.height * 4);
		textures.push_back(std::move(t));
	}
	// sRGB EOTF, evaluated in double and rounded once (texture format is R8G8B8A8_SRGB, HostImage.cpp:22)
	for (int i = 0; i < 256; i++) {
		double c = i / 255.0;
		srgbToLinear[i] = float(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
	}

	twoLevel = (d.flags & RPT_SCENE_TWO_LEVEL) != 0;
	sets.clear(); records.clear();
	auto lightSet = [&](TriSet& set) {
		for (uint32_t i = 0; i < lights.size(); i++) {   // light BLAS: custom index 0, primitive i
			const RptTriangleLight& L = lights[i];
			WorldTri t;
			t.v0 = V3(L.v0); t.e1 = V3(L.v1) - V3(L.v0); t.e2 = V3(L.v2) - V3(L.v0);
			t.instanceIdx = 0; t.triangleIdx = i;
			set.tris.push_back(t);
		}
	};
	if (!twoLevel) {
		// flattened world-space triangle list: lights first (instance 0), then every object instance in order
		sets.emplace_back();
		TriSet& set = sets[0];
		lightSet(set);
		for (uint32_t k = 0; k < instances.size(); k++) {
			const RptObjectInstance& inst = instances[k];
			for (uint32_t j = 0; j < inst.indexCount / 3; j++) {
				vec3 w[3];
				for (int c = 0; c < 3; c++) {
					const RptMeshVertex& mv = vertices[indices[inst.indexOffset + j * 3 + c]];
					w[c] = xformPoint(inst.transform, V3(mv.pos));
				}
				WorldTri t;
				t.v0 = w[0]; t.e1 = w[1] - w[0]; t.e2 = w[2] - w[0];
				t.instanceIdx = k + 1; t.triangleIdx = j;
				set.tris.push_back(t);
			}
		}
		numFlatTris = uint32_t(set.tris.size());
		set.buildTree();
		return;
	}
	// two-level: set 0 = lights (world space), then one object-space set per unique (indexOffset, indexCount)
	sets.emplace_back();
	lightSet(sets[0]);
	InstanceRecord lr{};
	lr.r[0][0] = lr.r[1][1] = lr.r[2][2] = 1.0f;
	lr.set = 0; lr.customIndex = 0; lr.flatBase = 0;
	records.push_back(lr);
	std::vector<std::pair<uint32_t, uint32_t>> ranges;
	uint32_t flat = uint32_t(lights.size());
	for (uint32_t k = 0; k < instances.size(); k++) {
		const RptObjectInstance& inst = instances[k];
		const auto key = std::make_pair(inst.indexOffset, inst.indexCount);
		uint32_t m = uint32_t(std::find(ranges.begin(), ranges.end(), key) - ranges.begin());
		if (m == ranges.size()) {
			ranges.push_back(key);
			sets.emplace_back();
			TriSet& set = sets.back();
			for (uint32_t j = 0; j < inst.indexCount / 3; j++) {
				vec3 w[3];
				for (int c = 0; c < 3; c++) w[c] = V3(vertices[indices[inst.indexOffset + j * 3 + c]].pos);
				WorldTri t;
				t.v0 = w[0]; t.e1 = w[1] - w[0]; t.e2 = w[2] - w[0];
				t.instanceIdx = 0; t.triangleIdx = j;
				set.tris.push_back(t);
			}
		}
		InstanceRecord r{};
		const float* mi = inst.transformInv;   // column-major
		for (int row = 0; row < 3; row++) for (int c = 0; c < 4; c++) r.r[row][c] = mi[c * 4 + row];
		r.set = m + 1; r.customIndex = k + 1; r.flatBase = flat;
		records.push_back(r);
		flat += inst.indexCount / 3;
	}
	numFlatTris = flat;
	for (TriSet& set : sets) set.buildTree();
}

void Scene::TriSet::buildTree() {
	const uint32_t n = uint32_t(tris.size());
	order.resize(n);
	std::vector<vec3> cen(n);
	for (uint32_t i = 0; i < n; i++) {
		order[i] = i;
		const WorldTri& t = tris[i];
		vec3 a = t.v0, b = t.v0 + t.e1, c = t.v0 + t.e2;
		cen[i] = V3((a.x + b.x + c.x) / 3.0f, (a.y + b.y + c.y) / 3.0f, (a.z + b.z + c.z) / 3.0f);
	}
	nodes.clear();
	nodes.reserve(2 * size_t(n) + 1);
	nodes.push_back(Node{});
	if (n > 0) buildNode(0, 0, n, cen, 0);
}

static void triBounds(const WorldTri& t, float lo[3], float hi[3]) {
	vec3 p[3] = { t.v0, t.v0 + t.e1, t.v0 + t.e2 };
	float tl[3] = { 1e30f, 1e30f, 1e30f }, th[3] = { -1e30f, -1e30f, -1e30f };
	for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) {
		float v = (&p[k].x)[c];
		tl[c] = std::min(tl[c], v); th[c] = std::max(th[c], v);
	}
	// pad: e1/e2 are rounded differences and the intersector accepts barycentrics up to BaryEps outside
	// the triangle, so the box must contain that slightly fattened surface
	float ext = std::max(th[0] - tl[0], std::max(th[1] - tl[1], th[2] - tl[2]));
	for (int c = 0; c < 3; c++) {
		float pad = 4.0f * BaryEps * ext + 1e-5f * (std::max(std::fabs(tl[c]), std::fabs(th[c])) + 1.0f);
		lo[c] = std::min(lo[c], tl[c] - pad);
		hi[c] = std::max(hi[c], th[c] + pad);
	}
}

void Scene::TriSet::buildNode(uint32_t nodeIdx, uint32_t begin, uint32_t end, std::vector<vec3>& cen, int depth) {
	float lo[3] = { 1e30f, 1e30f, 1e30f }, hi[3] = { -1e30f, -1e30f, -1e30f };
	float clo[3] = { 1e30f, 1e30f, 1e30f }, chi[3] = { -1e30f, -1e30f, -1e30f };
	for (uint32_t i = begin; i < end; i++) {
		triBounds(tris[order[i]], lo, hi);
		const vec3& c = cen[order[i]];
		for (int k = 0; k < 3; k++) { clo[k] = std::min(clo[k], (&c.x)[k]); chi[k] = std::max(chi[k], (&c.x)[k]); }
	}
	Node nd;
	for (int k = 0; k < 3; k++) { nd.lo[k] = lo[k]; nd.hi[k] = hi[k]; }
	uint32_t count = end - begin;
	int axis = 0;
	for (int k = 1; k < 3; k++) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
	if (count <= 4 || depth > 60 || !(chi[axis] > clo[axis])) {
		nd.left = begin; nd.count = count;
		nodes[nodeIdx] = nd;
		return;
	}
	// binned SAH along the widest centroid axis
	const int B = 16;
	struct Bin { float lo[3], hi[3]; uint32_t n; };
	Bin bins[B];
	for (auto& b : bins) { for (int k = 0; k < 3; k++) { b.lo[k] = 1e30f; b.hi[k] = -1e30f; } b.n = 0; }
	float scale = float(B) / (chi[axis] - clo[axis]);
	auto binOf = [&](uint32_t id) {
		int b = int(((&cen[id].x)[axis] - clo[axis]) * scale);
		return std::min(std::max(b, 0), B - 1);
	};
	for (uint32_t i = begin; i < end; i++) {
		Bin& b = bins[binOf(order[i])];
		triBounds(tris[order[i]], b.lo, b.hi);
		b.n++;
	}
	auto area = [](const float* l, const float* h) {
		float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
		return dx * dy + dy * dz + dz * dx;
	};
	float rightArea[B]; uint32_t rightN[B];
	{
		float l[3] = { 1e30f, 1e30f, 1e30f }, h[3] = { -1e30f, -1e30f, -1e30f }; uint32_t c = 0;
		for (int i = B - 1; i > 0; i--) {
			if (bins[i].n) for (int k = 0; k < 3; k++) { l[k] = std::min(l[k], bins[i].lo[k]); h[k] = std::max(h[k], bins[i].hi[k]); }
			c += bins[i].n;
			rightArea[i] = c ? area(l, h) : 0.f; rightN[i] = c;
		}
	}
	int best = -1; float bestCost = 1e30f;
	{
		float l[3] = { 1e30f, 1e30f, 1e30f }, h[3] = { -1e30f, -1e30f, -1e30f }; uint32_t c = 0;
		for (int i = 0; i < B - 1; i++) {
			if (bins[i].n) for (int k = 0; k < 3; k++) { l[k] = std::min(l[k], bins[i].lo[k]); h[k] = std::max(h[k], bins[i].hi[k]); }
			c += bins[i].n;
			if (c == 0 || rightN[i + 1] == 0) continue;
			float cost = area(l, h) * c + rightArea[i + 1] * rightN[i + 1];
			if (cost < bestCost) { bestCost = cost; best = i; }
		}
	}
	uint32_t mid;
	if (best < 0) {
		mid = begin + count / 2;
		std::nth_element(order.begin() + begin, order.begin() + mid, order.begin() + end,
			[&](uint32_t a, uint32_t b) { return (&cen[a].x)[axis] < (&cen[b].x)[axis]; });
	}
	else {
		mid = uint32_t(std::partition(order.begin() + begin, order.begin() + end,
			[&](uint32_t id) { return binOf(id) <= best; }) - order.begin());
		if (mid == begin || mid == end) mid = begin + count / 2;
	}
	uint32_t left = uint32_t(nodes.size());
	nodes.push_back(Node{});
	nodes.push_back(Node{});
	nd.left = left; nd.count = 0;
	nodes[nodeIdx] = nd;
	buildNode(left, begin, mid, cen, depth + 1);
	buildNode(left + 1, mid, end, cen, depth + 1);
}

// conservative slab test in double precision (the BVH must never cull a triangle the intersector would hit)
static inline bool hitBox(const Scene::Node& n, const double o[3], const double inv[3], double tmin, double tmax) {
	double t0 = tmin, t1 = tmax;
	for (int k = 0; k < 3; k++) {
		double a = (double(n.lo[k]) - o[k]) * inv[k];
		double b = (double(n.hi[k]) - o[k]) * inv[k];
		if (a > b) std::swap(a, b);
		// NaN (0 * inf) leaves the interval untouched
		if (a > t0) t0 = a;
		if (b < t1) t1 = b;
	}
	return t0 <= t1 * 1.0000001 + 1e-9;
}

// NaN / infinite rays and empty intervals cannot pass the triangle test (all comparisons on NaN are false); the
// accelerated paths skip them instead of walking the whole tree (the NaN-blind slab test would accept every box).
// The brute-force path keeps testing every triangle, which proves the equivalence in the tests.
static inline bool degenerateRay(vec3 o, float tmin, vec3 d, float tmax) {
	return !(tmin < tmax) || !(abs_(o.x) + abs_(o.y) + abs_(o.z) + abs_(d.x) + abs_(d.y) + abs_(d.z) < 3.0e38f);
}

template <typename TFar, typename Fn>
void Scene::TriSet::candidates(vec3 o, vec3 d, float tmin, TFar tfar, bool brute, Fn fn) const {
	if (brute) {
		for (uint32_t i = 0; i < tris.size(); i++) fn(i);
		return;
	}
	if (tris.empty()) return;
	const double od[3] = { o.x, o.y, o.z };
	const double inv[3] = { 1.0 / double(d.x), 1.0 / double(d.y), 1.0 / double(d.z) };
	uint32_t stack[128];
	int sp = 0;
	stack[sp++] = 0;
	while (sp) {
		const Node& n = nodes[stack[--sp]];
		if (!hitBox(n, od, inv, double(tmin), double(tfar()))) continue;
		if (n.count) {
			for (uint32_t i = 0; i < n.count; i++) fn(order[n.left + i]);
		}
		else {
			stack[sp++] = n.left;
			stack[sp++] = n.left + 1;
		}
	}
}

// the ray in an instance's object space, in the fixed operation order of the numeric contract (CUDA: toObjectSpace)
static inline void toObjectSpace(const Scene::InstanceRecord& rec, vec3 o, vec3 d, vec3& oo, vec3& od) {
	const float (*r)[4] = rec.r;
	oo = V3(fma_(r[0][2], o.z, fma_(r[0][1], o.y, fma_(r[0][0], o.x, r[0][3]))),
	        fma_(r[1][2], o.z, fma_(r[1][1], o.y, fma_(r[1][0], o.x, r[1][3]))),
	        fma_(r[2][2], o.z, fma_(r[2][1], o.y, fma_(r[2][0], o.x, r[2][3]))));
	od = V3(fma_(r[0][2], d.z, fma_(r[0][1], d.y, r[0][0] * d.x)),
	        fma_(r[1][2], d.z, fma_(r[1][1], d.y, r[1][0] * d.x)),
	        fma_(r[2][2], d.z, fma_(r[2][1], d.y, r[2][0] * d.x)));
}

template <typename TFar, typename Fn>
void Scene::forCandidates(vec3 o, vec3 d, float tmin, TFar tfar, Fn fn) const {
	if (!twoLevel) {
		const TriSet& set = sets[0];
		set.candidates(o, d, tmin, tfar, bruteForce, [&](uint32_t i) { return fn(set.tris[i], i, set.tris[i].instanceIdx, o, d); });
		return;
	}
	for (const InstanceRecord& rec : records) {
		vec3 oo, od;
		toObjectSpace(rec, o, d, oo, od);
		// a NaN / infinite object-space ray (degenerate instance matrix) cannot pass the triangle test; the accelerated path skips it
		if (!bruteForce && degenerateRay(oo, tmin, od, 1e30f)) continue;
		const TriSet& set = sets[rec.set];
		set.candidates(oo, od, tmin, tfar, bruteForce, [&](uint32_t i) { return fn(set.tris[i], rec.flatBase + i, rec.customIndex, oo, od); });
	}
}

// debugging aid for the lock-step comparisons with the reference's shaders: prints every ray of the calling thread (ORC_RAY_LOG=1)
static const bool gRayLog = std::getenv("ORC_RAY_LOG") != nullptr;
static void logRay(const char* kind, vec3 o, float tmin, vec3 d, float tmax) {
	std::fprintf(stderr, "%s o %.9g %.9g %.9g tmin %.9g d %.9g %.9g %.9g tmax %.9g\n", kind, o.x, o.y, o.z, tmin, d.x, d.y, d.z, tmax);
}

Intersection Scene::traceClosestHit(vec3 o, float tmin, vec3 d, float tmax, bool skipLights) const {
	counters.closestRays.fetch_add(1, std::memory_order_relaxed);
	if (gRayLog) logRay("closest", o, tmin, d, tmax);
	Intersection best;
	best.bary = { 0.f, 0.f };
	best.instanceIdx = InvalidHitIndex;
	best.triangleIdx = 0;
	float bestT = tmax;
	uint32_t bestId = 0xffffffffu;
	if (!bruteForce && degenerateRay(o, tmin, d, tmax)) return best;
	forCandidates(o, d, tmin, [&] { return bestT; }, [&](const WorldTri& t, uint32_t id, uint32_t inst, vec3 ro, vec3 rd) {
		if (skipLights && inst == 0) return;
		float tt, u, v;
		// candidates at exactly the current best distance are still considered (tie rule below)
		if (intersectTri(t, ro, rd, tmin, tmax, tt, u, v)) {
			if (tt < bestT || (tt == bestT && id < bestId)) {
				bestT = tt; bestId = id;
				best.bary = { u, v };
				best.instanceIdx = inst;
				best.triangleIdx = t.triangleIdx;
			}
		}
	});
	return best;
}

bool Scene::traceShadow(vec3 o, float tmin, vec3 d, float tmax) const {
	counters.shadowRays.fetch_add(1, std::memory_order_relaxed);
	if (gRayLog) logRay("shadow ", o, tmin, d, tmax);
	if (!bruteForce && degenerateRay(o, tmin, d, tmax)) return false;
	bool hit = false;
	forCandidates(o, d, tmin, [&] { return hit ? -1.0f : tmax; }, [&](const WorldTri& t, uint32_t, uint32_t, vec3 ro, vec3 rd) {
		float tt, u, v;
		if (!hit && intersectTri(t, ro, rd, tmin, tmax, tt, u, v)) hit = true;   // (tfar < tmin from here on: every remaining box is culled)
	});
	return hit;
}

uint32_t Scene::countCandidates(vec3 o, vec3 d) const {
	uint32_t count = 0;
	if (degenerateRay(o, MinRayDistance, d, MaxRayDistance)) return 0;
	forCandidates(o, d, MinRayDistance, [] { return MaxRayDistance; }, [&](const WorldTri& t, uint32_t, uint32_t, vec3 ro, vec3 rd) {
		float tt, u, v;
		if (intersectTri(t, ro, rd, MinRayDistance, MaxRayDistance, tt, u, v)) count++;
	});
	return count;
}

static inline int wrapRepeat(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

// bilinear (weights quantised to 8 fractional bits like the sampler hardware) or nearest, REPEAT addressing,
// sRGB decode per texel before filtering
vec3 Scene::sampleTexture(uint32_t texIdx, float u, float v) const {
	const Texture& t = textures[texIdx];
	const int W = int(t.width), H = int(t.height);
	auto texel = [&](int x, int y) {
		const uint8_t* p = &t.rgba8[(size_t(y) * W + x) * 4];
		return V3(srgbToLinear[p[0]], srgbToLinear[p[1]], srgbToLinear[p[2]]);
	};
	if (!(abs_(u) < 1e6f) || !(abs_(v) < 1e6f)) { u = 0.f; v = 0.f; }
	if (t.filter == 1) {
		int x = wrapRepeat(int(std::floor(u * float(W))), W);
		int y = wrapRepeat(int(std::floor(v * float(H))), H);
		return texel(x, y);
	}
	float x = u * float(W) - 0.5f, y = v * float(H) - 0.5f;
	float fx = std::floor(x), fy = std::floor(y);
	float ax = std::floor((x - fx) * 256.0f + 0.5f) * 0.00390625f;
	float ay = std::floor((y - fy) * 256.0f + 0.5f) * 0.00390625f;
	int x0 = wrapRepeat(int(fx), W), x1 = wrapRepeat(int(fx) + 1, W);
	int y0 = wrapRepeat(int(fy), H), y1 = wrapRepeat(int(fy) + 1, H);
	vec3 top = texel(x0, y0) * (1.0f - ax) + texel(x1, y0) * ax;
	vec3 bot = texel(x0, y1) * (1.0f - ax) + texel(x1, y1) * ax;
	return top * (1.0f - ay) + bot * ay;
}

} // namespace orc
