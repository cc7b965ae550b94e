// Ray traversal over the compressed 8-wide BVH.  Replaces rayQueryEXT closest-hit / any-hit traversal of the
// reference (src/shader/ray_query.glsl:6-70, driver BVH + RT cores) with hand-written sm_100a code.
//
// Semantics (identical to the CPU oracle's brute-force definition):
//   * triangle test: Möller–Trumbore on (v0, e1, e2) with a fixed operation order and a BaryEps tolerance on
//     the barycentric bounds; accepted iff tmin < t < tmax;
//   * closest hit: minimum t, ties broken towards the lower flattened triangle index — so the answer does not
//     depend on traversal order;
//   * the box tests are conservative (error-padded slabs, outward-rounded quantisation), so the BVH never
//     culls a triangle the test above accepts.
#pragma once
#include "rt_math.cuh"
#include "rt_types.cuh"

namespace rt {

struct Hit {
	float u, v;
	uint32_t instanceIdx, triangleIdx;
};

enum TraceMode { TraceClosest = 0, TraceAny = 1, TraceClosestNoLights = 2, TraceCount = 3 };

constexpr int TraversalStackSize = 48;

// Byte j of q dropped into mantissa bits 8..15 of the float `unit` (a power of two 2^k): 2^k (1 + b 2^-15), by one PRMT
// on the ALU pipe instead of an I2F.U8 on the XU pipe (a quarter of the ALU rate; 48 of them per node made XU the
// busiest pipe of the traversal kernels, profiles/r1_03_wavefront_tracequeue_details.txt)
template <int J>
RT_DEV float byteIntoMantissa(uint32_t q, uint32_t unit) {
	return __uint_as_float(__byte_perm(q, unit, 0x7604u | (uint32_t(J) << 4)));
}

// one 4-child half of a node: hit bits of slots SLOT0 .. SLOT0+3.
// Plane distances t = b * (2^e / d) + n (b = quantised byte, 2^e = the node's grid step) are evaluated as
// fma(2^(e+15) (1 + b 2^-15), 1/d, N) with N = n - 2^(e+15) / d; the rounding of N (<= 2^-8 of one grid step in t) is
// covered by the slab padding in nodeStep.  ux/uy/uz = bits of 2^(e+15) per axis (0 for a flat axis).
template <int SLOT0>
RT_DEV uint32_t slabHits4(uint32_t qnx, uint32_t qny, uint32_t qnz, uint32_t qfx, uint32_t qfy, uint32_t qfz,
                          uint32_t ux, uint32_t uy, uint32_t uz, float idx, float idy, float idz,
                          float nx, float ny, float nz, float fx, float fy, float fz, float tmin, float tmax) {
	uint32_t hits = 0;
#define RT_SLAB_CHILD(J) { \
		float tnx = fma_(byteIntoMantissa<J>(qnx, ux), idx, nx); \
		float tny = fma_(byteIntoMantissa<J>(qny, uy), idy, ny); \
		float tnz = fma_(byteIntoMantissa<J>(qnz, uz), idz, nz); \
		float tfx = fma_(byteIntoMantissa<J>(qfx, ux), idx, fx); \
		float tfy = fma_(byteIntoMantissa<J>(qfy, uy), idy, fy); \
		float tfz = fma_(byteIntoMantissa<J>(qfz, uz), idz, fz); \
		float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin)); \
		float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax)); \
		if (tn <= tf) hits |= 1u << (SLOT0 + J); \
	}
	RT_SLAB_CHILD(0) RT_SLAB_CHILD(1) RT_SLAB_CHILD(2) RT_SLAB_CHILD(3)
#undef RT_SLAB_CHILD
	return hits;
}

// bit (s ^ o) of the result = bit s of x, for an 8-bit mask and o in 0..7: slot order -> near-to-far order of a ray whose
// direction octant is o (three conditional swaps: neighbours, pairs, nibbles)
RT_DEV uint32_t permuteByOctant(uint32_t x, uint32_t o) {
	// (each swap is one bit-select of the two shifted copies; the bits a left shift carries past bit 7 fall on the side of the
	// select that takes the right-shifted copy, so the result stays an 8-bit mask)
	const uint32_t x1 = ((x << 1) & 0xaaaaaaaau) | ((x >> 1) & 0x55555555u);
	x = (o & 1u) ? x1 : x;
	const uint32_t x2 = ((x << 2) & 0xccccccccu) | ((x >> 2) & 0x33333333u);
	x = (o & 2u) ? x2 : x;
	const uint32_t x4 = ((x << 4) & 0xf0f0f0f0u) | ((x >> 4) & 0x0f0f0f0fu);
	return (o & 4u) ? x4 : x;
}

// Per-ray constants of a traversal: origin, direction, padded reciprocal direction, octant.
struct TravRay {
	float3 o, d;
	float idx, idy, idz;
	float tmin;
	uint32_t octinv;
};

// A ray with a NaN / infinite component or an empty interval can never satisfy the triangle test (every
// comparison on NaN is false).  Without this early-out such a ray would walk the whole tree: fmaxf/fminf drop
// NaN operands, so every slab test would pass.  The shaders do produce such rays now and then (e.g.
// sqrt(1 - |uv|^2) of a disk sample that rounds to |uv| > 1); the result is a miss either way.
RT_DEV bool rayIsDegenerate(float3 o, float tmin, float3 d, float tmax) {
	return !(tmin < tmax) || !(abs_(o.x) + abs_(o.y) + abs_(o.z) + abs_(d.x) + abs_(d.y) + abs_(d.z) < 3.0e38f);
}

RT_DEV float fastRcp(float x) {
	float y;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}

RT_DEV TravRay makeTravRay(float3 o, float tmin, float3 d) {
	TravRay r;
	r.o = o; r.d = d; r.tmin = tmin;
	// reciprocal direction; components too close to zero are pushed away from it so 1/d stays finite
	const float tiny = 1e-20f;
	// the reciprocal direction only feeds the conservative box tests (never the triangle test): MUFU.RCP's 1 ulp is inside the
	// slab padding of nodeStep (1e-6 relative against ~5e-7 of accumulated rounding), and saves three IEEE divisions per ray
	r.idx = fastRcp(abs_(d.x) > tiny ? d.x : copysignf(tiny, d.x));
	r.idy = fastRcp(abs_(d.y) > tiny ? d.y : copysignf(tiny, d.y));
	r.idz = fastRcp(abs_(d.z) > tiny ? d.z : copysignf(tiny, d.z));
	r.octinv = 7u ^ ((r.idx < 0.0f ? 1u : 0u) | (r.idy < 0.0f ? 2u : 0u) | (r.idz < 0.0f ? 4u : 0u));
	return r;
}

// Triangles of the leaf children of one node that the ray's interval entered.  Bit 3s + k stands for triangle k of the leaf in
// slot s; `valid` = the bits that exist in this node; the triangles are stored from triBase on in bit order.
struct LeafHits {
	uint32_t bits, valid, triBase;
};

// bit s of an 8-bit mask -> bits 3s, 3s+1, 3s+2
RT_DEV uint32_t expandSlotsToTriples(uint32_t x) {
	x = (x * 0x101u) & 0x00f00fu;
	x = (x * 0x11u) & 0x0c30c3u;
	x = (x * 0x5u) & 0x249249u;
	return x * 7u;
}

// Pops the nearest-octant child of the node group, tests its 8 child boxes against [tmin, tfar] and returns the
// new node group (inner children hit, near to far) plus the hit leaf children.  The remainder of the old
// group is pushed first.  Precondition: ngroup.y > 0x00ffffff.
template <typename Stack>
RT_DEV void nodeStep(const WideNode* __restrict__ nodes, const TravRay& r, float tfar, uint2& ngroup, Stack& stack, int& sp, LeafHits& leaves) {
	const bool negx = !(r.octinv & 1u), negy = !(r.octinv & 2u), negz = !(r.octinv & 4u);   // (makeTravRay: octinv = 7 ^ sign bits of 1/d)
	const uint32_t hits = ngroup.y;
	const uint32_t bit = 31u - uint32_t(__clz(int(hits)));
	ngroup.y &= ~(1u << bit);
	if (ngroup.y > 0x00ffffffu && sp < TraversalStackSize) stack[sp++] = ngroup;
	const uint32_t slot = (bit - 24u) ^ r.octinv;
	const uint32_t rel = __popc(hits & 0xffu & ~(0xffffffffu << slot));
	const float4* np = reinterpret_cast<const float4*>(nodes + (ngroup.x + rel));
	const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);

	const uint32_t ebits = __float_as_uint(n0.w);
	// U = 2^(e+15) per axis (the node stores e + 15; 0: flat axis, step 0), S = U / d, h = (p - o) / d: the plane of byte b is at
	// t = h + b s with s = S 2^-15.  The slab interval is padded so that it can only grow:
	//   1e-6 (255 |s| + |h|)  rounding of S, h and of the plane fma;   0.005 |s| >= 2 x 2^-9 |s|  the two roundings of N = (h - S) -+ pad
	const uint32_t ux = (ebits & 0xffu) << 23, uy = (ebits & 0xff00u) << 15, uz = (ebits & 0xff0000u) << 7;
	const float Sx = __uint_as_float(ux) * r.idx, Sy = __uint_as_float(uy) * r.idy, Sz = __uint_as_float(uz) * r.idz;
	const float hx = (n0.x - r.o.x) * r.idx, hy = (n0.y - r.o.y) * r.idy, hz = (n0.z - r.o.z) * r.idz;   // (not p/d - o/d: that cancels)
	constexpr float PadPerUnit = (255.0e-6f + 0.005f) / 32768.0f;
	const float padx = fma_(abs_(hx), 1e-6f, abs_(Sx) * PadPerUnit);
	const float pady = fma_(abs_(hy), 1e-6f, abs_(Sy) * PadPerUnit);
	const float padz = fma_(abs_(hz), 1e-6f, abs_(Sz) * PadPerUnit);
	const float cx = hx - Sx, cy = hy - Sy, cz = hz - Sz;
	const float nx = cx - padx, ny = cy - pady, nz = cz - padz;
	const float fx = cx + padx, fy = cy + pady, fz = cz + padz;

	const uint32_t qlox0 = __float_as_uint(n2.x), qlox1 = __float_as_uint(n2.y);
	const uint32_t qloy0 = __float_as_uint(n2.z), qloy1 = __float_as_uint(n2.w);
	const uint32_t qloz0 = __float_as_uint(n3.x), qloz1 = __float_as_uint(n3.y);
	const uint32_t qhix0 = __float_as_uint(n3.z), qhix1 = __float_as_uint(n3.w);
	const uint32_t qhiy0 = __float_as_uint(n4.x), qhiy1 = __float_as_uint(n4.y);
	const uint32_t qhiz0 = __float_as_uint(n4.z), qhiz1 = __float_as_uint(n4.w);

	uint32_t hitSlots = slabHits4<0>(
		negx ? qhix0 : qlox0, negy ? qhiy0 : qloy0, negz ? qhiz0 : qloz0,
		negx ? qlox0 : qhix0, negy ? qloy0 : qhiy0, negz ? qloz0 : qhiz0,
		ux, uy, uz, r.idx, r.idy, r.idz, nx, ny, nz, fx, fy, fz, r.tmin, tfar);
	hitSlots |= slabHits4<4>(
		negx ? qhix1 : qlox1, negy ? qhiy1 : qloy1, negz ? qhiz1 : qloz1,
		negx ? qlox1 : qhix1, negy ? qloy1 : qhiy1, negz ? qloz1 : qhiz1,
		ux, uy, uz, r.idx, r.idy, r.idz, nx, ny, nz, fx, fy, fz, r.tmin, tfar);

	const uint32_t imask = ebits >> 24;
	ngroup.x = __float_as_uint(n1.x);
	ngroup.y = (permuteByOctant(hitSlots & imask, r.octinv) << 24) | imask;
	leaves.valid = __float_as_uint(n1.z);
	leaves.bits = expandSlotsToTriples(hitSlots) & leaves.valid;   // (an inner or empty slot has no valid bits)
	leaves.triBase = __float_as_uint(n1.y);
}

struct TriHit {
	float t, u, v;
	uint32_t instanceIdx, triangleIdx, flat;
};

// Möller–Trumbore in the fixed operation order of the numeric contract; accepted iff tmin < t < tmax.
// The 48-byte record is loaded apart from the test so that a loop can have the next record in flight (triLoop).
struct TriData { float4 t0, t1, t2; };
RT_DEV TriData loadTri(const SceneView& s, uint32_t triIndex) {
	const float4* tp = reinterpret_cast<const float4*>(s.tris + triIndex);
	TriData d;
	d.t0 = __ldg(tp + 0); d.t1 = __ldg(tp + 1); d.t2 = __ldg(tp + 2);
	return d;
}
RT_DEV bool triTestData(const TriData& td, float3 o, float3 d, float tmin, float tmax, TriHit& h) {
	const float3 v0 = f3(td.t0), e1 = f3(td.t1), e2 = f3(td.t2);
	const float3 p = cross(d, e2);
	const float det = dot(e1, p);
	const float inv = 1.0f / det;
	const float3 sv = o - v0;
	const float u = dot(sv, p) * inv;
	const float3 q = cross(sv, e1);
	const float v = dot(d, q) * inv;
	const float t = dot(e2, q) * inv;
	h.t = t; h.u = u; h.v = v;
	h.instanceIdx = __float_as_uint(td.t0.w); h.triangleIdx = __float_as_uint(td.t1.w); h.flat = __float_as_uint(td.t2.w);
	return u >= -BaryEps && v >= -BaryEps && (u + v) <= 1.0f + BaryEps && t > tmin && t < tmax;
}
RT_DEV bool triTestRay(const SceneView& s, float3 o, float3 d, float tmin, uint32_t triIndex, float tmax, TriHit& h) {
	return triTestData(loadTri(s, triIndex), o, d, tmin, tmax, h);
}
RT_DEV bool triTest(const SceneView& s, const TravRay& r, uint32_t triIndex, float tmax, TriHit& h) {
	return triTestRay(s, r.o, r.d, r.tmin, triIndex, tmax, h);
}

// Running result of one ray (closest hit with the order-independent tie rule, any hit, or candidate count)
struct TravResult {
	Hit best;
	float bestT;
	uint32_t bestFlat, count;
	RT_DEV void init(float tmax) {
		best.u = 0.0f; best.v = 0.0f; best.instanceIdx = InvalidHitIndex; best.triangleIdx = 0;
		bestT = tmax; bestFlat = 0xffffffffu; count = 0;
	}
	// returns true when the traversal may stop (any-hit found)
	template <int MODE>
	RT_DEV bool accept(const TriHit& h) {
		if (MODE == TraceAny) {
			best.instanceIdx = h.instanceIdx; best.triangleIdx = h.triangleIdx;
			return true;
		}
		if (MODE == TraceCount) { count++; return false; }
		if (MODE == TraceClosestNoLights && h.instanceIdx == 0u) return false;
		if (h.t < bestT || (h.t == bestT && h.flat < bestFlat)) {
			bestT = h.t; bestFlat = h.flat;
			best.u = h.u; best.v = h.v;
			best.instanceIdx = h.instanceIdx; best.triangleIdx = h.triangleIdx;
		}
		return false;
	}
};

// The hit leaf triangles of one node against the ray; returns true when the traversal may stop (any hit accepted).
// INSTANCED (two-level scenes): the ray is in the object space of an instance whose BLAS these triangles belong to — the
// record's ids are mesh-local, the hit takes the instance's custom index and its place in the flattened tie order.
template <int MODE, bool INSTANCED = false>
RT_DEV bool triLoop(const SceneView& s, const TravRay& r, LeafHits leaves, float tmaxOrig, TravResult& res, uint32_t& triTests,
                    uint32_t customIndex = 0, uint32_t flatBase = 0) {
	while (leaves.bits) {
		const uint32_t one = 1u << (31u - uint32_t(__clz(int(leaves.bits))));
		leaves.bits ^= one;
		triTests++;
		TriHit h;
		if (triTest(s, r, leaves.triBase + uint32_t(__popc(leaves.valid & (one - 1u))), tmaxOrig, h)) {
			if (INSTANCED) { h.instanceIdx = customIndex; h.flat += flatBase; }
			if (res.accept<MODE>(h)) return true;
		}
	}
	return false;
}

// ---- two-level scenes (SceneView::tlasNodes != nullptr) ------------------------------------------------------------------------
// The reference's arrangement (src/Scene.cpp:448-547): a TLAS over the instances, a BLAS per mesh in object space, the ray taken
// into object space at the instance boundary — WITHOUT renormalising the direction, so t means the same on both sides and the
// running closest t prunes across instances.  Hit definition = the triangle test above applied to the object-space ray;
// closest hit = minimum t, ties to the lower flattened index (instance's first triangle + triangle within the mesh), as the
// CPU oracle's instanced brute force defines it.
struct ObjectRay { float3 o, d; };
RT_DEV InstanceRecord loadInstanceRecord(const SceneView& s, uint32_t tlasLeaf) {
	const uint32_t recIdx = __float_as_uint(__ldg(&s.tlasLeaves[tlasLeaf].t0).w);
	const float4* p = reinterpret_cast<const float4*>(s.instRecords + recIdx);
	InstanceRecord rec;
	rec.r0 = __ldg(p + 0); rec.r1 = __ldg(p + 1); rec.r2 = __ldg(p + 2);
	const float4 ids = __ldg(p + 3);
	rec.rootNode = __float_as_uint(ids.x); rec.customIndex = __float_as_uint(ids.y); rec.flatBase = __float_as_uint(ids.z); rec.pad = 0;
	return rec;
}
RT_DEV ObjectRay toObjectSpace(const InstanceRecord& rec, float3 o, float3 d) {   // fixed operation order (numeric contract)
	ObjectRay r;
	r.o = make_float3(fma_(rec.r0.z, o.z, fma_(rec.r0.y, o.y, fma_(rec.r0.x, o.x, rec.r0.w))),
	                  fma_(rec.r1.z, o.z, fma_(rec.r1.y, o.y, fma_(rec.r1.x, o.x, rec.r1.w))),
	                  fma_(rec.r2.z, o.z, fma_(rec.r2.y, o.y, fma_(rec.r2.x, o.x, rec.r2.w))));
	r.d = make_float3(fma_(rec.r0.z, d.z, fma_(rec.r0.y, d.y, rec.r0.x * d.x)),
	                  fma_(rec.r1.z, d.z, fma_(rec.r1.y, d.y, rec.r1.x * d.x)),
	                  fma_(rec.r2.z, d.z, fma_(rec.r2.y, d.y, rec.r2.x * d.x)));
	return r;
}

// One ray per thread through TLAS and BLASes, run to completion.  A real call (not inlined): it is compiled into every pass
// kernel but only two-level scenes take it, and the single-level path below must not pay registers for it.
template <int MODE>
__device__ __noinline__ Hit traceRayTwoLevel(const SceneView& s, float3 o, float tmin, float3 d, float tmax, uint32_t* candidateCount) {
	TravResult res;
	res.init(tmax);
	uint32_t nodeVisits = 0, triTests = 0;
	const TravRay rw = makeTravRay(o, tmin, d);
	const float tmaxOrig = tmax;
	uint2 stack[TraversalStackSize];
	int sp = 0;
	uint2 tgroup = make_uint2(0u, 0x80000000u);
	bool done = false;
	while (!done) {
		LeafHits inst{ 0u, 0u, 0u };
		if (tgroup.y > 0x00ffffffu) {
			nodeStep(s.tlasNodes, rw, res.bestT, tgroup, stack, sp, inst);
			nodeVisits++;
		}
		while (inst.bits && !done) {
			const uint32_t one = 1u << (31u - uint32_t(__clz(int(inst.bits))));
			inst.bits ^= one;
			const InstanceRecord rec = loadInstanceRecord(s, inst.triBase + uint32_t(__popc(inst.valid & (one - 1u))));
			if (rec.rootNode == 0xffffffffu) continue;
			const ObjectRay ob = toObjectSpace(rec, o, d);
			if (rayIsDegenerate(ob.o, tmin, ob.d, tmax)) continue;
			const TravRay r = makeTravRay(ob.o, tmin, ob.d);
			const int base = sp;
			uint2 ngroup = make_uint2(rec.rootNode, 0x80000000u);
			for (;;) {
				LeafHits leaves{ 0u, 0u, 0u };
				if (ngroup.y > 0x00ffffffu) {
					nodeStep(s.nodes, r, res.bestT, ngroup, stack, sp, leaves);
					nodeVisits++;
				}
				if (triLoop<MODE, true>(s, r, leaves, tmaxOrig, res, triTests, rec.customIndex, rec.flatBase)) { done = true; break; }
				if (ngroup.y <= 0x00ffffffu) {
					if (sp == base) break;
					ngroup = stack[--sp];
				}
			}
		}
		if (!done && tgroup.y <= 0x00ffffffu) {
			if (sp == 0) break;
			tgroup = stack[--sp];
		}
	}
	if (s.counters != nullptr) {
		atomicAdd(&s.counters[MODE == TraceAny ? 1 : 0], 1ull);
		atomicAdd(&s.counters[2], (unsigned long long)nodeVisits);
		if (MODE == TraceAny) { atomicAdd(&s.counters[5], (unsigned long long)nodeVisits); atomicAdd(&s.counters[6], (unsigned long long)triTests); }
		atomicAdd(&s.counters[3], (unsigned long long)triTests);
	}
	if (MODE == TraceCount && candidateCount) *candidateCount = res.count;
	return res.best;
}

// One ray per thread, run to completion (the per-pixel passes call this in line).
template <int MODE>
RT_DEV Hit traceRay(const SceneView& s, float3 o, float tmin, float3 d, float tmax, uint32_t* candidateCount = nullptr) {
	TravResult res;
	res.init(tmax);
	uint32_t nodeVisits = 0, triTests = 0;
	if (rayIsDegenerate(o, tmin, d, tmax)) {
		if (s.counters != nullptr) atomicAdd(&s.counters[MODE == TraceAny ? 1 : 0], 1ull);
		if (MODE == TraceCount && candidateCount) *candidateCount = 0;
		return res.best;
	}
	if (s.tlasNodes != nullptr) return traceRayTwoLevel<MODE>(s, o, tmin, d, tmax, candidateCount);
	const TravRay r = makeTravRay(o, tmin, d);
	const float tmaxOrig = tmax;
	uint2 stack[TraversalStackSize];
	int sp = 0;
	uint2 ngroup = make_uint2(0u, 0x80000000u);

	for (;;) {
		LeafHits leaves{ 0u, 0u, 0u };
		if (ngroup.y > 0x00ffffffu) {
			nodeStep(s.nodes, r, res.bestT, ngroup, stack, sp, leaves);
			nodeVisits++;
		}
		if (triLoop<MODE>(s, r, leaves, tmaxOrig, res, triTests)) break;
		if (ngroup.y <= 0x00ffffffu) {
			if (sp == 0) break;
			ngroup = stack[--sp];
		}
	}
	if (s.counters != nullptr) {
		atomicAdd(&s.counters[MODE == TraceAny ? 1 : 0], 1ull);
		atomicAdd(&s.counters[2], (unsigned long long)nodeVisits);
		if (MODE == TraceAny) { atomicAdd(&s.counters[5], (unsigned long long)nodeVisits); atomicAdd(&s.counters[6], (unsigned long long)triTests); }
		atomicAdd(&s.counters[3], (unsigned long long)triTests);
	}
	if (MODE == TraceCount && candidateCount) *candidateCount = res.count;
	return res.best;
}

// wrappers with the reference's names (ray_query.glsl:6-70)
RT_DEV Hit traceClosestHit(const SceneView& s, float3 o, float tmin, float3 d, float tmax) {
	return traceRay<TraceClosest>(s, o, tmin, d, tmax);
}
RT_DEV bool traceShadow(const SceneView& s, float3 o, float tmin, float3 d, float tmax) {
	return traceRay<TraceAny>(s, o, tmin, d, tmax).instanceIdx != InvalidHitIndex;
}
RT_DEV bool traceVisibility(const SceneView& s, float3 from, float3 to) {   // ray_query.glsl:27-38
	return !traceShadow(s, from, MinRayDistance, normalize(to - from), distance(to, from) - MinRayDistance);
}

} // namespace rt
