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
#ifdef RT_SLAB_I2F   // A/B experiment build (profiles/README.md): float(b) through I2F.U8, `unit` = bits of the slope s
	return __uint2float_rn((q >> (8 * J)) & 0xffu);
#else
	return __uint_as_float(__byte_perm(q, unit, 0x7604u | (uint32_t(J) << 4)));
#endif
}

// one 4-child half of a node: accumulate hit bits for children j = 0..3 of the half.
// Plane distances t = b * (2^e / d) + n (b = quantised byte, 2^e = the node's grid step) are evaluated as
// fma(2^(e+15) (1 + b 2^-15), 1/d, N) with N = n - 2^(e+15) / d; the rounding of N (<= 2^-9 of one grid step in t) is
// covered by the slab padding in nodeStep.  ux/uy/uz = bits of 2^(e+15) per axis (0 for a flat axis).
RT_DEV uint32_t slabHits4(uint32_t meta4, uint32_t octinv4,
                          uint32_t qnx, uint32_t qny, uint32_t qnz, uint32_t qfx, uint32_t qfy, uint32_t qfz,
                          uint32_t ux, uint32_t uy, uint32_t uz, float idx, float idy, float idz,
                          float nx, float ny, float nz, float fx, float fy, float fz, float tmin, float tmax) {
	uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
	uint32_t innerMask4 = (isInner4 >> 4) * 0xffu;
	uint32_t bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1f1f1f1fu;
	uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
	uint32_t hits = 0;
#define RT_SLAB_CHILD(J) { \
		const int sh = 8 * J; \
		float tnx = fma_(byteIntoMantissa<J>(qnx, ux), idx, nx); \
		float tny = fma_(byteIntoMantissa<J>(qny, uy), idy, ny); \
		float tnz = fma_(byteIntoMantissa<J>(qnz, uz), idz, nz); \
		float tfx = fma_(byteIntoMantissa<J>(qfx, ux), idx, fx); \
		float tfy = fma_(byteIntoMantissa<J>(qfy, uy), idy, fy); \
		float tfz = fma_(byteIntoMantissa<J>(qfz, uz), idz, fz); \
		float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin)); \
		float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax)); \
		if (tn <= tf) hits |= ((childBits4 >> sh) & 0xffu) << ((bitIndex4 >> sh) & 0xffu); \
	}
	RT_SLAB_CHILD(0) RT_SLAB_CHILD(1) RT_SLAB_CHILD(2) RT_SLAB_CHILD(3)
#undef RT_SLAB_CHILD
	return hits;
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
#ifdef RT_FAST_RCP
	// the reciprocal direction only feeds the conservative box tests (never the triangle test): MUFU.RCP's 1 ulp is inside the
	// slab padding of nodeStep (1e-6 relative against ~3.6e-7 of accumulated rounding), and saves three IEEE divisions per ray
	r.idx = fastRcp(abs_(d.x) > tiny ? d.x : copysignf(tiny, d.x));
	r.idy = fastRcp(abs_(d.y) > tiny ? d.y : copysignf(tiny, d.y));
	r.idz = fastRcp(abs_(d.z) > tiny ? d.z : copysignf(tiny, d.z));
#else
	r.idx = 1.0f / (abs_(d.x) > tiny ? d.x : copysignf(tiny, d.x));
	r.idy = 1.0f / (abs_(d.y) > tiny ? d.y : copysignf(tiny, d.y));
	r.idz = 1.0f / (abs_(d.z) > tiny ? d.z : copysignf(tiny, d.z));
#endif
	r.octinv = 7u ^ ((r.idx < 0.0f ? 1u : 0u) | (r.idy < 0.0f ? 2u : 0u) | (r.idz < 0.0f ? 4u : 0u));
	return r;
}

// Pops the nearest-octant child of the node group, tests its 8 child boxes against [tmin, tfar] and returns the
// new node group (inner children hit) plus the triangle hits of its leaf children.  The remainder of the old
// group is pushed first.  Precondition: ngroup.y > 0x00ffffff.
template <typename Stack>
RT_DEV void nodeStep(const SceneView& s, const TravRay& r, float tfar, uint2& ngroup, Stack& stack, int& sp, uint32_t& triBase, uint32_t& triHits) {
	const bool negx = r.idx < 0.0f, negy = r.idy < 0.0f, negz = r.idz < 0.0f;
	const uint32_t octinv4 = r.octinv * 0x01010101u;
	const uint32_t hits = ngroup.y;
	const uint32_t bit = 31u - uint32_t(__clz(int(hits)));
	ngroup.y &= ~(1u << bit);
	if (ngroup.y > 0x00ffffffu && sp < TraversalStackSize) stack[sp++] = ngroup;
	const uint32_t slot = (bit - 24u) ^ r.octinv;
	const uint32_t rel = __popc(hits & 0xffu & ~(0xffffffffu << slot));
	const float4* np = reinterpret_cast<const float4*>(s.nodes + (ngroup.x + rel));
	const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);

	const uint32_t ebits = __float_as_uint(n0.w);
	// grid step 2^(e-127) per axis (e == 0: flat axis, step 0); s = step / d and h = (p - o) / d are the plane
	// distances' slope and offset.  The slab interval is padded so that it can only grow:
	//   1e-6 (255 |s| + |h|)  rounding of s, h and of the plane fma;   0.003 |s| >= 2^-9 |s|  rounding of N (slabHits4)
	const uint32_t ebx = ebits & 0xffu, eby = (ebits >> 8) & 0xffu, ebz = (ebits >> 16) & 0xffu;
	const float sx = __uint_as_float(ebx << 23) * r.idx, sy = __uint_as_float(eby << 23) * r.idy, sz = __uint_as_float(ebz << 23) * r.idz;
	const uint32_t ux = ebx ? (ebx + 15u) << 23 : 0u, uy = eby ? (eby + 15u) << 23 : 0u, uz = ebz ? (ebz + 15u) << 23 : 0u;
	const float hx = (n0.x - r.o.x) * r.idx, hy = (n0.y - r.o.y) * r.idy, hz = (n0.z - r.o.z) * r.idz;
	const float padx = fma_(fma_(abs_(sx), 255.0f, abs_(hx)), 1e-6f, abs_(sx) * 0.003f);
	const float pady = fma_(fma_(abs_(sy), 255.0f, abs_(hy)), 1e-6f, abs_(sy) * 0.003f);
	const float padz = fma_(fma_(abs_(sz), 255.0f, abs_(hz)), 1e-6f, abs_(sz) * 0.003f);
	const float Sx = __uint_as_float(ux) * r.idx, Sy = __uint_as_float(uy) * r.idy, Sz = __uint_as_float(uz) * r.idz;
#ifdef RT_SLAB_I2F
	const float nx = hx - padx, ny = hy - pady, nz = hz - padz;
	const float fx = hx + padx, fy = hy + pady, fz = hz + padz;
	(void)Sx; (void)Sy; (void)Sz;
#define RT_SLAB_ARGS ux, uy, uz, sx, sy, sz
#else
	const float nx = (hx - padx) - Sx, ny = (hy - pady) - Sy, nz = (hz - padz) - Sz;
	const float fx = (hx + padx) - Sx, fy = (hy + pady) - Sy, fz = (hz + padz) - Sz;
#define RT_SLAB_ARGS ux, uy, uz, r.idx, r.idy, r.idz
#endif

	const uint32_t qlox0 = __float_as_uint(n2.x), qlox1 = __float_as_uint(n2.y);
	const uint32_t qloy0 = __float_as_uint(n2.z), qloy1 = __float_as_uint(n2.w);
	const uint32_t qloz0 = __float_as_uint(n3.x), qloz1 = __float_as_uint(n3.y);
	const uint32_t qhix0 = __float_as_uint(n3.z), qhix1 = __float_as_uint(n3.w);
	const uint32_t qhiy0 = __float_as_uint(n4.x), qhiy1 = __float_as_uint(n4.y);
	const uint32_t qhiz0 = __float_as_uint(n4.z), qhiz1 = __float_as_uint(n4.w);

	uint32_t hitmask = slabHits4(__float_as_uint(n1.z), octinv4,
		negx ? qhix0 : qlox0, negy ? qhiy0 : qloy0, negz ? qhiz0 : qloz0,
		negx ? qlox0 : qhix0, negy ? qloy0 : qhiy0, negz ? qloz0 : qhiz0,
		RT_SLAB_ARGS, nx, ny, nz, fx, fy, fz, r.tmin, tfar);
	hitmask |= slabHits4(__float_as_uint(n1.w), octinv4,
		negx ? qhix1 : qlox1, negy ? qhiy1 : qloy1, negz ? qhiz1 : qloz1,
		negx ? qlox1 : qhix1, negy ? qloy1 : qhiy1, negz ? qloz1 : qhiz1,
		RT_SLAB_ARGS, nx, ny, nz, fx, fy, fz, r.tmin, tfar);

	ngroup.x = __float_as_uint(n1.x);
	ngroup.y = (hitmask & 0xff000000u) | (ebits >> 24);
	triBase = __float_as_uint(n1.y);
	triHits = hitmask & 0x00ffffffu;
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

// The triangles of one node's leaf children (bits of triHits, relative to triBase) against the ray; returns true when the
// traversal may stop (any hit accepted).  RT_TRI_PIPE: the next triangle's record is requested before the current one is
// tested, so its L1 / L2 latency overlaps the test instead of following it (the loop runs at ~8 of 32 lanes and a third of the
// traversal kernels' stall samples sit on these loads, profiles/r2_03_*).
template <int MODE>
RT_DEV bool triLoop(const SceneView& s, const TravRay& r, uint32_t triBase, uint32_t triHits, float tmaxOrig, TravResult& res, uint32_t& triTests) {
#ifdef RT_TRI_PIPE
	if (!triHits) return false;
	TriData cur = loadTri(s, triBase + uint32_t(__ffs(int(triHits))) - 1u);
	triHits &= triHits - 1u;
	for (;;) {
		const bool more = triHits != 0u;
		TriData nxt = cur;
		if (more) {
			nxt = loadTri(s, triBase + uint32_t(__ffs(int(triHits))) - 1u);
			triHits &= triHits - 1u;
		}
		triTests++;
		TriHit h;
		if (triTestData(cur, r.o, r.d, r.tmin, tmaxOrig, h)) {
			if (res.accept<MODE>(h)) return true;
		}
		if (!more) return false;
		cur = nxt;
	}
#else
	while (triHits) {
		const uint32_t i = uint32_t(__ffs(int(triHits))) - 1u;
		triHits &= triHits - 1u;
		triTests++;
		TriHit h;
		if (triTest(s, r, triBase + i, tmaxOrig, h)) {
			if (res.accept<MODE>(h)) return true;
		}
	}
	return false;
#endif
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
	const TravRay r = makeTravRay(o, tmin, d);
	const float tmaxOrig = tmax;
	uint2 stack[TraversalStackSize];
	int sp = 0;
	uint2 ngroup = make_uint2(0u, 0x80000000u);

	for (;;) {
		uint32_t triBase = 0, triHits = 0;
		if (ngroup.y > 0x00ffffffu) {
			nodeStep(s, r, res.bestT, ngroup, stack, sp, triBase, triHits);
			nodeVisits++;
		}
		while (triHits) {
			const uint32_t i = uint32_t(__ffs(int(triHits))) - 1u;
			triHits &= triHits - 1u;
			triTests++;
			TriHit h;
			if (triTest(s, r, triBase + i, tmaxOrig, h)) {
				if (res.accept<MODE>(h)) goto done;
			}
		}
		if (ngroup.y <= 0x00ffffffu) {
			if (sp == 0) break;
			ngroup = stack[--sp];
		}
	}
done:
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
