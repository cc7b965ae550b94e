// Wavefront ray traversal: a persistent kernel that pulls rays from a queue in global memory.
//
// Replaces the ray-query calls of the reference (src/shader/ray_query.glsl:6-70, executed by the driver's RT cores)
// for the wavefront passes.  The RT cores hide the fact that the rays of one SIMD group need very different numbers
// of BVH steps; on B200 that shows up as idle lanes (profiles/: 7-9 of 32 lanes active when each thread keeps its
// ray until the slowest ray of the warp is done).  Here rays are decoupled from threads: every lane runs the
// while-while traversal loop on its current ray, and as soon as enough lanes of the warp have finished theirs the
// idle lanes fetch the next rays of the queue (one atomic per warp) — "persistent threads with dynamic fetch"
// (Aila & Laine 2009), over the compressed 8-wide BVH.
//
// Queue entry i: rays[2i] = {origin, tmin}, rays[2i+1] = {direction, tmax}.  Results land at index i: a 16-byte
// RptIntersection (closest hit) or one byte (occluded).  Per-ray results are identical to traceRay<> (same box and
// triangle arithmetic, order-independent closest-hit rule).
#include "passes.h"
#include "persist.cuh"
#include "bvh_traverse.cuh"

namespace rt {

namespace {

constexpr int TraceBlock = 128;
constexpr uint32_t NoRay = 0xffffffffu;
constexpr int FetchThreshold = 8;   // idle lanes per warp that trigger a fetch

// resident blocks per SM the kernel is compiled for (0 = leave the register count to the compiler: 64 / 56 registers)
#ifndef RT_TRACE_MINBLOCKS
#define RT_TRACE_MINBLOCKS 0
#endif
#if RT_TRACE_MINBLOCKS > 0
#define RT_TRACE_BOUNDS __launch_bounds__(TraceBlock, RT_TRACE_MINBLOCKS)
#else
#define RT_TRACE_BOUNDS __launch_bounds__(TraceBlock)
#endif
template <int MODE>
__global__ void RT_TRACE_BOUNDS traceQueueKernel(const __grid_constant__ SceneView s, const float4* __restrict__ rays,
                                                                const uint32_t* __restrict__ countPtr, uint32_t countHost, uint32_t* __restrict__ head,
                                                                RptIntersection* __restrict__ hits, uint8_t* __restrict__ occluded) {
	const uint32_t n = countPtr ? *countPtr : countHost;
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t rayIdx = NoRay;
	bool dry = false;
	TravRay r = makeTravRay(f3(0.0f), 0.0f, f3(1.0f));
	TravResult res;
	res.init(0.0f);
	float tmaxOrig = 0.0f;
	uint2 stack[TraversalStackSize];
	int sp = 0;
	uint2 ngroup = make_uint2(0u, 0u);
	uint32_t nodeVisits = 0, triTests = 0, visitsAtFetch = 0;

	for (;;) {
		// ---- dynamic fetch ----------------------------------------------------------------------------------
		// (Measured and rejected, profiles/r1_12_*: issuing the queue atomic one loop iteration before its result is used, so
		// that its round trip overlaps a traversal step — the lanes that wait idle for it cost more than the stall saves.)
		const unsigned idleMask = __ballot_sync(FullWarp, rayIdx == NoRay);
		if (!dry && __popc(idleMask) >= FetchThreshold) {
			const int leader = __ffs(int(idleMask)) - 1;
			const uint32_t want = uint32_t(__popc(idleMask));
			uint32_t base = 0;
			if (int(lane) == leader) base = atomicAdd(head, want);
			base = __shfl_sync(FullWarp, base, leader);
			if (base + want >= n) dry = true;
			if (rayIdx == NoRay) {
				const uint32_t idx = base + uint32_t(__popc(idleMask & ((1u << lane) - 1u)));
				if (idx < n) {
					const float4 a = __ldcs(rays + 2 * size_t(idx)), b = __ldcs(rays + 2 * size_t(idx) + 1);
					if (rayIsDegenerate(f3(a), a.w, f3(b), b.w)) {
						if (MODE == TraceAny) occluded[idx] = 0;
						else { RptIntersection o; o.bary[0] = 0.f; o.bary[1] = 0.f; o.instanceIdx = InvalidHitIndex; o.triangleIdx = 0; hits[idx] = o; }
					}
					else {
						if (s.counters != nullptr) atomicAdd(&s.counters[MODE == TraceAny ? 1 : 0], 1ull);   // (empty-interval placeholders are not rays)
						rayIdx = idx;
						r = makeTravRay(f3(a), a.w, f3(b));
						tmaxOrig = b.w;
						res.init(b.w);
						sp = 0;
						ngroup = make_uint2(0u, 0x80000000u);
						visitsAtFetch = nodeVisits;
					}
				}
			}
		}
		if (__all_sync(FullWarp, rayIdx == NoRay)) {
			if (dry) break;
			continue;
		}

		// ---- one traversal step of every lane that holds a ray -----------------------------------------------
		if (rayIdx != NoRay) {
			bool finished = false;
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
					if (res.accept<MODE>(h)) { finished = true; break; }
				}
			}
			if (!finished && ngroup.y <= 0x00ffffffu) {
				if (sp == 0) finished = true;
				else ngroup = stack[--sp];
			}
			if (finished) {
				if (s.counters != nullptr) atomicMax(&s.counters[7], (unsigned long long)(nodeVisits - visitsAtFetch));
				if (MODE == TraceAny) occluded[rayIdx] = res.best.instanceIdx != InvalidHitIndex ? 1 : 0;
				else {
					RptIntersection o;
					o.bary[0] = res.best.u; o.bary[1] = res.best.v; o.instanceIdx = res.best.instanceIdx; o.triangleIdx = res.best.triangleIdx;
					hits[rayIdx] = o;
				}
				rayIdx = NoRay;
			}
		}
	}
	if (s.counters != nullptr) {
		// per-thread totals (a thread serves many rays); the rays themselves were counted at fetch time
		atomicAdd(&s.counters[2], (unsigned long long)nodeVisits);
		if (MODE == TraceAny) { atomicAdd(&s.counters[5], (unsigned long long)nodeVisits); atomicAdd(&s.counters[6], (unsigned long long)triTests); }
		atomicAdd(&s.counters[3], (unsigned long long)triTests);
	}
}


// ---- variant: warp-cooperative triangle rounds (RT_TRI_COMPACT) ---------------------------------------------------
// The node step above runs at ~28 of 32 lanes, the per-lane triangle loop at ~8 (profiles/r1_14_*: most node steps yield no
// triangle, a few yield several, so the loop runs as long as the fullest lane).  Here every round's (ray, triangle) pairs are
// listed in shared memory and tested by full warps: lane j tests pair j with the OWNER lane's ray (origin / direction staged
// in shared memory at fetch time).  Results return through native 32-bit shared atomics only (a 64-bit shared atomicMin is
// a CAS loop, which is what sank the r1_16 attempt): pass 1 atomicMin on the order-mapped bits of t, pass 2 atomicMin on the
// flattened index among the lanes that hold the minimum t, pass 3 the unique winner publishes {u, v, ids}; the owner merges
// it with the order-independent closest-hit rule.  Same triangle arithmetic, same rule => same bits as traceRay<>.
constexpr uint32_t NoCand = 0xffffffffu;
struct __align__(16) WarpScratch {
	float4 rayA[32];        // origin, tmin      of the lane's current ray
	float4 rayB[32];        // direction, tmax
	uint4 win[32];          // closest hit: {u, v, instanceIdx, triangleIdx} of the round's winner for this owner
	uint32_t keyT[32];      // closest hit: order-mapped bits of the best t so far; any hit: 1 = occluded
	uint32_t candFlat[32];  // lowest flattened index among this chunk's candidates at t == keyT (NoCand: none)
	uint32_t triBase[32];   // first triangle of the owner's current node
	uint32_t count;         // pairs listed this round
	uint32_t pad[3];
	uint16_t list[32 * 24]; // owner lane << 5 | triangle bit of the owner's hit mask
};

RT_DEV uint32_t orderKey(float t) {   // monotone map float -> uint32
	const uint32_t b = __float_as_uint(t);
	return b ^ (uint32_t(int32_t(b) >> 31) | 0x80000000u);
}
RT_DEV float orderKeyInv(uint32_t k) {
	return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

template <int MODE>
__global__ void RT_TRACE_BOUNDS traceQueueCompactKernel(const __grid_constant__ SceneView s, const float4* __restrict__ rays,
                                                        const uint32_t* __restrict__ countPtr, uint32_t countHost, uint32_t* __restrict__ head,
                                                        RptIntersection* __restrict__ hits, uint8_t* __restrict__ occluded) {
	__shared__ WarpScratch scratch[TraceBlock / 32];
	WarpScratch& ws = scratch[threadIdx.x >> 5];
	const uint32_t n = countPtr ? *countPtr : countHost;
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t rayIdx = NoRay;
	bool dry = false;
	TravRay r = makeTravRay(f3(0.0f), 0.0f, f3(1.0f));
	float bestT = 0.0f, bestU = 0.0f, bestV = 0.0f;
	uint32_t bestKey = 0, bestFlat = NoCand, bestInst = InvalidHitIndex, bestTri = 0;
#ifdef RT_SMEM_STACK
	__shared__ uint2 stackShared[RT_SMEM_STACK][TraceBlock];
	struct HybridStack {
		uint2 (*sh)[TraceBlock];
		uint2 spill[TraversalStackSize - RT_SMEM_STACK];
		struct Ref {
			HybridStack& st; int i;
			RT_DEV void operator=(uint2 v) { if (i < RT_SMEM_STACK) st.sh[i][threadIdx.x] = v; else st.spill[i - RT_SMEM_STACK] = v; }
			RT_DEV operator uint2() const { return i < RT_SMEM_STACK ? st.sh[i][threadIdx.x] : st.spill[i - RT_SMEM_STACK]; }
		};
		RT_DEV Ref operator[](int i) { return Ref{ *this, i }; }
	} stack;
	stack.sh = stackShared;
#else
	uint2 stack[TraversalStackSize];
#endif
	int sp = 0;
	uint2 ngroup = make_uint2(0u, 0u);
	uint32_t nodeVisits = 0, triTests = 0, visitsAtFetch = 0;
	if (lane == 0) ws.count = 0;
	__syncwarp();

	for (;;) {
		// ---- dynamic fetch (as in traceQueueKernel) ---------------------------------------------------------------
		const unsigned idleMask = __ballot_sync(FullWarp, rayIdx == NoRay);
		if (!dry && __popc(idleMask) >= FetchThreshold) {
			const int leader = __ffs(int(idleMask)) - 1;
			const uint32_t want = uint32_t(__popc(idleMask));
			uint32_t base = 0;
			if (int(lane) == leader) base = atomicAdd(head, want);
			base = __shfl_sync(FullWarp, base, leader);
			if (base + want >= n) dry = true;
			if (rayIdx == NoRay) {
				const uint32_t idx = base + uint32_t(__popc(idleMask & ((1u << lane) - 1u)));
				if (idx < n) {
					const float4 a = __ldcs(rays + 2 * size_t(idx)), b = __ldcs(rays + 2 * size_t(idx) + 1);
					if (rayIsDegenerate(f3(a), a.w, f3(b), b.w)) {
						if (MODE == TraceAny) occluded[idx] = 0;
						else { RptIntersection o; o.bary[0] = 0.f; o.bary[1] = 0.f; o.instanceIdx = InvalidHitIndex; o.triangleIdx = 0; hits[idx] = o; }
					}
					else {
						if (s.counters != nullptr) atomicAdd(&s.counters[MODE == TraceAny ? 1 : 0], 1ull);
						rayIdx = idx;
						r = makeTravRay(f3(a), a.w, f3(b));
						ws.rayA[lane] = a; ws.rayB[lane] = b;
						bestT = b.w; bestU = 0.0f; bestV = 0.0f; bestFlat = NoCand; bestInst = InvalidHitIndex; bestTri = 0;
						bestKey = MODE == TraceAny ? 0u : orderKey(b.w);
						ws.keyT[lane] = bestKey;
						ws.candFlat[lane] = NoCand;
						sp = 0;
						ngroup = make_uint2(0u, 0x80000000u);
						visitsAtFetch = nodeVisits;
					}
				}
			}
			__syncwarp();
		}
		if (__all_sync(FullWarp, rayIdx == NoRay)) {
			if (dry) break;
			continue;
		}

		// ---- node step of every lane that holds a ray -----------------------------------------------------------------
		uint32_t triBase = 0, triHits = 0;
		if (rayIdx != NoRay && ngroup.y > 0x00ffffffu) {
			nodeStep(s, r, bestT, ngroup, stack, sp, triBase, triHits);
			nodeVisits++;
		}

		// ---- the round's triangles, tested by full warps ---------------------------------------------------------------
		if (__any_sync(FullWarp, triHits != 0u)) {
			if (triHits) {
				uint32_t p = atomicAdd(&ws.count, uint32_t(__popc(triHits)));
				ws.triBase[lane] = triBase;
				uint32_t m = triHits;
				do {
					const uint32_t i = uint32_t(__ffs(int(m))) - 1u;
					m &= m - 1u;
					ws.list[p++] = uint16_t((lane << 5) | i);
				} while (m);
			}
			__syncwarp();
			const uint32_t total = ws.count;
			for (uint32_t c = 0; c < total; c += 32u) {
				const uint32_t j = c + lane;
				bool ok = false;
				uint32_t owner = 0, key = 0;
				TriHit h;
				if (j < total) {
					const uint32_t e = ws.list[j];
					owner = e >> 5;
					const float4 a = ws.rayA[owner], b = ws.rayB[owner];
					triTests++;
					ok = triTestRay(s, f3(a), f3(b), a.w, ws.triBase[owner] + (e & 31u), b.w, h);
				}
				if (MODE == TraceAny) {
					if (ok) ws.keyT[owner] = 1u;
					__syncwarp();
				}
				else {
					if (ok) { key = orderKey(h.t); atomicMin(&ws.keyT[owner], key); }
					__syncwarp();
					const bool cand = ok && ws.keyT[owner] == key;
					if (cand) atomicMin(&ws.candFlat[owner], h.flat);
					__syncwarp();
					if (cand && ws.candFlat[owner] == h.flat)
						ws.win[owner] = make_uint4(__float_as_uint(h.u), __float_as_uint(h.v), h.instanceIdx, h.triangleIdx);
					__syncwarp();
					const uint32_t cf = ws.candFlat[lane];
					if (rayIdx != NoRay && cf != NoCand) {
						const uint32_t kt = ws.keyT[lane];
						if (kt < bestKey || cf < bestFlat) {
							const uint4 w = ws.win[lane];
							bestKey = kt; bestFlat = cf; bestT = orderKeyInv(kt);
							bestU = __uint_as_float(w.x); bestV = __uint_as_float(w.y); bestInst = w.z; bestTri = w.w;
						}
						ws.candFlat[lane] = NoCand;
					}
					__syncwarp();
				}
			}
			if (lane == 0) ws.count = 0;
			__syncwarp();
		}

		if (rayIdx != NoRay) {
			bool finished = false;
			if (MODE == TraceAny && ws.keyT[lane] != 0u) finished = true;
			else if (ngroup.y <= 0x00ffffffu) {
				if (sp == 0) finished = true;
				else ngroup = stack[--sp];
			}
			if (finished) {
				if (s.counters != nullptr) atomicMax(&s.counters[7], (unsigned long long)(nodeVisits - visitsAtFetch));
				if (MODE == TraceAny) occluded[rayIdx] = ws.keyT[lane] != 0u ? 1 : 0;
				else {
					RptIntersection o;
					o.bary[0] = bestU; o.bary[1] = bestV; o.instanceIdx = bestInst; o.triangleIdx = bestTri;
					hits[rayIdx] = o;
				}
				rayIdx = NoRay;
			}
		}
	}
	if (s.counters != nullptr) {
		atomicAdd(&s.counters[2], (unsigned long long)nodeVisits);
		if (MODE == TraceAny) { atomicAdd(&s.counters[5], (unsigned long long)nodeVisits); atomicAdd(&s.counters[6], (unsigned long long)triTests); }
		atomicAdd(&s.counters[3], (unsigned long long)triTests);
	}
}

template <int MODE>
void launchQueue(const SceneView& s, const float4* rays, const uint32_t* countPtr, uint32_t countHost, uint32_t* head,
                 RptIntersection* hits, uint8_t* occluded, cudaStream_t st) {
#ifdef RT_TRI_COMPACT
	static const int blocks = persistentBlocks(reinterpret_cast<const void*>(traceQueueCompactKernel<MODE>), TraceBlock);
	traceQueueCompactKernel<MODE><<<blocks, TraceBlock, 0, st>>>(s, rays, countPtr, countHost, head, hits, occluded);
#else
	static const int blocks = persistentBlocks(reinterpret_cast<const void*>(traceQueueKernel<MODE>), TraceBlock);
	traceQueueKernel<MODE><<<blocks, TraceBlock, 0, st>>>(s, rays, countPtr, countHost, head, hits, occluded);
#endif
}

} // namespace

void launchTraceQueueClosest(const SceneView& s, const float4* rays, const uint32_t* countPtr, uint32_t countHost, uint32_t* head,
                             RptIntersection* hits, cudaStream_t st) {
	launchQueue<TraceClosest>(s, rays, countPtr, countHost, head, hits, nullptr, st);
}
void launchTraceQueueAny(const SceneView& s, const float4* rays, const uint32_t* countPtr, uint32_t countHost, uint32_t* head,
                         uint8_t* occluded, cudaStream_t st) {
	launchQueue<TraceAny>(s, rays, countPtr, countHost, head, nullptr, occluded, st);
}

} // namespace rt
