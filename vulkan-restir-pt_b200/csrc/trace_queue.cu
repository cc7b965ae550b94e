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
			LeafHits leaves{ 0u, 0u, 0u };
			if (ngroup.y > 0x00ffffffu) {
				nodeStep(s.nodes, r, res.bestT, ngroup, stack, sp, leaves);
				nodeVisits++;
			}
			finished = triLoop<MODE>(s, r, leaves, tmaxOrig, res, triTests);
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


// ---- two-level scenes ---------------------------------------------------------------------------------------------------------
// The same kernel over a TLAS of instances and object-space BLASes (bvh_traverse.cuh, traceRayTwoLevel): every lane is a small
// state machine — level 0 walks the TLAS with the world-space ray, a hit instance takes the lane to level 1 (ray transformed,
// BLAS root on top of the same stack, the TLAS group saved below it), and when the BLAS is exhausted the lane drops back and
// re-reads its world-space ray from the queue.
template <int MODE>
__global__ void RT_TRACE_BOUNDS traceQueueTwoLevelKernel(const __grid_constant__ SceneView s, const float4* __restrict__ rays,
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
	int sp = 0, baseSp = 0;
	uint2 ngroup = make_uint2(0u, 0u);
	LeafHits inst{ 0u, 0u, 0u };          // level 0: hit instances of the last TLAS node, not yet entered
	bool inBlas = false;
	uint32_t customIndex = 0, flatBase = 0;
	uint32_t nodeVisits = 0, triTests = 0, visitsAtFetch = 0;

	for (;;) {
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
						tmaxOrig = b.w;
						res.init(b.w);
						sp = 0; inBlas = false; inst.bits = 0;
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

		if (rayIdx != NoRay) {
			bool finished = false;
			// ONE node step per iteration for the lanes of both levels (the node format is the same; only the array differs), so a
			// warp with lanes in the TLAS and lanes in a BLAS does not run the box tests twice
			LeafHits leaves{ 0u, 0u, 0u };
			const bool step = ngroup.y > 0x00ffffffu && (inBlas || inst.bits == 0u);
			if (step) {
				nodeStep(inBlas ? s.nodes : s.tlasNodes, r, res.bestT, ngroup, stack, sp, leaves);
				nodeVisits++;
			}
			if (!inBlas) {
				if (step) inst = leaves;
				if (inst.bits) {   // enter the next hit instance
					const uint32_t one = 1u << (31u - uint32_t(__clz(int(inst.bits))));
					inst.bits ^= one;
					const InstanceRecord rec = loadInstanceRecord(s, inst.triBase + uint32_t(__popc(inst.valid & (one - 1u))));
					const ObjectRay ob = toObjectSpace(rec, r.o, r.d);
					if (rec.rootNode != 0xffffffffu && !rayIsDegenerate(ob.o, r.tmin, ob.d, tmaxOrig) && sp < TraversalStackSize) {
						stack[sp++] = ngroup;   // the TLAS group, empty or not: popped when the BLAS is done
						baseSp = sp;
						r = makeTravRay(ob.o, r.tmin, ob.d);
						ngroup = make_uint2(rec.rootNode, 0x80000000u);
						customIndex = rec.customIndex; flatBase = rec.flatBase;
						inBlas = true;
					}
				}
				else if (ngroup.y <= 0x00ffffffu) {
					if (sp == 0) finished = true;
					else ngroup = stack[--sp];
				}
			}
			else {
				finished = triLoop<MODE, true>(s, r, leaves, tmaxOrig, res, triTests, customIndex, flatBase);
				if (!finished && ngroup.y <= 0x00ffffffu) {
					if (sp > baseSp) ngroup = stack[--sp];
					else {   // back to the TLAS with the world-space ray
						ngroup = stack[--sp];
						const float4 a = __ldg(rays + 2 * size_t(rayIdx)), b = __ldg(rays + 2 * size_t(rayIdx) + 1);
						r = makeTravRay(f3(a), a.w, f3(b));
						inBlas = false;
					}
				}
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
		atomicAdd(&s.counters[2], (unsigned long long)nodeVisits);
		if (MODE == TraceAny) { atomicAdd(&s.counters[5], (unsigned long long)nodeVisits); atomicAdd(&s.counters[6], (unsigned long long)triTests); }
		atomicAdd(&s.counters[3], (unsigned long long)triTests);
	}
}

// Measured and rejected variants of this kernel (profiles/README.md, r1_05 / r1_12 / r1_16 / r2_03 / r2_04): postponing triangle
// tests until enough lanes have one; warp-cooperative triangle rounds through shared memory (64-bit and 32-bit atomics);
// a warp-local pool of prefetched, prepared rays in shared memory (cp.async, refill at any number of idle lanes); a
// software-pipelined triangle loop; register caps; prefetching.  All of them raise lane utilisation or hide latency and all
// of them are slower: the kernel is bound by instruction issue (ALU pipe), so only fewer instructions per node help
// (bvh_traverse.cuh nodeStep).

template <int MODE>
void launchQueue(const SceneView& s, const float4* rays, const uint32_t* countPtr, uint32_t countHost, uint32_t* head,
                 RptIntersection* hits, uint8_t* occluded, cudaStream_t st) {
	if (s.tlasNodes != nullptr) {
		static const int blocks2 = persistentBlocks(reinterpret_cast<const void*>(traceQueueTwoLevelKernel<MODE>), TraceBlock);
		traceQueueTwoLevelKernel<MODE><<<blocks2, TraceBlock, 0, st>>>(s, rays, countPtr, countHost, head, hits, occluded);
		return;
	}
	static const int blocks = persistentBlocks(reinterpret_cast<const void*>(traceQueueKernel<MODE>), TraceBlock);
	traceQueueKernel<MODE><<<blocks, TraceBlock, 0, st>>>(s, rays, countPtr, countHost, head, hits, occluded);
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
