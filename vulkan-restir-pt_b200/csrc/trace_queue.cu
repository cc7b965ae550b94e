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
			finished = triLoop<MODE>(s, r, triBase, triHits, tmaxOrig, res, triTests);
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


// ---- variant: warp-local ray pool in shared memory (RT_RAY_POOL) ----------------------------------------------------
// In the kernel above a fetch is a dependent chain — queue atomic (L2 round trip) -> ray record (HBM, read once) -> three IEEE
// reciprocals — that the whole warp sits through whenever 8 lanes are idle: 30 % of the kernel's stall samples
// (profiles/r2_03_*), and the node step runs at 28 of 32 lanes while lanes wait for the threshold.  Here each warp owns a
// two-deep pool of 32-ray batches in shared memory: the queue atomic for batch k+2 is issued one loop iteration before its
// result is used, the batch's records travel HBM -> shared memory as asynchronous copies (cp.async, bypassing L1: they are
// read once) while the warp traverses, and when a batch lands all 32 lanes prepare one ray each (degenerate test, reciprocal
// direction, octant) at full width.  A lane whose ray ends takes the next prepared ray from shared memory in the same
// iteration (three 16-byte shared loads), so lanes are never parked waiting for a threshold.
// Per-ray arithmetic and results are those of traceQueueKernel.
constexpr uint32_t PoolBatch = 32;
#ifndef RT_POOL_THRESHOLD
#define RT_POOL_THRESHOLD 1   // idle lanes per warp that trigger a refill from the pool
#endif
struct __align__(16) RayPool {
	float4 a[2][PoolBatch];   // origin, tmin
	float4 b[2][PoolBatch];   // direction, tmax
	float4 c[PoolBatch];      // prepared: reciprocal direction, w = octinv (bits) or 0xffffffff for a degenerate ray (already answered)
};

RT_DEV void cpAsync16(void* sharedDst, const void* globalSrc) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(uint32_t(__cvta_generic_to_shared(sharedDst))), "l"(globalSrc) : "memory");
}
RT_DEV void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
RT_DEV void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int MODE>
__global__ void RT_TRACE_BOUNDS traceQueuePoolKernel(const __grid_constant__ SceneView s, const float4* __restrict__ rays,
                                                     const uint32_t* __restrict__ countPtr, uint32_t countHost, uint32_t* __restrict__ head,
                                                     RptIntersection* __restrict__ hits, uint8_t* __restrict__ occluded) {
	__shared__ RayPool pools[TraceBlock / 32];
	RayPool& pool = pools[threadIdx.x >> 5];
	const uint32_t n = countPtr ? *countPtr : countHost;
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t rayIdx = NoRay;
	TravRay r = makeTravRay(f3(0.0f), 0.0f, f3(1.0f));
	TravResult res;
	res.init(0.0f);
	float tmaxOrig = 0.0f;
	uint2 stack[TraversalStackSize];
	int sp = 0;
	uint2 ngroup = make_uint2(0u, 0u);
	uint32_t nodeVisits = 0, triTests = 0, visitsAtFetch = 0;

	// warp-uniform pool state
	uint32_t curBase = 0, curCount = 0, curPos = 0, buf = 0;   // batch being handed out: rays curBase + [curPos, curCount) of pool.a/b[buf]
	uint32_t nxtBase = 0, nxtCount = 0;
	int nxtState = 0;          // 0: nothing in flight, 1: queue atomic issued (result in lane 0's ticket), 2: copies issued into pool.a/b[buf ^ 1]
	uint32_t ticket = 0;       // lane 0: result of the queue atomic
	bool dry = false;          // the queue has been handed out completely (no further atomics)

	for (;;) {
		const unsigned idleMask = __ballot_sync(FullWarp, rayIdx == NoRay);
		// ---- the next batch: keep the pipeline one step ahead of its use ----------------------------------------------------
		if (nxtState == 1) {
			const uint32_t base = __shfl_sync(FullWarp, ticket, 0);
			nxtBase = base;
			nxtCount = base < n ? min(PoolBatch, n - base) : 0u;
			if (base + PoolBatch >= n) dry = true;
			if (lane < nxtCount) {
				cpAsync16(&pool.a[buf ^ 1u][lane], rays + 2 * size_t(base + lane));
				cpAsync16(&pool.b[buf ^ 1u][lane], rays + 2 * size_t(base + lane) + 1);
			}
			cpAsyncCommit();
			nxtState = 2;
		}
		else if (nxtState == 0 && !dry) {
			if (lane == 0) ticket = atomicAdd(head, PoolBatch);
			nxtState = 1;
		}
		// ---- the current batch is used up: take over the next one and prepare its rays with all 32 lanes ----------------------
		if (curPos == curCount && nxtState == 2 && idleMask != 0u) {
			cpAsyncWaitAll();
			__syncwarp();
			buf ^= 1u;
			curBase = nxtBase; curCount = nxtCount; curPos = 0;
			nxtState = 0;
			if (lane < curCount) {
				const float4 a = pool.a[buf][lane], b = pool.b[buf][lane];
				float4 c;
				if (rayIsDegenerate(f3(a), a.w, f3(b), b.w)) {
					const uint32_t idx = curBase + lane;
					if (MODE == TraceAny) occluded[idx] = 0;
					else { RptIntersection o; o.bary[0] = 0.f; o.bary[1] = 0.f; o.instanceIdx = InvalidHitIndex; o.triangleIdx = 0; hits[idx] = o; }
					c = make_float4(0.f, 0.f, 0.f, __uint_as_float(0xffffffffu));
				}
				else {
					if (s.counters != nullptr) atomicAdd(&s.counters[MODE == TraceAny ? 1 : 0], 1ull);
					const TravRay t = makeTravRay(f3(a), a.w, f3(b));
					c = make_float4(t.idx, t.idy, t.idz, __uint_as_float(t.octinv));
				}
				pool.c[lane] = c;
			}
			__syncwarp();
		}
		// ---- idle lanes take prepared rays ------------------------------------------------------------------------------------
		if (curPos < curCount && __popc(idleMask) >= RT_POOL_THRESHOLD) {
			const uint32_t slot = curPos + uint32_t(__popc(idleMask & ((1u << lane) - 1u)));
			if (rayIdx == NoRay && slot < curCount) {
				const float4 c = pool.c[slot];
				const uint32_t oct = __float_as_uint(c.w);
				if (oct != 0xffffffffu) {
					const float4 a = pool.a[buf][slot], b = pool.b[buf][slot];
					rayIdx = curBase + slot;
					r.o = f3(a); r.d = f3(b); r.tmin = a.w;
					r.idx = c.x; r.idy = c.y; r.idz = c.z; r.octinv = oct;
					tmaxOrig = b.w;
					res.init(b.w);
					sp = 0;
					ngroup = make_uint2(0u, 0x80000000u);
					visitsAtFetch = nodeVisits;
				}
			}
			curPos = min(curCount, curPos + uint32_t(__popc(idleMask)));
		}
		if (__all_sync(FullWarp, rayIdx == NoRay)) {
			if (curPos == curCount && nxtState == 0 && dry) break;
			continue;
		}

		// ---- one traversal step of every lane that holds a ray (as in traceQueueKernel) -----------------------------------------
		if (rayIdx != NoRay) {
			bool finished = false;
			uint32_t triBase = 0, triHits = 0;
			if (ngroup.y > 0x00ffffffu) {
				nodeStep(s, r, res.bestT, ngroup, stack, sp, triBase, triHits);
				nodeVisits++;
			}
			finished = triLoop<MODE>(s, r, triBase, triHits, tmaxOrig, res, triTests);
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
		atomicAdd(&s.counters[2], (unsigned long long)nodeVisits);
		if (MODE == TraceAny) { atomicAdd(&s.counters[5], (unsigned long long)nodeVisits); atomicAdd(&s.counters[6], (unsigned long long)triTests); }
		atomicAdd(&s.counters[3], (unsigned long long)triTests);
	}
}

template <int MODE>
void launchQueue(const SceneView& s, const float4* rays, const uint32_t* countPtr, uint32_t countHost, uint32_t* head,
                 RptIntersection* hits, uint8_t* occluded, cudaStream_t st) {
#ifdef RT_RAY_POOL
	static const int blocks = persistentBlocks(reinterpret_cast<const void*>(traceQueuePoolKernel<MODE>), TraceBlock);
	traceQueuePoolKernel<MODE><<<blocks, TraceBlock, 0, st>>>(s, rays, countPtr, countHost, head, hits, occluded);
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
