// GPU construction of the acceleration structure.  Replaces the driver-side BLAS / TLAS builds of the reference
// (zvk/core/AccelerationStructure.cpp:7-136, call sites src/Scene.cpp:448-547) with:
//   1. flatten: every instance's triangles (and the light triangles, custom index 0) to world space — the
//      reference never shares geometry between instances (src/Resource.cpp:183-184) and a B200 has 180 GB, so a
//      single-level structure over world-space triangles replaces BLAS+TLAS;
//   2. 63-bit Morton sort (CUB radix sort);
//   3. PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) binary BVH;
//   4. greedy surface-area collapse to 8-wide nodes, octant-ordered child slots, 80-byte compressed nodes
//      (Ylitie, Karras, Laine 2017), triangles re-emitted in leaf order.
#include <cstdlib>
#include <cub/cub.cuh>
#include <cfloat>
#include "bvh_build.h"
#include "rt_math.cuh"

namespace rt {

namespace {

#define CK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)

struct Box { float3 lo, hi; };

RT_DEV float boxArea(float3 lo, float3 hi) {
	float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
	return dx * dy + dy * dz + dz * dx;
}

// order-preserving float <-> uint for atomicMin/Max
RT_DEV uint32_t floatFlip(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ inline float floatUnflip(uint32_t u) {
	uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
	float f;
#ifdef __CUDA_ARCH__
	f = __uint_as_float(v);
#else
	memcpy(&f, &v, 4);
#endif
	return f;
}

// ---- 1. flatten ------------------------------------------------------------------------------------------
// flat triangle i: i < numLights -> light triangle i (instance 0); else object instance k with
// triOffsets[k] <= i < triOffsets[k+1] (triOffsets[0] == numLights)
__global__ void flattenKernel(BuildInputs in, TriRecord* __restrict__ trisFlat, float4* __restrict__ leafLo,
                              float4* __restrict__ leafHi, uint32_t* __restrict__ bounds) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	float3 lo = f3(FLT_MAX), hi = f3(-FLT_MAX);
	if (i < in.numTris) {
		float3 w0, w1, w2;
		uint32_t inst, tri;
		if (i < in.numLights) {
			const RptTriangleLight& L = in.lights[i];
			w0 = f3(L.v0); w1 = f3(L.v1); w2 = f3(L.v2);
			inst = 0; tri = i;
		}
		else {
			uint32_t a = 0, b = in.numInstances;   // last k with triOffsets[k] <= i
			while (b - a > 1) { uint32_t m = (a + b) >> 1; if (in.triOffsets[m] <= i) a = m; else b = m; }
			const RptObjectInstance& I = in.instances[a];
			tri = i - in.triOffsets[a];
			inst = a + 1;
			const uint32_t* idx = in.indices + I.indexOffset + tri * 3;
			w0 = xformPoint(I.transform, f3(in.vertices[idx[0]].pos));
			w1 = xformPoint(I.transform, f3(in.vertices[idx[1]].pos));
			w2 = xformPoint(I.transform, f3(in.vertices[idx[2]].pos));
		}
		const float3 e1 = w1 - w0, e2 = w2 - w0;
		TriRecord r;
		r.t0 = make_float4(w0.x, w0.y, w0.z, __uint_as_float(inst));
		r.t1 = make_float4(e1.x, e1.y, e1.z, __uint_as_float(tri));
		r.t2 = make_float4(e2.x, e2.y, e2.z, __uint_as_float(i));
		trisFlat[i] = r;
		// bounds of what the intersector sees (v0, v0+e1, v0+e2), padded for the BaryEps tolerance and rounding
		const float3 p1 = w0 + e1, p2 = w0 + e2;
		lo = make_float3(fminf(w0.x, fminf(p1.x, p2.x)), fminf(w0.y, fminf(p1.y, p2.y)), fminf(w0.z, fminf(p1.z, p2.z)));
		hi = make_float3(fmaxf(w0.x, fmaxf(p1.x, p2.x)), fmaxf(w0.y, fmaxf(p1.y, p2.y)), fmaxf(w0.z, fmaxf(p1.z, p2.z)));
		const float ext = fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z));
		const float3 pad = make_float3(
			4.0f * BaryEps * ext + 1e-5f * (fmaxf(fabsf(lo.x), fabsf(hi.x)) + 1.0f),
			4.0f * BaryEps * ext + 1e-5f * (fmaxf(fabsf(lo.y), fabsf(hi.y)) + 1.0f),
			4.0f * BaryEps * ext + 1e-5f * (fmaxf(fabsf(lo.z), fabsf(hi.z)) + 1.0f));
		lo = lo - pad; hi = hi + pad;
		if (!(lo.x <= hi.x) || !isfinite(lo.x + lo.y + lo.z + hi.x + hi.y + hi.z)) {   // degenerate / NaN input
			lo = f3(0.0f); hi = f3(0.0f);
		}
		leafLo[i] = make_float4(lo.x, lo.y, lo.z, __uint_as_float(i));
		leafHi[i] = make_float4(hi.x, hi.y, hi.z, __uint_as_float(0xffffffffu));
	}
	// scene bounds: warp reduce, then atomics
	for (int o = 16; o > 0; o >>= 1) {
		lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o)); lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
		hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
	}
	if ((threadIdx.x & 31) == 0 && lo.x <= hi.x) {
		atomicMin(&bounds[0], floatFlip(lo.x)); atomicMin(&bounds[1], floatFlip(lo.y)); atomicMin(&bounds[2], floatFlip(lo.z));
		atomicMax(&bounds[3], floatFlip(hi.x)); atomicMax(&bounds[4], floatFlip(hi.y)); atomicMax(&bounds[5], floatFlip(hi.z));
	}
}

// ---- 2. Morton codes --------------------------------------------------------------------------------------
RT_DEV uint64_t expandBits21(uint64_t v) {
	v &= 0x1fffffull;
	v = (v | (v << 32)) & 0x1f00000000ffffull;
	v = (v | (v << 16)) & 0x1f0000ff0000ffull;
	v = (v | (v << 8)) & 0x100f00f00f00f00full;
	v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
	v = (v | (v << 2)) & 0x1249249249249249ull;
	return v;
}

__global__ void mortonKernel(uint32_t n, const float4* __restrict__ leafLo, const float4* __restrict__ leafHi,
                             const uint32_t* __restrict__ bounds, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float3 slo = make_float3(floatUnflip(bounds[0]), floatUnflip(bounds[1]), floatUnflip(bounds[2]));
	const float3 shi = make_float3(floatUnflip(bounds[3]), floatUnflip(bounds[4]), floatUnflip(bounds[5]));
	const float4 lo = leafLo[i], hi = leafHi[i];
	const float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
	const float sx = shi.x > slo.x ? 2097151.0f / (shi.x - slo.x) : 0.0f;
	const float sy = shi.y > slo.y ? 2097151.0f / (shi.y - slo.y) : 0.0f;
	const float sz = shi.z > slo.z ? 2097151.0f / (shi.z - slo.z) : 0.0f;
	const uint64_t qx = uint64_t(fminf(fmaxf((cx - slo.x) * sx, 0.0f), 2097151.0f));
	const uint64_t qy = uint64_t(fminf(fmaxf((cy - slo.y) * sy, 0.0f), 2097151.0f));
	const uint64_t qz = uint64_t(fminf(fmaxf((cz - slo.z) * sz, 0.0f), 2097151.0f));
	keys[i] = (expandBits21(qx) << 2) | (expandBits21(qy) << 1) | expandBits21(qz);
	vals[i] = i;
}

// node arrays: [0, N) leaves in Morton order, [N, 2N-1) inner nodes in creation order.
//   nodeLo[i] = {lo, left}, nodeHi[i] = {hi, right}; leaf: left = flat triangle index, right = 0xffffffff
__global__ void initLeavesKernel(uint32_t n, const uint32_t* __restrict__ sortedVals, const float4* __restrict__ leafLo,
                                 const float4* __restrict__ leafHi, float4* __restrict__ nodeLo, float4* __restrict__ nodeHi,
                                 uint32_t* __restrict__ nodeCount, uint32_t* __restrict__ clusters) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t src = sortedVals[i];
	nodeLo[i] = leafLo[src];
	nodeHi[i] = leafHi[src];
	nodeCount[i] = 1;
	clusters[i] = i;
}

// ---- 3. PLOC ------------------------------------------------------------------------------------------------
constexpr int PlocRadius = 16;
constexpr int PlocBlock = 256;

__global__ void __launch_bounds__(PlocBlock) plocNearestKernel(uint32_t n, const uint32_t* __restrict__ clusters,
                                                                const float4* __restrict__ nodeLo, const float4* __restrict__ nodeHi,
                                                                uint32_t* __restrict__ nearest) {
	__shared__ float3 sLo[PlocBlock + 2 * PlocRadius];
	__shared__ float3 sHi[PlocBlock + 2 * PlocRadius];
	const int base = int(blockIdx.x) * PlocBlock - PlocRadius;
	for (int k = threadIdx.x; k < PlocBlock + 2 * PlocRadius; k += PlocBlock) {
		const int g = base + k;
		if (g >= 0 && g < int(n)) {
			const uint32_t c = clusters[g];
			sLo[k] = f3(nodeLo[c]); sHi[k] = f3(nodeHi[c]);
		}
	}
	__syncthreads();
	const int i = int(blockIdx.x) * PlocBlock + int(threadIdx.x);
	if (i >= int(n)) return;
	const int li = int(threadIdx.x) + PlocRadius;
	const float3 lo = sLo[li], hi = sHi[li];
	float best = FLT_MAX;
	int bestJ = -1;
	for (int dj = -PlocRadius; dj <= PlocRadius; dj++) {
		const int j = i + dj;
		if (dj == 0 || j < 0 || j >= int(n)) continue;
		const float3 a = sLo[li + dj], b = sHi[li + dj];
		const float area = boxArea(make_float3(fminf(lo.x, a.x), fminf(lo.y, a.y), fminf(lo.z, a.z)),
		                           make_float3(fmaxf(hi.x, b.x), fmaxf(hi.y, b.y), fmaxf(hi.z, b.z)));
		if (area < best) { best = area; bestJ = j; }   // ascending j: ties keep the lower index
	}
	nearest[i] = uint32_t(bestJ);
}

// flags: low 32 bits = cluster survives, high 32 bits = cluster creates a new node
__global__ void plocFlagKernel(uint32_t n, const uint32_t* __restrict__ nearest, uint64_t* __restrict__ flags) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t j = nearest[i];
	const bool mutual = j < n && nearest[j] == i;
	uint64_t f = 1ull;
	if (mutual) f = (i < j) ? (1ull | (1ull << 32)) : 0ull;
	flags[i] = f;
}

__global__ void plocMergeKernel(uint32_t n, uint32_t numLeaves, uint32_t nodesCreated, const uint32_t* __restrict__ clusters,
                                const uint32_t* __restrict__ nearest, const uint64_t* __restrict__ flags,
                                const uint64_t* __restrict__ prefix, float4* __restrict__ nodeLo, float4* __restrict__ nodeHi,
                                uint32_t* __restrict__ nodeCount, uint32_t* __restrict__ clustersOut) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t f = flags[i];
	if ((f & 1ull) == 0) return;
	const uint64_t p = prefix[i];
	const uint32_t outPos = uint32_t(p & 0xffffffffull);
	if (f >> 32) {
		const uint32_t a = clusters[i], b = clusters[nearest[i]];
		const uint32_t id = numLeaves + nodesCreated + uint32_t(p >> 32);
		const float4 alo = nodeLo[a], ahi = nodeHi[a], blo = nodeLo[b], bhi = nodeHi[b];
		nodeLo[id] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), __uint_as_float(a));
		nodeHi[id] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), __uint_as_float(b));
		nodeCount[id] = nodeCount[a] + nodeCount[b];
		clustersOut[outPos] = id;
	}
	else {
		clustersOut[outPos] = clusters[i];
	}
}

// ---- 4. collapse to 8-wide + compression --------------------------------------------------------------------
struct CollapseCounters { uint32_t numNodes, numTris, nextCount, pad; };

RT_DEV uint32_t exponentFor(float extent) {
	// smallest e with 2^(e-127) * 255 >= extent (with a safety factor); 0 encodes scale 0 for flat axes
	if (!(extent > 0.0f)) return 0;
	float q = extent * (1.0f / 255.0f) * 1.000002f;
	uint32_t b = __float_as_uint(q);
	uint32_t e = (b + 0x7fffffu) >> 23;
	return min(max(e, 1u), 239u);   // (+15 must fit the node's exponent byte; 2^112 is beyond any scene)
}

__global__ void __launch_bounds__(64) collapseKernel(uint32_t numTasks, const uint2* __restrict__ tasks, uint2* __restrict__ nextTasks,
                                                      uint32_t numLeaves, const float4* __restrict__ nodeLo, const float4* __restrict__ nodeHi,
                                                      const uint32_t* __restrict__ nodeCount, const TriRecord* __restrict__ trisFlat,
                                                      WideNode* __restrict__ wide, TriRecord* __restrict__ trisOut,
                                                      CollapseCounters* __restrict__ counters, uint32_t leafMax) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= numTasks) return;
	const uint2 task = tasks[t];
	const uint32_t self = task.x;

	uint32_t child[8];
	float area[8];
	int cnt = 0;
	const float4 selfLo = nodeLo[self], selfHi = nodeHi[self];
	if (self < numLeaves) {   // single-triangle scene: the root is a leaf
		child[cnt++] = self;
	}
	else {
		child[0] = __float_as_uint(selfLo.w); child[1] = __float_as_uint(selfHi.w);
		cnt = 2;
	}
	for (int k = 0; k < cnt; k++) area[k] = boxArea(f3(nodeLo[child[k]]), f3(nodeHi[child[k]]));
	// greedy: keep opening the inner child with the largest surface area until 8 slots are used
	while (cnt < 8) {
		int pick = -1; float bestA = -1.0f;
		for (int k = 0; k < cnt; k++) if (child[k] >= numLeaves && area[k] > bestA) { bestA = area[k]; pick = k; }
		if (pick < 0) break;
		const uint32_t c = child[pick];
		const uint32_t l = __float_as_uint(nodeLo[c].w), r = __float_as_uint(nodeHi[c].w);
		child[pick] = l; area[pick] = boxArea(f3(nodeLo[l]), f3(nodeHi[l]));
		child[cnt] = r; area[cnt] = boxArea(f3(nodeLo[r]), f3(nodeHi[r]));
		cnt++;
	}

	// slot assignment: child whose centroid lies towards (+x,+y,+z) prefers the slot with those bits set, so
	// that visiting slots in (slot ^ octant) order is roughly front to back for any ray direction
	const float3 centre = make_float3(0.5f * (selfLo.x + selfHi.x), 0.5f * (selfLo.y + selfHi.y), 0.5f * (selfLo.z + selfHi.z));
	float3 off[8];
	for (int k = 0; k < cnt; k++) {
		const float4 a = nodeLo[child[k]], b = nodeHi[child[k]];
		off[k] = make_float3(0.5f * (a.x + b.x) - centre.x, 0.5f * (a.y + b.y) - centre.y, 0.5f * (a.z + b.z) - centre.z);
	}
	int slotOf[8];
	uint32_t slotUsed = 0, childDone = 0;
	for (int round = 0; round < cnt; round++) {
		float bestC = -FLT_MAX; int bk = 0, bs = 0;
		for (int k = 0; k < cnt; k++) {
			if (childDone & (1u << k)) continue;
			for (int sl = 0; sl < 8; sl++) {
				if (slotUsed & (1u << sl)) continue;
				const float c = ((sl & 1) ? off[k].x : -off[k].x) + ((sl & 2) ? off[k].y : -off[k].y) + ((sl & 4) ? off[k].z : -off[k].z);
				if (c > bestC) { bestC = c; bk = k; bs = sl; }
			}
		}
		slotOf[bk] = bs; slotUsed |= 1u << bs; childDone |= 1u << bk;
	}
	uint32_t slotChild[8];
	for (int sl = 0; sl < 8; sl++) slotChild[sl] = 0xffffffffu;
	for (int k = 0; k < cnt; k++) slotChild[slotOf[k]] = child[k];

	// classify: subtrees of <= 3 triangles become leaves of the wide node
	uint32_t imask = 0, numInner = 0, numTri = 0;
	uint32_t triOffset[8], triCount[8];
	for (int sl = 0; sl < 8; sl++) {
		triOffset[sl] = 0; triCount[sl] = 0;
		const uint32_t c = slotChild[sl];
		if (c == 0xffffffffu) continue;
		const uint32_t n = nodeCount[c];
		if (n <= leafMax) { triOffset[sl] = numTri; triCount[sl] = n; numTri += n; }
		else { imask |= 1u << sl; numInner++; }
	}
	const uint32_t childBase = numInner ? atomicAdd(&counters->numNodes, numInner) : 0u;
	const uint32_t triBase = numTri ? atomicAdd(&counters->numTris, numTri) : 0u;
	const uint32_t queueBase = numInner ? atomicAdd(&counters->nextCount, numInner) : 0u;

	// quantisation frame
	const float3 p = f3(selfLo);
	const uint32_t ex = exponentFor(selfHi.x - selfLo.x), ey = exponentFor(selfHi.y - selfLo.y), ez = exponentFor(selfHi.z - selfLo.z);
	const double isx = ex ? 1.0 / double(__uint_as_float(ex << 23)) : 0.0;
	const double isy = ey ? 1.0 / double(__uint_as_float(ey << 23)) : 0.0;
	const double isz = ez ? 1.0 / double(__uint_as_float(ez << 23)) : 0.0;

	uint32_t qlo[3][8], qhi[3][8];
	uint32_t innerSeen = 0, leafTris = 0;
	for (int sl = 0; sl < 8; sl++) {
		for (int a = 0; a < 3; a++) { qlo[a][sl] = 255; qhi[a][sl] = 0; }   // empty slot: inverted box
		const uint32_t c = slotChild[sl];
		if (c == 0xffffffffu) continue;
		const float4 lo = nodeLo[c], hi = nodeHi[c];
		// exact in double: differences of floats, power-of-two scaling; floor / ceil round outwards
		qlo[0][sl] = uint32_t(fmin(fmax(floor((double(lo.x) - double(p.x)) * isx), 0.0), 255.0));
		qlo[1][sl] = uint32_t(fmin(fmax(floor((double(lo.y) - double(p.y)) * isy), 0.0), 255.0));
		qlo[2][sl] = uint32_t(fmin(fmax(floor((double(lo.z) - double(p.z)) * isz), 0.0), 255.0));
		qhi[0][sl] = uint32_t(fmin(fmax(ceil((double(hi.x) - double(p.x)) * isx), 0.0), 255.0));
		qhi[1][sl] = uint32_t(fmin(fmax(ceil((double(hi.y) - double(p.y)) * isy), 0.0), 255.0));
		qhi[2][sl] = uint32_t(fmin(fmax(ceil((double(hi.z) - double(p.z)) * isz), 0.0), 255.0));
		if (imask & (1u << sl)) {
			nextTasks[queueBase + innerSeen] = make_uint2(c, childBase + innerSeen);
			innerSeen++;
		}
		else {
			leafTris |= ((1u << triCount[sl]) - 1u) << (3 * sl);
			// gather the subtree's triangles (<= 3 leaves)
			uint32_t stack[4]; int sp = 0; uint32_t w = 0;
			stack[sp++] = c;
			while (sp) {
				const uint32_t x = stack[--sp];
				if (x < numLeaves) {
					trisOut[triBase + triOffset[sl] + w] = trisFlat[__float_as_uint(nodeLo[x].w)];
					w++;
				}
				else {
					stack[sp++] = __float_as_uint(nodeHi[x].w);
					stack[sp++] = __float_as_uint(nodeLo[x].w);
				}
			}
		}
	}
	auto pack4 = [](const uint32_t* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); };
	const uint32_t Ex = ex ? ex + 15u : 0u, Ey = ey ? ey + 15u : 0u, Ez = ez ? ez + 15u : 0u;
	WideNode out;
	out.n0 = make_float4(p.x, p.y, p.z, __uint_as_float(Ex | (Ey << 8) | (Ez << 16) | (imask << 24)));
	out.n1 = make_float4(__uint_as_float(childBase), __uint_as_float(triBase), __uint_as_float(leafTris), 0.0f);
	out.n2 = make_float4(__uint_as_float(pack4(qlo[0])), __uint_as_float(pack4(qlo[0] + 4)), __uint_as_float(pack4(qlo[1])), __uint_as_float(pack4(qlo[1] + 4)));
	out.n3 = make_float4(__uint_as_float(pack4(qlo[2])), __uint_as_float(pack4(qlo[2] + 4)), __uint_as_float(pack4(qhi[0])), __uint_as_float(pack4(qhi[0] + 4)));
	out.n4 = make_float4(__uint_as_float(pack4(qhi[1])), __uint_as_float(pack4(qhi[1] + 4)), __uint_as_float(pack4(qhi[2])), __uint_as_float(pack4(qhi[2] + 4)));
	wide[task.y] = out;
}

// scratch allocations of one build: freed on every path out of it (an early return on a CUDA error — the likeliest being out of
// memory on a large scene — must not leak gigabytes the context could never get back)
struct Scratch {
	std::vector<void*> ptrs;
	~Scratch() { for (void* p : ptrs) cudaFree(p); }
	template <typename T>
	cudaError_t alloc(T** p, size_t n) {
		void* q = nullptr;
		cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
		if (e == cudaSuccess) { ptrs.push_back(q); *p = static_cast<T*>(q); }
		return e;
	}
	template <typename T>
	T* keep(T* p) {   // ownership passes to the caller
		ptrs.erase(std::remove(ptrs.begin(), ptrs.end(), static_cast<void*>(p)), ptrs.end());
		return p;
	}
};

struct Events {
	cudaEvent_t a = nullptr, b = nullptr;
	~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

// Steps 2-4 for N primitives given as boxes (leafLo / leafHi, w lanes as flattenKernel writes them), their 48-byte records
// in the same order and the bounds of all of them: Morton sort, PLOC, collapse.  The records come out in leaf order.
cudaError_t buildFromLeaves(uint32_t N, const TriRecord* prims, const float4* leafLo, const float4* leafHi, const uint32_t* bounds,
                            cudaStream_t stream, BuildOutputs* out) {
	Scratch sc;
	float4 *nodeLo = nullptr, *nodeHi = nullptr;
	uint32_t *valsIn = nullptr, *valsOut = nullptr, *nodeCount = nullptr, *clA = nullptr, *clB = nullptr, *nearest = nullptr;
	uint64_t *keysIn = nullptr, *keysOut = nullptr, *flags = nullptr, *prefix = nullptr;
	void* cubTemp = nullptr;
	CK(sc.alloc(&nodeLo, 2 * size_t(N))); CK(sc.alloc(&nodeHi, 2 * size_t(N))); CK(sc.alloc(&nodeCount, 2 * size_t(N)));
	CK(sc.alloc(&valsIn, N)); CK(sc.alloc(&valsOut, N));
	CK(sc.alloc(&keysIn, N)); CK(sc.alloc(&keysOut, N));
	CK(sc.alloc(&clA, N)); CK(sc.alloc(&clB, N)); CK(sc.alloc(&nearest, N));
	CK(sc.alloc(&flags, size_t(N) + 1)); CK(sc.alloc(&prefix, size_t(N) + 1));

	const uint32_t B = 256, G = (N + B - 1) / B;
	mortonKernel<<<G, B, 0, stream>>>(N, leafLo, leafHi, bounds, keysIn, valsIn);

	size_t sortBytes = 0, scanBytes = 0;
	CK(cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, keysIn, keysOut, valsIn, valsOut, int(N), 0, 63, stream));
	CK(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, flags, prefix, int(N) + 1, stream));
	const size_t tempBytes = std::max(sortBytes, scanBytes);
	CK(sc.alloc(reinterpret_cast<char**>(&cubTemp), tempBytes));
	size_t tb = tempBytes;
	CK(cub::DeviceRadixSort::SortPairs(cubTemp, tb, keysIn, keysOut, valsIn, valsOut, int(N), 0, 63, stream));
	initLeavesKernel<<<G, B, 0, stream>>>(N, valsOut, leafLo, leafHi, nodeLo, nodeHi, nodeCount, clA);

	// PLOC iterations
	uint32_t n = N, nodesCreated = 0;
	uint32_t* cur = clA; uint32_t* nxt = clB;
	int guard = 0;
	while (n > 1) {
		const uint32_t g = (n + PlocBlock - 1) / PlocBlock;
		plocNearestKernel<<<g, PlocBlock, 0, stream>>>(n, cur, nodeLo, nodeHi, nearest);
		plocFlagKernel<<<g, PlocBlock, 0, stream>>>(n, nearest, flags);
		CK(cudaMemsetAsync(flags + n, 0, sizeof(uint64_t), stream));
		tb = tempBytes;
		CK(cub::DeviceScan::ExclusiveSum(cubTemp, tb, flags, prefix, int(n) + 1, stream));
		plocMergeKernel<<<g, PlocBlock, 0, stream>>>(n, N, nodesCreated, cur, nearest, flags, prefix, nodeLo, nodeHi, nodeCount, nxt);
		uint64_t totals = 0;
		CK(cudaMemcpyAsync(&totals, prefix + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
		CK(cudaStreamSynchronize(stream));
		const uint32_t kept = uint32_t(totals & 0xffffffffull), created = uint32_t(totals >> 32);
		if (created == 0 || ++guard > 4096) return cudaErrorUnknown;   // cannot happen: a mutual pair always exists
		n = kept; nodesCreated += created;
		std::swap(cur, nxt);
	}
	const uint32_t root = (N == 1) ? 0u : (N + nodesCreated - 1);
	float4 rootBox[2];
	CK(cudaMemcpyAsync(&rootBox[0], nodeLo + root, sizeof(float4), cudaMemcpyDeviceToHost, stream));
	CK(cudaMemcpyAsync(&rootBox[1], nodeHi + root, sizeof(float4), cudaMemcpyDeviceToHost, stream));

	// collapse, level by level
	WideNode* wide = nullptr; TriRecord* trisOut = nullptr; uint2 *qA = nullptr, *qB = nullptr; CollapseCounters* cc = nullptr;
	CK(sc.alloc(&wide, size_t(N) + 1)); CK(sc.alloc(&trisOut, N)); CK(sc.alloc(&qA, size_t(N) + 1)); CK(sc.alloc(&qB, size_t(N) + 1));
	CK(sc.alloc(&cc, 1));
	const CollapseCounters cc0 = { 1u, 0u, 0u, 0u };
	CK(cudaMemcpyAsync(cc, &cc0, sizeof(cc0), cudaMemcpyHostToDevice, stream));
	const uint2 rootTask = make_uint2(root, 0u);
	CK(cudaMemcpyAsync(qA, &rootTask, sizeof(rootTask), cudaMemcpyHostToDevice, stream));
	const uint32_t leafMax = getenv("RPT_LEAF_MAX") ? std::min(std::max(uint32_t(atoi(getenv("RPT_LEAF_MAX"))), 1u), 3u) : 3u;   // (build-time experiment switch)
	uint32_t numTasks = 1, depth = 0;
	while (numTasks) {
		collapseKernel<<<(numTasks + 63) / 64, 64, 0, stream>>>(numTasks, qA, qB, N, nodeLo, nodeHi, nodeCount, prims, wide, trisOut, cc, leafMax);
		CollapseCounters h;
		CK(cudaMemcpyAsync(&h, cc, sizeof(h), cudaMemcpyDeviceToHost, stream));
		CK(cudaStreamSynchronize(stream));
		numTasks = h.nextCount;
		out->numNodes = h.numNodes;
		CK(cudaMemsetAsync(&cc->nextCount, 0, sizeof(uint32_t), stream));
		std::swap(qA, qB);
		depth++;
	}
	CK(cudaGetLastError());
	// The traversal pushes at most one deferred node group per level of the wide tree (the rest of the group a child was popped
	// from), so the depth bounds the stack.  A tree deeper than the stack (PLOC puts no bound on pathological input) would
	// silently drop subtrees, so it is an error here instead.
	if (depth > uint32_t(MaxWideTreeDepth)) return cudaErrorInvalidValue;
	out->depth = depth;
	out->rootLo = make_float3(rootBox[0].x, rootBox[0].y, rootBox[0].z);
	out->rootHi = make_float3(rootBox[1].x, rootBox[1].y, rootBox[1].z);

	// shrink the node array to its final size
	WideNode* nodes = nullptr;
	CK(sc.alloc(&nodes, out->numNodes));
	CK(cudaMemcpyAsync(nodes, wide, size_t(out->numNodes) * sizeof(WideNode), cudaMemcpyDeviceToDevice, stream));
	CK(cudaStreamSynchronize(stream));
	out->nodes = sc.keep(nodes);
	out->tris = sc.keep(trisOut);
	out->numTris = N;
	return cudaSuccess;
}


// ---- two-level: instance records, TLAS leaf boxes ---------------------------------------------------------------------------
__global__ void rebaseKernel(WideNode* __restrict__ nodes, uint32_t n, uint32_t nodeOffset, uint32_t triOffset) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 n1 = nodes[i].n1;
	n1.x = __uint_as_float(__float_as_uint(n1.x) + nodeOffset);
	n1.y = __uint_as_float(__float_as_uint(n1.y) + triOffset);
	nodes[i].n1 = n1;
}

// record k: k == 0 the light triangles (world space), k >= 1 object instance k - 1
__global__ void instanceRecordKernel(uint32_t count, const RptObjectInstance* __restrict__ instances, const uint32_t* __restrict__ meshOfRecord,
                                     const BlasInfo* __restrict__ blas, const uint32_t* __restrict__ triOffsets,
                                     InstanceRecord* __restrict__ records, TriRecord* __restrict__ prims,
                                     float4* __restrict__ leafLo, float4* __restrict__ leafHi, uint32_t* __restrict__ bounds) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= count) return;
	const BlasInfo b = blas[meshOfRecord[k]];
	InstanceRecord r;
	r.r0 = make_float4(1.f, 0.f, 0.f, 0.f); r.r1 = make_float4(0.f, 1.f, 0.f, 0.f); r.r2 = make_float4(0.f, 0.f, 1.f, 0.f);
	r.rootNode = b.numTris ? b.rootNode : 0xffffffffu;
	r.customIndex = k; r.flatBase = 0; r.pad = 0;
	float3 lo = b.lo, hi = b.hi;
	if (k > 0) {
		const RptObjectInstance& I = instances[k - 1];
		const float* m = I.transformInv;   // column-major
		r.r0 = make_float4(m[0], m[4], m[8], m[12]);
		r.r1 = make_float4(m[1], m[5], m[9], m[13]);
		r.r2 = make_float4(m[2], m[6], m[10], m[14]);
		r.flatBase = triOffsets[k - 1];
		lo = f3(FLT_MAX); hi = f3(-FLT_MAX);
		for (int c = 0; c < 8; c++) {
			const float3 p = xformPoint(I.transform, make_float3((c & 1) ? b.hi.x : b.lo.x, (c & 2) ? b.hi.y : b.lo.y, (c & 4) ? b.hi.z : b.lo.z));
			lo = make_float3(fminf(lo.x, p.x), fminf(lo.y, p.y), fminf(lo.z, p.z));
			hi = make_float3(fmaxf(hi.x, p.x), fmaxf(hi.y, p.y), fmaxf(hi.z, p.z));
		}
		// the world-space box only has to contain what the object-space traversal can hit: padded for the rounding of the two
		// transforms (the ray goes through transformInv, the box through transform; the two need not be exact inverses)
		const float ext = fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z));
		const float3 pad = make_float3(1e-4f * ext + 1e-5f * (fmaxf(fabsf(lo.x), fabsf(hi.x)) + 1.0f),
		                               1e-4f * ext + 1e-5f * (fmaxf(fabsf(lo.y), fabsf(hi.y)) + 1.0f),
		                               1e-4f * ext + 1e-5f * (fmaxf(fabsf(lo.z), fabsf(hi.z)) + 1.0f));
		lo = lo - pad; hi = hi + pad;
	}
	if (!(lo.x <= hi.x) || !isfinite(lo.x + lo.y + lo.z + hi.x + hi.y + hi.z)) {   // empty mesh / NaN transform: never entered
		lo = f3(0.0f); hi = f3(0.0f);
		r.rootNode = 0xffffffffu;
	}
	records[k] = r;
	TriRecord t;
	t.t0 = make_float4(0.f, 0.f, 0.f, __uint_as_float(k)); t.t1 = make_float4(0.f, 0.f, 0.f, 0.f); t.t2 = make_float4(0.f, 0.f, 0.f, __uint_as_float(k));
	prims[k] = t;
	leafLo[k] = make_float4(lo.x, lo.y, lo.z, __uint_as_float(k));
	leafHi[k] = make_float4(hi.x, hi.y, hi.z, __uint_as_float(0xffffffffu));
	atomicMin(&bounds[0], floatFlip(lo.x)); atomicMin(&bounds[1], floatFlip(lo.y)); atomicMin(&bounds[2], floatFlip(lo.z));
	atomicMax(&bounds[3], floatFlip(hi.x)); atomicMax(&bounds[4], floatFlip(hi.y)); atomicMax(&bounds[5], floatFlip(hi.z));
}

} // namespace

cudaError_t buildBvh(const BuildInputs& in, cudaStream_t stream, BuildOutputs* out) {
	const uint32_t N = in.numTris;
	*out = BuildOutputs{};
	Scratch sc;
	if (N == 0) {
		// empty scene: one node without children, so traversal terminates immediately
		WideNode* nodes = nullptr; TriRecord* tris = nullptr;
		CK(sc.alloc(&nodes, 1));
		CK(sc.alloc(&tris, 1));
		CK(cudaMemsetAsync(nodes, 0, sizeof(WideNode), stream));
		CK(cudaStreamSynchronize(stream));
		out->nodes = sc.keep(nodes); out->tris = sc.keep(tris);
		out->numNodes = 1;
		return cudaSuccess;
	}
	Events ev;
	CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
	CK(cudaEventRecord(ev.a, stream));

	TriRecord* trisFlat = nullptr; float4 *leafLo = nullptr, *leafHi = nullptr; uint32_t* bounds = nullptr;
	CK(sc.alloc(&trisFlat, N)); CK(sc.alloc(&leafLo, N)); CK(sc.alloc(&leafHi, N)); CK(sc.alloc(&bounds, 8));
	const uint32_t initBounds[8] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u };
	CK(cudaMemcpyAsync(bounds, initBounds, sizeof(initBounds), cudaMemcpyHostToDevice, stream));
	flattenKernel<<<(N + 255) / 256, 256, 0, stream>>>(in, trisFlat, leafLo, leafHi, bounds);
	CK(buildFromLeaves(N, trisFlat, leafLo, leafHi, bounds, stream, out));
	CK(cudaEventRecord(ev.b, stream));
	CK(cudaStreamSynchronize(stream));
	CK(cudaEventElapsedTime(&out->buildMs, ev.a, ev.b));
	return cudaSuccess;
}

void TwoLevelState::release() {
	cudaFree(blasNodes); cudaFree(blasTris); cudaFree(tlasNodes); cudaFree(tlasLeaves); cudaFree(records); cudaFree(blasInfo); cudaFree(meshOfRecord);
	*this = TwoLevelState{};
}

cudaError_t rebuildTlas(const BuildInputs& in, cudaStream_t stream, TwoLevelState* out) {
	const uint32_t count = out->numRecords;
	Scratch sc;
	Events ev;
	CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
	CK(cudaEventRecord(ev.a, stream));
	TriRecord* prims = nullptr; float4 *leafLo = nullptr, *leafHi = nullptr; uint32_t* bounds = nullptr;
	CK(sc.alloc(&prims, count)); CK(sc.alloc(&leafLo, count)); CK(sc.alloc(&leafHi, count)); CK(sc.alloc(&bounds, 8));
	const uint32_t initBounds[8] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u };
	CK(cudaMemcpyAsync(bounds, initBounds, sizeof(initBounds), cudaMemcpyHostToDevice, stream));
	instanceRecordKernel<<<(count + 127) / 128, 128, 0, stream>>>(count, in.instances, out->meshOfRecord, out->blasInfo, in.triOffsets,
	                                                              out->records, prims, leafLo, leafHi, bounds);
	BuildOutputs tl;
	CK(buildFromLeaves(count, prims, leafLo, leafHi, bounds, stream, &tl));
	cudaFree(out->tlasNodes); cudaFree(out->tlasLeaves);
	out->tlasNodes = tl.nodes; out->tlasLeaves = tl.tris; out->numTlasNodes = tl.numNodes; out->tlasDepth = tl.depth;
	if (out->tlasDepth + out->maxBlasDepth + 2 > uint32_t(MaxWideTreeDepth) + 2) return cudaErrorInvalidValue;   // (one stack for both levels)
	CK(cudaEventRecord(ev.b, stream));
	CK(cudaStreamSynchronize(stream));
	CK(cudaEventElapsedTime(&out->tlasMs, ev.a, ev.b));
	return cudaSuccess;
}

cudaError_t buildTwoLevel(const TwoLevelInputs& in, cudaStream_t stream, TwoLevelState* out) {
	*out = TwoLevelState{};
	struct Guard { TwoLevelState* s; bool ok = false; ~Guard() { if (!ok) s->release(); } } guard{ out };
	Scratch sc;
	Events ev;
	CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
	CK(cudaEventRecord(ev.a, stream));
	const uint32_t numMeshes = uint32_t(in.meshes.size());
	const uint32_t numInstances = in.base.numInstances;

	// one BLAS per unique mesh, built by the single-level builder from an identity pseudo-instance over the mesh's index
	// range (object space), and one over the light triangles (world space)
	std::vector<RptObjectInstance> pseudo(std::max(numMeshes, 1u));
	for (uint32_t m = 0; m < numMeshes; m++) {
		RptObjectInstance I{};
		for (int d = 0; d < 4; d++) { I.transform[d * 5] = 1.0f; I.transformInv[d * 5] = 1.0f; I.transformInvT[d * 5] = 1.0f; }
		I.indexOffset = in.meshes[m].indexOffset; I.indexCount = in.meshes[m].indexCount;
		pseudo[m] = I;
	}
	RptObjectInstance* dPseudo = nullptr; uint32_t* dOffsets = nullptr;
	CK(sc.alloc(&dPseudo, pseudo.size())); CK(sc.alloc(&dOffsets, 2));
	CK(cudaMemcpyAsync(dPseudo, pseudo.data(), pseudo.size() * sizeof(RptObjectInstance), cudaMemcpyHostToDevice, stream));

	std::vector<BuildOutputs> blas(numMeshes + 1);
	struct BlasGuard { std::vector<BuildOutputs>& v; ~BlasGuard() { for (auto& b : v) { cudaFree(b.nodes); cudaFree(b.tris); } } } blasGuard{ blas };
	std::vector<BlasInfo> info(numMeshes + 1);
	uint64_t totalNodes = 0, totalTris = 0;
	for (uint32_t m = 0; m <= numMeshes; m++) {
		BuildInputs bi = in.base;
		if (m < numMeshes) {
			const uint32_t offs[2] = { 0u, in.meshes[m].indexCount / 3u };
			CK(cudaMemcpyAsync(dOffsets, offs, sizeof(offs), cudaMemcpyHostToDevice, stream));
			CK(cudaStreamSynchronize(stream));   // (offs is a stack variable)
			bi.instances = dPseudo + m; bi.triOffsets = dOffsets; bi.numInstances = 1; bi.numLights = 0; bi.numTris = offs[1];
		}
		else { bi.numInstances = 0; bi.numTris = in.base.numLights; }
		CK(buildBvh(bi, stream, &blas[m]));
		info[m].rootNode = uint32_t(totalNodes); info[m].numTris = blas[m].numTris;
		info[m].lo = blas[m].rootLo; info[m].hi = blas[m].rootHi;
		totalNodes += blas[m].numNodes; totalTris += std::max(blas[m].numTris, 1u);
		out->maxBlasDepth = std::max(out->maxBlasDepth, blas[m].depth);
	}
	if (totalNodes > 0x7fffffffull || totalTris > 0x7fffffffull) return cudaErrorInvalidValue;
	CK(cudaMalloc(reinterpret_cast<void**>(&out->blasNodes), totalNodes * sizeof(WideNode)));
	CK(cudaMalloc(reinterpret_cast<void**>(&out->blasTris), totalTris * sizeof(TriRecord)));
	uint32_t nodeAt = 0, triAt = 0;
	for (uint32_t m = 0; m <= numMeshes; m++) {
		CK(cudaMemcpyAsync(out->blasNodes + nodeAt, blas[m].nodes, size_t(blas[m].numNodes) * sizeof(WideNode), cudaMemcpyDeviceToDevice, stream));
		if (blas[m].numTris) CK(cudaMemcpyAsync(out->blasTris + triAt, blas[m].tris, size_t(blas[m].numTris) * sizeof(TriRecord), cudaMemcpyDeviceToDevice, stream));
		rebaseKernel<<<(blas[m].numNodes + 127) / 128, 128, 0, stream>>>(out->blasNodes + nodeAt, blas[m].numNodes, nodeAt, triAt);
		nodeAt += blas[m].numNodes; triAt += std::max(blas[m].numTris, 1u);
	}
	out->numBlasNodes = uint32_t(totalNodes); out->numBlasTris = uint32_t(totalTris);

	// records: [0] = lights (the last BLAS), [k + 1] = object instance k
	std::vector<uint32_t> meshOfRecord(size_t(numInstances) + 1);
	meshOfRecord[0] = numMeshes;
	for (uint32_t k = 0; k < numInstances; k++) meshOfRecord[k + 1] = in.meshOfInstance[k];
	out->numRecords = numInstances + 1;
	CK(cudaMalloc(reinterpret_cast<void**>(&out->records), size_t(out->numRecords) * sizeof(InstanceRecord)));
	CK(cudaMalloc(reinterpret_cast<void**>(&out->blasInfo), info.size() * sizeof(BlasInfo)));
	CK(cudaMalloc(reinterpret_cast<void**>(&out->meshOfRecord), meshOfRecord.size() * sizeof(uint32_t)));
	CK(cudaMemcpyAsync(out->blasInfo, info.data(), info.size() * sizeof(BlasInfo), cudaMemcpyHostToDevice, stream));
	CK(cudaMemcpyAsync(out->meshOfRecord, meshOfRecord.data(), meshOfRecord.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
	CK(cudaEventRecord(ev.b, stream));
	CK(cudaStreamSynchronize(stream));
	CK(cudaEventElapsedTime(&out->blasMs, ev.a, ev.b));
	CK(rebuildTlas(in.base, stream, out));
	guard.ok = true;
	return cudaSuccess;
}

} // namespace rt
