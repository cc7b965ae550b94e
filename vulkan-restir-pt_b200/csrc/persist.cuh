// Persistent-kernel work distribution with warp-level regeneration.
//
// The reference dispatches one shader invocation per pixel (8x8 work groups, src/shader/*.comp) and leaves SIMD
// packing of divergent path lengths to the driver's RT-core scheduling.  On B200 there is no such hardware: a warp
// whose 32 pixels follow paths of 1..15 bounces runs at the length of the longest one (measured: 8.8 of 32 lanes
// active, profiles/r1_megakernel_baseline.md).  These kernels instead launch one resident grid (SM count x
// occupancy) whose warps pull pixels from a global queue: whenever a lane's path ends, the lane is refilled with the
// next pixel at the top of the bounce loop, so every traversal call is entered by (nearly) all 32 lanes.
// Pixels are handed out in 8x4-tile-major order, so the lanes refilled together touch neighbouring G-buffer texels.
// Results do not depend on the schedule: a pixel's arithmetic and RNG stream are a function of (seed, x, y) only.
#pragma once
#include "rt_math.cuh"
#include "rt_types.cuh"

namespace rt {

constexpr unsigned FullWarp = 0xffffffffu;
constexpr int PersistBlock = 128;

struct PixelQueue {
	uint32_t* head;
	uint32_t tilesX, total;   // total = tilesX * tilesY * 32 queue slots (edge tiles contain slots outside the film)
	uint32_t width, rowBegin, rowEnd;

	__device__ __forceinline__ PixelQueue(const FrameView& f, uint32_t* counter)
		: head(counter), tilesX((f.width + 7u) / 8u), width(f.width), rowBegin(f.rowBegin), rowEnd(f.rowEnd) {
		total = tilesX * ((f.rowEnd - f.rowBegin + 3u) / 4u) * 32u;
	}

	// Must be called by all 32 lanes.  Lanes with need == true receive the next queue slots (one atomic per warp).
	// Returns true with (x, y) set when the lane got a pixel inside the film; sets exhausted (warp-uniform) once the
	// queue has run dry.
	__device__ __forceinline__ bool fetch(bool need, uint32_t& x, uint32_t& y, bool& exhausted) const {
		const unsigned mask = __ballot_sync(FullWarp, need);
		if (mask == 0u) return false;
		const uint32_t lane = threadIdx.x & 31u;
		const int leader = __ffs(int(mask)) - 1;
		uint32_t base = 0;
		if (int(lane) == leader) base = atomicAdd(head, uint32_t(__popc(mask)));
		base = __shfl_sync(FullWarp, base, leader);
		if (base + uint32_t(__popc(mask)) >= total) exhausted = true;
		if (!need) return false;
		const uint32_t slot = base + uint32_t(__popc(mask & ((1u << lane) - 1u)));
		if (slot >= total) return false;
		const uint32_t tile = slot >> 5, within = slot & 31u;
		x = (tile % tilesX) * 8u + (within & 7u);
		y = rowBegin + (tile / tilesX) * 4u + (within >> 3);
		return x < width && y < rowEnd;
	}
};

// appends one entry per calling lane to a device queue: one atomic per converged group of lanes
RT_DEV uint32_t queueAppend(uint32_t* __restrict__ count) {
	const unsigned mask = __activemask();
	const uint32_t lane = threadIdx.x & 31u;
	const int leader = __ffs(int(mask)) - 1;
	uint32_t base = 0;
	if (int(lane) == leader) base = atomicAdd(count, uint32_t(__popc(mask)));
	base = __shfl_sync(mask, base, leader);
	return base + uint32_t(__popc(mask & ((1u << lane) - 1u)));
}

} // namespace rt
