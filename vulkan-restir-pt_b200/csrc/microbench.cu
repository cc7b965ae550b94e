// Memory-system microbenchmarks behind rpt_membench ("L2 peak to be measured once by a pointer-chase / stream
// microbench on the GPU box").  VeachAjar's BVH + triangles (23 MB) live in B200's 126 MB L2, so the roof that applies to the
// traversal kernels on that scene is the L2's, not HBM's; bench.py reports the kernels against both.
//   stream: every SM reads the whole buffer `iterations` times with 16-byte loads (coalesced, grid = SMs x resident blocks);
//           a buffer that fits L2 gives the L2 -> SM bandwidth, one much larger than L2 gives the HBM read bandwidth.
//   chase:  one thread follows a random cyclic permutation of 128-byte lines: the dependent-load latency of the level the
//           buffer lives in (what a traversal step pays per node when nothing hides it).
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "passes.h"

namespace rt {

namespace {

__global__ void __launch_bounds__(256) streamReadKernel(const uint4* __restrict__ buf, size_t n16, int iterations, uint32_t* sink) {
	uint32_t acc = 0;
	const size_t stride = size_t(gridDim.x) * blockDim.x;
	for (int it = 0; it < iterations; it++) {
		// rotate the starting point per iteration so that a block does not re-read the lines its own L1 still holds
		const size_t shift = (size_t(it) * 7919u * blockDim.x) % n16;
		for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
			size_t j = i + shift;
			if (j >= n16) j -= n16;
			const uint4 v = __ldcg(buf + j);   // cache at L2 only
			acc += v.x ^ v.y ^ v.z ^ v.w;
		}
	}
	if (acc == 0x12345678u) *sink = acc;   // keeps the loads alive
}

__global__ void chaseKernel(const uint32_t* __restrict__ next, uint32_t steps, uint32_t* sink, long long* cycles) {
	uint32_t p = 0;
	for (uint32_t i = 0; i < 1024; i++) p = __ldcg(next + size_t(p) * 32);   // warm the TLB and the first lines
	const long long t0 = clock64();
	for (uint32_t i = 0; i < steps; i++) p = __ldcg(next + size_t(p) * 32);
	const long long t1 = clock64();
	*cycles = t1 - t0;
	*sink = p;
}

} // namespace

cudaError_t runMemBench(size_t bytes, int iterations, float* streamGBs, float* chaseNs, cudaStream_t st) {
	if (bytes < (1u << 20)) bytes = 1u << 20;
	bytes &= ~size_t(4095);
	uint4* buf = nullptr; uint32_t* sink = nullptr; long long* cycles = nullptr;
	cudaError_t e = cudaMalloc(&buf, bytes);
	if (e != cudaSuccess) return e;
	if ((e = cudaMalloc(&sink, 4)) != cudaSuccess) { cudaFree(buf); return e; }
	if ((e = cudaMalloc(&cycles, 8)) != cudaSuccess) { cudaFree(buf); cudaFree(sink); return e; }
	auto done = [&](cudaError_t r) { cudaFree(buf); cudaFree(sink); cudaFree(cycles); return r; };

	// pointer chase first (it needs the buffer's first word of every 128-byte line): a random cyclic permutation of the lines
	const uint32_t lines = uint32_t(bytes / 128);
	{
		std::vector<uint32_t> order(lines), host(size_t(lines) * 32, 0u);
		for (uint32_t i = 0; i < lines; i++) order[i] = i;
		uint64_t s = 0x9e3779b97f4a7c15ull;
		for (uint32_t i = lines - 1; i > 0; i--) {   // Fisher-Yates with splitmix64
			s += 0x9e3779b97f4a7c15ull;
			uint64_t z = s; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; z ^= z >> 31;
			const uint32_t j = uint32_t(z % (i + 1));
			const uint32_t t = order[i]; order[i] = order[j]; order[j] = t;
		}
		for (uint32_t i = 0; i < lines; i++) host[size_t(order[i]) * 32] = order[(i + 1) % lines];
		if ((e = cudaMemcpyAsync(buf, host.data(), bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) return done(e);
		if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(e);
	}
	int dev = 0, sms = 0, perSm = 0, khz = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, streamReadKernel, 256, 0);
	const int grid = (sms > 0 ? sms : 148) * (perSm > 0 ? perSm : 1);

	if (chaseNs) {
		const uint32_t steps = 20000;
		chaseKernel<<<1, 1, 0, st>>>(reinterpret_cast<const uint32_t*>(buf), steps, sink, cycles);   // (also pulls the lines into L2)
		chaseKernel<<<1, 1, 0, st>>>(reinterpret_cast<const uint32_t*>(buf), steps, sink, cycles);
		long long c = 0;
		if ((e = cudaMemcpyAsync(&c, cycles, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return done(e);
		if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(e);
		*chaseNs = float(double(c) / steps / (double(khz) * 1e-6));   // cycles / (cycles per ns); clock rate = the SM's maximum
	}
	if (streamGBs) {
		cudaEvent_t e0, e1;
		cudaEventCreate(&e0); cudaEventCreate(&e1);
		streamReadKernel<<<grid, 256, 0, st>>>(buf, bytes / 16, 2, sink);   // warm-up: the buffer settles in L2 (if it fits)
		cudaEventRecord(e0, st);
		streamReadKernel<<<grid, 256, 0, st>>>(buf, bytes / 16, iterations, sink);
		cudaEventRecord(e1, st);
		if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return done(e); }
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		cudaEventDestroy(e0); cudaEventDestroy(e1);
		*streamGBs = float(double(bytes) * iterations / (double(ms) * 1e-3) / 1e9);
	}
	return done(cudaGetLastError());
}

} // namespace rt
