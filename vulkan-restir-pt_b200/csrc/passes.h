// Host launchers of the per-pixel passes (one per reference shader entry point).
#pragma once
#include <cuda_runtime.h>
#include "rt_types.cuh"

namespace rt {

void launchGBuffer(const FrameView& f, const SceneView& s, cudaStream_t st);
void launchVisualizeAS(const FrameView& f, const SceneView& s, cudaStream_t st);
void launchPostProcess(const FrameView& f, const RptPostSettings& p, uchar4* rgba8, cudaStream_t st);
void launchDINaive(const FrameView& f, const SceneView& s, cudaStream_t st);
void launchDINaiveRT(const FrameView& f, const SceneView& s, cudaStream_t st);
void launchGINaive(const FrameView& f, const SceneView& s, cudaStream_t st);
void launchDIPathGen(const FrameView& f, const SceneView& s, const RptDISettings& p, cudaStream_t st);
void launchDITemporal(const FrameView& f, const SceneView& s, const RptDISettings& p, cudaStream_t st);
void launchDISpatial(const FrameView& f, const SceneView& s, const RptDISettings& p, cudaStream_t st);
void launchGIReSTIR(const FrameView& f, const SceneView& s, cudaStream_t st, cudaStream_t side = nullptr, cudaEvent_t fork = nullptr, cudaEvent_t join = nullptr);
// path tracing bounces [firstBounce, lastBounce] of the wavefront (bounce 0 = the G-buffer vertex); the C ABI layer
// runs the long tail [WavefrontTailStart, 15] on a second stream
struct KernelClock { virtual void tick(int rptKernelId) = 0; virtual ~KernelClock() = default; };   // called before every launch
void launchGRISPathTraceBounces(const FrameView& f, const SceneView& s, const RptGRISSettings& p, int firstBounce, int lastBounce, cudaStream_t st,
                                KernelClock* clock = nullptr, cudaStream_t side = nullptr, cudaEvent_t fork = nullptr, cudaEvent_t join = nullptr);
void launchGRISPathTraceTail(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st);
// tailMode 0: every pixel; 1: every pixel whose path is not in the tail; 2: only the pixels of the tail list
void launchGRISTemporal(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st, int tailMode = 0, KernelClock* clock = nullptr,
                        cudaStream_t side = nullptr, cudaEvent_t fork = nullptr);
// side / fork / join: a second stream and two events for the kernel that runs next to the main sequence (all NULL: one stream)
// (side2 / join2: a third stream for the replay wavefront, so that it runs next to the in-line replay list as well)
void launchGRISSpatial(const FrameView& f, const SceneView& s, const RptGRISSettings& p, cudaStream_t st, KernelClock* clock = nullptr,
                       cudaStream_t side = nullptr, cudaEvent_t fork = nullptr, cudaEvent_t join = nullptr,
                       cudaStream_t side2 = nullptr, cudaEvent_t join2 = nullptr,
                       cudaStream_t shadeStream = nullptr, cudaEvent_t shadeFork = nullptr, cudaEvent_t shadeDone = nullptr);
void launchTraceRays(const SceneView& s, const float4* rays, uint32_t n, RptIntersection* out, uint8_t* occluded, cudaStream_t st);

// wavefront traversal (trace_queue.cu): rays[2i] = {o, tmin}, rays[2i+1] = {d, tmax}; the ray count is read from
// countPtr on the device when it is not NULL (queues filled by a previous kernel), else countHost; head is the
// queue's fetch counter and must be zero at launch
void launchTraceQueueClosest(const SceneView& s, const float4* rays, const uint32_t* countPtr, uint32_t countHost, uint32_t* head,
                             RptIntersection* hits, cudaStream_t st);
void launchTraceQueueAny(const SceneView& s, const float4* rays, const uint32_t* countPtr, uint32_t countHost, uint32_t* head,
                         uint8_t* occluded, cudaStream_t st);

// pixel of this thread for a pass over rows [row0, row1); false when outside the film
constexpr int PassBlockX = 8, PassBlockY = 8;
inline dim3 passGrid(uint32_t width, uint32_t rows) {
	return dim3((width + PassBlockX - 1) / PassBlockX, (rows + PassBlockY - 1) / PassBlockY, 1);
}

// memory-system microbenchmarks (microbench.cu): bandwidth of `iterations` full reads of a `bytes` buffer by all SMs, and the
// dependent-load latency of a pointer chase through it
cudaError_t runMemBench(size_t bytes, int iterations, float* streamGBs, float* chaseNs, cudaStream_t st);

// grid of a persistent kernel: every SM filled to the kernel's occupancy limit (148 SMs on B200)
inline int persistentBlocks(const void* kernel, int blockSize) {
	int dev = 0, sms = 0, perSm = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, blockSize, 0);
	return (sms > 0 ? sms : 148) * (perSm > 0 ? perSm : 1);
}

} // namespace rt
