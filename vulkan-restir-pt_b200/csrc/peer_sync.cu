// Device-side hand-over between neighbouring strips of a multi-GPU film: monotonically increasing epoch flags
// written into the neighbour's memory over NVLink (system-scope stores after a system fence) and polled locally.
// Ordering per frame on every GPU (all on the frame's stream, no host involvement):
//   wait(neighbours finished spatial of epoch-1) -> temporal pass (pushes boundary rows into the neighbours' halo
//   rows) -> signal(temporal epoch) -> wait(neighbours' temporal epoch) -> spatial pass -> signal(spatial epoch)
#include "peer_sync.h"

namespace rt {

__global__ void peerSignalKernel(uint32_t* a, uint32_t* b, uint32_t epoch) {
	__threadfence_system();
	if (a) *reinterpret_cast<volatile uint32_t*>(a) = epoch;
	if (b) *reinterpret_cast<volatile uint32_t*>(b) = epoch;
	__threadfence_system();
}

__global__ void peerWaitKernel(const uint32_t* a, const uint32_t* b, uint32_t epoch, uint32_t* error) {
	const long long start = clock64();
	const long long limit = 8000000000ll;   // ~4 s at 2 GHz: never hang the GPU if a neighbour died
	for (;;) {
		const bool okA = !a || *reinterpret_cast<const volatile uint32_t*>(a) >= epoch;
		const bool okB = !b || *reinterpret_cast<const volatile uint32_t*>(b) >= epoch;
		if (okA && okB) break;
		if (clock64() - start > limit) { *error = 1u; break; }
		__nanosleep(200);
	}
	__threadfence_system();
}

void launchPeerSignal(uint32_t* a, uint32_t* b, uint32_t epoch, cudaStream_t st) {
	if (a || b) peerSignalKernel<<<1, 1, 0, st>>>(a, b, epoch);
}
void launchPeerWait(const uint32_t* a, const uint32_t* b, uint32_t epoch, uint32_t* error, cudaStream_t st) {
	if (a || b) peerWaitKernel<<<1, 1, 0, st>>>(a, b, epoch, error);
}

} // namespace rt
