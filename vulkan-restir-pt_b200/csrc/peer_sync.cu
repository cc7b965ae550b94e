// Device-side hand-over between neighbouring strips of a multi-GPU film: monotonically increasing epoch flags
// written into the neighbour's memory over NVLink (system-scope stores after a system fence) and polled locally.
// Ordering per frame on every GPU (all on the frame's stream, no host involvement):
//   wait(neighbours finished spatial of epoch-1) -> temporal pass (pushes boundary rows into the neighbours' halo
//   rows) -> signal(temporal epoch) -> wait(neighbours' temporal epoch) -> spatial pass (mirrors the boundary rows of its
//   output into the neighbours' halo rows) -> signal(spatial epoch)
// A wait that runs into its time limit (a neighbour died, or the strips were not driven in lock step) raises an error word on
// the device and in mapped host memory; the host turns it into RPT_ERR_PEER on the next pass call (capi.cu).
#include "peer_sync.h"

namespace rt {

namespace {
constexpr long long WaitLimitClocks = 8000000000ll;   // ~4 s at 2 GHz: never hang the GPU

__device__ __forceinline__ void raise(uint32_t* error, uint32_t* hostError) {
	*error = 1u;
	if (hostError) *reinterpret_cast<volatile uint32_t*>(hostError) = 1u;
	__threadfence_system();
}
}

__global__ void peerSignalKernel(uint32_t* a, uint32_t* b, uint32_t epoch) {
	__threadfence_system();
	if (a) *reinterpret_cast<volatile uint32_t*>(a) = epoch;
	if (b) *reinterpret_cast<volatile uint32_t*>(b) = epoch;
	__threadfence_system();
}

__global__ void peerWaitKernel(const uint32_t* a, const uint32_t* b, uint32_t epoch, uint32_t* error, uint32_t* hostError) {
	const long long start = clock64();
	for (;;) {
		const bool okA = !a || *reinterpret_cast<const volatile uint32_t*>(a) >= epoch;
		const bool okB = !b || *reinterpret_cast<const volatile uint32_t*>(b) >= epoch;
		if (okA && okB) break;
		if (clock64() - start > WaitLimitClocks) { raise(error, hostError); break; }
		__nanosleep(200);
	}
	__threadfence_system();
}

// one thread per flag: all of flags[0 .. count) must reach `epoch`
__global__ void peerWaitManyKernel(const uint32_t* flags, uint32_t count, uint32_t epoch, uint32_t* error, uint32_t* hostError) {
	if (threadIdx.x < count) {
		const long long start = clock64();
		while (*reinterpret_cast<const volatile uint32_t*>(flags + threadIdx.x) < epoch) {
			if (clock64() - start > WaitLimitClocks) { raise(error, hostError); break; }
			__nanosleep(500);
		}
	}
	__threadfence_system();
}

void launchPeerSignal(uint32_t* a, uint32_t* b, uint32_t epoch, cudaStream_t st) {
	if (a || b) peerSignalKernel<<<1, 1, 0, st>>>(a, b, epoch);
}
void launchPeerWait(const uint32_t* a, const uint32_t* b, uint32_t epoch, uint32_t* error, uint32_t* hostError, cudaStream_t st) {
	if (a || b) peerWaitKernel<<<1, 1, 0, st>>>(a, b, epoch, error, hostError);
}
void launchPeerWaitMany(const uint32_t* flags, uint32_t count, uint32_t epoch, uint32_t* error, uint32_t* hostError, cudaStream_t st) {
	if (flags && count) peerWaitManyKernel<<<1, 64, 0, st>>>(flags, count, epoch, error, hostError);
}

} // namespace rt
