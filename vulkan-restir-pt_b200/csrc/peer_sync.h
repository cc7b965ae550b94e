#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rt {

// flag slots inside a frame's flags buffer (written by the neighbours, polled by the owner)
enum PeerFlag { GrisTemporalFromUp = 0, GrisTemporalFromDown = 1, GrisSpatialFromUp = 2, GrisSpatialFromDown = 3,
                DiTemporalFromUp = 4, DiTemporalFromDown = 5, DiSpatialFromUp = 6, DiSpatialFromDown = 7, PeerError = 8,
                GiFromUp = 9, GiFromDown = 10, PeerFlagCount = 16 };
// slots of the gather flags on the root strip's GPU: [0] = release epoch (written by the root, polled by the strips over
// NVLink), [1 + i] = arrival epoch of strip i (written by strip i)
constexpr int GatherReleaseFlag = 0, GatherArrivalFlag0 = 1, GatherMaxStrips = 62;

void launchPeerSignal(uint32_t* a, uint32_t* b, uint32_t epoch, cudaStream_t st);
// `error` = device word, `hostError` = the same condition in mapped host memory (polled by the pass prologues without a sync)
void launchPeerWait(const uint32_t* a, const uint32_t* b, uint32_t epoch, uint32_t* error, uint32_t* hostError, cudaStream_t st);
void launchPeerWaitMany(const uint32_t* flags, uint32_t count, uint32_t epoch, uint32_t* error, uint32_t* hostError, cudaStream_t st);

} // namespace rt
