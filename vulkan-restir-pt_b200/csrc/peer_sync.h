#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rt {

// flag slots inside a frame's 16-word flags buffer (written by the neighbours, polled by the owner)
enum PeerFlag { GrisTemporalFromUp = 0, GrisTemporalFromDown = 1, GrisSpatialFromUp = 2, GrisSpatialFromDown = 3,
                DiTemporalFromUp = 4, DiTemporalFromDown = 5, DiSpatialFromUp = 6, DiSpatialFromDown = 7, PeerError = 8, PeerFlagCount = 16 };

void launchPeerSignal(uint32_t* a, uint32_t* b, uint32_t epoch, cudaStream_t st);
void launchPeerWait(const uint32_t* a, const uint32_t* b, uint32_t epoch, uint32_t* error, cudaStream_t st);

} // namespace rt
