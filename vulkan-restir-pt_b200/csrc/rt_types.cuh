// Device-side views of the scene and of one frame's buffers, plus the compressed wide BVH node layout.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/restirpt.h"

namespace rt {

constexpr uint32_t InvalidHitIndex = 0xffffffffu;
constexpr uint32_t SpecialHitIndex = 0xfffffffeu;
constexpr uint32_t InvalidResourceIdx = 0xffffffffu;
constexpr float MinRayDistance = 1e-4f;
constexpr float MaxRayDistance = 1e7f;
constexpr int WorkCounterCount = 16;
constexpr float BaryEps = 1e-4f;   // tolerance of the triangle test, see bvh_traverse.cuh

// 80-byte compressed 8-wide BVH node (after Ylitie, Karras, Laine 2017): child boxes quantised to 8 bits relative
// to the node origin p with per-axis power-of-two scale 2^(e-127).
//   n0 = {p.x, p.y, p.z, E.x | E.y<<8 | E.z<<16 | imask<<24}     E = e + 15 (0: flat axis, step 0) — the biased exponent is
//                                                                  the traversal's mantissa-insertion unit (bvh_traverse.cuh)
//   n1 = {childBase, triBase, leafTris, 0}
//   n2 = {qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7]}
//   n3 = {qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7]}
//   n4 = {qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7]}
// imask bit s: slot s holds an inner child (the inner children are nodes childBase + rank of s among the imask bits);
// leafTris bit 3s+k: the leaf in slot s has a triangle k (at most 3); the triangles are stored from triBase on in bit order;
// an empty slot has an inverted box (qlo = 255, qhi = 0), which no ray interval can enter.
// The box test yields one hit bit per SLOT; nothing per child is decoded unless it is hit (the 2017 layout's per-child
// meta byte cost five ALU instructions per child and node, hit or not: profiles/r2_06_*).
struct __align__(16) WideNode {
	float4 n0, n1, n2, n3, n4;
};
static_assert(sizeof(WideNode) == 80, "CWBVH node must be 80 bytes");

// 48-byte triangle record in leaf order: v0 | e1 = v1-v0 | e2 = v2-v0, ids in the w lanes
//   t0.w = instanceIdx (0 = light), t1.w = triangleIdx within the instance, t2.w = flattened index (tie order)
struct __align__(16) TriRecord {
	float4 t0, t1, t2;
};

// One entry per TLAS leaf primitive (two-level scenes): the ray is taken into the instance's object space and the BLAS below
// rootNode is traversed there (bvh_traverse.cuh)
struct __align__(16) InstanceRecord {
	float4 r0, r1, r2;        // world -> object: rows of the 3x4 matrix
	uint32_t rootNode;        // the BLAS root in SceneView::nodes (0xffffffff: nothing to traverse)
	uint32_t customIndex;     // Intersection.instanceIdx: 0 = the light triangles, k + 1 = object instance k
	uint32_t flatBase;        // flattened index of this instance's triangle 0 (tie order across instances)
	uint32_t pad;
};

struct TextureView {
	const uchar4* texels;
	uint32_t width, height, filter;
	uint32_t pad;
};

struct SceneView {
	const WideNode* nodes;
	const TriRecord* tris;
	// two-level scenes only (tlasNodes != nullptr): nodes / tris then hold the BLASes in object space
	const WideNode* tlasNodes;
	const TriRecord* tlasLeaves;
	const InstanceRecord* instRecords;
	const RptMeshVertex* vertices;
	const uint32_t* indices;
	const RptMaterial* materials;
	const int32_t* materialIndices;
	const RptObjectInstance* instances;
	const RptObjectInstance* prevInstances;   // last frame's placements while an rpt_scene_update_instances is "in motion", else nullptr
	const RptTriangleLight* lights;
	const RptLightSampleTableElement* lightTable;
	const TextureView* textures;
	const float* srgbToLinear;    // 256 entries
	uint32_t numLights;
	unsigned long long* counters; // RptCounters layout, or nullptr when counting is off
};

// Wavefront path-tracing buffers of one frame (passes_gris.cu, trace_queue.cu); layout described there.
constexpr int PathStateWords = 11;    // float4 planes of the per-slot path state
constexpr int WavefrontMaxBounces = 16;
struct WavefrontView {
	float4* state[2];                 // [bounce & 1]: PathStateWords planes of `capacity` float4, indexed by queue slot
	float4* cold;                     // 2 float4 per pixel (storage index): rarely touched words of the path sample
	float4* rays[2];                  // extension-ray queues: {o, tmin}, {d, tmax} per slot
	uint32_t* pix[2];                 // pixel storage index of each slot
	RptIntersection* hits;            // closest hit of each slot of the current queue
	float4* shadowRays[2];            // shadow-ray queues
	uint8_t* occluded[2];
	uint32_t* counters;               // [bounce][4]: extension count, shadow count, extension head, shadow head
	uint32_t capacity;                // slots per queue = pixels of the owned rows
	// "tail": the few long paths still alive after bounce WavefrontTailStart - 1 finish on a second stream while the
	// temporal pass already runs for every other pixel (passes_gris.cu)
	uint32_t* tailMark;               // per pixel (storage index): == epoch when the pixel's path is in the tail
	uint32_t* tailList;               // pixels of the tail, one per slot of the extension queue of bounce WavefrontTailStart
	uint32_t epoch;                   // changes with every path-tracing pass
};
#ifndef RT_TAIL_START
#define RT_TAIL_START 7
#endif
constexpr int WavefrontTailStart = RT_TAIL_START;

// Buffers of the wavefront reuse passes (temporal: 1 candidate per pixel, spatial: 3), indexed by candidate * capacity +
// owned-pixel index (row-major over the owned rows): no compaction, every access is a full coalesced run.
constexpr int ShiftTaskWords = 7;
struct ReuseView {
	float4* task;                     // ShiftTaskWords planes of 3 * capacity float4
	float4* rays;                     // visibility rays, 2 float4 per candidate
	uint8_t* occluded;
	uint32_t* shadeList;              // pixels whose final shading needs a replay ray (rare) ...  (3 * capacity entries: the spatial
	                                  // pass first uses it for the (pixel, neighbour) pairs whose shift needs replay rays, counter [3])
	uint32_t* redoList;               // ... and pixels whose speculated random-number sequence did not hold (very rare)
	uint32_t* counters;               // [0] shade list size, [1] redo list size, [2] ray queue head, [3] spatial shift replay list size (in-line kernel),
	                                  // [4] temporal replay list size, [5] wavefront replay list size
	// wavefront replay (passes_gris.cu, "replay wavefront"): the path prefixes the reuse passes have to replay, one queue per bounce
	float4* rwRays[2];                // [round & 1]: {o, tmin}, {d, tmax} per slot
	float4* rwState[2];               // [round & 1]: {throughput, rng}, {entry, reconnection vertex id, -, -} per slot
	RptIntersection* rwHits;          // closest hit of each slot of the current round
	float4* rwRc;                     // result per list position: 3 float4 planes of `capacity` (the reference's GRISReconnectionData)
	uint32_t* rwList;                 // the replays of this pass: (candidate * capacity + pixel) entries, at most `capacity`
	uint32_t* rwCounters;             // [round][4]: queue size, -, traversal head, -
	uint32_t capacity;                // owned pixels
	uint32_t noReplayWavefront;       // A/B switch (RPT_NO_REPLAY_WAVEFRONT): every replay in the in-line list kernels
	uint32_t noShadeFromTask;         // A/B switch (RPT_NO_SHADE_FROM_TASK): the spatial pass replays every selected sample for its final shading
};

struct FrameView {
	uint32_t width, height;       // full film
	uint32_t rowBegin, rowEnd;    // rows this frame owns
	uint32_t storeBegin, storeEnd;// rows held in memory (owned + halo)
	float4* directOutput;
	float4* indirectOutput;
	float4* depthNormal;          // this frame
	const float4* depthNormalPrev;
	uint2* albedoMatId;
	const uint2* albedoMatIdPrev;
	float2* motion;
	RptDIReservoir* diThis;  const RptDIReservoir* diPrev;  RptDIReservoir* diTemp;
	RptGIReservoir* giThis;  const RptGIReservoir* giPrev;
	RptGRISReservoir* grisThis;  const RptGRISReservoir* grisPrev;  RptGRISReservoir* grisTemp;
	const RptGRISReservoir* grisStale;   // the reservoirs of two frames ago (background pixels of the path-tracing pass keep them)
	RptIntersection* primaryIsec;
	RptCamera camera, prevCamera;

	// multi-GPU strips: the neighbours' temp reservoir buffers in peer memory (nullptr at a film edge / single GPU)
	uint32_t halo;
	uint32_t peerUpStoreBegin, peerDownStoreBegin;
	RptGRISReservoir* peerGrisUp;  RptGRISReservoir* peerGrisDown;
	RptDIReservoir* peerDiUp;      RptDIReservoir* peerDiDown;
	// ... and their final ("this") reservoir buffers of the current frame: the boundary rows of the spatial / GI pass output are
	// mirrored into the neighbours' halo rows, so that next frame's previous-frame lookups near a cut find what a single GPU finds
	RptGRISReservoir* peerGrisThisUp;  RptGRISReservoir* peerGrisThisDown;
	RptDIReservoir* peerDiThisUp;      RptDIReservoir* peerDiThisDown;
	RptGIReservoir* peerGiThisUp;      RptGIReservoir* peerGiThisDown;
	uint32_t prevRowBegin, prevRowEnd;   // rows whose previous-frame reservoirs (and bilinear G-buffer taps) are valid on this GPU
	uchar4* gatherImage;                 // full-film RGBA8 image on the root strip's GPU (peer memory), or nullptr
	bool striped;                 // true when this frame is one strip of a larger film
	uint32_t* work;               // work-queue heads of the persistent kernels (WorkCounterCount words)
	WavefrontView wf;
	ReuseView ru;

	__device__ __forceinline__ size_t index(uint32_t x, uint32_t y) const { return size_t(y - storeBegin) * width + x; }
};

} // namespace rt
