// C ABI of librestirpt.so (include/restirpt.h): contexts, scene upload + BVH build, frame buffers with the
// reference's ping-pong wiring (src/Renderer.cpp:324-347), pass dispatch, read-back.  No CPU fallback exists:
// every entry point needs a CUDA device.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <map>
#include <vector>
#include "../../include/restirpt.h"
#include "bvh_build.h"
#include "passes.h"
#include "peer_sync.h"
#include <unistd.h>

#define RPT_API extern "C" __attribute__((visibility("default")))

using namespace rt;

struct RptCtx {
	int device = 0;
	std::string lastError;
	cudaStream_t stream = nullptr;         // scene builds and raw ray queries
	unsigned long long* counters = nullptr;
	bool countersOn = false;
	// A/B switches of the measurements in profiles/, read from the environment ONCE at context creation (never in a pass)
	bool traceOneStream = false, spatialOneStream = false, noFrameOverlap = false;
	int tailForm = 0;   // 0: by the frame's size, 1: in-line tail kernel, 2: wavefront rounds
	bool noShadeFromTask = false, noReplayWavefront = false, noShadeApart = false;
	uint32_t rwMinList = 80000;    // replay pairs per spatial pass from which the replay wavefront is used (profiles/r2_24_*)
	int priorityMode = 0;   // stream priorities (profiles/r2_20_*): 0 = late set, tail and path-tracer side stream above the frame's stream; 1 = all equal; 2 = frame's stream + its side stream above the late set
};

struct RptScene {
	RptCtx* ctx = nullptr;
	SceneView view{};
	std::vector<void*> allocations;   // everything but the acceleration structure
	BuildInputs buildInputs{};        // device views kept for rpt_scene_update_instances
	RptBvhStats stats{};
	bool twoLevel = false;
	TwoLevelState tl{};               // two-level scenes: BLASes, TLAS, instance records (view.nodes / tris alias tl.blas*)
	std::vector<MeshRange> geometry;  // (indexOffset, indexCount) of every instance, as created
	RptObjectInstance* prevInstances = nullptr;   // device copy of the placements before the last update (per-instance motion vectors)
};

struct RptFrame {
	RptCtx* ctx = nullptr;
	uint32_t width = 0, height = 0, rowBegin = 0, rowEnd = 0, storeBegin = 0, storeEnd = 0;
	// Frames flipped since the last clear.  The DI / GI reservoirs are ping-pong pairs (slot flips & 1 = this frame); the G-buffer
	// images, the motion vectors and the GRIS reservoirs rotate through THREE slots (flips % 3 = this frame, the one before = the
	// previous frame), and the wavefront path-tracing state exists twice (flips & 1): the third slot / the other set is what the
	// G-buffer and path-tracing passes of the NEXT frame fill while this frame's reuse passes still read the other two (below).
	uint32_t flips = 0;
	uint32_t c2() const { return flips & 1u; }
	uint32_t p2() const { return (flips & 1u) ^ 1u; }
	uint32_t c3() const { return flips % 3u; }
	uint32_t p3() const { return (flips + 2u) % 3u; }
	uint32_t n3() const { return (flips + 1u) % 3u; }   // the slot of two frames ago = the slot of the next frame
	cudaStream_t stream = nullptr;
	float4 *directOutput = nullptr, *indirectOutput = nullptr, *depthNormal[3] = { nullptr, nullptr, nullptr };
	uint2* albedoMatId[3] = { nullptr, nullptr, nullptr };
	float2* motion[3] = { nullptr, nullptr, nullptr };
	RptDIReservoir *di[2] = { nullptr, nullptr }, *diTemp = nullptr;
	RptGIReservoir* gi[2] = { nullptr, nullptr };
	RptGRISReservoir *gris[3] = { nullptr, nullptr, nullptr }, *grisTemp = nullptr;
	RptIntersection* primaryIsec = nullptr;
	uchar4* rgba8 = nullptr;
	// pipelined read-back (rpt_postprocess_async): a second device image, a copy stream, one event pair per image
	static constexpr int ReadbackDepth = 3;    // images in flight: with the reuse passes one frame behind the path tracer the host
	uchar4* rgba8Ring[ReadbackDepth - 1] = {}; // has to run two frames ahead of the image it waits for (profiles/r2_25_*); ring = f->rgba8 + these
	cudaStream_t copyStream = nullptr;
	cudaEvent_t postDone[ReadbackDepth] = {}, copyDone[ReadbackDepth] = {};
	uint64_t asyncTicket = 0;                  // read-backs issued so far; ticket t used image / events [t % ReadbackDepth]
	RptCamera camera{}, prevCamera{};
	size_t pixels() const { return size_t(width) * (storeEnd - storeBegin); }

	// multi-GPU strips: neighbours' buffers in peer memory + device-side epoch flags (peer_sync.h)
	uint32_t halo = 0;
	uint32_t* flags = nullptr;                 // PeerFlagCount words, written by the neighbours
	uint32_t* work = nullptr;                  // WorkCounterCount queue heads of the persistent kernels
	WavefrontView wfSet[2]{};                  // wavefront path-tracing queues (owned rows only), one set per frame parity
	WavefrontView& wf() { return wfSet[flips & 1u]; }
	ReuseView ru{};                            // wavefront temporal / spatial reuse
	cudaStream_t tailStream = nullptr;         // the long tail of the path-tracing pass runs here ...
	cudaEvent_t tailFork = nullptr, tailDone = nullptr;
	bool tailPending = false;                  // ... until the next pass joins it back into `stream`
	// Two frames in flight: the reuse passes of a ReSTIR PT frame — temporal, spatial, post-process, whose replay kernels and whose
	// wait for the path tracer's tail are a few long dependent chains that leave most of the GPU idle — run on a stream set of their
	// own ("late"), in order, while the frame's stream goes on with the G-buffer and the path tracer of the NEXT frame.  Those two
	// passes touch nothing the late passes read or write: they fill the third slot of the G-buffer / motion / GRIS rotation and the
	// other set of wavefront queues.  The frame's stream is at most one frame ahead: before the G-buffer of frame k+2 reuses the
	// slots of frame k-1 it waits for the late passes of frame k (lateFrameDone[k & 1]), the last readers of those slots.  Any other
	// pass, read-back or query joins everything first.
	cudaStream_t lateStream = nullptr, lateSide = nullptr, lateSide2 = nullptr;
	cudaEvent_t lateFork = nullptr, lateDone = nullptr, lateSideFork = nullptr, lateSideDone = nullptr, lateSide2Done = nullptr, lateHead = nullptr;
	cudaEvent_t lateFrameDone[2] = { nullptr, nullptr };
	bool latePending = false;
	// the spatial pass's shade list and the post-process behind it run on the late set's third stream, next to the NEXT frame's
	// temporal pass (they gate nothing but the image); the next spatial pass and every join wait for shadeDone
	cudaEvent_t shadeFork = nullptr, shadeDone = nullptr;
	bool shadePending = false;
	// the path tracer's paired launches (any-hit next to closest-hit) have a side stream of their own: the tail stream still
	// carries the previous frame's tail when the next path tracer starts
	cudaStream_t ptSide = nullptr;
	cudaEvent_t ptFork = nullptr, ptJoin = nullptr;
	uint32_t* hostReuseCounters = nullptr;     // pinned: the list sizes of the last spatial pass, copied back asynchronously
	bool wavefrontTail = true;                 // the path tracer's tail as wavefront rounds (see rpt_ctx_create)
	uint32_t lastWfSet = 0;                    // the wavefront set of the last path-tracing pass (rpt_wavefront_counters)
	struct Peer {
		bool connected = false, ipc = false;
		RptGRISReservoir* grisTemp = nullptr; RptDIReservoir* diTemp = nullptr; uint32_t* flags = nullptr;
		RptGRISReservoir* gris[3] = { nullptr, nullptr, nullptr };   // the neighbour's buffers of final reservoirs, indexed like
		RptDIReservoir* di[2] = { nullptr, nullptr };                // OUR slots (the phase differences at connect time are
		RptGIReservoir* gi[2] = { nullptr, nullptr };                // folded in)
		uint32_t storeBegin = 0;
	} up, down;
	uint32_t grisEpoch = 0, diEpoch = 0, giEpoch = 0;
	uint32_t* hostError = nullptr;             // mapped pinned word: a device-side wait timed out (sticky)
	uint32_t* hostErrorDev = nullptr;          // its device address
	// final image gather (rpt_frame_gather_*): the film image + flags on the root strip's GPU
	struct Gather {
		bool connected = false, ipc = false, root = false;
		uchar4* image = nullptr; uint32_t* flags = nullptr;
		uint32_t strip = 0, numStrips = 0;
		uint32_t epoch = 0;                    // post-process passes since the connection (strip side)
		uint32_t gathered = 0;                 // rpt_gather_output calls since the creation (root side)
	} gather;
	uchar4* gatherImageOwned = nullptr;        // root: the allocation behind gather.image
	uint32_t* gatherFlagsOwned = nullptr;

	// per-pass timing
	struct Pending { int pass; cudaEvent_t a, b; bool poolB; };
	bool timing = false;
	std::vector<Pending> pending;
	std::vector<cudaEvent_t> eventPool;
	RptPassStats stats{};
};

// the tail of the last path-tracing pass (second stream) must be complete before anything else touches the frame
static void joinTail(RptFrame* f) {
	if (f->tailPending) { cudaStreamWaitEvent(f->stream, f->tailDone, 0); f->tailPending = false; }
}

// the late passes of the previous frame (spatial reuse, post-process) must be complete before anything but the next frame's
// G-buffer and path-tracing passes touches the frame
static void joinLate(RptFrame* f) {
	if (f->latePending) { cudaStreamWaitEvent(f->stream, f->lateDone, 0); f->latePending = false; }
	if (f->shadePending) { cudaStreamWaitEvent(f->stream, f->shadeDone, 0); f->shadePending = false; }
}
static cudaError_t syncFrame(RptFrame* f) {
	joinTail(f);
	joinLate(f);
	return cudaStreamSynchronize(f->stream);
}
// runs the calling pass on the late stream set: everything the helpers reach through f->stream / f->tailStream goes there
struct LateScope {
	RptFrame* f; bool on; cudaStream_t s0 = nullptr, t0 = nullptr; cudaEvent_t a0 = nullptr, b0 = nullptr;
	explicit LateScope(RptFrame* f_) : f(f_), on(f_->lateStream != nullptr && !f_->ctx->noFrameOverlap) {
		if (!on) return;
		cudaEventRecord(f->lateFork, f->stream);              // after everything enqueued for this frame so far
		cudaStreamWaitEvent(f->lateStream, f->lateFork, 0);
		s0 = f->stream; t0 = f->tailStream; a0 = f->tailFork; b0 = f->tailDone;
		f->stream = f->lateStream; f->tailStream = f->lateSide; f->tailFork = f->lateSideFork; f->tailDone = f->lateSideDone;
	}
	~LateScope() {
		if (!on) return;
		cudaEventRecord(f->lateDone, f->stream);
		cudaEventRecord(f->lateFrameDone[f->flips & 1u], f->stream);
		f->stream = s0; f->tailStream = t0; f->tailFork = a0; f->tailDone = b0;
		f->latePending = true;
	}
};
// the post-process after a spatial pass whose shade list runs on the third stream follows it THERE (inside a LateScope)
struct ShadeScope {
	RptFrame* f; bool on; cudaStream_t s0 = nullptr;
	ShadeScope(RptFrame* f_, bool inLateScope) : f(f_), on(inLateScope && f_->shadePending) {
		if (!on) return;
		s0 = f->stream;
		f->stream = f->lateSide2;
	}
	~ShadeScope() {
		if (!on) return;
		cudaEventRecord(f->shadeDone, f->stream);
		f->stream = s0;
	}
};
static bool pipelined(const RptFrame* f) { return f->lateStream != nullptr && !f->ctx->noFrameOverlap; }
// G-buffer / path tracer of frame k: the slots they fill were last read by the late passes of frame k-2
static void waitSlotReaders(RptFrame* f) {
	if (pipelined(f)) cudaStreamWaitEvent(f->stream, f->lateFrameDone[f->flips & 1u], 0);
}

static cudaEvent_t takeEvent(RptFrame* f) {
	if (!f->eventPool.empty()) { cudaEvent_t e = f->eventPool.back(); f->eventPool.pop_back(); return e; }
	cudaEvent_t e = nullptr;
	cudaEventCreate(&e);
	return e;
}
static void drainTiming(RptFrame* f) {
	for (auto& p : f->pending) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
			if (p.pass < RPT_PASS_COUNT) { f->stats.ms[p.pass] += ms; f->stats.launches[p.pass] += 1; }
			else { f->stats.kernelMs[p.pass - RPT_PASS_COUNT] += ms; f->stats.kernelLaunches[p.pass - RPT_PASS_COUNT] += 1; }
		}
		f->eventPool.push_back(p.a);
		if (p.poolB) f->eventPool.push_back(p.b);   // kernel spans share their end event with the next span's start
	}
	f->pending.clear();
}
struct PassTimer {
	RptFrame* f; int pass; cudaEvent_t a = nullptr, b = nullptr;
	PassTimer(RptFrame* f_, int pass_) : f(f_), pass(pass_) {
		if (f->timing) { a = takeEvent(f); b = takeEvent(f); cudaEventRecord(a, f->stream); }
	}
	~PassTimer() {
		if (f->timing) {
			cudaEventRecord(b, f->stream);
			f->pending.push_back({ pass, a, b, true });
			if (f->pending.size() >= 4096) { cudaStreamSynchronize(f->stream); if (f->lateStream) cudaStreamSynchronize(f->lateStream); drainTiming(f); }
		}
	}
};

// per-kernel timing inside a multi-kernel pass: an event before every launch, the span up to the next event is
// booked on that kernel (passes.h KernelClock)
struct FrameKernelClock : KernelClock {
	RptFrame* f; int last = -1; cudaEvent_t lastEv = nullptr;
	explicit FrameKernelClock(RptFrame* f_) : f(f_) {}
	void tick(int kernelId) override {
		cudaEvent_t e = takeEvent(f);
		cudaEventRecord(e, f->stream);
		if (last >= 0) f->pending.push_back({ RPT_PASS_COUNT + last, lastEv, e, kernelId < 0 });   // the last span also returns its end event
		last = kernelId; lastEv = e;
	}
	~FrameKernelClock() { if (last >= 0) tick(-1); }
};

static thread_local std::string gThreadError;

static int fail(RptCtx* ctx, int code, const std::string& msg) {
	gThreadError = msg;
	if (ctx) ctx->lastError = msg;
	return code;
}
static int cudaFail(RptCtx* ctx, cudaError_t e, const char* what) {
	return fail(ctx, e == cudaErrorMemoryAllocation ? RPT_ERR_OOM : RPT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(ctx, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return cudaFail(ctx, e_, #expr); } while (0)

RPT_API int rpt_version(void) { return 1; }

RPT_API const char* rpt_last_error(const RptCtx* ctx) { return ctx ? ctx->lastError.c_str() : gThreadError.c_str(); }

RPT_API int rpt_ctx_create(int cudaDevice, RptCtx** out) {
	if (!out) return fail(nullptr, RPT_ERR_INVALID, "rpt_ctx_create: out is NULL");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		return fail(nullptr, RPT_ERR_NO_DEVICE, std::string("rpt_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
	}
	if (cudaDevice < 0 || cudaDevice >= n) return fail(nullptr, RPT_ERR_INVALID, "rpt_ctx_create: device index out of range");
	RptCtx* ctx = new RptCtx;
	ctx->device = cudaDevice;
	CU(ctx, cudaSetDevice(cudaDevice));
	CU(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	CU(ctx, cudaMalloc(&ctx->counters, 8 * sizeof(unsigned long long)));
	CU(ctx, cudaMemset(ctx->counters, 0, 8 * sizeof(unsigned long long)));
	ctx->traceOneStream = getenv("RPT_TRACE_ONE_STREAM") != nullptr;
	// The path tracer's tail (bounces >= WavefrontTailStart, on the tail stream): further wavefront rounds, or one kernel that runs
	// every surviving path to its end with in-line traversal.  The rounds spend a quarter of the instructions (the in-line kernel
	// runs at 3.5 lanes) but take 1.7 ms end to end against 1.1 ms.  Since the reuse passes run one frame behind the path tracer
	// the tail's latency only matters where the reuse passes are the longer chain: small strips of a multi-GPU film (4K / 8:
	// in-line 5.13 ms per strip, rounds 5.30 ms); on a 1080p film the rounds win (8.81 against 8.89 ms; profiles/r2_21_*, r2_26_*).
	// Chosen per frame by its pixel count; RPT_INLINE_TAIL=1 / RPT_WAVEFRONT_TAIL=1 force one form (A/B, tests).
	ctx->tailForm = getenv("RPT_INLINE_TAIL") ? 1 : (getenv("RPT_WAVEFRONT_TAIL") ? 2 : 0);
	ctx->spatialOneStream = getenv("RPT_SPATIAL_ONE_STREAM") != nullptr;
	ctx->noFrameOverlap = getenv("RPT_NO_FRAME_OVERLAP") != nullptr;   // A/B switch (profiles/r2_16_*)
	ctx->noShadeFromTask = getenv("RPT_NO_SHADE_FROM_TASK") != nullptr;   // A/B switch (profiles/r2_21_*)
	ctx->noReplayWavefront = getenv("RPT_NO_REPLAY_WAVEFRONT") != nullptr;   // A/B switch (profiles/r2_24_*)
	ctx->noShadeApart = getenv("RPT_NO_SHADE_APART") != nullptr;             // A/B switch (profiles/r2_28_*)
	if (const char* m = getenv("RPT_RW_MIN_LIST")) ctx->rwMinList = uint32_t(strtoul(m, nullptr, 10));
	if (const char* pm = getenv("RPT_PRIORITY_MODE")) ctx->priorityMode = atoi(pm);
	*out = ctx;
	return RPT_OK;
}

RPT_API void rpt_ctx_destroy(RptCtx* ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->counters) cudaFree(ctx->counters);
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

template <typename T>
static cudaError_t upload(RptScene* sc, const T* host, size_t n, const T** dev, cudaStream_t st) {
	void* p = nullptr;
	cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
	if (e != cudaSuccess) return e;
	sc->allocations.push_back(p);
	if (n) e = cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, st);
	*dev = static_cast<const T*>(p);
	return e;
}

static void adoptBvh(RptScene* sc, const BuildOutputs& bo) {
	sc->view.nodes = bo.nodes; sc->view.tris = bo.tris;
	sc->view.tlasNodes = nullptr; sc->view.tlasLeaves = nullptr; sc->view.instRecords = nullptr;
	sc->stats = RptBvhStats{};
	sc->stats.numTriangles = bo.numTris;
	sc->stats.numNodes = bo.numNodes;
	sc->stats.nodeBytes = uint64_t(bo.numNodes) * sizeof(WideNode);
	sc->stats.triBytes = uint64_t(bo.numTris) * sizeof(TriRecord);
	sc->stats.buildMs = bo.buildMs;
	sc->stats.sahCost = 0.0f;
}

static void adoptTwoLevel(RptScene* sc, uint32_t numMeshes) {
	const TwoLevelState& t = sc->tl;
	sc->view.nodes = t.blasNodes; sc->view.tris = t.blasTris;
	sc->view.tlasNodes = t.tlasNodes; sc->view.tlasLeaves = t.tlasLeaves; sc->view.instRecords = t.records;
	RptBvhStats& st = sc->stats;
	st = RptBvhStats{};
	st.numTriangles = t.numBlasTris; st.numNodes = t.numBlasNodes;
	st.nodeBytes = uint64_t(t.numBlasNodes + t.numTlasNodes) * sizeof(WideNode);
	st.triBytes = uint64_t(t.numBlasTris) * sizeof(TriRecord) + uint64_t(t.numRecords) * (sizeof(TriRecord) + sizeof(InstanceRecord));
	st.buildMs = t.blasMs + t.tlasMs; st.tlasBuildMs = t.tlasMs;
	st.twoLevel = 1; st.numMeshes = numMeshes; st.numTlasNodes = t.numTlasNodes; st.numInstanceRecords = t.numRecords;
}

// The arrays cross the integration boundary unchecked by anyone else (only the bundled loaders validate their own output): an
// index past its array would be an out-of-bounds device read that poisons the CUDA context.  One linear host pass.
static std::string validateDesc(const RptSceneDesc* d) {
	auto bad = [](const char* what, uint64_t i) { return std::string("rpt_scene_create: ") + what + " (element " + std::to_string(i) + ")"; };
	if ((d->numVertices && !d->vertices) || (d->numIndices && !d->indices) || (d->numMaterials && !d->materials) ||
	    (d->numMaterialIndices && !d->materialIndices) || (d->numInstances && !d->instances) || !d->triangleLights || !d->lightSampleTable ||
	    (d->numTextures && !d->textures))
		return "rpt_scene_create: an array pointer is NULL while its count is not 0";
	if (d->numMaterials == 0) return "rpt_scene_create: at least one material is needed (materials[0] is the fallback of light surfaces)";
	if (d->numMaterialIndices != d->numIndices / 3) return "rpt_scene_create: numMaterialIndices must be numIndices / 3 (one per object triangle)";
	for (uint32_t i = 0; i < d->numIndices; i++) if (d->indices[i] >= d->numVertices) return bad("an index points past the vertex array", i);
	for (uint32_t i = 0; i < d->numMaterialIndices; i++)
		if (d->materialIndices[i] < 0 || uint32_t(d->materialIndices[i]) >= d->numMaterials) return bad("a material index is out of range", i);
	for (uint32_t i = 0; i < d->numMaterials; i++)
		if (d->materials[i].textureIdx != 0xffffffffu && d->materials[i].textureIdx >= d->numTextures) return bad("a material's textureIdx is out of range", i);
	for (uint32_t i = 0; i < d->numInstances; i++) {
		const RptObjectInstance& I = d->instances[i];
		if (I.indexOffset % 3 != 0 || I.indexCount % 3 != 0 || uint64_t(I.indexOffset) + I.indexCount > d->numIndices)
			return bad("an instance's index range is not whole triangles inside the index array", i);
	}
	for (uint32_t i = 1; i <= d->numTriangleLights; i++)
		if (d->lightSampleTable[i].failId < 1 || d->lightSampleTable[i].failId > d->numTriangleLights) return bad("a light sample table failId is not in 1..N", i);
	for (uint32_t i = 0; i < d->numTextures; i++)
		if (!d->textures[i].rgba8 || d->textures[i].width == 0 || d->textures[i].height == 0) return bad("an empty texture", i);
	return std::string();
}

RPT_API int rpt_scene_create(RptCtx* ctx, const RptSceneDesc* d, RptScene** out) {
	if (!ctx || !d || !out) return fail(ctx, RPT_ERR_INVALID, "rpt_scene_create: NULL argument");
	*out = nullptr;
	if (d->numTriangleLights == 0) return fail(ctx, RPT_ERR_INVALID, "rpt_scene_create: the scene needs at least one triangle light (the light sample table divides by its total power)");
	if (d->numIndices % 3 != 0) return fail(ctx, RPT_ERR_INVALID, "rpt_scene_create: numIndices must be a multiple of 3");
	{ const std::string why = validateDesc(d); if (!why.empty()) return fail(ctx, RPT_ERR_INVALID, why); }
	CU(ctx, cudaSetDevice(ctx->device));
	RptScene* sc = new RptScene;
	sc->ctx = ctx;
	cudaStream_t st = ctx->stream;
	SceneView& v = sc->view;
	auto bail = [&](cudaError_t e, const char* what) { int r = cudaFail(ctx, e, what); rpt_scene_destroy(sc); return r; };
	cudaError_t e;
	if ((e = upload(sc, d->vertices, d->numVertices, &v.vertices, st)) != cudaSuccess) return bail(e, "upload vertices");
	if ((e = upload(sc, d->indices, d->numIndices, &v.indices, st)) != cudaSuccess) return bail(e, "upload indices");
	if ((e = upload(sc, d->materials, d->numMaterials, &v.materials, st)) != cudaSuccess) return bail(e, "upload materials");
	if ((e = upload(sc, d->materialIndices, d->numMaterialIndices, &v.materialIndices, st)) != cudaSuccess) return bail(e, "upload materialIndices");
	if ((e = upload(sc, d->instances, d->numInstances, &v.instances, st)) != cudaSuccess) return bail(e, "upload instances");
	if ((e = upload(sc, d->triangleLights, d->numTriangleLights, &v.lights, st)) != cudaSuccess) return bail(e, "upload lights");
	if ((e = upload(sc, d->lightSampleTable, size_t(d->numTriangleLights) + 1, &v.lightTable, st)) != cudaSuccess) return bail(e, "upload light table");
	v.numLights = d->numTriangleLights;

	// textures: RGBA8 sRGB texels + a 256-entry decode table (exact EOTF evaluated in double, rounded once)
	std::vector<TextureView> tv(d->numTextures);
	for (uint32_t i = 0; i < d->numTextures; i++) {
		const uchar4* texels = nullptr;
		const size_t n = size_t(d->textures[i].width) * d->textures[i].height;
		if ((e = upload(sc, reinterpret_cast<const uchar4*>(d->textures[i].rgba8), n, &texels, st)) != cudaSuccess) return bail(e, "upload texture");
		tv[i] = TextureView{ texels, d->textures[i].width, d->textures[i].height, d->textures[i].filter, 0 };
	}
	if ((e = upload(sc, tv.data(), tv.size(), &v.textures, st)) != cudaSuccess) return bail(e, "upload texture table");
	float lut[256];
	for (int i = 0; i < 256; i++) {
		const double c = i / 255.0;
		lut[i] = float(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
	}
	if ((e = upload(sc, lut, 256, &v.srgbToLinear, st)) != cudaSuccess) return bail(e, "upload sRGB table");

	// flattened triangle numbering: lights first (instance 0), then every object instance in order
	std::vector<uint32_t> triOffsets(size_t(d->numInstances) + 1);
	uint64_t total = d->numTriangleLights;
	for (uint32_t k = 0; k < d->numInstances; k++) {
		triOffsets[k] = uint32_t(total);
		total += d->instances[k].indexCount / 3;
	}
	triOffsets[d->numInstances] = uint32_t(total);
	if (total > 0x7fffffffull) { rpt_scene_destroy(sc); return fail(ctx, RPT_ERR_UNSUPPORTED, "rpt_scene_create: more than 2^31 triangles"); }
	const uint32_t* dTriOffsets = nullptr;
	if ((e = upload(sc, triOffsets.data(), triOffsets.size(), &dTriOffsets, st)) != cudaSuccess) return bail(e, "upload triOffsets");

	BuildInputs in{};
	in.vertices = v.vertices; in.indices = v.indices; in.instances = v.instances; in.lights = v.lights;
	in.triOffsets = dTriOffsets; in.numInstances = d->numInstances; in.numLights = d->numTriangleLights; in.numTris = uint32_t(total);
	sc->buildInputs = in;
	sc->geometry.resize(d->numInstances);
	for (uint32_t k = 0; k < d->numInstances; k++) sc->geometry[k] = MeshRange{ d->instances[k].indexOffset, d->instances[k].indexCount };
	sc->twoLevel = (d->flags & RPT_SCENE_TWO_LEVEL) != 0;
	if (sc->twoLevel) {
		TwoLevelInputs ti;
		ti.base = in;
		ti.meshOfInstance.resize(d->numInstances);
		std::map<std::pair<uint32_t, uint32_t>, uint32_t> seen;
		for (uint32_t k = 0; k < d->numInstances; k++) {
			const auto key = std::make_pair(d->instances[k].indexOffset, d->instances[k].indexCount);
			auto it = seen.find(key);
			if (it == seen.end()) { it = seen.emplace(key, uint32_t(ti.meshes.size())).first; ti.meshes.push_back(MeshRange{ key.first, key.second }); }
			ti.meshOfInstance[k] = it->second;
		}
		if ((e = buildTwoLevel(ti, st, &sc->tl)) != cudaSuccess) return bail(e, "buildTwoLevel");
		adoptTwoLevel(sc, uint32_t(ti.meshes.size()) + 1);
	}
	else {
		BuildOutputs bo;
		if ((e = buildBvh(in, st, &bo)) != cudaSuccess) return bail(e, "buildBvh");
		adoptBvh(sc, bo);
	}
	v.counters = nullptr;
	*out = sc;
	return RPT_OK;
}

// Dynamic scenes (the reference is static): new transforms / radiance for the object instances, geometry
// unchanged.  The single-level structure is rebuilt from scratch on the GPU — flatten, sort, PLOC, collapse: 11 ms for the
// 383 k triangles of VeachAjar — so there is no refit drift to manage.  Synchronises the device: no pass may be in flight.
// (Triangle lights are stored in world space and are not moved by this call.)
RPT_API int rpt_scene_update_instances(RptScene* s, const RptObjectInstance* instances, uint32_t numInstances) {
	if (!s || !instances) return fail(s ? s->ctx : nullptr, RPT_ERR_INVALID, "rpt_scene_update_instances: NULL argument");
	RptCtx* ctx = s->ctx;
	if (numInstances != s->buildInputs.numInstances) return fail(ctx, RPT_ERR_INVALID, "rpt_scene_update_instances: the instance count cannot change");
	for (uint32_t k = 0; k < numInstances; k++) {
		if (s->geometry[k].indexOffset != instances[k].indexOffset || s->geometry[k].indexCount != instances[k].indexCount)
			return fail(ctx, RPT_ERR_INVALID, "rpt_scene_update_instances: an instance's geometry range (indexOffset / indexCount) cannot change");
	}
	CU(ctx, cudaSetDevice(ctx->device));
	CU(ctx, cudaDeviceSynchronize());
	// the new instances go to a staged copy; the scene's own array and its structure change together, and only on success
	RptObjectInstance* staged = nullptr;
	CU(ctx, cudaMalloc(reinterpret_cast<void**>(&staged), std::max<size_t>(numInstances, 1) * sizeof(RptObjectInstance)));
	cudaError_t e = cudaMemcpy(staged, instances, size_t(numInstances) * sizeof(RptObjectInstance), cudaMemcpyHostToDevice);
	BuildInputs in = s->buildInputs;
	in.instances = staged;
	if (e == cudaSuccess && s->twoLevel) {
		// the BLASes are in object space and do not move: new instance records and a new TLAS over them
		TwoLevelState next = s->tl;
		next.tlasNodes = nullptr; next.tlasLeaves = nullptr; next.records = nullptr;
		e = cudaMalloc(reinterpret_cast<void**>(&next.records), size_t(next.numRecords) * sizeof(InstanceRecord));
		if (e == cudaSuccess) e = rebuildTlas(in, ctx->stream, &next);
		if (e != cudaSuccess) { cudaFree(next.tlasNodes); cudaFree(next.tlasLeaves); cudaFree(next.records); }
		else {
			cudaFree(s->tl.tlasNodes); cudaFree(s->tl.tlasLeaves); cudaFree(s->tl.records);
			s->tl = next;
			adoptTwoLevel(s, s->stats.numMeshes);
		}
	}
	else if (e == cudaSuccess) {
		BuildOutputs bo;
		e = buildBvh(in, ctx->stream, &bo);
		if (e == cudaSuccess) {
			cudaFree(const_cast<WideNode*>(s->view.nodes));
			cudaFree(const_cast<TriRecord*>(s->view.tris));
			adoptBvh(s, bo);
		}
	}
	if (e == cudaSuccess) {
		// the placements being replaced become "last frame's" for the G-buffer's motion vectors, until rpt_scene_end_motion
		if (!s->prevInstances) e = cudaMalloc(reinterpret_cast<void**>(&s->prevInstances), std::max<size_t>(numInstances, 1) * sizeof(RptObjectInstance));
		if (e == cudaSuccess) e = cudaMemcpy(s->prevInstances, s->view.instances, size_t(numInstances) * sizeof(RptObjectInstance), cudaMemcpyDeviceToDevice);
		if (e == cudaSuccess) s->view.prevInstances = s->prevInstances;
	}
	if (e == cudaSuccess) e = cudaMemcpy(const_cast<RptObjectInstance*>(s->view.instances), staged, size_t(numInstances) * sizeof(RptObjectInstance), cudaMemcpyDeviceToDevice);
	cudaFree(staged);
	if (e != cudaSuccess) return cudaFail(ctx, e, "rpt_scene_update_instances: rebuild");
	return RPT_OK;
}

RPT_API int rpt_scene_end_motion(RptScene* s) {
	if (!s) return fail(nullptr, RPT_ERR_INVALID, "rpt_scene_end_motion: NULL scene");
	s->view.prevInstances = nullptr;   // (kernel parameters are by value: frames already enqueued keep the view they were given)
	return RPT_OK;
}

RPT_API void rpt_scene_destroy(RptScene* s) {
	if (!s) return;
	cudaSetDevice(s->ctx->device);
	if (s->prevInstances) cudaFree(s->prevInstances);
	for (void* p : s->allocations) cudaFree(p);
	if (s->twoLevel) s->tl.release();
	else {
		if (s->view.nodes) cudaFree(const_cast<WideNode*>(s->view.nodes));
		if (s->view.tris) cudaFree(const_cast<TriRecord*>(s->view.tris));
	}
	delete s;
}

RPT_API int rpt_scene_bvh_stats(const RptScene* s, RptBvhStats* out) {
	if (!s || !out) return fail(nullptr, RPT_ERR_INVALID, "rpt_scene_bvh_stats: NULL argument");
	*out = s->stats;
	return RPT_OK;
}

// ---- frames -------------------------------------------------------------------------------------------------
RPT_API size_t rpt_buffer_stride(RptBufferId id) {
	switch (id) {
	case RPT_BUF_DIRECT_OUTPUT: case RPT_BUF_INDIRECT_OUTPUT: case RPT_BUF_DEPTH_NORMAL: case RPT_BUF_DEPTH_NORMAL_PREV: return 16;
	case RPT_BUF_ALBEDO_MATID: case RPT_BUF_ALBEDO_MATID_PREV: case RPT_BUF_MOTION: return 8;
	case RPT_BUF_DI_THIS: case RPT_BUF_DI_PREV: case RPT_BUF_DI_TEMP: return sizeof(RptDIReservoir);
	case RPT_BUF_GI_THIS: case RPT_BUF_GI_PREV: return sizeof(RptGIReservoir);
	case RPT_BUF_GRIS_THIS: case RPT_BUF_GRIS_PREV: case RPT_BUF_GRIS_TEMP: return sizeof(RptGRISReservoir);
	case RPT_BUF_PRIMARY_ISEC: return sizeof(RptIntersection);
	default: return 0;
	}
}

static void* framePtr(RptFrame* f, RptBufferId id) {
	const uint32_t c = f->c2(), p = f->p2(), c3 = f->c3(), p3 = f->p3();
	switch (id) {
	case RPT_BUF_DIRECT_OUTPUT: return f->directOutput;
	case RPT_BUF_INDIRECT_OUTPUT: return f->indirectOutput;
	case RPT_BUF_DEPTH_NORMAL: return f->depthNormal[c3];
	case RPT_BUF_DEPTH_NORMAL_PREV: return f->depthNormal[p3];
	case RPT_BUF_ALBEDO_MATID: return f->albedoMatId[c3];
	case RPT_BUF_ALBEDO_MATID_PREV: return f->albedoMatId[p3];
	case RPT_BUF_MOTION: return f->motion[c3];
	case RPT_BUF_DI_THIS: return f->di[c];
	case RPT_BUF_DI_PREV: return f->di[p];
	case RPT_BUF_DI_TEMP: return f->diTemp;
	case RPT_BUF_GI_THIS: return f->gi[c];
	case RPT_BUF_GI_PREV: return f->gi[p];
	case RPT_BUF_GRIS_THIS: return f->gris[c3];
	case RPT_BUF_GRIS_PREV: return f->gris[p3];
	case RPT_BUF_GRIS_TEMP: return f->grisTemp;
	case RPT_BUF_PRIMARY_ISEC: return f->primaryIsec;
	default: return nullptr;
	}
}

struct FrameSlot { void** ptr; size_t stride; bool wrapRows; };
// (the depthNormal images carry two extra rows: film rows 0 and H-1 for REPEAT-wrapped taps of a strip)
static std::vector<FrameSlot> frameSlots(RptFrame* f) {
	return { { (void**)&f->directOutput, 16, false }, { (void**)&f->indirectOutput, 16, false },
	         { (void**)&f->depthNormal[0], 16, true }, { (void**)&f->depthNormal[1], 16, true }, { (void**)&f->depthNormal[2], 16, true },
	         { (void**)&f->albedoMatId[0], 8, false }, { (void**)&f->albedoMatId[1], 8, false }, { (void**)&f->albedoMatId[2], 8, false },
	         { (void**)&f->motion[0], 8, false }, { (void**)&f->motion[1], 8, false }, { (void**)&f->motion[2], 8, false },
	         { (void**)&f->di[0], 64, false }, { (void**)&f->di[1], 64, false }, { (void**)&f->diTemp, 64, false },
	         { (void**)&f->gi[0], 48, false }, { (void**)&f->gi[1], 48, false },
	         { (void**)&f->gris[0], 96, false }, { (void**)&f->gris[1], 96, false }, { (void**)&f->gris[2], 96, false }, { (void**)&f->grisTemp, 96, false },
	         { (void**)&f->primaryIsec, 16, false }, { (void**)&f->rgba8, 4, false } };
}
static size_t slotBytes(const RptFrame* f, const FrameSlot& s) {
	return (f->pixels() + (s.wrapRows ? 2 * size_t(f->width) : 0)) * s.stride;
}
struct WavefrontSlot { void** ptr; size_t bytes; };
static std::vector<WavefrontSlot> wavefrontSlots(RptFrame* f) {
	const size_t px = size_t(f->width) * (f->rowEnd - f->rowBegin);
	std::vector<WavefrontSlot> v;
	for (WavefrontView& w : f->wfSet) {
		v.insert(v.end(), { { (void**)&w.state[0], px * PathStateWords * 16 }, { (void**)&w.state[1], px * PathStateWords * 16 }, { (void**)&w.cold, f->pixels() * 32 },
		                    { (void**)&w.rays[0], px * 32 }, { (void**)&w.rays[1], px * 32 }, { (void**)&w.pix[0], px * 4 }, { (void**)&w.pix[1], px * 4 }, { (void**)&w.hits, px * 16 },
		                    { (void**)&w.shadowRays[0], px * 32 }, { (void**)&w.shadowRays[1], px * 32 }, { (void**)&w.occluded[0], px }, { (void**)&w.occluded[1], px },
		                    { (void**)&w.counters, size_t(WavefrontMaxBounces) * 4 * sizeof(uint32_t) }, { (void**)&w.tailMark, f->pixels() * 4 }, { (void**)&w.tailList, px * 4 } });
	}
	v.insert(v.end(), { { (void**)&f->ru.task, px * 3 * ShiftTaskWords * 16 }, { (void**)&f->ru.rays, px * 3 * 32 }, { (void**)&f->ru.occluded, px * 3 }, { (void**)&f->ru.shadeList, px * 3 * 4 },
	                    { (void**)&f->ru.redoList, px * 4 }, { (void**)&f->ru.counters, 16 * sizeof(uint32_t) },
	                    { (void**)&f->ru.rwRays[0], px * 32 }, { (void**)&f->ru.rwRays[1], px * 32 }, { (void**)&f->ru.rwState[0], px * 32 }, { (void**)&f->ru.rwState[1], px * 32 },
	                    { (void**)&f->ru.rwHits, px * 16 }, { (void**)&f->ru.rwRc, px * 48 }, { (void**)&f->ru.rwList, px * 4 }, { (void**)&f->ru.rwCounters, 64 * sizeof(uint32_t) } });
	return v;
}

RPT_API int rpt_frame_clear(RptFrame* f) {
	if (!f) return fail(nullptr, RPT_ERR_INVALID, "rpt_frame_clear: NULL frame");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	joinTail(f);
	joinLate(f);
	for (const FrameSlot& sl : frameSlots(f)) CU(f->ctx, cudaMemsetAsync(*sl.ptr, 0, slotBytes(f, sl), f->stream));
	f->flips = 0;
	return RPT_OK;
}

RPT_API int rpt_frame_create(RptCtx* ctx, uint32_t fullWidth, uint32_t fullHeight, uint32_t rowBegin, uint32_t rowEnd, uint32_t halo, RptFrame** out) {
	if (!ctx || !out) return fail(ctx, RPT_ERR_INVALID, "rpt_frame_create: NULL argument");
	*out = nullptr;
	if (fullWidth == 0 || fullHeight == 0 || rowBegin >= rowEnd || rowEnd > fullHeight || fullWidth > 16384 || fullHeight > 16384) {
		return fail(ctx, RPT_ERR_INVALID, "rpt_frame_create: bad film size or row range");
	}
	CU(ctx, cudaSetDevice(ctx->device));
	RptFrame* f = new RptFrame;
	f->ctx = ctx;
	f->width = fullWidth; f->height = fullHeight; f->rowBegin = rowBegin; f->rowEnd = rowEnd;
	f->storeBegin = rowBegin > halo ? rowBegin - halo : 0;
	f->storeEnd = std::min(fullHeight, rowEnd + halo);
	f->halo = halo;
	f->wavefrontTail = ctx->tailForm == 2 || (ctx->tailForm == 0 && size_t(fullWidth) * (rowEnd - rowBegin) >= 1500000u);
	int prLo = 0, prHi = 0;
	cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
	const int pm = ctx->priorityMode;
	const int prMain = pm == 2 ? prHi : prLo, prPtSide = pm == 1 ? prLo : prHi, prTail = pm == 1 ? prLo : prHi, prLate = pm == 0 ? prHi : prLo;
	cudaError_t e = cudaStreamCreateWithPriority(&f->stream, cudaStreamNonBlocking, prMain);
	if (e != cudaSuccess) { delete f; return cudaFail(ctx, e, "cudaStreamCreate"); }
	for (const FrameSlot& sl : frameSlots(f)) {
		e = cudaMalloc(sl.ptr, slotBytes(f, sl));
		if (e != cudaSuccess) { rpt_frame_destroy(f); return cudaFail(ctx, e, "cudaMalloc frame buffer"); }
	}
	e = cudaMalloc(&f->flags, PeerFlagCount * sizeof(uint32_t));
	if (e == cudaSuccess) e = cudaMemset(f->flags, 0, PeerFlagCount * sizeof(uint32_t));
	if (e != cudaSuccess) { rpt_frame_destroy(f); return cudaFail(ctx, e, "cudaMalloc peer flags"); }
	if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&f->hostError), sizeof(uint32_t), cudaHostAllocMapped);
	if (e == cudaSuccess) { *f->hostError = 0u; e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&f->hostErrorDev), f->hostError, 0); }
	if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&f->hostReuseCounters), 16 * sizeof(uint32_t), cudaHostAllocDefault);
	if (e == cudaSuccess) std::memset(f->hostReuseCounters, 0, 16 * sizeof(uint32_t));
	if (e != cudaSuccess) { rpt_frame_destroy(f); return cudaFail(ctx, e, "cudaHostAlloc peer error word"); }
	e = cudaMalloc(&f->work, WorkCounterCount * sizeof(uint32_t));
	if (e == cudaSuccess) e = cudaMemset(f->work, 0, WorkCounterCount * sizeof(uint32_t));
	if (e != cudaSuccess) { rpt_frame_destroy(f); return cudaFail(ctx, e, "cudaMalloc work counters"); }
	{
		const size_t px = size_t(f->width) * (f->rowEnd - f->rowBegin);
		for (const WavefrontSlot& ws : wavefrontSlots(f)) {
			e = cudaMalloc(ws.ptr, ws.bytes);
			if (e != cudaSuccess) { rpt_frame_destroy(f); return cudaFail(ctx, e, "cudaMalloc wavefront buffer"); }
		}
		f->ru.capacity = uint32_t(px);
		for (WavefrontView& w : f->wfSet) {
			w.capacity = uint32_t(px);
			if (e == cudaSuccess) e = cudaMemset(w.tailMark, 0, f->pixels() * 4);
		}
		if (e == cudaSuccess) {   // highest priority: its small kernels must slip in between the blocks of the big pass on `stream`
			e = cudaStreamCreateWithPriority(&f->tailStream, cudaStreamNonBlocking, prTail);
		}
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->tailFork, cudaEventDisableTiming);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->tailDone, cudaEventDisableTiming);
		if (e != cudaSuccess) { rpt_frame_destroy(f); return cudaFail(ctx, e, "tail stream"); }
		{
			e = cudaStreamCreateWithPriority(&f->lateStream, cudaStreamNonBlocking, prLate);
			if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&f->lateSide, cudaStreamNonBlocking, prLate);
			if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&f->lateSide2, cudaStreamNonBlocking, prLate);
			if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&f->ptSide, cudaStreamNonBlocking, prPtSide);
			for (cudaEvent_t* ev : { &f->lateFork, &f->lateDone, &f->lateSideFork, &f->lateSideDone, &f->lateSide2Done, &f->shadeFork, &f->shadeDone, &f->lateHead, &f->lateFrameDone[0], &f->lateFrameDone[1], &f->ptFork, &f->ptJoin })
				if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
			if (e != cudaSuccess) { rpt_frame_destroy(f); return cudaFail(ctx, e, "late stream"); }
		}
	}
	int r = rpt_frame_clear(f);
	if (r != RPT_OK) { rpt_frame_destroy(f); return r; }
	CU(ctx, cudaStreamSynchronize(f->stream));
	*out = f;
	return RPT_OK;
}

static void disconnectPeers(RptFrame* f) {
	for (RptFrame::Peer* p : { &f->up, &f->down }) {
		if (p->connected && p->ipc) {
			for (void* q : { (void*)p->grisTemp, (void*)p->diTemp, (void*)p->flags, (void*)p->gris[0], (void*)p->gris[1], (void*)p->gris[2],
			                 (void*)p->di[0], (void*)p->di[1], (void*)p->gi[0], (void*)p->gi[1] })
				if (q) cudaIpcCloseMemHandle(q);
		}
		*p = RptFrame::Peer{};
	}
	// a timed-out hand-over is forgiven with the connection that caused it
	if (f->hostError) *f->hostError = 0u;
	if (f->flags) cudaMemset(f->flags, 0, PeerFlagCount * sizeof(uint32_t));
	f->grisEpoch = f->diEpoch = f->giEpoch = 0;
}
static void disconnectGather(RptFrame* f) {
	if (f->gather.connected && f->gather.ipc) { cudaIpcCloseMemHandle(f->gather.image); cudaIpcCloseMemHandle(f->gather.flags); }
	f->gather = RptFrame::Gather{};
}

RPT_API void rpt_frame_destroy(RptFrame* f) {
	if (!f) return;
	cudaSetDevice(f->ctx->device);
	if (f->tailStream) cudaStreamSynchronize(f->tailStream);
	if (f->lateSide2) cudaStreamSynchronize(f->lateSide2);
	if (f->lateSide) cudaStreamSynchronize(f->lateSide);
	if (f->lateStream) cudaStreamSynchronize(f->lateStream);
	if (f->ptSide) cudaStreamSynchronize(f->ptSide);
	if (f->stream) cudaStreamSynchronize(f->stream);
	for (cudaEvent_t ev : { f->lateFork, f->lateDone, f->lateSideFork, f->lateSideDone, f->lateSide2Done, f->shadeFork, f->shadeDone, f->lateHead, f->lateFrameDone[0], f->lateFrameDone[1], f->ptFork, f->ptJoin }) if (ev) cudaEventDestroy(ev);
	if (f->ptSide) cudaStreamDestroy(f->ptSide);
	if (f->lateSide2) cudaStreamDestroy(f->lateSide2);
	if (f->lateSide) cudaStreamDestroy(f->lateSide);
	if (f->lateStream) cudaStreamDestroy(f->lateStream);
	if (f->tailFork) cudaEventDestroy(f->tailFork);
	if (f->tailDone) cudaEventDestroy(f->tailDone);
	if (f->tailStream) cudaStreamDestroy(f->tailStream);
	disconnectPeers(f);
	disconnectGather(f);
	if (f->gatherImageOwned) cudaFree(f->gatherImageOwned);
	if (f->gatherFlagsOwned) cudaFree(f->gatherFlagsOwned);
	if (f->hostError) cudaFreeHost(f->hostError);
	if (f->hostReuseCounters) cudaFreeHost(f->hostReuseCounters);
	if (f->copyStream) { cudaStreamSynchronize(f->copyStream); cudaStreamDestroy(f->copyStream); }
	for (int i = 0; i < RptFrame::ReadbackDepth; i++) { if (f->postDone[i]) cudaEventDestroy(f->postDone[i]); if (f->copyDone[i]) cudaEventDestroy(f->copyDone[i]); }
	for (uchar4* img : f->rgba8Ring) if (img) cudaFree(img);
	for (const FrameSlot& sl : frameSlots(f)) if (*sl.ptr) cudaFree(*sl.ptr);
	if (f->flags) cudaFree(f->flags);
	if (f->work) cudaFree(f->work);
	for (const WavefrontSlot& ws : wavefrontSlots(f)) if (*ws.ptr) cudaFree(*ws.ptr);
	drainTiming(f);
	for (cudaEvent_t e : f->eventPool) cudaEventDestroy(e);
	if (f->stream) cudaStreamDestroy(f->stream);
	delete f;
}

RPT_API int rpt_frame_flip(RptFrame* f) {
	if (!f) return fail(nullptr, RPT_ERR_INVALID, "rpt_frame_flip: NULL frame");
	f->flips++;
	return RPT_OK;
}

RPT_API void* rpt_frame_stream(RptFrame* f) { return f ? (void*)f->stream : nullptr; }

RPT_API int rpt_set_camera(RptFrame* f, const RptCamera* cur, const RptCamera* prev) {
	if (!f || !cur || !prev) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_set_camera: NULL argument");
	if (cur->filmSize[0] != f->width || cur->filmSize[1] != f->height) {
		return fail(f->ctx, RPT_ERR_INVALID, "rpt_set_camera: camera filmSize does not match the frame");
	}
	// the reference memcpy's both cameras into a host-visible UBO; here they travel as kernel parameters
	f->camera = *cur;
	f->prevCamera = *prev;
	return RPT_OK;
}

static FrameView makeView(RptFrame* f) {
	FrameView v{};
	const uint32_t c = f->c2(), p = f->p2(), c3 = f->c3(), p3 = f->p3();
	v.width = f->width; v.height = f->height; v.rowBegin = f->rowBegin; v.rowEnd = f->rowEnd;
	v.storeBegin = f->storeBegin; v.storeEnd = f->storeEnd;
	v.directOutput = f->directOutput; v.indirectOutput = f->indirectOutput;
	v.depthNormal = f->depthNormal[c3]; v.depthNormalPrev = f->depthNormal[p3];
	v.albedoMatId = f->albedoMatId[c3]; v.albedoMatIdPrev = f->albedoMatId[p3];
	v.motion = f->motion[c3];
	v.diThis = f->di[c]; v.diPrev = f->di[p]; v.diTemp = f->diTemp;
	v.giThis = f->gi[c]; v.giPrev = f->gi[p];
	v.grisThis = f->gris[c3]; v.grisPrev = f->gris[p3]; v.grisTemp = f->grisTemp;
	v.grisStale = f->gris[f->n3()];   // (what a ping-pong pair would still hold in "this": the reservoirs of two frames ago)
	v.primaryIsec = f->primaryIsec;
	v.camera = f->camera; v.prevCamera = f->prevCamera;
	v.halo = f->halo;
	v.work = f->work;
	v.wf = f->wf();
	v.ru = f->ru;
	v.ru.noShadeFromTask = f->ctx->noShadeFromTask ? 1u : 0u;
	v.ru.noReplayWavefront = f->ctx->noReplayWavefront ? 1u : 0u;
	v.striped = f->rowBegin != 0 || f->rowEnd != f->height;
	v.peerGrisUp = f->up.connected ? f->up.grisTemp : nullptr;
	v.peerDiUp = f->up.connected ? f->up.diTemp : nullptr;
	v.peerUpStoreBegin = f->up.storeBegin;
	v.peerGrisDown = f->down.connected ? f->down.grisTemp : nullptr;
	v.peerDiDown = f->down.connected ? f->down.diTemp : nullptr;
	v.peerDownStoreBegin = f->down.storeBegin;
	// the neighbours' final-reservoir buffers of this frame, and the rows whose previous-frame state is therefore valid here:
	// a connected side contributes its halo rows except the outermost one (the bilinear G-buffer tap of a lookup in row y also
	// reads row y-1 or y+1; at a film edge the wrap rows are stored, see depthNormalRow)
	v.peerGrisThisUp = f->up.connected ? f->up.gris[c3] : nullptr;    v.peerGrisThisDown = f->down.connected ? f->down.gris[c3] : nullptr;
	v.peerDiThisUp = f->up.connected ? f->up.di[c] : nullptr;         v.peerDiThisDown = f->down.connected ? f->down.di[c] : nullptr;
	v.peerGiThisUp = f->up.connected ? f->up.gi[c] : nullptr;         v.peerGiThisDown = f->down.connected ? f->down.gi[c] : nullptr;
	v.prevRowBegin = f->up.connected ? (f->storeBegin == 0 ? 0 : f->storeBegin + 1) : f->rowBegin;
	v.prevRowEnd = f->down.connected ? (f->storeEnd == f->height ? f->height : f->storeEnd - 1) : f->rowEnd;
	v.gatherImage = f->gather.connected ? f->gather.image : nullptr;
	return v;
}

static SceneView sceneView(const RptScene* s) {
	SceneView v = s->view;
	v.counters = s->ctx->countersOn ? s->ctx->counters : nullptr;
	return v;
}

#define PASS_PROLOGUE_EARLY(name) \
	if (!f || !s) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, name ": NULL argument"); \
	if (f->ctx != s->ctx) return fail(f->ctx, RPT_ERR_INVALID, name ": frame and scene belong to different contexts"); \
	if (f->hostError && *reinterpret_cast<volatile uint32_t*>(f->hostError)) return fail(f->ctx, RPT_ERR_PEER, name ": a multi-GPU hand-over of an earlier pass timed out on the device (a neighbouring strip died or was not driven in lock step); disconnect the peers to recover"); \
	CU(f->ctx, cudaSetDevice(f->ctx->device)); \
	joinTail(f);
// every pass but the G-buffer and the ReSTIR PT path tracer (PASS_PROLOGUE_EARLY + waitSlotReaders) waits for the late passes in flight
#define PASS_PROLOGUE(name) \
	PASS_PROLOGUE_EARLY(name) \
	joinLate(f);
#define PASS_EPILOGUE(name) \
	CU(f->ctx, cudaGetLastError()); \
	return RPT_OK;

#define SIMPLE_PASS(fn, passId, launcher) \
	RPT_API int fn(RptFrame* f, const RptScene* s) { \
		PASS_PROLOGUE(#fn) \
		{ PassTimer timer(f, passId); launcher(makeView(f), sceneView(s), f->stream); } \
		PASS_EPILOGUE(#fn) \
	}
// (writes the other half of the G-buffer ping-pong pair + the motion vectors, which only a temporal pass reads: it may run next
// to the previous frame's late passes)
RPT_API int rpt_gbuffer(RptFrame* f, const RptScene* s) {
	PASS_PROLOGUE_EARLY("rpt_gbuffer")
	waitSlotReaders(f);
	{ PassTimer timer(f, RPT_PASS_GBUFFER); launchGBuffer(makeView(f), sceneView(s), f->stream); }
	PASS_EPILOGUE("rpt_gbuffer")
}
SIMPLE_PASS(rpt_di_naive, RPT_PASS_DI_NAIVE, launchDINaive)
SIMPLE_PASS(rpt_di_naive_rt, RPT_PASS_DI_NAIVE, launchDINaiveRT)
SIMPLE_PASS(rpt_gi_naive, RPT_PASS_GI_NAIVE, launchGINaive)
static void peerBefore(RptFrame* f, int h);
static void peerAfter(RptFrame* f, int h);
enum PeerHook { HookNone, HookGrisTemporal, HookGrisSpatial, HookDiTemporal, HookDiSpatial, HookGi };
RPT_API int rpt_gi_restir(RptFrame* f, const RptScene* s) {
	PASS_PROLOGUE("rpt_gi_restir")
	peerBefore(f, HookGi);
	{
		PassTimer timer(f, RPT_PASS_GI_RESTIR);
		const bool overlap = !f->ctx->traceOneStream;   // A/B switch
		launchGIReSTIR(makeView(f), sceneView(s), f->stream, overlap ? f->tailStream : nullptr, f->tailFork, f->tailDone);
	}
	peerAfter(f, HookGi);
	PASS_EPILOGUE("rpt_gi_restir")
}
SIMPLE_PASS(rpt_visualize_as, RPT_PASS_VISUALIZE_AS, launchVisualizeAS)

// hand-over hooks around the temporal / spatial passes of a striped frame (no-ops without connected peers)
static void peerBefore(RptFrame* f, int h) {
	if (!f->up.connected && !f->down.connected) return;
	uint32_t* err = f->flags + PeerError;
	const uint32_t* fromUp = f->up.connected ? f->flags : nullptr;
	const uint32_t* fromDown = f->down.connected ? f->flags : nullptr;
	auto wait = [&](int upSlot, int downSlot, uint32_t epoch) {
		launchPeerWait(fromUp ? fromUp + upSlot : nullptr, fromDown ? fromDown + downSlot : nullptr, epoch, err, f->hostErrorDev, f->stream);
	};
	switch (h) {
	case HookGrisTemporal:   // neighbours must have finished last frame's spatial pass: they no longer read the halo rows the temporal
		f->grisEpoch++;      // pass is about to overwrite, and the mirrored rows of their final reservoirs have landed in ours
		wait(GrisSpatialFromUp, GrisSpatialFromDown, f->grisEpoch - 1);
		break;
	case HookGrisSpatial:    // neighbours' boundary rows of this frame must have landed in our halo rows
		wait(GrisTemporalFromUp, GrisTemporalFromDown, f->grisEpoch);
		break;
	case HookDiTemporal:
		f->diEpoch++;
		wait(DiSpatialFromUp, DiSpatialFromDown, f->diEpoch - 1);
		break;
	case HookDiSpatial:
		wait(DiTemporalFromUp, DiTemporalFromDown, f->diEpoch);
		break;
	case HookGi:             // one pass per frame: the neighbours' previous frame must be complete — its mirrored rows are what this
		f->giEpoch++;        // pass reads as history, and this pass mirrors into rows the neighbours read as history last frame
		wait(GiFromUp, GiFromDown, f->giEpoch - 1);
		break;
	default: break;
	}
}
static void peerAfter(RptFrame* f, int h) {
	if (!f->up.connected && !f->down.connected) return;
	// we are the "down" neighbour of the strip above and the "up" neighbour of the strip below
	uint32_t* toUp = f->up.connected ? f->up.flags : nullptr;
	uint32_t* toDown = f->down.connected ? f->down.flags : nullptr;
	auto signal = [&](int slotInUp, int slotInDown, uint32_t epoch) {
		launchPeerSignal(toUp ? toUp + slotInUp : nullptr, toDown ? toDown + slotInDown : nullptr, epoch, f->stream);
	};
	switch (h) {
	case HookGrisTemporal: signal(GrisTemporalFromDown, GrisTemporalFromUp, f->grisEpoch); break;
	case HookGrisSpatial: signal(GrisSpatialFromDown, GrisSpatialFromUp, f->grisEpoch); break;
	case HookDiTemporal: signal(DiTemporalFromDown, DiTemporalFromUp, f->diEpoch); break;
	case HookDiSpatial: signal(DiSpatialFromDown, DiSpatialFromUp, f->diEpoch); break;
	case HookGi: signal(GiFromDown, GiFromUp, f->giEpoch); break;
	default: break;
	}
}

#define SETTINGS_PASS(fn, T, passId, launcher, hook) \
	RPT_API int fn(RptFrame* f, const RptScene* s, const T* st) { \
		PASS_PROLOGUE(#fn) \
		if (!st) return fail(f->ctx, RPT_ERR_INVALID, #fn ": NULL settings"); \
		peerBefore(f, hook); \
		{ PassTimer timer(f, passId); launcher(makeView(f), sceneView(s), *st, f->stream); } \
		peerAfter(f, hook); \
		PASS_EPILOGUE(#fn) \
	}
SETTINGS_PASS(rpt_di_pathgen, RptDISettings, RPT_PASS_DI_PATHGEN, launchDIPathGen, HookNone)
SETTINGS_PASS(rpt_di_temporal, RptDISettings, RPT_PASS_DI_TEMPORAL, launchDITemporal, HookDiTemporal)
SETTINGS_PASS(rpt_di_spatial, RptDISettings, RPT_PASS_DI_SPATIAL, launchDISpatial, HookDiSpatial)

// GRISReSTIR::render step 1.  Bounces 0..WavefrontTailStart-1 (all but a few percent of the rays) run on the frame's
// stream, with a side stream of their own for the paired any-hit launches; the long tail of the few paths that live on is
// enqueued on the tail stream and joined by the temporal pass (its pixels' temporal step follows it there).
RPT_API int rpt_gris_pathtrace(RptFrame* f, const RptScene* s, const RptGRISSettings* st) {
	PASS_PROLOGUE_EARLY("rpt_gris_pathtrace")   // (reads the new G-buffer, writes the reservoirs the previous frame read as history)
	if (!st) return fail(f->ctx, RPT_ERR_INVALID, "rpt_gris_pathtrace: NULL settings");
	waitSlotReaders(f);
	f->lastWfSet = f->flips & 1u;
	f->wf().epoch++;
	const FrameView view = makeView(f);
	const SceneView scene = sceneView(s);
	{
		PassTimer timer(f, RPT_PASS_GRIS_PATHTRACE);
		FrameKernelClock clock(f);
		const bool overlap = !f->ctx->traceOneStream;   // A/B switch (profiles/r1_18_*)
		launchGRISPathTraceBounces(view, scene, *st, 0, WavefrontTailStart - 1, f->stream, f->timing ? &clock : nullptr,
		                           overlap ? f->ptSide : nullptr, f->ptFork, f->ptJoin);
	}
	CU(f->ctx, cudaEventRecord(f->tailFork, f->stream));
	CU(f->ctx, cudaStreamWaitEvent(f->tailStream, f->tailFork, 0));
	cudaEvent_t t0 = nullptr, t1 = nullptr;
	if (f->timing) { t0 = takeEvent(f); t1 = takeEvent(f); cudaEventRecord(t0, f->tailStream); }
	if (f->wavefrontTail) launchGRISPathTraceBounces(view, scene, *st, WavefrontTailStart, 15, f->tailStream);   // (form by frame size: rpt_ctx_create)
	else launchGRISPathTraceTail(view, scene, *st, f->tailStream);
	if (f->timing) { cudaEventRecord(t1, f->tailStream); f->pending.push_back({ RPT_PASS_COUNT + RPT_KERNEL_GRIS_TAIL, t0, t1, true }); }
	CU(f->ctx, cudaEventRecord(f->tailDone, f->tailStream));
	f->tailPending = true;
	PASS_EPILOGUE("rpt_gris_pathtrace")
}

// GRISReSTIR::render step 2.  While the path-tracing tail is still running, every pixel outside it is processed
// first; the tail's pixels follow once it is done.
RPT_API int rpt_gris_temporal(RptFrame* f, const RptScene* s, const RptGRISSettings* st) {
	if (!f || !s) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_gris_temporal: NULL argument");
	if (f->ctx != s->ctx) return fail(f->ctx, RPT_ERR_INVALID, "rpt_gris_temporal: frame and scene belong to different contexts");
	if (!st) return fail(f->ctx, RPT_ERR_INVALID, "rpt_gris_temporal: NULL settings");
	if (f->hostError && *reinterpret_cast<volatile uint32_t*>(f->hostError)) return fail(f->ctx, RPT_ERR_PEER, "rpt_gris_temporal: a multi-GPU hand-over of an earlier pass timed out on the device; disconnect the peers to recover");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	const FrameView view = makeView(f);
	const SceneView scene = sceneView(s);
	const bool hadTail = f->tailPending;
	if (pipelined(f) && !f->ctx->traceOneStream) {
		// The whole pass goes to the late stream set, behind the previous frame's spatial pass (which wrote this frame's history and
		// read the buffers this pass writes) and behind everything enqueued on the frame's stream so far (the path tracer).  Three
		// streams inside: the late stream carries gen / visibility rays / merge of every pixel outside the path tracer's tail; the
		// tail's pixels — independent of all others in this pass (own reservoir, own history pixel) — take their temporal step right
		// behind the tail kernel on ITS stream; the pixels whose history sample needs replay rays run as a list kernel on the late
		// set's side stream.  All three are joined here, because the hand-over signal to neighbouring strips and the spatial pass
		// (next on the late stream) need every pixel of the pass.
		LateScope late(f);
		peerBefore(f, HookGrisTemporal);
		{
			PassTimer timer(f, RPT_PASS_GRIS_TEMPORAL);
			FrameKernelClock clock(f);
			KernelClock* ck = f->timing ? &clock : nullptr;
			if (hadTail) {
				CU(f->ctx, cudaEventRecord(f->lateHead, f->stream));
				CU(f->ctx, cudaStreamWaitEvent(late.t0, f->lateHead, 0));
				launchGRISTemporal(view, scene, *st, late.t0, 2, nullptr);
				CU(f->ctx, cudaEventRecord(late.b0, late.t0));
			}
			launchGRISTemporal(view, scene, *st, f->stream, hadTail ? 1 : 0, ck, f->tailStream, f->tailFork);
			CU(f->ctx, cudaEventRecord(f->tailDone, f->tailStream));
			if (ck) ck->tick(RPT_KERNEL_TAIL_WAIT);
			CU(f->ctx, cudaStreamWaitEvent(f->stream, f->tailDone, 0));
			if (hadTail) { CU(f->ctx, cudaStreamWaitEvent(f->stream, late.b0, 0)); f->tailPending = false; }
		}
		peerAfter(f, HookGrisTemporal);
		PASS_EPILOGUE("rpt_gris_temporal")
	}
	joinLate(f);   // the previous frame's spatial pass wrote this frame's history and read the buffers this pass writes
	peerBefore(f, HookGrisTemporal);
	{
		PassTimer timer(f, RPT_PASS_GRIS_TEMPORAL);
		FrameKernelClock clock(f);
		KernelClock* ck = f->timing ? &clock : nullptr;
		if (f->ctx->traceOneStream) {
			if (hadTail) {
				launchGRISTemporal(view, scene, *st, f->stream, 1, ck);
				if (ck) ck->tick(RPT_KERNEL_TAIL_WAIT);
				joinTail(f);
				launchGRISTemporal(view, scene, *st, f->stream, 2, ck);
			}
			else launchGRISTemporal(view, scene, *st, f->stream, 0, ck);
		}
		else {
			// one frame at a time (RPT_NO_FRAME_OVERLAP): the same three streams, joined by the next pass (at once when neighbouring
			// strips are connected: the hand-over signal below must follow every pixel of the pass)
			if (hadTail) launchGRISTemporal(view, scene, *st, f->tailStream, 2, nullptr);
			cudaStream_t listStream = f->lateSide ? f->lateSide : f->tailStream;
			launchGRISTemporal(view, scene, *st, f->stream, hadTail ? 1 : 0, ck, listStream, f->tailFork);
			if (listStream != f->tailStream) {
				CU(f->ctx, cudaEventRecord(f->lateSideDone, listStream));
				CU(f->ctx, cudaStreamWaitEvent(f->tailStream, f->lateSideDone, 0));
			}
			CU(f->ctx, cudaEventRecord(f->tailDone, f->tailStream));
			f->tailPending = true;
			if (f->up.connected || f->down.connected) joinTail(f);
		}
	}
	peerAfter(f, HookGrisTemporal);
	PASS_EPILOGUE("rpt_gris_temporal")
}
RPT_API int rpt_gris_spatial(RptFrame* f, const RptScene* s, const RptGRISSettings* st) {
	if (!f || !s) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_gris_spatial: NULL argument");
	if (f->ctx != s->ctx) return fail(f->ctx, RPT_ERR_INVALID, "rpt_gris_spatial: frame and scene belong to different contexts");
	if (!st) return fail(f->ctx, RPT_ERR_INVALID, "rpt_gris_spatial: NULL settings");
	if (f->hostError && *reinterpret_cast<volatile uint32_t*>(f->hostError)) return fail(f->ctx, RPT_ERR_PEER, "rpt_gris_spatial: a multi-GPU hand-over of an earlier pass timed out on the device; disconnect the peers to recover");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	// (late stream set: in order behind the temporal pass; the frame's stream is not held up — it may already carry the next frame)
	if (!pipelined(f)) { joinTail(f); joinLate(f); }
	const bool tailOnMainSet = f->tailPending;   // a tail (or a temporal pass run on the frame's streams) not joined yet
	cudaEvent_t mainTailDone = f->tailDone;
	LateScope late(f);
	if (late.on && tailOnMainSet) { CU(f->ctx, cudaStreamWaitEvent(f->stream, mainTailDone, 0)); f->tailPending = false; }
	// (the previous frame's shade list and post-process read the lists and write the outputs this pass is about to overwrite)
	if (late.on && f->shadePending) { CU(f->ctx, cudaStreamWaitEvent(f->stream, f->shadeDone, 0)); f->shadePending = false; }
	const bool shadeApart = late.on && !f->ctx->spatialOneStream && !f->ctx->noShadeApart;
	peerBefore(f, HookGrisSpatial);
	{
		PassTimer timer(f, RPT_PASS_GRIS_SPATIAL);
		FrameKernelClock clock(f);
		const bool side = !f->ctx->spatialOneStream;   // A/B switch
		FrameView view = makeView(f);
		// The replays of this pass as a wavefront or as in-line chains?  The wavefront pays a fixed latency per bounce (its longest
		// ray), the in-line list kernel a time that grows with the list: the size of the last pass's list decides (read back
		// asynchronously — a frame or two stale, never waited for; both forms produce the same bits).
		const volatile uint32_t* hc = f->hostReuseCounters;
		if (hc[3] + hc[5] < f->ctx->rwMinList) view.ru.noReplayWavefront = 1u;
		launchGRISSpatial(view, sceneView(s), *st, f->stream, f->timing ? &clock : nullptr,
		                  side ? f->tailStream : nullptr, f->tailFork, f->tailDone, side ? f->lateSide2 : nullptr, f->lateSide2Done,
		                  shadeApart ? f->lateSide2 : nullptr, f->shadeFork, f->shadeDone);
		if (shadeApart) f->shadePending = true;
		CU(f->ctx, cudaMemcpyAsync(f->hostReuseCounters, f->ru.counters, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
	}
	peerAfter(f, HookGrisSpatial);
	PASS_EPILOGUE("rpt_gris_spatial")
}

static int postprocessInto(RptFrame* f, const RptPostSettings* st, uchar4* image) {
	joinTail(f);
	if (f->gather.connected) {   // the film image on the root must have been released by the gather of the previous frame
		f->gather.epoch++;
		launchPeerWait(f->gather.flags + GatherReleaseFlag, nullptr, f->gather.epoch - 1, f->flags + PeerError, f->hostErrorDev, f->stream);
	}
	{ PassTimer timer(f, RPT_PASS_POSTPROCESS); launchPostProcess(makeView(f), *st, image, f->stream); }
	if (f->gather.connected) launchPeerSignal(f->gather.flags + GatherArrivalFlag0 + f->gather.strip, nullptr, f->gather.epoch, f->stream);
	CU(f->ctx, cudaGetLastError());
	return RPT_OK;
}

RPT_API int rpt_postprocess(RptFrame* f, const RptPostSettings* st, uint8_t* rgba8Out) {
	if (!f || !st) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_postprocess: NULL argument");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	// device-only run after a late spatial pass: the post-process follows it there; with a read-back the streams are joined first
	struct MaybeLate { LateScope* l = nullptr; ~MaybeLate() { delete l; } } late;
	if (f->latePending && !rgba8Out) late.l = new LateScope(f);
	else joinLate(f);
	ShadeScope shade(f, late.l != nullptr && late.l->on);
	if (f->asyncTicket) for (cudaEvent_t ev : f->copyDone) CU(f->ctx, cudaStreamWaitEvent(f->stream, ev, 0));   // (f->rgba8 may still be being read back)
	const int rc = postprocessInto(f, st, f->rgba8);
	if (rc != RPT_OK) return rc;
	if (rgba8Out) {
		CU(f->ctx, cudaMemcpyAsync(rgba8Out, f->rgba8, size_t(f->width) * (f->rowEnd - f->rowBegin) * 4, cudaMemcpyDeviceToHost, f->stream));
		CU(f->ctx, cudaStreamSynchronize(f->stream));
	}
	return RPT_OK;
}

// The same pass with the read-back taken off the frame's stream: the image goes to one of two device buffers, a copy stream
// carries it to (pinned) host memory while the frame's stream is already rendering the next frame, and the caller collects it
// with rpt_readback_wait(ticket).  At most three read-backs are in flight: issuing ticket t waits (on the device) for the copy of
// ticket t - 3, which used the same device image — the caller must have collected that one, or not care about it.
RPT_API int rpt_postprocess_async(RptFrame* f, const RptPostSettings* st, uint8_t* rgba8Out, uint64_t* ticket) {
	if (!f || !st || !rgba8Out || !ticket) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_postprocess_async: NULL argument");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	const size_t bytes = size_t(f->width) * (f->rowEnd - f->rowBegin) * 4;
	if (!f->copyStream) {
		CU(f->ctx, cudaStreamCreateWithFlags(&f->copyStream, cudaStreamNonBlocking));
		for (int i = 0; i < RptFrame::ReadbackDepth; i++) {
			CU(f->ctx, cudaEventCreateWithFlags(&f->postDone[i], cudaEventDisableTiming));
			CU(f->ctx, cudaEventCreateWithFlags(&f->copyDone[i], cudaEventDisableTiming));
		}
		for (uchar4*& img : f->rgba8Ring) CU(f->ctx, cudaMalloc(reinterpret_cast<void**>(&img), std::max<size_t>(bytes, 4)));
	}
	const uint64_t t = f->asyncTicket;
	const int slot = int(t % RptFrame::ReadbackDepth);
	uchar4* image = slot ? f->rgba8Ring[slot - 1] : f->rgba8;
	// after a late spatial pass the post-process follows it there (it reads what that pass accumulated); otherwise the frame's stream
	struct MaybeLate { LateScope* l = nullptr; ~MaybeLate() { delete l; } } late;
	if (f->latePending) late.l = new LateScope(f);
	ShadeScope shade(f, late.l != nullptr && late.l->on);
	if (t >= uint64_t(RptFrame::ReadbackDepth)) CU(f->ctx, cudaStreamWaitEvent(f->stream, f->copyDone[slot], 0));
	const int rc = postprocessInto(f, st, image);
	if (rc != RPT_OK) return rc;
	CU(f->ctx, cudaEventRecord(f->postDone[slot], f->stream));
	CU(f->ctx, cudaStreamWaitEvent(f->copyStream, f->postDone[slot], 0));
	CU(f->ctx, cudaMemcpyAsync(rgba8Out, image, bytes, cudaMemcpyDeviceToHost, f->copyStream));
	CU(f->ctx, cudaEventRecord(f->copyDone[slot], f->copyStream));
	f->asyncTicket = t + 1;
	*ticket = t;
	return RPT_OK;
}

RPT_API int rpt_readback_wait(RptFrame* f, uint64_t ticket) {
	if (!f) return fail(nullptr, RPT_ERR_INVALID, "rpt_readback_wait: NULL frame");
	if (ticket >= f->asyncTicket) return fail(f->ctx, RPT_ERR_INVALID, "rpt_readback_wait: no such read-back");
	// (if its device image has been handed to a later read-back since, the event below stands for that later copy: waiting for it
	// covers this one, the copies of one image being ordered on the copy stream)
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	CU(f->ctx, cudaEventSynchronize(f->copyDone[ticket % RptFrame::ReadbackDepth]));
	return RPT_OK;
}

RPT_API int rpt_sync(RptFrame* f) {
	if (!f) return fail(nullptr, RPT_ERR_INVALID, "rpt_sync: NULL frame");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	joinTail(f);
	joinLate(f);
	CU(f->ctx, syncFrame(f));
	return RPT_OK;
}

RPT_API int rpt_frame_join(RptFrame* f) {
	if (!f) return fail(nullptr, RPT_ERR_INVALID, "rpt_frame_join: NULL frame");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	joinTail(f);
	joinLate(f);
	return RPT_OK;
}

RPT_API int rpt_frame_timing(RptFrame* f, int enable) {
	if (!f) return fail(nullptr, RPT_ERR_INVALID, "rpt_frame_timing: NULL frame");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	CU(f->ctx, syncFrame(f));
	drainTiming(f);
	f->timing = enable != 0;
	if (enable) f->stats = RptPassStats{};
	return RPT_OK;
}

RPT_API int rpt_frame_pass_stats(RptFrame* f, RptPassStats* out) {
	if (!f || !out) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_frame_pass_stats: NULL argument");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	joinTail(f);
	joinLate(f);
	CU(f->ctx, syncFrame(f));
	drainTiming(f);
	*out = f->stats;
	return RPT_OK;
}

RPT_API int rpt_frame_rows(const RptFrame* f, uint32_t* b, uint32_t* e) {
	if (!f || !b || !e) return fail(nullptr, RPT_ERR_INVALID, "rpt_frame_rows: NULL argument");
	*b = f->storeBegin; *e = f->storeEnd;
	return RPT_OK;
}

RPT_API int rpt_read(RptFrame* f, RptBufferId id, void* dst, size_t bytes) {
	if (!f || !dst) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_read: NULL argument");
	void* p = framePtr(f, id);
	if (!p || bytes != f->pixels() * rpt_buffer_stride(id)) return fail(f->ctx, RPT_ERR_INVALID, "rpt_read: bad buffer id or size");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	joinTail(f);
	joinLate(f);
	CU(f->ctx, cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, f->stream));
	CU(f->ctx, syncFrame(f));
	return RPT_OK;
}

RPT_API int rpt_write(RptFrame* f, RptBufferId id, const void* src, size_t bytes) {
	if (!f || !src) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_write: NULL argument");
	void* p = framePtr(f, id);
	if (!p || bytes != f->pixels() * rpt_buffer_stride(id)) return fail(f->ctx, RPT_ERR_INVALID, "rpt_write: bad buffer id or size");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	joinTail(f);
	joinLate(f);
	CU(f->ctx, cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, f->stream));
	CU(f->ctx, syncFrame(f));
	return RPT_OK;
}

RPT_API void* rpt_device_ptr(RptFrame* f, RptBufferId id) { return f ? framePtr(f, id) : nullptr; }

// ---- multi-GPU strips ------------------------------------------------------------------------------------------
RPT_API int rpt_frame_export_peer(RptFrame* f, RptPeerInfo* out) {
	if (!f || !out) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_frame_export_peer: NULL argument");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	std::memset(out, 0, sizeof(*out));
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	auto handle = [&](uint8_t* dst, void* p) { return cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(dst), p); };
	CU(f->ctx, handle(out->grisTempHandle, f->grisTemp));
	CU(f->ctx, handle(out->diTempHandle, f->diTemp));
	CU(f->ctx, handle(out->flagsHandle, f->flags));
	for (int i = 0; i < 3; i++) {
		CU(f->ctx, handle(out->grisHandle[i], f->gris[i]));
		out->grisPtr[i] = reinterpret_cast<uint64_t>(f->gris[i]);
	}
	for (int i = 0; i < 2; i++) {
		CU(f->ctx, handle(out->diHandle[i], f->di[i]));
		CU(f->ctx, handle(out->giHandle[i], f->gi[i]));
		out->diPtr[i] = reinterpret_cast<uint64_t>(f->di[i]);
		out->giPtr[i] = reinterpret_cast<uint64_t>(f->gi[i]);
	}
	out->grisTempPtr = reinterpret_cast<uint64_t>(f->grisTemp);
	out->diTempPtr = reinterpret_cast<uint64_t>(f->diTemp);
	out->flagsPtr = reinterpret_cast<uint64_t>(f->flags);
	out->pid = uint64_t(getpid());
	out->device = f->ctx->device;
	out->rowBegin = f->rowBegin; out->rowEnd = f->rowEnd; out->storeBegin = f->storeBegin; out->storeEnd = f->storeEnd;
	out->cur = f->flips;
	return RPT_OK;
}

static int connectOne(RptFrame* f, RptFrame::Peer& p, const RptPeerInfo* info) {
	p = RptFrame::Peer{};
	if (!info) return RPT_OK;
	void* q[10] = {};
	if (info->pid == uint64_t(getpid())) {
		// same process: plain pointers, peer access enabled when the neighbour lives on another device
		if (info->device != f->ctx->device) {
			int can = 0;
			CU(f->ctx, cudaDeviceCanAccessPeer(&can, f->ctx->device, info->device));
			if (!can) return fail(f->ctx, RPT_ERR_UNSUPPORTED, "rpt_frame_connect_peers: no peer access between the two devices");
			cudaError_t e = cudaDeviceEnablePeerAccess(info->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cudaFail(f->ctx, e, "cudaDeviceEnablePeerAccess");
			cudaGetLastError();
		}
		const uint64_t raw[10] = { info->grisTempPtr, info->diTempPtr, info->flagsPtr, info->grisPtr[0], info->grisPtr[1], info->grisPtr[2],
		                           info->diPtr[0], info->diPtr[1], info->giPtr[0], info->giPtr[1] };
		for (int i = 0; i < 10; i++) q[i] = reinterpret_cast<void*>(raw[i]);
	}
	else {
		const uint8_t* h[10] = { info->grisTempHandle, info->diTempHandle, info->flagsHandle, info->grisHandle[0], info->grisHandle[1], info->grisHandle[2],
		                         info->diHandle[0], info->diHandle[1], info->giHandle[0], info->giHandle[1] };
		for (int i = 0; i < 10; i++) {
			cudaError_t e = cudaIpcOpenMemHandle(&q[i], *reinterpret_cast<const cudaIpcMemHandle_t*>(h[i]), cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) {
				for (int k = 0; k < i; k++) cudaIpcCloseMemHandle(q[k]);
				return cudaFail(f->ctx, e, "cudaIpcOpenMemHandle");
			}
		}
		p.ipc = true;
	}
	p.grisTemp = static_cast<RptGRISReservoir*>(q[0]); p.diTemp = static_cast<RptDIReservoir*>(q[1]); p.flags = static_cast<uint32_t*>(q[2]);
	// index the neighbour's buffers by OUR phases: both frames flip once per frame from here on
	const uint32_t phase2 = (info->cur ^ f->flips) & 1u;
	const uint32_t phase3 = (info->cur % 3u + 3u - f->flips % 3u) % 3u;   // its slot = (ours + phase3) mod 3
	for (uint32_t i = 0; i < 3; i++) p.gris[i] = static_cast<RptGRISReservoir*>(q[3 + (i + phase3) % 3u]);
	for (uint32_t i = 0; i < 2; i++) {
		p.di[i] = static_cast<RptDIReservoir*>(q[6 + (i ^ phase2)]);
		p.gi[i] = static_cast<RptGIReservoir*>(q[8 + (i ^ phase2)]);
	}
	p.storeBegin = info->storeBegin;
	p.connected = true;
	return RPT_OK;
}

// All strips of a film connect between the same two frames, and a barrier (or, in one process, program order) separates the
// connection from the first pass: the epoch counters and flag words of the hand-over restart at zero here on every strip, so
// frames that rendered different numbers of frames before (re-partitioning, a replaced GPU) start in step again.
RPT_API int rpt_frame_connect_peers(RptFrame* f, const RptPeerInfo* up, const RptPeerInfo* down) {
	if (!f) return fail(nullptr, RPT_ERR_INVALID, "rpt_frame_connect_peers: NULL frame");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	if (up && (up->rowEnd != f->rowBegin || up->storeEnd < f->rowBegin + f->halo)) return fail(f->ctx, RPT_ERR_INVALID, "rpt_frame_connect_peers: `up` is not the strip directly above with a matching halo");
	if (down && (down->rowBegin != f->rowEnd || down->storeBegin + f->halo > f->rowEnd)) return fail(f->ctx, RPT_ERR_INVALID, "rpt_frame_connect_peers: `down` is not the strip directly below with a matching halo");
	// a neighbour fills our halo rows with ITS rows only: it must own all of them (or reach the film edge)
	if (up && up->rowBegin != 0 && up->rowEnd - up->rowBegin < f->halo) return fail(f->ctx, RPT_ERR_INVALID, "rpt_frame_connect_peers: the strip above is shorter than the halo");
	if (down && down->rowEnd != f->height && down->rowEnd - down->rowBegin < f->halo) return fail(f->ctx, RPT_ERR_INVALID, "rpt_frame_connect_peers: the strip below is shorter than the halo");
	joinTail(f);
	joinLate(f);
	CU(f->ctx, syncFrame(f));
	disconnectPeers(f);   // also: epochs, flag words and the sticky error back to zero
	int r = connectOne(f, f->up, up);
	if (r != RPT_OK) return r;
	r = connectOne(f, f->down, down);
	if (r != RPT_OK) disconnectPeers(f);
	return r;
}

RPT_API int rpt_frame_disconnect_peers(RptFrame* f) {
	if (!f) return fail(nullptr, RPT_ERR_INVALID, "rpt_frame_disconnect_peers: NULL frame");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	joinTail(f);
	joinLate(f);
	CU(f->ctx, syncFrame(f));
	disconnectPeers(f);
	return RPT_OK;
}

RPT_API int rpt_frame_peers_in_process(const RptFrame* f) {
	if (!f) return 0;
	return ((f->up.connected && !f->up.ipc) || (f->down.connected && !f->down.ipc)) ? 1 : 0;
}

// ---- final image gather ------------------------------------------------------------------------------------------
RPT_API int rpt_frame_gather_create(RptFrame* f, uint32_t numStrips, RptGatherInfo* out) {
	if (!f || !out || numStrips == 0 || numStrips > uint32_t(GatherMaxStrips)) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_frame_gather_create: bad argument");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	if (f->gatherImageOwned) return fail(f->ctx, RPT_ERR_INVALID, "rpt_frame_gather_create: this frame already is a gather root");
	const size_t bytes = size_t(f->width) * f->height * 4;
	CU(f->ctx, cudaMalloc(&f->gatherImageOwned, bytes));
	CU(f->ctx, cudaMemset(f->gatherImageOwned, 0, bytes));
	CU(f->ctx, cudaMalloc(&f->gatherFlagsOwned, 64 * sizeof(uint32_t)));
	CU(f->ctx, cudaMemset(f->gatherFlagsOwned, 0, 64 * sizeof(uint32_t)));
	std::memset(out, 0, sizeof(*out));
	CU(f->ctx, cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(out->imageHandle), f->gatherImageOwned));
	CU(f->ctx, cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(out->flagsHandle), f->gatherFlagsOwned));
	out->imagePtr = reinterpret_cast<uint64_t>(f->gatherImageOwned);
	out->flagsPtr = reinterpret_cast<uint64_t>(f->gatherFlagsOwned);
	out->pid = uint64_t(getpid());
	out->device = f->ctx->device;
	out->width = f->width; out->height = f->height; out->numStrips = numStrips;
	return RPT_OK;
}

RPT_API int rpt_frame_gather_connect(RptFrame* f, const RptGatherInfo* root, uint32_t stripIndex) {
	if (!f || !root) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_frame_gather_connect: NULL argument");
	if (root->width != f->width || root->height != f->height || stripIndex >= root->numStrips) return fail(f->ctx, RPT_ERR_INVALID, "rpt_frame_gather_connect: film size or strip index does not match the root");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	disconnectGather(f);
	RptFrame::Gather g;
	if (root->pid == uint64_t(getpid())) {
		if (root->device != f->ctx->device) {
			int can = 0;
			CU(f->ctx, cudaDeviceCanAccessPeer(&can, f->ctx->device, root->device));
			if (!can) return fail(f->ctx, RPT_ERR_UNSUPPORTED, "rpt_frame_gather_connect: no peer access to the root's device");
			cudaError_t e = cudaDeviceEnablePeerAccess(root->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cudaFail(f->ctx, e, "cudaDeviceEnablePeerAccess");
			cudaGetLastError();
		}
		g.image = reinterpret_cast<uchar4*>(root->imagePtr);
		g.flags = reinterpret_cast<uint32_t*>(root->flagsPtr);
	}
	else {
		void *a = nullptr, *b = nullptr;
		CU(f->ctx, cudaIpcOpenMemHandle(&a, *reinterpret_cast<const cudaIpcMemHandle_t*>(root->imageHandle), cudaIpcMemLazyEnablePeerAccess));
		cudaError_t e = cudaIpcOpenMemHandle(&b, *reinterpret_cast<const cudaIpcMemHandle_t*>(root->flagsHandle), cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) { cudaIpcCloseMemHandle(a); return cudaFail(f->ctx, e, "cudaIpcOpenMemHandle"); }
		g.image = static_cast<uchar4*>(a); g.flags = static_cast<uint32_t*>(b);
		g.ipc = true;
	}
	g.connected = true;
	g.root = f->gatherImageOwned != nullptr && reinterpret_cast<uint64_t>(f->gatherImageOwned) == root->imagePtr && root->pid == uint64_t(getpid());
	g.strip = stripIndex; g.numStrips = root->numStrips;
	f->gather = g;
	return RPT_OK;
}

RPT_API int rpt_frame_gather_disconnect(RptFrame* f) {
	if (!f) return fail(nullptr, RPT_ERR_INVALID, "rpt_frame_gather_disconnect: NULL frame");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	joinTail(f);
	joinLate(f);
	CU(f->ctx, syncFrame(f));
	disconnectGather(f);
	return RPT_OK;
}

// The root's side of the gather: wait (on the device) until every strip's rows of gather epoch k have arrived, copy the film to
// the host, then release the image: the strips' next post-process passes wait for that release before they overwrite it.
RPT_API int rpt_gather_output(RptFrame* f, uint8_t* rgba8FullFilm) {
	if (!f || !rgba8FullFilm) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_gather_output: NULL argument");
	if (!f->gatherImageOwned || !f->gather.connected) return fail(f->ctx, RPT_ERR_INVALID, "rpt_gather_output: this frame is not a connected gather root");
	if (f->hostError && *reinterpret_cast<volatile uint32_t*>(f->hostError)) return fail(f->ctx, RPT_ERR_PEER, "rpt_gather_output: a multi-GPU hand-over timed out on the device");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	joinTail(f);
	joinLate(f);
	const uint32_t epoch = ++f->gather.gathered;
	launchPeerWaitMany(f->gatherFlagsOwned + GatherArrivalFlag0, f->gather.numStrips, epoch, f->flags + PeerError, f->hostErrorDev, f->stream);
	CU(f->ctx, cudaMemcpyAsync(rgba8FullFilm, f->gatherImageOwned, size_t(f->width) * f->height * 4, cudaMemcpyDeviceToHost, f->stream));
	launchPeerSignal(f->gatherFlagsOwned + GatherReleaseFlag, nullptr, epoch, f->stream);
	CU(f->ctx, cudaGetLastError());
	CU(f->ctx, syncFrame(f));
	if (*reinterpret_cast<volatile uint32_t*>(f->hostError)) return fail(f->ctx, RPT_ERR_PEER, "rpt_gather_output: a strip's rows did not arrive within the time limit");
	return RPT_OK;
}

RPT_API int rpt_frame_peer_error(RptFrame* f) {
	if (!f) return 1;
	cudaSetDevice(f->ctx->device);
	uint32_t e = 0;
	if (syncFrame(f) != cudaSuccess) return 1;
	if (cudaMemcpy(&e, f->flags + PeerError, sizeof(e), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
	return int(e);
}

// ---- raw ray queries ------------------------------------------------------------------------------------------
static int traceCommon(RptCtx* ctx, const RptScene* s, const float* rays, uint32_t n, RptIntersection* out, uint8_t* occ) {
	if (!ctx || !s || (!rays && n) || (!out && !occ)) return fail(ctx, RPT_ERR_INVALID, "rpt_trace: NULL argument");
	if (n == 0) return RPT_OK;
	CU(ctx, cudaSetDevice(ctx->device));
	float4* dRays = nullptr; RptIntersection* dOut = nullptr; uint8_t* dOcc = nullptr;
	CU(ctx, cudaMalloc(&dRays, size_t(n) * 32));
	if (out) CU(ctx, cudaMalloc(&dOut, size_t(n) * sizeof(RptIntersection)));
	if (occ) CU(ctx, cudaMalloc(&dOcc, n));
	CU(ctx, cudaMemcpyAsync(dRays, rays, size_t(n) * 32, cudaMemcpyHostToDevice, ctx->stream));
	launchTraceRays(sceneView(s), dRays, n, dOut, dOcc, ctx->stream);
	CU(ctx, cudaGetLastError());
	if (out) CU(ctx, cudaMemcpyAsync(out, dOut, size_t(n) * sizeof(RptIntersection), cudaMemcpyDeviceToHost, ctx->stream));
	if (occ) CU(ctx, cudaMemcpyAsync(occ, dOcc, n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(dRays); if (dOut) cudaFree(dOut); if (dOcc) cudaFree(dOcc);
	return RPT_OK;
}
// Traversal microbenchmark: the same rays through either traversal kernel, timed with CUDA events on the context's
// stream.  kernel 0 = one ray per thread run to completion (what the per-pixel passes do in line), 1 = the persistent
// queue kernel with dynamic fetch (what the wavefront passes use).  Results of the last iteration are returned.
RPT_API int rpt_trace_bench(RptCtx* ctx, const RptScene* s, const float* rays, uint32_t n, int anyHit, int kernel, int iterations,
                            float* msPerIteration, RptIntersection* out, uint8_t* occ) {
	if (!ctx || !s || !rays || n == 0 || iterations < 1 || !msPerIteration) return fail(ctx, RPT_ERR_INVALID, "rpt_trace_bench: bad argument");
	CU(ctx, cudaSetDevice(ctx->device));
	float4* dRays = nullptr; RptIntersection* dOut = nullptr; uint8_t* dOcc = nullptr; uint32_t* dHead = nullptr;
	CU(ctx, cudaMalloc(&dRays, size_t(n) * 32));
	CU(ctx, cudaMalloc(&dOut, size_t(n) * sizeof(RptIntersection)));
	CU(ctx, cudaMalloc(&dOcc, n));
	CU(ctx, cudaMalloc(&dHead, sizeof(uint32_t)));
	CU(ctx, cudaMemcpyAsync(dRays, rays, size_t(n) * 32, cudaMemcpyHostToDevice, ctx->stream));
	cudaEvent_t e0, e1;
	CU(ctx, cudaEventCreate(&e0)); CU(ctx, cudaEventCreate(&e1));
	const SceneView sv = sceneView(s);
	for (int it = -2; it < iterations; it++) {   // two warm-up iterations
		if (it == 0) CU(ctx, cudaEventRecord(e0, ctx->stream));
		if (kernel == 0) launchTraceRays(sv, dRays, n, anyHit ? nullptr : dOut, anyHit ? dOcc : nullptr, ctx->stream);
		else {
			CU(ctx, cudaMemsetAsync(dHead, 0, sizeof(uint32_t), ctx->stream));
			if (anyHit) launchTraceQueueAny(sv, dRays, nullptr, n, dHead, dOcc, ctx->stream);
			else launchTraceQueueClosest(sv, dRays, nullptr, n, dHead, dOut, ctx->stream);
		}
	}
	CU(ctx, cudaEventRecord(e1, ctx->stream));
	CU(ctx, cudaGetLastError());
	if (out && !anyHit) CU(ctx, cudaMemcpyAsync(out, dOut, size_t(n) * sizeof(RptIntersection), cudaMemcpyDeviceToHost, ctx->stream));
	if (occ && anyHit) CU(ctx, cudaMemcpyAsync(occ, dOcc, n, cudaMemcpyDeviceToHost, ctx->stream));
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	float ms = 0.f;
	CU(ctx, cudaEventElapsedTime(&ms, e0, e1));
	*msPerIteration = ms / float(iterations);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaFree(dRays); cudaFree(dOut); cudaFree(dOcc); cudaFree(dHead);
	return RPT_OK;
}

RPT_API int rpt_trace_closest(RptCtx* ctx, const RptScene* s, const float* rays, uint32_t n, RptIntersection* out) { return traceCommon(ctx, s, rays, n, out, nullptr); }
RPT_API int rpt_trace_shadow(RptCtx* ctx, const RptScene* s, const float* rays, uint32_t n, uint8_t* occ) { return traceCommon(ctx, s, rays, n, nullptr, occ); }

// queue sizes of the last wavefront path-tracing pass: out[4*b + {0,1}] = extension / shadow rays of bounce b
RPT_API int rpt_wavefront_counters(RptFrame* f, uint32_t* out64) {
	if (!f || !out64) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_wavefront_counters: NULL argument");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	joinTail(f);
	joinLate(f);
	CU(f->ctx, syncFrame(f));
	CU(f->ctx, cudaMemcpy(out64, f->wfSet[f->lastWfSet].counters, size_t(WavefrontMaxBounces) * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	return RPT_OK;
}

RPT_API int rpt_reuse_counters(RptFrame* f, uint32_t* out16) {
	if (!f || !out16) return fail(f ? f->ctx : nullptr, RPT_ERR_INVALID, "rpt_reuse_counters: NULL argument");
	CU(f->ctx, cudaSetDevice(f->ctx->device));
	CU(f->ctx, syncFrame(f));
	std::memcpy(out16, f->hostReuseCounters, 16 * sizeof(uint32_t));
	return RPT_OK;
}

RPT_API int rpt_membench(RptCtx* ctx, size_t bytes, int iterations, float* streamGBs, float* chaseNs) {
	if (!ctx || iterations < 1 || (!streamGBs && !chaseNs)) return fail(ctx, RPT_ERR_INVALID, "rpt_membench: bad argument");
	CU(ctx, cudaSetDevice(ctx->device));
	CU(ctx, runMemBench(bytes, iterations, streamGBs, chaseNs, ctx->stream));
	return RPT_OK;
}

// ---- counters ---------------------------------------------------------------------------------------------------
RPT_API int rpt_counters_enable(RptCtx* ctx, int on) {
	if (!ctx) return fail(nullptr, RPT_ERR_INVALID, "rpt_counters_enable: NULL ctx");
	ctx->countersOn = on != 0;
	return RPT_OK;
}
RPT_API int rpt_counters_reset(RptCtx* ctx) {
	if (!ctx) return fail(nullptr, RPT_ERR_INVALID, "rpt_counters_reset: NULL ctx");
	CU(ctx, cudaSetDevice(ctx->device));
	CU(ctx, cudaDeviceSynchronize());
	CU(ctx, cudaMemset(ctx->counters, 0, 8 * sizeof(unsigned long long)));
	return RPT_OK;
}
RPT_API int rpt_counters_read(RptCtx* ctx, RptCounters* out) {
	if (!ctx || !out) return fail(ctx, RPT_ERR_INVALID, "rpt_counters_read: NULL argument");
	CU(ctx, cudaSetDevice(ctx->device));
	CU(ctx, cudaDeviceSynchronize());
	unsigned long long h[8];
	CU(ctx, cudaMemcpy(h, ctx->counters, sizeof(h), cudaMemcpyDeviceToHost));
	out->closestRays = h[0]; out->shadowRays = h[1]; out->nodeVisits = h[2]; out->triTests = h[3]; out->shadedHits = h[4];
	out->shadowNodeVisits = h[5]; out->shadowTriTests = h[6]; out->maxNodeVisits = h[7];
	return RPT_OK;
}
