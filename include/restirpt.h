/*
 * restirpt.h — C ABI of the B200 (sm_100a) ReSTIR PT device library (librestirpt.so).
 *
 * This is the drop-in boundary of SURVEY.md §8(b): the reference's host (Renderer / Scene / Camera) reaches
 * its device code through "create pipeline / execute(extent, bindings, push-constant)" pass objects
 * (reference src/RayTracing.h:27-48).  Every entry point below names the reference interface it replaces.
 * Plain pointers and sizes only; no C++ / torch types.  All structs are byte-identical to the reference's
 * std430 layouts (reference src/shader/layouts.glsl) so host code written against the reference keeps working.
 *
 * Conventions: every function returns 0 on success or a negative RptStatus; rpt_last_error() gives the
 * message.  Handles are opaque, not thread-safe, and bound to one CUDA device.  All passes enqueue on the
 * frame's CUDA stream in call order (stream order replaces the reference's pipeline barriers,
 * reference src/GRISReSTIR.cpp:33-47); rpt_sync() blocks.  There is NO CPU fallback: without a CUDA device
 * rpt_ctx_create fails with RPT_ERR_NO_DEVICE.
 */
#ifndef RESTIRPT_H
#define RESTIRPT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum RptStatus {
	RPT_OK = 0,
	RPT_ERR_INVALID = -1,    /* bad argument / bad handle */
	RPT_ERR_NO_DEVICE = -2,  /* no CUDA device: the library never falls back to the CPU */
	RPT_ERR_CUDA = -3,       /* sticky CUDA error, see rpt_last_error */
	RPT_ERR_OOM = -4,
	RPT_ERR_UNSUPPORTED = -5,
	RPT_ERR_PEER = -6        /* a multi-GPU hand-over timed out on the device (sticky until the peers are disconnected) */
} RptStatus;

/* ---- data layouts (reference src/shader/layouts.glsl, host mirrors in the src headers) ------------------------- */

/* layouts.glsl:6-25 with RESTIR_PT_MATERIAL=1, host src/Material.h:14-38.  32 B */
typedef struct RptMaterial {
	float baseColor[3];
	uint32_t type;        /* 1 Lambert, 2 MetallicWorkflow, 3 Metal, 4 Dielectric, 6 Fake (material.glsl:13-20) */
	uint32_t textureIdx;  /* 0xffffffff = none */
	float metallic;
	float roughness;
	float ior;
} RptMaterial;

/* layouts.glsl:53-58, src/Model.h:15-33.  32 B */
typedef struct RptMeshVertex {
	float pos[3];
	float uvx;
	float norm[3];
	float uvy;
} RptMeshVertex;

/* layouts.glsl:60-72, src/Scene.h:15-27.  224 B.  Matrices are column-major (glm). */
typedef struct RptObjectInstance {
	float transform[16];
	float transformInv[16];
	float transformInvT[16];
	float radiance[3];
	float pad0;
	uint32_t indexOffset;
	uint32_t indexCount;
	uint32_t matIndex;   /* unused by the shaders; host pad */
	float pad2;
} RptObjectInstance;

/* layouts.glsl:74-83, src/Scene.h:29-38.  64 B, world space */
typedef struct RptTriangleLight {
	float v0[3]; float nx;
	float v1[3]; float ny;
	float v2[3]; float nz;
	float radiance[3]; float area;
} RptTriangleLight;

/* layouts.glsl:85-88, src/util/AliasTable.h:7-10.  8 B; N+1 entries, [0] = {sum power, N}; failId is 1-based */
typedef struct RptLightSampleTableElement {
	float prob;
	uint32_t failId;
} RptLightSampleTableElement;

/* layouts.glsl:27-51; the reference memcpy's its Camera object as the UBO (src/Camera.h:47-69,
 * src/Renderer.cpp:359-360).  352 B */
typedef struct RptCamera {
	float view[16];
	float proj[16];
	float projView[16];
	float lastProjView[16];
	float pos[3];   float FOV;
	float angle[3]; float nearZ;
	float front[3]; float farZ;
	float right[3]; float lensRadius;
	float up[3];    float focalDist;
	uint32_t filmSize[2];
	uint32_t frameIndex;   /* bit 31 = clear flag (camera.glsl:6-7) */
	uint32_t seed;
} RptCamera;

/* layouts.glsl:90-94.  16 B.  instanceIdx 0 = light, i+1 = object instance i,
 * 0xffffffff invalid, 0xfffffffe "special" = primary hit (ray_layouts.glsl:12-13) */
typedef struct RptIntersection {
	float bary[2];
	uint32_t instanceIdx;
	uint32_t triangleIdx;
} RptIntersection;

/* layouts.glsl:96-115.  64 B */
typedef struct RptDIReservoir {
	RptIntersection isec;
	float Li[3]; float pad0;
	float jacobian; float samplePdf; uint32_t rng; uint32_t isLightSample;
	uint32_t sampleCount; float resampleWeight; float contribWeight; float weight;
} RptDIReservoir;

/* layouts.glsl:117-130.  48 B */
typedef struct RptGIReservoir {
	RptIntersection rcIsec;
	float rcLo[3]; uint32_t rcPrevCoord;
	uint32_t sampleCount; float resampleWeight; float contribWeight; float pad0;
} RptGIReservoir;

/* layouts.glsl:132-156.  96 B (GLSL stride; the reference host over-allocates 112, src/Renderer.cpp:35) */
typedef struct RptGRISReservoir {
	RptIntersection rcIsec;
	float rcLi[3]; uint32_t rcRng;
	float rcWi[3]; uint32_t flags;   /* bits 0-7 rcVertexId, 8-15 pathLength, 16-23 rcVertexType */
	float pad[2]; float rcPrevSamplePdf; float rcJacobian;
	float F[3]; uint32_t primaryRng;
	float sampleCount; float resampleWeight; float contribWeight; float pad0;
} RptGRISReservoir;

/* push constants */
typedef struct RptDISettings {   /* src/TestReSTIR.h:12-17, default {Reconnection, Light, 0, 1} */
	uint32_t shiftType;          /* 0 Reconnection, 1 Replay, 2 Hybrid (no-op for DI) */
	uint32_t sampleType;         /* 0 Light, 1 BSDF, 2 Both */
	uint32_t temporalReuse;
	uint32_t spatialReuse;
} RptDISettings;

typedef struct RptGRISSettings { /* src/GRISReSTIR.h:11-17, default {Hybrid, 1, 0, 1, 20} */
	uint32_t shiftType;
	float rrScale;
	uint32_t temporalReuse;
	uint32_t spatialReuse;
	uint32_t cap;
} RptGRISSettings;

typedef struct RptPostSettings { /* src/PostProcessFrag.h:7-12 */
	uint32_t toneMapping;        /* 0 none, 1 filmic, 2 ACES */
	uint32_t correctGamma;
	uint32_t noDirect;
	uint32_t noIndirect;
} RptPostSettings;

/* One RGBA8 texture; texels are sRGB-encoded (reference zvk/core/HostImage.cpp:22), sampled bilinear or
 * nearest with REPEAT addressing (zvk/core/Memory.cpp:75-92) */
typedef struct RptTextureDesc {
	const uint8_t* rgba8;
	uint32_t width, height;
	uint32_t filter;             /* 0 linear, 1 nearest */
} RptTextureDesc;

/* What DeviceScene uploads (reference src/Scene.cpp:368-446) and what its acceleration-structure build
 * consumes (src/Scene.cpp:448-547).  Host memory is borrowed for the duration of the call only. */
typedef struct RptSceneDesc {
	const RptMeshVertex* vertices;           uint32_t numVertices;
	const uint32_t* indices;                 uint32_t numIndices;        /* absolute into vertices[] */
	const RptMaterial* materials;            uint32_t numMaterials;
	const int32_t* materialIndices;          uint32_t numMaterialIndices;/* one per object triangle */
	const RptObjectInstance* instances;      uint32_t numInstances;
	const RptTriangleLight* triangleLights;  uint32_t numTriangleLights;
	const RptLightSampleTableElement* lightSampleTable;                  /* numTriangleLights + 1 entries */
	const RptTextureDesc* textures;          uint32_t numTextures;
	uint32_t flags;                          /* RptSceneFlags */
} RptSceneDesc;

/* Acceleration-structure arrangement.  Default (0): every instance flattened to world space under ONE wide BVH (the reference's
 * loader never shares geometry between instances, src/Resource.cpp:183-184).  RPT_SCENE_TWO_LEVEL: the reference's own
 * arrangement (src/Scene.cpp:448-547): one BLAS per unique (indexOffset, indexCount) range in object space, a BLAS of the light
 * triangles (custom index 0) and a TLAS over the instances; rays are transformed at the instance boundary.  Memory is per
 * unique mesh, and rpt_scene_update_instances only rebuilds the TLAS. */
typedef enum RptSceneFlags { RPT_SCENE_TWO_LEVEL = 1 } RptSceneFlags;

/* "current frame" / THIS = what this frame's passes have written since the last rpt_frame_flip; PREV = what the previous frame
 * left.  THIS is only defined once the frame's own pass has written it (the reference's ping-pong pair would show the frame
 * before last there; the G-buffer, motion and GRIS buffers of this library rotate through three slots).  The one place the
 * pipeline itself depends on such old content — background pixels of gris_path_trace keep their reservoir — is reproduced. */
typedef enum RptBufferId {
	RPT_BUF_DIRECT_OUTPUT = 0,    /* float4 / px   (layouts.glsl:180)      */
	RPT_BUF_INDIRECT_OUTPUT = 1,  /* float4 / px   (layouts.glsl:181)      */
	RPT_BUF_DEPTH_NORMAL = 2,     /* float4 / px, current frame (binding 2)*/
	RPT_BUF_DEPTH_NORMAL_PREV = 3,
	RPT_BUF_ALBEDO_MATID = 4,     /* uint2 / px                            */
	RPT_BUF_ALBEDO_MATID_PREV = 5,
	RPT_BUF_MOTION = 6,           /* float2 / px, values rounded through fp16 (RG16F target) */
	RPT_BUF_DI_THIS = 7, RPT_BUF_DI_PREV = 8, RPT_BUF_DI_TEMP = 9,          /* 64 B / px */
	RPT_BUF_GI_THIS = 10, RPT_BUF_GI_PREV = 11,                             /* 48 B / px */
	RPT_BUF_GRIS_THIS = 12, RPT_BUF_GRIS_PREV = 13, RPT_BUF_GRIS_TEMP = 14, /* 96 B / px */
	RPT_BUF_PRIMARY_ISEC = 15,    /* RptIntersection / px of the G-buffer primary ray (parity aid, new) */
	RPT_BUF_COUNT = 16
} RptBufferId;

/* ray / traversal counters accumulated since rpt_counters_reset (new; needed for Mrays/s and
 * bytes-per-ray).  Only maintained when counting is enabled. */
typedef struct RptCounters {
	uint64_t closestRays;
	uint64_t shadowRays;
	uint64_t nodeVisits;     /* CWBVH nodes fetched (80 B each)  */
	uint64_t triTests;       /* triangles fetched (48 B each)    */
	uint64_t shadedHits;     /* loadSurfaceInfo gathers (272 B)  */
	uint64_t shadowNodeVisits;   /* the share of nodeVisits / triTests spent on occlusion rays */
	uint64_t shadowTriTests;
	uint64_t maxNodeVisits;  /* most nodes any single queued ray fetched (the length of a traversal kernel's tail) */
} RptCounters;

typedef struct RptBvhStats {
	uint32_t numTriangles;
	uint32_t numNodes;        /* 80-byte CWBVH nodes */
	uint64_t nodeBytes;
	uint64_t triBytes;
	float buildMs;            /* GPU time of the build (CUDA events) */
	float sahCost;
	/* two-level scenes (0 otherwise): numNodes / numTriangles / *Bytes above then count the BLASes, once per unique mesh */
	uint32_t twoLevel;
	uint32_t numMeshes;       /* BLASes, including the one of the light triangles */
	uint32_t numTlasNodes;
	uint32_t numInstanceRecords;
	float tlasBuildMs;        /* instance records + TLAS: what rpt_scene_update_instances costs on a two-level scene */
	uint32_t pad;
} RptBvhStats;

/* per-pass device timing (new): CUDA events recorded on the frame's stream around every pass launch */
typedef enum RptPassId {
	RPT_PASS_GBUFFER = 0, RPT_PASS_DI_NAIVE = 1, RPT_PASS_GI_NAIVE = 2, RPT_PASS_DI_PATHGEN = 3,
	RPT_PASS_DI_TEMPORAL = 4, RPT_PASS_DI_SPATIAL = 5, RPT_PASS_GI_RESTIR = 6, RPT_PASS_GRIS_PATHTRACE = 7,
	RPT_PASS_GRIS_TEMPORAL = 8, RPT_PASS_GRIS_SPATIAL = 9, RPT_PASS_VISUALIZE_AS = 10, RPT_PASS_POSTPROCESS = 11,
	RPT_PASS_COUNT = 12
} RptPassId;

/* kernels inside the multi-kernel wavefront passes, timed individually (events between the launches on the frame's
 * stream; the path-tracing tail that runs on the second stream is not included) */
typedef enum RptKernelId {
	RPT_KERNEL_TRACE_CLOSEST = 0, RPT_KERNEL_TRACE_ANY = 1, RPT_KERNEL_GRIS_BEGIN = 2, RPT_KERNEL_GRIS_BOUNCE = 3,
	RPT_KERNEL_GRIS_TAIL = 4,   /* the in-line tail of the path tracer, timed on its own (concurrent) stream */
	RPT_KERNEL_REUSE_GEN = 5,   /* temporal / spatial reuse: candidate + shift preparation kernels */
	RPT_KERNEL_REUSE_MERGE = 6, /* temporal / spatial reuse: merge, shading and list kernels */
	RPT_KERNEL_TAIL_WAIT = 7,   /* not a kernel: time the frame's stream waited for the path tracer's tail */
	RPT_KERNEL_TRACE_PAIR = 8,  /* two traversal launches side by side on two streams: the extension rays of bounce b (closest hit)
	                               and the shadow rays of vertex b-1 (any hit); one count per pair */
	RPT_KERNEL_COUNT = 9
} RptKernelId;

typedef struct RptPassStats {
	double ms[RPT_PASS_COUNT];         /* accumulated device time per pass since rpt_frame_timing(frame, 1) */
	uint64_t launches[RPT_PASS_COUNT]; /* kernel launches per pass */
	double kernelMs[RPT_KERNEL_COUNT];         /* the same per kernel of the wavefront passes */
	uint64_t kernelLaunches[RPT_KERNEL_COUNT];
} RptPassStats;

typedef struct RptCtx RptCtx;
typedef struct RptScene RptScene;
typedef struct RptFrame RptFrame;

/* ---- context (replaces zvk::Instance/Context creation, reference src/Renderer.cpp:82-109) -------------- */
int rpt_ctx_create(int cudaDevice, RptCtx** out);
void rpt_ctx_destroy(RptCtx* ctx);
const char* rpt_last_error(const RptCtx* ctx);  /* ctx may be NULL: last error of the calling thread */
int rpt_version(void);

/* ---- scene (replaces DeviceScene ctor, reference src/Scene.cpp:324-330: buffer upload + BLAS/TLAS build).
 * Builds the compressed wide BVH(s) on the GPU: flattened to world space, or BLAS + TLAS (desc->flags).  The arrays are checked
 * first (index / instance / material / texture / alias-table ranges): RPT_ERR_INVALID instead of an out-of-bounds device read. */
int rpt_scene_create(RptCtx* ctx, const RptSceneDesc* desc, RptScene** out);
void rpt_scene_destroy(RptScene* scene);
/* dynamic scenes (new; the reference is static): new transforms / radiance of the object instances, same
 * geometry ranges.  Flattened scenes rebuild the whole structure on the GPU; two-level scenes rebuild the instance records and
 * the TLAS only.  The new state is committed only if the rebuild succeeds.  Synchronises the device. */
int rpt_scene_update_instances(RptScene* scene, const RptObjectInstance* instances, uint32_t numInstances);
/* Per-instance motion vectors (new).  After an update the scene is "in motion": rpt_gbuffer follows every surface point back
 * through its instance's PREVIOUS placement before it applies lastProjView, so the motion image (and with it the temporal
 * reprojection of every ReSTIR pass) carries the objects' movement as well as the camera's — the reference's GBuffer.frag:40-44
 * reprojects the current position only, its scenes being static.  Call rpt_scene_end_motion once the frame that shows the
 * movement has been issued; until then "previous" stays the placement before the last update. */
int rpt_scene_end_motion(RptScene* scene);
int rpt_scene_bvh_stats(const RptScene* scene, RptBvhStats* out);

/* ---- frame resources (replaces Renderer::createRayImage + GBufferPass::createResource,
 * reference src/Renderer.cpp:193-257, src/GBufferPass.cpp:81-137).
 * A frame owns the rows [rowBegin, rowEnd) of a fullWidth x fullHeight film plus `halo` guard rows on each
 * interior edge (multi-GPU strips); single GPU: rowBegin=0,rowEnd=fullHeight,halo=0. */
int rpt_frame_create(RptCtx* ctx, uint32_t fullWidth, uint32_t fullHeight,
                     uint32_t rowBegin, uint32_t rowEnd, uint32_t halo, RptFrame** out);
void rpt_frame_destroy(RptFrame* frame);
int rpt_frame_clear(RptFrame* frame);            /* zero every buffer, reset ping-pong */
int rpt_frame_flip(RptFrame* frame);             /* mCurFrame ^= 1, reference src/Renderer.cpp:567 (THIS becomes PREV) */
void* rpt_frame_stream(RptFrame* frame);         /* cudaStream_t the passes run on (for event timing) */
/* Frames overlap (new; the reference keeps one frame in flight, HostDevice.h:7): the reuse passes of a ReSTIR PT frame —
 * rpt_gris_temporal, rpt_gris_spatial and the post-process that follows — run on a second stream set, in order, next to
 * rpt_gbuffer and rpt_gris_pathtrace of the NEXT frame, which touch none of their buffers (the G-buffer, motion and GRIS
 * reservoir buffers rotate through three slots, the path tracer's queues exist twice); the frame's stream is at most one
 * frame ahead.  Every other call joins them first, so results do not change and buffers read between passes hold what the
 * pass wrote.  rpt_frame_join makes the frame's stream wait (on the device, the host does not block) for everything enqueued
 * on the frame's other streams so far: an event recorded on rpt_frame_stream() after it covers all of the frame's work.
 * Environment RPT_NO_FRAME_OVERLAP=1 keeps one frame at a time. */
int rpt_frame_join(RptFrame* frame);

/* 2 x 352-byte Camera upload, reference src/Renderer.cpp:358-361 */
int rpt_set_camera(RptFrame* frame, const RptCamera* cur, const RptCamera* prev);

/* ---- passes.  Each replaces one RayTracing::execute / pass ::render call of the reference ------------- */
int rpt_gbuffer(RptFrame* f, const RptScene* s);                               /* GBufferPass::render, src/GBufferPass.cpp:22-56 */
int rpt_di_naive(RptFrame* f, const RptScene* s);                              /* mNaiveDIPass,  shader di_naive.comp */
int rpt_di_naive_rt(RptFrame* f, const RptScene* s);                           /* mNaiveDIPass in RayTracing-pipeline mode (src/RayTracing.h:28-30): shader di_naive.rgen,
                                                                                   the one entry point whose estimator differs from its .comp twin */
int rpt_gi_naive(RptFrame* f, const RptScene* s);                              /* mNaiveGIPass,  shader gi_naive.comp */
int rpt_di_pathgen(RptFrame* f, const RptScene* s, const RptDISettings* st);   /* TestReSTIR::render step 1, src/TestReSTIR.cpp:9-36 */
int rpt_di_temporal(RptFrame* f, const RptScene* s, const RptDISettings* st);  /*   step 2 */
int rpt_di_spatial(RptFrame* f, const RptScene* s, const RptDISettings* st);   /*   step 3 */
int rpt_gi_restir(RptFrame* f, const RptScene* s);                             /* mResampledGIPass, shader gi_resample_temporal.comp */
int rpt_gris_pathtrace(RptFrame* f, const RptScene* s, const RptGRISSettings* st); /* GRISReSTIR::render step 1, src/GRISReSTIR.cpp:9-53 */
int rpt_gris_temporal(RptFrame* f, const RptScene* s, const RptGRISSettings* st);  /*   step 2 */
int rpt_gris_spatial(RptFrame* f, const RptScene* s, const RptGRISSettings* st);   /*   step 3 */
int rpt_visualize_as(RptFrame* f, const RptScene* s);                          /* as_visualize.comp (triangle tests / primary ray / 100) */
/* PostProcessFrag::render, src/PostProcessFrag.cpp:156-184.  rgba8Out may be NULL (device-only run);
 * otherwise it receives ownedRows*fullWidth*4 bytes (R,G,B,A order) after an implicit sync. */
int rpt_postprocess(RptFrame* f, const RptPostSettings* st, uint8_t* rgba8Out);
/* The same pass with a pipelined read-back (new; the reference presents from the GPU and only reads back for a screenshot,
 * src/Renderer.cpp:758-793): returns at once with a ticket; the image travels to rgba8Out (pinned host memory, for the copy to
 * be asynchronous) on a copy stream while the next frame renders.  rpt_readback_wait(ticket) blocks until rgba8Out is complete.
 * Three read-backs may be in flight (the host should stay two frames ahead of the image it waits for: the reuse passes of a frame
 * run behind the next frame's path tracer); rotate through three rgba8Out buffers. */
int rpt_postprocess_async(RptFrame* f, const RptPostSettings* st, uint8_t* rgba8Out, uint64_t* ticket);
int rpt_readback_wait(RptFrame* f, uint64_t ticket);

int rpt_sync(RptFrame* f);

/* enable (1) / disable (0) per-pass event timing; enabling resets the accumulators */
int rpt_frame_timing(RptFrame* f, int enable);
int rpt_frame_pass_stats(RptFrame* f, RptPassStats* out);   /* implicit sync */

/* ---- read-back / upload of any frame buffer (new; parity tests and multi-GPU halo plumbing) ----------- */
/* Rows are frame-local storage rows: row 0 is film row max(rowBegin-halo,0). */
size_t rpt_buffer_stride(RptBufferId id);         /* bytes per pixel */
int rpt_frame_rows(const RptFrame* f, uint32_t* storageRowBegin, uint32_t* storageRowEnd);
int rpt_read(RptFrame* f, RptBufferId id, void* dst, size_t bytes);      /* whole storage, implicit sync */
int rpt_write(RptFrame* f, RptBufferId id, const void* src, size_t bytes);
void* rpt_device_ptr(RptFrame* f, RptBufferId id); /* raw device pointer of the storage (P2P / NCCL halo exchange) */

/* ---- multi-GPU strips (new) ----------------------------------------------------------------
 * One frame per GPU owns a horizontal strip plus `halo` guard rows.  After rpt_frame_connect_peers the temporal
 * passes store the reservoirs of their boundary rows straight into the neighbours' halo rows through NVLink peer
 * memory (CUDA IPC across processes, plain peer access inside one process), and the hand-over is ordered on the
 * device by epoch flags in peer memory — no host round trip, no collective.  Temporal reuse stays GPU-local. */
typedef struct RptPeerInfo {
	uint8_t grisTempHandle[64];   /* cudaIpcMemHandle_t of the GRIS temp reservoir buffer */
	uint8_t diTempHandle[64];     /* ... of the DI temp reservoir buffer */
	uint8_t flagsHandle[64];      /* ... of the epoch flags */
	uint8_t grisHandle[3][64];    /* ... of the buffers of the final GRIS (three, in rotation: this frame, the previous one, and the one the */
	uint8_t diHandle[2][64];      /*     next frame's path tracer already fills) / DI / GI (ping-pong pairs) reservoirs: the boundary rows of the */
	uint8_t giHandle[2][64];      /*     spatial (GI: temporal) pass output are mirrored into the neighbours' halo rows, so that previous-frame */
	                              /*     lookups that cross a cut stay GPU-local and still find their history */
	uint64_t grisTempPtr, diTempPtr, flagsPtr;   /* raw device pointers, used when pid matches */
	uint64_t grisPtr[3], diPtr[2], giPtr[2];
	uint64_t pid;
	int32_t device;
	uint32_t rowBegin, rowEnd, storeBegin, storeEnd;
	uint32_t cur;                 /* frames flipped at export time: the buffer phases (both strips must flip in lock step from here on) */
	uint32_t pad[2];
} RptPeerInfo;
int rpt_frame_export_peer(RptFrame* f, RptPeerInfo* out);
/* up = the strip above (smaller rows), down = the strip below; NULL at the film edge */
int rpt_frame_connect_peers(RptFrame* f, const RptPeerInfo* up, const RptPeerInfo* down);
/* drop the mappings of the neighbours' buffers again (every rank calls this, then a barrier, before any strip is
 * destroyed or re-partitioned); also done by rpt_frame_destroy */
int rpt_frame_disconnect_peers(RptFrame* f);
/* non-zero if a device-side hand-over wait timed out (a neighbour died or was not driven in lock step).  The condition is
 * sticky: from then on every pass of this frame fails with RPT_ERR_PEER until the peers are disconnected. */
int rpt_frame_peer_error(RptFrame* f);
/* 1 when a connected neighbour lives in this process (its passes are enqueued by the same host thread): such strips must be
 * driven stage by stage — every strip's temporal pass before any strip's spatial pass (rh_draw_strips does) */
int rpt_frame_peers_in_process(const RptFrame* f);

/* ---- final image gather (new; SURVEY.md §8(e) step 4, "rpt_gather_output" of §8(b)) -----------------------------
 * The root strip allocates one full-film RGBA8 image; every strip (the root included) connects to it, after which
 * rpt_postprocess stores its rows straight into that image through NVLink peer memory (fused post-process + gather: no
 * staging copy, no collective) and raises an arrival flag.  rpt_gather_output on the root waits on the device for all
 * strips of the current frame, copies the film to the host and releases the image for the next frame.  Once connected,
 * the root must gather every frame (the strips wait for the release before they overwrite the image). */
typedef struct RptGatherInfo {
	uint8_t imageHandle[64];      /* cudaIpcMemHandle_t of the film image on the root's GPU */
	uint8_t flagsHandle[64];      /* ... of the arrival / release flags */
	uint64_t imagePtr, flagsPtr;  /* raw device pointers, used when pid matches */
	uint64_t pid;
	int32_t device;
	uint32_t width, height, numStrips;
} RptGatherInfo;
int rpt_frame_gather_create(RptFrame* root, uint32_t numStrips, RptGatherInfo* out);
int rpt_frame_gather_connect(RptFrame* f, const RptGatherInfo* root, uint32_t stripIndex);
int rpt_frame_gather_disconnect(RptFrame* f);   /* every strip, then a barrier, before the root frame is destroyed */
int rpt_gather_output(RptFrame* root, uint8_t* rgba8FullFilm);   /* width*height*4 bytes on the host; implicit sync */

/* ---- ray queries exposed directly (new; closest-hit primitive-ID parity, traversal microbench) -------- */
/* rays: n x {ox,oy,oz,tmin, dx,dy,dz,tmax} floats on the HOST; out: n RptIntersection on the host */
int rpt_trace_closest(RptCtx* ctx, const RptScene* s, const float* rays, uint32_t n, RptIntersection* out);
int rpt_trace_shadow(RptCtx* ctx, const RptScene* s, const float* rays, uint32_t n, uint8_t* occludedOut);

/* traversal microbenchmark (new): `iterations` timed passes over the same n host rays after two warm-up passes.
 * anyHit 0 = closest hit (out receives n RptIntersection), 1 = occlusion (occ receives n bytes); out / occ may be NULL.
 * kernel 0 = one ray per thread run to completion, 1 = persistent queue kernel with dynamic fetch (wavefront passes). */
int rpt_trace_bench(RptCtx* ctx, const RptScene* s, const float* rays, uint32_t n, int anyHit, int kernel, int iterations,
                    float* msPerIteration, RptIntersection* out, uint8_t* occludedOut);

/* queue sizes of the last wavefront path-tracing pass (new; diagnostics): out64[4*b + 0] = extension rays traced
 * for bounce b, out64[4*b + 1] = shadow rays of bounce b; implicit sync */
int rpt_wavefront_counters(RptFrame* f, uint32_t* out64);
/* list sizes of the last spatial-reuse pass (new, diagnostics): out[0] = pixels whose final shading needed replay rays, out[1] =
 * pixels recomputed sequentially, out[3] = (pixel, neighbour) pairs replayed by the in-line list kernel, out[5] = ... by the
 * replay wavefront.  The library reads the same numbers back asynchronously to pick the replay form for the next frame: the
 * wavefront pays a fixed latency per bounce and wins when the list is long (RPT_RW_MIN_LIST, default 80000 pairs). */
int rpt_reuse_counters(RptFrame* f, uint32_t* out16);

/* memory-system microbenchmark (new): all SMs read a buffer of `bytes` bytes `iterations` times with 16-byte
 * loads -> GB/s (a buffer that fits B200's 126 MB L2 gives the L2 bandwidth, the roof of the traversal kernels on an L2-resident
 * scene; a multi-GB one the HBM read bandwidth), and one thread chases pointers through it -> ns per dependent load.
 * Either output may be NULL. */
int rpt_membench(RptCtx* ctx, size_t bytes, int iterations, float* streamGBs, float* chaseNs);

int rpt_counters_enable(RptCtx* ctx, int on);
int rpt_counters_reset(RptCtx* ctx);
int rpt_counters_read(RptCtx* ctx, RptCounters* out);

#ifdef __cplusplus
}
#endif
#endif /* RESTIRPT_H */
