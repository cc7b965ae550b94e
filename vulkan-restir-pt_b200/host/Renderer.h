// Headless frame orchestration.  Mirrors the reference Renderer's per-frame sequence
// (src/Renderer.cpp:400-503 recordRenderCommand, :358-368 memorySyncHostAndDevice, :539-569 drawFrame,
// :571-577 initSettings) minus window / swapchain / GUI, and the composite passes GRISReSTIR::render
// (src/GRISReSTIR.cpp:9-53) and TestReSTIR::render (src/TestReSTIR.cpp:9-36).  All device work goes through
// the C ABI of include/restirpt.h; there is no other path.
#pragma once
#include <string>
#include <vector>
#include "Scene.h"

namespace rpt {

struct RayTracingMethod {   // reference src/Renderer.h:22-26
	enum { None = 0, Naive = 1, ResampledDI = 2, ResampledGI = 2, ResampledPT = 3, VisualizeAS = 3 };
};

struct RendererSettings {   // reference src/Renderer.h:29-36, values after initSettings()
	int directMethod = RayTracingMethod::None;
	int indirectMethod = RayTracingMethod::ResampledPT;
	int toneMapping = 1;
	bool correctGamma = true;
	bool accumulate = false;
	// RayTracing::Mode (reference src/RayTracing.h:28-30,48): RayQuery = the .comp entry points (default), RayTracing = the
	// .rgen ones.  Only the naive direct pass runs a different estimator in the two modes (di_naive.rgen vs di_naive.comp).
	int pipelineMode = 0;   // 0 RayQuery, 1 RayTracing
};

// called between the temporal and the spatial pass when the frame is one strip of a multi-GPU film:
// must make the neighbours' boundary rows of the given buffer visible in this frame's halo rows
typedef void (*HaloExchangeFn)(void* user, RptFrame* frame, RptBufferId buffer);

class Renderer {
public:
	Renderer(const Scene& scene, uint32_t width, uint32_t height, int cudaDevice,
	         uint32_t rowBegin = 0, uint32_t rowEnd = 0, uint32_t halo = 0);
	~Renderer();
	Renderer(const Renderer&) = delete;
	Renderer& operator=(const Renderer&) = delete;

	// one frame; rgba8Out may be null (no read-back).  seed = this frame's Camera::seed
	void drawFrame(uint32_t seed, uint8_t* rgba8Out);
	// the same frame with a pipelined read-back: returns once the frame is enqueued; the image is complete in rgba8Out (pinned
	// host memory) after waitReadback(ticket).  Alternate between two host buffers: two read-backs may be in flight.
	uint64_t drawFrameAsync(uint32_t seed, uint8_t* rgba8Out);
	void waitReadback(uint64_t ticket);
	// all strips of one film that live in this process (multi-GPU from one host thread, or several strips on one device): the
	// same frame on every strip, stage by stage, so that every device-side hand-over wait finds its signal already enqueued
	static void drawStrips(Renderer* const* strips, uint32_t count, uint32_t seed, uint8_t* const* rgba8Outs);
	// dynamic scenes: push the scene's current object instances to the device and rebuild the acceleration structure
	void updateInstances(const Scene& scene);
	void clearReservoirs() { mClearNext = true; }   // GUI "clear" → Camera::setClearFlag
	void setHaloExchange(HaloExchangeFn fn, void* user) { mHaloFn = fn; mHaloUser = user; }

	Camera& camera() { return mCamera; }
	RptFrame* frame() { return mFrame; }
	RptScene* deviceScene() { return mDeviceScene; }
	RptCtx* ctx() { return mCtx; }
	uint32_t frameCount() const { return mFrameCount; }

	RendererSettings settings;
	RptGRISSettings grisSettings = { 2 /*Hybrid*/, 1.f, 0, 1, 20 };   // src/GRISReSTIR.h:29
	RptDISettings diSettings = { 0 /*Reconnection*/, 0 /*Light*/, 0, 1 };   // src/TestReSTIR.h:29

private:
	void check(int status, const char* what);
	void drawStage(int stage, uint32_t seed, uint8_t* rgba8Out, uint64_t* asyncTicket = nullptr);

	RptCtx* mCtx = nullptr;
	RptScene* mDeviceScene = nullptr;
	RptFrame* mFrame = nullptr;
	Camera mCamera, mPrevCamera;
	bool mMotionPending = false;
	bool mClearNext = false;
	uint32_t mFrameCount = 0;
	HaloExchangeFn mHaloFn = nullptr;
	void* mHaloUser = nullptr;
};

} // namespace rpt
