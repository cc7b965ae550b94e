#include "Renderer.h"
#include <stdexcept>

namespace rpt {

void Renderer::check(int status, const char* what) {
	if (status != RPT_OK) {
		throw std::runtime_error(std::string(what) + ": " + rpt_last_error(mCtx));
	}
}

Renderer::Renderer(const Scene& scene, uint32_t width, uint32_t height, int cudaDevice,
                   uint32_t rowBegin, uint32_t rowEnd, uint32_t halo) {
	if (rowEnd == 0) rowEnd = height;
	check(rpt_ctx_create(cudaDevice, &mCtx), "rpt_ctx_create");
	RptSceneDesc desc = scene.desc();
	check(rpt_scene_create(mCtx, &desc, &mDeviceScene), "rpt_scene_create");
	check(rpt_frame_create(mCtx, width, height, rowBegin, rowEnd, halo, &mFrame), "rpt_frame_create");

	// reference src/Renderer.cpp:121-125: the window size overrides the XML film size; planes 0.001 / 200
	mCamera = scene.camera;
	mCamera.setFilmSize(width, height);
	mCamera.setPlanes(0.001f, 200.f);
	std::memcpy(mCamera.data().lastProjView, mCamera.data().projView, 64);
	mPrevCamera = mCamera;
}

Renderer::~Renderer() {
	if (mFrame) rpt_frame_destroy(mFrame);
	if (mDeviceScene) rpt_scene_destroy(mDeviceScene);
	if (mCtx) rpt_ctx_destroy(mCtx);
}

void Renderer::updateInstances(const Scene& scene) {
	check(rpt_scene_update_instances(mDeviceScene, scene.objectInstances.data(), uint32_t(scene.objectInstances.size())),
	      "rpt_scene_update_instances");
	mMotionPending = true;   // the next frame's G-buffer carries the objects' motion; it ends with that frame
}

void Renderer::drawFrame(uint32_t seed, uint8_t* rgba8Out) {
	// Strips of one film that live in this process are enqueued by one host thread: a strip's spatial pass waits (on the device)
	// for the neighbours' temporal passes, which must therefore be enqueued first — Renderer::drawStrips does that.
	if (rpt_frame_peers_in_process(mFrame)) {
		throw std::runtime_error("Renderer::drawFrame: this strip is connected to a neighbouring strip of the same process; "
		                         "draw all strips of the film together with Renderer::drawStrips (rh_draw_strips)");
	}
	drawStage(0, seed, nullptr);
	drawStage(1, seed, rgba8Out);
}

uint64_t Renderer::drawFrameAsync(uint32_t seed, uint8_t* rgba8Out) {
	if (rpt_frame_peers_in_process(mFrame)) throw std::runtime_error("Renderer::drawFrameAsync: strips of one process are drawn together with Renderer::drawStrips");
	if (!rgba8Out) throw std::runtime_error("Renderer::drawFrameAsync: no host buffer");
	uint64_t ticket = 0;
	drawStage(0, seed, nullptr);
	drawStage(1, seed, rgba8Out, &ticket);
	return ticket;
}

void Renderer::waitReadback(uint64_t ticket) { check(rpt_readback_wait(mFrame, ticket), "rpt_readback_wait"); }

void Renderer::drawStrips(Renderer* const* strips, uint32_t count, uint32_t seed, uint8_t* const* rgba8Outs) {
	for (uint32_t i = 0; i < count; i++) strips[i]->drawStage(0, seed, nullptr);
	for (uint32_t i = 0; i < count; i++) strips[i]->drawStage(1, seed, rgba8Outs ? rgba8Outs[i] : nullptr);
}

// stage 0: camera upload, G-buffer and everything up to and including the temporal passes (candidate generation, path tracing,
//          temporal reuse, ReSTIR GI); stage 1: the spatial passes (they read the neighbours' temporal output), post-process, flip.
// The direct and the indirect method write disjoint buffers, so running di_spatial after gris_temporal changes no result.
void Renderer::drawStage(int stage, uint32_t seed, uint8_t* rgba8Out, uint64_t* asyncTicket) {
	if (stage == 0) {
		// processGUI tail (src/Renderer.cpp:654-660): without accumulation the camera is re-updated every
		// frame, which zeroes frameIndex so every frame is shown un-accumulated
		if (!settings.accumulate) {
			mCamera.update();
		}
		if (mClearNext) {
			mCamera.setClearFlag();
			mClearNext = false;
		}
		// memorySyncHostAndDevice (src/Renderer.cpp:358-368)
		mCamera.data().seed = seed;
		check(rpt_set_camera(mFrame, &mCamera.data(), &mPrevCamera.data()), "rpt_set_camera");
		mPrevCamera = mCamera;
		mCamera.nextFrame(seed);

		// recordRenderCommand (src/Renderer.cpp:400-503)
		check(rpt_gbuffer(mFrame, mDeviceScene), "rpt_gbuffer");
	}

	if (settings.directMethod == RayTracingMethod::Naive) {
		if (stage == 0) {
			if (settings.pipelineMode == 1) check(rpt_di_naive_rt(mFrame, mDeviceScene), "rpt_di_naive_rt");
			else check(rpt_di_naive(mFrame, mDeviceScene), "rpt_di_naive");
		}
	}
	else if (settings.directMethod == RayTracingMethod::ResampledDI) {
		// TestReSTIR::render (src/TestReSTIR.cpp:9-36)
		if (stage == 0) {
			check(rpt_di_pathgen(mFrame, mDeviceScene, &diSettings), "rpt_di_pathgen");
			check(rpt_di_temporal(mFrame, mDeviceScene, &diSettings), "rpt_di_temporal");
		}
		else {
			if (mHaloFn) mHaloFn(mHaloUser, mFrame, RPT_BUF_DI_TEMP);
			check(rpt_di_spatial(mFrame, mDeviceScene, &diSettings), "rpt_di_spatial");
		}
	}
	else if (settings.directMethod == RayTracingMethod::VisualizeAS) {
		if (stage == 0) check(rpt_visualize_as(mFrame, mDeviceScene), "rpt_visualize_as");
	}

	if (settings.indirectMethod == RayTracingMethod::Naive) {
		if (stage == 0) check(rpt_gi_naive(mFrame, mDeviceScene), "rpt_gi_naive");
	}
	else if (settings.indirectMethod == RayTracingMethod::ResampledGI) {
		if (stage == 0) check(rpt_gi_restir(mFrame, mDeviceScene), "rpt_gi_restir");
	}
	else if (settings.indirectMethod == RayTracingMethod::ResampledPT) {
		// GRISReSTIR::render (src/GRISReSTIR.cpp:9-53)
		if (stage == 0) {
			check(rpt_gris_pathtrace(mFrame, mDeviceScene, &grisSettings), "rpt_gris_pathtrace");
			check(rpt_gris_temporal(mFrame, mDeviceScene, &grisSettings), "rpt_gris_temporal");
		}
		else {
			if (mHaloFn) mHaloFn(mHaloUser, mFrame, RPT_BUF_GRIS_TEMP);
			check(rpt_gris_spatial(mFrame, mDeviceScene, &grisSettings), "rpt_gris_spatial");
		}
	}
	if (stage == 0) return;

	RptPostSettings post;
	post.toneMapping = uint32_t(settings.toneMapping);
	post.correctGamma = settings.correctGamma ? 1u : 0u;
	post.noDirect = settings.directMethod == RayTracingMethod::None;
	post.noIndirect = settings.indirectMethod == RayTracingMethod::None;
	if (asyncTicket) check(rpt_postprocess_async(mFrame, &post, rgba8Out, asyncTicket), "rpt_postprocess_async");
	else check(rpt_postprocess(mFrame, &post, rgba8Out), "rpt_postprocess");

	if (mMotionPending) { check(rpt_scene_end_motion(mDeviceScene), "rpt_scene_end_motion"); mMotionPending = false; }
	check(rpt_frame_flip(mFrame), "rpt_frame_flip");   // mCurFrame ^= 1 (src/Renderer.cpp:567)
	mFrameCount++;
}

} // namespace rpt
