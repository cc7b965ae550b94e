// TEST INFRASTRUCTURE.  The reference's compute shaders — src/shader/*.comp with everything they #include — compiled by g++ from
// where they lie under /root/reference (oracle/ref/Makefile: glsl_to_cpp.py --shader runs the C preprocessor and the syntactic
// rewrite into a temporary directory; glsl_compat.h supplies the GLSL types and built-ins) and executed on the CPU, one invocation
// of main() per pixel, against plain buffers with the reference's layouts.  This is the reference's own algorithm running here:
// tests/test_cpu_ref_shaders.py compares every buffer it writes with what the oracle's restatement writes from the same inputs.
//
// What is NOT the reference's in this build, because the Vulkan driver supplies it there: ray / triangle intersection, texture and
// G-buffer filtering (call-backs, RefDriver — the test passes the oracle's definitions) and the built-in function library
// (glsl_compat.h, see its header for the two builds).  This file is compiled twice: libref.so carries `ref_run_shader` (IEEE
// built-ins + libm) and `refc_run_shader` (-DGLSL_BUILTINS_CONTRACT).
#define GLSL_COMPAT_RESOURCES
#include "glsl_compat.h"
#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>

#ifdef GLSL_BUILTINS_CONTRACT
#define REF_NAME(x) refc_##x
#else
#define REF_NAME(x) ref_##x
thread_local const RefDriver* glslDriver = nullptr;
thread_local uvec3 gl_GlobalInvocationID;
#endif

// plain pointers and sizes; the per-pixel buffers have the layouts of include/restirpt.h (= layouts.glsl)
struct RefBindings {
	const void* camera; const void* prevCamera;                       // 352 B each
	const void* materials; const int* materialIndices; const void* vertices; const uint* indices;
	const void* instances; const void* lights; const void* lightTable;
	uint numMaterials, numMaterialIndices, numVertices, numIndices, numInstances, numLights, numTextures;
	uint width, height;
	float* directOutput; float* indirectOutput;                       // rgba32f here (the reference's are rgba16f)
	const float* depthNormal; const float* depthNormalPrev;           // rgba32f
	const uint* albedoMatId; const uint* albedoMatIdPrev;             // rg32ui
	const float* motion;                                              // rg32f
	void* di; void* diPrev; void* diTemp; void* gi; void* giPrev; void* gris; void* grisPrev; void* grisTemp; void* grisRc;
	const void* push; uint pushBytes;                                 // the shader's push-constant block
	RefDriver driver;
};

namespace {

template <typename Fn>
void parallelRows(const RefBindings& b, int threads, Fn fn) {
	int nt = threads > 0 ? threads : int(std::thread::hardware_concurrency());
	if (nt < 1) nt = 1;
	std::atomic<uint> next{ 0 };
	auto worker = [&]() {
		glslDriver = &b.driver;
		for (;;) {
			const uint y = next.fetch_add(1);
			if (y >= b.height) break;
			fn(y);
		}
	};
	if (nt == 1) { worker(); return; }
	std::vector<std::thread> pool;
	for (int i = 0; i < nt; i++) pool.emplace_back(worker);
	for (auto& t : pool) t.join();
}

}  // namespace

// inside a shader's namespace, after its text: bind the interface variables of layouts.glsl / ray_layouts.glsl and run main() per pixel
#define REF_SHADER_GLUE(PUSH_STMT) \
	static_assert(sizeof(Camera) == 352 && sizeof(Material) == 32 && sizeof(MeshVertex) == 32 && sizeof(ObjectInstance) == 224 && \
	              sizeof(TriangleLight) == 64 && sizeof(LightSampleTableElement) == 8 && sizeof(Intersection) == 16 && \
	              sizeof(DIReservoir) == 64 && sizeof(GIReservoir) == 48 && sizeof(GRISReservoir) == 96 && \
	              sizeof(GRISReconnectionData) == 48, "std430 layouts"); \
	static std::vector<sampler2D> textures; \
	static int run(const RefBindings& b, int threads) { \
		std::memcpy(&uCamera, b.camera, sizeof(Camera)); std::memcpy(&uPrevCamera, b.prevCamera, sizeof(Camera)); \
		textures.assign(b.numTextures, sampler2D()); \
		for (uint i = 0; i < b.numTextures; i++) textures[i].index = i; \
		const size_t px = size_t(b.width) * b.height; \
		uTextures = { textures.data(), b.numTextures }; \
		uMaterials = { (Material*)b.materials, b.numMaterials }; uMaterialIndices = { (int*)b.materialIndices, b.numMaterialIndices }; \
		uVertices = { (MeshVertex*)b.vertices, b.numVertices }; uIndices = { (uint*)b.indices, b.numIndices }; \
		uObjectInstances = { (ObjectInstance*)b.instances, b.numInstances }; uTriangleLights = { (TriangleLight*)b.lights, b.numLights }; \
		uLightSampleTable = { (LightSampleTableElement*)b.lightTable, size_t(b.numLights) + 1 }; \
		uDirectOutput = image2D{ b.directOutput, b.width, b.height }; uIndirectOutput = image2D{ b.indirectOutput, b.width, b.height }; \
		uDepthNormal = sampler2D{ 1, 0, 0, b.depthNormal, 4, b.width, b.height }; \
		uDepthNormalPrev = sampler2D{ 1, 0, 1, b.depthNormalPrev, 4, b.width, b.height }; \
		uAlbedoMatId = usampler2D{ b.albedoMatId, 2, b.width, b.height }; \
		uAlbedoMatIdPrev = usampler2D{ b.albedoMatIdPrev, 2, b.width, b.height }; \
		uMotionVector = sampler2D{ 2, 0, 0, b.motion, 2, b.width, b.height }; \
		uDIReservoir = { (DIReservoir*)b.di, px }; uDIReservoirPrev = { (DIReservoir*)b.diPrev, px }; uDIReservoirTemp = { (DIReservoir*)b.diTemp, px }; \
		uGIReservoir = { (GIReservoir*)b.gi, px }; uGIReservoirPrev = { (GIReservoir*)b.giPrev, px }; \
		uGRISReservoir = { (GRISReservoir*)b.gris, px }; uGRISReservoirPrev = { (GRISReservoir*)b.grisPrev, px }; \
		uGRISReservoirTemp = { (GRISReservoir*)b.grisTemp, px }; uGRISReconnectionData = { (GRISReconnectionData*)b.grisRc, px }; \
		PUSH_STMT \
		parallelRows(b, threads, [&](uint y) { \
			for (uint x = 0; x < b.width; x++) { gl_GlobalInvocationID = uvec3(x, y, 0); main(); } \
		}); \
		return 0; \
	}
#define REF_PUSH { if (b.pushBytes != sizeof(uSettings)) return -2; std::memcpy(&uSettings, b.push, sizeof(uSettings)); }
#define REF_NO_PUSH { if (b.pushBytes != 0) return -2; }

namespace REF_NAME(sh_di_naive) {
#include "di_naive.inc"
REF_SHADER_GLUE(REF_NO_PUSH)
}
namespace REF_NAME(sh_gi_naive) {
#include "gi_naive.inc"
REF_SHADER_GLUE(REF_NO_PUSH)
}
namespace REF_NAME(sh_di_path_gen) {
#include "di_path_gen.inc"
REF_SHADER_GLUE(REF_PUSH)
}
namespace REF_NAME(sh_di_temporal) {
#include "di_temporal.inc"
REF_SHADER_GLUE(REF_PUSH)
}
namespace REF_NAME(sh_di_spatial) {
#include "di_spatial.inc"
REF_SHADER_GLUE(REF_PUSH)
}
namespace REF_NAME(sh_gi_resample_temporal) {
#include "gi_resample_temporal.inc"
REF_SHADER_GLUE(REF_NO_PUSH)
}
namespace REF_NAME(sh_gris_path_trace) {
#include "gris_path_trace.inc"
REF_SHADER_GLUE(REF_PUSH)
}
namespace REF_NAME(sh_gris_resample_temporal) {
#include "gris_resample_temporal.inc"
REF_SHADER_GLUE(REF_PUSH)
}
namespace REF_NAME(sh_gris_resample_spatial) {
#include "gris_resample_spatial.inc"
REF_SHADER_GLUE(REF_PUSH)
}
namespace REF_NAME(sh_as_visualize) {
#include "as_visualize.inc"
REF_SHADER_GLUE(REF_NO_PUSH)
}

// runs one compute shader over the whole film; 0 = ok, -1 = unknown shader, -2 = push-constant block of the wrong size
extern "C" __attribute__((visibility("default"))) int REF_NAME(run_shader)(const char* name, const RefBindings* b, int threads) {
#define REF_CASE(n) if (!std::strcmp(name, #n)) return REF_NAME(sh_##n)::run(*b, threads);
	REF_CASE(di_naive) REF_CASE(gi_naive) REF_CASE(di_path_gen) REF_CASE(di_temporal) REF_CASE(di_spatial)
	REF_CASE(gi_resample_temporal) REF_CASE(gris_path_trace) REF_CASE(gris_resample_temporal) REF_CASE(gris_resample_spatial)
	REF_CASE(as_visualize)
	return -1;
}
