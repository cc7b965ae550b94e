// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points of liboracle.so, shaped like include/restirpt.h (orc_*
// instead of rpt_*) so the parity tests drive the CUDA library and the oracle with the same calls.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may load this.
// Rows of a pass are split over std::thread workers (the reference's unit of parallelism is the pixel).
#include <algorithm>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>
#include "oracle_shading.h"

using namespace orc;

namespace orc {

void Frame2D::create(uint32_t w, uint32_t h) {
	width = w; height = h;
	size_t n = size_t(w) * h;
	directOutput.assign(n, vec4{ 0, 0, 0, 0 });
	indirectOutput.assign(n, vec4{ 0, 0, 0, 0 });
	for (int i = 0; i < 2; i++) {
		depthNormal[i].assign(n, vec4{ 0, 0, 0, 0 });
		albedoMatId[i].assign(n, uvec2{ 0, 0 });
		di[i].assign(n, RptDIReservoir{});
		gi[i].assign(n, RptGIReservoir{});
		gris[i].assign(n, RptGRISReservoir{});
	}
	motion.assign(n, vec2{ 0, 0 });
	diTemp.assign(n, RptDIReservoir{});
	grisTemp.assign(n, RptGRISReservoir{});
	primaryIsec.assign(n, RptIntersection{ { 0, 0 }, InvalidHitIndex, 0 });
	cur = 0;
}

void Frame2D::clear() { create(width, height); }

void* Frame2D::bufferPtr(RptBufferId id, size_t* bytes) {
	size_t n = size_t(width) * height;
	switch (id) {
	case RPT_BUF_DIRECT_OUTPUT: *bytes = n * 16; return directOutput.data();
	case RPT_BUF_INDIRECT_OUTPUT: *bytes = n * 16; return indirectOutput.data();
	case RPT_BUF_DEPTH_NORMAL: *bytes = n * 16; return depthNormal[cur].data();
	case RPT_BUF_DEPTH_NORMAL_PREV: *bytes = n * 16; return depthNormal[cur ^ 1].data();
	case RPT_BUF_ALBEDO_MATID: *bytes = n * 8; return albedoMatId[cur].data();
	case RPT_BUF_ALBEDO_MATID_PREV: *bytes = n * 8; return albedoMatId[cur ^ 1].data();
	case RPT_BUF_MOTION: *bytes = n * 8; return motion.data();
	case RPT_BUF_DI_THIS: *bytes = n * 64; return di[cur].data();
	case RPT_BUF_DI_PREV: *bytes = n * 64; return di[cur ^ 1].data();
	case RPT_BUF_DI_TEMP: *bytes = n * 64; return diTemp.data();
	case RPT_BUF_GI_THIS: *bytes = n * 48; return gi[cur].data();
	case RPT_BUF_GI_PREV: *bytes = n * 48; return gi[cur ^ 1].data();
	case RPT_BUF_GRIS_THIS: *bytes = n * 96; return gris[cur].data();
	case RPT_BUF_GRIS_PREV: *bytes = n * 96; return gris[cur ^ 1].data();
	case RPT_BUF_GRIS_TEMP: *bytes = n * 96; return grisTemp.data();
	case RPT_BUF_PRIMARY_ISEC: *bytes = n * 16; return primaryIsec.data();
	default: *bytes = 0; return nullptr;
	}
}

} // namespace orc

struct OrcScene { Scene scene; };
struct OrcFrame { Frame2D frame; };

static int gThreads = 0;

static void parallelRows(uint32_t height, const std::function<void(uint32_t, uint32_t)>& fn) {
	int nt = gThreads > 0 ? gThreads : int(std::thread::hardware_concurrency());
	if (nt < 1) nt = 1;
	const uint32_t chunk = 4;
	std::atomic<uint32_t> next{ 0 };
	auto worker = [&]() {
		for (;;) {
			uint32_t y0 = next.fetch_add(chunk);
			if (y0 >= height) break;
			fn(y0, std::min(y0 + chunk, height));
		}
	};
	if (nt == 1) { worker(); return; }
	std::vector<std::thread> pool;
	for (int i = 0; i < nt; i++) pool.emplace_back(worker);
	for (auto& t : pool) t.join();
}

extern "C" {

int orc_set_threads(int n) { gThreads = n; return int(std::thread::hardware_concurrency()); }

OrcScene* orc_scene_create(const RptSceneDesc* desc) {
	OrcScene* s = new OrcScene;
	s->scene.build(*desc);
	return s;
}
void orc_scene_destroy(OrcScene* s) { delete s; }
// the oracle's twin of rpt_scene_update_instances / rpt_scene_end_motion: a scene created from the NEW placements is told the
// previous ones (numInstances == 0 ends the motion)
void orc_scene_set_prev_instances(OrcScene* s, const RptObjectInstance* prev, uint32_t numInstances) {
	s->scene.prevInstances.assign(prev, prev + numInstances);
}
void orc_scene_set_brute_force(OrcScene* s, int on) { s->scene.bruteForce = on != 0; }
uint32_t orc_scene_num_triangles(const OrcScene* s) { return s->scene.numFlatTris; }

OrcFrame* orc_frame_create(uint32_t w, uint32_t h) {
	OrcFrame* f = new OrcFrame;
	f->frame.create(w, h);
	return f;
}
void orc_frame_destroy(OrcFrame* f) { delete f; }
void orc_frame_clear(OrcFrame* f) { f->frame.clear(); }
void orc_frame_flip(OrcFrame* f) { f->frame.cur ^= 1u; }
void orc_set_camera(OrcFrame* f, const RptCamera* cur, const RptCamera* prev) { f->frame.camera = *cur; f->frame.prevCamera = *prev; }

#define ORC_PASS(NAME, CALL) \
	void NAME { Frame2D& fr = f->frame; const Scene& sc = s->scene; parallelRows(fr.height, [&](uint32_t y0, uint32_t y1) { CALL; }); }

ORC_PASS(orc_gbuffer(OrcFrame* f, const OrcScene* s), passGBuffer(sc, fr, y0, y1))
ORC_PASS(orc_di_naive(OrcFrame* f, const OrcScene* s), passDINaive(sc, fr, y0, y1))
ORC_PASS(orc_di_naive_rt(OrcFrame* f, const OrcScene* s), passDINaiveRT(sc, fr, y0, y1))
ORC_PASS(orc_gi_naive(OrcFrame* f, const OrcScene* s), passGINaive(sc, fr, y0, y1))
ORC_PASS(orc_di_pathgen(OrcFrame* f, const OrcScene* s, const RptDISettings* st), passDIPathGen(sc, fr, *st, y0, y1))
ORC_PASS(orc_di_temporal(OrcFrame* f, const OrcScene* s, const RptDISettings* st), passDITemporal(sc, fr, *st, y0, y1))
ORC_PASS(orc_di_spatial(OrcFrame* f, const OrcScene* s, const RptDISettings* st), passDISpatial(sc, fr, *st, y0, y1))
ORC_PASS(orc_gi_restir(OrcFrame* f, const OrcScene* s), passGIReSTIR(sc, fr, y0, y1))
ORC_PASS(orc_gris_pathtrace(OrcFrame* f, const OrcScene* s, const RptGRISSettings* st), passGRISPathTrace(sc, fr, *st, y0, y1))
ORC_PASS(orc_gris_temporal(OrcFrame* f, const OrcScene* s, const RptGRISSettings* st), passGRISTemporal(sc, fr, *st, y0, y1))
ORC_PASS(orc_gris_spatial(OrcFrame* f, const OrcScene* s, const RptGRISSettings* st), passGRISSpatial(sc, fr, *st, y0, y1))
ORC_PASS(orc_visualize_as(OrcFrame* f, const OrcScene* s), passVisualizeAS(sc, fr, y0, y1))

void orc_postprocess(OrcFrame* f, const RptPostSettings* st, uint8_t* rgba8Out) {
	Frame2D& fr = f->frame;
	parallelRows(fr.height, [&](uint32_t y0, uint32_t y1) { passPostProcess(fr, *st, rgba8Out, y0, y1); });
}

int orc_read(OrcFrame* f, int id, void* dst, size_t bytes) {
	size_t have;
	void* p = f->frame.bufferPtr(RptBufferId(id), &have);
	if (!p || bytes != have) return -1;
	std::memcpy(dst, p, bytes);
	return 0;
}
int orc_write(OrcFrame* f, int id, const void* src, size_t bytes) {
	size_t have;
	void* p = f->frame.bufferPtr(RptBufferId(id), &have);
	if (!p || bytes != have) return -1;
	std::memcpy(p, src, bytes);
	return 0;
}

// rays: n x {ox,oy,oz,tmin, dx,dy,dz,tmax}
void orc_trace_closest(const OrcScene* s, const float* rays, uint32_t n, RptIntersection* out) {
	parallelRows(n, [&](uint32_t a, uint32_t b) {
		for (uint32_t i = a; i < b; i++) {
			const float* r = rays + size_t(i) * 8;
			out[i] = fromIsec(s->scene.traceClosestHit(V3(r[0], r[1], r[2]), r[3], V3(r[4], r[5], r[6]), r[7]));
		}
	});
}
void orc_trace_shadow(const OrcScene* s, const float* rays, uint32_t n, uint8_t* out) {
	parallelRows(n, [&](uint32_t a, uint32_t b) {
		for (uint32_t i = a; i < b; i++) {
			const float* r = rays + size_t(i) * 8;
			out[i] = s->scene.traceShadow(V3(r[0], r[1], r[2]), r[3], V3(r[4], r[5], r[6]), r[7]) ? 1 : 0;
		}
	});
}

void orc_counters_reset(const OrcScene* s) { s->scene.counters.closestRays = 0; s->scene.counters.shadowRays = 0; }
void orc_counters_read(const OrcScene* s, uint64_t* closest, uint64_t* shadow) {
	*closest = s->scene.counters.closestRays; *shadow = s->scene.counters.shadowRays;
}

// known-answer helpers for the integer RNG and small math (golden-vector tests)
uint32_t orc_hash2(uint32_t v) { return hash2(v); }
uint32_t orc_make_seed(uint32_t seed, uint32_t x, uint32_t y) { return makeSeed(seed, uvec2{ x, y }); }
float orc_sample1f(uint32_t* rng) { return sample1f(*rng); }
void orc_sincos(float x, float* s, float* c) { sincos_(x, *s, *c); }
float orc_round_through_half(float f) { return roundThroughHalf(f); }
void orc_concentric_disk(float u, float v, float* out) { vec2 d = toConcentricDisk({ u, v }); out[0] = d.x; out[1] = d.y; }
void orc_eval_bsdf(const RptMaterial* m, const float* albedo, const float* n, const float* wo, const float* wi, float* out3, float* pdf) {
	vec3 f = evalBSDF(*m, V3(albedo), V3(n), V3(wo), V3(wi));
	out3[0] = f.x; out3[1] = f.y; out3[2] = f.z;
	*pdf = evalPdf(*m, V3(n), V3(wo), V3(wi));
}
// light_sampling.glsl:24-53 on the scene's light table (for the pin against the reference's own text, tests/test_cpu_ref_pins.py)
void orc_sample_light(const OrcScene* s, const float* ref, const float* r4, float* radiance, float* wi, float* dist, float* pdf,
                      float* jacobian, float* bary, uint32_t* id) {
	LightSample L = sampleLight(s->scene, V3(ref), vec4{ r4[0], r4[1], r4[2], r4[3] });
	radiance[0] = L.radiance.x; radiance[1] = L.radiance.y; radiance[2] = L.radiance.z;
	wi[0] = L.wi.x; wi[1] = L.wi.y; wi[2] = L.wi.z;
	*dist = L.dist; *pdf = L.pdf; *jacobian = L.jacobian; bary[0] = L.bary.x; bary[1] = L.bary.y; *id = L.id;
}
int orc_is_bsdf_delta(const RptMaterial* m) { return isBSDFDelta(*m) ? 1 : 0; }
int orc_is_bsdf_connectible(const RptMaterial* m) { return isBSDFConnectible(*m) ? 1 : 0; }
int orc_sample_bsdf(const RptMaterial* m, const float* albedo, const float* n, const float* wo, const float* r3, float* wi, float* bsdf, float* pdf, uint32_t* type) {
	BSDFSample s;
	bool ok = sampleBSDF(*m, V3(albedo), V3(n), V3(wo), V3(r3), s);
	wi[0] = s.wi.x; wi[1] = s.wi.y; wi[2] = s.wi.z;
	bsdf[0] = s.bsdf.x; bsdf[1] = s.bsdf.y; bsdf[2] = s.bsdf.z;
	*pdf = s.pdf; *type = s.type;
	return ok ? 1 : 0;
}

// ---- "driver" call-backs for the reference's shaders compiled for the CPU (oracle/ref/ref_shaders.cpp, RefDriver) ----------------
// What the Vulkan driver supplies to the reference — ray / triangle intersection, texture and G-buffer filtering — is supplied to
// its shader text here by the oracle's definitions of them, so that tests/test_cpu_ref_shaders.py compares shader text with
// restatement and nothing else.
struct OrcDriverCtx { const OrcScene* scene; const OrcFrame* frame; };
OrcDriverCtx* orc_driver_create(const OrcScene* s, const OrcFrame* f) { return new OrcDriverCtx{ s, f }; }
void orc_driver_destroy(OrcDriverCtx* c) { delete c; }
int orc_cb_trace_closest(void* user, const float* o, float tmin, const float* d, float tmax, float* bary, uint32_t* instance, uint32_t* primitive) {
	const Scene& sc = static_cast<OrcDriverCtx*>(user)->scene->scene;
	Intersection i = sc.traceClosestHit(V3(o), tmin, V3(d), tmax);
	if (i.instanceIdx == InvalidHitIndex) return 0;
	bary[0] = i.bary.x; bary[1] = i.bary.y; *instance = i.instanceIdx; *primitive = i.triangleIdx;
	return 1;
}
int orc_cb_trace_any(void* user, const float* o, float tmin, const float* d, float tmax) {
	return static_cast<OrcDriverCtx*>(user)->scene->scene.traceShadow(V3(o), tmin, V3(d), tmax) ? 1 : 0;
}
uint32_t orc_cb_count_candidates(void* user, const float* o, const float* d) {
	return static_cast<OrcDriverCtx*>(user)->scene->scene.countCandidates(V3(o), V3(d));
}
void orc_cb_sample_texture(void* user, uint32_t tex, float u, float v, float* rgb) {
	vec3 c = static_cast<OrcDriverCtx*>(user)->scene->scene.sampleTexture(tex, u, v);
	rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
void orc_cb_sample_depth_normal(void* user, int which, float u, float v, float* out4) {
	const Frame2D& f = static_cast<OrcDriverCtx*>(user)->frame->frame;
	vec4 r = fetchDepthNormalBilinear(f.depthNormal[which ? (f.cur ^ 1u) : f.cur], f.width, f.height, vec2{ u, v });
	out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
// the frame's buffers in place (the reference's shaders write their reservoirs / outputs into a second oracle frame)
void* orc_frame_ptr(OrcFrame* f, int id) { size_t bytes; return f->frame.bufferPtr(RptBufferId(id), &bytes); }

} // extern "C"
