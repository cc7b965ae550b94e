// C view of the host library, see include/restirpt_host.h
#include "../../include/restirpt_host.h"
#include "Renderer.h"

#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <sstream>
#include "XmlLite.h"
#include <string>
#include <vector>

using namespace rpt;

struct RhScene { Scene scene; };
struct RhRenderer { Renderer* r; };

static thread_local std::string gLastError;

template <typename F>
static RhScene* makeScene(F&& fill) {
	RhScene* s = nullptr;
	try {
		s = new RhScene;
		fill(s->scene);
		return s;
	}
	catch (const std::exception& e) {
		gLastError = e.what();
		delete s;
		return nullptr;
	}
}

static Camera wrap(const RptCamera* c) { Camera cam; cam.data() = *c; return cam; }

extern "C" {

const char* rh_last_error(void) { return gLastError.c_str(); }

RhScene* rh_scene_load_xml(const char* path) { return makeScene([&](Scene& s) { s.load(path); }); }
RhScene* rh_scene_cornell(void) { return makeScene([&](Scene& s) { makeCornellBox(s); }); }
RhScene* rh_scene_room(uint32_t tris, uint32_t seed) { return makeScene([&](Scene& s) { makeAjarLikeRoom(s, tris, seed); }); }
RhScene* rh_scene_field(uint32_t subdiv, uint32_t gridN, uint32_t seed) { return makeScene([&](Scene& s) { makeInstancedField(s, subdiv, gridN, seed); }); }
RhScene* rh_scene_field_shared(uint32_t subdiv, uint32_t gridN, uint32_t seed) { return makeScene([&](Scene& s) { makeInstancedField(s, subdiv, gridN, seed, true); }); }
void rh_scene_set_two_level(RhScene* s, int on) { s->scene.twoLevel = on != 0; }
void rh_scene_destroy(RhScene* s) { delete s; }
void rh_scene_desc(const RhScene* s, RptSceneDesc* out) { *out = s->scene.desc(); }
void rh_scene_camera(const RhScene* s, RptCamera* out) { *out = s->scene.camera.data(); }
uint32_t rh_scene_num_triangles(const RhScene* s) { return s->scene.numTriangles(); }
int rh_scene_set_object_transform(RhScene* s, uint32_t objectIdx, const float pos[3], const float scale[3], const float rotDeg[3]) {
	try {
		s->scene.setObjectTransform(objectIdx, vec3(pos[0], pos[1], pos[2]), vec3(scale[0], scale[1], scale[2]), vec3(rotDeg[0], rotDeg[1], rotDeg[2]));
		return 0;
	}
	catch (const std::exception& e) { gLastError = e.what(); return -1; }
}

void rh_camera_init(RptCamera* cam, const float pos[3], const float angle[3], float fov,
                    uint32_t w, uint32_t h, float nearZ, float farZ) {
	Camera c(vec3(pos[0], pos[1], pos[2]), vec3(angle[0], angle[1], angle[2]));
	c.setFOV(fov);
	c.setFilmSize(w, h);
	c.setPlanes(nearZ, farZ);
	std::memcpy(c.data().lastProjView, c.data().projView, 64);
	*cam = c.data();
}
void rh_camera_look_at(RptCamera* cam, const float t[3]) { Camera c = wrap(cam); c.lookAt(vec3(t[0], t[1], t[2])); *cam = c.data(); }
void rh_camera_set_film(RptCamera* cam, uint32_t w, uint32_t h) { Camera c = wrap(cam); c.setFilmSize(w, h); *cam = c.data(); }
void rh_camera_set_planes(RptCamera* cam, float n, float f) { Camera c = wrap(cam); c.setPlanes(n, f); *cam = c.data(); }
void rh_camera_move(RptCamera* cam, const float d[3]) { Camera c = wrap(cam); c.move(vec3(d[0], d[1], d[2])); *cam = c.data(); }
void rh_camera_update(RptCamera* cam) { Camera c = wrap(cam); c.update(); *cam = c.data(); }
void rh_camera_next_frame(RptCamera* cam, uint32_t seed) { Camera c = wrap(cam); c.nextFrame(seed); *cam = c.data(); }

void rh_build_alias_table(const float* power, uint32_t n, RptLightSampleTableElement* out) {
	auto t = buildAliasTable(std::vector<float>(power, power + n));
	std::memcpy(out, t.data(), t.size() * sizeof(RptLightSampleTableElement));
}

RhRenderer* rh_renderer_create(const RhScene* s, uint32_t w, uint32_t h, int dev,
                               uint32_t rowBegin, uint32_t rowEnd, uint32_t halo) {
	try {
		return new RhRenderer{ new Renderer(s->scene, w, h, dev, rowBegin, rowEnd, halo) };
	}
	catch (const std::exception& e) {
		gLastError = e.what();
		return nullptr;
	}
}
void rh_renderer_destroy(RhRenderer* r) { if (r) { delete r->r; delete r; } }
void rh_renderer_set_methods(RhRenderer* r, int d, int i, int tm, int gamma, int acc) {
	r->r->settings.directMethod = d; r->r->settings.indirectMethod = i; r->r->settings.toneMapping = tm;
	r->r->settings.correctGamma = gamma != 0; r->r->settings.accumulate = acc != 0;
}
void rh_renderer_set_pipeline_mode(RhRenderer* r, int mode) { r->r->settings.pipelineMode = mode == 1 ? 1 : 0; }
void rh_renderer_set_gris(RhRenderer* r, const RptGRISSettings* st) { r->r->grisSettings = *st; }
void rh_renderer_set_di(RhRenderer* r, const RptDISettings* st) { r->r->diSettings = *st; }
void rh_renderer_clear_reservoirs(RhRenderer* r) { r->r->clearReservoirs(); }
void rh_renderer_camera_move(RhRenderer* r, const float d[3]) { r->r->camera().move(vec3(d[0], d[1], d[2])); }
void rh_renderer_camera(RhRenderer* r, RptCamera* out) { *out = r->r->camera().data(); }
void rh_renderer_set_halo_exchange(RhRenderer* r, RhHaloExchangeFn fn, void* user) { r->r->setHaloExchange(fn, user); }
int rh_renderer_update_instances(RhRenderer* r, const RhScene* s) {
	try { r->r->updateInstances(s->scene); return 0; }
	catch (const std::exception& e) { gLastError = e.what(); return -1; }
}
int rh_renderer_draw_frame(RhRenderer* r, uint32_t seed, uint8_t* rgba8Out) {
	try {
		r->r->drawFrame(seed, rgba8Out);
		return 0;
	}
	catch (const std::exception& e) {
		gLastError = e.what();
		return -1;
	}
}
int rh_renderer_draw_frame_async(RhRenderer* r, uint32_t seed, uint8_t* rgba8Out, uint64_t* ticket) {
	try {
		*ticket = r->r->drawFrameAsync(seed, rgba8Out);
		return 0;
	}
	catch (const std::exception& e) {
		gLastError = e.what();
		return -1;
	}
}
int rh_renderer_wait_readback(RhRenderer* r, uint64_t ticket) {
	try {
		r->r->waitReadback(ticket);
		return 0;
	}
	catch (const std::exception& e) {
		gLastError = e.what();
		return -1;
	}
}
int rh_draw_strips(RhRenderer* const* strips, uint32_t count, uint32_t seed, uint8_t* const* rgba8Outs) {
	try {
		std::vector<Renderer*> rs(count);
		for (uint32_t i = 0; i < count; i++) rs[i] = strips[i]->r;
		Renderer::drawStrips(rs.data(), count, seed, rgba8Outs);
		return 0;
	}
	catch (const std::exception& e) {
		gLastError = e.what();
		return -1;
	}
}
RptFrame* rh_renderer_frame(RhRenderer* r) { return r->r->frame(); }
RptScene* rh_renderer_scene(RhRenderer* r) { return r->r->deviceScene(); }
RptCtx* rh_renderer_ctx(RhRenderer* r) { return r->r->ctx(); }

// canonical dump of a scene file's element tree as host/XmlLite.h reads it ("<depth> <name> <attr>=<value> ...\n", document order):
// compared in the tests with the same dump made by the reference's parser (pugixml, oracle/ref/ref_pugi.cpp)
static void dumpXml(const XmlNode& n, int depth, std::string& out) {
	out += std::to_string(depth) + " " + n.name;
	for (auto& a : n.attrs) out += " " + a.first + "=" + a.second;
	out += "\n";
	for (auto& c : n.children) dumpXml(*c, depth + 1, out);
}
size_t rh_xml_dump(const char* path, char* out, size_t capacity) {
	try {
		std::ifstream f(path, std::ios::binary);
		if (!f) { gLastError = std::string("rh_xml_dump: cannot open ") + path; return 0; }
		std::stringstream ss;
		ss << f.rdbuf();
		const std::string text = ss.str();
		auto root = XmlParser(text).parseDocument();
		std::string s;
		dumpXml(*root, 0, s);
		if (out && capacity) { std::strncpy(out, s.c_str(), capacity - 1); out[capacity - 1] = 0; }
		return s.size() + 1;
	}
	catch (const std::exception& e) { gLastError = e.what(); return 0; }
}

int rh_write_png(const char* path, const uint8_t* rgba8, uint32_t w, uint32_t h) { return writePNG(path, rgba8, w, h) ? 0 : -1; }
uint8_t* rh_read_image(const char* path, uint32_t* width, uint32_t* height) {
	HostImage img;
	std::string err;
	if (!path || !width || !height || !readImage(path, img, &err)) { gLastError = err.empty() ? "rh_read_image: NULL argument" : err; return nullptr; }
	uint8_t* out = static_cast<uint8_t*>(std::malloc(img.rgba8.size()));
	if (!out) { gLastError = "rh_read_image: out of memory"; return nullptr; }
	std::memcpy(out, img.rgba8.data(), img.rgba8.size());
	*width = img.width; *height = img.height;
	return out;
}
void rh_free_image(uint8_t* rgba8) { std::free(rgba8); }

} // extern "C"
