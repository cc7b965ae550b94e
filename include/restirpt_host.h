/*
 * restirpt_host.h — C view of the C++ host library (librestirpt_host.so): Scene / Camera / alias table /
 * headless Renderer, i.e. the host classes of the reference (src/Scene.*, src/Resource.*, src/Model.*,
 * src/Material.*, src/Camera.*, src/util/AliasTable.h, src/Renderer.*) rebuilt over include/restirpt.h.
 * It exists so that tests, bench.py and other-language callers can drive the same host code; C++ callers
 * use the headers under vulkan-restir-pt_b200/host/ directly.
 */
#ifndef RESTIRPT_HOST_H
#define RESTIRPT_HOST_H

#include "restirpt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RhScene RhScene;
typedef struct RhRenderer RhRenderer;

const char* rh_last_error(void);

/* Scene::load (reference src/Scene.cpp:104-123); NULL on failure */
RhScene* rh_scene_load_xml(const char* path);
/* procedural scenes: BASELINE.json config 1 (Cornell), the VeachAjar stand-in, config 5 (instanced field) */
RhScene* rh_scene_cornell(void);
RhScene* rh_scene_room(uint32_t trisTarget, uint32_t seed);
RhScene* rh_scene_field(uint32_t meshSubdiv, uint32_t gridN, uint32_t seed);
/* the same field with ONE copy of the mesh referenced by every instance (BASELINE config 5's instanced variant) */
RhScene* rh_scene_field_shared(uint32_t meshSubdiv, uint32_t gridN, uint32_t seed);
/* ask rpt_scene_create for BLAS + TLAS (RPT_SCENE_TWO_LEVEL in the desc's flags) instead of the flattened structure */
void rh_scene_set_two_level(RhScene* s, int on);
void rh_scene_destroy(RhScene* s);
void rh_scene_desc(const RhScene* s, RptSceneDesc* out);   /* pointers stay valid until rh_scene_destroy */
void rh_scene_camera(const RhScene* s, RptCamera* out);
uint32_t rh_scene_num_triangles(const RhScene* s);
/* dynamic scenes (new): place object model `objectIdx` anew (translate, scale, rotate in degrees: the XML <transform>
 * attributes, reference src/Model.cpp:11-21); returns 0, or -1 with rh_last_error */
int rh_scene_set_object_transform(RhScene* s, uint32_t objectIdx, const float pos[3], const float scale[3], const float rotDeg[3]);

/* Camera (reference src/Camera.cpp) operating on the raw 352-byte block */
void rh_camera_init(RptCamera* cam, const float pos[3], const float angleDeg[3], float fovDeg,
                    uint32_t width, uint32_t height, float nearZ, float farZ);
void rh_camera_look_at(RptCamera* cam, const float target[3]);
void rh_camera_set_film(RptCamera* cam, uint32_t width, uint32_t height);
void rh_camera_set_planes(RptCamera* cam, float nearZ, float farZ);
void rh_camera_move(RptCamera* cam, const float delta[3]);
void rh_camera_update(RptCamera* cam);                      /* Camera::update: frameIndex = 0 */
void rh_camera_next_frame(RptCamera* cam, uint32_t seed);   /* Camera::nextFrame */

/* DiscreteSampler1D<float>::build (reference src/util/AliasTable.h:26-71); out has n+1 entries */
void rh_build_alias_table(const float* power, uint32_t n, RptLightSampleTableElement* out);

/* headless Renderer (reference src/Renderer.cpp drawFrame path); NULL on failure */
RhRenderer* rh_renderer_create(const RhScene* s, uint32_t width, uint32_t height, int cudaDevice,
                               uint32_t rowBegin, uint32_t rowEnd, uint32_t halo);
void rh_renderer_destroy(RhRenderer* r);
void rh_renderer_set_methods(RhRenderer* r, int directMethod, int indirectMethod, int toneMapping,
                             int correctGamma, int accumulate);
/* RayTracing::Mode of the reference (src/RayTracing.h:28-30): 0 = RayQuery (.comp shaders, default), 1 = RayTracing (.rgen) */
void rh_renderer_set_pipeline_mode(RhRenderer* r, int mode);
void rh_renderer_set_gris(RhRenderer* r, const RptGRISSettings* st);
void rh_renderer_set_di(RhRenderer* r, const RptDISettings* st);
void rh_renderer_clear_reservoirs(RhRenderer* r);
void rh_renderer_camera_move(RhRenderer* r, const float delta[3]);
void rh_renderer_camera(RhRenderer* r, RptCamera* out);
typedef void (*RhHaloExchangeFn)(void* user, RptFrame* frame, RptBufferId buffer);
void rh_renderer_set_halo_exchange(RhRenderer* r, RhHaloExchangeFn fn, void* user);
/* push the scene's object instances to the device and rebuild the acceleration structure (rpt_scene_update_instances) */
int rh_renderer_update_instances(RhRenderer* r, const RhScene* s);
int rh_renderer_draw_frame(RhRenderer* r, uint32_t seed, uint8_t* rgba8Out);   /* 0 ok, -1 error */
/* the same frame with a pipelined read-back (rpt_postprocess_async): returns when the frame is enqueued; rgba8Out (pinned host
 * memory; rotate through three buffers) is complete after rh_renderer_wait_readback(ticket) */
int rh_renderer_draw_frame_async(RhRenderer* r, uint32_t seed, uint8_t* rgba8Out, uint64_t* ticket);
int rh_renderer_wait_readback(RhRenderer* r, uint64_t ticket);
/* One frame on all strips of a film that live in this process (connected with rpt_frame_connect_peers): stage by stage —
 * every strip's G-buffer / candidate / temporal passes first, then every strip's spatial pass, post-process and flip — so that
 * each device-side hand-over wait finds its signal already enqueued.  rh_renderer_draw_frame refuses such strips.
 * rgba8Outs: NULL, or one pointer (possibly NULL) per strip. */
int rh_draw_strips(RhRenderer* const* strips, uint32_t count, uint32_t seed, uint8_t* const* rgba8Outs);
RptFrame* rh_renderer_frame(RhRenderer* r);
RptScene* rh_renderer_scene(RhRenderer* r);
RptCtx* rh_renderer_ctx(RhRenderer* r);

/* canonical text dump of the element tree the host's XML reader sees ("<depth> <name> <attr>=<value> ...\n"); returns the bytes
 * needed including the terminator (0 on error) and writes at most `capacity` bytes.  Test aid: compared with the reference's
 * parser (pugixml) by tests/test_cpu_ref_pins.py */
size_t rh_xml_dump(const char* path, char* out, size_t capacity);
int rh_write_png(const char* path, const uint8_t* rgba8, uint32_t width, uint32_t height);
/* Texture decoding of the scene front-end: zvk::HostImage::createFromFile(path, Int8, filter, 4) (reference src/Resource.cpp:26,
   stb_image underneath).  PNG, JPEG (baseline and progressive) or binary PPM -> width x height RGBA8, released with
   rh_free_image; NULL on failure (rh_last_error). */
uint8_t* rh_read_image(const char* path, uint32_t* width, uint32_t* height);
void rh_free_image(uint8_t* rgba8);

#ifdef __cplusplus
}
#endif
#endif
