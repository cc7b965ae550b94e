// ORACLE — TEST INFRASTRUCTURE ONLY.  Per-frame buffers with the reference's layouts and ping-pong wiring
// (reference src/Renderer.cpp:193-257 createRayImage, :324-347 initDescriptor, src/GBufferPass.h:23-26).
#pragma once
#include <vector>
#include "oracle_scene.h"

namespace orc {

struct Frame2D {
	uint32_t width = 0, height = 0;
	uint32_t cur = 0;   // mCurFrame: which of the ping-pong pair is "this"

	std::vector<vec4> directOutput, indirectOutput;
	std::vector<vec4> depthNormal[2];
	std::vector<uvec2> albedoMatId[2];
	std::vector<vec2> motion;
	std::vector<RptDIReservoir> di[2], diTemp;
	std::vector<RptGIReservoir> gi[2];
	std::vector<RptGRISReservoir> gris[2], grisTemp;
	std::vector<RptIntersection> primaryIsec;

	RptCamera camera{}, prevCamera{};

	void create(uint32_t w, uint32_t h);
	void clear();

	void* bufferPtr(RptBufferId id, size_t* bytes);
};

struct Ray { vec3 ori, dir; };

// all passes; each is the CPU restatement of one reference shader entry point
void passGBuffer(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1);
void passDINaive(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1);
void passDINaiveRT(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1);
void passGINaive(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1);
void passDIPathGen(const Scene& s, Frame2D& f, const RptDISettings& st, uint32_t y0, uint32_t y1);
void passDITemporal(const Scene& s, Frame2D& f, const RptDISettings& st, uint32_t y0, uint32_t y1);
void passDISpatial(const Scene& s, Frame2D& f, const RptDISettings& st, uint32_t y0, uint32_t y1);
void passGIReSTIR(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1);
void passGRISPathTrace(const Scene& s, Frame2D& f, const RptGRISSettings& st, uint32_t y0, uint32_t y1);
void passGRISTemporal(const Scene& s, Frame2D& f, const RptGRISSettings& st, uint32_t y0, uint32_t y1);
void passGRISSpatial(const Scene& s, Frame2D& f, const RptGRISSettings& st, uint32_t y0, uint32_t y1);
void passVisualizeAS(const Scene& s, Frame2D& f, uint32_t y0, uint32_t y1);
void passPostProcess(const Frame2D& f, const RptPostSettings& st, uint8_t* rgba8, uint32_t y0, uint32_t y1);

} // namespace orc
