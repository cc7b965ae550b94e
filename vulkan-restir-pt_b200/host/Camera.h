// Host camera.  Mirrors the reference Camera (src/Camera.h:10-70, src/Camera.cpp): the object's first 352
// bytes ARE the uniform block the device reads (RptCamera, include/restirpt.h), so it is memcpy'd as-is
// exactly like reference src/Renderer.cpp:358-361.  World is Z-up (src/Camera.cpp:84,93).
#pragma once
#include <cstring>
#include "rmath.h"
#include "../../include/restirpt.h"

namespace rpt {

class Camera {
public:
	Camera(vec3 pos = vec3(0, 0, 0), vec3 angle = vec3(90.0f, 0.0f, 0.0f));

	void move(vec3 v) { setPos(pos() + v); }
	void rotate(vec3 a);
	void setFOV(float fov);
	void lookAt(vec3 focus) { setDir(focus - pos()); }
	void setDir(vec3 dir);
	void setPos(vec3 p);
	void setAngle(vec3 a);
	void setFilmSize(uint32_t w, uint32_t h);
	void setPlanes(float nearZ, float farZ);
	void setLensRadius(float r) { d.lensRadius = r; }
	void setFocalDist(float f) { d.focalDist = f; }

	vec3 pos() const { return { d.pos[0], d.pos[1], d.pos[2] }; }
	vec3 front() const { return { d.front[0], d.front[1], d.front[2] }; }
	vec3 right() const { return { d.right[0], d.right[1], d.right[2] }; }
	vec3 up() const { return { d.up[0], d.up[1], d.up[2] }; }
	float FOV() const { return d.FOV; }
	float aspect() const { return static_cast<float>(d.filmSize[0]) / d.filmSize[1]; }

	// reference Camera::nextFrame (src/Camera.cpp:68-72) draws the seed from std::default_random_engine,
	// which is implementation-defined; here the caller supplies it so runs are reproducible.
	void nextFrame(uint32_t seed);
	void setClearFlag() { d.frameIndex = 0x80000000u; }
	void update();

	const RptCamera& data() const { return d; }
	RptCamera& data() { return d; }

private:
	RptCamera d;
};

} // namespace rpt
