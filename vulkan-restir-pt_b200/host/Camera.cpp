#include "Camera.h"

namespace rpt {

static void store(float* dst, const mat4& m) { std::memcpy(dst, &m.c[0].x, 64); }
static void store(float* dst, vec3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

Camera::Camera(vec3 pos, vec3 angle) {
	std::memset(&d, 0, sizeof(d));
	// defaults of reference src/Camera.h:53-69
	d.FOV = 45.0f;
	d.nearZ = 1e-3f;
	d.farZ = 1e3f;
	d.lensRadius = 0.0f;
	d.focalDist = 1.0f;
	d.up[2] = 1.0f;
	d.filmSize[0] = d.filmSize[1] = 1;
	d.seed = 0;  // uninitialised in the reference (src/Camera.h:69); defined as 0 here
	store(d.pos, pos);
	store(d.angle, angle);
	update();
}

void Camera::rotate(vec3 a) {
	d.angle[0] += a.x; d.angle[1] += a.y; d.angle[2] += a.z;
	d.angle[1] = clampf(d.angle[1], -89.f, 89.f);
	update();
}

void Camera::setFOV(float fov) {
	d.FOV = clampf(fov, .1f, 90.f);
	update();
}

// reference src/Camera.cpp:39-49
void Camera::setDir(vec3 dir) {
	dir = normalize(dir);
	d.angle[1] = degrees(std::asin(dir.z / length(dir)));
	float lenXY = std::sqrt(dir.x * dir.x + dir.y * dir.y);
	d.angle[0] = degrees(std::asin(dir.y / lenXY)) - 90.f;
	if (dir.x < 0) {
		d.angle[0] += 360.f;
	}
	update();
}

void Camera::setPos(vec3 p) { store(d.pos, p); update(); }
void Camera::setAngle(vec3 a) { store(d.angle, a); update(); }
void Camera::setFilmSize(uint32_t w, uint32_t h) { d.filmSize[0] = w; d.filmSize[1] = h; update(); }
void Camera::setPlanes(float n, float f) { d.nearZ = n; d.farZ = f; update(); }

void Camera::nextFrame(uint32_t seed) {
	d.seed = seed;
	std::memcpy(d.lastProjView, d.projView, 64);
	d.frameIndex++;
}

// reference src/Camera.cpp:78-98
void Camera::update() {
	float ax = radians(d.angle[0]), ay = radians(d.angle[1]);
	float x = std::sin(ax) * std::cos(ay);
	float y = std::cos(ax) * std::cos(ay);
	float z = std::sin(ay);

	vec3 front = normalize(vec3(x, y, z));
	const vec3 u(0.0f, 0.0f, 1.0f);
	vec3 right = normalize(cross(front, u));

	mat4 rot = rpt::rotate(mat4(1.0f), d.angle[2], front);
	vec4 r4 = rot * vec4(right, 1.0f);
	right = normalize(vec3(r4.x, r4.y, r4.z));
	vec3 up = normalize(cross(right, front));

	vec3 p = pos();
	vec3 lookingAt = p + vec3(x, y, z);

	mat4 view = rpt::lookAt(p, lookingAt, up);
	mat4 proj = perspective(radians(d.FOV), aspect(), d.nearZ, d.farZ);
	proj[1][1] *= -1.f;
	mat4 projView = proj * view;

	store(d.front, front);
	store(d.right, right);
	store(d.up, up);
	store(d.view, view);
	store(d.proj, proj);
	store(d.projView, projView);
	d.frameIndex = 0;
}

} // namespace rpt
