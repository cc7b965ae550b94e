// Small host-side vector / matrix library standing in for glm (which the reference
// takes from the Vulkan SDK and does not vendor; SURVEY.md §2.1 "cmake/").
// Column-major mat4 like glm, right-handed, clip z in [-1,1] (glm defaults, no GLM_FORCE_*).
// Only what Camera / Model / Scene need (reference: src/Camera.cpp, src/Model.cpp:11-21, src/Scene.cpp).
#pragma once
#include <cmath>
#include <cstdint>

namespace rpt {

struct vec2 { float x = 0, y = 0; };
struct uvec2 { uint32_t x = 0, y = 0; };

struct vec3 {
	float x = 0, y = 0, z = 0;
	vec3() = default;
	vec3(float s) : x(s), y(s), z(s) {}
	vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
	float& operator[](int i) { return (&x)[i]; }
	float operator[](int i) const { return (&x)[i]; }
};

inline vec3 operator+(vec3 a, vec3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline vec3 operator-(vec3 a, vec3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline vec3 operator-(vec3 a) { return { -a.x, -a.y, -a.z }; }
inline vec3 operator*(vec3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline vec3 operator*(float s, vec3 a) { return a * s; }
inline vec3 operator/(vec3 a, float s) { return { a.x / s, a.y / s, a.z / s }; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a / length(a); }

struct vec4 {
	float x = 0, y = 0, z = 0, w = 0;
	vec4() = default;
	vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
	vec4(vec3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
	float& operator[](int i) { return (&x)[i]; }
	float operator[](int i) const { return (&x)[i]; }
};

inline vec4 operator*(vec4 a, float s) { return { a.x * s, a.y * s, a.z * s, a.w * s }; }
inline vec4 operator+(vec4 a, vec4 b) { return { a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }

// column-major: c[j] is column j; element (row i, col j) = c[j][i]
struct mat4 {
	vec4 c[4];
	mat4() = default;
	explicit mat4(float d) {
		c[0] = { d, 0, 0, 0 }; c[1] = { 0, d, 0, 0 }; c[2] = { 0, 0, d, 0 }; c[3] = { 0, 0, 0, d };
	}
	vec4& operator[](int j) { return c[j]; }
	const vec4& operator[](int j) const { return c[j]; }
};

inline vec4 operator*(const mat4& m, vec4 v) {
	return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w;
}

inline mat4 operator*(const mat4& a, const mat4& b) {
	mat4 r;
	for (int j = 0; j < 4; j++) r[j] = a * b[j];
	return r;
}

inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline float degrees(float rad) { return rad * 57.295779513082320876798154814105f; }

inline mat4 translate(const mat4& m, vec3 v) {
	mat4 r = m;
	r[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
	return r;
}

// glm::rotate semantics: m * R(angle, axis), angle in radians
inline mat4 rotate(const mat4& m, float angle, vec3 axisIn) {
	float cs = std::cos(angle), sn = std::sin(angle);
	vec3 axis = normalize(axisIn);
	vec3 t = axis * (1.0f - cs);
	float R[3][3];
	R[0][0] = cs + t.x * axis.x;          R[0][1] = t.x * axis.y + sn * axis.z; R[0][2] = t.x * axis.z - sn * axis.y;
	R[1][0] = t.y * axis.x - sn * axis.z; R[1][1] = cs + t.y * axis.y;          R[1][2] = t.y * axis.z + sn * axis.x;
	R[2][0] = t.z * axis.x + sn * axis.y; R[2][1] = t.z * axis.y - sn * axis.x; R[2][2] = cs + t.z * axis.z;
	mat4 r;
	r[0] = m[0] * R[0][0] + m[1] * R[0][1] + m[2] * R[0][2];
	r[1] = m[0] * R[1][0] + m[1] * R[1][1] + m[2] * R[1][2];
	r[2] = m[0] * R[2][0] + m[1] * R[2][1] + m[2] * R[2][2];
	r[3] = m[3];
	return r;
}

inline mat4 scale(const mat4& m, vec3 v) {
	mat4 r;
	r[0] = m[0] * v.x; r[1] = m[1] * v.y; r[2] = m[2] * v.z; r[3] = m[3];
	return r;
}

inline mat4 transpose(const mat4& m) {
	mat4 r;
	for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[j][i] = m[i][j];
	return r;
}

// general 4x4 inverse by cofactors
inline mat4 inverse(const mat4& m) {
	const float* a = &m.c[0].x;
	float inv[16];
	inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
	inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
	inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
	inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
	inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
	inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
	inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
	inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
	inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
	inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
	inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
	inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
	inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
	inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
	inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
	inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
	float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
	float id = 1.0f / det;
	mat4 r;
	float* o = &r.c[0].x;
	for (int i = 0; i < 16; i++) o[i] = inv[i] * id;
	return r;
}

// glm::lookAtRH
inline mat4 lookAt(vec3 eye, vec3 center, vec3 up) {
	vec3 f = normalize(center - eye);
	vec3 s = normalize(cross(f, up));
	vec3 u = cross(s, f);
	mat4 r(1.0f);
	r[0][0] = s.x; r[1][0] = s.y; r[2][0] = s.z;
	r[0][1] = u.x; r[1][1] = u.y; r[2][1] = u.z;
	r[0][2] = -f.x; r[1][2] = -f.y; r[2][2] = -f.z;
	r[3][0] = -dot(s, eye); r[3][1] = -dot(u, eye); r[3][2] = dot(f, eye);
	return r;
}

// glm::perspectiveRH_NO (fovy in radians)
inline mat4 perspective(float fovy, float aspect, float zNear, float zFar) {
	float t = std::tan(fovy * 0.5f);
	mat4 r(0.0f);
	r[0][0] = 1.0f / (aspect * t);
	r[1][1] = 1.0f / t;
	r[2][2] = -(zFar + zNear) / (zFar - zNear);
	r[2][3] = -1.0f;
	r[3][2] = -(2.0f * zFar * zNear) / (zFar - zNear);
	return r;
}

inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

} // namespace rpt
