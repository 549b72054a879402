// vec3.hpp -- minimal float3 algebra for the host side of the B200 render path.
//
// Layout-compatible with Eigen::Vector3f (three packed floats), so a maintainer of the reference
// can pass `v.data()` straight through the C ABI.  The reductions follow the
// evaluation order the reference build gets from Eigen 3.3.7 for fixed-size 3-vectors,
// x*x' + (y*y' + z*z'), because the scene bake must reproduce the reference's floats bit for bit
// (SURVEY.md App. A.9).  Compile host code with -ffp-contract=off.
#pragma once

#include <cmath>

namespace rt {

struct Vec3f {
  float x = 0.f, y = 0.f, z = 0.f;
  Vec3f() = default;
  Vec3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  explicit Vec3f(const float *p) : x(p[0]), y(p[1]), z(p[2]) {}
  float &operator[](int i) { return (&x)[i]; }
  float operator[](int i) const { return (&x)[i]; }
  const float *data() const { return &x; }
  float *data() { return &x; }
};

inline Vec3f operator+(Vec3f a, Vec3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3f operator-(Vec3f a, Vec3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3f operator*(float s, Vec3f a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3f operator*(Vec3f a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3f operator/(Vec3f a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline Vec3f &operator+=(Vec3f &a, Vec3f b) { a = a + b; return a; }

inline float dot(Vec3f a, Vec3f b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline Vec3f cross(Vec3f a, Vec3f b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float norm(Vec3f a) { return std::sqrt(dot(a, a)); }
// Eigen's normalized(): divide by the norm, leave the zero vector alone.
inline Vec3f normalized(Vec3f a) {
  float z = dot(a, a);
  if (z > 0.f) return a / std::sqrt(z);
  return a;
}

}  // namespace rt
