// host/tpt_math.h -- value types of the scene-description front end.
//
// Same public surface as the reference's headers/vec3.h:6-157, headers/ray.h:5-18 and
// headers/aabb.h:7-21 (names, constructors, public data members), written from scratch.
// Nothing here runs on the hot path: intersection and shading live in csrc/ on the GPU; the
// host only needs these types to DESCRIBE a scene, build bounding boxes and the BVH.
#ifndef TPT_HOST_MATH_H_
#define TPT_HOST_MATH_H_

#include <algorithm>
#include <cmath>
#include <iostream>

struct vec3 {
  float e[3];

  vec3() : e{0.0f, 0.0f, 0.0f} {}
  vec3(float a, float b, float c) : e{a, b, c} {}
  vec3(float s) : e{s, s, s} {}

  float x() const { return e[0]; }
  float y() const { return e[1]; }
  float z() const { return e[2]; }
  float r() const { return e[0]; }
  float g() const { return e[1]; }
  float b() const { return e[2]; }

  const vec3 &operator+() const { return *this; }
  vec3 operator-() const { return vec3(-e[0], -e[1], -e[2]); }
  float operator[](int i) const { return e[i]; }
  float &operator[](int i) { return e[i]; }

#define TPT_VEC3_COMPOUND(OP)                                                                      \
  vec3 &operator OP(const vec3 &o) {                                                               \
    for (int i = 0; i < 3; ++i) e[i] OP o.e[i];                                                    \
    return *this;                                                                                  \
  }
  TPT_VEC3_COMPOUND(+=)
  TPT_VEC3_COMPOUND(-=)
  TPT_VEC3_COMPOUND(*=)
  TPT_VEC3_COMPOUND(/=)
#undef TPT_VEC3_COMPOUND
  vec3 &operator*=(float s) {
    for (float &c : e) c *= s;
    return *this;
  }
  vec3 &operator/=(float s) {
    for (float &c : e) c /= s;
    return *this;
  }

  float squared_length() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
  float length() const { return std::sqrt(squared_length()); }
  void make_unit_vector() {
    float k = 1.0 / length(); // double reciprocal rounded to float, as headers/vec3.h:96
    *this *= k;
  }
  void normalize() { make_unit_vector(); }
};

#define TPT_VEC3_BINARY(OP)                                                                        \
  inline vec3 operator OP(const vec3 &a, const vec3 &b) {                                          \
    return vec3(a.e[0] OP b.e[0], a.e[1] OP b.e[1], a.e[2] OP b.e[2]);                             \
  }
TPT_VEC3_BINARY(+)
TPT_VEC3_BINARY(-)
TPT_VEC3_BINARY(*)
TPT_VEC3_BINARY(/)
#undef TPT_VEC3_BINARY

inline vec3 operator*(const vec3 &a, float s) { return vec3(a.e[0] * s, a.e[1] * s, a.e[2] * s); }
inline vec3 operator*(float s, const vec3 &a) { return a * s; }
inline vec3 operator/(const vec3 &a, float s) { return vec3(a.e[0] / s, a.e[1] / s, a.e[2] / s); }

inline float dot(const vec3 &a, const vec3 &b) {
  return a.e[0] * b.e[0] + a.e[1] * b.e[1] + a.e[2] * b.e[2];
}
inline vec3 cross(const vec3 &a, const vec3 &b) {
  return vec3(a.e[1] * b.e[2] - b.e[1] * a.e[2], a.e[2] * b.e[0] - b.e[2] * a.e[0],
              a.e[0] * b.e[1] - b.e[0] * a.e[1]);
}
inline vec3 normalize(vec3 v) { return v / v.length(); }
inline vec3 unit_vector(vec3 v) { return normalize(v); }
inline vec3 vec_min(const vec3 &a, const vec3 &b) {
  return vec3(std::min(a.x(), b.x()), std::min(a.y(), b.y()), std::min(a.z(), b.z()));
}
inline vec3 vec_max(const vec3 &a, const vec3 &b) {
  return vec3(std::max(a.x(), b.x()), std::max(a.y(), b.y()), std::max(a.z(), b.z()));
}
inline std::istream &operator>>(std::istream &is, vec3 &v) { return is >> v.e[0] >> v.e[1] >> v.e[2]; }
inline std::ostream &operator<<(std::ostream &os, const vec3 &v) {
  return os << v.e[0] << " " << v.e[1] << " " << v.e[2];
}

inline float float_min(float a, float b) { return a < b ? a : b; }
inline float float_max(float a, float b) { return a > b ? a : b; }

class ray {
public:
  ray() {}
  ray(const vec3 &pos, const vec3 &direction, float time = 0.0f)
      : pos_(pos), direction_(direction), time_(time) {}
  vec3 origin() const { return pos_; }
  vec3 direction() const { return direction_; }
  float time() const { return time_; }
  vec3 point_at_parameter(float t) const { return pos_ + t * direction_; }

  vec3 pos_, direction_;
  float time_ = 0.0f;
};

// Bounds only: the slab test (reference src/aabb.cc:3-19) runs on the GPU (csrc/tpt_device.cuh).
class AABB {
public:
  AABB() {}
  AABB(const vec3 &lo, const vec3 &hi) : min_(lo), max_(hi) {}
  vec3 min() const { return min_; }
  vec3 max() const { return max_; }
  vec3 min_, max_;
};

inline AABB surrounding_box(AABB a, AABB b) {
  return AABB(vec3(float_min(a.min().x(), b.min().x()), float_min(a.min().y(), b.min().y()),
                   float_min(a.min().z(), b.min().z())),
              vec3(float_max(a.max().x(), b.max().x()), float_max(a.max().y(), b.max().y()),
                   float_max(a.max().z(), b.max().z())));
}

#endif // TPT_HOST_MATH_H_
