// hvec.h — host-side f32 vector helpers with cgmath 0.16 evaluation order.
// Built with -ffp-contract=off so every expression rounds like rustc's non-fused f32 code.
//   dot       = (ax*bx + ay*by) + az*bz          (cgmath Vector3::dot -> mul_element_wise().sum())
//   normalize = v * (1 / |v|)                    (InnerSpace::normalize_to)
#pragma once
#include <cmath>

namespace b200rt_host {

struct V3 {
    float x, y, z;
};
static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
static inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
static inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
static inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline float magnitude(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 normalize(V3 a) { return a * (1.0f / magnitude(a)); }
static inline void store(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

}  // namespace b200rt_host
