// oracle.cpp — CPU oracle: TEST INFRASTRUCTURE ONLY (see oracle.h for who may load it).
//
// A reference-faithful restatement of foriequal0/homework-18-graphics-raytracer's render path:
// AoS scene, face_normal() recomputed per ray-triangle pair (main.rs:184, 202), true recursion,
// brute-force World::cast.  Every function cites the reference lines it follows.  Arithmetic is
// IEEE f32 with no contraction (build: -ffp-contract=off -fno-fast-math), matching rustc.
//
// Third-party semantics restated (crate sources are not vendored in the reference; versions from
// Cargo.toml:6-14, no Cargo.lock => patch level unpinned):
//   cgmath 0.16.1  dot=(ax*bx+ay*by)+az*bz; normalize=v*(1/|v|); cross; Vector3::angle=atan2(|axb|,a.b);
//                  Quaternion::from_arc / quat*vec; Deg->Rad
//   palette 0.4    LinSrgb +,*,/ component-wise; Mix::mix; into_luma (XYZ Y row); sRGB OETF + u8
//   rand 0.5       NOT reproduced: IsaacRng + ziggurat Normal are replaced by a Philox4x32-10 counter
//                  stream keyed (seed; x, y, epoch) with Box-Muller normals.  The CUDA path uses the
//                  same stream, so stochastic samples can be compared one by one.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle.h"

namespace {

// ------------------------------------------------------------------------------------------------
// cgmath / palette restatement
// ------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
struct V2 { float x, y; };
struct Rgb { float r, g, b; };

inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
inline V3 mk(const float* p) { return V3{p[0], p[1], p[2]}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float magnitude(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a * (1.0f / magnitude(a)); }
inline float distance(V3 a, V3 b) { return magnitude(b - a); }  // MetricSpace for Point3
inline V2 operator+(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
inline V2 operator*(V2 a, float s) { return V2{a.x * s, a.y * s}; }

inline Rgb rgb(float r, float g, float b) { return Rgb{r, g, b}; }
inline Rgb rgb(const float* p) { return Rgb{p[0], p[1], p[2]}; }
inline Rgb black() { return Rgb{0.0f, 0.0f, 0.0f}; }
inline Rgb operator+(Rgb a, Rgb b) { return Rgb{a.r + b.r, a.g + b.g, a.b + b.b}; }
inline Rgb operator-(Rgb a, Rgb b) { return Rgb{a.r - b.r, a.g - b.g, a.b - b.b}; }
inline Rgb operator*(Rgb a, Rgb b) { return Rgb{a.r * b.r, a.g * b.g, a.b * b.b}; }
inline Rgb operator*(Rgb a, float s) { return Rgb{a.r * s, a.g * s, a.b * s}; }
inline Rgb operator/(Rgb a, float s) { return Rgb{a.r / s, a.g / s, a.b / s}; }
// palette Mix::mix: self + (other - self) * clamp(factor, 0, 1)
inline Rgb mix(Rgb a, Rgb b, float f) {
    f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
    return a + (b - a) * f;
}

const float F32_EPSILON = 1.1920929e-7f;  // std::f32::EPSILON
const float PI_F = 3.14159265358979323846f;  // std::f32::consts::PI

// approx 0.1 ulps_eq!(a, b) with the defaults cgmath uses: epsilon = f32::EPSILON, max_ulps = 4
inline float signum(float v) { return std::isnan(v) ? v : (std::signbit(v) ? -1.0f : 1.0f); }
inline bool ulps_eq(float a, float b) {
    if (std::fabs(a - b) <= F32_EPSILON) return true;
    if (signum(a) != signum(b)) return false;
    int32_t ia, ib;
    std::memcpy(&ia, &a, 4);
    std::memcpy(&ib, &b, 4);
    int64_t diff = (int64_t)ia - (int64_t)ib;
    if (diff < 0) diff = -diff;
    return diff <= 4;
}

struct Quat { float s; V3 v; };
// cgmath Quaternion::from_arc(src, dst, None)
Quat from_arc(V3 src, V3 dst) {
    float mag_avg = std::sqrt(dot(src, src) * dot(dst, dst));
    float d = dot(src, dst);
    if (ulps_eq(d, mag_avg)) {
        return Quat{1.0f, mk(0.0f, 0.0f, 0.0f)};
    } else if (ulps_eq(d, -mag_avg)) {
        V3 v = cross(mk(1.0f, 0.0f, 0.0f), src);
        if (ulps_eq(v.x, 0.0f) && ulps_eq(v.y, 0.0f) && ulps_eq(v.z, 0.0f)) v = cross(mk(0.0f, 1.0f, 0.0f), src);
        V3 axis = normalize(v);
        // from_axis_angle(axis, Rad::turn_div_2()): (s, c) = sin_cos(angle * 0.5)
        float half = PI_F * 0.5f;
        float s = std::sin(half), c = std::cos(half);
        return Quat{c, axis * s};
    } else {
        Quat q{mag_avg + d, cross(src, dst)};
        float mag = std::sqrt(q.s * q.s + dot(q.v, q.v));
        float inv = 1.0f / mag;
        return Quat{q.s * inv, q.v * inv};
    }
}
// cgmath Quaternion * Vector3
inline V3 rotate(Quat q, V3 vec) {
    V3 tmp = cross(q.v, vec) + (vec * q.s);
    return (cross(q.v, tmp) * 2.0f) + vec;
}

// Rust `f as i32`: saturating, NaN -> 0
inline int32_t f32_as_i32(float f) {
    if (std::isnan(f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

// f32::is_normal
inline bool is_normal(float f) { return std::fpclassify(f) == FP_NORMAL; }

// ------------------------------------------------------------------------------------------------
// Sample stream: Philox4x32-10, counter = (x, y, epoch, block), key = (seed_lo, seed_hi)
// ------------------------------------------------------------------------------------------------
inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    std::memcpy(out, c, sizeof c);
}

struct Rng {
    uint32_t key[2];
    uint32_t x, y, epoch;
    uint32_t draws;
    uint32_t block[4];
    Rng(uint64_t seed, uint32_t y_, uint32_t x_, uint32_t epoch_) : x(x_), y(y_), epoch(epoch_), draws(0) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
    }
    uint32_t next_u32() {
        if ((draws & 3u) == 0u) {
            uint32_t ctr[4] = {x, y, epoch, draws >> 2};
            philox4x32_10(ctr, key, block);
        }
        return block[draws++ & 3u];
    }
    // uniform in [0,1): top 24 bits
    float uniform() { return (float)(next_u32() >> 8) * 5.9604644775390625e-8f; }
    // stands in for rand 0.5 gen_range::<f32>(low, high)
    float gen_range(float low, float high) { return low + (high - low) * uniform(); }
};

// ------------------------------------------------------------------------------------------------
// reference types
// ------------------------------------------------------------------------------------------------
enum Face : uint32_t { Front = 0, Back = 1, Both = 2 };
inline Face invert(Face f) { return f == Front ? Back : (f == Back ? Front : Both); }  // main.rs:59-66

struct Ray {  // main.rs:69-81
    V3 origin, direction;
    Face face_direction;
    bool has_exclude;
    int32_t ex_index;
    Face ex_face;
};

struct At { V3 position, normal; V2 uv; };  // PositionNormalUV

struct Hit {  // main.rs:139-147
    uint32_t object;
    Ray ray;
    int32_t index;
    At at;
    Face face_direction;
    float distance;
};

struct Counters { uint64_t casts = 0, tri = 0, sph = 0, samples = 0; };

struct World {
    const b200rt_scene* s;
    Counters* cnt;
};

inline V3 vpos(const b200rt_triangle& t, int i) { return mk(t.vertices[i].position); }

// primitives.rs:37-42
inline V3 face_normal(const b200rt_triangle& t) {
    V3 a = vpos(t, 1) - vpos(t, 0);
    V3 b = vpos(t, 2) - vpos(t, 1);
    return normalize(cross(a, b));
}
// primitives.rs:44-46
inline bool backface(const b200rt_triangle& t, V3 ray_dir) { return dot(face_normal(t), ray_dir) > 0.0f; }

// main.rs:180-326
bool cast(const World& w, const Ray& ray, Hit* out) {
    const b200rt_scene& sc = *w.s;
    if (w.cnt) {
        w.cnt->casts += 1;
        w.cnt->tri += sc.n_triangles;
        w.cnt->sph += sc.n_spheres;
    }
    bool have_nearest = false;
    float nearest_distance = 0.0f;
    Hit nearest{};
    for (uint32_t i = 0; i < sc.n_triangles; ++i) {
        const b200rt_triangle& triangle = sc.triangles[i];
        bool bf = backface(triangle, ray.direction);                                   // :184
        if ((bf && ray.face_direction == Front) || (!bf && ray.face_direction == Back)) continue;  // :185-188
        if (ray.has_exclude) {                                                          // :190-200
            bool same_face = ray.ex_index == (int32_t)i;
            bool criteria = ray.ex_face == Front ? !bf : (ray.ex_face == Back ? bf : true);
            if (same_face && criteria) continue;
        }
        V3 fn = face_normal(triangle);                                                  // :202
        float d = dot(fn, vpos(triangle, 0));                                           // :203
        float travel_distance = (d - dot(fn, ray.origin)) / dot(fn, ray.direction);     // :204
        if (travel_distance <= 0.0f) continue;                                          // :205
        V3 position = ray.origin + ray.direction * travel_distance;                     // :210
        V3 v[3] = {vpos(triangle, 0), vpos(triangle, 1), vpos(triangle, 2)};
        float area[3] = {
            dot(cross(v[2] - v[1], position - v[1]), fn),                               // :219
            dot(cross(v[0] - v[2], position - v[2]), fn),                               // :220
            dot(cross(v[1] - v[0], position - v[0]), fn),                               // :221
        };
        if (area[0] < 0.0f || area[1] < 0.0f || area[2] < 0.0f) continue;                // :224
        if (have_nearest && nearest_distance < travel_distance) continue;               // :229-233
        float area_of_triangle = dot(cross(v[1] - v[0], v[2] - v[0]), fn);              // :235
        V3 barycentric = mk(area[0], area[1], area[2]) / area_of_triangle;              // :236
        V3 n0 = mk(triangle.vertices[0].normal), n1 = mk(triangle.vertices[1].normal),
           n2 = mk(triangle.vertices[2].normal);
        V3 tmp = (n0 * barycentric.x + n1 * barycentric.y) + n2 * barycentric.z;        // :249 Matrix3*Vector3
        V3 normal = bf ? -tmp : tmp;                                                    // :250
        V2 uv0{triangle.vertices[0].uv[0], triangle.vertices[0].uv[1]};
        V2 uv1{triangle.vertices[1].uv[0], triangle.vertices[1].uv[1]};
        V2 uv2{triangle.vertices[2].uv[0], triangle.vertices[2].uv[1]};
        V2 uv = uv0 * barycentric.x + uv1 * barycentric.y + uv2 * barycentric.z;        // :252
        have_nearest = true;
        nearest_distance = travel_distance;                                             // :253
        nearest.object = triangle.object_index;
        nearest.ray = ray;
        nearest.index = (int32_t)i;
        nearest.at = At{position, normal, uv};
        nearest.distance = travel_distance;
        nearest.face_direction = bf ? Back : Front;                                     // :260
    }
    for (uint32_t i = 0; i < sc.n_spheres; ++i) {
        const b200rt_sphere& sphere = sc.spheres[i];
        V3 center = mk(sphere.center);
        float line_sphere_distance = magnitude(cross(center - ray.origin, ray.direction));  // :265
        if (line_sphere_distance > sphere.radius) continue;                             // :266
        V3 displacement = center - ray.origin;                                          // :270
        float tc = dot(ray.direction, displacement);                                    // :271
        float k = std::sqrt(sphere.radius * sphere.radius - line_sphere_distance * line_sphere_distance);  // :272
        float travel_distance;
        bool bf;
        if (ray.face_direction == Front) { travel_distance = tc - k; bf = false; }      // :274
        else if (ray.face_direction == Back) { travel_distance = tc + k; bf = true; }   // :275
        else if (tc < k) { travel_distance = tc + k; bf = true; }                       // :276-277
        else { travel_distance = tc - k; bf = false; }                                  // :279
        if (travel_distance <= 0.0f) continue;                                          // :282
        const int32_t prim = (int32_t)(sc.n_triangles + i);
        if (ray.has_exclude) {                                                          // :286-296
            bool same_face = ray.ex_index == prim;
            bool criteria = ray.ex_face == Front ? !bf : (ray.ex_face == Back ? bf : true);
            if (same_face && criteria) continue;
        }
        if (have_nearest && nearest_distance < travel_distance) continue;               // :298-302
        V3 position = ray.origin + ray.direction * travel_distance;                     // :304
        V3 tmp = normalize(position - center);                                          // :306
        V3 normal = bf ? -tmp : tmp;
        V2 uv{std::acos(normal.y) / PI_F,                                               // :311
              std::atan2(normal.z, normal.x) / (PI_F * 2.0f) + 0.5f};                   // :312
        have_nearest = true;
        nearest_distance = travel_distance;
        nearest.object = sphere.object_index;
        nearest.ray = ray;
        nearest.index = prim;
        nearest.at = At{position, normal, uv};
        nearest.distance = travel_distance;
        nearest.face_direction = bf ? Back : Front;
    }
    if (have_nearest && out) *out = nearest;
    return have_nearest;
}

// materials.rs:33-37 (ColorMaterial) and 85-103 (GenerativeMaterial); closures main.rs:848-863, 1019-1026
b200rt_material material_approx(const b200rt_material& m, V2 uv) {
    if (m.kind != B200RT_MATERIAL_GENERATIVE) return m;
    b200rt_material o = m;
    o.kind = B200RT_MATERIAL_COLOR;
    const float* p = m.fn_params;
    switch (m.diffuse_fn) {
        case B200RT_DIFFUSE_STRIPE_V: {
            bool even = f32_as_i32(uv.y * p[0]) % 2 == 0;
            const float* c = even ? p + 1 : p + 4;
            o.diffuse_color[0] = c[0]; o.diffuse_color[1] = c[1]; o.diffuse_color[2] = c[2];
            break;
        }
        case B200RT_DIFFUSE_CHECKER_UPV: {
            bool even = f32_as_i32((uv.x + uv.y) * p[0]) % 2 == 0;
            const float* c = even ? p + 1 : p + 4;
            o.diffuse_color[0] = c[0]; o.diffuse_color[1] = c[1]; o.diffuse_color[2] = c[2];
            break;
        }
        default: break;
    }
    switch (m.normal_fn) {
        case B200RT_NORMAL_SINCOS_U: {
            float angle = uv.x * p[7] * 2.0f * PI_F;                                    // main.rs:856
            V3 v = mk(std::sin(angle), 0.0f, std::cos(angle));
            if (dot(v, mk(0.0f, 0.0f, 1.0f)) <= 0.0f) v = -v;                           // main.rs:858-862
            o.normal[0] = v.x; o.normal[1] = v.y; o.normal[2] = v.z;
            break;
        }
        default: break;
    }
    return o;
}

// materials.rs:40-44
inline V3 adjust_normal(const b200rt_material& m, V3 normal) {
    return rotate(from_arc(mk(0.0f, 0.0f, 1.0f), normal), mk(m.normal));
}
// materials.rs:46-53   (probe.at.normal = n, probe.light_direction = l)
inline Rgb get_diffuse(const b200rt_material& m, V3 n, V3 l) {
    float cosine = dot(l, n);
    if (cosine > 0.0f) return rgb(m.diffuse_color) * cosine;
    return black();
}
// materials.rs:55-66
inline Rgb get_specular(const b200rt_material& m, V3 n, V3 view, V3 l) {
    float cosine = dot(l, n);
    if (cosine <= 0.0f) return black();
    V3 reflected_ray = 2.0f * cosine * n - l;
    float specular = 1.0f / (m.smoothness + F32_EPSILON);
    float energy_conserving = (specular + 8.0f) / (8.0f * PI_F);
    float specular_amount = std::pow(std::fmax(dot(reflected_ray, view), 0.0f), specular) * energy_conserving;
    return rgb(m.specular_color) * specular_amount;
}

struct Directional { bool has_origin; V3 origin, direction; Rgb color; };

// lights.rs:48-93
bool approximate_into_directional(const b200rt_light& L, V3 position, Directional* out) {
    switch (L.kind) {
        case B200RT_LIGHT_DIRECTIONAL:  // lights.rs:48-52
            out->has_origin = L.has_origin != 0;
            out->origin = mk(L.origin);
            out->direction = mk(L.direction);
            out->color = rgb(L.color);
            return true;
        case B200RT_LIGHT_SPOT: {  // lights.rs:54-72
            V3 origin = mk(L.origin);
            V3 offset = position - origin;
            V3 sd = mk(L.direction);
            float angle = std::fabs(std::atan2(magnitude(cross(sd, offset)), dot(sd, offset)));
            float spot_spread = L.angle;
            if (angle > spot_spread) return false;
            float angular_attenuation = std::pow(1.0f - angle / spot_spread, L.softness + F32_EPSILON);
            float distance_attenuation = 1.0f / (magnitude(offset) + F32_EPSILON);
            out->has_origin = true;
            out->origin = origin;
            out->direction = normalize(position - origin);
            out->color = rgb(L.color) * angular_attenuation * distance_attenuation;
            return true;
        }
        case B200RT_LIGHT_POINT: {  // lights.rs:74-84
            V3 origin = mk(L.origin);
            V3 offset = position - origin;
            float distance_attenuation = 1.0f / (magnitude(offset) + F32_EPSILON);
            out->has_origin = true;
            out->origin = origin;
            out->direction = normalize(offset);
            out->color = rgb(L.color) * distance_attenuation;
            return true;
        }
        default: return false;
    }
}

// main.rs:328-341
Ray get_reflect(const Hit& hit) {
    V3 n = hit.at.normal, l = hit.ray.direction;
    V3 reflected = l - 2.0f * dot(l, n) * n;
    Ray r;
    r.origin = hit.at.position;
    r.direction = normalize(reflected);
    r.face_direction = hit.ray.face_direction;
    r.has_exclude = true;
    r.ex_index = hit.index;
    r.ex_face = invert(hit.face_direction);
    return r;
}

// closure at main.rs:344-352
bool refract(V3 n, V3 l, float k, V3* out) {
    float cos = -dot(l, n);
    if (k * k >= 1.0f - cos * cos) {
        V3 x = (l + n * cos) / k - n * std::sqrt(1.0f - (1.0f - cos * cos) / (k * k));
        *out = normalize(x);
        return true;
    }
    return false;
}

enum RefractionKind { Escaped, Infinite, Trapped };
struct Refraction { RefractionKind kind; float travel_distance; Ray escape_ray; };

// main.rs:343-405
Refraction get_refract(const World& w, const Hit& hit, float max_distance, uint32_t tir_retries) {
    Refraction res{};
    float k = material_approx(w.s->materials[hit.object], hit.at.uv).refraction_index;   // :354
    V3 refract_in;
    if (!refract(hit.at.normal, hit.ray.direction, k, &refract_in)) { res.kind = Trapped; return res; }
    Ray ray_inside;
    ray_inside.origin = hit.at.position;
    ray_inside.direction = normalize(refract_in);                                        // :362
    ray_inside.face_direction = Back;
    ray_inside.has_exclude = true;
    ray_inside.ex_index = hit.index;
    ray_inside.ex_face = Front;
    Hit hit_inside;
    if (!cast(w, ray_inside, &hit_inside)) { res.kind = Infinite; return res; }           // :371-374
    float travel_distance = distance(hit_inside.at.position, hit.at.position);           // :375
    V3 refract_out;
    bool have_out = refract(hit_inside.at.normal, hit_inside.ray.direction, 1.0f / k, &refract_out);  // :376
    uint32_t retry = 0;
    while (!have_out && travel_distance <= max_distance && retry < tir_retries) {         // :378
        V3 previous_hit_position = hit_inside.at.position;
        Ray total_reflect = get_reflect(hit_inside);
        if (!cast(w, total_reflect, &hit_inside)) { res.kind = Infinite; return res; }    // :381-384
        travel_distance += distance(previous_hit_position, hit_inside.at.position);      // :385
        have_out = refract(hit_inside.at.normal, hit_inside.ray.direction, 1.0f / k, &refract_out);
        retry += 1;
    }
    if (!have_out) { res.kind = Trapped; return res; }
    res.kind = Escaped;
    res.travel_distance = travel_distance;
    res.escape_ray.origin = hit_inside.at.position;
    res.escape_ray.direction = normalize(refract_out);                                   // :395
    res.escape_ray.face_direction = Front;
    res.escape_ray.has_exclude = true;
    res.escape_ray.ex_index = hit_inside.index;
    res.escape_ray.ex_face = Back;
    return res;
}

// main.rs:407-464
Rgb get_shade(const World& w, const Hit& hit) {
    b200rt_material material = material_approx(w.s->materials[hit.object], hit.at.uv);   // :408
    const Ray& ray = hit.ray;
    V3 normal = adjust_normal(material, hit.at.normal);                                  // :410
    Rgb sum = black();
    for (uint32_t li = 0; li < w.s->n_lights; ++li) {
        Directional light;
        if (!approximate_into_directional(w.s->lights[li], hit.at.position, &light)) continue;  // :414-417
        float cosine = -dot(light.direction, normal);                                    // :420
        if (cosine <= 0.0f) continue;
        Ray shadow_ray;                                                                  // :425-433
        shadow_ray.origin = hit.at.position;
        shadow_ray.direction = -light.direction;
        shadow_ray.face_direction = Back;
        shadow_ray.has_exclude = true;
        shadow_ray.ex_index = hit.index;
        shadow_ray.ex_face = Back;
        Hit occlusion;
        if (cast(w, shadow_ray, &occlusion)) {                                           // :435-448
            if (light.has_origin) {
                float occlusion_distance = distance(hit.at.position, occlusion.at.position);
                float light_distance = distance(hit.at.position, light.origin);
                if (occlusion_distance < light_distance) continue;
            } else {
                continue;
            }
        }
        V3 view_direction = -ray.direction, light_direction = -light.direction;          // :450-454
        float shiness = material.shiness;
        Rgb diffuse = get_diffuse(material, normal, light_direction) * light.color;      // :458
        Rgb specular = get_specular(material, normal, view_direction, light_direction) * light.color;  // :459
        sum = sum + diffuse * (1.0f - shiness) + specular * shiness;                     // :461
    }
    return sum;
}

struct TraceState { int32_t depth; float contribution; };  // main.rs:668-680
inline TraceState nested(const TraceState& s, float decay) { return TraceState{s.depth - 1, s.contribution * decay}; }

struct Cfg { float threshold, refract_max_distance; uint32_t tir_retries; };

// main.rs:466-519
Rgb ray_trace(const World& w, const Cfg& cfg, const TraceState& state, const Ray& ray, int32_t* primary_id) {
    const float THRESHOLD = cfg.threshold;
    if (state.contribution < THRESHOLD) return black();                                  // :469
    Hit hit;
    if (!cast(w, ray, &hit)) return black();                                             // :473-476
    if (primary_id) *primary_id = hit.index;
    b200rt_material material = material_approx(w.s->materials[hit.object], hit.at.uv);   // :478
    float shade_contribution = (1.0f - material.shiness) * (1.0f - material.transparency);  // :480
    TraceState shade_state = nested(state, shade_contribution);
    Rgb shade = shade_state.contribution >= THRESHOLD ? get_shade(w, hit) : black();      // :482-486
    if (state.depth <= 0) return shade;                                                  // :488-490
    float reflection_contribution = material.shiness * (1.0f - material.transparency);   // :493
    TraceState reflection_state = nested(state, reflection_contribution);
    Rgb reflection = black();
    if (reflection_state.contribution >= THRESHOLD) {                                    // :495
        Ray reflected_ray = get_reflect(hit);
        reflection = ray_trace(w, cfg, reflection_state, reflected_ray, nullptr);
    }
    float refraction_contribution = material.transparency;                               // :502
    TraceState refraction_state = nested(state, refraction_contribution);
    Rgb refraction = black();
    if (refraction_state.contribution > THRESHOLD) {                                     // :504
        Refraction r = get_refract(w, hit, cfg.refract_max_distance, cfg.tir_retries);   // :505
        if (r.kind == Escaped) {
            Rgb s = ray_trace(w, cfg, refraction_state, r.escape_ray, nullptr);          // :507
            refraction = s * std::pow(material.opaque_decay, r.travel_distance);         // :508
        }
    }
    return shade * shade_contribution + reflection * reflection_contribution +           // :516-518
           refraction * refraction_contribution;
}

struct DistributeState { int32_t depth; float contribution; Rng* rng; };  // main.rs:682-698
inline DistributeState nested(DistributeState& s, float decay) {
    return DistributeState{s.depth - 1, s.contribution * decay, s.rng};
}

enum RayType { Diffuse, Reflection, RefractionT };

// main.rs:652-666
RayType weighted_select(Rng& rng, const float w[3]) {
    float sum = (w[0] + w[1]) + w[2];
    float r = rng.gen_range(0.0f, sum);
    float accum = 0.0f;
    for (int i = 0; i < 3; ++i) {
        accum += w[i];
        if (r < accum) return (RayType)i;
    }
    return RefractionT;
}

// main.rs:539-554
Hit scatter_hit(DistributeState& state, const Hit& hit, V3 direction, float exponent) {
    float phi = std::acos(std::pow(1.0f - state.rng->gen_range(0.0f, 1.0f), exponent));  // :543
    float theta = state.rng->gen_range(-PI_F, PI_F);                                     // :544
    V3 z = mk(0.0f, 0.0f, 1.0f);
    Quat from_z = from_arc(z, normalize(direction));                                     // :546
    V3 new_dir = rotate(from_z, mk(std::sin(phi) * std::cos(theta), std::sin(phi) * std::sin(theta), std::cos(phi)));
    Hit out = hit;
    out.ray.direction = new_dir;                                                         // :552
    return out;
}

// main.rs:521-614
Rgb distributed_ray_trace(const World& w, const Cfg& cfg, DistributeState& state, const Hit& hit) {
    Rgb shade = get_shade(w, hit);                                                       // :524
    if (state.depth <= 0) return shade;
    b200rt_material material = material_approx(w.s->materials[hit.object], hit.at.uv);   // :529
    const float weights[3] = {(1.0f - material.shiness) * (1.0f - material.transparency),  // :534
                              material.shiness * (1.0f - material.transparency),        // :535
                              material.transparency};                                    // :536
    RayType selected = weighted_select(*state.rng, weights);
    switch (selected) {
        case Diffuse: {                                                                  // :557-575
            Hit scattered_hit = scatter_hit(state, hit, -hit.at.normal, 1.0f);
            float cosine = -dot(hit.at.normal, scattered_hit.ray.direction);
            if (cosine <= 0.0f) return black();
            Ray reflected = get_reflect(scattered_hit);
            Hit reflected_hit;
            if (cast(w, reflected, &reflected_hit)) {
                DistributeState ns = nested(state, 1.0f);
                Rgb x = distributed_ray_trace(w, cfg, ns, reflected_hit);
                Rgb s = x * get_diffuse(material, scattered_hit.at.normal, reflected.direction);  // :566-570
                return mix(get_shade(w, reflected_hit), s, 0.5f);                         // :571
            }
            return get_shade(w, scattered_hit);                                          // :573
        }
        case Reflection: {                                                               // :576-594
            Hit scattered_hit = scatter_hit(state, hit, hit.ray.direction, material.smoothness);
            float cosine = -dot(hit.at.normal, scattered_hit.ray.direction);
            if (cosine <= 0.0f) return black();
            Ray reflected = get_reflect(scattered_hit);
            Hit reflected_hit;
            if (cast(w, reflected, &reflected_hit)) {
                DistributeState ns = nested(state, 1.0f);
                Rgb x = distributed_ray_trace(w, cfg, ns, reflected_hit);
                Rgb s = x * get_specular(material, scattered_hit.at.normal, -hit.ray.direction, reflected.direction);  // :585-589
                return mix(get_shade(w, reflected_hit), s, 0.5f);                         // :590
            }
            return get_shade(w, scattered_hit);                                          // :592
        }
        default: {                                                                       // :595-612
            Hit scattered_hit = scatter_hit(state, hit, hit.ray.direction, material.smoothness);
            float cosine = -dot(hit.at.normal, scattered_hit.ray.direction);
            if (cosine <= 0.0f) return black();
            Refraction r = get_refract(w, scattered_hit, cfg.refract_max_distance, cfg.tir_retries);
            if (r.kind == Escaped) {
                Hit refracted_hit;
                if (cast(w, r.escape_ray, &refracted_hit)) {
                    DistributeState ns = nested(state, 1.0f);
                    Rgb x = distributed_ray_trace(w, cfg, ns, refracted_hit);
                    return (x + get_shade(w, refracted_hit)) * std::pow(material.opaque_decay, r.travel_distance);  // :605
                }
                return black();
            }
            return black();
        }
    }
}

// main.rs:83-99
struct CameraBasis { V3 toward, x, y, origin; };
CameraBasis camera_basis(const b200rt_camera& cam) {
    CameraBasis b;
    V3 toward = normalize(mk(cam.toward));                                               // :85
    V3 right = normalize(cross(toward, mk(cam.up)));                                     // :86
    V3 up = normalize(cross(right, toward));                                             // :87
    float t = std::tan(cam.fovy / 2.0f);
    b.toward = toward;
    b.x = t * right;                                                                     // :89
    b.y = t * up;                                                                        // :90
    b.origin = mk(cam.center) + toward * cam.near;                                       // :92
    return b;
}
Ray shoot(const b200rt_camera& cam, float clip_x, float clip_y) {
    CameraBasis b = camera_basis(cam);
    Ray r;
    r.direction = normalize(clip_x * b.x + clip_y * b.y + b.toward);                     // :91
    r.origin = b.origin;
    r.has_exclude = false;
    r.ex_index = -1;
    r.ex_face = Front;
    r.face_direction = Front;
    return r;
}
// main.rs:101-127.  Normal::new(0, blur) is sampled by Box-Muller on two stream uniforms
// (x offset from the cosine branch, y offset from the sine branch) instead of rand's ziggurat.
Ray shoot_focus(const b200rt_camera& cam, float clip_x, float clip_y, Rng& rng, float focus, float blur) {
    CameraBasis b = camera_basis(cam);
    V3 direction = normalize(clip_x * b.x + clip_y * b.y + b.toward);                    // :110
    float u1 = 1.0f - rng.uniform();
    float u2 = rng.uniform();
    float radius = std::sqrt(-2.0f * std::log(u1));
    float ang = 2.0f * PI_F * u2;
    float xoffset = blur * (radius * std::cos(ang));                                     // :112
    float yoffset = blur * (radius * std::sin(ang));                                     // :113
    V3 direction_offset = normalize(direction * focus + b.x * xoffset + b.y * yoffset);  // :115-117
    V3 origin = mk(cam.center) + normalize(b.toward) * cam.near - (b.x * xoffset + b.y * yoffset);  // :118-120
    Ray r;
    r.origin = origin;
    r.direction = direction_offset;
    r.has_exclude = false;
    r.ex_index = -1;
    r.ex_face = Front;
    r.face_direction = Front;
    return r;
}

inline void clip_of(uint32_t x, uint32_t y, uint32_t width, uint32_t height, float* cx, float* cy) {
    *cy = ((float)height / 2.0f - (float)y) / (float)height;                             // main.rs:1094
    *cx = ((float)x - (float)width / 2.0f) / (float)height;                              // main.rs:1095
}

inline void rows_of(const b200rt_params& p, uint32_t* r0, uint32_t* r1) {
    if (p.row_count == 0) { *r0 = 0; *r1 = p.height; }
    else { *r0 = p.row_begin; *r1 = std::min(p.height, p.row_begin + p.row_count); }
}

Rgb sample_distributed(const World& w, const b200rt_camera& cam, const b200rt_params& p, const Cfg& cfg,
                       uint32_t y, uint32_t x, uint32_t epoch) {
    float cx, cy;
    clip_of(x, y, p.width, p.height, &cx, &cy);
    Rng rng(p.seed, y, x, epoch);
    DistributeState state{p.depth, 1.0f, &rng};                                           // main.rs:1137-1142
    Ray ray = shoot_focus(cam, cx, cy, rng, p.focus, p.blur);                             // main.rs:1144-1149
    Hit hit;
    if (cast(w, ray, &hit)) return distributed_ray_trace(w, cfg, state, hit);            // main.rs:1150-1152
    return black();                                                                      // main.rs:1154
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C API
// ------------------------------------------------------------------------------------------------
extern "C" {

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// main.rs:1089-1109
int oracle_render_whitted(const b200rt_scene* scene, const b200rt_camera* cam, const b200rt_params* params,
                          float* out_rgb, int32_t* out_prim_id, uint64_t counters[4], int n_threads) {
    if (!scene || !cam || !params || !out_rgb) return B200RT_ERR_INVALID;
    const b200rt_params p = *params;
    uint32_t r0, r1;
    rows_of(p, &r0, &r1);
    const Cfg cfg{p.threshold, p.refract_max_distance, p.tir_retries};
    const int64_t n = (int64_t)(r1 - r0) * p.width;
    uint64_t c_casts = 0, c_tri = 0, c_sph = 0, c_samples = 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads) reduction(+ : c_casts, c_tri, c_sph, c_samples)
#endif
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t y = r0 + (uint32_t)(i / p.width), x = (uint32_t)(i % p.width);
        Counters cnt;
        World w{scene, &cnt};
        float cx, cy;
        clip_of(x, y, p.width, p.height, &cx, &cy);
        Ray ray = shoot(*cam, cx, cy);                                                   // main.rs:1096
        TraceState state{p.depth, 1.0f};                                                 // main.rs:1097-1100
        int32_t prim = -1;
        Rgb photon = ray_trace(w, cfg, state, ray, &prim);                               // main.rs:1101
        const size_t at = (size_t)y * p.width + x;
        out_rgb[3 * at + 0] = 0.0f + photon.r;                                           // main.rs:1107
        out_rgb[3 * at + 1] = 0.0f + photon.g;
        out_rgb[3 * at + 2] = 0.0f + photon.b;
        if (out_prim_id) out_prim_id[at] = prim;
        c_casts += cnt.casts; c_tri += cnt.tri; c_sph += cnt.sph; c_samples += 1;
    }
    if (counters) { counters[0] = c_casts; counters[1] = c_tri; counters[2] = c_sph; counters[3] = c_samples; }
    return B200RT_OK;
}

// main.rs:1129-1167 for epochs [epoch_begin, epoch_begin+epoch_count); accum += per photon.rs:29-32
int oracle_render_distributed(const b200rt_scene* scene, const b200rt_camera* cam, const b200rt_params* params,
                              uint32_t epoch_begin, uint32_t epoch_count, float* accum, uint64_t counters[4],
                              int n_threads) {
    if (!scene || !cam || !params || !accum) return B200RT_ERR_INVALID;
    const b200rt_params p = *params;
    uint32_t r0, r1;
    rows_of(p, &r0, &r1);
    const Cfg cfg{p.threshold, p.refract_max_distance, p.tir_retries};
    const int64_t n = (int64_t)(r1 - r0) * p.width;
    uint64_t c_casts = 0, c_tri = 0, c_sph = 0, c_samples = 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads) reduction(+ : c_casts, c_tri, c_sph, c_samples)
#endif
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t y = r0 + (uint32_t)(i / p.width), x = (uint32_t)(i % p.width);
        Counters cnt;
        World w{scene, &cnt};
        const size_t at = (size_t)y * p.width + x;
        float* acc = accum + 4 * at;
        for (uint32_t e = epoch_begin; e < epoch_begin + epoch_count; ++e) {
            Rgb photon = sample_distributed(w, *cam, p, cfg, y, x, e);
            if (is_normal(photon.r) && is_normal(photon.g) && is_normal(photon.b)) {     // main.rs:1157-1160
                acc[0] += photon.r; acc[1] += photon.g; acc[2] += photon.b;              // photon.rs:30
                acc[3] += 1.0f;                                                          // photon.rs:31
                c_samples += 1;
            }
        }
        c_casts += cnt.casts; c_tri += cnt.tri; c_sph += cnt.sph;
    }
    if (counters) { counters[0] = c_casts; counters[1] = c_tri; counters[2] = c_sph; counters[3] = c_samples; }
    return B200RT_OK;
}

int oracle_sample_distributed(const b200rt_scene* scene, const b200rt_camera* cam, const b200rt_params* params,
                              uint32_t y, uint32_t x, uint32_t epoch, float out[3]) {
    if (!scene || !cam || !params || !out) return B200RT_ERR_INVALID;
    World w{scene, nullptr};
    const Cfg cfg{params->threshold, params->refract_max_distance, params->tir_retries};
    Rgb c = sample_distributed(w, *cam, *params, cfg, y, x, epoch);
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
    return B200RT_OK;
}

int oracle_intersect(const b200rt_scene* scene, const b200rt_ray* rays, size_t n, b200rt_hit* hits) {
    if (!scene || (!rays && n) || (!hits && n)) return B200RT_ERR_INVALID;
    World w{scene, nullptr};
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        Ray r;
        r.origin = mk(rays[i].origin);
        r.direction = mk(rays[i].direction);
        r.face_direction = (Face)rays[i].face_direction;
        r.has_exclude = rays[i].exclude_prim >= 0;
        r.ex_index = rays[i].exclude_prim;
        r.ex_face = (Face)rays[i].exclude_face;
        Hit h;
        b200rt_hit o;
        std::memset(&o, 0, sizeof o);
        if (cast(w, r, &h)) {
            o.prim_id = h.index;
            o.object_index = h.object;
            o.face_direction = h.face_direction;
            o.distance = h.distance;
            o.position[0] = h.at.position.x; o.position[1] = h.at.position.y; o.position[2] = h.at.position.z;
            o.normal[0] = h.at.normal.x; o.normal[1] = h.at.normal.y; o.normal[2] = h.at.normal.z;
            o.uv[0] = h.at.uv.x; o.uv[1] = h.at.uv.y;
        } else {
            o.prim_id = -1;
        }
        hits[i] = o;
    }
    return B200RT_OK;
}

// main.rs:748-762.  palette 0.4 into_luma: Y row of the linear sRGB -> XYZ (D65) matrix.
float oracle_post_process(float* rgbv, size_t n_pixels) {
    std::vector<float> luma;
    luma.reserve(n_pixels);
    for (size_t i = 0; i < n_pixels; ++i) {
        float l = rgbv[3 * i] * 0.2126729f + rgbv[3 * i + 1] * 0.7151522f + rgbv[3 * i + 2] * 0.0721750f;
        if (is_normal(l)) luma.push_back(l);
    }
    if (luma.empty()) return 0.0f;  // the reference would panic on the index (main.rs:754)
    std::sort(luma.begin(), luma.end());
    size_t idx = (size_t)((float)luma.size() * 0.99f);
    if (idx >= luma.size()) idx = luma.size() - 1;
    float p98 = luma[idx];
    if (p98 > F32_EPSILON) {
        for (size_t i = 0; i < 3 * n_pixels; ++i) rgbv[i] = rgbv[i] / p98;
        return p98;
    }
    return 0.0f;
}

// image.rs:55-66 -> palette Srgb::from_linear + into_format::<u8>()
void oracle_encode_srgb8(const float* v, size_t n_values, uint8_t* out) {
    for (size_t i = 0; i < n_values; ++i) {
        float x = v[i];
        float e = x <= 0.0031308f ? 12.92f * x : 1.055f * std::pow(x, 1.0f / 2.4f) - 0.055f;
        float s = e * 255.0f;
        s = s < 0.0f ? 0.0f : (s > 255.0f ? 255.0f : s);
        if (std::isnan(s)) s = 0.0f;
        out[i] = (uint8_t)std::round(s);
    }
}

// photon.rs:18-21
void oracle_resolve(const float* accum, size_t n_pixels, float* out_rgb) {
    for (size_t i = 0; i < n_pixels; ++i) {
        const float* a = accum + 4 * i;
        if (a[3] < F32_EPSILON) { out_rgb[3 * i] = out_rgb[3 * i + 1] = out_rgb[3 * i + 2] = 0.0f; }
        else { out_rgb[3 * i] = a[0] / a[3]; out_rgb[3 * i + 1] = a[1] / a[3]; out_rgb[3 * i + 2] = a[2] / a[3]; }
    }
}

void oracle_camera_shoot(const b200rt_camera* cam, float clip_x, float clip_y, b200rt_ray* out) {
    Ray r = shoot(*cam, clip_x, clip_y);
    out->origin[0] = r.origin.x; out->origin[1] = r.origin.y; out->origin[2] = r.origin.z;
    out->direction[0] = r.direction.x; out->direction[1] = r.direction.y; out->direction[2] = r.direction.z;
    out->face_direction = r.face_direction;
    out->exclude_prim = -1;
    out->exclude_face = 0;
}

int oracle_refract(const float n[3], const float l[3], float k, float out[3]) {
    V3 o;
    if (!refract(mk(n), mk(l), k, &o)) return 0;
    out[0] = o.x; out[1] = o.y; out[2] = o.z;
    return 1;
}

void oracle_from_arc_rotate(const float src[3], const float dst[3], const float v[3], float out[3]) {
    V3 o = rotate(from_arc(mk(src), mk(dst)), mk(v));
    out[0] = o.x; out[1] = o.y; out[2] = o.z;
}

int oracle_light_approx(const b200rt_light* light, const float position[3], float out_dir[3], float out_color[3],
                        float out_origin[3], int* out_has_origin) {
    Directional d;
    if (!approximate_into_directional(*light, mk(position), &d)) return 0;
    out_dir[0] = d.direction.x; out_dir[1] = d.direction.y; out_dir[2] = d.direction.z;
    out_color[0] = d.color.r; out_color[1] = d.color.g; out_color[2] = d.color.b;
    out_origin[0] = d.origin.x; out_origin[1] = d.origin.y; out_origin[2] = d.origin.z;
    *out_has_origin = d.has_origin ? 1 : 0;
    return 1;
}

void oracle_material_approx(const b200rt_material* m, const float uv[2], b200rt_material* out) {
    *out = material_approx(*m, V2{uv[0], uv[1]});
}
void oracle_get_diffuse(const b200rt_material* m, const float normal[3], const float light_dir[3], float out[3]) {
    Rgb c = get_diffuse(*m, mk(normal), mk(light_dir));
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}
void oracle_get_specular(const b200rt_material* m, const float normal[3], const float view_dir[3],
                         const float light_dir[3], float out[3]) {
    Rgb c = get_specular(*m, mk(normal), mk(view_dir), mk(light_dir));
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
void oracle_sample_uniforms(uint64_t seed, uint32_t y, uint32_t x, uint32_t epoch, int n, float* out) {
    Rng rng(seed, y, x, epoch);
    for (int i = 0; i < n; ++i) out[i] = rng.uniform();
}

}  // extern "C"
