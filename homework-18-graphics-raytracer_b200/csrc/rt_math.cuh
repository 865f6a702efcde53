// rt_math.cuh — device f32 math in the reference's evaluation order.
//
// This translation unit is compiled with -fmad=false: nvcc never contracts a*b+c, so every
// expression below rounds exactly like rustc's (and the oracle's) non-fused IEEE f32 code.
// Division and sqrt use the default -prec-div=true / -prec-sqrt=true (correctly rounded).
// The only fused arithmetic in the library is written explicitly with __fmaf_rn in the filter
// stage of the two-phase cast (rt_cast.cuh), whose results never reach an output.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace b200rt {

struct f3 { float x, y, z; };
struct f2 { float x, y; };

#define RT_DI __device__ __forceinline__
// heavy leaf math (IEEE div/sqrt expansions, libm calls): ONE copy in the kernel, keeps the phase machine inside
// the instruction cache (ncu: 70% of stalls were no_instruction when these were inlined at every site)
#define RT_DN static __device__ __noinline__

RT_DI f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_DI f3 mk3(const float* p) { return mk3(p[0], p[1], p[2]); }
RT_DI f3 mk3(float4 v) { return mk3(v.x, v.y, v.z); }
RT_DI f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DI f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DI f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
RT_DI f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RT_DI f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
RT_DI f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }  // LinSrgb * LinSrgb
RT_DI f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
// cgmath 0.16 Vector3::dot = mul_element_wise().sum() = (x + y) + z
RT_DI float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
RT_DI f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RT_DI float magnitude(f3 a) { return sqrtf(dot(a, a)); }
RT_DN f3 normalize(f3 a) { return a * (1.0f / magnitude(a)); }  // InnerSpace::normalize_to
RT_DI float distance(f3 a, f3 b) { return magnitude(b - a); }   // MetricSpace for Point3

// libm entry points: one out-of-line copy each (scalar register ABI, no stack traffic)
RT_DN float nl_powf(float a, float b) { return powf(a, b); }
// x^e where the result only enters a COLOUR (Phong lobe materials.rs:63, spot cone lights.rs:62-64, opaque decay
// main.rs:508 / 605 - never a ray or a decision): x^e = 2^(e log2 x) in 17 instructions instead of the ~100 of powf and
// its call.  x = m 2^k with m in [sqrt(1/2), sqrt(2)); f = m - 1 is exact; ln m = 2 atanh(s), s = f / (2 + f), as
// 2 s (1 + z/3 + z^2/5 + z^3/7 + z^4/9), z = s^2 <= 0.0295 (truncation 2e-9).  Measured on the device against f64 pow over
// x in (0, 1], e in [0.5, 8.4e6] (tests/test_gpu_image.py::test_color_pow_error_bound): absolute error <= 1.3e-7, relative
// error <= 2.4e-7 max(1, |e log2 x|) - powf itself is specified to 4 ulp, and the colours of the path are compared at 1e-4.  Outside 0 < e < inf, FLT_MIN <= x < inf the libm call decides.
RT_DI float color_pow(float x, float e) {
    if (!(e > 0.0f && e < CUDART_INF_F && x >= 1.17549435e-38f && x < CUDART_INF_F)) {
        if (x == 0.0f && e > 0.0f) return 0.0f;
        return nl_powf(x, e);
    }
    const int32_t ix = __float_as_int(x);
    const int32_t k = (ix - 0x3f3504f3) >> 23;
    const float f = __int_as_float(ix - (k << 23)) - 1.0f;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(2.0f + f));
    const float s = f * r, z = s * s;
    float p = __fmaf_rn(0.32059889f, z, 0.41219857f);      // (2 / ln 2) / 9, / 7
    p = __fmaf_rn(p, z, 0.57707802f);                      //             / 5
    p = __fmaf_rn(p, z, 0.96179669f);                      //             / 3
    p = __fmaf_rn(p, z, 2.88539008f);                      //  2 / ln 2
    const float y = e * ((float)k + s * p);
    // results below 2^-126 are formed 2^64 too large and scaled down (an exact multiply but for the subnormal's own rounding):
    // MUFU.EX2 flushes subnormal results to zero, and a flushed Phong term times its normalisation (up to 3e5) can be the only
    // contribution to a channel - the sample's is_normal filter (main.rs:1157-1160) would then see 0 where the reference sees a number
    const bool deep = y < -100.0f;
    float out;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(out) : "f"(deep ? y + 64.0f : y));
    return deep ? out * 5.42101086e-20f : out;
}
RT_DN float nl_acosf(float a) { return acosf(a); }
RT_DN float nl_atan2f(float a, float b) { return atan2f(a, b); }
RT_DN float nl_logf(float a) { return logf(a); }
#ifndef RT_SINCOS_JOINT
#define RT_SINCOS_JOINT 1   // sincosf: one argument reduction for both (the same polynomials as sinf / cosf)
#endif
RT_DN float2 nl_sincosf(float a) {
    float2 r;
#if RT_SINCOS_JOINT
    sincosf(a, &r.x, &r.y);
#else
    r.x = sinf(a); r.y = cosf(a);
#endif
    return r;
}

constexpr float kF32Epsilon = 1.1920929e-7f;       // std::f32::EPSILON
constexpr float kPi = 3.14159265358979323846f;     // std::f32::consts::PI

// approx 0.1 ulps_eq! with cgmath's defaults (epsilon = f32::EPSILON, max_ulps = 4)
RT_DI float signum(float v) { return isnan(v) ? v : (signbit(v) ? -1.0f : 1.0f); }
RT_DI bool ulps_eq(float a, float b) {
    if (fabsf(a - b) <= kF32Epsilon) return true;
    if (signum(a) != signum(b)) return false;
    long long diff = (long long)__float_as_int(a) - (long long)__float_as_int(b);
    if (diff < 0) diff = -diff;
    return diff <= 4;
}

struct quat { float s; f3 v; };
// cgmath Quaternion::from_arc(src, dst, None)
RT_DN quat from_arc(f3 src, f3 dst) {
    float mag_avg = sqrtf(dot(src, src) * dot(dst, dst));
    float d = dot(src, dst);
    quat q;
    if (ulps_eq(d, mag_avg)) {
        q.s = 1.0f; q.v = mk3(0.0f, 0.0f, 0.0f);
    } else if (ulps_eq(d, -mag_avg)) {
        f3 v = cross(mk3(1.0f, 0.0f, 0.0f), src);
        if (ulps_eq(v.x, 0.0f) && ulps_eq(v.y, 0.0f) && ulps_eq(v.z, 0.0f)) v = cross(mk3(0.0f, 1.0f, 0.0f), src);
        f3 axis = normalize(v);
        // from_axis_angle(axis, pi): sin_cos(pi/2 in f32) = (1.0, -4.371139e-8)
        q.s = -4.371139e-8f; q.v = axis * 1.0f;
    } else {
        float qs = mag_avg + d;
        f3 qv = cross(src, dst);
        float inv = 1.0f / sqrtf(qs * qs + dot(qv, qv));
        q.s = qs * inv; q.v = qv * inv;
    }
    return q;
}
// cgmath Quaternion * Vector3
RT_DI f3 rotate(quat q, f3 vec) {
    f3 tmp = cross(q.v, vec) + (vec * q.s);
    return (cross(q.v, tmp) * 2.0f) + vec;
}

// f32::is_normal
RT_DI bool is_normal_f32(float f) {
    uint32_t e = (__float_as_uint(f) >> 23) & 0xffu;
    return e != 0u && e != 0xffu;
}

// ---- Philox4x32-10 sample stream: counter (x, y, epoch, block), key (seed_lo, seed_hi) ---------
struct Rng {
    uint32_t k0, k1, x, y, epoch, draws;
    uint32_t b[4];
};
// one out-of-line copy; by-value ABI (no stack traffic)
RT_DN uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
RT_DI void philox_block(Rng& r, uint32_t block) {
    const uint4 v = philox4x32_10(r.x, r.y, r.epoch, block, r.k0, r.k1);
    r.b[0] = v.x; r.b[1] = v.y; r.b[2] = v.z; r.b[3] = v.w;
}
RT_DI void rng_init(Rng& r, uint32_t seed_lo, uint32_t seed_hi, uint32_t y, uint32_t x, uint32_t epoch) {
    r.k0 = seed_lo; r.k1 = seed_hi; r.x = x; r.y = y; r.epoch = epoch; r.draws = 0;
}
RT_DI float rng_uniform(Rng& r) {
    if ((r.draws & 3u) == 0u) philox_block(r, r.draws >> 2);
    uint32_t w = r.draws & 3u;
    uint32_t v = w == 0 ? r.b[0] : (w == 1 ? r.b[1] : (w == 2 ? r.b[2] : r.b[3]));
    r.draws++;
    return (float)(v >> 8) * 5.9604644775390625e-8f;
}
RT_DI float rng_range(Rng& r, float low, float high) { return low + (high - low) * rng_uniform(r); }

}  // namespace b200rt
