// rt_shade.cuh — material / light evaluators and ray constructors (reference L1 + parts of L2).
#pragma once
#include "rt_cast.cuh"

namespace b200rt {

// ColorMaterial after Material::approx (materials.rs:33-37, 85-103)
struct MatEval {
    f3 normal_ts, diffuse, specular;
    float shiness, smoothness, transparency, refraction_index, opaque_decay;
};

RT_DI MatEval material_approx(const DMaterial* __restrict__ mats, uint32_t object, f2 uv) {
    const DMaterial& m = mats[object];
    MatEval e;
    e.normal_ts = mk3(m.normal);
    e.diffuse = mk3(m.diffuse);
    e.specular = mk3(m.specular);
    e.shiness = m.shiness;
    e.smoothness = m.smoothness;
    e.transparency = m.transparency;
    e.refraction_index = m.refraction_index;
    e.opaque_decay = m.opaque_decay;
    if (m.kind == B200RT_MATERIAL_GENERATIVE) {
        const float* p = m.fn_params;
        if (m.diffuse_fn == B200RT_DIFFUSE_STRIPE_V) {
            // `(uv.y * 20.0) as i32 % 2 == 0`, main.rs:849; (int) is cvt.rzi: saturating, NaN -> 0 like Rust
            const bool even = ((int)(uv.y * p[0])) % 2 == 0;
            e.diffuse = even ? mk3(p + 1) : mk3(p + 4);
        } else if (m.diffuse_fn == B200RT_DIFFUSE_CHECKER_UPV) {
            const bool even = ((int)((uv.x + uv.y) * p[0])) % 2 == 0;                    // main.rs:1020
            e.diffuse = even ? mk3(p + 1) : mk3(p + 4);
        }
        if (m.normal_fn == B200RT_NORMAL_SINCOS_U) {
            const float angle = uv.x * p[7] * 2.0f * kPi;                                // main.rs:856
            const float2 sc2 = nl_sincosf(angle);
            f3 v = mk3(sc2.x, 0.0f, sc2.y);
            if (dot(v, mk3(0.0f, 0.0f, 1.0f)) <= 0.0f) v = -v;                           // main.rs:858-862
            e.normal_ts = v;
        }
    }
    return e;
}

// materials.rs:40-44
RT_DI f3 adjust_normal(const MatEval& m, f3 normal) { return rotate(from_arc(mk3(0.0f, 0.0f, 1.0f), normal), m.normal_ts); }

// powf(x, e) of the Phong lobe (materials.rs:63: x = max(R.V, 0) in [0, 1 + ulps], e = 1 / (smoothness + eps) up to 8.4e6):
// a colour (rt_math.cuh: color_pow)
RT_DI float phong_pow(float x, float e) { return color_pow(x, e); }

// materials.rs:46-53
RT_DI f3 get_diffuse(const MatEval& m, f3 n, f3 l) {
    const float cosine = dot(l, n);
    if (cosine > 0.0f) return m.diffuse * cosine;
    return mk3(0.0f, 0.0f, 0.0f);
}
// materials.rs:55-66
RT_DI f3 get_specular(const MatEval& m, f3 n, f3 view, f3 l) {
    const float cosine = dot(l, n);
    if (cosine <= 0.0f) return mk3(0.0f, 0.0f, 0.0f);
    const f3 reflected_ray = 2.0f * cosine * n - l;
    const float specular = 1.0f / (m.smoothness + kF32Epsilon);
    const float energy_conserving = (specular + 8.0f) / (8.0f * kPi);
    const float amount = phong_pow(fmaxf(dot(reflected_ray, view), 0.0f), specular) * energy_conserving;
    return m.specular * amount;
}
// The two material constants of get_specular (materials.rs:60-62), for callers that evaluate several lights: the same
// expressions, hence the same bits, without an IEEE division per light.
struct SpecConst { float exponent, energy; };
RT_DI SpecConst spec_const(const MatEval& m) {
    SpecConst c;
    c.exponent = 1.0f / (m.smoothness + kF32Epsilon);
    c.energy = (c.exponent + 8.0f) / (8.0f * kPi);
    return c;
}
RT_DI f3 get_specular(const MatEval& m, const SpecConst& c, f3 n, f3 view, f3 l) {
    const float cosine = dot(l, n);
    if (cosine <= 0.0f) return mk3(0.0f, 0.0f, 0.0f);
    const f3 reflected_ray = 2.0f * cosine * n - l;
    const float amount = phong_pow(fmaxf(dot(reflected_ray, view), 0.0f), c.exponent) * c.energy;
    return m.specular * amount;
}

struct DirLight {  // lights.rs:6-11
    bool has_origin;
    f3 origin, dir, color;
    float angular;   // spot lights: (1 - angle/spread)^(softness + eps), lights.rs:62-64 (1 otherwise)
};

// lights.rs:48-93
RT_DI bool approx_light(const DLight& L, f3 position, DirLight& out) {
    if (L.kind == B200RT_LIGHT_DIRECTIONAL) {
        out.has_origin = L.has_origin != 0u;
        out.origin = mk3(L.origin);
        out.dir = mk3(L.direction);
        out.color = mk3(L.color);
        out.angular = 1.0f;
        return true;
    }
    const f3 origin = mk3(L.origin);
    const f3 offset = position - origin;
    if (L.kind == B200RT_LIGHT_SPOT) {
        const f3 sd = mk3(L.direction);
        const float angle = fabsf(nl_atan2f(magnitude(cross(sd, offset)), dot(sd, offset)));  // Vector3::angle
        if (angle > L.angle) return false;
        const float angular = color_pow(1.0f - angle / L.angle, L.softness + kF32Epsilon);
        const float dist_att = 1.0f / (magnitude(offset) + kF32Epsilon);
        out.has_origin = true;
        out.origin = origin;
        out.dir = normalize(position - origin);
        out.color = mk3(L.color) * angular * dist_att;
        out.angular = angular;
        return true;
    }
    if (L.kind == B200RT_LIGHT_POINT) {
        const float dist_att = 1.0f / (magnitude(offset) + kF32Epsilon);
        out.has_origin = true;
        out.origin = origin;
        out.dir = normalize(offset);
        out.color = mk3(L.color) * dist_att;
        out.angular = 1.0f;
        return true;
    }
    return false;
}

// The same Directional as approx_light() for a light that is known to reach `position`, from what the wavefront kept
// when it requested the shadow ray: its direction (dir = normalize(position - origin), lights.rs:66 / 80) and the spot's
// angular factor.  Skips the atan2 / powf / normalize of the second evaluation; every value has the bits of the first.
RT_DI void approx_light_cached(const DLight& L, f3 position, f3 dir, float angular, DirLight& out, float& dist_to_origin) {
    out.dir = dir;
    out.angular = angular;
    if (L.kind == B200RT_LIGHT_DIRECTIONAL) {
        out.has_origin = L.has_origin != 0u;
        out.origin = mk3(L.origin);
        out.color = mk3(L.color);
        dist_to_origin = -1.0f;   // not formed here
        return;
    }
    const f3 origin = mk3(L.origin);
    const f3 offset = position - origin;
    const float dist = magnitude(offset);
    const float dist_att = 1.0f / (dist + kF32Epsilon);
    out.has_origin = true;
    out.origin = origin;
    out.color = L.kind == B200RT_LIGHT_SPOT ? mk3(L.color) * angular * dist_att : mk3(L.color) * dist_att;
    dist_to_origin = dist;   // = distance(position, origin) bit for bit: |a - b| and |b - a| square the same components
}

// main.rs:328-341 (normal = hit.at.normal, l = hit.ray.direction)
RT_DI DRay make_reflect(f3 pos, f3 normal, f3 l, uint32_t ray_face, int32_t prim, uint32_t hit_face) {
    DRay r;
    const f3 reflected = l - 2.0f * dot(l, normal) * normal;
    r.o = pos;
    r.d = normalize(reflected);
    r.face = ray_face;
    r.ex_prim = prim;
    r.ex_face = face_invert(hit_face);
    return r;
}

// closure main.rs:344-352
RT_DI bool refract_dir(f3 n, f3 l, float k, f3& out) {
    const float c = -dot(l, n);
    if (k * k >= 1.0f - c * c) {
        const f3 x = (l + n * c) / k - n * sqrtf(1.0f - (1.0f - c * c) / (k * k));
        out = normalize(x);
        return true;
    }
    return false;
}

}  // namespace b200rt
