// rt_cast.cuh — World::cast (main.rs:180-326) on the device.
//
// Two implementations with identical results (bit-identical prim id, face, t, position):
//
//   cast_brute_exact   every ray x primitive pair goes through the exact test, written in the
//                      reference's operation order with non-fused IEEE arithmetic.
//   cast_two_phase     phase 1: a branch-free FMA "plane + 3 edge planes" filter over the packed
//                      64-byte tri_filter records staged in shared memory produces a 64-bit
//                      candidate mask per 64-triangle tile (the FP32-roofline loop: 20 FFMA-pipe
//                      instructions per pair).  The filter is conservative: its slack (folded into
//                      the edge-plane offsets at upload) exceeds the worst-case rounding gap between
//                      the fused filter and the reference's formula, so it never rejects a pair the
//                      exact test would accept.
//                      phase 2: the few surviving candidates (typically 1-4 of 64) are confirmed in
//                      increasing index order by the same exact test as cast_brute_exact, which keeps
//                      the reference's "later primitive wins ties, spheres after triangles" rule.
#pragma once
#include "rt_math.cuh"
#include "rt_types.h"

namespace b200rt {

enum : uint32_t { kFront = 0u, kBack = 1u, kBoth = 2u };
RT_DI uint32_t face_invert(uint32_t f) { return f == kFront ? kBack : (f == kBack ? kFront : kBoth); }  // main.rs:59-66

struct DRay {  // main.rs:69-81
    f3 o, d;
    uint32_t face;
    int32_t ex_prim;   // -1 = no exclusion
    uint32_t ex_face;
};

struct DHit {  // main.rs:139-147
    int32_t prim;      // -1 = None
    uint32_t face;     // kFront / kBack
    uint32_t object;
    float t;
    f3 pos, normal;
    f2 uv;
};

// nearest-so-far record carried through a cast
struct Best {
    int32_t prim;
    uint32_t bf;
    float t;
    f3 pos;
    float a0, a1, a2;  // edge areas of the winning triangle (barycentric numerators, main.rs:218-222)
};

// exclusion criteria, main.rs:190-200 / 286-296
RT_DI bool excluded(const DRay& r, int32_t prim, bool bf) {
    if (r.ex_prim != prim) return false;
    return r.ex_face == kFront ? !bf : (r.ex_face == kBack ? bf : true);
}

// One ray x triangle pair, exactly as main.rs:184-233 evaluates it.  n and d are the per-triangle
// values the reference recomputes for every pair (primitives.rs:37-42, main.rs:203); they were
// computed once on the host with the same operation order, so the bits are the same.
RT_DI void tri_exact_test(const float4* __restrict__ rec, int32_t i, const DRay& r, Best& best) {
    const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q3 = rec[3];
    const f3 n = mk3(q0);
    const float nd = dot(n, r.d);
    const bool bf = nd > 0.0f;                                                    // primitives.rs:45
    if ((bf && r.face == kFront) || (!bf && r.face == kBack)) return;             // main.rs:185-188
    if (excluded(r, i, bf)) return;                                               // main.rs:190-200
    const float t = (q0.w - dot(n, r.o)) / nd;                                    // main.rs:204
    if (t <= 0.0f) return;                                                        // main.rs:205
    const f3 p = r.o + r.d * t;                                                   // main.rs:210
    const f3 v0 = mk3(q1), v1 = mk3(q2), v2 = mk3(q3);
    const float a0 = dot(cross(v2 - v1, p - v1), n);                              // main.rs:219
    const float a1 = dot(cross(v0 - v2, p - v2), n);                              // main.rs:220
    const float a2 = dot(cross(v1 - v0, p - v0), n);                              // main.rs:221
    if (a0 < 0.0f || a1 < 0.0f || a2 < 0.0f) return;                              // main.rs:224
    if (best.prim >= 0 && best.t < t) return;                                     // main.rs:229-233
    best.prim = i; best.bf = bf ? 1u : 0u; best.t = t; best.pos = p;
    best.a0 = a0; best.a1 = a1; best.a2 = a2;
}

// One ray x sphere pair, main.rs:265-302
RT_DI void sphere_exact_test(float4 s, int32_t prim, const DRay& r, Best& best) {
    const f3 c = mk3(s);
    const float lsd = magnitude(cross(c - r.o, r.d));                             // main.rs:265
    if (lsd > s.w) return;                                                        // main.rs:266
    const f3 disp = c - r.o;                                                      // main.rs:270
    const float tc = dot(r.d, disp);                                              // main.rs:271
    const float k = sqrtf(s.w * s.w - lsd * lsd);                                 // main.rs:272
    float t; bool bf;
    if (r.face == kFront) { t = tc - k; bf = false; }                             // main.rs:274
    else if (r.face == kBack) { t = tc + k; bf = true; }                          // main.rs:275
    else if (tc < k) { t = tc + k; bf = true; }                                   // main.rs:276-277
    else { t = tc - k; bf = false; }                                              // main.rs:279
    if (t <= 0.0f) return;                                                        // main.rs:282
    if (excluded(r, prim, bf)) return;                                            // main.rs:286-296
    if (best.prim >= 0 && best.t < t) return;                                     // main.rs:298-302
    best.prim = prim; best.bf = bf ? 1u : 0u; best.t = t;
    best.pos = r.o + r.d * t;                                                     // main.rs:304
}

// Winner-only work: barycentric normal / uv (main.rs:235-252) or sphere normal / uv (main.rs:305-313)
RT_DI void finalize_hit(const DScene& sc, const DRay& r, const Best& best, DHit& h) {
    h.prim = best.prim;
    if (best.prim < 0) return;
    h.face = best.bf;
    h.t = best.t;
    h.pos = best.pos;
    if ((uint32_t)best.prim < sc.n_tris) {
        const float4* ex = sc.tri_exact + 4 * (size_t)best.prim;
        const float4* at = sc.tri_attr + 4 * (size_t)best.prim;
        const float4 q0 = ex[0], q1 = ex[1], q2 = ex[2], q3 = ex[3];
        const float4 t0 = at[0], t1 = at[1], t2 = at[2], t3 = at[3];
        const f3 n = mk3(q0), v0 = mk3(q1), v1 = mk3(q2), v2 = mk3(q3);
        h.object = __float_as_uint(q1.w);
        const float area = dot(cross(v1 - v0, v2 - v0), n);                       // main.rs:235
        const float b0 = best.a0 / area, b1 = best.a1 / area, b2 = best.a2 / area;  // main.rs:236
        const f3 tmp = (mk3(t0) * b0 + mk3(t1) * b1) + mk3(t2) * b2;              // main.rs:249
        h.normal = best.bf ? -tmp : tmp;                                          // main.rs:250
        const float u0 = t0.w, w0 = t1.w, u1 = t2.w, w1 = t3.x, u2 = t3.y, w2 = t3.z;
        h.uv.x = (u0 * b0 + u1 * b1) + u2 * b2;                                   // main.rs:252
        h.uv.y = (w0 * b0 + w1 * b1) + w2 * b2;
    } else {
        const uint32_t j = (uint32_t)best.prim - sc.n_tris;
        const float4 s = sc.sph[j];
        h.object = sc.sph_obj[j];
        const f3 tmp = normalize(best.pos - mk3(s));                              // main.rs:306
        h.normal = best.bf ? -tmp : tmp;
        h.uv.x = acosf(h.normal.y) / kPi;                                         // main.rs:311
        h.uv.y = atan2f(h.normal.z, h.normal.x) / (kPi * 2.0f) + 0.5f;            // main.rs:312
    }
}

// ---- brute-force exact cast (validation path, B200RT_CAST_BRUTE_EXACT) -------------------------
RT_DI void cast_brute_exact(const DScene& sc, const DRay& r, DHit& h) {
    Best best;
    best.prim = -1; best.bf = 0; best.t = 0.0f; best.pos = mk3(0.f, 0.f, 0.f); best.a0 = best.a1 = best.a2 = 0.0f;
    for (uint32_t i = 0; i < sc.n_tris; ++i) tri_exact_test(sc.tri_exact + 4 * (size_t)i, (int32_t)i, r, best);
    for (uint32_t j = 0; j < sc.n_sph; ++j) sphere_exact_test(sc.sph[j], (int32_t)(sc.n_tris + j), r, best);
    finalize_hit(sc, r, best, h);
}

// ---- two-phase cast -------------------------------------------------------------------------------
// Filter math for one pair (all explicit FMAs; 20 FFMA-pipe + 1 MUFU + 2 FMNMX3 + 1 SHF):
//   nd  = n.d                      num = d - n.o                t = num * rcp(nd)
//   p   = o + t*d                  e_k = m_k.p - c_k            (c_k already holds -slack)
//   cul = cull_eps - s*nd          (s = +1 front rays, -1 back rays, 0 both)
//   keep = min(e0, e1, e2, t + t_eps, cul) >= 0                 (NaN keeps: min() drops NaN operands
//                                                                only if another operand is not NaN;
//                                                                an all-NaN/Inf pair is caught by the
//                                                                |nd| guard folded into `cul`)
// The reject bit is the sign bit of that min, funnel-shifted into the tile mask.
struct FilterRay {
    float ox, oy, oz, dx, dy, dz;
    float s;         // cull sign
};

RT_DI float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }

template <int kCount>
RT_DI void filter_tile(const float4* __restrict__ tile, const FilterRay& fr, float t_eps, float cull_eps,
                       uint32_t& rej_lo, uint32_t& rej_hi) {
    // tile: kCount records of 4 float4 in shared memory; all lanes read the same address (broadcast).
    uint32_t lo = 0u, hi = 0u;
#pragma unroll 8
    for (int i = 0; i < kCount; ++i) {
        const float4 q0 = tile[4 * i + 0], q1 = tile[4 * i + 1], q2 = tile[4 * i + 2], q3 = tile[4 * i + 3];
        const float nd = __fmaf_rn(q0.z, fr.dz, __fmaf_rn(q0.y, fr.dy, q0.x * fr.dx));
        const float num = __fmaf_rn(-q0.z, fr.oz, __fmaf_rn(-q0.y, fr.oy, __fmaf_rn(-q0.x, fr.ox, q0.w)));
        const float t = num * __frcp_rn(nd);
        const float px = __fmaf_rn(t, fr.dx, fr.ox), py = __fmaf_rn(t, fr.dy, fr.oy), pz = __fmaf_rn(t, fr.dz, fr.oz);
        const float e0 = __fmaf_rn(q1.z, pz, __fmaf_rn(q1.y, py, __fmaf_rn(q1.x, px, -q1.w)));
        const float e1 = __fmaf_rn(q2.z, pz, __fmaf_rn(q2.y, py, __fmaf_rn(q2.x, px, -q2.w)));
        const float e2 = __fmaf_rn(q3.z, pz, __fmaf_rn(q3.y, py, __fmaf_rn(q3.x, px, -q3.w)));
        const float cul = __fmaf_rn(-fr.s, nd, cull_eps);
        const float m = fminf(min3f(e0, e1, e2), fminf(t + t_eps, cul));
        // reject iff m < 0 (sign bit set and not -0/NaN-with-sign issues: handled below)
        const uint32_t sgn = __float_as_uint(m);
        if (i < 32) lo = __funnelshift_l(sgn, lo, 1); else hi = __funnelshift_l(sgn, hi, 1);
    }
    rej_lo = lo; rej_hi = hi;
}

}  // namespace b200rt
