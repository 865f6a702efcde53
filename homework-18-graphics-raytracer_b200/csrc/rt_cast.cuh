// rt_cast.cuh — World::cast (main.rs:180-326) on the device.
//
// Two implementations with identical results (bit-identical prim id, face, t, position):
//
//   cast_brute_exact   every ray x primitive pair goes through the exact test, written in the
//                      reference's operation order with non-fused IEEE arithmetic (validation path).
//
//   warp_cast          the production path, a WARP-COLLECTIVE two-phase cast:
//     phase 1 (filter)  lanes hold TRIANGLES, the warp loops over its active RAYS.  Each lane keeps one
//                       packed pair of triangle records (tri a = 64*tile + lane, tri b = a + 32) in
//                       registers: 16 float2 = the plane {n, d} and three unit edge planes {m_k, w_k}.
//                       Every lane's ray is staged once in a 1 KB per-warp shared-memory slot; for each
//                       active ray j the warp reads it back as a broadcast and evaluates, with Blackwell's
//                       packed FFMA2 (two triangles per instruction, the ray as scalar-broadcast operand):
//                           nd = n.dir   num = d - n.o   t = num * rcp(nd)   p = o + t dir
//                           e_k = m_k.p + w_k            (signed in-plane distance to edge k, + slack)
//                           keep = min(e0,e1,e2,t) + A|rcp(nd)| >= 0   or   |nd| < g
//                       Two __ballot_sync give ray j's 64-bit candidate mask, kept by lane j.
//                       Idle lanes (finished pixels, other recursion branches) still work as triangle
//                       lanes, so SIMT divergence in the tracer never wastes filter throughput, and for
//                       scenes of <= 64 triangles the records never leave the register file.
//     phase 2 (confirm) each lane walks its own ray's candidates in increasing primitive index through the
//                       exact test — the same code as cast_brute_exact — which preserves the reference's
//                       "later primitive wins exact ties, spheres after triangles" rule (main.rs:229-233,
//                       298-302) and its NaN behaviour.
//
// The filter is CONSERVATIVE: it may keep a pair the exact test rejects, never the converse.  Bound
// (u = 2^-24, S = O + Vmax + Emax: ray-origin bound + largest vertex norm + longest edge):
//   |t_filter - t_ref| <= (u(14 O + 11 V + 6 E)) / |nd| + 5u T      for |nd| >= 8u
//   |e_filter - e_ref| <= |t_filter - t_ref| + u(6 O + 19 V + 30 E)
// so A = 64u*2S and B = 128u*2S (folded into w_k) leave a > 4x margin; pairs with |nd| < g = 2^-18 and
// rays that violate the assumptions (|o| > O, |dir|^2 far from 1, non-finite) skip the filter.
#pragma once
#include "rt_math.cuh"
#include "rt_types.h"

namespace b200rt {

enum : uint32_t { kFront = 0u, kBack = 1u, kBoth = 2u };
RT_DI uint32_t face_invert(uint32_t f) { return f == kFront ? kBack : (f == kBack ? kFront : kBoth); }  // main.rs:59-66

constexpr unsigned kFullMask = 0xffffffffu;

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

RT_DI void best_init(Best& b) {
    b.prim = -1; b.bf = 0; b.t = 0.0f; b.pos = mk3(0.f, 0.f, 0.f); b.a0 = b.a1 = b.a2 = 0.0f;
}

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

// Winner-only work: barycentric normal / uv (main.rs:235-252) or sphere normal / uv (main.rs:305-313).
//   exact / attr / sph: the record arrays (sc.tri_exact ... or shared-memory copies of them, same bits).
//   sphere_uv = false skips acos / atan2 of a sphere hit whose material never reads uv (the caller knows: only
//   GenerativeMaterial::approx takes uv, materials.rs:85-103); uv is then left 0.
RT_DI void finalize_hit(const DScene& sc, const Best& best, DHit& h, bool want_attrs, const float4* __restrict__ exact,
                        const float4* __restrict__ attr, const float4* __restrict__ sph, bool all_sphere_uv = true) {
    h.prim = best.prim;
    if (best.prim < 0) return;
    h.face = best.bf;
    h.t = best.t;
    h.pos = best.pos;
    if (!want_attrs) return;    // shadow rays: only "is there a hit, and how far" is used (main.rs:435-447)
    if ((uint32_t)best.prim < sc.n_tris) {
        const float4* ex = exact + 4 * (size_t)best.prim;
        const float4* at = attr + 4 * (size_t)best.prim;
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
        const float4 s = sph[j];
        h.object = sc.sph_obj[j];
        const f3 tmp = normalize(best.pos - mk3(s));                              // main.rs:306
        h.normal = best.bf ? -tmp : tmp;
        if (all_sphere_uv || sc.materials[h.object].kind == B200RT_MATERIAL_GENERATIVE) {
            h.uv.x = nl_acosf(h.normal.y) / kPi;                                         // main.rs:311
            h.uv.y = nl_atan2f(h.normal.z, h.normal.x) / (kPi * 2.0f) + 0.5f;            // main.rs:312
        }
    }
}
RT_DI void finalize_hit(const DScene& sc, const Best& best, DHit& h, bool want_attrs = true) {
    finalize_hit(sc, best, h, want_attrs, sc.tri_exact, sc.tri_attr, sc.sph);
}

// ---- brute-force exact cast (validation path, B200RT_CAST_BRUTE_EXACT) -------------------------
RT_DI void cast_brute_exact(const DScene& sc, const DRay& r, DHit& h) {
    Best best;
    best_init(best);
#pragma unroll 1
    for (uint32_t i = 0; i < sc.n_tris; ++i) tri_exact_test(sc.tri_exact + 4 * (size_t)i, (int32_t)i, r, best);
#pragma unroll 1
    for (uint32_t j = 0; j < sc.n_sph; ++j) sphere_exact_test(sc.sph[j], (int32_t)(sc.n_tris + j), r, best);
    finalize_hit(sc, best, h);
}

// ---- warp-collective two-phase cast ---------------------------------------------------------------
RT_DI float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // one MUFU.RCP, rel. error <= 2^-23
    return r;
}
RT_DI float2 pk(float a, float b) { return make_float2(a, b); }
RT_DI float2 bc2(float a) { return make_float2(a, a); }      // becomes a scalar-broadcast FFMA2 operand (R.F32)

// Packed f32x2 values held as ONE 64-bit register pair from the moment they are built.  (As float2 the compiler
// treats the halves as independent floats and re-packs them with MOVs before every FFMA2 that uses them.)
typedef unsigned long long P2;
RT_DI P2 p2_pack(float lo, float hi) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
RT_DI P2 p2_bc(float a) { P2 r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(a)); return r; }   // scalar-broadcast operand (R.F32)
RT_DI void p2_unpack(P2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
RT_DI P2 p2_fma(P2 a, P2 b, P2 c) { P2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
RT_DI P2 p2_mul(P2 a, P2 b) { P2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// One lane's packed pair of filter records (triangles a | b in the .x | .y halves): 32 registers.
struct TriPair {
    float2 nx, ny, nz, d;
    float2 m0x, m0y, m0z, w0;
    float2 m1x, m1y, m1z, w1;
    float2 m2x, m2y, m2z, w2;
};

// tri_filter layout: [tile][k = 0..7][lane] float4, so each of the 8 loads is one coalesced 512 B request.
RT_DI void load_tripair(const float4* __restrict__ tri_filter, uint32_t tile, uint32_t lane, TriPair& c) {
    const float4* base = tri_filter + (size_t)tile * 256 + lane;
    const float4 q0 = base[0], q1 = base[32], q2 = base[64], q3 = base[96], q4 = base[128], q5 = base[160],
                 q6 = base[192], q7 = base[224];
    c.nx = pk(q0.x, q0.y); c.ny = pk(q0.z, q0.w); c.nz = pk(q1.x, q1.y); c.d = pk(q1.z, q1.w);
    c.m0x = pk(q2.x, q2.y); c.m0y = pk(q2.z, q2.w); c.m0z = pk(q3.x, q3.y); c.w0 = pk(q3.z, q3.w);
    c.m1x = pk(q4.x, q4.y); c.m1y = pk(q4.z, q4.w); c.m1z = pk(q5.x, q5.y); c.w1 = pk(q5.z, q5.w);
    c.m2x = pk(q6.x, q6.y); c.m2y = pk(q6.z, q6.w); c.m2z = pk(q7.x, q7.y); c.w2 = pk(q7.z, q7.w);
}

// Filter of this lane's two triangles against one (warp-uniform) ray.  Returns keep flags.
//   cf: the ray's face-cull factor (-K for Front rays, +K for Back rays, 0 for Both; main.rs:185-188): the term
//   c = cf * nd is negative exactly for the faces the ray's mode culls, and -K|nd| + A/|nd| < 0 for every
//   |nd| >= g, so culled pairs leave the candidate set here (1 FMUL2 per triangle pair, no extra min: FMNMX3).
//   nd is a FUSED dot product: its sign can differ from the reference's non-fused n.dir when |nd| <~ 3e-7, far
//   inside the |nd| < g band that always passes; the exact culling decision is taken in confirm_tile.
//   (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false, so the reference's own
//   non-fused dot cannot be formed with packed instructions: measured, see DESIGN.md.)
RT_DI void filter_pair(const TriPair& c, float ox, float oy, float oz, float cf, float dx, float dy, float dz, float A,
                       float g, bool& keep_a, bool& keep_b) {
    const float2 nd = __ffma2_rn(c.nz, bc2(dz), __ffma2_rn(c.ny, bc2(dy), __fmul2_rn(c.nx, bc2(dx))));
    const float2 num = __ffma2_rn(c.nz, bc2(-oz), __ffma2_rn(c.ny, bc2(-oy), __ffma2_rn(c.nx, bc2(-ox), c.d)));
    const float2 r = pk(rcp_approx(nd.x), rcp_approx(nd.y));
    const float2 t = __fmul2_rn(num, r);
    const float2 cull = __fmul2_rn(nd, bc2(cf));
    const float2 px = __ffma2_rn(t, bc2(dx), bc2(ox));
    const float2 py = __ffma2_rn(t, bc2(dy), bc2(oy));
    const float2 pz = __ffma2_rn(t, bc2(dz), bc2(oz));
    const float2 e0 = __ffma2_rn(c.m0z, pz, __ffma2_rn(c.m0y, py, __ffma2_rn(c.m0x, px, c.w0)));
    const float2 e1 = __ffma2_rn(c.m1z, pz, __ffma2_rn(c.m1y, py, __ffma2_rn(c.m1x, px, c.w1)));
    const float2 e2 = __ffma2_rn(c.m2z, pz, __ffma2_rn(c.m2y, py, __ffma2_rn(c.m2x, px, c.w2)));
    const float ma = fminf(fminf(fminf(e0.x, e1.x), e2.x), fminf(t.x, cull.x));
    const float mb = fminf(fminf(fminf(e0.y, e1.y), e2.y), fminf(t.y, cull.y));
    const float2 ms = __ffma2_rn(bc2(A), pk(fabsf(r.x), fabsf(r.y)), pk(ma, mb));
    keep_a = (ms.x >= 0.0f) | (fabsf(nd.x) < g);
    keep_b = (ms.y >= 0.0f) | (fabsf(nd.y) < g);
}
constexpr float kCullK = 1099511627776.0f;   // 2^40
#ifndef B200RT_FILTER_RAYS
#define B200RT_FILTER_RAYS 2
#endif

struct CastStats {
    unsigned long long casts, confirms, fallbacks;
};

// Per-warp shared-memory slot of the cast: 32 staged rays (2 float4 each) + the 32 x 64-bit candidate masks.
constexpr int kCastSlotFloat4 = 64 + 16 + 2;      // rays, masks (uint2[32] = 16 float4), 2 float4 of tail padding
RT_DI uint2* cast_slot_masks(float4* s_rays) { return reinterpret_cast<uint2*>(s_rays + 64); }

// Phase 2 for one tile: nearest exact hit among this ray's filter candidates (bit i of `cand` = triangle
// base + i), merged into `best`.
//
// CERTIFIED SELECT.  The reference walks the candidates in index order through the exact test.  Here each
// candidate is first classified with a few instructions:
//   * face culling and the exclusion (main.rs:185-200) depend only on the sign of n.dir, which the fused dot product
//     shares with the reference's whenever |n.dir| >= g: culled / excluded candidates are dropped, exactly as the
//     walk would;
//   * the filter's estimate t_i (|t_i - t_exact| <= delta_i = A/|nd|, the bound the filter itself relies on)
//     drops candidates with t_i + delta_i < 0 (main.rs:205) or t_i - delta_i > best.t (main.rs:229-233);
//   * near-parallel candidates (|nd| < g) have no usable estimate: set U.
// Of the survivors only the one with the smallest lower bound, w, can be the nearest hit unless bounds overlap.
// U and w go through the exact test in index order; if afterwards every other survivor j has
// t_j - delta_j > best.t, the full ordered walk would have ended with the same `best`: those j lose the
// strict distance test whatever their inside test says, and no tie with them is possible.  A NaN distance (ray
// inside a triangle's plane, 0/0 at main.rs:204) makes the walk order-dependent (main.rs:229-231 accepts NaN),
// so it, like any overlap of bounds, restores `best` and takes the reference's walk over all candidates.
//   planes: the {n, d} rows (stride 4 float4) of the tile's triangles for the classification — sc.tri_exact, or a
//   shared-memory copy of the plain filter records, whose first row holds the same bits (degenerate triangles are
//   all-zero there: n.dir = 0, set U).
//   exact: the exact records {n,d}{v0,obj}{v1}{v2} of the tile's triangles (sc.tri_exact + 4 * base, or a shared-memory copy).
//   nan_out (optional): set when a NaN distance was accepted on the way - the result then depends on the nearest-so-far
//   the walk STARTED from (a caller that folds the results of separate triangle ranges must not use this one).
RT_DI void confirm_tile(const DScene& sc, uint32_t base, unsigned long long cand, bool certify, const DRay& ray,
                        Best& best, CastStats& cs, const float4* __restrict__ planes, const float4* __restrict__ exact,
                        bool* nan_out = nullptr) {
    if (cand == 0ull) return;
    unsigned long long todo = cand;
    bool fast = false;
    float lo2 = CUDART_INF_F;
    if (certify) {
        float lo1 = CUDART_INF_F;
        int32_t w = -1;
        unsigned long long rem = cand, amb = 0ull;
#pragma unroll 1
        while (rem) {
            const uint32_t i = (uint32_t)__ffsll((long long)rem) - 1u;
            rem &= rem - 1ull;
            const float4 q0 = planes[4 * i];
            const float nd = __fmaf_rn(q0.z, ray.d.z, __fmaf_rn(q0.y, ray.d.y, q0.x * ray.d.x));
            if (!(fabsf(nd) >= sc.filter_g)) {
                // near-parallel: only the reference's own non-fused dot knows the face (exactly 0 for axis-aligned
                // walls against an axis-aligned light: front faces, which every shadow ray culls — main.rs:185-188)
                const bool bfx = dot(mk3(q0), ray.d) > 0.0f;                          // primitives.rs:45
                if ((bfx && ray.face == kFront) || (!bfx && ray.face == kBack)) continue;
                if (excluded(ray, (int32_t)(base + i), bfx)) continue;
                amb |= 1ull << i;
                continue;
            }
            // |nd| >= g = 2^-18 is 16x the rounding of either form of the dot product (|n| = 1, |dir|^2 within 1e-3 of 1):
            // the sign of the fused nd is the sign of the reference's non-fused n.dir (primitives.rs:45)
            const bool bf = nd > 0.0f;
            if ((bf && ray.face == kFront) || (!bf && ray.face == kBack)) continue;   // main.rs:185-188
            if (excluded(ray, (int32_t)(base + i), bf)) continue;                     // main.rs:190-200
            const float num = __fmaf_rn(q0.z, -ray.o.z, __fmaf_rn(q0.y, -ray.o.y, __fmaf_rn(q0.x, -ray.o.x, q0.w)));
            const float r = rcp_approx(nd);
            const float t = num * r, delta = sc.filter_A * fabsf(r);
            if (t + delta < 0.0f) continue;                                           // t_exact < 0, main.rs:205
            const float lo = t - delta;
            if (best.prim >= 0 && lo > best.t) continue;                              // main.rs:229-233
            if (lo < lo1) { lo2 = lo1; lo1 = lo; w = (int32_t)i; }
            else if (lo < lo2) lo2 = lo;
        }
        todo = amb | (w >= 0 ? (1ull << w) : 0ull);
        if (todo == 0ull) return;                     // every candidate is certain to be rejected
        fast = true;
    }
    const Best saved = best;
    for (;;) {
        cs.confirms += (unsigned long long)__popcll(todo);
        bool nan_seen = false;
#pragma unroll 1
        while (todo) {                                // increasing primitive index
            const uint32_t i = (uint32_t)__ffsll((long long)todo) - 1u;
            todo &= todo - 1ull;
            tri_exact_test(exact + 4 * (size_t)i, (int32_t)(base + i), ray, best);
            nan_seen |= best.t != best.t;
        }
        if (nan_out && nan_seen) *nan_out = true;
        if (!fast) break;
        // certified: every survivor that was not tested is strictly farther than the nearest hit so far
        if (!nan_seen && (lo2 == CUDART_INF_F || (best.prim >= 0 && lo2 > best.t))) break;
        best = saved;
        todo = cand;
        fast = false;
        cs.fallbacks += 1ull;
    }
}

// ---- building blocks of the warp-collective cast ---------------------------------------------------------------
// The 32 x 64-bit candidate masks of one tile for the rays staged in s_rays (warp-collective; __syncwarp() before
// and after by the caller).
RT_DI void filter_tile(const DScene& sc, float4* __restrict__ s_rays, const TriPair& c, uint32_t n_act, uint32_t lane) {
    uint2* s_mask = cast_slot_masks(s_rays);
#if B200RT_FILTER_RAYS == 4
    // four rays per iteration: four independent FFMA2 dependency chains in flight per warp
    // (rays beyond n_act read stale slots of the 32-slot staging area, their masks are never used)
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < n_act; i0 += 4u) {
        const uint32_t i2 = min(i0 + 2u, 30u);                       // i0 = 28 is the last possible group: no overrun
        const float4 ro0 = s_rays[2 * i0 + 0], rd0 = s_rays[2 * i0 + 1];   // broadcast reads
        const float4 ro1 = s_rays[2 * i0 + 2], rd1 = s_rays[2 * i0 + 3];
        const float4 ro2 = s_rays[2 * i2 + 0], rd2 = s_rays[2 * i2 + 1];
        const float4 ro3 = s_rays[2 * i2 + 2], rd3 = s_rays[2 * i2 + 3];
        bool ka0, kb0, ka1, kb1, ka2, kb2, ka3, kb3;
        filter_pair(c, ro0.x, ro0.y, ro0.z, ro0.w, rd0.x, rd0.y, rd0.z, sc.filter_A, sc.filter_g, ka0, kb0);
        filter_pair(c, ro1.x, ro1.y, ro1.z, ro1.w, rd1.x, rd1.y, rd1.z, sc.filter_A, sc.filter_g, ka1, kb1);
        filter_pair(c, ro2.x, ro2.y, ro2.z, ro2.w, rd2.x, rd2.y, rd2.z, sc.filter_A, sc.filter_g, ka2, kb2);
        filter_pair(c, ro3.x, ro3.y, ro3.z, ro3.w, rd3.x, rd3.y, rd3.z, sc.filter_A, sc.filter_g, ka3, kb3);
        const unsigned ba0 = __ballot_sync(kFullMask, ka0), bb0 = __ballot_sync(kFullMask, kb0);
        const unsigned ba1 = __ballot_sync(kFullMask, ka1), bb1 = __ballot_sync(kFullMask, kb1);
        const unsigned ba2 = __ballot_sync(kFullMask, ka2), bb2 = __ballot_sync(kFullMask, kb2);
        const unsigned ba3 = __ballot_sync(kFullMask, ka3), bb3 = __ballot_sync(kFullMask, kb3);
        if (lane == 0u) {
            *reinterpret_cast<uint4*>(s_mask + i0) = make_uint4(ba0, bb0, ba1, bb1);
            *reinterpret_cast<uint4*>(s_mask + i2) = make_uint4(ba2, bb2, ba3, bb3);
        }
    }
#else
    // two rays per iteration: two independent FFMA2 dependency chains in flight per warp
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < n_act; i0 += 2u) {
        const float4 ro0 = s_rays[2 * i0 + 0], rd0 = s_rays[2 * i0 + 1];   // broadcast reads
        const float4 ro1 = s_rays[2 * i0 + 2], rd1 = s_rays[2 * i0 + 3];   // (odd count: a stale slot, its mask is unused)
        bool ka0, kb0, ka1, kb1;
        filter_pair(c, ro0.x, ro0.y, ro0.z, ro0.w, rd0.x, rd0.y, rd0.z, sc.filter_A, sc.filter_g, ka0, kb0);
        filter_pair(c, ro1.x, ro1.y, ro1.z, ro1.w, rd1.x, rd1.y, rd1.z, sc.filter_A, sc.filter_g, ka1, kb1);
        const unsigned ba0 = __ballot_sync(kFullMask, ka0), bb0 = __ballot_sync(kFullMask, kb0);
        const unsigned ba1 = __ballot_sync(kFullMask, ka1), bb1 = __ballot_sync(kFullMask, kb1);
        if (lane == 0u) *reinterpret_cast<uint4*>(s_mask + i0) = make_uint4(ba0, bb0, ba1, bb1);   // one STS.128
    }
#endif
}

// stage a ray in slot `slot` of the warp's staging area (with its face-cull factor)
RT_DI void stage_ray(float4* __restrict__ s_rays, uint32_t slot, const DRay& ray) {
    const float cf = ray.face == kFront ? -kCullK : (ray.face == kBack ? kCullK : 0.0f);
    s_rays[2 * slot + 0] = make_float4(ray.o.x, ray.o.y, ray.o.z, cf);
    s_rays[2 * slot + 1] = make_float4(ray.d.x, ray.d.y, ray.d.z, 0.0f);
}

// rays outside the filter's assumptions (|o| beyond the packed bound, |dir| != 1, non-finite) go through every pair exactly
RT_DI bool ray_trusted(const DScene& sc, const DRay& ray, float& dd) {
    const float oo = ray.o.x * ray.o.x + ray.o.y * ray.o.y + ray.o.z * ray.o.z;
    dd = ray.d.x * ray.d.x + ray.d.y * ray.d.y + ray.d.z * ray.d.z;
    return (oo <= sc.origin_bound * sc.origin_bound) && (fabsf(dd - 1.0f) <= 1e-3f);  // false for NaN/Inf
}

// A ray with a NaN component (the reference produces them: a zero vector normalised, 0/0 in a refraction) makes every
// plane distance NaN, and main.rs:205 / 224 / 229-233 reject nothing on NaN: each triangle that the face and exclusion
// tests (main.rs:185-200) let through replaces the nearest-so-far, whatever came before.  The ordered walk over all
// triangles therefore ends with the LAST such triangle — found here from the end, instead of 100 000 exact tests by
// one lane.  (Infinite components do not qualify: their distances can be +-inf, which compare.)
RT_DI bool ray_has_nan(const DRay& r) {
    const float oo = r.o.x * r.o.x + r.o.y * r.o.y + r.o.z * r.o.z, dd = r.d.x * r.d.x + r.d.y * r.d.y + r.d.z * r.d.z;
    return oo != oo || dd != dd;
}
RT_DI void cast_nan_ray_triangles(const DScene& sc, const DRay& r, Best& best, CastStats& cs) {
#pragma unroll 1
    for (int32_t i = (int32_t)sc.n_tris - 1; i >= 0; --i) {
        Best b;
        best_init(b);
        tri_exact_test(sc.tri_exact + 4 * (size_t)i, i, r, b);
        cs.confirms += 1ull;
        if (b.prim >= 0) { best = b; return; }
    }
}

// this ray's candidates of one tile from its filter mask (padding lanes masked off; untrusted rays: every triangle)
RT_DI unsigned long long tile_candidates(const DScene& sc, uint32_t tile, uint2 m, bool trust) {
    const uint32_t left = sc.n_tris - tile * kTileTris;              // >= 1
    const uint32_t v_lo = left >= 32u ? 0xffffffffu : ((1u << left) - 1u);
    const uint32_t v_hi = left >= 64u ? 0xffffffffu : (left > 32u ? ((1u << (left - 32u)) - 1u) : 0u);
    const uint32_t c_lo = trust ? (m.x & v_lo) : v_lo;
    const uint32_t c_hi = trust ? (m.y & v_hi) : v_hi;
    return ((unsigned long long)c_hi << 32) | (unsigned long long)c_lo;
}

// (Round 2 measured this pre-filter in packed form inside the rays-in-lanes loop - the same operations on ray pairs, FFMA2 /
// FMUL2, the spheres beside the tile in the kernel parameter: the cast of a 16-epoch 4K batch took 66.6 ms against 63.4.  A
// packed instruction occupies the FMA pipe for two issue slots: on an issue-bound kernel it saves nothing, and the four
// spheres leave it two dependent chains where this loop has four.  Not built.  Nor is one test against a sphere around all
// the scene's spheres ahead of the loop: 64.1 ms against 63.1 - most rays of the fixture room pass the enclosure.)
// spheres (main.rs:264-324): a conservative pre-filter of main.rs:265-268 in fused arithmetic for 32 spheres at a
// time — squared line-sphere distance |disp|^2 |dir|^2 - (disp.dir)^2 against r^2 with a 64u (r^2 + |disp|^2)
// slack (both sides' rounding is <= 13u of that; NaNs pass) — then the exact test of the survivors in index
// order.  Walking a mask lets lanes that pass DIFFERENT spheres run their exact tests in the same iteration.
RT_DI void cast_spheres(const DScene& sc, const DRay& ray, bool trust, float dd, Best& best, const float4* __restrict__ sph) {
    for (uint32_t j0 = 0; j0 < sc.n_sph; j0 += 32u) {
        const uint32_t nj = min(32u, sc.n_sph - j0);
        uint32_t smask = 0u;
#pragma unroll 4
        for (uint32_t j = 0; j < nj; ++j) {
            const float4 s4 = sph[j0 + j];
            const float ex = s4.x - ray.o.x, ey = s4.y - ray.o.y, ez = s4.z - ray.o.z;
            const float b = __fmaf_rn(ez, ray.d.z, __fmaf_rn(ey, ray.d.y, ex * ray.d.x));
            const float e2 = __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, ex * ex));
            const float r2 = s4.w * s4.w;
            const float d2 = __fmaf_rn(-b, b, e2 * dd);
            const float bound = __fmaf_rn(3.8146973e-6f, r2 + e2, r2);
            if (!(trust && d2 > bound)) smask |= 1u << j;
        }
        // a sphere the ray's own exclusion always rules out (main.rs:286-296: the only face a Front / Back ray can see of it
        // is the excluded one - every shadow ray that starts on a sphere) never changes `best`: sphere_exact_test returns
        // at or before its exclusion test
        const int32_t es = ray.ex_prim - (int32_t)(sc.n_tris + j0);
        if (es >= 0 && es < 32 && ((ray.face == kFront && ray.ex_face == kFront) || (ray.face == kBack && ray.ex_face == kBack) || ray.ex_face == kBoth))
            smask &= ~(1u << es);
#pragma unroll 1
        while (smask) {
            const uint32_t j = (uint32_t)__ffs((int)smask) - 1u;
            smask &= smask - 1u;
            sphere_exact_test(sph[j0 + j], (int32_t)(sc.n_tris + j0 + j), ray, best);
        }
    }
}
RT_DI void cast_spheres(const DScene& sc, const DRay& ray, bool trust, float dd, Best& best) { cast_spheres(sc, ray, trust, dd, best, sc.sph); }

// Warp-collective cast: EVERY lane of the warp must call it (converged).  `active` lanes carry a ray.
// s_rays: this warp's kCastSlotFloat4 staging slot in shared memory.  tile0: the lane's records of tile 0,
// loaded once per kernel (scenes of <= 64 triangles never reload them).
RT_DI void warp_cast(const DScene& sc, float4* __restrict__ s_rays, const TriPair& tile0, uint32_t lane, bool active,
                     const DRay& ray, DHit& hit, CastStats& cs, bool want_attrs = true) {
    const unsigned act = __ballot_sync(kFullMask, active);
    Best best;
    best_init(best);
    if (act == 0u) { hit.prim = -1; return; }
    // stage the rays COMPACTED: the k-th active lane writes slot k, so the filter loop is a plain counted
    // loop over slots (no find-first-set / mask bookkeeping per iteration)
    const uint32_t n_act = (uint32_t)__popc(act);
    const uint32_t rank = (uint32_t)__popc(act & ((1u << lane) - 1u));
    if (active) stage_ray(s_rays, rank, ray);
    const uint2* s_mask = cast_slot_masks(s_rays);
    float dd;
    const bool trust = ray_trusted(sc, ray, dd);
    __syncwarp();
    const uint32_t n_tiles = sc.n_tris_padded / kTileTris;
    const bool nan_ray = active && !trust && ray_has_nan(ray);
    for (uint32_t tile = 0; tile < n_tiles; ++tile) {
        TriPair c;
        if (tile == 0) c = tile0; else load_tripair(sc.tri_filter, tile, lane, c);
        filter_tile(sc, s_rays, c, n_act, lane);
        __syncwarp();
        if (active && !nan_ray) confirm_tile(sc, tile * kTileTris, tile_candidates(sc, tile, s_mask[rank], trust), trust, ray, best, cs,
                                 sc.tri_exact + 4 * (size_t)(tile * kTileTris), sc.tri_exact + 4 * (size_t)(tile * kTileTris));
        __syncwarp();   // masks (and, after the last tile, the ray slots) are free again
    }
    if (active) {
        if (nan_ray) cast_nan_ray_triangles(sc, ray, best, cs);
        cast_spheres(sc, ray, trust, dd, best);
        finalize_hit(sc, best, hit, want_attrs);
        cs.casts += 1ull;
    } else {
        hit.prim = -1;
    }
}

}  // namespace b200rt
