// rt_bvh.cuh — World::cast (main.rs:180-326) through the acceleration structure of rt_bvh_build.h (SURVEY 8f, N1;
// B200RT_CAST_BVH).  Same hit as the reference's walk over every triangle, bit for bit - prim id, face, t, position,
// barycentric numerators - for ANY ray; what changes is how many triangles reach the exact test.
//
// One ray per lane, two passes (the argument is in rt_bvh_build.h):
//   1. the NORMAL tree: every triangle whose normal may satisfy |n.dir| < g - where main.rs:204 can overflow or divide
//      0 by 0 and the reference registers hits at t = +inf / NaN wherever the triangle lies - goes through the exact test;
//   2. the SPATIAL tree: a node is skipped when the ray's line misses its box inflated by rho_geom + the ray's rounding slack, or meets
//      it only at parameters below zero (main.rs:205) or beyond the nearest hit so far (main.rs:229-233); the triangles
//      of the leaves reached go through the exact test.
// The exact test is the reference's own (tri_exact_eval = main.rs:184-227 verbatim).  The nearest rule of the walk, "skip
// if best.t < t" in index order (main.rs:229-233), is applied in its order-independent form: the smallest t wins, the
// LATER primitive wins a tie.  That form is the walk's result unless a NaN distance is accepted (a NaN never loses a
// comparison: the walk is then order-dependent) - such rays, and rays outside the cast's assumptions (non-finite, far
// origins, non-unit directions), take the ordered walk over every triangle instead, by the whole warp
// (rl_coop_exact_tile).  Spheres follow the triangles as in the reference (cast_spheres).
#pragma once
#include "rt_cast.cuh"
#include "rt_cast_rl.cuh"

namespace b200rt {

constexpr float kBvhBand = 5.0e-6f;           // > g = 2^-18 plus the rounding of the bound formed from a normal box
constexpr uint32_t kBvhLeafBit = 0x80000000u;
constexpr int kBvhStack = 64;

// main.rs:184-227 for one pair: everything but the nearest-so-far rule.  Same expressions as tri_exact_test (rt_cast.cuh).
RT_DI bool tri_exact_eval(const float4* __restrict__ rec, int32_t i, const DRay& r, Best& cand) {
    const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q3 = rec[3];
    const f3 n = mk3(q0);
    const float nd = dot(n, r.d);
    const bool bf = nd > 0.0f;                                                    // primitives.rs:45
    if ((bf && r.face == kFront) || (!bf && r.face == kBack)) return false;       // main.rs:185-188
    if (excluded(r, i, bf)) return false;                                         // main.rs:190-200
    const float t = (q0.w - dot(n, r.o)) / nd;                                    // main.rs:204
    if (t <= 0.0f) return false;                                                  // main.rs:205
    const f3 p = r.o + r.d * t;                                                   // main.rs:210
    const f3 v0 = mk3(q1), v1 = mk3(q2), v2 = mk3(q3);
    const float a0 = dot(cross(v2 - v1, p - v1), n);                              // main.rs:219
    const float a1 = dot(cross(v0 - v2, p - v2), n);                              // main.rs:220
    const float a2 = dot(cross(v1 - v0, p - v0), n);                              // main.rs:221
    if (a0 < 0.0f || a1 < 0.0f || a2 < 0.0f) return false;                        // main.rs:224
    cand.prim = i; cand.bf = bf ? 1u : 0u; cand.t = t; cand.pos = p;
    cand.a0 = a0; cand.a1 = a1; cand.a2 = a2;
    return true;
}

// The conservative filter of the two-phase cast (rt_cast_rl.cuh: rl_filter_tile) for ONE pair, on the same plain record
// and with the same operations: false only where the exact test is certain to reject (rays inside the filter's
// assumptions; near-parallel pairs, degenerate records and NaNs pass).  ~25 instructions against the ~100 and two
// divisions of the exact test, which 98 % of the triangles in the leaves a ray reaches do not need.
RT_DI bool bvh_filter_pair(const DScene& sc, uint32_t i, const DRay& ray) {
    const float4* __restrict__ f = sc.tri_filter_plain + 4 * (size_t)i;
    const float4 q0 = f[0], q1 = f[1], q2 = f[2], q3 = f[3];
    const float nd = __fmaf_rn(q0.z, ray.d.z, __fmaf_rn(q0.y, ray.d.y, q0.x * ray.d.x));
    const float num = __fmaf_rn(-q0.z, ray.o.z, __fmaf_rn(-q0.y, ray.o.y, __fmaf_rn(-q0.x, ray.o.x, q0.w)));
    const float r = rcp_approx(nd);
    const float t = num * r;
    const float px = __fmaf_rn(t, ray.d.x, ray.o.x), py = __fmaf_rn(t, ray.d.y, ray.o.y), pz = __fmaf_rn(t, ray.d.z, ray.o.z);
    const float e0 = __fmaf_rn(q1.z, pz, __fmaf_rn(q1.y, py, __fmaf_rn(q1.x, px, q1.w)));
    const float e1 = __fmaf_rn(q2.z, pz, __fmaf_rn(q2.y, py, __fmaf_rn(q2.x, px, q2.w)));
    const float e2 = __fmaf_rn(-q3.y, e1, __fmaf_rn(-q3.x, e0, q3.z));
    const float cull = ray.face == kFront ? -r : (ray.face == kBack ? r : CUDART_INF_F);
    const float m = fminf(fminf(fminf(e0, e1), e2), fminf(t, cull));
    const float ms = __fmaf_rn(sc.filter_As, fabsf(r), m);
    return (__float_as_uint(ms) >> 31) == 0u;
}

// main.rs:229-233 over the accepted triangles in index order == the smallest t, the later index on a tie (no NaN)
template <bool FILTER = true>
RT_DI void bvh_try_triangle(const DScene& sc, uint32_t i, const DRay& ray, Best& best, bool& nan_seen, uint32_t& tested) {
    if (FILTER && !bvh_filter_pair(sc, i, ray)) return;
    Best cand;
    tested += 1u;
    if (!tri_exact_eval(sc.tri_exact + 4 * (size_t)i, (int32_t)i, ray, cand)) return;
    if (cand.t != cand.t) { nan_seen = true; return; }
    if (best.prim < 0 || cand.t < best.t || (cand.t == best.t && cand.prim > best.prim)) best = cand;
}

// Nearest triangle hit of a ray INSIDE the cast's assumptions.  nan_seen: a NaN distance was accepted - the caller must
// redo the ray with the ordered walk.  tested: exact tests run (statistics).
RT_DI void bvh_cast_triangles(const DScene& sc, const DRay& ray, Best& best, bool& nan_seen, uint32_t& tested) {
    nan_seen = false;
    uint32_t stack[kBvhStack];
    // ---- pass 1: triangles that may be (nearly) parallel to the ray --------------------------------------------------
    if (sc.nbvh_n_nodes != 0u) {
        int sp = 0;
        uint32_t node = 0u;
        for (;;) {
            const float4 a = sc.nbvh_nodes[2 * (size_t)node], b = sc.nbvh_nodes[2 * (size_t)node + 1];
            // the range of n.dir over the node's normal box
            const float lo = (fminf(a.x * ray.d.x, b.x * ray.d.x) + fminf(a.y * ray.d.y, b.y * ray.d.y)) + fminf(a.z * ray.d.z, b.z * ray.d.z);
            const float hi = (fmaxf(a.x * ray.d.x, b.x * ray.d.x) + fmaxf(a.y * ray.d.y, b.y * ray.d.y)) + fmaxf(a.z * ray.d.z, b.z * ray.d.z);
            const uint32_t w1 = __float_as_uint(a.w), w2 = __float_as_uint(b.w);
            if (!(lo > kBvhBand) && !(hi < -kBvhBand)) {          // |n.dir| < g possible in this node
                if (!(w2 & kBvhLeafBit)) {
                    if (sp < kBvhStack) stack[sp++] = w2;
                    node = w1;
                    continue;
                }
                const uint32_t count = w2 & ~kBvhLeafBit;
#pragma unroll 1
                for (uint32_t k = 0; k < count; ++k) {
                    // of the leaf's triangles only those that ARE nearly parallel (the reference's own n.dir), and those the
                    // spatial tree does not hold (flagged), are this pass's business; NaN normals compare false and stay
                    const uint32_t e = sc.nbvh_tris[w1 + k], i = e & ~kBvhLeafBit;
                    const float nd = dot(mk3(sc.tri_exact[4 * (size_t)i]), ray.d);        // the reference's own n.dir (primitives.rs:45)
                    if (!(e & kBvhLeafBit) && fabsf(nd) >= kBvhBand) continue;
                    // main.rs:185-188 decided here (the exact test starts with the same comparison on the same bits): a light
                    // that lies in the plane of a thousand triangles sends every one of its shadow rays through this loop
                    const bool bf = nd > 0.0f;
                    if ((bf && ray.face == kFront) || (!bf && ray.face == kBack)) continue;
                    bvh_try_triangle<false>(sc, i, ray, best, nan_seen, tested);        // (near-parallel pairs pass the filter anyway)
                }
            }
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
    // ---- pass 2: triangles near the ray's line --------------------------------------------------------------------
    if (sc.bvh_n_nodes == 0u) return;
    const float ix = 1.0f / ray.d.x, iy = 1.0f / ray.d.y, iz = 1.0f / ray.d.z;
    // what separates the reference's plane point from the line and from the plane (rt_bvh_build.h), for THIS ray: with
    // |o|, |t| <= 2 |o| + V and u = 2^-24 it is below u (10 V + 25 |o|); four times that, plus the slab arithmetic's own rounding
    const float slack = 5.9604645e-8f * (40.0f * sc.scene_extent + 100.0f * ((fabsf(ray.o.x) + fabsf(ray.o.y)) + fabsf(ray.o.z)));
    int sp = 0;
    uint32_t node = 0u;
    for (;;) {
        const float4* __restrict__ N = sc.bvh_nodes + 3 * (size_t)node;
        const float4 a = N[0], b = N[1], c = N[2];
        const float infl = a.w + slack;
        // slabs of the inflated box; fminf / fmaxf drop the NaN of 0 * inf (origin on a slab plane, direction in it)
        const float x0 = ((a.x - infl) - ray.o.x) * ix, x1 = ((b.x + infl) - ray.o.x) * ix;
        const float y0 = ((a.y - infl) - ray.o.y) * iy, y1 = ((b.y + infl) - ray.o.y) * iy;
        const float z0 = ((a.z - infl) - ray.o.z) * iz, z1 = ((b.z + infl) - ray.o.z) * iz;
        const float t_in = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
        const float t_out = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
        // (main.rs:229-233: a farther t loses; main.rs:205: t <= 0 loses; 1e-5 relative: the rounding of the slab arithmetic)
        const float t_hi = best.prim >= 0 ? best.t + 1.0e-5f * fabsf(best.t) : CUDART_INF_F;
        const bool reach = !(t_in > t_out) && !(t_out < 0.0f) && !(t_in > t_hi);
        const uint32_t w1 = __float_as_uint(c.x), w2 = __float_as_uint(c.y);
        if (reach && !(w2 & kBvhLeafBit)) {
            // the child on the ray's side of the split first: nearer hits shrink the parameter range early
            const uint32_t axis = __float_as_uint(c.z);
            const float da = axis == 0u ? ray.d.x : (axis == 1u ? ray.d.y : ray.d.z);
            const uint32_t first = da >= 0.0f ? w1 : w2, second = da >= 0.0f ? w2 : w1;
            if (sp < kBvhStack) stack[sp++] = second;
            node = first;
            continue;
        }
        if (reach) {
            const uint32_t count = w2 & ~kBvhLeafBit;
#pragma unroll 1
            for (uint32_t k = 0; k < count; ++k) bvh_try_triangle(sc, sc.bvh_tris[w1 + k], ray, best, nan_seen, tested);
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
}

// The reference's ordered walk over every triangle for ONE ray, by the whole warp (every lane calls it with the same ray).
RT_DI void bvh_coop_walk(const DScene& sc, const DRay& r, uint32_t lane, Best& out, CastStats& cs) {
    best_init(out);
    bool bvalid = false, changed = false;
    float bt = 0.0f;
    const uint32_t n_tiles = sc.n_tris_padded / kTileTris;
#pragma unroll 1
    for (uint32_t tile = 0; tile < n_tiles; ++tile) rl_coop_exact_tile(sc, tile, r, lane, bvalid, bt, out, changed, cs);
    if (!changed) best_init(out);
}

// Warp-collective (all 32 lanes, converged): World::cast for every active lane's ray.
RT_DI void bvh_warp_cast(const DScene& sc, uint32_t lane, bool active, const DRay& ray, DHit& hit, CastStats& cs, bool want_attrs = true,
                         bool all_sphere_uv = true) {
    Best best;
    best_init(best);
    float dd = 1.0f;
    const bool trust = active && ray_trusted(sc, ray, dd);
    const bool has_nan = active && !trust && ray_has_nan(ray);
    bool nan_seen = false;
    uint32_t tested = 0u;
    if (trust) bvh_cast_triangles(sc, ray, best, nan_seen, tested);
    cs.confirms += tested;
    if (has_nan) cast_nan_ray_triangles(sc, ray, best, cs);
    // rays the tree cannot answer: the ordered walk, one ray at a time, by the whole warp
    unsigned redo = __ballot_sync(kFullMask, active && !has_nan && (!trust || nan_seen));
#pragma unroll 1
    while (redo) {
        const int l = __ffs((int)redo) - 1;
        redo &= redo - 1u;
        DRay r;
        r.o.x = __shfl_sync(kFullMask, ray.o.x, l); r.o.y = __shfl_sync(kFullMask, ray.o.y, l); r.o.z = __shfl_sync(kFullMask, ray.o.z, l);
        r.d.x = __shfl_sync(kFullMask, ray.d.x, l); r.d.y = __shfl_sync(kFullMask, ray.d.y, l); r.d.z = __shfl_sync(kFullMask, ray.d.z, l);
        r.face = __shfl_sync(kFullMask, ray.face, l); r.ex_prim = __shfl_sync(kFullMask, ray.ex_prim, l); r.ex_face = __shfl_sync(kFullMask, ray.ex_face, l);
        Best w;
        bvh_coop_walk(sc, r, lane, w, cs);
        if ((int)lane == l) { best = w; cs.fallbacks += 1ull; }
    }
    hit.prim = -1;
    if (active) {
        cast_spheres(sc, ray, trust, dd, best);
        finalize_hit(sc, best, hit, want_attrs, sc.tri_exact, sc.tri_attr, sc.sph, all_sphere_uv);
        cs.casts += 1ull;
    }
}

}  // namespace b200rt
