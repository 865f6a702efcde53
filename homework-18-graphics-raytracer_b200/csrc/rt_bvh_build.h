// rt_bvh_build.h — host-side builder of the acceleration structure for World::cast.
//
// The reference tests every ray against every triangle (main.rs:183).  B200RT_CAST_BVH keeps the RESULT of that walk, bit
// for bit, and skips the triangles a ray provably cannot hit (rt_bvh.cuh).  Two facts about the reference's own test
// (main.rs:184-227, f32, non-fused) make a skip provable for a ray inside the cast's assumptions (finite, |o| within the
// packed origin bound O, |dir| = 1 to 1e-3) and a well-shaped triangle (kappa = 1 / sin(theta_min / 2) <= 32):
//
//   * |n.dir| >= g = 2^-18: t is finite, and the test accepts only if the plane point p = o + t dir it forms - a point of
//     the ray's line, to rounding - projects into the triangle or within 32 u kappa E' of it (u = 2^-24, E' = the
//     triangle's diameter): an edge function (e x (p - v)).n carries a rounding error of at most 8 u |e| |p - v|, and
//     outside the triangle at distance D one of the three is below -|e| D / kappa.  p is also within u (10 V + 25 |o|) of
//     the plane and of the line (the error of t scales with 1 / |n.dir|, but it moves p ALONG the line: off the plane it
//     is |d_plane| u + 4 u |n.o| from the numerator and 5.2 u |t| from the rounding of n.dir, |t| <= 2 |o| + V for a point
//     near the triangle; p = o + t dir adds 3.5 u (|o| + |t|)).  So the LINE passes within
//         rho = 1e-4 kappa E' + 4 u (10 (V + E) + 25 |o|)
//     of the triangle, at a parameter t that lies in the line's parameter range inside the triangle's bounding box
//     inflated by rho (50x and 4x the bounds).
//   * |n.dir| < g: the division of main.rs:204 may overflow or be 0/0 - the reference then registers hits at t = +inf or
//     NaN whatever the triangle's position (main.rs:205, 224, 229-231 reject nothing on inf / NaN).  Such pairs must
//     reach the exact test wherever the triangle lies.
//
// Hence TWO trees over the same triangles:
//   the SPATIAL tree  bounds positions; a node stores the exact bounding box of its triangles' vertices and rho_geom =
//                     the largest 1e-4 kappa E' of its triangles.  It prunes pairs of the first kind.
//   the NORMAL tree   bounds the unit normals (the {n} of the exact records, the bits n.dir is formed from); the
//                     traversal walks the nodes whose normal box admits |n.dir| < g for the ray at hand - a thin band
//                     around a great circle, O(sqrt N) leaves of a smooth mesh - and finds the pairs of the second kind.
// Triangles that are not well shaped (a zero-area or needle triangle, a non-finite vertex, no unit normal) get the
// normal box [-1, 1]^3: every ray reaches them through the normal tree, as every ray tests them in the reference.
//
// Both are binary trees built by binned SAH (16 bins on the longest centroid axis) with leaves of <= 4 triangles:
//   spatial node (48 B)  {bmin.xyz, rho_geom} {bmax.xyz, -} {u32 left | first, u32 right | 0x80000000 + count, u32 axis, -}
//   normal node  (32 B)  {nmin.xyz, u32 left | first} {nmax.xyz, u32 right | 0x80000000 + count}
// (the normal tree's index entries carry 0x80000000 for the triangles that are not in the spatial tree)
// The left child holds the smaller centroids along `axis`.  Triangle order inside the trees is free: the traversal applies
// the reference's nearest / tie rule (main.rs:229-233) in its order-independent form.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace b200rt {

struct BvhBuild {
    std::vector<float4> nodes;         // spatial tree: 3 float4 per node, root = node 0
    std::vector<uint32_t> tri_index;   // its leaves index this permutation of the triangles
    std::vector<float4> nnodes;        // normal tree: 2 float4 per node
    std::vector<uint32_t> ntri_index;
    uint32_t max_depth = 0, n_leaves = 0, nmax_depth = 0, nn_leaves = 0;
};

constexpr uint32_t kBvhLeafFlag = 0x80000000u;
constexpr uint32_t kBvhLeafTris = 4u;
constexpr uint32_t kBvhMaxDepth = 62u;       // the traversal's stack
constexpr double kBvhKappaMax = 32.0;        // well-shaped: theta_min >= 3.6 degrees

namespace bvh_detail {
struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; } }
    void add(const float* p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    void add(const Box& b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    double area() const {
        const double e[3] = {(double)hi[0] - lo[0], (double)hi[1] - lo[1], (double)hi[2] - lo[2]};
        return (e[0] < 0 || e[1] < 0 || e[2] < 0) ? 0.0 : 2.0 * (e[0] * e[1] + e[1] * e[2] + e[2] * e[0]);
    }
};
struct Prim { Box box; float centroid[3]; float rho; };
struct TreeNode { Box box; float rho; uint32_t w1, w2, axis; };
inline float u2f_bits(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }

// binary tree over prims[index[...]] (binned SAH, leaves of <= kBvhLeafTris); index is permuted in place
inline void build_tree(const std::vector<Prim>& prims, std::vector<uint32_t>& index, std::vector<TreeNode>& nodes,
                       uint32_t& max_depth, uint32_t& n_leaves) {
    nodes.clear(); max_depth = 0; n_leaves = 0;
    if (index.empty()) return;
    struct Task { uint32_t node, first, count, depth; };
    std::vector<Task> todo;
    nodes.resize(1);
    todo.push_back(Task{0u, 0u, (uint32_t)index.size(), 0u});
    constexpr int kBins = 16;
    while (!todo.empty()) {
        const Task t = todo.back();
        todo.pop_back();
        max_depth = std::max(max_depth, t.depth);
        Box box, cbox;
        box.reset(); cbox.reset();
        float rho = 0.0f;
        for (uint32_t k = 0; k < t.count; ++k) {
            const Prim& p = prims[index[t.first + k]];
            box.add(p.box); cbox.add(p.centroid);
            rho = std::max(rho, p.rho);
        }
        int axis = 0;
        float ext = -1.0f;
        for (int c = 0; c < 3; ++c) if (cbox.hi[c] - cbox.lo[c] > ext) { ext = cbox.hi[c] - cbox.lo[c]; axis = c; }
        uint32_t n_left = 0;
        const bool make_leaf = t.count <= kBvhLeafTris || t.depth >= kBvhMaxDepth - 1u;
        if (!make_leaf) {
            if (ext > 0.0f && std::isfinite(ext)) {
                Box bb[kBins]; uint32_t bn[kBins];
                for (int b = 0; b < kBins; ++b) { bb[b].reset(); bn[b] = 0; }
                const double scale = (double)kBins / ((double)cbox.hi[axis] - cbox.lo[axis]);
                auto bin_of = [&](const Prim& p) { return std::min(kBins - 1, std::max(0, (int)(((double)p.centroid[axis] - cbox.lo[axis]) * scale))); };
                for (uint32_t k = 0; k < t.count; ++k) {
                    const Prim& p = prims[index[t.first + k]];
                    const int b = bin_of(p);
                    bb[b].add(p.box); bn[b]++;
                }
                double right_area[kBins]; uint32_t right_n[kBins];
                Box acc; acc.reset(); uint32_t cn = 0;
                for (int b = kBins - 1; b > 0; --b) { acc.add(bb[b]); cn += bn[b]; right_area[b] = acc.area(); right_n[b] = cn; }
                acc.reset(); cn = 0;
                double best = INFINITY; int best_b = -1;
                for (int b = 0; b + 1 < kBins; ++b) {
                    acc.add(bb[b]); cn += bn[b];
                    if (cn == 0 || right_n[b + 1] == 0) continue;
                    const double cost = acc.area() * cn + right_area[b + 1] * right_n[b + 1];
                    if (cost < best) { best = cost; best_b = b; }
                }
                if (best_b >= 0) {
                    auto mid = std::partition(index.begin() + t.first, index.begin() + t.first + t.count,
                                              [&](uint32_t i) { return bin_of(prims[i]) <= best_b; });
                    n_left = (uint32_t)(mid - (index.begin() + t.first));
                }
            }
            if (n_left == 0 || n_left == t.count) {   // (coincident centroids: split the list in the middle)
                n_left = t.count / 2;
                std::nth_element(index.begin() + t.first, index.begin() + t.first + n_left, index.begin() + t.first + t.count,
                                 [&](uint32_t a, uint32_t b) { return prims[a].centroid[axis] < prims[b].centroid[axis] || (prims[a].centroid[axis] == prims[b].centroid[axis] && a < b); });
            }
        }
        TreeNode nd;
        nd.box = box; nd.rho = rho; nd.axis = (uint32_t)axis;
        if (n_left == 0) {
            nd.w1 = t.first; nd.w2 = kBvhLeafFlag | t.count;
            n_leaves++;
        } else {
            const uint32_t left = (uint32_t)nodes.size(), right = left + 1;
            nodes.resize(nodes.size() + 2);
            nd.w1 = left; nd.w2 = right;
            todo.push_back(Task{right, t.first + n_left, t.count - n_left, t.depth + 1});
            todo.push_back(Task{left, t.first, n_left, t.depth + 1});
        }
        nodes[t.node] = nd;
    }
}
}  // namespace bvh_detail

// tri_exact: [tri][4] float4 {n, d} {v0, obj} {v1} {v2} (the exact records of the cast)
inline void build_bvh(const float4* tri_exact, uint32_t n_tris, BvhBuild& out) {
    using namespace bvh_detail;
    out = BvhBuild();
    if (n_tris == 0) return;
    std::vector<Prim> sp, np;          // spatial / normal-space primitives
    std::vector<uint32_t> sp_ids;      // triangles the spatial tree holds (the well-shaped ones)
    sp.resize(n_tris); np.resize(n_tris);
    for (uint32_t i = 0; i < n_tris; ++i) {
        const float4* r = tri_exact + 4 * (size_t)i;
        const float v[3][3] = {{r[1].x, r[1].y, r[1].z}, {r[2].x, r[2].y, r[2].z}, {r[3].x, r[3].y, r[3].z}};
        bool finite = true;
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 3; ++c) finite = finite && std::isfinite(v[k][c]);
        // kappa = 1 / sin(theta_min / 2) and the diameter E' (f64)
        double kappa = INFINITY, diam = 0.0;
        if (finite) {
            double smin = 1.0;
            for (int k = 0; k < 3; ++k) {
                const float* a = v[k]; const float* b = v[(k + 1) % 3]; const float* c = v[(k + 2) % 3];
                const double e1[3] = {(double)b[0] - a[0], (double)b[1] - a[1], (double)b[2] - a[2]};
                const double e2[3] = {(double)c[0] - a[0], (double)c[1] - a[1], (double)c[2] - a[2]};
                const double l1 = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]), l2 = std::sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
                diam = std::max(diam, std::max(l1, l2));
                if (!(l1 > 0.0) || !(l2 > 0.0)) { smin = 0.0; break; }
                double cs = (e1[0] * e2[0] + e1[1] * e2[1] + e1[2] * e2[2]) / (l1 * l2);
                cs = std::min(1.0, std::max(-1.0, cs));
                smin = std::min(smin, std::sqrt(0.5 * (1.0 - cs)));       // sin(theta / 2)
            }
            if (smin > 0.0) kappa = 1.0 / smin;
        }
        const float n[3] = {r[0].x, r[0].y, r[0].z};
        const double nn = (double)n[0] * n[0] + (double)n[1] * n[1] + (double)n[2] * n[2];
        const bool regular = finite && std::isfinite(nn) && std::fabs(nn - 1.0) < 1e-3 && std::isfinite(r[0].w) && kappa <= kBvhKappaMax &&
                             std::isfinite(diam);
        Prim& q = np[i];
        q.box.reset(); q.rho = 0.0f;
        if (regular) {
            q.box.add(n);
            for (int c = 0; c < 3; ++c) q.centroid[c] = n[c];
            Prim& p = sp[i];
            p.box.reset();
            for (int k = 0; k < 3; ++k) p.box.add(v[k]);
            for (int c = 0; c < 3; ++c) p.centroid[c] = (float)(((double)v[0][c] + v[1][c] + v[2][c]) / 3.0);
            p.rho = (float)(1.0e-4 * kappa * diam * 1.000001);
            sp_ids.push_back(i);
        } else {
            const float m1[3] = {-1.f, -1.f, -1.f}, p1[3] = {1.f, 1.f, 1.f};
            q.box.add(m1); q.box.add(p1);
            for (int c = 0; c < 3; ++c) q.centroid[c] = 0.0f;
        }
    }
    std::vector<TreeNode> tn;
    out.tri_index = sp_ids;
    build_tree(sp, out.tri_index, tn, out.max_depth, out.n_leaves);
    out.nodes.resize(3 * tn.size());
    for (size_t k = 0; k < tn.size(); ++k) {
        const TreeNode& t = tn[k];
        out.nodes[3 * k + 0] = make_float4(t.box.lo[0], t.box.lo[1], t.box.lo[2], t.rho);
        out.nodes[3 * k + 1] = make_float4(t.box.hi[0], t.box.hi[1], t.box.hi[2], 0.0f);
        out.nodes[3 * k + 2] = make_float4(u2f_bits(t.w1), u2f_bits(t.w2), u2f_bits(t.axis), 0.0f);
    }
    out.ntri_index.resize(n_tris);
    for (uint32_t i = 0; i < n_tris; ++i) out.ntri_index[i] = i;
    build_tree(np, out.ntri_index, tn, out.nmax_depth, out.nn_leaves);
    // (a triangle that is not well shaped is not in the spatial tree: its entry carries a flag, and every ray tests it)
    {
        std::vector<uint8_t> in_spatial(n_tris, 0);
        for (uint32_t i : sp_ids) in_spatial[i] = 1;
        for (uint32_t& e : out.ntri_index) if (!in_spatial[e]) e |= kBvhLeafFlag;
    }
    out.nnodes.resize(2 * tn.size());
    for (size_t k = 0; k < tn.size(); ++k) {
        const TreeNode& t = tn[k];
        out.nnodes[2 * k + 0] = make_float4(t.box.lo[0], t.box.lo[1], t.box.lo[2], u2f_bits(t.w1));
        out.nnodes[2 * k + 1] = make_float4(t.box.hi[0], t.box.hi[1], t.box.hi[2], u2f_bits(t.w2));
    }
}

}  // namespace b200rt
