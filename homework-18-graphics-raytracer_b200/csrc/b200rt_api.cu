// b200rt_api.cu — the C ABI of libb200rt.so: context, scene packing, render entry points.
//
// Host code in this file is built with -ffp-contract=off: the per-triangle invariants it hoists
// (face normal, plane offset; primitives.rs:37-42, main.rs:203) and the camera basis (main.rs:85-92)
// must have the same bits the reference computes per pair / per pixel.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b200rt.h"
#include "b200rt_dev.h"
#include "hvec.h"
#include "rt_types.h"
#include "rt_bvh_build.h"
#include "rt_wavefront.h"

using namespace b200rt;
using namespace b200rt_host;

struct b200rt_ctx {
    int device = -1;
    int sm_count = 0;
    int sm_clock_khz = 0;
    size_t hbm_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    std::string last_cuda_error;

    // device scene
    bool have_scene = false;
    DScene scene{};
    void* d_scene_blob = nullptr;
    size_t scene_blob_bytes = 0;

    // host copy of the exact records (for re-deriving the filter records when the origin bound grows)
    std::vector<float4> h_tri_exact;
    std::vector<float4> h_sph;
    float scene_radius = 0.0f;   // max |vertex|, max |centre| + r
    float max_edge = 0.0f;
    void* d_filter = nullptr;  size_t d_filter_bytes = 0;
    void* d_bvh = nullptr;     size_t d_bvh_bytes = 0;      // B200RT_CAST_BVH: nodes, then the triangle permutation
    uint32_t bvh_depth = 0, bvh_leaves = 0;
    RlTileParam h_tile0{};     // tile 0 of the rays-in-lanes filter records, passed to the cast kernels by value

    // device scratch for the host-buffer entry points
    void* d_out = nullptr;  size_t d_out_bytes = 0;
    void* d_aux = nullptr;  size_t d_aux_bytes = 0;
    DCounters* d_cnt = nullptr;
    void* d_pp = nullptr;   // post_process control block
    // wavefront tracer: path state / ray queues in HBM, a pinned word and an event to poll the retired counter
    void* d_wf = nullptr;   size_t d_wf_bytes = 0;
    uint32_t* h_poll = nullptr;
    cudaEvent_t ev_poll = nullptr;
    uint32_t last_rounds = 0, last_launches = 0;
    bool kernel_timing = false;
    WfKernelTiming wf_timing;

    b200rt_stats stats{};
};

namespace {

int cuda_fail(b200rt_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) {
        ctx->last_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
    }
    return B200RT_ERR_CUDA;
}

#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

int ensure(b200rt_ctx* ctx, void** p, size_t* have, size_t need) {
    if (*have >= need && *p) return B200RT_OK;
    if (*p) { CU(cudaFree(*p)); *p = nullptr; *have = 0; }
    CU(cudaMalloc(p, need));
    *have = need;
    return B200RT_OK;
}

// Filter records of the two-phase cast (rt_cast.cuh): unit edge planes in f64, slack B folded into w_k.
// Layout [tile][k][lane] float4 with triangles a = 64*tile + lane, b = a + 32 interleaved.
int pack_filter(b200rt_ctx* ctx, float origin_bound) {
    const uint32_t nt = ctx->scene.n_tris;
    const uint32_t n_tiles = ctx->scene.n_tris_padded / kTileTris;
    const double u = 5.9604644775390625e-8;   // 2^-24
    const double S = 2.0 * ((double)origin_bound + (double)ctx->scene_radius + (double)ctx->max_edge);
    const double A = 64.0 * u * S, B = 128.0 * u * S;
    std::vector<float4> rec((size_t)std::max(n_tiles, 1u) * 256 * 2, make_float4(0.f, 0.f, 0.f, 0.f));
    std::memset(&ctx->h_tile0, 0, sizeof ctx->h_tile0);
    float4* plain = rec.data() + (size_t)std::max(n_tiles, 1u) * 256;   // second half: [tri][4] plain records
    auto up = [](double v) { float f = (float)v; if ((double)f < v) f = std::nextafter(f, INFINITY); return f; };
    for (uint32_t tile = 0; tile < n_tiles; ++tile) {
        for (uint32_t lane = 0; lane < 32; ++lane) {
            float vals[2][16];
            for (int h = 0; h < 2; ++h) {
                for (int q = 0; q < 16; ++q) vals[h][q] = 0.0f;   // all-zero record: n.dir = 0 -> always a candidate
                const uint32_t idx = tile * kTileTris + (uint32_t)h * 32u + lane;
                if (idx >= nt) continue;
                const float4* ex = &ctx->h_tri_exact[4 * (size_t)idx];
                const double n[3] = {ex[0].x, ex[0].y, ex[0].z};
                const double v[3][3] = {{ex[1].x, ex[1].y, ex[1].z}, {ex[2].x, ex[2].y, ex[2].z}, {ex[3].x, ex[3].y, ex[3].z}};
                // edge k and its anchor vertex, in the order of main.rs:219-221
                const int ea[3] = {2, 0, 1}, eb[3] = {1, 2, 0};   // e_k = v[ea] - v[eb], anchor v[eb]
                float out[16];
                out[0] = ex[0].x; out[1] = ex[0].y; out[2] = ex[0].z; out[3] = ex[0].w;
                bool ok = std::isfinite(ex[0].x) && std::isfinite(ex[0].y) && std::isfinite(ex[0].z) && std::isfinite(ex[0].w);
                double L[3] = {0.0, 0.0, 0.0}, w_plain[3] = {0.0, 0.0, 0.0};   // edge lengths |n x e_k| and unslacked offsets
                for (int k = 0; k < 3 && ok; ++k) {
                    const double e[3] = {v[ea[k]][0] - v[eb[k]][0], v[ea[k]][1] - v[eb[k]][1], v[ea[k]][2] - v[eb[k]][2]};
                    const double M[3] = {n[1] * e[2] - n[2] * e[1], n[2] * e[0] - n[0] * e[2], n[0] * e[1] - n[1] * e[0]};
                    const double len = std::sqrt(M[0] * M[0] + M[1] * M[1] + M[2] * M[2]);
                    if (!(len > 0.0) || !std::isfinite(len)) { ok = false; break; }
                    const double m[3] = {M[0] / len, M[1] / len, M[2] / len};
                    const double c = m[0] * v[eb[k]][0] + m[1] * v[eb[k]][1] + m[2] * v[eb[k]][2];
                    out[4 + 4 * k + 0] = (float)m[0]; out[4 + 4 * k + 1] = (float)m[1]; out[4 + 4 * k + 2] = (float)m[2];
                    out[4 + 4 * k + 3] = up(-c + B);
                    L[k] = len; w_plain[k] = -c;
                }
                if (ok) for (int q = 0; q < 16; ++q) vals[h][q] = out[q];
                // plain record (rays-in-lanes loop): the edge function of the LONGEST edge r follows from the other two,
                // L_p e_p + L_q e_q + L_r e_r = 2 Area  =>  e_r = c - a e_p - b e_q, a = L_p / L_r, b = L_q / L_r (<= 1).
                // The loop evaluates it on the slacked e_p + B, e_q + B: c' = c + (a + b) B undoes their slack, + B is the
                // edge's own (a > 4x margin over the rounding of a direct evaluation), + 16 u S covers what the
                // dependent form adds to that rounding: a eps_p + b eps_q + two FMAs, eps <= 5 u S each.
                // The plane row is scaled by 2^-108 (exact; an all-zero record stays all-zero: always a candidate).
                float pl[16];
                for (int q = 0; q < 16; ++q) pl[q] = vals[h][q];
                if (ok) {
                    const int r = (L[0] >= L[1] && L[0] >= L[2]) ? 0 : (L[1] >= L[2] ? 1 : 2);
                    const int pq[2] = {(r + 1) % 3, (r + 2) % 3};
                    const double a = L[pq[0]] / L[r], b = L[pq[1]] / L[r];
                    const double c = (L[0] * w_plain[0] + L[1] * w_plain[1] + L[2] * w_plain[2]) / L[r];
                    for (int q = 0; q < 4; ++q) { pl[q] = out[q] * kRlPlaneScale; pl[4 + q] = out[4 + 4 * pq[0] + q]; pl[8 + q] = out[4 + 4 * pq[1] + q]; }
                    pl[12] = (float)a; pl[13] = (float)b; pl[14] = up(c + B * (1.0 + a + b) + 16.0 * u * S); pl[15] = 0.0f;
                }
                for (int k = 0; k < 4; ++k)
                    plain[4 * (size_t)idx + k] = make_float4(pl[4 * k], pl[4 * k + 1], pl[4 * k + 2], pl[4 * k + 3]);
                if (tile == 0) {   // the same record as a kernel parameter: multipliers | addends
                    float4* tp = ctx->h_tile0.rec + 4 * (size_t)idx;
                    tp[0] = make_float4(pl[0], pl[1], pl[2], pl[4]);
                    tp[1] = make_float4(pl[5], pl[6], pl[8], pl[9]);
                    tp[2] = make_float4(pl[10], pl[12], pl[13], 0.0f);
                    tp[3] = make_float4(pl[3], pl[7], pl[11], pl[14]);
                }
            }
            // interleave {a,b}: entry q -> float2; two entries per float4
            for (int k = 0; k < 8; ++k)
                rec[((size_t)tile * 8 + k) * 32 + lane] =
                    make_float4(vals[0][2 * k], vals[1][2 * k], vals[0][2 * k + 1], vals[1][2 * k + 1]);
        }
    }
    // Plane runs of tile 0 (the kernel-parameter records): CONSECUTIVE triangles of (nearly) one plane - the two halves of a
    // square, the fan of a polygon - share n.dir, d - n.o, the reciprocal, t and the plane point in the loop, which
    // evaluates them once per run, on the plane of the run's first triangle (flag in the spare slot of the record).
    // Consecutive only: candidate bits stay in primitive order, which the exact walk's tie rule needs (main.rs:229-233).
    // A member's own plane {n, d} (the reference's bits) differs from the run's by dn, dd (a few ulps: its normal comes
    // from other edges): that moves the filter's t by at most (dd + |dn| (O + T)) / |n.dir|, T <= O + V + E - a slack of
    // the A kind, added (twice over) to the A the run loop uses.  Phase 2 classifies with each triangle's own plane.
    double a_runs = 0.0;
    {
        const double O = origin_bound, V = ctx->scene_radius, E = ctx->max_edge;
        bool have_run = false;
        double rn[3] = {0, 0, 0}, rd = 0;
        for (uint32_t idx = 0; idx < (uint32_t)kTileTris; ++idx) {
            float4* tp = ctx->h_tile0.rec + 4 * (size_t)idx;
            bool start = true;
            const bool live = idx < nt && (tp[0].x != 0.0f || tp[0].y != 0.0f || tp[0].z != 0.0f);   // (degenerate: all-zero record)
            if (live) {
                const float4 e0 = ctx->h_tri_exact[4 * (size_t)idx];
                const double n[3] = {e0.x, e0.y, e0.z}, d = e0.w;
                if (have_run) {
                    const double dn = std::sqrt((n[0] - rn[0]) * (n[0] - rn[0]) + (n[1] - rn[1]) * (n[1] - rn[1]) + (n[2] - rn[2]) * (n[2] - rn[2]));
                    const double dd = std::fabs(d - rd);
                    if (dn <= 2e-6 && dd <= 2e-6 * std::max(1.0, std::fabs(rd))) {
                        start = false;
                        a_runs = std::max(a_runs, dd + dn * (2.0 * O + V + E));
                    }
                }
                if (start) { rn[0] = n[0]; rn[1] = n[1]; rn[2] = n[2]; rd = d; }
                have_run = true;
            } else have_run = false;
            tp[2].w = start ? 1.0f : 0.0f;
        }
    }
    CU(cudaDeviceSynchronize());   // a kernel of an earlier launch may still read the old records
    int rc = ensure(ctx, &ctx->d_filter, &ctx->d_filter_bytes, rec.size() * sizeof(float4));
    if (rc != B200RT_OK) return rc;
    CU(cudaMemcpyAsync(ctx->d_filter, rec.data(), rec.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->scene.tri_filter = reinterpret_cast<const float4*>(ctx->d_filter);
    ctx->scene.tri_filter_plain = ctx->scene.tri_filter + (size_t)std::max(n_tiles, 1u) * 256;
    ctx->scene.origin_bound = origin_bound;
    ctx->scene.filter_A = up(A);
    ctx->scene.filter_As = ctx->scene.filter_A * kRlPlaneScale;
    ctx->scene.filter_As_runs = up(A + 2.0 * a_runs) * kRlPlaneScale;
    ctx->scene.h_tile0 = &ctx->h_tile0;
    ctx->scene.filter_g = 3.814697265625e-6f;   // 2^-18
    ctx->scene.filter_B = up(B);
    ctx->scene.scene_extent = up((double)ctx->scene_radius + (double)ctx->max_edge);
    return B200RT_OK;
}

// the exact record of a triangle {n, d} {v0, obj} {v1} {v2}: Triangle::face_normal (primitives.rs:37-42) and d = n . v0
// (main.rs:203) in the reference's arithmetic, once per triangle instead of once per pair
void exact_record(const b200rt_triangle& t, float4* out) {
    const V3 v0 = v3(t.vertices[0].position), v1 = v3(t.vertices[1].position), v2 = v3(t.vertices[2].position);
    const V3 a = v1 - v0, b = v2 - v1;
    const V3 n = normalize(cross(a, b));
    const float d = dot(n, v0);
    float obj_bits;
    std::memcpy(&obj_bits, &t.object_index, 4);
    out[0] = make_float4(n.x, n.y, n.z, d);
    out[1] = make_float4(v0.x, v0.y, v0.z, obj_bits);
    out[2] = make_float4(v1.x, v1.y, v1.z, 0.0f);
    out[3] = make_float4(v2.x, v2.y, v2.z, 0.0f);
}

// B200RT_CAST_BVH: the hierarchy over the scene's triangles (host build, rt_bvh_build.h), uploaded behind the scene
int build_and_upload_bvh(b200rt_ctx* ctx) {
    DScene& sc = ctx->scene;
    sc.bvh_nodes = sc.nbvh_nodes = nullptr; sc.bvh_tris = sc.nbvh_tris = nullptr; sc.bvh_n_nodes = sc.nbvh_n_nodes = 0u;
    ctx->bvh_depth = ctx->bvh_leaves = 0u;
    if (sc.n_tris == 0u) return B200RT_OK;
    BvhBuild bvh;
    build_bvh(ctx->h_tri_exact.data(), sc.n_tris, bvh);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b0 = bvh.nodes.size() * sizeof(float4), b1 = bvh.nnodes.size() * sizeof(float4);
    const size_t b2 = bvh.tri_index.size() * sizeof(uint32_t), b3 = bvh.ntri_index.size() * sizeof(uint32_t);
    const size_t o1 = al(b0), o2 = o1 + al(b1), o3 = o2 + al(b2);
    int rc = ensure(ctx, &ctx->d_bvh, &ctx->d_bvh_bytes, o3 + al(b3) + 256);
    if (rc != B200RT_OK) return rc;
    char* base = (char*)ctx->d_bvh;
    if (b0) CU(cudaMemcpyAsync(base, bvh.nodes.data(), b0, cudaMemcpyHostToDevice, ctx->stream));
    if (b1) CU(cudaMemcpyAsync(base + o1, bvh.nnodes.data(), b1, cudaMemcpyHostToDevice, ctx->stream));
    if (b2) CU(cudaMemcpyAsync(base + o2, bvh.tri_index.data(), b2, cudaMemcpyHostToDevice, ctx->stream));
    if (b3) CU(cudaMemcpyAsync(base + o3, bvh.ntri_index.data(), b3, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    sc.bvh_nodes = reinterpret_cast<const float4*>(base);
    sc.nbvh_nodes = reinterpret_cast<const float4*>(base + o1);
    sc.bvh_tris = reinterpret_cast<const uint32_t*>(base + o2);
    sc.nbvh_tris = reinterpret_cast<const uint32_t*>(base + o3);
    sc.bvh_n_nodes = (uint32_t)(bvh.nodes.size() / 3);
    sc.nbvh_n_nodes = (uint32_t)(bvh.nnodes.size() / 2);
    ctx->bvh_depth = std::max(bvh.max_depth, bvh.nmax_depth); ctx->bvh_leaves = bvh.n_leaves;
    return B200RT_OK;
}

// Rays leave the camera from center + toward*near (- lens offsets): make sure the filter's origin bound covers it.
int ensure_origin_bound(b200rt_ctx* ctx, const b200rt_camera& cam, const b200rt_params& p) {
    const double c = std::sqrt((double)cam.center[0] * cam.center[0] + (double)cam.center[1] * cam.center[1] +
                               (double)cam.center[2] * cam.center[2]);
    const double need = c + std::fabs((double)cam.near) * 2.0 + 16.0 * std::fabs((double)p.blur) + 1.0;
    if (need <= (double)ctx->scene.origin_bound) return B200RT_OK;
    return pack_filter(ctx, (float)(2.0 * need));
}

// Camera::shoot hoisted (main.rs:85-92)
void make_camera(const b200rt_camera& c, DCamera& out) {
    V3 toward = normalize(v3(c.toward));
    V3 right = normalize(cross(toward, v3(c.up)));
    V3 up = normalize(cross(right, toward));
    float t = std::tan(c.fovy / 2.0f);
    store(out.toward, toward);
    store(out.x, t * right);
    store(out.y, t * up);
    store(out.origin, v3(c.center) + toward * c.near);
    store(out.center, v3(c.center));
    out.near = c.near;
}

int make_params(const b200rt_params& p, uint32_t epoch_begin, uint32_t epoch_count, DParams& o) {
    if (p.width == 0 || p.height == 0) return B200RT_ERR_INVALID;
    if (p.depth < 0 || p.depth > B200RT_MAX_DEPTH) return B200RT_ERR_UNSUPPORTED;
    if (p.row_count && (p.row_begin >= p.height || p.row_count > p.height - p.row_begin)) return B200RT_ERR_INVALID;   // (no u32 wrap)
    if (p.cast_mode > B200RT_CAST_BVH) return B200RT_ERR_INVALID;
    if (p.tracer > B200RT_TRACER_MEGAKERNEL) return B200RT_ERR_INVALID;
    o.width = p.width; o.height = p.height;
    o.row_begin = p.row_count ? p.row_begin : 0u;
    o.row_count = p.row_count ? p.row_count : p.height;
    o.strip_rows = 1u << 30; o.strip_parts = 1u; o.strip_part = 0u; o.local_row0 = 0u;   // contiguous rows
    o.depth = p.depth;
    o.threshold = p.threshold;
    o.refract_max_distance = p.refract_max_distance;
    o.tir_retries = p.tir_retries;
    o.focus = p.focus; o.blur = p.blur;
    o.seed_lo = (uint32_t)p.seed; o.seed_hi = (uint32_t)(p.seed >> 32);
    o.cast_mode = p.cast_mode;
    o.epoch_begin = epoch_begin; o.epoch_count = epoch_count;
    return B200RT_OK;
}

int fetch_counters(b200rt_ctx* ctx) {
    DCounters h;
    CU(cudaMemcpyAsync(&h, ctx->d_cnt, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stats.casts = h.casts;
    ctx->stats.tri_pair_tests = h.tri_pairs;
    ctx->stats.sph_pair_tests = h.sph_pairs;
    ctx->stats.exact_confirms = h.confirms;
    ctx->stats.samples = h.samples;
    ctx->stats.certify_fallbacks = h.fallbacks;
    ctx->stats.wavefront_rounds = ctx->last_rounds + (uint32_t)h.rounds;       // host-enqueued rounds + the device loop's
    ctx->stats.cast_kernel_ms = (float)ctx->wf_timing.cast_ms;
    ctx->stats.logic_kernel_ms = (float)ctx->wf_timing.logic_ms;
    ctx->stats.cast_kernel_launches = (uint32_t)ctx->wf_timing.cast_launches;
    ctx->stats.primary_kernel_ms = (float)ctx->wf_timing.primary_ms;
    ctx->stats.kernel_launches = ctx->last_launches + (uint32_t)h.launches;
    return B200RT_OK;
}

}  // namespace

extern "C" {

const char* b200rt_strerror(int code) {
    switch (code) {
        case B200RT_OK: return "ok";
        case B200RT_ERR_INVALID: return "invalid argument";
        case B200RT_ERR_CUDA: return "CUDA runtime error (see b200rt_last_cuda_error)";
        case B200RT_ERR_NO_SCENE: return "no scene uploaded";
        case B200RT_ERR_NO_DEVICE: return "no usable CUDA device (libb200rt has no CPU fallback)";
        case B200RT_ERR_IO: return "OBJ file could not be read or parsed";
        case B200RT_ERR_UNSUPPORTED: return "unsupported parameter (depth > B200RT_MAX_DEPTH?)";
        case B200RT_ERR_NCCL: return "NCCL unavailable or an NCCL call failed (see b200rt_group_last_error)";
        default: return "unknown error";
    }
}

const char* b200rt_last_cuda_error(const b200rt_ctx* ctx) { return ctx ? ctx->last_cuda_error.c_str() : ""; }

int b200rt_create(int device_id, b200rt_ctx** out_ctx) {
    if (!out_ctx) return B200RT_ERR_INVALID;
    *out_ctx = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return B200RT_ERR_NO_DEVICE;
    if (device_id < 0 || device_id >= n) return B200RT_ERR_NO_DEVICE;
    b200rt_ctx* ctx = new b200rt_ctx();
    ctx->device = device_id;
    cudaDeviceProp prop;
    if (cudaSetDevice(device_id) != cudaSuccess || cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) {
        delete ctx;
        return B200RT_ERR_NO_DEVICE;
    }
    ctx->sm_count = prop.multiProcessorCount;
    // (cudaLimitMaxL2FetchGranularity = 32 / 64 / 128 was measured on B200 in round 2: no effect on any kernel of the tracer.)
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device_id);
    ctx->sm_clock_khz = khz;
    ctx->hbm_bytes = prop.totalGlobalMem;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreate(&ctx->ev2) != cudaSuccess || cudaEventCreate(&ctx->ev3) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_poll, cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_poll, sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_cnt, sizeof(DCounters)) != cudaSuccess ||
        cudaMalloc(&ctx->d_pp, post_process_workspace_bytes() + sizeof(float)) != cudaSuccess ||
        cudaMemset(ctx->d_cnt, 0, sizeof(DCounters)) != cudaSuccess) {
        b200rt_destroy(ctx);
        return B200RT_ERR_CUDA;
    }
    *out_ctx = ctx;
    return B200RT_OK;
}

int b200rt_destroy(b200rt_ctx* ctx) {
    if (!ctx) return B200RT_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->d_scene_blob) cudaFree(ctx->d_scene_blob);
    if (ctx->d_filter) cudaFree(ctx->d_filter);
    if (ctx->d_bvh) cudaFree(ctx->d_bvh);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_aux) cudaFree(ctx->d_aux);
    if (ctx->d_cnt) cudaFree(ctx->d_cnt);
    if (ctx->d_pp) cudaFree(ctx->d_pp);
    if (ctx->d_wf) cudaFree(ctx->d_wf);
    if (ctx->h_poll) cudaFreeHost(ctx->h_poll);
    if (ctx->ev_poll) cudaEventDestroy(ctx->ev_poll);
    for (cudaEvent_t ev : ctx->wf_timing.pool) cudaEventDestroy(ev);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return B200RT_OK;
}

int b200rt_device_info(const b200rt_ctx* ctx, int* sm_count, int* sm_clock_khz, size_t* hbm_bytes) {
    if (!ctx) return B200RT_ERR_INVALID;
    if (sm_count) *sm_count = ctx->sm_count;
    if (sm_clock_khz) *sm_clock_khz = ctx->sm_clock_khz;
    if (hbm_bytes) *hbm_bytes = ctx->hbm_bytes;
    return B200RT_OK;
}

int b200rt_upload_scene(b200rt_ctx* ctx, const b200rt_scene* s) {
    if (!ctx || !s) return B200RT_ERR_INVALID;
    if ((s->n_triangles && !s->triangles) || (s->n_spheres && !s->spheres) || (s->n_materials && !s->materials) ||
        (s->n_lights && !s->lights))
        return B200RT_ERR_INVALID;
    // limits of the packed device formats: primitive id + 1 in 28 bits (ray meta), object id in 24 bits (hit meta),
    // first light of a shadow chunk in 12 bits (path flags)
    if ((uint64_t)s->n_triangles + s->n_spheres >= (1ull << 28) || s->n_materials >= (1u << 24) || s->n_lights >= 4096u)
        return B200RT_ERR_UNSUPPORTED;
    for (uint32_t i = 0; i < s->n_triangles; ++i)
        if (s->triangles[i].object_index >= s->n_materials) return B200RT_ERR_INVALID;
    for (uint32_t i = 0; i < s->n_spheres; ++i)
        if (s->spheres[i].object_index >= s->n_materials) return B200RT_ERR_INVALID;
    for (uint32_t k = 0; k < s->n_materials; ++k) {
        const b200rt_material& m = s->materials[k];
        if (m.kind > B200RT_MATERIAL_GENERATIVE) return B200RT_ERR_INVALID;
        if (m.kind == B200RT_MATERIAL_GENERATIVE && (m.diffuse_fn > B200RT_DIFFUSE_CHECKER_UPV || m.normal_fn > B200RT_NORMAL_SINCOS_U))
            return B200RT_ERR_INVALID;
    }
    for (uint32_t k = 0; k < s->n_lights; ++k)
        if (s->lights[k].kind > B200RT_LIGHT_POINT) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));

    const uint32_t nt = s->n_triangles, ns = s->n_spheres, nm = s->n_materials, nl = s->n_lights;
    const uint32_t nt_pad = ((nt + kTileTris - 1) / kTileTris) * kTileTris;

    // ---- pack (host) ------------------------------------------------------------------------------
    std::vector<float4> tri_exact(4 * (size_t)std::max(nt, 1u)), tri_attr(4 * (size_t)std::max(nt, 1u));
    std::vector<float4> sph(std::max(ns, 1u));
    std::vector<uint32_t> sph_obj(std::max(ns, 1u));
    std::vector<DMaterial> mats(std::max(nm, 1u));
    std::vector<DLight> lights(std::max(nl, 1u));
    for (uint32_t i = 0; i < nt; ++i) {
        const b200rt_triangle& t = s->triangles[i];
        exact_record(t, &tri_exact[4 * (size_t)i]);
        const b200rt_vertex* v = t.vertices;
        tri_attr[4 * (size_t)i + 0] = make_float4(v[0].normal[0], v[0].normal[1], v[0].normal[2], v[0].uv[0]);
        tri_attr[4 * (size_t)i + 1] = make_float4(v[1].normal[0], v[1].normal[1], v[1].normal[2], v[0].uv[1]);
        tri_attr[4 * (size_t)i + 2] = make_float4(v[2].normal[0], v[2].normal[1], v[2].normal[2], v[1].uv[0]);
        tri_attr[4 * (size_t)i + 3] = make_float4(v[1].uv[1], v[2].uv[0], v[2].uv[1], 0.0f);
    }
    for (uint32_t j = 0; j < ns; ++j) {
        sph[j] = make_float4(s->spheres[j].center[0], s->spheres[j].center[1], s->spheres[j].center[2], s->spheres[j].radius);
        sph_obj[j] = s->spheres[j].object_index;
    }
    for (uint32_t k = 0; k < nm; ++k) {
        const b200rt_material& m = s->materials[k];
        DMaterial& o = mats[k];
        std::memset(&o, 0, sizeof o);
        std::memcpy(o.normal, m.normal, 12);
        std::memcpy(o.diffuse, m.diffuse_color, 12);
        o.shiness = m.shiness;
        std::memcpy(o.specular, m.specular_color, 12);
        o.smoothness = m.smoothness;
        o.transparency = m.transparency;
        o.refraction_index = m.refraction_index;
        o.opaque_decay = m.opaque_decay;
        o.kind = m.kind; o.diffuse_fn = m.diffuse_fn; o.normal_fn = m.normal_fn;
        std::memcpy(o.fn_params, m.fn_params, sizeof o.fn_params);
    }
    for (uint32_t k = 0; k < nl; ++k) {
        const b200rt_light& l = s->lights[k];
        DLight& o = lights[k];
        std::memset(&o, 0, sizeof o);
        o.kind = l.kind; o.has_origin = l.has_origin;
        std::memcpy(o.origin, l.origin, 12);
        std::memcpy(o.direction, l.direction, 12);
        o.angle = l.angle; o.softness = l.softness;
        std::memcpy(o.color, l.color, 12);
    }

    // Shadow rays of a directional light (main.rs:423-431: direction = -light.direction, back faces only) have the same
    // direction wherever they start, so the face test of main.rs:184-188 is a property of the (light, triangle) pair:
    // bf = n . (-direction) > 0 in the reference's own arithmetic; !bf triangles are culled for every such ray.  The
    // cast kernels drop them from the candidate sets (this includes triangles exactly parallel to the light, which the
    // conservative filter has to keep for every ray).  Only for scenes of one light chunk (slot s <-> light s).
    const uint32_t n_tiles_sc = nt_pad / kTileTris;
    std::vector<uint2> shadow_cull((size_t)std::max(nl <= 4u ? nl : 0u, 1u) * std::max(n_tiles_sc, 1u), make_uint2(0u, 0u));
    if (nl <= 4u) {
        for (uint32_t k = 0; k < nl; ++k) {
            if (s->lights[k].kind != B200RT_LIGHT_DIRECTIONAL) continue;
            const V3 rd = -v3(s->lights[k].direction);
            for (uint32_t i = 0; i < nt; ++i) {
                const float4 q0 = tri_exact[4 * (size_t)i];
                const bool bf = dot(v3(q0.x, q0.y, q0.z), rd) > 0.0f;          // primitives.rs:45
                if (!bf) {
                    uint2& m = shadow_cull[(size_t)k * n_tiles_sc + i / kTileTris];
                    const uint32_t bit = i % kTileTris;
                    if (bit < 32u) m.x |= 1u << bit; else m.y |= 1u << (bit - 32u);
                }
            }
        }
    }

    // ---- one device blob, 256-byte aligned sections --------------------------------------------------
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    const size_t off_exact = off;  off = align(off + tri_exact.size() * sizeof(float4));
    const size_t off_attr = off;   off = align(off + tri_attr.size() * sizeof(float4));
    const size_t off_sph = off;    off = align(off + sph.size() * sizeof(float4));
    const size_t off_sobj = off;   off = align(off + sph_obj.size() * sizeof(uint32_t));
    const size_t off_mat = off;    off = align(off + mats.size() * sizeof(DMaterial));
    const size_t off_light = off;  off = align(off + lights.size() * sizeof(DLight));
    const size_t off_cull = off;   off = align(off + shadow_cull.size() * sizeof(uint2));
    const size_t total = off;
    std::vector<unsigned char> blob(total, 0);
    std::memcpy(blob.data() + off_exact, tri_exact.data(), tri_exact.size() * sizeof(float4));
    std::memcpy(blob.data() + off_attr, tri_attr.data(), tri_attr.size() * sizeof(float4));
    std::memcpy(blob.data() + off_sph, sph.data(), sph.size() * sizeof(float4));
    std::memcpy(blob.data() + off_sobj, sph_obj.data(), sph_obj.size() * sizeof(uint32_t));
    std::memcpy(blob.data() + off_mat, mats.data(), mats.size() * sizeof(DMaterial));
    std::memcpy(blob.data() + off_light, lights.data(), lights.size() * sizeof(DLight));
    std::memcpy(blob.data() + off_cull, shadow_cull.data(), shadow_cull.size() * sizeof(uint2));

    ctx->have_scene = false;
    int rc = ensure(ctx, &ctx->d_scene_blob, &ctx->scene_blob_bytes, total);
    if (rc != B200RT_OK) return rc;
    CU(cudaMemcpyAsync(ctx->d_scene_blob, blob.data(), total, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    unsigned char* base = static_cast<unsigned char*>(ctx->d_scene_blob);
    DScene& d = ctx->scene;
    d.tri_filter = nullptr;
    d.tri_filter_plain = nullptr;
    d.tri_exact = reinterpret_cast<const float4*>(base + off_exact);
    d.tri_attr = reinterpret_cast<const float4*>(base + off_attr);
    d.sph = reinterpret_cast<const float4*>(base + off_sph);
    d.sph_obj = reinterpret_cast<const uint32_t*>(base + off_sobj);
    d.materials = reinterpret_cast<const DMaterial*>(base + off_mat);
    d.lights = reinterpret_cast<const DLight*>(base + off_light);
    d.shadow_cull = nl <= 4u && nl > 0u ? reinterpret_cast<const uint2*>(base + off_cull) : nullptr;
    d.n_tris = nt; d.n_sph = ns; d.n_lights = nl; d.n_materials = nm;
    d.n_tris_padded = nt_pad;
    // bounds for the filter slack
    double rad = 0.0, edge = 0.0;
    for (uint32_t i = 0; i < nt; ++i) {
        const float4* ex = &tri_exact[4 * (size_t)i];
        for (int k = 1; k <= 3; ++k) {
            const double r2 = (double)ex[k].x * ex[k].x + (double)ex[k].y * ex[k].y + (double)ex[k].z * ex[k].z;
            if (std::isfinite(r2)) rad = std::max(rad, std::sqrt(r2));
            const float4 a = ex[k], b = ex[k == 3 ? 1 : k + 1];
            const double e2 = ((double)a.x - b.x) * ((double)a.x - b.x) + ((double)a.y - b.y) * ((double)a.y - b.y) +
                              ((double)a.z - b.z) * ((double)a.z - b.z);
            if (std::isfinite(e2)) edge = std::max(edge, std::sqrt(e2));
        }
    }
    for (uint32_t j = 0; j < ns; ++j) {
        const double r2 = (double)sph[j].x * sph[j].x + (double)sph[j].y * sph[j].y + (double)sph[j].z * sph[j].z;
        if (std::isfinite(r2) && std::isfinite(sph[j].w)) rad = std::max(rad, std::sqrt(r2) + std::fabs((double)sph[j].w));
    }
    ctx->scene_radius = (float)rad;
    ctx->max_edge = (float)edge;
    ctx->h_tri_exact = tri_exact;
    ctx->h_sph = sph;
    // every secondary ray starts on a primitive, i.e. within scene_radius (+ rounding); cameras further out re-pack
    rc = pack_filter(ctx, (float)(4.0 * rad + 8.0));
    if (rc != B200RT_OK) return rc;
    rc = build_and_upload_bvh(ctx);
    if (rc != B200RT_OK) return rc;
    ctx->have_scene = true;
    return B200RT_OK;
}

int b200rt_render_whitted_device(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                                 float* d_out_rgb, int32_t* d_out_prim_id, void* cuda_stream) {
    if (!ctx || !cam || !params || !d_out_rgb) return B200RT_ERR_INVALID;
    if (!ctx->have_scene) return B200RT_ERR_NO_SCENE;
    DCamera dc;
    DParams dp;
    make_camera(*cam, dc);
    int rc = make_params(*params, 0, 1, dp);
    if (rc != B200RT_OK) return rc;
    CU(cudaSetDevice(ctx->device));
    rc = ensure_origin_bound(ctx, *cam, *params);
    if (rc != B200RT_OK) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;  // NULL = the CUDA default stream
    CU(cudaEventRecord(ctx->ev0, st));
    CU(launch_whitted(ctx->scene, dc, dp, d_out_rgb, d_out_prim_id, ctx->d_cnt, st));
    CU(cudaEventRecord(ctx->ev1, st));
    return B200RT_OK;
}

int b200rt_render_whitted(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params, float* out_rgb,
                          int32_t* out_prim_id) {
    if (!ctx || !cam || !params || !out_rgb) return B200RT_ERR_INVALID;
    if (!ctx->have_scene) return B200RT_ERR_NO_SCENE;
    CU(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)params->width * params->height;
    int rc = ensure(ctx, &ctx->d_out, &ctx->d_out_bytes, npx * 3 * sizeof(float));
    if (rc != B200RT_OK) return rc;
    if (out_prim_id) {
        rc = ensure(ctx, &ctx->d_aux, &ctx->d_aux_bytes, npx * sizeof(int32_t));
        if (rc != B200RT_OK) return rc;
    }
    rc = b200rt_render_whitted_device(ctx, cam, params, (float*)ctx->d_out, out_prim_id ? (int32_t*)ctx->d_aux : nullptr,
                                      ctx->stream);
    if (rc != B200RT_OK) return rc;
    // only the rendered rows are defined; copy exactly those
    const uint32_t r0 = params->row_count ? params->row_begin : 0u;
    const uint32_t rn = params->row_count ? params->row_count : params->height;
    const size_t o = (size_t)r0 * params->width, cnt = (size_t)rn * params->width;
    CU(cudaEventRecord(ctx->ev2, ctx->stream));
    CU(cudaMemcpyAsync(out_rgb + 3 * o, (float*)ctx->d_out + 3 * o, cnt * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_prim_id)
        CU(cudaMemcpyAsync(out_prim_id + o, (int32_t*)ctx->d_aux + o, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev3, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.kernel_ms, ctx->ev0, ctx->ev1);
    cudaEventElapsedTime(&ctx->stats.d2h_ms, ctx->ev2, ctx->ev3);
    ctx->stats.h2d_ms = 0.0f;
    return B200RT_OK;
}

static int render_distributed_device_impl(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                                          uint32_t epoch_begin, uint32_t epoch_count, float* d_accum, void* cuda_stream,
                                          uint32_t strip_rows, uint32_t strip_parts, uint32_t strip_part);

int b200rt_render_distributed_device(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                                     uint32_t epoch_begin, uint32_t epoch_count, float* d_accum, void* cuda_stream) {
    return render_distributed_device_impl(ctx, cam, params, epoch_begin, epoch_count, d_accum, cuda_stream, 0u, 1u, 0u);
}

int b200rt_render_distributed_strips_device(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                                            uint32_t epoch_begin, uint32_t epoch_count, float* d_accum, void* cuda_stream,
                                            uint32_t strip_rows, uint32_t n_parts, uint32_t part) {
    if (strip_rows == 0u || n_parts == 0u || part >= n_parts) return B200RT_ERR_INVALID;
    if (params && (params->tracer == B200RT_TRACER_MEGAKERNEL || params->cast_mode == B200RT_CAST_BRUTE_EXACT)) return B200RT_ERR_UNSUPPORTED;
    return render_distributed_device_impl(ctx, cam, params, epoch_begin, epoch_count, d_accum, cuda_stream, strip_rows, n_parts, part);
}

static int render_distributed_device_impl(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                                          uint32_t epoch_begin, uint32_t epoch_count, float* d_accum, void* cuda_stream,
                                          uint32_t strip_rows, uint32_t strip_parts, uint32_t strip_part) {
    if (!ctx || !cam || !params || !d_accum) return B200RT_ERR_INVALID;
    if (!ctx->have_scene) return B200RT_ERR_NO_SCENE;
    if (((uintptr_t)d_accum & 15u) != 0) return B200RT_ERR_INVALID;  // float4 accumulators
    DCamera dc;
    DParams dp;
    make_camera(*cam, dc);
    int rc = make_params(*params, epoch_begin, epoch_count, dp);
    if (rc != B200RT_OK) return rc;
    CU(cudaSetDevice(ctx->device));
    rc = ensure_origin_bound(ctx, *cam, *params);
    if (rc != B200RT_OK) return rc;
    if (strip_parts > 1u) {
        // this launch owns the strips s of the band with s % strip_parts == strip_part: count its rows
        const uint32_t R = dp.row_count;
        uint32_t mine = 0u;
        for (uint32_t s0 = strip_part * strip_rows; s0 < R; s0 += strip_parts * strip_rows) mine += std::min(strip_rows, R - s0);
        dp.strip_rows = strip_rows; dp.strip_parts = strip_parts; dp.strip_part = strip_part;
        dp.row_count = mine;                 // local rows; frame rows through wf_frame_row
    }
    cudaStream_t st = (cudaStream_t)cuda_stream;  // NULL = the CUDA default stream
    CU(cudaEventRecord(ctx->ev0, st));
    ctx->last_rounds = 0;
    ctx->last_launches = 0;
    CU(cudaMemsetAsync(&ctx->d_cnt->rounds, 0, 2 * sizeof(unsigned long long), st));   // the device loop's counters are per call
    ctx->wf_timing.cast_ms = ctx->wf_timing.logic_ms = ctx->wf_timing.primary_ms = 0.0;
    ctx->wf_timing.cast_launches = 0;
    if (epoch_count) {
        // (the wavefront's control row packs the pixel in 16 + 16 bits: frames beyond 65535 take the megakernel)
        if (params->tracer == B200RT_TRACER_MEGAKERNEL || params->cast_mode == B200RT_CAST_BRUTE_EXACT || dp.width > 65535u || dp.height > 65535u) {
            CU(launch_distributed(ctx->scene, dc, dp, d_accum, ctx->d_cnt, st));
            ctx->last_launches = 1;
        } else {
            // wavefront: the epochs are rendered in batches of `epar` epochs, every epoch of a batch in flight at once
            // (one path slot per pixel and epoch), and - only for frames too large for that - in bands of rows.
            // Both are functions of the frame and the epoch range alone, never of the row band a caller renders,
            // so a band is bitwise the same rows of the full frame.
            const uint32_t epar = wf_epochs_in_flight(dp.width, dp.height, epoch_count, ctx->hbm_bytes);
            const uint64_t max_paths_mem = std::max<uint64_t>(1, (uint64_t)(0.4 * (double)ctx->hbm_bytes) / wf_workspace_bytes_per_path());
            uint32_t band_rows = (uint32_t)std::max<uint64_t>(1, max_paths_mem / ((uint64_t)dp.width * epar));
            band_rows = std::min(band_rows, dp.row_count);
            const uint32_t max_paths = band_rows * dp.width * epar;
            rc = ensure(ctx, &ctx->d_wf, &ctx->d_wf_bytes, wf_workspace_bytes(max_paths));
            if (rc != B200RT_OK) return rc;
            ctx->last_rounds = 0;
            for (uint32_t e0 = 0; e0 < epoch_count; e0 += epar) {
                const uint32_t en = std::min(epar, epoch_count - e0);
                for (uint32_t r = 0; r < dp.row_count; r += band_rows) {
                    DParams band = dp;
                    if (dp.strip_parts > 1u) band.local_row0 = r;          // (strips: sub-bands of the launch's LOCAL rows)
                    else band.row_begin = dp.row_begin + r;
                    band.row_count = std::min(band_rows, dp.row_count - r);
                    band.epoch_begin = epoch_begin + e0;
                    band.epoch_count = en;
                    uint32_t rounds = 0;
                    CU(launch_distributed_wavefront(ctx->scene, dc, band, d_accum, ctx->d_cnt, ctx->d_wf,
                                                    band.row_count * dp.width * en, en, ctx->sm_count, ctx->h_poll,
                                                    ctx->ev_poll, st, &rounds, &ctx->last_launches, ctx->kernel_timing ? &ctx->wf_timing : nullptr));
                    ctx->last_rounds += rounds;
                }
            }
        }
    }
    CU(cudaEventRecord(ctx->ev1, st));
    return B200RT_OK;
}

int b200rt_render_distributed(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                              uint32_t epoch_begin, uint32_t epoch_count, float* accum) {
    if (!ctx || !cam || !params || !accum) return B200RT_ERR_INVALID;
    if (!ctx->have_scene) return B200RT_ERR_NO_SCENE;
    CU(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)params->width * params->height;
    int rc = ensure(ctx, &ctx->d_out, &ctx->d_out_bytes, npx * 4 * sizeof(float));
    if (rc != B200RT_OK) return rc;
    const uint32_t r0 = params->row_count ? params->row_begin : 0u;
    const uint32_t rn = params->row_count ? params->row_count : params->height;
    if (params->row_count && (r0 >= params->height || rn > params->height - r0)) return B200RT_ERR_INVALID;
    const size_t o = (size_t)r0 * params->width, cnt = (size_t)rn * params->width;
    // the accumulation buffer is ADDED to: it travels to the device and back
    CU(cudaEventRecord(ctx->ev2, ctx->stream));
    CU(cudaMemcpyAsync((float*)ctx->d_out + 4 * o, accum + 4 * o, cnt * 4 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->ev3, ctx->stream));
    rc = b200rt_render_distributed_device(ctx, cam, params, epoch_begin, epoch_count, (float*)ctx->d_out, ctx->stream);
    if (rc != B200RT_OK) return rc;
    CU(cudaMemcpyAsync(accum + 4 * o, (float*)ctx->d_out + 4 * o, cnt * 4 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.kernel_ms, ctx->ev0, ctx->ev1);
    cudaEventElapsedTime(&ctx->stats.h2d_ms, ctx->ev2, ctx->ev3);
    return B200RT_OK;
}

int b200rt_resolve_device(b200rt_ctx* ctx, const float* d_accum, float* d_out_rgb, size_t n_pixels, void* cuda_stream) {
    if (!ctx || !d_accum || !d_out_rgb) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;  // NULL = the CUDA default stream
    CU(launch_resolve(d_accum, d_out_rgb, n_pixels, st));
    return B200RT_OK;
}

int b200rt_post_process_device(b200rt_ctx* ctx, float* d_rgb, size_t n_pixels, float* d_p98_out, void* cuda_stream) {
    if (!ctx || (n_pixels && !d_rgb)) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(launch_post_process(d_rgb, n_pixels, ctx->d_pp, d_p98_out, ctx->sm_count, (cudaStream_t)cuda_stream));
    return B200RT_OK;
}

int b200rt_post_process(b200rt_ctx* ctx, float* rgb, size_t n_pixels, float* p98_out) {
    if (!ctx || (n_pixels && !rgb)) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure(ctx, &ctx->d_out, &ctx->d_out_bytes, std::max<size_t>(n_pixels, 1) * 3 * sizeof(float));
    if (rc != B200RT_OK) return rc;
    float* d_p98 = reinterpret_cast<float*>(static_cast<unsigned char*>(ctx->d_pp) + post_process_workspace_bytes());
    CU(cudaMemcpyAsync(ctx->d_out, rgb, n_pixels * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    rc = b200rt_post_process_device(ctx, (float*)ctx->d_out, n_pixels, d_p98, ctx->stream);
    if (rc != B200RT_OK) return rc;
    CU(cudaMemcpyAsync(rgb, ctx->d_out, n_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    float p98 = 0.0f;
    CU(cudaMemcpyAsync(&p98, d_p98, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (p98_out) *p98_out = p98;
    return B200RT_OK;
}

int b200rt_encode_srgb8_device(b200rt_ctx* ctx, const float* d_rgb, size_t n_values, uint8_t* d_out, void* cuda_stream) {
    if (!ctx || (n_values && (!d_rgb || !d_out))) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(launch_encode_srgb8(d_rgb, n_values, d_out, ctx->sm_count, (cudaStream_t)cuda_stream));
    return B200RT_OK;
}

int b200rt_encode_srgb8(b200rt_ctx* ctx, const float* rgb, size_t n_values, uint8_t* out) {
    if (!ctx || (n_values && (!rgb || !out))) return B200RT_ERR_INVALID;
    if (n_values == 0) return B200RT_OK;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure(ctx, &ctx->d_out, &ctx->d_out_bytes, n_values * sizeof(float));
    if (rc != B200RT_OK) return rc;
    rc = ensure(ctx, &ctx->d_aux, &ctx->d_aux_bytes, n_values);
    if (rc != B200RT_OK) return rc;
    CU(cudaMemcpyAsync(ctx->d_out, rgb, n_values * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    rc = b200rt_encode_srgb8_device(ctx, (const float*)ctx->d_out, n_values, (uint8_t*)ctx->d_aux, ctx->stream);
    if (rc != B200RT_OK) return rc;
    CU(cudaMemcpyAsync(out, ctx->d_aux, n_values, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200RT_OK;
}

int b200rt_intersect_device(b200rt_ctx* ctx, const b200rt_ray* d_rays, size_t n, uint32_t cast_mode, b200rt_hit* d_hits,
                            void* cuda_stream) {
    if (!ctx || (n && (!d_rays || !d_hits))) return B200RT_ERR_INVALID;
    if (!ctx->have_scene) return B200RT_ERR_NO_SCENE;
    if (cast_mode > B200RT_CAST_BVH) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;  // NULL = the CUDA default stream
    CU(cudaEventRecord(ctx->ev0, st));
    CU(launch_intersect(ctx->scene, d_rays, n, cast_mode, d_hits, ctx->d_cnt, st));
    CU(cudaEventRecord(ctx->ev1, st));
    return B200RT_OK;
}

int b200rt_intersect(b200rt_ctx* ctx, const b200rt_ray* rays, size_t n, uint32_t cast_mode, b200rt_hit* hits) {
    if (!ctx || (n && (!rays || !hits))) return B200RT_ERR_INVALID;
    if (!ctx->have_scene) return B200RT_ERR_NO_SCENE;
    if (n == 0) return B200RT_OK;
    if (n > 0xffffffffull) return B200RT_ERR_UNSUPPORTED;
    // FaceDirection is an enum in the reference (main.rs:52-57) and the exclusion an Option<PrimitiveIndex> (main.rs:69-75):
    // reject what they cannot hold.  An exclusion index beyond the scene's primitives is legal (it never matches).
    for (size_t i = 0; i < n; ++i)
        if (rays[i].face_direction > B200RT_FACE_BOTH || rays[i].exclude_face > B200RT_FACE_BOTH ||
            rays[i].exclude_prim < -1 || rays[i].exclude_prim >= (1 << 28) - 1)
            return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure(ctx, &ctx->d_aux, &ctx->d_aux_bytes, n * sizeof(b200rt_ray));
    if (rc != B200RT_OK) return rc;
    rc = ensure(ctx, &ctx->d_out, &ctx->d_out_bytes, n * sizeof(b200rt_hit));
    if (rc != B200RT_OK) return rc;
    CU(cudaMemcpyAsync(ctx->d_aux, rays, n * sizeof(b200rt_ray), cudaMemcpyHostToDevice, ctx->stream));
    rc = b200rt_intersect_device(ctx, (const b200rt_ray*)ctx->d_aux, n, cast_mode, (b200rt_hit*)ctx->d_out, ctx->stream);
    if (rc != B200RT_OK) return rc;
    CU(cudaMemcpyAsync(hits, ctx->d_out, n * sizeof(b200rt_hit), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.kernel_ms, ctx->ev0, ctx->ev1);
    return B200RT_OK;
}

int b200rt_get_stats(b200rt_ctx* ctx, b200rt_stats* out) {
    if (!ctx || !out) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = fetch_counters(ctx);
    if (rc != B200RT_OK) return rc;
    // kernel_ms of a *_device call: the events were recorded on the caller's stream
    if (cudaEventQuery(ctx->ev1) == cudaSuccess) cudaEventElapsedTime(&ctx->stats.kernel_ms, ctx->ev0, ctx->ev1);
    *out = ctx->stats;
    return B200RT_OK;
}

int b200rt_set_kernel_timing(b200rt_ctx* ctx, int enabled) {
    if (!ctx) return B200RT_ERR_INVALID;
    ctx->kernel_timing = enabled != 0;
    return B200RT_OK;
}

int b200rt_reset_stats(b200rt_ctx* ctx) {
    if (!ctx) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(DCounters), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::memset(&ctx->stats, 0, sizeof ctx->stats);
    return B200RT_OK;
}

// dev / test entry (b200rt_dev.h): the acceleration structure of a scene, built on the host exactly as
// b200rt_upload_scene builds it (no GPU needed)
int b200rt_dev_build_bvh(const b200rt_scene* s, int which, float* nodes_out, uint32_t max_nodes, uint32_t* tri_index_out,
                         uint32_t* n_nodes, uint32_t* n_indexed, uint32_t* depth, uint32_t* n_leaves) {
    if (!s || !n_nodes || (s->n_triangles && !s->triangles) || which < 0 || which > 1) return B200RT_ERR_INVALID;
    std::vector<float4> ex(4 * (size_t)std::max(s->n_triangles, 1u));
    for (uint32_t i = 0; i < s->n_triangles; ++i) exact_record(s->triangles[i], &ex[4 * (size_t)i]);
    BvhBuild bvh;
    build_bvh(ex.data(), s->n_triangles, bvh);
    const std::vector<float4>& nodes = which == 0 ? bvh.nodes : bvh.nnodes;
    const std::vector<uint32_t>& index = which == 0 ? bvh.tri_index : bvh.ntri_index;
    const size_t per = which == 0 ? 3 : 2;
    *n_nodes = (uint32_t)(nodes.size() / per);
    if (n_indexed) *n_indexed = (uint32_t)index.size();
    if (depth) *depth = which == 0 ? bvh.max_depth : bvh.nmax_depth;
    if (n_leaves) *n_leaves = which == 0 ? bvh.n_leaves : bvh.nn_leaves;
    if (nodes_out) {
        if (max_nodes < *n_nodes) return B200RT_ERR_INVALID;
        if (!nodes.empty()) std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(float4));
    }
    if (tri_index_out && !index.empty()) std::memcpy(tri_index_out, index.data(), index.size() * sizeof(uint32_t));
    return B200RT_OK;
}

int b200rt_filter_bench(b200rt_ctx* ctx, int variant, int blocks_per_sm, int iters, float* kernel_ms, uint64_t* pair_tests) {
    if (!ctx || !kernel_ms || !pair_tests || blocks_per_sm <= 0 || iters <= 0) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * blocks_per_sm;
    const size_t n_rays = (size_t)blocks * 256 * 4;
    // synthetic tile: an 8x8 patch of small triangles in the z=0 plane; rays start above it pointing down
    std::vector<float4> recs(4 * 64), rays(2 * n_rays);
    for (int i = 0; i < 64; ++i) {
        const float cx = (float)(i % 8) * 0.25f, cy = (float)(i / 8) * 0.25f;
        recs[4 * i + 0] = make_float4(0.f, 0.f, 1.f, 0.f);
        recs[4 * i + 1] = make_float4(1.f, 0.f, 0.f, -cx);
        recs[4 * i + 2] = make_float4(0.f, 1.f, 0.f, -cy);
        recs[4 * i + 3] = make_float4(-0.70710678f, -0.70710678f, 0.f, 0.70710678f * (cx + cy + 0.25f));
    }
    uint32_t s = 12345u;
    auto rnd = [&s]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); };
    for (size_t i = 0; i < n_rays; ++i) {
        float dx = rnd() - 0.5f, dy = rnd() - 0.5f, dz = -1.0f;
        const float inv = 1.0f / std::sqrt(dx * dx + dy * dy + dz * dz);
        rays[2 * i] = make_float4(rnd() * 2.f, rnd() * 2.f, 1.0f + rnd(), 0.f);
        rays[2 * i + 1] = make_float4(dx * inv, dy * inv, dz * inv, 0.f);
    }
    const size_t rec_bytes = recs.size() * sizeof(float4), ray_bytes = rays.size() * sizeof(float4);
    const size_t out_bytes = (size_t)blocks * 256 * sizeof(uint32_t);
    int rc = ensure(ctx, &ctx->d_aux, &ctx->d_aux_bytes, rec_bytes + ray_bytes);
    if (rc != B200RT_OK) return rc;
    rc = ensure(ctx, &ctx->d_out, &ctx->d_out_bytes, out_bytes);
    if (rc != B200RT_OK) return rc;
    float4* d_recs = (float4*)ctx->d_aux;
    float4* d_rays = d_recs + recs.size();
    CU(cudaMemcpyAsync(d_recs, recs.data(), rec_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_rays, rays.data(), ray_bytes, cudaMemcpyHostToDevice, ctx->stream));
    unsigned long long pairs = 0;
    const float A = 2e-6f, B = 4e-6f, g = 2.44140625e-4f;
    CU(launch_filter_bench(variant, d_recs, d_rays, (uint32_t*)ctx->d_out, blocks, iters, A, B, g, ctx->stream, &pairs));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        CU(launch_filter_bench(variant, d_recs, d_rays, (uint32_t*)ctx->d_out, blocks, iters, A, B, g, ctx->stream, &pairs));
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        best = std::min(best, ms);
    }
    *kernel_ms = best;
    *pair_tests = pairs;
    return B200RT_OK;
}

// dev / test entry (b200rt_dev.h): out[i] = color_pow(x[i], e[i]) on the device
int b200rt_dev_color_pow(b200rt_ctx* ctx, const float* x, const float* e, float* out, size_t n) {
    if (!ctx || !x || !e || !out) return B200RT_ERR_INVALID;
    if (n == 0) return B200RT_OK;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure(ctx, &ctx->d_aux, &ctx->d_aux_bytes, 2 * n * sizeof(float));
    if (rc != B200RT_OK) return rc;
    rc = ensure(ctx, &ctx->d_out, &ctx->d_out_bytes, n * sizeof(float));
    if (rc != B200RT_OK) return rc;
    float* d_x = (float*)ctx->d_aux;
    CU(cudaMemcpyAsync(d_x, x, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_x + n, e, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_color_pow(d_x, d_x + n, (float*)ctx->d_out, n, ctx->stream));
    CU(cudaMemcpyAsync(out, ctx->d_out, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200RT_OK;
}

int b200rt_pipe_bench(b200rt_ctx* ctx, int variant, float* kernel_ms, double* inst_per_clk_per_smsp) {
    if (!ctx || !kernel_ms || !inst_per_clk_per_smsp) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, iters = 2048;
    int rc = ensure(ctx, &ctx->d_aux, &ctx->d_aux_bytes, (size_t)blocks * 256 * sizeof(float));
    if (rc != B200RT_OK) return rc;
    CU(launch_pipe_bench(variant, (float*)ctx->d_aux, blocks, iters, ctx->stream));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        CU(launch_pipe_bench(variant, (float*)ctx->d_aux, blocks, iters, ctx->stream));
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        best = std::min(best, ms);
    }
    *kernel_ms = best;
    // 64 "main" instructions per thread per iteration; warps = blocks*8
    const double warp_insts = (double)blocks * 8.0 * iters * 64.0;
    const double clocks = best * 1e-3 * (double)ctx->sm_clock_khz * 1e3;
    *inst_per_clk_per_smsp = warp_insts / clocks / ((double)ctx->sm_count * 4.0);
    return B200RT_OK;
}

int b200rt_measure_fp32_peak(b200rt_ctx* ctx, double* tflops, double* sm_mhz_effective) {
    if (!ctx || !tflops) return B200RT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    const int threads = 256, blocks = ctx->sm_count * 8, iters = 4096;
    int rc = ensure(ctx, &ctx->d_aux, &ctx->d_aux_bytes, (size_t)blocks * threads * sizeof(float));
    if (rc != B200RT_OK) return rc;
    for (int w = 0; w < 2; ++w) CU(launch_fp32_peak((float*)ctx->d_aux, blocks, threads, iters, ctx->stream));
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        CU(launch_fp32_peak((float*)ctx->d_aux, blocks, threads, iters, ctx->stream));
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        best = std::min(best, ms);
    }
    const double fmas = (double)blocks * threads * (double)iters * 64.0;
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    if (sm_mhz_effective) *sm_mhz_effective = (fmas / (best * 1e-3)) / ((double)ctx->sm_count * 128.0) / 1e6;
    return B200RT_OK;
}

}  // extern "C"
