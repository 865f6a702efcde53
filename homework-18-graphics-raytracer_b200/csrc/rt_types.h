// rt_types.h — device-resident scene layout shared by the host packer (b200rt_api.cu) and the
// kernels (rt_kernels.cu).  All records are 16-byte float4 so every load is a 128-bit LDG/LDS.
//
// HBM layout (see DESIGN.md "Data layout"):
//   tri_filter[(tile*8 + k)*32 + lane]   k = 0..7, one float4 each: the packed record PAIR of triangles
//        a = 64*tile + lane and b = a + 32, interleaved {a,b} so each 64-bit half-pair is an FFMA2 operand:
//          k0 {nx_a,nx_b, ny_a,ny_b}  k1 {nz_a,nz_b, d_a,d_b}
//          k2 {m0x, m0y}  k3 {m0z, w0}   k4,k5 edge 1   k6,k7 edge 2          (each entry an {a,b} pair)
//        n = unit face normal, d = n.v0 (the reference's per-pair values, bit-identical); m_k = unit
//        in-plane normal of edge k (n x e_k normalised, computed in f64), w_k = -(m_k.v_k) + B with the
//        conservative slack B folded in.  Read by the warp-transposed FFMA2 filter: 8 coalesced 512 B
//        loads per warp per tile, and not at all after kernel start for scenes of <= 64 triangles.
//   tri_exact[4*i + 0..3]   {n.xyz, d}  {v0.xyz, object}  {v1.xyz, 0}  {v2.xyz, 0}
//        Read only for pairs that survive the filter (exact reference-order confirm).
//   tri_attr[4*i + 0..3]    {n0.xyz, uv0.x} {n1.xyz, uv0.y} {n2.xyz, uv1.x} {uv1.y, uv2.x, uv2.y, 0}
//        Read once per cast, for the winning triangle only (normal / uv interpolation).
//   sph[j]                  {c.xyz, r}     sph_obj[j] = object index
//   tri_filter_plain[4*i + 0..3]   the rays-in-lanes filter records, one triangle after the other (tiles stream through
//        shared memory):  {n.xyz, d} * 2^-108   {m_p.xyz, w_p}   {m_q.xyz, w_q}   {a, b, c, 0}
//        * the plane row is SCALED by 2^-108 (exact): n.dir comes out as nd 2^-108, which is subnormal exactly when
//          |n.dir| < g = 2^-18; rcp.approx.ftz flushes a subnormal input to zero and returns +-inf, every estimate of the
//          pair turns inf / NaN and the pair is kept - the filter's near-parallel guard costs no instruction.  t = num / nd
//          is unchanged (both scaled), the slack A |1/nd| uses filter_As = A 2^-108.
//        * the signed distance to the third (the LONGEST) edge follows from the other two, L_p e_p + L_q e_q + L_r e_r =
//          2 Area for every point: e_r = c - a e_p - b e_q with a = L_p / L_r <= 1, b = L_q / L_r <= 1 (2 packed FMAs
//          instead of 3; slack folded into c).
//   RlTileParam (scenes of one tile)   the same 16 numbers per triangle, multipliers and addends apart, handed to the cast
//        kernels as a KERNEL PARAMETER (constant bank): {n.xyz 2^-108, m_p.x} {m_p.yz, m_q.xy} {m_q.z, a, b, 0} and
//        {d 2^-108, w_p, w_q, c}.  The multipliers reach the FFMA2s through uniform registers (LDCU -> FFMA2 R, R, UR, R):
//        no shared-memory loads in the loop and one vector-register operand less per instruction.  The spare slot is 1.0
//        where a triangle starts a new PLANE RUN: consecutive triangles of one plane (the halves of a square, the fan of a
//        polygon) share the plane terms of the run's first triangle in the loop (b200rt_api.cu, pack_filter).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200rt.h"

// Measured alternatives of the rays-in-lanes filter loop that are NOT built (B200, cast of a 16-epoch 4K batch):
// a slack A|r|(1 + |r|) instead of the near-parallel guard widened the band around every triangle at grazing incidence,
// the nearest candidate then failed its exact test and the certified select fell back ten times as often (96.2 ms vs 90.1).

namespace b200rt {

struct DMaterial {  // ColorMaterial / GenerativeMaterial, materials.rs:20-31, 70-83
    float normal[3];
    float diffuse[3];
    float shiness;
    float specular[3];
    float smoothness;
    float transparency;
    float refraction_index;
    float opaque_decay;
    uint32_t kind, diffuse_fn, normal_fn;
    float fn_params[8];
    float pad;
};

struct DLight {  // lights.rs:6-30
    uint32_t kind, has_origin;
    float origin[3];
    float direction[3];
    float angle, softness;
    float color[3];
    float pad;
};

struct DScene {
    const float4* tri_filter;
    const float4* tri_filter_plain;   // [tri][4]: the same filter records, one triangle after the other (rays-in-lanes cast)
    const float4* tri_exact;
    const float4* tri_attr;
    const float4* sph;
    const uint32_t* sph_obj;
    const DMaterial* materials;
    const DLight* lights;
    // [light][tile] 64-bit masks (scenes of <= 4 lights, else null): triangles that the shadow ray of a DIRECTIONAL light
    // culls whatever its origin — the ray's direction is the light's, so main.rs:185-188 is decided per triangle.
    const uint2* shadow_cull;
    uint32_t n_tris, n_sph, n_lights, n_materials;
    uint32_t n_tris_padded;  // tri_filter is padded to a multiple of kTileTris (padding is masked off)
    float origin_bound;      // filter slack was derived for ray origins with |o| <= origin_bound
    float filter_A;          // slack per unit |1/(n.dir)|   (64u * 2S)
    float filter_As;         // filter_A * 2^-108: the same slack against the scaled plane rows of the rays-in-lanes records
    float filter_As_runs;    // ... for the loop over RlTileParam, whose coplanar runs share the plane of their first triangle
    const struct RlTileParam* h_tile0;   // HOST pointer (launchers only): tile 0 of the scene as a kernel parameter
    float filter_g;          // |n.dir| below this always passes the filter (2^-18)
    float filter_B;          // the absolute slack folded into the filter's edge offsets (128u * 2S)
    float scene_extent;      // V + E: largest vertex norm + longest edge
    // B200RT_CAST_BVH (rt_bvh_build.h / rt_bvh.cuh): the spatial tree (3 float4 per node) and the tree over the normals
    // (2 float4 per node), root = node 0, leaves index bvh_tris / nbvh_tris
    const float4* bvh_nodes;
    const uint32_t* bvh_tris;
    const float4* nbvh_nodes;
    const uint32_t* nbvh_tris;
    uint32_t bvh_n_nodes, nbvh_n_nodes;
};

// Camera::shoot hoisted per frame (main.rs:85-92): computed on the host with the same libm tanf
// and the same non-fused f32 expression order as the reference.
struct DCamera {
    float toward[3];
    float x[3];       // tan(fovy/2) * right
    float y[3];       // tan(fovy/2) * up
    float origin[3];  // center + toward * near
    float center[3];
    float near;
};

struct DParams {
    uint32_t width, height, row_begin, row_count;   // row_count LOCAL rows are rendered, local row L = frame row wf_frame_row(p, L)
    // Rows in STRIPS (wavefront tracer only): the band starting at row_begin is cut into strips of strip_rows rows and this
    // launch owns the strips s with s % strip_parts == strip_part; local_row0 = index of the launch's first local row among
    // the owned rows.  strip_parts = 1: every row, contiguous (strip_rows is then 2^30).
    uint32_t strip_rows, strip_parts, strip_part, local_row0;
    int32_t depth;
    float threshold, refract_max_distance;
    uint32_t tir_retries;
    float focus, blur;
    uint32_t seed_lo, seed_hi;
    uint32_t cast_mode;
    uint32_t epoch_begin, epoch_count;
};

// frame row of local row L of a launch (see DParams)
__host__ __device__ inline uint32_t wf_frame_row(const DParams& p, uint32_t L) {
    const uint32_t l = p.local_row0 + L, si = l / p.strip_rows;
    return p.row_begin + (si * p.strip_parts + p.strip_part) * p.strip_rows + (l - si * p.strip_rows);
}

struct DCounters {  // device-side statistics, one 64-bit atomic per CTA at kernel end
    unsigned long long casts, tri_pairs, sph_pairs, confirms, samples, fallbacks;
    unsigned long long rounds, launches;   // of the wavefront's device-side round loop, since the last render call began
};

constexpr int kTileTris = 64;  // triangles per shared-memory tile (64 x 64 B = 4 KB)
constexpr float kRlPlaneScale = 3.0814879110195774e-33f;   // 2^-108: g * 2^-108 = 2^-126, the smallest normal f32

// tile 0 of a scene as a kernel parameter (see the layout notes above): 4 KB in the constant bank
struct RlTileParam {
    float4 rec[4 * kTileTris];
};

// launchers (rt_kernels.cu)
cudaError_t launch_whitted(const DScene& sc, const DCamera& cam, const DParams& p, float* d_rgb, int32_t* d_prim,
                           DCounters* d_cnt, cudaStream_t stream);
cudaError_t launch_distributed(const DScene& sc, const DCamera& cam, const DParams& p, float* d_accum,
                               DCounters* d_cnt, cudaStream_t stream);
cudaError_t launch_intersect(const DScene& sc, const b200rt_ray* d_rays, size_t n, uint32_t cast_mode,
                             b200rt_hit* d_hits, DCounters* d_cnt, cudaStream_t stream);
cudaError_t launch_resolve(const float* d_accum, float* d_rgb, size_t n_pixels, cudaStream_t stream);
cudaError_t launch_filter_bench(int variant, const float4* d_recs, const float4* d_rays, uint32_t* d_out, int blocks,
                                int iters, float A, float B, float g, cudaStream_t stream, unsigned long long* pairs);
cudaError_t launch_pipe_bench(int variant, float* d_sink, int blocks, int iters, cudaStream_t stream);
cudaError_t launch_fp32_peak(float* d_sink, int blocks, int threads, int iters, cudaStream_t stream);
// rt_image.cu
size_t post_process_workspace_bytes();
cudaError_t launch_post_process(float* d_rgb, size_t n_pixels, void* d_workspace, float* d_p98_out, int sm_count, cudaStream_t stream);
cudaError_t launch_color_pow(const float* d_x, const float* d_e, float* d_out, size_t n, cudaStream_t stream);
cudaError_t launch_encode_srgb8(const float* d_rgb, size_t n_values, uint8_t* d_out, int sm_count, cudaStream_t stream);

}  // namespace b200rt
