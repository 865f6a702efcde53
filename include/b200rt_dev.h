/*
 * b200rt_dev.h — development micro-benchmarks exported by libb200rt.so.  NOT part of the product ABI (b200rt.h):
 * pipe calibration and filter-loop variants used while tuning the cast kernels (tools/filter_bench.py).
 */
#ifndef B200RT_DEV_H
#define B200RT_DEV_H

#include "b200rt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* K2 micro-benchmark: the ray x triangle filter loop of the two-phase cast in isolation (one 64-triangle
 * shared-memory tile, `iters` passes per ray).  variant 0 = scalar FFMA, 1 = FFMA2 over triangle pairs,
 * 2 / 3 = FFMA2 over ray pairs with 2 / 4 rays per thread.  Returns the kernel time and pair-test count. */
int b200rt_filter_bench(b200rt_ctx* ctx, int variant, int blocks_per_sm, int iters, float* kernel_ms,
                        uint64_t* pair_tests);

/* Pipe calibration loops (dev tool): returns warp-instructions per clock per SM sub-partition at sm_mhz. */
int b200rt_pipe_bench(b200rt_ctx* ctx, int variant, float* kernel_ms, double* inst_per_clk_per_smsp);

/* The acceleration structure of B200RT_CAST_BVH as b200rt_upload_scene builds it, on the host (no GPU;
 * csrc/rt_bvh_build.h).  which = 0: the spatial tree, 12 floats per node {bmin.xyz, rho_geom}{bmax.xyz, -}{u32 left |
 * first, u32 right | 0x80000000 + count, u32 axis, -}; which = 1: the tree over the triangles' unit normals, 8 floats per
 * node {nmin.xyz, u32 left | first}{nmax.xyz, u32 right | 0x80000000 + count}.  tri_index_out receives the n_indexed
 * triangle ids the leaves index (every triangle for the normal tree, the well-shaped ones for the spatial tree).
 * nodes_out / tri_index_out may be NULL. */
int b200rt_dev_build_bvh(const b200rt_scene* scene, int which, float* nodes_out, uint32_t max_nodes, uint32_t* tri_index_out,
                         uint32_t* n_nodes, uint32_t* n_indexed, uint32_t* depth, uint32_t* n_leaves);

/* out[i] = color_pow(x[i], e[i]) on the device (csrc/rt_math.cuh): the power the shading code uses where the result is only
 * ever a colour - Phong lobe, spot cone, opaque decay - so that its error bound is measured, not asserted. */
int b200rt_dev_color_pow(b200rt_ctx* ctx, const float* x, const float* e, float* out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_DEV_H */
