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

#ifdef __cplusplus
}
#endif
#endif /* B200RT_DEV_H */
