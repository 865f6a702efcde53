/*
 * oracle.h — C API of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * The oracle is a line-by-line C++ restatement of the reference's render path (main.rs:83-128,
 * 180-614, 652-698, 748-762; primitives.rs:36-47; materials.rs:33-103; lights.rs:48-93;
 * photon.rs:15-33; image.rs:55-66).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libb200rt.so) never does.
 *
 * Pinning status: the reference ships no numeric golden vectors and cannot be built here (Rust,
 * no toolchain).  The oracle is pinned against report/out_single_epoch.png (the reference's own
 * render of the deterministic pass) through tests/golden/out_single_epoch_probe.json, and against
 * hand-derived known-answer tests.  Stochastic passes use a Philox counter RNG instead of rand 0.5's
 * ISAAC (crate source absent): for those the parity is statistical and "unpinned" by the reference.
 */
#ifndef B200RT_ORACLE_H
#define B200RT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "b200rt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* counters[0]=World::cast calls, [1]=ray x triangle pairs, [2]=ray x sphere pairs, [3]=accepted samples */
int oracle_render_whitted(const b200rt_scene* scene, const b200rt_camera* cam, const b200rt_params* params,
                          float* out_rgb, int32_t* out_prim_id, uint64_t counters[4], int n_threads);
int oracle_render_distributed(const b200rt_scene* scene, const b200rt_camera* cam, const b200rt_params* params,
                              uint32_t epoch_begin, uint32_t epoch_count, float* accum, uint64_t counters[4],
                              int n_threads);
int oracle_intersect(const b200rt_scene* scene, const b200rt_ray* rays, size_t n, b200rt_hit* hits);
/* one stochastic sample (pixel y,x; epoch e), before the is_normal filter */
int oracle_sample_distributed(const b200rt_scene* scene, const b200rt_camera* cam, const b200rt_params* params,
                              uint32_t y, uint32_t x, uint32_t epoch, float rgb[3]);

/* main.rs:748-762 (in place, [n_pixels][3]); returns the p99 luma used (0 if none applied) */
float oracle_post_process(float* rgb, size_t n_pixels);
/* image.rs:55-66: linear -> sRGB u8 */
void oracle_encode_srgb8(const float* rgb, size_t n_values, uint8_t* out);
/* photon.rs:18-21 */
void oracle_resolve(const float* accum, size_t n_pixels, float* out_rgb);

/* unit-level hooks for known-answer tests */
void oracle_camera_shoot(const b200rt_camera* cam, float clip_x, float clip_y, b200rt_ray* out);
int oracle_refract(const float n[3], const float l[3], float k, float out[3]);      /* 1 = Some */
void oracle_from_arc_rotate(const float src[3], const float dst[3], const float v[3], float out[3]);
int oracle_light_approx(const b200rt_light* light, const float position[3], float out_dir[3],
                        float out_color[3], float out_origin[3], int* out_has_origin);  /* 1 = Some */
void oracle_material_approx(const b200rt_material* m, const float uv[2], b200rt_material* out);
void oracle_get_diffuse(const b200rt_material* m, const float normal[3], const float light_dir[3], float out[3]);
void oracle_get_specular(const b200rt_material* m, const float normal[3], const float view_dir[3],
                         const float light_dir[3], float out[3]);
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* first n uniform draws (as f32 in [0,1)) of sample stream (seed, y, x, epoch) */
void oracle_sample_uniforms(uint64_t seed, uint32_t y, uint32_t x, uint32_t epoch, int n, float* out);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
