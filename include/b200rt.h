/*
 * b200rt.h — C ABI of libb200rt.so, the B200-native (sm_100a) render core that replaces the
 * per-pixel ray-trace loop of foriequal0/homework-18-graphics-raytracer.
 *
 * Citations are file:line in the reference crate (src/...).  The reference has no FFI of its
 * own: it is one binary whose render loop is the two rayon blocks main.rs:1090-1109 (Whitted
 * frame) and main.rs:1131-1167 (one stochastic epoch).  Those two blocks are the seam; a Rust
 * `-sys` crate binds exactly the entry points below (see INTEGRATION.md for the binding).
 *
 * Conventions
 *   - every entry point returns int: 0 = B200RT_OK, negative = error (b200rt_strerror()).
 *     The reference panics instead (main.rs:785, 767-775); no exception crosses this boundary.
 *   - caller owns every host buffer; the library owns device memory.  `*_device` variants take
 *     device pointers that the caller (e.g. a torch tensor) owns and enqueue on `cuda_stream`
 *     (a cudaStream_t; NULL = the CUDA default stream) without synchronising.
 *   - one context per host thread and per GPU.  Multi-GPU: a b200rt_group (below) shards a frame over
 *     the GPUs of one process or over one rank per process and reduces / gathers on rank 0 with NCCL;
 *     a caller can also shard by hand through b200rt_params.row_begin/row_count and the
 *     epoch_begin/epoch_count arguments of the single-GPU entry points.
 *   - there is no CPU fallback: without a CUDA device b200rt_create fails with B200RT_ERR_NO_DEVICE.
 *   - primitive ids (reference PrimitiveIndex, primitives.rs:31-34) are int32:
 *       triangle i -> i, sphere j -> n_triangles + j, miss -> -1   (reference iteration order,
 *       main.rs:183 then main.rs:264).
 */
#ifndef B200RT_H
#define B200RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RT_VERSION 2

/* ---- error codes ------------------------------------------------------------------------ */
enum {
    B200RT_OK = 0,
    B200RT_ERR_INVALID = -1,    /* bad argument (NULL pointer, zero size, index out of range) */
    B200RT_ERR_CUDA = -2,       /* a CUDA runtime call failed; see b200rt_last_cuda_error()    */
    B200RT_ERR_NO_SCENE = -3,   /* render/intersect before b200rt_upload_scene                 */
    B200RT_ERR_NO_DEVICE = -4,  /* no usable CUDA device: there is no CPU fallback            */
    B200RT_ERR_IO = -5,         /* OBJ file could not be read / parsed (reference: main.rs:785) */
    B200RT_ERR_UNSUPPORTED = -6,/* e.g. recursion depth above B200RT_MAX_DEPTH                 */
    B200RT_ERR_NCCL = -7        /* device groups: libnccl.so.2 missing or an NCCL call failed; b200rt_group_last_error() */
};

/* ---- scene PODs: 1:1 with the reference's Rust types -------------------------------------- */

/* FaceDirection, main.rs:52-67 */
enum { B200RT_FACE_FRONT = 0, B200RT_FACE_BACK = 1, B200RT_FACE_BOTH = 2 };

/* geometric.rs:43-47  PositionNormalUV */
typedef struct b200rt_vertex {
    float position[3];
    float normal[3];
    float uv[2];
} b200rt_vertex;

/* primitives.rs:26-29  Triangle<PositionNormalUV> */
typedef struct b200rt_triangle {
    b200rt_vertex vertices[3];
    uint32_t object_index;
} b200rt_triangle;

/* primitives.rs:15-24  Sphere { object_index, geometry: SphereGeometry{center, radius} } */
typedef struct b200rt_sphere {
    float center[3];
    float radius;
    uint32_t object_index;
} b200rt_sphere;

/* materials.rs:20-31 ColorMaterial and materials.rs:70-83 GenerativeMaterial.
 * Rust closures cannot cross a C ABI, so the two closure pairs the reference instantiates
 * (main.rs:848-863 and main.rs:1019-1026) are enumerated procedural kinds. */
enum { B200RT_MATERIAL_COLOR = 0, B200RT_MATERIAL_GENERATIVE = 1 };
enum {
    B200RT_DIFFUSE_CONST = 0,       /* diffuse_color                                          */
    B200RT_DIFFUSE_STRIPE_V = 1,    /* ((uv.y*freq) as i32 % 2 == 0) ? c0 : c1  main.rs:848-854 */
    B200RT_DIFFUSE_CHECKER_UPV = 2  /* (((uv.x+uv.y)*freq) as i32 % 2 == 0) ? c0 : c1  main.rs:1019-1025 */
};
enum {
    B200RT_NORMAL_CONST = 0,        /* normal                                                 */
    B200RT_NORMAL_SINCOS_U = 1      /* a = uv.x*nfreq*2*PI; v=(sin a,0,cos a); v.z<=0 ? -v : v  main.rs:855-863 */
};
typedef struct b200rt_material {
    uint32_t kind;           /* B200RT_MATERIAL_*                                             */
    float normal[3];         /* tangent-space normal (ColorMaterial.normal)                    */
    float diffuse_color[3];
    float shiness;
    float specular_color[3];
    float smoothness;
    float transparency;
    float refraction_index;
    float opaque_decay;
    uint32_t diffuse_fn;     /* B200RT_DIFFUSE_*  (GENERATIVE only)                            */
    uint32_t normal_fn;      /* B200RT_NORMAL_*   (GENERATIVE only)                            */
    float fn_params[8];      /* [0]=freq, [1..3]=c0, [4..6]=c1, [7]=nfreq                      */
} b200rt_material;

/* lights.rs:6-30  Directional / Spot / Point */
enum { B200RT_LIGHT_DIRECTIONAL = 0, B200RT_LIGHT_SPOT = 1, B200RT_LIGHT_POINT = 2 };
typedef struct b200rt_light {
    uint32_t kind;
    uint32_t has_origin;     /* Directional.origin: Option<Point3>; Spot/Point always have one */
    float origin[3];
    float direction[3];      /* Directional, Spot                                             */
    float angle;             /* Spot: half-angle in radians (lights.rs:17)                    */
    float softness;          /* Spot                                                          */
    float color[3];
} b200rt_light;

/* World, main.rs:130-137.  objects[i] owns materials[i] (primitives.rs:8-10). */
typedef struct b200rt_scene {
    const b200rt_triangle* triangles;  uint32_t n_triangles;
    const b200rt_sphere*   spheres;    uint32_t n_spheres;
    const b200rt_material* materials;  uint32_t n_materials;   /* one per object */
    const b200rt_light*    lights;     uint32_t n_lights;
} b200rt_scene;

/* Camera, main.rs:43-49 */
typedef struct b200rt_camera {
    float fovy;        /* radians */
    float center[3];
    float toward[3];
    float up[3];
    float near;
} b200rt_camera;

/* Ray + Exclusion, main.rs:69-81 */
typedef struct b200rt_ray {
    float origin[3];
    float direction[3];
    uint32_t face_direction;       /* B200RT_FACE_*                                            */
    int32_t exclude_prim;          /* -1 = None, else primitive id                             */
    uint32_t exclude_face;         /* B200RT_FACE_* of the exclusion                           */
} b200rt_ray;

/* Hit, main.rs:139-147 (object reference replaced by its index) */
typedef struct b200rt_hit {
    int32_t prim_id;               /* -1 = Option::None                                        */
    uint32_t object_index;
    uint32_t face_direction;
    float distance;
    float position[3];
    float normal[3];
    float uv[2];
} b200rt_hit;

/* Every literal of main() that shapes the render loop, as a runtime parameter. */
#define B200RT_MAX_DEPTH 16
enum {
    B200RT_CAST_TWO_PHASE = 0,   /* FMA filter over packed plane records + exact confirm (default) */
    B200RT_CAST_BRUTE_EXACT = 1, /* every pair through the exact reference-order test              */
    /* the same hits (bit for bit: id, face, distance, position, normal, uv) through a bounding-volume
     * hierarchy over the triangles, built at b200rt_upload_scene.  Not the brute-force walk of main.rs:183 any more, so
     * not the workload the FP32-roofline figure is quoted on: a mode for scenes of 1e4+ triangles (rt_bvh.cuh). */
    B200RT_CAST_BVH = 2
};
/* How b200rt_render_distributed* schedules the stochastic tracer on the GPU (same samples, same bits):
 * WAVEFRONT keeps every path's state in HBM and alternates one cast kernel with one shading kernel per
 * round (all lanes of every warp busy); MEGAKERNEL runs each pixel's whole recursion in one thread. */
enum {
    B200RT_TRACER_WAVEFRONT = 0,  /* default */
    B200RT_TRACER_MEGAKERNEL = 1
};
typedef struct b200rt_params {
    uint32_t width, height;        /* main.rs:1084-1085                                        */
    uint32_t row_begin, row_count; /* rows [row_begin,row_begin+row_count) are rendered; 0,0 = all */
    int32_t depth;                 /* main.rs:1098, 1139 (5)                                   */
    float threshold;               /* main.rs:467 (0.001)                                      */
    float refract_max_distance;    /* main.rs:505, 601 (100.0)                                 */
    uint32_t tir_retries;          /* main.rs:378 (10)                                         */
    float focus, blur;             /* main.rs:1147-1148 (3.0, 0.04)                            */
    uint64_t seed;                 /* key of the counter-based sample generator                */
    uint32_t cast_mode;            /* B200RT_CAST_*                                            */
    uint32_t tracer;               /* B200RT_TRACER_* (stochastic tracer only)                 */
} b200rt_params;

typedef struct b200rt_stats {
    uint64_t casts;                /* World::cast calls                                        */
    uint64_t tri_pair_tests;       /* ray x triangle pairs put through the filter / exact test */
    uint64_t sph_pair_tests;       /* ray x sphere pairs                                       */
    uint64_t exact_confirms;       /* pairs that went on to the exact confirm                  */
    uint64_t samples;              /* pixel samples produced (reference "rays", main.rs:1108)  */
    float kernel_ms;               /* device time of the last render/intersect kernel          */
    float h2d_ms, d2h_ms;
    uint32_t wavefront_rounds;     /* cast + shading rounds of the last wavefront render        */
    uint64_t certify_fallbacks;    /* casts whose certified select fell back to the ordered walk */
    /* only with b200rt_set_kernel_timing(ctx, 1): device time of the last wavefront render split by kernel */
    float cast_kernel_ms;          /* sum over the cast launches (World::cast)                    */
    float logic_kernel_ms;         /* sum over the shading / scatter kernels between them         */
    uint32_t cast_kernel_launches;
    uint32_t kernel_launches;      /* kernels launched by the last render call (always counted) */
    float primary_kernel_ms;       /* (kernel timing) the round-0 cast that also generates the camera rays: one primary cast
                                      per pixel sample, NOT part of cast_kernel_ms / cast_kernel_launches */
    uint32_t reserved;
} b200rt_stats;

typedef struct b200rt_ctx b200rt_ctx;

/* ---- context ---------------------------------------------------------------------------- */
int b200rt_create(int device_id, b200rt_ctx** out_ctx);
int b200rt_destroy(b200rt_ctx* ctx);
const char* b200rt_strerror(int code);
const char* b200rt_last_cuda_error(const b200rt_ctx* ctx);
int b200rt_device_info(const b200rt_ctx* ctx, int* sm_count, int* sm_clock_khz, size_t* hbm_bytes);

/* Replaces World::push_* having built the immutable world (main.rs:812-1075): copies the scene,
 * packs the SoA plane / vertex / attribute records and precomputes per-triangle invariants. */
int b200rt_upload_scene(b200rt_ctx* ctx, const b200rt_scene* scene);

/* ---- render entry points ------------------------------------------------------------------ */
/* main.rs:1090-1109: one Whitted frame.  out_rgb = [h][w][3] linear f32 (image.rs:35-40 layout),
 * out_prim_id = [h][w] primary hit ids or NULL.  Only rows of params->row_* are written. */
int b200rt_render_whitted(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                          float* out_rgb, int32_t* out_prim_id);
int b200rt_render_whitted_device(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                                 float* d_out_rgb, int32_t* d_out_prim_id, void* cuda_stream);

/* main.rs:1131-1167: epochs [epoch_begin, epoch_begin+epoch_count) of the thin-lens + scatter
 * tracer, accumulated with PhotonAccumulator semantics (photon.rs:15-33):
 * accum = [h][w][4] = {sum.r, sum.g, sum.b, weight_sum}; samples with a non-normal channel are
 * dropped (main.rs:1157-1160).  The buffer is ADDED to (zero it first for a fresh render). */
int b200rt_render_distributed(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                              uint32_t epoch_begin, uint32_t epoch_count, float* accum);
int b200rt_render_distributed_device(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                                     uint32_t epoch_begin, uint32_t epoch_count, float* d_accum,
                                     void* cuda_stream);
/* The same for an INTERLEAVED share of the rows: the band params->row_* is cut into strips of strip_rows rows and the call
 * renders the strips s with s % n_parts == part (wavefront tracer only).  Strips balance a frame whose cost varies from
 * top to bottom over several GPUs; every pixel sample is what the whole-frame call renders (bitwise). */
int b200rt_render_distributed_strips_device(b200rt_ctx* ctx, const b200rt_camera* cam, const b200rt_params* params,
                                            uint32_t epoch_begin, uint32_t epoch_count, float* d_accum, void* cuda_stream,
                                            uint32_t strip_rows, uint32_t n_parts, uint32_t part);
/* Note: the *_device calls are ASYNCHRONOUS on a stream the caller created: with B200RT_TRACER_WAVEFRONT the rounds of
 * the tracer repeat on the device (a CUDA graph WHILE node that ends when every path has retired), the call enqueues
 * round 0, that graph and the accumulate kernel and returns; B200RT_TRACER_MEGAKERNEL enqueues one kernel and returns.
 * On the legacy default stream (cuda_stream = NULL: it cannot be captured into a graph), and while per-kernel timing
 * is on (b200rt_set_kernel_timing), the wavefront call enqueues its rounds from the host and synchronises with the
 * stream a few times to read the retired-paths counter; on return all but the accumulate kernel have completed. */

/* photon.rs:18-21 into_rgb_internal: out_rgb[i] = weight_sum < EPSILON ? 0 : sum / weight_sum */
int b200rt_resolve_device(b200rt_ctx* ctx, const float* d_accum, float* d_out_rgb, size_t n_pixels,
                          void* cuda_stream);

/* ---- image finishers --------------------------------------------------------- */
/* post_process, main.rs:748-762: divide every channel by the 99th-percentile luma of the image (the element at
 * index (len as f32 * 0.99) of the sorted NORMAL lumas), if it exceeds f32::EPSILON.  In place on [n_pixels][3]
 * linear f32; exact (radix select instead of the sort).  *p98_out = the divisor, 0 if the image was left as is
 * (p98_out may be NULL; for the _device variant it is a device pointer). */
int b200rt_post_process(b200rt_ctx* ctx, float* rgb, size_t n_pixels, float* p98_out);
int b200rt_post_process_device(b200rt_ctx* ctx, float* d_rgb, size_t n_pixels, float* d_p98_out, void* cuda_stream);
/* image.rs:55-66: linear f32 -> sRGB u8 (palette Srgb::from_linear + into_format::<u8>), n_values = 3 * pixels. */
int b200rt_encode_srgb8(b200rt_ctx* ctx, const float* rgb, size_t n_values, uint8_t* out);
int b200rt_encode_srgb8_device(b200rt_ctx* ctx, const float* d_rgb, size_t n_values, uint8_t* d_out, void* cuda_stream);
/* write_to_file, main.rs:764-776 (next row N4): [height][width][3] sRGB u8 -> an RGB8 PNG, written to tmp.png in the
 * target's directory and renamed over `path` (a reader never sees a partial file).  Host-only: no context, no GPU. */
int b200rt_write_png_rgb8(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height);

/* World::cast, main.rs:180-326, for n rays.  face_direction / exclude_face must be B200RT_FACE_* and exclude_prim in
 * [-1, 2^28 - 2] (-1 = no exclusion; an id beyond the scene's primitives never matches, as in the reference): the
 * host-buffer entry returns B200RT_ERR_INVALID otherwise; the _device entry cannot look at the rays and treats an
 * out-of-range face as BOTH and an out-of-range exclusion as none. */
int b200rt_intersect(b200rt_ctx* ctx, const b200rt_ray* rays, size_t n, uint32_t cast_mode,
                     b200rt_hit* hits);
int b200rt_intersect_device(b200rt_ctx* ctx, const b200rt_ray* d_rays, size_t n, uint32_t cast_mode,
                            b200rt_hit* d_hits, void* cuda_stream);

int b200rt_get_stats(b200rt_ctx* ctx, b200rt_stats* out);
/* Measure the wavefront tracer's kernels separately (CUDA events around every cast launch on the launching
 * stream; a few microseconds of host work per round).  Off by default. */
int b200rt_set_kernel_timing(b200rt_ctx* ctx, int enabled);
int b200rt_reset_stats(b200rt_ctx* ctx);

/* FP32-pipe calibration: runs a dependent-free FFMA loop on every SM and returns the measured
 * TFLOP/s (2 flop per FFMA lane) — the live denominator bench.py reports beside the nominal one. */
int b200rt_measure_fp32_peak(b200rt_ctx* ctx, double* tflops, double* sm_mhz_effective);

/* ---- device groups: one frame over several GPUs --------------------------------- */
/* The reference's render loop is host code (main.rs:1086-1173); its pixel samples are independent, so a group shards
 * them without any exchange in the data path:
 *   epochs (stochastic pass, main.rs:1129): rank g renders epochs [g*E/G, (g+1)*E/G) of the full frame into its own
 *          PhotonAccumulator buffer; ONE ncclReduce(sum) to rank 0 over NVLink ends the render;
 *   rows   (Whitted pass, main.rs:1090): rank g renders rows [g*H/G, (g+1)*H/G) of the requested band; the disjoint
 *          bands are gathered on rank 0 (ncclSend / ncclRecv): bitwise the single-GPU frame.
 * A group is every GPU of ONE process (b200rt_group_create: ncclCommInitAll; the library runs one host thread per
 * device while a render is in flight) or ONE RANK per process (b200rt_group_create_rank: ncclCommInitRank with an id
 * from b200rt_group_unique_id on rank 0 that the host program hands to every rank).  Calls are collective: every rank
 * makes the same call with the same camera / params / epoch range.  Results land on rank 0 only; they are FRESH
 * frames (accumulators start at zero on the devices: nothing but the arguments travels host -> device).
 * NCCL (libnccl.so.2) is loaded at run time by the first group call; groups of one GPU do not need it. */
typedef struct b200rt_group b200rt_group;
#define B200RT_GROUP_ID_BYTES 128
int b200rt_group_create(const int* device_ids, int n_devices, b200rt_group** out_group);
int b200rt_group_unique_id(void* id_out, size_t id_bytes);                    /* id_bytes >= B200RT_GROUP_ID_BYTES */
int b200rt_group_create_rank(int device_id, int rank, int n_ranks, const void* unique_id, size_t id_bytes,
                             b200rt_group** out_group);
int b200rt_group_destroy(b200rt_group* group);
int b200rt_group_size(const b200rt_group* group, int* n_ranks, int* n_local);
/* the context of local member `local_index` (stats, device info); owned by the group */
int b200rt_group_ctx(b200rt_group* group, int local_index, b200rt_ctx** out_ctx);
const char* b200rt_group_last_error(const b200rt_group* group);
int b200rt_group_upload_scene(b200rt_group* group, const b200rt_scene* scene);   /* replicated on every member */
/* out_accum = [h][w][4] {sum.rgb, weight_sum} on the process that holds rank 0 (ignored elsewhere, may be NULL). */
int b200rt_group_render_distributed(b200rt_group* group, const b200rt_camera* cam, const b200rt_params* params,
                                    uint32_t epoch_begin, uint32_t epoch_count, float* out_accum);
/* d_accum_root: device buffer on rank 0's GPU that receives the reduced frame (NULL elsewhere); synchronous. */
int b200rt_group_render_distributed_device(b200rt_group* group, const b200rt_camera* cam, const b200rt_params* params,
                                           uint32_t epoch_begin, uint32_t epoch_count, float* d_accum_root);
/* The same frame sharded by ROWS instead (every rank renders all the epochs of its rows): the split for frames of few
 * epochs, e.g. the 10 M one-sample "photons" of a 4000 x 2500 frame.  Rank g takes the 16-row strips
 * s with s % G == g (interleaved: the cost of a frame varies from top to bottom), and one ncclReduce(sum) of the
 * accumulators, whose supports are disjoint, assembles the frame on rank 0: bitwise the single-GPU frame. */
int b200rt_group_render_distributed_rows(b200rt_group* group, const b200rt_camera* cam, const b200rt_params* params,
                                         uint32_t epoch_begin, uint32_t epoch_count, float* out_accum);
int b200rt_group_render_distributed_rows_device(b200rt_group* group, const b200rt_camera* cam, const b200rt_params* params,
                                                uint32_t epoch_begin, uint32_t epoch_count, float* d_accum_root);
/* out_rgb = [h][w][3], out_prim_id = [h][w] (only with want_prim_ids, which every rank passes alike), on rank 0. */
int b200rt_group_render_whitted(b200rt_group* group, const b200rt_camera* cam, const b200rt_params* params,
                                float* out_rgb, int32_t* out_prim_id, int want_prim_ids);
int b200rt_group_render_whitted_device(b200rt_group* group, const b200rt_camera* cam, const b200rt_params* params,
                                       float* d_rgb_root);
/* device time of the last group render on the slowest local member, render + collective, without the D2H copy */
int b200rt_group_last_render_ms(const b200rt_group* group, float* ms);

/* ---- host-side scene construction (World builder; no GPU needed) --------------------------- */
/* Mirrors World::new / push_object / ObjectProxy::push_triangle(s) / push_sphere / push_light
 * (main.rs:161-178, 705-728), triangle()/square() (main.rs:730-746) and load_obj (main.rs:778-807). */
typedef struct b200rt_world b200rt_world;
b200rt_world* b200rt_world_new(void);
void b200rt_world_free(b200rt_world* w);
int b200rt_world_push_object(b200rt_world* w, const b200rt_material* material);       /* -> object index */
int b200rt_world_push_triangle(b200rt_world* w, uint32_t object_index, const b200rt_vertex v[3]);
/* triangle(): flat normal = normalize((v1-v0) x (v2-v1)), main.rs:730-739 */
int b200rt_world_push_flat_triangle(b200rt_world* w, uint32_t object_index,
                                    const float pos[3][3], const float uv[3][2]);
/* square(): two flat triangles (0,1,2),(0,2,3), main.rs:741-746 */
int b200rt_world_push_square(b200rt_world* w, uint32_t object_index,
                             const float pos[4][3], const float uv[4][2]);
int b200rt_world_push_sphere(b200rt_world* w, uint32_t object_index, const float center[3], float radius);
int b200rt_world_push_light(b200rt_world* w, const b200rt_light* light);
/* load_obj(): first model only, flat triangles, uv=(0,0), position = p/scale_div + offset
 * (the reference hard-codes /3.0 + (0.7,1.0,-0.5), main.rs:802). Returns #triangles or <0. */
int b200rt_world_load_obj(b200rt_world* w, uint32_t object_index, const char* path,
                          float scale_div, const float offset[3]);
/* The rest of what tobj::load_obj (tobj 0.1.6, Cargo.toml:14) returns and load_obj drops (main.rs:786-790): the other
 * models of the file (model_index = k, or -1 for all of them in file order; a model is a run of faces between o / g
 * statements) and, with the flags, the file's own `vt` as uv and `vn` as vertex normals for the corners that name them
 * (the others keep uv = (0,0) / the flat normal of triangle(), main.rs:730-739, 797-799).  flags = 0, model_index = 0
 * is b200rt_world_load_obj.  Returns #triangles or <0. */
#define B200RT_OBJ_USE_TEXCOORDS 1u
#define B200RT_OBJ_USE_NORMALS   2u
int b200rt_world_load_obj_ex(b200rt_world* w, uint32_t object_index, const char* path, float scale_div,
                             const float offset[3], int32_t model_index, uint32_t flags);
int b200rt_obj_model_count(const char* path);   /* models in the file (>= 1), or <0 */
int b200rt_world_scene(const b200rt_world* w, b200rt_scene* out);   /* view, valid until next push/free */
/* The scene literal of main() (main.rs:810-1075) and its camera (main.rs:1077-1083).
 * obj_path NULL = use the built-in dodecahedron mesh (same 20 v / 36 f as dodecahedron.obj). */
int b200rt_world_fixture(b200rt_world* w, const char* obj_path);
void b200rt_fixture_camera(b200rt_camera* cam);
void b200rt_default_params(b200rt_params* p);   /* 1280x960, depth 5, 0.001, 100.0, 10, 3.0, 0.04 */

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
