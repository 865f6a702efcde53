//! FFI declarations of `libb200rt.so` (`include/b200rt.h`), the B200-native replacement of the two rayon blocks of
//! the reference's `main()` (main.rs:1089-1109 and 1131-1167).  UNCOMPILED SOURCE: the build image has no Rust
//! toolchain; field order, names and types are checked against the C header's ctypes mirror by
//! `tests/test_rust_sys_layout.py`.  Every `#[repr(C)]` struct mirrors one Rust type of the reference 1:1.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const B200RT_OK: c_int = 0;
pub const B200RT_ERR_INVALID: c_int = -1;
pub const B200RT_ERR_CUDA: c_int = -2;
pub const B200RT_ERR_NO_SCENE: c_int = -3;
pub const B200RT_ERR_NO_DEVICE: c_int = -4; // there is no CPU fallback
pub const B200RT_ERR_IO: c_int = -5;
pub const B200RT_ERR_UNSUPPORTED: c_int = -6;
pub const B200RT_ERR_NCCL: c_int = -7; // device groups only

pub const B200RT_FACE_FRONT: u32 = 0; // FaceDirection, main.rs:52-66
pub const B200RT_FACE_BACK: u32 = 1;
pub const B200RT_FACE_BOTH: u32 = 2;
pub const B200RT_MATERIAL_COLOR: u32 = 0; // ColorMaterial, materials.rs:20-31
pub const B200RT_MATERIAL_GENERATIVE: u32 = 1; // GenerativeMaterial, materials.rs:70-83 (closures enumerated)
pub const B200RT_DIFFUSE_CONST: u32 = 0;
pub const B200RT_DIFFUSE_STRIPE_V: u32 = 1; // main.rs:848-854
pub const B200RT_DIFFUSE_CHECKER_UPV: u32 = 2; // main.rs:1019-1025
pub const B200RT_NORMAL_CONST: u32 = 0;
pub const B200RT_NORMAL_SINCOS_U: u32 = 1; // main.rs:855-863
pub const B200RT_LIGHT_DIRECTIONAL: u32 = 0; // lights.rs:6-30
pub const B200RT_LIGHT_SPOT: u32 = 1;
pub const B200RT_LIGHT_POINT: u32 = 2;
pub const B200RT_CAST_TWO_PHASE: u32 = 0;
pub const B200RT_CAST_BRUTE_EXACT: u32 = 1;
pub const B200RT_CAST_BVH: u32 = 2;
pub const B200RT_TRACER_WAVEFRONT: u32 = 0;
pub const B200RT_TRACER_MEGAKERNEL: u32 = 1;
pub const B200RT_OBJ_USE_TEXCOORDS: u32 = 1;
pub const B200RT_OBJ_USE_NORMALS: u32 = 2;

/// `PositionNormalUV`, geometric.rs:43-47
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_vertex {
    pub position: [f32; 3],
    pub normal: [f32; 3],
    pub uv: [f32; 2],
}

/// `Triangle<PositionNormalUV>`, primitives.rs:26-29
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_triangle {
    pub vertices: [b200rt_vertex; 3],
    pub object_index: u32,
}

/// `Sphere`, primitives.rs:15-24
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_sphere {
    pub center: [f32; 3],
    pub radius: f32,
    pub object_index: u32,
}

/// `ColorMaterial` / `GenerativeMaterial`, materials.rs:20-31, 70-83
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_material {
    pub kind: u32,
    pub normal: [f32; 3],
    pub diffuse_color: [f32; 3],
    pub shiness: f32,
    pub specular_color: [f32; 3],
    pub smoothness: f32,
    pub transparency: f32,
    pub refraction_index: f32,
    pub opaque_decay: f32,
    pub diffuse_fn: u32,
    pub normal_fn: u32,
    pub fn_params: [f32; 8],
}

/// `Directional` / `Spot` / `Point`, lights.rs:6-30
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_light {
    pub kind: u32,
    pub has_origin: u32,
    pub origin: [f32; 3],
    pub direction: [f32; 3],
    pub angle: f32,
    pub softness: f32,
    pub color: [f32; 3],
}

/// `World`, main.rs:130-137 (a view: the arrays stay owned by the caller)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_scene {
    pub triangles: *const b200rt_triangle,
    pub n_triangles: u32,
    pub spheres: *const b200rt_sphere,
    pub n_spheres: u32,
    pub materials: *const b200rt_material,
    pub n_materials: u32,
    pub lights: *const b200rt_light,
    pub n_lights: u32,
}

/// `Camera`, main.rs:43-49
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_camera {
    pub fovy: f32,
    pub center: [f32; 3],
    pub toward: [f32; 3],
    pub up: [f32; 3],
    pub near: f32,
}

/// `Ray` + `Exclusion`, main.rs:69-81
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_ray {
    pub origin: [f32; 3],
    pub direction: [f32; 3],
    pub face_direction: u32,
    pub exclude_prim: i32,
    pub exclude_face: u32,
}

/// `Hit`, main.rs:139-147 (`prim_id`: triangle i -> i, sphere j -> n_triangles + j, None -> -1)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_hit {
    pub prim_id: i32,
    pub object_index: u32,
    pub face_direction: u32,
    pub distance: f32,
    pub position: [f32; 3],
    pub normal: [f32; 3],
    pub uv: [f32; 2],
}

/// the literals of `main()`: main.rs:1084-1085, 1098, 467, 505, 378, 1147-1148
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_params {
    pub width: u32,
    pub height: u32,
    pub row_begin: u32,
    pub row_count: u32,
    pub depth: i32,
    pub threshold: f32,
    pub refract_max_distance: f32,
    pub tir_retries: u32,
    pub focus: f32,
    pub blur: f32,
    pub seed: u64,
    pub cast_mode: u32,
    pub tracer: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200rt_stats {
    pub casts: u64,
    pub tri_pair_tests: u64,
    pub sph_pair_tests: u64,
    pub exact_confirms: u64,
    pub samples: u64,
    pub kernel_ms: f32,
    pub h2d_ms: f32,
    pub d2h_ms: f32,
    pub wavefront_rounds: u32,
    pub certify_fallbacks: u64,
    pub cast_kernel_ms: f32,
    pub logic_kernel_ms: f32,
    pub cast_kernel_launches: u32,
    pub kernel_launches: u32,
    pub primary_kernel_ms: f32,
    pub reserved: u32,
}

/// opaque: one context per host thread and GPU
pub enum b200rt_ctx {}
/// opaque: a device group (several GPUs render one frame; NCCL inside the library)
pub enum b200rt_group {}
/// opaque: the host-side `World` builder (main.rs:161-178, 705-746, 778-807)
pub enum b200rt_world {}

extern "C" {
    pub fn b200rt_create(device_id: c_int, out_ctx: *mut *mut b200rt_ctx) -> c_int;
    pub fn b200rt_destroy(ctx: *mut b200rt_ctx) -> c_int;
    pub fn b200rt_strerror(code: c_int) -> *const c_char;
    pub fn b200rt_last_cuda_error(ctx: *const b200rt_ctx) -> *const c_char;
    pub fn b200rt_device_info(ctx: *const b200rt_ctx, sm_count: *mut c_int, sm_clock_khz: *mut c_int, hbm_bytes: *mut usize) -> c_int;
    pub fn b200rt_upload_scene(ctx: *mut b200rt_ctx, scene: *const b200rt_scene) -> c_int;
    /// replaces main.rs:1089-1109; `out_rgb`: [height][width][3] f32, `out_prim_id`: [height][width] i32 or null
    pub fn b200rt_render_whitted(ctx: *mut b200rt_ctx, cam: *const b200rt_camera, params: *const b200rt_params,
                                 out_rgb: *mut f32, out_prim_id: *mut i32) -> c_int;
    pub fn b200rt_render_whitted_device(ctx: *mut b200rt_ctx, cam: *const b200rt_camera, params: *const b200rt_params,
                                        d_out_rgb: *mut f32, d_out_prim_id: *mut i32, cuda_stream: *mut c_void) -> c_int;
    /// replaces main.rs:1131-1167 for epochs [epoch_begin, epoch_begin + epoch_count); `accum`: [height][width][4] f32
    /// {sum.rgb, weight_sum} (photon.rs:9-12), added to
    pub fn b200rt_render_distributed(ctx: *mut b200rt_ctx, cam: *const b200rt_camera, params: *const b200rt_params,
                                     epoch_begin: u32, epoch_count: u32, accum: *mut f32) -> c_int;
    pub fn b200rt_render_distributed_device(ctx: *mut b200rt_ctx, cam: *const b200rt_camera, params: *const b200rt_params,
                                            epoch_begin: u32, epoch_count: u32, d_accum: *mut f32, cuda_stream: *mut c_void) -> c_int;
    pub fn b200rt_render_distributed_strips_device(ctx: *mut b200rt_ctx, cam: *const b200rt_camera, params: *const b200rt_params,
                                                   epoch_begin: u32, epoch_count: u32, d_accum: *mut f32, cuda_stream: *mut c_void,
                                                   strip_rows: u32, n_parts: u32, part: u32) -> c_int;
    pub fn b200rt_resolve_device(ctx: *mut b200rt_ctx, d_accum: *const f32, d_out_rgb: *mut f32, n_pixels: usize, cuda_stream: *mut c_void) -> c_int;
    /// post_process, main.rs:748-762
    pub fn b200rt_post_process(ctx: *mut b200rt_ctx, rgb: *mut f32, n_pixels: usize, p98_out: *mut f32) -> c_int;
    pub fn b200rt_post_process_device(ctx: *mut b200rt_ctx, d_rgb: *mut f32, n_pixels: usize, d_p98_out: *mut f32, cuda_stream: *mut c_void) -> c_int;
    /// image.rs:55-66
    pub fn b200rt_encode_srgb8(ctx: *mut b200rt_ctx, rgb: *const f32, n_values: usize, out: *mut u8) -> c_int;
    pub fn b200rt_encode_srgb8_device(ctx: *mut b200rt_ctx, d_rgb: *const f32, n_values: usize, d_out: *mut u8, cuda_stream: *mut c_void) -> c_int;
    /// write_to_file, main.rs:764-776
    pub fn b200rt_write_png_rgb8(path: *const c_char, rgb: *const u8, width: u32, height: u32) -> c_int;
    /// World::cast, main.rs:180-326
    pub fn b200rt_intersect(ctx: *mut b200rt_ctx, rays: *const b200rt_ray, n: usize, cast_mode: u32, hits: *mut b200rt_hit) -> c_int;
    pub fn b200rt_intersect_device(ctx: *mut b200rt_ctx, d_rays: *const b200rt_ray, n: usize, cast_mode: u32, d_hits: *mut b200rt_hit,
                                   cuda_stream: *mut c_void) -> c_int;
    pub fn b200rt_get_stats(ctx: *mut b200rt_ctx, out: *mut b200rt_stats) -> c_int;
    pub fn b200rt_reset_stats(ctx: *mut b200rt_ctx) -> c_int;
    pub fn b200rt_set_kernel_timing(ctx: *mut b200rt_ctx, enabled: c_int) -> c_int;

    // device groups: the epoch loop of main.rs:1129-1173 (or the rows of main.rs:1090) sharded over GPUs, reduced on rank 0
    pub fn b200rt_group_create(device_ids: *const c_int, n_devices: c_int, out_group: *mut *mut b200rt_group) -> c_int;
    pub fn b200rt_group_unique_id(id_out: *mut c_void, id_bytes: usize) -> c_int;
    pub fn b200rt_group_create_rank(device_id: c_int, rank: c_int, n_ranks: c_int, unique_id: *const c_void, id_bytes: usize,
                                    out_group: *mut *mut b200rt_group) -> c_int;
    pub fn b200rt_group_destroy(group: *mut b200rt_group) -> c_int;
    pub fn b200rt_group_size(group: *const b200rt_group, n_ranks: *mut c_int, n_local: *mut c_int) -> c_int;
    pub fn b200rt_group_ctx(group: *mut b200rt_group, local_index: c_int, out_ctx: *mut *mut b200rt_ctx) -> c_int;
    pub fn b200rt_group_last_error(group: *const b200rt_group) -> *const c_char;
    pub fn b200rt_group_upload_scene(group: *mut b200rt_group, scene: *const b200rt_scene) -> c_int;
    pub fn b200rt_group_render_distributed(group: *mut b200rt_group, cam: *const b200rt_camera, params: *const b200rt_params,
                                           epoch_begin: u32, epoch_count: u32, out_accum: *mut f32) -> c_int;
    pub fn b200rt_group_render_distributed_device(group: *mut b200rt_group, cam: *const b200rt_camera, params: *const b200rt_params,
                                                  epoch_begin: u32, epoch_count: u32, d_accum_root: *mut f32) -> c_int;
    pub fn b200rt_group_render_distributed_rows(group: *mut b200rt_group, cam: *const b200rt_camera, params: *const b200rt_params,
                                                epoch_begin: u32, epoch_count: u32, out_accum: *mut f32) -> c_int;
    pub fn b200rt_group_render_distributed_rows_device(group: *mut b200rt_group, cam: *const b200rt_camera, params: *const b200rt_params,
                                                       epoch_begin: u32, epoch_count: u32, d_accum_root: *mut f32) -> c_int;
    pub fn b200rt_group_render_whitted(group: *mut b200rt_group, cam: *const b200rt_camera, params: *const b200rt_params,
                                       out_rgb: *mut f32, out_prim_id: *mut i32, want_prim_ids: c_int) -> c_int;
    pub fn b200rt_group_render_whitted_device(group: *mut b200rt_group, cam: *const b200rt_camera, params: *const b200rt_params,
                                              d_rgb_root: *mut f32) -> c_int;
    pub fn b200rt_group_last_render_ms(group: *const b200rt_group, ms: *mut f32) -> c_int;

    // the builder surface for hosts without the Rust World (the reference crate lowers its own World instead)
    pub fn b200rt_world_new() -> *mut b200rt_world;
    pub fn b200rt_world_free(w: *mut b200rt_world);
    pub fn b200rt_world_push_object(w: *mut b200rt_world, material: *const b200rt_material) -> c_int;
    pub fn b200rt_world_push_triangle(w: *mut b200rt_world, object_index: u32, v: *const b200rt_vertex) -> c_int;
    pub fn b200rt_world_push_sphere(w: *mut b200rt_world, object_index: u32, center: *const f32, radius: f32) -> c_int;
    pub fn b200rt_world_push_light(w: *mut b200rt_world, light: *const b200rt_light) -> c_int;
    pub fn b200rt_world_load_obj(w: *mut b200rt_world, object_index: u32, path: *const c_char, scale_div: f32, offset: *const f32) -> c_int;
    pub fn b200rt_world_load_obj_ex(w: *mut b200rt_world, object_index: u32, path: *const c_char, scale_div: f32, offset: *const f32,
                                    model_index: i32, flags: u32) -> c_int;
    pub fn b200rt_obj_model_count(path: *const c_char) -> c_int;
    pub fn b200rt_world_scene(w: *const b200rt_world, out: *mut b200rt_scene) -> c_int;
    pub fn b200rt_world_fixture(w: *mut b200rt_world, obj_path: *const c_char) -> c_int;
    pub fn b200rt_fixture_camera(cam: *mut b200rt_camera);
    pub fn b200rt_default_params(p: *mut b200rt_params);
}

/// `check(code)`: the reference panics on errors (main.rs:785, 767-775); so does this helper.
pub fn check(code: c_int) -> c_int {
    if code < 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(b200rt_strerror(code)) };
        panic!("b200rt: error {} ({})", code, msg.to_string_lossy());
    }
    code
}
