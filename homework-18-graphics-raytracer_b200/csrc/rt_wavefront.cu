// rt_wavefront.cu — the stochastic tracer (distributed_ray_trace, main.rs:521-614, driven by the epoch loop
// main.rs:1129-1167) as a WAVEFRONT pipeline: every pixel sample is a path whose state lives in HBM, and the
// frame advances in rounds of two kernels
//
//     wf_cast_rl_kernel World::cast (main.rs:180-326) for every ray requested in the previous round — the
//      (_tiled / wf_cast_kernel)  packed-FFMA2 filter with rays in lanes (rt_cast_rl.cuh; tiles by TMA for scenes of
//                       more than 64 triangles; warp-transposed alternative: rt_cast.cuh) + the certified exact
//                       select, all 32 lanes of every warp carrying rays of one kind (the FP32-roofline kernel);
//     wf_logic_kernel   everything between two casts (camera, get_shade, scatter, reflect / refract,
//                       accumulation), run on queues that are binned by WHAT the finished cast was for
//                       (primary / bounce / shadow rays / refraction step), so each warp executes one
//                       branch of the reference's recursion with full lanes.  With fused levels a hit's shadow
//                       rays and the next level's ray travel in the same round (one pass per level).
//
// The per-path transitions are the same ones, in the same arithmetic, as the phase machine of rt_kernels.cu
// (trace_kernel<kModeDistributed>): with one epoch in flight per pixel both tracers produce bit-identical
// {sum.rgb, count} accumulators.
//
// HBM layout (n = paths in flight = pixels of the rendered rows x epochs-in-flight `epar`).  Queues permute the
// paths, so everything a path owns is ARRAY-OF-STRUCTS in whole 32-byte sectors: a lane that reads its path's
// rows uses every byte of every sector it touches, however the queue ordered the paths.
//   st  [n][12]  path state, float4 rows (see ROW_*)                                   192 B / path
//   req [n][6]   {origin, meta}{direction} of the path ray + the directions of up to 4 shadow rays
//                (a shadow ray's origin / exclusion is the current hit: ROW_HPOS)      96 B
//   res [n][2]   result of the path ray {prim, meta, t, uv.x}{normal, uv.y}             32 B
//   sres[n][4]   float2 {prim, t} per shadow ray                                        32 B
//   q   [2][6][n] path ids per consumer segment (double buffered by round parity)       48 B
//   work[2][5][n] cast work: one list of path ids per ray slot (path ray, 4 shadow rays)  40 B
//   sums[n]      every slot's PhotonAccumulator {sum.rgb, weight_sum}, dense [slot][pixel]   16 B
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include <string>

#include "rt_cast.cuh"
#include "rt_cast_rl.cuh"
#include "rt_bvh.cuh"
#include "rt_shade.cuh"
#include "rt_types.h"
#include "rt_wavefront.h"

namespace b200rt {

namespace {

// rows that are written together are neighbours, so every store pair fills one whole 32-byte sector
enum : int {
    ROW_CTRL = 0,     // flags, depth | rng draws << 16, pixel x | y << 16, sample index (4 x u32)
    ROW_RNG,          // Philox block of the sample stream
    ROW_ACC,          // radiance of the sample so far
    ROW_T,            // throughput: what a unit of radiance at the current hit adds to the sample
    ROW_HPOS,         // hit.at.position, hit.index
    ROW_HNORMAL,      // hit.at.normal, meta (face | ray_face << 1 | object << 8)
    ROW_HDIR,         // hit.ray.direction (after scatter_hit), uv.x
    ROW_HDIR0,        // hit.ray.direction before scatter_hit (view direction of the BRDF probes), uv.y
    ROW_PEND,         // pending factor (BRDF probe value or decay^distance); w = get_refract travel distance
    ROW_NADJ,         // adjust_normal(hit.at.normal) of the current hit, between get_shade's entry and its sum
    ROW_HI_POS,       // get_refract: previous inside hit position, retry count | get_shade: sum of earlier light chunks
    ROW_SPARE2,       // (the slot's PhotonAccumulator lived here; it is a dense array of its own now: WfBuffers::sums)
    kStateRows
};
static_assert(kStateRows == WF_STATE_ROWS, "state rows");
static_assert(ROW_CTRL % 2 == 0 && ROW_RNG == ROW_CTRL + 1 && ROW_ACC % 2 == 0 && ROW_T == ROW_ACC + 1 && ROW_HPOS % 2 == 0 &&
              ROW_HNORMAL == ROW_HPOS + 1 && ROW_HDIR % 2 == 0 && ROW_HDIR0 == ROW_HDIR + 1 && ROW_PEND % 2 == 0 &&
              ROW_NADJ == ROW_PEND + 1 && ROW_HI_POS % 2 == 0 && ROW_SPARE2 == ROW_HI_POS + 1, "rows that travel together share a sector");
#if WF_REQ_HP
enum : int { REQ_O = 0, REQ_D = 1, REQ_SD3 = 2, REQ_HP = 4, REQ_SD0 = 5 };
RT_DI int req_shadow_row(uint32_t s) { return s < 3u ? REQ_SD0 + (int)s : REQ_SD3; }
#else
enum : int { REQ_O = 0, REQ_D = 1, REQ_SHADOW_D = 2 };
RT_DI int req_shadow_row(uint32_t s) { return REQ_SHADOW_D + (int)s; }
#endif

// flags word of ROW_CTRL
enum : uint32_t {
    F_A_KNOWN = 1u << 0,
    F_RAYTYPE_SHIFT = 1,        // 2 bits: RayType (main.rs:532): 0 diffuse, 1 reflection, 2 refraction
    F_PURPOSE_SHIFT = 3,        // 2 bits: SH_FINAL / SH_NEXT_MIX / SH_NEXT_REFR
    F_TIR = 1u << 5,            // refraction step in flight is a total-internal-reflection bounce (else the first inside ray)
    F_NEED_SHIFT = 6,           // 4 bits: lights of the current chunk with a shadow ray in flight
    F_LI0_SHIFT = 10,           // 12 bits: first light of the current chunk
    F_PARTIAL = 1u << 22,       // ROW_HI_POS holds the get_shade sum of earlier chunks
    // fused levels (scenes of <= 4 lights): the next level's select / scatter ran when the hit was reached, and its
    // ray travels in the same round as the hit's shadow rays
    F_PRE = 1u << 23,           // a path ray (bounce, or the first inside ray of get_refract) is in flight with the shadow rays
    F_BLACK = 1u << 24,         // the next level ends the sample black (main.rs:559-561 / 366-368): close it after get_shade
    F_FRESH = 1u << 25          // radiance / throughput are still the opening values (0, 1): their rows were not written yet
};
enum : uint32_t { SH_FINAL = 0, SH_NEXT_MIX = 1, SH_NEXT_REFR = 2 };

enum : int { OUT_NONE = 0, OUT_PRIMARY, OUT_BOUNCE, OUT_REFR, OUT_SHADE, OUT_SHB, OUT_RETIRE };

RT_DI float u2f(uint32_t v) { return __uint_as_float(v); }
RT_DI uint32_t f2u(float v) { return __float_as_uint(v); }

RT_DI uint32_t pack_ray_meta(uint32_t face, int32_t ex_prim, uint32_t ex_face) {
    return face | (ex_face << 2) | ((uint32_t)(ex_prim + 1) << 4);
}

// Two neighbouring float4 rows that share a 32-byte sector, as ONE 256-bit access (LDG.E.256 / STG.E.256 on sm_100):
// half the load / store instructions and L1 requests of the path-row gathers.  p must be 32-byte aligned.
#ifndef WF_ROWS_256
#define WF_ROWS_256 1
#endif
RT_DI void ld_rows2(const float4* p, float4& a, float4& b) {
#if WF_ROWS_256
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
    a = p[0]; b = p[1];
#endif
}
template <bool W256 = true>
RT_DI void st_rows2(float4* p, float4 a, float4 b) {
#if WF_ROWS_256
    if (W256) {
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
        return;
    }
#endif
    p[0] = a; p[1] = b;
}

struct PathMem {
    float4* st; float4* req; uint32_t pid;
    RT_DI float4 ld(int row) const { return st[(size_t)pid * kStateRows + row]; }
    RT_DI void sv(int row, float4 v) const { st[(size_t)pid * kStateRows + row] = v; }
    RT_DI void ld2(int row, float4& a, float4& b) const { ld_rows2(st + (size_t)pid * kStateRows + row, a, b); }   // row even
    template <bool W256 = true>
    RT_DI void sv2(int row, float4 a, float4 b) const { st_rows2<W256>(st + (size_t)pid * kStateRows + row, a, b); }
    template <bool W256 = true>
    RT_DI void put_ray(const DRay& r) const {
        st_rows2<W256>(req + (size_t)pid * WF_REQ_ROWS + REQ_O,
                 make_float4(r.o.x, r.o.y, r.o.z, u2f(pack_ray_meta(r.face, r.ex_prim, r.ex_face))), make_float4(r.d.x, r.d.y, r.d.z, 0.0f));
    }
    RT_DI void get_ray(DRay& r) const {
        float4 a, b;
        ld_rows2(req + (size_t)pid * WF_REQ_ROWS + REQ_O, a, b);
        const uint32_t m = f2u(a.w);
        r.o = mk3(a); r.d = mk3(b); r.face = m & 3u; r.ex_face = (m >> 2) & 3u; r.ex_prim = (int32_t)(m >> 4) - 1;
    }
    // .w: the spot light's angular factor, so that get_shade does not evaluate the light a second time
    RT_DI void put_shadow_dir(uint32_t s, f3 d, float angular = 0.0f) const { req[(size_t)pid * WF_REQ_ROWS + req_shadow_row(s)] = make_float4(d.x, d.y, d.z, angular); }
    RT_DI float4 get_shadow_dir(uint32_t s) const { return req[(size_t)pid * WF_REQ_ROWS + req_shadow_row(s)]; }
#if WF_REQ_HP
    // the origin / exclusion of the shadow rays: a copy of the current hit next to their directions
    RT_DI void put_shadow_origin(f3 p, int32_t prim) const { req[(size_t)pid * WF_REQ_ROWS + REQ_HP] = make_float4(p.x, p.y, p.z, __int_as_float(prim)); }
    RT_DI void put_req_pad() const { req[(size_t)pid * WF_REQ_ROWS + REQ_SD3 + 1] = make_float4(0.f, 0.f, 0.f, 0.f); }
#endif
    // shadow ray of light slot s (main.rs:423-431): from the current hit, back faces only, the hit primitive excluded
    RT_DI void get_shadow_ray(uint32_t s, DRay& r) const {
#if WF_REQ_HP
        const float4 a = req[(size_t)pid * WF_REQ_ROWS + REQ_HP], b = req[(size_t)pid * WF_REQ_ROWS + req_shadow_row(s)];
#else
        const float4 a = ld(ROW_HPOS), b = req[(size_t)pid * WF_REQ_ROWS + req_shadow_row(s)];
#endif
        r.o = mk3(a); r.d = mk3(b); r.face = kBack; r.ex_prim = __float_as_int(a.w); r.ex_face = kBack;
    }
};

// The cast work of one round: ONE LIST PER RAY SLOT (0 = the path rays, s = the shadow rays of light slot s - 1), so
// that a warp of the cast kernels carries rays of one kind (shadow rays of one light are coherent and cull the same
// triangles; only path rays need hit attributes) and the shading kernels append with coalesced stores.  The cast
// kernels walk a virtual index space in which every list is padded to whole 128-ray blocks.
struct WfWork {
    const uint32_t* __restrict__ lists;   // [WF_WORK_PER_PATH][n]
    uint32_t n;
    uint32_t count[WF_WORK_PER_PATH];
    uint32_t first_block[WF_WORK_PER_PATH + 1];
    RT_DI uint32_t n_virtual() const { return first_block[WF_WORK_PER_PATH] * 128u; }
    RT_DI uint32_t n_real() const { return count[0] + count[1] + count[2] + count[3] + count[4]; }
    // (path << 3 | slot) of virtual index v, or 0xffffffff for the padding of a list
    RT_DI uint32_t item(uint32_t v) const {
        const uint32_t block = v >> 7;
        const uint32_t k = (block >= first_block[1] ? 1u : 0u) + (block >= first_block[2] ? 1u : 0u) +
                           (block >= first_block[3] ? 1u : 0u) + (block >= first_block[4] ? 1u : 0u);
        const uint32_t fb = k == 0u ? first_block[0] : k == 1u ? first_block[1] : k == 2u ? first_block[2] : k == 3u ? first_block[3] : first_block[4];
        const uint32_t cn = k == 0u ? count[0] : k == 1u ? count[1] : k == 2u ? count[2] : k == 3u ? count[3] : count[4];
        const uint32_t i = v - fb * 128u;
        return i < cn ? (lists[(size_t)k * n + i] << 3) | k : 0xffffffffu;
    }
};
RT_DI WfWork wf_work(const WfBuffers& wb, uint32_t buf) {
    WfWork w;
    w.lists = wb.work + (size_t)buf * WF_WORK_PER_PATH * wb.n;
    w.n = wb.n;
    uint32_t fb = 0u;
#pragma unroll
    for (uint32_t k = 0; k < WF_WORK_PER_PATH; ++k) {
        w.count[k] = wb.ctl->c[buf].work[k];
        w.first_block[k] = fb;
        fb += (w.count[k] + 127u) >> 7;
    }
    w.first_block[WF_WORK_PER_PATH] = fb;
    return w;
}

// Camera::shoot_focus for pixel (px, py), sample `sample_idx` of the launch (main.rs:101-127, 1143-1149): seeds the
// sample's stream and draws the lens offsets (Box-Muller on two stream uniforms).  ONE definition for the
// pass that opens a sample and for the round-0 cast / consumer that regenerate it: same expressions, same bits.
RT_DI void wf_open_sample(const DCamera& cam, const DParams& p, uint32_t px, uint32_t py, uint32_t sample_idx, Rng& rng, DRay& ray) {
    const f3 cam_toward = mk3(cam.toward), cam_x = mk3(cam.x), cam_y = mk3(cam.y);
    float clip_x, clip_y;
    clip_y = ((float)p.height / 2.0f - (float)py) / (float)p.height;   // main.rs:1094
    clip_x = ((float)px - (float)p.width / 2.0f) / (float)p.height;    // main.rs:1095
    const f3 pinhole_dir = normalize(clip_x * cam_x + clip_y * cam_y + cam_toward);   // main.rs:110
    rng_init(rng, p.seed_lo, p.seed_hi, py, px, p.epoch_begin + sample_idx);
    const float u1 = 1.0f - rng_uniform(rng);
    const float u2 = rng_uniform(rng);
    const float radius = sqrtf(-2.0f * nl_logf(u1));
    const float ang = 2.0f * kPi * u2;
    const float2 sca = nl_sincosf(ang);
    const float xoffset = p.blur * (radius * sca.y);
    const float yoffset = p.blur * (radius * sca.x);
    ray.d = normalize(pinhole_dir * p.focus + cam_x * xoffset + cam_y * yoffset);        // main.rs:115-117
    ray.o = mk3(cam.center) + normalize(cam_toward) * cam.near - (cam_x * xoffset + cam_y * yoffset);  // :118-120
    ray.face = kFront; ray.ex_prim = -1; ray.ex_face = kFront;
}

}  // namespace

// ---- cast --------------------------------------------------------------------------------------------------
#ifndef WF_CAST_MIN_BLOCKS
#define WF_CAST_MIN_BLOCKS 4
#endif
#ifndef WF_CAST_PREFETCH
#define WF_CAST_PREFETCH 0   // measured on B200: the register-held prefetch of the next chunk spills and gains nothing (235 vs 238 ms)
#endif
__global__ void __launch_bounds__(128, WF_CAST_MIN_BLOCKS) wf_cast_kernel(const DScene sc, const WfBuffers wb, const uint32_t buf,
                                                         DCounters* __restrict__ cnt) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    __shared__ float4 s_rays_all[4][kCastSlotFloat4];
    float4* s_rays = s_rays_all[warp];
    // the counters the logic kernel of this round appends to
    if (blockIdx.x == 0 && threadIdx.x < sizeof(WfCounters) / 4) reinterpret_cast<uint32_t*>(&wb.ctl->c[buf ^ 1u])[threadIdx.x] = 0u;
    const WfWork wk = wf_work(wb, buf);
    const uint32_t n_real = wk.n_real(), n_work = wk.n_virtual();
    if (n_real == 0u) return;
    TriPair tile0;
    if (sc.n_tris_padded) load_tripair(sc.tri_filter, 0, lane, tile0);
    else {
        const float2 z = make_float2(0.f, 0.f);
        tile0.nx = tile0.ny = tile0.nz = tile0.d = tile0.m0x = tile0.m0y = tile0.m0z = tile0.w0 = tile0.m1x = tile0.m1y = tile0.m1z =
            tile0.w1 = tile0.m2x = tile0.m2y = tile0.m2z = tile0.w2 = z;
    }
    CastStats cs;
    cs.casts = cs.confirms = cs.fallbacks = 0ull;
    const uint32_t stride = gridDim.x * 4u * 32u;
    // software pipeline: the work item and ray of the NEXT chunk are fetched before the current chunk's filter loop
    // (a dependent chain work[] -> path -> ray rows of ~2 us that 4 warps per sub-partition cannot hide)
    auto fetch = [&](uint32_t idx, uint32_t& item, DRay& r) {
        r.o = mk3(0.f, 0.f, 0.f); r.d = mk3(0.f, 0.f, 1.f); r.face = kFront; r.ex_prim = -1; r.ex_face = kFront;
        item = 0xffffffffu;                                  // padding of a work list: no ray
        if (idx < n_work) item = wk.item(idx);
        if (item != 0xffffffffu) {
            const PathMem pm{wb.st, wb.req, item >> 3};
            if ((item & 7u) == 0u) pm.get_ray(r);
            else pm.get_shadow_ray((item & 7u) - 1u, r);
        }
    };
    uint32_t base = (blockIdx.x * 4u + warp) * 32u;
    uint32_t item_next;
    DRay r_next;
    fetch(base + lane, item_next, r_next);
    for (; base < n_work; base += stride) {
#if WF_CAST_PREFETCH
        const uint32_t item = item_next;
        const DRay r = r_next;
        fetch(base + stride + lane, item_next, r_next);
#else
        uint32_t item;
        DRay r;
        fetch(base + lane, item, r);
#endif
        const bool active = item != 0xffffffffu;
        const uint32_t pid = item >> 3, slot = item & 7u;
        DHit h;
        h.prim = -1; h.face = 0; h.object = 0; h.t = 0.f; h.pos = h.normal = mk3(0.f, 0.f, 0.f); h.uv.x = h.uv.y = 0.f;
        warp_cast(sc, s_rays, tile0, lane, active, r, h, cs, slot == 0u);
        if (active) {
            if (slot == 0u) {
                const uint32_t meta = h.prim >= 0 ? (h.face | (h.object << 8)) : 0u;
                wb.res[(size_t)pid * 2u + 0u] = make_float4(__int_as_float(h.prim), u2f(meta), h.t, h.uv.x);
                wb.res[(size_t)pid * 2u + 1u] = make_float4(h.normal.x, h.normal.y, h.normal.z, h.uv.y);
            } else {
                wb.sres[(size_t)pid * 4u + (slot - 1u)] = make_float2(__int_as_float(h.prim), h.t);
            }
        }
    }
    if (cnt) {
        unsigned long long n_conf = cs.confirms, n_fb = cs.fallbacks;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_conf += __shfl_xor_sync(0xffffffffu, n_conf, o);
            n_fb += __shfl_xor_sync(0xffffffffu, n_fb, o);
        }
        if (lane == 0u && n_conf) atomicAdd(&cnt->confirms, n_conf);
        if (lane == 0u && n_fb) atomicAdd(&cnt->fallbacks, n_fb);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicAdd(&cnt->casts, (unsigned long long)n_real);
            atomicAdd(&cnt->tri_pairs, (unsigned long long)n_real * sc.n_tris);
            atomicAdd(&cnt->sph_pairs, (unsigned long long)n_real * sc.n_sph);
        }
    }
}

// ---- cast, rays in lanes (scenes of one tile: <= 64 triangles): rt_cast_rl.cuh ---------------------------------------
// The work list of the round is the ray source: item = path << 3 | slot; slot 0 = the path ray, 1..4 = shadow rays whose
// origin / exclusion is the path's current hit.  The items of the warp's NEXT block are re-read after the filter loop to
// pull their rows into L2 while phase 2 runs (the chain work[] -> path -> rows is two DRAM round trips otherwise).
// (A prefetch.global.L2 of the rows of the warp's next block while phase 2 runs was measured on B200 in round 2: 86.0 ms
// per 16-epoch 4K batch with it, 78.5 without - it pulls whole 128-byte lines where the loads touch one 32-byte sector.)
RT_DI void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#ifndef WF_CAST_RL_MIN_BLOCKS
#define WF_CAST_RL_MIN_BLOCKS 6   // measured on B200, cast of a 16-epoch 4K batch: 4 -> 108.1 ms, 5 -> 103.1, 6 -> 101.1 (latency-bound phase 2)
#endif
namespace {
struct WfRayIO {
    WfBuffers wb;
    WfWork work;
    // the list the current 128-ray block lies in (lists are padded to whole blocks: warp-uniform, set by begin_block).
    // (Visiting the lists interleaved in groups, so that the slots of a path are cast close in time, was measured on B200
    // in round 2: the DRAM traffic did not drop - the rows are over-fetched by DRAM granularity, not re-read - and the
    // cast of a 16-epoch 4K batch took 94.7 ms instead of 81.3.)
    const uint32_t* cur_list = nullptr;
    uint32_t cur_first = 0u, cur_count = 0u, cur_slot = 0u;
    RT_DI void begin_block(uint32_t base) {
        const uint32_t block = base >> 7;
        const uint32_t k = (block >= work.first_block[1] ? 1u : 0u) + (block >= work.first_block[2] ? 1u : 0u) +
                           (block >= work.first_block[3] ? 1u : 0u) + (block >= work.first_block[4] ? 1u : 0u);
        cur_slot = k;
        cur_first = (k == 0u ? work.first_block[0] : k == 1u ? work.first_block[1] : k == 2u ? work.first_block[2] : k == 3u ? work.first_block[3] : work.first_block[4]) * 128u;
        cur_count = k == 0u ? work.count[0] : k == 1u ? work.count[1] : k == 2u ? work.count[2] : k == 3u ? work.count[3] : work.count[4];
        cur_list = work.lists + (size_t)k * work.n;
    }
    RT_DI uint32_t item(uint32_t idx) const {
        const uint32_t i = idx - cur_first;
        return i < cur_count ? (cur_list[i] << 3) | cur_slot : 0xffffffffu;     // (padding of the list: no ray)
    }
    // prefetching (rt_cast_rl.cuh): where a future block's items lie, the item of an index from its prefetched list entry,
    // and the rows fetch() will read for a path of ray slot `slot`, pulled into L2 by 16-byte cp.async copies
#if RL_PREFETCH
    static constexpr bool kPrefetch = true;
    struct Loc { const uint32_t* list; uint32_t first, count, slot; };
    RT_DI Loc locate(uint32_t block) const {
        const uint32_t k = (block >= work.first_block[1] ? 1u : 0u) + (block >= work.first_block[2] ? 1u : 0u) +
                           (block >= work.first_block[3] ? 1u : 0u) + (block >= work.first_block[4] ? 1u : 0u);
        Loc l;
        l.slot = k;
        l.first = (k == 0u ? work.first_block[0] : k == 1u ? work.first_block[1] : k == 2u ? work.first_block[2] : k == 3u ? work.first_block[3] : work.first_block[4]) * 128u;
        l.count = k == 0u ? work.count[0] : k == 1u ? work.count[1] : k == 2u ? work.count[2] : k == 3u ? work.count[3] : work.count[4];
        l.list = work.lists + (size_t)k * work.n;
        return l;
    }
    RT_DI uint32_t item_from(uint32_t idx, uint32_t list_entry) const {
        const uint32_t i = idx - cur_first;
        return i < cur_count ? (list_entry << 3) | cur_slot : 0xffffffffu;
    }
    RT_DI void touch(uint32_t pid, uint32_t slot, void* scratch) const {
        if (slot == 0u) rl_cp_async16(scratch, wb.req + (size_t)pid * WF_REQ_ROWS + REQ_O);      // {origin, meta}{direction}: one sector
        else {
#if WF_REQ_HP
            rl_cp_async16(scratch, wb.req + (size_t)pid * WF_REQ_ROWS + REQ_HP);
#else
            rl_cp_async16(scratch, wb.st + (size_t)pid * kStateRows + ROW_HPOS);
#endif
            rl_cp_async16(scratch, wb.req + (size_t)pid * WF_REQ_ROWS + req_shadow_row(slot - 1u));
        }
    }
#endif
    RT_DI void fetch(uint32_t tag, DRay& r) const {
        const PathMem pm{wb.st, wb.req, tag >> 3};
        if (cur_slot == 0u) pm.get_ray(r);
        else pm.get_shadow_ray(cur_slot - 1u, r);
    }
    RT_DI bool want_attrs(uint32_t tag) const { return (tag & 7u) == 0u; }   // shadow rays: main.rs:435-447
    RT_DI bool all_sphere_uv() const { return false; }                       // uv of a sphere hit only where a material reads it
    // shadow slot s of a one-chunk scene is light s: a directional light's shadow ray culls the same triangles everywhere
    RT_DI uint2 culled(const DScene& sc, uint32_t tag, uint32_t tile) const {
        const uint32_t slot = tag & 7u;
        if (slot == 0u || sc.shadow_cull == nullptr) return make_uint2(0u, 0u);
        return sc.shadow_cull[(size_t)(slot - 1u) * (sc.n_tris_padded / kTileTris) + tile];
    }
    // one-tile scenes: the class of a ray is its slot, the masks are copied to shared memory once per CTA
    RT_DI uint32_t cull_class(uint32_t tag) const { return tag & 7u; }
    RT_DI uint2 cull_mask(const DScene& sc, uint32_t cls) const {
        if (cls == 0u || cls > 4u || sc.shadow_cull == nullptr || cls > sc.n_lights) return make_uint2(0u, 0u);
        return sc.shadow_cull[(size_t)(cls - 1u) * (sc.n_tris_padded / kTileTris)];
    }
    RT_DI void store(uint32_t tag, const DHit& h) const {
        const uint32_t pid = tag >> 3, slot = tag & 7u;
        if (slot == 0u) {
            const uint32_t m2 = h.prim >= 0 ? (h.face | (h.object << 8)) : 0u;
            st_rows2(wb.res + (size_t)pid * 2u, make_float4(__int_as_float(h.prim), u2f(m2), h.t, h.uv.x),
                     make_float4(h.normal.x, h.normal.y, h.normal.z, h.uv.y));
        } else {
            wb.sres[(size_t)pid * 4u + (slot - 1u)] = make_float2(__int_as_float(h.prim), h.t);
        }
    }
};
}  // namespace
__global__ void __launch_bounds__(kRlThreads, WF_CAST_RL_MIN_BLOCKS) wf_cast_rl_kernel(const DScene sc, const __grid_constant__ RlTileParam tp,
                                                                                      const WfBuffers wb, const uint32_t buf,
                                                                                      DCounters* __restrict__ cnt) {
    __shared__ RlShared sh;
    const uint32_t lane = threadIdx.x & 31u;
    if (blockIdx.x == 0 && threadIdx.x < sizeof(WfCounters) / 4) reinterpret_cast<uint32_t*>(&wb.ctl->c[buf ^ 1u])[threadIdx.x] = 0u;
    const WfWork wk = wf_work(wb, buf);
    const uint32_t n_work = wk.n_real();
    if (n_work == 0u) return;
    CastStats cs;
    cs.casts = cs.confirms = cs.fallbacks = 0ull;
    WfRayIO io;
    io.wb = wb; io.work = wk;
    cast_rays_in_lanes(sc, tp, io, wk.n_virtual(), sh, cs);
    if (cnt) {
        unsigned long long n_conf = cs.confirms, n_fb = cs.fallbacks;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_conf += __shfl_xor_sync(0xffffffffu, n_conf, o);
            n_fb += __shfl_xor_sync(0xffffffffu, n_fb, o);
        }
        if (lane == 0u && n_conf) atomicAdd(&cnt->confirms, n_conf);
        if (lane == 0u && n_fb) atomicAdd(&cnt->fallbacks, n_fb);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicAdd(&cnt->casts, (unsigned long long)n_work);
            atomicAdd(&cnt->tri_pairs, (unsigned long long)n_work * sc.n_tris);
            atomicAdd(&cnt->sph_pairs, (unsigned long long)n_work * sc.n_sph);
        }
    }
}

// ---- cast, rays in lanes, scenes of more than one tile: the tiles stream through shared memory by TMA --------------------
#ifndef WF_CAST_RL_TILED_MIN_BLOCKS
#define WF_CAST_RL_TILED_MIN_BLOCKS 5
#endif
// Rounds of few rays (at most split_below blocks of 128) go to wf_cast_rl_tiled_split_kernel instead, which is launched right
// after this one: the work counters live on the device, so both kernels are enqueued every round and one of them returns.
__global__ void __launch_bounds__(kRlThreads, WF_CAST_RL_TILED_MIN_BLOCKS) wf_cast_rl_tiled_kernel(const DScene sc, const WfBuffers wb,
                                                                                                  const uint32_t buf,
                                                                                                  DCounters* __restrict__ cnt,
                                                                                                  const uint32_t split_below) {
    __shared__ RlTiledShared sh;
    const uint32_t lane = threadIdx.x & 31u;
    if (blockIdx.x == 0 && threadIdx.x < sizeof(WfCounters) / 4) reinterpret_cast<uint32_t*>(&wb.ctl->c[buf ^ 1u])[threadIdx.x] = 0u;
    const WfWork wk = wf_work(wb, buf);
    const uint32_t n_work = wk.n_real();
    if (n_work == 0u) return;
    if ((wk.n_virtual() >> 7) <= split_below) return;
    CastStats cs;
    cs.casts = cs.confirms = cs.fallbacks = 0ull;
    WfRayIO io;
    io.wb = wb; io.work = wk;
    cast_rays_in_lanes_tiled(sc, io, wk.n_virtual(), sh, cs);
    if (cnt) {
        unsigned long long n_conf = cs.confirms, n_fb = cs.fallbacks;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_conf += __shfl_xor_sync(0xffffffffu, n_conf, o);
            n_fb += __shfl_xor_sync(0xffffffffu, n_fb, o);
        }
        if (lane == 0u && n_conf) atomicAdd(&cnt->confirms, n_conf);
        if (lane == 0u && n_fb) atomicAdd(&cnt->fallbacks, n_fb);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicAdd(&cnt->casts, (unsigned long long)n_work);
            atomicAdd(&cnt->tri_pairs, (unsigned long long)n_work * sc.n_tris);
            atomicAdd(&cnt->sph_pairs, (unsigned long long)n_work * sc.n_sph);
        }
    }
}

// the same cast for rounds of few rays: one block of 128 rays per CTA, the tile range split over its four warps
// (rt_cast_rl.cuh: cast_rays_in_lanes_tiled_split).  51 KB of dynamic shared memory.
__global__ void __launch_bounds__(kRlThreads, 4) wf_cast_rl_tiled_split_kernel(const DScene sc, const WfBuffers wb, const uint32_t buf,
                                                                              DCounters* __restrict__ cnt, const uint32_t split_below) {
    extern __shared__ __align__(128) unsigned char wf_split_smem[];
    RlSplitShared& sh = *reinterpret_cast<RlSplitShared*>(wf_split_smem);
    const WfWork wk = wf_work(wb, buf);
    const uint32_t n_work = wk.n_real();
    if (n_work == 0u || (wk.n_virtual() >> 7) > split_below) return;
    CastStats cs;
    cs.casts = cs.confirms = cs.fallbacks = 0ull;
    WfRayIO io;
    io.wb = wb; io.work = wk;
    cast_rays_in_lanes_tiled_split(sc, io, wk.n_virtual(), sh, cs);
    if (cnt) {
        const uint32_t lane = threadIdx.x & 31u;
        unsigned long long n_conf = cs.confirms, n_fb = cs.fallbacks;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_conf += __shfl_xor_sync(0xffffffffu, n_conf, o);
            n_fb += __shfl_xor_sync(0xffffffffu, n_fb, o);
        }
        if (lane == 0u && n_conf) atomicAdd(&cnt->confirms, n_conf);
        if (lane == 0u && n_fb) atomicAdd(&cnt->fallbacks, n_fb);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicAdd(&cnt->casts, (unsigned long long)n_work);
            atomicAdd(&cnt->tri_pairs, (unsigned long long)n_work * sc.n_tris);
            atomicAdd(&cnt->sph_pairs, (unsigned long long)n_work * sc.n_sph);
        }
    }
}

// ---- round 0 of the rays-in-lanes casts: the first camera ray of every slot, generated in the cast (no INIT pass, no
// request rows, no work list).  Work index = path id; the result goes where the consumer (WF_SEG_PRIMARY0) reads it.
namespace {
struct WfPrimaryIO {
    WfBuffers wb;
    DCamera cam;
    DParams p;
#if RL_PREFETCH
    static constexpr bool kPrefetch = false;   // rays are generated, not gathered
    struct Loc { const uint32_t* list; uint32_t first, count, slot; };
    RT_DI Loc locate(uint32_t) const { return Loc{nullptr, 0u, 0u, 0u}; }
    RT_DI uint32_t item_from(uint32_t idx, uint32_t) const { return idx; }
    RT_DI void touch(uint32_t, uint32_t, void*) const {}
#endif
    RT_DI void begin_block(uint32_t) const {}
    RT_DI uint32_t item(uint32_t idx) const { return idx; }
    RT_DI void fetch(uint32_t pid, DRay& r) const {
        const uint32_t pix = pid % wb.n_pixels, e_lane = pid / wb.n_pixels;
        Rng rng;
        wf_open_sample(cam, p, pix % p.width, wf_frame_row(p, pix / p.width), e_lane, rng, r);
    }
    RT_DI bool want_attrs(uint32_t) const { return true; }
    RT_DI bool all_sphere_uv() const { return false; }
    RT_DI uint2 culled(const DScene&, uint32_t, uint32_t) const { return make_uint2(0u, 0u); }
    RT_DI uint32_t cull_class(uint32_t) const { return 0u; }
    RT_DI uint2 cull_mask(const DScene&, uint32_t) const { return make_uint2(0u, 0u); }
    RT_DI void store(uint32_t pid, const DHit& h) const {
        const uint32_t m2 = h.prim >= 0 ? (h.face | (h.object << 8)) : 0u;
        st_rows2(wb.res + (size_t)pid * 2u, make_float4(__int_as_float(h.prim), u2f(m2), h.t, h.uv.x),
                 make_float4(h.normal.x, h.normal.y, h.normal.z, h.uv.y));
    }
};
RT_DI void wf_cast_tail(const DScene& sc, const CastStats& cs, uint32_t n_work, DCounters* __restrict__ cnt) {
    if (!cnt) return;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long n_conf = cs.confirms, n_fb = cs.fallbacks;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_conf += __shfl_xor_sync(0xffffffffu, n_conf, o);
        n_fb += __shfl_xor_sync(0xffffffffu, n_fb, o);
    }
    if (lane == 0u && n_conf) atomicAdd(&cnt->confirms, n_conf);
    if (lane == 0u && n_fb) atomicAdd(&cnt->fallbacks, n_fb);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&cnt->casts, (unsigned long long)n_work);
        atomicAdd(&cnt->tri_pairs, (unsigned long long)n_work * sc.n_tris);
        atomicAdd(&cnt->sph_pairs, (unsigned long long)n_work * sc.n_sph);
    }
}
}  // namespace
__global__ void __launch_bounds__(kRlThreads, WF_CAST_RL_MIN_BLOCKS) wf_cast_rl_primary_kernel(const DScene sc, const __grid_constant__ RlTileParam tp,
                                                                                              const DCamera cam, const DParams p, const WfBuffers wb,
                                                                                              const uint32_t buf, DCounters* __restrict__ cnt) {
    __shared__ RlShared sh;
    if (blockIdx.x == 0 && threadIdx.x < sizeof(WfCounters) / 4) reinterpret_cast<uint32_t*>(&wb.ctl->c[buf ^ 1u])[threadIdx.x] = 0u;
    CastStats cs;
    cs.casts = cs.confirms = cs.fallbacks = 0ull;
    const WfPrimaryIO io{wb, cam, p};
    cast_rays_in_lanes(sc, tp, io, wb.n, sh, cs);
    wf_cast_tail(sc, cs, wb.n, cnt);
}
__global__ void __launch_bounds__(kRlThreads, WF_CAST_RL_TILED_MIN_BLOCKS) wf_cast_rl_tiled_primary_kernel(const DScene sc, const DCamera cam, const DParams p,
                                                                                                          const WfBuffers wb, const uint32_t buf,
                                                                                                          DCounters* __restrict__ cnt) {
    __shared__ RlTiledShared sh;
    if (blockIdx.x == 0 && threadIdx.x < sizeof(WfCounters) / 4) reinterpret_cast<uint32_t*>(&wb.ctl->c[buf ^ 1u])[threadIdx.x] = 0u;
    CastStats cs;
    cs.casts = cs.confirms = cs.fallbacks = 0ull;
    const WfPrimaryIO io{wb, cam, p};
    cast_rays_in_lanes_tiled(sc, io, wb.n, sh, cs);
    wf_cast_tail(sc, cs, wb.n, cnt);
}

// ---- cast through the acceleration structure (B200RT_CAST_BVH, rt_bvh.cuh): one ray per lane -----------------------------------
namespace {
template <class IO>
RT_DI void wf_cast_rays_bvh(const DScene& sc, IO io, const uint32_t n_work, CastStats& cs) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_chunks = (n_work + 31u) >> 5;
    for (uint32_t chunk = warp0; chunk < n_chunks; chunk += n_warps) {
        const uint32_t idx = chunk * 32u + lane;
        io.begin_block(idx & ~127u);                       // (the lists are padded to 128-index blocks: warp-uniform)
        const uint32_t tag = idx < n_work ? io.item(idx) : kRlNoRay;
        DRay r;
        r.o = mk3(0.f, 0.f, 0.f); r.d = mk3(0.f, 0.f, 1.f); r.face = kFront; r.ex_prim = -1; r.ex_face = kFront;
        if (tag != kRlNoRay) io.fetch(tag, r);
        DHit h;
        h.prim = -1; h.face = 0; h.object = 0; h.t = 0.f; h.pos = h.normal = mk3(0.f, 0.f, 0.f); h.uv.x = h.uv.y = 0.f;
        bvh_warp_cast(sc, lane, tag != kRlNoRay, r, h, cs, tag != kRlNoRay && io.want_attrs(tag), io.all_sphere_uv());
        if (tag != kRlNoRay) io.store(tag, h);
    }
}
}  // namespace
// the traversal chases pointers through L2 (a few hundred dependent node loads per ray): it wants warps, not registers.
// Measured on B200, C5 frame: 4 CTAs of 128 threads per SM 1146 ms, 6: 947, 8: 815, 10: 778, 12: 976, 16: 899
#ifndef WF_BVH_MIN_BLOCKS
#define WF_BVH_MIN_BLOCKS 10
#endif
__global__ void __launch_bounds__(128, WF_BVH_MIN_BLOCKS) wf_cast_bvh_kernel(const DScene sc, const WfBuffers wb, const uint32_t buf, DCounters* __restrict__ cnt) {
    if (blockIdx.x == 0 && threadIdx.x < sizeof(WfCounters) / 4) reinterpret_cast<uint32_t*>(&wb.ctl->c[buf ^ 1u])[threadIdx.x] = 0u;
    const WfWork wk = wf_work(wb, buf);
    const uint32_t n_work = wk.n_real();
    if (n_work == 0u) return;
    CastStats cs;
    cs.casts = cs.confirms = cs.fallbacks = 0ull;
    WfRayIO io;
    io.wb = wb; io.work = wk;
    wf_cast_rays_bvh(sc, io, wk.n_virtual(), cs);
    wf_cast_tail(sc, cs, n_work, cnt);
}
__global__ void __launch_bounds__(128, WF_BVH_MIN_BLOCKS) wf_cast_bvh_primary_kernel(const DScene sc, const DCamera cam, const DParams p, const WfBuffers wb,
                                                                     const uint32_t buf, DCounters* __restrict__ cnt) {
    if (blockIdx.x == 0 && threadIdx.x < sizeof(WfCounters) / 4) reinterpret_cast<uint32_t*>(&wb.ctl->c[buf ^ 1u])[threadIdx.x] = 0u;
    CastStats cs;
    cs.casts = cs.confirms = cs.fallbacks = 0ull;
    const WfPrimaryIO io{wb, cam, p};
    wf_cast_rays_bvh(sc, io, wb.n, cs);
    wf_cast_tail(sc, cs, wb.n, cnt);
}

// ---- logic -------------------------------------------------------------------------------------------------

// One instantiation per segment: each is a small kernel (the code of the other segments is compiled out), so it keeps
// few registers and many warps in flight — these kernels are bound by the latency of gathering path rows from HBM.
#ifndef WF_FUSED_SHADE_MIN_BLOCKS
#define WF_FUSED_SHADE_MIN_BLOCKS 2
#endif
#ifndef WF_FUSED_SHADE_THREADS
#define WF_FUSED_SHADE_THREADS 256
#endif
#ifndef WF_PRIMARY_MIN_BLOCKS
#define WF_PRIMARY_MIN_BLOCKS 3
#endif
#ifndef WF_FUSED_BOUNCE_MIN_BLOCKS
#define WF_FUSED_BOUNCE_MIN_BLOCKS 3
#endif
template <int SEG, bool FUSED = false> struct LogicCfg { static constexpr int kMinBlocks = 3, kThreads = 256; };
template <> struct LogicCfg<WF_SEG_PRIMARY, false> { static constexpr int kMinBlocks = WF_PRIMARY_MIN_BLOCKS, kThreads = 256; };
template <> struct LogicCfg<WF_SEG_BOUNCE, true> { static constexpr int kMinBlocks = WF_FUSED_BOUNCE_MIN_BLOCKS, kThreads = 256; };
template <> struct LogicCfg<WF_SEG_INIT, false> { static constexpr int kMinBlocks = 4, kThreads = 256; };
template <> struct LogicCfg<WF_SEG_PRIMARY0, false> { static constexpr int kMinBlocks = WF_PRIMARY_MIN_BLOCKS, kThreads = 256; };
template <> struct LogicCfg<WF_SEG_REFR, false> { static constexpr int kMinBlocks = 4, kThreads = 256; };
// shade + arrival + next level in one pass
template <> struct LogicCfg<WF_SEG_SHADE, true> { static constexpr int kMinBlocks = WF_FUSED_SHADE_MIN_BLOCKS, kThreads = WF_FUSED_SHADE_THREADS; };

// (Round 2 also measured pulling the rows a fused SHADE pass reads late - the ray and hit that travelled with the shadow
// rays, the shadow directions: a third dependent round trip - into L2 with 16-byte cp.async copies while the state rows are
// in flight: 74.5 ms of shading kernels per 16-epoch 4K batch against 74.7; and the queue entry of the warp's next chunk
// one iteration ahead (WF_LOGIC_PREFETCH = 2): 75.3.  These kernels wait on dependent arithmetic, instruction fetch and
// branches as much as on memory (ncu: long scoreboard 36 % of the stall samples, wait 22 %, no-instruction 14 %).)
#ifndef WF_SHADE_SUM_UNROLL
// the light loops of get_shade (its sum, its entry): unrolled copies measured on B200 - shading kernels of a 16-epoch 4K
// batch 74.9 ms rolled; sum x2 75.4, sum x4 76.6, entry x4 78.0, both x4 78.9 (code size: these kernels already wait on
// instruction fetch)
#define WF_SHADE_SUM_UNROLL 1
#endif
#ifndef WF_SHADE_BEGIN_UNROLL
#define WF_SHADE_BEGIN_UNROLL 1
#endif
constexpr int kShadeSumUnroll = WF_SHADE_SUM_UNROLL, kShadeBeginUnroll = WF_SHADE_BEGIN_UNROLL;
#ifndef WF_LOGIC_PREFETCH
#define WF_LOGIC_PREFETCH 0   // measured on B200: pulling the next chunk's rows into L2 one iteration ahead is SLOWER (94.7 vs 89.2 ms per batch)
#endif
template <int SEG, bool FUSED>
__global__ void __launch_bounds__(LogicCfg<SEG, FUSED>::kThreads, LogicCfg<SEG, FUSED>::kMinBlocks) wf_logic_kernel(const DScene sc, const DCamera cam, const DParams p,
                                                          const WfBuffers wb, const uint32_t buf,
                                                          DCounters* __restrict__ cnt) {
    constexpr int seg = SEG;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nbuf = buf ^ 1u;
    constexpr bool kIndexed = SEG == WF_SEG_INIT || SEG == WF_SEG_PRIMARY0;      // no queue: path id = index
    const uint32_t n_this = kIndexed ? wb.n : wb.ctl->c[buf].seg[SEG];
    const uint32_t total_chunks = (n_this + 31u) >> 5;
    const uint32_t n_epochs = p.epoch_count;
    unsigned long long n_samples = 0ull;

    // the path of this lane in the warp's NEXT chunk is fetched one iteration ahead, and its rows are pulled into L2
    // before the current chunk's stores and queue atomics: the chain queue -> path -> rows is otherwise two DRAM round
    // trips at the top of every iteration of a kernel that runs 4-6 warps per sub-partition
    const uint32_t* __restrict__ q_this = wb.q + ((size_t)buf * WF_SEG_COUNT + (kIndexed ? 0 : seg)) * wb.n;
    uint32_t pid_ahead = 0u;
    if (!kIndexed && WF_LOGIC_PREFETCH) {
        const uint32_t k0 = gw * 32u + lane;
        if (gw < total_chunks && k0 < n_this) pid_ahead = q_this[k0];
    }
    for (uint32_t chunk = gw; chunk < total_chunks; chunk += n_warps) {
        const uint32_t k = chunk * 32u + lane;
        const bool valid = k < n_this;
        const uint32_t pid = !valid ? 0u : (kIndexed ? k : (WF_LOGIC_PREFETCH ? pid_ahead : q_this[k]));
        const PathMem pm{wb.st, wb.req, pid};
        bool valid_ahead = false;
        if (!kIndexed && WF_LOGIC_PREFETCH) {
            const uint32_t k_ahead = (chunk + n_warps) * 32u + lane;
            valid_ahead = chunk + n_warps < total_chunks && k_ahead < n_this;
            if (valid_ahead) pid_ahead = q_this[k_ahead];
        }

        // ---- per-path registers (the names of trace_kernel) --------------------------------------------------
        f3 acc = mk3(0.f, 0.f, 0.f), T = mk3(1.f, 1.f, 1.f), a_shade = mk3(0.f, 0.f, 0.f);
        int32_t depth = 0;
        uint32_t flags = 0u, sample_idx = 0u;
        Rng rng;
        rng.k0 = p.seed_lo; rng.k1 = p.seed_hi; rng.x = rng.y = rng.epoch = rng.draws = 0u; rng.b[0] = rng.b[1] = rng.b[2] = rng.b[3] = 0u;
        DHit h;
        h.prim = -1; h.face = 0; h.object = 0; h.t = 0.f; h.pos = mk3(0.f, 0.f, 0.f); h.normal = mk3(0.f, 0.f, 1.f); h.uv.x = h.uv.y = 0.f;
        f3 h_dir = mk3(0.f, 0.f, 1.f), h_dir_orig = h_dir, pend = mk3(0.f, 0.f, 0.f);
        float rf_travel = 0.0f;
        uint32_t h_rayface = kFront;
        f3 shade = mk3(0.f, 0.f, 0.f);
        bool do_level = false, do_shade_begin = false, do_finish = false;
        f3 nadj_in = mk3(0.f, 0.f, 1.f), nadj_out = nadj_in;   // adjusted normal: read by get_shade's sum, written by its entry
        bool w_nadj = false;
        bool pre_level = false;     // fused levels: do_level runs for the level AFTER the pending get_shade (depth - 1)
        bool pre_ray = false;       // ... and requested a path ray that travels with the shadow rays
        bool w_hit = false, w_dirs = false, w_acc = false, w_pend = false, w_rng = false;   // rows to write back
        int out = OUT_NONE;
        DRay ray;
        ray.o = mk3(0.f, 0.f, 0.f); ray.d = mk3(0.f, 0.f, 1.f); ray.face = kFront; ray.ex_prim = -1; ray.ex_face = kFront;

        // pixel of this path: slot e_lane of pixel pix renders samples e_lane, e_lane + epar, ...  The passes that open a
        // slot (path id = index) derive the pixel from the index; it then travels in the control row, so that the queued
        // passes do not pay three integer divisions per path (pid / n_pixels, pix / width, the strip of the row)
        uint32_t px = 0u, py = 0u;
        if (kIndexed) {
            const uint32_t pix = pid % wb.n_pixels;
            px = pix % p.width; py = wf_frame_row(p, pix / p.width);
        }

        // ---- load: only the rows this segment reads -----------------------------------------------------------------
        if (valid && !kIndexed) {
            constexpr bool kNeedRng = seg == WF_SEG_PRIMARY || seg == WF_SEG_SHADE || (FUSED && seg == WF_SEG_BOUNCE);
            float4 r0, r1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kNeedRng) pm.ld2(ROW_CTRL, r0, r1); else r0 = pm.ld(ROW_CTRL);
            flags = f2u(r0.x); depth = (int32_t)(int16_t)(f2u(r0.y) & 0xffffu); rng.draws = f2u(r0.y) >> 16; sample_idx = f2u(r0.w);
            px = f2u(r0.z) & 0xffffu; py = f2u(r0.z) >> 16;
            rng.x = px; rng.y = py; rng.epoch = p.epoch_begin + sample_idx;
            if (kNeedRng) { rng.b[0] = f2u(r1.x); rng.b[1] = f2u(r1.y); rng.b[2] = f2u(r1.z); rng.b[3] = f2u(r1.w); }
            if (seg != WF_SEG_PRIMARY) {
                float4 r5, r6, r7, r8;
                pm.ld2(ROW_HPOS, r5, r6);
                pm.ld2(ROW_HDIR, r7, r8);
                h.pos = mk3(r5); h.prim = __float_as_int(r5.w);
                h.normal = mk3(r6);
                const uint32_t meta = f2u(r6.w);
                h.face = meta & 1u; h_rayface = (meta >> 1) & 3u; h.object = meta >> 8;
                h_dir = mk3(r7); h.uv.x = r7.w;
                h_dir_orig = mk3(r8); h.uv.y = r8.w;
            }
            if (seg == WF_SEG_SHADE && !(flags & F_FRESH)) {   // (fresh: acc = 0, T = 1 as initialised above)
                float4 ra, rt;
                pm.ld2(ROW_ACC, ra, rt);
                acc = mk3(ra); T = mk3(rt);
            }
            if (seg == WF_SEG_SHADE || seg == WF_SEG_BOUNCE || seg == WF_SEG_REFR) {
                float4 r4;
                if (seg == WF_SEG_SHADE) { float4 r9; pm.ld2(ROW_PEND, r4, r9); nadj_in = mk3(r9); }
                else r4 = pm.ld(ROW_PEND);
                pend = mk3(r4); rf_travel = r4.w;
            }
        }

        // ---- consume the finished cast(s) ----------------------------------------------------------------------------
        // a path ray reached a new hit (main.rs:564-574 / 583-593 / 603-608): BRDF probe / decay of the CURRENT material,
        // then the hit is replaced and its get_shade starts
        auto arrive = [&](const DHit& hc) {
            const uint32_t ray_type = (flags >> F_RAYTYPE_SHIFT) & 3u;
            const MatEval mat = material_approx(sc.materials, h.object, h.uv);
            uint32_t purpose;
            if (ray_type == 2u) {
                pend = mk3(color_pow(mat.opaque_decay, rf_travel), 0.f, 0.f);
                purpose = SH_NEXT_REFR;
            } else {
                pend = ray_type == 0u ? get_diffuse(mat, h.normal, ray.d)                    // main.rs:566-570
                                      : get_specular(mat, h.normal, -h_dir_orig, ray.d);     // main.rs:585-589
                purpose = SH_NEXT_MIX;
            }
            w_pend = true;
            flags = (flags & ~((3u << F_PURPOSE_SHIFT) | F_PRE | F_BLACK)) | (purpose << F_PURPOSE_SHIFT);
            h = hc; h_dir = ray.d; h_dir_orig = ray.d; h_rayface = ray.face; w_hit = true; w_dirs = true;
            do_shade_begin = true;
            // fused levels: distributed_ray_trace's next level at this hit draws its branch and direction now (the
            // sample's draw order is unchanged: nothing else draws in between) and its ray is cast with the shadow rays
            if (FUSED && depth - 1 > 0) { do_level = true; pre_level = true; }
        };
        // one step of get_refract after an inside ray came back (main.rs:371-402)
        auto refract_step = [&](const float4 a, const float4 b) {
            const int32_t hprim = __float_as_int(a.x);
            if (hprim < 0) { if (seg == WF_SEG_REFR && !(flags & F_FRESH)) acc = mk3(pm.ld(ROW_ACC)); do_finish = true; return; }   // Infinite
            const uint32_t meta = f2u(a.y);
            const uint32_t hi_face = meta & 1u;
            const f3 hi_pos = ray.o + ray.d * a.z, hi_normal = mk3(b);
            uint32_t rf_retry;
            if (!(flags & F_TIR)) { rf_travel = distance(hi_pos, h.pos); rf_retry = 0u; }       // main.rs:375
            else {
                const float4 prev = pm.ld(ROW_HI_POS);
                rf_travel = rf_travel + distance(mk3(prev), hi_pos);                            // main.rs:385
                rf_retry = f2u(prev.w) + 1u;                                                    // main.rs:387
            }
            const float rf_k = sc.materials[h.object].refraction_index;
            f3 rout;
            const bool have_out = refract_dir(hi_normal, ray.d, 1.0f / rf_k, rout);             // main.rs:376 / 386
            if (!have_out && rf_travel <= p.refract_max_distance && rf_retry < p.tir_retries) {  // main.rs:378
                pm.sv(ROW_HI_POS, make_float4(hi_pos.x, hi_pos.y, hi_pos.z, u2f(rf_retry)));
                ray = make_reflect(hi_pos, hi_normal, ray.d, ray.face, hprim, hi_face);         // main.rs:379-381
                flags |= F_TIR;
                w_pend = true;
                out = OUT_REFR;
            } else if (!have_out) {
                if (seg == WF_SEG_REFR && !(flags & F_FRESH)) acc = mk3(pm.ld(ROW_ACC));
                do_finish = true;                                                  // Trapped
            } else {                                                               // main.rs:392-402, then 603
                DRay e;
                e.o = hi_pos; e.d = normalize(rout); e.face = kFront; e.ex_prim = hprim; e.ex_face = kBack;
                ray = e;
                w_pend = true;
                out = OUT_BOUNCE;
            }
        };
        auto load_path_hit = [&](float4& a, float4& b, DHit& hc) {
            pm.get_ray(ray);
            ld_rows2(wb.res + (size_t)pid * 2u, a, b);
            hc.prim = __float_as_int(a.x);
            const uint32_t meta = f2u(a.y);
            hc.face = meta & 1u; hc.object = meta >> 8; hc.t = a.z;
            hc.pos = ray.o + ray.d * hc.t;                                         // main.rs:210 / 304
            hc.normal = mk3(b); hc.uv.x = a.w; hc.uv.y = b.w;
        };

        if (seg == WF_SEG_INIT) {
            do_finish = valid;
        } else if (seg == WF_SEG_PRIMARY0) {
            // the slot's first sample: the round-0 cast generated its camera ray from the index; regenerate it (and the
            // sample's stream) instead of reading what an INIT pass would have stored (main.rs:1133-1155)
            if (valid) {
                sample_idx = pid / wb.n_pixels;         // slot e_lane opens with sample e_lane
                depth = p.depth; flags = 0u;
                wf_open_sample(cam, p, px, py, sample_idx, rng, ray);
                w_rng = true;
                float4 a, b;
                ld_rows2(wb.res + (size_t)pid * 2u, a, b);
                DHit hc;
                hc.prim = __float_as_int(a.x);
                const uint32_t meta = f2u(a.y);
                hc.face = meta & 1u; hc.object = meta >> 8; hc.t = a.z;
                hc.pos = ray.o + ray.d * hc.t;                                         // main.rs:210 / 304
                hc.normal = mk3(b); hc.uv.x = a.w; hc.uv.y = b.w;
                flags |= F_FRESH;                       // acc = 0, T = 1: the first pass that changes them stores them
                if (hc.prim < 0) do_finish = true;
                else { h = hc; h_dir = ray.d; h_dir_orig = ray.d; h_rayface = ray.face; w_hit = true; w_dirs = true; do_level = true; }
            }
        } else if (seg == WF_SEG_PRIMARY || seg == WF_SEG_BOUNCE) {
            if (valid) {
                float4 a, b;
                DHit hc;
                load_path_hit(a, b, hc);
                const bool hit = hc.prim >= 0;
                if (seg == WF_SEG_PRIMARY) {                                           // main.rs:1150-1155
                    // a fresh sample: acc = 0, T = 1 (not stored until they change: F_FRESH)
                    flags |= F_FRESH;
                    if (!hit) do_finish = true;
                    else { h = hc; h_dir = ray.d; h_dir_orig = ray.d; h_rayface = ray.face; w_hit = true; w_dirs = true; do_level = true; }
                } else {
                    const uint32_t ray_type = (flags >> F_RAYTYPE_SHIFT) & 3u;
                    if (!hit) {
                        if (ray_type == 2u) { if (!(flags & F_FRESH)) acc = mk3(pm.ld(ROW_ACC)); do_finish = true; }   // main.rs:606-608
                        else {                                                                               // main.rs:572-574 / 591-593
                            flags = (flags & ~((3u << F_PURPOSE_SHIFT) | F_PRE | F_BLACK)) | (SH_FINAL << F_PURPOSE_SHIFT);
                            do_shade_begin = true;
                        }
                    } else arrive(hc);
                }
            }
        } else if (seg == WF_SEG_SHB) {
            do_shade_begin = valid;
        } else if (seg == WF_SEG_REFR) {                                               // main.rs:371-402
            if (valid) {
                pm.get_ray(ray);
                float4 ra, rb;
                ld_rows2(wb.res + (size_t)pid * 2u, ra, rb);
                refract_step(ra, rb);
            }
        } else {   // WF_SEG_SHADE: the shadow rays of the current light chunk are back (main.rs:435-461)
            if (valid) {
                const MatEval mat = material_approx(sc.materials, h.object, h.uv);
                const f3 nadj = nadj_in;                   // adjust_normal(mat, h.normal), kept by get_shade's entry
                const SpecConst spc = spec_const(mat);
                const uint32_t li0 = (flags >> F_LI0_SHIFT) & 0xfffu, need = (flags >> F_NEED_SHIFT) & 15u;
                const uint32_t purpose = (flags >> F_PURPOSE_SHIFT) & 3u;
                // fused levels: the path ray that travelled with these shadow rays
                const bool pre = FUSED && (flags & F_PRE);
                const uint32_t ray_type = (flags >> F_RAYTYPE_SHIFT) & 3u;
                float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
                DHit hc;
                hc.prim = -1; hc.face = 0; hc.object = 0; hc.t = 0.f; hc.pos = hc.normal = mk3(0.f, 0.f, 0.f); hc.uv.x = hc.uv.y = 0.f;
                if (pre) load_path_hit(ra, rb, hc);
                // a bounce ray that left the scene ends the sample with get_shade of THIS hit seen along the scattered
                // direction (main.rs:572-574 / 591-593): same shadow rays, a second specular term
                const bool want_final = pre && ray_type != 2u && hc.prim < 0;
                f3 shade2 = mk3(0.f, 0.f, 0.f);
                // view direction of the Phong probe (main.rs:449-456): the hit's ray direction — the scattered one only for
                // the get_shade that closes a sample after a missed bounce (hit.ray.direction is replaced at main.rs:552)
                const f3 view = purpose == SH_FINAL ? -h_dir : -h_dir_orig;
                if (flags & F_PARTIAL) shade = mk3(pm.ld(ROW_HI_POS));
                // the four shadow results of the path: one 32-byte sector, read before the loop
                float4 sr01, sr23;
                ld_rows2(reinterpret_cast<const float4*>(wb.sres) + (size_t)pid * 2u, sr01, sr23);
#pragma unroll kShadeSumUnroll
                for (uint32_t s = 0; s < 4u; ++s) {
                    if (!((need >> s) & 1u)) continue;
                    DirLight L;
                    const float4 sd = pm.get_shadow_dir(s);                            // {-L.dir, angular} kept by get_shade's entry
                    float dist_light;
                    approx_light_cached(sc.lights[li0 + s], h.pos, -mk3(sd), sd.w, L, dist_light);
                    const float2 sr = s == 0u ? make_float2(sr01.x, sr01.y) : s == 1u ? make_float2(sr01.z, sr01.w)
                                    : s == 2u ? make_float2(sr23.x, sr23.y) : make_float2(sr23.z, sr23.w);
                    bool occluded = false;
                    if (__float_as_int(sr.x) >= 0) {
                        if (L.has_origin) {
                            const f3 occ = h.pos + (-L.dir) * sr.y;                    // the shadow ray's hit point
                            if (distance(h.pos, occ) < (dist_light >= 0.0f ? dist_light : distance(h.pos, L.origin))) occluded = true;
                        } else occluded = true;
                    }
                    if (!occluded) {
                        const f3 ldir = -L.dir;
                        const f3 diffuse = get_diffuse(mat, nadj, ldir) * L.color;         // main.rs:458
                        const f3 specular = get_specular(mat, spc, nadj, view, ldir) * L.color; // main.rs:459
                        shade = shade + diffuse * (1.0f - mat.shiness) + specular * mat.shiness;  // main.rs:461
                        if (FUSED && want_final) {
                            const f3 specular2 = get_specular(mat, spc, nadj, -h_dir, ldir) * L.color;
                            shade2 = shade2 + diffuse * (1.0f - mat.shiness) + specular2 * mat.shiness;
                        }
                    }
                }
                if (!FUSED && li0 + 4u < sc.n_lights) {     // more lights: next chunk
                    pm.sv(ROW_HI_POS, make_float4(shade.x, shade.y, shade.z, 0.f));
                    flags = (flags & ~((0xfffu << F_LI0_SHIFT) | (15u << F_NEED_SHIFT))) | ((li0 + 4u) << F_LI0_SHIFT) | F_PARTIAL;
                    do_shade_begin = true;
                } else {
                    w_acc = true;
                    flags &= ~F_FRESH;
                    if (purpose == SH_FINAL) {
                        acc = acc + T * shade;
                        do_finish = true;
                    } else {
                        if (purpose == SH_NEXT_MIX) {
                            // mix(get_shade(next), x*probe, 0.5) = a + (x*probe - a)*0.5   (main.rs:571, 590)
                            acc = acc + T * (shade - shade * 0.5f);
                            T = T * (pend * 0.5f);
                        } else {
                            // (x + get_shade(next)) * decay^distance   (main.rs:605)
                            acc = acc + T * (shade * pend.x);
                            T = T * pend.x;
                        }
                        a_shade = shade; flags |= F_A_KNOWN; depth -= 1;
                        if (!FUSED) do_level = true;
                        else if (depth <= 0) { acc = acc + T * a_shade; do_finish = true; }       // main.rs:525-527
                        else if (flags & F_BLACK) do_finish = true;                               // main.rs:559-561, 366-368
                        else if (ray_type == 2u) refract_step(ra, rb);                            // first inside ray of get_refract
                        else if (hc.prim < 0) { acc = acc + T * shade2; do_finish = true; }       // main.rs:572-574 / 591-593
                        else arrive(hc);
                    }
                }
            }
        }

        // ---- top of distributed_ray_trace for the current hit (main.rs:521-554), then the bounce ray ---------------------
        if (seg == WF_SEG_PRIMARY || seg == WF_SEG_PRIMARY0 || seg == WF_SEG_SHADE || (FUSED && seg == WF_SEG_BOUNCE)) {
            if (do_level) {
                if (!pre_level && depth <= 0) {
                    if (flags & F_A_KNOWN) { acc = acc + T * a_shade; do_finish = true; }    // main.rs:525-527
                    else {                                                                    // depth 0 at the primary hit
                        flags = (flags & ~(3u << F_PURPOSE_SHIFT)) | (SH_FINAL << F_PURPOSE_SHIFT);
                        out = OUT_SHB;
                    }
                } else {
                    const MatEval mat = material_approx(sc.materials, h.object, h.uv);
                    const float w0 = (1.0f - mat.shiness) * (1.0f - mat.transparency);
                    const float w1 = mat.shiness * (1.0f - mat.transparency);
                    const float w2 = mat.transparency;
                    // weighted_select, main.rs:652-666
                    const float wsum = (w0 + w1) + w2;
                    const float rsel = rng_range(rng, 0.0f, wsum);
                    float accum = 0.0f;
                    accum += w0;
                    uint32_t ray_type;
                    if (rsel < accum) ray_type = 0u;
                    else { accum += w1; ray_type = rsel < accum ? 1u : 2u; }
                    // scatter_hit, main.rs:539-554
                    const f3 base_dir = ray_type == 0u ? -h.normal : h_dir;
                    const float exponent = ray_type == 0u ? 1.0f : mat.smoothness;
                    // (a diffuse bounce scatters with exponent 1: powf(x, 1) = x exactly, as the IEEE pow of the reference)
                    const float su = 1.0f - rng_range(rng, 0.0f, 1.0f);
                    const float phi = nl_acosf(exponent == 1.0f ? su : nl_powf(su, exponent));
                    const float theta = rng_range(rng, -kPi, kPi);
                    const quat from_z = from_arc(mk3(0.0f, 0.0f, 1.0f), normalize(base_dir));
                    const float2 scp = nl_sincosf(phi), sct = nl_sincosf(theta);
                    const f3 new_dir = rotate(from_z, mk3(scp.x * sct.y, scp.x * sct.x, scp.y));
                    h_dir_orig = h_dir;
                    h_dir = new_dir;                                                   // main.rs:552
                    w_dirs = true; w_rng = true;
                    flags = (flags & ~((3u << F_RAYTYPE_SHIFT) | F_TIR)) | (ray_type << F_RAYTYPE_SHIFT);
                    const float cosine = -dot(h.normal, h_dir);                        // main.rs:559 / 578 / 597
                    // pre_level: the sample still owes the get_shade of this hit; what the level decides is kept in the flags
                    if (cosine <= 0.0f) { if (pre_level) flags |= F_BLACK; else do_finish = true; }   // black
                    else if (ray_type == 2u) {                                         // get_refract, main.rs:354-368
                        f3 rin;
                        if (refract_dir(h.normal, h_dir, mat.refraction_index, rin)) {
                            ray.o = h.pos; ray.d = normalize(rin); ray.face = kBack; ray.ex_prim = h.prim; ray.ex_face = kFront;
                            if (pre_level) { pre_ray = true; flags |= F_PRE; } else out = OUT_REFR;
                        } else if (pre_level) flags |= F_BLACK;                        // Trapped
                        else do_finish = true;
                    } else {
                        ray = make_reflect(h.pos, h.normal, h_dir, h_rayface, h.prim, h.face);   // main.rs:563 / 582
                        if (pre_level) { pre_ray = true; flags |= F_PRE; } else out = OUT_BOUNCE;
                    }
                }
            }
        }

        // ---- get_shade entry (main.rs:408-433): shadow rays of the next chunk of up to 4 lights ------------------------
        uint32_t n_shadow = 0u;
        if (seg == WF_SEG_BOUNCE || seg == WF_SEG_SHB || seg == WF_SEG_SHADE) {
            if (do_shade_begin) {
                const MatEval mat = material_approx(sc.materials, h.object, h.uv);
                const f3 nadj = adjust_normal(mat, h.normal);
                nadj_out = nadj; w_nadj = true;                                // the consuming pass does not rotate it again
                // (fused levels run scenes of one light chunk: get_shade always starts at light 0)
                uint32_t li0 = (seg == WF_SEG_SHADE && !FUSED) ? ((flags >> F_LI0_SHIFT) & 0xfffu) : 0u;
                if (seg != WF_SEG_SHADE || FUSED) flags &= ~F_PARTIAL;
                uint32_t need = 0u;
                // chunks without any shadow ray are skipped here (their lights contribute nothing, main.rs:418-421)
                for (;;) {
#pragma unroll kShadeBeginUnroll
                    for (uint32_t s = 0; s < 4u && li0 + s < sc.n_lights; ++s) {
                        DirLight L;
                        if (!approx_light(sc.lights[li0 + s], h.pos, L)) continue;
                        const float cosine = -dot(L.dir, nadj);                        // main.rs:420
                        if (cosine <= 0.0f) continue;
                        pm.put_shadow_dir(s, -L.dir, L.angular);                       // main.rs:423-431
                        need |= 1u << s;
                    }
                    if (need || li0 + 4u >= sc.n_lights) break;
                    li0 += 4u;
                }
#if WF_REQ_HP
                // whole sectors: {hit, dir 0} {dir 1, dir 2} {dir 3, -}
                if (need) { pm.put_shadow_origin(h.pos, h.prim); if (!(need & 1u)) pm.put_shadow_dir(0, mk3(0.f, 0.f, 0.f)); }
                if (need & 0x6u) { if (!(need & 2u)) pm.put_shadow_dir(1, mk3(0.f, 0.f, 0.f)); if (!(need & 4u)) pm.put_shadow_dir(2, mk3(0.f, 0.f, 0.f)); }
                if (need & 0x8u) pm.put_req_pad();
#else
                if (need & 0x3u) { if (!(need & 1u)) pm.put_shadow_dir(0, mk3(0.f, 0.f, 0.f)); if (!(need & 2u)) pm.put_shadow_dir(1, mk3(0.f, 0.f, 0.f)); }   // whole sectors
                if (need & 0xcu) { if (!(need & 4u)) pm.put_shadow_dir(2, mk3(0.f, 0.f, 0.f)); if (!(need & 8u)) pm.put_shadow_dir(3, mk3(0.f, 0.f, 0.f)); }
#endif
                flags = (flags & ~((0xfffu << F_LI0_SHIFT) | (15u << F_NEED_SHIFT))) | (li0 << F_LI0_SHIFT) | (need << F_NEED_SHIFT);
                n_shadow = (uint32_t)__popc(need);
                out = OUT_SHADE;      // with need == 0 the path passes through the SHADE segment of the next round with no cast
            }
        }

        // ---- close the sample and open the next one (main.rs:1133-1166; photon.rs:28-33) ------------------------------------
        if (__any_sync(kFullMask, do_finish)) {
            if (do_finish) {
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                const uint32_t e_lane = pid / wb.n_pixels;
                if (seg != WF_SEG_INIT) {
                    // (a slot's first sample starts its accumulator: no INIT pass zeroed it when round 0 ran fused)
                    const bool first = wb.fused_primary != 0u && sample_idx == e_lane;
                    if (!first) sum = wb.sums[pid];
                    const bool accepted = is_normal_f32(acc.x) && is_normal_f32(acc.y) && is_normal_f32(acc.z);   // main.rs:1157-1160
                    if (accepted) {
                        sum.x += acc.x; sum.y += acc.y; sum.z += acc.z; sum.w += 1.0f;            // photon.rs:30-31
                        n_samples += 1ull;
                    }
                    if (accepted || first) wb.sums[pid] = sum;
                    sample_idx += wb.epar;
                } else {
                    sample_idx = e_lane;
                    wb.sums[pid] = sum;
                }
                if (sample_idx >= n_epochs) out = OUT_RETIRE;
                else {
                    acc = mk3(0.f, 0.f, 0.f); T = mk3(1.f, 1.f, 1.f); depth = p.depth; flags = 0u;
                    w_acc = false; w_hit = false; w_dirs = false; w_pend = false; w_nadj = false;
                    wf_open_sample(cam, p, px, py, sample_idx, rng, ray);   // Camera::shoot_focus, main.rs:101-127
                    w_rng = true;
                    out = OUT_PRIMARY;
                }
            }
        }

        if (seg != WF_SEG_INIT && WF_LOGIC_PREFETCH == 1 && valid_ahead) {   // (2: only the path id travels ahead)
            const float4* st_a = wb.st + (size_t)pid_ahead * kStateRows;
            const float4* rq_a = wb.req + (size_t)pid_ahead * WF_REQ_ROWS;
            prefetch_l2(st_a + ROW_CTRL);                                             // + ROW_RNG
            if (seg != WF_SEG_PRIMARY) { prefetch_l2(st_a + ROW_HPOS); prefetch_l2(st_a + ROW_HDIR); prefetch_l2(st_a + ROW_PEND); }
            if (seg == WF_SEG_SHADE) {
                prefetch_l2(st_a + ROW_ACC);                                          // + ROW_T
                prefetch_l2(wb.sres + (size_t)pid_ahead * 4u);
                prefetch_l2(rq_a + req_shadow_row(0)); prefetch_l2(rq_a + req_shadow_row(2));
            }
            if (seg != WF_SEG_SHADE || FUSED) { prefetch_l2(rq_a + REQ_O); prefetch_l2(wb.res + (size_t)pid_ahead * 2u); }
        }
        // ---- store what changed and append the path to the next round's queues -------------------------------------------
        if (valid && out != OUT_RETIRE) {
            const float4 ctrl = make_float4(u2f(flags), u2f(((uint32_t)depth & 0xffffu) | (rng.draws << 16)), u2f(px | (py << 16)), u2f(sample_idx));
            // (the INIT pass streams over consecutive paths: measured faster with 128-bit stores, 8.5 vs 12.4 ms per batch)
            constexpr bool kW = seg != WF_SEG_INIT;   // (the INIT pass streams over consecutive paths)
            const float4 rng_row = make_float4(u2f(rng.b[0]), u2f(rng.b[1]), u2f(rng.b[2]), u2f(rng.b[3]));
            if (kW) { if (w_rng) pm.sv2(ROW_CTRL, ctrl, rng_row); else pm.sv(ROW_CTRL, ctrl); }
            else { pm.sv(ROW_CTRL, ctrl); if (w_rng) pm.sv(ROW_RNG, rng_row); }   // (merged branches of 128-bit stores were scalarised)
            if (w_acc) pm.sv2(ROW_ACC, make_float4(acc.x, acc.y, acc.z, 0.f), make_float4(T.x, T.y, T.z, 0.f));
            if (w_pend && w_nadj) pm.sv2(ROW_PEND, make_float4(pend.x, pend.y, pend.z, rf_travel), make_float4(nadj_out.x, nadj_out.y, nadj_out.z, 0.0f));
            else if (w_pend) pm.sv(ROW_PEND, make_float4(pend.x, pend.y, pend.z, rf_travel));
            else if (w_nadj) pm.sv(ROW_NADJ, make_float4(nadj_out.x, nadj_out.y, nadj_out.z, 0.0f));
            if (w_hit)
                pm.sv2(ROW_HPOS, make_float4(h.pos.x, h.pos.y, h.pos.z, __int_as_float(h.prim)),
                       make_float4(h.normal.x, h.normal.y, h.normal.z, u2f(h.face | (h_rayface << 1) | (h.object << 8))));
            if (w_dirs)
                pm.sv2(ROW_HDIR, make_float4(h_dir.x, h_dir.y, h_dir.z, h.uv.x), make_float4(h_dir_orig.x, h_dir_orig.y, h_dir_orig.z, h.uv.y));
            if (out == OUT_PRIMARY || out == OUT_BOUNCE || out == OUT_REFR || pre_ray) pm.template put_ray<kW>(ray);
        }
        // ---- route: every reservation of this chunk (5 queues, the 5 cast work lists, the retired counter) is one
        // atomic issued by a different lane, so the warp pays one round trip to L2 instead of eleven
        {
            const bool path_ray = out == OUT_PRIMARY || out == OUT_BOUNCE || out == OUT_REFR;
            const uint32_t need_out = (valid && out == OUT_SHADE) ? ((flags >> F_NEED_SHIFT) & 15u) : 0u;
            const int my_seg = out == OUT_PRIMARY ? WF_SEG_PRIMARY : out == OUT_BOUNCE ? WF_SEG_BOUNCE : out == OUT_REFR ? WF_SEG_REFR
                             : out == OUT_SHADE ? WF_SEG_SHADE : out == OUT_SHB ? WF_SEG_SHB : 0;   // 0 = none (INIT is never a target)
            unsigned m_seg[WF_SEG_COUNT];
#pragma unroll
            for (int sgi = 1; sgi < WF_SEG_COUNT; ++sgi) m_seg[sgi] = __ballot_sync(kFullMask, valid && my_seg == sgi);
            const unsigned m_ret = __ballot_sync(kFullMask, valid && out == OUT_RETIRE);
            // cast work: slot 0 = the path ray (also the one that travels with shadow rays), slots 1..4 = shadow rays
            unsigned m_work[WF_WORK_PER_PATH];
            m_work[0] = __ballot_sync(kFullMask, valid && (path_ray || (out == OUT_SHADE && pre_ray)));
#pragma unroll
            for (uint32_t sl = 0; sl < 4u; ++sl) m_work[sl + 1u] = __ballot_sync(kFullMask, ((need_out >> sl) & 1u) != 0u);
            uint32_t my_n = 0u;
            uint32_t* my_addr = nullptr;
#pragma unroll
            for (int sgi = 1; sgi < WF_SEG_COUNT; ++sgi)
                if ((int)lane == sgi) { my_n = (uint32_t)__popc(m_seg[sgi]); my_addr = &wb.ctl->c[nbuf].seg[sgi]; }
            if (lane == (uint32_t)WF_SEG_COUNT) { my_n = (uint32_t)__popc(m_ret); my_addr = &wb.ctl->retired; }
#pragma unroll
            for (uint32_t k = 0; k < WF_WORK_PER_PATH; ++k)
                if (lane == 8u + k) { my_n = (uint32_t)__popc(m_work[k]); my_addr = &wb.ctl->c[nbuf].work[k]; }
            uint32_t my_base = 0u;
            if (my_n) my_base = atomicAdd(my_addr, my_n);
            uint32_t seg_base = 0u;
            unsigned seg_mask = 0u;
#pragma unroll
            for (int sgi = 1; sgi < WF_SEG_COUNT; ++sgi) {
                const uint32_t bse = __shfl_sync(kFullMask, my_base, sgi);
                if (my_seg == sgi) { seg_base = bse; seg_mask = m_seg[sgi]; }
            }
            const uint32_t lt = (1u << lane) - 1u;
            if (valid && my_seg != 0)
                wb.q[((size_t)nbuf * WF_SEG_COUNT + my_seg) * wb.n + seg_base + (uint32_t)__popc(seg_mask & lt)] = pid;
            uint32_t* w = wb.work + (size_t)nbuf * WF_WORK_PER_PATH * wb.n;
#pragma unroll
            for (uint32_t k = 0; k < WF_WORK_PER_PATH; ++k) {
                const uint32_t bse = __shfl_sync(kFullMask, my_base, 8 + (int)k);
                if ((m_work[k] >> lane) & 1u) w[(size_t)k * wb.n + bse + (uint32_t)__popc(m_work[k] & lt)] = pid;
            }
        }
    }
    if (cnt) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_samples += __shfl_xor_sync(0xffffffffu, n_samples, o);
        if (lane == 0u && n_samples) atomicAdd(&cnt->samples, n_samples);
    }
}

// accum[pixel] += the PhotonAccumulators of the pixel's slots, in slot order (photon.rs:28-33)
__global__ void wf_combine_kernel(const WfBuffers wb, const DParams p, float4* __restrict__ accum) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= wb.n_pixels) return;
    float4 s = wb.sums[pix];
    for (uint32_t e = 1; e < wb.epar; ++e) {
        const float4 v = wb.sums[(size_t)e * wb.n_pixels + pix];      // (dense per slot: coalesced over the pixels of a warp)
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const size_t at = (size_t)wf_frame_row(p, pix / p.width) * p.width + pix % p.width;
    float4 v = accum[at];
    v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    accum[at] = v;
}

// the condition of the round loop's WHILE node (launch_distributed_wavefront): go on while a slot has samples left
__global__ void wf_loop_kernel(cudaGraphConditionalHandle handle, const WfControl* __restrict__ ctl, uint32_t n_paths,
                               DCounters* __restrict__ cnt, uint32_t launches) {
    const bool more = ctl->retired < n_paths;
    if (cnt) { cnt->rounds += 2ull; cnt->launches += (unsigned long long)launches; }
    // (every path retires after <= epochs * (depth + 1) * (14 + lights) rounds; the cap only guards a corrupted counter)
    cudaGraphSetConditional(handle, (more && (!cnt || cnt->rounds < (1ull << 22))) ? 1u : 0u);
}

// ---- host side -----------------------------------------------------------------------------------------------
size_t wf_workspace_bytes_per_path() {
    return (size_t)kStateRows * 16 + WF_REQ_ROWS * 16 + 32 + 32 + 2 * WF_SEG_COUNT * 4 + 2 * WF_WORK_PER_PATH * 4 + 16;
}
size_t wf_workspace_bytes(uint32_t n_paths) {
    return sizeof(WfControl) + 256 + (size_t)n_paths * wf_workspace_bytes_per_path() + 16 * 256;   // (every array starts on a 256-byte boundary)
}

// Epochs rendered at once: every epoch of a batch has its own path slot per pixel, so no slot ever renders two samples
// in sequence.  (Measured on B200, 3840x2160 x 32 epochs: 16 in flight 536 ms, 8: 628 ms, 4: 693 ms - slots that chain
// samples drift apart, the queues lose their path order and the shading kernels their DRAM locality.)
uint32_t wf_epochs_in_flight(uint32_t width, uint32_t height, uint32_t epoch_count, size_t hbm_bytes) {
    // a function of the FRAME (not of the rendered row band), so that bands are bitwise the same rows of the full frame
    const unsigned long long px = (unsigned long long)width * height;
    if (const char* env = getenv("B200RT_WF_EPAR")) {            // tuning override
        const unsigned long long v = strtoull(env, nullptr, 10);
        if (v >= 1) return (uint32_t)(v > epoch_count ? (epoch_count ? epoch_count : 1u) : v);
    }
    const unsigned long long budget = (unsigned long long)(0.4 * (double)hbm_bytes);   // path state + queues
    unsigned long long e = budget / ((px ? px : 1ull) * wf_workspace_bytes_per_path());
    if (e < 1ull) e = 1ull;
    if (e > 64ull) e = 64ull;
    if (e > epoch_count) e = epoch_count ? epoch_count : 1u;
    return (uint32_t)e;
}

static WfBuffers wf_carve(void* workspace, uint32_t n_paths, uint32_t n_pixels, uint32_t epar) {
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    unsigned char* b = static_cast<unsigned char*>(workspace);
    const size_t n = n_paths;
    WfBuffers wb;
    size_t off = 0;
    wb.ctl = reinterpret_cast<WfControl*>(b + off);            off = align(off + sizeof(WfControl));
    wb.st = reinterpret_cast<float4*>(b + off);                off = align(off + n * kStateRows * 16);
    wb.req = reinterpret_cast<float4*>(b + off);               off = align(off + n * WF_REQ_ROWS * 16);
    wb.res = reinterpret_cast<float4*>(b + off);               off = align(off + n * 32);
    wb.sres = reinterpret_cast<float2*>(b + off);              off = align(off + n * 32);
    wb.q = reinterpret_cast<uint32_t*>(b + off);               off = align(off + n * 2 * WF_SEG_COUNT * 4);
    wb.work = reinterpret_cast<uint32_t*>(b + off);            off = align(off + n * 2 * WF_WORK_PER_PATH * 4);
    wb.sums = reinterpret_cast<float4*>(b + off);              off = align(off + n * 16);
    wb.n = n_paths; wb.n_pixels = n_pixels; wb.epar = epar; wb.fused_primary = 0u;
    return wb;
}

cudaError_t launch_distributed_wavefront(const DScene& sc, const DCamera& cam, const DParams& p, float* d_accum,
                                         DCounters* d_cnt, void* workspace, uint32_t n_paths, uint32_t epar, int sm_count,
                                         uint32_t* h_pinned_retired, cudaEvent_t ev_poll, cudaStream_t stream,
                                         uint32_t* rounds_out, uint32_t* launches_out, WfKernelTiming* timing) {
    const uint32_t n_pixels = p.width * p.row_count;
    WfBuffers wb = wf_carve(workspace, n_paths, n_pixels, epar);
    cudaError_t e = cudaMemsetAsync(wb.ctl, 0, sizeof(WfControl), stream);
    if (e != cudaSuccess) return e;
    const int cast_blocks = sm_count * WF_CAST_MIN_BLOCKS;
    // small scenes: phase 1 and phase 2 of the cast as two kernels (B200RT_WF_FUSED_CAST=1 keeps them fused: tuning)
    const uint32_t n_tiles = sc.n_tris_padded / kTileTris;
    // B200RT_WF_CAST = rl (default: rays in lanes, tiles by TMA for larger scenes) | fused (warp-transposed): measurement
    const char* cast_env = getenv("B200RT_WF_CAST");
    const std::string cast_sel = cast_env ? cast_env : "";
    const bool bvh = p.cast_mode == B200RT_CAST_BVH && sc.bvh_n_nodes != 0u;      // the acceleration structure (rt_bvh.cuh)
    const bool rays_in_lanes = bvh || (n_tiles == 1 && (cast_sel.empty() || cast_sel == "rl"));   // (bvh: the same round-0 flow)
    const bool rays_in_lanes_tiled = !bvh && n_tiles > 1 && (cast_sel.empty() || cast_sel == "rl");
    // fused levels (one light chunk): a hit's shadow rays and the next level's ray are cast in the same round, and one
    // kernel pass per level consumes both (B200RT_WF_FUSED_LEVELS=0: one pass per cast, as for scenes of > 4 lights)
    const char* fused_env = getenv("B200RT_WF_FUSED_LEVELS");
    const bool fused = sc.n_lights <= 4u && !(fused_env && fused_env[0] == '0');
    auto logic_blocks = [&](int min_blocks) { return sm_count * min_blocks; };
    // Round 0.  With a rays-in-lanes cast the first camera ray of every slot is generated inside the cast and consumed by
    // the PRIMARY0 pass (path id = index): no INIT pass, no request rows, no work list (B200RT_WF_FUSED_PRIMARY=0 keeps
    // the INIT pass, as the warp-transposed cast does).  Otherwise: every slot opens its first sample in an INIT pass.
    const char* fp_env = getenv("B200RT_WF_FUSED_PRIMARY");
    const bool fused_primary = (rays_in_lanes || rays_in_lanes_tiled) && !(fp_env && fp_env[0] == '0');
    wb.fused_primary = fused_primary ? 1u : 0u;
    if (!fused_primary)
        wf_logic_kernel<WF_SEG_INIT, false><<<logic_blocks(LogicCfg<WF_SEG_INIT>::kMinBlocks), 256, 0, stream>>>(sc, cam, p, wb, 1u, d_cnt);
    // Scenes of many tiles: rounds of at most split_below 128-ray blocks take the split form of the cast (a block per CTA,
    // the tile range over its warps).  The unsplit kernel fills the GPU with sm_count x 5 x 4 blocks; below half of that the
    // split form is faster (B200RT_WF_SPLIT_BELOW overrides, 0 = never).
    uint32_t split_below = 0u;
    if (rays_in_lanes_tiled) {
        split_below = (uint32_t)sm_count * WF_CAST_RL_TILED_MIN_BLOCKS * 2u;
        if (const char* sb = getenv("B200RT_WF_SPLIT_BELOW")) split_below = (uint32_t)strtoul(sb, nullptr, 10);
        if (split_below) {
            e = cudaFuncSetAttribute(wf_cast_rl_tiled_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RlSplitShared));
            if (e != cudaSuccess) return e;
        }
    }
    // one round on `stream`: the cast of the rays requested in the round before, then the passes that consume it
    auto launch_round = [&](uint32_t round, uint32_t buf) {
        if (bvh) {
            if (round == 0u && fused_primary) wf_cast_bvh_primary_kernel<<<sm_count * 2 * WF_BVH_MIN_BLOCKS, 128, 0, stream>>>(sc, cam, p, wb, buf, d_cnt);
            else wf_cast_bvh_kernel<<<sm_count * 2 * WF_BVH_MIN_BLOCKS, 128, 0, stream>>>(sc, wb, buf, d_cnt);
        } else if (round == 0u && fused_primary) {
            if (rays_in_lanes) wf_cast_rl_primary_kernel<<<sm_count * WF_CAST_RL_MIN_BLOCKS, kRlThreads, 0, stream>>>(sc, *sc.h_tile0, cam, p, wb, buf, d_cnt);
            else wf_cast_rl_tiled_primary_kernel<<<sm_count * WF_CAST_RL_TILED_MIN_BLOCKS, kRlThreads, 0, stream>>>(sc, cam, p, wb, buf, d_cnt);
        } else if (rays_in_lanes) {
            wf_cast_rl_kernel<<<sm_count * WF_CAST_RL_MIN_BLOCKS, kRlThreads, 0, stream>>>(sc, *sc.h_tile0, wb, buf, d_cnt);
        } else if (rays_in_lanes_tiled) {
            wf_cast_rl_tiled_kernel<<<sm_count * WF_CAST_RL_TILED_MIN_BLOCKS, kRlThreads, 0, stream>>>(sc, wb, buf, d_cnt, split_below);
            if (split_below) wf_cast_rl_tiled_split_kernel<<<sm_count * 4, kRlThreads, sizeof(RlSplitShared), stream>>>(sc, wb, buf, d_cnt, split_below);
        } else {
            wf_cast_kernel<<<cast_blocks, 128, 0, stream>>>(sc, wb, buf, d_cnt);
        }
    };
    auto launch_round_logic = [&](uint32_t round, uint32_t buf) {
        if (round == 0u && fused_primary)
            wf_logic_kernel<WF_SEG_PRIMARY0, false><<<logic_blocks(LogicCfg<WF_SEG_PRIMARY0>::kMinBlocks), 256, 0, stream>>>(sc, cam, p, wb, buf, d_cnt);
        wf_logic_kernel<WF_SEG_PRIMARY, false><<<logic_blocks(LogicCfg<WF_SEG_PRIMARY>::kMinBlocks), 256, 0, stream>>>(sc, cam, p, wb, buf, d_cnt);
        if (fused) {
            wf_logic_kernel<WF_SEG_SHADE, true><<<logic_blocks(LogicCfg<WF_SEG_SHADE, true>::kMinBlocks), LogicCfg<WF_SEG_SHADE, true>::kThreads, 0, stream>>>(sc, cam, p, wb, buf, d_cnt);
            wf_logic_kernel<WF_SEG_BOUNCE, true><<<logic_blocks(LogicCfg<WF_SEG_BOUNCE, true>::kMinBlocks), 256, 0, stream>>>(sc, cam, p, wb, buf, d_cnt);
        } else {
            wf_logic_kernel<WF_SEG_SHADE, false><<<logic_blocks(LogicCfg<WF_SEG_SHADE>::kMinBlocks), 256, 0, stream>>>(sc, cam, p, wb, buf, d_cnt);
            wf_logic_kernel<WF_SEG_BOUNCE, false><<<logic_blocks(LogicCfg<WF_SEG_BOUNCE>::kMinBlocks), 256, 0, stream>>>(sc, cam, p, wb, buf, d_cnt);
        }
        wf_logic_kernel<WF_SEG_REFR, false><<<logic_blocks(LogicCfg<WF_SEG_REFR>::kMinBlocks), 256, 0, stream>>>(sc, cam, p, wb, buf, d_cnt);
        if (p.depth <= 0)   // get_shade of depth-0 primary hits (the only source of this segment)
            wf_logic_kernel<WF_SEG_SHB, false><<<logic_blocks(LogicCfg<WF_SEG_SHB>::kMinBlocks), 256, 0, stream>>>(sc, cam, p, wb, buf, d_cnt);
    };
    const uint32_t launches_per_round = (p.depth <= 0 ? 6u : 5u) + ((rays_in_lanes_tiled && split_below) ? 1u : 0u);

    // The round loop ON THE DEVICE: a CUDA graph whose WHILE node repeats {round on buffer 1, round on buffer 0, "any
    // path left?"} until every slot has retired (wf_loop_kernel sets the node's condition from the retired counter).  The
    // host enqueues round 0, the graph and the combine pass and returns: no polling, the *_device entry points are
    // asynchronous.  Not on the legacy default stream (it cannot be captured), not when the per-kernel timing of the
    // roofline pass is on, and B200RT_WF_GRAPH=0 keeps the host loop (measurement).
    const char* graph_env = getenv("B200RT_WF_GRAPH");
    const bool use_graph = timing == nullptr && stream != nullptr && stream != cudaStreamLegacy && !(graph_env && graph_env[0] == '0');
    bool round0_done = false;
    if (use_graph) {
        launch_round(0u, 0u);
        launch_round_logic(0u, 0u);
        round0_done = true;
        cudaGraph_t graph = nullptr, body = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaGraphConditionalHandle handle;
        cudaGraphNode_t node;
        cudaGraphNodeParams np = {};
        bool captured = false;
        e = cudaGraphCreate(&graph, 0);
        if (e == cudaSuccess) e = cudaGraphConditionalHandleCreate(&handle, graph, 1u, cudaGraphCondAssignDefault);
        if (e == cudaSuccess) {
            np.type = cudaGraphNodeTypeConditional;
            np.conditional.handle = handle;
            np.conditional.type = cudaGraphCondTypeWhile;
            np.conditional.size = 1;
            e = cudaGraphAddNode(&node, graph, nullptr, 0, &np);
        }
        if (e == cudaSuccess) {
            body = np.conditional.phGraph_out[0];
            e = cudaStreamBeginCaptureToGraph(stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        }
        if (e == cudaSuccess) {
            captured = true;
            launch_round(1u, 1u); launch_round_logic(1u, 1u);
            launch_round(2u, 0u); launch_round_logic(2u, 0u);
            wf_loop_kernel<<<1, 1, 0, stream>>>(handle, wb.ctl, n_paths, d_cnt, 2u * launches_per_round + 1u);
            cudaGraph_t out = nullptr;
            e = cudaStreamEndCapture(stream, &out);
        }
        if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
        if (e == cudaSuccess) e = cudaGraphLaunch(exec, stream);
        if (exec) cudaGraphExecDestroy(exec);      // (in flight: freed when the launch completes)
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            if (captured) return e;                // the stream's state is the capture's: report
            (void)cudaGetLastError();              // (no conditional-node support: the host loop below takes over after round 0)
        } else {
            wf_combine_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(wb, p, reinterpret_cast<float4*>(d_accum));
            if (rounds_out) *rounds_out = 1u;      // round 0; the loop's rounds and launches are counted on the device (DCounters)
            if (launches_out) *launches_out += 2u + launches_per_round;
            return cudaGetLastError();
        }
    }
    uint32_t round = round0_done ? 1u : 0u, buf = round0_done ? 1u : 0u;
    uint32_t group = 8;
    for (;;) {
        const uint32_t first_round_of_group = round;
        for (uint32_t g = 0; g < group; ++g, ++round, buf ^= 1u) {
            // optional per-kernel timing (bench.py's roofline pass): events around every cast launch
            cudaEvent_t ev_a = nullptr, ev_b = nullptr;
            if (timing) {
                while (timing->pool.size() < 2 * (size_t)(round - first_round_of_group + 1)) {
                    cudaEvent_t ev;
                    e = cudaEventCreate(&ev);
                    if (e != cudaSuccess) return e;
                    timing->pool.push_back(ev);
                }
                ev_a = timing->pool[2 * (round - first_round_of_group)];
                ev_b = timing->pool[2 * (round - first_round_of_group) + 1];
                cudaEventRecord(ev_a, stream);
            }
            launch_round(round, buf);
            if (timing) cudaEventRecord(ev_b, stream);
            launch_round_logic(round, buf);
        }
        e = cudaMemcpyAsync(h_pinned_retired, &wb.ctl->retired, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaEventRecord(ev_poll, stream);
        if (e != cudaSuccess) return e;
        e = cudaEventSynchronize(ev_poll);
        if (e != cudaSuccess) return e;
        if (timing) {
            for (uint32_t g = 0; g < round - first_round_of_group; ++g) {
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, timing->pool[2 * g], timing->pool[2 * g + 1]);
                if (first_round_of_group + g == 0u && fused_primary) timing->primary_ms += ms;   // a different kernel: camera rays + cast
                else { timing->cast_ms += ms; timing->cast_launches += 1; }
                if (g + 1 < round - first_round_of_group) {       // cast end -> next cast start = the logic kernels of the round
                    cudaEventElapsedTime(&ms, timing->pool[2 * g + 1], timing->pool[2 * g + 2]);
                    timing->logic_ms += ms;
                }
            }
        }
        if (*h_pinned_retired >= n_paths) break;
        if (group < 64) group *= 2;
        if (round > (1u << 21)) return cudaErrorLaunchTimeout;   // cannot happen: every path retires after <= epochs * (depth+1) * (14 + lights) rounds
    }
    wf_combine_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(wb, p, reinterpret_cast<float4*>(d_accum));
    if (rounds_out) *rounds_out = round;
    if (launches_out) *launches_out += 2u + round * ((p.depth <= 0 ? 6u : 5u));   // (INIT or PRIMARY0) + combine + per round
    return cudaGetLastError();
}

}  // namespace b200rt
