// rt_cast_rl.cuh — World::cast (main.rs:180-326) with RAYS IN LANES, for scenes of one tile (<= 64 triangles) and
// kernels in which every lane has a ray (wavefront rounds, b200rt_intersect).
//
// Every lane owns FOUR rays, packed as two FFMA2 pairs, and the CTA's 64 plain filter records sit in 4 KB of shared
// memory; each record is read once per warp as four broadcast LDS.128 and feeds 2 x 21 FFMA2 (128 ray x triangle pairs
// per record read).  Measured on B200 (tools/filter_bench.py, the filter loop alone): 53.5 % of the FP32 roofline in this
// form against 42 % for the warp-transposed form of rt_cast.cuh, whose FFMA2 read three live register pairs each.
// Keep / reject is the sign bit of max(min(e0,e1,e2,t,c) + A|r|, g - |nd|), shifted into a per-ray mask (no predicates,
// no ballots).  The packed ray operands are 64-bit values built ONCE per block of rays (P2, rt_cast.cuh): as float2
// arrays the compiler re-packed them from scattered registers before every use (77 MOVs per 168 FFMA2 issue slots).
// The loop is 143 instructions per 8 ray x triangle pairs: 84 FFMA2/FMUL2 (168 issue slots, 144 of them algorithmic)
// + 8 LDS.128 + 8 MUFU.RCP + 24 FMNMX(3) + 8 FADD + 8 SHF + 3: an instruction-mix ceiling of 63 % of the FP32 roofline.
// Phase 2 (certified select, spheres, attributes: rt_cast.cuh) then runs per lane for its four rays, which it reads
// back from a per-thread shared-memory slot (no dynamically indexed register arrays, no local memory).
//
// IO (a small struct, by value) connects the loop to its rays and results:
//     bool     load(uint32_t idx, DRay& r, uint32_t& tag)   ray of work index idx (idx < n_work); tag travels to store()
//     void     prefetch(uint32_t idx)                        hint: idx will be loaded by this lane in its next block
//     bool     want_attrs(uint32_t tag)                      false: only "is there a hit, and how far" is needed
//     void     store(uint32_t tag, const DHit& h)
#pragma once
#include "rt_cast.cuh"

namespace b200rt {

constexpr int kRlThreads = 128;                 // CTA size of the kernels built on cast_rays_in_lanes
constexpr uint32_t kRlNoRay = 0xffffffffu;      // tag of an empty ray slot (IO tags must not use it)

struct RlShared {                               // 24.5 KB per CTA
    float4 tile[4 * kTileTris];                 // plain filter records
    float4 ro[4][kRlThreads];                   // {origin, ray meta}   of ray j of thread t
    float4 rd[4][kRlThreads];                   // {direction, tag}
    uint2 mk[4][kRlThreads];                    // candidate mask
};

RT_DI uint32_t rl_pack_ray_meta(uint32_t face, int32_t ex_prim, uint32_t ex_face) {
    return face | (ex_face << 2) | ((uint32_t)(ex_prim + 1) << 4);
}

// CTA-collective (kRlThreads threads): casts rays [0, n_work) of `io`, 128 rays per warp and iteration, work split over
// the whole grid.  Call once per kernel; n_work may be 0.
template <bool PREFETCH, class IO>
RT_DI void cast_rays_in_lanes(const DScene& sc, IO io, const uint32_t n_work, RlShared& sh, CastStats& cs) {
    const uint32_t lane = threadIdx.x & 31u, tid = threadIdx.x;
    if (n_work == 0u) return;
    for (uint32_t i = threadIdx.x; i < 4u * kTileTris; i += blockDim.x) sh.tile[i] = sc.tri_filter_plain[i];
    __syncthreads();
    const P2 A2 = p2_bc(sc.filter_A);
    const float g = sc.filter_g;
    // exactly 1.0f, but opaque to ptxas: a packed multiply by it MATERIALISES each ray operand in its own aligned
    // register pair (a plain pack is coalesced with the LDG.128 destination quads and re-packed inside the loop)
    const P2 one2 = p2_bc(__fmaf_rn(sc.filter_g, 0.0f, 1.0f));
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const uint32_t stride = warps_total * 128u;
    // (n_work + stride may exceed 2^32: the block loop counts blocks, not indices)
    const uint32_t n_blocks = (n_work + 127u) / 128u;
    for (uint32_t blk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); blk < n_blocks; blk += warps_total) {
        const uint32_t base = blk * 128u;
        // this lane's four rays: work indices base + lane + 32 j
        P2 ox[2], oy[2], oz[2], dx[2], dy[2], dz[2], cf[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            DRay r[2];
            float c[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * k + h;
                const uint32_t idx = base + lane + 32u * (uint32_t)j;
                r[h].o = mk3(0.f, 0.f, 0.f); r[h].d = mk3(0.f, 0.f, 1.f); r[h].face = kFront; r[h].ex_prim = -1; r[h].ex_face = kFront;
                uint32_t tag = kRlNoRay;
                if (idx < n_work && !io.load(idx, r[h], tag)) tag = kRlNoRay;
                c[h] = r[h].face == kFront ? -kCullK : (r[h].face == kBack ? kCullK : 0.0f);
                sh.ro[j][tid] = make_float4(r[h].o.x, r[h].o.y, r[h].o.z, __uint_as_float(rl_pack_ray_meta(r[h].face, r[h].ex_prim, r[h].ex_face)));
                sh.rd[j][tid] = make_float4(r[h].d.x, r[h].d.y, r[h].d.z, __uint_as_float(tag));
            }
            ox[k] = p2_mul(p2_pack(r[0].o.x, r[1].o.x), one2); oy[k] = p2_mul(p2_pack(r[0].o.y, r[1].o.y), one2);
            oz[k] = p2_mul(p2_pack(r[0].o.z, r[1].o.z), one2);
            dx[k] = p2_mul(p2_pack(r[0].d.x, r[1].d.x), one2); dy[k] = p2_mul(p2_pack(r[0].d.y, r[1].d.y), one2);
            dz[k] = p2_mul(p2_pack(r[0].d.z, r[1].d.z), one2);
            cf[k] = p2_mul(p2_pack(c[0], c[1]), one2);
        }
        // phase 1: reject masks (bit set = rejected), triangle i in bit (31 - i) of its half
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint32_t rj0 = 0u, rj1 = 0u, rj2 = 0u, rj3 = 0u;
#pragma unroll 2
            for (int i = 0; i < 32; ++i) {
                const float4* q = sh.tile + 4 * (32 * half + i);
                const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const P2 nd = p2_fma(p2_bc(q0.z), dz[k], p2_fma(p2_bc(q0.y), dy[k], p2_mul(p2_bc(q0.x), dx[k])));
                    const P2 num = p2_fma(p2_bc(-q0.z), oz[k], p2_fma(p2_bc(-q0.y), oy[k], p2_fma(p2_bc(-q0.x), ox[k], p2_bc(q0.w))));
                    float nda, ndb;
                    p2_unpack(nd, nda, ndb);
                    const float ra = rcp_approx(nda), rb = rcp_approx(ndb);
                    const P2 t = p2_mul(num, p2_pack(ra, rb));
                    const P2 cull = p2_mul(nd, cf[k]);
                    const P2 px = p2_fma(t, dx[k], ox[k]), py = p2_fma(t, dy[k], oy[k]), pz = p2_fma(t, dz[k], oz[k]);
                    const P2 e0 = p2_fma(p2_bc(q1.z), pz, p2_fma(p2_bc(q1.y), py, p2_fma(p2_bc(q1.x), px, p2_bc(q1.w))));
                    const P2 e1 = p2_fma(p2_bc(q2.z), pz, p2_fma(p2_bc(q2.y), py, p2_fma(p2_bc(q2.x), px, p2_bc(q2.w))));
                    const P2 e2 = p2_fma(p2_bc(q3.z), pz, p2_fma(p2_bc(q3.y), py, p2_fma(p2_bc(q3.x), px, p2_bc(q3.w))));
                    float e0a, e0b, e1a, e1b, e2a, e2b, ta, tb, ca, cb;
                    p2_unpack(e0, e0a, e0b); p2_unpack(e1, e1a, e1b); p2_unpack(e2, e2a, e2b); p2_unpack(t, ta, tb); p2_unpack(cull, ca, cb);
                    const float ma = fminf(fminf(fminf(e0a, e1a), e2a), fminf(ta, ca));
                    const float mb = fminf(fminf(fminf(e0b, e1b), e2b), fminf(tb, cb));
                    const P2 ms = p2_fma(A2, p2_pack(fabsf(ra), fabsf(rb)), p2_pack(ma, mb));
                    float msa, msb;
                    p2_unpack(ms, msa, msb);
                    // keep iff ms >= 0 or |nd| < g  <=>  max(ms, g - |nd|) is not negative (NaN ms: the second operand decides)
                    const float ka = fmaxf(msa, g - fabsf(nda)), kb = fmaxf(msb, g - fabsf(ndb));
                    if (k == 0) { rj0 = __funnelshift_l(__float_as_uint(ka), rj0, 1); rj1 = __funnelshift_l(__float_as_uint(kb), rj1, 1); }
                    else        { rj2 = __funnelshift_l(__float_as_uint(ka), rj2, 1); rj3 = __funnelshift_l(__float_as_uint(kb), rj3, 1); }
                }
            }
            const uint32_t k0 = ~__brev(rj0), k1 = ~__brev(rj1), k2 = ~__brev(rj2), k3 = ~__brev(rj3);
            if (half == 0) { sh.mk[0][tid].x = k0; sh.mk[1][tid].x = k1; sh.mk[2][tid].x = k2; sh.mk[3][tid].x = k3; }
            else           { sh.mk[0][tid].y = k0; sh.mk[1][tid].y = k1; sh.mk[2][tid].y = k2; sh.mk[3][tid].y = k3; }
        }
        if (PREFETCH && blk + warps_total < n_blocks) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t idx = base + stride + lane + 32u * (uint32_t)j;
                if (idx < n_work) io.prefetch(idx);
            }
        }
        // phase 2, ray by ray (every thread reads only its own slots: no barrier)
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const float4 a = sh.ro[j][tid], b = sh.rd[j][tid];
            const uint32_t tag = __float_as_uint(b.w);
            if (tag == kRlNoRay) continue;
            const uint32_t meta = __float_as_uint(a.w);
            DRay r;
            r.o = mk3(a); r.d = mk3(b);
            r.face = meta & 3u; r.ex_face = (meta >> 2) & 3u; r.ex_prim = (int32_t)(meta >> 4) - 1;
            float dd;
            const bool trust = ray_trusted(sc, r, dd);
            Best best;
            best_init(best);
            confirm_tile(sc, 0u, tile_candidates(sc, 0u, sh.mk[j][tid], trust), trust, r, best, cs, sh.tile);
            cast_spheres(sc, r, trust, dd, best);
            DHit h;
            h.prim = -1; h.face = 0; h.object = 0; h.t = 0.f; h.pos = h.normal = mk3(0.f, 0.f, 0.f); h.uv.x = h.uv.y = 0.f;
            finalize_hit(sc, best, h, io.want_attrs(tag));
            cs.casts += 1ull;
            io.store(tag, h);
        }
    }
}

}  // namespace b200rt
