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
//     void     begin_block(uint32_t base)                   the warp is about to load work indices [base, base + 128)
//     uint32_t item(uint32_t idx)                           the tag of work index idx (idx < n_work), kRlNoRay for an empty slot
//     void     fetch(uint32_t tag, DRay& r)                 the ray behind a tag; the tag travels to store().  A block's four
//                                                           item() loads are issued before its fetch()es: two round trips to
//                                                           memory per block instead of eight
//     bool     want_attrs(uint32_t tag)                      false: only "is there a hit, and how far" is needed
//     uint2    culled(const DScene&, uint32_t tag, uint32_t tile)   triangles of the tile this ray is known to cull
//     void     store(uint32_t tag, const DHit& h)
#pragma once
#include "rt_cast.cuh"

namespace b200rt {

constexpr int kRlThreads = 128;                 // CTA size of the kernels built on cast_rays_in_lanes
constexpr uint32_t kRlNoRay = 0xffffffffu;      // tag of an empty ray slot (IO tags must not use it)
constexpr uint32_t kRlSphShared = 16u;          // spheres copied to shared memory (more: read from global memory)
constexpr uint32_t kRlCullClasses = 8u;         // classes of statically culled triangles an IO can name (0 = none)

// RL_PREFETCH = 1: ray sources that gather (IO::kPrefetch) keep the work items of a warp's next two blocks in a shared-memory
// ring and pull the next block's ray rows into L2 one block ahead (cast_rays_in_lanes).  Measured on B200 (round 2, cast of a
// 16-epoch 4K batch): 70.0 ms with it, 66.6 ms with the 6 KB of extra shared memory alone, 63.9 ms without either - the
// long-scoreboard stalls at the top of a block (25 % of the stall samples) are covered by the other warps already, the
// kernel is bound by issue slots, and the shared memory comes out of the L1 that serves the ray gathers: OFF.
#ifndef RL_PREFETCH
#define RL_PREFETCH 0
#endif
struct RlShared {                               // 28.4 KB per CTA (the filter records are a kernel parameter)
    float4 exact[4 * kTileTris];                // exact records {n,d}{v0,obj}{v1}{v2}  (phase 2 never waits for L1 / L2)
    float4 attr[4 * kTileTris];                 // vertex normals / uvs of the tile's triangles (winner only)
    float4 sph[kRlSphShared];
    uint2 cull[kRlCullClasses];
    float4 ro[4][kRlThreads];                   // {origin, ray meta}   of ray j of thread t
    float4 rd[4][kRlThreads];                   // {direction, tag}
    uint2 mk[4][kRlThreads];                    // candidate mask
#if RL_PREFETCH
    uint32_t ring[2][4][kRlThreads];            // IO::kPrefetch: the work items of the warp's next two blocks (cp.async destinations)
    float4 touch[kRlThreads];                   // IO::kPrefetch: where the cp.async "touches" of the next block's ray rows land (never read)
#endif
};

// cp.async (LDGSTS): global -> shared without a register in between.  The 4-byte form carries the work items of future
// blocks; the 16-byte .cg form is used as a PREFETCH INTO L2 at sector granularity: the copy pulls the 32-byte sector of
// its source into L2 (prefetch.global.L2 pulls whole 128-byte lines: measured 3x the DRAM reads) and costs no register.
RT_DI void rl_cp_async4(void* smem_dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}
RT_DI void rl_cp_async16(void* smem_dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}
RT_DI void rl_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

RT_DI uint32_t rl_pack_ray_meta(uint32_t face, int32_t ex_prim, uint32_t ex_face) {
    return face | (ex_face << 2) | ((uint32_t)(ex_prim + 1) << 4);
}

#ifndef RL_LOOP_UNROLL
#define RL_LOOP_UNROLL 2   // triangles per iteration of the filter loop (x 2 ray pairs = independent FFMA2 chains in flight)
#endif
constexpr int kRlLoopUnroll = RL_LOOP_UNROLL;

// face modes of a block of rays: every ray Front / every ray Back / anything (the cull factor is a multiply)
enum : int { kRlFront = 0, kRlBack = 1, kRlMixed = 2 };

// one filter record as named scalars (multipliers n, m_p, m_q, a, b; addends d, w_p, w_q, c), and where it comes from
struct RlRec { float nx, ny, nz, d, px, py, pz, wp, qx, qy, qz, wq, a, b, c; };
struct RlRecShared {     // tri_filter_plain layout in shared memory: four broadcast LDS.128 per record
    const float4* __restrict__ tile;
    RT_DI RlRec operator()(int i) const {
        const float4 q0 = tile[4 * i], q1 = tile[4 * i + 1], q2 = tile[4 * i + 2], q3 = tile[4 * i + 3];
        return RlRec{q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z};
    }
};
struct RlRecParam {      // RlTileParam layout in the constant bank: the multipliers travel through uniform registers
    const RlTileParam& tp;
    RT_DI RlRec operator()(int i) const {
        const float4 u0 = tp.rec[4 * i], u1 = tp.rec[4 * i + 1], u2 = tp.rec[4 * i + 2], q3 = tp.rec[4 * i + 3];
        return RlRec{u0.x, u0.y, u0.z, q3.x, u0.w, u1.x, u1.y, q3.y, u1.z, u1.w, u2.x, q3.z, u2.y, u2.z, q3.w};
    }
};

// Phase 1 for one 64-triangle tile and this lane's four rays (two packed pairs): keep[j] = the 64-bit candidate mask of
// ray j (bit i = triangle i of the tile).  Per ray x triangle pair, with the plane row scaled by s = 2^-108:
//     nd = s n.dir   num = s (d - n.o)   r = rcp.approx.ftz(nd)   t = num r   p = o + t dir
//     e_p = m_p.p + w_p   e_q = m_q.p + w_q   e_r = c - a e_p - b e_q
//     keep = sign bit of  min(e_p, e_q, e_r, t, cull) + A s |r|
// 19 FFMA2 / FMUL2 + 1 MUFU.RCP + 2 FMNMX3 + 1 SHF per pair of rays and triangle (20 with the cull multiply).
//   * near-parallel guard: |n.dir| < g = 2^-18 makes nd subnormal, the reciprocal's FTZ turns it into r = +-inf, t, p and
//     the edge functions into inf / NaN, and inf - inf (or any NaN) is the canonical NaN 0x7fffffff, sign bit clear: the
//     pair is kept, as the filter's error analysis needs - without the FADD + FMNMX the guard used to cost;
//   * cull = -r for Front rays, +r for Back rays (main.rs:185-188; FACE picks the sign at compile time, a block of mixed
//     rays multiplies by the ray's own factor, 0 for Both): a pair whose face the ray culls has cull <= -2^108 and is
//     rejected whenever |n.dir| >= g; below that the pair is kept whatever the sign of the fused dot product says.
template <int FACE, class REC>
RT_DI void rl_filter_tile(REC rec, const P2 (&ox)[2], const P2 (&oy)[2], const P2 (&oz)[2], const P2 (&dx)[2],
                          const P2 (&dy)[2], const P2 (&dz)[2], const P2 (&cf)[2], const P2 As2, uint32_t (&keep)[4][2]) {
    // reject masks (bit set = rejected), triangle i in bit (31 - i) of its half
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t rj0 = 0u, rj1 = 0u, rj2 = 0u, rj3 = 0u;
#pragma unroll kRlLoopUnroll
        for (int i = 0; i < 32; ++i) {
            const RlRec q = rec(32 * half + i);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const P2 nd = p2_fma(p2_bc(q.nz), dz[k], p2_fma(p2_bc(q.ny), dy[k], p2_mul(p2_bc(q.nx), dx[k])));
                const P2 num = p2_fma(p2_bc(-q.nz), oz[k], p2_fma(p2_bc(-q.ny), oy[k], p2_fma(p2_bc(-q.nx), ox[k], p2_bc(q.d))));
                float nda, ndb;
                p2_unpack(nd, nda, ndb);
                const float ra = rcp_approx(nda), rb = rcp_approx(ndb);
                const P2 r2 = p2_pack(ra, rb);
                const P2 t = p2_mul(num, r2);
                const P2 px = p2_fma(t, dx[k], ox[k]), py = p2_fma(t, dy[k], oy[k]), pz = p2_fma(t, dz[k], oz[k]);
                const P2 e0 = p2_fma(p2_bc(q.pz), pz, p2_fma(p2_bc(q.py), py, p2_fma(p2_bc(q.px), px, p2_bc(q.wp))));
                const P2 e1 = p2_fma(p2_bc(q.qz), pz, p2_fma(p2_bc(q.qy), py, p2_fma(p2_bc(q.qx), px, p2_bc(q.wq))));
                const P2 e2 = p2_fma(p2_bc(-q.b), e1, p2_fma(p2_bc(-q.a), e0, p2_bc(q.c)));
                float e0a, e0b, e1a, e1b, e2a, e2b, ta, tb, ca, cb;
                p2_unpack(e0, e0a, e0b); p2_unpack(e1, e1a, e1b); p2_unpack(e2, e2a, e2b); p2_unpack(t, ta, tb);
                if (FACE == kRlFront) { ca = -ra; cb = -rb; }
                else if (FACE == kRlBack) { ca = ra; cb = rb; }
                else { const P2 cull = p2_mul(r2, cf[k]); p2_unpack(cull, ca, cb); }
                const float ma = fminf(fminf(fminf(e0a, e1a), e2a), fminf(ta, ca));
                const float mb = fminf(fminf(fminf(e0b, e1b), e2b), fminf(tb, cb));
                const P2 ms = p2_fma(As2, p2_pack(fabsf(ra), fabsf(rb)), p2_pack(ma, mb));
                float ka, kb;
                p2_unpack(ms, ka, kb);
                if (k == 0) { rj0 = __funnelshift_l(__float_as_uint(ka), rj0, 1); rj1 = __funnelshift_l(__float_as_uint(kb), rj1, 1); }
                else        { rj2 = __funnelshift_l(__float_as_uint(ka), rj2, 1); rj3 = __funnelshift_l(__float_as_uint(kb), rj3, 1); }
            }
        }
        keep[0][half] = ~__brev(rj0); keep[1][half] = ~__brev(rj1); keep[2][half] = ~__brev(rj2); keep[3][half] = ~__brev(rj3);
    }
}

// The same filter over the kernel-parameter tile with PLANE RUNS: a record whose flag is set starts a run and computes
// nd, num, r, t and the plane point p for the lane's rays; the records after it (same plane up to ulps, see pack_filter)
// reuse them and only evaluate their edge functions: 9 instead of 19 FFMA2 per pair of rays and triangle, no MUFU.  The
// fixture scene's 64 triangles are 26 planes (14 squares, 12 pentagon fans).  The flag is warp-uniform (a uniform branch).
template <int FACE>
RT_DI void rl_filter_tile_runs(const RlTileParam& tp, const P2 (&ox)[2], const P2 (&oy)[2], const P2 (&oz)[2], const P2 (&dx)[2],
                               const P2 (&dy)[2], const P2 (&dz)[2], const P2 (&cf)[2], const P2 As2, uint32_t (&keep)[4][2]) {
    P2 T[2], PX[2], PY[2], PZ[2], R2[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) T[k] = PX[k] = PY[k] = PZ[k] = R2[k] = 0ull;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t rj0 = 0u, rj1 = 0u, rj2 = 0u, rj3 = 0u;
#pragma unroll kRlLoopUnroll
        for (int i = 0; i < 32; ++i) {
            const int ti = 32 * half + i;
            const float4 u2 = tp.rec[4 * ti + 2], q3 = tp.rec[4 * ti + 3];
            if (__float_as_uint(u2.w) != 0u) {                      // a new plane
                const float4 u0 = tp.rec[4 * ti];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const P2 nd = p2_fma(p2_bc(u0.z), dz[k], p2_fma(p2_bc(u0.y), dy[k], p2_mul(p2_bc(u0.x), dx[k])));
                    const P2 num = p2_fma(p2_bc(-u0.z), oz[k], p2_fma(p2_bc(-u0.y), oy[k], p2_fma(p2_bc(-u0.x), ox[k], p2_bc(q3.x))));
                    float nda, ndb;
                    p2_unpack(nd, nda, ndb);
                    R2[k] = p2_pack(rcp_approx(nda), rcp_approx(ndb));
                    T[k] = p2_mul(num, R2[k]);
                    PX[k] = p2_fma(T[k], dx[k], ox[k]); PY[k] = p2_fma(T[k], dy[k], oy[k]); PZ[k] = p2_fma(T[k], dz[k], oz[k]);
                }
            }
            const float4 u0 = tp.rec[4 * ti], u1 = tp.rec[4 * ti + 1];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const P2 e0 = p2_fma(p2_bc(u1.y), PZ[k], p2_fma(p2_bc(u1.x), PY[k], p2_fma(p2_bc(u0.w), PX[k], p2_bc(q3.y))));
                const P2 e1 = p2_fma(p2_bc(u2.x), PZ[k], p2_fma(p2_bc(u1.w), PY[k], p2_fma(p2_bc(u1.z), PX[k], p2_bc(q3.z))));
                const P2 e2 = p2_fma(p2_bc(-u2.z), e1, p2_fma(p2_bc(-u2.y), e0, p2_bc(q3.w)));
                float e0a, e0b, e1a, e1b, e2a, e2b, ta, tb, ra, rb, ca, cb;
                p2_unpack(e0, e0a, e0b); p2_unpack(e1, e1a, e1b); p2_unpack(e2, e2a, e2b); p2_unpack(T[k], ta, tb); p2_unpack(R2[k], ra, rb);
                if (FACE == kRlFront) { ca = -ra; cb = -rb; }
                else if (FACE == kRlBack) { ca = ra; cb = rb; }
                else { const P2 cull = p2_mul(R2[k], cf[k]); p2_unpack(cull, ca, cb); }
                const float ma = fminf(fminf(fminf(e0a, e1a), e2a), fminf(ta, ca));
                const float mb = fminf(fminf(fminf(e0b, e1b), e2b), fminf(tb, cb));
                const P2 ms = p2_fma(As2, p2_pack(fabsf(ra), fabsf(rb)), p2_pack(ma, mb));
                float ka, kb;
                p2_unpack(ms, ka, kb);
                if (k == 0) { rj0 = __funnelshift_l(__float_as_uint(ka), rj0, 1); rj1 = __funnelshift_l(__float_as_uint(kb), rj1, 1); }
                else        { rj2 = __funnelshift_l(__float_as_uint(ka), rj2, 1); rj3 = __funnelshift_l(__float_as_uint(kb), rj3, 1); }
            }
        }
        keep[0][half] = ~__brev(rj0); keep[1][half] = ~__brev(rj1); keep[2][half] = ~__brev(rj2); keep[3][half] = ~__brev(rj3);
    }
}
RT_DI void rl_filter_tile_runs_mode(int mode, const RlTileParam& tp, const P2 (&ox)[2], const P2 (&oy)[2], const P2 (&oz)[2], const P2 (&dx)[2],
                                    const P2 (&dy)[2], const P2 (&dz)[2], const P2 (&cf)[2], const P2 As2, uint32_t (&keep)[4][2]) {
    if (mode == kRlFront) rl_filter_tile_runs<kRlFront>(tp, ox, oy, oz, dx, dy, dz, cf, As2, keep);
    else if (mode == kRlBack) rl_filter_tile_runs<kRlBack>(tp, ox, oy, oz, dx, dy, dz, cf, As2, keep);
    else rl_filter_tile_runs<kRlMixed>(tp, ox, oy, oz, dx, dy, dz, cf, As2, keep);
}

// the loop for the face modes of a block of rays (warp-uniform `mode`)
template <class REC>
RT_DI void rl_filter_tile_mode(int mode, REC rec, const P2 (&ox)[2], const P2 (&oy)[2], const P2 (&oz)[2], const P2 (&dx)[2],
                               const P2 (&dy)[2], const P2 (&dz)[2], const P2 (&cf)[2], const P2 As2, uint32_t (&keep)[4][2]) {
    if (mode == kRlFront) rl_filter_tile<kRlFront>(rec, ox, oy, oz, dx, dy, dz, cf, As2, keep);
    else if (mode == kRlBack) rl_filter_tile<kRlBack>(rec, ox, oy, oz, dx, dy, dz, cf, As2, keep);
    else rl_filter_tile<kRlMixed>(rec, ox, oy, oz, dx, dy, dz, cf, As2, keep);
}
// cull factor of a ray for the mixed loop, and the block's mode from the warp's rays (empty slots fit any mode)
RT_DI float rl_cull_factor(uint32_t face) { return face == kFront ? -1.0f : (face == kBack ? 1.0f : 0.0f); }
RT_DI int rl_block_mode(uint32_t have_front, uint32_t have_back, uint32_t have_both) {
    const bool f = __any_sync(kFullMask, have_front != 0u), b = __any_sync(kFullMask, have_back != 0u),
               o = __any_sync(kFullMask, have_both != 0u);
    return (!b && !o) ? kRlFront : ((!f && !o) ? kRlBack : kRlMixed);
}

// CTA-collective (kRlThreads threads): casts rays [0, n_work) of `io`, 128 rays per warp and iteration, work split over
// the whole grid.  Call once per kernel; n_work may be 0.  The scene's one tile (filter, exact and attribute records), its
// spheres and the IO's classes of statically culled triangles are copied to shared memory once per CTA.
template <class IO>
RT_DI void cast_rays_in_lanes(const DScene& sc, const RlTileParam& tp, IO io, const uint32_t n_work, RlShared& sh, CastStats& cs) {
    const uint32_t lane = threadIdx.x & 31u, tid = threadIdx.x;
    if (n_work == 0u) return;
    for (uint32_t i = threadIdx.x; i < 4u * kTileTris; i += blockDim.x) {
        const bool in = i < 4u * sc.n_tris;
        sh.exact[i] = in ? sc.tri_exact[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        sh.attr[i] = in ? sc.tri_attr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x < kRlSphShared) sh.sph[threadIdx.x] = threadIdx.x < sc.n_sph ? sc.sph[threadIdx.x] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < kRlCullClasses) sh.cull[threadIdx.x] = io.cull_mask(sc, threadIdx.x);
    __syncthreads();
    const float4* __restrict__ sph = sc.n_sph <= kRlSphShared ? sh.sph : sc.sph;
    // candidates are triangles of the scene: the padding of the tile is masked off once
    const uint32_t v_lo = sc.n_tris >= 32u ? 0xffffffffu : ((1u << sc.n_tris) - 1u);
    const uint32_t v_hi = sc.n_tris >= 64u ? 0xffffffffu : (sc.n_tris > 32u ? ((1u << (sc.n_tris - 32u)) - 1u) : 0u);
#ifndef RL_PLANE_RUNS
#define RL_PLANE_RUNS 1
#endif
#if RL_PLANE_RUNS
    const P2 As2 = p2_bc(sc.filter_As_runs);
#else
    const P2 As2 = p2_bc(sc.filter_As);
#endif
    // exactly 1.0f, but opaque to ptxas: a packed multiply by it MATERIALISES each ray operand in its own aligned
    // register pair (a plain pack is coalesced with the LDG.128 destination quads and re-packed inside the loop)
    const P2 one2 = p2_bc(__fmaf_rn(sc.filter_g, 0.0f, 1.0f));
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    // (n_work + stride may exceed 2^32: the block loop counts blocks, not indices)
    const uint32_t n_blocks = (n_work + 127u) / 128u;
    // Ray sources that GATHER (a work list of path ids, rows the queues have permuted) pay two dependent DRAM round trips
    // at the top of every block (ncu: 25 % of the kernel's stall samples with 6 warps per scheduler).  With IO::kPrefetch the
    // warp keeps the work items of its next two blocks in a shared-memory ring (4-byte cp.async, issued two blocks ahead)
    // and, one block ahead, pulls the ray rows of the next block into L2 with 16-byte cp.async copies into a scratch slot:
    // the block's own loads then find items in shared memory and rows in L2.  No registers live across the filter loop.
    const uint32_t blk_first = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
#if RL_PREFETCH
    constexpr bool kPf = IO::kPrefetch;
    auto pf_items = [&](uint32_t blk, uint32_t stage) {
        if (blk >= n_blocks) return;
        const auto loc = io.locate(blk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t i = blk * 128u + lane + 32u * (uint32_t)j - loc.first;
            if (i < loc.count) rl_cp_async4(&sh.ring[stage][j][tid], loc.list + i);
        }
    };
    auto pf_rows = [&](uint32_t blk, uint32_t stage) {
        if (blk >= n_blocks) return;
        const auto loc = io.locate(blk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t i = blk * 128u + lane + 32u * (uint32_t)j - loc.first;
            if (i < loc.count) io.touch(sh.ring[stage][j][tid], loc.slot, &sh.touch[tid]);
        }
    };
    if (kPf) { pf_items(blk_first, 0u); pf_items(blk_first + warps_total, 1u); rl_cp_async_wait_all(); }
#endif
    uint32_t it = 0u;
    for (uint32_t blk = blk_first; blk < n_blocks; blk += warps_total, ++it) {
        const uint32_t base = blk * 128u;
        io.begin_block(base);
        // this lane's four rays: work indices base + lane + 32 j
        P2 ox[2], oy[2], oz[2], dx[2], dy[2], dz[2], cf[2];
        uint32_t trust4 = 0u, have_front = 0u, have_back = 0u, have_both = 0u;
        uint32_t tags[4];
        DRay rays4[4];
#ifdef RL_SEQ_LOAD
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = base + lane + 32u * (uint32_t)j;
            tags[j] = idx < n_work ? io.item(idx) : kRlNoRay;
            rays4[j].o = mk3(0.f, 0.f, 0.f); rays4[j].d = mk3(0.f, 0.f, 1.f); rays4[j].face = kFront; rays4[j].ex_prim = -1; rays4[j].ex_face = kFront;
            if (tags[j] != kRlNoRay) io.fetch(tags[j], rays4[j]);
        }
#else
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = base + lane + 32u * (uint32_t)j;
#if RL_PREFETCH
            if (kPf) tags[j] = idx < n_work ? io.item_from(idx, sh.ring[it & 1u][j][tid]) : kRlNoRay;
            else
#endif
            tags[j] = idx < n_work ? io.item(idx) : kRlNoRay;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rays4[j].o = mk3(0.f, 0.f, 0.f); rays4[j].d = mk3(0.f, 0.f, 1.f); rays4[j].face = kFront; rays4[j].ex_prim = -1; rays4[j].ex_face = kFront;
            if (tags[j] != kRlNoRay) io.fetch(tags[j], rays4[j]);
        }
#if RL_PREFETCH
        if (kPf) {
            rl_cp_async_wait_all();                       // the items of the next block (requested a block ago)
            pf_rows(blk + warps_total, (it + 1u) & 1u);   // ... whose ray rows start their way into L2 now
            pf_items(blk + 2u * warps_total, it & 1u);    // (this block's items are in registers: their ring stage is free)
        }
#endif
#endif
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            DRay (&r)[2] = reinterpret_cast<DRay (&)[2]>(rays4[2 * k]);
            float c[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * k + h;
                const uint32_t tag = tags[j];
                c[h] = rl_cull_factor(r[h].face);
                if (tag != kRlNoRay) {
                    have_front |= r[h].face == kFront; have_back |= r[h].face == kBack; have_both |= r[h].face > kBack;
                    float dd;
                    if (ray_trusted(sc, r[h], dd)) trust4 |= 1u << j;
                }
                sh.ro[j][tid] = make_float4(r[h].o.x, r[h].o.y, r[h].o.z, __uint_as_float(rl_pack_ray_meta(r[h].face, r[h].ex_prim, r[h].ex_face)));
                sh.rd[j][tid] = make_float4(r[h].d.x, r[h].d.y, r[h].d.z, __uint_as_float(tag));
            }
            ox[k] = p2_mul(p2_pack(r[0].o.x, r[1].o.x), one2); oy[k] = p2_mul(p2_pack(r[0].o.y, r[1].o.y), one2);
            oz[k] = p2_mul(p2_pack(r[0].o.z, r[1].o.z), one2);
            dx[k] = p2_mul(p2_pack(r[0].d.x, r[1].d.x), one2); dy[k] = p2_mul(p2_pack(r[0].d.y, r[1].d.y), one2);
            dz[k] = p2_mul(p2_pack(r[0].d.z, r[1].d.z), one2);
            cf[k] = p2_mul(p2_pack(c[0], c[1]), one2);
        }
        const int mode = rl_block_mode(have_front, have_back, have_both);
        // phase 1: the candidate masks of this lane's four rays
        uint32_t keep[4][2];
#if RL_PLANE_RUNS
        rl_filter_tile_runs_mode(mode, tp, ox, oy, oz, dx, dy, dz, cf, As2, keep);
#else
        rl_filter_tile_mode(mode, RlRecParam{tp}, ox, oy, oz, dx, dy, dz, cf, As2, keep);
#endif
#ifdef RL_DIAG_LOOP2   // timing experiment: the loop twice (the second result is masked by a run-time zero)
        {
            uint32_t keep2[4][2];
            const P2 A2b = p2_bc(sc.filter_As + __uint_as_float(sc.n_tris_padded >> 31));
            rl_filter_tile_mode(mode, RlRecParam{tp}, ox, oy, oz, dx, dy, dz, cf, A2b, keep2);
            const uint32_t zero_rt = sc.n_tris_padded >> 31;
#pragma unroll
            for (int j = 0; j < 4; ++j) { keep[j][0] |= keep2[j][0] & zero_rt; keep[j][1] |= keep2[j][1] & zero_rt; }
        }
#endif
#pragma unroll
        for (int j = 0; j < 4; ++j) sh.mk[j][tid] = make_uint2(keep[j][0] & v_lo, keep[j][1] & v_hi);
        // phase 2, ray by ray (every thread reads only its own slots: no barrier)
#ifdef RL_DIAG_P2X2    // timing experiment: phase 2 twice (the first pass stores nothing)
#pragma unroll 1
        for (int jj = 0; jj < 8; ++jj) {
            const int j = jj & 3;
            const bool really = jj >= 4 || (sc.n_tris_padded >> 31) != 0u;
#else
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const bool really = true;
#endif
            const float4 a = sh.ro[j][tid], b = sh.rd[j][tid];
            const uint32_t tag = __float_as_uint(b.w);
            if (tag == kRlNoRay) continue;
            const uint32_t meta = __float_as_uint(a.w);
            DRay r;
            r.o = mk3(a); r.d = mk3(b);
            r.face = meta & 3u; r.ex_face = (meta >> 2) & 3u; r.ex_prim = (int32_t)(meta >> 4) - 1;
            const bool trust = (trust4 >> j) & 1u;
            Best best;
            best_init(best);
            const uint2 km = sh.mk[j][tid], cm = sh.cull[io.cull_class(tag)];
            // untrusted rays (outside the filter's assumptions): every triangle of the scene
            const uint32_t c_lo = trust ? (km.x & ~cm.x) : v_lo, c_hi = trust ? (km.y & ~cm.y) : v_hi;
            confirm_tile(sc, 0u, ((unsigned long long)c_hi << 32) | (unsigned long long)c_lo, trust, r, best, cs, sh.exact, sh.exact);
            cast_spheres(sc, r, trust, (r.d.x * r.d.x + r.d.y * r.d.y) + r.d.z * r.d.z, best, sph);
            DHit h;
            h.prim = -1; h.face = 0; h.object = 0; h.t = 0.f; h.pos = h.normal = mk3(0.f, 0.f, 0.f); h.uv.x = h.uv.y = 0.f;
            finalize_hit(sc, best, h, io.want_attrs(tag), sh.exact, sh.attr, sph, io.all_sphere_uv());
            cs.casts += 1ull;
            if (really) io.store(tag, h);
            else if (h.t == -12345.0f) io.store(tag ^ 1u, h);    // (keeps the discarded pass alive; never true)
        }
    }
}

// ---- scenes of more than one tile ------------------------------------------------------------------------------------
// The same loop with the tiles of plain records STREAMED through shared memory: every tile is one contiguous 4 KB
// block, fetched by one TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx, issued by thread 0) into a two-deep
// ring while the CTA filters the previous tile; an mbarrier per buffer tells the CTA the bytes have landed.  The four
// warps of a CTA walk the tiles together (one __syncthreads per tile frees the buffer for the copy after next); each
// still owns its own 128 rays.  The nearest hit so far of every ray lives in shared memory between tiles (the
// position is o + t*dir, main.rs:210 / 304: recomputed, not stored); only rays with candidates in a tile touch it.
constexpr uint32_t kRlTileBytes = 4u * kTileTris * 16u;

struct RlTiledShared {                          // 38 KB per CTA
    alignas(128) float4 tile[2][4 * kTileTris]; // TMA destinations
    unsigned long long bar[2];                  // "tile landed" mbarriers
    float4 ro[4][kRlThreads];
    float4 rd[4][kRlThreads];
    uint2 mk[4][kRlThreads];
    float4 best[4][kRlThreads];                 // {(prim + 1) << 1 | backface, t, a0, a1}
    float best_a2[4][kRlThreads];
};

RT_DI uint32_t rl_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
RT_DI void rl_mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rl_smem_addr(bar)), "r"(count) : "memory");
}
RT_DI void rl_tma_load_tile(float4* dst, const float4* src, unsigned long long* bar) {
    const uint32_t b = rl_smem_addr(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(kRlTileBytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rl_smem_addr(dst)), "l"(src), "r"(kRlTileBytes), "r"(b) : "memory");
}
RT_DI void rl_mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t b = rl_smem_addr(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    } while (!ok);
}

RT_DI void rl_best_store(RlTiledShared& sh, int j, uint32_t tid, const Best& b) {
    sh.best[j][tid] = make_float4(__uint_as_float(((uint32_t)(b.prim + 1) << 1) | (b.bf & 1u)), b.t, b.a0, b.a1);
    sh.best_a2[j][tid] = b.a2;
}
RT_DI void rl_best_load(const RlTiledShared& sh, int j, uint32_t tid, const DRay& r, Best& b) {
    const float4 v = sh.best[j][tid];
    const uint32_t w = __float_as_uint(v.x);
    b.prim = (int32_t)(w >> 1) - 1; b.bf = w & 1u; b.t = v.y; b.a0 = v.z; b.a1 = v.w; b.a2 = sh.best_a2[j][tid];
    b.pos = b.prim >= 0 ? r.o + r.d * b.t : mk3(0.f, 0.f, 0.f);                   // main.rs:210 / 304
}

// One UNTRUSTED ray (origin at infinity after a t = +inf hit, |dir| != 1: the filter's bounds do not cover it) against
// the tile, by the whole warp: every lane puts two triangles through the exact test on its own (main.rs:184-227 do not
// look at the nearest-so-far), then the nearest-so-far rule (main.rs:229-233) is folded over the triangles that passed,
// in index order, by all lanes alike.  32x the speed of one lane walking 64 exact tests, same result as the walk for
// any distances (inf and NaN included).  Warp-collective: every lane calls it with the SAME ray and the owner's best.
RT_DI void rl_coop_exact_tile(const DScene& sc, uint32_t tile, const DRay& r, uint32_t lane, bool& best_valid, float& best_t,
                              Best& winner, bool& changed, CastStats& cs, bool* nan_out = nullptr) {
    const uint32_t base = tile * kTileTris;
    Best ba, bb;
    best_init(ba); best_init(bb);
    const uint32_t ia = base + lane, ib = base + 32u + lane;
    if (ia < sc.n_tris) tri_exact_test(sc.tri_exact + 4 * (size_t)ia, (int32_t)ia, r, ba);
    if (ib < sc.n_tris) tri_exact_test(sc.tri_exact + 4 * (size_t)ib, (int32_t)ib, r, bb);
    unsigned ma = __ballot_sync(kFullMask, ba.prim >= 0), mb = __ballot_sync(kFullMask, bb.prim >= 0);
    int win = -1;                                   // 0..31: triangle a of that lane, 32..63: triangle b
#pragma unroll 1
    while (ma) {
        const int l = __ffs((int)ma) - 1;
        ma &= ma - 1u;
        const float t = __shfl_sync(kFullMask, ba.t, l);
        if (best_valid && best_t < t) continue;     // main.rs:229-233
        best_valid = true; best_t = t; win = l;
        if (nan_out && t != t) *nan_out = true;     // (an accepted NaN: the walk is order-dependent from here)
    }
#pragma unroll 1
    while (mb) {
        const int l = __ffs((int)mb) - 1;
        mb &= mb - 1u;
        const float t = __shfl_sync(kFullMask, bb.t, l);
        if (best_valid && best_t < t) continue;
        best_valid = true; best_t = t; win = 32 + l;
        if (nan_out && t != t) *nan_out = true;
    }
    if (lane == 0u) cs.confirms += (unsigned long long)min(kTileTris, (int)(sc.n_tris - base));
    if (win >= 0) {
        const int l = win & 31;
        const Best mine = win < 32 ? ba : bb;
        winner.prim = __shfl_sync(kFullMask, mine.prim, l);
        winner.bf = __shfl_sync(kFullMask, mine.bf, l);
        winner.t = __shfl_sync(kFullMask, mine.t, l);
        winner.pos.x = __shfl_sync(kFullMask, mine.pos.x, l);
        winner.pos.y = __shfl_sync(kFullMask, mine.pos.y, l);
        winner.pos.z = __shfl_sync(kFullMask, mine.pos.z, l);
        winner.a0 = __shfl_sync(kFullMask, mine.a0, l);
        winner.a1 = __shfl_sync(kFullMask, mine.a1, l);
        winner.a2 = __shfl_sync(kFullMask, mine.a2, l);
        changed = true;
    }
}

// CTA-collective (kRlThreads threads, ALL of them must call it, converged): casts rays [0, n_work) of `io` against
// every tile of the scene.  Call once per kernel.
template <class IO>
RT_DI void cast_rays_in_lanes_tiled(const DScene& sc, IO io, const uint32_t n_work, RlTiledShared& sh, CastStats& cs) {
    const uint32_t lane = threadIdx.x & 31u, tid = threadIdx.x, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = sc.n_tris_padded / kTileTris;
    if (n_work == 0u || n_tiles == 0u) return;
    if (tid == 0u) { rl_mbar_init(&sh.bar[0], 1u); rl_mbar_init(&sh.bar[1], 1u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t parity0 = 0u, parity1 = 0u;
    const P2 As2 = p2_bc(sc.filter_As);
    const P2 one2 = p2_bc(__fmaf_rn(sc.filter_g, 0.0f, 1.0f));   // see cast_rays_in_lanes
    const uint32_t n_blocks = (n_work + 127u) / 128u;
    const uint32_t warps_per_cta = blockDim.x >> 5;
    const uint32_t n_iters = (n_blocks + warps_per_cta - 1u) / warps_per_cta;
    for (uint32_t it = blockIdx.x; it < n_iters; it += gridDim.x) {
        // the first two tiles of this pass (both buffers are free: the pass before ended on a barrier)
        if (tid == 0u) {
            rl_tma_load_tile(sh.tile[0], sc.tri_filter_plain, &sh.bar[0]);
            if (n_tiles > 1u) rl_tma_load_tile(sh.tile[1], sc.tri_filter_plain + 4u * kTileTris, &sh.bar[1]);
        }
        const uint32_t base = (it * warps_per_cta + warp) * 128u;   // >= n_work: this warp idles through the pass
        io.begin_block(base);
        P2 ox[2], oy[2], oz[2], dx[2], dy[2], dz[2], cf[2];
        uint32_t valid4 = 0u, trust4 = 0u, nan4 = 0u;   // nan4: rays with a NaN component (cast_nan_ray_triangles)
        uint32_t have_front = 0u, have_back = 0u, have_both = 0u;
        uint32_t tags[4];
        DRay rays4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = base + lane + 32u * (uint32_t)j;
            tags[j] = (base < n_work && idx < n_work) ? io.item(idx) : kRlNoRay;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rays4[j].o = mk3(0.f, 0.f, 0.f); rays4[j].d = mk3(0.f, 0.f, 1.f); rays4[j].face = kFront; rays4[j].ex_prim = -1; rays4[j].ex_face = kFront;
            if (tags[j] != kRlNoRay) io.fetch(tags[j], rays4[j]);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            DRay (&r)[2] = reinterpret_cast<DRay (&)[2]>(rays4[2 * k]);
            float c[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * k + h;
                const uint32_t tag = tags[j];
                c[h] = rl_cull_factor(r[h].face);
                float dd;
                if (tag != kRlNoRay) {
                    have_front |= r[h].face == kFront; have_back |= r[h].face == kBack; have_both |= r[h].face > kBack;
                    valid4 |= 1u << j;
                    if (ray_trusted(sc, r[h], dd)) trust4 |= 1u << j;
                    else if (ray_has_nan(r[h])) nan4 |= 1u << j;
                }
                sh.ro[j][tid] = make_float4(r[h].o.x, r[h].o.y, r[h].o.z, __uint_as_float(rl_pack_ray_meta(r[h].face, r[h].ex_prim, r[h].ex_face)));
                sh.rd[j][tid] = make_float4(r[h].d.x, r[h].d.y, r[h].d.z, __uint_as_float(tag));
                Best b0;
                best_init(b0);
                rl_best_store(sh, j, tid, b0);
            }
            ox[k] = p2_mul(p2_pack(r[0].o.x, r[1].o.x), one2); oy[k] = p2_mul(p2_pack(r[0].o.y, r[1].o.y), one2);
            oz[k] = p2_mul(p2_pack(r[0].o.z, r[1].o.z), one2);
            dx[k] = p2_mul(p2_pack(r[0].d.x, r[1].d.x), one2); dy[k] = p2_mul(p2_pack(r[0].d.y, r[1].d.y), one2);
            dz[k] = p2_mul(p2_pack(r[0].d.z, r[1].d.z), one2);
            cf[k] = p2_mul(p2_pack(c[0], c[1]), one2);
        }
        const bool warp_valid = __any_sync(kFullMask, valid4 != 0u);
        const int mode = rl_block_mode(have_front, have_back, have_both);
#pragma unroll 1
        for (uint32_t tile = 0; tile < n_tiles; ++tile) {
            const uint32_t buf = tile & 1u;
            if (buf == 0u) { rl_mbar_wait(&sh.bar[0], parity0); parity0 ^= 1u; }
            else           { rl_mbar_wait(&sh.bar[1], parity1); parity1 ^= 1u; }
            const float4* __restrict__ recs = sh.tile[buf];
            if (warp_valid) {                                       // warp-uniform: lanes without rays filter a dummy ray
                uint32_t keep[4][2];
                rl_filter_tile_mode(mode, RlRecShared{recs}, ox, oy, oz, dx, dy, dz, cf, As2, keep);
                // untrusted rays without NaNs: the whole warp tests the tile for each of them (rl_coop_exact_tile)
                const uint32_t coop4 = valid4 & ~trust4 & ~nan4;
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    unsigned owners = __ballot_sync(kFullMask, (coop4 >> j) & 1u);
#pragma unroll 1
                    while (owners) {
                        const int l = __ffs((int)owners) - 1;
                        owners &= owners - 1u;
                        // the owner's ray and nearest-so-far, from its shared-memory slots
                        const uint32_t ot = (tid & ~31u) + (uint32_t)l;
                        const float4 a = sh.ro[j][ot], b = sh.rd[j][ot];
                        const uint32_t meta = __float_as_uint(a.w);
                        DRay r;
                        r.o = mk3(a); r.d = mk3(b);
                        r.face = meta & 3u; r.ex_face = (meta >> 2) & 3u; r.ex_prim = (int32_t)(meta >> 4) - 1;
                        Best cur;
                        rl_best_load(sh, j, ot, r, cur);
                        bool bvalid = cur.prim >= 0, changed = false;
                        float bt = cur.t;
                        Best winner = cur;
                        rl_coop_exact_tile(sc, tile, r, lane, bvalid, bt, winner, changed, cs);
                        if (changed && lane == (uint32_t)l) rl_best_store(sh, j, tid, winner);
                        __syncwarp();
                    }
                }
                // phase 2 of this tile, for the trusted rays that have candidates in it
                uint32_t todo4 = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (((valid4 & trust4) >> j) & 1u) { if ((keep[j][0] | keep[j][1]) != 0u) todo4 |= 1u << j; }
                if (todo4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) sh.mk[j][tid] = make_uint2(keep[j][0], keep[j][1]);
#pragma unroll 1
                    while (todo4) {
                        const int j = __ffs((int)todo4) - 1;
                        todo4 &= todo4 - 1u;
                        const float4 a = sh.ro[j][tid], b = sh.rd[j][tid];
                        const uint32_t meta = __float_as_uint(a.w);
                        DRay r;
                        r.o = mk3(a); r.d = mk3(b);
                        r.face = meta & 3u; r.ex_face = (meta >> 2) & 3u; r.ex_prim = (int32_t)(meta >> 4) - 1;
                        const bool trust = (trust4 >> j) & 1u;
                        Best best;
                        rl_best_load(sh, j, tid, r, best);
                        const uint2 km = sh.mk[j][tid], cm = io.culled(sc, __float_as_uint(b.w), tile);
                        confirm_tile(sc, tile * kTileTris, tile_candidates(sc, tile, make_uint2(km.x & ~cm.x, km.y & ~cm.y), trust),
                                     trust, r, best, cs, sc.tri_exact + 4 * (size_t)(tile * kTileTris),
                                     sc.tri_exact + 4 * (size_t)(tile * kTileTris));   // (the tile's plane rows are scaled: planes from the exact records)
                        rl_best_store(sh, j, tid, best);
                    }
                }
            }
            __syncthreads();                                        // every warp is done with this buffer
            if (tid == 0u && tile + 2u < n_tiles)
                rl_tma_load_tile(sh.tile[buf], sc.tri_filter_plain + (size_t)(tile + 2u) * 4u * kTileTris, &sh.bar[buf]);
        }
        // spheres, attributes, results
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            if (!((valid4 >> j) & 1u)) continue;
            const float4 a = sh.ro[j][tid], b = sh.rd[j][tid];
            const uint32_t tag = __float_as_uint(b.w), meta = __float_as_uint(a.w);
            DRay r;
            r.o = mk3(a); r.d = mk3(b);
            r.face = meta & 3u; r.ex_face = (meta >> 2) & 3u; r.ex_prim = (int32_t)(meta >> 4) - 1;
            float dd;
            const bool trust = ray_trusted(sc, r, dd);
            Best best;
            rl_best_load(sh, j, tid, r, best);
            if ((nan4 >> j) & 1u) cast_nan_ray_triangles(sc, r, best, cs);
#ifdef RL_DIAG_UNTRUSTED   // diagnostics build: untrusted rays (every pair through the exact test) in bits 40.. of the fallback counter
            if (!trust) {
                const float oo = r.o.x * r.o.x + r.o.y * r.o.y + r.o.z * r.o.z;
                cs.fallbacks += ray_has_nan(r) ? 0ull : (!(oo <= sc.origin_bound * sc.origin_bound) ? (1ull << 40) : (1ull << 52));
            }
#endif
            cast_spheres(sc, r, trust, dd, best);
            DHit h;
            h.prim = -1; h.face = 0; h.object = 0; h.t = 0.f; h.pos = h.normal = mk3(0.f, 0.f, 0.f); h.uv.x = h.uv.y = 0.f;
            finalize_hit(sc, best, h, io.want_attrs(tag));
            cs.casts += 1ull;
            io.store(tag, h);
        }
    }
}

// ---- scenes of more than one tile, FEW rays: the tile range of a ray block split over the warps of a CTA ----------------
// A unit of work of cast_rays_in_lanes_tiled is 128 rays x ALL tiles by one warp (8 ms on the 1570 tiles of C5): a round of
// a few ten thousand rays - the later rounds of a frame, every round of an eighth of a frame on one of 8 GPUs - keeps a
// fraction of the SMs busy for that long whatever its size.  Here a CTA takes ONE block of 128 rays and its four warps walk
// a quarter of the tiles each, every warp with its own two-deep TMA ring (lane 0 issues, an mbarrier per buffer,
// __syncwarp between tiles), and the four partial results of a ray are folded in tile order with the reference's own rule,
// "skip if best.t < t" (main.rs:229-233) - which is the walk's result, because a range's winner is the last of its nearest
// triangles and a later range only wins with t <= the earlier one's.  The one exception is a NaN distance accepted inside a
// range (a ray lying in a triangle's plane, main.rs:204): it makes the walk depend on the nearest-so-far the range started
// from.  confirm_tile / rl_coop_exact_tile flag it, and such a ray takes the ordered walk over every triangle by the whole
// warp.  Spheres, attributes and the result follow as in the unsplit cast: warp w finishes ray w of every lane.
struct RlSplitShared {                              // 51 KB per CTA
    alignas(128) float4 tile[4][2][4 * kTileTris];  // per warp: TMA destinations
    unsigned long long bar[4][2];
    float4 ro[4][32];                               // the block's rays [j][lane] (every warp stages the same values)
    float4 rd[4][32];
    uint2 mk[4][4][32];                             // per warp: candidate masks of the current tile [warp][j][lane]
    float4 best[4][4][32];                          // per warp: nearest of its tile range [warp][j][lane]
    float best_a2[4][4][32];
    uint32_t bad[4][32];                            // per warp: bit j = a NaN distance was accepted for ray j in this range
};

template <class IO>
RT_DI void cast_rays_in_lanes_tiled_split(const DScene& sc, IO io, const uint32_t n_work, RlSplitShared& sh, CastStats& cs) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = sc.n_tris_padded / kTileTris;
    if (n_work == 0u || n_tiles == 0u) return;
    if (lane == 0u) { rl_mbar_init(&sh.bar[warp][0], 1u); rl_mbar_init(&sh.bar[warp][1], 1u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t parity0 = 0u, parity1 = 0u;
    const P2 As2 = p2_bc(sc.filter_As);
    const P2 one2 = p2_bc(__fmaf_rn(sc.filter_g, 0.0f, 1.0f));   // see cast_rays_in_lanes
    const uint32_t n_blocks = (n_work + 127u) / 128u;
    const uint32_t t0 = (warp * n_tiles) / 4u, t1 = ((warp + 1u) * n_tiles) / 4u;      // this warp's tiles
    auto best_store = [&](int j, const Best& b) {
        sh.best[warp][j][lane] = make_float4(__uint_as_float(((uint32_t)(b.prim + 1) << 1) | (b.bf & 1u)), b.t, b.a0, b.a1);
        sh.best_a2[warp][j][lane] = b.a2;
    };
    auto best_load = [&](uint32_t w, int j, uint32_t l, const DRay& r, Best& b) {
        const float4 v = sh.best[w][j][l];
        const uint32_t u = __float_as_uint(v.x);
        b.prim = (int32_t)(u >> 1) - 1; b.bf = u & 1u; b.t = v.y; b.a0 = v.z; b.a1 = v.w; b.a2 = sh.best_a2[w][j][l];
        b.pos = b.prim >= 0 ? r.o + r.d * b.t : mk3(0.f, 0.f, 0.f);                   // main.rs:210
    };
    auto ray_of = [&](int j, uint32_t l, DRay& r, uint32_t& tag) {
        const float4 a = sh.ro[j][l], b = sh.rd[j][l];
        const uint32_t meta = __float_as_uint(a.w);
        r.o = mk3(a); r.d = mk3(b);
        r.face = meta & 3u; r.ex_face = (meta >> 2) & 3u; r.ex_prim = (int32_t)(meta >> 4) - 1;
        tag = __float_as_uint(b.w);
    };
    for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        if (lane == 0u && t0 < t1) {
            rl_tma_load_tile(sh.tile[warp][0], sc.tri_filter_plain + (size_t)t0 * 4u * kTileTris, &sh.bar[warp][0]);
            if (t0 + 1u < t1) rl_tma_load_tile(sh.tile[warp][1], sc.tri_filter_plain + (size_t)(t0 + 1u) * 4u * kTileTris, &sh.bar[warp][1]);
        }
        const uint32_t base = blk * 128u;
        io.begin_block(base);
        P2 ox[2], oy[2], oz[2], dx[2], dy[2], dz[2], cf[2];
        uint32_t valid4 = 0u, trust4 = 0u, nan4 = 0u;
        uint32_t have_front = 0u, have_back = 0u, have_both = 0u;
        uint32_t tags[4];
        DRay rays4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = base + lane + 32u * (uint32_t)j;
            tags[j] = idx < n_work ? io.item(idx) : kRlNoRay;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rays4[j].o = mk3(0.f, 0.f, 0.f); rays4[j].d = mk3(0.f, 0.f, 1.f); rays4[j].face = kFront; rays4[j].ex_prim = -1; rays4[j].ex_face = kFront;
            if (tags[j] != kRlNoRay) io.fetch(tags[j], rays4[j]);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            DRay (&r)[2] = reinterpret_cast<DRay (&)[2]>(rays4[2 * k]);
            float c[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * k + h;
                const uint32_t tag = tags[j];
                c[h] = rl_cull_factor(r[h].face);
                float dd;
                if (tag != kRlNoRay) {
                    have_front |= r[h].face == kFront; have_back |= r[h].face == kBack; have_both |= r[h].face > kBack;
                    valid4 |= 1u << j;
                    if (ray_trusted(sc, r[h], dd)) trust4 |= 1u << j;
                    else if (ray_has_nan(r[h])) nan4 |= 1u << j;
                }
                if (warp == 0u) {       // (the four warps hold the same rays: one of them stages them for the per-ray passes)
                    sh.ro[j][lane] = make_float4(r[h].o.x, r[h].o.y, r[h].o.z, __uint_as_float(rl_pack_ray_meta(r[h].face, r[h].ex_prim, r[h].ex_face)));
                    sh.rd[j][lane] = make_float4(r[h].d.x, r[h].d.y, r[h].d.z, __uint_as_float(tag));
                }
                Best b0;
                best_init(b0);
                best_store(j, b0);
            }
            ox[k] = p2_mul(p2_pack(r[0].o.x, r[1].o.x), one2); oy[k] = p2_mul(p2_pack(r[0].o.y, r[1].o.y), one2);
            oz[k] = p2_mul(p2_pack(r[0].o.z, r[1].o.z), one2);
            dx[k] = p2_mul(p2_pack(r[0].d.x, r[1].d.x), one2); dy[k] = p2_mul(p2_pack(r[0].d.y, r[1].d.y), one2);
            dz[k] = p2_mul(p2_pack(r[0].d.z, r[1].d.z), one2);
            cf[k] = p2_mul(p2_pack(c[0], c[1]), one2);
        }
        uint32_t bad4 = 0u;
        const int mode = rl_block_mode(have_front, have_back, have_both);
        __syncthreads();                                            // the staged rays are visible to every warp
#pragma unroll 1
        for (uint32_t tile = t0; tile < t1; ++tile) {
            const uint32_t buf = (tile - t0) & 1u;
            if (buf == 0u) { rl_mbar_wait(&sh.bar[warp][0], parity0); parity0 ^= 1u; }
            else           { rl_mbar_wait(&sh.bar[warp][1], parity1); parity1 ^= 1u; }
            const float4* __restrict__ recs = sh.tile[warp][buf];
            uint32_t keep[4][2];
            rl_filter_tile_mode(mode, RlRecShared{recs}, ox, oy, oz, dx, dy, dz, cf, As2, keep);
            // untrusted rays without NaNs: the whole warp tests the tile for each of them (rl_coop_exact_tile)
            const uint32_t coop4 = valid4 & ~trust4 & ~nan4;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                unsigned owners = __ballot_sync(kFullMask, (coop4 >> j) & 1u);
#pragma unroll 1
                while (owners) {
                    const int l = __ffs((int)owners) - 1;
                    owners &= owners - 1u;
                    DRay r;
                    uint32_t tag;
                    ray_of(j, (uint32_t)l, r, tag);
                    Best cur;
                    best_load(warp, j, (uint32_t)l, r, cur);
                    bool bvalid = cur.prim >= 0, changed = false, nan_here = false;
                    float bt = cur.t;
                    Best winner = cur;
                    rl_coop_exact_tile(sc, tile, r, lane, bvalid, bt, winner, changed, cs, &nan_here);
                    if (lane == (uint32_t)l) { if (changed) best_store(j, winner); if (nan_here) bad4 |= 1u << j; }
                    __syncwarp();
                }
            }
            // phase 2 of this tile, for the trusted rays that have candidates in it
            uint32_t todo4 = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (((valid4 & trust4) >> j) & 1u) { if ((keep[j][0] | keep[j][1]) != 0u) todo4 |= 1u << j; }
            if (todo4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) sh.mk[warp][j][lane] = make_uint2(keep[j][0], keep[j][1]);
#pragma unroll 1
                while (todo4) {
                    const int j = __ffs((int)todo4) - 1;
                    todo4 &= todo4 - 1u;
                    DRay r;
                    uint32_t tag;
                    ray_of(j, lane, r, tag);
                    Best best;
                    best_load(warp, j, lane, r, best);
                    const uint2 km = sh.mk[warp][j][lane], cm = io.culled(sc, tag, tile);
                    bool nan_here = false;
                    confirm_tile(sc, tile * kTileTris, tile_candidates(sc, tile, make_uint2(km.x & ~cm.x, km.y & ~cm.y), true),
                                 true, r, best, cs, sc.tri_exact + 4 * (size_t)(tile * kTileTris),
                                 sc.tri_exact + 4 * (size_t)(tile * kTileTris), &nan_here);
                    best_store(j, best);
                    if (nan_here) bad4 |= 1u << j;
                }
            }
            __syncwarp();                                           // every lane is done with this buffer
            if (lane == 0u && tile + 2u < t1)
                rl_tma_load_tile(sh.tile[warp][buf], sc.tri_filter_plain + (size_t)(tile + 2u) * 4u * kTileTris, &sh.bar[warp][buf]);
        }
        sh.bad[warp][lane] = bad4;
        __syncthreads();                                            // the four partial results of every ray are in shared memory
        // warp w finishes ray w of every lane: fold the ranges in tile order, then spheres, attributes, result
        {
            const int j = (int)warp;
            DRay r;
            uint32_t tag;
            ray_of(j, lane, r, tag);
            const bool valid = tag != kRlNoRay;
            float dd = 1.0f;
            const bool trust = valid && ray_trusted(sc, r, dd);
            const bool has_nan = valid && !trust && ray_has_nan(r);
            Best best;
            best_init(best);
            bool bad = false;
            if (valid && !has_nan) {
#pragma unroll 1
                for (uint32_t w = 0; w < 4u; ++w) {
                    bad |= ((sh.bad[w][lane] >> j) & 1u) != 0u;
                    Best part;
                    best_load(w, j, lane, r, part);
                    if (part.prim >= 0 && !(best.prim >= 0 && best.t < part.t)) best = part;      // main.rs:229-233
                }
            }
            // an accepted NaN somewhere: the ordered walk over every triangle, one ray at a time, by the whole warp
            unsigned redo = __ballot_sync(kFullMask, bad);
#pragma unroll 1
            while (redo) {
                const int l = __ffs((int)redo) - 1;
                redo &= redo - 1u;
                DRay rr;
                uint32_t tg;
                ray_of(j, (uint32_t)l, rr, tg);
                Best w;
                best_init(w);
                bool bvalid = false, changed = false;
                float bt = 0.0f;
#pragma unroll 1
                for (uint32_t tile = 0; tile < n_tiles; ++tile) rl_coop_exact_tile(sc, tile, rr, lane, bvalid, bt, w, changed, cs);
                if ((int)lane == l) { best_init(best); if (changed) best = w; cs.fallbacks += 1ull; }
            }
            if (valid) {
                if (has_nan) cast_nan_ray_triangles(sc, r, best, cs);
                cast_spheres(sc, r, trust, dd, best);
                DHit h;
                h.prim = -1; h.face = 0; h.object = 0; h.t = 0.f; h.pos = h.normal = mk3(0.f, 0.f, 0.f); h.uv.x = h.uv.y = 0.f;
                finalize_hit(sc, best, h, io.want_attrs(tag));
                cs.casts += 1ull;
                io.store(tag, h);
            }
        }
        __syncthreads();                                            // shared memory is free for the next block
    }
}

}  // namespace b200rt
