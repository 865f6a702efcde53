// rt_filter_bench.cu — K2 micro-kernels: the ray x triangle FILTER loop in isolation, in the variants
// considered for the two-phase cast.  Each variant runs `iters` passes of one 64-triangle shared-memory
// tile for every ray and writes an xor of its candidate masks (so nothing is dead code).
//
//   variant 0  scalar FFMA,  1 ray / thread,  plain records          (4 LDS.128 per pair test)
//   variant 1  FFMA2 over TRIANGLE pairs (ray scalar-broadcast), 1 ray / thread   (4 LDS.128 per pair test)
//   variant 2  FFMA2 over RAY pairs (coefficient scalar-broadcast), 2 rays / thread (2 LDS.128 per pair test)
//   variant 3  FFMA2 over RAY pairs, 4 rays / thread                 (1 LDS.128 per pair test)
//
// Filter per pair (see rt_cast.cuh): nd = n.d, num = d - n.o, t = num * rcp(nd), p = o + t d,
// e_k = m_k.p - c_k, S = A|rcp| + B, keep = max(min(e0,e1,e2,t) + S, g - |nd|) >= 0.
#include <cuda_runtime.h>

#include "rt_types.h"

namespace b200rt {

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }

constexpr int kT = 64;  // triangles per tile

// ---------------- variant 0: scalar --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t filter_scalar_32(const float4* __restrict__ tile, int base, float ox, float oy, float oz,
                                                     float dx, float dy, float dz, float A, float B, float g) {
    uint32_t rej = 0u;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
        const float4 q0 = tile[4 * (base + i) + 0], q1 = tile[4 * (base + i) + 1], q2 = tile[4 * (base + i) + 2],
                     q3 = tile[4 * (base + i) + 3];
        const float nd = __fmaf_rn(q0.z, dz, __fmaf_rn(q0.y, dy, q0.x * dx));
        const float num = __fmaf_rn(-q0.z, oz, __fmaf_rn(-q0.y, oy, __fmaf_rn(-q0.x, ox, q0.w)));
        const float r = rcp_approx(nd);
        const float t = num * r;
        const float px = __fmaf_rn(t, dx, ox), py = __fmaf_rn(t, dy, oy), pz = __fmaf_rn(t, dz, oz);
        const float e0 = __fmaf_rn(q1.z, pz, __fmaf_rn(q1.y, py, __fmaf_rn(q1.x, px, q1.w)));
        const float e1 = __fmaf_rn(q2.z, pz, __fmaf_rn(q2.y, py, __fmaf_rn(q2.x, px, q2.w)));
        const float e2 = __fmaf_rn(q3.z, pz, __fmaf_rn(q3.y, py, __fmaf_rn(q3.x, px, q3.w)));
        const float S = __fmaf_rn(A, fabsf(r), B);
        const float m = fminf(fminf(fminf(e0, e1), e2), t) + S;
        const float gm = g - fabsf(nd);
        const float k = fmaxf(m, gm);
        rej = __funnelshift_l(__float_as_uint(k), rej, 1);
    }
    return rej;
}

__global__ void __launch_bounds__(256) filter_bench_v0(const float4* __restrict__ recs, const float4* __restrict__ rays,
                                                       uint32_t* __restrict__ out, int iters, float A, float B, float g) {
    __shared__ float4 tile[4 * kT];
    for (int i = threadIdx.x; i < 4 * kT; i += blockDim.x) tile[i] = recs[i];
    __syncthreads();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float4 o = rays[2 * tid], d = rays[2 * tid + 1];
    uint32_t acc = 0u;
    float ox = o.x;
    for (int it = 0; it < iters; ++it) {
        acc ^= filter_scalar_32(tile, 0, ox, o.y, o.z, d.x, d.y, d.z, A, B, g);
        acc ^= filter_scalar_32(tile, 32, ox, o.y, o.z, d.x, d.y, d.z, A, B, g);
        ox += 1e-3f;
    }
    out[tid] = acc;
}

// ---------------- variant 1: FFMA2 over triangle pairs --------------------------------------------------------
// pair record: 8 float4 = {nx_a,nx_b,ny_a,ny_b} {nz_a,nz_b,d_a,d_b} {m0x,m0x',m0y,m0y'} {m0z,m0z',c0,c0'} ... m1, m2
__device__ __forceinline__ uint32_t filter_tripair_32(const float4* __restrict__ tile, int pair_base, float ox, float oy,
                                                      float oz, float dx, float dy, float dz, float A, float B, float g) {
    uint32_t rej = 0u;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const float4* q = tile + 8 * (pair_base + i);
        const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5], q6 = q[6], q7 = q[7];
        const float2 nd = __ffma2_rn(f2(q1.x, q1.y), bc(dz), __ffma2_rn(f2(q0.z, q0.w), bc(dy), __fmul2_rn(f2(q0.x, q0.y), bc(dx))));
        const float2 num = __ffma2_rn(f2(q1.x, q1.y), bc(-oz), __ffma2_rn(f2(q0.z, q0.w), bc(-oy), __ffma2_rn(f2(q0.x, q0.y), bc(-ox), f2(q1.z, q1.w))));
        const float2 r = f2(rcp_approx(nd.x), rcp_approx(nd.y));
        const float2 t = __fmul2_rn(num, r);
        const float2 px = __ffma2_rn(t, bc(dx), bc(ox)), py = __ffma2_rn(t, bc(dy), bc(oy)), pz = __ffma2_rn(t, bc(dz), bc(oz));
        const float2 e0 = __ffma2_rn(f2(q3.x, q3.y), pz, __ffma2_rn(f2(q2.z, q2.w), py, __ffma2_rn(f2(q2.x, q2.y), px, f2(q3.z, q3.w))));
        const float2 e1 = __ffma2_rn(f2(q5.x, q5.y), pz, __ffma2_rn(f2(q4.z, q4.w), py, __ffma2_rn(f2(q4.x, q4.y), px, f2(q5.z, q5.w))));
        const float2 e2 = __ffma2_rn(f2(q7.x, q7.y), pz, __ffma2_rn(f2(q6.z, q6.w), py, __ffma2_rn(f2(q6.x, q6.y), px, f2(q7.z, q7.w))));
        const float2 S = __ffma2_rn(bc(A), f2(fabsf(r.x), fabsf(r.y)), bc(B));
        const float ma = fminf(fminf(fminf(e0.x, e1.x), e2.x), t.x), mb = fminf(fminf(fminf(e0.y, e1.y), e2.y), t.y);
        const float2 m = __fadd2_rn(f2(ma, mb), S);
        const float2 gm = __fadd2_rn(bc(g), f2(-fabsf(nd.x), -fabsf(nd.y)));
        rej = __funnelshift_l(__float_as_uint(fmaxf(m.x, gm.x)), rej, 1);
        rej = __funnelshift_l(__float_as_uint(fmaxf(m.y, gm.y)), rej, 1);
    }
    return rej;
}

__global__ void __launch_bounds__(256) filter_bench_v1(const float4* __restrict__ recs, const float4* __restrict__ rays,
                                                       uint32_t* __restrict__ out, int iters, float A, float B, float g) {
    __shared__ float4 tile[4 * kT];
    for (int i = threadIdx.x; i < 4 * kT; i += blockDim.x) tile[i] = recs[i];
    __syncthreads();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float4 o = rays[2 * tid], d = rays[2 * tid + 1];
    uint32_t acc = 0u;
    float ox = o.x;
    for (int it = 0; it < iters; ++it) {
        acc ^= filter_tripair_32(tile, 0, ox, o.y, o.z, d.x, d.y, d.z, A, B, g);
        acc ^= filter_tripair_32(tile, 16, ox, o.y, o.z, d.x, d.y, d.z, A, B, g);
        ox += 1e-3f;
    }
    out[tid] = acc;
}

// ---------------- variants 2/3: FFMA2 over ray pairs, NP ray pairs per thread -----------------------------------
template <int NP>
struct RayPairs {
    float2 ox[NP], oy[NP], oz[NP], dx[NP], dy[NP], dz[NP];
};

template <int NP>
__device__ __forceinline__ void filter_raypair_32(const float4* __restrict__ tile, int base, const RayPairs<NP>& R, float A,
                                                  float B, float g, uint32_t (&rej)[2 * NP]) {
#pragma unroll
    for (int k = 0; k < 2 * NP; ++k) rej[k] = 0u;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
        const float4 q0 = tile[4 * (base + i) + 0], q1 = tile[4 * (base + i) + 1], q2 = tile[4 * (base + i) + 2],
                     q3 = tile[4 * (base + i) + 3];
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const float2 nd = __ffma2_rn(bc(q0.z), R.dz[k], __ffma2_rn(bc(q0.y), R.dy[k], __fmul2_rn(bc(q0.x), R.dx[k])));
            const float2 no = __ffma2_rn(bc(q0.z), R.oz[k], __ffma2_rn(bc(q0.y), R.oy[k], __fmul2_rn(bc(q0.x), R.ox[k])));
            const float2 num = __fadd2_rn(bc(q0.w), f2(-no.x, -no.y));
            const float2 r = f2(rcp_approx(nd.x), rcp_approx(nd.y));
            const float2 t = __fmul2_rn(num, r);
            const float2 px = __ffma2_rn(t, R.dx[k], R.ox[k]), py = __ffma2_rn(t, R.dy[k], R.oy[k]), pz = __ffma2_rn(t, R.dz[k], R.oz[k]);
            const float2 e0 = __ffma2_rn(bc(q1.z), pz, __ffma2_rn(bc(q1.y), py, __ffma2_rn(bc(q1.x), px, bc(q1.w))));
            const float2 e1 = __ffma2_rn(bc(q2.z), pz, __ffma2_rn(bc(q2.y), py, __ffma2_rn(bc(q2.x), px, bc(q2.w))));
            const float2 e2 = __ffma2_rn(bc(q3.z), pz, __ffma2_rn(bc(q3.y), py, __ffma2_rn(bc(q3.x), px, bc(q3.w))));
            const float2 S = __ffma2_rn(bc(A), f2(fabsf(r.x), fabsf(r.y)), bc(B));
            const float ma = fminf(fminf(fminf(e0.x, e1.x), e2.x), t.x), mb = fminf(fminf(fminf(e0.y, e1.y), e2.y), t.y);
            const float2 m = __fadd2_rn(f2(ma, mb), S);
            const float2 gm = __fadd2_rn(bc(g), f2(-fabsf(nd.x), -fabsf(nd.y)));
            rej[2 * k] = __funnelshift_l(__float_as_uint(fmaxf(m.x, gm.x)), rej[2 * k], 1);
            rej[2 * k + 1] = __funnelshift_l(__float_as_uint(fmaxf(m.y, gm.y)), rej[2 * k + 1], 1);
        }
    }
}

template <int NP>
__global__ void __launch_bounds__(256) filter_bench_rp(const float4* __restrict__ recs, const float4* __restrict__ rays,
                                                       uint32_t* __restrict__ out, int iters, float A, float B, float g) {
    __shared__ float4 tile[4 * kT];
    for (int i = threadIdx.x; i < 4 * kT; i += blockDim.x) tile[i] = recs[i];
    __syncthreads();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    RayPairs<NP> R;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const float4 oa = rays[2 * (tid * 2 * NP + 2 * k)], da = rays[2 * (tid * 2 * NP + 2 * k) + 1];
        const float4 ob = rays[2 * (tid * 2 * NP + 2 * k + 1)], db = rays[2 * (tid * 2 * NP + 2 * k + 1) + 1];
        R.ox[k] = f2(oa.x, ob.x); R.oy[k] = f2(oa.y, ob.y); R.oz[k] = f2(oa.z, ob.z);
        R.dx[k] = f2(da.x, db.x); R.dy[k] = f2(da.y, db.y); R.dz[k] = f2(da.z, db.z);
    }
    uint32_t acc = 0u;
    for (int it = 0; it < iters; ++it) {
        uint32_t rej[2 * NP];
        filter_raypair_32<NP>(tile, 0, R, A, B, g, rej);
#pragma unroll
        for (int k = 0; k < 2 * NP; ++k) acc ^= rej[k] * (k + 1);
        filter_raypair_32<NP>(tile, 32, R, A, B, g, rej);
#pragma unroll
        for (int k = 0; k < 2 * NP; ++k) acc ^= rej[k] * (k + 3);
#pragma unroll
        for (int k = 0; k < NP; ++k) R.ox[k].x += 1e-3f;
    }
    out[tid] = acc;
}

// ---------------- pipe calibration: isolated instruction-throughput loops ---------------------------------------
// variant 0: FFMA  a = a*b + c                 (b, c loop-invariant)
// variant 1: FFMA2 a2 = a2*b2 + c2
// variant 2: FFMA2 a2 = b.F32 * a2 + c2        (scalar-broadcast operand)
// variant 3: FFMA  x_i = x_j * x_k + x_i       (three live, changing registers)
// variant 4: FMNMX chains                      (ALU pipe)
// variant 5: MUFU.RCP chains                   (XU pipe)
// variant 6: 3 FFMA2 : 1 FMNMX mix
// variant 7: FFMA2 x2_i = x2_j * x2_k + x2_i   (three live packed registers)
template <int V>
__global__ void __launch_bounds__(256) pipe_bench_kernel(float* sink, int iters) {
    float2 a[8];
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i); s[i] = a[i].x + 0.5f; }
    const float b = 0.999f + 1e-6f * threadIdx.x, c = 1e-3f;
    const float2 b2 = make_float2(b, b * 0.9999f), c2 = make_float2(c, c * 2.0f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (V == 0) s[i] = __fmaf_rn(s[i], b, c);
                if (V == 1) a[i] = __ffma2_rn(a[i], b2, c2);
                if (V == 2) a[i] = __ffma2_rn(make_float2(b, b), a[i], c2);
                if (V == 3) s[i] = __fmaf_rn(s[(i + 1) & 7], s[(i + 3) & 7], s[i]);
                if (V == 4) s[i] = fminf(s[i], s[(i + 1) & 7] + 0.0f);
                if (V == 5) s[i] = rcp_approx(s[i]);
                if (V == 6) { a[i] = __ffma2_rn(a[i], b2, c2); if ((i & 3) == 3) s[i] = fmaxf(s[i], a[i].x); }
                if (V == 7) a[i] = __ffma2_rn(a[(i + 1) & 7], a[(i + 3) & 7], a[i]);
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y + s[i];
    if (r == 123456.789f) sink[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

cudaError_t launch_pipe_bench(int variant, float* d_sink, int blocks, int iters, cudaStream_t stream) {
    switch (variant) {
        case 0: pipe_bench_kernel<0><<<blocks, 256, 0, stream>>>(d_sink, iters); break;
        case 1: pipe_bench_kernel<1><<<blocks, 256, 0, stream>>>(d_sink, iters); break;
        case 2: pipe_bench_kernel<2><<<blocks, 256, 0, stream>>>(d_sink, iters); break;
        case 3: pipe_bench_kernel<3><<<blocks, 256, 0, stream>>>(d_sink, iters); break;
        case 4: pipe_bench_kernel<4><<<blocks, 256, 0, stream>>>(d_sink, iters); break;
        case 5: pipe_bench_kernel<5><<<blocks, 256, 0, stream>>>(d_sink, iters); break;
        case 6: pipe_bench_kernel<6><<<blocks, 256, 0, stream>>>(d_sink, iters); break;
        case 7: pipe_bench_kernel<7><<<blocks, 256, 0, stream>>>(d_sink, iters); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// returns the number of ray x triangle pair tests performed
// ---------------- variants 4 / 5: rays in lanes, triangle records in CONSTANT memory -----------------------------
// Every lane owns RAYS rays; the 64 plain records {n,d}{m0,w0}{m1,w1}{m2,w2} are warp-uniform and come from the
// constant bank / uniform registers, so an FFMA reads two registers (ray component, accumulator), not three.
__constant__ float4 c_recs[4 * kT];

template <int RAYS, int UNROLL>
__global__ void __launch_bounds__(256) filter_bench_const(const float4* __restrict__ rays, uint32_t* __restrict__ out, int iters,
                                                          float A, float B, float g) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float ox[RAYS], oy[RAYS], oz[RAYS], dx[RAYS], dy[RAYS], dz[RAYS];
#pragma unroll
    for (int r = 0; r < RAYS; ++r) {
        const float4 o = rays[2 * (tid * RAYS + r)], d = rays[2 * (tid * RAYS + r) + 1];
        ox[r] = o.x; oy[r] = o.y; oz[r] = o.z; dx[r] = d.x; dy[r] = d.y; dz[r] = d.z;
    }
    uint32_t acc = 0u;
    for (int it = 0; it < iters; ++it) {
        unsigned long long mask[RAYS];
#pragma unroll
        for (int r = 0; r < RAYS; ++r) mask[r] = 0ull;
#pragma unroll UNROLL
        for (int i = 0; i < kT; ++i) {
            const float4 q0 = c_recs[4 * i + 0], q1 = c_recs[4 * i + 1], q2 = c_recs[4 * i + 2], q3 = c_recs[4 * i + 3];
#pragma unroll
            for (int r = 0; r < RAYS; ++r) {
                const float nd = __fmaf_rn(q0.z, dz[r], __fmaf_rn(q0.y, dy[r], q0.x * dx[r]));
                const float num = __fmaf_rn(-q0.z, oz[r], __fmaf_rn(-q0.y, oy[r], __fmaf_rn(-q0.x, ox[r], q0.w)));
                const float rc = rcp_approx(nd);
                const float t = num * rc;
                const float px = __fmaf_rn(t, dx[r], ox[r]), py = __fmaf_rn(t, dy[r], oy[r]), pz = __fmaf_rn(t, dz[r], oz[r]);
                const float e0 = __fmaf_rn(q1.z, pz, __fmaf_rn(q1.y, py, __fmaf_rn(q1.x, px, q1.w)));
                const float e1 = __fmaf_rn(q2.z, pz, __fmaf_rn(q2.y, py, __fmaf_rn(q2.x, px, q2.w)));
                const float e2 = __fmaf_rn(q3.z, pz, __fmaf_rn(q3.y, py, __fmaf_rn(q3.x, px, q3.w)));
                const float m = __fmaf_rn(A, fabsf(rc), fminf(fminf(fminf(e0, e1), e2), t));
                const bool keep = (m >= 0.0f) | (fabsf(nd) < g);
                if (keep) mask[r] |= 1ull << i;
            }
        }
#pragma unroll
        for (int r = 0; r < RAYS; ++r) { acc ^= (uint32_t)mask[r] ^ (uint32_t)(mask[r] >> 32); ox[r] += 1e-3f; }
    }
    out[tid] = acc;
}

cudaError_t launch_filter_bench(int variant, const float4* d_recs, const float4* d_rays, uint32_t* d_out, int blocks,
                                int iters, float A, float B, float g, cudaStream_t stream, unsigned long long* pairs) {
    const unsigned long long threads = (unsigned long long)blocks * 256ull;
    switch (variant) {
        case 0: filter_bench_v0<<<blocks, 256, 0, stream>>>(d_recs, d_rays, d_out, iters, A, B, g); *pairs = threads * kT * iters; break;
        case 1: filter_bench_v1<<<blocks, 256, 0, stream>>>(d_recs, d_rays, d_out, iters, A, B, g); *pairs = threads * kT * iters; break;
        case 2: filter_bench_rp<1><<<blocks, 256, 0, stream>>>(d_recs, d_rays, d_out, iters, A, B, g); *pairs = threads * 2 * kT * iters; break;
        case 3: filter_bench_rp<2><<<blocks, 256, 0, stream>>>(d_recs, d_rays, d_out, iters, A, B, g); *pairs = threads * 4 * kT * iters; break;
        case 4: case 5: case 6: case 7: {
            cudaError_t e = cudaMemcpyToSymbolAsync(c_recs, d_recs, sizeof(float4) * 4 * kT, 0, cudaMemcpyDeviceToDevice, stream);
            if (e != cudaSuccess) return e;
            if (variant == 4) { filter_bench_const<1, 4><<<blocks, 256, 0, stream>>>(d_rays, d_out, iters, A, B, g); *pairs = threads * kT * iters; }
            if (variant == 5) { filter_bench_const<2, 4><<<blocks, 256, 0, stream>>>(d_rays, d_out, iters, A, B, g); *pairs = threads * 2 * kT * iters; }
            if (variant == 6) { filter_bench_const<1, 64><<<blocks, 256, 0, stream>>>(d_rays, d_out, iters, A, B, g); *pairs = threads * kT * iters; }
            if (variant == 7) { filter_bench_const<2, 64><<<blocks, 256, 0, stream>>>(d_rays, d_out, iters, A, B, g); *pairs = threads * 2 * kT * iters; }
            break;
        }
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace b200rt
