// loop_bench.cu — dev micro-benchmark of the rays-in-lanes filter loop (rt_cast_rl.cuh) in isolation: where do the
// triangle records come from?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o loop_bench loop_bench.cu
//   smem   : 4 broadcast LDS.128 per record, every FFMA2 reads its scalar from a vector register (production form, round 1)
//   param  : records in the kernel-parameter constant bank, multipliers through UNIFORM registers (LDCU.64 -> FFMA2 R, R, UR, R)
// Each thread owns 4 rays (2 packed pairs); `iters` passes over one 64-triangle tile.  Prints Gpairs/s and the fraction
// of the FP32 roofline at 36 flop per pair.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned long long P2;
#define DI __device__ __forceinline__
DI P2 p2_pack(float lo, float hi) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
DI P2 p2_bc(float a) { P2 r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(a)); return r; }
DI void p2_unpack(P2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DI P2 p2_fma(P2 a, P2 b, P2 c) { P2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
DI P2 p2_mul(P2 a, P2 b) { P2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
DI float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

struct Tile { float4 rec[256]; };
// one record as named scalars: multipliers n, m0, m1, m2 (or a, b with E2DEP) and addends d, w0, w1, w2 (c)
struct Rec { float nx, ny, nz, d, m0x, m0y, m0z, w0, m1x, m1y, m1z, w1, m2x, m2y, m2z, w2; };

// FACE: 0 front (cull = -nd), 1 back, 2 mixed (multiply);  E2DEP: third edge from the other two
template <int FACE, bool E2DEP, bool FTZG, class Fetch>
DI void filter_tile(Fetch fetch, const P2 (&ox)[2], const P2 (&oy)[2], const P2 (&oz)[2], const P2 (&dx)[2], const P2 (&dy)[2],
                    const P2 (&dz)[2], const P2 (&cf)[2], const P2 A2, const float g, uint32_t (&keep)[4][2]) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t rj0 = 0u, rj1 = 0u, rj2 = 0u, rj3 = 0u;
#pragma unroll 2
        for (int i = 0; i < 32; ++i) {
            const Rec q = fetch(32 * half + i);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const P2 nd = p2_fma(p2_bc(q.nz), dz[k], p2_fma(p2_bc(q.ny), dy[k], p2_mul(p2_bc(q.nx), dx[k])));
                const P2 num = p2_fma(p2_bc(-q.nz), oz[k], p2_fma(p2_bc(-q.ny), oy[k], p2_fma(p2_bc(-q.nx), ox[k], p2_bc(q.d))));
                float nda, ndb; p2_unpack(nd, nda, ndb);
                const float ra = rcp_approx(nda), rb = rcp_approx(ndb);
                const P2 t = p2_mul(num, p2_pack(ra, rb));
                const P2 px = p2_fma(t, dx[k], ox[k]), py = p2_fma(t, dy[k], oy[k]), pz = p2_fma(t, dz[k], oz[k]);
                const P2 e0 = p2_fma(p2_bc(q.m0z), pz, p2_fma(p2_bc(q.m0y), py, p2_fma(p2_bc(q.m0x), px, p2_bc(q.w0))));
                const P2 e1 = p2_fma(p2_bc(q.m1z), pz, p2_fma(p2_bc(q.m1y), py, p2_fma(p2_bc(q.m1x), px, p2_bc(q.w1))));
                P2 e2;
                if (E2DEP) e2 = p2_fma(p2_bc(-q.m2y), e1, p2_fma(p2_bc(-q.m2x), e0, p2_bc(q.w2)));
                else e2 = p2_fma(p2_bc(q.m2z), pz, p2_fma(p2_bc(q.m2y), py, p2_fma(p2_bc(q.m2x), px, p2_bc(q.w2))));
                float e0a, e0b, e1a, e1b, e2a, e2b, ta, tb, ca, cb;
                p2_unpack(e0, e0a, e0b); p2_unpack(e1, e1a, e1b); p2_unpack(e2, e2a, e2b); p2_unpack(t, ta, tb);
                if (FTZG) {      // cull term from the reciprocal: -r for front rays, +r for back rays (sign of n.dir, huge)
                    if (FACE == 0) { ca = -ra; cb = -rb; }
                    else if (FACE == 1) { ca = ra; cb = rb; }
                    else { const P2 cull = p2_mul(p2_pack(ra, rb), cf[k]); p2_unpack(cull, ca, cb); }
                } else {
                    if (FACE == 0) { ca = -nda; cb = -ndb; }
                    else if (FACE == 1) { ca = nda; cb = ndb; }
                    else { const P2 cull = p2_mul(nd, cf[k]); p2_unpack(cull, ca, cb); }
                }
                const float ma = fminf(fminf(fminf(e0a, e1a), e2a), fminf(ta, ca));
                const float mb = fminf(fminf(fminf(e0b, e1b), e2b), fminf(tb, cb));
                const P2 ms = p2_fma(A2, p2_pack(fabsf(ra), fabsf(rb)), p2_pack(ma, mb));
                float msa, msb; p2_unpack(ms, msa, msb);
                float ka, kb;
                if (FTZG) { ka = msa; kb = msb; }
                else { ka = fmaxf(msa, g - fabsf(nda)); kb = fmaxf(msb, g - fabsf(ndb)); }
                if (k == 0) { rj0 = __funnelshift_l(__float_as_uint(ka), rj0, 1); rj1 = __funnelshift_l(__float_as_uint(kb), rj1, 1); }
                else        { rj2 = __funnelshift_l(__float_as_uint(ka), rj2, 1); rj3 = __funnelshift_l(__float_as_uint(kb), rj3, 1); }
            }
        }
        keep[0][half] = ~__brev(rj0); keep[1][half] = ~__brev(rj1); keep[2][half] = ~__brev(rj2); keep[3][half] = ~__brev(rj3);
    }
}

struct FetchSmem {   // plain layout {n,d}{m0,w0}{m1,w1}{m2,w2}
    const float4* tile;
    DI Rec operator()(int i) const {
        const float4 a = tile[4 * i], b = tile[4 * i + 1], c = tile[4 * i + 2], d = tile[4 * i + 3];
        return Rec{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
    }
};
struct FetchParam {  // split layout: {n.xyz, m0.x}{m0.yz, m1.xy}{m1.z, m2.xyz} multipliers, {d, w0, w1, w2} addends
    const Tile& tile;
    DI Rec operator()(int i) const {
        const float4 a = tile.rec[4 * i], b = tile.rec[4 * i + 1], c = tile.rec[4 * i + 2], d = tile.rec[4 * i + 3];
        return Rec{a.x, a.y, a.z, d.x, a.w, b.x, b.y, d.y, b.z, b.w, c.x, d.z, c.y, c.z, c.w, d.w};
    }
};

template <class F>
DI void load_rays(const float4* __restrict__ rays, size_t tid, P2 (&ox)[2], P2 (&oy)[2], P2 (&oz)[2], P2 (&dx)[2], P2 (&dy)[2],
                  P2 (&dz)[2], P2 (&cf)[2], float one, F) {
    const P2 one2 = p2_bc(one);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float4 a = rays[2 * (4 * tid + 2 * k)], b = rays[2 * (4 * tid + 2 * k) + 1];
        const float4 c = rays[2 * (4 * tid + 2 * k + 1)], d = rays[2 * (4 * tid + 2 * k + 1) + 1];
        ox[k] = p2_mul(p2_pack(a.x, c.x), one2); oy[k] = p2_mul(p2_pack(a.y, c.y), one2); oz[k] = p2_mul(p2_pack(a.z, c.z), one2);
        dx[k] = p2_mul(p2_pack(b.x, d.x), one2); dy[k] = p2_mul(p2_pack(b.y, d.y), one2); dz[k] = p2_mul(p2_pack(b.z, d.z), one2);
        cf[k] = p2_mul(p2_pack(-1.0f, -1.0f), one2);
    }
}

template <int FACE, bool E2DEP, bool FTZG, int MINB>
__global__ void __launch_bounds__(128, MINB) k_smem(const float4* __restrict__ recs, const float4* __restrict__ rays, uint32_t* out,
                                                    int iters, float A, float g, float one) {
    __shared__ float4 tile[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tile[i] = recs[i];
    __syncthreads();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    P2 ox[2], oy[2], oz[2], dx[2], dy[2], dz[2], cf[2];
    load_rays(rays, tid, ox, oy, oz, dx, dy, dz, cf, one, 0);
    uint32_t acc = 0u;
    for (int it = 0; it < iters; ++it) {
        uint32_t keep[4][2];
        filter_tile<FACE, E2DEP, FTZG>(FetchSmem{tile}, ox, oy, oz, dx, dy, dz, cf, p2_bc(A), g, keep);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc ^= keep[j][0] * (j + 1) ^ keep[j][1] * (j + 5);
        ox[0] = p2_fma(ox[0], p2_bc(one), p2_bc(1e-3f));
    }
    out[tid] = acc;
}
template <int FACE, bool E2DEP, bool FTZG, int MINB>
__global__ void __launch_bounds__(128, MINB) k_param(const __grid_constant__ Tile tile, const float4* __restrict__ rays, uint32_t* out,
                                                     int iters, float A, float g, float one) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    P2 ox[2], oy[2], oz[2], dx[2], dy[2], dz[2], cf[2];
    load_rays(rays, tid, ox, oy, oz, dx, dy, dz, cf, one, 0);
    uint32_t acc = 0u;
    for (int it = 0; it < iters; ++it) {
        uint32_t keep[4][2];
        filter_tile<FACE, E2DEP, FTZG>(FetchParam{tile}, ox, oy, oz, dx, dy, dz, cf, p2_bc(A), g, keep);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc ^= keep[j][0] * (j + 1) ^ keep[j][1] * (j + 5);
        ox[0] = p2_fma(ox[0], p2_bc(one), p2_bc(1e-3f));
    }
    out[tid] = acc;
}

// ---- plane runs: consecutive coplanar triangles share nd, num, r, t, p (flag in u2.w: this record starts a new plane) ----
template <int FACE, int UNROLL>
DI void filter_tile_runs(const Tile& tile, const P2 (&ox)[2], const P2 (&oy)[2], const P2 (&oz)[2], const P2 (&dx)[2], const P2 (&dy)[2],
                         const P2 (&dz)[2], const P2 (&cf)[2], const P2 A2, uint32_t (&keep)[4][2]) {
    P2 T[2], PX[2], PY[2], PZ[2], R2[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) T[k] = PX[k] = PY[k] = PZ[k] = R2[k] = 0ull;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t rj0 = 0u, rj1 = 0u, rj2 = 0u, rj3 = 0u;
#pragma unroll UNROLL
        for (int i = 0; i < 32; ++i) {
            const int ti = 32 * half + i;
            const float4 u2 = tile.rec[4 * ti + 2], q3 = tile.rec[4 * ti + 3];
            if (__float_as_uint(u2.w) != 0u) {                      // uniform: a new plane
                const float4 u0 = tile.rec[4 * ti];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const P2 nd = p2_fma(p2_bc(u0.z), dz[k], p2_fma(p2_bc(u0.y), dy[k], p2_mul(p2_bc(u0.x), dx[k])));
                    const P2 num = p2_fma(p2_bc(-u0.z), oz[k], p2_fma(p2_bc(-u0.y), oy[k], p2_fma(p2_bc(-u0.x), ox[k], p2_bc(q3.x))));
                    float nda, ndb; p2_unpack(nd, nda, ndb);
                    R2[k] = p2_pack(rcp_approx(nda), rcp_approx(ndb));
                    T[k] = p2_mul(num, R2[k]);
                    PX[k] = p2_fma(T[k], dx[k], ox[k]); PY[k] = p2_fma(T[k], dy[k], oy[k]); PZ[k] = p2_fma(T[k], dz[k], oz[k]);
                }
            }
            const float4 u0 = tile.rec[4 * ti], u1 = tile.rec[4 * ti + 1];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const P2 e0 = p2_fma(p2_bc(u1.y), PZ[k], p2_fma(p2_bc(u1.x), PY[k], p2_fma(p2_bc(u0.w), PX[k], p2_bc(q3.y))));
                const P2 e1 = p2_fma(p2_bc(u2.x), PZ[k], p2_fma(p2_bc(u1.w), PY[k], p2_fma(p2_bc(u1.z), PX[k], p2_bc(q3.z))));
                const P2 e2 = p2_fma(p2_bc(-u2.z), e1, p2_fma(p2_bc(-u2.y), e0, p2_bc(q3.w)));
                float e0a, e0b, e1a, e1b, e2a, e2b, ta, tb, ra, rb, ca, cb;
                p2_unpack(e0, e0a, e0b); p2_unpack(e1, e1a, e1b); p2_unpack(e2, e2a, e2b); p2_unpack(T[k], ta, tb); p2_unpack(R2[k], ra, rb);
                if (FACE == 0) { ca = -ra; cb = -rb; }
                else if (FACE == 1) { ca = ra; cb = rb; }
                else { const P2 cull = p2_mul(R2[k], cf[k]); p2_unpack(cull, ca, cb); }
                const float ma = fminf(fminf(fminf(e0a, e1a), e2a), fminf(ta, ca));
                const float mb = fminf(fminf(fminf(e0b, e1b), e2b), fminf(tb, cb));
                const P2 ms = p2_fma(A2, p2_pack(fabsf(ra), fabsf(rb)), p2_pack(ma, mb));
                float ka, kb; p2_unpack(ms, ka, kb);
                if (k == 0) { rj0 = __funnelshift_l(__float_as_uint(ka), rj0, 1); rj1 = __funnelshift_l(__float_as_uint(kb), rj1, 1); }
                else        { rj2 = __funnelshift_l(__float_as_uint(ka), rj2, 1); rj3 = __funnelshift_l(__float_as_uint(kb), rj3, 1); }
            }
        }
        keep[0][half] = ~__brev(rj0); keep[1][half] = ~__brev(rj1); keep[2][half] = ~__brev(rj2); keep[3][half] = ~__brev(rj3);
    }
}
template <int FACE, int UNROLL, int MINB>
__global__ void __launch_bounds__(128, MINB) k_runs(const __grid_constant__ Tile tile, const float4* __restrict__ rays, uint32_t* out,
                                                    int iters, float A, float g, float one) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    P2 ox[2], oy[2], oz[2], dx[2], dy[2], dz[2], cf[2];
    load_rays(rays, tid, ox, oy, oz, dx, dy, dz, cf, one, 0);
    uint32_t acc = 0u;
    for (int it = 0; it < iters; ++it) {
        uint32_t keep[4][2];
        filter_tile_runs<FACE, UNROLL>(tile, ox, oy, oz, dx, dy, dz, cf, p2_bc(A), keep);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc ^= keep[j][0] * (j + 1) ^ keep[j][1] * (j + 5);
        ox[0] = p2_fma(ox[0], p2_bc(one), p2_bc(1e-3f));
    }
    out[tid] = acc;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 256;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    // synthetic tile: 8x8 patch of small triangles in z = 0 (plain layout), rays above it pointing down
    std::vector<float4> plain(256), split(256);
    for (int i = 0; i < 64; ++i) {
        const float cx = (float)(i % 8) * 0.25f, cy = (float)(i / 8) * 0.25f;
        const float r[16] = {0.f, 0.f, 1.f, 0.f, 1.f, 0.f, 0.f, -cx, 0.f, 1.f, 0.f, -cy, -0.70710678f, -0.70710678f, 0.f, 0.70710678f * (cx + cy + 0.25f)};
        for (int k = 0; k < 4; ++k) plain[4 * i + k] = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        split[4 * i + 0] = make_float4(r[0], r[1], r[2], r[4]);
        split[4 * i + 1] = make_float4(r[5], r[6], r[8], r[9]);
        split[4 * i + 2] = make_float4(r[10], r[12], r[13], r[14]);
        split[4 * i + 3] = make_float4(r[3], r[7], r[11], r[15]);
    }
    Tile tp; for (int i = 0; i < 256; ++i) tp.rec[i] = split[i];
    Tile tp_runs = tp, tp_all = tp;     // u2.w: new-plane flag
    for (int i = 0; i < 64; ++i) {
        const bool start = i < 36 ? (i % 3 == 0) : ((i - 36) % 2 == 0);
        tp_runs.rec[4 * i + 2].w = start ? 1.0f : 0.0f;
        tp_all.rec[4 * i + 2].w = 1.0f;
    }
    const int max_blocks = sms * 8;
    const size_t n_rays = (size_t)max_blocks * 128 * 4;
    std::vector<float4> rays(2 * n_rays);
    uint32_t s = 12345u;
    auto rnd = [&s]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); };
    for (size_t i = 0; i < n_rays; ++i) {
        float dx = rnd() - 0.5f, dy = rnd() - 0.5f, dz = -1.0f;
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        rays[2 * i] = make_float4(rnd() * 2.f, rnd() * 2.f, 1.0f + rnd(), 0.f);
        rays[2 * i + 1] = make_float4(dx * inv, dy * inv, dz * inv, 0.f);
    }
    float4 *d_recs, *d_rays; uint32_t* d_out;
    cudaMalloc(&d_recs, 4096); cudaMalloc(&d_rays, rays.size() * 16); cudaMalloc(&d_out, (size_t)max_blocks * 128 * 4);
    cudaMemcpy(d_recs, plain.data(), 4096, cudaMemcpyHostToDevice);
    cudaMemcpy(d_rays, rays.data(), rays.size() * 16, cudaMemcpyHostToDevice);
    const float A = 2e-6f, g = 2.44140625e-4f, one = 1.0f;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double peak = (double)sms * 128 * 2 * 1.965e9;
    auto report = [&](const char* name, int bps, float ms, uint32_t chk) {
        const double pairs = (double)sms * bps * 128 * 4 * 64 * iters;
        printf("%-28s %d CTAs/SM: %8.3f ms  %7.1f Gpairs/s  %5.1f %% of FP32 roofline  (chk %08x)\n", name, bps, ms, pairs / ms / 1e6,
               100.0 * pairs * 36.0 / (ms * 1e-3) / peak, chk);
    };
#define RUN(NAME, KERNEL, BPS, FIRST)                                                               \
    {                                                                                               \
        float best = 1e30f;                                                                         \
        for (int r = 0; r < 4; ++r) {                                                               \
            cudaEventRecord(e0);                                                                    \
            KERNEL<<<sms * BPS, 128>>>(FIRST, d_rays, d_out, iters, A, g, one);                     \
            cudaEventRecord(e1); cudaEventSynchronize(e1);                                          \
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;             \
        }                                                                                           \
        uint32_t chk = 0; cudaMemcpy(&chk, d_out + 77, 4, cudaMemcpyDeviceToHost);                  \
        cudaError_t err = cudaGetLastError(); if (err != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(err)); \
        report(NAME, BPS, best, chk);                                                               \
    }
    RUN("smem  mixed e2full guard", (k_smem<2, false, false, 6>), 6, d_recs)
    RUN("smem  front e2dep  guard", (k_smem<0, true, false, 6>), 6, d_recs)
    RUN("smem  mixed e2dep  ftz", (k_smem<2, true, true, 6>), 6, d_recs)
    RUN("smem  front e2dep  ftz", (k_smem<0, true, true, 6>), 6, d_recs)
    RUN("param mixed e2dep  ftz", (k_param<2, true, true, 6>), 6, tp)
    RUN("param front e2dep  ftz", (k_param<0, true, true, 6>), 6, tp)
    RUN("smem  front e2dep  ftz", (k_smem<0, true, true, 5>), 5, d_recs)
    RUN("param front e2dep  ftz", (k_param<0, true, true, 5>), 5, tp)
    RUN("smem  front e2dep  ftz", (k_smem<0, true, true, 4>), 4, d_recs)
    RUN("param front e2dep  ftz", (k_param<0, true, true, 4>), 4, tp)
    RUN("smem  front e2dep  ftz", (k_smem<0, true, true, 3>), 3, d_recs)
    RUN("param front e2dep  ftz", (k_param<0, true, true, 3>), 3, tp)
    RUN("runs 26 planes front u1", (k_runs<0, 1, 6>), 6, tp_runs)
    RUN("runs 26 planes front u2", (k_runs<0, 2, 6>), 6, tp_runs)
    RUN("runs 26 planes mixed u2", (k_runs<2, 2, 6>), 6, tp_runs)
    RUN("runs 64 planes front u2", (k_runs<0, 2, 6>), 6, tp_all)
    RUN("runs 26 planes front u2", (k_runs<0, 2, 5>), 5, tp_runs)
    RUN("runs 26 planes front u4", (k_runs<0, 4, 5>), 5, tp_runs)
    return 0;
}
