// rt_image.cu — the image finishers of main(): post_process (main.rs:748-762: divide the image by its 99th-percentile
// luma) and the linear -> sRGB u8 encode (image.rs:55-66, palette 0.4 Srgb::from_linear + into_format::<u8>) on the
// device, so a frame can stay in HBM from the accumulators to the bytes a PNG writer takes.
//
// post_process is exact: the reference sorts all NORMAL lumas and takes the element at index (len as f32 * 0.99) as
// usize.  The same element is found without sorting by a 4-pass radix select over the order-preserving integer
// image of the floats (8-bit digits, most significant first): each pass histograms one digit of the lumas that
// match the prefix found so far (HBM-bound: 12 B per pixel per pass, luma recomputed in the reference's
// evaluation order), one small block picks the bucket that holds the wanted rank.  Then every channel is divided
// by that value (IEEE division) if it exceeds f32::EPSILON.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "rt_math.cuh"
#include "rt_types.h"

namespace b200rt {

namespace {

struct SelectState {          // device-resident control block of one post_process call
    unsigned int hist[256];
    unsigned int prefix;      // digits found so far (high bits of the key)
    unsigned int rank;        // wanted rank among the keys that match the prefix
    unsigned int n_normal;
    float p98;
};

// palette 0.4 into_luma: Y row of the linear sRGB -> XYZ (D65) matrix, evaluated left to right, non-fused
RT_DI float luma_of(const float* __restrict__ rgb, size_t i) {
    return (rgb[3 * i] * 0.2126729f + rgb[3 * i + 1] * 0.7151522f) + rgb[3 * i + 2] * 0.0721750f;
}
// order-preserving map float -> uint (total order of partial_cmp on non-NaN values)
RT_DI unsigned int key_of(float v) {
    const unsigned int b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
RT_DI float value_of(unsigned int k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

}  // namespace

__global__ void pp_clear_kernel(SelectState* st) {
    if (threadIdx.x < 256) st->hist[threadIdx.x] = 0u;
    if (threadIdx.x == 0) { st->prefix = 0u; st->rank = 0u; st->n_normal = 0u; st->p98 = 0.0f; }
}

// pass = 0..3: histogram of digit `pass` (from the top) over the normal lumas whose higher digits equal the prefix
__global__ void __launch_bounds__(256) pp_hist_kernel(const float* __restrict__ rgb, size_t n, int pass, SelectState* st) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const unsigned int prefix = st->prefix;
    const unsigned int mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float l = luma_of(rgb, i);
        if (!is_normal_f32(l)) continue;                                  // main.rs:751
        const unsigned int k = key_of(l);
        if ((k & mask) == (prefix & mask)) atomicAdd(&h[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

// one block: pick the bucket that holds the wanted rank, extend the prefix, clear the histogram for the next pass
__global__ void pp_pick_kernel(int pass, SelectState* st) {
    if (threadIdx.x != 0) return;
    if (pass == 0) {
        unsigned int total = 0u;
        for (int b = 0; b < 256; ++b) total += st->hist[b];
        st->n_normal = total;
        if (total == 0u) { st->rank = 0u; st->p98 = 0.0f; for (int b = 0; b < 256; ++b) st->hist[b] = 0u; return; }
        unsigned int idx = (unsigned int)((float)total * 0.99f);          // (len as f32 * 0.99) as usize, main.rs:754
        if (idx >= total) idx = total - 1u;
        st->rank = idx;
    }
    if (st->n_normal == 0u) return;
    unsigned int r = st->rank, b = 0u;
    for (; b < 255u; ++b) {
        if (r < st->hist[b]) break;
        r -= st->hist[b];
    }
    const int shift = 24 - 8 * pass;
    st->prefix |= b << shift;
    st->rank = r;
    for (int q = 0; q < 256; ++q) st->hist[q] = 0u;
    if (pass == 3) st->p98 = value_of(st->prefix);
}

__global__ void pp_scale_kernel(float* __restrict__ rgb, size_t n_values, const SelectState* __restrict__ st, float* __restrict__ p98_out) {
    const float p98 = st->n_normal ? st->p98 : 0.0f;
    const bool scale = st->n_normal && p98 > kF32Epsilon;                 // main.rs:755
    if (blockIdx.x == 0 && threadIdx.x == 0 && p98_out) *p98_out = scale ? p98 : 0.0f;
    if (!scale) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_values; i += (size_t)gridDim.x * blockDim.x)
        rgb[i] = rgb[i] / p98;                                            // main.rs:757
}

// image.rs:55-66 (palette: x <= 0.0031308 ? 12.92 x : 1.055 x^(1/2.4) - 0.055; clamp; round to u8).  powf is CUDA's
// (<= 2 ulp): a value within an ulp of a rounding boundary can differ from glibc's by one code.
__global__ void encode_srgb8_kernel(const float* __restrict__ v, size_t n_values, uint8_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_values; i += (size_t)gridDim.x * blockDim.x) {
        const float x = v[i];
        const float e = x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
        float s = e * 255.0f;
        s = s < 0.0f ? 0.0f : (s > 255.0f ? 255.0f : s);
        if (isnan(s)) s = 0.0f;
        out[i] = (uint8_t)roundf(s);
    }
}

size_t post_process_workspace_bytes() { return sizeof(SelectState); }

cudaError_t launch_post_process(float* d_rgb, size_t n_pixels, void* d_workspace, float* d_p98_out, int sm_count, cudaStream_t stream) {
    SelectState* st = static_cast<SelectState*>(d_workspace);
    pp_clear_kernel<<<1, 256, 0, stream>>>(st);
    if (n_pixels) {
        const unsigned blocks = (unsigned)std::min<size_t>((n_pixels + 255) / 256, (size_t)sm_count * 8);
        for (int pass = 0; pass < 4; ++pass) {
            pp_hist_kernel<<<blocks, 256, 0, stream>>>(d_rgb, n_pixels, pass, st);
            pp_pick_kernel<<<1, 32, 0, stream>>>(pass, st);
        }
        pp_scale_kernel<<<(unsigned)std::min<size_t>((3 * n_pixels + 255) / 256, (size_t)sm_count * 8), 256, 0, stream>>>(d_rgb, 3 * n_pixels, st, d_p98_out);
    } else {
        pp_scale_kernel<<<1, 32, 0, stream>>>(d_rgb, 0, st, d_p98_out);
    }
    return cudaGetLastError();
}

cudaError_t launch_encode_srgb8(const float* d_rgb, size_t n_values, uint8_t* d_out, int sm_count, cudaStream_t stream) {
    if (n_values == 0) return cudaSuccess;
    encode_srgb8_kernel<<<(unsigned)std::min<size_t>((n_values + 255) / 256, (size_t)sm_count * 8), 256, 0, stream>>>(d_rgb, n_values, d_out);
    return cudaGetLastError();
}

// dev / test: color_pow (rt_math.cuh) element by element, so that its error bound can be measured against f64 pow
__global__ void color_pow_kernel(const float* __restrict__ x, const float* __restrict__ e, float* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = color_pow(x[i], e[i]);
}
cudaError_t launch_color_pow(const float* d_x, const float* d_e, float* d_out, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    color_pow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_x, d_e, d_out, n);
    return cudaGetLastError();
}

}  // namespace b200rt
