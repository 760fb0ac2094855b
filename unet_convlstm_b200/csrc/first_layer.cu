// First convolution of the UNet (DoubleConv `inc`, reference train/unet.py:70 with in_channels = 2 satellites):
// K = 2 x 3 x 3 = 18 -- far too shallow for the tensor-core pipeline (one 16-deep MMA per tap of a zero-padded
// 16-channel tensor ran at ~100 TFLOP/s, bound by the 32-byte-row TMA boxes) and purely HBM-bound: the layer
// writes 64 channels per pixel and reads 2.  CUDA-core kernels:
//   conv_first_fwd   : 4 threads per pixel, 16 output channels each (4 lanes write one 128-byte pixel row),
//                      weights [18][N] fp32 in shared memory, fp32 accumulation, bf16 output
//   conv_first_wgrad : dW[n][c][tap] = sum_p dz[p,n] * x[p+tap,c]; 16 threads per pixel (4 dz channels x 18 taps
//                      = 72 register accumulators each), grid-stride over pixels, block reduction + fp32 atomics
// x: bf16 [IMG][H][W][Cx] (Cx = padded channel count of the activation tensor, the first `cin` are real).
#include "../../include/b200_convlstm.h"
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {


template <int CIN>
__global__ void __launch_bounds__(256) conv_first_fwd_kernel(const __nv_bfloat16* __restrict__ x, int Cx,
                                                             const float* __restrict__ w,     // [N][CIN][3][3]
                                                             const float* __restrict__ bias,  // [N] or nullptr
                                                             __nv_bfloat16* __restrict__ y, long long npix, int H, int W,
                                                             int N) {
    constexpr int K = CIN * 9;
    extern __shared__ float ws[];  // [K][N]
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
        const int k = i / N, n = i - k * N;  // k = c * 9 + tap
        ws[i] = w[static_cast<long long>(n) * K + k];
    }
    __syncthreads();
    const int quads = N / 16;  // threads per pixel
    const long long total = npix * quads;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long p = i / quads;
        const int n0 = static_cast<int>(i - p * quads) * 16;
        const int wq = static_cast<int>(p % W);
        const int hq = static_cast<int>((p / W) % H);
        float xin[K];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int hh = hq + ky - 1, ww = wq + kx - 1;
                const bool in = (hh >= 0) && (hh < H) && (ww >= 0) && (ww < W);
                const __nv_bfloat16* xp = x + (p + (ky - 1) * W + (kx - 1)) * Cx;
#pragma unroll
                for (int c = 0; c < CIN; ++c) xin[c * 9 + ky * 3 + kx] = in ? __bfloat162float(xp[c]) : 0.f;
            }
        }
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = bias ? __ldg(bias + n0 + j) : 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float4* wr = reinterpret_cast<const float4*>(ws + k * N + n0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 wv = wr[q];
                acc[4 * q] = fmaf(xin[k], wv.x, acc[4 * q]);
                acc[4 * q + 1] = fmaf(xin[k], wv.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(xin[k], wv.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(xin[k], wv.w, acc[4 * q + 3]);
            }
        }
        st_bf16x16(y + p * N + n0, acc);
    }
}

// 16 threads per pixel: thread q owns dz channels [4q, 4q+4); accumulators acc[4][K] in registers.
template <int CIN>
__global__ void __launch_bounds__(256) conv_first_wgrad_kernel(const __nv_bfloat16* __restrict__ dz, int N,
                                                               const __nv_bfloat16* __restrict__ x, int Cx, long long npix,
                                                               int H, int W, float* __restrict__ dw /* [N][CIN][3][3] */) {
    constexpr int K = CIN * 9;
    const int tpp = N / 4;                       // threads per pixel
    const int q = threadIdx.x % tpp;             // channel quad of this thread
    const int pl = threadIdx.x / tpp;            // pixel lane within the block
    const int ppb = blockDim.x / tpp;            // pixels per block iteration
    float acc[4][K];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[a][k] = 0.f;
    for (long long p = static_cast<long long>(blockIdx.x) * ppb + pl; p < npix; p += static_cast<long long>(gridDim.x) * ppb) {
        const int wq = static_cast<int>(p % W);
        const int hq = static_cast<int>((p / W) % H);
        const uint2 draw = *reinterpret_cast<const uint2*>(dz + p * N + 4 * q);
        const float d[4] = {__uint_as_float(draw.x << 16), __uint_as_float(draw.x & 0xffff0000u),
                            __uint_as_float(draw.y << 16), __uint_as_float(draw.y & 0xffff0000u)};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int hh = hq + ky - 1, ww = wq + kx - 1;
                const bool in = (hh >= 0) && (hh < H) && (ww >= 0) && (ww < W);
                const __nv_bfloat16* xp = x + (p + (ky - 1) * W + (kx - 1)) * Cx;
#pragma unroll
                for (int c = 0; c < CIN; ++c) {
                    const float xv = in ? __bfloat162float(xp[c]) : 0.f;
#pragma unroll
                    for (int a = 0; a < 4; ++a) acc[a][c * 9 + ky * 3 + kx] = fmaf(d[a], xv, acc[a][c * 9 + ky * 3 + kx]);
                }
            }
        }
    }
    // block reduction over the pixel lanes (shared memory, one slice of K values at a time), then atomics
    extern __shared__ float red[];  // [ppb][tpp * 4]
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a) red[pl * (tpp * 4) + 4 * q + a] = acc[a][k];
        __syncthreads();
        if (threadIdx.x < N) {
            float s = 0.f;
            for (int r = 0; r < ppb; ++r) s += red[r * N + threadIdx.x];
            atomicAdd(dw + static_cast<long long>(threadIdx.x) * K + k, s);
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_conv_first_supported(int cin, int N) {
    return (cin >= 1 && cin <= 4 && N >= 16 && N <= 256 && N % 16 == 0) ? 1 : 0;
}

extern "C" int b200_conv_first_fwd(const void* x, int Cx, int cin, const float* w, const float* bias, void* y, long long IMG,
                                   int H, int W, int N, void* stream) {
    if (!x || !w || !y || IMG <= 0 || H <= 0 || W <= 0 || !b200_conv_first_supported(cin, N) || Cx < cin) {
        set_last_error("b200_conv_first_fwd: bad arguments (cin=%d N=%d Cx=%d)", cin, N, Cx);
        return B200_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(y) & 31) != 0) {
        set_last_error("b200_conv_first_fwd: output must be 32-byte aligned");
        return B200_ERR_ALIGN;
    }
    const long long npix = IMG * H * W;
    const long long total = npix * (N / 16);
    long long grid = (total + 255) / 256;
    const long long cap = 16LL * num_sms();
    if (grid > cap) grid = cap;
    const size_t smem = static_cast<size_t>(cin) * 9 * N * sizeof(float);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const __nv_bfloat16* xs = static_cast<const __nv_bfloat16*>(x);
    __nv_bfloat16* ys = static_cast<__nv_bfloat16*>(y);
    switch (cin) {
        case 1: conv_first_fwd_kernel<1><<<static_cast<unsigned>(grid), 256, smem, st>>>(xs, Cx, w, bias, ys, npix, H, W, N); break;
        case 2: conv_first_fwd_kernel<2><<<static_cast<unsigned>(grid), 256, smem, st>>>(xs, Cx, w, bias, ys, npix, H, W, N); break;
        case 3: conv_first_fwd_kernel<3><<<static_cast<unsigned>(grid), 256, smem, st>>>(xs, Cx, w, bias, ys, npix, H, W, N); break;
        default: conv_first_fwd_kernel<4><<<static_cast<unsigned>(grid), 256, smem, st>>>(xs, Cx, w, bias, ys, npix, H, W, N); break;
    }
    B200_CUDA_CHECK(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_conv_first_wgrad(const void* dz, int N, const void* x, int Cx, int cin, long long IMG, int H, int W,
                                     float* dw, void* stream) {
    if (!dz || !x || !dw || IMG <= 0 || H <= 0 || W <= 0 || !b200_conv_first_supported(cin, N) || Cx < cin || N % 64 != 0 ||
        N > 256) {
        set_last_error("b200_conv_first_wgrad: bad arguments (cin=%d N=%d Cx=%d; N must be a multiple of 64)", cin, N, Cx);
        return B200_ERR_ARG;
    }
    const long long npix = IMG * H * W;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B200_CUDA_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * N * cin * 9, st));
    const int tpp = N / 4, ppb = 256 / tpp;
    long long grid = (npix + ppb - 1) / ppb;
    const long long cap = 4LL * num_sms();
    if (grid > cap) grid = cap;
    const size_t smem = static_cast<size_t>(ppb) * N * sizeof(float);
    const __nv_bfloat16* ds = static_cast<const __nv_bfloat16*>(dz);
    const __nv_bfloat16* xs = static_cast<const __nv_bfloat16*>(x);
    switch (cin) {
        case 1: conv_first_wgrad_kernel<1><<<static_cast<unsigned>(grid), 256, smem, st>>>(ds, N, xs, Cx, npix, H, W, dw); break;
        case 2: conv_first_wgrad_kernel<2><<<static_cast<unsigned>(grid), 256, smem, st>>>(ds, N, xs, Cx, npix, H, W, dw); break;
        case 3: conv_first_wgrad_kernel<3><<<static_cast<unsigned>(grid), 256, smem, st>>>(ds, N, xs, Cx, npix, H, W, dw); break;
        default: conv_first_wgrad_kernel<4><<<static_cast<unsigned>(grid), 256, smem, st>>>(ds, N, xs, Cx, npix, H, W, dw); break;
    }
    B200_CUDA_CHECK(cudaGetLastError());
    return B200_OK;
}
