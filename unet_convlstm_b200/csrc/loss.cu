// Training loss of the reference (main.py:28-72, `compute_loss`): weighted L1 + 0.005 x spatial-gradient
// loss, masked means with a 1e-8 epsilon (or plain means without a mask), as TWO passes over the fp32
// prediction / target / mask maps instead of ~25 element-wise launches forward and ~40 backward:
//   b200_wl1_grad_loss_fwd : the four global sums  S0 = sum |d| m w,  S1 = sum m w,  S2 = sum (|dxe|+|dye|) mc,
//                            S3 = sum mc   (d = y_pred - y, w = 1 + 4|y|^3, e = d, crop [H-1, W-1]) + the loss
//   b200_wl1_grad_loss_bwd : d loss / d y_pred from the same maps and the stored sums
// HBM-bound: 12 bytes per element forward, 16 backward; fp32 partials per thread, fp64 block sums and atomics.
#include "../../include/b200_convlstm.h"
#include "common.cuh"

namespace b200 {

struct LossGeom {
    long long n;  // IMG * H * W
    int H, W;
    int masked;
};

__global__ void __launch_bounds__(256) wl1_grad_loss_reduce_kernel(const float* __restrict__ yp, const float* __restrict__ y,
                                                                   const float* __restrict__ mask, LossGeom g,
                                                                   double* __restrict__ sums) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < g.n; i += stride) {
        const int w = static_cast<int>(i % g.W);
        const int h = static_cast<int>((i / g.W) % g.H);
        const float yt = __ldg(y + i);
        const float e = __ldg(yp + i) - yt;
        const float m = g.masked ? __ldg(mask + i) : 1.f;
        const float a = fabsf(yt);
        const float wt = fmaf(4.f * a * a, a, 1.f);
        s0 = fmaf(fabsf(e) * m, wt, s0);
        s1 = fmaf(m, wt, s1);
        if (h + 1 < g.H && w + 1 < g.W) {
            const float er = __ldg(yp + i + 1) - __ldg(y + i + 1);
            const float ed = __ldg(yp + i + g.W) - __ldg(y + i + g.W);
            s2 = fmaf(fabsf(er - e) + fabsf(ed - e), m, s2);
            s3 += m;
        }
    }
    __shared__ double red[4][8];
    double v[4] = {s0, s1, s2, s3};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) red[k][warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int wi = 0; wi < 8; ++wi) t += red[threadIdx.x][wi];
        atomicAdd(sums + threadIdx.x, t);
    }
}

// sums[4] -> sums[4..5] = the two denominators actually used, loss
__global__ void wl1_grad_loss_finalize_kernel(double* sums, LossGeom g, long long n_crop, float* loss) {
    const double den1 = g.masked ? sums[1] + 1e-8 : static_cast<double>(g.n);
    const double den2 = g.masked ? sums[3] + 1e-8 : static_cast<double>(n_crop);
    sums[4] = den1;
    sums[5] = n_crop > 0 ? den2 : 1.0;  // H == 1 or W == 1: no crop cell, no gradient term
    *loss = static_cast<float>(sums[0] / den1 + 0.005 * (n_crop > 0 ? sums[2] / den2 : 0.0));
}

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(256) wl1_grad_loss_bwd_kernel(const float* __restrict__ yp, const float* __restrict__ y,
                                                                const float* __restrict__ mask, LossGeom g,
                                                                const double* __restrict__ sums,
                                                                const float* __restrict__ gout, float* __restrict__ dyp) {
    const float go = gout ? __ldg(gout) : 1.f;
    const float k1 = go / static_cast<float>(sums[4]);
    const float k2 = go * 0.005f / static_cast<float>(sums[5]);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < g.n; i += stride) {
        const int w = static_cast<int>(i % g.W);
        const int h = static_cast<int>((i / g.W) % g.H);
        const float yt = __ldg(y + i);
        const float e = __ldg(yp + i) - yt;
        const float m = g.masked ? __ldg(mask + i) : 1.f;
        const float a = fabsf(yt);
        float gsum = sgn(e) * m * fmaf(4.f * a * a, a, 1.f) * k1;
        float gg = 0.f;
        if (h + 1 < g.H && w + 1 < g.W) {  // this pixel is the base of a crop cell
            const float er = __ldg(yp + i + 1) - __ldg(y + i + 1);
            const float ed = __ldg(yp + i + g.W) - __ldg(y + i + g.W);
            gg -= (sgn(er - e) + sgn(ed - e)) * m;
        }
        if (w > 0 && h + 1 < g.H) {        // right neighbour of the cell based at (h, w-1)
            const float el = __ldg(yp + i - 1) - __ldg(y + i - 1);
            gg += sgn(e - el) * (g.masked ? __ldg(mask + i - 1) : 1.f);
        }
        if (h > 0 && w + 1 < g.W) {        // lower neighbour of the cell based at (h-1, w)
            const float eu = __ldg(yp + i - g.W) - __ldg(y + i - g.W);
            gg += sgn(e - eu) * (g.masked ? __ldg(mask + i - g.W) : 1.f);
        }
        dyp[i] = fmaf(gg, k2, gsum);
    }
}

}  // namespace b200

using namespace b200;

static unsigned loss_grid(long long n) {
    long long gsz = (n + 255) / 256;
    const long long cap = 8LL * num_sms();
    if (gsz > cap) gsz = cap;
    return static_cast<unsigned>(gsz < 1 ? 1 : gsz);
}

extern "C" int b200_wl1_grad_loss_fwd(const float* y_pred, const float* y, const float* mask, long long IMG, int H, int W,
                                      double* sums, float* loss, void* stream) {
    if (!y_pred || !y || !sums || !loss || IMG <= 0 || H <= 0 || W <= 0) {
        set_last_error("b200_wl1_grad_loss_fwd: bad arguments");
        return B200_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LossGeom g{IMG * H * W, H, W, mask != nullptr};
    B200_CUDA_CHECK(cudaMemsetAsync(sums, 0, 6 * sizeof(double), st));
    wl1_grad_loss_reduce_kernel<<<loss_grid(g.n), 256, 0, st>>>(y_pred, y, mask, g, sums);
    wl1_grad_loss_finalize_kernel<<<1, 1, 0, st>>>(sums, g, IMG * (H - 1) * (W - 1), loss);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_wl1_grad_loss_bwd(const float* y_pred, const float* y, const float* mask, long long IMG, int H, int W,
                                      const double* sums, const float* grad_out, float* d_y_pred, void* stream) {
    if (!y_pred || !y || !sums || !d_y_pred || IMG <= 0 || H <= 0 || W <= 0) {
        set_last_error("b200_wl1_grad_loss_bwd: bad arguments");
        return B200_ERR_ARG;
    }
    LossGeom g{IMG * H * W, H, W, mask != nullptr};
    wl1_grad_loss_bwd_kernel<<<loss_grid(g.n), 256, 0, static_cast<cudaStream_t>(stream)>>>(y_pred, y, mask, g, sums,
                                                                                           grad_out, d_y_pred);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200_OK;
}
