// Gradient clipping + AdamW of the reference's training step (main.py:106 `clip_grad_norm_(params, 1.0)`,
// main.py:275 `torch.optim.AdamW`) as multi-tensor kernels: ONE pass over all gradients for the global norm and ONE
// pass over {param, grad, exp_avg, exp_avg_sq} for the update, with the clip coefficient read from device memory
// (no host synchronisation, no separate grad *= coef pass).  The tensor table travels in the kernel parameters
// (48 tensors per launch); a block owns one 4096-element chunk of one tensor.
// HBM-bound: 4 B/param for the norm, 16 B read + 12 B written per param for the update.
#include "../../include/b200_convlstm.h"
#include "common.cuh"
#include "pack.cuh"

namespace b200 {

constexpr int MT_MAX = 48;
constexpr int MT_CHUNK = 4096;  // elements per block: 256 threads x 4 float4

struct MtTable {
    float* p[MT_MAX];
    float* g[MT_MAX];
    float* m[MT_MAX];
    float* v[MT_MAX];
    long long n[MT_MAX];
    int first_chunk[MT_MAX + 1];
    int count;
};

struct AdamCfg {
    float lr, beta1, beta2, eps, weight_decay;
    float step_size;      // lr / (1 - beta1^t)
    float inv_bc2_sqrt;   // 1 / sqrt(1 - beta2^t)
    float max_norm;       // <= 0: no clipping
};

__device__ __forceinline__ int mt_locate(const MtTable& tb, int blk) {
    int t = 0;
    while (t + 1 < tb.count && blk >= tb.first_chunk[t + 1]) ++t;
    return t;
}

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(256) grad_sqnorm_multi_kernel(const __grid_constant__ MtTable tb, double* __restrict__ out) {
    const int t = mt_locate(tb, blockIdx.x);
    const long long off = static_cast<long long>(blockIdx.x - tb.first_chunk[t]) * MT_CHUNK;
    const long long rem = tb.n[t] - off;
    const int len = rem < MT_CHUNK ? static_cast<int>(rem) : MT_CHUNK;
    const float* g = tb.g[t] + off;
    float s = 0.f;
    if (len == MT_CHUNK && aligned16(g)) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(g) + k * 256 + threadIdx.x);
            s = fmaf(a.x, a.x, s), s = fmaf(a.y, a.y, s), s = fmaf(a.z, a.z, s), s = fmaf(a.w, a.w, s);
        }
    } else {
        for (int i = threadIdx.x; i < len; i += 256) s = fmaf(g[i], g[i], s);
    }
    double d = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        atomicAdd(out, tot);
    }
}

// torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
__device__ __forceinline__ float clip_coef(const double* sqnorm, float max_norm) {
    if (!sqnorm || max_norm <= 0.f) return 1.f;
    const float c = max_norm / (static_cast<float>(sqrt(*sqnorm)) + 1e-6f);
    return c < 1.f ? c : 1.f;
}

// Every operation is spelled out with rounding intrinsics: left to the compiler, the multiply-adds were contracted
// differently in the vectorised body, the scalar tail and the tile kernel (adamw_pack_kernel), so the same parameter got
// last-bit different updates depending on which code path touched it.  One arithmetic, three callers, identical bits.
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamCfg& c, float clip) {
    g = __fmul_rn(g, clip);
    p = __fmaf_rn(-__fmul_rn(c.lr, c.weight_decay), p, p);                        // p -= lr * wd * p
    m = __fmaf_rn(__fsub_rn(g, m), __fsub_rn(1.f, c.beta1), m);                   // exp_avg.lerp_(grad, 1 - beta1)
    v = __fmaf_rn(c.beta2, v, __fmul_rn(__fmul_rn(__fsub_rn(1.f, c.beta2), g), g));
    const float denom = __fmaf_rn(__fsqrt_rn(v), c.inv_bc2_sqrt, c.eps);
    p = __fmaf_rn(-c.step_size, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(256) adamw_multi_kernel(const __grid_constant__ MtTable tb, const AdamCfg c,
                                                          const double* __restrict__ sqnorm) {
    const int t = mt_locate(tb, blockIdx.x);
    const long long off = static_cast<long long>(blockIdx.x - tb.first_chunk[t]) * MT_CHUNK;
    const long long rem = tb.n[t] - off;
    const int len = rem < MT_CHUNK ? static_cast<int>(rem) : MT_CHUNK;
    float* p = tb.p[t] + off;
    const float* g = tb.g[t] + off;
    float* m = tb.m[t] + off;
    float* v = tb.v[t] + off;
    const float clip = clip_coef(sqnorm, c.max_norm);
    if (len == MT_CHUNK && aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v)) {
        float4 P[4], G[4], M[4], V[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // all loads before the first use
            const int i = k * 256 + threadIdx.x;
            P[k] = reinterpret_cast<const float4*>(p)[i];
            G[k] = __ldg(reinterpret_cast<const float4*>(g) + i);
            M[k] = reinterpret_cast<const float4*>(m)[i];
            V[k] = reinterpret_cast<const float4*>(v)[i];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = k * 256 + threadIdx.x;
            adamw_one(P[k].x, G[k].x, M[k].x, V[k].x, c, clip);
            adamw_one(P[k].y, G[k].y, M[k].y, V[k].y, c, clip);
            adamw_one(P[k].z, G[k].z, M[k].z, V[k].z, c, clip);
            adamw_one(P[k].w, G[k].w, M[k].w, V[k].w, c, clip);
            reinterpret_cast<float4*>(p)[i] = P[k];
            reinterpret_cast<float4*>(m)[i] = M[k];
            reinterpret_cast<float4*>(v)[i] = V[k];
        }
    } else {
        for (int i = threadIdx.x; i < len; i += 256) {
            float pp = p[i], mm = m[i], vv = v[i];
            adamw_one(pp, g[i], mm, vv, c, clip);
            p[i] = pp, m[i] = mm, v[i] = vv;
        }
    }
}

__global__ void __launch_bounds__(256) grad_scale_multi_kernel(const __grid_constant__ MtTable tb, const double* __restrict__ sqnorm,
                                                               float max_norm) {
    const float clip = clip_coef(sqnorm, max_norm);
    if (clip == 1.f) return;
    const int t = mt_locate(tb, blockIdx.x);
    const long long off = static_cast<long long>(blockIdx.x - tb.first_chunk[t]) * MT_CHUNK;
    const long long rem = tb.n[t] - off;
    const int len = rem < MT_CHUNK ? static_cast<int>(rem) : MT_CHUNK;
    float* g = tb.g[t] + off;
    for (int i = threadIdx.x; i < len; i += 256) g[i] *= clip;
}

// AdamW of convolution weights [A][B][taps] that also EMITS the GEMM-operand copies of the weights it has just updated
// ("fused AdamW that also emits the packed bf16 weights").  A block owns a 32 x 32 x taps tile --
// 32 runs of 32 * taps contiguous floats of param / grad / exp_avg / exp_avg_sq -- updates it with the arithmetic of
// adamw_multi_kernel, keeps the new values in shared memory and stores them in up to two packed layouts (pack.cuh; the
// same stores as pack_weight_kernel).  Multi-tensor: the table of up to AP_MAX weights travels in the kernel parameters
// and blockIdx.x walks the tiles of all of them, so the 64-parameter first layer and the 75 M-parameter cell weight share
// one launch (one launch per weight was latency-bound on the small ones: 40 us each for 4 .. 128 blocks, ncu round 2).
// Replaces the multi-tensor update of these weights plus two or three pack launches per weight and step.
struct PackDst {
    void* dst;      // nullptr: unused
    int fp32;
    PackGeom g;
};

constexpr int AP_MAX = 16;
struct ApEntry {
    float* p;
    const float* g;
    float* m;
    float* v;
    int A, B, taps, tiles_b;
    PackDst d0, d1;
};
struct ApTable {
    int count;
    int first_tile[AP_MAX + 1];
    ApEntry e[AP_MAX];
};
static_assert(sizeof(ApTable) + sizeof(AdamCfg) + 16 <= 4096, "the table must fit the classic 4 KB kernel-parameter space");

constexpr int AP_U = 4;  // elements per thread and batch: every load of a batch is issued before the first use

__global__ void __launch_bounds__(256, 4) adamw_pack_kernel(const __grid_constant__ ApTable tb, const AdamCfg c,
                                                         const double* __restrict__ sqnorm) {
    __shared__ float tile[PK_T][PK_PITCH];
    int t = 0;
    while (t + 1 < tb.count && static_cast<int>(blockIdx.x) >= tb.first_tile[t + 1]) ++t;
    const ApEntry& e = tb.e[t];
    const int tl = blockIdx.x - tb.first_tile[t];
    const int ty = tl / e.tiles_b;
    const int a0 = ty * PK_T, b0 = (tl - ty * e.tiles_b) * PK_T;
    const int na = min(PK_T, e.A - a0), nb = min(PK_T, e.B - b0);
    const int run = nb * e.taps;              // floats of one row of the tile, contiguous in memory
    const int run_p = (run + 31) & ~31;       // rows padded to whole warps: a warp never straddles two rows
    const int total = na * run_p;
    const unsigned inv_run_p = static_cast<unsigned>((0x100000000ULL + run_p - 1) / run_p);  // idx / run_p = umulhi(idx, inv): exact for idx < 2^32 / run_p
    const float clip = clip_coef(sqnorm, c.max_norm);
    float* __restrict__ p = e.p;
    const float* __restrict__ g = e.g;
    float* __restrict__ m = e.m;
    float* __restrict__ v = e.v;
    for (int base = threadIdx.x; base < total; base += 256 * AP_U) {
        long long off[AP_U];
        int ta[AP_U], col[AP_U];
        float pp[AP_U], gg[AP_U], mm[AP_U], vv[AP_U];
#pragma unroll
        for (int u = 0; u < AP_U; ++u) {
            const int idx = base + u * 256;
            ta[u] = static_cast<int>(__umulhi(static_cast<unsigned>(idx), inv_run_p));
            col[u] = idx - ta[u] * run_p;
            const bool ok = idx < total && col[u] < run;
            if (!ok) ta[u] = -1;
            off[u] = ok ? (static_cast<long long>(a0 + ta[u]) * e.B + b0) * e.taps + col[u] : 0;
            if (ok) {
                pp[u] = p[off[u]];
                gg[u] = __ldg(g + off[u]);
                mm[u] = m[off[u]];
                vv[u] = v[off[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < AP_U; ++u) {
            if (ta[u] >= 0) {
                adamw_one(pp[u], gg[u], mm[u], vv[u], c, clip);
                p[off[u]] = pp[u], m[off[u]] = mm[u], v[off[u]] = vv[u];
                tile[ta[u]][col[u]] = pp[u];
            }
        }
    }
    __syncthreads();
    const PackDst* ds[2] = {&e.d0, &e.d1};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (!ds[k]->dst) continue;
        if (ds[k]->fp32)
            pack_store_tile<float>(tile, static_cast<float*>(ds[k]->dst), ds[k]->g, a0, b0, na, nb);
        else
            pack_store_tile<__nv_bfloat16>(tile, static_cast<__nv_bfloat16*>(ds[k]->dst), ds[k]->g, a0, b0, na, nb);
    }
}

// Calls launch(table) for consecutive groups of <= MT_MAX non-empty tensors.
template <class F>
static int for_each_group(int n, float* const* p, float* const* g, float* const* m, float* const* v, const long long* numel,
                          F&& launch) {
    int i = 0;
    while (i < n) {
        MtTable tb;
        tb.count = 0;
        tb.first_chunk[0] = 0;
        while (i < n && tb.count < MT_MAX) {
            if (numel[i] > 0) {
                const long long chunks = (numel[i] + MT_CHUNK - 1) / MT_CHUNK;
                if (tb.first_chunk[tb.count] + chunks > 0x7fffffffLL) return B200_ERR_ARG;
                const int k = tb.count++;
                tb.p[k] = p ? p[i] : nullptr;
                tb.g[k] = g[i];
                tb.m[k] = m ? m[i] : nullptr;
                tb.v[k] = v ? v[i] : nullptr;
                tb.n[k] = numel[i];
                tb.first_chunk[k + 1] = tb.first_chunk[k] + static_cast<int>(chunks);
            }
            ++i;
        }
        if (tb.count > 0) {
            const int rc = launch(tb);
            if (rc != B200_OK) return rc;
        }
    }
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_grad_sqnorm_multi(int n, const void* const* grads, const long long* numel, double* sqnorm, void* stream) {
    if (n < 0 || (n > 0 && (!grads || !numel)) || !sqnorm) {
        set_last_error("b200_grad_sqnorm_multi: bad arguments");
        return B200_ERR_ARG;
    }
    for (int i = 0; i < n; ++i)
        if (numel[i] < 0 || (numel[i] > 0 && !grads[i])) {
            set_last_error("b200_grad_sqnorm_multi: tensor %d: null pointer or negative size", i);
            return B200_ERR_ARG;
        }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B200_CUDA_CHECK(cudaMemsetAsync(sqnorm, 0, sizeof(double), st));
    return for_each_group(n, nullptr, reinterpret_cast<float* const*>(const_cast<void* const*>(grads)), nullptr, nullptr, numel,
                          [&](const MtTable& tb) {
                              grad_sqnorm_multi_kernel<<<tb.first_chunk[tb.count], 256, 0, st>>>(tb, sqnorm);
                              B200_CUDA_CHECK(cudaGetLastError());
                              return B200_OK;
                          });
}

extern "C" int b200_grad_clip_multi(int n, void* const* grads, const long long* numel, const double* sqnorm, float max_norm,
                                    void* stream) {
    if (n < 0 || (n > 0 && (!grads || !numel)) || !sqnorm || !(max_norm > 0.f)) {
        set_last_error("b200_grad_clip_multi: bad arguments");
        return B200_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return for_each_group(n, nullptr, reinterpret_cast<float* const*>(grads), nullptr, nullptr, numel, [&](const MtTable& tb) {
        grad_scale_multi_kernel<<<tb.first_chunk[tb.count], 256, 0, st>>>(tb, sqnorm, max_norm);
        B200_CUDA_CHECK(cudaGetLastError());
        return B200_OK;
    });
}

extern "C" int b200_adamw_multi(int n, void* const* params, const void* const* grads, void* const* exp_avg,
                                void* const* exp_avg_sq, const long long* numel, float lr, float beta1, float beta2, float eps,
                                float weight_decay, long long step, const double* sqnorm, float max_norm, void* stream) {
    if (n < 0 || (n > 0 && (!params || !grads || !exp_avg || !exp_avg_sq || !numel)) || step < 1 || !(beta1 >= 0.f && beta1 < 1.f) ||
        !(beta2 >= 0.f && beta2 < 1.f)) {
        set_last_error("b200_adamw_multi: bad arguments");
        return B200_ERR_ARG;
    }
    for (int i = 0; i < n; ++i)
        if (numel[i] < 0 || (numel[i] > 0 && (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i]))) {
            set_last_error("b200_adamw_multi: tensor %d: null pointer or negative size", i);
            return B200_ERR_ARG;
        }
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
    AdamCfg c{lr, beta1, beta2, eps, weight_decay, static_cast<float>(lr / bc1), static_cast<float>(1.0 / sqrt(bc2)), max_norm};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return for_each_group(n, reinterpret_cast<float* const*>(params), reinterpret_cast<float* const*>(const_cast<void* const*>(grads)),
                          reinterpret_cast<float* const*>(exp_avg), reinterpret_cast<float* const*>(exp_avg_sq), numel,
                          [&](const MtTable& tb) {
                              adamw_multi_kernel<<<tb.first_chunk[tb.count], 256, 0, st>>>(tb, c, sqnorm);
                              B200_CUDA_CHECK(cudaGetLastError());
                              return B200_OK;
                          });
}

static bool pack_dst_ok(const void* dst, int A, long long tap_pitch, long long row_pitch, long long perm_ch, long long perm_cht) {
    if (!dst) return true;
    return tap_pitch >= 0 && row_pitch > 0 && (perm_ch == 0 || (perm_cht > 0 && perm_ch % perm_cht == 0 && A == 4 * perm_ch));
}

static AdamCfg adam_cfg(float lr, float beta1, float beta2, float eps, float weight_decay, long long step, float max_norm) {
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
    return AdamCfg{lr, beta1, beta2, eps, weight_decay, static_cast<float>(lr / bc1), static_cast<float>(1.0 / sqrt(bc2)), max_norm};
}

// geom: dst_fp32, a_contig, flip, tap_pitch, row_pitch, perm_ch, perm_cht
static PackDst pack_dst(void* dst, const long long* geom, int A, int B, int taps) {
    if (!dst) return PackDst{nullptr, 0, PackGeom{A, B, taps, 0, 0, 0, 1, 0, 0}};
    return PackDst{dst, static_cast<int>(geom[0]),
                   PackGeom{A, B, taps, geom[1] != 0, geom[2] != 0, geom[3], geom[4], static_cast<int>(geom[5]), static_cast<int>(geom[6])}};
}

extern "C" int b200_adamw_pack_multi(int n, void* const* params, const void* const* grads, void* const* exp_avg,
                                     void* const* exp_avg_sq, const long long* dims, float lr, float beta1, float beta2,
                                     float eps, float weight_decay, long long step, const double* sqnorm, float max_norm,
                                     void* const* dst0, const long long* geom0, void* const* dst1, const long long* geom1,
                                     void* stream) {
    if (n < 0 || (n > 0 && (!params || !grads || !exp_avg || !exp_avg_sq || !dims || !dst0 || !geom0 || !dst1 || !geom1)) ||
        step < 1 || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f)) {
        set_last_error("b200_adamw_pack_multi: bad arguments");
        return B200_ERR_ARG;
    }
    for (int i = 0; i < n; ++i) {
        const long long A = dims[3 * i], B = dims[3 * i + 1], taps = dims[3 * i + 2];
        const long long* g0 = geom0 + 7 * i;
        const long long* g1 = geom1 + 7 * i;
        if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i] || A <= 0 || B <= 0 || taps <= 0 || taps > PK_MAX_TAPS ||
            A > 0x7fffffffLL || B > 0x7fffffffLL || !pack_dst_ok(dst0[i], static_cast<int>(A), g0[3], g0[4], g0[5], g0[6]) ||
            !pack_dst_ok(dst1[i], static_cast<int>(A), g1[3], g1[4], g1[5], g1[6])) {
            set_last_error("b200_adamw_pack_multi: weight %d: bad arguments (taps <= %d; the gate interleave needs A = 4*Ch, Ch %% cht = 0)",
                           i, PK_MAX_TAPS);
            return B200_ERR_ARG;
        }
    }
    const AdamCfg c = adam_cfg(lr, beta1, beta2, eps, weight_decay, step, max_norm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int i = 0;
    while (i < n) {
        ApTable tb;
        tb.count = 0;
        tb.first_tile[0] = 0;
        while (i < n && tb.count < AP_MAX) {
            const int A = static_cast<int>(dims[3 * i]), B = static_cast<int>(dims[3 * i + 1]), taps = static_cast<int>(dims[3 * i + 2]);
            const long long tiles_b = (B + PK_T - 1) / PK_T, tiles_a = (A + PK_T - 1) / PK_T;
            if (tb.first_tile[tb.count] + tiles_a * tiles_b > 0x7fffffffLL) {
                set_last_error("b200_adamw_pack_multi: too many tiles");
                return B200_ERR_SHAPE;
            }
            ApEntry& e = tb.e[tb.count];
            e.p = static_cast<float*>(params[i]);
            e.g = static_cast<const float*>(grads[i]);
            e.m = static_cast<float*>(exp_avg[i]);
            e.v = static_cast<float*>(exp_avg_sq[i]);
            e.A = A; e.B = B; e.taps = taps; e.tiles_b = static_cast<int>(tiles_b);
            e.d0 = pack_dst(dst0[i], geom0 + 7 * i, A, B, taps);
            e.d1 = pack_dst(dst1[i], geom1 + 7 * i, A, B, taps);
            tb.first_tile[tb.count + 1] = tb.first_tile[tb.count] + static_cast<int>(tiles_a * tiles_b);
            ++tb.count;
            ++i;
        }
        adamw_pack_kernel<<<tb.first_tile[tb.count], 256, 0, st>>>(tb, c, sqnorm);
        B200_CUDA_CHECK(cudaGetLastError());
    }
    return B200_OK;
}

extern "C" int b200_adamw_pack(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int A, int B, int taps,
                               float lr, float beta1, float beta2, float eps, float weight_decay, long long step,
                               const double* sqnorm, float max_norm, void* dst0, int dst0_fp32, int a_contig0, int flip0,
                               long long tap_pitch0, long long row_pitch0, int perm_ch0, int perm_cht0, void* dst1,
                               int dst1_fp32, int a_contig1, int flip1, long long tap_pitch1, long long row_pitch1,
                               int perm_ch1, int perm_cht1, void* stream) {
    void* ps[1] = {param};
    const void* gs[1] = {grad};
    void* ms[1] = {exp_avg};
    void* vs[1] = {exp_avg_sq};
    void* d0[1] = {dst0};
    void* d1[1] = {dst1};
    const long long dims[3] = {A, B, taps};
    const long long g0[7] = {dst0_fp32, a_contig0, flip0, tap_pitch0, row_pitch0, perm_ch0, perm_cht0};
    const long long g1[7] = {dst1_fp32, a_contig1, flip1, tap_pitch1, row_pitch1, perm_ch1, perm_cht1};
    return b200_adamw_pack_multi(1, ps, gs, ms, vs, dims, lr, beta1, beta2, eps, weight_decay, step, sqnorm, max_norm, d0, g0, d1,
                                 g1, stream);
}
