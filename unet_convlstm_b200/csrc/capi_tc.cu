// C-ABI entry points of the tensor-core (bf16) path.  Declared in include/b200_convlstm.h.
#include "../../include/b200_convlstm.h"
#include "common.cuh"
#include "conv_tc.cuh"

using namespace b200;

namespace b200 {
int launch_wgrad_tc(const void* dz, int Nz, const void* src, int Csrc, int T, int B, int H, int W,
                    int ksize, float* dw, long long ldk, int koff, cudaStream_t stream);
}

extern "C" int b200_device_error(void) {
    int* f = device_error_flag();
    if (!f) return 0;
    int v = 0;
    if (cudaMemcpy(&v, f, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return B200_ERR_CUDA;
    if (v != 0) cudaMemset(f, 0, sizeof(int));
    return v;
}

extern "C" int b200_conv_tc_supported(int B, int H, int W, int C0, int C1, int N, int lstm) {
    MTile mt;
    if (!plan_mtile(B, H, W, 128, &mt)) return 0;
    if (C0 <= 0 || C0 % 16 != 0 || C1 % 16 != 0) return 0;
    if (pick_block_n(N, lstm ? EPI_LSTM : EPI_STORE) == 0) return 0;
    return 1;
}

extern "C" int b200_conv_tc_fwd(const void* src0, int C0, const void* src1, int C1, int T, int B, int H,
                                int W, const void* wpacked, const float* bias, int N, int ksize,
                                void* dst0, long long ld0, int split, void* dst1, long long ld1,
                                int out_fp32, int relu, int accumulate, void* stream) {
    if (!src0 || !wpacked || !dst0 || T <= 0 || N <= 0 || (ksize & 1) == 0) {
        set_last_error("b200_conv_tc_fwd: bad arguments");
        return B200_ERR_ARG;
    }
    if (split < 0 || split > N || (split < N && !dst1) || split % 16 != 0 || ld0 % 8 != 0 ||
        (split < N && ld1 % 8 != 0)) {
        set_last_error("b200_conv_tc_fwd: bad split/ld (split=%d N=%d ld0=%lld ld1=%lld)", split, N, ld0,
                       ld1);
        return B200_ERR_ARG;
    }
    if (accumulate && !out_fp32) {
        set_last_error("b200_conv_tc_fwd: accumulate requires fp32 output");
        return B200_ERR_ARG;
    }
    ConvTcParams p = {};
    p.T = T; p.B = B; p.H = H; p.W = W;
    p.C0 = C0; p.C1 = C1; p.N = N; p.ksize = ksize;
    p.dst0 = dst0; p.dst1 = dst1; p.ld0 = ld0; p.ld1 = ld1; p.split = split;
    p.out_fp32 = out_fp32; p.relu = relu; p.accumulate = accumulate; p.bias = bias;
    return launch_conv_tc(src0, src1, wpacked, p, EPI_STORE, static_cast<cudaStream_t>(stream));
}

extern "C" int b200_convlstm_cell_fwd_tc(const void* x, int Cin, const void* h_prev, int Ch, int B, int H,
                                         int W, const void* wpacked, const float* bias_packed,
                                         const float* c_prev, float* c_next, void* h_next,
                                         void* gates_out, int ksize, void* stream) {
    if (!x || !wpacked || !c_next || !h_next || Ch <= 0 || Cin <= 0) {
        set_last_error("b200_convlstm_cell_fwd_tc: bad arguments");
        return B200_ERR_ARG;
    }
    ConvTcParams p = {};
    p.T = 1; p.B = B; p.H = H; p.W = W;
    p.C0 = Cin; p.C1 = h_prev ? Ch : 0;
    p.N = 4 * Ch; p.ksize = ksize;
    p.wK = Cin + Ch;
    p.bias = bias_packed;
    p.c_prev = c_prev; p.c_next = c_next;
    p.h_next = static_cast<__nv_bfloat16*>(h_next);
    p.gates_out = static_cast<__nv_bfloat16*>(gates_out);
    // With h_prev == NULL (zero initial state, unet.py:23-25) the h half of K is skipped; the
    // packed weight rows still span Cin+Ch columns, so the weight map must keep the full K.
    return launch_conv_tc(x, h_prev, wpacked, p, EPI_LSTM, static_cast<cudaStream_t>(stream));
}

extern "C" int b200_wgrad_tc(const void* dz, int Nz, const void* src, int Csrc, int T, int B, int H, int W,
                             int ksize, float* dw, long long ldk, int koff, void* stream) {
    if (!dz || !src || !dw || Nz <= 0 || Csrc <= 0 || (ksize & 1) == 0 || koff < 0 || koff + Csrc > ldk) {
        set_last_error("b200_wgrad_tc: bad arguments");
        return B200_ERR_ARG;
    }
    return launch_wgrad_tc(dz, Nz, src, Csrc, T, B, H, W, ksize, dw, ldk, koff,
                           static_cast<cudaStream_t>(stream));
}
