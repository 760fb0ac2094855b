/* C ABI of the B200-native ConvLSTM / UNet-block hot path (libb200convlstm.so).
 *
 * The reference (dordanino12/unet-convlstm) has no native interface: its hot path is the Python
 * nn.Module code in train/unet.py, which lowers to ATen library kernels.  Each entry point below
 * names the reference lines whose arithmetic it replaces.  Conventions:
 *   - every pointer is a device pointer owned by the caller; nothing is allocated or freed here
 *     except one 4-byte per-device watchdog flag;
 *   - activations are NHWC ("channels last"), bf16 on the tensor-core path (*_tc), fp32 or bf16 on
 *     the generic SIMT path; a sequence tensor is [T][B][H][W][C];
 *   - `stream` is a cudaStream_t passed as void*; kernels are launched on it, no hidden sync;
 *   - return value 0 = success, negative = B200_ERR_* below; b200_last_error() gives the text.
 *     No exception ever crosses this boundary.
 */
#ifndef B200_CONVLSTM_H
#define B200_CONVLSTM_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_SHAPE -1
#define B200_ERR_ALIGN -2
#define B200_ERR_CUDA -3
#define B200_ERR_ARG -4
#define B200_ERR_PIPELINE -5

/* Text of the last error raised on the calling thread. */
const char* b200_last_error(void);
/* Reads and clears the device watchdog flag (non-zero: id of the mbarrier a kernel timed out on). */
int b200_device_error(void);

/* 1 if the tcgen05 path can tile this problem (else use the *_simt entry points). */
int b200_conv_tc_supported(int B, int H, int W, int C0, int C1, int N, int lstm);

/* Convolution over the virtual concat [src0 ; src1] (nn.Conv2d(.., k, padding=k//2), unet.py:19,70-71;
 * the torch.cat of unet.py:28 and :98 is never materialised).  wpacked: bf16 [k*k][N][C0+C1].
 * Output columns [0,split) go to dst0 (row stride ld0), [split,N) to dst1 (row stride ld1). */
int b200_conv_tc_fwd(const void* src0, int C0, const void* src1, int C1, int T, int B, int H, int W,
                     const void* wpacked, const float* bias, int N, int ksize, void* dst0,
                     long long ld0, int split, void* dst1, long long ld1, int out_fp32, int relu,
                     int accumulate, void* stream);

/* One ConvLSTM cell step, ConvLSTMCell.forward (unet.py:21-36): gate conv over [x ; h_prev] fused
 * with sigmoid/tanh and the c/h update.  wpacked: bf16 [k*k][4*Ch gate-interleaved][Cin+Ch];
 * bias_packed fp32 [4*Ch] in the same row order (b200_pack_lstm_weights).  h_prev / c_prev may be
 * NULL (zero state, unet.py:23-25).  gates_out (bf16 [P][4][Ch], activated i,f,g,o) may be NULL. */
int b200_convlstm_cell_fwd_tc(const void* x, int Cin, const void* h_prev, int Ch, int B, int H, int W,
                              const void* wpacked, const float* bias_packed, const float* c_prev,
                              float* c_next, void* h_next, void* gates_out, int ksize, void* stream);

/* Weight gradient of a convolution, summed over all T*B images (autograd of nn.Conv2d, unet.py:19
 * and :70-71; for the ConvLSTM this is the BPTT sum over timesteps):
 *   dw[tap][n][koff + k] += sum_{t,p} dz[t,p,n] * src[t, p+tap, k]
 * dz: bf16 [T][B][H][W][Nz], src: bf16 [T][B][H][W][Csrc], dw: fp32 [k*k][Nz][ldk], zero-initialised
 * by the caller (partial sums are combined with fp32 reductions). */
int b200_wgrad_tc(const void* dz, int Nz, const void* src, int Csrc, int T, int B, int H, int W,
                  int ksize, float* dw, long long ldk, int koff, void* stream);

#ifdef __cplusplus
}
#endif
#endif
