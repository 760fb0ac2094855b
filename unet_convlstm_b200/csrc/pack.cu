// Weight layout changes between the reference's parameter layout (OIHW / IOHW fp32, the tap index fastest) and the
// GEMM operand layouts of the tensor-core kernels ([tap][row][col], bf16 or fp32), and back for the weight
// gradients.  They run once per weight and training step (after every optimizer update); the generic strided copy
// reads the 9 taps of one (n, k) pair with a stride of taps floats per thread (0.7 TB/s on the 75 M-parameter cell
// weight), so here a block stages a 32 x 32 x taps tile through shared memory: coalesced reads of whole
// [32 * taps] float runs, coalesced writes of 32 consecutive columns per (tap, row).
// HBM-bound: 4 B read + 2 B written per parameter (pack), 4 + 4 (unpack).
#include "../../include/b200_convlstm.h"
#include "pack.cuh"

namespace b200 {

// grid (ceil(B/32), ceil(A/32)), 256 threads
template <typename TOut>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ src, TOut* __restrict__ dst, PackGeom g) {
    __shared__ float tile[PK_T][PK_PITCH];
    const int a0 = blockIdx.y * PK_T, b0 = blockIdx.x * PK_T;
    const int na = min(PK_T, g.A - a0), nb = min(PK_T, g.B - b0);
    const int run = nb * g.taps;  // floats of one a row of the tile, contiguous in src
    for (int ta = threadIdx.x >> 5; ta < na; ta += 8) {
        const float* s = src + (static_cast<long long>(a0 + ta) * g.B + b0) * g.taps;
        for (int i = threadIdx.x & 31; i < run; i += 32) tile[ta][i] = __ldg(s + i);
    }
    __syncthreads();
    pack_store_tile<TOut>(tile, dst, g, a0, b0, na, nb);
}

// packed fp32 [taps][A][ldb] (columns koff.. of each row) -> dst fp32 [A][B][taps];  grid (ceil(B/32), ceil(A/32))
__global__ void __launch_bounds__(256) unpack_wgrad_kernel(const float* __restrict__ packed, float* __restrict__ dst, int A, int B,
                                                           int taps, long long ldb) {
    __shared__ float tile[PK_T][PK_PITCH];
    const int a0 = blockIdx.y * PK_T, b0 = blockIdx.x * PK_T;
    const int na = min(PK_T, A - a0), nb = min(PK_T, B - b0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < taps * na; r += 8) {
        const int tap = r / na, ta = r - tap * na;
        if (lane < nb) tile[ta][lane * taps + tap] = __ldg(packed + (static_cast<long long>(tap) * A + a0 + ta) * ldb + b0 + lane);
    }
    __syncthreads();
    const int run = nb * taps;
    for (int ta = warp; ta < na; ta += 8) {
        float* d = dst + (static_cast<long long>(a0 + ta) * B + b0) * taps;
        for (int i = lane; i < run; i += 32) d[i] = tile[ta][i];
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_pack_weight(const float* src, int A, int B, int taps, void* dst, int dst_fp32, int a_contig, int flip,
                                long long tap_pitch, long long row_pitch, int perm_ch, int perm_cht, void* stream) {
    if (!src || !dst || A <= 0 || B <= 0 || taps <= 0 || taps > PK_MAX_TAPS || tap_pitch < 0 || row_pitch <= 0 ||
        (perm_ch != 0 && (perm_cht <= 0 || perm_ch % perm_cht != 0 || A != 4 * perm_ch))) {
        set_last_error("b200_pack_weight: bad arguments (taps <= %d; the gate interleave needs A = 4*Ch, Ch %% cht = 0)", PK_MAX_TAPS);
        return B200_ERR_ARG;
    }
    PackGeom g{A, B, taps, a_contig != 0, flip != 0, tap_pitch, row_pitch, perm_ch, perm_cht};
    dim3 grid((B + PK_T - 1) / PK_T, (A + PK_T - 1) / PK_T);
    if (grid.y > 65535) {
        set_last_error("b200_pack_weight: A too large");
        return B200_ERR_SHAPE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dst_fp32)
        pack_weight_kernel<float><<<grid, 256, 0, st>>>(src, static_cast<float*>(dst), g);
    else
        pack_weight_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), g);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_unpack_wgrad(const float* packed, int A, int B, int taps, long long ldb, float* dst, void* stream) {
    if (!packed || !dst || A <= 0 || B <= 0 || taps <= 0 || taps > PK_MAX_TAPS || ldb < B) {
        set_last_error("b200_unpack_wgrad: bad arguments");
        return B200_ERR_ARG;
    }
    dim3 grid((B + PK_T - 1) / PK_T, (A + PK_T - 1) / PK_T);
    if (grid.y > 65535) {
        set_last_error("b200_unpack_wgrad: A too large");
        return B200_ERR_SHAPE;
    }
    unpack_wgrad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, dst, A, B, taps, ldb);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200_OK;
}
