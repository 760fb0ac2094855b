// Pieces shared by the weight layout-change kernels (pack.cu) and the AdamW kernel that emits the GEMM-operand copies
// of the weight it has just updated (optimizer.cu): the 32 x 32 x taps tile and the description of a packed layout.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int PK_T = 32;        // tile edge over both index dimensions
constexpr int PK_MAX_TAPS = 9;
constexpr int PK_PITCH = PK_T * PK_MAX_TAPS + 1;  // +1: column reads of the tile are conflict-free

struct PackGeom {
    int A, B, taps;             // src [A][B][taps] fp32 contiguous (pack) / dst of the same shape (unpack)
    int a_contig;               // 0: packed[tap'][pa(a)][b]   1: packed[tap'][b][pa(a)]
    int flip;                   // tap' = taps-1-tap (data-gradient weights) instead of tap
    long long tap_pitch;        // elements between consecutive tap' planes of the packed tensor
    long long row_pitch;        // elements between consecutive rows of the packed tensor
    int perm_ch, perm_cht;      // gate interleave of a (ConvLSTM rows): a = g*Ch + nt*cht + j -> (nt*4 + g)*cht + j; 0 = none
};

__device__ __forceinline__ int perm_a(int a, const PackGeom& g) {
    if (g.perm_ch == 0) return a;
    const int gate = a / g.perm_ch, r = a - gate * g.perm_ch;
    const int nt = r / g.perm_cht, j = r - nt * g.perm_cht;
    return (nt * 4 + gate) * g.perm_cht + j;
}

template <typename TOut>
__device__ __forceinline__ TOut to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// Writes the tile (tile[a][b * taps + tap], na x nb valid) of a [A][B][taps] weight to one packed layout.  The row
// permutation and every division are hoisted out of the tap loop (the first version decoded (tap, row) from one counter
// and permuted per store: 236 warp instructions per 32 parameters in the update-and-pack kernel, ncu round 2).
template <typename TOut>
__device__ __forceinline__ void pack_store_tile(const float (*tile)[PK_PITCH], TOut* __restrict__ dst, const PackGeom& g, int a0,
                                                int b0, int na, int nb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!g.a_contig) {
        // one warp per (a, tap) row: 32 consecutive b
        if (lane >= nb) return;
        for (int ta = warp; ta < na; ta += 8) {
            TOut* row = dst + static_cast<long long>(perm_a(a0 + ta, g)) * g.row_pitch + b0 + lane;
            const float* src = &tile[ta][lane * g.taps];
            for (int tap = 0; tap < g.taps; ++tap) {
                const int tp = g.flip ? g.taps - 1 - tap : tap;
                row[tp * g.tap_pitch] = to_out<TOut>(src[tap]);
            }
        }
    } else {
        // one warp per (b, tap) row: 32 consecutive a (the gate interleave keeps runs of perm_cht >= 16 together)
        if (lane >= na) return;
        TOut* col = dst + perm_a(a0 + lane, g);
        for (int tb = warp; tb < nb; tb += 8) {
            TOut* row = col + static_cast<long long>(b0 + tb) * g.row_pitch;
            const float* src = &tile[lane][tb * g.taps];
            for (int tap = 0; tap < g.taps; ++tap) {
                const int tp = g.flip ? g.taps - 1 - tap : tap;
                row[tp * g.tap_pitch] = to_out<TOut>(src[tap]);
            }
        }
    }
}

}  // namespace b200
