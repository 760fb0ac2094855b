// Weight-gradient GEMM on the sm_100a tensor cores (conv backward-weight, batched over the
// whole sequence / frame batch):
//
//   dW[tap][n][koff + k] (+)= sum_{t, pixel p} dz[t, p, n] * src[t, p + tap, k]
//
// i.e. the autograd weight gradient of nn.Conv2d (reference train/unet.py:19 for the ConvLSTM
// gate conv -- summed over all T timesteps of BPTT --, :70-71 for the UNet blocks).
//
// GEMM view: M = 128 rows of dz channels, N = up to 256 source channels, reduction over pixels.
// Both operands are read straight from their NHWC tensors, where the *channel* (M resp. N)
// index is contiguous and the pixel (reduction) index is strided: both UMMA operands are
// therefore "MN-major".  A TMA box {cw channels, Wt, Ht, Bt} lands in smem as [pixel][cw] rows
// with the 128/64/32-byte swizzle, which is exactly the canonical MN-major layout
//     ((8 elem, cw/8, m), (8 pixels, k)) : ((1, 8, LBO), (cw, SBO))
// with SBO = 8 pixel rows and LBO = one whole box (next channel chunk).  The tap shift and the
// zero padding are the TMA coordinates / out-of-bounds fill of the source box, as in conv_tc.cu.
//
// Stacked mode (source with 16/32/64 channels, the full-resolution UNet layers): an M tile of
// 128 dz channels would be half empty and every tap would re-read dz, so the roles are swapped:
// A = G = 128/Csrc tap-shifted source boxes stacked along M (rows = (tap, source channel)),
// B = dz channels (N).  One unit then covers G taps with one dz load per K block.
//
// Work units = (reduction split s) x (tap, M tile, N tile); partial sums of different splits
// are combined with vectorised fp32 reductions (red.global.add.v4.f32) into the caller's
// zero-initialised buffer.  Warp roles are those of conv_tc.cu.
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

static constexpr int WG_BLOCK_M = 128;
static constexpr int WG_RB = 64;  // pixels per pipeline stage
static constexpr int WG_THREADS = 256;

struct WgradParams {
    int T, B, H, W;
    int Nz;          // channels of dz (rows of dW)
    int Csrc;        // channels of the source (columns written)
    int ksize, pad;
    int cwA, cwB;    // channel chunk widths (64/32/16)
    int Wt, Ht, Bt, tiles_w, tiles_h, tiles_b;  // 64-pixel box geometry
    int num_rblocks; // T * tiles_w * tiles_h * tiles_b
    int rb_per_split, splits;
    int num_m_tiles, num_n_tiles, out_tiles;
    int stacked;     // 1: A = tap-stacked source boxes (G taps x Csrc = 128 rows), B = dz; see below
    int G;           // taps per unit in stacked mode
    float* dw;       // [taps][Nz][ldk]
    long long ldk;
    int koff;
    int* err_flag;
};

template <int BLOCK_N>
struct WgCfg {
    static constexpr int A_BYTES = WG_BLOCK_M * WG_RB * 2;  // 16 KB
    static constexpr int B_BYTES = BLOCK_N * WG_RB * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = 2 * BLOCK_N;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}

__device__ __forceinline__ void red_add_f32(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

template <int BLOCK_N>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_src,
                const WgradParams p) {
    using Cfg = WgCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_units = p.out_tiles * p.splits;
    const int boxesA = WG_BLOCK_M / p.cwA;
    const uint32_t boxA_bytes = WG_RB * p.cwA * 2;
    const uint32_t boxB_bytes = WG_RB * p.cwB * 2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_dz);
        prefetch_tmap(&tm_src);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    // unit -> (split, tap, m tile, n tile); n columns actually present in this N tile
    auto decode = [&](int unit, int& split, int& tap, int& m0, int& n0, int& ncols) {
        split = unit / p.out_tiles;
        int ot = unit - split * p.out_tiles;
        int nt = ot % p.num_n_tiles;
        ot /= p.num_n_tiles;
        int mt = ot % p.num_m_tiles;
        tap = ot / p.num_m_tiles;
        m0 = mt * WG_BLOCK_M;
        n0 = nt * BLOCK_N;
        // stacked mode: "tap" is the tap group, N runs over dz channels
        ncols = min(BLOCK_N, (p.stacked ? p.Nz : p.Csrc) - n0);
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
                int split, tap, m0, n0, ncols;
                decode(unit, split, tap, m0, n0, ncols);
                const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
                const int boxesB = (ncols + p.cwB - 1) / p.cwB;
                const int rb_begin = split * p.rb_per_split;
                const int rb_end = min(rb_begin + p.rb_per_split, p.num_rblocks);
                for (int rb = rb_begin; rb < rb_end; ++rb) {
                    int m = rb;
                    const int wt = m % p.tiles_w;
                    m /= p.tiles_w;
                    const int ht = m % p.tiles_h;
                    m /= p.tiles_h;
                    const int bt = m % p.tiles_b;
                    const int t = m / p.tiles_b;
                    const int w0 = wt * p.Wt, h0 = ht * p.Ht, b0 = bt * p.Bt;
                    mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 500 + stage);
                    mbar_arrive_expect_tx(full_bar(stage), boxesA * boxA_bytes + boxesB * boxB_bytes);
                    const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t b_dst = a_dst + Cfg::A_BYTES;
                    if (p.stacked) {
                        const int taps = p.ksize * p.ksize;
                        for (int i = 0; i < boxesA; ++i) {
                            // taps beyond the filter repeat the last one; their rows are never stored
                            const int tp = min(tap * p.G + i, taps - 1);
                            const int sy = tp / p.ksize, sx = tp - sy * p.ksize;
                            tma_load_5d(a_dst + i * boxA_bytes, &tm_src, full_bar(stage), 0, w0 + sx - p.pad,
                                        h0 + sy - p.pad, b0, t);
                        }
                        for (int i = 0; i < boxesB; ++i)
                            tma_load_5d(b_dst + i * boxB_bytes, &tm_dz, full_bar(stage), n0 + i * p.cwB, w0, h0,
                                        b0, t);
                    } else {
                        for (int i = 0; i < boxesA; ++i)
                            tma_load_5d(a_dst + i * boxA_bytes, &tm_dz, full_bar(stage), m0 + i * p.cwA, w0,
                                        h0, b0, t);
                        for (int i = 0; i < boxesB; ++i)
                            tma_load_5d(b_dst + i * boxB_bytes, &tm_src, full_bar(stage), n0 + i * p.cwB,
                                        w0 + kx - p.pad, h0 + ky - p.pad, b0, t);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t ltA = (p.cwA == 64) ? 2u : (p.cwA == 32 ? 4u : 6u);
            const uint32_t ltB = (p.cwB == 64) ? 2u : (p.cwB == 32 ? 4u : 6u);
            const uint32_t sboA = 8u * p.cwA * 2, sboB = 8u * p.cwB * 2;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
                int split, tap, m0, n0, ncols;
                decode(unit, split, tap, m0, n0, ncols);
                const int nmma = ((ncols + p.cwB - 1) / p.cwB) * p.cwB;  // whole loaded boxes
                const uint32_t idesc = make_idesc_bf16(WG_BLOCK_M, nmma, 1, 1);
                const int rb_begin = split * p.rb_per_split;
                const int rb_end = min(rb_begin + p.rb_per_split, p.num_rblocks);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err_flag, 700 + acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int rb = rb_begin; rb < rb_end; ++rb) {
                    mbar_wait(full_bar(stage), phase, p.err_flag, 600 + stage);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < WG_RB / 16; ++k) {
                        // 16 pixels = two 8-row atoms further down the box
                        const uint64_t adesc =
                            make_smem_desc(a_addr + k * 16 * p.cwA * 2, boxA_bytes, sboA, ltA);
                        const uint64_t bdesc =
                            make_smem_desc(b_addr + k * 16 * p.cwB * 2, boxB_bytes, sboB, ltB);
                        umma_bf16(d_tmem, adesc, bdesc, idesc, (rb > rb_begin || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull_bar(acc));
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int r = q * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
            int split, tap, m0, n0, ncols;
            decode(unit, split, tap, m0, n0, ncols);
            const int co = m0 + r;
            bool valid = co < p.Nz;
            float* row = p.dw + (static_cast<long long>(tap) * p.Nz + co) * p.ldk + p.koff + n0;
            if (p.stacked) {
                // row r = (tap within the group, source channel); column = dz channel
                const int g = r / p.Csrc, c = r - g * p.Csrc;
                const int tp = tap * p.G + g;
                valid = tp < p.ksize * p.ksize;
                row = p.dw + (static_cast<long long>(tp) * p.Nz + n0) * p.ldk + p.koff + c;
            }
            mbar_wait(tfull_bar(acc), acc_phase, p.err_flag, 800 + acc);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BLOCK_N + (uint32_t(q * 32) << 16);
#pragma unroll 1
            for (int c16 = 0; c16 * 16 < ncols; ++c16) {
                uint32_t v[16];
                tmem_ld16(t_row + c16 * 16, v);
                tmem_ld_wait();
                if (valid && p.stacked) {
                    // transposed store: consecutive lanes hold consecutive source channels
                    float* o = row + static_cast<long long>(c16) * 16 * p.ldk;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (p.splits > 1)
                            red_add_f32(o + j * p.ldk, __uint_as_float(v[j]));
                        else
                            o[j * p.ldk] = __uint_as_float(v[j]);
                    }
                } else if (valid) {
                    float* o = row + c16 * 16;
                    if (p.splits > 1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            red_add_v4(o + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            reinterpret_cast<float4*>(o)[j] =
                                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int BLOCK_N>
static int launch_wgrad_impl(const CUtensorMap& tz, const CUtensorMap& ts, const WgradParams& p,
                             cudaStream_t stream) {
    using Cfg = WgCfg<BLOCK_N>;
    auto kern = wgrad_tc_kernel<BLOCK_N>;
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES));
        attr_set = true;
    }
    int total = p.out_tiles * p.splits;
    int grid = total < num_sms() ? total : num_sms();
    kern<<<grid, WG_THREADS, Cfg::SMEM_BYTES, stream>>>(tz, ts, p);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200_OK;
}

static int chunk_width(int C) {
    if (C % 64 == 0) return 64;
    if (C % 32 == 0) return 32;
    if (C % 16 == 0) return 16;
    return 0;
}

// dw must be zero-initialised by the caller when more than one reduction split is used; the
// function reports the split count through *splits_out so the caller can tell (it is >1 only
// for layers with few output tiles).  To keep the contract simple the caller always zeroes.
int launch_wgrad_tc(const void* dz, int Nz, const void* src, int Csrc, int T, int B, int H, int W,
                    int ksize, float* dw, long long ldk, int koff, cudaStream_t stream) {
    WgradParams p = {};
    p.T = T; p.B = B; p.H = H; p.W = W;
    p.Nz = Nz; p.Csrc = Csrc; p.ksize = ksize; p.pad = ksize / 2;
    if (chunk_width(Nz) == 0 || chunk_width(Csrc) == 0) {
        set_last_error("wgrad_tc: channels Nz=%d Csrc=%d not multiples of 16", Nz, Csrc);
        return B200_ERR_SHAPE;
    }
    if ((ldk % 4) != 0 || (koff % 4) != 0) {
        set_last_error("wgrad_tc: ldk/koff must be multiples of 4");
        return B200_ERR_ALIGN;
    }
    MTile mt;
    if (!plan_mtile(B, H, W, WG_RB, &mt)) {
        set_last_error("wgrad_tc: spatial shape B=%d H=%d W=%d cannot be tiled", B, H, W);
        return B200_ERR_SHAPE;
    }
    p.Wt = mt.Wt; p.Ht = mt.Ht; p.Bt = mt.Bt;
    p.tiles_w = mt.tiles_w; p.tiles_h = mt.tiles_h; p.tiles_b = mt.tiles_b;
    p.num_rblocks = T * mt.tiles_w * mt.tiles_h * mt.tiles_b;
    const int taps = ksize * ksize;
    // few source channels: stack G taps of the source along M, dz becomes the N operand
    p.stacked = (Csrc == 16 || Csrc == 32 || Csrc == 64) && taps > 1 ? 1 : 0;
    int block_n;
    if (p.stacked) {
        p.G = WG_BLOCK_M / Csrc;
        p.cwA = Csrc;
        p.cwB = chunk_width(Nz);
        block_n = Nz > 128 ? 256 : (Nz > 64 ? 128 : 64);
        p.num_m_tiles = 1;
        p.num_n_tiles = (Nz + block_n - 1) / block_n;
        p.out_tiles = ((taps + p.G - 1) / p.G) * p.num_n_tiles;
    } else {
        p.G = 1;
        p.cwA = chunk_width(Nz);
        p.cwB = chunk_width(Csrc);
        block_n = Csrc > 128 ? 256 : (Csrc > 64 ? 128 : 64);
        p.num_m_tiles = (Nz + WG_BLOCK_M - 1) / WG_BLOCK_M;
        p.num_n_tiles = (Csrc + block_n - 1) / block_n;
        p.out_tiles = taps * p.num_m_tiles * p.num_n_tiles;
    }
    // reduction split: fill two waves of CTAs without spilling into a third (floor, not ceil), and
    // keep at least 8 reduction blocks per unit
    int splits = (2 * num_sms()) / p.out_tiles;
    int max_splits = (p.num_rblocks + 7) / 8;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.rb_per_split = (p.num_rblocks + splits - 1) / splits;
    p.splits = (p.num_rblocks + p.rb_per_split - 1) / p.rb_per_split;
    p.dw = dw; p.ldk = ldk; p.koff = koff;
    p.err_flag = device_error_flag();

    CUtensorMap tz, ts;
    int rc = make_act_tmap(&tz, dz, Nz, W, H, B, T, p.stacked ? p.cwB : p.cwA, mt.Wt, mt.Ht, mt.Bt);
    if (rc != B200_OK) return rc;
    rc = make_act_tmap(&ts, src, Csrc, W, H, B, T, p.stacked ? p.cwA : p.cwB, mt.Wt, mt.Ht, mt.Bt);
    if (rc != B200_OK) return rc;
    switch (block_n) {
        case 256: return launch_wgrad_impl<256>(tz, ts, p, stream);
        case 128: return launch_wgrad_impl<128>(tz, ts, p, stream);
        default: return launch_wgrad_impl<64>(tz, ts, p, stream);
    }
}

}  // namespace b200
