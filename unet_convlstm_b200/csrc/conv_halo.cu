// 3x3 implicit-GEMM convolution on tcgen05 with the activation tile loaded ONCE per output tile
// ("halo" variant of conv_tc.cu, used for the narrow layers -- N <= 128 output channels -- where the
// generic kernel is bound by the L2 -> shared-memory fill rate because every tap re-loads its A box).
//
// Idea.  Pad every image row with a zero column on each side: pitch P = W + 2.  In that padded,
// flattened position space the input row needed by output position q for tap (ky, kx) is simply
//     q + ky * P + kx
// i.e. a UNIFORM shift of the whole 128-position M tile.  One TMA box {64 channels, P, R rows} that
// starts at (w = -1, h = first_row - 1) delivers exactly that padded layout into shared memory -- the
// zero halo columns / rows are the TMA out-of-bounds fill -- and the nine taps are nine UMMA A
// descriptors whose start address is offset by (ky * P + kx) * 128 bytes.  A K-major 128B-swizzled
// operand may start at any 128-byte row: the swizzle is a function of the absolute shared-memory
// address (measured on B200 with tools/exp_desc_offset.cu, base-offset field = 0).
// Output positions that fall on a halo column (2 of every P) are computed and discarded.
//
// Activation traffic per 128 outputs drops from 9 x 16 KB to one ~42 KB box (W = 64); when the packed
// weights fit (9 * K * N * 2 bytes next to >= 2 activation stages) they stay resident in shared memory
// for the lifetime of the persistent CTA, otherwise they stream through their own ring.
//
// Warp roles as in conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (EPI_STORE semantics: bias / ReLU, bf16 or fp32, split over two destinations).
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace b200 {

static constexpr int H_BLOCK_M = 128;
static constexpr int H_THREADS = 384;  // 4 control warps + 8 epilogue warps
static constexpr int H_MAX_SA = 4;
static constexpr int H_MAX_SB = 8;

struct ConvHaloParams {
    int IMG, H, W;       // IMG = T*B images
    int C0, C1, N;
    int P;               // padded pitch W + 2
    int RB;              // padded rows per activation box
    int tiles_per_img;   // ceil(H * P / 128)
    int num_m_tiles, num_n_tiles;
    int kc;              // channels per K chunk: 64 (128-byte rows) or 16 (the 2-channel first layer padded to 16)
    int chunks0, chunks; // K chunks of source 0 / both sources
    int SA, SB;          // activation / weight ring depths (SB unused when wres)
    int wres;            // weights resident in smem
    uint32_t a_stage_bytes, a_box_bytes, b_box_bytes;
    void* dst0;
    void* dst1;
    long long ld0, ld1;
    int split, out_fp32, relu, accumulate;
    const float* bias;
    const float* scale;  // per-channel multiplier applied before the bias (folded eval-mode BatchNorm) or nullptr
    // fused BatchNorm statistics (see ConvTcParams::stat_sum): [IMG / imgs_per_t][N]
    double* stat_sum;
    double* stat_sumsq;
    int imgs_per_t;
    int* err_flag;
};

__device__ __forceinline__ uint32_t h_pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int BLOCK_N>
__global__ void __launch_bounds__(H_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                 const __grid_constant__ CUtensorMap tm_b, const ConvHaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + p.SA * p.a_stage_bytes;
    const int b_slots = p.wres ? 9 * p.chunks : p.SB;
    const uint32_t bar_base = b_base + b_slots * p.b_box_bytes;
    auto afull = [&](int s) { return bar_base + 8u * s; };
    auto aempty = [&](int s) { return bar_base + 8u * (H_MAX_SA + s); };
    auto bfull = [&](int s) { return bar_base + 8u * (2 * H_MAX_SA + s); };
    auto bempty = [&](int s) { return bar_base + 8u * (2 * H_MAX_SA + H_MAX_SB + s); };
    const uint32_t wfull = bar_base + 8u * (2 * H_MAX_SA + 2 * H_MAX_SB);
    auto tfull = [&](int s) { return wfull + 8u * (1 + s); };
    auto tempty = [&](int s) { return wfull + 8u * (3 + s); };
    const uint32_t tmem_ptr_addr = wfull + 8u * 5;
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    // [8 epilogue warps][2][BLOCK_N / 2] floats after the 256-byte barrier area
    float* stat_smem = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches use the uniform datapath
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.num_m_tiles * p.num_n_tiles;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a0);
        if (p.C1 > 0) prefetch_tmap(&tm_a1);
        prefetch_tmap(&tm_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < H_MAX_SA; ++s) {
            mbar_init(afull(s), 1);
            mbar_init(aempty(s), 1);
        }
        for (int s = 0; s < H_MAX_SB; ++s) {
            mbar_init(bfull(s), 1);
            mbar_init(bempty(s), 1);
        }
        mbar_init(wfull, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull(s), 1);
            mbar_init(tempty(s), 8);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_addr, 2 * BLOCK_N);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    // tile -> (n tile, image, first position q0, first padded row r0)
    auto decode = [&](int tile, int& n0, int& img, int& q0, int& r0) {
        const int nt = tile / p.num_m_tiles;
        const int m = tile - nt * p.num_m_tiles;
        n0 = nt * BLOCK_N;
        img = m / p.tiles_per_img;
        q0 = (m - img * p.tiles_per_img) * H_BLOCK_M;
        r0 = q0 / p.P;
    };

    if (warp == 0) {
        // =================================== TMA producer ===================================
        // The whole warp runs the loop converged and ONE elected lane issues (elect.sync): ptxas then
        // emits straight-line UTMALDG code instead of the per-instruction ELECT/branch loop it needs
        // under a divergent `lane == 0` guard.
        if (p.wres) {
            // resident weights: all (tap, chunk) boxes of N tile 0, once
            if (elect_one()) {
                mbar_arrive_expect_tx(wfull, 9u * p.chunks * p.b_box_bytes);
                uint32_t dst = b_base;
                for (int tap = 0; tap < 9; ++tap)
                    for (int c = 0; c < p.chunks; ++c, dst += p.b_box_bytes) tma_load_3d(dst, &tm_b, wfull, c * p.kc, 0, tap);
            }
            __syncwarp();
        }
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int n0, img, q0, r0;
            decode(tile, n0, img, q0, r0);
            for (int c = 0; c < p.chunks; ++c) {
                mbar_wait(aempty(sa), pa ^ 1u, p.err_flag, 3000 + 100 + sa);
                if (elect_one()) {
                    mbar_arrive_expect_tx(afull(sa), p.a_box_bytes);
                    const uint32_t dst = a_base + sa * p.a_stage_bytes;
                    if (c < p.chunks0)
                        tma_load_5d(dst, &tm_a0, afull(sa), c * p.kc, -1, r0 - 1, img, 0);
                    else
                        tma_load_5d(dst, &tm_a1, afull(sa), (c - p.chunks0) * p.kc, -1, r0 - 1, img, 0);
                }
                __syncwarp();
                if (++sa == p.SA) {
                    sa = 0;
                    pa ^= 1u;
                }
                if (!p.wres) {
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(bempty(sb), pb ^ 1u, p.err_flag, 3000 + 120 + sb);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(bfull(sb), p.b_box_bytes);
                            tma_load_3d(b_base + sb * p.b_box_bytes, &tm_b, bfull(sb), c * p.kc, n0, tap);
                        }
                        __syncwarp();
                        if (++sb == p.SB) {
                            sb = 0;
                            pb ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =================================== MMA issuer =====================================
        // converged warp, one elected lane issues (see the producer): back-to-back UTCHMMA
        const uint32_t idesc = make_idesc_bf16(H_BLOCK_M, BLOCK_N, 0, 0);
        // K-major operand rows of kc * 2 bytes: 128 B rows / 128B swizzle (kc = 64) or 32 B rows / 32B swizzle (kc = 16)
        const uint32_t row_bytes = p.kc * 2;
        const uint64_t desc_hi = make_smem_desc(0, 16, 8u * row_bytes, p.kc == 64 ? 2u : 6u);
        const int mma_per_tap = p.kc / 16;
        const uint32_t a_lo0 = (a_base & 0x3FFFFu) >> 4, b_lo0 = (b_base & 0x3FFFFu) >> 4;
        const uint32_t a_stage_lo = p.a_stage_bytes >> 4, b_box_lo = p.b_box_bytes >> 4;
        const uint32_t row_lo = row_bytes >> 4;  // one padded position = one operand row
        if (p.wres) {
            mbar_wait(wfull, 0, p.err_flag, 3000 + 250);
            tc_fence_after();
        }
        int sa = 0, sb = 0, acc = 0;
        uint32_t pa = 0, pb = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int n0, img, q0, r0;
            decode(tile, n0, img, q0, r0);
            const uint32_t off = q0 - r0 * p.P;  // first output position inside the loaded box
            mbar_wait(tempty(acc), acc_phase ^ 1u, p.err_flag, 3000 + 300 + acc);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            uint32_t accum = 0;
            for (int c = 0; c < p.chunks; ++c) {
                mbar_wait(afull(sa), pa, p.err_flag, 3000 + 200 + sa);
                tc_fence_after();
                const uint32_t a_row0 = a_lo0 + sa * a_stage_lo + off * row_lo;  // tap (0, 0)
                if (p.wres) {
                    if (elect_one()) {
                        uint32_t a_row = a_row0;
                        uint32_t b_lo = b_lo0 + c * b_box_lo;
                        const uint32_t b_tap_lo = p.chunks * b_box_lo;
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky, a_row += (p.P - 3) * row_lo) {
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx, a_row += row_lo, b_lo += b_tap_lo) {
                                const uint64_t adesc = desc_hi | a_row;
                                const uint64_t bdesc = desc_hi | b_lo;
                                umma_bf16(d_tmem, adesc, bdesc, idesc, accum);
                                if (mma_per_tap == 4) {
                                    umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                                    umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                                    umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                                }
                                accum = 1u;
                            }
                        }
                        umma_commit(aempty(sa));  // every tap of this chunk has read the activation box
                    }
                    __syncwarp();
                    accum = 1u;
                } else {
                    uint32_t a_row = a_row0;
                    for (int ky = 0; ky < 3; ++ky, a_row += (p.P - 3) * row_lo) {
                        for (int kx = 0; kx < 3; ++kx, a_row += row_lo) {
                            mbar_wait(bfull(sb), pb, p.err_flag, 3000 + 220 + sb);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t adesc = desc_hi | a_row;
                                const uint64_t bdesc = desc_hi | (b_lo0 + sb * b_box_lo);
                                umma_bf16(d_tmem, adesc, bdesc, idesc, accum);
                                if (mma_per_tap == 4) {
                                    umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                                    umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                                    umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                                }
                                umma_commit(bempty(sb));
                            }
                            __syncwarp();
                            accum = 1u;
                            if (++sb == p.SB) {
                                sb = 0;
                                pb ^= 1u;
                            }
                        }
                    }
                    if (elect_one()) umma_commit(aempty(sa));
                    __syncwarp();
                }
                if (++sa == p.SA) {
                    sa = 0;
                    pa ^= 1u;
                }
            }
            if (elect_one()) umma_commit(tfull(acc));
            __syncwarp();
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // =================================== epilogue =======================================
        // 8 warps: a TMEM lane quarter (warp % 4) is shared by two warps that split the columns.  With
        // a single warp per scheduler the TMEM-load / bias-load / store latencies were fully exposed and
        // the N = 64 layers were epilogue-bound (tensor pipe 28 % busy, profiles/r01_ncu_conv_halo_*).
        const int e = static_cast<int>(threadIdx.x >> 5) - 4;  // vector-register copy: keeps the epilogue arithmetic off the uniform datapath
        const int qw = e & 3;
        const int half = e >> 2;
        const int r = qw * 32 + lane;
        constexpr int COLS = BLOCK_N / 2;  // columns handled by this warp
        // 256-bit stores (one transaction per row and 16-channel piece) when every row piece is 32-byte aligned
        const bool wide = !p.out_fp32 && (p.ld0 % 16 == 0) && (p.dst1 == nullptr || p.ld1 % 16 == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.dst0) | reinterpret_cast<uintptr_t>(p.dst1)) & 31) == 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        // fused BatchNorm statistics: per-warp partial sums in shared memory (plain read-modify-write by
        // the owning lane), flushed with fp64 atomics when the (t, N tile) key changes (see conv_tc.cu)
        const bool do_stats = p.stat_sum != nullptr;
        float* stat_row = stat_smem + e * (2 * COLS);
        int stat_t = -1, stat_n0 = 0;
        auto stat_flush = [&]() {
            __syncwarp();
            for (int i = lane; i < 2 * COLS; i += 32) {
                const float v = stat_row[i];
                stat_row[i] = 0.f;
                const int col = stat_n0 + half * COLS + (i % COLS);
                if (col < p.N && stat_t >= 0)
                    atomicAdd((i < COLS ? p.stat_sum : p.stat_sumsq) + static_cast<long long>(stat_t) * p.N + col,
                              static_cast<double>(v));
            }
            __syncwarp();
        };
        if (do_stats) stat_flush();  // zeroes the buffer
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int n0, img, q0, r0;
            decode(tile, n0, img, q0, r0);
            if (do_stats) {
                const int tt = img / p.imgs_per_t;
                if (tt != stat_t || n0 != stat_n0) {
                    if (stat_t >= 0) stat_flush();
                    stat_t = tt;
                    stat_n0 = n0;
                }
            }
            const int q = q0 + r;
            const int h = q / p.P;
            const int w = q - h * p.P;
            const bool valid = (w < p.W) && (h < p.H);  // halo columns / rows past the image are discarded
            const long long pix = (static_cast<long long>(img) * p.H + h) * p.W + w;
            mbar_wait(tfull(acc), acc_phase, p.err_flag, 3000 + 400 + acc);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BLOCK_N + half * COLS + (uint32_t(qw * 32) << 16);
#pragma unroll 1
            for (int c32 = 0; c32 < COLS / 32; ++c32) {
                const int ncol = n0 + half * COLS + c32 * 32;
                if (ncol >= p.N) break;  // warp-uniform
                uint32_t v[32];
                tmem_ld32(t_row + c32 * 32, v);
                tmem_ld_wait();
                if (do_stats) {
#pragma unroll
                    for (int g16 = 0; g16 < 2; ++g16) {
                        if (ncol + g16 * 16 >= p.N) break;
                        float s1[16], s2[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float x = __uint_as_float(v[g16 * 16 + j]);
                            if (p.bias) x += __ldg(p.bias + ncol + g16 * 16 + j);
                            x = valid ? __bfloat162float(__float2bfloat16_rn(x)) : 0.f;
                            s1[j] = x;
                            s2[j] = x * x;
                        }
                        warp_colsum16(s1, lane);
                        warp_colsum16(s2, lane);
                        if ((lane & 1) == 0) {
                            const int col = c32 * 32 + g16 * 16 + stat_col(lane);
                            stat_row[col] += s1[0];
                            stat_row[COLS + col] += s2[0];
                        }
                    }
                }
                if (!valid) continue;
                const int nv = min(32, p.N - ncol);  // multiple of 16
#pragma unroll
                for (int g16 = 0; g16 < 2; ++g16) {
                    if (g16 * 16 >= nv) break;
                    const int nc = ncol + g16 * 16;
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[g16 * 16 + j]);
                    if (p.scale) {
                        const float4* s4 = reinterpret_cast<const float4*>(p.scale + nc);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 ss = __ldg(s4 + j);
                            f[4 * j] *= ss.x; f[4 * j + 1] *= ss.y; f[4 * j + 2] *= ss.z; f[4 * j + 3] *= ss.w;
                        }
                    }
                    if (p.bias) {
                        const float4* b4 = reinterpret_cast<const float4*>(p.bias + nc);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 bb = __ldg(b4 + j);
                            f[4 * j] += bb.x; f[4 * j + 1] += bb.y; f[4 * j + 2] += bb.z; f[4 * j + 3] += bb.w;
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    const bool second = nc >= p.split;
                    const long long o = second ? pix * p.ld1 + (nc - p.split) : pix * p.ld0 + nc;
                    void* base = second ? p.dst1 : p.dst0;
                    if (p.out_fp32) {
                        float4* d4 = reinterpret_cast<float4*>(static_cast<float*>(base) + o);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float4 val = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                            if (p.accumulate) {
                                const float4 old = d4[j];
                                val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
                            }
                            d4[j] = val;
                        }
                    } else if (wide) {
                        st_bf16x16(static_cast<__nv_bfloat16*>(base) + o, f);
                    } else {
                        uint4* d4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + o);
                        d4[0] = make_uint4(h_pack_bf16x2(f[0], f[1]), h_pack_bf16x2(f[2], f[3]),
                                           h_pack_bf16x2(f[4], f[5]), h_pack_bf16x2(f[6], f[7]));
                        d4[1] = make_uint4(h_pack_bf16x2(f[8], f[9]), h_pack_bf16x2(f[10], f[11]),
                                           h_pack_bf16x2(f[12], f[13]), h_pack_bf16x2(f[14], f[15]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        if (do_stats && stat_t >= 0) stat_flush();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BLOCK_N);
    }
}

// ------------------------------------------------------------------------------------------------
static constexpr int H_SMEM_LIMIT = 227 * 1024;

// Geometry / shared-memory plan; returns false if the halo kernel cannot (or should not) run this problem.
static bool plan_halo(int IMG, int H, int W, int C0, int C1, int N, int ksize, int block_n, ConvHaloParams* p,
                      int* smem_bytes) {
    if (ksize != 3 || C0 <= 0 || N % 16 != 0 || W + 2 > 256 || W < 4) return false;
    // 64-channel K chunks, or the single 16-channel chunk of the zero-padded first layer
    const int kc = (C0 == 16 && C1 == 0) ? 16 : 64;
    if (C0 % kc != 0 || C1 % kc != 0) return false;
    p->kc = kc;
    p->IMG = IMG; p->H = H; p->W = W; p->C0 = C0; p->C1 = C1; p->N = N;
    p->P = W + 2;
    p->RB = 3 + (129 + p->P - 1) / p->P;
    if (p->RB > 256) return false;
    p->tiles_per_img = (H * p->P + H_BLOCK_M - 1) / H_BLOCK_M;
    p->num_m_tiles = IMG * p->tiles_per_img;
    p->num_n_tiles = (N + block_n - 1) / block_n;
    p->chunks0 = C0 / kc;
    p->chunks = (C0 + C1) / kc;
    p->a_box_bytes = static_cast<uint32_t>(p->RB) * p->P * static_cast<uint32_t>(kc * 2);
    p->a_stage_bytes = (p->a_box_bytes + 1023u) & ~1023u;
    p->b_box_bytes = static_cast<uint32_t>(block_n) * static_cast<uint32_t>(kc * 2);
    const int fixed = 1024 + 256 + 8 * block_n * 4;  // alignment slack + barriers + BatchNorm partial sums
    // resident weights if they leave room for >= 2 activation stages (single N tile only)
    const long long wbytes = 9LL * p->chunks * p->b_box_bytes;
    p->wres = 0;
    if (p->num_n_tiles == 1 && wbytes + 2LL * p->a_stage_bytes + fixed <= H_SMEM_LIMIT) {
        p->wres = 1;
        p->SB = 0;
        long long sa = (H_SMEM_LIMIT - fixed - wbytes) / p->a_stage_bytes;
        p->SA = static_cast<int>(sa > H_MAX_SA ? H_MAX_SA : sa);
        *smem_bytes = static_cast<int>(fixed + wbytes + static_cast<long long>(p->SA) * p->a_stage_bytes);
        return true;
    }
    p->SB = block_n == 256 ? 4 : 6;
    long long sa = (H_SMEM_LIMIT - fixed - static_cast<long long>(p->SB) * p->b_box_bytes) / p->a_stage_bytes;
    if (sa < 2) return false;
    p->SA = static_cast<int>(sa > H_MAX_SA ? H_MAX_SA : sa);
    *smem_bytes = static_cast<int>(fixed + static_cast<long long>(p->SB) * p->b_box_bytes +
                                   static_cast<long long>(p->SA) * p->a_stage_bytes);
    return true;
}

static int halo_block_n(int N) { return N > 128 ? 256 : (N > 64 ? 128 : 64); }

bool conv_halo_supported(int IMG, int H, int W, int C0, int C1, int N, int ksize) {
    ConvHaloParams p = {};
    int smem = 0;
    return plan_halo(IMG, H, W, C0, C1, N, ksize, halo_block_n(N), &p, &smem);
}

template <int BLOCK_N>
static int launch_halo_impl(const CUtensorMap& ta0, const CUtensorMap& ta1, const CUtensorMap& tb,
                            const ConvHaloParams& p, int smem_bytes, cudaStream_t stream) {
    auto kern = conv_halo_kernel<BLOCK_N>;
    static PerDeviceOnce attr_once;  // kernel attributes are per device
    if (attr_once.first()) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM_LIMIT));
    }
    const int total = p.num_m_tiles * p.num_n_tiles;
    const int grid = total < num_sms() ? total : num_sms();
    kern<<<grid, H_THREADS, smem_bytes, stream>>>(ta0, ta1, tb, p);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200_OK;
}

// src0/src1: bf16 [IMG][H][W][C*]; wpacked: bf16 [9][N][C0+C1]; output semantics of EPI_STORE.
int launch_conv_halo(const void* src0, const void* src1, const void* wpacked, int IMG, int H, int W, int C0, int C1,
                     int N, const float* bias, void* dst0, long long ld0, int split, void* dst1, long long ld1,
                     int out_fp32, int relu, int accumulate, double* stat_sum, double* stat_sumsq, int imgs_per_t,
                     const float* scale, cudaStream_t stream) {
    ConvHaloParams p = {};
    int smem = 0;
    const int block_n = halo_block_n(N);
    if (!plan_halo(IMG, H, W, C0, C1, N, 3, block_n, &p, &smem)) {
        set_last_error("conv_halo: shape not supported");
        return B200_ERR_SHAPE;
    }
    p.dst0 = dst0; p.dst1 = dst1; p.ld0 = ld0; p.ld1 = ld1; p.split = split;
    p.out_fp32 = out_fp32; p.relu = relu; p.accumulate = accumulate; p.bias = bias; p.scale = scale;
    p.stat_sum = stat_sum; p.stat_sumsq = stat_sumsq; p.imgs_per_t = imgs_per_t > 0 ? imgs_per_t : 1;
    p.err_flag = device_error_flag();
    CUtensorMap ta0, ta1, tb;
    {
        const uint64_t dims[5] = {uint64_t(C0), uint64_t(W), uint64_t(H), uint64_t(IMG), 1};
        const uint64_t str[4] = {uint64_t(C0), uint64_t(C0) * W, uint64_t(C0) * W * H, uint64_t(C0) * W * H * IMG};
        const uint32_t box[5] = {uint32_t(p.kc), uint32_t(p.P), uint32_t(p.RB), 1u, 1u};
        int rc = make_tmap_5d(&ta0, src0, dims, str, box);
        if (rc != B200_OK) return rc;
    }
    if (C1 > 0) {
        const uint64_t dims[5] = {uint64_t(C1), uint64_t(W), uint64_t(H), uint64_t(IMG), 1};
        const uint64_t str[4] = {uint64_t(C1), uint64_t(C1) * W, uint64_t(C1) * W * H, uint64_t(C1) * W * H * IMG};
        const uint32_t box[5] = {uint32_t(p.kc), uint32_t(p.P), uint32_t(p.RB), 1u, 1u};
        int rc = make_tmap_5d(&ta1, src1, dims, str, box);
        if (rc != B200_OK) return rc;
    } else {
        ta1 = ta0;
    }
    int rc = make_w_tmap(&tb, wpacked, C0 + C1, N, 9, p.kc, block_n);
    if (rc != B200_OK) return rc;
    switch (block_n) {
        case 256: return launch_halo_impl<256>(ta0, ta1, tb, p, smem, stream);
        case 128: return launch_halo_impl<128>(ta0, ta1, tb, p, smem, stream);
        default: return launch_halo_impl<64>(ta0, ta1, tb, p, smem, stream);
    }
}

}  // namespace b200
