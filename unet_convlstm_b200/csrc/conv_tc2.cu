// Implicit-GEMM convolution on a CTA PAIR (tcgen05 cta_group::2): the 2-SM variant of conv_tc.cu's
// EPI_STORE kernel.
//
// A cluster of two CTAs (two SMs of one TPC) computes one 256 x BLOCK_N tile: CTA r owns the 128 output
// pixels of M tile 2*pair + r (its own A boxes, its own TMEM lanes, its own epilogue) and HALF of the
// weight tile (BLOCK_N / 2 rows).  One thread of the leader CTA issues tcgen05.mma.cta_group::2 with
// M = 256: the tensor cores of both SMs read the A rows of their own shared memory and the B rows of
// both, so every weight byte is fetched from L2 once per pair instead of once per CTA and the per-SM
// shared-memory read traffic per MMA drops from A + B to A + B/2 -- which is what bounds the 1-CTA
// kernel at N <= 128 (A 4 KB + B 4 KB per 64-cycle MMA = the 128 B/clk of one SM's shared memory) and
// what keeps the large-K layers at the L2 -> SM fill limit.
//
// Synchronisation (all mbarriers have the same shared-memory offset in both CTAs):
//   full[s]   lives in the LEADER: 2 arrivals (one arrive.expect_tx per CTA's producer, the peer's through
//             the cluster address space) + the bytes of both CTAs' TMA copies (cp.async.bulk.tensor
//             .cta_group::2 with the barrier address mapped to the leader)
//   empty[s]  one per CTA, signalled in BOTH by the leader's tcgen05.commit.cta_group::2 ... multicast
//   tfull[a]  one per CTA (multicast commit), tempty[a] in the leader: 4 epilogue warps x 2 CTAs arrive
// EPI_LSTM (+ the timestep-persistent mode of conv_tc.cu, ConvTcParams::seq_T): the fused ConvLSTM cell.  The
// gate-interleaved N tile [i f g o] x 64 hidden channels is split [i f] / [g o] between the two CTAs' weight
// loads; after the 2-CTA MMA each CTA holds all four gates of its own 128 pixels in TMEM, so the cell update is
// the same lane-local epilogue.  Steps are separated by the grid-wide counter (cooperative cluster launch).
// Warp roles per CTA as in conv_tc.cu (warp 0 TMA producer, warp 1 MMA issuer -- leader only --, warp 2
// TMEM allocator, warps 4-7 epilogue); every wait is bounded (watchdog -> trap).
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace b200 {

static constexpr int P_BLOCK_M = 128;  // rows per CTA (256 per pair)
// Epilogue warps per CTA: two warps share each TMEM lane quarter and split the columns of the tile -- the low-K
// layers (ConvTranspose GEMMs, K = Cin; the 16/64/128-channel 3x3 layers) spend more time draining a tile than
// accumulating it, and the cell update is latency bound.
static constexpr int P_EPI_WARPS = 8;
static constexpr int P_THREADS = (4 + P_EPI_WARPS) * 32;

template <int BLOCK_N>
struct Tc2Cfg {
    static constexpr int A_BYTES = P_BLOCK_M * 128;
    static constexpr int B_BYTES = (BLOCK_N / 2) * 128;  // this CTA's half of the weight tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BLOCK_N == 256) ? 6 : 8;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ uint32_t p_pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

struct PairTile {
    int n0, t, b0, h0, w0;
    bool in_range;  // false: the odd last M tile has no partner -- loads are out of bounds (zero fill), nothing stored
};

__device__ __forceinline__ PairTile decode_pair_tile(const ConvTcParams& p, int ptile, int num_m_pairs, int rank,
                                                     int block_n) {
    PairTile tc;
    const int nt = ptile / num_m_pairs;
    int m = (ptile - nt * num_m_pairs) * 2 + rank;
    tc.n0 = nt * block_n;
    tc.in_range = m < p.num_m_tiles;
    if (!tc.in_range) m = p.num_m_tiles - 1;  // any valid tile: the data is discarded
    const int wt = m % p.tiles_w;
    m /= p.tiles_w;
    const int ht = m % p.tiles_h;
    m /= p.tiles_h;
    const int bt = m % p.tiles_b;
    tc.t = m / p.tiles_b;
    tc.w0 = wt * p.Wt;
    tc.h0 = ht * p.Ht;
    tc.b0 = bt * p.Bt;
    return tc;
}

__device__ __forceinline__ void grid_wait2(const unsigned* ctr, unsigned target, int* err_flag) {
    const long long t0 = clock64();
    unsigned v;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) break;
        if (clock64() - t0 > 8000000000LL) {
            if (err_flag) *reinterpret_cast<volatile int*>(err_flag) = 2901;
            __threadfence_system();
            asm volatile("trap;");
        }
    } while (true);
}

template <int BLOCK_N, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                const __grid_constant__ CUtensorMap tm_b, const ConvTcParams p) {
    using Cfg = Tc2Cfg<BLOCK_N>;
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

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    const int taps = p.ksize * p.ksize;
    const int chunks0 = p.C0 / p.kc;
    const int chunks = chunks0 + p.C1 / p.kc;
    const int num_kb = taps * chunks;
    const bool seq = p.seq_T > 0;
    const int nsteps = seq ? p.seq_T : 1;
    const int num_m_pairs = (p.num_m_tiles + 1) >> 1;
    const int total_ptiles = num_m_pairs * p.num_n_tiles;
    const int num_clusters = gridDim.x >> 1;
    const int cluster_id = blockIdx.x >> 1;
    const uint32_t a_bytes = P_BLOCK_M * p.kc * 2;
    const uint32_t b_bytes = (BLOCK_N / 2) * p.kc * 2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a0);
        if (p.C1 > 0) prefetch_tmap(&tm_a1);
        prefetch_tmap(&tm_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 2);   // one arrive.expect_tx per CTA of the pair (used in the leader only)
            mbar_init(empty_bar(s), 1);  // the leader's multicast commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);   // multicast commit
            mbar_init(tempty_bar(s), 2 * P_EPI_WARPS);  // the epilogue warps of both CTAs (used in the leader only)
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc2(tmem_ptr_addr, Cfg::TMEM_COLS);
        tmem_relinquish2();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before anybody arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0) {
        // =================================== TMA producer (both CTAs) ===================================
        int stage = 0;
        uint32_t phase = 0;
        uint32_t a_dst = smem_base;
        const uint32_t tx_bytes = a_bytes + b_bytes;
        const uint32_t full0_leader = map_to_cta(full_bar(0), 0);  // the leader's full barriers, cluster address
        for (int step = 0; step < nsteps; ++step) {
        // zero initial state (unet.py:23-25): the h half of K is skipped at t = 0
        const int chunks_t = (seq && step == 0 && !p.seq_have_h0) ? chunks0 : chunks;
        if (seq && step > 0) {
            // h_{t-1} was written by the epilogue warps of every CTA in the previous step
            if (lane == 0) grid_wait2(p.sync_ctr, gridDim.x * step, p.err_flag);
            __syncwarp();
            fence_proxy_async_all();
        }
        for (int pt = cluster_id; pt < total_ptiles; pt += num_clusters) {
            PairTile tc = decode_pair_tile(p, pt, num_m_pairs, rank, BLOCK_N);
            if (seq) tc.t = step;
            const int bn0 = tc.n0 + static_cast<int>(rank) * (BLOCK_N / 2);  // this CTA's half of the weight rows
            int ky = 0, kx = 0;
            for (int tap = 0; tap < taps; ++tap) {
                const int cw = tc.w0 + kx - p.pad, chh = tc.h0 + ky - p.pad;
                int kofs = 0;
                for (int c = 0; c < chunks_t; ++c, kofs += p.kc) {
                    mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 2000 + 100 + stage);
                    if (elect_one()) {
                        const uint32_t fb = full0_leader + 8u * stage;
                        mbar_arrive_expect_tx_cluster(fb, tx_bytes);
                        if (c < chunks0)
                            tma2_load_5d(a_dst, &tm_a0, fb, kofs, cw, chh, tc.b0, tc.t);
                        else
                            tma2_load_5d(a_dst, &tm_a1, fb, kofs - p.C0, cw, chh, tc.b0, tc.t);
                        tma2_load_3d(a_dst + Cfg::A_BYTES, &tm_b, fb, kofs, bn0, tap);
                    }
                    __syncwarp();
                    a_dst += Cfg::STAGE_BYTES;
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                        a_dst = smem_base;
                    }
                }
                if (++kx == p.ksize) {
                    kx = 0;
                    ++ky;
                }
            }
        }
        }
    } else if (warp == 1) {
        // =================================== MMA issuer (leader CTA only) ===============================
        if (leader) {
            const uint32_t idesc = make_idesc_bf16(2 * P_BLOCK_M, BLOCK_N, 0, 0);
            const uint32_t row_bytes = p.kc * 2;
            const uint32_t layout_type = (p.kc == 64) ? 2u : (p.kc == 32 ? 4u : 6u);
            const uint64_t desc_hi = make_smem_desc(0, 16, 8u * row_bytes, layout_type);
            const int mma_per_kb = p.kc / 16;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            uint32_t a_lo = (smem_base & 0x3FFFFu) >> 4;
            const uint32_t a_lo0 = a_lo;
            constexpr uint32_t STAGE_LO = Cfg::STAGE_BYTES >> 4, B_LO = Cfg::A_BYTES >> 4;
            for (int step = 0; step < nsteps; ++step) {
            const int num_kb_t = (seq && step == 0 && !p.seq_have_h0) ? taps * chunks0 : num_kb;
            for (int pt = cluster_id; pt < total_ptiles; pt += num_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err_flag, 2000 + 300 + acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                uint32_t accum = 0;
                for (int kb = 0; kb < num_kb_t; ++kb) {
                    mbar_wait(full_bar(stage), phase, p.err_flag, 2000 + 200 + stage);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t adesc = desc_hi | a_lo;
                        const uint64_t bdesc = desc_hi | (a_lo + B_LO);
                        for (int k = 0; k < mma_per_kb; ++k)
                            umma2_bf16(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc,
                                       k == 0 ? accum : 1u);
                        umma2_commit_both(empty_bar(stage));  // frees the slot in both CTAs
                    }
                    __syncwarp();
                    accum = 1u;
                    a_lo += STAGE_LO;
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                        a_lo = a_lo0;
                    }
                }
                if (elect_one()) umma2_commit_both(tfull_bar(acc));  // both CTAs' epilogues
                __syncwarp();
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
            }
        }
    } else if (warp >= 4) {
        // =================================== epilogue (both CTAs, own TMEM lanes) =======================
        const int ew = static_cast<int>(threadIdx.x >> 5) - 4;
        const int q = ew & 3;        // TMEM lane quarter == warp % 4
        const int chalf = ew >> 2;   // which half of the tile's columns this warp drains
        constexpr int NHALF = P_EPI_WARPS / 4;
        const int r = q * 32 + lane;
        const int wi = r % p.Wt;
        const int hi = (r / p.Wt) % p.Ht;
        const int bi = r / (p.Wt * p.Ht);
        const uint32_t tempty0_leader = map_to_cta(tempty_bar(0), 0);
        // 256-bit stores need 32-byte aligned rows: row strides that are multiples of 16 bf16 and aligned bases
        const bool wide = !p.out_fp32 && (p.ld0 % 16 == 0) && (p.dst1 == nullptr || p.ld1 % 16 == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.dst0) | reinterpret_cast<uintptr_t>(p.dst1)) & 31) == 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int step = 0; step < nsteps; ++step) {
        const bool zero_state = seq && step == 0 && !p.seq_have_h0;
        for (int pt = cluster_id; pt < total_ptiles; pt += num_clusters) {
            PairTile tc = decode_pair_tile(p, pt, num_m_pairs, rank, BLOCK_N);
            if (seq) tc.t = step;
            const bool valid = tc.in_range && (tc.h0 + hi < p.H) && (tc.b0 + bi < p.B);
            const long long pix =
                ((static_cast<long long>(tc.t) * p.B + tc.b0 + bi) * p.H + tc.h0 + hi) * p.W + tc.w0 + wi;
            mbar_wait(tfull_bar(acc), acc_phase, p.err_flag, 2000 + 400 + acc);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BLOCK_N + (uint32_t(q * 32) << 16);
            if constexpr (EPI == EPI_LSTM) {
                // ---- fused LSTM cell update (reference train/unet.py:29-35), as in conv_tc.cu ----
                constexpr int CHT = BLOCK_N / 4;
                const int Ch = p.N >> 2;
                const int ch0 = (tc.n0 >> 2);
                // 16 hidden channels per iteration: with CHT = 16 (BLOCK_N = 64) only the first warp of a quarter works
                constexpr int CH_PER = (CHT / NHALF >= 16) ? CHT / NHALF : CHT;
                const int j_begin = (CHT / NHALF >= 16) ? chalf * CH_PER : (chalf == 0 ? 0 : CHT);
#pragma unroll 1
                for (int j0 = j_begin; j0 < j_begin + CH_PER && j0 < CHT; j0 += 16) {
                    uint32_t vi[16], vf[16], vg[16], vo[16];
                    tmem_ld16(t_row + 0 * CHT + j0, vi);
                    tmem_ld16(t_row + 1 * CHT + j0, vf);
                    tmem_ld16(t_row + 2 * CHT + j0, vg);
                    tmem_ld16(t_row + 3 * CHT + j0, vo);
                    tmem_ld_wait();
                    if (!valid) continue;
                    const int ch = ch0 + j0;
                    // pix already contains the step (tc.t = step); the state slots are indexed by step as well
                    const long long coff = pix * Ch + ch;
                    float cp[16];
                    if (p.c_prev && !zero_state) {
                        const float4* c4 = reinterpret_cast<const float4*>(p.c_prev + coff);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 t4 = c4[j];
                            cp[4 * j] = t4.x; cp[4 * j + 1] = t4.y; cp[4 * j + 2] = t4.z; cp[4 * j + 3] = t4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) cp[j] = 0.f;
                    }
                    float gi[16], gf[16], gg[16], go[16], cn[16], hn[16];
                    const float* bp = p.bias ? p.bias + tc.n0 + j0 : nullptr;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float zi = __uint_as_float(vi[j]);
                        float zf = __uint_as_float(vf[j]);
                        float zg = __uint_as_float(vg[j]);
                        float zo = __uint_as_float(vo[j]);
                        if (bp) {
                            zi += __ldg(bp + j);
                            zf += __ldg(bp + CHT + j);
                            zg += __ldg(bp + 2 * CHT + j);
                            zo += __ldg(bp + 3 * CHT + j);
                        }
                        gi[j] = fast_sigmoid(zi);
                        gf[j] = fast_sigmoid(zf);
                        gg[j] = fast_tanh(zg);
                        go[j] = fast_sigmoid(zo);
                        cn[j] = fmaf(gf[j], cp[j], gi[j] * gg[j]);
                        hn[j] = go[j] * fast_tanh(cn[j]);
                    }
                    // Ch is a multiple of 16 and the state buffers come from the caching allocator (512-byte
                    // aligned): every 16-channel piece is 32-byte aligned -> 256-bit stores
                    auto st16 = [](__nv_bfloat16* dst, const float* sv) { st_bf16x16(dst, sv); };
                    // c_next / h_next are NULL in the gate-recompute pass of BPTT (b200_convlstm_gates_recompute_tc)
                    if (p.c_next) {
                        st_f32x8(p.c_next + coff, cn);
                        st_f32x8(p.c_next + coff + 8, cn + 8);
                    }
                    if (p.h_next) st16(p.h_next + coff, hn);
                    if (p.gates_out) {
                        __nv_bfloat16* gb = p.gates_out + pix * (4LL * Ch) + ch;
                        st16(gb, gi);
                        st16(gb + Ch, gf);
                        st16(gb + 2 * Ch, gg);
                        st16(gb + 3 * Ch, go);
                    }
                }
            } else {
#pragma unroll 1
            for (int c16 = chalf * (BLOCK_N / 16 / NHALF); c16 < (chalf + 1) * (BLOCK_N / 16 / NHALF); ++c16) {
                const int ncol = tc.n0 + c16 * 16;
                if (ncol >= p.N) break;  // warp-uniform
                uint32_t v[16];
                tmem_ld16(t_row + c16 * 16, v);
                tmem_ld_wait();
                if (!valid) continue;
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                if (p.scale) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] *= __ldg(p.scale + ncol + j);
                }
                if (p.bias) {
                    const float* bp = p.bias + (p.shuf_C > 0 ? ncol % p.shuf_C : ncol);
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] += __ldg(bp + j);
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                if (p.shuf_C > 0) {
                    // ConvTranspose 2x2 stride 2 (see conv_tc.cu): the tap block of this column chunk selects the
                    // output pixel of the 2x2 patch; shuf_C is a multiple of 16, every piece is 32-byte aligned
                    const int tap = ncol / p.shuf_C, co = ncol - tap * p.shuf_C;
                    const long long img = static_cast<long long>(tc.t) * p.B + tc.b0 + bi;
                    const long long orow = (img * p.shuf_Hd + 2 * (tc.h0 + hi) + (tap >> 1) + p.shuf_oy) * p.shuf_Wd +
                                           2 * (tc.w0 + wi) + (tap & 1) + p.shuf_ox;
                    st_bf16x16(static_cast<__nv_bfloat16*>(p.dst0) + orow * p.shuf_C + co, f);
                    continue;
                }
                const bool second = ncol >= p.split;
                const long long off = second ? pix * p.ld1 + (ncol - p.split) : pix * p.ld0 + ncol;
                void* base = second ? p.dst1 : p.dst0;
                if (p.out_fp32) {
                    float4* o = reinterpret_cast<float4*>(static_cast<float*>(base) + off);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 val = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        if (p.accumulate) {
                            const float4 old = o[j];
                            val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
                        }
                        o[j] = val;
                    }
                } else if (wide) {
                    st_bf16x16(static_cast<__nv_bfloat16*>(base) + off, f);
                } else {
                    uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + off);
                    o[0] = make_uint4(p_pack_bf16x2(f[0], f[1]), p_pack_bf16x2(f[2], f[3]), p_pack_bf16x2(f[4], f[5]),
                                      p_pack_bf16x2(f[6], f[7]));
                    o[1] = make_uint4(p_pack_bf16x2(f[8], f[9]), p_pack_bf16x2(f[10], f[11]), p_pack_bf16x2(f[12], f[13]),
                                      p_pack_bf16x2(f[14], f[15]));
                }
            }
            }
            // release the accumulator stage: the MMA issuer (leader) waits for the epilogue warps of BOTH CTAs
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty0_leader + 8u * acc);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        if (seq && step + 1 < nsteps) {
            // publish h_t / c_t of this CTA's tiles, then signal the grid-wide step counter (see conv_tc.cu)
            asm volatile("bar.sync 1, %0;" ::"n"(P_EPI_WARPS * 32) : "memory");
            if (threadIdx.x == 4 * 32) {
                if (step > 0) grid_wait2(p.sync_ctr, gridDim.x * step, p.err_flag);
                __threadfence();
                atomicAdd(p.sync_ctr, 1u);
            }
        }
        }
    }

    // ---- teardown: nobody leaves while the partner may still read its shared memory / signal its barriers ----
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int EPI>
static int launch_pair_impl(const CUtensorMap& ta0, const CUtensorMap& ta1, const CUtensorMap& tb, const ConvTcParams& p,
                            cudaStream_t stream) {
    using Cfg = Tc2Cfg<BLOCK_N>;
    auto kern = conv_tc2_kernel<BLOCK_N, EPI>;
    static PerDeviceOnce attr_once;  // kernel attributes are per device
    if (attr_once.first()) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    }
    const int pairs = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    const int max_clusters = num_sms() / 2;
    const int clusters = pairs < max_clusters ? pairs : max_clusters;
    if (p.seq_T > 0) {
        // the steps are separated by a grid-wide counter: every CTA must be resident -> cooperative launch (the
        // cluster shape comes from the kernel's __cluster_dims__)
        B200_CUDA_CHECK(cudaMemsetAsync(p.sync_ctr, 0, sizeof(unsigned), stream));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(P_THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        B200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta0, ta1, tb, p));
        return B200_OK;
    }
    kern<<<2 * clusters, P_THREADS, Cfg::SMEM_BYTES, stream>>>(ta0, ta1, tb, p);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200_OK;
}

bool conv_tc2_supported(const ConvTcParams& p) {
    // plain store and ConvTranspose pixel-shuffle epilogues; the fused-statistics variant stays on the 1-CTA kernel
    return p.stat_sum == nullptr && p.seq_T == 0 && p.N % 16 == 0;
}

// Same contract as launch_conv_tc(.., epi, ..) for EPI_STORE problems conv_tc2_supported() accepts and for
// EPI_LSTM (single step or timestep-persistent).
int launch_conv_tc2(const void* src0, const void* src1, const void* wpacked, ConvTcParams p, int epi, cudaStream_t stream) {
    if (p.C1 > 0 && !src1) {
        set_last_error("conv_tc2: C1 > 0 but src1 is null");
        return B200_ERR_ARG;
    }
    if (p.seq_T > 0) {
        p.sync_ctr = device_sync_counter(stream);
        if (!p.sync_ctr) {
            set_last_error("conv_tc2: no step counter");
            return B200_ERR_CUDA;
        }
    }
    const int Ct = p.wK > 0 ? p.wK : p.C0 + p.C1;
    int kc = 64;
    while (kc >= 16 && ((p.C0 % kc) != 0 || (p.C1 % kc) != 0)) kc >>= 1;
    if (kc < 16 || p.C0 <= 0) {
        set_last_error("conv_tc2: channel counts C0=%d C1=%d are not multiples of 16", p.C0, p.C1);
        return B200_ERR_SHAPE;
    }
    p.kc = kc;
    const int block_n = epi == EPI_LSTM ? pick_block_n(p.N, EPI_LSTM) : (p.N > 128 ? 256 : (p.N > 64 ? 128 : 64));
    if (block_n == 0) {
        set_last_error("conv_tc2: N=%d not supported for the LSTM epilogue", p.N);
        return B200_ERR_SHAPE;
    }
    MTile mt;
    if (!plan_mtile(p.B, p.H, p.W, P_BLOCK_M, &mt)) {
        set_last_error("conv_tc2: spatial shape B=%d H=%d W=%d cannot be tiled", p.B, p.H, p.W);
        return B200_ERR_SHAPE;
    }
    p.Wt = mt.Wt; p.Ht = mt.Ht; p.Bt = mt.Bt;
    p.tiles_w = mt.tiles_w; p.tiles_h = mt.tiles_h; p.tiles_b = mt.tiles_b;
    const int map_T = p.seq_T > 0 ? p.seq_T : p.T;
    if (p.seq_T > 0) p.T = 1;  // tiles of ONE step; the kernel iterates over the steps itself
    p.num_m_tiles = p.T * mt.tiles_w * mt.tiles_h * mt.tiles_b;
    p.num_n_tiles = (p.N + block_n - 1) / block_n;
    p.pad = p.ksize / 2;
    p.err_flag = device_error_flag();
    CUtensorMap ta0, ta1, tb;
    int rc = make_act_tmap(&ta0, src0, p.C0, p.W, p.H, p.B, map_T, kc, mt.Wt, mt.Ht, mt.Bt);
    if (rc != B200_OK) return rc;
    if (p.C1 > 0) {
        rc = make_act_tmap(&ta1, src1, p.C1, p.W, p.H, p.B, map_T, kc, mt.Wt, mt.Ht, mt.Bt);
        if (rc != B200_OK) return rc;
    } else {
        ta1 = ta0;
    }
    // each CTA loads block_n / 2 weight rows; rows beyond N are out-of-bounds reads => zero filled
    rc = make_w_tmap(&tb, wpacked, Ct, p.N, p.ksize * p.ksize, kc, block_n / 2);
    if (rc != B200_OK) return rc;
    if (epi == EPI_LSTM) {
        switch (block_n) {
            case 256: return launch_pair_impl<256, EPI_LSTM>(ta0, ta1, tb, p, stream);
            case 128: return launch_pair_impl<128, EPI_LSTM>(ta0, ta1, tb, p, stream);
            default: return launch_pair_impl<64, EPI_LSTM>(ta0, ta1, tb, p, stream);
        }
    }
    switch (block_n) {
        case 256: return launch_pair_impl<256, EPI_STORE>(ta0, ta1, tb, p, stream);
        case 128: return launch_pair_impl<128, EPI_STORE>(ta0, ta1, tb, p, stream);
        default: return launch_pair_impl<64, EPI_STORE>(ta0, ta1, tb, p, stream);
    }
}

}  // namespace b200
