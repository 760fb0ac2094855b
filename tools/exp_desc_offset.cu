// Experiment: can a K-major 128B-swizzled UMMA A operand start at a row that is NOT a multiple of 8
// (i.e. a start address that is 128-byte but not 1024-byte aligned)?  A [256 rows][64 bf16] tile is
// TMA-loaded (SWIZZLE_128B) at a 1024-aligned smem address; B is a 64x64 identity (K-major, SW128).
// D = A[r .. r+127, :] * I must reproduce rows r..r+127 for every row offset r if the hardware swizzle
// is a function of the absolute smem address (or of the descriptor's base_offset field).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/exp tools/exp_desc_offset.cu -lcuda && /tmp/exp
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../unet_convlstm_b200/csrc/common.cuh"
#include "../unet_convlstm_b200/csrc/ptx.cuh"

using namespace b200;

__global__ void __launch_bounds__(128, 1)
exp_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* out,
           int row_off, int use_base_offset) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_addr = base;               // 256 rows * 128 B = 32 KB
    const uint32_t b_addr = base + 32768;       // 64 rows * 128 B = 8 KB
    const uint32_t bar = base + 32768 + 8192;
    const uint32_t bar2 = bar + 8;
    const uint32_t tptr = bar + 16;
    volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar2, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tptr, 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr_gen;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 32768 + 8192);
        tma_load_3d(a_addr, &tm_a, bar, 0, 0, 0);
        tma_load_3d(a_addr + 16384, &tm_a, bar, 0, 128, 0);
        tma_load_3d(b_addr, &tm_b, bar, 0, 0, 0);
        mbar_wait(bar, 0, nullptr, 1);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
        const uint32_t start = a_addr + row_off * 128;
        uint64_t adesc = make_smem_desc(start, 16, 1024, 2);
        if (use_base_offset) adesc |= uint64_t((start >> 7) & 7) << 49;
        const uint64_t bdesc = make_smem_desc(b_addr, 16, 1024, 2);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
        umma_commit(bar2);
    }
    mbar_wait(bar2, 0, nullptr, 2);
    tc_fence_after();
    const int r = warp * 32 + lane;
    for (int c16 = 0; c16 < 4; ++c16) {
        uint32_t v[16];
        tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c16 * 16, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[r * 64 + c16 * 16 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 64);
    }
}

int main() {
    const int ROWS = 256, K = 64;
    std::vector<__nv_bfloat16> ha(ROWS * K), hb(64 * K);
    for (int r = 0; r < ROWS; ++r)
        for (int k = 0; k < K; ++k) ha[r * K + k] = __float2bfloat16(float((r * 7 + (k / 8) * 3 + k % 8 * 29) % 251));  // small ints: exact
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < K; ++k) hb[n * K + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *da, *db;
    float* dout;
    cudaMalloc(&da, ha.size() * 2);
    cudaMalloc(&db, hb.size() * 2);
    cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ta, tb;
    if (make_w_tmap(&ta, da, K, ROWS, 1, 64, 128) != 0 || make_w_tmap(&tb, db, K, 64, 1, 64, 64) != 0) {
        printf("tmap failed\n");
        return 1;
    }
    cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    std::vector<float> ho(128 * 64);
    for (int ubo = 0; ubo < 2; ++ubo) {
        for (int off : {0, 1, 2, 3, 5, 7, 8, 9, 66, 67, 127}) {
            cudaMemset(dout, 0, 128 * 64 * 4);
            exp_kernel<<<1, 128, 65536>>>(ta, tb, dout, off, ubo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("base_offset=%d off=%d: CUDA error %s\n", ubo, off, cudaGetErrorString(e));
                return 2;
            }
            cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < 128; ++r)
                for (int k = 0; k < 64; ++k) {
                    const float want = __bfloat162float(ha[(r + off) * K + k]);
                    if (ho[r * 64 + k] != want) ++bad;
                }
            printf("base_offset_field=%d row_off=%3d : %s (%d mismatches; out[0][0..2]= %.3f %.3f %.3f, out[1][0]=%.3f)\n",
                   ubo, off, bad ? "MISMATCH" : "ok", bad, ho[0], ho[1], ho[2], ho[64]);
        }
    }
    return 0;
}
