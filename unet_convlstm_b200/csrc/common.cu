// Host-side utilities: error reporting, device properties, TMA tensor-map encoding.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

namespace b200 {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_last_error; }

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int* device_error_flag() {
    static int* flags[64] = {nullptr};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!flags[dev]) {
        // pinned, device-mapped host memory (UVA: same pointer on both sides): the host can still read
        // the barrier id after a watchdog trap has killed the context
        int* p = nullptr;
        if (cudaHostAlloc(&p, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
        *p = 0;
        flags[dev] = p;
    }
    return flags[dev];
}

unsigned* device_sync_counter(cudaStream_t stream) {
    // one counter per (device, stream): two timestep-persistent kernels on different streams may run side by side
    static std::map<std::pair<int, cudaStream_t>, unsigned*> ctrs;
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    unsigned*& slot = ctrs[std::make_pair(dev, stream)];
    if (!slot) {
        unsigned* p = nullptr;
        if (cudaMalloc(&p, sizeof(unsigned)) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, sizeof(unsigned));
        slot = p;
    }
    return slot;
}

// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault,
                                             &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

static CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
    switch (inner_bytes) {
        case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
        case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
        case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
        default: return CU_TENSOR_MAP_SWIZZLE_NONE;
    }
}

int make_tmap_5d(CUtensorMap* out, const void* base, const uint64_t dims[5],
                 const uint64_t strides_elems[4], const uint32_t box[5]) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point not available");
        return B200_ERR_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
        set_last_error("TMA base pointer %p not 16-byte aligned", base);
        return B200_ERR_ALIGN;
    }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < 5; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i < 4; ++i) {
        gstr[i] = strides_elems[i] * 2;  // bf16
        if (gstr[i] % 16 != 0) {
            set_last_error("TMA stride %d = %llu bytes not a multiple of 16", i,
                           (unsigned long long)gstr[i]);
            return B200_ERR_ALIGN;
        }
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, bx,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(int(box[0]) * 2),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error(
            "cuTensorMapEncodeTiled(5d) failed: %d dims={%llu,%llu,%llu,%llu,%llu} box={%u,%u,%u,%u,%u}",
            int(r), (unsigned long long)gdim[0], (unsigned long long)gdim[1],
            (unsigned long long)gdim[2], (unsigned long long)gdim[3], (unsigned long long)gdim[4],
            bx[0], bx[1], bx[2], bx[3], bx[4]);
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

int make_act_tmap(CUtensorMap* out, const void* base, int C, int W, int H, int B, int T, int box_c,
                  int Wt, int Ht, int Bt) {
    uint64_t dims[5] = {uint64_t(C), uint64_t(W), uint64_t(H), uint64_t(B), uint64_t(T)};
    uint64_t str[4] = {uint64_t(C), uint64_t(C) * W, uint64_t(C) * W * H, uint64_t(C) * W * H * B};
    uint32_t box[5] = {uint32_t(box_c), uint32_t(Wt), uint32_t(Ht), uint32_t(Bt), 1u};
    return make_tmap_5d(out, base, dims, str, box);
}

int make_w_tmap(CUtensorMap* out, const void* base, int K, int rows, int taps, int box_k,
                int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point not available");
        return B200_ERR_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (K % 8) != 0) {
        set_last_error("weight TMA map: base %p / K=%d misaligned", base, K);
        return B200_ERR_ALIGN;
    }
    cuuint64_t gdim[3] = {cuuint64_t(K), cuuint64_t(rows), cuuint64_t(taps)};
    cuuint64_t gstr[2] = {cuuint64_t(K) * 2, cuuint64_t(K) * 2 * rows};
    cuuint32_t bx[3] = {cuuint32_t(box_k), cuuint32_t(box_rows), 1u};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, bx,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_k * 2),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(3d weights) failed: %d K=%d rows=%d taps=%d box={%d,%d}",
                       int(r), K, rows, taps, box_k, box_rows);
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
static int pow2_ceil(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

bool plan_mtile(int B, int H, int W, int rows, MTile* mt) {
    if (B <= 0 || H <= 0 || W <= 0) return false;
    int Wt;
    if (W >= rows) {
        if (W % rows != 0) return false;
        Wt = rows;
    } else {
        if (!is_pow2(W)) return false;
        Wt = W;
    }
    int Ht = rows / Wt;
    int hp = pow2_ceil(H);
    if (Ht > hp) Ht = hp;
    int Bt = rows / (Wt * Ht);
    if (Bt > 256 || Ht > 256 || Wt > 256) return false;
    mt->Wt = Wt;
    mt->Ht = Ht;
    mt->Bt = Bt;
    mt->tiles_w = W / Wt;
    mt->tiles_h = (H + Ht - 1) / Ht;
    mt->tiles_b = (B + Bt - 1) / Bt;
    return true;
}

}  // namespace b200

extern "C" const char* b200_last_error(void) { return b200::last_error(); }
