// Error reporting, launch accounting and cached device attributes for libcir_b200.so.
#include "common.cuh"

namespace cir {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += n; }

const DeviceInfo& device_info() {
    static thread_local DeviceInfo info[16];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) dev = 0;
    DeviceInfo& d = info[dev];
    if (d.device != dev) {
        cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&d.coop, cudaDevAttrCooperativeLaunch, dev);
        d.device = dev;
    }
    return d;
}

typedef CUresult (*EncodeTiledFnT)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows) {
    static EncodeTiledFnT fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnT>(p);
    }
    CIR_REQUIRE(fn, CIR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {64u, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CIR_REQUIRE(r == CUDA_SUCCESS, CIR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CIR_OK;
}

}  // namespace cir

extern "C" {

const char* cir_last_error(void) { return cir::g_err; }

int cir_version(void) { return 210; }   // 200: round 2 (pooled_out; training, radix-sort and loss entry points); 210: + cir_tail_fwd_train, cir_search_topk_exchange_merge

int64_t cir_launch_count(int reset) {
    int64_t v = cir::g_launches;
    if (reset) cir::g_launches = 0;
    return v;
}

}  // extern "C"
