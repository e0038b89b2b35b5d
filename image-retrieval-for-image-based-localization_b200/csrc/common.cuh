// Shared host/device helpers for libcir_b200.so (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/cir_b200.h"

namespace cir {

// ---------------------------------------------------------------- host-side plumbing
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

struct DeviceInfo {
    int device = -1;
    int num_sms = 0;
    int max_smem_optin = 0;
    int coop = 0;
};
const DeviceInfo& device_info();   // cached per calling thread's current device

// K-major bf16 matrix [rows, cols] with row pitch `pitch_elems` -> TMA map with box {64 cols, box_rows},
// 128 B swizzle, zero fill out of bounds (cuTensorMapEncodeTiled through the runtime's driver entry point)
int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows);

#define CIR_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            cir::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                           __FILE__, __LINE__);                                           \
            return CIR_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define CIR_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            cir::set_error(__VA_ARGS__);                                                  \
            return (code);                                                                \
        }                                                                                 \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit load: read-only path, do not allocate in L1 (data is touched once)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// order-preserving map fp32 -> u32 (larger float <=> larger unsigned)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// 64-bit sort key: descending key order == (score desc, index asc).  key 0 = "empty".
__device__ __forceinline__ unsigned long long make_key(float s, uint32_t idx) {
    return ((unsigned long long)float_to_ordered(s) << 32) | (unsigned long long)(0xffffffffu - idx);
}
// candidate-list entries as the search epilogue stores them: (index << 32) | raw fp32 score bits
__device__ __forceinline__ unsigned long long raw_to_key(unsigned long long raw) {
    return make_key(__uint_as_float((uint32_t)raw), (uint32_t)(raw >> 32));
}
__device__ __forceinline__ float key_score(unsigned long long k) {
    return ordered_to_float((uint32_t)(k >> 32));
}
__device__ __forceinline__ uint32_t key_index(unsigned long long k) {
    return 0xffffffffu - (uint32_t)(k & 0xffffffffull);
}

__device__ __forceinline__ unsigned long long key_to_raw(unsigned long long k) {
    return ((unsigned long long)key_index(k) << 32) | (unsigned long long)__float_as_uint(key_score(k));
}

#endif  // __CUDACC__

}  // namespace cir
