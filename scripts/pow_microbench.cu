// Compute-only microbenchmark of the GeM row reduction (sum over a row of max(x, eps)^p) on B200: one 512-thread CTA per
// SM, 15 consumer warps reading rows from shared memory exactly as tail.cu's phase A does (8 float4 per lane per row),
// no HBM traffic.  Prints cycles per 32-element warp step per SM sub-partition for several formulations of x^p, next to
// the budget the HBM stream leaves (64 x 2048 x 1024 elements in ~80 us).  Also raw pipe throughputs (MUFU.EX2, MUFU.LG2,
// FFMA, FMNMX).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/build/pow_microbench scripts/pow_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float lg2f_(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2f_(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float poly_ex2(float t) {
    t = fminf(fmaxf(t, -126.0f), 127.0f);
    const float r = t + 12582912.0f;
    const float f = t - (r - 12582912.0f);
    float q = 0.009570102207362652f;
    q = fmaf(q, f, 0.05591786280274391f);
    q = fmaf(q, f, 0.240247443318367f);
    q = fmaf(q, f, 0.6931217908859253f);
    q = fmaf(q, f, 0.9999992847442627f);
    return __int_as_float(__float_as_int(q) + (__float_as_int(r) << 23));
}
// cheaper polynomial variant: no clamp (caller guarantees range), degree 3
__device__ __forceinline__ float poly_ex2_d3(float t) {
    const float r = t + 12582912.0f;
    const float f = t - (r - 12582912.0f);
    float q = 0.0551716685295105f;
    q = fmaf(q, f, 0.2426111400127411f);
    q = fmaf(q, f, 0.6932609677314758f);
    q = fmaf(q, f, 0.9999280571937561f);
    return __int_as_float(__float_as_int(q) + (__float_as_int(r) << 23));
}

enum { V_P3 = 0, V_MUFU = 1, V_POLY1 = 2, V_POLY2 = 3, V_POLY3 = 4, V_POLY4 = 5, V_MUFU_4ACC = 6, V_P3_TIMES_FRAC = 7, V_MUFU_NOCLAMP = 8,
       V_POLY2_D3 = 9, V_LG2_ONLY = 10, V_EX2_ONLY = 11, V_HALF2 = 12, NVAR = 13 };
static const char* names[NVAR] = {"p=3 (FMNMX FMUL FFMA)", "general: lg2+ex2 on MUFU (r1)", "1 of 4 ex2 polynomial", "2 of 4 ex2 polynomial",
                                  "3 of 4 ex2 polynomial", "4 of 4 ex2 polynomial", "MUFU, 4 accumulators", "x^3 * ex2((p-3) lg2 x)",
                                  "MUFU, clamp after lg2", "2 of 4 polynomial, degree 3, no range clamp", "lg2 only (1 MUFU)",
                                  "ex2 only (1 MUFU)", "ex2 as f16x2 pairs (1.5 MUFU, inexact)"};

template <int V>
__device__ __forceinline__ float fold1(float acc, float v, float eps, float p, int which) {
    const float t = fmaxf(v, eps);
    if (V == V_P3) return fmaf(t * t, t, acc);
    if (V == V_LG2_ONLY) return acc + lg2f_(t);
    if (V == V_EX2_ONLY) return acc + ex2f_(t);
    if (V == V_P3_TIMES_FRAC) return fmaf(t * t * t, ex2f_((p - 3.0f) * lg2f_(t)), acc);
    if (V == V_MUFU_NOCLAMP) return acc + ex2f_(p * fmaxf(lg2f_(v), -19.931568f));
    const float e = p * lg2f_(t);
    bool poly = false;
    if (V == V_POLY1) poly = which == 1;
    if (V == V_POLY2 || V == V_POLY2_D3) poly = which & 1;
    if (V == V_POLY3) poly = which != 0;
    if (V == V_POLY4) poly = true;
    if (V == V_POLY2_D3) return acc + (poly ? poly_ex2_d3(e) : ex2f_(e));
    return acc + (poly ? poly_ex2(e) : ex2f_(e));
}

template <int V>
__device__ __forceinline__ float row_partial(const float4* v, int lane, float eps, float p) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    float4 u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = v[lane + 32 * j];
    if (V == V_HALF2) {
        // lg2 in fp32 (MUFU), the two products packed to half2, ONE MUFU.EX2.F16x2 for both (accuracy ~1e-3: throughput probe only)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float e0 = p * lg2f_(fmaxf(u[j].x, eps)), e1 = p * lg2f_(fmaxf(u[j].y, eps));
            const float e2 = p * lg2f_(fmaxf(u[j].z, eps)), e3 = p * lg2f_(fmaxf(u[j].w, eps));
            unsigned h01, h23, r01, r23;
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h01) : "f"(e1), "f"(e0));
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h23) : "f"(e3), "f"(e2));
            asm("ex2.approx.f16x2 %0, %1;" : "=r"(r01) : "r"(h01));
            asm("ex2.approx.f16x2 %0, %1;" : "=r"(r23) : "r"(h23));
            float f0, f1, f2, f3;
            asm("{.reg .f16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(f0), "=f"(f1) : "r"(r01));
            asm("{.reg .f16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(f2), "=f"(f3) : "r"(r23));
            a0 += f0; a1 += f1; a0 += f2; a1 += f3;
        }
        return a0 + a1;
    }
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        if (V == V_MUFU_4ACC) {
            a0 = fold1<V_MUFU>(a0, u[j].x, eps, p, 0); a1 = fold1<V_MUFU>(a1, u[j].y, eps, p, 1);
            a2 = fold1<V_MUFU>(a2, u[j].z, eps, p, 2); a3 = fold1<V_MUFU>(a3, u[j].w, eps, p, 3);
            a0 = fold1<V_MUFU>(a0, u[j + 1].x, eps, p, 0); a1 = fold1<V_MUFU>(a1, u[j + 1].y, eps, p, 1);
            a2 = fold1<V_MUFU>(a2, u[j + 1].z, eps, p, 2); a3 = fold1<V_MUFU>(a3, u[j + 1].w, eps, p, 3);
        } else {
            a0 = fold1<V>(a0, u[j].x, eps, p, 0); a0 = fold1<V>(a0, u[j].y, eps, p, 1);
            a0 = fold1<V>(a0, u[j].z, eps, p, 2); a0 = fold1<V>(a0, u[j].w, eps, p, 3);
            a1 = fold1<V>(a1, u[j + 1].x, eps, p, 0); a1 = fold1<V>(a1, u[j + 1].y, eps, p, 1);
            a1 = fold1<V>(a1, u[j + 1].z, eps, p, 2); a1 = fold1<V>(a1, u[j + 1].w, eps, p, 3);
        }
    }
    return (a0 + a1) + (a2 + a3);
}

// 512 threads: warps 0..nw-1 each reduce `rows` rows of 1024 floats held in shared memory (24 rows = 96 KB, like the ring)
template <int V>
__global__ void __launch_bounds__(512, 1) pow_kernel(const float* __restrict__ src, float* __restrict__ out, int rows, int nw, float eps,
                                                       const float* __restrict__ pp, long long* cycles) {
    extern __shared__ float4 ring[];
    for (int i = threadIdx.x; i < 24 * 256; i += blockDim.x) ring[i] = reinterpret_cast<const float4*>(src)[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float p = pp[0];
    float acc = 0.f;
    const long long t0 = clock64();
    if (warp < nw)
        for (int r = 0; r < rows; ++r) {
            float a = row_partial<V>(ring + ((r + warp) % 24) * 256, lane, eps, p);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            acc += a;
        }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (lane == 0 && warp < nw) out[blockIdx.x * 16 + warp] = acc;
}

template <int V>
static void run(const float* src, float* out, const float* pp, long long* cyc, int nw) {
    const int rows = 200;
    cudaFuncSetAttribute(pow_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 4096);
    pow_kernel<V><<<148, 512, 24 * 4096>>>(src, out, rows, nw, 1e-6f, pp, cyc);
    pow_kernel<V><<<148, 512, 24 * 4096>>>(src, out, rows, nw, 1e-6f, pp, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += (double)h[i];
    mean /= 148;
    // warp steps (32 elements each) per SM: nw warps * rows * 32 float4-steps * 4 elements; per sub-partition: / 4
    const double steps_per_smsp = (double)nw * rows * 32.0 / 4.0;
    const double cyc_per_step = mean / steps_per_smsp;
    // budget: 64*2048 rows over 148 SMs in 80 us at 1.965 GHz
    const double budget = 80e-6 * 1.965e9 / ((64.0 * 2048 / 148) * 32.0 / 4.0);
    printf("%-52s warps %2d  %7.2f cycles per 32-element step per sub-partition  (HBM budget %.1f)  -> %5.1f us for 64x2048x32x32 at 1.965 GHz\n",
           names[V], nw, cyc_per_step, budget, cyc_per_step * (64.0 * 2048 / 148) * 8.0 / 1.965e3);
}

int main() {
    float *src, *out, *pp;
    long long* cyc;
    cudaMalloc(&src, 24 * 4096);
    cudaMalloc(&out, 148 * 16 * 4);
    cudaMalloc(&pp, 4);
    cudaMalloc(&cyc, 148 * 8);
    float h[24 * 1024];
    for (int i = 0; i < 24 * 1024; ++i) h[i] = (i % 3 == 0) ? 0.f : 0.01f + 0.001f * (i % 977);
    cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
    const float p = 2.7f;
    cudaMemcpy(pp, &p, 4, cudaMemcpyHostToDevice);
    for (int nw : {15, 16}) {
        run<V_P3>(src, out, pp, cyc, nw);
        run<V_MUFU>(src, out, pp, cyc, nw);
        run<V_POLY1>(src, out, pp, cyc, nw);
        run<V_POLY2>(src, out, pp, cyc, nw);
        run<V_POLY3>(src, out, pp, cyc, nw);
        run<V_POLY4>(src, out, pp, cyc, nw);
        run<V_MUFU_4ACC>(src, out, pp, cyc, nw);
        run<V_P3_TIMES_FRAC>(src, out, pp, cyc, nw);
        run<V_MUFU_NOCLAMP>(src, out, pp, cyc, nw);
        run<V_POLY2_D3>(src, out, pp, cyc, nw);
        run<V_LG2_ONLY>(src, out, pp, cyc, nw);
        run<V_EX2_ONLY>(src, out, pp, cyc, nw);
        run<V_HALF2>(src, out, pp, cyc, nw);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
