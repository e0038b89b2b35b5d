// Backward of GeM pooling over the feature map: the only part of the tail's backward that touches
// N*C*H*W elements (everything else is [N, C] / [N, D] sized and stays in stock torch ops + cuBLAS).
//
//   forward (cirtorch/modules/pools.py:37-38):   g = ( mean_hw t^p )^(1/p),  t = max(x, eps)
//   dx[n,c,h,w] = dg[n,c] * g^(1-p) * t^(p-1) / HW      where x >= eps (the clamp passes no gradient below)
//   S[n,c]      = sum_hw t^p * ln t                      (needed for dL/dp; optional)
//
// One pass: reads x once, writes dx once (2 * N*C*H*W*4 bytes, HBM-bound); one warp per (n, c) row,
// 128-bit streaming loads / stores, integer p in {1, 2, 3, 4} without transcendentals.
#include "common.cuh"

namespace cir {

template <int PI>   // PI = integer exponent 1..4, 0 = general
__device__ __forceinline__ void gem_bwd_elem(float x, float eps, float p, float coef, bool want_s, float& dx, float& s) {
    const float t = fmaxf(x, eps);
    float tpm1;     // t^(p-1)
    if (PI == 0) {
        // general exponent: ONE lg2 serves both t^(p-1) = ex2((p-1) lg2 t) and ln t = lg2 t * ln 2 (two MUFU per element; the
        // libm-style exp2f / logf pair cost ~3x the instructions and made the general backward 0.13 ms slower than p = 3)
        float l, e;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(t));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((p - 1.0f) * l));
        tpm1 = e;
        dx = x >= eps ? coef * tpm1 : 0.0f;
        if (want_s) s = fmaf(tpm1 * t, l * 0.6931471805599453f, s);
        return;
    }
    if (PI == 1) tpm1 = 1.0f;
    else if (PI == 2) tpm1 = t;
    else if (PI == 3) tpm1 = t * t;
    else tpm1 = t * t * t;
    dx = x >= eps ? coef * tpm1 : 0.0f;
    if (want_s) s = fmaf(tpm1 * t, __logf(t), s);
}

template <int PI>
__device__ __forceinline__ float gem_bwd_row(const float* __restrict__ x, float* __restrict__ dx, int HW, bool vec, int lane,
                                             float eps, float p, float coef, bool want_s) {
    float s = 0.0f;
    if (vec) {
        const float4* xv = reinterpret_cast<const float4*>(x);
        float4* dv = reinterpret_cast<float4*>(dx);
        const int nvec = HW >> 2;
        int i = lane;
        for (; i + 32 * 3 < nvec; i += 128) {
            float4 u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) u[j] = ld_stream_f4(xv + i + 32 * j);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 d;
                gem_bwd_elem<PI>(u[j].x, eps, p, coef, want_s, d.x, s);
                gem_bwd_elem<PI>(u[j].y, eps, p, coef, want_s, d.y, s);
                gem_bwd_elem<PI>(u[j].z, eps, p, coef, want_s, d.z, s);
                gem_bwd_elem<PI>(u[j].w, eps, p, coef, want_s, d.w, s);
                if (dx) __stcs(dv + i + 32 * j, d);
            }
        }
        for (; i < nvec; i += 32) {
            const float4 u = ld_stream_f4(xv + i);
            float4 d;
            gem_bwd_elem<PI>(u.x, eps, p, coef, want_s, d.x, s);
            gem_bwd_elem<PI>(u.y, eps, p, coef, want_s, d.y, s);
            gem_bwd_elem<PI>(u.z, eps, p, coef, want_s, d.z, s);
            gem_bwd_elem<PI>(u.w, eps, p, coef, want_s, d.w, s);
            if (dx) __stcs(dv + i, d);
        }
    } else {
        for (int i = lane; i < HW; i += 32) {
            float d;
            gem_bwd_elem<PI>(ld_stream_f1(x + i), eps, p, coef, want_s, d, s);
            if (dx) dx[i] = d;
        }
    }
    return s;
}

__global__ void __launch_bounds__(256)
gem_bwd_kernel(const float* __restrict__ x, long long rows, int C, int HW, const float* __restrict__ p, int p_stride, float eps,
               const float* __restrict__ g, const float* __restrict__ dg, float* __restrict__ dx, float* __restrict__ S, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * 8;
    for (long long r = w0; r < rows; r += nw) {
        const int c = (int)(r % C);
        const float pr = __ldg(p + c * p_stride);
        const float gr = __ldg(g + r), dgr = __ldg(dg + r);
        // dg * g^(1-p) / HW
        const float coef = dgr * powf(gr, 1.0f - pr) / (float)HW;
        const float* xr = x + r * HW;
        float* dr = dx ? dx + r * HW : nullptr;
        const bool want_s = S != nullptr;
        float s;
        if (pr == 3.0f) s = gem_bwd_row<3>(xr, dr, HW, vec_ok, lane, eps, pr, coef, want_s);
        else if (pr == 2.0f) s = gem_bwd_row<2>(xr, dr, HW, vec_ok, lane, eps, pr, coef, want_s);
        else if (pr == 1.0f) s = gem_bwd_row<1>(xr, dr, HW, vec_ok, lane, eps, pr, coef, want_s);
        else if (pr == 4.0f) s = gem_bwd_row<4>(xr, dr, HW, vec_ok, lane, eps, pr, coef, want_s);
        else s = gem_bwd_row<0>(xr, dr, HW, vec_ok, lane, eps, pr, coef, want_s);
        if (want_s) {
            s = warp_sum(s);
            if (lane == 0) S[r] = s;
        }
    }
}

}  // namespace cir

using namespace cir;

extern "C" int cir_gem_bwd(const float* x, int N, int C, int H, int W, const float* p, int p_stride, float eps_gem,
                           const float* g, const float* dg, float* dx, float* S, void* stream) {
    CIR_REQUIRE(x && p && g && dg && (dx || S) && N > 0 && C > 0 && H > 0 && W > 0, CIR_ERR_INVALID_ARG, "cir_gem_bwd: bad arguments");
    CIR_REQUIRE(p_stride == 0 || p_stride == 1, CIR_ERR_INVALID_ARG, "cir_gem_bwd: p_stride must be 0 or 1");
    const int HW = H * W;
    const int vec_ok = ((HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15) == 0);   // dx may be NULL (only S wanted)
    const long long rows = (long long)N * C;
    const DeviceInfo& dev = device_info();
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)dev.num_sms * 8;      // 8 blocks x 256 threads per SM
    if (blocks > cap) blocks = cap;
    gem_bwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, C, HW, p, p_stride, eps_gem, g, dg, dx, S,
                                                                                 vec_ok);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}
