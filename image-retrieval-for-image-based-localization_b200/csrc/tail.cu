// Fused descriptor tail:  pool (GeM / MAC / SPoC) -> L2N -> whitening Linear -> L2N.
//
// Replaces cirtorch/modules/pools.py:37-38, normalizations.py:15-16 and
// heads/global_head.py:52-67 (seven eager PyTorch ops, three full passes over the map)
// with ONE cooperative launch:
//   phase A  every warp streams (n, c) rows of the NCHW map with 128-bit no-allocate
//            loads, two rows (8 KB) in flight per warp, clamp + x^p + shuffle reduce ->
//            pooled[n, c] (N*C floats, stays in L2).  This is the HBM-bound part.
//   barrier  cooperative grid sync
//   phase B  CTA i owns a slice of <= 16 output dims; its W slice (<= 128 KB) was
//            prefetched into shared memory with cp.async while phase A ran.  Each warp
//            takes 4 images, lanes split K, exact fp32 FMA, shuffle reduce.  The first
//            L2N is folded in as a scale of the accumulators.
//   barrier  cooperative grid sync (per-chunk partial sums of squares)
//   phase C  every thread rescales the outputs it wrote: second L2N.
//
// Algorithmic HBM bytes per launch: N*C*H*W*4 (x) + D_out*C*4 (W) + D_out*4 (b) + N*D_out*4 (out).
#include "common.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace cir {

constexpr int TAIL_THREADS = 512;
constexpr int TAIL_WARPS = TAIL_THREADS / 32;
constexpr int TAIL_JMAX = 16;    // output dims per phase-B chunk
constexpr int TAIL_KC = 2048;    // K extent of the W slice staged in shared memory
constexpr int TAIL_IMG = 4;      // images per warp in phase B
constexpr size_t TAIL_STAMP_BYTES = 1024 * 8 * 8;   // debug time stamps: up to 1024 CTAs x 8 slots

struct TailParams {
    const float* x;
    int N, C, HW;
    const float* p;
    int p_stride;
    float eps_gem, eps_l2;
    int pool_mode;
    const float* Wt;
    const float* bias;
    int D_out;
    float* out;
    int out_ld;
    float* pooled;     // [N, pooled_ld]
    int pooled_ld;
    float* partial;    // [n_chunks, N] sums of squares of the un-normalised outputs
    int jch, n_chunks;
    unsigned flags;
    int vec_ok;        // rows are 16 B aligned and HW % 4 == 0
    unsigned long long* stamps;   // optional [gridDim][8] globaltimer stamps (CIR_TAIL_DEBUG_STAMPS)
};

__device__ __forceinline__ void stamp(const TailParams& P, int slot) {
    if (P.stamps && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.stamps[(size_t)blockIdx.x * 8 + slot] = t;
    }
}

__device__ __forceinline__ float fast_lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// exponent classes: small integer p avoids the two MUFU ops per element
enum { PM_GENERAL = 0, PM_1 = 1, PM_2 = 2, PM_3 = 3, PM_4 = 4, PM_MAX = 5, PM_MEAN = 6 };

template <int PM>
__device__ __forceinline__ float term(float v, float eps, float p) {
    if (PM == PM_MAX || PM == PM_MEAN) return v;
    float t = fmaxf(v, eps);
    if (PM == PM_1) return t;
    if (PM == PM_2) return t * t;
    if (PM == PM_3) return t * t * t;
    if (PM == PM_4) { float t2 = t * t; return t2 * t2; }
    return fast_ex2(p * fast_lg2(t));
}

template <int PM>
__device__ __forceinline__ float fold(float acc, float v, float eps, float p) {
    if (PM == PM_MAX) return fmaxf(acc, v);
    return acc + term<PM>(v, eps, p);
}

template <int PM>
__device__ __forceinline__ float fold4(float acc, const float4& v, float eps, float p) {
    acc = fold<PM>(acc, v.x, eps, p);
    acc = fold<PM>(acc, v.y, eps, p);
    acc = fold<PM>(acc, v.z, eps, p);
    acc = fold<PM>(acc, v.w, eps, p);
    return acc;
}

__device__ __forceinline__ int classify_p(int pool_mode, float p) {
    if (pool_mode == CIR_POOL_MAC) return PM_MAX;
    if (pool_mode == CIR_POOL_SPOC) return PM_MEAN;
    if (p == 3.0f) return PM_3;
    if (p == 2.0f) return PM_2;
    if (p == 1.0f) return PM_1;
    if (p == 4.0f) return PM_4;
    return PM_GENERAL;
}

// warp-uniform dispatch of one row's 8 preloaded vectors
__device__ __forceinline__ float fold_row8(int pm, const float4 (&v)[8], const bool (&ok)[8],
                                           float acc, float eps, float p) {
#define CIR_FOLD_CASE(PMV)                                                     \
    case PMV: {                                                                \
        _Pragma("unroll") for (int j = 0; j < 8; ++j)                          \
            if (ok[j]) acc = fold4<PMV>(acc, v[j], eps, p);                    \
        break;                                                                 \
    }
    switch (pm) {
        CIR_FOLD_CASE(PM_3)
        CIR_FOLD_CASE(PM_2)
        CIR_FOLD_CASE(PM_1)
        CIR_FOLD_CASE(PM_4)
        CIR_FOLD_CASE(PM_MAX)
        CIR_FOLD_CASE(PM_MEAN)
        default: {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (ok[j]) acc = fold4<PM_GENERAL>(acc, v[j], eps, p);
        }
    }
#undef CIR_FOLD_CASE
    return acc;
}

__device__ __forceinline__ float fold_scalar(int pm, float acc, float v, float eps, float p) {
    switch (pm) {
        case PM_3: return fold<PM_3>(acc, v, eps, p);
        case PM_2: return fold<PM_2>(acc, v, eps, p);
        case PM_1: return fold<PM_1>(acc, v, eps, p);
        case PM_4: return fold<PM_4>(acc, v, eps, p);
        case PM_MAX: return fold<PM_MAX>(acc, v, eps, p);
        case PM_MEAN: return fold<PM_MEAN>(acc, v, eps, p);
        default: return fold<PM_GENERAL>(acc, v, eps, p);
    }
}

__device__ __forceinline__ float finish_row(int pm, float acc, int HW, float p) {
    // acc already reduced over the warp
    if (pm == PM_MAX) return acc;
    float mean = acc / (float)HW;
    if (pm == PM_MEAN || pm == PM_1) return mean;
    // reference: .pow(1. / self.p) with the reciprocal rounded to fp32 (pools.py:38)
    return powf(mean, 1.0f / p);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) tail_fused_kernel(TailParams P) {
    extern __shared__ __align__(16) float Ws[];   // [TAIL_JMAX][TAIL_KC] (phase B only)
    cg::grid_group grid = cg::this_grid();

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool pool_only = (P.flags & CIR_TAIL_POOL_ONLY) != 0;
    const bool whiten = !pool_only && !(P.flags & CIR_TAIL_NO_WHITEN);

    auto stage_W = [&](int chunk, int kc0) {
        const int j0 = chunk * P.jch;
        const int J = min(P.jch, P.D_out - j0);
        const int kw = min(TAIL_KC, P.C - kc0);
        const int nvr = kw >> 2;
        for (int i = tid; i < J * nvr; i += TAIL_THREADS) {
            const int j = i / nvr, kv = i - j * nvr;
            cp_async16(&Ws[j * TAIL_KC + kv * 4], P.Wt + (size_t)(j0 + j) * P.C + kc0 + kv * 4);
        }
        cp_async_commit();
    };

    stamp(P, 0);
    int staged_chunk = -1, staged_kc0 = -1;
    if (whiten && (int)blockIdx.x < P.n_chunks) {
        stage_W(blockIdx.x, 0);            // lands while phase A streams the map
        staged_chunk = blockIdx.x;
        staged_kc0 = 0;
    }

    // ------------------------------------------------------------------ phase A
    {
        const long long rows = (long long)P.N * P.C;
        const long long pairs = (rows + 1) >> 1;
        const long long gw = (long long)blockIdx.x * TAIL_WARPS + warp;
        const long long nw = (long long)gridDim.x * TAIL_WARPS;
        const int HW = P.HW;
        for (long long pr = gw; pr < pairs; pr += nw) {
            const long long r0 = pr * 2, r1 = r0 + 1;
            const bool has1 = r1 < rows;
            const int c0 = (int)(r0 % P.C), c1 = (int)(r1 % P.C);
            const float p0 = P.pool_mode == CIR_POOL_GEM ? __ldg(P.p + c0 * P.p_stride) : 1.0f;
            const float p1 = (P.pool_mode == CIR_POOL_GEM && has1) ? __ldg(P.p + c1 * P.p_stride) : p0;
            const int pm0 = classify_p(P.pool_mode, p0), pm1 = classify_p(P.pool_mode, p1);
            float a0 = pm0 == PM_MAX ? -INFINITY : 0.0f;
            float a1 = a0;
            const float* x0 = P.x + r0 * HW;
            const float* x1 = P.x + (has1 ? r1 : r0) * HW;
            if (P.vec_ok) {
                const int nvec = HW >> 2;
                const float4* v0 = reinterpret_cast<const float4*>(x0);
                const float4* v1 = reinterpret_cast<const float4*>(x1);
                for (int base = 0; base < nvec; base += 256) {
                    float4 u0[8], u1[8];
                    bool ok[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i = base + j * 32 + lane;
                        ok[j] = i < nvec;
                        if (ok[j]) u0[j] = ld_stream_f4(v0 + i);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i = base + j * 32 + lane;
                        if (ok[j]) u1[j] = ld_stream_f4(v1 + i);
                    }
                    a0 = fold_row8(pm0, u0, ok, a0, P.eps_gem, p0);
                    a1 = fold_row8(pm1, u1, ok, a1, P.eps_gem, p1);
                }
            } else {
                for (int i = lane; i < HW; i += 32) {
                    a0 = fold_scalar(pm0, a0, ld_stream_f1(x0 + i), P.eps_gem, p0);
                    a1 = fold_scalar(pm1, a1, ld_stream_f1(x1 + i), P.eps_gem, p1);
                }
            }
            a0 = pm0 == PM_MAX ? warp_max(a0) : warp_sum(a0);
            a1 = pm1 == PM_MAX ? warp_max(a1) : warp_sum(a1);
            if (lane == 0) {
                const long long n0 = r0 / P.C;
                P.pooled[n0 * P.pooled_ld + c0] = finish_row(pm0, a0, HW, p0);
                if (has1) {
                    const long long n1 = r1 / P.C;
                    P.pooled[n1 * P.pooled_ld + c1] = finish_row(pm1, a1, HW, p1);
                }
            }
        }
    }
    stamp(P, 1);
    if (pool_only) return;

    grid.sync();
    stamp(P, 2);

    // ------------------------------------------------------------------ no whitening
    if (!whiten) {
        __shared__ float red[TAIL_WARPS];
        for (int n = blockIdx.x; n < P.N; n += gridDim.x) {
            const float* g = P.pooled + (size_t)n * P.pooled_ld;
            float ss = 0.0f;
            for (int c = tid; c < P.C; c += TAIL_THREADS) { float v = __ldcg(g + c); ss += v * v; }
            ss = warp_sum(ss);
            __syncthreads();
            if (lane == 0) red[warp] = ss;
            __syncthreads();
            float tot = 0.0f;
#pragma unroll
            for (int w = 0; w < TAIL_WARPS; ++w) tot += red[w];
            const float denom = sqrtf(tot) + P.eps_l2;
            for (int c = tid; c < P.C; c += TAIL_THREADS)
                P.out[(size_t)n * P.out_ld + c] = __ldcg(g + c) / denom;
        }
        return;
    }

    // ------------------------------------------------------------------ phase B
    // Every CTA reads all pooled vectors from L2; CTAs start at different K offsets so that they do not
    // all hit the same L2 lines at the same moment, and each lane prefetches its next 4 x float4 while it
    // multiplies the current ones.
    for (int chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
        const int j0 = chunk * P.jch;
        const int J = min(P.jch, P.D_out - j0);
        for (int n0 = 0; n0 < P.N; n0 += TAIL_WARPS * TAIL_IMG) {
            const int nb = n0 + warp * TAIL_IMG;
            float acc[TAIL_IMG][TAIL_JMAX];
            float ss[TAIL_IMG];
#pragma unroll
            for (int i = 0; i < TAIL_IMG; ++i) {
                ss[i] = 0.0f;
#pragma unroll
                for (int j = 0; j < TAIL_JMAX; ++j) acc[i][j] = 0.0f;
            }
            for (int kc0 = 0; kc0 < P.C; kc0 += TAIL_KC) {
                if (staged_chunk != chunk || staged_kc0 != kc0) {
                    __syncthreads();      // everyone is done with the previous tile
                    stage_W(chunk, kc0);
                    staged_chunk = chunk;
                    staged_kc0 = kc0;
                }
                cp_async_wait_all();
                __syncthreads();
                const int kw = min(TAIL_KC, P.C - kc0);
                const int iters = (kw + 127) >> 7;
                if (nb < P.N) {
                    const float* gbase = P.pooled + (size_t)nb * P.pooled_ld + kc0 + lane * 4;
                    auto load_g = [&](int it, float4 (&g)[TAIL_IMG]) {
                        const int k = it * 128 + lane * 4;
#pragma unroll
                        for (int i = 0; i < TAIL_IMG; ++i) {
                            g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (k < kw && nb + i < P.N)
                                g[i] = __ldcg(reinterpret_cast<const float4*>(gbase + (size_t)i * P.pooled_ld + it * 128));
                        }
                    };
                    int it = (int)(blockIdx.x % (unsigned)iters);
                    float4 g_next[TAIL_IMG];
                    load_g(it, g_next);
                    for (int t = 0; t < iters; ++t) {
                        float4 g[TAIL_IMG];
#pragma unroll
                        for (int i = 0; i < TAIL_IMG; ++i) g[i] = g_next[i];
                        const int k = it * 128 + lane * 4;
                        if (++it == iters) it = 0;
                        if (t + 1 < iters) load_g(it, g_next);
                        if (k < kw) {
#pragma unroll
                            for (int i = 0; i < TAIL_IMG; ++i) ss[i] += dot4(g[i], g[i]);
#pragma unroll
                            for (int j = 0; j < TAIL_JMAX; ++j) {
                                if (j < J) {
                                    const float4 w = *reinterpret_cast<const float4*>(&Ws[j * TAIL_KC + k]);
#pragma unroll
                                    for (int i = 0; i < TAIL_IMG; ++i) acc[i][j] += dot4(w, g[i]);
                                }
                            }
                        }
                    }
                }
            }
            if (nb < P.N) {
#pragma unroll
                for (int i = 0; i < TAIL_IMG; ++i) {
                    ss[i] = warp_sum(ss[i]);
#pragma unroll
                    for (int j = 0; j < TAIL_JMAX; ++j) acc[i][j] = warp_sum(acc[i][j]);
                }
                const float bj = (lane < J && P.bias) ? __ldg(P.bias + j0 + lane) : 0.0f;
#pragma unroll
                for (int i = 0; i < TAIL_IMG; ++i) {
                    if (nb + i < P.N) {
                        // first L2N folded in: W.(g/(|g|+eps)) == (W.g)/(|g|+eps)
                        const float inv = 1.0f / (sqrtf(ss[i]) + P.eps_l2);
                        float y = 0.0f;
#pragma unroll
                        for (int j = 0; j < TAIL_JMAX; ++j)
                            if (lane == j) y = acc[i][j] * inv + bj;
                        if (lane < J) P.out[(size_t)(nb + i) * P.out_ld + j0 + lane] = y;
                        const float sq = warp_sum(lane < J ? y * y : 0.0f);
                        if (lane == 0) P.partial[(size_t)chunk * P.N + nb + i] = sq;
                    }
                }
            }
        }
    }

    stamp(P, 3);
    grid.sync();
    stamp(P, 4);

    // ------------------------------------------------------------------ phase C: second L2N
    // per image: sum the chunks' partial sums of squares in a fixed order (deterministic), then rescale
    // the J columns this CTA wrote.
    {
        constexpr int CB = 64;                       // images per pass
        constexpr int PARTS = TAIL_THREADS / CB;     // 8 chunk subsets
        __shared__ float redc[PARTS][CB];
        __shared__ float denom_s[CB];
        for (int chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
            const int j0 = chunk * P.jch;
            const int J = min(P.jch, P.D_out - j0);
            for (int n0 = 0; n0 < P.N; n0 += CB) {
                const int nl = tid & (CB - 1), part = tid / CB;
                float sacc = 0.0f;
                if (n0 + nl < P.N)
                    for (int ch = part; ch < P.n_chunks; ch += PARTS) sacc += __ldcg(P.partial + (size_t)ch * P.N + n0 + nl);
                redc[part][nl] = sacc;
                __syncthreads();
                if (tid < CB) {
                    float tot = 0.0f;
#pragma unroll
                    for (int pp = 0; pp < PARTS; ++pp) tot += redc[pp][tid];
                    denom_s[tid] = sqrtf(tot) + P.eps_l2;
                }
                __syncthreads();
                const int cnt = min(CB, P.N - n0) * J;
                for (int e = tid; e < cnt; e += TAIL_THREADS) {
                    const int n = e / J, j = e - n * J;
                    float* o = P.out + (size_t)(n0 + n) * P.out_ld + j0 + j;
                    *o = __ldcg(o) / denom_s[n];
                }
                __syncthreads();
            }
        }
    }
    stamp(P, 5);
}

// ---------------------------------------------------------------------- row L2N (A2)
__global__ void __launch_bounds__(256) l2n_rows_kernel(const float* __restrict__ X, long long N, int C,
                                                       long long ldx, float eps, float* __restrict__ out,
                                                       long long out_ld) {
    // one warp per row
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float* x = X + row * ldx;
    float ss = 0.0f;
    for (int c = lane; c < C; c += 32) { float v = x[c]; ss += v * v; }
    ss = warp_sum(ss);
    const float denom = sqrtf(ss) + eps;
    float* o = out + row * out_ld;
    for (int c = lane; c < C; c += 32) o[c] = x[c] / denom;
}

}  // namespace cir

using namespace cir;

extern "C" int cir_tail_workspace_bytes(int N, int C, int D_out, size_t* bytes) {
    CIR_REQUIRE(bytes && N > 0 && C > 0 && D_out > 0, CIR_ERR_INVALID_ARG, "cir_tail_workspace_bytes: bad arguments");
    // pooled [N, C] + partial [D_out, N] (n_chunks <= D_out)
    *bytes = align_up((size_t)N * C * 4, 256) + align_up((size_t)D_out * N * 4, 256) + TAIL_STAMP_BYTES;
    return CIR_OK;
}

extern "C" int cir_tail_fwd(const float* x, int N, int C, int H, int W, const float* p, int p_stride,
                            float eps_gem, float eps_l2, int pool_mode, const float* Wt,
                            const float* bias, int D_out, float* out, int out_ld, void* workspace,
                            size_t workspace_bytes, unsigned flags, void* stream) {
    CIR_REQUIRE(x && out && N > 0 && C > 0 && H > 0 && W > 0, CIR_ERR_INVALID_ARG,
                "cir_tail_fwd: null pointer or empty shape (N=%d C=%d H=%d W=%d)", N, C, H, W);
    CIR_REQUIRE(pool_mode >= CIR_POOL_GEM && pool_mode <= CIR_POOL_SPOC, CIR_ERR_INVALID_ARG,
                "cir_tail_fwd: unknown pool_mode %d", pool_mode);
    CIR_REQUIRE(pool_mode != CIR_POOL_GEM || p, CIR_ERR_INVALID_ARG, "cir_tail_fwd: GeM needs p");
    CIR_REQUIRE(p_stride == 0 || p_stride == 1, CIR_ERR_INVALID_ARG, "cir_tail_fwd: p_stride must be 0 or 1");
    const bool pool_only = flags & CIR_TAIL_POOL_ONLY;
    const bool whiten = !pool_only && !(flags & CIR_TAIL_NO_WHITEN);
    if (!whiten) D_out = C;
    CIR_REQUIRE(out_ld >= D_out, CIR_ERR_INVALID_ARG, "cir_tail_fwd: out_ld %d < D_out %d", out_ld, D_out);
    const DeviceInfo& dev = device_info();
    CIR_REQUIRE(dev.coop, CIR_ERR_UNSUPPORTED, "cir_tail_fwd: device lacks cooperative launch");

    TailParams P{};
    P.x = x; P.N = N; P.C = C; P.HW = H * W;
    P.p = p; P.p_stride = p_stride; P.eps_gem = eps_gem; P.eps_l2 = eps_l2; P.pool_mode = pool_mode;
    P.Wt = Wt; P.bias = bias; P.D_out = D_out; P.out = out; P.out_ld = out_ld; P.flags = flags;
    P.vec_ok = ((P.HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);

    const int grid = dev.num_sms;
    size_t smem = 0;
    if (pool_only) {
        P.pooled = out; P.pooled_ld = out_ld;
    } else {
        size_t need = 0;
        cir_tail_workspace_bytes(N, C, D_out, &need);
        CIR_REQUIRE(workspace && workspace_bytes >= need, CIR_ERR_WORKSPACE,
                    "cir_tail_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
        CIR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, CIR_ERR_INVALID_ARG,
                    "cir_tail_fwd: workspace must be 16 B aligned");
        P.pooled = static_cast<float*>(workspace);
        P.pooled_ld = C;
        P.partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + align_up((size_t)N * C * 4, 256));
        if (flags & CIR_TAIL_DEBUG_STAMPS)
            P.stamps = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + need - TAIL_STAMP_BYTES);
    }
    if (whiten) {
        CIR_REQUIRE(Wt, CIR_ERR_INVALID_ARG, "cir_tail_fwd: whitening needs Wt");
        CIR_REQUIRE((C & 3) == 0 && (reinterpret_cast<uintptr_t>(Wt) & 15) == 0, CIR_ERR_UNSUPPORTED,
                    "cir_tail_fwd: whitening needs C %% 4 == 0 and a 16 B aligned Wt (C=%d)", C);
        int jch = (D_out + grid - 1) / grid;
        if (jch > TAIL_JMAX) jch = TAIL_JMAX;
        P.jch = jch;
        P.n_chunks = (D_out + jch - 1) / jch;
        smem = (size_t)TAIL_JMAX * TAIL_KC * sizeof(float);
        CIR_REQUIRE((int)smem <= dev.max_smem_optin, CIR_ERR_UNSUPPORTED, "cir_tail_fwd: shared memory");
    }
    static thread_local int attr_set_dev = -1;
    if (attr_set_dev != dev.device) {
        CIR_CHECK_CUDA(cudaFuncSetAttribute(tail_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            TAIL_JMAX * TAIL_KC * (int)sizeof(float)));
        attr_set_dev = dev.device;
    }
    void* args[] = {&P};
    if (pool_only) {
        // no grid barrier on this path: a plain launch is enough
        tail_fused_kernel<<<grid, TAIL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(P);
        CIR_CHECK_CUDA(cudaGetLastError());
    } else {
        CIR_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)tail_fused_kernel, dim3(grid), dim3(TAIL_THREADS),
                                                   args, smem, static_cast<cudaStream_t>(stream)));
    }
    count_launch();
    return CIR_OK;
}

extern "C" int cir_l2n_rows(const float* X, int64_t N, int C, int64_t ldx, float eps, float* out,
                            int64_t out_ld, void* stream) {
    CIR_REQUIRE(X && out && N >= 0 && C > 0 && ldx >= C && out_ld >= C, CIR_ERR_INVALID_ARG, "cir_l2n_rows: bad arguments");
    if (N == 0) return CIR_OK;
    const long long blocks = (N + 7) / 8;
    l2n_rows_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, N, C, ldx, eps, out, out_ld);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}
