// Small bandwidth-bound kernels around the search path:
//   pack_bf16_kernel     fp32 descriptor rows -> K-major bf16 search operands (plain or bf16x3 split)
//   bias_l2n_rows_kernel y = (x + b) / (||x + b||_2 + eps) per row: the tail of whitenapply
//                        (cirtorch/utils/whiten.py:8-12) after the projection GEMM
//   qe_aggregate_kernel  alpha-QE / DBA neighbour aggregation
//   mine_filter_kernel   greedy cluster-exclusion walk of create_epoch_tuples
//                        (cirtorch/datasets/globalFeatures/tuples_dataset.py:328-345)
#include "common.cuh"

namespace cir {

// ------------------------------------------------------------------------------ pack
// one thread per 4 source elements; dst row = n_split segments of Dp = round_up(D, 64)
__global__ void __launch_bounds__(256)
pack_bf16_kernel(const float* __restrict__ src, long long rows, int D, long long src_ld, __nv_bfloat16* __restrict__ dst,
                 long long dst_ld, int Dp, int n_split, int role) {
    const int groups = Dp >> 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= rows * groups) return;
    const long long r = gid / groups;
    const int d0 = (int)(gid - r * groups) * 4;
    float v[4];
    const float* s = src + r * src_ld + d0;
    if (d0 + 3 < D && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) {
        const float4 t = ld_stream_f4(reinterpret_cast<const float4*>(s));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = d0 + i < D ? __ldg(s + i) : 0.0f;
    }
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hi[i] = __float2bfloat16_rn(v[i]);
        lo[i] = __float2bfloat16_rn(v[i] - __bfloat162float(hi[i]));
    }
    __nv_bfloat16* o = dst + r * dst_ld + d0;
    auto store4 = [&](__nv_bfloat16* p, const __nv_bfloat16 (&x)[4]) {
        uint2 u;
        u.x = (uint32_t)__bfloat16_as_ushort(x[0]) | ((uint32_t)__bfloat16_as_ushort(x[1]) << 16);
        u.y = (uint32_t)__bfloat16_as_ushort(x[2]) | ((uint32_t)__bfloat16_as_ushort(x[3]) << 16);
        *reinterpret_cast<uint2*>(p) = u;
    };
    if (n_split == 1) {
        store4(o, hi);
    } else {
        // query side [hi, hi, lo], database side [hi, lo, hi]:  q'.d' = hi.hi + hi.lo + lo.hi
        store4(o, hi);
        if (role == 0) { store4(o + Dp, hi); store4(o + 2 * Dp, lo); }
        else           { store4(o + Dp, lo); store4(o + 2 * Dp, hi); }
    }
}

// ------------------------------------------------------------------------------ bias + L2N
__global__ void __launch_bounds__(256)
bias_l2n_rows_kernel(const float* __restrict__ X, long long N, int C, long long ldx, const float* __restrict__ bias,
                     float eps, float* __restrict__ out, long long out_ld) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float* x = X + row * ldx;
    float ss = 0.0f;
    for (int c = lane; c < C; c += 32) {
        const float v = x[c] + (bias ? __ldg(bias + c) : 0.0f);
        ss += v * v;
    }
    ss = warp_sum(ss);
    const float denom = eps >= 0.0f ? sqrtf(ss) + eps : 1.0f;
    float* o = out + row * out_ld;
    for (int c = lane; c < C; c += 32) o[c] = (x[c] + (bias ? __ldg(bias + c) : 0.0f)) / denom;
}

// ------------------------------------------------------------------------------ alpha-QE
constexpr int QE_THREADS = 256;
__global__ void __launch_bounds__(QE_THREADS)
qe_aggregate_kernel(const float* __restrict__ q32, const float* __restrict__ db32, long long N, int D,
                    const int32_t* __restrict__ idx, const float* __restrict__ scores, int klist, int ld_k, int k_use,
                    float alpha, long long self_base, float eps_l2, float* __restrict__ out) {
    __shared__ float red[QE_THREADS / 32];
    __shared__ float s_tot;
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    // each thread owns elements d = tid, tid + 256, ... (kept in global `out` between passes)
    float ss = 0.0f;
    for (int d = tid; d < D; d += QE_THREADS) {
        float acc = q32 ? __ldg(q32 + (size_t)q * D + d) : 0.0f;      // q32 == NULL: the neighbour sum alone (one shard's part)
        int used = 0;
        for (int i = 0; i < klist && used < k_use; ++i) {
            const int32_t ix = __ldg(idx + (size_t)q * ld_k + i);
            if (ix < 0 || ix >= N) continue;
            if (self_base >= 0 && (long long)ix == self_base + q) continue;
            ++used;
            const float s = __ldg(scores + (size_t)q * ld_k + i);
            const float w = s > 0.0f ? powf(s, alpha) : 0.0f;
            acc = fmaf(w, __ldg(db32 + (size_t)ix * D + d), acc);
        }
        out[(size_t)q * D + d] = acc;
        ss += acc * acc;
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
        for (int w = 0; w < QE_THREADS / 32; ++w) t += red[w];
        s_tot = t;
    }
    __syncthreads();
    if (eps_l2 < 0.0f) return;                            // raw sum: the caller adds the shards' parts, then normalises
    const float denom = sqrtf(s_tot) + eps_l2;
    for (int d = tid; d < D; d += QE_THREADS) out[(size_t)q * D + d] /= denom;
}

// ------------------------------------------------------------------------------ mining
// one warp per query: lane 0 walks the ranked candidates, all lanes compute the distances
constexpr int MINE_MAX_NNUM = 32;
__global__ void __launch_bounds__(128)
mine_filter_kernel(const int32_t* __restrict__ cand, int Q, int Kc, const int32_t* __restrict__ pool_cluster, long long P,
                   const int32_t* __restrict__ q_cluster, int nnum, const float* __restrict__ q32,
                   const float* __restrict__ pool32, int D, int32_t* __restrict__ out_sel, int32_t* __restrict__ out_count,
                   float* __restrict__ out_dist, const float* __restrict__ cand_score, const float* __restrict__ scan_tail,
                   float margin, int32_t* __restrict__ out_open) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= Q) return;
    int taken_cluster[MINE_MAX_NNUM];
    int count = 0, last_pos = -1;
    const int qc = __ldg(q_cluster + q);
    const bool vec = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(q32) | reinterpret_cast<uintptr_t>(pool32)) & 15) == 0;
    // 32 candidates and their clusters are fetched at once (one lane each); the greedy walk itself then runs on
    // registers -- the serial version paid two dependent memory latencies per candidate
    for (int r0 = 0; r0 < Kc && count < nnum; r0 += 32) {
        const int32_t jl = r0 + lane < Kc ? __ldg(cand + (size_t)q * Kc + r0 + lane) : -1;
        const int cl = (jl >= 0 && jl < P) ? __ldg(pool_cluster + jl) : -1;
        const int lim = min(32, Kc - r0);
        for (int t = 0; t < lim && count < nnum; ++t) {
            const int32_t j = __shfl_sync(0xffffffffu, jl, t);
            const int c = __shfl_sync(0xffffffffu, cl, t);
            if (j < 0 || j >= P) continue;
            bool clash = c == qc;
#pragma unroll 1
            for (int u = 0; u < count; ++u) clash |= taken_cluster[u] == c;
            if (clash) continue;
            taken_cluster[count] = c;
            last_pos = r0 + t;
            if (lane == 0) out_sel[(size_t)q * nnum + count] = j;
            if (q32 && pool32 && out_dist) {
                const float* a = q32 + (size_t)q * D;
                const float* b = pool32 + (size_t)j * D;
                float ss = 0.0f;
                if (vec) {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
                    for (int d = lane * 4; d < D; d += 128) {
                        const float4 x = __ldg(reinterpret_cast<const float4*>(a + d));
                        const float4 y = __ldg(reinterpret_cast<const float4*>(b + d));
                        const float d0 = x.x - y.x + 1e-6f, d1 = x.y - y.y + 1e-6f, d2 = x.z - y.z + 1e-6f, d3 = x.w - y.w + 1e-6f;
                        s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
                    }
                    ss = (s0 + s1) + (s2 + s3);
                } else {
                    for (int d = lane; d < D; d += 32) {
                        const float df = __ldg(a + d) - __ldg(b + d) + 1e-6f;
                        ss = fmaf(df, df, ss);
                    }
                }
                ss = warp_sum(ss);
                if (lane == 0) out_dist[(size_t)q * nnum + count] = sqrtf(ss);
            }
            ++count;
        }
    }
    if (lane == 0) {
        out_count[q] = count;
        if (out_open) {
            // Is the walk conclusive?  The list is the best Kc rows of an approximate scan, re-ordered by exact scores: a row
            // just outside it could outrank the walk's tail unless the exact score of the last negative taken clears the
            // scan score of the list's last entry by `margin` (-inf there = every row the query may take is in the list).
            const float tail = scan_tail ? __ldg(scan_tail + q) : -INFINITY;
            bool open = false;
            if (tail > -INFINITY)
                open = count < nnum || last_pos < 0 || !cand_score || __ldg(cand_score + (size_t)q * Kc + last_pos) - tail < margin;
            out_open[q] = open ? 1 : 0;
        }
        for (int t = count; t < nnum; ++t) {
            out_sel[(size_t)q * nnum + t] = -1;
            if (out_dist) out_dist[(size_t)q * nnum + t] = 0.0f;
        }
    }
}

// ------------------------------------------------------------------------------ PowerLaw
// cirtorch/modules/normalizations.py:19-27:  y = sign(x + eps) * sqrt(|x + eps|)
__global__ void __launch_bounds__(256)
powerlaw_kernel(const float* __restrict__ x, long long n, float eps, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i] + eps;
    out[i] = copysignf(sqrtf(fabsf(v)), v) * (v == 0.0f ? 0.0f : 1.0f);
}

}  // namespace cir

using namespace cir;

extern "C" int cir_powerlaw(const float* x, int64_t n, float eps, float* out, void* stream) {
    CIR_REQUIRE(x && out && n >= 0, CIR_ERR_INVALID_ARG, "cir_powerlaw: bad arguments");
    if (n == 0) return CIR_OK;
    const long long blocks = (n + 255) / 256;
    CIR_REQUIRE(blocks <= 0x7fffffffll, CIR_ERR_UNSUPPORTED, "cir_powerlaw: too many elements");
    powerlaw_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, eps, out);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}


extern "C" int cir_pack_bf16(const float* src, int64_t rows, int D, int64_t src_ld, void* dst, int64_t dst_ld, int n_split,
                             int role, void* stream) {
    CIR_REQUIRE(src && dst && rows >= 0 && D > 0 && src_ld >= D, CIR_ERR_INVALID_ARG, "cir_pack_bf16: bad arguments");
    CIR_REQUIRE(n_split == 1 || n_split == 3, CIR_ERR_INVALID_ARG, "cir_pack_bf16: n_split must be 1 or 3");
    CIR_REQUIRE(role == 0 || role == 1, CIR_ERR_INVALID_ARG, "cir_pack_bf16: role must be 0 (query) or 1 (database)");
    const int Dp = (D + 63) / 64 * 64;
    CIR_REQUIRE(dst_ld >= (int64_t)n_split * Dp && (dst_ld % 4) == 0 && ((uintptr_t)dst & 7) == 0, CIR_ERR_INVALID_ARG,
                "cir_pack_bf16: dst_ld=%lld must be >= %d and the destination 8 B aligned", (long long)dst_ld, n_split * Dp);
    if (rows == 0) return CIR_OK;
    const long long total = rows * (Dp >> 2);
    const long long blocks = (total + 255) / 256;
    CIR_REQUIRE(blocks <= 0x7fffffffll, CIR_ERR_UNSUPPORTED, "cir_pack_bf16: too many elements");
    pack_bf16_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, rows, D, src_ld, static_cast<__nv_bfloat16*>(dst), dst_ld, Dp, n_split, role);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}

extern "C" int cir_bias_l2n_rows(const float* X, int64_t N, int C, int64_t ldx, const float* bias, float eps_l2, float* out,
                                 int64_t out_ld, void* stream) {
    CIR_REQUIRE(X && out && N >= 0 && C > 0 && ldx >= C && out_ld >= C, CIR_ERR_INVALID_ARG, "cir_bias_l2n_rows: bad arguments");
    if (N == 0) return CIR_OK;
    const long long blocks = (N + 7) / 8;
    bias_l2n_rows_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, N, C, ldx, bias, eps_l2, out, out_ld);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}

extern "C" int cir_qe_aggregate(const float* q32, int Q, const float* db32, int64_t N, int D, const int32_t* idx,
                                const float* scores, int klist, int ld_k, int k_use, float alpha, int64_t self_base,
                                float eps_l2, float* out, void* stream) {
    CIR_REQUIRE(db32 && idx && scores && out, CIR_ERR_INVALID_ARG, "cir_qe_aggregate: null pointer");
    CIR_REQUIRE(Q >= 0 && N > 0 && D > 0 && klist >= 0 && ld_k >= klist && k_use >= 0, CIR_ERR_INVALID_ARG,
                "cir_qe_aggregate: bad shape");
    if (Q == 0) return CIR_OK;
    qe_aggregate_kernel<<<Q, QE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(q32, db32, N, D, idx, scores, klist, ld_k, k_use,
                                                                                alpha, self_base, eps_l2, out);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}

extern "C" int cir_mine_filter(const int32_t* cand, int Q, int Kc, const int32_t* pool_cluster, int64_t P,
                               const int32_t* q_cluster, int nnum, const float* q32, const float* pool32, int D,
                               int32_t* out_sel, int32_t* out_count, float* out_dist, const float* cand_score,
                               const float* scan_tail, float margin, int32_t* out_open, void* stream) {
    CIR_REQUIRE(cand && pool_cluster && q_cluster && out_sel && out_count, CIR_ERR_INVALID_ARG, "cir_mine_filter: null pointer");
    CIR_REQUIRE(Q >= 0 && Kc >= 1 && P > 0 && nnum >= 1 && nnum <= MINE_MAX_NNUM, CIR_ERR_INVALID_ARG,
                "cir_mine_filter: bad shape (nnum <= %d)", MINE_MAX_NNUM);
    if (Q == 0) return CIR_OK;
    mine_filter_kernel<<<(Q + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(cand, Q, Kc, pool_cluster, P, q_cluster, nnum,
                                                                                   q32, pool32, D, out_sel, out_count, out_dist,
                                                                                   cand_score, scan_tail, margin, out_open);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}
