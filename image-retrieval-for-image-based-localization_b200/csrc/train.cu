// Training-side kernels of the descriptor head:
//   tuple_loss_kernel     contrastive_loss / triplet_loss of cirtorch/modules/losses.py:7-46 on tuples (q, p, n1..nk):
//                         the loss AND its gradient w.r.t. every descriptor in one launch
//   l2n_bwd_rows_kernel   gradient of v / (||v|| + eps) per row (the two L2Ns of globalHead.forward, global_head.py:57-64)
//   colsum_rows_kernel    column sums of a small [N, C] matrix (dL/db of the whitening Linear)
//   gem_dp_kernel         dL/dp of GeM from the pooled values, their gradient and S = sum t^p ln t (tail_bwd.cu)
// All of them touch [N, C]-sized data only (N = images of a batch): latency-bound, one pass, deterministic reductions.
#include "common.cuh"

namespace cir {

constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    // fixed shuffle tree, then the 8 warp sums in order: deterministic
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < LOSS_THREADS / 32; ++w) t += red[w];
    return t;
}

// One block per tuple of S consecutive rows of x [n_tuples * S, D].  label: -1 query, 1 positive, 0 negative.
//   kind 0 (contrastive, losses.py:7-23): every non-query row j pairs with the tuple's query (its FIRST row, :13):
//       d = sqrt(sum (xq - xj + eps)^2),  y = 0.5 l d^2 + 0.5 (1 - l) max(margin - d, 0)^2
//   kind 1 (triplet, losses.py:26-46): a = the query row, p = the positive row, every negative n:
//       y = max(|a - p|^2 - |a - n|^2 + margin, 0)
// partial[t] = loss of tuple t; the last block to finish adds the partials in tuple order into loss[0].
__global__ void __launch_bounds__(LOSS_THREADS)
tuple_loss_kernel(const float* __restrict__ x, long long ldx, int n_tuples, int S, int D, const int32_t* __restrict__ label,
                  int kind, float margin, float eps, float* __restrict__ loss, float* __restrict__ grad, long long ldg,
                  float* __restrict__ partial, unsigned* __restrict__ counter) {
    __shared__ float red[LOSS_THREADS / 32];
    __shared__ int s_q, s_p;
    __shared__ bool s_last;
    const int t = blockIdx.x;
    const int tid = threadIdx.x;
    const long long row0 = (long long)t * S;
    if (tid == 0) {
        int qi = 0, pi = -1;
        if (kind == 1) {
            qi = -1;
            for (int j = 0; j < S; ++j) {
                const int l = label[row0 + j];
                if (l == -1 && qi < 0) qi = j;
                if (l == 1 && pi < 0) pi = j;
            }
        }
        s_q = qi;
        s_p = pi;
    }
    __syncthreads();
    const int qi = s_q, pi = s_p;
    float tuple_loss = 0.0f;
    const float* xq = x + (row0 + (qi < 0 ? 0 : qi)) * ldx;
    if (grad) {                                  // the query / positive rows collect contributions from several pairs
        for (int j = 0; j < S; ++j)
            for (int d = tid; d < D; d += LOSS_THREADS) grad[(row0 + j) * ldg + d] = 0.0f;
        __syncthreads();
    }
    if (kind == 0) {
        for (int j = 0; j < S; ++j) {
            const int l = label[row0 + j];
            if (l == -1) continue;
            const float* xj = x + (row0 + j) * ldx;
            float ss = 0.0f;
            for (int d = tid; d < D; d += LOSS_THREADS) {
                const float df = xq[d] - xj[d] + eps;
                ss = fmaf(df, df, ss);
            }
            const float dist = sqrtf(block_sum_256(ss, red));
            const float lf = (float)l;
            const float hinge = fmaxf(margin - dist, 0.0f);
            tuple_loss += 0.5f * lf * dist * dist + 0.5f * (1.0f - lf) * hinge * hinge;
            if (grad) {
                // dy/d(dist) = l dist - (1 - l) hinge;  d(dist)/d(dif) = (dif + eps) / dist
                const float c = dist > 0.0f ? (lf * dist - (1.0f - lf) * hinge) / dist : 0.0f;
                for (int d = tid; d < D; d += LOSS_THREADS) {
                    const float gq = c * (xq[d] - xj[d] + eps);
                    grad[(row0 + qi) * ldg + d] += gq;           // thread-private column d: no race
                    grad[(row0 + j) * ldg + d] = -gq;
                }
            }
        }
    } else if (qi >= 0 && pi >= 0) {
        const float* xp = x + (row0 + pi) * ldx;
        float sp = 0.0f;
        for (int d = tid; d < D; d += LOSS_THREADS) {
            const float df = xq[d] - xp[d];
            sp = fmaf(df, df, sp);
        }
        const float dist_pos = block_sum_256(sp, red);
        for (int j = 0; j < S; ++j) {
            if (label[row0 + j] != 0) continue;
            const float* xn = x + (row0 + j) * ldx;
            float sn = 0.0f;
            for (int d = tid; d < D; d += LOSS_THREADS) {
                const float df = xq[d] - xn[d];
                sn = fmaf(df, df, sn);
            }
            const float h = dist_pos - block_sum_256(sn, red) + margin;
            if (h > 0.0f) {
                tuple_loss += h;
                if (grad) {
                    for (int d = tid; d < D; d += LOSS_THREADS) {
                        const float a = xq[d], p = xp[d], n = xn[d];
                        grad[(row0 + qi) * ldg + d] += 2.0f * (n - p);
                        grad[(row0 + pi) * ldg + d] += -2.0f * (a - p);
                        grad[(row0 + j) * ldg + d] = 2.0f * (a - n);
                    }
                }
            }
        }
    }
    if (tid == 0) {
        partial[t] = tuple_loss;
        __threadfence();
        s_last = atomicAdd(counter, 1u) == (unsigned)(n_tuples - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        float acc = 0.0f;                          // tuple order, chunked by thread then the fixed block tree
        for (int i = tid; i < n_tuples; i += LOSS_THREADS) acc += __ldcg(partial + i);
        acc = block_sum_256(acc, red);
        if (tid == 0) {
            loss[0] = acc;
            *counter = 0u;
        }
    }
}

// out[n, :] = gout / (s + eps) - v (v . gout) / (s (s + eps)^2),  s = ||v[n, :]||;  one 256-thread block per row (a batch
// has 64 .. 448 rows: a warp per row left 8 blocks on 148 SMs and 16 us per call).
// unit_out (optional) = v / (s + eps): the forward value, needed by the caller for dL/dW.
__global__ void __launch_bounds__(256)
l2n_bwd_rows_kernel(const float* __restrict__ v, const float* __restrict__ gout, long long N, int C, float eps,
                    float* __restrict__ out, float* __restrict__ unit_out) {
    __shared__ float red[2][8];
    const long long row = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* a = v + row * C;
    const float* g = gout ? gout + row * C : nullptr;
    float ss = 0.0f, dot = 0.0f;
    for (int c = threadIdx.x; c < C; c += 256) {
        const float x = a[c];
        ss = fmaf(x, x, ss);
        if (g) dot = fmaf(x, g[c], dot);
    }
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    if (lane == 0) { red[0][warp] = ss; red[1][warp] = dot; }
    __syncthreads();
    ss = 0.0f; dot = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { ss += red[0][w]; dot += red[1][w]; }      // fixed order: deterministic
    const float s = sqrtf(ss);
    const float inv = 1.0f / (s + eps);
    const float k = s > 0.0f ? dot / (s * (s + eps) * (s + eps)) : 0.0f;
    for (int c = threadIdx.x; c < C; c += 256) {
        const float x = a[c];
        if (out) out[row * C + c] = g[c] * inv - x * k;
        if (unit_out) unit_out[row * C + c] = x * inv;
    }
}

// out[c] = sum_n X[n, c] (rows added in order: deterministic)
__global__ void __launch_bounds__(256)
colsum_rows_kernel(const float* __restrict__ X, long long N, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float acc = 0.0f;
    for (long long n = 0; n < N; ++n) acc += X[n * C + c];
    out[c] = acc;
}

// dL/dp of GeM (autograd of pools.py:37-38):  g = (mean t^p)^(1/p), t = max(x, eps), S = sum_hw t^p ln t
//   d g / d p = g * ( -ln g / p + S / (p HW g^p) ),   dL/dp = sum dg * dg/dp
// p_stride 1: one exponent per channel -> out[c], one thread per channel.  p_stride 0: one exponent -> out[0]: one block per
// image sums its C terms, the last block to finish adds the per-image partials in image order (deterministic).  (The first
// version summed all N*C terms in ONE block: 165 us of a 600 us training step.)
__device__ __forceinline__ float gem_dp_term(float gv, float dgv, float Sv, float pp, float HW) {
    if (!(gv > 0.0f)) return 0.0f;
    float l, gp;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(gv));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(gp) : "f"(pp * l));          // g^p
    return dgv * gv * (-(l * 0.6931471805599453f) / pp + Sv / (pp * HW * gp));
}

__global__ void __launch_bounds__(256)
gem_dp_kernel(const float* __restrict__ g, const float* __restrict__ dg, const float* __restrict__ S, const float* __restrict__ p,
              int p_stride, int N, int C, float HW, float* __restrict__ out, float* __restrict__ partial, unsigned* __restrict__ counter) {
    if (p_stride) {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        if (c >= C) return;
        const float pp = p[c];
        float acc = 0.0f;
        for (int n = 0; n < N; ++n) acc += gem_dp_term(g[(size_t)n * C + c], dg[(size_t)n * C + c], S[(size_t)n * C + c], pp, HW);
        out[c] = acc;
        return;
    }
    __shared__ float red[8];
    __shared__ bool s_last;
    const int n = blockIdx.x;
    const float pp = p[0];
    float acc = 0.0f;
    for (int c = threadIdx.x; c < C; c += 256) acc += gem_dp_term(g[(size_t)n * C + c], dg[(size_t)n * C + c], S[(size_t)n * C + c], pp, HW);
    acc = block_sum_256(acc, red);
    if (threadIdx.x == 0) {
        partial[n] = acc;
        __threadfence();
        s_last = atomicAdd(counter, 1u) == (unsigned)(N - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        float t = 0.0f;
        for (int i = threadIdx.x; i < N; i += 256) t += __ldcg(partial + i);
        t = block_sum_256(t, red);
        if (threadIdx.x == 0) {
            out[0] = t;
            *counter = 0u;
        }
    }
}

}  // namespace cir

using namespace cir;

extern "C" int cir_tuple_loss_workspace_bytes(int n_tuples, size_t* bytes) {
    CIR_REQUIRE(bytes && n_tuples > 0, CIR_ERR_INVALID_ARG, "cir_tuple_loss_workspace_bytes: bad arguments");
    *bytes = align_up((size_t)n_tuples * 4, 16) + 16;
    return CIR_OK;
}

extern "C" int cir_tuple_loss(const float* x, int64_t ldx, int n_tuples, int S, int D, const int32_t* label, int kind,
                              float margin, float eps, float* loss, float* grad, int64_t ldg, void* workspace,
                              size_t workspace_bytes, void* stream) {
    CIR_REQUIRE(x && label && loss, CIR_ERR_INVALID_ARG, "cir_tuple_loss: null pointer");
    CIR_REQUIRE(n_tuples > 0 && S >= 2 && D > 0 && ldx >= D && (!grad || ldg >= D), CIR_ERR_INVALID_ARG,
                "cir_tuple_loss: bad shape (tuples=%d S=%d D=%d)", n_tuples, S, D);
    CIR_REQUIRE(kind == CIR_LOSS_CONTRASTIVE || kind == CIR_LOSS_TRIPLET, CIR_ERR_INVALID_ARG, "cir_tuple_loss: unknown kind %d", kind);
    size_t need = 0;
    cir_tuple_loss_workspace_bytes(n_tuples, &need);
    CIR_REQUIRE(workspace && workspace_bytes >= need, CIR_ERR_WORKSPACE, "cir_tuple_loss: workspace %zu < %zu bytes", workspace_bytes, need);
    float* partial = static_cast<float*>(workspace);
    unsigned* counter = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + align_up((size_t)n_tuples * 4, 16));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CIR_CHECK_CUDA(cudaMemsetAsync(counter, 0, 4, st));
    tuple_loss_kernel<<<n_tuples, LOSS_THREADS, 0, st>>>(x, ldx, n_tuples, S, D, label, kind, margin, eps, loss, grad, ldg, partial,
                                                         counter);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}

extern "C" int cir_l2n_bwd_rows(const float* v, const float* gout, int64_t N, int C, float eps, float* out, float* unit_out,
                                void* stream) {
    CIR_REQUIRE(v && N >= 0 && C > 0 && (out || unit_out) && (!out || gout), CIR_ERR_INVALID_ARG, "cir_l2n_bwd_rows: bad arguments");
    if (N == 0) return CIR_OK;
    CIR_REQUIRE(N <= 0x7fffffffll, CIR_ERR_UNSUPPORTED, "cir_l2n_bwd_rows: too many rows");
    l2n_bwd_rows_kernel<<<(unsigned)N, 256, 0, static_cast<cudaStream_t>(stream)>>>(v, gout, N, C, eps, out, unit_out);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}

extern "C" int cir_colsum_rows(const float* X, int64_t N, int C, float* out, void* stream) {
    CIR_REQUIRE(X && out && N >= 0 && C > 0, CIR_ERR_INVALID_ARG, "cir_colsum_rows: bad arguments");
    colsum_rows_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, N, C, out);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}

extern "C" int cir_gem_dp(const float* g, const float* dg, const float* S, const float* p, int p_stride, int N, int C, int HW,
                          float* out, void* workspace, size_t workspace_bytes, void* stream) {
    CIR_REQUIRE(g && dg && S && p && out && N > 0 && C > 0 && HW > 0 && (p_stride == 0 || p_stride == 1), CIR_ERR_INVALID_ARG,
                "cir_gem_dp: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p_stride) {
        gem_dp_kernel<<<(C + 255) / 256, 256, 0, st>>>(g, dg, S, p, 1, N, C, (float)HW, out, nullptr, nullptr);
    } else {
        const size_t need = align_up((size_t)N * 4, 16) + 16;          // per-image partial sums + the arrival counter
        CIR_REQUIRE(workspace && workspace_bytes >= need, CIR_ERR_WORKSPACE, "cir_gem_dp: workspace %zu < %zu bytes (N * 4 + 32)",
                    workspace_bytes, need);
        float* partial = static_cast<float*>(workspace);
        unsigned* counter = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + align_up((size_t)N * 4, 16));
        CIR_CHECK_CUDA(cudaMemsetAsync(counter, 0, 4, st));
        gem_dp_kernel<<<N, 256, 0, st>>>(g, dg, S, p, 0, N, C, (float)HW, out, partial, counter);
    }
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}
