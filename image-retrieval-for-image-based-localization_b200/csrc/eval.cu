// mAP / mP@k of ranked lists on the device: the consumer of the ranking step.
//
// Replaces compute_ap / compute_map (cirtorch/utils/evaluation/ParisOxfordEval.py:4-113) for
// ranks that already live in HBM (full N x Q rankings or top-k lists), so a 10k-query
// evaluation does not have to ship Q x N indices to the host.  One block per query:
// membership of every ranked index in the query's sorted `ok` / `junk` lists by binary
// search, block-wide prefix counts (junk entries ranked before a positive move it up,
// :86-96), trapezoidal AP (:24-36) and precision at kappa (:103-107) in fp64.
#include "common.cuh"

namespace cir {

constexpr int EVAL_THREADS = 256;
constexpr int EVAL_MAX_KAPPA = 8;

__device__ __forceinline__ bool sorted_contains(const int32_t* a, int n, int32_t v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int32_t x = __ldg(a + mid);
        if (x < v) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(a + lo) == v;
}

// exclusive block-wide prefix of two counters packed in one int (pos << 16 | junk would overflow: use two scans)
__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot, int tid, int& total) {
    const int lane = tid & 31, warp = tid >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < EVAL_THREADS / 32; ++w) {
        const int t = warp_tot[w];
        if (w < warp) base += t;
        tot += t;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(EVAL_THREADS)
eval_ap_kernel(const int32_t* __restrict__ ranks, int Q, long long R, long long ld, const int32_t* __restrict__ ok_off,
               const int32_t* __restrict__ ok_idx, const int32_t* __restrict__ junk_off,
               const int32_t* __restrict__ junk_idx, const int32_t* __restrict__ kappas, int nk,
               double* __restrict__ ap_out, double* __restrict__ prs_out) {
    __shared__ int warp_tot[EVAL_THREADS / 32];
    __shared__ double red[EVAL_THREADS / 32];
    __shared__ int red_i[EVAL_THREADS / 32][EVAL_MAX_KAPPA + 1];
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* ok = ok_idx + ok_off[q];
    const int nok = ok_off[q + 1] - ok_off[q];
    const int32_t* jk = junk_idx + junk_off[q];
    const int njk = junk_off[q + 1] - junk_off[q];
    if (nok == 0) {     // no positives: excluded from the averages (:69-73)
        if (tid == 0) ap_out[q] = nan("");
        if (tid < nk) prs_out[(size_t)q * nk + tid] = nan("");
        return;
    }
    const double step = 1.0 / (double)nok;
    double ap = 0.0;
    int cnt_le[EVAL_MAX_KAPPA];     // positives with 1-based adjusted position <= kappa_j
#pragma unroll
    for (int j = 0; j < EVAL_MAX_KAPPA; ++j) cnt_le[j] = 0;
    int max_pos1 = 0;
    int seen_pos = 0, seen_junk = 0;     // running totals of the previous tiles
    for (long long r0 = 0; r0 < R; r0 += EVAL_THREADS) {
        const long long r = r0 + tid;
        int is_pos = 0, is_junk = 0;
        if (r < R) {
            const int32_t v = __ldg(ranks + (size_t)q * ld + r);
            if (v >= 0) {
                is_pos = sorted_contains(ok, nok, v) ? 1 : 0;
                is_junk = (!is_pos && njk > 0 && sorted_contains(jk, njk, v)) ? 1 : 0;
            }
        }
        int tot_p, tot_j;
        const int before_p = block_excl_scan(is_pos, warp_tot, tid, tot_p);
        const int before_j = block_excl_scan(is_junk, warp_tot, tid, tot_j);
        if (is_pos) {
            const int j = seen_pos + before_p;                               // index among the positives
            const long long rank = r - (seen_junk + before_j);               // junk ranked before it moves it up
            const double p0 = rank == 0 ? 1.0 : (double)j / (double)rank;
            const double p1 = (double)(j + 1) / (double)(rank + 1);
            ap += (p0 + p1) * step / 2.0;
            const int pos1 = (int)(rank + 1);
            max_pos1 = max(max_pos1, pos1);
#pragma unroll
            for (int kk = 0; kk < EVAL_MAX_KAPPA; ++kk)
                if (kk < nk && pos1 <= __ldg(kappas + kk)) ++cnt_le[kk];
        }
        seen_pos += tot_p;
        seen_junk += tot_j;
    }
    // deterministic block reduction
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ap += __shfl_xor_sync(0xffffffffu, ap, o);
        max_pos1 = max(max_pos1, __shfl_xor_sync(0xffffffffu, max_pos1, o));
#pragma unroll
        for (int kk = 0; kk < EVAL_MAX_KAPPA; ++kk) cnt_le[kk] += __shfl_xor_sync(0xffffffffu, cnt_le[kk], o);
    }
    if (lane == 0) {
        red[warp] = ap;
        red_i[warp][EVAL_MAX_KAPPA] = max_pos1;
#pragma unroll
        for (int kk = 0; kk < EVAL_MAX_KAPPA; ++kk) red_i[warp][kk] = cnt_le[kk];
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;
        int mx = 0;
        for (int w = 0; w < EVAL_THREADS / 32; ++w) { a += red[w]; mx = max(mx, red_i[w][EVAL_MAX_KAPPA]); }
        ap_out[q] = a;
        for (int kk = 0; kk < nk; ++kk) {
            int c = 0;
            for (int w = 0; w < EVAL_THREADS / 32; ++w) c += red_i[w][kk];
            // kq = min(max(pos), kappa); c = positives with pos <= kappa == positives with pos <= kq (:104-106)
            const int kap = kappas[kk];
            const int kq = min(mx, kap);
            prs_out[(size_t)q * nk + kk] = kq > 0 ? (double)c / (double)kq : 0.0;
        }
    }
}

}  // namespace cir

using namespace cir;

extern "C" int cir_eval_ap(const int32_t* ranks, int Q, int64_t R, int64_t ld, const int32_t* ok_off, const int32_t* ok_idx,
                           const int32_t* junk_off, const int32_t* junk_idx, const int32_t* kappas, int nk, double* ap_out,
                           double* prs_out, void* stream) {
    CIR_REQUIRE(ranks && ok_off && ok_idx && junk_off && junk_idx && ap_out, CIR_ERR_INVALID_ARG, "cir_eval_ap: null pointer");
    CIR_REQUIRE(Q >= 0 && R >= 0 && ld >= R && nk >= 0 && nk <= EVAL_MAX_KAPPA && (nk == 0 || (kappas && prs_out)),
                CIR_ERR_INVALID_ARG, "cir_eval_ap: bad shape (at most %d kappas)", EVAL_MAX_KAPPA);
    if (Q == 0) return CIR_OK;
    eval_ap_kernel<<<Q, EVAL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(ranks, Q, R, ld, ok_off, ok_idx, junk_off, junk_idx,
                                                                             kappas, nk, ap_out, prs_out);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}
