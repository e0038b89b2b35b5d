// Full descending argsort of long score rows: ranks = np.argsort(-scores, axis=0) of the reference
// (scripts/train_globalF.py:734, scripts/test.py:246-258) for databases far beyond one shared-memory tile.
//
// Segmented (one segment = one query's row) least-significant-digit radix sort, 4 passes of 8 bits over the 32-bit
// order-preserving image of the score, STABLE, so equal scores keep ascending database index -- the tie rule of every
// other kernel here.  Per pass three launches over a (chunk, row) grid:
//   radix_hist_kernel     256-bin histogram of every 4096-key chunk               (4 B read per key)
//   radix_scan_kernel     per row, exclusive prefix over (digit, chunk)           (tiny)
//   radix_scatter_kernel  re-reads the chunk in the same order, ranks every key inside its digit with warp MATCH.ANY
//                         votes (no atomics, deterministic), orders the chunk by digit in shared memory and writes every
//                         digit's run of keys + indices contiguously to its place of the pass
//                                                                                 (8 B read + 8 B written per key)
// Pass 0 builds keys from the fp32 scores on the fly, pass 3 writes the int32 ranks (and optionally the sorted scores)
// straight to the outputs.  Algorithmic traffic: 4 x 20 B per element (70 x 1M: 5.6 GB); HBM-bound, scattered writes.
#include "common.cuh"

namespace cir {

constexpr int RDX_THREADS = 256;
constexpr int RDX_WARPS = RDX_THREADS / 32;
constexpr int RDX_PER_LANE = 16;
constexpr int RDX_CHUNK = RDX_THREADS * RDX_PER_LANE;       // 4096 keys per block
constexpr int RDX_WARP_SPAN = 32 * RDX_PER_LANE;            // 512 consecutive keys per warp

// ascending radix order == descending score: complement of the order-preserving map
// (-0.0 + 0.0 = +0.0: the two zeros compare equal in np.argsort and must tie here as well)
__device__ __forceinline__ uint32_t score_to_radix(float s) { return ~float_to_ordered(s + 0.0f); }
__device__ __forceinline__ float radix_to_score(uint32_t k) { return ordered_to_float(~k); }

struct RadixArgs {
    const float* scores;      // pass 0 source
    long long ld;
    const uint32_t* in_keys;  // passes 1..3 source
    const int32_t* in_vals;
    uint32_t* out_keys;       // passes 0..2 destination
    int32_t* out_vals;
    int32_t* final_idx;       // pass 3 destination
    float* final_sorted;      // pass 3, optional
    int* hist;                // [Q][256][nblk]
    long long N;
    int nblk;
    int pass;
};

__device__ __forceinline__ uint32_t radix_load_key(const RadixArgs& a, int q, long long i) {
    return a.pass == 0 ? score_to_radix(__ldg(a.scores + (size_t)q * a.ld + i)) : __ldg(a.in_keys + (size_t)q * a.N + i);
}

__global__ void __launch_bounds__(RDX_THREADS) radix_hist_kernel(const RadixArgs a) {
    __shared__ int h[256];
    const int q = blockIdx.y, blk = blockIdx.x;
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blk * RDX_CHUNK;
    const int shift = a.pass * 8;
#pragma unroll 4
    for (int t = 0; t < RDX_PER_LANE; ++t) {
        const long long i = base + t * RDX_THREADS + threadIdx.x;
        if (i < a.N) atomicAdd(&h[(radix_load_key(a, q, i) >> shift) & 255u], 1);
    }
    __syncthreads();
    a.hist[((size_t)q * 256 + threadIdx.x) * a.nblk + blk] = h[threadIdx.x];
}

// exclusive prefix sum over the row's [256][nblk] counts in (digit, chunk) order, in place
__global__ void __launch_bounds__(1024) radix_scan_kernel(int* __restrict__ hist, int nblk) {
    __shared__ int warp_tot[32];
    int* h = hist + (size_t)blockIdx.x * 256 * nblk;
    const int total = 256 * nblk;
    const int per = (total + 1023) / 1024;
    const int lo = min(total, (int)threadIdx.x * per), hi = min(total, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += h[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += v;
        }
        warp_tot[lane] = winc - w;
    }
    __syncthreads();
    int run = warp_tot[warp] + inc - s;
    for (int i = lo; i < hi; ++i) {
        const int c = h[i];
        h[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(RDX_THREADS) radix_scatter_kernel(const RadixArgs a) {
    __shared__ int cnt[RDX_WARPS][256];
    __shared__ int gbase[256];              // where the chunk's keys of digit d start in the row (this pass's output)
    __shared__ int lbase[256 + 1];          // where they start inside the chunk once it is ordered by digit
    __shared__ uint32_t skey[RDX_CHUNK];    // the chunk ordered by digit: a digit's keys leave as ONE contiguous run
    __shared__ int32_t sval[RDX_CHUNK];
    const int q = blockIdx.y, blk = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < RDX_WARPS * 256; t += RDX_THREADS) (&cnt[0][0])[t] = 0;
    __syncthreads();
    const long long base = (long long)blk * RDX_CHUNK + (long long)warp * RDX_WARP_SPAN;
    const int shift = a.pass * 8;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t key[RDX_PER_LANE];
    int32_t val[RDX_PER_LANE];
    // warp `warp` owns keys [base, base + 512) in 16 batches of 32 consecutive keys: order inside the chunk = (warp, batch, lane)
#pragma unroll
    for (int t = 0; t < RDX_PER_LANE; ++t) {
        const long long i = base + t * 32 + lane;
        const bool valid = i < a.N;
        key[t] = valid ? radix_load_key(a, q, i) : 0u;
        val[t] = valid ? (a.pass == 0 ? (int32_t)i : __ldg(a.in_vals + (size_t)q * a.N + i)) : -1;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const uint32_t d = (key[t] >> shift) & 255u;
            const unsigned peers = __match_any_sync(vmask, d);
            if ((peers & lt) == 0) cnt[warp][d] += __popc(peers);       // one leader per digit, batches in sequence
        }
        __syncwarp();
    }
    __syncthreads();
    int tot = 0;
    {   // digit d = threadIdx.x: the warps' counts become their start inside the digit's run; tot = keys of the digit in the chunk
        const int d = threadIdx.x;
        gbase[d] = a.hist[((size_t)q * 256 + d) * a.nblk + blk];
#pragma unroll
        for (int w = 0; w < RDX_WARPS; ++w) {
            const int c = cnt[w][d];
            cnt[w][d] = tot;
            tot += c;
        }
    }
    {   // exclusive prefix of the digit totals over the 256 threads -> lbase
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        __shared__ int wsum[RDX_WARPS];
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int before = 0;
#pragma unroll
        for (int w = 0; w < RDX_WARPS; ++w) before += w < warp ? wsum[w] : 0;
        lbase[threadIdx.x] = before + inc - tot;
        if (threadIdx.x == RDX_THREADS - 1) lbase[256] = before + inc;
    }
    __syncthreads();
    // The scatter used to write every key straight to its place: the 32 lanes of a store hit ~27 different digit runs, i.e. 27
    // sectors for 128 bytes.  Ordered by digit in shared memory first, consecutive threads write consecutive places of a run.
#pragma unroll
    for (int t = 0; t < RDX_PER_LANE; ++t) {
        const bool valid = val[t] >= 0;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const uint32_t d = (key[t] >> shift) & 255u;
            const unsigned peers = __match_any_sync(vmask, d);
            const int l = lbase[d] + cnt[warp][d] + __popc(peers & lt);
            __syncwarp(vmask);
            if ((peers & lt) == 0) cnt[warp][d] += __popc(peers);
            skey[l] = key[t];
            sval[l] = val[t];
        }
        __syncwarp();
    }
    __syncthreads();
    const size_t row = (size_t)q * a.N;
    const int n_chunk = lbase[256];
    for (int l = threadIdx.x; l < n_chunk; l += RDX_THREADS) {
        const uint32_t k = skey[l];
        const uint32_t d = (k >> shift) & 255u;
        const size_t pos = row + (size_t)(gbase[d] + (l - lbase[d]));
        if (a.pass == 3) {
            a.final_idx[pos] = sval[l];
            if (a.final_sorted) a.final_sorted[pos] = radix_to_score(k);
        } else {
            a.out_keys[pos] = k;
            a.out_vals[pos] = sval[l];
        }
    }
}

size_t radix_sort_workspace_bytes(int Q, long long N) {
    const size_t nblk = (size_t)((N + RDX_CHUNK - 1) / RDX_CHUNK);
    return align_up((size_t)Q * (size_t)N * 4, 256) * 4 + align_up((size_t)Q * 256 * nblk * 4, 256);
}

int radix_sort_rows_desc(const float* scores, int Q, long long N, long long ld, int32_t* out_idx, float* out_sorted,
                         void* workspace, cudaStream_t stream) {
    const size_t arr = align_up((size_t)Q * (size_t)N * 4, 256);
    char* ws = static_cast<char*>(workspace);
    uint32_t* keys[2] = {reinterpret_cast<uint32_t*>(ws), reinterpret_cast<uint32_t*>(ws + arr)};
    int32_t* vals[2] = {reinterpret_cast<int32_t*>(ws + 2 * arr), reinterpret_cast<int32_t*>(ws + 3 * arr)};
    RadixArgs a{};
    a.scores = scores; a.ld = ld; a.N = N;
    a.nblk = (int)((N + RDX_CHUNK - 1) / RDX_CHUNK);
    a.hist = reinterpret_cast<int*>(ws + 4 * arr);
    a.final_idx = out_idx; a.final_sorted = out_sorted;
    const dim3 grid((unsigned)a.nblk, (unsigned)Q);
    for (int pass = 0; pass < 4; ++pass) {
        a.pass = pass;
        a.in_keys = keys[(pass + 1) & 1]; a.in_vals = vals[(pass + 1) & 1];       // written by the previous pass
        a.out_keys = keys[pass & 1]; a.out_vals = vals[pass & 1];
        radix_hist_kernel<<<grid, RDX_THREADS, 0, stream>>>(a);
        radix_scan_kernel<<<Q, 1024, 0, stream>>>(a.hist, a.nblk);
        radix_scatter_kernel<<<grid, RDX_THREADS, 0, stream>>>(a);
    }
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch(12);
    return CIR_OK;
}

}  // namespace cir
