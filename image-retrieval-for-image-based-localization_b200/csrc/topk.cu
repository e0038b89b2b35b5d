// Selection / sorting kernels of the search path.
//
//   topk_select_kernel   per query: gather the candidate lists the search kernel produced
//                        (or G already-sorted top-k lists from G shards / GPUs), bitonic-sort
//                        them in shared memory in batches of 4096 keys, emit the sorted top-k
//   rescore_kernel       exact fp32 re-scoring of a candidate list + final ordering
//   sort_*               full descending argsort of every score row
//                        (ranks = np.argsort(-scores, axis=0), scripts/train_globalF.py:734)
//
// Ordering everywhere: 64-bit keys (ordered score << 32 | ~index), sorted descending ==
// (score descending, index ascending).  Key 0 = empty slot.
#include "common.cuh"
#include "search.cuh"

#include <stdlib.h>

namespace cir {

constexpr int SEL_MAX_PEERS = 16;
constexpr int SEL_N = 4096;          // keys per shared-memory tile of the sorts / re-scoring
constexpr int SEL_CAP = 8192;        // keys the block-per-query selection holds at once
constexpr int SEL_THREADS = 512;

// in-place descending bitonic sort of buf[0, P) (P a power of two <= SEL_N); `base` is the
// global position of buf[0] (direction bits above the tile come from it), phases
// k = kfirst .. klast, strides start at min(k/2, P/2).
__device__ __forceinline__ void block_bitonic(unsigned long long* buf, int P, long long base, long long kfirst,
                                              long long klast, int tid, int nthreads) {
    for (long long k = kfirst; k <= klast; k <<= 1) {
        int j0 = (k >> 1) < (long long)(P >> 1) ? (int)(k >> 1) : (P >> 1);
        for (int j = j0; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += nthreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool desc = ((base + i) & k) == 0;
                const unsigned long long a = buf[i], b = buf[l];
                if ((a < b) == desc) { buf[i] = b; buf[l] = a; }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 2;
    while (p < v) p <<= 1;
    return p;
}

struct SelParams {
    // SRC 0: candidate lists of the search kernel
    const unsigned long long* lists;
    const int* counts;
    int Qpad, cap;
    // SRC 1: G sorted (score, index) lists [G][Q][kin]
    const float* in_scores;
    const int32_t* in_idx;
    int kin;
    long long g_stride;   // elements between consecutive lists of one query set
    int G, Q, k;
    float* out_scores;
    int32_t* out_idx;
    int out_ld;
    int32_t idx_offset;
    // fused exchange: the final lists are also stored into every peer GPU's exchange buffer [G][2][Q][k] (32-bit words:
    // scores then indices of rank g) through NVLink-mapped pointers -- an all-gather without a separate collective
    int n_peers, my_rank;
    void* peer[SEL_MAX_PEERS];
    // fused merge (flag_target != 0; block-per-query kernel, every block resident): behind the lists every exchange buffer
    // holds one arrival counter per query.  After its stores the block adds 1 to counter q of every peer (release, system
    // scope), waits until its own counter reaches flag_target (= uses of this buffer so far x n_peers: the counters are
    // never reset) and merges the n_peers lists of its query itself -- no cross-GPU barrier, no merge launch.  The merged
    // lists go to out_scores / out_idx (the local lists are then not written there).
    uint32_t flag_target;
};

__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int SEL_MAX_LISTS = 1024;   // lists per query handled with shared-memory prefix sums
constexpr int SEL_WARPS = SEL_THREADS / 32;
constexpr int SEL_MAX_KOUT = 1024;    // largest k the select / merge kernel emits
constexpr int SEL_SMEM_BYTES = SEL_CAP * 8 + SEL_MAX_KOUT * 8 + SEL_WARPS * 256 * 4 + (2 * SEL_MAX_LISTS + 1 + 8 + 3) * 4;

template <int SRC> __global__ void topk_select_kernel(const struct SelParams p);
template <int SRC>
static int launch_select(const struct SelParams& p, int Q, cudaStream_t stream);

// Bin search shared by the radix selects: lane l of ONE warp owns bins 255 - 8 l ... 248 - 8 l of a 256-bin histogram
// (descending); finds the bin where the running count from the top reaches krem.  Returns (on every lane) the bin, the
// number of keys in higher bins and the number of keys in the bin itself.
__device__ __forceinline__ void warp_find_bin(const int* hist, int krem, int lane, int* bin_out, int* before_out, int* cnt_out) {
    int c[8], tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i] = hist[255 - 8 * lane - i]; tot += c[i]; }
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    const int before = inc - tot;                     // keys in higher bins than this lane's
    const bool mine = before < krem && krem <= inc;
    int bin = 0, run = before, cb = 0;
    if (mine) {
        bin = 255 - 8 * lane;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (run + c[i] >= krem) { bin = 255 - 8 * lane - i; cb = c[i]; break; }
            run += c[i];
        }
    }
    const unsigned who = __ballot_sync(0xffffffffu, mine);
    const int src = who ? __ffs(who) - 1 : 0;          // who == 0 only if fewer than krem keys exist (callers exclude it)
    *bin_out = __shfl_sync(0xffffffffu, bin, src);
    *before_out = __shfl_sync(0xffffffffu, run, src);
    *cnt_out = __shfl_sync(0xffffffffu, cb, src);
}

__device__ __forceinline__ unsigned long long warp_and64(unsigned long long v) {
    const uint32_t lo = __reduce_and_sync(0xffffffffu, (uint32_t)v), hi = __reduce_and_sync(0xffffffffu, (uint32_t)(v >> 32));
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long warp_or64(unsigned long long v) {
    const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)v), hi = __reduce_or_sync(0xffffffffu, (uint32_t)(v >> 32));
    return ((unsigned long long)hi << 32) | lo;
}

// Keep the k largest of buf[0, fill) (fill > k, keys distinct): MSB-first radix select of a threshold key T with exactly k
// keys >= T, then the survivors are moved to buf[0, k) (unordered).  The bits all keys share are skipped (one AND / OR
// pass: candidate scores sit in a narrow range, so the top ~10 bits are common) and the select stops as soon as the
// boundary bin is taken whole -- typically 2 histogram passes instead of 8.
__device__ __forceinline__ void block_keep_topk(unsigned long long* buf, int fill, int k, int* hist /*[SEL_WARPS][256]*/,
                                                unsigned long long* stage /*[SEL_MAX_KOUT]*/, int* s_misc /*[8]*/, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    unsigned long long a = ~0ull, o = 0ull;
    for (int t = tid; t < fill; t += SEL_THREADS) { const unsigned long long key = buf[t]; a &= key; o |= key; }
    a = warp_and64(a); o = warp_or64(o);
    unsigned long long* s_ao = reinterpret_cast<unsigned long long*>(hist);      // [2][SEL_WARPS], before hist is used
    if (lane == 0) { s_ao[warp] = a; s_ao[SEL_WARPS + warp] = o; }
    __syncthreads();
    a = ~0ull; o = 0ull;
#pragma unroll
    for (int w = 0; w < SEL_WARPS; ++w) { a &= s_ao[w]; o |= s_ao[SEL_WARPS + w]; }
    __syncthreads();
    const unsigned long long diff = a ^ o;
    int bits = diff ? 64 - __clzll((long long)diff) : 0;         // undecided low bits
    unsigned long long prefix = bits >= 64 ? 0ull : (a >> bits) << bits;
    int krem = k;
    while (bits > 0) {
        const int w = bits < 8 ? bits : 8;
        const int shift = bits - w;
        const unsigned long long himask = bits >= 64 ? 0ull : ~((1ull << bits) - 1ull);
        for (int t = tid; t < SEL_WARPS * 256; t += SEL_THREADS) hist[t] = 0;
        __syncthreads();
        for (int t = tid; t < fill; t += SEL_THREADS) {
            const unsigned long long key = buf[t];
            if ((key & himask) == prefix) atomicAdd(&hist[warp * 256 + (int)((key >> shift) & (unsigned long long)((1 << w) - 1))], 1);
        }
        __syncthreads();
        if (tid < 256) {
            int c = 0;
#pragma unroll
            for (int w2 = 0; w2 < SEL_WARPS; ++w2) c += hist[w2 * 256 + tid];
            hist[tid] = c;                     // row 0 = block histogram
        }
        __syncthreads();
        if (warp == 0) {
            int bin, before, cb;
            warp_find_bin(hist, krem, lane, &bin, &before, &cb);
            if (lane == 0) { s_misc[0] = bin; s_misc[1] = krem - before; s_misc[3] = cb; }
        }
        __syncthreads();
        prefix |= (unsigned long long)s_misc[0] << shift;
        krem = s_misc[1];
        const int cb = s_misc[3];
        bits = shift;
        __syncthreads();
        if (cb == krem) break;                  // the whole boundary bin is needed: every key >= prefix survives
    }
    // survivors: everything above the boundary bin, then keys of the boundary bin until k are kept (the whole bin after an
    // early exit; with duplicated keys -- empty slots -- whichever come first)
    const unsigned long long thi = prefix | (bits >= 64 ? ~0ull : ((1ull << bits) - 1ull));
    if (tid == 0) { s_misc[2] = 0; }
    __syncthreads();
    for (int t = tid; t < fill; t += SEL_THREADS) {
        const unsigned long long key = buf[t];
        if (key > thi) stage[atomicAdd(&s_misc[2], 1)] = key;
    }
    __syncthreads();
    for (int t = tid; t < fill; t += SEL_THREADS) {
        const unsigned long long key = buf[t];
        if (key >= prefix && key <= thi) {
            const int pos = atomicAdd(&s_misc[2], 1);
            if (pos < k) stage[pos] = key;
        }
    }
    __syncthreads();
    const int kept = min(s_misc[2], k);
    for (int t = tid; t < k; t += SEL_THREADS) buf[t] = t < kept ? stage[t] : 0ull;
    __syncthreads();
}

// The same selection by ONE warp on its own slice of shared memory (no block barriers): survivors are compacted in place
// to buf[0, k).  Returns the number kept.
// ``a`` / ``o``: this lane's AND / OR over (a superset of) the keys it stored -- tracked while the keys are gathered, so no
// separate pass is needed to find the bits all keys share.
__device__ __forceinline__ int warp_keep_topk(unsigned long long* buf, int n, int k, int* hist /*[256]*/, int lane,
                                              unsigned long long a, unsigned long long o, unsigned long long* floor_key) {
    a = warp_and64(a); o = warp_or64(o);
    const unsigned long long diff = a ^ o;
    int bits = diff ? 64 - __clzll((long long)diff) : 0;
    unsigned long long prefix = bits >= 64 ? 0ull : (a >> bits) << bits;
    int krem = k;
    while (bits > 0) {
        const int w = bits < 8 ? bits : 8;
        const int shift = bits - w;
        const unsigned long long himask = bits >= 64 ? 0ull : ~((1ull << bits) - 1ull);
        const unsigned long long dmask = (unsigned long long)((1 << w) - 1);
#pragma unroll
        for (int t = 0; t < 8; ++t) hist[t * 32 + lane] = 0;
        __syncwarp();
        for (int t = lane; t < n; t += 32) {
            const unsigned long long key = buf[t];
            if ((key & himask) == prefix) atomicAdd(&hist[(int)((key >> shift) & dmask)], 1);
        }
        __syncwarp();
        int bin, before, cb;
        warp_find_bin(hist, krem, lane, &bin, &before, &cb);
        __syncwarp();
        prefix |= (unsigned long long)bin << shift;
        krem -= before;
        bits = shift;
        if (cb == krem) break;
    }
    // in-place compaction: everything above the boundary bin + the first krem keys of the bin (all of it after an early exit)
    const unsigned long long thi = prefix | (bits >= 64 ? ~0ull : ((1ull << bits) - 1ull));
    int base = 0, taken = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int t0 = 0; t0 < n; t0 += 32) {
        const int t = t0 + lane;
        const unsigned long long key = t < n ? buf[t] : 0ull;
        const bool edge = t < n && key >= prefix && key <= thi;
        const unsigned bal_e = __ballot_sync(0xffffffffu, edge);
        const bool keep = t < n && (key > thi || (edge && taken + __popc(bal_e & lt) < krem));
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = base + __popc(bal & lt);
            if (pos < k) buf[pos] = key;             // pos <= t: never ahead of the keys still to be read
        }
        base += __popc(bal);
        taken += __popc(bal_e);
        __syncwarp();
    }
    *floor_key = prefix;          // k keys >= prefix are kept: anything smaller can never enter the top-k again
    return min(base, k);
}

template <int SRC>
__global__ void __launch_bounds__(SEL_THREADS) topk_select_kernel(const SelParams p) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(sel_smem);          // [SEL_CAP]
    unsigned long long* stage = buf + SEL_CAP;                                          // [SEL_MAX_KOUT]
    int* hist = reinterpret_cast<int*>(stage + SEL_MAX_KOUT);                            // [SEL_WARPS][256]
    int* s_cnt = hist + SEL_WARPS * 256;      // [SEL_MAX_LISTS] list sizes of the current window of lists
    int* s_off = s_cnt + SEL_MAX_LISTS;       // [SEL_MAX_LISTS + 1] exclusive prefix sums
    int* s_misc = s_off + SEL_MAX_LISTS + 1;  // [8]
    const int q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int fill = 0;
    for (int g0 = 0; g0 < p.G;) {
        // sizes of up to SEL_MAX_LISTS lists starting at g0, loaded in parallel
        const int win = min(SEL_MAX_LISTS, p.G - g0);
        for (int t = tid; t < win; t += SEL_THREADS)
            s_cnt[t] = SRC == 0 ? __ldg(p.counts + (size_t)(g0 + t) * p.Qpad + q) : p.kin;
        __syncthreads();
        if (warp == 0) {   // exclusive scan by one warp
            int run = 0;
            for (int base = 0; base < win; base += 32) {
                const int c = base + lane < win ? s_cnt[base + lane] : 0;
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (base + lane < win) s_off[base + lane] = run + inc - c;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) s_off[win] = run;
        }
        __syncthreads();
        int done = 0;     // lists of this window already consumed
        while (done < win) {
            // Take as many lists as fit behind the `fill` survivors of the previous round: the largest `take` with
            // s_off[take] - start <= room, by bisection (every thread used to walk the offsets one by one: 148 lists per
            // query at 70 queries x 1M -- 15 % of the kernel's samples).  A list always fits: fill <= k and k + cap <= SEL_CAP.
            const int room = SEL_CAP - fill;
            const int start = s_off[done];
            int take = done, hi = win;
            while (take < hi) {
                const int mid = (take + hi + 1) >> 1;
                if (s_off[mid] - start <= room) take = mid; else hi = mid - 1;
            }
            // Gather: warp w copies lists done + w, done + w + 16, ...  The first 64 entries of EIGHT lists are loaded before
            // any is stored (16 independent global loads in flight per lane); the plain loop had one dependent L2 round
            // trip per list and query (23 % of the samples of 70 x 1M waited there), a flat gather with a bisection per
            // entry traded that for 8 x more instructions.  Longer lists finish in a remainder loop.
            constexpr int LPW = 8;
            for (int gb = done + warp; gb < take; gb += SEL_WARPS * LPW) {
                unsigned long long v0[LPW], v1[LPW];
#pragma unroll
                for (int u = 0; u < LPW; ++u) {
                    const int g = gb + u * SEL_WARPS;
                    v0[u] = v1[u] = 0ull;
                    if (g < take) {
                        const int c = s_cnt[g];
                        if (SRC == 0) {
                            const unsigned long long* L = p.lists + ((size_t)(g0 + g) * p.Qpad + q) * p.cap;
                            if (lane < c) v0[u] = __ldcg(L + lane);
                            if (lane + 32 < c) v1[u] = __ldcg(L + lane + 32);
                        } else {
                            const size_t o = (size_t)(g0 + g) * (size_t)p.g_stride + (size_t)q * p.kin;
                            if (lane < c) {
                                const int32_t ix = __ldg(p.in_idx + o + lane);
                                v0[u] = ix < 0 ? 0ull : make_key(__ldg(p.in_scores + o + lane), (uint32_t)ix);
                            }
                            if (lane + 32 < c) {
                                const int32_t ix = __ldg(p.in_idx + o + lane + 32);
                                v1[u] = ix < 0 ? 0ull : make_key(__ldg(p.in_scores + o + lane + 32), (uint32_t)ix);
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < LPW; ++u) {
                    const int g = gb + u * SEL_WARPS;
                    if (g < take) {
                        const int c = s_cnt[g];
                        const int dst = fill + s_off[g] - start;
                        if (lane < c) buf[dst + lane] = SRC == 0 ? raw_to_key(v0[u]) : v0[u];
                        if (lane + 32 < c) buf[dst + lane + 32] = SRC == 0 ? raw_to_key(v1[u]) : v1[u];
                        if (c > 64) {
                            if (SRC == 0) {
                                const unsigned long long* L = p.lists + ((size_t)(g0 + g) * p.Qpad + q) * p.cap;
                                for (int t = lane + 64; t < c; t += 32) buf[dst + t] = raw_to_key(__ldcg(L + t));
                            } else {
                                const size_t o = (size_t)(g0 + g) * (size_t)p.g_stride + (size_t)q * p.kin;
                                for (int t = lane + 64; t < c; t += 32) {
                                    const int32_t ix = __ldg(p.in_idx + o + t);
                                    buf[dst + t] = ix < 0 ? 0ull : make_key(__ldg(p.in_scores + o + t), (uint32_t)ix);
                                }
                            }
                        }
                    }
                }
            }
            fill += s_off[take] - start;
            done = take;
            __syncthreads();
            if (fill > p.k) {
                block_keep_topk(buf, fill, p.k, hist, stage, s_misc, tid);
                fill = p.k;
            }
        }
        g0 += win;
        __syncthreads();
    }
    // Final ordering of the <= k survivors: rank sort.  The keys are distinct (the index is part of the key), so the rank of a
    // key is the number of larger keys: `fill` broadcast reads per thread and two barriers, where the bitonic network over a
    // few hundred keys cost log^2 stages of block barriers (28 for k = 100) -- a large part of this kernel's time when it
    // serves a handful of queries (70 x 1M: 38 us of selection on a 0.6 ms scan, the same on the 78 us scan of an 8-GPU shard).
    const bool fused_merge = SRC == 0 && p.flag_target != 0u;
    for (int round = 0; round < (fused_merge ? 2 : 1); ++round) {
        for (int t = tid; t < p.k; t += SEL_THREADS) stage[t] = 0ull;       // empty slots (fewer than k candidates) sort last
        __syncthreads();
        for (int t = tid; t < fill; t += SEL_THREADS) {
            const unsigned long long key = buf[t];
            if (key != 0ull) {
                int rank = 0;
                for (int u = 0; u < fill; ++u) { const unsigned long long o = buf[u]; rank += (o > key || (o == key && u < t)) ? 1 : 0; }
                stage[rank] = key;
            }
        }
        __syncthreads();
        // round 0: this rank's list (indices + idx_offset); round 1 of a fused merge: the merged list (indices already global)
        const int32_t add = round == 0 ? p.idx_offset : 0;
        const bool to_out = p.out_scores && (!fused_merge || round == 1);
        for (int t = tid; t < p.k; t += SEL_THREADS) {
            const unsigned long long key = stage[t];
            const bool ok = key != 0ull;
            const float sc = ok ? key_score(key) : -INFINITY;
            const int32_t ix = ok ? (int32_t)key_index(key) + add : -1;
            if (to_out) {
                p.out_scores[(size_t)q * p.out_ld + t] = sc;
                p.out_idx[(size_t)q * p.out_ld + t] = ix;
            }
            if (round == 0) {
                for (int g = 0; g < p.n_peers; ++g) {
                    // peer buffer layout [G][2][Q][k]; this rank fills slot my_rank of every peer (plain stores over NVLink)
                    uint32_t* base = static_cast<uint32_t*>(p.peer[g]) + (size_t)p.my_rank * 2 * p.Q * p.k;
                    base[(size_t)q * p.k + t] = __float_as_uint(sc);
                    base[(size_t)p.Q * p.k + (size_t)q * p.k + t] = (uint32_t)ix;
                }
            }
        }
        if (!fused_merge || round == 1) break;
        // ---- arrival counters: signal every peer, wait for all of them, gather the n_peers lists of this query
        const size_t list_words = (size_t)p.n_peers * 2 * p.Q * p.k;
        __syncthreads();                                       // every store of this block is issued ...
        if (tid < p.n_peers) {
            __threadfence_system();                            // ... and ordered before the signal (cumulative over the barrier)
            uint32_t* f = static_cast<uint32_t*>(p.peer[tid]) + list_words + q;
            asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
        }
        if (tid == 0) {
            const uint32_t* f = static_cast<const uint32_t*>(p.peer[p.my_rank]) + list_words + q;
            const uint64_t t0 = global_timer_ns();
            for (;;) {
                uint32_t v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if ((int32_t)(v - p.flag_target) >= 0) break;
                if (global_timer_ns() - t0 > 60000000000ull) __trap();     // a peer never arrived (60 s): fail loudly, do not hang
                __nanosleep(64);
            }
        }
        __syncthreads();
        const uint32_t* mine = static_cast<const uint32_t*>(p.peer[p.my_rank]);
        fill = p.n_peers * p.k;
        for (int t = tid; t < fill; t += SEL_THREADS) {
            const int g = t / p.k, j = t - g * p.k;
            const uint32_t* L = mine + (size_t)g * 2 * p.Q * p.k + (size_t)q * p.k + j;
            const uint32_t sc = __ldcg(L);                                   // written by peers: L2, never a stale L1 line
            const int32_t ix = (int32_t)__ldcg(L + (size_t)p.Q * p.k);
            buf[t] = ix < 0 ? 0ull : make_key(__uint_as_float(sc), (uint32_t)ix);
        }
        __syncthreads();
        if (fill > p.k) {
            block_keep_topk(buf, fill, p.k, hist, stage, s_misc, tid);
            fill = p.k;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Warp-per-query variant for large query batches: 8 queries per block, one warp each, no block barriers.  With
// thousands of queries in flight the serial latency of one warp (list loads, ~2 histogram passes, a 128-key bitonic
// sort) is hidden by the other warps; the block-per-query kernel above needs ~70 block barriers per query and ran
// 10k queries in 0.7 ms, this one in well under 0.1 ms.  Window = WN keys of shared memory per warp (>= k + one list).
// ---------------------------------------------------------------------------------------
constexpr int WSEL_WARPS = 8;
constexpr int WSEL_MAX_WN = 2048;

template <int SRC>
__global__ void __launch_bounds__(WSEL_WARPS * 32) topk_select_warp_kernel(const SelParams p, const int WN) {
    extern __shared__ __align__(16) unsigned char wsel_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(wsel_smem) + (size_t)warp * WN;
    int* hist = reinterpret_cast<int*>(wsel_smem + (size_t)WSEL_WARPS * WN * 8) + warp * 256;
    const int q = blockIdx.x * WSEL_WARPS + warp;
    if (q >= p.Q) return;                       // warps are independent: no block-level barrier below
    const unsigned lt = (1u << lane) - 1u;
    int fill = 0;
    unsigned long long and_acc = ~0ull, or_acc = 0ull;      // bits shared by every key gathered so far (lane-local)
    unsigned long long floor_key = 0ull;                    // after a selection: k keys >= floor_key are in the window
    // Gather in batches of <= 256 entries (8 per lane, all loads of a batch in flight together); the loads of batch
    // i + 1 are issued BEFORE batch i is written to shared memory, so one warp pays one memory round trip per batch
    // instead of one per 32 entries.
    constexpr int BPL = 8;                                  // entries per lane per batch
    unsigned long long pend[BPL];
    int pend_n = 0;
    auto commit = [&]() {                                   // pending batch -> window (making room first)
        if (pend_n == 0) return;
        if (fill + pend_n > WN) {
            __syncwarp();
            fill = warp_keep_topk(buf, fill, p.k, hist, lane, and_acc, or_acc, &floor_key);
        }
        if (SRC == 0 && floor_key == 0ull) {                // first window: everything is stored
#pragma unroll
            for (int u = 0; u < BPL; ++u) {
                const int t = u * 32 + lane;
                if (t < pend_n) {
                    const unsigned long long key = raw_to_key(pend[u]);
                    buf[fill + t] = key;
                    and_acc &= key; or_acc |= key;
                }
            }
            fill += pend_n;
        } else {
            // empty slots (key 0) and, once a selection has run, keys below its k-th best are dropped here: like a
            // streaming top-k the window then fills ~k ln(n) slowly and one more selection at the end is enough
#pragma unroll
            for (int u = 0; u < BPL; ++u) {
                const int t = u * 32 + lane;
                unsigned long long key = t < pend_n ? pend[u] : 0ull;
                if (SRC == 0 && key != 0ull) key = raw_to_key(key);
                const bool keep = key != 0ull && key >= floor_key;
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    buf[fill + __popc(bal & lt)] = key;
                    and_acc &= key; or_acc |= key;
                }
                fill += __popc(bal);
            }
        }
        pend_n = 0;
    };
    for (int g0 = 0; g0 < p.G; g0 += 32) {
        const int ng = min(32, p.G - g0);
        int myc = 0;
        if (lane < ng) myc = SRC == 0 ? __ldg(p.counts + (size_t)(g0 + lane) * p.Qpad + q) : p.kin;
        for (int i = 0; i < ng; ++i) {
            const int c = __shfl_sync(0xffffffffu, myc, i);
            for (int b0 = 0; b0 < c; b0 += BPL * 32) {
                const int nb = min(BPL * 32, c - b0);
                unsigned long long cur[BPL];
                if (SRC == 0) {
                    const unsigned long long* L = p.lists + ((size_t)(g0 + i) * p.Qpad + q) * p.cap + b0;
#pragma unroll
                    for (int u = 0; u < BPL; ++u) {
                        const int t = u * 32 + lane;
                        cur[u] = t < nb ? __ldcg(L + t) : 0ull;
                    }
                } else {
                    const size_t o = (size_t)(g0 + i) * (size_t)p.g_stride + (size_t)q * p.kin + b0;
                    int32_t ix[BPL];
                    float sc[BPL];
#pragma unroll
                    for (int u = 0; u < BPL; ++u) {
                        const int t = u * 32 + lane;
                        ix[u] = t < nb ? __ldg(p.in_idx + o + t) : -1;
                        sc[u] = t < nb ? __ldg(p.in_scores + o + t) : 0.0f;
                    }
#pragma unroll
                    for (int u = 0; u < BPL; ++u) cur[u] = ix[u] < 0 ? 0ull : make_key(sc[u], (uint32_t)ix[u]);
                }
                commit();                                   // the previous batch, while this one is in flight
#pragma unroll
                for (int u = 0; u < BPL; ++u) pend[u] = cur[u];
                pend_n = nb;
            }
        }
    }
    commit();
    __syncwarp();
    if (fill > p.k) fill = warp_keep_topk(buf, fill, p.k, hist, lane, and_acc, or_acc, &floor_key);
    // final ordering of the <= k survivors
    const int P = next_pow2(fill);
    for (int t = fill + lane; t < P; t += 32) buf[t] = 0ull;
    __syncwarp();
    for (int kk = 2; kk <= P; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool desc = (i & kk) == 0;
                const unsigned long long a = buf[i], b = buf[l];
                if ((a < b) == desc) { buf[i] = b; buf[l] = a; }
            }
            __syncwarp();
        }
    }
    for (int t = lane; t < p.k; t += 32) {
        const unsigned long long key = t < fill ? buf[t] : 0ull;
        const bool ok = key != 0ull;
        const float sc = ok ? key_score(key) : -INFINITY;
        const int32_t ix = ok ? (int32_t)key_index(key) + p.idx_offset : -1;
        if (p.out_scores) {
            p.out_scores[(size_t)q * p.out_ld + t] = sc;
            p.out_idx[(size_t)q * p.out_ld + t] = ix;
        }
        for (int g = 0; g < p.n_peers; ++g) {
            uint32_t* base = static_cast<uint32_t*>(p.peer[g]) + (size_t)p.my_rank * 2 * p.Q * p.k;
            base[(size_t)q * p.k + t] = __float_as_uint(sc);
            base[(size_t)p.Q * p.k + (size_t)q * p.k + t] = (uint32_t)ix;
        }
    }
}

// ---------------------------------------------------------------------------------------
// k-th largest value of every row of a dense [Q, n] score block (n <= KTH_MAX_N): the warm
// start tau0 of the fused top-k.  One block per row, row held in shared memory as ordered
// u32, up to 4 x 8-bit radix-select passes with per-warp histograms (stops once the boundary bin is taken whole).
// ---------------------------------------------------------------------------------------
template <int KTH_THREADS>
__global__ void __launch_bounds__(KTH_THREADS)
row_kth_largest_kernel(const float* __restrict__ scores, int n, long long ld, int k, float* __restrict__ out) {
    constexpr int KTH_WARPS = KTH_THREADS / 32;
    extern __shared__ __align__(16) unsigned char kth_smem[];
    uint32_t* vals = reinterpret_cast<uint32_t*>(kth_smem);                 // [n]
    int* hist = reinterpret_cast<int*>(kth_smem + (size_t)n * 4);           // [KTH_WARPS][256]
    __shared__ uint32_t s_prefix;
    __shared__ int s_krem, s_done;
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
    if (n < k) {
        if (tid == 0) out[q] = -INFINITY;
        return;
    }
    const float* row = scores + (size_t)q * ld;
    for (int t = tid; t < n; t += KTH_THREADS) vals[t] = float_to_ordered(__ldcs(row + t));
    if (tid == 0) { s_prefix = 0u; s_krem = k; s_done = 0; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int t = tid; t < KTH_WARPS * 256; t += KTH_THREADS) hist[t] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int t = tid; t < n; t += KTH_THREADS) {
            const uint32_t v = vals[t];
            if ((v & himask) == prefix) atomicAdd(&hist[warp * 256 + ((v >> shift) & 255u)], 1);
        }
        __syncthreads();
        if (tid < 256) {
            int c = 0;
#pragma unroll 8
            for (int w = 0; w < KTH_WARPS; ++w) c += hist[w * 256 + tid];
            hist[tid] = c;      // row 0 now holds the block histogram
        }
        __syncthreads();
        if (warp == 0) {       // (was a serial walk over the 256 bins by one thread: 4 us per pass)
            int bin, before, cb;
            warp_find_bin(hist, s_krem, tid, &bin, &before, &cb);
            __syncwarp();
            if (tid == 0) {
                s_krem -= before;
                s_prefix = prefix | ((uint32_t)bin << shift);
                s_done = (cb == s_krem) ? 1 : 0;      // the whole boundary bin is needed: exactly k values >= prefix
            }
        }
        __syncthreads();
        if (s_done) break;
    }
    if (tid == 0) {
        const float t = ordered_to_float(s_prefix);     // low bits zero after an early exit: still a lower bound
        out[q] = t == t ? t : -INFINITY;
    }
}

int launch_row_kth_largest(const float* scores, int Q, int n, long long ld, int k, float* out, cudaStream_t stream) {
    const bool small = n <= 4096;                       // 256 threads (8 blocks / SM) for short rows
    const size_t smem = (size_t)n * 4 + (small ? 8 : 32) * 256 * 4;
    const DeviceInfo& dev = device_info();
    CIR_REQUIRE(n <= KTH_MAX_N && (int)smem + 1024 <= dev.max_smem_optin, CIR_ERR_UNSUPPORTED,
                "row_kth_largest: n=%d too large for shared memory", n);
    static thread_local int attr_dev = -1;
    if (attr_dev != dev.device) {
        CIR_CHECK_CUDA(cudaFuncSetAttribute(row_kth_largest_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            KTH_MAX_N * 4 + 32 * 256 * 4));
        attr_dev = dev.device;
    }
    if (small) row_kth_largest_kernel<256><<<Q, 256, smem, stream>>>(scores, n, ld, k, out);
    else row_kth_largest_kernel<1024><<<Q, 1024, smem, stream>>>(scores, n, ld, k, out);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}

template <int SRC>
static int launch_select(const SelParams& p, int Q, cudaStream_t stream) {
    static thread_local int attr_dev = -1;
    const DeviceInfo& dev = device_info();
    if (attr_dev != dev.device) {
        CIR_CHECK_CUDA(cudaFuncSetAttribute(topk_select_kernel<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_SMEM_BYTES));
        CIR_CHECK_CUDA(cudaFuncSetAttribute(topk_select_warp_kernel<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            WSEL_WARPS * (WSEL_MAX_WN * 8 + 1024)));
        attr_dev = dev.device;
    }
    // large batches: one warp per query (window >= k + the longest list); few queries: one block per query
    const int longest = SRC == 0 ? p.cap : p.kin;
    int WN = 512;
    while (WN < p.k + longest) WN <<= 1;
    static const char* dbg = getenv("CIR_DEBUG_SELECT");              // experiments: "block" / "warp"
    const bool want_warp = p.flag_target ? false : (dbg ? dbg[0] == 'w' : Q >= 4 * dev.num_sms);
    if (WN <= WSEL_MAX_WN && want_warp) {
        const size_t smem = (size_t)WSEL_WARPS * ((size_t)WN * 8 + 1024);
        topk_select_warp_kernel<SRC><<<(Q + WSEL_WARPS - 1) / WSEL_WARPS, WSEL_WARPS * 32, smem, stream>>>(p, WN);
    } else {
        topk_select_kernel<SRC><<<Q, SEL_THREADS, SEL_SMEM_BYTES, stream>>>(p);
    }
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}

int launch_topk_select_lists(const unsigned long long* lists, const int* counts, int S, int Qpad, int cap, int Q, int k,
                             float* out_scores, int32_t* out_idx, int out_ld, int32_t idx_offset, cudaStream_t stream,
                             void* const* peers, int n_peers, int my_rank, uint32_t flag_target) {
    CIR_REQUIRE(k + cap <= SEL_CAP, CIR_ERR_UNSUPPORTED, "topk select: k + cap = %d exceeds %d", k + cap, SEL_CAP);
    if (flag_target) {
        // every block spins on its peers: all Q blocks must be resident at once (one per SM is always possible)
        CIR_REQUIRE(n_peers >= 1 && Q <= device_info().num_sms && n_peers * k <= SEL_CAP && out_scores && out_idx, CIR_ERR_UNSUPPORTED,
                    "topk select: the fused merge needs Q <= %d queries, n_peers * k <= %d and an output", device_info().num_sms, SEL_CAP);
    }
    SelParams p{};
    p.lists = lists; p.counts = counts; p.Qpad = Qpad; p.cap = cap;
    p.G = S; p.Q = Q; p.k = k;
    p.out_scores = out_scores; p.out_idx = out_idx; p.out_ld = out_ld; p.idx_offset = idx_offset;
    CIR_REQUIRE(n_peers >= 0 && n_peers <= SEL_MAX_PEERS, CIR_ERR_UNSUPPORTED, "topk select: at most %d peers", SEL_MAX_PEERS);
    p.n_peers = n_peers; p.my_rank = my_rank;
    for (int g = 0; g < n_peers; ++g) p.peer[g] = peers[g];
    p.flag_target = flag_target;
    return launch_select<0>(p, Q, stream);
}

// ---------------------------------------------------------------------------------------
// exact fp32 re-scoring
// ---------------------------------------------------------------------------------------
constexpr int RESCORE_THREADS = 256;

__global__ void __launch_bounds__(RESCORE_THREADS)
rescore_kernel(const float* __restrict__ q32, const float* __restrict__ db32, long long N, int D,
               const int32_t* __restrict__ cand, int Kc, int Kp, int32_t idx_offset, float* __restrict__ out_scores,
               int32_t* __restrict__ out_idx, int k_out) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(rs_smem);      // [Kp]
    float* qv = reinterpret_cast<float*>(rs_smem + (size_t)Kp * 8);                 // [D]
    const int q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int d = tid; d < D; d += RESCORE_THREADS) qv[d] = __ldg(q32 + (size_t)q * D + d);
    for (int t = Kc + tid; t < Kp; t += RESCORE_THREADS) keys[t] = 0ull;
    __syncthreads();
    const bool vec = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(db32) & 15) == 0);
    for (int c = warp; c < Kc; c += RESCORE_THREADS / 32) {
        const int32_t ix = __ldg(cand + (size_t)q * Kc + c);
        const long long row = (long long)ix - idx_offset;
        unsigned long long key = 0ull;
        if (ix >= 0 && row >= 0 && row < N) {
            const float* v = db32 + (size_t)row * D;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (vec) {
                // 8 x 512 B of the row in flight per warp (the gather is latency-bound otherwise)
                int d = lane * 4;
                for (; d + 7 * 128 < D; d += 8 * 128) {
                    float4 x[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) x[u] = ld_stream_f4(reinterpret_cast<const float4*>(v + d + u * 128));
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 y = *reinterpret_cast<const float4*>(qv + d + u * 128);
                        a0 = fmaf(x[u].x, y.x, a0); a1 = fmaf(x[u].y, y.y, a1);
                        a2 = fmaf(x[u].z, y.z, a2); a3 = fmaf(x[u].w, y.w, a3);
                    }
                }
                for (; d < D; d += 128) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(v + d));
                    const float4 y = *reinterpret_cast<const float4*>(qv + d);
                    a0 = fmaf(x.x, y.x, a0); a1 = fmaf(x.y, y.y, a1);
                    a2 = fmaf(x.z, y.z, a2); a3 = fmaf(x.w, y.w, a3);
                }
            } else {
                for (int d = lane; d < D; d += 32) a0 = fmaf(__ldg(v + d), qv[d], a0);
            }
            const float dot = warp_sum((a0 + a1) + (a2 + a3));
            key = make_key(dot, (uint32_t)ix);
        }
        if (lane == 0) keys[c] = key;
    }
    __syncthreads();
    block_bitonic(keys, Kp, 0, 2, Kp, tid, RESCORE_THREADS);
    for (int t = tid; t < k_out; t += RESCORE_THREADS) {
        const unsigned long long key = t < Kp ? keys[t] : 0ull;
        const bool ok = key != 0ull;
        out_scores[(size_t)q * k_out + t] = ok ? key_score(key) : -INFINITY;
        out_idx[(size_t)q * k_out + t] = ok ? (int32_t)key_index(key) : -1;
    }
}

// ---------------------------------------------------------------------------------------
// full argsort of score rows
// ---------------------------------------------------------------------------------------
__global__ void sort_build_keys(const float* __restrict__ scores, long long N, long long ld, long long Npad,
                                unsigned long long* __restrict__ keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (i >= Npad) return;
    // + 0.0f: -0.0 and +0.0 compare equal in np.argsort and must tie (index order) here as well
    keys[(size_t)q * Npad + i] = i < N ? make_key(__ldg(scores + (size_t)q * ld + i) + 0.0f, (uint32_t)i) : 0ull;
}

// sorts / merges one 4096-key tile in shared memory: phases kfirst..klast
__global__ void __launch_bounds__(SEL_THREADS)
sort_local_kernel(unsigned long long* __restrict__ keys, long long Npad, long long kfirst, long long klast) {
    __shared__ unsigned long long buf[SEL_N];
    const long long base = (long long)blockIdx.x * SEL_N;
    unsigned long long* g = keys + (size_t)blockIdx.y * Npad + base;
    for (int t = threadIdx.x; t < SEL_N; t += SEL_THREADS) buf[t] = g[t];
    __syncthreads();
    block_bitonic(buf, SEL_N, base, kfirst, klast, threadIdx.x, SEL_THREADS);
    for (int t = threadIdx.x; t < SEL_N; t += SEL_THREADS) g[t] = buf[t];
}

// one compare-exchange stage with stride j >= 4096 in global memory
__global__ void sort_global_stage(unsigned long long* __restrict__ keys, long long Npad, long long k, long long j) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (Npad >> 1)) return;
    unsigned long long* g = keys + (size_t)blockIdx.y * Npad;
    const long long i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const long long l = i | j;
    const bool desc = (i & k) == 0;
    const unsigned long long a = g[i], b = g[l];
    if ((a < b) == desc) { g[i] = b; g[l] = a; }
}

__global__ void sort_extract(const unsigned long long* __restrict__ keys, long long N, long long Npad,
                             int32_t* __restrict__ out_idx, float* __restrict__ out_sorted) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (i >= N) return;
    const unsigned long long key = keys[(size_t)q * Npad + i];
    out_idx[(size_t)q * N + i] = (int32_t)key_index(key);
    if (out_sorted) out_sorted[(size_t)q * N + i] = key_score(key);
}

static long long sort_npad(long long N) {
    long long p = SEL_N;
    while (p < N) p <<= 1;
    return p;
}

// rows longer than this go to the segmented radix sort (radix.cu): the bitonic network needs log^2 global stages
// (70 x 1M: 14.0 ms; torch.sort 7.4 ms), the radix sort 4 passes
constexpr long long SORT_RADIX_MIN_N = 4 * SEL_N;
size_t radix_sort_workspace_bytes(int Q, long long N);
int radix_sort_rows_desc(const float* scores, int Q, long long N, long long ld, int32_t* out_idx, float* out_sorted,
                         void* workspace, cudaStream_t stream);

}  // namespace cir

using namespace cir;

extern "C" int cir_topk_merge(const float* scores, const int32_t* idx, int G, int Q, int k, int64_t g_stride,
                              float* out_scores, int32_t* out_idx, int k_out, void* stream) {
    CIR_REQUIRE(scores && idx && out_scores && out_idx, CIR_ERR_INVALID_ARG, "cir_topk_merge: null pointer");
    CIR_REQUIRE(G >= 1 && Q >= 0 && k >= 1 && k_out >= 1, CIR_ERR_INVALID_ARG, "cir_topk_merge: bad shape");
    CIR_REQUIRE(k_out + k <= SEL_CAP && k_out <= SEL_MAX_KOUT, CIR_ERR_UNSUPPORTED,
                "cir_topk_merge: k_out + k = %d exceeds %d (or k_out > %d)", k_out + k, SEL_CAP, SEL_MAX_KOUT);
    if (Q == 0) return CIR_OK;
    SelParams p{};
    CIR_REQUIRE(g_stride == 0 || g_stride >= (int64_t)Q * k, CIR_ERR_INVALID_ARG, "cir_topk_merge: g_stride too small");
    p.in_scores = scores; p.in_idx = idx; p.kin = k;
    p.g_stride = g_stride ? g_stride : (long long)Q * k;
    p.G = G; p.Q = Q; p.k = k_out;
    p.out_scores = out_scores; p.out_idx = out_idx; p.out_ld = k_out; p.idx_offset = 0;
    return launch_select<1>(p, Q, static_cast<cudaStream_t>(stream));
}

extern "C" int cir_rescore_topk(const float* q32, int Q, const float* db32, int64_t N, int D, const int32_t* cand,
                                int Kc, int32_t idx_offset, float* out_scores, int32_t* out_idx, int k_out,
                                void* stream) {
    CIR_REQUIRE(q32 && db32 && cand && out_scores && out_idx, CIR_ERR_INVALID_ARG, "cir_rescore_topk: null pointer");
    CIR_REQUIRE(Q >= 0 && N > 0 && D > 0 && Kc >= 1 && k_out >= 1 && k_out <= Kc, CIR_ERR_INVALID_ARG,
                "cir_rescore_topk: bad shape (Q=%d D=%d Kc=%d k_out=%d)", Q, D, Kc, k_out);
    CIR_REQUIRE(Kc <= SEL_N, CIR_ERR_UNSUPPORTED, "cir_rescore_topk: Kc=%d exceeds %d", Kc, SEL_N);
    if (Q == 0) return CIR_OK;
    int Kp = 2;
    while (Kp < Kc) Kp <<= 1;
    const size_t smem = (size_t)Kp * 8 + (size_t)D * 4;
    const DeviceInfo& dev = device_info();
    CIR_REQUIRE((int)smem <= dev.max_smem_optin, CIR_ERR_UNSUPPORTED, "cir_rescore_topk: D=%d too large for shared memory", D);
    if (smem > 48 * 1024)
        CIR_CHECK_CUDA(cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rescore_kernel<<<Q, RESCORE_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(q32, db32, N, D, cand, Kc, Kp, idx_offset,
                                                                                   out_scores, out_idx, k_out);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}

extern "C" int cir_sort_rows_workspace_bytes(int Q, int64_t N, size_t* bytes) {
    CIR_REQUIRE(bytes && Q > 0 && N > 0, CIR_ERR_INVALID_ARG, "cir_sort_rows_workspace_bytes: bad arguments");
    *bytes = N > SORT_RADIX_MIN_N ? radix_sort_workspace_bytes(Q, N) : (size_t)Q * (size_t)sort_npad(N) * 8;
    return CIR_OK;
}

extern "C" int cir_sort_rows_desc(const float* scores, int Q, int64_t N, int64_t ld, int32_t* out_idx, float* out_sorted,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
    CIR_REQUIRE(scores && out_idx && Q > 0 && N > 0 && ld >= N, CIR_ERR_INVALID_ARG, "cir_sort_rows_desc: bad arguments");
    CIR_REQUIRE(N <= 0x7fffff00ll && Q <= 65535, CIR_ERR_UNSUPPORTED, "cir_sort_rows_desc: N or Q too large");
    size_t need = 0;
    cir_sort_rows_workspace_bytes(Q, N, &need);
    CIR_REQUIRE(workspace && workspace_bytes >= need, CIR_ERR_WORKSPACE, "cir_sort_rows_desc: workspace %zu < %zu bytes",
                workspace_bytes, need);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    static const char* dbg = getenv("CIR_DEBUG_SORT");               // experiments: "bitonic" forces the network
    if (N > SORT_RADIX_MIN_N && !(dbg && dbg[0] == 'b' && workspace_bytes >= (size_t)Q * (size_t)sort_npad(N) * 8))
        return radix_sort_rows_desc(scores, Q, N, ld, out_idx, out_sorted, workspace, stream);
    const long long Npad = sort_npad(N);
    unsigned long long* keys = static_cast<unsigned long long*>(workspace);
    int launches = 0;
    {
        dim3 grid((unsigned)((Npad + 255) / 256), Q);
        sort_build_keys<<<grid, 256, 0, stream>>>(scores, N, ld, Npad, keys);
        ++launches;
    }
    const dim3 tiles((unsigned)(Npad / SEL_N), Q);
    sort_local_kernel<<<tiles, SEL_THREADS, 0, stream>>>(keys, Npad, 2, SEL_N);
    ++launches;
    for (long long k = 2ll * SEL_N; k <= Npad; k <<= 1) {
        for (long long j = k >> 1; j >= SEL_N; j >>= 1) {
            dim3 grid((unsigned)(((Npad >> 1) + 255) / 256), Q);
            sort_global_stage<<<grid, 256, 0, stream>>>(keys, Npad, k, j);
            ++launches;
        }
        sort_local_kernel<<<tiles, SEL_THREADS, 0, stream>>>(keys, Npad, k, k);
        ++launches;
    }
    {
        dim3 grid((unsigned)((N + 255) / 256), Q);
        sort_extract<<<grid, 256, 0, stream>>>(keys, N, Npad, out_idx, out_sorted);
        ++launches;
    }
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch(launches);
    return CIR_OK;
}
