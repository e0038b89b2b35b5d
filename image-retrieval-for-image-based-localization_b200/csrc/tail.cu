// Fused descriptor tail:  pool (GeM / MAC / SPoC) -> L2N -> whitening Linear -> L2N.
//
// Replaces cirtorch/modules/pools.py:37-38, normalizations.py:15-16 and
// heads/global_head.py:52-67 (seven eager PyTorch ops, three full passes over the map)
// with ONE cooperative launch of one 512-thread CTA per SM:
//   phase A  the HBM stream.  Warp 15 is a producer: one thread feeds a ring of 16 KB
//            shared-memory slots with cp.async.bulk (TMA bulk copies of whole (n, c) rows,
//            L2 evict-first) completing on mbarriers; warps 0..14 consume rows from shared
//            memory (clamp, x^p, sum / max; warp-shuffle reduce) and finish 32 rows at a
//            time (mean^(1/p)) -> pooled[n, c] (N*C floats, stays in L2).  The bytes in
//            flight are the ring, not registers.  Meanwhile the CTA's slice of W (<= 16
//            output rows, <= 128 KB) is prefetched into shared memory with cp.async.
//            Rows that TMA cannot move (H*W % 4 != 0, > 16 KB, unaligned) take a direct-load path.
//   barrier  cooperative grid sync
//   phase B  CTA i owns a slice of <= 16 output dims.  Y[64 x 16] = G[64 x K] . Wslice^T on
//            mma.sync m16n8k16 bf16 with every fp32 operand split into hi + lo bf16
//            (hi.hi + hi.lo + lo.hi, fp32 accumulate: ~1e-5 relative, inside the 1e-4 bar).
//            16 warps = 2 image halves x 8 K ranges; A fragments are loaded straight from the
//            L2-resident pooled vectors (one full 128 B line per 4 lanes, next block
//            prefetched), B fragments from the shared-memory W slice; a shared-memory
//            reduction over the K ranges; the first L2N is folded in as a scale.
//   barrier  cooperative grid sync (per-chunk partial sums of squares)
//   phase C  second L2N: partial sums added in a fixed order (deterministic), outputs rescaled.
//
// Algorithmic HBM bytes per launch: N*C*H*W*4 (x) + D_out*C*4 (W) + D_out*4 (b) + N*D_out*4 (out).
#include "common.cuh"
#include "ptx.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace cir {

using namespace ptx;

constexpr int TAIL_THREADS = 512;
constexpr int TAIL_WARPS = TAIL_THREADS / 32;
constexpr int TAIL_CONSUMERS = TAIL_WARPS - 1;       // phase A: warps 0..14 consume, warp 15 produces
constexpr int TAIL_JMAX = 16;                        // output dims per phase-B chunk (two n8 MMA tiles)
constexpr int TAIL_KC = 2048;                        // K extent of the W slice staged in shared memory
constexpr int TAIL_SLOT_BYTES = 16384;               // one ring slot: whole rows, <= 16 KB
constexpr int TAIL_MAX_SLOTS = 8;
constexpr int TAIL_MB = 64;                          // images per phase-B pass
constexpr size_t TAIL_STAMP_BYTES = 1024 * 8 * 8;    // debug time stamps: up to 1024 CTAs x 8 slots
// phase-B reduction scratch (aliases the ring): [8 K ranges][8 tile combos][32 lanes][4] + [8][64] + [64][16] + [64]
constexpr int TAIL_RED_FLOATS = 8 * 8 * 32 * 4 + 8 * TAIL_MB + TAIL_MB * TAIL_JMAX + TAIL_MB;

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
    int bulk_ok;       // rows can be moved by cp.async.bulk
    int vec_ok;        // rows are 16 B aligned and HW % 4 == 0
    int n_slots;       // ring slots
    int w_bytes;       // shared-memory bytes reserved for the W slice
    unsigned long long* stamps;   // optional [gridDim][8] globaltimer stamps (CIR_TAIL_DEBUG_STAMPS)
};

__device__ __forceinline__ void stamp(const TailParams& P, int slot) {
    if (P.stamps && threadIdx.x == 0) P.stamps[(size_t)blockIdx.x * 8 + slot] = global_timer_ns();
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
__device__ __forceinline__ float fold(float acc, float v, float eps, float p) {
    if (PM == PM_MAX) return fmaxf(acc, v);
    if (PM == PM_MEAN) return acc + v;
    const float t = fmaxf(v, eps);
    if (PM == PM_1) return acc + t;
    if (PM == PM_2) return fmaf(t, t, acc);
    if (PM == PM_3) return fmaf(t * t, t, acc);
    if (PM == PM_4) { const float t2 = t * t; return fmaf(t2, t2, acc); }
    return acc + fast_ex2(p * fast_lg2(t));
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

// per-lane partial of one row held as float4s (shared or global memory), lanes stride the vectors
template <int PM, bool GLOBAL>
__device__ __forceinline__ float row_partial_vec(const float4* v, int nvec, int lane, float eps, float p) {
    float a0 = PM == PM_MAX ? -INFINITY : 0.0f, a1 = a0;
    int i = lane;
    for (; i + 32 * 7 < nvec; i += 256) {
        float4 u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = GLOBAL ? ld_stream_f4(v + i + 32 * j) : v[i + 32 * j];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            a0 = fold4<PM>(a0, u[j], eps, p);
            a1 = fold4<PM>(a1, u[j + 1], eps, p);
        }
    }
    for (; i < nvec; i += 32) a0 = fold4<PM>(a0, GLOBAL ? ld_stream_f4(v + i) : v[i], eps, p);
    return PM == PM_MAX ? fmaxf(a0, a1) : a0 + a1;
}
template <int PM>
__device__ __forceinline__ float row_partial_scalar(const float* x, int n, int lane, float eps, float p) {
    float a = PM == PM_MAX ? -INFINITY : 0.0f;
    for (int i = lane; i < n; i += 32) a = fold<PM>(a, ld_stream_f1(x + i), eps, p);
    return a;
}

// warp-uniform dispatch on the exponent class; returns the warp-reduced row statistic in every lane
template <bool GLOBAL>
__device__ __forceinline__ float row_reduce(int pm, const float* row, int HW, bool vec, int lane, float eps, float p) {
    float a;
#define CIR_ROW_CASE(PMV)                                                                               \
    case PMV:                                                                                           \
        a = vec ? row_partial_vec<PMV, GLOBAL>(reinterpret_cast<const float4*>(row), HW >> 2, lane, eps, p) \
                : row_partial_scalar<PMV>(row, HW, lane, eps, p);                                       \
        break;
    switch (pm) {
        CIR_ROW_CASE(PM_3)
        CIR_ROW_CASE(PM_2)
        CIR_ROW_CASE(PM_1)
        CIR_ROW_CASE(PM_4)
        CIR_ROW_CASE(PM_MAX)
        CIR_ROW_CASE(PM_MEAN)
        default:
            a = vec ? row_partial_vec<PM_GENERAL, GLOBAL>(reinterpret_cast<const float4*>(row), HW >> 2, lane, eps, p)
                    : row_partial_scalar<PM_GENERAL>(row, HW, lane, eps, p);
    }
#undef CIR_ROW_CASE
    return pm == PM_MAX ? warp_max(a) : warp_sum(a);
}

__device__ __forceinline__ float finish_row(int pm, float acc, int HW, float p) {
    if (pm == PM_MAX) return acc;
    const float mean = acc / (float)HW;
    if (pm == PM_MEAN || pm == PM_1) return mean;
    // reference: .pow(1. / self.p) with the reciprocal rounded to fp32 (pools.py:38)
    return powf(mean, 1.0f / p);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// TMA bulk copy global -> shared (1-D, size % 16 == 0), completes on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                          uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_k, float hi_k) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo_k, hi_k);     // .x (low half) = lower k index
    return *reinterpret_cast<uint32_t*>(&v);
}
// split two fp32 values into packed bf16 hi parts and packed bf16 residuals
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = pack_bf16x2(a - __low2float(h), b - __high2float(h));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) tail_fused_kernel(const TailParams P) {
    extern __shared__ __align__(128) unsigned char tail_smem[];
    float* Ws = reinterpret_cast<float*>(tail_smem);                               // [jch][TAIL_KC]
    unsigned char* ring = tail_smem + P.w_bytes;                                   // [n_slots][16 KB]
    float* red = reinterpret_cast<float*>(ring);                                   // phase B scratch (aliases the ring)
    __shared__ __align__(8) uint64_t full_bar[TAIL_MAX_SLOTS];
    __shared__ __align__(8) uint64_t empty_bar[TAIL_MAX_SLOTS];
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
        const long long r0 = rows * blockIdx.x / gridDim.x, r1 = rows * (blockIdx.x + 1) / gridDim.x;
        const int my_rows = (int)(r1 - r0);
        const int HW = P.HW;
        const bool gem = P.pool_mode == CIR_POOL_GEM;

        // lane l of a consumer warp parks the statistic of the l-th row of the current batch of 32
        float held = 0.0f;
        long long held_row = -1;
        int nheld = 0;
        auto flush = [&]() {
            if (held_row >= 0) {
                const int c = (int)(held_row % P.C);
                const long long n = held_row / P.C;
                const float pl = gem ? __ldg(P.p + c * P.p_stride) : 1.0f;
                P.pooled[n * P.pooled_ld + c] = finish_row(classify_p(P.pool_mode, pl), held, HW, pl);
            }
            held_row = -1;
        };
        auto take = [&](float a, long long row) {
            if (lane == (nheld & 31)) { held = a; held_row = row; }
            if ((++nheld & 31) == 0) flush();
        };

        if (P.bulk_ok) {
            const int rps = max(1, TAIL_SLOT_BYTES / (HW * 4));            // rows per slot
            const int iters = (my_rows + rps - 1) / rps;
            if (tid == 0) {
                for (int s = 0; s < P.n_slots; ++s) {
                    mbar_init(&full_bar[s], 1);
                    mbar_init(&empty_bar[s], TAIL_CONSUMERS);
                }
                fence_barrier_init();
            }
            __syncthreads();
            if (warp == TAIL_CONSUMERS) {
                if (lane == 0) {
                    const uint64_t pol = policy_evict_first();
                    int slot = 0;
                    uint32_t phase = 0;
                    for (int t = 0; t < iters; ++t) {
                        mbar_wait(&empty_bar[slot], phase ^ 1u);
                        const int nr = min(rps, my_rows - t * rps);
                        const uint32_t bytes = (uint32_t)nr * (uint32_t)HW * 4u;
                        mbar_arrive_expect_tx(&full_bar[slot], bytes);
                        bulk_load(ring + (size_t)slot * TAIL_SLOT_BYTES, P.x + (r0 + (long long)t * rps) * HW, bytes,
                                  &full_bar[slot], pol);
                        if (++slot == P.n_slots) { slot = 0; phase ^= 1u; }
                    }
                }
            } else {
                int slot = 0;
                uint32_t phase = 0;
                // the exponent class is warp-uniform per row; with a shared p it is constant
                for (int t = 0; t < iters; ++t) {
                    mbar_wait(&full_bar[slot], phase);
                    const int nr = min(rps, my_rows - t * rps);
                    const int base = t * rps;
                    int j = (warp - base % TAIL_CONSUMERS + TAIL_CONSUMERS) % TAIL_CONSUMERS;   // first local row with (base + j) % 15 == warp
                    for (; j < nr; j += TAIL_CONSUMERS) {
                        const long long row = r0 + base + j;
                        const float pr = gem ? __ldg(P.p + (int)(row % P.C) * P.p_stride) : 1.0f;
                        const float* src = reinterpret_cast<const float*>(ring + (size_t)slot * TAIL_SLOT_BYTES) + (size_t)j * HW;
                        const float a = row_reduce<false>(classify_p(P.pool_mode, pr), src, HW, true, lane, P.eps_gem, pr);
                        take(a, row);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[slot]);
                    if (++slot == P.n_slots) { slot = 0; phase ^= 1u; }
                }
                flush();
            }
        } else {
            // direct loads: every warp takes rows r0 + warp, r0 + warp + 16, ...
            for (int j = warp; j < my_rows; j += TAIL_WARPS) {
                const long long row = r0 + j;
                const float pr = gem ? __ldg(P.p + (int)(row % P.C) * P.p_stride) : 1.0f;
                const float a = row_reduce<true>(classify_p(P.pool_mode, pr), P.x + row * HW, HW, P.vec_ok != 0, lane, P.eps_gem, pr);
                take(a, row);
            }
            flush();
        }
    }
    stamp(P, 1);
    if (pool_only) return;

    grid.sync();
    stamp(P, 2);

    // ------------------------------------------------------------------ no whitening
    if (!whiten) {
        __shared__ float redn[TAIL_WARPS];
        for (int n = blockIdx.x; n < P.N; n += gridDim.x) {
            const float* g = P.pooled + (size_t)n * P.pooled_ld;
            float ss = 0.0f;
            for (int c = tid; c < P.C; c += TAIL_THREADS) { float v = __ldcg(g + c); ss += v * v; }
            ss = warp_sum(ss);
            __syncthreads();
            if (lane == 0) redn[warp] = ss;
            __syncthreads();
            float tot = 0.0f;
#pragma unroll
            for (int w = 0; w < TAIL_WARPS; ++w) tot += redn[w];
            const float denom = sqrtf(tot) + P.eps_l2;
            for (int c = tid; c < P.C; c += TAIL_THREADS)
                P.out[(size_t)n * P.out_ld + c] = __ldcg(g + c) / denom;
        }
        return;
    }

    // ------------------------------------------------------------------ phase B
    {
        float* red_acc = red;                               // [8 wk][8 combo][32 lanes][4]
        float* red_ss = red_acc + 8 * 8 * 32 * 4;           // [8 wk][64]
        float* ysq = red_ss + 8 * TAIL_MB;                  // [64][16]
        float* inv_s = ysq + TAIL_MB * TAIL_JMAX;           // [64]
        const int g = lane >> 2, tig = lane & 3;
        const int wm = warp & 1, wk = warp >> 1;            // image half, K range
        for (int chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
            const int j0 = chunk * P.jch;
            const int J = min(P.jch, P.D_out - j0);
            for (int n0 = 0; n0 < P.N; n0 += TAIL_MB) {
                float acc[2][2][4];
                float ss[2][2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int x2 = 0; x2 < 2; ++x2) {
                        ss[mt][x2] = 0.0f;
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[mt][x2][r] = 0.0f;
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
                    // this warp's K range inside the tile, in blocks of 32
                    const int kr0 = wk * (TAIL_KC / 8);
                    const int nblk = max(0, min(TAIL_KC / 8, kw - kr0) + 31) >> 5;
                    // rows (images) of the 4 fragment row slots: [mt][rsel]
                    const float* arow[2][2];
                    bool aok[2][2];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int rs = 0; rs < 2; ++rs) {
                            const int n = n0 + wm * 32 + mt * 16 + g + rs * 8;
                            aok[mt][rs] = n < P.N;
                            arow[mt][rs] = P.pooled + (size_t)(aok[mt][rs] ? n : 0) * P.pooled_ld + kc0 + kr0 + tig * 8;
                        }
                    auto load_a = [&](int blk, float4 (&a)[2][2][2]) {
                        const int kk = kr0 + blk * 32 + tig * 8;
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                            for (int rs = 0; rs < 2; ++rs)
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    a[mt][rs][h] = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (aok[mt][rs] && kk + h * 4 < kw)
                                        a[mt][rs][h] = __ldcg(reinterpret_cast<const float4*>(arow[mt][rs] + blk * 32 + h * 4));
                                }
                    };
                    // CTAs start at different blocks so that they do not hit the same L2 lines at the same moment
                    int blk = nblk > 0 ? (int)(blockIdx.x % (unsigned)nblk) : 0;
                    float4 a_next[2][2][2];
                    if (nblk > 0) load_a(blk, a_next);
                    for (int t = 0; t < nblk; ++t) {
                        float4 a[2][2][2];
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                            for (int rs = 0; rs < 2; ++rs)
#pragma unroll
                                for (int h = 0; h < 2; ++h) a[mt][rs][h] = a_next[mt][rs][h];
                        const int kk = kr0 + blk * 32 + tig * 8;
                        if (++blk == nblk) blk = 0;
                        if (t + 1 < nblk) load_a(blk, a_next);
                        // B fragments of the two 8-dim tiles: W rows j = nt*8 + g, the same 8 consecutive k as A
                        uint32_t bhi[2][2][2], blo[2][2][2];     // [nt][kstep][reg]
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) {
                            const int j = nt * 8 + g;
                            float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
                            if (j < J && kk < kw) w0 = *reinterpret_cast<const float4*>(&Ws[j * TAIL_KC + kk]);
                            if (j < J && kk + 4 < kw) w1 = *reinterpret_cast<const float4*>(&Ws[j * TAIL_KC + kk + 4]);
                            split2(w0.x, w0.y, bhi[nt][0][0], blo[nt][0][0]);
                            split2(w0.z, w0.w, bhi[nt][0][1], blo[nt][0][1]);
                            split2(w1.x, w1.y, bhi[nt][1][0], blo[nt][1][0]);
                            split2(w1.z, w1.w, bhi[nt][1][1], blo[nt][1][1]);
                        }
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                            for (int rs = 0; rs < 2; ++rs) {
                                const float4 u = a[mt][rs][0], v = a[mt][rs][1];
                                ss[mt][rs] += (u.x * u.x + u.y * u.y) + (u.z * u.z + u.w * u.w) + (v.x * v.x + v.y * v.y) +
                                              (v.z * v.z + v.w * v.w);
                            }
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                // k step ks uses the floats 4*ks .. 4*ks+3 of the lane's 8: slots (2 tig, 2 tig + 1) and (+8, +9)
                                const float4 r0v = a[mt][0][ks], r1v = a[mt][1][ks];
                                uint32_t ahi[4], alo[4];
                                split2(r0v.x, r0v.y, ahi[0], alo[0]);
                                split2(r1v.x, r1v.y, ahi[1], alo[1]);
                                split2(r0v.z, r0v.w, ahi[2], alo[2]);
                                split2(r1v.z, r1v.w, ahi[3], alo[3]);
#pragma unroll
                                for (int nt = 0; nt < 2; ++nt) {
                                    mma_bf16_16816(acc[mt][nt], ahi, bhi[nt][ks]);
                                    mma_bf16_16816(acc[mt][nt], ahi, blo[nt][ks]);
                                    mma_bf16_16816(acc[mt][nt], alo, bhi[nt][ks]);
                                }
                            }
                        }
                    }
                }
                // ---- reduce the 8 K ranges through shared memory
                __syncthreads();   // the ring / previous pass scratch is free
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const int combo = (wm * 2 + mt) * 2 + nt;
                        *reinterpret_cast<float4*>(&red_acc[((wk * 8 + combo) * 32 + lane) * 4]) =
                            make_float4(acc[mt][nt][0], acc[mt][nt][1], acc[mt][nt][2], acc[mt][nt][3]);
                    }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int rs = 0; rs < 2; ++rs) {
                        float s = ss[mt][rs];
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        if (tig == 0) red_ss[wk * TAIL_MB + wm * 32 + mt * 16 + g + rs * 8] = s;
                    }
                __syncthreads();
                if (tid < TAIL_MB) {
                    float s = 0.0f;
#pragma unroll
                    for (int k8 = 0; k8 < 8; ++k8) s += red_ss[k8 * TAIL_MB + tid];
                    inv_s[tid] = 1.0f / (sqrtf(s) + P.eps_l2);      // first L2N: W.(g/(|g|+eps)) == (W.g)/(|g|+eps)
                }
                __syncthreads();
                {
                    // thread (warp, lane): tile combo = warp & 7, fragment row half = warp >> 3
                    const int combo = warp & 7, rp = warp >> 3;
                    const int cwm = combo >> 2, cmt = (combo >> 1) & 1, cnt = combo & 1;
                    float y0 = 0.0f, y1 = 0.0f;
#pragma unroll
                    for (int k8 = 0; k8 < 8; ++k8) {
                        const float2 v = *reinterpret_cast<const float2*>(&red_acc[((k8 * 8 + combo) * 32 + lane) * 4 + rp * 2]);
                        y0 += v.x;
                        y1 += v.y;
                    }
                    const int nl = cwm * 32 + cmt * 16 + g + rp * 8;
                    const int jl = cnt * 8 + tig * 2;
                    const int n = n0 + nl;
                    const float inv = inv_s[nl];
                    y0 = y0 * inv + ((jl < J && P.bias) ? __ldg(P.bias + j0 + jl) : 0.0f);
                    y1 = y1 * inv + ((jl + 1 < J && P.bias) ? __ldg(P.bias + j0 + jl + 1) : 0.0f);
                    if (jl >= J) y0 = 0.0f;
                    if (jl + 1 >= J) y1 = 0.0f;
                    if (n < P.N) {
                        if (jl < J) P.out[(size_t)n * P.out_ld + j0 + jl] = y0;
                        if (jl + 1 < J) P.out[(size_t)n * P.out_ld + j0 + jl + 1] = y1;
                    }
                    ysq[nl * TAIL_JMAX + jl] = y0 * y0;
                    ysq[nl * TAIL_JMAX + jl + 1] = y1 * y1;
                }
                __syncthreads();
                if (tid < TAIL_MB && n0 + tid < P.N) {
                    float s = 0.0f;
#pragma unroll
                    for (int j = 0; j < TAIL_JMAX; ++j) s += ysq[tid * TAIL_JMAX + j];
                    P.partial[(size_t)chunk * P.N + n0 + tid] = s;
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
    // pooled [N, C] + partial [D_out, N] (n_chunks <= D_out) + debug stamps
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
    P.bulk_ok = P.vec_ok && (size_t)P.HW * 4 <= (size_t)TAIL_SLOT_BYTES;

    const int grid = dev.num_sms;
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
        P.w_bytes = (int)align_up((size_t)jch * TAIL_KC * sizeof(float), 1024);
    }
    // ring: as many 16 KB slots as fit beside the W slice (and at least the phase-B scratch)
    const int static_smem = 4096;     // barriers + phase-C arrays, with slack
    int slots = (dev.max_smem_optin - static_smem - P.w_bytes) / TAIL_SLOT_BYTES;
    if (slots > TAIL_MAX_SLOTS) slots = TAIL_MAX_SLOTS;
    const int min_slots = (int)((TAIL_RED_FLOATS * sizeof(float) + TAIL_SLOT_BYTES - 1) / TAIL_SLOT_BYTES);
    CIR_REQUIRE(slots >= (whiten ? min_slots : 1), CIR_ERR_UNSUPPORTED, "cir_tail_fwd: shared memory");
    P.n_slots = slots;
    const size_t smem = (size_t)P.w_bytes + (size_t)slots * TAIL_SLOT_BYTES;
    static thread_local int attr_set_dev = -1;
    if (attr_set_dev != dev.device) {
        CIR_CHECK_CUDA(cudaFuncSetAttribute(tail_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            dev.max_smem_optin - static_smem));
        attr_set_dev = dev.device;
    }
    void* args[] = {&P};
    if (pool_only) {
        // no grid barrier on this path: a plain launch is enough
        tail_fused_kernel<<<grid, TAIL_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(P);
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
