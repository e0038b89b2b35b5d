// Fused descriptor tail:  pool (GeM / MAC / SPoC) -> L2N -> whitening Linear -> L2N.
//
// Replaces cirtorch/modules/pools.py:37-38, normalizations.py:15-16 and
// heads/global_head.py:52-67 (seven eager PyTorch ops, three full passes over the map)
// with ONE cooperative launch of one CTA per SM, in one of two shapes (TAIL_THREADS_LIGHT / _HEAVY: 512 threads for cheap rows,
// 640 for a non-integer exponent):
//   phase A  the HBM stream.  The last warp is a producer: one thread feeds a ring of 16 KB shared-memory slots with
//            cp.async.bulk (TMA bulk copies of whole (n, c) rows, L2 evict-first) completing on mbarriers; the other 15 / 19
//            warps consume rows from shared memory (clamp, x^p, sum / max; warp-shuffle reduce) and finish 32 rows at a time
//            (mean^(1/p)).  The bytes in flight are the ring, not registers.  The consumer loop is instantiated per
//            exponent class (p = 1, 2, 3, 4, general, max, mean), so its body is the inlined arithmetic; the (image,
//            channel) of a row is tracked incrementally; with a non-integer exponent a slot is handed back as soon as the
//            row sits in registers.  Between rows the consumer lanes convert this CTA's W tile
//            (128 output dims x 256 k, fp32, read from HBM once) into bf16 hi / lo UMMA tiles in shared memory (128 KB).
//            Rows that TMA cannot move (H*W % 4 != 0, > 16 KB, unaligned) take a direct-load path.
//            Phase A stores the pooled vectors already split into bf16 hi + lo parts (and, on request, as fp32 for the
//            backward pass).
//   barrier  cooperative grid sync
//   phase B  split-K projection on tcgen05: unit (nt, ks) = 128 output dims x a 256-wide K slice;
//            part[ks][n][nt*128 ..] = G[128 images x 256] . W_tile^T with every fp32 operand split into
//            hi + lo bf16 (hi.hi + hi.lo + lo.hi, fp32 accumulator in TMEM: ~1e-5 relative, inside
//            the 1e-4 bar).  A CTA reads only ITS K slice of the pooled vectors (64 KB by TMA, 128 B
//            swizzle).  Warp 0 TMA, warp 1 MMA issue, warp 2 TMEM alloc, warps 8-11 epilogue (tcgen05.ld -> partial sums).
//   barrier  cooperative grid sync
//   phase C  one CTA per image: adds the K-slice partial sums in a fixed order (deterministic), applies
//            the first L2N as a scale (||g|| from the pooled vector), the bias and the second L2N.
//
// Algorithmic HBM bytes per launch: N*C*H*W*4 (x) + D_out*C*4 (W) + D_out*4 (b) + N*D_out*4 (out).
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace cir {

using namespace ptx;

// Two launch shapes of the same kernel.  512 threads (15 consumer warps + 1 producer, 120 registers): the integer exponents,
// MAC and SPoC -- a row is a handful of FMAs per element.  640 threads (19 + 1, 96 registers): everything else.  With a
// non-integer exponent a row is two MUFU operations per element, the dependent chain of a row is long and 3.75 warps per
// scheduler leave the MUFU pipe idle between rows; a fifth warp per scheduler is worth 4-5 us per launch (64 x 2048 x 32 x 32,
// p = 2.7: 102.6 -> 98.1 us), and costs the integer path 0.4 us (94.0 -> 94.5) -- so the host picks, from a hint, and the result
// never depends on the hint being right (the exponent is classified on the device either way).  576: 104.0 us, 704 / 768
// threads (80 registers, spills): 99.8 / 97.2 us at p = 2.7 but 95.4-95.7 us at p = 3.
constexpr int TAIL_THREADS_LIGHT = 512;
constexpr int TAIL_THREADS_HEAVY = 640;
constexpr int TAIL_NT = 128;                         // output dims per projection unit (UMMA N)
constexpr int TAIL_KS = 256;                         // K slice per projection unit (4 k blocks of 64)
constexpr int TAIL_SLOT_BYTES = 16384;               // one ring slot: whole rows (phase A) / one 128 x 64 bf16 A tile (phase B)
constexpr int TAIL_MAX_SLOTS = 8;
constexpr int TAIL_MAX_BARS = 32;
constexpr int TAIL_MB = 128;                         // images per phase-B pass (UMMA M)
constexpr int TAIL_BT_BYTES = TAIL_NT * 128;         // one B tile: 128 rows x 64 k bf16, 128 B swizzle (16 KB)
constexpr int TAIL_W_BYTES = (TAIL_KS / 64) * 2 * TAIL_BT_BYTES;   // hi + lo tiles of a unit's W tile: 128 KB
constexpr int TAIL_NPOLY_DEFAULT = 1;                 // general exponent: 1 of every 4 ex2 on the FMA pipe
constexpr size_t TAIL_STAMP_BYTES = 1024 * 8 * 8;    // debug time stamps: up to 1024 CTAs x 8 slots

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
    float* pooled;     // [N, pooled_ld] fp32 (pool-only / no-whiten modes)
    int pooled_ld;
    __nv_bfloat16* pooled_hi;   // whiten mode: pooled vectors as bf16 hi + lo parts, [N, C] each (same bytes as fp32)
    __nv_bfloat16* pooled_lo;
    float* pooled_out; // optional caller-owned [N, C] fp32 copy of the pooled values (kept for the backward pass)
    float* z_out;      // optional caller-owned [N, D_out] fp32: the projection W.L2N(g) + b BEFORE the last L2N (backward pass)
    float* part;       // [n_kslices][N][D_out] partial projections of the K slices
    int n_ntiles, n_kslices, units;   // projection units: u = ks * n_ntiles + nt
    unsigned flags;
    int bulk_ok;       // rows can be moved by cp.async.bulk
    int vec_ok;        // rows are 16 B aligned and HW % 4 == 0
    int n_slots;       // ring slots
    int w_bytes;       // shared-memory bytes reserved for the W slice
    int gen_mode;      // exponent class of a non-integer p on vector rows: PM_GENERAL or PM_GENERAL_POLY0 + NPOLY
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
enum { PM_GENERAL = 0, PM_1 = 1, PM_2 = 2, PM_3 = 3, PM_4 = 4, PM_MAX = 5, PM_MEAN = 6,
       PM_GENERAL_POLY0 = 16 /* + NPOLY in 0..4: vector rows with NPOLY of 4 ex2 on the FMA pipe */ };

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
// 2^t on the FMA / ALU pipes (no MUFU): t = n + f with n = rne(t) taken from the low mantissa bits of t + 1.5 * 2^23,
// 2^f by a degree-4 minimax polynomial on [-0.5, 0.5] (max relative error 2.7e-6), 2^n added into the exponent field.
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
// general exponent, second flavour: lg2 on the MUFU pipe, 2^t on the FMA pipe
__device__ __forceinline__ float fold_poly(float acc, float v, float eps, float p) {
    return acc + poly_ex2(p * fast_lg2(fmaxf(v, eps)));
}
// General exponent: x^p = ex2(p * lg2 x) costs two MUFU operations per element (4 lanes per clock and SM sub-partition:
// 8.7 cycles per warp instruction measured, scripts/pow_microbench.cu).  PM_GENERAL_POLY0 + NPOLY sends NPOLY of every 4 ex2
// to the FMA pipe instead (poly_ex2).  Compute only (rows resident in shared memory, no HBM stream) the 32-element step
// costs 17.3 cycles per sub-partition with NPOLY = 0, 15.3 / 13.9 / 15.2 / 17.7 with 1 / 2 / 3 / 4 -- all inside the 22.2
// cycles the HBM stream leaves.  Inside the kernel (per launch, 64 x 2048 x 32 x 32, p = 2.7): 110.7 / 109.3 / 113.3 /
// 120.3 us for NPOLY 0 / 1 / 2 / 3: the extra FMA-pipe instructions compete with the per-row bookkeeping for issue slots,
// so only NPOLY = 1 pays.
template <int PM>
__device__ __forceinline__ float fold4(float acc, const float4& v, float eps, float p) {
    if (PM >= PM_GENERAL_POLY0) {
        constexpr int NPOLY = PM - PM_GENERAL_POLY0;
        acc = NPOLY >= 4 ? fold_poly(acc, v.x, eps, p) : fold<PM_GENERAL>(acc, v.x, eps, p);
        acc = NPOLY >= 1 ? fold_poly(acc, v.y, eps, p) : fold<PM_GENERAL>(acc, v.y, eps, p);
        acc = NPOLY >= 3 ? fold_poly(acc, v.z, eps, p) : fold<PM_GENERAL>(acc, v.z, eps, p);
        acc = NPOLY >= 2 ? fold_poly(acc, v.w, eps, p) : fold<PM_GENERAL>(acc, v.w, eps, p);
        return acc;
    }
    acc = fold<PM>(acc, v.x, eps, p);
    acc = fold<PM>(acc, v.y, eps, p);
    acc = fold<PM>(acc, v.z, eps, p);
    acc = fold<PM>(acc, v.w, eps, p);
    return acc;
}

__device__ __forceinline__ int classify_p(int pool_mode, float p, int gen_mode = PM_GENERAL) {
    if (pool_mode == CIR_POOL_MAC) return PM_MAX;
    if (pool_mode == CIR_POOL_SPOC) return PM_MEAN;
    if (p == 3.0f) return PM_3;
    if (p == 2.0f) return PM_2;
    if (p == 1.0f) return PM_1;
    if (p == 4.0f) return PM_4;
    return gen_mode;
}

// 128-bit load from shared memory by its 32-bit shared address.  The ring pointer is derived from an aligned-up uintptr_t,
// so plain C++ loads through it compile to GENERIC loads (LD.E.128: address-space check, long-scoreboard latency) --
// ncu showed the consumers' first arithmetic instruction of every row waiting on them.  volatile: ordered after the
// mbarrier wait that makes the row visible and before the arrive that hands the slot back.
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
    return r;
}

// per-lane partial of one row held as float4s (shared or global memory), lanes stride the vectors
template <int PM, bool GLOBAL>
__device__ __forceinline__ float row_partial_vec(const float4* v, int nvec, int lane, float eps, float p) {
    float a0 = PM == PM_MAX ? -INFINITY : 0.0f, a1 = a0;
    const uint32_t sa = GLOBAL ? 0u : smem_u32(v);
    int i = lane;
    for (; i + 32 * 7 < nvec; i += 256) {
        float4 u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = GLOBAL ? ld_stream_f4(v + i + 32 * j) : lds_f4(sa + (uint32_t)(i + 32 * j) * 16u);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            a0 = fold4<PM>(a0, u[j], eps, p);
            a1 = fold4<PM>(a1, u[j + 1], eps, p);
        }
    }
    for (; i < nvec; i += 32) a0 = fold4<PM>(a0, GLOBAL ? ld_stream_f4(v + i) : lds_f4(sa + (uint32_t)i * 16u), eps, p);
    return PM == PM_MAX ? fmaxf(a0, a1) : a0 + a1;
}
// The same for a row in a ring slot, handing the slot back as soon as the row sits in REGISTERS: with a non-integer exponent
// the folding is most of a row's time, and the producer can refill the slot that much earlier.  The arrival must not be issued
// while a load is in flight -- the refilling bulk copy (async proxy) overtakes it (measured: launches differed) -- so one
// component of each of the eight vectors is clamped first (the fold clamps it anyway): those FMNMX cannot issue before the
// loads have returned, and the arrival is issued after them.
// Returns true if the slot was released here (rows whose vector count is not a multiple of 256 are released by the caller).
template <int PM>
__device__ __forceinline__ bool row_partial_vec_release(const float4* v, int nvec, int lane, float eps, float p, uint64_t* empty_bar,
                                                        float* out) {
    float a0 = 0.0f, a1 = 0.0f;
    const uint32_t sa = smem_u32(v);
    const bool whole = (nvec & 255) == 0;
    int i = lane;
    for (; i + 32 * 7 < nvec; i += 256) {
        float4 u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = lds_f4(sa + (uint32_t)(i + 32 * j) * 16u);
        if (whole && i + 256 >= nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j].x = fmaxf(u[j].x, eps);
            asm volatile("" : "+f"(u[0].x), "+f"(u[1].x), "+f"(u[2].x), "+f"(u[3].x), "+f"(u[4].x), "+f"(u[5].x), "+f"(u[6].x), "+f"(u[7].x));
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar);
        }
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            a0 = fold4<PM>(a0, u[j], eps, p);
            a1 = fold4<PM>(a1, u[j + 1], eps, p);
        }
    }
    for (; i < nvec; i += 32) a0 = fold4<PM>(a0, lds_f4(sa + (uint32_t)i * 16u), eps, p);
    *out = a0 + a1;
    return whole && nvec > 0;
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
        case PM_GENERAL_POLY0 + 1: a = row_partial_vec<PM_GENERAL_POLY0 + 1, GLOBAL>(reinterpret_cast<const float4*>(row), HW >> 2, lane, eps, p); break;
        case PM_GENERAL_POLY0 + 2: a = row_partial_vec<PM_GENERAL_POLY0 + 2, GLOBAL>(reinterpret_cast<const float4*>(row), HW >> 2, lane, eps, p); break;
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
// sum of squares of 8 values given as packed bf16 pairs of hi parts and of lo parts (value = hi + lo)
__device__ __forceinline__ float sumsq8(const uint4& h, const uint4& l) {
    float s = 0.0f;
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = __uint_as_float(hh[i] << 16) + __uint_as_float(ll[i] << 16);
        const float b = __uint_as_float(hh[i] & 0xffff0000u) + __uint_as_float(ll[i] & 0xffff0000u);
        s = fmaf(a, a, s);
        s = fmaf(b, b, s);
    }
    return s;
}
// fp32 W tile of projection unit u = (nt, ks) -> bf16 hi / lo UMMA B tiles (K-major, 128 B swizzle):
// tile (kb, part) at Bt + (kb * 2 + part) * 16 KB, row r at (r / 8) * 1024 + (r % 8) * 128, 16-byte chunk c at
// ((c ^ (r % 8)) * 16).  One task = one chunk (8 consecutive k of one row); rows >= D_out and k >= C become zeros.
// BATCH tasks are loaded before any is converted, so a lane has 2 * BATCH 16-byte loads in flight.
constexpr int TAIL_W_TASKS = TAIL_NT * (TAIL_KS / 8);         // 4096 chunks per W tile

// tasks t0, t0 + nworkers, ... (at most max_batches * BATCH of them)
template <int BATCH>
__device__ __forceinline__ void convert_w_tile(const TailParams& P, unsigned char* Bt, int unit, int worker, int nworkers,
                                               int first_batch = 0, int max_batches = 1 << 30) {
    const int nt = unit % P.n_ntiles, ks = unit / P.n_ntiles;
    const int j0 = nt * TAIL_NT, k0 = ks * TAIL_KS;
    constexpr int tasks = TAIL_W_TASKS;
    int nb = 0;
    for (int t0 = worker + first_batch * nworkers * BATCH; t0 < tasks && nb < max_batches; t0 += nworkers * BATCH, ++nb) {
        float4 w0[BATCH], w1[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const int t = t0 + u * nworkers;
            const int r = t >> 5, c8 = t & 31;
            const int j = j0 + r, k = k0 + c8 * 8;
            w0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            w1[u] = w0[u];
            if (t < tasks && j < P.D_out && k < P.C) {
                const float* w = P.Wt + (size_t)j * P.C + k;
                w0[u] = __ldg(reinterpret_cast<const float4*>(w));
                w1[u] = __ldg(reinterpret_cast<const float4*>(w + 4));
            }
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const int t = t0 + u * nworkers;
            if (t < tasks) {
                const int r = t >> 5, c8 = t & 31;
                const int kb = c8 >> 3, c = c8 & 7;
                uint4 hi, lo;
                split2(w0[u].x, w0[u].y, hi.x, lo.x);
                split2(w0[u].z, w0[u].w, hi.y, lo.y);
                split2(w1[u].x, w1[u].y, hi.z, lo.z);
                split2(w1[u].z, w1[u].w, hi.w, lo.w);
                unsigned char* dst = Bt + (size_t)(kb * 2) * TAIL_BT_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + TAIL_BT_BYTES) = lo;
            }
        }
    }
    fence_proxy_async();       // the tiles are read by the tensor core (async proxy)
}

template <int TAIL_THREADS>
__global__ void __launch_bounds__(TAIL_THREADS, 1)
tail_fused_kernel(const __grid_constant__ CUtensorMap tmHi, const __grid_constant__ CUtensorMap tmLo, const TailParams P) {
    constexpr int TAIL_WARPS = TAIL_THREADS / 32;
    constexpr int TAIL_CONSUMERS = TAIL_WARPS - 1;       // phase A: all warps but the last consume, the last one produces
    extern __shared__ __align__(1024) unsigned char tail_smem_raw[];
    unsigned char* tail_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tail_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* Bt = tail_smem;                                                 // W slice as UMMA B tiles (whiten)
    unsigned char* ring = tail_smem + P.w_bytes;                                   // [n_slots][16 KB]
    // Phase A barriers: a ring of nbar = M * n_slots (full, empty) pairs over the n_slots data slots.  A consumer warp
    // only visits the slots that hold its rows, so it does not observe every phase of a barrier; with a single ring a
    // parity wait could then be satisfied by an OLDER, still incomplete phase (bulk copies complete out of order).
    // With (M - 1) * n_slots >= the largest gap between two rows of a warp, the previous use of a barrier is complete
    // before anyone waits for its next use.
    __shared__ __align__(8) uint64_t full_bar[TAIL_MAX_BARS];
    __shared__ __align__(8) uint64_t empty_bar[TAIL_MAX_BARS];
    __shared__ __align__(8) uint64_t fullB[TAIL_MAX_SLOTS];      // phase B ring: TMA -> MMA / sum-of-squares warps
    __shared__ __align__(8) uint64_t emptyB[TAIL_MAX_SLOTS];
    __shared__ __align__(8) uint64_t accB;                       // accumulator complete
    __shared__ uint32_t tmem_slot;
    cg::grid_group grid = cg::this_grid();

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool pool_only = (P.flags & CIR_TAIL_POOL_ONLY) != 0;
    const bool whiten = !pool_only && !(P.flags & CIR_TAIL_NO_WHITEN);
    const bool accumulate = (P.flags & CIR_TAIL_ACCUMULATE) != 0;     // out += descriptor (multi-scale sum, GF_net.py:74-92)

    stamp(P, 0);
    if (P.stamps && threadIdx.x == 0) {          // profiling aid: which SM ran this CTA
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        P.stamps[(size_t)blockIdx.x * 8 + 7] = smid;
    }
    int conv_unit = -1;                      // which projection unit's W tile currently sits in Bt
    const bool own_unit = whiten && (int)blockIdx.x < P.units;

    // ------------------------------------------------------------------ phase A
    {
        const long long rows = (long long)P.N * P.C;
        const long long r0 = rows * blockIdx.x / gridDim.x, r1 = rows * (blockIdx.x + 1) / gridDim.x;
        const int my_rows = (int)(r1 - r0);
        const int HW = P.HW;
        const bool gem = P.pool_mode == CIR_POOL_GEM;

        // lane l of a consumer warp parks the statistic of the l-th row of the current batch of 32, with its (image, channel)
        float held = 0.0f;
        int held_n = -1, held_c = 0;
        int nheld = 0;
        auto flush = [&]() {
            if (held_n >= 0) {
                const int c = held_c;
                const long long n = held_n;
                const float pl = gem ? __ldg(P.p + c * P.p_stride) : 1.0f;
                const float v = finish_row(classify_p(P.pool_mode, pl), held, HW, pl);
                if (P.pooled_out) P.pooled_out[n * P.C + c] = v;
                if (P.pooled_hi) {
                    // the projection consumes bf16 hi + lo operands: split once here instead of in all 148 CTAs
                    const __nv_bfloat16 h = __float2bfloat16_rn(v);
                    P.pooled_hi[n * P.C + c] = h;
                    P.pooled_lo[n * P.C + c] = __float2bfloat16_rn(v - __bfloat162float(h));
                } else {
                    P.pooled[n * P.pooled_ld + c] = v;
                }
            }
            held_n = -1;
        };
        auto take = [&](float a, int n, int c) {
            if (lane == (nheld & 31)) { held = a; held_n = n; held_c = c; }
            if ((++nheld & 31) == 0) flush();
        };
        // The (image, channel) of a warp's current row is tracked incrementally: a 64-bit division / modulo per row (and two
        // per flushed row) cost more issue slots than the arithmetic of the row itself (ncu, r2c: ~170 of ~360 warp
        // instructions per row were index arithmetic, which held the MUFU pipe of the general exponent at 51 %).
        const int row_step = P.bulk_ok ? TAIL_CONSUMERS : TAIL_WARPS;     // rows between two rows of one warp
        int cur_n = 0, cur_c = 0;
        {
            const long long first = r0 + warp;
            cur_n = (int)(first / P.C);
            cur_c = (int)(first - (long long)cur_n * P.C);
        }
        const bool wide = P.C >= row_step;            // the usual case: at most one image boundary per step
        auto next_row = [&]() {
            cur_c += row_step;
            if (wide) {
                const bool wrap = cur_c >= P.C;
                cur_c -= wrap ? P.C : 0;
                cur_n += wrap ? 1 : 0;
            } else {
                while (cur_c >= P.C) { cur_c -= P.C; ++cur_n; }
            }
        };
        // one exponent for all channels: load and classify it once
        const float p_shared = (gem && P.p_stride == 0) ? __ldg(P.p) : 1.0f;

        if (P.bulk_ok) {
            // Slot size: 16 KB.  Smaller slots pin less of the ring while rows are being reduced, but the stream then moves
            // in smaller bulk copies and collapses (measured at 64 x 2048 x 32 x 32, p = 3: 16 KB slots 98.8 us, 8 KB
            // 120.9 us, 4 KB 204.8 us per launch).
            const int n_slots = P.n_slots;
            const int rps = max(1, TAIL_SLOT_BYTES / (HW * 4));              // rows per slot
            const int iters = (my_rows + rps - 1) / rps;
            // gap between two rows of a consumer warp: at most TAIL_CONSUMERS iterations (+1)
            const int nbar = min(TAIL_MAX_BARS, n_slots * ((TAIL_CONSUMERS + 1 + n_slots - 1) / n_slots + 1));
            if (tid == 0) {
                for (int s = 0; s < nbar; ++s) {
                    mbar_init(&full_bar[s], 1);
                    mbar_init(&empty_bar[s], rps);       // one arrival per consumed row (a short last slot is never refilled)
                }
                fence_barrier_init();
            }
            __syncthreads();
            if (warp == TAIL_CONSUMERS) {
                if (lane == 0) {
                    const uint64_t pol = policy_evict_first();
                    int slot = 0, bar = 0, ebar = 0;          // data slot, barrier of iteration t, barrier of iteration t - n_slots
                    uint32_t epar = 0;
                    for (int t = 0; t < iters; ++t) {
                        if (t >= n_slots) {                  // the slot's previous rows (iteration t - n_slots) are consumed
                            mbar_wait(&empty_bar[ebar], epar);
                            if (++ebar == nbar) { ebar = 0; epar ^= 1u; }
                        }
                        const int nr = min(rps, my_rows - t * rps);
                        const uint32_t bytes = (uint32_t)nr * (uint32_t)HW * 4u;
                        mbar_arrive_expect_tx(&full_bar[bar], bytes);
                        bulk_load(ring + (size_t)slot * TAIL_SLOT_BYTES, P.x + (r0 + (long long)t * rps) * HW, bytes,
                                  &full_bar[bar], pol);
                        if (++slot == n_slots) slot = 0;
                        if (++bar == nbar) bar = 0;
                    }
                }
            } else {
                // consumer warps; the consumer count is a compile-time constant in either variant (no runtime modulo)
                // PMC >= 0: one exponent class for every row, known before the loop -- the row arithmetic is inlined and the
                // loop carries no classification / jump table (those cost ~50 of the ~150 bookkeeping instructions per
                // row); PMC < 0: per-channel exponents (GeMmp), classified row by row.
                auto consume = [&](auto nc_tag, auto pm_tag) {
                    constexpr int NC = decltype(nc_tag)::value;
                    constexpr int PMC = decltype(pm_tag)::value;
                    // this CTA's W tile -> bf16 hi / lo UMMA tiles, spread over the 480 consumer lanes in three
                    // batches of 4 chunks interleaved with the row stream (each batch: 8 loads in flight per lane)
                    constexpr int CONV_BATCHES = (TAIL_W_TASKS + NC * 32 * 4 - 1) / (NC * 32 * 4);
                    const int my_count = (my_rows - warp + NC - 1) / NC;            // rows this warp will take
                    const int conv_every = max(1, my_count / (CONV_BATCHES + 1));
                    int conv_done = 0, taken = 0;
                    // warp w takes local rows w, w + NC, ...; it only touches the barriers of the slots holding them
                    // slot iterations / rows inside the slot between two rows of this warp (NC = q * rps + r); the two possible
                    // advances (q or q + 1 iterations) are reduced modulo the ring sizes once, so that the per-row update is a
                    // few selects instead of loops
                    const int step_q = NC / rps, step_r = NC - step_q * rps;
                    const int sl0 = step_q % n_slots, sl1 = (step_q + 1) % n_slots;
                    const int ba0 = step_q % nbar, ba1 = (step_q + 1) % nbar;
                    const uint32_t fl0 = (uint32_t)(step_q / nbar) & 1u, fl1 = (uint32_t)((step_q + 1) / nbar) & 1u;
                    int j = warp % rps;
                    int slot = (warp / rps) % n_slots, bar = (warp / rps) % nbar;
                    uint32_t par = (uint32_t)((warp / rps) / nbar) & 1u;
                    for (int i = warp; i < my_rows; i += NC) {
                        mbar_wait(&full_bar[bar], par);
                        const float* src = reinterpret_cast<const float*>(ring + (size_t)slot * TAIL_SLOT_BYTES) + (size_t)j * HW;
                        float a;
                        bool released = false;
                        if constexpr (PMC == PM_GENERAL || PMC >= PM_GENERAL_POLY0) {
                            released = row_partial_vec_release<PMC>(reinterpret_cast<const float4*>(src), HW >> 2, lane, P.eps_gem, p_shared,
                                                                    &empty_bar[bar], &a);
                            a = warp_sum(a);
                        } else if constexpr (PMC >= 0) {
                            a = row_partial_vec<PMC, false>(reinterpret_cast<const float4*>(src), HW >> 2, lane, P.eps_gem, p_shared);
                            a = PMC == PM_MAX ? warp_max(a) : warp_sum(a);
                        } else {
                            const float pr = __ldg(P.p + cur_c);
                            a = row_reduce<false>(classify_p(P.pool_mode, pr, P.gen_mode), src, HW, true, lane, P.eps_gem, pr);
                        }
                        if (!released) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&empty_bar[bar]);    // one arrival per row: count = rows per slot
                        }
                        take(a, cur_n, cur_c);
                        next_row();
                        j += step_r;                            // advance to the slot iteration holding row i + NC
                        const bool carry = j >= rps;
                        j -= carry ? rps : 0;
                        slot += carry ? sl1 : sl0;
                        slot -= slot >= n_slots ? n_slots : 0;
                        bar += carry ? ba1 : ba0;
                        par ^= carry ? fl1 : fl0;
                        if (bar >= nbar) { bar -= nbar; par ^= 1u; }
                        if (own_unit && conv_done < CONV_BATCHES && ++taken == (conv_done + 1) * conv_every) {
                            convert_w_tile<4>(P, Bt, blockIdx.x, warp * 32 + lane, NC * 32, conv_done, 1);
                            ++conv_done;
                        }
                    }
                    if (own_unit && conv_done < CONV_BATCHES)      // short streams: whatever is left
                        convert_w_tile<4>(P, Bt, blockIdx.x, warp * 32 + lane, NC * 32, conv_done, CONV_BATCHES - conv_done);
                    flush();
                };
                using NCt = std::integral_constant<int, TAIL_CONSUMERS>;
#define CIR_CONSUME(PMV) consume(NCt{}, std::integral_constant<int, PMV>{})
                if (gem && P.p_stride) CIR_CONSUME(-1);
                else switch (classify_p(P.pool_mode, p_shared, P.gen_mode)) {
                    case PM_3: CIR_CONSUME(PM_3); break;
                    case PM_2: CIR_CONSUME(PM_2); break;
                    case PM_1: CIR_CONSUME(PM_1); break;
                    case PM_4: CIR_CONSUME(PM_4); break;
                    case PM_MAX: CIR_CONSUME(PM_MAX); break;
                    case PM_MEAN: CIR_CONSUME(PM_MEAN); break;
                    case PM_GENERAL_POLY0 + 1: CIR_CONSUME(PM_GENERAL_POLY0 + 1); break;
                    case PM_GENERAL_POLY0 + 2: CIR_CONSUME(PM_GENERAL_POLY0 + 2); break;
                    default: CIR_CONSUME(PM_GENERAL); break;
                }
#undef CIR_CONSUME
            }
        } else {
            if (own_unit) convert_w_tile<2>(P, Bt, blockIdx.x, tid, TAIL_THREADS);
            // direct loads: every warp takes rows r0 + warp, r0 + warp + 16, ...
            for (int j = warp; j < my_rows; j += TAIL_WARPS) {
                const long long row = r0 + j;
                const float pr = (gem && P.p_stride) ? __ldg(P.p + cur_c) : p_shared;
                const float a = row_reduce<true>(classify_p(P.pool_mode, pr, P.vec_ok ? P.gen_mode : PM_GENERAL), P.x + row * HW, HW, P.vec_ok != 0, lane,
                                               P.eps_gem, pr);
                take(a, cur_n, cur_c);
                next_row();
            }
            flush();
        }
        if (own_unit) conv_unit = blockIdx.x;
    }
    stamp(P, 1);
    if (pool_only) return;

    // phase-B setup (barriers, TMEM, tensor-map prefetch) does not depend on the pooled vectors: do it before the grid
    // barrier instead of on the critical path behind it
    const bool projecting = whiten && (int)blockIdx.x < P.units;
    if (projecting) {
        if (tid == 0) {
            for (int s2 = 0; s2 < P.n_slots; ++s2) {
                mbar_init(&fullB[s2], 1);
                mbar_init(&emptyB[s2], 1);          // freed by tcgen05.commit
            }
            mbar_init(&accB, 1);
            fence_barrier_init();
            prefetch_tmap(&tmHi);
            prefetch_tmap(&tmLo);
        }
        if (warp == 2) {
            tmem_alloc(&tmem_slot, TAIL_NT);
            tmem_relinquish();
        }
        tc_fence_before();
    }
    grid.sync();
    if (projecting) tc_fence_after();
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
            for (int c = tid; c < P.C; c += TAIL_THREADS) {
                float* o = P.out + (size_t)n * P.out_ld + c;
                const float y = __ldcg(g + c) / denom;
                *o = accumulate ? *o + y : y;
            }
        }
        return;
    }

    // ------------------------------------------------------------------ phase B: split-K projection units
    if ((int)blockIdx.x < P.units) {
        constexpr uint32_t idesc = make_idesc_bf16(TAIL_MB, TAIL_NT);
        const uint32_t tmem_acc = tmem_slot;       // allocated before the grid barrier
        int slot = 0;               // ring position of the producer / the MMA issuer (same tile sequence)
        uint32_t phase = 0;
        uint32_t acc_phase = 0;
        for (int unit = blockIdx.x; unit < P.units; unit += gridDim.x) {
            const int nt = unit % P.n_ntiles, ks = unit / P.n_ntiles;
            if (conv_unit != unit) {
                // (more units than CTAs) the previous unit's MMAs have retired: its epilogue waited for them
                __syncthreads();
                convert_w_tile<2>(P, Bt, unit, tid, TAIL_THREADS);
                conv_unit = unit;
                __syncthreads();
            }
            const int kw = min(TAIL_KS, P.C - ks * TAIL_KS);
            const int nkb = (kw + 63) >> 6;
            for (int n0 = 0; n0 < P.N; n0 += TAIL_MB) {
                if (warp == 0) {
                    if (lane == 0) {
                        for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll 1
                            for (int part = 0; part < 2; ++part) {
                                mbar_wait(&emptyB[slot], phase ^ 1u);
                                mbar_arrive_expect_tx(&fullB[slot], TAIL_SLOT_BYTES);
                                tma_load_2d(ring + (size_t)slot * TAIL_SLOT_BYTES, part ? &tmLo : &tmHi, &fullB[slot],
                                            ks * TAIL_KS + kb * 64, n0);
                                if (++slot == P.n_slots) { slot = 0; phase ^= 1u; }
                            }
                        }
                    }
                } else if (warp == 1) {
                    if (lane == 0) {
                        for (int kb = 0; kb < nkb; ++kb) {
                            const int s_hi = slot;
                            const uint32_t ph_hi = phase;
                            if (++slot == P.n_slots) { slot = 0; phase ^= 1u; }
                            const int s_lo = slot;
                            const uint32_t ph_lo = phase;
                            if (++slot == P.n_slots) { slot = 0; phase ^= 1u; }
                            mbar_wait(&fullB[s_hi], ph_hi);
                            mbar_wait(&fullB[s_lo], ph_lo);
                            tc_fence_after();
                            const uint64_t a_hi = make_sw128_desc(smem_u32(ring + (size_t)s_hi * TAIL_SLOT_BYTES));
                            const uint64_t a_lo = make_sw128_desc(smem_u32(ring + (size_t)s_lo * TAIL_SLOT_BYTES));
                            const uint64_t b_hi = make_sw128_desc(smem_u32(Bt + (size_t)(kb * 2) * TAIL_BT_BYTES));
                            const uint64_t b_lo = make_sw128_desc(smem_u32(Bt + (size_t)(kb * 2 + 1) * TAIL_BT_BYTES));
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const uint64_t o = (uint64_t)(kk * 2);
                                umma_bf16(tmem_acc, a_hi + o, b_hi + o, idesc, (uint32_t)((kb | kk) != 0));
                                umma_bf16(tmem_acc, a_hi + o, b_lo + o, idesc, 1u);
                                umma_bf16(tmem_acc, a_lo + o, b_hi + o, idesc, 1u);
                            }
                            umma_commit(&emptyB[s_hi]);
                            umma_commit(&emptyB[s_lo]);
                        }
                        umma_commit(&accB);           // accumulator of this (unit, image block) complete
                    }
                } else if (warp >= 8 && warp < 12) {
                    // epilogue: TMEM -> this K slice's partial projection (fp32, stays in L2 for phase C)
                    mbar_wait(&accB, acc_phase);
                    tc_fence_after();
                    const int quad = warp & 3;
                    const int n = n0 + quad * 32 + lane;
                    float* dst = P.part + ((size_t)ks * P.N + (size_t)(n < P.N ? n : 0)) * P.D_out + nt * TAIL_NT;
#pragma unroll 1
                    for (int cchunk = 0; cchunk < TAIL_NT / 32; ++cchunk) {
                        uint32_t v[32];
                        tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cchunk * 32), v);
                        tmem_ld_wait();
                        if (n < P.N) {
                            const int d0 = nt * TAIL_NT + cchunk * 32;
                            if (d0 + 32 <= P.D_out && (P.D_out & 3) == 0) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    __stcg(reinterpret_cast<float4*>(dst + cchunk * 32 + j),
                                           make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                       __uint_as_float(v[j + 3])));
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (d0 + j < P.D_out) dst[cchunk * 32 + j] = __uint_as_float(v[j]);
                            }
                        }
                    }
                    tc_fence_before();
                }
                acc_phase ^= 1u;
                __syncthreads();          // TMEM is free for the next image block / unit
            }
        }
        tc_fence_before();
        __syncthreads();
        if (warp == 2) tmem_dealloc(tmem_acc, TAIL_NT);
    }

    stamp(P, 3);
    grid.sync();
    stamp(P, 4);

    // ------------------------------------------------------------------ phase C: one CTA per image
    {
        __shared__ float redc[2][TAIL_WARPS];
        // deterministic block sum: fixed shuffle tree, then every thread adds the 16 warp sums in order (ONE barrier; the
        // two sums of an image use different scratch rows, a third use waits for the loop's trailing barrier)
        auto block_sum = [&](float v, int which) -> float {
            v = warp_sum(v);
            if (lane == 0) redc[which][warp] = v;
            __syncthreads();
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < TAIL_WARPS; ++w) t += redc[which][w];
            return t;
        };
        for (int n = blockIdx.x; n < P.N; n += gridDim.x) {
            // issue everything that does not depend on a reduction first, so that ONE L2 round trip covers it all: the pooled
            // vector (first L2N), the bias, then the K-slice partial sums (fixed order)
            constexpr int YR = 8;
            float yreg[YR], breg[YR];
            const bool in_regs = P.D_out <= YR * TAIL_THREADS;
            uint4 gh = make_uint4(0, 0, 0, 0), gl = gh;
            const int c_first = tid * 8;
            if (c_first < P.C) {
                gh = __ldcg(reinterpret_cast<const uint4*>(P.pooled_hi + (size_t)n * P.C + c_first));
                gl = __ldcg(reinterpret_cast<const uint4*>(P.pooled_lo + (size_t)n * P.C + c_first));
            }
#pragma unroll
            for (int i = 0; i < YR; ++i) {
                const int d = tid + i * TAIL_THREADS;
                yreg[i] = 0.0f;
                breg[i] = (P.bias && d < P.D_out) ? __ldg(P.bias + d) : 0.0f;
            }
            // 4 slices x 8 dims = up to 32 independent L2 loads in flight per thread, added in slice order (an
            // accumulate-as-you-load loop serialised 32 L2 latencies: ~6 us on the critical path)
            for (int ks0 = 0; ks0 < P.n_kslices; ks0 += 4) {
                float t[4][YR];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int i = 0; i < YR; ++i) {
                        const int d = tid + i * TAIL_THREADS;
                        t[u][i] = (ks0 + u < P.n_kslices && d < P.D_out)
                                      ? __ldcg(P.part + ((size_t)(ks0 + u) * P.N + n) * P.D_out + d) : 0.0f;
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int i = 0; i < YR; ++i) yreg[i] += t[u][i];
            }
            // first L2N: ||g_n|| from the split pooled vector
            float sg = c_first < P.C ? sumsq8(gh, gl) : 0.0f;
            for (int c = c_first + TAIL_THREADS * 8; c < P.C; c += TAIL_THREADS * 8)
                sg += sumsq8(__ldcg(reinterpret_cast<const uint4*>(P.pooled_hi + (size_t)n * P.C + c)),
                             __ldcg(reinterpret_cast<const uint4*>(P.pooled_lo + (size_t)n * P.C + c)));
            const float inv = 1.0f / (sqrtf(block_sum(sg, 0)) + P.eps_l2);     // W.(g/(|g|+eps)) == (W.g)/(|g|+eps)
            float sy = 0.0f;
#pragma unroll
            for (int i = 0; i < YR; ++i) {
                const int d = tid + i * TAIL_THREADS;
                if (d < P.D_out) {
                    yreg[i] = yreg[i] * inv + breg[i];
                    sy = fmaf(yreg[i], yreg[i], sy);
                    if (P.z_out) P.z_out[(size_t)n * P.D_out + d] = yreg[i];
                }
            }
            for (int d = tid + YR * TAIL_THREADS; d < P.D_out; d += TAIL_THREADS) {
                float acc = 0.0f;
                for (int ks = 0; ks < P.n_kslices; ++ks) acc += __ldcg(P.part + ((size_t)ks * P.N + n) * P.D_out + d);
                const float y = acc * inv + (P.bias ? __ldg(P.bias + d) : 0.0f);
                sy = fmaf(y, y, sy);
                if (P.z_out) P.z_out[(size_t)n * P.D_out + d] = y;
            }
            const float denom = sqrtf(block_sum(sy, 1)) + P.eps_l2;
#pragma unroll
            for (int i = 0; i < YR; ++i) {
                const int d = tid + i * TAIL_THREADS;
                if (d < P.D_out) {
                    float* o = P.out + (size_t)n * P.out_ld + d;
                    const float y = yreg[i] / denom;
                    *o = accumulate ? *o + y : y;
                }
            }
            if (!in_regs) {
                for (int d = tid + YR * TAIL_THREADS; d < P.D_out; d += TAIL_THREADS) {
                    float acc = 0.0f;
                    for (int ks = 0; ks < P.n_kslices; ++ks) acc += __ldcg(P.part + ((size_t)ks * P.N + n) * P.D_out + d);
                    float* o = P.out + (size_t)n * P.out_ld + d;
                    const float y = (acc * inv + (P.bias ? __ldg(P.bias + d) : 0.0f)) / denom;
                    *o = accumulate ? *o + y : y;
                }
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
    // pooled [N, C] (fp32, or bf16 hi + lo) + part [ceil(C / 256)][N][D_out] + debug stamps
    const size_t kslices = ((size_t)C + TAIL_KS - 1) / TAIL_KS;
    *bytes = align_up((size_t)N * C * 4, 256) + align_up(kslices * (size_t)N * D_out * 4, 256) + TAIL_STAMP_BYTES;
    return CIR_OK;
}

static int tail_fwd_impl(const float* x, int N, int C, int H, int W, const float* p, int p_stride,
                         float eps_gem, float eps_l2, int pool_mode, const float* Wt,
                         const float* bias, int D_out, float* out, int out_ld, float* pooled_out, float* z_out, void* workspace,
                         size_t workspace_bytes, unsigned flags, void* stream);

extern "C" int cir_tail_fwd(const float* x, int N, int C, int H, int W, const float* p, int p_stride,
                            float eps_gem, float eps_l2, int pool_mode, const float* Wt,
                            const float* bias, int D_out, float* out, int out_ld, float* pooled_out, void* workspace,
                            size_t workspace_bytes, unsigned flags, void* stream) {
    return tail_fwd_impl(x, N, C, H, W, p, p_stride, eps_gem, eps_l2, pool_mode, Wt, bias, D_out, out, out_ld, pooled_out, nullptr,
                         workspace, workspace_bytes, flags, stream);
}

extern "C" int cir_tail_fwd_train(const float* x, int N, int C, int H, int W, const float* p, int p_stride,
                                  float eps_gem, float eps_l2, int pool_mode, const float* Wt,
                                  const float* bias, int D_out, float* out, int out_ld, float* pooled_out, float* z_out,
                                  void* workspace, size_t workspace_bytes, unsigned flags, void* stream) {
    CIR_REQUIRE(!z_out || !(flags & (CIR_TAIL_POOL_ONLY | CIR_TAIL_NO_WHITEN)), CIR_ERR_INVALID_ARG,
                "cir_tail_fwd_train: z_out is the whitening projection; not available with CIR_TAIL_POOL_ONLY / CIR_TAIL_NO_WHITEN");
    return tail_fwd_impl(x, N, C, H, W, p, p_stride, eps_gem, eps_l2, pool_mode, Wt, bias, D_out, out, out_ld, pooled_out, z_out,
                         workspace, workspace_bytes, flags, stream);
}

static int tail_fwd_impl(const float* x, int N, int C, int H, int W, const float* p, int p_stride,
                         float eps_gem, float eps_l2, int pool_mode, const float* Wt,
                         const float* bias, int D_out, float* out, int out_ld, float* pooled_out, float* z_out, void* workspace,
                         size_t workspace_bytes, unsigned flags, void* stream) {
    CIR_REQUIRE(x && out && N > 0 && C > 0 && H > 0 && W > 0, CIR_ERR_INVALID_ARG,
                "cir_tail_fwd: null pointer or empty shape (N=%d C=%d H=%d W=%d)", N, C, H, W);
    CIR_REQUIRE(pool_mode >= CIR_POOL_GEM && pool_mode <= CIR_POOL_SPOC, CIR_ERR_INVALID_ARG,
                "cir_tail_fwd: unknown pool_mode %d", pool_mode);
    CIR_REQUIRE(pool_mode != CIR_POOL_GEM || p, CIR_ERR_INVALID_ARG, "cir_tail_fwd: GeM needs p");
    CIR_REQUIRE(p_stride == 0 || p_stride == 1, CIR_ERR_INVALID_ARG, "cir_tail_fwd: p_stride must be 0 or 1");
    const bool pool_only = flags & CIR_TAIL_POOL_ONLY;
    const bool whiten = !pool_only && !(flags & CIR_TAIL_NO_WHITEN);
    CIR_REQUIRE(!(pool_only && (flags & CIR_TAIL_ACCUMULATE)), CIR_ERR_INVALID_ARG,
                "cir_tail_fwd: CIR_TAIL_ACCUMULATE applies to descriptors, not to CIR_TAIL_POOL_ONLY");
    if (!whiten) D_out = C;
    CIR_REQUIRE(out_ld >= D_out, CIR_ERR_INVALID_ARG, "cir_tail_fwd: out_ld %d < D_out %d", out_ld, D_out);
    const DeviceInfo& dev = device_info();
    CIR_REQUIRE(dev.coop, CIR_ERR_UNSUPPORTED, "cir_tail_fwd: device lacks cooperative launch");

    TailParams P{};
    P.x = x; P.N = N; P.C = C; P.HW = H * W;
    P.p = p; P.p_stride = p_stride; P.eps_gem = eps_gem; P.eps_l2 = eps_l2; P.pool_mode = pool_mode;
    P.Wt = Wt; P.bias = bias; P.D_out = D_out; P.out = out; P.out_ld = out_ld; P.flags = flags;
    P.pooled_out = pool_only ? nullptr : pooled_out;
    P.z_out = z_out;
    P.vec_ok = ((P.HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    P.bulk_ok = P.vec_ok && (size_t)P.HW * 4 <= (size_t)TAIL_SLOT_BYTES;
    {
        // non-integer exponent: how many of every 4 ex2 go to the FMA pipe (CIR_TAIL_NPOLY = 0 / 1 for experiments; 2..4
        // were measured and lost, see the comment at fold4)
        static const char* dbg = getenv("CIR_TAIL_NPOLY");
        const int npoly = (dbg && dbg[0] >= '0' && dbg[0] <= '2') ? dbg[0] - '0' : TAIL_NPOLY_DEFAULT;
        P.gen_mode = npoly == 0 ? PM_GENERAL : PM_GENERAL_POLY0 + npoly;
    }

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
        P.part = reinterpret_cast<float*>(static_cast<char*>(workspace) + align_up((size_t)N * C * 4, 256));
        if (flags & CIR_TAIL_DEBUG_STAMPS)
            P.stamps = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + need - TAIL_STAMP_BYTES);
    }
    if (whiten) {
        CIR_REQUIRE(Wt, CIR_ERR_INVALID_ARG, "cir_tail_fwd: whitening needs Wt");
        CIR_REQUIRE((C & 7) == 0 && (reinterpret_cast<uintptr_t>(Wt) & 15) == 0, CIR_ERR_UNSUPPORTED,
                    "cir_tail_fwd: whitening needs C %% 8 == 0 and a 16 B aligned Wt (C=%d)", C);
        P.pooled_hi = reinterpret_cast<__nv_bfloat16*>(P.pooled);
        P.pooled_lo = P.pooled_hi + (size_t)N * C;
        P.n_ntiles = (D_out + TAIL_NT - 1) / TAIL_NT;
        P.n_kslices = (C + TAIL_KS - 1) / TAIL_KS;
        P.units = P.n_ntiles * P.n_kslices;
        P.w_bytes = TAIL_W_BYTES;
    }
    // ring: as many 16 KB slots as fit beside the W slice (and at least the phase-B scratch)
    const int static_smem = 2048;     // barriers + reduction scratch (1.5 KB today), with slack
    int slots = (dev.max_smem_optin - static_smem - 1024 - P.w_bytes) / TAIL_SLOT_BYTES;
    if (slots > TAIL_MAX_SLOTS) slots = TAIL_MAX_SLOTS;
    CIR_REQUIRE(slots >= (whiten ? 2 : 1), CIR_ERR_UNSUPPORTED, "cir_tail_fwd: shared memory");
    P.n_slots = slots;
    const size_t smem = (size_t)P.w_bytes + (size_t)slots * TAIL_SLOT_BYTES + 1024 /* 1 KB alignment of the tiles */;
    CUtensorMap tmHi, tmLo;
    memset(&tmHi, 0, sizeof(tmHi));
    memset(&tmLo, 0, sizeof(tmLo));
    if (whiten) {
        int rc = make_tmap_bf16(&tmHi, P.pooled_hi, (uint64_t)N, (uint64_t)C, (uint64_t)C, TAIL_MB);
        if (rc) return rc;
        rc = make_tmap_bf16(&tmLo, P.pooled_lo, (uint64_t)N, (uint64_t)C, (uint64_t)C, TAIL_MB);
        if (rc) return rc;
    }
    static thread_local int attr_set_dev = -1;
    if (attr_set_dev != dev.device) {
        CIR_CHECK_CUDA(cudaFuncSetAttribute(tail_fused_kernel<TAIL_THREADS_LIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            dev.max_smem_optin - static_smem));
        CIR_CHECK_CUDA(cudaFuncSetAttribute(tail_fused_kernel<TAIL_THREADS_HEAVY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            dev.max_smem_optin - static_smem));
        attr_set_dev = dev.device;
    }
    // launch shape (see TAIL_THREADS_*): light rows = MAC / SPoC, or GeM when the caller vouches for an integer exponent
    static const char* dbg_threads = getenv("CIR_TAIL_THREADS");       // experiments: 512 / 640
    bool light = pool_mode != CIR_POOL_GEM || (flags & CIR_TAIL_HINT_INTEGER_P);
    if (dbg_threads) light = atoi(dbg_threads) == TAIL_THREADS_LIGHT;
    const void* kern = light ? (const void*)tail_fused_kernel<TAIL_THREADS_LIGHT> : (const void*)tail_fused_kernel<TAIL_THREADS_HEAVY>;
    const int threads = light ? TAIL_THREADS_LIGHT : TAIL_THREADS_HEAVY;
    void* args[] = {&tmHi, &tmLo, &P};
    if (pool_only) {
        // no grid barrier on this path: a plain launch is enough
        if (light) tail_fused_kernel<TAIL_THREADS_LIGHT><<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(tmHi, tmLo, P);
        else tail_fused_kernel<TAIL_THREADS_HEAVY><<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(tmHi, tmLo, P);
        CIR_CHECK_CUDA(cudaGetLastError());
    } else {
        CIR_CHECK_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(threads), args, smem, static_cast<cudaStream_t>(stream)));
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
