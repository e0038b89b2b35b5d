// Exhaustive inner-product search with a fused per-query top-k (and a dense-score mode).
//
// Replaces  scores = np.dot(V.T, Q); ranks = np.argsort(-scores, axis=0)
//           (scripts/train_globalF.py:733-734, scripts/test.py:246-258) and
//           torch.mm + torch.sort (cirtorch/datasets/globalFeatures/tuples_dataset.py:317-319).
//
// One persistent, warp-specialised kernel per search:
//   warp 0      TMA producer: K-major bf16 tiles of the query block (128 x 64) and of the
//               database block (256 x 64), 128-byte swizzle, 4-stage mbarrier ring
//   warp 1      tcgen05.mma issuer (one thread): 128 x 256 x 16 UMMAs, fp32 accumulators in
//               TMEM, two accumulator stages (2 x 256 columns = all 512 TMEM columns)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue: tcgen05.ld 32 lanes x 32 columns; thread r owns query row r of the
//               tile.  A score enters the row's candidate list only if it beats the row's
//               running threshold tau (a lower bound of the row's k-th best score), so the
//               Q x N score matrix never exists in memory.  When a list is about to
//               overflow, the warp radix-selects its k-th best score, drops everything
//               below it and raises tau.
// Work unit = (query tile m, split s of the database tiles); unit u = s * mt + m, CTA c
// takes units c, c + grid, ... so CTAs running at the same time share database tiles
// through L2.  Every (split, query) list is then reduced to the final sorted top-k by
// topk_select_kernel (topk.cu).
//
// Algorithmic work per launch: 2 * Q * N * Kd flop; N * Kd * 2 bytes of database.
#include "common.cuh"
#include "ptx.cuh"
#include "search.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace cir {

using namespace ptx;

constexpr int BM = SEARCH_BM;        // 128 queries  (TMEM lanes)
constexpr int BN = SEARCH_BN;        // 256 database rows (TMEM columns per accumulator)
constexpr int BK = 64;               // bf16 per K block = one 128 B swizzle row
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2; // 16 KB
constexpr int B_BYTES = BN * BK * 2; // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SEARCH_THREADS = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

enum { MODE_TOPK = 0, MODE_DENSE = 1, MODE_GROUPMAX = 2 };
constexpr int GROUP = 8;            // MODE_GROUPMAX: one value per 8 consecutive database rows

struct SearchParams {
    int Q;
    int N;
    int kblocks;                 // Kd / 64
    int mt, nt, S, tps, units;
    int k, cap;
    unsigned long long* lists;   // [S][Qpad][cap]
    int* counts;                 // [S][Qpad]
    int Qpad;
    const float* tau0;           // optional [Q]: a caller-supplied lower bound of the k-th score
    const int32_t* q_label;      // optional [Q]
    const int32_t* db_label;     // optional [N]: candidates with db_label == q_label are skipped
    float* dense_out;            // MODE_DENSE: [Q, dense_ld]
    long long dense_ld;
    int b_policy;                // database tiles: 0 evict_last, 1 evict_first, 2 evict_normal
    long long piece_stride;      // MODE_GROUPMAX: the problem's rows are 32-row pieces of the matrix, piece i at row i * piece_stride
};

// ---------------------------------------------------------------------------------------
// list compaction (rare path): keep the entries whose score is >= the k-th best score
// ---------------------------------------------------------------------------------------
template <int KPL>
__device__ __noinline__ void compact_row(unsigned long long* L, int n, int k, int cap, int lane, int* out_cnt,
                                         float* out_tau) {
    unsigned long long key[KPL];
#pragma unroll
    for (int r = 0; r < KPL; ++r) {
        const int e = r * 32 + lane;
        key[r] = e < n ? raw_to_key(__ldcg(L + e)) : 0ull;
    }
    // largest T with |{hi >= T}| >= k  ==  the k-th largest ordered score
    uint32_t T = 0;
    for (int b = 31; b >= 0; --b) {
        const uint32_t cand = T | (1u << b);
        int c = 0;
#pragma unroll
        for (int r = 0; r < KPL; ++r) c += ((uint32_t)(key[r] >> 32) >= cand) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= k) T = cand;
    }
    int kept = 0;
#pragma unroll
    for (int r = 0; r < KPL; ++r) kept += ((uint32_t)(key[r] >> 32) >= T && key[r] != 0ull) ? 1 : 0;
    kept = __reduce_add_sync(0xffffffffu, kept);
    unsigned long long T64 = (unsigned long long)T << 32;
    if (kept + 64 > cap) {
        // a flood of equal scores: select on the full (score, index) key -> exactly k survive
        T64 = 0ull;
        for (int b = 63; b >= 0; --b) {
            const unsigned long long cand = T64 | (1ull << b);
            int c = 0;
#pragma unroll
            for (int r = 0; r < KPL; ++r) c += (key[r] >= cand) ? 1 : 0;
            c = __reduce_add_sync(0xffffffffu, c);
            if (c >= k) T64 = cand;
        }
    }
    __syncwarp();
    int base = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < KPL; ++r) {
        const bool keep = key[r] >= T64 && key[r] != 0ull;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) L[base + __popc(bal & lt)] = key_to_raw(key[r]);
        base += __popc(bal);
    }
    __syncwarp();
    *out_cnt = base;
    *out_tau = ordered_to_float(T);
}

__device__ __forceinline__ void compact_dispatch(unsigned long long* L, int n, int k, int cap, int lane, int* c,
                                                 float* t) {
    switch (cap) {
        case 128: compact_row<4>(L, n, k, cap, lane, c, t); break;
        case 256: compact_row<8>(L, n, k, cap, lane, c, t); break;
        case 512: compact_row<16>(L, n, k, cap, lane, c, t); break;
        default: compact_row<32>(L, n, k, cap, lane, c, t); break;
    }
}

// ---------------------------------------------------------------------------------------
// the search kernel
// ---------------------------------------------------------------------------------------
// LABELS: the variant with label exclusion (mining) is a separate instantiation, so the plain search keeps its registers
template <int MODE, bool LABELS>
__global__ void __launch_bounds__(SEARCH_THREADS, 1)
search_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SearchParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                 // [STAGES]  TMA -> MMA
    uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> TMA
    uint64_t* tfull = bars + 2 * STAGES;   // [2]       MMA -> epilogue
    uint64_t* tempty = tfull + 2;          // [2]       epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================ TMA producer
        if (lane == 0) {
            // queries: small and re-read for every database tile -> keep; database: streamed.  With several
            // query tiles the database tiles are shared through L2 by the CTAs of the same split (ncu, 10k x 1M:
            // 9.7 GB of DRAM reads for a 4.1 GB database, 95 % L2 hits); with one query tile they are read once.
            const uint64_t pol_a = policy_evict_last();
            const uint64_t pol_b = P.b_policy == 1 ? policy_evict_first() : (P.b_policy == 2 ? policy_evict_normal() : policy_evict_last());
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
                const int m = u % P.mt, s = u / P.mt;
                const int n0 = s * P.tps, n1 = min(P.nt, n0 + P.tps);
                for (int n = n0; n < n1; ++n) {
                    for (int kb = 0; kb < P.kblocks; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        uint8_t* a = smem + stage * STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                        tma_load_2d_hint(a, &tmA, &full[stage], kb * BK, m * BM, pol_a);
                        if (MODE == MODE_GROUPMAX) {
                            // threshold sample: the 256 rows of a tile are eight 32-row pieces from eight places of the matrix
                            // (tmB has a 32-row box here); a piece is 4 KB of the swizzled tile, 1 KB aligned
#pragma unroll
                            for (int pc = 0; pc < BN / 32; ++pc)
                                tma_load_2d_hint(a + A_BYTES + pc * 32 * 128, &tmB, &full[stage], kb * BK,
                                                 (int)(((long long)n * (BN / 32) + pc) * P.piece_stride), pol_b);
                        } else {
                            tma_load_2d_hint(a + A_BYTES, &tmB, &full[stage], kb * BK, n * BN, pol_b);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
                const int s = u / P.mt;
                const int n0 = s * P.tps, n1 = min(P.nt, n0 + P.tps);
                for (int n = n0; n < n1; ++n) {
                    mbar_wait(&tempty[acc], acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = 0; kb < P.kblocks; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t da = make_sw128_desc(a_addr);
                        const uint64_t db = make_sw128_desc(a_addr + A_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)
                            umma_bf16(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc,
                                      (uint32_t)((kb | kk) != 0));
                        umma_commit(&empty[stage]);     // frees the smem slot once these MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit(&tfull[acc]);           // accumulator complete
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        // ================================================================ epilogue
        const int quad = warp & 3;                  // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
            const int m = u % P.mt, s = u / P.mt;
            const int n0 = s * P.tps, n1 = min(P.nt, n0 + P.tps);
            const int qg = m * BM + row;
            const bool valid = qg < P.Q;
            unsigned long long* L = nullptr;
            int cnt = 0;
            float tau = -INFINITY;
            int qlab = 0;
            if (MODE == MODE_TOPK) {
                L = P.lists + ((size_t)s * P.Qpad + qg) * P.cap;
                if (valid && P.tau0) tau = nextafterf(__ldg(P.tau0 + qg), -INFINITY);
            }
            if (LABELS && MODE != MODE_DENSE && valid) qlab = __ldg(P.q_label + qg);
            for (int n = n0; n < n1; ++n) {
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                const int col0 = n * BN;
                float gmax[MODE == MODE_GROUPMAX ? 32 : 1];          // MODE_GROUPMAX: running maxima over the tile's eight pieces
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(lane_addr + (uint32_t)(acc * BN + c * 32), v);
                    tmem_ld_wait();
                    const int cb = col0 + c * 32;
                    if (MODE == MODE_TOPK) {
                        if (valid) {
                            if (cb + 32 > P.N) {
                                // the ragged last tile: checked path
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const float sc = __uint_as_float(v[j]);
                                    if (sc > tau) {
                                        const int idx = cb + j;
                                        bool ok = idx < P.N;
                                        if (LABELS && ok) ok = __ldg(P.db_label + idx) != qlab;
                                        if (ok) L[cnt++] = ((unsigned long long)(uint32_t)idx << 32) | (unsigned long long)v[j];
                                    }
                                }
                            } else {
                                // Branch-free: one epilogue warp runs alone on its scheduler, so every taken branch costs
                                // its full latency (measured: ~165 cycles per append, which put the epilogue on the MMA's
                                // critical path).  Predicated store + predicated bump of a 32-bit byte offset instead.
                                uint32_t off = (uint32_t)cnt * 8u;
                                if (LABELS) {
                                    // mining: candidates of the query's own cluster are skipped; the label is only loaded
                                    // (predicated) for scores that beat the threshold
                                    const int32_t* lab = P.db_label + cb;
#pragma unroll
                                    for (int j = 0; j < 32; ++j) {
                                        const uint32_t ix = (uint32_t)(cb + j);
                                        asm volatile(
                                            "{\n\t"
                                            ".reg .pred p;\n\t"
                                            ".reg .u64 a;\n\t"
                                            ".reg .s32 l;\n\t"
                                            "setp.gt.f32 p, %2, %3;\n\t"
                                            "@p ld.global.nc.s32 l, [%6];\n\t"
                                            "@p setp.ne.s32 p, l, %7;\n\t"
                                            "@p cvt.u64.u32 a, %0;\n\t"
                                            "@p add.u64 a, a, %1;\n\t"
                                            "@p st.global.v2.b32 [a], {%4, %5};\n\t"
                                            "@p add.u32 %0, %0, 8;\n\t"
                                            "}"
                                            : "+r"(off)
                                            : "l"(L), "f"(__uint_as_float(v[j])), "f"(tau), "r"(v[j]), "r"(ix), "l"(lab + j), "r"(qlab)
                                            : "memory");
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) {
                                        const uint32_t ix = (uint32_t)(cb + j);
                                        asm volatile(
                                            "{\n\t"
                                            ".reg .pred p;\n\t"
                                            ".reg .u64 a;\n\t"
                                            "setp.gt.f32 p, %2, %3;\n\t"
                                            "@p cvt.u64.u32 a, %0;\n\t"
                                            "@p add.u64 a, a, %1;\n\t"
                                            "@p st.global.v2.b32 [a], {%4, %5};\n\t"
                                            "@p add.u32 %0, %0, 8;\n\t"
                                            "}"
                                            : "+r"(off)
                                            : "l"(L), "f"(__uint_as_float(v[j])), "f"(tau), "r"(v[j]), "r"(ix)
                                            : "memory");
                                    }
                                }
                                cnt = (int)(off >> 3);
                            }
                        }
                        const unsigned need = __ballot_sync(0xffffffffu, valid && (cnt + 32 > P.cap));
                        if (need) {
                            __syncwarp();
                            unsigned todo = need;
                            while (todo) {
                                const int l = __ffs(todo) - 1;
                                todo &= todo - 1;
                                const int n_l = __shfl_sync(0xffffffffu, cnt, l);
                                unsigned long long* Ll =
                                    P.lists + ((size_t)s * P.Qpad + (m * BM + quad * 32 + l)) * P.cap;
                                int nc;
                                float nt_;
                                compact_dispatch(Ll, n_l, P.k, P.cap, lane, &nc, &nt_);
                                if (lane == l) { cnt = nc; tau = fmaxf(tau, nt_); }
                            }
                        }
                    } else if (MODE == MODE_GROUPMAX) {
                        // Threshold pre-pass: of every 8 rows only the maximum leaves the SM (N % 256 == 0 here).  Each
                        // maximum is the score of a distinct row, so the k-th largest of them is a lower bound of the query's
                        // k-th best score -- at 1/8 of the bytes and 1/8 of the selection work of the dense block.  A group is
                        // the SAME offset j in the tile's eight 32-row pieces, i.e. 8 rows from 8 different places of the
                        // matrix: the maximum of 8 neighbouring rows of a centre-sorted database is hardly an extreme value
                        // (the rows of a centre score alike), and the bound then admitted 8 x more candidates.
                        if (valid) {
                            if (LABELS) {                   // mining: rows of the query's own cluster do not count
                                const int32_t* lab = P.db_label + ((long long)n * (BN / 32) + c) * P.piece_stride;     // the sampled piece's rows
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (__ldg(lab + j) == qlab) v[j] = 0xff800000u;      // -inf
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                gmax[j] = c == 0 ? __uint_as_float(v[j]) : fmaxf(gmax[j], __uint_as_float(v[j]));
                            if (c == BN / 32 - 1) {
                                float* o = P.dense_out + (long long)qg * P.dense_ld + col0 / GROUP;
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<float4*>(o + j) = make_float4(gmax[j], gmax[j + 1], gmax[j + 2], gmax[j + 3]);
                            }
                        }
                    } else {
                        if (valid) {
                            float* o = P.dense_out + (long long)qg * P.dense_ld + cb;
                            if (cb + 32 <= P.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<float4*>(o + j) =
                                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (cb + j < P.N) o[j] = __uint_as_float(v[j]);
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            if (MODE == MODE_TOPK) P.counts[(size_t)s * P.Qpad + qg] = valid ? cnt : 0;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// K-major bf16 matrix [rows, Kd] (row pitch = Kd elements) -> box {64, box_rows}, 128 B swizzle
static int make_tmap(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t Kd, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    CIR_REQUIRE(fn, CIR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[2] = {Kd, rows};
    cuuint64_t strides[1] = {Kd * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CIR_REQUIRE(r == CUDA_SUCCESS, CIR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CIR_OK;
}

// candidate list capacity: room for ~3k appends beyond the k kept, so that after the threshold warm start a
// list almost never has to be compacted inside the GEMM kernel (compaction stalls the MMA pipeline)
int search_cap_for_k(int k) {
    int cap = 128;
    while (cap < k + 96 || (cap < 4 * k && cap < 1024)) cap <<= 1;
    return cap;
}

SearchPlan plan_search(int Q, long long N, int num_sms) {
    SearchPlan p{};
    p.mt = (Q + BM - 1) / BM;
    p.nt = (int)((N + BN - 1) / BN);
    p.Qpad = p.mt * BM;
    // choose tiles-per-split: minimise waves * tps (+ a per-unit overhead of ~1 tile)
    long long best_cost = -1;
    int best_tps = 1;
    for (int tps = 1; tps <= p.nt; ++tps) {
        const int S = (p.nt + tps - 1) / tps;
        const long long units = (long long)p.mt * S;
        const long long waves = (units + num_sms - 1) / num_sms;
        const long long cost = waves * (tps + 1);
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && tps > best_tps)) {
            best_cost = cost;
            best_tps = tps;
        }
        if (S == 1) break;
    }
    p.tps = best_tps;
    p.S = (p.nt + best_tps - 1) / best_tps;
    p.units = p.mt * p.S;
    return p;
}

// N = rows of the problem; MODE_GROUPMAX (the threshold sample): the matrix behind `db` has N_map >= N rows and the problem's
// rows are its 32-row pieces i * piece_stride .. + 31 (piece_stride = 32: the first N rows)
static int launch_search(int mode, const void* q, int Q, const void* db, long long N, int Kd, SearchParams& P,
                         const SearchPlan& plan, cudaStream_t stream, long long N_map = 0, long long piece_stride = 32) {
    const DeviceInfo& dev = device_info();
    CIR_REQUIRE(dev.max_smem_optin >= SMEM_BYTES, CIR_ERR_UNSUPPORTED, "search: device offers %d B shared memory, need %d",
                dev.max_smem_optin, SMEM_BYTES);
    CUtensorMap tmA, tmB;
    int rc = make_tmap(&tmA, q, (uint64_t)Q, (uint64_t)Kd, BM);
    if (rc) return rc;
    rc = make_tmap(&tmB, db, (uint64_t)(N_map > N ? N_map : N), (uint64_t)Kd, mode == MODE_GROUPMAX ? 32 : BN);
    if (rc) return rc;
    P.Q = Q;
    P.N = (int)N;
    P.kblocks = Kd / BK;
    P.mt = plan.mt; P.nt = plan.nt; P.S = plan.S; P.tps = plan.tps; P.units = plan.units; P.Qpad = plan.Qpad;
    P.b_policy = plan.mt == 1 ? 1 : 0;
    P.piece_stride = piece_stride;
    static thread_local int attr_dev = -1;
    if (attr_dev != dev.device) {
        CIR_CHECK_CUDA(cudaFuncSetAttribute(search_kernel<MODE_TOPK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CIR_CHECK_CUDA(cudaFuncSetAttribute(search_kernel<MODE_TOPK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CIR_CHECK_CUDA(cudaFuncSetAttribute(search_kernel<MODE_DENSE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CIR_CHECK_CUDA(cudaFuncSetAttribute(search_kernel<MODE_GROUPMAX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CIR_CHECK_CUDA(cudaFuncSetAttribute(search_kernel<MODE_GROUPMAX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_dev = dev.device;
    }
    const int grid = plan.units < dev.num_sms ? plan.units : dev.num_sms;
    const bool labels = P.db_label != nullptr;
    if (mode == MODE_TOPK) {
        if (labels) search_kernel<MODE_TOPK, true><<<grid, SEARCH_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, P);
        else search_kernel<MODE_TOPK, false><<<grid, SEARCH_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, P);
    } else if (mode == MODE_GROUPMAX) {
        if (labels) search_kernel<MODE_GROUPMAX, true><<<grid, SEARCH_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, P);
        else search_kernel<MODE_GROUPMAX, false><<<grid, SEARCH_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, P);
    } else {
        search_kernel<MODE_DENSE, false><<<grid, SEARCH_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, P);
    }
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}

static int check_operands(const char* who, const void* q, int Q, const void* db, long long N, int Kd) {
    CIR_REQUIRE(q && db, CIR_ERR_INVALID_ARG, "%s: null operand", who);
    CIR_REQUIRE(Q > 0 && N > 0, CIR_ERR_INVALID_ARG, "%s: empty problem (Q=%d N=%lld)", who, Q, N);
    CIR_REQUIRE(N <= 0x7fffff00ll, CIR_ERR_UNSUPPORTED, "%s: N=%lld exceeds the int32 index range", who, N);
    CIR_REQUIRE(Kd > 0 && Kd % BK == 0, CIR_ERR_INVALID_ARG, "%s: Kd=%d must be a positive multiple of 64", who, Kd);
    CIR_REQUIRE(((uintptr_t)q & 15) == 0 && ((uintptr_t)db & 15) == 0, CIR_ERR_INVALID_ARG,
                "%s: operands must be 16 B aligned", who);
    return CIR_OK;
}

}  // namespace cir

using namespace cir;

// Warm start of the running thresholds: the k-th best score of every query over (a subset of) n0 database rows --
// n0 / 256 whole tiles spread evenly over the matrix -- is a valid lower bound of its k-th best over all rows.  It removes almost all list compactions
// and most appends (measured 53.6 ms -> 30.9 ms on 10k x 1M with an exact tau0).
static int sample_rows(int Q, long long N, int k) {
    // Large databases: ~N/32 rows (3 % extra scan), a power of two in [2048, 32768], with at least 4 k groups of 8 rows so
    // that the k-th largest group maximum stays a tight bound (10k x 1M: n0 = 32768 -> 32.3 ms; an exact threshold: 31.5 ms).
    // Mid-size databases (8,192 .. 65,535 rows, more than one query tile: mining, 2,000 x 20,000): 2 k groups (>= 1024
    // rows, <= N / 4) -- without a threshold every (split, query) list starts by taking every score and compacting
    // every ~400 appends, which cost more than the whole GEMM.  Small or single-tile problems skip the pre-pass.
    if (N < 8192 || (Q <= SEARCH_BM && N < 65536)) return 0;
    static const char* dbg = getenv("CIR_DEBUG_SAMPLE_ROWS");        // experiments only
    int n0 = 2048;
    if (dbg && atoi(dbg) >= 1024 && atoi(dbg) <= SAMPLE_MAX_ROWS) {
        n0 = atoi(dbg) / 256 * 256;
        if ((long long)n0 * 2 > N) n0 = 1024;
    } else if (N >= 65536) {
        // measured at 1M rows: 10k queries 32768 -> 31.4 ms, 16384 -> 35.1, 8192 -> 36.9 (appends stall the MMA);
        // 1,024 queries 32768 -> 3.38 ms, 16384 -> 3.16, 8192 -> 3.20: half the sample below ~2k queries
        const long long frac = Q > 2048 ? 32 : 64;
        while (n0 * 2 <= SAMPLE_MAX_ROWS && (long long)n0 * 2 * frac <= N + N / 2) n0 *= 2;
        while (k > n0 / GROUP / 4 && n0 * 2 <= SAMPLE_MAX_ROWS && (long long)n0 * 2 * 4 <= N) n0 *= 2;
    } else {
        n0 = 1024;
        while (k > n0 / GROUP / 2 && (long long)n0 * 2 * 4 <= N) n0 *= 2;
    }
    if (k > n0 / GROUP / 2) return 0;
    return n0;
}

struct SearchWs {
    size_t lists, counts, tau, dense, total;
};
static SearchWs search_ws_layout(const SearchPlan& plan, int Q, int cap, int n0) {
    SearchWs w{};
    w.lists = 0;
    w.counts = w.lists + align_up((size_t)plan.S * plan.Qpad * cap * 8, 256);
    w.tau = w.counts + align_up((size_t)plan.S * plan.Qpad * 4, 256);
    w.dense = w.tau + align_up((size_t)plan.Qpad * 4, 256);
    w.total = w.dense + align_up((size_t)Q * (n0 / GROUP) * 4, 256);
    return w;
}

// Host-only view of the planner (no device needed): how a search of Q queries over N rows is cut into work units on a GPU
// with num_sms SMs, the candidate-list capacity for k and the size of the threshold sample.
extern "C" int cir_search_plan(int Q, int64_t N, int k, int num_sms, int32_t* out8) {
    CIR_REQUIRE(out8 && Q > 0 && N > 0 && k >= 1 && k <= SEARCH_MAX_K && num_sms > 0, CIR_ERR_INVALID_ARG,
                "cir_search_plan: bad arguments (Q=%d N=%lld k=%d num_sms=%d)", Q, (long long)N, k, num_sms);
    const SearchPlan p = plan_search(Q, N, num_sms);
    out8[0] = p.mt; out8[1] = p.nt; out8[2] = p.S; out8[3] = p.tps; out8[4] = p.units; out8[5] = p.Qpad;
    out8[6] = search_cap_for_k(k);
    out8[7] = sample_rows(Q, N, k);
    return CIR_OK;
}

extern "C" int cir_search_workspace_bytes(int Q, int64_t N, int Kd, int k, size_t* bytes) {
    (void)Kd;
    CIR_REQUIRE(bytes && Q > 0 && N > 0 && k >= 1 && k <= SEARCH_MAX_K, CIR_ERR_INVALID_ARG,
                "cir_search_workspace_bytes: bad arguments (Q=%d N=%lld k=%d, k <= %d)", Q, (long long)N, k, SEARCH_MAX_K);
    const SearchPlan plan = plan_search(Q, N, device_info().num_sms);
    *bytes = search_ws_layout(plan, Q, search_cap_for_k(k), sample_rows(Q, N, k)).total;
    return CIR_OK;
}

static int search_topk_impl(const void* q, int Q, const void* db, int64_t N, int Kd, int k, const float* tau0,
                            const int32_t* q_label, const int32_t* db_label, float* out_scores, int32_t* out_idx,
                            int32_t idx_offset, void* workspace, size_t workspace_bytes, unsigned flags, void* stream,
                            void* const* peers, int n_peers, int my_rank, uint32_t flag_target = 0);

extern "C" int cir_search_topk(const void* q, int Q, const void* db, int64_t N, int Kd, int k, const float* tau0,
                               const int32_t* q_label, const int32_t* db_label, float* out_scores, int32_t* out_idx,
                               int32_t idx_offset, void* workspace, size_t workspace_bytes, unsigned flags,
                               void* stream) {
    CIR_REQUIRE(out_scores && out_idx, CIR_ERR_INVALID_ARG, "cir_search_topk: null output");
    return search_topk_impl(q, Q, db, N, Kd, k, tau0, q_label, db_label, out_scores, out_idx, idx_offset, workspace,
                            workspace_bytes, flags, stream, nullptr, 0, 0);
}

extern "C" int cir_search_topk_exchange(const void* q, int Q, const void* db, int64_t N, int Kd, int k, int32_t idx_offset,
                                        void* const* peer_bufs, int n_peers, int my_rank, void* workspace,
                                        size_t workspace_bytes, unsigned flags, void* stream) {
    CIR_REQUIRE(peer_bufs && n_peers >= 1 && my_rank >= 0 && my_rank < n_peers, CIR_ERR_INVALID_ARG,
                "cir_search_topk_exchange: bad peer arguments");
    for (int g = 0; g < n_peers; ++g) CIR_REQUIRE(peer_bufs[g], CIR_ERR_INVALID_ARG, "cir_search_topk_exchange: null peer buffer");
    return search_topk_impl(q, Q, db, N, Kd, k, nullptr, nullptr, nullptr, nullptr, nullptr, idx_offset, workspace,
                            workspace_bytes, flags, stream, peer_bufs, n_peers, my_rank);
}

extern "C" int cir_search_topk_exchange_merge(const void* q, int Q, const void* db, int64_t N, int Kd, int k, int32_t idx_offset,
                                              void* const* peer_bufs, int n_peers, int my_rank, uint32_t arrivals,
                                              float* out_scores, int32_t* out_idx, void* workspace, size_t workspace_bytes,
                                              unsigned flags, void* stream) {
    CIR_REQUIRE(peer_bufs && n_peers >= 1 && my_rank >= 0 && my_rank < n_peers && arrivals != 0 && out_scores && out_idx,
                CIR_ERR_INVALID_ARG, "cir_search_topk_exchange_merge: bad peer arguments");
    for (int g = 0; g < n_peers; ++g)
        CIR_REQUIRE(peer_bufs[g], CIR_ERR_INVALID_ARG, "cir_search_topk_exchange_merge: null peer buffer");
    // every selection block waits for its peers: all Q blocks must be resident at once.  Checked before anything is launched.
    CIR_REQUIRE(Q <= device_info().num_sms && (long long)n_peers * k <= 8192, CIR_ERR_UNSUPPORTED,
                "cir_search_topk_exchange_merge: the fused merge needs Q <= %d queries and n_peers * k <= 8192 (Q=%d n_peers=%d k=%d)",
                device_info().num_sms, Q, n_peers, k);
    return search_topk_impl(q, Q, db, N, Kd, k, nullptr, nullptr, nullptr, out_scores, out_idx, idx_offset, workspace,
                            workspace_bytes, flags, stream, peer_bufs, n_peers, my_rank, arrivals);
}

static int search_topk_impl(const void* q, int Q, const void* db, int64_t N, int Kd, int k, const float* tau0,
                            const int32_t* q_label, const int32_t* db_label, float* out_scores, int32_t* out_idx,
                            int32_t idx_offset, void* workspace, size_t workspace_bytes, unsigned flags, void* stream,
                            void* const* peers, int n_peers, int my_rank, uint32_t flag_target) {
    int rc = check_operands("cir_search_topk", q, Q, db, N, Kd);
    if (rc) return rc;
    CIR_REQUIRE(k >= 1 && k <= SEARCH_MAX_K, CIR_ERR_UNSUPPORTED, "cir_search_topk: k=%d outside [1, %d]", k, SEARCH_MAX_K);
    CIR_REQUIRE((q_label == nullptr) == (db_label == nullptr), CIR_ERR_INVALID_ARG,
                "cir_search_topk: q_label and db_label go together");
    size_t need = 0;
    cir_search_workspace_bytes(Q, N, Kd, k, &need);
    CIR_REQUIRE(workspace && workspace_bytes >= need, CIR_ERR_WORKSPACE, "cir_search_topk: workspace %zu < %zu bytes",
                workspace_bytes, need);
    CIR_REQUIRE(((uintptr_t)workspace & 15) == 0, CIR_ERR_INVALID_ARG, "cir_search_topk: workspace must be 16 B aligned");
    const SearchPlan plan = plan_search(Q, N, device_info().num_sms);
    const int n0_ws = sample_rows(Q, N, k);
    const SearchWs w = search_ws_layout(plan, Q, search_cap_for_k(k), n0_ws);
    char* ws = static_cast<char*>(workspace);
    const int n0 = (tau0 || (flags & CIR_SEARCH_NO_PREPASS)) ? 0 : n0_ws;
    if (n0 > 0) {
        // pre-pass: maxima of every 8 consecutive rows of the first n0 rows, then the k-th largest of them per query
        float* dense = reinterpret_cast<float*>(ws + w.dense);
        float* tau = reinterpret_cast<float*>(ws + w.tau);
        SearchParams D{};
        D.dense_out = dense;
        D.dense_ld = n0 / GROUP;
        D.q_label = q_label;          // mining: the threshold only counts rows the query may take
        D.db_label = db_label;
        // The sample = n0 / 32 pieces of 32 consecutive rows spread evenly over the matrix, so that the threshold does not
        // depend on how the rows are ordered.  Whole 256-row tiles were not fine enough: on a shard stored centre by
        // centre (125k rows of an 8-GPU search: a 4,096-row sample = 16 tiles = ~40 distinct centres) the k-th largest
        // of 512 group maxima was the ~9th best CENTRE, which admitted 20 % of the rows (5.11 vs 3.96 ms).
        const long long pieces = n0 / 32;
        const long long stride = (flags & CIR_SEARCH_SAMPLE_FIRST_ROWS) ? 32 : N / pieces;
        rc = launch_search(MODE_GROUPMAX, q, Q, db, n0, Kd, D, plan_search(Q, n0, device_info().num_sms),
                           static_cast<cudaStream_t>(stream), N, stride < 32 ? 32 : stride);
        if (rc) return rc;
        rc = launch_row_kth_largest(dense, Q, n0 / GROUP, n0 / GROUP, k, tau, static_cast<cudaStream_t>(stream));
        if (rc) return rc;
        tau0 = tau;
    }
    SearchParams P{};
    P.k = k;
    P.cap = search_cap_for_k(k);
    P.lists = reinterpret_cast<unsigned long long*>(ws + w.lists);
    P.counts = reinterpret_cast<int*>(ws + w.counts);
    P.tau0 = tau0;
    P.q_label = q_label;
    P.db_label = db_label;
    rc = launch_search(MODE_TOPK, q, Q, db, N, Kd, P, plan, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    return launch_topk_select_lists(P.lists, P.counts, plan.S, plan.Qpad, P.cap, Q, k, out_scores, out_idx, k, idx_offset,
                                    static_cast<cudaStream_t>(stream), peers, n_peers, my_rank, flag_target);
}

extern "C" int cir_scores_dense(const void* q, int Q, const void* db, int64_t N, int Kd, float* out, int64_t ld_out,
                                void* stream) {
    int rc = check_operands("cir_scores_dense", q, Q, db, N, Kd);
    if (rc) return rc;
    CIR_REQUIRE(out && ld_out >= N, CIR_ERR_INVALID_ARG, "cir_scores_dense: bad output (ld_out=%lld N=%lld)",
                (long long)ld_out, (long long)N);
    const SearchPlan plan = plan_search(Q, N, device_info().num_sms);
    SearchParams P{};
    P.dense_out = out;
    P.dense_ld = ld_out;
    return launch_search(MODE_DENSE, q, Q, db, N, Kd, P, plan, static_cast<cudaStream_t>(stream));
}
