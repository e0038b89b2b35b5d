// Shared declarations of the search path (search.cu <-> topk.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cir {

constexpr int SEARCH_BM = 128;     // queries per tile
constexpr int SEARCH_BN = 256;     // database rows per tile
constexpr int SEARCH_MAX_K = 512;  // largest k of the fused top-k path

struct SearchPlan {
    int mt, nt;     // query tiles, database tiles
    int S, tps;     // splits of the database tiles, tiles per split
    int units;      // mt * S work units
    int Qpad;       // mt * 128
};
SearchPlan plan_search(int Q, long long N, int num_sms);
int search_cap_for_k(int k);       // per-(split, query) candidate list capacity

// final reduction of the candidate lists [S][Qpad][cap] (+ counts [S][Qpad]) to sorted top-k
int launch_topk_select_lists(const unsigned long long* lists, const int* counts, int S, int Qpad, int cap, int Q, int k,
                             float* out_scores, int32_t* out_idx, int out_ld, int32_t idx_offset, cudaStream_t stream,
                             void* const* peers = nullptr, int n_peers = 0, int my_rank = 0, uint32_t flag_target = 0);

// tau0[q] = k-th largest of scores[q, 0:n) (n * 4 + 32 KB of shared memory per block)
int launch_row_kth_largest(const float* scores, int Q, int n, long long ld, int k, float* out, cudaStream_t stream);
constexpr int KTH_MAX_N = 32768;
constexpr int SAMPLE_MAX_ROWS = 32768;   // rows of the threshold pre-pass

}  // namespace cir
