/*
 * cir_b200.h -- C-ABI of the B200-native cirtorch global-descriptor retrieval hot path.
 *
 * One shared library (libcir_b200.so), extern "C", plain pointers and sizes, no torch or
 * C++ types.  The reference (/root/reference, 100 % Python) has no FFI; each entry point
 * below replaces the arithmetic of the reference function cited next to it, and the
 * Python host layer (cirtorch_b200/) binds it with ctypes behind the reference's own
 * module / function names.
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller;
 *   - matrices are ROW-MAJOR "one descriptor per row" ([rows, D], D contiguous): this is
 *     the physical layout behind cirtorch's public D x N column-descriptor tensors
 *     (global_head.py:67 returns a permute view of a contiguous N x D buffer);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream), allocates nothing persistent and keeps no state besides
 *     cached device attributes and the driver entry point used to encode TMA maps;
 *   - workspaces are caller-owned; query the size with the matching *_workspace_bytes;
 *   - return value: 0 = ok, negative = CIR_ERR_*; cir_last_error() returns a
 *     thread-local, human-readable message for the last failing call.
 */
#ifndef CIR_B200_H
#define CIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CIR_OK 0
#define CIR_ERR_INVALID_ARG (-1)   /* bad shape / alignment / null pointer            */
#define CIR_ERR_WORKSPACE (-2)     /* workspace missing or too small                  */
#define CIR_ERR_CUDA (-3)          /* a CUDA runtime / driver call failed             */
#define CIR_ERR_UNSUPPORTED (-4)   /* shape outside what the kernels implement        */

/* pooling modes of cir_tail_fwd (cirtorch/modules/pools.py POOLING_LAYERS) */
#define CIR_POOL_GEM 0   /* GeM   pools.py:30-38 (p shared) / GeMmp :43-54 (p per channel) */
#define CIR_POOL_MAC 1   /* MAC   pools.py:10-16 */
#define CIR_POOL_SPOC 2  /* SPoC  pools.py:20-26 */

/* flags of cir_tail_fwd */
#define CIR_TAIL_NO_WHITEN 1u   /* globalHead.forward(x, do_whitening=False)            */
#define CIR_TAIL_POOL_ONLY 2u   /* stop after pooling: GeM.forward alone (no L2N)       */
#define CIR_TAIL_ACCUMULATE 4u  /* out += descriptor instead of out = descriptor: the sum over scales of the multi-scale
                                   mean, ImageRetrievalNet.forward cirtorch/models/GF_net.py:74-92 (divide by S afterwards) */
#define CIR_TAIL_HINT_INTEGER_P 8u /* performance hint, never changes the result: the caller knows that the GeM exponent is
                                    * 1, 2, 3 or 4 (p lives on the device; the library does not read it on the host).  Rows are
                                    * then cheap and the launch uses 512 threads / 120 registers; without the hint 640 / 96,
                                    * which suits the two MUFU operations per element of a non-integer exponent */
#define CIR_TAIL_DEBUG_STAMPS 0x80000000u /* profiling aid: every CTA writes %globaltimer at its phase
                                   boundaries into the last 64 KB of the workspace ([cta][8] u64) */

/* flags of cir_search_topk */
#define CIR_SEARCH_SORTED 0u      /* (default) lists sorted by (score desc, index asc)  */
#define CIR_SEARCH_NO_PREPASS 1u  /* skip the threshold warm start (sample of the rows)   */
#define CIR_SEARCH_SAMPLE_FIRST_ROWS 2u /* threshold sample = the first rows instead of 32-row pieces spread over the matrix */

const char* cir_last_error(void);
int cir_version(void);            /* 100 = round 1, 200 = round 2 (cir_tail_fwd gained pooled_out; section 7 added), 210 = cir_tail_fwd_train, cir_search_topk_exchange_merge */
/* number of kernels launched by this library on the calling thread since the last reset
 * (bench.py's gpu_launches counter) */
int64_t cir_launch_count(int reset);

/* ------------------------------------------------------------------------------------
 * 1. Fused descriptor tail
 *    replaces GeM.forward            cirtorch/modules/pools.py:37-38
 *             L2N.forward            cirtorch/modules/normalizations.py:15-16
 *             globalHead.forward     cirtorch/modules/heads/global_head.py:52-67
 *    out[n, :] = L2N( L2N(pool(x[n])) . W^T + b )         (row n = descriptor of image n)
 *
 *    x        device, [N, C, H, W] fp32 contiguous (NCHW)
 *    p        device, GeM exponent: 1 float (p_stride = 0) or C floats (p_stride = 1);
 *             read on the device, so a learnable Parameter needs no host sync
 *    Wt       device, [D_out, C] fp32 row-major (= nn.Linear.weight), may be NULL with
 *             CIR_TAIL_NO_WHITEN / CIR_TAIL_POOL_ONLY (then D_out must equal C)
 *    bias     device, [D_out] or NULL (= zeros)
 *    out      device, [N, out_ld] fp32, out_ld >= D_out
 *    pooled_out  optional device [N, C] fp32: the pooled values (before the first L2N), which the backward pass
 *             needs (cir_gem_bwd); NULL = not wanted
 *    One cooperative launch: phase A streams x once (HBM-bound), a grid barrier, phase B
 *    does the projection from L2-resident pooled vectors, a second barrier, final L2N.
 * ------------------------------------------------------------------------------------ */
int cir_tail_workspace_bytes(int N, int C, int D_out, size_t* bytes);
int cir_tail_fwd(const float* x, int N, int C, int H, int W,
                 const float* p, int p_stride, float eps_gem, float eps_l2, int pool_mode,
                 const float* Wt, const float* bias, int D_out,
                 float* out, int out_ld, float* pooled_out,
                 void* workspace, size_t workspace_bytes, unsigned flags, void* stream);

/* The same launch for the training forward: additionally stores z = W . L2N(pool(x)) + b, the projection BEFORE the last L2N,
 * into z_out [N, D_out] fp32 (NULL = not wanted) -- the backward of the final L2N and of the Linear need it, and
 * recomputing it is a 64 x 2048 x 2048 fp32 GEMM. */
int cir_tail_fwd_train(const float* x, int N, int C, int H, int W,
                       const float* p, int p_stride, float eps_gem, float eps_l2, int pool_mode,
                       const float* Wt, const float* bias, int D_out,
                       float* out, int out_ld, float* pooled_out, float* z_out,
                       void* workspace, size_t workspace_bytes, unsigned flags, void* stream);

/* Backward of the GeM pooling over the feature map (training, scripts/train_globalF.py:480-488; the autograd of
 * pools.py:37-38): dx[n,c,:,:] = dg[n,c] * g[n,c]^(1-p) * max(x,eps)^(p-1) / (H W) where x >= eps, and, when S != NULL,
 * S[n,c] = sum_hw max(x,eps)^p ln max(x,eps) (for dL/dp).  g = the forward's pooled values [N, C], dg = their gradient.
 * One pass over x (read) and dx (write); everything else of the tail's backward is [N, C]-sized. */
int cir_gem_bwd(const float* x, int N, int C, int H, int W, const float* p, int p_stride, float eps_gem,
                const float* g, const float* dg, float* dx, float* S, void* stream);

/* ------------------------------------------------------------------------------------
 * 2. Bias + L2N of projected descriptors
 *    replaces the tail of whitenapply   cirtorch/utils/whiten.py:8-12
 *    whitenapply(X, m, P) = L2N(P[:dims] (X - m)) = L2N(X . W^T + b), W = P[:dims], b = -W m:
 *    the host packs X and W as bf16x3 operands (cir_pack_bf16, n_split = 3), runs the
 *    projection through cir_scores_dense (tcgen05 GEMM, ~fp32-accurate) and finishes here:
 *    out[n, :] = (X[n, :] + bias) / (|| X[n, :] + bias ||_2 + eps_l2); eps_l2 < 0 skips the L2N,
 *    bias may be NULL.  X [N, C] fp32 (ldx), out [N, out_ld]; in place allowed.
 * ------------------------------------------------------------------------------------ */
int cir_bias_l2n_rows(const float* X, int64_t N, int C, int64_t ldx, const float* bias,
                      float eps_l2, float* out, int64_t out_ld, void* stream);

/* PowerLaw.forward (cirtorch/modules/normalizations.py:19-27): out = sign(x + eps) * sqrt(|x + eps|), elementwise */
int cir_powerlaw(const float* x, int64_t n, float eps, float* out, void* stream);

/* Regional pooling: replaces the region loop of Rpool.roipool (cirtorch/modules/pools.py:126-167) and the max-pools of
 * RMAC.forward (pools.py:64-113).  x [N, C, H, W] fp32 device (read once); regions [R][4] int32 in HOST memory =
 * (row0, col0, height, width) of every region (inside the map, R <= 64); pool_mode / p / p_stride / eps as in
 * cir_tail_fwd, applied per region; out [N][R][C] fp32 device (row = one region descriptor). */
int cir_region_pool(const float* x, int N, int C, int H, int W, const int32_t* regions, int R,
                    const float* p, int p_stride, float eps, int pool_mode, float* out, void* stream);

/* row-wise L2N in place or out of place: L2N.forward on [N, C] rows (normalizations.py:15) */
int cir_l2n_rows(const float* X, int64_t N, int C, int64_t ldx, float eps,
                 float* out, int64_t out_ld, void* stream);

/* ------------------------------------------------------------------------------------
 * 3. Exhaustive inner-product search with fused per-query top-k
 *    replaces scores = np.dot(V.T, Q); ranks = np.argsort(-scores, axis=0)
 *                                         scripts/train_globalF.py:733-734, scripts/test.py:246-258
 *             torch.mm + torch.sort       tuples_dataset.py:317-319
 *    The score matrix never reaches HBM: a TMA-fed tcgen05 GEMM (bf16 in, fp32 accumulate
 *    in TMEM) whose epilogue keeps a running top-k per query.
 *
 *    cir_pack_bf16: fp32 rows -> K-major bf16 rows (the search operand format).
 *      n_split = 1: dst[r, 0:D)         = bf16(src[r])
 *      n_split = 3: "bf16x3" operands, K' = 3 D, ~fp32-accurate scores:
 *           role 0 (query side):  [hi, hi, lo]      role 1 (database side): [hi, lo, hi]
 *           so  q'.d' = hi.hi + hi.lo + lo.hi
 *      D is padded with zeros to a multiple of 64 per segment (dst_ld = n_split*round_up(D,64)).
 * ------------------------------------------------------------------------------------ */
int cir_pack_bf16(const float* src, int64_t rows, int D, int64_t src_ld,
                  void* dst, int64_t dst_ld, int n_split, int role, void* stream);

/* host-only planner view (no device needed): out8 = {query tiles, database tiles, splits, tiles per split, work units,
 * padded query count, candidate-list capacity for k, rows of the threshold pre-pass (0 = none)} for a GPU with num_sms SMs */
int cir_search_plan(int Q, int64_t N, int k, int num_sms, int32_t* out8);
int cir_search_workspace_bytes(int Q, int64_t N, int Kd, int k, size_t* bytes);
/*  q   device bf16 [Q, Kd] (ld = Kd), db device bf16 [N, Kd]; Kd % 64 == 0; 1 <= k <= 512
 *  tau0      optional device [Q] fp32: a caller-supplied lower bound of each query's k-th
 *            best score (e.g. from a previous search of a subset); NULL = none
 *  q_label / db_label  optional device int32 [Q] / [N] (both or neither): database rows whose
 *            label equals the query's label are skipped -- the "same cluster as the query"
 *            exclusion of tuples_dataset.py:330-339 applied before the top-k
 *  out_scores [Q, k] fp32, out_idx [Q, k] int32 sorted by (score desc, index asc);
 *  idx_offset is added to every index; -inf / -1 pad when fewer than k rows qualify */
int cir_search_topk(const void* q, int Q, const void* db, int64_t N, int Kd, int k,
                    const float* tau0, const int32_t* q_label, const int32_t* db_label,
                    float* out_scores, int32_t* out_idx, int32_t idx_offset,
                    void* workspace, size_t workspace_bytes, unsigned flags, void* stream);

/* Sharded search with the exchange fused into the final selection: instead of returning its lists, rank `my_rank`
 * stores them into slot `my_rank` of EVERY peer's exchange buffer -- peer_bufs[g] is a device pointer (mapped into
 * this process, e.g. by torch symmetric memory / CUDA IPC over NVLink) to an int32 buffer [n_peers][2][Q][k]
 * (fp32 score bits, then indices).  After a cross-GPU barrier every rank merges its own buffer with
 * cir_topk_merge(scores = buf, idx = buf + Q*k, G = n_peers, g_stride = 2*Q*k).  peer_bufs is a HOST array. */
int cir_search_topk_exchange(const void* q, int Q, const void* db, int64_t N, int Kd, int k, int32_t idx_offset,
                             void* const* peer_bufs, int n_peers, int my_rank,
                             void* workspace, size_t workspace_bytes, unsigned flags, void* stream);

/* The same with the merge fused in as well (Q <= the number of SMs, n_peers * k <= 8192): every exchange buffer is
 * [n_peers][2][Q][k] int32 words of lists followed by Q 32-bit arrival counters (zeroed once, never reset).  The block that
 * finishes query q stores its list into every peer's buffer, adds 1 to counter q of every peer (release, system scope), waits
 * until its OWN counter q reaches `arrivals` (= n_peers x the number of searches that have used this buffer, this one
 * included) and merges the n_peers lists itself: out_scores / out_idx [Q, k] receive the GLOBAL top-k on every rank.  No
 * cross-GPU barrier and no merge launch; alternate two buffers between consecutive searches.  Every rank must make the
 * matching call (a peer that never arrives traps the kernel after 60 s instead of hanging). */
int cir_search_topk_exchange_merge(const void* q, int Q, const void* db, int64_t N, int Kd, int k, int32_t idx_offset,
                                   void* const* peer_bufs, int n_peers, int my_rank, uint32_t arrivals,
                                   float* out_scores, int32_t* out_idx,
                                   void* workspace, size_t workspace_bytes, unsigned flags, void* stream);

/* dense scores through the same GEMM (the reference's full `scores` matrix, for full
 * ranking of small databases): out [Q, ld_out] fp32, out[q, n] = q . db[n] */
int cir_scores_dense(const void* q, int Q, const void* db, int64_t N, int Kd,
                     float* out, int64_t ld_out, void* stream);

/* full descending argsort of every row (np.argsort(-scores) per query): idx [Q, N] int32.
 * workspace: cir_sort_rows_workspace_bytes.  Ties: lower index first. */
int cir_sort_rows_workspace_bytes(int Q, int64_t N, size_t* bytes);
int cir_sort_rows_desc(const float* scores, int Q, int64_t N, int64_t ld,
                       int32_t* out_idx, float* out_sorted /* may be NULL */,
                       void* workspace, size_t workspace_bytes, void* stream);

/* merge G sorted top-k lists per query (shards of one GPU or the all-gathered lists of G
 * GPUs) into one: in [G, Q, k] -> out [Q, k_out], k_out <= k*G.  Exact:
 * top-k(union of shard top-k) == top-k(global).  Entries with idx < 0 are ignored;
 * k_out + k <= 4096, k_out <= 1024.  g_stride = elements between list g and list g + 1 (0 = Q * k, dense);
 * an all-gathered buffer that interleaves scores and indices per rank is merged in place with g_stride = 2 Q k. */
int cir_topk_merge(const float* scores, const int32_t* idx, int G, int Q, int k, int64_t g_stride,
                   float* out_scores, int32_t* out_idx, int k_out, void* stream);

/* exact fp32 re-scoring of candidate lists and final ordering:
 * q32 [Q, D], db32 [N, D] fp32; cand [Q, Kc] int32 (entries < 0 ignored, idx_offset is
 * subtracted to address db32) -> out [Q, k_out] sorted by (fp32 score desc, index asc) */
int cir_rescore_topk(const float* q32, int Q, const float* db32, int64_t N, int D,
                     const int32_t* cand, int Kc, int32_t idx_offset,
                     float* out_scores, int32_t* out_idx, int k_out, void* stream);

/* ------------------------------------------------------------------------------------
 * 4. alpha-QE / DBA aggregation (not in the reference)
 *    out[q] = L2N( q32[q] + sum_i max(s[q,i],0)^alpha * db32[idx[q,i]] )
 *    over the first k_use usable entries of the query's neighbour list idx[q, 0:klist):
 *    entries with idx < 0, and the self match idx == self_base + q (when self_base >= 0,
 *    database-side augmentation), are skipped and do not count.  out [Q, D] fp32.
 *    Row-sharded databases: q32 == NULL starts from zero and eps_l2 < 0 skips the L2N, which gives the sum over the
 *    neighbours of ONE shard (entries of other shards masked to -1); the parts are added (all_reduce) and normalised after.
 * ------------------------------------------------------------------------------------ */
int cir_qe_aggregate(const float* q32, int Q, const float* db32, int64_t N, int D,
                     const int32_t* idx, const float* scores, int klist, int ld_k, int k_use,
                     float alpha, int64_t self_base, float eps_l2,
                     float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * 5. Hard-negative selection
 *    replaces the greedy loop of TuplesDataset.create_epoch_tuples
 *                                         tuples_dataset.py:328-345
 *    For each query walk its ranked candidate list (pool positions, best first); take a
 *    candidate iff its cluster differs from the query's cluster and from every cluster
 *    already taken; stop at nnum.  out_sel [Q, nnum] pool positions (-1 if exhausted),
 *    out_count [Q] = number taken, out_dist [Q, nnum] = || q - n + 1e-6 ||_2 (:342).
 *    Optional exactness check (all four or none): cand_score [Q, Kc] = exact scores of the candidates, scan_tail [Q] = the
 *    approximate scan's score of the list's last entry (-inf if the list holds every admissible row); out_open[q] = 1 if
 *    the walk found fewer than nnum negatives or the exact score of the last one taken does not clear scan_tail by
 *    `margin` -- such a query must be re-run with a longer list.
 * ------------------------------------------------------------------------------------ */
int cir_mine_filter(const int32_t* cand, int Q, int Kc,
                    const int32_t* pool_cluster, int64_t P, const int32_t* q_cluster,
                    int nnum, const float* q32, const float* pool32, int D,
                    int32_t* out_sel, int32_t* out_count, float* out_dist,
                    const float* cand_score, const float* scan_tail, float margin, int32_t* out_open, void* stream);

/* ------------------------------------------------------------------------------------
 * 6. Evaluation of ranked lists (the consumer of the ranking step)
 *    replaces compute_ap / compute_map    cirtorch/utils/evaluation/ParisOxfordEval.py:4-113
 *    ranks [Q, ld] int32 row-major (row q = database indices best first, R entries used;
 *    negative entries = padding), positives / junk per query as CSR lists of SORTED indices
 *    (ok_off [Q+1], ok_idx; junk_off [Q+1], junk_idx).  Junk entries ranked before a positive
 *    move it up (:86-96); ap_out [Q] fp64 (NaN for a query without positives, :69-73);
 *    prs_out [Q, nk] = precision at kappas[j] with kq = min(max rank of a positive, kappa).
 * ------------------------------------------------------------------------------------ */
int cir_eval_ap(const int32_t* ranks, int Q, int64_t R, int64_t ld,
                const int32_t* ok_off, const int32_t* ok_idx,
                const int32_t* junk_off, const int32_t* junk_idx,
                const int32_t* kappas, int nk, double* ap_out, double* prs_out, void* stream);

/* ------------------------------------------------------------------------------------
 * 7. Training side of the head
 *    cir_tuple_loss replaces contrastive_loss / triplet_loss   cirtorch/modules/losses.py:7-23 / :26-46
 *    (called through globalFeatureLoss, cirtorch/algos/GF_algo.py:11-34,64-83) on the descriptors of
 *    n_tuples training tuples (query, positive, negatives): x [n_tuples * S, ldx] fp32 rows (the physical
 *    layout behind the reference's D x N tensor), label [n_tuples * S] int32 with -1 = query, 1 = positive,
 *    0 = negative.  ONE launch computes loss[0] (the sum over all pairs / triplets) and, when grad != NULL,
 *    d loss / d x into grad [n_tuples * S, ldg].  contrastive: the query is the FIRST row of a tuple
 *    (x[:, ::S], losses.py:13), eps is added to the difference (:20); triplet: squared distances, one
 *    positive per tuple (:38-44).  Deterministic (fixed-order reductions).
 * ------------------------------------------------------------------------------------ */
#define CIR_LOSS_CONTRASTIVE 0
#define CIR_LOSS_TRIPLET 1
int cir_tuple_loss_workspace_bytes(int n_tuples, size_t* bytes);
int cir_tuple_loss(const float* x, int64_t ldx, int n_tuples, int S, int D, const int32_t* label, int kind,
                   float margin, float eps, float* loss, float* grad, int64_t ldg,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the [N, C]-sized part of globalHead.forward (global_head.py:57-64), the pieces between the cuBLAS-sized
 * products:  cir_l2n_bwd_rows: out = d/dv [ v / (||v|| + eps) ] applied to gout, per row; unit_out (optional) = the
 * forward value v / (||v|| + eps); gout / out may be NULL when only unit_out is wanted.  cir_colsum_rows: out[c] =
 * sum_n X[n, c] (dL/db).  cir_gem_dp: dL/dp of GeM (pools.py:37-38) from the pooled values g, their gradient dg and
 * S = sum t^p ln t of cir_gem_bwd: out[0] (p_stride 0) or out[c] (p_stride 1). */
int cir_l2n_bwd_rows(const float* v, const float* gout, int64_t N, int C, float eps, float* out, float* unit_out,
                     void* stream);
int cir_colsum_rows(const float* X, int64_t N, int C, float* out, void* stream);
int cir_gem_dp(const float* g, const float* dg, const float* S, const float* p, int p_stride, int N, int C, int HW,
               float* out, void* workspace /* >= N * 4 + 32 bytes, needed for p_stride 0 */, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CIR_B200_H */
