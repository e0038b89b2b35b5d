"""Training-time hard-negative mining: the arithmetic of TuplesDataset.create_epoch_tuples
(cirtorch/datasets/globalFeatures/tuples_dataset.py:213-350, arithmetic :317-345).

The reference does ``torch.mm`` (pool x queries), a full ``torch.sort`` of the 20000 x 2000 score
matrix and a Python loop with one GPU->CPU sync per probe.  Here: one search launch whose
epilogue already skips pool images of the query's own cluster (:330-332,:339), a top-K' list per
query, exact fp32 re-scoring, and a greedy "at most one negative per cluster" walk on the device
(:335-345).  Queries whose list is exhausted before ``neg_num`` negatives are found are re-run
with a longer list and finally with a full ranking, so the result equals the reference's.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from . import search as S


def mine_filter(cand, pool_cluster, q_cluster, nnum, q_rows=None, pool_rows=None):
    """Greedy cluster-exclusion walk (cir_mine_filter) -> (sel [Q, nnum] pool positions, count [Q], dist [Q, nnum])."""
    lib = _lib.load()
    cand = cand.to(torch.int32).contiguous()
    Q, Kc = cand.shape
    dev = cand.device
    sel = torch.empty((Q, nnum), dtype=torch.int32, device=dev)
    cnt = torch.empty((Q,), dtype=torch.int32, device=dev)
    dist = torch.zeros((Q, nnum), dtype=torch.float32, device=dev)
    D = 0 if q_rows is None else q_rows.shape[1]
    rc = lib.cir_mine_filter(_lib.ptr(cand), Q, Kc, _lib.ptr(pool_cluster), pool_cluster.shape[0], _lib.ptr(q_cluster),
                             nnum, _lib.ptr(q_rows), _lib.ptr(pool_rows), D, _lib.ptr(sel), _lib.ptr(cnt),
                             _lib.ptr(dist), _lib.stream_of(cand))
    _lib.check(rc, "cir_mine_filter")
    return sel, cnt, dist


def mine_hard_negatives_rows(q_rows, pool_rows, q_cluster, pool_cluster, neg_num, mode="bf16x3", kc=None):
    """Device-level mining.  q_rows [Q, D], pool_rows [P, D] fp32; *_cluster int32 device tensors.

    Returns (sel [Q, neg_num] int32 pool positions best-first, dist [Q, neg_num] = ||q - n + 1e-6||_2).
    """
    S._check_rows(q_rows)
    S._check_rows(pool_rows)
    Q = q_rows.shape[0]
    P = pool_rows.shape[0]
    q_cluster = q_cluster.to(device=q_rows.device, dtype=torch.int32).contiguous()
    pool_cluster = pool_cluster.to(device=q_rows.device, dtype=torch.int32).contiguous()
    kc = kc or min(S.MAX_K, max(64, 16 * neg_num))
    qp = S.pack_rows(q_rows, "query", mode)
    pp = S.pack_rows(pool_rows, "db", mode)
    sel = cnt = dist = None
    todo = None               # indices of queries still short of neg_num
    while True:
        sub_q = q_rows if todo is None else q_rows[todo].contiguous()
        sub_qp = qp if todo is None else qp[todo].contiguous()
        sub_qc = q_cluster if todo is None else q_cluster[todo].contiguous()
        kk = min(kc, S.MAX_K)
        _, cand = S.search_packed(sub_qp, pp, kk, q_label=sub_qc, db_label=pool_cluster)
        _, cand = S.rescore_rows(sub_q, pool_rows, cand, kk)          # exact fp32 order
        s2, c2, d2 = mine_filter(cand, pool_cluster, sub_qc, neg_num, sub_q, pool_rows)
        if todo is None:
            sel, cnt, dist = s2, c2, d2
        else:
            sel[todo], cnt[todo], dist[todo] = s2, c2, d2
        # a list is conclusive if it found neg_num negatives or already covers the whole pool
        short = (cnt < neg_num)
        if kk >= P or not bool(short.any()):
            break
        todo = torch.nonzero(short).flatten()
        if kc >= S.MAX_K:
            # full ranking for the stragglers: dense scores + full sort, then the same walk
            sub_q = q_rows[todo].contiguous()
            sub_qc = q_cluster[todo].contiguous()
            dense = S.scores_dense_rows(sub_q, pool_rows, mode="bf16x3")
            order = S.argsort_rows_desc(dense)
            s2, c2, d2 = mine_filter(order, pool_cluster, sub_qc, neg_num, sub_q, pool_rows)
            sel[todo], cnt[todo], dist[todo] = s2, c2, d2
            break
        kc = min(S.MAX_K, kc * 4)
    return sel, cnt, dist


def mine_hard_negatives(qvecs, poolvecs, clusters, query_indices, idxs2images, neg_num, mode="bf16x3"):
    """Reference-shaped entry point (tuples_dataset.py:317-350).

    qvecs D x Q, poolvecs D x P (CUDA fp32, descriptors as columns); ``clusters[i]`` = cluster id of
    dataset image i; ``query_indices[q]`` = dataset index of query q; ``idxs2images[j]`` = dataset
    index of pool entry j.  Returns (negative_indices: list of Q lists of ``neg_num`` dataset
    indices, average negative L2 distance) exactly like the reference's ``self.nidxs`` / return value.
    """
    dev = qvecs.device
    clusters_t = torch.as_tensor(np.asarray(clusters), dtype=torch.int64)
    idxs2images_t = torch.as_tensor(np.asarray(idxs2images), dtype=torch.int64)
    q_idx_t = torch.as_tensor(np.asarray(query_indices), dtype=torch.int64)
    q_cluster = clusters_t[q_idx_t].to(torch.int32).to(dev)
    pool_cluster = clusters_t[idxs2images_t].to(torch.int32).to(dev)
    sel, cnt, dist = mine_hard_negatives_rows(S._rows(qvecs), S._rows(poolvecs), q_cluster, pool_cluster, neg_num, mode)
    sel_c, cnt_c, dist_c = sel.cpu(), cnt.cpu(), dist.cpu()           # ONE device->host transfer
    if bool((cnt_c < neg_num).any()):
        q_bad = int(torch.nonzero(cnt_c < neg_num)[0])
        raise IndexError("query %d: only %d of %d negatives with distinct clusters exist in the pool"
                         % (q_bad, int(cnt_c[q_bad]), neg_num))
    negs = idxs2images_t[sel_c.long()]
    avg = float(dist_c.sum() / max(int(cnt_c.sum()), 1))
    return negs.tolist(), avg


class TuplesMiner:
    """Host-side state of the reference's TuplesDataset that mining touches (tuples_dataset.py:50-103):
    ``clusters`` per image, the query / positive pools, epoch sizes; ``create_epoch_tuples`` re-draws
    the epoch's queries and negative pool (:223-229) and mines ``nnum`` negatives per query."""

    def __init__(self, clusters, qpool, ppool, nnum=5, qsize=2000, poolsize=20000):
        self.clusters = np.asarray(clusters)
        self.qpool = np.asarray(qpool)
        self.ppool = np.asarray(ppool)
        self.nnum = nnum
        self.qsize = min(qsize, len(self.qpool))
        self.poolsize = min(poolsize, len(self.clusters))
        self.qidxs = self.pidxs = self.nidxs = None

    def create_epoch_tuples(self, extract_fn, generator=None):
        """``extract_fn(image_indices) -> D x n CUDA descriptors`` (e.g. a closure over extract_vectors)."""
        idxs2qpool = torch.randperm(len(self.qpool), generator=generator)[:self.qsize]
        self.qidxs = [int(self.qpool[i]) for i in idxs2qpool]
        self.pidxs = [int(self.ppool[i]) for i in idxs2qpool]
        if self.nnum == 0:
            self.nidxs = [[] for _ in self.qidxs]
            return 0.0
        idxs2images = torch.randperm(len(self.clusters), generator=generator)[:self.poolsize]
        qvecs = extract_fn(self.qidxs)
        poolvecs = extract_fn(idxs2images.tolist())
        self.nidxs, avg = mine_hard_negatives(qvecs, poolvecs, self.clusters, self.qidxs, idxs2images, self.nnum)
        return avg
