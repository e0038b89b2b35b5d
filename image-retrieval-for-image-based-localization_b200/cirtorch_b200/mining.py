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


def mine_filter(cand, pool_cluster, q_cluster, nnum, q_rows=None, pool_rows=None, cand_scores=None, scan_tail=None, margin=0.0):
    """Greedy cluster-exclusion walk (cir_mine_filter) -> (sel [Q, nnum] pool positions, count [Q], dist [Q, nnum], open [Q]).

    ``cand_scores`` [Q, Kc] (exact scores of ``cand``) and ``scan_tail`` [Q] (the scan's score of the list's last entry)
    switch on the exactness check: ``open[q]`` = 1 where the walk is not conclusive (see mine_hard_negatives_rows)."""
    lib = _lib.load()
    cand = cand.to(torch.int32).contiguous()
    Q, Kc = cand.shape
    dev = cand.device
    sel = torch.empty((Q, nnum), dtype=torch.int32, device=dev)
    cnt = torch.empty((Q,), dtype=torch.int32, device=dev)
    dist = torch.zeros((Q, nnum), dtype=torch.float32, device=dev)
    open_ = None
    if cand_scores is not None:
        cand_scores = cand_scores.float().contiguous()
        scan_tail = scan_tail.float().contiguous()
        open_ = torch.empty((Q,), dtype=torch.int32, device=dev)
    D = 0 if q_rows is None else q_rows.shape[1]
    rc = lib.cir_mine_filter(_lib.ptr(cand), Q, Kc, _lib.ptr(pool_cluster), pool_cluster.shape[0], _lib.ptr(q_cluster),
                             nnum, _lib.ptr(q_rows), _lib.ptr(pool_rows), D, _lib.ptr(sel), _lib.ptr(cnt),
                             _lib.ptr(dist), _lib.ptr(cand_scores), _lib.ptr(scan_tail), float(margin), _lib.ptr(open_),
                             _lib.stream_of(cand))
    _lib.check(rc, "cir_mine_filter")
    return sel, cnt, dist, open_


# |scan score - fp32 score| for descriptors of norm <= 1 (what the head produces).  bf16x3: measured <= 1e-5 (DESIGN.md
# section 5).  bf16: the RIGOROUS bound 2^-8 (every factor carries a relative rounding error <= 2^-9, Cauchy-Schwarz), not the
# ~3e-4 seen in practice -- the walk's margin test must never pass wrongly.
TOL = {"bf16x3": 2e-5, "bf16": 2.0 ** -8}
KC_DEFAULT = {"bf16x3": 80, "bf16": 128}   # candidates per query of the first pass (bf16: 1/3 of the scan flops, a longer fp32 re-score)


def _full_ranking_fp32(sub_q, pool_rows):
    """Exact fp32 full ranking of a few straggler queries: bf16x3 dense scores order the pool, every 4096-entry window
    of that order is re-scored exactly in fp32 (cir_rescore_topk), the exact scores are scattered back into a dense
    [Q, P] matrix and sorted (torch.sort of the reference, tuples_dataset.py:319)."""
    dense = S.scores_dense_rows(sub_q, pool_rows, mode="bf16x3")
    order = S.argsort_rows_desc(dense)
    P = pool_rows.shape[0]
    for a in range(0, P, 4096):
        b = min(P, a + 4096)
        sc, ix = S.rescore_rows(sub_q, pool_rows, order[:, a:b], b - a)
        dense.scatter_(1, ix.long(), sc)
    return S.argsort_rows_desc(dense)


def mine_hard_negatives_rows(q_rows, pool_rows, q_cluster, pool_cluster, neg_num, mode="bf16x3", kc=None):
    """Device-level mining.  q_rows [Q, D], pool_rows [P, D] fp32; *_cluster int32 device tensors.

    Returns (sel [Q, neg_num] int32 pool positions best-first, count [Q], dist [Q, neg_num] = ||q - n + 1e-6||_2).

    ``mode``: "bf16x3" (default; scan error ~1e-5, 80 candidates) or "bf16" (a third of the scan's flops, 128 candidates and
    a wider margin from the rigorous rounding bound: 0.65 vs 0.81 ms at 2,000 x 20,000 x 2048, same sets).

    Exactness.  The candidates of a query are its kc best pool rows by the scan (own cluster masked), re-ordered by
    exact fp32 scores.  The walk over them is conclusive only if (a) it found neg_num negatives and (b) the fp32 score of
    the LAST negative taken clears the scan score of the kc-th candidate by 2 * TOL[mode] -- otherwise a row just outside
    the list could outrank the walk's tail -- or the list already holds every row the query may take.  Anything else is
    re-run with a 4x longer list and finally with an exact fp32 full ranking.
    """
    S._check_rows(q_rows)
    S._check_rows(pool_rows)
    Q = q_rows.shape[0]
    P = pool_rows.shape[0]
    q_cluster = q_cluster.to(device=q_rows.device, dtype=torch.int32).contiguous()
    pool_cluster = pool_cluster.to(device=q_rows.device, dtype=torch.int32).contiguous()
    kc = kc or min(S.MAX_K, max(KC_DEFAULT[mode], 16 * neg_num))
    qp = S.pack_rows(q_rows, "query", mode)
    pp = S.pack_rows(pool_rows, "db", mode)
    sel = cnt = dist = None
    todo = None               # indices of queries still open
    while True:
        sub_q = q_rows if todo is None else q_rows[todo].contiguous()
        sub_qp = qp if todo is None else qp[todo].contiguous()
        sub_qc = q_cluster if todo is None else q_cluster[todo].contiguous()
        kk = min(kc, S.MAX_K)
        s3, cand = S.search_packed(sub_qp, pp, kk, q_label=sub_qc, db_label=pool_cluster)
        sc, cand = S.rescore_rows(sub_q, pool_rows, cand, kk)         # exact fp32 order
        # the margin of the last negative taken over the first row that did NOT make the list is checked in the walk kernel
        s2, c2, d2, open_ = mine_filter(cand, pool_cluster, sub_qc, neg_num, sub_q, pool_rows, cand_scores=sc,
                                        scan_tail=s3[:, kk - 1], margin=2 * TOL[mode])
        if todo is None:
            sel, cnt, dist = s2, c2, d2
        else:
            sel[todo], cnt[todo], dist[todo] = s2, c2, d2
        if kk >= P or not bool(open_.any()):
            break
        todo = torch.nonzero(open_).flatten() if todo is None else todo[torch.nonzero(open_).flatten()]
        if kc >= S.MAX_K:
            # exact full ranking for the stragglers, then the same walk
            sub_q = q_rows[todo].contiguous()
            sub_qc = q_cluster[todo].contiguous()
            order = _full_ranking_fp32(sub_q, pool_rows)
            s2, c2, d2, _ = mine_filter(order, pool_cluster, sub_qc, neg_num, sub_q, pool_rows)
            sel[todo], cnt[todo], dist[todo] = s2, c2, d2
            break
        kc = min(S.MAX_K, kc * 4)
    return sel, cnt, dist


def mine_hard_negatives(qvecs, poolvecs, clusters, query_indices, idxs2images, neg_num, mode="bf16x3"):
    """Reference-shaped entry point (tuples_dataset.py:317-350).

    qvecs D x Q, poolvecs D x P (CUDA fp32, descriptors as columns); ``clusters[i]`` = cluster id of
    dataset image i; ``query_indices[q]`` = dataset index of query q; ``idxs2images[j]`` = dataset
    index of pool entry j.  Returns (negative_indices: list of Q lists of ``neg_num`` dataset
    indices, average negative L2 distance) exactly like the reference's ``self.negative_indices`` / return value.
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
    """Host-side state of the reference's ``TuplesDataset`` that mining touches (tuples_dataset.py:50-103), with its
    attribute names: ``clusters`` per image, ``query_pool`` / ``positive_pool``, ``neg_num``, ``query_size``,
    ``pool_size``; after :meth:`create_epoch_tuples`: ``query_indices``, ``positive_indices``, ``negative_indices``.

    ``images``: what the epoch's descriptors are extracted from -- a sequence of 3 x H x W tensors (or image paths,
    loaded with PIL) indexed like ``clusters``; or pass ``extract_fn`` to :meth:`create_epoch_tuples`."""

    def __init__(self, clusters, query_pool, positive_pool, neg_num=5, query_size=2000, pool_size=20000, images=None,
                 transform=None, name="retrieval-SfM-120k", mode="train", *, nnum=None, qsize=None, poolsize=None):
        self.name, self.mode = name, mode
        self.clusters = np.asarray(clusters)
        self.query_pool = np.asarray(query_pool)
        self.positive_pool = np.asarray(positive_pool)
        self.images = images
        self.transform = transform
        self.neg_num = neg_num if nnum is None else nnum
        self.query_size = min(query_size if qsize is None else qsize, len(self.query_pool))
        self.pool_size = min(pool_size if poolsize is None else poolsize, len(self.clusters))
        self.query_indices = self.positive_indices = self.negative_indices = None

    # upstream cnnimageretrieval-pytorch names of the same lists
    qidxs = property(lambda self: self.query_indices)
    pidxs = property(lambda self: self.positive_indices)
    nidxs = property(lambda self: self.negative_indices)

    def _load(self, i):
        item = self.images[i]
        if not torch.is_tensor(item):
            from PIL import Image
            with Image.open(item) as im:
                item = im.convert("RGB")
        return self.transform(item) if self.transform is not None else item

    def _extract(self, model, indices, output_dim, batch_size, device, rank, world_size):
        """The extraction loops of tuples_dataset.py:259-272 / :300-312: descriptors as columns of zeros(D, n); rank r of
        ``world_size`` fills the batches it owns (the reference's distributed sampler) and the parts are summed."""
        from .utils.sequence import PackedSequence
        vecs = torch.zeros(output_dim, len(indices), device=device)
        for b, a in enumerate(range(0, len(indices), batch_size)):
            if b % world_size != rank:
                continue
            batch = [self._load(i).to(device, non_blocking=True) for i in indices[a:a + batch_size]]
            out = model(img=PackedSequence(batch), do_prediction=True)
            pred = out[1]["ret_pred"] if isinstance(out, tuple) else out
            vecs[:, a:a + len(batch)] = pred
        if world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(vecs)
        return vecs

    def create_epoch_tuples(self, model, log_info=None, log_debug=None, *, extract_fn=None, generator=None, **varargs):
        """tuples_dataset.py:213-350 with its call shape ``(model, log_info, log_debug, output_dim=, world_size=, rank=,
        device=, data_config=)``: draws the epoch's (query, positive) pairs and the negative pool, extracts their
        descriptors with ``model`` (``model(img=PackedSequence, do_prediction=True)`` -> ``(_, {"ret_pred": D x B})``),
        mines ``neg_num`` hard negatives per query on the device, sets ``self.negative_indices``, logs and returns the
        average negative L2 distance.

        ``extract_fn(image_indices) -> D x n CUDA descriptors`` replaces the extraction (descriptors already in HBM)."""
        log_info = log_info or (lambda *a: None)
        log_debug = log_debug or (lambda *a: None)
        log_debug('Creating tuples for an epoch of {%s}--{%s}', self.name, self.mode)
        if hasattr(model, "eval"):
            model.eval()
        idxs2qpool = torch.randperm(len(self.query_pool), generator=generator)[:self.query_size]
        self.query_indices = [int(self.query_pool[i]) for i in idxs2qpool]
        self.positive_indices = [int(self.positive_pool[i]) for i in idxs2qpool]
        if self.neg_num == 0:
            self.negative_indices = [[] for _ in self.query_indices]
            return 0.0
        idxs2images = torch.randperm(len(self.clusters), generator=generator)[:self.pool_size]
        with torch.no_grad():
            if extract_fn is not None:
                qvecs = extract_fn(self.query_indices)
                poolvecs = extract_fn(idxs2images.tolist())
            else:
                if self.images is None:
                    raise ValueError("TuplesMiner needs images= (or extract_fn=) to extract the epoch's descriptors")
                cfg = varargs.get("data_config")
                batch_size = (cfg.getint("test_batch_size") * 2) if cfg is not None else varargs.get("batch_size", 16)
                device = varargs.get("device") or next(model.parameters()).device
                args = (varargs["output_dim"], batch_size, device, varargs.get("rank", 0), varargs.get("world_size", 1))
                log_debug('Extracting descriptors for query images :')
                qvecs = self._extract(model, self.query_indices, *args)
                log_debug('Extracting descriptors for negative pool :')
                poolvecs = self._extract(model, idxs2images.tolist(), *args)
            log_debug('Searching for hard negatives :')
            self.negative_indices, avg = mine_hard_negatives(qvecs, poolvecs, self.clusters, self.query_indices, idxs2images,
                                                             self.neg_num)
        log_info('Average negative l2-distance = %f', avg)
        return avg
