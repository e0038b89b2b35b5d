"""Exhaustive inner-product search with fused top-k -- the ranking step of the reference

    scores = np.dot(database_vecs.T, qvecs); ranks = np.argsort(-scores, axis=0)
                                   (scripts/train_globalF.py:729-734, scripts/test.py:243-258)

on the tcgen05 search kernel (csrc/search.cu).  Public tensors follow cirtorch's convention
(descriptors are COLUMNS: D x N, results are k x Q / N x Q with the best match in row 0); the
``*_rows`` functions take the physical row-per-descriptor layout the kernels use.

Numerics ("mode"):
  "bf16"    operands rounded to bf16, fp32 accumulate.  |score error| <= ~5e-4 for unit-norm
            D=2048 descriptors; with ``rescore=True`` (default when the fp32
            rows are available) the top-(k+pad) candidates are re-scored exactly in fp32 and
            re-ordered, so the returned lists match an fp32 reference except at fp32 ties.
  "bf16x3"  each fp32 value split into hi + lo bf16; hi.hi + hi.lo + lo.hi accumulated in
            fp32 (K' = 3 D): ~1e-6 relative error, 3x the tensor work.  Used by mining.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

MAX_K = 512
_SPLIT = {"bf16": 1, "bf16x3": 3}


def _rows(x: torch.Tensor) -> torch.Tensor:
    """D x N column-descriptor tensor -> contiguous N x D rows (free when x is the head's permute view)."""
    if x.dim() != 2:
        raise ValueError("expected a D x N matrix, got shape %s" % (tuple(x.shape),))
    r = x.t()
    if r.dtype != torch.float32:
        r = r.float()
    return r if r.is_contiguous() else r.contiguous()


def _check_rows(x):
    _lib.require_cuda(x)
    if x.dim() != 2 or x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("expected a contiguous fp32 [rows, D] CUDA tensor")


def packed_width(D: int, mode: str) -> int:
    return _SPLIT[mode] * ((D + 63) // 64 * 64)


def pack_rows(x_rows: torch.Tensor, role: str, mode: str = "bf16", out: torch.Tensor = None) -> torch.Tensor:
    """fp32 [rows, D] -> K-major bf16 search operand [rows, Kd] (cir_pack_bf16)."""
    _check_rows(x_rows)
    lib = _lib.load()
    rows, D = x_rows.shape
    Kd = packed_width(D, mode)
    if out is None:
        out = torch.empty((rows, Kd), dtype=torch.bfloat16, device=x_rows.device)
    if rows == 0:
        return out
    rc = lib.cir_pack_bf16(_lib.ptr(x_rows), rows, D, D, _lib.ptr(out), Kd, _SPLIT[mode],
                           0 if role == "query" else 1, _lib.stream_of(x_rows))
    _lib.check(rc, "cir_pack_bf16")
    return out


NO_PREPASS, SAMPLE_FIRST_ROWS = 1, 2        # flags of cir_search_topk (include/cir_b200.h)


def search_packed(qp: torch.Tensor, dbp: torch.Tensor, k: int, idx_offset: int = 0, tau0=None,
                  q_label=None, db_label=None, out=None, flags: int = 0):
    """Top-k of packed operands: (scores [Q, k] fp32, idx [Q, k] int32), sorted (score desc, idx asc).

    ``flags``: NO_PREPASS / SAMPLE_FIRST_ROWS change how the running thresholds are warm-started, never the result."""
    _lib.require_cuda(qp, dbp, tau0, q_label, db_label)
    lib = _lib.load()
    Q, Kd = qp.shape
    N = dbp.shape[0]
    if dbp.shape[1] != Kd:
        raise ValueError("query / database operand widths differ: %d vs %d" % (Kd, dbp.shape[1]))
    if not 1 <= k <= MAX_K:
        raise ValueError("k=%d outside [1, %d]; use rank() for full rankings" % (k, MAX_K))
    if out is not None:        # caller-provided contiguous [Q, k] buffers (fp32 scores, int32 indices)
        scores, idx = out
    else:
        scores = torch.empty((Q, k), dtype=torch.float32, device=qp.device)
        idx = torch.empty((Q, k), dtype=torch.int32, device=qp.device)
    if Q == 0:
        return scores, idx
    if N == 0:
        scores.fill_(float("-inf"))
        idx.fill_(-1)
        return scores, idx
    need = C.c_size_t(0)
    _lib.check(lib.cir_search_workspace_bytes(Q, N, Kd, k, C.byref(need)), "cir_search_workspace_bytes")
    ws = _lib.workspace(qp.device, need.value, "search")
    if q_label is not None:
        q_label = q_label.to(torch.int32).contiguous()
        db_label = db_label.to(torch.int32).contiguous()
    rc = lib.cir_search_topk(_lib.ptr(qp), Q, _lib.ptr(dbp), N, Kd, k, _lib.ptr(tau0), _lib.ptr(q_label),
                             _lib.ptr(db_label), _lib.ptr(scores), _lib.ptr(idx), int(idx_offset),
                             _lib.ptr(ws), ws.numel(), int(flags), _lib.stream_of(qp))
    _lib.check(rc, "cir_search_topk")
    return scores, idx


def rescore_rows(q_rows, db_rows, cand, k_out, idx_offset: int = 0):
    """Exact fp32 scores of candidate lists + final ordering (cir_rescore_topk)."""
    _check_rows(q_rows)
    _check_rows(db_rows)
    lib = _lib.load()
    Q, D = q_rows.shape
    cand = cand.to(torch.int32).contiguous()
    Kc = cand.shape[1]
    scores = torch.empty((Q, k_out), dtype=torch.float32, device=q_rows.device)
    idx = torch.empty((Q, k_out), dtype=torch.int32, device=q_rows.device)
    rc = lib.cir_rescore_topk(_lib.ptr(q_rows), Q, _lib.ptr(db_rows), db_rows.shape[0], D, _lib.ptr(cand), Kc,
                              int(idx_offset), _lib.ptr(scores), _lib.ptr(idx), k_out, _lib.stream_of(q_rows))
    _lib.check(rc, "cir_rescore_topk")
    return scores, idx


def rescore_pad(k: int) -> int:
    """How many bf16 candidates are re-scored for a final top-k."""
    return min(MAX_K, k + max(28, k // 4))


def search_topk_rows(q_rows, db_rows, k, mode="bf16", rescore=None, db_packed=None, idx_offset=0,
                     q_label=None, db_label=None):
    """(scores [Q, k], idx [Q, k]) of fp32 row matrices; see the module docstring for ``mode``."""
    _check_rows(q_rows)
    if db_rows is not None:
        _check_rows(db_rows)
    if rescore is None:
        rescore = mode == "bf16" and db_rows is not None
    if rescore and db_rows is None:
        raise ValueError("rescore=True needs the fp32 database rows (db_rows); this index keeps only the packed operand")
    if db_rows is None and db_packed is None:
        raise ValueError("search_topk_rows needs db_rows or db_packed")
    qp = pack_rows(q_rows, "query", mode)
    dbp = db_packed if db_packed is not None else pack_rows(db_rows, "db", mode)
    if not rescore or q_rows.shape[0] == 0 or dbp.shape[0] == 0:
        return search_packed(qp, dbp, k, idx_offset, q_label=q_label, db_label=db_label)
    kc = min(rescore_pad(k), max(k, dbp.shape[0]))
    kc = min(kc, MAX_K)
    _, cand = search_packed(qp, dbp, kc, idx_offset, q_label=q_label, db_label=db_label)
    return rescore_rows(q_rows, db_rows, cand, k, idx_offset)


def search_topk(qvecs, database_vecs, k, mode="bf16", rescore=None):
    """cirtorch convention: qvecs D x Q, database_vecs D x N -> (scores k x Q, ranks k x Q int64)."""
    s, i = search_topk_rows(_rows(qvecs), _rows(database_vecs), k, mode=mode, rescore=rescore)
    return s.t(), i.t().long()


def scores_dense_rows(q_rows, db_rows, mode="bf16x3", out=None):
    """Full score matrix [Q, N] through the same GEMM (for small databases / full rankings)."""
    _check_rows(q_rows)
    _check_rows(db_rows)
    lib = _lib.load()
    Q, N = q_rows.shape[0], db_rows.shape[0]
    if out is None:
        out = torch.empty((Q, N), dtype=torch.float32, device=q_rows.device)
    if Q == 0 or N == 0:
        return out
    qp = pack_rows(q_rows, "query", mode)
    dbp = pack_rows(db_rows, "db", mode)
    rc = lib.cir_scores_dense(_lib.ptr(qp), Q, _lib.ptr(dbp), N, qp.shape[1], _lib.ptr(out), out.stride(0),
                              _lib.stream_of(q_rows))
    _lib.check(rc, "cir_scores_dense")
    return out


def argsort_rows_desc(scores: torch.Tensor, return_sorted=False):
    """Per-row full descending argsort (ties: lower index first) -> int32 [Q, N]."""
    _lib.require_cuda(scores)
    lib = _lib.load()
    if scores.dtype != torch.float32 or scores.stride(1) != 1:
        scores = scores.float().contiguous()
    Q, N = scores.shape
    idx = torch.empty((Q, N), dtype=torch.int32, device=scores.device)
    srt = torch.empty((Q, N), dtype=torch.float32, device=scores.device) if return_sorted else None
    if Q == 0 or N == 0:
        return (idx, srt) if return_sorted else idx
    need = C.c_size_t(0)
    _lib.check(lib.cir_sort_rows_workspace_bytes(Q, N, C.byref(need)), "cir_sort_rows_workspace_bytes")
    ws = _lib.workspace(scores.device, need.value, "sort")
    rc = lib.cir_sort_rows_desc(_lib.ptr(scores), Q, N, scores.stride(0), _lib.ptr(idx), _lib.ptr(srt),
                                _lib.ptr(ws), ws.numel(), _lib.stream_of(scores))
    _lib.check(rc, "cir_sort_rows_desc")
    return (idx, srt) if return_sorted else idx


def rank(database_vecs, qvecs, mode="bf16x3"):
    """The reference's full ranking (train_globalF.py:733-734): returns (scores N x Q, ranks N x Q int64)."""
    s = scores_dense_rows(_rows(qvecs), _rows(database_vecs), mode=mode)      # [Q, N]
    order = argsort_rows_desc(s)
    return s.t(), order.t().long()


def merge_topk(scores, idx, k_out=None):
    """Merge G sorted top-k lists per query: [G, Q, k] -> [Q, k_out] (cir_topk_merge).

    ``scores`` / ``idx`` may be strided views along G (e.g. the two halves of one all-gathered buffer)."""
    _lib.require_cuda(scores, idx)
    lib = _lib.load()
    G, Q, k = scores.shape
    k_out = k if k_out is None else k_out
    def dense_lists(t, dtype):
        if t.dtype != dtype:
            t = t.to(dtype)
        ok = t.stride(2) == 1 and t.stride(1) == k and (G == 1 or t.stride(0) >= Q * k)
        return t if ok else t.contiguous()
    scores = dense_lists(scores, torch.float32)
    idx = dense_lists(idx, torch.int32)
    if G > 1 and scores.stride(0) != idx.stride(0):
        scores, idx = scores.contiguous(), idx.contiguous()
    out_s = torch.empty((Q, k_out), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((Q, k_out), dtype=torch.int32, device=scores.device)
    if Q == 0:
        return out_s, out_i
    g_stride = scores.stride(0) if G > 1 else Q * k
    rc = lib.cir_topk_merge(_lib.ptr(scores), _lib.ptr(idx), G, Q, k, g_stride, _lib.ptr(out_s), _lib.ptr(out_i), k_out,
                            _lib.stream_of(scores))
    _lib.check(rc, "cir_topk_merge")
    return out_s, out_i


class Index:
    """A database resident in HBM: K-major bf16 rows for the scan (+ the fp32 rows for re-scoring).

    ``db_rows``: fp32 [N, D] CUDA tensor (or pass ``database_vecs`` D x N to :meth:`from_columns`).
    ``row_offset`` is added to every returned index (global ids of a shard).
    """

    def __init__(self, db_rows, mode="bf16", keep_fp32=True, row_offset=0, labels=None):
        _check_rows(db_rows)
        self.mode = mode
        self.N, self.D = db_rows.shape
        self.row_offset = int(row_offset)
        self.packed = pack_rows(db_rows, "db", mode)
        self.rows32 = db_rows if keep_fp32 else None
        self.labels = None if labels is None else labels.to(device=db_rows.device, dtype=torch.int32).contiguous()

    @classmethod
    def from_columns(cls, database_vecs, **kw):
        return cls(_rows(database_vecs), **kw)

    def search_rows(self, q_rows, k, rescore=None, q_label=None):
        if rescore is None:
            rescore = self.mode == "bf16" and self.rows32 is not None
        return search_topk_rows(q_rows, self.rows32 if rescore else None, k, mode=self.mode, rescore=rescore,
                                db_packed=self.packed, idx_offset=self.row_offset, q_label=q_label,
                                db_label=self.labels if q_label is not None else None)

    def search(self, qvecs, k, rescore=None):
        s, i = self.search_rows(_rows(qvecs), k, rescore=rescore)
        return s.t(), i.t().long()
