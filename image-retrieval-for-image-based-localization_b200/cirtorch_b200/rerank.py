"""alpha query expansion and database-side augmentation on top of the search kernel.

Neither exists in the reference; BASELINE.json config 5 asks for them.
Definition (Radenovic-Tolias-Chum TPAMI'18; Arandjelovic-Zisserman'12 / Gordo'17):
    q' = L2N( q + sum_{i<=k} max(s_i, 0)^alpha v_i ),   v_i = top-k neighbours of q, s_i = q^T v_i
DBA applies the same update to every database vector against the database, the vector itself
(its own rank-0 neighbour) excluded from the sum.  Parity is pinned to oracle/ only.
"""
from __future__ import annotations

import torch

from . import _lib
from .search import Index, _rows, _check_rows


def qe_aggregate_rows(q_rows, db_rows, idx, scores, k_use, alpha, self_base=-1, idx_offset=0, eps=1e-6, normalize=True):
    """cir_qe_aggregate.  ``q_rows=None`` and ``normalize=False`` give the raw neighbour sum of one database shard
    (entries that live in other shards masked to -1): parts are added across ranks and normalised afterwards."""
    _check_rows(db_rows)
    lib = _lib.load()
    idx = idx.to(torch.int32)
    if idx_offset:
        idx = torch.where(idx >= 0, idx - idx_offset, idx)
    idx = idx.contiguous()
    scores = scores.float().contiguous()
    Q, D = idx.shape[0], db_rows.shape[1]
    if q_rows is not None:
        _check_rows(q_rows)
        if tuple(q_rows.shape) != (Q, D):
            raise ValueError("queries are %s, lists / database say %d x %d" % (tuple(q_rows.shape), Q, D))
    out = torch.empty((Q, D), dtype=torch.float32, device=db_rows.device)
    rc = lib.cir_qe_aggregate(_lib.ptr(q_rows), Q, _lib.ptr(db_rows), db_rows.shape[0], D, _lib.ptr(idx),
                              _lib.ptr(scores), idx.shape[1], idx.shape[1], int(k_use), float(alpha), int(self_base),
                              float(eps) if normalize else -1.0, _lib.ptr(out), _lib.stream_of(db_rows))
    _lib.check(rc, "cir_qe_aggregate")
    return out


def alpha_qe_rows(q_rows, index: Index, k=10, alpha=3.0):
    """Expanded queries [Q, D] (search with them afterwards: ``index.search_rows(q2, k)``)."""
    if index.rows32 is None:
        raise ValueError("alpha-QE needs the fp32 database rows (Index(keep_fp32=True))")
    s, i = index.search_rows(q_rows, k)
    return qe_aggregate_rows(q_rows, index.rows32, i, s, k, alpha, idx_offset=index.row_offset)


def alpha_qe(qvecs, database_vecs, k=10, alpha=3.0, mode="bf16"):
    """D x Q queries, D x N database -> D x Q expanded, L2-normalised queries."""
    index = Index.from_columns(database_vecs, mode=mode)
    return alpha_qe_rows(_rows(qvecs), index, k, alpha).t()


def dba_rows(db_rows, k=10, alpha=3.0, mode="bf16", chunk=16384, index: Index = None, row_begin=0, row_end=None):
    """Augmented database rows [row_begin, row_end) of ``db_rows`` (all rows by default)."""
    _check_rows(db_rows)
    index = index if index is not None else Index(db_rows, mode=mode)
    row_end = db_rows.shape[0] if row_end is None else row_end
    out = torch.empty((row_end - row_begin, db_rows.shape[1]), dtype=torch.float32, device=db_rows.device)
    for a in range(row_begin, row_end, chunk):
        b = min(row_end, a + chunk)
        q = db_rows[a:b]
        s, i = index.search_rows(q, min(k + 1, index.N))
        out[a - row_begin:b - row_begin] = qe_aggregate_rows(q, db_rows, i, s, k, alpha, self_base=a)
    return out


def dba(database_vecs, k=10, alpha=3.0, mode="bf16"):
    """D x N database -> D x N augmented, L2-normalised database."""
    return dba_rows(_rows(database_vecs), k, alpha, mode).t()
