"""mAP / mP@k evaluation with the reference's function names and return values
(cirtorch/utils/evaluation/ParisOxfordEval.py:4-195), computed on the device from ranks that
are already in HBM (full ``N x Q`` rankings or ``k x Q`` top-k lists).

``ranks`` follows the reference: ``db_size x #queries`` (column i = ranking of query i, best
first), numpy or torch.  With top-k lists shorter than the database, positives that are not
in the list simply do not contribute (AP is then a lower bound).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def compute_ap(ranks, nres):
    """ParisOxfordEval.py:4-38: trapezoidal AP from the zero-based ranks of the positives (host helper)."""
    ap = 0.0
    recall_step = 1.0 / nres
    for j, rank in enumerate(np.asarray(ranks)):
        precision_0 = 1.0 if rank == 0 else float(j) / rank
        precision_1 = float(j + 1) / (rank + 1)
        ap += (precision_0 + precision_1) * recall_step / 2.0
    return ap


def _csr(lists, dev):
    off = np.zeros(len(lists) + 1, dtype=np.int32)
    flat = []
    for i, l in enumerate(lists):
        a = np.unique(np.asarray(l, dtype=np.int64).reshape(-1)).astype(np.int32)      # sorted
        flat.append(a)
        off[i + 1] = off[i] + a.shape[0]
    idx = np.concatenate(flat) if flat else np.zeros(0, np.int32)
    if idx.shape[0] == 0:
        idx = np.zeros(1, np.int32)
    return torch.from_numpy(off).to(dev), torch.from_numpy(idx.astype(np.int32)).to(dev)


def compute_map(ranks, gnd, kappas=[]):
    """ParisOxfordEval.py:41-113.  Returns (map, aps, pr, prs) exactly like the reference."""
    lib = _lib.load()
    if not torch.is_tensor(ranks):
        if not torch.cuda.is_available():
            raise _lib.CirError("cirtorch_b200.evaluate needs a CUDA device (no CPU fallback)")
        ranks = torch.as_tensor(np.asarray(ranks)).cuda()
    _lib.require_cuda(ranks)
    dev = ranks.device
    nq = len(gnd)
    if ranks.shape[1] != nq:
        raise ValueError("ranks has %d columns, gnd has %d queries" % (ranks.shape[1], nq))
    rows = ranks.t().to(torch.int32).contiguous()                     # [Q, R]
    ok_off, ok_idx = _csr([g["ok"] for g in gnd], dev)
    jk_off, jk_idx = _csr([g.get("junk", []) if isinstance(g, dict) else [] for g in gnd], dev)
    nk = len(kappas)
    kap = torch.as_tensor(list(kappas) if nk else [0], dtype=torch.int32, device=dev)
    aps = torch.empty(nq, dtype=torch.float64, device=dev)
    prs = torch.empty((nq, max(nk, 1)), dtype=torch.float64, device=dev)
    rc = lib.cir_eval_ap(_lib.ptr(rows), nq, rows.shape[1], rows.shape[1], _lib.ptr(ok_off), _lib.ptr(ok_idx),
                         _lib.ptr(jk_off), _lib.ptr(jk_idx), _lib.ptr(kap), nk, _lib.ptr(aps), _lib.ptr(prs),
                         _lib.stream_of(rows))
    _lib.check(rc, "cir_eval_ap")
    aps_h = aps.cpu().numpy()
    prs_h = prs.cpu().numpy()[:, :nk]
    valid = ~np.isnan(aps_h)
    nvalid = int(valid.sum())
    map_ = float(aps_h[valid].sum() / nvalid) if nvalid else float("nan")
    pr = prs_h[valid].sum(axis=0) / nvalid if nvalid else np.full(nk, np.nan)
    return map_, aps_h, pr, prs_h


def compute_map_and_print(dataset, ranks, gnd, log_info, kappas=[1, 5, 10]):
    """ParisOxfordEval.py:116-195: old (oxford5k / paris6k) and revisited (roxford5k / rparis6k) protocols."""
    if dataset.startswith("oxford5k") or dataset.startswith("paris6k"):
        map_, _, _, _ = compute_map(ranks, gnd)
        log_info("{%s}: mAP = {%f}", dataset, np.around(map_ * 100, decimals=2))
        return {"mAP": 100 * map_}
    if dataset.startswith("roxford5k") or dataset.startswith("rparis6k"):
        def regroup(ok_keys, junk_keys):
            return [{"ok": np.concatenate([g[k] for k in ok_keys]), "junk": np.concatenate([g[k] for k in junk_keys])}
                    for g in gnd]
        mapE, _, mprE, _ = compute_map(ranks, regroup(["easy"], ["junk", "hard"]), kappas)
        mapM, _, mprM, _ = compute_map(ranks, regroup(["easy", "hard"], ["junk"]), kappas)
        mapH, _, mprH, _ = compute_map(ranks, regroup(["hard"], ["junk", "easy"]), kappas)
        log_info("{%s}: mAP E: {%f}, M: {%f}, H: {%f}", dataset, np.around(mapE * 100, decimals=2),
                 np.around(mapM * 100, decimals=2), np.around(mapH * 100, decimals=2))
        for j in range(min(3, len(kappas))):
            log_info("{%s}: mP@k{%f} E: {%f}, M: {%f}, H: {%f}", dataset, kappas[j],
                     np.around(mprE * 100, decimals=2)[j], np.around(mprM * 100, decimals=2)[j],
                     np.around(mprH * 100, decimals=2)[j])
        return {"mAP": 100 * (mapM + mapH) / 2.0}
    raise ValueError("unknown dataset protocol: %s" % dataset)
