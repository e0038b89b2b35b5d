"""CPU oracle for the cirtorch global-descriptor retrieval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: it may be
imported from ``tests/``, from ``__graft_entry__.smoke()`` and from ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, and only as the checker / the reported CPU
baseline.  The product (``cirtorch_b200``) never imports it and has no CPU fallback.

Every function restates, in numpy / torch-CPU, the arithmetic of one reference function
(paths relative to /root/reference, cited per function).  The reference ships no tests
and no golden vectors, so parity is pinned by running the
*imported* reference functions on seeded inputs in this container
(``tests/golden/make_golden.py``) and committing inputs+outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this oracle against those fixtures.

Parity status
  * tail (GeM / MAC / SPoC / L2N / globalHead), regional pooling (Rpool / RMAC), whitenapply / whitenlearn /
    pcawhitenlearn / cholesky, compute_ap / compute_map, multi-scale mean:
    PINNED by fixtures generated from the imported reference code.
  * contrastive_loss / triplet_loss (cirtorch/modules/losses.py:7-46) and the padded-batch forward of
    ImageRetrievalNet + globalFeatureAlgo (cirtorch/models/GF_net.py:94-126, cirtorch/algos/GF_algo.py:51-94):
    PINNED by fixtures generated from the imported reference code (losses.npz, net.npz).
  * ranking (scripts/train_globalF.py:733-734) and hard-negative mining
    (cirtorch/datasets/globalFeatures/tuples_dataset.py:317-350): the reference inlines
    them in a driver / a dataset method that cannot be called standalone; the fixture
    generator executes the same numpy/torch statements on seeded tensors.  PINNED to
    those statements, not to a callable.
  * alpha-QE / DBA: NOT PRESENT in the reference -> literature
    definition (Radenovic-Tolias-Chum TPAMI'18 aQE; Arandjelovic-Zisserman'12 / Gordo'17
    DBA).  PARITY UNPINNED by the reference; two independent restatements are
    cross-checked instead.
"""
from __future__ import annotations

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# Descriptor tail
# --------------------------------------------------------------------------------------


def gem(x: torch.Tensor, p, eps: float = 1e-6) -> torch.Tensor:
    """GeM pooling, cirtorch/modules/pools.py:30-38.

    y[n,c] = (mean_{h,w} max(x[n,c,h,w], eps) ** p) ** (1/p), result shaped N x C x 1 x 1.
    ``p`` may be a float, a 1-element tensor, or a per-channel tensor of shape [C]
    (GeMmp, pools.py:43-54).
    """
    p = torch.as_tensor(p, dtype=x.dtype)
    if p.numel() > 1:
        p = p.reshape(1, -1, 1, 1)
    powered = torch.clamp(x, min=eps) ** p
    pooled = powered.mean(dim=(-2, -1), keepdim=True)
    return pooled ** (1.0 / p)


def mac(x: torch.Tensor) -> torch.Tensor:
    """Global max pooling, cirtorch/modules/pools.py:10-16."""
    return x.amax(dim=(-2, -1), keepdim=True)


def spoc(x: torch.Tensor) -> torch.Tensor:
    """Global average pooling, cirtorch/modules/pools.py:20-26."""
    return x.mean(dim=(-2, -1), keepdim=True)


def l2n(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Channel L2 normalisation with eps ADDED to the norm, normalizations.py:9-16."""
    return x / (torch.linalg.vector_norm(x, ord=2, dim=1, keepdim=True) + eps)


def head_forward(x, p=3.0, eps=1e-6, weight=None, bias=None, do_whitening=True,
                 pooling="GeM", l2_eps=1e-6) -> torch.Tensor:
    """globalHead.forward, cirtorch/modules/heads/global_head.py:52-67.

    pool -> L2N -> squeeze -> Linear(weight [out,in], bias) -> L2N -> transpose; returns
    D x N (descriptors are columns).  do_whitening=False stops after the first L2N.
    """
    if pooling == "GeM" or pooling == "GeMmp":
        v = gem(x, p, eps)
    elif pooling == "MAC":
        v = mac(x)
    elif pooling == "SPoC":
        v = spoc(x)
    else:
        raise KeyError(pooling)
    v = l2n(v, l2_eps).flatten(1)
    if do_whitening:
        v = v @ weight.t()
        if bias is not None:
            v = v + bias
        v = l2n(v, l2_eps)
    return v.t()


def rmac_region_grid(H: int, W: int, L: int = 3):
    """Region grid shared by Rpool.roipool (cirtorch/modules/pools.py:126-165) and RMAC.forward (:64-103).

    Returns one (wl, tops, lefts) triple per level l = 1..L.  The number of extra regions along the long side is the
    candidate in {2..7} whose neighbour overlap is closest to 0.4 (:127-145); level l has side wl = floor(2 w / (l + 1))
    and its top-left corners are floor(wl2 + i * b) - wl2 with b = (extent - wl) / (count - 1) (:150-160).  float32
    torch arithmetic as in the reference (the 0.4 comparison has near-ties, e.g. 11 x 47)."""
    import math
    w = min(W, H)
    steps = torch.Tensor([2, 3, 4, 5, 6, 7])
    b = (max(H, W) - w) / (steps - 1)
    idx = torch.min(torch.abs(((w ** 2 - w * b) / w ** 2) - 0.4), 0)[1].item()
    Wd, Hd = (idx + 1, 0) if H < W else ((0, idx + 1) if H > W else (0, 0))
    out = []
    for l in range(1, L + 1):
        wl = math.floor(2 * w / (l + 1))
        wl2 = math.floor(wl / 2 - 1)
        bw = 0 if l + Wd == 1 else (W - wl) / (l + Wd - 1)
        bh = 0 if l + Hd == 1 else (H - wl) / (l + Hd - 1)
        lefts = (torch.floor(wl2 + torch.Tensor(range(l + Wd)) * bw).int() - wl2).tolist()
        tops = (torch.floor(wl2 + torch.Tensor(range(l + Hd)) * bh).int() - wl2).tolist()
        out.append((wl, tops, lefts))
    return out


def rpool_forward(x: torch.Tensor, pool, whiten_weight=None, whiten_bias=None, L: int = 3, aggregate: bool = True,
                  l2_eps: float = 1e-6) -> torch.Tensor:
    """Rpool.forward, cirtorch/modules/pools.py:169-197: ``pool`` (a callable N x C x h x w -> N x C x 1 x 1) over the
    whole map and every grid region (roipool :126-167), L2N per region, optional Linear + L2N per region, sum over
    regions, L2N.  Returns N x D x 1 x 1 (aggregate) or N x R x D x 1 x 1."""
    N, _, H, W = x.shape
    vecs = [pool(x)]
    for wl, tops, lefts in rmac_region_grid(H, W, L):
        if wl == 0:
            continue
        for i in tops:
            for j in lefts:
                vecs.append(pool(x[:, :, i:i + wl, j:j + wl]))
    o = torch.stack([v.reshape(N, -1) for v in vecs], dim=1)            # N x R x C
    R = o.shape[1]
    o = l2n(o.reshape(N * R, -1), l2_eps)
    if whiten_weight is not None:
        o = torch.nn.functional.linear(o, whiten_weight, whiten_bias)
        o = l2n(o, l2_eps)
    o = o.reshape(N, R, -1)
    if aggregate:
        return l2n(o.sum(dim=1), l2_eps)[:, :, None, None]
    return o[:, :, :, None, None]


def rmac_forward(x: torch.Tensor, L: int = 3, eps: float = 1e-6) -> torch.Tensor:
    """RMAC.forward AS WRITTEN, cirtorch/modules/pools.py:64-113.  The loops over the region centres (:105-113) sit
    outside the level loop and the pooling statements outside the column loop, so what is computed is
    L2N(MAC(x)) + sum over the rows i of level L of L2N(MAC(x[.., i:i+wl, j:j+wl])) with j = the LAST column centre
    of level L.  (L = 1 with H >= W leaves ``cenW`` undefined in the reference -> NameError.)"""
    N, C, H, W = x.shape
    if L == 1 and not H < W:
        raise NameError("cenW")
    v = l2n(mac(x), eps)
    wl, tops, lefts = rmac_region_grid(H, W, L)[-1]
    j = lefts[-1]
    for i in tops:
        v = v + l2n(mac(x[:, :, i:i + wl, j:j + wl]), eps)
    return v


def multiscale_mean(desc_per_scale) -> torch.Tensor:
    """Fork-style multi-scale aggregation, cirtorch/models/GF_net.py:74-92.

    Plain mean over scales of already-normalised D x N descriptors, no re-normalisation
    (SURVEY.md quirk Q3).
    """
    return torch.stack(list(desc_per_scale), dim=0).mean(dim=0)


# --------------------------------------------------------------------------------------
# Whitening helpers (numpy, like the reference)
# --------------------------------------------------------------------------------------


def whitenapply(X, m, P, dimensions=None):
    """cirtorch/utils/whiten.py:4-12: Y = P[:dims] (X - m); columns / (||.||_2 + 1e-6)."""
    dims = P.shape[0] if not dimensions else dimensions
    Y = P[:dims, :] @ (X - m)
    return Y / (np.sqrt((Y * Y).sum(axis=0, keepdims=True)) + 1e-6)


def cholesky(S):
    """cirtorch/utils/whiten.py:50-65: add 1e-10 * 10^j to the diagonal until PD."""
    jitter = 0.0
    eye = np.eye(S.shape[0], S.shape[1])
    while True:
        try:
            return np.linalg.cholesky(S + jitter * eye)
        except np.linalg.LinAlgError:
            jitter = 1e-10 if jitter == 0.0 else jitter * 10.0


def _eig_desc(M):
    w, V = np.linalg.eig(M)
    order = np.argsort(w)[::-1]
    return w[order], V[:, order]


def pcawhitenlearn(X):
    """cirtorch/utils/whiten.py:14-30: m = mean, P = diag(lambda)^-1/2 V^T (desc. eigval)."""
    n = X.shape[1]
    m = X.mean(axis=1, keepdims=True)
    Xc = X - m
    C = Xc @ Xc.T
    C = (C + C.T) / (2 * n)
    w, V = _eig_desc(C)
    P = np.linalg.inv(np.sqrt(np.diag(w))) @ V.T
    return m, P


def whitenlearn(X, qidxs, pidxs):
    """cirtorch/utils/whiten.py:32-48 (supervised Lw)."""
    m = X[:, qidxs].mean(axis=1, keepdims=True)
    d = X[:, qidxs] - X[:, pidxs]
    S = (d @ d.T) / d.shape[1]
    P = np.linalg.inv(cholesky(S))
    Z = P @ (X - m)
    _, V = _eig_desc(Z @ Z.T)
    return m, V.T @ P


# --------------------------------------------------------------------------------------
# Ranking / search
# --------------------------------------------------------------------------------------


def rank(database_vecs: np.ndarray, qvecs: np.ndarray):
    """scripts/train_globalF.py:733-734 (also scripts/test.py:246-248).

    database_vecs D x N, qvecs D x Q -> scores N x Q, ranks N x Q (best first along axis 0).
    """
    scores = database_vecs.T @ qvecs
    ranks = np.argsort(-scores, axis=0)
    return scores, ranks


def topk(database_vecs: np.ndarray, qvecs: np.ndarray, k: int):
    """Top-k prefix of :func:`rank` with a deterministic tie rule (score desc, index asc).

    Returns (scores k x Q, idx k x Q).  fp64 accumulate so it can referee fp32 results.
    """
    scores = database_vecs.T.astype(np.float64) @ qvecs.astype(np.float64)
    order = np.argsort(-scores, axis=0, kind="stable")[:k]
    return np.take_along_axis(scores, order, axis=0), order


# --------------------------------------------------------------------------------------
# Hard-negative mining
# --------------------------------------------------------------------------------------


def mine_hard_negatives(qvecs: torch.Tensor, poolvecs: torch.Tensor, clusters,
                        query_indices, idxs2images, neg_num: int):
    """cirtorch/datasets/globalFeatures/tuples_dataset.py:317-350.

    qvecs D x Q, poolvecs D x P (torch, fp32).  clusters[i] = cluster id of dataset image
    i; query_indices[q] = dataset index of query q; idxs2images[j] = dataset index of pool
    entry j.  Walk each query's ranking by descending similarity, take an image only if
    its cluster differs from the query's and from every negative already taken.
    Returns (negative_indices: list of lists of dataset indices, mean ||q - n + 1e-6||_2).
    """
    sims = poolvecs.t() @ qvecs
    _, order = torch.sort(sims, dim=0, descending=True)
    total = 0.0
    count = 0
    out = []
    idxs2images = [int(i) for i in idxs2images]
    for q in range(qvecs.shape[1]):
        seen = {clusters[query_indices[q]]}
        picked = []
        r = 0
        while len(picked) < neg_num:
            j = int(order[r, q])
            cand = idxs2images[j]
            if clusters[cand] not in seen:
                picked.append(cand)
                seen.add(clusters[cand])
                diff = qvecs[:, q] - poolvecs[:, j] + 1e-6
                total += float(torch.sqrt((diff * diff).sum()))
                count += 1
            r += 1
        out.append(picked)
    return out, total / max(count, 1)


# --------------------------------------------------------------------------------------
# alpha-QE / DBA (literature definition; NOT in the reference -> parity unpinned)
# --------------------------------------------------------------------------------------


def alpha_qe(qvecs: np.ndarray, database_vecs: np.ndarray, k: int = 10, alpha: float = 3.0):
    """alpha query expansion.

    q' = L2N(q + sum_{i<=k} max(s_i, 0)^alpha v_i) over the top-k neighbours v_i of q with
    s_i = q^T v_i.  D x Q in, D x Q out (fp64).
    """
    Q = qvecs.astype(np.float64)
    V = database_vecs.astype(np.float64)
    s, idx = topk(V, Q, k)
    out = Q.copy()
    for j in range(Q.shape[1]):
        w = np.maximum(s[:, j], 0.0) ** alpha
        out[:, j] += V[:, idx[:, j]] @ w
    return out / (np.sqrt((out * out).sum(axis=0, keepdims=True)) + 1e-6)


def alpha_qe_torch(qvecs: torch.Tensor, database_vecs: torch.Tensor, k=10, alpha=3.0):
    """Second, independent restatement of :func:`alpha_qe` (gather + bmm in torch fp64)."""
    Q = qvecs.double().t()
    V = database_vecs.double().t()
    sims = Q @ V.t()
    s, idx = torch.sort(sims, dim=1, descending=True, stable=True)
    s, idx = s[:, :k], idx[:, :k]
    w = s.clamp(min=0) ** alpha
    agg = Q + torch.einsum("qk,qkd->qd", w, V[idx])
    agg = agg / (agg.norm(dim=1, keepdim=True) + 1e-6)
    return agg.t()


def dba(database_vecs: np.ndarray, k: int = 10, alpha: float = 3.0):
    """Database-side augmentation: alpha-QE applied to every DB vector against the DB.

    The vector itself is its own rank-0 neighbour (s = 1 for unit-norm input) and is
    excluded from the neighbour sum: v' = L2N(v + sum_{i<=k, i != self} max(s_i,0)^alpha v_i).
    """
    V = database_vecs.astype(np.float64)
    s, idx = topk(V, V, k + 1)
    out = V.copy()
    for j in range(V.shape[1]):
        keep = idx[:, j] != j
        nb = idx[keep, j][:k]
        w = np.maximum(s[keep, j][:k], 0.0) ** alpha
        out[:, j] += V[:, nb] @ w
    return out / (np.sqrt((out * out).sum(axis=0, keepdims=True)) + 1e-6)


# --------------------------------------------------------------------------------------
# Evaluation (consumer of ranks)
# --------------------------------------------------------------------------------------


def compute_ap(pos_ranks, n_pos):
    """cirtorch/utils/evaluation/ParisOxfordEval.py:4-38 (trapezoidal AP)."""
    ap = 0.0
    step = 1.0 / n_pos
    for j, r in enumerate(pos_ranks):
        before = 1.0 if r == 0 else j / float(r)
        after = (j + 1) / float(r + 1)
        ap += 0.5 * (before + after) * step
    return ap


def compute_map(ranks, gnd, kappas=()):
    """cirtorch/utils/evaluation/ParisOxfordEval.py:41-113.

    ranks: N x Q (or k x Q) best-first; gnd[i]['ok'] / ['junk'] index lists.
    Returns (mAP, aps, mP@kappas, per-query P@kappas).
    """
    nq = len(gnd)
    aps = np.zeros(nq)
    prs = np.zeros((nq, len(kappas)))
    pr = np.zeros(len(kappas))
    total = 0.0
    empty = 0
    for i in range(nq):
        ok = np.asarray(gnd[i]["ok"])
        if ok.shape[0] == 0:
            aps[i] = np.nan
            prs[i, :] = np.nan
            empty += 1
            continue
        junk = np.asarray(gnd[i].get("junk", []))
        col = ranks[:, i]
        pos = np.flatnonzero(np.isin(col, ok))
        jnk = np.flatnonzero(np.isin(col, junk))
        if len(jnk):
            # each positive moves up by the number of junk entries ranked before it
            pos = pos - np.searchsorted(jnk, pos, side="left")
        ap = compute_ap(pos, len(ok))
        total += ap
        aps[i] = ap
        pos1 = pos + 1
        for j, kap in enumerate(kappas):
            kq = min(int(pos1.max()), kap)
            prs[i, j] = (pos1 <= kq).sum() / kq
        pr = pr + prs[i, :]
    denom = nq - empty
    return total / denom, aps, pr / denom, prs


def compute_map_revisited(ranks, gnd, kappas=(1, 5, 10)):
    """Easy/Medium/Hard protocol of compute_map_and_print, ParisOxfordEval.py:130-193."""
    def regroup(ok_keys, junk_keys):
        return [{"ok": np.concatenate([g[k] for k in ok_keys]),
                 "junk": np.concatenate([g[k] for k in junk_keys])} for g in gnd]
    e = compute_map(ranks, regroup(["easy"], ["junk", "hard"]), kappas)
    m = compute_map(ranks, regroup(["easy", "hard"], ["junk"]), kappas)
    h = compute_map(ranks, regroup(["hard"], ["junk", "easy"]), kappas)
    return {"mAP": 100.0 * (m[0] + h[0]) / 2.0, "E": e, "M": m, "H": h}


# --------------------------------------------------------------------------------------
# Tuple losses (training side of the head) and the padded-batch forward of the network
# --------------------------------------------------------------------------------------


def contrastive_loss(x: torch.Tensor, label: torch.Tensor, margin: float = 0.7, eps: float = 1e-6) -> torch.Tensor:
    """cirtorch/modules/losses.py:7-23.  x: D x N columns, a tuple = S consecutive columns starting with its query
    (label -1); every other column j of the tuple pairs with the query:
    d = ||x_q - x_j + eps||, y = 0.5 l d^2 + 0.5 (1 - l) max(margin - d, 0)^2, summed."""
    nq = int((label == -1).sum())
    S = x.shape[1] // nq
    total = x.new_zeros(())
    for t in range(nq):
        xq = x[:, t * S]
        for j in range(t * S, (t + 1) * S):
            if label[j] == -1:
                continue
            d = ((xq - x[:, j] + eps) ** 2).sum().sqrt()
            l = label[j].to(x.dtype)
            total = total + 0.5 * l * d ** 2 + 0.5 * (1 - l) * torch.clamp(margin - d, min=0) ** 2
    return total


def triplet_loss(x: torch.Tensor, label: torch.Tensor, label_msk: torch.Tensor, margin: float = 0.1) -> torch.Tensor:
    """cirtorch/modules/losses.py:26-46.  Per tuple: anchor (label -1), positive (1), negatives (0);
    sum over negatives of max(|a - p|^2 - |a - n|^2 + margin, 0)."""
    nt = len(torch.unique(label_msk))
    S = x.shape[1] // nt
    total = x.new_zeros(())
    for t in range(nt):
        cols = range(t * S, (t + 1) * S)
        a = x[:, [j for j in cols if label[j] == -1][0]]
        p = x[:, [j for j in cols if label[j] == 1][0]]
        dp = ((a - p) ** 2).sum()
        for j in cols:
            if label[j] == 0:
                total = total + torch.clamp(dp - ((a - x[:, j]) ** 2).sum() + margin, min=0.0)
    return total


def pad_images(images, pad_value: float = 0.0) -> torch.Tensor:
    """cirtorch/utils/sequence.py:4-58: ragged C x H x W images, top-left aligned in one N x C x Hmax x Wmax tensor."""
    H = max(t.shape[-2] for t in images)
    W = max(t.shape[-1] for t in images)
    out = images[0].new_full((len(images), images[0].shape[0], H, W), pad_value)
    for i, t in enumerate(images):
        out[i, :, :t.shape[-2], :t.shape[-1]] = t
    return out


def net_forward(body, images, p, eps, weight, bias, scales=(1,)) -> torch.Tensor:
    """ImageRetrievalNet.forward for inference (cirtorch/models/GF_net.py:63-126 with GF_algo.py:85-94): per scale the
    (rescaled) ragged images are zero-padded to one batch, run through the body, the "mod5" map goes through the head --
    pooling INCLUDES the padded area (the valid sizes are ignored, GF_algo.py:85-90); scales are averaged without
    re-normalisation (GF_net.py:74-92).  -> D x B."""
    preds = []
    for s in scales:
        ims = images if s == 1 else [torch.nn.functional.interpolate(t[None], scale_factor=s, mode="bilinear",
                                                                     align_corners=False)[0] for t in images]
        fmap = body(pad_images(ims))
        fmap = fmap["mod5"] if isinstance(fmap, dict) else fmap
        preds.append(head_forward(fmap, p, eps, weight, bias))
    return torch.stack(preds, 0).mean(0)
