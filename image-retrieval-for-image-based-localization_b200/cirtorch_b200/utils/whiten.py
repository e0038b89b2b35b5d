"""cirtorch/utils/whiten.py on the GPU: same names, arguments and return values.

``whitenapply`` (the per-descriptor hot part: N projections of D -> dims, then L2N) runs on the
tcgen05 GEMM in bf16x3 mode + the bias/L2N kernel.  ``whitenlearn`` / ``pcawhitenlearn`` are
one-off D x D factorizations; they run in fp64 through torch.linalg on the device (cuSOLVER --
library code, not a hot kernel) and keep the reference's semantics, including the growing
diagonal jitter of ``cholesky`` (whiten.py:50-65) and fp64 results for fp32 input.
Inputs may be numpy arrays (like the reference) or torch tensors; the result type follows X.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import _lib
from .. import search as _search


def _device():
    if not torch.cuda.is_available():
        raise _lib.CirError("cirtorch_b200.utils.whiten needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(a, dtype):
    if torch.is_tensor(a):
        return a.to(device=a.device if a.is_cuda else _device(), dtype=dtype)
    return torch.as_tensor(np.asarray(a), dtype=dtype, device=_device())


def _like(result: torch.Tensor, template):
    return result if torch.is_tensor(template) else result.cpu().numpy()


def whitenapply(X, m, P, dimensions=None):
    """whiten.py:4-12:  Y = P[:dimensions] (X - m);  Y /= (||Y||_2 over axis 0 + 1e-6).   X: D x N."""
    lib = _lib.load()
    if not dimensions:
        dimensions = P.shape[0]
    Xd = _to_dev(X, torch.float32)
    md = _to_dev(m, torch.float32).reshape(1, -1)
    Wd = _to_dev(P, torch.float32)[:dimensions].contiguous()
    rows = (Xd.t() - md).contiguous()                      # [N, D] centred descriptors
    N = rows.shape[0]
    out = torch.empty((N, dimensions), dtype=torch.float32, device=rows.device)
    if N:
        _search.scores_dense_rows(rows, Wd, mode="bf16x3", out=out)     # rows . W^T
        rc = lib.cir_bias_l2n_rows(_lib.ptr(out), N, dimensions, dimensions, None, 1e-6, _lib.ptr(out), dimensions,
                                   _lib.stream_of(out))
        _lib.check(rc, "cir_bias_l2n_rows")
    return _like(out.t(), X)


def cholesky(S):
    """whiten.py:50-65: Cholesky factor, adding 1e-10 * 10^j to the diagonal until S is positive definite."""
    Sd = _to_dev(S, torch.float64)
    eye = torch.eye(Sd.shape[0], dtype=torch.float64, device=Sd.device)
    alpha = 0.0
    while True:
        L, info = torch.linalg.cholesky_ex(Sd + alpha * eye)
        if int(info) == 0:
            return _like(L, S)
        alpha = 1e-10 if alpha == 0 else alpha * 10
        print(">>>> {}::cholesky: Matrix is not positive definite, adding {:.0e} on the diagonal"
              .format(os.path.basename(__file__), alpha))


def _eig_desc(M):
    # M is symmetric PSD by construction; eigh == the reference's eig up to eigenvector sign
    w, V = torch.linalg.eigh((M + M.t()) / 2)
    order = torch.argsort(w, descending=True)
    return w[order], V[:, order]


def _mean_like_numpy(X, Xd):
    """Row means in X's own precision (numpy keeps fp32 for fp32 input), returned as fp64."""
    src_dtype = X.dtype if torch.is_tensor(X) else torch.from_numpy(np.asarray(X)[:0]).dtype
    if src_dtype == torch.float32:
        return Xd.float().mean(dim=1, keepdim=True).double()
    return Xd.mean(dim=1, keepdim=True)


def pcawhitenlearn(X):
    """whiten.py:14-30 (unsupervised PCA whitening).  X: D x N -> (m D x 1, P D x D)."""
    Xd = _to_dev(X, torch.float64)
    N = Xd.shape[1]
    m = _mean_like_numpy(X, Xd)
    Xc = Xd - m
    cov = Xc @ Xc.t()
    cov = (cov + cov.t()) / (2 * N)
    w, V = _eig_desc(cov)
    P = torch.linalg.inv(torch.sqrt(torch.diag(w))) @ V.t()
    return _like(m, X), _like(P, X)


def whitenlearn(X, qidxs, pidxs):
    """whiten.py:32-48 (supervised Lw whitening from query / positive index pairs)."""
    Xd = _to_dev(X, torch.float64)
    q = torch.as_tensor(np.asarray(qidxs), dtype=torch.long, device=Xd.device)
    p = torch.as_tensor(np.asarray(pidxs), dtype=torch.long, device=Xd.device)
    m = _mean_like_numpy(X, Xd[:, q])
    df = Xd[:, q] - Xd[:, p]
    S = (df @ df.t()) / df.shape[1]            # fp64 here; the reference forms S in X's precision
    P = torch.linalg.inv(_to_dev(cholesky(S), torch.float64))
    df = P @ (Xd - m)
    _, V = _eig_desc(df @ df.t())
    P = V.t() @ P
    return _like(m, X), _like(P, X)
