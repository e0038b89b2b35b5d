"""cirtorch/utils/whiten.py on the GPU: same names, arguments and return values.

``whitenapply`` (the per-descriptor hot part: N projections of D -> dims, then L2N) runs on the
tcgen05 GEMM in bf16x3 mode + the bias/L2N kernel.  ``whitenlearn`` / ``pcawhitenlearn`` are
one-off D x D factorizations; they run through torch.linalg on the device (cuSOLVER -- library
code, not a hot kernel) and keep the reference's semantics: the covariance in the descriptors' own
precision, everything after ``cholesky`` in fp64, the growing diagonal jitter of ``cholesky``
(whiten.py:50-65) and fp64 results for fp32 input.
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


def _result_dtype(*arrays):
    """numpy's promotion of the reference expression (fp32 X with the fp64 m, P of whitenlearn gives fp64)."""
    return np.result_type(*[np.asarray(a).dtype for a in arrays])


def whitenapply(X, m, P, dimensions=None):
    """whiten.py:4-12:  Y = P[:dimensions] (X - m);  Y /= (||Y||_2 over axis 0 + 1e-6).   X: D x N.

    One pass: Y = X W^T + b with W = P[:dimensions] and b = -W m folded into the bias of the L2N kernel (no centred
    copy of X is materialised).  The projection runs on the tcgen05 GEMM with bf16x3 operands (~1e-6 relative).
    numpy in -> numpy out with numpy's result dtype of the reference expression (fp64 for the fp64 m, P of whitenlearn);
    a torch tensor X stays on the device and the result is the fp32 D' x N view of the kernel's output."""
    lib = _lib.load()
    if not dimensions:
        dimensions = P.shape[0]
    Xd = _to_dev(X, torch.float32)
    rows = Xd.t()
    rows = rows if rows.is_contiguous() else rows.contiguous()      # [N, D]; free for the head's permute view
    W64 = _to_dev(P, torch.float64)[:dimensions]
    bias = -(W64 @ _to_dev(m, torch.float64).reshape(-1, 1)).reshape(-1).float().contiguous()
    Wd = W64.float().contiguous()
    N = rows.shape[0]
    out = torch.empty((N, dimensions), dtype=torch.float32, device=rows.device)
    if N:
        _search.scores_dense_rows(rows, Wd, mode="bf16x3", out=out)     # rows . W^T
        rc = lib.cir_bias_l2n_rows(_lib.ptr(out), N, dimensions, dimensions, _lib.ptr(bias), 1e-6, _lib.ptr(out), dimensions,
                                   _lib.stream_of(out))
        _lib.check(rc, "cir_bias_l2n_rows")
    if torch.is_tensor(X):
        return out.t()
    return out.t().cpu().numpy().astype(_result_dtype(X, m, P), copy=False)


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


def lw_covariance(X, qidxs, pidxs):
    """whiten.py:35-37: m = mean of the query descriptors, S = cov of the (query - positive) differences.  Formed in X's
    OWN precision like the reference (fp32 products and accumulation for fp32 descriptors; the summation order of a BLAS
    is not reproducible, the precision is)."""
    src = X.dtype if torch.is_tensor(X) else torch.from_numpy(np.asarray(X)[:0]).dtype
    Xw = _to_dev(X, torch.float32 if src == torch.float32 else torch.float64)
    q = torch.as_tensor(np.asarray(qidxs), dtype=torch.long, device=Xw.device)
    p = torch.as_tensor(np.asarray(pidxs), dtype=torch.long, device=Xw.device)
    m = Xw[:, q].mean(dim=1, keepdim=True)
    df = Xw[:, q] - Xw[:, p]
    return m, (df @ df.t()) / df.shape[1]


def lw_from_covariance(S, X, m):
    """whiten.py:38-46 in fp64 (the reference's cholesky promotes to fp64): P = inv(chol(S)), then the eigenvectors of
    the projected scatter, by descending eigenvalue, rotate P."""
    Xd = _to_dev(X, torch.float64)
    P = torch.linalg.inv(_to_dev(cholesky(_to_dev(S, torch.float64)), torch.float64))
    df = P @ (Xd - _to_dev(m, torch.float64))
    _, V = _eig_desc(df @ df.t())
    return V.t() @ P


def whitenlearn(X, qidxs, pidxs):
    """whiten.py:32-48 (supervised Lw whitening from query / positive index pairs) -> (m D x 1, P D x D)."""
    m, S = lw_covariance(X, qidxs, pidxs)
    P = lw_from_covariance(S, X, m)
    return _like(m, X), _like(P, X)
