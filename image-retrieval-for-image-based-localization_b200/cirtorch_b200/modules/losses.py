"""Tuple losses of the retrieval head with the reference's names and signatures
(cirtorch/modules/losses.py:7-23 ``contrastive_loss``, :26-46 ``triplet_loss``), each ONE kernel launch that returns
the loss and keeps d loss / d x for the backward pass (cir_tuple_loss, csrc/train.cu).

``x`` is the reference's D x N tensor of tuple descriptors (columns; N = tuples * S, a tuple = query, positive,
negatives in consecutive columns) -- i.e. exactly what ``globalHead.forward`` returns, whose physical layout is the
row-per-descriptor N x D buffer the kernel reads.  ``label``: -1 query, 1 positive, 0 negative.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib

CONTRASTIVE, TRIPLET = 0, 1


def _launch(rows, label, n_tuples, kind, margin, eps, want_grad):
    lib = _lib.load()
    n, D = rows.shape
    S = n // n_tuples
    loss = torch.empty((1,), dtype=torch.float32, device=rows.device)
    grad = torch.empty_like(rows) if want_grad else None
    need = C.c_size_t(0)
    _lib.check(lib.cir_tuple_loss_workspace_bytes(n_tuples, C.byref(need)), "cir_tuple_loss_workspace_bytes")
    ws = _lib.workspace(rows.device, need.value, "loss")
    rc = lib.cir_tuple_loss(_lib.ptr(rows), D, n_tuples, S, D, _lib.ptr(label), kind, float(margin), float(eps),
                            _lib.ptr(loss), _lib.ptr(grad), D, _lib.ptr(ws), ws.numel(), _lib.stream_of(rows))
    _lib.check(rc, "cir_tuple_loss")
    return loss, grad


class _TupleLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, label, n_tuples, kind, margin, eps):
        rows = x.t()
        if rows.dtype != torch.float32 or not rows.is_contiguous():
            rows = rows.float().contiguous()
        loss, grad = _launch(rows, label, n_tuples, kind, margin, eps, x.requires_grad)
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return (grad * gout).t(), None, None, None, None, None


def _prepare(x, label):
    _lib.require_cuda(x, label)
    if x.dim() != 2:
        raise ValueError("expected the D x N matrix of tuple descriptors, got shape %s" % (tuple(x.shape),))
    if label is None or label.numel() != x.shape[1]:
        raise ValueError("label must hold one entry per column of x")
    return label.to(torch.int32).contiguous()


def contrastive_loss(x, label=None, margin=0.7, eps=1e-6):
    """losses.py:7-23: sum over the (query, other) pairs of every tuple of
    0.5 l d^2 + 0.5 (1 - l) max(margin - d, 0)^2,  d = || x_q - x_j + eps ||_2."""
    lab = _prepare(x, label)
    nq = int((lab == -1).sum())                      # number of tuples (one host sync, as in the reference :10-11)
    if nq == 0 or x.shape[1] % nq:
        raise ValueError("labels hold %d queries for %d descriptors" % (nq, x.shape[1]))
    return _TupleLossFn.apply(x, lab, nq, CONTRASTIVE, margin, eps)


def triplet_loss(x, label=None, label_msk=None, margin=0.1):
    """losses.py:26-46: sum over tuples and negatives of max(|a - p|^2 - |a - n|^2 + margin, 0)."""
    lab = _prepare(x, label)
    n_tuples = int(torch.unique(label_msk).numel())  # :28
    if n_tuples == 0 or x.shape[1] % n_tuples:
        raise ValueError("label_msk names %d tuples for %d descriptors" % (n_tuples, x.shape[1]))
    return _TupleLossFn.apply(x, lab, n_tuples, TRIPLET, margin, 0.0)
