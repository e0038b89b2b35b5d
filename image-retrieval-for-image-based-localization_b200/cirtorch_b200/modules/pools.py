"""Pooling layers with the reference's names and constructor signatures
(cirtorch/modules/pools.py:10-54, registry :200-207), running the fused CUDA tail kernel."""
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from .. import functional as LF
from .normalizations import L2N as _L2N


class MAC(nn.Module):
    """cirtorch/modules/pools.py:10-16."""

    def forward(self, x):
        return LF.mac(x)

    def __repr__(self):
        return self.__class__.__name__ + "()"


class SPoC(nn.Module):
    """cirtorch/modules/pools.py:20-26."""

    def forward(self, x):
        return LF.spoc(x)

    def __repr__(self):
        return self.__class__.__name__ + "()"


class GeM(nn.Module):
    """cirtorch/modules/pools.py:30-38: learnable exponent ``p`` (Parameter of shape [1]), eps clamp."""

    def __init__(self, p=3, eps=1e-6):
        super().__init__()
        self.p = Parameter(torch.ones(1) * p)
        self.eps = eps

    def forward(self, x):
        return LF.gem(x, p=self.p, eps=self.eps)

    def __repr__(self):
        return self.__class__.__name__ + "(p=%.4f, eps=%s)" % (float(self.p.detach().flatten()[0]), self.eps)


class GeMmp(nn.Module):
    """cirtorch/modules/pools.py:43-54: one exponent per channel (``mp`` = number of channels)."""

    def __init__(self, p=3, mp=1, eps=1e-6):
        super().__init__()
        self.mp = mp
        self.p = Parameter(torch.ones(self.mp) * p)
        self.eps = eps

    def forward(self, x):
        return LF.descriptor_tail(x, p=self.p, eps=self.eps, pooling="GeMmp", pool_only=True)[:, :, None, None]

    def __repr__(self):
        return self.__class__.__name__ + "(p=[%d], eps=%s)" % (self.mp, self.eps)


class RMAC(nn.Module):
    """cirtorch/modules/pools.py:57-113.  As written in the reference the region loops are mis-indented: the grid is
    computed for every level, but only the regions of the LAST level in its LAST column are pooled (one per row), so
    the result is L2N(MAC(x)) + sum_i L2N(MAC(x over region (i, last column) of level L)).  That behaviour is
    reproduced, with ONE pass over the map for all the max-pools; like the reference, L = 1 on a map with H >= W
    raises NameError (its ``cenW`` is never assigned)."""

    def __init__(self, L=3, eps=1e-6):
        super().__init__()
        self.L = L
        self.eps = eps

    def forward(self, x):
        N, C, H, W = x.shape
        if self.L < 1 or (self.L == 1 and not H < W):
            raise NameError("cenW is not defined for L=%d on a %dx%d map (reference behaviour, pools.py:94-98)" % (self.L, H, W))
        wl, cenH, cenW = LF.rmac_grid(H, W, self.L)[-1]
        if wl == 0:
            raise ValueError("RMAC: empty region at level %d of a %dx%d map" % (self.L, H, W))
        regions = [(0, 0, H, W)] + [(i, cenW[-1], wl, wl) for i in cenH]
        o = LF.region_pool(x, regions, pooling="MAC")                         # N x R x C
        R = len(regions)
        o = LF.l2n(o.reshape(N * R, C), eps=self.eps).reshape(N, R, C)
        return o.sum(dim=1)[:, :, None, None]

    def __repr__(self):
        return self.__class__.__name__ + "(L=%d)" % self.L


class Rpool(nn.Module):
    """cirtorch/modules/pools.py:116-197: ``rpool`` over the whole map and over every region of the L-level grid,
    L2N, optional whitening + L2N per region, sum over regions, L2N.

    With one of this package's pooling layers as ``rpool`` the map is read ONCE for all regions (cir_region_pool);
    the per-region L2N -> whiten (an ``nn.Linear``) -> L2N runs on the row-L2N kernel, the tcgen05 GEMM (bf16x3) and the
    bias + L2N kernel.
    Any other ``rpool`` / ``whiten`` callable takes the reference's region-by-region composition."""

    def __init__(self, rpool, whiten=None, L=3, eps=1e-6):
        super().__init__()
        self.rpool = rpool
        self.L = L
        self.whiten = whiten
        self.norm = _L2N(eps=1e-6)
        self.eps = eps

    def roipool(self, x):
        """-> N x R x C x 1 x 1 (pools.py:126-167)."""
        N, C, H, W = x.shape
        regions = [(0, 0, H, W)] + LF.rmac_regions(H, W, self.L)
        if isinstance(self.rpool, (GeM, GeMmp)):
            o = LF.region_pool(x, regions, p=self.rpool.p, eps=self.rpool.eps, pooling="GeM")
        elif isinstance(self.rpool, MAC):
            o = LF.region_pool(x, regions, pooling="MAC")
        elif isinstance(self.rpool, SPoC):
            o = LF.region_pool(x, regions, pooling="SPoC")
        else:
            vecs = [self.rpool(x.narrow(2, i, h).narrow(3, j, w).contiguous()).reshape(N, 1, C) for (i, j, h, w) in regions]
            o = torch.cat(vecs, dim=1)
        return o[:, :, :, None, None]

    def forward(self, x, aggregate=True):
        o = self.roipool(x)
        N, R, C = o.shape[:3]
        o = o.reshape(N * R, C)
        needs_grad = torch.is_grad_enabled() and (o.requires_grad or (isinstance(self.whiten, nn.Linear) and any(
            t.requires_grad for t in self.whiten.parameters())))
        if isinstance(self.whiten, nn.Linear) and not needs_grad:
            # L2N -> Linear -> L2N of the N*R region vectors: row L2N, the tcgen05 GEMM in bf16x3 mode (~fp32 accurate,
            # the path whitenapply takes) and the fused bias + L2N kernel
            from .. import _lib, search as _search
            o = LF.l2n(o, eps=self.norm.eps)
            Wt = self.whiten.weight.detach()
            z = torch.empty((N * R, Wt.shape[0]), dtype=torch.float32, device=o.device)
            _search.scores_dense_rows(o, Wt, mode="bf16x3", out=z)
            bias = self.whiten.bias.detach().float().contiguous() if self.whiten.bias is not None else None
            rc = _lib.load().cir_bias_l2n_rows(_lib.ptr(z), N * R, z.shape[1], z.shape[1], _lib.ptr(bias), float(self.norm.eps),
                                               _lib.ptr(z), z.shape[1], _lib.stream_of(z))
            _lib.check(rc, "cir_bias_l2n_rows")
            o = z
        else:
            o = LF.l2n(o, eps=self.norm.eps)
            if self.whiten is not None:
                o = LF.l2n(self.whiten(o), eps=self.norm.eps)
        D = o.shape[1]
        o = o.reshape(N, R, D, 1, 1)
        if aggregate:
            o = LF.l2n(o.sum(dim=1).reshape(N, D), eps=self.norm.eps).reshape(N, D, 1, 1)
        return o

    def __repr__(self):
        return super().__repr__() + "(L=%d)" % self.L


POOLING_LAYERS = {
    "MAC": MAC,
    "SPoC": SPoC,
    "GeM": GeM,
    "GeMmp": GeMmp,
    "RMAC": RMAC,
    "ROIpool": Rpool,
}
