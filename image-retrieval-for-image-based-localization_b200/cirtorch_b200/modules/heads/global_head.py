"""globalHead (cirtorch/modules/heads/global_head.py:16-67) as ONE fused kernel launch.

Same constructor, same state-dict keys (``pool.p``, ``whiten.weight``, ``whiten.bias``), same
init (xavier_normal gain 0.1, zero bias, :41-50), same return value: a D x N permute view of a
contiguous N x D buffer (:67).
"""
import torch.nn as nn

from .. import pools as _pools
from ..pools import POOLING_LAYERS
from ..normalizations import NORMALIZATION_LAYERS
from ... import functional as LF


class globalHead(nn.Module):

    def __init__(self, pooling=None, normal=None, dim=None, norm_act=None):
        super().__init__()
        self.dim = dim
        self.whiten = nn.Linear(dim, dim, bias=True)
        if pooling["name"] == "GeMmp":
            self.pool = POOLING_LAYERS[pooling["name"]](**pooling["params"], mp=self.dim)
        else:
            self.pool = POOLING_LAYERS[pooling["name"]](**pooling["params"])
        self.pool_name = pooling["name"]
        self.norm = NORMALIZATION_LAYERS[normal["name"]](eps=1e-6)
        self.reset_parameters()

    def reset_parameters(self):
        for _, mod in self.named_modules():
            if isinstance(mod, nn.Linear):
                nn.init.xavier_normal_(mod.weight, 0.1)
            if hasattr(mod, "bias") and mod.bias is not None:
                nn.init.constant_(mod.bias, 0.)

    def forward(self, x, do_whitening=True, *, out=None, accumulate=False):
        """``out`` / ``accumulate`` (extensions, inference only): write the descriptors into / add them to an existing
        D x N result of this head -- the running sum of the multi-scale mean (GF_net.py:74-92)."""
        from ..normalizations import L2N
        if not isinstance(self.norm, L2N):
            # any other registered normalisation (PowerLaw): the reference's composition pool -> norm -> whiten -> norm
            # (global_head.py:57-64) layer by layer; only L2N is fused into the tail kernel
            if out is not None or accumulate:
                raise ValueError("out= / accumulate= need the fused L2N tail")
            v = self.norm(self.pool(x)).squeeze(-1).squeeze(-1)
            if do_whitening:
                v = self.norm(self.whiten(v))
            return v.permute(1, 0)
        phys = None if out is None else out.permute(1, 0)          # the physical N x D buffer behind a D x N result
        kw = dict(weight=self.whiten.weight, bias=self.whiten.bias, do_whitening=do_whitening, l2_eps=self.norm.eps,
                  out=phys, accumulate=accumulate)
        if self.pool_name in ("GeM", "GeMmp", "MAC", "SPoC"):
            y = LF.descriptor_tail(x, p=getattr(self.pool, "p", None), eps=getattr(self.pool, "eps", 1e-6),
                                   pooling=self.pool_name, **kw)
        else:
            # any other registered pooling (RMAC, ...): its own kernel(s), then L2N -> whiten -> L2N fused on the pooled
            # N x C x 1 x 1 vectors (the mean of one value is the value)
            v = self.pool(x)
            y = LF.descriptor_tail(v.reshape(v.shape[0], v.shape[1], 1, 1), pooling="SPoC", **kw)
        return y.permute(1, 0)


# the reference ships an identical copy under another class name (heads/ir_head.py)
ImageRetrievalHead = globalHead
