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

    def forward(self, x, do_whitening=True):
        y = LF.descriptor_tail(
            x, p=getattr(self.pool, "p", None), eps=getattr(self.pool, "eps", 1e-6),
            weight=self.whiten.weight, bias=self.whiten.bias, pooling=self.pool_name,
            do_whitening=do_whitening, l2_eps=self.norm.eps)
        return y.permute(1, 0)


# the reference ships an identical copy under another class name (heads/ir_head.py)
ImageRetrievalHead = globalHead
