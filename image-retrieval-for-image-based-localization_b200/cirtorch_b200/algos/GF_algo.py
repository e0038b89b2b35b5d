"""The algorithm glue between a backbone's feature maps and the fused head, with the reference's class names,
constructor arguments and return values (cirtorch/algos/GF_algo.py:10-94), so that the reference's
``ImageRetrievalNet`` (cirtorch/models/GF_net.py:10-126) takes this package's algo + head unchanged.
"""
from __future__ import annotations

from ..modules.losses import contrastive_loss, triplet_loss
from ..utils.sequence import PackedSequence


class Empty(Exception):
    """cirtorch/utils/misc.py:18-20: raised by a head that has nothing to predict on."""


class globalFeatureLoss:
    """GF_algo.py:10-34: ``name`` in {"triplet", "contrastive"}, ``sigma`` = margin, ``epsilon``."""

    def __init__(self, name=None, sigma=0.1, epsilon=1e-6):
        if name not in ("triplet", "contrastive"):
            raise ValueError("unknown loss %r" % (name,))
        self.name = name
        self.sigma = sigma
        self.epsilon = epsilon

    def _triplet_loss(self, x, label, label_msk):
        return triplet_loss(x, label=label, label_msk=label_msk, margin=self.sigma)

    def _contrastive_loss(self, x, label, label_msk):
        # the reference forwards label_msk to a function that does not take it (GF_algo.py:28, a TypeError there)
        return contrastive_loss(x, label=label, margin=self.sigma, eps=self.epsilon)

    def __call__(self, x, label, label_msk):
        return getattr(self, "_" + self.name + "_loss")(x, label, label_msk)


class globalFeatureAlgo:
    """GF_algo.py:37-94."""

    def __init__(self, loss, min_level, fpn_levels):
        self.loss = loss
        self.min_level = min_level
        self.fpn_levels = fpn_levels

    def _get_level(self, x):
        if isinstance(x, list):                       # FPN outputs: the first of the configured levels (:52-53)
            return x[self.min_level:self.min_level + self.fpn_levels][0]
        if isinstance(x, dict):                       # plain backbone: {"mod1" .. "mod5"} (:54-55)
            return x["mod5"]
        raise NameError("unknown input type")

    def _head(self, head, x):
        return head(x)

    def training(self, head, x, labels, img_size):
        """-> (ret_loss, ret_pred).  ``labels``: PackedSequence of per-tuple label tensors (:64-83)."""
        x = self._get_level(x)
        try:
            labels, labels_idx = labels.contiguous
            ret_pred = self._head(head, x)
            ret_loss = self.loss(ret_pred, labels, labels_idx)
        except Empty:
            ret_loss = sum(x_i.sum() for x_i in x) * 0
            ret_pred = None
        return ret_loss, ret_pred

    def inference(self, head, x, img_size):
        """-> D x B descriptors (:85-94).  ``img_size`` (the valid sizes of a padded batch) is accepted and, like in the
        reference, not used: GeM pools over the zero padding of a ragged batch."""
        x = self._get_level(x)
        try:
            return self._head(head, x)
        except Empty:
            return PackedSequence([None for _ in range(x[0].size(0))])
