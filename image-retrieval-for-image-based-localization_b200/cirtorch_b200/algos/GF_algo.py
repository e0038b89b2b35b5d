"""The algorithm glue between a backbone's feature maps and the fused head, with the reference's class names,
constructor arguments and return values (cirtorch/algos/GF_algo.py:10-94), so that the reference's
``ImageRetrievalNet`` (cirtorch/models/GF_net.py:10-126) takes this package's algo + head unchanged.
"""
from __future__ import annotations

from ..modules import losses as _losses
from ..utils.sequence import PackedSequence


class Empty(Exception):
    """cirtorch/utils/misc.py:18-20: raised by a head that has nothing to predict on."""


class globalFeatureLoss:
    """GF_algo.py:10-34: ``name`` in {"triplet", "contrastive"}, ``sigma`` = margin, ``epsilon``.
    Called as ``loss(x, label, label_msk)`` with x = D x N tuple descriptors."""

    _KINDS = ("triplet", "contrastive")

    def __init__(self, name=None, sigma=0.1, epsilon=1e-6):
        if name not in self._KINDS:
            raise ValueError("unknown loss %r (expected one of %s)" % (name, ", ".join(self._KINDS)))
        self.name, self.sigma, self.epsilon = name, sigma, epsilon

    def __call__(self, x, label, label_msk):
        if self.name == "triplet":
            return _losses.triplet_loss(x, label=label, label_msk=label_msk, margin=self.sigma)
        # the reference forwards label_msk to contrastive_loss, which does not take it (GF_algo.py:28: a TypeError there)
        return _losses.contrastive_loss(x, label=label, margin=self.sigma, eps=self.epsilon)

    # the reference's per-loss method names
    def _triplet_loss(self, x, label, label_msk):
        return globalFeatureLoss("triplet", self.sigma, self.epsilon)(x, label, label_msk)

    def _contrastive_loss(self, x, label, label_msk):
        return globalFeatureLoss("contrastive", self.sigma, self.epsilon)(x, label, label_msk)


class globalFeatureAlgo:
    """GF_algo.py:37-94: picks the feature level, runs the head, and -- in training -- the tuple loss."""

    def __init__(self, loss, min_level, fpn_levels):
        self.loss, self.min_level, self.fpn_levels = loss, min_level, fpn_levels

    def _get_level(self, x):
        """A plain backbone hands over {"mod1" .. "mod5"} (:54-55), an FPN a list of levels of which the first configured
        one is used (:52-53); anything else is an error (:56-57)."""
        if isinstance(x, dict):
            return x["mod5"]
        if isinstance(x, list):
            window = x[self.min_level:self.min_level + self.fpn_levels]
            return window[0]
        raise NameError("unknown input type")

    def _head(self, head, x):
        return head(x)

    def inference(self, head, x, img_size):
        """-> D x B descriptors (:85-94).  ``img_size`` (the valid sizes of a padded batch) is accepted and, like in the
        reference, not used: GeM pools over the zero padding of a ragged batch."""
        fmap = self._get_level(x)
        try:
            return self._head(head, fmap)
        except Empty:
            return PackedSequence([None] * fmap[0].size(0))

    def training(self, head, x, labels, img_size):
        """-> (ret_loss, ret_pred).  ``labels``: PackedSequence of per-tuple label tensors (:64-83)."""
        fmap = self._get_level(x)
        ret_pred = None
        try:
            flat_labels, tuple_of_label = labels.contiguous
            ret_pred = self._head(head, fmap)
            ret_loss = self.loss(ret_pred, flat_labels, tuple_of_label)
        except Empty:
            ret_loss = sum(level.sum() for level in fmap) * 0
        return ret_loss, ret_pred
