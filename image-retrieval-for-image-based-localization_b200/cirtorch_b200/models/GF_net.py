"""``ImageRetrievalNet(body, ret_algo, ret_head, augment=None)`` with the reference's forward contract
(cirtorch/models/GF_net.py:10-126): ``forward(img=PackedSequence, scales=[1], do_prediction=True)`` returns
``(OrderedDict(ret_loss=...), OrderedDict(ret_pred=D x B))``; ``do_loss=True`` packs (query, positive, negatives)
tuples and returns the tuple loss.  The body is any module returning ``{"mod5": map}`` or a list of FPN levels (stock
PyTorch, not part of the product); everything after it runs on the fused tail / loss kernels.
Multi-scale inference averages the per-scale descriptors WITHOUT re-normalisation (:74-92)."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from ..utils.sequence import PackedSequence, pad_packed_images


class ImageRetrievalNet(nn.Module):

    def __init__(self, body, ret_algo, ret_head, augment=None):
        super().__init__()
        self.augment = augment
        self.body = body
        self.ret_algo = ret_algo
        self.ret_head = ret_head

    @staticmethod
    def _prepare_pyramid_inputs(img, scales):
        def resized(t, s):
            return nn.functional.interpolate(t[None], scale_factor=s, mode="bilinear", align_corners=False)[0]
        return [img if s == 1 else PackedSequence([resized(t, s) for t in img]) for s in scales]

    @staticmethod
    def _prepare_inputs(img, positive_img, negative_img, labels=None):
        imgs, lbls = [], []
        for q, p, negs, lab in zip(img, positive_img, negative_img, labels):
            imgs += [q, p] + list(negs)
            lbls.append(lab)
        return PackedSequence(imgs), PackedSequence(lbls)

    def forward(self, img=None, positive_img=None, negative_img=None, scales=[1], do_augmentaton=False, do_loss=False,
                do_prediction=True, **varargs):
        labels = None
        if do_loss:
            if positive_img is None or negative_img is None:
                raise IOError(" Tuples is not correctly created ")
            img, labels = self._prepare_inputs(img, positive_img, negative_img, labels=varargs["tuple_labels"])
        if len(scales) > 1:                                    # evaluation only: image pyramid, mean over the scales
            preds = [self.forward(img=im, scales=[1], do_prediction=True, do_loss=False)[1]["ret_pred"]
                     for im in self._prepare_pyramid_inputs(img, scales)]
            return OrderedDict([("ret_loss", None)]), OrderedDict([("ret_pred", torch.stack(preds, 0).mean(0))])
        if self.augment:
            img = self.augment(img, masking=True, do_augmentation=do_augmentaton)
        img, valid_size = pad_packed_images(img)
        x = self.body(img)
        ret_loss = ret_pred = None
        if do_loss:
            ret_loss, ret_pred = self.ret_algo.training(self.ret_head, x, labels, valid_size)
        elif do_prediction:
            ret_pred = self.ret_algo.inference(self.ret_head, x, valid_size)
        return OrderedDict([("ret_loss", ret_loss)]), OrderedDict([("ret_pred", ret_pred)])
