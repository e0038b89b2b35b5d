"""extract_vectors and the retrieval network wrapper.

``extract_vectors(net, images, image_size, transform, bbxs=None, ms=[1], msp=1)`` is imported by
scripts/test.py:12 and called at :200,:236,:238 but never defined in the fork; its in-tree
behaviour is the pair of loops scripts/train_globalF.py:666-680 / 708-722 (preallocate
``zeros(D, N)``, run the model batch by batch, write descriptor columns).  The backbone is stock
torchvision (not part of the product); everything after it is the fused tail kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .modules.heads.global_head import globalHead


class ImageRetrievalNet(nn.Module):
    """backbone -> globalHead.  Mirrors cirtorch/models/GF_net.py:63-126 for inference:
    ``forward(img, scales=[1])`` returns D x B; with several scales the per-scale descriptors are
    averaged WITHOUT re-normalisation (the fork's rule, GF_net.py:74-92 / SURVEY.md quirk Q3)."""

    def __init__(self, body: nn.Module, head: globalHead):
        super().__init__()
        self.body = body
        self.ret_head = head

    @staticmethod
    def _rescale(img, s):
        if s == 1:
            return img
        return F.interpolate(img, scale_factor=s, mode="bilinear", align_corners=False)   # GF_net.py:20-40

    def forward(self, img, scales=(1,), do_whitening=True):
        fused_sum = not (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()))
        acc, descs = None, []
        for s in scales:
            fmap = self.body(self._rescale(img, s))
            if isinstance(fmap, dict):          # the reference body returns {"mod1".."mod5"} (GF_algo.py:54-55)
                fmap = fmap["mod5"]
            if not fused_sum:
                descs.append(self.ret_head(fmap, do_whitening=do_whitening))
            elif acc is None:
                acc = self.ret_head(fmap, do_whitening=do_whitening)
            else:                               # later scales are added in the tail kernel's last phase
                self.ret_head(fmap, do_whitening=do_whitening, out=acc, accumulate=True)
        if not fused_sum:
            return descs[0] if len(descs) == 1 else torch.stack(descs, 0).mean(0)
        if len(scales) > 1:
            acc.mul_(1.0 / len(scales))         # avg_pool1d over the scales, no re-normalisation (GF_net.py:84-85)
        return acc


def resnet50_gem(dim=2048, p=3.0, pretrained=False):
    """torchvision ResNet50 (children[:-2], random init) + GeM/L2N/whiten head: BASELINE.json config 1."""
    import torchvision
    body = nn.Sequential(*list(torchvision.models.resnet50(weights=None).children())[:-2])
    head = globalHead(pooling={"name": "GeM", "params": {"p": p, "eps": 1e-6}},
                      normal={"name": "L2N", "params": {}}, dim=dim)
    return ImageRetrievalNet(body, head)


def _load(item, image_size, transform, bbx):
    if torch.is_tensor(item):
        img = item
    else:                                     # a path: PIL is optional in this image
        from PIL import Image
        img = Image.open(item).convert("RGB")
        if bbx is not None:
            img = img.crop(bbx)
        if image_size:
            img.thumbnail((image_size, image_size), Image.BILINEAR)
    return transform(img) if transform is not None else img


@torch.no_grad()
def extract_vectors(net, images, image_size=1024, transform=None, bbxs=None, ms=(1,), msp=1, batch_size=1,
                    device=None, rank=0, world_size=1, print_freq=0, pad_ragged=False):
    """images: list of image tensors (3 x H x W) / paths, or one B x 3 x H x W tensor -> D x N fp32 (CPU, like upstream).

    ``msp`` is accepted for signature compatibility; the fork averages scales without the
    generalized-mean power.  With ``world_size > 1`` rank r extracts the
    contiguous slice [r*n/W, (r+1)*n/W) and the caller all-gathers (parallel.extract_vectors_dp).

    Images of different sizes in one batch: by default every image is run at its own size (one launch per image, the
    descriptor does not depend on its batch mates).  ``pad_ragged=True`` reproduces the fork instead: the batch is
    zero-padded to Hmax x Wmax (cirtorch/utils/sequence.py:4-67) and GeM pools over the padding, because
    ``globalFeatureAlgo.inference`` ignores the valid sizes (GF_algo.py:85-90).
    """
    device = torch.device(device) if device is not None else next(net.parameters()).device
    net.eval()
    n = len(images)
    lo, hi = (rank * n) // world_size, ((rank + 1) * n) // world_size
    D = net.ret_head.dim
    vecs = torch.zeros(D, hi - lo, device=device)
    # equal-sized images are stacked straight into one of two pinned staging buffers and copied asynchronously: the host
    # side of a batch (stack + H2D) then overlaps the backbone of the previous one
    stage, staged = [None, None], [None, None]
    turn = 0
    i = lo
    while i < hi:
        j = min(hi, i + batch_size)
        items = [_load(images[t], image_size, transform, None if bbxs is None else bbxs[t]) for t in range(i, j)]
        shapes = {tuple(t.shape) for t in items}
        if len(shapes) == 1:
            shape = (batch_size,) + tuple(items[0].shape)
            if device.type == "cuda" and not items[0].is_cuda and items[0].dtype == torch.float32:
                if stage[turn] is None or tuple(stage[turn].shape) != shape:
                    stage[turn] = torch.empty(shape, dtype=torch.float32).pin_memory()
                elif staged[turn] is not None:
                    staged[turn].synchronize()          # the copy that last read this buffer has finished
                torch.stack(items, out=stage[turn][:len(items)])
                batch = stage[turn][:len(items)].to(device, non_blocking=True)
                staged[turn] = torch.cuda.Event()
                staged[turn].record()
                turn ^= 1
            else:
                batch = torch.stack(items).to(device, non_blocking=True)
            groups = [(list(range(i, j)), batch)]
        elif pad_ragged:                       # the fork's behaviour: one zero-padded batch
            from .utils.sequence import PackedSequence, pad_packed_images
            groups = [(list(range(i, j)), pad_packed_images(PackedSequence(items))[0].to(device, non_blocking=True))]
        else:                                  # ragged sizes: one launch per image
            groups = [([i + t], it[None].to(device, non_blocking=True)) for t, it in enumerate(items)]
        for ids, batch in groups:
            out = net(batch, scales=tuple(ms))
            vecs[:, ids[0] - lo:ids[-1] - lo + 1] = out
        i = j
    return vecs.cpu() if world_size == 1 else vecs
