"""Ragged image batches either side of the path: ``PackedSequence`` (cirtorch/utils/parallel/packed_sequence.py:8-99)
and the zero padding of a ragged batch to one N x C x Hmax x Wmax tensor (cirtorch/utils/sequence.py:4-67).
Host-side containers only; same names and behaviour as the reference's so its ``ImageRetrievalNet`` / data loaders
work with either."""
from __future__ import annotations

import torch


class PackedSequence:
    """A list of tensors (or None) of one dtype and device that may differ in their leading / spatial sizes."""

    def __init__(self, *args):
        items = args[0] if (len(args) == 1 and isinstance(args[0], list)) else list(args)
        real = [t for t in items if t is not None]
        if any(not isinstance(t, torch.Tensor) for t in real):
            raise TypeError("All args must be tensors")
        if len({t.dtype for t in real}) > 1:
            raise TypeError("All tensors must have the same type")
        if len({t.device for t in real}) > 1:
            raise TypeError("All tensors must reside on the same device")
        self._tensors = items
        self._compatible = len({tuple(t.shape[1:]) for t in real}) <= 1
        self._all_none = not real

    def __add__(self, other):
        if not isinstance(other, PackedSequence):
            raise TypeError("other must be a PackedSequence")
        return PackedSequence(list(self._tensors) + list(other._tensors))

    def __iadd__(self, other):
        if not isinstance(other, PackedSequence):
            raise TypeError("other must be a PackedSequence")
        self._tensors = list(self._tensors) + list(other._tensors)
        return self

    def __len__(self):
        return len(self._tensors)

    def __getitem__(self, item):
        return PackedSequence(list(self._tensors[item])) if isinstance(item, slice) else self._tensors[item]

    def __iter__(self):
        return iter(self._tensors)

    def _map(self, fn):
        self._tensors = [None if t is None else fn(t) for t in self._tensors]
        return self

    def cuda(self, device=None, non_blocking=False):
        return self._map(lambda t: t.cuda(device, non_blocking))

    def cpu(self):
        return self._map(lambda t: t.cpu())

    @property
    def all_none(self):
        return self._all_none

    def _first(self):
        return next((t for t in self._tensors if t is not None), None)

    @property
    def dtype(self):
        t = self._first()
        return None if t is None else t.dtype

    @property
    def device(self):
        t = self._first()
        return None if t is None else t.device

    @property
    def contiguous(self):
        """(all entries concatenated along dim 0, the index of the entry every row came from)."""
        if not self._compatible:
            raise ValueError("The tensors in the sequence are not compatible for contiguous view")
        if self._all_none:
            return None, None
        parts = [(i, t) for i, t in enumerate(self._tensors) if t is not None]
        return (torch.cat([t for _, t in parts], dim=0),
                torch.cat([t.new_full((t.size(0),), i, dtype=torch.long) for i, t in parts], dim=0))


def pad_packed_images(packed_images, pad_value=0., snap_size_to=None):
    """Ragged images (2-D or C x H x W, top-left aligned) -> (padded N x C x Hmax x Wmax tensor, list of valid sizes)."""
    if packed_images.all_none:
        raise ValueError("at least one image in packed_images should be non-None")
    real = [t for t in packed_images if t is not None]
    nd = real[0].dim()
    if nd not in (2, 3):
        raise ValueError("The input sequence must contain 2D or 3D tensors")
    if any(t.dim() != nd for t in real):
        raise ValueError("All tensors in the input sequence must have the same number of dimensions")
    if nd == 3 and any(t.shape[0] != real[0].shape[0] for t in real):
        raise ValueError("3D tensors must all have the same number of channels")
    hw = [max(t.shape[-2] for t in real), max(t.shape[-1] for t in real)]
    if snap_size_to is not None:
        hw = [(s + snap_size_to - 1) // snap_size_to * snap_size_to for s in hw]
    lead = [len(packed_images)] + ([real[0].shape[0]] if nd == 3 else [])
    padded = real[0].new_full(lead + hw, pad_value)
    sizes = []
    for i, t in enumerate(packed_images):
        if t is None:
            sizes.append((0, 0))
            continue
        padded[i, ..., :t.shape[-2], :t.shape[-1]] = t
        sizes.append(t.shape[-2:])
    return padded, sizes


def pack_padded_images(padded_images, sizes):
    """Inverse of :func:`pad_packed_images`."""
    return PackedSequence([img[..., :int(h), :int(w)].contiguous() for img, (h, w) in zip(padded_images, sizes)])
