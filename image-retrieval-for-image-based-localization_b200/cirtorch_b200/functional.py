"""Functional form of the fused descriptor tail (the upstream ``cirtorch.layers.functional`` names).

All functions take CUDA fp32 tensors and launch libcir_b200 kernels on the current stream.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import (CIR_POOL_GEM, CIR_POOL_MAC, CIR_POOL_SPOC, CIR_TAIL_ACCUMULATE, CIR_TAIL_HINT_INTEGER_P, CIR_TAIL_NO_WHITEN,
                   CIR_TAIL_POOL_ONLY)

_POOL = {"GeM": CIR_POOL_GEM, "GeMmp": CIR_POOL_GEM, "MAC": CIR_POOL_MAC, "SPoC": CIR_POOL_SPOC}


def _as_f32_contig(x: torch.Tensor) -> torch.Tensor:
    if x.dtype != torch.float32:
        x = x.float()
    return x if x.is_contiguous() else x.contiguous()


_tail_ws_bytes = {}


def _integer_p_hint(p) -> int:
    """CIR_TAIL_HINT_INTEGER_P if the (single, frozen) GeM exponent is a small integer.  The exponent lives on the device; it is
    read back ONCE per tensor object and version -- an inference-time constant -- and only picks the launch shape, never the
    result.  The answer is kept on the tensor itself (a dictionary keyed by address would outlive the tensor)."""
    hit = getattr(p, "_cir_integer_p", None)
    if hit is None or hit[0] != p._version:
        if torch.cuda.is_current_stream_capturing():
            return 0                  # no read-back inside a CUDA graph capture: the general launch shape
        hit = (p._version, float(p.detach().reshape(-1)[0]) in (1.0, 2.0, 3.0, 4.0))
        try:
            p._cir_integer_p = hit
        except AttributeError:        # a tensor type that takes no attributes: read back every time rather than guess
            pass
    return CIR_TAIL_HINT_INTEGER_P if hit[1] else 0


def _tail_launch(x, p, eps, weight, bias, pool_mode, flags, l2_eps=1e-6, out=None, pooled_out=None, z_out=None, allow_hint=True):
    """One cooperative launch of the fused tail.  Returns the physical [N, D] buffer.
    ``pooled_out`` / ``z_out``: optional contiguous fp32 [N, C] / [N, D] buffers that receive the pooled values and the
    projection before the last L2N (both kept for the backward pass; cir_tail_fwd_train).

    Kept lean on the host (a launch is ~0.1 ms of GPU time): no detach / reshape copies, one stream query."""
    _lib.require_cuda(x, p, weight, bias)
    lib = _lib.load()
    x = _as_f32_contig(x)
    if x.dim() != 4:
        raise ValueError("expected an N x C x H x W feature map, got shape %s" % (tuple(x.shape),))
    N, Cc, H, W = x.shape
    whiten = not (flags & (CIR_TAIL_NO_WHITEN | CIR_TAIL_POOL_ONLY))
    D = weight.shape[0] if whiten else Cc
    if out is None:
        if flags & CIR_TAIL_ACCUMULATE:
            raise ValueError("accumulate=True needs the buffer to add to (out=...)")
        out = torch.empty((N, D), dtype=torch.float32, device=x.device)
    elif (tuple(out.shape) != (N, D) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device):
        raise ValueError("out must be a contiguous fp32 %d x %d tensor on %s" % (N, D, x.device))
    if N == 0:
        return out
    p_stride = 0
    if p is not None:
        if p.dtype != torch.float32 or not p.is_contiguous():
            p = p.detach().float().contiguous()
        elif allow_hint and p.numel() == 1 and not (p.requires_grad and torch.is_grad_enabled()):
            flags |= _integer_p_hint(p)
        np_ = p.numel()
        if np_ not in (1, Cc):
            raise ValueError("GeM exponent must have 1 or C=%d elements, got %d" % (Cc, np_))
        p_stride = 0 if np_ == 1 else 1
    if whiten:
        if weight.dtype != torch.float32 or not weight.is_contiguous():
            weight = weight.detach().float().contiguous()
        if weight.shape[1] != Cc:
            raise ValueError("whitening weight is %s, feature map has %d channels" % (tuple(weight.shape), Cc))
        if bias is not None and (bias.dtype != torch.float32 or not bias.is_contiguous()):
            bias = bias.detach().float().contiguous()
    key = (N, Cc, D)
    need = _tail_ws_bytes.get(key)
    if need is None:
        import ctypes as C
        c_need = C.c_size_t(0)
        _lib.check(lib.cir_tail_workspace_bytes(N, Cc, D, C.byref(c_need)), "cir_tail_workspace_bytes")
        need = _tail_ws_bytes[key] = c_need.value
    ws = _lib.workspace(x.device, need, "tail")
    if z_out is not None:
        rc = lib.cir_tail_fwd_train(x.data_ptr(), N, Cc, H, W, p.data_ptr() if p is not None else None, p_stride, eps, l2_eps,
                                    pool_mode, weight.data_ptr() if whiten else None,
                                    bias.data_ptr() if (whiten and bias is not None) else None, D,
                                    out.data_ptr(), D, pooled_out.data_ptr() if pooled_out is not None else None,
                                    z_out.data_ptr(), ws.data_ptr(), ws.numel(), flags,
                                    torch.cuda.current_stream(x.device).cuda_stream)
    else:
        rc = lib.cir_tail_fwd(x.data_ptr(), N, Cc, H, W, p.data_ptr() if p is not None else None, p_stride, eps, l2_eps,
                              pool_mode, weight.data_ptr() if whiten else None,
                              bias.data_ptr() if (whiten and bias is not None) else None, D,
                              out.data_ptr(), D, pooled_out.data_ptr() if pooled_out is not None else None,
                              ws.data_ptr(), ws.numel(), flags, torch.cuda.current_stream(x.device).cuda_stream)
    if rc:
        _lib.check(rc, "cir_tail_fwd")
    return out


def _tail_reference_formula(x, p, eps, weight, bias, pooling, flags, l2_eps):
    """The tail written with stock differentiable torch ops -- used ONLY to obtain gradients
    (recompute-in-backward); the forward values always come from the CUDA kernel."""
    if pooling in ("GeM", "GeMmp"):
        pp = p.reshape(1, -1, 1, 1) if p.numel() > 1 else p
        v = x.clamp(min=eps).pow(pp).mean(dim=(-2, -1)).pow(1.0 / (p if p.numel() > 1 else p))
    elif pooling == "MAC":
        v = x.amax(dim=(-2, -1))
    else:
        v = x.mean(dim=(-2, -1))
    if flags & CIR_TAIL_POOL_ONLY:
        return v
    v = v / (v.norm(p=2, dim=1, keepdim=True) + l2_eps)
    if not (flags & CIR_TAIL_NO_WHITEN):
        v = torch.nn.functional.linear(v, weight, bias)
        v = v / (v.norm(p=2, dim=1, keepdim=True) + l2_eps)
    return v


def _gem_bwd_launch(x, p, eps, g, dg, want_dx, want_s):
    """cir_gem_bwd: the N*C*H*W-sized part of the backward (one read of x, one write of dx)."""
    lib = _lib.load()
    N, Cc, H, W = x.shape
    dx = torch.empty_like(x) if want_dx else None
    S = torch.empty((N, Cc), dtype=torch.float32, device=x.device) if want_s else None
    pc = p.detach().reshape(-1).float().contiguous()
    rc = lib.cir_gem_bwd(_lib.ptr(x), N, Cc, H, W, _lib.ptr(pc), 0 if pc.numel() == 1 else 1, float(eps),
                         _lib.ptr(g), _lib.ptr(dg), _lib.ptr(dx), _lib.ptr(S), _lib.stream_of(x))
    _lib.check(rc, "cir_gem_bwd")
    return dx, S


def _l2n_bwd_rows(v, gout, l2_eps, want_unit=False):
    """cir_l2n_bwd_rows: gradient of v / (||v|| + eps) w.r.t. v applied to gout (rows); optionally also the forward value."""
    lib = _lib.load()
    N, Cc = v.shape
    out = torch.empty_like(v) if gout is not None else None
    unit = torch.empty_like(v) if want_unit else None
    rc = lib.cir_l2n_bwd_rows(_lib.ptr(v), _lib.ptr(gout), N, Cc, float(l2_eps), _lib.ptr(out), _lib.ptr(unit), _lib.stream_of(v))
    _lib.check(rc, "cir_l2n_bwd_rows")
    return out, unit


def _colsum_rows(x2d):
    lib = _lib.load()
    out = torch.empty((x2d.shape[1],), dtype=torch.float32, device=x2d.device)
    _lib.check(lib.cir_colsum_rows(_lib.ptr(x2d), x2d.shape[0], x2d.shape[1], _lib.ptr(out), _lib.stream_of(x2d)), "cir_colsum_rows")
    return out


_NO_ZOUT = bool(int(os.environ.get("CIR_DEBUG_NO_ZOUT", "0")))      # experiments: recompute the projection in backward


class _TailFn(torch.autograd.Function):
    """Forward = the fused CUDA kernel (which also hands back the pooled values g).  Backward for GeM pooling: two
    row kernels for the L2N gradients (cir_l2n_bwd_rows), the three [N, C] x [C, D]-sized products of the Linear through
    cuBLAS (plain library GEMMs on 64-row operands), dL/db by cir_colsum_rows, ONE streaming kernel (cir_gem_bwd) for
    everything that touches the feature map (dx and the sum needed for dL/dp) and cir_gem_dp.  MAC / SPoC recompute the
    tail with differentiable torch ops."""

    @staticmethod
    def forward(ctx, x, p, weight, bias, eps, pooling, flags, l2_eps):
        ctx.cfg = (eps, pooling, flags, l2_eps)
        xc = _as_f32_contig(x)
        g = z = None
        gem_path = pooling in ("GeM", "GeMmp") and xc.shape[0] > 0
        if gem_path and not (flags & CIR_TAIL_POOL_ONLY):
            g = torch.empty((xc.shape[0], xc.shape[1]), dtype=torch.float32, device=xc.device)
            if not (flags & CIR_TAIL_NO_WHITEN) and not _NO_ZOUT:
                # the projection before the last L2N, stored by the kernel's last phase: the backward would otherwise recompute
                # it with a 64 x 2048 x 2048 fp32 GEMM (51 us of the 0.49 ms forward + backward)
                z = torch.empty((xc.shape[0], weight.shape[0]), dtype=torch.float32, device=xc.device)
        # training: the exponent moves with every optimiser step -- no read-back, the general launch shape
        out = _tail_launch(xc, p, eps, weight, bias, _POOL[pooling], flags, l2_eps, pooled_out=g, z_out=z, allow_hint=False)
        if gem_path and (flags & CIR_TAIL_POOL_ONLY):
            g = out
        ctx.save_for_backward(xc, p, weight, bias, g, z)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, p, weight, bias, g, z = ctx.saved_tensors
        eps, pooling, flags, l2_eps = ctx.cfg
        need = ctx.needs_input_grad
        if g is None:
            return _TailFn._backward_recompute(ctx, gout)
        gout = gout.contiguous().float()
        dW = db = None
        with torch.no_grad():
            if flags & CIR_TAIL_POOL_ONLY:
                dg = gout
            elif flags & CIR_TAIL_NO_WHITEN:
                dg, _ = _l2n_bwd_rows(g, gout, l2_eps)
            else:
                _, u = _l2n_bwd_rows(g, None, l2_eps, want_unit=True)          # u = g / (||g|| + eps)
                if z is None:
                    z = torch.addmm(bias, u, weight.t()) if bias is not None else u @ weight.t()
                dz, _ = _l2n_bwd_rows(z, gout, l2_eps)
                if need[2]:
                    dW = dz.t() @ u
                if bias is not None and need[3]:
                    db = _colsum_rows(dz)
                dg, _ = _l2n_bwd_rows(g, dz @ weight, l2_eps)
            dx = dp = None
            if need[0] or need[1]:
                dx, S = _gem_bwd_launch(x, p, eps, g, dg, need[0], need[1])
                if need[1]:
                    lib = _lib.load()
                    pc = p.detach().reshape(-1).float().contiguous()
                    dp = torch.empty_like(pc)
                    ws = _lib.workspace(g.device, g.shape[0] * 4 + 64, "gem_dp")
                    rc = lib.cir_gem_dp(_lib.ptr(g), _lib.ptr(dg), _lib.ptr(S), _lib.ptr(pc), 0 if pc.numel() == 1 else 1,
                                        g.shape[0], g.shape[1], x.shape[2] * x.shape[3], _lib.ptr(dp), _lib.ptr(ws), ws.numel(),
                                        _lib.stream_of(g))
                    _lib.check(rc, "cir_gem_dp")
                    dp = dp.reshape(p.shape)
        return dx, dp, dW, db, None, None, None, None

    @staticmethod
    def _backward_recompute(ctx, g):
        x, p, weight, bias, _, _ = ctx.saved_tensors
        eps, pooling, flags, l2_eps = ctx.cfg
        ins = []
        with torch.enable_grad():
            leaves = []
            for i, t in enumerate((x, p, weight, bias)):
                if t is None:
                    leaves.append(None)
                    continue
                t2 = t.detach().requires_grad_(ctx.needs_input_grad[i])
                leaves.append(t2)
                if ctx.needs_input_grad[i]:
                    ins.append(t2)
            y = _tail_reference_formula(leaves[0].float(), leaves[1], eps, leaves[2], leaves[3], pooling, flags, l2_eps)
            grads = torch.autograd.grad(y, ins, g.contiguous(), allow_unused=True) if ins else ()
        it = iter(grads)
        out = [next(it) if (t is not None and ctx.needs_input_grad[i]) else None
               for i, t in enumerate((x, p, weight, bias))]
        return (*out, None, None, None, None)


def descriptor_tail(x, p=None, eps=1e-6, weight=None, bias=None, pooling="GeM", do_whitening=True,
                    pool_only=False, l2_eps=1e-6, out=None, accumulate=False):
    """pool -> L2N -> (whiten -> L2N).  Returns the physical N x D buffer (row = descriptor).

    globalHead.forward (cirtorch/modules/heads/global_head.py:52-67) returns its transpose view.
    ``out``: write into this N x D buffer; with ``accumulate`` the descriptors are ADDED to it in the kernel's last phase
    (the sum over scales of the fork's multi-scale mean, GF_net.py:74-92) -- inference only.
    """
    if pooling not in _POOL:
        raise KeyError(pooling)
    flags = CIR_TAIL_POOL_ONLY if pool_only else (0 if do_whitening else CIR_TAIL_NO_WHITEN)
    if accumulate:
        if pool_only:
            raise ValueError("accumulate applies to descriptors, not to pool_only")
        flags |= CIR_TAIL_ACCUMULATE
    if pooling in ("GeM", "GeMmp"):
        if p is None:
            raise ValueError("GeM pooling needs the exponent p")
        if not torch.is_tensor(p):
            pf = float(p)
            p = torch.full((1,), pf, dtype=torch.float32, device=x.device)
            p._cir_integer_p = (p._version, pf in (1.0, 2.0, 3.0, 4.0))                 # known on the host: no read-back
    else:
        p = None
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and torch.is_tensor(t) and t.requires_grad for t in (x, p, weight, bias))
    if needs_grad:
        if out is not None:
            raise ValueError("out= / accumulate= are inference-only (no autograd through an in-place sum)")
        return _TailFn.apply(x, p, weight if do_whitening and not pool_only else None,
                             bias if do_whitening and not pool_only else None, eps, pooling, flags, l2_eps)
    return _tail_launch(x, p, eps, weight, bias, _POOL[pooling], flags, l2_eps, out=out)


def rmac_grid(H, W, L=3):
    """Per level l = 1..L of the R-MAC grid: (wl, cenH, cenW) = region side, top rows, left columns
    (cirtorch/modules/pools.py:126-160 / :64-103).  Six-element HOST arithmetic, done with the same float32 torch
    ops in the same order as the reference so that near-ties (e.g. an 11 x 47 map) round the same way."""
    import math
    short, long_ = min(W, H), max(W, H)
    n_extra = torch.tensor([2., 3., 4., 5., 6., 7.])          # candidate region counts along the long side
    stride = (long_ - short) / (n_extra - 1)
    overlap = (short ** 2 - short * stride) / short ** 2
    best = int(torch.min(torch.abs(overlap - 0.4), 0)[1])      # desired overlap of neighbouring regions: 0.4
    Wd = best + 1 if H < W else 0
    Hd = best + 1 if H > W else 0
    levels = []
    for l in range(1, L + 1):
        wl = math.floor(2 * short / (l + 1))
        wl2 = math.floor(wl / 2 - 1)

        def starts(extent, extra):
            n = l + extra
            step = 0 if n == 1 else (extent - wl) / (n - 1)
            return (torch.floor(wl2 + torch.arange(n, dtype=torch.float32) * step).int() - wl2).tolist()

        levels.append((int(wl), starts(H, Hd), starts(W, Wd)))
    return levels


def rmac_regions(H, W, L=3):
    """All regions of Rpool.roipool in its order (levels, rows outer / columns inner) as (row0, col0, height, width);
    the whole map (roipool's first vector) is NOT included."""
    return [(i, j, wl, wl) for (wl, cenH, cenW) in rmac_grid(H, W, L) if wl > 0 for i in cenH for j in cenW]


def _region_pool_launch(x, regions, p, eps, pooling):
    import ctypes as C
    lib = _lib.load()
    N, Cc, H, W = x.shape
    R = len(regions)
    p_stride = 0
    if p is not None:
        p = p.detach().reshape(-1).float().contiguous()
        p_stride = 0 if p.numel() == 1 else 1
    flat = [int(v) for r in regions for v in r]
    reg = (C.c_int32 * len(flat))(*flat)
    out = torch.empty((N, R, Cc), dtype=torch.float32, device=x.device)
    rc = lib.cir_region_pool(_lib.ptr(x), N, Cc, H, W, C.cast(reg, C.c_void_p), R, _lib.ptr(p), p_stride, float(eps),
                             _POOL[pooling], _lib.ptr(out), _lib.stream_of(x))
    _lib.check(rc, "cir_region_pool")
    return out


def _region_pool_formula(x, regions, p, eps, pooling):
    """The same pooling with stock differentiable torch ops (pools.py:126-167 region by region) -- gradients only."""
    vecs = []
    for (i, j, h, w) in regions:
        r = x[:, :, i:i + h, j:j + w]
        if pooling in ("GeM", "GeMmp"):
            pp = p.reshape(1, -1, 1, 1) if p.numel() > 1 else p
            v = r.clamp(min=eps).pow(pp).mean(dim=(-2, -1)).pow(1.0 / p)
        elif pooling == "MAC":
            v = r.amax(dim=(-2, -1))
        else:
            v = r.mean(dim=(-2, -1))
        vecs.append(v)
    return torch.stack(vecs, dim=1)


class _RegionPoolFn(torch.autograd.Function):
    """Forward = cir_region_pool (one pass over the map); backward recomputes the regions with differentiable torch ops."""

    @staticmethod
    def forward(ctx, x, p, regions, eps, pooling):
        ctx.cfg = (regions, eps, pooling)
        ctx.save_for_backward(x, p)
        return _region_pool_launch(x, regions, p, eps, pooling)

    @staticmethod
    def backward(ctx, gout):
        x, p = ctx.saved_tensors
        regions, eps, pooling = ctx.cfg
        with torch.enable_grad():
            xl = x.detach().requires_grad_(ctx.needs_input_grad[0])
            pl = None if p is None else p.detach().requires_grad_(ctx.needs_input_grad[1])
            ins = [t for t, need in ((xl, ctx.needs_input_grad[0]), (pl, p is not None and ctx.needs_input_grad[1])) if need]
            grads = torch.autograd.grad(_region_pool_formula(xl, regions, pl, eps, pooling), ins, gout.contiguous()) if ins else ()
        it = iter(grads)
        gx = next(it) if ctx.needs_input_grad[0] else None
        gp = next(it) if (p is not None and ctx.needs_input_grad[1]) else None
        return gx, gp, None, None, None


def region_pool(x, regions, p=None, eps=1e-6, pooling="GeM"):
    """Pool every (row0, col0, height, width) region of an N x C x H x W map in ONE pass over the map
    (cir_region_pool) -> N x R x C.  ``pooling`` / ``p`` / ``eps`` as in descriptor_tail.  Differentiable in x and p."""
    if pooling not in _POOL:
        raise KeyError(pooling)
    _lib.require_cuda(x, p if torch.is_tensor(p) else None)
    if x.dim() != 4:
        raise ValueError("expected an N x C x H x W feature map, got shape %s" % (tuple(x.shape),))
    Cc = x.shape[1]
    if len(regions) < 1:
        raise ValueError("region_pool needs at least one region")
    for r in regions:
        if len(r) != 4:
            raise ValueError("a region is (row0, col0, height, width)")
    if pooling in ("GeM", "GeMmp"):
        if p is None:
            raise ValueError("GeM pooling needs the exponent p")
        if not torch.is_tensor(p):
            p = torch.full((1,), float(p), dtype=torch.float32, device=x.device)
        if p.numel() not in (1, Cc):
            raise ValueError("GeM exponent must have 1 or C=%d elements, got %d" % (Cc, p.numel()))
    else:
        p = None
    regions = [tuple(int(v) for v in r) for r in regions]
    if torch.is_grad_enabled() and (x.requires_grad or (p is not None and p.requires_grad)):
        return _RegionPoolFn.apply(_as_f32_contig(x), p, regions, eps, pooling)
    return _region_pool_launch(_as_f32_contig(x), regions, p, eps, pooling)


def gem(x, p=3, eps=1e-6):
    """GeM pooling, cirtorch/modules/pools.py:37-38 -> N x C x 1 x 1."""
    return descriptor_tail(x, p=p, eps=eps, pooling="GeM", pool_only=True)[:, :, None, None]


def mac(x):
    """cirtorch/modules/pools.py:15."""
    return descriptor_tail(x, pooling="MAC", pool_only=True)[:, :, None, None]


def spoc(x):
    """cirtorch/modules/pools.py:25."""
    return descriptor_tail(x, pooling="SPoC", pool_only=True)[:, :, None, None]


class _L2NFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, eps):
        ctx.eps = eps
        ctx.save_for_backward(x2d)
        return _l2n_rows(x2d, eps)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        with torch.enable_grad():
            x2 = x.detach().requires_grad_(True)
            y = x2 / (x2.norm(p=2, dim=1, keepdim=True) + ctx.eps)
            (gx,) = torch.autograd.grad(y, x2, g)
        return gx, None


def _l2n_rows(x2d, eps):
    _lib.require_cuda(x2d)
    lib = _lib.load()
    x2d = _as_f32_contig(x2d)
    out = torch.empty_like(x2d)
    if x2d.numel():
        rc = lib.cir_l2n_rows(_lib.ptr(x2d), x2d.shape[0], x2d.shape[1], x2d.shape[1], float(eps),
                              _lib.ptr(out), x2d.shape[1], _lib.stream_of(x2d))
        _lib.check(rc, "cir_l2n_rows")
    return out


def l2n(x, eps=1e-6):
    """x / (||x||_2 over dim 1 + eps), cirtorch/modules/normalizations.py:15-16.  Any rank >= 2."""
    if x.dim() < 2:
        raise ValueError("l2n expects at least 2 dimensions")
    shape = x.shape
    if x.dim() == 2 or all(s == 1 for s in shape[2:]):
        x2d = x.reshape(shape[0], shape[1])
        moved = None
    else:   # norm over channels at every spatial position: rows = (n, h, w)
        moved = x.movedim(1, -1)
        x2d = moved.reshape(-1, shape[1])
    y = _L2NFn.apply(x2d, eps) if (torch.is_grad_enabled() and x.requires_grad) else _l2n_rows(x2d, eps)
    if moved is None:
        return y.reshape(shape)
    return y.reshape(moved.shape).movedim(-1, 1)


def _powerlaw_launch(x, eps):
    lib = _lib.load()
    xc = _as_f32_contig(x)
    out = torch.empty_like(xc)
    rc = lib.cir_powerlaw(_lib.ptr(xc), xc.numel(), float(eps), _lib.ptr(out), _lib.stream_of(xc))
    _lib.check(rc, "cir_powerlaw")
    return out


class _PowerLawFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        ctx.eps = eps
        ctx.save_for_backward(x)
        return _powerlaw_launch(x, eps)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        with torch.enable_grad():
            x2 = x.detach().requires_grad_(True)
            t = x2 + ctx.eps
            (gx,) = torch.autograd.grad(t.abs().sqrt().mul(t.sign()), x2, g)       # normalizations.py:25-27
        return gx, None


def powerlaw(x, eps=1e-6):
    """sign(x + eps) * sqrt(|x + eps|), cirtorch/modules/normalizations.py:25-27.  Differentiable in x."""
    _lib.require_cuda(x)
    if torch.is_grad_enabled() and x.requires_grad:
        return _PowerLawFn.apply(x, eps)
    return _powerlaw_launch(x, eps)
