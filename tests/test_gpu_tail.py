"""Fused descriptor tail (cir_tail_fwd through the reference-shaped modules) vs the CPU oracle and
the committed golden fixtures.  Tolerance: 1e-4 relative per descriptor (north-star, fp32)."""
import numpy as np
import pytest
import torch

from oracle import cirtorch_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-4


def _rel(a, b, dim=0):
    return float(((a - b).norm(dim=dim) / b.norm(dim=dim).clamp_min(1e-30)).max())


def _head(dim, pooling="GeM", p=3.0):
    from cirtorch_b200.modules.heads.global_head import globalHead
    params = {"p": p, "eps": 1e-6} if pooling in ("GeM", "GeMmp") else ({"L": 3} if pooling == "RMAC" else {})
    return globalHead(pooling={"name": pooling, "params": params}, normal={"name": "L2N", "params": {}}, dim=dim)


def test_golden_fixtures(golden):
    from cirtorch_b200.modules.pools import GeM, MAC, SPoC
    from cirtorch_b200.modules.normalizations import L2N
    g = golden("tail")
    for c in "abcd":
        x = torch.from_numpy(g[f"{c}_x"]).to(DEV)
        p = float(g[f"{c}_p"])
        head = _head(x.shape[1], p=p).to(DEV)
        with torch.no_grad():
            head.whiten.weight.copy_(torch.from_numpy(g[f"{c}_W"]))
            head.whiten.bias.copy_(torch.from_numpy(g[f"{c}_b"]))
            gem = GeM(p=p).to(DEV)(x)
            assert gem.shape == g[f"{c}_gem"].shape
            np.testing.assert_allclose(gem.cpu().numpy(), g[f"{c}_gem"], rtol=RTOL, atol=1e-7)
            np.testing.assert_allclose(L2N()(gem).cpu().numpy(), g[f"{c}_l2n"], rtol=RTOL, atol=1e-7)
            np.testing.assert_array_equal(MAC()(x).cpu().numpy(), g[f"{c}_mac"])
            np.testing.assert_allclose(SPoC()(x).cpu().numpy(), g[f"{c}_spoc"], rtol=1e-5, atol=1e-8)
            out = head(x)
            assert out.shape == g[f"{c}_head"].shape
            assert _rel(out.cpu(), torch.from_numpy(g[f"{c}_head"])) < RTOL
            out2 = head(x, do_whitening=False)
            assert _rel(out2.cpu(), torch.from_numpy(g[f"{c}_head_nowhiten"])) < RTOL


@pytest.mark.parametrize("shape,p", [
    ((64, 2048, 8, 8), 3.0),       # config-2 channel count, small maps
    ((5, 2048, 32, 32), 3.0),      # config-2 map size
    ((3, 2048, 32, 32), 2.7),      # learnable, non-integer exponent
    ((7, 512, 19, 23), 3.0),       # HW % 4 != 0: scalar path
    ((130, 96, 4, 4), 2.0),        # more images than one phase-B pass (64)
    ((1, 64, 1, 1), 3.0),
    ((2, 64, 80, 80), 3.0),        # rows of 25.6 KB: larger than a TMA ring slot -> direct-load path
    ((3, 4096, 8, 8), 3.0),        # C > 2048: two K tiles of the projection weights
    ((70, 32, 2, 2), 4.0),         # tiny rows, many per ring slot
])
def test_head_vs_oracle(shape, p):
    torch.manual_seed(0)
    n, c, h, w = shape
    x = torch.relu(torch.randn(shape))
    head = _head(c, p=p)
    with torch.no_grad():
        head.whiten.bias.normal_(0, 0.02)
    ref = O.head_forward(x, p, 1e-6, head.whiten.weight.detach(), head.whiten.bias.detach())
    ref_nw = O.head_forward(x, p, 1e-6, do_whitening=False)
    head = head.to(DEV)
    with torch.no_grad():
        out = head(x.to(DEV))
        out_nw = head(x.to(DEV), do_whitening=False)
    assert out.shape == (c, n) and out.t().is_contiguous()
    assert _rel(out.cpu(), ref) < RTOL
    assert _rel(out_nw.cpu(), ref_nw) < RTOL
    np.testing.assert_allclose(out.norm(dim=0).cpu().numpy(), 1.0, atol=1e-4)


def test_pooling_variants():
    torch.manual_seed(1)
    x = torch.relu(torch.randn(4, 256, 16, 16))
    for pooling in ("MAC", "SPoC"):
        head = _head(256, pooling)
        ref = O.head_forward(x, None, 1e-6, head.whiten.weight.detach(), head.whiten.bias.detach(), pooling=pooling)
        with torch.no_grad():
            out = head.to(DEV)(x.to(DEV))
        assert _rel(out.cpu(), ref) < RTOL
    head = _head(256, "GeMmp", p=3.0)
    with torch.no_grad():
        head.pool.p.copy_(torch.linspace(1.5, 4.5, 256))
    ref = O.head_forward(x, head.pool.p.detach(), 1e-6, head.whiten.weight.detach(), head.whiten.bias.detach())
    with torch.no_grad():
        out = head.to(DEV)(x.to(DEV))
    assert _rel(out.cpu(), ref) < RTOL


def test_regional_pooling_golden(golden):
    """Rpool / RMAC on the one-pass region kernel vs the fixtures produced by the reference (pools.py:57-197)."""
    from cirtorch_b200.modules.pools import GeM, MAC, SPoC, RMAC, Rpool, POOLING_LAYERS
    assert POOLING_LAYERS["RMAC"] is RMAC and POOLING_LAYERS["ROIpool"] is Rpool
    g = golden("regional")
    for name in ("sq", "wide", "tall", "tiny"):
        x = torch.from_numpy(g[f"{name}_x"]).to(DEV)
        C = x.shape[1]
        lin = torch.nn.Linear(C, C).to(DEV)
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(g[f"{name}_W"]))
            lin.bias.copy_(torch.from_numpy(g[f"{name}_b"]))
            for L in (1, 2, 3):
                for key, mod, kw in ((f"{name}_gem3_L{L}", Rpool(GeM(p=3), L=L), {}),
                                     (f"{name}_gem25_white_L{L}", Rpool(GeM(p=2.5), whiten=lin, L=L), {}),
                                     (f"{name}_mac_L{L}", Rpool(MAC(), L=L), {}),
                                     (f"{name}_spoc_regions_L{L}", Rpool(SPoC(), L=L), {"aggregate": False})):
                    out = mod.to(DEV)(x, **kw)
                    want = torch.from_numpy(g[key])
                    assert out.shape == want.shape, key
                    D = want.shape[-3]           # per-descriptor relative error (north-star bar), elementwise with a floor
                    assert _rel(out.cpu().reshape(-1, D), want.reshape(-1, D), dim=1) < RTOL, key
                    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-3, atol=2e-6, err_msg=key)
                want = g[f"{name}_rmac_L{L}"]
                if want.size == 0:
                    with pytest.raises(NameError):
                        RMAC(L=L)(x)
                else:
                    out = RMAC(L=L)(x)
                    assert out.shape == want.shape
                    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-5, atol=1e-7)


def test_regional_pooling_vs_oracle_layer4_shape():
    """ResNet layer-4 sized maps (2048 x 32 x 32 and a non-square 24 x 40), per-channel exponents, a generic rpool callable."""
    from cirtorch_b200.modules.pools import GeM, GeMmp, Rpool
    from cirtorch_b200 import functional as LF
    torch.manual_seed(3)
    for shape in ((3, 2048, 32, 32), (2, 256, 24, 40)):
        x = torch.relu(torch.randn(*shape))
        C = shape[1]
        lin = torch.nn.Linear(C, C)
        ref = O.rpool_forward(x, lambda t: O.gem(t, 3.0), lin.weight.detach(), lin.bias.detach(), L=3)
        with torch.no_grad():
            out = Rpool(GeM(p=3), whiten=lin.to(DEV), L=3).to(DEV)(x.to(DEV))
        assert _rel(out.cpu().reshape(shape[0], C).t(), ref.reshape(shape[0], C).t()) < RTOL
        pc = 2.0 + 2.0 * torch.rand(C)
        ref = O.rpool_forward(x, lambda t: O.gem(t, pc), L=2)
        mp = GeMmp(p=3, mp=C)
        with torch.no_grad():
            mp.p.copy_(pc)
            out = Rpool(mp.to(DEV), L=2)(x.to(DEV))
            # any other callable takes the region-by-region composition
            out2 = Rpool(lambda t: LF.gem(t, p=3.0), L=2)(x.to(DEV))
        assert _rel(out.cpu().reshape(shape[0], C).t(), ref.reshape(shape[0], C).t()) < RTOL
        ref2 = O.rpool_forward(x, lambda t: O.gem(t, 3.0), L=2)
        assert _rel(out2.cpu().reshape(shape[0], C).t(), ref2.reshape(shape[0], C).t()) < RTOL
    with pytest.raises(ValueError):
        LF.region_pool(torch.zeros(1, 4, 8, 8, device=DEV), [(0, 0, 9, 8)], pooling="MAC")


def test_accumulate_flag_and_multiscale_sum(golden):
    """CIR_TAIL_ACCUMULATE: descriptors added in the kernel's last phase == the fork's multi-scale mean (GF_net.py:74-92,
    fixture produced by the reference's avg_pool1d), for whitened and un-whitened heads; inference only."""
    from cirtorch_b200.extract import ImageRetrievalNet
    from cirtorch_b200 import functional as LF
    torch.manual_seed(5)
    for do_whitening in (True, False):
        head = _head(96).to(DEV).eval()
        xs = [torch.relu(torch.randn(5, 96, h, w)) for (h, w) in ((8, 8), (6, 5), (11, 7))]
        refs = [O.head_forward(x, 3.0, 1e-6, head.whiten.weight.detach().cpu(), head.whiten.bias.detach().cpu(),
                               do_whitening=do_whitening) for x in xs]
        want = O.multiscale_mean(refs)
        with torch.no_grad():
            acc = head(xs[0].to(DEV), do_whitening=do_whitening)
            for x in xs[1:]:
                ret = head(x.to(DEV), do_whitening=do_whitening, out=acc, accumulate=True)
                assert ret.data_ptr() == acc.data_ptr()
            acc.mul_(1.0 / len(xs))
        assert _rel(acc.cpu(), want) < RTOL
    g = golden("multiscale")           # [S, D, N] per-scale descriptors -> [D, N] from the reference
    np.testing.assert_allclose(O.multiscale_mean([torch.from_numpy(p_) for p_ in g["preds"]]).numpy(), g["out"], rtol=1e-6)

    class Body(torch.nn.Module):       # a stand-in backbone: fixed 1x1 conv + ReLU, output follows the image size
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 96, 1)

        def forward(self, img):
            return {"mod5": torch.relu(self.conv(img))}

    net = ImageRetrievalNet(Body(), _head(96)).to(DEV).eval()
    img = torch.randn(4, 3, 32, 24, device=DEV)
    scales = (1, 0.5, 2 ** -0.5)
    with torch.no_grad():
        got = net(img, scales=scales)
        per_scale = [net.ret_head(net.body(net._rescale(img, s))["mod5"]).clone() for s in scales]
    want = torch.stack(per_scale, 0).mean(0)
    assert _rel(got.cpu(), want.cpu()) < 1e-6
    assert float(got.norm(dim=0).max()) <= 1.0 + 1e-5
    with pytest.raises(ValueError):
        LF.descriptor_tail(torch.zeros(2, 8, 4, 4, device=DEV), p=3.0, do_whitening=False, accumulate=True)


def test_global_head_with_rmac_pooling():
    """Any registered pooling goes through globalHead: RMAC (its own one-pass kernel) -> fused L2N / whiten / L2N."""
    torch.manual_seed(6)
    x = torch.relu(torch.randn(3, 64, 12, 10))
    head = _head(64, pooling="RMAC").to(DEV).eval()
    with torch.no_grad():
        got = head(x.to(DEV))
    v = O.rmac_forward(x, L=3)
    want = O.l2n(torch.nn.functional.linear(O.l2n(v.reshape(3, 64)), head.whiten.weight.detach().cpu(), head.whiten.bias.detach().cpu()))
    assert got.shape == (64, 3)
    assert _rel(got.cpu(), want.t()) < RTOL


def test_state_dict_keys_and_empty_batch():
    head = _head(64)
    assert set(head.state_dict().keys()) == {"pool.p", "whiten.weight", "whiten.bias"}
    head = head.to(DEV)
    out = head(torch.zeros(0, 64, 4, 4, device=DEV))
    assert out.shape == (64, 0)


def test_l2n_general_rank():
    from cirtorch_b200.modules.normalizations import L2N
    torch.manual_seed(2)
    x = torch.randn(3, 40, 5, 6)
    ref = O.l2n(x)
    out = L2N()(x.to(DEV))
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-7)


def test_powerlaw_matches_reference_formula():
    from cirtorch_b200.modules.normalizations import NORMALIZATION_LAYERS
    x = torch.randn(4, 33, 3, 5)
    x[0, 0, 0, 0] = -1e-6                      # x + eps == 0
    ref = (x + 1e-6).abs().sqrt().mul((x + 1e-6).sign())          # normalizations.py:25-27
    out = NORMALIZATION_LAYERS["PowerLaw"]()(x.to(DEV))
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-6, atol=1e-9)


def test_gradients_match_autograd():
    """Training: p, W, b and x receive the gradients of the reference formula."""
    torch.manual_seed(3)
    x = torch.relu(torch.randn(6, 32, 5, 5)) + 0.01
    head = _head(32, p=2.5)
    xr = x.clone().requires_grad_(True)
    ref = O.head_forward(xr, head.pool.p, 1e-6, head.whiten.weight, head.whiten.bias)
    tgt = torch.randn_like(ref)
    (ref * tgt).sum().backward()
    gref = [xr.grad.clone(), head.pool.p.grad.clone(), head.whiten.weight.grad.clone(), head.whiten.bias.grad.clone()]
    head.zero_grad()
    head = head.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    out = head(xg)
    (out * tgt.to(DEV)).sum().backward()
    got = [xg.grad, head.pool.p.grad, head.whiten.weight.grad, head.whiten.bias.grad]
    for a, b in zip(got, gref):     # error relative to the largest gradient entry of each tensor
        scale = float(b.abs().max()) + 1e-12
        assert float((a.cpu() - b).abs().max()) / scale < 2e-3


@pytest.mark.parametrize("pooling,p,mode", [("GeM", 3.0, "full"), ("GeM", 2.5, "nowhiten"), ("GeMmp", 3.0, "full"),
                                            ("GeM", 2.2, "pool"), ("MAC", None, "full")])
def test_gradients_all_modes(pooling, p, mode):
    """cir_gem_bwd + the [N, C]-sized chain vs autograd of the oracle formula, every mode of the tail."""
    from cirtorch_b200 import functional as LF
    torch.manual_seed(5)
    x = torch.relu(torch.randn(5, 64, 6, 8)) + 0.02
    x[0, 0] = 0.0                       # a fully clamped row: zero gradient, finite dp
    head = _head(64, pooling, p=p if p else 3.0)
    if pooling == "GeMmp":
        with torch.no_grad():
            head.pool.p.copy_(torch.linspace(2.0, 4.0, 64))
    pt = getattr(head.pool, "p", None)

    def run(xin, pp, W, b, device):
        kw = dict(p=pp, eps=1e-6, weight=W, bias=b, pooling=pooling)
        if device == "ref":
            if mode == "pool":
                return O.gem(xin, pp).flatten(1)
            return O.head_forward(xin, pp, 1e-6, W, b, do_whitening=(mode == "full"), pooling=pooling).t()
        return LF.descriptor_tail(xin, do_whitening=(mode == "full"), pool_only=(mode == "pool"), **kw)

    xr = x.clone().requires_grad_(True)
    ref = run(xr, pt, head.whiten.weight, head.whiten.bias, "ref")
    tgt = torch.randn_like(ref)
    (ref * tgt).sum().backward()
    gref = {"x": xr.grad.clone()}
    if pt is not None:
        gref["p"] = pt.grad.clone()
    if mode == "full":
        gref["W"], gref["b"] = head.whiten.weight.grad.clone(), head.whiten.bias.grad.clone()
    head.zero_grad()
    head = head.to(DEV)
    pt = getattr(head.pool, "p", None)
    xg = x.to(DEV).requires_grad_(True)
    out = run(xg, pt, head.whiten.weight, head.whiten.bias, "dev")
    (out * tgt.to(DEV)).sum().backward()
    got = {"x": xg.grad}
    if pt is not None:
        got["p"] = pt.grad
    if mode == "full":
        got["W"], got["b"] = head.whiten.weight.grad, head.whiten.bias.grad
    for k in gref:
        scale = float(gref[k].abs().max()) + 1e-12
        err = float((got[k].cpu() - gref[k]).abs().max()) / scale
        assert err < 2e-3, (k, err)


def test_training_forward_saves_projection():
    """cir_tail_fwd_train: the descriptors are those of cir_tail_fwd (bit-equal), pooled_out = the GeM values and z_out = the
    projection before the last L2N (global_head.py:62-64), so that out == L2N(z_out)."""
    from cirtorch_b200 import functional as LF
    torch.manual_seed(9)
    x = (torch.relu(torch.randn(7, 256, 12, 16)) + 0.01).to(DEV)
    head = _head(256, "GeM", p=2.6).to(DEV)
    W, b, pt = head.whiten.weight.detach(), head.whiten.bias.detach(), head.pool.p.detach()
    g = torch.empty((7, 256), device=DEV)
    z = torch.empty((7, 256), device=DEV)
    out_t = LF._tail_launch(x, pt, 1e-6, W, b, LF._POOL["GeM"], 0, pooled_out=g, z_out=z)
    out = LF._tail_launch(x, pt, 1e-6, W, b, LF._POOL["GeM"], 0)
    assert torch.equal(out, out_t)
    g_ref = O.gem(x.cpu(), pt.cpu()).flatten(1)
    u = g_ref / (g_ref.norm(dim=1, keepdim=True) + 1e-6)
    z_ref = u.double() @ W.cpu().double().t() + b.cpu().double()
    np.testing.assert_allclose(g.cpu().numpy(), g_ref.numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(z.cpu().numpy(), z_ref.numpy(), rtol=0, atol=2e-5)
    zn = z / (z.norm(dim=1, keepdim=True) + 1e-6)
    np.testing.assert_allclose(out.cpu().numpy(), zn.cpu().numpy(), rtol=0, atol=1e-6)
    # z_out is the whitening projection: refused where there is none
    import ctypes as C
    from cirtorch_b200 import _lib
    lib = _lib.load()
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=DEV)
    rc = lib.cir_tail_fwd_train(x.data_ptr(), 7, 256, 12, 16, pt.data_ptr(), 0, 1e-6, 1e-6, 0, None, None, 256, out.data_ptr(), 256,
                                None, z.data_ptr(), ws.data_ptr(), ws.numel(), LF.CIR_TAIL_NO_WHITEN, None)
    assert rc != 0 and b"z_out" in lib.cir_last_error()


@pytest.mark.parametrize("p", [3.0, 2.7])
def test_launches_are_bit_identical(p):
    """Ring hand-over races (a slot refilled while a row is still being read) show up as launches that differ: 40 launches of
    the full-size batch, other data through the ring in between, pool-only and full -- all bit-identical to the first."""
    from cirtorch_b200 import functional as LF
    torch.manual_seed(21)
    x = torch.relu(torch.randn(64, 2048, 32, 32, device=DEV))
    other = torch.relu(torch.randn(64, 2048, 32, 32, device=DEV))
    head = _head(2048, p=p).to(DEV)
    pt = torch.full((1,), p, device=DEV)
    with torch.no_grad():
        ref = head(x).clone()
        ref_pool = LF.descriptor_tail(x, p=pt, pooling="GeM", pool_only=True).clone()
        for i in range(40):
            if i % 3 == 0:
                head(other)
            assert torch.equal(head(x), ref), i
            assert torch.equal(LF.descriptor_tail(x, p=pt, pooling="GeM", pool_only=True), ref_pool), i


@pytest.mark.parametrize("p", [3.0, 2.7])
def test_launch_shape_hint_never_changes_the_result(p):
    """CIR_TAIL_HINT_INTEGER_P only picks the launch shape (512 threads for cheap rows, 640 for two MUFU operations per
    element); the exponent is classified on the device either way.  Right hint, wrong hint, no hint: the same descriptors
    (the row sums are bit-equal; the final norms are summed in a thread-count-dependent order, hence 1e-6) == the oracle."""
    from cirtorch_b200 import functional as LF
    torch.manual_seed(13)
    x = (torch.relu(torch.randn(9, 256, 20, 24)) + 0.01).to(DEV)
    head = _head(256, "GeM", p=p).to(DEV)
    W, b, pt = head.whiten.weight.detach(), head.whiten.bias.detach(), head.pool.p.detach()
    ref = O.head_forward(x.cpu(), pt.cpu(), 1e-6, W.cpu(), b.cpu()).t().numpy()
    outs = []
    for flags in (0, LF.CIR_TAIL_HINT_INTEGER_P):
        o = LF._tail_launch(x, pt, 1e-6, W, b, LF._POOL["GeM"], flags, allow_hint=False)
        g = LF._tail_launch(x, pt, 1e-6, None, None, LF._POOL["GeM"], flags | LF.CIR_TAIL_POOL_ONLY, allow_hint=False)
        np.testing.assert_allclose(o.cpu().numpy(), ref, rtol=0, atol=1e-5)
        outs.append((o, g))
    assert torch.equal(outs[0][1], outs[1][1])                                # pooled values: bit-equal across launch shapes
    np.testing.assert_allclose(outs[0][0].cpu().numpy(), outs[1][0].cpu().numpy(), rtol=0, atol=1e-6)
    # the module path picks the hint itself from the frozen exponent (one read-back per tensor version)
    with torch.no_grad():
        np.testing.assert_allclose(head(x).t().cpu().numpy(), ref, rtol=0, atol=1e-5)
    assert head.pool.p._cir_integer_p == (head.pool.p._version, p == 3.0)
    with torch.no_grad():
        head.pool.p.fill_(2.5 if p == 3.0 else 4.0)            # a new version of the tensor: read back again
        head(x)
    assert head.pool.p._cir_integer_p == (head.pool.p._version, p != 3.0)


def test_inference_inside_a_cuda_graph():
    """The fused tail (a cooperative launch) can be captured and replayed; a head whose exponent has not been read back yet is
    captured without a device-to-host copy (general launch shape) and replays to the eager result."""
    torch.manual_seed(17)
    x = (torch.relu(torch.randn(6, 128, 16, 16)) + 0.01).to(DEV)
    warm = _head(128, "GeM", p=3.0).to(DEV)
    with torch.no_grad():
        warm(x)                                             # library state (attributes, workspace) exists before the capture
    head = _head(128, "GeM", p=3.0).to(DEV)                 # fresh exponent tensor: no cached hint
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            y = head(x)
    assert not hasattr(head.pool.p, "_cir_integer_p")
    graph.replay()
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = head(x)
    np.testing.assert_allclose(y.cpu().numpy(), ref.cpu().numpy(), rtol=0, atol=1e-6)
    assert head.pool.p._cir_integer_p[1] is True


def test_cpu_tensor_is_rejected():
    from cirtorch_b200._lib import CirError
    head = _head(16)
    with pytest.raises(CirError):
        head(torch.zeros(1, 16, 2, 2))


@pytest.mark.parametrize("p", [3.0, 2.7])
def test_full_size_all_images(p):
    """BASELINE.json config 2 at full size (64 x 2048 x 32 x 32, whitening 2048 -> 2048), p = 3 and a non-integer exponent:
    ALL 64 descriptors against the oracle; unit norms, batch-order equivariance, determinism."""
    torch.manual_seed(4)
    x = torch.relu(torch.randn(64, 2048, 32, 32, device=DEV))
    head = _head(2048, p=p).to(DEV)
    with torch.no_grad():
        head.whiten.bias.normal_(0, 0.02)
        a = head(x)
        b = head(x)
        perm = torch.randperm(64, device=DEV)
        c = head(x[perm].contiguous())
        a_nw = head(x, do_whitening=False)
    assert torch.equal(a, b)
    np.testing.assert_allclose(a.norm(dim=0).cpu().numpy(), 1.0, atol=1e-4)
    assert float((a[:, perm] - c).abs().max()) < 1e-6
    xc = x.cpu()
    ref = O.head_forward(xc, p, 1e-6, head.whiten.weight.detach().cpu(), head.whiten.bias.detach().cpu())
    assert _rel(a.cpu(), ref) < RTOL
    np.testing.assert_allclose(a.cpu().numpy(), ref.numpy(), rtol=RTOL, atol=1e-6)
    assert _rel(a_nw.cpu(), O.head_forward(xc, p, 1e-6, do_whitening=False)) < RTOL


@pytest.mark.parametrize("p", [2.2, 2.7, 3.3, 4.5, 1.5, 5.5])
def test_general_exponent_accuracy(p):
    """Non-integer p: half of the ex2 of x^p = ex2(p lg2 x) run as a degree-4 polynomial on the FMA pipe (tail.cu
    poly_ex2).  Pooled values within 1e-5 relative, descriptors within 1e-4, over a wide dynamic range of activations
    (exact zeros, values at the eps clamp, 1e-4 .. 60)."""
    from cirtorch_b200.modules.pools import GeM
    torch.manual_seed(int(p * 100))
    x = torch.relu(torch.randn(6, 256, 16, 16)) * torch.logspace(-4, 1.3, 256).reshape(1, 256, 1, 1)
    x[0, :8] = 0.0
    x[1, :8] = 1e-6
    x[2, 3] = 60.0
    ref_g = O.gem(x, p)
    with torch.no_grad():
        got_g = GeM(p=p).to(DEV)(x.to(DEV))
    np.testing.assert_allclose(got_g.cpu().numpy(), ref_g.numpy(), rtol=1e-5, atol=1e-30)
    head = _head(256, p=p)
    ref = O.head_forward(x, p, 1e-6, head.whiten.weight.detach(), head.whiten.bias.detach())
    with torch.no_grad():
        out = head.to(DEV)(x.to(DEV))
    assert _rel(out.cpu(), ref) < RTOL
