"""The CPU oracle (oracle/cirtorch_oracle.py) against fixtures produced by the imported
reference code (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import cirtorch_oracle as O


def test_tail_matches_reference(golden):
    g = golden("tail")
    for c in "abcd":
        x = torch.from_numpy(g[f"{c}_x"])
        p = float(g[f"{c}_p"])
        W, b = torch.from_numpy(g[f"{c}_W"]), torch.from_numpy(g[f"{c}_b"])
        np.testing.assert_allclose(O.gem(x, p).numpy(), g[f"{c}_gem"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(O.l2n(O.gem(x, p)).numpy(), g[f"{c}_l2n"], rtol=1e-5, atol=1e-7)
        np.testing.assert_array_equal(O.mac(x).numpy(), g[f"{c}_mac"])
        np.testing.assert_allclose(O.spoc(x).numpy(), g[f"{c}_spoc"], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(O.head_forward(x, p, 1e-6, W, b).numpy(), g[f"{c}_head"],
                                   rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(O.head_forward(x, p, 1e-6, W, b, do_whitening=False).numpy(),
                                   g[f"{c}_head_nowhiten"], rtol=1e-5, atol=1e-7)


def test_regional_pooling_matches_reference(golden):
    """Rpool / RMAC restatements (and the region grid) against the imported reference (pools.py:57-197)."""
    g = golden("regional")
    for name in ("sq", "wide", "tall", "tiny"):
        x = torch.from_numpy(g[f"{name}_x"])
        W, b = torch.from_numpy(g[f"{name}_W"]), torch.from_numpy(g[f"{name}_b"])
        for L in (1, 2, 3):
            np.testing.assert_allclose(O.rpool_forward(x, lambda t: O.gem(t, 3.0), L=L).numpy(), g[f"{name}_gem3_L{L}"],
                                       rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(O.rpool_forward(x, lambda t: O.gem(t, 2.5), W, b, L=L).numpy(),
                                       g[f"{name}_gem25_white_L{L}"], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(O.rpool_forward(x, O.mac, L=L).numpy(), g[f"{name}_mac_L{L}"], rtol=1e-6, atol=1e-8)
            np.testing.assert_allclose(O.rpool_forward(x, O.spoc, L=L, aggregate=False).numpy(),
                                       g[f"{name}_spoc_regions_L{L}"], rtol=1e-5, atol=1e-8)
            want = g[f"{name}_rmac_L{L}"]
            if want.size == 0:            # the reference raises NameError (cenW never assigned)
                try:
                    O.rmac_forward(x, L=L)
                    raise AssertionError("expected NameError")
                except NameError:
                    pass
            else:
                np.testing.assert_allclose(O.rmac_forward(x, L=L).numpy(), want, rtol=1e-6, atol=1e-8)
    grid = g["grid"]
    for (H, W_, L) in sorted({tuple(r[:3]) for r in grid.tolist()}):
        want = [tuple(r[3:]) for r in grid.tolist() if tuple(r[:3]) == (H, W_, L)]
        got = [(0, 0, H, W_)] + [(i, j, wl, wl) for (wl, tops, lefts) in O.rmac_region_grid(H, W_, L) if wl > 0
                                 for i in tops for j in lefts]
        assert got == want, (H, W_, L)


def test_whiten_matches_reference(golden):
    g = golden("whiten")
    X = g["X"]
    np.testing.assert_allclose(O.whitenapply(X, g["m"], g["P"]), g["apply"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(O.whitenapply(X, g["m"], g["P"], dimensions=16), g["apply_16"],
                               rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(O.cholesky(g["S"]), g["L"], rtol=1e-12)
    m, P = O.whitenlearn(X, g["qidxs"], g["pidxs"])
    np.testing.assert_allclose(m, g["m"], rtol=1e-12)
    # eigenvectors are defined up to sign: compare P^T P and the applied result up to row sign
    np.testing.assert_allclose(P.T @ P, g["P"].T @ g["P"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(np.abs(O.whitenapply(X, m, P)), np.abs(g["apply"]), rtol=1e-6, atol=1e-7)
    mp, Pp = O.pcawhitenlearn(X)
    np.testing.assert_allclose(mp, g["m_pca"], rtol=1e-6)
    np.testing.assert_allclose(np.abs(O.whitenapply(X, mp, Pp)), np.abs(g["apply_pca"]), rtol=1e-5, atol=1e-6)


def test_rank_matches_reference(golden):
    g = golden("rank")
    scores, ranks = O.rank(g["database_vecs"], g["qvecs"])
    np.testing.assert_array_equal(scores, g["scores"])
    np.testing.assert_array_equal(ranks, g["ranks"])
    s, idx = O.topk(g["database_vecs"], g["qvecs"], 10)
    np.testing.assert_array_equal(idx, g["ranks"][:10])     # no ties in this fixture
    np.testing.assert_allclose(s, np.take_along_axis(g["scores"], g["ranks"][:10], 0), atol=1e-6)


def test_mining_matches_reference(golden):
    g = golden("mining")
    neg, avg = O.mine_hard_negatives(torch.from_numpy(g["qvecs"]), torch.from_numpy(g["poolvecs"]),
                                     g["clusters"].tolist(), g["query_indices"].tolist(),
                                     g["idxs2images"], int(g["neg_num"]))
    assert neg == g["negative_indices"].tolist()
    assert abs(avg - float(g["avg_dist"])) < 1e-5


def test_eval_matches_reference(golden):
    g = golden("eval")
    gnd = [{"ok": g[f"ok{i}"], "junk": g[f"junk{i}"]} for i in range(int(g["nq"]))]
    mp, aps, pr, prs = O.compute_map(g["ranks"], gnd, [1, 5, 10])
    np.testing.assert_allclose(mp, g["map"], rtol=1e-12)
    np.testing.assert_allclose(aps, g["aps"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(pr, g["pr"], rtol=1e-12)
    np.testing.assert_allclose(prs, g["prs"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(O.compute_ap(np.array([0, 3, 4, 10]), 6), g["ap_direct"], rtol=1e-12)


def test_multiscale_matches_reference(golden):
    g = golden("multiscale")
    out = O.multiscale_mean([torch.from_numpy(p) for p in g["preds"]])
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-6, atol=1e-8)


def test_alpha_qe_two_restatements_agree():
    # alpha-QE / DBA are absent from the reference (parity unpinned): cross-check two
    # independent restatements and the defining properties instead.
    rs = np.random.RandomState(0)
    V = rs.randn(16, 200); V /= np.linalg.norm(V, axis=0, keepdims=True)
    Q = rs.randn(16, 7); Q /= np.linalg.norm(Q, axis=0, keepdims=True)
    a = O.alpha_qe(Q, V, k=10, alpha=3.0)
    b = O.alpha_qe_torch(torch.from_numpy(Q), torch.from_numpy(V), 10, 3.0).numpy()
    np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(a, axis=0), 1.0, atol=1e-5)
    # k = 0 -> identity (up to the eps in the norm)
    np.testing.assert_allclose(O.alpha_qe(Q, V, k=0), Q / (1 + 1e-6), rtol=1e-9)
    d = O.dba(V, k=5, alpha=3.0)
    np.testing.assert_allclose(np.linalg.norm(d, axis=0), 1.0, atol=1e-5)


def test_losses_match_reference(golden):
    """contrastive_loss / triplet_loss restatements (values and autograd gradients) vs losses.py:7-46."""
    g = golden("losses")
    for name in "abc":
        label, msk = torch.from_numpy(g[f"{name}_label"]), torch.from_numpy(g[f"{name}_msk"])
        for margin in (0.7, 0.1):
            for kind, fn in (("contrastive", lambda x: O.contrastive_loss(x, label, margin, 1e-6)),
                             ("triplet", lambda x: O.triplet_loss(x, label, msk, margin))):
                x = torch.from_numpy(g[f"{name}_x"]).requires_grad_(True)
                y = fn(x)
                np.testing.assert_allclose(y.detach().numpy(), g[f"{name}_{kind}_{margin}"], rtol=1e-5, atol=1e-6)
                if y.requires_grad:
                    (gx,) = torch.autograd.grad(y, x)
                    np.testing.assert_allclose(gx.numpy(), g[f"{name}_{kind}_{margin}_grad"], rtol=1e-4, atol=1e-6)


def _net_fixture_body(g):
    conv = torch.nn.Conv2d(3, 32, 3, stride=2, padding=1)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(g["conv_w"]))
        conv.bias.copy_(torch.from_numpy(g["conv_b"]))
    return conv, (lambda img: {"mod5": torch.relu(conv(img))})


def test_net_forward_matches_reference(golden):
    """Padded ragged batch through body + head (GF_net.py / GF_algo.py): single- and multi-scale descriptors."""
    g = golden("net")
    _, body = _net_fixture_body(g)
    imgs = [torch.from_numpy(g[f"img{i}"]) for i in range(4)]
    W, b = torch.from_numpy(g["W"]), torch.from_numpy(g["b"])
    with torch.no_grad():
        np.testing.assert_allclose(O.net_forward(body, imgs, float(g["p"]), 1e-6, W, b).numpy(), g["pred"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(O.net_forward(body, imgs, float(g["p"]), 1e-6, W, b, scales=(1, 2 ** -0.5, 0.5)).numpy(),
                                   g["pred_ms"], rtol=1e-5, atol=1e-7)
        timgs = [torch.from_numpy(g[f"timg{i}"]) for i in range(8)]
        pred = O.net_forward(body, timgs, float(g["p"]), 1e-6, W, b)
        np.testing.assert_allclose(pred.numpy(), g["train_pred"], rtol=1e-5, atol=1e-7)
        label = torch.tensor([-1., 1., 0., 0.] * 2)
        loss = O.triplet_loss(pred, label, torch.arange(2).repeat_interleave(4), 0.5)
        np.testing.assert_allclose(loss.numpy(), g["train_loss"], rtol=1e-5)


def test_reference_net_takes_the_drop_in_algo():
    """The reference's OWN ImageRetrievalNet (cirtorch/models/GF_net.py) built around this package's globalFeatureAlgo /
    globalFeatureLoss: same call contract, same result as with the reference's algo.  Needs /root/reference (build
    container only); the head here is a CPU stand-in so that no kernel is involved."""
    import os
    import sys
    import types
    import pytest
    if not os.path.isdir("/root/reference/cirtorch"):
        pytest.skip("the reference tree is not present on this machine")
    sys.path.insert(0, "/root/reference")
    if "inplace_abn" not in sys.modules:
        stub = types.ModuleType("inplace_abn")
        stub.ABN = stub.InPlaceABN = stub.InPlaceABNSync = type("ABN", (torch.nn.Module,), {})
        stub.active_group = stub.set_active_group = lambda *a, **k: None
        sys.modules["inplace_abn"] = stub
    from cirtorch.models.GF_net import ImageRetrievalNet as RefNet
    from cirtorch.algos.GF_algo import globalFeatureAlgo as RefAlgo, globalFeatureLoss as RefLoss
    from cirtorch.utils.parallel import PackedSequence as RefSeq
    from cirtorch_b200.algos.GF_algo import globalFeatureAlgo, globalFeatureLoss

    class Head(torch.nn.Module):
        def forward(self, x):
            return O.head_forward(x, 3.0, 1e-6, do_whitening=False)

    class Body(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 16, 3, padding=1)

        def forward(self, img):
            return {"mod5": torch.relu(self.conv(img))}

    torch.manual_seed(0)
    body, head = Body(), Head()
    imgs = [torch.randn(3, 20, 16), torch.randn(3, 18, 24)]
    mine = RefNet(body, globalFeatureAlgo(globalFeatureLoss("triplet", 0.5), min_level=2, fpn_levels=1), head, augment=None).eval()
    ref = RefNet(body, RefAlgo(RefLoss("triplet", 0.5), min_level=2, fpn_levels=1), head, augment=None).eval()
    with torch.no_grad():
        for scales in ([1], [1, 0.5]):
            a = mine(img=RefSeq(imgs), scales=scales, do_prediction=True)[1]["ret_pred"]
            b = ref(img=RefSeq(imgs), scales=scales, do_prediction=True)[1]["ret_pred"]
            assert torch.equal(a, b)
    # FPN-style list input and the error for anything else (GF_algo.py:51-59)
    algo = globalFeatureAlgo(None, min_level=1, fpn_levels=2)
    assert algo._get_level(["l0", "l1", "l2", "l3"]) == "l1"
    try:
        algo._get_level(torch.zeros(1))
        raise AssertionError("expected NameError")
    except NameError:
        pass
