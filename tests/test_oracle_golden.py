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
