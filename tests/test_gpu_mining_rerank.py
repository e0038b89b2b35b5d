"""Hard-negative mining, alpha-QE / DBA and the whitening helpers vs the CPU oracle."""
import numpy as np
import pytest
import torch

from helpers import clustered_unit_rows, check_topk_against_exact
from oracle import cirtorch_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_mining_golden_fixture(golden):
    from cirtorch_b200.mining import mine_hard_negatives
    g = golden("mining")
    neg, avg = mine_hard_negatives(_dev(g["qvecs"]), _dev(g["poolvecs"]), g["clusters"].tolist(),
                                   g["query_indices"].tolist(), g["idxs2images"], int(g["neg_num"]))
    assert neg == g["negative_indices"].tolist()
    assert abs(avg - float(g["avg_dist"])) < 1e-5


def _mining_case(Q, P, D, n_clusters, n_images, seed):
    rs = np.random.RandomState(seed)
    clusters = rs.randint(0, n_clusters, size=n_images)
    idxs2images = rs.permutation(n_images)[:P]
    query_indices = rs.permutation(n_images)[:Q]
    centres = rs.randn(n_clusters, D).astype(np.float32)
    allv = centres[clusters] * 0.6 + rs.randn(n_images, D).astype(np.float32)
    allv /= np.linalg.norm(allv, axis=1, keepdims=True)
    return allv[query_indices].T.copy(), allv[idxs2images].T.copy(), clusters, query_indices, idxs2images


@pytest.mark.parametrize("Q,P,D,ncl,nimg,nnum", [
    (64, 1500, 128, 40, 4000, 5),
    (200, 3000, 64, 12, 5000, 5),      # few clusters: long walks, forces the longer-list / full-rank fallback
    (10, 200, 32, 8, 300, 3),
])
def test_mining_matches_oracle(Q, P, D, ncl, nimg, nnum):
    from cirtorch_b200.mining import mine_hard_negatives
    qv, pv, clusters, qidx, i2i = _mining_case(Q, P, D, ncl, nimg, seed=Q)
    ref_neg, ref_avg = O.mine_hard_negatives(torch.from_numpy(qv), torch.from_numpy(pv), clusters.tolist(),
                                             qidx.tolist(), i2i, nnum)
    neg, avg = mine_hard_negatives(_dev(qv), _dev(pv), clusters, qidx, i2i, nnum)
    if neg != ref_neg:
        # only fp32 near-ties may reorder the walk: the sets must still carry equal scores
        sims = pv.T.astype(np.float64) @ qv.astype(np.float64)
        pos = {int(v): j for j, v in enumerate(i2i)}
        for q, (a, b) in enumerate(zip(neg, ref_neg)):
            sa = sorted(sims[pos[x], q] for x in a)
            sb = sorted(sims[pos[x], q] for x in b)
            np.testing.assert_allclose(sa, sb, atol=2e-6)
    assert abs(avg - ref_avg) < 1e-4


def test_mining_config3_full_size():
    """BASELINE.json config 3 (2000 q x 20000 pool, nnum 5, D 2048): ALL 2,000 hard-negative sets identical to the reference
    restatement (torch.mm + torch.sort + greedy loop on the host)."""
    from cirtorch_b200.mining import mine_hard_negatives
    qv, pv, clusters, qidx, i2i = _mining_case(2000, 20000, 2048, 700, 91642, seed=11)
    neg, avg = mine_hard_negatives(_dev(qv), _dev(pv), clusters, qidx, i2i, 5)
    ref_neg, ref_avg = O.mine_hard_negatives(torch.from_numpy(qv), torch.from_numpy(pv), clusters.tolist(), qidx.tolist(), i2i, 5)
    assert neg == ref_neg
    assert abs(avg - ref_avg) < 1e-5
    for q in range(2000):                                 # invariants on every query
        cl = [clusters[x] for x in neg[q]]
        assert len(set(cl)) == 5 and clusters[qidx[q]] not in cl


def test_mining_boundary_near_tie_escalates():
    """The walk's last negative sits exactly at the end of the first candidate list (kc = 80) and a row of another cluster
    just OUTSIDE the list has an fp32 score 2e-6 higher or lower -- below what the bf16x3 scan resolves.  The margin rule
    (mining.py) must re-run those queries with a longer list; the sets must equal the reference's fp32 ranking."""
    from cirtorch_b200.mining import mine_hard_negatives
    rs = np.random.RandomState(3)
    D, nq, kc, nnum = 128, 16, 80, 5
    rows, clusters = [], []

    def planted(j, score):
        u = np.zeros(D, np.float32)
        u[64:] = rs.randn(64)
        u[64:] /= np.linalg.norm(u[64:])
        r = np.sqrt(1.0 - score ** 2) * u
        r[j] = score
        return r.astype(np.float32)

    for j in range(nq):
        for t in range(kc - 1):                           # ranks 0 .. kc-2: four clusters only
            rows.append(planted(j, 0.9 - 0.3 * t / kc))
            clusters.append(100 + 4 * j + t % 4)
        gap = (2e-6, -2e-6, 5e-7, -5e-7)[j % 4]
        rows.append(planted(j, 0.5))                      # cluster E
        clusters.append(1000 + 2 * j)
        rows.append(planted(j, np.float32(0.5) + np.float32(gap)))   # cluster F: the true winner when gap > 0
        clusters.append(1001 + 2 * j)
    for _ in range(2500):                                 # filler: score exactly 0 for every query
        rows.append(planted(127, 0.0))
        clusters.append(int(rs.randint(2000, 2100)))
    pool = np.stack(rows)
    perm = rs.permutation(len(pool))
    pool, pool_clusters = pool[perm], np.asarray(clusters)[perm]
    qv = np.eye(D, dtype=np.float32)[:nq]
    all_clusters = np.concatenate([pool_clusters, np.arange(nq) + 5000])        # queries: their own clusters
    i2i = np.arange(len(pool))
    qidx = np.arange(nq) + len(pool)
    ref_neg, _ = O.mine_hard_negatives(torch.from_numpy(qv.T.copy()), torch.from_numpy(pool.T.copy()), all_clusters.tolist(),
                                       qidx.tolist(), i2i, nnum)
    neg, _ = mine_hard_negatives(_dev(qv.T.copy()), _dev(pool.T.copy()), all_clusters, qidx, i2i, nnum)
    assert neg == ref_neg
    for j in range(nq):                                   # the fifth negative is the truly higher of E / F
        want = 1001 + 2 * j if (2e-6, -2e-6, 5e-7, -5e-7)[j % 4] > 0 else 1000 + 2 * j
        assert all_clusters[neg[j][4]] == want


def test_mining_exhausted_pool_raises():
    from cirtorch_b200.mining import mine_hard_negatives
    qv, pv, clusters, qidx, i2i = _mining_case(4, 50, 32, 3, 100, seed=5)
    with pytest.raises(IndexError):
        mine_hard_negatives(_dev(qv), _dev(pv), clusters, qidx, i2i, 5)     # only 2 foreign clusters exist


def test_alpha_qe_and_dba_match_oracle():
    from cirtorch_b200 import rerank, search as S
    db, _ = clustered_unit_rows(3000, 128, 60, 0.6, seed=8)
    q, _ = clustered_unit_rows(50, 128, 60, 0.6, seed=8)
    ref = O.alpha_qe(q.T, db.T, k=10, alpha=3.0)
    out = rerank.alpha_qe(_dev(q.T), _dev(db.T), k=10, alpha=3.0)
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=2e-5)
    # the re-ranked lists agree with the oracle's re-search
    s_ref, i_ref = O.topk(db.T, ref, 20)
    s, i = S.search_topk(out.contiguous(), _dev(db.T), 20)
    # tie-window rule on the re-ranked lists: exact scores of the oracle's expanded queries, device scores within 2e-5
    exact = ref.T.astype(np.float64) @ db.T.astype(np.float64)
    check_topk_against_exact(i.cpu().numpy().T, s.cpu().numpy().T, exact, 20, 2e-5)
    ref_d = O.dba(db.T[:, :800], k=5, alpha=3.0)
    out_d = rerank.dba(_dev(db.T[:, :800].copy()), k=5, alpha=3.0)
    np.testing.assert_allclose(out_d.cpu().numpy(), ref_d, atol=2e-5)
    # k = 0 -> identity
    idn = rerank.qe_aggregate_rows(_dev(q), _dev(db), torch.zeros((50, 1), dtype=torch.int32, device=DEV),
                                   torch.zeros((50, 1), device=DEV), 0, 3.0)
    np.testing.assert_allclose(idn.cpu().numpy(), q / (1 + 1e-6), atol=1e-6)


def test_whiten_helpers_golden(golden):
    from cirtorch_b200.utils import whiten as W
    g = golden("whiten")
    X = g["X"]
    out = W.whitenapply(X, g["m"], g["P"])
    assert isinstance(out, np.ndarray) and out.shape == g["apply"].shape
    rel = np.linalg.norm(out - g["apply"], axis=0) / np.linalg.norm(g["apply"], axis=0)
    assert rel.max() < 1e-4
    out16 = W.whitenapply(X, g["m"], g["P"], dimensions=16)
    rel = np.linalg.norm(out16 - g["apply_16"], axis=0) / np.linalg.norm(g["apply_16"], axis=0)
    assert rel.max() < 1e-4
    np.testing.assert_allclose(W.cholesky(g["S"]), g["L"], rtol=1e-9, atol=1e-12)
    assert out.dtype == g["apply"].dtype                     # fp32 X with the fp64 m, P of whitenlearn -> fp64, like numpy
    m, P = W.whitenlearn(X, g["qidxs"], g["pidxs"])
    assert P.dtype == np.float64 and m.dtype == g["m"].dtype
    np.testing.assert_allclose(m, g["m"], rtol=1e-5, atol=1e-8)
    # (1) the covariance is formed in the descriptors' precision, like whiten.py:35-37: equal to the fixture's to fp32 rounding
    _, S = W.lw_covariance(X, g["qidxs"], g["pidxs"])
    assert S.dtype == torch.float32
    assert np.abs(S.cpu().numpy() - g["S_lw"]).max() <= 2e-6 * np.abs(g["S_lw"]).max()
    # (2) everything after it, from the SAME covariance, basis-invariantly to 1e-6: P^T P (invariant to the eigenvector
    #     rotation) and |P (X - m)| (invariant to eigenvector signs)
    P2 = W.lw_from_covariance(g["S_lw"], X, g["m"]).cpu().numpy()
    ptp, ptp_ref = P2.T @ P2, g["P"].T @ g["P"]
    assert np.abs(ptp - ptp_ref).max() <= 1e-6 * np.abs(ptp_ref).max()
    ya, yb = np.abs(O.whitenapply(X, g["m"], P2)), np.abs(g["apply"])
    assert (np.linalg.norm(ya - yb, axis=0) / np.linalg.norm(yb, axis=0)).max() < 1e-6
    # (3) end to end the two fp32 covariances differ by summation order (1e-7 relative), which inv(S) amplifies by
    #     cond(S) = 3.2e3: the bound below is that amplification, not a tolerance of the factorisation
    ptp = P.T @ P
    assert np.abs(ptp - ptp_ref).max() < 2e-3 * np.abs(ptp_ref).max()
    ya = np.abs(O.whitenapply(X, m, P))
    assert (np.linalg.norm(ya - yb, axis=0) / np.linalg.norm(yb, axis=0)).max() < 2e-3
    mp, Pp = W.pcawhitenlearn(X)
    np.testing.assert_allclose(mp, g["m_pca"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(np.abs(O.whitenapply(X, mp, Pp)), np.abs(g["apply_pca"]), rtol=1e-4, atol=1e-5)


def test_whitenapply_equals_tail_with_folded_bias():
    """whitenapply(X, m, P) == the tail's Linear+L2N with W = P, b = -P m."""
    from cirtorch_b200.utils import whiten as W
    rs = np.random.RandomState(9)
    X = rs.randn(256, 500).astype(np.float32)
    X /= np.linalg.norm(X, axis=0, keepdims=True)
    m, P = O.pcawhitenlearn(X.astype(np.float64))
    ref = O.whitenapply(X.astype(np.float64), m, P)
    out = W.whitenapply(torch.from_numpy(X).to(DEV), m, P)
    assert torch.is_tensor(out) and out.is_cuda
    rel = np.linalg.norm(out.cpu().numpy() - ref, axis=0) / np.linalg.norm(ref, axis=0)
    assert rel.max() < 1e-4


def test_extract_vectors_resnet50_end_to_end():
    """Config-1 slice: torchvision ResNet50 (stock, random init) -> fused tail -> ranking, against the same
    backbone followed by the oracle head, on a handful of small synthetic images."""
    torchvision = pytest.importorskip("torchvision")
    from cirtorch_b200.extract import resnet50_gem, extract_vectors
    from cirtorch_b200 import search as S
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False          # the stock backbone must be reproducible across batch sizes
    torch.backends.cuda.matmul.allow_tf32 = False
    net = resnet50_gem().to(DEV).eval()
    imgs = [torch.randn(3, 256, 256) for _ in range(6)] + [torch.randn(3, 224, 192)]
    vecs = extract_vectors(net, imgs, 256, None, batch_size=3)
    assert vecs.shape == (2048, 7) and not vecs.is_cuda
    with torch.no_grad():
        fm = [net.body(torch.stack(imgs[0:3]).to(DEV)).cpu(), net.body(torch.stack(imgs[3:6]).to(DEV)).cpu(),
              net.body(imgs[6][None].to(DEV)).cpu()]
    ref = torch.cat([O.head_forward(f, 3.0, 1e-6, net.ret_head.whiten.weight.detach().cpu(),
                                    net.ret_head.whiten.bias.detach().cpu()) for f in fm], 1)
    rel = (vecs - ref).norm(dim=0) / ref.norm(dim=0)
    assert float(rel.max()) < 1e-4
    ms = extract_vectors(net, imgs[:2], 256, None, ms=[1, 2 ** -0.5])
    assert float(ms.norm(dim=0).max()) <= 1.0 + 1e-5          # mean of unit vectors, no re-normalisation (Q3)
    scores, ranks = S.rank(vecs.to(DEV), vecs[:, :3].to(DEV))
    s_ref, r_ref = O.rank(ref.numpy(), ref[:, :3].numpy())
    assert (ranks[0].cpu().numpy() == np.arange(3)).all()


def test_compute_map_golden_and_end_to_end(golden):
    """N1: device mAP vs the reference fixture, then ranks from the search kernel -> identical mAP to the oracle."""
    from cirtorch_b200 import evaluate as E, search as S
    g = golden("eval")
    nq = int(g["nq"])
    gnd = [{"ok": g[f"ok{i}"], "junk": g[f"junk{i}"]} for i in range(nq)]
    mp, aps, pr, prs = E.compute_map(g["ranks"], gnd, [1, 5, 10])
    np.testing.assert_allclose(mp, g["map"], rtol=1e-12)
    np.testing.assert_allclose(aps, g["aps"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(pr, g["pr"], rtol=1e-12)
    np.testing.assert_allclose(prs, g["prs"], rtol=1e-12, equal_nan=True)
    assert abs(E.compute_ap(np.array([0, 3, 4, 10]), 6) - float(g["ap_direct"])) < 1e-12
    # end to end on synthetic descriptors (rOxford shape, revisited protocol)
    db, lab = clustered_unit_rows(4993, 256, 60, 0.7, seed=21)
    q, qlab = clustered_unit_rows(70, 256, 60, 0.7, seed=21)
    rs = np.random.RandomState(3)
    gnd2 = []
    for i in range(70):
        same = np.flatnonzero(lab == qlab[i])
        rs.shuffle(same)
        a, b = len(same) // 3, 2 * len(same) // 3
        gnd2.append({"easy": same[:a], "hard": same[a:b], "junk": same[b:]})
    if len(gnd2[0]["easy"]) == 0:
        gnd2[0]["easy"] = gnd2[0]["hard"][:1]
    scores, ranks = S.rank(_dev(db.T.copy()), _dev(q.T.copy()))
    ref_scores, ref_ranks = O.rank(db.T, q.T)
    logs = []
    out = E.compute_map_and_print("roxford5k", ranks, gnd2, lambda fmt, *a: logs.append(fmt % a))
    ref = O.compute_map_revisited(ref_ranks, gnd2, (1, 5, 10))
    assert abs(out["mAP"] - ref["mAP"]) < 1e-6 and len(logs) == 4
    # top-k lists are accepted too: with k >= the last positive's rank the AP is unchanged
    mp_full, _, _, _ = E.compute_map(ranks, [{"ok": np.concatenate([g2["easy"], g2["hard"]]), "junk": g2["junk"]} for g2 in gnd2])
    mp_top, _, _, _ = E.compute_map(ranks[:4000], [{"ok": np.concatenate([g2["easy"], g2["hard"]]), "junk": g2["junk"]} for g2 in gnd2])
    assert mp_top <= mp_full + 1e-12
