"""Properties of the CPU oracle itself (no GPU, no fixtures): limits and invariants that follow from the reference's
formulas, so that a slip in the restatement shows up even where no golden vector covers it.  Small sizes, seconds."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import cirtorch_oracle as O

SHAPES = st.tuples(st.integers(1, 4), st.integers(1, 9), st.integers(1, 7), st.integers(1, 7))


def _map(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(shape, generator=g)) + 1e-3


@settings(max_examples=25, deadline=None, derandomize=True)
@given(SHAPES, st.integers(0, 10_000))
def test_gem_limits(shape, seed):
    """pools.py:30-38: GeM with p = 1 is the mean of the clamped map (SPoC on a positive map), grows with p and stays below
    the maximum (MAC); a constant plane pools to its value for every p."""
    x = _map(shape, seed)
    g1 = O.gem(x, 1.0).flatten(1)
    np.testing.assert_allclose(g1.numpy(), O.spoc(x).flatten(1).numpy(), rtol=1e-5, atol=1e-7)
    g3, g8 = O.gem(x, 3.0).flatten(1), O.gem(x, 8.0).flatten(1)
    mx = O.mac(x).flatten(1)
    assert bool((g1 <= g3 * (1 + 1e-5)).all()) and bool((g3 <= g8 * (1 + 1e-5)).all()) and bool((g8 <= mx * (1 + 1e-5)).all())
    const = torch.full(shape, 0.37)
    for p in (1.0, 2.5, 3.0, 6.0):
        np.testing.assert_allclose(O.gem(const, p).numpy(), 0.37, rtol=1e-5)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(SHAPES, st.floats(1.5, 5.0), st.integers(0, 10_000))
def test_head_is_scale_invariant_and_unit_norm(shape, p, seed):
    """global_head.py:52-67: the descriptor is a unit vector; GeM is positively homogeneous and the first L2N removes the
    scale (up to its eps), so a rescaled map gives the same descriptor; images of a batch do not interact."""
    x = _map(shape, seed)
    C = shape[1]
    g = torch.Generator().manual_seed(seed + 1)
    W = torch.randn((C + 2, C), generator=g)
    b = torch.randn((C + 2,), generator=g) * 0.1
    d = O.head_forward(x, p, 1e-6, W, b)                       # D x N
    assert d.shape == (C + 2, shape[0])
    np.testing.assert_allclose(d.norm(dim=0).numpy(), 1.0, atol=2e-5)
    d2 = O.head_forward(x * 7.5, p, 1e-6, W, b)
    np.testing.assert_allclose(d2.numpy(), d.numpy(), atol=5e-4)
    alone = torch.cat([O.head_forward(x[i:i + 1], p, 1e-6, W, b) for i in range(shape[0])], 1)
    np.testing.assert_allclose(alone.numpy(), d.numpy(), atol=1e-6)
    nw = O.head_forward(x, p, 1e-6, None, None, do_whitening=False)
    np.testing.assert_allclose(nw.numpy(), O.l2n(O.gem(x, p)).flatten(1).t().numpy(), atol=1e-7)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.integers(1, 40), st.integers(1, 6), st.integers(2, 12), st.integers(0, 10_000))
def test_rank_is_a_sorted_permutation_and_topk_is_its_prefix(N, Q, D, seed):
    """train_globalF.py:733-734: every column of `ranks` is a permutation of the database sorted by descending score; the
    top-k helper is the prefix of the stable ranking (ties keep ascending index)."""
    rng = np.random.default_rng(seed)
    db = rng.standard_normal((D, N)).astype(np.float32)
    db[:, N // 2:] = db[:, : N - N // 2]                       # duplicated columns: exact ties
    q = rng.standard_normal((D, Q)).astype(np.float32)
    scores, ranks = O.rank(db, q)
    assert scores.shape == (N, Q) and ranks.shape == (N, Q)
    for j in range(Q):
        assert sorted(ranks[:, j].tolist()) == list(range(N))
        s = scores[ranks[:, j], j]
        assert bool((s[:-1] >= s[1:]).all())
    k = min(N, 7)
    ts, ti = O.topk(db, q, k)
    s64 = db.T.astype(np.float64) @ q.astype(np.float64)
    for j in range(Q):
        ref = sorted(range(N), key=lambda i: (-s64[i, j], i))[:k]
        assert ti[:, j].tolist() == ref
        np.testing.assert_allclose(ts[:, j], s64[ref, j], rtol=0, atol=0)


@settings(max_examples=20, deadline=None, derandomize=True)
@given(st.integers(3, 9), st.integers(1, 5), st.integers(1, 3), st.integers(0, 10_000))
def test_mining_invariants(n_clusters, per_cluster, neg_num, seed):
    """tuples_dataset.py:317-345: the negatives of a query come from `neg_num` DIFFERENT clusters, none of them the query's own,
    each is the most similar pool image of its cluster among those not yet excluded, and they come out in descending
    similarity."""
    rng = np.random.default_rng(seed)
    D = 8
    n_img = n_clusters * per_cluster
    clusters = [i // per_cluster for i in range(n_img)]
    vecs = torch.from_numpy(rng.standard_normal((D, n_img)).astype(np.float32))
    vecs = vecs / vecs.norm(dim=0, keepdim=True)
    qidx = [0, n_img - 1]
    pool = list(range(n_img))
    neg_num = min(neg_num, n_clusters - 1)
    negs, avg = O.mine_hard_negatives(vecs[:, qidx], vecs, clusters, qidx, pool, neg_num)
    assert len(negs) == 2 and avg >= 0.0
    sims = (vecs.t() @ vecs[:, qidx]).numpy()
    for j, q in enumerate(qidx):
        got = negs[j]
        assert len(got) == neg_num
        cl = [clusters[i] for i in got]
        assert len(set(cl)) == neg_num and clusters[q] not in cl
        s = [sims[i, j] for i in got]
        assert all(a >= b for a, b in zip(s, s[1:]))
        banned = {clusters[q]}
        for i in got:                                           # greedy: best image of any cluster not yet banned
            best = max((m for m in range(n_img) if clusters[m] not in banned), key=lambda m: (sims[m, j], -m))
            assert sims[i, j] == sims[best, j]
            banned.add(clusters[i])


@settings(max_examples=20, deadline=None, derandomize=True)
@given(st.integers(2, 30), st.integers(1, 5), st.integers(0, 10_000))
def test_average_precision_bounds(N, n_pos, seed):
    """ParisOxfordEval.py:4-38: AP is 1 when the positives lead the list, does not grow when one of them is pushed back, and for
    a single positive at 0-based rank r it is the trapezoid between the precision before it (1 if r == 0 else 0) and at it
    (1 / (r + 1))."""
    n_pos = min(n_pos, N)
    assert abs(O.compute_ap(np.arange(n_pos), n_pos) - 1.0) < 1e-12
    rng = np.random.default_rng(seed)
    ranks = np.sort(rng.choice(N, size=n_pos, replace=False))
    ap = O.compute_ap(ranks, n_pos)
    assert 0.0 < ap <= 1.0 + 1e-12
    worse = ranks.copy()
    worse[-1] += 5
    assert O.compute_ap(worse, n_pos) <= ap + 1e-12
    r = int(ranks[0])
    single = O.compute_ap(np.array([r]), 1)
    p0 = 1.0 if r == 0 else 0.0
    assert abs(single - (p0 + 1.0 / (r + 1)) / 2.0) < 1e-12
