"""Search / top-k / ranking kernels vs the CPU oracle (np.dot + np.argsort restatement of
scripts/train_globalF.py:733-734).  Tolerances: bf16 scores 5e-4, bf16x3 and fp32 re-scored 2e-6
(unit-norm D <= 2048); index lists identical except inside the tie window.

Measured on B200: the tcgen05 fp32 accumulator truncates instead of rounding, so a K = 3 x 2048 bf16x3
dot product of magnitude ~0.7 carries up to ~1e-5 absolute error (384 sequential accumulations x 2^-24);
TOL_X3 states that.  Lists that must be fp32-exact go through the fp32 re-score (TOL_F32)."""
import numpy as np
import pytest
import torch

from helpers import clustered_unit_rows, check_topk_against_exact, check_topk_on_device
from oracle import cirtorch_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_BF16 = 5e-4
TOL_X3 = 2e-5
TOL_F32 = 2e-6


def _exact(q, db):
    return q.astype(np.float64) @ db.astype(np.float64).T


def _dev(a):
    return torch.from_numpy(a).to(DEV)


def test_pack_bf16_roundtrip():
    from cirtorch_b200 import search as S
    rs = np.random.RandomState(0)
    x = rs.randn(37, 100).astype(np.float32)
    p1 = S.pack_rows(_dev(x), "db", "bf16").float().cpu().numpy()
    assert p1.shape == (37, 128)
    ref = torch.from_numpy(x).bfloat16().float().numpy()
    np.testing.assert_array_equal(p1[:, :100], ref)
    assert (p1[:, 100:] == 0).all()
    pq = S.pack_rows(_dev(x), "query", "bf16x3").float().cpu().numpy()
    pd = S.pack_rows(_dev(x), "db", "bf16x3").float().cpu().numpy()
    hi, lo = pq[:, :100], pq[:, 256:356]
    np.testing.assert_array_equal(pq[:, 128:228], hi)
    np.testing.assert_array_equal(pd[:, 128:228], lo)
    np.testing.assert_array_equal(pd[:, 256:356], hi)
    assert np.abs(hi + lo - x).max() < 2e-5 * np.abs(x).max()


@pytest.mark.parametrize("Q,N,D,k", [
    (70, 4993, 2048, 100),     # rOxford5k shape (config 1)
    (1, 300, 64, 10),
    (129, 1000, 128, 5),       # two query tiles, one ragged
    (33, 257, 192, 257 if 257 <= 512 else 512),   # k == N
    (16, 50, 64, 64),          # k > N: -1 / -inf padding
    (40, 3000, 100, 512),      # D not a multiple of 64 (zero-padded operands), largest fused k
])
def test_topk_matches_oracle(Q, N, D, k):
    from cirtorch_b200 import search as S
    db, _ = clustered_unit_rows(N, D, max(4, N // 50), 0.8, seed=1)
    q, _ = clustered_unit_rows(Q, D, max(4, N // 50), 0.8, seed=1)   # same centres -> meaningful neighbours
    ex = _exact(q, db)
    kk = min(k, N)
    tol_bf16 = TOL_BF16 * (2048.0 / D) ** 0.5      # bf16 rounding noise of unit-norm rows grows as 1/sqrt(D)
    for mode, rescore, tol, stol in (("bf16", False, tol_bf16, tol_bf16), ("bf16", True, TOL_F32, TOL_F32),
                                     ("bf16x3", False, TOL_X3, TOL_X3)):
        s, i = S.search_topk_rows(_dev(q), _dev(db), k, mode=mode, rescore=rescore)
        s, i = s.cpu().numpy(), i.cpu().numpy()
        assert s.shape == (Q, k) and i.dtype == np.int32
        if k > N:
            assert (i[:, N:] == -1).all() and np.isneginf(s[:, N:]).all()
        assert (np.diff(s[:, :kk], axis=1) <= 0).all(), "scores must be non-increasing"
        if mode == "bf16" and rescore and k < N:
            # candidates come from a bf16 scan: allow a swap at the boundary inside the bf16 window
            check_topk_against_exact(i[:, :kk], s[:, :kk], ex, kk, tol_bf16, stol)
        else:
            check_topk_against_exact(i[:, :kk], s[:, :kk], ex, kk, tol, stol)


def test_golden_rank_fixture(golden):
    """The committed fixture produced by the reference statements (np.dot + np.argsort)."""
    from cirtorch_b200 import search as S
    g = golden("rank")
    V, Qv = g["database_vecs"], g["qvecs"]
    scores, ranks = S.rank(_dev(V), _dev(Qv))
    assert ranks.shape == g["ranks"].shape and ranks.dtype == torch.int64
    np.testing.assert_allclose(scores.cpu().numpy(), g["scores"], atol=TOL_X3)
    np.testing.assert_array_equal(ranks.cpu().numpy(), g["ranks"])          # no ties in this fixture
    s, i = S.search_topk(_dev(Qv), _dev(V), 10, mode="bf16x3")
    np.testing.assert_array_equal(i.cpu().numpy(), g["ranks"][:10])


def test_full_rank_large_and_ties():
    from cirtorch_b200 import search as S
    rs = np.random.RandomState(5)
    # exact ties: duplicated database rows -> (score desc, index asc) must hold
    db = rs.randn(6000, 64).astype(np.float32)
    db[3000:] = db[:3000]
    q = rs.randn(9, 64).astype(np.float32)
    sc = S.scores_dense_rows(_dev(q), _dev(db), mode="bf16x3")
    order, srt = S.argsort_rows_desc(sc, return_sorted=True)
    sc_h, order_h, srt_h = sc.cpu().numpy(), order.cpu().numpy(), srt.cpu().numpy()
    np.testing.assert_allclose(sc_h, _exact(q, db), atol=1e-4 * np.abs(_exact(q, db)).max())
    ref = np.argsort(-sc_h, axis=1, kind="stable")
    np.testing.assert_array_equal(order_h, ref)
    np.testing.assert_array_equal(srt_h, np.take_along_axis(sc_h, ref, 1))
    # a power-of-two boundary and a multi-stage global merge
    x = torch.randn(3, 20000, device=DEV)
    o = S.argsort_rows_desc(x).cpu().numpy()
    np.testing.assert_array_equal(o, np.argsort(-x.cpu().numpy(), axis=1, kind="stable"))


@pytest.mark.parametrize("Q,N", [(3, 16_385), (5, 100_003), (2, 1_000_000)])
def test_full_rank_long_rows_radix(Q, N):
    """Rows beyond 16,384 entries take the segmented radix sort (radix.cu): identical to np.argsort(-scores, kind="stable"),
    i.e. (score descending, index ascending), with heavy ties, negative scores, a strided score matrix and ragged N."""
    from cirtorch_b200 import search as S
    g = torch.Generator(device=DEV).manual_seed(N)
    big = torch.randn((Q, N + 7), device=DEV, generator=g)
    big[:, ::3] = torch.round(big[:, ::3] * 4) / 4              # a third of the entries collide on a coarse grid
    big[0, :100] = 0.0
    x = big[:, :N]                                              # row stride N + 7
    order, srt = S.argsort_rows_desc(x, return_sorted=True)
    xh = x.cpu().numpy()
    ref = np.argsort(-xh, axis=1, kind="stable")
    np.testing.assert_array_equal(order.cpu().numpy(), ref)
    np.testing.assert_array_equal(srt.cpu().numpy(), np.take_along_axis(xh, ref, 1))


def test_label_exclusion_and_tau0():
    from cirtorch_b200 import search as S
    db, lab = clustered_unit_rows(3000, 128, 30, 0.5, seed=2)
    q, qlab = clustered_unit_rows(40, 128, 30, 0.5, seed=2)
    ex = _exact(q, db)
    ex_masked = np.where(lab[None, :] == qlab[:, None], -np.inf, ex)
    qp, dbp = S.pack_rows(_dev(q), "query", "bf16x3"), S.pack_rows(_dev(db), "db", "bf16x3")
    s, i = S.search_packed(qp, dbp, 20, q_label=_dev(qlab.astype(np.int32)), db_label=_dev(lab.astype(np.int32)))
    i = i.cpu().numpy()
    assert (lab[i] != qlab[:, None]).all()
    check_topk_against_exact(i, s.cpu().numpy(), ex_masked, 20, TOL_X3)
    # a valid lower bound of the k-th score must not change the result
    kth = np.sort(ex, axis=1)[:, -20].astype(np.float32) - 1e-3
    s2, i2 = S.search_packed(qp, dbp, 20, tau0=_dev(kth))
    s3, i3 = S.search_packed(qp, dbp, 20)
    assert torch.equal(i2, i3) and torch.equal(s2, s3)


def test_merge_equals_global_topk():
    """Shard algebra on one GPU: top-k(global) == merge(top-k per shard)."""
    from cirtorch_b200 import search as S
    db, _ = clustered_unit_rows(5000, 256, 50, 0.7, seed=3)
    q, _ = clustered_unit_rows(65, 256, 50, 0.7, seed=3)
    qd, dbd = _dev(q), _dev(db)
    s_ref, i_ref = S.search_topk_rows(qd, dbd, 50, mode="bf16x3")
    parts_s, parts_i = [], []
    bounds = [0, 1200, 1201, 3700, 5000]       # ragged shards, one of a single row
    for a, b in zip(bounds[:-1], bounds[1:]):
        s, i = S.search_topk_rows(qd, dbd[a:b].contiguous(), 50, mode="bf16x3", idx_offset=a)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = S.merge_topk(torch.stack(parts_s), torch.stack(parts_i), 50)
    assert torch.equal(mi, i_ref)
    assert torch.equal(ms, s_ref)
    # the all-gathered exchange buffer [G, 2, Q, k] (scores and indices interleaved per rank) merges in place
    G = len(parts_s)
    buf = torch.empty((G, 2, 65, 50), dtype=torch.int32, device=DEV)
    for r in range(G):
        buf[r, 0] = parts_s[r].view(torch.int32)
        buf[r, 1] = parts_i[r]
    ms2, mi2 = S.merge_topk(buf[:, 0].view(torch.float32), buf[:, 1], 50)
    assert torch.equal(mi2, i_ref) and torch.equal(ms2, s_ref)


def test_fused_exchange_merge_two_ranks_on_one_gpu():
    """cir_search_topk_exchange_merge without a second GPU: two "ranks" (two shards, two streams, two exchange buffers in the
    same device memory) run concurrently; each selection block stores its list into both buffers, signals both arrival
    counters, waits for its own and merges.  The result on both ranks == the search over the whole database.  Then the
    degenerate world of one rank.  (The real thing over NVLink: tests/test_multi_gpu.py, bench.py's parity object.)"""
    import ctypes as C
    from cirtorch_b200 import _lib, search as S
    lib = _lib.load()
    db, _ = clustered_unit_rows(20_000, 256, 60, 0.7, seed=11)
    q, _ = clustered_unit_rows(70, 256, 60, 0.7, seed=12)
    Q, k, G = 70, 50, 2
    qp = S.pack_rows(_dev(q), "query", "bf16")
    dbp = S.pack_rows(_dev(db), "db", "bf16")
    s_ref, i_ref = S.search_packed(qp, dbp, k)
    bounds = [0, 9_000, 20_000]
    words = G * 2 * Q * k
    bufs = [torch.zeros(words + Q, dtype=torch.int32, device=DEV) for _ in range(G)]          # lists + arrival counters
    peers = (C.c_void_p * G)(*[b.data_ptr() for b in bufs])
    shards = [dbp[a:b].contiguous() for a, b in zip(bounds[:-1], bounds[1:])]
    streams = [torch.cuda.Stream() for _ in range(G)]
    outs = [(torch.empty((Q, k), dtype=torch.float32, device=DEV), torch.empty((Q, k), dtype=torch.int32, device=DEV)) for _ in range(G)]
    wss = []
    for r in range(G):
        need = C.c_size_t(0)
        _lib.check(lib.cir_search_workspace_bytes(Q, shards[r].shape[0], qp.shape[1], k, C.byref(need)), "ws")
        wss.append(torch.empty(need.value, dtype=torch.uint8, device=DEV))
    torch.cuda.synchronize()
    for use in (1, 2, 3):                       # the counters only grow: the target of use u is u * G
        for r in range(G):
            outs[r][1].fill_(-7)
        torch.cuda.synchronize()
        for r in range(G):
            with torch.cuda.stream(streams[r]):
                rc = lib.cir_search_topk_exchange_merge(_lib.ptr(qp), Q, _lib.ptr(shards[r]), shards[r].shape[0], qp.shape[1], k,
                                                        bounds[r], peers, G, r, use * G, _lib.ptr(outs[r][0]), _lib.ptr(outs[r][1]),
                                                        _lib.ptr(wss[r]), wss[r].numel(), 0, streams[r].cuda_stream)
                _lib.check(rc, "cir_search_topk_exchange_merge")
        torch.cuda.synchronize()
        for r in range(G):
            assert torch.equal(outs[r][1], i_ref) and torch.equal(outs[r][0], s_ref)
    assert [int(b[words:].min()) for b in bufs] == [3 * G] * G and [int(b[words:].max()) for b in bufs] == [3 * G] * G
    # one rank: its own buffer is the only peer
    solo = torch.zeros(2 * Q * k + Q, dtype=torch.int32, device=DEV)
    need = C.c_size_t(0)
    _lib.check(lib.cir_search_workspace_bytes(Q, dbp.shape[0], qp.shape[1], k, C.byref(need)), "ws")
    ws = torch.empty(need.value, dtype=torch.uint8, device=DEV)
    so, io = torch.empty((Q, k), dtype=torch.float32, device=DEV), torch.empty((Q, k), dtype=torch.int32, device=DEV)
    rc = lib.cir_search_topk_exchange_merge(_lib.ptr(qp), Q, _lib.ptr(dbp), dbp.shape[0], qp.shape[1], k, 0,
                                            (C.c_void_p * 1)(solo.data_ptr()), 1, 0, 1, _lib.ptr(so), _lib.ptr(io),
                                            _lib.ptr(ws), ws.numel(), 0, _lib.stream_of(qp))
    _lib.check(rc, "cir_search_topk_exchange_merge")
    assert torch.equal(io, i_ref) and torch.equal(so, s_ref)
    # more queries than SMs: refused (every block must be resident while it waits), nothing launched that could hang
    big = S.pack_rows(_dev(np.repeat(q, 3, 0)), "query", "bf16")
    rc = lib.cir_search_topk_exchange_merge(_lib.ptr(big), 210, _lib.ptr(dbp), dbp.shape[0], qp.shape[1], k, 0,
                                            (C.c_void_p * 1)(solo.data_ptr()), 1, 0, 2, _lib.ptr(so), _lib.ptr(io),
                                            _lib.ptr(ws), ws.numel(), 0, _lib.stream_of(qp))
    assert rc != 0 and b"fused merge" in lib.cir_last_error()


def test_many_splits_small_k_large_n():
    """More database tiles than SMs x a few: exercises multi-tile units, list compaction and the select rounds."""
    from cirtorch_b200 import search as S
    db, _ = clustered_unit_rows(200_000, 64, 500, 0.9, seed=4)
    q, _ = clustered_unit_rows(200, 64, 500, 0.9, seed=4)
    ex = _exact(q, db)
    for k in (1, 10, 100, 400):
        s, i = S.search_topk_rows(_dev(q), _dev(db), k, mode="bf16x3")
        check_topk_against_exact(i.cpu().numpy(), s.cpu().numpy(), ex, k, TOL_X3)


@pytest.mark.parametrize("k", [1, 37, 100, 512])
def test_large_query_batch_warp_select(k):
    """Q >= 4 x SM count takes the warp-per-query selection kernel (and N >= 65,536 the threshold pre-pass): same lists as
    the oracle, and merging ragged shard lists (one shorter than k: -1 padding) reproduces the global list bit for bit."""
    from cirtorch_b200 import search as S
    N, Q, D = 70_000, 700, 64
    db, _ = clustered_unit_rows(N, D, 300, 0.9, seed=21)
    q, _ = clustered_unit_rows(Q, D, 300, 0.9, seed=21)
    ex = _exact(q, db)
    qd, dbd = _dev(q), _dev(db)
    s, i = S.search_topk_rows(qd, dbd, k, mode="bf16x3")
    check_topk_against_exact(i.cpu().numpy(), s.cpu().numpy(), ex, k, TOL_X3)
    parts_s, parts_i = [], []
    bounds = [0, 30, 20_000, 20_001, 55_555, N]
    for a, b in zip(bounds[:-1], bounds[1:]):
        ps, pi = S.search_topk_rows(qd, dbd[a:b].contiguous(), k, mode="bf16x3", idx_offset=a)
        parts_s.append(ps)
        parts_i.append(pi)
    ms, mi = S.merge_topk(torch.stack(parts_s), torch.stack(parts_i), k)
    assert torch.equal(mi, i) and torch.equal(ms, s)


def test_equal_scores_flood_large_batch():
    """The flood of ties through the warp-per-query selection: lowest indices win for every one of 600 queries."""
    from cirtorch_b200 import search as S
    row = np.random.RandomState(6).randn(1, 64).astype(np.float32)
    db = np.repeat(row, 3000, axis=0)
    q = np.random.RandomState(7).randn(600, 64).astype(np.float32)
    s, i = S.search_topk_rows(_dev(q), _dev(db), 100, mode="bf16")
    np.testing.assert_array_equal(i.cpu().numpy(), np.tile(np.arange(100, dtype=np.int32), (600, 1)))


@pytest.mark.parametrize("Q,N,D,k", [(300, 20_000, 128, 80), (200, 9_000, 64, 10), (640, 40_000, 64, 100)])
def test_mid_size_database_threshold_prepass(Q, N, D, k):
    """8,192 <= N < 65,536 with several query tiles runs the group-max threshold pre-pass (2 k groups of 8 rows); with
    labels (mining) the threshold only counts rows outside the query's cluster.  Lists == oracle, with and without labels."""
    from cirtorch_b200 import search as S
    db, db_lab = clustered_unit_rows(N, D, 200, 0.9, seed=31)
    q, q_lab = clustered_unit_rows(Q, D, 200, 0.9, seed=31)
    ex = _exact(q, db)
    qp, dbp = S.pack_rows(_dev(q), "query", "bf16x3"), S.pack_rows(_dev(db), "db", "bf16x3")
    s, i = S.search_packed(qp, dbp, k)
    check_topk_against_exact(i.cpu().numpy(), s.cpu().numpy(), ex, k, TOL_X3)
    ql = torch.from_numpy(q_lab.astype(np.int32)).to(DEV)
    dl = torch.from_numpy(db_lab.astype(np.int32)).to(DEV)
    s2, i2 = S.search_packed(qp, dbp, k, q_label=ql, db_label=dl)
    ex_masked = ex.copy()
    ex_masked[q_lab[:, None] == db_lab[None, :]] = -np.inf          # the query's own cluster is not a candidate
    i2n = i2.cpu().numpy()
    assert not (db_lab[i2n] == q_lab[:, None]).any()
    check_topk_against_exact(i2n, s2.cpu().numpy(), ex_masked, k, TOL_X3)


def test_equal_scores_flood():
    """All database rows identical: every score ties; the lowest indices must win."""
    from cirtorch_b200 import search as S
    row = np.random.RandomState(6).randn(1, 64).astype(np.float32)
    db = np.repeat(row, 5000, axis=0)
    q = np.random.RandomState(7).randn(3, 64).astype(np.float32)
    s, i = S.search_topk_rows(_dev(q), _dev(db), 100, mode="bf16")
    np.testing.assert_array_equal(i.cpu().numpy(), np.tile(np.arange(100, dtype=np.int32), (3, 1)))


def test_empty_inputs():
    from cirtorch_b200 import search as S
    db = torch.randn(100, 64, device=DEV)
    s, i = S.search_topk_rows(torch.zeros(0, 64, device=DEV), db, 5)
    assert s.shape == (0, 5) and i.shape == (0, 5)
    s, i = S.search_packed(S.pack_rows(db[:3].contiguous(), "query"), torch.zeros(0, 64, device=DEV, dtype=torch.bfloat16), 4)
    assert bool((i == -1).all()) and bool(torch.isinf(s).all())
    assert S.merge_topk(torch.zeros(2, 0, 4, device=DEV), torch.zeros(2, 0, 4, device=DEV, dtype=torch.int32)).__len__() == 2


def test_ties_across_splits():
    """Equal scores in different database tiles / splits: the lower index must win everywhere."""
    from cirtorch_b200 import search as S
    rs = np.random.RandomState(11)
    base = rs.randn(50, 64).astype(np.float32)
    db = np.tile(base, (2000, 1))                        # 100,000 rows, every row repeated 2000 times
    q = base[:7] + 0.01 * rs.randn(7, 64).astype(np.float32)
    s, i = S.search_topk_rows(_dev(q), _dev(db), 300, mode="bf16", rescore=False)
    i = i.cpu().numpy()
    for r in range(7):
        # best match = base row r: its 2000 copies sit at r, r + 50, r + 100, ... -> the first 300 of them, ascending
        np.testing.assert_array_equal(i[r], r + 50 * np.arange(300))


def test_bad_arguments_raise():
    from cirtorch_b200 import search as S
    q = torch.zeros(4, 64, device=DEV)
    with pytest.raises(ValueError):
        S.search_topk_rows(q, torch.zeros(10, 64, device=DEV), 0)
    with pytest.raises(ValueError):
        S.search_topk_rows(q, torch.zeros(10, 64, device=DEV), 513)
    with pytest.raises(ValueError):
        S.search_packed(torch.zeros(4, 64, device=DEV, dtype=torch.bfloat16),
                        torch.zeros(10, 128, device=DEV, dtype=torch.bfloat16), 5)


def _million_rows(kind, N, D, gen):
    """1M unit rows on the device: "iid", or SURVEY.md 8(d) clustered rows (10k centres), optionally stored centre by centre."""
    db = torch.empty((N, D), device=DEV)
    centres = which = None
    if kind != "iid":
        centres = torch.randn((10_000, D), device=DEV, generator=gen)
        centres /= centres.norm(dim=1, keepdim=True)
        which = torch.randint(0, 10_000, (N,), device=DEV, generator=gen)
        if kind == "cluster_sorted":
            which = torch.sort(which).values
    for a in range(0, N, 100_000):
        blk = torch.randn((100_000, D), device=DEV, generator=gen)
        if centres is not None:
            blk = centres[which[a:a + 100_000]] + 0.5 * blk / D ** 0.5
        db[a:a + 100_000] = blk / blk.norm(dim=1, keepdim=True)
    return db


@pytest.mark.parametrize("Q,kind", [(70, "iid"), (1000, "iid"), (10_000, "iid"), (10_000, "clustered"), (70, "cluster_sorted"),
                                    (10_000, "cluster_sorted")])
def test_million_row_database(Q, kind):
    """BASELINE.json config 4 at full size (1M x 2048, top-100), i.i.d. / clustered / centre-sorted rows: EVERY query against
    a brute-force fp32 ranking computed on the device in query chunks (tie-window rule), plus size-independent
    properties: sortedness, fp32-exact returned scores, planted neighbours at rank 0, shard-merge == global."""
    from cirtorch_b200 import search as S
    N, D, k = 1_000_000, 2048, 100
    g = torch.Generator(device=DEV).manual_seed(0)
    db = _million_rows(kind, N, D, g)
    planted = torch.randint(0, N, (Q,), device=DEV, generator=g)
    q = db[planted] + 0.3 * torch.randn((Q, D), device=DEV, generator=g) / D ** 0.5
    q = q / q.norm(dim=1, keepdim=True)
    index = S.Index(db, mode="bf16")
    s, i = index.search_rows(q, k)                      # bf16 scan + fp32 re-score
    assert bool((i[:, 0].long() == planted).all())
    assert bool((s[:, 1:] <= s[:, :-1]).all())
    s_g, i_g = index.search_rows(q, k, rescore=False)   # raw bf16 lists
    frac = check_topk_on_device(i, s, q, db, k, tol=3e-6, score_tol=5e-6)          # both sides fp32: ties only
    frac_bf16 = check_topk_on_device(i_g, s_g, q, db, k, tol=TOL_BF16, score_tol=TOL_BF16)
    assert frac < 0.01 and frac_bf16 < 0.5
    # the threshold warm start never changes the result: spread sample (default) == first-rows sample == no pre-pass
    qp = S.pack_rows(q, "query", "bf16")
    for flags in (S.SAMPLE_FIRST_ROWS, S.NO_PREPASS):
        if flags == S.NO_PREPASS and Q > 1000:
            continue
        s_f, i_f = S.search_packed(qp, index.packed, k, flags=flags)
        assert torch.equal(i_f, i_g) and torch.equal(s_f, s_g)
    # shard-merge == global (4 shards, raw bf16 lists so both sides see identical scores)
    parts = []
    for r in range(4):
        a, b = r * N // 4, (r + 1) * N // 4
        sh = S.Index.__new__(S.Index)
        sh.mode, sh.N, sh.D, sh.row_offset, sh.packed, sh.rows32, sh.labels = "bf16", b - a, D, a, index.packed[a:b], None, None
        parts.append(sh.search_rows(q, k, rescore=False))
    ms, mi = S.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), k)
    assert torch.equal(mi, i_g) and torch.equal(ms, s_g)


def test_threshold_sample_is_order_independent():
    """Ragged, centre-sorted database (the first rows are a handful of centres): default spread sample, first-rows sample and no
    pre-pass return bit-identical lists, all equal to the oracle's."""
    from cirtorch_b200 import search as S
    db, which = clustered_unit_rows(140_123, 64, 400, 0.5, seed=21)
    db = db[np.argsort(which, kind="stable")]
    q, _ = clustered_unit_rows(333, 64, 400, 0.5, seed=21)
    qp, dbp = S.pack_rows(_dev(q), "query", "bf16"), S.pack_rows(_dev(db), "db", "bf16")
    s0, i0 = S.search_packed(qp, dbp, 100)
    for flags in (S.SAMPLE_FIRST_ROWS, S.NO_PREPASS):
        s1, i1 = S.search_packed(qp, dbp, 100, flags=flags)
        assert torch.equal(i0, i1) and torch.equal(s0, s1)
    check_topk_on_device(i0, s0, _dev(q), _dev(db), 100, tol=TOL_BF16 * (2048 / 64) ** 0.5, score_tol=TOL_BF16 * (2048 / 64) ** 0.5)


def test_index_save_load_roundtrip(tmp_path):
    from cirtorch_b200 import search as S, store
    db, _ = clustered_unit_rows(3000, 128, 30, 0.6, seed=12)
    q, _ = clustered_unit_rows(20, 128, 30, 0.6, seed=12)
    index = S.Index(_dev(db), mode="bf16", row_offset=1000)
    s0, i0 = index.search_rows(_dev(q), 10)
    store.save_index(str(tmp_path / "shard.pt"), index, extra={"name": "shard0"})
    index2 = store.load_index(str(tmp_path / "shard.pt"), device=DEV)
    s1, i1 = index2.search_rows(_dev(q), 10)
    assert torch.equal(i0, i1) and torch.equal(s0, s1) and int(i1.min()) >= 1000
