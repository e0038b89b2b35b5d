"""Shared helpers of the parity tests: seeded synthetic descriptors and the tie-window rule."""
import numpy as np


def clustered_unit_rows(n, d, n_centres, sigma, seed):
    """Unit-norm rows drawn around ``n_centres`` Gaussian centres (non-degenerate top-k)."""
    rs = np.random.RandomState(seed)
    centres = rs.randn(n_centres, d).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    which = rs.randint(0, n_centres, size=n)
    x = centres[which] + sigma * rs.randn(n, d).astype(np.float32) / np.sqrt(d)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32), which


def check_topk_against_exact(idx, scores, exact_scores, k, tol, score_tol=None):
    """idx/scores [Q, k] from the device; exact_scores [Q, N] fp64 reference.

    Rule (north-star): lists identical except where the reference scores of the swapped items
    differ by <= 2*tol; returned scores within ``score_tol`` of the reference scores of the
    returned items; the k-th returned item's reference score is within 2*tol of the true k-th."""
    score_tol = tol if score_tol is None else score_tol
    Q = idx.shape[0]
    order = np.argsort(-exact_scores, axis=1, kind="stable")[:, :k]
    ref_s = np.take_along_axis(exact_scores, order, 1)
    got_s = np.take_along_axis(exact_scores, idx.astype(np.int64), 1)
    assert np.abs(scores - got_s).max() <= score_tol, "returned scores off by %g" % np.abs(scores - got_s).max()
    # positional agreement up to the tie window
    bad = idx != order
    assert np.abs(got_s - ref_s)[bad].max(initial=0.0) <= 2 * tol, \
        "rank mismatch outside the tie window: %g" % np.abs(got_s - ref_s)[bad].max(initial=0.0)
    for q in range(Q):
        assert len(set(idx[q].tolist())) == k, "duplicate index in a top-k list"
    return float(bad.mean())


def check_topk_on_device(idx, scores, q, db, k, tol, score_tol, chunk=512):
    """The same rule as check_topk_against_exact for problems too large for the host: idx / scores [Q, k], q [Q, D] and
    db [N, D] are CUDA fp32 tensors; the exact scores are a brute-force fp32 product computed on the device in chunks of
    ``chunk`` queries (checker only -- stock torch ops).  Returns the fraction of positions that differ from the
    brute-force ranking (all of them inside the tie window)."""
    import torch
    assert not torch.backends.cuda.matmul.allow_tf32
    Q = idx.shape[0]
    differing = 0
    for a in range(0, Q, chunk):
        b = min(Q, a + chunk)
        ex = q[a:b] @ db.t()                                           # [chunk, N] fp32
        ref_s, ref_i = torch.topk(ex, k, dim=1)
        got_i = idx[a:b].long()
        assert int(got_i.min()) >= 0 and int(got_i.max()) < db.shape[0]
        got_s = torch.gather(ex, 1, got_i)
        err = float((scores[a:b] - got_s).abs().max())
        assert err <= score_tol, "returned scores off by %g" % err
        bad = got_i != ref_i
        worst = float(((got_s - ref_s).abs() * bad).max())
        assert worst <= 2 * tol, "rank mismatch outside the tie window: %g" % worst
        srt = torch.sort(got_i, dim=1).values
        assert bool((srt[:, 1:] != srt[:, :-1]).all()), "duplicate index in a top-k list"
        differing += int(bad.sum())
        del ex
    return differing / float(Q * k)
