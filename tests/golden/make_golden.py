"""Generate tests/golden/*.npz by running the IMPORTED reference code on seeded inputs.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):      python tests/golden/make_golden.py

The reference ships no golden vectors; these fixtures are the pin
for oracle/cirtorch_oracle.py and, through it, for the CUDA path.  Sizes are small so the
fixtures stay a few hundred KB.  Where the reference inlines a step in a driver script
(ranking, mining) the generator executes the same statements, quoted with file:line.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    # inplace_abn (3rd-party CUDA ext, absent here) is only used for an isinstance check
    stub = types.ModuleType("inplace_abn")

    class _ABN(torch.nn.Module):
        pass

    stub.ABN = stub.InPlaceABN = stub.InPlaceABNSync = _ABN
    stub.active_group = stub.set_active_group = lambda *a, **k: None
    sys.modules["inplace_abn"] = stub


def tail_fixtures():
    from cirtorch.modules.pools import GeM, MAC, SPoC
    from cirtorch.modules.normalizations import L2N
    from cirtorch.modules.heads.global_head import globalHead

    g = torch.Generator().manual_seed(1234)
    cases = {}
    # (N, C, H, W, p)
    for name, (n, c, h, w, p) in {
        "a": (3, 64, 8, 8, 3.0),
        "b": (2, 48, 7, 5, 2.7),       # HW not a multiple of 4 -> scalar-load path
        "c": (5, 128, 4, 12, 1.0),
        "d": (1, 32, 16, 16, 4.5),
    }.items():
        x = torch.relu(torch.randn(n, c, h, w, generator=g))
        x[0, 0] = 0.0                  # an all-zero map row: pure eps clamp
        head = globalHead(pooling={"name": "GeM", "params": {"p": p, "eps": 1e-6}},
                          normal={"name": "L2N", "params": {}}, dim=c)
        torch.manual_seed(7)
        head.reset_parameters()
        with torch.no_grad():
            head.whiten.bias.copy_(0.05 * torch.randn(c, generator=g))
            cases[f"{name}_x"] = x.numpy()
            cases[f"{name}_p"] = np.float32(p)
            cases[f"{name}_W"] = head.whiten.weight.numpy().copy()
            cases[f"{name}_b"] = head.whiten.bias.numpy().copy()
            cases[f"{name}_gem"] = GeM(p=p)(x).numpy()
            cases[f"{name}_l2n"] = L2N()(GeM(p=p)(x)).numpy()
            cases[f"{name}_mac"] = MAC()(x).numpy()
            cases[f"{name}_spoc"] = SPoC()(x).numpy()
            cases[f"{name}_head"] = head(x).contiguous().numpy()
            cases[f"{name}_head_nowhiten"] = head(x, do_whitening=False).contiguous().numpy()
    np.savez_compressed(os.path.join(OUT, "tail.npz"), **cases)


def whiten_fixtures():
    from cirtorch.utils.whiten import whitenapply, whitenlearn, pcawhitenlearn, cholesky

    rs = np.random.RandomState(5)
    D, N = 24, 400
    basis = rs.randn(D, D)
    X = basis @ rs.randn(D, N) * np.linspace(2.0, 0.2, D)[:, None]
    X = (X / np.linalg.norm(X, axis=0, keepdims=True)).astype(np.float32)
    qidxs = rs.randint(0, N, size=150)
    pidxs = rs.randint(0, N, size=150)
    m, P = whitenlearn(X, qidxs, pidxs)
    mp, Pp = pcawhitenlearn(X)
    S = np.cov(rs.randn(6, 40))
    df = X[:, qidxs] - X[:, pidxs]                     # whiten.py:36-37, the fp32 covariance whitenlearn factorises
    S_lw = np.dot(df, df.T) / df.shape[1]
    out = dict(X=X, qidxs=qidxs, pidxs=pidxs, m=m, P=P, m_pca=mp, P_pca=Pp, S_lw=S_lw,
               apply=whitenapply(X, m, P), apply_16=whitenapply(X, m, P, dimensions=16),
               apply_pca=whitenapply(X, mp, Pp), S=S, L=cholesky(S))
    np.savez_compressed(os.path.join(OUT, "whiten.npz"), **out)


def rank_fixtures():
    rs = np.random.RandomState(11)
    D, N, Q = 32, 300, 9
    centres = rs.randn(20, D)
    V = centres[rs.randint(0, 20, N)] + 0.3 * rs.randn(N, D)
    Qv = centres[rs.randint(0, 20, Q)] + 0.3 * rs.randn(Q, D)
    database_vecs = (V / np.linalg.norm(V, axis=1, keepdims=True)).T.astype(np.float32).copy()
    qvecs = (Qv / np.linalg.norm(Qv, axis=1, keepdims=True)).T.astype(np.float32).copy()
    # scripts/train_globalF.py:733-734
    scores = np.dot(database_vecs.T, qvecs)
    ranks = np.argsort(-scores, axis=0)
    np.savez_compressed(os.path.join(OUT, "rank.npz"), database_vecs=database_vecs,
                        qvecs=qvecs, scores=scores, ranks=ranks)


def mining_fixtures():
    g = torch.Generator().manual_seed(21)
    D, Q, Pn, n_img, n_clu, neg_num = 32, 12, 200, 500, 15, 5
    clusters = torch.randint(0, n_clu, (n_img,), generator=g).tolist()
    idxs2images = torch.randperm(n_img, generator=g)[:Pn]
    query_indices = torch.randperm(n_img, generator=g)[:Q].tolist()
    centres = torch.randn(n_clu, D, generator=g)

    def vec(i):
        v = centres[clusters[i]] + 0.5 * torch.randn(D, generator=g)
        return v / v.norm()

    qvecs = torch.stack([vec(i) for i in query_indices], dim=1)
    poolvecs = torch.stack([vec(int(i)) for i in idxs2images], dim=1)

    # cirtorch/datasets/globalFeatures/tuples_dataset.py:317-345, executed on CPU tensors
    scores = torch.mm(poolvecs.t(), qvecs)
    scores, scores_indices = torch.sort(scores, dim=0, descending=True)
    average_negative_distance = torch.tensor(0).float()
    negative_distance = torch.tensor(0).float()
    negative_indices = []
    for q in range(len(query_indices)):
        qcluster = clusters[query_indices[q]]
        clus = [qcluster]
        nidxs = []
        r = 0
        while len(nidxs) < neg_num:
            potential = idxs2images[scores_indices[r, q]]
            if not clusters[potential] in clus:
                nidxs.append(potential)
                clus.append(clusters[potential])
                average_negative_distance += torch.pow(
                    qvecs[:, q] - poolvecs[:, scores_indices[r, q]] + 1e-6, 2).sum(dim=0).sqrt()
                negative_distance += 1
            r += 1
        negative_indices.append([int(i) for i in nidxs])
    np.savez_compressed(os.path.join(OUT, "mining.npz"), qvecs=qvecs.numpy(),
                        poolvecs=poolvecs.numpy(), clusters=np.array(clusters),
                        idxs2images=idxs2images.numpy(), query_indices=np.array(query_indices),
                        neg_num=neg_num, negative_indices=np.array(negative_indices),
                        avg_dist=float(average_negative_distance / negative_distance))


def eval_fixtures():
    from cirtorch.utils.evaluation.ParisOxfordEval import compute_ap, compute_map

    rs = np.random.RandomState(3)
    N, Q = 120, 6
    ranks = np.stack([rs.permutation(N) for _ in range(Q)], axis=1)
    gnd, flat = [], {}
    for i in range(Q):
        perm = rs.permutation(N)
        easy, hard, junk = perm[:5], perm[5:9], perm[9:15]
        if i == 4:
            easy = easy[:0]
            hard = hard[:0]      # query without positives -> skipped
        gnd.append({"ok": np.concatenate([easy, hard]), "junk": junk})
        flat[f"ok{i}"] = gnd[-1]["ok"]
        flat[f"junk{i}"] = junk
    kappas = [1, 5, 10]
    mp, aps, pr, prs = compute_map(ranks, gnd, kappas)
    ap_direct = compute_ap(np.array([0, 3, 4, 10]), 6)
    np.savez_compressed(os.path.join(OUT, "eval.npz"), ranks=ranks, map=mp, aps=aps, pr=pr,
                        prs=prs, ap_direct=ap_direct, nq=Q, **flat)


def multiscale_fixture():
    # cirtorch/models/GF_net.py:84-85 on synthetic per-scale predictions
    g = torch.Generator().manual_seed(9)
    preds = [torch.nn.functional.normalize(torch.randn(16, 5, generator=g), dim=0) for _ in range(3)]
    pred = torch.cat([p.unsqueeze(0) for p in preds], dim=0).permute(1, 2, 0)
    pred = torch.nn.functional.avg_pool1d(pred, kernel_size=3).squeeze(-1)
    np.savez_compressed(os.path.join(OUT, "multiscale.npz"),
                        preds=torch.stack(preds).numpy(), out=pred.numpy())


def regional_fixtures():
    """Rpool / RMAC (cirtorch/modules/pools.py:57-197), run from the imported reference."""
    from cirtorch.modules.pools import GeM, MAC, SPoC, RMAC, Rpool

    g = torch.Generator().manual_seed(4321)
    cases = {}
    shapes = {"sq": (2, 16, 12, 12), "wide": (3, 16, 9, 14), "tall": (2, 8, 17, 10), "tiny": (1, 8, 5, 5)}
    with torch.no_grad():
        for name, shp in shapes.items():
            x = torch.relu(torch.randn(*shp, generator=g))
            x[0, 0] = 0.0
            cases[f"{name}_x"] = x.numpy()
            C = shp[1]
            lin = torch.nn.Linear(C, C)          # Rpool.forward views the result back as C-dimensional (:190)
            lin.weight.copy_(0.3 * torch.randn(C, C, generator=g))
            lin.bias.copy_(0.05 * torch.randn(C, generator=g))
            cases[f"{name}_W"] = lin.weight.numpy().copy()
            cases[f"{name}_b"] = lin.bias.numpy().copy()
            for L in (1, 2, 3):
                cases[f"{name}_gem3_L{L}"] = Rpool(GeM(p=3), L=L)(x).numpy()
                cases[f"{name}_gem25_white_L{L}"] = Rpool(GeM(p=2.5), whiten=lin, L=L)(x).numpy()
                cases[f"{name}_mac_L{L}"] = Rpool(MAC(), L=L)(x).numpy()
                cases[f"{name}_spoc_regions_L{L}"] = Rpool(SPoC(), L=L)(x, aggregate=False).numpy()
                try:
                    cases[f"{name}_rmac_L{L}"] = RMAC(L=L)(x).numpy()
                except NameError:      # pools.py:94-98: cenW never assigned when L + Wd == 1
                    cases[f"{name}_rmac_L{L}"] = np.zeros(0, dtype=np.float32)
    # the region grid itself, recovered from the reference by pooling a position-coded map
    grid = []
    for (H, W, L) in [(12, 12, 3), (9, 14, 3), (17, 10, 2), (11, 47, 3), (47, 11, 4), (32, 32, 3), (24, 32, 3), (5, 5, 3), (1, 7, 2)]:
        xpos = torch.arange(H * W, dtype=torch.float32).view(1, 1, H, W)
        rec = []

        def probe(t, rec=rec, W=W):
            v = int(t[0, 0, 0, 0].item())
            rec.append((v // W, v % W, t.shape[2], t.shape[3]))
            return torch.zeros(1, 1, 1, 1)

        Rpool(probe, L=L).roipool(xpos, probe, L=L)
        for r in rec:
            grid.append((H, W, L) + r)
    cases["grid"] = np.array(grid, dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "regional.npz"), **cases)


def losses_fixtures():
    """contrastive_loss / triplet_loss (cirtorch/modules/losses.py:7-46) on tuple descriptors, with autograd gradients."""
    from cirtorch.modules.losses import contrastive_loss, triplet_loss
    g = torch.Generator().manual_seed(77)
    out = {}
    for name, (D, nt, nneg) in {"a": (32, 3, 2), "b": (128, 5, 5), "c": (8, 1, 1)}.items():
        S = 2 + nneg
        x = torch.randn(D, nt * S, generator=g)
        x = (x / x.norm(dim=0, keepdim=True)).requires_grad_(True)
        label = torch.tensor(([-1, 1] + [0] * nneg) * nt, dtype=torch.float32)
        msk = torch.arange(nt).repeat_interleave(S)
        out[f"{name}_x"] = x.detach().numpy().copy()
        out[f"{name}_label"] = label.numpy()
        out[f"{name}_msk"] = msk.numpy()
        for margin in (0.7, 0.1):
            y = contrastive_loss(x, label=label, margin=margin, eps=1e-6)
            (gx,) = torch.autograd.grad(y, x)
            out[f"{name}_contrastive_{margin}"] = y.detach().numpy()
            out[f"{name}_contrastive_{margin}_grad"] = gx.numpy()
            y = triplet_loss(x, label=label, label_msk=msk, margin=margin)
            (gx,) = torch.autograd.grad(y, x)
            out[f"{name}_triplet_{margin}"] = y.detach().numpy()
            out[f"{name}_triplet_{margin}_grad"] = gx.numpy()
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)


def net_fixtures():
    """The reference's own ImageRetrievalNet (cirtorch/models/GF_net.py:10-126, augment=None) around its globalFeatureAlgo
    (cirtorch/algos/GF_algo.py) and globalHead, on a ragged PackedSequence: single- and multi-scale inference and a
    training step with the triplet loss.  The body is a one-layer stand-in backbone returning {"mod5": map}."""
    from cirtorch.models.GF_net import ImageRetrievalNet
    from cirtorch.algos.GF_algo import globalFeatureAlgo, globalFeatureLoss
    from cirtorch.modules.heads.global_head import globalHead
    from cirtorch.utils.parallel import PackedSequence

    class Body(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 32, 3, stride=2, padding=1)

        def forward(self, img):
            return {"mod5": torch.relu(self.conv(img))}

    torch.manual_seed(21)
    body = Body()
    head = globalHead(pooling={"name": "GeM", "params": {"p": 2.5, "eps": 1e-6}}, normal={"name": "L2N", "params": {}}, dim=32)
    with torch.no_grad():
        head.whiten.bias.normal_(0, 0.05)
    net = ImageRetrievalNet(body, globalFeatureAlgo(globalFeatureLoss("triplet", 0.5), min_level=2, fpn_levels=1), head,
                            augment=None).eval()
    g = torch.Generator().manual_seed(22)
    sizes = [(40, 32), (36, 48), (40, 48), (24, 24)]
    imgs = [torch.randn(3, h, w, generator=g) for (h, w) in sizes]
    out = {"conv_w": body.conv.weight.detach().numpy().copy(), "conv_b": body.conv.bias.detach().numpy().copy(),
           "W": head.whiten.weight.detach().numpy().copy(), "b": head.whiten.bias.detach().numpy().copy(), "p": np.float32(2.5)}
    for i, t in enumerate(imgs):
        out[f"img{i}"] = t.numpy()
    with torch.no_grad():
        _, pred = net(img=PackedSequence(imgs), scales=[1], do_prediction=True)
        out["pred"] = pred["ret_pred"].contiguous().numpy()
        _, pred = net(img=PackedSequence(imgs), scales=[1, 2 ** -0.5, 0.5], do_prediction=True)
        out["pred_ms"] = pred["ret_pred"].contiguous().numpy()
    # a training step: 2 tuples (query, positive, 2 negatives), all images of one size
    timgs = [torch.randn(3, 32, 32, generator=g) for _ in range(8)]
    for i, t in enumerate(timgs):
        out[f"timg{i}"] = t.numpy()
    q, p_, negs = [timgs[0], timgs[4]], [timgs[1], timgs[5]], [[timgs[2], timgs[3]], [timgs[6], timgs[7]]]
    labels = [torch.tensor([-1., 1., 0., 0.]), torch.tensor([-1., 1., 0., 0.])]
    net.train()
    loss, pred = net(img=q, positive_img=p_, negative_img=negs, do_loss=True, do_prediction=False, tuple_labels=labels)
    loss["ret_loss"].backward()
    out["train_loss"] = loss["ret_loss"].detach().numpy()
    out["train_pred"] = pred["ret_pred"].detach().contiguous().numpy()
    out["train_grad_W"] = head.whiten.weight.grad.numpy().copy()
    out["train_grad_b"] = head.whiten.bias.grad.numpy().copy()
    out["train_grad_p"] = head.pool.p.grad.numpy().copy()
    out["train_grad_conv_w"] = body.conv.weight.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "net.npz"), **out)


if __name__ == "__main__":
    _import_reference()
    torch.set_num_threads(1)
    only = sys.argv[1:]            # e.g. ``make_golden.py regional`` regenerates one fixture file
    for name, fn in (("tail", tail_fixtures), ("whiten", whiten_fixtures), ("rank", rank_fixtures), ("mining", mining_fixtures),
                     ("eval", eval_fixtures), ("multiscale", multiscale_fixture), ("regional", regional_fixtures),
                     ("losses", losses_fixtures), ("net", net_fixtures)):
        if not only or name in only:
            fn()
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))
