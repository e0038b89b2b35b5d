#!/usr/bin/env python
"""Round-2 side measurements on one GPU (device time, CUDA events): head forward + backward, tuple losses, whitenapply at
1M x 2048, the full N x Q ranking at 70 x 1M, the region kernel.  One JSON object per line (kept under profiles/)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
from cirtorch_b200 import search as S, functional as LF, _lib
from cirtorch_b200.modules.heads.global_head import globalHead
from cirtorch_b200.modules.losses import contrastive_loss, triplet_loss
from cirtorch_b200.utils import whiten as W
dev = torch.device("cuda:0")
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
    {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
what = set(sys.argv[1:]) or {"train", "loss", "whiten", "rank", "regions", "eager"}


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def unit(n, d, g):
    x = torch.randn((n, d), device=dev, generator=g)
    return x / x.norm(dim=1, keepdim=True)

g = torch.Generator(device=dev).manual_seed(0)
if "train" in what:
    for p in (3.0, 2.7):
        torch.manual_seed(0)
        head = globalHead(pooling={"name": "GeM", "params": {"p": p, "eps": 1e-6}}, normal={"name": "L2N", "params": {}}, dim=2048).to(dev)
        x = torch.relu(torch.randn((64, 2048, 32, 32), device=dev, generator=g)).requires_grad_(True)
        tgt = torch.randn((2048, 64), device=dev, generator=g)

        def step():
            head.zero_grad(set_to_none=True); x.grad = None
            (head(x) * tgt).sum().backward()
        _lib.launch_count(reset=True)
        ms = timeit(step, n=20)
        launches = _lib.launch_count() / 23
        with torch.no_grad():
            ms_f = timeit(lambda: head(x), n=20)
        print(json.dumps({"what": "globalHead forward + backward, 64x2048x32x32, whitening 2048->2048, p=%.1f" % p, "ms": ms, "forward_only_ms": ms_f,
                          "own_launches_per_step": launches,
                          "note": "grads to x (512 MiB written), p, W, b; includes the (out * tgt).sum() of the test loss and torch's autograd glue"}))
        del x, head
if "loss" in what:
    nt, S_, D = 64, 7, 2048
    rows = unit(nt * S_, D, g).requires_grad_(True)
    label = torch.tensor([-1, 1, 0, 0, 0, 0, 0] * nt, dtype=torch.float32, device=dev)
    msk = torch.arange(nt, device=dev).repeat_interleave(S_)
    for name, fn in (("contrastive", lambda: contrastive_loss(rows.t(), label=label, margin=0.7)),
                     ("triplet", lambda: triplet_loss(rows.t(), label=label, label_msk=msk, margin=0.1))):
        def step():
            rows.grad = None
            fn().backward()
        print(json.dumps({"what": "%s loss forward + backward, 64 tuples x 7 x 2048" % name, "ms": timeit(step, n=50),
                          "note": "one kernel (loss + gradient) + the host sync that counts the tuples, like the reference"}))
if "whiten" in what:
    N, D = 1_000_000, 2048
    X = torch.empty((N, D), device=dev)
    for a in range(0, N, 125_000):
        X[a:a + 125_000] = unit(125_000, D, g)
    m = torch.randn((D, 1), device=dev, generator=g, dtype=torch.float64) * 0.01
    P = torch.randn((D, D), device=dev, generator=g, dtype=torch.float64) / D ** 0.5
    Xt = X.t()                                            # the D x N view cirtorch passes around
    ms = timeit(lambda: W.whitenapply(Xt, m, P), n=3, warm=1)
    fl = 2.0 * N * D * D
    print(json.dumps({"what": "whitenapply 1M x 2048 -> 2048 (A6)", "ms": ms, "algorithmic_tflops": fl / (ms * 1e-3) / 1e12,
                      "executed_bf16_tflops": 3 * fl / (ms * 1e-3) / 1e12, "frac_of_sustained_bf16_peak": 3 * fl / (ms * 1e-3) / 1e12 / PK["bf16_tflops_sustained"],
                      "note": "pack X and W as bf16x3 (K = 6144), tcgen05 GEMM, bias (-W m) + L2N, fp64 result copy; 8.6 algorithmic TFLOP"}))
    del X, Xt
    torch.cuda.empty_cache()
if "rank" in what:
    N, D, Q = 1_000_000, 2048, 70
    db = torch.empty((N, D), device=dev)
    for a in range(0, N, 125_000):
        db[a:a + 125_000] = unit(125_000, D, g)
    q = unit(Q, D, g)
    ms = timeit(lambda: S.rank(db.t(), q.t()), n=3, warm=1)
    sc = S.scores_dense_rows(q, db, mode="bf16x3")
    ms_sort = timeit(lambda: S.argsort_rows_desc(sc), n=3, warm=1)
    ms_torch = timeit(lambda: torch.sort(sc, dim=1, descending=True), n=3, warm=1)
    print(json.dumps({"what": "full N x Q ranking (np.argsort(-scores, axis=0)), 70 x 1M x 2048", "ms": ms, "sort_only_ms": ms_sort,
                      "torch_sort_ms": ms_torch, "note": "rank() = pack + dense bf16x3 scores + cir_sort_rows_desc"}))
    del db, sc
    torch.cuda.empty_cache()
if "regions" in what:
    x = torch.relu(torch.randn((64, 2048, 32, 32), device=dev, generator=g))
    regs = [(0, 0, 32, 32)] + LF.rmac_regions(32, 32, 3)
    p3 = torch.full((1,), 3.0, device=dev)
    ms = timeit(lambda: LF.region_pool(x, regs, p=p3, pooling="GeM"), n=10)
    print(json.dumps({"what": "region_pool 15 regions, 64x2048x32x32", "ms": ms, "gbs": x.numel() * 4 / (ms * 1e-3) / 1e9,
                      "frac_of_hbm_peak": x.numel() * 4 / (ms * 1e-3) / 1e9 / PK["hbm_gbs"]}))
if "eager" in what:
    # the reference's head as stock eager PyTorch ops ON THE B200 (the "unfused GPU" comparison):
    # clamp / pow / avg_pool2d / pow, norm / div, Linear, norm / div -- pools.py:37-38, normalizations.py:15-16, global_head.py:52-67
    import torch.nn.functional as F
    x = torch.relu(torch.randn((64, 2048, 32, 32), device=dev, generator=g))
    Wt = torch.randn((2048, 2048), device=dev, generator=g) * 0.02
    b = torch.zeros(2048, device=dev)
    for p in (3.0, 2.7):
        pt = torch.ones(1, device=dev) * p

        def eager():
            v = F.avg_pool2d(x.clamp(min=1e-6).pow(pt), (32, 32)).pow(1.0 / pt)
            v = (v / (torch.norm(v, p=2, dim=1, keepdim=True) + 1e-6).expand_as(v)).squeeze(-1).squeeze(-1)
            v = F.linear(v, Wt, b)
            return (v / (torch.norm(v, p=2, dim=1, keepdim=True) + 1e-6).expand_as(v)).permute(1, 0)
        with torch.no_grad():
            ms = timeit(eager, n=20)
        print(json.dumps({"what": "reference head as eager PyTorch ops on the B200 (unfused), 64x2048x32x32, p=%.1f" % p, "ms": ms,
                          "gbs_algorithmic": 554_180_608 / (ms * 1e-3) / 1e9}))
