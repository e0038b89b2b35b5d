#!/usr/bin/env python
"""Phase-A probe of the tail kernel: pool-only (plain launch, no projection) and full launches at p = 3 / 2.7 for a
batch that streams from HBM (64 images, 512 MiB) and one that stays in L2 (8 images, 64 MiB)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
from cirtorch_b200 import functional as LF
from cirtorch_b200.modules.heads.global_head import globalHead
dev = torch.device("cuda:0")


def timeit(fn, n=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n

for N in (64, 8):
    xs = [torch.relu(torch.randn((N, 2048, 32, 32), device=dev)) for _ in range(2 if N == 64 else 1)]
    for p in (3.0, 2.7):
        pt = torch.full((1,), p, device=dev)
        state = {"i": 0}
        def pool():
            state["i"] += 1
            return LF.descriptor_tail(xs[state["i"] % len(xs)], p=pt, pooling="GeM", pool_only=True)
        head = globalHead(pooling={"name": "GeM", "params": {"p": p, "eps": 1e-6}}, normal={"name": "L2N", "params": {}}, dim=2048).to(dev).eval()
        def full():
            state["i"] += 1
            with torch.no_grad():
                return head(xs[state["i"] % len(xs)])
        with torch.no_grad():
            print("N=%d p=%.1f pool_only %.2f us   full %.2f us   (%.0f MB per launch)" % (N, p, timeit(pool), timeit(full), xs[0].numel() * 4 / 1e6))
