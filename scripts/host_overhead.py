#!/usr/bin/env python
"""Host-side cost of one fused-tail call through the module API (not part of the product)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
from bench import make_head, B, C, H, W
dev = torch.device("cuda:0")
head = make_head(dev)
x = torch.relu(torch.randn((B, C, H, W), device=dev))
small = torch.relu(torch.randn((1, 64, 4, 4), device=dev))
from cirtorch_b200.modules.heads.global_head import globalHead
hs = globalHead(pooling={"name": "GeM", "params": {"p": 3, "eps": 1e-6}}, normal={"name": "L2N", "params": {}}, dim=64).to(dev).eval()
with torch.no_grad():
    for _ in range(20): head(x); hs(small)
    torch.cuda.synchronize()
    # tiny problem: GPU time ~10 us, so the wall time per call is the host cost
    t0 = time.perf_counter()
    for _ in range(300): hs(small)
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print("host cost per call (tiny problem, async): %.1f us" % ((t1 - t0) / 300 * 1e6))
    t0 = time.perf_counter()
    for _ in range(300): head(x)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("full-size wall per call incl. sync: %.1f us  (OMP_NUM_THREADS=%s)" % ((t1 - t0) / 300 * 1e6, os.environ.get("OMP_NUM_THREADS")))
