#!/usr/bin/env python
"""Small driver for ncu captures: runs a few launches of one hot kernel at its BASELINE.json size.
    python scripts/profile_target.py tail|search10k|search70|regions|bwd|mining [n_launches]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
if what == "tail":
    from bench import make_head, B, C, H, W
    head = make_head(dev, p=float(os.environ.get("CIR_PROFILE_P", "3.0")))
    xs = [torch.relu(torch.randn((B, C, H, W), device=dev)) for _ in range(2)]
    with torch.no_grad():
        for i in range(n):
            head(xs[i & 1])
elif what in ("search10k", "search70"):
    from cirtorch_b200 import search as S
    N, D = 1_000_000, 2048
    Q = 10_000 if what == "search10k" else 70
    if os.environ.get("CIR_PROFILE_Q"):
        Q = int(os.environ["CIR_PROFILE_Q"])
    dbp = torch.empty((N, D), dtype=torch.bfloat16, device=dev)
    for a in range(0, N, 125_000):
        blk = torch.randn((125_000, D), device=dev)
        S.pack_rows(blk / blk.norm(dim=1, keepdim=True), "db", "bf16", out=dbp[a:a + 125_000])
    q = torch.randn((Q, D), device=dev)
    qp = S.pack_rows(q / q.norm(dim=1, keepdim=True), "query", "bf16")
    tau = None
    if len(sys.argv) > 3 and sys.argv[3] == "tau":
        s0, _ = S.search_packed(qp, dbp, 100)
        tau = (s0[:, -1] - 1e-4).contiguous()
    for i in range(n):
        S.search_packed(qp, dbp, 100, tau0=tau)
elif what == "regions":
    from cirtorch_b200 import functional as LF
    x = torch.relu(torch.randn((64, 2048, 32, 32), device=dev))
    regs = [(0, 0, 32, 32)] + LF.rmac_regions(32, 32, 3)
    p3 = torch.full((1,), 3.0, device=dev)
    for i in range(n):
        LF.region_pool(x, regs, p=p3, pooling="GeM")
elif what == "bwd":
    from cirtorch_b200.functional import _gem_bwd_launch
    x = torch.relu(torch.randn((64, 2048, 32, 32), device=dev))
    gg = torch.rand((64, 2048), device=dev) + 0.1
    dg = torch.randn((64, 2048), device=dev)
    pt = torch.full((1,), 3.0, device=dev)
    for i in range(n):
        _gem_bwd_launch(x, pt, 1e-6, gg, dg, True, True)
elif what == "mining":
    from cirtorch_b200.mining import mine_hard_negatives_rows
    q = torch.randn((2000, 2048), device=dev)
    p = torch.randn((20000, 2048), device=dev)
    q, p = q / q.norm(dim=1, keepdim=True), p / p.norm(dim=1, keepdim=True)
    qc = torch.randint(0, 700, (2000,), device=dev, dtype=torch.int32)
    pc = torch.randint(0, 700, (20000,), device=dev, dtype=torch.int32)
    for i in range(n):
        mine_hard_negatives_rows(q, p, qc, pc, 5)
torch.cuda.synchronize()
print("ok", what)
