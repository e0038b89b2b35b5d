#!/usr/bin/env python
"""Experiment: the CTA-pair (cta_group::2) search kernel against the single-CTA kernel.  Run twice:
    python scripts/pair_check.py ref  -> writes reference lists;   CIR_SEARCH_2CTA=1 python scripts/pair_check.py cmp"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
from cirtorch_b200 import search as S
dev = torch.device("cuda:0")
mode = sys.argv[1]
cases = [(300, 5000, 128, 10), (129, 70000, 64, 100), (1000, 200000, 256, 50), (257, 1000, 64, 5)]
if len(sys.argv) > 2: cases = cases[:int(sys.argv[2])]
out = {}
for ci, (Q, N, D, k) in enumerate(cases):
    g = torch.Generator(device=dev).manual_seed(ci)
    db = torch.randn((N, D), device=dev, generator=g); db = db / db.norm(dim=1, keepdim=True)
    q = torch.randn((Q, D), device=dev, generator=g); q = q / q.norm(dim=1, keepdim=True)
    qp, dbp = S.pack_rows(q, "query", "bf16"), S.pack_rows(db, "db", "bf16")
    s, i = S.search_packed(qp, dbp, k)
    torch.cuda.synchronize()
    out[ci] = (s.cpu(), i.cpu())
    print("case", ci, (Q, N, D, k), "done", flush=True)
path = "/tmp/pair_ref.pt"
if mode == "ref":
    torch.save(out, path)
else:
    ref = torch.load(path)
    for ci in out:
        same_i = bool(torch.equal(ref[ci][1], out[ci][1])); same_s = bool(torch.equal(ref[ci][0], out[ci][0]))
        print("case", ci, "idx equal", same_i, "scores equal", same_s, "max |ds|", float((ref[ci][0] - out[ci][0]).abs().max()))
