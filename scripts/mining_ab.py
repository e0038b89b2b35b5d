import sys, os, json
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
from bench import mining_case, timed_region
from cirtorch_b200 import mining as M
dev=torch.device("cuda:0")
q, pool, clusters, qidx, i2i = mining_case()
qd, pd = q.to(dev), pool.to(dev)
qc = torch.from_numpy(clusters[qidx]).to(dev, torch.int32); pc = torch.from_numpy(clusters[i2i]).to(dev, torch.int32)
ref = None
for mode, kc in (("bf16x3", None), ("bf16", 128), ("bf16", 192), ("bf16x3", None)):
    f = lambda: M.mine_hard_negatives_rows(qd, pd, qc, pc, 5, mode=mode, kc=kc)
    ms = timed_region(f, 10, 3, 1) / 10
    sel, cnt, dist = f()
    if ref is None: ref = sel.clone()
    print(mode, kc, "ms %.3f" % ms, "same sets", bool(torch.equal(sel, ref)))
