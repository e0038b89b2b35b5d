"""Stress: the fused tail must return bit-identical descriptors on every launch (ring hand-over races show up here)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
from cirtorch_b200.modules.heads.global_head import globalHead
from cirtorch_b200 import functional as LF
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
torch.manual_seed(0)
bad = 0
for shape in ((64, 2048, 32, 32), (16, 512, 16, 16), (5, 256, 24, 40)):
    x = torch.relu(torch.randn(shape, device=dev))
    other = torch.relu(torch.randn(shape, device=dev))
    for p in (3.0, 2.7):
        head = globalHead(pooling={"name": "GeM", "params": {"p": p, "eps": 1e-6}}, normal={"name": "L2N", "params": {}}, dim=shape[1]).to(dev).eval()
        pt = torch.full((1,), p, device=dev)
        with torch.no_grad():
            ref = head(x).clone()
            ref_pool = LF.descriptor_tail(x, p=pt, pooling="GeM", pool_only=True).clone()
            diff = 0
            for i in range(n):
                if i % 3 == 0:
                    head(other)                                  # different data through the ring in between
                diff += int(not torch.equal(head(x), ref))
                diff += int(not torch.equal(LF.descriptor_tail(x, p=pt, pooling="GeM", pool_only=True), ref_pool))
        print("shape %s p=%.1f: %d of %d launches differ" % (shape, p, diff, 2 * n), flush=True)
        bad += diff
sys.exit(1 if bad else 0)
