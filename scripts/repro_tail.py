import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
from cirtorch_b200 import functional as LF
dev = torch.device("cuda:0")
B, C, H, W = 64, 2048, 32, 32
x = torch.relu(torch.randn((B, C, H, W), device=dev))
p3 = torch.full((1,), 3.0, device=dev)
torch.cuda.synchronize()
mode = sys.argv[1] if len(sys.argv) > 1 else "pool"
Wt = torch.randn(2048, 2048, device=dev) * 0.01
b = torch.zeros(2048, device=dev)
for it in range(40):
    t0 = time.time()
    try:
        if mode == "pool":
            LF.descriptor_tail(x, p=p3, pooling="GeM", pool_only=True)
        else:
            LF.descriptor_tail(x, p=p3, weight=Wt, bias=b, pooling="GeM")
        torch.cuda.synchronize()
    except Exception as e:
        print("FAILED at iteration", it, "after %.3f s" % (time.time() - t0), str(e)[:80])
        sys.exit(1)
print("ok", mode)
