"""GPU time of the head's forward + backward without the Python launch overhead: the step is captured in a CUDA graph and
replayed (the eager loop of scripts/bench_round2.py is bound by the host at ~0.5 ms per step)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
from cirtorch_b200.modules.heads.global_head import globalHead
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for p in (3.0, 2.7):
    torch.manual_seed(0)
    head = globalHead(pooling={"name": "GeM", "params": {"p": p, "eps": 1e-6}}, normal={"name": "L2N", "params": {}}, dim=2048).to(dev)
    x = torch.relu(torch.randn((64, 2048, 32, 32), device=dev, generator=g)).requires_grad_(True)
    tgt = torch.randn((2048, 64), device=dev, generator=g)
    params = [x] + list(head.parameters())

    def step():
        grads = torch.autograd.grad((head(x) * tgt).sum(), params)
        return grads

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    # eager timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 20
    try:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = step()
        torch.cuda.synchronize()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            graph.replay()
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"what": "globalHead forward + backward as a CUDA graph, 64x2048x32x32, p=%.1f" % p,
                          "ms_graph": e0.elapsed_time(e1) / 20, "ms_eager": eager}))
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"what": "graph capture failed", "error": repr(e)[:300], "ms_eager": eager}))
    del x, head
