"""Where the time of a 70-query search over one SHARD goes (125k rows = 1/8 of 1M, the 8-GPU case of VERDICT item 5).

  python scripts/q70_probe.py [iters]     prints ms per search for N in (125000, 500000, 1000000), with and without the
                                          threshold pre-pass, the host time per call, and the bytes / HBM-peak floor.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch  # noqa: E402
from cirtorch_b200 import search as S  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
D, K = 2048, 100
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sizes = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (125_000, 500_000, 1_000_000)
variants = (("default", 0), ("no_prepass", S.NO_PREPASS)) if len(sys.argv) <= 3 else (("default", 0),)
for N in sizes:
    db = torch.nn.functional.normalize(torch.randn((N, D), device=dev, generator=g), dim=1)
    dbp = S.pack_rows(db, "db", "bf16")
    del db
    q = torch.nn.functional.normalize(torch.randn((70, D), device=dev, generator=g), dim=1)
    qp = S.pack_rows(q, "query", "bf16")
    ref = None
    for name, flags in variants:
        out = (torch.empty((70, K), dtype=torch.float32, device=dev), torch.empty((70, K), dtype=torch.int32, device=dev))
        for _ in range(5):
            S.search_packed(qp, dbp, K, out=out, flags=flags)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(iters):
            S.search_packed(qp, dbp, K, out=out, flags=flags)
        e1.record()
        host = (time.perf_counter() - t0) / iters * 1e3
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        # cold variant: L2 flushed between searches (each search timed on its own)
        cold = 0.0
        for _ in range(10):
            flush.zero_()
            e0.record()
            S.search_packed(qp, dbp, K, out=out, flags=flags)
            e1.record()
            torch.cuda.synchronize()
            cold += e0.elapsed_time(e1) / 10
        if ref is None:
            ref = out[1].clone()
        floor = N * D * 2 / 6548.2e9 * 1e3
        print("N=%d %-10s %.4f ms/search (cold %.4f)  host %.4f ms/call  floor %.4f ms  frac %.3f  same %s" % (
            N, name, ms, cold, host, floor, floor / ms, bool(torch.equal(out[1], ref))), flush=True)
    del dbp
