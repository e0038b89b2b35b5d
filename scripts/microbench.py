#!/usr/bin/env python
"""Phase breakdown timings used while optimising (not part of the product)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
from cirtorch_b200 import functional as LF, search as S
dev = torch.device("cuda:0")

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "tail"):
    B, C, H, W = 64, 2048, 32, 32
    xs = [torch.relu(torch.randn((B, C, H, W), device=dev)) for _ in range(2)]
    Wt = torch.randn(2048, 2048, device=dev) * 0.01
    b = torch.zeros(2048, device=dev)
    p3 = torch.full((1,), 3.0, device=dev); p27 = torch.full((1,), 2.7, device=dev)
    i = [0]
    def nxt():
        i[0] ^= 1; return xs[i[0]]
    gb = B * C * H * W * 4 / 1e6
    for name, fn in [
        ("pool_only p=3", lambda: LF.descriptor_tail(nxt(), p=p3, pooling="GeM", pool_only=True)),
        ("pool_only p=2.7", lambda: LF.descriptor_tail(nxt(), p=p27, pooling="GeM", pool_only=True)),
        ("pool_only MAC", lambda: LF.descriptor_tail(nxt(), pooling="MAC", pool_only=True)),
        ("no_whiten p=3", lambda: LF.descriptor_tail(nxt(), p=p3, pooling="GeM", do_whitening=False)),
        ("full p=3", lambda: LF.descriptor_tail(nxt(), p=p3, weight=Wt, bias=b, pooling="GeM")),
        ("full p=2.7", lambda: LF.descriptor_tail(nxt(), p=p27, weight=Wt, bias=b, pooling="GeM")),
        ("torch clone (copy)", lambda: nxt().clone()),
        ("torch sum (read)", lambda: nxt().sum()),
    ]:
        ms = timeit(fn)
        print(f"tail {name:22s} {ms*1e3:8.1f} us  {gb/ms:8.1f} GB/s(x only)")
if what in ("all", "bwd"):
    from cirtorch_b200.functional import _gem_bwd_launch
    B, C, H, W = 64, 2048, 32, 32
    x = torch.relu(torch.randn((B, C, H, W), device=dev))
    g = torch.rand((B, C), device=dev) + 0.1
    dg = torch.randn((B, C), device=dev)
    for pv in (3.0, 2.7):
        pt = torch.full((1,), pv, device=dev)
        for want_s in (False, True):
            ms = timeit(lambda: _gem_bwd_launch(x, pt, 1e-6, g, dg, True, want_s), n=10)
            print(f"gem_bwd p={pv} dp={want_s}: {ms*1e3:8.1f} us  {2*B*C*H*W*4/ms/1e6:8.1f} GB/s (read x + write dx)")
    from cirtorch_b200.modules.heads.global_head import globalHead
    head = globalHead(pooling={"name": "GeM", "params": {"p": 3, "eps": 1e-6}}, normal={"name": "L2N", "params": {}}, dim=2048).to(dev)
    xg = x.clone().requires_grad_(True)
    tgt = torch.randn(2048, B, device=dev)
    def fb():
        out = head(xg)
        (out * tgt).sum().backward()
        xg.grad = None
    ms = timeit(fb, n=10)
    print(f"globalHead forward + backward (x, p, W, b grads): {ms*1e3:8.1f} us")
if what in ("all", "tail", "stamps"):
    import ctypes as C
    from cirtorch_b200 import _lib
    lib = _lib.load()
    B, Cc, H, W = 64, 2048, 32, 32
    x = torch.relu(torch.randn((B, Cc, H, W), device=dev))
    Wt = torch.randn(2048, 2048, device=dev) * 0.01
    b = torch.zeros(2048, device=dev); p3 = torch.full((1,), 3.0, device=dev)
    out = torch.empty((B, 2048), device=dev)
    need = C.c_size_t(0); lib.cir_tail_workspace_bytes(B, Cc, 2048, C.byref(need))
    ws = torch.zeros(need.value, dtype=torch.uint8, device=dev)
    for it in range(3):
        rc = lib.cir_tail_fwd(_lib.ptr(x), B, Cc, H, W, _lib.ptr(p3), 0, 1e-6, 1e-6, 0, _lib.ptr(Wt), _lib.ptr(b), 2048,
                              _lib.ptr(out), 2048, None, _lib.ptr(ws), ws.numel(), 0x80000000, None)
        assert rc == 0
        torch.cuda.synchronize()
    st = ws[need.value - 65536:].view(torch.int64).view(-1, 8)[:148, :8].cpu().double()
    st = st[st[:, 3] > st[:, 2]]      # CTAs that own an output chunk
    t0 = st[:, 0].min()
    names = ["start", "endA", "sync1", "endB", "sync2", "end", "B:Wready", "B:kloop"]
    for j in range(8):
        col = (st[:, j] - t0) / 1e3
        print(f"stamp {names[j]:6s}: min {col.min():8.1f} us  mean {col.mean():8.1f}  max {col.max():8.1f}")
if what in ("all", "search"):
    N, D = 1_000_000, 2048
    dbp = torch.empty((N, D), dtype=torch.bfloat16, device=dev)
    rows = None
    for a in range(0, N, 125_000):
        blk = torch.randn((125_000, D), device=dev)
        S.pack_rows(blk / blk.norm(dim=1, keepdim=True), "db", "bf16", out=dbp[a:a + 125_000])
    for Q in (70, 128, 1024, 10_000):
        q = torch.randn((Q, D), device=dev); q = q / q.norm(dim=1, keepdim=True)
        qp = S.pack_rows(q, "query", "bf16")
        s, i = S.search_packed(qp, dbp, 100)
        tau = (s[:, -1] - 1e-4).contiguous()
        n = 3 if Q > 2000 else 10
        ms = timeit(lambda: S.search_packed(qp, dbp, 100), n=n)
        ms_t = timeit(lambda: S.search_packed(qp, dbp, 100, tau0=tau), n=n)
        s2, i2 = S.search_packed(qp, dbp, 100, tau0=tau)
        print(f"search Q={Q:6d}: {ms:8.3f} ms ({2*Q*N*D/ms/1e9:7.1f} TF, {N*D*2/ms/1e6:7.1f} GB/s) | with exact tau0: {ms_t:8.3f} ms "
              f"({2*Q*N*D/ms_t/1e9:7.1f} TF, {N*D*2/ms_t/1e6:7.1f} GB/s) same={bool(torch.equal(i, i2))}")
    ms = timeit(lambda: dbp.sum(dtype=torch.float32), n=5)
    print(f"torch sum over bf16 db: {ms:.3f} ms {N*D*2/ms/1e6:.1f} GB/s")

if what == "shard":
    # one shard of an 8/4/2-way split 1M database: how the threshold sample size moves the 10k-query search
    D = 2048
    for N in [int(a) for a in (sys.argv[2] if len(sys.argv) > 2 else "125000,250000,500000").split(",")]:
        dbp = torch.empty((N, D), dtype=torch.bfloat16, device=dev)
        for a in range(0, N, 125_000):
            blk = torch.randn((min(125_000, N - a), D), device=dev)
            S.pack_rows(blk / blk.norm(dim=1, keepdim=True), "db", "bf16", out=dbp[a:a + blk.shape[0]])
        for Q in (10_000, 70):
            q = torch.randn((Q, D), device=dev); q = q / q.norm(dim=1, keepdim=True)
            qp = S.pack_rows(q, "query", "bf16")
            s, i = S.search_packed(qp, dbp, 100)
            tau = (s[:, -1] - 1e-4).contiguous()
            ms = timeit(lambda: S.search_packed(qp, dbp, 100), n=5)
            ms_t = timeit(lambda: S.search_packed(qp, dbp, 100, tau0=tau), n=5)
            print(f"shard N={N} Q={Q} sample={os.environ.get('CIR_DEBUG_SAMPLE_ROWS','rule')}: {ms:8.3f} ms ({2*Q*N*D/ms/1e9:7.1f} TF) | exact tau0 {ms_t:8.3f} ms")
        del dbp
if what == "kernels":
    # per-kernel device times (CUPTI through torch.profiler) of one search: N rows, Q queries
    from torch.profiler import profile, ProfilerActivity
    D = 2048
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    dbp = torch.empty((N, D), dtype=torch.bfloat16, device=dev)
    for a in range(0, N, 125_000):
        blk = torch.randn((min(125_000, N - a), D), device=dev)
        S.pack_rows(blk / blk.norm(dim=1, keepdim=True), "db", "bf16", out=dbp[a:a + blk.shape[0]])
    for Q in (10_000, 70):
        q = torch.randn((Q, D), device=dev); q = q / q.norm(dim=1, keepdim=True)
        qp = S.pack_rows(q, "query", "bf16")
        for _ in range(3): S.search_packed(qp, dbp, 100)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5): S.search_packed(qp, dbp, 100)
            torch.cuda.synchronize()
        for e in prof.key_averages():
            if e.device_time_total > 0:
                print(f"kernels N={N} Q={Q}: {e.key[:70]:70s} x{e.count:3d}  {e.device_time_total / e.count:10.1f} us")
if what == "regions":
    from cirtorch_b200.modules.pools import GeM, MAC, Rpool
    B, C, H, W = 64, 2048, 32, 32
    xs = [torch.relu(torch.randn((B, C, H, W), device=dev)) for _ in range(2)]
    i = [0]
    def nxt():
        i[0] ^= 1; return xs[i[0]]
    regs = [(0, 0, H, W)] + LF.rmac_regions(H, W, 3)
    p3 = torch.full((1,), 3.0, device=dev); p27 = torch.full((1,), 2.7, device=dev)
    gb = B * C * H * W * 4 / 1e6
    lin = torch.nn.Linear(C, C).to(dev)
    rp = Rpool(GeM(p=3), whiten=lin, L=3).to(dev)
    with torch.no_grad():
        for name, fn in [
            ("region_pool GeM p=3 (%d regions)" % len(regs), lambda: LF.region_pool(nxt(), regs, p=p3, pooling="GeM")),
            ("region_pool GeM p=2.7", lambda: LF.region_pool(nxt(), regs, p=p27, pooling="GeM")),
            ("region_pool MAC", lambda: LF.region_pool(nxt(), regs, pooling="MAC")),
            ("Rpool(GeM)+whiten+aggregate", lambda: rp(nxt())),
        ]:
            ms = timeit(fn)
            print(f"regions {name:36s} {ms*1e3:8.1f} us  {gb/ms:8.1f} GB/s(x only)")
if what == "mining":
    from torch.profiler import profile, ProfilerActivity
    from cirtorch_b200.mining import mine_hard_negatives_rows
    g = torch.Generator(device=dev).manual_seed(0)
    def unit(n, d):
        x = torch.randn((n, d), device=dev, generator=g); return x / x.norm(dim=1, keepdim=True)
    centres = unit(700, 2048)
    pc = torch.randint(0, 700, (20000,), device=dev, generator=g); qc = torch.randint(0, 700, (2000,), device=dev, generator=g)
    pool = centres[pc] * 0.6 + unit(20000, 2048); pool = (pool / pool.norm(dim=1, keepdim=True)).contiguous()
    q = centres[qc] * 0.6 + unit(2000, 2048); q = (q / q.norm(dim=1, keepdim=True)).contiguous()
    for _ in range(3): mine_hard_negatives_rows(q, pool, qc.int(), pc.int(), 5)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): mine_hard_negatives_rows(q, pool, qc.int(), pc.int(), 5)
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if e.device_time_total > 0:
            print(f"mining: {e.key[:80]:80s} x{e.count:3d}  {e.device_time_total / e.count:10.1f} us")
    print("mining total per call: %.1f us" % (timeit(lambda: mine_hard_negatives_rows(q, pool, qc.int(), pc.int(), 5), n=10) * 1e3))
if what == "midsize":
    D = 2048
    for (Q, N, mode) in ((2000, 20000, "bf16"), (2000, 20000, "bf16x3"), (5000, 50000, "bf16"), (1000, 10000, "bf16")):
        x = torch.randn((N, D), device=dev); x = x / x.norm(dim=1, keepdim=True)
        q = torch.randn((Q, D), device=dev); q = q / q.norm(dim=1, keepdim=True)
        dbp = S.pack_rows(x, "db", mode); qp = S.pack_rows(q, "query", mode)
        ms = timeit(lambda: S.search_packed(qp, dbp, 100), n=10)
        Kd = dbp.shape[1]
        print(f"midsize Q={Q} N={N} {mode}: {ms*1e3:8.1f} us  {2*Q*N*Kd/ms/1e9:7.1f} TF  sample={os.environ.get('CIR_DEBUG_SAMPLE_ROWS','rule')}")
if what == "smid":
    # is the phase-A straggler pattern tied to the SM?  per-CTA end of phase A over several launches + the SM id
    import ctypes as C
    from cirtorch_b200 import _lib
    lib = _lib.load()
    B, Cc, H, W = 64, 2048, 32, 32
    xs = [torch.relu(torch.randn((B, Cc, H, W), device=dev)) for _ in range(2)]
    Wt = torch.randn(2048, 2048, device=dev) * 0.01
    b = torch.zeros(2048, device=dev); p3 = torch.full((1,), 3.0, device=dev)
    out = torch.empty((B, 2048), device=dev)
    need = C.c_size_t(0); lib.cir_tail_workspace_bytes(B, Cc, 2048, C.byref(need))
    ws = torch.zeros(need.value, dtype=torch.uint8, device=dev)
    runs = []
    for it in range(8):
        x = xs[it & 1]
        rc = lib.cir_tail_fwd(_lib.ptr(x), B, Cc, H, W, _lib.ptr(p3), 0, 1e-6, 1e-6, 0, _lib.ptr(Wt), _lib.ptr(b), 2048,
                              _lib.ptr(out), 2048, None, _lib.ptr(ws), ws.numel(), 0x80000000, None)
        assert rc == 0
        torch.cuda.synchronize()
        st = ws[need.value - 65536:].view(torch.int64).view(-1, 8)[:148, :8].cpu().double()
        runs.append(st)
    import numpy as np
    durA = np.stack([(r[:, 1] - r[:, 0]).numpy() / 1e3 for r in runs[2:]])       # [runs, cta] us
    smid = np.stack([r[:, 7].numpy() for r in runs[2:]])
    print("smid: cta->sm mapping identical across launches:", bool((smid == smid[0]).all()))
    print("smid: per-launch mean %s  max %s" % (np.round(durA.mean(1), 1), np.round(durA.max(1), 1)))
    m = durA.mean(0)
    order = np.argsort(-m)
    print("smid: slowest CTAs (cta, sm, mean us, std):", [(int(c), int(smid[0][c]), round(float(m[c]), 1), round(float(durA[:, c].std()), 2)) for c in order[:12]])
    print("smid: fastest CTAs:", [(int(c), int(smid[0][c]), round(float(m[c]), 1)) for c in order[-8:]])
    print("smid: corr between launches of per-CTA duration:", np.round(np.corrcoef(durA)[0, 1:], 2))
    bysm = {}
    for c in range(148): bysm.setdefault(int(smid[0][c]) // 2 // 9, []).append(m[c])
    print("smid: mean by group of 18 SMs:", {k: round(float(np.mean(v)), 1) for k, v in sorted(bysm.items())})
if what == "interleave":
    # does a cooperative tail launch cost more right after an ordinary kernel (the real pipeline: backbone -> tail)?
    from bench import make_head
    B, C, H, W = 64, 2048, 32, 32
    head = make_head(dev)
    xs = [torch.relu(torch.randn((B, C, H, W), device=dev)) for _ in range(2)]
    small = torch.zeros(1024, device=dev)
    big = torch.randn((64, 2048, 32, 32), device=dev)
    i = [0]
    def tail_only():
        i[0] ^= 1
        with torch.no_grad(): head(xs[i[0]])
    def small_only(): small.add_(1.0)
    def tail_small():
        tail_only(); small.add_(1.0)
    def relu_only(): torch.relu_(big)
    def tail_relu():
        tail_only(); torch.relu_(big)
    for name, fn in [("tail", tail_only), ("small kernel", small_only), ("tail + small kernel", tail_small),
                     ("relu 512MiB", relu_only), ("tail + relu 512MiB", tail_relu)]:
        ms = timeit(fn, n=50, warm=5)
        print(f"interleave {name:24s} {ms*1e3:8.1f} us per iteration")
    # first steps after an idle gap
    import time as _t
    for gap in (0.0, 0.002, 0.05):
        torch.cuda.synchronize(); _t.sleep(gap)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
        evs[0].record()
        for k in range(12):
            tail_only(); evs[k + 1].record()
        torch.cuda.synchronize()
        print(f"interleave after {gap*1e3:.0f} ms idle, per-step us:", [round(evs[k].elapsed_time(evs[k + 1]) * 1e3, 1) for k in range(12)])
