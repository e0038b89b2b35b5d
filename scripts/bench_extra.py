#!/usr/bin/env python
"""Timings of the remaining BASELINE.json configs (3: hard-negative mining, 5: alpha-QE + DBA) on one GPU.
Writes one JSON object per config to stdout (kept under profiles/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import numpy as np
import torch
from cirtorch_b200 import search as S, rerank, mining, _lib
dev = torch.device("cuda:0")


def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def unit(n, d, g=None):
    x = torch.randn((n, d), device=dev, generator=g)
    return x / x.norm(dim=1, keepdim=True)

# ---- config 3: 2000 queries x 20000 pool, nnum 5, D 2048
g = torch.Generator(device=dev).manual_seed(0)
centres = unit(700, 2048, g)
pc = torch.randint(0, 700, (20000,), device=dev, generator=g)
qc = torch.randint(0, 700, (2000,), device=dev, generator=g)
pool = centres[pc] * 0.6 + unit(20000, 2048, g); pool = (pool / pool.norm(dim=1, keepdim=True)).contiguous()
q = centres[qc] * 0.6 + unit(2000, 2048, g); q = (q / q.norm(dim=1, keepdim=True)).contiguous()
_lib.launch_count(reset=True)
ms = timeit(lambda: mining.mine_hard_negatives_rows(q, pool, qc.int(), pc.int(), 5))
launches = _lib.launch_count() // 7
t0 = time.perf_counter()
sims = pool.cpu() @ q.cpu().t()            # reference path on the host: torch.mm + torch.sort (+ Python loop, not timed)
torch.sort(sims, dim=0, descending=True)
cpu_s = time.perf_counter() - t0
print(json.dumps({"config": "mining 2000 q x 20000 pool x 2048, nnum=5 (BASELINE configs[2])", "gpu_ms": ms, "launches": launches,
                  "queries_per_s": 2000 / (ms * 1e-3), "flop": 2 * 2000 * 20000 * 2048 * 3,
                  "cpu_reference_mm_sort_s": cpu_s, "cpu_threads": torch.get_num_threads(),
                  "note": "bf16x3 search (K=6144) with cluster exclusion + fp32 re-score + greedy filter; CPU = torch.mm + torch.sort "
                          "of tuples_dataset.py:317-319 without its Python greedy loop"}))
del pool, q, sims

# ---- config 5: alpha-QE (k=10, alpha=3) with 10k queries over a 1M x 2048 database, and DBA
N, D = 1_000_000, 2048
db = torch.empty((N, D), device=dev)
for a in range(0, N, 125_000):
    db[a:a + 125_000] = unit(125_000, D, g)
index = S.Index(db, mode="bf16")
q = unit(10_000, D, g)
def qe():
    q2 = rerank.alpha_qe_rows(q, index, k=10, alpha=3.0)
    return index.search_rows(q2, 100)
ms_qe = timeit(qe, n=3, warm=1)
ms_plain = timeit(lambda: index.search_rows(q, 100), n=3, warm=1)
print(json.dumps({"config": "alpha-QE k=10 alpha=3, 10k queries x 1M x 2048, then top-100 re-search (BASELINE configs[4], 1 GPU)",
                  "gpu_ms": ms_qe, "queries_per_s": 10_000 / (ms_qe * 1e-3), "plain_search_ms": ms_plain,
                  "note": "search top-10 (bf16 + fp32 re-score) + cir_qe_aggregate + search top-100 (bf16 + fp32 re-score)"}))
# DBA over a 20k-row slice of the 1M database (the full pass is 50 such slices)
ms_dba = timeit(lambda: rerank.dba_rows(db, k=10, alpha=3.0, index=index, row_begin=0, row_end=20_000, chunk=10_000), n=2, warm=1)
print(json.dumps({"config": "DBA k=10 alpha=3 over 1M x 2048: 20,000-row slice (2 x 10k-query searches + aggregation)", "gpu_ms": ms_dba,
                  "rows_per_s": 20_000 / (ms_dba * 1e-3), "full_1M_pass_s_extrapolated": ms_dba * 50 / 1e3}))
