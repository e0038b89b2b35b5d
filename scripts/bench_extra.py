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
del db, index, q
torch.cuda.empty_cache()

# ---- config 1: ResNet50-GeM extract_vectors + exhaustive search, rOxford5k shape (70 queries vs 4,993 database images),
# synthetic 1024 px images, random-init weights.  The torchvision backbone is stock PyTorch (not part of the product);
# a 32-image sample is timed and scaled to the 5,063 images of the set.
from cirtorch_b200.extract import resnet50_gem, extract_vectors
net = resnet50_gem().to(dev).eval()
imgs = torch.randn((32, 3, 1024, 1024), generator=torch.Generator().manual_seed(0))
t_sample = None
for rep in range(2):            # first pass = cuDNN autotune / warm-up
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vecs = extract_vectors(net, imgs, image_size=1024, transform=None, batch_size=8, device=dev)
    torch.cuda.synchronize(); t_sample = time.perf_counter() - t0
fm = torch.relu(torch.randn((8, 2048, 32, 32), device=dev))
with torch.no_grad():
    ms_tail8 = timeit(lambda: net.ret_head(fm), n=20)
dbv = unit(4993, 2048, g); qv = unit(70, 2048, g)
ms_rank = timeit(lambda: S.rank(dbv.t(), qv.t()), n=5)                      # full N x Q ranking like np.argsort(-scores, axis=0)
ms_topk = timeit(lambda: S.search_topk(qv.t(), dbv.t(), 100), n=10)
# the reference's CPU path on the same host, 2 images (torchvision backbone + oracle head), then np.dot + np.argsort
sys.path.insert(0, ROOT)
from oracle import cirtorch_oracle as O
net_cpu = resnet50_gem().eval()
with torch.no_grad():
    t0 = time.perf_counter()
    f = net_cpu.body(imgs[:2])
    O.head_forward(f, 3.0, 1e-6, net_cpu.ret_head.whiten.weight, net_cpu.ret_head.whiten.bias)
    cpu_img_s = (time.perf_counter() - t0) / 2
t0 = time.perf_counter(); O.rank(dbv.t().cpu().numpy(), qv.t().cpu().numpy()); cpu_rank_s = time.perf_counter() - t0
print(json.dumps({"config": "ResNet50-GeM extract_vectors + exhaustive search, 70 q vs 4,993 db, 1024 px (BASELINE configs[0])",
                  "gpu_images_per_s": 32 / t_sample, "gpu_extract_5063_images_s_extrapolated": 5063 * t_sample / 32,
                  "gpu_tail_ms_per_8_images": ms_tail8, "gpu_full_ranking_ms": ms_rank, "gpu_top100_ms": ms_topk,
                  "cpu_reference_s_per_image": cpu_img_s, "cpu_extract_5063_images_s_extrapolated": 5063 * cpu_img_s,
                  "cpu_rank_s": cpu_rank_s, "cpu_threads": torch.get_num_threads(),
                  "note": "backbone = stock torchvision fp32 (excluded from the roofline); sample of 32 images (GPU) / 2 images (CPU) "
                          "scaled linearly; includes the H2D copy of the images and the D2H copy of the descriptors"}))
