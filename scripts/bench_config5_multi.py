#!/usr/bin/env python
"""BASELINE.json configs[4] on N GPUs (torchrun): alpha-QE (k=10, alpha=3) of 10k queries over a row-sharded 1M x 2048
database followed by the top-100 re-search, and a full DBA pass over the 1M rows (each rank augments its own rows).
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_config5_multi.py [--dba-rows R]
Prints one JSON line on rank 0 (device time, max over ranks)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
import torch.distributed as dist
from cirtorch_b200 import parallel as P, search as S, rerank as R

ap = argparse.ArgumentParser()
ap.add_argument("--dba-rows", type=int, default=0, help="augment only this many rows per rank (0 = the whole shard)")
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, D, Q = 1_000_000, 2048, 10_000
lo, hi = P.shard_bounds(N, world, rank)
g = torch.Generator(device=dev).manual_seed(100 + rank)
rows = torch.empty((hi - lo, D), device=dev)
for a in range(0, hi - lo, 62_500):
    b = min(hi - lo, a + 62_500)
    blk = torch.randn((b - a, D), device=dev, generator=g)
    rows[a:b] = blk / blk.norm(dim=1, keepdim=True)
sh = P.ShardedIndex(rows, N, mode="bf16")
gq = torch.Generator(device=dev).manual_seed(7)
q = torch.randn((Q, D), device=dev, generator=gq)
q = q / q.norm(dim=1, keepdim=True)


def timed(fn, n, warm):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def qe():
    q2 = P.alpha_qe_sharded_rows(q, sh, k=10, alpha=3.0)
    return sh.search_rows(q2, 100)

ms_qe = timed(qe, 3, 1)
ms_plain = timed(lambda: sh.search_rows(q, 100), 3, 1)
# DBA: replicate the rows once, then every rank augments its own rows against the whole database
full = P.all_gather_rows(rows, N)
index = S.Index(full, mode="bf16")
n_aug = (hi - lo) if args.dba_rows <= 0 else min(args.dba_rows, hi - lo)
ms_dba = timed(lambda: R.dba_rows(full, k=10, alpha=3.0, index=index, row_begin=lo, row_end=lo + n_aug, chunk=10_000), 1, 0)
if rank == 0:
    print(json.dumps({"config": "alpha-QE k=10 alpha=3 + top-100 re-search, 10k queries x 1M x 2048 row-sharded; DBA over the 1M rows "
                                "(BASELINE configs[4])", "n_gpus": world, "alpha_qe_ms": ms_qe, "alpha_qe_queries_per_s": Q / (ms_qe * 1e-3),
                      "plain_search_ms": ms_plain, "dba_rows_per_rank": n_aug, "dba_ms": ms_dba,
                      "dba_full_pass_s": ms_dba * 1e-3 * ((hi - lo) / n_aug), "dba_rows_per_s": world * n_aug / (ms_dba * 1e-3),
                      "note": "search_rows = bf16 scan + fp32 re-score per shard, NCCL all_gather + merge; alpha-QE = one all_reduce of the "
                              "[Q, D] neighbour sums; DBA = rows replicated by one all_gather, no communication afterwards"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
