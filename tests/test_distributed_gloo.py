"""world_size-2 gloo test of the sharded-search host logic (partition -> per-shard top-k with global
ids -> all_gather -> merge).  No GPU here, so the per-shard top-k and the merge are the CPU oracle's;
what is under test is cirtorch_b200.parallel's partitioning, gather order and index offsets."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cirtorch_b200 import parallel as P
    from oracle import cirtorch_oracle as O
    rs = np.random.RandomState(0)
    N, D, Q, k = 1001, 32, 9, 12
    db = rs.randn(N, D).astype(np.float32)
    q = rs.randn(Q, D).astype(np.float32)
    lo, hi = P.shard_bounds(N, world, rank)
    s, i = O.topk(db[lo:hi].T, q.T, k)                    # k x Q, local ids
    s_loc = torch.from_numpy(s.T.astype(np.float32).copy())
    i_loc = torch.from_numpy((i.T + lo).astype(np.int32).copy())
    s_all, i_all = P.gather_topk(s_loc, i_loc)
    assert s_all.shape == (world, Q, k)
    # merge on the host: sort the union by (score desc, idx asc)
    su = s_all.permute(1, 0, 2).reshape(Q, -1).numpy()
    iu = i_all.permute(1, 0, 2).reshape(Q, -1).numpy()
    order = np.lexsort((iu, -su), axis=1)[:, :k]
    merged = np.take_along_axis(iu, order, 1)
    s_ref, i_ref = O.topk(db.T, q.T, k)
    ok = bool((merged == i_ref.T).all())
    # row shards -> the replicated matrix (ragged: 500 + 501 rows), the plumbing of the sharded DBA / DP extraction
    full = P.all_gather_rows(torch.from_numpy(db[lo:hi].copy()), N)
    ok = ok and tuple(full.shape) == (N, D) and bool((full.numpy() == db).all())
    # host rows replicated with one slice per rank + all_gather (the end-to-end query path of the sharded search)
    rep = P.replicate_host_rows(torch.from_numpy(q.copy()), torch.device("cpu"))
    ok = ok and tuple(rep.shape) == (Q, D) and bool((rep.numpy() == q).all())
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_two_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1
