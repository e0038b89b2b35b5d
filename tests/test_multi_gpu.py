"""Two-GPU NCCL / NVLink tests (skipped on a single-GPU box): the sharded search (NCCL all_gather path and the fused
peer-store exchange) and the sharded alpha-QE / DBA must equal the single-GPU results."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from cirtorch_b200 import parallel as P, search as S
    g = torch.Generator(device="cpu").manual_seed(0)
    N, D, Q, k = 70_001, 256, 300, 50
    db = torch.randn((N, D), generator=g)
    db = (db / db.norm(dim=1, keepdim=True)).to(dev)
    q = torch.randn((Q, D), generator=g)
    q = (q / q.norm(dim=1, keepdim=True)).to(dev)
    s_ref, i_ref = S.search_topk_rows(q, db, k, mode="bf16", rescore=False)
    lo, hi = P.shard_bounds(N, world, rank)
    sh = P.ShardedIndex(db[lo:hi].contiguous(), N, mode="bf16")
    s1, i1 = sh.search_rows(q, k, rescore=False)                      # NCCL all_gather + merge
    ok = bool(torch.equal(i1, i_ref)) and bool(torch.equal(s1, s_ref))
    # with fp32 re-scoring the sharded search returns exactly what one GPU holding the whole database returns
    s_rr, i_rr = S.Index(db, mode="bf16").search_rows(q, k)
    s3, i3 = sh.search_rows(q, k)
    ok = ok and bool(torch.equal(i3, i_rr)) and bool(torch.equal(s3, s_rr))
    qh = q.cpu().pin_memory()
    ok = ok and bool(torch.equal(P.replicate_host_rows(qh, dev), q))
    qp = S.pack_rows(q, "query", "bf16")
    for _ in range(3):                                               # alternates the two exchange buffers
        s2, i2 = sh.search_packed_p2p(qp, k)                         # fused peer-store exchange
        ok = ok and bool(torch.equal(i2, i_ref)) and bool(torch.equal(s2, s_ref))
    # a small batch (one block per query, all resident): exchange AND merge inside the selection kernel -- arrival counters
    # in the exchange buffers instead of a barrier.  One rank is held back on purpose: the other one waits inside its kernel.
    qp70 = S.pack_rows(q[:70].contiguous(), "query", "bf16")
    for it in range(6):
        if it % 2 == rank:
            torch.cuda._sleep(30_000_000)                            # ~15 ms
        s4, i4 = sh.search_packed_p2p(qp70, k, fused_merge=(it != 3))   # one search through the barrier path in between
        ok = ok and bool(torch.equal(i4, i_ref[:70])) and bool(torch.equal(s4, s_ref[:70]))
    # config 5 on shards: alpha-QE with the neighbour sum split over the ranks, DBA with each rank augmenting its rows
    from cirtorch_b200 import rerank as R
    full_index = S.Index(db, mode="bf16")
    q2_ref = R.alpha_qe_rows(q, full_index, k=10, alpha=3.0)
    q2 = P.alpha_qe_sharded_rows(q, sh, k=10, alpha=3.0)
    ok = ok and float((q2 - q2_ref).abs().max()) < 2e-6
    small = db[:6000].contiguous()
    lo2, hi2 = P.shard_bounds(6000, world, rank)
    aug_ref = R.dba_rows(small, k=10, alpha=3.0)
    aug = P.dba_sharded_rows(small[lo2:hi2].contiguous(), 6000, k=10, alpha=3.0)
    ok = ok and aug.shape == (hi2 - lo2, D) and float((aug - aug_ref[lo2:hi2]).abs().max()) < 2e-6
    # data-parallel extraction: every rank extracts its slice of the images, one all_gather of the descriptor rows
    from cirtorch_b200.extract import ImageRetrievalNet, extract_vectors
    from cirtorch_b200.modules.heads.global_head import globalHead

    class Body(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 64, 3, padding=1)

        def forward(self, img):
            return {"mod5": torch.relu(self.conv(img))}

    torch.manual_seed(11)
    net = ImageRetrievalNet(Body(), globalHead(pooling={"name": "GeM", "params": {"p": 3, "eps": 1e-6}},
                                               normal={"name": "L2N", "params": {}}, dim=64)).to(dev).eval()
    imgs = [torch.randn(3, 40, 32, generator=g) for _ in range(7)]
    v_ref = extract_vectors(net, imgs, 40, None, batch_size=2, device=dev)            # D x 7 on the CPU, like upstream
    v_dp = P.extract_vectors_dp(net, imgs, image_size=40, transform=None, batch_size=2, device=dev)
    ok = ok and tuple(v_dp.shape) == (64, 7) and float((v_dp.cpu() - v_ref).abs().max()) < 1e-6
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    if rank == 0:
        out.put(int(flag))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_search_nccl_and_p2p():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1
