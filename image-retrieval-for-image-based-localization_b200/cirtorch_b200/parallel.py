"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box,
gloo in the CPU tests).

Search: the database is sharded by rows, every rank searches its shard with global indices
(``row_offset``), then ONE all_gather of the per-query top-k lists (scores fp32 + idx int32,
Q*k*8 bytes per rank) and a k-way merge on every rank.  Exact: top-k(global) == top-k(union of
shard top-k).  Extraction is data-parallel over contiguous image slices followed by an
all_gather of the descriptor columns (the reference never gathers -- SURVEY.md quirk Q1).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int):
    """Contiguous row range [lo, hi) of shard ``rank``."""
    return (rank * n) // world_size, ((rank + 1) * n) // world_size


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def topk_exchange_buffer(Q: int, k: int, device):
    """One 32-bit buffer [2, Q, k] whose halves are the local (scores fp32, idx int32) lists: the search writes
    into it directly and it travels in ONE collective."""
    both = torch.empty((2, Q, k), dtype=torch.int32, device=device)
    return both, (both[0].view(torch.float32), both[1])


def gather_topk(scores: torch.Tensor, idx: torch.Tensor, group=None, both: torch.Tensor = None):
    """all_gather of per-shard lists [Q, k] -> ([G, Q, k] scores, [G, Q, k] idx), shard order = rank order.

    Scores (fp32) and indices (int32) travel in ONE collective as a [2, Q, k] 32-bit buffer (pass ``both`` from
    :func:`topk_exchange_buffer` to skip the packing copy); the results are strided views of the gathered buffer."""
    _, ws = world(group)
    if ws == 1:
        return scores[None], idx[None]
    Q, k = scores.shape
    if both is None:
        both = torch.empty((2, Q, k), dtype=torch.int32, device=scores.device)
        both[0] = scores.contiguous().view(torch.int32)
        both[1] = idx.to(torch.int32)
    out = torch.empty((ws * 2, Q, k), dtype=torch.int32, device=scores.device)
    dist.all_gather_into_tensor(out, both, group=group)                       # concatenation along dim 0
    out = out.view(ws, 2, Q, k)
    return out[:, 0].view(torch.float32), out[:, 1]


class PeerExchange:
    """NVLink exchange buffers for the fused search + all-gather (cir_search_topk_exchange).

    Two symmetric-memory buffers per rank, used alternately: [G][2][Q][k] int32 words of lists followed by Q arrival
    counters.  The final selection kernel of every rank stores its lists straight into slot `rank` of all peers' buffers.
    Small query batches (Q <= the number of SMs): the same kernel then signals the peers' counters, waits on its own and
    merges (cir_search_topk_exchange_merge) -- no barrier, no merge launch.  Larger batches: one device-side barrier, then
    every rank merges its own buffer.  Double buffering makes one synchronisation per search sufficient (a rank can only
    get to search i+1's exchange after every peer's list of search i reached it, i.e. after every peer finished search i-1)."""

    def __init__(self, Q: int, k: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world_size = world(group)
        self.Q, self.k = Q, k
        self.bufs, self.handles, self.ptrs = [], [], []
        import ctypes as C
        words = self.world_size * 2 * Q * k
        self.uses = [0, 0]
        self._flat = []
        for _ in range(2):
            flat = symm_mem.empty((words + Q,), dtype=torch.int32, device=device)
            flat[words:].zero_()                                           # the arrival counters start at 0 and only grow
            h = symm_mem.rendezvous(flat, self.group)
            self._flat.append(flat)
            self.bufs.append(flat[:words].view(self.world_size, 2, Q, k))
            self.handles.append(h)
            self.ptrs.append((C.c_void_p * self.world_size)(*[int(a) for a in h.buffer_ptrs]))
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)        # no peer signals a counter before everybody has zeroed theirs
        self.turn = 0

    def next(self, fused_merge: bool):
        """The buffer of this search: (lists view, handle, peer pointers, arrival count its counters reach when every rank
        has delivered -- only searches with the fused merge signal them)."""
        i = self.turn
        self.turn ^= 1
        if fused_merge:
            self.uses[i] += 1
        return self.bufs[i], self.handles[i], self.ptrs[i], self.uses[i] * self.world_size


class ShardedIndex:
    """Row-sharded database: this rank holds rows [lo, hi) of a global N-row database."""

    def __init__(self, local_rows: torch.Tensor, n_global: int, group=None, mode="bf16", keep_fp32=True):
        from .search import Index
        self._exchange = {}
        self.group = group
        self.rank, self.world_size = world(group)
        self.lo, self.hi = shard_bounds(n_global, self.world_size, self.rank)
        if local_rows.shape[0] != self.hi - self.lo:
            raise ValueError("rank %d expects %d rows, got %d" % (self.rank, self.hi - self.lo, local_rows.shape[0]))
        self.n_global = n_global
        self.index = Index(local_rows, mode=mode, keep_fp32=keep_fp32, row_offset=self.lo)

    def search_local(self, q_rows, k, rescore=None):
        return self.index.search_rows(q_rows, k, rescore=rescore)

    def search_rows(self, q_rows, k, rescore=None):
        """Global top-k on every rank: (scores [Q, k], idx [Q, k] global row ids).

        With re-scoring (the default for a bf16 index that keeps its fp32 rows) the result is exactly what one GPU holding the
        whole database returns: the shards' bf16 lists of k + 28 candidates are merged into the GLOBAL bf16 candidate list,
        every rank re-scores in fp32 only the candidates that live in its shard (1/G of the gather traffic instead of a full
        list per shard), and a second all_gather + merge orders them."""
        from .search import MAX_K, merge_topk, rescore_pad, rescore_rows
        if rescore is None:
            rescore = self.index.mode == "bf16" and self.index.rows32 is not None
        if self.world_size == 1:
            return self.index.search_rows(q_rows, k, rescore=rescore)
        if not rescore:
            s, i = self.index.search_rows(q_rows, k, rescore=False)
            s_all, i_all = gather_topk(s, i, self.group)
            return merge_topk(s_all, i_all, k)      # strided views of the gathered buffer, merged in place
        kc = min(rescore_pad(k), max(k, self.n_global), MAX_K)
        s, i = self.index.search_rows(q_rows, kc, rescore=False)
        s_all, i_all = gather_topk(s, i, self.group)
        _, cand = merge_topk(s_all, i_all, kc)       # the global bf16 top-kc, identical on every rank
        s2, i2 = rescore_rows(q_rows, self.index.rows32, cand, kc, idx_offset=self.lo)    # rows of other shards are skipped
        s_all, i_all = gather_topk(s2, i2, self.group)
        return merge_topk(s_all, i_all, k)

    def search_packed_p2p(self, qp, k, fused_merge=None):
        """Global top-k of packed queries with the exchange fused into the search: every rank's selection kernel
        writes its lists into all peers' buffers over NVLink.  No NCCL call.  Up to one query per SM the merge is fused in
        too (arrival counters in the exchange buffers; ``fused_merge=False`` forces the other path); larger batches:
        one device barrier, then a merge launch."""
        import ctypes as C
        from . import _lib
        from .search import merge_topk
        lib = _lib.load()
        Q, Kd = qp.shape
        key = (Q, k)
        ex = self._exchange.get(key)
        if ex is None:
            ex = self._exchange[key] = PeerExchange(Q, k, qp.device, self.group)
        if fused_merge is None:
            fused_merge = Q <= torch.cuda.get_device_properties(qp.device).multi_processor_count and self.world_size * k <= 8192
        buf, handle, ptrs, arrivals = ex.next(fused_merge)
        need = C.c_size_t(0)
        _lib.check(lib.cir_search_workspace_bytes(Q, self.index.N, Kd, k, C.byref(need)), "cir_search_workspace_bytes")
        ws = _lib.workspace(qp.device, need.value, "search")
        if fused_merge:
            scores = torch.empty((Q, k), dtype=torch.float32, device=qp.device)
            idx = torch.empty((Q, k), dtype=torch.int32, device=qp.device)
            rc = lib.cir_search_topk_exchange_merge(_lib.ptr(qp), Q, _lib.ptr(self.index.packed), self.index.N, Kd, k, self.lo,
                                                    ptrs, self.world_size, self.rank, arrivals, _lib.ptr(scores), _lib.ptr(idx),
                                                    _lib.ptr(ws), ws.numel(), 0, _lib.stream_of(qp))
            _lib.check(rc, "cir_search_topk_exchange_merge")
            return scores, idx
        rc = lib.cir_search_topk_exchange(_lib.ptr(qp), Q, _lib.ptr(self.index.packed), self.index.N, Kd, k, self.lo,
                                          ptrs, self.world_size, self.rank, _lib.ptr(ws), ws.numel(), 0,
                                          _lib.stream_of(qp))
        _lib.check(rc, "cir_search_topk_exchange")
        handle.barrier()                                    # device-side, on the current stream
        return merge_topk(buf[:, 0].view(torch.float32), buf[:, 1], k)


def all_gather_rows(local_rows: torch.Tensor, n_global: int, group=None) -> torch.Tensor:
    """Row shards (rank r holds rows shard_bounds(n_global, W, r)) -> the full [n_global, D] matrix on every rank."""
    rank, ws = world(group)
    if ws == 1:
        return local_rows
    D = local_rows.shape[1]
    sizes = [shard_bounds(n_global, ws, r)[1] - shard_bounds(n_global, ws, r)[0] for r in range(ws)]
    mx = max(sizes)
    pad = local_rows if local_rows.shape[0] == mx else torch.cat(
        [local_rows, local_rows.new_zeros((mx - local_rows.shape[0], D))], 0)
    allv = torch.empty((ws * mx, D), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(allv, pad.contiguous(), group=group)
    if all(sz == mx for sz in sizes):
        return allv
    allv = allv.view(ws, mx, D)
    return torch.cat([allv[r, :sizes[r]] for r in range(ws)], 0)


def replicate_host_rows(host_rows: torch.Tensor, device, group=None) -> torch.Tensor:
    """Host rows (pinned, the same on every rank) -> device rows on every rank, moving each row over PCIe ONCE per node:
    rank r copies its 1/G slice host -> device, one all_gather over NVLink rebuilds the matrix."""
    rank, ws = world(group)
    if ws == 1:
        return host_rows.to(device, non_blocking=True)
    n = host_rows.shape[0]
    lo, hi = shard_bounds(n, ws, rank)
    return all_gather_rows(host_rows[lo:hi].to(device, non_blocking=True), n, group)


def alpha_qe_sharded_rows(q_rows, sharded: "ShardedIndex", k=10, alpha=3.0):
    """alpha-QE against a ROW-SHARDED database (BASELINE.json config 5): one sharded search for the global top-k lists, every
    rank sums the neighbours that live in its shard (cir_qe_aggregate without the query and without the L2N), ONE
    all_reduce of the [Q, D] parts, then q' = L2N(q + sum).  Every rank returns all expanded queries."""
    from .rerank import qe_aggregate_rows
    from . import functional as LF
    if sharded.index.rows32 is None:
        raise ValueError("alpha-QE needs the fp32 database rows (keep_fp32=True)")
    s, i = sharded.search_rows(q_rows, k)
    mine = (i >= sharded.lo) & (i < sharded.hi)
    i_loc = torch.where(mine, i - sharded.lo, torch.full_like(i, -1))
    part = qe_aggregate_rows(None, sharded.index.rows32, i_loc, s, k, alpha, normalize=False)
    if sharded.world_size > 1:
        dist.all_reduce(part, group=sharded.group)
    return LF.l2n(q_rows + part)


def dba_sharded_rows(local_rows, n_global, k=10, alpha=3.0, group=None, mode="bf16", chunk=16384):
    """Database-side augmentation of a row-sharded database: the rows are replicated once (one all_gather; 1M x 2048 fp32 =
    8.2 GB + 4.1 GB bf16 per GPU), then every rank augments ITS OWN rows against the whole database with no further
    communication.  Returns this rank's augmented rows [hi - lo, D]."""
    from .rerank import dba_rows
    from .search import Index
    rank, ws = world(group)
    lo, hi = shard_bounds(n_global, ws, rank)
    full = all_gather_rows(local_rows, n_global, group)
    index = Index(full, mode=mode)
    return dba_rows(full, k, alpha, index=index, row_begin=lo, row_end=hi, chunk=chunk)


def extract_vectors_dp(net, images, group=None, **kw):
    """Data-parallel extract_vectors: every rank returns the full D x N matrix (on its device)."""
    from .extract import extract_vectors
    rank, ws = world(group)
    local = extract_vectors(net, images, rank=rank, world_size=ws, **kw)
    if ws == 1:
        return local
    n = len(images)
    D = local.shape[0]
    sizes = [shard_bounds(n, ws, r)[1] - shard_bounds(n, ws, r)[0] for r in range(ws)]
    mx = max(sizes)
    pad = torch.zeros((mx, D), dtype=local.dtype, device=local.device)
    pad[:local.shape[1]] = local.t()
    allv = torch.empty((ws * mx, D), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(allv, pad, group=group)
    allv = allv.view(ws, mx, D)
    return torch.cat([allv[r, :sizes[r]] for r in range(ws)], 0).t()
