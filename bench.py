#!/usr/bin/env python
"""Benchmark of the global-descriptor retrieval hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-search]

Headline line (`metric` = descriptors/s): one step = one pass of the fused GeM+L2N+whiten+L2N tail
over a batch of 64 x 2048 x 32 x 32 fp32 layer-4 maps (BASELINE.json configs[1]), whitening
2048 -> 2048, synthetic data, random-init weights.  N > 1 (torchrun): data parallel, one batch per
rank per step (weak scaling).  The same JSON line carries the other BASELINE.json configs:
"tail_p27" (the same step with a non-integer GeM exponent), "search" (configs[3]: queries/s of the
1M x 2048 top-100 search, database row-sharded over the N GPUs, strong scaling; i.i.d., clustered
and cluster-sorted rows), "mining" (configs[2]), "alpha_qe" / "dba" (configs[4], the DBA pass over
all 1M rows), "extract_rank" (configs[0]) -- each with its roofline -- and, at N > 1, "parity":
sharded results checked against single-rank / brute-force results outside the timed regions
(exit code 1 if any check fails).

`--impl reference` times the reference's CPU implementation of the same step (the oracle port of
globalHead.forward, all host threads) -- on rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

B, C, H, W, DOUT = 64, 2048, 32, 32, 2048
TAIL_BYTES = B * C * H * W * 4 + DOUT * C * 4 + DOUT * 4 + B * DOUT * 4          # 554,180,608
TAIL_NCU_TRAFFIC = 553_788_672 + 3_611_392      # dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full (profiles/r2_tail_p3.txt)
DB_N, DB_D, TOPK = 1_000_000, 2048, 100
WORKLOAD = "fused tail: batch 64x2048x32x32 fp32 maps -> GeM(p=3)+L2N+whiten 2048->2048+L2N"
METRIC = "descriptors/s (GeM+whiten tail) & queries/s vs 1M×2048 DB at 1/2/4/8 B200, %roofline"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


_SAMPLER_CODE = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
out = open(sys.argv[2], "w", buffering=1)
out.write("max %d\n" % nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
while True:
    try:
        mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        out.write("%.6f %d %d\n" % (time.time(), mhz, r))
    except Exception:
        pass
    time.sleep(0.0005)
"""


class ClockSampler:
    """SM clock / throttle reasons sampled by NVML in a SEPARATE process (no GIL contention with the launch
    loop) while the timed region runs; samples are matched to the region by wall-clock time stamps."""

    def __init__(self, index):
        import subprocess
        import tempfile
        self.path = tempfile.mktemp(prefix="cir_clocks_")
        self.t0 = self.t1 = None
        try:
            phys = index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except ValueError:
                    phys = index
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_CODE, str(phys), self.path],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def start(self):
        # give the child time to initialise NVML
        t_end = time.time() + 5.0
        while self.proc and time.time() < t_end and not (os.path.exists(self.path) and os.path.getsize(self.path) > 20):
            time.sleep(0.01)

    @property
    def mark(self):
        return 0

    @mark.setter
    def mark(self, v):
        if v == 1:
            self.t0 = time.time()
        elif v == 2:
            self.t1 = time.time()

    def finish(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        time.sleep(0.005)
        self.proc.kill()
        self.proc.wait()
        max_mhz, samples = None, []
        try:
            for line in open(self.path):
                f = line.split()
                if f and f[0] == "max":
                    max_mhz = int(f[1])
                elif len(f) == 3:
                    samples.append((float(f[0]), int(f[1]), int(f[2])))
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": max_mhz, "reasons": ["nvml unavailable"]}
        timed = [s for s in samples if self.t0 is not None and self.t0 <= s[0] <= self.t1]
        inside = len(timed)
        if not timed:   # region shorter than the sampling period: take the samples closest to it
            timed = sorted(samples, key=lambda s: abs(s[0] - (self.t0 or s[0])))[:3]
        names = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
                 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
                 0x100: "display_clock_setting"}
        bits = 0
        for s in timed:
            bits |= s[2]
        mhz = sorted(s[1] for s in timed)
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": max_mhz, "reasons": [n for b, n in names.items() if bits & b],
                "samples": inside}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


LAST_PER_RANK_MS = []          # device time of the last timed region on every rank (the reported time is their maximum)


def timed_region(fn, steps, warmup, world, sampler=None):
    """W untimed steps, then exactly K timed steps between barrier + synchronize; max over ranks (device time).

    One more untimed step runs AFTER the barrier: the first launch into a drained queue costs ~100 us (a cooperative
    launch more), which is not part of the steady state a K-step loop measures."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.mark = 1
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if sampler:
        sampler.mark = 2
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        LAST_PER_RANK_MS[:] = [float(v) for v in every]
        ms = max(LAST_PER_RANK_MS)
    else:
        LAST_PER_RANK_MS[:] = [ms]
    return ms


def make_head(dev, p=3.0):
    from cirtorch_b200.modules.heads.global_head import globalHead
    torch.manual_seed(0)
    head = globalHead(pooling={"name": "GeM", "params": {"p": p, "eps": 1e-6}},
                      normal={"name": "L2N", "params": {}}, dim=DOUT)
    return head.to(dev).eval()


def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


# ------------------------------------------------------------------------------------------ CPU baselines (oracle port)
def _cpu_tail_inputs():
    torch.manual_seed(0)
    x = torch.relu(torch.randn(B, C, H, W))
    Wt = torch.empty(DOUT, C)
    torch.nn.init.xavier_normal_(Wt, 0.1)
    return x, Wt, torch.zeros(DOUT)


def cpu_tail_baseline(budget_s=10.0, max_iters=30):
    """The oracle port of globalHead.forward on the host cores, batch 64 x 2048 x 32 x 32."""
    from oracle import cirtorch_oracle as O
    host_threads()
    x, Wt, b = _cpu_tail_inputs()
    O.head_forward(x, 3.0, 1e-6, Wt, b)              # warm-up
    t0 = time.perf_counter()
    it = 0
    while it < max_iters and (time.perf_counter() - t0 < budget_s or it < 2):
        O.head_forward(x, 3.0, 1e-6, Wt, b)
        it += 1
    dt = time.perf_counter() - t0
    return {"value": B * it / dt, "unit": "descriptors/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d batches of %dx%dx%dx%d fp32 through oracle.head_forward (torch CPU eager, %d threads)"
                      % (it, B, C, H, W, torch.get_num_threads())}


def cpu_search_baseline(n_rows=100_000, q=70, budget_s=8.0):
    """np.dot + np.argsort (scripts/train_globalF.py:733-734) on a 100k-row slice of the database."""
    from oracle import cirtorch_oracle as O
    rs = np.random.RandomState(0)
    db = rs.randn(DB_D, n_rows).astype(np.float32)
    qs = rs.randn(DB_D, q).astype(np.float32)
    t0 = time.perf_counter()
    it = 0
    while it < 3 and (time.perf_counter() - t0 < budget_s or it < 1):
        O.rank(db, qs)
        it += 1
    dt = (time.perf_counter() - t0) / it
    return {"value": q / (dt * DB_N / n_rows), "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "np.dot + np.argsort, %d queries x %d rows x %d, time scaled linearly to 1M rows" % (q, n_rows, DB_D)}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    host_threads()
    from oracle import cirtorch_oracle as O
    x, Wt, b = _cpu_tail_inputs()
    for _ in range(args.warmup):
        O.head_forward(x, 3.0, 1e-6, Wt, b)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.head_forward(x, 3.0, 1e-6, Wt, b)
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    threads = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "descriptors/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "note": "reference CPU path = oracle port of globalHead.forward (global_head.py:52-67), torch CPU eager, "
                           "one 64-image batch per step on rank 0"},
        "cpu_baseline": {"value": val, "unit": "descriptors/s", "cores": threads, "kind": "port",
                         "sample": "%d steps of one 64-image batch, %d host threads" % (args.steps, threads)},
        "e2e": {"value": val, "unit": "descriptors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ synthetic databases
N_CENTRES, SIGMA = 10_000, 0.5


def _unit(x):
    return x / x.norm(dim=1, keepdim=True)


def fill_database(kind, rows32, lo, hi, dev, rank):
    """Rows [lo, hi) of the global 1M x 2048 database, written into rows32 (fp32, unit norm).

    "iid"            independent Gaussian directions (the round-1 workload)
    "clustered"      SURVEY.md 8(d): 10k Gaussian centres + sigma-noise, rows in random order
    "cluster_sorted" the same distribution stored centre by centre (a database written scene by scene)"""
    n_loc = hi - lo
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    centres = which = None
    if kind != "iid":
        gc = torch.Generator(device=dev).manual_seed(4321)              # the same centres / assignment on every rank
        centres = _unit(torch.randn((N_CENTRES, DB_D), device=dev, generator=gc))
        which = torch.randint(0, N_CENTRES, (DB_N,), device=dev, generator=gc)
        if kind == "cluster_sorted":
            which = torch.sort(which).values
        which = which[lo:hi]
    for a in range(0, n_loc, 62_500):
        b = min(n_loc, a + 62_500)
        blk = torch.randn((b - a, DB_D), device=dev, generator=g)
        if centres is not None:
            blk = centres[which[a:b]] + blk * (SIGMA / DB_D ** 0.5)
        rows32[a:b] = _unit(blk)
    return centres


def make_queries(kind, Q, dev, centres):
    gq = torch.Generator(device=dev).manual_seed(99)
    q = torch.randn((Q, DB_D), device=dev, generator=gq)
    if kind != "iid":
        pick = torch.randint(0, N_CENTRES, (Q,), device=dev, generator=gq)
        q = centres[pick] + q * (SIGMA / DB_D ** 0.5)
    return _unit(q)


def _local_index(rows32, dbp, lo):
    from cirtorch_b200 import search as S
    idx = S.Index.__new__(S.Index)
    idx.mode, idx.N, idx.D, idx.row_offset, idx.packed, idx.rows32, idx.labels = "bf16", rows32.shape[0], DB_D, lo, dbp, rows32, None
    return idx


def _sharded(rows32, dbp, lo, hi, rank, world):
    from cirtorch_b200 import parallel as P
    sh = P.ShardedIndex.__new__(P.ShardedIndex)
    sh._exchange, sh.group, sh.rank, sh.world_size, sh.lo, sh.hi, sh.n_global = {}, None, rank, world, lo, hi, DB_N
    sh.index = _local_index(rows32, dbp, lo)
    return sh


def bench_search(dev, rank, world, pk, steps, warmup, state):
    """1M x 2048 database row-sharded over the ranks, top-100: 10k-query batch (tensor-bound) and 70 queries (HBM-bound)
    on i.i.d. rows, plus the 10k-query batch on clustered rows in random and in centre-sorted order."""
    import torch.distributed as dist
    from cirtorch_b200 import search as S, parallel as P, _lib
    lo, hi = P.shard_bounds(DB_N, world, rank)
    n_loc = hi - lo
    dbp = torch.empty((n_loc, DB_D), dtype=torch.bfloat16, device=dev)
    rows32 = torch.empty((n_loc, DB_D), dtype=torch.float32, device=dev)
    out = {}
    p2p_ok = world > 1
    for kind, cases in (("iid", (("q10k", 10_000), ("q70", 70))), ("clustered", (("q10k_clustered", 10_000),)),
                        ("cluster_sorted", (("q10k_cluster_sorted", 10_000),))):
        centres = fill_database(kind, rows32, lo, hi, dev, rank)
        S.pack_rows(rows32, "db", "bf16", out=dbp)
        sh = _sharded(rows32, dbp, lo, hi, rank, world)
        for name, Q in cases:
            q32 = make_queries(kind, Q, dev, centres)
            qp = S.pack_rows(q32, "query", "bf16")
            both, views = P.topk_exchange_buffer(Q, TOPK, dev)
            exchange = "none" if world == 1 else "nccl all_gather"
            use_p2p = False
            if p2p_ok:
                # fused exchange: the selection kernel stores the lists into every peer's buffer over NVLink
                try:
                    sh.search_packed_p2p(qp, TOPK)
                    torch.cuda.synchronize()
                    use_p2p = True
                except Exception as e:  # noqa: BLE001  (no symmetric memory on this box: NCCL all_gather instead)
                    exchange = "nccl all_gather (%s)" % type(e).__name__
                flag = torch.tensor([1 if use_p2p else 0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                use_p2p = p2p_ok = bool(int(flag))
                if use_p2p:
                    exchange = "peer stores over NVLink fused into the selection kernel (symmetric memory)"

            def step_local():
                return S.search_packed(qp, dbp, TOPK, idx_offset=lo, out=views)

            def step():
                if use_p2p:
                    return sh.search_packed_p2p(qp, TOPK)
                s, i = step_local()
                if world > 1:
                    s_all, i_all = P.gather_topk(s, i, both=both)
                    s, i = S.merge_topk(s_all, i_all, TOPK)
                return s, i

            k_steps = max(3, steps // 10) if Q >= 1000 else max(5, steps)
            _lib.launch_count(reset=True)
            ms = timed_region(step, k_steps, max(3, warmup), world) / k_steps
            per_rank = [round(v / k_steps, 4) for v in LAST_PER_RANK_MS]
            launches = _lib.launch_count()
            ms_kernel = timed_region(step_local, k_steps, 3, world) / k_steps
            flops = 2.0 * Q * n_loc * DB_D
            res = {"queries_per_s": Q / (ms * 1e-3), "ms_per_search": ms, "ms_local_kernels": ms_kernel, "rows": kind,
                   "steps": k_steps, "launches_per_search": launches // (k_steps + max(3, warmup) + 1), "Q": Q, "N": DB_N, "k": TOPK,
                   "exchange": exchange, "per_rank_ms_per_search": per_rank}
            if Q >= 1000:
                tf = flops / (ms_kernel * 1e-3) / 1e12
                res["roofline"] = {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                   "frac": tf / pk["bf16_tflops_sustained"], "frac_of_burst": tf / pk["bf16_tflops"], "traffic": None,
                                   "note": "per GPU: 2*Q*N_local*D flop / local search time (threshold pre-pass + GEMM + top-k "
                                           "select); peak = " + pk["src"] + " sustained cuBLAS bf16"}
            else:
                gb = (n_loc * DB_D * 2 + Q * DB_D * 2 + Q * TOPK * 8) / (ms_kernel * 1e-3) / 1e9
                res["roofline"] = {"bound": "hbm", "achieved": gb, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                   "frac": gb / pk["hbm_gbs"], "traffic": None,
                                   "note": "per GPU: bf16 shard scan bytes / local search time; peak = " + pk["src"] + " copy bandwidth"}
            # end to end through the public API with host buffers: H2D queries, pack, scan, fp32 re-score, D2H lists
            if name == "q10k":
                q_host = q32.cpu().pin_memory()
                s_host = torch.empty((Q, TOPK), dtype=torch.float32).pin_memory()
                i_host = torch.empty((Q, TOPK), dtype=torch.int32).pin_memory()

                def step_e2e():
                    # every rank needs all queries: each copies 1/G of them over PCIe, one all_gather over NVLink
                    qd = P.replicate_host_rows(q_host, dev)
                    s, i = sh.search_rows(qd, TOPK)           # bf16 scan, global top-128, exact fp32 re-score, merge
                    s_host.copy_(s, non_blocking=True)
                    i_host.copy_(i, non_blocking=True)

                ms_e = timed_region(step_e2e, k_steps, 3, world) / k_steps
                res["e2e"] = {"value": Q / (ms_e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": Q * DB_D * 4,
                              "d2h_bytes_per_step": world * Q * TOPK * 8,
                              "includes": "whole job: every query row crosses PCIe once (1/G per rank + all_gather over NVLink), every "
                                          "rank reads the final lists back; fp32 re-score of 128 candidates"}
            if name == "q10k" and world > 1:
                state["parity"].update(parity_search(sh, q32, qp, dev, world, use_p2p))
            if name == "q70" and use_p2p:
                # the small batch goes through the fused exchange + merge (arrival counters, no barrier): against NCCL + merge
                s_l, i_l = step_local()
                s_all, i_all = P.gather_topk(s_l.clone(), i_l.clone())
                s_n, i_n = S.merge_topk(s_all, i_all, TOPK)
                ok = True
                for _ in range(4):
                    s_p, i_p = sh.search_packed_p2p(qp, TOPK)
                    ok = ok and bool(torch.equal(i_p, i_n)) and bool(torch.equal(s_p, s_n))
                state["parity"]["search_q70_fused_merge_equals_nccl"] = _all_true(ok, dev)
                res["exchange"] = "peer stores + arrival counters + merge, all inside the selection kernel (symmetric memory; no barrier)"
            out[name] = res
            del q32, qp
        if kind == "iid":
            out.update(bench_rerank(sh, rows32, dev, rank, world, pk, state))
    del dbp, rows32
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------ config 5: alpha-QE, DBA
def bench_rerank(sh, rows32, dev, rank, world, pk, state):
    """BASELINE.json configs[4]: alpha-QE (k=10, alpha=3) of 10k queries + top-100 re-search over the row-sharded 1M x 2048
    database, and one FULL database-side augmentation pass over the 1M rows (every rank augments its own rows)."""
    from cirtorch_b200 import parallel as P, rerank as R, search as S
    Q = 10_000
    q = make_queries("iid", Q, dev, None)

    def qe():
        q2 = P.alpha_qe_sharded_rows(q, sh, k=10, alpha=3.0)
        return sh.search_rows(q2, TOPK)

    ms_qe = timed_region(qe, 3, 2, world) / 3
    n_loc = rows32.shape[0]
    tf = 2 * 2.0 * Q * n_loc * DB_D / (ms_qe * 1e-3) / 1e12
    res = {"alpha_qe": {"ms": ms_qe, "queries_per_s": Q / (ms_qe * 1e-3), "Q": Q, "N": DB_N, "k": 10, "alpha": 3.0,
                        "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                     "frac": tf / pk["bf16_tflops_sustained"], "traffic": None,
                                     "note": "per GPU: 2 searches x 2*Q*N_local*D flop / time of (sharded top-10 search with fp32 "
                                             "re-score, neighbour sums + one all_reduce, L2N, top-100 re-search with fp32 re-score)"}}}
    # DBA: the rows are replicated once (all_gather), then every rank augments ITS rows against the whole database
    full = P.all_gather_rows(rows32, DB_N)
    index = S.Index(full, mode="bf16")
    lo, hi = sh.lo, sh.hi
    R.dba_rows(full, k=10, alpha=3.0, index=index, row_begin=lo, row_end=lo + 4096, chunk=4096)        # warm-up slice
    holder = {}

    def dba():
        holder["aug"] = R.dba_rows(full, k=10, alpha=3.0, index=index, row_begin=lo, row_end=hi, chunk=16384)

    ms_dba = timed_region(dba, 1, 0, world)
    tf = 2.0 * n_loc * DB_N * DB_D / (ms_dba * 1e-3) / 1e12
    res["dba"] = {"ms": ms_dba, "rows_per_s": DB_N / (ms_dba * 1e-3), "rows_total": DB_N, "rows_per_gpu": n_loc, "k": 10, "alpha": 3.0,
                  "full_pass": True,
                  "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                               "frac": tf / pk["bf16_tflops_sustained"], "traffic": None,
                               "note": "per GPU: 2*N_local*N*D flop / time of the whole pass over this rank's rows (top-11 search with "
                                       "fp32 re-score in 16k-row chunks + neighbour aggregation), after one untimed pass"}}
    if world > 1:
        state["parity"].update(parity_rerank(rows32, dev, rank, world))
    del full, index, holder
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------ config 3: mining
def mining_case(dev=None):
    """BASELINE.json configs[2] / SURVEY.md 8(d): 2,000 queries vs a 20,000-image pool, clusters in [0, 700), nnum 5."""
    rs = np.random.RandomState(11)
    n_images, n_cl, Qm, Pm = 91_642, 700, 2000, 20_000
    clusters = rs.randint(0, n_cl, size=n_images)
    idxs2images = rs.permutation(n_images)[:Pm]
    query_indices = rs.permutation(n_images)[:Qm]
    g = torch.Generator().manual_seed(11)
    centres = torch.randn((n_cl, DB_D), generator=g)
    def rows(ids):
        v = centres[torch.from_numpy(clusters[ids])] * 0.6 + torch.randn((len(ids), DB_D), generator=g)
        return v / v.norm(dim=1, keepdim=True)
    return rows(query_indices), rows(idxs2images), clusters, query_indices, idxs2images


def bench_mining(dev, world, pk, with_cpu):
    from cirtorch_b200 import mining as M, _lib
    q, pool, clusters, qidx, i2i = mining_case()
    qd, pd = q.to(dev), pool.to(dev)
    qc = torch.from_numpy(clusters[qidx]).to(dev, torch.int32)
    pc = torch.from_numpy(clusters[i2i]).to(dev, torch.int32)
    _lib.launch_count(reset=True)
    n = 10
    ms = timed_region(lambda: M.mine_hard_negatives_rows(qd, pd, qc, pc, 5), n, 3, world) / n
    launches = _lib.launch_count() // (n + 4)
    ms_bf16 = timed_region(lambda: M.mine_hard_negatives_rows(qd, pd, qc, pc, 5, mode="bf16"), n, 3, world) / n
    same = bool(torch.equal(M.mine_hard_negatives_rows(qd, pd, qc, pc, 5, mode="bf16")[0], M.mine_hard_negatives_rows(qd, pd, qc, pc, 5)[0]))
    flop = 2.0 * q.shape[0] * pool.shape[0] * DB_D
    tf = flop / (ms * 1e-3) / 1e12
    gb = (q.numel() + pool.numel()) * 4 / (ms * 1e-3) / 1e9
    res = {"ms": ms, "queries_per_s": q.shape[0] / (ms * 1e-3), "Q": q.shape[0], "pool": pool.shape[0], "nnum": 5,
           "launches": launches, "ms_mode_bf16": ms_bf16, "mode_bf16_same_sets": same,
           "mode_note": "default mode bf16x3 (scan error ~1e-5, 80 candidates); mode bf16 = a third of the scan flops, 128 candidates, "
                        "margin from the rigorous bf16 rounding bound",
           "scaling": "replicas (every rank mines the same epoch; SURVEY.md 8e: multi-GPU optional)",
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops"],
                        "hbm_frac": gb / pk["hbm_gbs"], "traffic": None,
                        "note": "latency-bound: 2*Q*P*D = 1.64e11 algorithmic flop (executed 3x as bf16x3) and 180 MB of fp32 inputs over "
                                "the whole step (pack, label-masked pre-pass, search, select, fp32 re-score, greedy walk)"}}
    if with_cpu:
        from oracle import cirtorch_oracle as O
        host_threads()
        t0 = time.perf_counter()
        ref_neg, _ = O.mine_hard_negatives(q.t().contiguous(), pool.t().contiguous(), clusters.tolist(), qidx.tolist(), i2i, 5)
        dt = time.perf_counter() - t0
        neg, _ = M.mine_hard_negatives(qd.t(), pd.t(), clusters, qidx, i2i, 5)
        res["cpu_baseline"] = {"value": q.shape[0] / dt, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "the full epoch once: torch.mm + torch.sort + greedy loop (tuples_dataset.py:317-345) on CPU tensors"}
        res["sets_identical_to_cpu_port"] = bool(neg == ref_neg)
    return res


# ------------------------------------------------------------------------------------------ config 1: extract + rank
def bench_extract_rank(dev, rank, world, with_cpu):
    """BASELINE.json configs[0]: ResNet50-GeM extract_vectors (stock torchvision fp32 backbone, outside the product and the
    roofline) + exhaustive search at the rOxford5k shape.  A 32-image sample per GPU is timed end to end (H2D of the
    images, backbone, fused tail, D2H / all_gather of the descriptors); ranking = the reference's full N x Q argsort."""
    from cirtorch_b200 import parallel as P, search as S
    from cirtorch_b200.extract import resnet50_gem, extract_vectors
    try:          # torchrun exports OMP_NUM_THREADS=1: give every rank its share of the host cores for the image staging
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0)) // world))
    except AttributeError:
        pass
    torch.manual_seed(0)
    net = resnet50_gem().to(dev).eval()
    n_img = 32 * world
    base = torch.randn((32, 3, 1024, 1024), generator=torch.Generator().manual_seed(rank))
    imgs = [base[i % 32] for i in range(n_img)]          # rank r only reads its own slice [32 r, 32 r + 32)
    t = None
    for _ in range(2):                      # first pass = cuDNN autotune / warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        P.extract_vectors_dp(net, imgs, image_size=1024, transform=None, batch_size=8, device=dev)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
    tt = torch.tensor([t], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt)
    g = torch.Generator(device=dev).manual_seed(5)
    dbv = _unit(torch.randn((4993, 2048), device=dev, generator=g))
    qv = _unit(torch.randn((70, 2048), device=dev, generator=g))
    ms_rank = timed_region(lambda: S.rank(dbv.t(), qv.t()), 5, 3, 1) / 5
    ms_topk = timed_region(lambda: S.search_topk(qv.t(), dbv.t(), TOPK), 5, 3, 1) / 5
    res = {"images_per_s": n_img / t, "images_timed": n_img, "extract_5063_images_s_extrapolated": 5063 * t / n_img,
           "full_ranking_ms": ms_rank, "top100_ms": ms_topk, "shape": "70 queries x 4,993 database x 2048",
           "note": "wall clock around extract_vectors_dp (max over ranks): images from host memory, stock torchvision ResNet50 fp32 "
                   "at 1024 px, fused tail, descriptors gathered; 32 images per GPU, scaled linearly to the 5,063 images of the set"}
    if with_cpu:
        from oracle import cirtorch_oracle as O
        host_threads()
        net_cpu = resnet50_gem().eval()
        with torch.no_grad():
            t0 = time.perf_counter()
            f = net_cpu.body(base[:2])
            O.head_forward(f, 3.0, 1e-6, net_cpu.ret_head.whiten.weight, net_cpu.ret_head.whiten.bias)
            cpu_img = (time.perf_counter() - t0) / 2
        t0 = time.perf_counter()
        O.rank(dbv.t().cpu().numpy(), qv.t().cpu().numpy())
        cpu_rank = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": 1.0 / cpu_img, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                               "rank_s": cpu_rank, "sample": "2 images through torchvision ResNet50 + oracle.head_forward on the host; "
                                                             "np.dot + np.argsort of 70 x 4,993"}
    del net, imgs, base
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------ multi-GPU parity (N > 1)
def _all_true(flag, dev):
    import torch.distributed as dist
    t = torch.tensor([1 if flag else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t))


def parity_search(sh, q32, qp, dev, world, use_p2p, n_check=256):
    """Outside the timed regions: the sharded lists (NCCL path, fused peer-store path, re-scored path) against each other and
    against a brute-force fp32 top-100 of the first 256 queries over all shards (tie-window rule: lists identical except
    where the exact scores of the swapped items differ by <= 2 * tol)."""
    import torch.distributed as dist
    from cirtorch_b200 import search as S, parallel as P
    out = {}
    s_l, i_l = S.search_packed(qp, sh.index.packed, TOPK, idx_offset=sh.lo)
    s_all, i_all = P.gather_topk(s_l, i_l)
    s_n, i_n = S.merge_topk(s_all, i_all, TOPK)                       # NCCL all_gather + merge
    if use_p2p:
        ok = True
        for _ in range(2):                                           # both exchange buffers
            s_p, i_p = sh.search_packed_p2p(qp, TOPK)
            ok = ok and bool(torch.equal(i_p, i_n)) and bool(torch.equal(s_p, s_n))
        out["search_p2p_equals_nccl"] = _all_true(ok, dev)
    s_r, i_r = sh.search_rows(q32, TOPK)                              # bf16 scan + fp32 re-score (the e2e path)
    # brute force: exact fp32 scores of this shard, local top-100, gathered, global top-100
    qs = q32[:n_check]
    loc = qs @ sh.index.rows32.t()                                    # [n_check, N_local] fp32 (cuBLAS, checker only)
    bs, bi = torch.topk(loc, TOPK, dim=1)
    bi = bi + sh.lo
    gs = [torch.empty_like(bs) for _ in range(world)]
    gi = [torch.empty_like(bi) for _ in range(world)]
    dist.all_gather(gs, bs)
    dist.all_gather(gi, bi)
    cs, ci = torch.cat(gs, 1), torch.cat(gi, 1)
    order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :TOPK]
    ref_s, ref_i = torch.gather(cs, 1, order), torch.gather(ci, 1, order)

    def exact_scores_of(idx):
        mine = (idx >= sh.lo) & (idx < sh.hi)
        loc_idx = torch.where(mine, idx - sh.lo, torch.zeros_like(idx)).long()
        sc = torch.gather(loc, 1, loc_idx) * mine
        dist.all_reduce(sc)
        return sc

    def window_ok(idx, tol):
        got = exact_scores_of(idx[:n_check])
        bad = idx[:n_check].long() != ref_i
        worst = float(((got - ref_s).abs() * bad).max())
        return worst <= 2 * tol, worst, float(bad.float().mean())

    ok_n, worst_n, frac_n = window_ok(i_n, 5e-4)
    ok_r, worst_r, frac_r = window_ok(i_r, 2e-6)
    out["search_sharded_vs_bruteforce_fp32"] = _all_true(ok_n, dev)
    out["search_rescored_vs_bruteforce_fp32"] = _all_true(ok_r, dev)
    out["search_detail"] = {"queries_checked": n_check, "bf16_worst_swap": worst_n, "bf16_positions_differing": frac_n,
                            "rescored_worst_swap": worst_r, "rescored_positions_differing": frac_r, "tol_bf16": 5e-4, "tol_fp32": 2e-6}
    return out


def parity_rerank(rows32, dev, rank, world, n_sub=65_536, n_q=512):
    """Sharded alpha-QE / DBA against the single-rank result on a 65,536-row sub-database (the first rows of every shard)."""
    from cirtorch_b200 import parallel as P, rerank as R, search as S
    per = n_sub // world
    n_sub = per * world
    local = rows32[:per].contiguous()
    full = P.all_gather_rows(local, n_sub)
    sh = P.ShardedIndex(local, n_sub, mode="bf16")
    q = make_queries("iid", n_q, dev, None)
    q2 = P.alpha_qe_sharded_rows(q, sh, k=10, alpha=3.0)
    q2_ref = R.alpha_qe_rows(q, S.Index(full, mode="bf16"), k=10, alpha=3.0)
    d_qe = float((q2 - q2_ref).abs().max())
    aug = P.dba_sharded_rows(local[:2048].contiguous(), 2048 * world, k=10, alpha=3.0)
    small = P.all_gather_rows(local[:2048].contiguous(), 2048 * world)
    aug_ref = R.dba_rows(small, k=10, alpha=3.0)[rank * 2048:(rank + 1) * 2048]
    d_dba = float((aug - aug_ref).abs().max())
    return {"alpha_qe_sharded_equals_single": _all_true(d_qe < 2e-6, dev), "dba_sharded_equals_single": _all_true(d_dba < 2e-6, dev),
            "rerank_detail": {"alpha_qe_max_abs_diff": d_qe, "dba_max_abs_diff": d_dba, "sub_database_rows": n_sub, "queries": n_q}}


def parity_extract(dev, rank, world):
    """extract_vectors_dp (every rank extracts a slice, one all_gather) against the same extraction done by one rank."""
    from cirtorch_b200 import parallel as P
    from cirtorch_b200.extract import resnet50_gem, extract_vectors
    torch.manual_seed(3)
    net = resnet50_gem().to(dev).eval()
    imgs = torch.randn((2 * world + 1, 3, 256, 256), generator=torch.Generator().manual_seed(3))
    v_dp = P.extract_vectors_dp(net, imgs, image_size=256, transform=None, batch_size=2, device=dev)
    v_one = extract_vectors(net, imgs, image_size=256, transform=None, batch_size=2, device=dev)       # D x n on the CPU
    d = float((v_dp.cpu() - v_one).abs().max())
    return {"extract_dp_equals_single": _all_true(d < 1e-5, dev), "extract_detail": {"max_abs_diff": d, "images": int(imgs.shape[0])}}


# ------------------------------------------------------------------------------------------ main
def bench_tail(head, xs, steps, warmup, world, sampler=None):
    state = {"i": 0}

    def step():
        state["i"] ^= 1
        with torch.no_grad():
            return head(xs[state["i"]])

    for _ in range(3):
        step()
    return timed_region(step, steps, warmup, world, sampler), step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-search", action="store_true", help="tail only (skips configs 0, 2, 3, 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cirtorch_b200 import _lib
    _lib.load()
    pk = peaks()
    with_cpu = (not args.no_cpu_baseline) and world == 1
    head = make_head(dev)

    # two input batches (2 x 512 MiB) used alternately: every step streams data that is not in the 126 MB L2
    g = torch.Generator(device=dev).manual_seed(rank)
    xs = [torch.relu(torch.randn((B, C, H, W), device=dev, generator=g)) for _ in range(2)]

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    _lib.launch_count(reset=True)
    ms, step = bench_tail(head, xs, args.steps, args.warmup, world, sampler)
    launches = _lib.launch_count() - args.warmup - 4           # minus the untimed launches (3 + W + 1)
    per_rank_us = [round(1e3 * v / args.steps, 2) for v in LAST_PER_RANK_MS]
    clocks = sampler.finish() if sampler else None
    ms_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)
    achieved = TAIL_BYTES / (ms_step * 1e-3) / 1e9

    # the same step with a non-integer exponent (what a trained checkpoint has: p is learnable)
    head27 = make_head(dev, p=2.7)
    ms27, _ = bench_tail(head27, xs, args.steps, args.warmup, world)
    ms27 /= args.steps
    gb27 = TAIL_BYTES / (ms27 * 1e-3) / 1e9
    tail_p27 = {"p": 2.7, "ms_per_step": ms27, "descriptors_per_s": world * B / (ms27 * 1e-3),
                "roofline": {"bound": "hbm", "achieved": gb27, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb27 / pk["hbm_gbs"],
                             "traffic": None, "note": "general exponent: x^p = ex2(p lg2 x), one of every four ex2 on the FMA pipe; same algorithmic bytes"}}
    del head27

    # end to end through the module API with HOST buffers: H2D of the maps, the fused tail, D2H of the descriptors
    x_host = xs[0].cpu().pin_memory()
    d_host = torch.empty((B, DOUT), dtype=torch.float32).pin_memory()
    x_dev = torch.empty_like(xs[0])

    def step_e2e():
        x_dev.copy_(x_host, non_blocking=True)
        with torch.no_grad():
            d = head(x_dev)
        d_host.copy_(d.t(), non_blocking=True)

    e_steps = max(3, min(args.steps, 10))
    ms_e = timed_region(step_e2e, e_steps, 3, world)
    e2e = {"value": world * B * e_steps / (ms_e * 1e-3), "unit": "descriptors/s",
           "h2d_bytes_per_step": world * B * C * H * W * 4, "d2h_bytes_per_step": world * B * DOUT * 4}     # whole job
    del x_host, x_dev, xs
    torch.cuda.empty_cache()

    state = {"parity": {}}
    search = mining = extract_rank = None
    if not args.no_search:
        search = bench_search(dev, rank, world, pk, args.steps, args.warmup, state)
        alpha_qe, dba = search.pop("alpha_qe"), search.pop("dba")
        mining = bench_mining(dev, world, pk, with_cpu and rank == 0)
        extract_rank = bench_extract_rank(dev, rank, world, with_cpu and rank == 0)
        if world > 1:
            state["parity"].update(parity_extract(dev, rank, world))
    parity = state["parity"]
    parity_ok = all(v for k, v in parity.items() if isinstance(v, bool))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "descriptors/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "note": "BASELINE.json configs[1]; one batch per GPU per step (data parallel, no collective in the step)",
                       "l2": "inputs larger than L2: two 512 MiB batches used alternately",
                       "search_workload": "configs[3]: 1M x 2048 bf16 database row-sharded over the GPUs, top-100, 10k-query batch and 70 "
                                          "queries; strong scaling; rows i.i.d. / clustered (10k centres) / clustered and stored centre by centre",
                       "other_workloads": "configs[2] mining 2,000 x 20,000 nnum 5; configs[4] alpha-QE (k=10, alpha=3) of 10k queries + one "
                                          "full DBA pass over the 1M rows; configs[0] ResNet50-GeM extract (32 images per GPU at 1024 px) + "
                                          "70 x 4,993 ranking"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "per_rank_us_per_step": per_rank_us,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"], "traffic": TAIL_NCU_TRAFFIC,
                         "note": "554,180,608 algorithmic bytes per launch / mean launch time; peak = %s copy bandwidth; "
                                 "traffic = dram read+write per launch from profiles/r2_tail_p3.txt" % pk["src"]},
            "tail_p27": tail_p27, "search": search,
        }
        if not args.no_search:
            line.update({"mining": mining, "alpha_qe": alpha_qe, "dba": dba, "extract_rank": extract_rank})
        if world > 1:
            line["parity"] = dict(parity, all_ok=parity_ok)
        if with_cpu:
            line["cpu_baseline"] = cpu_tail_baseline()
            if search is not None:
                line["search"]["cpu_baseline"] = cpu_search_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not parity_ok:
        raise SystemExit("multi-GPU parity check failed: %s" % json.dumps(parity))


if __name__ == "__main__":
    main()
