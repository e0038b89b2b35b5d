#!/usr/bin/env python
"""Benchmark of the global-descriptor retrieval hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-search]

Headline line (`metric` = descriptors/s): one step = one pass of the fused GeM+L2N+whiten+L2N tail
over a batch of 64 x 2048 x 32 x 32 fp32 layer-4 maps (BASELINE.json configs[1]), whitening
2048 -> 2048, synthetic data, random-init weights.  N > 1 (torchrun): data parallel, one batch per
rank per step (weak scaling).  The same JSON line carries, under "search", queries/s of the
1M x 2048 top-100 search (configs[3]) with the database row-sharded over the N GPUs (strong
scaling: one NCCL all_gather of the per-query lists + merge), and the roofline of both kernels.

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
TAIL_NCU_TRAFFIC = 553_768_448 + 4_396_288      # dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full (profiles/r1h_tail.txt)
DB_N, DB_D, TOPK = 1_000_000, 2048, 100
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
    """W untimed steps, then exactly K timed steps between barrier + synchronize; max over ranks (device time)."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
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


def make_head(dev):
    from cirtorch_b200.modules.heads.global_head import globalHead
    torch.manual_seed(0)
    head = globalHead(pooling={"name": "GeM", "params": {"p": 3, "eps": 1e-6}},
                      normal={"name": "L2N", "params": {}}, dim=DOUT)
    return head.to(dev).eval()


def cpu_tail_baseline(budget_s=12.0, max_iters=30):
    """The oracle port of globalHead.forward on the host cores, batch 64 x 2048 x 32 x 32."""
    from oracle import cirtorch_oracle as O
    host_threads()
    torch.manual_seed(0)
    x = torch.relu(torch.randn(B, C, H, W))
    Wt = torch.empty(DOUT, C)
    torch.nn.init.xavier_normal_(Wt, 0.1)
    b = torch.zeros(DOUT)
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


def cpu_search_baseline(n_rows=100_000, q=70, budget_s=10.0):
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


def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    host_threads()
    from oracle import cirtorch_oracle as O
    torch.manual_seed(0)
    x = torch.relu(torch.randn(B, C, H, W))
    Wt = torch.empty(DOUT, C)
    torch.nn.init.xavier_normal_(Wt, 0.1)
    b = torch.zeros(DOUT)
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
        "config": {"workload": "fused tail: batch 64x2048x32x32 fp32 maps -> GeM(p=3)+L2N+whiten 2048->2048+L2N",
                   "note": "reference CPU path = oracle port of globalHead.forward (global_head.py:52-67), torch CPU eager"},
        "cpu_baseline": {"value": val, "unit": "descriptors/s", "cores": threads, "kind": "port",
                         "sample": "%d steps of one 64-image batch, %d host threads" % (args.steps, threads)},
        "e2e": {"value": val, "unit": "descriptors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bench_search(dev, rank, world, pk, steps, warmup):
    """1M x 2048 database row-sharded over the ranks, top-100: 10k-query batch (tensor-bound) and 70 queries (HBM-bound)."""
    import torch.distributed as dist
    from cirtorch_b200 import search as S, parallel as P, _lib
    lo, hi = P.shard_bounds(DB_N, world, rank)
    n_loc = hi - lo
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    dbp = torch.empty((n_loc, DB_D), dtype=torch.bfloat16, device=dev)
    rows32 = torch.empty((n_loc, DB_D), dtype=torch.float32, device=dev)
    for a in range(0, n_loc, 62_500):
        b = min(n_loc, a + 62_500)
        blk = torch.randn((b - a, DB_D), device=dev, generator=g)
        rows32[a:b] = blk / blk.norm(dim=1, keepdim=True)
    S.pack_rows(rows32, "db", "bf16", out=dbp)
    gq = torch.Generator(device=dev).manual_seed(99)
    out = {}
    for name, Q in (("q10k", 10_000), ("q70", 70)):
        q32 = torch.randn((Q, DB_D), device=dev, generator=gq)
        q32 = q32 / q32.norm(dim=1, keepdim=True)
        qp = S.pack_rows(q32, "query", "bf16")

        both, views = P.topk_exchange_buffer(Q, TOPK, dev)
        shard = None
        exchange = "none" if world == 1 else "nccl all_gather"
        if world > 1:
            # fused exchange: the selection kernel stores the lists into every peer's buffer over NVLink
            try:
                shard = P.ShardedIndex.__new__(P.ShardedIndex)
                shard._exchange, shard.group, shard.rank, shard.world_size, shard.lo, shard.hi, shard.n_global = \
                    {}, None, rank, world, lo, hi, DB_N
                idx_obj = S.Index.__new__(S.Index)
                idx_obj.mode, idx_obj.N, idx_obj.D, idx_obj.row_offset, idx_obj.packed, idx_obj.rows32, idx_obj.labels = \
                    "bf16", n_loc, DB_D, lo, dbp, rows32, None
                shard.index = idx_obj
                shard.search_packed_p2p(qp, TOPK)
                torch.cuda.synchronize()
                exchange = "peer stores over NVLink fused into the selection kernel (symmetric memory)"
            except Exception as e:  # noqa: BLE001  (no symmetric memory on this box: NCCL all_gather instead)
                shard = None
                exchange = "nccl all_gather (%s)" % type(e).__name__
            flag = torch.tensor([1 if shard is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag) == 0:
                shard = None

        def step_local():
            return S.search_packed(qp, dbp, TOPK, idx_offset=lo, out=views)

        def step():
            if shard is not None:
                return shard.search_packed_p2p(qp, TOPK)
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
        res = {"queries_per_s": Q / (ms * 1e-3), "ms_per_search": ms, "ms_local_kernels": ms_kernel,
               "steps": k_steps, "launches_per_search": launches // (k_steps + max(3, warmup)), "Q": Q, "N": DB_N, "k": TOPK,
               "exchange": exchange, "per_rank_ms_per_search": per_rank}
        if Q >= 1000:
            tf = flops / (ms_kernel * 1e-3) / 1e12
            res["roofline"] = {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                               "frac": tf / pk["bf16_tflops_sustained"], "traffic": None,
                               "note": "per GPU: 2*Q*N_local*D flop / local search time (GEMM + top-k select); peak = "
                                       + pk["src"] + " sustained cuBLAS bf16"}
        else:
            gb = (n_loc * DB_D * 2 + Q * DB_D * 2 + Q * TOPK * 8) / (ms_kernel * 1e-3) / 1e9
            res["roofline"] = {"bound": "hbm", "achieved": gb, "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": gb / pk["hbm_gbs"], "traffic": None,
                               "note": "per GPU: bf16 shard scan bytes / local search time; peak = " + pk["src"] + " copy bandwidth"}
        # end to end through the public API with host buffers: H2D queries, pack, scan, fp32 re-score, D2H lists
        if name == "q10k":
            index = S.Index.__new__(S.Index)
            index.mode, index.N, index.D, index.row_offset, index.packed, index.rows32, index.labels = \
                "bf16", n_loc, DB_D, lo, dbp, rows32, None
            q_host = q32.cpu().pin_memory()
            s_host = torch.empty((Q, TOPK), dtype=torch.float32).pin_memory()
            i_host = torch.empty((Q, TOPK), dtype=torch.int32).pin_memory()

            sh_e2e = P.ShardedIndex.__new__(P.ShardedIndex)
            sh_e2e._exchange, sh_e2e.group, sh_e2e.rank, sh_e2e.world_size, sh_e2e.lo, sh_e2e.hi, sh_e2e.n_global = \
                {}, None, rank, world, lo, hi, DB_N
            sh_e2e.index = index

            def step_e2e():
                # every rank needs all queries: each copies 1/G of them over PCIe, one all_gather over NVLink
                qd = P.replicate_host_rows(q_host, dev)
                s, i = sh_e2e.search_rows(qd, TOPK)           # bf16 scan, global top-128, exact fp32 re-score, merge
                s_host.copy_(s, non_blocking=True)
                i_host.copy_(i, non_blocking=True)

            ms_e = timed_region(step_e2e, k_steps, 3, world) / k_steps
            res["e2e"] = {"value": Q / (ms_e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": Q * DB_D * 4,
                          "d2h_bytes_per_step": world * Q * TOPK * 8,
                          "includes": "whole job: every query row crosses PCIe once (1/G per rank + all_gather over NVLink), every rank "
                                      "reads the final lists back; fp32 re-score of 128 candidates"}
        out[name] = res
        del q32, qp
    del dbp, rows32
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-search", action="store_true")
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
    head = make_head(dev)

    # two input batches (2 x 512 MiB) used alternately: every step streams data that is not in the 126 MB L2
    g = torch.Generator(device=dev).manual_seed(rank)
    xs = [torch.relu(torch.randn((B, C, H, W), device=dev, generator=g)) for _ in range(2)]
    state = {"i": 0}

    def step():
        state["i"] ^= 1
        with torch.no_grad():
            return head(xs[state["i"]])

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(3):
        step()
    _lib.launch_count(reset=True)
    ms = timed_region(step, args.steps, args.warmup, world, sampler)
    launches = _lib.launch_count() - args.warmup
    per_rank_us = [round(1e3 * v / args.steps, 2) for v in LAST_PER_RANK_MS]
    clocks = sampler.finish() if sampler else None
    ms_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)

    # end to end through the module API with HOST buffers: H2D of the maps, the fused tail, D2H of the descriptors
    x_host = [x.cpu().pin_memory() for x in xs[:1]]
    d_host = torch.empty((B, DOUT), dtype=torch.float32).pin_memory()
    x_dev = torch.empty_like(xs[0])

    def step_e2e():
        x_dev.copy_(x_host[0], non_blocking=True)
        with torch.no_grad():
            d = head(x_dev)
        d_host.copy_(d.t(), non_blocking=True)

    e_steps = max(3, min(args.steps, 10))
    ms_e = timed_region(step_e2e, e_steps, 3, world)
    e2e = {"value": world * B * e_steps / (ms_e * 1e-3), "unit": "descriptors/s",
           "h2d_bytes_per_step": world * B * C * H * W * 4, "d2h_bytes_per_step": world * B * DOUT * 4}     # whole job
    del x_host, x_dev
    achieved = TAIL_BYTES / (ms_step * 1e-3) / 1e9

    search = None
    if not args.no_search:
        del xs
        torch.cuda.empty_cache()
        search = bench_search(dev, rank, world, pk, args.steps, args.warmup)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "descriptors/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "fused tail: batch 64x2048x32x32 fp32 maps -> GeM(p=3)+L2N+whiten 2048->2048+L2N "
                                   "(BASELINE.json configs[1]); one batch per GPU per step",
                       "l2": "inputs larger than L2: two 512 MiB batches used alternately",
                       "search_workload": "1M x 2048 bf16 database row-sharded over the GPUs, top-100, 10k-query batch and 70 queries "
                                          "(configs[3]); strong scaling"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "per_rank_us_per_step": per_rank_us,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"], "traffic": TAIL_NCU_TRAFFIC,
                         "note": "554,180,608 algorithmic bytes per launch / mean launch time; peak = %s copy bandwidth; "
                                 "traffic = dram read+write per launch from profiles/r1h_tail.txt" % pk["src"]},
            "search": search,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_tail_baseline()
            if search is not None:
                line["search"]["cpu_baseline"] = cpu_search_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
