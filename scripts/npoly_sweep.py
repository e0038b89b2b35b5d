#!/usr/bin/env python
"""Tail at 64 x 2048 x 32 x 32 with a non-integer exponent: time per launch for CIR_TAIL_NPOLY = 0..4 (how many of every
4 ex2 run as a polynomial on the FMA pipe).  One subprocess per setting (the variable is read once per process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "image-retrieval-for-image-based-localization_b200"))
from cirtorch_b200.modules.heads.global_head import globalHead
dev = torch.device("cuda:0")
for p in (2.7, 3.0):
    torch.manual_seed(0)
    head = globalHead(pooling={"name": "GeM", "params": {"p": p, "eps": 1e-6}}, normal={"name": "L2N", "params": {}}, dim=2048).to(dev).eval()
    xs = [torch.relu(torch.randn(64, 2048, 32, 32, device=dev)) for _ in range(2)]
    with torch.no_grad():
        for i in range(60): head(xs[i & 1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import pynvml as nv
        nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
        mhz, watts = [], []
        e0.record()
        for i in range(400):
            head(xs[i & 1])
        mhz.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)); watts.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
        e1.record(); torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
        ev[0].record()
        for i in range(40):
            head(xs[i & 1]); ev[i + 1].record()
        torch.cuda.synchronize()
        per = sorted(1e3 * ev[i].elapsed_time(ev[i + 1]) for i in range(40))
        print("   per-launch us: min %%.1f median %%.1f max %%.1f" %% (per[0], per[20], per[-1]))
    print("dynamic=%%s npoly=%%s p=%%.1f us_per_launch=%%.2f  sm_mhz(min/median)=%%d/%%d  power_w(max)=%%.0f" %% (os.environ.get("CIR_TAIL_DYNAMIC"), os.environ.get("CIR_TAIL_NPOLY", "default"), p,
          1e3 * e0.elapsed_time(e1) / 400, min(mhz), sorted(mhz)[len(mhz) // 2], max(watts)))
""" % (ROOT, ROOT)
for n, a in (("0", "0"), ("1", "0"), ("2", "0"), ("0", "0"), ("1", "0"), ("2", "0")):
    env = dict(os.environ, CIR_TAIL_NPOLY=n, CIR_TAIL_DYNAMIC=a)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout.strip() or r.stderr[-500:])
