"""Is the spread of phase A over the CTAs (the first grid barrier waits for the slowest) a property of the SM or noise?
Per-CTA phase-A durations from the kernel's time stamps over several launches: run-to-run correlation, per-SM mean."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200"))
import torch
from cirtorch_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
B, Cc, H, W = 64, 2048, 32, 32
xs = [torch.relu(torch.randn((B, Cc, H, W), device=dev)) for _ in range(2)]
Wt = torch.randn(2048, 2048, device=dev) * 0.01
b = torch.zeros(2048, device=dev)
out = torch.empty((B, 2048), device=dev)
need = C.c_size_t(0); lib.cir_tail_workspace_bytes(B, Cc, 2048, C.byref(need))
ws = torch.zeros(need.value, dtype=torch.uint8, device=dev)
for p in (3.0, 2.7):
    pt = torch.full((1,), p, device=dev)
    flags = 0x80000000 | (8 if p == 3.0 else 0)
    durs, ends, smids = [], [], None
    for it in range(12):
        rc = lib.cir_tail_fwd(_lib.ptr(xs[it % 2]), B, Cc, H, W, _lib.ptr(pt), 0, 1e-6, 1e-6, 0, _lib.ptr(Wt), _lib.ptr(b), 2048,
                              _lib.ptr(out), 2048, None, _lib.ptr(ws), ws.numel(), flags, None)
        assert rc == 0
        torch.cuda.synchronize()
        st = ws[need.value - 65536:].view(torch.int64).view(-1, 8)[:148, :8].cpu().double()
        if it >= 2:
            t0 = st[:, 0].min()
            durs.append((st[:, 1] - st[:, 0]) / 1e3)
            ends.append((st[:, 1] - t0) / 1e3)
            smids = st[:, 7]
    D = torch.stack(durs)            # [runs, 148]
    E = torch.stack(ends)
    mean_cta = E.mean(0)
    cc = torch.corrcoef(E)           # run-to-run correlation of the per-CTA end times
    off = cc[~torch.eye(cc.shape[0], dtype=torch.bool)]
    print("p=%.1f: phase-A end per run: min %.1f mean %.1f max %.1f us; spread of the per-CTA MEANS over 10 runs: min %.1f max %.1f (std %.2f); "
          "run-to-run correlation of the per-CTA end times: mean %.2f (min %.2f)" % (
              p, E.min(1).values.mean(), E.mean(), E.max(1).values.mean(), mean_cta.min(), mean_cta.max(), mean_cta.std(), off.mean(), off.min()))
    order = torch.argsort(mean_cta)
    print("   slowest CTAs (block: smid, mean end us):", [(int(i), int(smids[i]), round(float(mean_cta[i]), 1)) for i in order[-8:]])
    print("   fastest CTAs:", [(int(i), int(smids[i]), round(float(mean_cta[i]), 1)) for i in order[:8]])
    own = torch.arange(148) < 128
    print("   mean end: CTAs with a projection unit %.1f, without %.1f" % (mean_cta[own].mean(), mean_cta[~own].mean()))
