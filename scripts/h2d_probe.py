import torch, time
dev = torch.device("cuda:0")
n = 64 * 2048 * 32 * 32
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
def run(parts):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    sz = n // parts
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        for i, s in enumerate(streams):
            s.wait_event(e0)
            with torch.cuda.stream(s):
                d[i * sz:(i + 1) * sz].copy_(h[i * sz:(i + 1) * sz], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return n * 4 / best / 1e6
for parts in (1, 2, 4, 8):
    print("H2D 512 MiB in %d concurrent parts: %.1f GB/s" % (parts, run(parts)))
