"""Raw pinned host -> device copy bandwidth for one Case-B cube pair (what bounds bench.py's e2e arm)."""
import torch, time
n = 2 * 180 * 1024 * 1024
h = [torch.empty(n // 2, dtype=torch.int16).pin_memory() for _ in range(2)]
d = [torch.empty(n // 2, dtype=torch.int16, device="cuda") for _ in range(2)]
for t in h: t.zero_()
def run(reps=10, two_streams=False):
    s2 = torch.cuda.Stream()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d[0].copy_(h[0], non_blocking=True)
        if two_streams:
            with torch.cuda.stream(s2):
                d[1].copy_(h[1], non_blocking=True)
        else:
            d[1].copy_(h[1], non_blocking=True)
    if two_streams:
        torch.cuda.current_stream().wait_stream(s2)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return 2 * n / ms / 1e6
print("one stream   GB/s", run()); print("two streams  GB/s", run(two_streams=True))
