"""Host <-> device copy bandwidth of this box for the e2e leg's transfer sizes (pinned memory, CUDA events)."""
import torch

for mb in (1.5, 14, 64, 256):
    n = int(mb * 1e6 / 8)
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    for name, dst, src in (("d2h", h, d), ("h2d", d, h)):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name} {mb:6.1f} MB: {ms:.4f} ms  {n * 8 / ms / 1e6:.1f} GB/s")
# both directions at once (two streams)
n = int(14e6 / 8)
h1, h2 = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
d1, d2 = torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
print("duplex 14 MB each way: %.1f GB/s per direction (wall)" % (n * 8 * 10 / (e0.elapsed_time(e1)) / 1e6))
