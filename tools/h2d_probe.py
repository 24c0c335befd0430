"""Raw pinned host->device copy rate on this box, linear and 2-D (1000 rows), for the e2e ceiling."""
import torch, time
n = 1000 * 1000000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    d.copy_(h, non_blocking=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print("linear 1 GB pinned H2D: %.2f ms -> %.1f GB/s" % (ms, n / ms / 1e6))
h2 = h.view(1000, 1000000); d2 = torch.empty(1000, 1000064, dtype=torch.uint8, device="cuda")
a.record()
for _ in range(5):
    d2[:, :1000000].copy_(h2, non_blocking=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print("2-D (1000 rows, pitched dst) H2D: %.2f ms -> %.1f GB/s" % (ms, n / ms / 1e6))
