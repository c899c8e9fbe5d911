"""Developer probe: write-only HBM bandwidth of a plain fill, the ceiling of the assembly kernel."""
import torch
n = 15600
a = torch.empty((n, n), dtype=torch.float64, device="cuda")
for _ in range(3): a.fill_(1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(10):
    e0.record(); a.fill_(2.0); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
print(f"fill_ {a.numel()*8/1e9:.3f} GB: {best:.3f} ms -> {a.numel()*8/best/1e6:.0f} GB/s")
b = torch.empty_like(a)
for _ in range(10):
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize(); best2 = e0.elapsed_time(e1)
print(f"copy_: {best2:.3f} ms -> {2*a.numel()*8/best2/1e6:.0f} GB/s (read+write)")
