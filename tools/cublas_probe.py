"""Developer probe: which kernel cuBLAS runs for the trailing-update shape (profile under ncu)."""
import torch
m, K = 15344, 256
x = torch.randn(m, K, dtype=torch.float64, device="cuda"); y = torch.randn(K, m, dtype=torch.float64, device="cuda")
c = torch.randn(m, m, dtype=torch.float64, device="cuda")
for _ in range(3): torch.addmm(c, x, y, alpha=-1.0, out=c)
torch.cuda.synchronize()
print("ok")
