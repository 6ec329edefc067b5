import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ellspmv_b200 as E
rows, K = 32768, 4096
A = E.EllMatrix.generate(E.GEN_RANDOM, (rows, 1 << 20, K), (0.0, 0.0), 42, 32, flags=E.KERNEL_LONGROW)
x = torch.randn(1 << 20, dtype=torch.float64, device="cuda")
y = torch.zeros(rows, dtype=torch.float64, device="cuda")
for _ in range(3):
    A.spmv_device(y, x, E.OVERWRITE, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok")
