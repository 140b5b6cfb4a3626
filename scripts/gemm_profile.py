"""One tensor-core forward GEMM and one weight-gradient GEMM at cfg2 size, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_string_b200 import _lib, ops
M, K, N = 516776, 1024, 256
A = torch.randn(M, K, device="cuda"); W = torch.randn(K, N, device="cuda") / 32; b = torch.randn(N, device="cuda")
dH = torch.randn(M, N, device="cuda"); out = torch.empty(M, N, device="cuda")
for _ in range(3):
    ops.linear_fwd(A, W, b, out=out)
    ops.linear_bwd_weight(A, dH)
torch.cuda.synchronize()
print("ok")
