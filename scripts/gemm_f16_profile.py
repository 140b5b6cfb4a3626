"""One fp16-split tensor-core forward GEMM at cfg2 size (K = 1024), for ncu: -k regex:linear_tc_pair_kernel -s 2 -c 1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_string_b200 import _lib, ops
lib = _lib.load()
lib.gcs_debug_set_param(7, 2)          # fp16 kernel for the standalone op (|max| of A by an extra pass)
lib.gcs_debug_set_gemm_mode(2)
M, K, N = 516776, 1024, 256
A = torch.randn(M, K, device="cuda"); W = torch.randn(K, N, device="cuda") / 32; b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda")
for _ in range(3):
    ops.linear_fwd(A, W, b, out=out)
torch.cuda.synchronize()
print("ok")
