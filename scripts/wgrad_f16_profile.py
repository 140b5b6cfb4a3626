"""One fp16-split weight-gradient GEMM at cfg2 size (K_in = 1024), for ncu: -k regex:wgrad_tc_pair_kernel -s 2 -c 1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_string_b200 import _lib, ops
lib = _lib.load()
lib.gcs_debug_set_param(7, 2)          # fp16 kernels for the standalone ops (|max| of the operands by an extra pass)
lib.gcs_debug_set_gemm_mode(2)
M, K, N = 516776, 1024, 256
A = torch.randn(M, K, device="cuda"); dH = torch.randn(M, N, device="cuda") * 1e-4
for _ in range(3):
    ops.linear_bwd_weight(A, dH, want_db=False)
torch.cuda.synchronize()
print("ok")
