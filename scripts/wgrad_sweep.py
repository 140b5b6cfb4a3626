"""Weight-gradient GEMM time vs accumulation-chain length (debug param 3)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_string_b200 import _lib, ops
lib = _lib.load()
lib.gcs_debug_set_gemm_mode(2)
M, N = 516776, 256
for K in (256, 1024):
    A = torch.randn(M, K, device="cuda"); dH = torch.randn(M, N, device="cuda")
    ref = None
    for chain in (16, 32, 64, 128):
        lib.gcs_debug_set_param(3, chain)
        for _ in range(2): dw, db = ops.linear_bwd_weight(A, dH)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.linear_bwd_weight(A, dH)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"K": K, "chain": chain, "ms": e0.elapsed_time(e1) / 5}), flush=True)
