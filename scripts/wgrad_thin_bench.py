"""First-layer weight gradient (K = 32): thin kernel vs the tiled FFMA kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_string_b200 import _lib, ops
lib = _lib.load()
M, N = 514799, 256
for K in (32, 16):
    a = torch.rand(M, K, device="cuda"); dh = torch.randn(M, N, device="cuda")
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    for thin in (1, 0):
        lib.gcs_debug_set_param(13, thin)
        for _ in range(3): ops.linear_bwd_weight(a, dh, want_db=False)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.linear_bwd_weight(a, dh, want_db=False); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(json.dumps({"K": K, "thin": thin, "us": round(ts[len(ts)//2] * 1e3, 1), "GBs": round((M * (K + N) * 4) / ts[len(ts)//2] / 1e6, 1)}))
lib.gcs_debug_set_param(13, 1)
