"""Tensor-core (tcgen05 3xTF32) GEMM vs the CUDA-core path and a float64 reference."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gcn_string_b200 import _lib, ops

lib = _lib.load()

def check(M, K, N, accumulate=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(K, N, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    ref = (A.double() @ W.double() + b.double())
    out = {}
    for mode, name in ((1, "ffma"), (2, "tc")):
        lib.gcs_debug_set_gemm_mode(mode)
        out[name] = ops.linear_fwd(A, W, b)
    torch.cuda.synchronize()
    lib.gcs_debug_set_gemm_mode(0)
    den = ref.abs().max().item()
    res = {"M": M, "K": K, "N": N, "err_ffma": (out["ffma"].double() - ref).abs().max().item() / den,
           "err_tc": (out["tc"].double() - ref).abs().max().item() / den}
    # backward-input form with accumulate
    dH = torch.randn(M, N, device="cuda", generator=g)
    base = torch.randn(M, K, device="cuda", generator=g)
    if K % 256 == 0:
        ref2 = base.double() + dH.double() @ W.double().T
        lib.gcs_debug_set_gemm_mode(2)
        got = ops.linear_bwd_input(dH, W, out=base.clone(), accumulate=True)
        lib.gcs_debug_set_gemm_mode(0)
        res["err_tc_bwd_input"] = (got.double() - ref2).abs().max().item() / ref2.abs().max().item()
    # weight gradient
    if K % 128 == 0:
        ref3 = A.double().T @ dH.double()
        r3 = {}
        for mode, name in ((1, "ffma"), (2, "tc")):
            lib.gcs_debug_set_gemm_mode(mode)
            dw, db = ops.linear_bwd_weight(A, dH)
            r3[name] = (dw.double() - ref3).abs().max().item() / ref3.abs().max().item()
        lib.gcs_debug_set_gemm_mode(0)
        res["err_wgrad_ffma"], res["err_wgrad_tc"] = r3["ffma"], r3["tc"]
    return res

def bench_wgrad(M, K, N, iters=5):
    A = torch.randn(M, K, device="cuda"); dH = torch.randn(M, N, device="cuda")
    r = {"wgrad": 1, "M": M, "K": K, "N": N}
    for mode, name in ((1, "ffma"), (2, "tc")):
        lib.gcs_debug_set_gemm_mode(mode)
        for _ in range(2): ops.linear_bwd_weight(A, dH)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): ops.linear_bwd_weight(A, dH)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        r[name + "_ms"] = ms; r[name + "_tflops"] = 2.0 * M * K * N / ms / 1e9
    lib.gcs_debug_set_gemm_mode(0)
    return r

def bench(M, K, N, iters=10):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(K, N, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    r = {"M": M, "K": K, "N": N}
    for mode, name in ((1, "ffma"), (2, "tc")):
        lib.gcs_debug_set_gemm_mode(mode)
        for _ in range(2): ops.linear_fwd(A, W, b, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): ops.linear_fwd(A, W, b, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        r[name + "_ms"] = ms; r[name + "_tflops"] = 2.0 * M * K * N / ms / 1e9
    lib.gcs_debug_set_gemm_mode(0)
    return r

if __name__ == "__main__":
    for shape in [(128, 32, 256), (300, 64, 256), (1000, 256, 256), (4096, 1024, 256), (777, 1280, 512), (70001, 256, 256)]:
        print(json.dumps(check(*shape)), flush=True)
    if "--bench" in sys.argv:
        for shape in [(516776, 32, 256), (516776, 256, 256), (516776, 512, 256), (516776, 1024, 256)]:
            print(json.dumps(bench(*shape)), flush=True)
        for shape in [(516776, 256, 256), (516776, 1024, 256)]:
            print(json.dumps(bench_wgrad(*shape)), flush=True)
