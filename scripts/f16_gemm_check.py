"""fp16-split tensor-core GEMM (linear_tc_pair_kernel<true>) vs the tf32-split kernel and float64: accuracy on
well- and badly-scaled operands, then timings.  gcs_debug_set_param(7, 2) lets the standalone ops take the fp16
kernel (the |max| of A then comes from an extra pass, which the timings below include)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_string_b200 import _lib, ops

lib = _lib.load()


def run(mode, fn):
    lib.gcs_debug_set_param(7, mode)
    lib.gcs_debug_set_gemm_mode(2)
    try:
        return fn()
    finally:
        lib.gcs_debug_set_gemm_mode(0)
        lib.gcs_debug_set_param(7, 1)


def check(M, K, N, a_mag=1.0, w_mag=1.0, spread=0.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g) * a_mag
    if spread:
        A = A * torch.exp2(torch.randint(-int(spread), 1, (M, K), device="cuda", generator=g).float())
    W = torch.randn(K, N, device="cuda", generator=g) / K ** 0.5 * w_mag
    b = torch.randn(N, device="cuda", generator=g) * a_mag * w_mag
    ref = A.double() @ W.double() + b.double()
    den = ref.abs().max().item()
    res = {"M": M, "K": K, "N": N, "a_mag": a_mag, "w_mag": w_mag, "spread": spread}
    for mode, name in ((0, "tf32"), (2, "f16")):
        out = run(mode, lambda: ops.linear_fwd(A, W, b))
        res["fwd_" + name] = (out.double() - ref).abs().max().item() / den
    dH = torch.randn(M, N, device="cuda", generator=g) * a_mag
    base = torch.randn(M, K, device="cuda", generator=g) * a_mag * w_mag
    ref2 = base.double() + dH.double() @ W.double().T
    for mode, name in ((0, "tf32"), (2, "f16")):
        got = run(mode, lambda: ops.linear_bwd_input(dH, W, out=base.clone(), accumulate=True))
        res["dx_" + name] = (got.double() - ref2).abs().max().item() / ref2.abs().max().item()
    return res


def bench(M, K, N, iters=10):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(K, N, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    r = {"bench": 1, "M": M, "K": K, "N": N}
    for mode, name in ((0, "tf32"), (2, "f16")):
        def go():
            for _ in range(2): ops.linear_fwd(A, W, b, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters): ops.linear_fwd(A, W, b, out=out)
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        ms = run(mode, go)
        r[name + "_ms"] = round(ms, 4); r[name + "_tflops"] = round(2.0 * M * K * N / ms / 1e9, 1)
    return r


if __name__ == "__main__":
    for shape in [(256, 128, 128), (1000, 256, 256), (4096, 1024, 256), (777, 1280, 512), (70001, 256, 256), (300, 2048, 128)]:
        print(json.dumps(check(*shape)), flush=True)
    for kw in (dict(a_mag=1e-7), dict(a_mag=3e4), dict(w_mag=1e-6), dict(spread=30), dict(a_mag=1e-20, w_mag=1e5)):
        print(json.dumps(check(2048, 512, 256, **kw)), flush=True)
    if "--bench" in sys.argv:
        for shape in [(516776, 256, 256), (516776, 512, 256), (516776, 768, 256), (516776, 1024, 256)]:
            print(json.dumps(bench(*shape)), flush=True)
