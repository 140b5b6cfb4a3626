"""SpMM (GeneralConv aggregation) micro-benchmark: BASELINE.json configs[3] sweep and the
kernel ncu profiles are taken from this script.

    python scripts/spmm_bench.py [--graphs 1024] [--n-mean 500] [--deg 12] [--hidden 256]
                                 [--mode auto|rows|staged] [--iters 20] [--sweep]
Prints one JSON line per configuration: us/launch, algorithmic GB/s (SURVEY.md §8d formula),
fraction of the measured HBM peak, edges/s.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gcn_string_b200 as g
from gcn_string_b200 import _lib, ops, synthetic


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def make_batch(n_graphs, n_mean, deg, seed=0):
    ds = synthetic.make_dataset(n_graphs, seed=seed, n_mean=n_mean, deg=deg, n_feat=4)
    store = g.data.DeviceGraphStore(ds, symmetric=True)
    ids = np.arange(n_graphs, dtype=np.int64)
    _, a, _, _ = store.batch(torch.from_numpy(ids).cuda(), ids)
    return a


def run(a, H, mode, iters, transform=True, flush=None):
    lib = _lib.load()
    lib.gcs_debug_set_spmm_mode({"rows": 1, "rb4": 2}[mode])
    n = a.n_rows
    x = torch.randn(n, H, device="cuda")
    y = torch.empty(n, H, device="cuda")
    sc, sh, al = (torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda"), torch.rand(H, device="cuda") * 0.3)
    kw = dict(rb4=a.rb4) if mode == "rb4" else {}
    args = (sc, sh, al) if transform else (None, None, None)
    for _ in range(3):
        ops.spmm_sum(a.rowptr, a.colidx, x, *args, out=y, **kw)
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                                   # evict L2 between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.spmm_sum(a.rowptr, a.colidx, x, *args, out=y, **kw)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    lib.gcs_debug_set_spmm_mode(0)
    t = float(np.median(times)) / 1e3
    alg = 4.0 * n * H * 2 + 4.0 * a.nnz + 4.0 * (n + 1)
    return {"H": H, "n_rows": n, "nnz": a.nnz, "deg": a.nnz / n, "mode": mode, "us": t * 1e6,
            "rb4_ratio": (a.nnz / int(a.rb4[0][-1].item())) if mode == "rb4" else None, "alg_GBs": alg / t / 1e9, "frac_measured_hbm": alg / t / 1e9 / peak(), "edges_per_s": a.nnz / t}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", type=int, default=1024)
    ap.add_argument("--n-mean", type=int, default=500)
    ap.add_argument("--deg", type=int, default=12)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--no-transform", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sweep", action="store_true", help="BASELINE cfg4: H 16..512 x deg 4..64")
    ap.add_argument("--param", type=int, nargs=2, action="append", default=[], help="debug knob: id value")
    args = ap.parse_args()
    for pid, val in args.param:
        _lib.load().gcs_debug_set_param(pid, val)
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    if args.sweep:
        for deg in (4, 8, 16, 32, 64):
            a = make_batch(args.graphs, args.n_mean, deg)
            for H in (16, 32, 64, 128, 256, 512):
                for mode in ("rows", "rb4"):
                    if mode == "rb4" and 256 % (H // 4):
                        continue
                    print(json.dumps(run(a, H, mode, args.iters, flush=flush)), flush=True)
    else:
        a = make_batch(args.graphs, args.n_mean, args.deg)
        for mode in ([args.mode] if args.mode != "auto" else ["rows", "rb4"]):
            print(json.dumps(run(a, args.hidden, mode, args.iters, transform=not args.no_transform, flush=flush)), flush=True)
