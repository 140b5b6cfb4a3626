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


def run(a, H, mode, iters, transform=True, flush=None, check=None, ldy=None):
    """mode: rows | rb4 (global-memory kernels) | slab1 | slab2 | slab4 (per-graph shared-memory slabs; the digit is
    the row-block height, 1 = CSR).  The output is a column slice of a [n, ldy] buffer as in the 'cat' layout."""
    lib = _lib.load()
    lib.gcs_debug_set_spmm_mode({"rows": 1, "rb4": 2}.get(mode, 0))
    n = a.n_rows
    x = torch.randn(n, H, device="cuda")
    ybuf = torch.empty(n, ldy or H, device="cuda")
    y = ybuf[:, :H]
    sc, sh, al = (torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda"), torch.rand(H, device="cuda") * 0.3)
    args = (sc, sh, al) if transform else (None, None, None)
    if mode.startswith("slab"):
        hgt = int(mode[4:])
        rb = a.rb(hgt)
        call = lambda: ops.spmm_sum_graphs(a.graph_ptr, a.max_graph_nodes, a.rowptr, a.colidx, x, *args, out=y, rb=rb, rb_height=hgt)
        ratio = a.nnz / int(rb[0][-1].item()) if rb is not None else 1.0
    else:
        kw = dict(rb4=a.rb4) if mode == "rb4" else {}
        call = lambda: ops.spmm_sum(a.rowptr, a.colidx, x, *args, out=y, **kw)
        ratio = (a.nnz / int(a.rb4[0][-1].item())) if mode == "rb4" else None
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ok = None
    if check is not None:
        lib.gcs_debug_set_spmm_mode(1)
        ref = ops.spmm_sum(a.rowptr, a.colidx, x, *args)
        lib.gcs_debug_set_spmm_mode({"rows": 1, "rb4": 2}.get(mode, 0))
        ok = float((ref - y).abs().max() / ref.abs().max())   # a reordered float32 sum: ~1e-7
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                                   # evict L2 between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    lib.gcs_debug_set_spmm_mode(0)
    t = float(np.median(times)) / 1e3
    alg = 4.0 * n * H * 2 + 4.0 * a.nnz + 4.0 * (n + 1)
    out = {"H": H, "n_rows": n, "nnz": a.nnz, "deg": round(a.nnz / n, 2), "mode": mode, "prologue": bool(transform), "us": round(t * 1e6, 1),
           "union_ratio": ratio, "alg_GBs": round(alg / t / 1e9, 1), "frac_measured_hbm": round(alg / t / 1e9 / peak(), 4),
           "edges_per_s": a.nnz / t}
    if ok is not None:
        out["max_rel_diff_vs_rows_kernel"] = ok
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", type=int, default=1024)
    ap.add_argument("--n-mean", type=int, default=500)
    ap.add_argument("--deg", type=int, default=12)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--mode", default="auto", help="comma-separated list of rows,rb4,slab1,slab2,slab4")
    ap.add_argument("--both", action="store_true", help="time with and without the BN+PReLU prologue")
    ap.add_argument("--check", action="store_true", help="compare bit for bit with the CSR row kernel")
    ap.add_argument("--ldy", type=int, default=None)
    ap.add_argument("--no-transform", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sweep", action="store_true", help="BASELINE cfg4: H 16..512 x deg 4..64")
    ap.add_argument("--param", type=int, nargs=2, action="append", default=[], help="debug knob: id value")
    args = ap.parse_args()
    for pid, val in args.param:
        _lib.load().gcs_debug_set_param(pid, val)
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    modes = ["rows", "rb4", "slab2", "slab4"] if args.mode == "auto" else args.mode.split(",")
    chk = True if args.check else None
    if args.sweep:
        for deg in (4, 8, 16, 32, 64):
            a = make_batch(args.graphs, args.n_mean, deg)
            for H in (16, 32, 64, 128, 256, 512):
                for mode in modes:
                    if mode == "rb4" and 256 % (H // 4):
                        continue
                    for tr in (True, False):
                        print(json.dumps(run(a, H, mode, args.iters, transform=tr, flush=flush, check=chk)), flush=True)
    else:
        a = make_batch(args.graphs, args.n_mean, args.deg)
        for mode in modes:
            for tr in ((True, False) if args.both else (not args.no_transform,)):
                print(json.dumps(run(a, args.hidden, mode, args.iters, transform=tr, flush=flush, check=chk, ldy=args.ldy)), flush=True)
