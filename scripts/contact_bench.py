"""Contact-map / pair-graph construction benchmark: device kernels K11 / K12 vs the CPU restatement.

    python scripts/contact_bench.py [--pairs 20000] [--n-mean 250] [--iters 5]
Prints one JSON line: pairs/s and residue-pair distance evaluations/s on the GPU (CUDA events, inputs resident),
the vectorised NumPy restatement on a sample, and the reference's literal Python double loop on a smaller sample.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gcn_string_b200 import contact
from oracle import contact_ref


def walk(rng, n):
    step = rng.normal(size=(n, 3))
    step /= np.linalg.norm(step, axis=1, keepdims=True)
    return np.round(np.cumsum(3.8 * step, axis=0), 3).astype(np.float32)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=20000)
    ap.add_argument("--n-mean", type=int, default=250)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    n_chains = 2 * args.pairs
    lengths = np.clip(np.round(rng.lognormal(np.log(args.n_mean) - 0.08, 0.4, n_chains)), 30, 4 * args.n_mean).astype(np.int64)
    ca = np.concatenate([walk(rng, int(n)) for n in lengths])
    cp = contact.chain_offsets(lengths)
    pairs = np.arange(n_chains, dtype=np.int32).reshape(-1, 2)
    ba = np.concatenate([rng.integers(0, lengths[a], 20) for a in pairs[:, 0]]).astype(np.int32)
    bb = np.concatenate([rng.integers(0, lengths[b], 20) for b in pairs[:, 1]]).astype(np.int32)
    bptr = (20 * np.arange(args.pairs + 1)).astype(np.int32)
    ca_d, cp_d = torch.from_numpy(ca).cuda(), torch.from_numpy(cp).cuda()
    dev = [torch.from_numpy(v).cuda() for v in (pairs[:, 0].copy(), pairs[:, 1].copy(), bptr, ba, bb)]

    def run():
        r, c, _ = contact.contact_maps(ca_d, cp_d, 10)
        return contact.link_pairs(r, c, cp_d, *dev)

    for _ in range(2):
        out = run()
    torch.cuda.synchronize()
    times = []
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = run(); e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 1e3)
    t = float(np.median(times))
    evals = float((lengths.astype(np.float64) ** 2).sum())
    # CPU: vectorised restatement on 200 chains, literal reference loop on 2 chains
    sample = [ca[cp[k]:cp[k + 1]] for k in range(200)]
    t0 = time.perf_counter()
    for s in sample:
        contact_ref.contact_csr(s, 10)
    t_np = time.perf_counter() - t0
    ev_np = float(sum(len(s) ** 2 for s in sample))
    t0 = time.perf_counter()
    for s in sample[:2]:
        contact_ref.distance_matrix_loop(s)
    t_loop = time.perf_counter() - t0
    ev_loop = float(sum(len(s) ** 2 for s in sample[:2]))
    print(json.dumps({"pairs": args.pairs, "residues": int(lengths.sum()), "pair_graph_nnz": int(out[2].numel()),
                      "gpu_s": t, "gpu_pairs_per_s": args.pairs / t, "gpu_distance_evals_per_s": evals / t,
                      "gpu_includes": "contact count + scan + fill, link count + scan + fill, 2 host reads of nnz",
                      "numpy_vectorised_evals_per_s": ev_np / t_np, "reference_python_loop_evals_per_s": ev_loop / t_loop,
                      "speedup_vs_numpy_1core": (evals / t) / (ev_np / t_np),
                      "speedup_vs_reference_loop": (evals / t) / (ev_loop / t_loop)}))
