"""Diagnostic: per-tensor gradient error of the CUDA path and of the fp32 CPU restatement
(O2) against the float64 oracle (O1), with random PReLU slopes and with alpha = 1 (no kink)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gcn_string_b200 as g
from gcn_string_b200 import synthetic
from gcn_string_b200.params import GNNConfig, block_specs, named_slices
from oracle import batching_ref, model_ref_np as O1, model_ref_torch as O2

def run(n_graphs, n_mean, H, L, smooth, seed=0):
    ds = synthetic.make_dataset(n_graphs, seed=seed, n_mean=n_mean, deg=12, n_feat=32)
    graphs = [ds.graph(k) for k in range(n_graphs)]
    (xr, (idx, _, _), seg), yr = batching_ref.collate(graphs)
    cfg = GNNConfig(in_features=32, output=2, activation="softmax", hidden=H, message_passing=L)
    specs = block_specs(cfg)
    w, s = g.init_params(cfg, seed=4, perturb=True)
    if smooth:
        for b in specs:
            o, n = b.alpha
            w[o:o+n] = 1.0
    ref = O1.loss_and_grads(cfg, specs, w, s, xr, idx[:, 0], idx[:, 1], seg, yr, n_graphs)
    r2 = O2.loss_and_grads(cfg, specs, w, s, xr, idx[:, 0], idx[:, 1], seg, yr, n_graphs)
    (x, a, i), y = next(g.DisjointLoader(ds, batch_size=n_graphs, epochs=1, shuffle=False))
    m = g.GeneralGNN(2, activation="softmax", hidden=H, message_passing=L); m.build(32); m.load_flat(w, s)
    la, probs = m.train_step_grads([x, a, i], y)
    got = m.grads.cpu().numpy()
    nk = sum(int((np.abs(c["z"]) < 1e-5).sum()) for c, b in zip(ref["ctx"]["caches"], specs) if b.has_alpha)
    print(f"== graphs={n_graphs} n_mean={n_mean} H={H} L={L} smooth={smooth}  N={xr.shape[0]}  near-kink(|z|<1e-5)={nk}")
    print(f"   loss gpu {la[0].item():.8f} o2 {r2['loss']:.8f} o1 {ref['loss']:.8f}; probs err gpu {np.abs(probs.cpu().numpy()-ref['probs']).max():.2e} o2 {np.abs(r2['probs']-ref['probs']).max():.2e}")
    gmax = np.abs(ref["grads"]).max()
    print(f"   global max-rel: gpu {np.abs(got-ref['grads']).max()/gmax:.2e}  o2 {np.abs(r2['grads']-ref['grads']).max()/gmax:.2e};  L2-rel: gpu {np.linalg.norm(got-ref['grads'])/np.linalg.norm(ref['grads']):.2e} o2 {np.linalg.norm(r2['grads']-ref['grads'])/np.linalg.norm(ref['grads']):.2e}")
    for name, shape, off, buf in named_slices(cfg):
        if buf != "trainable": continue
        n = int(np.prod(shape)); r = ref["grads"][off:off+n]
        den = max(np.abs(r).max(), 1e-3 * gmax)
        print(f"   {name:22s} |ref|={np.abs(r).max():.2e} gpu {np.abs(got[off:off+n]-r).max()/den:.2e} o2 {np.abs(r2['grads'][off:off+n]-r).max()/den:.2e}")

if __name__ == "__main__":
    from gcn_string_b200 import _lib
    if len(sys.argv) > 1:
        for c in sys.argv[1:]:
            _lib.load().gcs_debug_set_param(3, int(c))
            print("#### wgrad chain", c)
            run(8, 500, 256, 4, True)
    else:
        run(8, 500, 256, 4, False)
        run(8, 500, 256, 4, True)
        run(6, 50, 32, 4, True, seed=12)
