"""Per-tensor gradient error of the hidden-512 / deg-32 case vs the float64 oracle for several chain lengths."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gcn_string_b200 as g
from gcn_string_b200 import synthetic, _lib
from gcn_string_b200.params import GNNConfig, block_specs, named_slices
from oracle import batching_ref, model_ref_np as O1, model_ref_torch as O2
ds = synthetic.make_dataset(3, seed=5, n_mean=1200, deg=32, n_feat=32)
graphs = [ds.graph(k) for k in range(3)]
(xr, (idx, _, _), seg), yr = batching_ref.collate(graphs)
cfg = GNNConfig(in_features=32, output=2, activation="softmax", hidden=512)
specs = block_specs(cfg)
w, s = g.init_params(cfg, seed=6, perturb=True)
for b in specs:
    o, n = b.alpha
    w[o:o+n] = 1.0
ref = O1.loss_and_grads(cfg, specs, w, s, xr, idx[:, 0], idx[:, 1], seg, yr, 3)
(x, a, i), y = next(g.DisjointLoader(ds, batch_size=3, epochs=1, shuffle=False))
floor = 0.1 * np.abs(ref["grads"]).max()
r2 = O2.loss_and_grads(cfg, specs, w, s, xr, idx[:, 0], idx[:, 1], seg, yr, 3)
e2 = {}
for name, shape, off, buf in named_slices(cfg):
    if buf != "trainable": continue
    n = int(np.prod(shape)); r = ref["grads"][off:off+n]
    e2[name] = np.abs(r2["grads"][off:off+n]-r).max()/max(np.abs(r).max(), floor)
print("O2 worst", sorted(e2.items(), key=lambda kv: -kv[1])[:4])
for chain_k, wchain in ((768, 16), (512, 16)):
    _lib.load().gcs_debug_set_param(5, chain_k); _lib.load().gcs_debug_set_param(3, wchain)
    m = g.GeneralGNN(2, activation="softmax", hidden=512); m.build(32); m.load_flat(w, s)
    la, probs = m.train_step_grads([x, a, i], y)
    got = m.grads.cpu().numpy()
    errs = {}
    for name, shape, off, buf in named_slices(cfg):
        if buf != "trainable": continue
        n = int(np.prod(shape)); r = ref["grads"][off:off+n]
        errs[name] = np.abs(got[off:off+n]-r).max()/max(np.abs(r).max(), floor)
    bad = {k: (f"{v:.2e}", f"{e2[k]:.2e}") for k, v in errs.items() if v > max(1e-5, 4 * e2[k])}
    print("   out of criterion:", bad)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    st = np.abs(m.state.cpu().numpy()-ref["new_state"]).max()/np.abs(ref["new_state"]).max()
    print(chain_k, wchain, "state", f"{st:.2e}", "worst", [(k, f"{v:.2e}") for k, v in worst], flush=True)
