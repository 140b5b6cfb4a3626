"""One training step at BASELINE cfg5 per-GPU size (64 graphs of ~5000 residues, deg ~32, hidden 512): time + memory."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gcn_string_b200 as g
from gcn_string_b200 import _lib, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
if len(sys.argv) > 2: _lib.load().gcs_debug_set_param(5, int(sys.argv[2]))
t0 = time.time()
ds = synthetic.make_dataset(B, seed=0, n_mean=5000, deg=32, n_feat=32)
loader = g.DisjointLoader(ds, batch_size=B, epochs=None, shuffle=False, symmetric=True)
model = g.GeneralGNN(2, activation="softmax", hidden=512, message_passing=4, seed=0); model.build(32)
(x, a, i), y = next(loader)
for _ in range(2): model.train_step_grads((x, a, i), y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): model.train_step_grads((x, a, i), y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
flops = 3 * 2.0 * a.n_rows * (32 * 512 + 512 * 512 * 11)
_lib.profile_begin(); model.train_step_grads((x, a, i), y); torch.cuda.synchronize(); prof = _lib.profile_end()
print(json.dumps({"config": "cfg5 per-GPU step", "graphs": B, "nodes": a.n_rows, "nnz": a.nnz, "hidden": 512, "ms_per_step": ms,
                  "graphs_per_s": B / ms * 1e3, "fp32_equiv_tflops": flops / ms / 1e9,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                  "ops_ms": {k: round(v[1], 2) for k, v in prof.items()}}))
