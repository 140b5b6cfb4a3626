"""Per-op device times of the inference forward (cfg2) and of one train step."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gcn_string_b200 as g
from gcn_string_b200 import _lib, synthetic
lib = _lib.load()
ds = synthetic.make_dataset(1024, seed=0, n_mean=500, deg=12, n_feat=32)
loader = g.DisjointLoader(ds, batch_size=1024, epochs=None, shuffle=False, symmetric=True, device_resident=True)
model = g.GeneralGNN(2, activation="softmax", hidden=256, message_passing=4, seed=0); model.build(32)
(x, a, i), y = next(loader)
for training in (False, True):
    for _ in range(3):
        model((x, a, i), training=training) if not training else model.train_step_grads((x, a, i), y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model((x, a, i), training=training) if not training else model.train_step_grads((x, a, i), y)
    e1.record(); torch.cuda.synchronize()
    print("training", training, "ms", e0.elapsed_time(e1) / 5)
    _lib.profile_begin()
    model((x, a, i), training=training) if not training else model.train_step_grads((x, a, i), y)
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    print({k: (c, round(ms / c, 3)) for k, (c, ms) in prof.items()})
