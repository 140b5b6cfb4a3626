import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_err(got, ref):
    """max |got - ref| / max |ref| (the '1e-5 relative' of BASELINE.json, per tensor)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.abs(ref).max() if ref.size else 0.0
    num = np.abs(got - ref).max() if ref.size else 0.0
    return num / den if den > 0 else num


@pytest.fixture(scope="session")
def small_case():
    """8 small synthetic graphs + a perturbed tiny model, collated by the oracle."""
    import gcn_string_b200 as g
    from gcn_string_b200 import synthetic
    from oracle import batching_ref
    ds = synthetic.make_dataset(8, seed=3, n_mean=60, deg=8, n_feat=12)
    graphs = [ds.graph(i) for i in range(ds.n_graphs)]
    (x, (idx, vals, shape), seg), y = batching_ref.collate(graphs)
    cfg = g.GNNConfig(in_features=12, output=2, activation="softmax", hidden=16, message_passing=3)
    w, s = g.init_params(cfg, seed=1, perturb=True)
    return dict(ds=ds, graphs=graphs, x=x, idx=idx, seg=seg, y=y, cfg=cfg, specs=g.block_specs(cfg), w=w, s=s)
