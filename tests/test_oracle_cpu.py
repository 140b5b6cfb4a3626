"""CPU suite: the oracle against itself / scipy / the golden vectors, the host logic, and
the C-ABI library's exported symbols.  No GPU compute here."""
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import gcn_string_b200 as g
from gcn_string_b200 import synthetic
from gcn_string_b200.params import GNNConfig, block_specs, n_state, n_trainable, named_slices
from oracle import batching_ref, model_ref_np as O1, model_ref_torch as O2

from conftest import ROOT, rel_err

GOLDEN = os.path.join(ROOT, "tests", "golden")


# ---------------------------------------------------------------- parameters / layout
def test_default_parameter_count_matches_survey():
    # 1 064 454 trainable + 3 588 non-trainable at F=32; 1 060 358 at F=16
    cfg = GNNConfig(in_features=32, output=2, activation="softmax")
    assert n_trainable(cfg) == 1_064_454 and n_state(cfg) == 3_588
    assert n_trainable(GNNConfig(in_features=16, output=2, activation="softmax")) == 1_060_358


def test_layout_is_dense_and_ordered():
    cfg = GNNConfig(in_features=7, output=3, hidden=12, message_passing=3, pre_process=1, post_process=3)
    off = {"trainable": 0, "state": 0}
    for name, shape, o, buf in named_slices(cfg):
        assert o == off[buf], name
        off[buf] += int(np.prod(shape))
    assert off["trainable"] == n_trainable(cfg) and off["state"] == n_state(cfg)
    widths = [b.k_in for b in block_specs(cfg) if b.name.startswith("gnn")]
    assert widths == [12, 24, 36]                       # 'cat' grows the conv input (a4)
    assert [b.has_alpha for b in block_specs(cfg)][-1] is False


@pytest.mark.parametrize("kw", [dict(aggregate="prod"), dict(aggregate="max", connectivity="sum"), dict(dropout=0.5), dict(hidden_activation="relu"),
                                dict(pool="max"), dict(batch_norm=False)])
def test_unsupported_configurations_raise(kw):
    with pytest.raises(NotImplementedError):
        GNNConfig(in_features=4, output=2, **kw).validate()


# ---------------------------------------------------------------- synthetic generator
def test_synthetic_graph_shape():
    ds = synthetic.make_dataset(16, seed=0, n_mean=500, deg=12, n_feat=32)
    assert ds.x.dtype == np.float32 and ds.x.shape[1] == 32
    assert 350 < ds.n_nodes.mean() < 650
    assert 10.5 < (ds.n_edges / ds.n_nodes).mean() < 13.5
    for gi in (0, 7):
        x, a, y = ds.graph(gi)
        assert (a != a.T).nnz == 0                      # symmetric (undirected nx graph)
        assert np.all(a.diagonal() == 1)                # self-loops (gcn_utills.py:225-227)
        assert a.has_sorted_indices and a.data.dtype == np.int64 and np.all(a.data == 1)
        assert y.sum() == 1
    assert ds.y.sum(0).tolist() == [8, 8]               # balanced labels (gcn.py:265-266)
    again = synthetic.make_dataset(16, seed=0, n_mean=500, deg=12, n_feat=32)
    assert np.array_equal(ds.col, again.col) and np.array_equal(ds.x, again.x)


def test_pack_graphs_roundtrip_and_find_semantics():
    rng = np.random.default_rng(0)
    graphs = []
    for n in (1, 5, 9):
        d = (rng.random((n, n)) < 0.4).astype(np.int64)
        a = sp.csr_matrix(d)
        if n == 5:                                      # explicit zero must be dropped (sp.find)
            a = sp.csr_matrix((np.array([0, 1]), (np.array([0, 1]), np.array([1, 2]))), shape=(5, 5))
        graphs.append(g.Graph(x=rng.random((n, 3)), a=a, y=np.array([1, 0])))
    packed = synthetic.pack_graphs(graphs)
    assert packed.n_nodes.tolist() == [1, 5, 9]
    x1, a1, _ = packed.graph(1)
    assert a1.nnz == 1 and a1[1, 2] == 1
    (x, (idx, _, _), seg), y = batching_ref.collate([packed.graph(i) for i in range(3)])
    (x2, (idx2, _, _), seg2), _ = batching_ref.collate([(gr.x, gr.a, gr.y) for gr in graphs])
    assert np.array_equal(idx, idx2) and np.array_equal(seg, seg2)


# ---------------------------------------------------------------- O3: scipy collate
def test_collate_is_row_major_block_diagonal(small_case):
    c = small_case
    idx, seg = c["idx"], c["seg"]
    n = c["x"].shape[0]
    assert idx.dtype == np.int64 and seg.dtype == np.int64
    keys = idx[:, 0] * n + idx[:, 1]
    assert np.all(np.diff(keys) > 0)                    # canonical order, no duplicates
    assert np.array_equal(seg[idx[:, 0]], seg[idx[:, 1]])   # no edge crosses graphs
    rowptr, colidx, deg = batching_ref.derived_csr(idx, n)
    dense = sp.block_diag([gr[1] for gr in c["graphs"]]).toarray()
    assert np.array_equal(deg, (dense != 0).sum(1))
    gp = batching_ref.graph_ptr(seg, len(c["graphs"]))
    assert np.array_equal(np.diff(gp), c["ds"].n_nodes)
    assert batching_ref.batch_slices(10, 4) == [(0, 4), (4, 8), (8, 10)]


# ---------------------------------------------------------------- O1 vs O2, finite differences
def test_float64_and_autograd_restatements_agree(small_case):
    c = small_case
    args = (c["cfg"], c["specs"], c["w"], c["s"], c["x"], c["idx"][:, 0], c["idx"][:, 1], c["seg"], c["y"], 8)
    r1 = O1.loss_and_grads(*args)
    r2 = O2.loss_and_grads(*args)
    assert abs(r1["loss"] - r2["loss"]) < 1e-5 * abs(r1["loss"])
    assert rel_err(r2["probs"], r1["probs"]) < 1e-5
    assert rel_err(r2["grads"], r1["grads"]) < 2e-5    # fp32 autograd noise floor
    for (m2, v2), cache in zip(r2["stats"], r1["ctx"]["caches"]):
        assert rel_err(m2.numpy(), cache["mean"]) < 1e-5 and rel_err(v2.numpy(), cache["var"]) < 1e-5


def test_manual_backward_matches_finite_differences(small_case):
    c = small_case
    cfg, specs, s = c["cfg"], c["specs"], c["s"]
    rows, cols = c["idx"][:, 0], c["idx"][:, 1]
    r = O1.loss_and_grads(cfg, specs, c["w"], s, c["x"], rows, cols, c["seg"], c["y"], 8)

    def loss_at(wv):
        _, ctx = O1.forward(cfg, specs, wv, s, c["x"], rows, cols, c["seg"], 8, training=True)
        return O1.xent_from_logits(ctx["logits"], c["y"].astype(np.float64))[0]

    rng = np.random.default_rng(5)
    scale = np.abs(r["grads"]).max()
    for j in rng.integers(0, c["w"].shape[0], 12):
        wp = c["w"].astype(np.float64).copy()
        wm = wp.copy()
        wp[j] += 1e-6
        wm[j] -= 1e-6
        fd = (loss_at(wp) - loss_at(wm)) / 2e-6
        assert abs(fd - r["grads"][j]) < 1e-6 * scale + 1e-9


@pytest.mark.parametrize("connectivity", ["sum", None])
def test_skip_connection_variants_in_both_restatements(small_case, connectivity):
    """GeneralGNN.call with connectivity='sum' (out = Add()([z, out])) and None (out = z): every layer stays `hidden`
    wide; the hand-derived backward agrees with autograd and with finite differences."""
    c = small_case
    cfg = GNNConfig(in_features=12, output=2, activation="softmax", hidden=16, message_passing=3, connectivity=connectivity)
    specs = block_specs(cfg)
    assert [b.k_in for b in specs if b.name.startswith("gnn")] == [16, 16, 16] and specs[-2].k_in == 16
    w, s = g.init_params(cfg, seed=11, perturb=True)
    rows, cols = c["idx"][:, 0], c["idx"][:, 1]
    args = (cfg, specs, w, s, c["x"], rows, cols, c["seg"], c["y"], 8)
    r1, r2 = O1.loss_and_grads(*args), O2.loss_and_grads(*args)
    assert abs(r1["loss"] - r2["loss"]) < 1e-5 * abs(r1["loss"])
    assert rel_err(r2["grads"], r1["grads"]) < 2e-5

    def loss_at(wv):
        _, ctx = O1.forward(cfg, specs, wv, s, c["x"], rows, cols, c["seg"], 8, training=True)
        return O1.xent_from_logits(ctx["logits"], c["y"].astype(np.float64))[0]

    scale = np.abs(r1["grads"]).max()
    for j in np.random.default_rng(6).integers(0, w.shape[0], 8):
        wp = w.astype(np.float64).copy()
        wm = wp.copy()
        wp[j] += 1e-6
        wm[j] -= 1e-6
        assert abs((loss_at(wp) - loss_at(wm)) / 2e-6 - r1["grads"][j]) < 1e-6 * scale + 1e-9
    # the three variants are different functions of the same inputs
    cat = GNNConfig(in_features=12, output=2, activation="softmax", hidden=16, message_passing=3)
    assert n_trainable(cat) > n_trainable(cfg)


@pytest.mark.parametrize("aggregate,weighted,connectivity", [("mean", False, "cat"), ("max", False, "cat"), ("sum", True, "cat"),
                                                           ("mean", True, None), ("max", True, None), ("sum", True, "sum")])
def test_general_aggregation_backward_matches_finite_differences(small_case, aggregate, weighted, connectivity):
    """scatter_mean / scatter_max and per-entry weights in the oracle: the hand-derived backward -
    max shares a row's gradient between the entries that attain the maximum, like tf's unsorted_segment_max gradient -
    against central finite differences of the float64 forward."""
    c = small_case
    cfg = GNNConfig(in_features=12, output=2, activation="softmax", hidden=16, message_passing=2, connectivity=connectivity,
                    aggregate=aggregate)
    specs = block_specs(cfg)
    w, s = g.init_params(cfg, seed=21, perturb=True)
    rows, cols = c["idx"][:, 0], c["idx"][:, 1]
    ew = np.random.default_rng(4).uniform(0.2, 1.5, rows.shape[0]) if weighted else None
    r = O1.loss_and_grads(cfg, specs, w, s, c["x"], rows, cols, c["seg"], c["y"], 8, edge_weight=ew)

    def loss_at(wv):
        _, ctx = O1.forward(cfg, specs, wv, s, c["x"], rows, cols, c["seg"], 8, training=True, edge_weight=ew)
        return O1.xent_from_logits(ctx["logits"], c["y"].astype(np.float64))[0]

    scale = np.abs(r["grads"]).max()
    for j in np.random.default_rng(8).integers(0, w.shape[0], 10):
        wp = w.astype(np.float64).copy()
        wm = wp.copy()
        wp[j] += 1e-6
        wm[j] -= 1e-6
        assert abs((loss_at(wp) - loss_at(wm)) / 2e-6 - r["grads"][j]) < 2e-6 * scale + 1e-9


def test_inference_uses_moving_statistics(small_case):
    c = small_case
    rows, cols = c["idx"][:, 0], c["idx"][:, 1]
    p_inf, _ = O1.forward(c["cfg"], c["specs"], c["w"], c["s"], c["x"], rows, cols, c["seg"], 8, training=False)
    p_tr, _ = O1.forward(c["cfg"], c["specs"], c["w"], c["s"], c["x"], rows, cols, c["seg"], 8, training=True)
    assert np.allclose(p_inf.sum(1), 1) and np.abs(p_inf - p_tr).max() > 1e-4


def test_optimizer_restatements():
    # gcn.py:321-325 with epochs=5: boundaries [0, 1] -> 0.02, 0.002, then 0.0002
    assert [O1.piecewise_constant(s, [0, 1], [0.02, 0.002, 0.0002]) for s in range(4)] == [0.02, 0.002, 0.0002, 0.0002]
    sched = g.optimizers.schedules.PiecewiseConstantDecay([0, 1], [0.02, 0.002, 0.0002])
    assert [sched(s) for s in range(4)] == [0.02, 0.002, 0.0002, 0.0002]
    w = np.array([1.0, -2.0]); gr = np.array([0.5, 0.25])
    assert np.allclose(O1.sgd_step(w, gr, 0.1), [0.95, -2.025])
    w1, m, v = O1.adam_step(w, gr, np.zeros(2), np.zeros(2), 1, 0.001)
    assert np.allclose(w1, w - 0.001 * np.sign(gr), atol=1e-6)   # first Adam step = lr * sign(g)


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("name", ["tiny_h8", "small_h32"])
def test_oracle_reproduces_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    F, C, H, L = (int(v) for v in z["cfg"])
    cfg = GNNConfig(in_features=F, output=C, activation="softmax", hidden=H, message_passing=L)
    specs = block_specs(cfg)
    nb = z["y"].shape[0]
    r = O1.loss_and_grads(cfg, specs, z["w"], z["s"], z["x"], z["indices"][:, 0], z["indices"][:, 1], z["seg"],
                          z["y"], nb)
    assert rel_err(r["grads"], z["grads"]) < 1e-12 and abs(r["loss"] - float(z["loss"])) < 1e-12
    assert rel_err(r["new_state"], z["new_state"]) < 1e-12
    # the fixture's integer structure is what the real scipy collate produces from the packed graphs
    ds = synthetic.PackedGraphs(z["node_off"], z["ds_rowptr"], z["ds_col"], z["ds_x"], z["ds_y"])
    (x, (idx, _, _), seg), y = batching_ref.collate([ds.graph(i) for i in range(nb)])
    assert np.array_equal(idx, z["indices"]) and np.array_equal(seg, z["seg"])
    rowptr, colidx, _ = batching_ref.derived_csr(idx, x.shape[0])
    assert np.array_equal(rowptr, z["rowptr"]) and np.array_equal(colidx, z["colidx"])
    r2 = O2.loss_and_grads(cfg, specs, z["w"], z["s"], z["x"], z["indices"][:, 0], z["indices"][:, 1], z["seg"],
                           z["y"], nb)
    assert rel_err(r2["grads"], z["grads"]) < 5e-5


# ---------------------------------------------------------------- C ABI surface
def _declared_functions():
    text = open(os.path.join(ROOT, "include", "gcnstring_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gcn_string_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/gcnstring_b200.h but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
    assert lib.gcs_version() == 100
    # size queries are host-only and must agree with the Python layout
    cfg = GNNConfig(in_features=32, output=2, activation="softmax")
    c = _lib.model_config(cfg)
    assert lib.gcs_model_num_params(c) == n_trainable(cfg) and lib.gcs_model_num_state(c) == n_state(cfg)
    assert lib.gcs_model_workspace_bytes(c, 1000, 12000, 4, 1) > lib.gcs_model_workspace_bytes(c, 1000, 12000, 4, 0) > 0


def test_ctypes_prototypes_match_the_header_signatures():
    """Every declaration of include/gcnstring_b200.h against its entry in _lib.PROTOTYPES: same number of parameters,
    and per parameter the same class (pointer / 32-bit / 64-bit integer / float / double) - a drifted binding would
    corrupt the stack silently."""
    import ctypes
    from gcn_string_b200 import _lib
    text = open(os.path.join(ROOT, "include", "gcnstring_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = re.findall(r"\b(?:int|int32_t|int64_t|const char\s*\*|void)\s+(gcs_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(decls) >= 30

    def klass_c(param):
        param = " ".join(param.split())
        if param in ("void", ""):
            return None
        if "*" in param or param.startswith("gcs_stream") or param.startswith("gcs_allreduce_fn"):
            return "ptr"
        ty = param.rsplit(" ", 1)[0].replace("const ", "")
        return {"int32_t": "i32", "int": "i32", "int64_t": "i64", "float": "f32", "double": "f64"}[ty]

    def klass_py(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
            return "ptr"
        return {ctypes.c_int32: "i32", ctypes.c_int64: "i64", ctypes.c_float: "f32", ctypes.c_double: "f64"}[t]

    for name, params in decls:
        want = [k for k in (klass_c(q) for q in params.split(",")) if k is not None]
        got = [klass_py(t) for t in _lib.PROTOTYPES[name][1]]
        assert got == want, f"{name}: header {want} vs ctypes {got}"


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gcn-string_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_native_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.GeneralGNN(2, activation="softmax")([torch.zeros(3, 4), None, None])
