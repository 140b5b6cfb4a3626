"""CPU tests of the §8(f) rows: the packed shard ingest format (shards.py) and the ROC metrics
(evaluate.py) against the libraries the reference calls (networkx adjacency, scikit-learn)."""
import os

import numpy as np
import pytest

import gcn_string_b200 as g
from gcn_string_b200 import shards, synthetic
from gcn_string_b200.evaluate import auc, roc_auc, roc_curve


def _same(a, b):
    return all(np.array_equal(getattr(a, k), getattr(b, k)) for k in ("node_off", "rowptr", "col", "x", "y"))


def test_shard_round_trip_mmap_and_checksum(tmp_path):
    ds = synthetic.make_dataset(9, seed=3, n_mean=40, deg=8, n_feat=5)
    path = shards.write_shard(str(tmp_path / "a.gcss"), ds)
    h = shards.read_header(path)
    assert (h.n_graphs, h.n_nodes, h.nnz, h.n_feat, h.n_classes) == (9, int(ds.node_off[-1]), ds.col.shape[0], 5, 2)
    assert h.symmetric                                     # undirected contact graphs
    for mmap in (True, False):
        back = shards.read_shard(path, mmap=mmap, verify=True)
        assert _same(back, ds) and back.symmetric
        assert back.x.dtype == np.float32 and back.col.dtype == np.int32 and back.rowptr.dtype == np.int64
    assert not shards.read_shard(path).x.flags.owndata      # a view of the mapped file, not a parsed copy
    # every section starts on a 64-byte boundary
    secs, end = shards._sections(h.n_graphs, h.n_nodes, h.nnz, h.n_feat, h.n_classes)
    assert all(off % 64 == 0 for _, _, _, off in secs) and end == os.path.getsize(path)


def test_shard_rejects_corruption(tmp_path):
    ds = synthetic.make_dataset(3, seed=1, n_mean=30, deg=8, n_feat=4)
    path = shards.write_shard(str(tmp_path / "a.gcss"), ds)
    raw = bytearray(open(path, "rb").read())
    bad = str(tmp_path / "bad.gcss")
    open(bad, "wb").write(b"NOTASHRD" + raw[8:])
    with pytest.raises(ValueError, match="bad magic"):
        shards.read_shard(bad)
    open(bad, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(ValueError, match="truncated"):
        shards.read_shard(bad)
    flipped = bytearray(raw)
    h = shards.read_header(path)
    x_off = [off for name, _, _, off in shards._sections(h.n_graphs, h.n_nodes, h.nnz, h.n_feat, h.n_classes)[0] if name == "x"][0]
    flipped[x_off + 5] ^= 0xFF
    open(bad, "wb").write(flipped)
    with pytest.raises(ValueError, match="checksum"):
        shards.read_shard(bad, verify=True)
    v2 = bytearray(raw)
    v2[8] = 2
    open(bad, "wb").write(v2)
    with pytest.raises(ValueError, match="version"):
        shards.read_shard(bad)
    broken = synthetic.PackedGraphs(ds.node_off, ds.rowptr, ds.col.copy(), ds.x, ds.y)
    broken.col[0] = 10 ** 6
    with pytest.raises(ValueError, match="outside its graph"):
        shards.write_shard(bad, broken)


def test_write_dataset_shards_and_reload_in_order(tmp_path):
    ds = synthetic.make_dataset(11, seed=5, n_mean=30, deg=8, n_feat=3)
    graphs = [g.Graph(*ds.graph(k)[:2], y=ds.graph(k)[2]) for k in range(11)]
    paths = shards.write_dataset(graphs, str(tmp_path / "out"), graphs_per_shard=4)
    assert [os.path.basename(p) for p in paths] == ["shard-00000.gcss", "shard-00001.gcss", "shard-00002.gcss"]
    assert [shards.read_header(p).n_graphs for p in paths] == [4, 4, 3]
    back = shards.load_dataset(str(tmp_path / "out"), verify=True)
    assert _same(back, ds) and back.symmetric
    assert _same(shards.load_dataset(paths[1:2]), synthetic.pack_graphs(graphs[4:8]))
    with pytest.raises(ValueError, match="empty"):
        shards.write_dataset([], str(tmp_path / "none"))


def test_asymmetric_pattern_is_flagged(tmp_path):
    import scipy.sparse as sp
    a = sp.csr_matrix(np.array([[1, 1, 0], [0, 1, 0], [0, 1, 1]]))
    p = synthetic.pack_graphs([g.Graph(x=np.ones((3, 2)), a=a, y=np.array([0.0, 1.0]))])
    assert not shards.pattern_is_symmetric(p)
    path = shards.write_shard(str(tmp_path / "d.gcss"), p)
    assert not shards.read_header(path).symmetric and not shards.read_shard(path).symmetric


def test_graph_from_networkx_matches_the_reference_ingest():
    """format_graph + generate_spektral_graph of gcn.py:161-197 on a toy protein-pair graph: string
    node names, a 'weight' attribute networkx added, a self-loop, features under 'x'."""
    nx = pytest.importorskip("networkx")
    rng = np.random.default_rng(0)
    G = nx.Graph()
    names = ["a-0", "a-1", "a-2", "b-0", "b-1"]
    for n in names:
        G.add_node(n, x=rng.random(4))
    G.add_edge("a-0", "a-1", weight=3.0, dca=0.1)
    G.add_edge("a-1", "a-2", dca=0.2)
    G.add_edge("a-2", "b-0", weight=0.5, dca=0.9)      # DCA bridge
    G.add_edge("b-0", "b-1", dca=0.3)
    G.add_edge("a-0", "a-0", dca=0.0)                  # self-loop (diagonal of the contact map)
    gr = shards.graph_from_networkx(G, [0, 1])
    F = nx.convert_node_labels_to_integers(G)
    for _, _, d in F.edges(data=True):
        d.pop("weight", None)
    want_a = nx.adjacency_matrix(F)
    assert np.array_equal(gr.a.toarray(), want_a.toarray())
    assert np.array_equal(gr.x, np.vstack([x[1] for x in F.nodes.data("x")]))
    assert np.array_equal(gr.y, np.array([0, 1]))
    p = synthetic.pack_graphs([gr])
    assert shards.pattern_is_symmetric(p) and p.col.shape[0] == want_a.nnz


@pytest.mark.parametrize("seed,n,ties", [(0, 50, False), (1, 400, True), (2, 7, True), (3, 1000, False)])
def test_roc_curve_and_auc_match_scikit_learn(seed, n, ties):
    skm = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(seed)
    y = rng.integers(0, 2, n)
    y[0], y[1] = 0, 1
    s = rng.random(n) * 0.6 + 0.3 * y
    if ties:
        s = np.round(s, 1)
    for drop in (True, False):
        fpr, tpr, thr = roc_curve(y, s, drop_intermediate=drop)
        f0, t0, h0 = skm.roc_curve(y, s, drop_intermediate=drop)
        assert np.allclose(fpr, f0, atol=1e-12) and np.allclose(tpr, t0, atol=1e-12) and np.array_equal(thr[1:], h0[1:])
    assert abs(roc_auc(y, s) - skm.roc_auc_score(y, s)) < 1e-12
    fpr, tpr, _ = roc_curve(y, s)
    assert abs(auc(fpr, tpr) - skm.auc(fpr, tpr)) < 1e-12
    # the reference calls auc(tpr, fpr) (gcn.py:397): same swapped-argument value
    assert abs(auc(tpr, fpr) - skm.auc(tpr, fpr)) < 1e-12
    with pytest.raises(ValueError):
        auc([0.0, 1.0, 0.5], [0.0, 1.0, 1.0])
    with pytest.raises(ValueError):
        auc([0.0], [1.0])


def test_bucketed_batch_tensor_sizes():
    """ops.bucket_rows (the allocation size behind every per-batch tensor): never less than asked for, at most 12.5 % more,
    and sizes that differ by a per cent or two - consecutive batches of a loader - share one bucket almost always."""
    from gcn_string_b200.ops import bucket_rows
    rng = np.random.default_rng(0)
    for n in [0, 1, 4095, 4096, 4097, 508180, 6120690, 2**24 - 1] + rng.integers(1, 10**8, 200).tolist():
        b = bucket_rows(n)
        assert b >= n and (n < 4096 and b == n or b <= n * 1.125 + 1)
        assert bucket_rows(b) == b                                   # a bucket boundary maps to itself
    sizes = (508180 * (1 + 0.008 * rng.standard_normal(2000))).astype(np.int64)      # cfg2: rows of a batch, +-0.8 %
    assert len({bucket_rows(int(v)) for v in sizes}) <= 2
    nnz = (6120690 * (1 + 0.008 * rng.standard_normal(2000))).astype(np.int64)
    assert len({bucket_rows(int(v)) for v in nnz}) <= 2
