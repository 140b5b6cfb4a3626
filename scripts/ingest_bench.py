"""Ingest measurement: per-graph objects -> packed arrays (what every run of the reference's
MyDataset.read + Spektral collate front-end has to do) versus opening packed shards.

    python scripts/ingest_bench.py [n_graphs] [out.json]
Host-only (no GPU needed); the upload of a packed dataset is one contiguous H2D copy per array."""
import json, os, pickle, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gcn_string_b200 as g
from gcn_string_b200 import shards, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
ds = synthetic.make_dataset(n, seed=0, n_mean=500, deg=12, n_feat=32)
graphs = [g.Graph(*ds.graph(k)[:2], y=ds.graph(k)[2]) for k in range(n)]
nbytes = sum(a.nbytes for a in (ds.node_off, ds.rowptr, ds.col, ds.x, ds.y))
out = {"n_graphs": n, "nodes": int(ds.node_off[-1]), "nnz": int(ds.col.shape[0]), "packed_bytes": int(nbytes)}
with tempfile.TemporaryDirectory() as d:
    # reference-style storage: one pickle per graph (the reference uses one .gpickle per protein pair)
    t0 = time.perf_counter()
    for k, gr in enumerate(graphs):
        with open(os.path.join(d, f"g{k}.pkl"), "wb") as f:
            pickle.dump((gr.x, gr.a, gr.y), f, protocol=4)
    out["write_pickles_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    loaded = []
    for k in range(n):
        with open(os.path.join(d, f"g{k}.pkl"), "rb") as f:
            x, a, y = pickle.load(f)
        loaded.append(g.Graph(x=x, a=a, y=y))
    packed = synthetic.pack_graphs(loaded)
    out["per_graph_objects_to_packed_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    paths = shards.write_dataset(graphs, os.path.join(d, "sh"), graphs_per_shard=1024)
    out["write_shards_s"] = time.perf_counter() - t0
    out["shard_files"] = len(paths)
    t0 = time.perf_counter()
    p = shards.load_dataset(os.path.join(d, "sh"))
    out["open_shards_mmap_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    touched = [np.ascontiguousarray(a) for a in (p.node_off, p.rowptr, p.col, p.x, p.y)]     # page everything in
    out["open_and_read_shards_s"] = out["open_shards_mmap_s"] + (time.perf_counter() - t0)
    t0 = time.perf_counter()
    shards.load_dataset(os.path.join(d, "sh"), verify=True)
    out["open_with_checksum_s"] = time.perf_counter() - t0
    assert all(np.array_equal(getattr(p, k), getattr(packed, k)) for k in ("node_off", "rowptr", "col", "x", "y"))
# the reference's own route for a sample: one pickled networkx graph per pair -> format_graph ->
# adjacency / node features (gcn.py:161-197) -> Graph
try:
    import networkx as nx
    m = min(n, 200)
    with tempfile.TemporaryDirectory() as d:
        for k in range(m):
            x, a, y = ds.graph(k)
            G = nx.from_scipy_sparse_array(a)
            for i in range(x.shape[0]):
                G.nodes[i]["x"] = x[i]
            G = nx.relabel_nodes(G, {i: f"r{i}" for i in G.nodes})
            with open(os.path.join(d, f"g{k}.gpickle"), "wb") as f:
                pickle.dump(G, f, protocol=4)
        t0 = time.perf_counter()
        gl = []
        for k in range(m):
            with open(os.path.join(d, f"g{k}.gpickle"), "rb") as f:
                G = pickle.load(f)
            gl.append(shards.graph_from_networkx(G, ds.y[k]))
        synthetic.pack_graphs(gl)
        out["networkx_route_sample_graphs"] = m
        out["networkx_route_s"] = time.perf_counter() - t0
        out["graphs_per_s_networkx_route"] = m / out["networkx_route_s"]
except ImportError:
    pass
out["graphs_per_s_per_graph_objects"] = n / out["per_graph_objects_to_packed_s"]
out["graphs_per_s_shards"] = n / out["open_and_read_shards_s"]
out["read_GBps_shards"] = nbytes / out["open_and_read_shards_s"] / 1e9
out["speedup"] = out["per_graph_objects_to_packed_s"] / out["open_and_read_shards_s"]
print(json.dumps(out))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
