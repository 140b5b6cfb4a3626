"""Synthetic "E. coli inter-protein shape" graphs as a packed store.

What the reference's offline generator emits per protein pair, restated as a sampler
(reference: src/utilities/gcn_utills.py):
  * two proteins, one residue per node, node order = protein a then protein b
    (``nx.union`` with a-/b- prefixes, gcn_utills.py:345; relabelled to ints at
    src/scripts/gcn.py:189);
  * per protein a symmetric CA-CA contact graph WITH self-loops (10 A threshold on a
    distance matrix whose diagonal is 0, gcn_utills.py:225-227,256-257) - here a band
    |i-j| <= w plus long-range contacts with P(d) ~ 1/d;
  * 20 symmetric inter-protein DCA "bridge" edges (gcn_utills.py:87,345-351);
  * adjacency values are all 1 (src/scripts/gcn.py:187-197 pops ``weight``);
  * node features: 16 NetSurfP columns in the real data (gcn_utills.py:293-311),
    mostly probabilities -> U[0,1); width F is a parameter (BASELINE.json uses 32);
  * labels one-hot [0,1] positive / [1,0] negative, balanced (src/scripts/gcn.py:259,262).

Everything is NumPy with ``default_rng(seed + graph_id)`` so any graph can be
regenerated alone.  The output is the packed layout the device batching kernel consumes
(``PackedGraphs``): per-graph CSR with LOCAL column indices, concatenated.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class PackedGraphs:
    """Concatenated per-graph CSR + features + labels (host, NumPy).

    node_off[G+1]  int64  node offset of each graph in ``x`` / ``rowptr``
    rowptr[sumN+1] int64  start of each node's neighbour list in ``col`` (dataset-global)
    col[sumNNZ]    int32  neighbour index LOCAL to the graph, ascending within a row
    x[sumN, F]     float32
    y[G, C]        float32 one-hot
    """
    node_off: np.ndarray
    rowptr: np.ndarray
    col: np.ndarray
    x: np.ndarray
    y: np.ndarray

    @property
    def n_graphs(self) -> int:
        return int(self.node_off.shape[0] - 1)

    @property
    def n_nodes(self) -> np.ndarray:
        return np.diff(self.node_off)

    @property
    def n_edges(self) -> np.ndarray:
        return self.rowptr[self.node_off[1:]] - self.rowptr[self.node_off[:-1]]

    def graph(self, g: int):
        """(x[n,F], scipy CSR int64 ones [n,n], y[C]) for graph g — the per-graph objects
        the reference's MyDataset hands to Spektral (src/scripts/gcn.py:161-181)."""
        import scipy.sparse as sp
        n0, n1 = int(self.node_off[g]), int(self.node_off[g + 1])
        e0, e1 = int(self.rowptr[n0]), int(self.rowptr[n1])
        n = n1 - n0
        a = sp.csr_matrix((np.ones(e1 - e0, dtype=np.int64), self.col[e0:e1].astype(np.int32),
                           (self.rowptr[n0:n1 + 1] - e0).astype(np.int32)), shape=(n, n))
        return self.x[n0:n1], a, self.y[g]


def _protein_pairs(rng, n, deg):
    """Undirected contact keys i*n+j (i<j) of one protein: band + 1/d long-range."""
    lr_per_node = 3 if deg >= 8 else 1
    w = max(1, (deg - 1 - lr_per_node) // 2)
    w = min(w, max(1, n - 1))
    i = np.repeat(np.arange(n, dtype=np.int64), w)
    d = np.tile(np.arange(1, w + 1, dtype=np.int64), n)
    keep = i + d < n
    bi, bj = i[keep], (i + d)[keep]
    band_nnz = n + 2 * bi.size
    m = max(0, int(round((deg * n - band_nnz) / 2.0)))
    if m > 0 and n > w + 2:
        u = rng.random(m)
        dd = np.floor((w + 1) * np.exp(u * np.log(n / (w + 1.0)))).astype(np.int64)
        dd = np.clip(dd, w + 1, n - 1)
        li = np.floor(rng.random(m) * (n - dd)).astype(np.int64)
        lj = li + dd
        bi = np.concatenate([bi, li])
        bj = np.concatenate([bj, lj])
    return bi, bj


def make_graph(seed: int, graph_id: int, n_mean: int = 500, deg: int = 12, n_feat: int = 32,
               n_bridges: int = 20):
    """One graph -> (rowptr_local int64 [n+1], col_local int32 [nnz], x f32 [n,F])."""
    rng = np.random.default_rng(seed + graph_id)
    half = n_mean / 2.0
    sigma = 0.35
    mu = np.log(half) - 0.5 * sigma * sigma
    lens = np.clip(np.rint(rng.lognormal(mu, sigma, size=2)), 15, 2 * n_mean).astype(np.int64)
    n1, n2 = int(lens[0]), int(lens[1])
    n = n1 + n2
    ai, aj = _protein_pairs(rng, n1, deg)
    bi, bj = _protein_pairs(rng, n2, deg)
    br_i = rng.integers(0, n1, size=n_bridges)
    br_j = rng.integers(0, n2, size=n_bridges) + n1
    ui = np.concatenate([ai, bi + n1, br_i])
    uj = np.concatenate([aj, bj + n1, br_j])
    diag = np.arange(n, dtype=np.int64)
    rows = np.concatenate([ui, uj, diag])
    cols = np.concatenate([uj, ui, diag])
    keys = np.unique(rows * n + cols)          # sorted row-major, duplicates merged
    r = keys // n
    c = (keys - r * n).astype(np.int32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=n), out=rowptr[1:])
    x = rng.random((n, n_feat), dtype=np.float32)
    return rowptr, c, x


def make_dataset(n_graphs: int, seed: int = 0, n_mean: int = 500, deg: int = 12,
                 n_feat: int = 32, n_classes: int = 2) -> PackedGraphs:
    """G(seed, n_mean, deg, F) of SURVEY.md §8d, packed."""
    rps: List[np.ndarray] = []
    cols: List[np.ndarray] = []
    xs: List[np.ndarray] = []
    node_off = np.zeros(n_graphs + 1, dtype=np.int64)
    e_off = 0
    for g in range(n_graphs):
        rp, c, x = make_graph(seed, g, n_mean, deg, n_feat)
        node_off[g + 1] = node_off[g] + x.shape[0]
        rps.append(rp[:-1] + e_off)
        e_off += int(rp[-1])
        cols.append(c)
        xs.append(x)
    rowptr = np.concatenate(rps + [np.array([e_off], dtype=np.int64)]) if n_graphs else np.zeros(1, np.int64)
    col = np.concatenate(cols) if n_graphs else np.zeros(0, np.int32)
    x = np.concatenate(xs) if n_graphs else np.zeros((0, n_feat), np.float32)
    y = np.zeros((n_graphs, n_classes), dtype=np.float32)
    lab = np.arange(n_graphs) % n_classes          # balanced (gcn.py:265-266)
    y[np.arange(n_graphs), lab] = 1.0
    return PackedGraphs(node_off, rowptr, col, x, y)


def concat_packed(parts: List[PackedGraphs]) -> PackedGraphs:
    """Concatenate packed datasets (graph order = order of the parts)."""
    if len(parts) == 1:
        return parts[0]
    node_off = [np.zeros(1, dtype=np.int64)]
    rowptr = []
    n0, e0 = 0, 0
    for p in parts:
        node_off.append(p.node_off[1:] + n0)
        rowptr.append(p.rowptr[:-1] + e0)
        n0 += int(p.node_off[-1])
        e0 += int(p.rowptr[-1])
    rowptr.append(np.array([e0], dtype=np.int64))
    return PackedGraphs(np.concatenate(node_off), np.concatenate(rowptr), np.concatenate([p.col for p in parts]),
                        np.concatenate([p.x for p in parts]), np.concatenate([p.y for p in parts]))


def _make_range(args):
    seed, g0, g1, n_mean, deg, n_feat, n_classes = args
    rps, cols, xs = [], [], []
    node_off = np.zeros(g1 - g0 + 1, dtype=np.int64)
    e_off = 0
    for k, gid in enumerate(range(g0, g1)):
        rp, c, x = make_graph(seed, gid, n_mean, deg, n_feat)
        node_off[k + 1] = node_off[k] + x.shape[0]
        rps.append(rp[:-1] + e_off)
        e_off += int(rp[-1])
        cols.append(c)
        xs.append(x)
    y = np.zeros((g1 - g0, n_classes), dtype=np.float32)
    y[np.arange(g1 - g0), np.arange(g0, g1) % n_classes] = 1.0
    return PackedGraphs(node_off, np.concatenate(rps + [np.array([e_off], dtype=np.int64)]), np.concatenate(cols),
                        np.concatenate(xs), y)


def make_dataset_parallel(n_graphs: int, seed: int = 0, n_mean: int = 500, deg: int = 12, n_feat: int = 32,
                          n_classes: int = 2, workers: int = 0) -> PackedGraphs:
    """make_dataset over a process pool (the same graphs: every graph is seeded by its id).  workers = 0: the cores this
    process may run on."""
    import multiprocessing as mp
    import os
    workers = workers or len(os.sched_getaffinity(0))
    if workers <= 1 or n_graphs < 4 * workers:
        return make_dataset(n_graphs, seed, n_mean, deg, n_feat, n_classes)
    chunk = -(-n_graphs // (4 * workers))
    jobs = [(seed, g0, min(g0 + chunk, n_graphs), n_mean, deg, n_feat, n_classes) for g0 in range(0, n_graphs, chunk)]
    with mp.get_context("fork").Pool(workers) as pool:
        parts = pool.map(_make_range, jobs)
    return concat_packed(parts)


def tile_dataset(packed: PackedGraphs, n_graphs: int) -> PackedGraphs:
    """The first n_graphs graphs of `packed` repeated end to end (a large epoch out of a smaller set of unique
    synthetic graphs; labels and features repeat with them)."""
    g = packed.n_graphs
    reps = -(-n_graphs // g)
    out = concat_packed([packed] * reps)
    if out.n_graphs == n_graphs:
        return out
    n1 = int(out.node_off[n_graphs])
    e1 = int(out.rowptr[n1])
    return PackedGraphs(out.node_off[:n_graphs + 1], out.rowptr[:n1 + 1], out.col[:e1], out.x[:n1], out.y[:n_graphs])


def pack_graphs(graphs) -> PackedGraphs:
    """Pack a list of per-graph objects exposing ``.x [n,F]``, ``.a`` (scipy sparse / dense
    [n,n]) and ``.y`` — what the reference's MyDataset.read returns (gcn.py:84-102) — into
    the device-ready layout.  Explicit zeros are dropped and duplicates merged, exactly as
    ``sp.find`` does in Spektral's collate; values are otherwise ignored
    (a5: GeneralConv never reads them)."""
    import scipy.sparse as sp
    node_off = np.zeros(len(graphs) + 1, dtype=np.int64)
    rps, cols, xs, ys = [], [], [], []
    e_off = 0
    for g, gr in enumerate(graphs):
        a = sp.csr_matrix(gr.a, copy=True)     # never mutate the caller's matrix
        a.sum_duplicates()
        a.eliminate_zeros()
        a.sort_indices()
        n = a.shape[0]
        x = np.asarray(gr.x)
        if x.shape[0] != n:
            raise ValueError(f"graph {g}: x has {x.shape[0]} rows but a is {a.shape}")
        node_off[g + 1] = node_off[g] + n
        rps.append(a.indptr[:-1].astype(np.int64) + e_off)
        e_off += int(a.indptr[-1])
        cols.append(a.indices.astype(np.int32))
        xs.append(x.astype(np.float32))
        ys.append(np.asarray(gr.y, dtype=np.float32).reshape(-1))
    rowptr = np.concatenate(rps + [np.array([e_off], dtype=np.int64)]) if graphs else np.zeros(1, np.int64)
    col = np.concatenate(cols) if graphs else np.zeros(0, np.int32)
    x = np.concatenate(xs) if graphs else np.zeros((0, 0), np.float32)
    y = np.stack(ys) if graphs else np.zeros((0, 0), np.float32)
    return PackedGraphs(node_off, rowptr, col, x, y)
