"""Golden vectors for the contact-map / pair-graph kernels, produced by the REFERENCE's own code.

    python tests/golden/make_contact_golden.py        (needs /root/reference; run in the build container only)

Imports /root/reference/src/utilities/gcn_utills.py unmodified.  Its module-level imports of Bio, seaborn and
matplotlib (absent here, and unused by the functions exercised) are satisfied with empty stub modules; two names the
pinned-era libraries still had are aliased (np.float -> float, nx.from_numpy_matrix -> nx.from_numpy_array, the
renamed equivalent).  Residues are stand-ins exposing ``residue["CA"].coord`` (float32[3]) like Bio.PDB's.  Then
``GraphMaker.generate_proximity_matrix`` (gcn_utills.py:203-238), ``generate_graphs`` (:240-270) and ``link_graphs``
(:319-377) run as written, followed by the formatting gcn.py applies (convert_node_labels_to_integers :184-197,
nx.adjacency_matrix :104-117).  Output: tests/golden/contact_pairs.npz.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/utilities"


class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any


class _Any(metaclass=_AnyMeta):
    """Stands in for any class of an absent third-party module (only ever used as a base class / default value)."""


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return sys.modules.get(self.__name__ + "." + name, _Any)


def load_reference():
    for name in ["Bio", "Bio.PDB", "Bio.PDB.StructureBuilder", "Bio.PDB.Residue", "seaborn", "matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(name, _Stub(name))
    import networkx as nx
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(nx, "from_numpy_matrix"):
        nx.from_numpy_matrix = nx.from_numpy_array
    sys.path.insert(0, REF)
    import gcn_utills
    return gcn_utills


class _Atom:
    def __init__(self, xyz):
        self.coord = np.asarray(xyz, dtype=np.float32)


def chain(rng, n):
    """A CA trace: 3.8 A steps with a persistent direction, float32 like PDB coordinates (3 decimals)."""
    d = rng.normal(size=3)
    pos, out = np.zeros(3), []
    for _ in range(n):
        d = d + 0.9 * rng.normal(size=3)
        d /= np.linalg.norm(d)
        pos = pos + 3.8 * d
        out.append(np.round(pos, 3))
    return np.asarray(out, dtype=np.float32)


def main():
    import networkx as nx
    import scipy.sparse as sp
    ref = load_reference()
    gm = ref.GraphMaker.__new__(ref.GraphMaker)             # __init__ only reads a csv path
    rng = np.random.default_rng(2024)
    lengths = [57, 33, 1, 96, 40, 64]
    cas = [chain(rng, n) for n in lengths]
    cas[4][7] = cas[4][3] + np.float32(10.0) * np.array([1, 0, 0], np.float32)    # a distance of exactly 10 A: not < 10
    adjs, dists = [], []
    for ca in cas:
        seq = [{"CA": _Atom(c)} for c in ca]
        adj, cmap = gm.generate_proximity_matrix(seq, seq, angstroms=10)
        adjs.append(adj)
        dists.append(cmap)
    pairs = [(0, 1), (3, 2), (4, 5), (1, 1)]
    bridges = [[(5, 2), (50, 30), (5, 2), (0, 0), (56, 32), (5, 31)], [(10, 0), (95, 0)], [], [(3, 3), (4, 20)]]
    out = {"lengths": np.asarray(lengths, np.int32), "ca": np.concatenate(cas).astype(np.float32),
           "pairs": np.asarray(pairs, np.int32)}
    for k, (adj, cmap) in enumerate(zip(adjs, dists)):
        a = sp.csr_matrix(adj)
        a.sort_indices()
        out[f"chain{k}_indptr"], out[f"chain{k}_indices"] = a.indptr.astype(np.int64), a.indices.astype(np.int32)
        out[f"chain{k}_dist"] = cmap.astype(np.float64)
    for k, ((a, b), br) in enumerate(zip(pairs, bridges)):
        g1, g2 = gm.generate_graphs(adjs[a], adjs[b])
        u = gm.link_graphs(g1, g2, br)
        f = nx.convert_node_labels_to_integers(u)
        m = sp.csr_matrix(nx.adjacency_matrix(f))
        m.data[:] = 1
        m.sort_indices()
        out[f"pair{k}_indptr"], out[f"pair{k}_indices"] = m.indptr.astype(np.int64), m.indices.astype(np.int32)
        out[f"pair{k}_bridges"] = np.asarray(br, np.int32).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "contact_pairs.npz"), **out)
    print("wrote contact_pairs.npz:", {k: v.shape for k, v in out.items() if not k.startswith("chain")})


if __name__ == "__main__":
    main()
