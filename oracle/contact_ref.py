"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's contact-map and
pair-graph construction, the checker for csrc/contact.cu.

Unlike the model math, this part of the path is the reference's OWN NumPy / networkx code, so parity is PINNED:
tests/golden/make_contact_golden.py imports /root/reference/src/utilities/gcn_utills.py (third-party imports that
are absent here - Bio, seaborn, matplotlib - stubbed, they are not touched by these functions) and runs the
unmodified ``GraphMaker.generate_proximity_matrix`` / ``generate_graphs`` / ``link_graphs`` on seeded chains; the
outputs are committed as tests/golden/contact_pairs.npz and this restatement is checked against them.

  residue_distance      gcn_utills.py:161-178  diff = a.coord - b.coord (float32[3]); np.sqrt(np.sum(diff * diff))
  distance_matrix       gcn_utills.py:180-201  d_mat[row, col] = residue_distance(seq[row], seq[col]), float64 store
  proximity_matrix      gcn_utills.py:203-238  adjacency[contact_map < angstroms] = 1
  pair_adjacency        gcn_utills.py:240-270 (nx.from_numpy_matrix), :319-377 (nx.union + add_edge per bridge),
                        gcn.py:184-197 (convert_node_labels_to_integers), :104-117 (nx.adjacency_matrix, 0/1 pattern)
"""
import numpy as np
import scipy.sparse as sp


def residue_distance(ca_i, ca_j):
    diff = np.asarray(ca_i, np.float32) - np.asarray(ca_j, np.float32)
    return np.sqrt(np.sum(diff * diff))


def distance_matrix_loop(ca):
    """The literal double loop (small chains only)."""
    n = len(ca)
    d = np.zeros((n, n), np.float64)
    for r in range(n):
        for c in range(n):
            d[r, c] = residue_distance(ca[r], ca[c])
    return d


def distance_matrix(ca):
    """Same float32 arithmetic, vectorised: np.sum over 3 float32 values is ((x + y) + z)."""
    ca = np.asarray(ca, np.float32)
    diff = ca[:, None, :] - ca[None, :, :]
    sq = diff * diff
    return np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2]).astype(np.float64)


def proximity_matrix(ca, angstroms=10):
    contact_map = distance_matrix(ca)
    adjacency = np.zeros(contact_map.shape)
    adjacency[contact_map < angstroms] = 1
    return adjacency, contact_map


def contact_csr(ca, angstroms=10):
    """(indptr int64, indices int32 ascending, dist float32) of one chain's proximity matrix."""
    adjacency, contact_map = proximity_matrix(ca, angstroms)
    a = sp.csr_matrix(adjacency)
    a.sort_indices()
    rows = np.repeat(np.arange(a.shape[0]), np.diff(a.indptr))
    return a.indptr.astype(np.int64), a.indices.astype(np.int32), contact_map[rows, a.indices].astype(np.float32)


def pair_adjacency(adj_a, adj_b, bridges):
    """0/1 CSR of the linked pair graph via networkx, as the reference builds it."""
    import networkx as nx
    g1, g2 = nx.from_numpy_array(np.asarray(adj_a)), nx.from_numpy_array(np.asarray(adj_b))
    u = nx.union(g1, g2, rename=("a-", "b-"))
    for b1, b2 in bridges:
        u.add_edge("a-" + str(b1), "b-" + str(b2))
    f = nx.convert_node_labels_to_integers(u)
    a = sp.csr_matrix(nx.adjacency_matrix(f))
    a.data[:] = 1
    a.sort_indices()
    return a
