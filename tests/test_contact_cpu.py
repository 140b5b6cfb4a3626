"""Contact-map / pair-graph oracle against the golden vectors the REFERENCE's own code produced
(tests/golden/make_contact_golden.py ran GraphMaker.generate_proximity_matrix / generate_graphs / link_graphs from
/root/reference unmodified).  CPU only; the CUDA kernels are checked in test_contact_gpu.py."""
import os

import numpy as np
import pytest

from oracle import contact_ref
from conftest import ROOT

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "contact_pairs.npz"))


def chains():
    off = np.concatenate([[0], np.cumsum(GOLD["lengths"])])
    return [GOLD["ca"][off[k]:off[k + 1]] for k in range(len(GOLD["lengths"]))]


def test_distance_restatement_is_bit_identical_to_the_reference():
    for k, ca in enumerate(chains()):
        d = contact_ref.distance_matrix(ca)
        assert d.dtype == np.float64 and np.array_equal(d, GOLD[f"chain{k}_dist"])       # float32 values, stored as float64
        if len(ca) <= 40:
            assert np.array_equal(contact_ref.distance_matrix_loop(ca), d)                 # the literal double loop
        indptr, indices, dist = contact_ref.contact_csr(ca, 10)
        assert np.array_equal(indptr, GOLD[f"chain{k}_indptr"]) and np.array_equal(indices, GOLD[f"chain{k}_indices"])
        assert np.all(dist < 10) and np.all(np.diff(indptr) >= 1)                          # self-loops: d = 0 < 10


def test_a_distance_of_exactly_the_threshold_is_not_a_contact():
    ca = np.array([[0, 0, 0], [10, 0, 0], [6, 8, 0], [9.999999, 0, 0]], np.float32)
    adj, d = contact_ref.proximity_matrix(ca, 10)
    assert d[0, 1] == 10.0 and adj[0, 1] == 0 and d[0, 2] == 10.0 and adj[0, 2] == 0 and adj[0, 3] == 1


def test_pair_graph_restatement_matches_the_reference():
    cas = chains()
    adjs = [contact_ref.proximity_matrix(ca, 10)[0] for ca in cas]
    for k, (a, b) in enumerate(GOLD["pairs"]):
        m = contact_ref.pair_adjacency(adjs[a], adjs[b], [tuple(q) for q in GOLD[f"pair{k}_bridges"]])
        assert np.array_equal(m.indptr, GOLD[f"pair{k}_indptr"]) and np.array_equal(m.indices, GOLD[f"pair{k}_indices"])
        assert (m != m.T).nnz == 0 and np.all(m.diagonal() == 1)
        na = len(cas[a])
        for b1, b2 in GOLD[f"pair{k}_bridges"]:
            assert m[b1, na + b2] == 1 and m[na + b2, b1] == 1


def test_vectorised_restatement_equals_the_literal_loop_on_random_chains():
    """Property check with hypothesis: for random small chains (including coincident residues and coordinates that make
    distances straddle the threshold) the vectorised float32 restatement equals the reference's literal double loop
    bit for bit, the adjacency is symmetric with a full diagonal, and linking adds exactly the bridge edges."""
    from hypothesis import given, settings, strategies as st

    coords = st.lists(st.tuples(*[st.floats(-30, 30, width=32) for _ in range(3)]), min_size=1, max_size=14)

    @settings(max_examples=40, deadline=None)
    @given(coords, coords, st.integers(0, 2 ** 31 - 1))
    def check(ca, cb, seed):
        ca, cb = np.asarray(ca, np.float32), np.asarray(cb, np.float32)
        d = contact_ref.distance_matrix(ca)
        assert np.array_equal(d, contact_ref.distance_matrix_loop(ca))
        adj_a, _ = contact_ref.proximity_matrix(ca, 10)
        adj_b, _ = contact_ref.proximity_matrix(cb, 10)
        assert np.array_equal(adj_a, adj_a.T) and np.all(np.diag(adj_a) == 1)
        rng = np.random.default_rng(seed)
        bridges = [(int(rng.integers(len(ca))), int(rng.integers(len(cb)))) for _ in range(int(rng.integers(0, 5)))]
        m = contact_ref.pair_adjacency(adj_a, adj_b, bridges).toarray()
        na = len(ca)
        expect = np.zeros_like(m)
        expect[:na, :na], expect[na:, na:] = adj_a, adj_b
        for b1, b2 in bridges:
            expect[b1, na + b2] = expect[na + b2, b1] = 1
        assert np.array_equal(m, expect)

    check()
