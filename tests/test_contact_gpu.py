"""K11 / K12 (csrc/contact.cu) through the C ABI: bit-exact against the reference-produced golden vectors and, at
larger sizes, against the NumPy / networkx restatement."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import gcn_string_b200 as g  # noqa: E402
from gcn_string_b200 import contact  # noqa: E402
from oracle import contact_ref  # noqa: E402
from conftest import ROOT  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "contact_pairs.npz"))


def host(t):
    return t.detach().cpu().numpy()


def walk(rng, n):
    step = rng.normal(size=(n, 3))
    step /= np.linalg.norm(step, axis=1, keepdims=True)
    return np.round(np.cumsum(3.8 * step, axis=0), 3).astype(np.float32)


def test_contact_maps_match_the_reference_golden_vectors():
    lengths = GOLD["lengths"]
    cp = contact.chain_offsets(lengths)
    rowptr, col, dist = contact.contact_maps(GOLD["ca"], cp, 10, want_dist=True)
    rowptr, col, dist = host(rowptr), host(col), host(dist)
    assert rowptr.dtype == np.int64 and col.dtype == np.int32
    for k in range(len(lengths)):
        r = rowptr[cp[k]:cp[k + 1] + 1]
        assert np.array_equal(r - r[0], GOLD[f"chain{k}_indptr"])
        c = col[r[0]:r[-1]]
        assert np.array_equal(c, GOLD[f"chain{k}_indices"])
        rows = np.repeat(np.arange(lengths[k]), np.diff(r))
        assert np.array_equal(dist[r[0]:r[-1]].astype(np.float64), GOLD[f"chain{k}_dist"][rows, c])   # bit-exact float32


def test_pair_graphs_match_the_reference_golden_vectors():
    lengths, pairs = GOLD["lengths"], GOLD["pairs"]
    cp = contact.chain_offsets(lengths)
    c_rowptr, c_col, _ = contact.contact_maps(GOLD["ca"], cp, 10)
    br = [GOLD[f"pair{k}_bridges"] for k in range(len(pairs))]
    bptr = np.concatenate([[0], np.cumsum([len(b) for b in br])])
    flat = np.concatenate(br).reshape(-1, 2)
    node_off, rowptr, col = (host(t) for t in contact.link_pairs(c_rowptr, c_col, cp, pairs[:, 0], pairs[:, 1], bptr,
                                                                  flat[:, 0], flat[:, 1]))
    for k, (a, b) in enumerate(pairs):
        assert node_off[k + 1] - node_off[k] == lengths[a] + lengths[b]
        r = rowptr[node_off[k]:node_off[k + 1] + 1]
        assert np.array_equal(r - r[0], GOLD[f"pair{k}_indptr"])
        assert np.array_equal(col[r[0]:r[-1]], GOLD[f"pair{k}_indices"])


def test_contact_maps_larger_batch_and_edge_cases():
    rng = np.random.default_rng(5)
    lengths = [300, 0, 1, 33, 513, 64, 0, 1000]
    cas = [walk(rng, n) if n else np.zeros((0, 3), np.float32) for n in lengths]
    cas[3][5] = np.nan                                                    # NaN coordinates never make a contact
    for thr in (10, 6.5):
        rowptr, col, dist = contact.contact_maps(np.concatenate(cas), contact.chain_offsets(lengths), thr, want_dist=True)
        rowptr, col, dist = host(rowptr), host(col), host(dist)
        off = np.concatenate([[0], np.cumsum(lengths)])
        for k, ca in enumerate(cas):
            if not lengths[k]:
                continue
            ip, ix, d = contact_ref.contact_csr(ca, thr)
            r = rowptr[off[k]:off[k + 1] + 1]
            assert np.array_equal(r - r[0], ip) and np.array_equal(col[r[0]:r[-1]], ix)
            assert np.array_equal(dist[r[0]:r[-1]], d)
    r0, c0, _ = contact.contact_maps(np.zeros((0, 3), np.float32), np.zeros(1, np.int32))
    assert host(r0).tolist() == [0] and c0.numel() == 0
    with pytest.raises(ValueError):
        contact.contact_maps(np.zeros((4, 3), np.float64), [0, 4])        # float64 would change the decisions
    with pytest.raises(ValueError):
        contact.contact_maps(np.zeros((4, 3), np.float32), [0, 3])


def test_pair_dataset_trains_through_the_loader():
    """CA coordinates -> packed dataset -> DisjointLoader -> one train step; the structure equals the networkx route."""
    rng = np.random.default_rng(9)
    lengths = [40, 55, 31, 62]
    cas = [walk(rng, n) for n in lengths]
    pairs = [(0, 1), (2, 3), (1, 2), (3, 0), (0, 2), (1, 3)]
    bridges = [[(int(rng.integers(lengths[a])), int(rng.integers(lengths[b]))) for _ in range(20)] for a, b in pairs]
    feats = rng.random((sum(lengths), 16)).astype(np.float32)
    ds = contact.pair_dataset(np.concatenate(cas), lengths, pairs, bridges, feats, [0, 1, 0, 1, 1, 0])
    adjs = [contact_ref.proximity_matrix(ca, 10)[0] for ca in cas]
    off = np.concatenate([[0], np.cumsum(lengths)])
    for k, (a, b) in enumerate(pairs):
        x, adj, y = ds.graph(k)
        m = contact_ref.pair_adjacency(adjs[a], adjs[b], bridges[k])
        assert np.array_equal(adj.indptr, m.indptr) and np.array_equal(adj.indices, m.indices)
        assert np.array_equal(x, np.concatenate([feats[off[a]:off[a + 1]], feats[off[b]:off[b + 1]]]))
    (x, a, i), y = next(g.DisjointLoader(ds, batch_size=6, epochs=1, shuffle=False))
    model = g.GeneralGNN(2, activation="softmax", hidden=32, message_passing=2, seed=0)
    loss_acc, probs = model.train_step_grads([x, a, i], y)
    assert np.isfinite(host(loss_acc)).all() and probs.shape == (6, 2)
    with pytest.raises(ValueError, match="bridge"):
        contact.pair_dataset(np.concatenate(cas), lengths, [(0, 1)], [[(40, 0)]], feats, [0])
    with pytest.raises(ValueError, match="chain id"):
        contact.pair_dataset(np.concatenate(cas), lengths, [(0, 4)], [[]], feats, [0])
