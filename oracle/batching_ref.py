"""O3 — the literal host collate of Spektral's ``DisjointLoader`` (TEST INFRASTRUCTURE).

Reference call sites: /root/reference/src/scripts/gcn.py:316-317 (loader construction),
:350,:367 (iteration).  Upstream functions restated [spektral/data/loaders.py,
spektral/data/utils.py, spektral/utils/sparse.py; SURVEY.md §8 a1]:

  to_disjoint:              x = np.vstack(x_list); a = sp.block_diag(a_list);
                            i = np.repeat(np.arange(B), n_nodes)
  sp_matrix_to_sp_tensor:   row, col, values = sp.find(a)   (sums duplicates, drops zeros)
                            SparseTensor(indices=[row, col].T, values, dense_shape)
                            tf.sparse.reorder -> canonical row-major order
  collate_labels_disjoint:  y = np.vstack(y_list)   (graph-level labels)
  batch_generator:          per epoch shuffle (np.random), consecutive slices of
                            batch_size, last batch short; steps = ceil(len / bs)

Because these are scipy/numpy calls, this file runs the real thing: integer outputs of the
CUDA batching kernel are compared bit-for-bit against it.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def to_disjoint(x_list, a_list):
    """Upstream ``to_disjoint`` for (x, a) without edge features."""
    x_out = np.vstack(x_list)
    a_out = sp.block_diag(a_list)
    n_nodes = np.array([x.shape[0] for x in x_list])
    i_out = np.repeat(np.arange(len(n_nodes)), n_nodes)
    return x_out, a_out, i_out


def sp_matrix_to_sp_tensor(a):
    """Upstream ``sp_matrix_to_sp_tensor`` + ``tf.sparse.reorder`` -> (indices[nnz,2] int64
    row-major, values[nnz], dense_shape[2] int64)."""
    row, col, values = sp.find(a)
    indices = np.array([row, col]).T.astype(np.int64)
    order = np.lexsort((indices[:, 1], indices[:, 0]))       # tf.sparse.reorder
    return indices[order], values[order], np.array(a.shape, dtype=np.int64)


def collate(graphs):
    """graphs: list of (x[n,F], a scipy [n,n], y[C]).  Returns the tuple the reference's
    train_step receives: ((x, (indices, values, dense_shape), i), y)."""
    x, a, i = to_disjoint([g[0] for g in graphs], [g[1] for g in graphs])
    indices, values, shape = sp_matrix_to_sp_tensor(a)
    y = np.vstack([np.asarray(g[2]) for g in graphs])
    return (x, (indices, values, shape), i.astype(np.int64)), y


def derived_csr(indices, n_rows):
    """rowptr / colidx / degree the CUDA path derives from the row-major COO."""
    counts = np.bincount(indices[:, 0], minlength=n_rows)
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, indices[:, 1].copy(), counts


def graph_ptr(i, n_graphs):
    """Segment offsets of the sorted batch index ``i``."""
    gp = np.zeros(n_graphs + 1, dtype=np.int64)
    np.cumsum(np.bincount(i, minlength=n_graphs), out=gp[1:])
    return gp


def batch_slices(n, batch_size):
    """(start, stop) of every batch of one epoch (upstream ``batch_generator``)."""
    steps = int(np.ceil(n / batch_size))
    return [(b * batch_size, min((b + 1) * batch_size, n)) for b in range(steps)]
