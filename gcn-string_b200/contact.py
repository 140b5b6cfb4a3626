"""Contact maps and inter-protein pair graphs on the device (kernels K11 / K12, csrc/contact.cu).

Replaces the graph generator's hot loops - ``GraphMaker.generate_proximity_matrix`` (an O(n^2) Python loop per chain,
src/utilities/gcn_utills.py:161-238), ``generate_graphs`` (:240-270) and ``link_graphs`` (:319-377) - for a whole batch
of chains / protein pairs per launch, and emits the packed layout ``DisjointLoader`` consumes, so structures go from
CA coordinates to training batches without networkx, gpickle files or a scipy collate.  Integer outputs are
bit-identical to the reference's NumPy / networkx results (tests/golden/contact_pairs.npz, produced by the reference's
own code).  No CPU fallback: every function needs the CUDA library.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check, ptr, stream_ptr
from .synthetic import PackedGraphs


def _i32(a, name):
    torch = _lib.require_cuda()
    t = _lib.as_tensor(a) if not isinstance(a, (list, tuple, np.ndarray)) else torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype not in (torch.int32, torch.int64):
        raise ValueError(f"{name} must be an integer array")
    return t.to(device="cuda", dtype=torch.int32).contiguous()


def chain_offsets(lengths: Sequence[int]) -> np.ndarray:
    """int32 residue offsets [n_chains + 1] of concatenated chains."""
    out = np.zeros(len(lengths) + 1, np.int64)
    np.cumsum(np.asarray(lengths, np.int64), out=out[1:])
    if out[-1] >= 2 ** 31:
        raise ValueError("more than 2^31 residues in one call")
    return out.astype(np.int32)


def contact_maps(ca, chain_ptr, angstroms: float = 10, want_dist: bool = False):
    """Proximity graphs of all chains: ``ca`` float32 [n_residues, 3] (CA coordinates, chains concatenated),
    ``chain_ptr`` [n_chains + 1].  Returns device tensors (rowptr int64 [n_residues + 1], col int32 [nnz] chain-local
    ascending, dist float32 [nnz] or None): entry (i, j) present iff float32 ||ca_i - ca_j|| < angstroms."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    ca = _lib.as_tensor(ca) if not isinstance(ca, np.ndarray) else torch.from_numpy(np.ascontiguousarray(ca))
    if ca.dim() != 2 or ca.shape[1] != 3:
        raise ValueError("ca must have shape [n_residues, 3]")
    if ca.dtype != torch.float32:
        raise ValueError("ca must be float32 (Bio.PDB coordinates are; casting float64 here would change d < angstroms)")
    ca = ca.to(device="cuda").contiguous()
    chain_ptr = _i32(chain_ptr, "chain_ptr")
    n_res, n_chains = ca.shape[0], chain_ptr.shape[0] - 1
    if n_chains < 0 or (n_chains == 0 and n_res) or (n_chains and int(chain_ptr[-1].item()) != n_res) or \
            (n_chains and (int(chain_ptr[0].item()) != 0 or bool((chain_ptr[1:] < chain_ptr[:-1]).any().item()))):
        raise ValueError("chain_ptr must rise from 0 to n_residues")
    rowptr = torch.empty(n_res + 1, dtype=torch.int64, device="cuda")
    ws = torch.empty(max(int(lib.gcs_contact_workspace_bytes(n_res)), 256), dtype=torch.uint8, device="cuda")
    check(lib.gcs_contact_map_rowptr(ptr(ca), ptr(chain_ptr), n_chains, n_res, float(angstroms), ptr(rowptr), ptr(ws),
                                     ws.numel(), stream_ptr()), "gcs_contact_map_rowptr")
    nnz = int(rowptr[-1].item())                               # the one host read: sizes the column array
    col = torch.empty(max(nnz, 1), dtype=torch.int32, device="cuda")[:nnz]
    dist = torch.empty(max(nnz, 1), dtype=torch.float32, device="cuda")[:nnz] if want_dist else None
    check(lib.gcs_contact_map_fill(ptr(ca), ptr(chain_ptr), n_chains, n_res, float(angstroms), ptr(rowptr), ptr(col),
                                   ptr(dist), stream_ptr()), "gcs_contact_map_fill")
    return rowptr, col, dist


def link_pairs(chain_rowptr, chain_col, chain_ptr, pair_a, pair_b, bridge_ptr, bridge_a, bridge_b):
    """Pair graphs = union of two chains' contact graphs + one symmetric edge per DCA bridge
    (``bridge_a[q]`` in chain ``pair_a[p]``, ``bridge_b[q]`` in chain ``pair_b[p]``, q in
    [bridge_ptr[p], bridge_ptr[p+1])).  Returns device tensors (node_off int64 [P+1], rowptr int64, col int32)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    chain_ptr, pair_a, pair_b = _i32(chain_ptr, "chain_ptr"), _i32(pair_a, "pair_a"), _i32(pair_b, "pair_b")
    bridge_ptr, bridge_a, bridge_b = _i32(bridge_ptr, "bridge_ptr"), _i32(bridge_a, "bridge_a"), _i32(bridge_b, "bridge_b")
    n_pairs, n_chains = pair_a.shape[0], chain_ptr.shape[0] - 1
    if pair_b.shape[0] != n_pairs or bridge_ptr.shape[0] != n_pairs + 1 or bridge_a.shape != bridge_b.shape:
        raise ValueError("pair_a / pair_b / bridge_ptr / bridge_a / bridge_b have inconsistent lengths")
    if n_pairs and (int(bridge_ptr[0].item()) != 0 or int(bridge_ptr[-1].item()) != bridge_a.shape[0]
                    or bool((bridge_ptr[1:] < bridge_ptr[:-1]).any().item())):
        raise ValueError("bridge_ptr must rise from 0 to the number of bridges")
    if chain_rowptr.dtype != torch.int64 or chain_col.dtype != torch.int32:
        raise ValueError("chain_rowptr must be int64 and chain_col int32 (the outputs of contact_maps)")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    node_off = torch.empty(n_pairs + 1, dtype=torch.int64, device="cuda")
    ws = torch.empty(max(int(lib.gcs_contact_workspace_bytes(n_pairs)), 256), dtype=torch.uint8, device="cuda")
    check(lib.gcs_link_pairs_offsets(ptr(chain_ptr), n_chains, ptr(pair_a), ptr(pair_b), n_pairs, ptr(node_off), ptr(flag),
                                     ptr(ws), ws.numel(), stream_ptr()), "gcs_link_pairs_offsets")
    if int(flag.item()) & 1:
        raise ValueError("pair_a / pair_b hold a chain id outside [0, n_chains)")
    n_rows = int(node_off[-1].item())
    rowptr = torch.empty(n_rows + 1, dtype=torch.int64, device="cuda")
    ws = torch.empty(max(int(lib.gcs_contact_workspace_bytes(n_rows)), 256), dtype=torch.uint8, device="cuda")
    args = (ptr(chain_rowptr), ptr(chain_col), ptr(chain_ptr), ptr(pair_a), ptr(pair_b), n_pairs, ptr(bridge_ptr),
            ptr(bridge_a), ptr(bridge_b), ptr(node_off), n_rows, ptr(rowptr))
    check(lib.gcs_link_pairs(*args, None, ptr(flag), ptr(ws), ws.numel(), stream_ptr()), "gcs_link_pairs")
    nnz = int(rowptr[-1].item())
    if int(flag.item()) & 2:
        raise ValueError("a DCA bridge points outside its chain (networkx would add a new node; not supported)")
    col = torch.empty(max(nnz, 1), dtype=torch.int32, device="cuda")[:nnz]
    check(lib.gcs_link_pairs(*args, ptr(col), ptr(flag), ptr(ws), ws.numel(), stream_ptr()), "gcs_link_pairs")
    return node_off, rowptr, col


def pair_dataset(ca, chain_lengths: Sequence[int], pairs, bridges: Sequence[Sequence[Tuple[int, int]]], chain_x, labels,
                 angstroms: float = 10, n_classes: int = 2) -> PackedGraphs:
    """CA coordinates + per-residue features + DCA bridges -> the packed dataset ``DisjointLoader`` trains on.
    ``pairs`` [P, 2] chain ids, ``bridges[p]`` list of (pos_in_a, pos_in_b), ``chain_x`` float [n_residues, F]
    (NetSurfP features, gcn_utills.py:272-317), ``labels`` [P] class ids (one-hot encoded like gcn.py:259-262)."""
    torch = _lib.require_cuda()
    chain_ptr = chain_offsets(chain_lengths)
    pairs = np.asarray(pairs, np.int64).reshape(-1, 2)
    if len(bridges) != pairs.shape[0] or len(labels) != pairs.shape[0]:
        raise ValueError("one bridge list and one label per pair")
    bptr = np.zeros(pairs.shape[0] + 1, np.int64)
    np.cumsum([len(b) for b in bridges], out=bptr[1:])
    flat = np.asarray([q for b in bridges for q in b], np.int64).reshape(-1, 2)
    c_rowptr, c_col, _ = contact_maps(ca, chain_ptr, angstroms)
    node_off, rowptr, col = link_pairs(c_rowptr, c_col, chain_ptr, pairs[:, 0], pairs[:, 1], bptr, flat[:, 0], flat[:, 1])
    # node features: residues of chain a then of chain b (nx.union order), gathered on the device
    cp = torch.from_numpy(chain_ptr.astype(np.int64)).cuda()
    pa, pb = torch.from_numpy(pairs[:, 0]).cuda(), torch.from_numpy(pairs[:, 1]).cuda()
    na = cp[pa + 1] - cp[pa]
    rows = torch.arange(int(node_off[-1].item()), device="cuda")
    p = torch.searchsorted(node_off, rows, right=True) - 1
    r = rows - node_off[p]
    src = torch.where(r < na[p], cp[pa[p]] + r, cp[pb[p]] + (r - na[p]))
    x = torch.as_tensor(np.ascontiguousarray(chain_x, dtype=np.float32)).cuda()[src]
    y = np.zeros((pairs.shape[0], n_classes), np.float32)
    y[np.arange(pairs.shape[0]), np.asarray(labels, np.int64)] = 1.0
    return PackedGraphs(node_off.cpu().numpy(), rowptr.cpu().numpy(), col.cpu().numpy(), x.cpu().numpy(), y)
