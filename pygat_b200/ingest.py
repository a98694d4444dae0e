"""Graph ingest without a dense adjacency.

The reference's loaders hand the layers a dense N x N float matrix (utils.py:55, load_data_ppi.py:153,
:86) of which only the sparsity pattern is ever read (layers.py:41, :129).  That matrix is the reason the
reference cannot reach the ogbn-products shape (24 TB).  The functions here build the SAME pattern --
bit-exact against `adj.nonzero()` of what the loaders would have produced -- straight from edge lists,
scipy sparse matrices and per-graph CSR pieces:

    pattern_from_edges      utils.py:30, 49-52: coo_matrix of the edge list, symmetrised, + identity
    pattern_from_scipy      utils.py:45-52 for the *_dgl datasets (adj_sparse.npz)
    batch_patterns          load_data_ppi.py:84-86: torch.block_diag of the batch's adjacencies
    graph_from_* / batch_graphs   the same, as pygat_b200.Graph handles on a CUDA device

The pattern arithmetic is a handful of torch ops on int64 keys (device agnostic, so it is tested on CPU
against scipy); only the Graph handle itself needs CUDA.  Every model entry point accepts a Graph wherever
the reference takes `adj`.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch


def _csr_from_sorted_keys(key: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    row = torch.div(key, n, rounding_mode="floor")
    col = (key - row * n).to(torch.int32)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=key.device)
    if key.numel():
        rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    return rowptr, col


def pattern_from_edges(row: torch.Tensor, col: torch.Tensor, n: int, symmetric: bool = True,
                       self_loops: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """CSR (rowptr int64 [n+1], col int32 [E]) of the pattern the reference's citation loader ends up with.

    utils.py:30 builds `coo_matrix(ones, (edges[:,0], edges[:,1]))` (duplicates add up: still one stored
    entry), utils.py:49 symmetrises it (`adj + adj.T.multiply(adj.T > adj) - adj.multiply(adj.T > adj)` is the
    element-wise maximum of adj and its transpose: entry (i,j) is stored iff (i,j) or (j,i) is) and utils.py:52
    adds the identity before the normalisation (which rescales but never zeroes an entry).  Entries come
    out in (row, col) order, i.e. the order of `adj.nonzero()` (layers.py:129)."""
    row = row.reshape(-1).to(torch.int64)
    col = col.reshape(-1).to(torch.int64)
    if row.numel() != col.numel():
        raise ValueError("row and col must have the same number of entries")
    if row.numel() and (int(row.min()) < 0 or int(col.min()) < 0 or int(row.max()) >= n or int(col.max()) >= n):
        raise ValueError(f"edge endpoint outside [0, {n})")
    keys = [row * n + col]
    if symmetric:
        keys.append(col * n + row)
    if self_loops:
        d = torch.arange(n, dtype=torch.int64, device=row.device)
        keys.append(d * n + d)
    key = torch.unique(torch.cat(keys))  # sorted: row-major (row, col) order, duplicates dropped
    return _csr_from_sorted_keys(key, n)


def pattern_from_scipy(mat, symmetric: bool = True, self_loops: bool = True, device="cpu"):
    """The same pattern from a scipy sparse matrix (utils.py:45: `load_npz('.../adj_sparse.npz')`).  Stored
    zeros are not edges (scipy's `>` / `multiply` and the final `todense().nonzero()` drop them too)."""
    coo = mat.tocoo()
    if coo.shape[0] != coo.shape[1]:
        raise ValueError(f"adjacency must be square, got {coo.shape}")
    keep = coo.data != 0
    row = torch.from_numpy(coo.row[keep].astype("int64")).to(device)
    col = torch.from_numpy(coo.col[keep].astype("int64")).to(device)
    return pattern_from_edges(row, col, coo.shape[0], symmetric, self_loops)


def batch_patterns(parts: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Block-diagonal union of per-graph CSR patterns: what `torch.block_diag(*adjacency_matrix_list)`
    (load_data_ppi.py:86) yields for the batch, without the dense blocks.  Node ids of graph k are shifted by
    the node count of graphs 0..k-1, as the reference's collate function does implicitly."""
    if not parts:
        raise ValueError("empty batch")
    dev = parts[0][0].device
    rowptrs, cols = [torch.zeros(1, dtype=torch.int64, device=dev)], []
    node_off, nnz_off = 0, 0
    for rowptr, col in parts:
        rowptr = rowptr.to(torch.int64)
        n_k, e_k = rowptr.numel() - 1, int(col.numel())
        if int(rowptr[-1]) != e_k:
            raise ValueError("rowptr does not match col")
        rowptrs.append(rowptr[1:] + nnz_off)
        cols.append(col.to(torch.int32) + node_off)
        node_off += n_k
        nnz_off += e_k
    return torch.cat(rowptrs), (torch.cat(cols) if cols else torch.zeros(0, dtype=torch.int32, device=dev))


# ---------------------------------------------------------------------- Graph handles (CUDA)
def graph_from_edges(row: torch.Tensor, col: torch.Tensor, n: int, symmetric: bool = True, self_loops: bool = True,
                     seg_len: Optional[int] = None):
    from .graph import Graph
    rowptr, c = pattern_from_edges(row, col, n, symmetric, self_loops)
    return Graph.from_csr(rowptr, c, seg_len=seg_len)


def graph_from_scipy(mat, device, symmetric: bool = True, self_loops: bool = True, seg_len: Optional[int] = None):
    from .graph import Graph
    rowptr, c = pattern_from_scipy(mat, symmetric, self_loops, device)
    return Graph.from_csr(rowptr, c, seg_len=seg_len)


def load_dgl_adjacency(folder: str, device, seg_len: Optional[int] = None):
    """`utils.load_data` for the pubmed / citeseer branches (utils.py:35-52), adjacency only: reads
    `<folder>/adj_sparse.npz` and returns the Graph of the symmetrised pattern plus identity."""
    import os

    from scipy.sparse import load_npz
    return graph_from_scipy(load_npz(os.path.join(folder, "adj_sparse.npz")), device, seg_len=seg_len)


def batch_graphs(graphs, seg_len: Optional[int] = None):
    """One Graph for a batch of independent graphs (PPI: load_data_ppi.py:84-86)."""
    from .graph import Graph
    rowptr, col = batch_patterns([(g.rowptr, g.col) for g in graphs])
    return Graph.from_csr(rowptr, col, seg_len=seg_len)
