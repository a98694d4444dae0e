"""Graph handle: the CSR form of the adjacency pattern the reference recomputes per head with
``adj.nonzero()`` (layers.py:129) or masks with ``adj > 0`` (layers.py:41), built once by the
K0 kernels and cached per adjacency tensor."""
from __future__ import annotations

import os
import weakref
from typing import Optional, Tuple

import torch

from . import _lib, _mem

RULE_NONZERO = 0   # SpGraphAttentionLayer: every entry != 0 (layers.py:129)
RULE_POSITIVE = 1  # GraphAttentionLayer:   entries > 0   (layers.py:41)

DEFAULT_SEG_LEN = int(os.environ.get("GATK_SEG_LEN", "512"))   # rows longer than this are split into segments
ITEM_EDGES = int(os.environ.get("GATK_ITEM_EDGES", "256"))   # target stored entries per scheduler work item


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


_process_device = None  # the one GPU this process drives (per-device kernel attributes are configured once)


def _require_cuda(t: torch.Tensor, what: str):
    """No CPU path, and one process per GPU (the library caches per-device launch configuration): a tensor on a
    second GPU of the same process is refused instead of launching with the first device's settings."""
    global _process_device
    if not t.is_cuda:
        raise RuntimeError(f"pygat_b200: {what} must be a CUDA tensor (the engine has no CPU path); got {t.device}")
    if _process_device is None:
        _process_device = t.device.index
    elif t.device.index != _process_device:
        raise RuntimeError(f"pygat_b200: {what} is on {t.device}, but this process already drives cuda:{_process_device} "
                           "(one process per GPU: launch with torchrun, one rank per device)")


def on_device(t: torch.Tensor):
    """Context manager making t's GPU the current device, so launches and the stream query go to the device the
    operands live on even when the caller's current device is another one."""
    return torch.cuda.device(t.device)


class HubPartition:
    """Rows longer than seg_len, cut into segments (see gatk_attn_fwd in include/gatk.h), and the edge-balanced work
    items of the dynamic scheduler.  Built on the HOST from one copy of the row pointers (`ptr_host`: the caller's
    single device->host read per pattern; a fresh PPI batch used to cost eight small syncs here) and uploaded."""

    def __init__(self, ptr: torch.Tensor, seg_len: int, ptr_host=None):
        import numpy as np
        dev = ptr.device
        ph = (ptr.cpu() if ptr_host is None else ptr_host).numpy()
        deg = ph[1:] - ph[:-1]
        rows = np.nonzero(deg > seg_len)[0]
        self.seg_len = int(seg_len)
        self.n_hub = int(rows.size)
        if self.n_hub:
            nseg = (deg[rows] + seg_len - 1) // seg_len
            seg_ptr = np.zeros(self.n_hub + 1, dtype=np.int64)
            np.cumsum(nseg, out=seg_ptr[1:])
            self.rows = torch.from_numpy(rows.astype(np.int32)).to(dev)
            self.seg_ptr = torch.from_numpy(seg_ptr.astype(np.int32)).to(dev)
            self.n_seg = int(seg_ptr[-1])
        else:
            self.rows = self.seg_ptr = None
            self.n_seg = 0
        # edge-balanced work items for the dynamic scheduler of the gather kernels: consecutive rows are
        # grouped until they hold ~ITEM_EDGES stored entries (hub rows are skipped by those kernels)
        n, e = ph.size - 1, int(ph[-1]) if ph.size > 1 else 0
        # small patterns (Pubmed: 1e5 entries) get finer items, or a 148-SM part would see only a few hundred warps:
        # aim at ~32 items per SM, between 32 and ITEM_EDGES entries each
        self.item_edges = int(min(ITEM_EDGES, max(32, e // (148 * 32)))) if e > 0 else ITEM_EDGES
        if n > 0 and e > 0:
            cuts = np.searchsorted(ph[:-1], np.arange(0, e, self.item_edges, dtype=np.int64), side="left")
            bounds = np.concatenate([cuts, np.array([n], dtype=cuts.dtype)])
            bounds = bounds[np.concatenate([[True], bounds[1:] != bounds[:-1]])]  # unique_consecutive
            if bounds[0] != 0:
                bounds = np.concatenate([np.zeros(1, dtype=bounds.dtype), bounds])
            self.items = torch.from_numpy(bounds.astype(np.int32)).to(dev)
            self.n_items = int(bounds.size) - 1
        else:
            self.items, self.n_items = None, 0
        self.n_empty = int((deg == 0).sum()) if n > 0 else 0
        self.empty_rows_host = np.nonzero(deg == 0)[0] if self.n_empty else None

    def args(self, scratch: Optional[torch.Tensor]):
        return (self.seg_len, _ptr(self.rows), _ptr(self.seg_ptr), self.n_hub, self.n_seg, _ptr(scratch))

    def item_args(self):
        return (_ptr(self.items), self.n_items)


class Graph:
    """CSR pattern (rows = destinations) + lazily built transpose for the backward pass.

    n_dst rows, n_src columns (equal unless the handle is a destination-row shard).
    """

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, n_src: Optional[int] = None,
                 seg_len: Optional[int] = None, rowptr_host: Optional[torch.Tensor] = None):
        _require_cuda(rowptr, "rowptr")
        _require_cuda(col, "col")
        assert rowptr.dtype == torch.int64 and col.dtype == torch.int32
        self.rowptr = rowptr.contiguous()
        self.col = col.contiguous()
        self.n_dst = rowptr.numel() - 1
        self.n_src = int(n_src) if n_src is not None else self.n_dst
        self.nnz = int(col.numel())
        self.device = rowptr.device
        self.seg_len = int(seg_len or DEFAULT_SEG_LEN)
        self.hubs = HubPartition(self.rowptr, self.seg_len, rowptr_host)
        self._counters = {}   # one scheduler counter per (graph, stream): see `counter`
        self._t: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, HubPartition]] = None
        self._iperm: Optional[torch.Tensor] = None

    @property
    def counter(self) -> torch.Tensor:
        """int32 scratch word of the dynamic row scheduler.  One per CUDA stream that uses this graph: a kernel zeroes
        and consumes it in stream order, so two streams working on the same pattern must not share it."""
        key = torch.cuda.current_stream(self.device).cuda_stream
        c = self._counters.get(key)
        if c is None:
            c = self._counters[key] = torch.zeros(1, dtype=torch.int32, device=self.device)
        return c

    # ------------------------------------------------------------------ constructors
    @staticmethod
    def from_dense(adj: torch.Tensor, rule: int = RULE_NONZERO, seg_len: Optional[int] = None) -> "Graph":
        _require_cuda(adj, "adj")
        if adj.dim() != 2 or adj.shape[0] != adj.shape[1]:
            raise RuntimeError(f"adj must be square, got {tuple(adj.shape)}")
        if adj.dtype != torch.float32:
            adj = adj.float()
        n = adj.shape[0]
        rowptr = _mem.empty(n + 1, dtype=torch.int64, device=adj.device)
        ws_bytes = _lib.query("gatk_scan_workspace_bytes", n)
        ws = _mem.empty(max(ws_bytes, 1), dtype=torch.uint8, device=adj.device)
        rs, cs = adj.stride()
        _lib.call("gatk_csr_from_dense_rowptr", adj.data_ptr(), n, rs, cs, rule, rowptr.data_ptr(),
                  ws.data_ptr(), ws_bytes, _stream())
        rowptr_host = rowptr.cpu()  # the ONE host read of the graph build (the reference syncs per head, per call)
        e = int(rowptr_host[-1])
        col = _mem.empty(e, dtype=torch.int32, device=adj.device)
        _lib.call("gatk_csr_from_dense_fill", adj.data_ptr(), n, rs, cs, rule, rowptr.data_ptr(),
                  col.data_ptr(), _stream())
        return Graph(rowptr, col, seg_len=seg_len, rowptr_host=rowptr_host)

    @staticmethod
    def from_coo(edge: torch.Tensor, n: int, seg_len: Optional[int] = None) -> "Graph":
        """edge: (2, E) int64, row-major sorted (what adj.nonzero().t() yields)."""
        _require_cuda(edge, "edge")
        edge = edge.contiguous()
        e = edge.shape[1]
        rowptr = _mem.empty(n + 1, dtype=torch.int64, device=edge.device)
        col = _mem.empty(e, dtype=torch.int32, device=edge.device)
        ws_bytes = _lib.query("gatk_scan_workspace_bytes", n)
        ws = _mem.empty(max(ws_bytes, 1), dtype=torch.uint8, device=edge.device)
        _lib.call("gatk_csr_from_coo", edge[0].data_ptr(), edge[1].data_ptr(), e, n, rowptr.data_ptr(),
                  col.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
        return Graph(rowptr, col, seg_len=seg_len)

    @staticmethod
    def from_csr(rowptr: torch.Tensor, col: torch.Tensor, n_src: Optional[int] = None,
                 seg_len: Optional[int] = None) -> "Graph":
        return Graph(rowptr.to(torch.int64), col.to(torch.int32), n_src=n_src, seg_len=seg_len)

    # ------------------------------------------------------------------ transpose (backward only)
    def transpose(self):
        if self._t is None:
            with on_device(self.rowptr):
                self._build_transpose()
        return self._t

    def _build_transpose(self):
        if self._t is None:
            tptr = _mem.empty(self.n_src + 1, dtype=torch.int64, device=self.device)
            trow = _mem.empty(self.nnz, dtype=torch.int32, device=self.device)
            perm = _mem.empty(self.nnz, dtype=torch.int32, device=self.device)
            ws_bytes = _lib.query("gatk_transpose_workspace_bytes", self.n_dst, self.n_src, self.nnz)
            ws = _mem.empty(max(ws_bytes, 1), dtype=torch.uint8, device=self.device)
            _lib.call("gatk_csr_transpose", self.n_dst, self.n_src, self.nnz, self.rowptr.data_ptr(),
                      _ptr(self.col), tptr.data_ptr(), _ptr(trow), _ptr(perm), ws.data_ptr(), ws_bytes, _stream())
            del ws
            self._t = (tptr, trow, perm, HubPartition(tptr, self.seg_len))
        return self._t

    def empty_rows(self) -> Optional[torch.Tensor]:
        """int64 ids of destination rows without a stored entry, or None (cached; one host read on first use)."""
        if not hasattr(self, "_empty_rows"):  # found on the host copy of the row pointers: no device sync here
            h = self.hubs.empty_rows_host
            self._empty_rows = None if h is None else torch.from_numpy(h).to(self.device)
        return self._empty_rows

    def inverse_perm(self) -> torch.Tensor:
        """int32 [E]: CSR entry -> its position in the transposed pattern (inverse of transpose()[2])."""
        if self._iperm is None:
            perm = self.transpose()[2]
            ip = _mem.empty_like(perm)
            ip[perm.long()] = torch.arange(self.nnz, dtype=torch.int32, device=self.device)
            self._iperm = ip
        return self._iperm

    def edge_index(self) -> torch.Tensor:
        """(2, E) int64 in adj.nonzero().t() order (layers.py:129)."""
        deg = self.rowptr[1:] - self.rowptr[:-1]
        row = torch.repeat_interleave(torch.arange(self.n_dst, device=self.device), deg)
        return torch.stack([row, self.col.long()])


# ---------------------------------------------------------------------- per-adjacency cache
# Keyed on the identity of the tensor OBJECT (weakly referenced, so a recycled data_ptr of a
# freed adjacency can never alias) plus its version counter, strides and the pattern rule.
_cache: "dict[tuple, tuple]" = {}


def graph_of(adj, rule: int = RULE_NONZERO) -> Graph:
    if isinstance(adj, Graph):
        return adj
    if not torch.is_tensor(adj):
        raise RuntimeError(f"adj must be a dense tensor, a sparse tensor or a pygat_b200.Graph, got {type(adj)}")
    key = (id(adj), rule)
    hit = _cache.get(key)
    sig = (adj._version, adj.data_ptr() if not adj.is_sparse else 0, tuple(adj.shape), adj.layout)
    if hit is not None and hit[0]() is adj and hit[1] == sig:
        return hit[2]
    _require_cuda(adj, "adj")
    with on_device(adj):
        if adj.layout == torch.strided:
            g = Graph.from_dense(adj, rule)
        else:
            coo = adj.coalesce() if adj.layout == torch.sparse_coo else adj.to_sparse_coo().coalesce()
            idx, val = coo.indices(), coo.values()
            sel = val != 0 if rule == RULE_NONZERO else val > 0
            g = Graph.from_coo(idx[:, sel], adj.shape[0])
    ref = weakref.ref(adj, lambda _r, k=key: _cache.pop(k, None))
    _cache[key] = (ref, sig, g)
    return g


def clear_cache():
    _cache.clear()
