"""Synthetic inputs of the benchmark shapes: power-law graphs in the
reference's adjacency convention (symmetric, one self-loop per node, row-major sorted;
utils.py:49-55), built directly as CSR on whatever device is asked for -- the dense N x N
matrix the reference's loaders produce never exists at these sizes."""
from __future__ import annotations

import torch


def power_law_csr(n: int, avg_deg: float, seed: int, exponent: float = 0.5, device="cpu"):
    """rowptr int64 [n+1], col int32 [E]; E ~ n*avg_deg before de-duplication.

    One endpoint of every sampled edge is Zipf(exponent) over node rank (inverse CDF: rank =
    n * u^(1/(1-exponent))), the other uniform.  exponent 0.5 at the ogbn-products shape gives a
    max degree of ~2e4, like the real graph's 17.5k."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    m = int(n * max(avg_deg - 1.0, 0.0) / 2.0)
    u = torch.rand(m, generator=g, dtype=torch.float64, device=dev)
    src = (u.pow(1.0 / (1.0 - exponent)) * n).long().clamp_(max=n - 1)
    del u
    dst = torch.randint(0, n, (m,), generator=g, device=dev)
    loops = torch.arange(n, device=dev)
    key = torch.cat([src * n + dst, dst * n + src, loops * n + loops])
    del src, dst
    key = torch.unique(key)  # sorted => row-major (row, col) order, duplicates dropped
    row = key // n
    col = (key - row * n).to(torch.int32)
    del key
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    return rowptr, col


def power_law_shard(n_total: int, lo: int, hi: int, avg_deg: float, seed: int, exponent: float = 0.5, device="cpu"):
    """Destination rows [lo, hi) of a power-law pattern over n_total nodes, generated WITHOUT the rest of the graph
    (papers100M-shaped shards: the whole pattern -- 1.6 G entries -- never exists on one device): rowptr int64
    [hi-lo+1] (local), col int32 [E_local] (GLOBAL source ids), row-major sorted, one self-loop per row.
    Half of the sampled entries take a Zipf(exponent) SOURCE over all n_total nodes and a uniform local row (hub
    columns: rows that every shard gathers), half a Zipf row inside the shard and a uniform global source (hub
    rows: the segment path).  Directed: the layer kernels never assume symmetry."""
    dev = torch.device(device)
    n = hi - lo
    g = torch.Generator(device=dev).manual_seed(seed * 1000003 + lo % 999983)
    m = int(n * max(avg_deg - 1.0, 0.0))
    h = m // 2
    u = torch.rand(h, generator=g, dtype=torch.float64, device=dev)
    src_a = (u.pow(1.0 / (1.0 - exponent)) * n_total).long().clamp_(max=n_total - 1)
    dst_a = torch.randint(0, n, (h,), generator=g, device=dev)
    u = torch.rand(m - h, generator=g, dtype=torch.float64, device=dev)
    dst_b = (u.pow(1.0 / (1.0 - exponent)) * n).long().clamp_(max=n - 1)
    del u
    src_b = torch.randint(0, n_total, (m - h,), generator=g, device=dev)
    loops = torch.arange(n, device=dev)
    key = torch.cat([dst_a * n_total + src_a, dst_b * n_total + src_b, loops * n_total + (loops + lo)])
    del src_a, dst_a, src_b, dst_b
    key = torch.unique(key)
    row = key // n_total
    col = (key - row * n_total).to(torch.int32)
    del key
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    return rowptr, col


def shard_rows_by_nnz(rowptr: torch.Tensor, world: int, row_cost: int = 0):
    """Contiguous destination-row ranges of ~equal COST, cost(row) = stored entries + row_cost (power-law
    graphs are balanced neither by row count nor, once per-row work matters, by entries alone: at the
    products shape the per-row kernels -- projections, ELU, the aggregated-row traffic -- weigh as much as
    ~25 stored entries per row).  Returns world+1 row boundaries (python ints)."""
    n = rowptr.numel() - 1
    cum = rowptr + row_cost * torch.arange(n + 1, device=rowptr.device, dtype=torch.int64)
    total = int(cum[-1].item())
    targets = torch.arange(1, world, device=rowptr.device, dtype=torch.int64) * total // world
    cuts = torch.searchsorted(cum, targets).clamp_(max=n).tolist()
    return [0] + [int(c) for c in cuts] + [n]


def init_layer_params(f_in: int, H: int, D: int, device, seed: int = 72):
    """Per-head parameters of one sparse-class layer (W xavier-normal gain 1.414, a (1,2D)
    xavier-normal; layers.py:111-115) as leaf tensors on `device`, split into (Ws, a_src, a_dst)."""
    import math
    g = torch.Generator(device="cpu").manual_seed(seed)
    Ws, a_src, a_dst = [], [], []
    for _ in range(H):
        w = torch.randn(f_in, D, generator=g) * (1.414 * math.sqrt(2.0 / (f_in + D)))
        a = torch.randn(2 * D, generator=g) * (1.414 * math.sqrt(2.0 / (1 + 2 * D)))
        Ws.append(w.to(device).requires_grad_(True))
        a_src.append(a[:D].clone().to(device).requires_grad_(True))
        a_dst.append(a[D:].clone().to(device).requires_grad_(True))
    return Ws, a_src, a_dst
