"""CPU oracle for the GAT-layer hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file restates, as plain functions over CPU tensors, the algorithm of the
reference's layer operators.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product path (``pygat_b200``) never does and fails loudly without its CUDA
library.

Parity status: PINNED.  The reference has no golden vectors of its own
(SURVEY.md section 4), so the pin is the reference itself, imported unmodified
from ``/root/reference`` by ``tests/golden/make_golden.py`` (with a six-line
``torch_scatter.scatter_max`` shim, the reference's one un-vendored dependency,
version unpinned) and run on seeded inputs; the resulting tensors are committed
under ``tests/golden/`` and ``tests/test_oracle_golden.py`` checks every
function below against them.

Reference lines each function follows (all under /root/reference):

* ``edge_list``                  layers.py:129  (``adj.nonzero().t()``) and
                                 layers.py:41   (``adj > 0`` for the dense class)
* ``segment_max``                layers.py:145  (``torch_scatter.scatter_max``)
* ``coo_matmul`` / ``CooMatmul`` layers.py:70-90 (``SpecialSpmmFunction``)
* ``sparse_head``                layers.py:125-173 (``SpGraphAttentionLayer.forward``)
* ``dense_head``                 layers.py:32-64  (``GraphAttentionLayer.forward``)
* ``sparse_head_v2``             layers.py:255-313 (``SpGraphAttentionLayerV2.forward``)
* ``dense_head_v2``              layers.py:203-229 (``GraphAttentionLayerV2.forward``)
* ``gat_forward``                models.py:29-35  (``GAT.forward``)
* ``init_head`` / ``init_gat``   layers.py:21-28, 111-119 and models.py:15-27

Arithmetic is whatever dtype the inputs carry (fp32 to mirror the reference,
fp64 for a tighter yardstick).  Dropout masks are explicit inputs (``keep``
tensors of 0/1) so that a test can hand the same mask to the CUDA path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

NEG_FILL = -9e15  # layers.py:40


# --------------------------------------------------------------------------- graph
def edge_list(adj: torch.Tensor, rule: str = "nonzero") -> torch.Tensor:
    """(2, E) int64 edges in row-major order.

    rule="nonzero": layers.py:129, every entry != 0 (negative entries count).
    rule="positive": the pattern the dense class masks with, layers.py:41.
    """
    if rule == "nonzero":
        return adj.nonzero().t()
    if rule == "positive":
        return (adj > 0).nonzero().t()
    raise ValueError(rule)


def csr_from_edges(edge: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """rowptr (int64, n+1) / col (int32, E) of a row-major sorted edge list."""
    counts = torch.bincount(edge[0], minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr, edge[1].to(torch.int32)


def csr_transpose(rowptr: torch.Tensor, col: torch.Tensor, n: int):
    """CSR of the transposed pattern plus the permutation into the original edge ids.

    Stable in the original edge order, so within one source column the
    destination rows appear ascending.
    """
    e = col.numel()
    row = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), rowptr[1:] - rowptr[:-1])
    perm = torch.argsort(col.to(torch.int64), stable=True)
    counts = torch.bincount(col.to(torch.int64), minlength=n)
    tptr = torch.zeros(n + 1, dtype=torch.int64)
    tptr[1:] = torch.cumsum(counts, 0)
    assert perm.numel() == e
    return tptr, row[perm].to(torch.int32), perm


def segment_max(src: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """torch_scatter.scatter_max(src, index)[0]: 1-D segment max, length index.max()+1,
    empty segments give 0 (layers.py:145).  torch_scatter is not vendored by the
    reference and its version is unpinned; this restates its documented behaviour."""
    size = int(index.max().item()) + 1 if index.numel() else 0
    out = torch.zeros(size, dtype=src.dtype)
    return out.scatter_reduce(0, index, src, reduce="amax", include_self=False)


class CooMatmul(torch.autograd.Function):
    """layers.py:70-90.  ``faithful=True`` keeps the reference's dense N x N backward
    (layers.py:85-87); ``faithful=False`` gathers the same E entries row by row
    without the N x N temporary (identical values, used where N^2 does not fit)."""

    @staticmethod
    def forward(ctx, indices, values, shape, b, faithful):
        assert not indices.requires_grad  # layers.py:74
        sp = torch.sparse_coo_tensor(indices, values, shape, check_invariants=False)
        ctx.save_for_backward(indices, values, b)
        ctx.shape = shape
        ctx.faithful = faithful
        return torch.matmul(sp, b)

    @staticmethod
    def backward(ctx, grad_out):
        indices, values, b = ctx.saved_tensors
        n = ctx.shape[0]
        g_values = g_b = None
        if ctx.needs_input_grad[1]:
            if ctx.faithful:
                dense = grad_out.matmul(b.t())
                g_values = dense.view(-1)[indices[0] * n + indices[1]]
            else:
                g_values = (grad_out[indices[0]] * b[indices[1]]).sum(1)
        if ctx.needs_input_grad[3]:
            sp_t = torch.sparse_coo_tensor(indices.flip(0), values, (ctx.shape[1], ctx.shape[0]),
                                           check_invariants=False)
            g_b = torch.matmul(sp_t, grad_out)
        return None, g_values, None, g_b, None


def coo_matmul(indices, values, shape, b, faithful=True):
    return CooMatmul.apply(indices, values, shape, b, faithful)


# --------------------------------------------------------------------------- heads
def _drop(t: torch.Tensor, keep: Optional[torch.Tensor], p: float) -> torch.Tensor:
    """F.dropout with an explicit keep mask: kept entries scaled by 1/(1-p)."""
    if keep is None or p == 0.0:
        return t
    return t * keep.to(t.dtype) / (1.0 - p)


def _leaky(t: torch.Tensor, alpha: float) -> torch.Tensor:
    return torch.where(t > 0, t, t * alpha)


def sparse_head(x, W, a, edge, alpha, concat=True, skip=None, p=0.0,
                keep_in=None, keep_wh=None, keep_att=None, faithful=True):
    """SpGraphAttentionLayer.forward (layers.py:125-173) for one head.

    x (N,F) W (F,D) a (1,2D) edge (2,E) int64; keep_* are 0/1 masks of shapes
    (N,F), (N,D), (E,) or None (eval / p == 0).
    """
    n = x.shape[0]
    d = W.shape[1]
    h = _drop(x, keep_in, p)                                   # :132
    wh = _drop(h.mm(W), keep_wh, p)                            # :134-136
    cat = torch.cat((wh[edge[0]], wh[edge[1]]), dim=1).t()     # :141
    logit = _leaky(a.reshape(1, 2 * d).mm(cat).squeeze(0), alpha)  # :144
    row_max = segment_max(logit, edge[0])                      # :145
    ex = torch.exp(logit - row_max[edge[0]])                   # :146
    ones = torch.ones(n, 1, dtype=x.dtype)
    rowsum = coo_matmul(edge, ex, torch.Size([n, n]), ones, faithful)   # :150
    ex = _drop(ex, keep_att, p)                                # :153
    out = coo_matmul(edge, ex, torch.Size([n, n]), wh, faithful)        # :156
    out = out.div(rowsum)                                      # :160
    if skip is not None:
        out = out + h.mm(skip)                                 # :166
    return torch.nn.functional.elu(out) if concat else out     # :168-173


def dense_head(x, W, a, adj, alpha, concat=True, skip=None, p=0.0,
               keep_in=None, keep_wh=None, keep_att=None):
    """GraphAttentionLayer.forward (layers.py:32-64) for one head.

    a is (2D,1); keep_att is the N x N mask applied to the normalised attention.
    """
    d = W.shape[1]
    h = _drop(x, keep_in, p)                                   # :34
    wh = _drop(h.mm(W), keep_wh, p)                            # :35-37
    a = a.reshape(2 * d, 1)
    e = _leaky(wh.matmul(a[:d]) + wh.matmul(a[d:]).t(), alpha)  # :60-64
    att = torch.where(adj > 0, e, torch.full_like(e, NEG_FILL))  # :40-41
    att = torch.softmax(att, dim=1)                            # :42
    att = _drop(att, keep_att, p)                              # :43
    out = att.matmul(wh)                                       # :44
    if skip is not None:
        out = out + h.mm(skip)                                 # :48
    return torch.nn.functional.elu(out) if concat else out     # :50-53


# --------------------------------------------------------------------------- GATv2 flavours
def sparse_head_v2(x, W, a, edge, alpha, concat=True, skip=None, faithful=True):
    """SpGraphAttentionLayerV2.forward (layers.py:255-313) for one head, eval mode / p = 0.

    x (N,F) W (2F,D) a (1,D) edge (2,E).  The score needs a D-wide operation per stored entry,
    e_ij = a . LeakyReLU(Whi_i + Whj_j) -- it does not split into f_i + g_j -- and the aggregated rows are
    the FIRST projection of the source, Whi_j (layers.py:295)."""
    n, f_in = x.shape
    whi = x.mm(W[:f_in])                                       # :265
    whj = x.mm(W[f_in:])                                       # :266
    edge_h = (whi[edge[0]] + whj[edge[1]]).t()                 # :275
    logit = a.reshape(1, -1).mm(_leaky(edge_h, alpha)).squeeze(0)  # :278
    row_max = segment_max(logit, edge[0])                      # :280
    ex = torch.exp(logit - row_max[edge[0]])                   # :281
    ones = torch.ones(n, 1, dtype=x.dtype)
    rowsum = coo_matmul(edge, ex, torch.Size([n, n]), ones, faithful)   # :285
    out = coo_matmul(edge, ex, torch.Size([n, n]), whi, faithful)       # :291
    out = out.div(rowsum)                                      # :295
    if skip is not None:
        out = out + x.mm(skip)                                 # :301
    return torch.nn.functional.elu(out) if concat else out     # :303-308


def dense_head_v2(x, W, a, adj, alpha, concat=True, skip=None):
    """GraphAttentionLayerV2.forward (layers.py:203-229) for one head, eval mode / p = 0.

    a is (D,1).  As shipped, the score is a per-node COLUMN (N,1) that `torch.where` broadcasts along each
    row (layers.py:214-217): every stored entry of row i gets the same score, i.e. attention is uniform
    over a node's neighbourhood, and the aggregated rows are the second projection (layers.py:220)."""
    f_in = x.shape[1]
    wh1 = x.mm(W[:f_in])                                       # :207
    wh2 = x.mm(W[f_in:])                                       # :208
    e = _leaky(wh1 + wh2, alpha).matmul(a.reshape(-1, 1))      # :212-214
    att = torch.where(adj > 0, e, torch.full_like(e, NEG_FILL))  # :216-217 (broadcast of the (N,1) column)
    att = torch.softmax(att, dim=1)                            # :218
    out = att.matmul(wh2)                                      # :220
    if skip is not None:
        out = out + x.mm(skip)                                 # :224
    return torch.nn.functional.elu(out) if concat else out     # :226-229


# --------------------------------------------------------------------------- model
def init_head(f_in: int, d: int, kind: str, skip: bool, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Parameters of one head with the reference's init calls in the reference's
    order (dense: layers.py:21-28, sparse: layers.py:111-119), so that a shared
    torch seed yields identical weights."""
    out: Dict[str, torch.Tensor] = {}
    if kind == "dense":
        w = torch.empty(f_in, d)
        torch.nn.init.xavier_uniform_(w, gain=1.414)
        a = torch.empty(2 * d, 1)
        torch.nn.init.xavier_uniform_(a, gain=1.414)
    elif kind == "sparse":
        w = torch.zeros(f_in, d)
        torch.nn.init.xavier_normal_(w, gain=1.414)
        a = torch.zeros(1, 2 * d)
        torch.nn.init.xavier_normal_(a, gain=1.414)
    else:
        raise ValueError(kind)
    out["W"], out["a"] = w.to(dtype), a.to(dtype)
    if skip:
        s = torch.empty(f_in, d)
        torch.nn.init.xavier_uniform_(s, gain=1.414)
        out["skip_projection"] = s.to(dtype)
    return out


def init_gat(nfeat: Sequence[int], nheads: Sequence[int], nlayers: int, kind: str,
             skip: bool, dtype=torch.float32) -> List[List[Dict[str, torch.Tensor]]]:
    """models.py:15-27: layer i head j sees in_features = nfeat[i] * nheads'[i]
    with nheads' = [1] + nheads; creation order is layer-major, head-minor."""
    hh = [1] + list(nheads)
    return [[init_head(nfeat[i] * hh[i], nfeat[i + 1], kind, skip, dtype) for _ in range(hh[i + 1])]
            for i in range(nlayers)]


def gat_forward(params, x, adj, alpha, kind="sparse", p=0.0, masks=None, faithful=True):
    """models.py:29-35.  Hidden layers concatenate their heads (ELU inside each
    head), the last layer averages them (no ELU).  ``masks[i][j]`` is a dict with
    optional keep_in / keep_wh / keep_att for layer i head j."""
    edge = edge_list(adj, "nonzero") if kind == "sparse" else None
    nl = len(params)
    for i, heads in enumerate(params):
        last = i == nl - 1
        outs = []
        for j, hp in enumerate(heads):
            mk = masks[i][j] if masks is not None else {}
            if kind == "sparse":
                o = sparse_head(x, hp["W"], hp["a"], edge, alpha, concat=not last,
                                skip=hp.get("skip_projection"), p=p, faithful=faithful, **mk)
            else:
                o = dense_head(x, hp["W"], hp["a"], adj, alpha, concat=not last,
                               skip=hp.get("skip_projection"), p=p, **mk)
            outs.append(o)
        x = torch.mean(torch.stack(outs, dim=1), dim=1) if last else torch.cat(outs, dim=1)
    return x


# --------------------------------------------------------------------------- synthetic graphs
def power_law_edges(n: int, avg_deg: float, seed: int, exponent: float = 0.5):
    """Synthetic power-law graph in the reference's adjacency convention: one endpoint drawn
    Zipf(exponent) over node rank by inverse CDF, the other uniform; symmetrised,
    de-duplicated, one self-loop per node, sorted row-major (what ``utils.load_data`` +
    ``adj.nonzero()`` would yield, utils.py:49-55, layers.py:129).  Same recipe as
    pygat_b200.synth.power_law_csr.  Returns rowptr int64, col int32."""
    g = torch.Generator().manual_seed(seed)
    m = int(n * max(avg_deg - 1.0, 0.0) / 2.0)
    u = torch.rand(m, generator=g, dtype=torch.float64)
    src = (u.pow(1.0 / (1.0 - exponent)) * n).long().clamp_(max=n - 1)
    dst = torch.randint(0, n, (m,), generator=g)
    r = torch.cat([src, dst, torch.arange(n)])
    c = torch.cat([dst, src, torch.arange(n)])
    key = torch.unique(r * n + c)  # sorted => row-major order
    r, c = key // n, key % n
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(r, minlength=n), 0)
    return rowptr, c.to(torch.int32)


class PatternAdj:
    """Duck-typed adjacency exposing only ``nonzero()`` (the single thing
    layers.py:129 reads), so the sparse oracle can run where N x N floats do not fit."""

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor):
        n = rowptr.numel() - 1
        row = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), rowptr[1:] - rowptr[:-1])
        self._nz = torch.stack([row, col.to(torch.int64)], dim=1)
        self.shape = (n, n)

    def nonzero(self):
        return self._nz


def xavier_std(fan_in: int, fan_out: int, gain: float = 1.414) -> float:
    return gain * math.sqrt(2.0 / (fan_in + fan_out))


# --------------------------------------------------------------------------- loss heads of the callers
def citation_loss_acc(logits: torch.Tensor, labels: torch.Tensor, idx: torch.Tensor):
    """train.py:151-152,159-160 and utils.py:92-96: nll_loss(log_softmax(elu(logits))[idx], labels[idx]) and the
    accuracy of the same rows."""
    out = torch.log_softmax(torch.nn.functional.elu(logits), dim=1)
    loss = torch.nn.functional.nll_loss(out[idx], labels[idx])
    preds = out[idx].max(1)[1].type_as(labels)
    acc = preds.eq(labels[idx]).sum() / len(labels[idx])
    return loss, acc


def ppi_loss_f1(logits: torch.Tensor, labels: torch.Tensor):
    """train_ppi.py:106-110,114,119-120: BCEWithLogitsLoss(mean) and sklearn's f1_score(gt, logits > 0, average='micro'),
    which for multilabel indicator input is 2 TP / (2 TP + FP + FN) over all (node, label) cells (0 when undefined)."""
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction="mean")
    pred = logits > 0
    pos = labels > 0.5
    tp = (pred & pos).sum().double()
    fp = (pred & ~pos).sum().double()
    fn = (~pred & pos).sum().double()
    denom = 2 * tp + fp + fn
    f1 = 2 * tp / denom if denom > 0 else torch.zeros((), dtype=torch.float64)
    return loss, f1
