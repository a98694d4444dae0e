"""Drop-in operator API of the reference's layers.py, executed by the B200 engine.

Same class names, constructor arguments, parameter names / shapes / init calls (so a shared
seed gives identical weights and reference checkpoints load), same forward signatures and
dropout / alpha / concat / skip semantics -- but forward is ONE fused CSR pass per layer
through libgatk.so instead of ~25 ATen launches per head, and backward is O(E*D) instead of
the reference's dense N x N (layers.py:85).  CUDA tensors only: there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .functional import gat_layer
from .graph import RULE_NONZERO, RULE_POSITIVE, _ptr, _require_cuda, _stream, graph_of


class _EngineHead(nn.Module):
    """State shared by both head flavours (reference layers.py:12-30 / :103-123)."""
    PATTERN_RULE = RULE_NONZERO

    def _setup(self, in_features, out_features, dropout, alpha, concat, skip_connection):
        self.dropout = dropout
        self.in_features = in_features
        self.out_features = out_features
        self.alpha = alpha
        self.concat = concat
        self.skip_connection = skip_connection

    def _maybe_skip(self):
        if self.skip_connection:
            self.skip_projection = nn.Parameter(torch.empty(size=(self.in_features, self.out_features)))
            nn.init.xavier_uniform_(self.skip_projection.data, gain=1.414)

    def attention_vectors(self):
        """(a_src, a_dst): the halves of `a` that multiply Wh_i and Wh_j."""
        a = self.a.reshape(-1)
        return a[: self.out_features], a[self.out_features:]

    def forward(self, h, adj):
        return fused_heads([self], h, adj, combine="cat")

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


class GraphAttentionLayer(_EngineHead):
    """Dense-adjacency GAT head (reference layers.py:8-67): neighbours are `adj > 0`, `a` is (2D, 1)."""
    PATTERN_RULE = RULE_POSITIVE

    def __init__(self, in_features, out_features, dropout, alpha, concat=True, skip_connection=False):
        super().__init__()
        self._setup(in_features, out_features, dropout, alpha, concat, skip_connection)
        self.W = nn.Parameter(torch.empty(size=(in_features, out_features)))
        nn.init.xavier_uniform_(self.W.data, gain=1.414)
        self.a = nn.Parameter(torch.empty(size=(2 * out_features, 1)))
        nn.init.xavier_uniform_(self.a.data, gain=1.414)
        self._maybe_skip()
        self.leakyrelu = nn.LeakyReLU(self.alpha)


class SpGraphAttentionLayer(_EngineHead):
    """Sparse GAT head (reference layers.py:98-176): neighbours are `adj.nonzero()`, `a` is (1, 2D)."""
    PATTERN_RULE = RULE_NONZERO

    def __init__(self, in_features, out_features, dropout, alpha, concat=True, skip_connection=False):
        super().__init__()
        self._setup(in_features, out_features, dropout, alpha, concat, skip_connection)
        self.W = nn.Parameter(torch.zeros(size=(in_features, out_features)))
        nn.init.xavier_normal_(self.W.data, gain=1.414)
        self.a = nn.Parameter(torch.zeros(size=(1, 2 * out_features)))
        nn.init.xavier_normal_(self.a.data, gain=1.414)
        self._maybe_skip()
        self.leakyrelu = nn.LeakyReLU(self.alpha)
        self.special_spmm = SpecialSpmm()


def can_fuse(heads) -> bool:
    """True when a list of head modules can run as one batched layer call."""
    h0 = heads[0]
    if not isinstance(h0, _EngineHead):
        return False
    key = (type(h0), h0.in_features, h0.out_features, h0.dropout, h0.alpha, h0.concat, h0.skip_connection,
           h0.training)
    return all(isinstance(h, _EngineHead) and
               (type(h), h.in_features, h.out_features, h.dropout, h.alpha, h.concat, h.skip_connection,
                h.training) == key for h in heads)


def fused_heads(heads, x, adj, combine="cat", masks=None):
    """Run every head of a layer in one engine call (models.py:29-35 loops over them)."""
    h0 = heads[0]
    _require_cuda(x, "input features")
    graph = graph_of(adj, h0.PATTERN_RULE)
    vecs = [h.attention_vectors() for h in heads]
    skips = [h.skip_projection for h in heads] if h0.skip_connection else None
    out = gat_layer(x, graph, [h.W for h in heads], [v[0] for v in vecs], [v[1] for v in vecs], skips,
                    h0.alpha, h0.concat, p=h0.dropout, training=h0.training, masks=masks, combine=combine)
    if h0.PATTERN_RULE == RULE_POSITIVE:
        empty = graph.empty_rows()
        if empty is not None:
            out = _dense_rows_without_neighbours(out, empty, heads, x, skips, combine)
    return out


def _dense_rows_without_neighbours(out, empty, heads, x, skips, combine):
    """Dense class only: a row with no `adj > 0` entry is all -9e15 before the softmax (layers.py:40-42), so the
    reference attends UNIFORMLY to all N nodes and h'_i is the mean of Wh over every node (+ skip, ELU).  The CSR
    engine leaves such rows at 0 (+ skip); they are patched here with parameter-sized torch ops (the mean is linear:
    mean_j(x_j W) = mean_j(x_j) W), which also routes their gradients.  Evaluated without dropout -- the reference's
    random masks cannot be reproduced anyway.  Every loader of the reference adds self-loops, so this is rare."""
    h0 = heads[0]
    x = x.float()
    xm = x.mean(0, keepdim=True)
    per_head = []
    for k, h in enumerate(heads):
        v = (xm @ h.W).expand(empty.numel(), -1)
        if skips is not None:
            v = v + x[empty] @ skips[k]
        per_head.append(torch.nn.functional.elu(v) if h0.concat else v)
    if combine == "mean":
        repl = torch.stack(per_head, dim=1).mean(dim=1)
    elif combine == "cat":
        repl = torch.cat(per_head, dim=1)
    else:
        raise RuntimeError("rows without neighbours are only patched for combine='cat' / 'mean'")
    return out.index_copy(0, empty, repl)


# ---------------------------------------------------------------------- SpecialSpmm (layers.py:70-95)
class SpecialSpmmFunction(torch.autograd.Function):
    """COO(indices, values, shape) @ b with gradients only on the stored entries.  Backward is
    O(E*k): one dot product per stored entry instead of the reference's dense N x N product."""

    @staticmethod
    def forward(ctx, indices, values, shape, b):
        assert indices.requires_grad == False  # noqa: E712  (reference precondition, layers.py:74)
        _require_cuda(values, "values")
        idx = indices.contiguous()
        vals = values.contiguous().float()
        bb = b.contiguous().float()
        k = bb.shape[1]
        out = torch.zeros(shape[0], k, dtype=torch.float32, device=bb.device)
        _lib.call("gatk_spmm_coo_fwd", idx[0].data_ptr(), idx[1].data_ptr(), vals.data_ptr(), idx.shape[1], k,
                  bb.data_ptr(), out.data_ptr(), _stream())
        ctx.save_for_backward(idx, vals, bb)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        idx, vals, bb = ctx.saved_tensors
        go = grad_output.contiguous().float()
        k = bb.shape[1]
        grad_values = torch.empty_like(vals) if ctx.needs_input_grad[1] else None
        grad_b = torch.zeros_like(bb) if ctx.needs_input_grad[3] else None
        _lib.call("gatk_spmm_coo_bwd", idx[0].data_ptr(), idx[1].data_ptr(), vals.data_ptr(), idx.shape[1], k,
                  bb.data_ptr(), go.data_ptr(), _ptr(grad_values), _ptr(grad_b), _stream())
        return None, grad_values, None, grad_b


class SpecialSpmm(nn.Module):
    def forward(self, indices, values, shape, b):
        return SpecialSpmmFunction.apply(indices, values, shape, b)
