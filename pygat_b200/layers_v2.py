"""GATv2 flavours of the reference (layers.py:179-316; train.py:17 and train_ppi.py:18 import all four layer names).

SpGraphAttentionLayerV2 runs on the engine: one projection GEMM for [Whi | Whj | skip] and the fused CSR attention
kernels of csrc/attn_v2.cu (gatk_attn_v2_fwd / _bwd) instead of the reference's per-call adj.nonzero(), E x 2D
gathers and dense N x N backward.  The dense GraphAttentionLayerV2 is degenerate as shipped (its score is a per-node
column, so attention is uniform over a node's neighbours, layers.py:212-217): it is kept as plain torch ops for
import compatibility."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .graph import RULE_NONZERO, graph_of


class _V2Base(nn.Module):
    def _setup(self, in_features, out_features, dropout, alpha, concat, skip_connection, a_shape, normal):
        self.dropout, self.alpha, self.concat = dropout, alpha, concat
        self.in_features, self.out_features, self.skip_connection = in_features, out_features, skip_connection
        init = nn.init.xavier_normal_ if normal else nn.init.xavier_uniform_
        self.W = nn.Parameter(torch.empty(size=(2 * in_features, out_features)))
        init(self.W.data, gain=1.414)
        self.a = nn.Parameter(torch.zeros(size=a_shape))
        init(self.a.data, gain=1.414)
        if skip_connection:
            self.skip_projection = nn.Parameter(torch.empty(size=(in_features, out_features)))
            nn.init.xavier_uniform_(self.skip_projection.data, gain=1.414)
        self.leakyrelu = nn.LeakyReLU(alpha)

    def _project(self, x):
        h = F.dropout(x, self.dropout, training=self.training)
        left = F.dropout(h @ self.W[: self.in_features], self.dropout, training=self.training)
        right = F.dropout(h @ self.W[self.in_features:], self.dropout, training=self.training)
        return h, left, right

    def _finish(self, agg, h):
        if self.skip_connection:
            agg = agg + h @ self.skip_projection
        return F.elu(agg) if self.concat else agg

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


class GraphAttentionLayerV2(_V2Base):
    """Dense GATv2 as the reference computes it (layers.py:203-229): the score is a per-node
    column (N x 1) broadcast across each row of the mask, so attention is uniform over a node's
    neighbours; the aggregated features are the second projection."""

    def __init__(self, in_features, out_features, dropout, alpha, concat=True, skip_connection=False):
        super().__init__()
        self._setup(in_features, out_features, dropout, alpha, concat, skip_connection, (out_features, 1), False)

    def forward(self, h, adj):
        h, left, right = self._project(h)
        score = self.leakyrelu(left + right) @ self.a
        att = torch.where(adj > 0, score.expand(-1, adj.shape[1]), torch.full_like(adj, -9e15))
        att = F.dropout(torch.softmax(att, dim=1), self.dropout, training=self.training)
        return self._finish(att @ right, h)


class SpGraphAttentionLayerV2(_V2Base):
    """Sparse GATv2 (layers.py:255-313): score_ij = a . LeakyReLU(Whi_i + Whj_j), softmax over the
    stored entries of row i, aggregation of the first projection Whi_j (as the reference does)."""

    def __init__(self, in_features, out_features, dropout, alpha, concat=True, skip_connection=False):
        super().__init__()
        self._setup(in_features, out_features, dropout, alpha, concat, skip_connection, (1, out_features), True)

    def forward(self, input, adj):
        return fused_heads_v2([self], input, adj, combine="cat")


def can_fuse_v2(heads) -> bool:
    h0 = heads[0]
    if not isinstance(h0, SpGraphAttentionLayerV2):
        return False
    key = (h0.in_features, h0.out_features, h0.dropout, h0.alpha, h0.concat, h0.skip_connection, h0.training)
    return all(isinstance(h, SpGraphAttentionLayerV2) and
               (h.in_features, h.out_features, h.dropout, h.alpha, h.concat, h.skip_connection, h.training) == key
               for h in heads)


def fused_heads_v2(heads, x, adj, combine="cat"):
    """Every SpGraphAttentionLayerV2 head of a layer in one engine call (models.py:29-35 loops over them)."""
    from .functional import gat_v2_layer
    from .graph import _require_cuda
    h0 = heads[0]
    _require_cuda(x, "input features")
    graph = graph_of(adj, RULE_NONZERO)
    skips = [h.skip_projection for h in heads] if h0.skip_connection else None
    return gat_v2_layer(x, graph, [h.W for h in heads], [h.a.reshape(-1) for h in heads], skips, h0.alpha, h0.concat,
                        p=h0.dropout, training=h0.training, combine=combine)
