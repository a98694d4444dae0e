"""GATv2 flavours of the reference (layers.py:179-316), kept importable because train.py:17 and
train_ppi.py:18 import all four layer names.  They are NOT on the accelerated path (SURVEY.md
section 8(f) ranks a fused GATv2 kernel as the next row): the math below is plain torch ops on the
caller's device, with the sparse variant using the engine's cached CSR edge list and O(E*D)
segment ops instead of the reference's per-call adj.nonzero() and dense N x N backward."""
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
        n = input.shape[0]
        dst, src = graph_of(adj, RULE_NONZERO).edge_index()
        h, left, right = self._project(input)
        score = self.leakyrelu(left[dst] + right[src]) @ self.a.reshape(-1)
        top = torch.zeros(n, dtype=score.dtype, device=score.device).scatter_reduce(
            0, dst, score.detach(), reduce="amax", include_self=False)
        ex = torch.exp(score - top[dst])
        denom = torch.zeros(n, dtype=ex.dtype, device=ex.device).index_add(0, dst, ex)
        ex = F.dropout(ex, self.dropout, training=self.training)
        agg = torch.zeros(n, self.out_features, dtype=ex.dtype, device=ex.device).index_add(
            0, dst, ex.unsqueeze(1) * left[src])
        return self._finish(agg / denom.unsqueeze(1), h)
