"""Drop-in `models.GAT` (reference models.py:7-35): same constructor, module names
(`attention_layer_{i}_head_{j}`, so state_dicts interchange) and forward(x, adj); the heads of
a layer are executed as one batched engine call instead of a Python loop."""
from __future__ import annotations

import torch
import torch.nn as nn

from .layers import GraphAttentionLayer, can_fuse, fused_heads
from .layers_v2 import can_fuse_v2, fused_heads_v2


class GAT(nn.Module):
    def __init__(self, nfeat, nheads, nlayers, dropout, alpha, layer_type=GraphAttentionLayer, skip_connection=False):
        super().__init__()
        self.dropout = dropout
        widths = [1] + list(nheads)
        self.gat_layers = []
        for i in range(nlayers):
            heads = []
            for j in range(widths[i + 1]):
                head = layer_type(in_features=nfeat[i] * widths[i], out_features=nfeat[i + 1], dropout=dropout,
                                  alpha=alpha, concat=i < nlayers - 1, skip_connection=skip_connection)
                heads.append(head)
                self.add_module('attention_layer_{}_head_{}'.format(i + 1, j + 1), head)
            self.gat_layers.append(heads)

    def forward(self, x, adj):
        last = len(self.gat_layers) - 1
        for i, heads in enumerate(self.gat_layers):
            if can_fuse(heads):
                x = fused_heads(heads, x, adj, combine="mean" if i == last else "cat")
            elif can_fuse_v2(heads):
                x = fused_heads_v2(heads, x, adj, combine="mean" if i == last else "cat")
            elif i < last:
                x = torch.cat([att(x, adj) for att in heads], dim=1)
            else:
                x = torch.mean(torch.stack([att(x, adj) for att in heads], dim=1), dim=1)
        return x
