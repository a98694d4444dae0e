"""`from layers import ...` as train.py:17 / train_ppi.py:18 do -- served by the B200 engine."""
from pygat_b200.layers import (GraphAttentionLayer, SpecialSpmm, SpecialSpmmFunction,  # noqa: F401
                               SpGraphAttentionLayer)
from pygat_b200.layers_v2 import GraphAttentionLayerV2, SpGraphAttentionLayerV2  # noqa: F401
