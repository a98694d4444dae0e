"""pygat_b200 -- B200-native (sm_100a) GAT layer engine behind the pyGAT operator API.

`layers` / `models` mirror the reference's modules; `functional.gat_layer` is the fused layer
call; `graph.Graph` is the cached CSR handle; the kernels live in libgatk.so (include/gatk.h).
"""
from . import _lib  # noqa: F401
from .functional import LayerMasks, gat_layer, pack_masks, padded_width  # noqa: F401
from .graph import Graph, graph_of  # noqa: F401
from .layers import GraphAttentionLayer, SpecialSpmm, SpecialSpmmFunction, SpGraphAttentionLayer  # noqa: F401
from .models import GAT  # noqa: F401

__version__ = "0.1.0"
