"""Stand-in for the reference's one un-vendored dependency (layers.py:5), used ONLY by
tests/golden/make_golden.py to import /root/reference unmodified in the build container.
torch_scatter.scatter_max(src, index) -> (out, argmax): 1-D segment max, length
index.max()+1, empty segments 0.  The reference discards argmax (layers.py:145)."""
import torch


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    size = int(index.max().item()) + 1 if dim_size is None else dim_size
    res = torch.zeros(size, dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(0, index, src, reduce="amax", include_self=False)
    return res, None
