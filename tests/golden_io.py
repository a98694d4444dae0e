"""Load the fixtures written by tests/golden/make_golden.py."""
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_ALL_HEADS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*_head_*.npz")))
HEAD_CASES = [n for n in _ALL_HEADS if n.startswith(("sp_", "de_"))]        # GAT heads: the accelerated path
V2_HEAD_CASES = [n for n in _ALL_HEADS if n.startswith(("sp2_", "de2_"))]   # GATv2 heads (layers.py:179-316)
GAT_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "gat_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {}
    for k in z.files:
        v = z[k]
        d[k] = torch.from_numpy(v) if v.ndim > 0 else v.item()
    return d


def dense_adj(d):
    n = int(d["n"])
    adj = torch.zeros(n, n)
    e = d["edge"].long()
    adj[e[:, 0], e[:, 1]] = d["val"]
    return adj


def unpack_mask(bits, shape):
    n = int(np.prod(shape))
    return torch.from_numpy(np.unpackbits(bits.numpy())[:n].reshape(shape).astype(np.bool_))


def head_masks(d, kind):
    if "keep_in" not in d:
        return {}
    n, f = d["x"].shape
    dd = d["W"].shape[1]
    e = d["edge"].shape[0]
    att_shape = (e,) if kind == "sparse" else (n, n)
    return {"keep_in": unpack_mask(d["keep_in"], (n, f)), "keep_wh": unpack_mask(d["keep_wh"], (n, dd)),
            "keep_att": unpack_mask(d["keep_att"], att_shape)}


def gat_params(d):
    """[[{W,a,skip_projection}]] in models.py:27 naming order."""
    nheads = [int(v) for v in d["nheads"]]
    params = []
    for i, h in enumerate(nheads):
        layer = []
        for j in range(h):
            pre = f"attention_layer_{i + 1}_head_{j + 1}."
            hp = {"W": d["param." + pre + "W"], "a": d["param." + pre + "a"]}
            if "param." + pre + "skip_projection" in d:
                hp["skip_projection"] = d["param." + pre + "skip_projection"]
            layer.append(hp)
        params.append(layer)
    return params


def rel_err(got, ref):
    """max-abs-diff / max-abs-ref (SURVEY.md section 8(c) tolerance definition)."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    denom = ref.abs().max().item()
    return (got - ref).abs().max().item() / (denom if denom > 0 else 1.0)
