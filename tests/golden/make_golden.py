"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports /root/reference/{layers,models}.py with the torch_scatter shim in
tests/golden/shims, runs seeded forward + backward passes on CPU in fp32 and stores
inputs, parameters, outputs and every gradient as compressed .npz files.  The fixtures
travel to the GPU box; the reference does not.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYGAT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, REF)

import layers as ref_layers  # noqa: E402  (the reference's own file)
import models as ref_models  # noqa: E402


def rand_graph(n, avg_deg, seed, symmetric=True, negatives=0, dense_vals=True):
    g = torch.Generator().manual_seed(seed)
    m = int(n * avg_deg)
    r = torch.randint(0, n, (m,), generator=g)
    c = torch.randint(0, n, (m,), generator=g)
    adj = torch.zeros(n, n)
    adj[r, c] = torch.rand(m, generator=g) + 0.1
    if symmetric:
        adj = torch.maximum(adj, adj.t())
    adj = adj + torch.eye(n)
    if negatives:
        rr = torch.randint(0, n, (negatives,), generator=g)
        cc = torch.randint(0, n, (negatives,), generator=g)
        off = rr != cc
        adj[rr[off], cc[off]] = -0.5
    return adj


def hub_graph(n, seed):
    """A few very high degree rows plus a sparse background (exercises row splitting)."""
    adj = rand_graph(n, 2.0, seed)
    adj[0, :] = 1.0
    adj[:, 0] = 1.0
    adj[3, : n // 2] = 1.0
    adj[: n // 2, 3] = 1.0
    return adj


def pack(adj):
    nz = adj.nonzero()
    return {"n": np.int64(adj.shape[0]), "edge": nz.numpy().astype(np.int32),
            "val": adj[nz[:, 0], nz[:, 1]].numpy().astype(np.float32)}


def save(name, d):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in d.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def replay_masks(seed, shapes, p):
    """Masks F.dropout will draw, in call order, for tensors of the given shapes."""
    torch.manual_seed(seed)
    return [(F.dropout(torch.ones(s), p, training=True) != 0) for s in shapes]


def head_case(name, cls, n, f_in, d, adj, concat, skip, seed, p=0.0, train=False, alpha=0.2):
    torch.manual_seed(seed)
    layer = cls(f_in, d, dropout=p, alpha=alpha, concat=concat, skip_connection=skip)
    x = torch.randn(n, f_in, requires_grad=True)
    gout = torch.randn(n, d)
    out = {"alpha": np.float32(alpha), "p": np.float32(p), "concat": np.int64(concat), "train": np.int64(train)}
    out.update(pack(adj))
    layer.train(train)
    if train and p > 0:
        e = int(adj.nonzero().shape[0])
        att_shape = (e,) if cls is ref_layers.SpGraphAttentionLayer else (n, n)
        masks = replay_masks(seed + 1000, [(n, f_in), (n, d), att_shape], p)
        out["keep_in"] = np.packbits(masks[0].numpy())
        out["keep_wh"] = np.packbits(masks[1].numpy())
        out["keep_att"] = np.packbits(masks[2].numpy())
        torch.manual_seed(seed + 1000)
    y = layer(x, adj)
    y.backward(gout)
    out.update({"x": x, "gout": gout, "y": y, "dx": x.grad, "W": layer.W, "a": layer.a,
                "dW": layer.W.grad, "da": layer.a.grad})
    if skip:
        out.update({"skip": layer.skip_projection, "dskip": layer.skip_projection.grad})
    save(name, out)


def gat_case(name, nfeat, nheads, cls, adj, skip, seed, p, train, alpha=0.2):
    torch.manual_seed(seed)
    model = ref_models.GAT(nfeat=nfeat, nheads=nheads, nlayers=len(nheads), dropout=p, alpha=alpha,
                           layer_type=cls, skip_connection=skip)
    n = adj.shape[0]
    x = torch.randn(n, nfeat[0], requires_grad=True)
    gout = torch.randn(n, nfeat[-1])
    model.train(train)
    y = model(x, adj)
    y.backward(gout)
    out = {"alpha": np.float32(alpha), "p": np.float32(p), "train": np.int64(train), "skip": np.int64(skip),
           "nfeat": np.array(nfeat), "nheads": np.array(nheads),
           "x": x, "gout": gout, "y": y, "dx": x.grad}
    out.update(pack(adj))
    for k, v in model.named_parameters():
        out["param." + k] = v
        out["grad." + k] = v.grad
    save(name, out)


def cora_adj():
    cites = np.genfromtxt(os.path.join(REF, "data/cora/cora.cites"), dtype=np.int64)
    ids = np.unique(cites)
    remap = {v: i for i, v in enumerate(ids)}
    e = np.vectorize(remap.get)(cites)
    n = len(ids)
    adj = torch.zeros(n, n)
    adj[e[:, 0], e[:, 1]] = 1.0
    adj = torch.maximum(adj, adj.t()) + torch.eye(n)
    return adj


def v2_cases():
    """GATv2 flavours (layers.py:179-316): eval mode and p = 0, as for the other parity cases."""
    sp2, de2 = ref_layers.SpGraphAttentionLayerV2, ref_layers.GraphAttentionLayerV2
    a96 = rand_graph(96, 3.0, 1)
    head_case("sp2_head_basic", sp2, 96, 24, 8, a96, True, False, 41)
    head_case("sp2_head_skip_last", sp2, 96, 24, 7, a96, False, True, 42)
    head_case("sp2_head_hub", sp2, 300, 12, 16, hub_graph(300, 3), True, True, 43)
    head_case("de2_head_basic", de2, 96, 24, 8, a96, True, False, 44)
    head_case("de2_head_skip_last", de2, 96, 24, 7, a96, False, True, 45)


def isolated_graph(n, seed):
    """Rows without any `adj > 0` entry (no self-loop either): the dense class attends uniformly to ALL nodes there
    (softmax of an all -9e15 row, layers.py:40-42).  Row 9 keeps only negative entries."""
    adj = rand_graph(n, 3.0, seed)
    for r in (5, 17, n - 1):
        adj[r, :] = 0.0
    adj[9, :] = 0.0
    adj[9, 2] = adj[9, 40] = -0.7
    return adj


def isolated_cases():
    de = ref_layers.GraphAttentionLayer
    iso = isolated_graph(96, 9)
    head_case("de_head_isolated", de, 96, 24, 8, iso, True, True, 51)
    head_case("de_head_isolated_last", de, 96, 24, 7, iso, False, False, 52)
    gat_case("gat_de_isolated", [12, 8, 5], [4, 3], de, iso, True, 53, 0.6, False)


def main():
    if os.environ.get("GOLDEN_ONLY") == "v2":  # add the GATv2 fixtures without rewriting the others
        return v2_cases()
    if os.environ.get("GOLDEN_ONLY") == "isolated":
        return isolated_cases()
    sp, de = ref_layers.SpGraphAttentionLayer, ref_layers.GraphAttentionLayer
    a96 = rand_graph(96, 3.0, 1)
    head_case("sp_head_basic", sp, 96, 24, 8, a96, True, False, 11)
    head_case("sp_head_skip_last", sp, 96, 24, 7, a96, False, True, 12)
    head_case("sp_head_neg_asym", sp, 80, 10, 12, rand_graph(80, 2.5, 2, symmetric=False, negatives=40),
              True, False, 13)
    head_case("sp_head_hub", sp, 700, 12, 16, hub_graph(700, 3), True, True, 14)
    head_case("sp_head_wide", sp, 64, 50, 256, rand_graph(64, 4.0, 4), True, True, 15)
    head_case("sp_head_train_p06", sp, 96, 24, 8, a96, True, False, 16, p=0.6, train=True)
    head_case("de_head_basic", de, 96, 24, 8, a96, True, False, 21)
    head_case("de_head_skip_last", de, 96, 24, 7, a96, False, True, 22)
    head_case("de_head_neg_asym", de, 80, 10, 12, rand_graph(80, 2.5, 2, symmetric=False, negatives=40),
              True, False, 23)
    head_case("de_head_train_p06", de, 96, 24, 8, a96, True, False, 24, p=0.6, train=True)
    gat_case("gat_sp_pubmed_like", [20, 8, 3], [8, 8], sp, rand_graph(150, 3.0, 5), False, 31, 0.6, False)
    gat_case("gat_de_cora_like", [30, 8, 7], [8, 1], de, rand_graph(120, 3.0, 6), False, 32, 0.6, False)
    blk = torch.block_diag(rand_graph(70, 4.0, 7), rand_graph(50, 4.0, 8))
    gat_case("gat_de_ppi_like", [10, 32, 32, 11], [4, 4, 6], de, blk, True, 33, 0.0, True)
    gat_case("gat_sp_ppi_like", [10, 32, 32, 11], [4, 4, 6], sp, blk, True, 34, 0.0, True)
    gat_case("gat_sp_cora_topology", [16, 8, 7], [8, 1], sp, cora_adj(), False, 72, 0.6, False)
    v2_cases()
    isolated_cases()


if __name__ == "__main__":
    main()
