"""The oracle restatement must reproduce the reference's own outputs and gradients
(fixtures generated from the unmodified reference by tests/golden/make_golden.py)."""
import pytest
import torch

from oracle import gat_oracle as O
from tests.golden_io import GAT_CASES, HEAD_CASES, V2_HEAD_CASES, dense_adj, gat_params, head_masks, load, rel_err

TOL = 2e-6  # same fp32 ops in the same order on the same CPU: only reduction-order noise


@pytest.mark.parametrize("name", HEAD_CASES)
def test_head_matches_reference(name):
    d = load(name)
    kind = "sparse" if name.startswith("sp_") else "dense"
    adj = dense_adj(d)
    x = d["x"].clone().requires_grad_(True)
    W = d["W"].clone().requires_grad_(True)
    a = d["a"].clone().requires_grad_(True)
    skip = d["skip"].clone().requires_grad_(True) if "skip" in d else None
    mk = head_masks(d, kind)
    if kind == "sparse":
        y = O.sparse_head(x, W, a, O.edge_list(adj), d["alpha"], bool(d["concat"]), skip, d["p"], **mk)
    else:
        y = O.dense_head(x, W, a, adj, d["alpha"], bool(d["concat"]), skip, d["p"], **mk)
    y.backward(d["gout"])
    assert rel_err(y, d["y"]) < TOL
    assert rel_err(x.grad, d["dx"]) < TOL
    assert rel_err(W.grad, d["dW"]) < TOL
    assert rel_err(a.grad, d["da"]) < TOL
    if skip is not None:
        assert rel_err(skip.grad, d["dskip"]) < TOL


@pytest.mark.parametrize("name", V2_HEAD_CASES)
def test_v2_head_matches_reference(name):
    """GATv2 flavours (layers.py:179-316): the checker for the next fused kernel."""
    d = load(name)
    adj = dense_adj(d)
    x = d["x"].clone().requires_grad_(True)
    W = d["W"].clone().requires_grad_(True)
    a = d["a"].clone().requires_grad_(True)
    skip = d["skip"].clone().requires_grad_(True) if "skip" in d else None
    if name.startswith("sp2_"):
        y = O.sparse_head_v2(x, W, a, O.edge_list(adj), d["alpha"], bool(d["concat"]), skip)
    else:
        y = O.dense_head_v2(x, W, a, adj, d["alpha"], bool(d["concat"]), skip)
    y.backward(d["gout"])
    assert rel_err(y, d["y"]) < TOL
    assert rel_err(x.grad, d["dx"]) < TOL
    assert rel_err(W.grad, d["dW"]) < TOL
    assert rel_err(a.grad, d["da"]) < TOL
    if skip is not None:
        assert rel_err(skip.grad, d["dskip"]) < TOL


@pytest.mark.parametrize("name", GAT_CASES)
def test_gat_matches_reference(name):
    d = load(name)
    kind = "sparse" if "_sp_" in name else "dense"
    adj = dense_adj(d)
    params = gat_params(d)
    leaves = []
    for layer in params:
        for hp in layer:
            for k in hp:
                hp[k] = hp[k].clone().requires_grad_(True)
                leaves.append(hp[k])
    x = d["x"].clone().requires_grad_(True)
    y = O.gat_forward(params, x, adj, d["alpha"], kind=kind, p=0.0)
    y.backward(d["gout"])
    assert rel_err(y, d["y"]) < TOL
    assert rel_err(x.grad, d["dx"]) < TOL
    for i, layer in enumerate(params):
        for j, hp in enumerate(layer):
            for k, v in hp.items():
                ref = d[f"grad.attention_layer_{i + 1}_head_{j + 1}.{k}"]
                assert rel_err(v.grad, ref) < TOL, (i, j, k)


def test_sparse_backward_variants_agree():
    """The O(E) backward used where N^2 does not fit equals the faithful N x N one."""
    d = load("sp_head_basic")
    adj = dense_adj(d)
    grads = []
    for faithful in (True, False):
        x = d["x"].clone().requires_grad_(True)
        W = d["W"].clone().requires_grad_(True)
        a = d["a"].clone().requires_grad_(True)
        y = O.sparse_head(x, W, a, O.edge_list(adj), d["alpha"], True, None, 0.0, faithful=faithful)
        y.backward(d["gout"])
        grads.append((x.grad, W.grad, a.grad))
    for g0, g1 in zip(*grads):
        assert rel_err(g1, g0) < TOL


def test_csr_helpers_roundtrip():
    d = load("sp_head_neg_asym")
    adj = dense_adj(d)
    n = adj.shape[0]
    e = O.edge_list(adj)
    assert torch.equal(e.t().to(torch.int32), d["edge"])
    assert O.edge_list(adj, "positive").shape[1] < e.shape[1]  # negatives dropped by the dense rule
    rowptr, col = O.csr_from_edges(e, n)
    assert rowptr[-1].item() == e.shape[1]
    tptr, trow, perm = O.csr_transpose(rowptr, col, n)
    # transposing twice gives the original pattern back
    tt_ptr, tt_row, _ = O.csr_transpose(tptr, trow, n)
    assert torch.equal(tt_ptr, rowptr) and torch.equal(tt_row, col)
    assert torch.equal(col[perm].long(), torch.repeat_interleave(torch.arange(n), tptr[1:] - tptr[:-1]))


def test_power_law_generator_shape():
    rowptr, col = O.power_law_edges(2000, 12.0, seed=72)
    n = 2000
    deg = rowptr[1:] - rowptr[:-1]
    assert rowptr[-1].item() == col.numel() and deg.min().item() >= 1
    row = torch.repeat_interleave(torch.arange(n), deg)
    key = row * n + col.long()
    assert torch.all(key[1:] > key[:-1])                       # sorted, no duplicates
    assert torch.equal(torch.unique(col.long() * n + row), key)  # symmetric
    assert deg.max().item() > 10 * deg.float().mean().item()   # heavy tail
