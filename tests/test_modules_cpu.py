"""Host-side mirror of the reference interface: names, shapes, init order, error behaviour."""
import pytest
import torch

import layers
import models
from pygat_b200.functional import pack_heads, padded_width
from pygat_b200.layers import can_fuse
from tests.golden_io import load

# seeds used by tests/golden/make_golden.py
GAT_SEEDS = {"gat_sp_pubmed_like": 31, "gat_de_cora_like": 32, "gat_de_ppi_like": 33, "gat_sp_ppi_like": 34,
             "gat_sp_cora_topology": 72}
HEAD_SEEDS = {"sp_head_basic": 11, "sp_head_skip_last": 12, "de_head_basic": 21, "de_head_skip_last": 22}


@pytest.mark.parametrize("name", sorted(GAT_SEEDS))
def test_gat_same_seed_same_parameters_as_reference(name):
    d = load(name)
    cls = layers.SpGraphAttentionLayer if "_sp_" in name else layers.GraphAttentionLayer
    torch.manual_seed(GAT_SEEDS[name])
    nheads = [int(v) for v in d["nheads"]]
    model = models.GAT(nfeat=[int(v) for v in d["nfeat"]], nheads=nheads, nlayers=len(nheads), dropout=d["p"],
                       alpha=d["alpha"], layer_type=cls, skip_connection=bool(d["skip"]))
    ref = {k[len("param."):]: v for k, v in d.items() if k.startswith("param.")}
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys())  # models.py:27 naming and order
    for k, v in ref.items():
        assert sd[k].shape == v.shape
        assert torch.equal(sd[k], v), k
    model.load_state_dict(ref)  # reference checkpoints load


@pytest.mark.parametrize("name", sorted(HEAD_SEEDS))
def test_head_same_seed_same_parameters_as_reference(name):
    d = load(name)
    cls = layers.SpGraphAttentionLayer if name.startswith("sp_") else layers.GraphAttentionLayer
    torch.manual_seed(HEAD_SEEDS[name])
    f_in, dd = d["W"].shape
    head = cls(f_in, dd, dropout=d["p"], alpha=d["alpha"], concat=bool(d["concat"]), skip_connection="skip" in d)
    assert torch.equal(head.W.data, d["W"]) and torch.equal(head.a.data, d["a"])
    if "skip" in d:
        assert torch.equal(head.skip_projection.data, d["skip"])
    assert repr(head) == f"{cls.__name__} ({f_in} -> {dd})"
    assert isinstance(head.leakyrelu, torch.nn.LeakyReLU)


def test_all_four_layer_names_importable():
    for n in ("GraphAttentionLayer", "GraphAttentionLayerV2", "SpGraphAttentionLayer", "SpGraphAttentionLayerV2",
              "SpecialSpmm", "SpecialSpmmFunction"):
        assert hasattr(layers, n)
    v2 = layers.SpGraphAttentionLayerV2(6, 4, 0.0, 0.2)
    assert v2.W.shape == (12, 4) and v2.a.shape == (1, 4)
    assert layers.GraphAttentionLayerV2(6, 4, 0.0, 0.2).a.shape == (4, 1)


def test_cpu_tensors_are_refused_loudly():
    head = layers.SpGraphAttentionLayer(5, 4, 0.0, 0.2)
    with pytest.raises(RuntimeError, match="CUDA"):
        head(torch.randn(6, 5), torch.eye(6))
    with pytest.raises(RuntimeError, match="CUDA"):
        layers.SpecialSpmm()(torch.zeros(2, 1, dtype=torch.long), torch.ones(1), torch.Size([3, 3]), torch.ones(3, 2))


def test_can_fuse_rules():
    a = [layers.SpGraphAttentionLayer(5, 4, 0.1, 0.2) for _ in range(3)]
    assert can_fuse(a)
    a[1].eval()
    assert not can_fuse(a)
    assert not can_fuse([layers.SpGraphAttentionLayer(5, 4, 0.1, 0.2), layers.GraphAttentionLayer(5, 4, 0.1, 0.2)])
    assert not can_fuse([layers.SpGraphAttentionLayerV2(5, 4, 0.1, 0.2)])


def test_padded_width_and_packing():
    assert [padded_width(d) for d in (1, 3, 4, 7, 8, 64, 100, 121, 256)] == [4, 4, 4, 8, 8, 64, 128, 128, 256]
    heads = [layers.GraphAttentionLayer(6, 7, 0.0, 0.2, skip_connection=True) for _ in range(3)]
    vec = [h.attention_vectors() for h in heads]
    w_ext, a_src, a_dst, D, Dp = pack_heads([h.W for h in heads], [v[0] for v in vec], [v[1] for v in vec],
                                            [h.skip_projection for h in heads])
    assert (D, Dp) == (7, 8) and w_ext.shape == (6, 48) and a_src.shape == (3, 8)
    assert torch.equal(w_ext[:, 8:15], heads[1].W) and torch.all(w_ext[:, 15] == 0)
    assert torch.equal(w_ext[:, 24:31], heads[0].skip_projection)
    assert torch.equal(a_src[2, :7], heads[2].a[:7, 0]) and torch.equal(a_dst[2, :7], heads[2].a[7:, 0])
    w_ext.sum().backward()  # gradients flow back to the per-head parameters
    assert heads[0].W.grad is not None and heads[2].skip_projection.grad is not None


# ------------------------------------------------------------------------------ GATv2 flavours
V2_SEEDS = {"sp2_head_basic": 41, "sp2_head_skip_last": 42, "sp2_head_hub": 43, "de2_head_basic": 44,
            "de2_head_skip_last": 45}


def _v2_head(d, name, seed=None):
    cls = layers.SpGraphAttentionLayerV2 if name.startswith("sp2_") else layers.GraphAttentionLayerV2
    two_f, dd = d["W"].shape
    if seed is not None:
        torch.manual_seed(seed)
    return cls(two_f // 2, dd, dropout=d["p"], alpha=d["alpha"], concat=bool(d["concat"]), skip_connection="skip" in d)


@pytest.mark.parametrize("name", sorted(V2_SEEDS))
def test_v2_head_same_seed_same_parameters_as_reference(name):
    d = load(name)
    head = _v2_head(d, name, V2_SEEDS[name])
    assert torch.equal(head.W.data, d["W"]) and torch.equal(head.a.data, d["a"])  # layers.py:190-195, 246-251
    if "skip" in d:
        assert torch.equal(head.skip_projection.data, d["skip"])


@pytest.mark.parametrize("name", [n for n in sorted(V2_SEEDS) if n.startswith("de2_")])
def test_dense_v2_head_matches_reference_golden(name):
    """The dense GATv2 class is plain torch ops on the caller's device (not on the accelerated path): checked on
    CPU against the unmodified reference's outputs and gradients (layers.py:203-229)."""
    from tests.golden_io import dense_adj, rel_err
    d = load(name)
    head = _v2_head(d, name)
    with torch.no_grad():
        head.W.copy_(d["W"])
        head.a.copy_(d["a"])
        if "skip" in d:
            head.skip_projection.copy_(d["skip"])
    head.train(bool(d["train"]))
    x = d["x"].clone().requires_grad_(True)
    y = head(x, dense_adj(d))
    y.backward(d["gout"])
    errs = {"y": rel_err(y, d["y"]), "dx": rel_err(x.grad, d["dx"]), "dW": rel_err(head.W.grad, d["dW"]),
            "da": rel_err(head.a.grad, d["da"])}
    if "skip" in d:
        errs["dskip"] = rel_err(head.skip_projection.grad, d["dskip"])
    assert all(v < 1e-5 for v in errs.values()), errs


def test_dropout_site_offsets_do_not_overlap():
    """The three dropout sites of a layer draw from disjoint counter ranges of one Philox stream
    (functional.site_offsets): a site of sz elements owns ceil(sz / 4) counters."""
    from pygat_b200.functional import site_offsets
    for n, f_in, H, Dp, nnz in [(7, 5, 3, 8, 19), (2708, 1433, 8, 8, 13264), (1, 1, 1, 4, 1)]:
        offs, sizes = site_offsets(n, f_in, H, Dp, nnz)
        assert sizes == (H * n * f_in, n * H * Dp, nnz * H) and offs[0] == 0
        ends = [o + (s + 3) // 4 for o, s in zip(offs, sizes)]
        assert offs[1] == ends[0] and offs[2] == ends[1]


def test_head_chunks_divide_the_heads(monkeypatch):
    """sharded.head_chunks: heads per exchange chunk; one chunk up to two ranks, two beyond, the override clipped to
    a divisor of H."""
    from pygat_b200.sharded import head_chunks
    monkeypatch.delenv("GATK_SHARD_CHUNKS", raising=False)
    assert head_chunks(8, 2) == 8 and head_chunks(8, 8) == 4 and head_chunks(1, 8) == 1 and head_chunks(3, 8) == 3
    monkeypatch.setenv("GATK_SHARD_CHUNKS", "4")
    assert head_chunks(8, 2) == 2 and head_chunks(6, 8) == 2 and head_chunks(2, 8) == 1
    monkeypatch.setenv("GATK_SHARD_CHUNKS", "100")
    assert head_chunks(8, 8) == 1


def test_long_row_count_of_the_byte_model():
    from benchmarks.layer import XBWD_SHORT_ROW, long_rows
    rowptr = torch.tensor([0, 1, 1 + XBWD_SHORT_ROW, 2 + 2 * XBWD_SHORT_ROW, 2 + 2 * XBWD_SHORT_ROW])
    assert long_rows(rowptr) == 1   # rows of 1, 32, 33 and 0 entries
