"""Graph ingest without a dense adjacency: the patterns built from edge lists / scipy
matrices / per-graph pieces equal `adj.nonzero()` of the dense matrices the reference's loaders produce
(utils.py:30,49-55; load_data_ppi.py:84-86,149-153), restated here with scipy / dense torch as the checker."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from pygat_b200 import ingest


def _reference_dense_adj(adj):
    """utils.py:49-55 on a scipy matrix (normalize_adj restated: utils.py:73-79)."""
    adj = adj + adj.T.multiply(adj.T > adj) - adj.multiply(adj.T > adj)
    mx = adj + sp.eye(adj.shape[0])
    rowsum = np.array(mx.sum(1))
    with np.errstate(divide="ignore"):
        r = np.power(rowsum, -0.5).flatten()
    r[np.isinf(r)] = 0.0
    d = sp.diags(r)
    return torch.FloatTensor(np.array(mx.dot(d).transpose().dot(d).todense()))


def _edges_of(rowptr, col):
    deg = rowptr[1:] - rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(rowptr.numel() - 1), deg)
    return torch.stack([row, col.long()])


def _random_edges(n, m, seed):
    g = np.random.default_rng(seed)
    e = g.integers(0, n, size=(m, 2)).astype(np.int32)
    e = np.concatenate([e, e[: m // 10]])  # duplicates: coo_matrix sums them, still one stored entry
    return e


@pytest.mark.parametrize("n,m,seed", [(1, 0, 0), (7, 0, 1), (50, 120, 2), (300, 2500, 3)])
def test_edge_list_pattern_equals_reference_loader(n, m, seed):
    e = _random_edges(n, m, seed)
    adj = sp.coo_matrix((np.ones(e.shape[0]), (e[:, 0], e[:, 1])), shape=(n, n), dtype=np.float32)  # utils.py:30
    want = _reference_dense_adj(adj).nonzero().t()  # layers.py:129
    rowptr, col = ingest.pattern_from_edges(torch.from_numpy(e[:, 0]), torch.from_numpy(e[:, 1]), n)
    assert rowptr.dtype == torch.int64 and col.dtype == torch.int32 and rowptr.numel() == n + 1
    assert torch.equal(_edges_of(rowptr, col), want)
    # the scipy entry point agrees, including explicitly stored zeros (not edges)
    adj2 = adj.tocsr().astype(np.float32)
    if adj2.nnz:
        adj2.data[0] = 0.0
        want2 = _reference_dense_adj(adj2).nonzero().t()
        rowptr2, col2 = ingest.pattern_from_scipy(adj2)
        assert torch.equal(_edges_of(rowptr2, col2), want2)


def test_directed_and_loop_free_variants():
    row, col = torch.tensor([0, 2, 2]), torch.tensor([1, 0, 0])
    rp, c = ingest.pattern_from_edges(row, col, 3, symmetric=False, self_loops=False)
    assert rp.tolist() == [0, 1, 1, 2] and c.tolist() == [1, 0]
    rp, c = ingest.pattern_from_edges(row, col, 3, symmetric=True, self_loops=False)
    assert _edges_of(rp, c).t().tolist() == [[0, 1], [0, 2], [1, 0], [2, 0]]
    with pytest.raises(ValueError):
        ingest.pattern_from_edges(torch.tensor([0]), torch.tensor([3]), 3)
    with pytest.raises(ValueError):
        ingest.pattern_from_edges(torch.tensor([0, 1]), torch.tensor([1]), 3)


def test_batch_patterns_equal_block_diag():
    parts, dense = [], []
    for k, (n, m) in enumerate([(5, 6), (1, 0), (12, 30)]):
        e = _random_edges(n, m, 10 + k)
        parts.append(ingest.pattern_from_edges(torch.from_numpy(e[:, 0]), torch.from_numpy(e[:, 1]), n))
        adj = sp.coo_matrix((np.ones(e.shape[0]), (e[:, 0], e[:, 1])), shape=(n, n), dtype=np.float32)
        dense.append(_reference_dense_adj(adj))
    rowptr, col = ingest.batch_patterns(parts)
    want = torch.block_diag(*dense).nonzero().t()  # load_data_ppi.py:86
    assert torch.equal(_edges_of(rowptr, col), want)
    with pytest.raises(ValueError):
        ingest.batch_patterns([])


@pytest.mark.gpu
def test_graph_handles_equal_dense_build_and_run_the_model():
    import models
    from pygat_b200.graph import Graph
    dev = torch.device("cuda")
    n = 400
    e = _random_edges(n, 3000, 5)
    adj = _reference_dense_adj(sp.coo_matrix((np.ones(e.shape[0]), (e[:, 0], e[:, 1])), shape=(n, n), dtype=np.float32))
    g_dense = Graph.from_dense(adj.to(dev))
    g = ingest.graph_from_edges(torch.from_numpy(e[:, 0]).to(dev), torch.from_numpy(e[:, 1]).to(dev), n)
    assert torch.equal(g.rowptr, g_dense.rowptr) and torch.equal(g.col, g_dense.col)  # bit-exact
    halves = [ingest.graph_from_edges(torch.from_numpy(e[:1500, 0] % 200).to(dev), torch.from_numpy(e[:1500, 1] % 200).to(dev), 200),
              ingest.graph_from_edges(torch.from_numpy(e[1500:, 0] % 150).to(dev), torch.from_numpy(e[1500:, 1] % 150).to(dev), 150)]
    gb = ingest.batch_graphs(halves)
    blocks = torch.block_diag(*[torch.zeros(h.n_dst, h.n_dst).index_put_(tuple(h.edge_index().cpu()), torch.tensor(1.0))
                                for h in halves])
    gd = Graph.from_dense(blocks.to(dev))
    assert torch.equal(gb.rowptr, gd.rowptr) and torch.equal(gb.col, gd.col)
    # the model takes the handle wherever the reference takes the dense adj (models.py:29)
    import layers
    torch.manual_seed(0)
    model = models.GAT(nfeat=[20, 8, 5], nheads=[4, 1], nlayers=2, dropout=0.0, alpha=0.2,
                       layer_type=layers.SpGraphAttentionLayer).to(dev)
    x = torch.randn(n, 20, device=dev)
    y_dense = model(x, adj.to(dev))
    y_graph = model(x, g)
    assert torch.equal(y_dense, y_graph)


def test_hub_partition_host_build_matches_the_definition():
    """HubPartition (rows longer than seg_len cut into segments; edge-balanced work items) is built on the host from one
    copy of the row pointers; check it against the plain definition on skewed and degenerate patterns."""
    import torch
    from pygat_b200 import graph as G
    from pygat_b200.synth import power_law_csr
    cases = [power_law_csr(3000, 14.0, seed=3, exponent=0.7)[0], torch.zeros(8, dtype=torch.int64),
             torch.tensor([0, 0, 5, 5, 5, 700, 700, 1300], dtype=torch.int64), torch.arange(0, 4001, 4, dtype=torch.int64)]
    for ptr in cases:
        for seg_len in (16, 64, 512):
            hp = G.HubPartition(ptr, seg_len)
            deg = ptr[1:] - ptr[:-1]
            rows = torch.nonzero(deg > seg_len).flatten()
            assert hp.n_hub == rows.numel()
            if hp.n_hub:
                assert torch.equal(hp.rows.long(), rows)
                nseg = (deg[rows] + seg_len - 1) // seg_len
                assert torch.equal(hp.seg_ptr.long(), torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(nseg, 0)]))
                assert hp.n_seg == int(nseg.sum())
            n, e = ptr.numel() - 1, int(ptr[-1])
            if e == 0:
                assert hp.items is None and hp.n_items == 0
            else:
                it = hp.items.long()
                assert it[0] == 0 and it[-1] == n and torch.all(it[1:] > it[:-1]) and hp.n_items == it.numel() - 1
                # every item starts at the first row whose first entry is >= a multiple of ITEM_EDGES
                assert 32 <= hp.item_edges <= G.ITEM_EDGES
                starts = torch.searchsorted(ptr[:-1].contiguous(), torch.arange(0, e, hp.item_edges))
                assert set(it[:-1].tolist()) == (set(starts.tolist()) - {n}) | {0}
            assert hp.n_empty == int((deg == 0).sum())
