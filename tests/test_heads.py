"""Loss heads + metrics of the callers' training steps: the oracle's restatement against sklearn
(CPU) and the fused kernels of csrc/heads.cu against the oracle (GPU)."""
import pytest
import torch

from oracle import gat_oracle as O


def _citation_case(n=700, C=7, seed=0):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, C, generator=g) * 2.0
    labels = torch.randint(0, C, (n,), generator=g)
    idx = torch.randperm(n, generator=g)[:140]
    return logits, labels, idx


def _ppi_case(n=900, C=121, seed=1):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, C, generator=g) * 3.0
    labels = (torch.rand(n, C, generator=g) < 0.3).float()
    return logits, labels


def test_oracle_micro_f1_is_sklearns():
    from sklearn.metrics import f1_score
    logits, labels = _ppi_case()
    _, f1 = O.ppi_loss_f1(logits, labels)
    want = f1_score(labels.numpy(), (logits > 0).float().numpy(), average="micro")  # train_ppi.py:106-110
    assert abs(f1.item() - want) < 1e-12
    z = torch.full((5, 3), -1.0)
    assert O.ppi_loss_f1(z, torch.zeros(5, 3))[1].item() == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("n,C", [(700, 7), (2708, 7), (19717, 3), (300, 121)])
def test_citation_head_matches_oracle(n, C):
    from pygat_b200.heads import citation_head
    logits, labels, idx = _citation_case(n, C, seed=n)
    lo = logits.double().requires_grad_(True)
    loss_o, acc_o = O.citation_loss_acc(lo, labels, idx)
    (loss_o * 1.7).backward()
    ld = logits.cuda().requires_grad_(True)
    loss, acc = citation_head(ld, labels.cuda(), idx.cuda())
    (loss * 1.7).backward()
    assert abs(loss.item() - loss_o.item()) < 1e-6 * max(1.0, abs(loss_o.item()))
    assert abs(acc.item() - acc_o.item()) < 1e-7
    assert (ld.grad.cpu().double() - lo.grad).abs().max().item() < 1e-6 * lo.grad.abs().max().item()
    # all rows (idx = None) and repeated rows
    loss2, _ = citation_head(ld.detach(), labels.cuda(), None)
    loss2_o, _ = O.citation_loss_acc(logits.double(), labels, torch.arange(n))
    assert abs(loss2.item() - loss2_o.item()) < 1e-6 * max(1.0, abs(loss2_o.item()))


@pytest.mark.gpu
@pytest.mark.parametrize("n,C", [(900, 121), (4500, 121), (64, 5)])
def test_ppi_head_matches_oracle(n, C):
    from pygat_b200.heads import ppi_head
    logits, labels = _ppi_case(n, C, seed=n)
    lo = logits.double().requires_grad_(True)
    loss_o, f1_o = O.ppi_loss_f1(lo, labels.double())
    loss_o.backward()
    ld = logits.cuda().requires_grad_(True)
    loss, f1 = ppi_head(ld, labels.cuda())
    loss.backward()
    assert abs(loss.item() - loss_o.item()) < 1e-6 * max(1.0, abs(loss_o.item()))
    assert abs(f1.item() - f1_o.item()) < 1e-6
    assert (ld.grad.cpu().double() - lo.grad).abs().max().item() < 1e-6 * lo.grad.abs().max().item()


@pytest.mark.gpu
def test_sync_free_epochs_equal_the_reference_steps():
    """heads.citation_epoch / ppi_batch_step against the same steps written as the reference's scripts write them
    (train.py:154-179, train_ppi.py:117-124) on the same model state: identical parameters after the step."""
    import copy

    import torch.nn.functional as F

    import layers
    import models
    from pygat_b200.heads import citation_epoch, ppi_batch_step
    from tests.golden_io import dense_adj, load
    d = load("gat_sp_pubmed_like")
    adj = dense_adj(d).cuda()
    x = d["x"].cuda()
    n = x.shape[0]
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, 3, (n,), generator=g).cuda()
    idx_train, idx_val = torch.arange(40).cuda(), torch.arange(60, 120).cuda()
    nheads = [int(v) for v in d["nheads"]]
    torch.manual_seed(1)
    m1 = models.GAT(nfeat=[int(v) for v in d["nfeat"]], nheads=nheads, nlayers=2, dropout=0.0, alpha=0.2,
                    layer_type=layers.SpGraphAttentionLayer, skip_connection=False).cuda()
    m2 = copy.deepcopy(m1)
    o1 = torch.optim.Adam(m1.parameters(), lr=0.01, weight_decay=0.001)
    o2 = torch.optim.Adam(m2.parameters(), lr=0.01, weight_decay=0.001)
    lt, at, lv, av = citation_epoch(m1, o1, x, adj, labels, idx_train, idx_val)
    m2.train()
    o2.zero_grad()
    out = F.log_softmax(F.elu(m2(x, adj)), dim=1)
    loss = F.nll_loss(out[idx_train], labels[idx_train])
    loss.backward()
    o2.step()
    m2.eval()
    out = F.log_softmax(F.elu(m2(x, adj)), dim=1)
    loss_val = F.nll_loss(out[idx_val], labels[idx_val])
    assert abs(lt.item() - loss.item()) < 1e-5 and abs(lv.item() - loss_val.item()) < 1e-5
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert (a - b).abs().max().item() < 1e-5
    # PPI step
    d = load("gat_de_ppi_like")
    adj = dense_adj(d).cuda()
    x = d["x"].cuda()
    yl = (torch.rand(x.shape[0], 11, generator=g) < 0.3).float().cuda()
    torch.manual_seed(2)
    p1 = models.GAT(nfeat=[10, 32, 32, 11], nheads=[4, 4, 6], nlayers=3, dropout=0.0, alpha=0.2,
                    layer_type=layers.GraphAttentionLayer, skip_connection=True).cuda().train()
    p2 = copy.deepcopy(p1)
    q1 = torch.optim.Adam(p1.parameters(), lr=0.005)
    q2 = torch.optim.Adam(p2.parameters(), lr=0.005)
    l1, f1 = ppi_batch_step(p1, q1, x, yl, adj)
    out = p2(x, adj)
    l2 = torch.nn.BCEWithLogitsLoss(reduction="mean")(out, yl)
    q2.zero_grad()
    l2.backward()
    q2.step()
    from sklearn.metrics import f1_score
    f2 = f1_score(yl.cpu().numpy(), (out > 0).float().cpu().numpy(), average="micro")
    assert abs(l1.item() - l2.item()) < 1e-6 and abs(f1.item() - f2) < 1e-6
    for a, b in zip(p1.parameters(), p2.parameters()):
        assert (a - b).abs().max().item() < 1e-5


@pytest.mark.gpu
def test_cuda_graph_replays_equal_eager_steps():
    """pygat_b200.graphed.GraphedStep: the whole PPI-like training step (3 layers with skip, fused BCE head, backward,
    Adam) captured once; three replays leave the parameters where three eager steps leave them."""
    import copy

    import layers
    import models
    from pygat_b200.graphed import GraphedStep
    from pygat_b200.heads import ppi_head
    from tests.golden_io import dense_adj, load
    d = load("gat_de_ppi_like")
    adj = dense_adj(d).cuda()
    x = d["x"].cuda()
    g = torch.Generator().manual_seed(0)
    yl = (torch.rand(x.shape[0], 11, generator=g) < 0.3).float().cuda()
    torch.manual_seed(2)
    m1 = models.GAT(nfeat=[10, 32, 32, 11], nheads=[4, 4, 6], nlayers=3, dropout=0.0, alpha=0.2,
                    layer_type=layers.GraphAttentionLayer, skip_connection=True).cuda().train()
    m2 = copy.deepcopy(m1)
    o1 = torch.optim.Adam(m1.parameters(), lr=0.005, capturable=True)
    o2 = torch.optim.Adam(m2.parameters(), lr=0.005)

    def step():
        o1.zero_grad(set_to_none=True)
        loss, f1 = ppi_head(m1(x, adj), yl)
        loss.backward()
        o1.step()
        return loss.detach(), f1

    gs = GraphedStep(m1, o1, step)
    for a, b in zip(m1.parameters(), m2.parameters()):   # the warm-up steps were undone
        assert torch.equal(a, b)
    losses = []
    for _ in range(3):
        loss, _ = gs()
        losses.append(loss.item())
        o2.zero_grad()
        l2 = torch.nn.BCEWithLogitsLoss()(m2(x, adj), yl)
        l2.backward()
        o2.step()
        assert abs(losses[-1] - l2.item()) < 1e-5
    assert losses[2] < losses[0]
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert (a - b).abs().max().item() < 1e-5
    with pytest.raises(RuntimeError, match="capture-safe"):
        GraphedStep(m2, o2, lambda: None)
