"""Epoch-time workloads of BASELINE.json's configs 2 and 3, driven exactly as the reference's
training loops drive the model (train.py:154-179, train_ppi.py:112-132) but on synthetic data of
the same shapes (the reference's feature blobs are missing and nothing of it travels to the GPU
box).  Used by bench.py; returns milliseconds per epoch."""
from __future__ import annotations

import torch
import torch.nn.functional as F

# node counts of the 20 PPI training graphs (data/ppi/train_graph_id.npy of the reference)
PPI_TRAIN_GRAPH_NODES = [1767, 1377, 2263, 2339, 1578, 1021, 1823, 2488, 591, 3312, 2401, 1878, 1819, 3480, 2794,
                         2326, 2650, 2815, 3163, 3021]


def _dense_from_csr(rowptr, col, n, device, column_major=False):
    row = torch.repeat_interleave(torch.arange(n, device=device), rowptr[1:] - rowptr[:-1])
    adj = torch.zeros(n, n, device=device)
    adj[row, col.long()] = 1.0
    if column_major:  # what utils.load_data hands the model (utils.py:55, normalize_adj returns CSC)
        adj = adj.t().contiguous().t()
    return adj


def pubmed_epoch_ms(device, epochs: int = 20, warmup: int = 3, seed: int = 72):
    """`python train.py --dataset pubmed --model GAT_sparse` shape: N=19717, ~108k stored entries,
    500 features, 8x8 hidden heads, 8x3 output heads averaged, dropout 0.6, Adam lr 0.01 wd 0.001;
    one epoch = train step + eval forward (train.py:154-170)."""
    import layers
    import models
    from pygat_b200.synth import power_law_csr
    n, f_in, classes = 19717, 500, 3
    torch.manual_seed(seed)
    rowptr, col = power_law_csr(n, 5.5, seed=seed, device=device)
    adj = _dense_from_csr(rowptr, col, n, device, column_major=True)
    x = torch.rand(n, f_in, device=device)
    x = x / x.sum(1, keepdim=True)
    labels = torch.randint(0, classes, (n,), device=device)
    idx_train = torch.arange(60, device=device)
    idx_val = torch.arange(200, 700, device=device)
    model = models.GAT(nfeat=[f_in, 8, classes], nheads=[8, 8], nlayers=2, dropout=0.6, alpha=0.2,
                       layer_type=layers.SpGraphAttentionLayer, skip_connection=False).to(device)
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.001)

    def epoch():
        model.train()
        opt.zero_grad()
        out = F.log_softmax(F.elu(model(x, adj)), dim=1)
        loss = F.nll_loss(out[idx_train], labels[idx_train])
        loss.backward()
        opt.step()
        model.eval()
        with torch.no_grad():
            out = F.log_softmax(F.elu(model(x, adj)), dim=1)
            lv = F.nll_loss(out[idx_val], labels[idx_val])
        return loss.item() + lv.item()  # the reference prints both every epoch (host sync)

    for _ in range(warmup):
        epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(epochs):
        epoch()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / epochs, {"nodes": n, "stored_entries": int(col.numel()), "epochs": epochs}


def _ppi_setup(device, seed):
    import layers
    import models
    from pygat_b200.synth import power_law_csr
    torch.manual_seed(seed)
    graphs = []
    for k, n in enumerate(PPI_TRAIN_GRAPH_NODES):
        rowptr, col = power_law_csr(n, 29.0, seed=seed + k, exponent=0.3, device=device)
        adj = _dense_from_csr(rowptr, col, n, device)
        g = torch.Generator(device=device).manual_seed(seed + 100 + k)
        feats = torch.randn(n, 50, generator=g, device=device)
        labels = (torch.rand(n, 121, generator=g, device=device) < 0.3).float()
        graphs.append((feats, labels, adj))
    batches = [(torch.cat([graphs[i][0], graphs[i + 1][0]]), torch.cat([graphs[i][1], graphs[i + 1][1]]),
                torch.block_diag(graphs[i][2], graphs[i + 1][2])) for i in range(0, len(graphs), 2)]
    model = models.GAT(nfeat=[50, 256, 256, 121], nheads=[4, 4, 6], nlayers=3, dropout=0.0, alpha=0.2,
                       layer_type=layers.GraphAttentionLayer, skip_connection=True).to(device)
    return graphs, batches, model


def _time_epochs(epoch, epochs, warmup):
    for _ in range(warmup):
        epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(epochs):
        epoch()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / epochs


def ppi_epoch_ms(device, rank: int = 0, world: int = 1, epochs: int = 3, warmup: int = 1, seed: int = 72):
    """`python train_ppi.py` shape: 20 training graphs, batches of 2 merged by block_diag
    (load_data_ppi.py:84-86), 3 layers 50 -> 4x256 -> 4x256 -> 6x121 (mean), skip connections, dense
    class (the script's default), BCEWithLogits, Adam lr 0.005; driven as train_ppi.py:112-132 drives it, including
    its per-batch host read of the loss.  With world > 1 the batches of an epoch are dealt round-robin to the ranks
    (graph-level data parallelism) and gradients are all-reduced with node-count weights."""
    from pygat_b200.sharded import allreduce_gradients, rank_batch_schedule
    graphs, batches, model = _ppi_setup(device, seed)
    opt = torch.optim.Adam(model.parameters(), lr=0.005, weight_decay=0.0)
    loss_fn = torch.nn.BCEWithLogitsLoss(reduction="mean")
    mine = [batches[i] if i is not None else None for i in rank_batch_schedule(len(batches), rank, world)]

    def epoch():
        model.train()
        tot = 0.0
        for item in mine:  # every rank runs the same number of steps; None = this rank idles in the step
            opt.zero_grad(set_to_none=True)
            n_nodes = 0
            if item is not None:
                feats, labels, adj = item
                loss = loss_fn(model(feats, adj), labels)
                loss.backward()
                n_nodes = feats.shape[0]
                tot += loss.item()  # the reference prints the loss of every batch (host sync)
            if world > 1:
                allreduce_gradients(model.parameters(), n_nodes)
            opt.step()
        return tot

    ms = _time_epochs(epoch, epochs, warmup)
    return ms, {"graphs": len(graphs), "batches_per_rank": sum(m is not None for m in mine), "epochs": epochs,
                "nodes": sum(PPI_TRAIN_GRAPH_NODES)}


def ppi_epoch_sync_free_ms(device, rank: int = 0, world: int = 1, epochs: int = 3, warmup: int = 1, seed: int = 72):
    """The PPI epoch with the caller's side widened but WITHOUT graph capture (works at any rank count): fused BCE +
    on-device micro-F1 (heads.ppi_batch_step), the sync-free single-collective gradient all-reduce of
    sharded.allreduce_gradients, one host read per epoch."""
    from pygat_b200.heads import ppi_batch_step
    from pygat_b200.sharded import allreduce_gradients, rank_batch_schedule
    graphs, batches, model = _ppi_setup(device, seed)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=0.005, weight_decay=0.0)
    mine = [batches[i] if i is not None else None for i in rank_batch_schedule(len(batches), rank, world)]
    tot = torch.zeros((), device=device)

    def epoch():
        tot.zero_()
        for item in mine:
            if item is not None:
                feats, labels, adj = item
                loss, _f1 = ppi_batch_step(model, opt, feats, labels, adj,
                                           allreduce=(lambda ps, n: allreduce_gradients(ps, n)) if world > 1 else None)
                tot.add_(loss)
            else:   # this rank idles in the step but still takes part in the collective and the (zero) update
                opt.zero_grad(set_to_none=True)
                allreduce_gradients(model.parameters(), 0)
                opt.step()
        return tot.item()

    ms = _time_epochs(epoch, epochs, warmup)
    return ms, {"graphs": len(graphs), "batches_per_rank": sum(m is not None for m in mine), "epochs": epochs,
                "nodes": sum(PPI_TRAIN_GRAPH_NODES), "host_reads": "1 per epoch"}


def ppi_epoch_graphed_ms(device, epochs: int = 5, warmup: int = 1, seed: int = 72):
    """The same epoch with the caller's side widened: fused BCE + on-device micro-F1
    (pygat_b200.heads.ppi_head) instead of the per-batch .cpu().numpy() + sklearn, no per-batch .item(), and every
    batch's whole step (forward, loss, backward, Adam) replayed from a CUDA graph (pygat_b200.graphed).  One host
    read per EPOCH (the summed loss).  Single GPU."""
    from pygat_b200.graphed import GraphedStep
    from pygat_b200.heads import ppi_head
    graphs, batches, model = _ppi_setup(device, seed)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=0.005, weight_decay=0.0, capturable=True)
    steps, pool = [], None

    def make(feats, labels, adj):
        def step():
            opt.zero_grad(set_to_none=True)
            loss, f1 = ppi_head(model(feats, adj), labels)
            loss.backward()
            opt.step()
            return loss.detach(), f1
        return step

    for feats, labels, adj in batches:
        gs = GraphedStep(model, opt, make(feats, labels, adj), pool=pool)
        pool = pool or gs.pool()
        steps.append(gs)
    tot = torch.zeros((), device=device)

    def epoch():
        tot.zero_()
        for gs in steps:
            loss, _f1 = gs()
            tot.add_(loss)
        return tot.item()  # one host read per epoch

    ms = _time_epochs(epoch, epochs, warmup)
    return ms, {"graphs": len(graphs), "batches": len(batches), "epochs": epochs, "nodes": sum(PPI_TRAIN_GRAPH_NODES),
                "cuda_graphs": len(steps)}


def pubmed_epoch_sync_free_ms(device, epochs: int = 20, warmup: int = 3, seed: int = 72):
    """pubmed_epoch_ms with the caller's side widened: the fused loss head (heads.citation_epoch) and one host read
    per 10 epochs instead of four .item() calls per epoch.  Dropout 0.6 is active, so the step is not graph-captured."""
    import layers
    import models
    from pygat_b200.heads import citation_epoch
    from pygat_b200.synth import power_law_csr
    n, f_in, classes = 19717, 500, 3
    torch.manual_seed(seed)
    rowptr, col = power_law_csr(n, 5.5, seed=seed, device=device)
    adj = _dense_from_csr(rowptr, col, n, device, column_major=True)
    x = torch.rand(n, f_in, device=device)
    x = x / x.sum(1, keepdim=True)
    labels = torch.randint(0, classes, (n,), device=device)
    idx_train = torch.arange(60, device=device)
    idx_val = torch.arange(200, 700, device=device)
    model = models.GAT(nfeat=[f_in, 8, classes], nheads=[8, 8], nlayers=2, dropout=0.6, alpha=0.2,
                       layer_type=layers.SpGraphAttentionLayer, skip_connection=False).to(device)
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.001)
    k = [0]

    def epoch():
        lt, at, lv, av = citation_epoch(model, opt, x, adj, labels, idx_train, idx_val)
        k[0] += 1
        if k[0] % 10 == 0:
            return lv.item()
        return None

    ms = _time_epochs(epoch, epochs, warmup)
    return ms, {"nodes": n, "stored_entries": int(col.numel()), "epochs": epochs, "host_reads": "1 per 10 epochs"}


def products_model_epoch_ms(device, rank: int = 0, world: int = 1, epochs: int = 5, warmup: int = 2, seed: int = 72):
    """BASELINE.json metric (ii) at the products shape: one full-batch training epoch of a 2-layer models.GAT
    (100 -> 8 x 64 -> 47 classes, ogbn-products' widths; SpGraphAttentionLayer, p = 0) = forward of both layers,
    the fused loss head of train.py:151-160 on an 8 % training split, backward, Adam -- on one GPU through the
    drop-in modules, on N GPUs through sharded.sharded_gat_forward (layer 1 aggregate-first with the kept input
    rows, layer 2 hidden-layer form with the source-shard backward; every rank's loss term is weighted by its share of
    the training nodes, the layers' backward all-reduces the parameter gradients)."""
    import torch.distributed as dist

    import layers
    import models
    from pygat_b200.graph import Graph
    from pygat_b200.heads import citation_head
    from pygat_b200.sharded import ShardPlan, shard_rows_by_cost, sharded_gat_forward
    from pygat_b200.synth import power_law_csr
    n, f_in, classes = 2_449_029, 100, 47
    rowptr, col = power_law_csr(n, 25.26, seed=seed, exponent=0.5, device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(n, f_in, generator=g, device=device)
    labels = torch.randint(0, classes, (n,), generator=g, device=device)
    is_train = torch.rand(n, generator=g, device=device) < 0.08
    torch.manual_seed(seed)
    model = models.GAT(nfeat=[f_in, 64, classes], nheads=[8, 1], nlayers=2, dropout=0.0, alpha=0.2,
                       layer_type=layers.SpGraphAttentionLayer, skip_connection=False).to(device).train()
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.0)
    n_train = int(is_train.sum().item())
    if world > 1:
        plan = ShardPlan(shard_rows_by_cost(rowptr, world, 30.0), rank)
        graph = plan.local_graph(rowptr, col)
        graph_t = plan.source_shard(rowptr, col)
        x_loc = plan.rows(x).clone()
        lab_loc = plan.rows(labels).clone()
        idx = torch.nonzero(plan.rows(is_train)).flatten()
        del x, labels
        weight = idx.numel() / max(n_train, 1)

        def epoch():
            opt.zero_grad(set_to_none=True)
            out = sharded_gat_forward(model, x_loc, graph, plan, graph_t)
            loss, _acc = citation_head(out, lab_loc, idx)
            (loss * weight).backward()
            opt.step()
            return None
    else:
        graph = Graph.from_csr(rowptr, col)
        graph.transpose()
        idx = torch.nonzero(is_train).flatten()

        def epoch():
            opt.zero_grad(set_to_none=True)
            out = model(x, graph)
            loss, _acc = citation_head(out, labels, idx)
            loss.backward()
            opt.step()
            return None
    del rowptr, col
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    ms = _time_epochs(epoch, epochs, warmup)
    return ms, {"nodes": n, "stored_entries": 61_861_615, "layers": "100 -> 8x64 -> 47 (mean of 1 head)", "train_nodes": n_train,
                "epochs": epochs, "n_gpus": world}
