"""Worker for the multi-process tests (launched by torchrun / mp.spawn from tests/test_sharded_*.py).

mode "gloo":  host-side sharding logic on CPU tensors (row gather / reduce helpers, plan, weighted
              gradient all-reduce), world_size 2, backend gloo.
mode "nccl":  the sharded GAT layer on 2+ GPUs against the single-GPU layer (same inputs on every rank).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run_gloo(rank, world):
    from pygat_b200.sharded import ShardPlan, allreduce_, allreduce_gradients, gather_rows, reduce_rows
    from pygat_b200.synth import power_law_csr, shard_rows_by_nnz
    rowptr, col = power_law_csr(500, 8.0, seed=1)
    bounds = shard_rows_by_nnz(rowptr, world)
    assert bounds[0] == 0 and bounds[-1] == 500 and all(b1 >= b0 for b0, b1 in zip(bounds, bounds[1:]))
    nnz = [int(rowptr[b1] - rowptr[b0]) for b0, b1 in zip(bounds, bounds[1:])]
    assert max(nnz) - min(nnz) <= int((rowptr[1:] - rowptr[:-1]).max())  # balanced up to one row
    plan = ShardPlan(bounds, rank)
    # gather: every rank fills its slice, afterwards all ranks hold the same full matrix
    ref = torch.arange(500 * 3, dtype=torch.float32).view(500, 3)
    full = torch.full((500, 3), -1.0)
    plan.rows(full).copy_(plan.rows(ref))
    gather_rows(full, plan)
    assert torch.equal(full, ref)
    # reduce: rank r ends up with the sum over ranks of its row range
    part = ref * (rank + 1)
    got = reduce_rows(part.clone(), plan)
    assert torch.equal(got, plan.rows(ref) * sum(r + 1 for r in range(world)))
    a, b = torch.ones(3) * (rank + 1), torch.ones(2, 2) * rank
    allreduce_([a, b], plan)
    assert torch.equal(a, torch.ones(3) * sum(r + 1 for r in range(world)))
    assert torch.equal(b, torch.ones(2, 2) * sum(range(world)))
    # node-count weighted gradient all-reduce == gradient of the mean loss over the merged batch
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    sizes = [5, 11][:world] if world == 2 else [3 + r for r in range(world)]
    xs = [torch.randn(s, 4, generator=torch.Generator().manual_seed(10 + r)) for r, s in enumerate(sizes)]
    ys = [torch.rand(s, 3, generator=torch.Generator().manual_seed(20 + r)).round() for r, s in enumerate(sizes)]
    loss_fn = torch.nn.BCEWithLogitsLoss(reduction="mean")
    loss_fn(lin(torch.cat(xs)), torch.cat(ys)).backward()
    merged = [p.grad.clone() for p in lin.parameters()]
    lin.zero_grad()
    loss_fn(lin(xs[rank]), ys[rank]).backward()
    allreduce_gradients(lin.parameters(), sizes[rank])
    for p, m in zip(lin.parameters(), merged):
        assert torch.allclose(p.grad, m, atol=1e-6), (p.grad, m)
    # uneven epochs: 3 batches on 2 ranks = 2 steps, rank 1 idles in the second one but still all-reduces;
    # every step's gradient equals the merged gradient of the batches that ran in it
    from pygat_b200.sharded import rank_batch_schedule
    assert rank_batch_schedule(3, 0, 2) == [0, 2] and rank_batch_schedule(3, 1, 2) == [1, None]
    assert rank_batch_schedule(10, 3, 4) == [3, 7, None] and rank_batch_schedule(10, 1, 4) == [1, 5, 9]
    bx = [torch.randn(s, 4, generator=torch.Generator().manual_seed(30 + i)) for i, s in enumerate([6, 9, 4])]
    by = [torch.rand(s, 3, generator=torch.Generator().manual_seed(40 + i)).round() for i, s in enumerate([6, 9, 4])]
    for step, idx in enumerate(rank_batch_schedule(3, rank, world) if world == 2 else []):
        ran = [i for i in (step * world + r for r in range(world)) if i < 3]
        lin.zero_grad()
        loss_fn(lin(torch.cat([bx[i] for i in ran])), torch.cat([by[i] for i in ran])).backward()
        want = [p.grad.clone() for p in lin.parameters()]
        lin.zero_grad(set_to_none=True)
        n_nodes = 0
        if idx is not None:
            loss_fn(lin(bx[idx]), by[idx]).backward()
            n_nodes = bx[idx].shape[0]
        allreduce_gradients(lin.parameters(), n_nodes)
        for p, m in zip(lin.parameters(), want):
            assert torch.allclose(p.grad, m, atol=1e-6), (step, p.grad, m)
    # kept gather buffer: same key -> hit, new key -> same storage, miss
    from pygat_b200.sharded import _input_key
    xin = torch.zeros(7, 3)
    buf, hit = plan.cached_xg(_input_key(xin), (500, 4), torch.device("cpu"))
    assert not hit
    buf2, hit2 = plan.cached_xg(_input_key(xin), (500, 4), torch.device("cpu"))
    assert hit2 and buf2.data_ptr() == buf.data_ptr()
    xin.add_(1.0)  # in-place update: version counter moves, the kept rows are stale
    buf3, hit3 = plan.cached_xg(_input_key(xin), (500, 4), torch.device("cpu"))
    assert not hit3 and buf3.data_ptr() == buf.data_ptr()
    k_dead = _input_key(torch.zeros(7, 3))  # the tensor object is gone: never a hit, whatever reuses its storage
    plan.cached_xg(k_dead, (500, 4), torch.device("cpu"))
    assert not plan.cached_xg(k_dead, (500, 4), torch.device("cpu"))[1]
    from pygat_b200.sharded import reduce_rows_async
    got2, work = reduce_rows_async((ref * (rank + 1)).clone(), plan)
    if work is not None:
        work.wait()
    assert torch.equal(got2, plan.rows(ref) * sum(r + 1 for r in range(world)))
    # local graph slices reassemble the global CSR
    g_rows = [int(rowptr[b1] - rowptr[b0]) for b0, b1 in zip(bounds, bounds[1:])]
    assert sum(g_rows) == col.numel()


def run_nccl(rank, world):
    from pygat_b200.functional import gat_layer
    from pygat_b200.graph import Graph
    from pygat_b200.sharded import ShardPlan, sharded_gat_layer
    from pygat_b200.synth import init_layer_params, power_law_csr
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    for (n, H, D, f_in, skip, concat) in [(5000, 8, 64, 100, False, True), (3000, 4, 32, 36, True, False)]:
        rowptr, col = power_law_csr(n, 14.0, seed=5, exponent=0.7, device=dev)
        gen = torch.Generator(device=dev).manual_seed(3)
        x = torch.randn(n, f_in, generator=gen, device=dev)
        gout = torch.randn(n, H * D, generator=gen, device=dev)
        Ws, a_s, a_d = init_layer_params(f_in, H, D, dev, seed=9)
        Ss = [w.detach().clone().flip(0).requires_grad_(True) for w in Ws] if skip else None
        # single GPU reference on every rank
        xr = x.clone().requires_grad_(True)
        y_ref = gat_layer(xr, Graph.from_csr(rowptr, col, seg_len=256), Ws, a_s, a_d, Ss, 0.2, concat)
        y_ref.backward(gout)
        ref = {"dx": xr.grad.clone(), "dW": [w.grad.clone() for w in Ws], "da": [a.grad.clone() for a in a_s + a_d],
               "dS": [s.grad.clone() for s in Ss] if skip else []}
        for p in Ws + a_s + a_d + (Ss or []):
            p.grad = None
        plan = ShardPlan.by_nnz(rowptr, rank, world)
        graph = plan.local_graph(rowptr, col, seg_len=256)

        def rel(a, b):
            return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
        # hidden layer (input needs a gradient): [Wh | g] exchanged forward; backward either by reduce-scattering the
        # partial dWh rows (gt None) or on this rank's SOURCE rows from all-gathered destination records
        for gt in (None, plan.source_shard(rowptr, col, seg_len=256)):
            for p in Ws + a_s + a_d + (Ss or []):
                p.grad = None
            xl = plan.rows(x).clone().requires_grad_(True)
            y = sharded_gat_layer(xl, graph, plan, Ws, a_s, a_d, Ss, 0.2, concat, graph_t=gt)
            y.backward(plan.rows(gout))
            assert rel(y, plan.rows(y_ref)) < 1e-5, rel(y, plan.rows(y_ref))
            assert rel(xl.grad, plan.rows(ref["dx"])) < 1e-5
            for got, want in zip([w.grad for w in Ws], ref["dW"]):
                assert rel(got, want) < 1e-5, rel(got, want)
            for got, want in zip([a.grad for a in a_s + a_d], ref["da"]):
                assert rel(got, want) < 1e-5, rel(got, want)
            if skip:
                for got, want in zip([s.grad for s in Ss], ref["dS"]):
                    assert rel(got, want) < 1e-5
        # first-layer case: the input needs no gradient, so no dWh rows are exchanged at all
        for p in Ws + a_s + a_d + (Ss or []):
            p.grad = None
        # (narrow inputs take the aggregate-first form here while the single-GPU reference above ran the folded
        # form: two fp32 re-associations, each held to 1e-5 against the oracle in test_gpu_parity.py, so they
        # may differ from each other by up to the sum)
        y2 = sharded_gat_layer(plan.rows(x).clone(), graph, plan, Ws, a_s, a_d, Ss, 0.2, concat)
        y2.backward(plan.rows(gout))
        assert rel(y2, plan.rows(y_ref)) < 1e-5
        for got, want in zip([w.grad for w in Ws], ref["dW"]):
            assert rel(got, want) < 2e-5, rel(got, want)
        for got, want in zip([a.grad for a in a_s + a_d], ref["da"]):
            assert rel(got, want) < 2e-5, rel(got, want)
        if skip:
            for got, want in zip([s.grad for s in Ss], ref["dS"]):
                assert rel(got, want) < 2e-5
        # cached input gather: the same input tensor again hits the kept all-gathered rows and exchanges only g.
        # The first call runs with perturbed attention vectors, so the g columns it leaves behind are stale and the
        # second call must refresh them on every rank; an in-place update of the input must gather again.
        xs = plan.rows(x).clone()
        a_bad = [a.detach() * 1.7 for a in a_d]
        sharded_gat_layer(xs, graph, plan, Ws, a_s, a_bad, Ss, 0.2, concat)
        kept = getattr(plan, "_peer_rows", None)  # symmetric-memory exchange, or the NCCL path's kept buffer
        assert (kept.key if kept else plan._xg_cache[0])[:3] == (xs.data_ptr(), xs._version, tuple(xs.shape))
        for variant in ("hit", "invalidated"):
            for p in Ws + a_s + a_d + (Ss or []):
                p.grad = None
            if variant == "invalidated":
                xs.mul_(1.0)
            y2c = sharded_gat_layer(xs, graph, plan, Ws, a_s, a_d, Ss, 0.2, concat)
            y2c.backward(plan.rows(gout))
            assert rel(y2c, plan.rows(y_ref)) < 1e-5, (variant, rel(y2c, plan.rows(y_ref)))
            for got, want in zip([w.grad for w in Ws], ref["dW"]):
                assert rel(got, want) < 2e-5, (variant, rel(got, want))
            for got, want in zip([a.grad for a in a_s + a_d], ref["da"]):
                assert rel(got, want) < 2e-5, (variant, rel(got, want))
        # a second forward on the same plan rewrites the kept rows: the first forward's backward must refuse loudly
        # (raised before any collective, on every rank alike) instead of returning gradients of the wrong rows
        ya = sharded_gat_layer(xs, graph, plan, Ws, a_s, a_d, Ss, 0.2, concat)
        yb = sharded_gat_layer(xs, graph, plan, Ws, a_s, a_bad, Ss, 0.2, concat)
        try:
            ya.backward(plan.rows(gout))
            raise AssertionError("stale gathered rows were not detected")
        except RuntimeError as exc:
            assert "overwritten by a later forward" in str(exc) or "modified by an inplace operation" in str(exc), str(exc)
        yb.backward(plan.rows(gout))
        # and the project-first sharded kernels on the same no-gradient input
        for p in Ws + a_s + a_d + (Ss or []):
            p.grad = None
        y3 = sharded_gat_layer(plan.rows(x).clone(), graph, plan, Ws, a_s, a_d, Ss, 0.2, concat, form="project_first")
        y3.backward(plan.rows(gout))
        assert rel(y3, plan.rows(y_ref)) < 1e-5
        for got, want in zip([w.grad for w in Ws], ref["dW"]):
            assert rel(got, want) < 1e-5, rel(got, want)
        for got, want in zip([a.grad for a in a_s + a_d], ref["da"]):
            assert rel(got, want) < 1e-5, rel(got, want)
    # ---- whole model: models.GAT through sharded_gat_forward against the drop-in module on one GPU
    import copy

    import layers
    import models
    from pygat_b200.sharded import sharded_gat_forward
    n = 5000
    rowptr, col = power_law_csr(n, 14.0, seed=5, exponent=0.7, device=dev)
    gen = torch.Generator(device=dev).manual_seed(4)
    x = torch.randn(n, 36, generator=gen, device=dev)
    gout = torch.randn(n, 7, generator=gen, device=dev)
    for skip in (False, True):
        torch.manual_seed(11)
        m_ref = models.GAT(nfeat=[36, 16, 7], nheads=[4, 2], nlayers=2, dropout=0.0, alpha=0.2,
                           layer_type=layers.SpGraphAttentionLayer, skip_connection=skip).to(dev).train()
        m_sh = copy.deepcopy(m_ref)
        y_ref = m_ref(x, Graph.from_csr(rowptr, col, seg_len=256))
        y_ref.backward(gout)
        plan = ShardPlan.by_nnz(rowptr, rank, world)
        graph = plan.local_graph(rowptr, col, seg_len=256)
        y = sharded_gat_forward(m_sh, plan.rows(x).clone(), graph, plan, plan.source_shard(rowptr, col, seg_len=256))
        y.backward(plan.rows(gout))

        def rel(a, b):
            return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
        assert rel(y, plan.rows(y_ref)) < 1e-5, rel(y, plan.rows(y_ref))
        for (k, p_ref), p_sh in zip(m_ref.named_parameters(), m_sh.parameters()):
            assert rel(p_sh.grad, p_ref.grad) < 2e-5, (k, rel(p_sh.grad, p_ref.grad))
    torch.cuda.synchronize()


def main():
    import faulthandler
    # a rank that fails while its peer waits in a collective would otherwise hang the job: dump every thread's
    # stack and exit instead
    faulthandler.dump_traceback_later(int(os.environ.get("DIST_WORKER_TIMEOUT", "120")), exit=True)
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo" if mode == "gloo" else "nccl")
    try:
        (run_gloo if mode == "gloo" else run_nccl)(rank, world)
        dist.barrier()
        if rank == 0:
            print(f"DIST_OK mode={mode} world={world}")
    except BaseException:
        # report and leave at once: tearing the process group down would wait for the peer's pending collective
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
