"""Multi-process tests: host-side sharding logic on CPU (gloo, world_size 2) and, on a box with two or
more GPUs, the sharded GAT layer against the single-GPU layer (NCCL)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(mode, nproc, timeout=600, **extra_env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "dist_worker.py"), mode]
    env = dict(os.environ, OMP_NUM_THREADS="2", **extra_env)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0 and f"DIST_OK mode={mode} world={nproc}" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_sharding_host_logic_world2_gloo():
    _torchrun("gloo", 2)


def test_shard_plan_local_graph_shapes():
    from pygat_b200.synth import power_law_csr, shard_rows_by_nnz
    rowptr, col = power_law_csr(300, 6.0, seed=2)
    b = shard_rows_by_nnz(rowptr, 4)
    assert len(b) == 5 and b[0] == 0 and b[-1] == 300
    assert shard_rows_by_nnz(rowptr, 1) == [0, 300]
    # cost-balanced cuts: entries + row_cost per row; a large row cost tends to equal row counts
    c = shard_rows_by_nnz(rowptr, 4, row_cost=10_000)
    assert c[0] == 0 and c[-1] == 300 and max(c[i + 1] - c[i] for i in range(4)) - min(c[i + 1] - c[i] for i in range(4)) <= 2
    cost = lambda lo, hi, rc: int(rowptr[hi] - rowptr[lo]) + rc * (hi - lo)
    d = shard_rows_by_nnz(rowptr, 4, row_cost=25)
    costs = [cost(d[i], d[i + 1], 25) for i in range(4)]
    assert max(costs) - min(costs) <= int((rowptr[1:] - rowptr[:-1]).max()) + 25


@pytest.mark.gpu
def test_sharded_layer_world1_equals_single_gpu():
    from pygat_b200.functional import gat_layer
    from pygat_b200.graph import Graph
    from pygat_b200.sharded import ShardPlan, sharded_gat_layer
    from pygat_b200.synth import init_layer_params, power_law_csr
    dev = "cuda"
    n, H, D, f_in = 4000, 8, 64, 100
    rowptr, col = power_law_csr(n, 12.0, seed=5, exponent=0.7, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, f_in, generator=g, device=dev)
    gout = torch.randn(n, H * D, generator=g, device=dev)
    Ws, a_s, a_d = init_layer_params(f_in, H, D, dev)
    y0 = gat_layer(x, Graph.from_csr(rowptr, col, seg_len=128), Ws, a_s, a_d, None, 0.2, True)
    y0.backward(gout)
    g0 = [p.grad.clone() for p in Ws + a_s + a_d]
    for p in Ws + a_s + a_d:
        p.grad = None
    plan = ShardPlan([0, n], 0)
    y1 = sharded_gat_layer(x, plan.local_graph(rowptr, col, seg_len=128), plan, Ws, a_s, a_d, None, 0.2, True)
    y1.backward(gout)
    # the single-GPU layer runs the folded form, the sharded one the explicit kernels: same math, re-associated
    assert (y0 - y1).abs().max().item() <= 2e-6 * y0.abs().max().item()
    for a, b in zip(g0, [p.grad for p in Ws + a_s + a_d]):
        assert (a - b).abs().max().item() <= 3e-6 * b.abs().max().item()


@pytest.mark.gpu
@pytest.mark.parametrize("source_shard", [False, True])
@pytest.mark.parametrize("H,D,f_in,skip,concat", [(8, 64, 96, False, True), (6, 121, 64, True, False), (4, 32, 128, True, True),
                                                  (1, 16, 40, False, True), (3, 8, 20, False, False)])
def test_sharded_hidden_layer_world1_equals_single_gpu(H, D, f_in, skip, concat, source_shard):
    """The hidden-layer form (own-row projection, [Wh | g] exchanged in head chunks, ShardedGatLayerWhFunction) on a
    one-rank plan against functional.gat_layer: outputs, input gradient and every parameter gradient."""
    from pygat_b200.functional import gat_layer
    from pygat_b200.graph import Graph
    from pygat_b200.sharded import ShardPlan, sharded_gat_layer
    from pygat_b200.synth import init_layer_params, power_law_csr
    dev = "cuda"
    n = 3000
    rowptr, col = power_law_csr(n, 12.0, seed=5, exponent=0.7, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, f_in, generator=g, device=dev)
    gout = torch.randn(n, H * D if concat or True else D, generator=g, device=dev)
    Ws, a_s, a_d = init_layer_params(f_in, H, D, dev)
    Ss = [w.detach().clone().flip(0).requires_grad_(True) for w in Ws] if skip else None
    params = Ws + a_s + a_d + (Ss or [])
    x0 = x.clone().requires_grad_(True)
    y0 = gat_layer(x0, Graph.from_csr(rowptr, col, seg_len=128), Ws, a_s, a_d, Ss, 0.2, concat)
    y0.backward(gout)
    g0 = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    plan = ShardPlan([0, n], 0)
    x1 = x.clone().requires_grad_(True)
    # source_shard: the backward walks the rank's SOURCE rows (all-gathered destination records, df through reds)
    gt = plan.source_shard(rowptr, col, seg_len=128) if source_shard else None
    y1 = sharded_gat_layer(x1, plan.local_graph(rowptr, col, seg_len=128), plan, Ws, a_s, a_d, Ss, 0.2, concat, graph_t=gt)
    y1.backward(gout)
    rel = lambda a, b: (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    assert rel(y1, y0) < 3e-6
    assert rel(x1.grad, x0.grad) < 5e-6
    for a, b in zip([p.grad for p in params], g0):
        assert rel(a, b) < 5e-6


@pytest.mark.gpu
def test_second_forward_before_backward_is_refused():
    """The aggregate-first sharded layer keeps its gathered rows in a plan-wide buffer; a later forward rewrites it
    behind autograd's back, so the earlier forward's backward must raise rather than use the wrong rows."""
    from pygat_b200.sharded import ShardPlan, sharded_gat_layer
    from pygat_b200.synth import init_layer_params, power_law_csr
    dev = "cuda"
    n, H, D, f_in = 2000, 4, 32, 20
    rowptr, col = power_law_csr(n, 8.0, seed=2, exponent=0.7, device=dev)
    x = torch.randn(n, f_in, device=dev)
    gout = torch.randn(n, H * D, device=dev)
    Ws, a_s, a_d = init_layer_params(f_in, H, D, dev)
    plan = ShardPlan([0, n], 0)
    graph = plan.local_graph(rowptr, col, seg_len=128)
    ya = sharded_gat_layer(x, graph, plan, Ws, a_s, a_d, None, 0.2, True)
    yb = sharded_gat_layer(x, graph, plan, Ws, a_s, [a * 2 for a in a_d], None, 0.2, True)
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        ya.backward(gout)
    yb.backward(gout)
    # without the kept rows every call owns its buffer: both backwards are fine
    yc = sharded_gat_layer(x, graph, plan, Ws, a_s, a_d, None, 0.2, True, cache_input_gather=False)
    yd = sharded_gat_layer(x, graph, plan, Ws, a_s, a_d, None, 0.2, True, cache_input_gather=False)
    yc.backward(gout)
    yd.backward(gout)


@pytest.mark.gpu
def test_sharded_layer_matches_single_gpu_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _torchrun("nccl", 2, GATK_PEER_PUSH="0")  # NCCL exchanges only


@pytest.mark.gpu
def test_sharded_layer_peer_push_exchange_nccl():
    """Forward exchange fused into the pack kernel (stores into the peers' symmetric-memory copies over NVLink)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _torchrun("nccl", 2, GATK_PEER_PUSH="1")  # the default


def test_row_cost_fit_and_fractional_cuts():
    from pygat_b200.sharded import fit_row_cost, shard_rows_by_cost
    from pygat_b200.synth import power_law_csr, shard_rows_by_nnz
    rows = torch.tensor([100., 200., 300., 400.])
    ent = torch.tensor([4000., 3000., 2500., 2000.])
    assert abs(fit_row_cost(rows, ent, 7.0 * rows + 0.25 * ent) - 28.0) < 1e-6
    assert fit_row_cost(rows[:1], ent[:1], torch.tensor([1.0])) is None          # one rank: not identifiable
    assert fit_row_cost(rows, 10 * rows, 3.0 * rows) is None                     # collinear shard shapes
    assert fit_row_cost(rows, ent, 1.0 * rows - 0.5 * ent) is None               # negative coefficient: rejected
    rowptr, _ = power_law_csr(3000, 10.0, seed=1)
    assert shard_rows_by_cost(rowptr, 4, 25.0) == shard_rows_by_nnz(rowptr, 4, 25)
    b = shard_rows_by_cost(rowptr, 4, 12.5)
    assert b[0] == 0 and b[-1] == 3000 and all(x <= y for x, y in zip(b, b[1:]))
