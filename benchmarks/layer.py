"""Layer workloads of bench.py: one GAT layer (all heads) forward + backward over a synthetic power-law graph, on one
GPU through `functional.gat_layer` and on destination-row shards through `sharded.sharded_gat_layer`, the
host-buffer end-to-end loop, and the parity check of a sharded step against a single-GPU run of the same step."""
from __future__ import annotations

import torch
import torch.distributed as dist

from pygat_b200 import _lib
from pygat_b200.functional import gat_layer
from pygat_b200.graph import Graph
from pygat_b200.sharded import ShardPlan, fit_row_cost, shard_rows_by_cost, sharded_gat_layer
from pygat_b200.synth import init_layer_params, power_law_csr

XBWD_SHORT_ROW = 32  # rows up to two 16-entry chunks: the aggregate-first backward does not read their xagg_i (csrc/attn_x.cu)


def long_rows(rowptr: torch.Tensor) -> int:
    return int(((rowptr[1:] - rowptr[:-1]) > XBWD_SHORT_ROW).sum().item())


ROW_COST = 25  # measured on one GPU at the products shape: row-proportional kernels ~7.0 ns/row, edge passes ~0.275 ns/entry


def _inputs(cfg, dev):
    """Seeded inputs of a workload, identical on every rank and in the single-GPU run: (rowptr, col, x, gout)."""
    n, H, D, f_in = cfg["n"], cfg["H"], cfg["D"], cfg["f_in"]
    rowptr, col = power_law_csr(n, cfg["avg_deg"], seed=72, exponent=cfg["exponent"], device=dev)
    g = torch.Generator(device=dev).manual_seed(72)
    x = torch.randn(n, f_in, generator=g, device=dev)
    gout = torch.randn(n, H * D, generator=g, device=dev)
    return rowptr, col, x, gout


class SingleGpuLayerBench:
    """Whole graph on one GPU; the layer is driven through the public functional API.  cfg["needs_dx"]: the input
    needs a gradient (a hidden layer), which selects the project-first kernels."""

    def __init__(self, cfg, dev):
        n, H, D, f_in = cfg["n"], cfg["H"], cfg["D"], cfg["f_in"]
        rowptr, col, self.x, self.gout = _inputs(cfg, dev)
        self.graph = Graph.from_csr(rowptr, col)
        self.graph.transpose()  # cached per adjacency, like the CSR itself; not part of a step
        self.e_total = self.graph.nnz
        self.n_long_rows = long_rows(rowptr)
        self.needs_dx = bool(cfg.get("needs_dx", False))
        if self.needs_dx:
            self.x.requires_grad_(True)
        self.Ws, self.a_src, self.a_dst = init_layer_params(f_in, H, D, dev, seed=72)
        self.params = self.Ws + self.a_src + self.a_dst
        self.cfg = cfg
        self.x_host = None
        self.row_cost = float("nan")

    def _layer(self, x):
        return gat_layer(x, self.graph, self.Ws, self.a_src, self.a_dst, None, 0.2, concat=True)

    def step(self):
        for p in self.params:
            p.grad = None
        self.x.grad = None
        y = self._layer(self.x)
        y.backward(self.gout)
        return y

    def e2e(self, steps):
        return pipelined_e2e(self, steps)


class ShardedLayerBench:
    """The same workload on `world` GPUs: the whole synthetic graph is generated identically on every rank (same
    seed) and each rank keeps its destination-row shard.

    Shard boundaries equalise cost(row range) = stored entries + row_cost * rows.  row_cost starts from the
    single-GPU measurement and, with calibrate=True, is re-fitted once from what the ranks actually measure:
    a few untimed steps give every rank's kernel time t_r (collective waits excluded), least squares over the
    ranks gives t = a * rows + b * entries, and the shards are cut again with row_cost = a / b.  This is setup
    work, like the CSR build: it happens before the timed region."""

    def __init__(self, cfg, rank: int, world: int, dev, calibrate: bool = True):
        self.rank, self.world, self.dev, self.cfg = rank, world, dev, cfg
        rowptr, col, x, gout = _inputs(cfg, dev)
        self.e_total = int(col.numel())
        self.n_long_rows = long_rows(rowptr)
        self.needs_dx = bool(cfg.get("needs_dx", False))
        self.Ws, self.a_src, self.a_dst = init_layer_params(cfg["f_in"], cfg["H"], cfg["D"], dev, seed=72)
        self.params = self.Ws + self.a_src + self.a_dst
        self.row_cost = float(ROW_COST)
        self._cut(rowptr, col, x, gout)
        if calibrate and world > 1:
            fitted = self._fit_row_cost()
            if fitted is not None and abs(fitted - self.row_cost) > 0.5:
                self.row_cost = fitted
                self._cut(rowptr, col, x, gout)
        del rowptr, col, x, gout
        torch.cuda.empty_cache()
        self.x_host = None

    def _cut(self, rowptr, col, x, gout):
        self.plan = ShardPlan(shard_rows_by_cost(rowptr, self.world, self.row_cost), self.rank)
        self.graph = self.plan.local_graph(rowptr, col)
        self.graph_t = None
        if self.needs_dx:   # hidden layer: its backward walks this rank's SOURCE rows (sharded.SourceShard)
            self.graph_t = self.plan.source_shard(rowptr, col)
        else:
            self.graph.transpose()
        self.x = self.plan.rows(x).clone()
        if self.needs_dx:
            self.x.requires_grad_(True)
        self.gout = self.plan.rows(gout).clone()

    def _fit_row_cost(self, steps: int = 3):
        """Least-squares (a, b) of t_r = a * rows_r + b * entries_r over the ranks -> a / b, or None when the
        fit is unusable (fewer than 2 distinct shard shapes, non-positive coefficients)."""
        for _ in range(2):
            self.step()
        torch.cuda.synchronize()
        old = _lib.timer
        _lib.timer = _lib.KernelTimer()
        for _ in range(steps):
            self.step()
        kern = _lib.timer.summary()
        _lib.timer = old
        t = sum(v["ms_total"] for k, v in kern.items() if not k.startswith("comm:")) / steps
        mine = torch.tensor([float(self.plan.n_local), float(self.graph.nnz), t], dtype=torch.float64, device=self.dev)
        allr = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allr, mine)
        m = torch.stack(allr).cpu()
        return fit_row_cost(m[:, 0], m[:, 1], m[:, 2])

    def _layer(self, x):
        return sharded_gat_layer(x, self.graph, self.plan, self.Ws, self.a_src, self.a_dst, None, 0.2, concat=True,
                                 graph_t=self.graph_t)

    def step(self):
        for p in self.params:
            p.grad = None
        self.x.grad = None
        y = self._layer(self.x)
        y.backward(self.gout)
        return y

    def e2e(self, steps: int):
        return pipelined_e2e(self, steps, barrier=dist.barrier)


class ShardOnlyLayerBench:
    """papers100M-shaped workload (BASELINE.json config 5): every rank generates and keeps ONLY its own destination-row
    shard (cfg["n"] rows per rank, sources over all world * cfg["n"] nodes, synth.power_law_shard) -- the whole graph
    (1.6 G entries at 8 ranks) and the N x N adjacency of the reference never exist anywhere.  First layer with a
    static input, so the aggregate-first form is forced: the x columns of the gathered rows are exchanged once, a
    step moves g [N, H] forward and dg [N, H] backward.  Weak scaling by construction (per-rank shard fixed)."""

    def __init__(self, cfg, rank: int, world: int, dev):
        from pygat_b200.synth import power_law_shard
        self.rank, self.world, self.dev, self.cfg = rank, world, dev, cfg
        n, H, D, f_in = cfg["n"], cfg["H"], cfg["D"], cfg["f_in"]
        self.plan = ShardPlan([r * n for r in range(world + 1)], rank)
        rowptr, col = power_law_shard(world * n, rank * n, (rank + 1) * n, cfg["avg_deg"], seed=72,
                                      exponent=cfg["exponent"], device=dev)
        self.graph = Graph(rowptr, col, n_src=world * n)
        e = torch.tensor([self.graph.nnz, long_rows(rowptr)], dtype=torch.int64, device=dev)
        del rowptr, col
        torch.cuda.empty_cache()
        if world > 1:
            dist.all_reduce(e)
        self.e_total, self.n_long_rows = int(e[0].item()), int(e[1].item())
        g = torch.Generator(device=dev).manual_seed(72 + rank)
        self.x = torch.randn(n, f_in, generator=g, device=dev)
        self.gout = torch.randn(n, H * D, generator=g, device=dev)
        self.Ws, self.a_src, self.a_dst = init_layer_params(f_in, H, D, dev, seed=72)
        self.params = self.Ws + self.a_src + self.a_dst
        self.needs_dx = False
        self.row_cost = float("nan")
        self.x_host = None

    def _layer(self, x):
        return sharded_gat_layer(x, self.graph, self.plan, self.Ws, self.a_src, self.a_dst, None, 0.2, concat=True,
                                 form="agg_first")

    def step(self):
        for p in self.params:
            p.grad = None
        y = self._layer(self.x)
        y.backward(self.gout)
        return y

    def e2e(self, steps: int):
        """Host buffers in, gradients out, SINGLE device buffer (the shard's working set leaves no room for a second
        copy of the features at the papers shape): the copy of step k+1 waits for step k."""
        if self.x_host is None:
            self.x_host = self.x.cpu().pin_memory()
        n_par = sum(p.numel() for p in self.params)
        host_out = torch.empty(n_par + 1, dtype=torch.float32).pin_memory()

        def run(k):
            for _ in range(k):
                self.x.copy_(self.x_host, non_blocking=True)   # in place: bumps the version counter, the kept rows go stale
                y = self.step()
                flat = torch.cat([p.grad.reshape(-1) for p in self.params] + [y.detach()[:: max(1, y.shape[0] // 1024)].sum().reshape(1)])
                host_out.copy_(flat, non_blocking=True)
                del y
        run(1)
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, self.x_host.numel() * 4, host_out.numel() * 4


def pipelined_e2e(runner, steps, barrier=None):
    """End to end through the public API with HOST inputs: every step copies its input features from
    pinned host memory (double buffered on a copy stream, so step k+1's copy overlaps step k's compute,
    as a training input pipeline would) and reads the step's results (all parameter gradients + an
    output checksum; the layer output itself stays on the device for the next layer) back to pinned host
    memory.  The pipeline is timed in steady state: the input of the first timed step was copied while the
    last warm-up step ran, every timed step issues the copy for the step after it (the last one included) and
    the region ends only when that copy has landed -- `steps` copies and `steps` steps inside the region.
    Returns (ms per step, H2D bytes, D2H bytes)."""
    dev = runner.x.device
    if runner.x_host is None:
        runner.x_host = runner.x.detach().cpu().pin_memory()
    n_par = sum(p.numel() for p in runner.params)
    host_out = torch.empty(n_par + 1, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(runner.x), torch.empty_like(runner.x)]
    if runner.needs_dx:
        for b in bufs:
            b.requires_grad_(True)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    main = torch.cuda.current_stream()
    done = [0]

    def issue_copy(k):
        b = k & 1
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(freed[b])  # the step that last used this buffer is done with it
            bufs[b].copy_(runner.x_host, non_blocking=True)
            ready[b].record(copy_stream)

    def run(n_steps):
        for _ in range(n_steps):
            k = done[0]
            b = k & 1
            main.wait_event(ready[b])
            issue_copy(k + 1)
            for p in runner.params:
                p.grad = None
            bufs[b].grad = None
            y = runner._layer(bufs[b])
            y.backward(runner.gout)
            freed[b].record(main)
            flat = torch.cat([p.grad.reshape(-1) for p in runner.params] +
                             [y.detach()[:: max(1, y.shape[0] // 1024)].sum().reshape(1)])
            host_out.copy_(flat, non_blocking=True)
            done[0] = k + 1
        main.wait_event(ready[done[0] & 1])  # the copy the last step issued

    for b in (0, 1):
        freed[b].record(main)
    issue_copy(0)
    run(2)
    if barrier is not None:
        barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, runner.x_host.numel() * 4, host_out.numel() * 4


# ---------------------------------------------------------------------- sharded step == single-GPU step
CHECK_ROWS = 256
CHECK_TOL = 1e-5


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


def parity_check(runner: ShardedLayerBench, rank: int, world: int, dev):
    """bench.py --gpus N: one more (untimed) step on the shards, then rank 0 runs the SAME step on the whole graph
    by itself and compares: the first CHECK_ROWS output rows and the output column sums of every rank's shard, every
    parameter gradient and (hidden-layer workloads) the same rows of the input gradient.  Collective: all ranks
    call it.  Returns the report on rank 0 (None elsewhere)."""
    y = runner.step()
    torch.cuda.synchronize()
    k = min(CHECK_ROWS, y.shape[0])
    mine = {"y_rows": y.detach()[:k].cpu(), "y_colsum": y.detach().double().sum(0).cpu(),
            "grads": [p.grad.detach().cpu() for p in runner.params],
            "dx_rows": runner.x.grad[:k].cpu() if runner.needs_dx else None, "bounds": runner.plan.bounds}
    del y
    gathered = [None] * world
    dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
    report = None
    if rank == 0:
        single = SingleGpuLayerBench(runner.cfg, dev)
        y_ref = single.step()
        torch.cuda.synchronize()
        errs = {"y_rows": 0.0, "y_colsum": 0.0, "param_grads": 0.0}
        if runner.needs_dx:
            errs["dx_rows"] = 0.0
        scale_y = y_ref.detach().abs().max().item()
        bounds = mine["bounds"]
        for r, got in enumerate(gathered):
            lo, hi = bounds[r], bounds[r + 1]
            kk = got["y_rows"].shape[0]
            ref_rows = y_ref.detach()[lo:lo + kk].cpu()
            errs["y_rows"] = max(errs["y_rows"], (got["y_rows"].double() - ref_rows.double()).abs().max().item() / scale_y)
            cs = y_ref.detach()[lo:hi].double().sum(0).cpu()
            # a column sum over ~n/world rows of O(1) terms: compare relative to the sum of magnitudes
            mag = y_ref.detach()[lo:hi].double().abs().sum(0).cpu().clamp_min(1e-30)
            errs["y_colsum"] = max(errs["y_colsum"], ((got["y_colsum"] - cs).abs() / mag).max().item())
            for gp, ps in zip(got["grads"], single.params):
                errs["param_grads"] = max(errs["param_grads"], _rel(gp, ps.grad.cpu()))
            if runner.needs_dx:
                errs["dx_rows"] = max(errs["dx_rows"], _rel(got["dx_rows"], single.x.grad[lo:lo + kk].cpu()))
        report = {"parity_ok": all(v < CHECK_TOL for v in errs.values()), "tol": CHECK_TOL,
                  "max_rel_err": {k2: float(f"{v:.3e}") for k2, v in errs.items()},
                  "against": "single-GPU run of the same step on rank 0 (same seed, whole graph)",
                  "rows_per_rank": k}
        del single, y_ref
        torch.cuda.empty_cache()
    dist.barrier()
    return report
