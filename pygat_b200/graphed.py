"""CUDA-graph capture of a whole training step (forward, loss head, backward, optimizer) on a FIXED batch: the body of
the reference's per-batch loop, train_ppi.py:117-124 (zero_grad, model(features, adj), BCE loss, backward,
optimizer.step), and of its full-batch step, train.py:154-162.

A PPI batch is ~4.5 k nodes: its 3-layer step is a few hundred small launches (the reference: ~25 ATen launches per
head and direction; this engine: one projection + one attention kernel per layer and direction plus the
parameter-sized packing ops), so the step is bound by launch latency, not by bandwidth.  The
engine's calls are stream-ordered and host-sync free once the adjacency's CSR is cached, which makes the whole step
capturable: one graph per distinct batch, replayed every epoch.

Constraints (checked): dropout must be inactive (the engine seeds its masks from torch's CPU generator, which a graph
would freeze), the optimizer must be capture-safe (torch.optim.Adam(..., capturable=True)), and the step function
must not synchronise (use pygat_b200.heads for the loss and metric)."""
from __future__ import annotations

import torch


class GraphedStep:
    """step_fn() -> tuple of device tensors.  Warm-up runs (lazy kernel configuration, CSR cache, optimizer state
    allocation) happen on a side stream and are UNDONE afterwards (parameters and optimizer state restored in
    place), so replaying the graph k times equals k eager steps from the state the model was in."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, step_fn, warmup: int = 2, pool=None):
        for m in model.modules():
            if getattr(m, "dropout", 0.0) and m.training and isinstance(getattr(m, "dropout"), float) and m.dropout > 0.0:
                raise RuntimeError("GraphedStep: dropout is active (p > 0 in training mode); its masks cannot be captured")
        for grp in optimizer.param_groups:
            if not grp.get("capturable", False):
                raise RuntimeError("GraphedStep needs a capture-safe optimizer, e.g. torch.optim.Adam(..., capturable=True)")
        params = [p for grp in optimizer.param_groups for p in grp["params"]]
        saved_p = [p.detach().clone() for p in params]
        saved_s = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in optimizer.state.get(p, {}).items()} for p in params]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():
            for p, sp, ss in zip(params, saved_p, saved_s):
                p.copy_(sp)
                for k, v in optimizer.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        if k in ss:
                            v.copy_(ss[k])
                        else:
                            v.zero_()
        optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.out = step_fn()

    def pool(self):
        return self.graph.pool()

    def __call__(self):
        self.graph.replay()
        return self.out
