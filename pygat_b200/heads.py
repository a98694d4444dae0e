"""Fused loss heads + metrics of the reference's training steps, csrc/heads.cu.

    citation_head(logits, labels, idx)  ==  F.nll_loss(F.log_softmax(F.elu(logits), 1)[idx], labels[idx])  and the
                                            accuracy of those rows                    (train.py:151-160, utils.py:92-96)
    ppi_head(logits, labels)            ==  BCEWithLogitsLoss(mean)(logits, labels) and the micro-F1 of logits > 0
                                                                                      (train_ppi.py:106-120)

Both return device tensors only (loss with autograd, metric detached): the reference's `.item()` / `.cpu().numpy()`
per step is the caller's choice, not a side effect of computing the loss.  `fused_steps` holds sync-free versions of
the two training steps built on them."""
from __future__ import annotations

import torch

from . import _lib
from .graph import _ptr, _require_cuda, _stream, on_device


class CitationHeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, idx):
        _require_cuda(logits, "logits")
        logits = logits.contiguous().float()
        n, C = logits.shape
        idx = None if idx is None else idx.contiguous().long()
        labels = labels.contiguous().long()
        n_idx = n if idx is None else idx.numel()
        stats = torch.zeros(2, dtype=torch.float64, device=logits.device)
        with on_device(logits):
            _lib.call("gatk_nll_head_fwd", n_idx, _ptr(idx), logits.data_ptr(), C, labels.data_ptr(), C, stats.data_ptr(), _stream())
        ctx.save_for_backward(logits, labels, idx if idx is not None else torch.empty(0, device=logits.device))
        ctx.has_idx, ctx.n_idx = idx is not None, n_idx
        out = (stats / max(n_idx, 1)).float()
        loss, acc = out[0].clone(), out[1].clone()
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, gloss, _gacc):
        logits, labels, idx = ctx.saved_tensors
        n, C = logits.shape
        d = torch.zeros_like(logits)
        g = gloss.contiguous().float().reshape(1)
        with on_device(logits):
            _lib.call("gatk_nll_head_bwd", ctx.n_idx, idx.data_ptr() if ctx.has_idx else None, logits.data_ptr(), C,
                      labels.data_ptr(), C, g.data_ptr(), 1.0 / max(ctx.n_idx, 1), d.data_ptr(), C, _stream())
        return d, None, None


def citation_head(logits: torch.Tensor, labels: torch.Tensor, idx: "torch.Tensor | None" = None):
    """(loss, accuracy) of train.py:151-160 on the rows `idx` (all rows if None), as 0-d device tensors."""
    return CitationHeadFunction.apply(logits, labels, idx)


class PpiHeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        _require_cuda(logits, "logits")
        logits = logits.contiguous().float()
        labels = labels.contiguous().float()
        total = logits.numel()
        stats = torch.zeros(4, dtype=torch.float64, device=logits.device)
        with on_device(logits):
            _lib.call("gatk_bce_f1_fwd", total, logits.data_ptr(), labels.data_ptr(), stats.data_ptr(), _stream())
        ctx.save_for_backward(logits, labels)
        loss = (stats[0] / max(total, 1)).float()
        denom = 2 * stats[1] + stats[2] + stats[3]
        f1 = torch.where(denom > 0, 2 * stats[1] / denom.clamp_min(1.0), torch.zeros_like(denom)).float()
        ctx.mark_non_differentiable(f1)
        return loss, f1

    @staticmethod
    def backward(ctx, gloss, _gf1):
        logits, labels = ctx.saved_tensors
        d = torch.empty_like(logits)
        g = gloss.contiguous().float().reshape(1)
        with on_device(logits):
            _lib.call("gatk_bce_bwd", logits.numel(), logits.data_ptr(), labels.data_ptr(), g.data_ptr(),
                      1.0 / max(logits.numel(), 1), d.data_ptr(), _stream())
        return d, None


def ppi_head(logits: torch.Tensor, labels: torch.Tensor):
    """(loss, micro-F1) of train_ppi.py:106-120 as 0-d device tensors (sklearn's f1_score(average='micro') of logits > 0)."""
    return PpiHeadFunction.apply(logits, labels)


# ---------------------------------------------------------------------- sync-free training steps
def citation_epoch(model, optimizer, features, adj, labels, idx_train, idx_val, fastmode: bool = False):
    """One epoch of train.py:154-179 -- train step, then (unless fastmode) an eval-mode forward for the validation
    loss -- without a host synchronisation: returns (loss_train, acc_train, loss_val, acc_val) as device tensors.
    The reference prints all four every epoch (four .item() syncs) and keeps loss_val on the host for early stopping;
    a caller that wants that reads the tensors, ideally every k-th epoch."""
    model.train()
    optimizer.zero_grad(set_to_none=True)
    out = model(features, adj)
    loss_train, acc_train = citation_head(out, labels, idx_train)
    loss_train.backward()
    optimizer.step()
    if not fastmode:
        model.eval()
        with torch.no_grad():
            out = model(features, adj)
    with torch.no_grad():
        loss_val, acc_val = citation_head(out.detach(), labels, idx_val)
    return loss_train.detach(), acc_train, loss_val, acc_val


def ppi_batch_step(model, optimizer, features, labels, adj, allreduce=None):
    """One training batch of train_ppi.py:117-124 without the per-batch device->host copy of logits and labels for
    sklearn: returns (loss, micro-F1) as device tensors.  allreduce(params, n_nodes): graph-level data parallelism
    (sharded.allreduce_gradients)."""
    optimizer.zero_grad(set_to_none=True)
    out = model(features, adj)
    loss, f1 = ppi_head(out, labels)
    loss.backward()
    if allreduce is not None:
        allreduce(model.parameters(), features.shape[0])
    optimizer.step()
    return loss.detach(), f1
