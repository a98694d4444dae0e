"""Workspace allocation for the engine's host side.

Every buffer a kernel is expected to fully overwrite is allocated through `empty` / `empty_like`.
With GATK_POISON=1 those buffers are filled with NaN (float) or 0xFF bytes (integer) first, so a kernel
that reads a workspace element nobody wrote turns the result into NaN / an out-of-range index instead of
silently depending on whatever the caching allocator handed back (tests/test_gpu_determinism.py runs
the parity cases that way).  `trace` (a list, or None) collects every buffer for debugging scripts.
"""
from __future__ import annotations

import os

import torch

POISON = os.environ.get("GATK_POISON", "0") not in ("", "0")
trace = None


def _poison(t: torch.Tensor) -> torch.Tensor:
    if t.numel():
        if t.is_floating_point():
            t.fill_(float("nan"))
        else:
            t.view(torch.uint8).fill_(0xFF)
    return t


def empty(*size, dtype=torch.float32, device=None) -> torch.Tensor:
    t = torch.empty(*size, dtype=dtype, device=device)
    if POISON:
        _poison(t)
    if trace is not None:
        trace.append(t)
    return t


def empty_like(ref: torch.Tensor) -> torch.Tensor:
    t = torch.empty_like(ref)
    if POISON:
        _poison(t)
    if trace is not None:
        trace.append(t)
    return t


def zeros(*size, dtype=torch.float32, device=None) -> torch.Tensor:
    """A buffer that must START at zero (accumulation targets of reductions); traced like the others."""
    t = torch.zeros(*size, dtype=dtype, device=device)
    if trace is not None:
        trace.append(t)
    return t
