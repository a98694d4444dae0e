"""ctypes binding of libgatk.so (the C ABI declared in include/gatk.h).

There is no CPU fallback: importing works anywhere (so the symbol table can be checked
without a GPU), but every compute entry point needs CUDA tensors and raises otherwise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgatk.so")

P = c_void_p  # every device pointer / stream crosses the ABI as void*

# name -> (restype, argtypes); mirrors include/gatk.h one to one
PROTOTYPES = {
    "gatk_version": (c_int, []),
    "gatk_last_error": (c_char_p, []),
    "gatk_launch_count": (c_int64, []),
    "gatk_sm_count": (c_int, []),
    "gatk_scan_workspace_bytes": (c_size_t, [c_int64]),
    "gatk_csr_from_dense_rowptr": (c_int, [P, c_int64, c_int64, c_int64, c_int, P, P, c_size_t, P]),
    "gatk_csr_from_dense_fill": (c_int, [P, c_int64, c_int64, c_int64, c_int, P, P, P]),
    "gatk_csr_from_coo": (c_int, [P, P, c_int64, c_int64, P, P, P, c_size_t, P]),
    "gatk_transpose_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "gatk_csr_transpose": (c_int, [c_int64, c_int64, c_int64, P, P, P, P, P, P, c_size_t, P]),
    "gatk_dropout_keep_mask": (c_int, [P, c_int64, c_float, c_uint64, c_uint64, P]),
    "gatk_mask_scale": (c_int, [P, c_int64, P, c_float, P, c_int64, c_int64, c_int64, P]),
    "gatk_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int64, c_int64, c_int64]),
    "gatk_gemm_uses_tensor_cores": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int]),
    "gatk_gemm": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, P, c_int64, P, c_int64, P, c_int64, c_int,
                          P, c_size_t, P]),
    "gatk_gemm_batched_workspace_bytes": (c_size_t, [c_int, c_int, c_int64, c_int64, c_int64, c_int]),
    "gatk_gemm_batched_fuses_elu_grad": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_int, c_int64, c_int64, c_int64,
                                                 c_int64, c_int64, c_int64]),
    "gatk_gemm_batched": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_int, P, c_int64, c_int64, P, c_int64, c_int64,
                                  P, c_int64, c_int64, c_int, P, c_int64, P, c_size_t, P]),
    "gatk_logits_fwd": (c_int, [c_int64, c_int, c_int, P, c_int64, P, c_float, P, P, P, P, c_uint64, c_uint64, c_float, P]),
    "gatk_gemm_heads_dropout_ws_floats": (c_size_t, [c_int64, c_int, c_int, c_int, c_int]),
    "gatk_gemm_heads_dropout_fwd": (c_int, [c_int64, c_int, c_int, c_int, c_int, P, c_int64, P, c_int64, P, c_int64, c_uint64,
                                            c_uint64, c_float, P]),
    "gatk_gemm_heads_dropout_dw": (c_int, [c_int64, c_int, c_int, c_int, c_int, P, c_int64, P, c_int64, P, c_int64, P, c_uint64,
                                           c_uint64, c_float, P]),
    "gatk_gemm_heads_dropout_dx": (c_int, [c_int64, c_int, c_int, c_int, c_int, P, c_int64, P, c_int64, P, c_int64, c_uint64,
                                           c_uint64, c_float, P]),
    "gatk_hub_scratch_floats": (c_size_t, [c_int, c_int, c_int, c_int]),
    "gatk_attn_fwd": (c_int, [c_int64, P, P, c_int, c_int, P, c_int64, P, c_int64, P, c_int64, P, c_float, c_float,
                              P, c_int64, c_int, P, P, c_int64, P,
                              c_int, P, P, c_int, c_int, P, P, P, c_int, c_uint64, c_uint64, c_float, P]),
    "gatk_attn_bwd_record_ld": (c_int64, [c_int, c_int]),
    "gatk_attn_bwd_prep": (c_int, [c_int64, c_int, c_int, P, c_int64, P, c_int64, c_int, P, c_int64, P, c_int64, P,
                                   P, c_int64, P, c_int64, P]),
    "gatk_attn_bwd_fused": (c_int, [c_int64, P, P, P, c_int, c_int, P, c_int64, P, c_int64, P, c_int64, P, c_float,
                                    c_float, P, P, c_int64, P, c_int64, P, P, c_int64,
                                    c_int, P, P, c_int, c_int, P, P, P, c_int, c_uint64, c_uint64, c_float, P]),
    "gatk_attn_bwd_finish": (c_int, [c_int64, P, c_int, c_int, P, P, P, c_float, P, c_int64, P, c_int64,
                                     c_int, P, P, c_int, c_int, P, c_uint64, c_uint64, c_float, P]),
    "gatk_da_workspace_floats": (c_size_t, [c_int, c_int]),
    "gatk_da_reduce": (c_int, [c_int64, c_int, c_int, P, c_int64, P, P, P, P, P, P]),
    "gatk_xg_pitch": (c_int64, [c_int, c_int]),
    "gatk_logits_pack": (c_int, [c_int64, c_int, c_int, P, c_int64, P, c_int64, P, c_int64, P, c_int64, P]),
    "gatk_logits_pack_push": (c_int, [c_int64, c_int, c_int, P, c_int64, P, c_int64, P, c_int64, P, c_int64, c_int, P, c_int, P]),
    "gatk_attn_x_scratch_floats": (c_size_t, [c_int, c_int, c_int, c_int]),
    "gatk_attn_x_fwd": (c_int, [c_int64, c_int64, P, P, c_int, c_int, P, c_int64, P, c_int64, c_float, P, c_int64, P,
                                c_int, P, P, c_int, c_int, P, P, P, c_int, P]),
    "gatk_attn_x_bwd": (c_int, [c_int64, c_int64, P, P, c_int, c_int, P, c_int64, P, c_int64, P, c_float, P, c_int64, P,
                                c_int64, P, P, P, c_int64, P, c_int64, c_int, P, P, c_int, c_int, P, P, P, c_int, P]),
    "gatk_edge_tsum": (c_int, [c_int64, P, P, c_int, P, P, c_int64, c_int, P, c_int, P]),
    "gatk_elu_fwd": (c_int, [c_int64, c_int64, P, c_int64, P]),
    "gatk_elu_bwd": (c_int, [c_int64, c_int64, P, c_int64, P, c_int64, P, c_int64, P]),
    "gatk_attn_v2_fwd": (c_int, [c_int64, P, P, c_int, c_int, P, c_int64, P, P, c_float, c_float, c_int, c_int, P, P, c_int64, P,
                                 P, P]),
    "gatk_attn_v2_bwd": (c_int, [c_int64, P, P, c_int, c_int, P, c_int64, P, P, c_float, c_float, c_int, c_int, P, P, c_int64,
                                 P, P, c_int64, P, P, P, P]),
    "gatk_head_combine": (c_int, [c_int64, c_int, c_int, c_int, P, c_int64, c_int, P, P]),
    "gatk_head_combine_bwd": (c_int, [c_int64, c_int, c_int, c_int, P, c_int, P, c_int64, P]),
    "gatk_nll_head_fwd": (c_int, [c_int64, P, P, c_int64, P, c_int, P, P]),
    "gatk_nll_head_bwd": (c_int, [c_int64, P, P, c_int64, P, c_int, P, c_float, P, c_int64, P]),
    "gatk_bce_f1_fwd": (c_int, [c_int64, P, P, P, P]),
    "gatk_bce_bwd": (c_int, [c_int64, P, P, P, c_float, P, P]),
    "gatk_spmm_coo_fwd": (c_int, [P, P, P, c_int64, c_int64, P, P, P]),
    "gatk_spmm_coo_bwd": (c_int, [P, P, P, c_int64, c_int64, P, P, P, P, P]),
}

_lib = None


class GatkError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libgatk.so (building it first if the toolkit is present and it is stale/missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) or os.environ.get("GATK_REBUILD"):
        from . import _build
        _build.build(force=bool(os.environ.get("GATK_REBUILD")))
    if not os.path.exists(LIB_PATH):
        raise GatkError(f"{LIB_PATH} is missing: the CUDA extension is required (there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class KernelTimer:
    """Optional per-entry-point CUDA-event timing (bench.py's roofline numbers).  Events are
    recorded on the stream the call launches on; nothing synchronises until summary()."""

    def __init__(self):
        self.records = []  # (name, start_event, end_event)
        self.calls = 0

    def summary(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.records:
            tot, cnt = out.get(name, (0.0, 0))
            out[name] = (tot + e0.elapsed_time(e1), cnt + 1)
        return {k: {"ms_total": v[0], "calls": v[1], "ms_avg": v[0] / v[1]} for k, v in out.items()}


timer: "KernelTimer | None" = None  # set by bench.py around its timed region
call_count = 0                      # entry-point invocations (every one launches >= 1 kernel)


def call(name: str, *args, label: "str | None" = None):
    """Invoke an int-returning entry point; non-zero -> GatkError(gatk_last_error()).
    label: key the optional KernelTimer files this call under (default: the entry point's name)."""
    global call_count
    lib = load()
    call_count += 1
    if timer is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        timer.records.append((label or name, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise GatkError(f"{name} failed ({rc}): {lib.gatk_last_error().decode(errors='replace')}")


class timed:
    """Context manager filing a non-library region (a collective, a torch op) under `label` in the KernelTimer."""

    def __init__(self, label: str):
        self.label = label

    def __enter__(self):
        if timer is not None:
            import torch
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if timer is not None:
            import torch
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            timer.records.append((self.label, self.e0, e1))
        return False


def query(name: str, *args):
    """Invoke a size/value query (no error code)."""
    return getattr(load(), name)(*args)
