"""Host side of the fused GAT layer: one autograd.Function per LAYER (all heads batched)
driving the C-ABI kernels of libgatk.so.

What it replaces (reference /root/reference): the per-head Python loop of models.py:29-35
around SpGraphAttentionLayer.forward (layers.py:125-173) / GraphAttentionLayer.forward
(layers.py:32-53) and autograd through SpecialSpmmFunction (layers.py:70-90).

Layout: heads are packed side by side, each padded from D to Dp = 4*2^k columns, so the
projected features are one [N, H*Dp] matrix (== torch.cat over heads, models.py:32, when
D == Dp).  The skip projection (layers.py:48,166) rides in the same GEMM as extra columns.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib, _mem
from .graph import Graph, _ptr, _require_cuda, _stream, on_device


import os

# The folded form needs no logits pass, no da reduction and no dWh read-modify-write; GATK_FOLD=0 keeps
# the explicit kernels even without dropout (used by the tests to compare the two forms).
FOLD_LOGITS = os.environ.get("GATK_FOLD", "1") != "0"
# The aggregate-first form (neighbour sum before the projection) for narrow first-layer inputs; GATK_AGG_FIRST=0
# keeps the project-first kernels (the tests compare the forms).
AGG_FIRST = os.environ.get("GATK_AGG_FIRST", "1") != "0"
# Aggregate-first backward: dg_j = sum_i ds_ij is accumulated with red.global.add inside the edge pass (default), or --
# GATK_DETERMINISTIC=1 -- by writing ds and summing it along the transposed pattern in a fixed order (bit-reproducible;
# one more kernel and 2.3 GB more traffic at the products shape).
DETERMINISTIC = os.environ.get("GATK_DETERMINISTIC", "0") not in ("", "0")
# ELU' folded into the operand paths of the backward products (GATK_FUSE_ELU_GRAD=0: separate gatk_elu_bwd pass)
FUSE_ELU_GRAD = os.environ.get("GATK_FUSE_ELU_GRAD", "1") != "0"


def padded_width(d: int) -> int:
    """Per-head width the kernels use: 4 * 2^k >= d (float4 slots, power-of-two slots per head)."""
    l = 1
    while 4 * l < d:
        l *= 2
    return 4 * l


@dataclass
class LayerMasks:
    """Explicit keep masks (uint8, 1 = keep) for the three dropout sites of a layer
    (layers.py:34/132 input, :37/136 projected features, :43/153 attention).
    keep_in [H, N, F]; keep_wh [N, H*Dp]; keep_att [E, H].  None = site not dropped."""
    keep_in: Optional[torch.Tensor] = None
    keep_wh: Optional[torch.Tensor] = None
    keep_att: Optional[torch.Tensor] = None
    # seeded mode (training with p > 0 and no injected masks): nothing is materialised; the kernels evaluate the Philox
    # stream (seed, offset of the site) where they consume a decision.  offsets = (input, projected, attention).
    seed: Optional[int] = None
    offsets: tuple = (0, 0, 0)


def _gemm(ta, tb, M, N, K, A, lda, B, ldb, C, ldc, accumulate=0, a_off=0, b_off=0, c_off=0, label=None):
    """C[M,N] (+)= op(A) op(B) on raw pointers; *_off are element offsets into the tensors."""
    ws_bytes = _lib.query("gatk_gemm_workspace_bytes", ta, tb, M, N, K)
    ws = _mem.empty(ws_bytes, dtype=torch.uint8, device=C.device) if ws_bytes else None
    _lib.call("gatk_gemm", ta, tb, M, N, K, A.data_ptr() + 4 * a_off, lda, B.data_ptr() + 4 * b_off, ldb,
              C.data_ptr() + 4 * c_off, ldc, accumulate, _ptr(ws), ws_bytes, _stream(), label=label)


def _gemm_batched(ta, tb, M, N, K, batches, A, lda, a_bs, B, ldb, b_bs, C, ldc, c_bs, epilogue=0, label=None,
                  elu_out=None, ld_elu=0):
    """C_b = op(A_b) op(B_b) for the `batches` column blocks A_b = A + b*a_bs, ... (one launch for all heads).
    elu_out: the dh' operand (A of an NT product, B of a TN product) is gout * ELU'(.) formed inside the kernel."""
    ws_bytes = _lib.query("gatk_gemm_batched_workspace_bytes", ta, tb, M, N, K, batches)
    ws = _mem.empty(ws_bytes, dtype=torch.uint8, device=C.device) if ws_bytes else None
    _lib.call("gatk_gemm_batched", ta, tb, M, N, K, batches, A.data_ptr(), lda, a_bs, B.data_ptr(), ldb, b_bs,
              C.data_ptr(), ldc, c_bs, epilogue, _ptr(elu_out), ld_elu, _ptr(ws), ws_bytes, _stream(), label=label)


def _elu_grad_fusable(n, f_in, H, Dp, Fp, M_out, HD) -> bool:
    """True when both backward products of the aggregate-first form (dW_h = xagg_h^T dh'_h, dxagg_h = dh'_h W_h^T)
    take the batched tensor-core kernels that form dh' = gout * ELU'(out) on the fly."""
    tn = _lib.query("gatk_gemm_batched_fuses_elu_grad", 1, 0, f_in, Dp, n, H, H * Fp, Fp, HD, Dp, M_out, Dp)
    nt = _lib.query("gatk_gemm_batched_fuses_elu_grad", 0, 1, n, f_in, Dp, H, HD, Dp, M_out, Dp, H * Fp, Fp)
    return bool(tn and nt)


def _hub_scratch(which: int, H: int, Dp: int, n_seg: int, dev):
    if not n_seg:
        return None
    return _mem.empty(_lib.query("gatk_hub_scratch_floats", which, H, Dp, n_seg), dtype=torch.float32, device=dev)


def site_offsets(n: int, f_in: int, H: int, Dp: int, nnz: int):
    """Philox counter offsets of the three dropout sites of a layer: consecutive, non-overlapping ranges of the
    stream (a site of `sz` elements uses ceil(sz / 4) counters)."""
    sizes = (H * n * f_in, n * H * Dp, nnz * H)
    offs, off = [], 0
    for sz in sizes:
        offs.append(off)
        off += (sz + 3) // 4
    return tuple(offs), sizes


def random_masks(n: int, f_in: int, H: int, Dp: int, nnz: int, p: float, device, seed: Optional[int] = None) -> LayerMasks:
    """MATERIALISE the three masks of a layer from the in-library Philox stream (debugging / tests: the seeded mode
    of the layer consumes the same stream without ever building them)."""
    if seed is None:
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
    offs, sizes = site_offsets(n, f_in, H, Dp, nnz)
    bufs = []
    for sz, off in zip(sizes, offs):
        t = _mem.empty(sz, dtype=torch.uint8, device=device)
        _lib.call("gatk_dropout_keep_mask", t.data_ptr(), sz, float(p), seed, off, _stream())
        bufs.append(t)
    return LayerMasks(bufs[0].view(H, n, f_in), bufs[1].view(n, H * Dp), bufs[2].view(nnz, H))


def seeded_masks(n: int, f_in: int, H: int, Dp: int, nnz: int, seed: Optional[int] = None) -> LayerMasks:
    """Dropout sites as (seed, offsets): the 64-bit seed comes from torch's CPU generator, so torch.manual_seed
    (train.py:91-99) makes runs reproducible; no mask is materialised."""
    if seed is None:
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
    return LayerMasks(seed=seed & 0x7FFFFFFFFFFFFFFF, offsets=site_offsets(n, f_in, H, Dp, nnz)[0])


class GatLayerFunction(torch.autograd.Function):
    """out[N, H*Dp] = act( softmax_j(LeakyReLU(f_i + g_j)) @ Wh  (+ x_drop @ S) ), all heads."""

    @staticmethod
    def forward(ctx, x, w_ext, a_src, a_dst, graph: Graph, H: int, Dp: int, has_skip: bool, alpha: float,
                act_elu: bool, p: float, masks: Optional[LayerMasks]):
        _require_cuda(x, "input features")
        dev = x.device
        n, f_in = x.shape
        if graph.n_dst != n or graph.n_src != n:
            raise RuntimeError(f"adjacency is {graph.n_dst}x{graph.n_src} but the input has {n} rows")
        HD = H * Dp
        M_out = HD * (2 if has_skip else 1)
        assert w_ext.shape == (f_in, M_out) and a_src.shape == (H, Dp) and a_dst.shape == (H, Dp)
        x = x.contiguous()
        w_ext = w_ext.contiguous()
        a_src = a_src.contiguous()
        a_dst = a_dst.contiguous()
        masks = masks if (masks is not None and p > 0.0) else LayerMasks()
        inv_keep = 1.0 / (1.0 - p) if p > 0.0 else 1.0
        st = _stream()
        seeded = masks.seed is not None
        seed = masks.seed if seeded else 0
        o_in, o_wh, o_att = masks.offsets if seeded else (0, 0, 0)
        pk = float(p) if seeded else 0.0   # p handed to the kernels: > 0 only in seeded mode

        # ---- K1: projection (+ skip columns) ---------------------------------------------
        z = _mem.empty(n, M_out, dtype=torch.float32, device=dev)
        if seeded:   # every head drops the input with its own decisions (layers.py:34,132), all heads in one launch
            _lib.call("gatk_gemm_heads_dropout_fwd", n, f_in, H, Dp, int(has_skip), x.data_ptr(), f_in, w_ext.data_ptr(), M_out,
                      z.data_ptr(), M_out, seed, o_in, pk, st)
        elif masks.keep_in is None:
            _gemm(0, 0, n, M_out, f_in, x, f_in, w_ext, M_out, z, M_out)
        else:
            xh = _mem.empty_like(x)
            for h in range(H):  # every head drops the input with its own mask (layers.py:34,132)
                _lib.call("gatk_mask_scale", x.data_ptr(), f_in, masks.keep_in[h].data_ptr(), inv_keep,
                          xh.data_ptr(), f_in, n, f_in, st)
                for base in ((0, HD) if has_skip else (0,)):
                    off = base + h * Dp
                    _gemm(0, 0, n, Dp, f_in, xh, f_in, w_ext, M_out, z, M_out, b_off=off, c_off=off)
        wh_ptr = z.data_ptr()
        skip_ptr = z.data_ptr() + 4 * HD if has_skip else None

        # ---- logits (and the post-projection dropout, in place) --------------------------
        f = _mem.empty(n, H, dtype=torch.float32, device=dev)
        g = _mem.empty(n, H, dtype=torch.float32, device=dev)
        _lib.call("gatk_logits_fwd", n, H, Dp, wh_ptr, M_out, _ptr(masks.keep_wh), inv_keep,
                  a_src.data_ptr(), a_dst.data_ptr(), f.data_ptr(), g.data_ptr(), seed, o_wh, pk, st)

        # ---- K2: fused attention -----------------------------------------------------------
        need_grad = any(ctx.needs_input_grad[:4])
        out = _mem.empty(n, HD, dtype=torch.float32, device=dev)
        separate_hagg = need_grad and (has_skip or act_elu)
        hagg = _mem.empty(n, HD, dtype=torch.float32, device=dev) if separate_hagg else None
        lse = _mem.empty(n, H, dtype=torch.float32, device=dev) if need_grad else None
        hubs = graph.hubs
        scratch = _hub_scratch(0, H, Dp, hubs.n_seg, dev)
        _lib.call("gatk_attn_fwd", n, graph.rowptr.data_ptr(), _ptr(graph.col), H, Dp, wh_ptr, M_out,
                  f.data_ptr(), H, g.data_ptr(), H, _ptr(masks.keep_att), inv_keep, float(alpha),
                  skip_ptr, M_out, int(act_elu), _ptr(hagg), out.data_ptr(), HD, _ptr(lse),
                  *hubs.args(scratch), graph.counter.data_ptr(), *hubs.item_args(), seed, o_att, pk, st)

        if need_grad:
            ctx.graph, ctx.masks = graph, masks
            ctx.cfg = (H, Dp, has_skip, float(alpha), bool(act_elu), float(p), inv_keep)
            ctx.save_for_backward(x, w_ext, a_src, a_dst, z, f, g, lse, out, hagg if separate_hagg else out)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, w_ext, a_src, a_dst, z, f, g, lse, out, hagg = ctx.saved_tensors
        graph, masks = ctx.graph, ctx.masks
        H, Dp, has_skip, alpha, act_elu, p, inv_keep = ctx.cfg
        dev = x.device
        n, f_in = x.shape
        HD = H * Dp
        M_out = HD * (2 if has_skip else 1)
        st = _stream()
        gout = gout.contiguous()
        tptr, trow, perm, thubs = graph.transpose()
        seeded = masks.seed is not None
        seed = masks.seed if seeded else 0
        o_in, o_wh, o_att = masks.offsets if seeded else (0, 0, 0)
        pk = float(p) if seeded else 0.0

        # dZ = [dWh | dSkip]; with a skip projection dL/dh' IS dSkip, so prep writes it there as well.
        dz_rows = _mem.empty(n, M_out, dtype=torch.float32, device=dev)
        ldrec = _lib.query("gatk_attn_bwd_record_ld", H, Dp)
        rec = _mem.empty(n, ldrec, dtype=torch.float32, device=dev)
        df = _mem.empty(n, H, dtype=torch.float32, device=dev)
        dg = _mem.empty(n, H, dtype=torch.float32, device=dev)
        edge_dz = _mem.empty(graph.nnz, H, dtype=torch.float32, device=dev)

        # ---- K3 prep: per-destination records [dh' | f, lse, c] -------------------------------------
        _lib.call("gatk_attn_bwd_prep", n, H, Dp, gout.data_ptr(), HD, out.data_ptr() if (act_elu and has_skip) else None, HD,
                  int(act_elu), hagg.data_ptr(), HD, f.data_ptr(), H, lse.data_ptr(), rec.data_ptr(), ldrec,
                  dz_rows.data_ptr() + 4 * HD if has_skip else None, M_out, st)

        # ---- K4 fused source pass over the transposed pattern (one gather of the record per edge) -----
        scratch_t = _hub_scratch(1, H, Dp, thubs.n_seg, dev)
        _lib.call("gatk_attn_bwd_fused", n, tptr.data_ptr(), _ptr(trow), _ptr(perm), H, Dp, z.data_ptr(), M_out,
                  g.data_ptr(), H, rec.data_ptr(), ldrec, _ptr(masks.keep_att), inv_keep, alpha,
                  a_dst.data_ptr(), dz_rows.data_ptr(), M_out, dg.data_ptr(), H, edge_dz.data_ptr(), None, 0,
                  *thubs.args(scratch_t), graph.counter.data_ptr(), *thubs.item_args(), seed, o_att, pk, st)
        del rec

        # ---- finish: df = segmented sum of dz, dWh += df a_src, Wh-dropout mask ------------------
        hubs = graph.hubs
        scratch = _hub_scratch(2, H, Dp, hubs.n_seg, dev)
        _lib.call("gatk_attn_bwd_finish", n, graph.rowptr.data_ptr(), H, Dp, edge_dz.data_ptr(), a_src.data_ptr(),
                  _ptr(masks.keep_wh), inv_keep, dz_rows.data_ptr(), M_out, df.data_ptr(), H,
                  *hubs.args(scratch), seed, o_wh, pk, st)
        del edge_dz

        # ---- da ------------------------------------------------------------------------------
        da_src = _mem.empty(H, Dp, dtype=torch.float32, device=dev)
        da_dst = _mem.empty(H, Dp, dtype=torch.float32, device=dev)
        ws = _mem.empty(_lib.query("gatk_da_workspace_floats", H, Dp), dtype=torch.float32, device=dev)
        _lib.call("gatk_da_reduce", n, H, Dp, z.data_ptr(), M_out, df.data_ptr(), dg.data_ptr(),
                  da_src.data_ptr(), da_dst.data_ptr(), ws.data_ptr(), st)

        # ---- K5: projection backward -------------------------------------------------------
        need_dx = ctx.needs_input_grad[0]
        dw_ext = _mem.empty(f_in, M_out, dtype=torch.float32, device=dev)
        dx = _mem.empty(n, f_in, dtype=torch.float32, device=dev) if need_dx else None
        if seeded:   # the forward's per-head input decisions, re-evaluated inside one launch per product
            ws2 = _mem.empty(_lib.query("gatk_gemm_heads_dropout_ws_floats", n, f_in, H, Dp, int(has_skip)),
                             dtype=torch.float32, device=dev)
            _lib.call("gatk_gemm_heads_dropout_dw", n, f_in, H, Dp, int(has_skip), x.data_ptr(), f_in, dz_rows.data_ptr(), M_out,
                      dw_ext.data_ptr(), M_out, ws2.data_ptr(), seed, o_in, pk, st)
            if need_dx:
                _lib.call("gatk_gemm_heads_dropout_dx", n, f_in, H, Dp, int(has_skip), dz_rows.data_ptr(), M_out,
                          w_ext.data_ptr(), M_out, dx.data_ptr(), f_in, seed, o_in, pk, st)
        elif masks.keep_in is None:
            _gemm(1, 0, f_in, M_out, n, x, f_in, dz_rows, M_out, dw_ext, M_out)
            if need_dx:
                _gemm(0, 1, n, f_in, M_out, dz_rows, M_out, w_ext, M_out, dx, f_in)
        else:
            xh = _mem.empty_like(x)
            dxh = _mem.empty_like(x) if need_dx else None
            if need_dx:
                dx.zero_()
            for h in range(H):
                _lib.call("gatk_mask_scale", x.data_ptr(), f_in, masks.keep_in[h].data_ptr(), inv_keep,
                          xh.data_ptr(), f_in, n, f_in, st)
                for k, base in enumerate((0, HD) if has_skip else (0,)):
                    off = base + h * Dp
                    _gemm(1, 0, f_in, Dp, n, xh, f_in, dz_rows, M_out, dw_ext, M_out, b_off=off, c_off=off)
                    if need_dx:
                        _gemm(0, 1, n, f_in, Dp, dz_rows, M_out, w_ext, M_out, dxh, f_in, accumulate=int(k > 0),
                              a_off=off, b_off=off)
                if need_dx:
                    _lib.call("gatk_mask_scale", dxh.data_ptr(), f_in, masks.keep_in[h].data_ptr(), inv_keep,
                              dxh.data_ptr(), f_in, n, f_in, st)
                    dx.add_(dxh)
        return dx, dw_ext, da_src, da_dst, None, None, None, None, None, None, None, None


# Small graphs with wide rows (PPI: ~4.5 k nodes, 4 x 256 / 6 x 128 floats per row) are latency bound in the attention
# kernels: a warp owns a whole H*Dp row, rows that wide need 8+ float4 accumulators per lane, which limits an SM to four
# resident warps -- a few hundred warps chasing dependent gathers on a 148-SM part.  Below SMALL_GRAPH_ENTRIES stored
# entries such layers are run one head group at a time (rows of <= 64 slots: 2 accumulators per lane, 16 resident warps
# per SM and H/Hc times more work items).  MEASURED (PPI epoch, ncu launch list, 20 steps): once the work items are
# sized for small patterns (graph.HubPartition.item_edges) the unsplit kernels take 13.9 ms and the split ones 15.3 ms,
# and at the products scale the split is slower as well (bandwidth bound: narrower gathers lose efficiency) -- so the
# split is OFF by default (GATK_SMALL_GRAPH_ENTRIES=<entries> turns it on below that many stored entries).
SMALL_GRAPH_ENTRIES = int(os.environ.get("GATK_SMALL_GRAPH_ENTRIES", "0"))


def small_graph_head_chunk(H: int, Dp: int, nnz: int) -> int:
    """Heads per attention launch (H = no split)."""
    if nnz >= SMALL_GRAPH_ENTRIES or H * Dp <= 512:
        return H
    hc = max(1, 256 // Dp)
    while H % hc:
        hc -= 1
    return hc


class GatLayerFoldedFunction(torch.autograd.Function):
    """The layer when no dropout sits between projection and logits (eval, or p == 0 as in train_ppi.py:49
    and the benchmark shapes).  The attention-logit halves are linear in the input,
    f = (x W) a_src = x (W a_src), so they ride in the projection GEMM as 2H extra output columns and
    their gradients df, dg ride in its backward as 2H extra columns of dZ:

        Z  = x [W | S | W a_src | W a_dst]          (one GEMM;  w_full is built from the per-head parameters
        dW_full = x^T [dWh | dSkip | df | dg]        by parameter-sized torch ops, so autograd routes
        dx = dZ w_full^T                              dW_full back into dW, da_src, da_dst)

    which removes the logits pass, the da reduction over the nodes and the read-modify-write of dWh for
    the df a_src / dg a_dst terms.  Same math as GatLayerFunction up to fp32 re-association.
    The attention kernels run on all heads at once, or -- small graphs with wide rows, see
    small_graph_head_chunk -- on head groups of Hc heads (column blocks of the same buffers)."""

    @staticmethod
    def forward(ctx, x, w_full, graph: Graph, H: int, Dp: int, has_skip: bool, alpha: float, act_elu: bool):
        _require_cuda(x, "input features")
        dev = x.device
        n, f_in = x.shape
        if graph.n_dst != n or graph.n_src != n:
            raise RuntimeError(f"adjacency is {graph.n_dst}x{graph.n_src} but the input has {n} rows")
        HD = H * Dp
        M_out = HD * (2 if has_skip else 1)
        Mz = w_full.shape[1]
        assert Mz >= M_out + 2 * H and Mz % 4 == 0
        x = x.contiguous()
        w_full = w_full.contiguous()
        st = _stream()
        z = _mem.empty(n, Mz, dtype=torch.float32, device=dev)
        _gemm(0, 0, n, Mz, f_in, x, f_in, w_full, Mz, z, Mz, label="gemm:project_folded")
        zp = z.data_ptr()
        need_grad = any(ctx.needs_input_grad[:2])
        out = _mem.empty(n, HD, dtype=torch.float32, device=dev)
        separate_hagg = need_grad and (has_skip or act_elu)
        Hc = small_graph_head_chunk(H, Dp, graph.nnz)
        C, HDc = H // Hc, Hc * Dp
        # per head group: hagg [C, n, Hc*Dp] and lse [C, n, Hc] (C == 1: the plain [n, H*Dp] / [n, H] layouts)
        hagg = _mem.empty(C, n, HDc, dtype=torch.float32, device=dev) if separate_hagg else None
        lse = _mem.empty(C, n, Hc, dtype=torch.float32, device=dev) if need_grad else None
        hubs = graph.hubs
        scratch = _hub_scratch(0, Hc, Dp, hubs.n_seg, dev)
        for c in range(C):
            _lib.call("gatk_attn_fwd", n, graph.rowptr.data_ptr(), _ptr(graph.col), Hc, Dp, zp + 4 * c * HDc, Mz,
                      zp + 4 * (M_out + c * Hc), Mz, zp + 4 * (M_out + H + c * Hc), Mz, None, 1.0, float(alpha),
                      zp + 4 * (HD + c * HDc) if has_skip else None, Mz, int(act_elu),
                      hagg[c].data_ptr() if separate_hagg else None, out.data_ptr() + 4 * c * HDc, HD,
                      lse[c].data_ptr() if need_grad else None, *hubs.args(scratch), graph.counter.data_ptr(),
                      *hubs.item_args(), 0, 0, 0.0, st)
        if need_grad:
            ctx.graph = graph
            ctx.cfg = (H, Dp, Hc, has_skip, float(alpha), bool(act_elu), separate_hagg)
            ctx.save_for_backward(x, w_full, z, lse, out, hagg if separate_hagg else out)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, w_full, z, lse, out, hagg = ctx.saved_tensors
        graph = ctx.graph
        H, Dp, Hc, has_skip, alpha, act_elu, separate_hagg = ctx.cfg
        dev = x.device
        n, f_in = x.shape
        HD = H * Dp
        C, HDc = H // Hc, Hc * Dp
        M_out = HD * (2 if has_skip else 1)
        Mz = w_full.shape[1]
        st = _stream()
        gout = gout.contiguous()
        tptr, trow, perm, thubs = graph.transpose()[:4]
        zp = z.data_ptr()

        dz_rows = _mem.empty(n, Mz, dtype=torch.float32, device=dev)   # [dWh | dSkip | df | dg | pad]
        if Mz > M_out + 2 * H:
            dz_rows[:, M_out + 2 * H:].zero_()
        dzp = dz_rows.data_ptr()
        ldrec = _lib.query("gatk_attn_bwd_record_ld", Hc, Dp)
        rec = _mem.empty(n, ldrec, dtype=torch.float32, device=dev)
        edge_dz = _mem.empty(graph.nnz, Hc, dtype=torch.float32, device=dev)
        hubs = graph.hubs
        scratch_t = _hub_scratch(1, Hc, Dp, thubs.n_seg, dev)
        scratch = _hub_scratch(2, Hc, Dp, hubs.n_seg, dev)
        for c in range(C):   # stream order lets the head groups share rec / edge_dz / the hub scratch
            # without a separate hagg (no skip, no ELU) the layer output IS the aggregation
            hp, ldh = (hagg[c].data_ptr(), HDc) if separate_hagg else (out.data_ptr() + 4 * c * HDc, HD)
            _lib.call("gatk_attn_bwd_prep", n, Hc, Dp, gout.data_ptr() + 4 * c * HDc, HD,
                      out.data_ptr() + 4 * c * HDc if (act_elu and has_skip) else None, HD, int(act_elu), hp, ldh,
                      zp + 4 * (M_out + c * Hc), Mz, lse[c].data_ptr(), rec.data_ptr(), ldrec,
                      dzp + 4 * (HD + c * HDc) if has_skip else None, Mz, st)
            _lib.call("gatk_attn_bwd_fused", n, tptr.data_ptr(), _ptr(trow), _ptr(perm), Hc, Dp, zp + 4 * c * HDc, Mz,
                      zp + 4 * (M_out + H + c * Hc), Mz, rec.data_ptr(), ldrec, None, 1.0, alpha,
                      None, dzp + 4 * c * HDc, Mz, dzp + 4 * (M_out + H + c * Hc), Mz, edge_dz.data_ptr(), None, 0,
                      *thubs.args(scratch_t), graph.counter.data_ptr(), *thubs.item_args(), 0, 0, 0.0, st)
            _lib.call("gatk_attn_bwd_finish", n, graph.rowptr.data_ptr(), Hc, Dp, edge_dz.data_ptr(), None,
                      None, 1.0, None, 0, dzp + 4 * (M_out + c * Hc), Mz, *hubs.args(scratch), 0, 0, 0.0, st)
        del rec, edge_dz
        dw_full = _mem.empty(f_in, Mz, dtype=torch.float32, device=dev)
        _gemm(1, 0, f_in, Mz, n, x, f_in, dz_rows, Mz, dw_full, Mz, label="gemm:dW_folded")
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _mem.empty(n, f_in, dtype=torch.float32, device=dev)
            _gemm(0, 1, n, f_in, Mz, dz_rows, Mz, w_full, Mz, dx, f_in, label="gemm:dx_folded")
        return dx, dw_full, None, None, None, None, None, None


def agg_first_geometry(f_in: int, H: int, Dp: int):
    """(Fp, ok): padded input width and whether the aggregate-first form applies and pays: at most 8
    heads, accumulators fit the register file, and the input row is narrower than the projected row."""
    Fp = (f_in + 3) // 4 * 4
    hp = 1 if H <= 1 else 2 if H <= 2 else 4 if H <= 4 else 8
    ns = 1 if Fp <= 128 else 2 if Fp <= 256 else 4
    ok = H <= 8 and Fp <= 512 and hp * ns <= 16 and Fp < H * Dp
    return Fp, ok


class GatLayerAggFirstFunction(torch.autograd.Function):
    """The layer with the neighbour sum taken BEFORE the projection (see csrc/attn_x.cu):

        f, xg = x [W a_src], [x | x (W a_dst)]            gatk_logits_pack: logits are linear in the input; the
                                                          source half g rides behind the input row it is gathered with
        xagg  = softmax-weighted neighbour sum of x       gatk_attn_x_fwd, gathers F_in-wide rows
        out   = act( xagg_h W_h  (+ x S_h) )              H per-head GEMMs on the aggregated rows

    Same math as layers.py:134-170 up to fp32 re-association (sum_j alpha_ij (x_j W) = (sum_j alpha_ij x_j) W).
    Used when no dropout sits between projection and logits, the input needs no gradient (a first layer)
    and F_in < H*D: the edge passes then move F_in/(H*D) of the bytes (1/5 at the products shape)."""

    @staticmethod
    def forward(ctx, x, w_ext, w_uv, graph: Graph, H: int, Dp: int, has_skip: bool, alpha: float, act_elu: bool):
        _require_cuda(x, "input features")
        dev = x.device
        n, f_in = x.shape
        if graph.n_dst != n or graph.n_src != n:
            raise RuntimeError(f"adjacency is {graph.n_dst}x{graph.n_src} but the input has {n} rows")
        HD = H * Dp
        M_out = HD * (2 if has_skip else 1)
        Muv = w_uv.shape[1]
        Fp = (f_in + 3) // 4 * 4
        assert w_ext.shape == (f_in, M_out) and Muv >= 2 * H and Muv % 4 == 0
        x = x.contiguous()
        w_ext = w_ext.contiguous()
        w_uv = w_uv.contiguous()
        st = _stream()
        P = _lib.query("gatk_xg_pitch", Fp, H)
        xg = _mem.empty(n, P, dtype=torch.float32, device=dev)
        f = _mem.empty(n, H, dtype=torch.float32, device=dev)
        _lib.call("gatk_logits_pack", n, f_in, H, x.data_ptr(), f_in, w_uv.data_ptr(), Muv, xg.data_ptr(), P,
                  f.data_ptr(), H, st)
        need_grad = any(ctx.needs_input_grad[1:3])
        xagg = _mem.empty(n, H * Fp, dtype=torch.float32, device=dev)
        lse = _mem.empty(n, H, dtype=torch.float32, device=dev) if need_grad else None
        hubs = graph.hubs
        scratch = _x_scratch(0, H, Fp, hubs.n_seg, dev)
        _lib.call("gatk_attn_x_fwd", graph.n_src, n, graph.rowptr.data_ptr(), _ptr(graph.col), H, Fp, xg.data_ptr(), P,
                  f.data_ptr(), H, float(alpha), xagg.data_ptr(), H * Fp, _ptr(lse),
                  *hubs.args(scratch), graph.counter.data_ptr(), *hubs.item_args(), st)
        out = _mem.empty(n, HD, dtype=torch.float32, device=dev)
        fuse_elu = act_elu and not has_skip  # ELU rides in the projection's epilogue unless a skip term is added first
        _gemm_batched(0, 0, n, Dp, f_in, H, xagg, H * Fp, Fp, w_ext, M_out, Dp, out, HD, Dp, epilogue=int(fuse_elu),
                      label="gemm:project")
        if has_skip:
            _gemm(0, 0, n, HD, f_in, xg, P, w_ext, M_out, out, HD, accumulate=1, b_off=HD, label="gemm:skip")
            if act_elu:
                _lib.call("gatk_elu_fwd", n, HD, out.data_ptr(), HD, st)
        if need_grad:
            ctx.graph = graph
            ctx.cfg = (H, Dp, has_skip, float(alpha), bool(act_elu), f_in, Fp, Muv)
            ctx.save_for_backward(xg, w_ext, f, lse, xagg, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        xg, w_ext, f, lse, xagg, out = ctx.saved_tensors
        graph = ctx.graph
        H, Dp, has_skip, alpha, act_elu, f_in, Fp, Muv = ctx.cfg
        dev = xg.device
        n, P = xg.shape
        HD = H * Dp
        M_out = HD * (2 if has_skip else 1)
        st = _stream()
        gout = gout.contiguous()
        # dh' = gout * ELU'(out): formed inside the two products that consume it when both run on the batched
        # tensor-core kernels (no separate pass, dh' never written); a skip projection needs dh' for a third product
        fuse = act_elu and not has_skip and FUSE_ELU_GRAD and _elu_grad_fusable(n, f_in, H, Dp, Fp, M_out, HD)
        if act_elu and not fuse:
            dhp = _mem.empty(n, HD, dtype=torch.float32, device=dev)
            _lib.call("gatk_elu_bwd", n, HD, gout.data_ptr(), HD, out.data_ptr(), HD, dhp.data_ptr(), HD, st)
        else:
            dhp = gout
        eo, ldeo = (out, HD) if fuse else (None, 0)
        # value path: dW_h = xagg_h^T dh'_h, dS = x^T dh';  dxagg_h = dh'_h W_h^T feeds the softmax backward
        dw_ext = _mem.empty(f_in, M_out, dtype=torch.float32, device=dev)
        dxagg = (_mem.empty if Fp == f_in else torch.zeros)(n, H * Fp, dtype=torch.float32, device=dev)
        _gemm_batched(1, 0, f_in, Dp, n, H, xagg, H * Fp, Fp, dhp, HD, Dp, dw_ext, M_out, Dp, label="gemm:dW",
                      elu_out=eo, ld_elu=ldeo)
        _gemm_batched(0, 1, n, f_in, Dp, H, dhp, HD, Dp, w_ext, M_out, Dp, dxagg, H * Fp, Fp, label="gemm:dxagg",
                      elu_out=eo, ld_elu=ldeo)
        if has_skip:
            _gemm(1, 0, f_in, HD, n, xg, P, dhp, HD, dw_ext, M_out, c_off=HD, label="gemm:dskip")
        # logit path: ds per stored entry, df per destination, dg per source (transposed sum of ds)
        hubs = graph.hubs
        scratch = _x_scratch(1, H, Fp, hubs.n_seg, dev)
        if DETERMINISTIC:
            ds = _mem.empty(graph.nnz, H, dtype=torch.float32, device=dev)
            dfg = (_mem.empty if Muv == 2 * H else torch.zeros)(n, Muv, dtype=torch.float32, device=dev)
            _lib.call("gatk_attn_x_bwd", graph.n_src, n, graph.rowptr.data_ptr(), _ptr(graph.col), H, Fp, xg.data_ptr(), P,
                      f.data_ptr(), H, lse.data_ptr(), alpha, xagg.data_ptr(), H * Fp,
                      dxagg.data_ptr(), H * Fp, ds.data_ptr(), None, None, 0, dfg.data_ptr(), Muv,
                      *hubs.args(scratch), graph.counter.data_ptr(), *hubs.item_args(), st)
            tptr, _trow, perm, thubs = graph.transpose()[:4]
            _lib.call("gatk_edge_tsum", graph.n_src, tptr.data_ptr(), _ptr(perm), H, ds.data_ptr(),
                      dfg.data_ptr() + 4 * H, Muv, thubs.seg_len, _ptr(thubs.rows), thubs.n_hub, st)
        else:
            # dg_j is accumulated by the edge pass itself: vector reductions into a COMPACT zeroed [N, H] array (half the
            # footprint of the [df | dg] rows, so it fits the persisting L2 window the kernel sets up), merged afterwards
            dfg = _mem.zeros(n, Muv, dtype=torch.float32, device=dev)
            dg = torch.zeros(graph.n_src, H, dtype=torch.float32, device=dev)
            _lib.call("gatk_attn_x_bwd", graph.n_src, n, graph.rowptr.data_ptr(), _ptr(graph.col), H, Fp, xg.data_ptr(), P,
                      f.data_ptr(), H, lse.data_ptr(), alpha, xagg.data_ptr(), H * Fp,
                      dxagg.data_ptr(), H * Fp, None, None, dg.data_ptr(), H, dfg.data_ptr(), Muv,
                      *hubs.args(scratch), graph.counter.data_ptr(), *hubs.item_args(), st)
            dfg[:, H:2 * H] = dg
        dw_uv = _mem.empty(f_in, Muv, dtype=torch.float32, device=dev)
        _gemm_batched(1, 0, f_in, Muv, n, 1, xg, P, 0, dfg, Muv, 0, dw_uv, Muv, 0, label="gemm:dlogits")  # one "head": the TMEM-A TN kernel
        return None, dw_ext, dw_uv, None, None, None, None, None, None


def _x_scratch(which: int, H: int, Fp: int, n_seg: int, dev):
    if not n_seg:
        return None
    return _mem.empty(_lib.query("gatk_attn_x_scratch_floats", which, H, Fp, n_seg), dtype=torch.float32, device=dev)


class HeadCombineFunction(torch.autograd.Function):
    """[N, H*Dp] -> torch.cat of the unpadded heads (mode 0, models.py:32) or their mean
    (mode 1, models.py:34)."""

    @staticmethod
    def forward(ctx, rows, H: int, D: int, Dp: int, mode: int):
        rows = rows.contiguous()
        n = rows.shape[0]
        out = _mem.empty(n, H * D if mode == 0 else D, dtype=torch.float32, device=rows.device)
        _lib.call("gatk_head_combine", n, H, D, Dp, rows.data_ptr(), H * Dp, mode, out.data_ptr(), _stream())
        ctx.cfg = (n, H, D, Dp, mode)
        return out

    @staticmethod
    def backward(ctx, gout):
        n, H, D, Dp, mode = ctx.cfg
        gout = gout.contiguous()
        gin = _mem.empty(n, H * Dp, dtype=torch.float32, device=gout.device)
        _lib.call("gatk_head_combine_bwd", n, H, D, Dp, gout.data_ptr(), mode, gin.data_ptr(), H * Dp, _stream())
        return gin, None, None, None, None


def _pad_cols(t: torch.Tensor, Dp: int) -> torch.Tensor:
    d = t.shape[-1]
    return t if d == Dp else torch.nn.functional.pad(t, (0, Dp - d))


def pack_heads(Ws: Sequence[torch.Tensor], a_srcs: Sequence[torch.Tensor], a_dsts: Sequence[torch.Tensor],
               skips: Optional[Sequence[torch.Tensor]]):
    """Per-head parameters (the reference's nn.Parameters, models.py:27) -> packed operands.
    Plain torch ops on parameter-sized tensors, so autograd splits the packed gradients back; the heads are
    stacked first and padded once (a handful of launches per layer instead of a pad + copy per head: the PPI
    step spent a third of its launches here)."""
    D = Ws[0].shape[1]
    Dp = padded_width(D)
    f_in, H = Ws[0].shape[0], len(Ws)
    blocks = [torch.stack(list(Ws), dim=1)]                      # [F, H, D]
    if skips is not None:
        blocks.append(torch.stack(list(skips), dim=1))
    w3 = torch.cat(blocks, dim=1) if len(blocks) > 1 else blocks[0]   # [F, (2)H, D]
    w_ext = _pad_cols(w3, Dp).reshape(f_in, -1)                  # [F, (2)H*Dp]
    a2 = _pad_cols(torch.stack([a.reshape(-1) for a in list(a_srcs) + list(a_dsts)]), Dp)   # [2H, Dp]
    return w_ext, a2[:H], a2[H:], D, Dp


def gat_layer(x: torch.Tensor, graph: Graph, Ws, a_srcs, a_dsts, skips, alpha: float, concat: bool,
              p: float = 0.0, training: bool = False, masks: Optional[LayerMasks] = None,
              combine: str = "cat", form: str = "auto") -> torch.Tensor:
    """All heads of one GAT layer.  concat=True applies ELU inside each head (layers.py:50-53);
    combine="cat" -> [N, H*D] (models.py:32), "mean" -> [N, D] (models.py:34),
    "none" -> the padded [N, H*Dp] rows.
    form: "auto" picks among the three algebraically equal forms of the layer -- "explicit" (projection,
    logits pass, attention; the only one valid with dropout), "folded" (logits as projection columns) and
    "agg_first" (neighbour sum before the projection; narrow inputs that need no gradient)."""
    _require_cuda(x, "input features")
    if graph.device != x.device:
        raise RuntimeError(f"pygat_b200: the graph lives on {graph.device} but the input features are on {x.device}")
    with on_device(x):
        return _gat_layer(x, graph, Ws, a_srcs, a_dsts, skips, alpha, concat, p, training, masks, combine, form)


def _gat_layer(x, graph, Ws, a_srcs, a_dsts, skips, alpha, concat, p, training, masks, combine, form):
    H = len(Ws)
    if x.dtype != torch.float32:
        x = x.float()
    w_ext, a_src, a_dst, D, Dp = pack_heads(Ws, a_srcs, a_dsts, skips)
    p_eff = float(p) if training else 0.0
    if p_eff > 0.0 and masks is None:
        masks = seeded_masks(x.shape[0], x.shape[1], H, Dp, graph.nnz)   # nothing materialised: decisions live in the kernels
    use_agg_first = (p_eff == 0.0 and FOLD_LOGITS and AGG_FIRST and form in ("auto", "agg_first")
                     and not (x.requires_grad and torch.is_grad_enabled())
                     and (agg_first_geometry(x.shape[1], H, Dp)[1] or form == "agg_first"))
    if p_eff == 0.0 and FOLD_LOGITS and form != "explicit":
        # f = x (W a_src), g = x (W a_dst): linear in the input (see GatLayerFoldedFunction)
        w3 = w_ext[:, : H * Dp].reshape(x.shape[1], H, Dp)
        uv = [(w3 * a_src).sum(-1), (w3 * a_dst).sum(-1)]
        pad = (-2 * H) % 4
        if pad:
            uv.append(w_ext.new_zeros(x.shape[1], pad))
        if use_agg_first:
            rows = GatLayerAggFirstFunction.apply(x, w_ext, torch.cat(uv, dim=1), graph, H, Dp, skips is not None,
                                                  float(alpha), bool(concat))
        else:
            rows = GatLayerFoldedFunction.apply(x, torch.cat([w_ext] + uv, dim=1), graph, H, Dp, skips is not None,
                                                float(alpha), bool(concat))
    else:
        rows = GatLayerFunction.apply(x, w_ext, a_src, a_dst, graph, H, Dp, skips is not None, float(alpha),
                                      bool(concat), p_eff, masks)
    if combine == "none":
        return rows
    if combine == "mean":
        return HeadCombineFunction.apply(rows, H, D, Dp, 1)
    if D == Dp:
        return rows
    return HeadCombineFunction.apply(rows, H, D, Dp, 0)


def pack_masks(keep_in: Optional[List[torch.Tensor]], keep_wh: Optional[List[torch.Tensor]],
               keep_att: Optional[List[torch.Tensor]], Dp: int) -> LayerMasks:
    """Per-head boolean masks (as a test draws them for the oracle) -> the packed uint8 layout."""
    m = LayerMasks()
    if keep_in is not None:
        m.keep_in = torch.stack([k.to(torch.uint8) for k in keep_in]).contiguous()
    if keep_wh is not None:
        m.keep_wh = torch.cat([_pad_cols(k.to(torch.uint8), Dp) for k in keep_wh], dim=1).contiguous()
    if keep_att is not None:
        m.keep_att = torch.stack([k.to(torch.uint8) for k in keep_att], dim=1).contiguous()
    return m


# ====================================================================== GATv2 flavour (layers.py:234-316)
class EngineMatmul(torch.autograd.Function):
    """x @ w through gatk_gemm (and its autograd: dW = x^T dZ, dx = dZ w^T)."""

    @staticmethod
    def forward(ctx, x, w):
        _require_cuda(x, "input features")
        x, w = x.contiguous().float(), w.contiguous().float()
        n, k = x.shape
        m = w.shape[1]
        z = _mem.empty(n, m, dtype=torch.float32, device=x.device)
        _gemm(0, 0, n, m, k, x, k, w, m, z, m)
        ctx.save_for_backward(x, w)
        return z

    @staticmethod
    def backward(ctx, dz):
        x, w = ctx.saved_tensors
        dz = dz.contiguous()
        n, k = x.shape
        m = w.shape[1]
        dw = dx = None
        if ctx.needs_input_grad[1]:
            dw = _mem.empty(k, m, dtype=torch.float32, device=x.device)
            _gemm(1, 0, k, m, n, x, k, dz, m, dw, m)
        if ctx.needs_input_grad[0]:
            dx = _mem.empty(n, k, dtype=torch.float32, device=x.device)
            _gemm(0, 1, n, k, m, dz, m, w, m, dx, k)
        return dx, dw


class GatV2AttnFunction(torch.autograd.Function):
    """out[N, H*Dp] = act( sum_j softmax_j(a . LeakyReLU(Whi_i + Whj_j)) Whi_j (+ skip_i) ) from z = [Whi | Whj | (skip)]:
    the fused attention of SpGraphAttentionLayerV2 (layers.py:275-305), csrc/attn_v2.cu."""

    @staticmethod
    def forward(ctx, z, a, graph: Graph, H: int, Dp: int, has_skip: bool, alpha: float, act_elu: bool, keep_att, inv_keep: float):
        _require_cuda(z, "projected features")
        dev = z.device
        z, a = z.contiguous(), a.contiguous()
        n, ldz = z.shape
        if graph.n_dst != n or graph.n_src != n:
            raise RuntimeError(f"adjacency is {graph.n_dst}x{graph.n_src} but the input has {n} rows")
        HD = H * Dp
        assert ldz == HD * (3 if has_skip else 2) and a.shape == (H, Dp)
        need_grad = any(ctx.needs_input_grad[:2])
        out = _mem.empty(n, HD, dtype=torch.float32, device=dev)
        hagg = _mem.empty(n, HD, dtype=torch.float32, device=dev) if need_grad else None
        lse = _mem.empty(n, H, dtype=torch.float32, device=dev) if need_grad else None
        _lib.call("gatk_attn_v2_fwd", n, graph.rowptr.data_ptr(), _ptr(graph.col), H, Dp, z.data_ptr(), ldz, a.data_ptr(),
                  _ptr(keep_att), float(inv_keep), float(alpha), int(has_skip), int(act_elu), _ptr(hagg), out.data_ptr(), HD,
                  _ptr(lse), graph.counter.data_ptr(), _stream())
        if need_grad:
            ctx.graph, ctx.keep = graph, keep_att
            ctx.cfg = (H, Dp, has_skip, float(alpha), bool(act_elu), float(inv_keep))
            ctx.save_for_backward(z, a, hagg, out, lse)
        return out

    @staticmethod
    def backward(ctx, gout):
        z, a, hagg, out, lse = ctx.saved_tensors
        H, Dp, has_skip, alpha, act_elu, inv_keep = ctx.cfg
        graph = ctx.graph
        n, ldz = z.shape
        HD = H * Dp
        gout = gout.contiguous()
        dz = torch.zeros_like(z)
        da = torch.zeros_like(a)
        _lib.call("gatk_attn_v2_bwd", n, graph.rowptr.data_ptr(), _ptr(graph.col), H, Dp, z.data_ptr(), ldz, a.data_ptr(),
                  _ptr(ctx.keep), inv_keep, alpha, int(has_skip), int(act_elu), hagg.data_ptr(), out.data_ptr(), HD,
                  lse.data_ptr(), gout.data_ptr(), HD, dz.data_ptr(), da.data_ptr(), graph.counter.data_ptr(), _stream())
        return dz, da, None, None, None, None, None, None, None, None


def gat_v2_layer(x: torch.Tensor, graph: Graph, Ws, a_vecs, skips, alpha: float, concat: bool, p: float = 0.0,
                 training: bool = False, combine: str = "cat") -> torch.Tensor:
    """All heads of one SpGraphAttentionLayerV2 layer.  Ws: per head (2*F_in, D) with rows [:F_in] = the Whi projection and
    [F_in:] = Whj (layers.py:265-266); a_vecs: per head (D,).  Training-mode dropout follows the reference's sites: the
    input and both projections (layers.py:262-269) with torch's generator, the attention through a keep mask."""
    _require_cuda(x, "input features")
    with on_device(x):
        H = len(Ws)
        x = x.float()
        f_in = x.shape[1]
        D = Ws[0].shape[1]
        Dp = padded_width(D)
        p_eff = float(p) if training else 0.0
        left = [_pad_cols(w[:f_in], Dp) for w in Ws]
        right = [_pad_cols(w[f_in:], Dp) for w in Ws]
        cols = left + right + ([_pad_cols(s, Dp) for s in skips] if skips is not None else [])
        w_ext = torch.cat(cols, dim=1)
        a = torch.stack([_pad_cols(v.reshape(-1), Dp) for v in a_vecs])
        keep_att, inv_keep = None, 1.0
        if p_eff > 0.0:
            # every head of the reference drops the input with its own mask; a batched layer call shares one mask
            # between the heads it fuses (same distribution, different draws)
            h = torch.nn.functional.dropout(x, p_eff, training=True)
            z = EngineMatmul.apply(h, w_ext)
            HD = H * Dp
            z = torch.cat([torch.nn.functional.dropout(z[:, :2 * HD], p_eff, training=True), z[:, 2 * HD:]], dim=1)
            keep_att = (torch.rand(graph.nnz, H, device=x.device) >= p_eff).to(torch.uint8)
            inv_keep = 1.0 / (1.0 - p_eff)
        else:
            z = EngineMatmul.apply(x, w_ext)
        rows = GatV2AttnFunction.apply(z, a, graph, H, Dp, skips is not None, float(alpha), bool(concat), keep_att, inv_keep)
        if combine == "none":
            return rows
        if combine == "mean":
            return HeadCombineFunction.apply(rows, H, D, Dp, 1)
        return rows if D == Dp else HeadCombineFunction.apply(rows, H, D, Dp, 0)
