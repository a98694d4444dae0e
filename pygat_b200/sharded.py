"""Multi-GPU execution of the GAT layer; one process per GPU, NCCL through
torch.distributed.  The reference has no distributed code -- everything here is new design.

* One big graph (products / papers shapes): 1-D DESTINATION-ROW sharding, contiguous row ranges
  balanced by stored entries.  Rank r owns rows [b_r, b_{r+1}): its slice of the input features, its
  CSR rows (column ids stay global) and its outputs.  Softmax and aggregation of a destination row
  need only that row's edges plus Wh_j / g_j of its sources, so the one exchange per layer and
  direction is: forward an ALL-GATHER of the projected rows [Wh | g], backward a REDUCE-SCATTER of
  the partial dWh / dg rows; the parameter gradients (dW, da) are all-reduced.
* Independent graphs (PPI): graph-level data parallelism, `allreduce_gradients` with node-count
  weights so the result equals the reference's merged-batch mean loss (train_ppi.py:114-119).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, _mem
from . import functional as Fn
from .functional import (HeadCombineFunction, LayerMasks, _gemm, _gemm_batched, _hub_scratch, _x_scratch,
                         agg_first_geometry, pack_heads)
from .graph import Graph, _ptr, _stream
from .synth import shard_rows_by_nnz


class ShardPlan:
    """Row boundaries of every rank (python ints), identical on all ranks."""

    def __init__(self, bounds: Sequence[int], rank: int, group=None):
        self.bounds = [int(b) for b in bounds]
        self.world = len(self.bounds) - 1
        self.rank = rank
        self.group = group
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = self.hi - self.lo
        self.n_total = self.bounds[-1]

    @staticmethod
    def by_nnz(rowptr: torch.Tensor, rank: int, world: int, group=None, row_cost: int = 0) -> "ShardPlan":
        """row_cost: per-row work expressed in stored entries (see synth.shard_rows_by_nnz)."""
        return ShardPlan(shard_rows_by_nnz(rowptr, world, row_cost), rank, group)

    def cached_xg(self, key, shape, dev):
        """(buffer, hit): the persistent gathered-rows buffer for the input identified by `key`; hit says its x
        columns are already complete on every rank (see ShardedGatLayerAggFirstFunction.forward)."""
        c = getattr(self, "_xg_cache", None)
        same_buf = c is not None and tuple(c[1].shape) == tuple(shape) and c[1].device == dev
        if same_buf and _same_input(c[0], key):
            return c[1], True
        buf = c[1] if same_buf else _mem.empty(*shape, dtype=torch.float32, device=dev)
        self._xg_cache = (key, buf)
        return buf, False

    def peer_rows(self, shape, dev):
        """PeerRows for an [N, P] gathered-row buffer in symmetric memory (every rank's copy mapped into every
        process), or None when symmetric memory is unavailable or switched off (GATK_PEER_PUSH=0) / the group is too
        large: the caller then uses the NCCL all-gather.  Collective: every rank of the group must call it."""
        c = getattr(self, "_peer_rows", None)
        if c is not None and (c is False or c.shape == tuple(shape)):
            return c or None
        ok, made = 1, None
        if os.environ.get("GATK_PEER_PUSH", "1") == "0" or self.world - 1 > MAX_PEERS or dev.type != "cuda":
            ok = 0
        else:
            try:
                made = PeerRows(self, tuple(shape), dev)
            except Exception as exc:  # no fabric / IPC support on this box
                import warnings
                warnings.warn(f"symmetric memory unavailable ({exc!r}); using NCCL all-gather")
                ok = 0
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self._peer_rows = made if int(flag.item()) == 1 else False
        return self._peer_rows or None

    def rows(self, full: torch.Tensor, r: Optional[int] = None) -> torch.Tensor:
        r = self.rank if r is None else r
        return full[self.bounds[r]:self.bounds[r + 1]]

    def local_graph(self, rowptr: torch.Tensor, col: torch.Tensor, seg_len: Optional[int] = None) -> Graph:
        """This rank's destination rows of the global CSR (global column ids kept)."""
        e0, e1 = int(rowptr[self.lo].item()), int(rowptr[self.hi].item())
        return Graph((rowptr[self.lo:self.hi + 1] - e0).contiguous(), col[e0:e1].contiguous(),
                     n_src=self.n_total, seg_len=seg_len)


class SourceShard:
    """This rank's SOURCE rows of the transposed pattern: tptr int64 [n_local + 1] (local), trow int32 [E_src]
    (GLOBAL destination ids), work items / hub segments over the source rows.  The backward of a hidden layer walks
    it (ShardedGatLayerWhFunction): every rank then completes dWh_j / dg_j of ITS OWN sources from the all-gathered
    destination records, instead of producing partial rows for all N sources (which does not shrink with the rank
    count) and reduce-scattering them."""

    def __init__(self, tptr: torch.Tensor, trow: torch.Tensor, seg_len: Optional[int] = None):
        from .graph import DEFAULT_SEG_LEN, HubPartition
        self.tptr, self.trow = tptr.contiguous(), trow.contiguous()
        self.n_src = tptr.numel() - 1
        self.nnz = int(trow.numel())
        self.hubs = HubPartition(self.tptr, int(seg_len or DEFAULT_SEG_LEN))
        self.counter = torch.zeros(1, dtype=torch.int32, device=tptr.device)


def _plan_source_shard(self, rowptr: torch.Tensor, col: torch.Tensor, seg_len: Optional[int] = None) -> SourceShard:
    """Source-row shard [lo, hi) of the GLOBAL pattern (setup work, like the CSR build: one transpose of the whole
    pattern on this device, then a slice)."""
    tptr_g, trow_g, _perm, _ = Graph(rowptr, col).transpose()
    e0, e1 = int(tptr_g[self.lo].item()), int(tptr_g[self.hi].item())
    return SourceShard((tptr_g[self.lo:self.hi + 1] - e0).contiguous(), trow_g[e0:e1].contiguous(), seg_len)


ShardPlan.source_shard = _plan_source_shard

MAX_PEERS = 15  # GATK_MAX_PEERS in include/gatk.h


def _input_key(x: torch.Tensor):
    """Identity of a layer input for the kept gathered rows: storage address, version counter, shape, and a weak
    reference to the tensor object (a dead object's storage may have been handed to a different tensor)."""
    import weakref
    return (x.data_ptr(), x._version, tuple(x.shape), weakref.ref(x))


def _same_input(kept, new) -> bool:
    return (kept is not None and new is not None and kept[:3] == new[:3] and kept[3]() is not None
            and (kept[3]() is new[3]() or kept[3]().data_ptr() == new[0]))


class PeerRows:
    """The gathered rows [N, P] of a shard group in symmetric memory: `buf` is this rank's copy, `ptrs` the other
    ranks' copies (device pointers valid in THIS process, offset to this rank's first row) for
    gatk_logits_pack_push, `barrier()` a device-side barrier over the group on the current stream."""

    def __init__(self, plan: "ShardPlan", shape, dev):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem
        self.shape = shape
        n, pitch = shape
        self.flat = symm_mem.empty(n * pitch, dtype=torch.float32, device=dev)
        self.hdl = symm_mem.rendezvous(self.flat, plan.group if plan.group is not None else dist.group.WORLD)
        assert self.hdl.world_size == plan.world and self.hdl.rank == plan.rank
        self.buf = self.flat.view(n, pitch)
        peers = [int(self.hdl.buffer_ptrs[q]) + 4 * plan.lo * pitch for q in range(plan.world) if q != plan.rank]
        self.n_peers = len(peers)
        self.ptrs = (ctypes.c_void_p * max(self.n_peers, 1))(*peers)
        self.key = None   # identity of the input whose x columns every copy currently holds
        self._chan = 0

    def barrier(self):
        self.hdl.barrier(channel=self._chan)
        self._chan ^= 1


def _uneven_ok(group) -> bool:
    return dist.get_backend(group) == "nccl"


def gather_rows(full: torch.Tensor, plan: ShardPlan):
    """In-place all-gather of row shards: on entry rank r has filled full[b_r:b_{r+1}]."""
    if plan.world == 1:
        return
    if _uneven_ok(plan.group):
        dist.all_gather([plan.rows(full, r) for r in range(plan.world)], plan.rows(full), group=plan.group)
    else:  # gloo (CPU tests): all_gather needs equal sizes there
        for r in range(plan.world):
            dist.broadcast(plan.rows(full, r), src=dist.get_global_rank(plan.group, r) if plan.group else r,
                           group=plan.group)


def reduce_rows(partial: torch.Tensor, plan: ShardPlan) -> torch.Tensor:
    """Reduce-scatter of row shards: returns the sum over ranks of partial[b_r:b_{r+1}] on rank r."""
    if plan.world == 1:
        return plan.rows(partial)
    if _uneven_ok(plan.group):
        out = _mem.empty_like(plan.rows(partial))
        dist.reduce_scatter(out, [plan.rows(partial, r) for r in range(plan.world)], group=plan.group)
        return out
    for r in range(plan.world):
        dist.reduce(plan.rows(partial, r), dst=dist.get_global_rank(plan.group, r) if plan.group else r,
                    group=plan.group)
    return plan.rows(partial)


def reduce_rows_async(partial: torch.Tensor, plan: ShardPlan):
    """reduce_rows without blocking the launching stream: returns (owned rows, work handle or None).  The
    caller calls work.wait() (a stream-side wait, no host sync) before it reads the rows."""
    if plan.world == 1 or not _uneven_ok(plan.group):
        return reduce_rows(partial, plan), None
    out = _mem.empty_like(plan.rows(partial))
    work = dist.reduce_scatter(out, [plan.rows(partial, r) for r in range(plan.world)], group=plan.group, async_op=True)
    return out, work


def allreduce_(tensors: Sequence[torch.Tensor], plan_or_group=None):
    """Sum small tensors over ranks with one flat all-reduce."""
    group = plan_or_group.group if isinstance(plan_or_group, ShardPlan) else plan_or_group
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, group=group)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()


class ShardedGatLayerFunction(torch.autograd.Function):
    """GatLayerFunction for a destination-row shard (no dropout: the sharded shapes train with p = 0).

    What crosses NVLink per layer:
      forward   all-gather of the INPUT rows x [N, F] (F << H*D) and of g [N, H]; every rank then
                projects all N rows itself -- re-computing Wh costs one GEMM pass, all-gathering it
                would move H*D/F times more bytes than gathering x;
      backward  reduce-scatter of dg [N, H] and all-reduce of dW / da.  When the layer input needs no
                gradient (first layer), dW = x_full^T dWh_partial is formed from the PARTIAL dWh rows on
                every rank and only the F x H*D result is all-reduced; otherwise the partial dWh rows
                are reduce-scattered to their owners (dx needs the complete rows)."""

    @staticmethod
    def forward(ctx, x, w_ext, a_src, a_dst, graph: Graph, plan: ShardPlan, H: int, Dp: int, has_skip: bool,
                alpha: float, act_elu: bool):
        dev = x.device
        n, f_in = x.shape
        assert n == plan.n_local == graph.n_dst and graph.n_src == plan.n_total
        HD = H * Dp
        N = plan.n_total
        x, w_ext, a_src, a_dst = x.contiguous(), w_ext.contiguous(), a_src.contiguous(), a_dst.contiguous()
        M_out = w_ext.shape[1]
        st = _stream()
        if plan.world > 1:
            x_full = _mem.empty(N, f_in, dtype=torch.float32, device=dev)
            plan.rows(x_full).copy_(x)
            gather_rows(x_full, plan)
        else:
            x_full = x
        wh_full = _mem.empty(N, HD, dtype=torch.float32, device=dev)
        _gemm(0, 0, N, HD, f_in, x_full, f_in, w_ext, M_out, wh_full, HD)
        skipv = None
        if has_skip:
            skipv = _mem.empty(n, HD, dtype=torch.float32, device=dev)
            _gemm(0, 0, n, HD, f_in, x, f_in, w_ext, M_out, skipv, HD, b_off=HD)
        f = _mem.empty(n, H, dtype=torch.float32, device=dev)
        g_full = _mem.empty(N, H, dtype=torch.float32, device=dev)
        _lib.call("gatk_logits_fwd", n, H, Dp, plan.rows(wh_full).data_ptr(), HD, None, 1.0, a_src.data_ptr(),
                  a_dst.data_ptr(), f.data_ptr(), plan.rows(g_full).data_ptr(), 0, 0, 0.0, st)
        gather_rows(g_full, plan)

        need_grad = any(ctx.needs_input_grad[:4])
        out = _mem.empty(n, HD, dtype=torch.float32, device=dev)
        separate_hagg = need_grad and (has_skip or act_elu)
        hagg = _mem.empty(n, HD, dtype=torch.float32, device=dev) if separate_hagg else None
        lse = _mem.empty(n, H, dtype=torch.float32, device=dev) if need_grad else None
        hubs = graph.hubs
        scratch = _hub_scratch(0, H, Dp, hubs.n_seg, dev)
        _lib.call("gatk_attn_fwd", n, graph.rowptr.data_ptr(), _ptr(graph.col), H, Dp, wh_full.data_ptr(), HD,
                  f.data_ptr(), H, g_full.data_ptr(), H, None, 1.0, float(alpha), _ptr(skipv), HD, int(act_elu),
                  _ptr(hagg), out.data_ptr(), HD, _ptr(lse), *hubs.args(scratch), graph.counter.data_ptr(), *hubs.item_args(), 0, 0, 0.0, st)
        if need_grad:
            ctx.graph, ctx.plan = graph, plan
            ctx.cfg = (H, Dp, has_skip, float(alpha), bool(act_elu))
            ctx.save_for_backward(x, x_full, w_ext, a_src, a_dst, wh_full, g_full, f, lse, out,
                                  hagg if separate_hagg else out)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, x_full, w_ext, a_src, a_dst, wh_full, g_full, f, lse, out, hagg = ctx.saved_tensors
        graph, plan = ctx.graph, ctx.plan
        H, Dp, has_skip, alpha, act_elu = ctx.cfg
        dev = x.device
        n, f_in = x.shape
        HD = H * Dp
        M_out = w_ext.shape[1]
        N = plan.n_total
        st = _stream()
        gout = gout.contiguous()
        need_dx = ctx.needs_input_grad[0]
        tptr, trow, perm, thubs = graph.transpose()

        ldrec = _lib.query("gatk_attn_bwd_record_ld", H, Dp)
        rec = _mem.empty(n, ldrec, dtype=torch.float32, device=dev)
        dskip = _mem.empty(n, HD, dtype=torch.float32, device=dev) if has_skip else None
        _lib.call("gatk_attn_bwd_prep", n, H, Dp, gout.data_ptr(), HD, out.data_ptr() if (act_elu and has_skip) else None, HD,
                  int(act_elu), hagg.data_ptr(), HD, f.data_ptr(), H, lse.data_ptr(), rec.data_ptr(), ldrec,
                  _ptr(dskip), HD, st)

        # partial dWh / dg for EVERY source from this rank's destination rows
        dwh_part = _mem.empty(N, HD, dtype=torch.float32, device=dev)
        dg_part = _mem.empty(N, H, dtype=torch.float32, device=dev)
        edge_dz = _mem.empty(graph.nnz, H, dtype=torch.float32, device=dev)
        scratch_t = _hub_scratch(1, H, Dp, thubs.n_seg, dev)
        _lib.call("gatk_attn_bwd_fused", N, tptr.data_ptr(), _ptr(trow), _ptr(perm), H, Dp, wh_full.data_ptr(), HD,
                  g_full.data_ptr(), H, rec.data_ptr(), ldrec, None, 1.0, alpha,
                  a_dst.data_ptr(), dwh_part.data_ptr(), HD, dg_part.data_ptr(), H, edge_dz.data_ptr(), None, 0,
                  *thubs.args(scratch_t), graph.counter.data_ptr(), *thubs.item_args(), 0, 0, 0.0, st)
        del rec
        dg_loc = reduce_rows(dg_part, plan).contiguous()
        # owned rows: df = segmented sum of dz, dWh_i += df_i a_src (added once, on the owner's partial rows)
        df = _mem.empty(n, H, dtype=torch.float32, device=dev)
        hubs = graph.hubs
        scratch = _hub_scratch(2, H, Dp, hubs.n_seg, dev)
        dwh_own = plan.rows(dwh_part)
        _lib.call("gatk_attn_bwd_finish", n, graph.rowptr.data_ptr(), H, Dp, edge_dz.data_ptr(), a_src.data_ptr(),
                  None, 1.0, dwh_own.data_ptr(), HD, df.data_ptr(), H, *hubs.args(scratch), 0, 0, 0.0, st)
        del edge_dz

        da_src = _mem.empty(H, Dp, dtype=torch.float32, device=dev)
        da_dst = _mem.empty(H, Dp, dtype=torch.float32, device=dev)
        ws = _mem.empty(_lib.query("gatk_da_workspace_floats", H, Dp), dtype=torch.float32, device=dev)
        _lib.call("gatk_da_reduce", n, H, Dp, plan.rows(wh_full).data_ptr(), HD, df.data_ptr(), dg_loc.data_ptr(),
                  da_src.data_ptr(), da_dst.data_ptr(), ws.data_ptr(), st)

        dw_ext = _mem.empty(f_in, M_out, dtype=torch.float32, device=dev)
        dx = None
        if not need_dx:
            # dW = sum over ranks of x_full^T dWh_partial: no row exchange, only the F x H*D all-reduce below
            _gemm(1, 0, f_in, HD, N, x_full, f_in, dwh_part, HD, dw_ext, M_out)
            if has_skip:
                _gemm(1, 0, f_in, HD, n, x, f_in, dskip, HD, dw_ext, M_out, c_off=HD)
        else:
            dwh_loc = reduce_rows(dwh_part, plan)
            _gemm(1, 0, f_in, HD, n, x, f_in, dwh_loc, HD, dw_ext, M_out)
            dx = _mem.empty(n, f_in, dtype=torch.float32, device=dev)
            _gemm(0, 1, n, f_in, HD, dwh_loc, HD, w_ext, M_out, dx, f_in)
            if has_skip:
                _gemm(1, 0, f_in, HD, n, x, f_in, dskip, HD, dw_ext, M_out, c_off=HD)
                _gemm(0, 1, n, f_in, HD, dskip, HD, w_ext, M_out, dx, f_in, accumulate=1, b_off=HD)
        if plan.world > 1:
            allreduce_([dw_ext, da_src, da_dst], plan)
        return dx, dw_ext, da_src, da_dst, None, None, None, None, None, None, None


def gather_rows_async(full: torch.Tensor, plan: ShardPlan):
    """gather_rows without blocking the launching stream: returns a work handle (None when nothing is in flight);
    work.wait() is a stream-side wait."""
    if plan.world == 1:
        return None
    if _uneven_ok(plan.group):
        return dist.all_gather([plan.rows(full, r) for r in range(plan.world)], plan.rows(full), group=plan.group, async_op=True)
    gather_rows(full, plan)
    return None


def head_chunks(H: int, world: int = 8) -> int:
    """Heads per exchange chunk of the hidden-layer form.  Chunking lets chunk k+1's rows cross NVLink while chunk k's
    attention runs, but the attention kernels lose efficiency on narrower rows (measured at the products shape on
    2 GPUs, 8 x 64: one chunk 60.6 ms, two 62.4, four 64.6 per step -- the exposed exchange is only ~14 ms there).  So:
    one chunk on up to two ranks, two beyond (the exchange grows with the rank count while the compute per rank
    shrinks); GATK_SHARD_CHUNKS overrides the count."""
    want = int(os.environ.get("GATK_SHARD_CHUNKS", "1" if world <= 2 else "2"))
    c = max(1, min(want, H))
    while H % c:
        c -= 1
    return H // c


class ShardedGatLayerWhFunction(torch.autograd.Function):
    """A HIDDEN layer on a destination-row shard (wide input that needs a gradient; no dropout): every rank projects
    only ITS OWN rows and the projected rows cross NVLink, in head chunks that overlap with the attention kernels
    (north_star: "all-gather of Wh/g and reduce-scatter of dWh, overlapped with local-edge compute").

    One projection of the own rows, Z = x_own [W_0 | W_0 a_dst | .. | W_{C-1} | W_{C-1} a_dst | S | W a_src] (the
    logits are linear in the input, as in functional.GatLayerFoldedFunction).  Per chunk c of Hc heads (Pc = Hc*Dp +
    4*ceil(Hc/4) columns [Wh_c | g_c]):
      forward   own rows copied into the chunk's [N, Pc] buffer -> async all-gather -> K2 on the chunk's heads as soon
                as ITS rows have landed, while the next chunks are still in flight;
      backward  K3/K4 on the chunk give the partial [dWh_c | dg_c] of EVERY source from this rank's destination rows
                -> async reduce-scatter to the owners, overlapped with the next chunk's kernels.
    Then dW = x_own^T dZ_own and dx = dZ_own W^T on own rows (one product each) and one all-reduce of the
    parameter-sized gradients.  The previous version all-gathered the INPUT rows and projected all N rows on every
    rank (G-fold redundant GEMM work) and blocked on both reduce-scatters."""

    @staticmethod
    def forward(ctx, x, w_all, graph: Graph, plan: ShardPlan, H: int, Dp: int, Hc: int, has_skip: bool, alpha: float,
                act_elu: bool, graph_t: "Optional[SourceShard]" = None):
        dev = x.device
        n, f_in = x.shape
        assert n == plan.n_local == graph.n_dst and graph.n_src == plan.n_total
        N, HD, C = plan.n_total, H * Dp, H // Hc
        HDc = Hc * Dp
        Pc = HDc + 4 * ((Hc + 3) // 4)
        Mz = w_all.shape[1]                    # C*Pc | S (HD, if skip) | W a_src (H) | pad
        off_s = C * Pc
        off_f = off_s + (HD if has_skip else 0)
        assert Mz >= off_f + H and Mz % 4 == 0
        x, w_all = x.contiguous(), w_all.contiguous()
        st = _stream()
        z = _mem.empty(n, Mz, dtype=torch.float32, device=dev)
        _gemm(0, 0, n, Mz, f_in, x, f_in, w_all, Mz, z, Mz, label="gemm:project_own")
        whg, works = [], []
        for c in range(C):
            buf = _mem.empty(N, Pc, dtype=torch.float32, device=dev)
            plan.rows(buf).copy_(z[:, c * Pc:(c + 1) * Pc])
            with _lib.timed("comm:allgather_whg_issue"):
                works.append(gather_rows_async(buf, plan))
            whg.append(buf)
        need_grad = any(ctx.needs_input_grad[:2])
        out = _mem.empty(n, HD, dtype=torch.float32, device=dev)
        separate_hagg = need_grad and (has_skip or act_elu)
        haggs = [(_mem.empty(n, HDc, dtype=torch.float32, device=dev) if separate_hagg else None) for _ in range(C)]
        lses = [(_mem.empty(n, Hc, dtype=torch.float32, device=dev) if need_grad else None) for _ in range(C)]
        hubs = graph.hubs
        for c in range(C):
            if works[c] is not None:
                with _lib.timed("comm:allgather_whg_wait"):
                    works[c].wait()
            scratch = _hub_scratch(0, Hc, Dp, hubs.n_seg, dev)
            _lib.call("gatk_attn_fwd", n, graph.rowptr.data_ptr(), _ptr(graph.col), Hc, Dp, whg[c].data_ptr(), Pc,
                      z.data_ptr() + 4 * (off_f + c * Hc), Mz, whg[c].data_ptr() + 4 * HDc, Pc, None, 1.0, float(alpha),
                      z.data_ptr() + 4 * (off_s + c * HDc) if has_skip else None, Mz, int(act_elu), _ptr(haggs[c]),
                      out.data_ptr() + 4 * c * HDc, HD, _ptr(lses[c]), *hubs.args(scratch), graph.counter.data_ptr(),
                      *hubs.item_args(), 0, 0, 0.0, st)
        if need_grad:
            ctx.graph, ctx.plan, ctx.graph_t = graph, plan, graph_t
            ctx.cfg = (H, Dp, Hc, has_skip, float(alpha), bool(act_elu), separate_hagg)
            ctx.save_for_backward(x, w_all, z, out, *whg, *[h for h in haggs if h is not None], *lses)
        return out

    @staticmethod
    def backward(ctx, gout):
        H, Dp, Hc, has_skip, alpha, act_elu, separate_hagg = ctx.cfg
        C = H // Hc
        saved = ctx.saved_tensors
        x, w_all, z, out = saved[:4]
        whg = saved[4:4 + C]
        haggs = saved[4 + C:4 + 2 * C] if separate_hagg else None
        lses = saved[-C:]
        graph, plan = ctx.graph, ctx.plan
        dev = x.device
        n, f_in = x.shape
        N, HD = plan.n_total, H * Dp
        HDc = Hc * Dp
        Pc = HDc + 4 * ((Hc + 3) // 4)
        Mz = w_all.shape[1]
        off_s = C * Pc
        off_f = off_s + (HD if has_skip else 0)
        st = _stream()
        gout = gout.contiguous()
        hubs = graph.hubs
        dz = _mem.empty(n, Mz, dtype=torch.float32, device=dev)   # [dWh_c | dg_c | pad].. | dSkip | df | pad
        if Mz > off_f + H:
            dz[:, off_f + H:].zero_()
        ldrec = _lib.query("gatk_attn_bwd_record_ld", Hc, Dp)
        gt = ctx.graph_t

        def prep(c, rec_ptr):
            hagg_c = haggs[c] if separate_hagg else None
            # without a separate hagg (no skip, no ELU) the layer output IS the aggregation
            _lib.call("gatk_attn_bwd_prep", n, Hc, Dp, gout.data_ptr() + 4 * c * HDc, HD,
                      out.data_ptr() + 4 * c * HDc if (act_elu and has_skip) else None, HD, int(act_elu),
                      hagg_c.data_ptr() if hagg_c is not None else out.data_ptr() + 4 * c * HDc, HDc if hagg_c is not None else HD,
                      z.data_ptr() + 4 * (off_f + c * Hc), Mz, lses[c].data_ptr(), rec_ptr, ldrec,
                      dz.data_ptr() + 4 * (off_s + c * HDc) if has_skip else None, Mz, st)

        if gt is not None:
            # ---- source-row shards: all-gather the destination records [dh' | f, lse, c], then every rank COMPLETES
            # dWh_j / dg_j of its own sources (work and traffic shrink with the rank count); the per-destination
            # df partials are accumulated with reds and reduce-scattered ([N, Hc] floats: small)
            recs, works = [], []
            for c in range(C):
                rec_full = _mem.empty(N, ldrec, dtype=torch.float32, device=dev)
                prep(c, rec_full.data_ptr() + 4 * plan.lo * ldrec)
                with _lib.timed("comm:allgather_rec_issue"):
                    works.append(gather_rows_async(rec_full, plan))
                recs.append(rec_full)
            df_owned, df_works = [], []
            for c in range(C):
                if works[c] is not None:
                    with _lib.timed("comm:allgather_rec_wait"):
                        works[c].wait()
                if Pc > HDc + Hc:
                    dz[:, c * Pc + HDc + Hc:(c + 1) * Pc].zero_()
                df_part = torch.zeros(N, Hc, dtype=torch.float32, device=dev)
                scratch_t = _hub_scratch(1, Hc, Dp, gt.hubs.n_seg, dev)
                own_whg = whg[c].data_ptr() + 4 * plan.lo * Pc
                _lib.call("gatk_attn_bwd_fused", n, gt.tptr.data_ptr(), _ptr(gt.trow), None, Hc, Dp, own_whg, Pc,
                          own_whg + 4 * HDc, Pc, recs[c].data_ptr(), ldrec, None, 1.0, alpha, None,
                          dz.data_ptr() + 4 * c * Pc, Mz, dz.data_ptr() + 4 * (c * Pc + HDc), Mz, None, df_part.data_ptr(), Hc,
                          *gt.hubs.args(scratch_t), gt.counter.data_ptr(), *gt.hubs.item_args(), 0, 0, 0.0, st)
                with _lib.timed("comm:reduce_df_issue"):
                    own, work = reduce_rows_async(df_part, plan)
                df_owned.append(own)
                df_works.append(work)
            for c in range(C):
                if df_works[c] is not None:
                    with _lib.timed("comm:reduce_df_wait"):
                        df_works[c].wait()
                dz[:, off_f + c * Hc:off_f + (c + 1) * Hc].copy_(df_owned[c])
            del recs
        else:
            tptr, trow, perm, thubs = graph.transpose()
            owned, works = [], []
            for c in range(C):
                rec = _mem.empty(n, ldrec, dtype=torch.float32, device=dev)
                prep(c, rec.data_ptr())
                part = _mem.empty(N, Pc, dtype=torch.float32, device=dev)
                if Pc > HDc + Hc:
                    part[:, HDc + Hc:].zero_()   # pad columns behind dg meet zero weight columns in the products below
                edge_dz = _mem.empty(graph.nnz, Hc, dtype=torch.float32, device=dev)
                scratch_t = _hub_scratch(1, Hc, Dp, thubs.n_seg, dev)
                _lib.call("gatk_attn_bwd_fused", N, tptr.data_ptr(), _ptr(trow), _ptr(perm), Hc, Dp, whg[c].data_ptr(), Pc,
                          whg[c].data_ptr() + 4 * HDc, Pc, rec.data_ptr(), ldrec, None, 1.0, alpha, None,
                          part.data_ptr(), Pc, part.data_ptr() + 4 * HDc, Pc, edge_dz.data_ptr(), None, 0,
                          *thubs.args(scratch_t), graph.counter.data_ptr(), *thubs.item_args(), 0, 0, 0.0, st)
                with _lib.timed("comm:reduce_dwh_issue"):
                    own, work = reduce_rows_async(part, plan)
                owned.append(own)
                works.append(work)
                scratch = _hub_scratch(2, Hc, Dp, hubs.n_seg, dev)
                _lib.call("gatk_attn_bwd_finish", n, graph.rowptr.data_ptr(), Hc, Dp, edge_dz.data_ptr(), None, None, 1.0, None, 0,
                          dz.data_ptr() + 4 * (off_f + c * Hc), Mz, *hubs.args(scratch), 0, 0, 0.0, st)
                del rec, edge_dz
            for c in range(C):
                if works[c] is not None:
                    with _lib.timed("comm:reduce_dwh_wait"):
                        works[c].wait()
                dz[:, c * Pc:(c + 1) * Pc].copy_(owned[c])
        # own-row products (one each, as on a single GPU)
        dw_all = _mem.empty(f_in, Mz, dtype=torch.float32, device=dev)
        _gemm(1, 0, f_in, Mz, n, x, f_in, dz, Mz, dw_all, Mz, label="gemm:dW_own")
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _mem.empty(n, f_in, dtype=torch.float32, device=dev)
            _gemm(0, 1, n, f_in, Mz, dz, Mz, w_all, Mz, dx, f_in, label="gemm:dx_own")
        if plan.world > 1:
            with _lib.timed("comm:allreduce_dw"):
                allreduce_([dw_all], plan)
        return dx, dw_all, None, None, None, None, None, None, None, None, None


class ShardedGatLayerAggFirstFunction(torch.autograd.Function):
    """functional.GatLayerAggFirstFunction on a destination-row shard (narrow first-layer inputs, no dropout).

    What crosses NVLink per layer:
      forward   the GATHER ROWS xg = [x | g] (F_in + H floats per node, 64-byte aligned pitch): every rank packs its
                own rows and the pack kernel itself writes them into the other ranks' copies (gatk_logits_pack_push
                over symmetric memory; NCCL all-gather when that is unavailable).  While the input tensor is unchanged
                only the g columns are exchanged; nothing is recomputed;
      backward  reduce-scatter of the per-source logit gradients dg [N, H] and one all-reduce of the
                parameter-sized gradients (dW, d[W a_src | W a_dst]).
    Everything H*D wide (aggregated rows, outputs, their gradients) stays on the rank that owns the rows."""

    @staticmethod
    def forward(ctx, x, w_ext, w_uv, graph: Graph, plan: ShardPlan, H: int, Dp: int, has_skip: bool, alpha: float,
                act_elu: bool, x_key=None):
        dev = x.device
        n, f_in = x.shape
        assert n == plan.n_local == graph.n_dst and graph.n_src == plan.n_total
        N = plan.n_total
        HD = H * Dp
        M_out = HD * (2 if has_skip else 1)
        Muv = w_uv.shape[1]
        Fp = (f_in + 3) // 4 * 4
        x, w_ext, w_uv = x.contiguous(), w_ext.contiguous(), w_uv.contiguous()
        st = _stream()
        P = _lib.query("gatk_xg_pitch", Fp, H)
        # The x columns of the gathered rows depend only on the layer INPUT.  A first layer's input is the
        # dataset's feature matrix: constant across training steps, so its all-gather is done once and kept
        # (keyed on the input tensor's storage and version counter: any in-place update or a new tensor
        # gathers again).  Per step only the source logits g [N, H] cross NVLink.
        f = _mem.empty(n, H, dtype=torch.float32, device=dev)
        peer = plan.peer_rows((N, P), dev) if plan.world > 1 else None
        persistent = peer is not None or x_key is not None
        if persistent:
            # the gathered rows live in a plan-wide buffer that the NEXT forward rewrites through raw pointers (no
            # autograd version counter sees that): every forward takes a new generation and backward checks it
            plan.xg_generation = getattr(plan, "xg_generation", 0) + 1
        if peer is not None:
            # fused pack + exchange: the pack kernel writes every row (or, when the peers already hold this
            # input's x columns, just its g columns) into all the other GPUs' copies over NVLink; two device-side
            # barriers order it against the readers of the previous contents and of the new ones
            hit = _same_input(peer.key, x_key)
            xg_full = peer.buf
            xg_loc = plan.rows(xg_full)
            with _lib.timed("comm:pack_push"):
                peer.barrier()
                _lib.call("gatk_logits_pack_push", n, f_in, H, x.data_ptr(), f_in, w_uv.data_ptr(), Muv, xg_loc.data_ptr(), P,
                          f.data_ptr(), H, peer.n_peers, peer.ptrs, 0 if hit else 1, st, label="gatk_logits_pack")
                peer.barrier()
            peer.key = x_key
        else:
            cached = plan.cached_xg(x_key, (N, P), dev) if x_key is not None else None
            hit = cached is not None and cached[1]
            xg_full = cached[0] if cached is not None else _mem.empty(N, P, dtype=torch.float32, device=dev)
            xg_loc = plan.rows(xg_full)
            _lib.call("gatk_logits_pack", n, f_in, H, x.data_ptr(), f_in, w_uv.data_ptr(), Muv, xg_loc.data_ptr(), P,
                      f.data_ptr(), H, st)
            if hit and plan.world > 1:
                with _lib.timed("comm:allgather_g"):
                    g_all = _mem.empty(N, H, dtype=torch.float32, device=dev)
                    plan.rows(g_all).copy_(xg_loc[:, Fp:Fp + H])
                    gather_rows(g_all, plan)
                    xg_full[:, Fp:Fp + H].copy_(g_all)
            else:
                with _lib.timed("comm:allgather_xg"):
                    gather_rows(xg_full, plan)
        need_grad = any(ctx.needs_input_grad[1:3])
        xagg = _mem.empty(n, H * Fp, dtype=torch.float32, device=dev)
        lse = _mem.empty(n, H, dtype=torch.float32, device=dev) if need_grad else None
        hubs = graph.hubs
        scratch = _x_scratch(0, H, Fp, hubs.n_seg, dev)
        _lib.call("gatk_attn_x_fwd", graph.n_src, n, graph.rowptr.data_ptr(), _ptr(graph.col), H, Fp, xg_full.data_ptr(), P,
                  f.data_ptr(), H, float(alpha), xagg.data_ptr(), H * Fp, _ptr(lse),
                  *hubs.args(scratch), graph.counter.data_ptr(), *hubs.item_args(), st)
        out = _mem.empty(n, HD, dtype=torch.float32, device=dev)
        fuse_elu = act_elu and not has_skip
        _gemm_batched(0, 0, n, Dp, f_in, H, xagg, H * Fp, Fp, w_ext, M_out, Dp, out, HD, Dp, epilogue=int(fuse_elu),
                      label="gemm:project")
        if has_skip:
            _gemm(0, 0, n, HD, f_in, xg_loc, P, w_ext, M_out, out, HD, accumulate=1, b_off=HD)
            if act_elu:
                _lib.call("gatk_elu_fwd", n, HD, out.data_ptr(), HD, st)
        if need_grad:
            ctx.graph, ctx.plan = graph, plan
            ctx.cfg = (H, Dp, has_skip, float(alpha), bool(act_elu), f_in, Fp, Muv)
            ctx.xg_generation = plan.xg_generation if persistent else None
            ctx.save_for_backward(xg_full, w_ext, f, lse, xagg, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        xg_full, w_ext, f, lse, xagg, out = ctx.saved_tensors
        graph, plan = ctx.graph, ctx.plan
        if ctx.xg_generation is not None and ctx.xg_generation != plan.xg_generation:
            raise RuntimeError(
                "pygat_b200.sharded: the gathered rows saved by this forward were overwritten by a later forward on the "
                "same ShardPlan (generation %d, now %d).  Run backward before the next forward of this layer, or give "
                "the second call its own ShardPlan / pass cache_input_gather=False with GATK_PEER_PUSH=0."
                % (ctx.xg_generation, plan.xg_generation))
        H, Dp, has_skip, alpha, act_elu, f_in, Fp, Muv = ctx.cfg
        dev = xg_full.device
        N, P = xg_full.shape
        n = plan.n_local
        HD = H * Dp
        M_out = HD * (2 if has_skip else 1)
        st = _stream()
        gout = gout.contiguous()
        xg_loc = plan.rows(xg_full)
        fuse = act_elu and not has_skip and Fn.FUSE_ELU_GRAD and Fn._elu_grad_fusable(n, f_in, H, Dp, Fp, M_out, HD)
        if act_elu and not fuse:
            dhp = _mem.empty(n, HD, dtype=torch.float32, device=dev)
            _lib.call("gatk_elu_bwd", n, HD, gout.data_ptr(), HD, out.data_ptr(), HD, dhp.data_ptr(), HD, st)
        else:
            dhp = gout
        eo, ldeo = (out, HD) if fuse else (None, 0)
        dw_ext = _mem.empty(f_in, M_out, dtype=torch.float32, device=dev)
        dxagg = (_mem.empty if Fp == f_in else torch.zeros)(n, H * Fp, dtype=torch.float32, device=dev)
        _gemm_batched(0, 1, n, f_in, Dp, H, dhp, HD, Dp, w_ext, M_out, Dp, dxagg, H * Fp, Fp, label="gemm:dxagg",
                      elu_out=eo, ld_elu=ldeo)
        dfg = (_mem.empty if Muv == 2 * H else torch.zeros)(n, Muv, dtype=torch.float32, device=dev)
        hubs = graph.hubs
        scratch = _x_scratch(1, H, Fp, hubs.n_seg, dev)
        # dg: partial sums for EVERY source from this rank's stored entries, reduced to the rows' owners
        if Fn.DETERMINISTIC:
            ds = _mem.empty(graph.nnz, H, dtype=torch.float32, device=dev)
            _lib.call("gatk_attn_x_bwd", graph.n_src, n, graph.rowptr.data_ptr(), _ptr(graph.col), H, Fp, xg_full.data_ptr(), P,
                      f.data_ptr(), H, lse.data_ptr(), alpha, xagg.data_ptr(), H * Fp,
                      dxagg.data_ptr(), H * Fp, ds.data_ptr(), None, None, 0, dfg.data_ptr(), Muv,
                      *hubs.args(scratch), graph.counter.data_ptr(), *hubs.item_args(), st)
            tptr, _trow, perm, thubs = graph.transpose()[:4]
            dg_part = _mem.empty(N, H, dtype=torch.float32, device=dev)
            _lib.call("gatk_edge_tsum", N, tptr.data_ptr(), _ptr(perm), H, ds.data_ptr(), dg_part.data_ptr(), H,
                      thubs.seg_len, _ptr(thubs.rows), thubs.n_hub, st)
        else:  # the edge pass accumulates them itself (red.global.add into the zeroed [N, H] array)
            dg_part = torch.zeros(N, H, dtype=torch.float32, device=dev)
            _lib.call("gatk_attn_x_bwd", graph.n_src, n, graph.rowptr.data_ptr(), _ptr(graph.col), H, Fp, xg_full.data_ptr(), P,
                      f.data_ptr(), H, lse.data_ptr(), alpha, xagg.data_ptr(), H * Fp,
                      dxagg.data_ptr(), H * Fp, None, None, dg_part.data_ptr(), H, dfg.data_ptr(), Muv,
                      *hubs.args(scratch), graph.counter.data_ptr(), *hubs.item_args(), st)
        # the reduce-scatter runs on NCCL's stream while the value-path products (which need nothing from it)
        # keep this stream busy
        dg_own, work = reduce_rows_async(dg_part, plan)
        _gemm_batched(1, 0, f_in, Dp, n, H, xagg, H * Fp, Fp, dhp, HD, Dp, dw_ext, M_out, Dp, label="gemm:dW",
                      elu_out=eo, ld_elu=ldeo)
        if has_skip:
            _gemm(1, 0, f_in, HD, n, xg_loc, P, dhp, HD, dw_ext, M_out, c_off=HD)
        with _lib.timed("comm:reduce_dg_wait"):
            if work is not None:
                work.wait()
            dfg[:, H:2 * H] = dg_own
        dw_uv = _mem.empty(f_in, Muv, dtype=torch.float32, device=dev)
        _gemm_batched(1, 0, f_in, Muv, n, 1, xg_loc, P, 0, dfg, Muv, 0, dw_uv, Muv, 0, label="gemm:dlogits")  # one "head": the TMEM-A TN kernel
        if plan.world > 1:
            with _lib.timed("comm:allreduce_dw"):
                allreduce_([dw_ext, dw_uv], plan)
        return None, dw_ext, dw_uv, None, None, None, None, None, None, None, None


def sharded_gat_layer(x_local: torch.Tensor, graph: Graph, plan: ShardPlan, Ws, a_srcs, a_dsts, skips, alpha: float,
                      concat: bool, combine: str = "cat", form: str = "auto", cache_input_gather: bool = True,
                      graph_t: "Optional[SourceShard]" = None) -> torch.Tensor:
    """All heads of one GAT layer on this rank's destination rows (see functional.gat_layer).
    cache_input_gather: keep the all-gathered input rows of the aggregate-first form while the input tensor is
    unchanged (same storage, same version counter): a first layer's features are static across steps."""
    H = len(Ws)
    x_local = x_local.float()
    w_ext, a_src, a_dst, D, Dp = pack_heads(Ws, a_srcs, a_dsts, skips)
    agg_first = (Fn.AGG_FIRST and form in ("auto", "agg_first")
                 and not (x_local.requires_grad and torch.is_grad_enabled())
                 and (agg_first_geometry(x_local.shape[1], H, Dp)[1] or form == "agg_first"))
    if agg_first:
        w3 = w_ext[:, : H * Dp].reshape(x_local.shape[1], H, Dp)
        uv = [(w3 * a_src).sum(-1), (w3 * a_dst).sum(-1)]
        if (-2 * H) % 4:
            uv.append(w_ext.new_zeros(x_local.shape[1], (-2 * H) % 4))
        x_key = _input_key(x_local) if cache_input_gather else None
        rows = ShardedGatLayerAggFirstFunction.apply(x_local, w_ext, torch.cat(uv, dim=1), graph, plan, H, Dp,
                                                     skips is not None, float(alpha), bool(concat), x_key)
    elif form in ("auto", "exchange_wh") and (x_local.requires_grad and torch.is_grad_enabled() or form == "exchange_wh"):
        # hidden layer: project own rows, exchange [Wh | g] in head chunks (ShardedGatLayerWhFunction)
        f_in = x_local.shape[1]
        Hc = head_chunks(H, plan.world)
        C = H // Hc
        HDc, gp = Hc * Dp, 4 * ((Hc + 3) // 4)
        w3 = w_ext[:, : H * Dp].reshape(f_in, H, Dp)
        u, v = (w3 * a_src).sum(-1), (w3 * a_dst).sum(-1)            # f = x (W a_src), g = x (W a_dst)
        blocks = []
        for c in range(C):
            blocks.append(w_ext[:, c * HDc:(c + 1) * HDc])
            blocks.append(v[:, c * Hc:(c + 1) * Hc])
            if gp > Hc:
                blocks.append(w_ext.new_zeros(f_in, gp - Hc))
        if skips is not None:
            blocks.append(w_ext[:, H * Dp:])
        blocks.append(u)
        if (-H) % 4:
            blocks.append(w_ext.new_zeros(f_in, (-H) % 4))
        rows = ShardedGatLayerWhFunction.apply(x_local, torch.cat(blocks, dim=1), graph, plan, H, Dp, Hc, skips is not None,
                                               float(alpha), bool(concat), graph_t)
    else:
        rows = ShardedGatLayerFunction.apply(x_local, w_ext, a_src, a_dst, graph, plan, H, Dp, skips is not None,
                                             float(alpha), bool(concat))
    if combine == "none" or (combine == "cat" and D == Dp):
        return rows
    return HeadCombineFunction.apply(rows, H, D, Dp, 1 if combine == "mean" else 0)


def sharded_gat_forward(model, x_local: torch.Tensor, graph: Graph, plan: ShardPlan,
                        graph_t: "Optional[SourceShard]" = None) -> torch.Tensor:
    """models.GAT.forward (models.py:29-35) on a destination-row shard: every layer runs through
    sharded_gat_layer -- the first one in the aggregate-first form when its input is narrow and needs no gradient,
    the following ones in the hidden-layer form (own-row projection, [Wh | g] exchange, source-shard backward when
    graph_t is given); heads are concatenated, the last layer averaged.  The parameter gradients every layer's
    backward leaves are already summed over the ranks.  Dropout must be inactive (the sharded kernels train the
    large shapes, which use p = 0)."""
    from .layers import _EngineHead
    last = len(model.gat_layers) - 1
    x = x_local
    for i, heads in enumerate(model.gat_layers):
        h0 = heads[0]
        if not all(isinstance(h, _EngineHead) for h in heads):
            raise RuntimeError("sharded_gat_forward: only GraphAttentionLayer / SpGraphAttentionLayer heads are sharded")
        if h0.training and h0.dropout > 0.0:
            raise RuntimeError("sharded_gat_forward: dropout is active; the sharded layers have no dropout path")
        vecs = [h.attention_vectors() for h in heads]
        skips = [h.skip_projection for h in heads] if h0.skip_connection else None
        x = sharded_gat_layer(x, graph, plan, [h.W for h in heads], [v[0] for v in vecs], [v[1] for v in vecs], skips,
                              h0.alpha, h0.concat, combine="mean" if i == last else "cat", graph_t=graph_t)
    return x


# ---------------------------------------------------------------------- graph-level data parallelism (PPI)
def allreduce_gradients(params, n_local_nodes: int, group=None):
    """Make per-rank mean-loss gradients equal the gradient of the mean over ALL ranks' nodes: the
    reference's BCEWithLogitsLoss(reduction='mean') over a merged batch (train_ppi.py:114-119,
    load_data_ppi.py:84-86) weights every node equally, so each rank's gradient is scaled by
    n_local / n_total before the sum.  A rank that has no graphs in this step (the last step of an epoch when
    the batches do not divide evenly, see rank_batch_schedule) passes n_local_nodes = 0 and still takes part:
    its missing gradients count as zeros."""
    params = list(params)
    if not params:
        return
    dev = params[0].device
    # ONE collective and no host synchronisation: every rank contributes n_local * grad and, in the last slot, n_local;
    # the sum of the first part divided by the summed count is the merged-batch gradient.  (The first version
    # all-reduced the count separately and read it back with .item(): a device sync per training step, which
    # serialised the launch-bound PPI step behind the GPU and erased the data-parallel gain.)
    nl = float(n_local_nodes)
    pieces = [(p.grad.reshape(-1) if (p.grad is not None and n_local_nodes > 0)
               else torch.zeros(p.numel(), dtype=p.dtype, device=dev)) for p in params]
    flat = torch.cat(pieces + [torch.ones(1, dtype=params[0].dtype, device=dev)])
    flat.mul_(nl)
    dist.all_reduce(flat, group=group)
    total = flat[-1].clamp_min(1.0)
    flat.div_(total)
    views, off = [], 0
    for p in params:
        views.append(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    missing = [i for i, p in enumerate(params) if p.grad is None]
    for i in missing:
        params[i].grad = torch.empty_like(params[i])
    torch._foreach_copy_([p.grad for p in params], views)


def rank_batch_schedule(n_batches: int, rank: int, world: int):
    """Which batch this rank runs at every step of an epoch under graph-level data parallelism: step s takes
    batches [s*world, (s+1)*world), one per rank, so every rank runs the SAME number of steps (collectives
    stay matched); None marks a step in which this rank idles (and all-reduces zeros with weight 0)."""
    steps = (n_batches + world - 1) // world
    return [s * world + rank if s * world + rank < n_batches else None for s in range(steps)]


def shard_rows_by_cost(rowptr: torch.Tensor, world: int, row_cost: float):
    """synth.shard_rows_by_nnz with a fractional per-row cost (costs scaled to integers)."""
    n = rowptr.numel() - 1
    cum = rowptr * 16 + int(round(row_cost * 16)) * torch.arange(n + 1, device=rowptr.device, dtype=torch.int64)
    total = int(cum[-1].item())
    targets = torch.arange(1, world, device=rowptr.device, dtype=torch.int64) * total // world
    cuts = torch.searchsorted(cum, targets).clamp_(max=n).tolist()
    return [0] + [int(c) for c in cuts] + [n]


def fit_row_cost(rows: torch.Tensor, entries: torch.Tensor, t: torch.Tensor):
    """row_cost = a / b of the least-squares fit t = a * rows + b * entries (None if not identifiable)."""
    A = torch.stack([rows.double(), entries.double()], dim=1)
    if A.shape[0] < 2 or torch.linalg.matrix_rank(A) < 2:
        return None
    sol = torch.linalg.lstsq(A, t.double().unsqueeze(1)).solution.flatten()
    a, b = float(sol[0]), float(sol[1])
    if not (a > 0 and b > 0):
        return None
    return min(max(a / b, 0.0), 1000.0)
