"""Benchmark of the GAT-layer hot path (BASELINE.json metric: GAT layer fwd+bwd head-edges/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload products|pubmed]

A step is one forward + backward of one hidden GAT layer (all heads, ELU, no dropout) over the
whole synthetic graph.  Default workload: the ogbn-products shape BASELINE.json's target is
quoted on (2.45 M nodes, ~61.9 M stored entries incl. self-loops, 100 features, 8 heads x 64).
Rank 0 prints ONE JSON line (see the keys at the bottom).  `--impl reference` times the
reference's sparse CPU path (the oracle port of layers.py:125-173 + its autograd, all host
threads) on a bounded power-law sample of the same shape.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout; stdout carries ONE JSON line
_JSON_OUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle to the real stdout for it and point file descriptor 1
    at stderr, so that whatever a library prints (the NCCL banner still appears with some launchers) cannot precede it."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _JSON_OUT


def emit(line: dict):
    out = _claim_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()

WORKLOADS = {
    # name: nodes, average stored entries per row, in-features, heads, per-head width, zipf exponent
    "products": dict(n=2_449_029, avg_deg=25.26, f_in=100, H=8, D=64, exponent=0.5),
    "pubmed": dict(n=19_717, avg_deg=5.5, f_in=500, H=8, D=8, exponent=0.5),
    # papers100M shape (config 5): n = rows PER RANK (111,059,956 / 8), sources span all ranks' rows; weak scaling
    "papers_shard": dict(n=13_882_495, avg_deg=14.55, f_in=128, H=4, D=32, exponent=0.5, shard_only=True),
    "papers_tiny": dict(n=200_000, avg_deg=14.55, f_in=128, H=4, D=32, exponent=0.5, shard_only=True),
    # the products graph as a HIDDEN layer: the 8 x 64 output of layer 1 is the input, and it needs a gradient
    "products_hidden": dict(n=2_449_029, avg_deg=25.26, f_in=512, H=8, D=64, exponent=0.5, needs_dx=True),
}
# reference backward is O(N^2) memory (layers.py:85): 1 GiB per product at 16384 nodes (the environment
# variable only exists so that the contract test can run the arm in a second)
CPU_SAMPLE_NODES = int(os.environ.get("BENCH_CPU_SAMPLE_NODES", "16384"))


def algorithmic_bytes(n, e, H, D, f_in, need_dx=False, n_long=None):
    """Bytes each kernel's algorithm must move (no-reuse gather model: every stored entry moves its
    neighbour row once per pass; fp32, 4-byte col ids, 8-byte rowptr).

    `layer_survey` is SURVEY.md section 8(d)'s model of the layer (three H*D-wide row gathers per edge: K2,
    a destination-pass K3 and a source-pass K4).  `layer_project_first` sums the kernels of this engine's
    project-first (folded) form, whose backward gathers each row once; `layer_agg_first` sums the kernels of
    the aggregate-first form, which gathers the F_in-wide INPUT row twice."""
    k2 = e * H * (4 * D + 4) + 4 * e + n * H * (4 * D + 12) + 8 * n
    k3_survey = e * H * (4 * D + 4 + 4) + 4 * e + n * H * (8 * D + 16)
    k4_survey = e * H * (4 * D + 4 + 12) + 8 * e + n * H * (4 * D + 4)
    k1 = 4 * n * f_in + 4 * f_in * H * D + 4 * n * H * (D + 2)
    k5 = 4 * n * f_in + 4 * n * H * D + 4 * f_in * H * D + ((4 * n * H * D + 4 * n * f_in) if need_dx else 0)
    prep = n * H * (16 * D + 4)                      # gout, out, hagg read; dhp write; c write
    fused = e * H * (4 * D + 16 + 4) + 8 * e + n * H * (8 * D + 8)   # record gather (dh' + f,lse,c,pad), dz write, trow+perm
    finish = 4 * e * H + 4 * n * H + 8 * n            # dz read, df write (folded form: no dWh update, no da pass)
    # aggregate-first form: gather rows xg = [x (Fp) | g (H)] of pitch P floats
    Fp = (f_in + 3) // 4 * 4
    P = Fp + 4 * ((H + 3) // 4)
    HD = H * D
    pack = n * (4 * f_in + 4 * P + 4 * H)                                        # x read; xg, f written
    x_fwd = e * (4 * P + 4) + n * (4 * H * Fp + 8 * H + 8)                       # xg_j, col; xagg write, f, lse
    # xg_j, col, ds write; dxagg, f, lse, df per row; xagg_i only for rows of more than 32 stored entries (n_long: the
    # others take c_i = sum_j alpha_ij dalpha_ij from their own entries, csrc/attn_x.cu; None = every row, the R1/R2a kernel)
    n_xagg = n if n_long is None else n_long
    x_bwd = e * (4 * P + 4 + 4 * H) + n * (4 * H * Fp + 12 * H + 8) + n_xagg * 4 * H * Fp
    tsum = e * (4 * H + 4) + n * (4 * H + 8)                                     # ds through perm; dg write
    elu_b = 12 * n * HD                                                          # separate ELU' pass (only when not fused)
    g_proj = 4 * n * H * Fp + 4 * n * HD                                         # out_h = ELU(xagg_h W_h)
    # backward products with dh' = gout * ELU'(out) formed inside the kernel: each reads gout AND out (8 n HD)
    g_dw = 4 * n * H * Fp + 8 * n * HD                                           # dW_h = xagg_h^T dh'_h
    g_dx = 8 * n * HD + 4 * n * H * Fp                                           # dxagg_h = dh'_h W_h^T
    g_dl = 4 * n * f_in + 8 * n * H                                              # d[u|v] = x^T [df|dg]
    # x_bwd keeps its 4H bytes per entry: the ds values now travel to L2 as red.global.add operands instead of a store;
    # the transposed segmented sum (tsum) only runs on the bit-reproducible route (GATK_DETERMINISTIC=1)
    agg = pack + x_fwd + x_bwd + g_proj + g_dw + g_dx + g_dl
    return {"gatk_attn_fwd": k2, "gatk_attn_bwd_prep": prep, "gatk_attn_bwd_fused": fused,
            "gatk_attn_bwd_finish": finish, "projection_fwd": k1, "projection_bwd": k5,
            "gatk_logits_pack": pack, "gatk_attn_x_fwd": x_fwd, "gatk_attn_x_bwd": x_bwd, "gatk_edge_tsum": tsum,
            "gatk_elu_bwd": elu_b, "gemm:project": g_proj, "gemm:dW": g_dw, "gemm:dxagg": g_dx, "gemm:dlogits": g_dl,
            "layer_survey": k2 + k3_survey + k4_survey + k1 + k5,
            "layer_project_first": k2 + prep + fused + finish + k1 + k5,
            "layer_agg_first": agg, "layer_agg_first_deterministic": agg + tsum,
            "layer_agg_first_r01": agg + tsum + elu_b - 8 * n * HD}  # round 1: separate ELU' pass, dh' read twice


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ clocks
class NvmlClockSampler:
    """SM clock, power and clock-event reasons read through NVML on a thread, every ~2 ms, so that even the
    ~40 ms timed region of an 8-GPU run holds samples taken under load (`nvidia-smi -lms` needs ~0.5 s to
    deliver its first line).  Raises from __init__ when NVML is not usable; ClockSampler then takes over."""
    PERIOD_S = 0.002

    def __init__(self, dev):
        import pynvml
        import torch
        self.nv = pynvml
        pynvml.nvmlInit()
        handle = None
        uuid = getattr(torch.cuda.get_device_properties(dev), "uuid", None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except pynvml.NVMLError:
                handle = None
        if handle is None:   # no uuid on this torch: the visible-device list, then the plain index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip()]
            idx = int(ids[dev.index]) if ids and all(v.strip().isdigit() for v in ids) else dev.index
            handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.handle = handle
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        self.rows = []
        self.stop_flag = False
        self.thread = None

    def _read(self):
        nv = self.nv
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except (nv.NVMLError, AttributeError):
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)),
                          nv.nvmlDeviceGetPowerUsage(self.handle) / 1e3, int(mask)))

    def _loop(self):
        while not self.stop_flag:
            self._read()
            time.sleep(self.PERIOD_S)

    def start(self):
        import threading
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def sample_now(self):
        """One read from the calling thread (the benchmark takes it with the last step still in flight)."""
        self._read()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=5)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        rows = self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "source": "nvml"}
        reasons = sorted(nm for nm, bit in names.items() if any(r[2] & bit for r in rows))
        return {"sm_mhz": statistics.median(r[0] for r in rows), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(rows), "power_w_max": max(r[1] for r in rows), "source": "nvml"}


class ClockSampler:
    """`nvidia-smi -lms 100` beside the timed region (the profiling recipe's clocks line); used when NVML cannot
    be loaded in-process."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def sample_now(self):
        pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(",") for r in open(self.tmp.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.tmp.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "source": "nvidia-smi"}
        sm = [float(r[1]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({nm for r in rows for nm, v in zip(names, r[5:9]) if "Active" in v and "Not" not in v})
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows), "source": "nvidia-smi"}


def make_clock_sampler(dev):
    try:
        return NvmlClockSampler(dev)
    except Exception as exc:   # NVML missing or refusing: fall back to the nvidia-smi loop
        sys.stderr.write(f"[bench] NVML clock sampling unavailable ({type(exc).__name__}: {exc}); using nvidia-smi -lms\n")
        return ClockSampler(dev.index)


# ------------------------------------------------------------------------------ reference / CPU arm
def _load_true_reference():
    """The UNMODIFIED reference's layers.py from baseline/_ref (copied there by __graft_entry__.build() where
    /root/reference exists; git-ignored, travels with the snapshot), or None.  Its one un-vendored dependency,
    torch_scatter.scatter_max, comes from the shim the golden vectors were generated with."""
    path = os.path.join(ROOT, "baseline", "_ref", "layers.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    shim = os.path.join(ROOT, "tests", "golden", "shims")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    spec = importlib.util.spec_from_file_location("pygat_reference_layers", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_rate(cfg, steps: int, warmup: int, nodes: int = CPU_SAMPLE_NODES):
    """head-edges/s of the reference's sparse CPU path on a bounded sample, all host threads: the reference's own
    SpGraphAttentionLayer modules (layers.py:98-176, dense N x N adjacency in, its dense N x N backward) when
    baseline/_ref is present (kind "reference"), else the oracle port of the same lines (kind "port")."""
    import torch
    from oracle import gat_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H, D, f_in = cfg["H"], cfg["D"], cfg["f_in"]
    rowptr, col = O.power_law_edges(nodes, cfg["avg_deg"], seed=72, exponent=cfg["exponent"])
    adj = O.PatternAdj(rowptr, col)
    e = int(col.numel())
    torch.manual_seed(72)
    x = torch.randn(nodes, f_in)
    gout = torch.randn(nodes, H * D)
    ref = _load_true_reference() if os.environ.get("BENCH_REFERENCE_PORT", "0") == "0" else None
    if ref is not None:
        kind = "reference"
        heads = [ref.SpGraphAttentionLayer(f_in, D, dropout=0.0, alpha=0.2, concat=True) for _ in range(H)]
        edge = adj.nonzero()
        dense = torch.zeros(nodes, nodes)
        dense[edge[:, 0], edge[:, 1]] = 1.0   # what utils.load_data hands the model: a dense N x N float matrix (utils.py:55)

        def step():
            for h in heads:
                h.zero_grad(set_to_none=True)
            torch.cat([h(x, dense) for h in heads], dim=1).backward(gout)   # models.py:32 over the reference's heads
    else:
        kind = "port"
        heads = [O.init_head(f_in, D, "sparse", False) for _ in range(H)]

        def step():
            xs = x.clone().requires_grad_(True)
            ps = [{k: v.clone().requires_grad_(True) for k, v in hp.items()} for hp in heads]
            edge = adj.nonzero().t()  # layers.py:129 (the duck-typed adj makes this free)
            outs = [O.sparse_head(xs, hp["W"], hp["a"], edge, 0.2, True, None, 0.0, faithful=True) for hp in ps]
            torch.cat(outs, dim=1).backward(gout)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    what = ("the unmodified reference's SpGraphAttentionLayer x%d (baseline/_ref/layers.py), dense adj in" % H
            if kind == "reference" else "oracle port")
    return {"value": e * H / dt, "unit": "head-edges/s", "cores": cores, "kind": kind,
            "sample": f"power-law N={nodes} E={e} F={f_in} H={H} D={D}, fwd+bwd incl. the reference's dense "
                      f"N x N backward, {what}, {steps} step(s) of {dt:.2f} s", "ms_per_step": dt * 1e3, "edges": e}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    steps = max(1, min(args.steps, 3))
    r = cpu_reference_rate(cfg, steps=steps, warmup=max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": "gat_layer_fwd_bwd_head_edges_per_s", "value": r["value"],
            "unit": "head-edges/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}-shape power-law sample", "nodes": CPU_SAMPLE_NODES,
                       "edges": r["edges"], "f_in": cfg["f_in"], "heads": cfg["H"], "head_dim": cfg["D"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "head-edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------ our arm
def make_runner(cfg, rank, world, dev):
    from benchmarks.layer import ShardedLayerBench, ShardOnlyLayerBench, SingleGpuLayerBench
    if cfg.get("shard_only"):
        return ShardOnlyLayerBench(cfg, rank, world, dev)
    return ShardedLayerBench(cfg, rank, world, dev) if world > 1 else SingleGpuLayerBench(cfg, dev)


def time_layer(runner, steps, warmup, rank, world, dev, sample_clocks=False):
    """W warm-up steps, then exactly `steps` steps between barrier + synchronize on both sides, timed with CUDA
    events on the launching stream; max over ranks.  Returns ms/step, per-entry-point kernel times, launch counts,
    clocks (rank 0) and every rank's per-call times."""
    import torch
    import torch.distributed as dist

    from pygat_b200 import _lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        runner.step()
    barrier()
    clocks = make_clock_sampler(dev) if (sample_clocks and rank == 0) else None
    _lib.timer = _lib.KernelTimer()
    calls0 = _lib.call_count
    launches0 = _lib.query("gatk_launch_count")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if clocks:
        clocks.start()
    ev0.record()
    for _ in range(steps):
        runner.step()
    ev1.record()
    if clocks:
        clocks.sample_now()   # the host runs ahead of the device: the last steps are still executing
    barrier()
    ms = ev0.elapsed_time(ev1) / steps
    kern = _lib.timer.summary()
    _lib.timer = None
    out = {"launches": _lib.query("gatk_launch_count") - launches0, "calls": _lib.call_count - calls0,
           "clocks": clocks.stop() if clocks else None, "per_rank": None}
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # every rank's per-call times (rank 0's alone hide the load balance: a "comm:" wait is mostly skew)
        mine = {k: round(v["ms_total"] / steps, 4) for k, v in kern.items()}
        mine["rows"], mine["entries"] = runner.plan.n_local, runner.graph.nnz
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        out["per_rank"] = {k: [g.get(k) for g in gathered] for k in mine}
    out["ms"], out["kern"] = ms, kern
    return out


KERNEL_NAMES = ("gatk_attn_fwd", "gatk_attn_bwd_fused", "gatk_attn_bwd_prep", "gatk_attn_bwd_finish",
                "gatk_logits_pack", "gatk_attn_x_fwd", "gatk_attn_x_bwd", "gatk_edge_tsum", "gatk_elu_bwd",
                "gemm:project", "gemm:dW", "gemm:dxagg", "gemm:dlogits")


def kernel_table(kern, ab, world, steps, peak):
    per_kernel = {}
    for name in KERNEL_NAMES:
        if name in kern and name in ab:
            gbs = ab[name] / world / (kern[name]["ms_avg"] * 1e-3) / 1e9
            per_kernel[name] = {"ms": round(kern[name]["ms_avg"], 4), "algorithmic_GB": round(ab[name] / world / 1e9, 3),
                                "achieved_GBs": round(gbs, 1), "frac": round(gbs / peak, 4)}
    other = {k: round(v["ms_total"] / steps, 4) for k, v in kern.items() if k not in per_kernel}
    return per_kernel, other


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = WORKLOADS[args.workload]
    n, H, D, f_in = cfg["n"], cfg["H"], cfg["D"], cfg["f_in"]

    runner = make_runner(cfg, rank, world, dev)
    e_total = runner.e_total
    # rows whose xagg_i the backward edge pass reads: all of them, unless the opt-in short-row variant runs
    n_long = runner.n_long_rows if os.environ.get("GATK_XBWD_SHORT_C") == "1" else None
    res = time_layer(runner, args.steps, args.warmup, rank, world, dev, sample_clocks=True)
    ms, kern = res["ms"], res["kern"]

    # ---- end-to-end through the public API with host buffers (pinned H2D of the step's input
    # features, D2H of the step's results: parameter gradients + a checksum of the output)
    e2e_steps = max(3, min(args.steps, 10)) if not cfg.get("shard_only") else 2
    e2e_ms, h2d, d2h = runner.e2e(e2e_steps)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())

    # ---- N > 1: the sharded step against a single-GPU run of the same step (rank 0 holds the whole graph for it)
    check = None
    if world > 1 and not args.no_check and not cfg.get("shard_only"):
        from benchmarks.layer import parity_check
        try:
            check = parity_check(runner, rank, world, dev)
        except Exception as exc:  # the headline line must still be printed; a failed check is reported, not hidden
            check = {"parity_ok": False, "error": repr(exc)[:300]}
    row_cost = runner.row_cost
    del runner
    torch.cuda.empty_cache()

    # ---- the same layer as a HIDDEN layer (wide input that needs a gradient): the project-first kernels that
    # layer 2 of models.GAT and every PPI layer run (models.py:29-35)
    hidden = None
    if not args.no_hidden and args.workload == "products":  # (the papers shard has no hidden-layer add-on)
        hcfg = WORKLOADS["products_hidden"]
        try:
            hrun = make_runner(hcfg, rank, world, dev)
            hres = time_layer(hrun, max(3, min(args.steps, 5)), 3, rank, world, dev)
            if rank == 0:
                peak, _ = measured_peak()
                hab = algorithmic_bytes(hcfg["n"], hrun.e_total, hcfg["H"], hcfg["D"], hcfg["f_in"], need_dx=True)
                hk, hother = kernel_table(hres["kern"], hab, world, max(3, min(args.steps, 5)), peak)
                hidden = {"workload": "products-shape graph, hidden layer: F_in=512 (needs dx), 8 heads x 64, fwd+bwd",
                          "ms_per_step": round(hres["ms"], 3), "value": hrun.e_total * hcfg["H"] / (hres["ms"] * 1e-3),
                          "unit": "head-edges/s", "form": "project_first",
                          "layer_frac": round(hab["layer_project_first"] / world / (hres["ms"] * 1e-3) / 1e9 / peak, 4),
                          "layer_algorithmic_GB": round(hab["layer_project_first"] / 1e9, 2),
                          "kernels": hk, "other_ms_per_step": hother, "per_rank_ms_per_step": hres["per_rank"]}
                if world > 1:
                    hidden["note"] = ("per-kernel bytes are the single-GPU model / N: every gathered row is counted as "
                                      "moved, so a shard whose hub rows stay in L2 can read above 1.0, and the source-"
                                      "shard backward's prep kernel writes only the destination records (its model does "
                                      "not apply); ms_per_step and layer_frac are the figures to read at N > 1")
            del hrun
            torch.cuda.empty_cache()
        except Exception as exc:
            hidden = {"error": repr(exc)[:300]}

    # ---- epoch workloads: Pubmed (replicas only: one GPU) and PPI (graph-level data parallel over the ranks)
    epochs = None
    if not args.no_epochs and not cfg.get("shard_only"):
        epochs = {}
        try:
            from benchmarks import epochs as epoch_bench
            if world == 1:
                ms_e, info = epoch_bench.pubmed_epoch_ms(dev)
                epochs["pubmed_GAT_sparse_train_plus_eval"] = {"ms_per_epoch": round(ms_e, 3), **info}
                ms_e, info = epoch_bench.pubmed_epoch_sync_free_ms(dev)
                epochs["pubmed_GAT_sparse_fused_loss_head"] = {"ms_per_epoch": round(ms_e, 3), **info}
                ms_e, info = epoch_bench.ppi_epoch_graphed_ms(dev)
                epochs["ppi_GAT_dense_class_fused_head_cuda_graphs"] = {"ms_per_epoch": round(ms_e, 3), **info}
            ms_e, info = epoch_bench.ppi_epoch_ms(dev, rank=rank, world=world)
            if world > 1:
                t = torch.tensor([ms_e], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_e = float(t.item())
            epochs["ppi_GAT_dense_class_train"] = {"ms_per_epoch": round(ms_e, 3), "n_gpus": world, **info}
            ms_e, info = epoch_bench.ppi_epoch_sync_free_ms(dev, rank=rank, world=world)
            if world > 1:
                t = torch.tensor([ms_e], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_e = float(t.item())
            epochs["ppi_GAT_dense_class_fused_head_sync_free"] = {"ms_per_epoch": round(ms_e, 3), "n_gpus": world, **info}
            if args.workload == "products":
                ms_e, info = epoch_bench.products_model_epoch_ms(dev, rank=rank, world=world)
                if world > 1:
                    t = torch.tensor([ms_e], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms_e = float(t.item())
                epochs["products_2layer_GAT_full_batch_epoch"] = {"ms_per_epoch": round(ms_e, 3), **info}
        except Exception as exc:  # the headline metric must still be printed
            epochs["error"] = repr(exc)[:300]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    n_nodes = n * world if cfg.get("shard_only") else n
    ab = algorithmic_bytes(n_nodes, e_total, H, D, f_in, n_long=n_long)
    per_kernel, other = kernel_table(kern, ab, world, args.steps, peak)
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms"]) if per_kernel else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if dom and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {}).get(dom)
    roofline = None
    form = "agg_first" if "gatk_attn_x_fwd" in kern else "project_first"
    ab["layer"] = ab["layer_" + form]
    if dom:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["achieved_GBs"], "peak": peak,
                    "unit": "GB/s", "frac": per_kernel[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                    "form": form,
                    "layer_frac": round(ab["layer"] / world / (ms * 1e-3) / 1e9 / peak, 4),
                    "layer_algorithmic_GB": round(ab["layer"] / 1e9, 2),
                    "layer_frac_survey_model": round(ab["layer_survey"] / world / (ms * 1e-3) / 1e9 / peak, 4),
                    "layer_survey_model_GB": round(ab["layer_survey"] / 1e9, 2)}

    cpu = cpu_reference_rate(cfg, steps=1, warmup=1)
    line = {
        "metric": "gat_layer_fwd_bwd_head_edges_per_s", "value": e_total * H / (ms * 1e-3), "unit": "head-edges/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak" if cfg.get("shard_only") else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}-shape power-law graph, one hidden GAT layer fwd+bwd",
                   "nodes": n_nodes, "edges": e_total, "f_in": f_in, "heads": H, "head_dim": D, "dropout": 0.0,
                   "parallelism": (f"dst-row shards x{world}, {n} rows each, generated per rank (the whole pattern never exists); "
                                   "aggregate-first form; layer-1 features are static, their all-gather is kept across steps; "
                                   "per step g [N,H] is pushed, dg reduce-scattered, dW all-reduced") if cfg.get("shard_only") else
                                  (f"dst-row shards x{world} (cost-balanced, row_cost={row_cost:.1f} entries); layer-1 "
                                   "features are static, their all-gather is kept across steps; per step g [N,H] is "
                                   "all-gathered, dg reduce-scattered, dW all-reduced") if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (Wh alone is %.1f GB)" % (n * H * D * 4 / 1e9)},
        "e2e": {"value": e_total * H / (e2e_ms * 1e-3), "unit": "head-edges/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "input_pipeline": "pinned host -> device copy of step k+1 overlaps step k (double buffered), timed in steady "
                                  "state: `steps` copies and `steps` steps inside the region, which ends when the last copy has landed",
                "result": "D2H = every parameter gradient + an output checksum; the layer output (%.1f GB) stays on the "
                          "device as the next layer's input" % (n * H * D * 4 / 1e9)},
        "gpu_launches": res["launches"], "abi_calls": res["calls"], "clocks": res["clocks"], "roofline": roofline,
        "kernels": per_kernel, "other_ms_per_step": other, "per_rank_ms_per_step": res["per_rank"],
        "check": check, "hidden_layer": hidden,
        "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "epoch_times": epochs,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="products", choices=sorted(WORKLOADS))
    ap.add_argument("--no-epochs", action="store_true", help="skip the Pubmed / PPI epoch-time add-on measurements")
    ap.add_argument("--no-hidden", action="store_true", help="skip the hidden-layer (project-first) add-on measurement")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the comparison with a single-GPU run of the same step")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
