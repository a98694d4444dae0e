"""Host -> device copy rate of the benchmark's feature matrix (2,449,029 x 100 fp32 = 980 MB) from pinned memory:
one copy, the copy split over several streams, the copy beside a bandwidth-bound kernel, and the copy from pinned
memory allocated after binding the thread to the GPU's own CPU set (NUMA placement).  bench.py's `e2e` figure is
bounded by this number whenever 980 MB / rate exceeds the step time."""
import json
import os
import sys

import torch


def timed_copy(dst, src, streams, reps=5):
    n = dst.shape[0]
    k = len(streams)
    cuts = [n * i // k for i in range(k + 1)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        for i, s in enumerate(streams):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                dst[cuts[i]:cuts[i + 1]].copy_(src[cuts[i]:cuts[i + 1]], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    dev = torch.device("cuda", 0)
    out = {"cpu_affinity": sorted(os.sched_getaffinity(0))}
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        out["numa_nodes"] = {d: open(f"/sys/devices/system/node/{d}/cpulist").read().strip() for d in nodes}
    except OSError as exc:
        out["numa_nodes"] = repr(exc)
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    try:
        words = pynvml.nvmlDeviceGetCpuAffinity(h, 8)
        out["gpu_cpu_affinity_words"] = [hex(w) for w in words]
    except pynvml.NVMLError as exc:
        out["gpu_cpu_affinity_words"] = repr(exc)
    try:
        out["pcie"] = {"gen": pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), "width": pynvml.nvmlDeviceGetCurrPcieLinkWidth(h),
                       "max_gen": pynvml.nvmlDeviceGetMaxPcieLinkGeneration(h)}
    except pynvml.NVMLError as exc:
        out["pcie"] = repr(exc)

    n, f = 2449029, 100
    gb = n * f * 4 / 1e9
    dst = torch.empty(n, f, device=dev)
    src = torch.randn(n, f).pin_memory()
    main_s = [torch.cuda.Stream()]
    out["one_stream_GBs"] = round(gb / (timed_copy(dst, src, main_s) * 1e-3), 2)
    for k in (2, 4):
        out[f"{k}_streams_GBs"] = round(gb / (timed_copy(dst, src, [torch.cuda.Stream() for _ in range(k)]) * 1e-3), 2)

    # beside a bandwidth-bound kernel (device copies of 4 GB in a loop on the main stream)
    a = torch.empty(1 << 30, device=dev)
    b = torch.empty(1 << 30, device=dev)
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(40):
        b.copy_(a)
    with torch.cuda.stream(s):
        e0.record(s)
        dst.copy_(src, non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    out["beside_hbm_bound_kernel_GBs"] = round(gb / (e0.elapsed_time(e1) * 1e-3), 2)
    del a, b

    # device -> host, for completeness
    back = torch.empty(n, f).pin_memory()
    torch.cuda.synchronize()
    e0.record()
    back.copy_(dst, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    out["d2h_GBs"] = round(gb / (e0.elapsed_time(e1) * 1e-3), 2)

    # pinned memory first touched by a thread bound to the GPU's CPU set
    try:
        pynvml.nvmlDeviceSetCpuAffinity(h)
        out["affinity_after_bind"] = sorted(os.sched_getaffinity(0))
        src2 = torch.empty(n, f).pin_memory()
        src2.copy_(src)
        out["bound_alloc_one_stream_GBs"] = round(gb / (timed_copy(dst, src2, main_s) * 1e-3), 2)
    except Exception as exc:
        out["bound_alloc_one_stream_GBs"] = repr(exc)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
