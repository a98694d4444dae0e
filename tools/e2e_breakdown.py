"""Where the end-to-end step differs from the resident step: per-entry-point kernel times (CUDA events around each
C-ABI call) of plain steps and of the host-buffer pipeline of benchmarks.layer.pipelined_e2e, one GPU."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from benchmarks.layer import SingleGpuLayerBench  # noqa: E402
from pygat_b200 import _lib  # noqa: E402


def table(fn, steps):
    _lib.timer = _lib.KernelTimer()
    ms = fn(steps)
    kern = _lib.timer.summary()
    _lib.timer = None
    return ms, {k: round(v["ms_total"] / steps, 3) for k, v in kern.items()}


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    r = SingleGpuLayerBench(bench.WORKLOADS["products"], dev)
    for _ in range(3):
        r.step()

    def plain(k):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            r.step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    out = {}
    ms, t = table(plain, 10)
    out["resident"] = {"ms_per_step": round(ms, 3), "kernels": t, "sum": round(sum(t.values()), 3)}
    ms, t = table(lambda k: r.e2e(k)[0], 10)
    out["e2e_with_per_call_events"] = {"ms_per_step": round(ms, 3), "kernels": {k: round(v * 10 / 12, 3) for k, v in t.items()},
                                      "note": "kernel sums cover 12 steps (2 warm-up + 10), scaled to per step"}
    out["e2e_untimed_calls"] = {"ms_per_step": round(r.e2e(10)[0], 3)}
    out["e2e_30_steps"] = {"ms_per_step": round(r.e2e(30)[0], 3)}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
